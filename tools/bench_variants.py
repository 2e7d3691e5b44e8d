"""Tuning aid: time the packed K=8 step (2^20 envs, FP32) for every library variant given on the command line.
Each variant runs in its own process (AH_B200_LIB selects the library).  python tools/bench_variants.py lib1.so lib2.so ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch, json
sys.path.insert(0, %r)
from air_hockey_training_environment_b200 import BatchedAirHockey
n, K = 1 << 20, 8
env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=1234)
out = env.make_outputs(K, ("events", "obs_next"))
acts = [torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda") for _ in range(4)]
for i in range(300): env.step(acts[i %% 4], K=K, out=out)
torch.cuda.synchronize()
best = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(200): env.step(acts[i %% 4], K=K, out=out)
    b.record(); torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / 200 * 1e3)
from air_hockey_training_environment_b200 import InterleavedStepper
inter = {}
for parts in (2, 3, 4):
    st = InterleavedStepper(env, parts=parts)
    for i in range(50): st.step(acts[i %% 4], K, out)
    st.join(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(200): st.step(acts[i %% 4], K, out)
    st.join(); b.record(); torch.cuda.synchronize()
    inter[parts] = round(a.elapsed_time(b) / 200 * 1e3, 2)
print(json.dumps({"us": [round(x, 2) for x in best], "interleaved_us": inter, "stats_frames": env.stats()["frames"], "launch": env.launch_info()}))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, AH_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(os.path.basename(lib), line, flush=True)
