"""Tuning aid: time the packed K=8 step WITH the fused ring append (2^20 envs, FP32, config 4's per-GPU work) for every
library variant given on the command line (AH_B200_LIB selects the library; AH_RING_COMPACT=0 the rows layout).  python tools/bench_ring_variants.py lib1.so ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch, json
sys.path.insert(0, %r)
from air_hockey_training_environment_b200 import BatchedAirHockey, ReplayRing
n, K = 1 << 20, 8
env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=1234)
env.step(None, K=2000, out=env.make_outputs(2000, ()))
import os
ring = ReplayRing(2 * K, n, env.dtype, env.device, compact=os.environ.get("AH_RING_COMPACT", "1") == "1")
out = env.make_outputs(K, ("events", "obs_next"))
acts = [torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda") for _ in range(4)]
for i in range(40): env.step(acts[i %% 4], K=K, out=out, ring=ring)
torch.cuda.synchronize()
best = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(100): env.step(acts[i %% 4], K=K, out=out, ring=ring)
    b.record(); torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / 100 * 1e3)
print(json.dumps({"compact": ring.compact, "us": [round(x, 2) for x in best], "appended": ring.appended, "checksum": float(ring.state.double().sum())}))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, AH_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(os.path.basename(lib), line, flush=True)
