"""Tuning aid: BASELINE config 4's per-GPU step (2^20 envs, K = 8, compact ring append) issued as 1 / 2 / 3 / 4 interleaved slices."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from air_hockey_training_environment_b200 import BatchedAirHockey, InterleavedStepper, ReplayRing
n, K = 1 << 20, 8
env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=1234)
env.step(None, K=2000, out=env.make_outputs(2000, ()))
ring = ReplayRing(2 * K, n, env.dtype, env.device, compact=True)
out = env.make_outputs(K, ("events", "obs_next"))
acts = [torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda") for _ in range(4)]
res = {}
for parts in (1, 2, 3, 4):
    st = InterleavedStepper(env, parts=parts) if parts > 1 else None
    step = (lambda i: st.step(acts[i % 4], K, out, ring=ring)) if st else (lambda i: env.step(acts[i % 4], K=K, out=out, ring=ring))
    for i in range(30): step(i)
    if st: st.join()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(200): step(i)
    if st: st.join()
    b.record(); torch.cuda.synchronize()
    res[parts] = round(a.elapsed_time(b) / 200 * 1e3, 2)
print(json.dumps({"us_per_step_by_slices": res, "appended": ring.appended, "counters_equal": len(set(ring.counter.tolist())) == 1}))
