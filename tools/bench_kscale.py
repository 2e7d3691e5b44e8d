"""Tuning aid: packed step time vs K (frames per launch) at 2^20 envs: slope = steady-state cost per frame,
intercept = per-launch overhead (state load / store, prologue, tail)."""
import sys, os, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from air_hockey_training_environment_b200 import BatchedAirHockey
n = 1 << 20
env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=1234)
env.step(None, K=2400, out=env.make_outputs(2400, ()))
res = {}
for K in (4, 8, 16, 32, 64):
    out = env.make_outputs(K, ("events", "obs_next"))
    acts = [torch.randint(0, 4, ((K + 3) // 4, n, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for i in range(20): env.step(acts[i % 2], K=K, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, 800 // K)
    a.record()
    for i in range(reps): env.step(acts[i % 2], K=K, out=out)
    b.record(); torch.cuda.synchronize()
    res[K] = a.elapsed_time(b) / reps * 1e3
print(json.dumps(res))
ks = sorted(res)
slope = (res[ks[-1]] - res[ks[0]]) / (ks[-1] - ks[0])
print("us/frame slope %.3f, intercept %.2f us; K=8 predicted steady %.1f us of %.1f" % (slope, res[ks[0]] - slope * ks[0], 8 * slope, res[8]))
