"""Run ONE mode of the step kernel repeatedly (for ncu captures of the non-headline variants).
usage: python tools/run_mode.py k1 | fp64 | policy | ring | ringc | unpacked | philox"""
import sys
sys.path.insert(0, "/root/repo")
import torch
from air_hockey_training_environment_b200 import BatchedAirHockey
mode = sys.argv[1]
if mode == "k1":                       # HBM-bound rollout mode: persistent grid + prefetch variant
    n = 1 << 20
    env = BatchedAirHockey(n, seed=1)
    out = env.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next"))
    acts = [torch.randint(0, 4, (n,), dtype=torch.int8, device="cuda") for _ in range(4)]
    for i in range(300):
        env.step(acts[i % 4], K=1, out=out)
elif mode == "fp64":                   # validation build
    n = 1 << 18
    env = BatchedAirHockey(n, dtype=torch.float64, seed=1)
    out = env.make_outputs(8, ("reward", "done", "score", "game_over", "obs_next"))
    acts = [torch.randint(0, 4, (8, n), dtype=torch.int8, device="cuda") for _ in range(4)]
    for i in range(80):
        env.step(acts[i % 4], K=8, out=out)
elif mode == "policy":                 # in-kernel actor rollout (config 5)
    from air_hockey_training_environment_b200.rollout import RolloutCollector
    col = RolloutCollector(BatchedAirHockey(65536, seed=1), 128, single_launch=True)
    for i in range(12):
        col.collect()
elif mode in ("ring", "ringc"):        # config 4's per-GPU work: K = 8 step + fused replay-ring append (HBM-bound, write-heavy);
                                       # ringc = the compact tuple layout (AH_RING_COMPACT)
    from air_hockey_training_environment_b200 import ReplayRing
    n, K = 1 << 20, 8
    env = BatchedAirHockey(n, seed=1)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))
    ring = ReplayRing(2 * K, n, env.dtype, env.device, compact=(mode == "ringc"))
    from air_hockey_training_environment_b200.env import pack_actions
    out = env.make_outputs(K, ("events", "obs_next"))
    acts = [pack_actions(torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda")) for _ in range(4)]
    for i in range(60):
        env.step(acts[i % 4], K=K, out=out, ring=ring)
elif mode in ("unpacked", "philox"):   # the packed core behind the general API's tensors / the Philox simulation
    n, K = 1 << 20, 8
    env = BatchedAirHockey(n, seed=1)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))
    if mode == "unpacked":
        out = env.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
        acts = [torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda") for _ in range(4)]
    else:
        out, acts = env.make_outputs(K, ("obs_next",)), [None] * 4
    for i in range(60):
        env.step(acts[i % 4], K=K, out=out)
torch.cuda.synchronize()
print("ok", mode)
