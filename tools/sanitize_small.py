"""Small run that touches every kernel and variant once (for compute-sanitizer memcheck)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from air_hockey_training_environment_b200 import BatchedAirHockey, ReplayRing
from air_hockey_training_environment_b200.rollout import RolloutCollector
from air_hockey_training_environment_b200.host import HostStepper
for dt in (torch.float32, torch.float64):
    n = 1000                                   # ragged: not a multiple of 32 or 256
    env = BatchedAirHockey(n, dtype=dt, seed=1)
    dev = env.device
    env.step(None, K=50)
    env.step(torch.randint(0, 4, (7, n), dtype=torch.int8, device=dev), K=7)
    env.step(torch.randint(0, 4, (n,), dtype=torch.int8, device=dev), K=3, repeat=True)
    env.step((torch.rand(5, n, 2, device=dev, dtype=dt) * 6 - 3), K=5)
    ring = ReplayRing(3, n, dt, dev)
    env.step(torch.randint(0, 4, (4, n), dtype=torch.int8, device=dev), K=4, ring=ring,
             out=env.make_outputs(4, ("obs_state", "obs_new_state", "reward", "done", "score", "game_over", "obs_next", "actions"), True))
    tab = torch.rand(n, 4, 2, dtype=torch.float64, device=dev) * 20 - 10
    env.step(None, K=300, out=env.make_outputs(300, ()), reset_table=tab)
    env.update_state(torch.randint(0, 4, (n,), dtype=torch.int8, device=dev), "robot")
    env.update_state(torch.rand(n, 2, device=dev, dtype=dt) * 400 + 450, "human")
    env.update_state(agent_name="computer")
    env.get_state("robot"); env.get_state("opponent")
    r, s, d = env.rewards(); env.update_score(s); env.reset(total=True, mask=(s != 0))
    env.stats(clear=True); env.counters(); env.compiled_without_fma_contraction()
    col = RolloutCollector(env, 6, use_graph=False, single_launch=False); col.collect()
    col2 = RolloutCollector(env, 6, single_launch=True); col2.collect()
big = BatchedAirHockey(70000, seed=2)              # prefetch variant (K <= 2, n >= 65536)
big.step(torch.randint(0, 4, (70000,), dtype=torch.int8, device=big.device), K=1)
big.step(torch.randint(0, 4, (2, 70000), dtype=torch.int8, device=big.device), K=2)
st = HostStepper(big, K=2, slots=2)
st.submit(torch.randint(0, 4, (2, 70000), dtype=torch.int8).pin_memory()).wait()
torch.cuda.synchronize()
print("sanitize_small ok")
