"""Tuning aid: time BASELINE config 5's single-launch rollouts (65,536 envs x 128 frames) for every library variant given
on the command line (AH_B200_LIB selects the library).  python tools/bench_cfg5_variants.py lib1.so ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch, json
sys.path.insert(0, %r)
from air_hockey_training_environment_b200 import BatchedAirHockey, policy as ahp
from air_hockey_training_environment_b200.rollout import RolloutCollector
n, T = 65536, 128
res = {}
for name, kw in (("hard_coded_actor", lambda: dict(single_launch=True)),
                 ("exported_ppo_actor", lambda: dict(device_policy=ahp.DevicePolicy(ahp.ppo_actor_discrete().cuda(), "cuda"))),
                 ("exported_dueling", lambda: dict(device_policy=ahp.DevicePolicy(ahp.DuelingQNet(4).cuda(), "cuda", select="eps_greedy",
                                                                                   epsilon=0.1, initial_epsilon=0.1, final_epsilon=0.1)))):
    env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=7)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))
    col = RolloutCollector(env, T, seed=7, **kw())
    for _ in range(3): col.collect()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): col.collect()
    b.record(); torch.cuda.synchronize()
    res[name] = round(a.elapsed_time(b) / 20, 4)
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, AH_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(os.path.basename(lib), "ms per rollout:", line, flush=True)
