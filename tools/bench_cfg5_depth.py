"""Tuning aid: where does the exported-policy rollout (BASELINE config 5, one launch) spend its time?  Rollouts of 65,536
envs x 128 frames with exported networks of growing depth, against the hard-coded actor.  python tools/bench_cfg5_depth.py"""
import json, sys, os, torch
from torch import nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from air_hockey_training_environment_b200 import BatchedAirHockey, policy as ahp
from air_hockey_training_environment_b200.rollout import RolloutCollector
n, T = 65536, 128
nets = {"8-4": nn.Sequential(nn.Linear(8, 4), nn.Softmax(dim=-1)),
        "8-4-4": nn.Sequential(nn.Linear(8, 4), nn.ReLU(), nn.Linear(4, 4), nn.Softmax(dim=-1)),
        "8-8-4 (a2c)": ahp.a2c_actor(),
        "8-4-6-4-4 (ppo)": ahp.ppo_actor_discrete()}
res = {}
for name, net in [("hard_coded", None)] + list(nets.items()):
    env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=7)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))
    if net is None:
        kw = dict(single_launch=True)
    elif "continuous" in name:
        kw = dict(device_policy=ahp.DevicePolicy(net.cuda().eval(), "cuda", head="continuous", select="gaussian"))
    else:
        kw = dict(device_policy=ahp.DevicePolicy(net.cuda().eval(), "cuda"))
    col = RolloutCollector(env, T, seed=7, **kw)
    for _ in range(3): col.collect()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): col.collect()
    b.record(); torch.cuda.synchronize()
    res[name] = round(a.elapsed_time(b) / 20, 4)
print(json.dumps(res))
