#!/usr/bin/env python
"""End-to-end example: Dueling DDQN on the batched air-hockey environment, everything on the device.

    python examples/train_dueling_dqn.py [--envs 16384] [--frames 32] [--iters 30]

Each iteration = ONE kernel launch that plays `frames` frames of `envs` games with the dueling Q-network and the
reference's epsilon-greedy exploration evaluated INSIDE the step kernel (policy.DevicePolicy / AH_ACT_POLICY_MLP), the
Observation tuples appended to the device replay ring by the same launch (ReplayRing: lib/buffer.py on the device), then a
minibatch from the one-kernel sampler and a Dueling-DDQN update in torch (learners.DuelingLearner:
lib/agents/dueling/dueling.py:125-178)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from air_hockey_training_environment_b200 import BatchedAirHockey, ReplayRing, policy
from air_hockey_training_environment_b200.dist import summarize
from air_hockey_training_environment_b200.learners import DuelingLearner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--lr", type=float, default=1e-4)
    a = ap.parse_args()
    env = BatchedAirHockey(a.envs, dtype=torch.float32, device="cuda", seed=0)
    ring = ReplayRing(4 * a.frames, a.envs, env.dtype, env.device)
    learner = DuelingLearner(env, ring, lr=a.lr, batch_size=1 << 16)
    dp = policy.DevicePolicy(learner.online, env.device, select="eps_greedy", epsilon=1.0, initial_epsilon=1.0,
                             final_epsilon=0.05, observe=a.frames, explore=a.iters // 2 + 1)
    out = env.make_outputs(a.frames, ("reward", "done"))
    for it in range(a.iters):
        t0 = time.perf_counter()
        env.step(None, K=a.frames, out=out, policy=dp, ring=ring)       # net + epsilon-greedy + physics + ring append: one launch
        info = learner.update()
        dp.refresh()                                                    # the next launch plays the updated network
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        s = summarize(env.stats_tensor(clear=True))
        print(f"iter {it:3d}  {a.envs * a.frames / dt:10.3e} env-steps/s  loss {info['loss']:.4f}  mean reward {s['mean_reward']:+.3f}  "
              f"goals/1k frames robot {s['robot_goals_per_1k_frames']:.2f} opponent {s['opponent_goals_per_1k_frames']:.2f}")


if __name__ == "__main__":
    main()
