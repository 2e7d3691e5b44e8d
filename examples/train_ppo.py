#!/usr/bin/env python
"""End-to-end example: PPO on the batched air-hockey environment, everything on the device.

    python examples/train_ppo.py [--envs 65536] [--frames 128] [--iters 50]

Each iteration = ONE kernel launch that plays `frames` frames of `envs` games under the in-kernel actor
(rollout.RolloutCollector(single_launch=True)) + a torch PPO update on the zero-copy rollout tensors
(learners.PPOLearner; the reference's losses, lib/utils/helpers.py:11-46).  Prints env-steps/s and the statistics
the reference logs (goals, wins, mean reward)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from air_hockey_training_environment_b200 import BatchedAirHockey
from air_hockey_training_environment_b200.dist import summarize
from air_hockey_training_environment_b200.learners import PPOLearner
from air_hockey_training_environment_b200.rollout import RolloutCollector


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--gamma", type=float, default=0.8)          # lib/agents/ppo/model.py:37
    a = ap.parse_args()
    torch.manual_seed(0)
    env = BatchedAirHockey(a.envs, seed=0)
    col = RolloutCollector(env, a.frames, single_launch=True)
    learner = PPOLearner(col, gamma=a.gamma, actor_lr=a.lr, critic_lr=a.lr, epochs=4, minibatch=1 << 20)
    for it in range(a.iters):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        col.collect()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        log = learner.update()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        s = summarize(env.stats_tensor(clear=True))
        print(f"it {it:3d}  rollout {1e3 * (t1 - t0):6.2f} ms ({a.envs * a.frames / (t1 - t0):.2e} env-steps/s)  update {1e3 * (t2 - t1):7.1f} ms  "
              f"mean_reward {log['mean_reward']:+.3f}  robot goals/1k {s['robot_goals_per_1k_frames']:.2f}  "
              f"opponent goals/1k {s['opponent_goals_per_1k_frames']:.2f}  touches/1k {s['dones_per_1k_frames']:.2f}  "
              f"actor_loss {log['actor_loss']:+.4f}  critic_loss {log['critic_loss']:.4f}")


if __name__ == "__main__":
    main()
