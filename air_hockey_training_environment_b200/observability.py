"""Host observability adapters -- OFF the hot path by construction.

The reference pushes, on EVERY sub-update, a JSON payload to Redis for its web UI and appends a CSV row per frame
(25 % of its frame time).  Here one chosen game can be sampled into the same wire format whenever a
UI wants a frame, and the per-frame CSV rows become aggregates of the device-side statistics vector.

Wire format (paths relative to the reference root):
    key "components"   environment/airhockey.py:124-132   locations int()-truncated, velocities raw
    key "scores"       environment/airhockey.py:57,154,164,179
    scores.csv row     game.py:82-112                     robot_goal, opponent_goal, robot_win, opponent_win
    rewards_robot.csv  lib/rewards.py:129-140             mean_reward
Values are returned as plain dicts; the caller does `redis.set(key, json.dumps(value))` exactly like
lib/utils/connect.py:25-29 if it has a Redis.
"""
from __future__ import annotations

from typing import Dict

import torch

from .env import BatchedAirHockey


def components_payload(env: BatchedAirHockey, i: int = 0) -> Dict[str, Dict]:
    """{"puck": {"location": [x, y], "velocity": [dx, dy]}, "robot": {...}, "opponent": {...}} for game ``i``."""
    rows = torch.stack([env.puck[i], env.robot[i], env.opponent[i]]).double().cpu().tolist()   # one small D2H copy
    out = {}
    for name, (x, y, dx, dy) in zip(("puck", "robot", "opponent"), rows):
        out[name] = {"location": [int(x), int(y)], "velocity": [dx, dy]}
    return out


def scores_payload(env: BatchedAirHockey, i: int = 0) -> Dict[str, int]:
    r, o = env.scores[i].tolist()
    return {"robot_score": int(r), "opponent_score": int(o)}


def scores_csv_aggregate(stats: Dict[str, int]) -> Dict[str, float]:
    """The reference writes one scores.csv row per frame with 0/1 flags; their sums over the frames covered by a
    statistics vector (``BatchedAirHockey.stats()`` or the all-reduced one) are exactly these counters."""
    f = max(int(stats["frames"]), 1)
    return {"frames": f, "robot_goal": int(stats["robot_goals"]), "opponent_goal": int(stats["opponent_goals"]),
            "robot_win": int(stats["robot_wins"]), "opponent_win": int(stats["opponent_wins"]),
            "mean_reward": stats["reward_sum"] / f}
