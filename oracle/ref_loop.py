"""TEST INFRASTRUCTURE ONLY.  Canonical frame loop over the *live* Python reference.

Drives the real ``AirHockey`` (environment/airhockey.py), the real base ``Agent``
(lib/agents/agent.py, whose ``update`` is a no-op) and the real ``Rewards`` (lib/rewards.py)
exactly in the order of ``Game.play`` / ``Game.robot_player`` (game.py:114-177), with the policy
replaced by a supplied action table and ``np.random.uniform`` replaced by a supplied table of
reset velocities so the run is reproducible by the C oracle and by the CUDA kernels.

One *frame* (= one env-step):

    robot.move(a)                       game.py:136  -> agent.py:32-36 -> airhockey.py:96
    score, obs = robot.step(a)          game.py:139  -> agent.py:57-78 (get_state, move, get_state, Rewards)
    env.update_score(score)             game.py:142  -> airhockey.py:149-169
    stats()                             game.py:145  (robot_goal / opponent_goal / *_win flags)
    env.update_state(agent_name="computer")   game.py:166
    if opponent_score == 10: reset(total=True)   game.py:169-172
    if robot_score == 10:    reset(total=True)   game.py:174-177
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import refshim

STATE_FIELDS = ("puck", "robot", "opponent")


class _ResetFeed:
    """Replacement for ``np.random.uniform`` that pops pre-drawn values (puck.py:115-116)."""

    def __init__(self):
        self.table = None
        self.cursor = 0

    def bind(self, table: np.ndarray):
        self.table = table
        self.cursor = 0

    def __call__(self, low=0.0, high=1.0, size=None):
        assert size is None and low == -10 and high == 10, "only Puck.reset draws are expected"
        v = float(self.table[self.cursor])
        self.cursor += 1
        return v


def _obs8(state) -> np.ndarray:
    # lib/utils/helpers.py:72-81 flattens State(agent_location, puck_location, agent_velocity, puck_velocity)
    return np.asarray(state, dtype=np.float64).flatten()


def snapshot(env, tracker) -> np.ndarray:
    """13 reals: puck(x,y,dx,dy) robot(x,y,dx,dy) opponent(x,y,dx,dy) prev_mallet_to_puck_distance."""
    p, r, o = env.puck, env.robot, env.opponent
    return np.array([p.x, p.y, p.dx, p.dy, r.x, r.y, r.dx, r.dy, o.x, o.y, o.dx, o.dy,
                     tracker.prev_mallet_to_puck_distance], dtype=np.float64)


def run_reference(actions: np.ndarray, reset_table: np.ndarray,
                  init_state: Optional[np.ndarray] = None,
                  init_scores: Optional[np.ndarray] = None,
                  record_every: int = 1, init_frame: bool = False) -> Dict[str, np.ndarray]:
    """Run ``T`` frames of ``N`` independent reference games.

    actions      int   [T, N]      discrete actions (any int; 0..3 nudge, others no nudge)
                 or float [T, N, 2] continuous actions (added to the mallet velocity)
    reset_table  f64   [N, R, 2]   (dx, dy) per reset, consumed in order per env
    init_state   f64   [N, 13]     optional initial 13 reals (see ``snapshot``)
    init_scores  int   [N, 2]      optional (robot_score, opponent_score)
    init_frame   frame 0 is the reference's move-only first frame (game.py:118-130): `robot.move(a)` without
                 `robot.step`, then the opponent's sub-update; its reward / score / done rows are left at zero
    """
    refshim.install()
    refshim.silence_side_channels()
    from environment import AirHockey
    from lib.agents.agent import Agent

    continuous = actions.ndim == 3
    T, N = actions.shape[:2]
    n_rec = T // record_every
    out = {
        "states": np.zeros((n_rec, N, 13), np.float64),       # state after frame t (t = record_every-1, ...)
        "scores": np.zeros((n_rec, N, 2), np.int32),
        "obs_state": np.zeros((T, N, 8), np.float64),          # Observation.state      (agent.py:61)
        "obs_new_state": np.zeros((T, N, 8), np.float64),      # Observation.new_state  (agent.py:67)
        "obs_next": np.zeros((T, N, 8), np.float64),           # get_state() at the end of the frame
        "reward": np.zeros((T, N), np.int32),
        "score": np.zeros((T, N), np.int8),
        "done": np.zeros((T, N), np.uint8),
        "robot_win": np.zeros((T, N), np.uint8),               # game.py:101-104 stats flags
        "opponent_win": np.zeros((T, N), np.uint8),
        "reset_cursor": np.zeros((N,), np.int32),              # draws/2 consumed per env
        "final_state": np.zeros((N, 13), np.float64),
        "final_scores": np.zeros((N, 2), np.int32),
    }
    feed = _ResetFeed()
    saved_uniform = np.random.uniform
    np.random.uniform = feed
    try:
        for e in range(N):
            env = AirHockey()
            robot = Agent(env)
            robot.name = "robot"
            tracker = robot.reward_tracker
            if init_state is not None:
                s = init_state[e]
                (env.puck.x, env.puck.y, env.puck.dx, env.puck.dy) = (float(v) for v in s[0:4])
                (env.robot.x, env.robot.y, env.robot.dx, env.robot.dy) = (float(v) for v in s[4:8])
                (env.opponent.x, env.opponent.y, env.opponent.dx, env.opponent.dy) = (float(v) for v in s[8:12])
                tracker.prev_mallet_to_puck_distance = float(s[12])
            if init_scores is not None:
                env.robot_score, env.opponent_score = int(init_scores[e, 0]), int(init_scores[e, 1])
            feed.bind(reset_table[e].reshape(-1))
            for t in range(T):
                a = (float(actions[t, e, 0]), float(actions[t, e, 1])) if continuous else int(actions[t, e])
                if init_frame and t == 0:
                    robot.move(a)                               # game.py:128 (self.init: move only)
                    env.update_state(agent_name="computer")     # game.py:166
                    out["obs_next"][t, e] = _obs8(env.get_state("robot"))
                    out["states"][0, e] = snapshot(env, tracker)
                    continue
                robot.move(a)                                   # game.py:136
                score, obs = robot.step(a)                      # game.py:139
                env.update_score(score)                         # game.py:142
                rw = 1 if env.robot_score == 10 else 0          # game.py:101 (stats runs here, :145)
                ow = 1 if env.opponent_score == 10 else 0       # game.py:106
                env.update_state(agent_name="computer")         # game.py:166
                if env.opponent_score == 10:                    # game.py:169-172
                    env.reset(total=True)
                if env.robot_score == 10:                       # game.py:174-177
                    env.reset(total=True)
                out["obs_state"][t, e] = _obs8(obs.state)
                out["obs_new_state"][t, e] = _obs8(obs.new_state)
                out["obs_next"][t, e] = _obs8(env.get_state("robot"))
                out["reward"][t, e] = obs.reward
                out["score"][t, e] = score
                out["done"][t, e] = obs.done
                out["robot_win"][t, e] = rw
                out["opponent_win"][t, e] = ow
                if (t + 1) % record_every == 0:
                    out["states"][(t + 1) // record_every - 1, e] = snapshot(env, tracker)
                    out["scores"][(t + 1) // record_every - 1, e] = (env.robot_score, env.opponent_score)
            assert feed.cursor % 2 == 0
            out["reset_cursor"][e] = feed.cursor // 2
            out["final_state"][e] = snapshot(env, tracker)
            out["final_scores"][e] = (env.robot_score, env.opponent_score)
    finally:
        np.random.uniform = saved_uniform
        refshim.uninstall()                    # leave no stub modules behind in the calling process
    return out


def run_reference_substeps(calls, init_state: Optional[np.ndarray] = None) -> np.ndarray:
    """Apply a list of ``(action, agent_name)`` to one fresh reference env via ``update_state``
    (airhockey.py:96-134) and return the 12 state reals after each call -- pins the low-level API."""
    refshim.install()
    refshim.silence_side_channels()
    from environment import AirHockey

    env = AirHockey()
    if init_state is not None:
        s = init_state
        (env.puck.x, env.puck.y, env.puck.dx, env.puck.dy) = (float(v) for v in s[0:4])
        (env.robot.x, env.robot.y, env.robot.dx, env.robot.dy) = (float(v) for v in s[4:8])
        (env.opponent.x, env.opponent.y, env.opponent.dx, env.opponent.dy) = (float(v) for v in s[8:12])
    rows = []
    for action, name in calls:
        env.update_state(action, name)
        p, r, o = env.puck, env.robot, env.opponent
        rows.append([p.x, p.y, p.dx, p.dy, r.x, r.y, r.dx, r.dy, o.x, o.y, o.dx, o.dy])
    refshim.uninstall()
    return np.array(rows, dtype=np.float64)
