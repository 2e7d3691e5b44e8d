"""TEST INFRASTRUCTURE ONLY.  Times the reference's OWN Python implementation of the path on this host's cores.

BASELINE.md §4 / BASELINE.json config 1: a single ``AirHockey`` env, uniform-random robot (actions {0,1,2,3}) against
the ``computerAI`` heuristic, 10,000 frames of the ``Game.play`` loop (game.py:133-177, lib/agents/agent.py:32-78) with
the real ``AirHockey``, base ``Agent`` and ``Rewards`` classes:

    robot.move(a); score, obs = robot.step(a); env.update_score(score); env.update_state(agent_name="computer")
    if env.opponent_score == 10: env.reset(total=True)
    if env.robot_score == 10:    env.reset(total=True)

measured (i) in ONE process and (ii) with a ``multiprocessing.Pool`` over all host cores (one independent env per
worker; sum of per-worker rates and the wall-clock rate), each in two variants:

    "socket_stub"  the Redis client is stubbed at the client class (``redis.StrictRedis``), so the reference still
                   builds and JSON-encodes the ``components`` payload every sub-update (environment/airhockey.py:124-132,
                   lib/utils/connect.py:25-29) and still appends its per-``done`` CSV row (lib/rewards.py:129-140);
    "pure_sim"     ``RedisConnection.post/publish`` and ``record_data_csv`` no-op'd: the most favourable to the reference.

The reference comes from ``oracle/refshim.REFERENCE_ROOT``: the live sources in the development container, the
byte-compiled ``oracle/_ref`` (built by ``oracle/fetch_ref.py``) on the GPU box.  Runs in its own process tree
(``python -m oracle.time_reference``) so that ``bench.py`` never forks after initialising CUDA.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import platform
import time

import numpy as np

VARIANTS = ("socket_stub", "pure_sim")


def _worker(args):
    """One env, `frames` frames of the Game.play loop; returns (frames/s, robot goals, opponent goals, dones)."""
    seed, frames, variant = args
    from . import refshim
    refshim.install()
    if variant == "pure_sim":
        refshim.silence_side_channels()
    from environment import AirHockey
    from lib.agents.agent import Agent

    np.random.seed(seed)                                  # Puck.reset draws from the global stream (puck.py:115-116)
    actions = np.random.default_rng(seed).integers(0, 4, frames).tolist()
    env = AirHockey()
    robot = Agent(env)
    robot.name = "robot"
    goals_r = goals_o = dones = 0
    t0 = time.perf_counter()
    for a in actions:
        robot.move(a)                                     # game.py:136
        score, obs = robot.step(a)                        # game.py:139
        env.update_score(score)                           # game.py:142
        env.update_state(agent_name="computer")           # game.py:166
        goals_r += score > 0
        goals_o += score < 0
        dones += bool(obs.done)
        if env.opponent_score == 10:                      # game.py:169-172
            env.reset(total=True)
        if env.robot_score == 10:                         # game.py:174-177
            env.reset(total=True)
    dt = time.perf_counter() - t0
    return frames / dt, int(goals_r), int(goals_o), int(dones)


# ---- persistent pool: `bench.py --impl reference` times K "steps", each = every host core advancing its own env
_G = {}


def _pool_init(variant, seed0):
    from . import refshim
    refshim.install()
    if variant == "pure_sim":
        refshim.silence_side_channels()
    from environment import AirHockey
    from lib.agents.agent import Agent
    seed = seed0 + os.getpid()
    np.random.seed(seed % (2 ** 32))
    env = AirHockey()
    robot = Agent(env)
    robot.name = "robot"
    _G.update(env=env, robot=robot, rng=np.random.default_rng(seed))


def _pool_step(frames):
    env, robot = _G["env"], _G["robot"]
    for a in _G["rng"].integers(0, 4, frames).tolist():
        robot.move(a)
        score, _ = robot.step(a)
        env.update_score(score)
        env.update_state(agent_name="computer")
        if env.opponent_score == 10:
            env.reset(total=True)
        if env.robot_score == 10:
            env.reset(total=True)
    return frames


class ReferencePool:
    """`procs` worker processes, each owning one reference env; ``step(frames)`` advances every env by `frames`
    frames of the Game.play loop and returns the env-steps done (procs x frames)."""

    def __init__(self, procs: int = 0, variant: str = "pure_sim", seed: int = 1234):
        self.procs = procs or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.procs, initializer=_pool_init, initargs=(variant, seed))

    def step(self, frames: int) -> int:
        return sum(self.pool.map(_pool_step, [frames] * self.procs, chunksize=1))

    def close(self):
        self.pool.close()
        self.pool.join()


def measure(frames: int = 10000, procs: int = 0, variants=VARIANTS):
    from . import refshim
    procs = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    res = {"what": "the reference's own Python loop (AirHockey + Agent + Rewards as shipped), BASELINE config 1",
           "reference": refshim.reference_kind(), "reference_root": refshim.REFERENCE_ROOT,
           "frames_per_process": frames, "host_cores": procs, "os_cpu_count": os.cpu_count(),
           "cpu": platform.processor() or platform.machine(), "python": platform.python_version(),
           "numpy": np.__version__, "unit": "env-steps/s"}
    for v in variants:
        with ctx.Pool(1) as pool:
            one = pool.map(_worker, [(0, frames, v)])[0]
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            rows = pool.map(_worker, [(s + 1, frames, v) for s in range(procs)], chunksize=1)
        wall = time.perf_counter() - t0
        tot = sum(r[1] + r[2] for r in rows)
        res[v] = {"one_process": one[0], "pool_processes": procs, "pool_sum_of_rates": sum(r[0] for r in rows),
                  "pool_wall_clock_rate": procs * frames / wall, "pool_wall_s": wall,
                  "sub_updates_per_s_one_process": 3 * one[0],
                  "goals_per_1000_frames": 1000.0 * tot / (procs * frames),
                  "dones_per_1000_frames": 1000.0 * sum(r[3] for r in rows) / (procs * frames)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--variants", default=",".join(VARIANTS))
    a = ap.parse_args()
    print(json.dumps(measure(a.frames, a.procs, tuple(a.variants.split(",")))))


if __name__ == "__main__":
    main()
