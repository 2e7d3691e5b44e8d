"""TEST INFRASTRUCTURE ONLY.  Times the LIVE Python reference (development container only; the reference cannot
travel to the GPU box) on BASELINE config 1: single AirHockey env, uniform-random robot vs computerAI, 10,000 frames,
Redis and CSV side channels no-op'd -- (i) one process, (ii) a multiprocessing pool over all cores.

    python -m oracle.time_reference [--frames 10000] > profiles/r1_python_reference_timing_dev_container.json
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import platform
import time

import numpy as np


def _run(args):
    seed, frames = args
    from . import ref_loop
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (frames, 1))
    table = rng.uniform(-10, 10, (1, 4 + frames // 50, 2))
    t0 = time.perf_counter()
    ref_loop.run_reference(actions, table)
    return frames / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    a = ap.parse_args()
    one = _run((0, a.frames))
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        rates = pool.map(_run, [(s + 1, a.frames) for s in range(cores)])
    wall = time.perf_counter() - t0
    print(json.dumps({"what": "LIVE Python reference, config 1 (single env, random robot vs computerAI), side channels no-op'd",
                      "where": "development container (NOT the B200 box host)", "cpu": platform.processor() or platform.machine(),
                      "frames_per_process": a.frames, "one_process_env_steps_per_s": one, "pool_processes": cores,
                      "pool_sum_of_rates_env_steps_per_s": sum(rates), "pool_wall_clock_env_steps_per_s": cores * a.frames / wall,
                      "python": platform.python_version(), "numpy": np.__version__}, indent=1))


if __name__ == "__main__":
    main()
