"""TEST INFRASTRUCTURE ONLY.  Large live differential run: C oracle vs the Python reference at
BASELINE config-2 size (4,096 envs x 1,000 frames), parallelised over processes.

    python -m oracle.diff_vs_reference [--envs 4096] [--frames 1000] [--procs 8]

Prints one summary line; the result of the run in the development container is recorded in DESIGN.md.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import time

import numpy as np


def _chunk(args):
    seed, n, T = args
    from . import oracle, ref_loop
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 64, 2))
    ref = ref_loop.run_reference(actions, table)
    st, sc = oracle.initial_state(n)
    out = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=1)
    same = (np.array_equal(ref["states"].view(np.uint64), out["states_rec"].view(np.uint64))
            and np.array_equal(ref["scores"], out["scores_rec"])
            and np.array_equal(ref["reward"], out["reward"])
            and np.array_equal(ref["score"], out["score"])
            and np.array_equal(ref["done"], out["done"])
            and np.array_equal(ref["robot_win"] + 2 * ref["opponent_win"], out["game_over"])
            and all(np.array_equal(ref[k].view(np.uint64), out[k].view(np.uint64))
                    for k in ("obs_state", "obs_new_state", "obs_next")))
    maxabs = float(np.nanmax(np.abs(np.nan_to_num(ref["states"], posinf=0) - np.nan_to_num(out["states_rec"], posinf=0))))
    return same, maxabs, int((ref["score"] != 0).sum()), int((ref["robot_win"] + ref["opponent_win"]).sum()), int(ref["done"].sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--procs", type=int, default=8)
    a = ap.parse_args()
    per = 64
    jobs = [(1000 + i, per, a.frames) for i in range(a.envs // per)]
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        res = pool.map(_chunk, jobs)
    print({"envs": a.envs, "frames": a.frames, "bit_identical": all(r[0] for r in res),
           "max_abs_diff": max(r[1] for r in res), "goals": sum(r[2] for r in res),
           "games_finished": sum(r[3] for r in res), "dones": sum(r[4] for r in res),
           "seconds": round(time.time() - t0, 1)})


if __name__ == "__main__":
    main()
