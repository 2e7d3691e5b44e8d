"""TEST INFRASTRUCTURE ONLY.  Generates the committed fixtures under tests/golden/ by running the
LIVE Python reference (/root/reference, via oracle/refshim.py + oracle/ref_loop.py).

Run in the development container only (the reference does not exist on the GPU box):

    python -m oracle.gen_golden

Every file records the numpy version, the np.dot FMA self-probe and the seeds,
so a reader can tell which numpy build the trajectories belong to.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import ref_loop, refshim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _header(seed: int) -> dict:
    return {
        "numpy_version": np.array(np.__version__),
        "dot_fma": np.array(int(refshim.numpy_dot_is_fma())),
        "seed": np.array(seed),
        "generator": np.array("oracle/gen_golden.py"),
        "python": np.array(sys.version.split()[0]),
    }


def scrambled_states(rng: np.random.Generator, n: int) -> np.ndarray:
    """Random non-integer states that exercise contacts, wall clamps, both speed-clamp signs and goals."""
    s = np.zeros((n, 13), np.float64)
    s[:, 0] = rng.uniform(0, 900, n)
    s[:, 1] = rng.uniform(0, 480, n)
    s[:, 2:4] = rng.uniform(-13, 13, (n, 2))
    s[:, 4] = rng.uniform(20, 430, n)
    s[:, 5] = rng.uniform(20, 460, n)
    s[:, 6:8] = rng.uniform(-12, 12, (n, 2))
    s[:, 8] = rng.uniform(470, 880, n)
    s[:, 9] = rng.uniform(20, 460, n)
    s[:, 10:12] = rng.uniform(-6, 6, (n, 2))
    s[:, 12] = np.where(rng.random(n) < 0.3, np.inf, rng.uniform(0, 600, n))
    # put a mallet close to the puck in many envs
    near = rng.random(n)
    r_near = near < 0.35
    s[r_near, 0] = np.clip(s[r_near, 4] + rng.uniform(-38, 38, r_near.sum()), 0, 900)
    s[r_near, 1] = np.clip(s[r_near, 5] + rng.uniform(-38, 38, r_near.sum()), 0, 480)
    o_near = (near >= 0.35) & (near < 0.7)
    s[o_near, 0] = np.clip(s[o_near, 8] + rng.uniform(-38, 38, o_near.sum()), 0, 900)
    s[o_near, 1] = np.clip(s[o_near, 9] + rng.uniform(-38, 38, o_near.sum()), 0, 480)
    # a few integer-valued and a few exactly-coincident cases (distance == 0 branch, airhockey.py:217)
    s[: n // 8] = np.round(s[: n // 8])
    s[n // 8, 0:2] = s[n // 8, 4:6]
    return s


def gen_canonical(seed=101, n=16, T=1000):
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 64, 2))
    ref = ref_loop.run_reference(actions, table)
    assert ref["reset_cursor"].max() < 64
    np.savez_compressed(os.path.join(GOLDEN, "canonical_16x1000.npz"), actions=actions, reset_table=table,
                        **{k: ref[k] for k in ("states", "scores", "obs_state", "obs_new_state", "obs_next")},
                        reward=ref["reward"].astype(np.int8), score=ref["score"], done=ref["done"],
                        game_over=(ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8),
                        reset_cursor=ref["reset_cursor"], **_header(seed))


def gen_long(seed=202, n=256, T=1000, every=100):
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 64, 2))
    ref = ref_loop.run_reference(actions, table, record_every=every)
    assert ref["reset_cursor"].max() < 64
    np.savez_compressed(os.path.join(GOLDEN, "long_256x1000.npz"), actions=actions,
                        reset_table=table.astype(np.float64), record_every=np.array(every),
                        states=ref["states"], scores=ref["scores"],
                        reward=ref["reward"].astype(np.int8), score=ref["score"], done=ref["done"],
                        game_over=(ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8),
                        reset_cursor=ref["reset_cursor"], final_state=ref["final_state"],
                        final_scores=ref["final_scores"], **_header(seed))


def gen_scrambled(seed=303, n=96, T=200):
    rng = np.random.default_rng(seed)
    actions = rng.integers(-1, 6, (T, n)).astype(np.int8)      # includes out-of-range ints: no nudge
    table = rng.uniform(-10, 10, (n, 32, 2))
    init = scrambled_states(rng, n)
    init_scores = rng.integers(0, 10, (n, 2)).astype(np.int32)
    init_scores[: n // 4] = rng.integers(8, 10, (n // 4, 2))   # close to game over (game.py:169-177)
    ref = ref_loop.run_reference(actions, table, init_state=init, init_scores=init_scores)
    assert ref["reset_cursor"].max() < 32
    np.savez_compressed(os.path.join(GOLDEN, "scrambled_96x200.npz"), actions=actions, reset_table=table,
                        init_state=init, init_scores=init_scores,
                        **{k: ref[k] for k in ("states", "scores", "obs_state", "obs_new_state", "obs_next")},
                        reward=ref["reward"].astype(np.int8), score=ref["score"], done=ref["done"],
                        game_over=(ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8),
                        reset_cursor=ref["reset_cursor"], **_header(seed))


def gen_continuous(seed=404, n=16, T=300):
    rng = np.random.default_rng(seed)
    actions = rng.uniform(-3, 3, (T, n, 2))                     # game.py:122 range
    table = rng.uniform(-10, 10, (n, 32, 2))
    ref = ref_loop.run_reference(actions, table)
    assert ref["reset_cursor"].max() < 32
    np.savez_compressed(os.path.join(GOLDEN, "continuous_16x300.npz"), actions=actions, reset_table=table,
                        **{k: ref[k] for k in ("states", "scores", "obs_state", "obs_new_state", "obs_next")},
                        reward=ref["reward"].astype(np.int8), score=ref["score"], done=ref["done"],
                        game_over=(ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8),
                        reset_cursor=ref["reset_cursor"], **_header(seed))


def gen_substeps(seed=505, n_seq=24, L=40):
    """Sequences of raw ``update_state`` calls (airhockey.py:96-134) mixing robot / human / computer."""
    rng = np.random.default_rng(seed)
    agents = np.zeros((n_seq, L), np.int8)          # 0 robot, 1 human, 2 computer
    kinds = np.zeros((n_seq, L), np.int8)           # robot only: 0 discrete, 1 continuous
    ia = np.zeros((n_seq, L), np.int32)
    axy = np.zeros((n_seq, L, 2), np.float64)
    init = scrambled_states(rng, n_seq)
    init[0] = [450, 240, -3, 0, 350, 240, 0, 0, 550, 240, 0, 0, np.inf]
    states = np.zeros((n_seq, L, 12), np.float64)
    names = ("robot", "human", "computer")
    for s in range(n_seq):
        calls = []
        for t in range(L):
            ag = int(rng.integers(0, 3))
            agents[s, t] = ag
            if ag == 0:
                if rng.random() < 0.5:
                    ia[s, t] = int(rng.integers(0, 4))
                    calls.append((int(ia[s, t]), "robot"))
                else:
                    kinds[s, t] = 1
                    axy[s, t] = rng.uniform(-3, 3, 2)
                    calls.append(((float(axy[s, t, 0]), float(axy[s, t, 1])), "robot"))
            elif ag == 1:
                # the human's mallet is teleported to the given location (airhockey.py:104)
                axy[s, t] = np.round(rng.uniform([400, 0], [900, 480]))
                calls.append(((float(axy[s, t, 0]), float(axy[s, t, 1])), "human"))
            else:
                calls.append(("", "computer"))
        states[s] = ref_loop.run_reference_substeps(calls, init_state=init[s])
    # the reference's own (stale) pin, tests/test_environment.py:9-16, with "computer" for "opponent"
    pin = ref_loop.run_reference_substeps([((80, 100), "robot"), ("", "computer"), ((200, 234), "robot"),
                                           ("", "computer")])
    pin_human = ref_loop.run_reference_substeps([((80, 100), "robot"), ((350, 233), "human"),
                                                 ((200, 234), "robot"), ((380, 234), "human")])
    np.savez_compressed(os.path.join(GOLDEN, "substeps_24x40.npz"), agents=agents, kinds=kinds, ia=ia, axy=axy,
                        init_state=init, states=states, pin_computer=pin, pin_human=pin_human,
                        names=np.array(names), **_header(seed))


def gen_init_frame(seed=606, n=8, T=80):
    """Game.play from its very first call: frame 0 is move-only (game.py:118-130)."""
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 8, 2))
    ref = ref_loop.run_reference(actions, table, init_frame=True)
    np.savez_compressed(os.path.join(GOLDEN, "init_frame_8x80.npz"), actions=actions, reset_table=table,
                        states=ref["states"], scores=ref["scores"], obs_next=ref["obs_next"],
                        reward=ref["reward"].astype(np.int8), score=ref["score"], done=ref["done"], **_header(seed))


def gen_friction(seed=707, n=24, T=150):
    """Pins the opt-in friction flag to the reference's OWN (dead) ``Puck.friction_on_puck`` (environment/puck.py:73-88):
    (1) the function itself on a table of velocities around its +-1 thresholds, (2) a trajectory of the live reference
    with that function hooked in immediately before ``Puck.limit_puck_speed`` -- the place ``ah_config.friction`` and
    the oracle's ``friction`` switch apply it (every sub-update, airhockey.py:117)."""
    refshim.install()
    from environment.puck import Puck
    rng = np.random.default_rng(seed)
    v = np.concatenate([rng.uniform(-12, 12, (400, 2)),
                        np.array([[1.0, 1.0], [-1.0, -1.0], [1.5, -1.5], [-1.25, 1.25], [0.0, 0.0], [1.0000001, -1.0000001],
                                  [0.75, -0.75], [10.0, -10.0], [-10.5, 10.5]])])
    out = np.zeros_like(v)
    for i, (dx, dy) in enumerate(v):
        p = Puck.__new__(Puck)
        p.dx, p.dy = float(dx), float(dy)
        p.friction_on_puck()
        out[i] = (p.dx, p.dy)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 32, 2))
    init = scrambled_states(rng, n)
    original = Puck.limit_puck_speed

    def limit_with_friction(self):
        self.friction_on_puck()
        return original(self)

    Puck.limit_puck_speed = limit_with_friction
    try:
        ref = ref_loop.run_reference(actions, table, init_state=init)
    finally:
        Puck.limit_puck_speed = original
    np.savez_compressed(os.path.join(GOLDEN, "friction_24x150.npz"), fn_in=v, fn_out=out, actions=actions, reset_table=table,
                        init_state=init, states=ref["states"], scores=ref["scores"], reward=ref["reward"].astype(np.int8),
                        score=ref["score"], done=ref["done"],
                        game_over=(ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8),
                        reset_cursor=ref["reset_cursor"], **_header(seed))


def main():
    refshim.install()
    os.makedirs(GOLDEN, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "friction":      # add the friction fixture without regenerating the others
        gen_friction()
        return
    gen_canonical()
    gen_long()
    gen_scrambled()
    gen_continuous()
    gen_substeps()
    gen_init_frame()
    gen_friction()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
