"""LIVE differential test: the C oracle against the Python reference imported from /root/reference.
Skipped where the reference is absent (the GPU box).  AH_DIFF_ENVS / AH_DIFF_FRAMES scale it up
(4096 x 1000 = BASELINE config 2 takes ~6 CPU-minutes)."""
import os

import numpy as np
import pytest

from oracle import oracle, refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="live reference not present")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def test_numpy_dot_probe_matches_oracle_default():
    assert int(refshim.numpy_dot_is_fma(4000)) == oracle.DOT_FMA_DEFAULT


@pytest.mark.parametrize("seed", [1, 2])
def test_live_canonical(seed):
    from oracle import ref_loop
    n = int(os.environ.get("AH_DIFF_ENVS", 24))
    T = int(os.environ.get("AH_DIFF_FRAMES", 400))
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 64, 2))
    ref = ref_loop.run_reference(actions, table)
    st, sc = oracle.initial_state(n)
    out = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=1)
    assert np.array_equal(bits(ref["states"]), bits(out["states_rec"]))
    assert np.array_equal(ref["scores"], out["scores_rec"])
    assert np.array_equal(ref["reward"], out["reward"])
    assert np.array_equal(ref["score"], out["score"]) and np.array_equal(ref["done"], out["done"])
    assert np.array_equal(ref["robot_win"] + 2 * ref["opponent_win"], out["game_over"])
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(ref[k]), bits(out[k]))


def test_live_scrambled():
    from oracle import gen_golden, ref_loop
    rng = np.random.default_rng(77)
    n, T = 48, 120
    init = gen_golden.scrambled_states(rng, n)
    init_scores = rng.integers(7, 10, (n, 2)).astype(np.int32)
    actions = rng.integers(-2, 7, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 32, 2))
    ref = ref_loop.run_reference(actions, table, init_state=init, init_scores=init_scores)
    st, sc = init.copy(), init_scores.copy()
    out = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=1)
    assert np.array_equal(bits(ref["states"]), bits(out["states_rec"]))
    assert np.array_equal(ref["scores"], out["scores_rec"])
    assert np.array_equal(ref["reward"], out["reward"])
    assert np.array_equal(ref["robot_win"] + 2 * ref["opponent_win"], out["game_over"])
