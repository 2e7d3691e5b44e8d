"""The reference's own unit-test pins (/root/reference/tests/*.py), re-expressed against the oracle.
The same pins are asserted against the CUDA path in tests/test_gpu_api.py."""
import numpy as np

from oracle import oracle

L = oracle.lib


def test_puck_intersect_mallet():
    # tests/test_puck.py:18-35 : mallet at (350, 240); puck at offset (-10,-10) overlaps, (-50,-50) does not
    assert L().orc_puck_overlaps_mallet(340.0, 230.0, 350.0, 240.0)
    assert not L().orc_puck_overlaps_mallet(300.0, 190.0, 350.0, 240.0)


def test_puck_intersect_walls():
    # tests/test_puck.py:37-63
    assert L().orc_puck_hits_left_wall(10.0) and not L().orc_puck_hits_left_wall(100.0)
    assert L().orc_puck_hits_right_wall(890.0) and not L().orc_puck_hits_right_wall(800.0)


def test_goals():
    # tests/test_goal.py:14-40 : left goal (0,240), right goal (900,240)
    assert L().orc_puck_in_goal(0.0, 240.0, 10.0, 240.0) and not L().orc_puck_in_goal(0.0, 240.0, 10.0, 30.0)
    assert L().orc_puck_in_goal(900.0, 240.0, 900.0, 240.0) and not L().orc_puck_in_goal(900.0, 240.0, 900.0, 30.0)


def _upd(which, x, y):
    m = np.array([x, y, 0.0, 0.0])
    L().orc_mallet_update(which, m.ctypes.data)
    return m[0], m[1]


def test_mallet_limits():
    # tests/test_mallet.py:14-63
    assert _upd(0, -1, -1) == (20, 20) and _upd(0, 1000, 1000) == (880, 460)
    assert _upd(1, -1, -1) == (20, 20) and _upd(1, 1000, 1000) == (430, 460)
    assert _upd(2, -1, -1) == (470, 20) and _upd(2, 1000, 1000) == (880, 460)


def test_environment_pin(golden_dir):
    # tests/test_environment.py:9-16 expects puck.location() == (438, 240) after four sub-updates; the
    # test as written passes the stale agent name "opponent" (-> InvalidAgentError, airhockey.py:108-110),
    # with "computer" substituted the pin holds.
    st, sc = oracle.initial_state(1)
    oracle.update_state(st, sc, 0, continuous=np.array([[80.0, 100.0]]))
    oracle.update_state(st, sc, 2)
    oracle.update_state(st, sc, 0, continuous=np.array([[200.0, 234.0]]))
    oracle.update_state(st, sc, 2)
    assert (int(st[0, 0]), int(st[0, 1])) == (438, 240)
    st, sc = oracle.initial_state(1)
    assert oracle.update_state(st, sc, 3) == -1          # unknown agent -> error code


def test_initial_observation():
    # get_state("robot") on a fresh env
    st, _ = oracle.initial_state(1)
    assert oracle.get_state(st)[0].tolist() == [350, 240, 450, 240, 0, 0, -3, 0]
    assert oracle.get_state(st, agent_is_robot=False)[0].tolist() == [550, 240, 450, 240, 0, 0, -3, 0]


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() \
        == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---- the FP32 throughput build's reset generator (pcg3d keyed by splitmix64(seed); csrc/ah_rng.cuh, oracle/ah_oracle.c)
def _pcg3d_reset_pairs(seed, env_id, reset_idx):
    """Independent numpy restatement (uint32 wrap-around arithmetic), vectorised over env_id / reset_idx."""
    M = (1 << 64) - 1
    z = (seed + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z ^= z >> 31
    env_id = np.asarray(env_id, np.uint64)
    x = (env_id & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ np.uint32(z & 0xFFFFFFFF)
    y = np.asarray(reset_idx, np.uint32) ^ (env_id >> np.uint64(32)).astype(np.uint32)
    x, y = np.broadcast_arrays(x, y)
    zz = np.full(x.shape, z >> 32, np.uint32)
    with np.errstate(over="ignore"):
        a, c = np.uint32(1664525), np.uint32(1013904223)
        x = x * a + c; y = y * a + c; zz = zz * a + c
        x = x + y * zz; y = y + zz * x; zz = zz + x * y
        x = x ^ (x >> np.uint32(16)); y = y ^ (y >> np.uint32(16)); zz = zz ^ (zz >> np.uint32(16))
        x = x + y * zz; y = y + zz * x
    # 24-bit integer * 20 / 2^24 - 10: the product needs 29 bits, so a float64 evaluation rounded once to binary32 is the fma
    dx = ((x >> np.uint32(8)).astype(np.float64) * (20.0 / 16777216.0) - 10.0).astype(np.float32)
    dy = ((y >> np.uint32(8)).astype(np.float64) * (20.0 / 16777216.0) - 10.0).astype(np.float32)
    return dx, dy


def test_fp32_reset_generator_matches_its_restatement():
    for seed, env, idx in [(0, 0, 0), (5, 0, 1), (1234, 1048575, 7), (2 ** 64 - 1, 2 ** 40 + 3, 2 ** 32 - 1), (99, 8388607, 1000)]:
        dx, dy = _pcg3d_reset_pairs(seed, env, idx)
        got = oracle.rng_reset_pair_f32(seed, env, idx)
        assert got.tolist() == [float(dx), float(dy)], (seed, env, idx)
    # known answers (binary32 patterns), pinned when the generator was introduced
    kat = {(5, 0, 1): [0xbf302cdc, 0x40a1dae7], (1234, 1048575, 7): [0x3e5a79b0, 0xc10789bc],
           (2 ** 64 - 1, 2 ** 40 + 3, 2 ** 32 - 1): [0xc1053a5d, 0xbfa95ef2]}
    for args, want in kat.items():
        assert oracle.rng_reset_pair_f32(*args).astype(np.float32).view(np.uint32).tolist() == want, args


def test_fp32_reset_generator_statistics():
    """U(-10, 10) per component; no structure along the env axis, the reset axis, between components or between seeds."""
    env = np.arange(1 << 18, dtype=np.uint64)
    dx, dy = _pcg3d_reset_pairs(1234, env[:, None], np.arange(8, dtype=np.uint32)[None, :])      # [2^18, 8]
    for v in (dx, dy):
        assert v.min() >= -10.0 and v.max() < 10.0
        assert abs(v.mean()) < 0.02 and abs(v.var() - 100.0 / 3.0) < 0.1
        h = np.bincount(((v.astype(np.float64) + 10.0) * (256 / 20.0)).astype(np.int64).ravel(), minlength=256)
        chi2 = ((h - v.size / 256) ** 2 / (v.size / 256)).sum()
        assert chi2 < 255 + 6 * np.sqrt(2 * 255), chi2                                       # 255 dof, 6 sigma

    def corr(a, b):
        a = a.astype(np.float64).ravel(); b = b.astype(np.float64).ravel()
        return abs(np.corrcoef(a, b)[0, 1])
    lim = 6.0 / np.sqrt(dx.size)
    assert corr(dx, dy) < lim                                   # the two components of a draw
    assert corr(dx[:-1], dx[1:]) < lim and corr(dy[:-1], dy[1:]) < lim            # neighbouring envs
    assert corr(dx[:, :-1], dx[:, 1:]) < lim and corr(dy[:, :-1], dy[:, 1:]) < lim    # consecutive resets of an env
    ex, ey = _pcg3d_reset_pairs(1235, env[:, None], np.arange(8, dtype=np.uint32)[None, :])   # the next seed
    assert corr(dx, ex) < lim and corr(dy, ey) < lim
    both = np.stack([dx.ravel(), dy.ravel()], 1).view(np.uint64).ravel()
    assert np.unique(both).size > both.size - 8                # 48-bit draws: (almost) no repeated pair in 2^21


def test_oracle_frame_loop_uses_the_selected_reset_generator():
    """run_frames(reset_rng=...) draws table-less resets from the generator it names: put the puck in the robot's goal
    mouth, play one frame with no nudge, the puck's velocity is the first draw of that env's stream."""
    for rng_name, pair in (("philox", oracle.rng_reset_pair), ("f32", oracle.rng_reset_pair_f32)):
        st, sc = oracle.initial_state(3)
        st[:, 0] = 20.0; st[:, 2:4] = 0.0
        cur = np.array([0, 4, 9], np.int32)
        out = oracle.run_frames(st, sc, 1, actions=np.full((1, 3), 4, np.int8), seed=321, env_id0=1000, reset_cursor=cur,
                                reset_rng=rng_name, want=("score",))
        assert out["score"].tolist() == [[-1, -1, -1]] and sc[:, 1].tolist() == [1, 1, 1]
        assert out["reset_cursor"].tolist() == [1, 5, 10]
        for i, c in enumerate((0, 4, 9)):
            assert st[i, 2:4].tolist() == pair(321, 1000 + i, c).tolist(), (rng_name, i)
