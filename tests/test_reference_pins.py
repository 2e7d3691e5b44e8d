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
