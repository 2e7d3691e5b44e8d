"""TEST INFRASTRUCTURE ONLY.  ctypes front-end of the C oracle (oracle/ah_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this module.  The product package (air_hockey_training_environment_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import build as _build

_lib = None

# np.dot FMA behaviour of the numpy build that generated tests/golden (probed by refshim.numpy_dot_is_fma)
DOT_FMA_DEFAULT = 1

STAT_NAMES = ("frames", "robot_goals", "opponent_goals", "robot_wins", "opponent_wins", "dones",
              "reward_sum", "reward_sq_sum", "_8", "_9", "_10", "_11", "_12", "_13", "contacts_robot", "contacts_opponent",
              "done_ambiguous", "done_pow_vs_mul_disagree", "_18", "_19")


class _Outputs(C.Structure):
    _fields_ = [("obs_state", C.c_void_p), ("obs_new_state", C.c_void_p), ("obs_next", C.c_void_p),
                ("reward", C.c_void_p), ("score", C.c_void_p), ("done", C.c_void_p), ("game_over", C.c_void_p),
                ("states_rec", C.c_void_p), ("scores_rec", C.c_void_p), ("stats", C.c_void_p),
                ("contacts", C.c_void_p), ("contact_dmin", C.c_void_p)]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path) or (os.path.exists(_build.SRC) and
                                        os.path.getmtime(path) < os.path.getmtime(_build.SRC)):
            path = _build.build()
        _lib = C.CDLL(path)
        _lib.orc_frames.restype = C.c_int
        _lib.orc_frames.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Outputs)]
        _lib.orc_frames_ex.restype = C.c_int
        _lib.orc_frames_ex.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Outputs)]
        _lib.orc_rng_reset_pair_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        _lib.orc_init.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        _lib.orc_update_state.restype = C.c_int
        _lib.orc_update_state.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int]
        _lib.orc_get_state.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_rng_reset_pair.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        _lib.orc_rng_action.restype = C.c_int
        _lib.orc_rng_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        for name in ("orc_puck_overlaps_mallet",):
            getattr(_lib, name).restype = C.c_int
            getattr(_lib, name).argtypes = [C.c_double] * 4
        for name in ("orc_puck_hits_left_wall", "orc_puck_hits_right_wall"):
            getattr(_lib, name).restype = C.c_int
            getattr(_lib, name).argtypes = [C.c_double]
        _lib.orc_puck_in_goal.restype = C.c_int
        _lib.orc_puck_in_goal.argtypes = [C.c_double] * 4
        _lib.orc_mallet_update.argtypes = [C.c_int, C.c_void_p]
        _lib.orc_friction_on_puck.argtypes = [C.c_void_p]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def initial_state(n: int):
    """Reference initial state (airhockey.py:25-63): returns (state[n,13] f64, scores[n,2] i32)."""
    state = np.empty((n, 13), np.float64)
    scores = np.empty((n, 2), np.int32)
    lib().orc_init(n, _p(state), _p(scores))
    return state, scores


def run_frames(state: np.ndarray, scores: np.ndarray, T: int, *,
               actions: Optional[np.ndarray] = None,
               reset_table: Optional[np.ndarray] = None,
               reset_cursor: Optional[np.ndarray] = None,
               seed: int = 0, env_id0: int = 0, frame0: int = 0,
               dot_fma: int = DOT_FMA_DEFAULT, friction: bool = False,
               record_every: int = 0, nthreads: int = 1, reset_rng: str = "philox",
               want=("obs_state", "obs_new_state", "obs_next", "reward", "score", "done", "game_over"),
               ) -> Dict[str, np.ndarray]:
    """Advance ``state``/``scores`` IN PLACE by T frames; returns the per-frame outputs asked for.

    actions: int8 [T,n] (discrete), float64 [T,n,2] (continuous) or None (Philox actions).
    reset_table: float64 [n,R,2] or None (in-kernel resets keyed (seed, env_id0 + i): ``reset_rng`` "philox" restates the
    FP64 build's Philox4x32-10 draw, "f32" the FP32 throughput build's pcg3d draw -- csrc/ah_rng.cuh).
    """
    n = state.shape[0]
    assert state.dtype == np.float64 and state.flags.c_contiguous and state.shape == (n, 13)
    assert scores.dtype == np.int32 and scores.flags.c_contiguous and scores.shape == (n, 2)
    a_i8 = a_f64 = None
    if actions is not None:
        if actions.ndim == 3:
            a_f64 = np.ascontiguousarray(actions, np.float64)
            assert a_f64.shape == (T, n, 2)
        else:
            a_i8 = np.ascontiguousarray(actions, np.int8)
            assert a_i8.shape == (T, n)
    R = 0
    if reset_table is not None:
        reset_table = np.ascontiguousarray(reset_table, np.float64)
        assert reset_table.shape[0] == n and reset_table.shape[2] == 2
        R = reset_table.shape[1]
    if reset_cursor is None:
        reset_cursor = np.zeros(n, np.int32)
    assert reset_cursor.dtype == np.int32 and reset_cursor.shape == (n,)
    out: Dict[str, np.ndarray] = {}
    shapes = {"obs_state": ((T, n, 8), np.float64), "obs_new_state": ((T, n, 8), np.float64),
              "obs_next": ((T, n, 8), np.float64), "reward": ((T, n), np.int32), "score": ((T, n), np.int8),
              "done": ((T, n), np.uint8), "game_over": ((T, n), np.uint8),
              "contacts": ((T, n), np.uint8), "contact_dmin": ((T, n), np.float64)}
    for k in want:
        out[k] = np.zeros(*shapes[k])
    if record_every:
        out["states_rec"] = np.zeros((T // record_every, n, 13), np.float64)
        out["scores_rec"] = np.zeros((T // record_every, n, 2), np.int32)
    out["stats"] = np.zeros(20, np.int64)
    o = _Outputs(**{k: _p(out.get(k)) for k, _ in _Outputs._fields_})
    rc = lib().orc_frames_ex(n, T, _p(state), _p(scores), _p(a_i8), _p(a_f64), _p(reset_table), R,
                             _p(reset_cursor), seed, env_id0, frame0, int(dot_fma), int(friction),
                             int(record_every), int(nthreads), {"philox": 0, "f32": 1}[reset_rng], C.byref(o))
    out["reset_cursor"] = reset_cursor
    out["reset_overflow"] = bool(rc)
    return out


def update_state(state: np.ndarray, scores: np.ndarray, agent: int, *, discrete: Optional[np.ndarray] = None,
                 continuous: Optional[np.ndarray] = None, dot_fma: int = DOT_FMA_DEFAULT,
                 friction: bool = False) -> int:
    """One ``AirHockey.update_state`` sub-update in place.  agent: 0 robot, 1 human, 2 computer."""
    n = state.shape[0]
    ia = None if discrete is None else np.ascontiguousarray(discrete, np.int32)
    axy = None if continuous is None else np.ascontiguousarray(continuous, np.float64)
    kind = 1 if (continuous is not None and agent == 0) else 0
    return lib().orc_update_state(n, _p(state), _p(scores), agent, kind, _p(ia), _p(axy), int(dot_fma), int(friction))


def get_state(state: np.ndarray, agent_is_robot: bool = True) -> np.ndarray:
    n = state.shape[0]
    obs = np.empty((n, 8), np.float64)
    lib().orc_get_state(n, _p(state), int(agent_is_robot), _p(obs))
    return obs


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def rng_reset_pair(seed: int, env_id: int, reset_idx: int) -> np.ndarray:
    o = np.zeros(2, np.float64)
    lib().orc_rng_reset_pair(seed, env_id, reset_idx, _p(o))
    return o


def rng_reset_pair_f32(seed: int, env_id: int, reset_idx: int) -> np.ndarray:
    """The FP32 build's reset draw (pcg3d; binary32 values, exact in the returned float64)."""
    o = np.zeros(2, np.float64)
    lib().orc_rng_reset_pair_f32(seed, env_id, reset_idx, _p(o))
    return o


def rng_action(seed: int, env_id: int, frame: int) -> int:
    return lib().orc_rng_action(seed, env_id, frame)


def friction_on_puck(dx: float, dy: float):
    """``Puck.friction_on_puck`` (environment/puck.py:73-88) on one velocity."""
    v = np.array([dx, dy], np.float64)
    lib().orc_friction_on_puck(_p(v))
    return float(v[0]), float(v[1])
