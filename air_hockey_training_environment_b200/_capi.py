"""ctypes binding of the C ABI declared in ``include/ah_b200.h`` (library: ``libah_b200.so``, built in-tree
by ``build.py``).  There is no CPU fallback: a missing library is a hard error, and every compute call on a
machine without a CUDA device fails with ``AH_ECUDA``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AH_B200_LIB", os.path.join(HERE, "libah_b200.so"))   # override: tuning experiments only

AH_OK, AH_EINVAL, AH_ECUDA, AH_ENOMEM, AH_EAGENT, AH_ESTATE = 0, -1, -2, -3, -4, -5
AH_F32, AH_F64 = 0, 1
AH_AGENT_ROBOT, AH_AGENT_HUMAN, AH_AGENT_COMPUTER = 0, 1, 2
(AH_ACT_NONE, AH_ACT_DISCRETE, AH_ACT_CONTINUOUS, AH_ACT_PHILOX, AH_ACT_DISCRETE_REPEAT, AH_ACT_CONTINUOUS_REPEAT,
 AH_ACT_POLICY, AH_ACT_DISCRETE_PACKED4, AH_ACT_POLICY_MLP) = range(9)
AH_NSTATS = 20
AH_ACTOR_NWEIGHTS = 114
STAT_NAMES = ("frames", "robot_goals", "opponent_goals", "robot_wins", "opponent_wins", "dones", "reward_sum",
              "reward_sq_sum", "games", "game_len_sum", "game_len_sq_sum", "game_return_sum", "nonfinite",
              "reset_table_overflow", "contacts_robot", "contacts_opponent", "done_ambiguous", "_17", "_18", "_19")

# every symbol include/ah_b200.h declares (tests/test_capi_symbols.py checks header <-> library <-> this list)
SYMBOLS = ("ah_strerror", "ah_abi_version", "ah_last_cuda_error", "ah_create", "ah_destroy", "ah_bind_state",
           "ah_init_state", "ah_reset", "ah_update_state", "ah_get_state", "ah_update_score", "ah_rewards", "ah_step", "ah_pack_host", "ah_actor_sample", "ah_policy_act", "ah_c51_project", "ah_stats", "ah_stats_bank", "ah_peer_export", "ah_peer_attach", "ah_stats_allreduce",
           "ah_set_counters", "ah_get_counters", "ah_ring_sample", "ah_launch_info", "ah_selftest_nofma")


class AhConfig(C.Structure):
    _fields_ = [("dot_fma", C.c_int32), ("friction", C.c_int32), ("no_prefetch", C.c_int32), ("no_packed_core", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class AhRing(C.Structure):
    _fields_ = [("state", C.c_void_p), ("new_state", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p),
                ("done", C.c_void_p), ("capacity", C.c_int64), ("appended", C.c_void_p),
                ("format", C.c_int32), ("reserved0", C.c_int32), ("state_pos", C.c_void_p), ("new_state_pos", C.c_void_p)]


AH_RING_ROWS, AH_RING_COMPACT = 0, 1
AH_RING_COUNTERS = 4


AH_ACTV_LINEAR, AH_ACTV_RELU, AH_ACTV_TANH = 0, 1, 2
AH_HEAD_SOFTMAX, AH_HEAD_VALUES, AH_HEAD_DUELING, AH_HEAD_CONTINUOUS = 0, 1, 2, 3
(AH_SELECT_SAMPLE, AH_SELECT_GREEDY, AH_SELECT_EPS_GREEDY, AH_SELECT_BOLTZMANN, AH_SELECT_MAX_BOLTZMANN, AH_SELECT_GAUSSIAN,
 AH_SELECT_OU) = range(7)
AH_POLICY_MAX_LAYERS, AH_POLICY_MAX_WIDTH = 4, 32
AH_MAX_PEERS, AH_IPC_HANDLE_BYTES = 8, 64


class AhPolicy(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("width", C.c_int32 * 5), ("activation", C.c_int32 * 4),
                ("head", C.c_int32), ("select", C.c_int32), ("n_actions", C.c_int32),
                ("epsilon", C.c_float), ("initial_epsilon", C.c_float), ("final_epsilon", C.c_float),
                ("observe", C.c_int32), ("explore", C.c_int32),
                ("tau", C.c_float), ("clip_lo", C.c_float), ("clip_hi", C.c_float),
                ("mu", C.c_float), ("sigma", C.c_float), ("sigma_min", C.c_float), ("n_steps_annealing", C.c_int32),
                ("theta", C.c_float), ("dt", C.c_float),
                ("weights", C.c_void_p), ("noise_state", C.c_void_p)]


class AhStepIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("action_kind", C.c_int32), ("reset_table_len", C.c_int32),
                ("reset_table", C.c_void_p),
                ("obs_state", C.c_void_p), ("obs_new_state", C.c_void_p), ("reward", C.c_void_p),
                ("done", C.c_void_p), ("score", C.c_void_p), ("game_over", C.c_void_p),
                ("obs_next", C.c_void_p), ("obs_next_per_frame", C.c_int32), ("collect_stats", C.c_int32),
                ("ring", C.POINTER(AhRing)), ("actions_out", C.c_void_p), ("probs_out", C.c_void_p),
                ("events", C.c_void_p), ("policy", C.POINTER(AhPolicy)), ("actions_real_out", C.c_void_p),
                ("env_first", C.c_int64), ("env_count", C.c_int64),
                ("stats_bank", C.c_int32), ("ring_counter_first", C.c_int32), ("ring_counter_count", C.c_int32),
                ("reserved0", C.c_int32)]


class AhError(RuntimeError):
    def __init__(self, code: int, where: str, cuda_error: int = 0):
        self.code = code
        self.cuda_error = cuda_error
        msg = f"{where}: {strerror(code)} ({code})"
        if code == AH_ECUDA:
            msg += f" [cudaError_t={cuda_error}]"
        super().__init__(msg)


_lib = None


def lib() -> C.CDLL:
    """Load libah_b200.so.  Fails loudly if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m air_hockey_training_environment_b200.build` "
                          "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.ah_strerror.restype = C.c_char_p
    L.ah_strerror.argtypes = [i32]
    L.ah_abi_version.restype = i32
    L.ah_abi_version.argtypes = []
    L.ah_last_cuda_error.restype = i32
    L.ah_last_cuda_error.argtypes = [vp]
    L.ah_create.restype = i32
    L.ah_create.argtypes = [C.POINTER(vp), i64, i32, i32, u64, u64, C.POINTER(AhConfig)]
    L.ah_destroy.restype = i32
    L.ah_destroy.argtypes = [vp]
    L.ah_bind_state.restype = i32
    L.ah_bind_state.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ah_init_state.restype = i32
    L.ah_init_state.argtypes = [vp, vp]
    L.ah_reset.restype = i32
    L.ah_reset.argtypes = [vp, i32, vp, vp, i32, vp]
    L.ah_update_state.restype = i32
    L.ah_update_state.argtypes = [vp, vp, i32, i32, vp]
    L.ah_get_state.restype = i32
    L.ah_get_state.argtypes = [vp, i32, vp, vp]
    L.ah_update_score.restype = i32
    L.ah_update_score.argtypes = [vp, vp, vp, i32, vp]
    L.ah_rewards.restype = i32
    L.ah_rewards.argtypes = [vp, vp, vp, vp, vp]
    L.ah_step.restype = i32
    L.ah_step.argtypes = [vp, C.POINTER(AhStepIO), i32, vp]
    L.ah_pack_host.restype = i32
    L.ah_pack_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ah_actor_sample.restype = i32
    L.ah_actor_sample.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.ah_policy_act.restype = i32
    L.ah_policy_act.argtypes = [vp, C.POINTER(AhPolicy), vp, vp, vp, vp, vp, vp]
    L.ah_c51_project.restype = i32
    L.ah_c51_project.argtypes = [vp, i64, i32, i32, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]
    L.ah_stats.restype = i32
    L.ah_stats.argtypes = [vp, vp, i32, vp]
    L.ah_stats_bank.restype = i32
    L.ah_stats_bank.argtypes = [vp, i32, vp, i32, vp]
    L.ah_peer_export.restype = i32
    L.ah_peer_export.argtypes = [vp, vp]
    L.ah_peer_attach.restype = i32
    L.ah_peer_attach.argtypes = [vp, i32, i32, vp]
    L.ah_stats_allreduce.restype = i32
    L.ah_stats_allreduce.argtypes = [vp, i32, vp, i32, vp]
    L.ah_set_counters.restype = i32
    L.ah_set_counters.argtypes = [vp, u64, vp]
    L.ah_get_counters.restype = i32
    L.ah_get_counters.argtypes = [vp, vp, vp]
    L.ah_ring_sample.restype = i32
    L.ah_ring_sample.argtypes = [vp, C.POINTER(AhRing), i32, i64, u64, vp, vp, vp, vp, vp, vp, vp]
    L.ah_launch_info.restype = i32
    L.ah_launch_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.ah_selftest_nofma.restype = i32
    L.ah_selftest_nofma.argtypes = [vp, vp, vp]
    _lib = L
    return L


def strerror(code: int) -> str:
    return lib().ah_strerror(code).decode()


def check(code: int, where: str, handle=None) -> None:
    if code != AH_OK:
        cuda = lib().ah_last_cuda_error(handle) if handle else 0
        raise AhError(code, where, cuda)
