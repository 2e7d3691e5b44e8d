"""Host side of the batched environment: the reference's ``AirHockey`` verbs, vectorised over N games.

``BatchedAirHockey`` owns the state as torch CUDA tensors (structure of arrays, see include/ah_b200.h) and
forwards every verb to the C ABI with raw device pointers and the current CUDA stream.  PyTorch is used for
device memory and streams only; all arithmetic happens in the sm_100a kernels of ``csrc/ah_b200.cu``.

Reference interface mirrored here (paths relative to the reference root):
    AirHockey.__init__        environment/airhockey.py:25-63
    AirHockey.update_state    environment/airhockey.py:96-134   (InvalidAgentError for unknown agent names)
    AirHockey.get_state       environment/airhockey.py:136-147  (+ serialize_state, lib/utils/helpers.py:72-81)
    AirHockey.update_score    environment/airhockey.py:149-169
    AirHockey.reset           environment/airhockey.py:171-187
    Rewards.__call__          lib/rewards.py:118-142
    one frame of Game.play    game.py:133-177 + lib/agents/agent.py:57-78          -> ``step``
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Union

import torch

from . import _capi
from .exceptions import InvalidAgentError

_AGENTS = {"robot": _capi.AH_AGENT_ROBOT, "human": _capi.AH_AGENT_HUMAN, "computer": _capi.AH_AGENT_COMPUTER}

STEP_FIELDS = ("obs_state", "obs_new_state", "reward", "done", "score", "game_over", "obs_next", "actions", "probs", "events",
               "actions_real")


def pack_actions(actions: torch.Tensor) -> torch.Tensor:
    """int8/uint8 [K, N] per-frame discrete actions -> uint8 [ceil(K/4), N, 4] "quad-major" (AH_ACT_DISCRETE_PACKED4):
    byte j of word (q, i) is env i's action of frame 4q + j.  The input format of the packed rollout mode."""
    K, n = actions.shape
    q = (K + 3) // 4
    a = actions.to(torch.uint8)
    if q * 4 != K:
        a = torch.cat([a, torch.zeros(q * 4 - K, n, dtype=torch.uint8, device=a.device)], 0)
    return a.view(q, 4, n).permute(0, 2, 1).contiguous()


def unpack_events(events: torch.Tensor, K: int) -> Dict[str, torch.Tensor]:
    """uint8 [ceil(K/4), N, 4] event bytes (AH_EVENT_* in include/ah_b200.h) -> the reference's per-frame outputs,
    time-major [K, N]: reward (int8, lib/rewards.py:118-142), done (bool), score (int8), game_over (uint8)."""
    q, n, _ = events.shape
    b = events.permute(0, 2, 1).reshape(q * 4, n)[:K]
    return {"reward": ((b & 7).to(torch.int8) << 1) - 4, "done": ((b >> 3) & 1).to(torch.bool),
            "score": ((b >> 4) & 3).to(torch.int8) - 1, "game_over": (b >> 6) & 3}


@dataclass
class StepOutputs:
    """Per-frame outputs of ``BatchedAirHockey.step``; any field may be None (not produced)."""
    obs_state: Optional[torch.Tensor] = None       # real [K,N,8]  Observation.state      (agent.py:61)
    obs_new_state: Optional[torch.Tensor] = None   # real [K,N,8]  Observation.new_state  (agent.py:67)
    reward: Optional[torch.Tensor] = None          # real [K,N]
    done: Optional[torch.Tensor] = None            # u8   [K,N]
    score: Optional[torch.Tensor] = None           # i8   [K,N]
    game_over: Optional[torch.Tensor] = None       # u8   [K,N]    0 / 1 robot won / 2 opponent won
    obs_next: Optional[torch.Tensor] = None        # real [N,8] or [K,N,8]: the next frame's policy input
    actions: Optional[torch.Tensor] = None         # i8   [K,N]    the discrete actions applied (Philox / in-kernel policy)
    probs: Optional[torch.Tensor] = None           # real [K,N,4]  the in-kernel actor's distribution (policy mode only)
    actions_real: Optional[torch.Tensor] = None    # real [K,N,2]  the continuous actions a device policy chose
    events: Optional[torch.Tensor] = None          # u8   [ceil(K/4),N,4]  packed rollout mode: reward, done, score and
                                                   #               game_over of a frame in ONE byte (see unpack_events)


@dataclass
class StepPlan:
    """A validated, frozen ``ah_step`` call (``BatchedAirHockey.plan_step`` / ``launch``)."""
    io: object
    io_ref: object
    K: int
    out: "StepOutputs"
    keepalive: tuple


class ReplayRing:
    """Device-resident replacement of ``MemoryBuffer`` (lib/buffer.py:11-66) for ``Observation`` tuples
    (lib/types.py:7), filled from inside the step kernel.  Layout ``[capacity, N, ...]``; like the reference's
    ``appendleft`` + ``retreive`` the newest entry comes first when read back.

    The ring owns its append counter as a device tensor (``counter``, uint64 as int64[AH_RING_COUNTERS], equal copies): the step kernel reads the
    write slot from it and advances it, so it is correct under CUDA-graph replay, and ``purge()`` is a stream-ordered
    ``counter.zero_()``.  ``len(ring)`` / ``order()`` / ``retreive()`` read it back (one host sync); ``sample`` does not."""

    FIELDS = ("state", "action", "reward", "done", "new_state")

    def __init__(self, capacity: int, n_envs: int, dtype: torch.dtype, device, continuous: bool = False,
                 default: Optional[Dict[str, float]] = None, compact: bool = False):
        """compact: the AH_RING_COMPACT layout (include/ah_b200.h) -- 51 instead of 70 bytes per tuple, written by the packed
        FP32 rollout kernel only: the observations' int()-truncated locations as int16 ``state_pos`` / ``new_state_pos``
        [capacity, N, 4], their velocities as float32 ``state`` / ``new_state`` [capacity, N, 4], ``reward`` as int8.
        ``retreive`` and ``sample`` return the same tensors either way."""
        self.capacity, self.n = int(capacity), int(n_envs)
        self.continuous, self.compact, self.dtype = continuous, bool(compact), dtype
        if compact and (dtype != torch.float32 or continuous):
            raise ValueError("the compact ring layout holds float32 discrete-action tuples")
        obs_w = 4 if compact else 8
        self.state = torch.zeros(capacity, n_envs, obs_w, dtype=dtype, device=device)
        self.new_state = torch.zeros(capacity, n_envs, obs_w, dtype=dtype, device=device)
        self.action = (torch.zeros(capacity, n_envs, 2, dtype=dtype, device=device) if continuous
                       else torch.zeros(capacity, n_envs, dtype=torch.int8, device=device))
        self.reward = torch.zeros(capacity, n_envs, dtype=torch.int8 if compact else dtype, device=device)
        self.done = torch.zeros(capacity, n_envs, dtype=torch.uint8, device=device)
        # AH_RING_COUNTERS equal copies of the append count: a step issued as concurrent env slices (InterleavedStepper)
        # gives every slice its own copy to read and advance; counter[0] is the one the readers use
        self.counter = torch.zeros(_capi.AH_RING_COUNTERS, dtype=torch.int64, device=device)
        self._c = _capi.AhRing(self.state.data_ptr(), self.new_state.data_ptr(), self.action.data_ptr(),
                               self.reward.data_ptr(), self.done.data_ptr(), self.capacity, self.counter.data_ptr())
        self._arrays = ["state", "new_state", "action", "reward", "done", "counter"]
        if compact:
            self.state_pos = torch.zeros(capacity, n_envs, 4, dtype=torch.int16, device=device)
            self.new_state_pos = torch.zeros(capacity, n_envs, 4, dtype=torch.int16, device=device)
            self._c.format = _capi.AH_RING_COMPACT
            self._c.state_pos, self._c.new_state_pos = self.state_pos.data_ptr(), self.new_state_pos.data_ptr()
            self._arrays += ["state_pos", "new_state_pos"]
        self._draw = 0
        if default:                                    # lib/buffer.py:19-22: a buffer full of default entries
            self.fill(default)

    def _put(self, field: str, index, value) -> None:
        """Write ``value`` (a scalar or a tensor in the PUBLIC shape of ``field``) at ``index`` of the ring's arrays."""
        dst = getattr(self, field)
        if self.compact and field in ("state", "new_state"):
            pos = getattr(self, field + "_pos")
            if torch.is_tensor(value):
                value = value.to(dst.device)
                pos[index] = torch.trunc(value[..., :4]).to(torch.int16)
                dst[index] = value[..., 4:].to(dst.dtype)
            else:
                pos[index] = int(value)
                dst[index] = value
        else:
            dst[index] = value.to(device=dst.device, dtype=dst.dtype) if torch.is_tensor(value) else value

    def _get(self, field: str, index) -> torch.Tensor:
        """Rows ``index`` of ``field`` in its public shape and dtype (the compact layout decoded)."""
        v = getattr(self, field)[index]
        if self.compact and field in ("state", "new_state"):
            return torch.cat([getattr(self, field + "_pos")[index].to(v.dtype), v], dim=-1)
        if self.compact and field == "reward":
            return v.to(self.dtype)
        return v

    def fill(self, default: Dict[str, float]) -> None:
        """Fill every slot with a default tuple (``MemoryBuffer(capacity, default)``, lib/buffer.py:19-22)."""
        for k in self.FIELDS:
            self._put(k, slice(None), default.get(k, 0))
        self.counter.fill_(self.capacity)

    def append(self, state: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor,
               new_state: torch.Tensor) -> None:
        """``MemoryBuffer.append(Observation(...))`` (lib/buffer.py:24-32) from device tensors ([N, ...], one tuple per
        env) -- the host-side verb; rollouts append from inside the step kernel instead (``step(ring=...)``).
        Stream-ordered, no host sync: the slot is computed on the device from the ring's counter."""
        slot = self.counter[:1] % self.capacity
        for k, v in zip(self.FIELDS, (state, action, reward, done, new_state)):
            if self.compact and k in ("state", "new_state"):
                v = v.to(self.state.device)
                getattr(self, k + "_pos").index_copy_(0, slot, torch.trunc(v[..., :4]).to(torch.int16).reshape(1, self.n, 4))
                getattr(self, k).index_copy_(0, slot, v[..., 4:].to(self.dtype).reshape(1, self.n, 4))
                continue
            dst = getattr(self, k)
            dst.index_copy_(0, slot, v.to(dst.dtype).reshape((1,) + tuple(dst.shape[1:])))
        self.counter += 1

    @property
    def appended(self) -> int:
        """Observation tuples appended since the last purge (reads the device counter)."""
        return int(self.counter[0].item())

    def __len__(self) -> int:      # lib/buffer.py:63-66
        return min(self.appended, self.capacity)

    def order(self) -> torch.Tensor:
        """Slot indices newest-first, the order of ``MemoryBuffer.retreive`` (lib/buffer.py:34-45)."""
        appended = self.appended
        n = min(appended, self.capacity)
        newest = (appended - 1) % self.capacity
        return (newest - torch.arange(n, device=self.state.device)) % self.capacity

    def retreive(self, average: bool = False) -> Dict[str, torch.Tensor]:
        """``MemoryBuffer.retreive`` (lib/buffer.py:34-45): the held tuples, newest first, as ``[len, N, ...]`` tensors;
        ``average=True`` returns the one-entry column average ``int(sum(col) / len)`` (truncated, :37-38)."""
        idx = self.order()
        got = {k: self._get(k, idx) for k in self.FIELDS}
        if average:
            got = {k: torch.trunc(v.to(torch.float64).sum(0, keepdim=True) / max(1, idx.numel())).to(v.dtype)
                   for k, v in got.items()}
        return got

    def sample(self, batch_size: int, env: "BatchedAirHockey", draw: Optional[int] = None, check: bool = True,
               out: Optional[Dict[str, torch.Tensor]] = None, with_index: bool = False) -> Dict[str, torch.Tensor]:
        """``MemoryBuffer.sample(batch_size)`` (lib/buffer.py:47-50): uniform WITHOUT replacement over the held
        (tuple, env) pairs, in one kernel (``ah_ring_sample``: a keyed bijection of the population, no permutation is
        materialised).  ``draw`` numbers the sample (default: a counter that advances per call); ``check`` reads the
        device counter to raise what ``random.sample`` raises for an over-sized batch (skip it inside CUDA graphs)."""
        if check and batch_size > len(self) * self.n:
            raise ValueError("Sample larger than population or is negative")
        if draw is None:
            draw, self._draw = self._draw, self._draw + 1
        dev, dt, B = self.state.device, self.dtype, int(batch_size)
        if out is None:
            out = {"state": torch.empty(B, 8, dtype=dt, device=dev), "new_state": torch.empty(B, 8, dtype=dt, device=dev),
                   "action": (torch.empty(B, 2, dtype=dt, device=dev) if self.continuous
                              else torch.empty(B, dtype=torch.int8, device=dev)),
                   "reward": torch.empty(B, dtype=dt, device=dev), "done": torch.empty(B, dtype=torch.uint8, device=dev)}
            if with_index:
                out["index"] = torch.empty(B, dtype=torch.int64, device=dev)
        idx = out.get("index")
        _capi.check(env._lib.ah_ring_sample(env._h, C.byref(self._c), int(self.continuous), B, int(draw),
                                            out["state"].data_ptr(), out["action"].data_ptr(), out["reward"].data_ptr(),
                                            out["done"].data_ptr(), out["new_state"].data_ptr(),
                                            None if idx is None else idx.data_ptr(), env._stream()), "ah_ring_sample", env._h)
        return out

    def purge(self, default: Optional[Dict[str, float]] = None) -> None:       # lib/buffer.py:52-61
        self.counter.zero_()
        if default:
            self.fill(default)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, k).clone() for k in self._arrays}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k in self._arrays:
            getattr(self, k).copy_(sd[k])


class BatchedAirHockey:
    """N independent air-hockey games stepped on one GPU."""

    def __init__(self, n_envs: int, dtype: torch.dtype = torch.float32, device: Union[str, torch.device] = "cuda",
                 seed: int = 0, env_id0: int = 0, dot_fma: bool = True, friction: bool = False, prefetch: bool = True,
                 packed_core: bool = True):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 (throughput build) or torch.float64 (validation build)")
        self._lib = _capi.lib()                       # raises ImportError when the CUDA library is missing
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("BatchedAirHockey runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n, self.dtype, self.seed, self.env_id0 = int(n_envs), dtype, int(seed), int(env_id0)
        self._h = C.c_void_p()
        cfg = _capi.AhConfig(int(dot_fma), int(friction), int(not prefetch), int(not packed_core))
        _capi.check(self._lib.ah_create(C.byref(self._h), self.n, _capi.AH_F32 if dtype == torch.float32 else _capi.AH_F64,
                                        self.device.index, self.seed, self.env_id0, C.byref(cfg)), "ah_create")
        kw = dict(dtype=dtype, device=self.device)
        self.puck = torch.empty(self.n, 4, **kw)          # x, y, dx, dy
        self.robot = torch.empty(self.n, 4, **kw)
        self.opponent = torch.empty(self.n, 4, **kw)
        self.prev_dist = torch.empty(self.n, **kw)
        self.meta = torch.zeros(self.n, 4, dtype=torch.int32, device=self.device)
        self._bind()
        self._stats_buf = torch.zeros(_capi.AH_NSTATS, dtype=torch.int64, device=self.device)
        self._io_keepalive = None
        self.step_size = 10                               # airhockey.py:63
        _capi.check(self._lib.ah_init_state(self._h, self._stream()), "ah_init_state", self._h)

    # ------------------------------------------------------------------ plumbing
    def _bind(self) -> None:
        _capi.check(self._lib.ah_bind_state(self._h, self.puck.data_ptr(), self.robot.data_ptr(), self.opponent.data_ptr(),
                                            self.prev_dist.data_ptr(), self.meta.data_ptr()), "ah_bind_state", self._h)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ah_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_tensor(self, t: torch.Tensor, shape: Sequence[int], dtype: torch.dtype, name: str) -> torch.Tensor:
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype} {tuple(shape)} on {self.device}, "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device}")
        return t

    def _table(self, reset_table: Optional[torch.Tensor]):
        if reset_table is None:
            return None, 0
        if (reset_table.device != self.device or reset_table.dtype != torch.float64 or reset_table.dim() != 3
                or reset_table.shape[0] != self.n or reset_table.shape[2] != 2 or not reset_table.is_contiguous()):
            raise ValueError("reset_table must be a contiguous float64 [N, R, 2] tensor on the env's device")
        return reset_table.data_ptr(), int(reset_table.shape[1])

    # ------------------------------------------------------------------ views
    @property
    def scores(self) -> torch.Tensor:
        """int16 [N,2] view: (robot_score, opponent_score)  (airhockey.py:52)."""
        return self.meta.view(torch.int16)[:, 0:2]

    @property
    def robot_score(self) -> torch.Tensor:
        return self.scores[:, 0]

    @property
    def opponent_score(self) -> torch.Tensor:
        return self.scores[:, 1]

    @property
    def reset_count(self) -> torch.Tensor:
        return self.meta[:, 1]

    # ------------------------------------------------------------------ reference verbs
    def init_state(self) -> None:
        """Back to the constructor state (airhockey.py:25-63)."""
        _capi.check(self._lib.ah_init_state(self._h, self._stream()), "ah_init_state", self._h)

    def reset(self, total: bool = False, mask: Optional[torch.Tensor] = None,
              reset_table: Optional[torch.Tensor] = None) -> None:
        """``AirHockey.reset(total)``: puck to the centre with a fresh U(-10,10) velocity, mallet velocities
        zeroed (positions kept); ``total`` also clears both scores.  ``mask`` (uint8/bool [N]) selects envs."""
        mp = None
        if mask is not None:
            mask = self._check_tensor(mask.to(torch.uint8) if mask.dtype == torch.bool else mask, (self.n,), torch.uint8, "mask")
            mp = mask.data_ptr()
        tp, R = self._table(reset_table)
        _capi.check(self._lib.ah_reset(self._h, int(total), mp, tp, R, self._stream()), "ah_reset", self._h)

    def update_state(self, action: Optional[torch.Tensor] = None, agent_name: str = "") -> None:
        """``AirHockey.update_state(action, agent_name)``: one physics sub-update for every env.

        robot:    ``action`` int8/int64 [N] (discrete) or real [N,2] (continuous, added to the velocity)
        human:    ``action`` real [N,2] = new location of the opponent's mallet
        computer: no action (the heuristic opponent moves)
        """
        if agent_name not in _AGENTS:
            raise InvalidAgentError(agent_name)          # airhockey.py:108-110
        agent = _AGENTS[agent_name]
        ap, kind = None, _capi.AH_ACT_NONE
        if agent != _capi.AH_AGENT_COMPUTER and action is not None:
            if action.dim() == 2:
                action = self._check_tensor(action, (self.n, 2), self.dtype, "action")
                kind = _capi.AH_ACT_CONTINUOUS
            else:
                if action.dtype != torch.int8:
                    action = action.to(torch.int8)
                action = self._check_tensor(action, (self.n,), torch.int8, "action")
                kind = _capi.AH_ACT_DISCRETE
            ap = action.data_ptr()
        _capi.check(self._lib.ah_update_state(self._h, ap, kind, agent, self._stream()), "ah_update_state", self._h)

    def get_state(self, agent_name: str = "robot", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``serialize_state(AirHockey.get_state(agent_name))`` for every env: real [N,8] =
        int(agent.x), int(agent.y), int(puck.x), int(puck.y), agent.dx, agent.dy, puck.dx, puck.dy."""
        if out is None:
            out = torch.empty(self.n, 8, dtype=self.dtype, device=self.device)
        self._check_tensor(out, (self.n, 8), self.dtype, "out")
        agent = _capi.AH_AGENT_ROBOT if agent_name == "robot" else _capi.AH_AGENT_COMPUTER   # airhockey.py:139
        _capi.check(self._lib.ah_get_state(self._h, agent, out.data_ptr(), self._stream()), "ah_get_state", self._h)
        return out

    def update_score(self, score: torch.Tensor, reset_table: Optional[torch.Tensor] = None) -> None:
        """``AirHockey.update_score(score)``: +1 -> robot_score++, -1 -> opponent_score++, then ``reset()``."""
        if score.dtype != torch.int8:
            score = score.to(torch.int8)
        self._check_tensor(score, (self.n,), torch.int8, "score")
        tp, R = self._table(reset_table)
        _capi.check(self._lib.ah_update_score(self._h, score.data_ptr(), tp, R, self._stream()), "ah_update_score", self._h)

    def rewards(self):
        """``Rewards.__call__(puck, robot)`` on the current state -> (reward real [N], score i8 [N], done u8 [N])."""
        reward = torch.empty(self.n, dtype=self.dtype, device=self.device)
        score = torch.empty(self.n, dtype=torch.int8, device=self.device)
        done = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        _capi.check(self._lib.ah_rewards(self._h, reward.data_ptr(), score.data_ptr(), done.data_ptr(), self._stream()),
                    "ah_rewards", self._h)
        return reward, score, done

    def first_frame(self, actions: torch.Tensor, reset_table: Optional[torch.Tensor] = None) -> None:
        """The reference's very first frame (``Game.robot_player`` with ``self.init``, game.py:118-130, followed by
        the rest of ``Game.play``, game.py:162-177): the robot only MOVES (no ``Agent.step``: no reward, no score),
        then the opponent's sub-update and the two 10-goal checks.  Every later frame is ``step``."""
        self.update_state(actions, "robot")                          # game.py:128
        self.update_state(agent_name="computer")                     # game.py:166
        for col in (1, 0):                                           # opponent first, then robot (game.py:169-177)
            over = self.scores[:, col] == 10
            self.reset(total=True, mask=over, reset_table=reset_table)

    # ------------------------------------------------------------------ the fused frame
    def make_outputs(self, K: int = 1, fields: Sequence[str] = ("reward", "done", "score", "game_over", "obs_next"),
                     obs_next_per_frame: bool = False, n_values: int = 4) -> StepOutputs:
        """Preallocate output tensors for ``step`` (reuse them across steps: nothing is allocated per step)."""
        o = StepOutputs()
        for f in fields:
            if f not in STEP_FIELDS:
                raise ValueError(f"unknown output field {f!r}")
        r = dict(dtype=self.dtype, device=self.device)
        if "obs_state" in fields:
            o.obs_state = torch.empty(K, self.n, 8, **r)
        if "obs_new_state" in fields:
            o.obs_new_state = torch.empty(K, self.n, 8, **r)
        if "reward" in fields:
            o.reward = torch.empty(K, self.n, **r)
        if "done" in fields:
            o.done = torch.empty(K, self.n, dtype=torch.uint8, device=self.device)
        if "score" in fields:
            o.score = torch.empty(K, self.n, dtype=torch.int8, device=self.device)
        if "game_over" in fields:
            o.game_over = torch.empty(K, self.n, dtype=torch.uint8, device=self.device)
        if "obs_next" in fields:
            o.obs_next = torch.empty((K, self.n, 8) if obs_next_per_frame else (self.n, 8), **r)
        if "actions" in fields:
            o.actions = torch.empty(K, self.n, dtype=torch.int8, device=self.device)
        if "probs" in fields:
            o.probs = torch.empty(K, self.n, n_values, **r)     # a device policy's head values: n_actions (or 2: continuous)
        if "actions_real" in fields:
            o.actions_real = torch.empty(K, self.n, 2, **r)
        if "events" in fields:
            o.events = torch.empty((K + 3) // 4, self.n, 4, dtype=torch.uint8, device=self.device)
        return o

    def step(self, actions: Optional[torch.Tensor] = None, K: int = 1, *, out: Optional[StepOutputs] = None,
             stream: Optional[int] = None, **kw) -> StepOutputs:
        """``plan_step`` + ``launch``: see ``plan_step`` for the arguments.  ``stream``: raw ``cudaStream_t`` to enqueue on
        (default: torch's current stream)."""
        if out is None:
            out = self.make_outputs(K)
        plan = self.plan_step(actions, K, out=out, **kw)
        self.launch(plan, stream)
        return out

    def launch(self, plan: "StepPlan", stream: Optional[int] = None) -> None:
        """Enqueue a prepared step (``plan_step``) -- one C-ABI call, no argument checking, no allocation."""
        self._io_keepalive = plan
        _capi.check(self._lib.ah_step(self._h, plan.io_ref, plan.K, self._stream() if stream is None else stream),
                    "ah_step", self._h)

    def plan_step(self, actions: Optional[torch.Tensor] = None, K: int = 1, *, repeat: bool = False,
                  out: Optional[StepOutputs] = None, reset_table: Optional[torch.Tensor] = None,
                  ring: Optional[ReplayRing] = None, collect_stats: bool = True,
                  policy_weights: Optional[torch.Tensor] = None, env_range: Optional[Sequence[int]] = None,
                  stats_bank: int = 0, policy=None, ring_counters: Optional[Sequence[int]] = None) -> "StepPlan":
        """Validate the arguments of one step and freeze them into a ``StepPlan`` that ``launch`` can enqueue any number
        of times (the tensors are referenced, not copied: replays read / write the same memory).

        K fused frames of ``Game.play`` for every env, in ONE kernel launch (state stays in registers).

        actions: None                -> uniform random {0,1,2,3} drawn in-kernel (Philox)
                 int8 [K,N]          -> one discrete action per frame   ([N] when K == 1 or ``repeat``)
                 real [K,N,2]        -> one continuous action per frame ([N,2] when K == 1 or ``repeat``)
        repeat:  apply the single given action to all K frames (frame-skip)
        out:     preallocated ``StepOutputs`` (see ``make_outputs``); default: reward/done/score/game_over/obs_next
        policy:  a ``policy.DevicePolicy``: closed loop with ANY small dense policy + exploration strategy evaluated in the
                 kernel every frame (``actions`` must be None); ``out.actions`` (discrete) / ``out.actions_real``
                 (continuous head) / ``out.probs`` (head values, [K,N,n_actions]) receive what it did
        env_range: (first, count) -- packed mode only: step just that slice of the envs (tensor shapes stay those of the
                 full N), so one step can be issued as disjoint slices on different streams (host.InterleavedStepper)
        ring_counters: (first, count) -- with ``ring`` and ``env_range``: which of the ring's counter copies this slice reads
                 and advances (``ah_step_io.ring_counter_first``; InterleavedStepper assigns them)
        policy_weights: float32 [114] device tensor (``rollout.FusedActor.weights``): closed loop -- every frame the
                 reference-shaped actor is evaluated IN the kernel on the env's current observation and the action
                 is sampled from its softmax (``actions`` must be None); ``out.actions`` / ``out.probs`` receive them
        """
        if out is None:
            out = self.make_outputs(K)
        io = _capi.AhStepIO()
        if policy is not None:
            if actions is not None or policy_weights is not None:
                raise ValueError("give either actions, policy_weights or policy")
            if policy.module is None:
                raise ValueError("a selection-only DevicePolicy has no network to run in the kernel (use policy_act)")
            io.action_kind = _capi.AH_ACT_POLICY_MLP
            io.policy = C.pointer(policy.c)
        elif policy_weights is not None:
            if actions is not None:
                raise ValueError("give either actions or policy_weights")
            self._check_tensor(policy_weights, (_capi.AH_ACTOR_NWEIGHTS,), torch.float32, "policy_weights")
            io.action_kind = _capi.AH_ACT_POLICY
            io.actions = policy_weights.data_ptr()
        elif out.events is not None:
            # PACKED rollout mode: quad-major uint8 actions in, one event byte per env-frame out (+ obs_next [N,8])
            if actions is None or actions.dtype not in (torch.uint8, torch.int8):
                raise ValueError("the packed mode takes uint8 [ceil(K/4), N, 4] actions (see pack_actions)")
            self._check_tensor(actions, ((K + 3) // 4, self.n, 4), actions.dtype, "actions")
            io.action_kind = _capi.AH_ACT_DISCRETE_PACKED4
            io.actions = actions.data_ptr()
        elif actions is None:
            io.action_kind = _capi.AH_ACT_PHILOX
        else:
            per_env = repeat or K == 1
            if actions.dtype in (torch.float32, torch.float64):
                shape = (self.n, 2) if (per_env and actions.dim() == 2) else (K, self.n, 2)
                self._check_tensor(actions, shape, self.dtype, "actions")
                io.action_kind = _capi.AH_ACT_CONTINUOUS_REPEAT if repeat else _capi.AH_ACT_CONTINUOUS
            else:
                if actions.dtype != torch.int8:
                    actions = actions.to(torch.int8)
                shape = (self.n,) if (per_env and actions.dim() == 1) else (K, self.n)
                self._check_tensor(actions, shape, torch.int8, "actions")
                io.action_kind = _capi.AH_ACT_DISCRETE_REPEAT if repeat else _capi.AH_ACT_DISCRETE
            io.actions = actions.data_ptr()
        tp, R = self._table(reset_table)
        io.reset_table, io.reset_table_len = tp, R
        r = self.dtype
        if out.obs_state is not None:
            io.obs_state = self._check_tensor(out.obs_state, (K, self.n, 8), r, "obs_state").data_ptr()
        if out.obs_new_state is not None:
            io.obs_new_state = self._check_tensor(out.obs_new_state, (K, self.n, 8), r, "obs_new_state").data_ptr()
        if out.reward is not None:
            io.reward = self._check_tensor(out.reward, (K, self.n), r, "reward").data_ptr()
        if out.done is not None:
            io.done = self._check_tensor(out.done, (K, self.n), torch.uint8, "done").data_ptr()
        if out.score is not None:
            io.score = self._check_tensor(out.score, (K, self.n), torch.int8, "score").data_ptr()
        if out.game_over is not None:
            io.game_over = self._check_tensor(out.game_over, (K, self.n), torch.uint8, "game_over").data_ptr()
        if out.obs_next is not None:
            per_frame = out.obs_next.dim() == 3
            io.obs_next = self._check_tensor(out.obs_next, (K, self.n, 8) if per_frame else (self.n, 8), r, "obs_next").data_ptr()
            io.obs_next_per_frame = int(per_frame)
        if out.actions is not None:
            io.actions_out = self._check_tensor(out.actions, (K, self.n), torch.int8, "actions").data_ptr()
        if out.probs is not None:
            if policy_weights is None and policy is None:
                raise ValueError("out.probs is only produced in policy mode")
            nv = 4 if policy is None else policy.n_actions
            io.probs_out = self._check_tensor(out.probs, (K, self.n, nv), r, "probs").data_ptr()
        if out.actions_real is not None:
            if policy is None or not policy.continuous:
                raise ValueError("out.actions_real is only produced by a continuous device policy")
            io.actions_real_out = self._check_tensor(out.actions_real, (K, self.n, 2), r, "actions_real").data_ptr()
        if out.events is not None:
            io.events = self._check_tensor(out.events, ((K + 3) // 4, self.n, 4), torch.uint8, "events").data_ptr()
        io.collect_stats = int(collect_stats)
        io.stats_bank = int(stats_bank)
        if env_range is not None:
            io.env_first, io.env_count = int(env_range[0]), int(env_range[1])
        if ring_counters is not None:
            io.ring_counter_first, io.ring_counter_count = int(ring_counters[0]), int(ring_counters[1])
        if ring is not None:
            if ring.n != self.n or ring.state.dtype != self.dtype or ring.state.device != self.device:
                raise ValueError("ring does not match this environment")
            if ring.continuous != (io.action_kind in (_capi.AH_ACT_CONTINUOUS, _capi.AH_ACT_CONTINUOUS_REPEAT)
                                   or (policy is not None and policy.continuous)):
                raise ValueError("ring action layout (discrete/continuous) does not match the actions given")
            if ring.compact and (out is None or out.events is None):
                raise ValueError("a compact ring is appended by the packed rollout mode only (packed actions + out.events)")
            io.ring = C.pointer(ring._c)
        return StepPlan(io, C.byref(io), int(K), out, (actions, reset_table, ring, policy_weights, policy))

    def policy_act(self, policy, obs: Optional[torch.Tensor] = None, values: Optional[torch.Tensor] = None,
                   actions: Optional[torch.Tensor] = None, values_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One action per env from a ``policy.DevicePolicy`` at the env's current frame (``ah_policy_act``): the network is
        evaluated on ``obs`` (real [N,8]; default: the current observation), or -- for ANY torch network -- the head values
        it produced are passed as ``values`` (real [N,n_actions]) and only the exploration strategy runs on the device
        (replaces ``torch.multinomial`` / ``argmax`` + host-side epsilon schedules).  Returns int8 [N] (discrete) or real
        [N,2] (continuous head); ``values_out`` (real [N,n_actions]) receives the head values."""
        if values is None and obs is None:
            obs = self.get_state("robot")
        if obs is not None:
            self._check_tensor(obs, (self.n, 8), self.dtype, "obs")
        if values is not None:
            self._check_tensor(values, (self.n, policy.n_actions), self.dtype, "values")
        if actions is None:
            actions = (torch.empty(self.n, 2, dtype=self.dtype, device=self.device) if policy.continuous
                       else torch.empty(self.n, dtype=torch.int8, device=self.device))
        self._check_tensor(actions, (self.n, 2) if policy.continuous else (self.n,),
                           self.dtype if policy.continuous else torch.int8, "actions")
        if values_out is not None:
            self._check_tensor(values_out, (self.n, policy.n_actions), self.dtype, "values_out")
        _capi.check(self._lib.ah_policy_act(self._h, policy.ref(), None if obs is None else obs.data_ptr(),
                                            None if values is None else values.data_ptr(),
                                            None if policy.continuous else actions.data_ptr(),
                                            actions.data_ptr() if policy.continuous else None,
                                            None if values_out is None else values_out.data_ptr(), self._stream()),
                    "ah_policy_act", self._h)
        return actions

    def c51_project(self, z_opt: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor, n_actions: int,
                    gamma: float, v_min: float, v_max: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """C51's categorical projection of a replay minibatch (lib/agents/c51/c51.py:175-190) in one kernel
        (``ah_c51_project``).  z_opt real [B, atoms], action int8 [B], reward real [B], done u8 [B] (the tensors
        ``ReplayRing.sample`` returns) -> m_prob real [n_actions, B, atoms]."""
        B, atoms = z_opt.shape
        for t, shape, dt, name in ((z_opt, (B, atoms), self.dtype, "z_opt"), (action, (B,), torch.int8, "action"),
                                   (reward, (B,), self.dtype, "reward"), (done, (B,), torch.uint8, "done")):
            if t.device != self.device or t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous():
                raise ValueError(f"{name}: expected contiguous {dt} {shape} on {self.device}")
        if out is None:
            out = torch.empty(n_actions, B, atoms, dtype=self.dtype, device=self.device)
        _capi.check(self._lib.ah_c51_project(self._h, B, int(n_actions), atoms, float(v_min), float(v_max), float(gamma),
                                             z_opt.data_ptr(), action.data_ptr(), reward.data_ptr(), done.data_ptr(),
                                             out.data_ptr(), self._stream()), "ah_c51_project", self._h)
        return out

    # ------------------------------------------------------------------ statistics, counters, checkpoints
    def stats_tensor(self, clear: bool = False, out: Optional[torch.Tensor] = None, bank: int = 0) -> torch.Tensor:
        """int64 [AH_NSTATS] device tensor of the AH_STAT_* counters accumulated by ``step`` (no host sync).  ``bank``:
        which of the two statistics banks (``step(stats_bank=...)``; everything defaults to bank 0)."""
        out = self._stats_buf if out is None else out
        _capi.check(self._lib.ah_stats_bank(self._h, int(bank), out.data_ptr(), int(clear), self._stream()), "ah_stats_bank", self._h)
        return out

    def stats(self, clear: bool = False) -> Dict[str, int]:
        v = self.stats_tensor(clear).tolist()
        return dict(zip(_capi.STAT_NAMES, v))

    def set_counters(self, frame: int = 0) -> None:
        _capi.check(self._lib.ah_set_counters(self._h, frame, self._stream()), "ah_set_counters", self._h)

    def counters(self) -> Dict[str, int]:
        buf = torch.zeros(1, dtype=torch.int64, device=self.device)
        _capi.check(self._lib.ah_get_counters(self._h, buf.data_ptr(), self._stream()), "ah_get_counters", self._h)
        return {"frame": int(buf.item())}

    def launch_info(self) -> Dict[str, int]:
        b, t, s = C.c_int(), C.c_int(), C.c_int()
        _capi.check(self._lib.ah_launch_info(self._h, C.byref(b), C.byref(t), C.byref(s)), "ah_launch_info", self._h)
        return {"blocks": b.value, "threads": t.value, "sms": s.value}

    def compiled_without_fma_contraction(self) -> bool:
        buf = torch.ones(1, dtype=torch.float64, device=self.device)
        _capi.check(self._lib.ah_selftest_nofma(self._h, buf.data_ptr(), self._stream()), "ah_selftest_nofma", self._h)
        return float(buf.item()) == 0.0

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Exact checkpoint of the environment (the reference never checkpoints env state)."""
        c = self.counters()
        return {"puck": self.puck.clone(), "robot": self.robot.clone(), "opponent": self.opponent.clone(),
                "prev_dist": self.prev_dist.clone(), "meta": self.meta.clone(),
                "counters": torch.tensor([c["frame"]], dtype=torch.int64),
                "seed": torch.tensor([self.seed, self.env_id0], dtype=torch.int64)}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k in ("puck", "robot", "opponent", "prev_dist", "meta"):
            getattr(self, k).copy_(sd[k])
        self.set_counters(int(sd["counters"][0]))


class InterleavedStepper:
    """Issues every packed step as ``parts`` disjoint env slices on ``parts`` private streams (``env_range`` launches).

    One launch over all N envs pays a ramp (every resident block loads its state at once) and a tail (the last blocks
    finish on a mostly idle GPU) per step.  Slices on different streams are independent (an env only ever depends on
    its own previous step, which is earlier on the SAME stream), so the GPU never drains between steps: the tail of
    one slice's launch overlaps the head of the next slice's.  Results are identical to ``env.step`` (same kernel, same
    per-env arithmetic; tests/test_gpu_packed.py).

    ``step`` returns immediately; call ``join()`` before reading outputs / state / statistics on the caller's stream.
    """

    def __init__(self, env: BatchedAirHockey, parts: int = 2, align: int = 1024):
        self.env, self.parts = env, int(parts)
        per = -(-env.n // self.parts)
        per = -(-per // align) * align                       # whole blocks per slice
        self.ranges = [(lo, min(per, env.n - lo)) for lo in range(0, env.n, per)]
        self.streams = [torch.cuda.Stream(env.device) for _ in self.ranges]
        self._raw_streams = [s.cuda_stream for s in self.streams]
        self.stats_stream = torch.cuda.Stream(env.device)
        self.bank = 0
        self._forked = False
        self._stats_pending = False                          # close_rollout queued work on stats_stream since the last join
        self._plans: Dict[tuple, list] = {}

    def step(self, actions: torch.Tensor, K: int, out: StepOutputs, **kw) -> StepOutputs:
        if out.events is None:
            raise ValueError("InterleavedStepper drives the packed mode (out.events)")
        cur = torch.cuda.current_stream(self.env.device)
        if not self._forked:                                 # order the slices after whatever the caller queued
            for s in self.streams:
                s.wait_stream(cur)
            self._forked = True
        key = (actions.data_ptr(), id(out), K, self.bank, tuple(sorted((k, id(v)) for k, v in kw.items())))
        plans = self._plans.get(key)
        if plans is None:                                    # validated once per (actions, outputs, bank) combination
            if len(self._plans) > 64:
                self._plans.clear()
            P = len(self.ranges)
            if kw.get("ring") is not None and P > _capi.AH_RING_COUNTERS:
                raise ValueError(f"a ring append can be issued as at most {_capi.AH_RING_COUNTERS} slices")
            # with a ring: slice j reads and advances the ring's counter copy j (the last slice: all the remaining copies)
            rc = [(j, 1 if j < P - 1 else _capi.AH_RING_COUNTERS - j) if kw.get("ring") is not None else None for j in range(P)]
            plans = [self.env.plan_step(actions, K, out=out, env_range=rng, stats_bank=self.bank, ring_counters=rc[j], **kw)
                     for j, rng in enumerate(self.ranges)]
            self._plans[key] = plans
        for s, plan in zip(self._raw_streams, plans):
            self.env.launch(plan, s)
        return out

    def close_rollout(self, stats_out: torch.Tensor, peer=None) -> torch.cuda.Stream:
        """End of a rollout, WITHOUT draining the GPU: the statistics bank the steps so far accumulated into is copied
        to ``stats_out`` and cleared on a private stream that waits for exactly those steps, and later steps switch to
        the other bank -- they neither wait for the copy nor race with it.  Returns the stream the copy was queued on
        (queue the all-reduce there, or make a consumer stream wait for it).  ``peer`` (a ``dist.PeerStats``): the copy
        IS the all-reduce (fused over NVLink peer memory)."""
        done = []
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            done.append(ev)
        if not self._forked:                                 # nothing issued since the last join: order after the caller
            self.stats_stream.wait_stream(torch.cuda.current_stream(self.env.device))
        for ev in done:
            self.stats_stream.wait_event(ev)
        if peer is not None:                                 # dist.PeerStats: copy + clear + sum over the ranks, one kernel
            peer.allreduce(stats_out, bank=self.bank, clear=True, stream=self.stats_stream.cuda_stream)
        else:
            with torch.cuda.stream(self.stats_stream):
                self.env.stats_tensor(clear=True, out=stats_out, bank=self.bank)
        self.bank ^= 1
        self._stats_pending = True
        return self.stats_stream

    def join(self) -> None:
        """Make the caller's current stream wait for every slice issued so far (no host sync)."""
        cur = torch.cuda.current_stream(self.env.device)
        if self._forked:
            for s in self.streams:
                cur.wait_stream(s)
        if self._stats_pending:                              # (never wait on an idle stream: inside a CUDA-graph capture that
            cur.wait_stream(self.stats_stream)               #  would be a dependency on uncaptured work)
            self._stats_pending = False
        self._forked = False
