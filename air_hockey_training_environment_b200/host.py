"""Host-buffer front end: step the games from HOST action buffers and get the results back on the HOST.

This is the call a CPU-side consumer of the environment makes (the reference's own loop lives on the host:
game.py:133-177).  On this path PCIe, not the kernel, is the bound, so the results cross the bus in a compact,
lossless encoding (``ah_pack_host`` in include/ah_b200.h): per frame one byte per env (reward + done), per step the
observation as 4 x int16 locations (they are int()-truncated in the reference, airhockey.py:136-147) + 4 x float32
velocities -- 33 B/env at K = 8 instead of 72.  Copies are staged through pinned buffers on dedicated streams and
double-buffered, so step i's device->host copy overlaps step i+1's kernel and step i+2's host->device copy.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _capi
from .env import BatchedAirHockey


@dataclass
class HostResult:
    """Results of one step in pinned host memory (valid after ``wait()``)."""
    events: torch.Tensor        # uint8 [ceil(K/4), n, 4]: one AH_EVENT byte per env-frame (include/ah_b200.h), quad-major
    obs_pos: torch.Tensor       # int16 [n, 4]: int(agent.x), int(agent.y), int(puck.x), int(puck.y)
    obs_vel: torch.Tensor       # float32 [n, 4]: agent.dx, agent.dy, puck.dx, puck.dy
    K: int = 8
    _event: Optional[torch.cuda.Event] = None

    def wait(self) -> "HostResult":
        if self._event is not None:
            self._event.synchronize()
        return self

    def _frames(self) -> torch.Tensor:          # uint8 [K, n], time-major
        q, n, _ = self.events.shape
        return self.events.permute(0, 2, 1).reshape(q * 4, n)[:self.K]

    @property
    def reward(self) -> torch.Tensor:           # int8 [K, n]  (lib/rewards.py:118-142 returns integers)
        return ((self._frames() & 7).to(torch.int8) << 1) - 4

    @property
    def done(self) -> torch.Tensor:             # bool [K, n]
        return ((self._frames() >> 3) & 1).to(torch.bool)

    @property
    def score(self) -> torch.Tensor:            # int8 [K, n]: +1 robot scored, -1 opponent scored
        return ((self._frames() >> 4) & 3).to(torch.int8) - 1

    @property
    def game_over(self) -> torch.Tensor:        # uint8 [K, n]: 0 / 1 robot won / 2 opponent won
        return (self._frames() >> 6) & 3

    @property
    def obs(self) -> torch.Tensor:              # float32 [n, 8] in the reference's layout
        return torch.cat([self.obs_pos.to(torch.float32), self.obs_vel], dim=1)


class HostStepper:
    """``submit(actions_host)`` enqueues H2D -> fused step -> D2H for one step and returns immediately; the returned
    ``HostResult`` is complete after ``wait()``.  ``slots`` steps can be in flight.

    The step runs in the packed rollout mode: the actions cross the bus as the quad-major uint8 words the kernel reads
    (``pack_actions_host``), the per-frame results as the event bytes it writes -- nothing is re-encoded on the device
    except the observation (4 x int16 + 4 x float32 per env).

    Streams: the copies and the kernels run on three private streams.  ``submit`` first makes them wait for the
    caller's current stream (so work the caller queued on the env -- ``init_state``, an earlier ``env.step`` -- is
    ordered before this step), and ``join()`` makes the caller's current stream wait for everything submitted (call it
    before reading ``env`` state or statistics on the caller's stream)."""

    def __init__(self, env: BatchedAirHockey, K: int = 8, slots: int = 3):
        self.env, self.K, self.slots = env, int(K), int(slots)
        n, dev, q = env.n, env.device, (int(K) + 3) // 4
        self._lib = _capi.lib()
        self.d_act = [torch.empty(q, n, 4, dtype=torch.uint8, device=dev) for _ in range(slots)]
        # a step's results live in ONE buffer per slot -- [velocities f32 [n,4] | locations i16 [n,4] | event bytes u8 [q,n,4]],
        # every part 8-byte aligned -- on the device and in pinned host memory: one device->host copy per step
        nbytes = 16 * n + 8 * n + 4 * q * n

        def parts(buf):
            return (buf[24 * n:].view(q, n, 4), buf[16 * n:24 * n].view(torch.int16).view(n, 4), buf[:16 * n].view(torch.float32).view(n, 4))

        self.d_buf = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(slots)]
        self.h_buf = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.outs, self.d_pos, self.d_vel = [], [], []
        for b in self.d_buf:
            ev, pos, vel = parts(b)
            o = env.make_outputs(K, ("obs_next",))
            o.events = ev
            self.outs.append(o); self.d_pos.append(pos); self.d_vel.append(vel)
        self.h = [HostResult(*parts(b), K=int(K)) for b in self.h_buf]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.e_in = [torch.cuda.Event() for _ in range(slots)]
        self.e_run = [torch.cuda.Event() for _ in range(slots)]
        self.e_out: List[Optional[torch.cuda.Event]] = [None] * slots
        self._i = 0
        self.h2d_bytes_per_step = q * 4 * n
        self.d2h_bytes_per_step = q * 4 * n + n * 4 * 2 + n * 4 * 4
        self.s_run.wait_stream(torch.cuda.current_stream(dev))

    @staticmethod
    def pack_actions_host(actions: torch.Tensor) -> torch.Tensor:
        """int8 [K, n] CPU actions -> pinned uint8 [ceil(K/4), n, 4] in the kernel's quad-major layout."""
        from .env import pack_actions
        return pack_actions(actions).pin_memory()

    def submit(self, actions_host: torch.Tensor) -> HostResult:
        """actions_host: uint8 [ceil(K/4), n, 4] CPU tensor (``pack_actions_host``; pinned for an asynchronous copy)."""
        K, n, s = self.K, self.env.n, self._i % self.slots
        self._i += 1
        if (actions_host.dtype != torch.uint8 or tuple(actions_host.shape) != ((K + 3) // 4, n, 4)
                or actions_host.device.type != "cpu"):
            raise ValueError(f"actions_host must be a uint8 [{(K + 3) // 4}, {n}, 4] CPU tensor (see pack_actions_host)")
        env = self.env
        cur = torch.cuda.current_stream(env.device)
        self.s_run.wait_stream(cur)                           # whatever the caller queued on the env comes first
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.e_run[s])               # the slot's previous kernel has consumed its actions
            self.d_act[s].copy_(actions_host, non_blocking=True)
            self.e_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.e_in[s])
            if self.e_out[s] is not None:
                self.s_run.wait_event(self.e_out[s])          # the slot's previous results have left the device
            o = self.outs[s]
            env.step(self.d_act[s], K=K, out=o)
            _capi.check(self._lib.ah_pack_host(env._h, K, None, None, o.obs_next.data_ptr(), None, self.d_pos[s].data_ptr(),
                                               self.d_vel[s].data_ptr(), self.s_run.cuda_stream), "ah_pack_host", env._h)
            self.e_run[s].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.e_run[s])
            h = self.h[s]
            self.h_buf[s].copy_(self.d_buf[s], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
            self.e_out[s] = ev
            h._event = ev
        return h

    def join(self) -> None:
        """Order the caller's current stream after everything submitted so far (no host sync)."""
        cur = torch.cuda.current_stream(self.env.device)
        cur.wait_stream(self.s_run)
        cur.wait_stream(self.s_out)

    def drain(self) -> None:
        for ev in self.e_out:
            if ev is not None:
                ev.synchronize()
