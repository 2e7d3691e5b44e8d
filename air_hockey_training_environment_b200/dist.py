"""Multi-GPU plumbing: environments shard trivially, one process per GPU.

The reference is a single-process, single-game loop (game.py:179-196); nothing in the environment step reads or
writes another game's state, so N games are split into contiguous slices, one per rank, with NO collective in the
step itself.  The only exchange is the small per-rollout all-reduce of the episode / score statistics vector
(the device-side equivalent of scores.csv, game.py:82-112, and of the rewards CSV, lib/rewards.py:129-140).

RNG keys use the GLOBAL env id (``env_id0 = first env of this rank``), so the union of the shards is bit-identical
to one big environment (tests/test_gpu_api.py::test_sharding_invariance).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _capi


def shard_bounds(n_global: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global env index space owned by ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_global, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_numa_node(index: int) -> Tuple[str, int]:
    p = torch.cuda.get_device_properties(index)
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    path = f"/sys/bus/pci/devices/{bdf}/numa_node"
    return bdf, (int(open(path).read().strip()) if os.path.exists(path) else -1)


def bind_host_to_gpu(local_rank: int, local_world: int = 1) -> Dict[str, object]:
    """Pin this process to host cores next to its GPU, BEFORE it allocates pinned staging buffers (first touch puts
    them on that NUMA node): the host-buffer path (``host.HostStepper``) is bound by PCIe and host memory, and at 8
    ranks per box unpinned processes all land on one node.  The cores of the GPU's NUMA node (sysfs) that this
    process may use are split evenly among the local ranks whose GPUs share that node.  Best effort: returns what it
    did (``bound: False`` with the reason when sysfs / affinity calls are not available)."""
    info: Dict[str, object] = {"bound": False}
    try:
        bdf, node = _gpu_numa_node(local_rank)
        peers = [r for r in range(local_world) if _gpu_numa_node(r)[1] == node] or [local_rank]
        allowed = sorted(os.sched_getaffinity(0))
        info.update(pci=bdf, numa_node=node, allowed_cpus=len(allowed), ranks_on_node=len(peers))
        cpus = allowed
        cpulist = f"/sys/devices/system/node/node{node}/cpulist"
        if node >= 0 and os.path.exists(cpulist):
            local = [c for c in _parse_cpulist(open(cpulist).read()) if c in allowed]
            info["node_local_cpus"] = len(local)
            if local:
                cpus = local
        share = max(1, len(cpus) // len(peers))
        k = peers.index(local_rank) if local_rank in peers else 0
        mine = cpus[k * share:(k + 1) * share] or cpus
        os.sched_setaffinity(0, mine)
        info.update(bound=True, cpus=len(mine), first_cpu=mine[0])
    except Exception as e:                                   # no sysfs / no permission: run unbound
        info["error"] = repr(e)
    return info


def init_process_group(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (nccl on GPUs, gloo on CPU)."""
    rank, world, local = rank_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the int64[AH_NSTATS] statistics vector over all ranks, in place (one 160-byte all-reduce per rollout; NCCL
    over NVLink on GPUs).  A no-op for a single process."""
    if stats.dtype != torch.int64 or stats.numel() != _capi.AH_NSTATS:
        raise ValueError("stats must be the int64[AH_NSTATS] vector returned by BatchedAirHockey.stats_tensor()")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class PeerStats:
    """The per-rollout statistics exchange fused over NVLink peer memory (``ah_stats_allreduce``): copy + clear + sum
    over all ranks of ONE box in a single one-warp kernel -- each rank stores its vector straight into every rank's
    receive buffer (P2P), flags it and waits for the others' -- instead of a copy kernel followed by an NCCL all-reduce.
    Set-up (once, collective): the ranks exchange CUDA IPC handles of their receive buffers through the process group.
    ``available`` is False (and callers fall back to ``all_reduce_stats``) when IPC / peer access cannot be set up."""

    def __init__(self, env, group=None):
        import ctypes as C
        self.env, self.available, self.error = env, False, None
        if not (dist.is_available() and dist.is_initialized()):
            self.error = "no process group"
            return
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        handle = torch.zeros(_capi.AH_IPC_HANDLE_BYTES, dtype=torch.uint8)
        rc = env._lib.ah_peer_export(env._h, handle.data_ptr()) if world <= _capi.AH_MAX_PEERS else _capi.AH_EINVAL
        ok = torch.tensor([1 if rc == _capi.AH_OK else 0], device=env.device)
        gathered = [torch.zeros(_capi.AH_IPC_HANDLE_BYTES, dtype=torch.uint8, device=env.device) for _ in range(world)]
        dist.all_gather(gathered, handle.to(env.device), group=group)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            blob = torch.stack(gathered).cpu().contiguous()
            rc = env._lib.ah_peer_attach(env._h, rank, world, blob.data_ptr())
            ok = torch.tensor([1 if rc == _capi.AH_OK else 0], device=env.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks or none
        self.available = int(ok.item()) == 1
        if not self.available:
            self.error = f"IPC set-up failed (rc {rc})"

    def allreduce(self, out: torch.Tensor, bank: int = 0, clear: bool = True, stream: Optional[int] = None) -> torch.Tensor:
        """``out`` (int64 [AH_NSTATS], device) <- the sum over all ranks of the env's statistics bank; no host sync."""
        env = self.env
        _capi.check(env._lib.ah_stats_allreduce(env._h, int(bank), out.data_ptr(), int(clear),
                                                env._stream() if stream is None else stream), "ah_stats_allreduce", env._h)
        return out


def summarize(stats: torch.Tensor) -> Dict[str, float]:
    """Host-side summary of a (reduced) statistics vector: the aggregates the reference logs per frame / per
    episode (robot_goal, opponent_goal, robot_win, opponent_win, mean reward), as rates."""
    v = dict(zip(_capi.STAT_NAMES, stats.tolist()))
    f = max(v["frames"], 1)
    g = max(v["games"], 1)
    out = dict(v)
    out.update({
        "robot_goals_per_1k_frames": 1e3 * v["robot_goals"] / f,
        "opponent_goals_per_1k_frames": 1e3 * v["opponent_goals"] / f,
        "dones_per_1k_frames": 1e3 * v["dones"] / f,
        "mean_reward": v["reward_sum"] / f,
        "reward_var": v["reward_sq_sum"] / f - (v["reward_sum"] / f) ** 2,
        "robot_win_rate": v["robot_wins"] / g,
        "mean_game_length": v["game_len_sum"] / g,
        "game_length_var": v["game_len_sq_sum"] / g - (v["game_len_sum"] / g) ** 2,
        "mean_game_return": v["game_return_sum"] / g,
        "contacts_robot_per_1k_frames": 1e3 * v["contacts_robot"] / f,
        "contacts_opponent_per_1k_frames": 1e3 * v["contacts_opponent"] / f,
    })
    return out


class ShardedAirHockey:
    """This rank's slice of ``n_global`` games.  Thin convenience over ``BatchedAirHockey``: same verbs, plus
    ``reduced_stats`` which performs the per-rollout all-reduce."""

    def __init__(self, n_global: int, dtype: torch.dtype = torch.float32, seed: int = 0, **kw):
        from .env import BatchedAirHockey
        self.rank, self.world, self.local = rank_info()
        self.n_global = int(n_global)
        self.lo, self.hi = shard_bounds(self.n_global, self.world, self.rank)
        self.env = BatchedAirHockey(self.hi - self.lo, dtype=dtype, device=torch.device("cuda", self.local),
                                    seed=seed, env_id0=self.lo, **kw)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reduced_stats(self, clear: bool = True) -> torch.Tensor:
        return all_reduce_stats(self.env.stats_tensor(clear=clear).clone())
