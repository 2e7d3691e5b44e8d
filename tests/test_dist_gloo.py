"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: contiguous sharding, global-env-id RNG keying and
the per-rollout all-reduce of the statistics vector.  The per-rank "device" here is the C oracle (tests may use
it); on the GPU box the same logic is exercised with NCCL by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from air_hockey_training_environment_b200 import dist as ahdist


def test_shard_bounds_partition_the_index_space():
    for n in (1, 7, 8, 1000, 1 << 20, (1 << 23) + 5):
        for w in (1, 2, 3, 4, 8):
            b = [ahdist.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ahdist.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, T, seed, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from oracle import oracle
    r, w, _ = ahdist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    lo, hi = ahdist.shard_bounds(n_global, world, rank)
    st, sc = oracle.initial_state(hi - lo)
    out = oracle.run_frames(st, sc, T, seed=seed, env_id0=lo, want=())       # Philox actions + resets, keyed globally
    stats = torch.from_numpy(out["stats"].copy())
    ahdist.all_reduce_stats(stats)
    gathered = [torch.zeros(hi - lo if rr == rank else ahdist.shard_bounds(n_global, world, rr)[1]
                            - ahdist.shard_bounds(n_global, world, rr)[0], 13, dtype=torch.float64) for rr in range(world)]
    dist.all_gather(gathered, torch.from_numpy(st))
    if rank == 0:
        ret["stats"] = stats.numpy().copy()
        ret["state"] = torch.cat(gathered).numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_one_big_env():
    from oracle import oracle
    n_global, T, seed, world = 258, 600, 42, 2          # uneven-free split 129 + 129; long enough for goals
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_global, T, seed, ret), nprocs=world, join=True)
    st, sc = oracle.initial_state(n_global)
    full = oracle.run_frames(st, sc, T, seed=seed, env_id0=0, want=())
    assert np.array_equal(ret["stats"], full["stats"])                        # all-reduced stats == single-process stats
    assert np.array_equal(ret["state"].view(np.uint64), st.view(np.uint64))   # shards are bit-identical to the big env
    assert full["stats"][1] + full["stats"][2] > 0                            # goals happened
    s = ahdist.summarize(torch.from_numpy(ret["stats"]))
    assert s["frames"] == n_global * T and 0 < s["robot_goals_per_1k_frames"] + s["opponent_goals_per_1k_frames"] < 30


def test_all_reduce_is_a_noop_without_a_process_group():
    v = torch.arange(20, dtype=torch.int64)
    assert torch.equal(ahdist.all_reduce_stats(v.clone()), v)
    with pytest.raises(ValueError):
        ahdist.all_reduce_stats(torch.zeros(4))
