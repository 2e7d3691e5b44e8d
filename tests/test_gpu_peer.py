"""2-GPU test of the fused statistics all-reduce over NVLink peer memory (ah_peer_export / ah_peer_attach /
ah_stats_allreduce, dist.PeerStats): one process per GPU, NCCL only for the handle exchange and as the reference sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from air_hockey_training_environment_b200 import BatchedAirHockey, STAT_NAMES
    from air_hockey_training_environment_b200 import dist as ahdist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    n = 4096 + 512 * rank                                      # different shard sizes: the sums are not symmetric
    env = BatchedAirHockey(n, device=dev, seed=3, env_id0=10_000 * rank)
    peer = ahdist.PeerStats(env)
    assert peer.available, peer.error
    ok = True
    for epoch in range(7):                                     # odd count: both parities of the exchange buffer, twice+
        bank = epoch & 1
        env.step(None, K=40 + epoch, out=env.make_outputs(40 + epoch, ()), stats_bank=bank)
        want = env.stats_tensor(clear=False, out=torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=dev), bank=bank).clone()
        dist.all_reduce(want)                                  # reference: NCCL sum of the same vectors
        got = peer.allreduce(torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=dev), bank=bank, clear=True)
        torch.cuda.synchronize(dev)
        ok &= bool(torch.equal(got, want)) and int(got[0]) == (40 + epoch) * (4096 * world + 512 * sum(range(world)))
        ok &= int(env.stats_tensor(bank=bank)[0]) == 0          # cleared
    # back-to-back exchanges without host syncs in between (a fast rank runs ahead by at most one exchange)
    outs = [torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=dev) for _ in range(6)]
    for k, o in enumerate(outs):
        env.step(None, K=8, out=env.make_outputs(8, ()))
        peer.allreduce(o)
    torch.cuda.synchronize(dev)
    ok &= all(int(o[0]) == 8 * (4096 * world + 512 * sum(range(world))) for o in outs)
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")
def test_fused_p2p_stats_allreduce_equals_nccl():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
