"""CPU tests of the multi-GPU host logic (world / pair sharding and the two collectives of the
path) with the gloo backend, world_size 2 and 3 — the same code runs over NCCL on the GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diverse_conventions_b200 import sharding


def test_world_and_pair_shards_partition_exactly():
    for total, ws in [(16384 * 8, 8), (10, 3), (7, 8), (256, 8)]:
        got = [sharding.world_shard(total, r, ws) for r in range(ws)]
        assert got[0][0] == 0 and sum(c for _, c in got) == total
        for (f0, c0), (f1, _) in zip(got, got[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in got) - min(c for _, c in got) <= 1
    pairs = sharding.all_pairs(16)
    assert len(pairs) == 256 and pairs[17] == (1, 1) and pairs[255] == (15, 15)
    shards = [sharding.pair_shard(pairs, r, 8) for r in range(8)]
    assert all(len(s) == 32 for s in shards) and sum(shards, []) == pairs


def test_single_process_gather_is_a_plain_scatter():
    pairs = [(0, 1), (2, 2), (0, 1)]  # a pair evaluated twice is pooled
    mean, eps = sharding.gather_pair_matrix(pairs, torch.tensor([10, 40, 30]), torch.tensor([2, 4, 2]), 3)
    assert eps.tolist() == [[0, 4, 0], [0, 0, 0], [0, 0, 4]]
    assert float(mean[0, 1]) == 10.0 and float(mean[2, 2]) == 10.0 and torch.isnan(mean[1, 1])
    s, n = sharding.reduce_episode_stats(torch.tensor([5, 7]), torch.tensor([1, 1]))
    assert (int(s), int(n)) == (12, 2)
    with pytest.raises(ValueError):
        sharding.gather_pair_matrix([(0, 3)], torch.tensor([1]), torch.tensor([1]), 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stats_for(pair):  # deterministic fake statistics of a pair
    i, j = pair
    return 100 * i + 7 * j + 3, 2 + (i + j) % 3


def _worker(rank, world, port, n_pol, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs = sharding.all_pairs(n_pol)
        mine = sharding.pair_shard(pairs, rank, world)  # ragged: 9 pairs over 2 ranks -> 5 + 4
        rs = torch.tensor([_stats_for(p)[0] for p in mine], dtype=torch.int64)
        ep = torch.tensor([_stats_for(p)[1] for p in mine], dtype=torch.int32)
        mean, eps = sharding.gather_pair_matrix(mine, rs, ep, n_pol)
        first, count = sharding.world_shard(1000, rank, world)
        s, n = sharding.reduce_episode_stats(torch.arange(first, first + count), torch.ones(count, dtype=torch.int32))
        # plain lists, not tensors: a tensor in an mp.Queue is rebuilt from a file descriptor the SENDER must still hold when the
        # parent fetches it — a worker that had already exited made this test fail with EOFError on a loaded machine
        q.put((rank, mean.tolist(), eps.tolist(), int(s), int(n)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_across_ranks_gloo(world):
    n_pol = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pol, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_mean = torch.tensor([[_stats_for((i, j))[0] / _stats_for((i, j))[1] for j in range(n_pol)] for i in range(n_pol)],
                             dtype=torch.float64)
    want_eps = torch.tensor([[_stats_for((i, j))[1] for j in range(n_pol)] for i in range(n_pol)])
    for rank, mean, eps, s, n in results:  # every rank ends with the full matrix
        assert torch.equal(torch.tensor(mean, dtype=torch.float64), want_mean) and torch.equal(torch.tensor(eps), want_eps)
        assert (s, n) == (sum(range(1000)), 1000)
