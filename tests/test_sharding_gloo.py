"""CPU suite: the N>1 host logic (walker sharding, max-over-ranks timing, statistics gather) under torch.distributed
with the gloo backend, world_size 2 -- the same code path bench.py runs over NCCL on the GPU box."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from latticemontecarlo_b200 import sharding


def test_shard_ranges_partition_the_walkers():
    for total in (1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    t = np.concatenate([sharding.walker_temperatures(*sharding.shard_range(8192, r, 8), 8192) for r in range(8)])
    assert np.allclose(t, 400.0 + 200.0 * np.arange(8192) / 8191.0)          # independent of the rank count


def _worker(rank, world, port, total):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = sharding.shard_range(total, rank, world)
        # every rank "measures" a different time; the reported one is the slowest rank's
        assert sharding.max_over_ranks([10.0 + rank, 1.0 - rank]) == [10.0 + world - 1, 1.0]
        assert sharding.sum_over_ranks([count]) == [float(total)]
        stats = sharding.gather_walker_stats(np.arange(first, first + count, dtype=np.float64) * 2.0, total)
        if rank == 0:
            assert np.array_equal(stats, 2.0 * np.arange(total))
        else:
            assert stats is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)


def test_evaluation_owner_partitions_the_warp_groups():
    """Multi-GPU single-lattice CMC: every warp group of a thread block has exactly one evaluating rank, and the groups
    are spread evenly (cmc_grid_kernels.cuh: owner = warp % world)."""
    from latticemontecarlo_b200 import sharding
    for world in (1, 2, 3, 4, 8):
        owners = [sharding.evaluation_owner(w, world) for w in range(16)]
        assert all(0 <= o < world for o in owners)
        counts = [owners.count(r) for r in range(world)]
        assert max(counts) - min(counts) <= 1


def test_interleaved_walkers_and_domain_slabs_partition_their_units():
    """bench.py's strong-scaling KMC deals walker w to rank w mod N; the multi-GPU domain CMC driver gives rank r the x slab
    [ndx r / N, ndx (r + 1) / N) of the domain grid (cmc_domain.h: domain_slab_begin)."""
    for total in (1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            parts = [sharding.interleaved_walkers(total, r, world) for r in range(world)]
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            t = np.concatenate([sharding.temperatures_of(p, total) for p in parts])
            assert np.allclose(np.sort(t), 400.0 + 200.0 * np.arange(total) / max(1, total - 1))
            # every rank sees the whole temperature range (that is the point of dealing w mod N)
            if total >= 64:
                assert all(sharding.temperatures_of(p, total).max() - sharding.temperatures_of(p, total).min() > 190.0 for p in parts)
    for ndx in (1, 5, 10, 25):
        for world in (1, 2, 4, 8):
            slabs = [sharding.domain_slab(ndx, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and sum(c for _, c in slabs) == ndx
            for (f0, c0), (f1, _) in zip(slabs, slabs[1:]):
                assert f0 + c0 == f1
