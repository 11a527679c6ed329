"""Walker / replica sharding over the ranks of one box: independent units, no data-path collective.

Rank r of `world` owns a contiguous block of walkers; seeds, temperatures and occupancies are functions of the GLOBAL
walker index so that a run is independent of the number of ranks.  The only communication is control-plane: a barrier
around the timed region, a MAX-reduction of per-rank times and a final gather of per-walker statistics.  Works with any
torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """(first, count) of the contiguous block of `total` units owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def interleaved_walkers(total: int, rank: int, world: int):
    """Global indices of the walkers owned by `rank` when walker w lives on GPU w mod world: neighbouring
    walkers -- neighbouring temperatures, hence similar cost -- are dealt to different ranks, so every rank gets the same
    mix and no rank is the slow one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(int(rank), int(total), int(world), dtype=np.int64)


def temperatures_of(walkers, total: int, t_lo: float = 400.0, t_hi: float = 600.0):
    """T_w = t_lo + (t_hi - t_lo) * w / (total - 1) for global walker indices."""
    return t_lo + (t_hi - t_lo) * np.asarray(walkers, dtype=np.float64) / max(1, int(total) - 1)


def walker_temperatures(first: int, count: int, total: int, t_lo: float = 400.0, t_hi: float = 600.0):
    """T_w = t_lo + (t_hi - t_lo) * w / (total - 1) for the global walker indices of this shard."""
    w = first + np.arange(count)
    return t_lo + (t_hi - t_lo) * w / max(1, total - 1)


def max_over_ranks(values, device=None):
    """Element-wise MAX of a small list of floats over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]


def gather_walker_stats(local: np.ndarray, total: int, device=None):
    """Gather a per-walker float64 array from every rank into global walker order on rank 0 (None elsewhere).
    Shards may differ in length by one; they are padded to the longest for the collective."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        return local.copy()
    world, rank = dist.get_world_size(), dist.get_rank()
    longest = max(shard_range(total, r, world)[1] for r in range(world))
    pad = torch.zeros(longest, dtype=torch.float64, device=device or "cpu")
    pad[:len(local)] = torch.from_numpy(local).to(pad.device)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank != 0:
        return None
    return np.concatenate([out[r][:shard_range(total, r, world)[1]].cpu().numpy() for r in range(world)])


def evaluation_owner(warp: int, world: int) -> int:
    """Rank that evaluates warp-group `warp` of every thread block in the multi-GPU single-lattice CMC driver
    (cmc_grid_kernels.cuh: `owner = warp % world == rank`)."""
    return int(warp) % int(world)


def attach_cmc_peers(engine, rank: int, world: int, sm_count: int | None = None):
    """Collective set-up of the multi-GPU single-lattice CMC driver (include/lmc_b200.h, lmc_cmc_attach_peers): all-gather
    the 64-byte CUDA IPC handles of the ranks' exchange buffers and the SM counts (the grid must be identical on every
    rank), attach, barrier.  With world == 1 (or no process group) the engine is simply reset to a world of one."""
    import torch.distributed as dist
    handle = engine.cmc_exchange_handle()
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        engine.cmc_attach_peers(0, 1, [handle], 0)
        return 0
    if sm_count is None:
        import torch
        sm_count = torch.cuda.get_device_properties(engine.device).multi_processor_count
    gathered = [None] * world
    dist.all_gather_object(gathered, (bytes(handle), int(sm_count)))
    grid_ctas = min(g[1] for g in gathered)
    engine.cmc_attach_peers(rank, world, [g[0] for g in gathered], grid_ctas)
    dist.barrier()
    return grid_ctas


def domain_slab(n_domains_x: int, rank: int, world: int):
    """(first, count) of the x slab of the domain grid owned by `rank` in the domain-decomposed CMC / SA driver
    (cmc_domain_kernels.cuh: `domain_slab_begin(ndx, world, r) = ndx * r / world`)."""
    first = (int(n_domains_x) * int(rank)) // int(world)
    return first, (int(n_domains_x) * (int(rank) + 1)) // int(world) - first


def attach_cmc_domain_peers(engine, rank: int, world: int):
    """Collective set-up of the multi-GPU domain-decomposed CMC / SA driver (include/lmc_b200.h,
    lmc_cmc_domain_attach_peers): all-gather the 192-byte CUDA IPC handle blobs (two occupancy buffers + line buffer),
    attach, barrier."""
    import torch.distributed as dist
    blob = engine.cmc_domain_handles()
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        engine.cmc_domain_attach_peers(0, 1, [blob])
        return
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(blob))
    engine.cmc_domain_attach_peers(rank, world, gathered)
    dist.barrier()
