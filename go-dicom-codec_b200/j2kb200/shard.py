"""Frame / tile sharding across ranks (one process per GPU) — no data-path collective.

The reference's frame loops (jpeg2000/lossless/codec.go:246-261, jpeg2000/lossy/codec.go:149-176)
and tile loops (jpeg2000/encoder.go:2010-2014) iterate over independent units: a rank owns a
contiguous block of units and never exchanges samples with another rank.  The only
cross-rank traffic is control: a barrier and the max-over-ranks of a device time.
"""
from __future__ import annotations


def unit_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of `n_units` owned by `rank`: ceil(n/world) units per rank, the
    last ranks may own fewer (or none).  Same rule as run_host_batch() in csrc/j2k_b200.cu."""
    if n_units < 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    per = (n_units + world - 1) // world
    b = min(rank * per, n_units)
    return b, min(b + per, n_units)


def shard_sizes(n_units: int, world: int) -> list[int]:
    return [unit_range(n_units, r, world)[1] - unit_range(n_units, r, world)[0] for r in range(world)]


def max_over_ranks(value: float, dist=None) -> float:
    """Max of a per-rank device time over the job (timing rule: never wall clock, max over ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: list[int], ms_max: float) -> float:
    """Whole-job units per second: every rank's units over the slowest rank's time."""
    return sum(units_per_rank) / (ms_max * 1e-3)
