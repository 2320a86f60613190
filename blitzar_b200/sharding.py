"""Host-side sharding logic for the one-process-per-GPU layout.

MSM is linear, so the generator range [0, n) is split into `world_size` contiguous shards; rank r
computes a complete local MSM over its shard (its own buckets and reduction) and produces ONE
partial accumulator point per output column. The only exchange step is an all-gather of those
partials (<= 64 x 160 B), after which the partials are summed and canonicalised. NCCL has no
group-law reduction, hence all-gather + local point adds (never all-reduce).

The reference does the same split with host staging (sxt/multiexp/bucket_method/accumulation.h:
110-163, sxt/multiexp/pippenger2/multiexponentiation.h:105-137, sxt/base/iterator/split.cc:28-39).

The arithmetic is supplied by the caller (`partial_fn`, `combine_fn`): the CUDA library on a GPU
box, the CPU emulation harness in the gloo tests.
"""
import numpy as np


def shard_range(n, rank, world_size):
    """Contiguous, balanced [begin, end) of rank's generator range."""
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_columns(columns, rank, world_size):
    """columns: list of (uint8 [n_j, nbytes], is_signed). Rows [begin, end) of every column; the
    shard boundaries are those of the longest column so that generator i stays paired with term i."""
    n = max((c[0].shape[0] for c in columns), default=0)
    begin, end = shard_range(n, rank, world_size)
    out = []
    for data, is_signed in columns:
        lo, hi = min(begin, data.shape[0]), min(end, data.shape[0])
        out.append((data[lo:hi], is_signed))
    return out, begin, end


def sharded_commit(columns, generators, rank, world_size, point_bytes, partial_fn, combine_fn,
                   all_gather_fn):
    """Runs rank's share and returns the combined commitments (same on every rank).

    partial_fn(columns_shard, generators_shard) -> uint8 [num_columns, point_bytes]
    all_gather_fn(uint8 array) -> uint8 [world_size, ...]
    combine_fn(uint8 [world_size * num_columns, point_bytes], world_size, num_columns) -> commitments
    """
    shard, begin, end = shard_columns(columns, rank, world_size)
    gens = generators[begin:end] if generators is not None else None
    partial = partial_fn(shard, gens, begin)
    assert partial.shape == (len(columns), point_bytes)
    gathered = all_gather_fn(np.ascontiguousarray(partial))
    return combine_fn(gathered.reshape(world_size * len(columns), point_bytes), world_size,
                      len(columns))
