"""Env sharding across ranks (one process per GPU): contiguous global id ranges, sizes differ by at most 1.

Env instances share no state (rsrl_domains/src/mountain_car/discrete.rs:50-53), RNG draws are keyed by the global
env id, so the union of the shards is bit-identical to a single-engine run in PER_ENV mode and identical up to the
summation order of dW in SHARED mode."""


def shard_range(n_global, rank, world):
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
