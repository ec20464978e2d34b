"""Row-range sharding of ONE commit across GPUs (partitioning B).

Leaf k of the tree is natural LDE row bitrev(k), so the n-row leaf block b is the evaluation of
every polynomial on one sub-coset: a self-contained size-n transform of the coefficients.  Each
rank therefore produces its own contiguous range of leaves (already in leaf order), the digests of
the cap subtrees below it and their roots; the only exchange is an all-gather of those roots
(32 bytes per cap entry).  Replaces nothing in plonky2 (which is single-process); it is how the
"build Merkle tree" + "FFT + blinding" scopes of [P2] fri/oracle.rs from_coeffs spread over an
8xB200 box.  Results are bit-identical to the single-GPU commit.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    first_leaf: int      # first leaf (row of the leaf matrix) owned by this rank
    nleaves: int         # leaves owned
    first_cap: int       # first cap entry owned
    ncap: int            # cap entries (subtree roots) owned
    digest_offset: int   # offset (in hashes) of this rank's digests inside the global buffer
    ndigests: int        # hashes of digests owned


def shard_plan(log_n: int, rate_bits: int, cap_height: int, rank: int, world: int) -> ShardPlan:
    """Which rows / cap entries / digests rank `rank` of `world` owns.  Raises ValueError when the
    commit cannot be split that way (shards must be whole n-row LDE blocks and whole cap subtrees)."""
    if world <= 0 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    log_m = log_n + rate_bits
    if cap_height > log_m:
        raise ValueError("cap_height should be at most log2(leaves.len())")
    lw = world.bit_length() - 1
    if lw > rate_bits:
        raise ValueError("at most 2^rate_bits = %d shards (one n-row LDE block each)" % (1 << rate_bits))
    if lw > cap_height:
        raise ValueError("at most 2^cap_height = %d shards (one cap subtree each)" % (1 << cap_height))
    m = 1 << log_m
    nleaves = m >> lw
    ncap = (1 << cap_height) >> lw
    sub_digests = 2 * (m >> cap_height) - 2      # hashes per cap subtree
    return ShardPlan(rank, world, rank * nleaves, nleaves, rank * ncap, ncap,
                     rank * ncap * sub_digests, ncap * sub_digests)


def gather_cap(local_roots, world: int):
    """all_gather of the per-rank subtree roots -> the full cap (ncap_total x 4).  `local_roots` is a
    torch int64 tensor (ncap_local x 4) on the rank's device (NCCL) or on the CPU (gloo)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_roots.clone()
    out = torch.empty((world * local_roots.shape[0], 4), dtype=local_roots.dtype,
                      device=local_roots.device)
    dist.all_gather_into_tensor(out, local_roots.contiguous())
    return out


def commit_sharded(ctx, d_cols, ncols: int, log_n: int, rate_bits: int, cap_height: int,
                   inputs_are_coeffs: bool, rank: int, world: int, want_stats: bool = False):
    """One rank's part of a sharded commit on device tensors.  d_cols: torch int64 (ncols, n) on this
    rank's GPU holding ALL columns.  Returns (plan, leaves, digests, cap, coeffs, stats): leaves and
    digests are this rank's shard, cap is the full gathered cap.

    Stream contract: the kernels run on torch's CURRENT stream of d_cols' device (the context is
    switched to it for the call and switched back afterwards), so they are ordered after whatever
    produced d_cols and before the NCCL all_gather of the roots, which torch enqueues on / orders
    against that same stream.  If torch's current stream is the legacy default stream (handle 0,
    which the context cannot adopt), the call synchronises explicitly on both sides instead."""
    import torch
    from .plonky2_api import commit_shard_device
    plan = shard_plan(log_n, rate_bits, cap_height, rank, world)
    dev = d_cols.device
    n = 1 << log_n
    cur = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
    prev = ctx.stream
    if cur:
        ctx.set_stream(cur)
    else:
        torch.cuda.current_stream(dev).synchronize()  # d_cols is final before our own stream reads it
    coeffs = torch.empty((ncols, n), dtype=torch.int64, device=dev)
    leaves = torch.empty((plan.nleaves, ncols), dtype=torch.int64, device=dev)
    digests = torch.empty((max(plan.ndigests, 1), 4), dtype=torch.int64, device=dev)
    roots = torch.empty((plan.ncap, 4), dtype=torch.int64, device=dev)
    try:
        stats = commit_shard_device(ctx, d_cols.data_ptr(), ncols, log_n, rate_bits, cap_height,
                                    inputs_are_coeffs, plan.first_leaf, plan.nleaves,
                                    coeffs.data_ptr(), leaves.data_ptr(),
                                    digests.data_ptr() if plan.ndigests else 0, roots.data_ptr(),
                                    want_stats=want_stats)
        if not cur:
            ctx.sync()  # own stream: the roots are final before NCCL reads them
    finally:
        if cur:
            ctx.set_stream(prev)
    cap = gather_cap(roots, world)
    return plan, leaves, digests[:plan.ndigests], cap, coeffs, stats
