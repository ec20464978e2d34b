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


class ShardedProof:
    """Host-side glue for ONE proof whose resident batches are sharded by row range over the ranks
    of a torch.distributed group (Context.set_shard; NCCL when `device` is a CUDA device, gloo on
    the CPU): completes the caps the shards return, and collects the query openings each rank serves
    from its own rows.  The exchanged data is 32 bytes per cap entry per commit and the opened rows +
    Merkle paths (a few tens of KB per proof)."""

    def __init__(self, rank: int, world: int, device=None):
        import torch
        self.rank, self.world = rank, world
        self.device = device if device is not None else torch.device("cpu")
        self._buf = {}

    def _tensor(self, key, numel):
        import torch
        t = self._buf.get(key)
        if t is None or t.numel() != numel:
            t = self._buf[key] = torch.empty(numel, dtype=torch.int64, device=self.device)
        return t

    def _before_library_call(self, ctx):
        """torch / NCCL work is ordered on torch's current stream; the library runs on the context's.
        When they differ, what torch enqueued must be finished before the library reads it (library
        calls themselves return synchronised)."""
        import torch
        if self.device.type != "cuda":
            return
        cur = torch.cuda.current_stream(self.device)
        # ctx.stream == 0 means "the context's own stream" (never torch's, whose legacy stream is also 0)
        if ctx.stream == 0 or cur.cuda_stream != ctx.stream:
            cur.synchronize()

    def complete_cap(self, cap):
        """cap: (ncap, 4) uint64 numpy array as a sharded commit returns it (own entries filled, the
        rest zero), completed IN PLACE with the other shards' entries (all-gather).  ncap must be a
        multiple of the world size (a shard is whole cap subtrees)."""
        import numpy as np
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return cap
        ncap = cap.shape[0]
        if ncap % self.world:
            raise ValueError("cap entries (%d) must be a multiple of the world size" % ncap)
        own = ncap // self.world
        mine = torch.from_numpy(cap[self.rank * own:(self.rank + 1) * own].view(np.int64).reshape(-1))
        full = self._tensor(("cap", ncap), ncap * 4)
        dist.all_gather_into_tensor(full, mine.to(self.device, non_blocking=False))
        cap[:] = full.cpu().numpy().view(np.uint64).reshape(ncap, 4)
        return cap

    def commit_from_host(self, ctx, host_cols, rate_bits: int, cap_height: int,
                         inputs_are_coeffs: bool = False):
        """PolynomialBatch::from_values / from_coeffs of one proof sharded over the ranks, from host
        columns every rank can read (host_cols: (ncols, n) uint64, ideally page-locked).  Every shard
        needs ALL columns, but they cross PCIe only once: rank r uploads columns
        [r * C' / world, (r + 1) * C' / world) and the ranks exchange them GPU to GPU (NCCL
        all-gather over NVLink), then each commits its own row range from device memory
        (vpbs_batch_commit_dev under Context.set_shard) and the cap is completed by the all-gather of
        the subtree roots.  Returns (batch handle, full cap).  Fastest when the context runs on torch's
        current stream (Context.set_stream): copies, collective and kernels are then ordered on one
        stream; otherwise the torch side is synchronised before the library reads its results."""
        import ctypes
        import numpy as np
        import torch
        import torch.distributed as dist
        from . import _lib
        ncols, n = host_cols.shape
        log_n = n.bit_length() - 1
        per = -(-ncols // self.world)                       # columns per rank, last ranks may be short
        d_all = self._tensor(("cols", per * self.world, n), per * self.world * n).view(per * self.world, n)
        c0, c1 = min(self.rank * per, ncols), min((self.rank + 1) * per, ncols)
        mine = d_all[self.rank * per:(self.rank + 1) * per]
        if c1 > c0:
            mine[:c1 - c0].copy_(torch.from_numpy(host_cols[c0:c1].view(np.int64)), non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), mine.reshape(-1))
        cap = np.empty((1 << cap_height, 4), np.uint64)
        h = ctypes.c_void_p()
        self._before_library_call(ctx)
        ctx.check(ctx.lib.vpbs_batch_commit_dev(ctx.handle, d_all.data_ptr(), ncols, log_n, rate_bits,
                                                cap_height, int(inputs_are_coeffs),
                                                cap.ctypes.data_as(_lib.u64p), ctypes.byref(h), None))
        return h, self.complete_cap(cap)

    def quotient_polys(self, ctx, constants_sigmas, sigmas_first_col: int, wires, zs_pp, k_is,
                       max_degree: int, quotient_degree_bits: int, betas, gammas, alphas, rate_bits: int,
                       cap_height: int, log_n: int, gate_terms=None, program=None, public_inputs_hash=None):
        """prove() steps 6-7 with sharded batches (handles as vpbs_batch_commit* returned them): every
        rank computes the quotient values of its own rows (vpbs_batch_quotient_values), the buffers are
        added up across the ranks (all-reduce over NCCL / NVLink: 8 bytes per challenge and point), and
        every rank commits its shard of the quotient batch (vpbs_quotient_commit_values); the cap is
        completed like every sharded commit's.  Returns (batch handle, full cap)."""
        import ctypes
        import numpy as np
        import torch.distributed as dist
        from . import _lib
        from .plonky2_api import _as_u64, _ptr
        u64p = _lib.u64p
        k = _as_u64(k_is).reshape(-1)
        b, g, a = (_as_u64(v).reshape(-1) for v in (betas, gammas, alphas))
        nc, q = b.size, (1 << log_n) << quotient_degree_bits
        d_vals = self._tensor(("qvals", nc, q), nc * q)
        gt_arr, gtp = None, None
        if gate_terms is not None:
            gt_arr = _as_u64(gate_terms)
            gtp = (u64p * nc)(*[_ptr(gt_arr[c]) for c in range(nc)])
        h = lambda x: x.handle if hasattr(x, "handle") else x
        ctx.check(ctx.lib.vpbs_batch_quotient_values(
            h(constants_sigmas), sigmas_first_col, h(wires), h(zs_pp), _ptr(k), k.size, max_degree,
            quotient_degree_bits, _ptr(b), _ptr(g), _ptr(a), nc, gtp,
            program.handle if program is not None else None,
            _ptr(_as_u64(public_inputs_hash).reshape(4)) if public_inputs_hash is not None else None,
            d_vals.data_ptr()))
        if self.world > 1:
            dist.all_reduce(d_vals)     # int64 wrap-around addition of a value and zeros: exact
        cap = np.empty((1 << cap_height, 4), np.uint64)
        out = ctypes.c_void_p()
        self._before_library_call(ctx)
        ctx.check(ctx.lib.vpbs_quotient_commit_values(ctx.handle, d_vals.data_ptr(), nc, log_n,
                                                      quotient_degree_bits, rate_bits, cap_height,
                                                      cap.ctypes.data_as(u64p), ctypes.byref(out), None))
        return out, self.complete_cap(cap)

    def owned(self, leaf_indices, nleaves_total: int):
        """Boolean mask of the leaf indices this rank's shard holds."""
        import numpy as np
        per = nleaves_total // self.world
        idx = np.asarray(leaf_indices, dtype=np.uint64)
        return (idx // np.uint64(per)) == np.uint64(self.rank)

    def collect(self, arrays):
        """arrays: uint64 numpy arrays in which every rank has filled the entries it owns and left
        the others ZERO.  Sums them across the ranks in place (all-reduce; integer addition of a
        value and zeros is exact), so that afterwards every rank holds every entry."""
        import numpy as np
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return arrays
        total = sum(a.size for a in arrays)
        t = self._tensor(("collect", total), total)
        flat = np.concatenate([a.reshape(-1).view(np.int64) for a in arrays])
        t.copy_(torch.from_numpy(flat))
        dist.all_reduce(t)
        out = t.cpu().numpy().view(np.uint64)
        off = 0
        for a in arrays:
            a[...] = out[off:off + a.size].reshape(a.shape)
            off += a.size
        return arrays


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
