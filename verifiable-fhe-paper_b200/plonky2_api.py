"""Host-side mirror of the plonky2 0.2.0 interface for the commitment path, over the C ABI.

Names, argument meaning and failure behaviour follow upstream so the parity tests read like
plonky2 call sites (the reference reaches them through prove()/build(),
/root/reference/src/vtfhe/ivc_based_vpbs.rs:275,302,333,364):

  [P2] plonky2_field/src/fft.rs            fft, ifft, FftRootTable (kept on the device here)
  [P2] plonky2_field/src/polynomial/mod.rs PolynomialCoeffs::coset_fft / lde
  [P2] plonky2/src/hash/merkle_tree.rs     MerkleTree::{new, get, prove}, MerkleCap
  [P2] plonky2/src/fri/oracle.rs           PolynomialBatch::{from_values, from_coeffs,
                                           get_lde_values}, SALT_SIZE

Everything numeric is a numpy uint64 array (GoldilocksField is a transparent u64).  Where plonky2
`assert!`s/panics, these raise (ValueError for argument errors, VpbsError for device errors).
There is no CPU implementation behind any of it: every call goes to libvpbs_commit.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import VpbsError, VpbsStats, u64p, u64pp

P = 0xFFFFFFFF00000001
SALT_SIZE = 4
COSET_SHIFT = 7


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _as_u64(x) -> np.ndarray:
    a = np.asarray(x)
    if a.dtype != np.uint64:
        a = a.astype(np.uint64)
    return np.ascontiguousarray(a)


def log2_strict(n: int) -> int:
    """[P2] plonky2_util::log2_strict — panics (here: ValueError) unless n is a power of two."""
    if n <= 0 or n & (n - 1):
        raise ValueError("Not a power of two: %d" % n)
    return n.bit_length() - 1


def reverse_bits(x: int, bits: int) -> int:
    """[P2] plonky2_util::reverse_bits."""
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


class Context:
    """One vpbs_ctx: a device, a stream and the grow-only device arena reused across commits."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.vpbs_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise VpbsError(rc, (self._L.vpbs_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._stream = 0  # 0 = the context's own stream

    def close(self):
        if getattr(self, "_h", None):
            self._L.vpbs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def check(self, rc: int):
        if rc == 0:
            return
        msg = (self._L.vpbs_last_error(self._h) or b"").decode()
        if rc == _lib.VPBS_ERR_ARG:
            raise ValueError(msg)
        raise VpbsError(rc, msg)

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._L

    def set_stream(self, cuda_stream: int):
        """Run this context's kernels on an existing CUDA stream (0 / None = its own stream)."""
        self.check(self._L.vpbs_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))
        self._stream = cuda_stream or 0

    @property
    def stream(self) -> int:
        return self._stream


    def sync(self):
        self.check(self._L.vpbs_ctx_sync(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.vpbs_ctx_kernel_launches(self._h))

    def set_shard(self, index: int, count: int):
        """Row-range sharding of one proof over `count` GPUs (vpbs_ctx_set_shard): resident batches
        created afterwards hold leaves [index * m / count, (index + 1) * m / count); their caps carry
        only the own entries (others zero) until gathered across the shards."""
        self.check(self._L.vpbs_ctx_set_shard(self._h, int(index), int(count)))

    def set_host_threads(self, threads: int):
        """Copy threads of the pinned staging ring that pageable host columns travel through
        (vpbs_ctx_set_host_threads; 0: leave such copies to the driver)."""
        self.check(self._L.vpbs_ctx_set_host_threads(self._h, int(threads)))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ----------------------------------------------------------------------------- fft.rs
def fft(coeffs, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] fft(poly): values[i] = poly(w^i), natural order."""
    ctx = ctx or default_context()
    a = _as_u64(coeffs).copy()
    ctx.check(ctx.lib.vpbs_fft(ctx.handle, _ptr(a), log2_strict(a.size)))
    return a


def ifft(values, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] ifft(poly): coefficients of the interpolant over <w>."""
    ctx = ctx or default_context()
    a = _as_u64(values).copy()
    ctx.check(ctx.lib.vpbs_ifft(ctx.handle, _ptr(a), log2_strict(a.size)))
    return a


def coset_fft(coeffs, shift: int = COSET_SHIFT, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] PolynomialCoeffs::coset_fft(shift): values[i] = poly(shift * w^i)."""
    ctx = ctx or default_context()
    a = _as_u64(coeffs).copy()
    ctx.check(ctx.lib.vpbs_coset_fft(ctx.handle, _ptr(a), log2_strict(a.size), int(shift) % 2**64))
    return a


def lde_values(polys, rate_bits: int, inputs_are_coeffs: bool = True,
               ctx: Optional[Context] = None):
    """[P2] PolynomialBatch::lde_values without blinding: (ncols, n << rate_bits), natural order.
    Returns (coeffs, lde)."""
    ctx = ctx or default_context()
    a = _as_u64(polys)
    ncols, n = a.shape
    log_n = log2_strict(n)
    colp = (u64p * ncols)(*[_ptr(a[c]) for c in range(ncols)])
    coeffs = np.empty((ncols, n), np.uint64)
    cop = (u64p * ncols)(*[_ptr(coeffs[c]) for c in range(ncols)])
    out = np.empty((ncols, n << rate_bits), np.uint64)
    ctx.check(ctx.lib.vpbs_lde_batch(ctx.handle, colp, ncols, log_n, rate_bits,
                                     int(inputs_are_coeffs), cop, _ptr(out)))
    return coeffs, out


# ----------------------------------------------------------------------------- hashing
def poseidon(states, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] Poseidon::poseidon on a (count, 12) batch of states."""
    ctx = ctx or default_context()
    a = _as_u64(states).reshape(-1, 12).copy()
    ctx.check(ctx.lib.vpbs_poseidon_permute(ctx.handle, _ptr(a), a.shape[0]))
    return a


def hash_or_noop(rows, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] Hasher::hash_or_noop on every row of a (count, len) matrix -> (count, 4)."""
    ctx = ctx or default_context()
    a = _as_u64(rows)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    out = np.empty((a.shape[0], 4), np.uint64)
    ctx.check(ctx.lib.vpbs_hash_or_noop_batch(ctx.handle, _ptr(a) if a.size else None, a.shape[0],
                                              a.shape[1], _ptr(out)))
    return out


def two_to_one(left, right, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] PoseidonHash::two_to_one on (count, 4) batches."""
    ctx = ctx or default_context()
    l, r = _as_u64(left).reshape(-1, 4), _as_u64(right).reshape(-1, 4)
    if l.shape != r.shape:
        raise ValueError("left/right shape mismatch")
    out = np.empty_like(l)
    ctx.check(ctx.lib.vpbs_two_to_one_batch(ctx.handle, _ptr(l), _ptr(r), l.shape[0], _ptr(out)))
    return out


def eval_ext2(coeff_cols, points, ctx: Optional[Context] = None) -> np.ndarray:
    """[P2] PolynomialCoeffs::to_extension().eval(point) for every polynomial and every point of
    F[X]/(X^2 - 7): coeff_cols (ncols, n), points (npoints, 2) -> (npoints, ncols, 2)."""
    ctx = ctx or default_context()
    a = _as_u64(coeff_cols)
    pts = _as_u64(points).reshape(-1, 2)
    ncols, n = a.shape
    colp = (u64p * ncols)(*[_ptr(a[c]) for c in range(ncols)])
    out = np.empty((pts.shape[0], ncols, 2), np.uint64)
    ctx.check(ctx.lib.vpbs_eval_ext2(ctx.handle, colp, ncols, log2_strict(n), _ptr(pts), pts.shape[0],
                                     _ptr(out)))
    return out


def fri_layer_commit(values_ext, arity_bits: int, cap_height: int,
                     ctx: Optional[Context] = None) -> MerkleTree:
    """[P2] fri/prover.rs fri_committed_trees, the tree of one layer: values_ext (len, 2)."""
    ctx = ctx or default_context()
    v = _as_u64(values_ext).reshape(-1, 2)
    ln = v.shape[0]
    lg = log2_strict(ln)
    if arity_bits > lg or cap_height > lg - arity_bits:
        raise ValueError("cap_height should be at most log2(leaves.len())")
    nl = ln >> arity_bits
    leaves = np.empty((nl, 2 << arity_bits), np.uint64)
    ndig = 2 * (nl - (1 << cap_height))
    digests = np.empty((ndig, 4), np.uint64)
    cap = np.empty((1 << cap_height, 4), np.uint64)
    ctx.check(ctx.lib.vpbs_fri_layer_commit(ctx.handle, _ptr(v), ln, arity_bits, cap_height,
                                            _ptr(leaves), _ptr(digests) if ndig else None, _ptr(cap)))
    return MerkleTree(leaves, digests, cap)


def fri_fold(coeffs_ext, arity_bits: int, beta, shift_next: int, ctx: Optional[Context] = None):
    """[P2] fri_committed_trees fold: (coeffs' (len/arity, 2), values' = coeffs'.coset_fft(shift_next))."""
    ctx = ctx or default_context()
    c = _as_u64(coeffs_ext).reshape(-1, 2)
    ln = c.shape[0]
    log2_strict(ln)
    ol = ln >> arity_bits
    co, vo = np.empty((ol, 2), np.uint64), np.empty((ol, 2), np.uint64)
    ctx.check(ctx.lib.vpbs_fri_fold(ctx.handle, _ptr(c), ln, arity_bits, _ptr(_as_u64(beta).reshape(2)),
                                    int(shift_next) % 2**64, _ptr(co), _ptr(vo)))
    return co, vo


class FriCommitPhase:
    """[P2] fri/prover.rs fri_committed_trees as one device-resident chain (vpbs_fri_*): the
    polynomial and every layer's tree stay in HBM; the caller (the challenger) sees caps and supplies
    betas.  final_poly_coeffs: (ncoeffs, 2) extension coefficients, not yet padded."""

    def __init__(self, final_poly_coeffs, rate_bits: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        c = _as_u64(final_poly_coeffs).reshape(-1, 2)
        log2_strict(c.shape[0])
        self.rate_bits = rate_bits
        self.len = c.shape[0] << rate_bits
        self.layers = []  # (nleaves, arity_bits, cap_height)
        self.handle = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.vpbs_fri_begin(self.ctx.handle, _ptr(c), c.shape[0], rate_bits,
                                                   ctypes.byref(self.handle)))

    @classmethod
    def from_openings(cls, oracles, batches, points, alpha, rate_bits: int) -> "FriCommitPhase":
        """[P2] fri/oracle.rs PolynomialBatch::prove_openings up to lde_final_values, on the device
        (vpbs_fri_begin_openings).  oracles: ResidentPolynomialBatch list; batches: per FRI batch the
        list of (oracle_index, polynomial_index) ([P2] FriBatchInfo.polynomials); points: one extension
        point (re, im) per batch; alpha: the extension challenge."""
        self = cls.__new__(cls)
        self.ctx = oracles[0].ctx
        sizes = np.array([len(b) for b in batches], np.uint32)
        refs = np.array([r for b in batches for r in b], np.uint32).reshape(-1, 2)
        pts = _as_u64(points).reshape(-1, 2)
        if pts.shape[0] != len(batches) or sizes.size == 0 or (sizes == 0).any():
            raise ValueError("one opening point per non-empty FRI batch")
        hs = (ctypes.c_void_p * len(oracles))(*[o.handle for o in oracles])
        self.rate_bits = rate_bits
        self.len = (1 << oracles[0].degree_log) << rate_bits
        self.layers = []
        self.handle = ctypes.c_void_p()
        u32p = ctypes.POINTER(ctypes.c_uint32)
        self.ctx.check(self.ctx.lib.vpbs_fri_begin_openings(
            self.ctx.handle, hs, len(oracles), sizes.ctypes.data_as(u32p), sizes.size,
            np.ascontiguousarray(refs).ctypes.data_as(u32p), _ptr(pts), _ptr(_as_u64(alpha).reshape(2)),
            rate_bits, ctypes.byref(self.handle)))
        return self

    def commit_layer(self, arity_bits: int, cap_height: int) -> np.ndarray:
        cap = np.empty((1 << cap_height, 4), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_fri_commit_layer(self.handle, arity_bits, cap_height, _ptr(cap)))
        self.layers.append((self.len >> arity_bits, arity_bits, cap_height))
        return cap

    def fold(self, beta):
        self.ctx.check(self.ctx.lib.vpbs_fri_fold_layer(self.handle, _ptr(_as_u64(beta).reshape(2))))
        self.len >>= self.layers[-1][1]

    def final_poly(self) -> np.ndarray:
        out = np.empty((self.len >> self.rate_bits, 2), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_fri_final_poly(self.handle, self.rate_bits, _ptr(out)))
        return out

    def query(self, layer: int, leaf_indices):
        """(rows (count, 2 << arity_bits), siblings (count, layers, 4)) of one layer's tree."""
        nleaves, arity_bits, cap_height = self.layers[layer]
        idx = _as_u64(leaf_indices).reshape(-1)
        rows = np.empty((idx.size, 2 << arity_bits), np.uint64)
        sib = np.empty((idx.size, log2_strict(nleaves) - cap_height, 4), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_fri_query_layer(self.handle, layer, _ptr(idx), idx.size, _ptr(rows),
                                                         _ptr(sib) if sib.size else None))
        return rows, sib

    def close(self):
        if self.handle:
            self.ctx.lib.vpbs_fri_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fri_proof_of_work(state, witness_pos: int, min_leading_zeros: int, first_candidate: int = 0,
                      count: int = 1 << 32, response_lane: int = 7, ctx: Optional[Context] = None):
    """[P2] fri/prover.rs fri_proof_of_work on a duplex state: smallest witness in
    [first_candidate, first_candidate + count) whose response has >= min_leading_zeros leading
    zero bits, or None."""
    ctx = ctx or default_context()
    st = _as_u64(state).reshape(12)
    out = np.zeros(1, np.uint64)
    found = ctypes.c_int(0)
    ctx.check(ctx.lib.vpbs_pow_grind(ctx.handle, _ptr(st), witness_pos, response_lane,
                                     min_leading_zeros, first_candidate, count, _ptr(out),
                                     ctypes.byref(found)))
    return int(out[0]) if found.value else None


# ----------------------------------------------------------------------------- merkle_tree.rs
@dataclass
class MerkleProof:
    """[P2] hash/merkle_proofs.rs MerkleProof { siblings }."""
    siblings: np.ndarray  # (num_layers, 4)


class MerkleTree:
    """[P2] hash/merkle_tree.rs MerkleTree { leaves, digests, cap }."""

    def __init__(self, leaves: np.ndarray, digests: np.ndarray, cap: np.ndarray):
        self.leaves = leaves    # (nleaves, leaf_len), row k = leaf k
        self.digests = digests  # (2 * (nleaves - 2^cap_height), 4), plonky2 layout
        self.cap = cap          # (2^cap_height, 4)  == MerkleCap.0

    @classmethod
    def new(cls, leaves, cap_height: int, ctx: Optional[Context] = None) -> "MerkleTree":
        ctx = ctx or default_context()
        a = _as_u64(leaves)
        if a.ndim != 2:
            raise ValueError("leaves must be a (nleaves, leaf_len) matrix")
        nleaves, leaf_len = a.shape
        lg = log2_strict(nleaves)
        if cap_height > lg:
            # same condition plonky2 asserts on in MerkleTree::new
            raise ValueError("cap_height=%d should be at most log2(leaves.len())=%d"
                             % (cap_height, lg))
        ndig = 2 * (nleaves - (1 << cap_height))
        digests = np.empty((ndig, 4), np.uint64)
        cap = np.empty((1 << cap_height, 4), np.uint64)
        ctx.check(ctx.lib.vpbs_merkle_new(ctx.handle, _ptr(a) if a.size else None, nleaves,
                                          leaf_len, cap_height, _ptr(digests) if ndig else None,
                                          _ptr(cap)))
        return cls(a, digests, cap)

    def get(self, i: int) -> np.ndarray:
        return self.leaves[i]

    def prove(self, leaf_index: int) -> MerkleProof:
        """[P2] MerkleTree::prove — pure index arithmetic over `digests` (unchanged layout)."""
        cap_height = log2_strict(self.cap.shape[0])
        num_layers = log2_strict(self.leaves.shape[0]) - cap_height
        if leaf_index >> (cap_height + num_layers):
            raise ValueError("leaf_index out of range")
        tree_index = leaf_index >> num_layers
        tree_len = self.digests.shape[0] >> cap_height
        tree = self.digests[tree_len * tree_index: tree_len * (tree_index + 1)]
        pair_index = leaf_index & ((1 << num_layers) - 1)
        sibs = np.empty((num_layers, 4), np.uint64)
        for i in range(num_layers):
            parity = pair_index & 1
            pair_index >>= 1
            siblings_index = (pair_index << (i + 1)) + (1 << i) - 1
            sibs[i] = tree[2 * siblings_index + (1 - parity)]
        return MerkleProof(sibs)


def verify_merkle_proof_to_cap(leaf, leaf_index: int, cap: np.ndarray, proof: MerkleProof,
                               ctx: Optional[Context] = None) -> bool:
    """[P2] hash/merkle_proofs.rs verify_merkle_proof_to_cap (hashing done on the device)."""
    cur = hash_or_noop(np.asarray(leaf, dtype=np.uint64).reshape(1, -1), ctx)[0]
    idx = leaf_index
    for sib in proof.siblings:
        cur = (two_to_one(sib, cur, ctx) if idx & 1 else two_to_one(cur, sib, ctx))[0]
        idx >>= 1
    return bool(np.array_equal(cur, cap[idx]))


# ----------------------------------------------------------------------------- fri/oracle.rs
class PolynomialBatch:
    """[P2] fri/oracle.rs PolynomialBatch { polynomials, merkle_tree, degree_log, rate_bits,
    blinding }."""

    def __init__(self, polynomials, merkle_tree, degree_log, rate_bits, blinding, stats=None):
        self.polynomials = polynomials  # (ncols, n) coefficients
        self.merkle_tree = merkle_tree
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.blinding = blinding
        self.stats = stats              # per-phase timings (plonky2's TimingTree scopes)

    @classmethod
    def from_values(cls, values, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, *, ctx: Optional[Context] = None, salt=None,
                    rng: Optional[np.random.Generator] = None, ctxs=None) -> "PolynomialBatch":
        """values: (ncols, n) evaluations over <w_n>.  `timing` / `fft_root_table` are accepted for
        signature parity (timings come back in .stats; root tables are cached on the device).
        ctxs: a list of Contexts on distinct GPUs spreads this one commit over them
        (vpbs_commit_multi: row ranges per GPU, no GPU-to-GPU traffic, identical outputs)."""
        return cls._commit(values, rate_bits, blinding, cap_height, False, ctx, salt, rng, ctxs)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, *, ctx: Optional[Context] = None, salt=None,
                    rng: Optional[np.random.Generator] = None, ctxs=None) -> "PolynomialBatch":
        return cls._commit(polynomials, rate_bits, blinding, cap_height, True, ctx, salt, rng, ctxs)

    @classmethod
    def _commit(cls, cols, rate_bits, blinding, cap_height, are_coeffs, ctx, salt, rng, ctxs=None):
        if ctxs is not None:
            ctxs = list(ctxs)
            if not ctxs:
                raise ValueError("ctxs must not be empty")
            ctx = ctxs[0]
        ctx = ctx or default_context()
        a = _as_u64(cols)
        if a.ndim != 2 or a.shape[0] == 0:
            raise ValueError("need a non-empty (ncols, n) matrix")
        ncols, n = a.shape
        log_n = log2_strict(n)
        m = n << rate_bits
        if cap_height > log_n + rate_bits:
            raise ValueError("cap_height=%d should be at most log2(leaves.len())=%d"
                             % (cap_height, log_n + rate_bits))
        salt_arr, saltp = None, None
        if blinding:
            # [P2] lde_values: SALT_SIZE extra columns of F::rand_vec(m); RNG stays on the host
            if salt is None:
                rng = rng or np.random.default_rng()
                salt = rng.integers(0, P, size=(SALT_SIZE, m), dtype=np.uint64)
            salt_arr = _as_u64(salt)
            if salt_arr.shape != (SALT_SIZE, m):
                raise ValueError("salt must be (%d, %d)" % (SALT_SIZE, m))
            saltp = (u64p * SALT_SIZE)(*[_ptr(salt_arr[s]) for s in range(SALT_SIZE)])
        width = ncols + (SALT_SIZE if blinding else 0)
        colp = (u64p * ncols)(*[_ptr(a[c]) for c in range(ncols)])
        coeffs = np.empty((ncols, n), np.uint64)
        cop = (u64p * ncols)(*[_ptr(coeffs[c]) for c in range(ncols)])
        leaves = np.empty((m, width), np.uint64)
        ndig = 2 * (m - (1 << cap_height))
        digests = np.empty((ndig, 4), np.uint64)
        cap = np.empty((1 << cap_height, 4), np.uint64)
        st = VpbsStats()
        if ctxs is not None:
            handles = (ctypes.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
            ctx.check(ctx.lib.vpbs_commit_multi(handles, len(ctxs), colp, ncols, log_n, rate_bits,
                                                cap_height, int(are_coeffs), saltp, cop, _ptr(leaves),
                                                _ptr(digests) if ndig else None, _ptr(cap),
                                                ctypes.byref(st)))
        else:
            ctx.check(ctx.lib.vpbs_commit(ctx.handle, colp, ncols, log_n, rate_bits, cap_height,
                                          int(are_coeffs), saltp, cop, _ptr(leaves),
                                          _ptr(digests) if ndig else None, _ptr(cap),
                                          ctypes.byref(st)))
        if are_coeffs:
            coeffs = a % np.uint64(P) if (a >= np.uint64(P)).any() else a
        return cls(coeffs, MerkleTree(leaves, digests, cap), log_n, rate_bits, blinding,
                   st.as_dict())

    def get_lde_values(self, index: int, step: int = 1) -> np.ndarray:
        """[P2] PolynomialBatch::get_lde_values: leaf reverse_bits(index * step) minus the salt."""
        index = index * step
        index = reverse_bits(index, self.degree_log + self.rate_bits)
        row = self.merkle_tree.get(index)
        return row[: row.shape[0] - (SALT_SIZE if self.blinding else 0)]


# ----------------------------------------------------------------------------- resident batches
class ResidentMerkleTree:
    """MerkleTree whose leaves and digests stay in HBM (vpbs_batch_*): `cap` is on the host, `get` /
    `prove` fetch single rows / authentication paths on demand — what plonky2's FRI query phase
    reads ([P2] fri/prover.rs: merkle_tree.get(i) + merkle_tree.prove(i), 28 queries per batch)."""

    def __init__(self, batch, cap, nleaves, width):
        self._b = batch
        self.cap = cap
        self.nleaves = nleaves
        self.width = width

    def get_many(self, indices) -> np.ndarray:
        idx = _as_u64(indices).reshape(-1)
        out = np.empty((idx.size, self.width), np.uint64)
        b = self._b
        b.ctx.check(b.ctx.lib.vpbs_batch_get_leaves(b.handle, _ptr(idx), idx.size, _ptr(out)))
        return out

    def get(self, i: int) -> np.ndarray:
        return self.get_many([i])[0]

    def prove_many(self, indices) -> List[MerkleProof]:
        idx = _as_u64(indices).reshape(-1)
        layers = log2_strict(self.nleaves) - log2_strict(self.cap.shape[0])
        out = np.empty((idx.size, layers, 4), np.uint64)
        b = self._b
        b.ctx.check(b.ctx.lib.vpbs_batch_prove(b.handle, _ptr(idx), idx.size,
                                               _ptr(out) if out.size else None))
        return [MerkleProof(out[i]) for i in range(idx.size)]

    def prove(self, leaf_index: int) -> MerkleProof:
        return self.prove_many([leaf_index])[0]


class ResidentPolynomialBatch:
    """PolynomialBatch kept on the device: only the cap crosses PCIe at commit time."""

    def __init__(self, ctx, handle, cap, ncols, degree_log, rate_bits, blinding, stats):
        self.ctx = ctx
        self.handle = handle
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.blinding = blinding
        self.ncols = ncols
        self.stats = stats
        width = ncols + (SALT_SIZE if blinding else 0)
        self.merkle_tree = ResidentMerkleTree(self, cap, 1 << (degree_log + rate_bits), width)

    def close(self):
        if self.handle:
            self.ctx.lib.vpbs_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shard(self):
        """(first_leaf, nleaves): the rows this batch holds — (0, m) unless its context shards
        (Context.set_shard)."""
        first, nl = ctypes.c_uint64(), ctypes.c_uint64()
        self.ctx.check(self.ctx.lib.vpbs_batch_shard(self.handle, ctypes.byref(first), ctypes.byref(nl)))
        return int(first.value), int(nl.value)

    def get_lde_values(self, index: int, step: int = 1) -> np.ndarray:
        k = reverse_bits(index * step, self.degree_log + self.rate_bits)
        row = self.merkle_tree.get(k)
        return row[: row.shape[0] - (SALT_SIZE if self.blinding else 0)]

    def eval_ext2(self, points) -> np.ndarray:
        """Openings of every polynomial of the batch at extension points: (npoints, ncols, 2)."""
        pts = _as_u64(points).reshape(-1, 2)
        out = np.empty((pts.shape[0], self.ncols, 2), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_batch_eval_ext2(self.handle, _ptr(pts), pts.shape[0], _ptr(out)))
        return out

    def get_lde_rows(self, first: int, step: int, count: int) -> np.ndarray:
        """get_lde_values(first + i * step, 1) for i < count as a (count, ncols) block — what
        [P2] plonk/prover.rs compute_quotient_polys reads (get_lde_values_packed), pulled lazily."""
        out = np.empty((count, self.ncols), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_batch_get_lde_rows(self.handle, first, step, count,
                                                            _ptr(out) if count else None))
        return out

    def coefficients(self) -> np.ndarray:
        """PolynomialBatch.polynomials of the resident batch: (ncols, n) coefficient columns, copied to
        the host (leaves and digests stay on the device)."""
        n = 1 << self.degree_log
        coeffs = np.empty((self.ncols, n), np.uint64)
        cop = (u64p * self.ncols)(*[_ptr(coeffs[c]) for c in range(self.ncols)])
        self.ctx.check(self.ctx.lib.vpbs_batch_download(self.handle, cop, None, None))
        return coeffs

    def download(self) -> "PolynomialBatch":
        """Materialise the eager form (polynomials, leaves, digests) on the host; of a sharded batch:
        all polynomials, the shard's own rows and the digests of its own cap subtrees."""
        n, m_all = 1 << self.degree_log, 1 << (self.degree_log + self.rate_bits)
        cap = self.merkle_tree.cap
        m = self.shard[1]
        coeffs = np.empty((self.ncols, n), np.uint64)
        cop = (u64p * self.ncols)(*[_ptr(coeffs[c]) for c in range(self.ncols)])
        leaves = np.empty((m, self.merkle_tree.width), np.uint64)
        digests = np.empty((2 * (m - cap.shape[0] * m // m_all), 4), np.uint64)
        self.ctx.check(self.ctx.lib.vpbs_batch_download(self.handle, cop, _ptr(leaves),
                                                        _ptr(digests) if digests.size else None))
        return PolynomialBatch(coeffs, MerkleTree(leaves, digests, cap), self.degree_log,
                               self.rate_bits, self.blinding, self.stats)


def open_all_at_points(batches: Sequence["ResidentPolynomialBatch"], points) -> List[np.ndarray]:
    """[P2] plonk/proof.rs OpeningSet::new over all oracles in one round trip
    (vpbs_batches_eval_ext2): per batch the (npoints, ncols, 2) openings at the same points."""
    ctx = batches[0].ctx
    pts = _as_u64(points).reshape(-1, 2)
    outs = [np.empty((pts.shape[0], b.ncols, 2), np.uint64) for b in batches]
    hs = (ctypes.c_void_p * len(batches))(*[b.handle for b in batches])
    ops = (u64p * len(batches))(*[_ptr(o) for o in outs])
    ctx.check(ctx.lib.vpbs_batches_eval_ext2(hs, len(batches), _ptr(pts), pts.shape[0], ops))
    return outs


def open_all_at_leaves(batches: Sequence["ResidentPolynomialBatch"], leaf_indices):
    """[P2] fri/prover.rs fri_prover_query_round (initial_trees_proof) over all oracles in one round
    trip (vpbs_batches_open): per batch (rows (count, width), siblings (count, layers, 4))."""
    ctx = batches[0].ctx
    idx = _as_u64(leaf_indices).reshape(-1)
    t0 = batches[0].merkle_tree
    layers = log2_strict(t0.nleaves) - log2_strict(t0.cap.shape[0])
    rows = [np.empty((idx.size, b.merkle_tree.width), np.uint64) for b in batches]
    sibs = [np.empty((idx.size, layers, 4), np.uint64) for _ in batches]
    hs = (ctypes.c_void_p * len(batches))(*[b.handle for b in batches])
    rp = (u64p * len(batches))(*[_ptr(r) for r in rows])
    sp = (u64p * len(batches))(*[_ptr(x) for x in sibs])
    ctx.check(ctx.lib.vpbs_batches_open(hs, len(batches), _ptr(idx), idx.size, rp, sp))
    return list(zip(rows, sibs))


def commit_resident(cols, rate_bits: int, blinding: bool, cap_height: int,
                    inputs_are_coeffs: bool = False, *, ctx: Optional[Context] = None, salt=None,
                    rng: Optional[np.random.Generator] = None) -> ResidentPolynomialBatch:
    """PolynomialBatch::from_values / from_coeffs with the result left in HBM."""
    ctx = ctx or default_context()
    a = _as_u64(cols)
    if a.ndim != 2 or a.shape[0] == 0:
        raise ValueError("need a non-empty (ncols, n) matrix")
    ncols, n = a.shape
    log_n = log2_strict(n)
    m = n << rate_bits
    if cap_height > log_n + rate_bits:
        raise ValueError("cap_height=%d should be at most log2(leaves.len())=%d"
                         % (cap_height, log_n + rate_bits))
    saltp, salt_arr = None, None
    if blinding:
        if salt is None:
            rng = rng or np.random.default_rng()
            salt = rng.integers(0, P, size=(SALT_SIZE, m), dtype=np.uint64)
        salt_arr = _as_u64(salt)
        if salt_arr.shape != (SALT_SIZE, m):
            raise ValueError("salt must be (%d, %d)" % (SALT_SIZE, m))
        saltp = (u64p * SALT_SIZE)(*[_ptr(salt_arr[s]) for s in range(SALT_SIZE)])
    colp = (u64p * ncols)(*[_ptr(a[c]) for c in range(ncols)])
    cap = np.empty((1 << cap_height, 4), np.uint64)
    handle = ctypes.c_void_p()
    st = VpbsStats()
    ctx.check(ctx.lib.vpbs_batch_commit(ctx.handle, colp, ncols, log_n, rate_bits, cap_height,
                                        int(inputs_are_coeffs), saltp, _ptr(cap),
                                        ctypes.byref(handle), ctypes.byref(st)))
    return ResidentPolynomialBatch(ctx, handle, cap, ncols, log_n, rate_bits, blinding, st.as_dict())


def commit_resident_device(ctx: Context, d_cols: int, ncols: int, log_n: int, rate_bits: int,
                           cap_height: int, inputs_are_coeffs: bool = False) -> ResidentPolynomialBatch:
    """vpbs_batch_commit_dev: a resident batch from columns that are already in HBM."""
    cap = np.empty((1 << cap_height, 4), np.uint64)
    handle = ctypes.c_void_p()
    st = VpbsStats()
    ctx.check(ctx.lib.vpbs_batch_commit_dev(ctx.handle, d_cols, ncols, log_n, rate_bits, cap_height,
                                            int(inputs_are_coeffs), _ptr(cap), ctypes.byref(handle),
                                            ctypes.byref(st)))
    return ResidentPolynomialBatch(ctx, handle, cap, ncols, log_n, rate_bits, False, st.as_dict())


# ----------------------------------------------------------------------------- plonk/prover.rs
def get_unique_coset_shifts(subgroup_size: int, num_shifts: int) -> np.ndarray:
    """[P2] plonky2_field/src/cosets.rs get_unique_coset_shifts: g^0 .. g^(num_shifts-1), g = 7
    (CommonCircuitData.k_is)."""
    log2_strict(subgroup_size)
    out, x = [], 1
    for _ in range(num_shifts):
        out.append(x)
        x = x * COSET_SHIFT % P
    return np.array(out, dtype=np.uint64)


class Sigmas:
    """The sigma polynomials' values and the coset shifts k_is of one circuit, resident in HBM
    (vpbs_sigmas_upload): constant per circuit, uploaded once."""

    def __init__(self, sigma_cols, k_is, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        a, k = _as_u64(sigma_cols), _as_u64(k_is).reshape(-1)
        if a.ndim != 2 or a.shape[0] != k.size or a.shape[0] == 0:
            raise ValueError("sigma_cols must be (num_routed, n) with one k_i per routed wire")
        self.num_routed, n = a.shape
        self.degree_bits = log2_strict(n)
        colp = (u64p * self.num_routed)(*[_ptr(a[j]) for j in range(self.num_routed)])
        self.handle = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.vpbs_sigmas_upload(self.ctx.handle, colp, _ptr(k), self.num_routed,
                                                       self.degree_bits, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.ctx.lib.vpbs_sigmas_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def all_wires_permutation_partial_products(wire_cols, sigmas: Sigmas, betas, gammas,
                                           max_degree: int) -> np.ndarray:
    """[P2] plonk/prover.rs all_wires_permutation_partial_products, returned in the order prove()
    commits the columns: (num_challenges * K, n) = [Z_0 .. Z_{C-1}, partial products of challenge 0,
    of challenge 1, ...].  wire_cols: (num_routed, n) with wire_cols[j][i] = witness.get_wire(i, j)."""
    ctx = sigmas.ctx
    w = _as_u64(wire_cols)
    b, g = _as_u64(betas).reshape(-1), _as_u64(gammas).reshape(-1)
    if w.shape != (sigmas.num_routed, 1 << sigmas.degree_bits) or b.size != g.size or b.size == 0:
        raise ValueError("wire_cols must be (num_routed, n); one gamma per beta")
    if max_degree < 2:
        raise ValueError("max_degree must be at least 2")
    K = -(-sigmas.num_routed // max_degree)
    out = np.empty((b.size * K, w.shape[1]), np.uint64)
    wp = (u64p * w.shape[0])(*[_ptr(w[j]) for j in range(w.shape[0])])
    op = (u64p * out.shape[0])(*[_ptr(out[c]) for c in range(out.shape[0])])
    ctx.check(ctx.lib.vpbs_zs_partial_products(ctx.handle, wp, sigmas.handle, max_degree, _ptr(b),
                                               _ptr(g), b.size, op))
    return out


def commit_zs_partial_products(wires: ResidentPolynomialBatch, sigmas: Sigmas, betas, gammas,
                               max_degree: int, rate_bits: int, cap_height: int) -> ResidentPolynomialBatch:
    """prove() steps 4-5 on the device: Z and partial products from the resident wires batch,
    committed at once as a new resident batch (vpbs_batch_zs_partial_products)."""
    ctx = wires.ctx
    b, g = _as_u64(betas).reshape(-1), _as_u64(gammas).reshape(-1)
    if b.size != g.size or b.size == 0:
        raise ValueError("one gamma per beta")
    K = -(-sigmas.num_routed // max_degree)
    cap = np.empty((1 << cap_height, 4), np.uint64)
    handle = ctypes.c_void_p()
    st = VpbsStats()
    ctx.check(ctx.lib.vpbs_batch_zs_partial_products(wires.handle, sigmas.handle, max_degree, _ptr(b),
                                                     _ptr(g), b.size, rate_bits, cap_height, _ptr(cap),
                                                     ctypes.byref(handle), ctypes.byref(st)))
    return ResidentPolynomialBatch(ctx, handle, cap, b.size * K, wires.degree_log, rate_bits, False,
                                   st.as_dict())


class GateProgram:
    """The gate constraints of a circuit as a straight-line program on the device
    (vpbs_gate_program_upload; instruction format in include/vpbs_commit.h).  Built with
    GateProgramBuilder; constant per circuit, uploaded once."""

    def __init__(self, code, imms, nregs: int, num_constraints: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        c, im = _as_u64(code).reshape(-1), _as_u64(imms).reshape(-1)
        self.code, self.imms, self.nregs, self.num_constraints = c, im, nregs, num_constraints
        self.handle = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.vpbs_gate_program_upload(
            self.ctx.handle, _ptr(c) if c.size else None, c.size, _ptr(im) if im.size else None, im.size,
            nregs, num_constraints, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.ctx.lib.vpbs_gate_program_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GateProgramBuilder:
    """Assembler for gate programs: what a host-side compiler of plonky2's gates emits.  Operands are
    tuples (kind, index): reg(i), wire(col), const(col), imm(value), pih(i); add / sub / mul return a
    fresh register; emit(j, v) states constraint j of the current gate; end_gate(filter) closes it.
    Registers are recycled at end_gate (a gate's temporaries die with it)."""
    ADD, SUB, MUL, EMIT, ENDGATE, MAD = 0, 1, 2, 3, 4, 5

    def __init__(self):
        self.code: List[int] = []
        self.imms: List[int] = []
        self._imm_index = {}
        self._next = 0
        self.nregs = 1
        self.num_constraints = 1

    @staticmethod
    def wire(col): return (1, col)

    @staticmethod
    def const(col): return (2, col)

    @staticmethod
    def pih(i): return (4, i)

    def imm(self, value):
        v = int(value) % P
        if v not in self._imm_index:
            self._imm_index[v] = len(self.imms)
            self.imms.append(v)
        return (3, self._imm_index[v])

    def _ins(self, op, dst, a, b=(0, 0)):
        self.code.append(op | dst << 8 | a[0] << 16 | b[0] << 20 | a[1] << 24 | b[1] << 40)

    def _binary(self, op, a, b):
        dst = self._next
        self._next += 1
        if self._next > 224:
            raise ValueError("gate needs more than 224 live registers")
        self.nregs = max(self.nregs, self._next)
        self._ins(op, dst, a, b)
        return (0, dst)

    def into(self, dst: int, op: int, a, b):
        """dst <- a op b into a register the caller manages (long gates reuse registers; add / sub / mul
        hand out a fresh one per result and recycle only at end_gate)."""
        if not 0 <= dst < 224:
            raise ValueError("register index out of range")
        self.nregs = max(self.nregs, dst + 1)
        self._next = max(self._next, dst + 1)
        self._ins(op, dst, a, b)
        return (0, dst)

    def mad(self, acc, a, b):
        """acc <- acc + a b in place (acc: a register operand); returns acc."""
        if acc[0] != 0:
            raise ValueError("mad accumulates into a register")
        self._ins(self.MAD, acc[1], a, b)
        return acc

    def add(self, a, b): return self._binary(self.ADD, a, b)
    def sub(self, a, b): return self._binary(self.SUB, a, b)
    def mul(self, a, b): return self._binary(self.MUL, a, b)

    def emit(self, j: int, value):
        self.num_constraints = max(self.num_constraints, j + 1)
        self._ins(self.EMIT, 0, value, (0, j))

    def end_gate(self, filter_value):
        self._ins(self.ENDGATE, 0, filter_value)
        self._next = 0

    def selector_filter(self, selector_col: int, row: int, group, many_selectors: bool):
        """[P2] gates/gate.rs compute_filter: prod over the group's other gate indices i of (i - s),
        times (UNUSED_SELECTOR - s) when the circuit has several selector polynomials; s = the
        selector constant of this row."""
        s = self.const(selector_col)
        f = self.imm(1)
        for i in group:
            if i != row:
                f = self.mul(f, self.sub(self.imm(i), s))
        if many_selectors:
            f = self.mul(f, self.sub(self.imm(2**32 - 1), s))
        return f

    def build(self, ctx: Optional[Context] = None, num_constraints: Optional[int] = None) -> GateProgram:
        return GateProgram(np.array(self.code, dtype=np.uint64), np.array(self.imms, dtype=np.uint64),
                           self.nregs, num_constraints or self.num_constraints, ctx)


def commit_quotient_polys(constants_sigmas: "ResidentPolynomialBatch", sigmas_first_col: int,
                          wires: "ResidentPolynomialBatch", zs_pp: "ResidentPolynomialBatch", k_is,
                          max_degree: int, quotient_degree_bits: int, betas, gammas, alphas,
                          rate_bits: int, cap_height: int, gate_terms=None, program: "GateProgram" = None,
                          public_inputs_hash=None) -> "ResidentPolynomialBatch":
    """prove() steps 6-7 on the device, gate-independent part ([P2] plonk/prover.rs
    compute_quotient_polys / plonk/vanishing_poly.rs): the Z(1) = 1 terms and the partial-product
    checks over the quotient domain, reduced with the alphas, divided by Z_H, coset_ifft, chunked and
    committed (vpbs_batch_quotient_polys).  The gate constraints come either as gate_terms —
    (num_challenges, n << quotient_degree_bits) alpha-reduced values in natural order — or as a
    GateProgram the device evaluates at every point (with public_inputs_hash), or not at all."""
    ctx = wires.ctx
    k = _as_u64(k_is).reshape(-1)
    b, g, a = (_as_u64(v).reshape(-1) for v in (betas, gammas, alphas))
    if not (b.size == g.size == a.size) or b.size == 0:
        raise ValueError("one beta, gamma and alpha per challenge")
    gt_arr, gtp = None, None
    if gate_terms is not None:
        gt_arr = _as_u64(gate_terms)
        if gt_arr.shape != (b.size, (1 << wires.degree_log) << quotient_degree_bits):
            raise ValueError("gate_terms must be (num_challenges, n << quotient_degree_bits)")
        gtp = (u64p * b.size)(*[_ptr(gt_arr[c]) for c in range(b.size)])
    cap = np.empty((1 << cap_height, 4), np.uint64)
    handle = ctypes.c_void_p()
    st = VpbsStats()
    ctx.check(ctx.lib.vpbs_batch_quotient_polys(constants_sigmas.handle, sigmas_first_col, wires.handle,
                                                zs_pp.handle, _ptr(k), k.size, max_degree,
                                                quotient_degree_bits, _ptr(b), _ptr(g), _ptr(a), b.size, gtp,
                                                program.handle if program is not None else None,
                                                _ptr(_as_u64(public_inputs_hash).reshape(4))
                                                if public_inputs_hash is not None else None,
                                                rate_bits, cap_height, _ptr(cap), ctypes.byref(handle),
                                                ctypes.byref(st)))
    return ResidentPolynomialBatch(ctx, handle, cap, b.size << quotient_degree_bits, wires.degree_log,
                                   rate_bits, False, st.as_dict())


# ----------------------------------------------------------------------------- device-resident
def commit_device(ctx: Context, d_cols: int, ncols: int, log_n: int, rate_bits: int,
                  cap_height: int, inputs_are_coeffs: bool, d_coeffs: int, d_leaves: int,
                  d_digests: int, d_cap: int, d_salt: int = 0, want_stats: bool = False):
    """vpbs_commit_dev on raw device pointers (e.g. torch tensor .data_ptr())."""
    st = VpbsStats() if want_stats else None
    ctx.check(ctx.lib.vpbs_commit_dev(ctx.handle, d_cols, ncols, log_n, rate_bits, cap_height,
                                      int(inputs_are_coeffs), d_salt or None, d_coeffs or None,
                                      d_leaves, d_digests or None, d_cap,
                                      ctypes.byref(st) if st is not None else None))
    return st.as_dict() if st is not None else None


def commit_shard_device(ctx: Context, d_cols: int, ncols: int, log_n: int, rate_bits: int,
                        cap_height: int, inputs_are_coeffs: bool, first_leaf: int,
                        nleaves_shard: int, d_coeffs: int, d_leaves: int, d_digests: int,
                        d_roots: int, want_stats: bool = False):
    """vpbs_commit_shard_dev: the row range [first_leaf, first_leaf + nleaves_shard) of a commit."""
    st = VpbsStats() if want_stats else None
    ctx.check(ctx.lib.vpbs_commit_shard_dev(ctx.handle, d_cols, ncols, log_n, rate_bits, cap_height,
                                            int(inputs_are_coeffs), first_leaf, nleaves_shard,
                                            d_coeffs or None, d_leaves, d_digests or None, d_roots,
                                            ctypes.byref(st) if st is not None else None))
    return st.as_dict() if st is not None else None


# ----------------------------------------------------------------------------- synthetic inputs
def synthetic_columns(ncols: int, n: int, seed: int = 0x5EED0000, canonical: bool = True) -> np.ndarray:
    """SURVEY.md §8(d) workload: column c, row i = splitmix64(seed + c, counter i) (mod p)."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    out = np.empty((ncols, n), np.uint64)
    with np.errstate(over="ignore"):
        for c in range(ncols):
            z = np.uint64((seed + c) % 2**64) + i * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[c] = z
    if canonical:
        out = np.where(out >= np.uint64(P), out - np.uint64(P), out)
    return out
