"""ctypes binding for oracle/liboracle.so (CPU restatement of plonky2 0.2.0's commit path).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
P = 0xFFFFFFFF00000001

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u64pp = ctypes.POINTER(_u64p)


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so with the recipe in oracle/Makefile."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "poseidon_simd.inc", "Makefile")]
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u64, u32, sz = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_size_t
        for name in ("orc_gl_add", "orc_gl_sub", "orc_gl_mul", "orc_gl_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64]
        L.orc_gl_inv.restype = u64
        L.orc_gl_inv.argtypes = [u64]
        L.orc_primitive_root_of_unity.restype = u64
        L.orc_primitive_root_of_unity.argtypes = [ctypes.c_uint]
        for name in ("orc_fft", "orc_ifft"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [_u64p, ctypes.c_uint]
        L.orc_coset_fft.restype = None
        L.orc_coset_fft.argtypes = [_u64p, ctypes.c_uint, u64]
        L.orc_lde.restype = None
        L.orc_lde.argtypes = [_u64p, ctypes.c_uint, ctypes.c_uint, _u64p]
        L.orc_poseidon_round_constants.argtypes = [_u64p]
        L.orc_poseidon.argtypes = [_u64p]
        L.orc_hash_no_pad.argtypes = [_u64p, sz, _u64p]
        L.orc_hash_or_noop.argtypes = [_u64p, sz, _u64p]
        L.orc_two_to_one.argtypes = [_u64p, _u64p, _u64p]
        L.orc_merkle_new.restype = ctypes.c_int
        L.orc_merkle_new.argtypes = [_u64p, u64, u32, u32, _u64p, _u64p]
        L.orc_merkle_prove.restype = ctypes.c_int
        L.orc_merkle_prove.argtypes = [_u64p, u64, u32, u64, _u64p]
        L.orc_merkle_verify.restype = ctypes.c_int
        L.orc_merkle_verify.argtypes = [_u64p, u32, u64, _u64p, u32, _u64p, u32]
        L.orc_commit.restype = ctypes.c_int
        L.orc_commit.argtypes = [_u64pp, u32, u32, u32, u32, ctypes.c_int, _u64pp,
                                 _u64p, _u64p, _u64p, _u64p, _u64p]
        L.orc_eval_ext2.restype = None
        L.orc_eval_ext2.argtypes = [_u64pp, u32, u64, _u64p, _u64p]
        L.orc_fri_layer_commit.restype = ctypes.c_int
        L.orc_fri_layer_commit.argtypes = [_u64p, u64, u32, u32, _u64p, _u64p, _u64p]
        L.orc_fri_fold.restype = None
        L.orc_fri_fold.argtypes = [_u64p, u64, u32, _u64p, u64, _u64p, _u64p]
        L.orc_fri_final_poly.restype = ctypes.c_int
        L.orc_fri_final_poly.argtypes = [_u64pp, ctypes.POINTER(u32), u32, u64, _u64p, _u64p, _u64p]
        L.orc_zs_partial_products.restype = ctypes.c_int
        L.orc_zs_partial_products.argtypes = [_u64pp, _u64pp, _u64p, u32, u32, u32, u64, u64, _u64p]
        L.orc_quotient_polys.restype = ctypes.c_int
        L.orc_quotient_polys.argtypes = [_u64pp, _u64pp, _u64pp, _u64p, u32, u32, u32, u32, _u64p, _u64p,
                                         _u64p, u32, _u64pp, _u64p]
        L.orc_gate_program_eval.restype = ctypes.c_int
        L.orc_gate_program_eval.argtypes = [_u64p, u32, _u64p, u32, u32, u32, _u64pp, u32, _u64pp, u32, u32, u32,
                                            _u64p, _u64p, u32, _u64pp]
        L.orc_set_simd.argtypes = [ctypes.c_int]
        L.orc_get_simd.restype = ctypes.c_int
        L.orc_poseidon_batch.argtypes = [_u64p, u64]
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_get_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _arr(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def gl_mul(a, b): return lib().orc_gl_mul(int(a), int(b))
def gl_add(a, b): return lib().orc_gl_add(int(a), int(b))
def gl_sub(a, b): return lib().orc_gl_sub(int(a), int(b))
def gl_pow(a, e): return lib().orc_gl_pow(int(a), int(e))
def gl_inv(a): return lib().orc_gl_inv(int(a))
def primitive_root_of_unity(k): return lib().orc_primitive_root_of_unity(int(k))
def set_threads(n): lib().orc_set_threads(int(n))
def get_threads(): return lib().orc_get_threads()


def _log2(n):
    assert n > 0 and n & (n - 1) == 0, "length must be a power of two"
    return n.bit_length() - 1


def fft(v):
    a = _arr(v).copy(); lib().orc_fft(_p(a), _log2(a.size)); return a


def ifft(v):
    a = _arr(v).copy(); lib().orc_ifft(_p(a), _log2(a.size)); return a


def coset_fft(v, shift):
    a = _arr(v).copy(); lib().orc_coset_fft(_p(a), _log2(a.size), int(shift)); return a


def lde(coeffs, rate_bits):
    a = _arr(coeffs); out = np.empty(a.size << rate_bits, dtype=np.uint64)
    lib().orc_lde(_p(a), _log2(a.size), rate_bits, _p(out)); return out


def round_constants():
    out = np.empty(360, dtype=np.uint64); lib().orc_poseidon_round_constants(_p(out)); return out


def poseidon(state):
    a = _arr(state).copy(); assert a.size == 12; lib().orc_poseidon(_p(a)); return a


def hash_no_pad(x):
    a = _arr(x); out = np.empty(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_p(a) if a.size else None, a.size, _p(out)); return out


def hash_or_noop(x):
    a = _arr(x); out = np.empty(4, dtype=np.uint64)
    lib().orc_hash_or_noop(_p(a) if a.size else None, a.size, _p(out)); return out


def two_to_one(l, r):
    out = np.empty(4, dtype=np.uint64)
    lib().orc_two_to_one(_p(_arr(l)), _p(_arr(r)), _p(out)); return out


def merkle_new(leaves, cap_height):
    """leaves: (nleaves, leaf_len) uint64 -> (digests (2(n-2^h),4), cap (2^h,4))."""
    a = _arr(leaves); n, w = a.shape
    digests = np.empty((2 * (n - (1 << cap_height)) if n >= (1 << cap_height) else 0, 4), np.uint64)
    cap = np.empty((1 << cap_height, 4), np.uint64)
    rc = lib().orc_merkle_new(_p(a) if a.size else None, n, w, cap_height,
                              _p(digests) if digests.size else None, _p(cap))
    if rc != 0:
        raise ValueError("orc_merkle_new: bad arguments")
    return digests, cap


def merkle_prove(digests, nleaves, cap_height, leaf_index):
    d = _arr(digests)
    sib = np.empty((_log2(nleaves) - cap_height, 4), np.uint64)
    rc = lib().orc_merkle_prove(_p(d) if d.size else None, nleaves, cap_height, leaf_index,
                                _p(sib) if sib.size else None)
    if rc != 0:
        raise ValueError("orc_merkle_prove: bad arguments")
    return sib


def merkle_verify(leaf, leaf_index, siblings, cap):
    l, s, c = _arr(leaf), _arr(siblings), _arr(cap)
    return lib().orc_merkle_verify(_p(l) if l.size else None, l.size, leaf_index,
                                   _p(s) if s.size else None, s.shape[0] if s.size else 0,
                                   _p(c), _log2(c.shape[0])) == 0


def commit(cols, rate_bits, cap_height, inputs_are_coeffs=False, salt_cols=None, want_lde=False):
    """cols: (ncols, n) uint64.  Returns dict(coeffs, lde|None, leaves, digests, cap)."""
    a = _arr(cols); ncols, n = a.shape; log_n = _log2(n); m = n << rate_bits
    width = ncols + (4 if salt_cols is not None else 0)
    colp = (_u64p * ncols)(*[_p(a[c]) for c in range(ncols)])
    saltp = None
    if salt_cols is not None:
        s = _arr(salt_cols); assert s.shape == (4, m)
        saltp = (_u64p * 4)(*[_p(s[i]) for i in range(4)])
    coeffs = np.empty((ncols, n), np.uint64)
    ldec = np.empty((ncols, m), np.uint64) if want_lde else None
    leaves = np.empty((m, width), np.uint64)
    ncap = 1 << cap_height
    digests = np.empty((2 * (m - ncap) if m >= ncap else 0, 4), np.uint64)
    cap = np.empty((ncap, 4), np.uint64)
    rc = lib().orc_commit(colp, ncols, log_n, rate_bits, cap_height, int(inputs_are_coeffs), saltp,
                          _p(coeffs), _p(ldec) if want_lde else None, _p(leaves),
                          _p(digests) if digests.size else None, _p(cap))
    if rc != 0:
        raise ValueError("orc_commit: bad arguments (rc=%d)" % rc)
    return dict(coeffs=coeffs, lde=ldec, leaves=leaves, digests=digests, cap=cap)


def eval_ext2(cols, x):
    """cols: (ncols, n) coefficients; x = (x0, x1) in F[X]/(X^2 - 7) -> (ncols, 2)."""
    a = _arr(cols); ncols, n = a.shape
    colp = (_u64p * ncols)(*[_p(a[c]) for c in range(ncols)])
    xx = _arr(x); out = np.empty((ncols, 2), np.uint64)
    lib().orc_eval_ext2(colp, ncols, n, _p(xx), _p(out))
    return out


def fri_layer_commit(values_ext, arity_bits, cap_height):
    """values_ext: (len, 2) -> dict(leaves (len/arity, 2*arity), digests, cap)."""
    v = _arr(values_ext); ln = v.shape[0]; nl = ln >> arity_bits
    leaves = np.empty((nl, 2 << arity_bits), np.uint64)
    ncap = 1 << cap_height
    digests = np.empty((2 * (nl - ncap) if nl >= ncap else 0, 4), np.uint64)
    cap = np.empty((ncap, 4), np.uint64)
    rc = lib().orc_fri_layer_commit(_p(v), ln, arity_bits, cap_height, _p(leaves),
                                    _p(digests) if digests.size else None, _p(cap))
    if rc != 0:
        raise ValueError("orc_fri_layer_commit: bad arguments")
    return dict(leaves=leaves, digests=digests, cap=cap)


def fri_fold(coeffs_ext, arity_bits, beta, shift_next):
    """coeffs_ext: (len, 2) -> (coeffs' (len/arity, 2), values' (len/arity, 2))."""
    c = _arr(coeffs_ext); ln = c.shape[0]; ol = ln >> arity_bits
    co = np.empty((ol, 2), np.uint64); vo = np.empty((ol, 2), np.uint64)
    lib().orc_fri_fold(_p(c), ln, arity_bits, _p(_arr(beta)), int(shift_next), _p(co), _p(vo))
    return co, vo


def fri_final_poly(batches, points, alpha):
    """[P2] prove_openings up to final_poly.  batches: list of (len_b, n) coefficient arrays;
    points: (nbatches, 2); alpha: (2,) -> (n, 2)."""
    arrs = [_arr(b) for b in batches]
    n = arrs[0].shape[1]
    sizes = np.array([a.shape[0] for a in arrs], np.uint32)
    ptrs = [_p(a[j]) for a in arrs for j in range(a.shape[0])]
    pp = (_u64p * len(ptrs))(*ptrs)
    out = np.empty((n, 2), np.uint64)
    rc = lib().orc_fri_final_poly(pp, sizes.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(arrs), n,
                                  _p(_arr(points).reshape(-1)), _p(_arr(alpha).reshape(2)), _p(out))
    if rc != 0:
        raise ValueError("orc_fri_final_poly: bad arguments")
    return out


def zs_partial_products(wires, sigmas, k_is, max_degree, beta, gamma):
    """[P2] plonk/prover.rs wires_permutation_partial_products_and_zs for one (beta, gamma).
    wires, sigmas: (num_routed, n); returns (K, n) columns [Z, pp_0 .. pp_{K-2}]."""
    w, s, k = _arr(wires), _arr(sigmas), _arr(k_is)
    nr, n = w.shape
    assert s.shape == w.shape and k.size == nr
    K = -(-nr // max_degree)
    out = np.zeros((K, n), np.uint64)
    wp = (_u64p * nr)(*[_p(w[j]) for j in range(nr)])
    sp = (_u64p * nr)(*[_p(s[j]) for j in range(nr)])
    rc = lib().orc_zs_partial_products(wp, sp, _p(k), nr, _log2(n), max_degree, int(beta), int(gamma), _p(out))
    if rc == -2:
        raise ZeroDivisionError("a permutation denominator is zero (plonky2 panics here)")
    if rc != 0:
        raise ValueError("bad arguments")
    return out


def quotient_polys(wire_coeffs, sigma_coeffs, zs_pp_coeffs, k_is, max_degree, qdb, betas, gammas, alphas,
                   gate_terms=None):
    """[P2] compute_quotient_polys (gate-independent terms + optional alpha-reduced gate terms).
    Coefficient columns in: wires (num_routed, n), sigmas (num_routed, n), zs_pp (nc * K, n) in commit
    order; gate_terms: (nc, n << qdb) or None.  Returns (nc * 2^qdb, n) quotient chunks."""
    w, s, z, k = _arr(wire_coeffs), _arr(sigma_coeffs), _arr(zs_pp_coeffs), _arr(k_is)
    b, g, a = _arr(betas).reshape(-1), _arr(gammas).reshape(-1), _arr(alphas).reshape(-1)
    nr, n = w.shape
    nc = b.size
    out = np.zeros((nc << qdb, n), np.uint64)
    ptrs = lambda m: (_u64p * m.shape[0])(*[_p(m[j]) for j in range(m.shape[0])])
    gt = None
    if gate_terms is not None:
        gt_arr = _arr(gate_terms)
        gt = ptrs(gt_arr)
    rc = lib().orc_quotient_polys(ptrs(w), ptrs(s), ptrs(z), _p(k), nr, _log2(n), max_degree, qdb,
                                  _p(b), _p(g), _p(a), nc, gt, _p(out))
    if rc != 0:
        raise ValueError("bad arguments")
    return out


def gate_program_eval(code, imms, nregs, num_constraints, wire_coeffs, cs_coeffs, qdb, pih, alphas):
    """[P2] evaluate_gate_constraints_base_batch for gates given as a program: (nc, n << qdb)
    alpha-reduced gate constraints over the quotient domain (natural order)."""
    c, im = _arr(code).reshape(-1), _arr(imms).reshape(-1)
    w, s, a = _arr(wire_coeffs), _arr(cs_coeffs), _arr(alphas).reshape(-1)
    ph = _arr(pih if pih is not None else [0, 0, 0, 0]).reshape(4)
    n = w.shape[1]
    out = np.zeros((a.size, n << qdb), np.uint64)
    ptrs = lambda m: (_u64p * m.shape[0])(*[_p(m[j]) for j in range(m.shape[0])])
    rc = lib().orc_gate_program_eval(_p(c) if c.size else None, c.size, _p(im) if im.size else None, im.size,
                                     nregs, num_constraints, ptrs(w), w.shape[0], ptrs(s), s.shape[0],
                                     _log2(n), qdb, _p(ph), _p(a), a.size, ptrs(out))
    if rc != 0:
        raise ValueError("bad program")
    return out


def set_simd(width): lib().orc_set_simd(int(width))
def get_simd(): return lib().orc_get_simd()


def poseidon_batch(states):
    """count x 12 states through the SIMD path (4 / 8 permutations per call)."""
    a = _arr(states).reshape(-1, 12).copy()
    lib().orc_poseidon_batch(_p(a), a.shape[0])
    return a
