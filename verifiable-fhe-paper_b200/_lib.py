"""ctypes binding of libvpbs_commit.so (include/vpbs_commit.h).

The product path has no CPU fallback: if the shared library is missing or no CUDA device is
usable, importing is fine (so that CPU-only tests can check symbols) but every compute call
raises VpbsError.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvpbs_commit.so")

u64p = ctypes.POINTER(ctypes.c_uint64)
u64pp = ctypes.POINTER(u64p)


class VpbsStats(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("ifft_ms", ctypes.c_float),
                ("fft_ms", ctypes.c_float), ("merkle_ms", ctypes.c_float),
                ("leaf_hash_ms", ctypes.c_float),
                ("d2h_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("kernel_launches", ctypes.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VpbsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vpbs error %d: %s" % (code, msg))
        self.code = code


VPBS_OK, VPBS_ERR_ARG, VPBS_ERR_CUDA, VPBS_ERR_OOM, VPBS_ERR_STATE = 0, -1, -2, -3, -4

# name -> (restype, argtypes); must list every symbol include/vpbs_commit.h declares
_c = ctypes
_ctx = _c.c_void_p
SIGNATURES = {
    "vpbs_abi_version": (_c.c_int, []),
    "vpbs_device_count": (_c.c_int, []),
    "vpbs_ctx_create": (_c.c_int, [_c.c_int, _c.POINTER(_ctx)]),
    "vpbs_ctx_destroy": (None, [_ctx]),
    "vpbs_ctx_set_stream": (_c.c_int, [_ctx, _c.c_void_p]),
    "vpbs_ctx_sync": (_c.c_int, [_ctx]),
    "vpbs_last_error": (_c.c_char_p, [_ctx]),
    "vpbs_ctx_kernel_launches": (_c.c_uint64, [_ctx]),
    "vpbs_ctx_set_host_threads": (_c.c_int, [_ctx, _c.c_uint]),
    "vpbs_ctx_set_shard": (_c.c_int, [_ctx, _c.c_uint32, _c.c_uint32]),
    "vpbs_host_alloc": (_c.c_void_p, [_c.c_size_t]),
    "vpbs_host_free": (None, [_c.c_void_p]),
    "vpbs_fft": (_c.c_int, [_ctx, u64p, _c.c_uint32]),
    "vpbs_ifft": (_c.c_int, [_ctx, u64p, _c.c_uint32]),
    "vpbs_coset_fft": (_c.c_int, [_ctx, u64p, _c.c_uint32, _c.c_uint64]),
    "vpbs_poseidon_permute": (_c.c_int, [_ctx, u64p, _c.c_uint64]),
    "vpbs_hash_or_noop_batch": (_c.c_int, [_ctx, u64p, _c.c_uint64, _c.c_uint32, u64p]),
    "vpbs_two_to_one_batch": (_c.c_int, [_ctx, u64p, u64p, _c.c_uint64, u64p]),
    "vpbs_merkle_new": (_c.c_int, [_ctx, u64p, _c.c_uint64, _c.c_uint32, _c.c_uint32, u64p, u64p]),
    "vpbs_merkle_new_dev": (_c.c_int, [_ctx, _c.c_void_p, _c.c_uint64, _c.c_uint32, _c.c_uint32,
                                       _c.c_void_p, _c.c_void_p, _c.POINTER(VpbsStats)]),
    "vpbs_lde_batch": (_c.c_int, [_ctx, u64pp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_int,
                                  u64pp, u64p]),
    "vpbs_commit": (_c.c_int, [_ctx, u64pp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                               _c.c_int, u64pp, u64pp, u64p, u64p, u64p, _c.POINTER(VpbsStats)]),
    "vpbs_commit_multi": (_c.c_int, [_c.POINTER(_ctx), _c.c_int, u64pp, _c.c_uint32, _c.c_uint32,
                                     _c.c_uint32, _c.c_uint32, _c.c_int, u64pp, u64pp, u64p, u64p,
                                     u64p, _c.POINTER(VpbsStats)]),
    "vpbs_commit_dev": (_c.c_int, [_ctx, _c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                                   _c.c_uint32, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                   _c.c_void_p, _c.c_void_p, _c.POINTER(VpbsStats)]),
    "vpbs_commit_shard_dev": (_c.c_int, [_ctx, _c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                                         _c.c_uint32, _c.c_int, _c.c_uint64, _c.c_uint64,
                                         _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                         _c.POINTER(VpbsStats)]),
    "vpbs_eval_ext2": (_c.c_int, [_ctx, u64pp, _c.c_uint32, _c.c_uint32, u64p, _c.c_uint32, u64p]),
    "vpbs_batch_eval_ext2": (_c.c_int, [_c.c_void_p, u64p, _c.c_uint32, u64p]),
    "vpbs_fri_layer_commit": (_c.c_int, [_ctx, u64p, _c.c_uint64, _c.c_uint32, _c.c_uint32, u64p, u64p, u64p]),
    "vpbs_fri_fold": (_c.c_int, [_ctx, u64p, _c.c_uint64, _c.c_uint32, u64p, _c.c_uint64, u64p, u64p]),
    "vpbs_fri_begin": (_c.c_int, [_ctx, u64p, _c.c_uint64, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "vpbs_fri_begin_openings": (_c.c_int, [_ctx, _c.POINTER(_c.c_void_p), _c.c_uint32,
                                           _c.POINTER(_c.c_uint32), _c.c_uint32, _c.POINTER(_c.c_uint32),
                                           u64p, u64p, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "vpbs_fri_commit_layer": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, u64p]),
    "vpbs_fri_fold_layer": (_c.c_int, [_c.c_void_p, u64p]),
    "vpbs_fri_final_poly": (_c.c_int, [_c.c_void_p, _c.c_uint32, u64p]),
    "vpbs_fri_query_layer": (_c.c_int, [_c.c_void_p, _c.c_uint32, u64p, _c.c_uint64, u64p, u64p]),
    "vpbs_fri_destroy": (None, [_c.c_void_p]),
    "vpbs_pow_grind": (_c.c_int, [_ctx, u64p, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_uint64,
                                  _c.c_uint64, u64p, _c.POINTER(_c.c_int)]),
    "vpbs_batch_commit": (_c.c_int, [_ctx, u64pp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                                     _c.c_int, u64pp, u64p, _c.POINTER(_c.c_void_p),
                                     _c.POINTER(VpbsStats)]),
    "vpbs_batch_destroy": (None, [_c.c_void_p]),
    "vpbs_batch_get_leaves": (_c.c_int, [_c.c_void_p, u64p, _c.c_uint64, u64p]),
    "vpbs_batch_prove": (_c.c_int, [_c.c_void_p, u64p, _c.c_uint64, u64p]),
    "vpbs_batch_download": (_c.c_int, [_c.c_void_p, u64pp, u64p, u64p]),
    "vpbs_batch_commit_dev": (_c.c_int, [_ctx, _c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                                         _c.c_uint32, _c.c_int, u64p, _c.POINTER(_c.c_void_p),
                                         _c.POINTER(VpbsStats)]),
    "vpbs_batch_get_lde_rows": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_uint64, _c.c_uint64, u64p]),
    "vpbs_sigmas_upload": (_c.c_int, [_ctx, u64pp, u64p, _c.c_uint32, _c.c_uint32,
                                      _c.POINTER(_c.c_void_p)]),
    "vpbs_sigmas_destroy": (None, [_c.c_void_p]),
    "vpbs_zs_partial_products": (_c.c_int, [_ctx, u64pp, _c.c_void_p, _c.c_uint32, u64p, u64p,
                                            _c.c_uint32, u64pp]),
    "vpbs_batch_zs_partial_products": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint32, u64p, u64p,
                                                  _c.c_uint32, _c.c_uint32, _c.c_uint32, u64p,
                                                  _c.POINTER(_c.c_void_p), _c.POINTER(VpbsStats)]),
    "vpbs_batch_shape": (_c.c_int, [_c.c_void_p] + [_c.POINTER(_c.c_uint32)] * 5),
    "vpbs_batch_quotient_polys": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p, u64p,
                                             _c.c_uint32, _c.c_uint32, _c.c_uint32, u64p, u64p, u64p,
                                             _c.c_uint32, _c.POINTER(u64p), _c.c_void_p, u64p, _c.c_uint32,
                                             _c.c_uint32, u64p, _c.POINTER(_c.c_void_p),
                                             _c.POINTER(VpbsStats)]),
    "vpbs_batch_quotient_values": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p, u64p,
                                              _c.c_uint32, _c.c_uint32, _c.c_uint32, u64p, u64p, u64p,
                                              _c.c_uint32, _c.POINTER(u64p), _c.c_void_p, u64p, _c.c_void_p]),
    "vpbs_quotient_commit_values": (_c.c_int, [_ctx, _c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_uint32,
                                               _c.c_uint32, _c.c_uint32, u64p, _c.POINTER(_c.c_void_p),
                                               _c.POINTER(VpbsStats)]),
    "vpbs_gate_program_upload": (_c.c_int, [_ctx, u64p, _c.c_uint32, u64p, _c.c_uint32, _c.c_uint32,
                                            _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "vpbs_gate_program_destroy": (None, [_c.c_void_p]),
    "vpbs_batch_shard": (_c.c_int, [_c.c_void_p] + [_c.POINTER(_c.c_uint64)] * 2),
    "vpbs_batches_eval_ext2": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_uint32, u64p, _c.c_uint32,
                                          _c.POINTER(u64p)]),
    "vpbs_batches_open": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_uint32, u64p, _c.c_uint64,
                                     _c.POINTER(u64p), _c.POINTER(u64p)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VpbsError(VPBS_ERR_STATE,
                            "%s not built (run __graft_entry__.build()); there is no CPU fallback"
                            % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
