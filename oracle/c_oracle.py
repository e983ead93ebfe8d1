"""ctypes binding of oracle/libtzk_oracle.so (the C/OpenMP restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
from ctypes import c_float, c_int32, c_int64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtzk_oracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB)
        P = c_void_p
        _lib.orc_num_threads.restype = c_int32
        _lib.orc_pooled_lookup.argtypes = [P, P, P, P, P, P, P, P, c_int32, c_int32, P, c_int64]
        _lib.orc_fused_update.argtypes = [c_int32, P, c_int64, P, P, P, P, P, P, P, P, c_int32, c_int32, P, P,
                                          c_float, c_float, c_float, c_int32]
        _lib.orc_dot_interact_fwd.argtypes = [P, c_int64, P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, P,
                                              c_int64]
        _lib.orc_dot_interact_bwd.argtypes = [P, c_int64, P, c_int64, P, c_int64, c_int64, c_int32, c_int32, c_int32,
                                              c_int32, P, c_int64, P, c_int64]
        _lib.orc_fm_fwd.argtypes = [P, c_int64, c_int64, c_int32, c_int32, P, c_int64]
        _lib.orc_fm_bwd.argtypes = [P, c_int64, P, c_int64, c_int64, c_int32, c_int32, P, c_int64]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _i64(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def pooled_lookup(weights, lay, ids, offsets, B):
    out = np.empty((B, lay.total_dim), dtype=np.float32)
    a = [_i64(lay.w_off), _i64(lay.rows), _i32(lay.dim), _i32(lay.col), _i32(lay.pool)]
    lib().orc_pooled_lookup(_p(weights), *[_p(x) for x in a], _p(ids), _p(offsets), lay.num_features, B, _p(out),
                            lay.total_dim)
    return out


def fused_update(optimizer, grad_out, weights, state, lay, ids, offsets, B, lr, eps, grad_scale):
    a = [_i64(lay.w_off), _i64(lay.rows), _i32(lay.dim), _i32(lay.col), _i32(lay.pool), _i64(lay.key_base)]
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    lib().orc_fused_update(optimizer, _p(grad_out), grad_out.shape[1], *[_p(x) for x in a], _p(ids), _p(offsets),
                           lay.num_features, B, _p(weights), _p(state), lr, eps, grad_scale, lay.max_dim)


def dot_interact_fwd(dense, sparse, Ns, D, copy_dense, copy_sparse):
    B = sparse.shape[0]
    N = Ns + (dense is not None)
    width = N * (N - 1) // 2 + (D if (copy_dense and dense is not None) else 0) + (Ns * D if copy_sparse else 0)
    out = np.empty((B, width), dtype=np.float32)
    lib().orc_dot_interact_fwd(_p(dense), D, _p(sparse), Ns * D, B, Ns, D, int(copy_dense), int(copy_sparse), _p(out),
                               width)
    return out


def dot_interact_bwd(dense, sparse, d_out, Ns, D, copy_dense, copy_sparse):
    B = sparse.shape[0]
    d_dense = np.empty((B, D), dtype=np.float32) if dense is not None else None
    d_sparse = np.empty((B, Ns * D), dtype=np.float32)
    lib().orc_dot_interact_bwd(_p(dense), D, _p(sparse), Ns * D, _p(d_out), d_out.shape[1], B, Ns, D, int(copy_dense),
                               int(copy_sparse), _p(d_dense), D, _p(d_sparse), Ns * D)
    return d_dense, d_sparse


def fm_fwd(x2d, N, D):
    B = x2d.shape[0]
    y = np.empty((B, D), dtype=np.float32)
    lib().orc_fm_fwd(_p(x2d), N * D, B, N, D, _p(y), D)
    return y


def fm_bwd(x2d, dy, N, D):
    B = x2d.shape[0]
    dx = np.empty((B, N * D), dtype=np.float32)
    lib().orc_fm_bwd(_p(x2d), N * D, _p(dy), D, B, N, D, _p(dx), N * D)
    return dx
