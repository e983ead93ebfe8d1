"""The tensor-core DLRM interaction kernels (csrc/tzk_interact_tc.cuh), their SOURCE executed on the host: one std::thread
per CUDA thread (tests/native/cuda_cpu_shim.h) with an emulated mma.sync.m16n8k8.tf32 / cvt.rna.tf32.  Checks the fragment
index mapping (which lane owns which element, the permuted contraction index, the output-column permutation of the
backward) and the 3xTF32 split against float64 — before the kernels meet a GPU (tests/test_kernels_gpu.py then holds the
real thing to the reference-generated golden vectors)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

EXP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")
P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
N, D, PAIRS, ROW = 27, 16, 351, 784


@pytest.fixture(scope="module")
def itc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libitc_cpu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DTZK_CPU_SHIM", "-Wno-unknown-pragmas", "-I", EXP, "-x", "c++",
                    os.path.join(EXP, "interact_tc_standalone.cu"), "-shared", "-fPIC", "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.tzk_itc_fwd.argtypes = [P, I64, P, I64, I64, P, I64, I32]
    L.tzk_itc_bwd.argtypes = [P, I64, P, I64, P, I64, I64, P, I64, P, I64, I32]
    return L


def _inputs(B, pad, seed):
    rng = np.random.default_rng(seed)
    dense = rng.standard_normal((B, D + pad)).astype(np.float32)
    sparse = rng.standard_normal((B, 26 * D + pad)).astype(np.float32)
    x = np.concatenate([dense[:, None, :D], sparse[:, :26 * D].reshape(B, 26, D)], axis=1).astype(np.float64)   # [B, 27, 16]
    return dense, sparse, x


@pytest.mark.parametrize("B,grid,pad", [(1, 1, 0), (19, 1, 0), (37, 2, 4)])
def test_forward_source_matches_float64(itc, B, grid, pad):
    dense, sparse, x = _inputs(B, pad, 3 + B)
    out = np.full((B, ROW + pad), np.nan, dtype=np.float32)
    itc.tzk_itc_fwd(dense.ctypes.data, D + pad, sparse.ctypes.data, 26 * D + pad, B, out.ctypes.data, ROW + pad, grid)
    z = x @ x.transpose(0, 2, 1)
    iu = np.triu_indices(N, 1)                       # row-major strict upper triangle = interaction.py:67-71
    np.testing.assert_allclose(out[:, :PAIRS], z[:, iu[0], iu[1]], rtol=2e-6, atol=2e-6)
    assert np.all(out[:, PAIRS] == 0.0)
    np.testing.assert_array_equal(out[:, 352:368], dense[:, :D])
    np.testing.assert_array_equal(out[:, 368:ROW], sparse[:, :26 * D])
    if pad:
        assert np.isnan(out[:, ROW:]).all()


@pytest.mark.parametrize("B,grid,pad", [(1, 1, 0), (19, 1, 0), (37, 2, 4)])
def test_backward_source_matches_float64(itc, B, grid, pad):
    dense, sparse, x = _inputs(B, pad, 40 + B)
    rng = np.random.default_rng(B)
    d_out = rng.standard_normal((B, ROW + pad)).astype(np.float32)
    d_dense = np.full((B, D + pad), np.nan, dtype=np.float32)
    d_sparse = np.full((B, 26 * D + pad), np.nan, dtype=np.float32)
    itc.tzk_itc_bwd(dense.ctypes.data, D + pad, sparse.ctypes.data, 26 * D + pad, d_out.ctypes.data, ROW + pad, B,
                    d_dense.ctypes.data, D + pad, d_sparse.ctypes.data, 26 * D + pad, grid)
    iu = np.triu_indices(N, 1)
    G = np.zeros((B, N, N))
    G[:, iu[0], iu[1]] = d_out[:, :PAIRS]
    S = G + G.transpose(0, 2, 1)
    dx = S @ x + d_out[:, 352:ROW].astype(np.float64).reshape(B, N, D)
    np.testing.assert_allclose(d_dense[:, :D], dx[:, 0], rtol=2e-6, atol=4e-6)
    np.testing.assert_allclose(d_sparse[:, :26 * D], dx[:, 1:].reshape(B, -1), rtol=2e-6, atol=4e-6)
    if pad:
        assert np.isnan(d_dense[:, D:]).all() and np.isnan(d_sparse[:, 26 * D:]).all()
