"""The barrier-free narrow-tower backward (csrc/tzk_tower_bwd2.cuh), executed on the CPU from their actual CUDA source.

`cuda_cpu_shim.h` compiles a plain CUDA kernel (no PTX, no warp intrinsics) with g++ and runs it with one std::thread
per CUDA thread and a real barrier for __syncthreads.  That checks what can be wrong before the first GPU minute is
spent — index arithmetic, bounds guards, vector / scalar path selection, reduction coverage — against float64 numpy.
It checks nothing about speed or memory ordering, and none of this code is part of libtzk.so."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

EXP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")
P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


def _host_compile(tmp_path_factory, name):
    out = str(tmp_path_factory.mktemp("shim") / f"lib{name}_cpu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DTZK_CPU_SHIM", "-Wno-unknown-pragmas", "-I", EXP, "-x", "c++",
                    os.path.join(EXP, name + ".cu"), "-shared", "-fPIC", "-o", out], check=True)
    return ctypes.CDLL(out)


def _p(a):
    return None if a is None else a.ctypes.data


@pytest.fixture(autouse=True)
def _tile_kernel_by_default(monkeypatch):
    # (the library keeps the dW tile kernel switched off until it has been through a GPU validation pass)
    monkeypatch.setenv("TZK_SMALL_LINEAR_DW", "1")


# ---- narrow tower layers, backward v2 ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tower(tmp_path_factory):
    L = _host_compile(tmp_path_factory, "tower_bwd2_standalone")
    L.tzk_small_linear_bwd2_workspace_bytes.restype = ctypes.c_size_t
    L.tzk_small_linear_bwd2_workspace_bytes.argtypes = [I64, I32, I32]
    L.tzk_small_linear_bwd2.argtypes = [P, I64, P, P, I64, P, I64, I64, I32, I32, I32, P, I64, P, P, P, ctypes.c_size_t, P]
    return L


@pytest.mark.parametrize("K,N", [(13, 64), (64, 16), (64, 32), (32, 1), (64, 64), (7, 3), (64, 30)])
@pytest.mark.parametrize("M,relu,pad", [(1, 1, 0), (300, 0, 0), (77, 1, 3)])
def test_small_linear_bwd2_source_matches_numpy(tower, K, N, M, relu, pad):
    """The four DLRM tower shapes (13->64, 64->16, 64->32, 32->1), the 128-thread limit case 64x64 (NB x KB = 4 x 8) and
    odd shapes on the scalar paths; `pad` widens the leading dimensions so that rows lose their 16-B alignment."""
    rng = np.random.default_rng(K * 1000 + N * 10 + M + relu)
    ldx, ldy, ldd = K + pad, N + pad, N + pad
    x = rng.standard_normal((M, ldx)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    y = rng.standard_normal((M, ldy)).astype(np.float32)          # the layer's output: only its sign matters (ReLU mask)
    dy = rng.standard_normal((M, ldd)).astype(np.float32)
    dx = np.full((M, K + pad), np.nan, dtype=np.float32)
    dw = np.full((N, K), np.nan, dtype=np.float32)
    db = np.full(N, np.nan, dtype=np.float32)
    nb = tower.tzk_small_linear_bwd2_workspace_bytes(M, K, N)
    ws = np.zeros(nb // 4 + 1, dtype=np.float32)
    rc = tower.tzk_small_linear_bwd2(_p(x), ldx, _p(w), _p(y), ldy, _p(dy), ldd, M, K, N, relu, _p(dx), K + pad, _p(dw),
                                     _p(db), _p(ws), nb, None)
    assert rc == 0
    dz = dy[:, :N].astype(np.float64) * ((y[:, :N] > 0) if relu else 1.0)
    np.testing.assert_allclose(dx[:, :K], dz @ w.astype(np.float64), rtol=2e-5, atol=2e-5)
    if pad:
        assert np.isnan(dx[:, K:]).all()                            # nothing is written beyond the row
    np.testing.assert_allclose(dw, dz.T @ x[:, :K].astype(np.float64), rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(db, dz.sum(0), rtol=2e-5, atol=2e-4)
    # run-to-run deterministic (fixed-order folds)
    dw2, db2 = np.empty_like(dw), np.empty_like(db)
    tower.tzk_small_linear_bwd2(_p(x), ldx, _p(w), _p(y), ldy, _p(dy), ldd, M, K, N, relu, None, 0, _p(dw2), _p(db2),
                                _p(ws), nb, None)
    np.testing.assert_array_equal(dw, dw2)
    np.testing.assert_array_equal(db, db2)


@pytest.mark.parametrize("K,N,pad", [(13, 64, 1)])
def test_small_linear_dw_tiles_several_tiles_per_cta(tower, K, N, pad):
    """M > 592 x 32 rows: every CTA walks two tiles (32 rows + a short one) through both shared-memory stages."""
    M = 19500
    rng = np.random.default_rng(K + N)
    x = rng.standard_normal((M, K + pad)).astype(np.float32)
    w = rng.standard_normal((N, K)).astype(np.float32)
    y = rng.standard_normal((M, N + pad)).astype(np.float32)
    dy = rng.standard_normal((M, N + pad)).astype(np.float32)
    dw = np.full((N, K), np.nan, dtype=np.float32)
    db = np.full(N, np.nan, dtype=np.float32)
    nb = tower.tzk_small_linear_bwd2_workspace_bytes(M, K, N)
    ws = np.zeros(nb // 4 + 1, dtype=np.float32)
    rc = tower.tzk_small_linear_bwd2(_p(x), K + pad, _p(w), _p(y), N + pad, _p(dy), N + pad, M, K, N, 1, None, 0, _p(dw),
                                     _p(db), _p(ws), nb, None)
    assert rc == 0
    dz = dy[:, :N].astype(np.float64) * (y[:, :N] > 0)
    np.testing.assert_allclose(dw, dz.T @ x[:, :K].astype(np.float64), rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(db, dz.sum(0), rtol=2e-5, atol=2e-3)


def test_small_linear_bwd2_row_group_kernel_still_matches(tower, monkeypatch):
    """TZK_SMALL_LINEAR_DW=0 keeps the row-group dW kernel selectable for A/B timing; 64 x 30 (which it cannot map) falls
    through to the tile kernel."""
    monkeypatch.setenv("TZK_SMALL_LINEAR_DW", "0")
    rng = np.random.default_rng(5)
    for K, N in ((64, 32), (64, 30)):
        M = 200
        x = rng.standard_normal((M, K)).astype(np.float32)
        w = rng.standard_normal((N, K)).astype(np.float32)
        y = rng.standard_normal((M, N)).astype(np.float32)
        dy = rng.standard_normal((M, N)).astype(np.float32)
        dw = np.full((N, K), np.nan, dtype=np.float32)
        db = np.full(N, np.nan, dtype=np.float32)
        nb = tower.tzk_small_linear_bwd2_workspace_bytes(M, K, N)
        ws = np.zeros(nb // 4 + 1, dtype=np.float32)
        rc = tower.tzk_small_linear_bwd2(_p(x), K, _p(w), _p(y), N, _p(dy), N, M, K, N, 1, None, 0, _p(dw), _p(db), _p(ws),
                                         nb, None)
        assert rc == 0
        dz = dy.astype(np.float64) * (y > 0)
        np.testing.assert_allclose(dw, dz.T @ x.astype(np.float64), rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(db, dz.sum(0), rtol=2e-5, atol=2e-4)
