"""The fused tower tail (csrc/tzk_tower_tail.cuh: last Perceptron + Linear(N, 1) + mean BCE, forward and backward in one
pass), its SOURCE executed on the host (tests/native/cuda_cpu_shim.h) against a float64 restatement of
tzrec/modules/mlp.py Perceptron -> Linear -> BCEWithLogitsLoss(mean) and its autograd."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

EXP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")
P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


@pytest.fixture(scope="module")
def tail(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libtail_cpu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DTZK_CPU_SHIM", "-Wno-unknown-pragmas", "-I", EXP, "-x", "c++",
                    os.path.join(EXP, "tower_tail_standalone.cu"), "-shared", "-fPIC", "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.tzk_tail_ws.restype = ctypes.c_size_t
    L.tzk_tail_ws.argtypes = [I64, I32, I32]
    L.tzk_tail_run.argtypes = [P, I64, P, P, P, P, P, I64, I32, I32, P, P, I64, P, P, ctypes.c_size_t]
    return L


def reference(y1, w1, b1, w2, b2, lab):
    y1, w1, b1, w2 = (a.astype(np.float64) for a in (y1, w1, b1, w2))
    M = y1.shape[0]
    pre = y1 @ w1.T + b1
    h = np.maximum(pre, 0)
    z = h @ w2 + float(b2)
    loss = np.mean(np.maximum(z, 0) - z * lab + np.log1p(np.exp(-np.abs(z))))
    dz = (1 / (1 + np.exp(-z)) - lab) / M
    dh = np.outer(dz, w2) * (pre > 0)
    return z, loss, dh @ w1, dh.T @ y1, dh.sum(0), h.T @ dz, dz.sum()


@pytest.mark.parametrize("K,N", [(64, 32), (64, 64), (32, 16), (13, 7), (60, 33), (16, 1)])
@pytest.mark.parametrize("M,pad", [(1, 0), (129, 3), (300, 0)])
def test_tower_tail_source_matches_float64(tail, K, N, M, pad):
    rng = np.random.default_rng(K * 100 + N + M)
    y1 = np.maximum(rng.standard_normal((M, K + pad)), 0).astype(np.float32)      # (a ReLU output)
    w1 = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(N)).astype(np.float32)
    w2 = (rng.standard_normal(N) / np.sqrt(N)).astype(np.float32)
    b2 = np.array([0.05], dtype=np.float32)
    lab = (rng.random(M) < 0.3).astype(np.float32)
    logits = np.full(M, np.nan, dtype=np.float32)
    dy1 = np.full((M, K + pad), np.nan, dtype=np.float32)
    out = np.full(N * K + 2 * N + 2, np.nan, dtype=np.float32)
    nb = tail.tzk_tail_ws(M, K, N)
    ws = np.zeros(nb // 4 + 1, dtype=np.float32)
    rc = tail.tzk_tail_run(y1.ctypes.data, K + pad, w1.ctypes.data, b1.ctypes.data, w2.ctypes.data, b2.ctypes.data,
                           lab.ctypes.data, M, K, N, logits.ctypes.data, dy1.ctypes.data, K + pad, out.ctypes.data,
                           ws.ctypes.data, nb)
    assert rc == 0
    z, loss, d_y1, d_w1, d_b1, d_w2, d_b2 = reference(y1[:, :K], w1, b1, w2, b2[0], lab.astype(np.float64))
    np.testing.assert_allclose(logits, z, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dy1[:, :K], d_y1, rtol=1e-4, atol=1e-7)
    if pad:
        assert np.isnan(dy1[:, K:]).all()
    np.testing.assert_allclose(out[:N * K].reshape(N, K), d_w1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[N * K:N * K + N], d_b1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[N * K + N:N * K + 2 * N], d_w2, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[N * K + 2 * N], d_b2, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(out[N * K + 2 * N + 1], loss, rtol=1e-5)
    # run-to-run identical
    out2 = np.empty_like(out)
    tail.tzk_tail_run(y1.ctypes.data, K + pad, w1.ctypes.data, b1.ctypes.data, w2.ctypes.data, b2.ctypes.data,
                      lab.ctypes.data, M, K, N, logits.ctypes.data, dy1.ctypes.data, K + pad, out2.ctypes.data,
                      ws.ctypes.data, nb)
    np.testing.assert_array_equal(out, out2)
