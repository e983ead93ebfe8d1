"""EXPERIMENTAL drop-in for torcheasyrec_b200.dense_gemm._LinearFn on the one wide tower layer (N = 64 outputs, input
width a multiple of 112 — DLRM's 783-wide final-MLP input travels as [B, 784]): the three library GEMMs
(cuBLASLt BF16x9) become tzk_gemm3x.cu's forward (bias + ReLU fused in the epilogue), dgrad and wgrad kernels.

    y  = act(x @ W^T + b)            tzk_gemm3x(x, W)                 [M, 784] x [64, 784]^T
    dx = dz @ W                      tzk_gemm3x(dz, W^T)              [M, 64]  x [784, 64]^T   (W^T materialised: 200 KB)
    dW = dz^T @ x                    tzk_wgrad3x(x, dz)               fixed-order slab reduction
    dz = dy * (y > 0), db = colsum(dz)   unchanged (tzk act_bwd_colsum on the GPU)

`lib` is either the sm_100a build (`load_gpu_lib()`) or, in tests/test_experimental_gemm3x_emu.py, the host build that
runs the same source under the tcgen05 emulation — the autograd wiring below is exercised on the CPU that way.
Not imported by the package; round 2 wires it into dense_gemm._LinearFn behind an environment switch once
try_gemm3x.py has passed on hardware."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from torcheasyrec_b200 import dense_gemm as _G  # noqa: E402
from torcheasyrec_b200.dense_gemm import Gemm3xLinearFn, gemm3x, wgrad3x  # noqa: E402,F401

declare = _G._declare_gemm3x
supported = _G.gemm3x_supported


def __getattr__(name):          # SLABS lives in the product module (tests patch it there)
    return getattr(_G, name)
