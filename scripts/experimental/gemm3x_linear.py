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
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SLABS = 21          # 7 column tiles x 21 row slabs = 147 CTAs


def declare(L):
    P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.tzk_gemm3x.argtypes = [P, I64, P, I64, P, I64, I32, I32, I32, P, I64, P, P, P]
    L.tzk_wgrad3x.argtypes = [P, I64, P, I64, I64, I32, I32, P, P, I64, P]
    L.tzk_wgrad3x_partial_floats.restype = I64
    L.tzk_wgrad3x_partial_floats.argtypes = [I32, I32]
    return L


def load_gpu_lib():
    src, lib = os.path.join(HERE, "tzk_gemm3x.cu"), os.path.join(HERE, "libtzk_gemm3x.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                        "-std=c++17", "-Xcompiler", "-fPIC", "-shared", src, "-o", lib, "-lcuda"], check=True)
    return declare(ctypes.CDLL(lib))


def supported(M: int, N: int, Kx: int) -> bool:
    return N == 64 and Kx % 112 == 0 and Kx % 4 == 0 and M >= 1


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream().cuda_stream if t.is_cuda else None


def _check(rc, what):
    if rc:
        raise RuntimeError(f"{what} failed with code {rc}")


def gemm3x(lib, x, w, bias, relu):
    """act(x [M, K] @ w [N, K]^T + bias) -> [M, N]; x rows may be strided (ld = x.stride(0))."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    w_hi, w_lo = torch.empty_like(w), torch.empty_like(w)
    _check(lib.tzk_gemm3x(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), None if bias is None else bias.data_ptr(),
                          M, N, K, int(relu), y.data_ptr(), N, w_hi.data_ptr(), w_lo.data_ptr(), _stream(x)), "tzk_gemm3x")
    return y


def wgrad3x(lib, x, dz, slabs=SLABS):
    """dz [M, 64]^T @ x [M, K] -> [64, K]."""
    M, K = x.shape
    dw = torch.empty((64, K), dtype=torch.float32, device=x.device)
    part = torch.empty(lib.tzk_wgrad3x_partial_floats(K, slabs), dtype=torch.float32, device=x.device)
    _check(lib.tzk_wgrad3x(x.data_ptr(), x.stride(0), dz.data_ptr(), dz.stride(0), M, K, slabs, part.data_ptr(),
                           dw.data_ptr(), K, _stream(x)), "tzk_wgrad3x")
    return dw


class Gemm3xLinearFn(torch.autograd.Function):
    """Same contract as dense_gemm._LinearFn.apply(x, weight, bias, relu, in_map) with `lib` in front."""

    @staticmethod
    def forward(ctx, lib, x, weight, bias, relu, in_map):
        K, Kx = weight.shape[1], x.shape[1]
        w = weight.contiguous()
        if Kx != K:      # zero-padded / column-mapped input: lay the weight out the same way
            w = torch.zeros((weight.shape[0], Kx), dtype=weight.dtype, device=weight.device)
            for (src, dst, n) in (in_map or ((0, 0, K),)):
                w[:, dst:dst + n].copy_(weight[:, src:src + n])
        if not supported(x.shape[0], w.shape[0], Kx):
            raise ValueError(f"gemm3x covers N = 64 and input widths that are multiples of 112, got {tuple(w.shape)}")
        y = gemm3x(lib, x, w, bias, relu)
        ctx.lib, ctx.in_map, ctx.K = lib, in_map, K
        ctx.has_bias, ctx.relu = bias is not None, relu
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        lib = ctx.lib
        if dy.is_cuda and (ctx.has_bias or ctx.relu):
            from torcheasyrec_b200.kernels import default_kernels

            dz, colsum = default_kernels().act_bwd_colsum(dy.contiguous(), y, ctx.relu, want_dz=ctx.relu)
            dz = dz if ctx.relu else dy.contiguous()
            db = colsum if ctx.has_bias else None
        else:
            dz = (dy * (y > 0) if ctx.relu else dy).contiguous()
            db = dz.sum(0) if ctx.has_bias else None
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dx = gemm3x(lib, dz, w.t().contiguous(), None, False)            # [M, 64] x [Kx, 64]^T
        if ctx.needs_input_grad[2]:
            dw = wgrad3x(lib, x, dz)
            if w.shape[1] != ctx.K:
                segs = ctx.in_map or ((0, 0, ctx.K),)
                dw = torch.cat([dw[:, dst:dst + n] for (_, dst, n) in segs], dim=1) if len(segs) > 1 \
                    else dw[:, :ctx.K].contiguous()
        if not ctx.needs_input_grad[3]:
            db = None
        return None, dx, dw, db, None, None
