"""fp32-accurate tensor-core Linear for the dense towers (cuBLASLt 12.9 BF16x9 emulation, csrc/tzk_gemm.cpp).

The towers are caller code (plain PyTorch in the reference, tzrec/modules/mlp.py); this only swaps the GEMM
algorithm: same fp32 inputs/outputs, fp32-equivalent accuracy, tensor cores instead of CUDA cores.  Falls back to
torch.nn.functional.linear whenever the library, the device or the algorithm is unavailable (set
TZK_DENSE_GEMM=torch to force the fallback).
"""

import ctypes
import os
from typing import Optional

import torch
from torch import nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "csrc", "libtzk_gemm.so")
_CUBLASLT_CANDIDATES = ["/usr/local/cuda/lib64/libcublasLt.so.12", "/usr/local/cuda-12.9/lib64/libcublasLt.so.12"]
_state = {"lib": None, "version": 0, "tried": False, "ws": {}, "emulated_calls": 0, "plain_calls": 0}


def _load():
    if _state["tried"]:
        return _state["lib"]
    _state["tried"] = True
    if os.environ.get("TZK_DENSE_GEMM", "") == "torch" or not os.path.exists(_LIB_PATH) or not torch.cuda.is_available():
        return None
    try:
        lib = ctypes.CDLL(_LIB_PATH)
        lib.tzg_init.restype = ctypes.c_long
        lib.tzg_init.argtypes = [ctypes.c_char_p]
        lib.tzg_last_error.restype = ctypes.c_char_p
        P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.tzg_matmul.argtypes = [I, I, I, I, I, P, I, P, I, P, I, F, F, I, P, ctypes.c_size_t, P,
                                   ctypes.POINTER(ctypes.c_int)]
        for path in _CUBLASLT_CANDIDATES:
            if os.path.exists(path):
                v = lib.tzg_init(path.encode())
                if v >= 120900:      # BF16x9 emulation exists from cuBLAS 12.9 on
                    _state["lib"], _state["version"] = lib, v
                    break
    except OSError:
        _state["lib"] = None
    return _state["lib"]


def available() -> bool:
    return _load() is not None


def stats():
    return {"cublaslt": _state["version"], "emulated_calls": _state["emulated_calls"],
            "plain_calls": _state["plain_calls"]}


def _workspace(device) -> torch.Tensor:
    ws = _state["ws"].get(device)
    if ws is None:
        ws = torch.empty(64 << 20, dtype=torch.uint8, device=device)
        _state["ws"][device] = ws
    return ws


def _mm(a: torch.Tensor, ta: bool, b: torch.Tensor, tb: bool, M: int, N: int, K: int) -> torch.Tensor:
    """Row-major C[M,N] = op(a) @ op(b) through cuBLASLt BF16x9."""
    lib = _state["lib"]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws = _workspace(a.device)
    used = ctypes.c_int(0)
    rc = lib.tzg_matmul(int(ta), int(tb), M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                        out.data_ptr(), N, 1.0, 0.0, 1, ws.data_ptr(), ws.numel(),
                        torch.cuda.current_stream().cuda_stream, ctypes.byref(used))
    if rc != 0:
        raise RuntimeError(lib.tzg_last_error().decode())
    _state["emulated_calls" if used.value else "plain_calls"] += 1
    return out


class _LinearFn(torch.autograd.Function):
    """y = x @ W^T + b.  `x` may carry zero columns beyond W's in_features (width rounded up to a multiple of 4 by
    the producer): 16-B aligned rows are what lets cuBLASLt pick the tensor-core (BF16x9) kernels instead of the
    align1 SIMT ones — e.g. DLRM's 783-wide final-MLP input travels as [B, 784]."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        K, Kx = weight.shape[1], x.shape[1]
        w = weight
        if Kx != K:      # zero-padded input: pad the weight the same way (N x Kx, a few hundred KB)
            w = torch.zeros((weight.shape[0], Kx), dtype=weight.dtype, device=weight.device)
            w[:, :K].copy_(weight)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.K = K
        y = _mm(x, False, w, True, x.shape[0], w.shape[0], Kx)
        if bias is not None:
            y.add_(bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _mm(dy, False, w, False, dy.shape[0], w.shape[1], w.shape[0])
        if ctx.needs_input_grad[1]:
            dw = _mm(dy, True, x, False, w.shape[0], w.shape[1], dy.shape[0])
            if w.shape[1] != ctx.K:
                dw = dw[:, :ctx.K].contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def padded_width(n: int) -> int:
    return (n + 3) // 4 * 4


def _usable(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and x.shape[0] >= 256 and x.stride(1) == 1 and x.stride(0) >= x.shape[1]
            and not torch.backends.cuda.matmul.allow_tf32 and available())


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """F.linear; accepts `x` zero-padded to padded_width(in_features) columns."""
    K = weight.shape[1]
    if x.dim() == 2 and x.shape[1] != K:
        if x.shape[1] != padded_width(K):
            raise RuntimeError(f"linear: input width {x.shape[1]} does not match in_features {K}")
        if not _usable(x, weight):
            x = x[:, :K]
    if _usable(x, weight):
        return _LinearFn.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


class Linear(nn.Linear):
    """nn.Linear whose 2-D fp32 CUDA GEMMs run as BF16x9-emulated fp32 on the tensor cores."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return linear(x, self.weight, self.bias)
