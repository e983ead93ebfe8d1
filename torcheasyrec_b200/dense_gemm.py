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
    """y = act(x @ W^T + b), act = ReLU or identity (tzrec/modules/mlp.py Perceptron).  GEMMs: cuBLASLt BF16x9;
    bias+ReLU and ReLU-backward+bias-gradient are one tzk kernel each instead of four ATen passes.
    `x` may carry zero columns beyond W's in_features (width rounded up to a multiple of 4 by the producer):
    16-B aligned rows are what lets cuBLASLt pick the tensor-core kernels instead of the align1 SIMT ones —
    e.g. DLRM's 783-wide final-MLP input travels as [B, 784]."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, in_map):
        from .kernels import default_kernels

        K, Kx = weight.shape[1], x.shape[1]
        w = weight
        if Kx != K:      # zero-padded input: pad the weight the same way (N x Kx, a few hundred KB)
            w = torch.zeros((weight.shape[0], Kx), dtype=weight.dtype, device=weight.device)
            for (src, dst, n) in (in_map or ((0, 0, K),)):
                w[:, dst:dst + n].copy_(weight[:, src:src + n])
        ctx.in_map = in_map
        y = _mm(x, False, w, True, x.shape[0], w.shape[0], Kx)
        if bias is not None or relu:
            default_kernels().bias_act(y, bias, relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.has_bias, ctx.relu, ctx.K = bias is not None, relu, K
        return y

    @staticmethod
    def backward(ctx, dy):
        from .kernels import default_kernels

        x, w, y = ctx.saved_tensors
        dy = dy if (dy.stride(1) == 1 and dy.stride(0) >= dy.shape[1]) else dy.contiguous()
        N = w.shape[0]
        dx = dw = db = None
        fused = (256 % N == 0) and (ctx.has_bias or ctx.relu)
        if fused:       # dz = dy * (y > 0) and db = sum_rows(dz) in one pass
            dz, colsum = default_kernels().act_bwd_colsum(dy, y, ctx.relu, want_dz=ctx.relu)
            if not ctx.relu:
                dz = dy.contiguous()
            db = colsum if ctx.has_bias else None
        else:
            dz = dy * (y > 0) if ctx.relu else dy.contiguous()
            db = dz.sum(0) if ctx.has_bias else None
        if ctx.needs_input_grad[0]:
            dx = _mm(dz, False, w, False, dz.shape[0], w.shape[1], N)
        if ctx.needs_input_grad[1]:
            dw = _mm(dz, True, x, False, N, w.shape[1], dz.shape[0])
            if w.shape[1] != ctx.K:
                segs = ctx.in_map or ((0, 0, ctx.K),)
                dw = torch.cat([dw[:, dst:dst + n] for (_, dst, n) in segs], dim=1) if len(segs) > 1 \
                    else dw[:, :ctx.K].contiguous()
        if not ctx.needs_input_grad[2]:
            db = None
        return dx, dw, db, None, None


class _SmallLinearFn(torch.autograd.Function):
    """y = act(x @ W^T + b) for K, N <= 64: one tzk launch forward, one (+ partial reduction) backward
    (csrc/tzk_tower.cu) instead of the 4-10 library launches such a layer costs as GEMM + element-wise passes."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        from .kernels import default_kernels

        w = weight.contiguous()
        y = default_kernels().small_linear_fwd(x, w, bias, relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        from .kernels import default_kernels

        x, w, y = ctx.saved_tensors
        dy = dy if (dy.stride(1) == 1 and dy.stride(0) >= dy.shape[1]) else dy.contiguous()
        dx, dw, db = default_kernels().small_linear_bwd(
            x, w, y, dy, ctx.relu, want_dx=ctx.needs_input_grad[0],
            want_db=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, (dw if ctx.needs_input_grad[1] else None), db, None


# --------------------------------------------------------------------------------------------------------------------
# The one wide layer (N = 64 outputs, input width a multiple of 112: DLRM's 783-wide final-MLP input travels as
# [B, 784]) on hand-written tcgen05 3xTF32 kernels (csrc/tzk_gemm3x.cu -> libtzk_gemm3x.so): forward with bias + ReLU
# in the epilogue, dgrad on W^T, wgrad with a fixed-order slab reduction.  TZK_GEMM3X=1 selects it.
# --------------------------------------------------------------------------------------------------------------------
_G3 = {"lib": None, "tried": False}


def _gemm3x_lib():
    if not _G3["tried"]:
        _G3["tried"] = True
        path = os.path.join(_HERE, "csrc", "libtzk_gemm3x.so")
        if os.path.exists(path) and torch.cuda.is_available():
            _G3["lib"] = _declare_gemm3x(ctypes.CDLL(path))
    return _G3["lib"]


SLABS = 21          # 7 column tiles x 21 row slabs = 147 CTAs


def _declare_gemm3x(L):
    P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.tzk_gemm3x.argtypes = [P, I64, P, I64, P, I64, I32, I32, I32, P, I64, P, P, P]
    L.tzk_wgrad3x.argtypes = [P, I64, P, I64, I64, I32, I32, P, P, I64, P]
    L.tzk_wgrad3x_partial_floats.restype = I64
    L.tzk_wgrad3x_partial_floats.argtypes = [I32, I32]
    return L


def gemm3x_supported(M: int, N: int, Kx: int) -> bool:
    return N == 64 and Kx % 112 == 0 and Kx % 4 == 0 and M >= 1


def _g3_stream(t: torch.Tensor):
    return torch.cuda.current_stream().cuda_stream if t.is_cuda else None


def _g3_check(rc, what):
    if rc:
        raise RuntimeError(f"{what} failed with code {rc}")


def gemm3x(lib, x, w, bias, relu):
    """act(x [M, K] @ w [N, K]^T + bias) -> [M, N]; x rows may be strided (ld = x.stride(0))."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    w_hi, w_lo = torch.empty_like(w), torch.empty_like(w)
    _g3_check(lib.tzk_gemm3x(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), None if bias is None else bias.data_ptr(),
                          M, N, K, int(relu), y.data_ptr(), N, w_hi.data_ptr(), w_lo.data_ptr(), _g3_stream(x)), "tzk_gemm3x")
    return y


def wgrad3x(lib, x, dz, slabs=SLABS):
    """dz [M, 64]^T @ x [M, K] -> [64, K]."""
    M, K = x.shape
    dw = torch.empty((64, K), dtype=torch.float32, device=x.device)
    part = torch.empty(lib.tzk_wgrad3x_partial_floats(K, slabs), dtype=torch.float32, device=x.device)
    _g3_check(lib.tzk_wgrad3x(x.data_ptr(), x.stride(0), dz.data_ptr(), dz.stride(0), M, K, slabs, part.data_ptr(),
                           dw.data_ptr(), K, _g3_stream(x)), "tzk_wgrad3x")
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
        if not gemm3x_supported(x.shape[0], w.shape[0], Kx):
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
            from .kernels import default_kernels

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


def _use_gemm3x(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (os.environ.get("TZK_GEMM3X", "1") == "1" and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
            and weight.dtype == torch.float32 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and gemm3x_supported(x.shape[0], weight.shape[0], x.shape[1]) and _gemm3x_lib() is not None)


SMALL_MAX = 64      # tzk_small_linear_*: K, N <= 64


def _small(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and weight.shape[0] <= SMALL_MAX and weight.shape[1] <= SMALL_MAX and x.shape[1] == weight.shape[1]
            and x.shape[0] >= 1 and x.stride(1) == 1 and x.stride(0) >= x.shape[1]
            and os.environ.get("TZK_SMALL_LINEAR", "1") != "0")


class _BceFn(torch.autograd.Function):
    """mean BCE-with-logits; the forward kernel also writes dloss/dlogits, backward only scales it."""

    @staticmethod
    def forward(ctx, logits, labels):
        from .kernels import default_kernels

        loss, dz = default_kernels().bce_logits_fwd_bwd(logits.contiguous(), labels.contiguous())
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g).view(ctx.shape), None


class _TowerTailFn(torch.autograd.Function):
    """(loss, logits) = BCE(Linear(relu(Linear(y1)))) with every gradient computed in the forward kernel."""

    @staticmethod
    def forward(ctx, y1, w1, b1, w2, b2, labels):
        from .kernels import default_kernels

        loss, logits, dy1, dw1, db1, dw2, db2 = default_kernels().tower_tail_bce(y1, w1, b1, w2, b2, labels)
        ctx.save_for_backward(dy1, dw1, db1, dw2, db2)
        ctx.has = (b1 is not None, b2 is not None)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, g_loss, _g_logits):
        dy1, dw1, db1, dw2, db2 = ctx.saved_tensors
        # (g_loss is 1 in a plain `loss.backward()`; any other scale multiplies through)
        return (dy1 * g_loss, dw1 * g_loss, (db1 * g_loss) if ctx.has[0] else None, dw2 * g_loss,
                (db2 * g_loss) if ctx.has[1] else None, None)


def tower_tail_usable(y1: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, labels: torch.Tensor) -> bool:
    """The fused tail covers fp32 CUDA towers ending K -> N (ReLU) -> 1 with K, N <= 64 and a float label per row."""
    return (y1.is_cuda and y1.dim() == 2 and y1.dtype == torch.float32 and w1.dtype == torch.float32
            and w1.shape[1] == y1.shape[1] <= 64 and w1.shape[0] <= 64 and tuple(w2.shape) == (1, w1.shape[0])
            and labels.dtype == torch.float32 and labels.numel() == y1.shape[0] >= 1 and y1.stride(1) == 1)


def tower_tail_bce(y1, w1, b1, w2, b2, labels):
    """-> (mean BCE loss, logits [M]) of Linear(N, 1)(relu(Linear(K, N)(y1))) against `labels`."""
    return _TowerTailFn.apply(y1, w1, b1, w2, b2, labels)


def bce_with_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """F.binary_cross_entropy_with_logits(logits, labels) (mean); fused fwd+bwd kernel on CUDA fp32 inputs."""
    if (logits.is_cuda and logits.dtype == torch.float32 and labels.dtype == torch.float32 and logits.numel() >= 1
            and os.environ.get("TZK_SMALL_LINEAR", "1") != "0"):
        return _BceFn.apply(logits, labels)
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)


def padded_width(n: int) -> int:
    return (n + 3) // 4 * 4


def _usable(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and x.shape[0] >= 256 and x.stride(1) == 1 and x.stride(0) >= x.shape[1]
            and not torch.backends.cuda.matmul.allow_tf32 and available())


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           in_map=None) -> torch.Tensor:
    """act(F.linear(x, weight, bias)).  `x` may be wider than in_features: zero-padded at the end (width =
    padded_width(in_features)) or, with `in_map` = ((src_col, dst_col, n), ...), with zero columns in between —
    column src of the weight multiplies column dst of x."""
    K = weight.shape[1]
    if in_map is None and _small(x, weight):
        return _SmallLinearFn.apply(x, weight, bias, relu)
    if x.dim() == 2 and (x.shape[1] != K or in_map is not None):
        if in_map is None and x.shape[1] != padded_width(K):
            raise RuntimeError(f"linear: input width {x.shape[1]} does not match in_features {K}")
        if not _usable(x, weight):       # fallback: drop the padding columns
            x = x[:, :K] if in_map is None else torch.cat([x[:, d:d + n] for (_, d, n) in in_map], dim=1)
            in_map = None
    if _usable(x, weight):
        if _use_gemm3x(x, weight):
            return Gemm3xLinearFn.apply(_gemm3x_lib(), x, weight, bias, relu, in_map)
        return _LinearFn.apply(x, weight, bias, relu, in_map)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(y) if relu else y


class Linear(nn.Linear):
    """nn.Linear whose 2-D fp32 CUDA GEMMs run as BF16x9-emulated fp32 on the tensor cores."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return linear(x, self.weight, self.bias)
