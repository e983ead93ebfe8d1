"""All three passes of tzk_gemm3x.cu in one process (one torch import): error against float64 and CUDA-event time.

    timeout 60 python scripts/experimental/try_gemm3x_all.py [M]

Order = increasing risk; every result is printed (flushed) as soon as it exists."""
import ctypes
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from try_gemm3x import build  # noqa: E402

P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32


def timed(fn, n=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def gemm(L, M, K, N, relu, bias, tag):
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") if bias else None
    y = torch.empty(M, N, device="cuda")
    wh, wl = torch.empty_like(w), torch.empty_like(w)
    st = torch.cuda.current_stream().cuda_stream
    ref = x.double() @ w.double().T + (b.double() if bias else 0)
    ref = torch.relu(ref) if relu else ref
    f32 = torch.nn.functional.linear(x, w, b)
    f32 = ((torch.relu(f32) if relu else f32).double() - ref).abs().max().item()
    for stack, tw, raw, split, pf, ring in (("0", "4", "0", "0", "0", "0"), ("1", "4", "0", "0", "0", "0"),
                                            ("1", "8", "0", "0", "0", "0"), ("1", "4", "1", "0", "0", "0"),
                                            ("1", "4", "0", "1", "0", "0"), ("1", "4", "0", "0", "1", "0"),
                                            ("1", "4", "1", "1", "1", "0"), ("1", "8", "1", "1", "1", "0"),
                                            ("1", "4", "1", "1", "0", "1"), ("1", "4", "1", "1", "1", "1")):
        for var, val in zip(("TZK_GEMM3X_STACK", "TZK_GEMM3X_TW", "TZK_GEMM3X_RAW", "TZK_GEMM3X_SPLIT",
                             "TZK_GEMM3X_PREFETCH", "TZK_GEMM3X_RING"), (stack, tw, raw, split, pf, ring)):
            os.environ[var] = val

        def run():
            rc = L.tzk_gemm3x(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr() if bias else None, M, N, K, int(relu),
                              y.data_ptr(), N, wh.data_ptr(), wl.data_ptr(), st)
            assert rc == 0, rc

        y.fill_(float("nan"))
        run()
        torch.cuda.synchronize()
        err = (y.double() - ref).abs().max().item()
        print(f"{tag} M={M} {'stacked' if stack == '1' else '3-mma  '} tw{tw}{' raw' if raw == '1' else ''}{' split' if split == '1' else ''}{' pf' if pf == '1' else ''}{' RING' if ring == '1' else ''}: max err {err:.2e} (torch fp32 {f32:.2e}), "
              f"{timed(run):.1f} us", flush=True)


def wgrad(L, M, K=784, slabs=21):
    torch.manual_seed(M)
    x = torch.randn(M, K, device="cuda")
    dz = torch.randn(M, 64, device="cuda") / M ** 0.5
    dw = torch.full((64, K), float("nan"), device="cuda")
    part = torch.empty(L.tzk_wgrad3x_partial_floats(K, slabs), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = L.tzk_wgrad3x(x.data_ptr(), K, dz.data_ptr(), 64, M, K, slabs, part.data_ptr(), dw.data_ptr(), K, st)
        assert rc == 0, rc

    run()
    torch.cuda.synchronize()
    ref = dz.double().T @ x.double()
    err = (dw.double() - ref).abs().max().item()
    f32 = ((dz.T @ x).double() - ref).abs().max().item()
    first = dw.clone()
    run()
    torch.cuda.synchronize()
    print(f"wgrad M={M}{' pf' if os.environ.get('TZK_GEMM3X_PREFETCH') == '1' else ''}{' RING' if os.environ.get('TZK_GEMM3X_RING') == '1' else ''}: max err {err:.2e} (torch fp32 {f32:.2e}), repeatable {torch.equal(first, dw)}, "
          f"{timed(run):.1f} us", flush=True)


def main():
    t0 = time.time()
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    L = build()
    L.tzk_gemm3x.argtypes = [P, I64, P, I64, P, I64, I32, I32, I32, P, I64, P, P, P]
    L.tzk_wgrad3x.argtypes = [P, I64, P, I64, I64, I32, I32, P, P, I64, P]
    L.tzk_wgrad3x_partial_floats.restype = I64
    L.tzk_wgrad3x_partial_floats.argtypes = [I32, I32]
    gemm(L, 300, 784, 64, True, True, "fwd  ")
    gemm(L, M, 784, 64, True, True, "fwd  ")
    gemm(L, M, 64, 784, False, False, "dgrad")
    os.environ["TZK_GEMM3X_RING"] = "0"
    os.environ["TZK_GEMM3X_PREFETCH"] = "0"
    wgrad(L, 300)
    wgrad(L, M)
    os.environ["TZK_GEMM3X_PREFETCH"] = "1"
    wgrad(L, M)
    os.environ["TZK_GEMM3X_RING"] = "1"
    wgrad(L, 300)
    wgrad(L, M)
    os.environ["TZK_GEMM3X_PREFETCH"] = "0"
    wgrad(L, M)
    print(f"done in {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
