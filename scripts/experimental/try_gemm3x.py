"""First contact of scripts/experimental/tzk_gemm3x.cu with hardware (round-2 groundwork).

    timeout 120 python scripts/experimental/try_gemm3x.py [M] [fwd|dgrad|wgrad]

Builds the kernel next to its source, runs y = relu(x @ w^T + b) for x [M, 784] against a float64 reference and
prints the error (target: fp32-GEMM level, <= 1e-6 relative to |x||w| row norms) and the CUDA-event time.  ALWAYS run
it under `timeout`: a wrong mbarrier phase or descriptor hangs the kernel."""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtzk_gemm3x.so")
SRC = os.path.join(HERE, "tzk_gemm3x.cu")


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                        "-std=c++17", "-Xcompiler", "-fPIC", "-shared", SRC, "-o", LIB, "-lcuda"], check=True)
    return ctypes.CDLL(LIB)


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    # "fwd": y = relu(x[M,784] @ w[64,784]^T + b);  "dgrad": dx[M,784] = dz[M,64] @ wT[784,64]^T
    mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
    lib = build()
    P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    if mode == "wgrad":
        return wgrad(lib, M)
    K, N = (784, 64) if mode == "fwd" else (64, 784)
    lib.tzk_gemm3x.argtypes = [P, I64, P, I64, P, I64, I32, I32, I32, P, I64, P, P, P]
    torch.manual_seed(0)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") if mode == "fwd" else None
    y = torch.empty(M, N, device="cuda")
    w_hi, w_lo = torch.empty_like(w), torch.empty_like(w)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.tzk_gemm3x(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr() if b is not None else None, M, N, K,
                            1 if mode == "fwd" else 0, y.data_ptr(), N, w_hi.data_ptr(), w_lo.data_ptr(), st)
        assert rc == 0, rc

    run()
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T
    ref = torch.relu(ref + b.double()) if mode == "fwd" else ref
    err = (y.double() - ref).abs().max().item()
    scale = (x.double().norm(dim=1).max() * w.double().norm(dim=1).max()).item()
    lin = torch.nn.functional.linear(x, w, b)
    fp32 = ((torch.relu(lin) if mode == "fwd" else lin).double() - ref).abs().max().item()
    print(f"M={M}: max abs err {err:.3e} (fp32 F.linear: {fp32:.3e}), relative to |x||w| {err / scale:.3e}")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (incl. the W split and three tensor-map encodes)")
    if os.environ.get("TZK_GEMM3X_STACK") != "1":      # second variant: two MMAs per k-step (stacked W_hi / W_lo)
        os.environ["TZK_GEMM3X_STACK"] = "1"
        run()
        torch.cuda.synchronize()
        err2 = (y.double() - ref).abs().max().item()
        for _ in range(3):
            run()
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        print(f"stacked-B variant: max abs err {err2:.3e}, {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")


def wgrad(lib, M, K=784, slabs=21):
    """dw[64, K] = dz[M, 64]^T @ x[M, K]: 7 column tiles x 21 row slabs = 147 CTAs at K = 784."""
    P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.tzk_wgrad3x.argtypes = [P, I64, P, I64, I64, I32, I32, P, P, I64, P]
    lib.tzk_wgrad3x_partial_floats.restype = I64
    lib.tzk_wgrad3x_partial_floats.argtypes = [I32, I32]
    torch.manual_seed(0)
    x = torch.randn(M, K, device="cuda")
    dz = torch.randn(M, 64, device="cuda") / M ** 0.5
    dw = torch.empty(64, K, device="cuda")
    partial = torch.empty(lib.tzk_wgrad3x_partial_floats(K, slabs), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.tzk_wgrad3x(x.data_ptr(), K, dz.data_ptr(), 64, M, K, slabs, partial.data_ptr(), dw.data_ptr(), K, st)
        assert rc == 0, rc

    run()
    torch.cuda.synchronize()
    ref = dz.double().T @ x.double()
    err = (dw.double() - ref).abs().max().item()
    fp32 = ((dz.T @ x).double() - ref).abs().max().item()
    first = dw.clone()
    run()
    torch.cuda.synchronize()
    print(f"wgrad M={M}: max abs err {err:.3e} (fp32 matmul: {fp32:.3e}); bitwise repeatable: {torch.equal(first, dw)}")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (incl. two tensor-map encodes and the slab reduction)")


if __name__ == "__main__":
    main()
