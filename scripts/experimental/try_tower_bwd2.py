"""First contact of scripts/experimental/tzk_tower_bwd2.cu with hardware (round-2 groundwork).

    timeout 300 python scripts/experimental/try_tower_bwd2.py [M]

For the four narrow DLRM tower layers: dx / dW / db of the draft against float64 and against the library's
tzk_small_linear_bwd, and CUDA-event times of both (the draft has to beat 63 / 58 / 108 / 33 us at M = 65536 to be
worth promoting into tzk_tower.cu; its source already passes tests/test_experimental_kernels_cpu.py on the host)."""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
LIB, SRC = os.path.join(HERE, "libtzk_tower_bwd2.so"), os.path.join(HERE, "tzk_tower_bwd2.cu")


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                        "-std=c++17", "-Xcompiler", "-fPIC", "-shared", SRC, "-o", LIB], check=True)
    L = ctypes.CDLL(LIB)
    P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.tzk_small_linear_bwd2_workspace_bytes.restype = ctypes.c_size_t
    L.tzk_small_linear_bwd2_workspace_bytes.argtypes = [I64, I32, I32]
    L.tzk_small_linear_bwd2.argtypes = [P, I64, P, P, I64, P, I64, I64, I32, I32, I32, P, I64, P, P, P, ctypes.c_size_t, P]
    return L


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    from torcheasyrec_b200 import functional as Fn

    M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    L, k = build(), Fn.backend()
    torch.manual_seed(0)
    for K, N, relu, want_dx in [(13, 64, 1, False), (64, 16, 1, True), (64, 32, 1, True), (32, 1, 0, True)]:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") / K ** 0.5
        y = torch.randn(M, N, device="cuda")
        dy = torch.randn(M, N, device="cuda")
        dx = torch.empty(M, K, device="cuda")
        dw, db = torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
        nb = L.tzk_small_linear_bwd2_workspace_bytes(M, K, N)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def v2():
            rc = L.tzk_small_linear_bwd2(x.data_ptr(), K, w.data_ptr(), y.data_ptr(), N, dy.data_ptr(), N, M, K, N, relu,
                                         dx.data_ptr() if want_dx else None, K, dw.data_ptr(), db.data_ptr(),
                                         ws.data_ptr(), nb, st)
            assert rc == 0, rc

        def v1():
            return k.small_linear_bwd(x, w, y if relu else None, dy, bool(relu), want_dx, True)

        v2()
        torch.cuda.synchronize()
        dz = dy.double() * ((y > 0) if relu else 1.0)
        err = lambda got, ref: ((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
        e_dw, e_db = err(dw, dz.T @ x.double()), err(db, dz.sum(0))
        e_dx = err(dx, dz @ w.double()) if want_dx else 0.0
        dx1, dw1, db1 = v1()
        e1 = err(dw1, dz.T @ x.double())
        print(f"{K:>2}->{N:<2} M={M}: draft rel.err dx {e_dx:.1e} dW {e_dw:.1e} db {e_db:.1e} (library dW {e1:.1e}); "
              f"draft {timed(v2):.1f} us, library {timed(v1):.1f} us", flush=True)


if __name__ == "__main__":
    main()
