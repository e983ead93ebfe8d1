"""tcgen05 descriptor encodings of the round-2 GEMM draft (tests/native/tzk_umma_desc.h) against CuTe.

Host-only: `check_umma_desc.cu` includes the CUTLASS / CuTe headers vendored in this image, asks `make_instr_desc` /
`make_umma_desc` what they encode for the draft's tiles (kind::tf32, 128 x 64 / 112, K-major and MN-major,
SWIZZLE_128B), and compares field by field; it also walks every element of both tiles and checks that CuTe's canonical
swizzled layout puts it exactly where the kernel's TMA boxes do.  Nothing is launched, no GPU is needed."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

EXP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")


def _cutlass_include():
    for sp in sys.path:
        for sub in ("flashinfer/data/cutlass/include", "tilelang/3rdparty/cutlass/include"):
            d = os.path.join(sp, sub)
            if os.path.exists(os.path.join(d, "cute", "atom", "mma_traits_sm100.hpp")):
                return d
    hits = glob.glob("/opt/**/cute/atom/mma_traits_sm100.hpp", recursive=True)
    return os.path.dirname(os.path.dirname(os.path.dirname(hits[0]))) if hits else None


def test_umma_descriptors_match_cute(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = _cutlass_include()
    if not os.path.exists(nvcc) or inc is None:
        pytest.skip("needs nvcc and the vendored CuTe headers")
    exe = str(tmp_path / "check_umma_desc")
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torcheasyrec_b200", "csrc")
    subprocess.run([nvcc, "-std=c++17", "-I" + inc, "-I" + csrc, "--expt-relaxed-constexpr", "-o", exe,
                    os.path.join(EXP, "check_umma_desc.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "0 mismatches" in r.stdout and r.stdout.count(" ok") >= 15, r.stdout


def test_gemm3x_barrier_protocol_model():
    """Discrete model of the draft's warp-specialised pipeline (full / ready / empty / acc_full / acc_empty mbarriers,
    asynchronous in-order tensor-core queue) under random schedules: no deadlock, no stale or overwritten stage, every
    epilogue sees exactly its tile's k-blocks — for the forward, dgrad and wgrad loop shapes."""
    sys.path.insert(0, EXP)
    import model_gemm3x_pipeline as m

    for case in m.CASES:
        for seed in range(40):
            assert m.simulate(*case, seed=seed)
    # the model notices a wrong initial parity (this is what a hang on hardware would look like)
    with pytest.raises((m.Deadlock, AssertionError)):
        bad = m.MBar.done
        try:
            m.MBar.done = lambda self, parity: self.parity == parity
            m.simulate(2, 3, 4, 2, seed=0)
        finally:
            m.MBar.done = bad
