// interact_tc_standalone.cu — host build of torcheasyrec_b200/csrc/tzk_interact_tc.cuh for tests/test_interact_tc_cpu.py:
// the kernels' SOURCE runs on std::threads (cuda_cpu_shim.h) with an emulated mma.sync.m16n8k8 (fragment layouts of the
// PTX ISA: A row-major 16x8, a0/a1 rows g, g+8 at column t, a2/a3 at column t+4; B 8x8, b0/b1 rows t, t+4 at column g;
// C 16x8, c0/c1 row g columns 2t, 2t+1, c2/c3 row g+8) and cvt.rna.tf32.  Checks the index mapping — which lane owns
// what — before the first GPU minute; says nothing about speed.
#ifndef TZK_CPU_SHIM
#error "host-only test build"
#endif
#include "cuda_cpu_shim.h"
#include <stdint.h>

namespace tzk_itc {
inline uint32_t cvt_tf32(float x) {          // round to nearest, ties away from zero, 10 explicit mantissa bits
  uint32_t u = __float_as_uint(x);
  u += 0x1000u;
  return u & ~0x1fffu;
}
inline void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  unsigned A[4][32], Bv[2][32];
  for (int q = 0; q < 4; ++q) tzk_shim::warp_allgather(a[q], A[q]);
  for (int q = 0; q < 2; ++q) tzk_shim::warp_allgather(b[q], Bv[q]);
  const int lane = (int)tzk_shim::t_lane, g = lane >> 2, t = lane & 3;
  // this lane's outputs: rows g, g + 8; columns 2 t, 2 t + 1
  for (int q = 0; q < 4; ++q) {
    const int row = g + 8 * (q >> 1), col = 2 * t + (q & 1);
    float s = c[q];
    for (int k = 0; k < 8; ++k) {
      // A[row][k]: held by lane (row % 8) * 4 + (k % 4) in register (row / 8) + 2 * (k / 4)
      const float av = __uint_as_float(A[(row >> 3) + 2 * (k >> 2)][(row & 7) * 4 + (k & 3)]);
      // B[k][col]: held by lane col * 4 + (k % 4) in register k / 4
      const float bv = __uint_as_float(Bv[k >> 2][col * 4 + (k & 3)]);
      s += av * bv;
    }
    c[q] = s;
  }
}
}  // namespace tzk_itc

#include "../../torcheasyrec_b200/csrc/tzk_interact_tc.cuh"

extern "C" int tzk_itc_fwd(const float* dense, int64_t ld_dense, const float* sparse, int64_t ld_sparse, int64_t B,
                           float* out, int64_t ld_out, int grid) {
  TZK_LAUNCH((tzk_itc::dot_interact27_fwd_tc_kernel), grid, tzk_itc::kWarps * 32, tzk_itc::fwd_smem(), nullptr, dense,
             ld_dense, sparse, ld_sparse, B, out, ld_out);
  return 0;
}

extern "C" int tzk_itc_bwd(const float* dense, int64_t ld_dense, const float* sparse, int64_t ld_sparse,
                           const float* d_out, int64_t ld_dout, int64_t B, float* d_dense, int64_t ld_ddense,
                           float* d_sparse, int64_t ld_dsparse, int grid) {
  TZK_LAUNCH((tzk_itc::dot_interact27_bwd_tc_kernel), grid, tzk_itc::kWarps * 32, tzk_itc::bwd_smem(), nullptr, dense,
             ld_dense, sparse, ld_sparse, d_out, ld_dout, B, d_dense, ld_ddense, d_sparse, ld_dsparse);
  return 0;
}
