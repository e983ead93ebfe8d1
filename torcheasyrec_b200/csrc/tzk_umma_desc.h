// tzk_umma_desc.h — tcgen05 (UMMA) shared-memory and instruction descriptor encodings used by tzk_gemm3x.cu.
// Kept in a header of their own so that check_umma_desc.cu can compare them, on the host, with what CuTe's
// make_umma_desc / make_instr_desc produce for the same tiles (tests/test_umma_desc.py).
#pragma once
#include <stdint.h>

#ifndef TZK_HD
#ifdef __CUDACC__
#define TZK_HD __host__ __device__ __forceinline__
#else
#define TZK_HD inline
#endif
#endif

constexpr int BM = 128;          // rows per tile (UMMA M)

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major: 1) | [32,46) SBO >> 4 (1024 B between
//   8-row groups) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
TZK_HD uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand of 32-bit elements: the only shared-memory layout tcgen05 accepts is SWIZZLE_128B_BASE32B (layout
// type 1; CUTLASS sm100_common.inl sm100_smem_selector: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout") = 128-B rows whose 32-B chunks are XOR-ed with (row & 3) — what a TMA box written with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B holds.  Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-B units
// (mma_traits_sm100.hpp make_umma_desc<Major::MN>): a swizzle atom is 4 k-rows x 128 B of MN; LBO = bytes between
// consecutive 32-float MN groups, SBO = bytes between consecutive 4-row k groups (one tf32 MMA, K = 8, spans two).
TZK_HD uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;      // SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor (InstrDescriptor): c_format F32 = 1 @ [4,6), a/b format TF32 = 2 @ [7,10) / [10,13),
// K-major A and B (bits 15, 16 = 0), n_dim = N >> 3 @ [17,23), m_dim = M >> 4 @ [24,29)
template <int BN, bool MN_MAJOR = false>
TZK_HD constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (MN_MAJOR ? (1u << 15) | (1u << 16) : 0u) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

