// check_umma_desc.cu — HOST-side check of tzk_umma_desc.h against CuTe (no GPU needed: nothing is launched).
//
// tzk_gemm3x.cu builds its tcgen05 descriptors by hand.  This program asks the CUTLASS / CuTe headers vendored in the
// image what THEY would encode for the same tiles and compares:
//   * instruction descriptors (kind::tf32, M = 128, N = 64 / 112, K-major and MN-major),
//   * shared-memory descriptors (layout type, version, LBO, SBO) for one UMMA k-step of a K-major and an MN-major tile,
//   * the byte offset of every element of those tiles under CuTe's canonical SWIZZLE_128B layouts vs. where the
//     kernel's TMA boxes put it (box = 32 floats x rows, 128-B rows; K-major: 16-B chunk index XOR (row & 7),
//     MN-major 32-bit: SWIZZLE_128B_ATOM_32B, 32-B chunk index XOR (row & 3)),
//   * the start offset of the second k-step.
// Exit code = number of disagreements.  Build: nvcc -std=c++17 -I<cutlass/include> --expt-relaxed-constexpr.
#include <cstdio>

#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>

#include "tzk_umma_desc.h"
using namespace cute;

static int bad = 0;
static void expect(const char* what, unsigned long long cute_v, unsigned long long mine) {
  printf("%-44s cute %8llx  mine %8llx  %s\n", what, cute_v, mine, cute_v == mine ? "ok" : "MISMATCH");
  bad += cute_v != mine;
}
static unsigned lbo(uint64_t d) { return (unsigned)((d >> 16) & 0x3fff); }
static unsigned sbo(uint64_t d) { return (unsigned)((d >> 32) & 0x3fff); }
static unsigned ver(uint64_t d) { return (unsigned)((d >> 46) & 0x3); }
static unsigned lay(uint64_t d) { return (unsigned)((d >> 61) & 0x7); }

int main() {
  using T = tfloat32_t;
  expect("idesc tf32 128x64  K-major", (uint32_t)UMMA::make_instr_desc<T, T, float, 128, 64, UMMA::Major::K, UMMA::Major::K>(),
         make_idesc<64>());
  expect("idesc tf32 128x112 K-major", (uint32_t)UMMA::make_instr_desc<T, T, float, 128, 112, UMMA::Major::K, UMMA::Major::K>(),
         make_idesc<112>());
  expect("idesc tf32 128x64  MN-major", (uint32_t)UMMA::make_instr_desc<T, T, float, 128, 64, UMMA::Major::MN, UMMA::Major::MN>(),
         make_idesc<64, true>());

  alignas(1024) static float buf[128 * 32];
  {  // K-major operand tile: 128 rows (MN) x 32 floats (K) = one TMA box, rows of 128 B
    auto layout = tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, _32>{});
    auto t = make_tensor(make_smem_ptr(reinterpret_cast<T*>(buf)), layout);
    uint64_t c = UMMA::make_umma_desc<UMMA::Major::K>(local_tile(t, Shape<_128, _8>{}, make_coord(0, 0)));
    uint64_t m = make_desc(0);
    expect("K-major  smem desc: layout type", lay(c), lay(m));
    expect("K-major  smem desc: version", ver(c), ver(m));
    expect("K-major  smem desc: LBO >> 4", lbo(c), lbo(m));
    expect("K-major  smem desc: SBO >> 4", sbo(c), sbo(m));
    int diff = 0;
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 32; ++k) {
        const int off = (int)(&t(r, k) - &t(0, 0)) * 4;
        const int tma = r * 128 + ((((k * 4) >> 4) ^ (r & 7)) << 4) + ((k * 4) & 15);
        diff += off != tma;
      }
    expect("K-major  element offsets differing from TMA", 0, diff);
    auto s1 = local_tile(t, Shape<_128, _8>{}, make_coord(0, 1));
    expect("K-major  k-step 1 start offset (bytes)", (int)(&s1(0, 0) - &t(0, 0)) * 4, 8 * 4);
  }
  {  // MN-major operand tile: logical (MN = 128, K = 32 batch rows) = 4 TMA boxes [32 floats x 32 rows], box b at b * 4096,
     // written with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; CuTe's only layout for 32-bit MN-major operands
    auto layout = tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<T>{}, Shape<_128, _32>{}, Step<_2, _1>{});
    auto t = make_tensor(make_smem_ptr(reinterpret_cast<T*>(buf)), layout);
    uint64_t c = UMMA::make_umma_desc<UMMA::Major::MN>(local_tile(t, Shape<_128, _8>{}, make_coord(0, 0)));
    uint64_t m = make_desc_mn(0, 4096, 512);
    expect("MN-major smem desc: layout type (BASE32B)", lay(c), lay(m));
    expect("MN-major smem desc: version", ver(c), ver(m));
    expect("MN-major smem desc: LBO >> 4", lbo(c), lbo(m));
    expect("MN-major smem desc: SBO >> 4", sbo(c), sbo(m));
    int diff = 0;
    for (int mn = 0; mn < 128; ++mn)
      for (int k = 0; k < 32; ++k) {
        const int off = (int)(&t(mn, k) - &t(0, 0)) * 4;
        const int b = mn / 32, cb = (mn % 32) * 4;
        const int tma = b * 4096 + k * 128 + (((cb >> 5) ^ (k & 3)) << 5) + (cb & 31);
        diff += off != tma;
      }
    expect("MN-major element offsets differing from TMA", 0, diff);
    auto s1 = local_tile(t, Shape<_128, _8>{}, make_coord(0, 1));
    expect("MN-major k-step 1 start offset (bytes)", (int)(&s1(0, 0) - &t(0, 0)) * 4, 1024);
  }
  printf("%d mismatches\n", bad);
  return bad;
}
