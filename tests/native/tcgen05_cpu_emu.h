// tcgen05_cpu_emu.h — host emulation of the PTX used by tzk_gemm3x.cu (the functions of tzk_tcgen05_ptx.h), for the
// CPU tests only.  Goes with cuda_cpu_shim.h (one std::thread per CUDA thread, blocks one after another).
//
// What is emulated, from the documented semantics (PTX ISA; CUTLASS cute/arch/{copy_sm90_tma,mma_sm100_desc}.hpp):
//   mbarrier      arrival count + outstanding transaction bytes per phase; try_wait.parity(P) succeeds iff the
//                 current parity != P.
//   TMA 2-D load  box [32 floats x box_rows] from a row-major global tensor, out-of-bounds elements = 0, written to
//                 shared memory as 128-B rows with SWIZZLE_128B (16-B chunk index ^= bits [7,10) of the shared address)
//                 or SWIZZLE_128B_ATOM_32B (32-B chunk index ^= bits [7,9)),
//                 then complete_tx(box bytes) on the barrier.
//   tcgen05.mma   kind::tf32, cta_group::1: decodes the instruction descriptor (M, N, operand majors) and both
//                 shared-memory descriptors (start address, LBO, SBO, layout type), reads A [M x 8] and B [N x 8]
//                 through the canonical layouts — K-major SWIZZLE_128B, MN-major SWIZZLE_128B_BASE32B (the only one
//                 the hardware accepts for 32-bit MN-major operands; anything else aborts) — truncates the operands to TF32 (10 mantissa bits) and
//                 accumulates D[m][n] (+)= sum_k a*b into TMEM lane m, column base + n, rounding the fp32 accumulator
//                 TOWARD ZERO (as the hardware was measured to do).  Executes synchronously,
//                 so tcgen05.commit is a plain arrival.
//   TMEM          128 lanes x 512 columns per CTA; tcgen05.ld 32x32b.x16: thread `lane` of the warp reads TMEM lane
//                 (addr >> 16) + lane, 16 consecutive columns; the warp may only touch lanes [32 * (warp % 4), +32) —
//                 violations abort (hardware would fault or read garbage).
// What is NOT emulated: timing, proxies / fences (no-ops: the mutex inside every barrier operation orders memory),
// cluster scope, allocation contention.  Agreement here shows that the kernel's control flow, descriptor arithmetic,
// swizzle assumptions and epilogue mapping are self-consistent with these semantics — not that the semantics are right;
// the descriptor fields themselves are checked against CuTe separately (check_umma_desc.cu).
#pragma once
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

struct CUtensorMap {            // host stand-in: what cuTensorMapEncodeTiled was told
  const float* base;
  int64_t rows, cols, ld;
  int box_rows;
  bool atom32;                  // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B instead of SWIZZLE_128B
};

namespace tzk_emu {
struct Bar { int count = 0, pending = 0, parity = 0; long tx = 0; };
struct Cta {                     // state of the CTA that is currently running (blocks run one at a time)
  std::mutex mu;
  std::condition_variable cv;
  std::map<uint32_t, Bar> bars;  // keyed by shared-memory offset
  float tmem[128][512];
};
inline Cta& cta() { static Cta c; return c; }
inline void fail(const char* msg) { fprintf(stderr, "tcgen05 emulation: %s\n", msg); abort(); }
inline uint8_t* smem_base() { return tzk_shim::t_dyn; }
inline void maybe_flip(Bar& b) {
  if (b.pending == 0 && b.tx == 0) { b.parity ^= 1; b.pending = b.count; }
}
inline uint32_t swz(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }   // SWIZZLE_128B on a shared-memory address
inline uint32_t swz32(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }  // SWIZZLE_128B_ATOM_32B / _BASE32B
inline float tf32_trunc(float x) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x;
}
}  // namespace tzk_emu

inline uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - tzk_emu::smem_base());
}

// ---- mbarrier ------------------------------------------------------------------------------------------------------
inline void mbar_init(uint64_t* bar, uint32_t count) {
  auto& c = tzk_emu::cta();
  std::lock_guard<std::mutex> g(c.mu);
  tzk_emu::Bar b; b.count = b.pending = (int)count;
  c.bars[smem_u32(bar)] = b;
}
inline void mbar_arrive(uint64_t* bar) {
  auto& c = tzk_emu::cta();
  std::lock_guard<std::mutex> g(c.mu);
  auto& b = c.bars.at(smem_u32(bar));
  if (b.pending <= 0) tzk_emu::fail("more arrivals than the mbarrier's count");
  --b.pending;
  tzk_emu::maybe_flip(b);
  c.cv.notify_all();
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  auto& c = tzk_emu::cta();
  std::lock_guard<std::mutex> g(c.mu);
  auto& b = c.bars.at(smem_u32(bar));
  if (b.pending <= 0) tzk_emu::fail("more arrivals than the mbarrier's count (expect_tx)");
  b.tx += bytes;
  --b.pending;
  tzk_emu::maybe_flip(b);
  c.cv.notify_all();
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  auto& c = tzk_emu::cta();
  std::unique_lock<std::mutex> g(c.mu);
  auto& b = c.bars.at(smem_u32(bar));
  if (!c.cv.wait_for(g, std::chrono::seconds(60), [&] { return b.parity != (int)parity; }))
    tzk_emu::fail("mbarrier wait timed out (deadlock in the pipeline protocol)");
}

// warp-uniform wait (all 32 lanes call it): lane 0 waits, the warp follows — the lockstep real lanes have.  Without it a
// lane thread that the OS descheduled inside the wait can sleep through TWO phase flips of a barrier whose next phase
// does not depend on that lane (the MMA warp's `ready` wait) and then waits forever: an artefact of modelling lanes as
// threads, observed once per ~5000 emulated CTAs.
inline void mbar_wait_all(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// ---- TMA --------------------------------------------------------------------------------------------------------------
inline void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  const uint32_t d0 = smem_u32(dst);
  if (d0 & 1023) tzk_emu::fail("TMA destination of a SWIZZLE_128B box must be 1024-B aligned here");
  uint8_t* sm = tzk_emu::smem_base();
  for (int r = 0; r < map->box_rows; ++r)
    for (int c = 0; c < 32; ++c) {
      const int64_t gr = (int64_t)c1 + r, gc = (int64_t)c0 + c;
      const float v = (gr >= 0 && gr < map->rows && gc >= 0 && gc < map->cols) ? map->base[gr * map->ld + gc] : 0.f;
      const uint32_t lin = d0 + (uint32_t)r * 128u + (uint32_t)c * 4u;
      const uint32_t a = map->atom32 ? tzk_emu::swz32(lin) : tzk_emu::swz(lin);
      memcpy(sm + a, &v, 4);
    }
  auto& ct = tzk_emu::cta();
  std::lock_guard<std::mutex> g(ct.mu);
  auto& b = ct.bars.at(smem_u32(bar));
  b.tx -= (long)map->box_rows * 128;
  if (b.tx < 0) tzk_emu::fail("complete_tx exceeds the bytes the barrier expects");
  tzk_emu::maybe_flip(b);
  ct.cv.notify_all();
}

inline void tma_prefetch_2d(const CUtensorMap*, int, int) {}   // L2 prefetch: no architectural effect

// ---- tcgen05 ----------------------------------------------------------------------------------------------------------
inline void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  if (cols < 32 || cols > 512 || (cols & (cols - 1))) tzk_emu::fail("tcgen05.alloc: columns must be a power of two in [32, 512]");
  if ((threadIdx.x & 31) == 0) *dst_smem = 0;      // .sync.aligned: the whole warp executes it, one result
}
inline void tmem_free(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void fence_mbarrier_init() {}
inline void fence_proxy_async() {}
inline void tc_commit(uint64_t* bar) { mbar_arrive(bar); }   // MMAs execute synchronously in the emulation

inline void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const int M = (int)((idesc >> 24) & 0x1f) << 4, N = (int)((idesc >> 17) & 0x3f) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 2 || ((idesc >> 10) & 7) != 2)
    tzk_emu::fail("instruction descriptor is not F32 += TF32 x TF32");
  if (M != 128 || N < 16 || N > 256 || (N & 15)) tzk_emu::fail("unsupported UMMA shape for cta_group::1 (M = 128, N % 16 == 0)");
  const uint8_t* sm = tzk_emu::smem_base();
  auto elem = [&](uint64_t desc, bool mn_major, int r, int k) -> float {
    const uint32_t type = (uint32_t)((desc >> 61) & 7);
    if (((desc >> 46) & 3) != 1) tzk_emu::fail("smem descriptor: version must be 1 on sm_100");
    const uint32_t start = (uint32_t)(desc & 0x3fff) << 4, lbo = (uint32_t)((desc >> 16) & 0x3fff) << 4,
                   sbo = (uint32_t)((desc >> 32) & 0x3fff) << 4;
    uint32_t a;
    if (!mn_major) {     // K-major, SWIZZLE_128B: ((8,m),(T,2)):((8T,SBO),(1,T))
      if (type != 2) tzk_emu::fail("K-major operand: expected layout type 2 (SWIZZLE_128B)");
      a = tzk_emu::swz(start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * 128u + (uint32_t)k * 4u);
    } else {             // MN-major 32-bit elements: SWIZZLE_128B_BASE32B is the only layout the hardware takes
      if (type != 1)
        tzk_emu::fail("MN-major tf32 operand: layout type must be 1 (SWIZZLE_128B_BASE32B; CUTLASS sm100_smem_selector)");
      a = tzk_emu::swz32(start + (uint32_t)(r / 32) * lbo + (uint32_t)(k / 4) * sbo + (uint32_t)(k % 4) * 128u +
                         (uint32_t)(r % 32) * 4u);
    }
    float v;
    memcpy(&v, sm + a, 4);
    return tzk_emu::tf32_trunc(v);
  };
  auto& ct = tzk_emu::cta();
  const uint32_t col0 = tmem_d & 0xffff, lane0 = tmem_d >> 16;
  if (lane0 != 0 || col0 + (uint32_t)N > 512) tzk_emu::fail("accumulator outside TMEM");
  static thread_local float A[128 * 8], B[256 * 8];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 8; ++k) A[m * 8 + k] = elem(desc_a, a_mn, m, k);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 8; ++k) B[n * 8 + k] = elem(desc_b, b_mn, n, k);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      // the 8 products are exact (11-bit x 11-bit mantissas); their sum is added to the fp32 accumulator with
      // round-toward-zero — measured on a B200 (first hardware run of this kernel: 2.2e-5 max error over K = 784 where
      // round-to-nearest accumulation gives 2.6e-6; the error halves when the accumulate count halves)
      double s = accumulate ? (double)ct.tmem[m][col0 + n] : 0.0;
      for (int k = 0; k < 8; ++k) s += (double)A[m * 8 + k] * (double)B[n * 8 + k];
      float f = (float)s;
      if (std::fabs((double)f) > std::fabs(s)) f = std::nextafterf(f, 0.f);
      ct.tmem[m][col0 + n] = f;
    }
}

inline void tmem_ld16(uint32_t addr, float* v) {
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const uint32_t lane0 = addr >> 16, col = addr & 0xffff;
  if ((int)lane0 != 32 * (warp % 4)) tzk_emu::fail("tcgen05.ld: a warp may only read TMEM lanes 32 * (warp % 4) ...");
  if (col + 16 > 512) tzk_emu::fail("tcgen05.ld beyond TMEM columns");
  for (int i = 0; i < 16; ++i) v[i] = tzk_emu::cta().tmem[lane0 + lane][col + i];
}

inline float tf32_rna(float x) {          // cvt.rna.tf32.f32: round to nearest, ties away from zero
  uint32_t u;
  memcpy(&u, &x, 4);
  if (((u >> 23) & 0xff) != 0xff) u += 0x1000u;
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
