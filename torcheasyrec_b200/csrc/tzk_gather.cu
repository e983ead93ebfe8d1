// tzk_gather.cu — K4 pooled gather forward and K4-nobag sequence gather (sm_100a).
//
// HBM-bound byte mover.  Layout decisions:
//  * a CTA owns a tile of TB consecutive samples x all F features, so its output region
//    (TB rows x sum(D) floats) is one contiguous block of HBM written with full 128-B lines, while the
//    per-feature id runs it reads (ids are key-major) are TB*8-byte contiguous runs;
//  * a bag is served by G lanes (G*4 >= D for D <= 128): one 16-B vector load per lane per row, i.e.
//    a D=16 row is one 64-B request = two full 32-B sectors;
//  * U bags per lane-group are kept in flight (offsets -> id -> row is a 3-deep dependent chain, so
//    memory-level parallelism has to come from independent bags);
//  * feature descriptors are staged once per CTA in shared memory.
#include "tzk_common.cuh"

using namespace tzk;

namespace {

struct FeatDesc {
  int64_t w_off;
  int64_t rows;
  int32_t dim;
  int32_t col;
  int32_t pool;
  int32_t pad;
};

constexpr int kThreads = 256;
constexpr int kTB = 32;  // samples per tile
constexpr int kU = 4;    // bags in flight per lane group

template <int G, int VEC>
__global__ void __launch_bounds__(kThreads)
pooled_gather_fwd_kernel(const float* __restrict__ weights, const int64_t* __restrict__ feat_w_off,
                         const int64_t* __restrict__ feat_rows, const int32_t* __restrict__ feat_dim,
                         const int32_t* __restrict__ feat_col, const int32_t* __restrict__ feat_pool,
                         const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets, int F, int B,
                         float* __restrict__ out, int64_t ld_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FeatDesc* fd = reinterpret_cast<FeatDesc*>(smem_raw);
  for (int f = threadIdx.x; f < F; f += kThreads) {
    fd[f].w_off = feat_w_off[f];
    fd[f].rows = feat_rows[f];
    fd[f].dim = feat_dim[f];
    fd[f].col = feat_col[f];
    fd[f].pool = feat_pool[f];
  }
  __syncthreads();

  constexpr int NG = kThreads / G;  // lane groups per CTA
  const int g = threadIdx.x / G;
  const int lane = threadIdx.x % G;
  const int n_tiles = (B + kTB - 1) / kTB;
  const int items = F * kTB;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b0 = tile * kTB;
    for (int i0 = g; i0 < items; i0 += NG * kU) {
      int64_t s[kU], e[kU];
      int fidx[kU], bidx[kU];
      bool ok[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * NG;
        fidx[u] = i / kTB;
        bidx[u] = b0 + (i % kTB);
        ok[u] = (i < items) && (bidx[u] < B);
        s[u] = 0;
        e[u] = 0;
        if (ok[u]) {
          const int64_t bag = (int64_t)fidx[u] * B + bidx[u];
          s[u] = __ldg(offsets + bag);
          e[u] = __ldg(offsets + bag + 1);
        }
      }
      // first id of every bag (the L=1 fast path keeps kU row loads in flight)
      int64_t id0[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) id0[u] = (ok[u] && e[u] > s[u]) ? __ldg(ids + s[u]) : 0;

#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (!ok[u]) continue;
        const FeatDesc d = fd[fidx[u]];
        const int64_t L = e[u] - s[u];
        float* orow = out + (int64_t)bidx[u] * ld_out + d.col;
        for (int c = lane * VEC; c < d.dim; c += G * VEC) {
          if (VEC == 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (L > 0) {
              int64_t id = id0[u];
              if ((uint64_t)id >= (uint64_t)d.rows) id = 0;
              acc = ld_row_f4(weights + d.w_off + id * d.dim + c);
              for (int64_t l = s[u] + 1; l < e[u]; ++l) {
                int64_t idl = __ldg(ids + l);
                if ((uint64_t)idl >= (uint64_t)d.rows) idl = 0;
                acc = f4_add(acc, ld_row_f4(weights + d.w_off + idl * d.dim + c));
              }
              if (d.pool == TZK_POOL_MEAN) acc = f4_scale(acc, 1.0f / (float)L);
            }
            st_stream_f4(orow + c, acc);
          } else {
            float acc = 0.f;
            if (L > 0) {
              int64_t id = id0[u];
              if ((uint64_t)id >= (uint64_t)d.rows) id = 0;
              acc = __ldg(weights + d.w_off + id * d.dim + c);
              for (int64_t l = s[u] + 1; l < e[u]; ++l) {
                int64_t idl = __ldg(ids + l);
                if ((uint64_t)idl >= (uint64_t)d.rows) idl = 0;
                acc += __ldg(weights + d.w_off + idl * d.dim + c);
              }
              if (d.pool == TZK_POOL_MEAN) acc = acc * (1.0f / (float)L);
            }
            orow[c] = acc;
          }
        }
      }
    }
  }
}

// one lane group per id position; f found by binary search over the key boundaries offsets[f*B]
template <int G, int VEC>
__global__ void __launch_bounds__(kThreads)
seq_gather_fwd_kernel(const float* __restrict__ weights, const int64_t* __restrict__ feat_w_off,
                      const int64_t* __restrict__ feat_rows, const int64_t* __restrict__ ids,
                      const int64_t* __restrict__ offsets, int F, int B, int D, int64_t nnz,
                      float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t* key_start = reinterpret_cast<int64_t*>(smem_raw);  // [F+1]
  int64_t* w_off = key_start + (F + 1);
  int64_t* rows = w_off + F;
  for (int f = threadIdx.x; f <= F; f += kThreads) key_start[f] = offsets[(int64_t)f * B];
  for (int f = threadIdx.x; f < F; f += kThreads) {
    w_off[f] = feat_w_off[f];
    rows[f] = feat_rows[f];
  }
  __syncthreads();
  constexpr int NG = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * NG;
  for (int64_t l = (int64_t)blockIdx.x * NG + threadIdx.x / G; l < nnz; l += stride) {
    int lo = 0, hi = F;  // largest f with key_start[f] <= l
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (key_start[mid] <= l) lo = mid; else hi = mid;
    }
    int64_t id = __ldg(ids + l);
    if ((uint64_t)id >= (uint64_t)rows[lo]) id = 0;
    const float* src = weights + w_off[lo] + id * D;
    float* dst = out + l * D;
    for (int c = lane * VEC; c < D; c += G * VEC) {
      if (VEC == 4) st_stream_f4(dst + c, ld_row_f4(src + c));
      else dst[c] = __ldg(src + c);
    }
  }
}

inline int pick_lanes(int max_dim, int vec) {
  int need = (max_dim + vec - 1) / vec;
  int g = 1;
  while (g < need && g < 32) g <<= 1;
  return g;
}

}  // namespace

#define TZK_DISPATCH_G(G_, VEC_, KERNEL, ...)                                                    \
  switch (G_) {                                                                                   \
    case 1: KERNEL<1, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                    \
    case 2: KERNEL<2, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                    \
    case 4: KERNEL<4, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                    \
    case 8: KERNEL<8, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                    \
    case 16: KERNEL<16, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                  \
    default: KERNEL<32, VEC_><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break;                  \
  }

extern "C" int tzk_pooled_gather_fwd(const float* weights, const int64_t* feat_w_off,
                                     const int64_t* feat_rows, const int32_t* feat_dim,
                                     const int32_t* feat_col, const int32_t* feat_pool, const int64_t* ids,
                                     const int64_t* offsets, int32_t F, int32_t B, int32_t max_dim,
                                     int32_t vec_ok, float* out, int64_t ld_out, tzk_stream_t stream) {
  TZK_REQUIRE(F >= 0 && B >= 0, "pooled_gather_fwd: negative F/B");
  if (F == 0 || B == 0) return 0;
  TZK_REQUIRE(weights && feat_w_off && feat_rows && feat_dim && feat_col && feat_pool && offsets && out,
              "pooled_gather_fwd: NULL argument");
  TZK_REQUIRE(max_dim >= 1, "pooled_gather_fwd: max_dim < 1");
  TZK_REQUIRE(F <= 4096, "pooled_gather_fwd: F=%d > 4096 keys per collection", F);
  const int vec = (vec_ok && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)out % 16 == 0) && (ld_out % 4 == 0))
                      ? 4 : 1;
  const int G = pick_lanes(max_dim, vec);
  cudaStream_t st = as_stream(stream);
  const int n_tiles = (B + kTB - 1) / kTB;
  int grid = n_tiles < kSmCountB200 * 8 ? n_tiles : kSmCountB200 * 8;
  size_t smem = (size_t)F * sizeof(FeatDesc);
  if (vec == 4) {
    TZK_DISPATCH_G(G, 4, pooled_gather_fwd_kernel, weights, feat_w_off, feat_rows, feat_dim, feat_col,
                   feat_pool, ids, offsets, F, B, out, ld_out)
  } else {
    TZK_DISPATCH_G(G, 1, pooled_gather_fwd_kernel, weights, feat_w_off, feat_rows, feat_dim, feat_col,
                   feat_pool, ids, offsets, F, B, out, ld_out)
  }
  TZK_CHECK_LAUNCH("pooled_gather_fwd");
  return 0;
}

extern "C" int tzk_seq_gather_fwd(const float* weights, const int64_t* feat_w_off, const int64_t* feat_rows,
                                  const int64_t* ids, const int64_t* offsets, int32_t F, int32_t B,
                                  int32_t D, int64_t nnz, float* out, tzk_stream_t stream) {
  TZK_REQUIRE(F >= 0 && B >= 0 && nnz >= 0 && D >= 1, "seq_gather_fwd: bad sizes");
  if (F == 0 || nnz == 0) return 0;
  TZK_REQUIRE(weights && feat_w_off && feat_rows && ids && offsets && out, "seq_gather_fwd: NULL argument");
  TZK_REQUIRE(F <= 2048, "seq_gather_fwd: F=%d > 2048", F);
  const int vec = (D % 4 == 0 && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)out % 16 == 0)) ? 4 : 1;
  const int G = pick_lanes(D, vec);
  cudaStream_t st = as_stream(stream);
  const int NG = kThreads / G;
  int64_t blocks = ceil_div64(nnz, NG);
  int grid = blocks < kSmCountB200 * 16 ? (int)blocks : kSmCountB200 * 16;
  size_t smem = (size_t)(3 * F + 1) * sizeof(int64_t);
  if (vec == 4) {
    TZK_DISPATCH_G(G, 4, seq_gather_fwd_kernel, weights, feat_w_off, feat_rows, ids, offsets, F, B, D, nnz, out)
  } else {
    TZK_DISPATCH_G(G, 1, seq_gather_fwd_kernel, weights, feat_w_off, feat_rows, ids, offsets, F, B, D, nnz, out)
  }
  TZK_CHECK_LAUNCH("seq_gather_fwd");
  return 0;
}
