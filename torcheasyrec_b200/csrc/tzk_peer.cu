// tzk_peer.cu — the sharded sparse step over peer memory (part of libtzk.so; exchange="peer" in shard_model).
//
// Round-2 groundwork: the sharded sparse step without a single collective call.  Every rank's
// table arena, its bucketized wire buffers and its pooled-output gradient live in symmetric (peer-mapped) memory of
// the NVSwitch domain, and the kernels read them in place:
//
//   forward   peer_pooled_gather_fwd_kernel: the REQUESTER's gather reads each embedding row straight from the owning
//             rank's arena (owner = feat_owner + id / block, the same rule as tzk_dist.cu's dest_of) and pools
//             locally, in bag order — no id exchange, no row exchange, no staging, any bag length, same bits as the
//             unsharded gather.  (Replaces: bucketize -> ids all-to-all -> owner gather -> rows all-to-all -> local
//             pooling; the reference's KJTAllToAll + lookup + PooledEmbeddingsAllToAll)
//   backward  each rank bucketizes its ids into its OWN wire buffer (tzk_bucketize_rw, fixed capacity per
//             destination) and publishes its gradient [B, sum D]; after one barrier the OWNER pulls, per source
//             rank, its chunk of ids / positions and the matching 64-B gradient slices (peer_pull_kernel) into the
//             buffers tzk_fused_bwd already takes; a second barrier closes the step (tables are quiescent again).
//   barrier   peer_barrier_kernel: one flag per (src, dst) pair in symmetric memory, st.release.sys / ld.acquire.sys,
//             epoch kept on the device so the whole step replays as one CUDA graph.
//
// Build + try (next round, 2 GPUs):
//   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/experimental/try_peer.py
#ifdef TZK_CPU_SHIM
#include "../../scripts/experimental/cuda_cpu_shim.h"   // host execution for tests/test_experimental_kernels_cpu.py (not the barrier kernel)
#else
#include <cuda_runtime.h>
#define TZK_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define TZK_UNPAREN(...) __VA_ARGS__
#define TZK_LAUNCH(kernel, grid, block, smem, stream, ...) TZK_UNPAREN kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <stdint.h>

namespace {

constexpr int kMaxPeers = 16;
constexpr int kThreads = 256;
struct Peers { unsigned long long p[kMaxPeers]; };

struct FeatDesc {
  int64_t rows;    // global rows of the table (ids are clamped like the unsharded gather: out of range -> row 0)
  int64_t block;   // row-wise block (>= rows for table-wise)
  int32_t owner;   // first owner rank
  int32_t dim;
  int32_t col;
  int32_t pool;    // 0 sum, 1 mean
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- forward: requester-side gather over peer memory ------------------------------------------------------------
// A CTA owns 32 consecutive samples x all features (its output block is contiguous); a bag is served by G lanes,
// one 16-B load per lane per row; U bags per lane group are in flight because a remote row costs an NVLink round trip.
template <int G>
__global__ void __launch_bounds__(kThreads)
peer_pooled_gather_fwd_kernel(const __grid_constant__ Peers tables, const int64_t* __restrict__ rf_w_off, const int64_t* __restrict__ feat_rows,
                              const int64_t* __restrict__ feat_block, const int32_t* __restrict__ feat_owner,
                              const int32_t* __restrict__ feat_dim, const int32_t* __restrict__ feat_col,
                              const int32_t* __restrict__ feat_pool, const int64_t* __restrict__ ids,
                              const int64_t* __restrict__ offsets, int F, int B, int W, float* __restrict__ out,
                              int64_t ld_out) {
  constexpr int NG = kThreads / G, TB = 32, U = 8;
  TZK_DYN_SMEM(unsigned char, smem_raw);
  FeatDesc* fd = reinterpret_cast<FeatDesc*>(smem_raw);
  int64_t* w_off = reinterpret_cast<int64_t*>(fd + F);                      // [W * F] arena offsets per (rank, feature)
  unsigned long long* base = reinterpret_cast<unsigned long long*>(w_off + (size_t)W * F);   // [W]
  for (int f = threadIdx.x; f < F; f += kThreads) {
    fd[f].rows = feat_rows[f];
    fd[f].block = feat_block[f];
    fd[f].owner = feat_owner ? feat_owner[f] : 0;
    fd[f].dim = feat_dim[f];
    fd[f].col = feat_col[f];
    fd[f].pool = feat_pool[f];
  }
  for (int i = threadIdx.x; i < W * F; i += kThreads) w_off[i] = rf_w_off[i];
  if (threadIdx.x < W) base[threadIdx.x] = tables.p[threadIdx.x];
  __syncthreads();

  auto row_ptr = [&](int f, const FeatDesc& d, int64_t id) -> const float* {
    if ((uint64_t)id >= (uint64_t)d.rows) id = 0;
    int64_t q = id / d.block;
    int64_t r = d.owner + q;
    if (r >= W) { q -= r - (W - 1); r = W - 1; }
    return reinterpret_cast<const float*>(base[r]) + w_off[r * F + f] + (id - q * d.block) * d.dim;
  };

  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  const int n_tiles = (B + TB - 1) / TB;
  const int items = F * TB;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b0 = tile * TB;
    for (int i0 = g; i0 < items; i0 += NG * U) {
      int32_t s[U];                       // nnz < 2^31 (tzk_bucketize_rw's own limit)
      int32_t len[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NG, b = b0 + (i % TB);
        s[u] = 0;
        len[u] = -1;
        if (i < items && b < B) {
          const int64_t bag = (int64_t)(i / TB) * B + b;
          s[u] = (int32_t)__ldg(offsets + bag);
          len[u] = (int32_t)__ldg(offsets + bag + 1) - s[u];
        }
      }
      int64_t id0[U];
#pragma unroll
      for (int u = 0; u < U; ++u) id0[u] = len[u] > 0 ? __ldg(ids + s[u]) : 0;
      float4 acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (len[u] > 0) {
          const int f = (i0 + u * NG) / TB;
          const FeatDesc& d = fd[f];
          if (lane * 4 < d.dim) acc[u] = __ldg(reinterpret_cast<const float4*>(row_ptr(f, d, id0[u]) + lane * 4));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (len[u] < 0) continue;
        const int i = i0 + u * NG, f = i / TB;
        const FeatDesc d = fd[f];
        float* orow = out + (int64_t)(b0 + (i % TB)) * ld_out + d.col;
        for (int c = lane * 4; c < d.dim; c += G * 4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (len[u] > 0) {
            a = (c == lane * 4) ? acc[u] : __ldg(reinterpret_cast<const float4*>(row_ptr(f, d, id0[u]) + c));
            for (int l = 1; l < len[u]; ++l)
              a = f4_add(a, __ldg(reinterpret_cast<const float4*>(row_ptr(f, d, __ldg(ids + s[u] + l)) + c)));
            if (d.pool == 1) {
              const float inv = 1.0f / (float)len[u];
              a = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
            }
          }
          *reinterpret_cast<float4*>(orow + c) = a;
        }
      }
    }
  }
}

#ifndef TZK_CPU_SHIM
// ---- barrier over the NVSwitch domain ------------------------------------------------------------------------------
// pads.p[r] -> rank r's flag array (uint32 [W]) in symmetric memory; flag[src] on rank dst = last epoch src reached.
__global__ void peer_barrier_kernel(const __grid_constant__ Peers pads, int me, int W, uint32_t* __restrict__ epoch) {
  __shared__ uint32_t e;
  if (threadIdx.x == 0) {
    e = *epoch + 1;
    *epoch = e;
  }
  __syncthreads();
  __threadfence_system();               // everything this GPU wrote before the barrier is visible system-wide
  if ((int)threadIdx.x < W) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(pads.p[threadIdx.x]) + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(e) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[me]) + threadIdx.x;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while ((int32_t)(v - e) < 0);
  }
}

#endif  // TZK_CPU_SHIM

// ---- backward: owner-side pull ----------------------------------------------------------------------------------
// recv_counts[src, f] = counts of rank src for destination `me`
__global__ void peer_pull_counts_kernel(const __grid_constant__ Peers counts, int me, int W, int F, int32_t* __restrict__ recv_counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W * F) {
    const int r = i / F, f = i - r * F;
    recv_counts[i] = reinterpret_cast<const int32_t*>(counts.p[r])[me * F + f];
  }
}

// slot s = (src r, j): id / original position from r's wire chunk for me, gradient slice from r's gradient buffer.
// bounds = the owner-side offsets tzk_fused_bwd gets ([W * (F + 1) + 1]: per source F feature runs + the padding
// run); slots in the padding run get id 0 and a zero row (the layout marks that run rows = 0, the update skips it).
template <int G>
__global__ void __launch_bounds__(kThreads)
peer_pull_kernel(const __grid_constant__ Peers wire_ids, const __grid_constant__ Peers wire_pos,
                 const __grid_constant__ Peers grads, int me, int W, int cap, int F, int B, int D,
                 const int32_t* __restrict__ feat_col, const int64_t* __restrict__ bounds, int64_t ld_grad,
                 int64_t* __restrict__ recv_ids, float* __restrict__ recv_g) {
  constexpr int NG = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t n_slots = (int64_t)W * cap;
  for (int64_t s = (int64_t)blockIdx.x * NG + threadIdx.x / G; s < n_slots; s += (int64_t)gridDim.x * NG) {
    const int r = (int)(s / cap);
    const int64_t j = s - (int64_t)r * cap;
    const bool valid = s < __ldg(bounds + (int64_t)r * (F + 1) + F);
    int64_t id = 0;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      id = reinterpret_cast<const int64_t*>(wire_ids.p[r])[(int64_t)me * cap + j];
      const int32_t pos = reinterpret_cast<const int32_t*>(wire_pos.p[r])[(int64_t)me * cap + j];
      const int f = pos / B, b = pos - f * B;                  // one id per bag: position = f * B + b
      const float* src = reinterpret_cast<const float*>(grads.p[r]) + (int64_t)b * ld_grad + __ldg(feat_col + f);
      for (int c = lane * 4; c < D; c += G * 4) {
        gv = *reinterpret_cast<const float4*>(src + c);
        *reinterpret_cast<float4*>(recv_g + s * D + c) = gv;
      }
    } else {
      for (int c = lane * 4; c < D; c += G * 4) *reinterpret_cast<float4*>(recv_g + s * D + c) = gv;
    }
    if (lane == 0) recv_ids[s] = id;
  }
}

int fill(Peers* dst, const uint64_t* host_ptrs, int W) {
  if (W < 1 || W > kMaxPeers) return 1;
  for (int r = 0; r < kMaxPeers; ++r) dst->p[r] = r < W ? host_ptrs[r] : 0ull;
  return 0;
}
int grid_for(int64_t work_ctas) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = (int64_t)sms * 8;
  return (int)(work_ctas < cap ? (work_ctas > 0 ? work_ctas : 1) : cap);
}
}  // namespace

// table_ptrs: HOST array [W] of device addresses (rank r's arena as mapped in THIS process); rf_w_off: device
// [W * F] int64 arena element offset of feature f's table on rank r; the other feature arrays as in tzk.h.
extern "C" int tzk_peer_pooled_gather_fwd(const uint64_t* table_ptrs, const int64_t* rf_w_off, const int64_t* feat_rows,
                                          const int64_t* feat_block, const int32_t* feat_owner, const int32_t* feat_dim,
                                          const int32_t* feat_col, const int32_t* feat_pool, const int64_t* ids,
                                          const int64_t* offsets, int32_t F, int32_t B, int32_t W, int32_t max_dim,
                                          float* out, int64_t ld_out, void* stream) {
  Peers t;
  if (fill(&t, table_ptrs, W) || F <= 0 || B <= 0 || max_dim <= 0 || (max_dim % 4) || (ld_out % 4)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)F * sizeof(FeatDesc) + (size_t)W * F * 8 + (size_t)W * 8;
  const int grid = grid_for((B + 31) / 32);
#define TZK_PEER_LAUNCH(G)                                                                                             \
  TZK_LAUNCH((peer_pooled_gather_fwd_kernel<G>), grid, kThreads, smem, st, t, rf_w_off, feat_rows, feat_block,         \
             feat_owner, feat_dim, feat_col, feat_pool, ids, offsets, F, B, W, out, ld_out)
  if (max_dim <= 16) TZK_PEER_LAUNCH(4);
  else if (max_dim <= 32) TZK_PEER_LAUNCH(8);
  else if (max_dim <= 64) TZK_PEER_LAUNCH(16);
  else TZK_PEER_LAUNCH(32);
#undef TZK_PEER_LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

#ifndef TZK_CPU_SHIM
extern "C" int tzk_peer_barrier(const uint64_t* pad_ptrs, int32_t me, int32_t W, uint32_t* epoch, void* stream) {
  Peers p;
  if (fill(&p, pad_ptrs, W) || me < 0 || me >= W) return 1;
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, me, W, epoch);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
#endif

extern "C" int tzk_peer_pull_counts(const uint64_t* counts_ptrs, int32_t me, int32_t W, int32_t F, int32_t* recv_counts,
                                    void* stream) {
  Peers p;
  if (fill(&p, counts_ptrs, W) || F <= 0) return 1;
  TZK_LAUNCH((peer_pull_counts_kernel), (W * F + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream), p, me, W, F,
             recv_counts);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

extern "C" int tzk_peer_pull(const uint64_t* ids_ptrs, const uint64_t* pos_ptrs, const uint64_t* grad_ptrs, int32_t me,
                             int32_t W, int32_t cap, int32_t F, int32_t B, int32_t D, const int32_t* feat_col,
                             const int64_t* bounds, int64_t ld_grad, int64_t* recv_ids, float* recv_g, void* stream) {
  Peers pi, pp, pg;
  if (fill(&pi, ids_ptrs, W) || fill(&pp, pos_ptrs, W) || fill(&pg, grad_ptrs, W)) return 1;
  if (cap <= 0 || F <= 0 || B <= 0 || D <= 0 || (D % 4) || (ld_grad % 4)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t slots = (int64_t)W * cap;
#define TZK_PEER_LAUNCH(G)                                                                                          \
  TZK_LAUNCH((peer_pull_kernel<G>), grid_for((slots + kThreads / G - 1) / (kThreads / G)), kThreads, 0, st, pi, pp, pg, \
             me, W, cap, F, B, D, feat_col, bounds, ld_grad, recv_ids, recv_g)
  if (D <= 16) TZK_PEER_LAUNCH(4);
  else if (D <= 32) TZK_PEER_LAUNCH(8);
  else if (D <= 64) TZK_PEER_LAUNCH(16);
  else TZK_PEER_LAUNCH(32);
#undef TZK_PEER_LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
