// Micro-benchmark of the NVLink access patterns the peer-memory sparse step uses (not part of libtzk.so):
// random 64-B row reads from a peer (requester-side gather), random 64-B row writes to a peer (owner-side push),
// contiguous reads (the owner's pull of wire chunks).  Built in-tree, driven by scripts/peer_microbench.py.
#include <cuda_runtime.h>
#include <stdint.h>

template <int U>
__global__ void __launch_bounds__(256) rand_read64_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          int64_t n, float* __restrict__ dst) {
  const int lane = threadIdx.x & 3;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 2;
  for (int64_t i0 = g; i0 < n; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        const float* p = src + (int64_t)idx[i] * 16 + lane * 4;
        asm("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) *reinterpret_cast<float4*>(dst + i * 16 + lane * 4) = v[u];
    }
  }
}

// every warp instruction touches local AND peer rows (odd index -> a, even -> b), like a row-wise sharded gather
__global__ void __launch_bounds__(256) rand_read64_mixed_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                const int32_t* __restrict__ idx, int64_t n,
                                                                float* __restrict__ dst) {
  constexpr int U = 8;
  const int lane = threadIdx.x & 3;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 2;
  for (int64_t i0 = g; i0 < n; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        const int32_t x = idx[i];
        const float* p = ((x & 1) ? a : b) + (int64_t)x * 16 + lane * 4;
        asm("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) *reinterpret_cast<float4*>(dst + i * 16 + lane * 4) = v[u];
    }
  }
}

__global__ void __launch_bounds__(256) rand_write64_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                           int64_t n, float* __restrict__ dst) {
  const int lane = threadIdx.x & 3;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 2;
  for (int64_t i = g; i < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * 16 + lane * 4);
    *reinterpret_cast<float4*>(dst + (int64_t)idx[i] * 16 + lane * 4) = v;
  }
}

__global__ void __launch_bounds__(256) seq_copy_kernel(const float4* __restrict__ src, int64_t n4, float4* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int bench_rand_read64(const float* src, const int32_t* idx, int64_t n, float* dst, int U, int grid, void* st) {
  cudaStream_t s = (cudaStream_t)st;
  if (U == 1) rand_read64_kernel<1><<<grid, 256, 0, s>>>(src, idx, n, dst);
  else if (U == 4) rand_read64_kernel<4><<<grid, 256, 0, s>>>(src, idx, n, dst);
  else if (U == 8) rand_read64_kernel<8><<<grid, 256, 0, s>>>(src, idx, n, dst);
  else rand_read64_kernel<16><<<grid, 256, 0, s>>>(src, idx, n, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
extern "C" int bench_rand_read64_mixed(const float* a, const float* b, const int32_t* idx, int64_t n, float* dst, int grid, void* st) {
  rand_read64_mixed_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(a, b, idx, n, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
extern "C" int bench_rand_write64(const float* src, const int32_t* idx, int64_t n, float* dst, int grid, void* st) {
  rand_write64_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(src, idx, n, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
extern "C" int bench_seq_copy(const float* src, int64_t n_floats, float* dst, int grid, void* st) {
  seq_copy_kernel<<<grid, 256, 0, (cudaStream_t)st>>>((const float4*)src, n_floats / 4, (float4*)dst);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
