// cuda_cpu_shim.h — compile a plain CUDA kernel (no PTX, no warp intrinsics) with g++ and run it on the host, one
// std::thread per CUDA thread, one block at a time, __syncthreads() as a real barrier.
//
// TEST INFRASTRUCTURE for the drafts in this directory only (there is no GPU in the build container): it lets
// tests/test_experimental_kernels_cpu.py execute the kernels' actual source — index arithmetic, guards, reduction
// order — against numpy before their first contact with hardware.  It says nothing about performance, memory
// ordering or anything that needs PTX, and the product (libtzk.so) never sees it.
//
// A kernel source opts in with
//     #ifdef TZK_CPU_SHIM
//     #include "cuda_cpu_shim.h"
//     #else
//     #include <cuda_runtime.h>
//     #endif
// declares dynamic shared memory with TZK_DYN_SMEM(type, name) and launches with
//     TZK_LAUNCH((kernel<targs>), grid, block, smem_bytes, stream, args...);
#pragma once
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;
constexpr int cudaDevAttrMultiProcessorCount = 16;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }

namespace tzk_shim {
inline thread_local dim3 t_thread, t_block, t_bdim, t_gdim;
inline thread_local std::barrier<>* t_bar = nullptr;   // per emulated thread: launches may run concurrently
inline thread_local unsigned char* t_dyn = nullptr;
inline thread_local std::barrier<>* t_wbar = nullptr;  // the thread's warp (32 consecutive linear thread ids)

template <class Body>
void launch(dim3 grid, dim3 block, size_t smem, Body body) {
  std::vector<unsigned char> dyn(smem + 128);
  unsigned char* base = dyn.data() + (128 - reinterpret_cast<uintptr_t>(dyn.data()) % 128) % 128;
  const unsigned n = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(n);
        std::vector<std::unique_ptr<std::barrier<>>> warps;
        for (unsigned w = 0; w * 32 < n; ++w) warps.emplace_back(new std::barrier<>(n - w * 32 < 32 ? n - w * 32 : 32));
        std::vector<std::thread> threads;
        threads.reserve(n);
        for (unsigned t = 0; t < n; ++t)
          threads.emplace_back([&, t] {
            t_thread = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            t_block = dim3(bx, by, bz);
            t_bdim = block;
            t_gdim = grid;
            t_bar = &bar;
            t_dyn = base;
            t_wbar = warps[t / 32].get();
            body();
            bar.arrive_and_drop();   // a thread that returned early must not strand the others at a barrier
            t_wbar->arrive_and_drop();
          });
        for (auto& th : threads) th.join();
      }
}
}  // namespace tzk_shim

#define threadIdx tzk_shim::t_thread
#define blockIdx tzk_shim::t_block
#define blockDim tzk_shim::t_bdim
#define gridDim tzk_shim::t_gdim
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static
#define __syncthreads() tzk_shim::t_bar->arrive_and_wait()
#define __syncwarp() tzk_shim::t_wbar->arrive_and_wait()
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#define TZK_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(tzk_shim::t_dyn)
#define TZK_UNPAREN(...) __VA_ARGS__
#define TZK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  tzk_shim::launch(dim3(grid), dim3(block), (smem), [&] { TZK_UNPAREN kernel(__VA_ARGS__); })
