// TEST INFRASTRUCTURE ONLY — runs the generated CUDA kernels on host threads so that the
// schedule logic (rings, lags, phases, ghost writes, reductions) can be checked against the
// oracle on a machine without a GPU.  One std::thread per CUDA thread of a CTA, CTAs executed
// one after another; __syncthreads is a std::barrier, warp shuffles exchange through a per-warp
// buffer.  Nothing in paraiso_b200/ includes this file; the product path needs nvcc + a GPU.
#pragma once
#define OM_EMULATED_INTRINSICS 1
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3e { unsigned x, y, z; };
static thread_local uint3e threadIdx, blockIdx;
static dim3 blockDim, gridDim;

typedef int cudaError_t;
typedef void* cudaStream_t;
static const cudaError_t cudaSuccess = 0;
static const int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
template <class F> static cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
template <class F> static cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return {a, b}; }

using std::max;
using std::min;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline double __longlong_as_double(long long i) { double f; memcpy(&f, &i, 8); return f; }

alignas(16) static unsigned char om_smem[256 * 1024];

struct EmuBlock {
  std::barrier<>* bar;
  std::vector<std::barrier<>*> warp_bar;
  unsigned char xchg[64][32][8];
  unsigned vote = 0;
};
static EmuBlock* emu_block;

static inline void __syncthreads() { emu_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_block->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
static inline int __syncthreads_or(int pred) {
  if (pred) __atomic_fetch_or(&emu_block->vote, 1u, __ATOMIC_SEQ_CST);
  emu_block->bar->arrive_and_wait();
  const int r = __atomic_load_n(&emu_block->vote, __ATOMIC_SEQ_CST) != 0;
  emu_block->bar->arrive_and_wait();      // everybody has read the vote ...
  if (threadIdx.x == 0) emu_block->vote = 0;   // ... before it is cleared for the next one
  emu_block->bar->arrive_and_wait();
  return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  memcpy(emu_block->xchg[wid][lane], &v, sizeof(T));
  emu_block->warp_bar[wid]->arrive_and_wait();
  T r = v;
  if (lane + d < 32) memcpy(&r, emu_block->xchg[wid][lane + d], sizeof(T));
  emu_block->warp_bar[wid]->arrive_and_wait();
  return r;
}

template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  memcpy(emu_block->xchg[wid][lane], &v, sizeof(T));
  emu_block->warp_bar[wid]->arrive_and_wait();
  T r = v;
  if (lane - d >= 0) memcpy(&r, emu_block->xchg[wid][lane - d], sizeof(T));
  emu_block->warp_bar[wid]->arrive_and_wait();
  return r;
}

// cp.async: the copy happens immediately; commit/wait are no-ops
template <int BYTES> static inline void om_cp_async(void* smem, const void* gmem, int src_bytes) {
  if (src_bytes > 0) memcpy(smem, gmem, src_bytes);
  if (src_bytes < BYTES) memset((char*)smem + src_bytes, 0, BYTES - src_bytes);
}
static inline void om_cp_async_commit() {}
// TMA bulk copies complete immediately; the mbarrier calls are no-ops
static inline void om_mbar_init(uint64_t*, unsigned) {}
static inline void om_mbar_init_fence() {}
static inline void om_mbar_expect_tx(uint64_t*, unsigned) {}
static inline void om_bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t*) { memcpy(smem, gmem, bytes); }
static inline void om_mbar_wait(uint64_t*, unsigned) {}
template <int N> static inline void om_cp_async_wait() {}

template <class K, class... Args>
static void om_emu_launch(K kernel, dim3 grid, unsigned nt, size_t smem, Args... args) {
  (void)smem;
  if (nt % 32) { fprintf(stderr, "emu: block size must be a multiple of 32\n"); abort(); }
  blockDim = dim3(nt);
  gridDim = grid;
  EmuBlock blk;
  std::barrier<> bar(nt);
  blk.bar = &bar;
  std::vector<std::unique_ptr<std::barrier<>>> wb;
  for (unsigned w = 0; w < nt / 32; ++w) { wb.emplace_back(new std::barrier<>(32)); blk.warp_bar.push_back(wb.back().get()); }
  emu_block = &blk;
  std::vector<std::thread> ths;
  for (unsigned t = 0; t < nt; ++t) {
    ths.emplace_back([=]() {
      for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          threadIdx = {t, 0, 0};
          blockIdx = {bx, by, bz};
          kernel(args...);
          emu_block->bar->arrive_and_wait();   // CTAs run one after another
        }
    });
  }
  for (auto& th : ths) th.join();
}
#define OM_LAUNCH(kernel, grid, block, smem, stream, ...) om_emu_launch(kernel, grid, block, smem, __VA_ARGS__)
#define OM_DYNAMIC_SMEM(name) (void)0
