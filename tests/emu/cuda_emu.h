// cuda_emu.h -- TEST INFRASTRUCTURE.  A minimal host stand-in for the CUDA constructs csrc/proposals.cu uses, so the
// CPU suite can compile that very file with g++ (-x c++ -DYOLAT_HOST_EMU) and check its logic against the oracle
// without a GPU.  "Device" pointers are host pointers, CTAs run one after another.
//   default             one thread per CTA (blockDim.x == 1), __syncthreads() is a no-op: checks the logic.
//   -DEMU_THREADS=N     N host threads per CTA (N a power of two), __syncthreads() is a pthread barrier, __syncwarp() a
//                       barrier of the thread's group of min(N, 32) consecutive threads, atomics are
//                       real atomics: checks the same logic under real concurrency, and -- built with
//                       -fsanitize=thread -- reports shared / global accesses that no barrier orders (missing
//                       __syncthreads()).
// The -m gpu tests still run the real kernels.  Never linked into the product.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "../../include/yolat_b200.h"

#ifndef EMU_THREADS
#define EMU_THREADS 1
#endif

#define __global__ static
#define __device__ static
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx = {0, 0, 0}, blockDim = {EMU_THREADS, 1, 1}, gridDim = {1, 1, 1};

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }

#if EMU_THREADS == 1
static emu_dim3 threadIdx = {0, 0, 0};
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
#define PROP_LAUNCH(kern, grid, block, st, ...)                              \
  do {                                                                       \
    gridDim.x = (unsigned)(grid);                                            \
    for (unsigned _b = 0; _b < (unsigned)(grid); ++_b) {                     \
      blockIdx.x = _b;                                                       \
      kern(__VA_ARGS__);                                                     \
    }                                                                        \
  } while (0)
#else
#include <pthread.h>
#include <thread>
#include <vector>
static thread_local emu_dim3 threadIdx = {0, 0, 0};
#define EMU_GROUP (EMU_THREADS < 32 ? EMU_THREADS : 32)        /* threads per "warp" */
static pthread_barrier_t emu_barrier, emu_group_barrier[EMU_THREADS / EMU_GROUP];
static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }
static inline void __syncwarp() { pthread_barrier_wait(&emu_group_barrier[threadIdx.x / EMU_GROUP]); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
// one CTA at a time; its EMU_THREADS threads are joined before the next CTA / kernel starts (= stream order)
#define PROP_LAUNCH(kern, grid, block, st, ...)                                          \
  do {                                                                                   \
    gridDim.x = (unsigned)(grid);                                                        \
    for (unsigned _b = 0; _b < (unsigned)(grid); ++_b) {                                 \
      blockIdx.x = _b;                                                                   \
      pthread_barrier_init(&emu_barrier, nullptr, EMU_THREADS);                          \
      for (auto& _gb : emu_group_barrier) pthread_barrier_init(&_gb, nullptr, EMU_GROUP); \
      std::vector<std::thread> _ts;                                                      \
      for (unsigned _t = 0; _t < EMU_THREADS; ++_t)                                      \
        _ts.emplace_back([&, _t] { threadIdx.x = _t; kern(__VA_ARGS__); });              \
      for (auto& _th : _ts) _th.join();                                                  \
      pthread_barrier_destroy(&emu_barrier);                                             \
      for (auto& _gb : emu_group_barrier) pthread_barrier_destroy(&_gb);                 \
    }                                                                                    \
  } while (0)
#endif

#define YOLAT_CHECK_LAUNCH() do {} while (0)
#define YOLAT_TRY(expr)            \
  do {                             \
    int _s = (expr);               \
    if (_s != YOLAT_OK) return _s; \
  } while (0)

namespace yolat {
constexpr int kNumSMs = 148;
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }
}  // namespace yolat
