// sim_cuda.h — TEST INFRASTRUCTURE ONLY.
//
// A host stand-in for the handful of CUDA runtime calls and builtins that csrc/index.cu uses, so that the index's
// MUTATION logic (op queue -> sort -> per-key apply, the sequential at-capacity path with its LRU order array, rehash,
// lookup + recency stamps) can be compiled with g++ (-DKVB_HOST_SIM) and driven against the oracle on a box without a
// GPU (tests/test_index_sim.py).  "Device memory" is malloc, kernels run as nested loops over (block, thread), the
// radix sort is std::stable_sort.  Nothing in the product links or loads this: libkvb.so is built by nvcc from the
// same source without KVB_HOST_SIM and fails loudly without a device.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <numeric>
#include <vector>

#include "kvb.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

struct SimDim3 {
  unsigned x = 1, y = 1, z = 1;
};
inline thread_local SimDim3 blockIdx, threadIdx, blockDim, gridDim;

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0,
                   cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaEventDefault = 0;
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes {
  cudaMemoryType type = cudaMemoryTypeUnregistered;
  int device = 0;
  void* devicePointer = nullptr;
  void* hostPointer = nullptr;
};

template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
  *p = static_cast<T*>(std::malloc(n ? n : 1));
  return *p ? 0 : 2;
}
inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return 0;
}
template <class T>
inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) {
  return cudaMalloc(p, n);
}
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemset(void* p, int v, size_t n) {
  std::memset(p, v, n);
  return 0;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return 0;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) {
  return cudaMemcpy(d, s, n, k);
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = nullptr;
  return 0;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = nullptr;
  return 0;
}
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return 0;
}
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "sim"; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
  *a = cudaPointerAttributes();
  return 0;
}

// single-threaded stand-ins for the device atomics / fences the kernels use
inline uint32_t atomicCAS(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old = *p;
  if (old == cmp) *p = val;
  return old;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long old = *p;
  *p = old + v;
  return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = *p;
  if (v > old) *p = v;
  return old;
}
inline void __threadfence() {}

#define KVB_LAUNCH(kernel, grid, block, stream, ...)                  \
  do {                                                                \
    gridDim.x = (unsigned)(grid);                                     \
    blockDim.x = (unsigned)(block);                                   \
    for (unsigned _b = 0; _b < (unsigned)(grid); ++_b)                \
      for (unsigned _t = 0; _t < (unsigned)(block); ++_t) {           \
        blockIdx.x = _b;                                              \
        threadIdx.x = _t;                                             \
        kernel(__VA_ARGS__);                                          \
      }                                                               \
  } while (0)

namespace kvb {
inline thread_local char sim_err[1024] = "";
inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(sim_err, sizeof(sim_err), fmt, ap);
  va_end(ap);
}
inline const char* get_error() { return sim_err; }
#define KVB_CUDA_TRY(expr)                                      \
  do {                                                          \
    if ((expr) != cudaSuccess) {                                \
      ::kvb::set_error("%s failed (sim)", #expr);               \
      return KVB_ERR_CUDA;                                      \
    }                                                           \
  } while (0)
#define KVB_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      ::kvb::set_error(__VA_ARGS__); \
      return KVB_ERR_INVALID;        \
    }                                \
  } while (0)
template <class F>
static inline int guarded(F&& f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    set_error("out of host memory");
    return KVB_ERR_NOMEM;
  } catch (const std::exception& e) {
    set_error("unexpected exception: %s", e.what());
    return KVB_ERR_INVALID;
  } catch (...) {
    set_error("unexpected exception");
    return KVB_ERR_INVALID;
  }
}
struct DeviceGuard {
  bool ok = true;
  explicit DeviceGuard(int) {}
};
inline void count_launch(int64_t = 1) {}
inline cudaError_t host_alloc_near(int, void** out, size_t bytes, unsigned) { return cudaMalloc(out, bytes); }

// cub::DeviceRadixSort::SortPairs stand-in (stable)
template <class K, class V>
inline void sim_sort_pairs(const K* kin, K* kout, const V* vin, V* vout, int64_t n) {
  std::vector<int64_t> ord((size_t)n);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return kin[a] < kin[b]; });
  for (int64_t i = 0; i < n; ++i) {
    kout[i] = kin[ord[(size_t)i]];
    vout[i] = vin[ord[(size_t)i]];
  }
}
}  // namespace kvb
