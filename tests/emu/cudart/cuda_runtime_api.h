// TEST INFRASTRUCTURE ONLY — a host-memory stand-in for the handful of CUDA runtime calls the GENERATED HOST CLASS
// (<Name>.cpp) makes, so that the class (lazy mirrors, dirty tracking, pitched <-> reference layout copies, stage
// geometry) can be exercised together with the emulated kernels (tests/emu/cuda_emu.h) on a machine without a GPU.
// "Device" memory is malloc'ed host memory, streams and events are no-ops (every call completes immediately).  Nothing in paraiso_b200/ includes this file: the product links the real libcudart.
#pragma once
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
// OM_EMU_DEVICES=N in the environment: N "devices" (all of them host memory; tests/emu/cudart/nccl.h moves data between them)
static inline int om_emu_device_count() { const char* e = std::getenv("OM_EMU_DEVICES"); const int n = e ? std::atoi(e) : 1; return n > 0 ? n : 1; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = om_emu_device_count(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < om_emu_device_count() ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }   // four "SMs"
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t) {
  if (width > dpitch || width > spitch) return cudaErrorInvalidValue;
  for (size_t r = 0; r < height; ++r) std::memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
