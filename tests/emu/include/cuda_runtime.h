// TEST INFRASTRUCTURE -- not part of the product, never shipped, never loaded
// by pylabolt_b200.  See tests/emu/README.md.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the libplb sources
// (pylabolt_b200/csrc/*.cu, unchanged) into tests/emu/_build/libplb_emu.so:
// "device" memory is host memory, streams execute in issue order, and every
// kernel launch runs its threads as cooperative fibers with warp shuffles,
// votes and __syncthreads (simt.h).  Purpose: check the INDEXING and the
// control flow of the kernels and of the host-side step logic on a machine
// without a GPU.  It says nothing about performance and it is not a fallback:
// the product loader only ever opens pylabolt_b200/lib/libplb*.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#define PLB_EMU_RUNTIME 1

// ---- qualifiers ------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

// ---- runtime API (implemented in tests/emu/emu_runtime.cpp) ------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1,
       cudaErrorNotSupported = 801 };
typedef struct plb_emu_stream *cudaStream_t;
typedef struct plb_emu_event *cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1,
                      cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3,
                      cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2,
       cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrClockRate = 13 };

extern "C" {
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError(void);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
cudaError_t cudaSetDevice(int);
cudaError_t cudaDeviceGetAttribute(int *, cudaDeviceAttr, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int *, int *);
cudaError_t cudaDeviceGetPCIBusId(char *, int, int);
cudaError_t cudaFree(void *);
cudaError_t cudaMemset(void *, int, size_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *, unsigned, int);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *);
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void *);
cudaError_t plb_emu_malloc(void **, size_t);
cudaError_t plb_emu_host_alloc(void **, size_t);
}
template <typename T>
static inline cudaError_t cudaMalloc(T **p, size_t n)
{
    return plb_emu_malloc(reinterpret_cast<void **>(p), n);
}
template <typename T>
static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned)
{
    return plb_emu_host_alloc(reinterpret_cast<void **>(p), n);
}

#include "simt.h"
