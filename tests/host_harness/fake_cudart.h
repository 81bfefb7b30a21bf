// tests/host_harness/fake_cudart.h -- TEST INFRASTRUCTURE. The handful of CUDA runtime calls the product's host code makes,
// answered on the CPU for the emulated build of libdre (tests/host_harness/gen_emu.py): "device" memory is host memory, kernels
// run synchronously inside their launch (simt_emu.h), so streams / events / copies have nothing left to order. Included where the
// product sources include <cuda_runtime.h>; the real header is still used for the TYPES. Never shipped, never a fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

namespace emu_rt {
// cudaMalloc does not zero: poison the block, so that host code or kernels that rely on zeros fail in the emulated suite
inline cudaError_t Malloc(void **p, size_t n)
{
    *p = n ? malloc(n) : nullptr;
    if (*p) memset(*p, 0xA5, n);
    return (*p || !n) ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> inline cudaError_t Malloc(T **p, size_t n) { return Malloc((void **)p, n); }
inline cudaError_t Free(void *p) { free(p); return cudaSuccess; }
inline cudaError_t HostAlloc(void **p, size_t n, unsigned) { return Malloc(p, n); }
inline cudaError_t FreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t HostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t Memset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
inline cudaError_t MemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { return Memset(p, v, n); }
inline cudaError_t Memcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t MemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return Memcpy(d, s, n, k); }
inline cudaError_t StreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t StreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t DeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t StreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t StreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t StreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t DeviceSynchronize() { return cudaSuccess; }
inline cudaError_t EventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t EventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t EventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t EventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t EventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t EventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t EventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t GetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t DeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? 148 : 0; return cudaSuccess; }
inline cudaError_t GetLastError() { return cudaSuccess; }
template <class F> inline cudaError_t FuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t OccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
inline const char *GetErrorString(cudaError_t) { return "emulated CUDA runtime"; }
inline const char *GetErrorName(cudaError_t) { return "emulated"; }
}  // namespace emu_rt

#define cudaMalloc emu_rt::Malloc
#define cudaFree emu_rt::Free
#define cudaHostAlloc emu_rt::HostAlloc
#define cudaFreeHost emu_rt::FreeHost
#define cudaHostGetDevicePointer emu_rt::HostGetDevicePointer
#define cudaMemset emu_rt::Memset
#define cudaMemsetAsync emu_rt::MemsetAsync
#define cudaMemcpy emu_rt::Memcpy
#define cudaMemcpyAsync emu_rt::MemcpyAsync
#define cudaStreamCreateWithFlags emu_rt::StreamCreateWithFlags
#define cudaStreamCreateWithPriority emu_rt::StreamCreateWithPriority
#define cudaDeviceGetStreamPriorityRange emu_rt::DeviceGetStreamPriorityRange
#define cudaStreamDestroy emu_rt::StreamDestroy
#define cudaStreamSynchronize emu_rt::StreamSynchronize
#define cudaStreamWaitEvent emu_rt::StreamWaitEvent
#define cudaDeviceSynchronize emu_rt::DeviceSynchronize
#define cudaEventCreate emu_rt::EventCreate
#define cudaEventCreateWithFlags emu_rt::EventCreateWithFlags
#define cudaEventDestroy emu_rt::EventDestroy
#define cudaEventRecord emu_rt::EventRecord
#define cudaEventSynchronize emu_rt::EventSynchronize
#define cudaEventQuery emu_rt::EventQuery
#define cudaEventElapsedTime emu_rt::EventElapsedTime
#define cudaGetDevice emu_rt::GetDevice
#define cudaDeviceGetAttribute emu_rt::DeviceGetAttribute
#define cudaGetLastError emu_rt::GetLastError
#define cudaFuncSetAttribute emu_rt::FuncSetAttribute
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor emu_rt::OccupancyMaxActiveBlocksPerMultiprocessor
#define cudaGetErrorString emu_rt::GetErrorString
#define cudaGetErrorName emu_rt::GetErrorName
