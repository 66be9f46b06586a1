// cuda_emu.h — TEST-ONLY functional emulation of the CUDA execution model on the host CPU.
//
// Purpose: compile fastqpacker_b200/csrc/*.cu with g++ (-x c++ -include cuda_emu.h -DFQZ_EMU
// -fsanitize=address,undefined) so that kernel logic (indexing, warp collectives, barriers,
// buffer sizing) can be debugged here, where there is no GPU, before a gpurun call is spent.
// It is NOT a product path: the emulated build lands in tests/emu/_build/, is loaded only by
// tests/test_emu_*.py, is never loaded by fastqpacker_b200 (which fails loudly without CUDA),
// and is never timed.  One CTA runs at a time; every CUDA thread of the CTA is a ucontext
// fiber; warp collectives and __syncthreads are rendezvous points between fibers.
#pragma once
#ifndef FQZ_EMU
#define FQZ_EMU 1
#endif
#include <ucontext.h>
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __constant__ const
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct uint3_e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
extern uint3_e g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char *dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void yield();
void syncthreads();
int syncthreads_count(int pred);
unsigned long long collective(unsigned mask, unsigned long long v, unsigned long long *all /*[32]*/);
unsigned lane();
}  // namespace emu

#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
static const int warpSize = 32;

// ------------------------------------------------------------------ runtime API shims
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = malloc(n ? n : 1);  // exact size so that ASan sees overruns
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *d) { *d = 1; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; char name[64]; size_t totalGlobalMem; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p); p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; strcpy(p->name, "emu");
    p->totalGlobalMem = 1ull << 34; return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = 1ull << 34; return cudaSuccess; }

extern thread_local unsigned long long g_fqz_launches;
#define FQZ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (++g_fqz_launches, emu::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); }))
#define FQZ_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::dyn_smem)

// ------------------------------------------------------------------ synchronisation
static inline void __syncthreads() { emu::syncthreads(); }
static inline int __syncthreads_count(int p) { return emu::syncthreads_count(p); }
static inline int __syncthreads_or(int p) { return emu::syncthreads_count(p) != 0; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned long long a[32]; emu::collective(mask, 0, a); }

// ------------------------------------------------------------------ warp collectives
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned long long a[32], x = 0; memcpy(&x, &v, sizeof(T)); emu::collective(mask, x, a);
    int l = (int)emu::lane(); int base = l & ~(width - 1); int s = base + (src & (width - 1));
    T r; memcpy(&r, &a[s], sizeof(T)); return r;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned long long a[32], x = 0; memcpy(&x, &v, sizeof(T)); emu::collective(mask, x, a);
    int l = (int)emu::lane(); int base = l & ~(width - 1); int s = l - (int)d; if (s < base) s = l;
    T r; memcpy(&r, &a[s], sizeof(T)); return r;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned long long a[32], x = 0; memcpy(&x, &v, sizeof(T)); emu::collective(mask, x, a);
    int l = (int)emu::lane(); int base = l & ~(width - 1); int s = l + (int)d; if (s >= base + width) s = l;
    T r; memcpy(&r, &a[s], sizeof(T)); return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    unsigned long long a[32], x = 0; memcpy(&x, &v, sizeof(T)); emu::collective(mask, x, a);
    int l = (int)emu::lane(); int base = l & ~(width - 1); int s = l ^ m; if (s >= base + width || s < base) s = l;
    T r; memcpy(&r, &a[s], sizeof(T)); return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned long long a[32]; emu::collective(mask, pred ? 1 : 0, a);
    unsigned r = 0; for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r |= (unsigned)(a[i] & 1) << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    unsigned long long a[32], x = 0; memcpy(&x, &v, sizeof(T)); emu::collective(mask, x, a);
    unsigned r = 0; for (int i = 0; i < 32; i++) if (((mask >> i) & 1) && a[i] == x) r |= 1u << i; return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    unsigned long long a[32]; emu::collective(mask, v, a); unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r += (unsigned)a[i];
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    unsigned long long a[32]; emu::collective(mask, v, a); unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r = std::min(r, (unsigned)a[i]);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    unsigned long long a[32]; emu::collective(mask, v, a); unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r = std::max(r, (unsigned)a[i]);
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    unsigned long long a[32]; emu::collective(mask, v, a); unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r |= (unsigned)a[i];
    return r;
}

// ------------------------------------------------------------------ atomics (single host thread)
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ------------------------------------------------------------------ integer intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long t = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF; unsigned byte = (unsigned)(t >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    unsigned long long t = ((unsigned long long)hi << 32) | lo; return (unsigned)(t >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    unsigned long long t = ((unsigned long long)hi << 32) | lo; return (unsigned)((t << (sh & 31)) >> 32);
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r;
}
static inline unsigned __vcmpltu4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) < ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r;
}
static inline unsigned __vsub4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xFF) - ((b >> (8 * i)) & 0xFF)) & 0xFF) << (8 * i); return r;
}
static inline unsigned __vadd4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xFF) + ((b >> (8 * i)) & 0xFF)) & 0xFF) << (8 * i); return r;
}
static inline unsigned __vminu4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) r |= std::min((a >> (8 * i)) & 0xFF, (b >> (8 * i)) & 0xFF) << (8 * i); return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;
static inline unsigned min(unsigned a, int b) { return std::min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return std::min((unsigned)a, b); }
static inline unsigned long long min(unsigned long long a, unsigned b) { return std::min(a, (unsigned long long)b); }
static inline unsigned long long min(unsigned a, unsigned long long b) { return std::min((unsigned long long)a, b); }
static inline unsigned max(unsigned a, int b) { return std::max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return std::max((unsigned)a, b); }
