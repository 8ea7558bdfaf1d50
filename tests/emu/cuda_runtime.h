// cuda_runtime.h of the CPU kernel emulator -- TEST INFRASTRUCTURE ONLY (tests/emu/README.md).
//
// Shadows the toolkit header when the library's .cu sources are compiled with g++ for tests/_build/libsep_emu.so.
// It provides the subset of the CUDA language and runtime the seplib-b200 kernels use, with the execution model kept:
// every CUDA thread of a block is a fiber (ucontext) with its own stack; __syncthreads() and the *_sync warp
// primitives are rendezvous points of the fiber scheduler (tests/emu/emu_rt.cpp), so block- and warp-cooperative
// code (shared-memory staging, ballots, shuffles, scans) runs with the semantics it has on the device.  Blocks run one
// after the other.  Nothing here is linked into libsep.so.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <functional>

#define SEPGPU_EMU 1
#ifndef __CUDACC__
#define __CUDACC__ 1          // headers that define __host__/__device__ away for plain C++ must not do it twice
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local     // one block at a time PER HOST THREAD (ranks of a decomposed run are threads)
#define __constant__ static

// ---- vector types ----------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __align__(8) int2 { int x, y; };
struct __align__(8) uint2 { unsigned x, y; };
struct __align__(8) float2 { float x, y; };
struct __align__(16) double2 { double x, y; };
struct __align__(16) double4 { double x, y, z, w; };
struct __align__(16) int4 { int x, y, z, w; };
struct __align__(16) uint4 { unsigned x, y, z, w; };
struct __align__(16) float4 { float x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
static inline double4 make_double4(double x, double y, double z, double w) { double4 r = {x, y, z, w}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }

extern "C" long long sepgpu_emu_counter[8];      // work statistics kernels may bump under #ifdef SEPGPU_EMU
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

// ---- the fiber scheduler ----------------------------------------------------------------------------------------
namespace emu {
dim3 d3(dim3 v);
static inline dim3 d3(long long v) { return dim3((unsigned)v, 1, 1); }
static inline dim3 d3(int v) { return dim3((unsigned)v, 1, 1); }
static inline dim3 d3(unsigned v) { return dim3(v, 1, 1); }
static inline dim3 d3(size_t v) { return dim3((unsigned)v, 1, 1); }
void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void *dyn_smem();
const char *self_path();         // file name of the emulated library (dladdr), for the NCCL stand-in
void sync_block();
// every lane named in mask deposits 8 bytes; returns the warp's 32 deposit slots and, in *part, which lanes took part in
// this exchange (slots of other lanes are stale)
const uint64_t *warp_exchange(unsigned mask, uint64_t mine, unsigned *part);
int lane_id();
// gather statistics (SEPGPU_EMU_GATHER_STATS=1): kernels mark their scattered loads with SEPGPU_EMU_GATHER(ptr); the k-th
// marked load of every lane of a warp is taken as one warp-wide request, and the distinct 128-byte lines it touches are
// counted -- the L1 wavefront count the force kernels are bound by.  sepgpu_emu_counter[1] += requests, [2] += lines,
// [3] += lane loads, per launch.
void record_gather(const void *p);
static inline double rcp_approx(double x) { return (double)(1.0f / (float)x); }
static inline double rsqrt_approx(double x) { return (double)(1.0f / sqrtf((float)x)); }
}  // namespace emu

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned part; emu::warp_exchange(mask, 0, &part); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline long long clock64()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 2000000000LL + 2 * (long long)ts.tv_nsec;     // "2 GHz"
}

template <class T> static inline uint64_t emu_pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "8-byte shuffles only"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// a source lane that did not take part (exited, or not named in the mask) yields the caller's own value
template <class T> static inline T emu_pick(const uint64_t *slots, unsigned part, int src, T own)
{
    return (src >= 0 && src < 32 && (part >> src & 1u)) ? emu_unpack<T>(slots[src]) : own;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    unsigned part;
    const uint64_t *slots = emu::warp_exchange(mask, emu_pack(v), &part);
    const int lane = emu::lane_id();
    return emu_pick(slots, part, (lane & ~(width - 1)) + (src & (width - 1)), v);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    unsigned part;
    const uint64_t *slots = emu::warp_exchange(mask, emu_pack(v), &part);
    const int lane = emu::lane_id();
    const int src = lane ^ lanemask;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return emu_pick(slots, part, src, v);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    unsigned part;
    const uint64_t *slots = emu::warp_exchange(mask, emu_pack(v), &part);
    const int lane = emu::lane_id();
    const int src = lane - (int)delta;
    if (src < (lane & ~(width - 1))) return v;
    return emu_pick(slots, part, src, v);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    unsigned part;
    const uint64_t *slots = emu::warp_exchange(mask, emu_pack(v), &part);
    const int lane = emu::lane_id();
    const int src = lane + (int)delta;
    if (src > (lane | (width - 1))) return v;
    return emu_pick(slots, part, src, v);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    unsigned part;
    const uint64_t *slots = emu::warp_exchange(mask, pred ? 1 : 0, &part);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if ((part >> l & 1u) && (mask >> l & 1u) && slots[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred)
{
    // all taking-part lanes true <=> no taking-part lane false
    return __ballot_sync(mask, !pred) == 0;
}

// ---- scalar intrinsics ----------------------------------------------------------------------------------------------
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline int __float_as_int(float x) { int r; memcpy(&r, &x, 4); return r; }
static inline unsigned __float_as_uint(float x) { unsigned r; memcpy(&r, &x, 4); return r; }
static inline float __int_as_float(int x) { float r; memcpy(&r, &x, 4); return r; }
static inline float __uint_as_float(unsigned x) { float r; memcpy(&r, &x, 4); return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int k = 0; k < 32; k++) r |= ((x >> k) & 1u) << (31 - k);
    return r;
}
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#ifndef EMU_NO_MINMAX
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
#endif

// cache-hinted loads/stores are plain accesses here
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldca(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }

// atomics: the fibers of a block share one host thread and are only switched at a rendezvous, but kernels of different
// host threads (ranks of a decomposed run) run concurrently and may meet in peer-mapped memory
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double *p, double v)
{
    unsigned long long *q = (unsigned long long *)p, o = __atomic_load_n(q, __ATOMIC_SEQ_CST), n;
    double od;
    do { memcpy(&od, &o, 8); const double nd = od + v; memcpy(&n, &nd, 8); } while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return od;
}
static inline float atomicAdd(float *p, float v)
{
    unsigned *q = (unsigned *)p, o = __atomic_load_n(q, __ATOMIC_SEQ_CST), n;
    float of;
    do { memcpy(&of, &o, 4); const float nf = of + v; memcpy(&n, &nf, 4); } while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return of;
}
template <class T> static inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMax(T *p, T v)
{
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
template <class T> static inline T atomicMin(T *p, T v)
{
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

// ---- runtime API --------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct emuStream *cudaStream_t;
typedef struct emuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaIpcMemHandle_t { char reserved[64]; };

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMallocHost((void **)p, bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = 0);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t s = 0);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
