/*
 * tests/cuda_emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal CUDA-on-CPU execution model so that the product's kernel sources
 * (complete-striped-smith-waterman-library_b200/csrc/ *.cuh / *.cu) can be
 * compiled with g++ and executed in this GPU-less container by the
 * `-m "not gpu"` tests.  Every CUDA thread of a block is a ucontext fiber;
 * __syncthreads / __shfl_*_sync / __any_sync etc. are rendezvous points at
 * which fibers yield, so warp-synchronous code behaves as on the device.
 *
 * It is NOT a fallback: the shipped libssw.so is built by nvcc only, contains
 * none of this, and fails loudly without a GPU.  The emulated build lives in
 * tests/cuda_emu/libssw_emu.so and is loaded by tests only.
 */
#ifndef SSW_CUDA_EMU_H
#define SSW_CUDA_EMU_H

#include <memory>
#include <thread>
#include <atomic>
#include <map>
#include <mutex>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ucontext.h>
#include <algorithm>
#include <functional>
#include <vector>

/* ---- qualifiers ---------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct int2 { int x, y; };
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r = {a, b}; return r; }
static inline int4 make_int4(int a, int b, int c, int d) { int4 r = {a, b, c, d}; return r; }
static inline int2 make_int2(int a, int b) { int2 r = {a, b}; return r; }

/* ---- fiber scheduler ----------------------------------------------------- */
namespace cuemu {

struct Fiber {
	ucontext_t ctx;
	char* stack;
	dim3 tid;
	int linear;
	bool done;
};

struct WarpSync { int arrived; int phase; unsigned long long slot[32]; };

struct Block {
	dim3 bid, bdim, gdim;
	std::vector<Fiber> fibers;
	std::vector<WarpSync> warps;
	int bar_arrived, bar_phase;
	int current;
	ucontext_t sched;
	std::function<void()> body;
	char* dyn_smem;
};

extern Block* g_block;
void yield_now();
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
inline Fiber& self() { return g_block->fibers[g_block->current]; }
inline int lane_id() { return self().linear & 31; }
inline int warp_id() { return self().linear >> 5; }
int warp_width();                     /* live lanes in my warp (last warp of a block may be partial) */
void warp_barrier();
void block_barrier();

}  // namespace cuemu

#define threadIdx (cuemu::self().tid)
#define blockIdx (cuemu::g_block->bid)
#define blockDim (cuemu::g_block->bdim)
#define gridDim (cuemu::g_block->gdim)
#define warpSize 32

/* dynamic shared memory: `extern __shared__ T name[];` is rewritten by the sources as EMU_DYN_SMEM(T, name) */
#define SSW_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(cuemu::g_block->dyn_smem)

/* ---- synchronisation ------------------------------------------------------ */
static inline void __syncthreads() { cuemu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { return 0xffffffffu; }

template <class T> static inline T cuemu_exchange(T v, int src_lane) {
	static_assert(sizeof(T) <= 8, "shuffle payload too wide");
	cuemu::WarpSync& w = cuemu::g_block->warps[cuemu::warp_id()];
	unsigned long long raw = 0;
	memcpy(&raw, &v, sizeof(T));
	w.slot[cuemu::lane_id()] = raw;
	cuemu::warp_barrier();
	raw = w.slot[src_lane & 31];
	cuemu::warp_barrier();
	T out;
	memcpy(&out, &raw, sizeof(T));
	return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
	int lane = cuemu::lane_id(), base = lane & ~(width - 1);
	return cuemu_exchange(v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
	int lane = cuemu::lane_id(), base = lane & ~(width - 1), src = lane - (int)delta;
	return cuemu_exchange(v, src < base ? lane : src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
	int lane = cuemu::lane_id(), base = lane & ~(width - 1), src = lane + (int)delta;
	return cuemu_exchange(v, src >= base + width ? lane : src);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
	int lane = cuemu::lane_id(), base = lane & ~(width - 1), src = lane ^ m;
	return cuemu_exchange(v, (src >= base + width || src < base) ? lane : src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
	cuemu::WarpSync& w = cuemu::g_block->warps[cuemu::warp_id()];
	w.slot[cuemu::lane_id()] = pred ? 1 : 0;
	cuemu::warp_barrier();
	unsigned m = 0;
	for (int i = 0; i < cuemu::warp_width(); ++i) if (w.slot[i]) m |= 1u << i;
	cuemu::warp_barrier();
	return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
	unsigned full = cuemu::warp_width() == 32 ? 0xffffffffu : ((1u << cuemu::warp_width()) - 1);
	return __ballot_sync(m, pred) == full;
}

/* ---- atomics (fibers are cooperative: plain read-modify-write is atomic) -- */
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

/* ---- scalar + SIMD-in-word intrinsics ------------------------------------- */
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
	unsigned long long v = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) {
		unsigned sel = (s >> (4 * i)) & 0xf;
		unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
		if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
using std::max;
using std::min;

static inline int16_t emu_lo(unsigned a) { return (int16_t)(a & 0xffff); }
static inline int16_t emu_hi(unsigned a) { return (int16_t)(a >> 16); }
static inline unsigned emu_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
/* wrap-around 16-bit lane arithmetic, exactly like the hardware's .16x2 forms */
static inline int emu_w16(int v) { return (int16_t)(uint16_t)v; }

static inline unsigned __vadd2(unsigned a, unsigned b) { return emu_pack(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return emu_pack(emu_lo(a) - emu_lo(b), emu_hi(a) - emu_hi(b)); }
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return emu_pack(max<int>(emu_lo(a), emu_lo(b)), max<int>(emu_hi(a), emu_hi(b))); }
static inline unsigned __vmins2(unsigned a, unsigned b) { return emu_pack(min<int>(emu_lo(a), emu_lo(b)), min<int>(emu_hi(a), emu_hi(b))); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) {
	return emu_pack(max<int>(max<int>(emu_lo(a), emu_lo(b)), 0), max<int>(max<int>(emu_hi(a), emu_hi(b)), 0));
}
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimax3_s16x2_relu(unsigned a, unsigned b, unsigned c) { return __vimax_s16x2_relu(__vmaxs2(a, b), c); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) {
	return emu_pack(max<int>(emu_w16(emu_lo(a) + emu_lo(b)), emu_lo(c)), max<int>(emu_w16(emu_hi(a) + emu_hi(b)), emu_hi(c)));
}
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c) {
	return emu_pack(max<int>(max<int>(emu_w16(emu_lo(a) + emu_lo(b)), emu_lo(c)), 0),
	                max<int>(max<int>(emu_w16(emu_hi(a) + emu_hi(b)), emu_hi(c)), 0));
}
static inline int __viaddmax_s32(int a, int b, int c) { return max(a + b, c); }
static inline int __viaddmin_s32(int a, int b, int c) { return min(a + b, c); }
static inline int __vimax3_s32(int a, int b, int c) { return max(max(a, b), c); }
static inline int __vimax3_s32_relu(int a, int b, int c) { return max(max(max(a, b), c), 0); }
static inline int __viaddmax_s32_relu(int a, int b, int c) { return max(max(a + b, c), 0); }

/* ---- runtime API subset ---------------------------------------------------- */
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost, cudaMemcpyDefault };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; size_t sharedMemPerBlockOptin; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
/* SSW_EMU_DEVICES=N: report N identical devices (the host logic of device groups; all of them are this CPU) */
static inline cudaError_t cudaGetDeviceCount(int* n) { const char* v = getenv("SSW_EMU_DEVICES"); *n = v && atoi(v) > 0 ? atoi(v) : 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
	memset(p, 0, sizeof(*p)); strcpy(p->name, "cuda_emu (CPU fibers, tests only)");
	p->multiProcessorCount = 4; p->totalGlobalMem = (size_t)8 << 30; p->major = 10; p->minor = 0;
	p->sharedMemPerBlockOptin = 227 * 1024; return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return cudaSuccess; }
static const unsigned cudaEventDisableTiming = 2;
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
#define cudaStreamNonBlocking 1
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 4; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)4 << 30; *t = (size_t)8 << 30; return cudaSuccess; }

/* kernel launch: the product's ssw_launch() (csrc/ssw_common.cuh) calls cuemu::run_grid in the emulated build */

#endif /* SSW_CUDA_EMU_H */
