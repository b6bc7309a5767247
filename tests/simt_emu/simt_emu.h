// SIMT emulator: TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Lets the CUDA kernel sources under zarc_b200/csrc/ be compiled with plain g++ and executed on
// the CPU so that their *logic* can be checked against the oracle in the `-m "not gpu"` test
// suite (this container has no GPU).  Every CUDA thread of a CTA is a fiber (hand-rolled x86-64
// context switch); CTAs run one after another on the calling OS thread.  Warp collectives and
// __syncthreads are real rendezvous points, so divergent code, shuffles, ballots and shared-memory
// hand-offs behave as on hardware.  The lane scheduling order can be changed with
// ZG_EMU_ORDER={fwd,rev,rand} to shake out missing __syncwarp()s.
//
// The product path is libzarcgpu.so (nvcc, sm_100a) and fails loudly without a GPU; nothing in
// zarc_b200/ loads this emulator.
#pragma once
#ifndef ZG_EMU
#error "simt_emu.h is only for -DZG_EMU host builds"
#endif

#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <functional>
#include <algorithm>
#include <type_traits>

// ------------------------------------------------------------------------------------------
// CUDA keywords
#define __global__
#define __device__
#define __host__
#define __constant__ static const
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace zg_emu {
struct State {
	uint3 tIdx, bIdx;
	dim3 bDim, gDim;
	unsigned char* dyn_smem;
};
extern State g;
// rendezvous of the lanes in `mask` of the current warp (all must call with the same mask)
void warp_barrier(unsigned mask);
void cta_barrier();
// exchange slots of the current warp (one 64-bit value per lane)
unsigned long long* warp_slots();
unsigned lane_id();
// hand the processor to another fiber of the CTA (a spin-wait on what another warp writes must call this)
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace zg_emu

#define threadIdx (zg_emu::g.tIdx)
#define blockIdx (zg_emu::g.bIdx)
#define blockDim (zg_emu::g.bDim)
#define gridDim (zg_emu::g.gDim)
#define warpSize 32

// ------------------------------------------------------------------------------------------
// synchronisation + warp collectives
static inline void __syncthreads() { zg_emu::cta_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { zg_emu::warp_barrier(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

namespace zg_emu {
template <typename T>
static inline unsigned long long to_bits(T v) {
	static_assert(sizeof(T) <= 8, "shuffle of >64-bit value");
	unsigned long long b = 0;
	memcpy(&b, &v, sizeof(T));
	return b;
}
template <typename T>
static inline T from_bits(unsigned long long b) {
	T v;
	memcpy(&v, &b, sizeof(T));
	return v;
}
// publish v, rendezvous, read slot `src` (or own value if src is not in mask / out of range)
template <typename T>
static inline T exchange(unsigned mask, T v, int src, bool valid) {
	unsigned long long* s = warp_slots();
	unsigned me = lane_id();
	s[me] = to_bits(v);
	warp_barrier(mask);
	T r = valid ? from_bits<T>(s[src & 31]) : v;
	warp_barrier(mask);
	return r;
}
}  // namespace zg_emu

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
	int me = (int)zg_emu::lane_id();
	int base = me & ~(width - 1);
	return zg_emu::exchange(mask, v, base + (src & (width - 1)), true);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
	int me = (int)zg_emu::lane_id();
	int base = me & ~(width - 1);
	int src = me - (int)d;
	return zg_emu::exchange(mask, v, src, src >= base);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
	int me = (int)zg_emu::lane_id();
	int base = me & ~(width - 1);
	int src = me + (int)d;
	return zg_emu::exchange(mask, v, src, src < base + width);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
	int me = (int)zg_emu::lane_id();
	(void)width;
	return zg_emu::exchange(mask, v, me ^ x, true);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
	unsigned long long* s = zg_emu::warp_slots();
	unsigned me = zg_emu::lane_id();
	s[me] = pred ? 1 : 0;
	zg_emu::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1)
			if (s[i]) r |= 1u << i;
	zg_emu::warp_barrier(mask);
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <typename T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
	unsigned long long* s = zg_emu::warp_slots();
	unsigned me = zg_emu::lane_id();
	unsigned long long mine = zg_emu::to_bits(v);
	s[me] = mine;
	zg_emu::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if (((mask >> i) & 1) && s[i] == mine) r |= 1u << i;
	zg_emu::warp_barrier(mask);
	return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
	unsigned long long* s = zg_emu::warp_slots();
	s[zg_emu::lane_id()] = v;
	zg_emu::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1) r += (unsigned)s[i];
	zg_emu::warp_barrier(mask);
	return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
	unsigned long long* s = zg_emu::warp_slots();
	s[zg_emu::lane_id()] = v;
	zg_emu::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1) r = std::max(r, (unsigned)s[i]);
	zg_emu::warp_barrier(mask);
	return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
	unsigned long long* s = zg_emu::warp_slots();
	s[zg_emu::lane_id()] = v;
	zg_emu::warp_barrier(mask);
	unsigned r = 0xffffffffu;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1) r = std::min(r, (unsigned)s[i]);
	zg_emu::warp_barrier(mask);
	return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
	unsigned long long* s = zg_emu::warp_slots();
	s[zg_emu::lane_id()] = v;
	zg_emu::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1) r |= (unsigned)s[i];
	zg_emu::warp_barrier(mask);
	return r;
}

// ------------------------------------------------------------------------------------------
// atomics (fibers never pre-empt each other, so plain read-modify-write is atomic)
template <typename T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <typename T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T>
static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T>
static inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <typename T>
static inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T>
static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ------------------------------------------------------------------------------------------
// integer intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) {
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
	unsigned long long t = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		unsigned sel = (s >> (4 * i)) & 0xf;
		unsigned byte = (unsigned)(t >> (8 * (sel & 7))) & 0xff;
		if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
	unsigned long long t = ((unsigned long long)hi << 32) | lo;
	return (unsigned)(t >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
	unsigned long long t = ((unsigned long long)hi << 32) | lo;
	return (unsigned)((t << (sh & 31)) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) {
	unsigned long long t = ((unsigned long long)hi << 32) | lo;
	sh = sh > 32 ? 32 : sh;
	return sh == 32 ? hi : (unsigned)(t >> sh);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
	return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;

// ------------------------------------------------------------------------------------------
// the slice of the CUDA runtime API the host side uses, on plain host memory
typedef int cudaError_t;
typedef struct zg_emu_stream* cudaStream_t;
typedef struct zg_emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaEventDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[256]; size_t totalGlobalMem; int major, minor; };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
	memset(p, 0, sizeof(*p));
	p->multiProcessorCount = 2;
	strcpy(p->name, "simt-emu");
	p->major = 10;
	return cudaSuccess;
}
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
double zg_emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new zg_emu_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = zg_emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
