// Build-mode shim.  Product build: nvcc, sm_100a, real CUDA.  -DZG_EMU: the same kernel sources
// compiled by g++ against tests/simt_emu (test infrastructure; see that header).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef ZG_EMU
#include "simt_emu.h"
#define ZG_DYN_SMEM(type, name) type* name = (type*)zg_emu::g.dyn_smem
template <typename... KArgs, typename... Args>
static inline void zg_emu_launch_k(void (*k)(KArgs...), dim3 g, dim3 b, size_t smem, Args... args) {
	zg_emu::launch(g, b, smem, [&]() { k(args...); });
}
#define ZG_LAUNCH(kernel, grid, block, smem, stream, ...) \
	zg_emu_launch_k(kernel, dim3(grid), dim3(block), (smem), __VA_ARGS__)
#define ZG_UNROLL
#define ZG_CONST_TABLE static const
#define zg_prefetch_l2(p) ((void)(p))
#define zg_prefetch_l1(p) ((void)(p))
#else
#include <cuda_runtime.h>
#define ZG_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; type* name = (type*)name##_raw_
#define ZG_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ZG_UNROLL _Pragma("unroll")
#define ZG_CONST_TABLE static __device__ const
#define zg_prefetch_l2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define zg_prefetch_l1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#endif

#define ZG_DEV __device__ __forceinline__
#ifdef ZG_EMU
#define ZG_DEV_NOINLINE static __attribute__((noinline))
#define ZG_UNROLL1
#else
#define ZG_DEV_NOINLINE static __device__ __noinline__
#define ZG_UNROLL1 _Pragma("unroll 1")
#endif
#define ZG_HD __host__ __device__ __forceinline__
#define ZG_FULL 0xffffffffu

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;

template <typename T>
ZG_HD T zg_min(T a, T b) { return a < b ? a : b; }
template <typename T>
ZG_HD T zg_max(T a, T b) { return a > b ? a : b; }

ZG_DEV u32 zg_lane() { return threadIdx.x & 31u; }
ZG_DEV u32 zg_lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// inclusive warp prefix sum
ZG_DEV u32 zg_warp_incl_scan(u32 v) {
	u32 lane = zg_lane();
	ZG_UNROLL
	for (int d = 1; d < 32; d <<= 1) {
		u32 t = __shfl_up_sync(ZG_FULL, v, d);
		if (lane >= (u32)d) v += t;
	}
	return v;
}
// warp-wide sum / maximum of one u32 per lane: the hardware's one-instruction reduction (REDUX)
ZG_DEV u32 zg_warp_sum(u32 v) { return __reduce_add_sync(ZG_FULL, v); }
ZG_DEV u32 zg_warp_max(u32 v) { return __reduce_max_sync(ZG_FULL, v); }

// unaligned little-endian loads (global or shared)
ZG_DEV u32 zg_ld16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
ZG_DEV u32 zg_ld24(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16); }
ZG_DEV u32 zg_ld32(const u8* p) {
	uintptr_t a = (uintptr_t)p;
	const u32* w = (const u32*)(a & ~(uintptr_t)3);
	u32 sh = (u32)(a & 3) * 8;
	u32 lo = w[0];
	if (sh == 0) return lo;
	u32 hi = w[1];
	return __funnelshift_r(lo, hi, sh);
}
ZG_DEV u64 zg_ld64(const u8* p) {
	uintptr_t a = (uintptr_t)p;
	const u32* w = (const u32*)(a & ~(uintptr_t)3);
	u32 sh = (u32)(a & 3) * 8;
	u32 w0 = w[0], w1 = w[1];
	if (sh == 0) return ((u64)w1 << 32) | w0;
	u32 w2 = w[2];
	return ((u64)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
