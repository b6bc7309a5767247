// Bulk asynchronous copies global -> shared memory (the TMA unit's 1-D mode: cp.async.bulk) and the
// mbarrier objects their completion is counted on.  sm_100a inline PTX; under -DZG_EMU (CPU test build
// of the same kernel sources) a copy is a memcpy at issue time and the barrier calls do nothing.
//
// Rules of cp.async.bulk: source, destination and size are multiples of 16 bytes; completion is
// reported to an mbarrier as a byte count (complete_tx), which the issuing side announces with
// expect_tx when it arrives.  A phase of the barrier ends when every expected arrival has happened
// and the announced bytes have landed; waiters poll the phase parity.
#pragma once
#include "simt.h"

#ifdef ZG_EMU
ZG_DEV void zg_mbar_init(u64* bar, u32 count) {
	(void)count;
	*bar = 0;
}
ZG_DEV void zg_mbar_fence_init() {}
ZG_DEV void zg_mbar_arrive(u64* bar) { (void)bar; }
ZG_DEV void zg_mbar_arrive_tx(u64* bar, u32 bytes) {
	(void)bar;
	(void)bytes;
}
ZG_DEV void zg_mbar_wait(u64* bar, u32 parity) {
	(void)bar;
	(void)parity;
}
ZG_DEV void zg_bulk_g2s(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
	(void)bar;
	memcpy(smem_dst, gmem_src, bytes);
}
ZG_DEV void zg_fence_proxy_async() {}
ZG_DEV void zg_cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
ZG_DEV void zg_cp_async16_l1(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
ZG_DEV void zg_cp_async_commit() {}
template <int N>
ZG_DEV void zg_cp_async_wait() {}
ZG_DEV u32 zg_shl_clamp(u32 x, u32 s) { return s >= 32 ? 0u : x << s; }
#else
ZG_DEV u32 zg_smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
ZG_DEV void zg_mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(zg_smem_addr(bar)), "r"(count) : "memory"); }
// makes the initialised barrier visible to the asynchronous proxy (the copy unit) before the first copy names it
ZG_DEV void zg_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZG_DEV void zg_mbar_arrive(u64* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(zg_smem_addr(bar)) : "memory"); }
ZG_DEV void zg_mbar_arrive_tx(u64* bar, u32 bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zg_smem_addr(bar)), "r"(bytes) : "memory");
}
ZG_DEV void zg_mbar_wait(u64* bar, u32 parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "ZG_MBAR_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra ZG_MBAR_DONE;\n"
	    "bra ZG_MBAR_WAIT;\n"
	    "ZG_MBAR_DONE:\n"
	    "}\n" ::"r"(zg_smem_addr(bar)),
	    "r"(parity)
	    : "memory");
}
ZG_DEV void zg_bulk_g2s(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(zg_smem_addr(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(zg_smem_addr(bar))
	             : "memory");
}
// orders this thread's earlier ordinary shared-memory writes before its later bulk copies into the same bytes
ZG_DEV void zg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Per-thread asynchronous copies (cp.async, LDGSTS): 16 bytes, both addresses 16-byte aligned, L1 bypassed.  A thread
// commits its copies as a group and later waits until at most N of its most recent groups are still in flight.
ZG_DEV void zg_cp_async16(void* smem_dst, const void* gmem_src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(zg_smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
ZG_DEV void zg_cp_async16_l1(void* smem_dst, const void* gmem_src) {  // the same, allocating in L1
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(zg_smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
ZG_DEV void zg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
ZG_DEV void zg_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// shift left with the hardware's clamp: a count of 32 or more gives 0
ZG_DEV u32 zg_shl_clamp(u32 x, u32 s) {
	u32 r;
	asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
	return r;
}
#endif
