// INT32 issue-rate micro-benchmarks ("an INT32 peak is not in MEASURED_PEAKS.json, so measure one:
// dependent-free iadd3 / lop3 / shf / prmt mix, all SMs").  The ceiling BLAKE3 (≈13.1 int-ops per byte) and match
// finding are held against.  Measurement tooling: not part of the content path.
#include "common.h"

#define PK_CHAINS 8
#define PK_THREADS 256

// MODE 0: IADD3   1: LOP3   2: SHF (funnel shift)   3: PRMT   4: the BLAKE3 quarter-round mix (add3 / xor / rotates by 16, 12, 8, 7)
// Every thread runs PK_CHAINS independent chains (the neighbour chain's value is the second operand, so nothing folds to
// a closed form); ops per thread = iters * PK_CHAINS * (ops per chain step, see zg_internal_int_peak).
template <int MODE>
__global__ void __launch_bounds__(PK_THREADS) k_int_peak(u32* __restrict__ out, u32 iters, u32 seed) {
	u32 a[PK_CHAINS];
	u32 t = blockIdx.x * blockDim.x + threadIdx.x;
	ZG_UNROLL
	for (int k = 0; k < PK_CHAINS; k++) a[k] = seed * (2u * (u32)k + 1u) + t * 0x9E3779B9u;
	u32 c = seed ^ t;
	ZG_UNROLL1
	for (u32 i = 0; i < iters; i++) {
		ZG_UNROLL
		for (int r = 0; r < 4; r++) {
			ZG_UNROLL
			for (int k = 0; k < PK_CHAINS; k++) {
				u32 x = a[k], y = a[(k + 1) & (PK_CHAINS - 1)];
				if (MODE == 0) x = x + y + c;
				else if (MODE == 1) asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(x), "r"(y), "r"(c));  // bit select: one 3-input op
				else if (MODE == 2) x = __funnelshift_l(x, y, 7);
				else if (MODE == 3) x = __byte_perm(x, y, 0x2541);
				else {
					// one G half-step: a += b + m; d = rotr(d ^ a, 16); c += d; b = rotr(b ^ c, 12)  (chains stand in for a,b,c,d)
					u32 z = a[(k + 2) & (PK_CHAINS - 1)], w = a[(k + 3) & (PK_CHAINS - 1)];
					x = x + y + c;                              // IADD3
					w = __byte_perm(w ^ x, 0, 0x1032);          // LOP3 + PRMT (rotate by 16)
					z = z + w;                                  // IADD
					y = __funnelshift_r(y ^ z, y ^ z, 12);      // LOP3 + SHF
					a[(k + 1) & (PK_CHAINS - 1)] = y;
					a[(k + 2) & (PK_CHAINS - 1)] = z;
					a[(k + 3) & (PK_CHAINS - 1)] = w;
				}
				a[k] = x;
			}
		}
	}
	u32 s = 0;
	ZG_UNROLL
	for (int k = 0; k < PK_CHAINS; k++) s ^= a[k];
	if (s == 0x12345678u) out[t] = s;  // (practically never: keeps the chains alive)
}

// Runs one mode on the whole device; returns device milliseconds (CUDA events on `stream`) and the number of 32-lane
// integer instructions issued per thread (so lane-ops/s = threads * per_thread / time).
extern "C" size_t zg_internal_int_peak(void* stream, int mode, uint32_t iters, uint32_t ctas_per_sm, double* ms, uint64_t* threads,
                                       uint64_t* ops_per_thread) {
#ifdef ZG_EMU
	return ZG_ERR(ZG_error_no_device);
#else
	cudaStream_t s = (cudaStream_t)stream;
	u32 grid = (u32)zg_sm_count() * (ctas_per_sm ? ctas_per_sm : 8);
	ZgBuf out;
	if (out.reserve((size_t)grid * PK_THREADS * 4)) return ZG_ERR(ZG_error_memory_allocation);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int rep = 0; rep < 2; rep++) {  // first pass warms up
		cudaEventRecord(e0, s);
		switch (mode) {
		case 0: k_int_peak<0><<<grid, PK_THREADS, 0, s>>>(out.as<u32>(), iters, 12345u); break;
		case 1: k_int_peak<1><<<grid, PK_THREADS, 0, s>>>(out.as<u32>(), iters, 12345u); break;
		case 2: k_int_peak<2><<<grid, PK_THREADS, 0, s>>>(out.as<u32>(), iters, 12345u); break;
		case 3: k_int_peak<3><<<grid, PK_THREADS, 0, s>>>(out.as<u32>(), iters, 12345u); break;
		default: k_int_peak<4><<<grid, PK_THREADS, 0, s>>>(out.as<u32>(), iters, 12345u); break;
		}
		cudaEventRecord(e1, s);
		cudaEventSynchronize(e1);
	}
	float f = 0;
	cudaEventElapsedTime(&f, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	out.release();
	if (ms) *ms = f;
	if (threads) *threads = (u64)grid * PK_THREADS;
	// per chain step: modes 0..3 issue ONE instruction (IADD3 / LOP3 / SHF / PRMT: checked with cuobjdump -sass);
	// the mix performs 7 integer operations counted the way SURVEY.md App. B counts BLAKE3's 792 per compression
	// (3 additions, 2 xors, 2 rotations; ptxas issues them as ~6.2 instructions: IMAD.IADD x2.2, LOP3 x2, PRMT, LEA.HI)
	if (ops_per_thread) *ops_per_thread = (u64)iters * 4 * PK_CHAINS * (mode >= 4 ? 7 : 1);
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
#endif
}
