// Synthetic corpus materialisation on the device (workload generation for bench/tests).
#include "common.h"
#include "corpus.cuh"

__global__ void __launch_bounds__(128)
k_corpus(u8* __restrict__ out, const u64* __restrict__ seg_off, const u32* __restrict__ seg_len,
         const u8* __restrict__ seg_kind, const u64* __restrict__ seg_key, u64 nseg) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nseg) return;
	zg_gen_segment(out + seg_off[i], seg_len[i], seg_kind[i], seg_key[i]);
}

size_t zg_corpus_run(cudaStream_t s, u8* out, const u64* seg_off, const u32* seg_len, const u8* seg_kind, const u64* seg_key, u64 nseg) {
	if (nseg == 0) return 0;
	ZG_LAUNCH(k_corpus, (u32)((nseg + 127) / 128), 128, 0, s, out, seg_off, seg_len, seg_kind, seg_key, nseg);
	ZG_COUNT_LAUNCH();
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}

extern "C" size_t zg_corpus_generate_host(uint8_t* out, const uint64_t* seg_off, const uint32_t* seg_len,
                                          const uint8_t* seg_kind, const uint64_t* seg_key, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) zg_gen_segment(out + seg_off[i], seg_len[i], seg_kind[i], seg_key[i]);
	return 0;
}
