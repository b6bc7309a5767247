// Host build of the synthetic-corpus generator (zarc_b200/csrc/corpus.cuh: counter-based, integer-only, the same
// code the device kernel runs) for bench.py's CPU reference arm, which must not load the product library.
// Workload generation only.  Built by zarc_b200/build.py:build_corpus_host -> tools/libzarc_corpus.so.
#include "corpus.cuh"

extern "C" size_t zc_generate_host(uint8_t* out, const uint64_t* seg_off, const uint32_t* seg_len, const uint8_t* seg_kind,
                                   const uint64_t* seg_key, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) zg_gen_segment(out + seg_off[i], seg_len[i], seg_kind[i], seg_key[i]);
	return 0;
}
