// philox.h -- Philox4x32-10 (Salmon et al., SC'11), the native counter-based generator.
//
// Native-RNG mode replaces the reference's single sequential rnglib stream
// (rnglib.cpp:573-683) with a keyed function of (read index, draw index): every read's
// randomness is addressable without walking the reads before it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QT_PHD __host__ __device__ __forceinline__
#else
#define QT_PHD inline
#endif

namespace qt {

struct Philox4 {
	uint32_t v[4];
};

// counter domains (4th counter word): what the block of random bits is used for
static constexpr uint32_t PHILOX_DOM_PLAN = 0x100u;    // template + offset of a read
static constexpr uint32_t PHILOX_DOM_BASE = 0x200u;    // base draws of a read
static constexpr uint32_t PHILOX_DOM_WINDOW = 0x300u;  // per-window binomial counts
static constexpr uint32_t PHILOX_DOM_RMATE = 0x400u;   // bases of a bad-end pair's random mate (below)

QT_PHD uint32_t ph_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

QT_PHD Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int i = 0; i < 10; i++) {
		const uint32_t hi0 = ph_mulhi(M0, c0), lo0 = M0 * c0;
		const uint32_t hi1 = ph_mulhi(M1, c2), lo1 = M1 * c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0;
		c1 = lo1;
		c2 = n2;
		c3 = lo0;
		k0 += W0;
		k1 += W1;
	}
	return Philox4{{c0, c1, c2, c3}};
}

// The random mate of a bad-end pair (init_random, simplesim.h:89-117) needs opp_len uniform bases and nothing else:
// two bits each.  Base j of the mate (in the order the reference draws them) is bits 2 (j mod 16) of word
// (j / 16) mod 4 of block j / 64 -- so the blocks of a mate, laid end to end, ARE its 2-bit packed sequence in the
// layout of the packed genome, and the simulate kernel prints it with the code that prints genome spans.
QT_PHD Philox4 philox_rmate_block(uint64_t read_index, uint32_t block, int cat, uint64_t seed) {
	return philox4x32((uint32_t)read_index, (uint32_t)(read_index >> 32), block, PHILOX_DOM_RMATE | (uint32_t)cat, (uint32_t)seed,
	                  (uint32_t)(seed >> 32));
}
QT_PHD uint32_t philox_rmate_code(const Philox4 &blk, uint32_t j) { return (blk.v[(j >> 4) & 3u] >> (2u * (j & 15u))) & 3u; }

}  // namespace qt
