// pack_swar.h -- word-parallel classification of FASTA text bytes for K0 (pack.cu).
//
// Same classes as the reference's parser: A/C/G/T in either case are bases (dna_upper,
// fasta.cpp:18-35), whitespace is isspace() in the C locale (fasta.cpp:113: ' ' and 9..13), every
// other byte is 'N'.  Four text bytes are classified at a time inside one 32-bit word (no byte
// extraction, no branches); a 16-byte vector becomes three small bit sets:
//   codes : 2 bits per byte, A C G T = 0 1 2 3, 0 for everything else
//   nbits : 1 bit per byte, set for 'N'
//   ws    : 1 bit per byte, set for whitespace
// Host-callable too, so that `pytest -m "not gpu"` can check it byte by byte (hosttest.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define QT_SW __host__ __device__ __forceinline__
#else
#define QT_SW static inline
#endif

namespace qt {

struct VecClass {
	uint32_t codes, nbits, ws;
};

QT_SW uint32_t sw_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
	return (uint32_t)__popc(x);
#else
	return (uint32_t)__builtin_popcount(x);
#endif
}
QT_SW uint32_t sw_ctz(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
	return (uint32_t)__ffs((int)x) - 1u;
#else
	return (uint32_t)__builtin_ctz(x);
#endif
}
// bytes of "ACGT" selected by four 2-bit codes held in the low four nibbles of sel
QT_SW uint32_t sw_letters(uint32_t sel) {
#ifdef __CUDA_ARCH__
	return __byte_perm(0x54474341u, 0u, sel);
#else
	uint32_t r = 0;
	for (int k = 0; k < 4; k++) r |= ((0x54474341u >> (8 * ((sel >> (4 * k)) & 3u))) & 0xffu) << (8 * k);
	return r;
#endif
}
// bit 7 of each of four bytes -> four adjacent bits
QT_SW uint32_t sw_flags4(uint32_t hbits) { return ((hbits >> 7) * 0x01020408u) >> 24; }
QT_SW uint32_t sw_low(uint32_t n) { return n >= 32u ? 0xffffffffu : (1u << n) - 1u; }

QT_SW void classify_word(uint32_t w, uint32_t &code8, uint32_t &n4, uint32_t &ws4) {
	const uint32_t H = 0x80808080u, L7 = 0x7f7f7f7fu;
	// whitespace: byte == 0x20, or 9 <= byte <= 13.  (x & L7) + L7 carries into bit 7 exactly when the low seven bits
	// are not all zero and never out of the byte; (b | 0x80) - k keeps bit 7 exactly when (b & 0x7f) >= k, no borrow
	const uint32_t xs = w ^ 0x20202020u;
	const uint32_t not_sp = ((xs & L7) + L7) | xs;
	const uint32_t wh = w | H;
	const uint32_t ge9 = wh - 0x09090909u, ge14 = wh - 0x0e0e0e0eu;
	const uint32_t ws = (~not_sp | (ge9 & ~ge14 & ~w)) & H;
	// bases: ((b >> 1) ^ (b >> 2)) & 3 maps A C G T (either case) to 0 1 2 3; a byte is a base exactly when its
	// upper-cased value equals the letter its own code selects
	const uint32_t c2 = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
	uint32_t sel = (c2 | (c2 >> 4)) & 0x00ff00ffu;
	sel = (sel | (sel >> 8)) & 0xffffu;
	const uint32_t d = (w & 0xdfdfdfdfu) ^ sw_letters(sel);
	const uint32_t not_base = (((d & L7) + L7) | d) & H;
	const uint32_t cm = c2 & (((not_base ^ H) >> 7) * 3u);
	uint32_t x = (cm | (cm >> 6)) & 0x000f000fu;
	code8 = (x | (x >> 12)) & 0xffu;
	n4 = sw_flags4(not_base & ~ws);
	ws4 = sw_flags4(ws);
}

QT_SW VecClass classify_vec(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
	uint32_t c[4], n[4], s[4];
	classify_word(w0, c[0], n[0], s[0]);
	classify_word(w1, c[1], n[1], s[1]);
	classify_word(w2, c[2], n[2], s[2]);
	classify_word(w3, c[3], n[3], s[3]);
	VecClass r;
	r.codes = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
	r.nbits = n[0] | (n[1] << 4) | (n[2] << 8) | (n[3] << 12);
	r.ws = s[0] | (s[1] << 4) | (s[2] << 8) | (s[3] << 12);
	return r;
}

// whitespace flags of the vector only (the count pass)
QT_SW uint32_t ws_vec(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) { return classify_vec(w0, w1, w2, w3).ws; }

// Keeps bytes [a, b) of the vector (0 <= a < b <= 16) and squeezes the whitespace out: afterwards element j of
// codes / nbits is the j-th non-whitespace byte in range, everything above the returned count is zero.
QT_SW uint32_t squeeze_vec(VecClass &v, uint32_t a, uint32_t b) {
	const uint32_t n = b - a;
	uint32_t codes = (v.codes >> (2u * a)) & sw_low(2u * n);
	uint32_t nb = (v.nbits >> a) & sw_low(n);
	uint32_t ws = (v.ws >> a) & sw_low(n);
	const uint32_t cnt = n - sw_popc(ws);
	while (ws) {
		const uint32_t k = sw_ctz(ws);
		const uint32_t low = (1u << k) - 1u, low2 = (1u << (2u * k)) - 1u;
		nb = (nb & low) | ((nb >> 1) & ~low);
		codes = (codes & low2) | ((codes >> 2) & ~low2);
		ws = (ws >> 1) & ~low;
	}
	v.codes = codes;
	v.nbits = nb;
	v.ws = 0;
	return cnt;
}

}  // namespace qt
