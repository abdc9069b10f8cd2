// simulate_common.cuh -- device helpers shared by the simulate+emit kernels (simulate_direct.cu,
// simulate_warp.cu): 2-bit genome -> ASCII expansion, the general (N / indel segment) chunk
// generator, quality re-alignment, descriptor block loads, read-name pieces.
#pragma once
#include <stdint.h>
#include "kernels.cuh"
#include "lecuyer.h"

namespace qt {

enum : uint32_t { RF_FW = 1, RF_RANDOM = 2, RF_GENERAL = 4 };

struct GenomeView {
	const uint32_t *bases, *nmask;
	uint64_t max_word;  // last valid index of bases[] minus 1
};

// bit i of m (i < 4) -> byte i = 0xff
__device__ __forceinline__ uint32_t spread4(uint32_t m) {
	return (((m & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}

// bytes [l, h) of a 16-byte chunk, l and h clamped to [0, 16]; lowmask[n] = bytes [0, n)
__device__ __forceinline__ uint4 range_mask(const uint4 *lowmask, int l, int h) {
	l = l < 0 ? 0 : (l > 16 ? 16 : l);
	h = h < 0 ? 0 : (h > 16 ? 16 : h);
	const uint4 a = lowmask[h], b = lowmask[l];
	return make_uint4(a.x & ~b.x, a.y & ~b.y, a.z & ~b.z, a.w & ~b.w);
}

// 16 two-bit codes -> 16 ASCII bases through the 4-byte table T ("ACGT", or "TGCA" to complement).
// A nibble of w holds codes 2k (bits 0-1) and 2k+1 (bits 2-3); masked to 2 bits, every nibble is
// directly a PRMT selector into T.  Even and odd codes are looked up separately (4 per PRMT),
// then the bytes are interleaved.
__device__ __forceinline__ uint4 expand16(uint32_t w, uint32_t T) {
	const uint32_t ev = w & 0x33333333u, od = (w >> 2) & 0x33333333u;
	const uint32_t E0 = __byte_perm(T, 0u, ev), O0 = __byte_perm(T, 0u, od);
	const uint32_t E1 = __byte_perm(T, 0u, ev >> 16), O1 = __byte_perm(T, 0u, od >> 16);
	return make_uint4(__byte_perm(E0, O0, 0x5140u), __byte_perm(E0, O0, 0x7362u), __byte_perm(E1, O1, 0x5140u),
	                  __byte_perm(E1, O1, 0x7362u));
}
// The same for a mate on either strand: '-' mates print the complement of the bases in reverse order
// (asc2dnacomp, simplesim.cpp:76-97).  Complement = the table "TGCA"; reversal costs one rotate: with the
// two 16-bit halves of w swapped, the byte-interleaving PRMTs pick their bytes in descending order
// (selectors 0x5140 / 0x7362 xor 0x7777 = 0x2637 / 0x0415).
__device__ __forceinline__ uint4 expand16_strand(uint32_t w, bool fw) {
	const uint32_t T = fw ? 0x54474341u : 0x41434754u;
	const uint32_t x7 = fw ? 0u : 0x7777u;
	w = __funnelshift_l(w, w, fw ? 0u : 16u);
	const uint32_t ev = w & 0x33333333u, od = (w >> 2) & 0x33333333u;
	const uint32_t E0 = __byte_perm(T, 0u, ev), O0 = __byte_perm(T, 0u, od);
	const uint32_t E1 = __byte_perm(T, 0u, ev >> 16), O1 = __byte_perm(T, 0u, od >> 16);
	const uint32_t sa = 0x5140u ^ x7, sb = 0x7362u ^ x7;
	return make_uint4(__byte_perm(E0, O0, sa), __byte_perm(E0, O0, sb), __byte_perm(E1, O1, sa), __byte_perm(E1, O1, sb));
}
static constexpr uint32_t TBL_FW = 0x54474341u;  // "ACGT"
static constexpr uint32_t TBL_RC = 0x41434754u;  // "TGCA": complement (asc2dnacomp, simplesim.cpp:76-97)

// reverse the order of 16 two-bit codes
__device__ __forceinline__ uint32_t rev16(uint32_t w) {
	w = __brev(w);
	return ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
}

// ---- general path (N, past-the-record, indel segments): everything from global memory ----
__device__ __forceinline__ uint32_t load16_bases(const GenomeView &gv, uint64_t g) {
	uint64_t wi = g >> 4;
	wi = wi < gv.max_word ? wi : gv.max_word;
	const uint32_t lo = __ldg(&gv.bases[wi]), hi = __ldg(&gv.bases[wi + 1]);
	return __funnelshift_r(lo, hi, 2u * (uint32_t)(g & 15));
}
__device__ __forceinline__ uint32_t load16_nmask(const GenomeView &gv, uint64_t g) {
	uint64_t mi = g >> 5;
	const uint64_t mmax = gv.max_word >> 1;
	mi = mi < mmax ? mi : mmax;
	const uint32_t mlo = __ldg(&gv.nmask[mi]), mhi = __ldg(&gv.nmask[mi + 1]);
	return __funnelshift_r(mlo, mhi, (uint32_t)(g & 31)) & 0xffffu;
}

// output positions [k0, k0+16) under one segment's mapping (d = k0 - segment's out_start), with the
// N mask applied and positions past the record end printed as N
__device__ __forceinline__ uint4 seq16_masked(const GenomeView &gv, bool fw, uint64_t gpos, uint64_t gend, int32_t ref0,
                                              int d) {
	uint32_t w, nm;
	if (fw) {
		const uint64_t g0 = (uint64_t)((int64_t)gpos + ref0 + d);
		w = load16_bases(gv, g0);
		nm = load16_nmask(gv, g0);
		if (g0 + 16 > gend) nm |= g0 >= gend ? 0xffffu : (0xffffu << (uint32_t)(gend - g0)) & 0xffffu;
	} else {
		const uint64_t ghi = (uint64_t)((int64_t)gpos + ref0 - d);
		w = rev16(load16_bases(gv, ghi - 15));
		nm = __brev(load16_nmask(gv, ghi - 15)) >> 16;
		if (ghi >= gend) nm |= (ghi - gend) >= 15 ? 0xffffu : (1u << ((uint32_t)(ghi - gend) + 1u)) - 1u;
	}
	uint4 v = expand16(w, fw ? TBL_FW : TBL_RC);
	if (nm) {
		const uint32_t N4 = 0x4e4e4e4eu;
		uint32_t m;
		m = spread4(nm);       v.x = (v.x & ~m) | (N4 & m);
		m = spread4(nm >> 4);  v.y = (v.y & ~m) | (N4 & m);
		m = spread4(nm >> 8);  v.z = (v.z & ~m) | (N4 & m);
		m = spread4(nm >> 12); v.w = (v.w & ~m) | (N4 & m);
	}
	return v;
}

// merges the contributions of every reference-copy segment that overlaps output positions
// [k0, k0+16); other bytes are zero
static __device__ __noinline__ uint4 seq16_general(const GenomeView &gv, const uint4 *lowmask, const SegDev *__restrict__ segs,
                                            int nseg, bool fw, int L, uint64_t gpos, uint64_t gend, int k0) {
	const int lo_k = k0 > 0 ? k0 : 0;
	const int hi_k = k0 + 16 < L ? k0 + 16 : L;
	uint4 out = make_uint4(0u, 0u, 0u, 0u);
	int a = 0, b = nseg;  // first segment whose end lies beyond lo_k
	while (a < b) {
		const int mid = (a + b) >> 1;
		if ((int)segs[mid].out_end <= lo_k) a = mid + 1;
		else b = mid;
	}
	for (int si = a; si < nseg; si++) {
		const SegDev sg = segs[si];
		if ((int)sg.out_start >= hi_k) break;
		const uint4 v = seq16_masked(gv, fw, gpos, gend, sg.ref0, k0 - (int)sg.out_start);
		const uint4 m = range_mask(lowmask, ((int)sg.out_start > lo_k ? (int)sg.out_start : lo_k) - k0,
		                           ((int)sg.out_end < hi_k ? (int)sg.out_end : hi_k) - k0);
		out.x |= v.x & m.x;
		out.y |= v.y & m.y;
		out.z |= v.z & m.z;
		out.w |= v.w & m.w;
	}
	return out;
}

// ---- fast path: chunk slot j of a read whose bases are all ACGT, inside the record, one segment ----
// geometry (step A): '+' mates  gadj = gpos + ref0 - (P & 15);       slot j starts at base gadj + 16 j
//                    '-' mates  gadj = gpos + ref0 + (P & 15) - 15;  slot j starts at base gadj - 16 j
// w0 = gadj >> 4, phase = gadj & 15 (meta bits 8..11)
struct SeqLoad {
	uint32_t lo, hi;
};
__device__ __forceinline__ SeqLoad seq_issue(const uint32_t *__restrict__ bases, uint32_t w0, uint32_t meta, int j) {
	const uint32_t wi = (meta & RF_FW) ? w0 + (uint32_t)j : w0 - (uint32_t)j;
	SeqLoad sl;
	sl.lo = __ldg(&bases[wi]);
	sl.hi = __ldg(&bases[wi + 1]);
	return sl;
}
__device__ __forceinline__ uint4 seq_finish(const SeqLoad &sl, uint32_t meta) {
	const bool fw = (meta & RF_FW) != 0;
	return expand16_strand(__funnelshift_r(sl.lo, sl.hi, (meta >> 7) & 30u), fw);
}

// quality chunk: two aligned 16-byte loads of the pre-oriented string ...
struct QualLoad {
	uint4 x, y;
};
__device__ __forceinline__ QualLoad qual_issue(const char *__restrict__ qual_out, uint32_t unit) {
	const uint4 *bp = (const uint4 *)qual_out + unit;
	QualLoad ql;
	ql.x = __ldg(bp);
	ql.y = __ldg(bp + 1);
	return ql;
}
// ... then a branch-free barrel shift by a bytes: two word-select stages, one funnel stage
__device__ __forceinline__ uint4 qual_finish(const QualLoad &ql, uint32_t a) {
	const uint4 &x = ql.x, &y = ql.y;
	const bool s2 = (a & 8u) != 0, s1 = (a & 4u) != 0;
	const uint32_t t0 = s2 ? x.z : x.x, t1 = s2 ? x.w : x.y, t2 = s2 ? y.x : x.z, t3 = s2 ? y.y : x.w, t4 = s2 ? y.z : y.x,
	               t5 = s2 ? y.w : y.y;
	const uint32_t u0 = s1 ? t1 : t0, u1 = s1 ? t2 : t1, u2 = s1 ? t3 : t2, u3 = s1 ? t4 : t3, u4 = s1 ? t5 : t4;
	const uint32_t sh = 8u * (a & 3u);
	return make_uint4(__funnelshift_r(u0, u1, sh), __funnelshift_r(u1, u2, sh), __funnelshift_r(u2, u3, sh),
	                  __funnelshift_r(u3, u4, sh));
}

// The same with three 8-byte loads: the bytes wanted start a (0..15) bytes into 16-byte unit `unit`; loading
// from the 8-byte boundary below them leaves one word-select stage instead of two (ALU-pipe instructions
// are what bounds the chunk loops).
struct QualLoad8 {
	uint2 p, q, r;
};
__device__ __forceinline__ QualLoad8 qual_issue8(const char *__restrict__ qual_out, uint32_t unit, uint32_t a) {
	const uint2 *bp = (const uint2 *)(qual_out + 16ull * unit + (a & 8u));
	QualLoad8 ql;
	ql.p = __ldg(bp);
	ql.q = __ldg(bp + 1);
	ql.r = __ldg(bp + 2);
	return ql;
}
__device__ __forceinline__ uint4 qual_finish8(const QualLoad8 &ql, uint32_t a) {
	const bool s1 = (a & 4u) != 0;
	const uint32_t u0 = s1 ? ql.p.y : ql.p.x, u1 = s1 ? ql.q.x : ql.p.y, u2 = s1 ? ql.q.y : ql.q.x, u3 = s1 ? ql.r.x : ql.q.y,
	               u4 = s1 ? ql.r.y : ql.r.x;
	const uint32_t sh = 8u * (a & 3u);
	return make_uint4(__funnelshift_r(u0, u1, sh), __funnelshift_r(u1, u2, sh), __funnelshift_r(u2, u3, sh),
	                  __funnelshift_r(u3, u4, sh));
}

// ---- template / record descriptors as 16-byte blocks (common.h lays them out for this) ----
struct MateHot {
	uint32_t qual_out_off, ops_off, seg_off;
	int32_t ref0;
	uint32_t len, rflen, nops, nseg;
	bool fw, single;
	uint32_t score_len;
	uint32_t score[3];
};
__device__ __forceinline__ MateHot mate_from_blocks(const uint4 &a, const uint4 &b, const uint4 &c);
__device__ __forceinline__ MateHot load_mate(const MateDesc *m) {
	const uint4 *p = (const uint4 *)m;
	const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
	return mate_from_blocks(a, b, c);
}
__device__ __forceinline__ MateHot mate_from_blocks(const uint4 &a, const uint4 &b, const uint4 &c) {
	MateHot h;
	h.qual_out_off = a.x;
	h.ops_off = a.y;
	h.seg_off = a.z;
	h.ref0 = (int32_t)a.w;
	h.len = b.x;
	h.rflen = b.y;
	h.nops = b.z & 0xffffu;
	h.nseg = b.z >> 16;
	h.fw = (b.w & 0xffu) != 0;
	h.single = ((b.w >> 8) & 0xffu) != 0;
	h.score_len = (b.w >> 16) & 0xffu;
	h.score[0] = c.x;
	h.score[1] = c.y;
	h.score[2] = c.z;
	return h;
}
struct TmplTail {
	uint32_t delta[2], opp_len, typ, typ_len;
	bool aligned_is_mate1;
};
__device__ __forceinline__ TmplTail tail_from_block(const uint4 &a);
__device__ __forceinline__ TmplTail load_tail(const TmplDesc *t) { return tail_from_block(__ldg((const uint4 *)t + 8)); }
__device__ __forceinline__ TmplTail tail_from_block(const uint4 &a) {
	TmplTail h;
	h.delta[0] = a.x;
	h.delta[1] = a.y;
	h.opp_len = a.z;
	h.typ = a.w & 0xffffu;
	h.typ_len = (a.w >> 16) & 0xffu;
	h.aligned_is_mate1 = (a.w >> 24) != 0;
	return h;
}

// ---- names ----
__device__ __forceinline__ uint32_t ndig32(uint32_t v) {
	return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u
	     : v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}
__device__ __forceinline__ uint32_t ndig64(uint64_t v) {
	if (v <= 0xffffffffull) return ndig32((uint32_t)v);
	uint32_t d = 1;
	while (v >= 10) {
		v /= 10;
		d++;
	}
	return d;
}
// decimal of v, n digits (n = number of digits of v)
__device__ __forceinline__ char *s_put_dec(char *p, uint64_t v, uint32_t n) {
	if (v <= 0xffffffffull) {
		uint32_t x = (uint32_t)v;
		for (int i = (int)n - 1; i >= 0; i--) {
			const uint32_t q = x / 10u;
			p[i] = (char)('0' + (int)(x - q * 10u));
			x = q;
		}
	} else {
		for (int i = (int)n - 1; i >= 0; i--) {
			p[i] = (char)('0' + (int)(v % 10));
			v /= 10;
		}
	}
	return p + n;
}
// decimal of v, n digits, also written d2 bytes further when TWO (the same name goes to both files of a pair)
template <bool TWO>
__device__ __forceinline__ char *s_put_dec2(char *p, ptrdiff_t d2, uint64_t v, uint32_t n) {
	if (v <= 0xffffffffull) {
		uint32_t x = (uint32_t)v;
		for (int i = (int)n - 1; i >= 0; i--) {
			const uint32_t q = x / 10u;
			const char ch = (char)('0' + (int)(x - q * 10u));
			p[i] = ch;
			if (TWO) p[i + d2] = ch;
			x = q;
		}
	} else {
		for (int i = (int)n - 1; i >= 0; i--) {
			const char ch = (char)('0' + (int)(v % 10));
			p[i] = ch;
			if (TWO) p[i + d2] = ch;
			v /= 10;
		}
	}
	return p + n;
}
// "<refid>:<+|->:<off>:<score>" (simplesim.cpp:104-110); the score and refids of up to 16 bytes come
// from registers (n16 = RecDev::name16).  TWO: every byte is also stored d2 bytes further (second file).
template <bool TWO>
__device__ __forceinline__ char *s_origin2(char *p, ptrdiff_t d2, const char *__restrict__ refid, const uint4 &n16, uint32_t rl,
                                           uint64_t off, uint32_t nd, const MateHot &m) {
	auto put = [&](char *q, char ch) {
		*q = ch;
		if (TWO) q[d2] = ch;
	};
	if (rl <= 16u) {
		const uint32_t nw[4] = {n16.x, n16.y, n16.z, n16.w};
#pragma unroll
		for (int k = 0; k < 16; k++)
			if (k < (int)rl) put(p + k, (char)(nw[k >> 2] >> (8 * (k & 3))));
	} else {
		for (uint32_t i = 0; i < rl; i++) put(p + i, refid[i]);
	}
	p += rl;
	put(p, ':');
	put(p + 1, m.fw ? '+' : '-');
	put(p + 2, ':');
	p = s_put_dec2<TWO>(p + 3, d2, off, nd);
	put(p, ':');
#pragma unroll
	for (int k = 0; k < 11; k++)
		if (k < (int)m.score_len) put(p + 1 + k, (char)(m.score[k >> 2] >> (8 * (k & 3))));
	return p + 1 + m.score_len;
}
__device__ __forceinline__ char *s_origin(char *p, const char *__restrict__ refid, const uint4 &n16, uint32_t rl, uint64_t off,
                                          uint32_t nd, const MateHot &m) {
	return s_origin2<false>(p, 0, refid, n16, rl, off, nd, m);
}

// state k steps ahead, from the tables of a^j mod m (j < JUMP_TBL) in global memory
__device__ __forceinline__ LState ljump_tbl(LState s, uint32_t k, const uint32_t *jt1, const uint32_t *jt2) {
	while (k >= (uint32_t)JUMP_TBL) {
		s.s1 = mulmod1(s.s1, __ldg(&jt1[JUMP_TBL - 1]));
		s.s2 = mulmod2(s.s2, __ldg(&jt2[JUMP_TBL - 1]));
		k -= (uint32_t)(JUMP_TBL - 1);
	}
	if (k) {
		s.s1 = mulmod1(s.s1, __ldg(&jt1[k]));
		s.s2 = mulmod2(s.s2, __ldg(&jt2[k]));
	}
	return s;
}

// ASCII base as staged -> 2-bit code of the REFERENCE base (4 for N); '-' mates are staged complemented
__device__ __forceinline__ uint32_t ref_code_of(uint32_t ch, bool fw) {
	if (ch == 'N') return 4u;
	uint32_t c = (ch >> 1) & 3u;  // A 0, C 1, T 2, G 3
	c ^= c >> 1;                  // A 0, C 1, G 2, T 3
	return fw ? c : 3u - c;
}

}  // namespace qt
