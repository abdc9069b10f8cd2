// simulate_tiled.cu -- the production simulate+emit kernel (sm_100a).
//
// Replaces SimulatedRead::init/mutate/init_random (simplesim.h:68-117, simplesim.cpp:31-71)
// and SimulatedRead::write/write_pair (simplesim.cpp:102-171) for a whole tile of reads at
// once.  The reference produces a read one byte at a time (a branch per transcript op, an
// fputc per base on the '-' strand); here the unit of work is 16 OUTPUT bytes.
//
// A CTA stages a tile of consecutive FASTQ records in shared memory at the same 16-byte
// phase they have in the output file, then ships the tile with one bulk asynchronous copy.
//
//   phase A  one thread per read: plan record + template -> per-read geometry in shared
//            memory (where the record sits in the tile, body length, strand, genome
//            position, single-segment shortcut, "any N / past the record end" flag)
//   phase B  three independent loops, no barrier between them (every store writes exactly
//            the bytes it owns, so they cannot clobber one another):
//              names      one thread per read: "@qtip!:refid:+:off:score:typ\n" and the
//                         "\n+\n" / "\n" separators
//              sequence   one thread per 16-byte chunk, READ-major (consecutive lanes =
//                         consecutive chunks of one read => coalesced genome loads,
//                         conflict-free shared stores): 32 bits of 2-bit genome, funnel-
//                         shifted from two packed words (reversed + complemented for '-'
//                         mates), expanded to 16 ASCII bases with PRMT; N mask and indel
//                         segments only on the slow path
//              qualities  one thread per 16-byte chunk, read-major: two aligned 16-byte
//                         loads of the template's pre-oriented quality string, re-aligned
//                         to the output phase with funnel shifts
//   phase C  one thread per read: the read's random draws -- X / I / S bases patched over
//            the staged sequence (REPLAY: the reference's L'Ecuyer stream stepped from the
//            plan's state, with X redraws; PHILOX: counter-based)
//   phase D  the staged tile goes to HBM: cp.async.bulk (shared -> global) for the 16-byte
//            aligned interior, byte stores for the ragged first/last chunk of the tile
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"

namespace qt {

// Debug build (-DQT_BOUNDS, `make dbg`): accesses of the tiled kernel are range-checked;
// the first violation is recorded in a device word.
#ifdef QT_BOUNDS
__device__ unsigned long long g_qt_violation[4];
#define QT_CHK(cond, id, v0, v1)                                                          \
	do {                                                                                   \
		if (!(cond)) {                                                                     \
			if (atomicCAS(&g_qt_violation[0], 0ull, (unsigned long long)(id)) == 0ull) {   \
				g_qt_violation[1] = (unsigned long long)(v0);                              \
				g_qt_violation[2] = (unsigned long long)(v1);                              \
				g_qt_violation[3] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;  \
			}                                                                              \
		}                                                                                  \
	} while (0)
#else
#define QT_CHK(cond, id, v0, v1) do { } while (0)
#endif

#ifndef QT_TL_THREADS
#define QT_TL_THREADS 256
#endif
static constexpr int TL_THREADS = QT_TL_THREADS;
#ifndef QT_SEQ_UNROLL
#define QT_SEQ_UNROLL 2
#endif
#ifndef QT_RAG_UNROLL
#define QT_RAG_UNROLL 2
#endif
#ifndef QT_MIN_BLOCKS
#define QT_MIN_BLOCKS (768 / QT_TL_THREADS)
#endif

// Profiling build (-DQT_PHASE_CLOCKS): thread 0 of every CTA accumulates the cycles between phase
// boundaries into g_qt_phase[] (A, B, C, D, whole CTA, CTA count).
#ifdef QT_PHASE_CLOCKS
__device__ unsigned long long g_qt_phase[8];
#define QT_MARK(slot)                                                   \
	do {                                                                \
		if (threadIdx.x == 0) {                                         \
			const long long now_ = clock64();                           \
			atomicAdd(&g_qt_phase[slot], (unsigned long long)(now_ - t_mark_)); \
			t_mark_ = now_;                                             \
		}                                                               \
	} while (0)
#else
#define QT_MARK(slot) do { } while (0)
#endif
static constexpr uint32_t TL_STAGE_DEFAULT = 52 * 1024;  // dynamic shared memory per CTA (4 CTAs per SM)

enum : uint8_t { BF_FW = 1, BF_RANDOM = 2, BF_SLOW = 4, BF_SINGLE = 8 };

// per-read, per-stream body geometry kept in shared memory between the phases (32 bytes)
struct BodyInfo {
	uint64_t gpos;    // genome position of the body's leftmost reference base (aligned mates)
	uint32_t P;       // staging byte position of the first sequence byte
	uint32_t segoff;  // into DevCategory::segs
	uint32_t qoff;    // into DevCategory::qual_out
	int32_t ref0;     // BF_SINGLE: reference offset of output position 0
	uint16_t L;       // body length
	uint16_t nseg;
	uint16_t nl;      // name length
	uint8_t flags;    // BF_*
	uint8_t pad;
};

// bit i of m (i < 4) -> byte i = 0xff
__device__ __forceinline__ uint32_t spread4(uint32_t m) {
	return (((m & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}

struct GenomeView {
	const uint32_t *bases, *nmask;
	uint64_t max_word;  // last valid index of bases[] minus 1
};

__device__ __forceinline__ uint32_t fetch_code_v(const GenomeView &gv, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;
	if ((__ldg(&gv.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&gv.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// bytes [l, h) of a 16-byte chunk, l and h clamped to [0, 16]; lowmask[n] = bytes [0, n)
__device__ __forceinline__ uint4 range_mask(const uint4 *lowmask, int l, int h) {
	l = l < 0 ? 0 : (l > 16 ? 16 : l);
	h = h < 0 ? 0 : (h > 16 ? 16 : h);
	const uint4 a = lowmask[h], b = lowmask[l];
	return make_uint4(a.x & ~b.x, a.y & ~b.y, a.z & ~b.z, a.w & ~b.w);
}

// 16 two-bit codes -> 16 ASCII bases.  A nibble of w holds codes 2k (bits 0-1) and 2k+1
// (bits 2-3); masked to 2 bits, every nibble is directly a PRMT selector into the table
// "ACGT".  Look up even and odd codes separately (4 per PRMT), then interleave the bytes.
__device__ __forceinline__ uint4 expand16(uint32_t w) {
	const uint32_t T = 0x54474341u;  // "ACGT"
	const uint32_t ev = w & 0x33333333u, od = (w >> 2) & 0x33333333u;
	const uint32_t E0 = __byte_perm(T, 0u, ev), O0 = __byte_perm(T, 0u, od);
	const uint32_t E1 = __byte_perm(T, 0u, ev >> 16), O1 = __byte_perm(T, 0u, od >> 16);
	return make_uint4(__byte_perm(E0, O0, 0x5140u), __byte_perm(E0, O0, 0x7362u), __byte_perm(E1, O1, 0x5140u),
	                  __byte_perm(E1, O1, 0x7362u));
}

// reverse the order of 16 two-bit codes and complement them (asc2dnacomp, simplesim.cpp:76-97)
__device__ __forceinline__ uint32_t revcomp16(uint32_t w) {
	w = __brev(w);
	w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
	return ~w;
}

// 16 consecutive bases starting at g (2-bit codes, base j in bits 2j)
__device__ __forceinline__ uint32_t load16_bases(const GenomeView &gv, uint64_t g) {
	uint64_t wi = g >> 4;
	QT_CHK(g < (gv.max_word + 2) * 16 + (1ull << 20), 1, g, gv.max_word);
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

// Fast path: sequence bytes for output positions [k0, k0+16) of a mate whose bases are all
// ACGT and inside the record.  d = k0 - (output position whose reference offset is ref0).
__device__ __forceinline__ uint4 seq16_fast(const GenomeView &gv, bool fw, uint64_t gpos, int32_t ref0, int d) {
	if (fw) return expand16(load16_bases(gv, (uint64_t)((int64_t)gpos + ref0 + d)));
	return expand16(revcomp16(load16_bases(gv, (uint64_t)((int64_t)gpos + ref0 - d) - 15)));
}

// Slow path: also applies the N mask and treats positions past the record end as N.
__device__ __forceinline__ uint4 seq16_slow(const GenomeView &gv, bool fw, uint64_t gpos, uint64_t gend, int32_t ref0,
                                            int d) {
	uint32_t w, nm;
	if (fw) {
		const uint64_t g0 = (uint64_t)((int64_t)gpos + ref0 + d);
		w = load16_bases(gv, g0);
		nm = load16_nmask(gv, g0);
		if (g0 + 16 > gend) nm |= g0 >= gend ? 0xffffu : (0xffffu << (uint32_t)(gend - g0)) & 0xffffu;
	} else {
		const uint64_t ghi = (uint64_t)((int64_t)gpos + ref0 - d);
		w = revcomp16(load16_bases(gv, ghi - 15));
		nm = __brev(load16_nmask(gv, ghi - 15)) >> 16;
		if (ghi >= gend) nm |= (ghi - gend) >= 15 ? 0xffffu : (1u << ((uint32_t)(ghi - gend) + 1u)) - 1u;
	}
	uint4 v = expand16(w);
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

// General path: merges the contributions of every reference-copy segment that overlaps
// output positions [k0, k0+16); other bytes are zero.
__device__ __noinline__ uint4 gen_seq_segments(const GenomeView &gv, const uint4 *lowmask,
                                               const SegDev *__restrict__ segs, int nseg, bool fw, int L,
                                               uint64_t gpos, uint64_t gend, int k0) {
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
		const uint4 v = seq16_slow(gv, fw, gpos, gend, sg.ref0, k0 - (int)sg.out_start);
		const uint4 m = range_mask(lowmask, ((int)sg.out_start > lo_k ? (int)sg.out_start : lo_k) - k0,
		                           ((int)sg.out_end < hi_k ? (int)sg.out_end : hi_k) - k0);
		out.x |= v.x & m.x;
		out.y |= v.y & m.y;
		out.z |= v.z & m.z;
		out.w |= v.w & m.w;
	}
	return out;
}

// Quality bytes for output positions [k0q, k0q+16): two aligned 16-byte loads from the
// pre-oriented quality arena, shifted to the output phase.
__device__ __forceinline__ uint4 gen_qual(const char *__restrict__ qual_out, uint32_t qbytes, uint32_t qoff, int k0q) {
	const int A = (int)qoff + k0q;  // >= 1 (the arena has 16 bytes of front padding)
	QT_CHK(A >= 0 && (uint32_t)A < qbytes, 2, A, qoff);
#ifdef QT_BOUNDS
	if (!(A >= 0 && (uint32_t)A < qbytes)) return make_uint4(0, 0, 0, 0);
#endif
	const uint4 *bp = (const uint4 *)qual_out + (A >> 4);
	const uint4 x = __ldg(bp), y = __ldg(bp + 1);
	const uint32_t sh = 8u * (uint32_t)(A & 3);
	switch ((A >> 2) & 3) {
	case 0:
		return make_uint4(__funnelshift_r(x.x, x.y, sh), __funnelshift_r(x.y, x.z, sh), __funnelshift_r(x.z, x.w, sh),
		                  __funnelshift_r(x.w, y.x, sh));
	case 1:
		return make_uint4(__funnelshift_r(x.y, x.z, sh), __funnelshift_r(x.z, x.w, sh), __funnelshift_r(x.w, y.x, sh),
		                  __funnelshift_r(y.x, y.y, sh));
	case 2:
		return make_uint4(__funnelshift_r(x.z, x.w, sh), __funnelshift_r(x.w, y.x, sh), __funnelshift_r(y.x, y.y, sh),
		                  __funnelshift_r(y.y, y.z, sh));
	default:
		return make_uint4(__funnelshift_r(x.w, y.x, sh), __funnelshift_r(y.x, y.y, sh), __funnelshift_r(y.y, y.z, sh),
		                  __funnelshift_r(y.z, y.w, sh));
	}
}

// Stores bytes [lo, hi) of v at p (16-byte aligned shared memory) and never disturbs the other
// bytes, which belong to a neighbouring component another thread may be writing right now:
// fully covered words are plain stores, partially covered words are cleared and set with
// shared-memory atomics (AND / OR touch only this chunk's own bits).
__device__ __forceinline__ void store_partial(unsigned char *p, const uint4 *lowmask, const uint4 &v, int lo, int hi) {
	const uint4 m = range_mask(lowmask, lo, hi);
	uint32_t *p32 = (uint32_t *)p;
	const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
	const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for (int k = 0; k < 4; k++) {
		if (mw[k] == 0xffffffffu) {
			p32[k] = vw[k];
		} else if (mw[k]) {
			atomicAnd(&p32[k], ~mw[k]);
			atomicOr(&p32[k], vw[k] & mw[k]);
		}
	}
}

__device__ __forceinline__ uint32_t ndig(uint64_t v) {
	uint32_t d = 1;
	while (v >= 10) {
		v /= 10;
		d++;
	}
	return d;
}

__device__ __forceinline__ char *s_put_u64(char *p, uint64_t v) {
	const uint32_t n = ndig(v);
	for (int i = (int)n - 1; i >= 0; i--) {
		p[i] = (char)('0' + (int)(v % 10));
		v /= 10;
	}
	return p + n;
}

__device__ __forceinline__ char *s_put(char *p, const char *__restrict__ s, uint32_t n) {
	for (uint32_t i = 0; i < n; i++) p[i] = s[i];
	return p + n;
}

__device__ __forceinline__ char *s_origin(char *p, const char *refid, uint32_t rl, bool fw, uint64_t off,
                                          const MateDesc &m) {
	p = s_put(p, refid, rl);
	*p++ = ':';
	*p++ = fw ? '+' : '-';
	*p++ = ':';
	p = s_put_u64(p, off);
	*p++ = ':';
	p = s_put(p, m.score, m.score_len);
	return p;
}

// ---- shared-memory genome cache -------------------------------------------------------
// Phase A copies, per aligned mate, the packed bases of [gpos-16, gpos+rflen+48) into shared
// memory with cp.async (16-byte pieces, 64-base aligned), so the sequence chunks and the X
// look-ups of phase C never wait on HBM.
__device__ __forceinline__ uint64_t gcache_origin(uint64_t gpos) { return (gpos - 16) & ~(uint64_t)63; }

__device__ __forceinline__ uint32_t cache16_bases(const uint32_t *cache, uint64_t gc0, uint64_t g) {
	const uint32_t off = (uint32_t)(g - gc0);
	const uint32_t lo = cache[off >> 4], hi = cache[(off >> 4) + 1];
	return __funnelshift_r(lo, hi, 2u * (off & 15u));
}

__device__ __forceinline__ uint4 cseq16_fast(const uint32_t *cache, uint64_t gc0, bool fw, uint64_t gpos, int32_t ref0,
                                             int d) {
	if (fw) return expand16(cache16_bases(cache, gc0, (uint64_t)((int64_t)gpos + ref0 + d)));
	return expand16(revcomp16(cache16_bases(cache, gc0, (uint64_t)((int64_t)gpos + ref0 - d) - 15)));
}

__device__ __forceinline__ uint4 cseq16_slow(const GenomeView &gv, const uint32_t *cache, uint64_t gc0, bool fw,
                                             uint64_t gpos, uint64_t gend, int32_t ref0, int d) {
	uint32_t w, nm;
	if (fw) {
		const uint64_t g0 = (uint64_t)((int64_t)gpos + ref0 + d);
		w = cache16_bases(cache, gc0, g0);
		nm = load16_nmask(gv, g0);
		if (g0 + 16 > gend) nm |= g0 >= gend ? 0xffffu : (0xffffu << (uint32_t)(gend - g0)) & 0xffffu;
	} else {
		const uint64_t ghi = (uint64_t)((int64_t)gpos + ref0 - d);
		w = revcomp16(cache16_bases(cache, gc0, ghi - 15));
		nm = __brev(load16_nmask(gv, ghi - 15)) >> 16;
		if (ghi >= gend) nm |= (ghi - gend) >= 15 ? 0xffffu : (1u << ((uint32_t)(ghi - gend) + 1u)) - 1u;
	}
	uint4 v = expand16(w);
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

// General path: merges the contributions of every reference-copy segment that overlaps
// output positions [k0, k0+16); other bytes are zero.
__device__ __noinline__ uint4 cgen_seq_segments(const GenomeView &gv, const uint32_t *cache, uint64_t gc0,
                                                const uint4 *lowmask, const SegDev *__restrict__ segs, int nseg, bool fw,
                                                int L, uint64_t gpos, uint64_t gend, int k0) {
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
		const uint4 v = cseq16_slow(gv, cache, gc0, fw, gpos, gend, sg.ref0, k0 - (int)sg.out_start);
		const uint4 m = range_mask(lowmask, ((int)sg.out_start > lo_k ? (int)sg.out_start : lo_k) - k0,
		                           ((int)sg.out_end < hi_k ? (int)sg.out_end : hi_k) - k0);
		out.x |= v.x & m.x;
		out.y |= v.y & m.y;
		out.z |= v.z & m.z;
		out.w |= v.w & m.w;
	}
	return out;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
	const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// quality chunk: issue the two aligned loads first (so several chunks can be in flight) ...
struct QualLoad {
	uint4 x, y;
	int A;
};
__device__ __forceinline__ QualLoad qual_issue(const char *__restrict__ qual_out, uint32_t qoff, int k0q) {
	QualLoad ql;
	ql.A = (int)qoff + k0q;  // >= 1 (the arena has 16 bytes of front padding)
	const uint4 *bp = (const uint4 *)qual_out + (ql.A >> 4);
	ql.x = __ldg(bp);
	ql.y = __ldg(bp + 1);
	return ql;
}
// ... then shift them to the output phase
__device__ __forceinline__ uint4 qual_finish(const QualLoad &ql) {
	const uint4 &x = ql.x, &y = ql.y;
	const uint32_t sh = 8u * (uint32_t)(ql.A & 3);
	switch ((ql.A >> 2) & 3) {
	case 0:
		return make_uint4(__funnelshift_r(x.x, x.y, sh), __funnelshift_r(x.y, x.z, sh), __funnelshift_r(x.z, x.w, sh),
		                  __funnelshift_r(x.w, y.x, sh));
	case 1:
		return make_uint4(__funnelshift_r(x.y, x.z, sh), __funnelshift_r(x.z, x.w, sh), __funnelshift_r(x.w, y.x, sh),
		                  __funnelshift_r(y.x, y.y, sh));
	case 2:
		return make_uint4(__funnelshift_r(x.z, x.w, sh), __funnelshift_r(x.w, y.x, sh), __funnelshift_r(y.x, y.y, sh),
		                  __funnelshift_r(y.y, y.z, sh));
	default:
		return make_uint4(__funnelshift_r(x.w, y.x, sh), __funnelshift_r(y.x, y.y, sh), __funnelshift_r(y.y, y.z, sh),
		                  __funnelshift_r(y.z, y.w, sh));
	}
}

// 32-bit decimal helpers: reference offsets fit 32 bits in every supported genome (< 4 Gbp per record);
// 64-bit division is ~100 instructions on the GPU, so the wide path is only taken when needed.
__device__ __forceinline__ uint32_t ndig32(uint32_t v) {
	return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u
	     : v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}
__device__ __forceinline__ uint32_t ndig_fast(uint64_t v) { return v <= 0xffffffffull ? ndig32((uint32_t)v) : ndig(v); }
__device__ __forceinline__ char *s_put_dec(char *p, uint64_t v) {
	if (v > 0xffffffffull) return s_put_u64(p, v);
	uint32_t x = (uint32_t)v;
	const uint32_t n = ndig32(x);
	for (int i = (int)n - 1; i >= 0; i--) {
		const uint32_t q = x / 10u;
		p[i] = (char)('0' + (int)(x - q * 10u));
		x = q;
	}
	return p + n;
}
__device__ __forceinline__ char *s_origin_fast(char *p, const char *refid, uint32_t rl, bool fw, uint64_t off,
                                               const MateDesc &m) {
	p = s_put(p, refid, rl);
	*p++ = ':';
	*p++ = fw ? '+' : '-';
	*p++ = ':';
	p = s_put_dec(p, off);
	*p++ = ':';
	p = s_put(p, m.score, m.score_len);
	return p;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
	const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// N bits of 16 consecutive bases starting at g, from the shared-memory copy of the N mask
__device__ __forceinline__ uint32_t cache16_nmask(const uint32_t *mcache, uint64_t gc0, uint64_t g) {
	const uint32_t off = (uint32_t)(g - gc0);
	const uint32_t lo = mcache[off >> 5], hi = mcache[(off >> 5) + 1];
	return __funnelshift_r(lo, hi, off & 31u) & 0xffffu;
}

__device__ __forceinline__ uint4 apply_n(uint4 v, uint32_t nm) {
	const uint32_t N4 = 0x4e4e4e4eu;
	uint32_t m;
	m = spread4(nm);       v.x = (v.x & ~m) | (N4 & m);
	m = spread4(nm >> 4);  v.y = (v.y & ~m) | (N4 & m);
	m = spread4(nm >> 8);  v.z = (v.z & ~m) | (N4 & m);
	m = spread4(nm >> 12); v.w = (v.w & ~m) | (N4 & m);
	return v;
}

// Sequence bytes for output positions [k0, k0+16) under one segment's mapping, everything from the
// shared-memory genome cache.  `past` = the mate may reach past its record (those bases print as N).
__device__ __forceinline__ uint4 cseq16(const uint32_t *cache, const uint32_t *mcache, uint64_t gc0, bool fw, bool past,
                                        uint64_t gpos, uint64_t gend, int32_t ref0, int d) {
	uint32_t w, nm;
	if (fw) {
		const uint64_t g0 = (uint64_t)((int64_t)gpos + ref0 + d);
		w = cache16_bases(cache, gc0, g0);
		nm = cache16_nmask(mcache, gc0, g0);
		if (past && g0 + 16 > gend) nm |= g0 >= gend ? 0xffffu : (0xffffu << (uint32_t)(gend - g0)) & 0xffffu;
	} else {
		const uint64_t ghi = (uint64_t)((int64_t)gpos + ref0 - d);
		w = revcomp16(cache16_bases(cache, gc0, ghi - 15));
		nm = __brev(cache16_nmask(mcache, gc0, ghi - 15)) >> 16;
		if (past && ghi >= gend) nm |= (ghi - gend) >= 15 ? 0xffffu : (1u << ((uint32_t)(ghi - gend) + 1u)) - 1u;
	}
	uint4 v = expand16(w);
	if (nm) v = apply_n(v, nm);
	return v;
}

// General path: merges the contributions of every reference-copy segment that overlaps
// output positions [k0, k0+16); other bytes are zero.
__device__ __noinline__ uint4 cseq16_segments(const uint32_t *cache, const uint32_t *mcache, uint64_t gc0,
                                              const uint4 *lowmask, const SegDev *__restrict__ segs, int nseg, bool fw,
                                              bool past, int L, uint64_t gpos, uint64_t gend, int k0) {
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
		const uint4 v = cseq16(cache, mcache, gc0, fw, past, gpos, gend, sg.ref0, k0 - (int)sg.out_start);
		const uint4 m = range_mask(lowmask, ((int)sg.out_start > lo_k ? (int)sg.out_start : lo_k) - k0,
		                           ((int)sg.out_end < hi_k ? (int)sg.out_end : hi_k) - k0);
		out.x |= v.x & m.x;
		out.y |= v.y & m.y;
		out.z |= v.z & m.z;
		out.w |= v.w & m.w;
	}
	return out;
}

static constexpr int OPS_PREFETCH = 6;  // draw ops per mate kept in shared memory by phase A

// what phase C needs per read and mate, so it never goes back to the template in HBM
struct DrawInfo {
	uint32_t ops_off;
	uint16_t nops;
	uint8_t fw;
	uint8_t pad;
	uint32_t ops[OPS_PREFETCH];
};

template <int MODE, int CAT>
__global__ void __launch_bounds__(TL_THREADS, QT_MIN_BLOCKS) k_simulate_tiled(SimArgs a, int R, uint32_t stage_stride, int K,
                                                                uint32_t gc_stride, uint32_t gc_mask_off, int skip) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;                     // output streams
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;   // aligned mates per read
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned char *stage = smem;                                                      // NS * stage_stride
	unsigned char *gcache = smem + (size_t)NS * stage_stride;                         // NM * R * gc_stride
	BodyInfo *s_body = (BodyInfo *)(gcache + (size_t)NM * R * gc_stride);             // [NS][R]
	DrawInfo *s_draw = (DrawInfo *)(s_body + (size_t)NS * R);                         // [NM][R]
	uint64_t *s_gend = (uint64_t *)(s_draw + (size_t)NM * R);                         // [R]
	LState *s_rng = (LState *)(s_gend + R);                                           // [R]
	uint32_t *s_opp = (uint32_t *)(s_rng + R);                                        // [R] (B: opp_len)
	__shared__ uint64_t s_G0[2], s_G1[2];
	__shared__ uint4 s_lowmask[17];

#ifdef QT_PHASE_CLOCKS
	long long t_mark_ = clock64();
	const long long t_start_ = t_mark_;
#endif
	const int tid = threadIdx.x;
	const uint64_t r0 = (uint64_t)blockIdx.x * (uint64_t)R;
	const int nr = (int)((a.nreads - r0) < (uint64_t)R ? (a.nreads - r0) : (uint64_t)R);
	const uint64_t *offs[2] = {a.off0, a.off1};
	const uint64_t *tots[2] = {a.tot0, a.tot1};
	char *outs[2] = {a.out0, a.out1};

	if (tid < NS) {
		s_G0[tid] = offs[tid][r0];
		s_G1[tid] = (r0 + nr < a.nreads) ? offs[tid][r0 + nr] : *tots[tid];
	}
	if (tid >= 32 && tid < 32 + 17) {
		const int n = tid - 32;
		uint32_t w[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int k = n - 4 * j;
			w[j] = k >= 4 ? 0xffffffffu : (k <= 0 ? 0u : (1u << (8 * k)) - 1u);
		}
		s_lowmask[n] = make_uint4(w[0], w[1], w[2], w[3]);
	}
	__syncthreads();

	// ---------------- phase A (one thread per read): geometry, genome prefetch, name ----------------
	for (int i = tid; i < nr; i += TL_THREADS) {
		const PlanRec pr = a.plan[r0 + i];
		uint64_t myoff[2];
#pragma unroll
		for (int s = 0; s < NS; s++) myoff[s] = offs[s][r0 + i];
		const TmplDesc &t = a.c.t[pr.tmpl];
		const RecDev rec = a.g.recs[pr.record];
		const uint64_t ro0 = pr.refoff + t.delta[0], ro1 = pr.refoff + t.delta[1];
		const uint64_t gend = rec.gstart + rec.len;
		uint32_t nl;
		if (CAT == CAT_U) nl = 13u + rec.name_len + ndig_fast(ro0) + t.m[0].score_len + t.typ_len;
		else if (CAT == CAT_B) nl = 18u + 2u * rec.name_len + 2u * ndig_fast(ro0) + 2u * t.m[0].score_len + t.typ_len;
		else nl = 18u + 2u * rec.name_len + ndig_fast(ro0) + ndig_fast(ro1) + t.m[0].score_len + t.m[1].score_len + t.typ_len;
		s_gend[i] = gend;
		s_rng[i] = LState{pr.s1, pr.s2};
		if (CAT == CAT_B) s_opp[i] = t.opp_len;
		// genome prefetch (bases + N mask), per aligned mate; draw ops for phase C
		uint8_t pastf[2] = {0, 0};
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const MateDesc &md = t.m[m];
			const uint64_t gpos = rec.gstart + (m == 0 ? ro0 : ro1);
			const uint64_t gc0 = gcache_origin(gpos);
			const uint32_t nch = (uint32_t)((gpos + md.rflen + 48 - gc0 + 63) >> 6);
			unsigned char *dst = gcache + ((size_t)m * R + i) * gc_stride;
			const unsigned char *src = (const unsigned char *)a.g.bases + (gc0 >> 2);
			const unsigned char *msrc = (const unsigned char *)a.g.nmask + (gc0 >> 3);
			const uint64_t src_limit = a.genome_words * 4;
			QT_CHK(nch * 16u <= gc_mask_off && gc_mask_off + nch * 8u <= gc_stride, 9, nch, gc_stride);
			for (uint32_t c = 0; c < nch; c++) {
				if ((gc0 >> 2) + 16ull * c + 16ull <= src_limit) {
					cp_async16(dst + 16 * c, src + 16 * c);
					cp_async8(dst + gc_mask_off + 8 * c, msrc + 8 * c);
				}
			}
			pastf[m] = (gpos + md.rflen + 16 > gend) ? BF_SLOW : 0;  // may reach past the record end
			DrawInfo di;
			di.ops_off = md.ops_off;
			di.nops = md.nops;
			di.fw = md.fw;
			di.pad = 0;
#pragma unroll
			for (int j = 0; j < OPS_PREFETCH; j++) di.ops[j] = j < (int)md.nops ? __ldg(&a.c.ops[md.ops_off + j]) : 0u;
			s_draw[m * R + i] = di;
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		char *rec_ptr[2] = {nullptr, nullptr};
		uint32_t blen[2] = {0, 0};
#pragma unroll
		for (int s = 0; s < NS; s++) {
			int mate = s;
			bool random = false;
			if (CAT == CAT_B) {
				mate = 0;
				random = (s == 0) != (t.aligned_is_mate1 != 0);
			}
			const MateDesc &md = t.m[mate];
			BodyInfo bi;
			bi.gpos = rec.gstart + (mate == 0 ? ro0 : ro1);
			bi.P = (uint32_t)(s_G0[s] & 15) + (uint32_t)(myoff[s] - s_G0[s]) + nl;
			bi.segoff = md.seg_off;
			bi.qoff = md.qual_out_off;
			bi.L = (uint16_t)(random ? t.opp_len : md.len);
			bi.nseg = md.nseg;
			bi.nl = (uint16_t)nl;
			bi.pad = 0;
			bi.ref0 = md.ref0;
			bi.flags = (md.fw ? BF_FW : 0) | (random ? BF_RANDOM : 0) | (md.single ? BF_SINGLE : 0) | pastf[mate];
			s_body[s * R + i] = bi;
			rec_ptr[s] = (char *)stage + (size_t)s * stage_stride + bi.P - nl;
			blen[s] = bi.L;
			QT_CHK(bi.P >= nl && bi.P + 2u * bi.L + 4u <= stage_stride, 5, bi.P, bi.L);
		}
		if (!(skip & 1)) {
			// name (simplesim.cpp:104-110, :140-150) and separators
			const char *refid = a.g.names + rec.name_off;
			char *p = s_put(rec_ptr[0], "@qtip!:", 7);
			p = s_origin_fast(p, refid, rec.name_len, t.m[0].fw != 0, ro0, t.m[0]);
			if (CAT == CAT_B) {
				*p++ = ':';
				p = s_origin_fast(p, refid, rec.name_len, t.m[0].fw != 0, ro0, t.m[0]);
			} else if (CAT != CAT_U) {
				*p++ = ':';
				p = s_origin_fast(p, refid, rec.name_len, t.m[1].fw != 0, ro1, t.m[1]);
			}
			*p++ = ':';
			p = s_put(p, t.typ, t.typ_len);
			*p++ = '\n';
#pragma unroll
			for (int s = 0; s < NS; s++) {
				char *sq = rec_ptr[s] + nl;
				if (s == 1)
					for (uint32_t j = 0; j < nl; j++) rec_ptr[1][j] = rec_ptr[0][j];
				sq[blen[s]] = '\n';
				sq[blen[s] + 1] = '+';
				sq[blen[s] + 2] = '\n';
				sq[2 * blen[s] + 3] = '\n';
			}
		}
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncthreads();
	QT_MARK(0);

	// ---------------- phase B: sequence and quality chunks, read-major, exact-byte stores ----------------
	{
		const int nitems = K * nr;
#pragma unroll
		for (int s = 0; s < NS; s++) {
			unsigned char *stg = stage + (size_t)s * stage_stride;
			const BodyInfo *body = s_body + s * R;
			const int cm = (CAT == CAT_C || CAT == CAT_D) ? s : 0;
			// (i, j) = (read, chunk) of item q = i*K + j, advanced without a division
			const int di = TL_THREADS / K, dj = TL_THREADS - di * K;
			// one sequence chunk (whole or ragged) of read i
			auto seq_chunk = [&](const BodyInfo &bi, int i, int k0) -> uint4 {
				const uint32_t *cache = (const uint32_t *)(gcache + ((size_t)cm * R + i) * gc_stride);
				const uint32_t *mcache = (const uint32_t *)((const unsigned char *)cache + gc_mask_off);
				const uint64_t gc0 = gcache_origin(bi.gpos);
				const bool fw = (bi.flags & BF_FW) != 0, past = (bi.flags & BF_SLOW) != 0;
				if (bi.flags & BF_SINGLE) return cseq16(cache, mcache, gc0, fw, past, bi.gpos, s_gend[i], bi.ref0, k0);
				return cseq16_segments(cache, mcache, gc0, s_lowmask, a.c.segs + bi.segoff, bi.nseg, fw, past, bi.L, bi.gpos,
				                       s_gend[i], k0);
			};
			// sequence: whole chunks only, two per iteration
			{
				int i = tid / K, j = tid - i * K;
				for (int q = tid; q < nitems && !(skip & 2); q += QT_SEQ_UNROLL * TL_THREADS) {
					uint4 v[QT_SEQ_UNROLL];
					int cc[QT_SEQ_UNROLL];
					bool live[QT_SEQ_UNROLL];
#pragma unroll
					for (int u = 0; u < QT_SEQ_UNROLL; u++) {
						live[u] = false;
						if (q + u * TL_THREADS < nitems) {
							const BodyInfo bi = body[i];
							cc[u] = (int)(bi.P >> 4) + j;
							const int k0 = 16 * cc[u] - (int)bi.P;  // output position of the chunk's first byte
							live[u] = !(bi.flags & BF_RANDOM) && k0 >= 0 && k0 + 16 <= (int)bi.L;
							if (live[u]) v[u] = seq_chunk(bi, i, k0);
						}
						i += di;
						j += dj;
						if (j >= K) {
							j -= K;
							i++;
						}
					}
#pragma unroll
					for (int u = 0; u < QT_SEQ_UNROLL; u++)
						if (live[u]) *(uint4 *)(stg + 16 * (size_t)cc[u]) = v[u];
				}
			}
			// qualities: whole chunks only, two per iteration with the loads issued before either is consumed
			{
				int i0 = tid / K, j0 = tid - i0 * K;
				for (int q = tid; q < nitems && !(skip & 4); q += 2 * TL_THREADS) {
					QualLoad ql[2];
					int cc[2];
					bool live[2], rnd[2];
#pragma unroll
					for (int u = 0; u < 2; u++) {
						live[u] = false;
						rnd[u] = false;
						if (q + u * TL_THREADS < nitems) {
							const BodyInfo bi = body[i0];
							const int L = bi.L;
							const uint32_t Pq = bi.P + (uint32_t)L + 3u;
							cc[u] = (int)(Pq >> 4) + j0;
							const int kq = 16 * cc[u] - (int)Pq;
							live[u] = kq >= 0 && kq + 16 <= L;
							rnd[u] = (bi.flags & BF_RANDOM) != 0;
							if (live[u] && !rnd[u]) ql[u] = qual_issue(a.c.qual_out, bi.qoff, kq);
						}
						i0 += di;
						j0 += dj;
						if (j0 >= K) {
							j0 -= K;
							i0++;
						}
					}
#pragma unroll
					for (int u = 0; u < 2; u++) {
						if (!live[u]) continue;
						QT_CHK(cc[u] >= 0 && 16u * (uint32_t)cc[u] + 16u <= stage_stride, 4, cc[u], u);
						const uint4 v = rnd[u] ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)  // 'I'
						                       : qual_finish(ql[u]);
						*(uint4 *)(stg + 16 * (size_t)cc[u]) = v;
					}
				}
			}
			// ragged ends: the first and last chunk of each sequence and each quality string (4 per read),
			// one uniform pass, exact-byte stores; two items per iteration, loads first
			for (int q = tid; q < 4 * nr && !(skip & 6); q += QT_RAG_UNROLL * TL_THREADS) {
				QualLoad ql[QT_RAG_UNROLL];
				uint4 v[QT_RAG_UNROLL];
				int cc[QT_RAG_UNROLL], kk[QT_RAG_UNROLL], LL[QT_RAG_UNROLL];
				bool live[QT_RAG_UNROLL], needq[QT_RAG_UNROLL];
#pragma unroll
				for (int u = 0; u < QT_RAG_UNROLL; u++) {
					const int qq = q + u * TL_THREADS;
					live[u] = false;
					needq[u] = false;
					if (qq >= 4 * nr) continue;
					const int i = qq >> 2, which = qq & 3;  // 0 seq head, 1 seq tail, 2 qual head, 3 qual tail
					const BodyInfo bi = body[i];
					const int L = bi.L;
					const bool isq = which >= 2;
					const uint32_t P0 = isq ? bi.P + (uint32_t)L + 3u : bi.P;
					const int c_first = (int)(P0 >> 4), c_last = (int)((P0 + (uint32_t)L - 1u) >> 4);
					const int c = (which & 1) ? c_last : c_first;
					if ((which & 1) && c_last == c_first) continue;  // single chunk: the head item handles it
					const int k0 = 16 * c - (int)P0;
					if (k0 >= 0 && k0 + 16 <= L) continue;  // whole chunk: done above
					cc[u] = c;
					kk[u] = k0;
					LL[u] = L;
					if (isq) {
						live[u] = true;
						if (bi.flags & BF_RANDOM) v[u] = make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u);
						else {
							needq[u] = true;
							ql[u] = qual_issue(a.c.qual_out, bi.qoff, k0);
						}
					} else if (!(bi.flags & BF_RANDOM)) {
						live[u] = true;
						v[u] = seq_chunk(bi, i, k0);
					}
				}
#pragma unroll
				for (int u = 0; u < QT_RAG_UNROLL; u++) {
					if (!live[u]) continue;
					if (needq[u]) v[u] = qual_finish(ql[u]);
					store_partial(stg + 16 * (size_t)cc[u], s_lowmask, v[u], -kk[u], LL[u] - kk[u]);
				}
			}
		}
	}
	__syncthreads();
	QT_MARK(1);

	// ---------------- phase C: random draws (shared memory only, except long draw-op lists) ----------------
	for (int i = tid; i < nr && !(skip & 8); i += TL_THREADS) {
		const BodyInfo b0 = s_body[i];
		const BodyInfo b1 = NS == 2 ? s_body[R + i] : b0;
		char *seq0 = (char *)stage + b0.P;
		char *seq1 = NS == 2 ? (char *)stage + stage_stride + b1.P : nullptr;
		// draws, in the reference's order: mate 1's transcript, mate 2's, then the random mate
		LState st = s_rng[i];
		const uint64_t gi = a.read_index_base + r0 + i;
		uint32_t dk = 0;
		Philox4 blk{};
		auto draw16 = [&]() -> uint32_t {  // PHILOX: 16 uniform bits for the read's next draw
			if ((dk & 7u) == 0u)
				blk = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), dk >> 3, PHILOX_DOM_BASE | (uint32_t)CAT,
				                 (uint32_t)a.philox_seed, (uint32_t)(a.philox_seed >> 32));
			const uint32_t w = blk.v[(dk >> 1) & 3];
			const uint32_t b = (dk & 1) ? (w >> 16) : (w & 0xffffu);
			dk++;
			return b;
		};
		const bool s0_random = (b0.flags & BF_RANDOM) != 0;
		const uint64_t gend = s_gend[i];
		for (int m = 0; m < NM; m++) {
			const DrawInfo &dinfo = s_draw[m * R + i];
			const BodyInfo &bm = (CAT == CAT_B) ? (s0_random ? b1 : b0) : (m == 0 ? b0 : b1);
			char *seq = (CAT == CAT_B) ? (s0_random ? seq1 : seq0) : (m == 0 ? seq0 : seq1);
			const uint32_t *ops = a.c.ops + dinfo.ops_off;
			const uint64_t gpos = bm.gpos;
			const uint64_t gc0 = gcache_origin(gpos);
			const uint32_t *cache = (const uint32_t *)(gcache + ((size_t)m * R + i) * gc_stride);
			const uint32_t *mcache = (const uint32_t *)((const unsigned char *)cache + gc_mask_off);
			const uint32_t T = dinfo.fw ? 0x54474341u : 0x41434754u;  // "ACGT" / "TGCA"
			const uint32_t nops = dinfo.nops;
			for (uint32_t j = 0; j < nops; j++) {
				const uint32_t e = j < (uint32_t)OPS_PREFETCH ? dinfo.ops[j] : __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				uint32_t code;
				if ((e & 3u) == OP_X) {
					const uint64_t gp = gpos + (e >> 17);
					const uint32_t off = (uint32_t)(gp - gc0);
					uint32_t ref = (cache[off >> 4] >> (2u * (off & 15u))) & 3u;
					if (gp >= gend || ((mcache[off >> 5] >> (off & 31u)) & 1u)) ref = 4u;
					if (MODE == QTIPSIM_RNG_REPLAY) {
						do {  // simplesim.cpp:47-49
							st = lstep(st);
							code = lbase_code(lfloat(lout(st)));
						} while (code == ref);
					} else {
						const uint32_t b = draw16();
						code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
					}
				} else {
					if (MODE == QTIPSIM_RNG_REPLAY) {
						st = lstep(st);
						code = lbase_code(lfloat(lout(st)));
					} else {
						code = draw16() >> 14;
					}
				}
				QT_CHK(outp < bm.L, 7, outp, bm.L);
				seq[outp] = (char)((T >> (8u * code)) & 0xffu);
			}
		}
		if (CAT == CAT_B) {
			// init_random (simplesim.h:89-117): opp_len draws, printed with the aligned mate's strand
			char *seq = s0_random ? seq0 : seq1;
			const uint32_t L = s_opp[i];
			const bool fw = s_draw[i].fw != 0;
			const uint32_t T = fw ? 0x54474341u : 0x41434754u;
			for (uint32_t j = 0; j < L; j++) {
				uint32_t code;
				if (MODE == QTIPSIM_RNG_REPLAY) {
					st = lstep(st);
					code = lbase_code(lfloat(lout(st)));
				} else {
					code = draw16() >> 14;
				}
				seq[fw ? j : L - 1 - j] = (char)((T >> (8u * code)) & 0xffu);
			}
		}
	}
	// make the generic-proxy shared-memory writes visible to the async (bulk copy) proxy
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	QT_MARK(2);

	// ---------------- phase D: staged tile -> HBM ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint64_t G0 = s_G0[s], G1 = s_G1[s];
		const int ph = (int)(G0 & 15);
		const int total = (int)(G1 - G0);
		const unsigned char *stg = stage + (size_t)s * stage_stride;
		char *gbase = outs[s] + (G0 - ph);  // 16-byte aligned
		// interior: whole 16-byte chunks [c_lo, c_hi)
		const int c_lo = ph ? 1 : 0;
		const int c_hi = (ph + total) >> 4;
		if (tid == s * 32 && c_hi > c_lo && !(skip & 16)) {
			const uint32_t nbytes = (uint32_t)(c_hi - c_lo) * 16u;
			QT_CHK(16u * c_hi <= stage_stride && total >= 0, 8, c_hi, total);
			const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stg + 16 * c_lo);
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase + 16 * (size_t)c_lo),
			             "r"(saddr), "r"(nbytes)
			             : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		if (tid >= 64 && tid < 64 + 32) {
			// staging positions [ph, ph+total) are this tile's; chunk 0 (if ph != 0) and chunk c_hi are partial
			const int j = tid - 64;
			const int end = ph + total;
			const int p = j < 16 ? j : 16 * c_hi + (j - 16);
			const int hi = j < 16 ? (16 * c_lo < end ? 16 * c_lo : end) : end;
			if (p >= ph && p < hi) gbase[p] = (char)stg[p];
		}
	}
	// the staging buffer must stay allocated until the bulk copies have read it
	if (tid == 0 || tid == 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
	QT_MARK(3);
#ifdef QT_PHASE_CLOCKS
	if (tid == 0) {
		atomicAdd(&g_qt_phase[4], (unsigned long long)(clock64() - t_start_));
		atomicAdd(&g_qt_phase[5], 1ull);
	}
#endif
}

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, uint32_t max_rflen, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;
	static uint32_t smem_budget = 0;
	if (!smem_budget) {
		const char *e = getenv("QTIPSIM_STAGE_KB");
		smem_budget = e ? (uint32_t)atoi(e) * 1024u : TL_STAGE_DEFAULT;
		if (smem_budget < 8 * 1024 || smem_budget > 200 * 1024) smem_budget = TL_STAGE_DEFAULT;
	}
	// genome cache per aligned mate: [gpos-16 rounded down to 64, gpos+rflen+48), in 16-byte pieces
	// (+1 piece so that the word after the last needed one exists), bases then N mask (8 bytes per piece)
	uint32_t gc_pieces = (max_rflen + 16 + 48 + 63 + 63) / 64 + 1;
	gc_pieces += gc_pieces & 1;  // even, so that gc_stride = 24 * pieces is a multiple of 16 (cp.async alignment)
	const uint32_t gc_mask_off = gc_pieces * 16;
	const uint32_t gc_stride = gc_pieces * 24;
	const uint32_t per_read = NS * max_rec_bytes + NM * gc_stride + NS * (uint32_t)sizeof(BodyInfo) +
	                          NM * (uint32_t)sizeof(DrawInfo) + 20;
	int R = (int)((smem_budget - NS * 64 - 64) / per_read);
	if (R > TL_THREADS) R = TL_THREADS;
	if (R < 4) return false;
	// a tile's bytes are at most R * max_rec_bytes; +48 for the alignment phase and the ragged tail
	const uint32_t stage_stride = ((uint32_t)R * max_rec_bytes + 48 + 15) & ~15u;
	const size_t smem = (size_t)NS * stage_stride + (size_t)NM * R * gc_stride + (size_t)NS * R * sizeof(BodyInfo) +
	                    (size_t)NM * R * sizeof(DrawInfo) + (size_t)R * 20 + 16;
	// chunks a sequence / quality string of maxL bytes can touch at any alignment
	const int K = (int)((max_body + 15u) / 16u) + 1;
	auto kern = k_simulate_tiled<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + R - 1) / R;
	// QTIPSIM_SKIP (debug / attribution only): bit mask of phases to leave out -- output is then incomplete
	static int skip = -1;
	if (skip < 0) {
		const char *e = getenv("QTIPSIM_SKIP");
		skip = e ? atoi(e) : 0;
	}
	kern<<<(unsigned)ntiles, TL_THREADS, smem, st>>>(a, R, stage_stride, K, gc_stride, gc_mask_off, skip);
	kernels_count_launch();
	return true;
}

template <int MODE>
static bool launch_mode(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, uint32_t max_rflen, cudaStream_t st) {
	switch (a.cat) {
	case CAT_U: return launch_cat<MODE, CAT_U>(a, max_rec_bytes, max_body, max_rflen, st);
	case CAT_B: return launch_cat<MODE, CAT_B>(a, max_rec_bytes, max_body, max_rflen, st);
	case CAT_C: return launch_cat<MODE, CAT_C>(a, max_rec_bytes, max_body, max_rflen, st);
	default: return launch_cat<MODE, CAT_D>(a, max_rec_bytes, max_body, max_rflen, st);
	}
}

#ifdef QT_PHASE_CLOCKS
void tiled_phase_report() {
	unsigned long long v[8];
	cudaMemcpyFromSymbol(v, g_qt_phase, sizeof(v));
	if (v[5])
		fprintf(stderr, "QT_PHASE_CLOCKS per CTA (cycles): A=%llu B=%llu C=%llu D=%llu total=%llu  (%llu CTAs)\n", v[0] / v[5],
		        v[1] / v[5], v[2] / v[5], v[3] / v[5], v[4] / v[5], v[5]);
	memset(v, 0, sizeof(v));
	cudaMemcpyToSymbol(g_qt_phase, v, sizeof(v));
}
#endif

#ifdef QT_BOUNDS
void tiled_debug_report() {
	unsigned long long v[4] = {0, 0, 0, 0};
	cudaMemcpyFromSymbol(v, g_qt_violation, sizeof(v));
	if (v[0])
		fprintf(stderr, "QT_BOUNDS violation id=%llu v0=%lld v1=%lld block=%llu thread=%llu\n", v[0], (long long)v[1],
		        (long long)v[2], v[3] >> 32, v[3] & 0xffffffffu);
	else fprintf(stderr, "QT_BOUNDS: no violation\n");
	memset(v, 0, sizeof(v));
	cudaMemcpyToSymbol(g_qt_violation, v, sizeof(v));
}
#endif

bool launch_simulate_tiled(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, uint32_t max_rflen,
                           cudaStream_t st) {
	if (!a.nreads) return true;
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) return launch_mode<QTIPSIM_RNG_REPLAY>(a, max_rec_bytes, max_body, max_rflen, st);
	return launch_mode<QTIPSIM_RNG_PHILOX>(a, max_rec_bytes, max_body, max_rflen, st);
}

}  // namespace qt

#ifdef QT_BOUNDS
extern "C" void qt_tiled_debug_report() { qt::tiled_debug_report(); }
#endif
#ifdef QT_PHASE_CLOCKS
extern "C" void qt_tiled_phase_report() { qt::tiled_phase_report(); }
#endif
