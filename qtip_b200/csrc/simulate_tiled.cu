// simulate_tiled.cu -- the production simulate+emit kernel (sm_100a).
//
// Replaces SimulatedRead::init/mutate/init_random (simplesim.h:68-117, simplesim.cpp:31-71)
// and SimulatedRead::write/write_pair (simplesim.cpp:102-171) for a whole tile of reads at
// once.  The reference produces a read one byte at a time (a branch per transcript op, an
// fputc per base on the '-' strand); here the unit of work is 16 OUTPUT bytes:
//
//   phase A  one thread per read: plan record + template -> geometry in shared memory
//            (record offset inside the tile, name length, genome position of each mate)
//   phase B  one thread per 16-byte chunk of the tile's FASTQ bytes: locate the read
//            (binary search over the tile's record offsets), then
//              sequence chunk = 32 bits of 2-bit genome (funnel-shifted from two packed
//                               words, reversed+complemented for '-' mates) expanded to 16
//                               ASCII bases with PRMT, N mask applied, indel segments merged
//              quality chunk  = 16 bytes of the template's (pre-oriented) quality string,
//                               re-aligned with funnel shifts
//            one 16-byte shared-memory store per chunk
//   phase C  one thread per read: read name (itoa), separators, and the read's random
//            draws -- X / I / S bases patched over the staged sequence (REPLAY: the
//            reference's L'Ecuyer stream stepped from the plan's state, with X redraws;
//            PHILOX: counter-based)
//   phase D  the staged tile goes to HBM as aligned 16-byte vectors (byte stores only for
//            the ragged first/last chunk of the tile)
//
// Shared-memory staging address == global address (mod 16), so phase D needs no shifting.
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"
#include <stdlib.h>

namespace qt {

static constexpr int TL_THREADS = 256;
static constexpr uint32_t TL_STAGE_DEFAULT = 44 * 1024;  // staging per CTA, all streams (4 CTAs per SM)

__device__ __forceinline__ uint32_t bytemask32(int l, int h) {
	// bytes [l, h) of a 32-bit word set to 0xff; l, h clamped to [0, 4]
	l = l < 0 ? 0 : (l > 4 ? 4 : l);
	h = h < 0 ? 0 : (h > 4 ? 4 : h);
	const uint64_t hi = (1ull << (8 * h)) - 1ull, lo = (1ull << (8 * l)) - 1ull;
	return (uint32_t)(hi & ~lo);
}

__device__ __forceinline__ uint4 bytemask128(int l, int h) {
	return make_uint4(bytemask32(l, h), bytemask32(l - 4, h - 4), bytemask32(l - 8, h - 8), bytemask32(l - 12, h - 12));
}

// bit i of m (i < 4) -> byte i = 0xff
__device__ __forceinline__ uint32_t spread4(uint32_t m) {
	return (((m & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}

struct GenomeView {
	const uint32_t *bases, *nmask;
	uint64_t max_word;  // last valid index of bases[] minus 1
};

// 16 consecutive bases starting at g (2-bit codes, base j in bits 2j) and their N bits
__device__ __forceinline__ void load16(const GenomeView &gv, uint64_t g, uint32_t &w, uint32_t &nm) {
	uint64_t wi = g >> 4;
	wi = wi < gv.max_word ? wi : gv.max_word;
	const uint32_t lo = __ldg(&gv.bases[wi]), hi = __ldg(&gv.bases[wi + 1]);
	w = __funnelshift_r(lo, hi, 2u * (uint32_t)(g & 15));
	const uint64_t mi = wi >> 1;
	const uint32_t mlo = __ldg(&gv.nmask[mi]), mhi = __ldg(&gv.nmask[mi + 1]);
	nm = __funnelshift_r(mlo, mhi, (uint32_t)(g & 31)) & 0xffffu;
}

__device__ __forceinline__ uint32_t fetch_code_v(const GenomeView &gv, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;
	if ((__ldg(&gv.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&gv.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// Sequence bytes for output positions [k0, k0+16) of one aligned mate; bytes outside the
// mate's reference-copy segments are zero.
__device__ __forceinline__ uint4 gen_seq(const GenomeView &gv, const SegDev *__restrict__ segs, int nseg, bool fw,
                                         int L, uint64_t gpos, uint64_t gend, int k0) {
	uint4 out = make_uint4(0u, 0u, 0u, 0u);
	const int lo_k = k0 > 0 ? k0 : 0;
	const int hi_k = k0 + 16 < L ? k0 + 16 : L;
	// first segment whose end lies beyond lo_k
	int a = 0, b = nseg;
	while (a < b) {
		const int mid = (a + b) >> 1;
		if ((int)segs[mid].out_end <= lo_k) a = mid + 1;
		else b = mid;
	}
	for (int si = a; si < nseg; si++) {
		const SegDev sg = segs[si];
		if ((int)sg.out_start >= hi_k) break;
		const int d = k0 - (int)sg.out_start;  // > -16
		uint32_t w, nm;
		int nbeyond_lo = 16, nbeyond_hi = 16;  // chunk positions [nbeyond_lo, nbeyond_hi) lie past the record
		if (fw) {
			const uint64_t g0 = (uint64_t)((int64_t)gpos + sg.ref0 + d);
			load16(gv, g0, w, nm);
			if (g0 + 16 > gend) {
				nbeyond_lo = g0 >= gend ? 0 : (int)(gend - g0);
				nbeyond_hi = 16;
			}
		} else {
			const uint64_t ghi = (uint64_t)((int64_t)gpos + sg.ref0 - d);
			const uint64_t g0 = ghi - 15;
			load16(gv, g0, w, nm);
			// reverse the order of the 16 bases, then complement (asc2dnacomp, simplesim.cpp:76-97)
			w = __brev(w);
			w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
			w = ~w;
			nm = __brev(nm) >> 16;
			if (ghi >= gend) {
				nbeyond_lo = 0;
				nbeyond_hi = (ghi - gend) >= 15 ? 16 : (int)(ghi - gend) + 1;
			}
		}
		if (nbeyond_lo < nbeyond_hi) nm |= ((1u << nbeyond_hi) - 1u) & ~((1u << nbeyond_lo) - 1u) & 0xffffu;
		// 16 codes -> 16 ASCII bases.  A nibble of w holds codes 2k (bits 0-1) and 2k+1 (bits 2-3);
		// masked to 2 bits, every nibble is directly a PRMT selector into the table "ACGT".
		// Look up even and odd codes separately (4 per PRMT), then interleave the bytes.
		const uint32_t T = 0x54474341u;  // "ACGT"
		const uint32_t ev = w & 0x33333333u, od = (w >> 2) & 0x33333333u;
		const uint32_t E0 = __byte_perm(T, 0u, ev), O0 = __byte_perm(T, 0u, od);
		const uint32_t E1 = __byte_perm(T, 0u, ev >> 16), O1 = __byte_perm(T, 0u, od >> 16);
		uint4 v = make_uint4(__byte_perm(E0, O0, 0x5140u), __byte_perm(E0, O0, 0x7362u), __byte_perm(E1, O1, 0x5140u),
		                     __byte_perm(E1, O1, 0x7362u));
		if (nm) {
			const uint32_t N4 = 0x4e4e4e4eu;
			uint32_t m;
			m = spread4(nm);       v.x = (v.x & ~m) | (N4 & m);
			m = spread4(nm >> 4);  v.y = (v.y & ~m) | (N4 & m);
			m = spread4(nm >> 8);  v.z = (v.z & ~m) | (N4 & m);
			m = spread4(nm >> 12); v.w = (v.w & ~m) | (N4 & m);
		}
		const int sa = ((int)sg.out_start > lo_k ? (int)sg.out_start : lo_k) - k0;
		const int sb = ((int)sg.out_end < hi_k ? (int)sg.out_end : hi_k) - k0;
		if (sa == 0 && sb == 16) return v;
		const uint4 m = bytemask128(sa, sb);
		out.x |= v.x & m.x;
		out.y |= v.y & m.y;
		out.z |= v.z & m.z;
		out.w |= v.w & m.w;
	}
	return out;
}

// Quality bytes for output positions [k0q, k0q+16), from the pre-oriented quality arena.
__device__ __forceinline__ uint4 gen_qual(const char *__restrict__ qual_out, uint32_t qoff, int L, int k0q) {
	const int64_t A = (int64_t)qoff + k0q;  // >= 1 (the arena has 16 bytes of front padding)
	const uint32_t *wp = (const uint32_t *)qual_out + (A >> 2);
	const uint32_t sh = 8u * (uint32_t)(A & 3);
	const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
	uint4 v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
	                     __funnelshift_r(w3, w4, sh));
	if (k0q >= 0 && k0q + 16 <= L) return v;
	const uint4 m = bytemask128(-k0q, L - k0q);
	return make_uint4(v.x & m.x, v.y & m.y, v.z & m.z, v.w & m.w);
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

template <int MODE, int CAT>
__global__ void __launch_bounds__(TL_THREADS) k_simulate_tiled(SimArgs a, int R, uint32_t stage_stride) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned char *stage = smem;  // NS * stage_stride
	uint32_t *s_recoff = (uint32_t *)(smem + (size_t)NS * stage_stride);  // [NS][R+1]
	uint64_t *s_gpos = (uint64_t *)(s_recoff + (size_t)NS * (R + 1) + ((NS * (R + 1)) & 1));  // [2][R]
	uint64_t *s_gend = s_gpos + 2 * (size_t)R;                                                 // [R]
	uint32_t *s_tmpl = (uint32_t *)(s_gend + R);                                                // [R]
	uint16_t *s_nl = (uint16_t *)(s_tmpl + R);                                                  // [R]
	__shared__ uint64_t s_G0[2], s_G1[2];

	const int tid = threadIdx.x;
	const uint64_t r0 = (uint64_t)blockIdx.x * (uint64_t)R;
	const int nr = (int)((a.nreads - r0) < (uint64_t)R ? (a.nreads - r0) : (uint64_t)R);
	const uint64_t *offs[2] = {a.off0, a.off1};
	const uint64_t *tots[2] = {a.tot0, a.tot1};
	char *outs[2] = {a.out0, a.out1};
	GenomeView gv{a.g.bases, a.g.nmask, a.genome_words - 2};

	if (tid < NS) {
		s_G0[tid] = offs[tid][r0];
		s_G1[tid] = (r0 + nr < a.nreads) ? offs[tid][r0 + nr] : *tots[tid];
	}
	__syncthreads();

	// ---------------- phase A: geometry ----------------
	for (int i = tid; i < nr; i += TL_THREADS) {
		const PlanRec pr = a.plan[r0 + i];
		const TmplDesc &t = a.c.t[pr.tmpl];
		const RecDev rec = a.g.recs[pr.record];
		const uint64_t ro0 = pr.refoff + t.delta[0], ro1 = pr.refoff + t.delta[1];
		uint32_t nl;
		if (CAT == CAT_U) nl = 13u + rec.name_len + ndig(ro0) + t.m[0].score_len + t.typ_len;
		else if (CAT == CAT_B) nl = 18u + 2u * rec.name_len + 2u * ndig(ro0) + 2u * t.m[0].score_len + t.typ_len;
		else nl = 18u + 2u * rec.name_len + ndig(ro0) + ndig(ro1) + t.m[0].score_len + t.m[1].score_len + t.typ_len;
		s_nl[i] = (uint16_t)nl;
		s_tmpl[i] = pr.tmpl;
		s_gpos[i] = rec.gstart + ro0;
		s_gpos[R + i] = rec.gstart + ro1;
		s_gend[i] = rec.gstart + rec.len;
#pragma unroll
		for (int s = 0; s < NS; s++) s_recoff[s * (R + 1) + i] = (uint32_t)(offs[s][r0 + i] - s_G0[s]);
	}
	if (tid < NS) s_recoff[tid * (R + 1) + nr] = (uint32_t)(s_G1[tid] - s_G0[tid]);
	__syncthreads();

	// ---------------- phase B: 16-byte chunks ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const int ph = (int)(s_G0[s] & 15);
		const int total = (int)(s_G1[s] - s_G0[s]);
		const int nchunks = (ph + total + 15) >> 4;
		const uint32_t *ro = s_recoff + s * (R + 1);
		unsigned char *stg = stage + (size_t)s * stage_stride;
		for (int c = tid; c < nchunks; c += TL_THREADS) {
			const int b0 = 16 * c - ph;
			const int bs = b0 > 0 ? b0 : 0;
			int lo = 0, hi = nr;
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				if ((int)ro[mid] <= bs) lo = mid;
				else hi = mid;
			}
			const int i = lo;
			const int rel = b0 - (int)ro[i];
			const TmplDesc &t = a.c.t[s_tmpl[i]];
			// which body does this stream carry for this read?
			int mate = s;
			bool random_body = false;
			if (CAT == CAT_B) {
				mate = 0;
				random_body = (s == 0) != (t.aligned_is_mate1 != 0);
			}
			const MateDesc &md = t.m[mate];
			const int L = random_body ? (int)t.opp_len : (int)md.len;
			const int k0 = rel - (int)s_nl[i];
			const int k0q = k0 - (L + 3);
			if (k0 + 16 <= 0 || k0q >= L) continue;  // name / trailing newline only: phase C
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if (k0 < L && !random_body)
				v = gen_seq(gv, a.c.segs + md.seg_off, md.nseg, md.fw != 0, L, s_gpos[mate * R + i], s_gend[i], k0);
			if (k0q + 16 > 0) {
				uint4 q;
				if (random_body) {
					const uint4 m = bytemask128(-k0q, L - k0q);
					q = make_uint4(0x49494949u & m.x, 0x49494949u & m.y, 0x49494949u & m.z, 0x49494949u & m.w);
				} else {
					q = gen_qual(a.c.qual_out, md.qual_out_off, L, k0q);
				}
				v.x |= q.x;
				v.y |= q.y;
				v.z |= q.z;
				v.w |= q.w;
			}
			*(uint4 *)(stg + 16 * (size_t)c) = v;
		}
	}
	__syncthreads();

	// ---------------- phase C: names, separators, random draws ----------------
	for (int i = tid; i < nr; i += TL_THREADS) {
		const TmplDesc &t = a.c.t[s_tmpl[i]];
		const PlanRec pr = a.plan[r0 + i];
		const RecDev rec = a.g.recs[pr.record];
		const char *refid = a.g.names + rec.name_off;
		const uint32_t nl = s_nl[i];
		char *rec0 = (char *)stage + (s_G0[0] & 15) + s_recoff[i];
		char *rec1 = NS == 2 ? (char *)stage + stage_stride + (s_G0[NS - 1] & 15) + s_recoff[(NS - 1) * (R + 1) + i] : nullptr;
		{
			char *p = s_put(rec0, "@qtip!:", 7);
			const uint64_t ro0 = pr.refoff + t.delta[0], ro1 = pr.refoff + t.delta[1];
			p = s_origin(p, refid, rec.name_len, t.m[0].fw != 0, ro0, t.m[0]);
			if (CAT == CAT_B) {
				*p++ = ':';
				p = s_origin(p, refid, rec.name_len, t.m[0].fw != 0, ro0, t.m[0]);
			} else if (CAT != CAT_U) {
				*p++ = ':';
				p = s_origin(p, refid, rec.name_len, t.m[1].fw != 0, ro1, t.m[1]);
			}
			*p++ = ':';
			p = s_put(p, t.typ, t.typ_len);
			*p++ = '\n';
			if (NS == 2)
				for (uint32_t j = 0; j < nl; j++) rec1[j] = rec0[j];
		}
		// body lengths per stream
		uint32_t bl0, bl1 = 0;
		bool s0_random = false, s1_random = false;
		if (CAT == CAT_U) {
			bl0 = t.m[0].len;
		} else if (CAT == CAT_B) {
			s0_random = !t.aligned_is_mate1;
			s1_random = !s0_random;
			bl0 = s0_random ? t.opp_len : t.m[0].len;
			bl1 = s1_random ? t.opp_len : t.m[0].len;
		} else {
			bl0 = t.m[0].len;
			bl1 = t.m[1].len;
		}
		char *seq0 = rec0 + nl, *seq1 = NS == 2 ? rec1 + nl : nullptr;
		seq0[bl0] = '\n';
		seq0[bl0 + 1] = '+';
		seq0[bl0 + 2] = '\n';
		seq0[2 * bl0 + 3] = '\n';
		if (NS == 2) {
			seq1[bl1] = '\n';
			seq1[bl1 + 1] = '+';
			seq1[bl1 + 2] = '\n';
			seq1[2 * bl1 + 3] = '\n';
		}

		// draws, in the reference's order: mate 1's transcript, mate 2's, then the random mate
		LState st{pr.s1, pr.s2};
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
		const int nm = CAT == CAT_C || CAT == CAT_D ? 2 : 1;
		for (int m = 0; m < nm; m++) {
			const MateDesc &md = t.m[m];
			char *seq = (CAT == CAT_B) ? (s0_random ? seq1 : seq0) : (m == 0 ? seq0 : seq1);
			const uint32_t *ops = a.c.ops + md.ops_off;
			const uint64_t gpos = s_gpos[m * R + i], gend = s_gend[i];
			const uint32_t T = md.fw ? 0x54474341u : 0x41434754u;  // "ACGT" / "TGCA"
			for (uint32_t j = 0; j < md.nops; j++) {
				const uint32_t e = __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				uint32_t code;
				if ((e & 3u) == OP_X) {
					const uint32_t ref = fetch_code_v(gv, gpos + (e >> 17), gend);
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
				seq[outp] = (char)((T >> (8u * code)) & 0xffu);
			}
		}
		if (CAT == CAT_B) {
			// init_random (simplesim.h:89-117): opp_len draws, printed with the aligned mate's strand
			char *seq = s0_random ? seq0 : seq1;
			const uint32_t L = t.opp_len;
			const bool fw = t.m[0].fw != 0;
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
	__syncthreads();

	// ---------------- phase D: staged tile -> HBM ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint64_t G0 = s_G0[s], G1 = s_G1[s];
		const int ph = (int)(G0 & 15);
		const int total = (int)(G1 - G0);
		const int nchunks = (ph + total + 15) >> 4;
		const unsigned char *stg = stage + (size_t)s * stage_stride;
		char *gbase = outs[s] + (G0 - ph);
		for (int c = tid; c < nchunks; c += TL_THREADS) {
			const int b0 = 16 * c - ph;
			if (b0 >= 0 && b0 + 16 <= total) {
				const uint4 v = *(const uint4 *)(stg + 16 * (size_t)c);
				__stcs((uint4 *)(gbase + 16 * (size_t)c), v);  // streaming store: written once, never re-read
			} else {
				for (int j = 0; j < 16; j++) {
					const int b = b0 + j;
					if (b >= 0 && b < total) gbase[16 * (size_t)c + j] = (char)stg[16 * (size_t)c + j];
				}
			}
		}
	}
}

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_rec_bytes, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	static uint32_t stage_bytes = 0;
	if (!stage_bytes) {
		const char *e = getenv("QTIPSIM_STAGE_KB");
		stage_bytes = e ? (uint32_t)atoi(e) * 1024u : TL_STAGE_DEFAULT;
		if (stage_bytes < 8 * 1024 || stage_bytes > 190 * 1024) stage_bytes = TL_STAGE_DEFAULT;
	}
	const uint32_t per_stream = stage_bytes / NS;
	int R = (int)((per_stream - 64) / max_rec_bytes);
	if (R > TL_THREADS) R = TL_THREADS;
	if (R < 4) return false;
	// a tile's bytes are at most R * max_rec_bytes; +32 for the alignment phase and the ragged tail
	const uint32_t stage_stride = ((uint32_t)R * max_rec_bytes + 32 + 15) & ~15u;
	const size_t meta = (size_t)NS * (R + 1) * 4 + 8 + 2 * (size_t)R * 8 + (size_t)R * 8 + (size_t)R * 4 + (size_t)R * 2 + 16;
	const size_t smem = (size_t)NS * stage_stride + meta;
	auto kern = k_simulate_tiled<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + R - 1) / R;
	kern<<<(unsigned)ntiles, TL_THREADS, smem, st>>>(a, R, stage_stride);
	kernels_count_launch();
	return true;
}

template <int MODE>
static bool launch_mode(const SimArgs &a, uint32_t max_rec_bytes, cudaStream_t st) {
	switch (a.cat) {
	case CAT_U: return launch_cat<MODE, CAT_U>(a, max_rec_bytes, st);
	case CAT_B: return launch_cat<MODE, CAT_B>(a, max_rec_bytes, st);
	case CAT_C: return launch_cat<MODE, CAT_C>(a, max_rec_bytes, st);
	default: return launch_cat<MODE, CAT_D>(a, max_rec_bytes, st);
	}
}

bool launch_simulate_tiled(const SimArgs &a, uint32_t max_rec_bytes, cudaStream_t st) {
	if (!a.nreads) return true;
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) return launch_mode<QTIPSIM_RNG_REPLAY>(a, max_rec_bytes, st);
	return launch_mode<QTIPSIM_RNG_PHILOX>(a, max_rec_bytes, st);
}

}  // namespace qt
