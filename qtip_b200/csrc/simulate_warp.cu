// simulate_warp.cu -- the production simulate+emit kernel (sm_100a): warp-autonomous tiles.
//
// Replaces SimulatedRead::init/mutate/init_random (simplesim.h:68-117, simplesim.cpp:31-71)
// and SimulatedRead::write/write_pair (simplesim.cpp:102-171).  The reference produces a read
// one byte at a time (a branch per transcript op, an fputc per base on the '-' strand); here
// the unit of work is 16 OUTPUT bytes, and a warp owns a tile of consecutive FASTQ records
// from plan record to HBM without ever meeting another warp at a barrier.
//
// A warp stages its tile (RW consecutive records, per output stream) in its private slice of
// shared memory at the same 16-byte phase the bytes have in the output file:
//
//   step A  lane = read: plan record + template -> per-read geometry (shared memory), L2
//           prefetch of the read's packed genome lines, the read's LAST sequence chunk (loads
//           issued here, used in step C), the read's first draw ops
//   step B  lane = 16-byte chunk, read-major (consecutive lanes = consecutive chunks of one read
//           => coalesced genome / quality loads, conflict-free 16-byte shared stores):
//             sequence   every chunk but the read's last: 32 bits of 2-bit genome funnel-
//                        shifted from two packed words (reversed for '-' mates), expanded to
//                        16 ASCII bases with PRMT ("ACGT" / "TGCA" as the byte table)
//             qualities  every chunk: two aligned 16-byte loads of the template's pre-oriented
//                        quality string, re-aligned to the output phase (branch-free word mux)
//           Chunks are stored WHOLE: bytes outside the string are garbage that lands only on
//           name / separator bytes, and those are written afterwards (step C).  Sequence and
//           quality stores never share a 16-byte chunk, so B needs no internal ordering.
//   step C  lane = read: last sequence chunk (exact bytes), name, separators, then the read's
//           random draws patched over the staged sequence (REPLAY: the reference's L'Ecuyer
//           stream stepped from the plan's state, with X redraws against the staged reference
//           base; PHILOX: counter-based)
//   step D  the staged tile goes to HBM: one cp.async.bulk (shared -> global) per stream for
//           the 16-byte aligned interior, byte stores for the ragged first/last chunk
//
// Reads that contain N, may reach past their record, or have indel segments take the general
// chunk generator (segment merge + N mask, from global memory) in steps A and B.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"

namespace qt {

#ifdef QT_BOUNDS
__device__ unsigned long long g_qw_violation[4];
#define QW_CHK(cond, id, v0, v1)                                                          \
	do {                                                                                   \
		if (!(cond)) {                                                                     \
			if (atomicCAS(&g_qw_violation[0], 0ull, (unsigned long long)(id)) == 0ull) {   \
				g_qw_violation[1] = (unsigned long long)(v0);                              \
				g_qw_violation[2] = (unsigned long long)(v1);                              \
				g_qw_violation[3] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;  \
			}                                                                              \
		}                                                                                  \
	} while (0)
#else
#define QW_CHK(cond, id, v0, v1) do { } while (0)
#endif

static constexpr int WT_MAX_WARPS = 8;
static constexpr int OPS_REG = 6;  // draw ops per mate held in registers from step A

enum : uint16_t { RF_FW = 1, RF_RANDOM = 2, RF_GENERAL = 4 };

// per-read, per-stream body geometry, shared memory (32 bytes; the fast path reads the first 16)
struct __align__(16) RdHot {
	uint32_t P;       // staging byte position of the first sequence byte
	uint16_t L;       // body length
	uint16_t flags;   // RF_* | (genome phase << 8)
	uint32_t w0;      // fast path: packed-genome word of chunk slot 0 (see geometry())
	uint32_t qoff;    // into DevCategory::qual_out
	uint64_t gpos;    // genome position of the mate's leftmost reference base
	uint32_t segoff;  // into DevCategory::segs
	uint16_t nseg;
	uint16_t pad;
};

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
__device__ __noinline__ uint4 seq16_general(const GenomeView &gv, const uint4 *lowmask, const SegDev *__restrict__ segs,
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
// geometry (step A): '+' mates  gadj = gpos + ref0 - (P & 15);       chunk j starts at base gadj + 16 j
//                    '-' mates  gadj = gpos + ref0 + (P & 15) - 15;  chunk j starts at base gadj - 16 j
// w0 = gadj >> 4, phase = gadj & 15 (in flags bits 8..11)
__device__ __forceinline__ uint4 seq16_fast(const uint32_t *__restrict__ bases, uint32_t w0, uint32_t flags, int j) {
	const bool fw = (flags & RF_FW) != 0;
	const uint32_t wi = fw ? w0 + (uint32_t)j : w0 - (uint32_t)j;
	const uint32_t lo = __ldg(&bases[wi]), hi = __ldg(&bases[wi + 1]);
	uint32_t w = __funnelshift_r(lo, hi, (flags >> 7) & 30u);
	if (!fw) w = rev16(w);
	return expand16(w, fw ? TBL_FW : TBL_RC);
}

// quality chunk: two aligned 16-byte loads ...
struct QualLoad {
	uint4 x, y;
	uint32_t a;
};
__device__ __forceinline__ QualLoad qual_issue(const char *__restrict__ qual_out, uint32_t qoff, int kq) {
	QualLoad ql;
	const uint32_t A = (uint32_t)((int)qoff + kq);  // >= 1 (the arena has 16 bytes of front padding)
	const uint4 *bp = (const uint4 *)qual_out + (A >> 4);
	ql.x = __ldg(bp);
	ql.y = __ldg(bp + 1);
	ql.a = A & 15u;
	return ql;
}
// ... then a branch-free barrel shift by a bytes: two word-select stages, one funnel stage
__device__ __forceinline__ uint4 qual_finish(const QualLoad &ql) {
	const uint4 &x = ql.x, &y = ql.y;
	const bool s2 = (ql.a & 8u) != 0, s1 = (ql.a & 4u) != 0;
	const uint32_t t0 = s2 ? x.z : x.x, t1 = s2 ? x.w : x.y, t2 = s2 ? y.x : x.z, t3 = s2 ? y.y : x.w, t4 = s2 ? y.z : y.x,
	               t5 = s2 ? y.w : y.y;
	const uint32_t u0 = s1 ? t1 : t0, u1 = s1 ? t2 : t1, u2 = s1 ? t3 : t2, u3 = s1 ? t4 : t3, u4 = s1 ? t5 : t4;
	const uint32_t sh = 8u * (ql.a & 3u);
	return make_uint4(__funnelshift_r(u0, u1, sh), __funnelshift_r(u1, u2, sh), __funnelshift_r(u2, u3, sh),
	                  __funnelshift_r(u3, u4, sh));
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
__device__ __forceinline__ char *s_put(char *p, const char *__restrict__ s, uint32_t n) {
	for (uint32_t i = 0; i < n; i++) p[i] = s[i];
	return p + n;
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
// "<refid>:<+|->:<off>:<score>" (simplesim.cpp:104-110)
__device__ __forceinline__ char *s_origin(char *p, const char *refid, uint32_t rl, uint64_t off, uint32_t nd,
                                          const MateDesc &m) {
	p = s_put(p, refid, rl);
	*p++ = ':';
	*p++ = m.fw ? '+' : '-';
	*p++ = ':';
	p = s_put_dec(p, off, nd);
	*p++ = ':';
	p = s_put(p, m.score, m.score_len);
	return p;
}

// ASCII base as staged -> 2-bit code of the REFERENCE base (4 for N); '-' mates are staged complemented
__device__ __forceinline__ uint32_t ref_code_of(uint32_t ch, bool fw) {
	if (ch == 'N') return 4u;
	uint32_t c = (ch >> 1) & 3u;  // A 0, C 1, T 2, G 3
	c ^= c >> 1;                  // A 0, C 1, G 2, T 3
	return fw ? c : 3u - c;
}

template <int MODE, int CAT>
__global__ void __launch_bounds__(WT_MAX_WARPS * 32) k_simulate_warp(SimArgs a, int RW, int K, uint32_t stage_stride,
                                                                      uint32_t warp_smem) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;                    // output streams
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;  // aligned mates per read
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ uint4 s_lowmask[17];

	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	if (tid < 17) {
		uint32_t w[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int k = tid - 4 * j;
			w[j] = k >= 4 ? 0xffffffffu : (k <= 0 ? 0u : (1u << (8 * k)) - 1u);
		}
		s_lowmask[tid] = make_uint4(w[0], w[1], w[2], w[3]);
	}
	__syncthreads();  // the only CTA-wide barrier: from here on every warp is on its own

	const uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
	const uint64_t r0 = tile * (uint64_t)RW;
	if (r0 >= a.nreads) return;
	const int nr = (int)((a.nreads - r0) < (uint64_t)RW ? (a.nreads - r0) : (uint64_t)RW);

	unsigned char *wsm = smem + (size_t)wib * warp_smem;
	unsigned char *stage = wsm;                                       // NS * stage_stride
	RdHot *s_hot = (RdHot *)(wsm + (size_t)NS * stage_stride);        // [NS][RW]
	uint64_t *s_gend = (uint64_t *)(s_hot + (size_t)NS * RW);         // [RW]

	const uint64_t *offs[2] = {a.off0, a.off1};
	const uint64_t *tots[2] = {a.tot0, a.tot1};
	char *outs[2] = {a.out0, a.out1};
	GenomeView gv;
	gv.bases = a.g.bases;
	gv.nmask = a.g.nmask;
	gv.max_word = a.genome_words - 2;

	const bool act = lane < nr;
	const uint64_t ri = r0 + (uint64_t)(act ? lane : 0);

	// ---------------- step A (lane = read) ----------------
	uint64_t G0[NS], G1[NS];
	uint32_t nl = 0, Pst[NS], Lst[NS];
	{
		uint64_t myoff[NS];
#pragma unroll
		for (int s = 0; s < NS; s++) {
			myoff[s] = offs[s][ri];
			G1[s] = (r0 + nr < a.nreads) ? offs[s][r0 + nr] : *tots[s];
			G0[s] = __shfl_sync(0xffffffffu, myoff[s], 0);
			uint64_t nxt = __shfl_down_sync(0xffffffffu, myoff[s], 1);
			if (lane == nr - 1) nxt = G1[s];
			Pst[s] = (uint32_t)(G0[s] & 15) + (uint32_t)(myoff[s] - G0[s]);  // record start; + nl below
			Lst[s] = (uint32_t)(nxt - myoff[s]);                             // record length; body length below
		}
	}
	const PlanRec pr = a.plan[ri];
	const TmplDesc &t = a.c.t[pr.tmpl];
	const RecDev rec = a.g.recs[pr.record];
	const uint64_t ro[2] = {pr.refoff + t.delta[0], pr.refoff + t.delta[1]};
	const uint64_t gend = rec.gstart + rec.len;
	uint32_t opsr[NM][OPS_REG];
	uint32_t nops[NM], ops_off[NM];
	uint4 tailv[NS];
	bool has_seq[NS];
	if (act) {
		s_gend[lane] = gend;
		// which mates need the general path: indel segments, reaching past the record end, or any N in the
		// 64-base blocks of the N-mask summary the mate touches (bases outside the mate only ever land on
		// garbage bytes of whole-chunk stores)
		bool general[2] = {false, false};
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const MateDesc &md = t.m[m];
			const uint64_t gpos = rec.gstart + ro[m];
			const uint64_t b0 = gpos >> 6, b1 = (gpos + (md.rflen ? md.rflen - 1u : 0u)) >> 6;  // blocks [b0, b1]
			bool anyn = false;
			for (uint64_t w = b0 >> 5; w <= (b1 >> 5); w++) {
				uint32_t bits = __ldg(&a.g.nblk[w]);
				const uint64_t lo = w << 5;
				if (b0 > lo) bits &= 0xffffffffu << (uint32_t)(b0 - lo);
				if (b1 < lo + 31) bits &= 0xffffffffu >> (uint32_t)(lo + 31 - b1);
				anyn |= bits != 0;
			}
			general[m] = anyn || !md.single || (gpos + md.rflen > gend);
			// L2 prefetch of the packed genome lines the chunk lanes are about to read
			const char *gl = (const char *)a.g.bases + ((gpos - 16) >> 2);
			const char *gh = (const char *)a.g.bases + ((gpos + md.rflen + 32) >> 2);
			for (const char *p = (const char *)((uintptr_t)gl & ~(uintptr_t)127); p <= gh; p += 128)
				asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
			nops[m] = md.nops;
			ops_off[m] = md.ops_off;
#pragma unroll
			for (int j = 0; j < OPS_REG; j++) opsr[m][j] = j < (int)md.nops ? __ldg(&a.c.ops[md.ops_off + j]) : 0u;
		}
#pragma unroll
		for (int s = 0; s < NS; s++) {
			int mate = s;
			bool random = false;
			if (CAT == CAT_B) {
				mate = 0;
				random = (s == 0) != (t.aligned_is_mate1 != 0);
			}
			const MateDesc &md = t.m[mate];
			const uint32_t L = random ? t.opp_len : md.len;
			if (s == 0) nl = Lst[0] - 2u * L - 4u;  // record = name + L + "\n+\n" + L + "\n"
			const uint32_t P = Pst[s] + nl;
			QW_CHK(P + 2u * L + 4u + 16u <= stage_stride, 5, P, L);
			const uint64_t gpos = rec.gstart + ro[mate];
			RdHot h;
			h.P = P;
			h.L = (uint16_t)L;
			h.qoff = md.qual_out_off;
			h.gpos = gpos;
			h.segoff = md.seg_off;
			h.nseg = md.nseg;
			h.pad = 0;
			const uint64_t gadj = md.fw ? gpos + (uint64_t)(int64_t)md.ref0 - (P & 15u)
			                            : gpos + (uint64_t)(int64_t)md.ref0 + (P & 15u) - 15u;
			h.w0 = (uint32_t)(gadj >> 4);
			h.flags = (uint16_t)((md.fw ? RF_FW : 0) | (random ? RF_RANDOM : 0) | (general[mate] ? RF_GENERAL : 0) |
			                     ((uint32_t)(gadj & 15u) << 8));
			s_hot[s * RW + lane] = h;
			Pst[s] = P;
			Lst[s] = L;
			// the read's last sequence chunk (loads issued now, stored in step C)
			has_seq[s] = !random && L > 0;
			tailv[s] = make_uint4(0u, 0u, 0u, 0u);
			if (has_seq[s]) {
				const int c_first = (int)(P >> 4), c_last = (int)((P + L - 1u) >> 4);
				if (general[mate])
					tailv[s] = seq16_general(gv, s_lowmask, a.c.segs + md.seg_off, md.nseg, md.fw != 0, (int)L, gpos, gend,
					                         16 * c_last - (int)P);
				else tailv[s] = seq16_fast(a.g.bases, h.w0, h.flags, c_last - c_first);
			}
		}
	}
	__syncwarp();

	// ---------------- step B (lane = chunk, read-major) ----------------
	{
		const int nitems = K * nr;
		const int di = 32 / K, dj = 32 - di * K;  // (read, slot) advance per 32 items, division-free
#pragma unroll
		for (int s = 0; s < NS; s++) {
			unsigned char *stg = stage + (size_t)s * stage_stride;
			const RdHot *hot = s_hot + s * RW;
			int i = lane / K, j = lane - i * K;
			for (int q = lane; q < nitems; q += 64) {
				uint4 sv[2];
				QualLoad ql[2];
				int sc[2], qc[2];
				bool slive[2], qlive[2], qrnd[2];
#pragma unroll
				for (int u = 0; u < 2; u++) {
					slive[u] = qlive[u] = qrnd[u] = false;
					if (q + 32 * u < nitems) {
						const uint4 h4 = *(const uint4 *)&hot[i];  // P | L, flags | w0 | qoff
						const uint32_t P = h4.x, L = h4.y & 0xffffu, flags = h4.y >> 16;
						// sequence: every chunk before the one holding the last base
						sc[u] = (int)(P >> 4) + j;
						slive[u] = !(flags & RF_RANDOM) && L > 0 && sc[u] < (int)((P + L - 1u) >> 4);
						if (slive[u]) {
							if (flags & RF_GENERAL) {
								const RdHot &hh = hot[i];
								sv[u] = seq16_general(gv, s_lowmask, a.c.segs + hh.segoff, hh.nseg, (flags & RF_FW) != 0, (int)L,
								                      hh.gpos, s_gend[i], 16 * sc[u] - (int)P);
							} else sv[u] = seq16_fast(a.g.bases, h4.z, flags, j);
						}
						// qualities: every chunk the string touches
						const uint32_t Pq = P + L + 3u;
						qc[u] = (int)(Pq >> 4) + j;
						qlive[u] = L > 0 && qc[u] <= (int)((Pq + L - 1u) >> 4);
						qrnd[u] = (flags & RF_RANDOM) != 0;
						if (qlive[u] && !qrnd[u]) ql[u] = qual_issue(a.c.qual_out, h4.w, 16 * qc[u] - (int)Pq);
					}
					i += di;
					j += dj;
					if (j >= K) {
						j -= K;
						i++;
					}
				}
#pragma unroll
				for (int u = 0; u < 2; u++) {
					if (slive[u]) *(uint4 *)(stg + 16 * (size_t)sc[u]) = sv[u];
					if (qlive[u]) {
						QW_CHK(16u * (uint32_t)qc[u] + 16u <= stage_stride, 4, qc[u], u);
						*(uint4 *)(stg + 16 * (size_t)qc[u]) =
						    qrnd[u] ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)  // 'I' (simplesim.h:105)
						            : qual_finish(ql[u]);
					}
				}
			}
		}
	}
	__syncwarp();

	// ---------------- step C (lane = read): last sequence chunk, name, separators, draws ----------------
	if (act) {
		char *recp[NS];
#pragma unroll
		for (int s = 0; s < NS; s++) {
			unsigned char *stg = stage + (size_t)s * stage_stride;
			recp[s] = (char *)stg + Pst[s] - nl;
			if (has_seq[s]) {
				// exact bytes: whole words where the word lies inside the string, single bytes otherwise
				const uint32_t P = Pst[s], L = Lst[s];
				const int c_last = (int)((P + L - 1u) >> 4);
				const int k0 = 16 * c_last - (int)P;
				const uint32_t vw[4] = {tailv[s].x, tailv[s].y, tailv[s].z, tailv[s].w};
				unsigned char *cp = stg + 16 * (size_t)c_last;
#pragma unroll
				for (int w = 0; w < 4; w++) {
					const int lo = k0 + 4 * w;  // output position of the word's first byte
					if (lo >= 0 && lo + 4 <= (int)L) *(uint32_t *)(cp + 4 * w) = vw[w];
					else {
#pragma unroll
						for (int b = 0; b < 4; b++)
							if (lo + b >= 0 && lo + b < (int)L) cp[4 * w + b] = (unsigned char)(vw[w] >> (8 * b));
					}
				}
			}
		}
		// name (simplesim.cpp:104-110, :140-150) and separators
		{
			const char *refid = a.g.names + rec.name_off;
			const uint32_t rl = rec.name_len;
			char *p = s_put(recp[0], "@qtip!:", 7);
			if (CAT == CAT_U) {
				const uint32_t nd = nl - 13u - rl - t.m[0].score_len - t.typ_len;
				p = s_origin(p, refid, rl, ro[0], nd, t.m[0]);
			} else if (CAT == CAT_B) {
				const uint32_t nd = (nl - 18u - 2u * rl - 2u * t.m[0].score_len - t.typ_len) >> 1;
				p = s_origin(p, refid, rl, ro[0], nd, t.m[0]);
				*p++ = ':';
				p = s_origin(p, refid, rl, ro[0], nd, t.m[0]);
			} else {
				const uint32_t nd0 = ndig64(ro[0]);
				const uint32_t nd1 = nl - 18u - 2u * rl - t.m[0].score_len - t.m[1].score_len - t.typ_len - nd0;
				p = s_origin(p, refid, rl, ro[0], nd0, t.m[0]);
				*p++ = ':';
				p = s_origin(p, refid, rl, ro[1], nd1, t.m[1]);
			}
			*p++ = ':';
			p = s_put(p, t.typ, t.typ_len);
			*p++ = '\n';
			QW_CHK(p == recp[0] + nl, 6, (uint32_t)(p - recp[0]), nl);
#pragma unroll
			for (int s = 0; s < NS; s++) {
				if (s == 1)
					for (uint32_t j = 0; j < nl; j++) recp[1][j] = recp[0][j];
				char *sq = recp[s] + nl;
				const uint32_t L = Lst[s];
				sq[L] = '\n';
				sq[L + 1] = '+';
				sq[L + 2] = '\n';
				sq[2 * L + 3] = '\n';
			}
		}
		// draws, in the reference's order: mate 1's transcript, mate 2's, then the random mate
		{
			LState st{pr.s1, pr.s2};
			const uint64_t gi = a.read_index_base + ri;
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
			auto draw_code = [&]() -> uint32_t {  // draw_base (simplesim.h:26-28)
				if (MODE == QTIPSIM_RNG_REPLAY) {
					st = lstep(st);
					return lbase_code(lfloat(lout(st)));
				}
				return draw16() >> 14;
			};
			// stream that holds the aligned mate m's sequence
			const bool s0_random = CAT == CAT_B && t.aligned_is_mate1 == 0;
#pragma unroll
			for (int m = 0; m < NM; m++) {
				const int s = CAT == CAT_B ? (s0_random ? 1 : 0) : m;
				char *seq = (char *)stage + (size_t)s * stage_stride + Pst[s];
				const bool fw = t.m[m].fw != 0;
				const uint32_t T = fw ? TBL_FW : TBL_RC;
				const uint32_t *ops = a.c.ops + ops_off[m];
				for (uint32_t j = 0; j < nops[m]; j++) {
					uint32_t e;
					if (j < (uint32_t)OPS_REG) {
						e = opsr[m][0];
#pragma unroll
						for (int k = 1; k < OPS_REG; k++) e = j == (uint32_t)k ? opsr[m][k] : e;
					} else e = __ldg(&ops[j]);
					const uint32_t outp = (e >> 2) & 0x7fffu;
					QW_CHK(outp < Lst[s], 7, outp, Lst[s]);
					uint32_t code;
					if ((e & 3u) == OP_X) {
						const uint32_t ref = ref_code_of((uint32_t)(unsigned char)seq[outp], fw);
						if (MODE == QTIPSIM_RNG_REPLAY) {
							do code = draw_code();  // simplesim.cpp:47-49
							while (code == ref);
						} else {
							const uint32_t b = draw16();
							code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
						}
					} else code = draw_code();
					seq[outp] = (char)((T >> (8u * code)) & 0xffu);
				}
			}
			if (CAT == CAT_B) {
				// init_random (simplesim.h:89-117): opp_len draws, printed with the aligned mate's strand
				const int s = s0_random ? 0 : 1;
				char *seq = (char *)stage + (size_t)s * stage_stride + Pst[s];
				const uint32_t L = Lst[s];
				const bool fw = t.m[0].fw != 0;
				const uint32_t T = fw ? TBL_FW : TBL_RC;
				for (uint32_t j = 0; j < L; j++) {
					const uint32_t code = draw_code();
					seq[fw ? j : L - 1 - j] = (char)((T >> (8u * code)) & 0xffu);
				}
			}
		}
	}
	// make the generic-proxy shared-memory writes visible to the async (bulk copy) proxy
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncwarp();

	// ---------------- step D: staged tile -> HBM ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const int ph = (int)(G0[s] & 15);
		const int total = (int)(G1[s] - G0[s]);
		const unsigned char *stg = stage + (size_t)s * stage_stride;
		char *gbase = outs[s] + (G0[s] - ph);  // 16-byte aligned
		// interior: whole 16-byte chunks [c_lo, c_hi)
		const int c_lo = ph ? 1 : 0;
		const int c_hi = (ph + total) >> 4;
		if (lane == 0 && c_hi > c_lo) {
			const uint32_t nbytes = (uint32_t)(c_hi - c_lo) * 16u;
			QW_CHK(16u * c_hi <= stage_stride && total >= 0, 8, c_hi, total);
			const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stg + 16 * c_lo);
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase + 16 * (size_t)c_lo),
			             "r"(saddr), "r"(nbytes)
			             : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		{
			// staging positions [ph, ph+total) are this tile's; chunk 0 (if ph != 0) and chunk c_hi are partial
			const int end = ph + total;
			const int p = lane < 16 ? lane : 16 * c_hi + (lane - 16);
			const int hi = lane < 16 ? (16 * c_lo < end ? 16 * c_lo : end) : end;
			if (p >= ph && p < hi) gbase[p] = (char)stg[p];
		}
	}
	// the staging buffer must stay allocated until the bulk copies have read it
	if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

struct WarpCfg {
	int rw = 0, warps = 0;
};

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	static int env_rw = -1, env_warps = -1, env_kb = -1;
	if (env_rw < 0) {
		const char *e = getenv("QTIPSIM_RW");
		env_rw = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_WARPS");
		env_warps = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_SMEM_KB");
		env_kb = e ? atoi(e) : 0;
	}
	auto warp_bytes = [&](int rw) -> uint32_t {
		const uint32_t stage_stride = ((uint32_t)rw * max_rec_bytes + 48 + 15) & ~15u;
		return NS * stage_stride + (uint32_t)rw * (NS * (uint32_t)sizeof(RdHot) + 8u);
	};
	// reads per warp tile: as many as keep >= 16 warps resident per SM (lane = read steps want full
	// warps, latency hiding wants resident warps), at least 4
	const uint32_t sm_budget = (env_kb > 0 ? (uint32_t)env_kb : 200u) * 1024u;
	int rw = env_rw > 0 ? env_rw : 32;
	if (rw > 32) rw = 32;
	if (env_rw <= 0)
		while (rw > 4 && warp_bytes(rw) * 16u > sm_budget) rw >>= 1;
	if (warp_bytes(rw) > sm_budget) return false;
	int warps = env_warps > 0 ? env_warps : 4;
	if (warps > WT_MAX_WARPS) warps = WT_MAX_WARPS;
	while (warps > 1 && warp_bytes(rw) * (uint32_t)warps > sm_budget) warps >>= 1;
	const uint32_t stage_stride = ((uint32_t)rw * max_rec_bytes + 48 + 15) & ~15u;
	const uint32_t wb = warp_bytes(rw);
	const size_t smem = (size_t)wb * warps;
	// chunks a sequence / quality string of max_body bytes can touch at any alignment
	const int K = (int)((max_body + 15u) / 16u) + 1;
	auto kern = k_simulate_warp<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + rw - 1) / rw;
	const uint64_t nblocks = (ntiles + warps - 1) / warps;
	kern<<<(unsigned)nblocks, warps * 32, smem, st>>>(a, rw, K, stage_stride, wb);
	kernels_count_launch();
	return true;
}

template <int MODE>
static bool launch_mode(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, cudaStream_t st) {
	switch (a.cat) {
	case CAT_U: return launch_cat<MODE, CAT_U>(a, max_rec_bytes, max_body, st);
	case CAT_B: return launch_cat<MODE, CAT_B>(a, max_rec_bytes, max_body, st);
	case CAT_C: return launch_cat<MODE, CAT_C>(a, max_rec_bytes, max_body, st);
	default: return launch_cat<MODE, CAT_D>(a, max_rec_bytes, max_body, st);
	}
}

bool launch_simulate_warp(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, cudaStream_t st) {
	if (!a.nreads) return true;
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) return launch_mode<QTIPSIM_RNG_REPLAY>(a, max_rec_bytes, max_body, st);
	return launch_mode<QTIPSIM_RNG_PHILOX>(a, max_rec_bytes, max_body, st);
}

// 64-base block summary of the N mask: bit b of nblk = any N in bases [64 b, 64 b + 64)
__global__ void __launch_bounds__(256) k_nblock(const uint32_t *__restrict__ nmask, uint64_t mask_words, uint32_t *nblk,
                                                uint64_t nblk_words) {
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // block index
	uint32_t any = 0;
	if (2 * b < mask_words) any = nmask[2 * b];
	if (2 * b + 1 < mask_words) any |= nmask[2 * b + 1];
	if (2 * b >= mask_words) any = 0xffffffffu;  // beyond the genome: N
	const uint32_t bits = __ballot_sync(0xffffffffu, any != 0);
	if ((threadIdx.x & 31) == 0 && (b >> 5) < nblk_words) nblk[b >> 5] = bits;
}

void launch_nblock(const uint32_t *nmask, uint64_t mask_words, uint32_t *nblk, uint64_t nblk_words, cudaStream_t st) {
	const uint64_t nthreads = nblk_words * 32;
	k_nblock<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(nmask, mask_words, nblk, nblk_words);
	kernels_count_launch();
}

#ifdef QT_BOUNDS
void warp_debug_report() {
	unsigned long long v[4] = {0, 0, 0, 0};
	cudaMemcpyFromSymbol(v, g_qw_violation, sizeof(v));
	if (v[0])
		fprintf(stderr, "QT_BOUNDS(warp) violation id=%llu v0=%lld v1=%lld block=%llu thread=%llu\n", v[0], (long long)v[1],
		        (long long)v[2], v[3] >> 32, v[3] & 0xffffffffu);
	else fprintf(stderr, "QT_BOUNDS(warp): no violation\n");
	memset(v, 0, sizeof(v));
	cudaMemcpyToSymbol(g_qw_violation, v, sizeof(v));
}
#endif

}  // namespace qt

#ifdef QT_BOUNDS
extern "C" void qt_warp_debug_report() { qt::warp_debug_report(); }
#endif
