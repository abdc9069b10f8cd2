// resolve2.cu -- K2, second generation: the reference's sequential draw chain (REPLAY mode) resolved by
// band speculation over many windows at once.
//
// The chain (simplesim.cpp:216-397): per window four binomial counts (ignbin, ranlib.cpp:1094-1389) and, per read,
// a template draw, an offset draw and a data-dependent number of base draws, all from ONE generator.  Where window
// w starts in the stream is only known after window w-1 -- but it is known *approximately*: the draws a window
// consumes have a mean and a variance that follow from the binomial parameters and the templates, so window w's
// start lies, with overwhelming probability, in a band of width ~ sigma * sqrt(distance to the last exactly known
// position).  Because the stream is addressable (lecuyer.h), a window's whole transfer function
//
//      T_w : start position  ->  position after the window        (for every start in the band)
//
// can be computed without knowing the true start:
//   1. every candidate start draws its binomial count(s) (one thread per candidate);
//   2. for each category with reads, "a read starting at stream position q consumes len(q) draws" is evaluated
//      densely over the range the candidates can reach (one thread per position) into shared memory;
//   3. five pointer-doubling passes turn len into the 32-read jump length; a candidate then finds the position
//      after its n reads in n/32 + (n mod 32) shared-memory look-ups.
// A STAGE is one kernel launch: CTA j builds T for the j-th window after the stage's anchor (the last exactly
// known position); the last CTA to finish walks the chain through the K tables (one L2 look-up per window),
// records every window's start and region ends, and publishes the next anchor.  A window whose true start misses
// its band, or whose candidate ran off the dense table, ends the stage early: it becomes the first window of the
// next stage, where the band is the single exact position and the table is extended round by round until the
// window is done (exact mode; this is also how windows too large for one table are resolved).  Nothing is
// approximate: a table entry is either the exact end position or INVALID.
//
// After the last stage k_resolve2_counts re-draws each window's binomial counts at the now known region starts
// (cheap, parallel), a scan gives every (window, category) its slice of the plan, and k_resolve2_emit -- one CTA
// per (window, category), all in parallel -- re-evaluates len densely from the known start, marks the orbit and
// writes the PlanRec of every read on it (template, record, offset, generator state after the two index draws).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "resolve.h"
#include "resolve_dev.cuh"

namespace qt {

static constexpr int R2_T = 1024;                  // threads of a stage CTA
static constexpr int R2_TE = 512;                  // threads of an emit CTA
static constexpr uint32_t R2_SINK = 0xFFFFu;       // table entry: the walk leaves the table / length does not fit
static constexpr uint32_t R2_INVALID = 0xFFFFFFFFu;
static constexpr uint32_t R2_POS_LIMIT = 0xE0000000u;  // stage-relative positions stay below this
static constexpr int R2_STRIDE_LOG = 5;            // the stride table holds 2^5-read jumps
static constexpr uint32_t R2_MAIN_PTS = R2_RCAP_MAX / 64 + 2;  // stride points of a main orbit (a read is >= 2 draws)

// stream position k (relative) from state s: two table look-ups for k < 2^24, square-and-multiply above
__device__ __forceinline__ LState jump2(LState s, uint32_t k, const uint32_t *jt1, const uint32_t *jt2, const uint32_t *jh1,
                                        const uint32_t *jh2) {
	const uint32_t lo = k & 4095u, hi = (k >> 12) & 4095u, top = k >> 24;
	if (lo) {
		s.s1 = mulmod1(s.s1, __ldg(&jt1[lo]));
		s.s2 = mulmod2(s.s2, __ldg(&jt2[lo]));
	}
	if (hi) {
		s.s1 = mulmod1(s.s1, __ldg(&jh1[hi]));
		s.s2 = mulmod2(s.s2, __ldg(&jh2[hi]));
	}
	if (top) s = ljump(s, (uint64_t)top << 24);
	return s;
}

// what a read of one window needs to know about its window
struct WinCtx {
	uint32_t record;
	uint64_t refoff;
	uint64_t nslots;
	uint64_t gwin, gend;
	bool has_n;
	const BinomSetup *su;
	const BinomTables *tab;  // may be null (then the direct sampler is used)
};

__device__ __forceinline__ WinCtx make_winctx(const ResolveArgs &a, uint32_t w, const BinomTables *tabs = nullptr) {
	WinCtx c;
	const Window wn = a.win[w];
	c.record = wn.record;
	c.refoff = wn.refoff;
	c.nslots = (uint64_t)wn.retsz - a.olap;
	// N mask reads are skipped when neither this window nor the next one of the record holds an N (a mate whose
	// reference span exceeds its fragment can reach past its window, SURVEY Q9, but never past the next)
	c.has_n = true;
	if (a.n_skip_ok && a.ncount[w] == 0) {
		const bool next_same = w + 1 < a.nwin && a.win[w + 1].record == wn.record;
		c.has_n = next_same && a.ncount[w + 1] != 0;
	}
	const RecDev rec = a.g.recs[wn.record];
	c.gwin = rec.gstart + wn.refoff;
	c.gend = rec.gstart + rec.len;
	c.su = a.setups + (size_t)a.pclass[w] * NCAT;
	c.tab = tabs ? tabs + (size_t)a.pclass[w] * NCAT : nullptr;
	return c;
}

// draws consumed by a read / pair of category cat whose template draw is the first draw after state s
// (simplesim.cpp:243-273, :280-334, :342-397)
__device__ __forceinline__ uint32_t read_len(const ResolveArgs &a, const WinCtx &wc, int cat, LState s) {
	s = lstep(s);
	const uint32_t ti = (uint32_t)lindex(lfloat(lout(s)), a.ntmpl[cat]);
	s = lstep(s);
	const uint64_t off = lindex(lfloat(lout(s)), wc.nslots);
	uint32_t len = 2;
	const XProgRec *xp = a.xp[cat];
	if (xp) {
		// packed programs: one 32-byte load per mate, issued before anything depends on them
		const uint4 *src = (const uint4 *)(xp + ti);
		uint32_t w0[8], w1[8];
		{
			const uint4 x = __ldg(src), y = __ldg(src + 1);
			w0[0] = x.x, w0[1] = x.y, w0[2] = x.z, w0[3] = x.w, w0[4] = y.x, w0[5] = y.y, w0[6] = y.z, w0[7] = y.w;
		}
		if (cat >= CAT_C) {
			const uint4 x = __ldg(src + 2), y = __ldg(src + 3);
			w1[0] = x.x, w1[1] = x.y, w1[2] = x.z, w1[3] = x.w, w1[4] = y.x, w1[5] = y.y, w1[6] = y.z, w1[7] = y.w;
		}
		const bool packed = ((w0[5] >> 8) & 0xffu) != 0xffu && (cat < CAT_C || ((w1[5] >> 8) & 0xffu) != 0xffu);
		if (packed) {
			const uint32_t d0 = cat >= CAT_C ? w0[7] : 0u;
			len += walk_mate_xp(s, a.g, w0, wc.gwin + off + d0, wc.gend, wc.has_n, a.jt1, a.jt2);
			if (cat >= CAT_C) len += walk_mate_xp(s, a.g, w1, wc.gwin + off + w1[6], wc.gend, wc.has_n, a.jt1, a.jt2);
			if (cat == CAT_B) len += w0[6];  // init_random, simplesim.h:108-113
			return len;
		}
	}
	const TmplDesc *tdesc = a.t[cat];
	const HostMateProg *prog = a.prog[cat];
	uint32_t d0 = 0, d1 = 0;  // the mates' offsets from the drawn position are 0 for U and B
	if (cat >= CAT_C) {
		d0 = tdesc[ti].delta[0];
		d1 = tdesc[ti].delta[1];
	}
	len += walk_mate_dev(s, a.g, a.xsteps[cat], prog[2 * ti], wc.gwin + off + d0, wc.gend, wc.has_n, a.jt1, a.jt2);
	if (cat >= CAT_C) len += walk_mate_dev(s, a.g, a.xsteps[cat], prog[2 * ti + 1], wc.gwin + off + d1, wc.gend, wc.has_n, a.jt1, a.jt2);
	if (cat == CAT_B) len += tdesc[ti].opp_len;  // init_random, simplesim.h:108-113
	return len;
}

// L1[q] = draws of a read starting at relative position q (state S0 = position 0), for q < R
template <int NT>
__device__ __forceinline__ void dense_eval(const ResolveArgs &a, const WinCtx &wc, int cat, LState S0, uint32_t R, uint16_t *L1,
                                           const uint32_t *jh1, const uint32_t *jh2) {
	const uint32_t tid = threadIdx.x;
	if (tid >= R) return;
	LState s = jump2(S0, tid, a.jt1, a.jt2, jh1, jh2);
	const uint32_t c1 = __ldg(&a.jt1[NT]), c2 = __ldg(&a.jt2[NT]);  // a^NT: this thread's next position
	for (uint32_t q = tid; q < R; q += NT) {
		const uint32_t len = read_len(a, wc, cat, s);
		L1[q] = (uint16_t)(len < R2_SINK ? len : R2_SINK);
		s.s1 = mulmod1(s.s1, c1);
		s.s2 = mulmod2(s.s2, c2);
	}
}

// five doubling passes: returns the table of 32-read jump lengths (in TA).  Needs L1 complete (barrier before);
// ends with a barrier.
template <int NT>
__device__ __forceinline__ const uint16_t *build_stride(const uint16_t *L1, uint16_t *TA, uint16_t *TB, uint32_t R) {
	const uint16_t *src = L1;
	uint16_t *dst = TA;
	for (int pass = 0; pass < R2_STRIDE_LOG; pass++) {
		for (uint32_t q = threadIdx.x; q < R; q += NT) {
			const uint32_t l = src[q];
			uint32_t r = R2_SINK;
			if (l != R2_SINK) {
				const uint32_t q2 = q + l;
				if (q2 < R) {
					const uint32_t l2 = src[q2];
					if (l2 != R2_SINK && l + l2 < R2_SINK) r = l + l2;
				}
			}
			dst[q] = (uint16_t)r;
		}
		__syncthreads();
		src = dst;
		dst = dst == TA ? TB : TA;
	}
	return src;
}

// n reads forward from relative position p, as far as the tables reach
__device__ __forceinline__ void advance(const uint16_t *L1, const uint16_t *LS, uint32_t R, uint32_t &p, uint32_t &n) {
	while (n > 0 && p < R) {
		if (n >= (1u << R2_STRIDE_LOG)) {
			const uint32_t l = LS[p];
			if (l != R2_SINK) {
				p += l;
				n -= 1u << R2_STRIDE_LOG;
				continue;
			}
		}
		const uint32_t l = L1[p];
		if (l == R2_SINK) break;
		p += l;
		n--;
	}
}

// block-wide min / max / max; every thread gets the results.  red: 3 * 32 words of shared memory.
template <int NT>
__device__ __forceinline__ void block_minmax(uint32_t &mn, uint32_t &mx, uint32_t &mxn, uint32_t *red) {
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	mn = __reduce_min_sync(0xffffffffu, mn);
	mx = __reduce_max_sync(0xffffffffu, mx);
	mxn = __reduce_max_sync(0xffffffffu, mxn);
	__syncthreads();  // red may still be read from the previous call
	if (lane == 0) {
		red[wid] = mn;
		red[32 + wid] = mx;
		red[64 + wid] = mxn;
	}
	__syncthreads();
	constexpr int NW = NT / 32;
	uint32_t a = lane < NW ? red[lane] : 0xffffffffu, b = lane < NW ? red[32 + lane] : 0u, c = lane < NW ? red[64 + lane] : 0u;
	mn = __reduce_min_sync(0xffffffffu, a);
	mx = __reduce_max_sync(0xffffffffu, b);
	mxn = __reduce_max_sync(0xffffffffu, c);
}

// expected reach of n reads beyond their start, generously
__device__ __forceinline__ uint64_t reach(const Resolve2Args &a, int cat, uint32_t n) {
	const float fn = (float)n;
	const float e = fn * a.m1[cat] + 6.0f * sqrtf(fn * a.v1[cat] + 1.0f) + 2.0f * (float)a.maxfixed[cat] + 64.0f;
	return (uint64_t)e;
}

__device__ __forceinline__ unsigned long long r2_now() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define R2_STAMP(k)                                                                                      \
	do {                                                                                                 \
		if (a2.dbg && s == a2.dbg_stage && threadIdx.x == 0) a2.dbg[(size_t)blockIdx.x * 16 + (k)] = r2_now(); \
	} while (0)

// All candidates of one window move over the reads of one category: P[i] = position (stage-relative) of candidate
// i's next draw, N[i] = reads it still has to pass.  On return P[i] is the position after those reads, or INVALID.
// exact (a single candidate, the true start): the dense table is extended round by round until the reads are done.
__device__ __forceinline__ void advance_region(const Resolve2Args &a2, const WinCtx &wc, int cat, LState SA, uint32_t nb, bool exact,
                                               uint32_t *P, uint32_t *N, uint16_t *L1, uint16_t *TA, uint16_t *TB, uint32_t *red,
                                               uint32_t *pts, uint32_t s = 0xffffffffu) {
	const ResolveArgs &a = a2.ra;
	const uint32_t tid = threadIdx.x;
	for (;;) {
		uint32_t mn = 0xffffffffu, mx = 0, mxn = 0;
		for (uint32_t i = tid; i < nb; i += R2_T) {
			const uint32_t pos = P[i], n = N[i];
			if (pos != R2_INVALID && n > 0) {
				mn = min(mn, pos);
				mx = max(mx, pos);
				mxn = max(mxn, n);
			}
		}
		block_minmax<R2_T>(mn, mx, mxn, red);
		if (mxn == 0) break;
		if (a.ntmpl[cat] == 0) {  // reads of a category without templates: the caller reports it (exact mode: stuck)
			for (uint32_t i = tid; i < nb; i += R2_T)
				if (N[i] > 0) {
					P[i] = R2_INVALID;
					N[i] = 0;
				}
			__syncthreads();
			break;
		}
		const uint32_t rlo = mn;
		const uint64_t want = (uint64_t)(mx - rlo) + reach(a2, cat, mxn) + 1;
		const uint32_t R = want < (uint64_t)a2.rcap ? (uint32_t)want : a2.rcap;
		const LState S0 = jump2(SA, rlo, a.jt1, a.jt2, a2.jh1, a2.jh2);
		R2_STAMP(12);
		dense_eval<R2_T>(a, wc, cat, S0, R, L1, a2.jh1, a2.jh2);
		__syncthreads();
		R2_STAMP(13);
		const uint16_t *LS = build_stride<R2_T>(L1, TA, TB, R);
		R2_STAMP(14);
		if (!exact) {
			// ---- every candidate's orbit merges, within a few reads, with the orbit of the lowest position: that one
			// is enumerated (rank in TB, positions in SEL), a candidate walks single reads until it meets it and looks
			// its end up by rank.  (Unmerged candidates just keep walking: still exact.) ----
			for (uint32_t q = tid; q < R; q += R2_T) TB[q] = 0;
			if (tid == 0) {  // every 32nd read of the main orbit
				uint32_t p = 0, k = 0;
				pts[0] = 0;
				while (p < R && k + 1 < R2_MAIN_PTS) {
					const uint32_t l = LS[p];
					if (l == R2_SINK) break;
					p += l;
					pts[++k] = p;
				}
				pts[R2_MAIN_PTS] = k;
			}
			__syncthreads();
			const uint32_t nfull = pts[R2_MAIN_PTS];
			uint16_t *SEL = TA;  // the stride table is not needed any more
			__syncthreads();
			constexpr uint32_t STR = 1u << R2_STRIDE_LOG;
			for (uint32_t k = tid; k <= nfull; k += R2_T) {
				uint32_t p = pts[k], t = 0;
				const uint32_t lim = k < nfull ? STR : 0xffffffffu;  // the last chunk runs to the table's end
				for (; t < lim && p < R; t++) {
					TB[p] = (uint16_t)(STR * k + t + 1);
					SEL[STR * k + t] = (uint16_t)p;
					const uint32_t l = L1[p];
					if (l == R2_SINK) {
						p = R2_SINK;  // the orbit cannot be followed past this read
						t++;
						break;
					}
					p += l;
				}
				if (k == nfull) {
					SEL[STR * nfull + t] = (uint16_t)(p < R2_SINK ? p : R2_SINK);  // position after the last read (may lie past the table)
					pts[R2_MAIN_PTS + 1] = STR * nfull + t;  // reads on the main orbit
				}
			}
			__syncthreads();
			const uint32_t olen = pts[R2_MAIN_PTS + 1];
			R2_STAMP(15);
			for (uint32_t i = tid; i < nb; i += R2_T) {
				const uint32_t pos = P[i];
				uint32_t n = N[i];
				if (pos == R2_INVALID || n == 0) continue;
				uint32_t p = pos - rlo;
				uint32_t rk = 0;
				bool ok = true;
				while (n > 0) {
					if (p >= R) {
						ok = false;
						break;
					}
					rk = TB[p];
					if (rk) break;
					const uint32_t l = L1[p];
					if (l == R2_SINK) {
						ok = false;
						break;
					}
					p += l;
					n--;
				}
				uint32_t endp = p;
				if (ok && n > 0) {  // met the main orbit at its read rk - 1 with n reads to go
					const uint32_t idx = rk - 1 + n;
					if (idx <= olen && SEL[idx] != R2_SINK) endp = SEL[idx];
					else ok = false;
				}
				if (ok && (uint64_t)rlo + endp < (uint64_t)R2_POS_LIMIT) P[i] = rlo + endp;
				else P[i] = R2_INVALID;
				N[i] = 0;
			}
			__syncthreads();
			break;
		}
		for (uint32_t i = tid; i < nb; i += R2_T) {
			const uint32_t pos = P[i];
			uint32_t n = N[i];
			if (pos == R2_INVALID || n == 0) continue;
			uint32_t p = pos - rlo;
			const uint32_t p_in = p;
			if (p < R) advance(L1, LS, R, p, n);
			const bool fits = (uint64_t)rlo + p < (uint64_t)R2_POS_LIMIT;
			if (n == 0 && fits) {
				P[i] = rlo + p;
				N[i] = 0;
			} else if (p != p_in && fits) {
				P[i] = rlo + p;  // the table ended first: the next round starts here
				N[i] = n;
			} else {
				P[i] = R2_INVALID;
				N[i] = 0;
			}
		}
		__syncthreads();
		if (!exact) break;
	}
}

__global__ void __launch_bounds__(R2_T, 1) k_resolve2_stage(Resolve2Args a2, uint32_t s) {
	extern __shared__ __align__(16) unsigned char r2_smem[];
	uint16_t *L1 = (uint16_t *)r2_smem;
	uint16_t *TA = L1 + a2.rcap;
	uint16_t *TB = TA + a2.rcap;
	__shared__ uint32_t s_red[96];
	__shared__ uint32_t s_b[4];
	__shared__ uint32_t s_start[R2_KMAX];
	__shared__ uint2 s_slot[R2_KMAX];
	__shared__ uint32_t s_pts[R2_MAIN_PTS + 2];
	// the window's four binomial set-ups and their BTPE tables: every candidate of the window samples from them
	__shared__ __align__(16) BinomSetup s_su[NCAT];
	__shared__ __align__(16) BinomTables s_tab[NCAT];

	const ResolveArgs &a = a2.ra;
	const uint32_t tid = threadIdx.x;
	const StageAnchor an = a2.anchor[s];
	if (an.status != R2_RUNNING) {  // done or stuck: later launches of the batch only pass the verdict on
		if (blockIdx.x == 0 && tid == 0) a2.anchor[s + 1] = an;
		return;
	}
	R2_STAMP(0);
	const LState SA{an.s1, an.s2};
	const uint32_t K = gridDim.x;
	// which window of the stage this CTA takes: a ticket, not blockIdx -- the relay below waits for the CTA of window
	// j - 1, which therefore has to have STARTED whenever this one has (tickets are handed out in starting order)
	if (tid == 0) s_b[3] = atomicAdd(&a2.ticket[s], 1u);
	__syncthreads();
	const uint32_t j = s_b[3];
	const uint32_t ai = an.first + j;
	const size_t nbmax = a2.nbmax;

	uint32_t my_lo = 0, my_nb = 0;
	if (ai < a2.na) {
		// ---- this window's band of candidate starts, relative to the anchor ----
		uint32_t lo = 0, nb = 1;
		if (j > 0) {
			const double dM = a2.cumM[ai] - a2.cumM[an.first], dV = a2.cumV[ai] - a2.cumV[an.first];
			const double h = (double)a2.band_sigma * sqrt(dV) + 16.0;
			double l = floor(dM - h);
			double want = 2.0 * h + 2.0;
			if (want > (double)a2.nbmax) {
				want = (double)a2.nbmax;
				l = floor(dM - 0.5 * want);
			}
			if (l < 0) l = 0;
			if (l + want + 1.0 < (double)R2_POS_LIMIT) {
				lo = (uint32_t)l;
				nb = (uint32_t)want;
			} else {
				nb = 0;  // out of the stage's coordinate range: the walker stops here
			}
		}
		if (tid == 0) a2.slot[j] = make_uint2(lo, nb);
		my_lo = lo;
		my_nb = nb;
		const uint32_t w = a2.aw[ai];
		WinCtx wc = make_winctx(a, w, a2.tabs);
		{
			const uint32_t *src = (const uint32_t *)wc.su;
			uint32_t *dst = (uint32_t *)s_su;
			for (uint32_t k = tid; k < NCAT * sizeof(BinomSetup) / 4; k += R2_T) dst[k] = __ldg(src + k);
			if (wc.tab) {
				src = (const uint32_t *)wc.tab;
				dst = (uint32_t *)s_tab;
				for (uint32_t k = tid; k < NCAT * sizeof(BinomTables) / 4; k += R2_T) dst[k] = __ldg(src + k);
			}
			__syncthreads();
			wc.su = s_su;
			if (wc.tab) wc.tab = s_tab;
		}
		uint32_t *endr[NCAT];
		for (int r = 0; r < NCAT; r++) endr[r] = a2.end + ((size_t)r * K + j) * nbmax;
		uint32_t *cn = a2.cn + (size_t)j * nbmax, *cnd = a2.cnd + (size_t)j * nbmax;
		const bool exact = nb == 1;
		// a speculated window whose expected reads cannot fit one dense table is left to exact mode
		bool hopeless = false;
		if (!exact) {
			float e = 0;
			for (int c = 0; c < NCAT; c++) e = fmaxf(e, (float)wc.su[c].n * wc.su[c].pp * a2.m1[c]);
			hopeless = (double)nb + (double)e > (double)a2.rcap;
		}
		if (hopeless) {
			for (uint32_t i = tid; i < nb; i += R2_T) endr[CAT_D][i] = R2_INVALID;
		} else {
			for (int region = 0; region < 3; region++) {  // U, B, then C and D (both counts are drawn before C's reads)
				const int cat0 = region == 2 ? CAT_C : region;
				if (region == 1 && wc.su[CAT_B].n == 0 && wc.su[CAT_C].n == 0 && wc.su[CAT_D].n == 0) {
					// No targets for bad-end, concordant and discordant pairs (an unpaired-only model): the three counts are
					// drawn one after the other from the end of the unpaired reads -- ignbin(0, p) takes one uniform, or more
					// when it rounds to 1.0f -- and no reads follow: one jump per candidate instead of two, no empty advances
					const uint32_t *prev = endr[CAT_U];
					uint32_t *eB = endr[CAT_B], *eC = endr[CAT_C], *eDd = endr[CAT_D];
					for (uint32_t i = tid; i < nb; i += R2_T) {
						const uint32_t pos = prev[i];
						uint32_t p1 = R2_INVALID, p2 = R2_INVALID;
						if (pos != R2_INVALID) {
							LState S = jump2(SA, pos, a.jt1, a.jt2, a2.jh1, a2.jh2);
							uint32_t used = 0;
							auto nx = [&]() {
								S = lstep(S);
								used++;
								return lfloat(lout(S));
							};
							(void)binom_sample(wc.su[CAT_B], nx);
							const uint32_t u1 = used;
							(void)binom_sample(wc.su[CAT_C], nx);
							(void)binom_sample(wc.su[CAT_D], nx);
							if ((uint64_t)pos + used < (uint64_t)R2_POS_LIMIT) {
								p1 = pos + u1;
								p2 = pos + used;
							}
						}
						eB[i] = p1;
						eC[i] = p2;
						eDd[i] = p2;
					}
					__syncthreads();
					break;
				}
				uint32_t *cur = endr[cat0];
				const uint32_t *prev = region == 0 ? nullptr : endr[cat0 - 1];
				// Region 0: the candidates are CONSECUTIVE stream positions, and the uniform a draw at stream position q
				// yields depends on q alone: the band's uniforms are tabulated once (one jump per thread, then a constant
				// multiplier per 1024 positions) in the space the dense tables use later, and every candidate's sampler
				// reads them instead of jumping to its own position and stepping the generator
				float *F = (float *)r2_smem;
				const uint32_t nF = nb + 96u;
				const bool tab_u = region == 0 && nb > 1 && (size_t)nF * 4 <= (size_t)a2.rcap * 6;
				if (tab_u) {
					LState Sq = jump2(SA, lo + 1u + tid, a.jt1, a.jt2, a2.jh1, a2.jh2);
					const uint32_t c1 = __ldg(&a.jt1[R2_T]), c2 = __ldg(&a.jt2[R2_T]);
					for (uint32_t k = tid; k < nF; k += R2_T) {
						F[k] = lfloat(lout(Sq));  // the uniform of the draw that takes the stream from lo + k to lo + k + 1
						Sq.s1 = mulmod1(Sq.s1, c1);
						Sq.s2 = mulmod2(Sq.s2, c2);
					}
					__syncthreads();
				}
				for (uint32_t i = tid; i < nb; i += R2_T) {
					uint32_t pos = prev ? prev[i] : lo + i;
					uint32_t n = 0, nd = 0;
					if (pos != R2_INVALID) {
						LState S{0u, 0u};
						if (!tab_u) S = jump2(SA, pos, a.jt1, a.jt2, a2.jh1, a2.jh2);
						uint32_t used = 0;
						auto nx = [&]() {
							if (tab_u) {
								const uint32_t k = i + used;
								used++;
								if (k < nF) return F[k];
								return lfloat(lout(jump2(SA, pos + used, a.jt1, a.jt2, a2.jh1, a2.jh2)));  // (a sampler that ran past the table)
							}
							S = lstep(S);
							used++;
							return lfloat(lout(S));
						};
						// draw_binomial, simplesim.cpp:242, :279, :340-341
						n = (uint32_t)(wc.tab ? binom_sample_tab(wc.su[cat0], wc.tab[cat0], nx) : binom_sample(wc.su[cat0], nx));
						if (region == 2) nd = (uint32_t)(wc.tab ? binom_sample_tab(wc.su[CAT_D], wc.tab[CAT_D], nx) : binom_sample(wc.su[CAT_D], nx));
						if ((uint64_t)pos + used < (uint64_t)R2_POS_LIMIT) {
							pos += used;
						} else {
							pos = R2_INVALID;
							n = nd = 0;
						}
					}
					cur[i] = pos;
					cn[i] = n;
					if (region == 2) cnd[i] = nd;
				}
				__syncthreads();
				R2_STAMP(1 + 2 * region);
				advance_region(a2, wc, cat0, SA, nb, exact, cur, cn, L1, TA, TB, s_red, s_pts, region == 0 ? s : 0xffffffffu);
				R2_STAMP(2 + 2 * region);
				if (region == 2) {
					uint32_t *ed = endr[CAT_D];
					for (uint32_t i = tid; i < nb; i += R2_T) ed[i] = cur[i];
					__syncthreads();
					advance_region(a2, wc, CAT_D, SA, nb, exact, ed, cnd, L1, TA, TB, s_red, s_pts);
				}
			}
		}
	}

	// ---- the chain through the stage's tables, as a relay: window j + 1 starts where window j's table says, and CTA j
	// looks that up itself as soon as its table is built and its own start has arrived from CTA j - 1 (which, with the
	// narrower band, is usually done first): the chain is complete about one hop after the last table, instead of 148
	// dependent L2 look-ups (0.33 us each) by one thread after everything else.  relay[j] = (stage + 1) << 32 | start of
	// window j relative to the anchor, or R2_INVALID when the chain broke before it (band missed, table entry invalid).
	// Window slots are taken by ticket, so the CTA waited for has always started. ----
	R2_STAMP(8);
	__threadfence();
	__syncthreads();
	if (tid == 0 && ai < a2.na) {
		const unsigned long long tag = (unsigned long long)(s + 1u) << 32;
		uint32_t pos = 0;
		if (j > 0) {
			unsigned long long v;
			do v = *(volatile unsigned long long *)&a2.relay[j];
			while ((v >> 32) != (unsigned long long)(s + 1u));
			pos = (uint32_t)v;
		}
		uint32_t e = R2_INVALID;
		if (pos != R2_INVALID && pos >= my_lo && pos - my_lo < my_nb)
			e = __ldcg(a2.end + ((size_t)CAT_D * K + j) * nbmax + (pos - my_lo));
		*(volatile unsigned long long *)&a2.relay[j + 1] = tag | e;
		__threadfence();
	}
	__syncthreads();
	if (tid == 0) {
		const uint32_t prev = atomicAdd(&a2.counter[s], 1u);
		s_b[0] = prev == gridDim.x - 1;
	}
	__syncthreads();
	if (!s_b[0]) return;
	__threadfence();
	R2_STAMP(9);
	const uint32_t nstage = min(K, a2.na - an.first);
	for (uint32_t jj = tid; jj < nstage; jj += R2_T) s_slot[jj] = __ldcg(&a2.slot[jj]);
	__syncthreads();
	// The last CTA to finish collects the relay: window jj resolved iff the chain reached window jj + 1
	for (uint32_t jj = tid; jj < nstage; jj += R2_T) s_start[jj] = jj ? (uint32_t)__ldcg(&a2.relay[jj]) : 0u;
	if (tid == 0) s_b[1] = nstage;
	__syncthreads();
	for (uint32_t jj = tid; jj < nstage; jj += R2_T)
		if ((uint32_t)__ldcg(&a2.relay[jj + 1]) == R2_INVALID) atomicMin(&s_b[1], jj);
	__syncthreads();
	if (tid == 0) s_b[2] = s_b[1] ? (uint32_t)__ldcg(&a2.relay[s_b[1]]) : 0u;
	__syncthreads();
	R2_STAMP(10);
	const uint32_t nres = s_b[1];
	for (uint32_t jj = tid; jj < nres; jj += R2_T) {
		WinResolved wr;
		wr.stage = s;
		wr.rel = s_start[jj];
		const uint32_t idx = s_start[jj] - s_slot[jj].x;
		for (int r = 0; r < NCAT; r++) wr.end[r] = __ldcg(a2.end + ((size_t)r * K + jj) * nbmax + idx);
		a2.wres[an.first + jj] = wr;
	}
	if (tid == 0) {
		StageAnchor nx = an;
		if (nres == 0) {
			nx.status = R2_STUCK;  // the first window failed in exact mode: the caller falls back / reports
			nx.stuck_win = a2.aw[an.first];
		} else {
			const uint32_t pos = s_b[2];
			nx.abs = an.abs + pos;
			const LState S = jump2(SA, pos, a.jt1, a.jt2, a2.jh1, a2.jh2);
			nx.s1 = S.s1;
			nx.s2 = S.s2;
			nx.first = an.first + nres;
			nx.status = nx.first >= a2.na ? R2_DONE : R2_RUNNING;
		}
		a2.anchor[s + 1] = nx;
	}
	R2_STAMP(11);
}

// Per resolved window: the four binomial counts, re-drawn at the region starts the walk established, and where each
// category's reads start.  One thread per active window.
__global__ void __launch_bounds__(128) k_resolve2_counts(Resolve2Args a2, uint32_t *cnt /*[NCAT][na]*/, uint32_t *start /*[NCAT][na]*/,
                                                       uint32_t *err) {
	const ResolveArgs &a = a2.ra;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a2.na) return;
	const WinResolved wr = a2.wres[i];
	const StageAnchor an = a2.anchor[wr.stage];
	const LState SA{an.s1, an.s2};
	const uint32_t w = a2.aw[i];
	const WinCtx wc = make_winctx(a, w);
	uint32_t n[NCAT], st[NCAT];
	const uint32_t at[3] = {wr.rel, wr.end[CAT_U], wr.end[CAT_B]};
	for (int region = 0; region < 3; region++) {
		LState S = jump2(SA, at[region], a.jt1, a.jt2, a2.jh1, a2.jh2);
		uint32_t used = 0;
		auto nx = [&]() {
			S = lstep(S);
			used++;
			return lfloat(lout(S));
		};
		const int cat0 = region == 2 ? CAT_C : region;
		n[cat0] = (uint32_t)binom_sample(wc.su[cat0], nx);
		if (region == 2) n[CAT_D] = (uint32_t)binom_sample(wc.su[CAT_D], nx);
		st[cat0] = at[region] + used;
	}
	st[CAT_D] = wr.end[CAT_C];
	for (int c = 0; c < NCAT; c++) {
		cnt[(size_t)c * a2.na + i] = n[c];
		start[(size_t)c * a2.na + i] = st[c];
		if (n[c] && a.ntmpl[c] == 0) atomicMax(err, 2u);
		if (n[c] > 4096u) atomicMax(err + 1, n[c]);  // err[1]: the largest region (the host picks the emit kernels by it)
	}
	if (n[CAT_U] && wc.nslots == 0) atomicMax(err, 1u);  // assert(nslots > 0), simplesim.cpp:252
}

// Plan records of one (window, category): the orbit of the known start under "next read", n reads long.
__global__ void __launch_bounds__(R2_TE) k_resolve2_emit(Resolve2Args a2, const uint32_t *cnt, const uint32_t *start,
                                                        const R2Offsets offp, uint32_t *err) {
	extern __shared__ __align__(16) unsigned char r2_smem[];
	uint16_t *L1 = (uint16_t *)r2_smem;
	uint16_t *TA = L1 + a2.rcap_emit;
	uint16_t *TB = TA + a2.rcap_emit;
	__shared__ uint32_t s_pts[R2_EMIT_PTS];
	__shared__ uint32_t s_b[4];
	const ResolveArgs &a = a2.ra;
	const uint32_t tid = threadIdx.x;
	const uint32_t i = blockIdx.x;
	const int cat = (int)blockIdx.y;
	uint32_t n = cnt[(size_t)cat * a2.na + i];
	if (n == 0 || n <= a2.walk_max || a.ntmpl[cat] == 0) return;  // (few reads: k_resolve2_emit_walk has them)
	const WinResolved wr = a2.wres[i];
	const StageAnchor an = a2.anchor[wr.stage];
	const LState SA{an.s1, an.s2};
	const uint32_t w = a2.aw[i];
	const WinCtx wc = make_winctx(a, w);
	uint32_t pos = start[(size_t)cat * a2.na + i];
	uint64_t base = offp.p[cat][i];
	PlanRec *plan = a.plan[cat];
	constexpr uint32_t STR = 1u << R2_STRIDE_LOG;
	while (n > 0) {
		uint64_t want = reach(a2, cat, n) + 1;
		const uint32_t R = want < (uint64_t)a2.rcap_emit ? (uint32_t)want : a2.rcap_emit;
		const LState S0 = jump2(SA, pos, a.jt1, a.jt2, a2.jh1, a2.jh2);
		dense_eval<R2_TE>(a, wc, cat, S0, R, L1, a2.jh1, a2.jh2);
		__syncthreads();
		const uint16_t *LS = build_stride<R2_TE>(L1, TA, TB, R);  // in TA; TB is free for the ranks
		for (uint32_t q = tid; q < R; q += R2_TE) TB[q] = 0;
		if (tid == 0) {  // every 32nd read of the orbit
			uint32_t p = 0, k = 0, rem = n;
			s_pts[0] = 0;
			while (rem >= STR && p < R && k + 1 < R2_EMIT_PTS) {
				const uint32_t l = LS[p];
				if (l == R2_SINK) break;
				p += l;
				rem -= STR;
				s_pts[++k] = p;
			}
			s_b[0] = k;
		}
		__syncthreads();
		const uint32_t nfull = s_b[0];
		// chunk k: the reads 32 k .. 32 k + 31 of the orbit (the last chunk: as far as the reads / the table go)
		for (uint32_t k = tid; k <= nfull; k += R2_TE) {
			uint32_t p = s_pts[k];
			const uint32_t lim = k < nfull ? STR : n - STR * nfull;
			uint32_t t = 0;
			for (; t < lim && p < R; t++) {
				TB[p] = (uint16_t)(STR * k + t + 1);
				const uint32_t l = L1[p];
				if (l == R2_SINK) {  // a read of >= 65535 draws: not representable here (the stage kernel got stuck first)
					atomicMax(err, 4u);
					t++;
					p = R;
					break;
				}
				p += l;
			}
			if (k == nfull) {
				s_b[1] = STR * nfull + t;  // reads taken in this round
				s_b[2] = p;                // where the next round starts
			}
		}
		__syncthreads();
		{
			LState sq = jump2(S0, tid, a.jt1, a.jt2, a2.jh1, a2.jh2);
			const uint32_t c1 = __ldg(&a.jt1[R2_TE]), c2 = __ldg(&a.jt2[R2_TE]);
			for (uint32_t q = tid; q < R; q += R2_TE) {
				const uint32_t rk = TB[q];
				if (rk) {
					LState t = lstep(sq);
					const uint32_t ti = (uint32_t)lindex(lfloat(lout(t)), a.ntmpl[cat]);
					t = lstep(t);
					const uint64_t off = lindex(lfloat(lout(t)), wc.nslots);
					PlanRec pr;
					pr.tmpl = ti;
					pr.record = wc.record;
					pr.refoff = wc.refoff + off;
					pr.s1 = t.s1;
					pr.s2 = t.s2;
					plan[base + rk - 1] = pr;
				}
				sq.s1 = mulmod1(sq.s1, c1);
				sq.s2 = mulmod2(sq.s2, c2);
			}
		}
		const uint32_t taken = s_b[1], endp = s_b[2];
		__syncthreads();  // s_b, s_pts and the tables are rewritten by the next round
		if (taken == 0 || (uint64_t)pos + endp >= (uint64_t)R2_POS_LIMIT) {
			if (tid == 0) atomicMax(err, 5u);
			break;
		}
		pos += endp;
		n -= taken;
		base += taken;
	}
	// the stage kernel's table-driven walk and this enumeration must agree on where the category's reads end
	if (tid == 0 && n == 0 && pos != wr.end[cat]) atomicMax(err, 6u);
}

// The same plan records for the (window, category) pairs with few reads -- the usual case -- by simply WALKING the orbit:
// one thread per window follows its n reads from the known start, evaluating only the reads that happen (the CTA version
// above evaluates every stream position of the region to find them in parallel: ~6 x the work at 5 draws per read, worth
// it only when one region holds thousands of reads).  Thousands of windows walk concurrently, so the per-read latency
// does not matter.
__global__ void __launch_bounds__(128) k_resolve2_emit_walk(Resolve2Args a2, int cat, const uint32_t *cnt, const uint32_t *start, const uint64_t *off,
                                                          uint32_t walk_max, uint32_t *err) {
	const ResolveArgs &a = a2.ra;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a2.na) return;
	uint32_t n = cnt[(size_t)cat * a2.na + i];
	if (n == 0 || n > walk_max || a.ntmpl[cat] == 0) return;
	const WinResolved wr = a2.wres[i];
	const StageAnchor an = a2.anchor[wr.stage];
	const WinCtx wc = make_winctx(a, a2.aw[i]);
	uint32_t pos = start[(size_t)cat * a2.na + i];
	LState s = jump2(LState{an.s1, an.s2}, pos, a.jt1, a.jt2, a2.jh1, a2.jh2);
	PlanRec *plan = a.plan[cat] + off[i];
	for (uint32_t r = 0; r < n; r++) {
		LState t = lstep(s);
		const uint32_t ti = (uint32_t)lindex(lfloat(lout(t)), a.ntmpl[cat]);
		t = lstep(t);
		const uint64_t of = lindex(lfloat(lout(t)), wc.nslots);
		PlanRec pr;
		pr.tmpl = ti;
		pr.record = wc.record;
		pr.refoff = wc.refoff + of;
		pr.s1 = t.s1;
		pr.s2 = t.s2;
		plan[r] = pr;
		const uint32_t len = read_len(a, wc, cat, s);
		s = jump2(s, len, a.jt1, a.jt2, a2.jh1, a2.jh2);
		pos += len;
	}
	if (pos != wr.end[cat]) atomicMax(err, 6u);  // must end where the stage's table said
}

void launch_resolve2_stage(const Resolve2Args &a, uint32_t s, uint32_t K, cudaStream_t st) {
	static bool attr = false;
	if (!attr) {
		cudaFuncSetAttribute(k_resolve2_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3u * R2_RCAP_MAX * 2u));
		attr = true;
	}
	k_resolve2_stage<<<K, R2_T, (size_t)3 * a.rcap * 2, st>>>(a, s);
	kernels_count_launch();
}

void launch_resolve2_counts(const Resolve2Args &a, uint32_t *cnt, uint32_t *start, uint32_t *err, cudaStream_t st) {
	k_resolve2_counts<<<(a.na + 127) / 128, 128, 0, st>>>(a, cnt, start, err);
	kernels_count_launch();
}

void launch_resolve2_emit(const Resolve2Args &a, const uint32_t *cnt, const uint32_t *start, const R2Offsets &offp, uint32_t *err,
                          cudaStream_t st) {
	static bool attr = false;
	if (!attr) {
		cudaFuncSetAttribute(k_resolve2_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3u * R2_RCAP_MAX * 2u));
		attr = true;
	}
	for (int c = 0; c < NCAT; c++) {
		if (!a.ra.cap[c] || !a.walk_max) continue;  // (cap[c] = reads of the category)
		k_resolve2_emit_walk<<<(a.na + 127) / 128, 128, 0, st>>>(a, c, cnt, start, offp.p[c], a.walk_max, err);
		kernels_count_launch();
	}
	if (a.any_large) {
		dim3 grid(a.na, NCAT);
		k_resolve2_emit<<<grid, R2_TE, (size_t)3 * a.rcap_emit * 2, st>>>(a, cnt, start, offp, err);
		kernels_count_launch();
	}
}

}  // namespace qt
