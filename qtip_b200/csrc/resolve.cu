// resolve.cu -- K2: the reference's sequential draw chain, resolved on the device (REPLAY mode).
//
// In the reference every random decision comes from one generator stepped in program order
// (simplesim.cpp:216-397): per window four binomial counts (ignbin), per read a template draw,
// an offset draw and a data-dependent number of base draws (an X redraws while it equals the
// reference base).  So where read r+1 starts in the stream is only known after read r.
//
// Because the stream is addressable (state_k = state_0 * a^k, lecuyer.h), the walk can be
// speculated: one CTA handles the windows in order, and inside a window it evaluates, for 1024
// consecutive stream positions at once, "a read starting HERE would end THERE" (J[k]); the reads
// that really happen are the orbit of the region's first position under J, found with pointer
// doubling + marking in shared memory; a prefix count over the marks numbers them.  Threads whose
// position is on the orbit still hold their read's template, offset and generator state in
// registers and write the plan record themselves.
//
// ignbin (ranlib.cpp:1094-1389) runs on thread 0 from set-ups the host computed with its own libm
// (lecuyer.h: binom_setup / binom_sample); its logf is the glibc-exact restatement.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "resolve.h"

namespace qt {

static constexpr int RS_T = 1024;    // stream positions speculated per round
static constexpr int RS_LOG = 10;

__device__ __forceinline__ LState jump_tbl(LState s, uint32_t k, const uint32_t *jt1, const uint32_t *jt2) {
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

__device__ __forceinline__ uint32_t rs_fetch_code(const DevGenome &g, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;
	if ((__ldg(&g.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// draws consumed by one mate's transcript starting at state s (simplesim.cpp:38-62); advances s
__device__ __forceinline__ uint32_t walk_mate_dev(LState &s, const DevGenome &g, const XStep *__restrict__ xs,
                                                  const HostMateProg hp, uint64_t gpos, uint64_t gend,
                                                  const uint32_t *jt1, const uint32_t *jt2) {
	uint32_t len = 0;
	for (uint32_t i = hp.xs_begin; i < hp.xs_end; i++) {
		const XStep x = xs[i];
		if (x.skip) {
			s = jump_tbl(s, x.skip, jt1, jt2);
			len += x.skip;
		}
		const uint32_t ref = rs_fetch_code(g, gpos + x.rf, gend);
		uint32_t code;
		do {
			s = lstep(s);
			len++;
			code = lbase_code(lfloat(lout(s)));
		} while (code == ref);
	}
	if (hp.tail_skip) {
		s = jump_tbl(s, hp.tail_skip, jt1, jt2);
		len += hp.tail_skip;
	}
	return len;
}

__global__ void __launch_bounds__(RS_T, 1) k_resolve_replay(ResolveArgs a) {
	__shared__ uint16_t s_jump[RS_LOG + 1][RS_T];  // J^(2^r), clamped to RS_T (sink)
	__shared__ uint8_t s_mark[RS_T];
	__shared__ uint32_t s_wsum[32];
	__shared__ LState s_S;            // generator state at the region's current position
	__shared__ uint64_t s_n;          // reads still to resolve in the current sub-region
	__shared__ uint64_t s_count[NCAT];
	__shared__ unsigned long long s_draws;
	__shared__ uint32_t s_p0;         // orbit start within the round
	__shared__ uint32_t s_err;
	__shared__ uint64_t s_nc, s_nd;
	__shared__ uint32_t s_total;

	const int tid = threadIdx.x;
	const uint32_t lane = tid & 31, wid = tid >> 5;
	if (tid == 0) {
		s_S = LState{a.s1, a.s2};
		for (int c = 0; c < NCAT; c++) s_count[c] = 0;
		s_draws = 0;
		s_err = 0;
	}
	__syncthreads();

	for (uint32_t w = a.win_lo; w < a.win_hi; w++) {
		const Window wn = a.win[w];
		if (wn.retsz < a.olap) continue;                                                  // simplesim.cpp:221-223
		if ((int)a.ncount[w] > (int)(0.9 * (double)wn.retsz)) continue;                   // simplesim.cpp:228-234
		const uint64_t nslots = (uint64_t)wn.retsz - a.olap;
		const RecDev rec = a.g.recs[wn.record];
		const uint64_t gwin = rec.gstart + wn.refoff;
		const uint64_t gend = rec.gstart + rec.len;
		const BinomSetup *su = a.setups + (size_t)a.pclass[w] * NCAT;

		// regions in the reference's order: U, B, then C and D counts followed by C reads and D reads
		for (int region = 0; region < 4; region++) {
			const int cat = region;  // region 2 = C (after drawing nc AND nd), region 3 = D
			__syncthreads();  // everyone has left the previous region's loop before its state is rewritten
			if (tid == 0) {
				LState S = s_S;
				unsigned long long nd = 0;
				auto nx = [&]() {
					S = lstep(S);
					nd++;
					return lfloat(lout(S));
				};
				if (region <= 1) {
					s_n = (uint64_t)binom_sample(su[cat], nx);
				} else if (region == 2) {
					s_nc = (uint64_t)binom_sample(su[CAT_C], nx);
					s_nd = (uint64_t)binom_sample(su[CAT_D], nx);
					s_n = s_nc;
				} else {
					s_n = s_nd;
				}
				s_S = S;
				s_draws += nd;
				s_p0 = 0;
				if (region == 0 && s_n && nslots == 0) s_err = 1;  // assert(nslots > 0), simplesim.cpp:252
				if (s_n && a.ntmpl[cat] == 0) s_err = 2;
				if (s_count[cat] + s_n > a.cap[cat]) s_err = 3;    // plan buffer too small: host retries larger
			}
			__syncthreads();
			if (s_err) break;
			const TmplDesc *tdesc = a.t[cat];
			const XStep *xs = a.xsteps[cat];
			const HostMateProg *prog = a.prog[cat];
			const uint32_t ntm = a.ntmpl[cat];

			while (s_n > 0) {
				if (s_p0 >= (uint32_t)RS_T) {
					// the previous read ran past this whole round (very long transcript): move on
					__syncthreads();
					if (tid == 0) {
						s_S = jump_tbl(s_S, (uint32_t)RS_T, a.jt1, a.jt2);
						s_draws += RS_T;
						s_p0 -= (uint32_t)RS_T;
					}
					__syncthreads();
					continue;
				}
				// ---- speculate: a read starting at stream position tid (relative to s_S) ----
				const LState S = s_S;
				LState s = jump_tbl(S, (uint32_t)tid, a.jt1, a.jt2);
				s = lstep(s);
				const uint32_t ti = (uint32_t)lindex(lfloat(lout(s)), ntm);
				s = lstep(s);
				const uint64_t off = lindex(lfloat(lout(s)), nslots);
				const LState after2 = s;
				uint32_t len = 2;
				{
					const TmplDesc &t = tdesc[ti];
					len += walk_mate_dev(s, a.g, xs, prog[2 * ti], gwin + off + t.delta[0], gend, a.jt1, a.jt2);
					if (cat >= CAT_C)
						len += walk_mate_dev(s, a.g, xs, prog[2 * ti + 1], gwin + off + t.delta[1], gend, a.jt1, a.jt2);
					if (cat == CAT_B) len += t.opp_len;  // init_random, simplesim.h:108-113
				}
				const uint32_t J = (uint32_t)tid + len;
				s_jump[0][tid] = (uint16_t)(J < (uint32_t)RS_T ? J : (uint32_t)RS_T);
				s_mark[tid] = 0;
				__syncthreads();
				// ---- pointer doubling: s_jump[r] = J^(2^r) ----
				for (int r = 1; r <= RS_LOG; r++) {
					const uint32_t h = s_jump[r - 1][tid];
					s_jump[r][tid] = h < (uint32_t)RS_T ? s_jump[r - 1][h] : (uint16_t)RS_T;
					__syncthreads();
				}
				// ---- mark the orbit of p0: all J^d(p0), d < 2^(RS_LOG+1) ----
				if (tid == (int)s_p0) s_mark[tid] = 1;
				__syncthreads();
				for (int r = RS_LOG; r >= 0; r--) {
					if (s_mark[tid]) {
						const uint32_t h = s_jump[r][tid];
						if (h < (uint32_t)RS_T) s_mark[h] = 1;
					}
					__syncthreads();
				}
				// ---- number the marked positions (orbit order == position order) ----
				const uint32_t mk = s_mark[tid];
				const uint32_t bal = __ballot_sync(0xffffffffu, mk);
				if (lane == 0) s_wsum[wid] = __popc(bal);
				__syncthreads();
				if (wid == 0) {
					const uint32_t v = s_wsum[lane];
					uint32_t x = v;
					for (int o = 1; o < 32; o <<= 1) {
						const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
						if (lane >= (uint32_t)o) x += y;
					}
					s_wsum[lane] = x - v;  // exclusive prefix over warps
					if (lane == 31) s_total = x;
				}
				__syncthreads();
				const uint32_t total_marked = s_total;
				const uint32_t rank = s_wsum[wid] + __popc(bal & ((1u << lane) - 1u));
				const uint64_t n_left = s_n;
				const uint64_t base = s_count[cat];
				const bool take = mk && (uint64_t)rank < n_left;
				if (take) {
					PlanRec pr;
					pr.tmpl = ti;
					pr.record = wn.record;
					pr.refoff = wn.refoff + off;
					pr.s1 = after2.s1;
					pr.s2 = after2.s2;
					a.plan[cat][base + rank] = pr;
				}
				__syncthreads();
				// ---- advance: the last taken read's end is the next position ----
				const uint64_t ntake = n_left < (uint64_t)total_marked ? n_left : (uint64_t)total_marked;
				if (take && (uint64_t)rank == ntake - 1) {
					// this thread holds the last read taken in this round
					const uint32_t endpos = (uint32_t)tid + len;  // may exceed RS_T
					if (ntake == n_left) {
						// region finished inside this round
						s_S = jump_tbl(S, endpos, a.jt1, a.jt2);
						s_draws += endpos;
						s_p0 = 0;
					} else {
						s_S = jump_tbl(S, (uint32_t)RS_T, a.jt1, a.jt2);
						s_draws += RS_T;
						s_p0 = endpos - (uint32_t)RS_T;  // >= 0: the orbit left the round here
					}
					s_n = n_left - ntake;
					s_count[cat] = base + ntake;
				}
				__syncthreads();
			}
			if (s_err) break;
		}
		if (s_err) break;
	}
	if (tid == 0) {
		for (int c = 0; c < NCAT; c++) a.result[c] = s_count[c];
		a.result[4] = s_S.s1;
		a.result[5] = s_S.s2;
		a.result[6] = s_draws;
		a.result[7] = s_err;
	}
}

void launch_resolve_replay(const ResolveArgs &a, cudaStream_t st) {
	k_resolve_replay<<<1, RS_T, 0, st>>>(a);
	kernels_count_launch();
}

}  // namespace qt
