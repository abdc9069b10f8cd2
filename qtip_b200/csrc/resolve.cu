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
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "resolve.h"
#include "resolve_dev.cuh"

namespace qt {

static constexpr int RS_T = 1024;    // stream positions speculated per round
static constexpr int RS_LOG = 10;

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
		// N mask reads are skipped when neither this window nor the next one of the record holds an N (a mate
		// whose reference span exceeds its fragment can reach past its window, SURVEY Q9, but never past the next)
		bool has_n = true;
		if (a.n_skip_ok && a.ncount[w] == 0) {
			const bool next_same = w + 1 < a.nwin && a.win[w + 1].record == wn.record;
			has_n = next_same && a.ncount[w + 1] != 0;
		}
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
					// the mates' offsets from the drawn position are 0 for U and B: no descriptor load there
					uint32_t d0 = 0, d1 = 0;
					if (cat >= CAT_C) {
						d0 = tdesc[ti].delta[0];
						d1 = tdesc[ti].delta[1];
					}
					len += walk_mate_dev(s, a.g, xs, prog[2 * ti], gwin + off + d0, gend, has_n, a.jt1, a.jt2);
					if (cat >= CAT_C) len += walk_mate_dev(s, a.g, xs, prog[2 * ti + 1], gwin + off + d1, gend, has_n, a.jt1, a.jt2);
					if (cat == CAT_B) len += tdesc[ti].opp_len;  // init_random, simplesim.h:108-113
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

// ---------------------------------------------------------------------------------------------------
// Cluster version: the speculation of one round is spread over the SMs of a thread-block cluster.
//
// Evaluating J for 1024 positions on ONE SM is bound by that SM's L1 (every look-up is a gather) and by
// its issue slots; the orbit resolution that follows is cheap.  Here CL_NC CTAs of a cluster evaluate
// CL_NC * CL_T consecutive stream positions together -- a whole window in one round at the usual read
// densities -- and write their J values into CTA 0's shared memory through distributed shared memory.
// CTA 0 resolves the orbit (pointer doubling, marking, prefix count) over all of them; every thread of the
// cluster then reads mark and rank of its own position back over DSMEM and writes its plan record from
// registers, exactly as in the single-CTA kernel.  Three cluster barriers per round.
// ---------------------------------------------------------------------------------------------------
static constexpr int CL_NC = 8;                 // CTAs per cluster (portable maximum)
static constexpr int CL_T = 512;                // threads per CTA = positions per CTA and round
static constexpr int CL_N = CL_NC * CL_T;       // positions per round
static constexpr int CL_LOG = 12;               // log2(CL_N)
static constexpr int CL_PER = CL_N / CL_T;      // entries per thread of CTA 0 in the orbit resolution

struct ClusterShared {
	uint16_t jump[CL_LOG + 1][CL_N];  // J^(2^r), clamped to CL_N (sink)
	uint16_t rank[CL_N];              // number of marked positions before this one
	uint8_t mark[CL_N];
	uint32_t wsum[CL_T / 32];
	LState S;                         // generator state at the round's first position
	uint64_t n;                       // reads still to resolve in the current region
	uint64_t count[NCAT];
	unsigned long long draws;
	uint32_t p0;                      // orbit start within the round
	uint32_t err;
	uint64_t nc, nd;
	uint32_t total;
	uint32_t w, region;               // window and region (category) being resolved
	uint32_t fresh, done;
};

__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of CTA 0's copy of a shared-memory object (distributed shared memory)
template <typename T>
__device__ __forceinline__ T *cta0(T *p) {
	uint64_t out;
	asm volatile("mapa.u64 %0, %1, 0;" : "=l"(out) : "l"((uint64_t)(uintptr_t)p));
	return (T *)(uintptr_t)out;
}

__global__ void __cluster_dims__(CL_NC, 1, 1) __launch_bounds__(CL_T, 1) k_resolve_replay_cluster(ResolveArgs a) {
	extern __shared__ __align__(16) unsigned char cl_smem[];
	ClusterShared *own = (ClusterShared *)cl_smem;
	ClusterShared *sh = cta0(own);  // every CTA works on CTA 0's copy
	uint32_t crank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
	const int tid = threadIdx.x;
	const uint32_t lane = tid & 31, wid = tid >> 5;
	const bool lead = crank == 0;
	const uint32_t pos = crank * (uint32_t)CL_T + (uint32_t)tid;  // this thread's position within a round

	if (lead && tid == 0) {
		own->S = LState{a.s1, a.s2};
		for (int c = 0; c < NCAT; c++) own->count[c] = 0;
		own->draws = 0;
		own->err = 0;
		own->n = 0;
		own->w = a.win_lo;
		own->region = 0;
		own->fresh = 1;
		own->done = 0;
	}

	for (;;) {
		cluster_sync_all();  // everyone has left the previous region's rounds before its state is rewritten
		// ---- thread 0 of CTA 0 walks the reference's sequence of binomial draws (simplesim.cpp:216-397) up
		// to the next region that has reads: regions without reads cost one draw and no barrier ----
		if (lead && tid == 0) {
			LState S = own->S;
			unsigned long long nd = 0;
			auto nx = [&]() {
				S = lstep(S);
				nd++;
				return lfloat(lout(S));
			};
			uint32_t w = own->w, region = own->region;
			if (!own->fresh) region++;  // the region just resolved
			own->fresh = 0;
			uint64_t n = 0;
			for (;;) {
				if (region == 4) {
					w++;
					region = 0;
				}
				if (w >= a.win_hi) {
					own->done = 1;
					break;
				}
				const Window wn = a.win[w];
				if (region == 0 && (wn.retsz < a.olap || (int)a.ncount[w] > (int)(0.9 * (double)wn.retsz))) {
					w++;  // simplesim.cpp:221-223, :228-234: the window is skipped, no draws
					continue;
				}
				const BinomSetup *su = a.setups + (size_t)a.pclass[w] * NCAT;
				if (region <= 1) {
					n = (uint64_t)binom_sample(su[region], nx);
				} else if (region == 2) {
					own->nc = (uint64_t)binom_sample(su[CAT_C], nx);
					own->nd = (uint64_t)binom_sample(su[CAT_D], nx);
					n = own->nc;
				} else {
					n = own->nd;
				}
				if (n) {
					if (region == 0 && (uint64_t)wn.retsz - a.olap == 0) own->err = 1;  // assert(nslots > 0), simplesim.cpp:252
					if (a.ntmpl[region] == 0) own->err = 2;
					if (own->count[region] + n > a.cap[region]) own->err = 3;  // plan buffer too small: host retries larger
					break;
				}
				region++;
			}
			own->S = S;
			own->draws += nd;
			own->p0 = 0;
			own->n = n;
			own->w = w;
			own->region = region;
		}
		cluster_sync_all();
		if (sh->done || sh->err) break;
		const uint32_t w = sh->w;
		const int cat = (int)sh->region;
		const Window wn = a.win[w];
		const uint64_t nslots = (uint64_t)wn.retsz - a.olap;
		// N mask reads are skipped when neither this window nor the next one of the record holds an N (a mate
		// whose reference span exceeds its fragment can reach past its window, SURVEY Q9, but never past the next)
		bool has_n = true;
		if (a.n_skip_ok && a.ncount[w] == 0) {
			const bool next_same = w + 1 < a.nwin && a.win[w + 1].record == wn.record;
			has_n = next_same && a.ncount[w + 1] != 0;
		}
		const RecDev rec = a.g.recs[wn.record];
		const uint64_t gwin = rec.gstart + wn.refoff;
		const uint64_t gend = rec.gstart + rec.len;
		const TmplDesc *tdesc = a.t[cat];
		const XStep *xs = a.xsteps[cat];
		const HostMateProg *prog = a.prog[cat];
		const uint32_t ntm = a.ntmpl[cat];

		while (sh->n > 0) {  // uniform over the cluster: read after a cluster barrier, rewritten only before the next
			const uint32_t p0 = sh->p0;
			if (p0 >= (uint32_t)CL_N) {
				// the previous read ran past this whole round (very long transcript): move on
				cluster_sync_all();
				if (lead && tid == 0) {
					own->S = jump_tbl(own->S, (uint32_t)CL_N, a.jt1, a.jt2);
					own->draws += CL_N;
					own->p0 -= (uint32_t)CL_N;
				}
				cluster_sync_all();
				continue;
			}
			// ---- speculate: a read starting at stream position pos (relative to S) ----
			const LState S = sh->S;
			const uint64_t n_left = sh->n;
			const uint64_t base = sh->count[cat];
			LState s = jump_tbl(S, pos, a.jt1, a.jt2);
			s = lstep(s);
			const uint32_t ti = (uint32_t)lindex(lfloat(lout(s)), ntm);
			s = lstep(s);
			const uint64_t off = lindex(lfloat(lout(s)), nslots);
			const LState after2 = s;
			uint32_t len = 2;
			{
				uint32_t d0 = 0, d1 = 0;  // the mates' offsets from the drawn position are 0 for U and B
				if (cat >= CAT_C) {
					d0 = tdesc[ti].delta[0];
					d1 = tdesc[ti].delta[1];
				}
				len += walk_mate_dev(s, a.g, xs, prog[2 * ti], gwin + off + d0, gend, has_n, a.jt1, a.jt2);
				if (cat >= CAT_C) len += walk_mate_dev(s, a.g, xs, prog[2 * ti + 1], gwin + off + d1, gend, has_n, a.jt1, a.jt2);
				if (cat == CAT_B) len += tdesc[ti].opp_len;  // init_random, simplesim.h:108-113
			}
			const uint32_t J = pos + len;
			sh->jump[0][pos] = (uint16_t)(J < (uint32_t)CL_N ? J : (uint32_t)CL_N);
			sh->mark[pos] = 0;
			cluster_sync_all();
			// ---- CTA 0: orbit of p0 under J ----
			if (lead) {
				for (int r = 1; r <= CL_LOG; r++) {  // pointer doubling: jump[r] = J^(2^r)
#pragma unroll
					for (int m = 0; m < CL_PER; m++) {
						const int e = tid + m * CL_T;
						const uint32_t h = own->jump[r - 1][e];
						own->jump[r][e] = h < (uint32_t)CL_N ? own->jump[r - 1][h] : (uint16_t)CL_N;
					}
					__syncthreads();
				}
				if (tid == 0) own->mark[p0] = 1;
				__syncthreads();
				for (int r = CL_LOG; r >= 0; r--) {  // mark all J^d(p0), d < 2^(CL_LOG+1)
#pragma unroll
					for (int m = 0; m < CL_PER; m++) {
						const int e = tid + m * CL_T;
						if (own->mark[e]) {
							const uint32_t h = own->jump[r][e];
							if (h < (uint32_t)CL_N) own->mark[h] = 1;
						}
					}
					__syncthreads();
				}
				// number the marked positions (orbit order == position order): thread tid owns positions
				// [tid * CL_PER, (tid + 1) * CL_PER)
				uint32_t mine = 0;
#pragma unroll
				for (int m = 0; m < CL_PER; m++) mine += own->mark[tid * CL_PER + m];
				uint32_t x = mine;
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
					if (lane >= (uint32_t)o) x += y;
				}
				if (lane == 31) own->wsum[wid] = x;
				__syncthreads();
				if (wid == 0) {
					const uint32_t v = lane < (uint32_t)(CL_T / 32) ? own->wsum[lane] : 0u;
					uint32_t z = v;
					for (int o = 1; o < 32; o <<= 1) {
						const uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
						if (lane >= (uint32_t)o) z += y;
					}
					if (lane < (uint32_t)(CL_T / 32)) own->wsum[lane] = z - v;  // exclusive prefix over warps
					if (lane == (uint32_t)(CL_T / 32) - 1) own->total = z;
				}
				__syncthreads();
				uint32_t run = own->wsum[wid] + x - mine;  // marked positions before tid * CL_PER
#pragma unroll
				for (int m = 0; m < CL_PER; m++) {
					own->rank[tid * CL_PER + m] = (uint16_t)run;
					run += own->mark[tid * CL_PER + m];
				}
			}
			cluster_sync_all();
			// ---- every thread: is my position on the orbit, and which read is it? ----
			const uint32_t total_marked = sh->total;
			const uint32_t mk = sh->mark[pos];
			const uint32_t rank = sh->rank[pos];
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
			// ---- advance: the last taken read's end is the next position ----
			const uint64_t ntake = n_left < (uint64_t)total_marked ? n_left : (uint64_t)total_marked;
			cluster_sync_all();  // everyone has read S, n, count, total, marks: they may be rewritten now
			if (take && (uint64_t)rank == ntake - 1) {
				const uint32_t endpos = pos + len;  // may exceed CL_N
				if (ntake == n_left) {  // region finished inside this round
					sh->S = jump_tbl(S, endpos, a.jt1, a.jt2);
					sh->draws += endpos;
					sh->p0 = 0;
				} else {
					sh->S = jump_tbl(S, (uint32_t)CL_N, a.jt1, a.jt2);
					sh->draws += CL_N;
					sh->p0 = endpos - (uint32_t)CL_N;  // >= 0: the orbit left the round here
				}
				sh->n = n_left - ntake;
				sh->count[cat] = base + ntake;
			}
			cluster_sync_all();
		}
	}
	cluster_sync_all();
	if (lead && tid == 0) {
		for (int c = 0; c < NCAT; c++) a.result[c] = own->count[c];
		a.result[4] = own->S.s1;
		a.result[5] = own->S.s2;
		a.result[6] = own->draws;
		a.result[7] = own->err;
	}
	cluster_sync_all();  // no CTA exits while its shared memory may still be read
}

void launch_resolve_replay(const ResolveArgs &a, cudaStream_t st) {
	static int mode = -1;  // QTIPSIM_RESOLVER=cta keeps the single-CTA kernel (tests compare the two)
	const char *e = getenv("QTIPSIM_RESOLVER");
	mode = (e && strcmp(e, "cta") == 0) ? 0 : 1;
	if (mode == 1) {
		static bool attr = false;
		if (!attr) {
			cudaFuncSetAttribute(k_resolve_replay_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ClusterShared));
			attr = true;
		}
		k_resolve_replay_cluster<<<CL_NC, CL_T, sizeof(ClusterShared), st>>>(a);
	} else {
		k_resolve_replay<<<1, RS_T, 0, st>>>(a);
	}
	kernels_count_launch();
}

}  // namespace qt
