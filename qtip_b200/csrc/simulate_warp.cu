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
#include "simulate_common.cuh"

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

#ifndef QT_WT_MAX_WARPS
#define QT_WT_MAX_WARPS 4
#endif
#ifndef QT_WT_MIN_BLOCKS
#define QT_WT_MIN_BLOCKS 5
#endif
#ifndef QT_UNROLL_B
#define QT_UNROLL_B 2
#endif
static constexpr int WT_MAX_WARPS = QT_WT_MAX_WARPS;
static constexpr int OPS_SMEM = 6;          // draw ops per mate kept in shared memory from step A
static constexpr int UNROLL_B = QT_UNROLL_B;  // step B: chunk iterations with their loads in flight together

// per-read, per-stream chunk geometry, shared memory: everything a chunk lane needs, precomputed by
// the read's lane in step A so that the per-chunk work is an add and a compare
struct __align__(16) RdHot {
	uint32_t sdst;    // staging offset of sequence chunk slot 0 (16 * (P >> 4))
	uint32_t w0;      // fast path: packed-genome word of slot 0, step +1 ('+') / -1 ('-')
	uint32_t qdst;    // staging offset of quality chunk slot 0
	uint32_t qsrc;    // quality arena, 16-byte unit of slot 0
	uint32_t nsq;     // sequence slots stored in step B (all chunks but the last) | quality slots << 16
	uint32_t meta;    // RF_* | genome phase << 8 | quality phase << 12
	uint32_t P;       // staging byte position of the first sequence byte
	uint32_t L;       // body length
};
// general path only
struct __align__(16) RdGen {
	uint64_t gpos;    // genome position of the mate's leftmost reference base
	uint64_t gend;    // end of the record
	uint32_t segoff;  // into DevCategory::segs
	uint32_t nseg;
	uint32_t pad[2];
};

template <int MODE, int CAT>
__global__ void __launch_bounds__(WT_MAX_WARPS * 32, QT_WT_MIN_BLOCKS) k_simulate_warp(SimArgs a, int RW, int lgG, int rounds,
                                                                      uint32_t stage_stride, uint32_t warp_smem) {
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
	RdGen *s_gen = (RdGen *)(s_hot + (size_t)NS * RW);                // [NS][RW]
	uint32_t *s_ops = (uint32_t *)(s_gen + (size_t)NS * RW);          // [NM][OPS_SMEM][RW]

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
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint64_t myoff = offs[s][ri];
		G1[s] = (r0 + nr < a.nreads) ? offs[s][r0 + nr] : *tots[s];
		G0[s] = __shfl_sync(0xffffffffu, myoff, 0);
		uint64_t nxt = __shfl_down_sync(0xffffffffu, myoff, 1);
		if (lane == nr - 1) nxt = G1[s];
		Pst[s] = (uint32_t)(G0[s] & 15) + (uint32_t)(myoff - G0[s]);  // record start; + nl below
		Lst[s] = (uint32_t)(nxt - myoff);                             // record length; body length below
	}
	const PlanRec pr = a.plan[ri];
	const TmplDesc *tp = a.c.t + pr.tmpl;
	const uint4 *recp4 = (const uint4 *)(a.g.recs + pr.record);
	const uint4 rec_a = __ldg(recp4), rec_b = __ldg(recp4 + 1), rec_c = __ldg(recp4 + 2);
	const uint64_t rec_gstart = (uint64_t)rec_a.x | ((uint64_t)rec_a.y << 32);
	const uint64_t gend = rec_gstart + ((uint64_t)rec_a.z | ((uint64_t)rec_a.w << 32));
	const uint32_t rec_name_off = rec_b.x, rl = rec_b.y;
	const TmplTail tt = load_tail(tp);
	MateHot mh[NM];
#pragma unroll
	for (int m = 0; m < NM; m++) mh[m] = load_mate(&tp->m[m]);
	const uint64_t ro[2] = {pr.refoff + tt.delta[0], pr.refoff + tt.delta[1]};
	uint4 tailv[NS];
	bool has_seq[NS];
	if (act) {
		// which mates need the general path: indel segments, reaching past the record end, or any N in the
		// 64-base blocks of the N-mask summary the mate touches (bases outside the mate only ever land on
		// garbage bytes of whole-chunk stores)
		bool general[NM];
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const uint64_t gpos = rec_gstart + ro[m];
			const uint64_t b0 = gpos >> 6, b1 = (gpos + (mh[m].rflen ? mh[m].rflen - 1u : 0u)) >> 6;  // blocks [b0, b1]
			bool anyn;
			if (b1 - b0 < 32) {  // the usual case: the blocks span at most two summary words
				const uint32_t lo = __ldg(&a.g.nblk[b0 >> 5]), hi = __ldg(&a.g.nblk[(b0 >> 5) + 1]);
				const uint32_t bits = __funnelshift_r(lo, hi, (uint32_t)(b0 & 31));
				anyn = (bits & (0xffffffffu >> (31u - (uint32_t)(b1 - b0)))) != 0;
			} else {
				anyn = false;
				for (uint64_t b = b0; b <= b1; b++) anyn |= ((__ldg(&a.g.nblk[b >> 5]) >> (b & 31)) & 1u) != 0;
			}
			general[m] = anyn || !mh[m].single || (gpos + mh[m].rflen > gend);
			// L2 prefetch of the packed genome lines the chunk lanes are about to read
			const uintptr_t gl = ((uintptr_t)a.g.bases + ((gpos - 16) >> 2)) & ~(uintptr_t)127;
			const uintptr_t gh = (uintptr_t)a.g.bases + ((gpos + mh[m].rflen + 32) >> 2);
			for (uintptr_t p = gl; p <= gh; p += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#pragma unroll
			for (int j = 0; j < OPS_SMEM; j++)
				if (j < (int)mh[m].nops) s_ops[(m * OPS_SMEM + j) * RW + lane] = __ldg(&a.c.ops[mh[m].ops_off + j]);
		}
#pragma unroll
		for (int s = 0; s < NS; s++) {
			const int mate = CAT == CAT_B ? 0 : s;
			const bool random = CAT == CAT_B && ((s == 0) != tt.aligned_is_mate1);
			const MateHot &md = mh[mate];
			const uint32_t L = random ? tt.opp_len : md.len;
			if (s == 0) nl = Lst[0] - 2u * L - 4u;  // record = name + L + "\n+\n" + L + "\n"
			const uint32_t P = Pst[s] + nl;
			QW_CHK(P + 2u * L + 4u + 16u <= stage_stride, 5, P, L);
			const uint64_t gpos = rec_gstart + ro[mate];
			const uint32_t Pq = P + L + 3u;
			const int c_first = (int)(P >> 4), c_last = L ? (int)((P + L - 1u) >> 4) : c_first;
			const int q_first = (int)(Pq >> 4), q_last = L ? (int)((Pq + L - 1u) >> 4) : q_first - 1;
			const uint32_t A0 = md.qual_out_off + 16u * (uint32_t)q_first - Pq;  // arena byte of quality slot 0, >= 1
			const uint64_t gadj = md.fw ? gpos + (uint64_t)(int64_t)md.ref0 - (P & 15u)
			                            : gpos + (uint64_t)(int64_t)md.ref0 + (P & 15u) - 15u;
			RdHot h;
			h.sdst = 16u * (uint32_t)c_first;
			h.w0 = (uint32_t)(gadj >> 4);
			h.qdst = 16u * (uint32_t)q_first;
			h.qsrc = A0 >> 4;
			h.nsq = (random ? 0u : (uint32_t)(c_last - c_first)) | ((uint32_t)(L ? q_last - q_first : 0) << 16);
			h.meta = (md.fw ? RF_FW : 0u) | (random ? RF_RANDOM : 0u) | (general[mate] ? RF_GENERAL : 0u) |
			         ((uint32_t)(gadj & 15u) << 8) | ((A0 & 15u) << 12);
			h.P = P;
			h.L = L;
			s_hot[s * RW + lane] = h;
			if (general[mate]) {
				RdGen g;
				g.gpos = gpos;
				g.gend = gend;
				g.segoff = md.seg_off;
				g.nseg = md.nseg;
				g.pad[0] = g.pad[1] = 0;
				s_gen[s * RW + lane] = g;
			}
			Pst[s] = P;
			Lst[s] = L;
			// the read's last quality chunk, whole (its garbage lands on the '\n' and the next name)
			if (L) {
				const uint4 v = random ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)
				                       : qual_finish(qual_issue(a.c.qual_out, h.qsrc + (uint32_t)(q_last - q_first)), A0 & 15u);
				*(uint4 *)(stage + (size_t)s * stage_stride + 16u * (uint32_t)q_last) = v;
			}
			// the read's last sequence chunk (loads issued now, stored in step C)
			has_seq[s] = !random && L > 0;
			tailv[s] = make_uint4(0u, 0u, 0u, 0u);
			if (has_seq[s]) {
				if (general[mate])
					tailv[s] = seq16_general(gv, s_lowmask, a.c.segs + md.seg_off, (int)md.nseg, md.fw, (int)L, gpos, gend,
					                         16 * c_last - (int)P);
				else tailv[s] = seq_finish(seq_issue(a.g.bases, h.w0, h.meta, c_last - c_first), h.meta);
			}
		}
	}
	__syncwarp();

	// ---------------- step B (lane = chunk): groups of G lanes per read, `rounds` passes of G slots ----------------
	{
		const int G = 1 << lgG, rpi = 32 >> lgG;  // reads per iteration
		const int sub = lane >> lgG, jl = lane & (G - 1);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			unsigned char *stg = stage + (size_t)s * stage_stride;
			const RdHot *hot = s_hot + s * RW;
			for (int rd = 0; rd < rounds; rd++) {
				const int j = rd * G + jl;
				for (int i0 = 0; i0 < nr; i0 += UNROLL_B * rpi) {
					SeqLoad sl[UNROLL_B];
					QualLoad ql[UNROLL_B];
					uint4 sgen[UNROLL_B];
					uint32_t meta[UNROLL_B], sdst[UNROLL_B], qdst[UNROLL_B];
					bool slive[UNROLL_B], qlive[UNROLL_B];
#pragma unroll
					for (int u = 0; u < UNROLL_B; u++) {
						const int i = i0 + u * rpi + sub;
						slive[u] = qlive[u] = false;
						meta[u] = 0;
						if (i < nr) {
							const uint4 h4 = *(const uint4 *)&hot[i];          // sdst | w0 | qdst | qsrc
							const uint2 h2 = *((const uint2 *)&hot[i] + 2);    // nsq | meta
							meta[u] = h2.y;
							sdst[u] = h4.x + 16u * (uint32_t)j;
							qdst[u] = h4.z + 16u * (uint32_t)j;
							slive[u] = j < (int)(h2.x & 0xffffu);
							qlive[u] = j < (int)(h2.x >> 16);
							if (slive[u]) {
								if (meta[u] & RF_GENERAL) {
									const RdHot &hh = hot[i];
									const RdGen &gg = s_gen[s * RW + i];
									sgen[u] = seq16_general(gv, s_lowmask, a.c.segs + gg.segoff, (int)gg.nseg, (meta[u] & RF_FW) != 0,
									                        (int)hh.L, gg.gpos, gg.gend, (int)sdst[u] - (int)hh.P);
								} else sl[u] = seq_issue(a.g.bases, h4.y, meta[u], j);
							}
							if (qlive[u] && !(meta[u] & RF_RANDOM)) ql[u] = qual_issue(a.c.qual_out, h4.w + (uint32_t)j);
						}
					}
#pragma unroll
					for (int u = 0; u < UNROLL_B; u++) {
						if (slive[u]) {
							QW_CHK(sdst[u] + 16u <= stage_stride, 3, sdst[u], u);
							*(uint4 *)(stg + sdst[u]) = (meta[u] & RF_GENERAL) ? sgen[u] : seq_finish(sl[u], meta[u]);
						}
						if (qlive[u]) {
							QW_CHK(qdst[u] + 16u <= stage_stride, 4, qdst[u], u);
							*(uint4 *)(stg + qdst[u]) =
							    (meta[u] & RF_RANDOM) ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)  // 'I' (simplesim.h:105)
							                          : qual_finish(ql[u], (meta[u] >> 12) & 15u);
						}
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
				// every word that holds a sequence byte; its other bytes are name or separator bytes, which
				// this lane writes next (never quality bytes: the separator is three bytes long)
				const uint32_t P = Pst[s], L = Lst[s];
				const int c_last = (int)((P + L - 1u) >> 4);
				const int k0 = 16 * c_last - (int)P;
				const uint32_t vw[4] = {tailv[s].x, tailv[s].y, tailv[s].z, tailv[s].w};
				uint32_t *cp = (uint32_t *)(stg + 16 * (size_t)c_last);
#pragma unroll
				for (int w = 0; w < 4; w++) {
					const int lo = k0 + 4 * w;  // output position of the word's first byte
					if (lo < (int)L && lo + 4 > 0) cp[w] = vw[w];
				}
			}
		}
		// name (simplesim.cpp:104-110, :140-150) and separators
		{
			const char *refid = a.g.names + rec_name_off;
			char *p = recp[0];
			p[0] = '@';
			p[1] = 'q';
			p[2] = 't';
			p[3] = 'i';
			p[4] = 'p';
			p[5] = '!';
			p[6] = ':';
			p += 7;
			if (CAT == CAT_U) {
				const uint32_t nd = nl - 13u - rl - mh[0].score_len - tt.typ_len;
				p = s_origin(p, refid, rec_c, rl, ro[0], nd, mh[0]);
			} else if (CAT == CAT_B) {
				const uint32_t nd = (nl - 18u - 2u * rl - 2u * mh[0].score_len - tt.typ_len) >> 1;
				p = s_origin(p, refid, rec_c, rl, ro[0], nd, mh[0]);
				*p++ = ':';
				p = s_origin(p, refid, rec_c, rl, ro[0], nd, mh[0]);
			} else {
				const uint32_t nd0 = ndig64(ro[0]);
				const uint32_t nd1 = nl - 18u - 2u * rl - mh[0].score_len - mh[NM - 1].score_len - tt.typ_len - nd0;
				p = s_origin(p, refid, rec_c, rl, ro[0], nd0, mh[0]);
				*p++ = ':';
				p = s_origin(p, refid, rec_c, rl, ro[1], nd1, mh[NM - 1]);
			}
			p[0] = ':';
			p[1] = (char)(tt.typ & 0xffu);
			if (tt.typ_len > 1) p[2] = (char)(tt.typ >> 8);
			p[1 + tt.typ_len] = '\n';
			QW_CHK(p + 2 + tt.typ_len == recp[0] + nl, 6, (uint32_t)(p - recp[0]), nl);
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
			const bool s0_random = CAT == CAT_B && !tt.aligned_is_mate1;
#pragma unroll
			for (int m = 0; m < NM; m++) {
				const int s = CAT == CAT_B ? (s0_random ? 1 : 0) : m;  // stream that holds aligned mate m
				char *seq = (char *)stage + (size_t)s * stage_stride + Pst[s];
				const bool fw = mh[m].fw;
				const uint32_t T = fw ? TBL_FW : TBL_RC;
				const uint32_t *ops = a.c.ops + mh[m].ops_off;
				const uint32_t nops = mh[m].nops;
				for (uint32_t j = 0; j < nops; j++) {
					const uint32_t e = j < (uint32_t)OPS_SMEM ? s_ops[(m * OPS_SMEM + j) * RW + lane] : __ldg(&ops[j]);
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
				const bool fw = mh[0].fw;
				const uint32_t T = fw ? TBL_FW : TBL_RC;
				Philox4 rb{};
				for (uint32_t j = 0; j < L; j++) {
					uint32_t code;
					if (MODE == QTIPSIM_RNG_REPLAY) code = draw_code();
					else {  // two bits per base, 64 bases per block (philox.h)
						if ((j & 63u) == 0u) rb = philox_rmate_block(gi, j >> 6, CAT, a.philox_seed);
						code = philox_rmate_code(rb, j);
					}
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

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;
	static int env_rw = -1, env_warps = -1, env_kb = -1, env_minw = -1;
	if (env_rw < 0) {
		const char *e = getenv("QTIPSIM_RW");
		env_rw = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_WARPS");
		env_warps = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_SMEM_KB");
		env_kb = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_MIN_WARPS");
		env_minw = e ? atoi(e) : 0;
	}
	auto warp_bytes = [&](int rw) -> uint32_t {
		const uint32_t stage_stride = ((uint32_t)rw * max_rec_bytes + 48 + 15) & ~15u;
		return NS * stage_stride + (uint32_t)rw * (NS * (uint32_t)(sizeof(RdHot) + sizeof(RdGen)) + NM * OPS_SMEM * 4u);
	};
	// reads per warp tile: 32 (full warps in the lane = read steps) as long as `min_warps` warps stay
	// resident per SM for latency hiding; fewer reads per tile for long records
	const uint32_t sm_budget = (env_kb > 0 ? (uint32_t)env_kb : 216u) * 1024u;
	const uint32_t min_warps = env_minw > 0 ? (uint32_t)env_minw : 12u;
	int rw = env_rw > 0 ? env_rw : 32;
	if (rw > 32) rw = 32;
	if (env_rw <= 0)
		while (rw > 4 && warp_bytes(rw) * min_warps > sm_budget) rw >>= 1;
	if (warp_bytes(rw) + 1024 > sm_budget) return false;
	int warps = env_warps > 0 ? env_warps : 2;
	if (warps > WT_MAX_WARPS) warps = WT_MAX_WARPS;
	while (warps > 1 && warp_bytes(rw) * (uint32_t)warps + 1024 > sm_budget) warps >>= 1;
	const uint32_t stage_stride = ((uint32_t)rw * max_rec_bytes + 48 + 15) & ~15u;
	const uint32_t wb = warp_bytes(rw);
	const size_t smem = (size_t)wb * warps;
	// step B: all chunks but the last of a string of max_body bytes, at any alignment -> at most `slots`
	// = floor((max_body + 14) / 16);
	// G lanes per read (a power of two), `rounds` passes
	const int slots = (int)((max_body + 14u) / 16u) > 0 ? (int)((max_body + 14u) / 16u) : 1;
	int lgG = 0;
	while ((1 << lgG) < slots && lgG < 5) lgG++;
	const int rounds = (slots + (1 << lgG) - 1) >> lgG;
	auto kern = k_simulate_warp<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + rw - 1) / rw;
	const uint64_t nblocks = (ntiles + warps - 1) / warps;
	kern<<<(unsigned)nblocks, warps * 32, smem, st>>>(a, rw, lgG, rounds, stage_stride, wb);
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
