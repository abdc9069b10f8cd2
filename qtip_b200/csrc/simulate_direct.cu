// simulate_direct.cu -- the production simulate+emit kernel (sm_100a): warp-autonomous tiles,
// interior chunks straight from registers to HBM.
//
// Replaces SimulatedRead::init/mutate/init_random (simplesim.h:68-117, simplesim.cpp:31-71)
// and SimulatedRead::write/write_pair (simplesim.cpp:102-171).  The reference produces a read
// one byte at a time (a branch per transcript op, an fputc per base on the '-' strand); here
// the unit of work is one aligned 16-byte chunk of the OUTPUT file.
//
// A FASTQ record is name, sequence, "\n+\n", qualities, "\n".  Cut at the 16-byte boundaries
// of the output file, nearly all of its chunks lie wholly inside the sequence or wholly inside
// the quality string ("interior" chunks); per record only a handful hold a string's first or last
// bytes next to name / separator bytes ("edge" chunks).  A warp owns a tile of RW consecutive
// records and never meets another warp at a barrier:
//
//   step A  lane = read: plan record + template -> chunk geometry (shared memory); L2 prefetch
//           of the read's packed genome lines; the read's draw ops with their reference base
//           looked up; the record's EDGE chunks composed exactly in a small per-read staging
//           slot: name, separators, first / last chunk of the sequence and of the qualities
//   step B  lane = interior chunk, G lanes per read (consecutive lanes = consecutive chunks of one
//           read => coalesced genome / quality loads and 16-byte global stores):
//             sequence   32 bits of 2-bit genome funnel-shifted from two packed words (reversed for
//                        '-' mates), expanded to 16 ASCII bases with PRMT ("ACGT" / "TGCA" as the
//                        byte table)
//             qualities  two aligned 16-byte loads of the template's pre-oriented quality string,
//                        re-aligned to the output phase (branch-free word mux)
//   step C  lane = read: the read's random draws patched over the sequence, byte stores into the
//           staging slot or HBM (REPLAY: the reference's L'Ecuyer stream stepped from the plan's
//           state, with X redraws against the reference base; PHILOX: counter-based)
//   step D  lane = edge chunk: staging slots -> HBM, 16 bytes per lane; the tile's first and last
//           chunk, shared with the neighbouring tiles, byte by byte
//
// Reads that contain N, reach past their record, or have indel segments take the general chunk
// generator (segment merge + N mask) in steps A and B.  Categories with a body shorter than 17
// bytes (edge chunks of one string could coincide) use the staged kernel in simulate_warp.cu.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"
#include "simulate_common.cuh"

namespace qt {

#ifdef QT_BOUNDS
__device__ unsigned long long g_qd_violation[4];
#define QD_CHK(cond, id, v0, v1)                                                          \
	do {                                                                                   \
		if (!(cond)) {                                                                     \
			if (atomicCAS(&g_qd_violation[0], 0ull, (unsigned long long)(id)) == 0ull) {   \
				g_qd_violation[1] = (unsigned long long)(v0);                              \
				g_qd_violation[2] = (unsigned long long)(v1);                              \
				g_qd_violation[3] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;  \
			}                                                                              \
		}                                                                                  \
	} while (0)
#else
#define QD_CHK(cond, id, v0, v1) do { } while (0)
#endif

#ifndef QT_DT_MAX_WARPS
#define QT_DT_MAX_WARPS 4
#endif
#ifndef QT_DT_MIN_BLOCKS
#define QT_DT_MIN_BLOCKS 6
#endif
#ifndef QT_DT_UNROLL
#define QT_DT_UNROLL 2
#endif
static constexpr int DT_MAX_WARPS = QT_DT_MAX_WARPS;
static constexpr int DT_OPS = 6;                 // draw ops per mate kept in shared memory from step A
static constexpr int DT_UNROLL = QT_DT_UNROLL;   // step B: chunk iterations with their loads in flight together
static constexpr uint32_t T_BYTES = 32;          // staging for the chunks of the sequence's last byte .. the qualities' first

// per-read, per-stream interior-chunk geometry (positions are relative to the tile base, which is
// the output offset of the tile's first byte rounded down to 16)
struct __align__(16) DHot {
	uint32_t sdst;    // byte position of interior sequence slot 0
	uint32_t w0;      // fast path: packed-genome word of that slot, step +1 ('+') / -1 ('-')
	uint32_t qdst;    // byte position of interior quality slot 0
	uint32_t qsrc;    // quality arena, 16-byte unit of that slot
	uint32_t nsq;     // interior sequence slots | interior quality slots << 16
	uint32_t meta;    // RF_* | genome phase << 8 | quality phase << 12
	uint32_t P;       // byte position of the first sequence byte
	uint32_t L;       // body length
};
struct __align__(16) DGen {  // general path only
	uint64_t gpos;    // genome position of the mate's leftmost reference base
	uint64_t gend;    // end of the record
	uint32_t segoff;  // into DevCategory::segs
	uint32_t nseg;
	uint32_t pad[2];
};
// what a staging slot holds: chunks [s0, s0 + nS) -- from the previous record's last quality chunk to
// this record's first sequence chunk -- and chunks [t0, t0 + nT): last sequence chunk .. first quality chunk
struct __align__(16) DShip {
	uint32_t s0, nS, t0, nT;
};

// words of a 16-byte chunk that hold at least one byte of output positions [0, L); k0 = output position
// of the chunk's first byte.  The word's other bytes belong to the name or a separator: the same lane
// writes those afterwards.
__device__ __forceinline__ void store_words(unsigned char *chunk, const uint4 &v, int k0, int L) {
	uint32_t *cp = (uint32_t *)chunk;
	const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for (int w = 0; w < 4; w++) {
		const int lo = k0 + 4 * w;
		if (lo < L && lo + 4 > 0) cp[w] = vw[w];
	}
}
// same, but a word that also holds bytes past L gets only its own bytes (the neighbouring bytes belong
// to the next record's name, which another lane writes)
__device__ __forceinline__ void store_words_exact_end(unsigned char *chunk, const uint4 &v, int k0, int L) {
	uint32_t *cp = (uint32_t *)chunk;
	const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for (int w = 0; w < 4; w++) {
		const int lo = k0 + 4 * w;
		if (lo + 4 <= L) {
			if (lo + 4 > 0) cp[w] = vw[w];
		} else {
#pragma unroll
			for (int b = 0; b < 3; b++)
				if (lo + b < L) chunk[4 * w + b] = (unsigned char)(vw[w] >> (8 * b));
		}
	}
}

// reference base at genome position gp as the reference's mutate() sees it: 0..3, or 4 for N and
// for positions past the record end
__device__ __forceinline__ uint32_t ref_code_at(const DevGenome &g, uint64_t gp, uint64_t gend, bool checked) {
	if (checked) {
		if (gp >= gend) return 4u;
		if ((__ldg(&g.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	}
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

template <int MODE, int CAT>
__global__ void __launch_bounds__(DT_MAX_WARPS * 32, QT_DT_MIN_BLOCKS) k_simulate_direct(SimArgs a, int RW, int lgG, int rounds,
                                                                                          int lgH, int hrounds, uint32_t s_bytes,
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

	const uint32_t slot_bytes = s_bytes + T_BYTES;
	unsigned char *wsm = smem + (size_t)wib * warp_smem;
	unsigned char *s_edge = wsm;                                                  // [NS][RW + 1] slots
	DHot *s_hot = (DHot *)(s_edge + (size_t)NS * (RW + 1) * slot_bytes);          // [NS][RW]
	DGen *s_gen = (DGen *)(s_hot + (size_t)NS * RW);                              // [NS][RW]
	DShip *s_ship = (DShip *)(s_gen + (size_t)NS * RW);                           // [NS][RW + 1]
	uint32_t *s_ops = (uint32_t *)(s_ship + (size_t)NS * (RW + 1));               // [NM][DT_OPS][RW]

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
	char *tb[NS];          // tile base in the output stream (16-byte aligned)
	uint32_t ph[NS], tend[NS];  // tile-relative position of the tile's first byte / one past its last
	uint32_t rstart[NS], rlen[NS];
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint64_t myoff = offs[s][ri];
		const uint64_t G1 = (r0 + nr < a.nreads) ? offs[s][r0 + nr] : *tots[s];
		const uint64_t G0 = __shfl_sync(0xffffffffu, myoff, 0);
		uint64_t nxt = __shfl_down_sync(0xffffffffu, myoff, 1);
		if (lane == nr - 1) nxt = G1;
		ph[s] = (uint32_t)(G0 & 15);
		tb[s] = outs[s] + (G0 - ph[s]);
		tend[s] = ph[s] + (uint32_t)(G1 - G0);
		rstart[s] = ph[s] + (uint32_t)(myoff - G0);
		rlen[s] = (uint32_t)(nxt - myoff);
	}
	const PlanRec pr = a.plan[ri];
	uint32_t nops[NM];
	uint32_t fwbits = 0;  // bit m: mate m is '+'
	bool s0_random = false;
	if (act) {
		const TmplDesc *tp = a.c.t + pr.tmpl;
		const uint4 *recp4 = (const uint4 *)(a.g.recs + pr.record);
		const uint4 rec_a = __ldg(recp4), rec_b = __ldg(recp4 + 1);
		const uint64_t rec_gstart = (uint64_t)rec_a.x | ((uint64_t)rec_a.y << 32);
		const uint64_t gend = rec_gstart + ((uint64_t)rec_a.z | ((uint64_t)rec_a.w << 32));
		const uint32_t rec_name_off = rec_b.x, rl = rec_b.y;
		const TmplTail tt = load_tail(tp);
		MateHot mh[NM];
#pragma unroll
		for (int m = 0; m < NM; m++) mh[m] = load_mate(&tp->m[m]);
		const uint64_t ro[2] = {pr.refoff + tt.delta[0], pr.refoff + tt.delta[1]};
		s0_random = CAT == CAT_B && !tt.aligned_is_mate1;
		// which mates need the general path: indel segments, reaching past the record end, or any N in the
		// 64-base blocks of the N-mask summary the mate touches (bases outside the mate only ever land on
		// bytes that are not stored)
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
			// the first draw ops, X ops with their reference base in place of the reference offset
			nops[m] = mh[m].nops;
			fwbits |= mh[m].fw ? (1u << m) : 0u;
#pragma unroll
			for (int j = 0; j < DT_OPS; j++)
				if (j < (int)mh[m].nops) {
					uint32_t e = __ldg(&a.c.ops[mh[m].ops_off + j]);
					if ((e & 3u) == OP_X) e = (e & 0x1ffffu) | (ref_code_at(a.g, gpos + (e >> 17), gend, general[m]) << 17);
					s_ops[(m * DT_OPS + j) * RW + lane] = e;
				}
		}
		uint32_t nl = 0;
		unsigned char *name0 = nullptr;
#pragma unroll
		for (int s = 0; s < NS; s++) {
			const int mate = CAT == CAT_B ? 0 : s;
			const bool random = CAT == CAT_B && ((s == 0) == s0_random);
			const MateHot &md = mh[mate];
			const uint32_t L = random ? tt.opp_len : md.len;
			if (s == 0) nl = rlen[0] - 2u * L - 4u;  // record = name + L + "\n+\n" + L + "\n"
			const uint32_t P = rstart[s] + nl, Pq = P + L + 3u;
			const uint64_t gpos = rec_gstart + ro[mate];
			const int c_first = (int)(P >> 4), c_last = (int)((P + L - 1u) >> 4);
			const int q_first = (int)(Pq >> 4), q_last = (int)((Pq + L - 1u) >> 4);
			const uint32_t A0 = md.qual_out_off + 16u * (uint32_t)q_first - Pq;  // arena byte of the first quality chunk, >= 1
			const uint64_t gadj = md.fw ? gpos + (uint64_t)(int64_t)md.ref0 - (P & 15u)
			                            : gpos + (uint64_t)(int64_t)md.ref0 + (P & 15u) - 15u;
			const uint32_t w0 = (uint32_t)(gadj >> 4);
			const uint32_t meta = (md.fw ? RF_FW : 0u) | (random ? RF_RANDOM : 0u) | (general[mate] ? RF_GENERAL : 0u) |
			                      ((uint32_t)(gadj & 15u) << 8) | ((A0 & 15u) << 12);
			DHot h;
			h.sdst = 16u * (uint32_t)(c_first + 1);
			h.w0 = md.fw ? w0 + 1u : w0 - 1u;
			h.qdst = 16u * (uint32_t)(q_first + 1);
			h.qsrc = (A0 >> 4) + 1u;
			h.nsq = (random ? 0u : (uint32_t)(c_last - c_first - 1)) | ((uint32_t)(q_last - q_first - 1) << 16);
			h.meta = meta;
			h.P = P;
			h.L = L;
			s_hot[s * RW + lane] = h;
			if (general[mate]) {
				DGen g;
				g.gpos = gpos;
				g.gend = gend;
				g.segoff = md.seg_off;
				g.nseg = md.nseg;
				g.pad[0] = g.pad[1] = 0;
				s_gen[s * RW + lane] = g;
			}
			// staging: this record's slot, and the head of the next slot (our last quality chunk and final '\n')
			const uint32_t s0c = lane == 0 ? 0u : (rstart[s] - 2u) >> 4;
			unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + lane) * slot_bytes;
			unsigned char *sS = slot - 16u * s0c;                       // S space: + tile-relative byte position
			unsigned char *sT = slot + s_bytes - 16u * (uint32_t)c_last;  // T space
			unsigned char *sN = slot + slot_bytes - 16u * (uint32_t)q_last;  // next slot's S space
			DShip sh;
			sh.s0 = s0c;
			sh.nS = (uint32_t)c_first - s0c + 1u;
			sh.t0 = (uint32_t)c_last;
			sh.nT = (uint32_t)(q_first - c_last) + 1u;
			s_ship[s * (RW + 1) + lane] = sh;
			QD_CHK(16u * sh.nS <= s_bytes && sh.nT <= 2u && c_last > c_first && q_last > q_first, 5, sh.nS, sh.nT);
			if (lane == nr - 1) {  // the slot after the tile's last record: its last quality chunk and '\n'
				DShip z;
				z.s0 = (uint32_t)q_last;
				z.nS = ((tend[s] - 1u) >> 4) - (uint32_t)q_last + 1u;
				z.t0 = 0;
				z.nT = 0;
				s_ship[s * (RW + 1) + nr] = z;
			}
			// edge chunks: words that hold string bytes first, then the name / separator bytes over them
			if (!random) {
				uint4 vh, vt;
				if (general[mate]) {
					vh = seq16_general(gv, s_lowmask, a.c.segs + md.seg_off, (int)md.nseg, md.fw, (int)L, gpos, gend, 16 * c_first - (int)P);
					vt = seq16_general(gv, s_lowmask, a.c.segs + md.seg_off, (int)md.nseg, md.fw, (int)L, gpos, gend, 16 * c_last - (int)P);
				} else {
					const SeqLoad lh = seq_issue(a.g.bases, w0, meta, 0), lt = seq_issue(a.g.bases, w0, meta, c_last - c_first);
					vh = seq_finish(lh, meta);
					vt = seq_finish(lt, meta);
				}
				store_words(sS + 16 * c_first, vh, 16 * c_first - (int)P, (int)L);
				store_words(sT + 16 * c_last, vt, 16 * c_last - (int)P, (int)L);
			}
			{
				uint4 vh, vt;
				if (random) vh = vt = make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u);  // 'I' (simplesim.h:105)
				else {
					const QualLoad lh = qual_issue(a.c.qual_out, A0 >> 4), lt = qual_issue(a.c.qual_out, (A0 >> 4) + (uint32_t)(q_last - q_first));
					vh = qual_finish(lh, A0 & 15u);
					vt = qual_finish(lt, A0 & 15u);
				}
				store_words(sT + 16 * q_first, vh, 16 * q_first - (int)Pq, (int)L);
				store_words_exact_end(sN + 16 * q_last, vt, 16 * q_last - (int)Pq, (int)L);
			}
			// name (simplesim.cpp:104-110, :140-150)
			unsigned char *np = sS + rstart[s];
			if (s == 0) {
				name0 = np;
				const char *refid = a.g.names + rec_name_off;
				char *p = (char *)np;
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
					p = s_origin(p, refid, rl, ro[0], nd, mh[0]);
				} else if (CAT == CAT_B) {
					const uint32_t nd = (nl - 18u - 2u * rl - 2u * mh[0].score_len - tt.typ_len) >> 1;
					p = s_origin(p, refid, rl, ro[0], nd, mh[0]);
					*p++ = ':';
					p = s_origin(p, refid, rl, ro[0], nd, mh[0]);
				} else {
					const uint32_t nd0 = ndig64(ro[0]);
					const uint32_t nd1 = nl - 18u - 2u * rl - mh[0].score_len - mh[NM - 1].score_len - tt.typ_len - nd0;
					p = s_origin(p, refid, rl, ro[0], nd0, mh[0]);
					*p++ = ':';
					p = s_origin(p, refid, rl, ro[1], nd1, mh[NM - 1]);
				}
				p[0] = ':';
				p[1] = (char)(tt.typ & 0xffu);
				if (tt.typ_len > 1) p[2] = (char)(tt.typ >> 8);
				p[1 + tt.typ_len] = '\n';
				QD_CHK(p + 2 + tt.typ_len == (char *)np + nl, 6, (uint32_t)(p - (char *)np), nl);
			} else {
				for (uint32_t j = 0; j < nl; j++) np[j] = name0[j];
			}
			// separators
			sT[P + L] = '\n';
			sT[P + L + 1u] = '+';
			sT[P + L + 2u] = '\n';
			sN[Pq + L] = '\n';
		}
	}
	__syncwarp();

	// ---------------- step B (lane = interior chunk): G lanes per read, `rounds` passes of G slots ----------------
	{
		const int G = 1 << lgG, rpi = 32 >> lgG;  // reads per iteration
		const int sub = lane >> lgG, jl = lane & (G - 1);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			char *tbs = tb[s];
			const DHot *hot = s_hot + s * RW;
			for (int rd = 0; rd < rounds; rd++) {
				const int j = rd * G + jl;
				for (int i0 = 0; i0 < nr; i0 += DT_UNROLL * rpi) {
					SeqLoad sl[DT_UNROLL];
					QualLoad ql[DT_UNROLL];
					uint4 sgen[DT_UNROLL];
					uint32_t meta[DT_UNROLL], sdst[DT_UNROLL], qdst[DT_UNROLL];
					bool slive[DT_UNROLL], qlive[DT_UNROLL];
#pragma unroll
					for (int u = 0; u < DT_UNROLL; u++) {
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
									const DHot &hh = hot[i];
									const DGen &gg = s_gen[s * RW + i];
									sgen[u] = seq16_general(gv, s_lowmask, a.c.segs + gg.segoff, (int)gg.nseg, (meta[u] & RF_FW) != 0,
									                        (int)hh.L, gg.gpos, gg.gend, (int)sdst[u] - (int)hh.P);
								} else sl[u] = seq_issue(a.g.bases, h4.y, meta[u], j);
							}
							if (qlive[u] && !(meta[u] & RF_RANDOM)) ql[u] = qual_issue(a.c.qual_out, h4.w + (uint32_t)j);
						}
					}
#pragma unroll
					for (int u = 0; u < DT_UNROLL; u++) {
						if (slive[u]) {
							QD_CHK(sdst[u] + 16u <= tend[s], 3, sdst[u], tend[s]);
							*(uint4 *)(tbs + sdst[u]) = (meta[u] & RF_GENERAL) ? sgen[u] : seq_finish(sl[u], meta[u]);
						}
						if (qlive[u]) {
							QD_CHK(qdst[u] + 16u <= tend[s], 4, qdst[u], tend[s]);
							*(uint4 *)(tbs + qdst[u]) =
							    (meta[u] & RF_RANDOM) ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)
							                          : qual_finish(ql[u], (meta[u] >> 12) & 15u);
						}
					}
				}
			}
		}
	}
	__syncwarp();

	// ---------------- step C (lane = read): draws, patched over staging slots / HBM ----------------
	if (act) {
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
		// where output position k of stream s's sequence lives: first chunk -> slot S space, last chunk ->
		// slot T space, interior -> HBM (already stored by step B: same warp, ordered by __syncwarp)
		auto put_base = [&](int s, uint32_t P, uint32_t c_first, uint32_t c_last, uint32_t s0c, uint32_t k, uint32_t ch) {
			const uint32_t x = P + k, cx = x >> 4;
			unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + lane) * slot_bytes;
			if (cx == c_first) slot[x - 16u * s0c] = (unsigned char)ch;
			else if (cx == c_last) slot[s_bytes + x - 16u * c_last] = (unsigned char)ch;
			else tb[s][x] = (char)ch;
		};
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const int s = CAT == CAT_B ? (s0_random ? 1 : 0) : m;  // stream that holds aligned mate m
			const DHot &hh = s_hot[s * RW + lane];
			const uint32_t P = hh.P, L = hh.L;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4, s0c = s_ship[s * (RW + 1) + lane].s0;
			const bool fw = (fwbits >> m) & 1u;
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			const uint32_t n = nops[m];
			for (uint32_t j = 0; j < n; j++) {
				uint32_t e;
				if (j < (uint32_t)DT_OPS) e = s_ops[(m * DT_OPS + j) * RW + lane];
				else {  // long draw-op lists: from the template, reference base looked up here
					const TmplDesc *tp = a.c.t + pr.tmpl;
					e = __ldg(&a.c.ops[tp->m[m].ops_off + j]);
					if ((e & 3u) == OP_X) {
						const DGen &gg = s_gen[s * RW + lane];
						const bool gen = (hh.meta & RF_GENERAL) != 0;
						uint64_t gpos, gend = 0;
						if (gen) {
							gpos = gg.gpos;
							gend = gg.gend;
						} else {
							const RecDev &rc = a.g.recs[pr.record];
							gpos = rc.gstart + pr.refoff + tp->delta[m];
						}
						e = (e & 0x1ffffu) | (ref_code_at(a.g, gpos + (e >> 17), gend, gen) << 17);
					}
				}
				const uint32_t outp = (e >> 2) & 0x7fffu;
				QD_CHK(outp < L, 7, outp, L);
				uint32_t code;
				if ((e & 3u) == OP_X) {
					const uint32_t ref = e >> 17;
					if (MODE == QTIPSIM_RNG_REPLAY) {
						do code = draw_code();  // simplesim.cpp:47-49
						while (code == ref);
					} else {
						const uint32_t b = draw16();
						code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
					}
				} else code = draw_code();
				put_base(s, P, c_first, c_last, s0c, outp, (T >> (8u * code)) & 0xffu);
			}
		}
		if (CAT == CAT_B) {
			// init_random (simplesim.h:89-117): opp_len draws, printed with the aligned mate's strand
			const int s = s0_random ? 0 : 1;
			const DHot &hh = s_hot[s * RW + lane];
			const uint32_t P = hh.P, L = hh.L;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4, s0c = s_ship[s * (RW + 1) + lane].s0;
			const bool fw = fwbits & 1u;
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			for (uint32_t j = 0; j < L; j++) {
				const uint32_t code = draw_code();
				put_base(s, P, c_first, c_last, s0c, fw ? j : L - 1 - j, (T >> (8u * code)) & 0xffu);
			}
		}
	}
	__syncwarp();

	// ---------------- step D (lane = edge chunk): staging slots -> HBM ----------------
	{
		const int H = 1 << lgH, spi = 32 >> lgH;  // slots per iteration
		const int sub = lane >> lgH, jl = lane & (H - 1);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			char *tbs = tb[s];
			const uint32_t end = tend[s], phs = ph[s];
			for (int rd = 0; rd < hrounds; rd++) {
				const uint32_t j = (uint32_t)(rd * H + jl);
				for (int i0 = 0; i0 <= nr; i0 += spi) {
					const int i = i0 + sub;
					if (i > nr) continue;
					const DShip sh = s_ship[s * (RW + 1) + i];
					const unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + i) * slot_bytes;
					uint32_t chunk;
					const unsigned char *src;
					if (j < sh.nS) {
						chunk = sh.s0 + j;
						src = slot + 16u * j;
					} else if (j < sh.nS + sh.nT) {
						chunk = sh.t0 + (j - sh.nS);
						src = slot + s_bytes + 16u * (j - sh.nS);
					} else continue;
					char *dst = tbs + 16u * (size_t)chunk;
					if (chunk == 0 && phs) {  // shared with the previous tile
						for (uint32_t b = phs; b < 16u; b++) dst[b] = (char)src[b];
					} else if (chunk == (end >> 4)) {  // shared with the next tile (end & 15 != 0, or nothing to ship)
						for (uint32_t b = 0; b < (end & 15u); b++) dst[b] = (char)src[b];
					} else *(uint4 *)dst = *(const uint4 *)src;
				}
			}
		}
	}
}

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_name, uint32_t max_body, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;
	static int env_rw = -1, env_warps = -1;
	if (env_rw < 0) {
		const char *e = getenv("QTIPSIM_RW");
		env_rw = e ? atoi(e) : 0;
		e = getenv("QTIPSIM_WARPS");
		env_warps = e ? atoi(e) : 0;
	}
	// S part of a slot: from the chunk of (record start - 2) to the chunk of the first sequence byte
	const uint32_t s_bytes = ((max_name + 2u + 15u + 15u) & ~15u) + 16u;
	auto warp_bytes = [&](int rw) -> uint32_t {
		return NS * (uint32_t)(rw + 1) * (s_bytes + T_BYTES + (uint32_t)sizeof(DShip)) +
		       (uint32_t)rw * (NS * (uint32_t)(sizeof(DHot) + sizeof(DGen)) + NM * DT_OPS * 4u);
	};
	const uint32_t sm_budget = 200u * 1024u;
	int rw = env_rw > 0 ? env_rw : 32;
	if (rw > 32) rw = 32;
	if (env_rw <= 0)
		while (rw > 4 && warp_bytes(rw) * 16u > sm_budget) rw >>= 1;
	if (warp_bytes(rw) + 1024 > sm_budget) return false;
	int warps = env_warps > 0 ? env_warps : DT_MAX_WARPS;
	if (warps > DT_MAX_WARPS) warps = DT_MAX_WARPS;
	while (warps > 1 && warp_bytes(rw) * (uint32_t)warps + 1024 > sm_budget) warps >>= 1;
	const uint32_t wb = warp_bytes(rw);
	const size_t smem = (size_t)wb * warps;
	// step B: interior chunks of a string of max_body bytes at any alignment: at most floor((max_body + 14) / 16) - 1
	int slots = (int)((max_body + 14u) / 16u) - 1;
	if (slots < 1) slots = 1;
	int lgG = 0;
	while ((1 << lgG) < slots && lgG < 5) lgG++;
	const int rounds = (slots + (1 << lgG) - 1) >> lgG;
	// step D: chunks per staging slot
	const int hslots = (int)(s_bytes / 16u) + 2;
	int lgH = 0;
	while ((1 << lgH) < hslots && lgH < 5) lgH++;
	const int hrounds = (hslots + (1 << lgH) - 1) >> lgH;
	auto kern = k_simulate_direct<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + rw - 1) / rw;
	const uint64_t nblocks = (ntiles + warps - 1) / warps;
	kern<<<(unsigned)nblocks, warps * 32, smem, st>>>(a, rw, lgG, rounds, lgH, hrounds, s_bytes, wb);
	kernels_count_launch();
	return true;
}

template <int MODE>
static bool launch_mode(const SimArgs &a, uint32_t max_name, uint32_t max_body, cudaStream_t st) {
	switch (a.cat) {
	case CAT_U: return launch_cat<MODE, CAT_U>(a, max_name, max_body, st);
	case CAT_B: return launch_cat<MODE, CAT_B>(a, max_name, max_body, st);
	case CAT_C: return launch_cat<MODE, CAT_C>(a, max_name, max_body, st);
	default: return launch_cat<MODE, CAT_D>(a, max_name, max_body, st);
	}
}

// max_name: upper bound on the read name (with '@' and '\n'); min_body / max_body: shortest / longest body
// of the category.  Returns false when the category cannot use this kernel (a body under 17 bytes, or
// names too long for the staging slots): the caller falls back to the staged kernel.
bool launch_simulate_direct(const SimArgs &a, uint32_t max_name, uint32_t min_body, uint32_t max_body, cudaStream_t st) {
	if (!a.nreads) return true;
	if (min_body < 17u) return false;
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) return launch_mode<QTIPSIM_RNG_REPLAY>(a, max_name, max_body, st);
	return launch_mode<QTIPSIM_RNG_PHILOX>(a, max_name, max_body, st);
}

#ifdef QT_BOUNDS
void direct_debug_report() {
	unsigned long long v[4] = {0, 0, 0, 0};
	cudaMemcpyFromSymbol(v, g_qd_violation, sizeof(v));
	if (v[0])
		fprintf(stderr, "QT_BOUNDS(direct) violation id=%llu v0=%lld v1=%lld block=%llu thread=%llu\n", v[0], (long long)v[1],
		        (long long)v[2], v[3] >> 32, v[3] & 0xffffffffu);
	else fprintf(stderr, "QT_BOUNDS(direct): no violation\n");
	memset(v, 0, sizeof(v));
	cudaMemcpyToSymbol(g_qd_violation, v, sizeof(v));
}
#endif

}  // namespace qt

#ifdef QT_BOUNDS
extern "C" void qt_direct_debug_report() { qt::direct_debug_report(); }
#endif
