// simulate_direct.cu -- the FIRST GENERATION of the simulate+emit kernel (sm_100a): warp-autonomous tiles,
// interior chunks straight from registers to HBM.  simulate_tile.cu (one edge image per record, names appended as
// 16-byte pieces, flat shipping loop) replaced it as the default; this one takes the categories whose tiles do not fit
// the tile kernel's budget (long names or reads, records over 4 Gbp) and stays selectable with QTIPSIM_KERNEL=direct.
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
//   step A   lane = read: plan record + template (16-byte descriptor blocks) -> chunk geometry in
//            shared memory; the N-block summary decides fast / general path per mate
//   step A2  lane = 16-byte piece: each read's packed genome span -> its row in shared memory,
//            cp.async, coalesced over the tile
//   step A3  lane = read, while those copies fly: the read's first draw ops, its name and the
//            separators into the slot IMAGE (the staged picture of the record's edge chunks)
//   step Bq  lane = quality chunk (item = read * K + slot, read-major: consecutive lanes = consecutive
//            chunks of one read => coalesced loads and 16-byte stores): two aligned 16-byte loads of
//            the template's pre-oriented quality string, re-aligned to the output phase by a
//            branch-free word mux
//   step Bs  lane = sequence chunk: two words of the row funnel-shifted to 32 bits of 2-bit genome,
//            expanded to 16 ASCII bases with PRMT; strand = byte table ("ACGT" / "TGCA") + reversed
//            interleave selectors after one rotate.  General-path reads (N, past the record end, indel
//            segments) take a second loop with the segment-merging generator, only in tiles that have any.
//            In both B steps interior chunks go from registers straight to HBM (streaming 16-byte
//            stores); a string's first / last chunk goes whole into one of the read's four AREAS
//   step C   lane = read (two lanes per read or pair in the native mode): the areas' in-string bytes
//            overlaid onto the image, then the read's draws patched over image or HBM bytes (REPLAY:
//            the reference's L'Ecuyer stream stepped from the plan's state, with X redraws against the
//            reference base read from the row; PHILOX: addressed by draw index)
//   step Br  bad-end pairs only, lane = chunk of the random mate: init_random's draws, 16 per lane
//            (Philox by index; L'Ecuyer by jump table, backwards for '-' mates); step C2 merges its edges
//   step D   lane = edge chunk: slot images -> HBM, 16 bytes per lane; the tile's first and last
//            chunk, shared with the neighbouring tiles, one byte per lane
//
// Categories with a body shorter than 17 bytes (the edge chunks of one string could coincide) use the
// staged kernel in simulate_warp.cu; records too large for either use the generic kernel (kernels.cu).
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
#define QT_DT_MIN_BLOCKS 8
#endif
#ifndef QT_DT_MIN_BLOCKS_PAIR
#define QT_DT_MIN_BLOCKS_PAIR 6
#endif
#ifndef QT_DT_UNROLL
#define QT_DT_UNROLL 2
#endif
// 16-byte stores of FASTQ chunks to HBM: written once, never read back by the kernel
// (streaming, evict-first: they should not push templates and quality strings out of L2)
#ifndef QT_DT_NO_STCS
#define ST16(ptr, v) __stcs((uint4 *)(ptr), (v))
#else
#define ST16(ptr, v) (*(uint4 *)(ptr) = (v))
#endif
static constexpr int DT_MAX_WARPS = QT_DT_MAX_WARPS;
static constexpr int DT_OPS = 6;                 // draw ops per mate kept in shared memory from step A
static constexpr int DT_UNROLL = QT_DT_UNROLL;   // step B: chunk iterations with their loads in flight together
static constexpr uint32_t T_BYTES = 32;          // staging for the chunks of the sequence's last byte .. the qualities' first

// per-read, per-stream chunk geometry (positions are relative to the tile base, which is the output
// offset of the tile's first byte rounded down to 16); slot j of a string = its j-th 16-byte chunk
struct __align__(16) DHot {   // what a sequence-chunk lane reads
	uint32_t grow;    // fast path: shared-memory address of slot 0's first packed genome word; slot j is
	                  // 4 j bytes further ('+' mates) or back ('-' mates)
	uint32_t sdst;    // byte position of sequence slot 0 (16 * (P >> 4))
	uint32_t ns;      // sequence slots | meta << 16; meta = RF_* | genome phase << 8 | quality phase << 12
	uint32_t area;    // shared-memory address of the read's four edge areas
};
struct __align__(16) DHotQ {  // what a quality-chunk lane reads
	uint32_t qsrc;    // quality arena, 16-byte unit of slot 0
	uint32_t qdst;    // byte position of quality slot 0
	uint32_t nq;      // quality slots | meta << 16
	uint32_t area;
};
// bad-end pairs: what the chunk lanes of step Br need to draw a read's random mate
struct __align__(8) DRnd {
	uint32_t r0, r1;  // REPLAY: generator state before the mate's first draw; PHILOX: index of its first draw, 0
	uint32_t flags;   // bit 0: stream of the random mate, bit 1: printed on the '+' strand
	uint32_t pad;
};
struct __align__(8) DPos {
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
// A staging slot holds the IMAGE (names, separators) of the record's edge chunks: chunks [s0, s0 + nS) --
// from the previous record's last quality chunk to this record's first sequence chunk -- and chunks
// [t0, t0 + nT): last sequence chunk .. first quality chunk.  The string bytes of those chunks are staged
// whole in the read's four AREAS (first / last chunk of the sequence, first / last chunk of the qualities)
// and merged over the image when the slot is shipped.
struct __align__(16) DShip {
	uint32_t s0, t0;
	uint32_t n;       // nS | nT << 8 | (1 << 16 when the slot is the one after the tile's last record)
	uint32_t rstart;  // byte position of the record's '@'
};
enum { AREA_SEQ_HEAD = 0, AREA_SEQ_TAIL = 1, AREA_QUAL_HEAD = 2, AREA_QUAL_TAIL = 3 };
// Per-read strides in shared memory are odd multiples of 16 bytes: when lane = read, the 32 lanes' accesses then
// spread over all eight 16-byte bank groups instead of piling onto two or four of them.
static constexpr int AREA_STRIDE = 5;  // uint4 per read (four areas + one of padding)

// reference base at genome position gp as the reference's mutate() sees it: 0..3, or 4 for N and
// for positions past the record end; general-path reads only
__device__ __forceinline__ uint32_t ref_code_at(const DevGenome &g, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;
	if ((__ldg(&g.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// out = out with bytes [lo, hi) of the chunk replaced by v's
__device__ __forceinline__ void overlay(uint4 &out, const uint4 &v, const uint4 *lowmask, int lo, int hi) {
	const uint4 m = range_mask(lowmask, lo, hi);
	out.x = (out.x & ~m.x) | (v.x & m.x);
	out.y = (out.y & ~m.y) | (v.y & m.y);
	out.z = (out.z & ~m.z) | (v.z & m.z);
	out.w = (out.w & ~m.w) | (v.w & m.w);
}

// shared memory through 32-bit shared-window addresses: one instruction per access, no generic-address arithmetic
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MODE, int CAT>
__global__ void __launch_bounds__(DT_MAX_WARPS * 32, CAT == CAT_U ? QT_DT_MIN_BLOCKS : QT_DT_MIN_BLOCKS_PAIR) k_simulate_direct(SimArgs a, int RW, int K, uint32_t kmagic,
                                                                                          int PMAX, uint32_t pmagic, int lgH,
                                                                                          int hrounds, uint32_t s_bytes,
                                                                                          uint32_t warp_smem, int skip) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;                    // output streams
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;  // aligned mates per read
	extern __shared__ __align__(16) unsigned char smem[];

	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	__shared__ uint4 s_lowmask[17];
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
	const int GROW = 4 * PMAX;  // words per s_gbuf row
	unsigned char *wsm = smem + (size_t)wib * warp_smem;
	unsigned char *s_edge = wsm;                                                  // [NS][RW + 1] slot images
	uint4 *s_area = (uint4 *)(s_edge + (size_t)NS * (RW + 1) * slot_bytes);       // [NS][RW][AREA_STRIDE]
	uint32_t *s_gbuf = (uint32_t *)(s_area + (size_t)NS * RW * AREA_STRIDE);                // [NS][RW][GROW] packed genome
	DHot *s_hot = (DHot *)(s_gbuf + (size_t)NS * RW * GROW);                      // [NS][RW]
	DHotQ *s_hotq = (DHotQ *)(s_hot + (size_t)NS * RW);                           // [NS][RW]
	DGen *s_gen = (DGen *)(s_hotq + (size_t)NS * RW);                             // [NS][RW]
	DShip *s_ship = (DShip *)(s_gen + (size_t)NS * RW);                           // [NS][RW + 1]
	uint32_t *s_ops = (uint32_t *)(s_ship + (size_t)NS * (RW + 1));               // [NM][DT_OPS][RW]
	uint32_t *s_gunit = s_ops + (size_t)NM * DT_OPS * RW;                         // [NS][RW] 16-byte unit of bases[] at the row's start
	uint32_t *s_gb0 = s_gunit + (size_t)NS * RW;                                  // [NS][RW]
	DPos *s_pos = (DPos *)(s_gb0 + (size_t)NS * RW);                              // [NS][RW]
	DRnd *s_rnd = (DRnd *)(s_pos + (size_t)NS * RW);                              // [RW] (bad-end pairs only)

	const uint64_t *offs[2] = {a.off0, a.off1};
	const uint64_t *tots[2] = {a.tot0, a.tot1};
	char *outs[2] = {a.out0, a.out1};
	GenomeView gv;
	gv.bases = a.g.bases;
	gv.nmask = a.g.nmask;
	gv.max_word = a.genome_words - 2;

	const bool act = lane < nr;
	const uint64_t ri = r0 + (uint64_t)(act ? lane : 0);

	// ---------------- step A (lane = read): geometry ----------------
	char *tb[NS];               // tile base in the output stream (16-byte aligned)
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
	const TmplDesc *tp = a.c.t + pr.tmpl;
	const uint4 *recp4 = (const uint4 *)(a.g.recs + pr.record);
	const uint4 rec_a = __ldg(recp4), rec_b = __ldg(recp4 + 1), rec_c = __ldg(recp4 + 2);
	const uint64_t rec_gstart = (uint64_t)rec_a.x | ((uint64_t)rec_a.y << 32);
	const uint64_t gend = rec_gstart + ((uint64_t)rec_a.z | ((uint64_t)rec_a.w << 32));
	const TmplTail tt = load_tail(tp);
	MateHot mh[NM];
#pragma unroll
	for (int m = 0; m < NM; m++) mh[m] = load_mate(&tp->m[m]);
	const uint64_t ro[2] = {pr.refoff + tt.delta[0], pr.refoff + tt.delta[1]};
	const bool s0_random = CAT == CAT_B && !tt.aligned_is_mate1;
	uint32_t nl = 0;
	bool my_general = false;
	if (act) {
		// which mates need the general path: indel segments, reaching past the record end, or any N in the
		// 64-base blocks of the N-mask summary the mate touches (bases outside the mate only ever land on
		// bytes that are masked away)
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
			my_general |= general[m];
		}
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
			const uint32_t ns = random ? 0u : (uint32_t)(c_last - c_first + 1);
			const bool fast = !random && !general[mate];
			// the read's row of packed genome words starts at a 16-byte unit of bases[]
			const uint64_t wmin = md.fw ? (gadj >> 4) : (gadj >> 4) - (uint64_t)(ns ? ns - 1u : 0u);
			const uint32_t gunit = (uint32_t)(wmin >> 2);
			s_gunit[s * RW + lane] = fast ? gunit : 0xffffffffu;
			s_gb0[s * RW + lane] = (uint32_t)(gpos - 64ull * gunit);
			const uint32_t meta = (md.fw ? RF_FW : 0u) | (random ? RF_RANDOM : 0u) | (general[mate] ? RF_GENERAL : 0u) |
			                      ((uint32_t)(gadj & 15u) << 8) | ((A0 & 15u) << 12);
			const uint32_t w0 = (uint32_t)((gadj >> 4) - 4ull * gunit);  // word of slot 0 in the read's row
			DHot h;
			h.grow = smem_u32(s_gbuf + ((size_t)s * RW + lane) * GROW) + 4u * w0;
			h.sdst = 16u * (uint32_t)c_first;
			h.ns = ns | (meta << 16);
			h.area = smem_u32(s_area + ((size_t)s * RW + lane) * AREA_STRIDE);
			s_hot[s * RW + lane] = h;
			DHotQ hq;
			hq.qsrc = A0 >> 4;
			hq.qdst = 16u * (uint32_t)q_first;
			hq.nq = (uint32_t)(q_last - q_first + 1) | (meta << 16);
			hq.area = h.area;
			s_hotq[s * RW + lane] = hq;
			s_pos[s * RW + lane] = DPos{P, L};
			QD_CHK(ns <= (uint32_t)K && (uint32_t)(q_last - q_first + 1) <= (uint32_t)K && (!fast || (md.fw ? w0 + ns : w0 + 1u) < (uint32_t)(4 * PMAX)), 9, ns, K);  // last row word a chunk lane touches
			if (general[mate]) {
				DGen g;
				g.gpos = gpos;
				g.gend = gend;
				g.segoff = md.seg_off;
				g.nseg = md.nseg;
				g.pad[0] = g.pad[1] = 0;
				s_gen[s * RW + lane] = g;
			}
			const uint32_t s0c = lane == 0 ? 0u : (rstart[s] - 2u) >> 4;
			DShip sh;
			sh.s0 = s0c;
			sh.t0 = (uint32_t)c_last;
			sh.n = ((uint32_t)c_first - s0c + 1u) | (((uint32_t)(q_first - c_last) + 1u) << 8);
			sh.rstart = rstart[s];
			s_ship[s * (RW + 1) + lane] = sh;
			QD_CHK(16u * (sh.n & 0xffu) <= s_bytes && (sh.n >> 8) <= 2u && c_last > c_first && q_last > q_first, 5, sh.n, s_bytes);
			if (lane == nr - 1) {  // the slot after the tile's last record: its last quality chunk and '\n'
				DShip z;
				z.s0 = (uint32_t)q_last;
				z.t0 = 0;
				z.n = (((tend[s] - 1u) >> 4) - (uint32_t)q_last + 1u) | (1u << 16);
				z.rstart = tend[s];
				s_ship[s * (RW + 1) + nr] = z;
			}
		}
	}
	__syncwarp();
	const bool any_general = __any_sync(0xffffffffu, my_general);

	// ---------------- step A2 (lane = 16-byte piece): the reads' packed genome rows, asynchronously ----------------
	{
		const uint32_t nunits = (uint32_t)(a.genome_words >> 2);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			const int nitems = nr * PMAX;
			for (int q = lane; q < nitems; q += 32) {
				const int i = (int)__umulhi((uint32_t)q, pmagic), p = q - i * PMAX;
				const uint32_t unit = s_gunit[s * RW + i];
				if (unit != 0xffffffffu && unit + (uint32_t)p < nunits) {
					const uint32_t d = (uint32_t)__cvta_generic_to_shared(s_gbuf + ((size_t)s * RW + i) * GROW + 4 * p);
					asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"((const uint4 *)a.g.bases + unit + p) : "memory");
				}
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	}

	// ---------------- step A3 (lane = read): draw ops, names and separators into the slot images ----------------
	if (act) {
#pragma unroll
		for (int m = 0; m < NM; m++) {
			// the first draw ops, four per load (templates.cpp aligns the lists)
			if (mh[m].nops) {
				const uint4 *op4 = (const uint4 *)(a.c.ops + mh[m].ops_off);
				const uint4 o0 = __ldg(op4);
				uint32_t *so = s_ops + (size_t)m * DT_OPS * RW + lane;
				so[0] = o0.x;
				so[RW] = o0.y;
				so[2 * RW] = o0.z;
				so[3 * RW] = o0.w;
				if (mh[m].nops > 4 && DT_OPS > 4) {
					const uint4 o1 = __ldg(op4 + 1);
					so[4 * RW] = o1.x;
					if (DT_OPS > 5) so[5 * RW] = o1.y;
					if (DT_OPS > 6) so[6 * RW] = o1.z;
					if (DT_OPS > 7) so[7 * RW] = o1.w;
				}
			}
		}
		const uint32_t rec_name_off = rec_b.x, rl = rec_b.y;
#pragma unroll
		for (int s = 0; s < NS; s++) {
			const bool random = CAT == CAT_B && ((s == 0) == s0_random);
			const uint32_t L = random ? tt.opp_len : mh[CAT == CAT_B ? 0 : s].len;
			const uint32_t P = rstart[s] + nl, Pq = P + L + 3u;
			const uint32_t c_last = (P + L - 1u) >> 4, q_last = (Pq + L - 1u) >> 4;
			const uint32_t s0c = lane == 0 ? 0u : (rstart[s] - 2u) >> 4;
			unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + lane) * slot_bytes;
			unsigned char *sS = slot - 16u * s0c;                  // S space: + tile-relative byte position
			unsigned char *sT = slot + s_bytes - 16u * c_last;     // T space
			unsigned char *sN = slot + slot_bytes - 16u * q_last;  // next slot's S space
			// name (simplesim.cpp:104-110, :140-150): composed once, every byte stored to both files of a pair
			unsigned char *np = sS + rstart[s];
			if (s == 0) {
				constexpr bool TWO = NS == 2;
				ptrdiff_t d2 = 0;
				if (TWO) {
					const uint32_t s0c1 = lane == 0 ? 0u : (rstart[NS - 1] - 2u) >> 4;
					unsigned char *np1 = s_edge + ((size_t)(NS - 1) * (RW + 1) + lane) * slot_bytes - 16u * s0c1 + rstart[NS - 1];
					d2 = np1 - np;
				}
				auto put = [&](char *q, char ch) {
					*q = ch;
					if (TWO) q[d2] = ch;
				};
				const char *refid = a.g.names + rec_name_off;
				char *p = (char *)np;
				put(p, '@');
				put(p + 1, 'q');
				put(p + 2, 't');
				put(p + 3, 'i');
				put(p + 4, 'p');
				put(p + 5, '!');
				put(p + 6, ':');
				p += 7;
				if (CAT == CAT_U) {
					const uint32_t nd = nl - 13u - rl - mh[0].score_len - tt.typ_len;
					p = s_origin2<TWO>(p, d2, refid, rec_c, rl, ro[0], nd, mh[0]);
				} else if (CAT == CAT_B) {
					const uint32_t nd = (nl - 18u - 2u * rl - 2u * mh[0].score_len - tt.typ_len) >> 1;
					p = s_origin2<TWO>(p, d2, refid, rec_c, rl, ro[0], nd, mh[0]);
					put(p++, ':');
					p = s_origin2<TWO>(p, d2, refid, rec_c, rl, ro[0], nd, mh[0]);
				} else {
					const uint32_t nd0 = ndig64(ro[0]);
					const uint32_t nd1 = nl - 18u - 2u * rl - mh[0].score_len - mh[NM - 1].score_len - tt.typ_len - nd0;
					p = s_origin2<TWO>(p, d2, refid, rec_c, rl, ro[0], nd0, mh[0]);
					put(p++, ':');
					p = s_origin2<TWO>(p, d2, refid, rec_c, rl, ro[1], nd1, mh[NM - 1]);
				}
				put(p, ':');
				put(p + 1, (char)(tt.typ & 0xffu));
				if (tt.typ_len > 1) put(p + 2, (char)(tt.typ >> 8));
				put(p + 1 + tt.typ_len, '\n');
				QD_CHK(p + 2 + tt.typ_len == (char *)np + nl, 6, (uint32_t)(p - (char *)np), nl);
			}
			// separators
			sT[P + L] = '\n';
			sT[P + L + 1u] = '+';
			sT[P + L + 2u] = '\n';
			sN[Pq + L] = '\n';
		}
	}
	if (skip & 1) return;  // attribution runs only (QTIPSIM_SKIP): output incomplete

	// ---------------- step Bq (lane = quality chunk): item q = read * K + slot, read-major ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		char *tbs = tb[s];
		const uint32_t hot_a = smem_u32(s_hotq + s * RW);
		const int nitems = nr * K;
		for (int q0 = lane; q0 < nitems && !(skip & 8); q0 += 32 * DT_UNROLL) {
			QualLoad ql[DT_UNROLL];
			uint32_t meta[DT_UNROLL], dst[DT_UNROLL], ar[DT_UNROLL];  // ar: area address for a string's first / last chunk, else 0
			bool live[DT_UNROLL];
#pragma unroll
			for (int u = 0; u < DT_UNROLL; u++) {
				const int q = q0 + 32 * u;
				live[u] = false;
				if (q < nitems) {
					const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
					const uint4 h4 = lds128(hot_a + 16u * (uint32_t)i);  // qsrc | qdst | nq, meta | area
					const int nq = (int)(h4.z & 0xffffu);
					meta[u] = h4.z >> 16;
					live[u] = j < nq;
					dst[u] = h4.y + 16u * (uint32_t)j;
					ar[u] = j == 0 ? h4.w + 16u * AREA_QUAL_HEAD : (j == nq - 1 ? h4.w + 16u * AREA_QUAL_TAIL : 0u);
					if (live[u] && !(CAT == CAT_B && (meta[u] & RF_RANDOM))) ql[u] = qual_issue(a.c.qual_out, h4.x + (uint32_t)j);
				}
			}
#pragma unroll
			for (int u = 0; u < DT_UNROLL; u++) {
				if (!live[u]) continue;
				const uint4 v = (CAT == CAT_B && (meta[u] & RF_RANDOM)) ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)  // 'I' (simplesim.h:105)
				                                      : qual_finish(ql[u], (meta[u] >> 12) & 15u);
				if (ar[u]) sts128(ar[u], v);
				else {
					QD_CHK(dst[u] + 16u <= tend[s], 4, dst[u], tend[s]);
					ST16(tbs + dst[u], v);
				}
			}
		}
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncwarp();

	// ---------------- step Bs (lane = sequence chunk): genome words from the rows in shared memory ----------------
	// Two loops, so that the hot one holds no call: fast-path reads first, then -- only in tiles that have any --
	// the reads of the general path (N, past the record end, indel segments).
#pragma unroll
	for (int s = 0; s < NS; s++) {
		char *tbs = tb[s];
		const uint32_t hot_a = smem_u32(s_hot + s * RW);
		const int nitems = nr * K;
		for (int q = lane; q < nitems && !(skip & 16); q += 32) {
			const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
			const uint4 h4 = lds128(hot_a + 16u * (uint32_t)i);  // grow | sdst | ns, meta | area
			const int ns = (int)(h4.z & 0xffffu);
			const uint32_t meta = h4.z >> 16;
			if (j >= ns || (meta & RF_GENERAL)) continue;
			const uint32_t wa = (meta & RF_FW) ? h4.x + 4u * (uint32_t)j : h4.x - 4u * (uint32_t)j;
			SeqLoad sl;
			sl.lo = lds32(wa);
			sl.hi = lds32(wa + 4u);
			const uint4 v = seq_finish(sl, meta);
			if (j == 0 || j == ns - 1) sts128(h4.w + (j ? 16u * AREA_SEQ_TAIL : 16u * AREA_SEQ_HEAD), v);
			else {
				QD_CHK(h4.y + 16u * (uint32_t)j + 16u <= tend[s], 3, h4.y, tend[s]);
				ST16(tbs + h4.y + 16u * (uint32_t)j, v);
			}
		}
		if (any_general) {
			for (int q = lane; q < nitems && !(skip & 16); q += 32) {
				const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
				const uint4 h4 = lds128(hot_a + 16u * (uint32_t)i);
				const int ns = (int)(h4.z & 0xffffu);
				const uint32_t meta = h4.z >> 16;
				if (j >= ns || !(meta & RF_GENERAL)) continue;
				const uint32_t dst = h4.y + 16u * (uint32_t)j;
				const DPos pp = s_pos[s * RW + i];
				const DGen &gg = s_gen[s * RW + i];
				const uint4 v = seq16_general(gv, s_lowmask, a.c.segs + gg.segoff, (int)gg.nseg, (meta & RF_FW) != 0, (int)pp.L, gg.gpos,
				                              gg.gend, (int)dst - (int)pp.P);
				if (j == 0 || j == ns - 1) sts128(h4.w + (j ? 16u * AREA_SEQ_TAIL : 16u * AREA_SEQ_HEAD), v);
				else ST16(tbs + dst, v);
			}
		}
	}
	__syncwarp();
	if (skip & 2) return;

	// ---------------- step C (lane = read): edge chunks merged into the slot images, then the draws ----------------
	// Unpaired reads with the native generator and tiles of at most 16 reads: TWO lanes per read.  Lanes 16-31 take
	// over half of the merges (first sequence chunk, last quality chunk: never the same 16 bytes as the other
	// half's last sequence / first quality chunk) and every second draw op -- Philox draws are addressed by their
	// index, so the ops of a read need not be walked in order.
	const bool dual = CAT == CAT_U && MODE == QTIPSIM_RNG_PHILOX && RW <= 16;
	if (dual) {
		const int half = lane >> 4, rl_ = lane & 15;
		const uint32_t c_nops = __shfl_sync(0xffffffffu, mh[0].nops, rl_);
		const uint32_t c_opsoff = __shfl_sync(0xffffffffu, mh[0].ops_off, rl_);
		const uint32_t c_fw = __shfl_sync(0xffffffffu, mh[0].fw ? 1u : 0u, rl_);
		uint32_t P = 0, L = 0, c_first = 0, c_last = 0, s0c = 0;
		unsigned char *slot = s_edge + (size_t)rl_ * slot_bytes;
		if (rl_ < nr) {
			const DPos pp = s_pos[rl_];
			P = pp.P;
			L = pp.L;
			const uint32_t Pq = P + L + 3u;
			c_first = P >> 4;
			c_last = (P + L - 1u) >> 4;
			const uint32_t q_first = Pq >> 4, q_last = (Pq + L - 1u) >> 4;
			s0c = s_ship[rl_].s0;
			const uint4 *ar = s_area + (size_t)rl_ * AREA_STRIDE;
			uint4 *img;
			uint4 v;
			if (half) {
				img = (uint4 *)(slot + 16u * (c_first - s0c));  // first sequence chunk: bytes from P on
				v = *img;
				overlay(v, ar[AREA_SEQ_HEAD], s_lowmask, (int)P - 16 * (int)c_first, 16);
				*img = v;
				img = (uint4 *)(slot + slot_bytes);  // last quality chunk (next slot): bytes before Pq + L
				v = *img;
				overlay(v, ar[AREA_QUAL_TAIL], s_lowmask, 0, (int)(Pq + L) - 16 * (int)q_last);
				*img = v;
			} else {
				img = (uint4 *)(slot + s_bytes);  // last sequence chunk: bytes before P + L
				v = *img;
				overlay(v, ar[AREA_SEQ_TAIL], s_lowmask, 0, (int)(P + L) - 16 * (int)c_last);
				*img = v;
				img = (uint4 *)(slot + s_bytes + 16u * (q_first - c_last));  // first quality chunk: bytes from Pq on
				v = *img;
				overlay(v, ar[AREA_QUAL_HEAD], s_lowmask, (int)Pq - 16 * (int)q_first, 16);
				*img = v;
			}
		}
		__syncwarp();  // the draws patch the merged images
		if (rl_ < nr && c_nops > (uint32_t)half) {
			const uint64_t gi = a.read_index_base + r0 + (uint64_t)rl_;
			const bool fw = c_fw != 0;
			const bool gen = ((s_hot[rl_].ns >> 16) & RF_GENERAL) != 0;
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			const uint32_t gb0 = s_gb0[rl_];
			const uint32_t *row = s_gbuf + (size_t)rl_ * GROW;
			const uint32_t *ops = a.c.ops + c_opsoff;
			uint32_t curb = 0xffffffffu;
			Philox4 blk{};
			for (uint32_t j = (uint32_t)half; j < c_nops; j += 2) {
				const uint32_t e = j < (uint32_t)DT_OPS ? s_ops[j * RW + rl_] : __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				QD_CHK(outp < L, 7, outp, L);
				if ((j >> 3) != curb) {  // draw j of the read = 16 bits of Philox block j / 8 (same indexing as the serial walk)
					curb = j >> 3;
					blk = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), curb, PHILOX_DOM_BASE | (uint32_t)CAT, (uint32_t)a.philox_seed,
					                 (uint32_t)(a.philox_seed >> 32));
				}
				const uint32_t wi = (j >> 1) & 3u;
				const uint32_t wv = wi == 0 ? blk.v[0] : wi == 1 ? blk.v[1] : wi == 2 ? blk.v[2] : blk.v[3];
				const uint32_t b = (j & 1u) ? (wv >> 16) : (wv & 0xffffu);
				uint32_t code;
				if ((e & 3u) == OP_X) {
					uint32_t ref;
					if (gen) {
						const DGen &gg = s_gen[rl_];
						ref = ref_code_at(a.g, gg.gpos + (e >> 17), gg.gend);
					} else {
						const uint32_t t = gb0 + (e >> 17);
						ref = (row[t >> 4] >> (2u * (t & 15u))) & 3u;
					}
					code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
				} else code = b >> 14;
				const uint32_t ch = (T >> (8u * code)) & 0xffu;
				const uint32_t x = P + outp, cx = x >> 4;
				if (cx == c_first) slot[x - 16u * s0c] = (unsigned char)ch;
				else if (cx == c_last) slot[s_bytes + x - 16u * c_last] = (unsigned char)ch;
				else tb[0][x] = (char)ch;
			}
		}
	} else if ((CAT == CAT_C || CAT == CAT_D) && MODE == QTIPSIM_RNG_PHILOX && RW <= 16) {
		// Pairs with the native generator: lanes 0-15 serve the first file and mate, lanes 16-31 the second (the
		// second mate's draws start at index nops of the first: addressable)
		const int sx = NS == 2 ? (lane >> 4) : 0, rl_ = lane & 15;
		const uint32_t n0 = __shfl_sync(0xffffffffu, mh[0].nops, rl_), n1 = __shfl_sync(0xffffffffu, mh[NM - 1].nops, rl_);
		const uint32_t o0 = __shfl_sync(0xffffffffu, mh[0].ops_off, rl_), o1 = __shfl_sync(0xffffffffu, mh[NM - 1].ops_off, rl_);
		const uint32_t f0 = __shfl_sync(0xffffffffu, mh[0].fw ? 1u : 0u, rl_), f1 = __shfl_sync(0xffffffffu, mh[NM - 1].fw ? 1u : 0u, rl_);
		const uint32_t c_nops = sx ? n1 : n0, c_opsoff = sx ? o1 : o0, c_fw = sx ? f1 : f0, d_first = sx ? n0 : 0u;
		uint32_t P = 0, L = 0, c_first = 0, c_last = 0, s0c = 0;
		unsigned char *slot = s_edge + ((size_t)sx * (RW + 1) + rl_) * slot_bytes;
		if (rl_ < nr) {
			const DPos pp = s_pos[sx * RW + rl_];
			P = pp.P;
			L = pp.L;
			const uint32_t Pq = P + L + 3u;
			c_first = P >> 4;
			c_last = (P + L - 1u) >> 4;
			const uint32_t q_first = Pq >> 4, q_last = (Pq + L - 1u) >> 4;
			s0c = s_ship[sx * (RW + 1) + rl_].s0;
			const uint4 *ar = s_area + ((size_t)sx * RW + rl_) * AREA_STRIDE;
			uint4 *img;
			uint4 v;
			img = (uint4 *)(slot + 16u * (c_first - s0c));
			v = *img;
			overlay(v, ar[AREA_SEQ_HEAD], s_lowmask, (int)P - 16 * (int)c_first, 16);
			*img = v;
			img = (uint4 *)(slot + s_bytes);
			v = *img;
			overlay(v, ar[AREA_SEQ_TAIL], s_lowmask, 0, (int)(P + L) - 16 * (int)c_last);
			*img = v;
			img = (uint4 *)(slot + s_bytes + 16u * (q_first - c_last));
			v = *img;
			overlay(v, ar[AREA_QUAL_HEAD], s_lowmask, (int)Pq - 16 * (int)q_first, 16);
			*img = v;
			img = (uint4 *)(slot + slot_bytes);
			v = *img;
			overlay(v, ar[AREA_QUAL_TAIL], s_lowmask, 0, (int)(Pq + L) - 16 * (int)q_last);
			*img = v;
			if (c_nops) {  // (same lane merged this stream's images: program order suffices)
				const uint64_t gi = a.read_index_base + r0 + (uint64_t)rl_;
				const bool fw = c_fw != 0;
				const bool gen = ((s_hot[sx * RW + rl_].ns >> 16) & RF_GENERAL) != 0;
				const uint32_t T = fw ? TBL_FW : TBL_RC;
				const uint32_t gb0 = s_gb0[sx * RW + rl_];
				const uint32_t *row = s_gbuf + ((size_t)sx * RW + rl_) * GROW;
				const uint32_t *ops = a.c.ops + c_opsoff;
				char *tbx = sx ? tb[NS - 1] : tb[0];
				uint32_t curb = 0xffffffffu;
				Philox4 blk{};
				for (uint32_t j = 0; j < c_nops; j++) {
					const uint32_t e = j < (uint32_t)DT_OPS ? s_ops[(sx * DT_OPS + j) * RW + rl_] : __ldg(&ops[j]);
					const uint32_t outp = (e >> 2) & 0x7fffu;
					QD_CHK(outp < L, 7, outp, L);
					const uint32_t d = d_first + j;  // draw index within the pair
					if ((d >> 3) != curb) {
						curb = d >> 3;
						blk = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), curb, PHILOX_DOM_BASE | (uint32_t)CAT, (uint32_t)a.philox_seed,
						                 (uint32_t)(a.philox_seed >> 32));
					}
					const uint32_t wi = (d >> 1) & 3u;
					const uint32_t wv = wi == 0 ? blk.v[0] : wi == 1 ? blk.v[1] : wi == 2 ? blk.v[2] : blk.v[3];
					const uint32_t b = (d & 1u) ? (wv >> 16) : (wv & 0xffffu);
					uint32_t code;
					if ((e & 3u) == OP_X) {
						uint32_t ref;
						if (gen) {
							const DGen &gg = s_gen[sx * RW + rl_];
							ref = ref_code_at(a.g, gg.gpos + (e >> 17), gg.gend);
						} else {
							const uint32_t t = gb0 + (e >> 17);
							ref = (row[t >> 4] >> (2u * (t & 15u))) & 3u;
						}
						code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
					} else code = b >> 14;
					const uint32_t ch = (T >> (8u * code)) & 0xffu;
					const uint32_t x = P + outp, cx = x >> 4;
					if (cx == c_first) slot[x - 16u * s0c] = (unsigned char)ch;
					else if (cx == c_last) slot[s_bytes + x - 16u * c_last] = (unsigned char)ch;
					else tbx[x] = (char)ch;
				}
			}
		}
	} else if (act) {
		// the four areas hold whole chunks; only the bytes inside the string go into the image (names and
		// separators are already there).  The last quality chunk belongs to the next slot's first chunk.
#pragma unroll
		for (int s = 0; s < NS; s++) {
			const DPos pp = s_pos[s * RW + lane];
			const uint32_t P = pp.P, L = pp.L, Pq = P + L + 3u;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4, q_first = Pq >> 4, q_last = (Pq + L - 1u) >> 4;
			const uint32_t s0c = s_ship[s * (RW + 1) + lane].s0;
			unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + lane) * slot_bytes;
			const uint4 *ar = s_area + ((size_t)s * RW + lane) * AREA_STRIDE;
			uint4 *img;
			uint4 v;
			if (!(CAT == CAT_B && ((s == 0) == s0_random))) {  // (a random mate's sequence is drawn in step Br)
				img = (uint4 *)(slot + 16u * (c_first - s0c));  // first sequence chunk: bytes from P on
				v = *img;
				overlay(v, ar[AREA_SEQ_HEAD], s_lowmask, (int)P - 16 * (int)c_first, 16);
				*img = v;
				img = (uint4 *)(slot + s_bytes);  // last sequence chunk: bytes before P + L
				v = *img;
				overlay(v, ar[AREA_SEQ_TAIL], s_lowmask, 0, (int)(P + L) - 16 * (int)c_last);
				*img = v;
			}
			img = (uint4 *)(slot + s_bytes + 16u * (q_first - c_last));  // first quality chunk: bytes from Pq on
			v = *img;
			overlay(v, ar[AREA_QUAL_HEAD], s_lowmask, (int)Pq - 16 * (int)q_first, 16);
			*img = v;
			img = (uint4 *)(slot + slot_bytes);  // last quality chunk: bytes before Pq + L
			v = *img;
			overlay(v, ar[AREA_QUAL_TAIL], s_lowmask, 0, (int)(Pq + L) - 16 * (int)q_last);
			*img = v;
		}
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
		// slot T space, interior -> HBM (already stored by step Bs: same warp, ordered by __syncwarp)
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
			const uint32_t n = mh[m].nops;
			if (!n) continue;
			const DPos pp = s_pos[s * RW + lane];
			const uint32_t P = pp.P, L = pp.L;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4, s0c = s_ship[s * (RW + 1) + lane].s0;
			const bool fw = mh[m].fw;
			const bool gen = ((s_hot[s * RW + lane].ns >> 16) & RF_GENERAL) != 0;
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			const uint32_t gb0 = s_gb0[s * RW + lane];
			const uint32_t *row = s_gbuf + ((size_t)s * RW + lane) * GROW;
			const uint32_t *ops = a.c.ops + mh[m].ops_off;
			for (uint32_t j = 0; j < n; j++) {
				const uint32_t e = j < (uint32_t)DT_OPS ? s_ops[(m * DT_OPS + j) * RW + lane] : __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				QD_CHK(outp < L, 7, outp, L);
				uint32_t code;
				if ((e & 3u) == OP_X) {
					uint32_t ref;
					if (gen) {
						const DGen &gg = s_gen[s * RW + lane];
						ref = ref_code_at(a.g, gg.gpos + (e >> 17), gg.gend);
					} else {
						const uint32_t t = gb0 + (e >> 17);
						ref = (row[t >> 4] >> (2u * (t & 15u))) & 3u;
					}
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
			// the random mate (init_random, simplesim.h:89-117) follows in the draw stream: hand the generator to step Br
			DRnd rd;
			rd.r0 = MODE == QTIPSIM_RNG_REPLAY ? st.s1 : dk;
			rd.r1 = MODE == QTIPSIM_RNG_REPLAY ? st.s2 : 0u;
			rd.flags = (s0_random ? 0u : 1u) | (mh[0].fw ? 2u : 0u);
			rd.pad = 0;
			s_rnd[lane] = rd;
		}
	}
	if (CAT == CAT_B) {
		__syncwarp();
		// ---------------- step Br (lane = 16-byte chunk of a random mate): opp_len draws per read, 16 per lane ----------------
		// Both generators are addressable: draw j of the mate is Philox draw (first + j), or the L'Ecuyer state
		// (j + 1) steps past the handed-over one (jump tables), walked forwards for '+' mates and backwards -- a
		// multiplication by a^-1 -- for '-' mates, whose bases are printed in reverse (simplesim.h:108-113).
		const int nitems = nr * K;
		for (int q = lane; q < nitems; q += 32) {
			const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
			const DRnd rd = s_rnd[i];
			const int sr = (int)(rd.flags & 1u);
			const bool fw = (rd.flags & 2u) != 0;
			const DPos pp = s_pos[sr * RW + i];
			const int P = (int)pp.P, L = (int)pp.L;
			const int c_first = P >> 4, ns = ((P + L - 1) >> 4) - c_first + 1;
			if (j >= ns) continue;
			const int k0 = 16 * (c_first + j) - P;       // output position of the chunk's first byte
			const int t_first = k0 < 0 ? -k0 : 0;        // first byte of the chunk inside the mate
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			uint32_t w[4] = {0u, 0u, 0u, 0u};
			if (MODE == QTIPSIM_RNG_REPLAY) {
				const int jd0 = fw ? k0 + t_first : L - 1 - (k0 + t_first);  // draw of the first byte inside
				LState st = ljump_tbl(LState{rd.r0, rd.r1}, (uint32_t)(fw ? jd0 : jd0 + 1), a.jt1, a.jt2);
#pragma unroll
				for (int t = 0; t < 16; t++) {
					if (t >= t_first && k0 + t < L) {
						if (fw) st = lstep(st);
						const uint32_t code = lbase_code(lfloat(lout(st)));
						w[t >> 2] |= ((T >> (8u * code)) & 0xffu) << (8 * (t & 3));
						if (!fw) st = lstep_back(st);
					}
				}
			} else {
				const uint64_t gi = a.read_index_base + r0 + (uint64_t)i;
				uint32_t curb = 0xffffffffu;
				Philox4 blk{};
#pragma unroll
				for (int t = 0; t < 16; t++) {
					if (t >= t_first && k0 + t < L) {
						const uint32_t d = (uint32_t)(fw ? k0 + t : L - 1 - (k0 + t));  // two bits per base, 64 bases per block (philox.h)
						if ((d >> 6) != curb) {
							curb = d >> 6;
							blk = philox_rmate_block(gi, curb, CAT, a.philox_seed);
						}
						const uint32_t wi = (d >> 4) & 3u;
						const uint32_t x = wi == 0 ? blk.v[0] : wi == 1 ? blk.v[1] : wi == 2 ? blk.v[2] : blk.v[3];
						const uint32_t code = (x >> (2u * (d & 15u))) & 3u;
						w[t >> 2] |= ((T >> (8u * code)) & 0xffu) << (8 * (t & 3));
					}
				}
			}
			const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
			if (j == 0 || j == ns - 1) s_area[((size_t)sr * RW + i) * AREA_STRIDE + (j ? AREA_SEQ_TAIL : AREA_SEQ_HEAD)] = v;
			else ST16((sr ? tb[NS - 1] : tb[0]) + 16u * (size_t)(c_first + j), v);
		}
		__syncwarp();
		// ---------------- step C2 (lane = read): the random mate's first and last chunk into the slot image ----------------
		if (act) {
			const int s = s0_random ? 0 : 1;
			const DPos pp = s_pos[s * RW + lane];
			const uint32_t P = pp.P, L = pp.L;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4, s0c = s_ship[s * (RW + 1) + lane].s0;
			unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + lane) * slot_bytes;
			const uint4 *ar = s_area + ((size_t)s * RW + lane) * AREA_STRIDE;
			uint4 *img = (uint4 *)(slot + 16u * (c_first - s0c));
			uint4 v = *img;
			overlay(v, ar[AREA_SEQ_HEAD], s_lowmask, (int)P - 16 * (int)c_first, 16);
			*img = v;
			img = (uint4 *)(slot + s_bytes);
			v = *img;
			overlay(v, ar[AREA_SEQ_TAIL], s_lowmask, 0, (int)(P + L) - 16 * (int)c_last);
			*img = v;
		}
	}
	__syncwarp();
	if (skip & 4) return;

	// ---------------- step D (lane = edge chunk): slot images -> HBM ----------------
	{
		const int H = 1 << lgH, spi = 32 >> lgH;  // slots per iteration
		const int sub = lane >> lgH, jl = lane & (H - 1);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			char *tbs = tb[s];
			const uint32_t end = tend[s], phs = ph[s];
			const uint32_t last_chunk = end >> 4;  // shared with the next tile when end & 15 != 0 (else nothing of ours)
			// hrounds < 0: the lane groups cover the S chunks only; the (one or two) T chunks of a slot go in a second
			// loop with two lanes per slot (launcher: when S + T would not fit the same power-of-two group)
			const bool split = hrounds < 0;
			const int hr = split ? -hrounds : hrounds;
			for (int rd = 0; rd < hr; rd++) {
				const uint32_t j = (uint32_t)(rd * H + jl);
				for (int i0 = 0; i0 <= nr; i0 += spi) {
					const int i = i0 + sub;
					if (i > nr) continue;
					const DShip sh = s_ship[s * (RW + 1) + i];
					const uint32_t nS = sh.n & 0xffu, nT = split ? 0u : (sh.n >> 8) & 0xffu;
					if (j >= nS + nT) continue;
					const unsigned char *slot = s_edge + ((size_t)s * (RW + 1) + i) * slot_bytes;
					const uint32_t chunk = j < nS ? sh.s0 + j : sh.t0 + (j - nS);
					const unsigned char *src = j < nS ? slot + 16u * j : slot + s_bytes + 16u * (j - nS);
					if ((chunk == 0 && phs) || chunk == last_chunk) continue;  // the tile's ragged ends: below
					ST16(tbs + 16u * (size_t)chunk, *(const uint4 *)src);
				}
			}
			if (split) {
				for (int i0 = 0; i0 < nr; i0 += 16) {
					const int i = i0 + (lane >> 1);
					if (i >= nr) continue;
					const DShip sh = s_ship[s * (RW + 1) + i];
					const uint32_t jt = (uint32_t)lane & 1u;
					if (jt >= ((sh.n >> 8) & 0xffu)) continue;
					const uint32_t chunk = sh.t0 + jt;
					if (chunk == last_chunk) continue;
					ST16(tbs + 16u * (size_t)chunk, *(const uint4 *)(s_edge + ((size_t)s * (RW + 1) + i) * slot_bytes + s_bytes + 16u * jt));
				}
			}
			// the tile's first and last chunk are shared with the neighbouring tiles: one byte per lane
			{
				const DShip z = s_ship[s * (RW + 1) + nr];
				const unsigned char *head = s_edge + (size_t)s * (RW + 1) * slot_bytes;                                   // slot 0, chunk 0
				const unsigned char *tail = s_edge + ((size_t)s * (RW + 1) + nr) * slot_bytes + 16u * (last_chunk - z.s0);  // in the last slot
				const uint32_t b = (uint32_t)lane & 15u;
				if (lane < 16) {
					if (phs && b >= phs) tbs[b] = (char)head[b];
				} else if (b < (end & 15u)) tbs[16u * (size_t)last_chunk + b] = (char)tail[b];
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
	// chunks a string of max_body bytes can touch at any alignment
	const uint32_t K = (max_body + 14u) / 16u + 1u;
	const uint32_t kmagic = (uint32_t)((0x100000000ull + K - 1) / K);  // q / K == umulhi(q, kmagic) for q < 2^16
	// 16-byte pieces of packed genome per read: K + 1 words plus up to 3 of alignment
	const uint32_t PMAX = ((K + 1u + 3u + 3u) / 4u) | 1u;  // odd: see AREA_STRIDE
	const uint32_t pmagic = (uint32_t)((0x100000000ull + PMAX - 1) / PMAX);
	// S part of a slot: from the chunk of (record start - 2) to the chunk of the first sequence byte
	uint32_t s_bytes = ((max_name + 2u + 15u + 15u) & ~15u) + 16u;
	if ((((s_bytes + T_BYTES) >> 4) & 1u) == 0) s_bytes += 16u;  // slot stride: an odd multiple of 16 bytes (see AREA_STRIDE)
	auto warp_bytes = [&](int rw) -> uint32_t {
		return NS * (uint32_t)(rw + 1) * (s_bytes + T_BYTES + (uint32_t)sizeof(DShip)) +
		       (uint32_t)rw * (NS * (uint32_t)(sizeof(DHot) + sizeof(DHotQ) + sizeof(DGen) + sizeof(DPos) + 16u * AREA_STRIDE + 16u * PMAX + 8u) + NM * DT_OPS * 4u + (CAT == CAT_B ? (uint32_t)sizeof(DRnd) : 0u));
	};
	const uint32_t sm_budget = 216u * 1024u;
	int rw = env_rw > 0 ? env_rw : 32;
	if (rw > 32) rw = 32;
	if (env_rw <= 0)
		// resident warps per SM wanted: 24; 12 for concordant / discordant pairs (their lane = read steps are long
		// and want full half-warps more than they want occupancy: measured)
		while (rw > 4 && warp_bytes(rw) * ((CAT == CAT_C || CAT == CAT_D) ? 12u : 24u) > sm_budget) rw >>= 1;
	if (warp_bytes(rw) + 1024 > sm_budget || (uint64_t)rw * K >= 65536u) return false;
	int warps = env_warps > 0 ? env_warps : (NS == 1 ? 1 : DT_MAX_WARPS);  // tiles per CTA: measured (QTIPSIM_WARPS overrides)
	if (warps > DT_MAX_WARPS) warps = DT_MAX_WARPS;
	while (warps > 1 && warp_bytes(rw) * (uint32_t)warps + 1024 > sm_budget) warps >>= 1;
	const uint32_t wb = warp_bytes(rw);
	const size_t smem = (size_t)wb * warps;
	// step D: chunks per staging slot
	// step D: lane groups of 2^lgH per slot; if the S chunks alone fill a power of two, the T chunks go separately
	const int hS = (int)(s_bytes / 16u);
	int lgH = 0;
	while ((1 << lgH) < hS && lgH < 5) lgH++;
	const bool split = hS + 2 > (1 << lgH);
	const int hslots = split ? hS : hS + 2;
	int hrounds = (hslots + (1 << lgH) - 1) >> lgH;
	if (split) hrounds = -hrounds;
	auto kern = k_simulate_direct<MODE, CAT>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + rw - 1) / rw;
	const uint64_t nblocks = (ntiles + warps - 1) / warps;
	static int skip = -1;  // QTIPSIM_SKIP (attribution runs only): bit 0 stop after step A, 1 after B, 2 after C
	if (skip < 0) {
		const char *e = getenv("QTIPSIM_SKIP");
		skip = e ? atoi(e) : 0;
	}
	kern<<<(unsigned)nblocks, warps * 32, smem, st>>>(a, rw, (int)K, kmagic, (int)PMAX, pmagic, lgH, hrounds, s_bytes, wb, skip);
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
