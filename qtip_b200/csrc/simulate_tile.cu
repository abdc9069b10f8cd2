// simulate_tile.cu -- the production simulate+emit kernel (sm_100a), second generation of
// simulate_direct.cu: the same warp-autonomous tiles and register-to-HBM interior chunks, with the
// per-record overhead restructured after the L1 data pipe turned out to be the kernel's limiter
// (profiles/r01_k_simulate_direct_ncu_summary.md: 83 % of the wavefront peak, 51 wavefronts per read).
//
// Replaces SimulatedRead::init/mutate/init_random (simplesim.h:68-117, simplesim.cpp:31-71)
// and SimulatedRead::write/write_pair (simplesim.cpp:102-171).
//
// What changed against simulate_direct.cu:
//   * one edge IMAGE per record instead of an image plus four areas: the chunk lanes of steps Bq / Bs store a
//     string's first and last 16-byte chunk WHOLE into the image entry of that chunk; the read name and the
//     separators are written over them afterwards (step N), so nothing is merged under a byte mask at lane = read.
//     The one chunk two strings can share (sequence tail, "\n+\n", quality head) keeps two entries, merged in step N
//   * names are assembled from 16-byte PIECES held in registers ("\n@qtip!:", refid, ":<strand>:<offset>",
//     ":<score>", ":<typ>\n") and appended with aligned 32-bit stores through a running accumulator -- about ten
//     word stores per name instead of one byte store per character
//   * step D is one flat loop over all image entries of the tile
//   * only the 16-byte pieces of packed genome a read's chunks touch are staged
//
//   step A   lane = read: plan record + template (16-byte descriptor blocks) -> chunk geometry in shared memory
//   step A2  lane = 16-byte piece: each read's packed genome span -> its row in shared memory, cp.async
//   step Bq  lane = quality chunk (item = read * K + slot, read-major)
//   step Bs  lane = sequence chunk; general-path reads (N, past the record end, indel segments) in a second loop
//   step C   lane = read (two lanes per read or pair in the native mode): the read's draws patched over image or
//            HBM bytes (REPLAY: the reference's L'Ecuyer stream stepped from the plan's state; PHILOX: by index)
//   step A2r bad-end pairs, native mode: the random mate's 2-bit codes (Philox blocks of 64 bases) into its row, to be
//            printed by step Bs like a genome span
//   step Br  bad-end pairs, REPLAY: lane = chunk of the random mate: init_random's draws, 16 per lane
//   step N   one lane per record and file: '\n' of the previous record + name, separators, the shared chunk's merge
//   step D   lane = image entry: images -> HBM, 16 bytes per lane; the tile's first and last chunk, shared with
//            the neighbouring tiles, one byte per lane
//
// Categories with a body shorter than 17 bytes use the staged kernel in simulate_warp.cu; records too large for
// either use the generic kernel (kernels.cu).  simulate_direct.cu stays selectable (QTIPSIM_KERNEL=direct).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"
#include "simulate_common.cuh"

namespace qt {

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

#ifndef QT_TL_MAX_WARPS
#define QT_TL_MAX_WARPS 4
#endif
#ifndef QT_TL_MIN_BLOCKS
#define QT_TL_MIN_BLOCKS 8
#endif
#ifndef QT_TL_MIN_BLOCKS_PAIR
#define QT_TL_MIN_BLOCKS_PAIR 6
#endif
#ifndef QT_TL_UNROLL
#define QT_TL_UNROLL 2
#endif
#define TST16(ptr, v) __stcs((uint4 *)(ptr), (v))
static constexpr int TL_MAX_WARPS = QT_TL_MAX_WARPS;
static constexpr int TL_OPS = 6;                 // draw ops per mate kept in shared memory from step A
static constexpr int TL_UNROLL = QT_TL_UNROLL;   // step Bq: chunk iterations with their loads in flight together

// bytes [0, n) of a 16-byte chunk, n = 0..16 (general path only: seq16_general reads it from global memory)
__device__ const uint32_t g_tl_lowmask[17][4] = {
	{0u, 0u, 0u, 0u},
	{0xffu, 0u, 0u, 0u}, {0xffffu, 0u, 0u, 0u}, {0xffffffu, 0u, 0u, 0u}, {0xffffffffu, 0u, 0u, 0u},
	{0xffffffffu, 0xffu, 0u, 0u}, {0xffffffffu, 0xffffu, 0u, 0u}, {0xffffffffu, 0xffffffu, 0u, 0u}, {0xffffffffu, 0xffffffffu, 0u, 0u},
	{0xffffffffu, 0xffffffffu, 0xffu, 0u}, {0xffffffffu, 0xffffffffu, 0xffffu, 0u}, {0xffffffffu, 0xffffffffu, 0xffffffu, 0u},
	{0xffffffffu, 0xffffffffu, 0xffffffffu, 0u},
	{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffu}, {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffu},
	{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffu}, {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}};

// Per-read, per-stream chunk geometry (positions are relative to the tile base, which is the output offset of the
// tile's first byte rounded down to 16); slot j of a string = its j-th 16-byte chunk.
// A record's image consists of STR entries of 16 bytes: entries [0, hS) hold chunks [s0, s0 + nS) -- from the
// previous record's last quality chunk to this record's first sequence chunk --, entry hS the chunk of the
// sequence's last byte (T_seq) and entry hS + 1 the chunk of the first quality byte (T_qual); the two are the same
// chunk of the file when the three separator bytes do not cross a 16-byte boundary.  `img` fields are byte offsets
// from the stream's image base.
struct __align__(16) THot {   // what a sequence-chunk lane reads
	uint32_t grow;    // fast path: shared-memory address of slot 0's first packed genome word
	uint32_t sdst;    // byte position of sequence slot 0
	uint32_t ns;      // sequence slots | meta << 16; meta = RF_* | genome phase << 8 | quality phase << 12
	uint32_t img;     // image entry of the first sequence chunk | of T_seq << 16
};
struct __align__(16) THotQ {  // what a quality-chunk lane reads
	uint32_t qsrc;    // quality arena, 16-byte unit of slot 0
	uint32_t qdst;    // byte position of quality slot 0
	uint32_t nq;      // quality slots | meta << 16
	uint32_t img;     // image entry T_qual | first entry of the next record's image << 16
};
struct __align__(8) TRnd {
	uint32_t r0, r1;  // REPLAY: generator state before the mate's first draw; PHILOX: index of its first draw, 0
	uint32_t flags;   // bit 0: stream of the random mate, bit 1: printed on the '+' strand
	uint32_t pad;
};
struct __align__(8) TPos {
	uint32_t P;       // byte position of the first sequence byte
	uint32_t L;       // body length
};
struct __align__(16) TGen {  // general path only; kept in the read's (otherwise unused) genome row
	uint64_t gpos;    // genome position of the mate's leftmost reference base
	uint64_t gend;    // end of the record
	uint32_t segoff;  // into DevCategory::segs
	uint32_t nseg;
	uint32_t pad[2];
};
struct __align__(8) TShip {
	uint32_t a;       // s0 | chunk of the sequence's last byte << 16
	uint32_t b;       // nS | merged << 8 | (Pq & 15) << 12 | final << 16 (the slot after the tile's last record: no T entries)
};

__device__ __forceinline__ uint32_t tl_ref_code_at(const DevGenome &g, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;
	if ((__ldg(&g.nmask[gp >> 5]) >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// shared memory through 32-bit shared-window addresses
__device__ __forceinline__ uint32_t tl_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 tl_lds128(uint32_t addr) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint2 tl_lds64(uint32_t addr) {
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t tl_lds32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void tl_sts128(uint32_t addr, const uint4 &v) {
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t tl_lds8(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void tl_sts64(uint32_t addr, const uint2 &v) {
	asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void tl_sts32(uint32_t addr, uint32_t v) {
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tl_sts8(uint32_t addr, uint32_t v) {
	asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ---- names: pieces of at most 16 bytes appended through an accumulator word ----
// wp: shared address of the word under construction, acc: its f low bytes (bytes above are zero)
template <int NS>
struct NameW {
	uint32_t wp[NS], acc[NS], f[NS];
};
template <int NS>
__device__ __forceinline__ void name_begin(NameW<NS> &w, int s, uint32_t byteaddr) {
	w.wp[s] = byteaddr & ~3u;
	w.f[s] = byteaddr & 3u;
	const uint32_t old = tl_lds32(w.wp[s]);  // the bytes below belong to the previous record's last quality chunk
	w.acc[s] = old & ~(0xffffffffu << (8u * w.f[s]));
}
// appends the n low bytes of (w0, w1, w2, w3) (bytes from n on must be zero); NW = words that can be non-zero
template <int NS, int NW>
__device__ __forceinline__ void name_append(NameW<NS> &w, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t n) {
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint32_t sh = 8u * w.f[s];
		const uint32_t o0 = w.acc[s] | (w0 << sh);
		const uint32_t o1 = __funnelshift_l(w0, NW > 1 ? w1 : 0u, sh);
		const uint32_t o2 = NW > 1 ? __funnelshift_l(w1, NW > 2 ? w2 : 0u, sh) : 0u;
		const uint32_t o3 = NW > 2 ? __funnelshift_l(w2, NW > 3 ? w3 : 0u, sh) : 0u;
		const uint32_t o4 = NW > 3 ? __funnelshift_l(w3, 0u, sh) : 0u;
		const uint32_t tot = w.f[s] + n, nw = tot >> 2;
		if (nw > 0u) tl_sts32(w.wp[s], o0);
		if (NW > 1 && nw > 1u) tl_sts32(w.wp[s] + 4u, o1);
		if (NW > 2 && nw > 2u) tl_sts32(w.wp[s] + 8u, o2);
		if (NW > 3 && nw > 3u) tl_sts32(w.wp[s] + 12u, o3);
		uint32_t acc = nw == 0u ? o0 : o1;
		if (NW > 1) acc = nw == 2u ? o2 : acc;
		if (NW > 2) acc = nw == 3u ? o3 : acc;
		if (NW > 3) acc = nw == 4u ? o4 : acc;
		w.acc[s] = acc;
		w.wp[s] += 4u * nw;
		w.f[s] = tot & 3u;
	}
}
// the word under construction goes out; the bytes above its fill are kept (first sequence bytes)
template <int NS>
__device__ __forceinline__ void name_flush(NameW<NS> &w) {
#pragma unroll
	for (int s = 0; s < NS; s++) {
		if (w.f[s]) {
			const uint32_t keep = 0xffffffffu << (8u * w.f[s]);
			tl_sts32(w.wp[s], (tl_lds32(w.wp[s]) & keep) | w.acc[s]);
		}
	}
}
// slow path (refids over 16 bytes): bytes one at a time
template <int NS, class Get>
__device__ __noinline__ void name_bytes(NameW<NS> &w, uint32_t n, Get get) {
	name_flush<NS>(w);
#pragma unroll
	for (int s = 0; s < NS; s++) {
		uint32_t at = w.wp[s] + w.f[s];
		for (uint32_t k = 0; k < n; k++) tl_sts8(at + k, (uint32_t)(unsigned char)get(k));
		name_begin<NS>(w, s, at + n);
	}
}
// ":<strand>:<decimal of v>" with nd = number of digits of v (v < 2^32): four words, 3 + nd bytes
__device__ __forceinline__ void name_dec_piece(uint32_t v, uint32_t nd, bool fw, uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &p3) {
	const uint32_t hi = v / 100000u, lo = v - hi * 100000u;
	// five digits each, most significant first
	uint32_t x = hi, q;
	q = x / 10u; const uint32_t h4 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t h3 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t h2 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t h1 = x - q * 10u; const uint32_t h0 = q;
	x = lo;
	q = x / 10u; const uint32_t l4 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t l3 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t l2 = x - q * 10u; x = q;
	q = x / 10u; const uint32_t l1 = x - q * 10u; const uint32_t l0 = q;
	const uint32_t W0 = (h0 | (h1 << 8) | (h2 << 16) | (h3 << 24)) + 0x30303030u;
	const uint32_t W1 = (h4 | (l0 << 8) | (l1 << 16) | (l2 << 24)) + 0x30303030u;
	const uint32_t W2 = (l3 | (l4 << 8)) + 0x3030u;
	// drop the 10 - nd leading zeros
	const uint32_t k = 10u - nd, kw = k >> 2, kb = 8u * (k & 3u);
	const uint32_t t0 = kw == 0u ? W0 : (kw == 1u ? W1 : W2);
	const uint32_t t1 = kw == 0u ? W1 : (kw == 1u ? W2 : 0u);
	const uint32_t t2 = kw == 0u ? W2 : 0u;
	const uint32_t r0 = __funnelshift_r(t0, t1, kb), r1 = __funnelshift_r(t1, t2, kb), r2 = t2 >> kb;
	const uint32_t head = 0x3a003au | ((fw ? (uint32_t)'+' : (uint32_t)'-') << 8);  // ":+:"
	p0 = head | (r0 << 24);
	p1 = __funnelshift_l(r0, r1, 24u);
	p2 = __funnelshift_l(r1, r2, 24u);
	p3 = r2 >> 8;
}
// "<refid>:<+|->:<off>:<score>" (simplesim.cpp:104-110) without the leading refid separator; lead = the piece in
// front of the refid ("\n@qtip!:" or ":"), given as words
template <int NS>
__device__ __forceinline__ void name_origin(NameW<NS> &w, const char *__restrict__ refid, const uint4 &n16, uint32_t rl, uint64_t off,
                                            uint32_t nd, const MateHot &m) {
	if (rl <= 16u) name_append<NS, 4>(w, n16.x, n16.y, n16.z, n16.w, rl);
	else name_bytes<NS>(w, rl, [&](uint32_t k) { return refid[k]; });
	{  // (offsets fit 32 bits: the launcher sends genomes with longer records to the first-generation kernel)
		uint32_t p0, p1, p2, p3;
		name_dec_piece((uint32_t)off, nd, m.fw, p0, p1, p2, p3);
		name_append<NS, 4>(w, p0, p1, p2, p3, 3u + nd);
	}
	// ":<score>"
	name_append<NS, 4>(w, 0x3au | (m.score[0] << 8), __funnelshift_l(m.score[0], m.score[1], 8u), __funnelshift_l(m.score[1], m.score[2], 8u),
	                   m.score[2] >> 24, 1u + m.score_len);
}

// Shared memory of one warp tile: everything whose size is fixed by the tile's read count sits at compile-time
// offsets; the genome rows (PMAX pieces per read) and the images (STR entries per record) follow.
template <int CAT, int RW>
struct TileLay {
	static constexpr int NS = CAT == CAT_U ? 1 : 2;
	static constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;
	static constexpr uint32_t hot = 0;                                        // [NS][RW] THot
	static constexpr uint32_t hotq = hot + NS * RW * 16u;                     // [NS][RW] THotQ
	static constexpr uint32_t ship = hotq + NS * RW * 16u;                    // [NS][RW + 1] TShip
	static constexpr uint32_t ops = ship + NS * (RW + 1) * 8u;                // [NM][TL_OPS][RW] u32
	static constexpr uint32_t gunit = ops + NM * TL_OPS * RW * 4u;            // [NS][RW] 16-byte unit of bases[] at the row's start | pieces << 28
	static constexpr uint32_t gb0 = gunit + NS * RW * 4u;                     // [NS][RW] first reference base of the mate, in bases from the row's start
	static constexpr uint32_t pos = gb0 + NS * RW * 4u;                       // [NS][RW] TPos
	static constexpr uint32_t rnd = pos + NS * RW * 8u;                       // [RW] TRnd (bad-end pairs only)
	static constexpr uint32_t gbuf = (rnd + (CAT == CAT_B ? RW * 16u : 0u) + 15u) & ~15u;  // [NS][RW][PMAX] 16-byte pieces of packed genome
	static __host__ __device__ constexpr uint32_t img(uint32_t pmax) { return gbuf + NS * RW * 16u * pmax; }  // [NS][1 + (RW + 1) * STR] entries
	static __host__ __device__ constexpr uint32_t bytes(uint32_t pmax, uint32_t str) { return img(pmax) + NS * 16u * ((RW + 1) * str + 1u); }
};

template <int MODE, int CAT, int RW>
__global__ void __launch_bounds__(TL_MAX_WARPS * 32, CAT == CAT_U ? QT_TL_MIN_BLOCKS : QT_TL_MIN_BLOCKS_PAIR)
k_simulate_tile(SimArgs a, int K, uint32_t kmagic, int PMAX, uint32_t pmagic, int hS, uint32_t strmagic, uint32_t warp_smem, int skip) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;                    // output streams
	constexpr int NM = (CAT == CAT_C || CAT == CAT_D) ? 2 : 1;  // aligned mates per read
	using LY = TileLay<CAT, RW>;
	extern __shared__ __align__(16) unsigned char smem[];

	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
	const uint64_t r0 = tile * (uint64_t)RW;
	if (r0 >= a.nreads) return;
	const int nr = (int)((a.nreads - r0) < (uint64_t)RW ? (a.nreads - r0) : (uint64_t)RW);

	const int STR = hS + 2;                                                        // image entries per record
	const uint32_t img_stream = 16u * ((uint32_t)(RW + 1) * (uint32_t)STR + 1u);  // bytes per stream, with the guard entry in front
	const uint32_t row_bytes = 16u * (uint32_t)PMAX;
	const uint32_t wb = tl_smem_u32(smem) + (uint32_t)wib * warp_smem;             // this warp's slice, as a shared-window address
	const uint32_t gbuf0 = wb + LY::gbuf;
	// image base of stream 0: entry 0 of record 0 (the guard entry lies 16 bytes below); stream s is img_stream further
	const uint32_t img0 = gbuf0 + (uint32_t)(NS * RW) * row_bytes + 16u;

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
	TmplTail tt;
	if (CAT == CAT_U) {  // nothing of the tail block is needed: no second mate, no opposite mate, typ = "u"
		tt.delta[0] = tt.delta[1] = 0;
		tt.opp_len = 0;
		tt.typ = (uint32_t)'u';
		tt.typ_len = 1;
		tt.aligned_is_mate1 = false;
	} else tt = load_tail(tp);
	MateHot mh[NM];
#pragma unroll
	for (int m = 0; m < NM; m++) mh[m] = load_mate(&tp->m[m]);
	const uint64_t ro[2] = {pr.refoff + tt.delta[0], pr.refoff + tt.delta[1]};
	const bool s0_random = CAT == CAT_B && !tt.aligned_is_mate1;
	uint32_t nl = 0;
	bool my_general = false;
	const uint32_t slot_off = 16u * (uint32_t)(lane * STR);  // this read's image, from the stream's image base
	if (act) {
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
		const uint32_t nunits = (uint32_t)(a.genome_words >> 2);
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
			// A random mate drawn by the native generator is printed like a genome span: its row holds the mate's 2-bit
			// codes from base 16 on (step A2r), read on the aligned mate's strand
			const bool rnd_row = random && MODE == QTIPSIM_RNG_PHILOX;
			const uint64_t gadj = rnd_row ? (md.fw ? 16u - (P & 15u) : 16u + (L - 1u) + (P & 15u) - 15u)
			                    : md.fw   ? gpos + (uint64_t)(int64_t)md.ref0 - (P & 15u)
			                              : gpos + (uint64_t)(int64_t)md.ref0 + (P & 15u) - 15u;
			const uint32_t ns = random && !rnd_row ? 0u : (uint32_t)(c_last - c_first + 1);
			const bool gen_s = general[mate] && !random;
			const bool fast = !random && !general[mate];
			// the read's row of packed genome words starts at a 16-byte unit of bases[]
			const uint64_t wmin = md.fw ? (gadj >> 4) : (gadj >> 4) - (uint64_t)(ns ? ns - 1u : 0u);
			const uint32_t gunit = rnd_row ? 0u : (uint32_t)(wmin >> 2);
			const uint32_t w0 = (uint32_t)((gadj >> 4) - 4ull * gunit);  // word of slot 0 in the read's row
			const uint32_t wmax = md.fw ? w0 + ns : w0 + 1u;             // last word a chunk lane touches
			uint32_t np = (wmax + 4u) >> 2;                              // 16-byte pieces staged
			if (gunit + np > nunits) np = gunit < nunits ? nunits - gunit : 0u;
			const uint32_t ix = (uint32_t)(s * RW + lane);
			tl_sts32(wb + LY::gunit + 4u * ix, fast ? (gunit | (np << 28)) : 0u);
			tl_sts32(wb + LY::gb0 + 4u * ix, (uint32_t)(gpos - 64ull * gunit));
			const uint32_t meta = (md.fw ? RF_FW : 0u) | (random ? RF_RANDOM : 0u) | (gen_s ? RF_GENERAL : 0u) |
			                      ((uint32_t)(gadj & 15u) << 8) | ((A0 & 15u) << 12);
			const uint32_t s0c = lane == 0 ? 0u : (rstart[s] - 2u) >> 4;
			const uint32_t nS = (uint32_t)c_first - s0c + 1u;
			// THot: grow | sdst | ns, meta | img
			tl_sts128(wb + LY::hot + 16u * ix, make_uint4(gbuf0 + ix * row_bytes + 4u * w0, 16u * (uint32_t)c_first, ns | (meta << 16),
			                                              (slot_off + 16u * (nS - 1u)) | ((slot_off + 16u * (uint32_t)hS) << 16)));
			// THotQ: qsrc | qdst | nq, meta | img
			tl_sts128(wb + LY::hotq + 16u * ix, make_uint4(A0 >> 4, 16u * (uint32_t)q_first, (uint32_t)(q_last - q_first + 1) | (meta << 16),
			                                               (slot_off + 16u * (uint32_t)(hS + 1)) | ((slot_off + 16u * (uint32_t)STR) << 16)));
			tl_sts64(wb + LY::pos + 8u * ix, make_uint2(P, L));
			QT_CHK(ns <= (uint32_t)K && (uint32_t)(q_last - q_first + 1) <= (uint32_t)K && (!fast || np <= (uint32_t)PMAX), 9, ns, K);
			QT_CHK(nS <= (uint32_t)hS && nS >= 2u && c_last > c_first && q_last > q_first, 5, nS, hS);
			if (CAT == CAT_B && random) tl_sts32(wb + LY::rnd + 16u * (uint32_t)lane + 8u, (uint32_t)s | (md.fw ? 2u : 0u));  // TRnd::flags
			if (gen_s) {
				// (no packed genome is staged for a general-path read: its row holds the TGen record instead)
				tl_sts128(gbuf0 + ix * row_bytes, make_uint4((uint32_t)gpos, (uint32_t)(gpos >> 32), (uint32_t)gend, (uint32_t)(gend >> 32)));
				tl_sts64(gbuf0 + ix * row_bytes + 16u, make_uint2(md.seg_off, md.nseg));
			}
			// TShip: s0 | c_last << 16, nS | T entries to ship << 8 (one when sequence tail and quality head share the chunk)
			tl_sts64(wb + LY::ship + 8u * (uint32_t)(s * (RW + 1) + lane), make_uint2(s0c | ((uint32_t)c_last << 16), nS | (q_first == c_last ? 0x100u : 0x200u)));
			if (lane == nr - 1)  // the slot after the tile's last record: its last quality chunk and '\n', no T entries
				tl_sts64(wb + LY::ship + 8u * (uint32_t)(s * (RW + 1) + nr), make_uint2((uint32_t)q_last, ((tend[s] - 1u) >> 4) - (uint32_t)q_last + 1u));
		}
	}
	__syncwarp();
	const bool any_general = __any_sync(0xffffffffu, my_general);

	// ---------------- step A2 (lane = 16-byte piece): the reads' packed genome rows, asynchronously ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const int nitems = nr * PMAX;
		for (int q = lane; q < nitems; q += 32) {
			const int i = (int)__umulhi((uint32_t)q, pmagic), p = q - i * PMAX;
			const uint32_t u = tl_lds32(wb + LY::gunit + 4u * (uint32_t)(s * RW + i));
			if ((uint32_t)p < (u >> 28)) {
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gbuf0 + (uint32_t)(s * RW) * row_bytes + 16u * (uint32_t)q),
				             "l"((const uint4 *)a.g.bases + (u & 0x0fffffffu) + p)
				             : "memory");
			}
		}
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
	if (CAT == CAT_B && MODE == QTIPSIM_RNG_PHILOX) {
		// ---- step A2r (lane = 64 bases of a random mate): Philox blocks = the mate's packed codes, into its row from word 1 on ----
		const int nbk = (K + 2) >> 2;  // blocks a mate can need
		for (int b0 = 0; b0 < nbk; b0 += 4) {
			for (int q = lane; q < 4 * nr; q += 32) {
				const int i = q >> 2, b = b0 + (q & 3);
				const uint32_t sr = tl_lds32(wb + LY::rnd + 16u * (uint32_t)i + 8u) & 1u;
				const uint32_t L = tl_lds32(wb + LY::pos + 8u * (sr * RW + (uint32_t)i) + 4u);
				if (64u * (uint32_t)b < L) {
					const Philox4 blk = philox_rmate_block(a.read_index_base + r0 + (uint64_t)i, (uint32_t)b, CAT, a.philox_seed);
					const uint32_t ra = gbuf0 + (sr * RW + (uint32_t)i) * row_bytes + 4u + 16u * (uint32_t)b;
					tl_sts32(ra, blk.v[0]);
					tl_sts32(ra + 4u, blk.v[1]);
					tl_sts32(ra + 8u, blk.v[2]);
					tl_sts32(ra + 12u, blk.v[3]);
				}
			}
		}
	}

	// ---------------- the first draw ops of each read (lane = read): loaded now, put into shared memory after step Bq ----------------
	uint4 op_a[NM];
	uint2 op_b[NM];
#pragma unroll
	for (int m = 0; m < NM; m++) {
		op_a[m] = make_uint4(0u, 0u, 0u, 0u);
		op_b[m] = make_uint2(0u, 0u);
		if (act && mh[m].nops) {
			const uint4 *op4 = (const uint4 *)(a.c.ops + mh[m].ops_off);
			op_a[m] = __ldg(op4);
			if (mh[m].nops > 4 && TL_OPS > 4) op_b[m] = __ldg((const uint2 *)(op4 + 1));
		}
	}
	if (skip & 1) return;  // attribution runs only (QTIPSIM_SKIP): output incomplete

	// ---------------- step Bq (lane = quality chunk): item q = read * K + slot, read-major ----------------
	// Interior chunks go from registers straight to HBM; a string's first and last chunk go whole into the image.
#pragma unroll
	for (int s = 0; s < NS; s++) {
		char *tbs = tb[s];
		const uint32_t hot_a = wb + LY::hotq + 16u * (uint32_t)(s * RW);
		const uint32_t imgs = img0 + (uint32_t)s * img_stream;
		const int nitems = nr * K;
		for (int q0 = lane; q0 < nitems && !(skip & 8); q0 += 32 * TL_UNROLL) {
			QualLoad ql[TL_UNROLL];
			uint32_t meta[TL_UNROLL], dst[TL_UNROLL], ar[TL_UNROLL];  // ar: image entry for a string's first / last chunk, else 0
			bool live[TL_UNROLL];
#pragma unroll
			for (int u = 0; u < TL_UNROLL; u++) {
				const int q = q0 + 32 * u;
				live[u] = false;
				if (q < nitems) {
					const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
					const uint4 h4 = tl_lds128(hot_a + 16u * (uint32_t)i);  // qsrc | qdst | nq, meta | img
					const int nq = (int)(h4.z & 0xffffu);
					meta[u] = h4.z >> 16;
					live[u] = j < nq;
					dst[u] = h4.y + 16u * (uint32_t)j;
					ar[u] = j == 0 ? imgs + (h4.w & 0xffffu) : (j == nq - 1 ? imgs + (h4.w >> 16) : 0u);
					if (live[u] && !(CAT == CAT_B && (meta[u] & RF_RANDOM))) ql[u] = qual_issue(a.c.qual_out, h4.x + (uint32_t)j);
				}
			}
#pragma unroll
			for (int u = 0; u < TL_UNROLL; u++) {
				if (!live[u]) continue;
				const uint4 v = (CAT == CAT_B && (meta[u] & RF_RANDOM)) ? make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u)  // 'I' (simplesim.h:105)
				                                      : qual_finish(ql[u], (meta[u] >> 12) & 15u);
				if (ar[u]) tl_sts128(ar[u], v);
				else {
					QT_CHK(dst[u] + 16u <= tend[s], 4, dst[u], tend[s]);
					TST16(tbs + dst[u], v);
				}
			}
		}
	}
	if (act) {
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const uint32_t so = wb + LY::ops + 4u * (uint32_t)(m * TL_OPS * RW + lane);
			tl_sts32(so, op_a[m].x);
			tl_sts32(so + 4u * RW, op_a[m].y);
			tl_sts32(so + 8u * RW, op_a[m].z);
			tl_sts32(so + 12u * RW, op_a[m].w);
			if (TL_OPS > 4) tl_sts32(so + 16u * RW, op_b[m].x);
			if (TL_OPS > 5) tl_sts32(so + 20u * RW, op_b[m].y);
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
		const uint32_t hot_a = wb + LY::hot + 16u * (uint32_t)(s * RW);
		const uint32_t imgs = img0 + (uint32_t)s * img_stream;
		const int nitems = nr * K;
		for (int q = lane; q < nitems && !(skip & 16); q += 32) {
			const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
			const uint4 h4 = tl_lds128(hot_a + 16u * (uint32_t)i);  // grow | sdst | ns, meta | img
			const int ns = (int)(h4.z & 0xffffu);
			const uint32_t meta = h4.z >> 16;
			if (j >= ns || (meta & RF_GENERAL)) continue;
			const uint32_t wa = (meta & RF_FW) ? h4.x + 4u * (uint32_t)j : h4.x - 4u * (uint32_t)j;
			SeqLoad sl;
			sl.lo = tl_lds32(wa);
			sl.hi = tl_lds32(wa + 4u);
			const uint4 v = seq_finish(sl, meta);
			if (j == 0 || j == ns - 1) tl_sts128(imgs + (j ? h4.w >> 16 : h4.w & 0xffffu), v);
			else {
				QT_CHK(h4.y + 16u * (uint32_t)j + 16u <= tend[s], 3, h4.y, tend[s]);
				TST16(tbs + h4.y + 16u * (uint32_t)j, v);
			}
		}
		if (any_general) {
			for (int q = lane; q < nitems && !(skip & 16); q += 32) {
				const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
				const uint4 h4 = tl_lds128(hot_a + 16u * (uint32_t)i);
				const int ns = (int)(h4.z & 0xffffu);
				const uint32_t meta = h4.z >> 16;
				if (j >= ns || !(meta & RF_GENERAL)) continue;
				const uint32_t dst = h4.y + 16u * (uint32_t)j;
				const uint2 pp = tl_lds64(wb + LY::pos + 8u * (uint32_t)(s * RW + i));
				const uint32_t ga = gbuf0 + (uint32_t)(s * RW + i) * row_bytes;  // (a general-path read's row holds its TGen)
				const uint4 g4 = tl_lds128(ga);
				const uint2 g2 = tl_lds64(ga + 16u);
				const uint4 v = seq16_general(gv, (const uint4 *)g_tl_lowmask, a.c.segs + g2.x, (int)g2.y, (meta & RF_FW) != 0, (int)pp.y,
				                              (uint64_t)g4.x | ((uint64_t)g4.y << 32), (uint64_t)g4.z | ((uint64_t)g4.w << 32), (int)dst - (int)pp.x);
				if (j == 0 || j == ns - 1) tl_sts128(imgs + (j ? h4.w >> 16 : h4.w & 0xffffu), v);
				else TST16(tbs + dst, v);
			}
		}
	}
	__syncwarp();
	if (skip & 2) return;

	// ---------------- step C: the draws, patched over image or HBM bytes ----------------
	// where byte x of a stream's sequence lives: first chunk -> the image's last S entry, last chunk -> T_seq,
	// interior -> HBM (already stored by step Bs: same warp, ordered by __syncwarp)
	auto put_at = [&](char *tbx, uint32_t simg, uint32_t timg, uint32_t c_first, uint32_t c_last, uint32_t x, uint32_t ch) {
		const uint32_t cx = x >> 4;
		if (cx == c_first) tl_sts8(simg + (x & 15u), ch);
		else if (cx == c_last) tl_sts8(timg + (x & 15u), ch);
		else tbx[x] = (char)ch;
	};
	// reference base under an X op of stream-slot ix (0..3, or 4 for N and positions past the record end)
	auto ref_base = [&](uint32_t ix, bool gen, uint32_t gb0, uint32_t roff) -> uint32_t {
		if (gen) {
			const uint4 g4 = tl_lds128(gbuf0 + ix * row_bytes);
			return tl_ref_code_at(a.g, ((uint64_t)g4.x | ((uint64_t)g4.y << 32)) + roff, (uint64_t)g4.z | ((uint64_t)g4.w << 32));
		}
		const uint32_t t = gb0 + roff;
		return (tl_lds32(gbuf0 + ix * row_bytes + 4u * (t >> 4)) >> (2u * (t & 15u))) & 3u;
	};
	constexpr bool dual = CAT == CAT_U && MODE == QTIPSIM_RNG_PHILOX && RW <= 16;
	constexpr bool pair_lanes = (CAT == CAT_C || CAT == CAT_D) && MODE == QTIPSIM_RNG_PHILOX && RW <= 16;
	if (dual || pair_lanes) {
		// Native generator, tiles of at most 16 reads: two lanes per read.  Unpaired: each lane takes every second draw op
		// (Philox draws are addressed by their index, so the ops of a read need not be walked in order).  Pairs: lanes
		// 0-15 serve the first file and mate, lanes 16-31 the second, whose draws start at index nops of the first.
		const int half = lane >> 4, rl_ = lane & 15;
		const int sx = pair_lanes ? half : 0;
		const uint32_t n0 = __shfl_sync(0xffffffffu, mh[0].nops, rl_), n1 = __shfl_sync(0xffffffffu, mh[NM - 1].nops, rl_);
		const uint32_t o0 = __shfl_sync(0xffffffffu, mh[0].ops_off, rl_), o1 = __shfl_sync(0xffffffffu, mh[NM - 1].ops_off, rl_);
		const uint32_t f0 = __shfl_sync(0xffffffffu, mh[0].fw ? 1u : 0u, rl_), f1 = __shfl_sync(0xffffffffu, mh[NM - 1].fw ? 1u : 0u, rl_);
		const uint32_t c_nops = sx ? n1 : n0, c_opsoff = sx ? o1 : o0, c_fw = sx ? f1 : f0;
		const uint32_t d_first = pair_lanes ? (sx ? n0 : 0u) : 0u;        // draw index of this lane's first op
		const uint32_t j_first = dual ? (uint32_t)half : 0u, j_step = dual ? 2u : 1u;
		if (rl_ < nr && c_nops > j_first) {
			const uint32_t ix = (uint32_t)(sx * RW + rl_);
			const uint2 pp = tl_lds64(wb + LY::pos + 8u * ix);
			const uint32_t P = pp.x, L = pp.y;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4;
			const uint4 hh = tl_lds128(wb + LY::hot + 16u * ix);
			const uint32_t imgs = img0 + (uint32_t)sx * img_stream;
			const uint32_t simg = imgs + (hh.w & 0xffffu), timg = imgs + (hh.w >> 16);
			const uint64_t gi = a.read_index_base + r0 + (uint64_t)rl_;
			const bool gen = ((hh.z >> 16) & RF_GENERAL) != 0;
			const uint32_t T = c_fw ? TBL_FW : TBL_RC;
			const uint32_t gb0 = tl_lds32(wb + LY::gb0 + 4u * ix);
			const uint32_t *ops = a.c.ops + c_opsoff;
			char *tbx = sx ? tb[NS - 1] : tb[0];
			uint32_t curb = 0xffffffffu;
			Philox4 blk{};
			for (uint32_t j = j_first; j < c_nops; j += j_step) {
				const uint32_t e = j < (uint32_t)TL_OPS ? tl_lds32(wb + LY::ops + 4u * ((uint32_t)(sx * TL_OPS) * RW + j * RW + (uint32_t)rl_)) : __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				QT_CHK(outp < L, 7, outp, L);
				const uint32_t d = d_first + j;  // draw index within the read / pair: 16 bits of Philox block d / 8
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
					const uint32_t ref = ref_base(ix, gen, gb0, e >> 17);
					code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
				} else code = b >> 14;
				put_at(tbx, simg, timg, c_first, c_last, P + outp, (T >> (8u * code)) & 0xffu);
			}
		}
	} else if (act) {
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
#pragma unroll
		for (int m = 0; m < NM; m++) {
			const int s = CAT == CAT_B ? (s0_random ? 1 : 0) : m;  // stream that holds aligned mate m
			const uint32_t n = mh[m].nops;
			if (!n) continue;
			const uint32_t ix = (uint32_t)(s * RW + lane);
			const uint2 pp = tl_lds64(wb + LY::pos + 8u * ix);
			const uint32_t P = pp.x, L = pp.y;
			const uint32_t c_first = P >> 4, c_last = (P + L - 1u) >> 4;
			const uint4 hh = tl_lds128(wb + LY::hot + 16u * ix);
			const uint32_t imgs = img0 + (uint32_t)s * img_stream;
			const uint32_t simg = imgs + (hh.w & 0xffffu), timg = imgs + (hh.w >> 16);
			const bool gen = ((hh.z >> 16) & RF_GENERAL) != 0;
			const uint32_t T = mh[m].fw ? TBL_FW : TBL_RC;
			const uint32_t gb0 = tl_lds32(wb + LY::gb0 + 4u * ix);
			const uint32_t *ops = a.c.ops + mh[m].ops_off;
			for (uint32_t j = 0; j < n; j++) {
				const uint32_t e = j < (uint32_t)TL_OPS ? tl_lds32(wb + LY::ops + 4u * ((uint32_t)(m * TL_OPS) * RW + j * RW + (uint32_t)lane)) : __ldg(&ops[j]);
				const uint32_t outp = (e >> 2) & 0x7fffu;
				QT_CHK(outp < L, 7, outp, L);
				uint32_t code;
				if ((e & 3u) == OP_X) {
					const uint32_t ref = ref_base(ix, gen, gb0, e >> 17);
					if (MODE == QTIPSIM_RNG_REPLAY) {
						do code = draw_code();  // simplesim.cpp:47-49
						while (code == ref);
					} else {
						const uint32_t b = draw16();
						code = ref < 4u ? ((ref + 1u + ((b * 3u) >> 16)) & 3u) : (b >> 14);
					}
				} else code = draw_code();
				put_at(tb[s], simg, timg, c_first, c_last, P + outp, (T >> (8u * code)) & 0xffu);
			}
		}
		if (CAT == CAT_B && MODE == QTIPSIM_RNG_REPLAY)
			// the random mate (init_random, simplesim.h:89-117) follows in the draw stream: hand the generator to step Br
			tl_sts64(wb + LY::rnd + 16u * (uint32_t)lane, make_uint2(st.s1, st.s2));
	}
	if (CAT == CAT_B && MODE == QTIPSIM_RNG_REPLAY) {
		__syncwarp();
		// ---------------- step Br (REPLAY; lane = 16-byte chunk of a random mate): opp_len draws per read, 16 per lane ----------------
		// The L'Ecuyer stream is addressable: draw j of the mate is the state (j + 1) steps past the handed-over one
		// (jump tables), walked forwards for '+' mates and backwards -- a multiplication by a^-1 -- for '-' mates, whose
		// bases are printed in reverse (simplesim.h:108-113).  (The native generator's random mates are rows: step A2r.)
		const int nitems = nr * K;
		for (int q = lane; q < nitems; q += 32) {
			const int i = (int)__umulhi((uint32_t)q, kmagic), j = q - i * K;
			const uint4 rd = tl_lds128(wb + LY::rnd + 16u * (uint32_t)i);  // r0 | r1 | flags: bit 0 stream of the random mate, bit 1 '+' strand
			const int sr = (int)(rd.z & 1u);
			const bool fw = (rd.z & 2u) != 0;
			const uint2 pp = tl_lds64(wb + LY::pos + 8u * (uint32_t)(sr * RW + i));
			const int P = (int)pp.x, L = (int)pp.y;
			const int c_first = P >> 4, ns = ((P + L - 1) >> 4) - c_first + 1;
			if (j >= ns) continue;
			const int k0 = 16 * (c_first + j) - P;       // output position of the chunk's first byte
			const int t_first = k0 < 0 ? -k0 : 0;        // first byte of the chunk inside the mate
			const uint32_t T = fw ? TBL_FW : TBL_RC;
			uint32_t w[4] = {0u, 0u, 0u, 0u};
			{
				const int jd0 = fw ? k0 + t_first : L - 1 - (k0 + t_first);  // draw of the first byte inside
				LState st = ljump_tbl(LState{rd.x, rd.y}, (uint32_t)(fw ? jd0 : jd0 + 1), a.jt1, a.jt2);
#pragma unroll
				for (int t = 0; t < 16; t++) {
					if (t >= t_first && k0 + t < L) {
						if (fw) st = lstep(st);
						const uint32_t code = lbase_code(lfloat(lout(st)));
						w[t >> 2] |= ((T >> (8u * code)) & 0xffu) << (8 * (t & 3));
						if (!fw) st = lstep_back(st);
					}
				}
			}
			const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
			if (j == 0 || j == ns - 1) {
				const uint32_t im = tl_lds32(wb + LY::hot + 16u * (uint32_t)(sr * RW + i) + 12u);
				tl_sts128(img0 + (uint32_t)sr * img_stream + (j ? im >> 16 : im & 0xffffu), v);
			} else TST16((sr ? tb[NS - 1] : tb[0]) + 16u * (size_t)(c_first + j), v);
		}
	}
	__syncwarp();

	// ---------------- step N (lane = record): '\n' of the previous record, name, separators ----------------
	// Everything that stores whole chunks into the images is done; what follows only touches its own bytes.  Pairs:
	// lanes 0-15 print the first file's copy of the name, lanes 16-31 the second file's.
	{
		static_assert(RW == 16, "step N: two lanes per pair");
		const int rl_ = NS == 2 ? (lane & 15) : lane, sx = NS == 2 ? (lane >> 4) : 0;
		uint32_t rs = __shfl_sync(0xffffffffu, rstart[0], rl_);
		if (NS == 2) {
			const uint32_t rs1 = __shfl_sync(0xffffffffu, rstart[NS - 1], rl_);
			rs = sx ? rs1 : rs;
		}
		const uint32_t nlx = __shfl_sync(0xffffffffu, nl, rl_);
		if (rl_ < nr) {
			// The name's inputs are loaded again (L1 / L2 hits) rather than held in registers across the chunk loops: the
			// plan record, the template's mates and tail block, the record's name blocks
			const PlanRec pr = a.plan[r0 + (uint64_t)rl_];
			const TmplDesc *tp = a.c.t + pr.tmpl;
			const uint4 *recp4 = (const uint4 *)(a.g.recs + pr.record);
			const uint4 rec_b = __ldg(recp4 + 1), rec_c = __ldg(recp4 + 2);
			TmplTail tt;
			if (CAT == CAT_U) {
				tt.delta[0] = tt.delta[1] = 0;
				tt.opp_len = 0;
				tt.typ = (uint32_t)'u';
				tt.typ_len = 1;
				tt.aligned_is_mate1 = false;
			} else tt = load_tail(tp);
			MateHot mh[NM];
#pragma unroll
			for (int m = 0; m < NM; m++) mh[m] = load_mate(&tp->m[m]);
			const uint64_t ro[2] = {pr.refoff + tt.delta[0], pr.refoff + tt.delta[1]};
			const bool s0_random = CAT == CAT_B && !tt.aligned_is_mate1;
			const uint32_t rec_name_off = rec_b.x, rl = rec_b.y;
			const char *refid = a.g.names + rec_name_off;
			const uint32_t imgx = img0 + (uint32_t)sx * img_stream + 16u * (uint32_t)(rl_ * STR);  // the record's image
			const uint32_t s0c = rl_ == 0 ? 0u : (rs - 2u) >> 4;
			NameW<1> w;
			// byte rstart - 1 (for the tile's first record it may lie in the guard entry: never shipped)
			name_begin<1>(w, 0, imgx + (rs - 1u - 16u * s0c));
			name_append<1, 2>(w, 0x7471400au, 0x3a217069u, 0u, 0u, 8u);  // "\n@qtip!:"
			if (CAT == CAT_U) {
				const uint32_t nd = nlx - 13u - rl - mh[0].score_len - tt.typ_len;
				name_origin<1>(w, refid, rec_c, rl, ro[0], nd, mh[0]);
			} else if (CAT == CAT_B) {
				const uint32_t nd = (nlx - 18u - 2u * rl - 2u * mh[0].score_len - tt.typ_len) >> 1;
				name_origin<1>(w, refid, rec_c, rl, ro[0], nd, mh[0]);
				name_append<1, 1>(w, 0x3au, 0u, 0u, 0u, 1u);
				name_origin<1>(w, refid, rec_c, rl, ro[0], nd, mh[0]);
			} else {
				const uint32_t nd0 = ndig64(ro[0]);
				const uint32_t nd1 = nlx - 18u - 2u * rl - mh[0].score_len - mh[NM - 1].score_len - tt.typ_len - nd0;
				name_origin<1>(w, refid, rec_c, rl, ro[0], nd0, mh[0]);
				name_append<1, 1>(w, 0x3au, 0u, 0u, 0u, 1u);
				name_origin<1>(w, refid, rec_c, rl, ro[1], nd1, mh[NM - 1]);
			}
			name_append<1, 1>(w, 0x3au | (tt.typ << 8) | (0x0au << (8u + 8u * tt.typ_len)), 0u, 0u, 0u, 2u + tt.typ_len);  // ":<typ>\n"
			name_flush<1>(w);
			const bool random = CAT == CAT_B && ((sx == 0) == s0_random);
			const uint32_t L = random ? tt.opp_len : (CAT == CAT_B || !sx ? mh[0].len : mh[NM - 1].len);
			const uint32_t P = rs + nlx, Pq = P + L + 3u;
			QT_CHK(w.wp[0] + w.f[0] == imgx + (P - 16u * s0c), 6, w.wp[0] + w.f[0], P);
			const uint32_t c_last = (P + L - 1u) >> 4;
			const uint32_t tseq = imgx + 16u * (uint32_t)hS;
			// separators: each byte into T_seq when it lies in the chunk of the last sequence byte, else into T_qual
			const uint32_t x0 = P + L, x1 = x0 + 1u, x2 = x0 + 2u;
			tl_sts8(tseq + ((x0 >> 4) == c_last ? 0u : 16u) + (x0 & 15u), 0x0au);
			tl_sts8(tseq + ((x1 >> 4) == c_last ? 0u : 16u) + (x1 & 15u), 0x2bu);
			tl_sts8(tseq + ((x2 >> 4) == c_last ? 0u : 16u) + (x2 & 15u), 0x0au);
			if ((Pq >> 4) == c_last) {
				// sequence tail, separators and quality head share one chunk of the file: its bytes from Pq & 15 on come
				// from the T_qual entry
				uint4 v = tl_lds128(tseq);
				const uint4 vq = tl_lds128(tseq + 16u);
				const uint32_t qph = Pq & 15u;
				const uint32_t wq = qph >> 2, lowm = ~(0xffffffffu << (8u * (qph & 3u)));  // word of the seam, its low bytes
				const uint32_t m0 = wq > 0u ? 0xffffffffu : lowm, m1 = wq > 1u ? 0xffffffffu : (wq == 1u ? lowm : 0u);
				const uint32_t m2 = wq > 2u ? 0xffffffffu : (wq == 2u ? lowm : 0u), m3 = wq == 3u ? lowm : 0u;
				v.x = (v.x & m0) | (vq.x & ~m0);
				v.y = (v.y & m1) | (vq.y & ~m1);
				v.z = (v.z & m2) | (vq.z & ~m2);
				v.w = (v.w & m3) | (vq.w & ~m3);
				tl_sts128(tseq, v);
			}
			if (rl_ == nr - 1) {  // no next record in this tile to print the record's last '\n'
				const uint32_t q_last = (Pq + L - 1u) >> 4;
				tl_sts8(imgx + 16u * (uint32_t)STR + (Pq + L - 16u * q_last), 0x0au);
			}
		}
	}
	__syncwarp();
	if (skip & 4) return;

	// ---------------- step D (lane = image entry): images -> HBM ----------------
#pragma unroll
	for (int s = 0; s < NS; s++) {
		char *tbs = tb[s];
		const uint32_t end = tend[s], phs = ph[s];
		const uint32_t last_chunk = end >> 4;         // shared with the next tile when end & 15 != 0 (else nothing of ours)
		const uint32_t head_chunk = phs ? 0u : ~0u;   // chunk 0 is shared with the previous tile when the tile starts inside it
		const uint32_t imgs = img0 + (uint32_t)s * img_stream;
		const uint32_t ship_a = wb + LY::ship + 8u * (uint32_t)(s * (RW + 1));
		const int nitems = (nr + 1) * STR;
		for (int e = lane; e < nitems; e += 32) {
			const int i = (int)__umulhi((uint32_t)e, strmagic), k = e - i * STR;
			const uint2 sh = tl_lds64(ship_a + 8u * (uint32_t)i);
			const bool isT = k >= hS;
			const uint32_t lim = isT ? (sh.y >> 8) + (uint32_t)hS : (sh.y & 0xffu);
			const uint32_t chunk = (isT ? (sh.x >> 16) - (uint32_t)hS : (sh.x & 0xffffu)) + (uint32_t)k;
			if ((uint32_t)k >= lim || chunk == head_chunk || chunk == last_chunk) continue;  // (the tile's ragged ends: below)
			QT_CHK(16u * chunk + 16u <= ((end + 15u) & ~15u), 8, chunk, end);
			TST16(tbs + 16u * (size_t)chunk, tl_lds128(imgs + 16u * (uint32_t)e));
		}
		// the tile's first and last chunk are shared with the neighbouring tiles: one byte per lane
		{
			const uint32_t zs0 = tl_lds32(ship_a + 8u * (uint32_t)nr);
			const uint32_t b = (uint32_t)lane & 15u;
			if (lane < 16) {
				if (phs && b >= phs) tbs[b] = (char)tl_lds8(imgs + b);  // record 0, entry 0
			} else if (b < (end & 15u)) tbs[16u * (size_t)last_chunk + b] = (char)tl_lds8(imgs + 16u * ((uint32_t)(nr * STR) + (last_chunk - zs0)) + b);
		}
	}
}

template <int MODE, int CAT>
static bool launch_cat(const SimArgs &a, uint32_t max_name, uint32_t max_body, cudaStream_t st) {
	constexpr int NS = CAT == CAT_U ? 1 : 2;
	constexpr int RW = 16;  // reads per warp tile (the sweeps of the first generation: profiles/r02_k3_sweep.log)
	using LY = TileLay<CAT, RW>;
	static int env_warps = -1;
	if (env_warps < 0) {
		const char *e = getenv("QTIPSIM_WARPS");
		env_warps = e ? atoi(e) : 0;
	}
	if ((a.genome_words >> 2) >= (1ull << 28)) return false;  // row descriptors keep the 16-byte unit in 28 bits
	// chunks a string of max_body bytes can touch at any alignment
	const uint32_t K = (max_body + 14u) / 16u + 1u;
	const uint32_t kmagic = (uint32_t)((0x100000000ull + K - 1) / K);  // q / K == umulhi(q, kmagic) for q < 2^16
	// 16-byte pieces of packed genome per read: K + 1 words plus up to 3 of alignment
	const uint32_t PMAX = ((K + 1u + 3u + 3u) / 4u) | 1u;  // odd: per-read strides are odd multiples of 16 bytes
	if (PMAX > 15u) return false;                          // (piece count kept in 4 bits)
	const uint32_t pmagic = (uint32_t)((0x100000000ull + PMAX - 1) / PMAX);
	// S entries of an image: from the chunk of (record start - 2) to the chunk of the first sequence byte
	uint32_t hS = ((max_name + 17u) >> 4) + 1u;  // name + 3 bytes (record start - 2 .. first sequence byte) at any alignment
	if (((hS + 2u) & 1u) == 0) hS++;  // entries per record: odd (bank spread of the lane = read steps)
	const uint32_t STR = hS + 2u;
	const uint32_t strmagic = (uint32_t)((0x100000000ull + STR - 1) / STR);
	const uint32_t wbytes = (LY::bytes(PMAX, STR) + 15u) & ~15u;
	const uint32_t sm_budget = 216u * 1024u;
	// resident warps per SM wanted: 24 (unpaired), 12 (pair categories: two images and rows per pair); longer records
	// and names go to the first-generation kernel
	if (wbytes * (CAT == CAT_U ? 24u : 12u) > sm_budget) return false;
	if ((uint64_t)RW * (2ull * K + STR + 2ull) >= 65536u) return false;  // chunk numbers within a tile are 16-bit
	if (16u * ((uint32_t)(RW + 1) * STR + 1u) >= 65536u) return false;   // image offsets are 16-bit
	int warps = env_warps > 0 ? env_warps : (NS == 1 ? 1 : TL_MAX_WARPS);  // tiles per CTA: measured (QTIPSIM_WARPS overrides)
	if (warps > TL_MAX_WARPS) warps = TL_MAX_WARPS;
	while (warps > 1 && wbytes * (uint32_t)warps + 1024 > sm_budget) warps >>= 1;
	const size_t smem = (size_t)wbytes * warps;
	auto kern = k_simulate_tile<MODE, CAT, RW>;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
		attr_set = true;
	}
	const uint64_t ntiles = (a.nreads + RW - 1) / RW;
	const uint64_t nblocks = (ntiles + warps - 1) / warps;
	static int skip = -1;  // QTIPSIM_SKIP (attribution runs only): bit 0 stop after step A, 1 after B, 2 after C / N
	if (skip < 0) {
		const char *e = getenv("QTIPSIM_SKIP");
		skip = e ? atoi(e) : 0;
	}
	kern<<<(unsigned)nblocks, warps * 32, smem, st>>>(a, (int)K, kmagic, (int)PMAX, pmagic, (int)hS, strmagic, wbytes, skip);
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
// of the category.  Returns false when the category cannot use this kernel (a body under 17 bytes, names too
// long for the images, a genome over 2^34 bases): the caller falls back.
bool launch_simulate_tile(const SimArgs &a, uint32_t max_name, uint32_t min_body, uint32_t max_body, cudaStream_t st) {
	if (!a.nreads) return true;
	if (min_body < 17u) return false;
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) return launch_mode<QTIPSIM_RNG_REPLAY>(a, max_name, max_body, st);
	return launch_mode<QTIPSIM_RNG_PHILOX>(a, max_name, max_body, st);
}

#ifdef QT_BOUNDS
void tile_debug_report() {
	unsigned long long v[4] = {0, 0, 0, 0};
	cudaMemcpyFromSymbol(v, g_qt_violation, sizeof(v));
	if (v[0])
		fprintf(stderr, "QT_BOUNDS(tile) violation id=%llu v0=%lld v1=%lld block=%llu thread=%llu\n", v[0], (long long)v[1],
		        (long long)v[2], v[3] >> 32, v[3] & 0xffffffffu);
	else fprintf(stderr, "QT_BOUNDS(tile): no violation\n");
	memset(v, 0, sizeof(v));
	cudaMemcpyToSymbol(g_qt_violation, v, sizeof(v));
}
#endif

}  // namespace qt

#ifdef QT_BOUNDS
extern "C" void qt_tile_debug_report() { qt::tile_debug_report(); }
#endif
