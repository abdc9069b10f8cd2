// pack.cu -- K0: FASTA text -> 2-bit packed genome + N mask, on the device.
//
// Replaces the byte-at-a-time FastaChunkwiseParser (fasta.cpp:42-127: getc_unlocked, dna_upper
// look-up, one window at a time) -- 22 s of the reference's 63 s at 3.1 Gbp.
// The host only locates the header lines; every sequence byte is classified, counted and
// packed by the GPU:
//
//   k_find_gt      positions of every '>' in the text (rare) -> host resolves header lines
//   k_count_bases  per 4 KiB unit of sequence text: number of non-whitespace bytes (word-parallel
//                  classification, pack_swar.h; all loads of a unit in flight together)
//   (scan)         exclusive prefix over units -> base index of every unit
//   k_pack_units   one warp per unit: 16 text bytes per lane per round classified four bytes per
//                  instruction (pack_swar.h), whitespace squeezed out of the lane's bit strings in
//                  registers, warp scan -> bit offsets, lanes OR their bits into a shared-memory
//                  image of the output words; interior words are plain coalesced stores, the two
//                  words a round shares with its neighbours are merged into HBM with RED.OR
//   k_fill_n       N padding in front of / between / after records
//
// Same classification as the reference: A/C/G/T in either case are bases, every other
// non-whitespace byte is 'N' (dna_upper, fasta.cpp:18-35); whitespace is isspace() in the C
// locale (fasta.cpp:113).
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"
#include "pack_swar.h"

namespace qt {

static constexpr uint32_t UNIT_BYTES = 4096;

__global__ void __launch_bounds__(256) k_find_gt(const uint8_t *__restrict__ text, uint64_t n, uint64_t *list,
                                                  uint32_t cap, uint32_t *count) {
	const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
	if (i0 >= n) return;
	uint32_t w[4] = {0, 0, 0, 0};
	if (i0 + 16 <= n) {
		const uint4 v = *(const uint4 *)(text + i0);
		w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
	} else {
		for (uint64_t k = i0; k < n; k++) w[(k - i0) >> 2] |= (uint32_t)text[k] << (8 * ((k - i0) & 3));
	}
#pragma unroll
	for (int k = 0; k < 4; k++) {
		// any byte == '>' (0x3e)?
		const uint32_t x = w[k] ^ 0x3e3e3e3eu;
		if (((x - 0x01010101u) & ~x & 0x80808080u) == 0u) continue;
		for (int b = 0; b < 4; b++) {
			if (((w[k] >> (8 * b)) & 0xffu) == 0x3eu) {
				const uint32_t slot = atomicAdd(count, 1u);
				if (slot < cap) list[slot] = i0 + 4 * k + b;
			}
		}
	}
}

// unit u covers text [seg_start[s] + 4096*p, min(.., seg_end[s])) where s = last segment with
// unit_base[s] <= u and p = u - unit_base[s]
struct SegTable {
	const uint64_t *seg_start, *seg_end;  // text ranges of the records' sequence lines
	const uint64_t *unit_base;            // [nseg+1] units before each segment
	uint32_t nseg;
};

__device__ __forceinline__ uint32_t unit_segment(const SegTable &st, uint64_t u) {
	uint32_t lo = 0, hi = st.nseg;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (st.unit_base[mid] <= u) lo = mid;
		else hi = mid;
	}
	return lo;
}

// 16 text bytes of a unit for one lane: the aligned vector at i, or blanks when it starts at or past t1
__device__ __forceinline__ uint4 load_text16(const uint8_t *__restrict__ text, uint64_t i, uint64_t t1) {
	if (i < t1) return *(const uint4 *)(text + i);
	return make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
}
// which bytes of the vector at i belong to [t0, t1): [a, b), b > a for every vector that starts below t1
__device__ __forceinline__ void vec_range(uint64_t i, uint64_t t0, uint64_t t1, uint32_t &a, uint32_t &b) {
	a = t0 > i ? (uint32_t)(t0 - i) : 0u;
	b = t1 - i < 16u ? (uint32_t)(t1 - i) : 16u;
}

static constexpr int UNIT_ROUNDS = UNIT_BYTES / 512 + 1;  // 512-byte warp rounds per unit (+1: t0 is not 16-aligned)

__global__ void __launch_bounds__(256) k_count_bases(const uint8_t *__restrict__ text, SegTable st, uint64_t nunits,
                                                      uint32_t *counts) {
	const uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (u >= nunits) return;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t s = unit_segment(st, u);
	const uint64_t t0 = st.seg_start[s] + (u - st.unit_base[s]) * UNIT_BYTES;
	const uint64_t t1 = t0 + UNIT_BYTES < st.seg_end[s] ? t0 + UNIT_BYTES : st.seg_end[s];
	// aligned 16-byte loads from the 16-byte boundary at or below t0 (the text buffer is padded, so the last load
	// may run past n); all of the unit's loads are issued before the first is looked at
	const uint64_t i0 = (t0 & ~(uint64_t)15) + lane * 16;
	uint4 v[UNIT_ROUNDS];
#pragma unroll
	for (int r = 0; r < UNIT_ROUNDS; r++) v[r] = load_text16(text, i0 + 512u * r, t1);
	uint32_t cnt = 0;
#pragma unroll
	for (int r = 0; r < UNIT_ROUNDS; r++) {
		const uint64_t i = i0 + 512u * r;
		if (i >= t1) continue;
		uint32_t a, b;
		vec_range(i, t0, t1, a, b);
		const uint32_t ws = (ws_vec(v[r].x, v[r].y, v[r].z, v[r].w) >> a) & sw_low(b - a);
		cnt += (b - a) - sw_popc(ws);
	}
	for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if (lane == 0) counts[u] = cnt;
}

// One warp per unit.  base_of_unit[u] = bases of the whole text before unit u (exclusive scan of
// counts); seg_gbase[s] = genome position of the segment's first base minus base_of_unit of its
// first unit, so that genome position = seg_gbase[s] + base_of_unit[u] + (bases before, in unit).
//
// Per round of 512 text bytes: every lane classifies its 16 bytes word-parallel (pack_swar.h) and squeezes the
// whitespace out of its 32 bits of codes / 16 N bits; an exclusive warp scan of the lanes' base counts gives each lane
// its bit offset in the round's output, the lanes OR their bits into a small shared-memory image of the output
// words (ATOMS.OR: a lane's bits straddle at most two words), and lane l ships image word l.  Words that lie
// wholly inside the round are plain stores; the first and the last one, shared with the neighbouring round or
// unit, are merged into HBM with RED.OR.
static constexpr int PK_BW = 36;  // image words per warp: 2-bit codes, (15 + 512 + 15) / 16 + spill
static constexpr int PK_NW = 20;  // N-mask words, (31 + 512 + 31) / 32 + spill

__global__ void __launch_bounds__(256) k_pack_units(const uint8_t *__restrict__ text, SegTable st, uint64_t nunits,
                                                     const uint64_t *__restrict__ base_of_unit,
                                                     const int64_t *__restrict__ seg_gbase, uint32_t *bases,
                                                     uint32_t *nmask) {
	__shared__ uint32_t s_img[8][PK_BW + PK_NW];
	const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint64_t u = (uint64_t)blockIdx.x * 8 + wib;
	if (u >= nunits) return;
	const uint32_t s = unit_segment(st, u);
	const uint64_t t0 = st.seg_start[s] + (u - st.unit_base[s]) * UNIT_BYTES;
	const uint64_t t1 = t0 + UNIT_BYTES < st.seg_end[s] ? t0 + UNIT_BYTES : st.seg_end[s];
	uint64_t g = (uint64_t)(seg_gbase[s] + (int64_t)base_of_unit[u]);  // genome position of the unit's first base
	uint32_t *img_b = s_img[wib], *img_n = img_b + PK_BW;
	for (uint32_t k = lane; k < (uint32_t)(PK_BW + PK_NW); k += 32) img_b[k] = 0;
	__syncwarp();
	const uint64_t i0 = (t0 & ~(uint64_t)15) + lane * 16;
	uint4 nxt = load_text16(text, i0, t1);
	for (uint64_t tile = t0 & ~(uint64_t)15; tile < t1; tile += 512) {
		const uint64_t i = tile + lane * 16;
		const uint4 v = nxt;
		nxt = load_text16(text, i + 512, t1);  // next round's bytes fly while this round is packed
		uint32_t cnt = 0;
		VecClass vc{0u, 0u, 0u};
		if (i < t1) {
			uint32_t a, b;
			vec_range(i, t0, t1, a, b);
			vc = classify_vec(v.x, v.y, v.z, v.w);
			cnt = squeeze_vec(vc, a, b);
		}
		// exclusive warp scan of the counts
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o) incl += y;
		}
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		const uint32_t excl = incl - cnt;
		const uint32_t gph = (uint32_t)(g & 15u), gpn = (uint32_t)(g & 31u);
		if (vc.codes) {
			const uint32_t ob = 2u * (gph + excl);
			const uint64_t x = (uint64_t)vc.codes << (ob & 31u);
			if ((uint32_t)x) atomicOr(&img_b[ob >> 5], (uint32_t)x);
			if ((uint32_t)(x >> 32)) atomicOr(&img_b[(ob >> 5) + 1], (uint32_t)(x >> 32));
		}
		if (vc.nbits) {
			const uint32_t on = gpn + excl;
			const uint64_t x = (uint64_t)vc.nbits << (on & 31u);
			if ((uint32_t)x) atomicOr(&img_n[on >> 5], (uint32_t)x);
			if ((uint32_t)(x >> 32)) atomicOr(&img_n[(on >> 5) + 1], (uint32_t)(x >> 32));
		}
		__syncwarp();
		// lane l ships image words l and l + 32 and clears them for the next round
		const uint64_t W0 = g >> 4;
		const uint32_t nwords = (gph + total + 15u) >> 4;
		for (uint32_t wl = lane; wl < nwords; wl += 32) {
			const uint32_t x = img_b[wl];
			img_b[wl] = 0;
			if (wl != 0 && wl + 1 != nwords) bases[W0 + wl] = x;
			else if (x) atomicOr(&bases[W0 + wl], x);
		}
		const uint64_t M0 = g >> 5;
		const uint32_t nmw = (gpn + total + 31u) >> 5;
		for (uint32_t wl = lane; wl < nmw; wl += 32) {
			const uint32_t x = img_n[wl];
			if (x) {
				img_n[wl] = 0;
				atomicOr(&nmask[M0 + wl], x);
			}
		}
		__syncwarp();
		g += total;
	}
}

// sets N bits for [lo[i], hi[i]) (record padding)
__global__ void k_fill_n(const uint64_t *lo, const uint64_t *hi, uint32_t n, uint32_t *nmask) {
	const uint32_t i = blockIdx.x;
	if (i >= n) return;
	for (uint64_t g = lo[i] + threadIdx.x; g < hi[i]; g += blockDim.x) atomicOr(&nmask[g >> 5], 1u << (g & 31));
}

void launch_find_gt(const uint8_t *text, uint64_t n, uint64_t *list, uint32_t cap, uint32_t *count, cudaStream_t st) {
	if (!n) return;
	const uint64_t threads = (n + 15) / 16;
	k_find_gt<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(text, n, list, cap, count);
	kernels_count_launch();
}

void launch_count_bases(const uint8_t *text, const uint64_t *seg_start, const uint64_t *seg_end,
                        const uint64_t *unit_base, uint32_t nseg, uint64_t nunits, uint32_t *counts, cudaStream_t st) {
	if (!nunits) return;
	SegTable stb{seg_start, seg_end, unit_base, nseg};
	k_count_bases<<<(unsigned)((nunits + 7) / 8), 256, 0, st>>>(text, stb, nunits, counts);
	kernels_count_launch();
}

void launch_pack_units(const uint8_t *text, const uint64_t *seg_start, const uint64_t *seg_end,
                       const uint64_t *unit_base, uint32_t nseg, uint64_t nunits, const uint64_t *base_of_unit,
                       const int64_t *seg_gbase, uint32_t *bases, uint32_t *nmask, cudaStream_t st) {
	if (!nunits) return;
	SegTable stb{seg_start, seg_end, unit_base, nseg};
	k_pack_units<<<(unsigned)((nunits + 7) / 8), 256, 0, st>>>(text, stb, nunits, base_of_unit, seg_gbase, bases, nmask);
	kernels_count_launch();
}

void launch_fill_n(const uint64_t *lo, const uint64_t *hi, uint32_t n, uint32_t *nmask, cudaStream_t st) {
	if (!n) return;
	k_fill_n<<<n, 128, 0, st>>>(lo, hi, n, nmask);
	kernels_count_launch();
}

}  // namespace qt
