// pack.cu -- K0: FASTA text -> 2-bit packed genome + N mask, on the device.
//
// Replaces the byte-at-a-time FastaChunkwiseParser (fasta.cpp:42-127: getc_unlocked, dna_upper
// look-up, one window at a time) -- 22 s of the reference's 63 s at 3.1 Gbp.
// The host only locates the header lines; every sequence byte is classified, counted and
// packed by the GPU:
//
//   k_find_gt      positions of every '>' in the text (rare) -> host resolves header lines
//   k_count_bases  per 4 KiB unit of sequence text: number of non-whitespace bytes
//   (scan)         exclusive prefix over units -> base index of every unit
//   k_pack_units   one warp per unit: 16 text bytes per lane per step, warp-scan compaction of
//                  the non-whitespace bytes through shared memory, then word-parallel assembly
//                  of 2-bit words and N-mask words, merged into HBM with RED.OR (unit boundaries
//                  fall inside words)
//   k_fill_n       N padding in front of / between / after records
//
// Same classification as the reference: A/C/G/T in either case are bases, every other
// non-whitespace byte is 'N' (dna_upper, fasta.cpp:18-35); whitespace is isspace() in the C
// locale (fasta.cpp:113).
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace qt {

static constexpr uint32_t UNIT_BYTES = 4096;

// class of a text byte: 0..3 = A C G T, 4 = N, 5 = whitespace
__device__ __forceinline__ uint32_t classify(uint32_t c) {
	const uint32_t u = c & 0xdfu;  // fold case
	if (u == 'A') return 0u;
	if (u == 'C') return 1u;
	if (u == 'G') return 2u;
	if (u == 'T') return 3u;
	if (c == ' ' || (c >= 9u && c <= 13u)) return 5u;
	return 4u;
}

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

__global__ void __launch_bounds__(256) k_count_bases(const uint8_t *__restrict__ text, SegTable st, uint64_t nunits,
                                                      uint32_t *counts) {
	const uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (u >= nunits) return;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t s = unit_segment(st, u);
	const uint64_t t0 = st.seg_start[s] + (u - st.unit_base[s]) * UNIT_BYTES;
	const uint64_t t1 = t0 + UNIT_BYTES < st.seg_end[s] ? t0 + UNIT_BYTES : st.seg_end[s];
	uint32_t cnt = 0;
	// aligned 16-byte loads from the 16-byte boundary at or below t0; bytes outside [t0, t1) are ignored
	// (the text buffer is padded, so the last load may run past n)
	for (uint64_t i = (t0 & ~(uint64_t)15) + lane * 16; i < t1; i += 32 * 16) {
		const uint4 v = *(const uint4 *)(text + i);
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int k = 0; k < 4; k++)
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const uint64_t p = i + 4 * k + b;
				cnt += (p >= t0 && p < t1 && classify((w[k] >> (8 * b)) & 0xffu) != 5u) ? 1u : 0u;
			}
	}
	for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if (lane == 0) counts[u] = cnt;
}

// One warp per unit.  base_of_unit[u] = bases of the whole text before unit u (exclusive scan of
// counts); seg_gbase[s] = genome position of the segment's first base minus base_of_unit of its
// first unit, so that genome position = seg_gbase[s] + base_of_unit[u] + (bases before, in unit).
__global__ void __launch_bounds__(256) k_pack_units(const uint8_t *__restrict__ text, SegTable st, uint64_t nunits,
                                                     const uint64_t *__restrict__ base_of_unit,
                                                     const int64_t *__restrict__ seg_gbase, uint32_t *bases,
                                                     uint32_t *nmask) {
	__shared__ uint8_t s_codes[8][512 + 16];
	const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint64_t u = (uint64_t)blockIdx.x * 8 + wib;
	if (u >= nunits) return;
	const uint32_t s = unit_segment(st, u);
	const uint64_t t0 = st.seg_start[s] + (u - st.unit_base[s]) * UNIT_BYTES;
	const uint64_t t1 = t0 + UNIT_BYTES < st.seg_end[s] ? t0 + UNIT_BYTES : st.seg_end[s];
	uint64_t g = (uint64_t)(seg_gbase[s] + (int64_t)base_of_unit[u]);  // genome position of the unit's first base
	uint8_t *codes = s_codes[wib];
	for (uint64_t tile = t0 & ~(uint64_t)15; tile < t1; tile += 512) {
		// classify 16 bytes per lane (one aligned 16-byte load; bytes outside [t0, t1) count as whitespace)
		const uint64_t i = tile + lane * 16;
		uint32_t cls[16];
		uint32_t cnt = 0;
		uint4 v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
		if (i < t1) v = *(const uint4 *)(text + i);
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const uint64_t p = i + k;
			cls[k] = (p >= t0 && p < t1) ? classify((w[k >> 2] >> (8 * (k & 3))) & 0xffu) : 5u;
			cnt += cls[k] != 5u;
		}
		// exclusive warp scan of the counts
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o) incl += y;
		}
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		uint32_t off = incl - cnt;
#pragma unroll
		for (int k = 0; k < 16; k++)
			if (cls[k] != 5u) codes[off++] = (uint8_t)cls[k];
		__syncwarp();
		// word-parallel assembly: lane l builds 2-bit word (g>>4)+l (+32 for a possible 33rd word)
		const uint64_t W0 = g >> 4;
		const uint32_t nwords = (uint32_t)(((g + total + 15) >> 4) - W0);
		for (uint32_t wl = lane; wl < nwords; wl += 32) {
			const int64_t first = (int64_t)((W0 + wl) << 4) - (int64_t)g;  // index in codes[] of the word's base 0
			uint32_t v = 0;
#pragma unroll
			for (int t = 0; t < 16; t++) {
				const int64_t idx = first + t;
				if (idx >= 0 && idx < (int64_t)total) {
					const uint32_t c = codes[idx];
					if (c < 4u) v |= c << (2 * t);
				}
			}
			if (v) atomicOr(&bases[W0 + wl], v);
		}
		const uint64_t M0 = g >> 5;
		const uint32_t nmw = (uint32_t)(((g + total + 31) >> 5) - M0);
		for (uint32_t wl = lane; wl < nmw; wl += 32) {
			const int64_t first = (int64_t)((M0 + wl) << 5) - (int64_t)g;
			uint32_t v = 0;
#pragma unroll
			for (int t = 0; t < 32; t++) {
				const int64_t idx = first + t;
				if (idx >= 0 && idx < (int64_t)total && codes[idx] == 4u) v |= 1u << t;
			}
			if (v) atomicOr(&nmask[M0 + wl], v);
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
