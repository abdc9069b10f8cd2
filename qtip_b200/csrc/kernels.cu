// kernels.cu -- sm_100a kernels of libqtipsim (generic path, one warp per read).
//
//   k_window_ncount : N count per FASTA window            (simplesim.cpp:228-234)
//   k_read_lengths  : FASTQ record length per read        (simplesim.cpp:102-171 byte counts)
//   k_scan_*        : exclusive scan of lengths -> byte offsets in the output streams
//   k_plan_philox_lengths : native-RNG plan (template, window, offset per read) + record lengths
//   k_simulate      : SimulatedRead::init/mutate/init_random + write/write_pair
//                     (simplesim.h:68-117, simplesim.cpp:31-71, :102-171)
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"

namespace qt {

static int g_launches = 0;
int kernels_launch_count() { return g_launches; }
void kernels_count_launch() { g_launches++; }

__global__ void __launch_bounds__(256) k_ctl_copy(unsigned char *dst, const unsigned char *src, size_t n, int vec) {
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	size_t done = 0;
	if (vec) {  // both 16-byte aligned
		const size_t n16 = n / 16;
		const uint4 *s = (const uint4 *)src;
		uint4 *d = (uint4 *)dst;
		for (size_t i = tid; i < n16; i += nth) d[i] = s[i];
		done = n16 * 16;
	}
	for (size_t i = done + tid; i < n; i += nth) dst[i] = src[i];
}

// Interleaves the two FASTQ streams of a pair category record by record (mate 1, mate 2, mate 1, ...): pair i's records are
// src0[off0[i], end0) and src1[off1[i], end1) and go to dst[off0[i] + off1[i] ...).  One warp per pair, bytes.
__global__ void __launch_bounds__(256) k_interleave(const char *src0, const char *src1, const uint64_t *off0, const uint64_t *off1, uint64_t n,
                                                  uint64_t tot0, uint64_t tot1, char *dst) {
	const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	if (i >= n) return;
	const uint64_t a0 = off0[i], a1 = off1[i];
	const uint64_t e0 = i + 1 < n ? off0[i + 1] : tot0, e1 = i + 1 < n ? off1[i + 1] : tot1;
	char *d = dst + a0 + a1;
	for (uint64_t k = lane; k < e0 - a0; k += 32) d[k] = src0[a0 + k];
	d += e0 - a0;
	for (uint64_t k = lane; k < e1 - a1; k += 32) d[k] = src1[a1 + k];
}

void launch_interleave(const char *src0, const char *src1, const uint64_t *off0, const uint64_t *off1, uint64_t n, uint64_t tot0, uint64_t tot1,
                       char *dst, cudaStream_t st) {
	if (!n) return;
	k_interleave<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(src0, src1, off0, off1, n, tot0, tot1, dst);
	kernels_count_launch();
}

void launch_ctl_copy(void *dst, const void *src, size_t n, cudaStream_t st) {
	if (!n) return;
	const int vec = (((uintptr_t)dst | (uintptr_t)src) & 15) == 0;
	const size_t work = vec ? n / 16 + 16 : n;
	unsigned blocks = (unsigned)std::min<size_t>((work + 255) / 256, 592);
	k_ctl_copy<<<blocks, 256, 0, st>>>((unsigned char *)dst, (const unsigned char *)src, n, vec);
	kernels_count_launch();
}

// ------------------------------------------------------------------ helpers --

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t fetch_code(const DevGenome &g, uint64_t gp, uint64_t gend) {
	if (gp >= gend) return 4u;  // past the record: treated as N (reference reads stale buffer)
	const uint32_t nm = __ldg(&g.nmask[gp >> 5]);
	if ((nm >> (gp & 31)) & 1u) return 4u;
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

__device__ __forceinline__ uint32_t ndigits_u64(uint64_t v) {
	if (v <= 0xffffffffull) {  // offsets fit 32 bits in practice: a compare chain, no 64-bit divisions
		const uint32_t x = (uint32_t)v;
		return x < 10u ? 1u : x < 100u ? 2u : x < 1000u ? 3u : x < 10000u ? 4u : x < 100000u ? 5u : x < 1000000u ? 6u
		     : x < 10000000u ? 7u : x < 100000000u ? 8u : x < 1000000000u ? 9u : 10u;
	}
	uint32_t d = 1;
	while (v >= 10) {
		v /= 10;
		d++;
	}
	return d;
}

// name length incl. '@' and '\n' (simplesim.cpp:104-110, :140-150)
__device__ __forceinline__ uint32_t name_len_unp(uint32_t refid_len, uint64_t refoff, const MateDesc &m,
                                                 uint32_t typ_len) {
	return 13u + refid_len + ndigits_u64(refoff) + m.score_len + typ_len;
}
__device__ __forceinline__ uint32_t name_len_pair(uint32_t refid_len, uint64_t off1, uint32_t sl1,
                                                  uint64_t off2, uint32_t sl2, uint32_t typ_len) {
	return 18u + 2u * refid_len + ndigits_u64(off1) + ndigits_u64(off2) + sl1 + sl2 + typ_len;
}

__device__ __forceinline__ char *put_u64(char *p, uint64_t v) {
	char tmp[20];
	int n = 0;
	do {
		tmp[n++] = (char)('0' + (int)(v % 10));
		v /= 10;
	} while (v);
	while (n) *p++ = tmp[--n];
	return p;
}
__device__ __forceinline__ char *put_bytes(char *p, const char *s, uint32_t n) {
	for (uint32_t i = 0; i < n; i++) p[i] = s[i];
	return p + n;
}

// "@qtip!:<refid>:<+|->:<refoff>:<score>" ... serial; one lane.
__device__ char *put_origin(char *p, const char *refid, uint32_t refid_len, bool fw, uint64_t refoff,
                            const char *score, uint32_t score_len) {
	p = put_bytes(p, refid, refid_len);
	*p++ = ':';
	*p++ = fw ? '+' : '-';
	*p++ = ':';
	p = put_u64(p, refoff);
	*p++ = ':';
	p = put_bytes(p, score, score_len);
	return p;
}

// ------------------------------------------------------------- window ncount --

__global__ void __launch_bounds__(256) k_window_ncount(DevGenome g, const Window *win, uint32_t nwin,
                                                        uint32_t *ncount) {
	const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (w >= nwin) return;
	const Window wn = win[w];
	const uint64_t b0 = g.recs[wn.record].gstart + wn.refoff;
	const uint64_t b1 = b0 + wn.retsz;  // exclusive
	const uint64_t w0 = b0 >> 5, w1 = (b1 + 31) >> 5;
	uint32_t cnt = 0;
	for (uint64_t i = w0 + lane_id(); i < w1; i += 32) {
		uint32_t v = __ldg(&g.nmask[i]);
		if (i == w0) v &= 0xffffffffu << (b0 & 31);
		if (i == w1 - 1 && (b1 & 31)) v &= 0xffffffffu >> (32 - (b1 & 31));
		cnt += __popc(v);
	}
	for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if (lane_id() == 0) ncount[w] = cnt;
}

void launch_window_ncount(const DevGenome &g, const Window *win, uint32_t nwin, uint32_t *ncount,
                          cudaStream_t st) {
	if (!nwin) return;
	k_window_ncount<<<(nwin + 7) / 8, 256, 0, st>>>(g, win, nwin, ncount);
	g_launches++;
}

// -------------------------------------------------------------- read lengths --

// Geometry of one planned read: where each output record goes and how long it is.
struct ReadGeom {
	uint32_t name_len;
	uint32_t body0, body1;  // read lengths written to stream 0 / 1
	uint64_t refoff[2];     // leftmost ref base of mate 0 / 1 (record coordinates)
};

__device__ __forceinline__ ReadGeom read_geom(const SimArgs &a, const PlanRec &pr, const TmplDesc &t) {
	ReadGeom gm;
	const uint32_t rl = a.g.recs[pr.record].name_len;
	gm.refoff[0] = pr.refoff + t.delta[0];
	gm.refoff[1] = pr.refoff + t.delta[1];
	if (a.cat == CAT_U) {
		gm.name_len = name_len_unp(rl, gm.refoff[0], t.m[0], t.typ_len);
		gm.body0 = t.m[0].len;
		gm.body1 = 0;
	} else if (a.cat == CAT_B) {
		// both name halves describe the aligned mate (simplesim.cpp:299-330)
		gm.name_len = name_len_pair(rl, gm.refoff[0], t.m[0].score_len, gm.refoff[0], t.m[0].score_len, t.typ_len);
		gm.body0 = t.aligned_is_mate1 ? t.m[0].len : t.opp_len;
		gm.body1 = t.aligned_is_mate1 ? t.opp_len : t.m[0].len;
	} else {
		gm.name_len = name_len_pair(rl, gm.refoff[0], t.m[0].score_len, gm.refoff[1], t.m[1].score_len, t.typ_len);
		gm.body0 = t.m[0].len;
		gm.body1 = t.m[1].len;
	}
	return gm;
}

__global__ void __launch_bounds__(256) k_read_lengths(SimArgs a, uint32_t *len0, uint32_t *len1,
                                                       unsigned long long *gbytes) {
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t gb = 0;
	if (r < a.nreads) {
		const PlanRec pr = a.plan[r];
		const TmplDesc &t = a.c.t[pr.tmpl];
		const ReadGeom gm = read_geom(a, pr, t);
		len0[r] = gm.name_len + 2u * gm.body0 + 4u;
		if (len1) len1[r] = gm.name_len + 2u * gm.body1 + 4u;
		gb = (t.m[0].rflen + 3u) / 4u;
		if (t.nmates == 2) gb += (t.m[1].rflen + 3u) / 4u;
	}
	for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
	__shared__ uint32_t ws[8];
	if (lane_id() == 0) ws[threadIdx.x >> 5] = gb;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long s = 0;
		for (int i = 0; i < 8; i++) s += ws[i];
		if (s) atomicAdd(gbytes, s);
	}
}

void launch_read_lengths(const SimArgs &a, uint32_t *len0, uint32_t *len1, unsigned long long *gbytes,
                         cudaStream_t st) {
	if (!a.nreads) return;
	k_read_lengths<<<(unsigned)((a.nreads + 255) / 256), 256, 0, st>>>(a, len0, len1, gbytes);
	g_launches++;
}

// ---------------------------------------------------------------------- scan --

static constexpr int SCAN_T = 256;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total) {
	__shared__ uint64_t wsum[SCAN_T / 32];
	const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
	uint64_t x = v;
	for (int o = 1; o < 32; o <<= 1) {
		uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= (uint32_t)o) x += y;
	}
	if (lane == 31) wsum[wid] = x;
	__syncthreads();
	if (wid == 0) {
		uint64_t s = lane < SCAN_T / 32 ? wsum[lane] : 0;
		for (int o = 1; o < SCAN_T / 32; o <<= 1) {
			uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
			if (lane >= (uint32_t)o) s += y;
		}
		if (lane < SCAN_T / 32) wsum[lane] = s;
	}
	__syncthreads();
	const uint64_t base = wid ? wsum[wid - 1] : 0;
	*total = wsum[SCAN_T / 32 - 1];
	__syncthreads();
	return base + x - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tile_sums(const uint32_t *len, uint64_t n, uint64_t *tsum) {
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint64_t s = 0;
	for (int i = 0; i < SCAN_ITEMS; i++)
		if (base + i < n) s += len[base + i];
	uint64_t tot;
	block_exclusive_scan(s, &tot);
	if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tile_prefix(uint64_t *tsum, uint64_t ntiles, uint64_t *total) {
	// single block: exclusive scan of tile sums in place
	uint64_t carry = 0;
	for (uint64_t b = 0; b < ntiles; b += SCAN_T) {
		const uint64_t i = b + threadIdx.x;
		const uint64_t v = i < ntiles ? tsum[i] : 0;
		uint64_t tot;
		const uint64_t ex = block_exclusive_scan(v, &tot);
		if (i < ntiles) tsum[i] = carry + ex;
		carry += tot;
	}
	if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_final(const uint32_t *len, uint64_t *off, uint64_t n,
                                                        const uint64_t *tsum) {
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	uint64_t s = 0;
	for (int i = 0; i < SCAN_ITEMS; i++) {
		v[i] = base + i < n ? len[base + i] : 0;
		s += v[i];
	}
	uint64_t tot;
	uint64_t ex = block_exclusive_scan(s, &tot) + tsum[blockIdx.x];
	for (int i = 0; i < SCAN_ITEMS; i++) {
		if (base + i < n) off[base + i] = ex;
		ex += v[i];
	}
}

void launch_scan(const uint32_t *len, uint64_t *off, uint64_t n, uint64_t *tmp, uint64_t *total,
                 cudaStream_t st) {
	if (!n) {
		cudaMemsetAsync(total, 0, sizeof(uint64_t), st);
		return;
	}
	const uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	k_scan_tile_sums<<<(unsigned)ntiles, SCAN_T, 0, st>>>(len, n, tmp);
	k_scan_tile_prefix<<<1, SCAN_T, 0, st>>>(tmp, ntiles, total);
	k_scan_final<<<(unsigned)ntiles, SCAN_T, 0, st>>>(len, off, n, tmp);
	g_launches += 3;
}

// -------------------------------------------------------------- philox plan --

// Per-window read counts in native mode: Binomial(target, p_window) -- the reference's own model
// (simplesim.cpp:224-226, 242, 279, 340-341) -- sampled with binom_sample from a Philox stream keyed
// by (seed, window, category).  One thread per window.
__global__ void __launch_bounds__(128) k_philox_counts(const Window *win, const uint32_t *ncount, const uint32_t *pclass,
                                                        const BinomSetup *setups, uint32_t nwin, uint64_t olap,
                                                        uint64_t target, uint64_t seed, int cat, uint32_t *counts) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwin) return;
	const Window wn = win[w];
	uint32_t n = 0;
	if (target && wn.retsz >= olap && !((int)ncount[w] > (int)(0.9 * (double)wn.retsz))) {
		const BinomSetup bs = setups[(size_t)pclass[w] * NCAT + cat];
		uint32_t ctr = 0, have = 0;
		Philox4 blk{};
		auto nx = [&]() -> float {
			if (!have) {
				blk = philox4x32(w, 0u, ctr++, PHILOX_DOM_WINDOW | (uint32_t)cat, (uint32_t)seed, (uint32_t)(seed >> 32));
				have = 4;
			}
			const uint32_t x = blk.v[4 - have];
			have--;
			return __fmul_rn((float)((x >> 8) + 1u), 1.0f / 16777216.0f);  // (0, 1]
		};
		n = (uint32_t)binom_sample(bs, nx);
	}
	counts[w] = n;
}

void launch_philox_counts(const Window *win, const uint32_t *ncount, const uint32_t *pclass, const BinomSetup *setups,
                          uint32_t nwin, uint64_t olap, uint64_t target, uint64_t seed, int cat, uint32_t *counts,
                          cudaStream_t st) {
	if (!nwin) return;
	k_philox_counts<<<(nwin + 127) / 128, 128, 0, st>>>(win, ncount, pclass, setups, nwin, olap, target, seed, cat, counts);
	g_launches++;
}

// Native plan, fused with k_read_lengths: the thread that plans read r (window by binary search over the
// cumulative per-window counts -- the length-weighted choice was made when those were drawn -- then template
// and offset from one Philox block) also computes its FASTQ record length(s).  cum[] holds GLOBAL cumulative
// counts for this shard's windows; a = the category's SimArgs: a.plan is written here, templates and records
// are read; a.read_index_base = cum[0].
__global__ void __launch_bounds__(256) k_plan_philox_lengths(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap,
                                                              uint32_t ntmpl, uint64_t seed, SimArgs a, PlanRec *plan,
                                                              uint32_t *len0, uint32_t *len1, unsigned long long *gbytes) {
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t gb = 0;
	// window of the warp's first read by binary search (lane 0), then every lane walks forward from there: the 32
	// consecutive reads of a warp lie in the same or the next few windows (invariant: cum[lo] <= gi < cum[lo + 1])
	uint32_t lo = 0;
	{
		const uint64_t r_first = r - lane_id();
		if (lane_id() == 0 && r_first < a.nreads) {
			const uint64_t g0 = a.read_index_base + r_first;
			uint32_t hi = nwin;
			while (hi - lo > 1) {
				const uint32_t mid = (lo + hi) >> 1;
				if (cum[mid] <= g0) lo = mid;
				else hi = mid;
			}
		}
		lo = __shfl_sync(0xffffffffu, lo, 0);
	}
	if (r < a.nreads) {
		const uint64_t gi = a.read_index_base + r;
		while (lo + 1 < nwin && cum[lo + 1] <= gi) lo++;
		const Window w = win[lo];
		const Philox4 x = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), 0u, PHILOX_DOM_PLAN | (uint32_t)a.cat,
		                             (uint32_t)seed, (uint32_t)(seed >> 32));
		const uint64_t nslots = (uint64_t)w.retsz - olap;
		PlanRec pr;
		pr.tmpl = __umulhi(x.v[0], ntmpl);
		pr.record = w.record;
		pr.refoff = w.refoff + (nslots ? (uint64_t)__umulhi(x.v[1], (uint32_t)nslots) : 0);
		pr.s1 = 0;
		pr.s2 = 0;
		plan[r] = pr;
		const TmplDesc &t = a.c.t[pr.tmpl];
		const ReadGeom gm = read_geom(a, pr, t);
		len0[r] = gm.name_len + 2u * gm.body0 + 4u;
		if (len1) len1[r] = gm.name_len + 2u * gm.body1 + 4u;
		gb = (t.m[0].rflen + 3u) / 4u;
		if (t.nmates == 2) gb += (t.m[1].rflen + 3u) / 4u;
	}
	for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
	__shared__ uint32_t ws[8];
	if (lane_id() == 0) ws[threadIdx.x >> 5] = gb;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long s = 0;
		for (int i = 0; i < 8; i++) s += ws[i];
		if (s) atomicAdd(gbytes, s);
	}
}

// The same plan, with the record lengths turned into stream offsets in the same pass: a single-pass prefix sum with
// decoupled look-back over tiles of PS_THREADS reads (tile ids from a ticket counter, so a tile only ever waits for
// tiles that have started).  status[s][tile] = flag << 62 | value: flag 1 = the tile's own sum, 2 = the sum of
// everything up to and including the tile.  Replaces k_plan_philox_lengths + three scan kernels per stream and the
// length arrays between them (28 bytes written and 8 read per read instead of 36 and 12).
static constexpr int PS_THREADS = 256, PS_ITEMS = 4, PS_TILE = PS_THREADS * PS_ITEMS;  // consecutive reads per thread
template <int NS>
__global__ void __launch_bounds__(PS_THREADS) k_plan_philox_scan(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap,
                                                                 uint32_t ntmpl, uint64_t seed, SimArgs a, PlanRec *plan, uint64_t *off0,
                                                                 uint64_t *off1, uint64_t *tot0, uint64_t *tot1, unsigned long long *gbytes,
                                                                 unsigned long long *status, unsigned int *ticket, uint32_t ntiles) {
	__shared__ uint32_t s_tile, s_lo;
	__shared__ uint32_t s_wsum[NS][PS_THREADS / 32];
	__shared__ uint64_t s_prefix[NS];
	__shared__ uint32_t s_gb[PS_THREADS / 32];
	const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	// window of the tile's first read: a 32-ary search by the first warp (three dependent probes for 32 k windows
	// where a binary search takes fifteen), then every thread walks forward from there -- a tile's reads lie in the
	// same or the next few windows (invariant: cum[lo] <= gi < cum[lo + 1])
	if (wid == 0) {
		const uint64_t g0 = a.read_index_base + (uint64_t)tile * PS_TILE;
		uint32_t l = 0, h = nwin;
		while (h - l > 1) {
			const uint32_t step = (h - l + 31u) / 32u;
			const uint32_t idx = l + (lane + 1u) * step;
			const bool le = idx < h && cum[idx] <= g0;
			const uint32_t n = (uint32_t)__popc(__ballot_sync(0xffffffffu, le));
			const uint32_t nh = l + (n + 1u) * step;
			l += n * step;
			h = nh < h ? nh : h;
		}
		if (lane == 0) s_lo = l;
	}
	__syncthreads();
	uint32_t lo = s_lo;
	const uint64_t rbase = (uint64_t)tile * PS_TILE + (uint64_t)tid * PS_ITEMS;
	uint32_t gb = 0, len[PS_ITEMS][NS];
	// the plan records first: what follows gathers from the templates, and those loads then fly together
	PlanRec pr[PS_ITEMS];
#pragma unroll
	for (int k = 0; k < PS_ITEMS; k++) {
		const uint64_t r = rbase + k;
		pr[k] = PlanRec{};
		if (r < a.nreads) {
			const uint64_t gi = a.read_index_base + r;
			while (lo + 1 < nwin && cum[lo + 1] <= gi) lo++;
			const Window w = win[lo];
			const Philox4 x = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), 0u, PHILOX_DOM_PLAN | (uint32_t)a.cat, (uint32_t)seed,
			                             (uint32_t)(seed >> 32));
			const uint64_t nslots = (uint64_t)w.retsz - olap;
			pr[k].tmpl = __umulhi(x.v[0], ntmpl);
			pr[k].record = w.record;
			pr[k].refoff = w.refoff + (nslots ? (uint64_t)__umulhi(x.v[1], (uint32_t)nslots) : 0);
		}
	}
	if (rbase + PS_ITEMS <= a.nreads) {  // a thread's four records are 96 contiguous, 16-byte aligned bytes: six vector stores
		static_assert(sizeof(PlanRec) == 24 && PS_ITEMS == 4, "plan store");
		uint4 *pd = (uint4 *)(plan + rbase);
		const uint32_t o0l = (uint32_t)pr[0].refoff, o0h = (uint32_t)(pr[0].refoff >> 32), o1l = (uint32_t)pr[1].refoff, o1h = (uint32_t)(pr[1].refoff >> 32);
		const uint32_t o2l = (uint32_t)pr[2].refoff, o2h = (uint32_t)(pr[2].refoff >> 32), o3l = (uint32_t)pr[3].refoff, o3h = (uint32_t)(pr[3].refoff >> 32);
		pd[0] = make_uint4(pr[0].tmpl, pr[0].record, o0l, o0h);
		pd[1] = make_uint4(0u, 0u, pr[1].tmpl, pr[1].record);
		pd[2] = make_uint4(o1l, o1h, 0u, 0u);
		pd[3] = make_uint4(pr[2].tmpl, pr[2].record, o2l, o2h);
		pd[4] = make_uint4(0u, 0u, pr[3].tmpl, pr[3].record);
		pd[5] = make_uint4(o3l, o3h, 0u, 0u);
	} else {
#pragma unroll
		for (int k = 0; k < PS_ITEMS; k++)
			if (rbase + k < a.nreads) plan[rbase + k] = pr[k];
	}
#pragma unroll
	for (int k = 0; k < PS_ITEMS; k++) {
#pragma unroll
		for (int s = 0; s < NS; s++) len[k][s] = 0u;
		if (rbase + k < a.nreads) {
			// what the lengths need of the template, as 16-byte blocks (common.h): block 1 of each mate (len, rflen, ...,
			// score_len) and, for pairs, the block behind the mates (delta, opp_len, typ_len, aligned_is_mate1)
			const uint4 *tp = (const uint4 *)(a.c.t + pr[k].tmpl);
			const uint32_t rl = __ldg(&a.g.recs[pr[k].record].name_len);
			const uint4 m0 = __ldg(tp + 1);
			const uint32_t len_0 = m0.x, sl0 = (m0.w >> 16) & 0xffu;
			gb += (m0.y + 3u) / 4u;
			if (NS == 1) {
				len[k][0] = 13u + rl + ndigits_u64(pr[k].refoff) + sl0 + 1u + 2u * len_0 + 4u;  // name_len_unp, typ "u"
			} else {
				const uint4 tl = __ldg(tp + 8);
				const uint32_t typ_len = (tl.w >> 16) & 0xffu;
				if (a.cat == CAT_B) {
					const bool m1 = (tl.w >> 24) != 0;
					const uint32_t nl = 18u + 2u * rl + 2u * ndigits_u64(pr[k].refoff) + 2u * sl0 + typ_len;
					len[k][0] = nl + 2u * (m1 ? len_0 : tl.z) + 4u;
					len[k][NS - 1] = nl + 2u * (m1 ? tl.z : len_0) + 4u;
				} else {
					const uint4 m1 = __ldg(tp + 5);
					const uint32_t nl = 18u + 2u * rl + ndigits_u64(pr[k].refoff + tl.x) + ndigits_u64(pr[k].refoff + tl.y) + sl0 + ((m1.w >> 16) & 0xffu) + typ_len;
					len[k][0] = nl + 2u * len_0 + 4u;
					len[k][NS - 1] = nl + 2u * m1.x + 4u;
					gb += (m1.y + 3u) / 4u;
				}
			}
		}
	}
	// exclusive scan inside the tile (32-bit: a tile's records stay far below 4 GiB)
	uint32_t tsum[NS], incl[NS];
#pragma unroll
	for (int s = 0; s < NS; s++) {
		uint32_t v = 0;
#pragma unroll
		for (int k = 0; k < PS_ITEMS; k++) v += len[k][s];
		tsum[s] = v;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
			if ((int)lane >= o) v += u;
		}
		incl[s] = v;
		if (lane == 31) s_wsum[s][wid] = v;
	}
	for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
	if (lane == 0) s_gb[wid] = gb;
	__syncthreads();
	if (wid == 0) {
		uint32_t g = lane < PS_THREADS / 32 ? s_gb[lane] : 0u;
		for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
		if (lane == 0 && g) atomicAdd(gbytes, (unsigned long long)g);
#pragma unroll
		for (int s = 0; s < NS; s++) {
			// warp sums -> exclusive warp offsets, tile sum
			uint32_t v = lane < PS_THREADS / 32 ? s_wsum[s][lane] : 0u;
			const uint32_t mine = v;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
				if ((int)lane >= o) v += u;
			}
			const uint64_t agg = __shfl_sync(0xffffffffu, v, 31);
			if (lane < PS_THREADS / 32) s_wsum[s][lane] = v - mine;
			unsigned long long *st = status + (size_t)s * ntiles;
			uint64_t excl = 0;
			if (tile) {
				if (lane == 0) *(volatile unsigned long long *)&st[tile] = (1ull << 62) | agg;
				int pos = (int)tile - 1;
				for (;;) {
					const int idx = pos - (int)lane;
					unsigned long long sv = 2ull << 62;  // in front of tile 0: an inclusive sum of zero
					if (idx >= 0) {
						do sv = *(volatile unsigned long long *)&st[idx];
						while ((sv >> 62) == 0);
					}
					const uint32_t done = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
					const uint32_t upto = done ? (uint32_t)__ffs(done) - 1u : 31u;  // nearest tile whose inclusive sum is known
					uint64_t part = lane <= upto ? (sv & ((1ull << 62) - 1)) : 0ull;
					for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
					excl += part;
					if (done) break;
					pos -= 32;
				}
			}
			if (lane == 0) {
				*(volatile unsigned long long *)&st[tile] = (2ull << 62) | (excl + agg);
				s_prefix[s] = excl;
				if (tile == ntiles - 1) *(s ? tot1 : tot0) = excl + agg;
			}
		}
	}
	__syncthreads();
#pragma unroll
	for (int s = 0; s < NS; s++) {
		uint64_t o = s_prefix[s] + s_wsum[s][wid] + (incl[s] - tsum[s]);
		uint64_t *off = s ? off1 : off0;
		if (rbase + PS_ITEMS <= a.nreads) {
			const uint64_t o1 = o + len[0][s], o2 = o1 + len[1][s], o3 = o2 + len[2][s];
			uint4 *od = (uint4 *)(off + rbase);
			od[0] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32));
			od[1] = make_uint4((uint32_t)o2, (uint32_t)(o2 >> 32), (uint32_t)o3, (uint32_t)(o3 >> 32));
		} else {
#pragma unroll
			for (int k = 0; k < PS_ITEMS; k++) {
				if (rbase + k < a.nreads) off[rbase + k] = o;
				o += len[k][s];
			}
		}
	}
}

// status: >= 2 * ntiles + 1 u64, zeroed by this call (the last word is the ticket counter)
void launch_plan_philox_scan(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap, uint32_t ntmpl, uint64_t seed,
                             const SimArgs &a, PlanRec *plan, uint64_t *off0, uint64_t *off1, uint64_t *tot0, uint64_t *tot1,
                             unsigned long long *gbytes, unsigned long long *status, cudaStream_t st) {
	if (!a.nreads) return;
	const uint32_t ntiles = (uint32_t)((a.nreads + PS_TILE - 1) / PS_TILE);
	cudaMemsetAsync(status, 0, ((size_t)2 * ntiles + 1) * 8, st);
	unsigned int *ticket = (unsigned int *)(status + (size_t)2 * ntiles);
	if (a.cat == CAT_U)
		k_plan_philox_scan<1><<<ntiles, PS_THREADS, 0, st>>>(win, cum, nwin, olap, ntmpl, seed, a, plan, off0, off1, tot0, tot1, gbytes, status, ticket, ntiles);
	else
		k_plan_philox_scan<2><<<ntiles, PS_THREADS, 0, st>>>(win, cum, nwin, olap, ntmpl, seed, a, plan, off0, off1, tot0, tot1, gbytes, status, ticket, ntiles);
	g_launches++;
}
size_t plan_philox_scan_status_bytes(uint64_t nreads) { return ((size_t)2 * ((nreads + PS_TILE - 1) / PS_TILE) + 2) * 8; }

void launch_plan_philox_lengths(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap, uint32_t ntmpl, uint64_t seed,
                                const SimArgs &a, PlanRec *plan, uint32_t *len0, uint32_t *len1, unsigned long long *gbytes,
                                cudaStream_t st) {
	if (!a.nreads) return;
	k_plan_philox_lengths<<<(unsigned)((a.nreads + 255) / 256), 256, 0, st>>>(win, cum, nwin, olap, ntmpl, seed, a, plan, len0, len1,
	                                                                        gbytes);
	g_launches++;
}


// ------------------------------------------------------------------ simulate --

static constexpr int SIM_WARPS = 8;
static constexpr int NAME_SMEM = 320;  // per-warp name staging; longer names go straight to global

__device__ __forceinline__ char code_char(uint32_t c) { return "ACGTN"[c]; }
__device__ __forceinline__ char code_char_comp(uint32_t c) { return "TGCAN"[c]; }  // asc2dnacomp, simplesim.cpp:76-97

// Where this read's base draws come from.
struct DrawSrc {
	// REPLAY: draw k (0-based, counted from the read's first base draw) is the generator's
	// output k+1 steps after `s` (the state left by the offset draw).
	LState s;
	const uint32_t *jt1, *jt2;
	// PHILOX
	uint64_t gi;
	uint32_t k0, k1, dom;
};

__device__ __forceinline__ uint32_t replay_code(const DrawSrc &d, uint32_t k) {
	const uint32_t j = k + 1;
	LState s;
	if (j < (uint32_t)JUMP_TBL) {
		s.s1 = mulmod1(d.s.s1, __ldg(&d.jt1[j]));
		s.s2 = mulmod2(d.s.s2, __ldg(&d.jt2[j]));
	} else {
		s = ljump(d.s, j);
	}
	return lbase_code(lfloat(lout(s)));
}

// 16 uniform bits for nominal draw k of the read
__device__ __forceinline__ uint32_t philox_bits16(const DrawSrc &d, uint32_t k) {
	const Philox4 x = philox4x32((uint32_t)d.gi, (uint32_t)(d.gi >> 32), k >> 3, d.dom, d.k0, d.k1);
	const uint32_t w = x.v[(k >> 1) & 3];
	return (k & 1) ? (w >> 16) : (w & 0xffffu);
}

// Applies one mate's edit transcript to the reference and writes sequence and qualities
// in output orientation (simplesim.cpp:31-71 and the body of write, :111-126).
// All 32 lanes of the warp call this together.  dcount = draws the read has consumed so far.
template <int MODE>
__device__ void emit_aligned_mate(const SimArgs &a, const MateDesc &m, uint64_t gpos, uint64_t gend,
                                  const DrawSrc &ds, uint32_t &dcount, char *seq, char *qual) {
	const uint32_t lane = lane_id();
	const uint32_t lt = (1u << lane) - 1u;
	const char *xs = a.c.xs + m.xs_off;
	const uint32_t L = m.len;
	const bool fw = m.fw != 0;
	uint32_t rf_base = 0, rd_base = 0;
	for (uint32_t i0 = 0; i0 < m.xs_len; i0 += 32) {
		const uint32_t i = i0 + lane;
		const char op = i < m.xs_len ? xs[i] : 0;
		const bool cons = (op == '=') | (op == 'X') | (op == 'D') | (op == 'S');
		const bool prod = (op == '=') | (op == 'X') | (op == 'I') | (op == 'S');
		const bool drw = (op == 'X') | (op == 'I') | (op == 'S');
		const uint32_t cm = __ballot_sync(0xffffffffu, cons);
		const uint32_t pm = __ballot_sync(0xffffffffu, prod);
		const uint32_t dm = __ballot_sync(0xffffffffu, drw);
		const uint32_t rf = rf_base + __popc(cm & lt);
		const uint32_t rd = rd_base + __popc(pm & lt);
		uint32_t ref = 4u;
		if (op == '=' || op == 'X') ref = fetch_code(a.g, gpos + rf, gend);
		uint32_t code = ref;
		uint32_t used = __popc(dm);  // draws this group of 32 ops consumes
		if (MODE == QTIPSIM_RNG_REPLAY) {
			// Draw positions: the e-th draw op of the group starts at dcount+e; every rejected
			// X draw (equal to the reference base, simplesim.cpp:47-49) pushes itself and all
			// later draw ops one draw further.  Resolve in order: lanes below the frontier
			// are final, the lowest rejecting lane at or above it moves the frontier.
			uint32_t pos = dcount + __popc(dm & lt);
			uint32_t frontier = 0;
			if (dm) {
				for (;;) {
					if (drw && lane >= frontier) code = replay_code(ds, pos);
					const bool rej = (op == 'X') && lane >= frontier && code == ref;
					const uint32_t rm = __ballot_sync(0xffffffffu, rej);
					if (!rm) break;
					frontier = __ffs(rm) - 1;
					if (drw && lane >= frontier) pos++;
					used++;
				}
			}
		} else {
			if (drw) {
				const uint32_t k = dcount + __popc(dm & lt);
				const uint32_t b = philox_bits16(ds, k);
				if (op == 'X' && ref < 4u) code = (ref + 1u + ((b * 3u) >> 16)) & 3u;
				else code = b >> 14;
			}
		}
		if (prod) {
			if (fw) seq[rd] = code_char(code);
			else seq[L - 1 - rd] = code_char_comp(code);
		}
		rf_base += __popc(cm);
		rd_base += __popc(pm);
		dcount += used;
	}
	const char *q = a.c.qual + m.qual_off;
	for (uint32_t j = lane; j < L; j += 32) qual[j] = fw ? q[j] : q[L - 1 - j];
}

// The unaligned mate of a bad-end pair: opp_len random bases, quality 'I'
// (simplesim.h:89-117), written with the aligned mate's strand (simplesim.cpp:154-159).
template <int MODE>
__device__ void emit_random_mate(uint32_t L, bool fw, const DrawSrc &ds, uint32_t &dcount, char *seq,
                                 char *qual) {
	const uint32_t lane = lane_id();
	for (uint32_t j = lane; j < L; j += 32) {
		uint32_t code;
		if (MODE == QTIPSIM_RNG_REPLAY) code = replay_code(ds, dcount + j);
		else code = philox_rmate_code(philox4x32((uint32_t)ds.gi, (uint32_t)(ds.gi >> 32), j >> 6, PHILOX_DOM_RMATE | (ds.dom & 0xffu), ds.k0, ds.k1), j);
		if (fw) seq[j] = code_char(code);
		else seq[L - 1 - j] = code_char_comp(code);
		qual[j] = 'I';
	}
	dcount += L;
}

template <int MODE>
__global__ void __launch_bounds__(SIM_WARPS * 32) k_simulate(SimArgs a) {
	__shared__ char s_name[SIM_WARPS][NAME_SMEM];
	const uint32_t lane = lane_id();
	const uint32_t wib = threadIdx.x >> 5;
	const uint64_t r = (uint64_t)blockIdx.x * SIM_WARPS + wib;
	if (r >= a.nreads) return;
	const PlanRec pr = a.plan[r];
	const TmplDesc &t = a.c.t[pr.tmpl];
	const RecDev rec = a.g.recs[pr.record];
	const ReadGeom gm = read_geom(a, pr, t);
	const uint64_t gend = rec.gstart + rec.len;
	const bool paired_out = a.cat != CAT_U;
	char *o0 = a.out0 + a.off0[r];
	char *o1 = paired_out ? a.out1 + a.off1[r] : nullptr;

	// ---- name: composed once by lane 0, copied by the warp to one or both files ----
	const bool name_in_smem = gm.name_len <= (uint32_t)NAME_SMEM;
	char *nb = name_in_smem ? s_name[wib] : o0;
	if (lane == 0) {
		const char *refid = a.g.names + rec.name_off;
		char *p = nb;
		p = put_bytes(p, "@qtip!:", 7);
		if (a.cat == CAT_U) {
			p = put_origin(p, refid, rec.name_len, t.m[0].fw, gm.refoff[0], t.m[0].score, t.m[0].score_len);
		} else if (a.cat == CAT_B) {
			p = put_origin(p, refid, rec.name_len, t.m[0].fw, gm.refoff[0], t.m[0].score, t.m[0].score_len);
			*p++ = ':';
			p = put_origin(p, refid, rec.name_len, t.m[0].fw, gm.refoff[0], t.m[0].score, t.m[0].score_len);
		} else {
			p = put_origin(p, refid, rec.name_len, t.m[0].fw, gm.refoff[0], t.m[0].score, t.m[0].score_len);
			*p++ = ':';
			p = put_origin(p, refid, rec.name_len, t.m[1].fw, gm.refoff[1], t.m[1].score, t.m[1].score_len);
		}
		*p++ = ':';
		p = put_bytes(p, t.typ, t.typ_len);
		*p++ = '\n';
	}
	__syncwarp();
	if (name_in_smem) {
		for (uint32_t j = lane; j < gm.name_len; j += 32) {
			const char ch = nb[j];
			o0[j] = ch;
			if (paired_out) o1[j] = ch;
		}
	} else if (paired_out) {
		__threadfence_block();
		for (uint32_t j = lane; j < gm.name_len; j += 32) o1[j] = o0[j];
	}

	// ---- bodies ----
	DrawSrc ds;
	ds.s = LState{pr.s1, pr.s2};
	ds.jt1 = a.jt1;
	ds.jt2 = a.jt2;
	ds.gi = a.read_index_base + r;
	ds.k0 = (uint32_t)a.philox_seed;
	ds.k1 = (uint32_t)(a.philox_seed >> 32);
	ds.dom = PHILOX_DOM_BASE | (uint32_t)a.cat;
	uint32_t dcount = 0;

	char *seq0 = o0 + gm.name_len;
	char *qual0 = seq0 + gm.body0 + 3;
	char *seq1 = paired_out ? o1 + gm.name_len : nullptr;
	char *qual1 = paired_out ? seq1 + gm.body1 + 3 : nullptr;

	if (a.cat == CAT_U) {
		emit_aligned_mate<MODE>(a, t.m[0], rec.gstart + gm.refoff[0], gend, ds, dcount, seq0, qual0);
	} else if (a.cat == CAT_B) {
		// aligned mate's draws first, then the random mate's (simplesim.cpp:299-331)
		char *aseq = t.aligned_is_mate1 ? seq0 : seq1, *aqual = t.aligned_is_mate1 ? qual0 : qual1;
		char *rseq = t.aligned_is_mate1 ? seq1 : seq0, *rqual = t.aligned_is_mate1 ? qual1 : qual0;
		emit_aligned_mate<MODE>(a, t.m[0], rec.gstart + gm.refoff[0], gend, ds, dcount, aseq, aqual);
		emit_random_mate<MODE>(t.opp_len, t.m[0].fw != 0, ds, dcount, rseq, rqual);
	} else {
		emit_aligned_mate<MODE>(a, t.m[0], rec.gstart + gm.refoff[0], gend, ds, dcount, seq0, qual0);
		emit_aligned_mate<MODE>(a, t.m[1], rec.gstart + gm.refoff[1], gend, ds, dcount, seq1, qual1);
	}
	// separators: "\n+\n" after the sequence, "\n" after the qualities
	if (lane < 4) {
		const char sep[4] = {'\n', '+', '\n', '\n'};
		const uint32_t p0 = lane < 3 ? gm.body0 + lane : 2u * gm.body0 + 3u;
		seq0[p0] = sep[lane];
		if (paired_out) {
			const uint32_t p1 = lane < 3 ? gm.body1 + lane : 2u * gm.body1 + 3u;
			seq1[p1] = sep[lane];
		}
	}
}

void launch_simulate(const SimArgs &a, cudaStream_t st) {
	if (!a.nreads) return;
	const unsigned grid = (unsigned)((a.nreads + SIM_WARPS - 1) / SIM_WARPS);
	if (a.rng_mode == QTIPSIM_RNG_REPLAY) k_simulate<QTIPSIM_RNG_REPLAY><<<grid, SIM_WARPS * 32, 0, st>>>(a);
	else k_simulate<QTIPSIM_RNG_PHILOX><<<grid, SIM_WARPS * 32, 0, st>>>(a);
	g_launches++;
}

}  // namespace qt
