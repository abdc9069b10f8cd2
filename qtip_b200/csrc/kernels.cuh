// kernels.cuh -- launch wrappers for the sm_100a kernels of libqtipsim.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qtipsim.h"
#include "common.h"
#include "lecuyer.h"

namespace qt {

static constexpr int JUMP_TBL = 4096;  // a^j mod m for j < JUMP_TBL, both components

struct DevGenome {
	const uint32_t *bases = nullptr;
	const uint32_t *nmask = nullptr;
	const uint32_t *nblk = nullptr;  // 1 bit per 64 bases: any N in the block
	const RecDev *recs = nullptr;
	const char *names = nullptr;
	uint32_t nrecs = 0;
};

struct DevCategory {
	const TmplDesc *t = nullptr;
	const char *qual = nullptr;
	const char *xs = nullptr;
	uint32_t n = 0;
	// compiled form (tiled kernel)
	const char *qual_out = nullptr;
	uint32_t qual_out_bytes = 0;
	const uint32_t *ops = nullptr;
	const SegDev *segs = nullptr;
};

struct SimArgs {
	DevGenome g;
	DevCategory c;
	const PlanRec *plan = nullptr;
	uint64_t nreads = 0;
	uint64_t read_index_base = 0;  // global index of plan[0] within its category (PHILOX counter)
	const uint64_t *off0 = nullptr;  // byte offset of each read's record in stream 0 / 1
	const uint64_t *off1 = nullptr;
	char *out0 = nullptr;
	char *out1 = nullptr;
	int cat = 0;
	int rng_mode = 0;
	uint64_t philox_seed = 0;
	const uint32_t *jt1 = nullptr;  // jump tables (REPLAY)
	const uint32_t *jt2 = nullptr;
	const uint64_t *tot0 = nullptr;  // total bytes of stream 0 / 1 (device)
	const uint64_t *tot1 = nullptr;
	uint64_t genome_words = 0;       // u32 words in g.bases (for clamping speculative loads)
};

// number of N per window (simplesim.cpp:228-231 histogram, N bin only)
void launch_window_ncount(const DevGenome &g, const Window *win, uint32_t nwin, uint32_t *ncount,
                          cudaStream_t st);

// FASTQ record lengths per read for the category's stream(s); also accumulates the
// algorithmic genome bytes (sum of ceil(rflen/4)) into *gbytes.
void launch_read_lengths(const SimArgs &a, uint32_t *len0, uint32_t *len1, unsigned long long *gbytes,
                         cudaStream_t st);

// exclusive scan u32 -> u64 offsets; total written to *total (device).  tmp: >= (n/1024+2) u64.
void launch_scan(const uint32_t *len, uint64_t *off, uint64_t n, uint64_t *tmp, uint64_t *total,
                 cudaStream_t st);

// simulate + emit (SimulatedRead::init/mutate/init_random/write/write_pair)
// generic path: one warp per read, byte-granular
void launch_simulate(const SimArgs &a, cudaStream_t st);
// staged path (simulate_warp.cu): a warp stages a tile of consecutive records in its own slice of shared
// memory and ships it with one bulk copy; no CTA barriers.  max_rec_bytes = upper bound on one FASTQ record of
// this category.  Returns false when the category cannot use it (record too large for the staging buffer) --
// the caller falls back to the generic kernel.
bool launch_simulate_warp(const SimArgs &a, uint32_t max_rec_bytes, uint32_t max_body, cudaStream_t st);
// direct path (simulate_direct.cu): interior chunks go from registers to HBM, only a record's edge chunks are
// staged.  max_name bounds the read name; returns false when the category cannot use it (caller falls back).
bool launch_simulate_direct(const SimArgs &a, uint32_t max_name, uint32_t min_body, uint32_t max_body, cudaStream_t st);
// tile path (simulate_tile.cu): second generation of the direct path and the production kernel -- one edge image per
// record written whole by the chunk lanes, names appended as 16-byte pieces, one flat shipping loop.  Same contract.
bool launch_simulate_tile(const SimArgs &a, uint32_t max_name, uint32_t min_body, uint32_t max_body, cudaStream_t st);
void launch_nblock(const uint32_t *nmask, uint64_t mask_words, uint32_t *nblk, uint64_t nblk_words, cudaStream_t st);
void kernels_count_launch();
// plain copy done by a kernel (either side may be pinned, mapped host memory): control data of the streaming path
void launch_ctl_copy(void *dst, const void *src, size_t n, cudaStream_t st);

// PHILOX planning: per-window counts on the device
void launch_philox_counts(const Window *win, const uint32_t *ncount, const uint32_t *pclass, const BinomSetup *setups,
                          uint32_t nwin, uint64_t olap, uint64_t target, uint64_t seed, int cat, uint32_t *counts,
                          cudaStream_t st);
// K0: FASTA text -> packed genome (pack.cu)
void launch_find_gt(const uint8_t *text, uint64_t n, uint64_t *list, uint32_t cap, uint32_t *count, cudaStream_t st);
void launch_count_bases(const uint8_t *text, const uint64_t *seg_start, const uint64_t *seg_end,
                        const uint64_t *unit_base, uint32_t nseg, uint64_t nunits, uint32_t *counts, cudaStream_t st);
void launch_pack_units(const uint8_t *text, const uint64_t *seg_start, const uint64_t *seg_end,
                       const uint64_t *unit_base, uint32_t nseg, uint64_t nunits, const uint64_t *base_of_unit,
                       const int64_t *seg_gbase, uint32_t *bases, uint32_t *nmask, cudaStream_t st);
void launch_fill_n(const uint64_t *lo, const uint64_t *hi, uint32_t n, uint32_t *nmask, cudaStream_t st);

// plan + record lengths in one pass over the reads (PHILOX)
void launch_plan_philox_lengths(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap, uint32_t ntmpl, uint64_t seed,
                                const SimArgs &a, PlanRec *plan, uint32_t *len0, uint32_t *len1, unsigned long long *gbytes,
                                cudaStream_t st);

// the same with the record lengths scanned into stream offsets in one pass (decoupled look-back); status: scratch of
// plan_philox_scan_status_bytes(nreads) bytes
void launch_plan_philox_scan(const Window *win, const uint64_t *cum, uint32_t nwin, uint64_t olap, uint32_t ntmpl, uint64_t seed,
                             const SimArgs &a, PlanRec *plan, uint64_t *off0, uint64_t *off1, uint64_t *tot0, uint64_t *tot1,
                             unsigned long long *gbytes, unsigned long long *status, cudaStream_t st);
size_t plan_philox_scan_status_bytes(uint64_t nreads);

// mate 1 / mate 2 records of a pair category interleaved into one stream
void launch_interleave(const char *src0, const char *src1, const uint64_t *off0, const uint64_t *off1, uint64_t n, uint64_t tot0, uint64_t tot1,
                       char *dst, cudaStream_t st);
// tandem-SAM correctness labelling (label.cu)
void launch_label_correctness(const char *qnames, const uint64_t *q_off, const char *rnames, const uint64_t *r_off, const int32_t *flag,
                              const uint64_t *pos, uint64_t n, int32_t wiggle, int8_t *correct, cudaStream_t st);

// SAM fields -> edit transcripts (xscript.cu)
void launch_xs_lengths(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off, const uint64_t *pos, uint64_t n,
                       uint32_t *len1, int32_t *lclip, int32_t *rclip, int32_t *status, cudaStream_t st);
void launch_xs_fill(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off, const uint64_t *pos, uint64_t n,
                    const uint64_t *x_off, const int32_t *status, char *xs, uint64_t *lpos, uint64_t *rpos, cudaStream_t st);

int kernels_launch_count();  // kernels launched through the wrappers above since process start

}  // namespace qt
