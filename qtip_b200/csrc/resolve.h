// resolve.h -- arguments of the device-side replay resolver (resolve.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.h"
#include "kernels.cuh"
#include "lecuyer.h"
#include "templates.h"

namespace qt {

struct ResolveArgs {
	DevGenome g;
	const Window *win;
	const uint32_t *ncount;
	const uint32_t *pclass;       // per window: row of the binomial set-up table
	const BinomSetup *setups;     // [npclass][NCAT]
	uint32_t win_lo, win_hi;
	uint32_t nwin;                // windows in the table
	bool n_skip_ok;               // chunksz - olap >= longest reference span: a mate never reaches past the next window
	uint64_t olap;
	// per category
	const TmplDesc *t[NCAT];
	const XStep *xsteps[NCAT];
	const HostMateProg *prog[NCAT];
	const XProgRec *xp[NCAT];     // packed draw programs (resolve2.cu); null: walk the XStep lists
	uint32_t ntmpl[NCAT];
	uint64_t target[NCAT];
	PlanRec *plan[NCAT];
	uint64_t cap[NCAT];
	const uint32_t *jt1, *jt2;
	// in/out
	uint32_t s1, s2;
	// results: [0..3] reads per category, [4] s1, [5] s2, [6] draws, [7] error code
	unsigned long long *result;
};

void launch_resolve_replay(const ResolveArgs &a, cudaStream_t st);

// ---- second-generation resolver (resolve2.cu): band speculation over many windows per launch ----
static constexpr uint32_t R2_KMAX = 512;          // most windows (= CTAs) per stage
static constexpr uint32_t R2_RCAP_MAX = 32768;    // most positions of a dense table (three u16 tables in shared memory)
static constexpr uint32_t R2_EMIT_PTS = 1025;     // stride points per emit round (32 reads each)
enum { R2_RUNNING = 0, R2_DONE = 1, R2_STUCK = 2 };

// the last exactly known position of the chain, handed from stage to stage
struct StageAnchor {
	uint64_t abs;         // draws consumed before it
	uint32_t s1, s2;      // generator state there
	uint32_t first;       // index (active-window list) of the first window after it
	uint32_t status;      // R2_*
	uint32_t stuck_win;   // R2_STUCK: the window that could not be resolved
	uint32_t pad;
};

// one resolved window: positions are relative to the anchor of the stage that resolved it
struct WinResolved {
	uint32_t stage;
	uint32_t rel;         // position of the window's first binomial draw
	uint32_t end[NCAT];   // position after the reads of U, B, C, D
};

struct R2Offsets {
	const uint64_t *p[NCAT];
};

struct Resolve2Args {
	ResolveArgs ra;               // genome, windows, set-ups, templates, jump tables (win_lo / win_hi / plan caps unused)
	const uint32_t *jh1, *jh2;    // a^(4096 k) mod m, k < 4096
	const BinomTables *tabs;      // [npclass][NCAT], parallel to ra.setups
	const uint32_t *aw;           // the shard's live windows, in order
	uint32_t na;
	const double *cumM, *cumV;    // [na + 1] prefix sums of the mean / variance of the draws a window consumes
	float m1[NCAT], v1[NCAT];     // draws per read: mean, variance
	uint32_t maxfixed[NCAT];
	float band_sigma;             // half-width of a band in standard deviations
	uint32_t rcap, rcap_emit;     // dense-table positions (stage / emit kernels)
	uint32_t nbmax;               // most candidates per window
	StageAnchor *anchor;          // [launches + 1]
	uint32_t *counter;            // [launches], zeroed: CTAs of the launch that are done
	uint32_t *ticket;             // [launches], zeroed: window slots of the launch handed out
	uint2 *slot;                  // [K] band of each CTA's window: {lo, nb}
	unsigned long long *relay;    // [K + 1] (stage + 1) << 32 | exact start of window j of the stage (zeroed per run)
	uint32_t *end;                // [NCAT][K][nbmax] candidate positions after each category
	uint32_t *cn, *cnd;           // [K][nbmax] candidate counts
	WinResolved *wres;            // [na]
	uint32_t walk_max;            // regions of at most this many reads are emitted by one walking thread, the others by a CTA
	uint32_t any_large;           // some region may exceed walk_max: launch the CTA version too
	unsigned long long *dbg;      // tracing only (QTIPSIM_R2_DBG): [K][16] globaltimer stamps of one stage's CTAs, or null
	uint32_t dbg_stage;
};

void launch_resolve2_stage(const Resolve2Args &a, uint32_t s, uint32_t K, cudaStream_t st);
void launch_resolve2_counts(const Resolve2Args &a, uint32_t *cnt, uint32_t *start, uint32_t *err, cudaStream_t st);
void launch_resolve2_emit(const Resolve2Args &a, const uint32_t *cnt, const uint32_t *start, const R2Offsets &off, uint32_t *err,
                          cudaStream_t st);

}  // namespace qt
