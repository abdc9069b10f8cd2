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

}  // namespace qt
