// planner.h -- resolves the reference's sequential draw chain into a per-read plan.
//
// In the reference every random decision comes from one global generator stepped in
// program order (simplesim.cpp:216-397): per window four binomial counts, and per read a
// template draw, an offset draw and a data-dependent number of base draws.  The plan
// records, for each read, what those two index draws selected and the generator state
// the read's base draws start from -- after which every read is independent and the
// simulate kernel can process them in any order.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "common.h"
#include "genome.h"
#include "lecuyer.h"
#include "templates.h"

namespace qt {

struct PlanParams {
	uint64_t chunksz = 0, olap = 0;
	uint64_t target[NCAT] = {0, 0, 0, 0};  // nu, nb, nc, nd after apply_function
	uint64_t tot_fasta_len = 0;
	uint64_t win_lo = 0, win_hi = 0;       // windows of this shard; the generator state passed in is
	                                       // the state at window win_lo (rank r-1 hands it to rank r)
};

struct Plan {
	std::vector<PlanRec> recs[NCAT];
	LState rng_out{0, 0};
	uint64_t draws = 0;
};

// simplesim.cpp:177-190
uint64_t apply_function(float fraction, int function, uint64_t mn, uint64_t n);

// simplesim.cpp:224-226: the per-window success probability
float window_binom_p(uint64_t retsz, uint64_t olap, uint64_t tot_fasta_len);
// simplesim.cpp:221-223, :228-234: is the window simulated from at all?
inline bool window_live(uint64_t retsz, uint64_t olap, uint64_t n_count) {
	if (retsz < olap) return false;
	return !((int)n_count > (int)(0.9 * (double)retsz));
}

// Host walk of the chain (REPLAY).  ncount[i] = number of N in window i.
// Returns QTIPSIM_OK or an error code with text in err.
int plan_replay_host(const HostGenome &g, const TemplateStore &ts, const std::vector<Window> &win,
                     const std::vector<uint32_t> &ncount, const PlanParams &pp, LState rng, Plan &out,
                     std::string &err);

}  // namespace qt
