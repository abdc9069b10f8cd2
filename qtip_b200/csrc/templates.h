// templates.h -- the input model on the host: validated, flattened templates.
//
// Replaces InputModelUnpaired / InputModelPaired (input_model.h:22-167) and the per-read
// rescans of the transcript the reference does (template.h:92-95, simplesim.cpp:32,37,
// 355-356): lengths, reference spans and draw counts are computed once here.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/qtipsim.h"
#include "common.h"

namespace qt {

// Where one mate's draws fall: between consecutive X ops there are `skip` draws that are
// always consumed exactly once (I, S); an X at reference offset `rf` consumes one draw
// plus one more for every draw equal to the reference base (simplesim.cpp:45-50).
struct XStep {
	uint32_t skip;  // I/S draws before this X
	uint32_t rf;    // reference offset of the X, relative to the mate's leftmost base
};

struct HostMateProg {
	uint32_t xs_begin = 0, xs_end = 0;  // range in Category::xsteps
	uint32_t tail_skip = 0;             // I/S draws after the last X
};

// One mate's draw program packed for the device resolvers: everything read_len needs about a template in one
// 32-byte load (the XStep lists cost one dependent load per X).  nx = 0xff: the mate does not fit (more than 7 X, a
// reference offset over 65535 or more than 255 skipped draws) and the XStep list is walked instead.
struct alignas(32) XProgHalf {
	uint16_t rf[7];    // reference offset of each X, relative to the mate's leftmost base
	uint8_t skip[7];   // I/S draws before each X
	uint8_t nx;
	uint8_t tail;      // I/S draws after the last X
	uint8_t pad;
	uint32_t extra0;   // half 0: opp_len (B);  half 1: delta[1]
	uint32_t extra1;   // half 0: delta[0]
};
struct alignas(64) XProgRec {
	XProgHalf h[2];
};
static_assert(sizeof(XProgHalf) == 32 && sizeof(XProgRec) == 64, "XProg layout");

struct Category {
	std::vector<TmplDesc> t;
	std::string qual;                 // arena
	std::string xs;                   // arena
	std::string qual_out;             // arena, output orientation, 16-byte aligned entries
	std::vector<uint32_t> ops;        // compiled draw-op lists
	std::vector<SegDev> segs;         // compiled reference-copy segments
	bool compiled_ok = true;          // false: some mate exceeds the compiled format's limits
	uint32_t max_body = 0;            // longest read body (either mate / random mate)
	uint32_t min_body = 0xffffffffu;  // shortest
	uint32_t max_score_len = 0;
	uint32_t max_rflen = 0;           // longest reference span of a mate
	std::vector<XStep> xsteps;        // host-side draw programs (replay planner)
	std::vector<HostMateProg> prog;   // 2 per template
	std::vector<XProgRec> xprog;      // 1 per template (built by finish_xprog)
	void finish_xprog();
	uint64_t added = 0;               // num_added(), input_model.h:68-70
	uint64_t max_len = 0;             // max_len(), input_model.h:82-84 / :155-157
	bool arena_overflow = false;      // an arena outgrew the 32-bit descriptor offsets
	bool valid = true;                // all transcripts well-formed
	std::string why;                  // first problem found
	double mean_bytes = 0;            // expected FASTQ bytes per read, excluding refid/offset digits
	double mean_draws = 0;            // expected draws per read
	// draws one read / pair consumes in REPLAY mode, for the band placement of the speculative resolver
	// (resolve2.cu): 2 index draws + one per I/S (+ opp_len) + Geometric(3/4) per X (simplesim.cpp:45-50)
	double draw_m1 = 0, draw_v1 = 0;  // mean and variance over a uniformly drawn template
	uint32_t draw_max_fixed = 0;      // most draws of a template when every X takes one draw
	double sum_bytes = 0, sum_draws = 0, sum_dm = 0, sum_dsq = 0;  // running sums behind the four means above
	void clear() { *this = Category(); }
};

struct TemplateStore {
	Category cat[NCAT];
	int set_unpaired(int c, const qtipsim_tmpl_unp *v, size_t kept, uint64_t added, std::string &err);
	int set_paired(int c, const qtipsim_tmpl_pair *v, size_t kept, uint64_t added, std::string &err);
	uint64_t olap() const;  // simplesim.h:228-230
};

}  // namespace qt
