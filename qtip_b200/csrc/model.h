// model.h -- the input model's reservoirs on this side of the boundary.
//
// The reference's SAM pass offers every aligned record to one of four reservoirs (qtip_parse.cpp:849-861 unpaired and
// bad-end, :1148-1166 concordant and discordant; ReservoirSampledEList::add_part1, ds.h:437-446) and the simulator is
// then built from what they kept (qtip_parse.cpp:1825-1829).  ModelBuilder takes the same records as batches of plain
// arrays -- the transcripts as qtipsim_edit_xscripts wrote them -- keeps what the reference keeps, consumes the
// generator the way the reference does (one r4_uni_01 per record once a reservoir is full, all four reservoirs on the
// one global stream, in SAM order), and hands the result to the template store.
//
// What the call sites do, restated: off = add_part1(); if (off < k) list().back().init(record).  While the reservoir
// is filling, add_part1 appends an element and the record lands in it.  Once it is full, off = min((size_t)(r4 * n),
// n - 1) with n the records offered so far, and a record with off < k overwrites the LAST element -- not element
// off -- because the call sites write through back().  That is the reference's behaviour and therefore ours.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/qtipsim.h"
#include "common.h"
#include "lecuyer.h"
#include "templates.h"

namespace qt {

struct ModelBuilder {
	struct Mate {
		int32_t score = 0, len = 0;
		char fw = 'T';
		std::string qual, xs;
	};
	struct Rec {
		Mate m[2];
		char mate_flag = '0';
		int32_t opp_len = 0;
		int32_t upstream1 = 0;
		uint64_t fraglen = 0;
	};
	uint64_t k[NCAT] = {0, 0, 0, 0};  // capacities (input-model-size, qtip_parse.cpp:1560)
	uint64_t n[NCAT] = {0, 0, 0, 0};  // records offered (ReservoirSampledEList::size())
	std::vector<Rec> kept[NCAT];
	bool open = false;

	void begin(const uint64_t cap[NCAT]);
	// offers the batch's records in order; rng is the global generator (advanced by the draws consumed)
	int add(const qtipsim_model_batch &b, LState &rng, std::string &err);
	// the four models as the simulator sees them: InputModel*(list(), size(), ...), qtip_parse.cpp:1826-1829
	int commit(TemplateStore &ts, std::string &err);
};

}  // namespace qt
