// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" face over the UNMODIFIED reference classes so tests and the
// bench's CPU-baseline leg can drive them through ctypes.  This file is ours; it is
// compiled together with the reference sources where they lie (oracle/Makefile,
// target _ref/libref.so) and contains no reference code.
//
//   ref_initialize/ref_set_seed/ref_get_state/ref_i4_uni/ref_r4_uni_01
//        -> rnglib.cpp:941, :1674, :516, :573, :1444
//   ref_ignbin            -> ranlib.cpp:1094
//   ref_reservoir_offsets -> ReservoirSampledEList::add_part1, ds.h:437-446
//   ref_simulate          -> StreamingSimulator ctor + simulate_batch
//                            (simplesim.h:214-257, simplesim.cpp:195-407), fed by
//                            InputModel{Unpaired,Paired} (input_model.h:26,100) built
//                            the way qtip_parse.cpp:849-861,1148-1166 builds them
//                            (EList::expand + back().init).
#include <stdio.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "ds.h"
#include "template.h"
#include "input_model.h"
#include "rnglib.hpp"
#include "ranlib.hpp"
#include "simplesim.h"

extern "C" {

// Same field order/meaning as include/qtipsim.h (kept in sync by tests/test_abi.py).
struct ref_tmpl_unp {
	int32_t best_score;
	int32_t len;
	char fw_flag;    // 'T' / 'F'
	char mate_flag;  // '0' / '1' / '2'
	int32_t opp_len;
	const char *qual;
	const char *xscript;
};

struct ref_tmpl_pair {
	int32_t score_12;
	int32_t score_1;
	int32_t len_1;
	char fw_flag_1;
	const char *qual_1;
	const char *xscript_1;
	int32_t score_2;
	int32_t len_2;
	char fw_flag_2;
	const char *qual_2;
	const char *xscript_2;
	int32_t upstream1;
	uint64_t fraglen;
};

void ref_initialize() { initialize(); }
void ref_set_seed(int a, int b) { set_seed(a, b); }
void ref_get_state(int *a, int *b) { int x, y; get_state(x, y); *a = x; *b = y; }
int ref_i4_uni() { return i4_uni(); }
float ref_r4_uni_01() { return r4_uni_01(); }

// ReservoirSampledEList<T>::add_part1 (ds.h:437-446) called n times on a reservoir of capacity k: the slot every call
// returns (the call sites compare it with k(), qtip_parse.cpp:849-850, :1148-1149).  Consumes the global generator.
uint64_t ref_reservoir_offsets(uint64_t k, uint64_t n, uint64_t *off) {
	ReservoirSampledEList<int> r((size_t)k);
	for (uint64_t i = 0; i < n; i++) off[i] = (uint64_t)r.add_part1();
	return (uint64_t)r.list().size();
}
int ref_ignbin(int n, float p) { return ignbin(n, p); }

// Runs the reference simulator.  paths[7] = u, b_1, b_2, c_1, c_2, d_1, d_2.
// *_added = number of alignments "seen" (InputModel n); may exceed the kept count.
// RNG state is whatever the process-global rnglib state currently is (set it with
// ref_set_seed first); it is advanced by the call exactly as in the reference.
int ref_simulate(
	const char *fasta, uint64_t chunksz,
	const ref_tmpl_unp *tu, uint64_t nu, uint64_t u_added,
	const ref_tmpl_unp *tb, uint64_t nb, uint64_t b_added,
	const ref_tmpl_pair *tc, uint64_t nc, uint64_t c_added,
	const ref_tmpl_pair *td, uint64_t nd, uint64_t d_added,
	float fraction, int function,
	uint64_t min_u, uint64_t min_c, uint64_t min_d, uint64_t min_b,
	const char *const *paths)
{
	EList<TemplateUnpaired> lu, lb;
	EList<TemplatePaired> lc, ld;
	for(uint64_t i = 0; i < nu; i++) {
		lu.expand();
		lu.back().init(tu[i].best_score, tu[i].len, tu[i].fw_flag, tu[i].mate_flag,
		               tu[i].opp_len, tu[i].qual, tu[i].xscript);
	}
	for(uint64_t i = 0; i < nb; i++) {
		lb.expand();
		lb.back().init(tb[i].best_score, tb[i].len, tb[i].fw_flag, tb[i].mate_flag,
		               tb[i].opp_len, tb[i].qual, tb[i].xscript);
	}
	for(uint64_t i = 0; i < nc; i++) {
		lc.expand();
		lc.back().init(tc[i].score_12, tc[i].score_1, tc[i].len_1, tc[i].fw_flag_1,
		               tc[i].qual_1, tc[i].xscript_1, tc[i].score_2, tc[i].len_2,
		               tc[i].fw_flag_2, tc[i].qual_2, tc[i].xscript_2,
		               tc[i].upstream1 != 0, (size_t)tc[i].fraglen);
	}
	for(uint64_t i = 0; i < nd; i++) {
		ld.expand();
		ld.back().init(td[i].score_12, td[i].score_1, td[i].len_1, td[i].fw_flag_1,
		               td[i].qual_1, td[i].xscript_1, td[i].score_2, td[i].len_2,
		               td[i].fw_flag_2, td[i].qual_2, td[i].xscript_2,
		               td[i].upstream1 != 0, (size_t)td[i].fraglen);
	}
	InputModelUnpaired mu(lu, (size_t)u_added, 1.0f, 1.0f);
	InputModelUnpaired mb(lb, (size_t)b_added, 1.0f, 1.0f);
	InputModelPaired mc(lc, (size_t)c_added, 1.0f, 1.0f);
	InputModelPaired md(ld, (size_t)d_added, 1.0f, 1.0f);
	FILE *fh[7];
	static char bufs[7][65536];
	for(int i = 0; i < 7; i++) {
		fh[i] = fopen(paths[i], "wb");
		if(fh[i] == NULL) {
			return -1;
		}
		setvbuf(fh[i], bufs[i], _IOFBF, 65536);  // qtip_parse.cpp:1535-1545
	}
	int rc = 0;
	try {
		std::vector<std::string> fns;
		fns.push_back(std::string(fasta));
		StreamingSimulator ss(fns, (size_t)chunksz, mu, mb, mc, md,
		                      fh[0], fh[1], fh[2], fh[3], fh[4], fh[5], fh[6]);
		ss.simulate_batch(fraction, function, (size_t)min_u, (size_t)min_c,
		                  (size_t)min_d, (size_t)min_b);
	} catch(int e) {
		rc = e ? e : 1;
	}
	for(int i = 0; i < 7; i++) {
		fclose(fh[i]);
	}
	return rc;
}

}  // extern "C"
