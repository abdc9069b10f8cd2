// src/simplesim_shim.cpp -- StreamingSimulator over libqtipsim's C ABI.
//
// Mirrors what the reference object does for its caller (qtip_parse.cpp:1840-1858):
//   ctor           : olap_ = max over models (simplesim.h:228-230), assert(chunksz > olap)
//                    (fasta.h:41), tot_fasta_len_ = sum of file sizes (simplesim.h:272-286)
//   simulate_batch : reads the global generator, runs the GPU simulator in replay mode (or Philox
//                    when QTIP_RNG=philox), appends the seven FASTQ streams to the caller's FILE*s,
//                    prints the four "Wrote ..." lines (simplesim.cpp:399-406), and leaves the global
//                    generator where the reference would have left it.
// Errors the reference raises with `throw 1` are raised the same way (uncaught -> abort -> the
// Python driver sees a non-zero exit, qtip:479-480).
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <cassert>
#include <iostream>
#include "../include/qtipsim.h"
#include "simplesim.h"

namespace {

void check(qtipsim *h, int rc) {
	if (rc != QTIPSIM_OK) {
		std::cerr << qtipsim_last_error(h) << std::endl;
		throw 1;
	}
}

template <class Model>
void upload_unpaired(qtipsim *h, int cat, const Model &m) {
	const EList<TemplateUnpaired> &ts = m.templates();
	std::vector<qtipsim_tmpl_unp> v(ts.size());
	for (size_t i = 0; i < ts.size(); i++) {
		v[i].best_score = ts[i].m.score;
		v[i].len = ts[i].m.len;
		v[i].fw_flag = ts[i].m.fw;
		v[i].mate_flag = ts[i].mate_flag;
		v[i].opp_len = ts[i].opp_len;
		v[i].qual = ts[i].m.qual.c_str();
		v[i].xscript = ts[i].m.xscript.c_str();
	}
	check(h, qtipsim_set_templates_unpaired(h, cat, v.data(), v.size(), m.num_added()));
}

void upload_paired(qtipsim *h, int cat, const InputModelPaired &m) {
	const EList<TemplatePaired> &ts = m.templates();
	std::vector<qtipsim_tmpl_pair> v(ts.size());
	for (size_t i = 0; i < ts.size(); i++) {
		v[i].score_12 = ts[i].score_12;
		v[i].score_1 = ts[i].m1.score;
		v[i].len_1 = ts[i].m1.len;
		v[i].fw_flag_1 = ts[i].m1.fw;
		v[i].qual_1 = ts[i].m1.qual.c_str();
		v[i].xscript_1 = ts[i].m1.xscript.c_str();
		v[i].score_2 = ts[i].m2.score;
		v[i].len_2 = ts[i].m2.len;
		v[i].fw_flag_2 = ts[i].m2.fw;
		v[i].qual_2 = ts[i].m2.qual.c_str();
		v[i].xscript_2 = ts[i].m2.xscript.c_str();
		v[i].upstream1 = ts[i].upstream1 ? 1 : 0;
		v[i].fraglen = ts[i].fraglen;
	}
	check(h, qtipsim_set_templates_paired(h, cat, v.data(), v.size(), m.num_added()));
}

}  // namespace

StreamingSimulator::StreamingSimulator(const std::vector<std::string> &fns, size_t chunksz,
                                       const InputModelUnpaired &model_u, const InputModelUnpaired &model_b,
                                       const InputModelPaired &model_c, const InputModelPaired &model_d, FILE *fh_u,
                                       FILE *fh_b_1, FILE *fh_b_2, FILE *fh_c_1, FILE *fh_c_2, FILE *fh_d_1,
                                       FILE *fh_d_2)
    : h_(NULL), chunksz_(chunksz) {
	fh_[0] = fh_u;
	fh_[1] = fh_b_1;
	fh_[2] = fh_b_2;
	fh_[3] = fh_c_1;
	fh_[4] = fh_c_2;
	fh_[5] = fh_d_1;
	fh_[6] = fh_d_2;
	olap_ = std::max(model_u.max_len(), std::max(model_b.max_len(), std::max(model_c.max_len(), model_d.max_len())));
	assert(chunksz > olap_);  // FastaChunkwiseParser ctor, fasta.h:41
	int rc = qtipsim_create(&h_, -1);
	if (rc != QTIPSIM_OK) {
		std::cerr << qtipsim_last_error(NULL) << std::endl;
		throw 1;
	}
	std::vector<const char *> paths(fns.size());
	for (size_t i = 0; i < fns.size(); i++) paths[i] = fns[i].c_str();
	check(h_, qtipsim_load_fasta(h_, paths.data(), paths.size()));
	tot_fasta_len_ = (size_t)qtipsim_num_estimated_bases(h_);
	upload_unpaired(h_, QTIPSIM_CAT_U, model_u);
	upload_unpaired(h_, QTIPSIM_CAT_B, model_b);
	upload_paired(h_, QTIPSIM_CAT_C, model_c);
	upload_paired(h_, QTIPSIM_CAT_D, model_d);
	assert(qtipsim_olap(h_) == olap_);
}

StreamingSimulator::~StreamingSimulator() { qtipsim_destroy(h_); }

void StreamingSimulator::simulate_batch(float fraction, int function, size_t min_u, size_t min_c, size_t min_d,
                                        size_t min_b) {
	qtipsim_run_params p;
	memset(&p, 0, sizeof(p));
	p.chunksz = chunksz_;
	p.fraction = fraction;
	p.function = function;
	p.min_u = min_u;
	p.min_c = min_c;
	p.min_d = min_d;
	p.min_b = min_b;
	const char *mode = getenv("QTIP_RNG");
	p.rng_mode = (mode && strcmp(mode, "philox") == 0) ? QTIPSIM_RNG_PHILOX : QTIPSIM_RNG_REPLAY;
	int cg1, cg2;
	get_state(cg1, cg2);
	p.cg1 = cg1;
	p.cg2 = cg2;
	p.philox_seed = ((uint64_t)(uint32_t)cg1 << 32) | (uint32_t)cg2;
	p.shard_rank = 0;
	p.shard_count = 1;
	qtipsim_stats st;
	check(h_, qtipsim_run(h_, &p, &st));
	if (p.rng_mode == QTIPSIM_RNG_REPLAY) set_seed(p.cg1, p.cg2);  // the global generator has advanced
	int fds[QTIPSIM_NOUT];
	for (int i = 0; i < QTIPSIM_NOUT; i++) {
		fds[i] = -1;
		if (fh_[i] != NULL) {
			fflush(fh_[i]);  // the caller's stdio buffer is empty here, but keep the order safe
			fds[i] = fileno(fh_[i]);
		}
	}
	check(h_, qtipsim_result_write_fds(h_, fds));
	std::cerr << "    Wrote " << st.wrote[0] << " unpaired tandem reads (target=" << st.target[0] << ")" << std::endl;
	std::cerr << "    Wrote " << st.wrote[1] << " bad-end tandem reads (target=" << st.target[1] << ")" << std::endl;
	std::cerr << "    Wrote " << st.wrote[2] << " concordant tandem pairs (target=" << st.target[2] << ")" << std::endl;
	std::cerr << "    Wrote " << st.wrote[3] << " discordant tandem pairs (target=" << st.target[3] << ")" << std::endl;
}
