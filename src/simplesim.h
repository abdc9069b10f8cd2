// src/simplesim.h -- drop-in for the reference's src/simplesim.h as seen by its one caller
// (qtip_parse.cpp:1825-1867): sim_startswith / sim_sep (:20-24), FUNC_* (:197-201) and
// StreamingSimulator (:203-299) with the same constructor and simulate_batch signature.  The
// class forwards to libqtipsim's C ABI (include/qtipsim.h); implementation in
// simplesim_shim.cpp.
#ifndef QTIP_B200_SIMPLESIM_H
#define QTIP_B200_SIMPLESIM_H

#include <stdio.h>
#include <string>
#include <vector>
#include "fasta.h"
#include "input_model.h"
#include "ranlib.hpp"

#define SIM_STARTSWITH_LITERAL "qtip!"
#define SIM_SEPARATOR_LITERAL ':'

static const char *sim_startswith = SIM_STARTSWITH_LITERAL;
static const char sim_sep = SIM_SEPARATOR_LITERAL;

enum { FUNC_LINEAR = 1, FUNC_SQRT, FUNC_CONST };

struct qtipsim;

class StreamingSimulator {
public:
	StreamingSimulator(const std::vector<std::string> &fns, size_t chunksz, const InputModelUnpaired &model_u,
	                   const InputModelUnpaired &model_b, const InputModelPaired &model_c,
	                   const InputModelPaired &model_d, FILE *fh_u, FILE *fh_b_1, FILE *fh_b_2, FILE *fh_c_1,
	                   FILE *fh_c_2, FILE *fh_d_1, FILE *fh_d_2);
	~StreamingSimulator();

	// one pass over the FASTA; argument order u, c, d, b as in the reference (simplesim.h:251-257)
	void simulate_batch(float fraction, int function, size_t min_u, size_t min_c, size_t min_d, size_t min_b);

	size_t num_estimated_bases() const { return tot_fasta_len_; }

private:
	StreamingSimulator(const StreamingSimulator &);
	StreamingSimulator &operator=(const StreamingSimulator &);
	qtipsim *h_;
	size_t chunksz_;
	size_t olap_;
	size_t tot_fasta_len_;
	FILE *fh_[7];
};

#endif
