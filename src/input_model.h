// src/input_model.h -- drop-in for the reference's src/input_model.h: InputModelUnpaired
// (:22-94) and InputModelPaired (:96-167).  A model only DESCRIBES the input here -- a reference to
// the caller's template list (not owned, input_model.h:88,161), how many alignments were offered to
// the reservoir, and the longest / average span -- because drawing from it happens on the GPU
// (qtip_b200/csrc/resolve.cu, kernels.cu).  One class template serves both kinds; the two names
// qtip_parse.cpp:1826-1829 constructs are thin subclasses that inherit its constructor.
#ifndef QTIP_B200_INPUT_MODEL_H
#define QTIP_B200_INPUT_MODEL_H

#include <algorithm>
#include <cassert>
#include "ds.h"
#include "rnglib.hpp"
#include "template.h"

template <class Tmpl>
class InputModelOf {
public:
	// (templates kept by the reservoir, alignments offered to it, two tuning knobs the reference stores and
	// never reads: input_model.h:26-30,100-104)
	InputModelOf(const EList<Tmpl> &kept, size_t offered, float fraction_even, float low_score_bias)
	    : kept_(kept), offered_(offered), knobs_{fraction_even, low_score_bias}, span_max_(0), span_avg_(0.0f) {
		const size_t k = kept.size();
		for (size_t i = 0; i < k; i++) {
			const size_t sp = kept[i].span();
			span_max_ = std::max(span_max_, sp);
			span_avg_ += (float)sp / k;  // same accumulation order as input_model.h:32-36
		}
	}
	bool empty() const { return kept_.empty(); }
	size_t num_added() const { return offered_; }  // input_model.h:68-70
	size_t max_len() const { return span_max_; }    // input_model.h:82-84: sets the windows' overlap
	float avg_len() const { return span_avg_; }
	const EList<Tmpl> &templates() const { return kept_; }
	// uniform pick with the global generator (input_model.h:49-54); kept for API compatibility -- the simulator
	// makes this draw on the device, in the reference's order
	const Tmpl &draw() const {
		assert(!empty());
		const size_t k = kept_.size();
		return kept_[std::min((size_t)(r4_uni_01() * k), k - 1)];
	}

private:
	const EList<Tmpl> &kept_;
	size_t offered_;
	float knobs_[2];
	size_t span_max_;
	float span_avg_;
};

class InputModelUnpaired : public InputModelOf<TemplateUnpaired> {
public:
	using InputModelOf<TemplateUnpaired>::InputModelOf;
};

class InputModelPaired : public InputModelOf<TemplatePaired> {
public:
	using InputModelOf<TemplatePaired>::InputModelOf;
};

#endif
