// src/input_model.h -- drop-in for the reference's src/input_model.h: InputModelUnpaired
// (:22-94) and InputModelPaired (:96-167).  They only describe the model (a reference to the
// caller's template list + counts); drawing from it happens on the GPU.
#ifndef QTIP_B200_INPUT_MODEL_H
#define QTIP_B200_INPUT_MODEL_H

#include <cassert>
#include <algorithm>
#include "ds.h"
#include "template.h"
#include "rnglib.hpp"

class InputModelUnpaired {
public:
	InputModelUnpaired(const EList<TemplateUnpaired> &ts, size_t n, float fraction_even, float low_score_bias)
	    : ts_(ts), n_(n), fraction_even_(fraction_even), low_score_bias_(low_score_bias) {
		len_avg_ = 0.0f;
		len_max_ = 0;
		for (size_t i = 0; i < ts.size(); i++) {
			len_avg_ += ((float)ts[i].reflen() / ts.size());
			len_max_ = std::max(len_max_, ts[i].reflen());
		}
	}
	// uniform draw, input_model.h:49-54 (kept for API compatibility; the simulator does not call it)
	const TemplateUnpaired &draw() const {
		assert(!empty());
		return ts_[std::min((size_t)(r4_uni_01() * ts_.size()), ts_.size() - 1)];
	}
	bool empty() const { return ts_.empty(); }
	size_t num_added() const { return n_; }
	float avg_len() const { return len_avg_; }
	size_t max_len() const { return len_max_; }
	const EList<TemplateUnpaired> &templates() const { return ts_; }

protected:
	const EList<TemplateUnpaired> &ts_;
	float len_avg_;
	size_t n_;
	size_t len_max_;
	float fraction_even_;   // unused, as in the reference
	float low_score_bias_;  // unused
};

class InputModelPaired {
public:
	InputModelPaired(const EList<TemplatePaired> &ts, size_t n, float fraction_even, float low_score_bias)
	    : ts_(ts), n_(n), fraction_even_(fraction_even), low_score_bias_(low_score_bias) {
		len_avg_ = 0.0f;
		len_max_ = 0;
		for (size_t i = 0; i < ts.size(); i++) {
			len_avg_ += ((float)ts[i].fraglen_ / ts.size());
			len_max_ = std::max(len_max_, ts[i].fraglen_);
		}
	}
	const TemplatePaired &draw() const {
		assert(!empty());
		return ts_[std::min((size_t)(r4_uni_01() * ts_.size()), ts_.size() - 1)];
	}
	bool empty() const { return ts_.empty(); }
	size_t num_added() const { return n_; }
	float avg_len() const { return len_avg_; }
	size_t max_len() const { return len_max_; }
	const EList<TemplatePaired> &templates() const { return ts_; }

protected:
	const EList<TemplatePaired> &ts_;
	float len_avg_;
	size_t n_;
	size_t len_max_;
	float fraction_even_;
	float low_score_bias_;
};

#endif
