// src/template.h -- drop-in for the reference's src/template.h (TemplateUnpaired :28-105,
// TemplatePaired :123-231): same constructors / init() signatures, value semantics.
#ifndef QTIP_B200_TEMPLATE_H
#define QTIP_B200_TEMPLATE_H

#include <cstddef>
#include <string>
#include "edit_xscript.h"

struct TemplateUnpaired {
	TemplateUnpaired() : best_score_(0), fw_flag_(0), len_(0), mate_flag_(0), opp_len_(0) {}
	TemplateUnpaired(int best_score, int len, char fw_flag, char mate_flag, int opp_len, const char *qual,
	                 const char *edit_xscript) {
		init(best_score, len, fw_flag, mate_flag, opp_len, qual, edit_xscript);
	}
	void init(int best_score, int len, char fw_flag, char mate_flag, int opp_len, const char *qual,
	          const char *edit_xscript) {
		best_score_ = best_score;
		len_ = len;
		fw_flag_ = fw_flag;
		mate_flag_ = mate_flag;
		opp_len_ = opp_len;
		qual_ = qual ? qual : "";
		edit_xscript_ = edit_xscript ? edit_xscript : "";
	}
	size_t reflen() const { return edit_xscript_to_rflen(edit_xscript_.c_str()); }

	int best_score_;
	char fw_flag_;    // 'T' / 'F'
	int len_;
	char mate_flag_;  // '0' / '1' / '2'
	int opp_len_;
	std::string qual_;
	std::string edit_xscript_;
};

struct TemplatePaired {
	TemplatePaired()
	    : score_12_(0), score_1_(0), len_1_(0), fw_flag_1_(0), score_2_(0), len_2_(0), fw_flag_2_(0),
	      upstream1_(false), fraglen_(0) {}
	TemplatePaired(int score_12, int score_1, int len_1, char fw_flag_1, const char *qual_1, const char *edit_xscript_1,
	               int score_2, int len_2, char fw_flag_2, const char *qual_2, const char *edit_xscript_2, bool upstream1,
	               size_t fraglen) {
		init(score_12, score_1, len_1, fw_flag_1, qual_1, edit_xscript_1, score_2, len_2, fw_flag_2, qual_2,
		     edit_xscript_2, upstream1, fraglen);
	}
	void init(int score_12, int score_1, int len_1, char fw_flag_1, const char *qual_1, const char *edit_xscript_1,
	          int score_2, int len_2, char fw_flag_2, const char *qual_2, const char *edit_xscript_2, bool upstream1,
	          size_t fraglen) {
		score_12_ = score_12;
		score_1_ = score_1;
		len_1_ = len_1;
		fw_flag_1_ = fw_flag_1;
		qual_1_ = qual_1 ? qual_1 : "";
		edit_xscript_1_ = edit_xscript_1 ? edit_xscript_1 : "";
		score_2_ = score_2;
		len_2_ = len_2;
		fw_flag_2_ = fw_flag_2;
		qual_2_ = qual_2 ? qual_2 : "";
		edit_xscript_2_ = edit_xscript_2 ? edit_xscript_2 : "";
		upstream1_ = upstream1;
		fraglen_ = fraglen;
	}

	int score_12_;
	int score_1_;
	int len_1_;
	char fw_flag_1_;
	std::string qual_1_;
	std::string edit_xscript_1_;
	int score_2_;
	int len_2_;
	char fw_flag_2_;
	std::string qual_2_;
	std::string edit_xscript_2_;
	bool upstream1_;
	size_t fraglen_;
};

#endif
