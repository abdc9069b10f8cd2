// src/template.h -- drop-in for the reference's src/template.h (TemplateUnpaired :28-105,
// TemplatePaired :123-231).  What qtip_parse.cpp needs from these types are the default and the
// full constructors and init(...) with the reference's argument order (qtip_parse.cpp:849-861,
// 1148-1166) and value semantics inside EList / ReservoirSampledEList; the fields are read only by
// src/input_model.h and src/simplesim_shim.cpp.  Both kinds are built from one mate record that owns
// its strings (the reference strdup()s them): no manual memory management, copy = deep copy.
#ifndef QTIP_B200_TEMPLATE_H
#define QTIP_B200_TEMPLATE_H

#include <cstddef>
#include <string>
#include "edit_xscript.h"

// one aligned mate of a template: alignment score, read length, strand flag ('T' / 'F'), qualities
// and edit transcript (ZT:Z), both in reference-forward orientation (qtip_parse.cpp:834,842-846)
struct TemplateMate {
	int score = 0;
	int len = 0;
	char fw = 0;
	std::string qual, xscript;

	void set(int score_, int len_, char fw_, const char *qual_, const char *xscript_) {
		score = score_;
		len = len_;
		fw = fw_;
		qual.assign(qual_ ? qual_ : "");
		xscript.assign(xscript_ ? xscript_ : "");
	}
	// reference bases the mate covers (edit_xscript.h:15-25)
	size_t reflen() const { return edit_xscript_to_rflen(xscript.c_str()); }
};

struct TemplateUnpaired {
	TemplateMate m;
	char mate_flag = 0;  // '0' unpaired, '1' / '2' the aligned end of a bad-end pair
	int opp_len = 0;     // bad-end pairs: length of the end that failed to align

	TemplateUnpaired() {}
	TemplateUnpaired(int best_score, int len, char fw_flag, char mate_flag_, int opp_len_, const char *qual,
	                 const char *edit_xscript) {
		init(best_score, len, fw_flag, mate_flag_, opp_len_, qual, edit_xscript);
	}
	void init(int best_score, int len, char fw_flag, char mate_flag_, int opp_len_, const char *qual,
	          const char *edit_xscript) {
		m.set(best_score, len, fw_flag, qual, edit_xscript);
		mate_flag = mate_flag_;
		opp_len = opp_len_;
	}
	size_t reflen() const { return m.reflen(); }
	// what the simulator's window overlap must cover (input_model.h:32-36 uses reflen())
	size_t span() const { return m.reflen(); }
};

struct TemplatePaired {
	TemplateMate m1, m2;
	int score_12 = 0;        // sum of the mates' scores as given by the caller
	bool upstream1 = false;  // mate 1 is the upstream one
	size_t fraglen = 0;      // as computed by the caller (one less than the true span)

	TemplatePaired() {}
	TemplatePaired(int score_12_, int score_1, int len_1, char fw_flag_1, const char *qual_1, const char *edit_xscript_1,
	               int score_2, int len_2, char fw_flag_2, const char *qual_2, const char *edit_xscript_2, bool upstream1_,
	               size_t fraglen_) {
		init(score_12_, score_1, len_1, fw_flag_1, qual_1, edit_xscript_1, score_2, len_2, fw_flag_2, qual_2,
		     edit_xscript_2, upstream1_, fraglen_);
	}
	void init(int score_12_, int score_1, int len_1, char fw_flag_1, const char *qual_1, const char *edit_xscript_1,
	          int score_2, int len_2, char fw_flag_2, const char *qual_2, const char *edit_xscript_2, bool upstream1_,
	          size_t fraglen_) {
		m1.set(score_1, len_1, fw_flag_1, qual_1, edit_xscript_1);
		m2.set(score_2, len_2, fw_flag_2, qual_2, edit_xscript_2);
		score_12 = score_12_;
		upstream1 = upstream1_;
		fraglen = fraglen_;
	}
	size_t span() const { return fraglen; }  // input_model.h:106-110 uses fraglen_
};

#endif
