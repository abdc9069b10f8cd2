// oracle/ref_label_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The UNMODIFIED reference's Alignment::set_correctness (qtip_parse.cpp:519-713) behind an extern "C" call, so tests
// can label alignments with the reference's own code.  Alignment is defined inside qtip_parse.cpp, next to main():
// this file (ours, no reference code in it) includes that translation unit where it lies under $(REF)/src -- with
// main renamed -- and adds the wrapper.  Built by oracle/Makefile into _ref/libreflabel.so.
#define main qtip_parse_reference_main
#include "qtip_parse.cpp"
#undef main

#include <stdint.h>
#include <string.h>
#include <string>

extern "C" int ref_set_correctness(const char *qname, const char *rname, int flag, uint64_t pos, int wig) {
	Alignment al;
	std::string q(qname), r(rname);  // the reference takes mutable char*
	al.qname = &q[0];
	al.rname = &r[0];
	al.flag = flag;
	al.pos = (size_t)pos;
	al.set_correctness(wig);
	return al.correct;
}

// The reference's own transcript construction for one alignment: parse_cigar (which builds the transcript of an
// extended CIGAR), mdz_to_list + cigar_and_mdz_to_edit_xscript for an M-style one, then lpos() / rpos().  The inputs
// must be well-formed: the reference asserts (and this build keeps asserts live, like src/Makefile:26).
extern "C" int ref_edit_xscript(const char *cigar, const char *mdz, uint64_t pos, char *out, uint64_t cap, int *left_clip,
                                int *right_clip, uint64_t *lpos, uint64_t *rpos) {
	Alignment al;
	std::string c(cigar), m(mdz ? mdz : "");
	al.cigar = &c[0];
	al.pos = (size_t)pos;
	al.parse_cigar();
	if (mdz) {
		al.mdz = &m[0];
		al.mdz_to_list();
		if (!al.cigar_equal_x) al.cigar_and_mdz_to_edit_xscript();
	}
	if (al.edit_xscript.empty()) return 3;
	const size_t n = strlen(al.edit_xscript.ptr());
	if (n + 1 > cap) return 5;
	memcpy(out, al.edit_xscript.ptr(), n + 1);
	*left_clip = al.left_clip;
	*right_clip = al.right_clip;
	*lpos = al.lpos();
	*rpos = al.rpos();
	return 0;
}
