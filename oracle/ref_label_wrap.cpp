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
