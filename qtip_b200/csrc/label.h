// label.h -- tandem-SAM correctness labelling: does an alignment of a simulated read lie where the read's name says
// the read came from?  Restates Alignment::set_correctness (qtip_parse.cpp:519-713), the consumer of the read-name
// wire format the simulate kernels emit (SimulatedRead::write / write_pair, simplesim.cpp:102-171):
//
//     qtip!:<refid>:<+|->:<refoff>:<score>:u
//     qtip!:<refid1>:<s1>:<off1>:<score1>:<refid2>:<s2>:<off2>:<score2>:<b1|b2|c|d>
//
// plus the wgsim-style names of simulated INPUT reads (qtip_parse.cpp:634-712).  One function for host and device:
// the host build is what tests/test_host_logic.py checks against the unmodified reference; the product runs it in
// k_label_correctness (label.cu), one thread per alignment.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QT_LHD __host__ __device__ __forceinline__
#else
#define QT_LHD inline
#endif

namespace qt {

QT_LHD bool lab_isdigit(char c) { return c >= '0' && c <= '9'; }
QT_LHD bool lab_isspace(char c) { return c == ' ' || (c >= 9 && c <= 13); }
// abs((int)(a - b)) of two size_t as the reference computes it (the difference truncated to int; abs(INT_MIN) stays negative)
QT_LHD int32_t lab_absdiff(uint64_t a, uint64_t b) {
	const int32_t d = (int32_t)(uint32_t)(a - b);
	return d < 0 ? (int32_t)(0u - (uint32_t)d) : d;
}
// strncmp(p, rname, rname_len) == 0, never reading past p's terminator
QT_LHD bool lab_prefix(const char *p, const char *rname, uint32_t rname_len) {
	for (uint32_t i = 0; i < rname_len; i++) {
		if (p[i] != rname[i]) return false;
		if (p[i] == '\0') return true;  // (both ended: cannot happen for i < rname_len, kept for safety)
	}
	return true;
}

// Returns 1 (correct), 0 (incorrect) or -1 (the name is in neither simulated-read format).  qname and rname are
// NUL-terminated; flag and pos are the SAM FLAG and 1-based POS of an ALIGNED record (the reference asserts it).
QT_LHD int32_t label_correctness(const char *qname, const char *rname, int32_t flag, uint64_t pos, int32_t wiggle) {
	uint32_t rname_len = 0;
	while (rname[rname_len] != '\0') rname_len++;
	const bool fw = (flag & 16) == 0;                                           // is_fw(), qtip_parse.cpp:134-136
	const char mate = (flag & 64) ? '1' : ((flag & 128) ? '2' : '0');           // mate_flag(), :146-148
	const char pre[6] = {'q', 't', 'i', 'p', '!', '\0'};                        // sim_startswith, simplesim.h:20-24
	bool sim = true;
	for (int i = 0; i < 5 && sim; i++) sim = qname[i] == pre[i];
	if (sim) {
		const char *c = qname + 5;
		// (the reference asserts *c == ':' here; a name that fails it is not one of ours)
		if (*c != ':') return 0;
		c++;
		if (mate != '2' && !lab_prefix(c, rname, rname_len)) return 0;
		// skipping rname_len bytes must not run past the terminator
		for (uint32_t i = 0; i < rname_len; i++)
			if (c[i] == '\0') return 0;
		c += rname_len;
		if (*(c++) != ':') return 0;
		if (mate != '2' && c[0] != (fw ? '+' : '-')) return 0;
		if (*c == '\0') return 0;
		c++;
		if (*(c++) != ':') return 0;
		uint64_t refoff = 0;
		while (lab_isdigit(*c)) refoff = refoff * 10 + (uint64_t)(*c++ - '0');
		if (mate != '2' && lab_absdiff(refoff, pos - 1) >= wiggle) return 0;
		if (*(c++) != ':') return 0;
		if (*c == '-') c++;
		while (lab_isdigit(*c)) c++;  // score1 (parsed, never compared)
		if (*(c++) != ':') return 0;
		if (c[0] == 'u' && (c[1] == '\0' || lab_isspace(c[1]))) return 1;      // unpaired and correct
		if (mate != '2') return 1;                                            // paired-end, mate 1, and correct
		if (!lab_prefix(c, rname, rname_len)) return 0;
		for (uint32_t i = 0; i < rname_len; i++)
			if (c[i] == '\0') return 0;
		c += rname_len;
		if (*(c++) != ':') return 0;
		if (c[0] != (fw ? '+' : '-')) return 0;
		c++;
		if (*(c++) != ':') return 0;
		refoff = 0;
		while (lab_isdigit(*c)) refoff = refoff * 10 + (uint64_t)(*c++ - '0');
		if (lab_absdiff(refoff, pos - 1) >= wiggle) return 0;
		if (*(c++) != ':') return 0;
		if (*c == '-') c++;
		while (lab_isdigit(*c)) c++;  // score2
		if (*(c++) != ':') return 0;
		return 1;                                                             // paired-end, mate 2, and correct
	}
	// wgsim-style name of a simulated input read, e.g. 11_25006153_25006410_0:0:0_0:0:0_100_100_1_1/1
	int nund = 0, ncolon = 0;
	for (const char *q = qname; *q; q++) {
		if (*q == '_') nund++;
		else if (*q == ':' && nund >= 3) ncolon++;
	}
	if (!(nund >= 8 && ncolon == 4)) return -1;
	const char *c = qname;
	if (!lab_prefix(c, rname, rname_len)) return 0;
	for (uint32_t i = 0; i < rname_len; i++)
		if (c[i] == '\0') return 0;
	c += rname_len;
	if (*c++ != '_') return 0;
	uint64_t frag_start = 0;
	while (lab_isdigit(*c)) frag_start = frag_start * 10 + (uint64_t)(*c++ - '0');
	if (*c++ != '_') return 0;
	uint64_t frag_end = 0;
	while (lab_isdigit(*c)) frag_end = frag_end * 10 + (uint64_t)(*c++ - '0');
	if (*c++ != '_') return 0;
	while (ncolon > 0) {
		if (*c == '\0') return 0;  // (the reference would run off the string here)
		if (*c++ == ':') ncolon--;
	}
	while (*c != '\0' && lab_isdigit(*c)) c++;  // the number after the last colon ...
	if (*c == '\0') return 0;
	c++;                                        // ... and the character after it (while(isdigit(*cur++));)
	int32_t len1 = 0;
	while (lab_isdigit(*c)) len1 = len1 * 10 + (int32_t)(*c++ - '0');
	if (*c++ != '_') return 0;
	int32_t len2 = 0;
	while (lab_isdigit(*c)) len2 = len2 * 10 + (int32_t)(*c++ - '0');
	if (*c++ != '_') return 0;
	const bool flip = *c == '1';
	const bool mate1 = mate != '2';
	const int32_t len = mate1 ? len1 : len2;
	if ((!flip) == mate1) return lab_absdiff(pos, frag_start) < wiggle ? 1 : 0;                          // left end of fragment
	return lab_absdiff(pos, frag_end - (uint64_t)(int64_t)len + 1) < wiggle ? 1 : 0;                     // right end
}

}  // namespace qt
