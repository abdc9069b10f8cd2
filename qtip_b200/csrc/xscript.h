// xscript.h -- SAM alignment fields -> edit transcript, the first half of the input model's template construction
// (SURVEY 8f2).  Restates, for one alignment:
//   parse_cigar                        qtip_parse.cpp:223-253   (runs, left / right soft clip, "=/X" detection)
//   cigar_to_edit_xscript              qtip_parse.cpp:307-319   (extended CIGAR: every op repeated run times)
//   mdz_to_list                        qtip_parse.cpp:258-300   (MD:Z elements: match run, mismatch stretch, ^deletion)
//   cigar_and_mdz_to_edit_xscript      qtip_parse.cpp:326-401   (M split into = / X by MD:Z; I, D, N, S copied; H dropped)
//   lpos / rpos                        qtip_parse.cpp:164-186   (incl. the quirk that rpos skips the op after the leading S run)
// The MD:Z string is consumed as a stream (no element list is built).  One function for host and device; count-only
// when out == nullptr.  Where the reference would assert / throw, a status code is returned instead.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QT_XHD __host__ __device__ __forceinline__
#else
#define QT_XHD inline
#endif

namespace qt {

enum {
	XS_OK = 0,
	XS_BAD_CIGAR = 1,     // a run without digits / an op that is no letter or '=' (asserts, :231, :238)
	XS_BAD_OP = 2,        // M or P in an extended CIGAR (:313); P, =, X or an unknown op next to MD:Z (:390-397)
	XS_NO_TRANSCRIPT = 3, // neither an extended CIGAR nor an MD:Z field (qtip_parse.cpp:825-830)
	XS_MD_MISMATCH = 4,   // MD:Z does not fit the CIGAR (asserts :340, :347, :366-368, :399)
	XS_OVERFLOW = 5       // the transcript does not fit the space given
};

struct XsInfo {
	uint32_t len;         // ops in the transcript (without the terminator)
	int32_t left_clip, right_clip;
	uint64_t lpos, rpos;
	int32_t status;
};

QT_XHD bool xs_isdigit(char c) { return c >= '0' && c <= '9'; }
QT_XHD bool xs_isalpha(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }

// the MD:Z string as a stream of elements: op 0 = match run, 1 = mismatch stretch, 2 = deletion (mdz_to_list)
struct MdStream {
	const char *p;
	int op, run;          // current element (run > 0), or op = -1 at the end
	QT_XHD void next() {
		for (;;) {
			if (!p || *p == '\0') {
				op = -1;
				run = 0;
				return;
			}
			if (xs_isdigit(*p)) {
				int r = 0;
				do r = r * 10 + (*p++ - '0');
				while (xs_isdigit(*p));
				if (r > 0) {  // zero-length match runs never enter the reference's list
					op = 0;
					run = r;
					return;
				}
			} else if (xs_isalpha(*p)) {
				int r = 0;
				do {
					p++;
					r++;
				} while (xs_isalpha(*p));
				op = 1;
				run = r;
				return;
			} else if (*p == '^') {
				p++;
				int r = 0;
				while (xs_isalpha(*p)) {
					p++;
					r++;
				}
				op = 2;
				run = r;
				return;
			} else {
				p++;  // (the reference prints a warning and does not advance: it would loop forever; skipped here)
			}
		}
	}
};

// cigar: NUL-terminated; mdz: NUL-terminated or nullptr; pos: 1-based POS.  out (may be nullptr): receives the
// transcript and a terminating NUL if cap > len.
QT_XHD XsInfo xs_build(const char *cigar, const char *mdz, uint64_t pos, char *out, uint32_t cap) {
	XsInfo r;
	r.len = 0;
	r.left_clip = r.right_clip = 0;
	r.lpos = r.rpos = pos;
	r.status = XS_OK;
	// pass over the CIGAR: clips and whether it is extended
	bool eqx = false;
	{
		const char *c = cigar;
		bool first = true;
		while (*c) {
			if (!xs_isdigit(*c)) {
				r.status = XS_BAD_CIGAR;
				return r;
			}
			int run = 0;
			do run = run * 10 + (*c++ - '0');
			while (xs_isdigit(*c));
			if (!(xs_isalpha(*c) || *c == '=')) {
				r.status = XS_BAD_CIGAR;
				return r;
			}
			if (first && *c == 'S') r.left_clip = run;
			else if (c[1] == '\0' && *c == 'S') r.right_clip = run;
			if (*c == 'X' || *c == '=') eqx = true;
			first = false;
			c++;
		}
	}
	if (!eqx && !mdz) {
		r.status = XS_NO_TRANSCRIPT;
		return r;
	}
	auto put = [&](char op, int n) {
		if (out) {
			for (int j = 0; j < n; j++) {
				if (r.len + (uint32_t)j < cap) out[r.len + j] = op;
			}
		}
		r.len += (uint32_t)n;
	};
	MdStream md;
	md.p = mdz;
	md.op = -1;
	md.run = 0;
	if (!eqx) md.next();
	for (const char *c = cigar; *c;) {
		int run = 0;
		do run = run * 10 + (*c++ - '0');
		while (xs_isdigit(*c));
		const char cop = *c++;
		if (eqx) {
			if (cop == 'M' || cop == 'P') {
				r.status = XS_BAD_OP;
				return r;
			}
			put(cop, run);
			continue;
		}
		if (cop == 'M') {
			int left = run;
			while (left > 0 && md.op >= 0) {
				if (md.op == 2) {
					r.status = XS_MD_MISMATCH;
					return r;
				}
				const int comb = left < md.run ? left : md.run;
				if (md.op == 0) {
					put('=', comb);
				} else {
					if (md.run != comb) {
						r.status = XS_MD_MISMATCH;
						return r;
					}
					put('X', comb);
				}
				left -= comb;
				if (comb < md.run) md.run -= comb;
				else md.next();
			}
		} else if (cop == 'I') {
			put('I', run);
		} else if (cop == 'D') {
			if (md.op != 2 || md.run != run) {
				r.status = XS_MD_MISMATCH;
				return r;
			}
			md.next();
			put('D', run);
		} else if (cop == 'N') {
			put('N', run);
		} else if (cop == 'S') {
			put('S', run);
		} else if (cop == 'H') {
		} else {
			r.status = XS_BAD_OP;
			return r;
		}
	}
	if (!eqx && md.op >= 0) {
		r.status = XS_MD_MISMATCH;
		return r;
	}
	if (out) {
		if (r.len < cap) out[r.len] = '\0';
		else r.status = XS_OVERFLOW;
	}
	return r;
}

// lpos / rpos from a finished transcript (qtip_parse.cpp:164-186): rpos skips the leading S run AND the op after it
QT_XHD void xs_span(const char *xs, uint32_t len, uint64_t pos, int32_t left_clip, uint64_t &lpos, uint64_t &rpos) {
	lpos = pos - (uint64_t)(int64_t)left_clip;
	uint32_t i = 0;
	while (i < len && xs[i] == 'S') i++;
	i++;  // while(*xs++ == 'S');
	uint64_t mv = 0;
	for (; i < len; i++) mv += (xs[i] == 'S' || xs[i] == 'D' || xs[i] == 'X' || xs[i] == '=');
	rpos = pos + mv - 1;
}

}  // namespace qt
