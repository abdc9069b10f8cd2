// templates.cpp -- see templates.h.
#include "templates.h"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>

namespace qt {

namespace {

// Flattens one (qual, transcript) into a MateDesc and its host draw program.
// Returns false (with why) when the reference would throw / assert on this template:
// unknown op (simplesim.cpp:58-60) or produced length != strlen(qual) (:63).
bool add_mate(Category &c, MateDesc &m, HostMateProg &hp, const char *qual, const char *xs,
              bool fw, int32_t score, std::string &why) {
	memset(&m, 0, sizeof(m));
	if (qual == NULL || xs == NULL) {
		why = "NULL qual or transcript";
		return false;
	}
	m.qual_off = (uint32_t)c.qual.size();
	m.xs_off = (uint32_t)c.xs.size();
	m.len = (uint32_t)strlen(qual);
	m.xs_len = (uint32_t)strlen(xs);
	m.fw = fw ? 1 : 0;
	m.score_len = (uint8_t)snprintf(m.score, sizeof(m.score), "%d", score);
	c.qual.append(qual, m.len);
	c.xs.append(xs, m.xs_len);
	hp.xs_begin = (uint32_t)c.xsteps.size();
	uint32_t rf = 0, rd = 0, skip = 0, ndraw = 0;
	bool ok = true;
	for (uint32_t i = 0; i < m.xs_len; i++) {
		switch (xs[i]) {
		case '=': rf++; rd++; break;
		case 'X': c.xsteps.push_back(XStep{skip, rf}); skip = 0; rf++; rd++; ndraw++; break;
		case 'I': rd++; skip++; ndraw++; break;
		case 'D': rf++; break;
		case 'S': rf++; rd++; skip++; ndraw++; break;
		default:
			if (ok) {
				char b[96];
				snprintf(b, sizeof(b), "Unknown operation in edit transcript: '%c'", xs[i]);
				why = b;
			}
			ok = false;
		}
	}
	hp.xs_end = (uint32_t)c.xsteps.size();
	hp.tail_skip = skip;
	m.rflen = rf;
	m.ndraw = ndraw;
	c.max_body = std::max(c.max_body, m.len);
	c.min_body = std::min(c.min_body, m.len);
	c.max_rflen = std::max(c.max_rflen, rf);
	c.max_score_len = std::max<uint32_t>(c.max_score_len, m.score_len);

	// ---- compiled, output-oriented form for the tiled kernel ----
	// qualities as they are printed (simplesim.cpp:119-125): reversed for '-' mates
	if (c.qual_out.empty()) c.qual_out.assign(16, '\0');
	m.qual_out_off = (uint32_t)c.qual_out.size();
	{
		const size_t at = c.qual_out.size();
		c.qual_out.resize(at + ((size_t)m.len + 15) / 16 * 16, '\0');
		if (fw) memcpy(&c.qual_out[at], qual, m.len);
		else std::reverse_copy(qual, qual + m.len, &c.qual_out[at]);
	}
	while (c.ops.size() % 4) c.ops.push_back(0u);  // 16-byte aligned lists: the kernel reads four ops per load
	m.ops_off = (uint32_t)c.ops.size();
	m.seg_off = (uint32_t)c.segs.size();
	if (ok && rd == m.len && m.len <= COMPILED_MAX_LEN && rf <= COMPILED_MAX_LEN) {
		const uint32_t L = m.len;
		uint32_t r_d = 0, r_f = 0;
		bool in_run = false;
		uint32_t run_rd = 0, run_rf = 0;
		std::vector<SegDev> segs;
		auto close_run = [&]() {
			if (!in_run) return;
			in_run = false;
			SegDev sg;
			if (fw) {
				sg.out_start = (uint16_t)run_rd;
				sg.out_end = (uint16_t)r_d;
				sg.ref0 = (int32_t)run_rf;
			} else {
				sg.out_start = (uint16_t)(L - r_d);
				sg.out_end = (uint16_t)(L - run_rd);
				sg.ref0 = (int32_t)(r_f - 1);
			}
			segs.push_back(sg);
		};
		for (uint32_t i = 0; i < m.xs_len; i++) {
			const char op = xs[i];
			const uint32_t outp = fw ? r_d : L - 1 - r_d;
			if (op == '=' || op == 'X' || op == 'S') {
				if (!in_run) {
					in_run = true;
					run_rd = r_d;
					run_rf = r_f;
				}
				if (op == 'X') c.ops.push_back(OP_X | (outp << 2) | (r_f << 17));
				if (op == 'S') c.ops.push_back(OP_IS | (outp << 2));
				r_d++;
				r_f++;
			} else if (op == 'I') {
				close_run();
				c.ops.push_back(OP_IS | (outp << 2));
				r_d++;
			} else {  // D
				close_run();
				r_f++;
			}
		}
		close_run();
		if (!fw) std::reverse(segs.begin(), segs.end());
		c.segs.insert(c.segs.end(), segs.begin(), segs.end());
		if (c.segs.size() - m.seg_off > 65535 || c.ops.size() - m.ops_off > 65535) c.compiled_ok = false;
		m.nseg = (uint16_t)(c.segs.size() - m.seg_off);
		m.nops = (uint16_t)(c.ops.size() - m.ops_off);
		if (segs.size() == 1 && segs[0].out_start == 0 && segs[0].out_end == L) {
			m.single = 1;
			m.ref0 = segs[0].ref0;
		}
	} else {
		c.compiled_ok = false;
	}
	if (ok && rd != m.len) {
		why = "edit transcript produces a read of different length than the quality string";
		ok = false;
	}
	return ok;
}

}  // namespace

void Category::finish_xprog() {
	xprog.assign(t.size(), XProgRec{});
	for (size_t i = 0; i < t.size(); i++) {
		for (int m = 0; m < 2; m++) {
			XProgHalf &hf = xprog[i].h[m];
			const HostMateProg &hp = prog[2 * i + m];
			const uint32_t nx = hp.xs_end - hp.xs_begin;
			bool fits = nx <= 7 && hp.tail_skip <= 255;
			for (uint32_t k = 0; fits && k < nx; k++) fits = xsteps[hp.xs_begin + k].rf <= 65535u && xsteps[hp.xs_begin + k].skip <= 255u;
			if (!fits) {
				hf.nx = 0xff;
			} else {
				hf.nx = (uint8_t)nx;
				hf.tail = (uint8_t)hp.tail_skip;
				for (uint32_t k = 0; k < nx; k++) {
					hf.rf[k] = (uint16_t)xsteps[hp.xs_begin + k].rf;
					hf.skip[k] = (uint8_t)xsteps[hp.xs_begin + k].skip;
				}
			}
		}
		xprog[i].h[0].extra0 = t[i].opp_len;
		xprog[i].h[0].extra1 = t[i].delta[0];
		xprog[i].h[1].extra0 = t[i].delta[1];
	}
}

namespace {

// templates [0, n) of v into a fresh category (arena offsets from zero; sums instead of means)
void fill_unpaired(Category &c, int ci, const qtipsim_tmpl_unp *v, size_t n) {
	c.t.resize(n);
	c.prog.resize(2 * n);
	if (n) {  // arenas sized from the first template: one allocation instead of a doubling chain
		const size_t l0 = v[0].qual ? strlen(v[0].qual) : 0;
		c.qual.reserve(n * (l0 + 8));
		c.xs.reserve(n * (l0 + 8));
		c.qual_out.reserve(n * (l0 + 24) + 64);
	}
	auto draw_stats = [&](double fixed, double nx) {  // per template: mean fixed + 4/3 nx, variance 4/9 nx
		const double m = fixed + nx * (4.0 / 3.0);
		c.sum_dm += m;
		c.sum_dsq += m * m + nx * (4.0 / 9.0);
		c.draw_max_fixed = std::max<uint32_t>(c.draw_max_fixed, (uint32_t)(fixed + nx));
	};
	for (size_t i = 0; i < n; i++) {
		TmplDesc &t = c.t[i];
		memset(&t, 0, sizeof(t));
		std::string why;
		if (!add_mate(c, t.m[0], c.prog[2 * i], v[i].qual, v[i].xscript, v[i].fw_flag == 'T', v[i].best_score, why) && c.valid) {
			c.valid = false;
			c.why = why;
		}
		t.nmates = 1;
		const double nx = c.prog[2 * i].xs_end - c.prog[2 * i].xs_begin;
		if (ci == CAT_B) {
			t.aligned_is_mate1 = (v[i].mate_flag == '1');
			t.opp_len = v[i].opp_len > 0 ? (uint32_t)v[i].opp_len : 0;
			c.max_body = std::max(c.max_body, t.opp_len);
			c.min_body = std::min(c.min_body, t.opp_len);
			if (t.opp_len > COMPILED_MAX_LEN) c.compiled_ok = false;
			if (v[i].opp_len <= 0 && c.valid) {  // assert(len > 0), simplesim.h:96
				c.valid = false;
				c.why = "bad-end template with opp_len <= 0";
			}
			t.typ[0] = 'b';
			t.typ[1] = t.aligned_is_mate1 ? '1' : '2';
			t.typ_len = 2;
			// name (both scores) on both files + two bodies
			c.sum_bytes += 2.0 * (7 + 2 * (5 + t.m[0].score_len) + 2 + 1) + 2.0 * t.m[0].len + 2.0 * t.opp_len + 8;
			c.sum_draws += 2 + t.m[0].ndraw + t.opp_len;
			draw_stats(2.0 + t.m[0].ndraw - nx + t.opp_len, nx);
		} else {
			t.typ[0] = 'u';
			t.typ_len = 1;
			c.sum_bytes += 7 + 5 + t.m[0].score_len + 1 + 1 + 2.0 * t.m[0].len + 4;
			c.sum_draws += 2 + t.m[0].ndraw;
			draw_stats(2.0 + t.m[0].ndraw - nx, nx);
		}
		c.max_len = std::max<uint64_t>(c.max_len, t.m[0].rflen);  // input_model.h:38-41
	}
}

void fill_paired(Category &c, int ci, const qtipsim_tmpl_pair *v, size_t n) {
	c.t.resize(n);
	c.prog.resize(2 * n);
	if (n) {
		const size_t l0 = (v[0].qual_1 ? strlen(v[0].qual_1) : 0) + (v[0].qual_2 ? strlen(v[0].qual_2) : 0);
		c.qual.reserve(n * (l0 + 16));
		c.xs.reserve(n * (l0 + 16));
		c.qual_out.reserve(n * (l0 + 48) + 64);
	}
	for (size_t i = 0; i < n; i++) {
		TmplDesc &t = c.t[i];
		memset(&t, 0, sizeof(t));
		std::string why;
		bool ok = add_mate(c, t.m[0], c.prog[2 * i], v[i].qual_1, v[i].xscript_1, v[i].fw_flag_1 == 'T', v[i].score_1, why);
		ok = ok && add_mate(c, t.m[1], c.prog[2 * i + 1], v[i].qual_2, v[i].xscript_2, v[i].fw_flag_2 == 'T', v[i].score_2, why);
		if (!ok && c.valid) {
			c.valid = false;
			c.why = why;
		}
		t.nmates = 2;
		t.typ[0] = ci == CAT_C ? 'c' : 'd';
		t.typ_len = 1;
		// simplesim.cpp:354-363
		const uint64_t fl = v[i].fraglen;
		if (v[i].upstream1) {
			t.delta[0] = 0;
			t.delta[1] = (uint32_t)(std::max<uint64_t>(fl, t.m[1].rflen) - t.m[1].rflen);
		} else {
			t.delta[1] = 0;
			t.delta[0] = (uint32_t)(std::max<uint64_t>(fl, t.m[0].rflen) - t.m[0].rflen);
		}
		c.max_len = std::max<uint64_t>(c.max_len, fl);  // input_model.h:112-115
		c.sum_bytes += 2.0 * (7 + 9 + t.m[0].score_len + t.m[1].score_len + 2) + 2.0 * t.m[0].len + 2.0 * t.m[1].len + 8;
		c.sum_draws += 2 + t.m[0].ndraw + t.m[1].ndraw;
		const double nx = (double)(c.prog[2 * i].xs_end - c.prog[2 * i].xs_begin) + (double)(c.prog[2 * i + 1].xs_end - c.prog[2 * i + 1].xs_begin);
		const double fixed = 2.0 + t.m[0].ndraw + t.m[1].ndraw - nx;
		const double m = fixed + nx * (4.0 / 3.0);
		c.sum_dm += m;
		c.sum_dsq += m * m + nx * (4.0 / 9.0);
		c.draw_max_fixed = std::max<uint32_t>(c.draw_max_fixed, (uint32_t)(fixed + nx));
	}
}

// appends part p (built from offset zero) to c: every arena offset of p's descriptors moves by the arena's size so far
void append_part(Category &c, Category &p) {
	while (c.ops.size() % 4) c.ops.push_back(0u);  // draw-op lists stay 16-byte aligned
	const uint32_t bq = (uint32_t)c.qual.size(), bx = (uint32_t)c.xs.size(), bo = (uint32_t)c.qual_out.size();
	const uint32_t bops = (uint32_t)c.ops.size(), bseg = (uint32_t)c.segs.size(), bxs = (uint32_t)c.xsteps.size();
	for (TmplDesc &t : p.t)
		for (int m = 0; m < 2; m++) {
			MateDesc &md = t.m[m];
			if (m >= t.nmates) continue;
			md.qual_off += bq;
			md.xs_off += bx;
			md.qual_out_off += bo;
			md.ops_off += bops;
			md.seg_off += bseg;
		}
	for (HostMateProg &hp : p.prog) {
		hp.xs_begin += bxs;
		hp.xs_end += bxs;
	}
	c.t.insert(c.t.end(), p.t.begin(), p.t.end());
	c.prog.insert(c.prog.end(), p.prog.begin(), p.prog.end());
	c.qual += p.qual;
	c.xs += p.xs;
	c.qual_out += p.qual_out;
	c.ops.insert(c.ops.end(), p.ops.begin(), p.ops.end());
	c.segs.insert(c.segs.end(), p.segs.begin(), p.segs.end());
	c.xsteps.insert(c.xsteps.end(), p.xsteps.begin(), p.xsteps.end());
	c.compiled_ok = c.compiled_ok && p.compiled_ok;
	c.max_body = std::max(c.max_body, p.max_body);
	c.min_body = std::min(c.min_body, p.min_body);
	c.max_score_len = std::max(c.max_score_len, p.max_score_len);
	c.max_rflen = std::max(c.max_rflen, p.max_rflen);
	c.max_len = std::max(c.max_len, p.max_len);
	c.draw_max_fixed = std::max(c.draw_max_fixed, p.draw_max_fixed);
	if (c.valid && !p.valid) {  // parts are appended in template order: the first problem found stays the reported one
		c.valid = false;
		c.why = p.why;
	}
	c.sum_bytes += p.sum_bytes;
	c.sum_draws += p.sum_draws;
	c.sum_dm += p.sum_dm;
	c.sum_dsq += p.sum_dsq;
}

// The per-template work (transcript scans, output-oriented quality copies, segment / draw-op compilation) is
// independent between templates: large models are compiled in slices by a few threads and the slices concatenated.
template <class V, class Fill>
void build_category(Category &c, int ci, const V *v, size_t kept, uint64_t added, Fill fill) {
	c.clear();
	c.added = added;
	unsigned nt = 1;
	if (kept >= 4096) {
		const char *e = getenv("QTIPSIM_TEMPLATE_THREADS");
		nt = e && *e ? (unsigned)atoi(e) : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		if (nt < 1) nt = 1;
	}
	if (nt == 1) {
		fill(c, ci, v, kept);
	} else {
		std::vector<Category> parts(nt);
		std::vector<std::thread> th;
		for (unsigned k = 0; k < nt; k++) {
			const size_t lo = kept * k / nt, hi = kept * (k + 1) / nt;
			th.emplace_back([&, k, lo, hi]() { fill(parts[k], ci, v + lo, hi - lo); });
		}
		for (auto &t : th) t.join();
		size_t q = 0, x = 0, o = 0;
		for (auto &p : parts) q += p.qual.size(), x += p.xs.size(), o += p.qual_out.size();
		c.qual.reserve(q);
		c.xs.reserve(x);
		c.qual_out.reserve(o + 64);
		c.t.reserve(kept);
		c.prog.reserve(2 * kept);
		c.added = added;
		for (auto &p : parts) append_part(c, p);
	}
	// descriptor offsets into the arenas are 32-bit: a model whose arenas outgrow that is refused, not wrapped
	if (c.qual.size() > 0xfffffff0ull || c.xs.size() > 0xfffffff0ull || c.qual_out.size() > 0xfffffff0ull || c.ops.size() > 0xfffffff0ull ||
	    c.segs.size() > 0xfffffff0ull || c.xsteps.size() > 0xfffffff0ull) {
		c.valid = false;
		c.arena_overflow = true;
		c.why = "input model too large: a template arena exceeds 4 GiB";
	}
	if (kept) {
		c.mean_bytes = c.sum_bytes / kept;
		c.mean_draws = c.sum_draws / kept;
		c.draw_m1 = c.sum_dm / kept;
		c.draw_v1 = std::max(0.0, c.sum_dsq / kept - c.draw_m1 * c.draw_m1);
	}
	c.finish_xprog();
}

}  // namespace

int TemplateStore::set_unpaired(int ci, const qtipsim_tmpl_unp *v, size_t kept, uint64_t added, std::string &err) {
	if (ci != CAT_U && ci != CAT_B) {
		err = "set_templates_unpaired: category must be U or B";
		return QTIPSIM_EINVAL;
	}
	build_category(cat[ci], ci, v, kept, added, fill_unpaired);
	if (cat[ci].arena_overflow) {
		err = cat[ci].why;
		return QTIPSIM_ERANGE;
	}
	return QTIPSIM_OK;
}

int TemplateStore::set_paired(int ci, const qtipsim_tmpl_pair *v, size_t kept, uint64_t added, std::string &err) {
	if (ci != CAT_C && ci != CAT_D) {
		err = "set_templates_paired: category must be C or D";
		return QTIPSIM_EINVAL;
	}
	build_category(cat[ci], ci, v, kept, added, fill_paired);
	if (cat[ci].arena_overflow) {
		err = cat[ci].why;
		return QTIPSIM_ERANGE;
	}
	return QTIPSIM_OK;
}

uint64_t TemplateStore::olap() const {
	uint64_t o = 0;
	for (int i = 0; i < NCAT; i++) o = std::max(o, cat[i].max_len);
	return o;
}

}  // namespace qt
