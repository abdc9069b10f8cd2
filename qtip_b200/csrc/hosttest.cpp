// hosttest.cpp -- CPU-only test hooks for the HOST logic of libqtipsim (FASTA parse/pack,
// window table, template flattening, replay chain walk).  Built into
// qtip_b200/libqtipsim_hosttest.so for `pytest -m "not gpu"`; not part of the product
// library and never produces reads.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/qtipsim.h"
#include "genome.h"
#include "label.h"
#include "lecuyer.h"
#include "pack_swar.h"
#include "planner.h"
#include "xscript.h"
#include "templates.h"
#include "model.h"

using namespace qt;

struct HostTest {
	HostGenome g;
	TemplateStore ts;
	std::vector<Window> win;
	std::vector<uint32_t> ncount;
	Plan plan;
	ModelBuilder model;
	std::string err;
	std::string scratch;
};

extern "C" {

HostTest *qthost_create() { return new HostTest(); }
void qthost_destroy(HostTest *h) { delete h; }
const char *qthost_error(HostTest *h) { return h->err.c_str(); }

int qthost_parse(HostTest *h, const void *text, uint64_t n) {
	h->g.parse((const uint8_t *)text, n);
	return 0;
}
uint64_t qthost_nrecords(HostTest *h) { return h->g.recs.size(); }
uint64_t qthost_record(HostTest *h, uint64_t i, char *name, uint64_t cap, uint64_t *len, uint64_t *gstart) {
	const RecDev &r = h->g.recs[i];
	uint64_t n = r.name_len < cap - 1 ? r.name_len : cap - 1;
	memcpy(name, h->g.names.data() + r.name_off, n);
	name[n] = 0;
	*len = r.len;
	*gstart = r.gstart;
	return r.name_len;
}
// unpack bases [g0, g0+n) of the concatenated space to ASCII
void qthost_unpack(HostTest *h, uint64_t g0, uint64_t n, char *dst) {
	for (uint64_t i = 0; i < n; i++) dst[i] = "ACGTN"[h->g.code_at(g0 + i)];
}
int qthost_set_unpaired(HostTest *h, int cat, const qtipsim_tmpl_unp *v, size_t kept, uint64_t added) {
	return h->ts.set_unpaired(cat, v, kept, added, h->err);
}
int qthost_set_paired(HostTest *h, int cat, const qtipsim_tmpl_pair *v, size_t kept, uint64_t added) {
	return h->ts.set_paired(cat, v, kept, added, h->err);
}
uint64_t qthost_olap(HostTest *h) { return h->ts.olap(); }

// the input model's reservoirs (model.h): the same three steps as qtipsim_model_begin / _add / _commit
int qthost_model_begin(HostTest *h, const uint64_t *k) {
	h->model.begin(k);
	return 0;
}
int qthost_model_add(HostTest *h, const qtipsim_model_batch *b, int32_t *cg1, int32_t *cg2) {
	LState rng{(uint32_t)*cg1, (uint32_t)*cg2};
	const int rc = h->model.add(*b, rng, h->err);
	*cg1 = (int32_t)rng.s1;
	*cg2 = (int32_t)rng.s2;
	return rc;
}
uint64_t qthost_model_kept(HostTest *h, int cat) { return h->model.kept[cat].size(); }
uint64_t qthost_model_offered(HostTest *h, int cat) { return h->model.n[cat]; }
// kept record i of a category as one line of text (for comparison with the test's restatement)
const char *qthost_model_record(HostTest *h, int cat, uint64_t i) {
	const ModelBuilder::Rec &r = h->model.kept[cat][i];
	std::string &o = h->scratch;
	o.clear();
	for (int m = 0; m < 2; m++) {
		o += std::to_string(r.m[m].score) + "," + std::to_string(r.m[m].len) + "," + r.m[m].fw + "," + r.m[m].qual + "," + r.m[m].xs + "|";
	}
	o += std::string(1, r.mate_flag) + "," + std::to_string(r.opp_len) + "," + std::to_string(r.upstream1) + "," + std::to_string(r.fraglen);
	return o.c_str();
}
int qthost_model_commit(HostTest *h) { return h->model.commit(h->ts, h->err); }
int qthost_cat_valid(HostTest *h, int cat) { return h->ts.cat[cat].valid ? 1 : 0; }

uint64_t qthost_windows(HostTest *h, uint64_t chunksz, uint64_t olap) {
	h->win = h->g.windows(chunksz, olap);
	h->ncount.assign(h->win.size(), 0);
	for (size_t i = 0; i < h->win.size(); i++) {
		const Window &w = h->win[i];
		const uint64_t g0 = h->g.recs[w.record].gstart + w.refoff;
		uint32_t c = 0;
		for (uint64_t j = 0; j < w.retsz; j++) c += h->g.code_at(g0 + j) == 4;
		h->ncount[i] = c;
	}
	return h->win.size();
}
void qthost_window(HostTest *h, uint64_t i, uint32_t *record, uint64_t *refoff, uint64_t *retsz, uint64_t *ncount) {
	*record = h->win[i].record;
	*refoff = h->win[i].refoff;
	*retsz = h->win[i].retsz;
	*ncount = h->ncount[i];
}

int qthost_plan(HostTest *h, uint64_t chunksz, float fraction, int function, uint64_t min_u, uint64_t min_c,
                uint64_t min_d, uint64_t min_b, int32_t *cg1, int32_t *cg2, uint64_t *draws, uint64_t nreads[4],
                uint64_t win_lo, uint64_t win_hi) {
	PlanParams pp;
	pp.chunksz = chunksz;
	pp.olap = h->ts.olap();
	pp.tot_fasta_len = h->g.file_bytes;
	pp.target[CAT_U] = apply_function(fraction, function, min_u, h->ts.cat[CAT_U].added);
	pp.target[CAT_B] = apply_function(fraction, function, min_b, h->ts.cat[CAT_B].added);
	pp.target[CAT_C] = apply_function(fraction, function, min_c, h->ts.cat[CAT_C].added);
	pp.target[CAT_D] = apply_function(fraction, function, min_d, h->ts.cat[CAT_D].added);
	qthost_windows(h, chunksz, pp.olap);
	pp.win_lo = win_lo;
	pp.win_hi = win_hi < h->win.size() ? win_hi : h->win.size();
	int rc = plan_replay_host(h->g, h->ts, h->win, h->ncount, pp, LState{(uint32_t)*cg1, (uint32_t)*cg2}, h->plan, h->err);
	if (rc) return rc;
	*cg1 = (int32_t)h->plan.rng_out.s1;
	*cg2 = (int32_t)h->plan.rng_out.s2;
	*draws = h->plan.draws;
	for (int c = 0; c < NCAT; c++) nreads[c] = h->plan.recs[c].size();
	return 0;
}
void qthost_plan_rec(HostTest *h, int cat, uint64_t i, uint32_t *tmpl, uint32_t *record, uint64_t *refoff,
                     uint32_t *s1, uint32_t *s2) {
	const PlanRec &p = h->plan.recs[cat][i];
	*tmpl = p.tmpl;
	*record = p.record;
	*refoff = p.refoff;
	*s1 = p.s1;
	*s2 = p.s2;
}

// RNG known-answer hooks
void qthost_rng_step(uint32_t *s1, uint32_t *s2, int32_t *i4, float *r4) {
	LState s = lstep(LState{*s1, *s2});
	*s1 = s.s1;
	*s2 = s.s2;
	*i4 = lout(s);
	*r4 = lfloat(lout(s));
}
void qthost_rng_jump(uint32_t *s1, uint32_t *s2, uint64_t k) {
	LState s = ljump(LState{*s1, *s2}, k);
	*s1 = s.s1;
	*s2 = s.s2;
}
int qthost_ignbin(uint32_t *s1, uint32_t *s2, int n, float p) {
	LState s{*s1, *s2};
	BinomSetup b = binom_setup(n, p);
	auto nx = [&]() {
		s = lstep(s);
		return lfloat(lout(s));
	};
	int v = binom_sample(b, nx);
	*s1 = s.s1;
	*s2 = s.s2;
	return v;
}
// the table-driven sampler the device resolver uses (lecuyer.h: binom_tables / binom_sample_tab)
int qthost_ignbin_tab(uint32_t *s1, uint32_t *s2, int n, float p) {
	LState s{*s1, *s2};
	BinomSetup b = binom_setup(n, p);
	static BinomTables t;
	binom_tables(b, t);
	auto nx = [&]() {
		s = lstep(s);
		return lfloat(lout(s));
	};
	int v = binom_sample_tab(b, t, nx);
	*s1 = s.s1;
	*s2 = s.s2;
	return v;
}
// lbase_code_z against lbase_code(lfloat(z)): re-derives the three thresholds by bisection, then compares the two on
// +-span around each threshold and on a stride over the whole range; returns the number of disagreements
uint64_t qthost_base_thresholds(int32_t *t, int32_t span, int32_t stride) {
	for (int c = 1; c <= 3; c++) {
		int64_t lo = 1, hi = 2147483562;
		while (lo < hi) {
			const int64_t m = (lo + hi) / 2;
			if ((int)lbase_code(lfloat((int32_t)m)) >= c) hi = m;
			else lo = m + 1;
		}
		t[c - 1] = (int32_t)lo;
	}
	uint64_t bad = 0;
	for (int c = 0; c < 3; c++)
		for (int64_t z = (int64_t)t[c] - span; z <= (int64_t)t[c] + span; z++)
			if (z >= 1 && z <= 2147483562 && lbase_code_z((int32_t)z) != lbase_code(lfloat((int32_t)z))) bad++;
	for (int64_t z = 1; z <= 2147483562; z += stride)
		if (lbase_code_z((int32_t)z) != lbase_code(lfloat((int32_t)z))) bad++;
	return bad;
}
// the product's labelling function (label.h), host build
int qthost_label(const char *qname, const char *rname, int flag, uint64_t pos, int wiggle) {
	return label_correctness(qname, rname, flag, pos, wiggle);
}
// the product's transcript builder (xscript.h), host build
int qthost_edit_xscript(const char *cigar, const char *mdz, uint64_t pos, char *out, uint32_t cap, int *lclip, int *rclip, uint64_t *lpos,
                        uint64_t *rpos) {
	const XsInfo r = xs_build(cigar, mdz, pos, out, cap);
	*lclip = r.left_clip;
	*rclip = r.right_clip;
	if (r.status == XS_OK) xs_span(out, r.len, pos, r.left_clip, *lpos, *rpos);
	return r.status;
}
float qthost_logf(float x) { return logf_glibc(x); }
// compares logf_glibc with the host libm over [lo, hi) bit patterns; returns mismatches
uint64_t qthost_logf_sweep(uint32_t lo, uint32_t hi, uint32_t step) {
	uint64_t bad = 0;
	for (uint64_t u = lo; u < hi; u += step) {
		uint32_t b = (uint32_t)u;
		float x, a, g;
		memcpy(&x, &b, 4);
		a = logf_glibc(x);
		g = logf(x);
		bad += memcmp(&a, &g, 4) != 0;
	}
	return bad;
}


// ---- K0's word-parallel classification (pack_swar.h), checked byte by byte by tests/test_host_logic.py ----

// bytes [a, b) of one 16-byte vector: squeezed codes / N bits, returns the count of non-whitespace bytes
uint32_t qthost_swar_vec(const uint8_t *bytes16, uint32_t a, uint32_t b, uint32_t *codes, uint32_t *nbits, uint32_t *ws) {
	uint32_t w[4];
	memcpy(w, bytes16, 16);
	VecClass vc = classify_vec(w[0], w[1], w[2], w[3]);
	*ws = vc.ws;
	const uint32_t cnt = squeeze_vec(vc, a, b);
	*codes = vc.codes;
	*nbits = vc.nbits;
	return cnt;
}

// The arithmetic of pack.cu's k_pack_units for one unit, lanes run one after the other: text [t0, t1) (text is
// 16-byte aligned and padded), first base at genome position g; ORs into bases[] / nmask[] like the kernel's
// image + ship steps.  Returns the number of bases.
uint64_t qthost_swar_pack_unit(const uint8_t *text, uint64_t t0, uint64_t t1, uint64_t g, uint32_t *bases, uint32_t *nmask) {
	const uint64_t g_first = g;
	for (uint64_t tile = t0 & ~(uint64_t)15; tile < t1; tile += 512) {
		uint32_t img_b[36] = {0}, img_n[20] = {0};
		uint32_t excl = 0;
		const uint32_t gph = (uint32_t)(g & 15u), gpn = (uint32_t)(g & 31u);
		for (uint32_t lane = 0; lane < 32; lane++) {
			const uint64_t i = tile + lane * 16;
			if (i >= t1) continue;
			const uint32_t a = t0 > i ? (uint32_t)(t0 - i) : 0u;
			const uint32_t b = t1 - i < 16u ? (uint32_t)(t1 - i) : 16u;
			uint32_t w[4];
			memcpy(w, text + i, 16);
			VecClass vc = classify_vec(w[0], w[1], w[2], w[3]);
			const uint32_t cnt = squeeze_vec(vc, a, b);
			if (vc.codes) {
				const uint32_t ob = 2u * (gph + excl);
				const uint64_t x = (uint64_t)vc.codes << (ob & 31u);
				img_b[ob >> 5] |= (uint32_t)x;
				if ((uint32_t)(x >> 32)) img_b[(ob >> 5) + 1] |= (uint32_t)(x >> 32);
			}
			if (vc.nbits) {
				const uint32_t on = gpn + excl;
				const uint64_t x = (uint64_t)vc.nbits << (on & 31u);
				img_n[on >> 5] |= (uint32_t)x;
				if ((uint32_t)(x >> 32)) img_n[(on >> 5) + 1] |= (uint32_t)(x >> 32);
			}
			excl += cnt;
		}
		const uint32_t total = excl;
		const uint32_t nwords = (gph + total + 15u) >> 4, nmw = (gpn + total + 31u) >> 5;
		for (uint32_t wl = 0; wl < nwords; wl++) {
			if (wl != 0 && wl + 1 != nwords) bases[(g >> 4) + wl] = img_b[wl];  // interior word: plain store
			else bases[(g >> 4) + wl] |= img_b[wl];
		}
		for (uint32_t wl = 0; wl < nmw; wl++) nmask[(g >> 5) + wl] |= img_n[wl];
		for (uint32_t wl = nwords; wl < 36; wl++)
			if (img_b[wl]) return ~0ull;  // bits beyond the round's words: would be lost by the kernel
		for (uint32_t wl = nmw; wl < 20; wl++)
			if (img_n[wl]) return ~0ull;
		g += total;
	}
	return g - g_first;
}

}  // extern "C"
