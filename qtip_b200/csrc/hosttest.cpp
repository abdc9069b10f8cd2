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
#include "lecuyer.h"
#include "planner.h"
#include "templates.h"

using namespace qt;

struct HostTest {
	HostGenome g;
	TemplateStore ts;
	std::vector<Window> win;
	std::vector<uint32_t> ncount;
	Plan plan;
	std::string err;
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

}  // extern "C"
