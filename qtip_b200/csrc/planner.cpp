// planner.cpp -- see planner.h.
#include "planner.h"
#include <math.h>
#include <map>
#include <utility>

namespace qt {

uint64_t apply_function(float fraction, int function, uint64_t mn, uint64_t n) {
	if (n == 0) return 0;
	double nn = (double)n;
	if (function == QTIPSIM_FUNC_SQRT) nn = (double)fraction * sqrt(nn);
	else if (function == QTIPSIM_FUNC_LINEAR) nn *= (double)fraction;
	else if (function == QTIPSIM_FUNC_CONST) nn = (double)fraction;
	const uint64_t v = (uint64_t)nn;
	return v > mn ? v : mn;
}

float window_binom_p(uint64_t retsz, uint64_t olap, uint64_t tot_fasta_len) {
	const uint64_t nchances = retsz - olap + 1;
	float p = fdiv_rn(fmul_rn((float)nchances, 1.1f), (float)tot_fasta_len);
	return 0.999f < p ? 0.999f : p;
}

namespace {

struct Stream {
	LState s;
	uint64_t n = 0;
	inline float next() {
		s = lstep(s);
		n++;
		return lfloat(lout(s));
	}
	inline void skip(uint32_t k) {
		for (uint32_t i = 0; i < k; i++) s = lstep(s);
		n += k;
	}
};

// Advances the stream over one mate's draws (simplesim.cpp:38-62).
inline void walk_mate(Stream &st, const HostGenome &g, const Category &c, const HostMateProg &hp,
                      uint64_t gpos, uint64_t gend) {
	for (uint32_t i = hp.xs_begin; i < hp.xs_end; i++) {
		const XStep &x = c.xsteps[i];
		st.skip(x.skip);
		const uint64_t gp = gpos + x.rf;
		const uint32_t ref = gp < gend ? g.code_at(gp) : 4u;
		while (lbase_code(st.next()) == ref) {
		}
	}
	st.skip(hp.tail_skip);
}

}  // namespace

int plan_replay_host(const HostGenome &g, const TemplateStore &ts, const std::vector<Window> &win,
                     const std::vector<uint32_t> &ncount, const PlanParams &pp, LState rng, Plan &out,
                     std::string &err) {
	Stream st;
	st.s = rng;
	std::map<std::pair<int32_t, uint32_t>, BinomSetup> setups;
	auto binom = [&](uint64_t n, float p) -> uint64_t {
		uint32_t pb;
		memcpy(&pb, &p, 4);
		auto key = std::make_pair((int32_t)n, pb);
		auto it = setups.find(key);
		if (it == setups.end()) it = setups.emplace(key, binom_setup((int32_t)n, p)).first;
		auto nx = [&]() { return st.next(); };
		return (uint64_t)binom_sample(it->second, nx);  // draw_binomial, simplesim.cpp:24-26
	};
	for (int c = 0; c < NCAT; c++) out.recs[c].clear();

	for (size_t wi = pp.win_lo; wi < win.size() && wi < pp.win_hi; wi++) {
		const Window &w = win[wi];
		if (!window_live(w.retsz, pp.olap, ncount[wi])) continue;
		const bool emit = true;
		const float p = window_binom_p(w.retsz, pp.olap, pp.tot_fasta_len);
		const uint64_t nslots = w.retsz - pp.olap;
		const RecDev &rec = g.recs[w.record];
		const uint64_t gwin = rec.gstart + w.refoff;
		const uint64_t gend = rec.gstart + rec.len;

		// unpaired (simplesim.cpp:242-273), bad-end (:279-334)
		for (int ci = CAT_U; ci <= CAT_B; ci++) {
			const Category &c = ts.cat[ci];
			const uint64_t ns = binom(pp.target[ci], p);
			if (ns && ci == CAT_U && nslots == 0) {
				err = "window with zero slots (assert(nslots > 0), simplesim.cpp:252)";
				return QTIPSIM_ERANGE;
			}
			for (uint64_t i = 0; i < ns; i++) {
				const uint64_t ti = lindex(st.next(), c.t.size());
				const uint64_t off = lindex(st.next(), nslots);
				if (emit) out.recs[ci].push_back(PlanRec{(uint32_t)ti, w.record, w.refoff + off, st.s.s1, st.s.s2});
				walk_mate(st, g, c, c.prog[2 * ti], gwin + off, gend);
				if (ci == CAT_B) st.skip(c.t[ti].opp_len);  // init_random, simplesim.h:108-113
			}
		}
		// concordant then discordant (simplesim.cpp:340-397)
		const uint64_t nc = binom(pp.target[CAT_C], p);
		const uint64_t nd = binom(pp.target[CAT_D], p);
		for (uint64_t i = 0; i < nc + nd; i++) {
			const int ci = i < nc ? CAT_C : CAT_D;
			const Category &c = ts.cat[ci];
			const uint64_t ti = lindex(st.next(), c.t.size());
			const uint64_t off = lindex(st.next(), nslots);
			if (emit) out.recs[ci].push_back(PlanRec{(uint32_t)ti, w.record, w.refoff + off, st.s.s1, st.s.s2});
			const TmplDesc &t = c.t[ti];
			walk_mate(st, g, c, c.prog[2 * ti], gwin + off + t.delta[0], gend);
			walk_mate(st, g, c, c.prog[2 * ti + 1], gwin + off + t.delta[1], gend);
		}
	}
	out.rng_out = st.s;
	out.draws = st.n;
	return QTIPSIM_OK;
}

}  // namespace qt
