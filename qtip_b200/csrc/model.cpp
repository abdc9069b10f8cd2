// model.cpp -- see model.h
#include "model.h"
#include <string.h>

namespace qt {

void ModelBuilder::begin(const uint64_t cap[NCAT]) {
	for (int c = 0; c < NCAT; c++) {
		k[c] = cap[c];
		n[c] = 0;
		kept[c].clear();
	}
	open = true;
}

int ModelBuilder::add(const qtipsim_model_batch &b, LState &rng, std::string &err) {
	if (!open) {
		err = "qtipsim_model_add before qtipsim_model_begin";
		return QTIPSIM_ESTATE;
	}
	if (b.n && (!b.cat || !b.a1 || !b.best_score || !b.len || !b.fw || !b.quals || !b.q_off || !b.xscripts || !b.x_off)) {
		err = "qtipsim_model_add: missing array";
		return QTIPSIM_EINVAL;
	}
	auto mate = [&](uint64_t a) {
		Mate m;
		m.score = b.best_score[a];
		m.len = b.len[a];
		m.fw = b.fw[a] ? 'T' : 'F';
		m.qual = b.quals + b.q_off[a];
		m.xs = b.xscripts + b.x_off[a];
		return m;
	};
	for (uint64_t i = 0; i < b.n; i++) {
		const int c = b.cat[i];
		if (c < 0 || c >= NCAT) {
			err = "qtipsim_model_add: bad category";
			return QTIPSIM_EINVAL;
		}
		const bool pair = c == CAT_C || c == CAT_D;
		if (pair ? (!b.a2 || !b.upstream1 || !b.fraglen) : (!b.mate_flag || !b.opp_len)) {
			err = "qtipsim_model_add: missing array for this category";
			return QTIPSIM_EINVAL;
		}
		if (b.a1[i] >= b.n_aln || (pair && b.a2[i] >= b.n_aln)) {
			err = "qtipsim_model_add: alignment index out of range";
			return QTIPSIM_EINVAL;
		}
		// add_part1, ds.h:437-446
		n[c]++;
		bool take;
		if (kept[c].size() < k[c]) {
			kept[c].emplace_back();
			take = true;
		} else {
			rng = lstep(rng);
			const float r4 = lfloat(lout(rng));
			uint64_t off = (uint64_t)(r4 * (float)n[c]);
			if (off > n[c] - 1) off = n[c] - 1;
			take = off < k[c];
		}
		if (!take) continue;
		// list().back().init(...), qtip_parse.cpp:851-859 / :1150-1164
		Rec &r = kept[c].back();
		r = Rec();
		r.m[0] = mate(b.a1[i]);
		if (pair) {
			r.m[1] = mate(b.a2[i]);
			r.upstream1 = b.upstream1[i] ? 1 : 0;
			r.fraglen = b.fraglen[i];
		} else {
			r.mate_flag = (char)b.mate_flag[i];
			r.opp_len = b.opp_len[i];
		}
	}
	return QTIPSIM_OK;
}

int ModelBuilder::commit(TemplateStore &ts, std::string &err) {
	if (!open) {
		err = "qtipsim_model_commit before qtipsim_model_begin";
		return QTIPSIM_ESTATE;
	}
	for (int c = 0; c < NCAT; c++) {
		int rc;
		if (c == CAT_U || c == CAT_B) {
			std::vector<qtipsim_tmpl_unp> v(kept[c].size());
			for (size_t i = 0; i < v.size(); i++) {
				const Rec &r = kept[c][i];
				v[i].best_score = r.m[0].score;
				v[i].len = r.m[0].len;
				v[i].fw_flag = r.m[0].fw;
				v[i].mate_flag = r.mate_flag;
				v[i].opp_len = r.opp_len;
				v[i].qual = r.m[0].qual.c_str();
				v[i].xscript = r.m[0].xs.c_str();
			}
			rc = ts.set_unpaired(c, v.data(), v.size(), n[c], err);
		} else {
			std::vector<qtipsim_tmpl_pair> v(kept[c].size());
			for (size_t i = 0; i < v.size(); i++) {
				const Rec &r = kept[c][i];
				v[i].score_12 = r.m[0].score + r.m[1].score;  // qtip_parse.cpp:1151
				v[i].score_1 = r.m[0].score;
				v[i].len_1 = r.m[0].len;
				v[i].fw_flag_1 = r.m[0].fw;
				v[i].qual_1 = r.m[0].qual.c_str();
				v[i].xscript_1 = r.m[0].xs.c_str();
				v[i].score_2 = r.m[1].score;
				v[i].len_2 = r.m[1].len;
				v[i].fw_flag_2 = r.m[1].fw;
				v[i].qual_2 = r.m[1].qual.c_str();
				v[i].xscript_2 = r.m[1].xs.c_str();
				v[i].upstream1 = r.upstream1;
				v[i].fraglen = r.fraglen;
			}
			rc = ts.set_paired(c, v.data(), v.size(), n[c], err);
		}
		if (rc) return rc;
	}
	open = false;
	return QTIPSIM_OK;
}

}  // namespace qt
