/*
 * oracle/qtipsim_oracle.c -- TEST INFRASTRUCTURE ONLY (see qtipsim_oracle.h).
 *
 * Plain-C CPU restatement of the reference's tandem-read simulator.  Written from
 * the behaviour of the reference, not from its text; every function cites the
 * reference lines it restates.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off:
 * the reference is built for baseline x86-64, i.e. without fused multiply-add, so
 * every float/double operation below rounds once, exactly as written).
 */
#define _GNU_SOURCE
#include "qtipsim_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RNG ---- */

#define M1 2147483563
#define M2 2147483399

/* rnglib.cpp:941-1009: the package starts every generator at this seed pair. */
void orc_rng_initialize(orc_rng *r) {
	r->cg1 = 1234567890;
	r->cg2 = 123456789;
	r->ndraws = 0;
}

/* rnglib.cpp:1674-1755: range-check, then overwrite the current state only
 * (the re-initialisation call is commented out in the reference). */
int orc_rng_set_seed(orc_rng *r, int32_t a, int32_t b) {
	if (a < 1 || a >= M1) return 1;
	if (b < 1 || b >= M2) return 1;
	r->cg1 = a;
	r->cg2 = b;
	return 0;
}

/* rnglib.cpp:573-683.  The reference uses Schrage's decomposition to stay inside
 * 32-bit ints; 64-bit modular multiplication gives the same residues. */
int32_t orc_i4_uni(orc_rng *r) {
	int64_t s1 = ((int64_t)r->cg1 * 40014) % M1;
	int64_t s2 = ((int64_t)r->cg2 * 40692) % M2;
	r->cg1 = (int32_t)s1;
	r->cg2 = (int32_t)s2;
	r->ndraws++;
	int64_t z = s1 - s2;
	if (z < 1) z += M1 - 1;
	return (int32_t)z;
}

/* rnglib.cpp:1444-1505: int -> float, times a double constant, back to float. */
float orc_r4_uni_01(orc_rng *r) {
	int32_t i = orc_i4_uni(r);
	float f = (float)i;
	double d = (double)f * 4.656613057E-10;
	return (float)d;
}

/* ranlib.cpp:1094-1389 (BTPE, Kachitvichyanukul & Schmeiser 1988).  The reference
 * is C++ with <cmath> overloads: sqrt/log/fabs on float stay float, pow(float,int)
 * is double; double literals promote the surrounding arithmetic. */
int orc_ignbin(orc_rng *rng, int n, float pp) {
	float al, alv, amaxp, c, f, f1, f2, ffm, fm, g, p, p1, p2, p3, p4, q, qn, r, t;
	float u, v, w, w2, x, x1, x2, xl, xll, xlr, xm, xnp, xnpq, xr, ynorm, z, z2;
	int i, ix, k, m;

	if ((double)pp <= 0.0 || 1.0 <= (double)pp) {
		fprintf(stderr, "orc_ignbin: pp out of range\n");
		exit(1);
	}
	{
		float other = (float)(1.0 - (double)pp);
		p = (other < pp) ? other : pp; /* r4_min, ranlib.cpp:2248 */
	}
	q = (float)(1.0 - (double)p);
	xnp = (float)n * p;

	if ((double)xnp < 30.0) {
		/* inversion, ranlib.cpp:1198-1231 */
		qn = (float)pow((double)q, (double)n);
		r = p / q;
		g = r * (float)(n + 1);
		for (;;) {
			ix = 0;
			f = qn;
			u = orc_r4_uni_01(rng);
			for (;;) {
				if (u < f) {
					if (0.5 < (double)pp) ix = n - ix;
					return ix;
				}
				if (110 < ix) break;
				u = u - f;
				ix = ix + 1;
				f = f * (g / (float)ix - r);
			}
		}
	}

	/* BTPE set-up, ranlib.cpp:1232-1251 */
	ffm = xnp + p;
	m = (int)ffm;
	fm = (float)m;
	xnpq = xnp * q;
	p1 = (float)((double)(int)(2.195 * (double)sqrtf(xnpq) - 4.6 * (double)q) + 0.5);
	xm = (float)((double)fm + 0.5);
	xl = xm - p1;
	xr = xm + p1;
	c = (float)(0.134 + 20.5 / (15.3 + (double)fm));
	al = (ffm - xl) / (ffm - xl * p);
	xll = (float)((double)al * (1.0 + 0.5 * (double)al));
	al = (xr - ffm) / (xr * q);
	xlr = (float)((double)al * (1.0 + 0.5 * (double)al));
	p2 = (float)((double)p1 * (1.0 + (double)c + (double)c));
	p3 = p2 + c / xll;
	p4 = p3 + c / xlr;

	for (;;) {
		u = orc_r4_uni_01(rng) * p4;
		v = orc_r4_uni_01(rng);
		if (u < p1) { /* triangle */
			ix = (int)(xm - p1 * v + u);
			if (0.5 < (double)pp) ix = n - ix;
			return ix;
		}
		if (u <= p2) { /* parallelogram */
			x = xl + (u - p1) / c;
			v = (float)((double)(v * c) + 1.0 - (double)(fabsf(xm - x) / p1));
			if ((double)v <= 0.0 || 1.0 < (double)v) continue;
			ix = (int)x;
		} else if (u <= p3) { /* left tail */
			ix = (int)(xl + logf(v) / xll);
			if (ix < 0) continue;
			v = v * (u - p2) * xll;
		} else { /* right tail */
			ix = (int)(xr - logf(v) / xlr);
			if (n < ix) continue;
			v = v * (u - p3) * xlr;
		}
		k = abs(ix - m);
		if (k <= 20 || (double)xnpq / 2.0 - 1.0 <= (double)k) {
			/* explicit evaluation */
			f = 1.0f;
			r = p / q;
			g = (float)(n + 1) * r;
			if (m < ix) {
				for (i = m + 1; i <= ix; i++) f = f * (g / (float)i - r);
			} else if (ix < m) {
				for (i = ix + 1; i <= m; i++) f = f / (g / (float)i - r);
			}
			if (v <= f) {
				if (0.5 < (double)pp) ix = n - ix;
				return ix;
			}
		} else {
			/* squeezing */
			amaxp = (float)((double)((float)k / xnpq) *
				(((double)k * ((double)k / 3.0 + 0.625) + 0.1666666666666) / (double)xnpq + 0.5));
			ynorm = (float)((double)(-(float)(k * k)) / (2.0 * (double)xnpq));
			alv = logf(v);
			if (alv < ynorm - amaxp) {
				if (0.5 < (double)pp) ix = n - ix;
				return ix;
			}
			if (ynorm + amaxp < alv) continue;
			x1 = (float)(ix + 1);
			f1 = (float)((double)fm + 1.0);
			z = (float)(n + 1) - fm;
			w = (float)(n - ix + 1);
			z2 = z * z;
			x2 = x1 * x1;
			f2 = f1 * f1;
			w2 = w * w;
			{
				double d = (double)(xm * logf(f1 / x1));
				d = d + ((double)(n - m) + 0.5) * (double)logf(z / w);
				d = d + (double)((float)(ix - m) * logf(w * p / (x1 * q)));
				d = d + (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / (double)f2) / (double)f2) / (double)f2) / (double)f2) / (double)f1 / 166320.0;
				d = d + (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / (double)z2) / (double)z2) / (double)z2) / (double)z2) / (double)z / 166320.0;
				d = d + (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / (double)x2) / (double)x2) / (double)x2) / (double)x2) / (double)x1 / 166320.0;
				d = d + (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / (double)w2) / (double)w2) / (double)w2) / (double)w2) / (double)w / 166320.0;
				t = (float)d;
			}
			if (alv <= t) {
				if (0.5 < (double)pp) ix = n - ix;
				return ix;
			}
		}
	}
}

/* ---------------------------------------------------------------- FASTA ---- */

struct orc_fasta {
	size_t nrec;
	char **name;      /* up to first whitespace */
	char **name_full; /* whole header line */
	char **seq;       /* upper-cased, non-ACGT -> 'N' (fasta.cpp:18-35) */
	uint64_t *len;
	uint64_t filesize;
};

static char norm_base(int c) {
	switch (c) {
	case 'A': case 'a': return 'A';
	case 'C': case 'c': return 'C';
	case 'G': case 'g': return 'G';
	case 'T': case 't': return 'T';
	default: return 'N';
	}
}

/* Restates what FastaChunkwiseParser::next (fasta.cpp:42-127) extracts from one file:
 * '>' anywhere starts a header; the id runs to the first whitespace, the full name to
 * end of line; every other non-whitespace byte is a base.  Bases that precede the
 * first header form a record with an empty name. */
orc_fasta *orc_fasta_load(const char *path) {
	FILE *fh = fopen(path, "rb");
	if (!fh) return NULL;
	fseek(fh, 0, SEEK_END);
	long fsz = ftell(fh);
	fseek(fh, 0, SEEK_SET);
	unsigned char *raw = (unsigned char *)malloc((size_t)fsz + 1);
	if (fsz > 0 && fread(raw, 1, (size_t)fsz, fh) != (size_t)fsz) {
		fclose(fh);
		free(raw);
		return NULL;
	}
	fclose(fh);
	orc_fasta *fa = (orc_fasta *)calloc(1, sizeof(orc_fasta));
	fa->filesize = (uint64_t)fsz;
	size_t cap = 0;
	long i = 0;
	int have = 0; /* is there a current record? */
	uint64_t seqcap = 0;
	while (i < fsz) {
		int c = raw[i];
		if (c == '>' || (!have && !isspace(c))) {
			if (fa->nrec == cap) {
				cap = cap ? cap * 2 : 16;
				fa->name = (char **)realloc(fa->name, cap * sizeof(char *));
				fa->name_full = (char **)realloc(fa->name_full, cap * sizeof(char *));
				fa->seq = (char **)realloc(fa->seq, cap * sizeof(char *));
				fa->len = (uint64_t *)realloc(fa->len, cap * sizeof(uint64_t));
			}
			size_t r = fa->nrec++;
			have = 1;
			fa->len[r] = 0;
			seqcap = 1024;
			fa->seq[r] = (char *)malloc(seqcap);
			if (c == '>') {
				long s = ++i;
				while (i < fsz && !isspace(raw[i])) i++;
				fa->name[r] = strndup((const char *)raw + s, (size_t)(i - s));
				while (i < fsz && raw[i] != '\n' && raw[i] != '\r') i++;
				fa->name_full[r] = strndup((const char *)raw + s, (size_t)(i - s));
			} else {
				fa->name[r] = strdup("");
				fa->name_full[r] = strdup("");
			}
			continue;
		}
		if (!isspace(c)) {
			size_t r = fa->nrec - 1;
			if (fa->len[r] + 1 >= seqcap) {
				seqcap *= 2;
				fa->seq[r] = (char *)realloc(fa->seq[r], seqcap);
			}
			fa->seq[r][fa->len[r]++] = norm_base(c);
		}
		i++;
	}
	for (size_t r = 0; r < fa->nrec; r++) fa->seq[r][fa->len[r]] = '\0';
	free(raw);
	return fa;
}

void orc_fasta_free(orc_fasta *fa) {
	if (!fa) return;
	for (size_t r = 0; r < fa->nrec; r++) {
		free(fa->name[r]);
		free(fa->name_full[r]);
		free(fa->seq[r]);
	}
	free(fa->name);
	free(fa->name_full);
	free(fa->seq);
	free(fa->len);
	free(fa);
}

size_t orc_fasta_nrecords(const orc_fasta *fa) { return fa->nrec; }
const char *orc_fasta_name(const orc_fasta *fa, size_t i) { return fa->name[i]; }
const char *orc_fasta_name_full(const orc_fasta *fa, size_t i) { return fa->name_full[i]; }
const char *orc_fasta_seq(const orc_fasta *fa, size_t i, uint64_t *len) {
	if (len) *len = fa->len[i];
	return fa->seq[i];
}
uint64_t orc_fasta_filesize(const orc_fasta *fa) { return fa->filesize; }

/* The windows fasta.cpp:42-127 hands out for one record of R bases: the first holds
 * min(R, chunksz) bases; each later one re-presents the last `olap` bases of its
 * predecessor (fasta.cpp:63-70) followed by up to chunksz-olap new ones, and exists
 * only if at least one new base remains (`first` flag, fasta.cpp:72-97).  A window
 * that ended a record short of `olap` bases carries nothing over, but nothing
 * follows it in that record either. */
orc_chunk *orc_fasta_chunks(const orc_fasta *fa, uint64_t chunksz, uint64_t olap, size_t *nchunks) {
	size_t cap = 64, n = 0;
	orc_chunk *out = (orc_chunk *)malloc(cap * sizeof(orc_chunk));
	for (size_t r = 0; r < fa->nrec; r++) {
		uint64_t R = fa->len[r];
		uint64_t consumed = 0; /* bases of the record read so far */
		uint64_t carry = 0;
		while (consumed < R) {
			uint64_t room = chunksz - carry;
			uint64_t take = R - consumed < room ? R - consumed : room;
			uint64_t retsz = carry + take;
			consumed += take;
			if (n == cap) {
				cap *= 2;
				out = (orc_chunk *)realloc(out, cap * sizeof(orc_chunk));
			}
			out[n].record = (uint32_t)r;
			out[n].retsz = retsz;
			out[n].refoff = consumed - retsz;
			uint64_t nn = 0;
			const char *s = fa->seq[r] + out[n].refoff;
			for (uint64_t j = 0; j < retsz; j++) nn += (s[j] == 'N');
			out[n].n_count = nn;
			n++;
			carry = (retsz >= olap) ? olap : 0;
		}
	}
	*nchunks = n;
	return out;
}

/* ------------------------------------------------------------ simulator ---- */

void orc_out_init(orc_out *o, int want_trace) {
	memset(o, 0, sizeof(*o));
	o->want_trace = want_trace;
}

void orc_out_free(orc_out *o) {
	for (int i = 0; i < 7; i++) free(o->data[i]);
	free(o->trace);
	memset(o, 0, sizeof(*o));
}

static void out_reserve(orc_out *o, int s, size_t extra) {
	if (o->len[s] + extra > o->cap[s]) {
		size_t nc = o->cap[s] ? o->cap[s] : (1u << 16);
		while (nc < o->len[s] + extra) nc *= 2;
		o->data[s] = (char *)realloc(o->data[s], nc);
		o->cap[s] = nc;
	}
}

static void out_bytes(orc_out *o, int s, const char *b, size_t n) {
	out_reserve(o, s, n);
	memcpy(o->data[s] + o->len[s], b, n);
	o->len[s] += n;
}

static void out_trace(orc_out *o, int cat, uint32_t tmpl, uint32_t chunk, uint64_t off, uint64_t di) {
	if (!o->want_trace) return;
	if (o->ntrace == o->captrace) {
		o->captrace = o->captrace ? o->captrace * 2 : 1024;
		o->trace = (orc_trace_rec *)realloc(o->trace, o->captrace * sizeof(orc_trace_rec));
	}
	orc_trace_rec *t = &o->trace[o->ntrace++];
	t->cat = (uint8_t)cat;
	t->tmpl = tmpl;
	t->chunk = chunk;
	t->off = off;
	t->draw_idx = di;
}

/* edit_xscript.h:15-25 */
uint64_t orc_xscript_rflen(const char *xs) {
	uint64_t n = 0;
	for (; *xs; xs++) n += (*xs == 'S' || *xs == '=' || *xs == 'X' || *xs == 'D');
	return n;
}

/* simplesim.cpp:177-190 */
uint64_t orc_apply_function(float fraction, int function, uint64_t mn, uint64_t n) {
	if (n == 0) return 0;
	double nn = (double)n;
	if (function == ORC_FUNC_SQRT) nn = (double)fraction * sqrt(nn);
	else if (function == ORC_FUNC_LINEAR) nn *= (double)fraction;
	else if (function == ORC_FUNC_CONST) nn = (double)fraction;
	uint64_t v = (uint64_t)nn;
	return v > mn ? v : mn;
}

/* simplesim.h:26-28 */
static char draw_base(orc_rng *rng) {
	int k = (int)(orc_r4_uni_01(rng) * 4);
	return "ACGT"[k < 3 ? k : 3];
}

/* index draw used for templates and offsets: input_model.h:51, simplesim.cpp:253.
 * float times size_t is a float multiply; the product truncates to size_t. */
static uint64_t draw_index(orc_rng *rng, uint64_t n) {
	float prod = orc_r4_uni_01(rng) * (float)n;
	uint64_t v = (uint64_t)prod;
	return v < n - 1 ? v : n - 1;
}

typedef struct {
	char *seq;   /* mutated read, read (= reference-forward) orientation */
	size_t cap;
	size_t len;
	const char *qual;
	char *qual_own; /* for random mates */
	size_t qual_cap;
	int fw;
	int score;
	const char *refid;
	uint64_t refoff;
} sim_read;

static void read_reserve(sim_read *r, size_t n) {
	if (n + 1 > r->cap) {
		r->cap = (n + 1) * 2;
		r->seq = (char *)realloc(r->seq, r->cap);
	}
}

/* SimulatedRead::init + mutate, simplesim.h:68-84, simplesim.cpp:31-71 */
static int read_init(sim_read *r, orc_rng *rng, const char *ref, const char *qual,
                     const char *xs, int fw, int score, const char *refid, uint64_t refoff) {
	size_t newsz = strlen(qual);
	read_reserve(r, newsz + strlen(xs));
	r->qual = qual;
	r->fw = fw;
	r->score = score;
	r->refid = refid;
	r->refoff = refoff;
	size_t rd = 0, rf = 0;
	for (; *xs; xs++) {
		switch (*xs) {
		case '=':
			r->seq[rd++] = ref[rf++];
			break;
		case 'X': {
			char b;
			do { b = draw_base(rng); } while (b == ref[rf]);
			r->seq[rd++] = b;
			rf++;
			break;
		}
		case 'I':
			r->seq[rd++] = draw_base(rng);
			break;
		case 'D':
			rf++;
			break;
		case 'S':
			r->seq[rd++] = draw_base(rng);
			rf++;
			break;
		default:
			fprintf(stderr, "Unknown operation in edit transcript: '%c'\n", *xs);
			return 1; /* reference: throw 1, simplesim.cpp:58-60 */
		}
	}
	r->len = rd;
	if (rd != newsz) return 2; /* reference: assert, simplesim.cpp:63 */
	return 0;
}

/* SimulatedRead::init_random, simplesim.h:89-117 */
static void read_init_random(sim_read *r, orc_rng *rng, size_t len, int fw, int score,
                             const char *refid, uint64_t refoff) {
	read_reserve(r, len);
	if (len + 1 > r->qual_cap) {
		r->qual_cap = (len + 1) * 2;
		r->qual_own = (char *)realloc(r->qual_own, r->qual_cap);
	}
	for (size_t i = 0; i < len; i++) {
		r->seq[i] = draw_base(rng);
		r->qual_own[i] = 'I';
	}
	r->qual_own[len] = '\0';
	r->len = len;
	r->qual = r->qual_own;
	r->fw = fw;
	r->score = score;
	r->refid = refid;
	r->refoff = refoff;
}

static char comp_base(char c) { /* asc2dnacomp, simplesim.cpp:76-97, on {A,C,G,T,N} */
	switch (c) {
	case 'A': return 'T';
	case 'C': return 'G';
	case 'G': return 'C';
	case 'T': return 'A';
	case 'N': return 'N';
	default: return 0;
	}
}

/* body shared by write and write_pair: simplesim.cpp:111-126, :151-169 */
static void emit_body(orc_out *o, int s, const sim_read *r) {
	size_t len = strlen(r->qual);
	out_reserve(o, s, 2 * len + 8);
	char *p = o->data[s] + o->len[s];
	if (r->fw) {
		memcpy(p, r->seq, len);
	} else {
		for (size_t i = 0; i < len; i++) p[i] = comp_base(r->seq[len - i - 1]);
	}
	p += len;
	*p++ = '\n'; *p++ = '+'; *p++ = '\n';
	if (r->fw) {
		memcpy(p, r->qual, len);
	} else {
		for (size_t i = 0; i < len; i++) p[i] = r->qual[len - i - 1];
	}
	p += len;
	*p++ = '\n';
	o->len[s] = (size_t)(p - o->data[s]);
}

/* SimulatedRead::write, simplesim.cpp:102-127 */
static void emit_unpaired(orc_out *o, int s, const sim_read *r, const char *typ) {
	char name[1024];
	int n = snprintf(name, sizeof(name), "@qtip!:%s:%c:%llu:%d:%s\n", r->refid,
	                 r->fw ? '+' : '-', (unsigned long long)r->refoff, r->score, typ);
	out_bytes(o, s, name, (size_t)n);
	emit_body(o, s, r);
}

/* SimulatedRead::write_pair, simplesim.cpp:132-171: both files get the same name. */
static void emit_pair(orc_out *o, int s1, int s2, const sim_read *r1, const sim_read *r2, const char *typ) {
	char name[2048];
	int n = snprintf(name, sizeof(name), "@qtip!:%s:%c:%llu:%d:%s:%c:%llu:%d:%s\n",
	                 r1->refid, r1->fw ? '+' : '-', (unsigned long long)r1->refoff, r1->score,
	                 r2->refid, r2->fw ? '+' : '-', (unsigned long long)r2->refoff, r2->score, typ);
	out_bytes(o, s1, name, (size_t)n);
	emit_body(o, s1, r1);
	out_bytes(o, s2, name, (size_t)n);
	emit_body(o, s2, r2);
}

static uint64_t max_u64(uint64_t a, uint64_t b) { return a > b ? a : b; }

int orc_simulate(
	const orc_fasta *fa, uint64_t chunksz,
	const orc_tmpl_unp *tu, size_t nu_t, uint64_t u_added,
	const orc_tmpl_unp *tb, size_t nb_t, uint64_t b_added,
	const orc_tmpl_pair *tc, size_t nc_t, uint64_t c_added,
	const orc_tmpl_pair *td, size_t nd_t, uint64_t d_added,
	float fraction, int function,
	uint64_t min_u, uint64_t min_c, uint64_t min_d, uint64_t min_b,
	orc_rng *rng, orc_out *out)
{
	/* olap_: simplesim.h:228-230 over input_model.h:38-41 (unpaired: max reference
	 * length) and :112-115 (paired: max fraglen_) -- SURVEY Q8 */
	uint64_t olap = 0;
	for (size_t i = 0; i < nu_t; i++) olap = max_u64(olap, orc_xscript_rflen(tu[i].xscript));
	for (size_t i = 0; i < nb_t; i++) olap = max_u64(olap, orc_xscript_rflen(tb[i].xscript));
	for (size_t i = 0; i < nc_t; i++) olap = max_u64(olap, tc[i].fraglen);
	for (size_t i = 0; i < nd_t; i++) olap = max_u64(olap, td[i].fraglen);
	if (!(chunksz > olap)) return 3; /* assert, fasta.h:41 */

	/* simplesim.cpp:205-209 */
	uint64_t nu = orc_apply_function(fraction, function, min_u, u_added);
	uint64_t nb = orc_apply_function(fraction, function, min_b, b_added);
	uint64_t nc = orc_apply_function(fraction, function, min_c, c_added);
	uint64_t nd = orc_apply_function(fraction, function, min_d, d_added);
	out->target[0] = nu; out->target[1] = nb; out->target[2] = nc; out->target[3] = nd;
	if (nu + nb + nc + nd == 0) return 4; /* assert, simplesim.cpp:209 */
	/* an empty model with a non-zero target would trip assert(!empty()) in draw() */
	if ((nu && !nu_t) || (nb && !nb_t) || (nc && !nc_t) || (nd && !nd_t)) return 5;

	const uint64_t tot_fasta_len = fa->filesize; /* simplesim.h:272-286, SURVEY Q6 */
	size_t nchunks = 0;
	orc_chunk *chunks = orc_fasta_chunks(fa, chunksz, olap, &nchunks);
	sim_read rd1, rd2;
	memset(&rd1, 0, sizeof(rd1));
	memset(&rd2, 0, sizeof(rd2));
	int rc = 0;

	for (size_t ci = 0; ci < nchunks && !rc; ci++) {
		const orc_chunk *ch = &chunks[ci];
		const uint64_t retsz = ch->retsz;
		const char *refid = fa->name[ch->record];
		const char *buf = fa->seq[ch->record] + ch->refoff;
		const uint64_t refoff = ch->refoff;
		if (retsz < olap) continue; /* simplesim.cpp:221-223 */
		const uint64_t nchances = retsz - olap + 1;
		/* simplesim.cpp:226: float * float / (float)size_t, clamped */
		float binom_p = (float)nchances * 1.1f / (float)tot_fasta_len;
		if (0.999f < binom_p) binom_p = 0.999f; /* std::min(a, b) is (b < a) ? b : a */
		/* simplesim.cpp:228-234 */
		if ((int)ch->n_count > (int)(0.9 * (double)retsz)) continue;
		const uint64_t nslots = retsz - olap;

		/* unpaired, simplesim.cpp:242-273 */
		uint64_t nu_samp = (uint64_t)orc_ignbin(rng, (int)nu, binom_p);
		for (uint64_t i = 0; i < nu_samp && !rc; i++) {
			uint64_t di = rng->ndraws;
			const uint64_t ti = draw_index(rng, nu_t);
			const orc_tmpl_unp *t = &tu[ti];
			if (nslots == 0) { rc = 6; break; } /* assert(nslots > 0), :252 */
			const uint64_t off = draw_index(rng, nslots);
			out_trace(out, 0, (uint32_t)ti, (uint32_t)ci, off, di);
			rc = read_init(&rd1, rng, buf + off, t->qual, t->xscript, t->fw_flag == 'T',
			               t->best_score, refid, refoff + off);
			if (rc) break;
			out->wrote[0]++;
			emit_unpaired(out, 0, &rd1, "u");
		}
		if (rc) break;

		/* bad-end, simplesim.cpp:279-334 */
		uint64_t nb_samp = (uint64_t)orc_ignbin(rng, (int)nb, binom_p);
		for (uint64_t i = 0; i < nb_samp && !rc; i++) {
			uint64_t di = rng->ndraws;
			const uint64_t ti = draw_index(rng, nb_t);
			const orc_tmpl_unp *t = &tb[ti];
			const int mate1 = (t->mate_flag == '1');
			const uint64_t off = draw_index(rng, nslots);
			out_trace(out, 1, (uint32_t)ti, (uint32_t)ci, off, di);
			sim_read *al = mate1 ? &rd1 : &rd2;
			sim_read *op = mate1 ? &rd2 : &rd1;
			rc = read_init(al, rng, buf + off, t->qual, t->xscript, t->fw_flag == 'T',
			               t->best_score, refid, refoff + off);
			if (rc) break;
			if (t->opp_len <= 0) { rc = 7; break; } /* assert(len > 0), simplesim.h:96 */
			read_init_random(op, rng, (size_t)t->opp_len, t->fw_flag == 'T', t->best_score,
			                 refid, refoff + off);
			out->wrote[1]++;
			emit_pair(out, 1, 2, &rd1, &rd2, mate1 ? "b1" : "b2");
		}
		if (rc) break;

		/* concordant + discordant, simplesim.cpp:340-397 */
		uint64_t nc_samp = (uint64_t)orc_ignbin(rng, (int)nc, binom_p);
		uint64_t nd_samp = (uint64_t)orc_ignbin(rng, (int)nd, binom_p);
		for (uint64_t i = 0; i < nc_samp + nd_samp && !rc; i++) {
			const int conc = i < nc_samp;
			uint64_t di = rng->ndraws;
			const uint64_t ti = draw_index(rng, conc ? nc_t : nd_t);
			const orc_tmpl_pair *t = conc ? &tc[ti] : &td[ti];
			const uint64_t off = draw_index(rng, nslots);
			out_trace(out, conc ? 2 : 3, (uint32_t)ti, (uint32_t)ci, off, di);
			const uint64_t rflen_1 = orc_xscript_rflen(t->xscript_1);
			const uint64_t rflen_2 = orc_xscript_rflen(t->xscript_2);
			uint64_t off_1, off_2;
			if (t->upstream1) { /* SURVEY Q9 */
				off_1 = off;
				off_2 = off + max_u64(t->fraglen, rflen_2) - rflen_2;
			} else {
				off_2 = off;
				off_1 = off + max_u64(t->fraglen, rflen_1) - rflen_1;
			}
			rc = read_init(&rd1, rng, buf + off_1, t->qual_1, t->xscript_1, t->fw_flag_1 == 'T',
			               t->score_1, refid, refoff + off_1);
			if (rc) break;
			rc = read_init(&rd2, rng, buf + off_2, t->qual_2, t->xscript_2, t->fw_flag_2 == 'T',
			               t->score_2, refid, refoff + off_2);
			if (rc) break;
			out->wrote[conc ? 2 : 3]++;
			emit_pair(out, conc ? 3 : 5, conc ? 4 : 6, &rd1, &rd2, conc ? "c" : "d");
		}
	}
	free(chunks);
	free(rd1.seq); free(rd1.qual_own);
	free(rd2.seq); free(rd2.qual_own);
	return rc;
}

/* ---------------------------------------------------------------------------
 * Tandem-SAM correctness labelling: Alignment::set_correctness, qtip_parse.cpp:519-713.
 * flag / pos: SAM FLAG and 1-based POS of an aligned record; returns the value the reference leaves in
 * `correct`: 1, 0, or -1 (name in neither simulated-read format).  Reads past a string's end (which the
 * reference can do on malformed names) are cut short and count as "incorrect".
 * ------------------------------------------------------------------------- */
static int lab_abs_int(uint64_t a, uint64_t b) {  /* abs((int)(a - b)) on size_t operands */
	int32_t d = (int32_t)(uint32_t)(a - b);
	return d < 0 ? (int32_t)(0u - (uint32_t)d) : d;
}

int orc_set_correctness(const char *qname, const char *rname, int flag, uint64_t pos, int wiggle) {
	const size_t rname_len = strlen(rname);
	const int is_fw = (flag & 16) == 0;                                      /* :134-136 */
	const char mate_flag = (flag & 64) ? '1' : ((flag & 128) ? '2' : '0');   /* :146-148 */
	const char *startswith = "qtip!";                                        /* simplesim.h:20-24 */
	if (strncmp(qname, startswith, 5) == 0) {                                /* :522 */
		const char *cur = qname + 5;
		const char *end = qname + strlen(qname);
		if (*cur != ':') return 0;                                           /* assert(*qname_cur == sim_sep), :526 */
		cur++;
		if (mate_flag != '2' && strncmp(cur, rname, rname_len) != 0) return 0;   /* :528-532 */
		if ((size_t)(end - cur) < rname_len) return 0;
		cur += rname_len;
		if (*(cur++) != ':') return 0;                                       /* :534 */
		if (mate_flag != '2' && cur[0] != (is_fw ? '+' : '-')) return 0;     /* :539 */
		if (cur >= end) return 0;
		cur++;
		if (*(cur++) != ':') return 0;
		uint64_t refoff = 0;                                                 /* :547-551 */
		while (isdigit((unsigned char)*cur)) refoff = refoff * 10 + (uint64_t)(*cur++ - '0');
		if (mate_flag != '2' && lab_abs_int(refoff, pos - 1) >= wiggle) return 0;   /* :552 */
		if (*(cur++) != ':') return 0;
		if (*cur == '-') cur++;                                              /* score1, :559-571 */
		while (isdigit((unsigned char)*cur)) cur++;
		if (*(cur++) != ':') return 0;
		if (cur[0] == 'u' && (cur[1] == '\0' || isspace((unsigned char)cur[1]))) return 1;   /* :575-578 */
		if (mate_flag != '2') return 1;                                      /* :580-583 */
		if (strncmp(cur, rname, rname_len) != 0) return 0;                   /* :585-588 */
		if ((size_t)(end - cur) < rname_len) return 0;
		cur += rname_len;
		if (*(cur++) != ':') return 0;
		if (cur[0] != (is_fw ? '+' : '-')) return 0;                         /* :594 */
		cur++;
		if (*(cur++) != ':') return 0;
		refoff = 0;
		while (isdigit((unsigned char)*cur)) refoff = refoff * 10 + (uint64_t)(*cur++ - '0');
		if (lab_abs_int(refoff, pos - 1) >= wiggle) return 0;                /* :607 */
		if (*(cur++) != ':') return 0;
		if (*cur == '-') cur++;                                              /* score2 */
		while (isdigit((unsigned char)*cur)) cur++;
		if (*(cur++) != ':') return 0;
		return 1;                                                            /* :631 */
	}
	/* wgsim-style names of simulated input reads, :634-712 */
	const size_t qname_len = strlen(qname);
	int nund = 0, ncolon = 0;
	for (size_t i = 0; i < qname_len; i++) {
		if (qname[i] == '_') nund++;
		else if (qname[i] == ':' && nund >= 3) ncolon++;
	}
	if (!(nund >= 8 && ncolon == 4)) return -1;
	const char *cur = qname, *end = qname + qname_len;
	if (strncmp(cur, rname, rname_len) != 0) return 0;
	if ((size_t)(end - cur) < rname_len) return 0;
	cur += rname_len;
	if (*cur++ != '_') return 0;
	uint64_t frag_start = 0, frag_end = 0;
	while (isdigit((unsigned char)*cur)) frag_start = frag_start * 10 + (uint64_t)(*cur++ - '0');
	if (*cur++ != '_') return 0;
	while (isdigit((unsigned char)*cur)) frag_end = frag_end * 10 + (uint64_t)(*cur++ - '0');
	if (*cur++ != '_') return 0;
	while (ncolon > 0) {                                                     /* :681-685 */
		if (cur >= end) return 0;
		if (*cur++ == ':') ncolon--;
	}
	while (cur < end && isdigit((unsigned char)*cur)) cur++;                  /* while(isdigit(*qname_cur++)); :686 */
	if (cur >= end) return 0;
	cur++;
	int len1 = 0, len2 = 0;
	while (isdigit((unsigned char)*cur)) len1 = len1 * 10 + (int)(*cur++ - '0');
	if (*cur++ != '_') return 0;
	while (isdigit((unsigned char)*cur)) len2 = len2 * 10 + (int)(*cur++ - '0');
	if (*cur++ != '_') return 0;
	const int flip = *cur == '1';
	const int mate1 = mate_flag != '2';
	const int len = mate1 ? len1 : len2;
	if ((!flip) == mate1) return lab_abs_int(pos, frag_start) < wiggle;       /* left end of fragment, :704-706 */
	return lab_abs_int(pos, frag_end - (uint64_t)(int64_t)len + 1) < wiggle;  /* right end, :707-709 */
}

/* ---------------------------------------------------------------------------
 * SAM fields -> edit transcript: parse_cigar (qtip_parse.cpp:223-253), mdz_to_list (:258-300),
 * cigar_to_edit_xscript (:307-319), cigar_and_mdz_to_edit_xscript (:326-401), lpos / rpos (:164-186).
 * Like the reference this builds the run / op lists and the MD:Z element list first.  Returns 0 and fills
 * out (NUL-terminated, cap bytes) or a status > 0 where the reference asserts / throws.
 * ------------------------------------------------------------------------- */
int orc_edit_xscript(const char *cigar, const char *mdz, uint64_t pos, char *out, size_t cap, int *left_clip,
                     int *right_clip, uint64_t *lpos, uint64_t *rpos) {
	const size_t clen = strlen(cigar);
	int *crun = (int *)malloc((clen + 1) * sizeof(int));
	char *cops = (char *)malloc(clen + 1);
	size_t nc = 0, i = 0;
	int eqx = 0, rc = 0;
	*left_clip = *right_clip = 0;
	while (i < clen) {                                                       /* parse_cigar */
		if (!isdigit((unsigned char)cigar[i])) { rc = 1; goto done; }
		int run = 0;
		do { run = run * 10 + (cigar[i] - '0'); i++; } while (isdigit((unsigned char)cigar[i]));
		if (!(isalpha((unsigned char)cigar[i]) || cigar[i] == '=')) { rc = 1; goto done; }
		if (nc == 0 && cigar[i] == 'S') *left_clip = run;
		else if (i + 1 >= clen && cigar[i] == 'S') *right_clip = run;
		if (cigar[i] == 'X' || cigar[i] == '=') eqx = 1;
		cops[nc] = cigar[i];
		crun[nc++] = run;
		i++;
	}
	size_t n = 0;
#define PUSH(ch) do { if (n + 1 < cap) out[n] = (ch); n++; } while (0)
	if (eqx) {                                                               /* cigar_to_edit_xscript */
		for (size_t k = 0; k < nc; k++) {
			if (cops[k] == 'M' || cops[k] == 'P') { rc = 2; goto done; }
			for (int j = 0; j < crun[k]; j++) PUSH(cops[k]);
		}
	} else {
		if (!mdz) { rc = 3; goto done; }
		const size_t mlen = strlen(mdz);                                     /* mdz_to_list */
		int *mop = (int *)malloc((mlen + 1) * sizeof(int)), *mrun = (int *)malloc((mlen + 1) * sizeof(int));
		size_t nm = 0, j = 0;
		while (j < mlen) {
			if (isdigit((unsigned char)mdz[j])) {
				int run = 0;
				do { run = run * 10 + (mdz[j] - '0'); j++; } while (j < mlen && isdigit((unsigned char)mdz[j]));
				if (run > 0) { mop[nm] = 0; mrun[nm++] = run; }
			} else if (isalpha((unsigned char)mdz[j])) {
				int run = 0;
				do { j++; run++; } while (j < mlen && isalpha((unsigned char)mdz[j]));
				mop[nm] = 1; mrun[nm++] = run;
			} else if (mdz[j] == '^') {
				j++;
				int run = 0;
				while (j < mlen && isalpha((unsigned char)mdz[j])) { j++; run++; }
				mop[nm] = 2; mrun[nm++] = run;
			} else {
				j++;
			}
		}
		size_t mdo = 0;                                                      /* cigar_and_mdz_to_edit_xscript */
		for (size_t k = 0; k < nc && !rc; k++) {
			const char cop = cops[k];
			if (cop == 'M') {
				int left = crun[k];
				while (left > 0 && mdo < nm) {
					if (mop[mdo] == 2) { rc = 4; break; }
					const int comb = left < mrun[mdo] ? left : mrun[mdo];
					left -= comb;
					if (mop[mdo] == 0) {
						for (int t = 0; t < comb; t++) PUSH('=');
					} else {
						if (mrun[mdo] != comb) { rc = 4; break; }
						for (int t = 0; t < comb; t++) PUSH('X');
					}
					if (comb < mrun[mdo]) mrun[mdo] -= comb;
					else mdo++;
				}
			} else if (cop == 'I') { for (int t = 0; t < crun[k]; t++) PUSH('I'); }
			else if (cop == 'D') {
				if (mdo >= nm || mop[mdo] != 2 || mrun[mdo] != crun[k]) { rc = 4; break; }
				mdo++;
				for (int t = 0; t < crun[k]; t++) PUSH('D');
			} else if (cop == 'N') { for (int t = 0; t < crun[k]; t++) PUSH('N'); }
			else if (cop == 'S') { for (int t = 0; t < crun[k]; t++) PUSH('S'); }
			else if (cop == 'H') { }
			else rc = 2;
		}
		if (!rc && mdo != nm) rc = 4;
		free(mop);
		free(mrun);
	}
#undef PUSH
	if (!rc) {
		if (n + 1 > cap) rc = 5;
		else {
			out[n] = '\0';
			*lpos = pos - (uint64_t)(int64_t)*left_clip;                     /* lpos, :164-167 */
			size_t k = 0;                                                    /* rpos, :173-186 */
			while (k < n && out[k] == 'S') k++;
			k++;
			uint64_t mv = 0;
			for (; k < n; k++) mv += (out[k] == 'S' || out[k] == 'D' || out[k] == 'X' || out[k] == '=');
			*rpos = pos + mv - 1;
		}
	}
done:
	free(crun);
	free(cops);
	return rc;
}
