/*
 * oracle/qtipsim_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's tandem-read simulator hot path
 * (BenLangmead/qtip src/simplesim.{h,cpp}, input_model.h, fasta.{h,cpp},
 * rnglib.cpp, ranlib.cpp::ignbin).  It is the checker for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing in the product (qtip_b200/, src/) links or calls it.
 *
 * Parity pin: this restatement is checked byte-for-byte against the UNMODIFIED
 * reference (oracle/_ref/qtip-parse and oracle/_ref/libref.so, built by
 * oracle/Makefile from /root/reference) and against golden FASTQ fixtures generated
 * from that reference (tests/golden/, generator tests/golden/make_golden.py); the
 * reference itself ships no golden vectors for this path.
 */
#ifndef QTIPSIM_ORACLE_H
#define QTIPSIM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* L'Ecuyer-Cote combined MLCG state of generator 0 (rnglib.cpp:573-683). */
typedef struct {
	int32_t cg1;
	int32_t cg2;
	uint64_t ndraws; /* number of i4_uni steps taken since creation (trace aid) */
} orc_rng;

void orc_rng_initialize(orc_rng *r);                  /* rnglib.cpp:941-1009 */
int orc_rng_set_seed(orc_rng *r, int32_t a, int32_t b); /* rnglib.cpp:1674-1755; !=0 on range error */
int32_t orc_i4_uni(orc_rng *r);                       /* rnglib.cpp:573-683 */
float orc_r4_uni_01(orc_rng *r);                      /* rnglib.cpp:1444-1505 */
int orc_ignbin(orc_rng *r, int n, float pp);          /* ranlib.cpp:1094-1389 */

/* Same layout as qtipsim_tmpl_unp / qtipsim_tmpl_pair in include/qtipsim.h
 * (mirrors template.h:97-104 and :218-230). */
typedef struct {
	int32_t best_score;
	int32_t len;
	char fw_flag;   /* 'T' or 'F' */
	char mate_flag; /* '0','1','2' */
	int32_t opp_len;
	const char *qual;
	const char *xscript;
} orc_tmpl_unp;

typedef struct {
	int32_t score_12;
	int32_t score_1;
	int32_t len_1;
	char fw_flag_1;
	const char *qual_1;
	const char *xscript_1;
	int32_t score_2;
	int32_t len_2;
	char fw_flag_2;
	const char *qual_2;
	const char *xscript_2;
	int32_t upstream1;
	uint64_t fraglen;
} orc_tmpl_pair;

/* One FASTA window as FastaChunkwiseParser::next returns it (fasta.cpp:42-127). */
typedef struct {
	uint32_t record;  /* index of the FASTA record */
	uint64_t refoff;  /* 0-based record offset of the first byte */
	uint64_t retsz;   /* number of bytes */
	uint64_t n_count; /* how many of them are 'N' */
} orc_chunk;

/* Parsed FASTA: records (name up to first whitespace, normalised bases) and the
 * windows the reference's parser would return for (chunksz, olap). */
typedef struct orc_fasta orc_fasta;
orc_fasta *orc_fasta_load(const char *path);
void orc_fasta_free(orc_fasta *fa);
size_t orc_fasta_nrecords(const orc_fasta *fa);
const char *orc_fasta_name(const orc_fasta *fa, size_t i);
const char *orc_fasta_name_full(const orc_fasta *fa, size_t i);
const char *orc_fasta_seq(const orc_fasta *fa, size_t i, uint64_t *len);
uint64_t orc_fasta_filesize(const orc_fasta *fa);
/* Returns a malloc'd array of windows; caller frees with free(). */
orc_chunk *orc_fasta_chunks(const orc_fasta *fa, uint64_t chunksz, uint64_t olap, size_t *nchunks);

/* Optional per-read trace (for debugging planners against the chain). */
typedef struct {
	uint8_t cat;       /* 0=u 1=b 2=c 3=d */
	uint32_t tmpl;
	uint32_t chunk;    /* index among the windows of orc_fasta_chunks */
	uint64_t off;      /* window-relative offset drawn */
	uint64_t draw_idx; /* rng.ndraws when the read started (before the template draw) */
} orc_trace_rec;

typedef struct {
	/* seven FASTQ byte streams in the reference's file order:
	 * u, b_1, b_2, c_1, c_2, d_1, d_2 (qtip_parse.cpp:1831-1837) */
	char *data[7];
	size_t len[7];
	size_t cap[7];
	uint64_t wrote[4];  /* u, b, c, d reads/pairs written (simplesim.cpp:399-406) */
	uint64_t target[4]; /* u, b, c, d targets after apply_function */
	orc_trace_rec *trace;
	size_t ntrace, captrace;
	int want_trace;
} orc_out;

void orc_out_init(orc_out *o, int want_trace);
void orc_out_free(orc_out *o);

enum { ORC_FUNC_LINEAR = 1, ORC_FUNC_SQRT = 2, ORC_FUNC_CONST = 3 }; /* simplesim.h:197-201 */

uint64_t orc_apply_function(float fraction, int function, uint64_t mn, uint64_t n); /* simplesim.cpp:177-190 */
/* CIGAR (+ MD:Z) -> edit transcript, clips, lpos / rpos: qtip_parse.cpp:164-186, :223-401; 0 or a status */
int orc_edit_xscript(const char *cigar, const char *mdz, uint64_t pos, char *out, size_t cap, int *left_clip,
                     int *right_clip, uint64_t *lpos, uint64_t *rpos);
/* Alignment::set_correctness, qtip_parse.cpp:519-713: 1 / 0 / -1 */
int orc_set_correctness(const char *qname, const char *rname, int flag, uint64_t pos, int wiggle);
uint64_t orc_xscript_rflen(const char *xs);                                         /* edit_xscript.h:15-25 */

/* StreamingSimulator ctor + simulate_batch (simplesim.h:214-245, simplesim.cpp:195-407).
 * Note the reference's argument order min_u, min_c, min_d, min_b.  rng is read and
 * advanced.  Returns 0, or non-zero where the reference would throw/abort. */
int orc_simulate(
	const orc_fasta *fa, uint64_t chunksz,
	const orc_tmpl_unp *tu, size_t nu, uint64_t u_added,
	const orc_tmpl_unp *tb, size_t nb, uint64_t b_added,
	const orc_tmpl_pair *tc, size_t nc, uint64_t c_added,
	const orc_tmpl_pair *td, size_t nd, uint64_t d_added,
	float fraction, int function,
	uint64_t min_u, uint64_t min_c, uint64_t min_d, uint64_t min_b,
	orc_rng *rng, orc_out *out);

#ifdef __cplusplus
}
#endif
#endif
