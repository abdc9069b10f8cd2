/*
 * include/qtipsim.h -- C ABI of libqtipsim, the B200-native tandem-read simulator.
 *
 * This is the drop-in boundary for the one data-parallel stage of Qtip: everything the
 * reference does between `StreamingSimulator ss(...)` and the seven fclose() calls in
 * qtip_parse.cpp:1825-1867.  The reference has no FFI layer for this path (it is a C++
 * class consumed by one caller); the entry points below are what a binding for it
 * would bind, one per reference interface, cited on each declaration.  src/ holds the
 * header-compatible C++ shim (simplesim.h, input_model.h, ...) that forwards the
 * reference's own class interface to these calls, so the unchanged qtip_parse.cpp
 * compiles and links against it.
 *
 * Plain pointers and sizes only: no C++ or torch types cross this line.  Every call
 * returns QTIPSIM_OK (0) or a negative error code; qtipsim_last_error() gives the text.
 * There is no CPU fallback: qtipsim_create fails with QTIPSIM_ENODEVICE when no CUDA
 * device (sm_100) is usable.
 */
#ifndef QTIPSIM_H
#define QTIPSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTIPSIM_ABI_VERSION 1

enum {
	QTIPSIM_OK = 0,
	QTIPSIM_EINVAL = -1,    /* bad argument */
	QTIPSIM_ENODEVICE = -2, /* no usable CUDA device / driver */
	QTIPSIM_ECUDA = -3,     /* CUDA runtime error (text in last_error) */
	QTIPSIM_EIO = -4,       /* FASTA unreadable: reference `throw 1`, fasta.cpp:54-58,83-85 */
	QTIPSIM_EXSCRIPT = -5,  /* unknown edit op / length mismatch: simplesim.cpp:58-63 */
	QTIPSIM_ERANGE = -6,    /* chunksz <= olap (fasta.h:41), empty model with target, seed out of range */
	QTIPSIM_ESTATE = -7,    /* call order violated (no genome, no run yet, ...) */
	QTIPSIM_ENOMEM = -8
};

/* read categories, in the order the reference simulates them per window
 * (simplesim.cpp:242, :279, :340-341) */
enum { QTIPSIM_CAT_U = 0, QTIPSIM_CAT_B = 1, QTIPSIM_CAT_C = 2, QTIPSIM_CAT_D = 3 };

/* output streams, in the reference's file order (qtip_parse.cpp:1831-1837):
 * _reads_u, _reads_b_1, _reads_b_2, _reads_c_1, _reads_c_2, _reads_d_1, _reads_d_2 */
enum {
	QTIPSIM_OUT_U = 0, QTIPSIM_OUT_B1 = 1, QTIPSIM_OUT_B2 = 2, QTIPSIM_OUT_C1 = 3,
	QTIPSIM_OUT_C2 = 4, QTIPSIM_OUT_D1 = 5, QTIPSIM_OUT_D2 = 6, QTIPSIM_NOUT = 7
};

/* simplesim.h:197-201 */
enum { QTIPSIM_FUNC_LINEAR = 1, QTIPSIM_FUNC_SQRT = 2, QTIPSIM_FUNC_CONST = 3 };

/* Where the random draws come from.
 * REPLAY : consumes the reference's rnglib/ranlib stream (L'Ecuyer-Cote combined MLCG,
 *          rnglib.cpp:573-683, ranlib.cpp:1094-1389) draw for draw; output is
 *          byte-identical to the reference for the same model, FASTA and RNG state.
 * PHILOX : counter-based Philox4x32-10 keyed by (seed, read index); same sampling
 *          model, different stream; fully parallel. */
enum { QTIPSIM_RNG_REPLAY = 0, QTIPSIM_RNG_PHILOX = 1 };

/* One unpaired / bad-end template.  Mirrors TemplateUnpaired, template.h:97-104.
 * qual and xscript are NUL-terminated and only read during the call. */
typedef struct {
	int32_t best_score;
	int32_t len;
	char fw_flag;   /* 'T' = forward strand */
	char mate_flag; /* '0', '1' or '2' */
	int32_t opp_len;
	const char *qual;
	const char *xscript; /* ops = X I D S, reference-forward orientation */
} qtipsim_tmpl_unp;

/* One concordant / discordant pair template.  Mirrors TemplatePaired, template.h:218-230. */
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
	int32_t upstream1; /* bool */
	uint64_t fraglen;
} qtipsim_tmpl_pair;

/* Arguments of StreamingSimulator::simulate_batch (simplesim.h:251-257; note the
 * reference's order u, c, d, b) plus the state the reference keeps in globals. */
typedef struct {
	uint64_t chunksz;  /* rolling FASTA window, qtip_parse.cpp:1840 passes 128*1024 */
	float fraction;    /* sim-factor */
	int32_t function;  /* QTIPSIM_FUNC_* */
	uint64_t min_u, min_c, min_d, min_b;
	int32_t rng_mode;  /* QTIPSIM_RNG_* */
	int32_t cg1, cg2;  /* REPLAY: rnglib state at entry (get_state, rnglib.cpp:516); the
	                      state at exit is written back here.  With shards the chain is
	                      sequential across them: shard r's entry state is shard r-1's exit
	                      state (the caller passes the two integers along, rank to rank) */
	uint64_t philox_seed; /* PHILOX */
	/* sharding across GPUs: this handle simulates the windows whose index i satisfies
	 * shard_rank*nwin/shard_count <= i < (shard_rank+1)*nwin/shard_count.  {0,1} = all. */
	uint32_t shard_rank, shard_count;
} qtipsim_run_params;

typedef struct {
	uint64_t wrote[4];          /* u, b, c, d reads / pairs written (simplesim.cpp:399-406) */
	uint64_t target[4];         /* u, b, c, d targets after apply_function (simplesim.cpp:205-208) */
	uint64_t bytes[QTIPSIM_NOUT]; /* FASTQ bytes per output stream */
	uint64_t draws;             /* REPLAY: rnglib draws consumed */
	uint64_t algorithmic_bytes; /* sum over reads of (FASTQ bytes written + ceil(rflen/4)) */
	uint32_t kernel_launches;   /* kernels launched by the call */
	float ms_pack;              /* device time, FASTA pack kernels (0 if genome was already resident) */
	float ms_plan;              /* device/host time, window table + draw-chain resolution + offsets */
	float ms_simulate;          /* device time, simulate+emit kernel(s) */
	float ms_total;             /* device time for the whole run call */
} qtipsim_stats;

typedef struct qtipsim qtipsim_t;

int qtipsim_abi_version(void);

/* Handle bound to one CUDA device.  device < 0 picks the current device. */
int qtipsim_create(qtipsim_t **out, int device);
void qtipsim_destroy(qtipsim_t *h);
const char *qtipsim_last_error(const qtipsim_t *h); /* h may be NULL: last create error */

/* Genome: replaces FastaChunkwiseParser (fasta.h:24-105, fasta.cpp:42-127) and
 * estimate_fasta_length (simplesim.h:272-286).  Loads (multi-)FASTA, upper-cases,
 * maps non-ACGT to N (fasta.cpp:18-35), packs to 2 bit/base + 1 bit/base N mask in HBM.
 * Only the first path of a multi-file list has defined behaviour in the reference
 * (SURVEY Q13); more than one path is rejected with QTIPSIM_EINVAL. */
int qtipsim_load_fasta(qtipsim_t *h, const char *const *paths, size_t npaths);
/* Same, from FASTA text already in host memory (pinned or pageable).
 * Threading: a handle is single-threaded like the reference (qtip_parse.cpp:73), with one exception a host may
 * use to overlap its own work with the upload: one thread may be inside qtipsim_load_fasta* while another calls
 * qtipsim_set_templates_* on the same handle (they touch disjoint state); both must have returned before
 * qtipsim_run. */
int qtipsim_load_fasta_bytes(qtipsim_t *h, const void *bytes, size_t nbytes);
/* size_t num_estimated_bases() const, simplesim.h:263-265 (= FASTA file bytes) */
uint64_t qtipsim_num_estimated_bases(const qtipsim_t *h);
uint64_t qtipsim_num_records(const qtipsim_t *h);

/* Packed genome exchange for multi-GPU replication: rank 0 packs, the
 * caller broadcasts these device buffers (NCCL over NVLink), other ranks adopt them.
 * Buffers: 0 = 2-bit bases (u32 words), 1 = N mask (u32 words), 2 = record table
 * (host-side blob; small).  export gives pointer+size; adopt takes ownership of
 * nothing -- it copies (device-to-device for 0/1). */
int qtipsim_genome_export(qtipsim_t *h, int which, void **ptr, uint64_t *nbytes, int *is_device);
int qtipsim_genome_adopt(qtipsim_t *h, const void *blob, uint64_t blob_bytes,
                         const void *d_bases, uint64_t bases_bytes,
                         const void *d_nmask, uint64_t nmask_bytes);

/* Packed-genome cache: a trial after the first need not parse and pack the FASTA again -- the reference
 * re-reads it byte by byte in every trial (fasta.cpp:42-127; 22 s of a 63 s run at 3.1 Gbp).  save: writes the loaded
 * genome in packed form (record table, 2 bit/base, N mask: 0.375 byte per base) to `cache_path`, stamped with the FASTA's
 * size and modification time.  load_cached: if `cache_path` holds a cache stamped with `fasta_path`'s current size and
 * mtime, uploads it (1.17 GB instead of 3.13 GB of text at 3.1 Gbp, no pack kernels); otherwise loads and packs the
 * FASTA and (re)writes the cache.  *hit (may be NULL) tells which happened. */
int qtipsim_genome_save(qtipsim_t *h, const char *cache_path, const char *fasta_path);
int qtipsim_load_fasta_cached(qtipsim_t *h, const char *fasta_path, const char *cache_path, int *hit);

/* Input model: replaces InputModelUnpaired / InputModelPaired (input_model.h:26-42,
 * :100-116).  kept = templates in v[], added = alignments seen (num_added(), :68-70). */
int qtipsim_set_templates_unpaired(qtipsim_t *h, int cat /* U or B */,
                                   const qtipsim_tmpl_unp *v, size_t kept, uint64_t added);
int qtipsim_set_templates_paired(qtipsim_t *h, int cat /* C or D */,
                                 const qtipsim_tmpl_pair *v, size_t kept, uint64_t added);
/* olap_ (simplesim.h:228-230): max over models of max reference length / fragment length */
uint64_t qtipsim_olap(const qtipsim_t *h);

/* StreamingSimulator::simulate_batch (simplesim.cpp:195-407): plan + simulate + emit on
 * the device.  FASTQ bytes stay in HBM until fetched. */
int qtipsim_run(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats);

/* Results of the last run.  copy: device-to-host into dst (cap >= bytes[stream]).
 * device_ptr: the HBM buffer itself (valid until the next run / destroy). */
int qtipsim_result_copy(qtipsim_t *h, int stream, void *dst, uint64_t cap);
int qtipsim_result_device_ptr(qtipsim_t *h, int stream, const void **ptr, uint64_t *nbytes);
/* Appends each stream to an open file descriptor (the shim passes fileno() of the
 * caller's FILE* after fflush; SimulatedRead::write/write_pair, simplesim.cpp:102-171). */
int qtipsim_result_write_fds(qtipsim_t *h, const int fds[QTIPSIM_NOUT]);

/* Output paths to the aligner that avoid the temporary-file round trip (the reference writes seven files and
 * the driver concatenates / passes them on, qtip:378-409).
 * stream_fds: like write_fds, but all seven streams are written CONCURRENTLY (one writer per descriptor, each with its
 * own device-to-host pipeline): the descriptors may be FIFOs or pipes whose reader -- an aligner taking _1 and _2 files --
 * consumes the paired files in lock step.
 * interleaved: the two streams of a pair category (QTIPSIM_CAT_B, _C or _D) as ONE stream, mate 1's record followed by mate
 * 2's for every pair (what `bowtie2 --interleaved` / `bwa mem -p` read, scripts/fastq_interleave.py): interleaved on the
 * device, copied into dst (cap bytes); *nbytes = bytes[_1] + bytes[_2]. */
int qtipsim_result_stream_fds(qtipsim_t *h, const int fds[QTIPSIM_NOUT]);
int qtipsim_result_interleaved(qtipsim_t *h, int cat, void *dst, uint64_t cap, uint64_t *nbytes);

/* Host-buffer convenience for end-to-end use: FASTA text + templates must already be
 * set; runs, then copies all seven streams into caller buffers dst[i] (cap[i] bytes). */
int qtipsim_run_to_host(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats,
                        void *const dst[QTIPSIM_NOUT], const uint64_t cap[QTIPSIM_NOUT]);

/* Streaming end-to-end path (one GPU): the same bytes as qtipsim_load_fasta_bytes + qtipsim_run_to_host, with the
 * upload of the FASTA text, the device pack, the simulation and the download of the FASTQ bytes overlapped --
 * the reference, too, streams the FASTA once and emits reads window by window (simplesim.cpp:216-397,
 * fasta.cpp:42-127).  begin: starts the chunked upload of `bytes` (must stay valid until stream_run returns;
 * pinned memory makes the copies asynchronous) and returns at once, so the caller can set the templates
 * meanwhile.  run: packs and simulates every record as soon as the next record's header has arrived and appends
 * its reads to dst[i] (cap[i] bytes); stats->bytes[i] = bytes written to dst[i].  Afterwards the genome is
 * resident as after qtipsim_load_fasta_bytes.  REPLAY mode chains the generator state through the parts. */
int qtipsim_stream_begin(qtipsim_t *h, const void *bytes, size_t nbytes);
int qtipsim_stream_run(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats,
                       void *const dst[QTIPSIM_NOUT], const uint64_t cap[QTIPSIM_NOUT]);

/* Tandem-SAM correctness labelling: Alignment::set_correctness (qtip_parse.cpp:519-713, called at :823, :944-945)
 * for a batch of ALIGNED records of the SAM the aligner produced from the tandem reads -- the consumer of the read
 * names the simulator emits.  qnames / rnames: the records' QNAME / RNAME strings, NUL-terminated, concatenated
 * (qnames_bytes / rnames_bytes in total, terminators included); q_off[i] / r_off[i]: where record i's strings start;
 * flag[i], pos[i]: its SAM FLAG and 1-based POS; wiggle: qtip_parse.cpp:797 (default 30).  correct[i] = 1 if the
 * alignment lies within `wiggle` of the origin the name records (reference, strand and offset of the mate the FLAG
 * names), 0 if not, -1 if the name is in neither simulated-read format (the reference leaves its field at -1).  All
 * pointers are host memory; the labels are computed on the device. */
int qtipsim_label_correctness(qtipsim_t *h, const char *qnames, uint64_t qnames_bytes, const uint64_t *q_off,
                              const char *rnames, uint64_t rnames_bytes, const uint64_t *r_off, const int32_t *flag,
                              const uint64_t *pos, uint64_t n, int32_t wiggle, int8_t *correct);

/* SAM alignment fields -> edit transcripts, the first half of the input model's template construction
 * (qtip_parse.cpp: parse_cigar :223-253, mdz_to_list :258-300, cigar_to_edit_xscript :307-319,
 * cigar_and_mdz_to_edit_xscript :326-401, lpos / rpos :164-186), for a batch of aligned records.
 * cigars / mdzs: NUL-terminated strings, concatenated; c_off[i] / m_off[i]: record i's CIGAR / MD:Z (m_off[i] =
 * UINT64_MAX: the record has no MD:Z tag); pos[i]: 1-based POS.  Out: xscripts (cap bytes) receives the transcripts,
 * NUL-terminated, record i's at x_off[i] (x_off has n + 1 entries; x_off[n] = bytes used); left_clip / right_clip;
 * lpos / rpos (the reference's, incl. rpos skipping the op after the leading soft clip), from which the
 * caller derives a pair's fraglen = min(max-allowed-fraglen, dn.rpos - up.lpos + 1) and upstream1 (:154-158, :955-957);
 * status[i] = 0, or why the reference would have asserted / thrown on the record (1 bad CIGAR, 2 unsupported op,
 * 3 neither extended CIGAR nor MD:Z, 4 MD:Z does not fit the CIGAR): such records get an empty transcript.
 * Undefined in the reference and answered here: a transcript with at most one op after its leading soft clip (a one-column
 * alignment, an all-S or an empty transcript) -- rpos() :173-186 steps over the leading S run and the op after it before
 * its assert, so it aborts or reads past the terminator; such a record gets status 0 and rpos = pos + span - 1.
 * Returns QTIPSIM_ERANGE if cap is too small (x_off[n] then holds the bytes needed).  Host pointers throughout. */
int qtipsim_edit_xscripts(qtipsim_t *h, const char *cigars, uint64_t cigars_bytes, const uint64_t *c_off, const char *mdzs,
                          uint64_t mdzs_bytes, const uint64_t *m_off, const uint64_t *pos, uint64_t n, char *xscripts,
                          uint64_t cap, uint64_t *x_off, int32_t *left_clip, int32_t *right_clip, uint64_t *lpos,
                          uint64_t *rpos, int32_t *status);

/* The input model's reservoirs, fed with plain arrays: the second half of the template construction.  The
 * reference's SAM pass offers every aligned record to one of four reservoirs -- unpaired and bad-end at
 * qtip_parse.cpp:849-861, concordant and discordant at :1148-1166, ReservoirSampledEList::add_part1 ds.h:437-446 --
 * and builds the four input models from what they kept (:1826-1829).  These three calls do the same for records
 * handed over in batches, so the caller no longer builds TemplateUnpaired / TemplatePaired records one by one:
 * begin(k): four empty reservoirs of capacity k[QTIPSIM_CAT_*] (input-model-size);
 * add(batch, &cg1, &cg2): offers the batch's records in the order given (SAM order) and keeps what the reference
 *   keeps -- including its habit of overwriting the LAST kept template, not a random one, once a reservoir is full (the
 *   call sites write through list().back()) -- consuming one r4_uni_01 draw per record offered to a full reservoir from
 *   the generator state cg1 / cg2 (get_state, rnglib.cpp:516), which is written back advanced (set_seed, :1674);
 * commit(kept, added): the kept templates become the handle's four input models, as after qtipsim_set_templates_*
 *   with added = records offered; kept / added (may be NULL) report list().size() / size() per category.
 * A batch has two levels.  Alignment level, one entry per aligned SAM record, indexed by a1 / a2: best_score, len
 * (Alignment::best_score, ::len), fw (1 = forward strand), QUAL strings and edit transcripts as NUL-terminated strings in
 * one buffer each with their offsets -- xscripts / x_off exactly as qtipsim_edit_xscripts wrote them.  Record level, n
 * entries: the category, the record's alignment(s), and mate_flag ('0' / '1' / '2') + opp_len for QTIPSIM_CAT_U / _B or
 * upstream1 + fraglen for QTIPSIM_CAT_C / _D (the arrays of the other kind may be NULL if no such record is in the
 * batch).  score_12 of a pair is best_score[a1] + best_score[a2] (:1151).  Host pointers, read during the call only. */
typedef struct {
	const int32_t *best_score;
	const int32_t *len;
	const uint8_t *fw;
	const char *quals;
	const uint64_t *q_off;
	const char *xscripts;
	const uint64_t *x_off;
	uint64_t n_aln; /* entries in the alignment-level arrays: a1[i], a2[i] < n_aln */
	uint64_t n;
	const uint8_t *cat;
	const uint64_t *a1;
	const uint64_t *a2;
	const uint8_t *mate_flag;
	const int32_t *opp_len;
	const uint8_t *upstream1;
	const uint64_t *fraglen;
} qtipsim_model_batch;

int qtipsim_model_begin(qtipsim_t *h, const uint64_t k[4]);
int qtipsim_model_add(qtipsim_t *h, const qtipsim_model_batch *batch, int32_t *cg1, int32_t *cg2);
int qtipsim_model_commit(qtipsim_t *h, uint64_t kept[4], uint64_t added[4]);

#ifdef __cplusplus
}
#endif
#endif /* QTIPSIM_H */
