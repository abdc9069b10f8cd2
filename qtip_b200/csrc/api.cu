// api.cu -- the C ABI of libqtipsim (include/qtipsim.h): handle, genome upload, template
// upload, run orchestration (plan -> lengths -> scan -> simulate), result access.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <map>
#include <string>
#include <vector>

#include "../../include/qtipsim.h"
#include "common.h"
#include "genome.h"
#include "kernels.cuh"
#include "lecuyer.h"
#include "philox.h"
#include "planner.h"
#include "resolve.h"
#include "templates.h"
#include "model.h"

using namespace qt;

namespace {

std::string g_create_err;

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) {
			cudaGetLastError();
			want = n;
			e = cudaMalloc(&p, want);
		}
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	template <class T>
	T *as() const { return (T *)p; }
};

double now_ms() {
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct qtipsim {
	int device = 0;
	cudaStream_t st = nullptr;
	std::string err;

	HostGenome hg;
	bool have_genome = false;
	bool host_packed = false;  // hg.bases/nmask valid on the host (needed by the host planner)
	DevBuf d_bases, d_nmask, d_nblk, d_recs, d_names;
	DevBuf d_text, d_pk_a, d_pk_b, d_pk_c, d_pk_d, d_pk_e;  // K0 scratch
	float ms_pack_pending = 0;

	TemplateStore ts;
	ModelBuilder model;  // the reservoirs of qtipsim_model_begin / _add / _commit
	bool tmpl_dirty[NCAT] = {true, true, true, true};
	DevBuf d_t[NCAT], d_qual[NCAT], d_xs[NCAT], d_qual_out[NCAT], d_ops[NCAT], d_segs[NCAT], d_xsteps[NCAT], d_prog[NCAT], d_xprog[NCAT];
	DevBuf d_setups, d_pclass, d_rsresult;

	DevBuf d_jt1, d_jt2, d_jh1, d_jh2;
	bool have_jt = false;
	// band-speculative replay resolver (resolve2.cu)
	DevBuf d_r2_aw, d_r2_cum, d_r2_anchor, d_r2_counter, d_r2_slot, d_r2_relay, d_r2_end, d_r2_cn, d_r2_wres, d_r2_cnt, d_r2_off, d_r2_misc;
	uint32_t r2_stages_last = 0;  // stage launches of the last replay run (diagnostics)

	DevBuf d_win, d_ncount, d_cum;
	// window table + per-window N counts: functions of (genome, chunksz, olap) only, kept across runs
	std::vector<Window> win_cache;
	std::vector<uint32_t> ncount_cache;
	uint64_t win_cache_chunksz = 0, win_cache_olap = ~0ull;
	bool win_cache_ok = false;
	// binomial set-ups per window class: functions of the above plus the targets and the shard's window range
	std::vector<BinomSetup> setup_cache;
	std::vector<BinomTables> tab_cache;  // BTPE acceptance bounds per set-up (lecuyer.h), for the band resolver
	DevBuf d_tabs;
	DevBuf d_label;  // qtipsim_label_correctness / qtipsim_edit_xscripts scratch
	DevBuf d_xs_out;
	std::vector<uint32_t> pclass_cache;
	uint64_t setup_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	bool setup_cache_ok = false;
	DevBuf d_plan[NCAT], d_len0, d_len1, d_off0[NCAT], d_off1[NCAT], d_scan_tmp, d_scan_status, d_small;
	DevBuf d_out[NOUT];
	uint64_t out_bytes[NOUT] = {0, 0, 0, 0, 0, 0, 0};
	uint64_t out_reads[NCAT] = {0, 0, 0, 0};  // reads / pairs per category in d_out (offsets in d_off0 / d_off1)
	DevBuf d_inter;
	bool have_result = false;

	void *h_stage = nullptr;  // pinned staging for fd writes
	size_t h_stage_cap = 0;

	// header-line scan (shared by the one-shot and the streaming loader): text ranges of the sequence lines of
	// every record of hg.recs, and the record whose end has not been seen yet
	std::vector<uint64_t> seg_start, seg_end;
	uint64_t hdr_end = 0;
	bool open_rec = false, open_is_leading = false;
	bool names_overflow = false;  // the refid arena outgrew the 32-bit name offsets (loader fails with QTIPSIM_ERANGE)
	RecDev open_desc{};
	uint64_t open_start = 0;
	uint32_t recs_packed = 0;
	// streaming loader (qtipsim_stream_begin / qtipsim_stream_run): upload, pack, simulate and download overlap
	cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
	const uint8_t *stream_text = nullptr;
	uint64_t stream_n = 0, stream_chunk = 0, stream_cap_bases = 0;
	std::vector<cudaEvent_t> stream_ev;
	bool stream_open = false;
	DevBuf d_out2[NOUT];
	std::thread feeder;  // issues the upload in small pieces, a few in flight: control copies never queue behind the bulk
	std::atomic<uint64_t> chunks_issued{0};
	std::atomic<int> bulk_err{0};
	double stream_t0 = 0;

	// Control data (tables of a few KB to a few MB, counts read back) while bulk copies are in flight on the copy engines
	// (streaming path): the copy engine of a direction serves copies in issue order, and pageable copies are staged
	// synchronously, so a small copy would wait behind megabytes of text or FASTQ.  Instead the data crosses the link
	// inside a kernel that reads / writes a pinned, mapped ring (no copy engine involved).
	unsigned char *ctl_ring = nullptr;
	size_t ctl_cap = 0, ctl_pos = 0;
	bool ctl_zero_copy = false;
	struct CtlPending {
		void *dst;
		const void *src;
		size_t n;
	};
	std::vector<CtlPending> ctl_pending;  // device-to-host results to move out of the ring at the next stream sync

	int fail(int code, const std::string &m) {
		err = m;
		return code;
	}
	int cuda_fail(cudaError_t e, const char *what) {
		err = std::string(what) + ": " + cudaGetErrorString(e);
		cudaGetLastError();
		return QTIPSIM_ECUDA;
	}
};

#define CK(call)                                                 \
	do {                                                         \
		cudaError_t e__ = (call);                                \
		if (e__ != cudaSuccess) return h->cuda_fail(e__, #call); \
	} while (0)

// synchronises the handle's stream and delivers the read-backs that went through the control ring
static cudaError_t ctl_sync(qtipsim *h) {
	const cudaError_t e = cudaStreamSynchronize(h->st);
	if (e == cudaSuccess)
		for (const auto &pd : h->ctl_pending) memcpy(pd.dst, pd.src, pd.n);
	h->ctl_pending.clear();
	h->ctl_pos = 0;
	return e;
}

static void *ctl_alloc(qtipsim *h, size_t n) {
	const size_t need = (n + 255) & ~(size_t)255;
	if (!h->ctl_ring || need > h->ctl_cap) return nullptr;
	if (h->ctl_pos + need > h->ctl_cap && ctl_sync(h) != cudaSuccess) return nullptr;  // everything in the ring has been consumed now
	void *p = h->ctl_ring + h->ctl_pos;
	h->ctl_pos += need;
	return p;
}

static cudaError_t ctl_h2d(qtipsim *h, void *dst, const void *src, size_t n) {
	void *p = h->ctl_zero_copy && n ? ctl_alloc(h, n) : nullptr;
	if (!p) return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, h->st);
	memcpy(p, src, n);
	launch_ctl_copy(dst, p, n, h->st);
	return cudaGetLastError();
}

static cudaError_t ctl_d2h(qtipsim *h, void *dst, const void *src, size_t n) {
	void *p = h->ctl_zero_copy && n ? ctl_alloc(h, n) : nullptr;
	if (!p) return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, h->st);
	launch_ctl_copy(p, src, n, h->st);
	h->ctl_pending.push_back(qtipsim::CtlPending{dst, p, n});
	return cudaGetLastError();
}

extern "C" {

int qtipsim_abi_version(void) { return QTIPSIM_ABI_VERSION; }

int qtipsim_create(qtipsim_t **out, int device) {
	if (!out) return QTIPSIM_EINVAL;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		g_create_err = std::string("no usable CUDA device: ") +
		               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
		               " (libqtipsim has no CPU fallback)";
		cudaGetLastError();
		return QTIPSIM_ENODEVICE;
	}
	if (device < 0) {
		if (cudaGetDevice(&device) != cudaSuccess) device = 0;
	}
	if (device >= ndev) {
		g_create_err = "device index out of range";
		return QTIPSIM_EINVAL;
	}
	e = cudaSetDevice(device);
	if (e != cudaSuccess) {
		g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
		return QTIPSIM_ENODEVICE;
	}
	qtipsim *h = new qtipsim();
	h->device = device;
	e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		g_create_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
		delete h;
		return QTIPSIM_ECUDA;
	}
	*out = h;
	return QTIPSIM_OK;
}

void qtipsim_destroy(qtipsim_t *h) {
	if (!h) return;
	cudaSetDevice(h->device);
	// an upload started by qtipsim_stream_begin and never consumed still writes into d_text: let it finish first
	if (h->feeder.joinable()) h->feeder.join();
	if (h->st_h2d) cudaStreamSynchronize(h->st_h2d);
	if (h->st_d2h) cudaStreamSynchronize(h->st_d2h);
	cudaStreamSynchronize(h->st);
	if (h->ctl_ring) cudaFreeHost(h->ctl_ring);
	DevBuf *all[] = {&h->d_inter, &h->d_xs_out, &h->d_label, &h->d_tabs, &h->d_jh1, &h->d_jh2, &h->d_r2_aw, &h->d_r2_cum, &h->d_r2_anchor, &h->d_r2_counter, &h->d_r2_slot, &h->d_r2_relay, &h->d_r2_end, &h->d_r2_cn,
	                 &h->d_r2_wres, &h->d_r2_cnt, &h->d_r2_off, &h->d_r2_misc, &h->d_setups, &h->d_pclass, &h->d_rsresult, &h->d_text, &h->d_pk_a, &h->d_pk_b, &h->d_pk_c, &h->d_pk_d, &h->d_pk_e, &h->d_bases, &h->d_nmask, &h->d_nblk, &h->d_recs, &h->d_names, &h->d_jt1, &h->d_jt2, &h->d_win,
	                 &h->d_ncount, &h->d_cum, &h->d_len0, &h->d_len1, &h->d_scan_tmp, &h->d_scan_status, &h->d_small};
	for (DevBuf *b : all) b->release();
	for (int c = 0; c < NCAT; c++) {
		h->d_t[c].release();
		h->d_qual[c].release();
		h->d_xs[c].release();
		h->d_qual_out[c].release();
		h->d_ops[c].release();
		h->d_segs[c].release();
		h->d_xsteps[c].release();
		h->d_prog[c].release();
		h->d_xprog[c].release();
		h->d_plan[c].release();
		h->d_off0[c].release();
		h->d_off1[c].release();
	}
	for (int s = 0; s < NOUT; s++) h->d_out[s].release();
	for (int s = 0; s < NOUT; s++) h->d_out2[s].release();
	if (h->feeder.joinable()) h->feeder.join();
	for (cudaEvent_t e : h->stream_ev) cudaEventDestroy(e);
	if (h->st_h2d) cudaStreamDestroy(h->st_h2d);
	if (h->st_d2h) cudaStreamDestroy(h->st_d2h);
	if (h->h_stage) cudaFreeHost(h->h_stage);
	cudaStreamDestroy(h->st);
	delete h;
}

const char *qtipsim_last_error(const qtipsim_t *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

// 64-base block summary of the N mask (bit b = any N in bases [64 b, 64 b + 64)): lets the simulate kernel
// decide per read, from one or two L2-resident words, whether the N-free fast path applies
static int build_nblock(qtipsim *h, uint64_t mask_words) {
	const uint64_t nblk_words = (h->hg.total_bases / 64) / 32 + 4;
	CK(h->d_nblk.ensure(nblk_words * 4));
	launch_nblock(h->d_nmask.as<uint32_t>(), mask_words, h->d_nblk.as<uint32_t>(), nblk_words, h->st);
	return QTIPSIM_OK;
}

static int upload_genome(qtipsim *h) {
	CK(cudaSetDevice(h->device));
	h->hg.fill_name16();
	const HostGenome &g = h->hg;
	CK(h->d_bases.ensure(g.bases.size() * 4));
	CK(h->d_nmask.ensure(g.nmask.size() * 4));
	CK(h->d_recs.ensure(g.recs.size() * sizeof(RecDev) + 16));
	CK(h->d_names.ensure(g.names.size() + 16));
	CK(cudaMemcpyAsync(h->d_bases.p, g.bases.data(), g.bases.size() * 4, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(h->d_nmask.p, g.nmask.data(), g.nmask.size() * 4, cudaMemcpyHostToDevice, h->st));
	CK(ctl_h2d(h, h->d_recs.p, g.recs.data(), g.recs.size() * sizeof(RecDev)));
	if (!g.names.empty())
		CK(ctl_h2d(h, h->d_names.p, g.names.data(), g.names.size()));
	{
		const int rc = build_nblock(h, g.nmask.size());
		if (rc) return rc;
	}
	CK(ctl_sync(h));
	h->have_genome = true;
	h->win_cache_ok = false;
	h->setup_cache_ok = false;
	h->host_packed = true;
	h->have_result = false;
	return QTIPSIM_OK;
}

// ---- K0 on the device (pack.cu).  The host only resolves header lines; classification, counting and packing
// of the sequence bytes run on the GPU.  Three pieces, shared by the one-shot loader (device_pack) and the
// streaming loader (qtipsim_stream_run): reset, header resolution, packing of a range of complete records. ----

static void genome_reset(qtipsim *h, uint64_t n) {
	// an upload started by qtipsim_stream_begin and never consumed must be over before its buffers are reused
	if (h->feeder.joinable()) h->feeder.join();
	h->stream_open = false;
	HostGenome &g = h->hg;
	g.recs.clear();
	g.names.clear();
	g.bases.clear();
	g.nmask.clear();
	g.file_bytes = n;
	g.total_bases = 0;
	h->seg_start.clear();
	h->seg_end.clear();
	h->hdr_end = 0;
	h->names_overflow = false;
	// sequence bytes before the first header form a record with an empty name (fasta.cpp:72-124)
	h->open_rec = true;
	h->open_is_leading = true;
	h->open_desc = RecDev{0, 0, 0, 0};
	h->open_start = 0;
	h->recs_packed = 0;
	h->have_genome = false;
	h->win_cache_ok = false;
	h->setup_cache_ok = false;
	h->host_packed = false;
	h->have_result = false;
}

static inline bool fa_is_sp(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// the open record ends at text position `end`: it joins hg.recs (the headerless leading record only if it has bases)
static void close_open_record(qtipsim *h, const uint8_t *text, uint64_t end) {
	if (!h->open_rec) return;
	bool keep = true;
	if (h->open_is_leading) {
		keep = false;
		for (uint64_t i = 0; i < end && !keep; i++) keep = !fa_is_sp(text[i]);
	}
	if (keep) {
		h->hg.recs.push_back(h->open_desc);
		h->seg_start.push_back(h->open_start);
		h->seg_end.push_back(end);
	}
	h->open_rec = false;
	h->open_is_leading = false;
}

// gts: ascending positions of '>' not yet looked at.  A header closes the open record and opens the next one.
static void resolve_headers(qtipsim *h, const uint8_t *text, uint64_t n, const std::vector<uint64_t> &gts, bool final) {
	HostGenome &g = h->hg;
	for (size_t k = 0; k < gts.size(); k++) {
		const uint64_t p = gts[k];
		if (p < h->hdr_end) continue;  // a '>' inside a header line is part of the name
		close_open_record(h, text, p);
		uint64_t i = p + 1;
		const uint64_t s0 = i;
		while (i < n && !fa_is_sp(text[i])) i++;
		if (g.names.size() + (i - s0) > 0xfffffff0ull) h->names_overflow = true;  // name offsets are 32-bit
		h->open_desc = RecDev{0, 0, (uint32_t)g.names.size(), (uint32_t)(i - s0)};
		g.names.append((const char *)text + s0, (size_t)(i - s0));
		while (i < n && text[i] != '\n' && text[i] != '\r') i++;
		h->hdr_end = i;
		h->open_rec = true;
		h->open_start = i;
	}
	if (final) close_open_record(h, text, n);
}

// positions of '>' in text [c0, c1), already on the device, ascending
static int find_headers_chunk(qtipsim *h, const uint8_t *text, uint64_t c0, uint64_t c1, std::vector<uint64_t> &gts) {
	gts.clear();
	if (c1 <= c0) return QTIPSIM_OK;
	const uint32_t cap = 1u << 22;
	CK(h->d_pk_a.ensure((size_t)cap * 8 + 16));
	CK(h->d_small.ensure(16 * 8));
	CK(cudaMemsetAsync(h->d_small.p, 0, 16 * 8, h->st));
	launch_find_gt(h->d_text.as<uint8_t>() + c0, c1 - c0, h->d_pk_a.as<uint64_t>(), cap, h->d_small.as<uint32_t>(), h->st);
	uint32_t ngt = 0;
	CK(ctl_d2h(h, &ngt, h->d_small.p, 4));
	CK(ctl_sync(h));
	if (ngt > cap) {  // pathological: fall back to a host scan for '>'
		for (const uint8_t *q0 = text + c0, *e = text + c1; q0 < e;) {
			const uint8_t *q = (const uint8_t *)memchr(q0, '>', (size_t)(e - q0));
			if (!q) break;
			gts.push_back((uint64_t)(q - text));
			q0 = q + 1;
		}
	} else {
		gts.resize(ngt);
		if (ngt) CK(cudaMemcpy(gts.data(), h->d_pk_a.p, (size_t)ngt * 8, cudaMemcpyDeviceToHost));
		for (uint64_t &v : gts) v += c0;
		std::sort(gts.begin(), gts.end());
	}
	return QTIPSIM_OK;
}

// Packs records [recs_packed, hg.recs.size()) -- their text is on the device -- behind what is packed already.
// presized: d_bases / d_nmask were allocated and cleared for stream_cap_bases positions by the caller.
static int pack_records(qtipsim *h, bool presized) {
	HostGenome &g = h->hg;
	const uint32_t r0 = h->recs_packed, r1 = (uint32_t)g.recs.size();
	const uint32_t nseg = r1 - r0;
	const bool first = r0 == 0 && g.total_bases == 0;
	const uint64_t *seg_start = h->seg_start.data() + r0, *seg_end = h->seg_end.data() + r0;
	std::vector<uint64_t> unit_base((size_t)nseg + 1, 0);
	for (uint32_t s = 0; s < nseg; s++) unit_base[s + 1] = unit_base[s] + (seg_end[s] - seg_start[s] + 4095) / 4096;
	const uint64_t nunits = unit_base[nseg];
	std::vector<uint64_t> unit_pref((size_t)nunits + 1, 0);
	if (nunits) {
		CK(h->d_pk_b.ensure((size_t)(3 * nseg + 1) * 8 + 16));
		uint64_t *d_ss = h->d_pk_b.as<uint64_t>(), *d_se = d_ss + nseg, *d_ub = d_se + nseg;
		CK(ctl_h2d(h, d_ss, seg_start, nseg * 8));
		CK(ctl_h2d(h, d_se, seg_end, nseg * 8));
		CK(ctl_h2d(h, d_ub, unit_base.data(), ((size_t)nseg + 1) * 8));
		CK(h->d_pk_c.ensure((size_t)nunits * 4 + 16));                   // counts
		CK(h->d_pk_d.ensure(((size_t)nunits + 1) * 8 + 16));             // exclusive prefix
		CK(h->d_pk_e.ensure(((size_t)nunits / 1024 + 4) * 8));           // scan scratch
		launch_count_bases(h->d_text.as<uint8_t>(), d_ss, d_se, d_ub, nseg, nunits, h->d_pk_c.as<uint32_t>(), h->st);
		launch_scan(h->d_pk_c.as<uint32_t>(), h->d_pk_d.as<uint64_t>(), nunits, h->d_pk_e.as<uint64_t>(),
		            h->d_pk_d.as<uint64_t>() + nunits, h->st);
		CK(ctl_d2h(h, unit_pref.data(), h->d_pk_d.p, ((size_t)nunits + 1) * 8));
		CK(ctl_sync(h));
	}
	// ---- genome coordinates: 64 N in front, every record 64-aligned with an N gap behind it ----
	std::vector<int64_t> seg_gbase(nseg);
	std::vector<uint64_t> pad_lo, pad_hi;
	uint64_t gp = first ? 64 : g.total_bases;
	if (first) {
		pad_lo.push_back(0);
		pad_hi.push_back(64);
	}
	for (uint32_t s = 0; s < nseg; s++) {
		const uint64_t len = unit_pref[unit_base[s + 1]] - unit_pref[unit_base[s]];
		g.recs[r0 + s].gstart = gp;
		g.recs[r0 + s].len = len;
		seg_gbase[s] = (int64_t)gp - (int64_t)unit_pref[unit_base[s]];
		gp += len;
		uint64_t end = (gp + 63) & ~(uint64_t)63;
		if (end == gp) end += 64;
		pad_lo.push_back(gp);
		pad_hi.push_back(end);
		gp = end;
	}
	g.total_bases = gp;
	const size_t bw = (size_t)((gp + 15) / 16 + 4), mw = (size_t)((gp + 31) / 32 + 4);
	if (!presized) {
		CK(h->d_bases.ensure(bw * 4));
		CK(h->d_nmask.ensure(mw * 4));
		CK(cudaMemsetAsync(h->d_bases.p, 0, bw * 4, h->st));
		CK(cudaMemsetAsync(h->d_nmask.p, 0, mw * 4, h->st));
	} else if (gp + 64 > h->stream_cap_bases) {
		return h->fail(QTIPSIM_ERANGE, "streaming loader: more record padding than reserved (too many records)");
	}
	if (nunits) {
		CK(h->d_pk_a.ensure((size_t)nseg * 8 + 16));
		CK(ctl_h2d(h, h->d_pk_a.p, seg_gbase.data(), nseg * 8));
		uint64_t *d_ss = h->d_pk_b.as<uint64_t>(), *d_se = d_ss + nseg, *d_ub = d_se + nseg;
		launch_pack_units(h->d_text.as<uint8_t>(), d_ss, d_se, d_ub, nseg, nunits, h->d_pk_d.as<uint64_t>(),
		                  h->d_pk_a.as<int64_t>(), h->d_bases.as<uint32_t>(), h->d_nmask.as<uint32_t>(), h->st);
	}
	if (!pad_lo.empty()) {
		const uint32_t np = (uint32_t)pad_lo.size();
		CK(h->d_pk_c.ensure((size_t)np * 16 + 16));
		uint64_t *d_lo = h->d_pk_c.as<uint64_t>(), *d_hi = d_lo + np;
		CK(ctl_h2d(h, d_lo, pad_lo.data(), np * 8));
		CK(ctl_h2d(h, d_hi, pad_hi.data(), np * 8));
		launch_fill_n(d_lo, d_hi, np, h->d_nmask.as<uint32_t>(), h->st);
	}
	{
		const int rc = build_nblock(h, mw);
		if (rc) return rc;
	}
	g.fill_name16();
	CK(h->d_recs.ensure(g.recs.size() * sizeof(RecDev) + 16));
	CK(h->d_names.ensure(g.names.size() + 16));
	if (!g.recs.empty())
		CK(ctl_h2d(h, h->d_recs.p, g.recs.data(), g.recs.size() * sizeof(RecDev)));
	if (!g.names.empty())
		CK(ctl_h2d(h, h->d_names.p, g.names.data(), g.names.size()));
	h->recs_packed = r1;
	h->win_cache_ok = false;
	h->setup_cache_ok = false;
	return QTIPSIM_OK;
}

// One-shot loader: `text` is the FASTA text in host memory.
static int device_pack(qtipsim *h, const uint8_t *text, uint64_t n) {
	CK(cudaSetDevice(h->device));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	CK(cudaEventRecord(e0, h->st));
	genome_reset(h, n);
	CK(h->d_text.ensure(n + 64));
	if (n) CK(cudaMemcpyAsync(h->d_text.p, text, n, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemsetAsync((char *)h->d_text.p + n, ' ', 64, h->st));
	// ---- header lines: every '>' found by the GPU, resolved against the host text ----
	std::vector<uint64_t> gts;
	int rc = find_headers_chunk(h, text, 0, n, gts);
	if (rc) return rc;
	resolve_headers(h, text, n, gts, true);
	if (h->names_overflow) return h->fail(QTIPSIM_ERANGE, "FASTA record names exceed 4 GiB in total");
	rc = pack_records(h, false);
	if (rc) return rc;
	CK(cudaEventRecord(e1, h->st));
	CK(ctl_sync(h));
	CK(cudaGetLastError());
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	h->ms_pack_pending = ms;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	h->have_genome = true;
	return QTIPSIM_OK;
}

// The host replay planner reads reference bases at X positions: fetch the packed genome back once.
static int ensure_host_packed(qtipsim *h) {
	if (h->host_packed) return QTIPSIM_OK;
	const size_t bw = (size_t)((h->hg.total_bases + 15) / 16 + 4), mw = (size_t)((h->hg.total_bases + 31) / 32 + 4);
	h->hg.bases.resize(bw);
	h->hg.nmask.resize(mw);
	CK(cudaMemcpyAsync(h->hg.bases.data(), h->d_bases.p, bw * 4, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(h->hg.nmask.data(), h->d_nmask.p, mw * 4, cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	h->host_packed = true;
	return QTIPSIM_OK;
}

static bool use_host_pack() {
	const char *e = getenv("QTIPSIM_HOST_PACK");
	return e && e[0] == '1';
}

int qtipsim_load_fasta_bytes(qtipsim_t *h, const void *bytes, size_t nbytes) {
	if (!h || (!bytes && nbytes)) return QTIPSIM_EINVAL;
	if (!use_host_pack()) return device_pack(h, (const uint8_t *)bytes, nbytes);
	if (h->feeder.joinable()) h->feeder.join();
	h->stream_open = false;
	h->hg.parse((const uint8_t *)bytes, nbytes);
	return upload_genome(h);
}

int qtipsim_load_fasta(qtipsim_t *h, const char *const *paths, size_t npaths) {
	if (!h || !paths) return QTIPSIM_EINVAL;
	if (npaths != 1)
		return h->fail(QTIPSIM_EINVAL, "exactly one FASTA file is supported (the reference crashes on more, SURVEY Q13)");
	FILE *fh = fopen(paths[0], "rb");
	if (!fh) return h->fail(QTIPSIM_EIO, std::string("Could not open FASTA file \"") + paths[0] + "\"");
	std::vector<uint8_t> buf;
	if (fseek(fh, 0, SEEK_END) == 0) {
		long sz = ftell(fh);
		fseek(fh, 0, SEEK_SET);
		if (sz > 0) buf.resize((size_t)sz);
	}
	size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), fh);
	const bool bad = ferror(fh) || got != buf.size();
	fclose(fh);
	if (bad) return h->fail(QTIPSIM_EIO, std::string("read error on FASTA file \"") + paths[0] + "\"");
	if (!use_host_pack()) return device_pack(h, buf.data(), buf.size());
	h->hg.parse(buf.data(), buf.size());
	return upload_genome(h);
}

uint64_t qtipsim_num_estimated_bases(const qtipsim_t *h) { return h ? h->hg.file_bytes : 0; }
uint64_t qtipsim_num_records(const qtipsim_t *h) { return h ? h->hg.recs.size() : 0; }

int qtipsim_genome_export(qtipsim_t *h, int which, void **ptr, uint64_t *nbytes, int *is_device) {
	if (!h || !ptr || !nbytes) return QTIPSIM_EINVAL;
	if (!h->have_genome) return h->fail(QTIPSIM_ESTATE, "no genome loaded");
	static thread_local std::string blob;
	switch (which) {
	case 0:
		*ptr = h->d_bases.p;
		*nbytes = (h->hg.total_bases + 15) / 16 * 4 + 16;
		if (is_device) *is_device = 1;
		return QTIPSIM_OK;
	case 1:
		*ptr = h->d_nmask.p;
		*nbytes = (h->hg.total_bases + 31) / 32 * 4 + 16;
		if (is_device) *is_device = 1;
		return QTIPSIM_OK;
	case 2:
		blob = h->hg.blob();
		*ptr = (void *)blob.data();
		*nbytes = blob.size();
		if (is_device) *is_device = 0;
		return QTIPSIM_OK;
	}
	return QTIPSIM_EINVAL;
}

int qtipsim_genome_adopt(qtipsim_t *h, const void *blob, uint64_t blob_bytes, const void *d_bases,
                         uint64_t bases_bytes, const void *d_nmask, uint64_t nmask_bytes) {
	if (!h || !blob || !d_bases || !d_nmask) return QTIPSIM_EINVAL;
	CK(cudaSetDevice(h->device));
	HostGenome incoming;  // swapped in only after every check has passed: a failed adopt leaves the handle as it was
	if (!incoming.from_blob(blob, blob_bytes)) return h->fail(QTIPSIM_EINVAL, "bad genome blob");
	const uint64_t need_b = (incoming.total_bases + 15) / 16 * 4 + 16, need_n = (incoming.total_bases + 31) / 32 * 4 + 16;
	if (bases_bytes < need_b || nmask_bytes < need_n) return h->fail(QTIPSIM_EINVAL, "packed genome buffers too small");
	genome_reset(h, incoming.file_bytes);  // joins a pending streaming upload and clears the streaming / header-scan state
	h->hg = incoming;
	h->hg.fill_name16();
	CK(h->d_bases.ensure(need_b));
	CK(h->d_nmask.ensure(need_n));
	CK(h->d_recs.ensure(h->hg.recs.size() * sizeof(RecDev) + 16));
	CK(h->d_names.ensure(h->hg.names.size() + 16));
	CK(cudaMemcpyAsync(h->d_bases.p, d_bases, need_b, cudaMemcpyDeviceToDevice, h->st));
	CK(cudaMemcpyAsync(h->d_nmask.p, d_nmask, need_n, cudaMemcpyDeviceToDevice, h->st));
	{
		const int rc = build_nblock(h, need_n / 4);
		if (rc) return rc;
	}
	CK(ctl_h2d(h, h->d_recs.p, h->hg.recs.data(), h->hg.recs.size() * sizeof(RecDev)));
	if (!h->hg.names.empty())
		CK(ctl_h2d(h, h->d_names.p, h->hg.names.data(), h->hg.names.size()));
	// (the host replay planner fetches a host copy of the packed arrays on demand: ensure_host_packed)
	h->hg.bases.clear();
	h->hg.nmask.clear();
	CK(ctl_sync(h));
	h->have_genome = true;
	h->win_cache_ok = false;
	h->setup_cache_ok = false;
	h->host_packed = false;
	h->have_result = false;
	return QTIPSIM_OK;
}

// ---- packed-genome cache: [header][record-table blob][bases][nmask] ----
namespace {
struct CacheHeader {
	char magic[8];  // "QTPG\1\0\0\0"
	uint64_t fasta_size;
	int64_t fasta_mtime_ns;
	uint64_t blob_bytes, bases_bytes, nmask_bytes;
};
const char CACHE_MAGIC[8] = {'Q', 'T', 'P', 'G', 1, 0, 0, 0};
bool fasta_stamp(const char *path, uint64_t &size, int64_t &mtime_ns) {
	struct stat sb;
	if (stat(path, &sb) != 0) return false;
	size = (uint64_t)sb.st_size;
	mtime_ns = (int64_t)sb.st_mtim.tv_sec * 1000000000ll + (int64_t)sb.st_mtim.tv_nsec;
	return true;
}
}  // namespace

static int ensure_stage(qtipsim *h) {
	const size_t STAGE = 32u << 20;
	if (!h->h_stage) {
		CK(cudaHostAlloc(&h->h_stage, 2 * STAGE, cudaHostAllocDefault));
		h->h_stage_cap = 2 * STAGE;
	}
	return QTIPSIM_OK;
}

int qtipsim_genome_save(qtipsim_t *h, const char *cache_path, const char *fasta_path) {
	if (!h || !cache_path || !fasta_path) return QTIPSIM_EINVAL;
	if (!h->have_genome) return h->fail(QTIPSIM_ESTATE, "genome_save: no genome loaded");
	CK(cudaSetDevice(h->device));
	CacheHeader hd;
	memcpy(hd.magic, CACHE_MAGIC, 8);
	if (!fasta_stamp(fasta_path, hd.fasta_size, hd.fasta_mtime_ns)) return h->fail(QTIPSIM_EIO, std::string("genome_save: cannot stat ") + fasta_path);
	const std::string blob = h->hg.blob();
	hd.blob_bytes = blob.size();
	hd.bases_bytes = (h->hg.total_bases + 15) / 16 * 4 + 16;
	hd.nmask_bytes = (h->hg.total_bases + 31) / 32 * 4 + 16;
	int rc = ensure_stage(h);
	if (rc) return rc;
	const std::string tmp = std::string(cache_path) + ".tmp";
	FILE *fh = fopen(tmp.c_str(), "wb");
	if (!fh) return h->fail(QTIPSIM_EIO, std::string("genome_save: cannot write ") + tmp);
	bool ok = fwrite(&hd, sizeof(hd), 1, fh) == 1 && (blob.empty() || fwrite(blob.data(), blob.size(), 1, fh) == 1);
	const size_t STAGE = 32u << 20;
	const void *src[2] = {h->d_bases.p, h->d_nmask.p};
	const uint64_t len[2] = {hd.bases_bytes, hd.nmask_bytes};
	for (int a = 0; a < 2 && ok; a++)
		for (uint64_t off = 0; off < len[a] && ok; off += STAGE) {
			const size_t n = (size_t)std::min<uint64_t>(STAGE, len[a] - off);
			if (cudaMemcpyAsync(h->h_stage, (const char *)src[a] + off, n, cudaMemcpyDeviceToHost, h->st) != cudaSuccess || ctl_sync(h) != cudaSuccess) ok = false;
			else ok = fwrite(h->h_stage, n, 1, fh) == 1;
		}
	ok = (fclose(fh) == 0) && ok;
	if (!ok || rename(tmp.c_str(), cache_path) != 0) {
		remove(tmp.c_str());
		cudaGetLastError();
		return h->fail(QTIPSIM_EIO, std::string("genome_save: writing ") + cache_path + " failed");
	}
	return QTIPSIM_OK;
}

// uploads a cache file whose stamp matches; false (with nothing changed on the handle's genome state) otherwise
static int load_cache(qtipsim *h, const char *cache_path, uint64_t fsize, int64_t fmtime, bool &hit) {
	hit = false;
	FILE *fh = fopen(cache_path, "rb");
	if (!fh) return QTIPSIM_OK;
	CacheHeader hd;
	bool ok = fread(&hd, sizeof(hd), 1, fh) == 1 && memcmp(hd.magic, CACHE_MAGIC, 8) == 0 && hd.fasta_size == fsize && hd.fasta_mtime_ns == fmtime &&
	          hd.blob_bytes < (1ull << 32);
	std::string blob;
	if (ok) {
		blob.resize((size_t)hd.blob_bytes);
		ok = blob.empty() || fread(&blob[0], blob.size(), 1, fh) == 1;
	}
	HostGenome tmp;
	ok = ok && tmp.from_blob(blob.data(), blob.size()) && hd.bases_bytes == (tmp.total_bases + 15) / 16 * 4 + 16 &&
	     hd.nmask_bytes == (tmp.total_bases + 31) / 32 * 4 + 16;
	if (!ok) {
		fclose(fh);
		return QTIPSIM_OK;  // absent, foreign or stale: the caller packs the FASTA
	}
	int rc = ensure_stage(h);
	if (rc) {
		fclose(fh);
		return rc;
	}
	genome_reset(h, fsize);
	if (h->d_bases.ensure(hd.bases_bytes) != cudaSuccess || h->d_nmask.ensure(hd.nmask_bytes) != cudaSuccess) {
		fclose(fh);
		return h->fail(QTIPSIM_ENOMEM, "genome cache: device allocation failed");
	}
	const size_t STAGE = 32u << 20;
	void *dst[2] = {h->d_bases.p, h->d_nmask.p};
	const uint64_t len[2] = {hd.bases_bytes, hd.nmask_bytes};
	cudaEvent_t ev[2];
	cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
	int slot = 0;
	bool used[2] = {false, false};
	for (int a = 0; a < 2 && ok; a++)
		for (uint64_t off = 0; off < len[a] && ok; off += STAGE) {  // double-buffered: read slice k + 1 while slice k uploads
			const size_t n = (size_t)std::min<uint64_t>(STAGE, len[a] - off);
			char *hp = (char *)h->h_stage + (size_t)slot * STAGE;
			if (used[slot]) cudaEventSynchronize(ev[slot]);
			ok = fread(hp, n, 1, fh) == 1 && cudaMemcpyAsync((char *)dst[a] + off, hp, n, cudaMemcpyHostToDevice, h->st) == cudaSuccess;
			cudaEventRecord(ev[slot], h->st);
			used[slot] = true;
			slot ^= 1;
		}
	fclose(fh);
	cudaStreamSynchronize(h->st);
	cudaEventDestroy(ev[0]);
	cudaEventDestroy(ev[1]);
	if (!ok) {
		cudaGetLastError();
		return h->fail(QTIPSIM_EIO, std::string("genome cache: reading ") + cache_path + " failed");
	}
	h->hg = tmp;
	h->hg.fill_name16();
	CK(h->d_recs.ensure(h->hg.recs.size() * sizeof(RecDev) + 16));
	CK(h->d_names.ensure(h->hg.names.size() + 16));
	rc = build_nblock(h, (size_t)(hd.nmask_bytes / 4));
	if (rc) return rc;
	CK(ctl_h2d(h, h->d_recs.p, h->hg.recs.data(), h->hg.recs.size() * sizeof(RecDev)));
	if (!h->hg.names.empty()) CK(ctl_h2d(h, h->d_names.p, h->hg.names.data(), h->hg.names.size()));
	CK(ctl_sync(h));
	h->have_genome = true;
	hit = true;
	return QTIPSIM_OK;
}

int qtipsim_load_fasta_cached(qtipsim_t *h, const char *fasta_path, const char *cache_path, int *hit) {
	if (!h || !fasta_path || !cache_path) return QTIPSIM_EINVAL;
	if (hit) *hit = 0;
	CK(cudaSetDevice(h->device));
	uint64_t fsize = 0;
	int64_t fmtime = 0;
	if (!fasta_stamp(fasta_path, fsize, fmtime)) return h->fail(QTIPSIM_EIO, std::string("cannot open FASTA ") + fasta_path);
	bool got = false;
	int rc = load_cache(h, cache_path, fsize, fmtime, got);
	if (rc) return rc;
	if (got) {
		if (hit) *hit = 1;
		return QTIPSIM_OK;
	}
	const char *paths[1] = {fasta_path};
	rc = qtipsim_load_fasta(h, paths, 1);
	if (rc) return rc;
	return qtipsim_genome_save(h, cache_path, fasta_path);
}

int qtipsim_set_templates_unpaired(qtipsim_t *h, int cat, const qtipsim_tmpl_unp *v, size_t kept, uint64_t added) {
	if (!h || (!v && kept)) return QTIPSIM_EINVAL;
	int rc = h->ts.set_unpaired(cat, v, kept, added, h->err);
	if (rc == QTIPSIM_OK) h->tmpl_dirty[cat] = true;
	return rc;
}

int qtipsim_set_templates_paired(qtipsim_t *h, int cat, const qtipsim_tmpl_pair *v, size_t kept, uint64_t added) {
	if (!h || (!v && kept)) return QTIPSIM_EINVAL;
	int rc = h->ts.set_paired(cat, v, kept, added, h->err);
	if (rc == QTIPSIM_OK) h->tmpl_dirty[cat] = true;
	return rc;
}

uint64_t qtipsim_olap(const qtipsim_t *h) { return h ? h->ts.olap() : 0; }

// ---- the input model's reservoirs fed with plain arrays (model.h; qtip_parse.cpp:849-861, 1148-1166, ds.h:437-446) ----

int qtipsim_model_begin(qtipsim_t *h, const uint64_t k[4]) {
	if (!h || !k) return QTIPSIM_EINVAL;
	h->model.begin(k);
	return QTIPSIM_OK;
}

int qtipsim_model_add(qtipsim_t *h, const qtipsim_model_batch *batch, int32_t *cg1, int32_t *cg2) {
	if (!h || !batch || !cg1 || !cg2) return QTIPSIM_EINVAL;
	if (*cg1 < 1 || (uint32_t)*cg1 >= LM1 || *cg2 < 1 || (uint32_t)*cg2 >= LM2)
		return h->fail(QTIPSIM_ERANGE, "rnglib state out of range (set_seed, rnglib.cpp:1714-1728)");
	LState rng{(uint32_t)*cg1, (uint32_t)*cg2};
	const int rc = h->model.add(*batch, rng, h->err);
	if (rc) return rc;
	*cg1 = (int32_t)rng.s1;
	*cg2 = (int32_t)rng.s2;
	return QTIPSIM_OK;
}

int qtipsim_model_commit(qtipsim_t *h, uint64_t kept[4], uint64_t added[4]) {
	if (!h) return QTIPSIM_EINVAL;
	for (int c = 0; c < NCAT; c++) {
		if (kept) kept[c] = h->model.kept[c].size();
		if (added) added[c] = h->model.n[c];
	}
	const int rc = h->model.commit(h->ts, h->err);
	if (rc) return rc;
	for (int c = 0; c < NCAT; c++) h->tmpl_dirty[c] = true;
	return QTIPSIM_OK;
}

static int upload_templates(qtipsim *h) {
	for (int c = 0; c < NCAT; c++) {
		if (!h->tmpl_dirty[c]) continue;
		const Category &cg = h->ts.cat[c];
		CK(h->d_t[c].ensure(cg.t.size() * sizeof(TmplDesc) + 16));
		CK(h->d_qual[c].ensure(cg.qual.size() + 16));
		CK(h->d_xs[c].ensure(cg.xs.size() + 16));
		if (!cg.t.empty()) {
			CK(ctl_h2d(h, h->d_t[c].p, cg.t.data(), cg.t.size() * sizeof(TmplDesc)));
			CK(ctl_h2d(h, h->d_qual[c].p, cg.qual.data(), cg.qual.size()));
			CK(ctl_h2d(h, h->d_xs[c].p, cg.xs.data(), cg.xs.size()));
			// compiled form; the quality arena is read with aligned 4-byte loads up to 16 bytes past an entry
			CK(h->d_qual_out[c].ensure(cg.qual_out.size() + 64));
			CK(h->d_ops[c].ensure(cg.ops.size() * 4 + 64));
			CK(h->d_segs[c].ensure(cg.segs.size() * sizeof(SegDev) + 16));
			CK(cudaMemsetAsync(h->d_qual_out[c].p, 0, cg.qual_out.size() + 64, h->st));
			CK(ctl_h2d(h, h->d_qual_out[c].p, cg.qual_out.data(), cg.qual_out.size()));
			if (!cg.ops.empty())
				CK(ctl_h2d(h, h->d_ops[c].p, cg.ops.data(), cg.ops.size() * 4));
			if (!cg.segs.empty())
				CK(ctl_h2d(h, h->d_segs[c].p, cg.segs.data(), cg.segs.size() * sizeof(SegDev)));
			// draw programs for the device-side replay resolver
			CK(h->d_xsteps[c].ensure(cg.xsteps.size() * sizeof(XStep) + 16));
			CK(h->d_prog[c].ensure(cg.prog.size() * sizeof(HostMateProg) + 16));
			if (!cg.xsteps.empty())
				CK(ctl_h2d(h, h->d_xsteps[c].p, cg.xsteps.data(), cg.xsteps.size() * sizeof(XStep)));
			if (!cg.prog.empty())
				CK(ctl_h2d(h, h->d_prog[c].p, cg.prog.data(), cg.prog.size() * sizeof(HostMateProg)));
			CK(h->d_xprog[c].ensure(cg.xprog.size() * sizeof(XProgRec) + 64));
			if (!cg.xprog.empty())
				CK(ctl_h2d(h, h->d_xprog[c].p, cg.xprog.data(), cg.xprog.size() * sizeof(XProgRec)));
		}
		h->tmpl_dirty[c] = false;
	}
	if (!h->have_jt) {
		std::vector<uint32_t> j1(JUMP_TBL), j2(JUMP_TBL);
		j1[0] = j2[0] = 1;
		for (int i = 1; i < JUMP_TBL; i++) {
			j1[i] = mulmod1(j1[i - 1], LA1);
			j2[i] = mulmod2(j2[i - 1], LA2);
		}
		CK(h->d_jt1.ensure(JUMP_TBL * 4));
		CK(h->d_jt2.ensure(JUMP_TBL * 4));
		CK(ctl_h2d(h, h->d_jt1.p, j1.data(), JUMP_TBL * 4));
		CK(ctl_h2d(h, h->d_jt2.p, j2.data(), JUMP_TBL * 4));
		// a^(4096 k): the high digit of a two-level jump (resolve2.cu)
		std::vector<uint32_t> g1(JUMP_TBL), g2(JUMP_TBL);
		const uint32_t b1 = mulmod1(j1[JUMP_TBL - 1], LA1), b2 = mulmod2(j2[JUMP_TBL - 1], LA2);
		g1[0] = g2[0] = 1;
		for (int i = 1; i < JUMP_TBL; i++) {
			g1[i] = mulmod1(g1[i - 1], b1);
			g2[i] = mulmod2(g2[i - 1], b2);
		}
		CK(h->d_jh1.ensure(JUMP_TBL * 4));
		CK(h->d_jh2.ensure(JUMP_TBL * 4));
		CK(ctl_h2d(h, h->d_jh1.p, g1.data(), JUMP_TBL * 4));
		CK(ctl_h2d(h, h->d_jh2.p, g2.data(), JUMP_TBL * 4));
		CK(ctl_sync(h));  // the vectors are stack-lifetime
		h->have_jt = true;
	}
	return QTIPSIM_OK;
}

// Binomial set-ups (ignbin's (n, p)-dependent constants, computed with the host's libm) for every
// distinct window probability in [lo, hi); pclass[w] = row of the table for window w.
static void window_setups(const std::vector<Window> &win, uint64_t olap, const PlanParams &pp, uint32_t lo, uint32_t hi,
                          std::vector<uint32_t> &pclass, std::vector<BinomSetup> &setups) {
	std::map<uint32_t, uint32_t> pclass_of;  // float bits of p -> class
	pclass.assign(win.size(), 0);
	setups.clear();
	uint32_t last_retsz = 0xffffffffu, last_class = 0;
	for (uint32_t w = lo; w < hi; w++) {
		if (win[w].retsz < olap) continue;
		if (win[w].retsz == last_retsz) {  // nearly every window is a full one
			pclass[w] = last_class;
			continue;
		}
		const float bp = window_binom_p(win[w].retsz, olap, pp.tot_fasta_len);
		uint32_t pb;
		memcpy(&pb, &bp, 4);
		auto it = pclass_of.find(pb);
		if (it == pclass_of.end()) {
			it = pclass_of.emplace(pb, (uint32_t)pclass_of.size()).first;
			for (int c = 0; c < NCAT; c++) setups.push_back(binom_setup((int32_t)pp.target[c], bp));
		}
		pclass[w] = it->second;
		last_retsz = win[w].retsz;
		last_class = it->second;
	}
	if (setups.empty()) setups.resize(NCAT);
}

// window classes + binomial set-ups on the device (d_pclass, d_setups), recomputed only when their inputs change
static int ensure_setups(qtipsim *h, const std::vector<Window> &win, uint64_t olap, const PlanParams &pp, uint32_t lo, uint32_t hi) {
	const uint64_t key[8] = {pp.target[0], pp.target[1], pp.target[2], pp.target[3], olap, pp.chunksz, ((uint64_t)lo << 32) | hi,
	                         pp.tot_fasta_len};
	if (h->setup_cache_ok && memcmp(key, h->setup_key, sizeof(key)) == 0) return QTIPSIM_OK;
	window_setups(win, olap, pp, lo, hi, h->pclass_cache, h->setup_cache);
	const size_t nwin = win.size();
	CK(h->d_setups.ensure(h->setup_cache.size() * sizeof(BinomSetup)));
	h->tab_cache.resize(h->setup_cache.size());
	for (size_t i = 0; i < h->setup_cache.size(); i++) binom_tables(h->setup_cache[i], h->tab_cache[i]);
	CK(h->d_tabs.ensure(h->tab_cache.size() * sizeof(BinomTables)));
	CK(ctl_h2d(h, h->d_tabs.p, h->tab_cache.data(), h->tab_cache.size() * sizeof(BinomTables)));
	CK(h->d_pclass.ensure(nwin * 4 + 16));
	CK(ctl_h2d(h, h->d_setups.p, h->setup_cache.data(), h->setup_cache.size() * sizeof(BinomSetup)));
	if (nwin) CK(ctl_h2d(h, h->d_pclass.p, h->pclass_cache.data(), nwin * 4));
	CK(ctl_sync(h));  // the cache vectors may be rebuilt by the next call
	memcpy(h->setup_key, key, sizeof(key));
	h->setup_cache_ok = true;
	return QTIPSIM_OK;
}

static void fill_resolve_args(qtipsim *h, const qtipsim_run_params *p, const PlanParams &pp, const DevGenome &dg, uint32_t nwin,
                              uint64_t olap, ResolveArgs &ra) {
	ra.g = dg;
	ra.win = h->d_win.as<Window>();
	ra.ncount = h->d_ncount.as<uint32_t>();
	ra.pclass = h->d_pclass.as<uint32_t>();
	ra.setups = h->d_setups.as<BinomSetup>();
	ra.nwin = nwin;
	uint32_t mr = 0;
	for (int c = 0; c < NCAT; c++) mr = std::max(mr, h->ts.cat[c].max_rflen);
	ra.n_skip_ok = p->chunksz - olap >= (uint64_t)mr;
	ra.win_lo = (uint32_t)pp.win_lo;
	ra.win_hi = (uint32_t)pp.win_hi;
	ra.olap = olap;
	for (int c = 0; c < NCAT; c++) {
		ra.t[c] = h->d_t[c].as<TmplDesc>();
		ra.xsteps[c] = h->d_xsteps[c].as<XStep>();
		ra.prog[c] = h->d_prog[c].as<HostMateProg>();
		ra.xp[c] = getenv("QTIPSIM_R2_NOXPROG") ? nullptr : h->d_xprog[c].as<XProgRec>();
		ra.ntmpl[c] = (uint32_t)h->ts.cat[c].t.size();
		ra.target[c] = pp.target[c];
		ra.cap[c] = 0;
		ra.plan[c] = nullptr;
	}
	ra.jt1 = h->d_jt1.as<uint32_t>();
	ra.jt2 = h->d_jt2.as<uint32_t>();
	ra.s1 = (uint32_t)p->cg1;
	ra.s2 = (uint32_t)p->cg2;
	ra.result = nullptr;
}

// K2, second generation (resolve2.cu): stages of band-speculated windows, then counts, offsets and the plan records.
// done = false (with rc OK): the resolver declined or got stuck -- the caller runs the first-generation kernel, which
// also produces the reference's error for the window in question.
static int resolve_replay_band(qtipsim *h, qtipsim_run_params *p, const PlanParams &pp, const DevGenome &dg, uint32_t nwin, uint64_t olap,
                               uint64_t nreads[NCAT], uint64_t &draws, bool &done) {
	done = false;
	const std::vector<Window> &win = h->win_cache;
	const std::vector<uint32_t> &ncount = h->ncount_cache;
	// ---- the shard's live windows and the statistics of the draws each one consumes ----
	std::vector<uint32_t> aw;
	std::vector<double> cum;  // [0 .. na]: mean, [na + 1 .. 2 na + 1]: variance
	for (uint64_t w = pp.win_lo; w < pp.win_hi; w++)
		if (window_live(win[w].retsz, olap, ncount[w])) aw.push_back((uint32_t)w);
	const uint32_t na = (uint32_t)aw.size();
	Resolve2Args a2;
	memset((void *)&a2, 0, sizeof(a2));
	fill_resolve_args(h, p, pp, dg, nwin, olap, a2.ra);
	for (int c = 0; c < NCAT; c++) {
		const Category &cg = h->ts.cat[c];
		a2.m1[c] = (float)cg.draw_m1;
		a2.v1[c] = (float)cg.draw_v1;
		a2.maxfixed[c] = cg.draw_max_fixed;
		if (cg.draw_max_fixed > 16000u) return QTIPSIM_OK;  // a read of this many draws does not fit the u16 tables: first-generation kernel
	}
	cum.assign(2 * (size_t)na + 2, 0.0);
	double typical = 0, vmax = 0, reach_max = 0;
	for (uint32_t i = 0; i < na; i++) {
		const BinomSetup *su = &h->setup_cache[(size_t)h->pclass_cache[aw[i]] * NCAT];
		double m = 0, v = 0, big = 0;
		for (int c = 0; c < NCAT; c++) {
			const double np = (double)su[c].n * (double)su[c].pp, q = 1.0 - (double)su[c].pp;
			m += 1.5 + np * a2.m1[c];
			v += np * q * (double)a2.m1[c] * a2.m1[c] + np * a2.v1[c] + 2.0;
			big = std::max(big, np * a2.m1[c]);
		}
		cum[i + 1] = cum[i] + m;
		cum[na + 1 + i + 1] = cum[na + 1 + i] + v;
		typical = std::max(typical, big);
		vmax = std::max(vmax, v);
		for (int c = 0; c < NCAT; c++) {
			const double np = (double)su[c].n * (double)su[c].pp;
			reach_max = std::max(reach_max, np * a2.m1[c] + 8.0 * sqrt(np * (1.0 - su[c].pp) * a2.m1[c] * a2.m1[c] + np * a2.v1[c] + 1.0) +
			                                    2.0 * a2.maxfixed[c] + 64.0);
		}
	}
	auto envu = [](const char *nm, uint32_t dflt) {
		const char *e = getenv(nm);
		return e && *e ? (uint32_t)strtoul(e, nullptr, 10) : dflt;
	};
	int nsm = 148;
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
	a2.rcap = std::min<uint32_t>(R2_RCAP_MAX, std::max<uint32_t>(1024u, envu("QTIPSIM_R2_RCAP", R2_RCAP_MAX)));
	a2.rcap_emit = std::min<uint32_t>(R2_RCAP_MAX, std::max<uint32_t>(1024u, envu("QTIPSIM_R2_RCAP_EMIT", 8192u)));
	a2.nbmax = std::max<uint32_t>(64u, envu("QTIPSIM_R2_NBMAX", 24576u));
	{
		const char *e = getenv("QTIPSIM_R2_SIGMA");
		// 3 sigma: a stage's work grows with the band width, and a miss only ends the stage early (the window restarts the
		// next one exactly): measured at C3, 4 sigma 34.6 ms / 162 stages, 3 sigma 30.3 ms / 162-170, 2.5 sigma 30.0 ms / 176-182,
		// 2 sigma 34.0 ms / 215-234 (profiles/r02_k_resolve2_ncu_summary.md)
		a2.band_sigma = e && *e ? (float)atof(e) : 3.0f;
	}
	uint32_t K = std::min<uint32_t>(R2_KMAX, std::max<uint32_t>(1u, envu("QTIPSIM_R2_K", (uint32_t)nsm)));
	// How many windows ahead can be speculated before the band (which widens with the square root of the distance)
	// plus one window's reads outgrow the largest dense table?  Workloads where that is only a handful -- many reads per
	// window, or bad-end reads whose ~150 skipped draws each make the position noise large -- are left to the
	// first-generation resolver, which walks them round by round on a cluster (QTIPSIM_R2_FORCE=1 keeps them here: tests).
	double kfit = 0;
	{
		const double room = (double)a2.rcap - reach_max * 1.1 - 64.0;
		if (room > 0 && vmax > 0) {
			const double r = (room / 2.2 - 16.0) / ((double)a2.band_sigma * sqrt(vmax));
			kfit = r > 0 ? r * r + 1.0 : 1.0;
		}
		if (kfit < (double)K) K = kfit < 1.0 ? 1u : (uint32_t)kfit;
		if (K < 16 && !getenv("QTIPSIM_R2_FORCE") && !getenv("QTIPSIM_R2_K")) return QTIPSIM_OK;
	}
	if (typical + 64.0 > (double)a2.rcap) K = 1;
	if (na && K > na) K = na;
	if (K > 1 && !getenv("QTIPSIM_R2_RCAP")) {
		// the dense tables only need to hold the widest band of a stage plus one window's reads: a smaller table leaves
		// more of the SM's shared memory to the L1 cache, where the window's packed bases live
		const double nbw = 2.0 * ((double)a2.band_sigma * sqrt(vmax * (double)(K - 1)) + 16.0) + 2.0;
		const double need = (std::min(nbw, (double)a2.nbmax) + reach_max) * 1.1;
		uint32_t rc = (uint32_t)std::min<double>(need, (double)R2_RCAP_MAX);
		rc = (rc + 1023u) & ~1023u;
		a2.rcap = std::min<uint32_t>(R2_RCAP_MAX, std::max<uint32_t>(4096u, rc));
	}
	if (!na) {
		for (int c = 0; c < NCAT; c++) nreads[c] = 0;
		draws = 0;
		done = true;
		return QTIPSIM_OK;
	}
	const uint32_t max_launch = na + 8;
	CK(h->d_r2_aw.ensure((size_t)na * 4));
	CK(h->d_r2_cum.ensure(cum.size() * 8));
	CK(h->d_r2_anchor.ensure(((size_t)max_launch + 1) * sizeof(StageAnchor)));
	CK(h->d_r2_counter.ensure((size_t)2 * max_launch * 4));
	CK(h->d_r2_slot.ensure((size_t)K * sizeof(uint2)));
	CK(h->d_r2_relay.ensure(((size_t)K + 1) * 8));
	CK(cudaMemsetAsync(h->d_r2_relay.p, 0, ((size_t)K + 1) * 8, h->st));
	CK(h->d_r2_end.ensure((size_t)NCAT * K * a2.nbmax * 4));
	CK(h->d_r2_cn.ensure((size_t)2 * K * a2.nbmax * 4));
	CK(h->d_r2_wres.ensure((size_t)na * sizeof(WinResolved)));
	CK(h->d_r2_cnt.ensure((size_t)2 * NCAT * na * 4));
	CK(h->d_r2_off.ensure((size_t)NCAT * ((size_t)na + 1) * 8));
	CK(h->d_r2_misc.ensure(64));
	CK(h->d_scan_tmp.ensure(((size_t)na / 1024 + 4) * 8));
	CK(ctl_h2d(h, h->d_r2_aw.p, aw.data(), (size_t)na * 4));
	CK(ctl_h2d(h, h->d_r2_cum.p, cum.data(), cum.size() * 8));
	CK(cudaMemsetAsync(h->d_r2_counter.p, 0, (size_t)2 * max_launch * 4, h->st));
	CK(cudaMemsetAsync(h->d_r2_misc.p, 0, 64, h->st));
	StageAnchor a0;
	memset(&a0, 0, sizeof(a0));
	a0.s1 = (uint32_t)p->cg1;
	a0.s2 = (uint32_t)p->cg2;
	a0.status = R2_RUNNING;
	CK(ctl_h2d(h, h->d_r2_anchor.p, &a0, sizeof(a0)));
	a2.jh1 = h->d_jh1.as<uint32_t>();
	a2.jh2 = h->d_jh2.as<uint32_t>();
	a2.tabs = h->d_tabs.as<BinomTables>();
	a2.aw = h->d_r2_aw.as<uint32_t>();
	a2.na = na;
	a2.cumM = h->d_r2_cum.as<double>();
	a2.cumV = h->d_r2_cum.as<double>() + na + 1;
	a2.anchor = h->d_r2_anchor.as<StageAnchor>();
	a2.counter = h->d_r2_counter.as<uint32_t>();
	a2.ticket = a2.counter + max_launch;
	a2.slot = h->d_r2_slot.as<uint2>();
	a2.relay = h->d_r2_relay.as<unsigned long long>();
	a2.end = h->d_r2_end.as<uint32_t>();
	a2.cn = h->d_r2_cn.as<uint32_t>();
	a2.cnd = a2.cn + (size_t)K * a2.nbmax;
	a2.wres = h->d_r2_wres.as<WinResolved>();

	DevBuf dbgbuf;  // tracing only: per-phase time stamps of the CTAs of one stage
	if (getenv("QTIPSIM_R2_DBG")) {
		a2.dbg_stage = envu("QTIPSIM_R2_DBG", 50);
		if (dbgbuf.ensure((size_t)K * 16 * 8) == cudaSuccess) {
			cudaMemsetAsync(dbgbuf.p, 0, (size_t)K * 16 * 8, h->st);
			a2.dbg = dbgbuf.as<unsigned long long>();
		}
	}
	// ---- stages: launched in batches without host round trips; a batch ends with a look at the last anchor ----
	uint32_t launched = 0;
	StageAnchor fin;
	memset(&fin, 0, sizeof(fin));
	for (;;) {
		const uint32_t left = na - fin.first;
		uint32_t batch = (left + K - 1) / K + 2;
		if (launched + batch > max_launch) batch = max_launch - launched;
		if (!batch) return h->fail(QTIPSIM_ERANGE, "replay resolver made no progress");
		for (uint32_t b = 0; b < batch; b++) launch_resolve2_stage(a2, launched + b, K, h->st);
		launched += batch;
		CK(ctl_d2h(h, &fin, a2.anchor + launched, sizeof(fin)));
		CK(ctl_sync(h));
		if (fin.status != R2_RUNNING) break;
	}
	h->r2_stages_last = launched;
	if (a2.dbg) {
		std::vector<unsigned long long> st((size_t)K * 16);
		cudaMemcpy(st.data(), a2.dbg, st.size() * 8, cudaMemcpyDeviceToHost);
		unsigned long long t0 = ~0ull;
		for (uint32_t c = 0; c < K; c++)
			if (st[c * 16]) t0 = std::min(t0, st[c * 16]);
		fprintf(stderr, "[resolve2 dbg] stage %u: per CTA us since the first CTA started: start | binomU advU | binomB advB | binomCD advC | built | walk-begin walk-end done\n", a2.dbg_stage);
		for (uint32_t c = 0; c < K; c += (K > 16 ? K / 8 : 1)) {
			fprintf(stderr, "  cta %3u:", c);
			for (int k = 0; k < 12; k++) fprintf(stderr, " %7.1f", st[c * 16 + k] ? (double)(st[c * 16 + k] - t0) / 1e3 : -1.0);
			fprintf(stderr, "\n");
		}
		uint32_t cl = K - 1;
		fprintf(stderr, "  cta %3u:", cl);
		for (int k = 0; k < 12; k++) fprintf(stderr, " %7.1f", st[cl * 16 + k] ? (double)(st[cl * 16 + k] - t0) / 1e3 : -1.0);
		fprintf(stderr, "\n  cta %3u, inside the U region: reduce done %.1f, dense table %.1f, stride table %.1f, main orbit %.1f us\n", cl,
		        (double)(st[cl * 16 + 12] - t0) / 1e3, (double)(st[cl * 16 + 13] - t0) / 1e3, (double)(st[cl * 16 + 14] - t0) / 1e3,
		        (double)(st[cl * 16 + 15] - t0) / 1e3);
		for (uint32_t c = 0; c < K; c++)
			if (st[c * 16 + 11]) {
				fprintf(stderr, "  walker was cta %u: walk %0.1f us, epilogue %0.1f us\n", c, (double)(st[c * 16 + 10] - st[c * 16 + 9]) / 1e3,
				        (double)(st[c * 16 + 11] - st[c * 16 + 10]) / 1e3);
			}
		dbgbuf.release();
	}
	if (getenv("QTIPSIM_R2_TRACE"))
		fprintf(stderr, "[resolve2] %u live windows, K = %u, rcap %u, %u stage launches, status %u, first %u, draws %llu\n", na, K, a2.rcap,
		        launched, fin.status, fin.first, (unsigned long long)fin.abs);
	if (fin.status == R2_STUCK) return QTIPSIM_OK;  // done stays false: first-generation kernel

	// ---- counts per (window, category), offsets, totals ----
	uint32_t *d_cnt = h->d_r2_cnt.as<uint32_t>();
	uint32_t *d_start = d_cnt + (size_t)NCAT * na;
	uint32_t *d_err = h->d_r2_misc.as<uint32_t>();
	launch_resolve2_counts(a2, d_cnt, d_start, d_err, h->st);
	R2Offsets offs;
	for (int c = 0; c < NCAT; c++) {
		uint64_t *off = h->d_r2_off.as<uint64_t>() + (size_t)c * ((size_t)na + 1);
		launch_scan(d_cnt + (size_t)c * na, off, na, h->d_scan_tmp.as<uint64_t>(), off + na, h->st);
		offs.p[c] = off;
	}
	uint64_t tot[NCAT];
	uint32_t errw2[2] = {0, 0};
	for (int c = 0; c < NCAT; c++)
		CK(ctl_d2h(h, &tot[c], offs.p[c] + na, 8));
	CK(ctl_d2h(h, errw2, d_err, 8));
	CK(ctl_sync(h));
	uint32_t errw = errw2[0];
	if (errw) return QTIPSIM_OK;  // the first-generation kernel reports the reference's error for it
	for (int c = 0; c < NCAT; c++) {
		if (tot[c]) CK(h->d_plan[c].ensure(tot[c] * sizeof(PlanRec)));
		a2.ra.plan[c] = h->d_plan[c].as<PlanRec>();
		a2.ra.cap[c] = tot[c];
	}
	a2.walk_max = std::min<uint32_t>(4096u, envu("QTIPSIM_R2_WALK_MAX", 4096u));  // (k_resolve2_counts reports regions above 4096 reads)
	a2.any_large = errw2[1] > a2.walk_max || a2.walk_max < 4096u;
	launch_resolve2_emit(a2, d_cnt, d_start, offs, d_err, h->st);
	CK(ctl_d2h(h, &errw, d_err, 4));
	CK(ctl_sync(h));
	if (errw) return QTIPSIM_OK;
	for (int c = 0; c < NCAT; c++) nreads[c] = tot[c];
	p->cg1 = (int32_t)fin.s1;
	p->cg2 = (int32_t)fin.s2;
	draws = fin.abs;
	done = true;
	return QTIPSIM_OK;
}

// One run over the windows [range_lo, range_hi) of the packed records (range_lo < 0: the shard's share of all windows).
static int run_impl(qtipsim *h, qtipsim_run_params *p, qtipsim_stats *stats, int64_t range_lo, int64_t range_hi) {
	if (!h || !p) return QTIPSIM_EINVAL;
	if (!h->have_genome && range_lo < 0) return h->fail(QTIPSIM_ESTATE, "qtipsim_run: no genome loaded");
	CK(cudaSetDevice(h->device));
	const int launches0 = kernels_launch_count();
	const uint64_t olap = h->ts.olap();
	// assert(chunksz > olap), fasta.h:41
	if (!(p->chunksz > olap)) return h->fail(QTIPSIM_ERANGE, "chunksz must exceed the longest template / fragment (fasta.h:41)");
	if (p->chunksz > 0x7fffffffu) return h->fail(QTIPSIM_EINVAL, "chunksz too large");
	if (p->rng_mode != QTIPSIM_RNG_REPLAY && p->rng_mode != QTIPSIM_RNG_PHILOX) return h->fail(QTIPSIM_EINVAL, "bad rng_mode");
	if (p->shard_count == 0) p->shard_count = 1;
	if (p->shard_rank >= p->shard_count) return h->fail(QTIPSIM_EINVAL, "shard_rank >= shard_count");

	PlanParams pp;
	pp.chunksz = p->chunksz;
	pp.olap = olap;
	pp.tot_fasta_len = h->hg.file_bytes;
	// simplesim.cpp:205-208 (argument order u, c, d, b at the call site)
	pp.target[CAT_U] = apply_function(p->fraction, p->function, p->min_u, h->ts.cat[CAT_U].added);
	pp.target[CAT_B] = apply_function(p->fraction, p->function, p->min_b, h->ts.cat[CAT_B].added);
	pp.target[CAT_C] = apply_function(p->fraction, p->function, p->min_c, h->ts.cat[CAT_C].added);
	pp.target[CAT_D] = apply_function(p->fraction, p->function, p->min_d, h->ts.cat[CAT_D].added);
	uint64_t tsum = 0;
	for (int c = 0; c < NCAT; c++) {
		tsum += pp.target[c];
		if (pp.target[c] > 0x7fffffffull) return h->fail(QTIPSIM_ERANGE, "per-category target exceeds int range (draw_binomial casts to int)");
		if (pp.target[c] && h->ts.cat[c].t.empty())
			return h->fail(QTIPSIM_ERANGE, "non-zero target for a category with no templates (assert(!empty()), input_model.h:50)");
		if (pp.target[c] && !h->ts.cat[c].valid) return h->fail(QTIPSIM_EXSCRIPT, h->ts.cat[c].why);
	}
	if (tsum == 0) return h->fail(QTIPSIM_ERANGE, "nothing to simulate (assert(nu+nb+nc+nd > 0), simplesim.cpp:209)");
	if (p->rng_mode == QTIPSIM_RNG_REPLAY) {
		if (p->cg1 < 1 || (uint32_t)p->cg1 >= LM1 || p->cg2 < 1 || (uint32_t)p->cg2 >= LM2)
			return h->fail(QTIPSIM_ERANGE, "rnglib state out of range (set_seed, rnglib.cpp:1714-1728)");
	}

	int rc = upload_templates(h);
	if (rc) return rc;

	h->have_result = false;  // whatever the previous run left in d_out is about to be overwritten
	struct Events {
		cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
		~Events() {
			for (auto &x : e)
				if (x) cudaEventDestroy(x);
		}
	} evs;  // destroyed on every return path
	cudaEvent_t(&ev)[4] = evs.e;
	for (auto &e : ev) CK(cudaEventCreate(&e));
	CK(cudaEventRecord(ev[0], h->st));

	// ---- window table + N counts (cached: they depend on the genome, chunksz and olap only) ----
	if (!h->win_cache_ok || h->win_cache_chunksz != p->chunksz || h->win_cache_olap != olap) {
		h->win_cache = h->hg.windows(p->chunksz, olap);
		const uint32_t nw = (uint32_t)h->win_cache.size();
		h->ncount_cache.assign(nw, 0);
		if (nw) {
			DevGenome dg0;
			dg0.bases = h->d_bases.as<uint32_t>();
			dg0.nmask = h->d_nmask.as<uint32_t>();
			dg0.nblk = h->d_nblk.as<uint32_t>();
			dg0.recs = h->d_recs.as<RecDev>();
			dg0.names = h->d_names.as<char>();
			dg0.nrecs = (uint32_t)h->hg.recs.size();
			CK(h->d_win.ensure(nw * sizeof(Window)));
			CK(h->d_ncount.ensure(nw * 4));
			CK(ctl_h2d(h, h->d_win.p, h->win_cache.data(), nw * sizeof(Window)));
			launch_window_ncount(dg0, h->d_win.as<Window>(), nw, h->d_ncount.as<uint32_t>(), h->st);
			CK(ctl_d2h(h, h->ncount_cache.data(), h->d_ncount.p, nw * 4));
			CK(ctl_sync(h));
		}
		h->win_cache_chunksz = p->chunksz;
		h->win_cache_olap = olap;
		h->win_cache_ok = true;
	}
	const std::vector<Window> &win = h->win_cache;
	const uint32_t nwin = (uint32_t)win.size();
	pp.win_lo = (uint64_t)p->shard_rank * nwin / p->shard_count;
	pp.win_hi = (uint64_t)(p->shard_rank + 1) * nwin / p->shard_count;
	if (range_lo >= 0) {
		pp.win_lo = (uint64_t)range_lo;
		pp.win_hi = (uint64_t)range_hi < nwin ? (uint64_t)range_hi : nwin;
	}
	const std::vector<uint32_t> &ncount = h->ncount_cache;
	DevGenome dg;
	dg.bases = h->d_bases.as<uint32_t>();
	dg.nmask = h->d_nmask.as<uint32_t>();
	dg.nblk = h->d_nblk.as<uint32_t>();
	dg.recs = h->d_recs.as<RecDev>();
	dg.names = h->d_names.as<char>();
	dg.nrecs = (uint32_t)h->hg.recs.size();

	// ---- plan ----
	const double t_plan0 = now_ms();
	uint64_t nreads[NCAT] = {0, 0, 0, 0};
	uint64_t read_base[NCAT] = {0, 0, 0, 0};
	uint64_t draws = 0;
	const char *planner_env = getenv("QTIPSIM_PLANNER");
	const bool host_planner = planner_env && strcmp(planner_env, "host") == 0;
	if (p->rng_mode == QTIPSIM_RNG_REPLAY && !host_planner) {
		// ---- K2 on the device: speculative resolution of the draw chain (resolve.cu) ----
		// binomial set-ups per distinct window probability, computed with the host's libm
		rc = ensure_setups(h, win, olap, pp, (uint32_t)pp.win_lo, (uint32_t)pp.win_hi);
		if (rc) return rc;
		// default: the band-speculative resolver (resolve2.cu); QTIPSIM_RESOLVER=cluster|cta forces the first-generation
		// kernel, which is also what a declined / stuck run falls back to
		const char *rsv = getenv("QTIPSIM_RESOLVER");
		bool band_done = false;
		if (!rsv || !*rsv || strcmp(rsv, "band") == 0) {
			rc = resolve_replay_band(h, p, pp, dg, nwin, olap, nreads, draws, band_done);
			if (rc) return rc;
		}
		if (!band_done) {
			CK(h->d_rsresult.ensure(8 * 8));
			double grow = 1.0;
			for (int attempt = 0;; attempt++) {
				ResolveArgs ra;
				ra.g = dg;
				ra.win = h->d_win.as<Window>();
				ra.ncount = h->d_ncount.as<uint32_t>();
				ra.pclass = h->d_pclass.as<uint32_t>();
				ra.setups = h->d_setups.as<BinomSetup>();
				ra.nwin = nwin;
				{
					uint32_t mr = 0;
					for (int c = 0; c < NCAT; c++) mr = std::max(mr, h->ts.cat[c].max_rflen);
					ra.n_skip_ok = p->chunksz - olap >= (uint64_t)mr;
				}
				ra.win_lo = (uint32_t)pp.win_lo;
				ra.win_hi = (uint32_t)pp.win_hi;
				ra.olap = olap;
				const double share = nwin ? (double)(pp.win_hi - pp.win_lo) / (double)nwin : 1.0;
				for (int c = 0; c < NCAT; c++) {
					ra.t[c] = h->d_t[c].as<TmplDesc>();
					ra.xsteps[c] = h->d_xsteps[c].as<XStep>();
					ra.prog[c] = h->d_prog[c].as<HostMateProg>();
					ra.xp[c] = nullptr;
					ra.ntmpl[c] = (uint32_t)h->ts.cat[c].t.size();
					ra.target[c] = pp.target[c];
					// expected reads <= 1.1 * target * (this shard's share of the windows); generous margin
					ra.cap[c] = pp.target[c] ? (uint64_t)((1.25 * share * (double)pp.target[c] + 65536.0) * grow) : 0;
					if (ra.cap[c]) CK(h->d_plan[c].ensure(ra.cap[c] * sizeof(PlanRec)));
					ra.plan[c] = h->d_plan[c].as<PlanRec>();
				}
				ra.jt1 = h->d_jt1.as<uint32_t>();
				ra.jt2 = h->d_jt2.as<uint32_t>();
				ra.s1 = (uint32_t)p->cg1;
				ra.s2 = (uint32_t)p->cg2;
				ra.result = h->d_rsresult.as<unsigned long long>();
				launch_resolve_replay(ra, h->st);
				unsigned long long res[8];
				CK(ctl_d2h(h, res, h->d_rsresult.p, sizeof(res)));
				CK(ctl_sync(h));
				if (res[7] == 3 && attempt < 6) {  // plan buffer too small: retry larger
					grow *= 2.0;
					continue;
				}
				if (res[7] == 1) return h->fail(QTIPSIM_ERANGE, "window with zero slots (assert(nslots > 0), simplesim.cpp:252)");
				if (res[7]) return h->fail(QTIPSIM_ERANGE, "replay resolver failed");
				for (int c = 0; c < NCAT; c++) nreads[c] = res[c];
				p->cg1 = (int32_t)res[4];
				p->cg2 = (int32_t)res[5];
				draws = res[6];
				break;
			}
		}
	} else if (p->rng_mode == QTIPSIM_RNG_REPLAY) {
		rc = ensure_host_packed(h);
		if (rc) return rc;
		Plan plan;
		rc = plan_replay_host(h->hg, h->ts, win, ncount, pp, LState{(uint32_t)p->cg1, (uint32_t)p->cg2}, plan, h->err);
		if (rc) return rc;
		for (int c = 0; c < NCAT; c++) {
			nreads[c] = plan.recs[c].size();
			if (!nreads[c]) continue;
			CK(h->d_plan[c].ensure(nreads[c] * sizeof(PlanRec)));
			CK(cudaMemcpyAsync(h->d_plan[c].p, plan.recs[c].data(), nreads[c] * sizeof(PlanRec), cudaMemcpyHostToDevice, h->st));
		}
		CK(ctl_sync(h));  // plan vectors die at scope end
		p->cg1 = (int32_t)plan.rng_out.s1;
		p->cg2 = (int32_t)plan.rng_out.s2;
		draws = plan.draws;
	} else {
		// ---- PHILOX: per-window counts ~ Binomial(target, p_window) drawn on the device from a Philox
		// stream keyed by (seed, window, category); prefix sums give each read's window ----
		rc = ensure_setups(h, win, olap, pp, 0, nwin);
		if (rc) return rc;
		CK(h->d_len0.ensure((size_t)nwin * 4 + 16));
		CK(h->d_scan_tmp.ensure(((size_t)nwin / 1024 + 4) * 8));
		CK(h->d_cum.ensure(((size_t)nwin + 1) * 8 * NCAT));
		uint64_t edge[NCAT][2];
		for (int c = 0; c < NCAT; c++) {
			edge[c][0] = edge[c][1] = 0;
			if (!pp.target[c]) continue;  // Binomial(0, p) = 0 in every window: nothing to launch
			uint64_t *cum = h->d_cum.as<uint64_t>() + (size_t)c * (nwin + 1);
			launch_philox_counts(h->d_win.as<Window>(), h->d_ncount.as<uint32_t>(), h->d_pclass.as<uint32_t>(),
			                     h->d_setups.as<BinomSetup>(), nwin, olap, pp.target[c], p->philox_seed, c,
			                     h->d_len0.as<uint32_t>(), h->st);
			launch_scan(h->d_len0.as<uint32_t>(), cum, nwin, h->d_scan_tmp.as<uint64_t>(), cum + nwin, h->st);
			CK(ctl_d2h(h, &edge[c][0], cum + pp.win_lo, 8));
			CK(ctl_d2h(h, &edge[c][1], cum + pp.win_hi, 8));
		}
		CK(ctl_sync(h));
		for (int c = 0; c < NCAT; c++) {
			nreads[c] = edge[c][1] - edge[c][0];
			read_base[c] = edge[c][0];
			if (!nreads[c]) continue;
			const uint64_t *cum = h->d_cum.as<uint64_t>() + (size_t)c * (nwin + 1);
			CK(h->d_plan[c].ensure(nreads[c] * sizeof(PlanRec)));  // filled together with the record lengths below
			(void)cum;
		}
	}
	const double t_plan_host = now_ms() - t_plan0;
	CK(cudaEventRecord(ev[1], h->st));

	// ---- lengths, offsets ----
	uint64_t maxn = 0;
	for (int c = 0; c < NCAT; c++) maxn = nreads[c] > maxn ? nreads[c] : maxn;
	CK(h->d_len0.ensure(maxn * 4 + 16));
	CK(h->d_len1.ensure(maxn * 4 + 16));
	CK(h->d_scan_tmp.ensure((maxn / 1024 + 4) * 8));
	CK(h->d_scan_status.ensure(plan_philox_scan_status_bytes(maxn)));
	CK(h->d_small.ensure(16 * 8));
	CK(cudaMemsetAsync(h->d_small.p, 0, 16 * 8, h->st));
	uint64_t *d_tot = h->d_small.as<uint64_t>();           // [0..6] stream totals
	unsigned long long *d_gb = h->d_small.as<unsigned long long>() + 8;  // genome bytes
	SimArgs sa[NCAT];
	for (int c = 0; c < NCAT; c++) {
		SimArgs &a = sa[c];
		a.g = dg;
		a.c.t = h->d_t[c].as<TmplDesc>();
		a.c.qual = h->d_qual[c].as<char>();
		a.c.xs = h->d_xs[c].as<char>();
		a.c.n = (uint32_t)h->ts.cat[c].t.size();
		a.c.qual_out = h->d_qual_out[c].as<char>();
		a.c.qual_out_bytes = (uint32_t)h->ts.cat[c].qual_out.size();
		a.c.ops = h->d_ops[c].as<uint32_t>();
		a.c.segs = h->d_segs[c].as<SegDev>();
		a.genome_words = (h->hg.total_bases + 15) / 16 + 4;
		a.plan = h->d_plan[c].as<PlanRec>();
		a.nreads = nreads[c];
		a.read_index_base = read_base[c];
		a.cat = c;
		a.rng_mode = p->rng_mode;
		a.philox_seed = p->philox_seed;
		a.jt1 = h->d_jt1.as<uint32_t>();
		a.jt2 = h->d_jt2.as<uint32_t>();
		if (!nreads[c]) continue;
		const int s0 = cat_stream0(c);
		CK(h->d_off0[c].ensure(nreads[c] * 8));
		if (c != CAT_U) CK(h->d_off1[c].ensure(nreads[c] * 8));
		const bool split_plan = getenv("QTIPSIM_PLAN_SPLIT") != nullptr;  // tests: plan + lengths, then separate scans
		if (p->rng_mode == QTIPSIM_RNG_PHILOX && !split_plan) {
			// native plan, record lengths and stream offsets in one pass (the read's window by binary search over the
			// cumulative counts, the offsets by a single-pass prefix sum)
			const uint64_t *cum = h->d_cum.as<uint64_t>() + (size_t)c * (nwin + 1);
			launch_plan_philox_scan(h->d_win.as<Window>() + pp.win_lo, cum + pp.win_lo, (uint32_t)(pp.win_hi - pp.win_lo), olap,
			                        (uint32_t)h->ts.cat[c].t.size(), p->philox_seed, a, h->d_plan[c].as<PlanRec>(), h->d_off0[c].as<uint64_t>(),
			                        c != CAT_U ? h->d_off1[c].as<uint64_t>() : nullptr, d_tot + s0, d_tot + s0 + 1, d_gb,
			                        h->d_scan_status.as<unsigned long long>(), h->st);
		} else {
			if (p->rng_mode == QTIPSIM_RNG_PHILOX) {
				const uint64_t *cum = h->d_cum.as<uint64_t>() + (size_t)c * (nwin + 1);
				launch_plan_philox_lengths(h->d_win.as<Window>() + pp.win_lo, cum + pp.win_lo, (uint32_t)(pp.win_hi - pp.win_lo), olap,
				                           (uint32_t)h->ts.cat[c].t.size(), p->philox_seed, a, h->d_plan[c].as<PlanRec>(),
				                           h->d_len0.as<uint32_t>(), c != CAT_U ? h->d_len1.as<uint32_t>() : nullptr, d_gb, h->st);
			} else {
				launch_read_lengths(a, h->d_len0.as<uint32_t>(), c != CAT_U ? h->d_len1.as<uint32_t>() : nullptr, d_gb, h->st);
			}
			launch_scan(h->d_len0.as<uint32_t>(), h->d_off0[c].as<uint64_t>(), nreads[c], h->d_scan_tmp.as<uint64_t>(), d_tot + s0, h->st);
			if (c != CAT_U)
				launch_scan(h->d_len1.as<uint32_t>(), h->d_off1[c].as<uint64_t>(), nreads[c], h->d_scan_tmp.as<uint64_t>(), d_tot + s0 + 1, h->st);
		}
		a.off0 = h->d_off0[c].as<uint64_t>();
		a.off1 = c != CAT_U ? h->d_off1[c].as<uint64_t>() : nullptr;
	}
	uint64_t small[16];
	CK(ctl_d2h(h, small, h->d_small.p, sizeof(small)));
	CK(ctl_sync(h));
	for (int s = 0; s < NOUT; s++) {
		h->out_bytes[s] = small[s];
		CK(h->d_out[s].ensure(small[s] + 64));
	}
	CK(cudaEventRecord(ev[2], h->st));

	// ---- simulate + emit ----
	for (int c = 0; c < NCAT; c++) {
		if (!nreads[c]) continue;
		const int s0 = cat_stream0(c);
		sa[c].out0 = h->d_out[s0].as<char>();
		sa[c].out1 = c != CAT_U ? h->d_out[s0 + 1].as<char>() : nullptr;
		sa[c].tot0 = d_tot + s0;
		sa[c].tot1 = d_tot + s0 + 1;
		// upper bound on one FASTQ record of this category (name + 2 * body + 4)
		uint32_t max_refid = 0;
		uint64_t max_reclen = 0;
		for (const RecDev &r : h->hg.recs) {
			max_refid = r.name_len > max_refid ? r.name_len : max_refid;
			max_reclen = r.len > max_reclen ? r.len : max_reclen;
		}
		uint32_t dig = 1;
		for (uint64_t v = max_reclen + p->chunksz; v >= 10; v /= 10) dig++;
		const Category &cg = h->ts.cat[c];
		const uint32_t max_rec = (c == CAT_U ? 13u + max_refid + dig + cg.max_score_len + 2u
		                                     : 18u + 2u * (max_refid + dig + cg.max_score_len) + 2u) +
		                         2u * cg.max_body + 4u;
		const char *force = getenv("QTIPSIM_KERNEL");
		bool done = false;
		// production: the tile kernel (then its first generation, the direct kernel); categories they cannot take (a body under 17 bytes) use the staged kernel,
		// records too large for either the generic warp-per-read kernel.  QTIPSIM_KERNEL forces one (tests).
		const bool want_tile = !force || strcmp(force, "tile") == 0;
		const bool want_direct = !force || strcmp(force, "tile") == 0 || strcmp(force, "direct") == 0;
		const bool want_staged = !force || strcmp(force, "tile") == 0 || strcmp(force, "direct") == 0 || strcmp(force, "staged") == 0;
		// (the tile kernel prints record offsets with a 32-bit decimal conversion: longer records take the first generation)
		if (cg.compiled_ok && want_tile && max_reclen + p->chunksz <= 0xffffffffull) done = launch_simulate_tile(sa[c], max_rec - 2u * cg.max_body - 4u, cg.min_body, cg.max_body, h->st);
		if (!done && cg.compiled_ok && want_direct) done = launch_simulate_direct(sa[c], max_rec - 2u * cg.max_body - 4u, cg.min_body, cg.max_body, h->st);
		if (!done && cg.compiled_ok && want_staged) done = launch_simulate_warp(sa[c], max_rec, cg.max_body, h->st);
		if (!done) launch_simulate(sa[c], h->st);
	}
	CK(cudaEventRecord(ev[3], h->st));
#ifdef QT_BOUNDS
	{
		extern void qt_direct_debug_report();
		extern void qt_tile_debug_report();
		extern void qt_warp_debug_report();
		ctl_sync(h);
		qt_direct_debug_report();
		qt_tile_debug_report();
		qt_warp_debug_report();
	}
#endif
	CK(ctl_sync(h));
	CK(cudaGetLastError());
	h->have_result = true;
	for (int c = 0; c < NCAT; c++) h->out_reads[c] = nreads[c];

	if (stats) {
		memset(stats, 0, sizeof(*stats));
		uint64_t fq = 0;
		for (int c = 0; c < NCAT; c++) {
			stats->wrote[c] = nreads[c];
			stats->target[c] = pp.target[c];
		}
		for (int s = 0; s < NOUT; s++) {
			stats->bytes[s] = h->out_bytes[s];
			fq += h->out_bytes[s];
		}
		stats->draws = draws;
		stats->algorithmic_bytes = fq + small[8];
		stats->kernel_launches = (uint32_t)(kernels_launch_count() - launches0);
		float ms = 0;
		stats->ms_pack = h->ms_pack_pending;
		h->ms_pack_pending = 0;
		cudaEventElapsedTime(&ms, ev[0], ev[2]);
		stats->ms_plan = ms;  // includes the host chain walk in REPLAY mode
		(void)t_plan_host;
		cudaEventElapsedTime(&ms, ev[2], ev[3]);
		stats->ms_simulate = ms;
		cudaEventElapsedTime(&ms, ev[0], ev[3]);
		stats->ms_total = ms;
	}
	return QTIPSIM_OK;
}

int qtipsim_run(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats) { return run_impl(h, p, stats, -1, -1); }

int qtipsim_result_copy(qtipsim_t *h, int stream, void *dst, uint64_t cap) {
	if (!h || stream < 0 || stream >= NOUT) return QTIPSIM_EINVAL;
	if (!h->have_result) return h->fail(QTIPSIM_ESTATE, "no result: call qtipsim_run first");
	if (cap < h->out_bytes[stream]) return h->fail(QTIPSIM_ERANGE, "destination buffer too small");
	if (!h->out_bytes[stream]) return QTIPSIM_OK;
	if (!dst) return QTIPSIM_EINVAL;
	CK(cudaSetDevice(h->device));
	CK(cudaMemcpyAsync(dst, h->d_out[stream].p, h->out_bytes[stream], cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	return QTIPSIM_OK;
}

int qtipsim_result_device_ptr(qtipsim_t *h, int stream, const void **ptr, uint64_t *nbytes) {
	if (!h || stream < 0 || stream >= NOUT || !ptr || !nbytes) return QTIPSIM_EINVAL;
	if (!h->have_result) return h->fail(QTIPSIM_ESTATE, "no result: call qtipsim_run first");
	*ptr = h->d_out[stream].p;
	*nbytes = h->out_bytes[stream];
	return QTIPSIM_OK;
}

int qtipsim_result_write_fds(qtipsim_t *h, const int fds[QTIPSIM_NOUT]) {
	if (!h || !fds) return QTIPSIM_EINVAL;
	if (!h->have_result) return h->fail(QTIPSIM_ESTATE, "no result: call qtipsim_run first");
	CK(cudaSetDevice(h->device));
	const size_t STAGE = 32u << 20;
	if (!h->h_stage) {
		CK(cudaHostAlloc(&h->h_stage, 2 * STAGE, cudaHostAllocDefault));
		h->h_stage_cap = 2 * STAGE;
	}
	cudaEvent_t done[2];
	CK(cudaEventCreate(&done[0]));
	CK(cudaEventCreate(&done[1]));
	int rc = QTIPSIM_OK;
	for (int s = 0; s < NOUT && rc == QTIPSIM_OK; s++) {
		const uint64_t n = h->out_bytes[s];
		if (!n || fds[s] < 0) continue;
		const char *src = h->d_out[s].as<char>();
		// double-buffered: copy slice k+1 while slice k is written to the file
		uint64_t issued = 0, written = 0;
		int slot = 0;
		uint64_t len_in_slot[2] = {0, 0};
		auto issue = [&](int sl) {
			const uint64_t len = n - issued < STAGE ? n - issued : STAGE;
			len_in_slot[sl] = len;
			cudaMemcpyAsync((char *)h->h_stage + sl * STAGE, src + issued, len, cudaMemcpyDeviceToHost, h->st);
			cudaEventRecord(done[sl], h->st);
			issued += len;
		};
		issue(0);
		while (written < n) {
			if (issued < n) issue(slot ^ 1);
			cudaError_t e = cudaEventSynchronize(done[slot]);
			if (e != cudaSuccess) {
				rc = h->cuda_fail(e, "cudaEventSynchronize");
				break;
			}
			const char *hp = (const char *)h->h_stage + slot * STAGE;
			uint64_t left = len_in_slot[slot];
			while (left) {
				ssize_t w = write(fds[s], hp, left);
				if (w < 0) {
					if (errno == EINTR) continue;
					rc = h->fail(QTIPSIM_EIO, std::string("write: ") + strerror(errno));
					break;
				}
				hp += w;
				left -= (uint64_t)w;
			}
			if (rc) break;
			written += len_in_slot[slot];
			slot ^= 1;
		}
		ctl_sync(h);
	}
	cudaEventDestroy(done[0]);
	cudaEventDestroy(done[1]);
	return rc;
}

int qtipsim_result_stream_fds(qtipsim_t *h, const int fds[QTIPSIM_NOUT]) {
	if (!h || !fds) return QTIPSIM_EINVAL;
	if (!h->have_result) return h->fail(QTIPSIM_ESTATE, "no result: call qtipsim_run first");
	CK(cudaSetDevice(h->device));
	CK(ctl_sync(h));
	std::atomic<int> bad{0};
	std::string first_err;
	std::mutex em;
	std::vector<std::thread> th;
	for (int s = 0; s < NOUT; s++) {
		if (!h->out_bytes[s] || fds[s] < 0) continue;
		th.emplace_back([&, s]() {
			// one writer per descriptor: double-buffered device-to-host slices on a stream of its own, then write()
			const size_t SL = 8u << 20;
			auto fail = [&](const std::string &m) {
				std::lock_guard<std::mutex> lk(em);
				if (!bad.exchange(1)) first_err = m;
			};
			if (cudaSetDevice(h->device) != cudaSuccess) return fail("cudaSetDevice");
			cudaStream_t st;
			void *hb = nullptr;
			cudaEvent_t ev[2];
			if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate");
			if (cudaHostAlloc(&hb, 2 * SL, cudaHostAllocDefault) != cudaSuccess) {
				cudaStreamDestroy(st);
				return fail("cudaHostAlloc");
			}
			cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
			cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
			const char *src = h->d_out[s].as<char>();
			const uint64_t n = h->out_bytes[s];
			uint64_t issued = 0, written = 0, len_in[2] = {0, 0};
			int slot = 0;
			auto issue = [&](int sl) {
				const uint64_t len = std::min<uint64_t>(SL, n - issued);
				len_in[sl] = len;
				cudaMemcpyAsync((char *)hb + (size_t)sl * SL, src + issued, len, cudaMemcpyDeviceToHost, st);
				cudaEventRecord(ev[sl], st);
				issued += len;
			};
			issue(0);
			while (written < n && !bad.load()) {
				if (issued < n) issue(slot ^ 1);
				if (cudaEventSynchronize(ev[slot]) != cudaSuccess) {
					fail("cudaEventSynchronize");
					break;
				}
				const char *hp = (const char *)hb + (size_t)slot * SL;
				uint64_t left = len_in[slot];
				while (left) {
					const ssize_t w = write(fds[s], hp, left);
					if (w < 0) {
						if (errno == EINTR) continue;
						fail(std::string("write: ") + strerror(errno));
						left = 0;
						break;
					}
					hp += w;
					left -= (uint64_t)w;
				}
				written += len_in[slot];
				slot ^= 1;
			}
			cudaStreamSynchronize(st);
			cudaEventDestroy(ev[0]);
			cudaEventDestroy(ev[1]);
			cudaFreeHost(hb);
			cudaStreamDestroy(st);
		});
	}
	for (auto &t : th) t.join();
	if (bad.load()) {
		cudaGetLastError();
		return h->fail(QTIPSIM_EIO, "result_stream_fds: " + first_err);
	}
	return QTIPSIM_OK;
}

int qtipsim_result_interleaved(qtipsim_t *h, int cat, void *dst, uint64_t cap, uint64_t *nbytes) {
	if (!h || !nbytes || cat < CAT_B || cat > CAT_D) return QTIPSIM_EINVAL;
	if (!h->have_result) return h->fail(QTIPSIM_ESTATE, "no result: call qtipsim_run first");
	const int s0 = cat_stream0(cat);
	const uint64_t t0 = h->out_bytes[s0], t1 = h->out_bytes[s0 + 1], n = h->out_reads[cat];
	*nbytes = t0 + t1;
	if (!(t0 + t1)) return QTIPSIM_OK;
	if (cap < t0 + t1 || !dst) return h->fail(QTIPSIM_ERANGE, "destination buffer too small");
	CK(cudaSetDevice(h->device));
	CK(h->d_inter.ensure(t0 + t1 + 16));
	launch_interleave(h->d_out[s0].as<char>(), h->d_out[s0 + 1].as<char>(), h->d_off0[cat].as<uint64_t>(), h->d_off1[cat].as<uint64_t>(), n, t0, t1,
	                  h->d_inter.as<char>(), h->st);
	CK(cudaMemcpyAsync(dst, h->d_inter.p, t0 + t1, cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	CK(cudaGetLastError());
	return QTIPSIM_OK;
}

int qtipsim_run_to_host(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats, void *const dst[QTIPSIM_NOUT],
                        const uint64_t cap[QTIPSIM_NOUT]) {
	if (!h || !dst || !cap) return QTIPSIM_EINVAL;
	int rc = qtipsim_run(h, p, stats);
	if (rc) return rc;
	for (int s = 0; s < NOUT; s++) {
		if (cap[s] < h->out_bytes[s]) return h->fail(QTIPSIM_ERANGE, "destination buffer too small");
	}
	for (int s = 0; s < NOUT; s++) {
		if (h->out_bytes[s])
			CK(cudaMemcpyAsync(dst[s], h->d_out[s].p, h->out_bytes[s], cudaMemcpyDeviceToHost, h->st));
	}
	CK(ctl_sync(h));
	return QTIPSIM_OK;
}

// ---- streaming end-to-end path: upload, pack, simulate and download overlap ----
//
// The reference streams the FASTA once and emits reads window by window (simplesim.cpp:216-397); everything a
// window's reads depend on -- the window table, the per-window counts, the read indices, the output offsets -- is a
// prefix quantity over the records before it.  So the text can be uploaded in chunks on one copy stream, every
// record packed and simulated as soon as the header of the NEXT record has arrived, and its FASTQ bytes sent back
// on a second copy stream while later chunks are still uploading: both directions of the link run at once.  The
// bytes are those of qtipsim_load_fasta_bytes + qtipsim_run_to_host (tests/test_gpu_parity.py).

static uint64_t env_u64(const char *name, uint64_t dflt) {
	const char *e = getenv(name);
	if (!e || !*e) return dflt;
	const unsigned long long v = strtoull(e, nullptr, 10);
	return v ? (uint64_t)v : dflt;
}

int qtipsim_stream_begin(qtipsim_t *h, const void *bytes, size_t nbytes) {
	if (!h || (!bytes && nbytes)) return QTIPSIM_EINVAL;
	CK(cudaSetDevice(h->device));
	if (!h->st_h2d) CK(cudaStreamCreateWithFlags(&h->st_h2d, cudaStreamNonBlocking));
	if (!h->st_d2h) CK(cudaStreamCreateWithFlags(&h->st_d2h, cudaStreamNonBlocking));
	CK(ctl_sync(h));
	if (!h->ctl_ring) {
		const size_t cap = (size_t)env_u64("QTIPSIM_CTL_RING", 64ull << 20);
		if (cudaHostAlloc((void **)&h->ctl_ring, cap, cudaHostAllocMapped) == cudaSuccess) h->ctl_cap = cap;
		else {
			h->ctl_ring = nullptr;
			cudaGetLastError();  // no ring: control data goes through the copy engines as in the one-shot path
		}
	}
	h->ctl_zero_copy = h->ctl_ring != nullptr && !getenv("QTIPSIM_NO_CTL_RING");
	const uint64_t n = nbytes;
	genome_reset(h, n);
	h->stream_text = (const uint8_t *)bytes;
	h->stream_n = n;
	h->stream_chunk = env_u64("QTIPSIM_STREAM_CHUNK", 128ull << 20) & ~(uint64_t)15;
	if (h->stream_chunk < 16) h->stream_chunk = 16;
	// every record takes at least two text bytes and adds at most 127 positions of padding
	const uint64_t max_recs = std::min<uint64_t>(1ull << 22, n / 2 + 2);
	h->stream_cap_bases = n + 128 + 128 * max_recs;
	CK(h->d_text.ensure(n + 64));
	const size_t bw = (size_t)(h->stream_cap_bases / 16 + 8), mw = (size_t)(h->stream_cap_bases / 32 + 8);
	CK(h->d_bases.ensure(bw * 4));
	CK(h->d_nmask.ensure(mw * 4));
	CK(cudaMemsetAsync(h->d_bases.p, 0, bw * 4, h->st));
	CK(cudaMemsetAsync(h->d_nmask.p, 0, mw * 4, h->st));
	CK(cudaMemsetAsync((char *)h->d_text.p + n, ' ', 64, h->st));
	if (h->feeder.joinable()) h->feeder.join();
	for (cudaEvent_t e : h->stream_ev) cudaEventDestroy(e);
	h->stream_ev.clear();
	for (uint64_t c0 = 0; c0 < n; c0 += h->stream_chunk) {
		cudaEvent_t e;
		CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		h->stream_ev.push_back(e);
	}
	// The copy engine of a direction serves copies in issue order: were the whole text enqueued at once, every small
	// control upload of the pack / plan steps would wait behind it.  A feeder thread issues the text in pieces (32 MB) and keeps three in flight; stream_ev[i] fires when chunk i (the unit of the header scan) has arrived.
	h->chunks_issued.store(0);
	h->bulk_err.store(0);
	const uint64_t piece = std::max<uint64_t>(env_u64("QTIPSIM_STREAM_PIECE", 8ull << 20) & ~(uint64_t)15, 16);
	h->feeder = std::thread([h, bytes, n, piece]() {
		if (cudaSetDevice(h->device) != cudaSuccess) h->bulk_err.store(1);
		constexpr int DEPTH = 3;
		cudaEvent_t ring[DEPTH];
		for (auto &e : ring) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
		uint64_t pos = 0, k = 0;
		for (size_t ci = 0; ci < h->stream_ev.size(); ci++) {
			const uint64_t chunk_end = std::min<uint64_t>((uint64_t)(ci + 1) * h->stream_chunk, n);
			while (pos < chunk_end) {
				const uint64_t len = std::min<uint64_t>(piece, chunk_end - pos);
				if (k >= DEPTH) cudaEventSynchronize(ring[k % DEPTH]);
				if (cudaMemcpyAsync((char *)h->d_text.p + pos, (const char *)bytes + pos, len, cudaMemcpyHostToDevice, h->st_h2d) != cudaSuccess)
					h->bulk_err.store(1);
				cudaEventRecord(ring[k % DEPTH], h->st_h2d);
				k++;
				pos += len;
			}
			cudaEventRecord(h->stream_ev[ci], h->st_h2d);
			h->chunks_issued.store(ci + 1, std::memory_order_release);
		}
		cudaStreamSynchronize(h->st_h2d);
		for (auto &e : ring) cudaEventDestroy(e);
	});
	h->stream_open = true;
	h->stream_t0 = now_ms();
	return QTIPSIM_OK;
}

int qtipsim_stream_run(qtipsim_t *h, qtipsim_run_params *p, qtipsim_stats *stats, void *const dst[QTIPSIM_NOUT],
                       const uint64_t cap[QTIPSIM_NOUT]) {
	if (!h || !p || !dst || !cap) return QTIPSIM_EINVAL;
	if (!h->stream_open) return h->fail(QTIPSIM_ESTATE, "qtipsim_stream_run: call qtipsim_stream_begin first");
	if (p->shard_count > 1) return h->fail(QTIPSIM_EINVAL, "the streaming path runs all windows on one GPU (shard_count must be 1)");
	CK(cudaSetDevice(h->device));
	h->stream_open = false;
	struct CtlGuard {  // control data goes back to the copy engines when the streaming run is over, however it ends
		qtipsim *h;
		~CtlGuard() {
			ctl_sync(h);
			h->ctl_zero_copy = false;
		}
	} ctl_guard{h};
	{
		// templates go up now, while the first records are still arriving, not inside the first part's run
		const int rc0 = upload_templates(h);
		if (rc0) return rc0;
	}
	if (getenv("QTIPSIM_STREAM_TRACE")) fprintf(stderr, "[stream] stream_run entered at %.1f ms, templates up at %.1f ms\n", 0.0, now_ms() - h->stream_t0);
	const uint8_t *text = h->stream_text;
	const uint64_t n = h->stream_n;
	const uint64_t min_part = env_u64("QTIPSIM_STREAM_MIN_PART", 512ull << 20);
	const size_t nchunks = h->stream_ev.size();
	// downloads: a drain thread issues each part's FASTQ bytes in pieces of a few MB, three in flight (the small
	// control read-backs of the next parts' pack / plan steps slip in between), and counts the parts that are home
	struct DrainJob {
		char *dst;
		const char *src;
		uint64_t n;
		bool last_of_part;
	};
	std::deque<DrainJob> dq;
	std::mutex dm;
	std::condition_variable dcv;
	bool drain_stop = false;
	int parts_home = 0;
	const uint64_t dpiece = std::max<uint64_t>(env_u64("QTIPSIM_STREAM_PIECE", 8ull << 20), 4096);
	std::thread drain([&]() {
		if (cudaSetDevice(h->device) != cudaSuccess) h->bulk_err.store(1);
		constexpr int DEPTH = 3;
		cudaEvent_t ring[DEPTH];
		for (auto &e : ring) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
		uint64_t k = 0;
		for (;;) {
			DrainJob j;
			{
				std::unique_lock<std::mutex> lk(dm);
				dcv.wait(lk, [&] { return drain_stop || !dq.empty(); });
				if (dq.empty()) break;
				j = dq.front();
				dq.pop_front();
			}
			for (uint64_t pos = 0; pos < j.n; pos += dpiece) {
				const uint64_t len = std::min<uint64_t>(dpiece, j.n - pos);
				if (k >= DEPTH) cudaEventSynchronize(ring[k % DEPTH]);
				if (cudaMemcpyAsync(j.dst + pos, j.src + pos, len, cudaMemcpyDeviceToHost, h->st_d2h) != cudaSuccess) h->bulk_err.store(1);
				cudaEventRecord(ring[k % DEPTH], h->st_d2h);
				k++;
			}
			if (j.last_of_part) {
				cudaStreamSynchronize(h->st_d2h);
				std::lock_guard<std::mutex> lk(dm);
				parts_home++;
				dcv.notify_all();
			}
		}
		cudaStreamSynchronize(h->st_d2h);
		for (auto &e : ring) cudaEventDestroy(e);
	});
	auto stop_drain = [&]() {
		{
			std::lock_guard<std::mutex> lk(dm);
			drain_stop = true;
		}
		dcv.notify_all();
		if (drain.joinable()) drain.join();
	};
	uint64_t out_off[NOUT] = {0, 0, 0, 0, 0, 0, 0};
	qtipsim_stats tot;
	memset(&tot, 0, sizeof(tot));
	uint64_t win_done = 0;
	int part = 0, rc = QTIPSIM_OK;
	const bool trace_parts = getenv("QTIPSIM_STREAM_TRACE") != nullptr;
	std::vector<uint64_t> gts;
	for (size_t i = 0; (i < nchunks || (i == 0 && nchunks == 0)) && rc == QTIPSIM_OK; i++) {
		const bool final = i + 1 >= nchunks;
		if (nchunks) {
			const uint64_t c0 = (uint64_t)i * h->stream_chunk, c1 = std::min<uint64_t>(c0 + h->stream_chunk, n);
			while (h->chunks_issued.load(std::memory_order_acquire) <= i) std::this_thread::sleep_for(std::chrono::microseconds(50));
			{
				const cudaError_t e = cudaEventSynchronize(h->stream_ev[i]);
				if (e != cudaSuccess) {
					rc = h->cuda_fail(e, "cudaEventSynchronize(upload chunk)");
					break;
				}
			}
			const double t_arr = now_ms();
			rc = find_headers_chunk(h, text, c0, c1, gts);
			if (rc) break;
			if (trace_parts)
				fprintf(stderr, "[stream]   chunk %zu arrived %.1f ms, headers found %.1f ms (%zu)\n", i, t_arr - h->stream_t0,
				        now_ms() - h->stream_t0, gts.size());
		} else {
			gts.clear();
		}
		resolve_headers(h, text, n, gts, final);
		if (h->names_overflow) {
			rc = h->fail(QTIPSIM_ERANGE, "FASTA record names exceed 4 GiB in total");
			break;
		}
		const double t_ready = now_ms();
		const uint32_t r1 = (uint32_t)h->hg.recs.size();
		if (r1 > h->recs_packed) {
			const uint64_t span = h->seg_end[r1 - 1] - h->seg_start[h->recs_packed];
			if (!final && span < min_part) continue;
			rc = pack_records(h, true);
			if (rc) break;
		} else if (final && h->hg.total_bases == 0) {
			rc = pack_records(h, true);  // no record at all: the leading N pad, as the one-shot loader leaves it
			if (rc) break;
		} else if (!final) {
			continue;
		}
		// windows of the records packed so far: [win_done, nwin) are new
		const uint64_t olap = h->ts.olap();
		if (!(p->chunksz > olap)) {
			rc = h->fail(QTIPSIM_ERANGE, "chunksz must exceed the longest template / fragment (fasta.h:41)");
			break;
		}
		const uint64_t nwin = h->hg.windows(p->chunksz, olap).size();
		if (nwin == win_done && !(final && part == 0)) continue;
		// the output buffers about to be written were last read by the download of part - 2
		{
			std::unique_lock<std::mutex> lk(dm);
			dcv.wait(lk, [&] { return parts_home >= part - 1; });
		}
		qtipsim_stats st;
		memset(&st, 0, sizeof(st));
		const double t_run0 = now_ms();
		rc = run_impl(h, p, &st, (int64_t)win_done, (int64_t)nwin);
		if (rc) break;
		if (trace_parts)
			fprintf(stderr, "[stream] part %d: chunk %zu, ready %.1f ms, run %.1f -> %.1f ms, windows %llu..%llu, home %d\n", part, i,
			        t_ready - h->stream_t0, t_run0 - h->stream_t0, now_ms() - h->stream_t0, (unsigned long long)win_done,
			        (unsigned long long)nwin, parts_home);
		win_done = nwin;
		for (int s = 0; s < NOUT; s++) {
			const uint64_t nb = h->out_bytes[s];
			if (nb && (out_off[s] + nb > cap[s] || !dst[s])) rc = h->fail(QTIPSIM_ERANGE, "destination buffer too small");
		}
		if (rc) break;
		{
			std::lock_guard<std::mutex> lk(dm);
			for (int s = 0; s < NOUT; s++) {
				const uint64_t nb = h->out_bytes[s];
				if (!nb) continue;
				dq.push_back(DrainJob{(char *)dst[s] + out_off[s], h->d_out[s].as<char>(), nb, false});
				out_off[s] += nb;
			}
			dq.push_back(DrainJob{nullptr, nullptr, 0, true});
		}
		dcv.notify_all();
		for (int s = 0; s < NOUT; s++) std::swap(h->d_out[s], h->d_out2[s]);
		for (int c = 0; c < NCAT; c++) {
			tot.wrote[c] += st.wrote[c];
			tot.target[c] = st.target[c];
		}
		tot.draws += st.draws;
		tot.algorithmic_bytes += st.algorithmic_bytes;
		tot.kernel_launches += st.kernel_launches;
		tot.ms_plan += st.ms_plan;
		tot.ms_simulate += st.ms_simulate;
		tot.ms_total += st.ms_total;
		tot.ms_pack += st.ms_pack;
		part++;
	}
	if (trace_parts) fprintf(stderr, "[stream] last part issued at %.1f ms\n", now_ms() - h->stream_t0);
	stop_drain();
	if (trace_parts) fprintf(stderr, "[stream] downloads home at %.1f ms\n", now_ms() - h->stream_t0);
	if (h->feeder.joinable()) h->feeder.join();
	cudaStreamSynchronize(h->st_h2d);
	cudaStreamSynchronize(h->st_d2h);
	ctl_sync(h);
	for (cudaEvent_t e : h->stream_ev) cudaEventDestroy(e);
	h->stream_ev.clear();
	if (rc) return rc;
	if (h->bulk_err.load()) return h->fail(QTIPSIM_ECUDA, "streaming path: a bulk copy failed");
	CK(cudaGetLastError());
	for (int s = 0; s < NOUT; s++) {
		tot.bytes[s] = out_off[s];
		h->out_bytes[s] = 0;
	}
	h->have_result = false;  // the parts' buffers were recycled: the bytes are in dst
	h->have_genome = true;   // the whole genome is packed: qtipsim_run works from here on
	if (stats) *stats = tot;
	return QTIPSIM_OK;
}

int qtipsim_label_correctness(qtipsim_t *h, const char *qnames, uint64_t qnames_bytes, const uint64_t *q_off, const char *rnames,
                              uint64_t rnames_bytes, const uint64_t *r_off, const int32_t *flag, const uint64_t *pos, uint64_t n,
                              int32_t wiggle, int8_t *correct) {
	if (!h) return QTIPSIM_EINVAL;
	if (!n) return QTIPSIM_OK;
	if (!qnames || !q_off || !rnames || !r_off || !flag || !pos || !correct || !qnames_bytes || !rnames_bytes) return QTIPSIM_EINVAL;
	// every string must end inside its arena: the kernel scans to the terminator
	if (qnames[qnames_bytes - 1] != '\0' || rnames[rnames_bytes - 1] != '\0') return h->fail(QTIPSIM_EINVAL, "label_correctness: string arenas must end with a NUL");
	for (uint64_t i = 0; i < n; i++)
		if (q_off[i] >= qnames_bytes || r_off[i] >= rnames_bytes) return h->fail(QTIPSIM_EINVAL, "label_correctness: string offset outside its arena");
	CK(cudaSetDevice(h->device));
	const size_t qa = (qnames_bytes + 15) & ~(size_t)15, ra = (rnames_bytes + 15) & ~(size_t)15, na8 = (n * 8 + 15) & ~(size_t)15,
	             na4 = (n * 4 + 15) & ~(size_t)15;
	CK(h->d_label.ensure(qa + ra + 3 * na8 + na4 + n + 64));
	char *d = h->d_label.as<char>();
	char *d_q = d, *d_r = d_q + qa;
	uint64_t *d_qo = (uint64_t *)(d_r + ra), *d_ro = (uint64_t *)((char *)d_qo + na8), *d_pos = (uint64_t *)((char *)d_ro + na8);
	int32_t *d_flag = (int32_t *)((char *)d_pos + na8);
	int8_t *d_out = (int8_t *)((char *)d_flag + na4);
	CK(cudaMemcpyAsync(d_q, qnames, qnames_bytes, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_r, rnames, rnames_bytes, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_qo, q_off, n * 8, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_ro, r_off, n * 8, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_pos, pos, n * 8, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_flag, flag, n * 4, cudaMemcpyHostToDevice, h->st));
	launch_label_correctness(d_q, d_qo, d_r, d_ro, d_flag, d_pos, n, wiggle, d_out, h->st);
	CK(cudaMemcpyAsync(correct, d_out, n, cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	CK(cudaGetLastError());
	return QTIPSIM_OK;
}

int qtipsim_edit_xscripts(qtipsim_t *h, const char *cigars, uint64_t cigars_bytes, const uint64_t *c_off, const char *mdzs,
                          uint64_t mdzs_bytes, const uint64_t *m_off, const uint64_t *pos, uint64_t n, char *xscripts, uint64_t cap,
                          uint64_t *x_off, int32_t *left_clip, int32_t *right_clip, uint64_t *lpos, uint64_t *rpos, int32_t *status) {
	if (!h || !x_off) return QTIPSIM_EINVAL;
	x_off[0] = 0;
	if (!n) return QTIPSIM_OK;
	if (!cigars || !cigars_bytes || !c_off || !m_off || !pos || !left_clip || !right_clip || !lpos || !rpos || !status) return QTIPSIM_EINVAL;
	if (cigars[cigars_bytes - 1] != '\0' || (mdzs_bytes && (!mdzs || mdzs[mdzs_bytes - 1] != '\0')))
		return h->fail(QTIPSIM_EINVAL, "edit_xscripts: string arenas must end with a NUL");
	for (uint64_t i = 0; i < n; i++)
		if (c_off[i] >= cigars_bytes || (m_off[i] != ~0ull && m_off[i] >= mdzs_bytes)) return h->fail(QTIPSIM_EINVAL, "edit_xscripts: string offset outside its arena");
	CK(cudaSetDevice(h->device));
	auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
	const size_t in_bytes = al(cigars_bytes) + al(mdzs_bytes + 1) + 3 * al(n * 8) + al((n + 1) * 8) + 2 * al(n * 8) + 4 * al(n * 4) + 64;
	CK(h->d_label.ensure(in_bytes));
	CK(h->d_scan_tmp.ensure((n / 1024 + 4) * 8));
	char *d = h->d_label.as<char>();
	char *d_c = d;
	char *d_m = d_c + al(cigars_bytes);
	uint64_t *d_co = (uint64_t *)(d_m + al(mdzs_bytes + 1));
	uint64_t *d_mo = (uint64_t *)((char *)d_co + al(n * 8));
	uint64_t *d_pos = (uint64_t *)((char *)d_mo + al(n * 8));
	uint64_t *d_xo = (uint64_t *)((char *)d_pos + al(n * 8));
	uint64_t *d_lp = (uint64_t *)((char *)d_xo + al((n + 1) * 8));
	uint64_t *d_rp = (uint64_t *)((char *)d_lp + al(n * 8));
	uint32_t *d_len = (uint32_t *)((char *)d_rp + al(n * 8));
	int32_t *d_lc = (int32_t *)((char *)d_len + al(n * 4));
	int32_t *d_rc = (int32_t *)((char *)d_lc + al(n * 4));
	int32_t *d_st = (int32_t *)((char *)d_rc + al(n * 4));
	CK(cudaMemcpyAsync(d_c, cigars, cigars_bytes, cudaMemcpyHostToDevice, h->st));
	if (mdzs_bytes) CK(cudaMemcpyAsync(d_m, mdzs, mdzs_bytes, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_co, c_off, n * 8, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_mo, m_off, n * 8, cudaMemcpyHostToDevice, h->st));
	CK(cudaMemcpyAsync(d_pos, pos, n * 8, cudaMemcpyHostToDevice, h->st));
	launch_xs_lengths(d_c, d_co, d_m, d_mo, d_pos, n, d_len, d_lc, d_rc, d_st, h->st);
	launch_scan(d_len, d_xo, n, h->d_scan_tmp.as<uint64_t>(), d_xo + n, h->st);
	uint64_t total = 0;
	CK(cudaMemcpyAsync(&total, d_xo + n, 8, cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	x_off[n] = total;
	if (total > cap || !xscripts) return h->fail(QTIPSIM_ERANGE, "edit_xscripts: transcript arena too small");
	CK(h->d_xs_out.ensure(total + 16));
	launch_xs_fill(d_c, d_co, d_m, d_mo, d_pos, n, d_xo, d_st, h->d_xs_out.as<char>(), d_lp, d_rp, h->st);
	CK(cudaMemcpyAsync(xscripts, h->d_xs_out.p, total, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(x_off, d_xo, (n + 1) * 8, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(left_clip, d_lc, n * 4, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(right_clip, d_rc, n * 4, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(lpos, d_lp, n * 8, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(rpos, d_rp, n * 8, cudaMemcpyDeviceToHost, h->st));
	CK(cudaMemcpyAsync(status, d_st, n * 4, cudaMemcpyDeviceToHost, h->st));
	CK(ctl_sync(h));
	CK(cudaGetLastError());
	return QTIPSIM_OK;
}

}  // extern "C"
