// xscript.cu -- SAM fields -> edit transcripts for a batch of alignments (see xscript.h).  Two passes, one
// thread per alignment: k_xs_lengths parses CIGAR + MD:Z and returns the transcript length, clips and status; a scan
// gives every transcript its place in one arena; k_xs_fill writes it and derives lpos / rpos.  HBM-bound byte work.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"
#include "xscript.h"

namespace qt {

static constexpr uint64_t XS_NONE = 0xffffffffffffffffull;

__global__ void __launch_bounds__(128) k_xs_lengths(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off,
                                                  const uint64_t *pos, uint64_t n, uint32_t *len1, int32_t *lclip, int32_t *rclip,
                                                  int32_t *status) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const char *md = m_off[i] == XS_NONE ? nullptr : mdzs + m_off[i];
	const XsInfo r = xs_build(cigars + c_off[i], md, pos[i], nullptr, 0);
	len1[i] = r.status == XS_OK ? r.len + 1u : 1u;  // with the terminator; failed alignments get an empty transcript
	lclip[i] = r.left_clip;
	rclip[i] = r.right_clip;
	status[i] = r.status;
}

__global__ void __launch_bounds__(128) k_xs_fill(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off,
                                               const uint64_t *pos, uint64_t n, const uint64_t *x_off, const int32_t *status, char *xs,
                                               uint64_t *lpos, uint64_t *rpos) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	char *out = xs + x_off[i];
	const uint32_t cap = (uint32_t)(x_off[i + 1] - x_off[i]);
	if (status[i] != XS_OK) {
		out[0] = '\0';
		lpos[i] = rpos[i] = pos[i];
		return;
	}
	const char *md = m_off[i] == XS_NONE ? nullptr : mdzs + m_off[i];
	const XsInfo r = xs_build(cigars + c_off[i], md, pos[i], out, cap);
	uint64_t lp, rp;
	xs_span(out, r.len, pos[i], r.left_clip, lp, rp);
	lpos[i] = lp;
	rpos[i] = rp;
}

void launch_xs_lengths(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off, const uint64_t *pos, uint64_t n,
                       uint32_t *len1, int32_t *lclip, int32_t *rclip, int32_t *status, cudaStream_t st) {
	if (!n) return;
	k_xs_lengths<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cigars, c_off, mdzs, m_off, pos, n, len1, lclip, rclip, status);
	kernels_count_launch();
}

void launch_xs_fill(const char *cigars, const uint64_t *c_off, const char *mdzs, const uint64_t *m_off, const uint64_t *pos, uint64_t n,
                    const uint64_t *x_off, const int32_t *status, char *xs, uint64_t *lpos, uint64_t *rpos, cudaStream_t st) {
	if (!n) return;
	k_xs_fill<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cigars, c_off, mdzs, m_off, pos, n, x_off, status, xs, lpos, rpos);
	kernels_count_launch();
}

}  // namespace qt
