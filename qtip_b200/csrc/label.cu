// label.cu -- k_label_correctness: Alignment::set_correctness (qtip_parse.cpp:519-713) for a batch of aligned
// tandem-SAM records, one thread per record.  HBM-bound byte work: ~100 name bytes read and one label written per
// record; see label.h for the grammar.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"
#include "label.h"

namespace qt {

__global__ void __launch_bounds__(256) k_label_correctness(const char *qnames, const uint64_t *q_off, const char *rnames, const uint64_t *r_off,
                                                         const int32_t *flag, const uint64_t *pos, uint64_t n, int32_t wiggle, int8_t *correct) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	correct[i] = (int8_t)label_correctness(qnames + q_off[i], rnames + r_off[i], flag[i], pos[i], wiggle);
}

void launch_label_correctness(const char *qnames, const uint64_t *q_off, const char *rnames, const uint64_t *r_off, const int32_t *flag,
                              const uint64_t *pos, uint64_t n, int32_t wiggle, int8_t *correct, cudaStream_t st) {
	if (!n) return;
	k_label_correctness<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(qnames, q_off, rnames, r_off, flag, pos, n, wiggle, correct);
	kernels_count_launch();
}

}  // namespace qt
