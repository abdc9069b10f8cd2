// resolve_dev.cuh -- device helpers shared by the two replay resolvers (resolve.cu, resolve2.cu): table jumps
// along the addressable L'Ecuyer stream and the draw count of one mate's transcript (simplesim.cpp:38-62).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"
#include "lecuyer.h"
#include "resolve.h"

namespace qt {

__device__ __forceinline__ LState jump_tbl(LState s, uint32_t k, const uint32_t *jt1, const uint32_t *jt2) {
	while (k >= (uint32_t)JUMP_TBL) {
		s.s1 = mulmod1(s.s1, __ldg(&jt1[JUMP_TBL - 1]));
		s.s2 = mulmod2(s.s2, __ldg(&jt2[JUMP_TBL - 1]));
		k -= (uint32_t)(JUMP_TBL - 1);
	}
	if (k) {
		s.s1 = mulmod1(s.s1, __ldg(&jt1[k]));
		s.s2 = mulmod2(s.s2, __ldg(&jt2[k]));
	}
	return s;
}

// has_n = the window holds at least one N (its N count is known): only then is the N mask read
__device__ __forceinline__ uint32_t rs_fetch_code(const DevGenome &g, uint64_t gp, uint64_t gend, bool has_n) {
	if (gp >= gend) return 4u;
	if (has_n && ((__ldg(&g.nmask[gp >> 5]) >> (gp & 31)) & 1u)) return 4u;
	return (__ldg(&g.bases[gp >> 4]) >> (2 * (gp & 15))) & 3u;
}

// draws consumed by one mate's transcript starting at state s (simplesim.cpp:38-62); advances s
__device__ __forceinline__ uint32_t walk_mate_dev(LState &s, const DevGenome &g, const XStep *__restrict__ xs,
                                                  const HostMateProg hp, uint64_t gpos, uint64_t gend, bool has_n,
                                                  const uint32_t *jt1, const uint32_t *jt2) {
	uint32_t len = 0;
	for (uint32_t i = hp.xs_begin; i < hp.xs_end; i++) {
		const XStep x = xs[i];
		if (x.skip) {
			s = jump_tbl(s, x.skip, jt1, jt2);
			len += x.skip;
		}
		const uint32_t ref = rs_fetch_code(g, gpos + x.rf, gend, has_n);
		uint32_t code;
		do {
			s = lstep(s);
			len++;
			code = lbase_code_z(lout(s));
		} while (code == ref);
	}
	if (hp.tail_skip) {
		s = jump_tbl(s, hp.tail_skip, jt1, jt2);
		len += hp.tail_skip;
	}
	return len;
}


// The same walk from a packed program (templates.h XProgHalf, eight words already in registers): every X's reference
// base is fetched before the first draw is made, so the loads overlap instead of alternating with the draws.
__device__ __forceinline__ uint32_t walk_mate_xp(LState &s, const DevGenome &g, const uint32_t (&w)[8], uint64_t gpos, uint64_t gend,
                                                 bool has_n, const uint32_t *jt1, const uint32_t *jt2) {
	const uint32_t nx = (w[5] >> 8) & 0xffu, tail = (w[5] >> 16) & 0xffu;
	uint32_t codes = 0;
#pragma unroll
	for (int k = 0; k < 7; k++) {
		if ((uint32_t)k < nx) {
			const uint32_t rf = (w[k >> 1] >> (16 * (k & 1))) & 0xffffu;
			codes |= rs_fetch_code(g, gpos + rf, gend, has_n) << (3 * k);
		}
	}
	uint32_t len = 0;
#pragma unroll
	for (int k = 0; k < 7; k++) {
		if ((uint32_t)k < nx) {
			const uint32_t sk = (w[(14 + k) >> 2] >> (8 * ((14 + k) & 3))) & 0xffu;
			if (sk) {
				s = jump_tbl(s, sk, jt1, jt2);
				len += sk;
			}
			const uint32_t ref = (codes >> (3 * k)) & 7u;
			uint32_t code;
			do {
				s = lstep(s);
				len++;
				code = lbase_code_z(lout(s));
			} while (code == ref);
		}
	}
	if (tail) {
		s = jump_tbl(s, tail, jt1, jt2);
		len += tail;
	}
	return len;
}

}  // namespace qt
