// lecuyer.h -- the reference's random stream, restated for host and device.
//
// Replay mode has to consume exactly the draws the reference's rnglib/ranlib would:
//   * i4_uni / r4_uni_01     rnglib.cpp:573-683, :1444-1505
//   * draw_base              simplesim.h:26-28
//   * index draws            input_model.h:51, simplesim.cpp:253
//   * ignbin                 ranlib.cpp:1094-1389
// Unlike the reference (one global state stepped one draw at a time) the stream here is
// *addressable*: state_k = state_0 * a^k mod m for each of the two component MLCGs, so
// any thread can compute any draw (jump tables below).  All float arithmetic that
// decides an outcome is written with explicit rounding (__f*_rn / __d*_rn on device)
// so it is immune to fused-multiply-add contraction and matches the reference's
// baseline-x86-64 build operation for operation.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#else
#define QT_HD inline
#endif

namespace qt {

static constexpr uint32_t LM1 = 2147483563u;  // 2^31 - 85
static constexpr uint32_t LM2 = 2147483399u;  // 2^31 - 249
static constexpr uint32_t LA1 = 40014u;
static constexpr uint32_t LA2 = 40692u;

// (a*b) mod (2^31 - 85), operands < 2^31.  2^31 == 85 (mod m1) folds the 62-bit product.
QT_HD uint32_t mulmod1(uint32_t a, uint32_t b) {
	uint64_t x = (uint64_t)a * b;
	uint64_t y = (x >> 31) * 85u + (x & 0x7fffffffu);
	uint32_t z = (uint32_t)((y >> 31) * 85u + (y & 0x7fffffffu));
	return z >= LM1 ? z - LM1 : z;
}
QT_HD uint32_t mulmod2(uint32_t a, uint32_t b) {
	uint64_t x = (uint64_t)a * b;
	uint64_t y = (x >> 31) * 249u + (x & 0x7fffffffu);
	uint32_t z = (uint32_t)((y >> 31) * 249u + (y & 0x7fffffffu));
	return z >= LM2 ? z - LM2 : z;
}

struct LState {
	uint32_t s1, s2;
};

// (a*w) mod m for a CONSTANT multiplier w with wp = floor(w * 2^32 / m) precomputed (Shoup): one high and two low
// multiplies and a conditional subtract instead of two 62-bit foldings.  Exact for a < m < 2^31.
QT_HD uint32_t mulmod_const(uint32_t a, uint32_t w, uint32_t wp, uint32_t m) {
#if defined(__CUDA_ARCH__)
	const uint32_t q = __umulhi(a, wp);
#else
	const uint32_t q = (uint32_t)(((uint64_t)a * wp) >> 32);
#endif
	const uint32_t r = a * w - q * m;
	return r >= m ? r - m : r;
}
static constexpr uint32_t LA1P = 80028u;  // floor(40014 * 2^32 / (2^31 - 85))
static constexpr uint32_t LA2P = 81384u;  // floor(40692 * 2^32 / (2^31 - 249))

QT_HD LState lstep(LState s) { return LState{mulmod_const(s.s1, LA1, LA1P, LM1), mulmod_const(s.s2, LA2, LA2P, LM2)}; }

// one step BACK: multiplication by a^-1 mod m (the stream is a group, so it can be walked in either direction)
static constexpr uint32_t LB1 = 2082061899u, LB1P = 4164123962u;  // 40014^-1 mod (2^31 - 85), floor(LB1 * 2^32 / m1)
static constexpr uint32_t LB2 = 1481316021u, LB2P = 2962632385u;  // 40692^-1 mod (2^31 - 249)
QT_HD LState lstep_back(LState s) { return LState{mulmod_const(s.s1, LB1, LB1P, LM1), mulmod_const(s.s2, LB2, LB2P, LM2)}; }

// i4_uni's output for the state *after* the step (rnglib.cpp:667-672)
QT_HD int32_t lout(LState s) {
	int32_t z = (int32_t)s.s1 - (int32_t)s.s2;
	if (z < 1) z += (int32_t)(LM1 - 1);
	return z;
}

// r4_uni_01's value for i4_uni output z: (float)z, times a double constant, to float
// (rnglib.cpp:1502).  Can be exactly 1.0f for the top ~100 values of z.
QT_HD float lfloat(int32_t z) {
#if defined(__CUDA_ARCH__)
	return __double2float_rn(__dmul_rn((double)__int2float_rn(z), 4.656613057E-10));
#else
	// host objects are built with -ffp-contract=off (Makefile): one rounding per operation
	float f = (float)z;
	double d = (double)f * 4.656613057E-10;
	return (float)d;
#endif
}

QT_HD float fmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fmul_rn(a, b);
#else
	return a * b;
#endif
}
QT_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fadd_rn(a, b);
#else
	return a + b;
#endif
}
QT_HD float fsub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fsub_rn(a, b);
#else
	return a - b;
#endif
}
QT_HD float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fdiv_rn(a, b);
#else
	return a / b;
#endif
}
QT_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
QT_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
QT_HD double dsub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}
QT_HD double ddiv_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __ddiv_rn(a, b);
#else
	return a / b;
#endif
}

// draw_base (simplesim.h:26-28): "ACGT"[min((int)(r4*4), 3)] as a 2-bit code.
QT_HD uint32_t lbase_code(float u) {
	int k = (int)fmul_rn(u, 4.0f);
	return (uint32_t)(k < 3 ? k : 3);
}

// The same code straight from i4_uni's integer z: lbase_code(lfloat(z)) is monotone in z (every step of it is), so it is
// three comparisons against the smallest z of each code (found by bisection over the exact expression; hosttest.cpp
// qthost_base_thresholds re-derives them and tests/test_host_logic.py checks them around every threshold).
static constexpr int32_t LBASE_T1 = 536870865, LBASE_T2 = 1073741729, LBASE_T3 = 1610612672;
QT_HD uint32_t lbase_code_z(int32_t z) { return (uint32_t)(z >= LBASE_T1) + (uint32_t)(z >= LBASE_T2) + (uint32_t)(z >= LBASE_T3); }

// index draws: min((size_t)(r4 * (float)n), n-1)  (input_model.h:51, simplesim.cpp:253)
QT_HD uint64_t lindex(float u, uint64_t n) {
	float prod = fmul_rn(u, (float)n);
	uint64_t v = (uint64_t)prod;
	return v < n - 1 ? v : n - 1;
}

// a^e mod m by square-and-multiply (host-side table construction and long jumps)
QT_HD uint32_t powmod1(uint32_t a, uint64_t e) {
	uint32_t r = 1;
	while (e) {
		if (e & 1) r = mulmod1(r, a);
		a = mulmod1(a, a);
		e >>= 1;
	}
	return r;
}
QT_HD uint32_t powmod2(uint32_t a, uint64_t e) {
	uint32_t r = 1;
	while (e) {
		if (e & 1) r = mulmod2(r, a);
		a = mulmod2(a, a);
		e >>= 1;
	}
	return r;
}
QT_HD LState ljump(LState s, uint64_t k) {
	return LState{mulmod1(s.s1, powmod1(LA1, k)), mulmod2(s.s2, powmod2(LA2, k))};
}

// ---------------------------------------------------------------------------
// logf, bit-identical to glibc's (sysdeps/ieee754/flt-32/e_logf.c, the ARM
// optimized-routines algorithm: 16-entry table, degree-3 polynomial in double).
// Checked against the host libm over every positive normal float by
// tests/test_logf_exhaustive.py (and on a stride in tests/test_host_logic.py);
// the fused and unfused evaluations agree after rounding to float on all of them.
// Valid for positive normal x, which is all ignbin ever passes.
// ---------------------------------------------------------------------------
#define QT_LOGF_TABLE                                                                                  \
	{{0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},     \
	 {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},     \
	 {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},        \
	 {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},     \
	 {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},                                  \
	 {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},       \
	 {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},       \
	 {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}}
#if defined(__CUDACC__)
// constant memory on the device: a table in the function body would be rebuilt on the stack at every call
static __constant__ double qt_logf_tab_dev[16][2] = QT_LOGF_TABLE;
#endif
static const double qt_logf_tab_host[16][2] = QT_LOGF_TABLE;

QT_HD float logf_glibc(float x) {
#if defined(__CUDA_ARCH__)
	const double(*T)[2] = qt_logf_tab_dev;
#else
	const double(*T)[2] = qt_logf_tab_host;
#endif
	const double Ln2 = 0x1.62e42fefa39efp-1;
	const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
	uint32_t ix;
#if defined(__CUDA_ARCH__)
	ix = __float_as_uint(x);
#else
	memcpy(&ix, &x, 4);
#endif
	if (ix == 0x3f800000u) return 0.0f;
	uint32_t tmp = ix - 0x3f330000u;
	int i = (int)((tmp >> 19) & 15u);
	int k = (int32_t)tmp >> 23;
	uint32_t iz = ix - (tmp & 0xff800000u);
	float zf;
#if defined(__CUDA_ARCH__)
	zf = __uint_as_float(iz);
#else
	memcpy(&zf, &iz, 4);
#endif
	double z = (double)zf;
	double r = dsub_rn(dmul_rn(z, T[i][0]), 1.0);
	double y0 = dadd_rn(T[i][1], dmul_rn((double)k, Ln2));
	double r2 = dmul_rn(r, r);
	double y = dadd_rn(dmul_rn(A1, r), A2);
	y = dadd_rn(dmul_rn(A0, r2), y);
	y = dadd_rn(dmul_rn(y, r2), dadd_rn(y0, r));
#if defined(__CUDA_ARCH__)
	return __double2float_rn(y);
#else
	return (float)y;
#endif
}

// ---------------------------------------------------------------------------
// ignbin (ranlib.cpp:1094-1389), split into a set-up that depends only on (n, pp)
// -- computed on the host with the host's libm (pow, sqrt) once per distinct pair --
// and a sampler that only needs the uniform stream, IEEE float/double arithmetic
// and logf_glibc, so the identical code runs on host and device.
// ---------------------------------------------------------------------------
struct BinomSetup {
	int32_t n;
	float pp;
	int32_t inversion;  // 1: xnp < 30 path, ranlib.cpp:1198-1231
	int32_t flip;       // 0.5 < pp
	float p, q, xnp;
	// inversion
	float qn, r, g;
	// BTPE, ranlib.cpp:1232-1251
	int32_t m;
	float ffm, fm, xnpq, p1, xm, xl, xr, c, xll, xlr, p2, p3, p4;
};

// Host only (uses libm pow/sqrtf exactly as the reference does).
inline BinomSetup binom_setup(int32_t n, float pp) {
	BinomSetup b;
	memset(&b, 0, sizeof(b));
	b.n = n;
	b.pp = pp;
	b.flip = (0.5 < (double)pp);
	float other = (float)(1.0 - (double)pp);
	b.p = (other < pp) ? other : pp;
	b.q = (float)(1.0 - (double)b.p);
	b.xnp = fmul_rn((float)n, b.p);
	if ((double)b.xnp < 30.0) {
		b.inversion = 1;
		b.qn = (float)pow((double)b.q, (double)n);
		b.r = fdiv_rn(b.p, b.q);
		b.g = fmul_rn(b.r, (float)(n + 1));
		return b;
	}
	b.ffm = fadd_rn(b.xnp, b.p);
	b.m = (int32_t)b.ffm;
	b.fm = (float)b.m;
	b.xnpq = fmul_rn(b.xnp, b.q);
	b.p1 = (float)((double)(int32_t)(2.195 * (double)sqrtf(b.xnpq) - 4.6 * (double)b.q) + 0.5);
	b.xm = (float)((double)b.fm + 0.5);
	b.xl = fsub_rn(b.xm, b.p1);
	b.xr = fadd_rn(b.xm, b.p1);
	b.c = (float)(0.134 + 20.5 / (15.3 + (double)b.fm));
	float al = fdiv_rn(fsub_rn(b.ffm, b.xl), fsub_rn(b.ffm, fmul_rn(b.xl, b.p)));
	b.xll = (float)((double)al * (1.0 + 0.5 * (double)al));
	al = fdiv_rn(fsub_rn(b.xr, b.ffm), fmul_rn(b.xr, b.q));
	b.xlr = (float)((double)al * (1.0 + 0.5 * (double)al));
	b.p2 = (float)((double)b.p1 * (1.0 + (double)b.c + (double)b.c));
	b.p3 = fadd_rn(b.p2, fdiv_rn(b.c, b.xll));
	b.p4 = fadd_rn(b.p3, fdiv_rn(b.c, b.xlr));
	return b;
}

QT_HD double stirling_tail(float s2, float s) {
	// (13860 - (462 - (132 - (99 - 140/s2)/s2)/s2)/s2) / s / 166320, all double
	double d2 = (double)s2;
	double t = ddiv_rn(140.0, d2);
	t = ddiv_rn(dsub_rn(99.0, t), d2);
	t = ddiv_rn(dsub_rn(132.0, t), d2);
	t = ddiv_rn(dsub_rn(462.0, t), d2);
	t = dsub_rn(13860.0, t);
	return ddiv_rn(ddiv_rn(t, (double)s), 166320.0);
}

// ---- the three acceptance bounds of BTPE's step 5 (ranlib.cpp:1274-1387): functions of ix (or k = |ix - m|) and the
// set-up only, so they can be tabulated per set-up (BinomTables) instead of being recomputed for every variate ----

// f(ix) / f(m) by the recurrence, ranlib.cpp:1284-1300
QT_HD float btpe_explicit_f(const BinomSetup &b, int32_t ix) {
	float f = 1.0f;
	float r = fdiv_rn(b.p, b.q);
	float g = fmul_rn((float)(b.n + 1), r);
	if (b.m < ix) {
		for (int32_t i = b.m + 1; i <= ix; i++) f = fmul_rn(f, fsub_rn(fdiv_rn(g, (float)i), r));
	} else if (ix < b.m) {
		for (int32_t i = ix + 1; i <= b.m; i++) f = fdiv_rn(f, fsub_rn(fdiv_rn(g, (float)i), r));
	}
	return f;
}

// squeeze bounds on log(v), ranlib.cpp:1315-1334: accept below lo = ynorm - amaxp, reject above hi = ynorm + amaxp
QT_HD void btpe_squeeze(const BinomSetup &b, int32_t k, float &lo, float &hi) {
	double dk = (double)k;
	double inner = dadd_rn(dmul_rn(dk, dadd_rn(ddiv_rn(dk, 3.0), 0.625)), 0.1666666666666);
	inner = dadd_rn(ddiv_rn(inner, (double)b.xnpq), 0.5);
	float amaxp = (float)dmul_rn((double)fdiv_rn((float)k, b.xnpq), inner);
	float ynorm = (float)ddiv_rn((double)(-(float)(k * k)), dmul_rn(2.0, (double)b.xnpq));
	lo = fsub_rn(ynorm, amaxp);
	hi = fadd_rn(ynorm, amaxp);
}

// the final bound on log(v) with Stirling's correction, ranlib.cpp:1340-1387
QT_HD float btpe_stirling_t(const BinomSetup &b, int32_t ix) {
	const int32_t n = b.n;
	float x1 = (float)(ix + 1);
	float f1 = (float)dadd_rn((double)b.fm, 1.0);
	float z = fsub_rn((float)(n + 1), b.fm);
	float w = (float)(n - ix + 1);
	float z2 = fmul_rn(z, z), x2 = fmul_rn(x1, x1), f2 = fmul_rn(f1, f1), w2 = fmul_rn(w, w);
	double d = (double)fmul_rn(b.xm, logf_glibc(fdiv_rn(f1, x1)));
	d = dadd_rn(d, dmul_rn(dadd_rn((double)(n - b.m), 0.5), (double)logf_glibc(fdiv_rn(z, w))));
	d = dadd_rn(d, (double)fmul_rn((float)(ix - b.m), logf_glibc(fdiv_rn(fmul_rn(w, b.p), fmul_rn(x1, b.q)))));
	d = dadd_rn(d, stirling_tail(f2, f1));
	d = dadd_rn(d, stirling_tail(z2, z));
	d = dadd_rn(d, stirling_tail(x2, x1));
	d = dadd_rn(d, stirling_tail(w2, w));
	return (float)d;
}

// Per BTPE set-up: the bounds above for every ix within BT_KT of the mode.  Built on the host (binom_tables); the
// device sampler binom_sample_tab reads them -- the values are those the direct computation gives, bit for bit.
static constexpr int BT_KT = 256;
struct BinomTables {
	float expl[41];               // btpe_explicit_f at ix = m - 20 + i
	float sq_lo[BT_KT];           // btpe_squeeze at k
	float sq_hi[BT_KT];
	float stir[2 * BT_KT - 1];    // btpe_stirling_t at ix = m - (BT_KT - 1) + i (0 where ix is outside [0, n])
};

inline void binom_tables(const BinomSetup &b, BinomTables &t) {
	memset(&t, 0, sizeof(t));
	if (b.inversion) return;
	for (int i = 0; i < 41; i++) t.expl[i] = btpe_explicit_f(b, b.m - 20 + i);
	for (int k = 0; k < BT_KT; k++) btpe_squeeze(b, k, t.sq_lo[k], t.sq_hi[k]);
	for (int i = 0; i < 2 * BT_KT - 1; i++) {
		const int64_t ix = (int64_t)b.m - (BT_KT - 1) + i;
		if (ix >= 0 && ix <= (int64_t)b.n) t.stir[i] = btpe_stirling_t(b, (int32_t)ix);
	}
}

// BTPE steps 1-4 (ranlib.cpp:1252-1273): the candidate ix and the v the acceptance test compares.  Returns 0: accepted
// outright (triangle), 1: go to step 5, 2: rejected, draw again.
QT_HD int btpe_candidate(const BinomSetup &b, float u, float &v, int32_t &ix) {
	if (u < b.p1) {
		ix = (int32_t)fadd_rn(fsub_rn(b.xm, fmul_rn(b.p1, v)), u);
		return 0;
	}
	if (u <= b.p2) {
		float x = fadd_rn(b.xl, fdiv_rn(fsub_rn(u, b.p1), b.c));
		double dv = dsub_rn(dadd_rn((double)fmul_rn(v, b.c), 1.0), (double)fdiv_rn(fabsf(fsub_rn(b.xm, x)), b.p1));
		v = (float)dv;
		if ((double)v <= 0.0 || 1.0 < (double)v) return 2;
		ix = (int32_t)x;
	} else if (u <= b.p3) {
		ix = (int32_t)fadd_rn(b.xl, fdiv_rn(logf_glibc(v), b.xll));
		if (ix < 0) return 2;
		v = fmul_rn(fmul_rn(v, fsub_rn(u, b.p2)), b.xll);
	} else {
		ix = (int32_t)fsub_rn(b.xr, fdiv_rn(logf_glibc(v), b.xlr));
		if (b.n < ix) return 2;
		v = fmul_rn(fmul_rn(v, fsub_rn(u, b.p3)), b.xlr);
	}
	return 1;
}

// Draws one Binomial(n, pp) variate from the stream.  `next()` must return the next
// r4_uni_01 value of the stream (and advance it).
template <class Next>
QT_HD int32_t binom_sample(const BinomSetup &b, Next &next) {
	const int32_t n = b.n;
	if (b.inversion) {
		for (;;) {
			int32_t ix = 0;
			float f = b.qn;
			float u = next();
			for (;;) {
				if (u < f) return b.flip ? n - ix : ix;
				if (110 < ix) break;
				u = fsub_rn(u, f);
				ix = ix + 1;
				f = fmul_rn(f, fsub_rn(fdiv_rn(b.g, (float)ix), b.r));
			}
		}
	}
	for (;;) {
		float u = fmul_rn(next(), b.p4);
		float v = next();
		int32_t ix;
		const int st = btpe_candidate(b, u, v, ix);
		if (st == 0) return b.flip ? n - ix : ix;
		if (st == 2) continue;
		int32_t k = ix - b.m;
		if (k < 0) k = -k;
		if (k <= 20 || dsub_rn(ddiv_rn((double)b.xnpq, 2.0), 1.0) <= (double)k) {
			if (v <= btpe_explicit_f(b, ix)) return b.flip ? n - ix : ix;
		} else {
			float lo, hi;
			btpe_squeeze(b, k, lo, hi);
			float alv = logf_glibc(v);
			if (alv < lo) return b.flip ? n - ix : ix;
			if (hi < alv) continue;
			if (alv <= btpe_stirling_t(b, ix)) return b.flip ? n - ix : ix;
		}
	}
}

// The same variate with step 5's bounds read from the set-up's tables (device resolver: thousands of candidates per
// window share one set-up, and the direct computation diverges badly across a warp).
template <class Next>
QT_HD int32_t binom_sample_tab(const BinomSetup &b, const BinomTables &t, Next &next) {
	if (b.inversion) return binom_sample(b, next);
	const int32_t n = b.n;
	const double khalf = dsub_rn(ddiv_rn((double)b.xnpq, 2.0), 1.0);
	for (;;) {
		float u = fmul_rn(next(), b.p4);
		float v = next();
		int32_t ix;
		const int st = btpe_candidate(b, u, v, ix);
		if (st == 0) return b.flip ? n - ix : ix;
		if (st == 2) continue;
		int32_t k = ix - b.m;
		if (k < 0) k = -k;
		if (k <= 20 || khalf <= (double)k) {
			const float f = k <= 20 ? t.expl[ix - b.m + 20] : btpe_explicit_f(b, ix);
			if (v <= f) return b.flip ? n - ix : ix;
		} else {
			float lo, hi;
			if (k < BT_KT) {
				lo = t.sq_lo[k];
				hi = t.sq_hi[k];
			} else {
				btpe_squeeze(b, k, lo, hi);
			}
			float alv = logf_glibc(v);
			if (alv < lo) return b.flip ? n - ix : ix;
			if (hi < alv) continue;
			const float tt = k < BT_KT ? t.stir[ix - b.m + (BT_KT - 1)] : btpe_stirling_t(b, ix);
			if (alv <= tt) return b.flip ? n - ix : ix;
		}
	}
}

}  // namespace qt
