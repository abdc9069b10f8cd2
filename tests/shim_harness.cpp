// tests/shim_harness.cpp -- TEST INFRASTRUCTURE: an extern "C" face over the drop-in headers in src/ (ds.h, rnglib.hpp)
// so that tests/test_shim_cpu.py can drive them without a GPU, the way oracle/ref_wrap.cpp drives the reference's.
#include <stdint.h>
#include "ds.h"
#include "rnglib.hpp"

extern int ignbin(int n, float pp);  // src/rnglib_shim.cpp (declared in src/ranlib.hpp)

extern "C" {
void shim_initialize() { initialize(); }
void shim_set_seed(int a, int b) { set_seed(a, b); }
void shim_get_state(int *a, int *b) {
	int x, y;
	get_state(x, y);
	*a = x;
	*b = y;
}
int shim_i4_uni() { return i4_uni(); }
float shim_r4_uni_01() { return r4_uni_01(); }
int shim_ignbin(int n, float p) { return ignbin(n, p); }
// ReservoirSampledEList<T>::add_part1 (src/ds.h; reference ds.h:437-446), n calls on capacity k
uint64_t shim_reservoir_offsets(uint64_t k, uint64_t n, uint64_t *off) {
	ReservoirSampledEList<int> r((size_t)k);
	for (uint64_t i = 0; i < n; i++) off[i] = (uint64_t)r.add_part1();
	return (uint64_t)r.list().size();
}
// ReservoirSampledEList<T>::add (ds.h:421-431): which of n offered items 0 .. n-1 are in the list afterwards
uint64_t shim_reservoir_add(uint64_t k, uint64_t n, int *kept) {
	ReservoirSampledEList<int> r((size_t)k);
	for (uint64_t i = 0; i < n; i++) r.add((int)i);
	for (size_t i = 0; i < r.list().size(); i++) kept[i] = r.list()[i];
	return (uint64_t)r.list().size();
}
}
