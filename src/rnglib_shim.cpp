// src/rnglib_shim.cpp -- the process-global generator behind src/rnglib.hpp.
#include <stdio.h>
#include <stdlib.h>
#include "rnglib.hpp"
#include "../qtip_b200/csrc/lecuyer.h"

static qt::LState g_state = {1234567890u, 123456789u};

void initialize() { g_state = qt::LState{1234567890u, 123456789u}; }

void set_seed(int cg1, int cg2) {
	if (cg1 < 1 || (unsigned)cg1 >= qt::LM1) {
		fprintf(stderr, "\nSET_SEED - Fatal error!\n  Input parameter CG1 out of bounds.\n");
		exit(1);
	}
	if (cg2 < 1 || (unsigned)cg2 >= qt::LM2) {
		fprintf(stderr, "\nSET_SEED - Fatal error!\n  Input parameter CG2 out of bounds.\n");
		exit(1);
	}
	g_state = qt::LState{(unsigned)cg1, (unsigned)cg2};
}

void get_state(int &cg1, int &cg2) {
	cg1 = (int)g_state.s1;
	cg2 = (int)g_state.s2;
}

int i4_uni() {
	g_state = qt::lstep(g_state);
	return qt::lout(g_state);
}

float r4_uni_01() { return qt::lfloat(i4_uni()); }

int ignbin(int n, float pp) {
	if (pp <= 0.0f || 1.0f <= pp) {
		fprintf(stderr, "\nIGNBIN - Fatal error!\n  PP is out of range.\n");
		exit(1);
	}
	qt::BinomSetup b = qt::binom_setup(n, pp);
	auto nx = []() { return r4_uni_01(); };
	return qt::binom_sample(b, nx);
}
