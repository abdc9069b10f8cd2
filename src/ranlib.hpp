// src/ranlib.hpp -- the reference's simplesim.h includes "ranlib.hpp" (simplesim.h:18); the only
// ranlib entry point on the path, ignbin (ranlib.cpp:1094), lives behind the C ABI now
// (qtip_b200/csrc/lecuyer.h: binom_setup / binom_sample).
#ifndef QTIP_B200_RANLIB_HPP
#define QTIP_B200_RANLIB_HPP
int ignbin(int n, float pp);
#endif
