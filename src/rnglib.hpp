// src/rnglib.hpp -- drop-in for the slice of the reference's rnglib API that the simulator's
// caller uses: initialize() (qtip_parse.cpp:1589), set_seed (:1673), r4_uni_01 (ds.h:444),
// plus get_state / i4_uni.  One process-global generator, as in rnglib.cpp:308-309; the
// stream itself is the shared restatement in qtip_b200/csrc/lecuyer.h.
#ifndef QTIP_B200_RNGLIB_HPP
#define QTIP_B200_RNGLIB_HPP

void initialize();                       // rnglib.cpp:941-1009
void set_seed(int cg1, int cg2);         // rnglib.cpp:1674-1755 (exit(1) on range error)
void get_state(int &cg1, int &cg2);      // rnglib.cpp:516-570
int i4_uni();                            // rnglib.cpp:573-683
float r4_uni_01();                       // rnglib.cpp:1444-1505

#endif
