// src/fasta.h -- the reference's simplesim.h includes "fasta.h" (simplesim.h:16).  The chunked
// parser (fasta.h:24-105, fasta.cpp:42-127) is replaced by whole-genome loading and 2-bit
// packing behind the C ABI (qtipsim_load_fasta); nothing of it is needed on the host.
#ifndef QTIP_B200_FASTA_H
#define QTIP_B200_FASTA_H
#endif
