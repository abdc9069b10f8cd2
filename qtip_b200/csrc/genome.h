// genome.h -- FASTA loading and 2-bit packing (host side), window table.
//
// Replaces FastaChunkwiseParser (fasta.h:24-105, fasta.cpp:42-127): instead of streaming
// 128 KiB windows through a char buffer, the whole genome is packed once and stays
// resident in HBM; the windows the reference would have produced are reduced to a table
// of (record, refoff, retsz) computed arithmetically from the record lengths.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "common.h"

namespace qt {

struct HostGenome {
	std::vector<RecDev> recs;
	std::string names;            // arena of refids
	std::vector<uint32_t> bases;  // 2 bit/base
	std::vector<uint32_t> nmask;  // 1 bit/base
	uint64_t total_bases = 0;     // padded length of the concatenated space
	uint64_t file_bytes = 0;      // tot_fasta_len_ (simplesim.h:272-286): FASTA size in bytes

	// FASTA text -> records + packed arrays.  Header: '>' to end of line, refid up to the
	// first whitespace; every other non-whitespace byte is a base; bases before the first
	// header form a record with an empty name (fasta.cpp:72-124).
	void parse(const uint8_t *text, uint64_t n);

	// The windows fasta.cpp:42-127 hands out for (chunksz, olap): per record, the first
	// holds min(R, chunksz) bases, each later one re-presents the last `olap` bases of its
	// predecessor (fasta.cpp:63-70) plus up to chunksz-olap new ones, and exists only if
	// at least one new base remains.
	std::vector<Window> windows(uint64_t chunksz, uint64_t olap) const;

	inline uint32_t code_at(uint64_t g) const {  // 0..3, or 4 for N
		if ((nmask[g >> 5] >> (g & 31)) & 1u) return 4;
		return (bases[g >> 4] >> (2 * (g & 15))) & 3u;
	}

	// copies the first 16 bytes of every refid into its record (RecDev::name16)
	void fill_name16();

	// serialise / restore the record table (multi-GPU replication)
	std::string blob() const;
	bool from_blob(const void *p, uint64_t n);
};

}  // namespace qt
