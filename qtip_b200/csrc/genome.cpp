// genome.cpp -- host FASTA parse + 2-bit pack, window table.  See genome.h.
#include "genome.h"
#include <string.h>

namespace qt {

namespace {
enum { CL_N = 4, CL_SPACE = 5, CL_HDR = 6 };

struct Lut {
	uint8_t v[256];
	Lut() {
		for (int i = 0; i < 256; i++) v[i] = CL_N;  // dna_upper: everything else is N
		v['A'] = v['a'] = 0;
		v['C'] = v['c'] = 1;
		v['G'] = v['g'] = 2;
		v['T'] = v['t'] = 3;
		// isspace() in the C locale (fasta.cpp:113)
		v[' '] = v['\t'] = v['\n'] = v['\v'] = v['\f'] = v['\r'] = CL_SPACE;
		v['>'] = CL_HDR;
	}
};
const Lut lut;
inline bool is_space(uint8_t c) { return lut.v[c] == CL_SPACE; }
}  // namespace

void HostGenome::parse(const uint8_t *text, uint64_t n) {
	recs.clear();
	names.clear();
	file_bytes = n;
	// upper bound on records: count '>' bytes
	uint64_t nhdr = 0;
	for (const uint8_t *p = text, *e = text + n; p < e;) {
		const uint8_t *q = (const uint8_t *)memchr(p, '>', (size_t)(e - p));
		if (!q) break;
		nhdr++;
		p = q + 1;
	}
	const uint64_t cap_bases = n + 64 * (nhdr + 4);
	bases.assign((size_t)(cap_bases / 16 + 4), 0u);
	nmask.assign((size_t)(cap_bases / 32 + 4), 0u);

	// 64 bases of N padding in front: kernels may read up to 15 bases before a record
	uint64_t g = 64;      // next base slot
	nmask[0] = nmask[1] = 0xffffffffu;
	bool have = false;    // inside a record?
	uint64_t i = 0;
	auto close_record = [&]() {
		if (!have) return;
		RecDev &r = recs.back();
		r.len = g - r.gstart;
		// pad to the next multiple of 64 with N
		uint64_t end = (g + 63) & ~(uint64_t)63;
		if (end == g) end += 64;  // always leave a full N gap between records
		for (; g < end; g++) nmask[g >> 5] |= 1u << (g & 31);
	};
	auto open_record = [&](uint32_t name_off, uint32_t name_len) {
		close_record();
		RecDev r{};
		r.gstart = g;
		r.len = 0;
		r.name_off = name_off;
		r.name_len = name_len;
		recs.push_back(r);
		have = true;
	};
	while (i < n) {
		const uint8_t c = text[i];
		const uint8_t cl = lut.v[c];
		if (cl <= CL_N) {
			if (!have) open_record((uint32_t)names.size(), 0);
			if (cl == CL_N) nmask[g >> 5] |= 1u << (g & 31);
			else bases[g >> 4] |= (uint32_t)cl << (2 * (g & 15));
			g++;
			i++;
		} else if (cl == CL_SPACE) {
			i++;
		} else {  // '>': header to end of line (fasta.cpp:91-112)
			uint64_t s = ++i;
			while (i < n && !is_space(text[i])) i++;
			uint32_t off = (uint32_t)names.size();
			names.append((const char *)text + s, (size_t)(i - s));
			open_record(off, (uint32_t)(i - s));
			while (i < n && text[i] != '\n' && text[i] != '\r') i++;
		}
	}
	close_record();
	total_bases = g;
	bases.resize((size_t)((g + 15) / 16 + 4));
	nmask.resize((size_t)((g + 31) / 32 + 4));
}

std::vector<Window> HostGenome::windows(uint64_t chunksz, uint64_t olap) const {
	std::vector<Window> out;
	for (size_t r = 0; r < recs.size(); r++) {
		const uint64_t R = recs[r].len;
		uint64_t consumed = 0, carry = 0;
		while (consumed < R) {
			const uint64_t room = chunksz - carry;
			const uint64_t take = R - consumed < room ? R - consumed : room;
			const uint64_t retsz = carry + take;
			consumed += take;
			Window w;
			w.record = (uint32_t)r;
			w.retsz = (uint32_t)retsz;
			w.refoff = consumed - retsz;
			out.push_back(w);
			carry = retsz >= olap ? olap : 0;
		}
	}
	return out;
}

void HostGenome::fill_name16() {
	for (RecDev &r : recs) {
		memset(r.name16, 0, sizeof(r.name16));
		const uint32_t n = r.name_len < 16u ? r.name_len : 16u;
		if (n) memcpy(r.name16, names.data() + r.name_off, n);
	}
}

std::string HostGenome::blob() const {
	std::string b;
	uint64_t hdr[4] = {(uint64_t)recs.size(), (uint64_t)names.size(), total_bases, file_bytes};
	b.append((const char *)hdr, sizeof(hdr));
	b.append((const char *)recs.data(), recs.size() * sizeof(RecDev));
	b.append(names);
	return b;
}

bool HostGenome::from_blob(const void *p, uint64_t n) {
	if (n < 32) return false;
	uint64_t hdr[4];
	memcpy(hdr, p, 32);
	const uint64_t need = 32 + hdr[0] * sizeof(RecDev) + hdr[1];
	if (n < need) return false;
	recs.resize((size_t)hdr[0]);
	memcpy(recs.data(), (const char *)p + 32, (size_t)hdr[0] * sizeof(RecDev));
	names.assign((const char *)p + 32 + hdr[0] * sizeof(RecDev), (size_t)hdr[1]);
	total_bases = hdr[2];
	file_bytes = hdr[3];
	bases.clear();
	nmask.clear();
	return true;
}

}  // namespace qt
