// common.h -- data layouts shared by the host side of libqtipsim and its kernels.
//
// HBM layout (see DESIGN.md "Data layout"):
//   genome   : all FASTA records in one base coordinate space; record r occupies
//              [gstart[r], gstart[r]+len[r]), gstart 64-base aligned.
//              bases : 2 bit/base, base i in bits 2*(i%16) of u32 word i/16,
//                      A=0 C=1 G=2 T=3 ("ACGT"[k], simplesim.h:27).
//              nmask : 1 bit/base, base i in bit i%32 of u32 word i/32; 1 = N
//                      (anything dna_upper maps to 'N', fasta.cpp:18-35); padding is N.
//   templates: one TmplDesc per template and category, mates as MateDesc, with one
//              byte arena for quality strings and one for edit transcripts.
//   plan     : one PlanRec per simulated read / pair (category-major, window order).
//   output   : seven FASTQ byte streams, each one contiguous buffer.
#pragma once
#include <stdint.h>

namespace qt {

struct alignas(16) RecDev {
	uint64_t gstart;    // first base in the concatenated space
	uint64_t len;       // bases in the record
	uint32_t name_off;  // into the names arena
	uint32_t name_len;  // refid = header up to first whitespace (fasta.cpp:98-104)
	uint32_t pad[2];
	char name16[16];    // first 16 bytes of the refid (the kernels print short names from registers)
};                      // 48 bytes: three 16-byte loads

// One FASTA window as the reference's chunk parser would return it (fasta.cpp:42-127).
struct Window {
	uint32_t record;
	uint32_t retsz;     // <= chunksz
	uint64_t refoff;    // record offset of the window's first base
};

// One SimulatedRead::init worth of template data (template.h:97-104 / :218-230).
// 64 bytes in four 16-byte blocks; the simulate kernel reads blocks 0-2 with vector loads.
struct alignas(16) MateDesc {
	// block 0: compiled, output-oriented form (see templates.cpp)
	uint32_t qual_out_off;  // qual arena in OUTPUT orientation (reversed for '-'), 16-byte aligned
	uint32_t ops_off;       // draw-op list (X/I/S in transcript order), u32 entries
	uint32_t seg_off;       // reference-copy segments, ascending output position
	int32_t ref0;           // single != 0: reference offset of output position 0
	// block 1
	uint32_t len;        // read length = strlen(qual)
	uint32_t rflen;      // reference bases consumed (edit_xscript.h:15-25)
	uint16_t nops;
	uint16_t nseg;
	uint8_t fw;          // 1 = '+'
	uint8_t single;      // one segment covering the whole read (no indels: =, X, S only)
	uint8_t score_len;
	uint8_t pad;
	// block 2
	char score[12];      // "%d" of the score
	uint32_t ndraw;      // X+I+S ops (draws without rejections)
	// block 3: as given (generic kernel, planners)
	uint32_t qual_off;   // qual arena, read orientation as given
	uint32_t xs_off;     // transcript arena, ASCII ops
	uint32_t xs_len;     // number of ops
	uint32_t pad2;
};

// A maximal run of transcript ops that copy reference to read (=, X, S), in output
// orientation: output positions [out_start, out_end) come from reference offsets
// ref0 + (k - out_start) for '+' mates and ref0 - (k - out_start) for '-' mates
// (relative to the mate's leftmost reference base).
struct SegDev {
	uint16_t out_start, out_end;
	int32_t ref0;
};

// draw-op entry: bits 0-1 kind (0 = X: redraw while equal to the reference base,
// 1 = I or S: one draw), bits 2-16 output position, bits 17-31 reference offset (X)
static constexpr uint32_t OP_X = 0u, OP_IS = 1u;
static constexpr uint32_t COMPILED_MAX_LEN = 32767u;

struct alignas(16) TmplDesc {
	MateDesc m[2];       // [0] = rd1's source (aligned mate for U/B), [1] = mate 2 (C/D)
	// one 16-byte block
	uint32_t delta[2];   // offset of each mate's leftmost ref base from the drawn `off`
	                     // (simplesim.cpp:354-363; 0 for U/B)
	uint32_t opp_len;    // B: length of the random opposite mate (simplesim.h:89-117)
	char typ[2];         // "u" "b1" "b2" "c" "d"
	uint8_t typ_len;
	uint8_t aligned_is_mate1;  // B: mate_flag == '1' (simplesim.cpp:288)
	// --
	uint8_t nmates;      // 1 (U, B) or 2 (C, D)
	uint8_t pad[15];
};
static_assert(sizeof(MateDesc) == 64 && sizeof(TmplDesc) == 160 && sizeof(RecDev) == 48, "descriptor layout");

// One simulated read (U), bad-end pair (B) or pair (C/D).
struct PlanRec {
	uint32_t tmpl;       // template index within the category
	uint32_t record;
	uint64_t refoff;     // record offset of the drawn position (window.refoff + off)
	uint32_t s1, s2;     // REPLAY: rnglib state after the template and offset draws
	                     // PHILOX: unused (draws are keyed by the read's global index)
};

enum { CAT_U = 0, CAT_B = 1, CAT_C = 2, CAT_D = 3, NCAT = 4, NOUT = 7 };

// First output stream of each category (qtip_parse.cpp:1831-1837 order).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int cat_stream0(int cat) { return cat == 0 ? 0 : 2 * cat - 1; }

}  // namespace qt
