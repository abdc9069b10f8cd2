// src/edit_xscript.h -- reference src/edit_xscript.h:15-25
#ifndef QTIP_B200_EDIT_XSCRIPT_H
#define QTIP_B200_EDIT_XSCRIPT_H
#include <cstddef>

// reference bases an edit transcript spans: soft clips, matches, mismatches, deletions
static inline size_t edit_xscript_to_rflen(const char *xs) {
	size_t n = 0;
	for (; *xs; ++xs) n += (*xs == 'S') | (*xs == '=') | (*xs == 'X') | (*xs == 'D');
	return n;
}
#endif
