// mpcr_hostio.h -- host-side text work around the device path, in plain C++ (no CUDA): the STS line rules of
// MerPCR.load_sts_file (core/engine.py:216-251) and the output line format of MerPCR.search (core/engine.py:437-444).
// Shared by libmerpcr_b200.so and the test emulation; the extern "C" wrappers live in those translation units.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/merpcr_b200.h"

namespace mpcr {

inline bool sts_is_term(uint8_t c) { return c == 10 || c == 13; }
// the characters str.strip() removes, restricted to ASCII (engine.py:218)
inline bool sts_is_space(uint8_t c) { return c == 32 || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

// engine.py:304-322 for the two shapes every real file uses: "123" and "100-200".  Anything else (signs, blanks,
// underscores, unicode digits, overflow) returns -1 and the Python host applies int() itself.
inline int32_t sts_fast_size(const uint8_t* s, uint32_t n, int32_t dflt) {
    if (n == 0 || n > 19) return -1;
    uint64_t a = 0, b = 0;
    uint32_t i = 0, da = 0, db = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9' && da < 9) { a = a * 10 + (s[i] - '0'); ++i; ++da; }
    if (i == n) return da ? (a > 0 ? (int32_t)a : dflt) : -1;          // :316-319 (<= 0 -> default)
    if (s[i] != '-' || da == 0) return -1;
    ++i;
    while (i < n && s[i] >= '0' && s[i] <= '9' && db < 9) { b = b * 10 + (s[i] - '0'); ++i; ++db; }
    if (i != n || db == 0) return -1;
    return (int32_t)((a + b) / 2);                                      // :310 (midpoint; may be 0 like the reference)
}

// Parse STS text (ASCII).  Returns MPCR_OK, or MPCR_EOVERFLOW with *n_lines = lines needed.
// *bad_line != 0: the 1-based number of the first line with fewer than 4 fields (engine.py:226-230) -- parsing stops.
inline int sts_parse_impl(const uint8_t* text, uint64_t n, int32_t wordsize, int32_t default_size, mpcr_sts_line* lines,
                          uint32_t max_lines, uint32_t* n_lines, uint32_t* bad_line, uint32_t* short_primers,
                          uint32_t* flags) {
    *n_lines = 0; *bad_line = 0; *short_primers = 0; *flags = 0;
    if (n >= 0xFFFFFFFFull) return MPCR_EINVAL;
    for (uint64_t i = 0; i < n; ++i)
        if (text[i] >= 128) { *flags = 1; return MPCR_OK; }
    uint32_t out = 0, line_no = 0;
    uint64_t pos = 0;
    while (pos < n) {
        uint64_t e = pos;
        while (e < n && !sts_is_term(text[e])) ++e;
        ++line_no;
        uint64_t a = pos, b = e;
        while (a < b && sts_is_space(text[a])) ++a;
        while (b > a && sts_is_space(text[b - 1])) --b;
        // next line: "\r\n" counts as one terminator (universal newlines)
        pos = e + 1;
        if (e < n && text[e] == 13 && pos < n && text[pos] == 10) ++pos;
        if (a == b || text[a] == '#') continue;                                   // :219-220
        uint32_t f_off[5], f_len[5], nf = 0;
        uint64_t s = a;
        for (uint64_t i = a; i <= b; ++i) {
            if (i == b || text[i] == '\t') {
                if (nf < 5) { f_off[nf] = (uint32_t)s; f_len[nf] = (uint32_t)(i - s); }
                ++nf;
                s = i + 1;
            }
        }
        if (nf < 4) { *bad_line = line_no; *n_lines = out; return MPCR_OK; }      // :226-230
        if (f_len[1] < (uint32_t)wordsize || f_len[2] < (uint32_t)wordsize) { ++*short_primers; continue; }   // :241-243
        if (out < max_lines) {
            mpcr_sts_line& L = lines[out];
            L.line_no = line_no;
            L.id_off = f_off[0]; L.id_len = f_len[0];
            L.p1_off = f_off[1]; L.p1_len = f_len[1];
            L.p2_off = f_off[2]; L.p2_len = f_len[2];
            L.size_off = f_off[3]; L.size_len = f_len[3];
            L.alias_off = nf > 4 ? f_off[4] : 0; L.alias_len = nf > 4 ? f_len[4] : 0;
            L.pcr_size = sts_fast_size(text + f_off[3], f_len[3], default_size);
        }
        ++out;
    }
    *n_lines = out;
    return out > max_lines ? MPCR_EOVERFLOW : MPCR_OK;
}

// upper-cased primers of the accepted lines, back to back (primer1 then primer2 per line), + 2n+1 offsets
inline void sts_blob_impl(const uint8_t* text, const mpcr_sts_line* lines, uint32_t n_lines, uint8_t* blob, uint64_t* off) {
    uint64_t p = 0;
    for (uint32_t i = 0; i < n_lines; ++i) {
        const mpcr_sts_line& L = lines[i];
        off[2 * (uint64_t)i] = p;
        for (uint32_t k = 0; k < L.p1_len; ++k) { uint8_t c = text[L.p1_off + k]; blob[p++] = (c >= 'a' && c <= 'z') ? c - 32 : c; }
        off[2 * (uint64_t)i + 1] = p;
        for (uint32_t k = 0; k < L.p2_len; ++k) { uint8_t c = text[L.p2_off + k]; blob[p++] = (c >= 'a' && c <= 'z') ? c - 32 : c; }
    }
    off[2 * (uint64_t)n_lines] = p;
}

inline uint32_t fmt_u32(uint8_t* dst, uint32_t v) {
    uint8_t tmp[10];
    uint32_t k = 0;
    do { tmp[k++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    for (uint32_t i = 0; i < k; ++i) dst[i] = tmp[k - 1 - i];
    return k;
}

// engine.py:442: f"{label}\t{pos1+1}..{pos2+1}\t{id}\t{alias}\t({direct})\n" for every hit, in order.
// Returns the number of bytes the text needs; it is written only if it fits out_cap.
inline uint64_t format_hits_impl(const mpcr_hit* hits, uint64_t n, const uint8_t* text, const mpcr_sts_line* lines,
                                 const uint8_t* labels, const uint64_t* label_off, uint8_t* out, uint64_t out_cap) {
    uint64_t need = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mpcr_sts_line& L = lines[hits[i].rec >> 1];
        need += (label_off[hits[i].contig + 1] - label_off[hits[i].contig]) + L.id_len + L.alias_len + 32;
    }
    if (need > out_cap || !out) return need;
    uint64_t p = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mpcr_hit& h = hits[i];
        const mpcr_sts_line& L = lines[h.rec >> 1];
        const uint64_t l0 = label_off[h.contig], l1 = label_off[h.contig + 1];
        memcpy(out + p, labels + l0, l1 - l0); p += l1 - l0;
        out[p++] = '\t';
        p += fmt_u32(out + p, h.pos1 + 1);
        out[p++] = '.'; out[p++] = '.';
        p += fmt_u32(out + p, h.pos2 + 1);
        out[p++] = '\t';
        memcpy(out + p, text + L.id_off, L.id_len); p += L.id_len;
        out[p++] = '\t';
        memcpy(out + p, text + L.alias_off, L.alias_len); p += L.alias_len;
        out[p++] = '\t'; out[p++] = '('; out[p++] = (h.rec & 1u) ? '-' : '+'; out[p++] = ')'; out[p++] = '\n';
    }
    return p;
}

}  // namespace mpcr
