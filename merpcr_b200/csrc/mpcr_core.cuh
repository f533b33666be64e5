// mpcr_core.cuh -- data layout + the per-position semantics of the STS search, shared by every kernel.
//
// Everything here is __host__ __device__ so that tests/host_emul.cpp can run the very same verification
// logic serially on a CPU against the oracle (test infrastructure only -- the product never runs it on the
// host).  Reference citations are relative to /root/reference/src/merpcr/.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define MPCR_HD __host__ __device__ __forceinline__
#else
#define MPCR_HD inline
#endif

namespace mpcr {

// ---------------------------------------------------------------------------------------------------------
// Layout
// ---------------------------------------------------------------------------------------------------------

static constexpr uint64_t kLane1 = 0x1111111111111111ull;  // LSB of every nibble
#ifndef MPCR_SCAN_THREADS
#define MPCR_SCAN_THREADS 512
#endif
#ifndef MPCR_SCAN_ILP
#define MPCR_SCAN_ILP 4
#endif
#ifndef MPCR_SCAN_QCAP
#define MPCR_SCAN_QCAP 256
#endif
static constexpr int kScanThreads = MPCR_SCAN_THREADS;     // threads of one scanner CTA (one CTA per SM)
static constexpr int kPosPerThread = 64;                   // hash positions per lane and unit
static constexpr int kTileBases = 32768;                   // hash positions per tile descriptor (multiple of 2048)
static constexpr int kTagBases = 8;                        // bases after the seed carried inline in the table

// One record = one strand of one STS line (core/models.py:17-29 + engine.py:253-281).
struct RecMeta {
    uint32_t pcr_size;  // engine.py:245-247 adjusted, clamped to 2^31-1
    uint32_t key;       // seed W-mer, little-endian digits (base i of the word in bits [2i,2i+1])
    uint32_t hash_be;   // the reference's hash value (engine.py:350), for read-back only
    uint32_t p1_word;   // offset (uint64 units) of primer1's words in the primer blob
    uint32_t p2_word;
    uint16_t len1, len2;
    uint16_t hash_off;  // engine.py:339-353
    uint16_t flags;     // bit0: the reference inserts the record (it has a clean W-mer); bit1: it is in THIS table
    uint32_t tag;       // primer1 bases right after the seed: 2-bit codes in bits [0,16), compare mask in [16,32)
};
static_assert(sizeof(RecMeta) == 32, "RecMeta layout");

// A tile = up to kTileBases consecutive hash positions of ONE contig.
struct TileDesc {
    int64_t gbase;     // plane-relative base index of the tile's first base (multiple of 128)
    uint32_t contig;   // contig index
    uint32_t lstart;   // contig-local coordinate of the tile's first base
    uint32_t length;   // true contig length L
    uint32_t nbases;   // bases of the contig inside this tile
};
static_assert(sizeof(TileDesc) == 24, "TileDesc layout");

struct SearchParams {
    int W, M, N, X, iupac;
    // Block tables (mpcr_ctx_set_seed_blocks): the key is the reference's seed (the first W - block letters) plus ONE block
    // of `block` letters that starts `gap` letters behind the seed.  gap = block = 0: an ordinary (contiguous) key.
    int gap = 0, block = 0;
};

MPCR_HD uint32_t wmask_of(int W) { return W >= 16 ? 0xFFFFFFFFu : ((1u << (2 * W)) - 1u); }
MPCR_HD uint32_t wmask_bits(int W) { return (1u << W) - 1u; }  // W one-bits (W <= 16)

// Three-input bitwise function by truth table (index = a << 2 | b << 1 | c), the GPU's LOP3.
template <uint32_t LUT>
MPCR_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < 8; ++i)
        if ((LUT >> i) & 1u) r |= ((i & 4u) ? a : ~a) & ((i & 2u) ? b : ~b) & ((i & 1u) ? c : ~c);
    return r;
#endif
}

// FASTA keep set (io/fasta.py:60: ACGTBDHKMNRSVWXY in either case), FOUR bytes at a time: bit 0 of every byte of the
// result says "kept".  A byte is kept iff b7 = 0, b6 = 1 and T[b4..b0] with T = the letter set as a 32-entry table (b5 is
// the case bit); the table is evaluated bit-sliced -- four 3-input sub-tables over (b2, b1, b0), muxed by b3 and b4 --
// on shifted copies of the word, 16 instructions for 4 bytes where a per-byte table look-up takes ~32.
MPCR_HD uint32_t fasta_keep_flags4(uint32_t w) {
    constexpr uint32_t T = 0x03DC699Eu;   // bit v: letter '@' + v is in the set (A=1 B=2 C=3 D=4 G=7 H=8 K=11 M=13 N=14 R=18 S=19 T=20 V=22 W=23 X=24 Y=25)
    const uint32_t x1 = w >> 1, x2 = w >> 2, x3 = w >> 3, x4 = w >> 4, x6 = w >> 6, x7 = w >> 7;
    const uint32_t h0 = lop3<(T >> 0) & 0xFFu>(x2, x1, w), h1 = lop3<(T >> 8) & 0xFFu>(x2, x1, w);
    const uint32_t h2 = lop3<(T >> 16) & 0xFFu>(x2, x1, w), h3 = lop3<(T >> 24) & 0xFFu>(x2, x1, w);
    const uint32_t m0 = lop3<0xCAu>(x3, h1, h0), m1 = lop3<0xCAu>(x3, h3, h2);   // 0xCA: a ? b : c
    const uint32_t f = lop3<0xCAu>(x4, m1, m0);
    return lop3<0x40u>(f, x6, x7) & 0x01010101u;                               // 0x40: a & b & ~c
}
// bytes equal to `c` (c replicated into every byte of c4): bit 7 of the byte is set for the LOWEST such byte and possibly
// for bytes above it (borrow) -- a trigger for a per-byte look, exact as "any?"
MPCR_HD uint32_t bytes_equal_trigger4(uint32_t w, uint32_t c4) {
    const uint32_t t = w ^ c4;
    return (t - 0x01010101u) & ~t & 0x80808080u;
}

// ---------------------------------------------------------------------------------------------------------
// Alphabet (engine.py:99-172).  The genome-side LUT is built by the host (merpcr_b200/alphabet.py) because it
// depends on the mode; the primer-side tables below are fixed.
// ---------------------------------------------------------------------------------------------------------

// engine.py:102-109 scode: 0..3 or 4 (= AMBIG)
MPCR_HD int scode_of(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 4;
    }
}

// engine.py:112-135,359: complement, unknown -> 'N'
MPCR_HD uint8_t complement_of(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A';
        case 'B': return 'V'; case 'D': return 'H'; case 'H': return 'D'; case 'K': return 'M'; case 'M': return 'K';
        case 'N': return 'N'; case 'R': return 'Y'; case 'S': return 'S'; case 'V': return 'B'; case 'W': return 'W';
        case 'X': return 'X'; case 'Y': return 'R';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a';
        case 'b': return 'v'; case 'd': return 'h'; case 'h': return 'd'; case 'k': return 'm'; case 'm': return 'k';
        case 'n': return 'n'; case 'r': return 'y'; case 's': return 's'; case 'v': return 'b'; case 'w': return 'w';
        case 'x': return 'x'; case 'y': return 'r';
        default: return 'N';
    }
}

// digit reversal: big-endian 2-bit pack (reference) <-> little-endian pack (plane2 order)
MPCR_HD uint32_t reverse_digits(uint32_t h, int W) {
    uint32_t r = 0;
    for (int i = 0; i < W; ++i) { r = (r << 2) | (h & 3u); h >>= 2; }
    return r;
}

// engine.py:331-355 _hash_value on an (already upper-cased) primer given as a char accessor.
// Returns the hash offset or -1; *hash_be gets the reference's value.
template <class CharAt>
MPCR_HD int first_clean_word(CharAt at, int len, int W, uint32_t* hash_be) {
    uint32_t h = 0, mask = wmask_of(W);
    int run = 0;
    for (int i = 0; i < len; ++i) {
        int code = scode_of(at(i));
        if (code > 3) { run = 0; continue; }
        h = ((h << 2) | (uint32_t)code) & mask;
        if (++run >= W) { *hash_be = h; return i - W + 1; }
    }
    *hash_be = 0;
    return -1;
}

// Seed extension (exact-match searches only, see mpcr_ctx_set_seed_extension): can the seed at offset ho be
// lengthened to w_ext plain A/C/G/T letters inside the primer?  Returns the extended word (little-endian digits).
template <class CharAt>
MPCR_HD bool extended_seed(CharAt at, int len, int ho, int w_ext, uint32_t* key_le) {
    if (ho < 0 || ho + w_ext > len) return false;
    uint32_t k = 0;
    for (int i = 0; i < w_ext; ++i) {
        const uint8_t c = at(ho + i);
        uint32_t code;
        if (c == 'A') code = 0; else if (c == 'C') code = 1; else if (c == 'G') code = 2; else if (c == 'T') code = 3;
        else return false;
        k |= code << (2 * i);
    }
    *key_le = k;
    return true;
}

// Block tables (searches that allow N >= 1 mismatches, see mpcr_ctx_set_seed_blocks).  The reference finds a site only
// where the seed word matches exactly and at most N of the primer's OTHER letters differ, so of N + 1 disjoint blocks of
// letters behind the seed at least one matches exactly: table i is keyed on seed + block i, and every site is found by
// the table of its first exact block.  A record takes part iff all `span` = W + n_blocks * block letters from its hash
// offset exist and are plain A/C/G/T.  Returns the key of the block that starts `gap` letters behind the seed
// (little-endian digits: seed in the low 2W bits, the block above it).
template <class CharAt>
MPCR_HD bool blocked_seed(CharAt at, int len, int ho, int W, int block, int gap, int span, uint32_t* key_le) {
    if (ho < 0 || ho + span > len) return false;
    uint32_t all = 0;   // codes of all span (<= 16) letters
    for (int i = 0; i < span; ++i) {
        const uint8_t c = at(ho + i);
        uint32_t code;
        if (c == 'A') code = 0; else if (c == 'C') code = 1; else if (c == 'G') code = 2; else if (c == 'T') code = 3;
        else return false;
        all |= code << (2 * i);
    }
    *key_le = (all & wmask_of(W)) | (((all >> (2 * (W + gap))) & wmask_of(block)) << (2 * W));
    return true;
}
// The same key out of a register that holds the 16 bases from a hash position on (2 bits each, little-endian); wk = key
// width W + block.  Bits above 2 * wk are garbage.
MPCR_HD uint32_t gap_key_raw(uint32_t x, uint32_t seed_mask, int gap) { return (x & seed_mask) | ((x >> (2 * gap)) & ~seed_mask); }

// Position sampling (exact searches only, see mpcr_ctx_set_sampling): with no mismatch allowed EVERY window of the
// primer matches where the primer does, so a table may hold the w-letter windows at offsets ho .. ho+S-1 and the
// scanner may look at every S-th position only.  True iff those S windows exist and hold plain A/C/G/T letters only.
template <class CharAt>
MPCR_HD bool sampleable_seed(CharAt at, int len, int ho, int w, int S) {
    if (ho < 0 || S < 1 || ho + S - 1 + w > len) return false;
    for (int i = ho; i < ho + S - 1 + w; ++i) {
        const uint8_t c = at(i);
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
    }
    return true;
}

// Encode a primer into nibble words + aux words.  lut[c] = nibble | never_match<<4 | zero_code_char<<5.
// dst[0..nw) nibbles, dst[nw..2nw) aux (bit0 of nibble i = never match, bit1 = "is the zero-code character").
template <class CharAt>
MPCR_HD void encode_primer(CharAt at, int len, const uint8_t* lut, uint64_t* dst) {
    int nw = (len + 15) >> 4;
    for (int w = 0; w < nw; ++w) {
        uint64_t q = 0, aux = 0;
        int cnt = len - 16 * w; if (cnt > 16) cnt = 16;
        for (int j = 0; j < cnt; ++j) {
            uint8_t e = lut[at(16 * w + j)];
            q |= (uint64_t)(e & 15u) << (4 * j);
            aux |= (uint64_t)((e >> 4) & 3u) << (4 * j);
        }
        dst[w] = q;
        dst[nw + w] = aux;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Plane access
// ---------------------------------------------------------------------------------------------------------

// 16 nibbles starting at plane-relative base b (any alignment).  Reads word b/16 and, if unaligned, b/16+1.
MPCR_HD uint64_t fetch16(const uint64_t* p4, int64_t b) {
    uint64_t i = (uint64_t)b >> 4;
    unsigned sh = ((unsigned)b & 15u) * 4u;
    uint64_t lo = p4[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (p4[i + 1] << (64u - sh));
}

// the W-mer starting at plane-relative base b, little-endian digits
MPCR_HD uint32_t extract_key(const uint64_t* p2, int64_t b, uint32_t wmask) {
    uint64_t i = (uint64_t)b >> 5;
    unsigned sh = ((unsigned)b & 31u) * 2u;
    uint64_t lo = p2[i];
    uint64_t v = sh ? ((lo >> sh) | (p2[i + 1] << (64u - sh))) : lo;
    return (uint32_t)v & wmask;
}

MPCR_HD uint64_t lanes_below(int n) {  // nibble-LSB mask of lanes [0, n), n in [0,16]
    return n >= 16 ? kLane1 : (n <= 0 ? 0ull : (((1ull << (4 * n)) - 1ull) & kLane1));
}

MPCR_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// ---------------------------------------------------------------------------------------------------------
// engine.py:599-642 _compare_seqs on packed data.
//   genome : 16 bases / uint64 IUPAC-mask nibbles starting at plane-relative base gb
//   primer : nibble words pw[0..nw) + aux words pw[nw..2nw)
//   plus   : strand "+" (3' protected zone = last X positions) else "-" (first X positions)   (:609-611)
// match test (:614-631): non-IUPAC -> identical letter == identical nibble (the nibble code is a bijection on
// the sequence alphabet); IUPAC -> mask AND != 0, or both are the zero-code letter (X matches only X).
// A primer letter that can match nothing in the genome carries the never_match aux bit.
// ---------------------------------------------------------------------------------------------------------
MPCR_HD bool compare_primer(const uint64_t* p4, int64_t gb, const uint64_t* pw, int len, bool plus,
                            const SearchParams& prm) {
    const int nw = (len + 15) >> 4;
    int prot_lo, prot_hi;  // protected index range [prot_lo, prot_hi)
    if (plus) { prot_lo = len - prm.X; if (prot_lo < 0) prot_lo = 0; prot_hi = len; }
    else { prot_lo = 0; prot_hi = prm.X < len ? prm.X : len; }
    int mism = 0;
    for (int w = 0; w < nw; ++w) {
        const uint64_t g = fetch16(p4, gb + 16 * w);
        const uint64_t q = pw[w], aux = pw[nw + w];
        uint64_t mis;
        if (prm.iupac) {
            uint64_t a = g & q;
            uint64_t nz = a | (a >> 1) | (a >> 2) | (a >> 3);
            uint64_t gz = ~(g | (g >> 1) | (g >> 2) | (g >> 3));
            mis = ~(nz | (gz & (aux >> 1)));
        } else {
            uint64_t x = g ^ q;
            mis = x | (x >> 1) | (x >> 2) | (x >> 3);
        }
        mis = (mis | aux) & lanes_below(len - 16 * w);
        const uint64_t prot = lanes_below(prot_hi - 16 * w) & ~lanes_below(prot_lo - 16 * w);
        if (mis & prot) return false;                    // :635-636
        mism += popc64(mis);
        if (mism > prm.N) return false;                  // :638-640
    }
    return true;
}

// The same compare for primers of up to 32 bases with everything that does not depend on the genome position
// hoisted: the verifier compares ONE primer at up to 2M+1 positions (engine.py:543-593).
struct PrimerView {
    uint64_t q[2], aux[2], lanes[2], prot[2];
    int nw;  // words in use (1 or 2); 0 = primer longer than 32 bases, use compare_primer
};
MPCR_HD PrimerView make_primer_view(const uint64_t* pw, int len, bool plus, const SearchParams& prm) {
    PrimerView v;
    v.nw = len <= 32 ? ((len + 15) >> 4) : 0;
    int prot_lo, prot_hi;
    if (plus) { prot_lo = len - prm.X; if (prot_lo < 0) prot_lo = 0; prot_hi = len; }
    else { prot_lo = 0; prot_hi = prm.X < len ? prm.X : len; }
    for (int w = 0; w < 2; ++w) {
        const bool on = w < v.nw;
        v.q[w] = on ? pw[w] : 0ull;
        v.aux[w] = on ? pw[v.nw + w] : 0ull;
        v.lanes[w] = on ? lanes_below(len - 16 * w) : 0ull;
        v.prot[w] = on ? (lanes_below(prot_hi - 16 * w) & ~lanes_below(prot_lo - 16 * w)) : 0ull;
    }
    return v;
}
MPCR_HD uint64_t mismatch_lanes(uint64_t g, uint64_t q, uint64_t aux, uint64_t lanes, int iupac) {
    uint64_t mis;
    if (iupac) {
        const uint64_t a = g & q;
        const uint64_t nz = a | (a >> 1) | (a >> 2) | (a >> 3);
        const uint64_t gz = ~(g | (g >> 1) | (g >> 2) | (g >> 3));
        mis = ~(nz | (gz & (aux >> 1)));
    } else {
        const uint64_t x = g ^ q;
        mis = x | (x >> 1) | (x >> 2) | (x >> 3);
    }
    return (mis | aux) & lanes;
}
// 8 nibbles starting at plane-relative base b, from the 32-bit view of plane4
MPCR_HD uint32_t fetch8(const uint64_t* p4, int64_t b) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p4);
    const uint64_t i = (uint64_t)b >> 3;
    const unsigned sh = ((unsigned)b & 7u) * 4u;
    const uint32_t lo = w[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[i + 1] << (32u - sh));
}
MPCR_HD uint32_t mismatch_lanes32(uint32_t g, uint32_t q, uint32_t aux, uint32_t lanes, int iupac) {
    uint32_t mis;
    if (iupac) {
        const uint32_t a = g & q;
        const uint32_t nz = a | (a >> 1) | (a >> 2) | (a >> 3);
        const uint32_t gz = ~(g | (g >> 1) | (g >> 2) | (g >> 3));
        mis = ~(nz | (gz & (aux >> 1)));
    } else {
        const uint32_t x = g ^ q;
        mis = x | (x >> 1) | (x >> 2) | (x >> 3);
    }
    return (mis | aux) & lanes;
}
MPCR_HD bool compare_view(const uint64_t* p4, int64_t gb, const PrimerView& v, const SearchParams& prm) {
    // the first 8 bases decide almost every position of the mate window (one 32-bit word): leave as early as the
    // reference's per-character loop does
    {
        const uint32_t m8 = mismatch_lanes32(fetch8(p4, gb), (uint32_t)v.q[0], (uint32_t)v.aux[0], (uint32_t)v.lanes[0],
                                             prm.iupac);
#ifdef __CUDA_ARCH__
        const int n8 = __popc(m8);
#else
        const int n8 = __builtin_popcount(m8);
#endif
        if ((m8 & (uint32_t)v.prot[0]) || n8 > prm.N) return false;   // :635-640
    }
    const uint64_t m0 = mismatch_lanes(fetch16(p4, gb), v.q[0], v.aux[0], v.lanes[0], prm.iupac);
    const int n0 = popc64(m0);
    if ((m0 & v.prot[0]) || n0 > prm.N) return false;           // :635-640
    if (v.nw < 2) return true;
    const uint64_t m1 = mismatch_lanes(fetch16(p4, gb + 16), v.q[1], v.aux[1], v.lanes[1], prm.iupac);
    if (m1 & v.prot[1]) return false;
    return n0 + popc64(m1) <= prm.N;
}

// compare_view's first check (the 8-base word that decides almost every position of a mate window) for m <= 32
// CONSECUTIVE positions starting at plane-relative base gb, in rolling form: six 32-bit words of plane4 are loaded once,
// aligned to gb, and every position's 8 nibbles are one funnel shift away -- where the per-position form pays address
// arithmetic and two loads each time.  Bit t of the result: position gb + t passes the check (exactly the same test, so
// a position that fails it fails compare_view).  Reads 48 bases from gb on.
MPCR_HD uint32_t mate_precheck32(const uint64_t* p4, int64_t gb, uint32_t m, const PrimerView& v, const SearchParams& prm) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p4);
    const uint64_t i0 = (uint64_t)gb >> 3;
    const unsigned a = ((unsigned)gb & 7u) * 4u;
    uint32_t r[6], al[5];
#pragma unroll
    for (int k = 0; k < 6; ++k) r[k] = w[i0 + k];
#pragma unroll
    for (int k = 0; k < 5; ++k) al[k] = a ? ((r[k] >> a) | (r[k + 1] << (32u - a))) : r[k];
    const uint32_t q8 = (uint32_t)v.q[0], aux8 = (uint32_t)v.aux[0], lanes8 = (uint32_t)v.lanes[0], prot8 = (uint32_t)v.prot[0];
    uint32_t out = 0;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        if ((uint32_t)t >= m) break;
        const int k = t >> 3, sft = 4 * (t & 7);
        const uint32_t x = sft ? ((al[k] >> sft) | (al[k + 1] << (32 - sft))) : al[k];
        const uint32_t m8 = mismatch_lanes32(x, q8, aux8, lanes8, prm.iupac);
#ifdef __CUDA_ARCH__
        const int n8 = __popc(m8);
#else
        const int n8 = __builtin_popcount(m8);
#endif
        if (!((m8 & prot8) || n8 > prm.N)) out |= 1u << t;
    }
    return out;
}

// Block tables: is one of the blocks IN FRONT of this table's block (letters [ws, ws + gap) behind the hash offset, in
// pieces of `block`) identical to the genome?  Then an earlier table finds this site and this one must not report it again.
// The blocks hold plain A/C/G/T primer letters, and the nibble code is a bijection on the sequence alphabet, so "identical
// letters" is nibble equality -- exactly the condition under which the earlier table's key matches.
// gb = plane-relative base of the primer's first letter, pw = its nibble words, ws = the reference's word size.
MPCR_HD bool earlier_block_exact(const uint64_t* p4, int64_t gb, const uint64_t* pw, int hash_off, int ws, int block, int gap) {
    for (int o = 0; o + block <= gap; o += block) {
        bool same = true;
        for (int i = 0; i < block; ++i) {
            const int pi = hash_off + ws + o + i;
            const int64_t gi = gb + pi;
            const uint32_t pn = (uint32_t)(pw[pi >> 4] >> (4 * (pi & 15))) & 15u;
            const uint32_t gn = (uint32_t)(p4[(uint64_t)gi >> 4] >> (4 * ((unsigned)gi & 15u))) & 15u;
            same = same && pn == gn;
        }
        if (same) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------
// engine.py:486-489 + 507-597: one bucket entry met at hash position p of a contig of true length L.
// gcontig = plane-relative base index of the contig's first base (may be negative for a shard that starts
// inside the contig; every base actually touched lies inside the shard + halo).
// emit(pos1, pos2, rank) is called once per matching delta, in the reference's order.
// ---------------------------------------------------------------------------------------------------------
template <class Emit>
MPCR_HD void verify_record(const uint64_t* p4, int64_t gcontig, int64_t L, int64_t p, const RecMeta& m,
                           const uint64_t* pwords, const SearchParams& prm, Emit&& emit) {
    const int l1 = m.len1, l2 = m.len2;
    const int64_t k = p - (int64_t)m.hash_off;                                   // :486
    if (k < 0 || k + l1 > L) return;                                              // :487
    if (!compare_primer(p4, gcontig + k, pwords + m.p1_word, l1, true, prm)) return;   // :515
    if (prm.gap > 0 && earlier_block_exact(p4, gcontig + k, pwords + m.p1_word, m.hash_off, prm.W - prm.block, prm.block, prm.gap))
        return;                                                                   // an earlier block table reports this site
    const int64_t avail = L - (k + l1);                                           // :521
    if (avail < l2) return;                                                       // :524
    int64_t E = (int64_t)m.pcr_size, hi, lo;
    if (E > L - k) { E = L - k; hi = 0; }                                         // :531-533 (end clamp, Q5)
    else { hi = L - k - E; if (hi > prm.M) hi = prm.M; }                          // :535
    lo = E - l1 - l2; if (lo > prm.M) lo = prm.M; if (lo < 0) lo = 0;             // :538-540
    const uint64_t* q2 = pwords + m.p2_word;
    const PrimerView v2 = make_primer_view(q2, l2, false, prm);
    auto mate = [&](int64_t q) {
        return v2.nw ? compare_view(p4, gcontig + q, v2, prm) : compare_primer(p4, gcontig + q, q2, l2, false, prm);
    };
    const int64_t p2 = k + E - l2;                                                // :543
    if (mate(p2)) emit(k, p2 + l2 - 1, (uint32_t)0);
    const int64_t mx = lo > hi ? lo : hi;
    for (int64_t i = 1; i <= mx; ++i) {                                           // :563
        if (i <= lo && mate(p2 - i)) emit(k, p2 - i + l2 - 1, (uint32_t)(2 * i - 1));   // :565-578
        if (i <= hi && mate(p2 + i)) emit(k, p2 + i + l2 - 1, (uint32_t)(2 * i));       // :581-593
    }
}

// ---------------------------------------------------------------------------------------------------------
// Table probing
//
// Level 1 (shared memory): a blocked Bloom filter over the seed keys, TWO bits per key inside ONE 32-bit word
//   word index = mulhi(key * cw, n_words), cw = golden-ratio constant << (32 - 2W)  (so garbage above the key's
//   2W bits cancels and the scanner never has to mask the funnel-shifted register), bit positions taken from
//   two disjoint 5-bit fields of the key so that a left shift by the raw register brings each to the MSB.
// Level 2 (L2-resident, one 16-byte gather): open-addressed slot table keyed by the exact seed, carrying the
//   record (or bucket start) and the inline tag of the single record.
// Level 3: CSR bucket entries {record | last << 31, tag} for seeds shared by several records.
// ---------------------------------------------------------------------------------------------------------

MPCR_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MPCR_HD uint32_t filter_mul(int W) { return 0x9E3779B1u << (32 - 2 * W); }
MPCR_HD uint32_t filter_word(uint32_t key_raw, uint32_t cw, uint32_t n_words) { return mulhi32(key_raw * cw, n_words); }
// bit positions: 31 - (key & 31) and, when the key has at least 11 bits (W >= 6), 31 - ((key >> 6) & 31).
// (key >> 6) is the raw register of the hash position three bases further on, so the scanner gets the second
// shift amount for free.)
MPCR_HD uint32_t filter_bits_of(uint32_t key, int W) {
    const uint32_t s1 = key & 31u, s2 = W >= 6 ? ((key >> 6) & 31u) : s1;
    return (0x80000000u >> s1) | (0x80000000u >> s2);
}
MPCR_HD bool filter_pass(uint32_t word, uint32_t key, int W) {
    const uint32_t m = filter_bits_of(key, W);
    return (word & m) == m;
}

// Linear filter map (tables keyed on 11 letters, the reference's default word size).  The scanner's stage 1 pays one
// issue slot per warp instruction whatever pipe executes it, and FP32 instructions come in a packed form (FFMA2: two
// lanes of work per instruction), so this map computes the word index in floating point: the 22-bit key goes into a
// float's mantissa (one LOP3: 2^23 + key, exact), ONE fma scales it to [2^23, 2^23 + n_words) -- at that magnitude a
// float's ulp is 1, so the rounded result's low mantissa bits ARE the word index -- and a second fma turns the index
// into the shared-memory byte address (a denormal whose bit pattern is the address).  The index is monotone in the key
// (keys of one word share their upper letters and differ in the low ~6.7 bits), so the two bit positions come from key
// bits [0,5) and [2,7): 122 of 128 neighbouring keys get a bit pair of their own.  False-positive rate with 2*10^5 random keys
// in 4*10^4 words: 6.2 % against 5.8 % for the multiplicative map (simulated; the scan kernel is 4.6 % faster with it).
#ifndef MPCR_LINEAR_FILTER
#define MPCR_LINEAR_FILTER 1
#endif
static constexpr int kLinearW = 11;
static constexpr uint32_t kLinearExp = 0x4B000000u;   // bit pattern of 2^23: mantissa = integer offset
MPCR_HD float bits_as_float(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
MPCR_HD uint32_t float_as_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
// float bits of the word index + kLinearExp (the scanner folds the subtraction into its address arithmetic)
MPCR_HD uint32_t filter_word_linear_raw(uint32_t key, float scale, float bias) {
    const float fk = bits_as_float(kLinearExp | key);   // 2^23 + key, key < 2^22
#ifdef __CUDA_ARCH__
    return __float_as_uint(__fmaf_rn(fk, scale, bias));
#else
    return float_as_bits(fmaf(fk, scale, bias));        // correctly rounded on the host too: same bits
#endif
}
MPCR_HD uint32_t filter_word_linear(uint32_t key, float scale, float bias) {
    return filter_word_linear_raw(key, scale, bias) - kLinearExp;
}
MPCR_HD uint32_t filter_bits_linear(uint32_t key) {
    // the paired scanner ROTATES the word left by the shift amount and looks at bit 23 (the float 2^-126)
    return (1u << ((23u - (key & 31u)) & 31u)) | (1u << ((23u - ((key >> 2) & 31u)) & 31u));
}
// scale / bias for n_words filter words; false if the range check fails (never, for n_words < 2^22)
inline bool filter_linear_setup(uint32_t n_words, float* scale, float* bias) {
    if (n_words < 2 || n_words >= (1u << 22)) return false;
    const double s = ((double)n_words - 1.0) / 4194304.0;
    float sf = (float)s;
    if ((double)sf > s) sf = bits_as_float(float_as_bits(sf) - 1u);          // round towards zero: never past the last word
    const float bf = (float)(8388608.0 + 0.25 - 8388608.0 * (double)sf);     // (2^23 + key) * s + bias = 2^23 + key * s + 0.25
    const uint32_t lo = filter_word_linear(0u, sf, bf), hi = filter_word_linear((1u << 22) - 1u, sf, bf);
    if (lo != 0u || hi >= n_words) return false;
    *scale = sf;
    *bias = bf;
    return true;
}

struct Slot {          // 16 bytes, one 128-bit gather; the scanner usually needs only the first 8 of them
    uint32_t code;     // survivor code: record index (one record) or kWalkBucket | first bucket entry; kSlotEmpty = free
    uint32_t tag_a;    // tag of the first record  (one record: tag_b == tag_a; three or more: both tags have mask 0,
    uint32_t tag_b;    // tag of the second record  i.e. they never reject)
    uint32_t key;      // exact seed key (little-endian digits); checked only in hashed mode
};
static_assert(sizeof(Slot) == 16, "Slot layout");
static constexpr uint32_t kSlotEmpty = 0xFFFFFFFFu;   // cudaMemset(0xFF) == all slots empty
static constexpr uint32_t kWalkBucket = 0x80000000u;  // survivor code flag: "walk the bucket starting at code & ~flag"
// Open-addressed tables only: set on every slot that the probe sequence of some key stored FURTHER ON passes through
// (mark_chains, after the table is complete).  A first probe that lands on another key's slot without this flag has
// proved its key absent -- which is what keeps the scanner's false-positive candidates off the synchronous probe path.
// Record indices stay below 2^30 (< 2^29 STS lines), so the bit is free; readers strip it from survivor codes.
static constexpr uint32_t kSlotChain = 0x40000000u;

struct BucketEntry {
    uint32_t rec_last;  // record index | last-of-bucket << 31
    uint32_t tag;
};

// Slot addressing: 4^W slots indexed by the key itself while that stays L2-sized (W <= 11: 64 MiB), otherwise
// open addressing with linear probing at load <= 1/8.
struct SlotMap {
    uint32_t mask;    // n_slots - 1
    uint32_t direct;  // 1: index = key
};

MPCR_HD uint32_t slot_hash(uint32_t key) {
    uint32_t h = key * 0x9E3779B1u;
    return h ^ (h >> 15);
}
MPCR_HD uint32_t slot_index(uint32_t key, SlotMap sm) { return sm.direct ? key : (slot_hash(key) & sm.mask); }

// Sampled tables (mpcr_ctx_set_sampling) are probed once per S positions of the whole genome through a Bloom filter that
// lives in global memory (L2): one 32-bit word per probe, two bits per key.
MPCR_HD uint32_t bloom_word(uint32_t key, uint32_t shift) { return slot_hash(key) >> shift; }
MPCR_HD uint32_t bloom_bits(uint32_t key) {
    const uint32_t h = (key ^ (key >> 13)) * 0x85EBCA77u;
    return (1u << (h >> 27)) | (1u << ((h >> 22) & 31u));
}


// engine.py:614-640 restricted to the tag: true iff the full primer-1 compare is CERTAIN to fail because the
// bases right after the seed already carry more than N mismatches.  tag = 2-bit codes of up to kTagBases
// A/C/G/T primer letters (bits [0,16)) and the mask of the 2-bit lanes that hold one (bits [16,32)); gcodes are
// the 2-bit codes of the genome bases following the seed.  A clean genome base whose code differs from an
// A/C/G/T primer letter mismatches in both compare modes; the caller must not use the verdict when one of the
// kTagBases genome bases is not clean (tag_window_clean) -- those positions go to the full compare.
MPCR_HD bool tag_rejects(uint32_t tag, uint32_t gcodes, int N) {
    uint32_t d = (tag ^ gcodes) & (tag >> 16);
    d = (d | (d >> 1)) & 0x5555u;
#ifdef __CUDA_ARCH__
    return __popc(d) > N;
#else
    return __builtin_popcount(d) > N;
#endif
}
MPCR_HD bool tag_window_clean(uint32_t gvalid) { return (gvalid & 0xFFu) == 0xFFu; }

// the tag of a primer: the kTagBases letters following the seed [ho+W, ho+W+kTagBases), lane i = letter i; lanes that
// hold a plain A/C/G/T letter are masked in, every other lane (degenerate or foreign letter, past the primer's end) is
// left out -- it could match or not, so it must not count as a mismatch
template <class CharAt>
MPCR_HD uint32_t make_tag(CharAt at, int len, int ho, int W) {
    uint32_t codes = 0, mask = 0;
    for (int i = ho + W, n = 0; i < len && n < kTagBases; ++i, ++n) {
        const uint8_t c = at(i);
        uint32_t code;
        if (c == 'A') code = 0; else if (c == 'C') code = 1; else if (c == 'G') code = 2; else if (c == 'T') code = 3;
        else continue;
        codes |= code << (2 * n);
        mask |= 3u << (2 * n);
    }
    return codes | (mask << 16);
}

// n bits of a little-endian bit plane starting at bit index b (n <= 32)
MPCR_HD uint32_t fetch_bits(const uint64_t* plane, int64_t b, int n) {
    const uint64_t i = (uint64_t)b >> 6;
    const unsigned sh = (unsigned)b & 63u;
    uint64_t v = plane[i] >> sh;
    if (sh) v |= plane[i + 1] << (64u - sh);
    return (uint32_t)v & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
}

MPCR_HD Slot load_slot(const Slot* p) {
#ifdef __CUDA_ARCH__
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return Slot{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}

// Does a (non-empty, key-matching) slot leave anything to verify at this position?  False iff every record of the
// seed is ruled out by its inline tag.
MPCR_HD bool slot_survives(const Slot& s, uint32_t gcodes, uint32_t gvalid, int N) {
    if (!tag_window_clean(gvalid)) return true;
    return !(tag_rejects(s.tag_a, gcodes, N) && tag_rejects(s.tag_b, gcodes, N));
}

// Find the slot of a key: returns false when the key is not in the table.  (The scanner issues the first probe
// itself as an async gather and only comes here for hashed-mode collisions.)
MPCR_HD bool find_slot(const Slot* slots, SlotMap sm, uint32_t key, Slot* out) {
    uint32_t i = slot_index(key, sm);
    for (;;) {
        const Slot s = load_slot(slots + i);
        if (s.code == kSlotEmpty) return false;
        if (sm.direct || s.key == key) { *out = s; out->code &= ~kSlotChain; return true; }
        i = (i + 1) & sm.mask;
    }
}

// engine.py:483-484 for one survivor code: visit, in bucket (= insertion) order, every record whose inline tag
// does not already rule out a primer-1 match at this position.
template <class OnRec>
MPCR_HD void for_each_survivor_record(const BucketEntry* bucket, uint32_t code, uint32_t gcodes, uint32_t gvalid, int N,
                                      OnRec&& on_rec) {
    if (!(code & kWalkBucket)) { on_rec(code); return; }
    const bool clean = tag_window_clean(gvalid);
    for (uint32_t e = code & ~kWalkBucket;; ++e) {
        const BucketEntry b = bucket[e];
        if (!clean || !tag_rejects(b.tag, gcodes, N)) on_rec(b.rec_last & 0x7FFFFFFFu);
        if (b.rec_last >> 31) return;
    }
}

// W-mer validity for the 64 hash positions of a strip: bit j set iff bases j .. j+W-1 are all clean.
// v0 = valid bits of the strip's 64 bases, v1 = valid bits of the following 64 bases (only W-1 are used).
MPCR_HD uint64_t window_valid(uint64_t v0, uint64_t v1, int W) {
    // run-length doubling on the 128-bit pair (lo, hi): a_b[j] = AND of 2^b consecutive bits from j
    uint64_t lo[5], hi[5];
    lo[0] = v0; hi[0] = v1;
    for (int b = 1; b < 5; ++b) {
        const int s = 1 << (b - 1);
        uint64_t slo = (lo[b - 1] >> s) | (hi[b - 1] << (64 - s));
        uint64_t shi = hi[b - 1] >> s;
        lo[b] = lo[b - 1] & slo;
        hi[b] = hi[b - 1] & shi;
    }
    uint64_t r = ~0ull;
    int off = 0;
    for (int b = 4; b >= 0; --b) {
        if ((W >> b) & 1) {
            uint64_t x = off == 0 ? lo[b] : ((lo[b] >> off) | (hi[b] << (64 - off)));
            r &= x;
            off += 1 << b;
        }
    }
    return r;
}

// The same for a block-table key: the seed's ws letters from j on and the block's letters from j + ws + gap on.
MPCR_HD uint64_t window_valid_gapped(uint64_t v0, uint64_t v1, int ws, int block, int gap) {
    const int off = ws + gap;   // 1 .. 15
    const uint64_t s0 = (v0 >> off) | (v1 << (64 - off)), s1 = v1 >> off;
    return window_valid(v0, v1, ws) & window_valid(s0, s1, block);
}

}  // namespace mpcr
