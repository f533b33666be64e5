// host_emul.cpp -- TEST INFRASTRUCTURE: a serial CPU emulation of the libmerpcr_b200.so C ABI.
//
// It implements include/merpcr_b200.h with "device" pointers being plain host pointers, and runs the SAME
// per-position semantics as the CUDA kernels because it includes merpcr_b200/csrc/mpcr_core.cuh (all of whose
// functions are __host__ __device__).  Purpose: let the CPU-only test tier exercise the Python host logic
// (parsing, layout, sharding, formatting) and the shared verification code against the oracle before any GPU
// time is spent.  It is injected explicitly by tests (merpcr_b200._capi._inject_backend_for_tests); the product
// never loads it and has no CPU fallback.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>
#include <vector>

#include "../../include/merpcr_b200.h"
#include "../../merpcr_b200/csrc/mpcr_core.cuh"
#include "../../merpcr_b200/csrc/mpcr_hostio.h"
#include "../../merpcr_b200/csrc/mpcr_hostpack.cpp"   // the product's own host packer (host code): mpcr_host_pack_nibbles

using namespace mpcr;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

struct mpcr_ctx {
    mpcr_params prm{};
    uint32_t n_rec = 0, n_valid = 0;
    std::vector<RecMeta> meta;
    std::vector<uint64_t> pwords;
    std::vector<Slot> slots;
    SlotMap smap{1023u, 0u};
    std::vector<BucketEntry> bucket;
    std::vector<uint32_t> filter;
    bool filter_linear = false;   // same rule as the CUDA library: contiguous 11-letter keys in a direct table
    float filter_scale = 0.f, filter_bias = 0.f;
    uint32_t max_hash_off = 0, max_len = 0;
    uint64_t max_pcr = 0;
    bool table_ready = false;
    uint64_t launches = 0;
    int ext_w = 0, ext_which = 0, scan_w = 0, true_strands = 0;
    int ext_block = 0, ext_gap = 0, ext_span = 0;
    int samp_w = 0, samp_s = 0, samp_role = 0;
    uint32_t part = 0, parts = 1;
    int append = 0;
};

extern "C" {

int mpcr_abi_version(void) { return MPCR_ABI_VERSION; }
const char* mpcr_last_error(void) { return g_err; }

int mpcr_ctx_create(int device, const mpcr_params* p, mpcr_ctx** out) {
    (void)device;
    if (!p || !out) return fail(MPCR_EINVAL, "null argument");
    if (p->wordsize < 3 || p->wordsize > 16) return fail(MPCR_EINVAL, "Word size must be between 3 and 16");
    if (p->mismatches < 0 || p->mismatches > 10) return fail(MPCR_EINVAL, "Number of mismatches must be between 0 and 10");
    if (p->margin < 0 || p->margin > 10000) return fail(MPCR_EINVAL, "Margin must be between 0 and 10000");
    if (p->three_prime_match < 0) return fail(MPCR_EINVAL, "Three prime match must be at least 0");
    mpcr_ctx* c = new mpcr_ctx();
    c->prm = *p;
    *out = c;
    return MPCR_OK;
}
void mpcr_ctx_destroy(mpcr_ctx* c) { delete c; }
int mpcr_ctx_set_true_strands(mpcr_ctx* c, int on) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    c->true_strands = on ? 1 : 0; c->table_ready = false;
    return MPCR_OK;
}
int mpcr_ctx_set_seed_extension(mpcr_ctx* c, int w_ext, int which) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (which < 0 || which > 2) return fail(MPCR_EINVAL, "which must be 0, 1 or 2");
    if (which != 0) {
        if (c->prm.mismatches != 0 || c->prm.iupac_mode != 0)
            return fail(MPCR_EINVAL, "seed extension needs an exact search (mismatches 0, no IUPAC mode)");
        if (w_ext <= c->prm.wordsize || w_ext > 16) return fail(MPCR_EINVAL, "extended word must be in (wordsize, 16]");
    }
    c->ext_w = which ? w_ext : 0; c->ext_which = which; c->table_ready = false;
    c->ext_block = c->ext_gap = c->ext_span = 0;
    return MPCR_OK;
}
int mpcr_ctx_set_seed_blocks(mpcr_ctx* c, int block, int n_blocks, int which) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (which < 0 || (which >= 2 && which - 2 >= n_blocks)) return fail(MPCR_EINVAL, "which must be 0, 1 or 2 + block index");
    if (which != 0) {
        if (c->prm.iupac_mode != 0) return fail(MPCR_EINVAL, "block tables need letter-identity compares (no IUPAC mode)");
        if (block < 1 || n_blocks < 1) return fail(MPCR_EINVAL, "block and n_blocks must be positive");
        if (n_blocks <= c->prm.mismatches) return fail(MPCR_EINVAL, "block tables need more blocks than mismatches");
        if (c->prm.wordsize + n_blocks * block > 16) return fail(MPCR_EINVAL, "wordsize + n_blocks * block must not exceed 16 letters");
    }
    c->ext_w = which ? c->prm.wordsize + block : 0;
    c->ext_which = which >= 2 ? 2 : which;
    c->ext_block = which ? block : 0;
    c->ext_gap = which >= 2 ? (which - 2) * block : 0;
    c->ext_span = which ? c->prm.wordsize + n_blocks * block : 0;
    c->table_ready = false;
    return MPCR_OK;
}
int mpcr_ctx_set_sampling(mpcr_ctx* c, int w_samp, int stride, int role) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (role < 0 || role > 2) return fail(MPCR_EINVAL, "role must be 0, 1 or 2");
    if (role != 0) {
        if (c->prm.mismatches != 0 || c->prm.iupac_mode != 0)
            return fail(MPCR_EINVAL, "position sampling needs an exact search (mismatches 0, no IUPAC mode)");
        if (w_samp <= c->prm.wordsize || w_samp > 16) return fail(MPCR_EINVAL, "sampled word must be in (wordsize, 16]");
        if (stride < 2 || stride > 64) return fail(MPCR_EINVAL, "stride must be in [2, 64]");
    }
    c->samp_w = role ? w_samp : 0; c->samp_s = role ? stride : 0; c->samp_role = role; c->table_ready = false;
    return MPCR_OK;
}
uint32_t mpcr_table_items(const mpcr_ctx* c) { return c && c->table_ready ? c->n_valid : 0; }
int mpcr_ctx_set_table_part(mpcr_ctx* c, uint32_t part, uint32_t parts) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (parts == 0) parts = 1;
    if (part >= parts) return fail(MPCR_EINVAL, "part must be < parts");
    c->part = part; c->parts = parts; c->table_ready = false;
    return MPCR_OK;
}
int mpcr_ctx_set_append(mpcr_ctx* c, int on) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    c->append = on ? 1 : 0;
    return MPCR_OK;
}
int mpcr_ctx_sm_count(const mpcr_ctx*) { return 1; }
uint64_t mpcr_launch_count(const mpcr_ctx* c) { return c ? c->launches : 0; }
float mpcr_last_scan_ms(mpcr_ctx*) { return 0.f; }
float mpcr_last_verify_ms(mpcr_ctx*) { return 0.f; }

int mpcr_pack_sequence(mpcr_ctx* c, const uint8_t* ascii, uint64_t n, uint64_t dst_base, uint64_t origin, void* plane2,
                       void* plane4, void* valid, const uint8_t* lut, void*) {
    if (!c || !plane2 || !plane4 || !valid || !lut) return fail(MPCR_EINVAL, "null argument");
    if ((dst_base & 63u) || (origin & 127u) || dst_base < origin) return fail(MPCR_EINVAL, "bad alignment");
    uint64_t *P2 = (uint64_t*)plane2, *P4 = (uint64_t*)plane4, *V = (uint64_t*)valid;
    const uint64_t rel = dst_base - origin;
    for (uint64_t s = 0; s * 64 < n; ++s) {
        uint64_t v = 0, p2[2] = {0, 0}, p4[4] = {0, 0, 0, 0};
        for (int j = 0; j < 64; ++j) {
            uint64_t i = s * 64 + j;
            uint32_t e = lut[i < n ? ascii[i] : 0];
            p4[j >> 4] |= (uint64_t)(e & 15u) << (4 * (j & 15));
            p2[j >> 5] |= (uint64_t)((e >> 4) & 3u) << (2 * (j & 31));
            v |= (uint64_t)((e >> 6) & 1u) << j;
        }
        const uint64_t w = rel / 64 + s;
        V[w] = v; P2[2 * w] = p2[0]; P2[2 * w + 1] = p2[1];
        for (int k = 0; k < 4; ++k) P4[4 * w + k] = p4[k];
    }
    c->launches++;
    return MPCR_OK;
}

int mpcr_derive_planes(mpcr_ctx* c, uint64_t n, uint64_t dst_base, uint64_t origin, const void* plane4, void* plane2,
                       void* valid, void*) {
    if (!c || !plane2 || !plane4 || !valid) return fail(MPCR_EINVAL, "null argument");
    if ((dst_base & 63u) || (origin & 127u) || dst_base < origin) return fail(MPCR_EINVAL, "bad alignment");
    const uint64_t* P4 = (const uint64_t*)plane4;
    uint64_t *P2 = (uint64_t*)plane2, *V = (uint64_t*)valid;
    const uint64_t rel = dst_base - origin;
    for (uint64_t s = 0; s * 64 < n; ++s) {
        const uint64_t w = rel / 64 + s;
        uint64_t v = 0, p2[2] = {0, 0};
        for (int j = 0; j < 64; ++j) {
            const uint32_t nib = (uint32_t)(P4[4 * w + (j >> 4)] >> (4 * (j & 15))) & 15u;
            const uint32_t code = nib == 2 ? 1u : nib == 4 ? 2u : nib == 8 ? 3u : 0u;
            if (nib == 1 || nib == 2 || nib == 4 || nib == 8) v |= 1ull << j;
            p2[j >> 5] |= (uint64_t)code << (2 * (j & 31));
        }
        V[w] = v; P2[2 * w] = p2[0]; P2[2 * w + 1] = p2[1];
    }
    c->launches++;
    return MPCR_OK;
}

int mpcr_sts_parse(const uint8_t* text, uint64_t n, int32_t wordsize, int32_t default_pcr_size, mpcr_sts_line* lines,
                   uint32_t max_lines, uint32_t* n_lines, uint32_t* bad_line, uint32_t* short_primers, uint32_t* flags) {
    return sts_parse_impl(text, n, wordsize, default_pcr_size, lines, max_lines, n_lines, bad_line, short_primers, flags);
}
int mpcr_sts_blob(const uint8_t* text, const mpcr_sts_line* lines, uint32_t n_lines, uint8_t* blob, uint64_t* off) {
    sts_blob_impl(text, lines, n_lines, blob, off);
    return MPCR_OK;
}
uint64_t mpcr_format_hits(const mpcr_hit* hits, uint64_t n, const uint8_t* text, const mpcr_sts_line* lines,
                          const uint8_t* labels, const uint64_t* label_off, uint8_t* out, uint64_t out_cap) {
    return format_hits_impl(hits, n, text, lines, labels, label_off, out, out_cap);
}

// ---- FASTA text ingest, serial restatement of io/fasta.py:43-66 for ASCII bytes ------------------------------------
static bool f_term(uint8_t c) { return c == 10 || c == 13; }
static bool f_blank(uint8_t c) { return c == 32 || c == 9 || c == 11 || c == 12 || (c >= 28 && c <= 31); }
static bool f_keep(uint8_t c) {
    static const char* K = "ACGTBDHKMNRSVWXYacgtbdhkmnrsvwxy";
    return c && strchr(K, c) != nullptr;
}
uint64_t mpcr_fasta_workspace_bytes(uint64_t n, uint32_t max_records) { return 64 + n / 512 + (uint64_t)max_records; }
int mpcr_fasta_index_ex(mpcr_ctx* c, uint8_t* text, uint64_t n, uint32_t mode, mpcr_fasta_record* recs, uint32_t max_records,
                        uint32_t* n_records, uint32_t* flags, void* ws, uint64_t ws_bytes, void*) {
    if (!c || !n_records || !flags) return fail(MPCR_EINVAL, "null argument");
    *n_records = 0; *flags = 0;
    if (n == 0) return MPCR_OK;
    if (!ws || ws_bytes < mpcr_fasta_workspace_bytes(n, max_records)) return fail(MPCR_EINVAL, "workspace too small");
    const bool skip_first_line = mode & 1u, keep_prologue = mode & 2u;
    uint64_t text_begin = 0;
    if (skip_first_line) {
        uint64_t t = 0;
        const uint64_t lim = n < (1u << 20) ? n : (1u << 20);
        while (t < lim && !f_term(text[t])) ++t;
        if (t >= lim) { *flags = 2u; return MPCR_OK; }
        text_begin = (text[t] == 13 && t + 1 < n && text[t + 1] == 10) ? t + 2 : t + 1;
        memset(text, '\n', text_begin);
    }
    for (uint64_t i = 0; i < n; ++i) if (text[i] >= 128) { *flags = 1; return MPCR_OK; }
    std::vector<std::pair<uint64_t, uint64_t>> hdr;
    uint64_t line = 0;
    while (line < n) {
        uint64_t e = line;
        while (e < n && !f_term(text[e])) ++e;
        uint64_t a = line;
        while (a < e && f_blank(text[a])) ++a;
        if (a < e && text[a] == '>') hdr.push_back({a, e});
        line = e + 1;   // "\r\n" yields an empty line in between, which changes nothing
    }
    c->launches++;
    const uint32_t lead = keep_prologue ? 1u : 0u;
    *n_records = (uint32_t)hdr.size() + lead;
    if (hdr.size() + lead > max_records) return MPCR_EOVERFLOW;
    if (hdr.empty() && !keep_prologue) { *n_records = 0; return MPCR_OK; }
    if (!hdr.empty() && !keep_prologue) memset(text, '\n', hdr[0].first);
    for (auto& h : hdr) memset(text + h.first, '\n', h.second - h.first);
    uint64_t kept = 0, pos = 0;
    if (lead) {
        const uint64_t stop = hdr.empty() ? n : hdr[0].first;
        recs[0].header_begin = recs[0].header_end = text_begin; recs[0].seq_offset = 0;
        for (; pos < stop; ++pos) kept += f_keep(text[pos]);
        recs[0].seq_length = kept;
    }
    for (size_t r = 0; r < hdr.size(); ++r) {
        const uint64_t stop = r + 1 < hdr.size() ? hdr[r + 1].first : n;
        for (; pos < hdr[r].second; ++pos) kept += f_keep(text[pos]);
        recs[r + lead].header_begin = hdr[r].first; recs[r + lead].header_end = hdr[r].second; recs[r + lead].seq_offset = kept;
        for (; pos < stop; ++pos) kept += f_keep(text[pos]);
        recs[r + lead].seq_length = kept - recs[r + lead].seq_offset;
    }
    return MPCR_OK;
}
int mpcr_fasta_index(mpcr_ctx* c, uint8_t* text, uint64_t n, mpcr_fasta_record* recs, uint32_t max_records,
                     uint32_t* n_records, uint32_t* flags, void* ws, uint64_t ws_bytes, void* st) {
    return mpcr_fasta_index_ex(c, text, n, 0u, recs, max_records, n_records, flags, ws, ws_bytes, st);
}
int mpcr_fasta_offsets_at(mpcr_ctx* c, const uint8_t* text, uint64_t n, const void*, const uint64_t* pos, uint32_t n_pos,
                          uint64_t* out, void*) {
    if (!c || (n_pos && (!pos || !out || !text))) return fail(MPCR_EINVAL, "null argument");
    for (uint32_t i = 0; i < n_pos; ++i) {
        const uint64_t p = pos[i] < n ? pos[i] : n;
        uint64_t k = 0;
        for (uint64_t j = 0; j < p; ++j) k += f_keep(text[j]);
        out[i] = k;
    }
    return MPCR_OK;
}
int mpcr_fasta_compact(mpcr_ctx* c, const uint8_t* text, uint64_t n, const void*, uint8_t* seq, void*) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) if (f_keep(text[i])) seq[k++] = text[i];
    c->launches++;
    return MPCR_OK;
}

struct Fwd { const uint8_t* p; uint8_t operator()(int i) const { return p[i]; } };
struct Rc { const uint8_t* p; int len; uint8_t operator()(int i) const { return complement_of(p[len - 1 - i]); } };

int mpcr_table_build(mpcr_ctx* c, const uint8_t* blob, const uint64_t* off, const uint32_t* pcr, uint32_t n_lines,
                     const uint8_t* plut, void*) {
    if (!c || !plut) return fail(MPCR_EINVAL, "null argument");
    const int W = c->prm.wordsize;
    const bool sampled = c->samp_role == 1;
    const int which = sampled ? 0 : c->ext_which;
    const int WS = sampled ? c->samp_w : (c->ext_which == 2 ? c->ext_w : W);
    const uint32_t per = sampled ? (uint32_t)c->samp_s : 1u;
    std::vector<uint32_t> tags((size_t)2 * n_lines * per, 0u);
    c->scan_w = WS;
    c->n_rec = 2 * n_lines; c->n_valid = 0; c->max_hash_off = 0; c->max_len = 0; c->max_pcr = 0;
    c->meta.assign(c->n_rec, RecMeta{});
    c->pwords.clear();
    std::vector<std::pair<uint32_t, uint32_t>> pairs;  // (key, rec) of inserted records
    for (uint32_t l = 0; l < n_lines; ++l) {
        const uint8_t *pr1 = blob + off[2 * l], *pr2 = blob + off[2 * l + 1];
        const int n1 = (int)(off[2 * l + 1] - off[2 * l]), n2 = (int)(off[2 * l + 2] - off[2 * l + 1]);
        if (n1 > 65535 || n2 > 65535) return fail(MPCR_EINVAL, "primer longer than 65535 bases at STS entry %u", l);
        c->max_len = std::max(c->max_len, (uint32_t)std::max(n1, n2));
        c->max_pcr = std::max<uint64_t>(c->max_pcr, pcr[l]);
        for (int minus = 0; minus < 2; ++minus) {
            const uint32_t r = 2 * l + minus;
            RecMeta& m = c->meta[r];
            m.pcr_size = pcr[l];
            const int l1 = minus ? n2 : n1, l2 = minus ? n1 : n2;
            m.len1 = (uint16_t)l1; m.len2 = (uint16_t)l2;
            m.p1_word = (uint32_t)c->pwords.size();
            c->pwords.resize(c->pwords.size() + 2 * ((l1 + 15) / 16));
            m.p2_word = (uint32_t)c->pwords.size();
            c->pwords.resize(c->pwords.size() + 2 * ((l2 + 15) / 16));
            uint32_t hbe = 0, kext = 0; int ho; bool ext;
            const Fwd q1{minus ? pr2 : pr1};
            ho = first_clean_word(q1, l1, W, &hbe);
            ext = which != 0 && (c->ext_block ? blocked_seed(q1, l1, ho, W, c->ext_block, c->ext_gap, c->ext_span, &kext)
                                              : extended_seed(q1, l1, ho, c->ext_w, &kext));
            if (ho >= 0) m.tag = make_tag(q1, l1, ho, which == 2 ? c->ext_w + c->ext_gap : W);
            encode_primer(q1, l1, plut, c->pwords.data() + m.p1_word);
            if (!minus) {
                if (c->true_strands) encode_primer(Rc{pr2, n2}, n2, plut, c->pwords.data() + m.p2_word);
                else encode_primer(Fwd{pr2}, n2, plut, c->pwords.data() + m.p2_word);
            } else {
                encode_primer(Rc{pr1, n1}, n1, plut, c->pwords.data() + m.p2_word);
            }
            const bool in_part = c->parts <= 1u || (r >> 1) % c->parts == c->part;
            bool here = ho >= 0 && (which == 0 || (which == 1 ? !ext : ext)) && in_part;
            m.hash_be = hbe; m.key = which == 2 ? kext : reverse_digits(hbe, W);
            m.hash_off = (uint16_t)(ho < 0 ? 0 : ho);
            if (ho >= 0) c->max_hash_off = std::max(c->max_hash_off, (uint32_t)ho);
            if (c->samp_role != 0) {
                const bool can = sampleable_seed(q1, l1, ho, c->samp_w, c->samp_s);
                if (sampled) {
                    here = false;
                    for (int d = 0; d < c->samp_s && can && in_part; ++d) {
                        uint32_t key = 0;
                        extended_seed(q1, l1, ho + d, c->samp_w, &key);
                        tags[(size_t)r * per + d] = make_tag(q1, l1, ho + d, c->samp_w);
                        pairs.push_back({key, r * per + (uint32_t)d});
                    }
                } else {
                    here = here && !can;
                }
            }
            m.flags = (ho >= 0 ? 1 : 0) | (here ? 2 : 0);
            if (!sampled) tags[r] = m.tag;
            if (here) pairs.push_back({m.key, r});
        }
    }
    c->pwords.resize(c->pwords.size() + 2);
    std::stable_sort(pairs.begin(), pairs.end(), [](auto& a, auto& b) { return a.first < b.first; });
    c->n_valid = (uint32_t)pairs.size();
    uint32_t words = 39616;  // what a B200 scanner CTA has room for; any multiple of 4 works
    if (const char* env = getenv("MPCR_FILTER_WORDS")) { long v = atol(env); if (v >= 4) words = (uint32_t)v & ~3u; }
    c->filter.assign(words, 0);
    const uint32_t cw = filter_mul(WS);
    uint32_t nslots = 1024;
    const bool direct = getenv("MPCR_EMUL_HASHED") ? false : WS <= 11;  // the env switch lets CPU tests cover both modes
    if (direct) nslots = 1u << (2 * WS);
    else while (nslots < 2u * c->n_valid + 2u) nslots <<= 1;
    c->smap = SlotMap{nslots - 1, direct ? 1u : 0u};
    c->filter_linear = MPCR_LINEAR_FILTER && WS == kLinearW && direct && !sampled && !(c->ext_which == 2 && c->ext_gap > 0) &&
                       filter_linear_setup(words, &c->filter_scale, &c->filter_bias);
    c->slots.assign(nslots, Slot{~0u, ~0u, ~0u, ~0u});
    c->bucket.assign(c->n_valid + 1, BucketEntry{0, 0});
    for (uint32_t i = 0; i < c->n_valid; ++i) {
        const uint32_t key = pairs[i].first, rec = pairs[i].second;
        const bool head = i == 0 || pairs[i - 1].first != key, last = i + 1 == c->n_valid || pairs[i + 1].first != key;
        c->bucket[i] = BucketEntry{rec | (last ? 0x80000000u : 0u), tags[rec]};
        if (head) {
            uint32_t n = 1, tag_b = 0;
            if (!last) {
                n = 2;
                tag_b = tags[pairs[i + 1].second];
                if (i + 2 < c->n_valid && pairs[i + 2].first == key) n = 3;
            }
            uint32_t s = slot_index(key, c->smap);
            if (!direct) while (c->slots[s].code != kSlotEmpty) s = (s + 1) & c->smap.mask;
            const uint32_t tag_a = tags[rec];
            c->slots[s] = n == 1 ? Slot{rec, tag_a, tag_a, key}
                          : n == 2 ? Slot{kWalkBucket | i, tag_a, tag_b, key} : Slot{kWalkBucket | i, 0u, 0u, key};
            if (c->filter_linear) c->filter[filter_word_linear(key, c->filter_scale, c->filter_bias)] |= filter_bits_linear(key);
            else c->filter[filter_word(key, cw, words)] |= filter_bits_of(key, WS);
        }
    }
    c->table_ready = true;
    c->launches++;
    return MPCR_OK;
}

int mpcr_table_records(mpcr_ctx* c, int32_t* ho, uint32_t* h) {
    if (!c || !c->table_ready) return fail(MPCR_ESTATE, "table not built");
    for (uint32_t r = 0; r < c->n_rec; ++r) {
        if (ho) ho[r] = (c->meta[r].flags & 1) ? (int32_t)c->meta[r].hash_off : -1;
        if (h) h[r] = c->meta[r].hash_be;
    }
    return MPCR_OK;
}

int mpcr_table_primer_words(mpcr_ctx* c, uint32_t rec, int which, uint64_t* out, uint32_t max_words, uint32_t* n_words) {
    if (!c || !c->table_ready) return fail(MPCR_ESTATE, "table not built");
    if (rec >= c->n_rec || (which != 1 && which != 2)) return fail(MPCR_EINVAL, "bad record / primer index");
    const RecMeta& m = c->meta[rec];
    const uint32_t len = which == 1 ? m.len1 : m.len2, nw = 2 * ((len + 15) / 16);
    if (n_words) *n_words = nw;
    if (nw > max_words) return fail(MPCR_EINVAL, "buffer too small");
    memcpy(out, c->pwords.data() + (which == 1 ? m.p1_word : m.p2_word), nw * 8);
    return MPCR_OK;
}

static uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
uint64_t mpcr_tile_bases(void) { return (uint64_t)kTileBases; }
uint64_t mpcr_halo_left(const mpcr_ctx* c) { return c ? round_up((uint64_t)c->max_hash_off + 64, 128) : 0; }
uint64_t mpcr_halo_right(const mpcr_ctx* c) {
    return c ? round_up(c->max_pcr + (uint64_t)c->prm.margin + c->max_len + 64 + 128, 128) + kTileBases : 0;
}

int mpcr_scan_prepare(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, uint64_t origin, uint64_t sb, uint64_t,
                      void*) {
    if (!c || (n_contigs && !contigs)) return fail(MPCR_EINVAL, "null argument");
    if ((origin & 127u) || (sb & 127u)) return fail(MPCR_EINVAL, "origin / shard_begin must be multiples of 128");
    return MPCR_OK;   // the emulation walks the contig table directly
}

int mpcr_scan(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, const void* plane2, const void* plane4,
              const void* valid, uint64_t origin, uint64_t plane_bases, uint64_t sb, uint64_t se, mpcr_hit* hits,
              uint64_t capacity, uint64_t* count, void*) {
    if (!c || !count) return fail(MPCR_EINVAL, "null argument");
    if (!c->table_ready) return fail(MPCR_ESTATE, "mpcr_scan called before mpcr_table_build");
    if ((origin & 127u) || (sb & 127u)) return fail(MPCR_EINVAL, "origin / shard_begin must be multiples of 128");
    const uint64_t *P2 = (const uint64_t*)plane2, *P4 = (const uint64_t*)plane4, *V = (const uint64_t*)valid;
    SearchParams prm{c->scan_w, c->prm.margin, c->prm.mismatches, c->prm.three_prime_match, c->prm.iupac_mode ? 1 : 0};
    const bool gapped = c->ext_which == 2 && c->ext_block > 0 && c->samp_role != 1;
    prm.gap = gapped ? c->ext_gap : 0;
    prm.block = gapped ? c->ext_block : 0;
    const uint32_t seed_mask = wmask_of(prm.W - prm.block);
    const int W_ref = c->prm.wordsize;   // the contig-length rule of engine.py:458 uses the reference's word size
    const uint32_t wmask = wmask_of(prm.W), cw = filter_mul(prm.W);
    uint64_t n = c->append ? *count : 0;   // append mode: keep counting behind the previous calls' hits
    c->launches++;
    for (uint32_t ci = 0; ci < n_contigs && c->n_valid; ++ci) {
        const uint64_t L = contigs[ci].length, g0 = contigs[ci].gstart;
        if (L <= (uint64_t)W_ref) continue;
        if (g0 & 127u) return fail(MPCR_EINVAL, "contig %u: gstart not a multiple of 128", ci);
        // ownership per 2048-position unit, by the unit's first base (same rule as the CUDA library's build_tiles)
        if (se <= g0 || sb >= g0 + L) continue;
        const uint64_t lo = sb > g0 ? round_up(sb - g0, 2048) : 0, stop = std::min<uint64_t>(se - g0, L);
        if (lo >= stop) continue;
        const uint64_t hi = std::min<uint64_t>(round_up(stop, 2048), L);
        for (uint64_t ls = lo; ls < hi; ls += kTileBases) {
            const uint64_t g = g0 + ls;
            const int64_t gbase = (int64_t)(g - origin);
            const uint32_t nb = (uint32_t)std::min<uint64_t>(hi - ls, kTileBases);
            for (uint32_t lp0 = 0; lp0 < nb; lp0 += 64) {
                const int64_t gb = gbase + lp0;
                if ((uint64_t)gb + 128 > plane_bases + 128) return fail(MPCR_EINVAL, "planes too small for the shard");
                uint64_t cand = prm.gap > 0 ? window_valid_gapped(V[gb >> 6], V[(gb >> 6) + 1], prm.W - prm.block, prm.block, prm.gap)
                                            : window_valid(V[gb >> 6], V[(gb >> 6) + 1], prm.W);
                for (int j = 0; j < 64 && cand; ++j) {
                    if (!((cand >> j) & 1)) continue;
                    if (c->samp_role == 1 && (ls + lp0 + j) % (uint64_t)c->samp_s != 0) continue;   // probed positions only
                    const uint32_t key = prm.gap > 0 ? (gap_key_raw(extract_key(P2, gb + j, 0xFFFFFFFFu), seed_mask, prm.gap) & wmask)
                                                     : extract_key(P2, gb + j, wmask);
                    if (c->filter_linear) {
                        const uint32_t fw = c->filter[filter_word_linear(key, c->filter_scale, c->filter_bias)], fm = filter_bits_linear(key);
                        if ((fw & fm) != fm) continue;
                    } else if (!filter_pass(c->filter[filter_word(key, cw, (uint32_t)c->filter.size())], key, prm.W)) continue;
                    const uint32_t gcodes = fetch_bits(P2, 2 * (gb + j + prm.W + prm.gap), 2 * kTagBases);
                    const uint32_t gvalid = fetch_bits(V, gb + j + prm.W + prm.gap, kTagBases);
                    Slot sl;
                    if (!find_slot(c->slots.data(), c->smap, key, &sl)) continue;
                    if (!slot_survives(sl, gcodes, gvalid, prm.N)) continue;
                    for_each_survivor_record(c->bucket.data(), sl.code, gcodes, gvalid, prm.N, [&](uint32_t item) {
                        const uint32_t S = c->samp_role == 1 ? (uint32_t)c->samp_s : 1u, rec = item / S, win = item % S;
                        const RecMeta& m = c->meta[rec];
                        verify_record(P4, gbase - (int64_t)ls, (int64_t)L, (int64_t)ls + lp0 + j - (int64_t)win, m, c->pwords.data(), prm,
                                      [&](int64_t p1, int64_t p2, uint32_t rank) {
                                          if (n < capacity) hits[n] = mpcr_hit{ci, (uint32_t)p1, (uint32_t)p2, rec, rank, m.hash_off};
                                          ++n;
                                      });
                    });
                }
            }
        }
    }
    *count = n;
    return MPCR_OK;
}

int mpcr_sort_hits(mpcr_ctx* c, mpcr_hit* hits, uint64_t n, void*) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    std::sort(hits, hits + n, [](const mpcr_hit& a, const mpcr_hit& b) {
        return std::make_tuple(a.contig, a.pos1, a.hash_off, a.rec, a.rank) < std::make_tuple(b.contig, b.pos1, b.hash_off, b.rec, b.rank);
    });
    c->launches++;
    return MPCR_OK;
}

int mpcr_sort_hits_dev(mpcr_ctx* c, mpcr_hit* hits, const uint64_t* count, uint64_t capacity, uint64_t, void* st) {
    if (!c || !count) return fail(MPCR_EINVAL, "null argument");
    return mpcr_sort_hits(c, hits, *count < capacity ? *count : capacity, st);
}

int mpcr_scan_sorted(mpcr_ctx* const* ctxs, uint32_t n_ctx, const mpcr_contig* contigs, uint32_t n_contigs, const void* p2,
                     const void* p4, const void* valid, uint64_t origin, uint64_t plane_bases, uint64_t sb, uint64_t se,
                     mpcr_hit* hits, uint64_t capacity, uint64_t* count, uint64_t* h_count, uint64_t, int sort, void* st) {
    if (!ctxs || n_ctx == 0 || !count) return fail(MPCR_EINVAL, "null argument");
    for (uint32_t i = 0; i < n_ctx; ++i) {
        const int saved = ctxs[i]->append;
        ctxs[i]->append = i == 0 ? 0 : 1;
        const int rc = mpcr_scan(ctxs[i], contigs, n_contigs, p2, p4, valid, origin, plane_bases, sb, se, hits, capacity, count, st);
        ctxs[i]->append = saved;
        if (rc) return rc;
    }
    if (sort) mpcr_sort_hits(ctxs[0], hits, *count < capacity ? *count : capacity, st);
    if (h_count) *h_count = *count;
    return MPCR_OK;
}

int mpcr_scan_sorted_async(mpcr_ctx* const* ctxs, uint32_t n_ctx, const mpcr_contig* contigs, uint32_t n_contigs, const void* p2,
                           const void* p4, const void* valid, uint64_t origin, uint64_t plane_bases, uint64_t sb, uint64_t se,
                           mpcr_hit* hits, uint64_t capacity, uint64_t* count, uint64_t* h_result, uint64_t hint, int sort,
                           int slot, void* st) {
    if (slot < 0 || slot >= MPCR_MAX_SLOTS) return fail(MPCR_EINVAL, "slot must be in [0, %d)", MPCR_MAX_SLOTS);
    uint64_t n = 0;
    const int rc = mpcr_scan_sorted(ctxs, n_ctx, contigs, n_contigs, p2, p4, valid, origin, plane_bases, sb, se, hits, capacity,
                                    count, &n, hint, sort, st);
    if (h_result) { h_result[0] = n; h_result[1] = 0; }
    return rc;
}
int mpcr_sort_finish(mpcr_ctx* c, mpcr_hit*, const uint64_t* h_result, uint64_t, void*) {
    if (!c || !h_result) return fail(MPCR_EINVAL, "null argument");
    return MPCR_OK;
}
float mpcr_slot_scan_ms(mpcr_ctx*, int) { return 0.f; }
float mpcr_slot_verify_ms(mpcr_ctx*, int) { return 0.f; }

// Test-only probe of the rolling mate pre-check (mpcr_core.cuh): the block form against compare_view's own first check,
// position by position.  Returns the number of positions where they differ.
uint32_t emul_mate_precheck_diff(const uint64_t* p4, int64_t gb, uint32_t m, const uint64_t* pw, int len, int N, int X, int iupac) {
    SearchParams prm{11, 50, N, X, iupac};
    const PrimerView v = make_primer_view(pw, len, false, prm);
    if (!v.nw) return 0;
    const uint32_t block = mate_precheck32(p4, gb, m, v, prm);
    uint32_t diff = 0;
    for (uint32_t t = 0; t < m; ++t) {
        const uint32_t m8 = mismatch_lanes32(fetch8(p4, gb + t), (uint32_t)v.q[0], (uint32_t)v.aux[0], (uint32_t)v.lanes[0], prm.iupac);
        const bool pass = !((m8 & (uint32_t)v.prot[0]) || __builtin_popcount(m8) > prm.N);
        if (pass != (((block >> t) & 1u) != 0)) ++diff;
        if (!pass && compare_view(p4, gb + t, v, prm)) ++diff;   // the pre-check is a necessary condition
    }
    if (m < 32 && (block >> m)) ++diff;
    return diff;
}
// Test-only probe of the four-bytes-at-a-time FASTA keep test (mpcr_core.cuh)
uint32_t emul_fasta_keep_flags4(uint32_t w) { return fasta_keep_flags4(w); }
uint32_t emul_bytes_equal_trigger4(uint32_t w, uint32_t c4) { return bytes_equal_trigger4(w, c4); }
// Test-only probes of the linear filter map (mpcr_core.cuh): word index and bit mask of a key for n_words filter words.
int emul_filter_linear(uint32_t n_words, uint32_t key, uint32_t* word, uint32_t* mask) {
    float scale = 0.f, bias = 0.f;
    if (!filter_linear_setup(n_words, &scale, &bias)) return 0;
    *word = filter_word_linear(key, scale, bias);
    *mask = filter_bits_linear(key);
    return 1;
}

}  // extern "C"
