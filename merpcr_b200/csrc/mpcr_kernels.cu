// mpcr_kernels.cu -- sm_100a kernels + the C ABI of libmerpcr_b200.so (include/merpcr_b200.h).
//
//   (1) pack_kernel        : filtered ASCII -> plane2 / plane4 / valid        (io/fasta.py:58-61, engine.py:455-478)
//   (2) encode_records ... : STS lines -> both-strand records, hashes, CSR bucket table + first-level filter
//                                                                              (engine.py:253-281,324-359)
//   (3)(4)(5) scan_kernel  : rolling W-mer keys, shared-memory filter probe, bucket walk, primer verify,
//                            mate search, hit append                           (engine.py:453-642)
//   sort                   : mpcr_sort.cuh                                     (engine.py:434 + tie order)
//
// Reference citations are relative to /root/reference/src/merpcr/.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/merpcr_b200.h"
#include "mpcr_core.cuh"
#include "mpcr_sort.cuh"
#include "mpcr_fasta.cuh"
#include "mpcr_hostio.h"

using namespace mpcr;

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(MPCR_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                           \
    } while (0)

// Entry points run on the context's device and leave the caller's current device as they found it (the Python
// host shares the process with PyTorch, whose notion of the current device must not change under it).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
        else prev = -1;   // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define GUARD(c)                                                                                          \
    DeviceGuard guard_((c)->device);                                                                      \
    if (!guard_.ok) return fail(MPCR_ECUDA, "cudaSetDevice(%d) failed", (c)->device)

struct Survivor;
struct LongRun;
struct mpcr_ctx {
    int device = 0;
    mpcr_params prm{};
    int sm_count = 0;
    int max_smem_optin = 0;
    // table
    uint32_t n_rec = 0, n_valid = 0;
    RecMeta* d_meta = nullptr;
    uint64_t* d_pwords = nullptr;
    uint64_t total_pwords = 0;
    Slot* d_slots = nullptr;
    SlotMap smap{1023u, 0u};
    BucketEntry* d_bucket = nullptr;
    uint32_t* d_filter = nullptr;
    uint32_t filter_words = 0;
    bool filter_linear = false;     // the filter of THIS table uses the linear map (mpcr_core.cuh: 11-letter keys)
    float filter_scale = 0.f, filter_bias = 0.f;
    uint32_t n_keys = 0;
    bool dense = false;
    int ext_w = 0, ext_which = 0;   // seed extension (mpcr_ctx_set_seed_extension / mpcr_ctx_set_seed_blocks)
    int ext_block = 0, ext_gap = 0, ext_span = 0;   // block tables: letters per block, letters between seed and block, letters needed
    int samp_w = 0, samp_s = 0, samp_role = 0;   // position sampling (mpcr_ctx_set_sampling)
    uint32_t* d_bloom = nullptr;    // sampled tables: the first-level filter in global memory
    uint32_t bloom_shift = 0;
    uint32_t part = 0, parts = 1;   // table partition (mpcr_ctx_set_table_part)
    int append = 0;                 // mpcr_ctx_set_append
    int true_strands = 0;           // mpcr_ctx_set_true_strands
    int scan_w = 0;                 // word width the scanner keys on: ext_w for an extended table, else wordsize
    uint32_t max_hash_off = 0, max_len = 0;
    uint64_t max_pcr = 0;
    bool table_ready = false;
    // tiles
    TileDesc* d_tiles = nullptr;
    size_t tiles_cap = 0;
    uint32_t n_tiles = 0;
    uint64_t tiles_sig = 0;         // layout (contigs, origin, word size) the descriptor array was built for
    uint64_t tiles_sb = 0, tiles_se = 0;   // ... and the range it covers
    std::vector<uint64_t> h_tile_g;        // global first base of every descriptor (ascending): sub-range views
    std::vector<TileDesc> h_tiles;         // host copy of the descriptors (extent checks of a view)
    uint32_t view_first = 0, view_count = 0;   // the descriptors the next scan walks
    uint32_t lay_contigs = 0, lay_max_len = 0;  // bounds of the last scanned layout (sort digit counts)
    uint32_t* d_tile_counter = nullptr;  // [0] tile counter, [1..2] survivor count / verify cursor
    Survivor* d_surv = nullptr;
    size_t surv_bytes = 0;
    uint32_t* d_surv_ctl = nullptr;
    // sort scratch
    void* d_sort_tmp = nullptr;
    LongRun* d_long_runs = nullptr;   // order_ties -> order_long_runs queue + its two counters behind it
    size_t sort_tmp_cap = 0;
    uint32_t* d_counts = nullptr;
    size_t counts_cap = 0;
    uint8_t* d_lut = nullptr;  // 256 B genome LUT for pack
    uint64_t launches = 0;
    // scan / verify timing events, one set per pipeline slot (a caller that keeps two steps in flight reads step k's
    // times while step k+1 records its own, see mpcr_scan_sorted_async)
    cudaEvent_t evs[MPCR_MAX_SLOTS][3] = {};
    int ev_slot = 0;
    bool scan_timed[MPCR_MAX_SLOTS] = {};
    int env_debug = 0;          // $MPCR_DEBUG, read once at context creation
    long env_surv_cap = -1;     // $MPCR_SURVIVOR_CAP (test hook), ditto
    bool ctl_dirty = false;     // the scanner's control words were left non-zero (debug runs skip the verifier)
    uint8_t lut_host[256] = {}; // what d_lut holds (re-uploaded only when the caller's LUT changes)
    bool lut_valid = false;
    // bucket sort of short hit lists (mpcr_sort.cuh): contig table of the last layout on the device + scratch
    unsigned long long* d_contig_g = nullptr;
    size_t contig_g_cap = 0;
    uint32_t n_contig_g = 0;
    uint64_t genome_end = 0;    // padded coordinate behind the last contig of the last layout
    uint32_t* d_bsort = nullptr;   // [cnt kSortBuckets][off kSortBuckets + 1][pad][slot kBucketSortMax][flag]
    uint32_t* h_flag = nullptr;    // pinned landing place of the bucket sort's fall-back flag
    const void* attr_kern = nullptr;   // scanner instantiation whose shared-memory attribute is set (and its size)
    size_t attr_smem = 0;
    uint64_t tiles_need = 0;    // plane extent (bases from the origin) the cached descriptors read up to
    int64_t tiles_min = 0;      // ... and down to
};

static int ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return MPCR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) return fail(MPCR_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    *cap = want;
    return MPCR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// (1) pack
// ---------------------------------------------------------------------------------------------------------
// One thread packs 64 bases: 4 x 16-byte loads -> 1 valid word, 2 plane2 words, 4 plane4 words.
// lut[c] = nibble | code2 << 4 | clean << 6.
// The source may start at any byte (records of the device-side FASTA ingest lie back to back in one buffer): a strip
// then reads the five ALIGNED 16-byte chunks that cover it and funnel-shifts them into place.  An aligned chunk that
// holds at least one byte of the record lies inside the caller's allocation (allocations begin and end on 16-byte
// boundaries), so nothing outside it is touched.
template <int WS>
__device__ __forceinline__ void pack_shift_words(const uint32_t (&a)[20], uint32_t byte_shift, uint32_t (&w)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __funnelshift_r(a[i + WS], a[i + WS + 1], byte_shift);
}
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ ascii, uint64_t n, uint64_t dst_rel,
                                                   uint64_t* __restrict__ plane2, uint64_t* __restrict__ plane4,
                                                   uint64_t* __restrict__ valid, const uint8_t* __restrict__ lut_g) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
    const uint64_t strip = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b0 = strip * 64;
    if (b0 >= n) return;
    uint64_t v = 0, p2[2] = {0, 0}, p4[4] = {0, 0, 0, 0};
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(ascii) & 15u);
    uint32_t w[16];
    if (b0 + 64 <= n && mis == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 t = *reinterpret_cast<const uint4*>(ascii + b0 + 16 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
    } else if (b0 + 64 <= n) {
        uint32_t a[20];
        const uint4* src = reinterpret_cast<const uint4*>(ascii - mis + b0);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const uint4 t = src[q];
            a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
        }
        const uint32_t bs = 8u * (mis & 3u);
        switch (mis >> 2) {   // the same for every thread of the launch
            case 0: pack_shift_words<0>(a, bs, w); break;
            case 1: pack_shift_words<1>(a, bs, w); break;
            case 2: pack_shift_words<2>(a, bs, w); break;
            default: pack_shift_words<3>(a, bs, w); break;
        }
    } else {   // the last strip of the call
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t x = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint64_t i = b0 + 4 * k + b;
                // bytes past n encode as 0 ('\0' maps to an invalid zero nibble in every LUT)
                const uint32_t c = i < n ? ascii[i] : 0u;
                x |= c << (8 * b);
            }
            w[k] = x;
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = 4 * k + b;  // base index inside the strip
            const uint32_t e = lut[(w[k] >> (8 * b)) & 0xFFu];
            p4[j >> 4] |= (uint64_t)(e & 15u) << (4 * (j & 15));
            p2[j >> 5] |= (uint64_t)((e >> 4) & 3u) << (2 * (j & 31));
            v |= (uint64_t)((e >> 6) & 1u) << j;
        }
    }
    const uint64_t s = dst_rel / 64 + strip;
    valid[s] = v;
    plane2[2 * s] = p2[0];
    plane2[2 * s + 1] = p2[1];
    plane4[4 * s] = p4[0];
    plane4[4 * s + 1] = p4[1];
    plane4[4 * s + 2] = p4[2];
    plane4[4 * s + 3] = p4[3];
}

// plane4 -> plane2 + valid for sequence that reached the device as packed nibbles (host packer, mpcr_hostpack.cpp):
// a nibble with exactly one bit set is A/C/G/T (clean, 2-bit code = the bit's index), everything else is not clean
// and has code 0 -- the same three planes pack_kernel writes from ASCII.  One thread = 64 bases: 32 bytes in, 24 out.
__device__ __forceinline__ uint32_t gather_lsb4(uint64_t x) {   // bit 0 of each of the 16 nibbles -> 16 bits
    x &= kLane1;
    x = (x | (x >> 3)) & 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    return (uint32_t)((x | (x >> 24)) & 0xFFFFull);
}
__device__ __forceinline__ uint32_t gather_low2(uint64_t x) {   // bits 0..1 of each nibble -> 32 bits
    x &= 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    return (uint32_t)((x | (x >> 16)) & 0xFFFFFFFFull);
}
__global__ void __launch_bounds__(256) derive_planes_kernel(const uint64_t* __restrict__ plane4, uint64_t n, uint64_t dst_rel,
                                                            uint64_t* __restrict__ plane2, uint64_t* __restrict__ valid) {
    const uint64_t strip = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (strip * 64 >= n) return;
    const uint64_t s = dst_rel / 64 + strip;
    const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(plane4 + 4 * s);
    const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(plane4 + 4 * s + 2);
    const uint64_t w[4] = {q0.x, q0.y, q1.x, q1.y};
    uint64_t v = 0, p2[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t x = w[k];
        const uint64_t b0 = x & kLane1, b1 = (x >> 1) & kLane1, b2 = (x >> 2) & kLane1, b3 = (x >> 3) & kLane1;
        const uint64_t t = (b0 + b1 + b2 + b3) ^ kLane1;                        // 0 in the nibbles that hold one bit
        const uint64_t one = ~(t | (t >> 1) | (t >> 2) | (t >> 3)) & kLane1;     // ... as a nibble-LSB mask
        const uint64_t code = (((b1 | b3) & one)) | (((b2 | b3) & one) << 1);     // C1 G2 T3 in bits 0..1 of the nibble
        v |= (uint64_t)gather_lsb4(one) << (16 * k);
        p2[k >> 1] |= (uint64_t)gather_low2(code) << (32 * (k & 1));
    }
    valid[s] = v;
    plane2[2 * s] = p2[0];
    plane2[2 * s + 1] = p2[1];
}

// ---------------------------------------------------------------------------------------------------------
// (2) table build
// ---------------------------------------------------------------------------------------------------------
struct BlobFwd {
    const uint8_t* p;
    __device__ uint8_t operator()(int i) const { return p[i]; }
};
struct BlobRc {  // engine.py:357-359 reverse complement, on the fly
    const uint8_t* p;
    int len;
    __device__ uint8_t operator()(int i) const { return complement_of(p[len - 1 - i]); }
};

// One thread per table item.  Ordinary tables: item = record slot r = 2*line + strand.  Sampled tables (samp_role 1): item =
// r * samp_s + d, the window of the record's first primer at offset hash_off + d.
//   "+" : (P1,P2) = (primer1, primer2)   (engine.py:265-268)      "-" : (P1,P2) = (primer2, revcomp(primer1)) (engine.py:273-279)
__global__ void __launch_bounds__(128) encode_records(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ off,
                                                      const uint32_t* __restrict__ pcr, uint32_t n_lines,
                                                      const uint8_t* __restrict__ plut,
                                                      const uint32_t* __restrict__ word_off,  // 2*n_rec+1 prefix
                                                      int W, int w_scan, int which, int ext_block, int ext_gap, int ext_span,
                                                      int true_strands,
                                                      uint32_t part, uint32_t parts, int samp_w, int samp_s, int samp_role,
                                                      RecMeta* __restrict__ meta,
                                                      uint64_t* __restrict__ pwords, Item<2>* __restrict__ pairs,
                                                      uint32_t* __restrict__ tags, uint32_t* __restrict__ stats) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per = samp_role == 1 ? (uint32_t)samp_s : 1u;
    if (t >= 2 * n_lines * per) return;
    const uint32_t r = t / per;
    const int d = (int)(t - r * per);
    const bool lead = d == 0;          // the item that writes the record's meta data and primer words
    const uint32_t line = r >> 1;
    const bool minus = r & 1u;
    const uint64_t a0 = off[2 * line], a1 = off[2 * line + 1], a2 = off[2 * line + 2];
    const uint8_t* pr1 = blob + a0;
    const uint8_t* pr2 = blob + a1;
    const int n1 = (int)(a1 - a0), n2 = (int)(a2 - a1);
    // P1 is literal on both strands: primer1 for "+", primer2 for "-"
    const BlobFwd q1{minus ? pr2 : pr1};
    const int l1 = minus ? n2 : n1, l2 = minus ? n1 : n2;
    RecMeta m;
    m.pcr_size = pcr[line];
    m.p1_word = word_off[2 * r];
    m.p2_word = word_off[2 * r + 1];
    m.len1 = (uint16_t)l1; m.len2 = (uint16_t)l2;
    uint32_t hbe = 0, kext = 0;
    const int ho = first_clean_word(q1, l1, W, &hbe);
    // which: 0 = every record keyed by its W-mer; 1 = only records whose seed cannot be lengthened to w_scan letters;
    //        2 = only those that can, keyed by the lengthened word (mpcr_ctx_set_seed_extension)
    // block tables (mpcr_ctx_set_seed_blocks): "can be lengthened" = all ext_span letters are plain, and the key is the seed
    // plus the block ext_gap letters behind it; the tag starts behind that block
    const bool ext = which != 0 && (ext_block ? blocked_seed(q1, l1, ho, W, ext_block, ext_gap, ext_span, &kext)
                                              : extended_seed(q1, l1, ho, w_scan, &kext));
    m.tag = ho >= 0 ? make_tag(q1, l1, ho, which == 2 ? w_scan + ext_gap : W) : 0u;
    if (lead) {
        encode_primer(q1, l1, plut, pwords + m.p1_word);
        if (!minus) {
            // reference (engine.py:267): primer2 literally; me-PCR-true strands: its reverse complement
            if (true_strands) encode_primer(BlobRc{pr2, n2}, n2, plut, pwords + m.p2_word);
            else encode_primer(BlobFwd{pr2}, n2, plut, pwords + m.p2_word);
        } else {
            encode_primer(BlobRc{pr1, n1}, n1, plut, pwords + m.p2_word);
        }
    }
    bool here = ho >= 0 && (which == 0 || (which == 1 ? !ext : ext)) && (parts <= 1u || line % parts == part);
    uint32_t key = which == 2 ? kext : reverse_digits(hbe, W), tag = m.tag;
    if (samp_role != 0) {   // mpcr_ctx_set_sampling: 1 = this table holds the sampled windows, 2 = it holds the rest
        const bool can = sampleable_seed(q1, l1, ho, samp_w, samp_s);
        if (samp_role == 1) {
            here = can && (parts <= 1u || line % parts == part);
            key = 0;
            if (can) extended_seed(q1, l1, ho + d, samp_w, &key);
            tag = can ? make_tag(q1, l1, ho + d, samp_w) : 0u;
        } else {
            here = here && !can;
        }
    }
    m.hash_be = hbe;
    m.key = which == 2 ? kext : reverse_digits(hbe, W);
    m.hash_off = (uint16_t)(ho < 0 ? 0 : ho);
    m.flags = (ho >= 0 ? 1u : 0u) | (here ? 2u : 0u);
    if (lead) meta[r] = m;
    pairs[t].f[0] = key;
    pairs[t].f[1] = t | (here ? 0u : 0x80000000u);
    tags[t] = tag;
    if (here) atomicAdd(&stats[2], 1u);
    if (ho >= 0 && lead) {
        atomicAdd(&stats[0], 1u);
        atomicMax(&stats[1], (uint32_t)ho);
    }
}

// After the pairs are sorted by (invalid, key, record): CSR bucket entries (with inline tags), the 16-byte slot
// table (direct-indexed or open-addressed) and the two-bit-per-key blocked Bloom filter.
__global__ void __launch_bounds__(256) build_buckets(const Item<2>* __restrict__ pairs, uint32_t n_valid,
                                                     const uint32_t* __restrict__ tags, BucketEntry* __restrict__ bucket,
                                                     Slot* __restrict__ slots, SlotMap sm,
                                                     uint32_t* __restrict__ filter, uint32_t filter_words, uint32_t cw,
                                                     int W, uint32_t* __restrict__ n_keys,
                                                     uint32_t* __restrict__ bloom, uint32_t bloom_shift, bool linear,
                                                     float lin_scale, float lin_bias) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    const uint32_t key = pairs[i].f[0];
    const uint32_t rec = pairs[i].f[1] & 0x7FFFFFFFu;
    const bool head = (i == 0) || (pairs[i - 1].f[0] != key);
    const bool last = (i + 1 == n_valid) || (pairs[i + 1].f[0] != key);
    const uint32_t tag = tags[rec];   // rec = item index: a record slot, or (record, window) of a sampled table
    bucket[i] = BucketEntry{rec | (last ? 0x80000000u : 0u), tag};
    if (head) {
        // bucket size, capped at 3, and the second record's tag
        uint32_t n = 1, tag_b = 0;
        if (!last) {
            n = 2;
            tag_b = tags[pairs[i + 1].f[1] & 0x7FFFFFFFu];
            if (i + 2 < n_valid && pairs[i + 2].f[0] == key) n = 3;
        }
        // one record: both tags are its tag; two: one tag each; three or more: mask 0 = never reject
        const uint32_t code = n == 1 ? rec : (kWalkBucket | i);
        const uint32_t ta = n >= 3 ? 0u : tag, tb = n == 1 ? tag : (n == 2 ? tag_b : 0u);
        const unsigned long long lo = (unsigned long long)code | ((unsigned long long)ta << 32);
        const unsigned long long hi = (unsigned long long)tb | ((unsigned long long)key << 32);
        uint32_t s = slot_index(key, sm);
        if (!sm.direct) {
            for (;;) {  // claim on the (code, tag_a) half -- code == kSlotEmpty marks a free slot
                unsigned long long* h = reinterpret_cast<unsigned long long*>(slots + s);
                if (atomicCAS(h, ~0ull, lo) == ~0ull) break;
                s = (s + 1) & sm.mask;
            }
        } else {
            reinterpret_cast<unsigned long long*>(slots + s)[0] = lo;
        }
        reinterpret_cast<unsigned long long*>(slots + s)[1] = hi;
        if (bloom) atomicOr(&bloom[bloom_word(key, bloom_shift)], bloom_bits(key));   // sampled table: global filter
        else if (linear) atomicOr(&filter[filter_word_linear(key, lin_scale, lin_bias)], filter_bits_linear(key));
        else atomicOr(&filter[filter_word(key, cw, filter_words)], filter_bits_of(key, W));
        atomicAdd(n_keys, 1u);
    }
}

// Open-addressed tables, after build_buckets: flag every slot a stored key had to step over (kSlotChain).
__global__ void __launch_bounds__(256) mark_chains(const Item<2>* __restrict__ pairs, uint32_t n_valid,
                                                   Slot* __restrict__ slots, SlotMap sm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    const uint32_t key = pairs[i].f[0];
    if (i != 0 && pairs[i - 1].f[0] == key) return;   // one walk per distinct key
    for (uint32_t s = slot_index(key, sm); slots[s].key != key; s = (s + 1) & sm.mask) atomicOr(&slots[s].code, kSlotChain);
}

// ---------------------------------------------------------------------------------------------------------
// (3)(4)(5) scan
// ---------------------------------------------------------------------------------------------------------
// A seed position that survived the Bloom filter, the exact table probe and the inline tag: (tile, offset, record).
struct Survivor {
    uint32_t tile, lp, code, pad;  // code: record index, or kWalkBucket | first bucket entry
};

struct ScanArgs {
    const uint64_t* p2;
    const uint64_t* p4;
    const uint64_t* valid;
    const TileDesc* tiles;
    uint32_t n_tiles;
    const Slot* slots;
    SlotMap smap;
    const BucketEntry* bucket;
    const RecMeta* meta;
    const uint64_t* pwords;
    const uint32_t* filter;
    uint32_t filter_words;
    uint32_t cw;  // filter_mul(W)
    float lin_scale, lin_bias;   // linear filter map (11-letter keys): word index = mantissa of fma(2^23 + key, scale, bias)
    uint32_t lin_exp;            // kLinearExp, as an argument: a register operand keeps (x & mask) | exp ONE LOP3
    SearchParams prm;
    mpcr_hit* hits;
    unsigned long long capacity;
    unsigned long long* count;
    uint32_t* tile_counter;
    Survivor* surv;        // kSurvLists sub-lists of surv_cap entries each
    uint32_t surv_cap;
    uint32_t* surv_ctl;    // per sub-list, 128 bytes apart: [0] entries appended (may exceed surv_cap), [1] verify cursor
    int debug;  // MPCR_DEBUG bit0: stop after the filter stage; bit1: probe the table but drop the survivors
    // Pipe balancing of stage 1: a multiplier the compiler cannot see through, so that the filter word's address stays a
    // multiply-add (IMAD, fma pipe) instead of becoming a LEA on the busier alu pipe (2.97 -> 2.89 ms).
    uint32_t k4;          // 4
    // sampled tables (mpcr_ctx_set_sampling): every samp_s-th position is probed through a Bloom filter in global memory
    uint32_t samp_s;      // 1 = ordinary table; else survivor codes are record * samp_s + window
    const uint32_t* bloom;
    uint32_t bloom_shift;
};

// shared-memory plan of the scanner CTA (dynamic shared memory): one private block per warp, then the filter
struct ScanSmem {
    static constexpr int kWarps = kScanThreads / 32;
    static constexpr int kUnitBases = 32 * kPosPerThread;    // hash positions one warp scans per step (2048)
    static constexpr int kUnitsPerTile = kTileBases / kUnitBases;
    static constexpr int kP2Bytes = kUnitBases / 4 + 32;     // plane2 of a unit + 128-base read-ahead
    static constexpr int kVBytes = kUnitBases / 8 + 16;      // valid bits of a unit + 128-base read-ahead
    static constexpr int kQCap = MPCR_SCAN_QCAP;             // candidate queue entries
    static constexpr int kIlp = MPCR_SCAN_ILP;               // slot gathers in flight per lane
    struct Warp {
        alignas(16) uint8_t p2[2][kP2Bytes];
        alignas(16) uint8_t v[2][kVBytes];
        alignas(16) uint4 landing[kIlp][32];                 // cp.async targets of the slot gathers
        uint16_t queue[kQCap];
        unsigned long long mbar[2];
    };
    static constexpr int kCtaBarOff = (int)(sizeof(Warp) * kWarps);                 // one mbarrier for the filter's arrival
    static constexpr int kFilterOff = (kCtaBarOff + 8 + 127) / 128 * 128;
    static_assert(sizeof(Warp) % 8 == 0, "mbarrier alignment");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 16-byte asynchronous gather global -> shared that bypasses L1 (LDGSTS.BYPASS): unlike a plain load it does not
// need an L1 line while in flight, which is what keeps ~0.75 gathers/clk/SM alive next to a 160 KB filter.
__device__ __forceinline__ void gather16_async(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gather_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gather_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct HitEmitter {
    mpcr_hit* hits;
    unsigned long long capacity;
    unsigned long long* count;
    uint32_t contig, rec, hash_off;
    __device__ void operator()(int64_t pos1, int64_t pos2, uint32_t rank) const {
        unsigned long long i = atomicAdd(count, 1ull);
        if (i < capacity) {
            mpcr_hit h;
            h.contig = contig; h.pos1 = (uint32_t)pos1; h.pos2 = (uint32_t)pos2; h.rec = rec; h.rank = rank;
            h.hash_off = hash_off;
            hits[i] = h;
        }
    }
};

// engine.py:483-489 + 507-597 for one survivor, serially in one thread (overflow path of the survivor list).
__device__ __noinline__ void verify_serial(const ScanArgs& a, uint32_t tile, uint32_t lp, uint32_t code) {
    const TileDesc td = a.tiles[tile];
    const int64_t gb = td.gbase + lp + a.prm.W + a.prm.gap;
    const uint32_t gcodes = fetch_bits(a.p2, 2 * gb, 2 * kTagBases), gvalid = fetch_bits(a.valid, gb, kTagBases);
    for_each_survivor_record(a.bucket, code, gcodes, gvalid, a.prm.N, [&](uint32_t item) {
        // sampled tables: item = record * samp_s + window; the window lies that many bases behind the seed
        const uint32_t rec = a.samp_s > 1 ? item / a.samp_s : item, d = a.samp_s > 1 ? item - rec * a.samp_s : 0u;
        const RecMeta m = a.meta[rec];
        verify_record(a.p4, td.gbase - (int64_t)td.lstart, (int64_t)td.length, (int64_t)td.lstart + lp - (int64_t)d, m,
                      a.pwords, a.prm, HitEmitter{a.hits, a.capacity, a.count, td.contig, rec, (uint32_t)m.hash_off});
    });
}

// A seed position matched a table entry whose tags did not rule it out: hand it to verify_kernel.  The survivor list
// is split into kSurvLists sub-lists (one counter each, on its own 128-byte line) chosen by the warp, because with
// N = 2 or ambiguity-rich sequence a human-sized scan appends 10^7 survivors and one hot counter would serialise them.
static constexpr uint32_t kSurvLists = 256;
static constexpr uint32_t kSurvCtlStride = 32;   // uint32 units = 128 bytes
__device__ __noinline__ void push_survivor(const ScanArgs& a, uint32_t tile, uint32_t lp, uint32_t code) {
    const uint32_t list = (blockIdx.x + (threadIdx.x >> 5) * gridDim.x) & (kSurvLists - 1u);
    const uint32_t k = atomicAdd(a.surv_ctl + list * kSurvCtlStride, 1u);
    if (k < a.surv_cap) a.surv[(size_t)list * a.surv_cap + k] = Survivor{tile, lp, code, 0u};
    else verify_serial(a, tile, lp, code);   // sub-list full: verify right here
}

// Hashed mode only: the first probe hit another key's slot -- continue the probe sequence with plain loads.
__device__ __noinline__ void probe_collision(const ScanArgs& a, uint32_t key, uint32_t gcodes, uint32_t gvalid,
                                             uint32_t tile, uint32_t lp) {
    Slot s;
    if (!find_slot(a.slots, a.smap, key, &s)) return;
    if (slot_survives(s, gcodes, gvalid, a.prm.N)) push_survivor(a, tile, lp, s.code);
}

// One filter probe: returns a word whose MSB is set iff both Bloom bits of the key are set.
// x holds the key in its low 2W bits (anything above cancels in the multiply by cw); x3 is the raw register of
// the position three bases further on, i.e. x >> 6 in its low bits (WIDE: W >= 6, else the second bit == the first).
struct FilterView {
    const uint32_t* words;   // the filter in shared memory
    uint32_t base32;         // its shared-window address
    uint32_t cw, n_words;
};
template <bool WIDE>
__device__ __forceinline__ uint32_t filter_probe(const FilterView& f, const ScanArgs& a, uint32_t x, uint32_t x3) {
    uint32_t word;
#ifdef MPCR_STAGE1_NOCONFLICT   // tuning build, WRONG results: every lane reads its own bank -- what stage 1 would cost without bank conflicts
    uint32_t lane_id;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane_id));
    const uint32_t addr = ((__umulhi(x * f.cw, f.n_words - 32u) & ~31u) | lane_id) * a.k4 + f.base32;
#else
    const uint32_t addr = __umulhi(x * f.cw, f.n_words) * a.k4 + f.base32;
#endif
    asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
    const uint32_t t1 = __funnelshift_l(0u, word, x);  // word << (x & 31)
    if (!WIDE) return t1;
    return t1 & __funnelshift_l(0u, word, x3);
}

// Stage 2 for up to 32*R queued positions of a warp starting at queue index base (R rounds of 32 async slot
// gathers in flight, then the tag check on what came back).  Straight-line code: lanes past the end of the queue
// re-probe entry 0 and are masked out, so the R rounds interleave freely.  CLEAN: every base of the unit (and its
// read-ahead) is A/C/G/T, so the tag verdict can be used without looking at the valid bits.
template <bool CLEAN, bool HASHED, int R, int WC, int NC, bool GAPPED>
__device__ __forceinline__ void probe_rounds(const ScanArgs& a, const uint16_t* __restrict__ queue, uint4 (*landing)[32],
                                             const uint32_t* __restrict__ s_p2, const uint32_t* __restrict__ s_v,
                                             uint32_t base, uint32_t cnt, int lane, uint32_t tile, uint32_t ubase,
                                             unsigned long long& n_dbg) {
    // WC / NC: word size and mismatch budget known at compile time (the reference's default -W 11 with -N 0..2), 0 / -1
    // = read them from the arguments
    const int W = WC ? WC : a.prm.W, N = NC >= 0 ? NC : a.prm.N;
    const uint32_t wmask = wmask_of(W);
    // block tables: the key's second part and the tag window lie `gap` letters further on (W + gap <= 16)
    const int gap = GAPPED ? a.prm.gap : 0;
    const uint32_t seed_mask = GAPPED ? wmask_of(W - a.prm.block) : 0u;
    uint32_t lpv[R], key[R], gcodes[R];
    bool ok[R], dirty[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const uint32_t qi = base + 32 * u + lane;
        ok[u] = qi < cnt;
        const uint32_t lp = queue[ok[u] ? qi : 0u];
        // key (2W bits) and the 8-base tag window (16 bits) out of three staged words
        const uint32_t wi = lp >> 4, sh = lp * 2u;   // the funnel shifts below wrap: only (2 * lp) & 31 counts
        const uint32_t w0 = s_p2[wi], w1 = s_p2[wi + 1], w2 = s_p2[wi + 2];
        const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
        key[u] = (GAPPED ? gap_key_raw(x0, seed_mask, gap) : x0) & wmask;
        gcodes[u] = __funnelshift_rc(x0, x1, 2 * (W + gap));  // clamped: 32 -> x1; only the low 16 bits are used
        lpv[u] = lp;
        dirty[u] = false;
        if (!CLEAN) {
            const uint32_t vb = lp + (uint32_t)(W + gap), vi = vb >> 5, vs = vb & 31u;
            dirty[u] = !tag_window_clean(__funnelshift_r(s_v[vi], s_v[vi + 1], vs));
        }
        {   // slot address as a wide multiply-add: mask + LEA + LEA.HI.X (the compiler's shift + mask + 64-bit add takes four)
            const uint32_t si = HASHED ? (slot_hash(key[u]) & a.smap.mask) : key[u];
            unsigned long long sa;
            asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(sa) : "r"(si), "l"(a.slots));   // becomes LEA + LEA.HI.X
            gather16_async(&landing[u][lane], reinterpret_cast<const void*>(sa));
        }
        gather_commit();
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
        if (u == 0) gather_wait<R - 1>();
        else if (u == 1) gather_wait<(R > 2 ? R - 2 : 0)>();
        else if (u == 2) gather_wait<(R > 3 ? R - 3 : 0)>();
        else gather_wait<0>();
        // (reading only the first 8 bytes -- code and first tag settle almost every candidate -- was measured and buys
        // nothing: 8-byte reads at a 16-byte stride cost the same shared-memory wavefronts as the 16-byte read)
        const uint4 sl = landing[u][lane];
        const uint2 v = make_uint2(sl.x, sl.y);
        const uint32_t tag_b = sl.z, skey = sl.w;
        // HASHED (open-addressed table, W >= 12): another key's slot ends the search unless a stored key's probe
        // sequence runs through it (kSlotChain); direct tables never collide
        const bool other = HASHED && skey != key[u];
        const bool collide = other && (v.x & kSlotChain);
        const bool pass = !other && (dirty[u] || !(tag_rejects(v.y, gcodes[u], N) && tag_rejects(tag_b, gcodes[u], N)));
        if (ok[u] && v.x != kSlotEmpty && (collide || pass)) {  // about one queued position in a thousand
            const uint32_t lp = ubase + lpv[u];
            if (a.debug & 2) ++n_dbg;
            else if (collide) probe_collision(a, key[u], gcodes[u] & 0xFFFFu, dirty[u] ? 0u : 0xFFu, tile, lp);
            else push_survivor(a, tile, lp, HASHED ? (v.x & ~kSlotChain) : v.x);
        }
    }
}

#ifdef MPCR_PROBE_NOINLINE
#define MPCR_PROBE_INLINE __noinline__
#else
#define MPCR_PROBE_INLINE __forceinline__
#endif
template <bool CLEAN, bool HASHED, int WC, int NC, bool GAPPED>
__device__ MPCR_PROBE_INLINE void probe_queue(const ScanArgs& a, const uint16_t* __restrict__ queue, uint4 (*landing)[32],
                                            const uint32_t* __restrict__ s_p2, const uint32_t* __restrict__ s_v,
                                            uint32_t cnt, int lane, uint32_t tile, uint32_t ubase,
                                            unsigned long long& n_dbg) {
    constexpr int kIlp = ScanSmem::kIlp;
    static_assert(kIlp >= 1 && kIlp <= 4, "kIlp");
    uint32_t base = 0;
    for (; base + 32 * kIlp <= cnt; base += 32 * kIlp)
        probe_rounds<CLEAN, HASHED, kIlp, WC, NC, GAPPED>(a, queue, landing, s_p2, s_v, base, cnt, lane, tile, ubase, n_dbg);
    const uint32_t rounds = (cnt - base + 31) >> 5;  // tail: only the rounds that hold something
    if (rounds == 1) probe_rounds<CLEAN, HASHED, 1, WC, NC, GAPPED>(a, queue, landing, s_p2, s_v, base, cnt, lane, tile, ubase, n_dbg);
    else if (rounds == 2) probe_rounds<CLEAN, HASHED, (kIlp >= 2 ? 2 : 1), WC, NC, GAPPED>(a, queue, landing, s_p2, s_v, base, cnt, lane, tile, ubase, n_dbg);
    else if (rounds == 3) probe_rounds<CLEAN, HASHED, (kIlp >= 3 ? 3 : 1), WC, NC, GAPPED>(a, queue, landing, s_p2, s_v, base, cnt, lane, tile, ubase, n_dbg);
    else if (rounds == 4) probe_rounds<CLEAN, HASHED, (kIlp >= 4 ? 4 : 1), WC, NC, GAPPED>(a, queue, landing, s_p2, s_v, base, cnt, lane, tile, ubase, n_dbg);
}

// Stage 1 for one unit of 2048 positions: lane l owns positions [64l, 64l+64) -- rolling keys by funnel shift out of
// five registers, one shared-memory Bloom probe (two bits of one word) per position.  c_lo / c_hi: pass masks.
template <bool WIDE, bool GAPPED, bool LINEAR>
__device__ __forceinline__ void stage1_unit(const FilterView& f, const ScanArgs& a, const uint32_t* __restrict__ s_p2,
                                            const uint32_t* __restrict__ s_v, int lane, uint32_t unit_nbases, int W,
                                            uint32_t& c_lo, uint32_t& c_hi, bool& my_clean) {
    const uint32_t lp0 = (uint32_t)lane * kPosPerThread;  // unit-local
    if (lp0 < unit_nbases) {
        const uint2 v0 = *reinterpret_cast<const uint2*>(s_v + 2 * lane);
        const uint2 v1 = *reinterpret_cast<const uint2*>(s_v + 2 * lane + 2);
        my_clean = (v0.x & v0.y & v1.x) == 0xFFFFFFFFu;  // own 64 bases + the 32 behind them (W + tag <= 24)
        // W-mer validity of the 64 positions: all ones on clean sequence, else log-doubling over the valid bits
        uint64_t wv = ~0ull;
        if (!my_clean)
            wv = GAPPED ? window_valid_gapped(((uint64_t)v0.y << 32) | v0.x, ((uint64_t)v1.y << 32) | v1.x, W - a.prm.block,
                                              a.prm.block, a.prm.gap)
                        : window_valid(((uint64_t)v0.y << 32) | v0.x, ((uint64_t)v1.y << 32) | v1.x, W);
        const uint32_t left = unit_nbases - lp0;
        if (left < 64u) wv &= (1ull << left) - 1ull;
        if (wv) {
            const uint4 q = *reinterpret_cast<const uint4*>(s_p2 + 4 * lane);
            const uint32_t r[6] = {q.x, q.y, q.z, q.w, s_p2[4 * lane + 4], 0u};
            // raw register of position j (its low 2W bits are the key); positions 64..66 only feed shift amounts
            auto raw = [&](int j) -> uint32_t {
                return (j & 15) ? __funnelshift_r(r[j >> 4], r[(j >> 4) + 1], 2 * (j & 15)) : r[j >> 4];
            };
            // block tables: seed letters + the block `gap` letters behind them (garbage above the key cancels in the probe)
            const uint32_t seed_mask = GAPPED ? wmask_of(W - a.prm.block) : 0u;
            const int gap = GAPPED ? a.prm.gap : 0;
            auto probe = [&](int j) -> uint32_t {
                if (!GAPPED) return filter_probe<WIDE>(f, a, raw(j), raw(j + 3));
                const uint32_t k = gap_key_raw(raw(j), seed_mask, gap);
                return filter_probe<WIDE>(f, a, k, k >> 6);
            };
            if (LINEAR) {
                // Linear map (mpcr_core.cuh).  Every warp instruction of this loop costs an issue slot of ~1.6 clocks whatever
                // pipe it runs on (scripts/ubench/pipes.cu), so the work is packed two positions per FP32 instruction
                // (FFMA2, sm_100): positions (j, j + 22) share the index fma, the address fma and the accumulate.  Per
                // position: SHF (key), LOP3 (key -> float 2^23 + key), 1/2 FFMA2 (word index in the mantissa), 1/2 FFMA2
                // (index -> byte address: (2^23 + i) * 2^-147 + (base - 2^25) * 2^-149 is the DENORMAL whose bit pattern is
                // 4 i + base), LDS, two rotates that bring the tested bits to bit 23 -- the float 2^-126 --, LOP3 (and),
                // 1/2 FFMA2 (Horner step acc = 2 acc + t, exact for 22 steps): 7.5 instructions against 10.
                const uint32_t lin_exp = a.lin_exp;
                const float addr_c = -(float)(33554432u - f.base32) * __uint_as_float(1u);   // (base - 2^25) * 2^-149, exact
                auto pair_step = [&](int ja, int jb, uint32_t& acc_a, uint32_t& acc_b) {
                    const uint32_t xa = raw(ja), xa1 = raw(ja + 1), xb = raw(jb), xb1 = raw(jb + 1);
                    uint32_t fa, fb, ga, gb;
                    asm("lop3.b32 %0, %1, 0x3FFFFF, %2, 0xEA;" : "=r"(fa) : "r"(xa), "r"(lin_exp));
                    asm("lop3.b32 %0, %1, 0x3FFFFF, %2, 0xEA;" : "=r"(fb) : "r"(xb), "r"(lin_exp));
                    asm("{\n .reg .b64 v, s, b;\n mov.b64 v, {%2, %3};\n mov.b64 s, {%4, %4};\n mov.b64 b, {%5, %5};\n"
                        " fma.rn.f32x2 v, v, s, b;\n mov.b64 {%0, %1}, v;\n}\n"
                        : "=r"(ga), "=r"(gb)
                        : "r"(fa), "r"(fb), "f"(a.lin_scale), "f"(a.lin_bias));
                    uint32_t addr_a, addr_b;
                    asm("{\n .reg .b64 v, s, b;\n mov.b64 v, {%2, %3};\n mov.b64 s, {%4, %4};\n mov.b64 b, {%5, %5};\n"
                        " fma.rn.f32x2 v, v, s, b;\n mov.b64 {%0, %1}, v;\n}\n"
                        : "=r"(addr_a), "=r"(addr_b)
                        : "r"(ga), "r"(gb), "f"(__uint_as_float(4u)), "f"(addr_c));
                    uint32_t wa, wb;
                    asm("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(addr_a));
                    asm("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(addr_b));
                    const uint32_t ta = __funnelshift_l(wa, wa, xa) & __funnelshift_l(wa, wa, xa1) & 0x00800000u;
                    const uint32_t tb = __funnelshift_l(wb, wb, xb) & __funnelshift_l(wb, wb, xb1) & 0x00800000u;
                    asm("{\n .reg .b64 v, s, t;\n mov.b64 v, {%0, %1};\n mov.b64 s, {%4, %4};\n mov.b64 t, {%2, %3};\n"
                        " fma.rn.f32x2 v, v, s, t;\n mov.b64 {%0, %1}, v;\n}\n"
                        : "+r"(acc_a), "+r"(acc_b)
                        : "r"(ta), "r"(tb), "f"(2.0f));
                };
                uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;   // positions 0..21, 22..43, 44..53, 54..63 (float bits)
#pragma unroll
                for (int i = 21; i >= 0; --i) pair_step(i, i + 22, acc0, acc1);
#pragma unroll
                for (int i = 9; i >= 0; --i) pair_step(44 + i, 54 + i, acc2, acc3);
                // acc = sum of pass bits * 2^(k - 126): times 2^126, plus 2^23, and the mantissa is the bit mask
                auto mask_of = [](uint32_t acc) -> uint32_t {
                    return __float_as_uint(__fmaf_rn(__uint_as_float(acc), __uint_as_float(0x7E800000u), 8388608.0f)) & 0x3FFFFFu;
                };
                const uint32_t m0 = mask_of(acc0), m1 = mask_of(acc1), m2 = mask_of(acc2), m3 = mask_of(acc3);
                c_lo = m0 | (m1 << 22);
                c_hi = (m1 >> 10) | (m2 << 12) | (m3 << 22);
            } else {
                // collect the MSB of each probe result as bit j of the pass mask (descending j: one funnel shift each)
#pragma unroll
                for (int j = 31; j >= 0; --j) c_lo = __funnelshift_l(probe(j), c_lo, 1);
#pragma unroll
                for (int j = 63; j >= 32; --j) c_hi = __funnelshift_l(probe(j), c_hi, 1);
            }
            c_lo &= (uint32_t)wv;
            c_hi &= (uint32_t)(wv >> 32);
        }
    }
}

// Persistent CTAs, one per SM, made of AUTONOMOUS warps: there is no CTA-wide barrier after the prologue, so
// the warps drift apart and the ALU/shared-memory work of stage 1 overlaps the gather latency of stage 2.
// Each warp walks units of 2048 hash positions (unit u of tile u/16), strided over all warps of the grid.
//   staging : plane2 + valid bits of a unit (+128 bases read-ahead) arrive by two TMA bulk copies on the warp's
//             own mbarrier, double buffered (the next unit loads while this one is scanned).
//   stage 1 : lane l owns positions [64l, 64l+64) -- rolling keys by funnel shift out of five registers, one
//             shared-memory Bloom probe (two bits of one word) per position.
//   stage 2 : surviving positions are compacted into the warp's queue and processed with full lanes, kIlp
//             16-byte L1-bypassing async gathers of the slot table in flight per lane; key + tag window come
//             from the staged unit; inline tag check.
//   What is left (about one position in 10^4) goes to the survivor list for verify_kernel.
template <bool WIDE, bool HASHED, int WC, int NC, bool GAPPED = false>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanArgs a) {
    // the instantiations for contiguous 11-letter keys in a direct table read the filter through the linear map
    constexpr bool LINEAR = MPCR_LINEAR_FILTER && WC == kLinearW && !HASHED && !GAPPED;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ScanSmem::Warp& ws = reinterpret_cast<ScanSmem::Warp*>(smem)[warp];
    uint32_t* s_filter = reinterpret_cast<uint32_t*>(smem + ScanSmem::kFilterOff);

    // Warp gw walks tiles gw, gw + stride, ... and inside a tile its units of 2048 positions in order.  Two cursors
    // run over that sequence: the load cursor (l_*) is one unit ahead of the process cursor (p_*).
    const uint32_t stride = gridDim.x * ScanSmem::kWarps;
    const uint32_t gw = (uint32_t)warp * gridDim.x + blockIdx.x;  // neighbouring tiles spread over the SMs
    uint32_t l_tile = gw, l_sub = 0, l_nsub = 0, p_tile = gw, p_sub = 0, p_nbases = 0;
    int64_t l_gbase = 0;
    if (gw < a.n_tiles) {
        const TileDesc td = a.tiles[gw];
        l_gbase = td.gbase;
        p_nbases = td.nbases;
        l_nsub = (td.nbases + ScanSmem::kUnitBases - 1) / ScanSmem::kUnitBases;
    }
    // lane 0 starts the two bulk copies of the load cursor's unit; then every lane advances the cursor
    auto issue_next = [&](int buf) {
        if (lane == 0) {
            const int64_t gb = l_gbase + (int64_t)l_sub * ScanSmem::kUnitBases;
            mbar_expect_tx(&ws.mbar[buf], ScanSmem::kP2Bytes + ScanSmem::kVBytes);
            tma_load_1d(ws.p2[buf], reinterpret_cast<const uint8_t*>(a.p2) + (gb >> 2), ScanSmem::kP2Bytes, &ws.mbar[buf]);
            tma_load_1d(ws.v[buf], reinterpret_cast<const uint8_t*>(a.valid) + (gb >> 3), ScanSmem::kVBytes, &ws.mbar[buf]);
        }
        if (++l_sub >= l_nsub) {
            l_tile += stride;
            l_sub = 0;
            if (l_tile < a.n_tiles) {
                l_gbase = __ldg(&a.tiles[l_tile].gbase);
                l_nsub = (__ldg(&a.tiles[l_tile].nbases) + ScanSmem::kUnitBases - 1) / ScanSmem::kUnitBases;
            }
        }
    };

    if (lane == 0) {
        mbar_init(&ws.mbar[0], 1);
        mbar_init(&ws.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // The filter (~160 KB) arrives by TMA bulk copies that one thread starts: a few large transfers per SM instead of
    // 10^4 16-byte loads through the LSU, and the warps' first units are already on their way behind it.
    unsigned long long* fbar = reinterpret_cast<unsigned long long*>(smem + ScanSmem::kCtaBarOff);
    if (tid == 0) {
        mbar_init(fbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = a.filter_words * 4u;
        mbar_expect_tx(fbar, bytes);
        for (uint32_t o = 0; o < bytes; o += 32768u)
            tma_load_1d(reinterpret_cast<uint8_t*>(s_filter) + o, reinterpret_cast<const uint8_t*>(a.filter) + o,
                        min(32768u, bytes - o), fbar);
    }
    __syncwarp();
    if (l_tile < a.n_tiles) issue_next(0);
    __syncthreads();            // fbar is initialised
    while (!mbar_try_wait(fbar, 0u)) {}

    const int W = WC ? WC : a.prm.W;
    const FilterView fv{s_filter, smem_u32(s_filter), a.cw, a.filter_words};
    unsigned long long n_dbg = 0;

    for (uint32_t it = 0; p_tile < a.n_tiles; ++it) {
        const int buf = it & 1;
        if (l_tile < a.n_tiles) issue_next(buf ^ 1);
        const uint32_t tile = p_tile;
        const uint32_t ubase = p_sub * ScanSmem::kUnitBases;  // tile-local offset of the unit
        const uint32_t unit_nbases = min(p_nbases - ubase, (uint32_t)ScanSmem::kUnitBases);
        if (ubase + ScanSmem::kUnitBases >= p_nbases) {  // last unit of the tile: move the process cursor on
            p_tile += stride;
            p_sub = 0;
            if (p_tile < a.n_tiles) p_nbases = __ldg(&a.tiles[p_tile].nbases);
        } else {
            ++p_sub;
        }
        while (!mbar_try_wait(&ws.mbar[buf], (it >> 1) & 1u)) {}
        const uint32_t* s_p2 = reinterpret_cast<const uint32_t*>(ws.p2[buf]);
        const uint32_t* s_v = reinterpret_cast<const uint32_t*>(ws.v[buf]);

        // ---------------- stage 1: rolling keys + Bloom probe ----------------
        const uint32_t lp0 = (uint32_t)lane * kPosPerThread;  // unit-local
        uint32_t c_lo = 0, c_hi = 0;
        bool my_clean = false;
        stage1_unit<WIDE, GAPPED, LINEAR>(fv, a, s_p2, s_v, lane, unit_nbases, W, c_lo, c_hi, my_clean);
#ifdef MPCR_STAGE1_ONLY   // tuning builds: the kernel without its stage 2 (what the register allocator does to stage 1 alone)
        n_dbg += __popc(c_lo) + __popc(c_hi);
        __syncwarp();
        continue;
#else
        if (a.debug & 1) {
            n_dbg += __popc(c_lo) + __popc(c_hi);
            __syncwarp();
            continue;
        }

        // ---------------- stage 2: warp queue, async slot gathers, tag check ----------------
        const bool all_clean = __all_sync(0xffffffffu, my_clean);
        for (;;) {
            // reserve queue space: lanes take their share in lane order while it fits (a lane has <= 64 <= kQCap)
            const uint32_t n = __popc(c_lo) + __popc(c_hi);
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) break;
            const bool fits = incl <= (uint32_t)ScanSmem::kQCap;
            if (fits && n) {
                uint16_t* q = ws.queue + (incl - n);
                // highest bit first (FLO alone; the lowest bit needs BREV + FLO): the order inside the queue is free
                while (c_lo) {
                    const uint32_t b = 31u - (uint32_t)__clz(c_lo);
                    *q++ = (uint16_t)(lp0 + b);
                    c_lo ^= 1u << b;
                }
                while (c_hi) {
                    const uint32_t b = 31u - (uint32_t)__clz(c_hi);
                    *q++ = (uint16_t)(lp0 + 32u + b);
                    c_hi ^= 1u << b;
                }
            }
            const uint32_t fit_mask = __ballot_sync(0xffffffffu, fits);
            const uint32_t cnt = total <= (uint32_t)ScanSmem::kQCap
                                     ? total
                                     : __shfl_sync(0xffffffffu, incl, 31 - __clz(fit_mask));  // lane 0 always fits
            __syncwarp();
            if (all_clean) probe_queue<true, HASHED, WC, NC, GAPPED>(a, ws.queue, ws.landing, s_p2, s_v, cnt, lane, tile, ubase, n_dbg);
            else probe_queue<false, HASHED, WC, NC, GAPPED>(a, ws.queue, ws.landing, s_p2, s_v, cnt, lane, tile, ubase, n_dbg);
            __syncwarp();
            if (fit_mask == 0xffffffffu) break;
        }
        __syncwarp();  // every lane is done with this buffer before lane 0 refills it two steps from now
#endif
    }
    if (a.debug) {
        for (int d = 16; d; d >>= 1) n_dbg += __shfl_xor_sync(0xffffffffu, n_dbg, d);
        if (lane == 0 && n_dbg) atomicAdd(a.count, n_dbg);
    }
}

// Dense tables: when a large share of all 4^W words are seeds (W = 8 with 10^5..10^6 STS, or W = 11 with 10^6) a
// Bloom filter cannot reject anything, so this scanner skips it and probes the slot table at EVERY valid position.
// No shared-memory carve-out is needed, so plain loads run at the full L1 rate (~1 random 16-byte load /clk/SM).
// A warp takes units of 2048 hash positions of a tile; lane l owns positions [64l, 64l+64):
//   phase A: the lane's bases arrive as 24 coalesced bytes of plane2 (+ 16 of the valid plane); rolling keys by
//            funnel shift like the sparse scanner, eight slot loads in flight per lane, inline tags for seeds with one
//            or two records, the tag window straight from the same registers;
//   phase B: seeds shared by three or more records are walked by the whole warp, 32 bucket entries per coalesced
//            256-byte load, four positions batched per round, each entry behind its own tag.
// Same survivor list and verify_kernel as the sparse path.
__global__ void __launch_bounds__(256, 2) dense_scan_kernel(const ScanArgs a) {
    constexpr int kUnit = 32 * kPosPerThread;
    constexpr int kChunk = 4;                      // slot loads in flight per lane (divides 16)
    const int lane = threadIdx.x & 31;
    const int W = a.prm.W, N = a.prm.N;
    const uint32_t wmask = wmask_of(W);
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) return;
        const TileDesc td = a.tiles[tile];
        for (uint32_t ubase = 0; ubase < td.nbases; ubase += kUnit) {
            const uint32_t lp0 = ubase + (uint32_t)lane * kPosPerThread;   // tile-local
            uint64_t walk = 0;    // positions whose seed is shared by >= 3 records (phase B)
            uint64_t dirty = 0;   // ... of those, the ones whose tag window is not clean
            uint32_t r[6] = {0, 0, 0, 0, 0, 0};
            if (lp0 < td.nbases) {
                const int64_t gpos = td.gbase + lp0;                    // multiple of 64
                const uint64_t* vp = a.valid + (gpos >> 6);
                const uint64_t v0 = __ldg(vp), v1 = __ldg(vp + 1);
                uint64_t wv = window_valid(v0, v1, W);                 // seed window hashable (engine.py:483)
                const uint64_t wt = window_valid(v0, v1, W + kTagBases); // ... and the tag window clean as well
                const uint32_t left = td.nbases - lp0;
                if (left < 64u) wv &= (1ull << left) - 1ull;
                if (wv) {
                    const uint32_t* pp = reinterpret_cast<const uint32_t*>(a.p2) + (gpos >> 4);
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(pp));
                    const uint2 q2 = __ldg(reinterpret_cast<const uint2*>(pp + 4));
                    r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; r[4] = q2.x; r[5] = q2.y;
                    // four groups of 16 positions over a rotating three-word window (keeps the unrolled body small)
                    uint32_t w0 = r[0], w1 = r[1], w2 = r[2];
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t ww[4] = {w0, w1, w2, 0u};
                        auto raw = [&](int j) -> uint32_t {   // 16 bases starting at position j of the group, j < 32
                            return (j & 15) ? __funnelshift_r(ww[j >> 4], ww[(j >> 4) + 1], 2 * (j & 15)) : ww[j >> 4];
                        };
                        const uint32_t wv16 = (uint32_t)(wv >> (16 * g)) & 0xFFFFu, wt16 = (uint32_t)(wt >> (16 * g)) & 0xFFFFu;
#pragma unroll
                        for (int c0 = 0; c0 < 16; c0 += kChunk) {
                            if (((wv16 >> c0) & ((1u << kChunk) - 1u)) == 0) continue;
                            Slot sl[kChunk];
#pragma unroll
                            for (int u = 0; u < kChunk; ++u) {
                                sl[u].code = kSlotEmpty;
                                if ((wv16 >> (c0 + u)) & 1u) {
                                    const uint32_t key = raw(c0 + u) & wmask;
                                    if (a.smap.direct) sl[u] = load_slot(a.slots + key);
                                    else if (!find_slot(a.slots, a.smap, key, &sl[u])) sl[u].code = kSlotEmpty;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < kChunk; ++u) {
                                const int j = c0 + u;
                                if (sl[u].code == kSlotEmpty) continue;
                                // the tag window: the 8 bases behind the seed (W == 16: the next register)
                                const uint32_t gc = __funnelshift_rc(raw(j), raw(j + 16), 2 * W);
                                const bool clean = (wt16 >> j) & 1u;
                                if (!(sl[u].code & kWalkBucket) || ((sl[u].tag_a | sl[u].tag_b) >> 16)) {
                                    if (!clean || !(tag_rejects(sl[u].tag_a, gc, N) && tag_rejects(sl[u].tag_b, gc, N)))
                                        push_survivor(a, tile, lp0 + 16 * g + j, sl[u].code);
                                } else {
                                    walk |= 1ull << (16 * g + j);
                                    if (!clean) dirty |= 1ull << (16 * g + j);
                                }
                            }
                        }
                        w0 = w1; w1 = w2; w2 = g == 0 ? r[3] : (g == 1 ? r[4] : r[5]);
                    }
                }
            }
            const uint32_t walk_lo = (uint32_t)walk, walk_hi = (uint32_t)(walk >> 32);
            const uint32_t dirty_lo = (uint32_t)dirty, dirty_hi = (uint32_t)(dirty >> 32);
            // ---- phase B: warp-cooperative bucket walks, lane by lane
            uint32_t lanes_todo = __ballot_sync(0xffffffffu, walk != 0);
            while (lanes_todo) {
                const int src = __ffs(lanes_todo) - 1;
                lanes_todo &= lanes_todo - 1;
                uint64_t todo = ((uint64_t)__shfl_sync(0xffffffffu, walk_hi, src) << 32) | __shfl_sync(0xffffffffu, walk_lo, src);
                const uint64_t dirt = ((uint64_t)__shfl_sync(0xffffffffu, dirty_hi, src) << 32) | __shfl_sync(0xffffffffu, dirty_lo, src);
                const uint32_t lps = ubase + (uint32_t)src * kPosPerThread;
                while (todo) {
                    constexpr int kBatch = 4;   // four positions per round: their first 32 entries are loaded together
                    int js[kBatch];
                    uint32_t es[kBatch], gs[kBatch];
                    BucketEntry bs[kBatch];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const bool on = todo != 0;
                        const int j = on ? __ffsll((long long)todo) - 1 : 0;
                        js[u] = on ? j : -1;
                        todo &= todo - 1;   // 0 stays 0
                        // key and tag window once more, straight from the planes (uniform addresses, L1 hits)
                        const int64_t gp = td.gbase + lps + j;
                        gs[u] = fetch_bits(a.p2, 2 * (gp + W), 2 * kTagBases);
                        Slot s;
                        s.code = kSlotEmpty;
                        if (on) find_slot(a.slots, a.smap, extract_key(a.p2, gp, wmask), &s);
                        es[u] = s.code & ~kWalkBucket;
                        bs[u] = on ? a.bucket[es[u] + lane] : BucketEntry{0x80000000u, 0u};   // the entry array is padded by 32
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int j = js[u] & 63;
                        const bool cj = !((dirt >> j) & 1ull);
                        BucketEntry b = bs[u];
                        for (uint32_t e = es[u];;) {
                            const uint32_t last = __ballot_sync(0xffffffffu, (b.rec_last >> 31) != 0);
                            const int n_in = last ? __ffs(last) : 32;
                            if (js[u] >= 0 && lane < n_in && (!cj || !tag_rejects(b.tag, gs[u], N)))
                                push_survivor(a, tile, lps + j, b.rec_last & 0x7FFFFFFFu);
                            if (last) break;
                            e += 32;
                            b = a.bucket[e + lane];
                        }
                    }
                }
            }
        }
    }
}

// Sampled tables (exact searches: mpcr_ctx_set_sampling): only every samp_s-th position of a contig is probed -- the
// table holds, for every record, the windows of its first primer at offsets hash_off .. hash_off + samp_s - 1, so exactly
// one probed position sees each site.  With 10^6 STS the keys (samp_s per record) outnumber anything a shared-memory
// filter can hold, so the first level is a Bloom filter in global memory (L2-resident): one 4-byte gather per probed
// position, no shared-memory carve-out, so plain loads run at the full L1 rate.  Neighbouring lanes take neighbouring
// probed positions: their plane2 / valid words come from the same few cache lines.  What passes the filter (a fraction
// of a percent) probes the open-addressed slot table and the inline tags like the other scanners; survivors carry the
// probed position and the item (record, window).
__global__ void __launch_bounds__(256, 4) sampled_scan_kernel(const ScanArgs a) {
    constexpr int kChunk = 4;                      // probes in flight per lane
    const int lane = threadIdx.x & 31;
    const uint32_t S = a.samp_s;
    const int W = a.prm.W, N = a.prm.N;
    const uint32_t wmask = wmask_of(W), vmask = wmask_bits(W);
    const uint32_t* __restrict__ p2w = reinterpret_cast<const uint32_t*>(a.p2);
    const uint32_t* __restrict__ vw = reinterpret_cast<const uint32_t*>(a.valid);
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) return;
        const TileDesc td = a.tiles[tile];
        const uint32_t first = (S - td.lstart % S) % S;       // probed positions are multiples of S in contig coordinates
        if (first >= td.nbases) continue;
        const uint32_t count = (td.nbases - first + S - 1) / S;
        for (uint32_t q0 = 0; q0 < count; q0 += 32 * kChunk) {
            uint32_t lp[kChunk], key[kChunk], word[kChunk];
            bool ok[kChunk];
#pragma unroll
            for (int u = 0; u < kChunk; ++u) {
                const uint32_t q = q0 + 32 * u + lane;
                ok[u] = q < count;
                lp[u] = first + S * (ok[u] ? q : 0u);
                const int64_t g = td.gbase + lp[u];
                const uint32_t wi = (uint32_t)(g >> 4), sh = ((uint32_t)g & 15u) * 2u;
                key[u] = __funnelshift_r(__ldg(p2w + wi), __ldg(p2w + wi + 1), sh) & wmask;
                const uint32_t vi = (uint32_t)(g >> 5), vs = (uint32_t)g & 31u;
                const uint32_t vv = __funnelshift_r(__ldg(vw + vi), __ldg(vw + vi + 1), vs);
                ok[u] = ok[u] && (vv & vmask) == vmask;       // the whole window is A/C/G/T (engine.py:483, N == 0)
                word[u] = ok[u] ? __ldg(a.bloom + bloom_word(key[u], a.bloom_shift)) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kChunk; ++u) {
                const uint32_t bits = bloom_bits(key[u]);
                if (!ok[u] || (word[u] & bits) != bits) continue;
                if (a.debug & 1) { atomicAdd(a.count, 1ull); continue; }
                Slot sl;
                if (!find_slot(a.slots, a.smap, key[u], &sl)) continue;
                const int64_t gb = td.gbase + lp[u] + W;
                const uint32_t gcodes = fetch_bits(a.p2, 2 * gb, 2 * kTagBases), gvalid = fetch_bits(a.valid, gb, kTagBases);
                if (!slot_survives(sl, gcodes, gvalid, N)) continue;
                if (a.debug & 2) atomicAdd(a.count, 1ull);
                else push_survivor(a, tile, lp[u], sl.code);
            }
        }
    }
}

// (4)(5) verifier + hit emitter: kVerifyLanes lanes (one) per survivor for the primer-1 half -- the work is a chain of
// dependent loads, so the narrower the groups the more chains a warp keeps in flight --, the whole warp for the mate
// search of whatever passes (verify_quad).
#ifndef MPCR_VERIFY_LANES
#define MPCR_VERIFY_LANES 1
#endif
// lanes per survivor in the primer-1 half; cfg3 verifier with 8 / 4 / 2 / 1: 0.183 / 0.156 / 0.139 / 0.133 ms, cfg4 0.57 /
// 0.45 / 0.39 / 0.36 ms, one eighth of cfg3 0.037 / 0.032 / 0.033 / 0.033 ms
static constexpr int kVerifyLanes = MPCR_VERIFY_LANES;

// What the primer-1 half of engine.py:507-597 leaves for the mate search of one (survivor, record) pair.
struct MateJob {
    int64_t k, p2, gcontig;
    int32_t lo, hi, l2;
    uint32_t rec, p2_word, hash_off, contig, n_rank;
    bool pass;
};

// engine.py:486-543 for one record at one seed position: primer 1, the end-of-sequence clamp, the mate window.
__device__ __forceinline__ MateJob primer1_phase(const ScanArgs& a, const TileDesc& td, uint32_t lp, uint32_t item) {
    MateJob j;
    j.pass = false;
    // sampled tables: item = record * samp_s + window, the probed position lies `window` bases behind the seed
    const uint32_t rec = a.samp_s > 1 ? item / a.samp_s : item, win = a.samp_s > 1 ? item - rec * a.samp_s : 0u;
    const RecMeta m = a.meta[rec];
    const int64_t gcontig = td.gbase - (int64_t)td.lstart, L = td.length;
    const int l1 = m.len1, l2 = m.len2;
    const int64_t k = (int64_t)td.lstart + lp - (int64_t)win - (int64_t)m.hash_off;          // engine.py:486
    if (k < 0 || k + l1 > L) return j;                                                        // :487
    if (!compare_primer(a.p4, gcontig + k, a.pwords + m.p1_word, l1, true, a.prm)) return j;  // :515
    if (a.prm.gap > 0 && earlier_block_exact(a.p4, gcontig + k, a.pwords + m.p1_word, m.hash_off, a.prm.W - a.prm.block,
                                             a.prm.block, a.prm.gap))
        return j;   // block tables: the table of an earlier block reports this site
    if (L - (k + l1) < l2) return j;                                                          // :521-524
    int64_t E = (int64_t)m.pcr_size, hi, lo;
    if (E > L - k) { E = L - k; hi = 0; }                                                     // :531-533
    else { hi = L - k - E; if (hi > a.prm.M) hi = a.prm.M; }                                  // :535
    lo = E - l1 - l2; if (lo > a.prm.M) lo = a.prm.M; if (lo < 0) lo = 0;                     // :538-540
    j.k = k; j.p2 = k + E - l2; j.gcontig = gcontig;                                          // :543
    j.lo = (int32_t)lo; j.hi = (int32_t)hi; j.l2 = l2;
    j.rec = rec; j.p2_word = m.p2_word; j.hash_off = (uint32_t)m.hash_off; j.contig = td.contig;
    j.n_rank = 2u * (uint32_t)(lo > hi ? lo : hi) + 1u;
    j.pass = true;
    return j;
}

// The 8-base pre-check of up to 32 consecutive mate positions (mpcr_core.cuh: mate_precheck32), out of line: the
// unrolled block is large and the verifier lives on 64 registers.
__device__ __noinline__ uint32_t mate_block(const uint64_t* __restrict__ p4, int64_t gb, uint32_t m, const PrimerView& v,
                                            const SearchParams& prm) {
    return mate_precheck32(p4, gb, m, v, prm);
}

// engine.py:545-593: the mate search of one job by a team of lanes.  The window's positions p2 - lo .. p2 + hi are dealt
// out in CONSECUTIVE stretches, one per lane, so that a lane can check its stretch in rolling form (blocks of 32
// positions: six words of plane4 loaded once, one funnel shift per position) and runs the full compare only where the
// first eight bases pass; a hit's rank is the reference's order of deltas (0, -1, +1, -2, ...), which the sort restores.
// Primers longer than 32 bases keep the position-by-position compare.
__device__ __forceinline__ void mate_phase(const ScanArgs& a, const MateJob& j, uint32_t lane_in_team, uint32_t team) {
    const uint64_t* q2 = a.pwords + j.p2_word;
    const PrimerView v2 = make_primer_view(q2, j.l2, false, a.prm);   // hoisted out of the delta loop
    const HitEmitter emit{a.hits, a.capacity, a.count, j.contig, j.rec, j.hash_off};
    const uint32_t n_off = (uint32_t)(j.lo + j.hi) + 1u;               // positions of the window
    const int64_t q_first = j.p2 - j.lo;
    const uint32_t per = (n_off + team - 1u) / team;
    const uint32_t t0 = lane_in_team * per;
    if (t0 >= n_off) return;
    const uint32_t cnt = min(per, n_off - t0);
    auto rank_of = [&](int64_t q) -> uint32_t {
        const int64_t d = q - j.p2;
        return d == 0 ? 0u : (d < 0 ? (uint32_t)(-2 * d - 1) : (uint32_t)(2 * d));
    };
    if (v2.nw) {
        for (uint32_t c0 = 0; c0 < cnt; c0 += 32u) {
            const int64_t qb = q_first + t0 + c0;
            uint32_t cand = mate_block(a.p4, j.gcontig + qb, min(32u, cnt - c0), v2, a.prm);
            while (cand) {
                const int64_t q = qb + (__ffs(cand) - 1);
                cand &= cand - 1;
                if (compare_view(a.p4, j.gcontig + q, v2, a.prm)) emit(j.k, q + j.l2 - 1, rank_of(q));
            }
        }
    } else {
        for (uint32_t t = 0; t < cnt; ++t) {
            const int64_t q = q_first + t0 + t;
            if (compare_primer(a.p4, j.gcontig + q, q2, j.l2, false, a.prm)) emit(j.k, q + j.l2 - 1, rank_of(q));
        }
    }
}

// One survivor per lane group of the warp (`have`: this group has one).  The groups walk their survivors' records side
// by side -- one record for an ordinary survivor, the bucket entries their tags do not rule out for a seed shared by
// several records -- and compare primer 1 on their own (the lanes of a group read the same addresses: broadcast loads;
// one chain of dependent loads per group).  The mate searches of the records
// that pass are then done by the WHOLE warp, one job after the other: a +-50 window is 13 rounds for a group of 8 lanes
// but 4 for the warp, and of the survivors a warp looks at side by side most do not get this far -- their groups used
// to sit through the others' rounds idle.  (cfg3 verifier 0.204 -> 0.133 ms, cfg4 0.76 -> 0.36, cfg5 5.2 -> 4.1.)
__device__ __forceinline__ void verify_quad(const ScanArgs& a, const Survivor* __restrict__ mine, bool have) {
    const uint32_t lane = threadIdx.x & 31;
    TileDesc td = {};
    uint32_t lp = 0, cur = 0, gcodes = 0;
    bool walk = false, clean = false, more = have;
    if (have) {
        const Survivor sv = *mine;
        td = a.tiles[sv.tile];
        lp = sv.lp;
        walk = (sv.code & kWalkBucket) != 0;
        cur = sv.code & ~kWalkBucket;          // the record, or the first bucket entry
        if (walk) {   // a seed shared by several records: bucket order, each entry behind its own tag
            const int64_t gb = td.gbase + sv.lp + a.prm.W + a.prm.gap;
            gcodes = fetch_bits(a.p2, 2 * gb, 2 * kTagBases);
            clean = tag_window_clean(fetch_bits(a.valid, gb, kTagBases));
        }
    }
    while (__any_sync(0xffffffffu, more)) {
        MateJob job;
        job.pass = false;
        if (more) {
            uint32_t item = cur;
            bool has = true;
            if (!walk) {
                more = false;
            } else {
                for (;;) {   // the next entry its tag does not rule out
                    const BucketEntry b = a.bucket[cur++];
                    const bool last = (b.rec_last >> 31) != 0;
                    has = !clean || !tag_rejects(b.tag, gcodes, a.prm.N);
                    item = b.rec_last & 0x7FFFFFFFu;
                    if (last) more = false;
                    if (has || last) break;
                }
            }
            if (has) job = primer1_phase(a, td, lp, item);
        }
        // the first lane of every group that has a job
        uint32_t todo = __ballot_sync(0xffffffffu, job.pass && (lane % kVerifyLanes) == 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            MateJob t;
            t.k = __shfl_sync(0xffffffffu, job.k, src);
            t.p2 = __shfl_sync(0xffffffffu, job.p2, src);
            t.gcontig = __shfl_sync(0xffffffffu, job.gcontig, src);
            t.lo = __shfl_sync(0xffffffffu, job.lo, src);
            t.hi = __shfl_sync(0xffffffffu, job.hi, src);
            t.l2 = __shfl_sync(0xffffffffu, job.l2, src);
            t.rec = __shfl_sync(0xffffffffu, job.rec, src);
            t.p2_word = __shfl_sync(0xffffffffu, job.p2_word, src);
            t.hash_off = __shfl_sync(0xffffffffu, job.hash_off, src);
            t.contig = __shfl_sync(0xffffffffu, job.contig, src);
            t.n_rank = __shfl_sync(0xffffffffu, job.n_rank, src);
            t.pass = true;
            mate_phase(a, t, lane, 32u);
        }
    }
}

__device__ __forceinline__ void verify_body(const ScanArgs& a) {
    constexpr int kGroups = 32 / kVerifyLanes;
    constexpr uint32_t kGrab = kGroups;       // survivors a warp takes per cursor atomic (one per lane group)
    const int lane = threadIdx.x & 31, group = lane / kVerifyLanes;
    const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // every warp starts at its own sub-list and moves on when that one is drained, until it has seen them all;
    // 32 sub-lists are looked at per round (one per lane) so that drained ones cost nothing
    for (uint32_t round = 0; round < kSurvLists / 32; ++round) {
      const uint32_t my_list = (warp_id + 32 * round + lane) & (kSurvLists - 1u);
      const uint32_t* my_ctl = a.surv_ctl + my_list * kSurvCtlStride;
      uint32_t open = __ballot_sync(0xffffffffu, __ldcg(my_ctl + 1) < min(__ldcg(my_ctl), a.surv_cap));
      while (open) {
        const int l = __ffs(open) - 1;
        open &= open - 1;
        const uint32_t list = (warp_id + 32 * round + l) & (kSurvLists - 1u);
        uint32_t* ctl = a.surv_ctl + list * kSurvCtlStride;
        const uint32_t n = min(__ldcg(ctl), a.surv_cap);
        const Survivor* surv = a.surv + (size_t)list * a.surv_cap;
        for (;;) {
            uint32_t first = 0;
            if (lane == 0) first = __ldcg(ctl + 1) < n ? atomicAdd(ctl + 1, kGrab) : n;   // look before touching the line
            first = __shfl_sync(0xffffffffu, first, 0);
            if (first >= n) break;
            const uint32_t idx = first + group;
            verify_quad(a, surv + idx, idx < n);
        }
      }
    }
}

#ifndef MPCR_VERIFY_CTAS_PER_SM
#define MPCR_VERIFY_CTAS_PER_SM 4
#endif
#ifndef MPCR_VERIFY_STATIC_DEPTH
#define MPCR_VERIFY_STATIC_DEPTH 64
#endif
// 64 registers x 256 threads: the verifier is a chain of dependent loads per survivor, so a fourth resident CTA per SM
// is worth the handful of spilled registers (whole genome 0.215 -> 0.202 ms, 1/8 shard 49 -> 45.5 us against three CTAs)
static constexpr int kVerifyCtasPerSm = MPCR_VERIFY_CTAS_PER_SM;
static constexpr uint32_t kStaticVerifyDepth = MPCR_VERIFY_STATIC_DEPTH;  // survivors per lane group dealt out statically
__global__ void __launch_bounds__(256, kVerifyCtasPerSm) verify_kernel(const ScanArgs a) {
    // Survivor lists of up to kStaticVerifyDepth entries per lane group are dealt out STATICALLY: every CTA sums the 256
    // sub-list counts itself (one load per thread, a block scan) and its lane groups stride over the concatenated lists
    // -- no cursor atomics, no polling of control lines that other warps are draining.  The grid is what fits the GPU at
    // once (kVerifyCtasPerSm per SM), so no CTA starts behind another.  Longer lists keep the dynamic schedule of
    // verify_body, which balances survivors of very different cost (mate windows, bucket walks).
    static_assert(kSurvLists == 256, "one sub-list per thread of the CTA");
    __shared__ uint32_t pre[kSurvLists + 1];
    __shared__ uint32_t wsum[8];
    {
        const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
        const uint32_t cnt = min(__ldcg(a.surv_ctl + tid * kSurvCtlStride), a.surv_cap);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < wid; ++w) base += wsum[w];
        pre[tid] = base + incl - cnt;
        if (tid == 255) pre[256] = base + incl;
        __syncthreads();
    }
    const uint32_t total = pre[kSurvLists];
    constexpr uint32_t kGroupsPerWarp = 32 / kVerifyLanes;
    // (the switch-over point is counted in survivors per warp, four to the unit, whatever the group width)
    if (total / (gridDim.x * (blockDim.x >> 5) * 4u) < kStaticVerifyDepth) {
        // survivor g of the concatenated sub-lists goes to warp g % n_warps; a warp takes its survivors as many at a time
        // as it has lane groups.  (Dealing consecutive survivors to the groups of ONE warp left most warps of a short list
        // -- a rank of an 8-GPU run -- without work once the groups became narrow.)
        const int lane = threadIdx.x & 31, group = lane / kVerifyLanes;
        const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
        const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        for (uint32_t slot0 = 0; (uint64_t)slot0 * n_warps + warp_id < total; slot0 += kGroupsPerWarp) {   // warp-uniform
            const uint64_t g64 = (uint64_t)(slot0 + group) * n_warps + warp_id;
            const bool have = g64 < total;
            const uint32_t g = have ? (uint32_t)g64 : 0u;
            uint32_t lo = 0, hi = kSurvLists;           // the sub-list that holds global index g
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (pre[mid] <= g) lo = mid; else hi = mid;
            }
            verify_quad(a, a.surv + ((size_t)lo * a.surv_cap + (g - pre[lo])), have);
        }
    } else {
        verify_body(a);
    }
    // The last CTA to finish zeroes the scanner's control words (survivor counters and cursors, tile counter), so the
    // next scan needs no memset launches in front of it.  tile_counter[4] counts finished CTAs.
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(a.tile_counter + 4, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    for (uint32_t i = threadIdx.x; i < kSurvLists; i += blockDim.x) {
        a.surv_ctl[i * kSurvCtlStride] = 0u;
        a.surv_ctl[i * kSurvCtlStride + 1] = 0u;
    }
    if (threadIdx.x == 0) {
        a.tile_counter[0] = 0u;
        a.tile_counter[4] = 0u;
    }
}

// After the radix passes over (pos1, contig): hits that share contig and pos1 (duplicate STS lines, several deltas of
// one primer-1 site) are put into the reference's discovery order (hash_off, rec, rank).  Runs of up to 32 hits are
// insertion-sorted by the thread that finds their start; longer ones (identical STS lines on a repeat, or a mate
// window inside an N-run in IUPAC mode: every offset matches) go to a queue for order_long_runs, because a lone thread
// sorting a hundred 24-byte records in global memory is a chain of thousands of dependent L2 round trips.
struct LongRun {
    uint32_t first, len;
};
static constexpr uint32_t kLongRunQueue = 1u << 16;   // queue entries (a full queue sorts in place, below)
static constexpr uint32_t kLongRunMax = 1024;         // hits one CTA sorts in shared memory

__device__ __forceinline__ bool tie_less(const mpcr_hit& a, const mpcr_hit& b) {
    if (a.hash_off != b.hash_off) return a.hash_off < b.hash_off;
    if (a.rec != b.rec) return a.rec < b.rec;
    return a.rank < b.rank;
}
// heap sort of one run by one thread, O(m log m): the fallback for runs no CTA can hold
__device__ __noinline__ void tie_heap_sort(mpcr_hit* h, uint64_t m) {
    auto sift = [&](uint64_t root, uint64_t end) {
        const mpcr_hit x = h[root];
        for (;;) {
            uint64_t child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && tie_less(h[child], h[child + 1])) ++child;
            if (!tie_less(x, h[child])) break;
            h[root] = h[child];
            root = child;
        }
        h[root] = x;
    };
    for (uint64_t k = m / 2; k-- > 0;) sift(k, m);
    for (uint64_t end = m - 1; end > 0; --end) {
        const mpcr_hit t = h[0];
        h[0] = h[end];
        h[end] = t;
        sift(0, end);
    }
}

__global__ void __launch_bounds__(256) order_ties(const mpcr_hit* __restrict__ in, mpcr_hit* __restrict__ hits,
                                                  uint64_t n_host, const unsigned long long* d_n, uint32_t skip_upto,
                                                  const uint32_t* __restrict__ skip_off, LongRun* __restrict__ queue,
                                                  uint32_t* __restrict__ queue_ctl /* [0] entries, [1] cursor */) {
    // `in` is where the radix passes left the records (the hit buffer itself, or the sort's scratch buffer after an
    // odd number of passes): the thread that owns a run moves it home first, so no separate copy pass is needed
    const uint64_t n = sort_count(d_n, n_host);
    if (n <= skip_upto && !(skip_off && *skip_off)) return;   // the bucket sort ordered this list completely
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mpcr_hit h0 = in[i];
    const uint32_t c = h0.contig, p = h0.pos1;
    if (i > 0 && in[i - 1].contig == c && in[i - 1].pos1 == p) return;   // not the start of a run
    uint64_t e = i + 1;
    while (e < n && in[e].contig == c && in[e].pos1 == p) ++e;
    const uint64_t m = e - i;
    if (in != hits) {
        hits[i] = h0;
        for (uint64_t k = i + 1; k < e; ++k) hits[k] = in[k];
    }
    if (m == 1) return;
    if (m <= 32) {
        for (uint64_t k = i + 1; k < e; ++k) {
            const mpcr_hit x = hits[k];
            uint64_t j = k;
            while (j > i && tie_less(x, hits[j - 1])) { hits[j] = hits[j - 1]; --j; }
            hits[j] = x;
        }
        return;
    }
    if (m <= kLongRunMax && i < 0xFFFFFFFFull) {
        const uint32_t q = atomicAdd(queue_ctl, 1u);
        if (q < kLongRunQueue) { queue[q] = LongRun{(uint32_t)i, (uint32_t)m}; return; }
    }
    tie_heap_sort(hits + i, m);
}

// One CTA per queued run: the records and their 61-bit tie keys (hash_off, rec, rank) go to shared memory, a bitonic
// network orders (key, index) pairs, the records are written back in that order.
__global__ void __launch_bounds__(256) order_long_runs(mpcr_hit* __restrict__ hits, const LongRun* __restrict__ queue,
                                                       uint32_t* __restrict__ queue_ctl /* [0] entries [1] cursor [2] CTAs done */) {
    __shared__ mpcr_hit rec[kLongRunMax];
    __shared__ unsigned long long key[kLongRunMax];
    __shared__ uint16_t idx[kLongRunMax];
    __shared__ uint32_t s_q;
    const uint32_t n_q = min(queue_ctl[0], kLongRunQueue);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_q = n_q ? atomicAdd(queue_ctl + 1, 1u) : 0u;
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= n_q) break;
        const LongRun r = queue[q];
        uint32_t np = 64;
        while (np < r.len) np <<= 1;
        for (uint32_t k = threadIdx.x; k < np; k += blockDim.x) {
            unsigned long long kk = ~0ull;   // padding sorts to the end
            if (k < r.len) {
                const mpcr_hit h = hits[r.first + k];
                rec[k] = h;
                kk = ((unsigned long long)h.hash_off << 45) | ((unsigned long long)h.rec << 15) | (unsigned long long)h.rank;
            }
            key[k] = kk;
            idx[k] = (uint16_t)k;
        }
        __syncthreads();
        for (uint32_t size = 2; size <= np; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = threadIdx.x; t < np / 2; t += blockDim.x) {
                    const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const unsigned long long a = key[lo], b = key[hi];
                    if ((a > b) == up) {
                        key[lo] = b; key[hi] = a;
                        const uint16_t ia = idx[lo]; idx[lo] = idx[hi]; idx[hi] = ia;
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t k = threadIdx.x; k < r.len; k += blockDim.x) hits[r.first + k] = rec[idx[k]];
    }
    // the last CTA to leave zeroes the control words for the next sort (every CTA has read n_q by then)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(queue_ctl + 2, 1u) == gridDim.x - 1) {
            queue_ctl[0] = 0; queue_ctl[1] = 0;
            __threadfence();
            queue_ctl[2] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int mpcr_abi_version(void) { return MPCR_ABI_VERSION; }
const char* mpcr_last_error(void) { return g_err; }

void mpcr_ctx_destroy(mpcr_ctx* c);
int mpcr_ctx_create(int device, const mpcr_params* p, mpcr_ctx** out) {
    if (!p || !out) return fail(MPCR_EINVAL, "null argument");
    // core/engine.py:80-97
    if (p->wordsize < 3 || p->wordsize > 16) return fail(MPCR_EINVAL, "Word size must be between 3 and 16");
    if (p->mismatches < 0 || p->mismatches > 10) return fail(MPCR_EINVAL, "Number of mismatches must be between 0 and 10");
    if (p->margin < 0 || p->margin > 10000) return fail(MPCR_EINVAL, "Margin must be between 0 and 10000");
    if (p->three_prime_match < 0) return fail(MPCR_EINVAL, "Three prime match must be at least 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MPCR_ECUDA, "no CUDA device available (%s); merpcr_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(MPCR_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    DeviceGuard guard_(device);
    if (!guard_.ok) return fail(MPCR_ECUDA, "cudaSetDevice(%d) failed", device);
    mpcr_ctx* c = new mpcr_ctx();
    c->device = device;
    c->prm = *p;
    // tuning / test switches are read once here, never on the scan path
    c->env_debug = getenv("MPCR_DEBUG") ? atoi(getenv("MPCR_DEBUG")) : 0;
    c->env_surv_cap = getenv("MPCR_SURVIVOR_CAP") ? atol(getenv("MPCR_SURVIVOR_CAP")) : -1;
    cudaError_t err = cudaSuccess;
    auto step = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    step(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    step(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    step(cudaMalloc(&c->d_tile_counter, 256));
    step(cudaMalloc(&c->d_surv_ctl, (size_t)kSurvLists * kSurvCtlStride * 4));   // kSurvLists control lines
    step(cudaMalloc(&c->d_lut, 256));
    if (err == cudaSuccess) step(cudaMemset(c->d_tile_counter, 0, 256));
    if (err == cudaSuccess) step(cudaMemset(c->d_surv_ctl, 0, (size_t)kSurvLists * kSurvCtlStride * 4));
    for (int sl = 0; sl < MPCR_MAX_SLOTS; ++sl)
        for (int k = 0; k < 3; ++k) step(cudaEventCreate(&c->evs[sl][k]));
    if (err != cudaSuccess) {
        mpcr_ctx_destroy(c);   // frees whatever was allocated
        return fail(MPCR_ECUDA, "context setup failed: %s", cudaGetErrorString(err));
    }
    *out = c;
    return MPCR_OK;
}

static void free_table(mpcr_ctx* c) {
    cudaFree(c->d_meta); cudaFree(c->d_pwords); cudaFree(c->d_slots); cudaFree(c->d_bucket); cudaFree(c->d_filter);
    cudaFree(c->d_bloom);
    c->d_meta = nullptr; c->d_pwords = nullptr; c->d_slots = nullptr; c->d_bucket = nullptr; c->d_filter = nullptr;
    c->d_bloom = nullptr;
    c->table_ready = false;
}

void mpcr_ctx_destroy(mpcr_ctx* c) {
    if (!c) return;
    DeviceGuard guard_(c->device);
    free_table(c);
    cudaFree(c->d_tiles); cudaFree(c->d_tile_counter); cudaFree(c->d_sort_tmp); cudaFree(c->d_long_runs); cudaFree(c->d_counts); cudaFree(c->d_lut);
    for (int sl = 0; sl < MPCR_MAX_SLOTS; ++sl)
        for (int k = 0; k < 3; ++k)
            if (c->evs[sl][k]) cudaEventDestroy(c->evs[sl][k]);
    cudaFree(c->d_surv);
    cudaFree(c->d_surv_ctl);
    cudaFree(c->d_contig_g);
    cudaFree(c->d_bsort);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    delete c;
}

int mpcr_ctx_set_seed_extension(mpcr_ctx* c, int w_ext, int which) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (which < 0 || which > 2) return fail(MPCR_EINVAL, "which must be 0, 1 or 2");
    if (which != 0) {
        if (c->prm.mismatches != 0 || c->prm.iupac_mode != 0)
            return fail(MPCR_EINVAL, "seed extension needs an exact search (mismatches 0, no IUPAC mode)");
        if (w_ext <= c->prm.wordsize || w_ext > 16) return fail(MPCR_EINVAL, "extended word must be in (wordsize, 16]");
    }
    c->ext_w = which ? w_ext : 0;
    c->ext_which = which;
    c->ext_block = c->ext_gap = c->ext_span = 0;
    free_table(c);
    return MPCR_OK;
}
int mpcr_ctx_set_seed_blocks(mpcr_ctx* c, int block, int n_blocks, int which) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (which < 0 || (which >= 2 && which - 2 >= n_blocks)) return fail(MPCR_EINVAL, "which must be 0, 1 or 2 + block index");
    if (which != 0) {
        if (c->prm.iupac_mode != 0) return fail(MPCR_EINVAL, "block tables need letter-identity compares (no IUPAC mode)");
        if (block < 1 || n_blocks < 1) return fail(MPCR_EINVAL, "block and n_blocks must be positive");
        if (n_blocks <= c->prm.mismatches)
            return fail(MPCR_EINVAL, "block tables need more blocks than mismatches (%d blocks, -N %d)", n_blocks, c->prm.mismatches);
        if (c->prm.wordsize + n_blocks * block > 16)
            return fail(MPCR_EINVAL, "wordsize + n_blocks * block must not exceed 16 letters");
    }
    c->ext_w = which ? c->prm.wordsize + block : 0;
    c->ext_which = which >= 2 ? 2 : which;
    c->ext_block = which ? block : 0;
    c->ext_gap = which >= 2 ? (which - 2) * block : 0;
    c->ext_span = which ? c->prm.wordsize + n_blocks * block : 0;
    free_table(c);
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
    c->samp_w = role ? w_samp : 0;
    c->samp_s = role ? stride : 0;
    c->samp_role = role;
    free_table(c);
    return MPCR_OK;
}
int mpcr_ctx_set_table_part(mpcr_ctx* c, uint32_t part, uint32_t parts) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (parts == 0) parts = 1;
    if (part >= parts) return fail(MPCR_EINVAL, "part must be < parts");
    c->part = part;
    c->parts = parts;
    free_table(c);
    return MPCR_OK;
}
int mpcr_ctx_set_append(mpcr_ctx* c, int on) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    c->append = on ? 1 : 0;
    return MPCR_OK;
}
int mpcr_ctx_set_true_strands(mpcr_ctx* c, int on) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    c->true_strands = on ? 1 : 0;
    free_table(c);
    return MPCR_OK;
}
int mpcr_ctx_sm_count(const mpcr_ctx* c) { return c ? c->sm_count : 0; }
uint64_t mpcr_launch_count(const mpcr_ctx* c) { return c ? c->launches : 0; }
uint32_t mpcr_table_items(const mpcr_ctx* c) { return c && c->table_ready ? c->n_valid : 0; }
uint64_t mpcr_fasta_workspace_bytes(uint64_t n, uint32_t max_records) {
    const uint64_t n_blk = (n + kFastaBlock - 1) / kFastaBlock, m = max_records ? max_records : 1;
    return (n_blk + 1) * 8 + ((n_blk + 1) & ~1ull) * 4 + 8 + m * (sizeof(FastaHeader) + 32) + 64;
}

int mpcr_pack_sequence(mpcr_ctx* c, const uint8_t* d_ascii, uint64_t n, uint64_t dst_base, uint64_t plane_origin,
                       void* d_plane2, void* d_plane4, void* d_valid, const uint8_t* h_lut, void* stream) {
    if (!c || !d_plane2 || !d_plane4 || !d_valid || !h_lut) return fail(MPCR_EINVAL, "null argument");
    if ((dst_base & 63u) || (plane_origin & 127u) || dst_base < plane_origin)
        return fail(MPCR_EINVAL, "dst_base must be a multiple of 64 and >= plane_origin (multiple of 128)");
    if (n == 0) return MPCR_OK;
    if (!d_ascii) return fail(MPCR_EINVAL, "null sequence pointer");
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    // the 256-byte LUT goes up only when it changes: a pageable upload per call queues on the H2D copy engine behind
    // the 64 MiB pieces of a streaming ingest and stalls the pack behind them
    if (!c->lut_valid || memcmp(c->lut_host, h_lut, 256) != 0) {
        memcpy(c->lut_host, h_lut, 256);
        CU(cudaMemcpyAsync(c->d_lut, c->lut_host, 256, cudaMemcpyHostToDevice, st));
        c->lut_valid = true;
    }
    const uint64_t strips = (n + 63) / 64;
    const uint32_t blocks = (uint32_t)((strips + 255) / 256);
    pack_kernel<<<blocks, 256, 0, st>>>(d_ascii, n, dst_base - plane_origin, (uint64_t*)d_plane2, (uint64_t*)d_plane4,
                                        (uint64_t*)d_valid, c->d_lut);
    c->launches++;
    CU(cudaGetLastError());
    return MPCR_OK;
}

int mpcr_derive_planes(mpcr_ctx* c, uint64_t n, uint64_t dst_base, uint64_t plane_origin, const void* d_plane4,
                       void* d_plane2, void* d_valid, void* stream) {
    if (!c || !d_plane2 || !d_plane4 || !d_valid) return fail(MPCR_EINVAL, "null argument");
    if ((dst_base & 63u) || (plane_origin & 127u) || dst_base < plane_origin)
        return fail(MPCR_EINVAL, "dst_base must be a multiple of 64 and >= plane_origin (multiple of 128)");
    if (n == 0) return MPCR_OK;
    GUARD(c);
    const uint64_t strips = (n + 63) / 64;
    derive_planes_kernel<<<(uint32_t)((strips + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint64_t*)d_plane4, n, dst_base - plane_origin, (uint64_t*)d_plane2, (uint64_t*)d_valid);
    c->launches++;
    CU(cudaGetLastError());
    return MPCR_OK;
}

// ---- host-side text helpers (mpcr_hostio.h) ------------------------------------------------------------------
int mpcr_sts_parse(const uint8_t* text, uint64_t n, int32_t wordsize, int32_t default_pcr_size, mpcr_sts_line* lines,
                   uint32_t max_lines, uint32_t* n_lines, uint32_t* bad_line, uint32_t* short_primers, uint32_t* flags) {
    if (!n_lines || !bad_line || !short_primers || !flags || (n && !text) || (max_lines && !lines))
        return fail(MPCR_EINVAL, "null argument");
    const int rc = sts_parse_impl(text, n, wordsize, default_pcr_size, lines, max_lines, n_lines, bad_line, short_primers, flags);
    if (rc == MPCR_EINVAL) return fail(MPCR_EINVAL, "STS text of 4 GiB or more is not supported");
    return rc;
}
int mpcr_sts_blob(const uint8_t* text, const mpcr_sts_line* lines, uint32_t n_lines, uint8_t* blob, uint64_t* off) {
    if (!off || (n_lines && (!text || !lines || !blob))) return fail(MPCR_EINVAL, "null argument");
    sts_blob_impl(text, lines, n_lines, blob, off);
    return MPCR_OK;
}
uint64_t mpcr_format_hits(const mpcr_hit* hits, uint64_t n, const uint8_t* text, const mpcr_sts_line* lines,
                          const uint8_t* labels, const uint64_t* label_off, uint8_t* out, uint64_t out_cap) {
    return format_hits_impl(hits, n, text, lines, labels, label_off, out, out_cap);
}

// ---- device-side FASTA text ingest (io/fasta.py:43-66) -------------------------------------------------------
int mpcr_fasta_index_ex(mpcr_ctx* c, uint8_t* d_text, uint64_t n, uint32_t mode, mpcr_fasta_record* h_records,
                        uint32_t max_records, uint32_t* n_records, uint32_t* flags, void* d_ws, uint64_t ws_bytes,
                        void* stream) {
    if (!c || !n_records || !flags) return fail(MPCR_EINVAL, "null argument");
    if (n && !d_text) return fail(MPCR_EINVAL, "null text pointer");
    if (max_records && !h_records) return fail(MPCR_EINVAL, "null record buffer");
    *n_records = 0;
    *flags = 0;
    if (n == 0) return MPCR_OK;
    if (mpcr_fasta_workspace_bytes(n, max_records) > ws_bytes || !d_ws)
        return fail(MPCR_EINVAL, "workspace too small (need %llu bytes)", (unsigned long long)mpcr_fasta_workspace_bytes(n, max_records));
    const bool skip_first_line = mode & 1u, keep_prologue = mode & 2u;
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    // workspace: [block offsets u64 (n_blk+1)] [block counts u32 n_blk] [headers max_records] [positions/offsets 4*max] [ctr]
    const uint64_t n_blk = (n + kFastaBlock - 1) / kFastaBlock;
    uint8_t* w = (uint8_t*)d_ws;
    uint64_t* d_off = (uint64_t*)w;                 w += (n_blk + 1) * 8;
    uint32_t* d_cnt = (uint32_t*)w;                 w += ((n_blk + 1) & ~1ull) * 4 + 8;
    FastaHeader* d_hdr = (FastaHeader*)w;           w += (size_t)(max_records ? max_records : 1) * sizeof(FastaHeader);
    uint64_t* d_pos = (uint64_t*)w;                 w += (size_t)(max_records ? max_records : 1) * 16;
    uint64_t* d_posoff = (uint64_t*)w;              w += (size_t)(max_records ? max_records : 1) * 16;
    uint32_t* d_ctr = (uint32_t*)w;
    // a slice that starts inside a line: that line (whatever it is a part of) is dropped first, so that header
    // detection only ever sees whole lines
    uint64_t text_begin = 0;
    if (skip_first_line) {
        CU(cudaMemsetAsync(d_ctr, 0xFF, 4, st));
        fasta_first_line_end<<<1, 256, 0, st>>>(d_text, n, 1u << 20, d_ctr);
        c->launches++;
        uint32_t fe = 0;
        CU(cudaMemcpyAsync(&fe, d_ctr, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (fe == 0xFFFFFFFFu) { *flags = 2u; return MPCR_OK; }   // no line end in the first MiB: the caller widens its slice
        text_begin = fe;
        CU(cudaMemsetAsync(d_text, '\n', (size_t)text_begin, st));
    }
    CU(cudaMemsetAsync(d_ctr, 0, 8, st));
    // one pass: headers, non-ASCII flag and the kept letters per 4096-byte block (header lines and the text in front
    // of the first header are still counted; blanking them below takes them out again)
    fasta_classify<<<(uint32_t)((n_blk + kFastaWarps - 1) / kFastaWarps), kFastaThreads, 0, st>>>(d_text, n, d_hdr, max_records, d_ctr, d_ctr + 1, d_cnt);
    c->launches++;
    CU(cudaGetLastError());
    uint32_t ctr[2] = {0, 0};
    CU(cudaMemcpyAsync(ctr, d_ctr, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint32_t lead = keep_prologue ? 1u : 0u;   // record 0 = the sequence in front of the first header
    *n_records = ctr[0] + lead;
    *flags = ctr[1];
    if (ctr[1] & 1u) return MPCR_OK;                      // non-ASCII: the host applies the locale rules itself
    if (ctr[0] + lead > max_records) return MPCR_EOVERFLOW;   // *n_records holds the required capacity
    const uint32_t nh = ctr[0];
    if (nh == 0 && !keep_prologue) { *n_records = 0; return MPCR_OK; }   // no header: everything is "before the first header"
    std::vector<FastaHeader> hdr(nh);
    if (nh) {
        CU(cudaMemcpy(hdr.data(), d_hdr, (size_t)nh * sizeof(FastaHeader), cudaMemcpyDeviceToHost));
        std::sort(hdr.begin(), hdr.end(), [](const FastaHeader& a, const FastaHeader& b) { return a.begin < b.begin; });
        CU(cudaMemcpyAsync(d_hdr, hdr.data(), (size_t)nh * sizeof(FastaHeader), cudaMemcpyHostToDevice, st));
        if (!keep_prologue && hdr[0].begin > text_begin) {   // data before the first header is discarded
            const uint64_t len = hdr[0].begin - text_begin;
            fasta_blank_range<<<(uint32_t)((len + 4095) / 4096), 256, 0, st>>>(d_text, text_begin, hdr[0].begin, d_cnt);
            c->launches++;
        }
        fasta_blank_headers<<<(nh + 127) / 128, 128, 0, st>>>(d_text, d_hdr, nh, d_cnt);
        c->launches++;
    }
    {
        const uint32_t n_chunks = (uint32_t)((n_blk + kScanChunk - 1) / kScanChunk);
        if (n_chunks > 1024) return fail(MPCR_EINVAL, "FASTA text of more than 64 GiB per call is not supported");
        uint64_t* d_chunk = d_posoff;   // scratch: the record-offset area is not in use yet (>= 1024 u64 checked below)
        if ((size_t)(max_records ? max_records : 1) * 16 < 1024 * 8 + 8) {
            // tiny record capacity: fall back to a private scratch
            CU(cudaMalloc(&d_chunk, 1025 * 8));
        }
        fasta_scan_reduce<<<n_chunks, 1024, 0, st>>>(d_cnt, n_blk, d_chunk);
        fasta_scan_top<<<1, 1024, 0, st>>>(d_chunk, n_chunks, d_off + n_blk);
        fasta_scan_down<<<n_chunks, 1024, 0, st>>>(d_cnt, n_blk, d_chunk, d_off);
        c->launches += 3;
        if (d_chunk != d_posoff) {
            CU(cudaStreamSynchronize(st));
            cudaFree(d_chunk);
        }
    }
    // kept-byte offsets at every record boundary: end of header r, start of header r+1 (or n)
    const uint32_t nr = nh + lead;
    std::vector<uint64_t> pos(2 * (size_t)nr);
    if (lead) {
        pos[0] = text_begin;
        pos[1] = nh ? hdr[0].begin : n;
    }
    for (uint32_t r = 0; r < nh; ++r) {
        pos[2 * (r + lead)] = hdr[r].end;
        pos[2 * (r + lead) + 1] = r + 1 < nh ? hdr[r + 1].begin : n;
    }
    CU(cudaMemcpyAsync(d_pos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice, st));
    fasta_offsets_at<<<(2 * nr + 3) / 4, 128, 0, st>>>(d_text, n, d_off, d_pos, 2 * nr, d_posoff);
    c->launches++;
    CU(cudaGetLastError());
    std::vector<uint64_t> po(2 * (size_t)nr);
    CU(cudaMemcpyAsync(po.data(), d_posoff, po.size() * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (uint32_t r = 0; r < nr; ++r) {
        const bool pseudo = lead && r == 0;
        h_records[r].header_begin = pseudo ? text_begin : hdr[r - lead].begin;
        h_records[r].header_end = pseudo ? text_begin : hdr[r - lead].end;
        h_records[r].seq_offset = po[2 * r];
        h_records[r].seq_length = po[2 * r + 1] - po[2 * r];
    }
    return MPCR_OK;
}

int mpcr_fasta_index(mpcr_ctx* c, uint8_t* d_text, uint64_t n, mpcr_fasta_record* h_records, uint32_t max_records,
                     uint32_t* n_records, uint32_t* flags, void* d_ws, uint64_t ws_bytes, void* stream) {
    return mpcr_fasta_index_ex(c, d_text, n, 0u, h_records, max_records, n_records, flags, d_ws, ws_bytes, stream);
}

int mpcr_fasta_offsets_at(mpcr_ctx* c, const uint8_t* d_text, uint64_t n, const void* d_ws, const uint64_t* h_pos,
                          uint32_t n_pos, uint64_t* h_out, void* stream) {
    if (!c || (n_pos && (!h_pos || !h_out))) return fail(MPCR_EINVAL, "null argument");
    if (n_pos == 0) return MPCR_OK;
    if (!d_text || !d_ws) return fail(MPCR_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    uint64_t* d_io = nullptr;
    CU(cudaMalloc(&d_io, (size_t)n_pos * 16));
    cudaError_t e = cudaMemcpyAsync(d_io, h_pos, (size_t)n_pos * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        fasta_offsets_at<<<(n_pos + 3) / 4, 128, 0, st>>>(d_text, n, (const uint64_t*)d_ws, d_io, n_pos, d_io + n_pos);
        c->launches++;
        e = cudaMemcpyAsync(h_out, d_io + n_pos, (size_t)n_pos * 8, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_io);
    if (e != cudaSuccess) return fail(MPCR_ECUDA, "mpcr_fasta_offsets_at: %s", cudaGetErrorString(e));
    return MPCR_OK;
}

int mpcr_fasta_compact(mpcr_ctx* c, const uint8_t* d_text, uint64_t n, const void* d_ws, uint8_t* d_seq, void* stream) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (n == 0) return MPCR_OK;
    if (!d_text || !d_ws || !d_seq) return fail(MPCR_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    const uint64_t n_blk = (n + kFastaBlock - 1) / kFastaBlock;
    fasta_compact<<<(uint32_t)((n_blk + kFastaWarps - 1) / kFastaWarps), kFastaThreads, 0, st>>>(d_text, n, (const uint64_t*)d_ws, d_seq);
    c->launches++;
    CU(cudaGetLastError());
    return MPCR_OK;
}

int mpcr_table_build(mpcr_ctx* c, const uint8_t* h_blob, const uint64_t* h_off, const uint32_t* h_pcr, uint32_t n_lines,
                     const uint8_t* h_plut, void* stream) {
    if (!c || !h_plut) return fail(MPCR_EINVAL, "null argument");
    if (n_lines && (!h_blob || !h_off || !h_pcr)) return fail(MPCR_EINVAL, "null argument");
    if (n_lines >= (1u << 29)) return fail(MPCR_EINVAL, "too many STS lines");
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    free_table(c);
    const int W = c->prm.wordsize;
    const bool sampled = c->samp_role == 1;
    const int WS = sampled ? c->samp_w : (c->ext_which == 2 ? c->ext_w : W);   // key width of THIS table
    c->scan_w = WS;
    const uint32_t n_rec = 2 * n_lines;
    const uint32_t per_rec = sampled ? (uint32_t)c->samp_s : 1u;   // table items per record
    if ((uint64_t)n_rec * per_rec >= (1ull << 30)) return fail(MPCR_EINVAL, "too many table items");
    const uint32_t n_items = n_rec * per_rec;
    c->n_rec = n_rec;
    c->n_valid = 0;
    c->max_hash_off = 0;
    c->max_len = 0;
    c->max_pcr = 0;
    // host-side prefix of primer word offsets: record r owns [word_off[2r], word_off[2r+1]) for P1 and
    // [word_off[2r+1], word_off[2r+2]) for P2, each primer = 2 * ceil(len/16) words (nibbles + aux).
    std::vector<uint32_t> word_off(2 * (size_t)n_rec + 1);
    uint64_t acc = 0;
    for (uint32_t l = 0; l < n_lines; ++l) {
        const uint64_t n1 = h_off[2 * l + 1] - h_off[2 * l], n2 = h_off[2 * l + 2] - h_off[2 * l + 1];
        if (n1 > 65535 || n2 > 65535) return fail(MPCR_EINVAL, "primer longer than 65535 bases at STS entry %u", l);
        if (n1 > c->max_len) c->max_len = (uint32_t)n1;
        if (n2 > c->max_len) c->max_len = (uint32_t)n2;
        if (h_pcr[l] > c->max_pcr) c->max_pcr = h_pcr[l];
        const uint32_t w1 = 2 * (uint32_t)((n1 + 15) / 16), w2 = 2 * (uint32_t)((n2 + 15) / 16);
        const size_t r = 2 * (size_t)l;
        // "+" record: P1 = primer1 (w1), P2 = primer2 (w2); "-" record: P1 = primer2 (w2), P2 = rc(primer1) (w1)
        word_off[2 * r] = (uint32_t)acc; acc += w1;
        word_off[2 * r + 1] = (uint32_t)acc; acc += w2;
        word_off[2 * r + 2] = (uint32_t)acc; acc += w2;
        word_off[2 * r + 3] = (uint32_t)acc; acc += w1;
        if (acc >= 0xFFFFFFFFull) return fail(MPCR_EINVAL, "primer blob too large");
    }
    word_off[2 * (size_t)n_rec] = (uint32_t)acc;
    c->total_pwords = acc;

    // first-level filter: as many words as the scanner CTA's shared memory leaves (a multiple of 4 words)
    {
        long words = ((long)c->max_smem_optin - ScanSmem::kFilterOff) / 4;
        if (const char* env = getenv("MPCR_FILTER_WORDS")) {
            long v = atol(env);
            if (v >= 4 && v < words) words = v;
        }
        if (words < 64) return fail(MPCR_ECUDA, "device shared memory too small for the scanner (%d bytes opt-in)", c->max_smem_optin);
        c->filter_words = (uint32_t)words & ~3u;
        // contiguous 11-letter keys in a direct table are scanned by the instantiations that read the filter through the
        // linear map (scan_kernel: LINEAR); split keys (a block table behind a gap) and sampled tables are not
        c->filter_linear = MPCR_LINEAR_FILTER && WS == kLinearW && !sampled && !(c->ext_which == 2 && c->ext_gap > 0) &&
                           filter_linear_setup(c->filter_words, &c->filter_scale, &c->filter_bias);
    }
    c->n_keys = 0;
    const size_t blob_bytes = n_lines ? (size_t)h_off[2 * (size_t)n_lines] : 0;
    uint8_t *d_blob = nullptr, *d_plut = nullptr;
    uint64_t* d_off = nullptr;
    uint32_t *d_pcr = nullptr, *d_woff = nullptr, *d_stats = nullptr, *d_tags = nullptr;
    Item<2>*d_pairs = nullptr, *d_pairs2 = nullptr, *d_sorted = nullptr;
    int rc = MPCR_OK;
#define CUG(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            rc = fail(MPCR_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            goto done;                                                                                   \
        }                                                                                                \
    } while (0)
    CUG(cudaMalloc(&c->d_filter, (size_t)c->filter_words * 4));
    CUG(cudaMemsetAsync(c->d_filter, 0, (size_t)c->filter_words * 4, st));
    if (n_lines == 0) {
        c->smap = SlotMap{1023u, 0u};
        CUG(cudaMalloc(&c->d_slots, 1024 * sizeof(Slot)));
        CUG(cudaMemsetAsync(c->d_slots, 0xFF, 1024 * sizeof(Slot), st));
        CUG(cudaStreamSynchronize(st));
        c->table_ready = true;
        goto done;
    }
    {
        CUG(cudaMalloc(&d_blob, blob_bytes + 16));
        CUG(cudaMalloc(&d_plut, 256));
        CUG(cudaMalloc(&d_off, (2 * (size_t)n_lines + 1) * sizeof(uint64_t)));
        CUG(cudaMalloc(&d_pcr, (size_t)n_lines * 4));
        CUG(cudaMalloc(&d_woff, word_off.size() * 4));
        CUG(cudaMalloc(&d_stats, 16));
        CUG(cudaMalloc(&d_pairs, (size_t)n_items * sizeof(Item<2>)));
        CUG(cudaMalloc(&d_pairs2, (size_t)n_items * sizeof(Item<2>)));
        CUG(cudaMalloc(&d_tags, (size_t)n_items * 4));
        CUG(cudaMalloc(&c->d_meta, (size_t)n_rec * sizeof(RecMeta)));
        CUG(cudaMalloc(&c->d_pwords, (size_t)(acc + 2) * sizeof(uint64_t)));
        CUG(cudaMemcpyAsync(d_blob, h_blob, blob_bytes, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_plut, h_plut, 256, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_off, h_off, (2 * (size_t)n_lines + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_pcr, h_pcr, (size_t)n_lines * 4, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_woff, word_off.data(), word_off.size() * 4, cudaMemcpyHostToDevice, st));
        CUG(cudaMemsetAsync(d_stats, 0, 16, st));
        encode_records<<<(n_items + 127) / 128, 128, 0, st>>>(d_blob, d_off, d_pcr, n_lines, d_plut, d_woff, W,
                                                               c->ext_which ? c->ext_w : W, sampled ? 0 : c->ext_which,
                                                               c->ext_block, c->ext_gap, c->ext_span, c->true_strands, c->part, c->parts, c->samp_w, c->samp_s,
                                                               c->samp_role, c->d_meta, c->d_pwords, d_pairs, d_tags, d_stats);
        c->launches++;
        CUG(cudaGetLastError());
        uint32_t stats[4] = {0, 0, 0, 0};
        CUG(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
        CUG(cudaStreamSynchronize(st));
        c->n_valid = stats[2];        // records in this table (stats[0] = records the reference inserts)
        c->max_hash_off = stats[1];
        // stable sort by (invalid flag, key): LSD passes over the key digits, then the flag bit
        PassDesc passes[8];
        int np = add_passes(passes, 0, 0, wmask_of(WS));
        passes[np].field = 1; passes[np].shift = 31; passes[np].mask = 1; ++np;
        const uint32_t nblk = (n_items + kSortItemsPerBlock - 1) / kSortItemsPerBlock;
        rc = ensure((void**)&c->d_counts, &c->counts_cap, (size_t)256 * nblk * 4);
        if (rc) goto done;
        c->launches += radix_sort<2>(d_pairs, d_pairs2, n_items, nullptr, 0, nullptr, passes, np, c->d_counts, st, &d_sorted);
        CUG(cudaGetLastError());
        // slot table: direct-indexed by the key while 4^W slots stay L2-sized (W <= 11 -> 64 MiB), else open
        // addressing at load <= 1/8 (distinct seeds <= min(records, 4^W))
        uint32_t nslots;
        if (WS <= 11) {
            nslots = 1u << (2 * WS);
            c->smap.direct = 1;
        } else {
            nslots = 1024;
            while (nslots < 8ull * c->n_valid && nslots < (1u << 25)) nslots <<= 1;
            while (nslots < 2ull * c->n_valid + 2) nslots <<= 1;
            c->smap.direct = 0;
        }
        c->smap.mask = nslots - 1;
        CUG(cudaMalloc(&c->d_slots, (size_t)nslots * sizeof(Slot)));
        CUG(cudaMemsetAsync(c->d_slots, 0xFF, (size_t)nslots * sizeof(Slot), st));
        CUG(cudaMalloc(&c->d_bucket, ((size_t)c->n_valid + 33) * sizeof(BucketEntry)));   // +32: dense walks read whole warps
        CUG(cudaMemsetAsync(c->d_bucket, 0, ((size_t)c->n_valid + 33) * sizeof(BucketEntry), st));
        if (sampled) {   // the first-level filter of a sampled table lives in global memory: >= 32 bits per key
            uint32_t lg = 10;
            while (lg < 26 && (1ull << lg) < (uint64_t)c->n_valid) ++lg;
            c->bloom_shift = 32 - lg;
            CUG(cudaMalloc(&c->d_bloom, ((size_t)1 << lg) * 4));
            CUG(cudaMemsetAsync(c->d_bloom, 0, ((size_t)1 << lg) * 4, st));
        }
        if (c->n_valid) {
            CUG(cudaMemsetAsync(d_stats, 0, 16, st));
            build_buckets<<<(c->n_valid + 255) / 256, 256, 0, st>>>(d_sorted, c->n_valid, d_tags, c->d_bucket,
                                                                     c->d_slots, c->smap, c->d_filter,
                                                                     c->filter_words, filter_mul(WS), WS, d_stats,
                                                                     c->d_bloom, c->bloom_shift, c->filter_linear,
                                                                     c->filter_scale, c->filter_bias);
            c->launches++;
            CUG(cudaGetLastError());
            if (!c->smap.direct) {
                mark_chains<<<(c->n_valid + 255) / 256, 256, 0, st>>>(d_sorted, c->n_valid, c->d_slots, c->smap);
                c->launches++;
                CUG(cudaGetLastError());
            }
            CUG(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
        }
        CUG(cudaStreamSynchronize(st));
        c->n_keys = stats[0];
        // a quarter of all words are seeds (or seeds are shared by several records on average): no filter can help,
        // use the dense scanner
        c->dense = c->n_keys && (c->n_valid >= 3ull * c->n_keys ||
                                 (2 * WS < 32 && 4ull * c->n_keys >= (1ull << (2 * WS))));
        if (const char* env = getenv("MPCR_DENSE")) c->dense = atoi(env) != 0;
        if (c->ext_which == 2 && c->ext_gap > 0) c->dense = false;   // only the sparse scanner assembles split keys
        c->table_ready = true;
    }
done:
    cudaFree(d_blob); cudaFree(d_plut); cudaFree(d_off); cudaFree(d_pcr); cudaFree(d_woff); cudaFree(d_stats);
    cudaFree(d_pairs); cudaFree(d_pairs2); cudaFree(d_tags);
    if (rc) free_table(c);
    return rc;
#undef CUG
}

int mpcr_table_records(mpcr_ctx* c, int32_t* h_hash_offset, uint32_t* h_hash) {
    if (!c || !c->table_ready) return fail(MPCR_ESTATE, "table not built");
    if (c->n_rec == 0) return MPCR_OK;
    GUARD(c);
    std::vector<RecMeta> m(c->n_rec);
    CU(cudaMemcpy(m.data(), c->d_meta, (size_t)c->n_rec * sizeof(RecMeta), cudaMemcpyDeviceToHost));
    for (uint32_t r = 0; r < c->n_rec; ++r) {
        if (h_hash_offset) h_hash_offset[r] = (m[r].flags & 1) ? (int32_t)m[r].hash_off : -1;
        if (h_hash) h_hash[r] = m[r].hash_be;
    }
    return MPCR_OK;
}

int mpcr_table_primer_words(mpcr_ctx* c, uint32_t rec, int which, uint64_t* h_words, uint32_t max_words,
                            uint32_t* n_words) {
    if (!c || !c->table_ready) return fail(MPCR_ESTATE, "table not built");
    if (rec >= c->n_rec || (which != 1 && which != 2)) return fail(MPCR_EINVAL, "bad record / primer index");
    GUARD(c);
    RecMeta m;
    CU(cudaMemcpy(&m, c->d_meta + rec, sizeof m, cudaMemcpyDeviceToHost));
    const uint32_t len = which == 1 ? m.len1 : m.len2, nw = 2 * ((len + 15) / 16);
    if (n_words) *n_words = nw;
    if (nw > max_words) return fail(MPCR_EINVAL, "buffer too small");
    CU(cudaMemcpy(h_words, c->d_pwords + (which == 1 ? m.p1_word : m.p2_word), (size_t)nw * 8, cudaMemcpyDeviceToHost));
    return MPCR_OK;
}

static uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

uint64_t mpcr_tile_bases(void) { return (uint64_t)kTileBases; }
uint64_t mpcr_halo_left(const mpcr_ctx* c) { return c ? round_up((uint64_t)c->max_hash_off + 64, 128) : 0; }
uint64_t mpcr_halo_right(const mpcr_ctx* c) {
    if (!c) return 0;
    // furthest base a position can touch: p - hash_off + pcr_size + margin (engine.py:543,582) and the
    // W-1 / primer overhang of the last position; a tile is owned by the shard holding its FIRST base, so the
    // last tile may run up to kTileBases past shard_end
    return round_up(c->max_pcr + (uint64_t)c->prm.margin + c->max_len + 64 + 128, 128) + kTileBases;
}

// Plane extent the current view touches.  Descriptors ascend in the padded coordinate, so the first one bounds the
// reads on the left (a primer-1 site starts at most max_hash_off bases in front of its seed, never in front of its
// contig) and the last one on the right (scanner: whole units + read-ahead; verifier: the end of the mate window,
// never past the contig's end; both plus the word over-read of the plane accessors).
static void view_extent(mpcr_ctx* c) {
    c->tiles_min = 0;
    c->tiles_need = 0;
    if (c->view_count == 0 || c->h_tiles.size() < (size_t)c->view_first + c->view_count) return;
    const TileDesc& f = c->h_tiles[c->view_first];
    const TileDesc& l = c->h_tiles[c->view_first + c->view_count - 1];
    const int64_t f_contig = f.gbase - (int64_t)f.lstart;
    const int64_t reach = f.gbase - (int64_t)c->max_hash_off - (c->samp_role == 1 ? (int64_t)c->samp_s : 0);
    c->tiles_min = reach > f_contig ? reach : f_contig;
    const uint64_t scan_end = (uint64_t)l.gbase + round_up(l.nbases, 2048) + 256;
    const uint64_t contig_end = (uint64_t)(l.gbase - (int64_t)l.lstart) + l.length;
    uint64_t mate_end = (uint64_t)l.gbase + l.nbases + c->max_pcr + (uint64_t)c->prm.margin + c->max_len;
    if (mate_end > contig_end) mate_end = contig_end;
    mate_end += 64;
    c->tiles_need = scan_end > mate_end ? scan_end : mate_end;
}

static int build_tiles_impl(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, uint64_t origin, uint64_t sb,
                            uint64_t se, cudaStream_t st);
static int build_tiles(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, uint64_t origin, uint64_t sb,
                       uint64_t se, cudaStream_t st) {
    const int rc = build_tiles_impl(c, contigs, n_contigs, origin, sb, se, st);
    if (rc == MPCR_OK) view_extent(c);
    return rc;
}
static int build_tiles_impl(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, uint64_t origin, uint64_t sb,
                            uint64_t se, cudaStream_t st) {
    // signature of the layout: rebuild the descriptor array only when it changes
    uint64_t sig = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    mix(n_contigs); mix(origin); mix((uint64_t)c->prm.wordsize);
    for (uint32_t i = 0; i < n_contigs; ++i) { mix(contigs[i].gstart); mix(contigs[i].length); }
    if (sig == c->tiles_sig && c->d_tiles) {
        if (sb == c->tiles_sb && se == c->tiles_se) {
            c->view_first = 0; c->view_count = c->n_tiles;
            return MPCR_OK;
        }
        // A sub-range of the cached one that is cut between contigs (or at the cached bounds) is a contiguous run of
        // the cached descriptors (tiles never span contigs): no rebuild, no upload, no host round trip -- this is what
        // lets a genome be scanned contig by contig while it is still uploading (mpcr_scan_prepare + append mode).
        auto clean = [&](uint64_t x) {
            if (x == c->tiles_sb || x == c->tiles_se) return true;
            uint32_t lo = 0, hi = n_contigs;   // last contig that starts before x
            while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (contigs[mid].gstart < x) lo = mid + 1; else hi = mid; }
            return lo == 0 || x >= contigs[lo - 1].gstart + contigs[lo - 1].length;
        };
        if (sb >= c->tiles_sb && se <= c->tiles_se && sb <= se && clean(sb) && clean(se)) {
            const auto b0 = std::lower_bound(c->h_tile_g.begin(), c->h_tile_g.end(), sb);
            const auto b1 = std::lower_bound(c->h_tile_g.begin(), c->h_tile_g.end(), se);
            c->view_first = (uint32_t)(b0 - c->h_tile_g.begin());
            c->view_count = (uint32_t)(b1 - b0);
            return MPCR_OK;
        }
    }
    // tile size: kTileBases for big inputs; halved (down to one 2048-position unit) while the scanner's warps would
    // get fewer than ~8 tiles each, so small inputs still spread over the whole GPU
    uint64_t span = 0;
    for (uint32_t i = 0; i < n_contigs; ++i) {
        const uint64_t g0 = contigs[i].gstart, g1 = g0 + contigs[i].length;
        const uint64_t a0 = g0 > sb ? g0 : sb, a1 = g1 < se ? g1 : se;
        if (a1 > a0) span += a1 - a0;
    }
    uint64_t tb = kTileBases;
    // (32 tiles per warp: the warps walk their tiles in lock step, so the launch ends with up to one tile-time of
    // partly idle SMs -- 3 % of the launch at 32 tiles per warp, 10 % at 10)
    // (tuning builds, scan kernel for one eighth of cfg3 / all of it: 16 tiles per warp 0.350 / 2.620 ms, 32: 0.353 / 2.620,
    // 64: 0.363 / 2.626, 128: 0.363 / 2.644 -- finer tiles cost more in descriptor fetches than they balance)
#ifndef MPCR_WANT_TILES
#define MPCR_WANT_TILES 32
#endif
    const uint64_t want_tiles = (uint64_t)MPCR_WANT_TILES * (uint64_t)c->sm_count * (kScanThreads / 32);
    while (tb > 2048 && span / tb < want_tiles) tb >>= 1;
    std::vector<TileDesc> tiles;
    for (uint32_t i = 0; i < n_contigs; ++i) {
        const uint64_t L = contigs[i].length, g0 = contigs[i].gstart;
        if (L <= (uint64_t)c->prm.wordsize) continue;  // engine.py:458 (Q3: len <= W is skipped)
        if (g0 & 127u) return fail(MPCR_EINVAL, "contig %u: gstart not a multiple of 128", i);
        if (L >= (1ull << 31)) return fail(MPCR_EINVAL, "contig %u longer than 2^31-1 bases", i);
        // Ownership is decided per 2048-position unit (unit starts are multiples of 2048 from the contig start): a
        // unit belongs to the range that holds its first base.  That is independent of the tile size chosen above,
        // so neighbouring ranges -- other ranks, or the ranges of a genome scanned while it uploads -- never overlap
        // and never leave a gap, whatever tile size each of them picks.
        if (se <= g0 || sb >= g0 + L) continue;
        const uint64_t unit = 2048;
        const uint64_t lo = sb > g0 ? round_up(sb - g0, unit) : 0;             // first owned unit
        const uint64_t stop = se - g0 < L ? se - g0 : L;                       // owned unit starts are < stop
        if (lo >= stop) continue;
        const uint64_t hi = round_up(stop, unit) < L ? round_up(stop, unit) : L;   // positions the owned units cover
        for (uint64_t ls = lo; ls < hi; ls += tb) {
            TileDesc t;
            if (g0 + ls < origin)
                return fail(MPCR_EINVAL, "shard_begin lies in front of plane_origin (contig %u, base %llu)", i,
                            (unsigned long long)ls);
            t.gbase = (int64_t)(g0 + ls - origin);
            t.contig = i;
            t.lstart = (uint32_t)ls;
            t.length = (uint32_t)L;
            t.nbases = (uint32_t)(hi - ls < tb ? hi - ls : tb);
            tiles.push_back(t);
        }
    }
    c->n_tiles = (uint32_t)tiles.size();
    c->view_first = 0; c->view_count = c->n_tiles;
    c->tiles_sb = sb; c->tiles_se = se;
    c->h_tile_g.resize(tiles.size());
    for (size_t i = 0; i < tiles.size(); ++i) c->h_tile_g[i] = (uint64_t)(tiles[i].gbase + (int64_t)origin);
    c->h_tiles = tiles;
    c->lay_contigs = n_contigs;
    c->lay_max_len = 0;
    for (uint32_t i = 0; i < n_contigs; ++i)
        if (contigs[i].length > c->lay_max_len) c->lay_max_len = contigs[i].length;
    int rc = ensure((void**)&c->d_tiles, &c->tiles_cap, (tiles.size() + 1) * sizeof(TileDesc));
    if (rc) return rc;
    if (!tiles.empty())
        CU(cudaMemcpyAsync(c->d_tiles, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    // the contig table of this layout (global start coordinates): what the bucket sort of the hits slices on
    std::vector<unsigned long long> cg(n_contigs ? n_contigs : 1, 0ull);
    c->genome_end = 0;
    for (uint32_t i = 0; i < n_contigs; ++i) {
        cg[i] = contigs[i].gstart;
        if (contigs[i].gstart + contigs[i].length > c->genome_end) c->genome_end = contigs[i].gstart + contigs[i].length;
    }
    rc = ensure((void**)&c->d_contig_g, &c->contig_g_cap, cg.size() * sizeof(unsigned long long));
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_contig_g, cg.data(), cg.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    c->n_contig_g = n_contigs;
    CU(cudaStreamSynchronize(st));  // `tiles` is pageable host memory going out of scope
    c->tiles_sig = sig;
    return MPCR_OK;
}

int mpcr_scan_prepare(mpcr_ctx* c, const mpcr_contig* h_contigs, uint32_t n_contigs, uint64_t plane_origin,
                      uint64_t shard_begin, uint64_t shard_end, void* stream) {
    if (!c || (n_contigs && !h_contigs)) return fail(MPCR_EINVAL, "null argument");
    if ((plane_origin & 127u) || (shard_begin & 127u)) return fail(MPCR_EINVAL, "origin / shard_begin must be multiples of 128");
    GUARD(c);
    return build_tiles(c, h_contigs, n_contigs, plane_origin, shard_begin, shard_end, (cudaStream_t)stream);
}

int mpcr_scan(mpcr_ctx* c, const mpcr_contig* h_contigs, uint32_t n_contigs, const void* d_plane2, const void* d_plane4,
              const void* d_valid, uint64_t plane_origin, uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end,
              mpcr_hit* d_hits, uint64_t capacity, uint64_t* d_count, void* stream) {
    if (!c || !d_count) return fail(MPCR_EINVAL, "null argument");
    if (!c->table_ready) return fail(MPCR_ESTATE, "mpcr_scan called before mpcr_table_build");
    if (n_contigs && (!h_contigs || !d_plane2 || !d_plane4 || !d_valid)) return fail(MPCR_EINVAL, "null argument");
    if ((plane_origin & 127u) || (shard_begin & 127u)) return fail(MPCR_EINVAL, "origin / shard_begin must be multiples of 128");
    if (capacity && !d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    cudaStream_t st = (cudaStream_t)stream;
    GUARD(c);
    int rc = build_tiles(c, h_contigs, n_contigs, plane_origin, shard_begin, shard_end, st);
    if (rc) return rc;
    if (!c->append) CU(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    c->scan_timed[c->ev_slot] = false;
    if (c->view_count == 0 || c->n_valid == 0) return MPCR_OK;
    // The planes must hold every base the kernels touch: units are staged whole (2048 positions + 128 bases of
    // read-ahead), the verifier reads up to the mate window's end (+ 64 bases of word over-read), and nothing lies in
    // front of the origin.  plane_bases = bases the three plane allocations hold, counted from plane_origin.
    if (c->tiles_min < 0)
        return fail(MPCR_EINVAL, "shard_begin - halo lies in front of plane_origin (planes start %lld bases too late)",
                    (long long)-c->tiles_min);
    if (c->tiles_need > plane_bases)
        return fail(MPCR_EINVAL, "planes too small for the shard: %llu bases given, %llu needed (last unit + read-ahead, "
                    "mate window, word over-read)", (unsigned long long)plane_bases,
                    (unsigned long long)c->tiles_need);
    if (c->ctl_dirty) {   // only after a debug run (the verifier, which re-zeroes the control words, did not run)
        CU(cudaMemsetAsync(c->d_tile_counter, 0, 32, st));
        CU(cudaMemsetAsync(c->d_surv_ctl, 0, (size_t)kSurvLists * kSurvCtlStride * 4, st));
        c->ctl_dirty = false;
    }
    // survivor list: ~1 position in 10^4 survives on random sequence; overflow falls back to in-kernel serial verify
    {
        uint64_t scanned = 0;
        for (uint32_t i = 0; i < n_contigs; ++i) scanned += h_contigs[i].length;
        // entries per sub-list: room for one survivor per 64 scanned positions overall
        size_t want = (size_t)(scanned / 64 / kSurvLists + 1024) * kSurvLists * sizeof(Survivor);
        if (want > ((size_t)1 << 30)) want = (size_t)1 << 30;
        if (c->surv_bytes < want) {
            int rc2 = ensure((void**)&c->d_surv, &c->surv_bytes, want);
            if (rc2) return rc2;
        }
    }
    ScanArgs a;
    a.p2 = (const uint64_t*)d_plane2; a.p4 = (const uint64_t*)d_plane4; a.valid = (const uint64_t*)d_valid;
    a.tiles = c->d_tiles + c->view_first; a.n_tiles = c->view_count;
    a.slots = c->d_slots; a.smap = c->smap; a.bucket = c->d_bucket; a.meta = c->d_meta; a.pwords = c->d_pwords;
    a.filter = c->d_filter; a.filter_words = c->filter_words; a.cw = filter_mul(c->scan_w);
    a.lin_scale = c->filter_scale; a.lin_bias = c->filter_bias; a.lin_exp = kLinearExp;
    a.prm.W = c->scan_w; a.prm.M = c->prm.margin; a.prm.N = c->prm.mismatches; a.prm.X = c->prm.three_prime_match;
    a.prm.iupac = c->prm.iupac_mode ? 1 : 0;
    const bool gapped = c->ext_which == 2 && c->ext_block > 0 && c->samp_role != 1;   // a block table (gap 0: the first block)
    a.prm.gap = gapped ? c->ext_gap : 0;
    a.prm.block = gapped ? c->ext_block : 0;
    a.debug = c->env_debug;
    a.k4 = 4u;
    a.samp_s = c->samp_role == 1 ? (uint32_t)c->samp_s : 1u;
    a.bloom = c->d_bloom;
    a.bloom_shift = c->bloom_shift;
    a.hits = d_hits; a.capacity = capacity; a.count = (unsigned long long*)d_count; a.tile_counter = c->d_tile_counter;
    a.surv = c->d_surv; a.surv_cap = (uint32_t)(c->surv_bytes / sizeof(Survivor) / kSurvLists); a.surv_ctl = c->d_surv_ctl;
    if (c->env_surv_cap >= 0 && (uint32_t)c->env_surv_cap < a.surv_cap)   // test hook: force the list-full path
        a.surv_cap = (uint32_t)c->env_surv_cap;
    const size_t smem = (size_t)ScanSmem::kFilterOff + (size_t)c->filter_words * 4;
    uint32_t grid = (uint32_t)c->sm_count;
    if (grid > c->view_count) grid = c->view_count;
    cudaEvent_t* ev = c->evs[c->ev_slot];
    CU(cudaEventRecord(ev[0], st));
    if (c->samp_role == 1) {
        sampled_scan_kernel<<<c->sm_count * 4, 256, 0, st>>>(a);
    } else if (c->dense) {
        dense_scan_kernel<<<c->sm_count * 8, 256, 0, st>>>(a);
    } else {
        // instantiations: the open-addressed table (W >= 12), the narrow filter (W < 6), the general direct table, and
        // the reference's default word size with the usual mismatch budgets fixed at compile time
        void (*kern)(const ScanArgs) = scan_kernel<true, false, 0, -1>;
        const bool w12n1 = a.prm.W == 12 && a.prm.N == 1;   // block tables of -W 8 / -W 9 searches with one mismatch
        if (a.prm.gap > 0)
            kern = c->smap.direct ? scan_kernel<true, false, 0, -1, true>
                                  : (w12n1 ? scan_kernel<true, true, 12, 1, true> : scan_kernel<true, true, 0, -1, true>);
        else if (!c->smap.direct) kern = w12n1 ? scan_kernel<true, true, 12, 1> : scan_kernel<true, true, 0, -1>;
        else if (a.prm.W < 6) kern = scan_kernel<false, false, 0, -1>;
        else if (a.prm.W == 11 && a.prm.N == 0) kern = scan_kernel<true, false, 11, 0>;
        else if (a.prm.W == 11 && a.prm.N == 1) kern = scan_kernel<true, false, 11, 1>;
        else if (a.prm.W == 11 && a.prm.N == 2) kern = scan_kernel<true, false, 11, 2>;
        else if (a.prm.W == 11) kern = scan_kernel<true, false, 11, -1>;
        // the filter was built for the map the chosen instantiation reads it through
        const bool kern_linear = MPCR_LINEAR_FILTER && a.prm.W == kLinearW && c->smap.direct && a.prm.gap == 0;
        if (kern_linear != c->filter_linear) return fail(MPCR_ESTATE, "filter map / scanner instantiation mismatch");
        if (c->attr_kern != (const void*)kern || c->attr_smem != smem) {   // once per (kernel, size), not per launch
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            c->attr_kern = (const void*)kern;
            c->attr_smem = smem;
        }
        kern<<<grid, kScanThreads, smem, st>>>(a);
    }
    CU(cudaEventRecord(ev[1], st));
    c->launches++;
    CU(cudaGetLastError());
    if (!a.debug) {
        verify_kernel<<<c->sm_count * kVerifyCtasPerSm, 256, 0, st>>>(a);
        c->launches++;
        CU(cudaGetLastError());
    } else {
        c->ctl_dirty = true;
    }
    CU(cudaEventRecord(ev[2], st));
    c->scan_timed[c->ev_slot] = true;
    return MPCR_OK;
}

static int sort_hits_impl(mpcr_ctx* c, mpcr_hit* d_hits, uint64_t n_host, const unsigned long long* d_n,
                          uint64_t n_hint, cudaStream_t st, bool optimistic, bool* only_bucket);
int mpcr_scan_sorted_async(mpcr_ctx* const* ctxs, uint32_t n_ctx, const mpcr_contig* h_contigs, uint32_t n_contigs,
                           const void* d_plane2, const void* d_plane4, const void* d_valid, uint64_t plane_origin,
                           uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end, mpcr_hit* d_hits,
                           uint64_t capacity, uint64_t* d_count, uint64_t* h_result, uint64_t n_hint, int sort, int slot,
                           void* stream) {
    if (!ctxs || n_ctx == 0 || !d_count) return fail(MPCR_EINVAL, "null argument");
    if (slot < 0 || slot >= MPCR_MAX_SLOTS) return fail(MPCR_EINVAL, "slot must be in [0, %d)", MPCR_MAX_SLOTS);
    for (uint32_t i = 0; i < n_ctx; ++i)
        if (!ctxs[i]) return fail(MPCR_EINVAL, "null context");
    int rc = MPCR_OK;
    std::vector<int> saved(n_ctx);
    for (uint32_t i = 0; i < n_ctx; ++i) saved[i] = ctxs[i]->append;
    for (uint32_t i = 0; i < n_ctx && rc == MPCR_OK; ++i) {
        ctxs[i]->append = i == 0 ? 0 : 1;    // the first table zeroes the count, the others append behind it
        ctxs[i]->ev_slot = slot;
        rc = mpcr_scan(ctxs[i], h_contigs, n_contigs, d_plane2, d_plane4, d_valid, plane_origin, plane_bases, shard_begin,
                       shard_end, d_hits, capacity, d_count, stream);
    }
    for (uint32_t i = 0; i < n_ctx; ++i) ctxs[i]->append = saved[i];
    if (rc) return rc;
    mpcr_ctx* c0 = ctxs[0];
    bool only_bucket = false;
    if (sort && capacity >= 2) {
        if (!d_hits) return fail(MPCR_EINVAL, "null hit buffer");
        GUARD(c0);
        // with a result buffer the caller looks at the bucket sort's fall-back flag itself (mpcr_sort_finish): a list
        // hinted short then gets the four bucket-sort launches and nothing else
        rc = sort_hits_impl(c0, d_hits, capacity, (const unsigned long long*)d_count, n_hint, (cudaStream_t)stream,
                            h_result != nullptr, &only_bucket);
        if (rc) return rc;
    }
    if (h_result) {
        GUARD(c0);
        cudaStream_t st = (cudaStream_t)stream;
        h_result[1] = 0;
        CU(cudaMemcpyAsync(h_result, d_count, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (only_bucket)   // the fall-back flag comes back with the count (4 bytes into the low half of h_result[1])
            CU(cudaMemcpyAsync(h_result + 1, c0->d_bsort + (size_t)kSortBuckets * 2 + 4 + kBucketSortMax, 4, cudaMemcpyDeviceToHost, st));
    }
    return MPCR_OK;
}

int mpcr_sort_finish(mpcr_ctx* c, mpcr_hit* d_hits, const uint64_t* h_result, uint64_t capacity, void* stream) {
    if (!c || !h_result) return fail(MPCR_EINVAL, "null argument");
    if (!h_result[1]) return MPCR_OK;
    // the list piled up in one slice (or outgrew the hint): radix passes, with the count known now
    const uint64_t n = h_result[0] < capacity ? h_result[0] : capacity;
    if (n < 2) return MPCR_OK;
    if (!d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    GUARD(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = sort_hits_impl(c, d_hits, n, nullptr, 0, st, false, nullptr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    return MPCR_OK;
}

int mpcr_scan_sorted(mpcr_ctx* const* ctxs, uint32_t n_ctx, const mpcr_contig* h_contigs, uint32_t n_contigs,
                     const void* d_plane2, const void* d_plane4, const void* d_valid, uint64_t plane_origin,
                     uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end, mpcr_hit* d_hits, uint64_t capacity,
                     uint64_t* d_count, uint64_t* h_count, uint64_t n_hint, int sort, void* stream) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return fail(MPCR_EINVAL, "null argument");
    mpcr_ctx* c0 = ctxs[0];
    uint64_t* h_result = nullptr;
    if (h_count) {
        if (!c0->h_flag) {
            GUARD(c0);
            CU(cudaMallocHost(&c0->h_flag, 32));
        }
        h_result = reinterpret_cast<uint64_t*>(c0->h_flag);
    }
    int rc = mpcr_scan_sorted_async(ctxs, n_ctx, h_contigs, n_contigs, d_plane2, d_plane4, d_valid, plane_origin, plane_bases,
                                    shard_begin, shard_end, d_hits, capacity, d_count, h_result, n_hint, sort, 0, stream);
    if (rc || !h_count) return rc;
    {
        GUARD(c0);
        CU(cudaStreamSynchronize((cudaStream_t)stream));
    }
    *h_count = h_result[0];
    return mpcr_sort_finish(c0, d_hits, h_result, capacity, stream);
}

static float event_ms(mpcr_ctx* c, int slot, int from, int to) {
    if (!c || slot < 0 || slot >= MPCR_MAX_SLOTS || !c->scan_timed[slot]) return 0.f;
    float ms = 0.f;
    if (cudaEventSynchronize(c->evs[slot][to]) != cudaSuccess) return 0.f;
    if (cudaEventElapsedTime(&ms, c->evs[slot][from], c->evs[slot][to]) != cudaSuccess) return 0.f;
    return ms;
}
float mpcr_last_scan_ms(mpcr_ctx* c) { return c ? event_ms(c, c->ev_slot, 0, 1) : 0.f; }
float mpcr_last_verify_ms(mpcr_ctx* c) { return c ? event_ms(c, c->ev_slot, 1, 2) : 0.f; }
float mpcr_slot_scan_ms(mpcr_ctx* c, int slot) { return event_ms(c, slot, 0, 1); }
float mpcr_slot_verify_ms(mpcr_ctx* c, int slot) { return event_ms(c, slot, 1, 2); }

// optimistic: the caller synchronises right behind the sort and looks at the bucket sort's fall-back flag itself
// (mpcr_scan_sorted) -- then a list hinted short gets the four bucket-sort launches and nothing else.
static int sort_hits_impl(mpcr_ctx* c, mpcr_hit* d_hits, uint64_t n_host, const unsigned long long* d_n,
                          uint64_t n_hint, cudaStream_t st, bool optimistic, bool* only_bucket) {
    if (only_bucket) *only_bucket = false;
    static_assert(sizeof(mpcr_hit) == sizeof(Item<6>), "hit layout");
    int rc = ensure(&c->d_sort_tmp, &c->sort_tmp_cap, n_host * sizeof(mpcr_hit));
    if (rc) return rc;
    const uint64_t nblk = (n_host + kSortItemsPerBlock - 1) / kSortItemsPerBlock;
    rc = ensure((void**)&c->d_counts, &c->counts_cap, (size_t)256 * nblk * 4);
    if (rc) return rc;
    if (!c->d_long_runs) {
        CU(cudaMalloc(&c->d_long_runs, (size_t)kLongRunQueue * sizeof(LongRun) + 16));
        CU(cudaMemsetAsync(c->d_long_runs + kLongRunQueue, 0, 16, st));   // order_long_runs re-zeroes them itself
    }
    Item<6>*hits = (Item<6>*)d_hits, *tmp = (Item<6>*)c->d_sort_tmp;
    // LSD radix passes over pos1 then contig (fields 1, 0), digits bounded by the layout of the last scan; the rest of
    // the key (hash_off, rec, rank) only matters inside runs of equal (contig, pos1), which order_ties settles
    PassDesc passes[24];
    int np = 0;
    np = add_passes(passes, np, 1, c->lay_max_len ? c->lay_max_len : 0x7FFFFFFFull);
    np = add_passes(passes, np, 0, c->lay_contigs ? c->lay_contigs - 1 : 0xFFFFFFFFull);
    // Lists of up to 2^17 hits: bucket sort on the global coordinate (four small launches, complete order).  Launched when
    // the count is known to be that short, or unknown (hint 0), and a layout is known; a list that piles up in one
    // slice raises the fall-back flag on the device and the radix passes below take over.
    uint32_t skip = 0;
    const uint32_t* skip_off = nullptr;
    const uint64_t n_known = d_n ? n_hint : n_host;
    if (n_known <= kBucketSortMax && c->d_contig_g && c->genome_end) {   // n_known 0 = unknown
        const size_t words = (size_t)kSortBuckets * 2 + 4 + kBucketSortMax + 4;
        if (!c->d_bsort) {
            CU(cudaMalloc(&c->d_bsort, words * 4));
            CU(cudaMemsetAsync(c->d_bsort, 0, words * 4, st));   // the counters are re-zeroed by bsort_finish itself
        }
        BucketSortArgs b;
        b.contig_g = c->d_contig_g;
        b.n_contigs = c->n_contig_g;
        b.g_lo = c->tiles_sb < c->genome_end ? c->tiles_sb : 0;
        const uint64_t hi = c->tiles_se < c->genome_end ? c->tiles_se : c->genome_end;
        b.span = hi > b.g_lo ? hi - b.g_lo : 1;
        b.cnt = c->d_bsort;
        b.off = b.cnt + kSortBuckets;               // 16-byte aligned (vector stores in bsort_scan)
        b.slot = b.off + kSortBuckets + 4;
        b.fallback = b.slot + kBucketSortMax;
        const uint64_t grid_n = n_host < kBucketSortMax ? n_host : kBucketSortMax;
        const uint32_t gb = (uint32_t)((grid_n + 255) / 256);
        bsort_count<<<gb, 256, 0, st>>>(hits, d_n, n_host, b);
        bsort_scan<<<1, 1024, 0, st>>>(d_n, n_host, b);
        bsort_scatter<<<gb, 256, 0, st>>>(hits, tmp, d_n, n_host, b);
        bsort_finish<<<kSortBuckets / 256, 256, 0, st>>>(tmp, hits, b);
        c->launches += 4;
        skip = kBucketSortMax;
        skip_off = b.fallback;
        if (optimistic && n_known != 0) {   // hinted short: the caller checks the flag after its synchronisation
            if (only_bucket) *only_bucket = true;
            CU(cudaGetLastError());
            return MPCR_OK;
        }
    }
    Item<6>* sorted = hits;
    c->launches += radix_sort<6>(hits, tmp, n_host, d_n, skip, skip_off, passes, np, c->d_counts, st, &sorted, &skip);
    uint32_t* queue_ctl = reinterpret_cast<uint32_t*>(c->d_long_runs + kLongRunQueue);
    order_ties<<<(uint32_t)((n_host + 255) / 256), 256, 0, st>>>((const mpcr_hit*)sorted, d_hits, n_host, d_n, skip, skip_off,
                                                                 c->d_long_runs, queue_ctl);
    order_long_runs<<<(uint32_t)c->sm_count, 256, 0, st>>>(d_hits, c->d_long_runs, queue_ctl);
    c->launches += 2;
    CU(cudaGetLastError());
    return MPCR_OK;
}

int mpcr_sort_hits(mpcr_ctx* c, mpcr_hit* d_hits, uint64_t n, void* stream) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (n < 2) return MPCR_OK;
    if (!d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    GUARD(c);
    return sort_hits_impl(c, d_hits, n, nullptr, 0, (cudaStream_t)stream, false, nullptr);
}

int mpcr_sort_hits_dev(mpcr_ctx* c, mpcr_hit* d_hits, const uint64_t* d_count, uint64_t capacity, uint64_t n_hint,
                       void* stream) {
    if (!c || !d_count) return fail(MPCR_EINVAL, "null argument");
    if (capacity < 2) return MPCR_OK;
    if (!d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    GUARD(c);
    return sort_hits_impl(c, d_hits, capacity, (const unsigned long long*)d_count, n_hint, (cudaStream_t)stream, false, nullptr);
}

}  // extern "C"
