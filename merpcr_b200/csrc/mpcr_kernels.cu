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

#include <vector>

#include "../../include/merpcr_b200.h"
#include "mpcr_core.cuh"
#include "mpcr_sort.cuh"

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
    uint64_t* d_slots = nullptr;
    uint32_t slot_mask = 0;
    uint32_t* d_bucket = nullptr;
    uint32_t* d_filter = nullptr;
    uint32_t filter_bits = 0, filter_words = 0;
    int filter_exact = 0;
    uint32_t max_hash_off = 0, max_len = 0;
    uint64_t max_pcr = 0;
    bool table_ready = false;
    // tiles
    TileDesc* d_tiles = nullptr;
    size_t tiles_cap = 0;
    uint32_t n_tiles = 0;
    uint64_t tiles_sig = 0;
    uint32_t lay_contigs = 0, lay_max_len = 0;  // bounds of the last scanned layout (sort digit counts)
    uint32_t* d_tile_counter = nullptr;
    // sort scratch
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_cap = 0;
    uint32_t* d_counts = nullptr;
    size_t counts_cap = 0;
    uint8_t* d_lut = nullptr;  // 256 B genome LUT for pack
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool scan_timed = false;
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
    const bool full = (b0 + 64 <= n) && ((reinterpret_cast<uintptr_t>(ascii) & 15u) == 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
        if (full) {
            uint4 t = *reinterpret_cast<const uint4*>(ascii + b0 + 16 * q);
            w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint64_t i = b0 + 16 * q + 4 * k + b;
                    // bytes past n encode as 0 ('\0' maps to an invalid zero nibble in every LUT)
                    uint32_t c = i < n ? ascii[i] : 0u;
                    x |= c << (8 * b);
                }
                w[k] = x;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = 16 * q + 4 * k + b;  // base index inside the strip
                const uint32_t e = lut[(w[k] >> (8 * b)) & 0xFFu];
                p4[j >> 4] |= (uint64_t)(e & 15u) << (4 * (j & 15));
                p2[j >> 5] |= (uint64_t)((e >> 4) & 3u) << (2 * (j & 31));
                v |= (uint64_t)((e >> 6) & 1u) << j;
            }
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

// One thread per record slot r = 2*line + strand.  "+" : (P1,P2) = (primer1, primer2)   (engine.py:265-268)
//                                                  "-" : (P1,P2) = (primer2, revcomp(primer1)) (engine.py:273-279)
__global__ void __launch_bounds__(128) encode_records(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ off,
                                                      const uint32_t* __restrict__ pcr, uint32_t n_lines,
                                                      const uint8_t* __restrict__ plut,
                                                      const uint32_t* __restrict__ word_off,  // 2*n_rec+1 prefix
                                                      int W, RecMeta* __restrict__ meta, uint64_t* __restrict__ pwords,
                                                      Item<2>* __restrict__ pairs, uint32_t* __restrict__ stats) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 2 * n_lines) return;
    const uint32_t line = r >> 1;
    const bool minus = r & 1u;
    const uint64_t a0 = off[2 * line], a1 = off[2 * line + 1], a2 = off[2 * line + 2];
    const uint8_t* pr1 = blob + a0;
    const uint8_t* pr2 = blob + a1;
    const int n1 = (int)(a1 - a0), n2 = (int)(a2 - a1);
    RecMeta m;
    m.pcr_size = pcr[line];
    m.p1_word = word_off[2 * r];
    m.p2_word = word_off[2 * r + 1];
    m.pad = 0;
    uint32_t hbe = 0;
    int ho;
    if (!minus) {
        m.len1 = (uint16_t)n1; m.len2 = (uint16_t)n2;
        ho = first_clean_word(BlobFwd{pr1}, n1, W, &hbe);
        encode_primer(BlobFwd{pr1}, n1, plut, pwords + m.p1_word);
        encode_primer(BlobFwd{pr2}, n2, plut, pwords + m.p2_word);
    } else {
        m.len1 = (uint16_t)n2; m.len2 = (uint16_t)n1;
        ho = first_clean_word(BlobFwd{pr2}, n2, W, &hbe);
        encode_primer(BlobFwd{pr2}, n2, plut, pwords + m.p1_word);
        encode_primer(BlobRc{pr1, n1}, n1, plut, pwords + m.p2_word);
    }
    m.hash_be = hbe;
    m.key = reverse_digits(hbe, W);
    m.hash_off = (uint16_t)(ho < 0 ? 0 : ho);
    m.flags = ho >= 0 ? 1u : 0u;
    meta[r] = m;
    pairs[r].f[0] = m.key;
    pairs[r].f[1] = r | (ho >= 0 ? 0u : 0x80000000u);
    if (ho >= 0) {
        atomicAdd(&stats[0], 1u);
        atomicMax(&stats[1], (uint32_t)ho);
    }
}

// After the pairs are sorted by (invalid, key, record): CSR bucket array + open-addressed slot table + filter.
__global__ void __launch_bounds__(256) build_buckets(const Item<2>* __restrict__ pairs, uint32_t n_valid,
                                                     uint32_t* __restrict__ bucket, uint64_t* __restrict__ slots,
                                                     uint32_t slot_mask, uint32_t* __restrict__ filter,
                                                     uint32_t filter_bits, int filter_exact) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    const uint32_t key = pairs[i].f[0];
    const bool head = (i == 0) || (pairs[i - 1].f[0] != key);
    const bool last = (i + 1 == n_valid) || (pairs[i + 1].f[0] != key);
    bucket[i] = (pairs[i].f[1] & 0x7FFFFFFFu) | (last ? 0x80000000u : 0u);
    if (head) {
        const unsigned long long val = ((unsigned long long)key << 32) | i;
        uint32_t s = slot_hash(key) & slot_mask;
        for (;;) {
            unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(slots + s), ~0ull, val);
            if (prev == ~0ull) break;
            s = (s + 1) & slot_mask;
        }
        const uint32_t fb = filter_index(key, filter_bits, filter_exact);
        atomicOr(&filter[fb >> 5], 1u << (fb & 31));
    }
}

// ---------------------------------------------------------------------------------------------------------
// (3)(4)(5) scan
// ---------------------------------------------------------------------------------------------------------
struct ScanArgs {
    const uint64_t* p2;
    const uint64_t* p4;
    const uint64_t* valid;
    const TileDesc* tiles;
    uint32_t n_tiles;
    const uint64_t* slots;
    uint32_t slot_mask;
    const uint32_t* bucket;
    const RecMeta* meta;
    const uint64_t* pwords;
    const uint32_t* filter;
    uint32_t filter_bits, filter_words;
    int filter_exact;
    SearchParams prm;
    mpcr_hit* hits;
    unsigned long long capacity;
    unsigned long long* count;
    uint32_t* tile_counter;
    int debug;  // MPCR_DEBUG bit0: hash/filter phase only (candidates are counted, not verified)
};

struct HitEmitter {
    const ScanArgs& a;
    uint32_t contig, rec, hash_off;
    __device__ void operator()(int64_t pos1, int64_t pos2, uint32_t rank) const {
        unsigned long long i = atomicAdd(a.count, 1ull);
        if (i < a.capacity) {
            mpcr_hit h;
            h.contig = contig; h.pos1 = (uint32_t)pos1; h.pos2 = (uint32_t)pos2; h.rec = rec; h.rank = rank;
            h.hash_off = hash_off;
            a.hits[i] = h;
        }
    }
};

// engine.py:483-489: probe the table with the key at tile-local offset lp and dispatch every bucket entry.
__device__ __noinline__ void process_candidate(const ScanArgs& a, const TileDesc& td, uint32_t lp, uint32_t wmask) {
    const int64_t gpos = td.gbase + lp;
    const uint32_t key = extract_key(a.p2, gpos, wmask);
    uint32_t i = find_bucket(a.slots, a.slot_mask, key);
    if (i == kEmptySlot) return;
    const int64_t gcontig = td.gbase - (int64_t)td.lstart;
    const int64_t p = (int64_t)td.lstart + lp;
    for (;;) {
        const uint32_t e = a.bucket[i];
        const uint32_t rec = e & 0x7FFFFFFFu;
        const RecMeta m = a.meta[rec];
        verify_record(a.p4, gcontig, (int64_t)td.length, p, m, a.pwords, a.prm,
                      HitEmitter{a, td.contig, rec, (uint32_t)m.hash_off});
        if (e & 0x80000000u) break;
        ++i;
    }
}

// Persistent CTAs; each pulls 32 768-base tiles from a global counter.  Thread t owns the 64 hash positions
// [64t, 64t+64) of the tile: one 128-bit load of plane2 (+ the next word for the W-1 overhang), rolling keys by
// funnel shift, one shared-memory filter probe per position, then the surviving positions are verified.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) scan_kernel(const ScanArgs a) {
    extern __shared__ uint32_t s_filter[];
    __shared__ uint32_t s_tile;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.filter);
        uint4* dst = reinterpret_cast<uint4*>(s_filter);
        for (uint32_t i = threadIdx.x; i < a.filter_words / 4; i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t wmask = wmask_of(a.prm.W);
    const uint32_t fbits = a.filter_bits;
    const int fexact = a.filter_exact;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        __syncthreads();
        if (t >= a.n_tiles) break;
        const TileDesc td = a.tiles[t];
        const uint32_t lp0 = threadIdx.x * 64u;
        if (lp0 >= td.nbases) continue;
        const int64_t gb = td.gbase + lp0;
        const uint64_t* vw = a.valid + (gb >> 6);
        const uint64_t wv = window_valid(vw[0], vw[1], a.prm.W);
        if (wv == 0) continue;
        const uint64_t* w = a.p2 + (gb >> 5);
        const uint4 q = *reinterpret_cast<const uint4*>(w);
        const uint64_t w2 = w[2];
        const uint32_t r[6] = {q.x, q.y, q.z, q.w, (uint32_t)w2, (uint32_t)(w2 >> 32)};
        uint32_t pass_lo = 0, pass_hi = 0;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const uint32_t key = __funnelshift_r(r[(2 * j) >> 5], r[((2 * j) >> 5) + 1], (2 * j) & 31) & wmask;
            const uint32_t fb = filter_index(key, fbits, fexact);
            const uint32_t bit = (s_filter[fb >> 5] >> (fb & 31)) & 1u;
            if (j < 32) pass_lo |= bit << j; else pass_hi |= bit << (j - 32);
        }
        uint64_t cand = wv & (((uint64_t)pass_hi << 32) | pass_lo);
        if (a.debug & 1) {
            if (cand) atomicAdd(a.count, (unsigned long long)__popcll(cand));
            continue;
        }
        while (cand) {
            const int j = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            process_candidate(a, td, lp0 + (uint32_t)j, wmask);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int mpcr_abi_version(void) { return MPCR_ABI_VERSION; }
const char* mpcr_last_error(void) { return g_err; }

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
    CU(cudaSetDevice(device));
    mpcr_ctx* c = new mpcr_ctx();
    c->device = device;
    c->prm = *p;
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU(cudaMalloc(&c->d_tile_counter, 256));
    CU(cudaMalloc(&c->d_lut, 256));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    *out = c;
    return MPCR_OK;
}

static void free_table(mpcr_ctx* c) {
    cudaFree(c->d_meta); cudaFree(c->d_pwords); cudaFree(c->d_slots); cudaFree(c->d_bucket); cudaFree(c->d_filter);
    c->d_meta = nullptr; c->d_pwords = nullptr; c->d_slots = nullptr; c->d_bucket = nullptr; c->d_filter = nullptr;
    c->table_ready = false;
}

void mpcr_ctx_destroy(mpcr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_table(c);
    cudaFree(c->d_tiles); cudaFree(c->d_tile_counter); cudaFree(c->d_sort_tmp); cudaFree(c->d_counts); cudaFree(c->d_lut);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
}

int mpcr_ctx_sm_count(const mpcr_ctx* c) { return c ? c->sm_count : 0; }
uint64_t mpcr_launch_count(const mpcr_ctx* c) { return c ? c->launches : 0; }

int mpcr_pack_sequence(mpcr_ctx* c, const uint8_t* d_ascii, uint64_t n, uint64_t dst_base, uint64_t plane_origin,
                       void* d_plane2, void* d_plane4, void* d_valid, const uint8_t* h_lut, void* stream) {
    if (!c || !d_plane2 || !d_plane4 || !d_valid || !h_lut) return fail(MPCR_EINVAL, "null argument");
    if ((dst_base & 63u) || (plane_origin & 127u) || dst_base < plane_origin)
        return fail(MPCR_EINVAL, "dst_base must be a multiple of 64 and >= plane_origin (multiple of 128)");
    if (n == 0) return MPCR_OK;
    if (!d_ascii) return fail(MPCR_EINVAL, "null sequence pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_lut, h_lut, 256, cudaMemcpyHostToDevice, st));
    const uint64_t strips = (n + 63) / 64;
    const uint32_t blocks = (uint32_t)((strips + 255) / 256);
    pack_kernel<<<blocks, 256, 0, st>>>(d_ascii, n, dst_base - plane_origin, (uint64_t*)d_plane2, (uint64_t*)d_plane4,
                                        (uint64_t*)d_valid, c->d_lut);
    c->launches++;
    CU(cudaGetLastError());
    return MPCR_OK;
}

int mpcr_table_build(mpcr_ctx* c, const uint8_t* h_blob, const uint64_t* h_off, const uint32_t* h_pcr, uint32_t n_lines,
                     const uint8_t* h_plut, void* stream) {
    if (!c || !h_plut) return fail(MPCR_EINVAL, "null argument");
    if (n_lines && (!h_blob || !h_off || !h_pcr)) return fail(MPCR_EINVAL, "null argument");
    if (n_lines >= (1u << 30)) return fail(MPCR_EINVAL, "too many STS lines");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    free_table(c);
    const int W = c->prm.wordsize;
    const uint32_t n_rec = 2 * n_lines;
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

    // first-level filter geometry: exact bitmap if it fits the shared-memory budget, else a fold
    const uint64_t space = 1ull << (2 * W);
    uint32_t budget_bits = 1u << 20;  // 128 KiB of shared memory
    if (const char* env = getenv("MPCR_FILTER_BITS")) {
        long v = atol(env);
        if (v >= 128) budget_bits = (uint32_t)v;
    }
    const uint32_t max_bits = (uint32_t)(((size_t)c->max_smem_optin - 2048) * 8);
    if (budget_bits > max_bits) budget_bits = max_bits;
    budget_bits &= ~127u;
    if (space <= budget_bits) { c->filter_exact = 1; c->filter_bits = (uint32_t)(space < 128 ? 128 : space); }
    else { c->filter_exact = 0; c->filter_bits = budget_bits; }
    c->filter_words = c->filter_bits / 32;

    CU(cudaMalloc(&c->d_filter, (size_t)c->filter_words * 4));
    CU(cudaMemsetAsync(c->d_filter, 0, (size_t)c->filter_words * 4, st));
    if (n_lines == 0) {
        c->slot_mask = 1023;
        CU(cudaMalloc(&c->d_slots, 1024 * sizeof(uint64_t)));
        CU(cudaMemsetAsync(c->d_slots, 0xFF, 1024 * sizeof(uint64_t), st));
        CU(cudaStreamSynchronize(st));
        c->table_ready = true;
        return MPCR_OK;
    }
    const size_t blob_bytes = (size_t)h_off[2 * (size_t)n_lines];
    uint8_t *d_blob = nullptr, *d_plut = nullptr;
    uint64_t* d_off = nullptr;
    uint32_t *d_pcr = nullptr, *d_woff = nullptr, *d_stats = nullptr;
    Item<2>*d_pairs = nullptr, *d_pairs2 = nullptr;
    int rc = MPCR_OK;
#define CUG(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            rc = fail(MPCR_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            goto done;                                                                                   \
        }                                                                                                \
    } while (0)
    {
        CUG(cudaMalloc(&d_blob, blob_bytes + 16));
        CUG(cudaMalloc(&d_plut, 256));
        CUG(cudaMalloc(&d_off, (2 * (size_t)n_lines + 1) * sizeof(uint64_t)));
        CUG(cudaMalloc(&d_pcr, (size_t)n_lines * 4));
        CUG(cudaMalloc(&d_woff, word_off.size() * 4));
        CUG(cudaMalloc(&d_stats, 16));
        CUG(cudaMalloc(&d_pairs, (size_t)n_rec * sizeof(Item<2>)));
        CUG(cudaMalloc(&d_pairs2, (size_t)n_rec * sizeof(Item<2>)));
        CUG(cudaMalloc(&c->d_meta, (size_t)n_rec * sizeof(RecMeta)));
        CUG(cudaMalloc(&c->d_pwords, (size_t)(acc + 2) * sizeof(uint64_t)));
        CUG(cudaMemcpyAsync(d_blob, h_blob, blob_bytes, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_plut, h_plut, 256, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_off, h_off, (2 * (size_t)n_lines + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_pcr, h_pcr, (size_t)n_lines * 4, cudaMemcpyHostToDevice, st));
        CUG(cudaMemcpyAsync(d_woff, word_off.data(), word_off.size() * 4, cudaMemcpyHostToDevice, st));
        CUG(cudaMemsetAsync(d_stats, 0, 16, st));
        encode_records<<<(n_rec + 127) / 128, 128, 0, st>>>(d_blob, d_off, d_pcr, n_lines, d_plut, d_woff, W, c->d_meta,
                                                             c->d_pwords, d_pairs, d_stats);
        c->launches++;
        CUG(cudaGetLastError());
        uint32_t stats[4] = {0, 0, 0, 0};
        CUG(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
        CUG(cudaStreamSynchronize(st));
        c->n_valid = stats[0];
        c->max_hash_off = stats[1];
        // stable sort by (invalid flag, key): LSD passes over the key digits, then the flag bit
        PassDesc passes[8];
        int np = add_passes(passes, 0, 0, wmask_of(W));
        passes[np].field = 1; passes[np].shift = 31; passes[np].mask = 1; ++np;
        const uint32_t nblk = (n_rec + kSortItemsPerBlock - 1) / kSortItemsPerBlock;
        rc = ensure((void**)&c->d_counts, &c->counts_cap, (size_t)256 * nblk * 4);
        if (rc) goto done;
        c->launches += radix_sort<2>(d_pairs, d_pairs2, n_rec, passes, np, c->d_counts, st);
        CUG(cudaGetLastError());
        uint32_t nslots = 1024;
        while (nslots < 2u * c->n_valid + 2u) nslots <<= 1;
        c->slot_mask = nslots - 1;
        CUG(cudaMalloc(&c->d_slots, (size_t)nslots * sizeof(uint64_t)));
        CUG(cudaMemsetAsync(c->d_slots, 0xFF, (size_t)nslots * sizeof(uint64_t), st));
        CUG(cudaMalloc(&c->d_bucket, ((size_t)c->n_valid + 1) * 4));
        if (c->n_valid) {
            build_buckets<<<(c->n_valid + 255) / 256, 256, 0, st>>>(d_pairs, c->n_valid, c->d_bucket, c->d_slots,
                                                                     c->slot_mask, c->d_filter, c->filter_bits,
                                                                     c->filter_exact);
            c->launches++;
            CUG(cudaGetLastError());
        }
        CUG(cudaStreamSynchronize(st));
        c->table_ready = true;
    }
done:
    cudaFree(d_blob); cudaFree(d_plut); cudaFree(d_off); cudaFree(d_pcr); cudaFree(d_woff); cudaFree(d_stats);
    cudaFree(d_pairs); cudaFree(d_pairs2);
    if (rc) free_table(c);
    return rc;
#undef CUG
}

int mpcr_table_records(mpcr_ctx* c, int32_t* h_hash_offset, uint32_t* h_hash) {
    if (!c || !c->table_ready) return fail(MPCR_ESTATE, "table not built");
    if (c->n_rec == 0) return MPCR_OK;
    CU(cudaSetDevice(c->device));
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
    CU(cudaSetDevice(c->device));
    RecMeta m;
    CU(cudaMemcpy(&m, c->d_meta + rec, sizeof m, cudaMemcpyDeviceToHost));
    const uint32_t len = which == 1 ? m.len1 : m.len2, nw = 2 * ((len + 15) / 16);
    if (n_words) *n_words = nw;
    if (nw > max_words) return fail(MPCR_EINVAL, "buffer too small");
    CU(cudaMemcpy(h_words, c->d_pwords + (which == 1 ? m.p1_word : m.p2_word), (size_t)nw * 8, cudaMemcpyDeviceToHost));
    return MPCR_OK;
}

static uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

uint64_t mpcr_halo_left(const mpcr_ctx* c) { return c ? round_up((uint64_t)c->max_hash_off + 64, 128) : 0; }
uint64_t mpcr_halo_right(const mpcr_ctx* c) {
    if (!c) return 0;
    // furthest base a position can touch: p - hash_off + pcr_size + margin (engine.py:543,582) and the
    // W-1 / primer overhang of the last position; a tile is owned by the shard holding its FIRST base, so the
    // last tile may run up to kTileBases past shard_end
    return round_up(c->max_pcr + (uint64_t)c->prm.margin + c->max_len + 64 + 128, 128) + kTileBases;
}

static int build_tiles(mpcr_ctx* c, const mpcr_contig* contigs, uint32_t n_contigs, uint64_t origin, uint64_t sb,
                       uint64_t se, cudaStream_t st) {
    // signature of the layout: rebuild the descriptor array only when it changes
    uint64_t sig = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    mix(n_contigs); mix(origin); mix(sb); mix(se); mix((uint64_t)c->prm.wordsize);
    for (uint32_t i = 0; i < n_contigs; ++i) { mix(contigs[i].gstart); mix(contigs[i].length); }
    if (sig == c->tiles_sig && c->d_tiles) return MPCR_OK;
    std::vector<TileDesc> tiles;
    for (uint32_t i = 0; i < n_contigs; ++i) {
        const uint64_t L = contigs[i].length, g0 = contigs[i].gstart;
        if (L <= (uint64_t)c->prm.wordsize) continue;  // engine.py:458 (Q3: len <= W is skipped)
        if (g0 & 127u) return fail(MPCR_EINVAL, "contig %u: gstart not a multiple of 128", i);
        if (L >= (1ull << 31)) return fail(MPCR_EINVAL, "contig %u longer than 2^31-1 bases", i);
        for (uint64_t ls = 0; ls < L; ls += kTileBases) {
            const uint64_t g = g0 + ls;
            if (g < sb || g >= se) continue;  // tile ownership by first base (tiles never straddle shards)
            TileDesc t;
            t.gbase = (int64_t)(g - origin);
            t.contig = i;
            t.lstart = (uint32_t)ls;
            t.length = (uint32_t)L;
            t.nbases = (uint32_t)(L - ls < (uint64_t)kTileBases ? L - ls : (uint64_t)kTileBases);
            tiles.push_back(t);
        }
    }
    c->n_tiles = (uint32_t)tiles.size();
    c->lay_contigs = n_contigs;
    c->lay_max_len = 0;
    for (uint32_t i = 0; i < n_contigs; ++i)
        if (contigs[i].length > c->lay_max_len) c->lay_max_len = contigs[i].length;
    int rc = ensure((void**)&c->d_tiles, &c->tiles_cap, (tiles.size() + 1) * sizeof(TileDesc));
    if (rc) return rc;
    if (!tiles.empty())
        CU(cudaMemcpyAsync(c->d_tiles, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // `tiles` is pageable host memory going out of scope
    c->tiles_sig = sig;
    return MPCR_OK;
}

static constexpr int kScanThreads = kTileBases / 64;  // 512

int mpcr_scan(mpcr_ctx* c, const mpcr_contig* h_contigs, uint32_t n_contigs, const void* d_plane2, const void* d_plane4,
              const void* d_valid, uint64_t plane_origin, uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end,
              mpcr_hit* d_hits, uint64_t capacity, uint64_t* d_count, void* stream) {
    if (!c || !d_count) return fail(MPCR_EINVAL, "null argument");
    if (!c->table_ready) return fail(MPCR_ESTATE, "mpcr_scan called before mpcr_table_build");
    if (n_contigs && (!h_contigs || !d_plane2 || !d_plane4 || !d_valid)) return fail(MPCR_EINVAL, "null argument");
    if ((plane_origin & 127u) || (shard_begin & 127u)) return fail(MPCR_EINVAL, "origin / shard_begin must be multiples of 128");
    if (capacity && !d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    (void)plane_bases;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    int rc = build_tiles(c, h_contigs, n_contigs, plane_origin, shard_begin, shard_end, st);
    if (rc) return rc;
    CU(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    c->scan_timed = false;
    if (c->n_tiles == 0 || c->n_valid == 0) return MPCR_OK;
    CU(cudaMemsetAsync(c->d_tile_counter, 0, 4, st));
    ScanArgs a;
    a.p2 = (const uint64_t*)d_plane2; a.p4 = (const uint64_t*)d_plane4; a.valid = (const uint64_t*)d_valid;
    a.tiles = c->d_tiles; a.n_tiles = c->n_tiles;
    a.slots = c->d_slots; a.slot_mask = c->slot_mask; a.bucket = c->d_bucket; a.meta = c->d_meta; a.pwords = c->d_pwords;
    a.filter = c->d_filter; a.filter_bits = c->filter_bits; a.filter_words = c->filter_words; a.filter_exact = c->filter_exact;
    a.prm.W = c->prm.wordsize; a.prm.M = c->prm.margin; a.prm.N = c->prm.mismatches; a.prm.X = c->prm.three_prime_match;
    a.prm.iupac = c->prm.iupac_mode ? 1 : 0;
    a.debug = getenv("MPCR_DEBUG") ? atoi(getenv("MPCR_DEBUG")) : 0;
    a.hits = d_hits; a.capacity = capacity; a.count = (unsigned long long*)d_count; a.tile_counter = c->d_tile_counter;
    const size_t smem = (size_t)c->filter_words * 4;
    CU(cudaFuncSetAttribute(scan_kernel<kScanThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t grid = (uint32_t)c->sm_count;
    if (grid > c->n_tiles) grid = c->n_tiles;
    CU(cudaEventRecord(c->ev0, st));
    scan_kernel<kScanThreads><<<grid, kScanThreads, smem, st>>>(a);
    CU(cudaEventRecord(c->ev1, st));
    c->launches++;
    c->scan_timed = true;
    CU(cudaGetLastError());
    return MPCR_OK;
}

float mpcr_last_scan_ms(mpcr_ctx* c) {
    if (!c || !c->scan_timed) return 0.f;
    float ms = 0.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return 0.f;
    return ms;
}

int mpcr_sort_hits(mpcr_ctx* c, mpcr_hit* d_hits, uint64_t n, void* stream) {
    if (!c) return fail(MPCR_EINVAL, "null argument");
    if (n < 2) return MPCR_OK;
    if (!d_hits) return fail(MPCR_EINVAL, "null hit buffer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    static_assert(sizeof(mpcr_hit) == sizeof(Item<6>), "hit layout");
    int rc = ensure(&c->d_sort_tmp, &c->sort_tmp_cap, n * sizeof(mpcr_hit));
    if (rc) return rc;
    const uint64_t nblk = (n + kSortItemsPerBlock - 1) / kSortItemsPerBlock;
    rc = ensure((void**)&c->d_counts, &c->counts_cap, (size_t)256 * nblk * 4);
    if (rc) return rc;
    // LSD order: rank, rec, hash_off, pos1, contig  (fields 4, 3, 5, 1, 0); digits bounded by what can occur
    PassDesc passes[24];
    int np = 0;
    np = add_passes(passes, np, 4, 2ull * (uint64_t)c->prm.margin);
    np = add_passes(passes, np, 3, c->n_rec ? c->n_rec - 1 : 0);
    np = add_passes(passes, np, 5, c->max_hash_off);
    // pos1 / contig digits are bounded by the layout of the last scan
    np = add_passes(passes, np, 1, c->lay_max_len ? c->lay_max_len : 0x7FFFFFFFull);
    np = add_passes(passes, np, 0, c->lay_contigs ? c->lay_contigs - 1 : 0xFFFFFFFFull);
    c->launches += radix_sort<6>((Item<6>*)d_hits, (Item<6>*)c->d_sort_tmp, n, passes, np, c->d_counts, st);
    CU(cudaGetLastError());
    return MPCR_OK;
}

}  // extern "C"
