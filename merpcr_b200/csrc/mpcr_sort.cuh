// mpcr_sort.cuh -- stable LSD radix sort of small fixed-size records (hits, table pairs) on the device.
//
// Hand-written (no CUB): three kernels per 8-bit digit pass -- per-block histogram, single-block exclusive
// scan, stable scatter (warp match_any ranking inside a block).  Passes are described by (field, shift,
// mask) so that callers skip digits that cannot vary (known value bounds), which is what keeps the hit sort
// at ~9 passes for a human-sized genome.  Replaces `hits.sort(key=pos1)` (core/engine.py:434) together with
// the discovery-order tie rule.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcr {

struct PassDesc {
    uint8_t field;   // index of the uint32 field inside the record
    uint8_t shift;
    uint16_t mask;   // <= 255
};

template <int NF>
struct Item {
    uint32_t f[NF];
};

static constexpr int kSortThreads = 256;
static constexpr int kSortItemsPerBlock = 2048;

// Record counts may live on the device: the scan leaves the number of hits in *d_n, and the sort is queued right
// behind it without a host round trip.  d_n == nullptr: the count is n_host; otherwise min(*d_n, n_host), with
// n_host the capacity of the buffer (the grids are sized for it, CTAs past the count find nothing to do).
__device__ __forceinline__ uint64_t sort_count(const unsigned long long* d_n, uint64_t n_host) {
    if (!d_n) return n_host;
    const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(d_n);
    return v < n_host ? v : n_host;
}

template <int NF>
__global__ void __launch_bounds__(kSortThreads) rs_hist(const Item<NF>* __restrict__ in, uint64_t n_host,
                                                        const unsigned long long* d_n, PassDesc pd,
                                                        uint32_t* __restrict__ counts, uint32_t nblk) {
    __shared__ uint32_t h[256];
    const uint64_t n = sort_count(d_n, n_host);
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortItemsPerBlock;
    for (int i = threadIdx.x; i < kSortItemsPerBlock; i += kSortThreads) {
        uint64_t idx = base + i;
        if (idx < n) atomicAdd(&h[(in[idx].f[pd.field] >> pd.shift) & pd.mask], 1u);
    }
    __syncthreads();
    counts[(uint64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `total` uint32 in place, one block of 1024 threads
__global__ void __launch_bounds__(1024) rs_scan(uint32_t* __restrict__ a, uint32_t total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t per = (total + 1023u) / 1024u;
    const uint32_t lo = min(total, (uint32_t)tid * per), hi = min(total, lo + per);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += a[i];
    // block exclusive scan of s
    uint32_t incl = s;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_sums[lane], wi = w;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        warp_sums[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - s;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t v = a[i];
        a[i] = run;
        run += v;
    }
}

template <int NF>
__global__ void __launch_bounds__(kSortThreads) rs_scatter(const Item<NF>* __restrict__ in, Item<NF>* __restrict__ out,
                                                           uint64_t n_host, const unsigned long long* d_n, PassDesc pd,
                                                           const uint32_t* __restrict__ offsets, uint32_t nblk) {
    const uint64_t n = sort_count(d_n, n_host);
    __shared__ uint32_t running[256];
    __shared__ uint32_t wc[kSortThreads / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    running[tid] = offsets[(uint64_t)tid * nblk + blockIdx.x];
    for (int w = 0; w < kSortThreads / 32; ++w) wc[w][tid] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortItemsPerBlock;
    for (int c = 0; c < kSortItemsPerBlock; c += kSortThreads) {
        const uint64_t idx = base + c + tid;
        const bool act = idx < n;
        Item<NF> it;
        uint32_t d = 0x100u + (uint32_t)lane;  // inactive lanes never match an active digit
        if (act) { it = in[idx]; d = (it.f[pd.field] >> pd.shift) & pd.mask; }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (act && rank == 0) wc[wid][d] = __popc(peers);
        __syncthreads();
        if (act) {
            uint32_t pre = 0;
            for (int w = 0; w < wid; ++w) pre += wc[w][d];
            out[running[d] + pre + rank] = it;
        }
        __syncthreads();
        {
            uint32_t tot = 0;
            for (int w = 0; w < kSortThreads / 32; ++w) { tot += wc[w][tid]; wc[w][tid] = 0; }
            running[tid] += tot;
        }
        __syncthreads();
    }
}

// All passes in ONE cooperative launch while the grid is co-resident (n_blk <= kFusedMaxBlocks chunks of 2048 records, one CTA
// each): per pass a block histogram, a grid barrier, every CTA derives its own 256 scatter offsets from the
// histograms of all CTAs, the same stable match_any scatter as rs_scatter, a grid barrier.  This is what orders
// the ~10^5 hits of a human-sized scan: 9 passes in one launch instead of 27 launches.
static constexpr int kFusedMaxBlocks = 1024;   // tried cooperatively; a grid that is not co-resident falls back below
static constexpr int kMaxPasses = 24;
struct PassList {
    PassDesc p[kMaxPasses];
    int n;
};

template <int NF>
__global__ void __launch_bounds__(kSortThreads) rs_sort_fused(Item<NF>* a, Item<NF>* b, uint32_t n_host,
                                                              const unsigned long long* d_n, uint32_t skip_upto,
                                                              const uint32_t* __restrict__ skip_off, PassList pl,
                                                              uint32_t* __restrict__ counts /* 256 * gridDim.x */) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const uint32_t n = (uint32_t)sort_count(d_n, n_host);
    // lists that short were bucket-sorted, unless that sort raised its fall-back flag (the whole grid branches together)
    if (n <= skip_upto && !(skip_off && *reinterpret_cast<const volatile uint32_t*>(skip_off))) return;
    __shared__ uint32_t hist[256];
    __shared__ uint32_t running[256];
    __shared__ uint32_t wc[kSortThreads / 32][256];
    __shared__ uint32_t warp_sums[kSortThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nblk = gridDim.x, blk = blockIdx.x;
    const uint32_t base = blk * kSortItemsPerBlock;
    Item<NF>*src = a, *dst = b;
    for (int p = 0; p < pl.n; ++p) {
        const PassDesc pd = pl.p[p];
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < kSortItemsPerBlock; i += kSortThreads) {
            const uint32_t idx = base + i;
            if (idx < n) atomicAdd(&hist[(src[idx].f[pd.field] >> pd.shift) & pd.mask], 1u);
        }
        __syncthreads();
        counts[(uint32_t)tid * nblk + blk] = hist[tid];
        __threadfence();
        grid.sync();
        // offset of (digit tid, this CTA) = records with a smaller digit + records with this digit in earlier CTAs
        uint32_t total = 0, before = 0;
        for (uint32_t bb = 0; bb < nblk; ++bb) {
            const uint32_t v = counts[(uint32_t)tid * nblk + bb];
            total += v;
            if (bb < blk) before += v;
        }
        uint32_t incl = total;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < wid; ++w) wbase += warp_sums[w];
        running[tid] = wbase + incl - total + before;
        for (int w = 0; w < kSortThreads / 32; ++w) wc[w][tid] = 0;
        __syncthreads();
        for (int c = 0; c < kSortItemsPerBlock; c += kSortThreads) {
            const uint32_t idx = base + c + tid;
            const bool act = idx < n;
            Item<NF> it;
            uint32_t d = 0x100u + (uint32_t)lane;  // inactive lanes never match an active digit
            if (act) { it = src[idx]; d = (it.f[pd.field] >> pd.shift) & pd.mask; }
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            if (act && rank == 0) wc[wid][d] = __popc(peers);
            __syncthreads();
            if (act) {
                uint32_t pre = 0;
                for (int w = 0; w < wid; ++w) pre += wc[w][d];
                dst[running[d] + pre + rank] = it;
            }
            __syncthreads();
            {
                uint32_t tot = 0;
                for (int w = 0; w < kSortThreads / 32; ++w) { tot += wc[w][tid]; wc[w][tid] = 0; }
                running[tid] += tot;
            }
            __syncthreads();
        }
        __threadfence();
        grid.sync();
        Item<NF>* t = src; src = dst; dst = t;
    }
}

// Host driver: sorts the records of d_a using d_b as the ping-pong buffer.  n_host is the record count, or -- with
// d_n != nullptr -- the capacity the grids are sized for while the true count min(*d_n, n_host) is read on the device.
// The result is left in *result (d_a after an even number of passes, d_b after an odd one: no copy back; the caller's
// next kernel reads from there).  Lists of up to skip_upto records are left untouched unless *skip_off is set (see the
// bucket sort below).
// d_counts must hold 256 * ceil(n_host / kSortItemsPerBlock) uint32.  Returns the number of kernels launched.
template <int NF>
inline int radix_sort(Item<NF>* d_a, Item<NF>* d_b, uint64_t n_host, const unsigned long long* d_n, uint32_t skip_upto,
                      const uint32_t* skip_off, const PassDesc* passes, int npass, uint32_t* d_counts, cudaStream_t st,
                      Item<NF>** result, uint32_t* skipped_upto = nullptr) {
    *result = d_a;
    if (skipped_upto) *skipped_upto = skip_upto;
    if (n_host < 2 || npass == 0) return 0;
    *result = (npass & 1) ? d_b : d_a;
    const uint32_t nblk = (uint32_t)((n_host + kSortItemsPerBlock - 1) / kSortItemsPerBlock);
    if (nblk <= (uint32_t)kFusedMaxBlocks && npass <= kMaxPasses) {
        PassList pl;
        pl.n = npass;
        for (int p = 0; p < npass; ++p) pl.p[p] = passes[p];
        uint32_t n32 = (uint32_t)n_host;
        void* args[] = {&d_a, &d_b, &n32, &d_n, &skip_upto, &skip_off, &pl, &d_counts};
        if (cudaLaunchCooperativeKernel((const void*)rs_sort_fused<NF>, dim3(nblk), dim3(kSortThreads), args, 0, st) ==
            cudaSuccess)
            return 1;
        (void)cudaGetLastError();  // not launchable cooperatively here: fall through to the three-kernel passes
    }
    // the three-kernel passes sort short lists as well (a stable re-sort of what rank_sort_small ordered): tell the
    // caller that nothing was skipped, so that its next kernel picks the records up from *result
    if (skipped_upto) *skipped_upto = 0;
    Item<NF>*src = d_a, *dst = d_b;
    int launches = 0;
    for (int p = 0; p < npass; ++p) {
        rs_hist<NF><<<nblk, kSortThreads, 0, st>>>(src, n_host, d_n, passes[p], d_counts, nblk);
        rs_scan<<<1, 1024, 0, st>>>(d_counts, 256u * nblk);
        rs_scatter<NF><<<nblk, kSortThreads, 0, st>>>(src, dst, n_host, d_n, passes[p], d_counts, nblk);
        launches += 3;
        Item<NF>* t = src; src = dst; dst = t;
    }
    return launches;
}

// Hit lists of up to 2^17 records (a human-sized scan yields ~10^5 hits, a rank of an 8-GPU run ~10^4): a bucket sort
// on the hits' GLOBAL coordinate instead of radix passes.  Hits spread over the scanned range, so 2^14 equal slices of
// that range hold a handful of records each: count per slice (one atomic per record), one-CTA scan of the counts,
// scatter, and one thread per slice puts its few records into the complete order (contig, pos1, hash_off, rec, rank)
// -- no tie pass needed afterwards.  Four small launches, ~15 us, against ~110 us for five cooperative radix passes
// (a rank sort and a one-CTA shared-memory radix were measured first: 88 us and 161 us for 10^4 hits; the same bucket
// sort by ONE CTA in one launch, counters and offsets in shared memory: ~100 us -- a lone CTA pays every L2 round trip
// of its ten records per thread in sequence).  A list that
// piles up in one slice (> kBucketMaxFill records: repeats, an N-run in IUPAC mode) raises *fallback and is left to
// the radix passes, which check the flag on the device.
static constexpr uint32_t kBucketSortMax = 1u << 17;
static constexpr uint32_t kSortBuckets = 1u << 14;
static constexpr uint32_t kBucketMaxFill = 64;

struct BucketSortArgs {
    const unsigned long long* contig_g;   // global start coordinate of every contig
    uint32_t n_contigs;
    unsigned long long g_lo, span;        // scanned range [g_lo, g_lo + span)
    uint32_t* cnt;                        // kSortBuckets counters (zero on entry, zeroed again by bsort_finish)
    uint32_t* off;                        // kSortBuckets + 1 offsets
    uint32_t* slot;                       // per record: position inside its slice << 14 | slice
    uint32_t* fallback;                   // [0] 1: some slice is too full, the radix passes take over
};

__device__ __forceinline__ bool hit_less(const Item<6>& a, const Item<6>& b) {   // (contig, pos1, hash_off, rec, rank)
    if (a.f[0] != b.f[0]) return a.f[0] < b.f[0];
    if (a.f[1] != b.f[1]) return a.f[1] < b.f[1];
    if (a.f[5] != b.f[5]) return a.f[5] < b.f[5];
    if (a.f[3] != b.f[3]) return a.f[3] < b.f[3];
    return a.f[4] < b.f[4];
}

__global__ void __launch_bounds__(256) bsort_count(const Item<6>* __restrict__ in, const unsigned long long* d_n, uint64_t n_host,
                                                   BucketSortArgs b) {
    const uint64_t n = sort_count(d_n, n_host);
    if (n > kBucketSortMax) return;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t contig = in[i].f[0], pos1 = in[i].f[1];
    unsigned long long g = (contig < b.n_contigs ? b.contig_g[contig] : 0ull) + pos1;
    g = g > b.g_lo ? g - b.g_lo : 0ull;
    unsigned long long s = b.span ? (g * kSortBuckets) / b.span : 0ull;   // monotone in (contig, pos1)
    if (s >= kSortBuckets) s = kSortBuckets - 1;
    const uint32_t at = atomicAdd(&b.cnt[(uint32_t)s], 1u);
    b.slot[i] = (at << 14) | (uint32_t)s;
}

// exclusive scan of the kSortBuckets counters by one CTA of 1024 threads (16 consecutive counters each, read and written
// as four 16-byte vectors) + the too-full check
__global__ void __launch_bounds__(1024) bsort_scan(const unsigned long long* d_n, uint64_t n_host, BucketSortArgs b) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_max;
    const uint64_t n = sort_count(d_n, n_host);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_max = 0;
    __syncthreads();
    constexpr int kPer = kSortBuckets / 1024;
    static_assert(kPer == 16, "16 counters per thread");
    uint32_t v[kPer], sum = 0, mx = 0;
    const uint4* src = reinterpret_cast<const uint4*>(b.cnt) + tid * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 t = src[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) { sum += v[k]; mx = max(mx, v[k]); }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    if (lane == 0) atomicMax(&s_max, mx);
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        warp_sums[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - sum;
    uint4* dst = reinterpret_cast<uint4*>(b.off) + tid * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 t;
        t.x = run; run += v[4 * q];
        t.y = run; run += v[4 * q + 1];
        t.z = run; run += v[4 * q + 2];
        t.w = run; run += v[4 * q + 3];
        dst[q] = t;
    }
    if (tid == 1023) b.off[kSortBuckets] = run;
    if (tid == 0) b.fallback[0] = (n > kBucketSortMax || s_max > kBucketMaxFill) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) bsort_scatter(const Item<6>* __restrict__ in, Item<6>* __restrict__ out,
                                                     const unsigned long long* d_n, uint64_t n_host, BucketSortArgs b) {
    const uint64_t n = sort_count(d_n, n_host);
    if (n > kBucketSortMax || b.fallback[0]) return;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sl = b.slot[i];
    out[b.off[sl & (kSortBuckets - 1u)] + (sl >> 14)] = in[i];
}

// one thread per slice: its records (in `tmp`) go home in the complete order; the counters are zeroed for the next sort
__global__ void __launch_bounds__(256) bsort_finish(Item<6>* __restrict__ tmp, Item<6>* __restrict__ hits, BucketSortArgs b) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= kSortBuckets) return;
    const uint32_t m = b.cnt[s];
    b.cnt[s] = 0;
    if (m == 0 || b.fallback[0]) return;
    const uint32_t o = b.off[s];
    for (uint32_t k = 1; k < m; ++k) {   // insertion sort in the scratch buffer (a handful of records)
        const Item<6> x = tmp[o + k];
        uint32_t j = k;
        while (j > 0 && hit_less(x, tmp[o + j - 1])) { tmp[o + j] = tmp[o + j - 1]; --j; }
        tmp[o + j] = x;
    }
    for (uint32_t k = 0; k < m; ++k) hits[o + k] = tmp[o + k];
}

// append the 8-bit digit passes needed to cover values in [0, max_value] of `field`
inline int add_passes(PassDesc* out, int np, int field, uint64_t max_value) {
    int bits = 0;
    while (bits < 32 && (max_value >> bits)) ++bits;
    for (int s = 0; s < bits; s += 8) {
        int w = bits - s < 8 ? bits - s : 8;
        out[np].field = (uint8_t)field;
        out[np].shift = (uint8_t)s;
        out[np].mask = (uint16_t)((1u << w) - 1u);
        ++np;
    }
    return np;
}

}  // namespace mpcr
