// mpcr_fasta.cuh -- device-side FASTA text ingest: raw file bytes in HBM -> header table + the filtered sequence
// bytes of every record, contiguous and in file order.  Same observable rules as the reference's
// FASTALoader.load_file (io/fasta.py:43-66) for ASCII files:
//   * lines end at "\n", "\r" or "\r\n" (text-mode universal newlines) and are strip()ped before the '>' test, so
//     a header is a '>' that is the first non-whitespace character of its line; it runs to the end of that line;
//   * everything before the first header is discarded;
//   * every other character is kept iff it is one of ACGTBDHKMNRSVWXY in either case (case preserved) -- line
//     structure is irrelevant for sequence lines because no terminator or blank is in the keep set.
// Bytes >= 128 are reported to the host, which then parses the (rare) non-ASCII file itself with the locale rules.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcr {

static constexpr int kFastaBlock = 4096;    // text bytes per counting / compaction block
static constexpr int kFastaThreads = 256;   // 16 bytes per thread

__device__ __forceinline__ bool fasta_is_term(uint8_t c) { return c == 10 || c == 13; }
// str.strip() blanks among ASCII, terminators excluded
__device__ __forceinline__ bool fasta_is_blank(uint8_t c) { return c == 32 || c == 9 || c == 11 || c == 12 || (c >= 28 && c <= 31); }
__device__ __forceinline__ bool fasta_keep(uint8_t c) {
    // ACGTBDHKMNRSVWXY: bit (c & 31) of the mask, for letters only
    const uint32_t u = c & 0xDFu;  // upper-case
    if (u < 'A' || u > 'Z') return false;
    return (0x01EE34CFu >> (u - 'A')) & 1u;  // bit i = letter 'A' + i is one of ACGTBDHKMNRSVWXY
}

struct FastaHeader {
    uint64_t begin;  // byte offset of '>'
    uint64_t end;    // byte offset of the line terminator (or n)
};

// One thread per 16 bytes: report every header '>' and whether any byte is >= 128.
__global__ void __launch_bounds__(256) fasta_find_headers(const uint8_t* __restrict__ text, uint64_t n,
                                                          FastaHeader* __restrict__ out, uint32_t cap,
                                                          uint32_t* __restrict__ count, uint32_t* __restrict__ flags) {
    const uint64_t b0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (b0 >= n) return;
    uint8_t c[16];
    if (b0 + 16 <= n && ((reinterpret_cast<uintptr_t>(text) & 15u) == 0)) {
        *reinterpret_cast<uint4*>(c) = *reinterpret_cast<const uint4*>(text + b0);
    } else {
        for (int k = 0; k < 16; ++k) c[k] = b0 + k < n ? text[b0 + k] : (uint8_t)'\n';
    }
    bool high = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        high |= c[k] >= 128;
        if (c[k] != '>') continue;
        const uint64_t i = b0 + k;
        // first non-blank character of its line?
        uint64_t j = i;
        bool first = true;
        while (j > 0) {
            const uint8_t p = text[j - 1];
            if (fasta_is_term(p)) break;
            if (!fasta_is_blank(p)) { first = false; break; }
            --j;
        }
        if (!first) continue;
        uint64_t e = i + 1;
        while (e < n && !fasta_is_term(text[e])) ++e;
        const uint32_t slot = atomicAdd(count, 1u);
        if (slot < cap) out[slot] = FastaHeader{i, e};
    }
    if (high) atomicOr(flags, 1u);
}

// Slices of a file (rank-local ingest): where does the first line terminator within the first `limit` bytes end?
// out[0] = index of the byte after it ("\r\n" counts as one terminator), 0xFFFFFFFF if there is none.
__global__ void __launch_bounds__(256) fasta_first_line_end(const uint8_t* __restrict__ text, uint64_t n, uint32_t limit,
                                                            uint32_t* __restrict__ out) {
    const uint32_t m = (uint32_t)(n < limit ? n : limit);
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        if (fasta_is_term(text[i])) { best = i; break; }
    }
    atomicMin(out, best);
    __syncthreads();
    if (threadIdx.x == 0 && out[0] != 0xFFFFFFFFu) {
        const uint32_t t = out[0];
        out[0] = (text[t] == 13 && (uint64_t)t + 1 < n && text[t + 1] == 10) ? t + 2 : t + 1;
    }
}

// Blank the header lines (sorted table) and everything before the first header, so that "kept" becomes a pure
// per-byte property.  One thread per header; header lines are short.
__global__ void __launch_bounds__(128) fasta_blank_headers(uint8_t* __restrict__ text, const FastaHeader* __restrict__ hdr,
                                                           uint32_t n_hdr) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hdr) return;
    for (uint64_t i = hdr[h].begin; i < hdr[h].end; ++i) text[i] = '\n';
}

__device__ __forceinline__ uint32_t fasta_keep_mask16(const uint8_t* __restrict__ text, uint64_t b0, uint64_t n, uint8_t* c) {
    if (b0 + 16 <= n && ((reinterpret_cast<uintptr_t>(text) & 15u) == 0)) {
        *reinterpret_cast<uint4*>(c) = *reinterpret_cast<const uint4*>(text + b0);
    } else {
        for (int k = 0; k < 16; ++k) c[k] = b0 + k < n ? text[b0 + k] : (uint8_t)'\n';
    }
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) m |= (fasta_keep(c[k]) ? 1u : 0u) << k;
    return m;
}

// kept bytes per 4096-byte block
__global__ void __launch_bounds__(kFastaThreads) fasta_count(const uint8_t* __restrict__ text, uint64_t n,
                                                             uint32_t* __restrict__ block_count) {
    __shared__ uint32_t warp_sum[kFastaThreads / 32];
    const uint64_t b0 = (uint64_t)blockIdx.x * kFastaBlock + (uint64_t)threadIdx.x * 16;
    uint8_t c[16];
    uint32_t k = b0 < n ? __popc(fasta_keep_mask16(text, b0, n, c)) : 0u;
    for (int d = 16; d; d >>= 1) k += __shfl_xor_sync(0xffffffffu, k, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kFastaThreads / 32; ++w) t += warp_sum[w];
        block_count[blockIdx.x] = t;
    }
}

// exclusive prefix of n_blk uint32 counts into uint64 offsets (one CTA; n_blk is ~1e6 for a human genome)
__global__ void __launch_bounds__(1024) fasta_scan_blocks(const uint32_t* __restrict__ cnt, uint64_t n_blk,
                                                          uint64_t* __restrict__ off /* n_blk + 1 */) {
    __shared__ uint64_t part[1024];
    const uint64_t per = (n_blk + 1023) / 1024;
    const uint64_t lo = min(n_blk, (uint64_t)threadIdx.x * per), hi = min(n_blk, lo + per);
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += cnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int t = 0; t < 1024; ++t) { const uint64_t v = part[t]; part[t] = run; run += v; }
        off[n_blk] = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint64_t i = lo; i < hi; ++i) { off[i] = run; run += cnt[i]; }
}

// compaction: block b writes its kept bytes at out[off[b] ...] in order
__global__ void __launch_bounds__(kFastaThreads) fasta_compact(const uint8_t* __restrict__ text, uint64_t n,
                                                               const uint64_t* __restrict__ off, uint8_t* __restrict__ out) {
    __shared__ uint32_t warp_sum[kFastaThreads / 32];
    __shared__ uint8_t stage[kFastaBlock];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t b0 = (uint64_t)blockIdx.x * kFastaBlock + (uint64_t)tid * 16;
    uint8_t c[16];
    const uint32_t m = b0 < n ? fasta_keep_mask16(text, b0, n, c) : 0u;
    const uint32_t k = __popc(m);
    uint32_t incl = k;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (int w = 0; w < kFastaThreads / 32; ++w) {
        if (w < wid) base += warp_sum[w];
        total += warp_sum[w];
    }
    uint32_t p = base + incl - k;
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if ((m >> j) & 1u) stage[p++] = c[j];
    __syncthreads();
    uint8_t* dst = out + off[blockIdx.x];
    for (uint32_t i = tid; i < total; i += kFastaThreads) dst[i] = stage[i];
}

// kept bytes before text position pos[i] (positions inside blanked/ordinary text; one thread each)
__global__ void __launch_bounds__(128) fasta_offsets_at(const uint8_t* __restrict__ text, uint64_t n,
                                                        const uint64_t* __restrict__ off, const uint64_t* __restrict__ pos,
                                                        uint32_t n_pos, uint64_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    const uint64_t p = pos[i] < n ? pos[i] : n;
    const uint64_t blk = p / kFastaBlock;
    uint64_t r = off[blk];
    for (uint64_t j = blk * kFastaBlock; j < p; ++j) r += fasta_keep(text[j]) ? 1u : 0u;
    out[i] = r;
}

}  // namespace mpcr
