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

// ONE pass over the text instead of two (fasta_find_headers + fasta_count).  One WARP per 4096-byte counting block, eight
// blocks per CTA and no CTA-wide barrier (one CTA per block with a __syncthreads reduction spent its time on CTA launches:
// 1.3 * 10^5 CTAs of 4 KB each for a 512 Mbp file): the warp walks its block in eight steps of 512 bytes, every lane
// classifies 16 bytes, four at a time in registers (mpcr_core.cuh: fasta_keep_flags4), kept letters are counted, '>' bytes
// take the (rare) header test, bytes >= 128 raise the flag.  The counts still include the letters of header lines and of
// the text in front of the first header -- fasta_blank_headers / fasta_blank_range take those out again when they blank
// them.
static constexpr int kFastaWarps = kFastaThreads / 32;   // counting blocks per CTA
__device__ __forceinline__ void fasta_load16(const uint8_t* __restrict__ text, uint64_t n, uint64_t b0, bool aligned, uint32_t w[4]) {
    if (aligned && b0 + 16 <= n) {
        const uint4 t = *reinterpret_cast<const uint4*>(text + b0);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    } else {
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
            for (int b = 0; b < 4; ++b) x |= (uint32_t)(b0 + 4 * k + b < n ? text[b0 + 4 * k + b] : (uint8_t)'\n') << (8 * b);
            w[k] = x;
        }
    }
}
__global__ void __launch_bounds__(kFastaThreads) fasta_classify(const uint8_t* __restrict__ text, uint64_t n,
                                                                FastaHeader* __restrict__ out, uint32_t cap,
                                                                uint32_t* __restrict__ count, uint32_t* __restrict__ flags,
                                                                uint32_t* __restrict__ block_count) {
    const int lane = threadIdx.x & 31;
    const uint64_t blk = (uint64_t)blockIdx.x * kFastaWarps + (threadIdx.x >> 5);
    const uint64_t base = blk * kFastaBlock;
    if (base >= n) return;   // the whole warp
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
    uint32_t kept = 0, any = 0;
#pragma unroll
    for (int it = 0; it < kFastaBlock / 512; ++it) {
        const uint64_t b0 = base + (uint64_t)it * 512 + (uint64_t)lane * 16;
        if (b0 >= n) continue;
        uint32_t w[4];
        fasta_load16(text, n, b0, aligned, w);
        uint32_t gt = 0;   // bit j: byte j is '>'
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            kept += __popc(fasta_keep_flags4(w[k]));
            any |= w[k];
            if (bytes_equal_trigger4(w[k], 0x3E3E3E3Eu)) {   // rare: which bytes exactly
                for (int b = 0; b < 4; ++b)
                    if (((w[k] >> (8 * b)) & 0xFFu) == (uint32_t)'>') gt |= 1u << (4 * k + b);
            }
        }
        while (gt) {   // rare: is this '>' the first non-blank character of its line?
            const int j = __ffs(gt) - 1;
            gt &= gt - 1;
            const uint64_t i = b0 + j;
            uint64_t q = i;
            bool first = true;
            while (q > 0) {
                const uint8_t p = text[q - 1];
                if (fasta_is_term(p)) break;
                if (!fasta_is_blank(p)) { first = false; break; }
                --q;
            }
            if (!first) continue;
            uint64_t e = i + 1;
            while (e < n && !fasta_is_term(text[e])) ++e;
            const uint32_t slot = atomicAdd(count, 1u);
            if (slot < cap) out[slot] = FastaHeader{i, e};
        }
    }
    __syncwarp();
    const bool high = __any_sync(0xffffffffu, (any & 0x80808080u) != 0);
    for (int d = 16; d; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
    if (lane == 0) {
        if (high) atomicOr(flags, 1u);
        block_count[blk] = kept;
    }
}

// Blank the bytes [a, b) (the text in front of the first header) and take their kept letters out of the block counts.
__global__ void __launch_bounds__(256) fasta_blank_range(uint8_t* __restrict__ text, uint64_t a, uint64_t b,
                                                         uint32_t* __restrict__ block_count) {
    const uint64_t i0 = a + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    for (uint64_t i = i0; i < i0 + 16 && i < b; ++i) {
        if (fasta_keep(text[i])) atomicSub(&block_count[i / kFastaBlock], 1u);
        text[i] = '\n';
    }
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
                                                           uint32_t n_hdr, uint32_t* __restrict__ block_count) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hdr) return;
    for (uint64_t i = hdr[h].begin; i < hdr[h].end; ++i) {
        // block_count (fasta_classify) still counts the letters of this line: take them out
        if (block_count && fasta_keep(text[i])) atomicSub(&block_count[i / kFastaBlock], 1u);
        text[i] = '\n';
    }
}

// exclusive prefix of n_blk uint32 counts into uint64 offsets, in three small launches (n_blk is ~10^6 for a human
// genome: one CTA walking it serially took 0.12 ms): per chunk of 16 384 counts a sum, one CTA scans the chunk sums (up
// to 1024 chunks = 64 GiB of text), then every chunk scans itself from its base.
static constexpr int kScanChunk = 16384;
__global__ void __launch_bounds__(1024) fasta_scan_reduce(const uint32_t* __restrict__ cnt, uint64_t n_blk,
                                                          uint64_t* __restrict__ chunk_sum) {
    __shared__ uint64_t ws[32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanChunk;
    uint64_t s = 0;
    for (int k = 0; k < kScanChunk / 1024; ++k) {
        const uint64_t i = base + (uint64_t)k * 1024 + threadIdx.x;
        if (i < n_blk) s += cnt[i];
    }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < 32; ++w) t += ws[w];
        chunk_sum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) fasta_scan_top(uint64_t* __restrict__ chunk_sum, uint32_t n_chunks,
                                                       uint64_t* __restrict__ total) {
    __shared__ uint64_t ws[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t v = threadIdx.x < n_chunks ? chunk_sum[threadIdx.x] : 0ull;
    uint64_t incl = v;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint64_t w = ws[lane];
        uint64_t wi = w;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        ws[lane] = wi - w;
    }
    __syncthreads();
    if (threadIdx.x < n_chunks) chunk_sum[threadIdx.x] = ws[wid] + incl - v;   // exclusive
    if (threadIdx.x == 1023) *total = ws[wid] + incl;
}
__global__ void __launch_bounds__(1024) fasta_scan_down(const uint32_t* __restrict__ cnt, uint64_t n_blk,
                                                        const uint64_t* __restrict__ chunk_base, uint64_t* __restrict__ off) {
    __shared__ uint32_t ws[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int kPer = kScanChunk / 1024;   // 16 consecutive counts per thread
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * kPer;
    uint32_t v[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) { v[k] = i0 + k < n_blk ? cnt[i0 + k] : 0u; sum += v[k]; }
    uint32_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = ws[lane];
        uint32_t wi = w;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        ws[lane] = wi - w;
    }
    __syncthreads();
    uint64_t run = chunk_base[blockIdx.x] + ws[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (i0 + k < n_blk) off[i0 + k] = run;
        run += v[k];
    }
}

// compaction: block b writes its kept bytes at out[off[b] ...] in order.  One warp per 4096-byte block (eight per CTA, no
// CTA-wide barrier): eight steps of 512 bytes -- keep flags four bytes at a time, a warp scan of the lanes' counts, the kept
// bytes into the warp's staging row -- and the row goes out in 16-byte vectors: it is filled from index (address of
// out[off[b]]) mod 16, so staging row and destination are aligned with each other and only the first and last few bytes
// are single-byte stores.
__global__ void __launch_bounds__(kFastaThreads) fasta_compact(const uint8_t* __restrict__ text, uint64_t n,
                                                               const uint64_t* __restrict__ off, uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t stage_all[kFastaWarps][kFastaBlock + 16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t blk = (uint64_t)blockIdx.x * kFastaWarps + wid;
    const uint64_t base = blk * kFastaBlock;
    if (base >= n) return;   // the whole warp
    uint8_t* stage = stage_all[wid];
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
    uint8_t* const dst0 = out + off[blk];
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(dst0) & 15u);
    uint32_t run = skew;
#pragma unroll 2
    for (int it = 0; it < kFastaBlock / 512; ++it) {
        const uint64_t b0 = base + (uint64_t)it * 512 + (uint64_t)lane * 16;
        uint32_t w[4] = {0u, 0u, 0u, 0u}, m = 0;
        if (b0 < n) {
            fasta_load16(text, n, b0, aligned, w);
            // the multiply gathers the four flag bits of a word (bit 0 of every byte) into a nibble
#pragma unroll
            for (int k = 0; k < 4; ++k) m |= (((fasta_keep_flags4(w[k]) * 0x01020408u) >> 24) & 0xFu) << (4 * k);
        }
        const uint32_t k = __popc(m);
        uint32_t incl = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        uint32_t p = run + incl - k;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if ((m >> j) & 1u) stage[p++] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // stage[skew, run) -> dst0[0, run - skew): index i of the row goes to address (dst0 - skew) + i, 16-byte aligned at i % 16 == 0
    uint8_t* const dst = dst0 - skew;
    const uint32_t lo = skew, hi = run, va = (lo + 15u) & ~15u, vb = hi & ~15u;
    if (va >= vb) {
        for (uint32_t i = lo + lane; i < hi; i += 32) dst[i] = stage[i];
    } else {
        if (lo + lane < va) dst[lo + lane] = stage[lo + lane];                      // at most 15 bytes
        for (uint32_t v = va + 16u * lane; v < vb; v += 512u)
            *reinterpret_cast<uint4*>(dst + v) = *reinterpret_cast<const uint4*>(stage + v);
        if (vb + lane < hi) dst[vb + lane] = stage[vb + lane];                      // at most 15 bytes
    }
}

// kept bytes before text position pos[i] (positions inside blanked / ordinary text): one warp per position, every lane
// counts 128 bytes of the position's block
__global__ void __launch_bounds__(128) fasta_offsets_at(const uint8_t* __restrict__ text, uint64_t n,
                                                        const uint64_t* __restrict__ off, const uint64_t* __restrict__ pos,
                                                        uint32_t n_pos, uint64_t* __restrict__ out) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_pos) return;
    const uint64_t p = pos[i] < n ? pos[i] : n;
    const uint64_t blk = p / kFastaBlock;
    uint32_t k = 0;
    const uint64_t j0 = blk * kFastaBlock + (uint64_t)lane * (kFastaBlock / 32);
    for (uint64_t j = j0; j < j0 + kFastaBlock / 32 && j < p; ++j) k += fasta_keep(text[j]) ? 1u : 0u;
    for (int d = 16; d; d >>= 1) k += __shfl_xor_sync(0xffffffffu, k, d);
    if (lane == 0) out[i] = off[blk] + k;
}

}  // namespace mpcr
