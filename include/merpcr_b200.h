/*
 * merpcr_b200.h -- C ABI of the B200-native STS-search path (libmerpcr_b200.so).
 *
 * The reference (FOI-Bioinformatics/merpcr, pure Python) has no native seam; the preserved surface is its
 * Python API and CLI (merpcr_b200/engine.py, merpcr_b200/cli.py mirror them).  This header is the boundary
 * between that thin host code and the hand-written sm_100a kernels.  Each entry point names the reference
 * function(s) it replaces (paths relative to /root/reference/src/merpcr/).
 *
 * Conventions
 *   - every function returns 0 on success or a negative MPCR_E* code; mpcr_last_error() gives the text
 *     (thread-local);
 *   - "d_" pointers are device memory owned by the caller (torch tensors in the Python host), "h_" pointers
 *     are host memory; the library never frees caller memory;
 *   - every launch goes onto the caller's stream (cudaStream_t passed as void*); functions are asynchronous
 *     unless documented otherwise;
 *   - one context per device; a context is not thread-safe, distinct contexts may be used concurrently;
 *   - there is NO CPU fallback: every compute entry point fails with MPCR_ECUDA when no device is usable.
 *
 * Genome layout in HBM ("planes").  Contigs are concatenated in a padded global base
 * coordinate: contig c starts at gstart[c], a multiple of 128, and gstart[c+1] >= gstart[c] + length[c] + 1;
 * the gaps are zero (invalid).  For a base with global coordinate g and plane origin o (multiple of 128):
 *   plane2  2 bit/base  : word (g-o)/32 of uint64, bits [2*((g-o)%32), +2)   A0 C1 G2 T3(U3), other -> 0
 *   plane4  4 bit/base  : word (g-o)/16 of uint64, bits [4*((g-o)%16), +4)   IUPAC mask A1 C2 G4 T8 ... N15, X/other 0
 *   valid   1 bit/base  : word (g-o)/64 of uint64, bit (g-o)%64              1 iff the base is A/C/G/T/U
 * Algorithmic traffic is 0.75 B/bp (plane2 + plane4); `valid` is a derived private acceleration plane.
 */
#ifndef MERPCR_B200_H
#define MERPCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCR_ABI_VERSION 15

enum {
    MPCR_OK = 0,
    MPCR_EINVAL = -1,   /* bad argument (the Python host raises ValueError)              */
    MPCR_ECUDA = -2,    /* CUDA runtime error / no device                                */
    MPCR_ENOMEM = -3,   /* allocation failure                                            */
    MPCR_ESTATE = -4,   /* call order violated (e.g. scan before table build)            */
    MPCR_EOVERFLOW = -5 /* hit buffer too small; *count holds the required capacity      */
};

typedef struct mpcr_ctx mpcr_ctx;

/* Search parameters: the constructor arguments of MerPCR (core/engine.py:47-57), validated as :80-97. */
typedef struct mpcr_params {
    int32_t wordsize;           /* -W  3..16   */
    int32_t margin;             /* -M  0..10000 */
    int32_t mismatches;         /* -N  0..10   */
    int32_t three_prime_match;  /* -X  >= 0    */
    int32_t iupac_mode;         /* -I  0/1     */
} mpcr_params;

/* One contig of the padded global coordinate (see layout above). */
typedef struct mpcr_contig {
    uint64_t gstart;  /* multiple of 128 */
    uint32_t length;  /* true contig length in bases, < 2^31 */
    uint32_t reserved;
} mpcr_contig;

/* One hit = one output line of MerPCR.search (core/engine.py:437-444, models.py:52-58 STSHit).
 * Positions are 0-based inclusive, contig-local; rec = 2*sts_line_index + (0 for "+", 1 for "-").
 * rank / hash_off complete the reference's output-order key. */
typedef struct mpcr_hit {
    uint32_t contig;
    uint32_t pos1;
    uint32_t pos2;
    uint32_t rec;
    uint32_t rank;      /* 0 for delta=0, 2i-1 for delta=-i, 2i for delta=+i (engine.py:543-593) */
    uint32_t hash_off;  /* hash_offset of the record (engine.py:486)                          */
} mpcr_hit;

/* ---- lifecycle ---------------------------------------------------------------------------------- */
int mpcr_abi_version(void);
const char *mpcr_last_error(void);

/* Replaces MerPCR.__init__/_validate_parameters (core/engine.py:47-97). MPCR_EINVAL <-> ValueError. */
int mpcr_ctx_create(int device, const mpcr_params *params, mpcr_ctx **out);
void mpcr_ctx_destroy(mpcr_ctx *ctx);
/* Seed extension for EXACT searches (mismatches 0, no IUPAC mode; MPCR_EINVAL otherwise).  With no mismatch
 * allowed, a record can only match where the letters following its seed word match too, so a table may be keyed on
 * a longer word w_ext (wordsize < w_ext <= 16) taken at the SAME hash offset the reference computes -- which is what
 * keeps candidate-heavy searches (small -W, 10^5..10^6 STS) off the bucket walk.  Records whose primer has no w_ext
 * plain A/C/G/T letters from that offset cannot be keyed that way, hence two tables (two contexts):
 *   which = 1 : this context's next table holds only the records that can NOT be extended, keyed by wordsize;
 *   which = 2 : only the records that can, keyed by w_ext (mpcr_scan then hashes w_ext-mers);
 *   which = 0 : back to one table with every record (default).
 * Scanning both tables over the same planes and sorting the concatenated hits gives exactly the one-table result.
 * Call before mpcr_table_build; drops the current table. */
int mpcr_ctx_set_seed_extension(mpcr_ctx *ctx, int w_ext, int which);
/* Block tables for searches that ALLOW mismatches (mismatches = N >= 0, no IUPAC mode; MPCR_EINVAL otherwise; no
 * reference counterpart -- the reference walks its whole bucket at every position, core/engine.py:483-489).  The
 * reference reports a site only where the seed word matches exactly and at most N of the first primer's other letters
 * differ (core/engine.py:467-489, 515, 599-642), so of n_blocks > N disjoint blocks of `block` letters right behind the
 * seed at least one is identical to the genome.  A record whose wordsize + n_blocks * block (<= 16) letters from its
 * hash offset exist and are plain A/C/G/T goes to n_blocks tables, table i keyed on seed + block i (a split key of
 * wordsize + block letters); a site is reported by the table of its FIRST identical block only (the verifier of table i
 * drops a site whose block j < i is identical), so the concatenated hits hold every site once.  The other records stay
 * in an ordinary table.
 *   which = 1     : this context's next table holds only the records that can NOT be keyed this way, keyed by wordsize;
 *   which = 2 + i : only those that can, keyed on seed + block i (0 <= i < n_blocks);
 *   which = 0     : back to one table with every record (default).
 * Scanning all n_blocks + 1 tables over the same planes and sorting the concatenated hits gives exactly the one-table
 * result; candidate-heavy searches (small -W, 10^5..10^6 STS, -N >= 1) then run on sparse keys instead of bucket walks.
 * Combines with mpcr_ctx_set_table_part.  Call before mpcr_table_build; drops the current table and replaces any
 * mpcr_ctx_set_seed_extension setting (and vice versa). */
int mpcr_ctx_set_seed_blocks(mpcr_ctx *ctx, int block, int n_blocks, int which);
/* Position sampling for EXACT searches (mismatches 0, no IUPAC mode; MPCR_EINVAL otherwise; no reference counterpart --
 * the reference probes its dict at every base, core/engine.py:467-489).  With no mismatch allowed every window of a
 * record's first primer matches wherever the primer does, so a table may hold, per record, the w_samp-letter windows at
 * offsets hash_offset .. hash_offset + stride - 1 (all plain A/C/G/T, inside the primer) and the scanner may probe
 * only the positions that are multiples of `stride` in contig coordinates: exactly one of them sees each site.  The
 * first-level filter of such a table lives in global memory (the keys of 10^6 STS do not fit a shared-memory filter).
 *   role = 1 : this context's next table holds the records that can be sampled, `stride` windows each;
 *   role = 2 : it holds only the records that can NOT (combine with mpcr_ctx_set_seed_extension / _table_part);
 *   role = 0 : off (default).
 * Scanning the role-1 and role-2 tables over the same planes and sorting the concatenated hits gives exactly the
 * one-table result.  Ownership of a site (shards, ranges) goes by the probed position.  Call before
 * mpcr_table_build; drops the current table. */
int mpcr_ctx_set_sampling(mpcr_ctx *ctx, int w_samp, int stride, int role);
/* Items in the context's current table (records, or records x stride for a sampled table); 0 before a build. */
uint32_t mpcr_table_items(const mpcr_ctx *ctx);
/* Table partitioning (no reference counterpart; the reference keeps one dict, core/engine.py:324-329): this context's
 * next table holds only the STS lines with (line index % parts) == part
 * (parts = 0 or 1: every line, the default).  The shared-memory filter of the scanner has room for about 6.5 bits per
 * key at 2*10^5 keys; an exact search with 10^6 STS is therefore split into several extended tables (which = 2) of
 * at most a few 10^5 records each, scanned one after the other over the same planes -- every record lives in exactly
 * one table, so the concatenated, sorted hits are again the one-table result.  Call before mpcr_table_build. */
int mpcr_ctx_set_table_part(mpcr_ctx *ctx, uint32_t part, uint32_t parts);
/* Append mode (off by default; the counterpart of the reference collecting the hits of all chunks into one list,
 * core/engine.py:415-423): with on != 0, mpcr_scan no longer zeroes *d_count first -- the caller zeroes it once,
 * passes the SAME d_hits / capacity to a series of calls (several tables, or the ranges of a genome that is still being
 * uploaded) and reads the total once at the end, so the calls queue up on the stream without a host round trip in
 * between.  *d_count keeps counting past capacity; nothing is written beyond it. */
int mpcr_ctx_set_append(mpcr_ctx *ctx, int on);
/* NOT reference behaviour, off by default: with on != 0 the "+" record of an STS line looks for
 * primer1 ... revcomp(primer2) -- a biologically normal forward amplicon, as NCBI me-PCR does -- instead of the
 * reference's primer1 ... primer2 (core/engine.py:267).  "-" records are unchanged.  Call before mpcr_table_build. */
int mpcr_ctx_set_true_strands(mpcr_ctx *ctx, int on);
/* Multiprocessor count of the context's device (grid sizing is a multiple of this). */
int mpcr_ctx_sm_count(const mpcr_ctx *ctx);

/* ---- (1) FASTA ingest --------------------------------------------------------------------------- */
/* Replaces the sequence half of FASTALoader.load_file (io/fasta.py:58-61) + the per-base scode lookup of
 * MerPCR._process_thread (core/engine.py:455,472,497): packs n filtered ASCII bases (device memory) into the
 * three planes at global base dst_base (multiple of 64).  h_lut[256] maps a byte to
 * (nibble | code2 << 4 | clean << 6); it is copied to the device by the call.  Words fully inside
 * [dst_base, dst_base + n) are overwritten, the last partial word is written with zero padding. */
int mpcr_pack_sequence(mpcr_ctx *ctx, const uint8_t *d_ascii, uint64_t n, uint64_t dst_base,
                       uint64_t plane_origin, void *d_plane2, void *d_plane4, void *d_valid,
                       const uint8_t *h_lut, void *stream);

/* The same ingest for sequence that starts in HOST memory, shipping half the bytes over PCIe:
 * mpcr_host_pack_nibbles (host cores, multi-threaded; AVX-512 / AVX2 / scalar) turns n ASCII bases into (n + 1) / 2
 * bytes in plane4's own layout -- base 2k in the low nibble of byte k, base 2k+1 in the high one, nibble = h_lut[c] & 15
 * -- so the caller's H2D copy lands them directly in d_plane4 at byte (dst_base - plane_origin) / 2 (dst_base even; a
 * pinned staging buffer keeps the copy asynchronous), and mpcr_derive_planes then rebuilds the 2-bit and the valid
 * plane from plane4 on the device for the bases [dst_base, dst_base + n) (dst_base a multiple of 64; whole 64-base
 * strips are written, so what follows the sequence inside its last strip must be zero in plane4).
 * threads <= 0: all hardware threads.  Returns 0; 1 when the input holds a byte whose 2-bit code / clean flag do not
 * follow from its nibble (the one case: 'U' outside IUPAC mode, which hashes like T but equals nothing) -- the caller
 * then takes the ASCII path (mpcr_pack_sequence) for that piece; -1 for a null argument. */
int mpcr_host_pack_nibbles(const uint8_t *h_ascii, uint64_t n, const uint8_t *h_lut, uint8_t *h_dst, int threads);
int mpcr_derive_planes(mpcr_ctx *ctx, uint64_t n, uint64_t dst_base, uint64_t plane_origin, const void *d_plane4,
                       void *d_plane2, void *d_valid, void *stream);

/* Host helper for the file half of FASTALoader.load_file (io/fasta.py:36-41 `open` + line iteration): n bytes of the
 * file at `offset` into h_dst (pinned staging memory in the Python host), read by `threads` concurrent pread streams
 * (<= 0: all hardware threads) -- one thread copies out of the page cache at a few GB/s, the PCIe link behind it
 * takes 55.  Returns the bytes read (short only at end of file) or -errno. */
long long mpcr_file_read(const char *path, uint64_t offset, uint64_t n, uint8_t *h_dst, int threads);

/* Device-side FASTA text ingest: replaces the whole of FASTALoader.load_file (io/fasta.py:43-66) for ASCII files.
 * d_text holds the raw file bytes (device memory, n bytes).  mpcr_fasta_index finds the header lines ('>' as the
 * first non-blank character of a line; lines end at \n, \r or \r\n), blanks them and everything before the first
 * header IN PLACE, and returns per record the byte range of its header line in the file and where its filtered
 * sequence (characters of ACGTBDHKMNRSVWXY in either case, case preserved) will sit in the compacted stream.
 * *flags bit0 = the file has bytes >= 128 (nothing else is computed then; the host applies the locale rules).
 * MPCR_EOVERFLOW: more than max_records headers, *n_records holds the count.  Synchronous.
 * mpcr_fasta_compact then writes the kept bytes of all records, contiguous and in file order, to d_seq
 * (sum of seq_length bytes); d_ws is the workspace mpcr_fasta_index filled (mpcr_fasta_workspace_bytes). */
typedef struct mpcr_fasta_record {
    uint64_t header_begin, header_end;  /* file bytes [begin, end) = the header line from '>' to its terminator */
    uint64_t seq_offset, seq_length;    /* position and size of the record's sequence in the compacted stream   */
} mpcr_fasta_record;
uint64_t mpcr_fasta_workspace_bytes(uint64_t n, uint32_t max_records);
int mpcr_fasta_index(mpcr_ctx *ctx, uint8_t *d_text, uint64_t n, mpcr_fasta_record *h_records, uint32_t max_records,
                     uint32_t *n_records, uint32_t *flags, void *d_ws, uint64_t ws_bytes, void *stream);
int mpcr_fasta_compact(mpcr_ctx *ctx, const uint8_t *d_text, uint64_t n, const void *d_ws, uint8_t *d_seq,
                       void *stream);
/* The same index for a SLICE of a file (rank-local ingest of a multi-GPU run: every rank reads its own byte range plus
 * margins; the reference has one reader, io/fasta.py:36-41).  mode bit0: the slice may start inside a line -- the bytes
 * up to and including the first line terminator are dropped first (*flags bit1 and nothing else when there is none in
 * the first MiB); bit1: the sequence bytes in front of the slice's first header are KEPT (they continue a record that
 * began earlier in the file) and reported as record 0 with header_begin == header_end == the first retained byte.
 * mode 0 == mpcr_fasta_index. */
int mpcr_fasta_index_ex(mpcr_ctx *ctx, uint8_t *d_text, uint64_t n, uint32_t mode, mpcr_fasta_record *h_records,
                        uint32_t max_records, uint32_t *n_records, uint32_t *flags, void *d_ws, uint64_t ws_bytes,
                        void *stream);
/* Kept (sequence) bytes in front of the given byte positions of an indexed text (after mpcr_fasta_index*, same d_ws):
 * where a byte range of the file begins and ends in the compacted stream.  Synchronous. */
int mpcr_fasta_offsets_at(mpcr_ctx *ctx, const uint8_t *d_text, uint64_t n, const void *d_ws, const uint64_t *h_pos,
                          uint32_t n_pos, uint64_t *h_out, void *stream);

/* ---- host-side text helpers (no device work) ---------------------------------------------------- */
/* One accepted STS line: byte ranges of its fields inside the file text, its 1-based line number, and the expected
 * size when the field is a plain "123" or "100-200" (-1: the host must apply int() itself, engine.py:304-322). */
typedef struct mpcr_sts_line {
    uint32_t line_no;
    uint32_t id_off, id_len, p1_off, p1_len, p2_off, p2_len, size_off, size_len, alias_off, alias_len;
    int32_t pcr_size;
} mpcr_sts_line;
/* Replaces the line loop of MerPCR.load_sts_file (core/engine.py:216-243) for ASCII files: universal newlines, strip(),
 * blank / '#' lines skipped, split on tabs, lines whose primers are shorter than wordsize counted in *short_primers
 * and dropped.  *bad_line != 0: 1-based number of the first line with fewer than 4 fields (the load fails there).
 * *flags bit0: the text has bytes >= 128 (nothing parsed; the host applies the locale rules).
 * MPCR_EOVERFLOW: more than max_lines accepted lines, *n_lines holds the count. */
int mpcr_sts_parse(const uint8_t *text, uint64_t n, int32_t wordsize, int32_t default_pcr_size, mpcr_sts_line *lines,
                   uint32_t max_lines, uint32_t *n_lines, uint32_t *bad_line, uint32_t *short_primers, uint32_t *flags);
/* Upper-cased primers of the accepted lines back to back (primer1, primer2 per line) + 2*n_lines+1 offsets: the
 * h_blob / h_off arguments of mpcr_table_build. */
int mpcr_sts_blob(const uint8_t *text, const mpcr_sts_line *lines, uint32_t n_lines, uint8_t *blob, uint64_t *off);
/* Replaces the formatting loop of MerPCR.search (core/engine.py:437-444): one line
 * "label\tpos1+1..pos2+1\tid\talias\t(+|-)\n" per hit, in the given order.  hits are host memory; rec >> 1 indexes
 * `lines`; labels[label_off[c] .. label_off[c+1]) is the label of contig c.  Returns the bytes needed; the text is
 * written only when it fits out_cap. */
uint64_t mpcr_format_hits(const mpcr_hit *hits, uint64_t n, const uint8_t *text, const mpcr_sts_line *lines,
                          const uint8_t *labels, const uint64_t *label_off, uint8_t *out, uint64_t out_cap);

/* ---- (2) primer word-hash table ----------------------------------------------------------------- */
/* Replaces the hashing half of MerPCR.load_sts_file + _hash_value + _reverse_complement + _insert_sts
 * (core/engine.py:253-281, 324-359).  Input = the accepted STS lines after the host-side text rules
 * (engine.py:216-251): for line i, upper-cased primer1 = blob[off[2i] .. off[2i+1]) and primer2 =
 * blob[off[2i+1] .. off[2i+2]); pcr_size[i] already adjusted (engine.py:245-247) and clamped to 2^31-1.
 * h_primer_lut[256]: byte -> (nibble | never_match << 4 | zero_code_char << 5) for the primer side of
 * _compare_seqs.  All pointers are HOST memory; the call copies them, builds both strand records per line
 * (record 2i = "+", 2i+1 = "-"), hashes, and builds the bucket table on the device.  Synchronous. */
int mpcr_table_build(mpcr_ctx *ctx, const uint8_t *h_blob, const uint64_t *h_off, const uint32_t *h_pcr_size,
                     uint32_t n_lines, const uint8_t *h_primer_lut, void *stream);
/* Read back per-record results of the build (2*n_lines entries each): hash_offset (-1 = record not
 * inserted: no clean W-mer, engine.py:266-270,275-281) and the reference's big-endian hash value. */
int mpcr_table_records(mpcr_ctx *ctx, int32_t *h_hash_offset, uint32_t *h_hash);
/* Debug/test readback of the encoded primers of one record (nibble words then aux words). */
int mpcr_table_primer_words(mpcr_ctx *ctx, uint32_t rec, int which /*1|2*/, uint64_t *h_words, uint32_t max_words,
                            uint32_t *n_words);

/* ---- (3)+(4)+(5) scanner, verifier, hit emitter ------------------------------------------------- */
/* Optional (the counterpart of the chunk table `search` draws up before it starts its workers,
 * core/engine.py:386-411), before a series of mpcr_scan calls over sub-ranges of [shard_begin, shard_end): builds and uploads the
 * scanner's work descriptors for the WHOLE range once (host pointers, synchronous on `stream`).  Later mpcr_scan calls
 * with the same contig table and origin whose range is cut between contigs (or at these bounds) then reuse them
 * without any upload or host round trip -- a genome can be scanned contig by contig, in append mode, while its
 * later contigs are still being copied to the device (the copy engine is never needed by the scan calls). */
int mpcr_scan_prepare(mpcr_ctx *ctx, const mpcr_contig *contigs, uint32_t n_contigs, uint64_t plane_origin,
                      uint64_t shard_begin, uint64_t shard_end, void *stream);
/* Replaces MerPCR._process_thread + _match_sts + _compare_seqs (core/engine.py:453-642) over the hash
 * positions whose global base coordinate lies in [shard_begin, shard_end) (multiples of 128; pass 0 and
 * UINT64_MAX for everything).  Ownership is decided per unit of 2048 hash positions (units start at multiples of
 * 2048 from their contig's first base): a unit belongs to the range that holds its first base, so consecutive
 * ranges -- other ranks, or the ranges of one genome scanned one after the other -- neither overlap nor leave a gap,
 * wherever they are cut.  The planes must cover every base the shard can touch:
 * [shard_begin - mpcr_halo_left(), shard_end + mpcr_halo_right()) clipped to the genome.
 * plane_bases = the number of bases the three plane ALLOCATIONS hold, counted from plane_origin (zero-filled where no
 * contig lies).  The call checks it against what the kernels read -- whole units of 2048 positions plus 256 bases of
 * read-ahead behind the last scanned position, the mate window of the last position (never past its contig's end)
 * plus 64 bases of word over-read, nothing in front of plane_origin -- and fails with MPCR_EINVAL instead of reading
 * out of bounds.  Allocating (last base touched - plane_origin) + mpcr_tile_bases() + 1024 bases always suffices.
 * Hits are appended (unordered) to d_hits; *d_count receives the TRUE number of hits even when it
 * exceeds capacity (the caller re-runs with a larger buffer; nothing is silently truncated). */
int mpcr_scan(mpcr_ctx *ctx, const mpcr_contig *h_contigs, uint32_t n_contigs, const void *d_plane2,
              const void *d_plane4, const void *d_valid, uint64_t plane_origin, uint64_t plane_bases,
              uint64_t shard_begin, uint64_t shard_end, mpcr_hit *d_hits, uint64_t capacity,
              uint64_t *d_count, void *stream);
uint64_t mpcr_halo_left(const mpcr_ctx *ctx);
/* Hash positions per scanner tile; the planes must be allocated with at least this many bases (+1024) of zeroed
 * slack behind the last base a shard can touch (tiles are staged whole). */
uint64_t mpcr_tile_bases(void);
uint64_t mpcr_halo_right(const mpcr_ctx *ctx);

/* Replaces the sort of MerPCR.search (core/engine.py:434) including its tie order: orders n hits by
 * (contig, pos1, hash_off, rec, rank) == "stable sort of discovery order by pos1" (SURVEY.md A.7). */
int mpcr_sort_hits(mpcr_ctx *ctx, mpcr_hit *d_hits, uint64_t n, void *stream);
/* The same sort queued right behind mpcr_scan with no host round trip in between (the reference sorts as soon as its
 * workers return, core/engine.py:424-434): the record count is read ON THE DEVICE as min(*d_count, capacity), with
 * d_count / capacity the arguments mpcr_scan was given.  n_hint = the count the caller expects (e.g. the previous
 * scan's), 0 = unknown; it only selects which kernels are launched (lists of up to ~12k hits are ordered by one
 * rank-sort launch instead of the radix passes), never the result. */
int mpcr_sort_hits_dev(mpcr_ctx *ctx, mpcr_hit *d_hits, const uint64_t *d_count, uint64_t capacity, uint64_t n_hint,
                       void *stream);

/* One step of the hot path in ONE call (what MerPCR.search does per sequence between reading it and printing,
 * core/engine.py:373-434): mpcr_scan with every table of `ctxs` (the first zeroes *d_count, the others append behind
 * it -- one context unless the search was split into several tables), then, if sort != 0, mpcr_sort_hits_dev with
 * ctxs[0], all queued on `stream` without a host round trip.  h_count != NULL (pinned host memory): the true hit count
 * is copied there and the call returns after the stream has drained, i.e. *h_count is valid on return (the caller
 * re-runs with a larger buffer when it exceeds capacity); h_count == NULL: fully asynchronous. */
int mpcr_scan_sorted(mpcr_ctx *const *ctxs, uint32_t n_ctx, const mpcr_contig *h_contigs, uint32_t n_contigs,
                     const void *d_plane2, const void *d_plane4, const void *d_valid, uint64_t plane_origin,
                     uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end, mpcr_hit *d_hits, uint64_t capacity,
                     uint64_t *d_count, uint64_t *h_count, uint64_t n_hint, int sort, void *stream);

/* The same step without the final synchronisation, for callers that keep several steps in flight (the host reads step
 * k's count while steps k+1 .. run; consecutive steps on one stream share the context's scratch, only d_hits / d_count /
 * h_result must differ between the steps in flight): h_result = two uint64 of PINNED host memory, [0] receives the true
 * hit count, [1] != 0 means the short-list sort gave up and mpcr_sort_finish must run -- both valid once the caller has
 * synchronised with `stream` (event, stream sync).  slot (0 .. MPCR_MAX_SLOTS - 1) selects the set of timing events
 * the step records (mpcr_slot_scan_ms / mpcr_slot_verify_ms read them back without waiting for a later step). */
#define MPCR_MAX_SLOTS 8
int mpcr_scan_sorted_async(mpcr_ctx *const *ctxs, uint32_t n_ctx, const mpcr_contig *h_contigs, uint32_t n_contigs,
                           const void *d_plane2, const void *d_plane4, const void *d_valid, uint64_t plane_origin,
                           uint64_t plane_bases, uint64_t shard_begin, uint64_t shard_end, mpcr_hit *d_hits,
                           uint64_t capacity, uint64_t *d_count, uint64_t *h_result, uint64_t n_hint, int sort, int slot,
                           void *stream);
/* After the caller has synchronised behind mpcr_scan_sorted_async: completes the sort when h_result[1] says so
 * (radix passes with the count now known; synchronous), a no-op otherwise. */
int mpcr_sort_finish(mpcr_ctx *ctx, mpcr_hit *d_hits, const uint64_t *h_result, uint64_t capacity, void *stream);
float mpcr_slot_scan_ms(mpcr_ctx *ctx, int slot);
float mpcr_slot_verify_ms(mpcr_ctx *ctx, int slot);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t mpcr_launch_count(const mpcr_ctx *ctx);
/* Name / elapsed-ms of the most recent scan kernel as timed by CUDA events on its stream (0 if none). */
float mpcr_last_scan_ms(mpcr_ctx *ctx);
/* Same for the verify kernel that follows it (primer compare + mate search of the surviving seed positions). */
float mpcr_last_verify_ms(mpcr_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MERPCR_B200_H */
