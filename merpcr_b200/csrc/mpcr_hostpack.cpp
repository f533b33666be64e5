// mpcr_hostpack.cpp -- host half of the FASTA ingest when the sequence starts in HOST memory: pack ASCII bases into
// the 4-bit plane's own byte layout (two bases per byte, low nibble first) on the host cores, so that only 0.5 byte
// per base crosses PCIe and the copy lands directly in plane4; the device then derives the 2-bit and the valid plane
// from it (derive_planes_kernel).  Replaces, for host-resident sequences, the per-base scode / upper() work of
// MerPCR._process_thread (core/engine.py:455,472,497 of the reference) exactly like mpcr_pack_sequence does.
//
// Compiled by g++ (AVX2 path behind a run-time CPU check, scalar path otherwise) and linked into libmerpcr_b200.so.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define MPCR_X86 1
#endif

namespace mpcr_hostpack {

// lut[c] = nibble | code2 << 4 | clean << 6 (merpcr_b200/alphabet.py genome_lut).  A byte is REGULAR when its 2-bit code
// and clean flag follow from its nibble alone (A1 C2 G4 T8 -> clean, code = log2; everything else not clean, code 0):
// only then can the device rebuild plane2 / valid from plane4.  (The one irregular case: 'U' in non-IUPAC mode, which
// hashes like T but equals nothing -- such sequences take the ASCII path.)
static inline bool regular(uint8_t e) {
    const uint32_t nib = e & 15u, code = (e >> 4) & 3u, clean = (e >> 6) & 1u;
    const bool one_hot = nib == 1 || nib == 2 || nib == 4 || nib == 8;
    const uint32_t want_code = nib == 2 ? 1u : nib == 4 ? 2u : nib == 8 ? 3u : 0u;
    return one_hot ? (clean == 1 && code == want_code) : (clean == 0 && code == 0);
}

struct Tables {
    uint8_t nib[256];        // nibble of every byte
    uint8_t irregular[256];  // 1: the byte cannot go through the nibble path
    bool letters_only;       // every byte outside 0x40..0x7F maps to nibble 0 and is regular, upper == lower case
    uint8_t nib32[32], irr32[32];   // by (c & 31), for bytes 0x40..0x7F
};

static void make_tables(const uint8_t* lut, Tables& t) {
    t.letters_only = true;
    for (int c = 0; c < 256; ++c) {
        t.nib[c] = lut[c] & 15u;
        t.irregular[c] = regular(lut[c]) ? 0 : 1;
        if ((c & 0xC0) != 0x40 && (t.nib[c] || t.irregular[c])) t.letters_only = false;
    }
    for (int i = 0; i < 32; ++i) {
        const int up = 0x40 + i, lo = 0x60 + i;
        if (t.nib[up] != t.nib[lo] || t.irregular[up] != t.irregular[lo]) t.letters_only = false;
        t.nib32[i] = t.nib[up];
        t.irr32[i] = t.irregular[up];
    }
}

// scalar: n bases (n even or the last byte gets a zero high nibble) -> (n + 1) / 2 bytes; returns 1 if irregular seen
static int pack_scalar(const Tables& t, const uint8_t* src, uint64_t n, uint8_t* dst) {
    uint32_t irr = 0;
    uint64_t i = 0;
    for (; i + 1 < n; i += 2) {
        const uint8_t a = src[i], b = src[i + 1];
        dst[i >> 1] = (uint8_t)(t.nib[a] | (t.nib[b] << 4));
        irr |= t.irregular[a] | t.irregular[b];
    }
    if (i < n) {
        dst[i >> 1] = t.nib[src[i]];
        irr |= t.irregular[src[i]];
    }
    return (int)irr;
}

#ifdef MPCR_X86
struct Avx2Consts {
    __m256i lo16, hi16, ilo16, ihi16, m0f, mc0, m40, m10;
};
// 32 bytes -> 32 nibbles (one per byte), case-insensitive 32-entry table by (c & 31) for bytes 0x40..0x7F, else 0
__attribute__((target("avx2"))) static inline __m256i nibbles32(const Avx2Consts& k, __m256i c, __m256i& irr) {
    const __m256i idx = _mm256_and_si256(c, k.m0f);                                    // low 4 bits of (c & 31)
    const __m256i sel = _mm256_cmpeq_epi8(_mm256_and_si256(c, k.m10), k.m10);          // bit 4: upper half of the table
    const __m256i letter = _mm256_cmpeq_epi8(_mm256_and_si256(c, k.mc0), k.m40);       // 0x40..0x7F
    const __m256i v = _mm256_blendv_epi8(_mm256_shuffle_epi8(k.lo16, idx), _mm256_shuffle_epi8(k.hi16, idx), sel);
    const __m256i r = _mm256_blendv_epi8(_mm256_shuffle_epi8(k.ilo16, idx), _mm256_shuffle_epi8(k.ihi16, idx), sel);
    irr = _mm256_or_si256(irr, _mm256_and_si256(r, letter));
    return _mm256_and_si256(v, letter);
}
__attribute__((target("avx2"))) static int pack_avx2(const Tables& t, const uint8_t* src, uint64_t n, uint8_t* dst) {
    Avx2Consts k;
    k.lo16 = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t.nib32)));
    k.hi16 = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t.nib32 + 16)));
    k.ilo16 = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t.irr32)));
    k.ihi16 = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t.irr32 + 16)));
    k.m0f = _mm256_set1_epi8(0x0F); k.mc0 = _mm256_set1_epi8((char)0xC0); k.m40 = _mm256_set1_epi8(0x40);
    k.m10 = _mm256_set1_epi8(0x10);
    const __m256i mul = _mm256_set1_epi16(0x1001);
    __m256i irr = _mm256_setzero_si256();
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = nibbles32(k, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), irr);
        const __m256i b = nibbles32(k, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32)), irr);
        // pairs (even | odd << 4) as 16-bit lanes, then back to bytes; packus works per 128-bit lane -> fix the order
        const __m256i pa = _mm256_maddubs_epi16(a, mul), pb = _mm256_maddubs_epi16(b, mul);
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi16(pa, pb), 0xD8);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 1)), pk);
    }
    int bad = !_mm256_testz_si256(irr, irr);
    if (i < n) bad |= pack_scalar(t, src + i, n - i, dst + (i >> 1));
    return bad;
}
#endif

#ifdef MPCR_X86
// AVX-512 (BW + VBMI): one 64-entry byte permute is the whole look-up for the 64 bytes 0x40..0x7F (index = c & 63,
// both letter cases at once); 128 bases in, 64 bytes out per iteration, streamed past the cache when the destination
// allows (the packed bytes are only ever read by the DMA engine).
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) static int pack_avx512(const Tables& t, const uint8_t* src, uint64_t n,
                                                                                uint8_t* dst) {
    const __m512i tab = _mm512_loadu_si512(t.nib + 0x40), itab = _mm512_loadu_si512(t.irregular + 0x40);
    const __m512i mc0 = _mm512_set1_epi8((char)0xC0), m40 = _mm512_set1_epi8(0x40), mul = _mm512_set1_epi16(0x1001);
    const __m512i fix = _mm512_setr_epi64(0, 2, 4, 6, 1, 3, 5, 7);
    const bool stream = ((uintptr_t)dst & 63u) == 0;
    __mmask64 irr = 0;
    uint64_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m512i c0 = _mm512_loadu_si512(src + i), c1 = _mm512_loadu_si512(src + i + 64);
        const __mmask64 l0 = _mm512_cmpeq_epi8_mask(_mm512_and_si512(c0, mc0), m40);
        const __mmask64 l1 = _mm512_cmpeq_epi8_mask(_mm512_and_si512(c1, mc0), m40);
        const __m512i a = _mm512_maskz_permutexvar_epi8(l0, c0, tab), b = _mm512_maskz_permutexvar_epi8(l1, c1, tab);
        irr |= _mm512_mask_test_epi8_mask(l0, _mm512_permutexvar_epi8(c0, itab), _mm512_set1_epi8(1));
        irr |= _mm512_mask_test_epi8_mask(l1, _mm512_permutexvar_epi8(c1, itab), _mm512_set1_epi8(1));
        const __m512i pk = _mm512_permutexvar_epi64(fix, _mm512_packus_epi16(_mm512_maddubs_epi16(a, mul), _mm512_maddubs_epi16(b, mul)));
        if (stream) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + (i >> 1)), pk);
        else _mm512_storeu_si512(dst + (i >> 1), pk);
    }
    if (stream) _mm_sfence();
    int bad = irr != 0;
    if (i < n) bad |= pack_scalar(t, src + i, n - i, dst + (i >> 1));
    return bad;
}
#endif

// ---- a small persistent worker pool (a thread spawn per 64 MiB piece would cost as much as the piece's copy) ----
class Pool {
public:
    explicit Pool(int n) : stop_(false), gen_(0), pending_(0) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int size() const { return (int)workers_.size(); }
    // run fn(part) for part in [0, parts) on the pool (parts <= size()), wait for all
    template <class F>
    void run(int parts, F&& fn) {
        std::function<void(int)> f = fn;
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &f;
            parts_ = parts;
            pending_ = parts;
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = id < parts_ ? job_ : nullptr;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_;
    uint64_t gen_;
    int pending_, parts_ = 0;
    std::function<void(int)>* job_ = nullptr;
};

}  // namespace mpcr_hostpack

using namespace mpcr_hostpack;

static std::mutex g_pool_mutex;
static Pool* g_pool = nullptr;

extern "C" {

// See include/merpcr_b200.h.  Returns 0, or 1 when the input holds a byte the nibble path cannot carry (dst is then
// unspecified for that call and the caller takes the ASCII path), or -1 for bad arguments.
int mpcr_host_pack_nibbles(const uint8_t* h_ascii, uint64_t n, const uint8_t* h_lut, uint8_t* h_dst, int threads) {
    if (n == 0) return 0;
    if (!h_ascii || !h_lut || !h_dst) return -1;
    Tables t;
    make_tables(h_lut, t);
    int (*kernel)(const Tables&, const uint8_t*, uint64_t, uint8_t*) = pack_scalar;
#ifdef MPCR_X86
    if (t.letters_only && __builtin_cpu_supports("avx2")) kernel = pack_avx2;
    if (t.letters_only && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi")) kernel = pack_avx512;
#endif
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (threads <= 0 || threads > hw) threads = hw;
    // one part per thread, cut on multiples of 128 bases: every part starts on a whole, 64-byte aligned output line
    const uint64_t per = ((n + threads - 1) / threads + 127) / 128 * 128;
    int parts = (int)((n + per - 1) / per);
    if (parts <= 1 || n < (1u << 16)) return kernel(t, h_ascii, n, h_dst);
    std::lock_guard<std::mutex> g(g_pool_mutex);   // one packing call at a time per process
    if (!g_pool || g_pool->size() < parts) {
        delete g_pool;
        g_pool = new Pool(parts > hw ? parts : hw);
    }
    std::atomic<int> bad(0);
    g_pool->run(parts, [&](int p) {
        const uint64_t a = (uint64_t)p * per, b = a + per < n ? a + per : n;
        if (kernel(t, h_ascii + a, b - a, h_dst + (a >> 1))) bad.store(1);
    });
    return bad.load();
}

// See include/merpcr_b200.h: n bytes of the file at `offset` into h_dst, the range cut into one pread stream per thread
// (a single thread copies out of the page cache at a few GB/s; the PCIe link behind it takes 55).  Returns the number
// of bytes read (short only at end of file) or -errno.
long long mpcr_file_read(const char* path, uint64_t offset, uint64_t n, uint8_t* h_dst, int threads) {
    if (!path || (n && !h_dst)) return -EINVAL;
    if (n == 0) return 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return -(long long)errno;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (threads <= 0 || threads > hw) threads = hw;
    const uint64_t per = ((n + threads - 1) / threads + 4095) / 4096 * 4096;
    const int parts = (int)((n + per - 1) / per);
    std::atomic<long long> total(0), err(0);
    auto part = [&](int p) {
        uint64_t a = (uint64_t)p * per;
        const uint64_t b = a + per < n ? a + per : n;
        while (a < b) {
            const ssize_t k = pread(fd, h_dst + a, b - a, (off_t)(offset + a));
            if (k < 0) { if (errno == EINTR) continue; err.store(-(long long)errno); return; }
            if (k == 0) break;   // end of file
            a += (uint64_t)k;
            total.fetch_add(k);
        }
    };
    if (parts <= 1) {
        part(0);
    } else {
        std::lock_guard<std::mutex> g(g_pool_mutex);
        if (!g_pool || g_pool->size() < parts) {
            delete g_pool;
            g_pool = new Pool(parts > hw ? parts : hw);
        }
        g_pool->run(parts, part);
    }
    close(fd);
    return err.load() ? err.load() : total.load();
}

}  // extern "C"
