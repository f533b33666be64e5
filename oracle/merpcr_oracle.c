/*
 * merpcr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the STS-search hot path of FOI-Bioinformatics/merpcr
 * (pure-Python reference mounted at /root/reference).  It exists so that the CUDA path
 * in merpcr_b200/ can be checked bit-for-bit at sizes CPython cannot finish, and so that
 * bench.py has a CPU baseline that travels to the GPU box (the Python reference cannot).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (merpcr_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   - the reference's own fixture golden line (tests/test_comprehensive.py:65-95),
 *   - its unit known-answers (tests/test_engine_internals.py:26-62, test_utils_comprehensive.py:173-181),
 *   - several hundred seeded fuzz cases whose expected output text was produced by importing the
 *     reference in the build container (tests/golden/make_golden.py -> tests/golden/ JSON files).
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src/merpcr/).  Nothing here is copied; the reference is Python.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_AMBIG 100                    /* core/engine.py:18 */
#define ORC_MIN_THREADING 100000         /* core/engine.py:19 MIN_FILESIZE_FOR_THREADING */

/* ------------------------------------------------------------------ small utils */

typedef struct { char *p; size_t n, cap; } sbuf;

static void sbuf_reserve(sbuf *b, size_t extra) {
    if (b->n + extra + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 256;
        while (nc < b->n + extra + 1) nc *= 2;
        b->p = (char *)realloc(b->p, nc);
        b->cap = nc;
    }
}
static void sbuf_put(sbuf *b, const char *s, size_t n) {
    sbuf_reserve(b, n);
    memcpy(b->p + b->n, s, n);
    b->n += n;
    b->p[b->n] = 0;
}
static char *xstrndup(const char *s, size_t n) {
    char *r = (char *)malloc(n + 1);
    memcpy(r, s, n);
    r[n] = 0;
    return r;
}

/* Python str.strip() whitespace for ASCII input: space, \t \n \v \f \r and the
 * separators \x1c-\x1f. */
static int py_isspace(unsigned char c) {
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}
static void py_strip(const char **s, size_t *n) {
    const char *p = *s; size_t m = *n;
    while (m && py_isspace((unsigned char)p[0])) { p++; m--; }
    while (m && py_isspace((unsigned char)p[m - 1])) m--;
    *s = p; *n = m;
}

/* Python int(str): optional surrounding whitespace, optional sign, decimal digits with single
 * underscores between digits.  Returns 0 on ValueError.  Saturates at +-2^62 (the reference has
 * unbounded ints; every consumer clamps long before that). */
static int py_int(const char *s, size_t n, long long *out) {
    py_strip(&s, &n);
    if (!n) return 0;
    int neg = 0;
    if (s[0] == '+' || s[0] == '-') { neg = (s[0] == '-'); s++; n--; }
    if (!n || !isdigit((unsigned char)s[0])) return 0;
    long long v = 0;
    int prev_us = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned char c = (unsigned char)s[i];
        if (c == '_') { if (prev_us) return 0; prev_us = 1; continue; }
        if (!isdigit(c)) return 0;
        prev_us = 0;
        if (v < (1LL << 62) / 10) v = v * 10 + (c - '0'); else v = (1LL << 62);
    }
    if (prev_us) return 0;
    *out = neg ? -v : v;
    return 1;
}

/* ------------------------------------------------------------------ engine state */

typedef struct {
    char *id, *alias;          /* models.py:17-29 STSRecord */
    char *primer1, *primer2;
    int len1, len2;
    long long pcr_size;
    int offset;                /* 1-based source line number */
    int hash_offset;
    uint32_t hash;             /* big-endian 2-bit pack, engine.py:350 */
    char direct;               /* '+' or '-' */
    int next;                  /* next record in the same bucket (insertion order), -1 = end */
} orc_sts;

typedef struct orc_engine {
    int W, M, N, X, I;
    long long Z;
    int scode[256];            /* engine.py:102-109 */
    unsigned char compl_[256]; /* engine.py:112-135 ; 0 = not in the map */
    unsigned short iupac[256]; /* engine.py:138-172 as a bit set over "ACGTURYMKSWBDHVN"; 0 = not a key */
    orc_sts *recs; int nrec, caprec;
    long long max_pcr_size;
    /* sts_table (engine.py:70,324-329): chained, insertion-ordered buckets.
     * direct-addressed when 4^W <= 2^26, else open addressing on the hash value. */
    int direct_table;
    int *head, *tail;          /* direct: size 4^W */
    uint32_t *okey; int *ohead, *otail; uint32_t omask; /* open addressing */
    int bad_short, bad_ambig, bad_size;
    char err[256];
} orc_engine;

static const char *ORC_IUPAC_KEYS = "ACGTURYMKSWBDHVN";

static unsigned short iupac_set(const char *members) {
    unsigned short s = 0;
    for (const char *m = members; *m; m++) {
        const char *q = strchr(ORC_IUPAC_KEYS, *m);
        if (q) s |= (unsigned short)(1u << (q - ORC_IUPAC_KEYS));
    }
    return s;
}

/* engine.py:99-191 _init_lookup_tables */
static void orc_init_tables(orc_engine *e) {
    for (int i = 0; i < 256; i++) { e->scode[i] = ORC_AMBIG; e->compl_[i] = 0; e->iupac[i] = 0; }
    e->scode['A'] = e->scode['a'] = 0;
    e->scode['C'] = e->scode['c'] = 1;
    e->scode['G'] = e->scode['g'] = 2;
    e->scode['T'] = e->scode['t'] = 3;
    e->scode['U'] = e->scode['u'] = 3;
    static const char *pairs[] = {"AT","CG","GC","TA","UA","BV","DH","HD","KM","MK","NN","RY","SS","VB","WW","XX","YR"};
    for (size_t i = 0; i < sizeof(pairs) / sizeof(pairs[0]); i++) {
        e->compl_[(unsigned char)pairs[i][0]] = (unsigned char)pairs[i][1];
        e->compl_[(unsigned char)tolower(pairs[i][0])] = (unsigned char)tolower(pairs[i][1]);
    }
    static const char *map[][2] = {
        {"A","A"},{"C","C"},{"G","G"},{"T","TU"},{"U","TU"},{"R","AGR"},{"Y","CTUY"},{"M","ACM"},
        {"K","GTUK"},{"S","CGS"},{"W","ATUW"},{"B","CGTUYKSB"},{"D","AGTURKWD"},{"H","ACTUYMWH"},
        {"V","ACGRMSV"},{"N","ACGTURYMKSWBDHVN"}};
    for (size_t i = 0; i < sizeof(map) / sizeof(map[0]); i++) {
        unsigned short s = iupac_set(map[i][1]);
        e->iupac[(unsigned char)map[i][0][0]] = s;
        e->iupac[(unsigned char)tolower(map[i][0][0])] = s;
    }
}

/* engine.py:80-97 _validate_parameters ; returns NULL with *err filled on ValueError */
orc_engine *orc_new(int W, int M, int N, int X, int I, long long Z, char *err, size_t errlen) {
    const char *msg = NULL;
    if (W < 3 || W > 16) msg = "Word size must be between 3 and 16";
    else if (N < 0 || N > 10) msg = "Number of mismatches must be between 0 and 10";
    else if (M < 0 || M > 10000) msg = "Margin must be between 0 and 10000";
    else if (X < 0) msg = "Three prime match must be at least 0";
    else if (Z < 1 || Z > 10000) msg = "Default PCR size must be between 1 and 10000";
    if (msg) { if (err && errlen) snprintf(err, errlen, "%s", msg); return NULL; }
    orc_engine *e = (orc_engine *)calloc(1, sizeof(*e));
    e->W = W; e->M = M; e->N = N; e->X = X; e->I = I; e->Z = Z;
    orc_init_tables(e);
    return e;
}

static void orc_clear_sts(orc_engine *e) {
    for (int i = 0; i < e->nrec; i++) {
        free(e->recs[i].id); free(e->recs[i].alias); free(e->recs[i].primer1); free(e->recs[i].primer2);
    }
    free(e->recs); e->recs = NULL; e->nrec = e->caprec = 0;
    free(e->head); free(e->tail); e->head = e->tail = NULL;
    free(e->okey); free(e->ohead); free(e->otail); e->okey = NULL; e->ohead = e->otail = NULL;
    e->max_pcr_size = 0;
}

void orc_free(orc_engine *e) {
    if (!e) return;
    orc_clear_sts(e);
    free(e);
}

/* engine.py:331-355 _hash_value : first offset whose W-mer is all A/C/G/T/U; big-endian pack */
int orc_hash_value(const orc_engine *e, const char *primer, int len, uint32_t *hash) {
    if (len < e->W) { *hash = 0; return -1; }
    for (int off = 0; off + e->W <= len; off++) {
        uint32_t h = 0; int ok = 1;
        for (int i = 0; i < e->W; i++) {
            int code = e->scode[(unsigned char)toupper((unsigned char)primer[off + i])];
            if (code == ORC_AMBIG) { ok = 0; break; }
            h = (h << 2) | (uint32_t)code;
        }
        if (ok) { *hash = h; return off; }
    }
    *hash = 0;
    return -1;
}

/* engine.py:357-359 _reverse_complement : unknown characters become 'N' */
void orc_reverse_complement(const orc_engine *e, const char *s, int len, char *out) {
    for (int i = 0; i < len; i++) {
        unsigned char c = e->compl_[(unsigned char)s[len - 1 - i]];
        out[i] = c ? (char)c : 'N';
    }
    out[len] = 0;
}

/* engine.py:304-322 _parse_pcr_size */
static long long orc_parse_pcr_size(const orc_engine *e, const char *s, size_t n) {
    if (memchr(s, '-', n)) {
        /* split("-") must give exactly two non-empty parts */
        const char *d = (const char *)memchr(s, '-', n);
        size_t n0 = (size_t)(d - s), n1 = n - n0 - 1;
        if (memchr(d + 1, '-', n1) || n0 == 0 || n1 == 0) return e->Z;
        long long lo, hi;
        if (!py_int(s, n0, &lo) || !py_int(d + 1, n1, &hi)) return e->Z;
        long long sum = lo + hi;                 /* Python floor division */
        return (sum >= 0) ? sum / 2 : -((-sum + 1) / 2);
    }
    long long v;
    if (!py_int(s, n, &v)) return e->Z;
    return v > 0 ? v : e->Z;
}

static uint32_t omix(uint32_t h) { h ^= h >> 15; h *= 0x9E3779B1u; h ^= h >> 13; return h; }

/* engine.py:324-329 _insert_sts */
static void orc_insert(orc_engine *e, const orc_sts *proto, const char *p1, int l1, const char *p2, int l2,
                       int hash_offset, uint32_t hash, char direct) {
    if (e->nrec == e->caprec) {
        e->caprec = e->caprec ? e->caprec * 2 : 64;
        e->recs = (orc_sts *)realloc(e->recs, (size_t)e->caprec * sizeof(orc_sts));
    }
    orc_sts *r = &e->recs[e->nrec];
    r->id = strdup(proto->id); r->alias = strdup(proto->alias);
    r->primer1 = xstrndup(p1, (size_t)l1); r->primer2 = xstrndup(p2, (size_t)l2);
    r->len1 = l1; r->len2 = l2; r->pcr_size = proto->pcr_size; r->offset = proto->offset;
    r->hash_offset = hash_offset; r->hash = hash; r->direct = direct; r->next = -1;
    e->nrec++;
}

static void orc_build_table(orc_engine *e) {
    int W = e->W;
    e->direct_table = (2 * W <= 26);
    if (e->direct_table) {
        size_t sz = (size_t)1 << (2 * W);
        e->head = (int *)malloc(sz * sizeof(int)); e->tail = (int *)malloc(sz * sizeof(int));
        memset(e->head, 0xff, sz * sizeof(int));
        for (int i = 0; i < e->nrec; i++) {
            uint32_t h = e->recs[i].hash;
            if (e->head[h] < 0) e->head[h] = i; else e->recs[e->tail[h]].next = i;
            e->tail[h] = i;
        }
    } else {
        uint32_t sz = 1024; while (sz < (uint32_t)e->nrec * 2u + 2u) sz <<= 1;
        e->omask = sz - 1;
        e->okey = (uint32_t *)malloc(sz * sizeof(uint32_t));
        e->ohead = (int *)malloc(sz * sizeof(int)); e->otail = (int *)malloc(sz * sizeof(int));
        memset(e->ohead, 0xff, sz * sizeof(int));
        for (int i = 0; i < e->nrec; i++) {
            uint32_t h = e->recs[i].hash, s = omix(h) & e->omask;
            while (e->ohead[s] >= 0 && e->okey[s] != h) s = (s + 1) & e->omask;
            if (e->ohead[s] < 0) { e->okey[s] = h; e->ohead[s] = i; } else e->recs[e->otail[s]].next = i;
            e->otail[s] = i;
        }
    }
}

static inline int orc_bucket_head(const orc_engine *e, uint32_t h) {
    if (e->direct_table) return e->head[h];
    uint32_t s = omix(h) & e->omask;
    while (e->ohead[s] >= 0) { if (e->okey[s] == h) return e->ohead[s]; s = (s + 1) & e->omask; }
    return -1;
}

/* engine.py:193-302 load_sts_file, operating on the file's text.
 * returns 1 (True), 0 (False: empty input or a line with < 4 fields). */
int orc_load_sts_text(orc_engine *e, const char *text, size_t n) {
    orc_clear_sts(e);
    e->bad_short = e->bad_ambig = e->bad_size = 0;
    if (n == 0) return 0;                                           /* :198-200 */
    int line_no = 0, ok = 1;
    size_t pos = 0;
    while (pos < n) {
        /* universal newlines: \n, \r\n, \r all end a line (text-mode open, :212-213) */
        size_t eol = pos;
        while (eol < n && text[eol] != '\n' && text[eol] != '\r') eol++;
        const char *ln = text + pos; size_t ll = eol - pos;
        pos = eol;
        if (pos < n) { if (text[pos] == '\r' && pos + 1 < n && text[pos + 1] == '\n') pos += 2; else pos += 1; }
        line_no++;
        py_strip(&ln, &ll);                                         /* :218 */
        if (ll == 0 || ln[0] == '#') continue;                      /* :221 */
        const char *f[6]; size_t fl[6]; int nf = 0;                 /* :225 split("\t"), only 5 used */
        {
            size_t s = 0;
            for (size_t i = 0; i <= ll; i++) {
                if (i == ll || ln[i] == '\t') {
                    if (nf < 6) { f[nf] = ln + s; fl[nf] = i - s; }
                    nf++; s = i + 1;
                }
            }
        }
        if (nf < 4) {                                               /* :226-230 */
            snprintf(e->err, sizeof e->err, "Bad STS file format at line %d. Expected at least 4 fields.", line_no);
            ok = 0; break;
        }
        orc_sts proto; memset(&proto, 0, sizeof proto);
        proto.id = xstrndup(f[0], fl[0]);
        char *p1 = xstrndup(f[1], fl[1]), *p2 = xstrndup(f[2], fl[2]);
        int l1 = (int)fl[1], l2 = (int)fl[2];
        for (int i = 0; i < l1; i++) p1[i] = (char)toupper((unsigned char)p1[i]);   /* :233-234 */
        for (int i = 0; i < l2; i++) p2[i] = (char)toupper((unsigned char)p2[i]);
        long long pcr = orc_parse_pcr_size(e, f[3], fl[3]);         /* :237 */
        proto.alias = nf > 4 ? xstrndup(f[4], fl[4]) : strdup("");  /* :238 */
        proto.offset = line_no;
        if (l1 < e->W || l2 < e->W) {                               /* :241-243 */
            e->bad_short++;
        } else {
            if ((long long)l1 + l2 > pcr) { e->bad_size++; pcr = (long long)l1 + l2; }   /* :245-247 */
            if (pcr > e->max_pcr_size) e->max_pcr_size = pcr;       /* :250-251 */
            proto.pcr_size = pcr;
            uint32_t h1, h2;
            int o1 = orc_hash_value(e, p1, l1, &h1);                /* :265-270 */
            if (o1 >= 0) orc_insert(e, &proto, p1, l1, p2, l2, o1, h1, '+'); else e->bad_ambig++;
            char *rc1 = (char *)malloc((size_t)l1 + 1);             /* :273-281 */
            orc_reverse_complement(e, p1, l1, rc1);
            int o2 = orc_hash_value(e, p2, l2, &h2);
            if (o2 >= 0) orc_insert(e, &proto, p2, l2, rc1, l1, o2, h2, '-'); else e->bad_ambig++;
            free(rc1);
        }
        free(proto.id); free(proto.alias); free(p1); free(p2);
    }
    if (!ok) { orc_clear_sts(e); return 0; }
    orc_build_table(e);
    return 1;
}

static char *read_file(const char *path, size_t *n) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    size_t got = fread(buf, 1, (size_t)sz, fp);
    fclose(fp);
    buf[got] = 0; *n = got;
    return buf;
}

/* returns 1/0 like the reference, -1 if the file cannot be opened (reference raises) */
int orc_load_sts_file(orc_engine *e, const char *path) {
    size_t n; char *t = read_file(path, &n);
    if (!t) return -1;
    int r = orc_load_sts_text(e, t, n);
    free(t);
    return r;
}

int orc_num_records(const orc_engine *e) { return e->nrec; }
long long orc_max_pcr_size(const orc_engine *e) { return e->max_pcr_size; }
const char *orc_rec_id(const orc_engine *e, int i) { return e->recs[i].id; }
const char *orc_rec_alias(const orc_engine *e, int i) { return e->recs[i].alias; }
const char *orc_rec_primer1(const orc_engine *e, int i) { return e->recs[i].primer1; }
const char *orc_rec_primer2(const orc_engine *e, int i) { return e->recs[i].primer2; }
long long orc_rec_pcr_size(const orc_engine *e, int i) { return e->recs[i].pcr_size; }
int orc_rec_hash_offset(const orc_engine *e, int i) { return e->recs[i].hash_offset; }
uint32_t orc_rec_hash(const orc_engine *e, int i) { return e->recs[i].hash; }
int orc_rec_line(const orc_engine *e, int i) { return e->recs[i].offset; }
char orc_rec_direct(const orc_engine *e, int i) { return e->recs[i].direct; }
/* Bulk read-back for large tables (tests map record indices to STS lines): source line number and '+' / '-' of
 * every record, in insertion order (engine.py:265-281). */
void orc_rec_lines(const orc_engine *e, int *lines, char *directs) {
    for (int i = 0; i < e->nrec; i++) { lines[i] = e->recs[i].offset; directs[i] = e->recs[i].direct; }
}
const char *orc_last_error(const orc_engine *e) { return e->err; }

/* ------------------------------------------------------------------ FASTA (io/fasta.py:19-71) */

typedef struct { char *defline, *label, *seq; size_t len; } orc_fa_rec;
typedef struct orc_fasta { orc_fa_rec *r; int n, cap; int error; } orc_fasta;

static int fasta_keep(unsigned char c) {                            /* fasta.py:60 */
    switch (toupper(c)) {
        case 'A': case 'C': case 'G': case 'T': case 'B': case 'D': case 'H': case 'K':
        case 'M': case 'N': case 'R': case 'S': case 'V': case 'W': case 'X': case 'Y': return 1;
        default: return 0;
    }
}

/* models.py:40-49 FASTARecord.__post_init__ ; returns NULL for the IndexError case (bare '>') */
static char *fasta_label(const char *defline, size_t n) {
    py_strip(&defline, &n);
    if (memchr(defline, '>', n)) { if (n) { defline++; n--; } }
    size_t i = 0;
    while (i < n && py_isspace((unsigned char)defline[i])) i++;
    size_t j = i;
    while (j < n && !py_isspace((unsigned char)defline[j])) j++;
    if (j == i) return NULL;
    return xstrndup(defline + i, j - i);
}

static void fasta_push(orc_fasta *fa, char *defline, sbuf *seq) {
    if (fa->n == fa->cap) { fa->cap = fa->cap ? fa->cap * 2 : 8; fa->r = (orc_fa_rec *)realloc(fa->r, (size_t)fa->cap * sizeof(orc_fa_rec)); }
    orc_fa_rec *r = &fa->r[fa->n++];
    r->defline = defline;
    r->label = fasta_label(defline, strlen(defline));
    if (!r->label) fa->error = 1;
    sbuf_reserve(seq, 0);
    r->seq = seq->p; r->len = seq->n;
    seq->p = NULL; seq->n = seq->cap = 0;
}

orc_fasta *orc_load_fasta_text(const char *text, size_t n) {
    orc_fasta *fa = (orc_fasta *)calloc(1, sizeof(*fa));
    if (n == 0) return fa;                                          /* fasta.py:32-34 */
    char *cur_def = NULL; sbuf seq = {0};
    size_t pos = 0;
    while (pos < n) {
        size_t eol = pos;
        while (eol < n && text[eol] != '\n' && text[eol] != '\r') eol++;
        const char *ln = text + pos; size_t ll = eol - pos;
        pos = eol;
        if (pos < n) { if (text[pos] == '\r' && pos + 1 < n && text[pos + 1] == '\n') pos += 2; else pos += 1; }
        py_strip(&ln, &ll);                                         /* :44 */
        if (!ll) continue;                                          /* :46-47 */
        if (ln[0] == '>') {                                         /* :49-57 */
            if (cur_def) fasta_push(fa, cur_def, &seq);
            else { free(seq.p); seq.p = NULL; seq.n = seq.cap = 0; } /* data before the first header is dropped */
            cur_def = xstrndup(ln, ll);
        } else {                                                    /* :58-61 */
            sbuf_reserve(&seq, ll);
            for (size_t i = 0; i < ll; i++) if (fasta_keep((unsigned char)ln[i])) seq.p[seq.n++] = ln[i];
            seq.p[seq.n] = 0;
        }
    }
    if (cur_def) fasta_push(fa, cur_def, &seq); else free(seq.p);   /* :64-66 */
    return fa;
}

orc_fasta *orc_load_fasta_file(const char *path) {
    size_t n; char *t = read_file(path, &n);
    if (!t) return NULL;
    orc_fasta *fa = orc_load_fasta_text(t, n);
    free(t);
    return fa;
}
int orc_fasta_count(const orc_fasta *fa) { return fa->n; }
int orc_fasta_error(const orc_fasta *fa) { return fa->error; }
const char *orc_fasta_label(const orc_fasta *fa, int i) { return fa->r[i].label; }
const char *orc_fasta_defline(const orc_fasta *fa, int i) { return fa->r[i].defline; }
const char *orc_fasta_seq(const orc_fasta *fa, int i) { return fa->r[i].seq; }
size_t orc_fasta_len(const orc_fasta *fa, int i) { return fa->r[i].len; }
void orc_fasta_free(orc_fasta *fa) {
    if (!fa) return;
    for (int i = 0; i < fa->n; i++) { free(fa->r[i].defline); free(fa->r[i].label); free(fa->r[i].seq); }
    free(fa->r); free(fa);
}

/* ------------------------------------------------------------------ search */

typedef struct { int64_t pos1, pos2; int rec; } orc_hit;               /* models.py:52-58 STSHit */
typedef struct { orc_hit *h; size_t n, cap; } hitvec;

static void hit_push(hitvec *v, int64_t p1, int64_t p2, int rec) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->h = (orc_hit *)realloc(v->h, v->cap * sizeof(orc_hit)); }
    v->h[v->n].pos1 = p1; v->h[v->n].pos2 = p2; v->h[v->n].rec = rec; v->n++;
}

/* engine.py:599-642 _compare_seqs ; seq1 = sequence slice (already upper-cased), seq2 = primer */
int orc_compare_seqs(const orc_engine *e, const char *seq1, int len1, const char *seq2, int len2, char strand) {
    if (len1 != len2) return 0;
    int mism = 0;
    for (int i = 0; i < len1; i++) {
        int prot = (strand == '+' && i >= len1 - e->X) || (strand == '-' && i < e->X);   /* :609-611 */
        unsigned char c1 = (unsigned char)toupper((unsigned char)seq1[i]);
        unsigned char c2 = (unsigned char)toupper((unsigned char)seq2[i]);
        int match;
        if (e->I) {                                                     /* :614-629 */
            if (e->iupac[c1] && e->iupac[c2]) match = (e->iupac[c1] & e->iupac[c2]) != 0;
            else match = (c1 == c2);
        } else match = (c1 == c2);                                      /* :631 */
        if (!match) {
            if (prot) return 0;                                         /* :635-636 */
            if (++mism > e->N) return 0;                                /* :638-640 */
        }
    }
    return 1;
}

/* engine.py:507-597 _match_sts */
static void orc_match_sts(const orc_engine *e, const char *seq, int64_t L, int64_t k, int ri, int64_t off, hitvec *out) {
    const orc_sts *r = &e->recs[ri];
    int l1 = r->len1, l2 = r->len2;
    if (!(k + l1 <= L && orc_compare_seqs(e, seq + k, l1, r->primer1, l1, '+'))) return;   /* :515 */
    int64_t exp = r->pcr_size;
    int64_t avail = L - (k + l1);                                       /* :521 */
    if (avail < l2) return;                                             /* :524-525 */
    int64_t actual = avail + l1, hi, lo;                                /* :528 */
    if (exp > actual) { exp = actual; hi = 0; }                         /* :531-533 */
    else { hi = L - k - exp; if (hi > e->M) hi = e->M; }                /* :535 */
    lo = exp - l1 - l2; if (lo > e->M) lo = e->M; if (lo < 0) lo = 0;   /* :538-540 */
    int64_t p2 = k + exp - l2;                                          /* :543 */
    if (k + l1 <= p2 && p2 + l2 <= L)                                   /* :546-555 */
        if (orc_compare_seqs(e, seq + p2, l2, r->primer2, l2, '-')) hit_push(out, k + off, p2 + l2 - 1 + off, ri);
    for (int i = 1; i <= e->M; i++) {                                   /* :563 */
        if (i <= lo) {                                                  /* :565-578 */
            p2 = k + exp - l2 - i;
            if (k + l1 <= p2 && p2 + l2 <= L)
                if (orc_compare_seqs(e, seq + p2, l2, r->primer2, l2, '-')) hit_push(out, k + off, p2 + l2 - 1 + off, ri);
        }
        if (i <= hi) {                                                  /* :581-593 */
            p2 = k + exp - l2 + i;
            if (p2 + l2 <= L)
                if (orc_compare_seqs(e, seq + p2, l2, r->primer2, l2, '-')) hit_push(out, k + off, p2 + l2 - 1 + off, ri);
        }
    }
}

/* engine.py:453-505 _process_thread ; seq must already be upper-cased (:455) */
static void orc_process_chunk(const orc_engine *e, const char *seq, int64_t L, int64_t off, hitvec *out) {
    int W = e->W;
    if (L <= W) return;                                                 /* :458 */
    uint32_t h = 0, mask = (W == 16) ? 0xffffffffu : ((1u << (2 * W)) - 1u);
    int N = 0;
    for (int i = 0; i < W; i++) {                                       /* :467-478 */
        h <<= 2;
        int code = e->scode[(unsigned char)seq[i]];
        if (code == ORC_AMBIG) N = W; else { if (N > 0) N--; h |= (uint32_t)code; }
    }
    h &= mask;
    for (int64_t pos = 0; pos + W <= L; pos++) {                        /* :481 */
        if (N == 0) {                                                   /* :483 */
            for (int ri = orc_bucket_head(e, h); ri >= 0; ri = e->recs[ri].next) {   /* :484 */
                int64_t k = pos - e->recs[ri].hash_offset;              /* :486 */
                if (k >= 0 && k + e->recs[ri].len1 <= L) orc_match_sts(e, seq, L, k, ri, off, out);   /* :487-489 */
            }
        }
        if (pos + W < L) {                                              /* :492-503 */
            h = (h << 2) & mask;
            int code = e->scode[(unsigned char)seq[pos + W]];
            if (code == ORC_AMBIG) N = W; else { if (N > 0) N--; h |= (uint32_t)code; }
        }
    }
}

/* stable merge sort by pos1 (engine.py:434 list.sort(key=pos1) is stable) */
static void hits_sort(orc_hit *a, size_t n) {
    if (n < 2) return;
    orc_hit *tmp = (orc_hit *)malloc(n * sizeof(orc_hit));
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (a[j].pos1 < a[i].pos1) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof(orc_hit));
    }
    free(tmp);
}

typedef struct { const orc_engine *e; const char *seq; int64_t len, off; hitvec hits; } chunk_job;
static void *chunk_main(void *p) {
    chunk_job *j = (chunk_job *)p;
    orc_process_chunk(j->e, j->seq, j->len, j->off, &j->hits);
    return NULL;
}

/*
 * engine.py:373-444 : the per-record body of search().  `threads` follows the reference's rule:
 * records shorter than 100 000 always run serially; chunk arithmetic is :386-411; the overlap
 * "de-duplication" (:425-431) is restated literally (it compares a global pos2 with the overlap
 * length, so it is ineffective -- SURVEY Q9).  Output lines appended to `out` exactly as :442.
 * Returns the number of hits; if hits_out != NULL the sorted hits are handed to the caller.
 */
long long orc_search_record(const orc_engine *e, const char *label, const char *sequence, size_t seq_len,
                            int threads, sbuf *out, orc_hit **hits_out) {
    int64_t L = (int64_t)seq_len;
    char *up = (char *)malloc(seq_len + 1);
    for (size_t i = 0; i < seq_len; i++) up[i] = (char)toupper((unsigned char)sequence[i]);   /* :455 */
    up[seq_len] = 0;
    int nt = threads;
    if (L < ORC_MIN_THREADING) nt = 1;                                  /* :381-384 */
    int64_t overlap = e->max_pcr_size + e->M - 1;                       /* :387 */
    while (nt > 1 && (int64_t)(nt + 1) * overlap > L) nt--;             /* :390-392 */
    /* :395 chunk_size = int((seq_len - (n+1)*overlap)/n) + 2*overlap ; true division then int() truncation */
    double q = (double)(L - (int64_t)(nt + 1) * overlap) / (double)nt;
    int64_t chunk = (int64_t)q + 2 * overlap;
    chunk_job *jobs = (chunk_job *)calloc((size_t)nt, sizeof(chunk_job));
    int64_t off = 0;
    for (int i = 0; i < nt; i++) {                                      /* :401-411 */
        int64_t length = (i < nt - 1) ? chunk : L - off;
        /* Python slicing clamps to the string; a negative start would wrap but cannot occur here */
        int64_t s = off < 0 ? 0 : (off > L ? L : off);
        int64_t epos = off + length; if (epos > L) epos = L; if (epos < s) epos = s;
        jobs[i].e = e; jobs[i].seq = up + s; jobs[i].len = epos - s; jobs[i].off = off;
        off += length - overlap;
    }
    if (nt > 1) {                                                       /* :414-419 */
        pthread_t *th = (pthread_t *)malloc((size_t)nt * sizeof(pthread_t));
        for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, chunk_main, &jobs[i]);
        for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
        free(th);
    } else chunk_main(&jobs[0]);
    hitvec all = {0};
    for (int i = 0; i < nt; i++) {                                      /* :425-431 */
        for (size_t j = 0; j < jobs[i].hits.n; j++) {
            orc_hit *h = &jobs[i].hits.h[j];
            if (jobs[i].off > 0 && h->pos2 < overlap && i > 0) continue;
            hit_push(&all, h->pos1, h->pos2, h->rec);
        }
        free(jobs[i].hits.h);
    }
    free(jobs); free(up);
    hits_sort(all.h, all.n);                                            /* :434 */
    if (out) {
        char line[64];
        for (size_t i = 0; i < all.n; i++) {                            /* :437-444 */
            const orc_sts *r = &e->recs[all.h[i].rec];
            sbuf_put(out, label, strlen(label));
            int m = snprintf(line, sizeof line, "\t%lld..%lld\t", (long long)all.h[i].pos1 + 1, (long long)all.h[i].pos2 + 1);
            sbuf_put(out, line, (size_t)m);
            sbuf_put(out, r->id, strlen(r->id));
            sbuf_put(out, "\t", 1);
            sbuf_put(out, r->alias, strlen(r->alias));
            m = snprintf(line, sizeof line, "\t(%c)\n", r->direct);
            sbuf_put(out, line, (size_t)m);
        }
    }
    long long n = (long long)all.n;
    if (hits_out) *hits_out = all.h; else free(all.h);
    return n;
}

/* -------- ctypes-friendly entry points -------- */

/* Search one in-memory sequence; returns malloc'ed output text (caller frees with orc_free_text). */
char *orc_search_seq_text(const orc_engine *e, const char *label, const char *seq, size_t len, int threads,
                          long long *n_hits) {
    sbuf out = {0}; sbuf_reserve(&out, 0); out.p[0] = 0;
    long long n = orc_search_record(e, label, seq, len, threads, &out, NULL);
    if (n_hits) *n_hits = n;
    return out.p;
}

/* Search one in-memory sequence; returns hits as flat int64 triples (pos1,pos2,rec), 0-based. */
long long orc_search_seq_hits(const orc_engine *e, const char *seq, size_t len, int threads, int64_t **triples) {
    orc_hit *h = NULL;
    long long n = orc_search_record(e, "", seq, len, threads, NULL, &h);
    if (triples) {
        int64_t *t = (int64_t *)malloc((size_t)(n ? n : 1) * 3 * sizeof(int64_t));
        for (long long i = 0; i < n; i++) { t[3 * i] = h[i].pos1; t[3 * i + 1] = h[i].pos2; t[3 * i + 2] = h[i].rec; }
        *triples = t;
    }
    free(h);
    return n;
}

/* Count-only search (timing legs): nothing is formatted. */
long long orc_search_seq_count(const orc_engine *e, const char *seq, size_t len, int threads) {
    return orc_search_record(e, "", seq, len, threads, NULL, NULL);
}

/* Whole-run restatement of cli.py:232-255 + engine.search: returns the output text, or NULL with
 * *status = 1 (the CLI's exit code) when a loader fails. */
char *orc_run_files(orc_engine *e, const char *sts_path, const char *fasta_path, int threads,
                    long long *n_hits, int *status) {
    *status = 0; if (n_hits) *n_hits = 0;
    int r = orc_load_sts_file(e, sts_path);
    if (r != 1) { *status = 1; return NULL; }
    orc_fasta *fa = orc_load_fasta_file(fasta_path);
    if (!fa || fa->error || fa->n == 0) { orc_fasta_free(fa); *status = 1; return NULL; }
    sbuf out = {0}; sbuf_reserve(&out, 0); out.p[0] = 0;
    long long total = 0;
    for (int i = 0; i < fa->n; i++)
        total += orc_search_record(e, fa->r[i].label, fa->r[i].seq, fa->r[i].len, threads, &out, NULL);
    orc_fasta_free(fa);
    if (n_hits) *n_hits = total;
    return out.p;
}

void orc_free_text(void *p) { free(p); }

double orc_now(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
