/*
 * fqz_oracle.c — CPU restatement of fqpack's block codec.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (fastqpacker_b200/libfqzgpu.so) never
 * links, loads or calls anything in oracle/.
 *
 * Parity status: the reference is pure Go and no Go toolchain exists in this image
 * (SURVEY.md F7), so oracle/_ref cannot be built.  The restatement is pinned against
 * every known-answer vector in the reference's own unit tests (tests/test_oracle_kat.py)
 * and against the hand-derived golden streams of testdata/sample.fq (SURVEY.md App. B,
 * tests/golden/).  The entropy stage of the reference is the third-party module
 * github.com/klauspost/compress v1.19.1 (go.mod:8), absent from /root/reference; the
 * oracle stands the system libzstd 1.5.5 in for it (same RFC 8878 format, a different
 * encoder), so COMPRESSED BYTES ARE PARITY-UNPINNED (no reference test pins them either).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/).  Nothing here is copied from the reference: it is C, the
 * reference is Go.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_E_HEADER_AT (-1)   /* "invalid FASTQ: header line must start with @"            parser.go:143 */
#define ORC_E_PLUS (-2)        /* "invalid FASTQ: separator line must start with +"         parser.go:164 */
#define ORC_E_LEN_MISMATCH (-3)/* "invalid FASTQ: sequence and quality lengths must match"  parser.go:180 */
#define ORC_E_LONG_N (-4)      /* "... ambiguous bases beyond position ..."                 compress.go:484 */
#define ORC_E_MAGIC (-5)       /* "invalid magic bytes: not an FQZ file"                    container.go:54 */
#define ORC_E_VERSION (-6)     /* "unsupported file version: %d"                            compress.go:572 */
#define ORC_E_TRUNC_FILE (-7)  /* io.ErrUnexpectedEOF while reading header / payload        compress.go:727,732 */
#define ORC_E_ZSTD (-8)        /* "decompressing <stream>: ..."                             compress.go:787-813 */
#define ORC_E_TRUNC_HEADER (-9)   /* "truncated header data"             compress.go:979 */
#define ORC_E_TRUNC_PLUS (-10)    /* "truncated plus-line payload data"  compress.go:1002 */
#define ORC_E_TRUNC_SEQ (-11)     /* "truncated sequence data"           compress.go:1020 */
#define ORC_E_TRUNC_QUAL (-12)    /* "truncated quality data"            compress.go:1033 */
#define ORC_E_TRUNC_LEN (-13)     /* "truncated length data"             compress.go:1048 */
#define ORC_E_TRUNC_NPOS (-14)    /* "truncated N position data"         compress.go:1057 */
#define ORC_E_NOSPACE (-15)
#define ORC_E_NOLIB (-16)
#define ORC_E_NPOS_RANGE (-17)    /* reference would panic: N position >= seqLen (sequence.go:218-220) */

#define ORC_BLOCK_RECORDS 100000u /* compress.go:71, batchPool compress.go:48-52 */
#define ORC_MAX_SEQ_LEN 65536u    /* sequence.go:11 */

/* ------------------------------------------------------------------ dynamic libs */
typedef size_t (*zstd_bound_fn)(size_t);
typedef void *(*zstd_createcctx_fn)(void);
typedef size_t (*zstd_freecctx_fn)(void *);
typedef size_t (*zstd_setparam_fn)(void *, int, int);
typedef size_t (*zstd_compress2_fn)(void *, void *, size_t, const void *, size_t);
typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
typedef unsigned long long (*zstd_fcs_fn)(const void *, size_t);
typedef size_t (*zstd_findframe_fn)(const void *, size_t);
typedef void *(*zstd_createdctx_fn)(void);
typedef size_t (*zstd_freedctx_fn)(void *);
typedef size_t (*zstd_decompressdctx_fn)(void *, void *, size_t, const void *, size_t);
typedef unsigned long long (*xxh64_fn)(const void *, size_t, unsigned long long);

static struct {
    int loaded;
    zstd_bound_fn bound;
    zstd_createcctx_fn createCCtx;
    zstd_freecctx_fn freeCCtx;
    zstd_setparam_fn setParam;
    zstd_compress2_fn compress2;
    zstd_decompress_fn decompress;
    zstd_iserror_fn isError;
    zstd_fcs_fn frameContentSize;
    zstd_findframe_fn findFrameCompressedSize;
    zstd_createdctx_fn createDCtx;
    zstd_freedctx_fn freeDCtx;
    zstd_decompressdctx_fn decompressDCtx;
    xxh64_fn xxh64;
} L;

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void load_libs_once(void) {
    void *z = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
    void *x = dlopen("libxxhash.so.0", RTLD_NOW | RTLD_GLOBAL);
    if (!z) return;
    L.bound = (zstd_bound_fn)dlsym(z, "ZSTD_compressBound");
    L.createCCtx = (zstd_createcctx_fn)dlsym(z, "ZSTD_createCCtx");
    L.freeCCtx = (zstd_freecctx_fn)dlsym(z, "ZSTD_freeCCtx");
    L.setParam = (zstd_setparam_fn)dlsym(z, "ZSTD_CCtx_setParameter");
    L.compress2 = (zstd_compress2_fn)dlsym(z, "ZSTD_compress2");
    L.decompress = (zstd_decompress_fn)dlsym(z, "ZSTD_decompress");
    L.isError = (zstd_iserror_fn)dlsym(z, "ZSTD_isError");
    L.frameContentSize = (zstd_fcs_fn)dlsym(z, "ZSTD_getFrameContentSize");
    L.findFrameCompressedSize = (zstd_findframe_fn)dlsym(z, "ZSTD_findFrameCompressedSize");
    L.createDCtx = (zstd_createdctx_fn)dlsym(z, "ZSTD_createDCtx");
    L.freeDCtx = (zstd_freedctx_fn)dlsym(z, "ZSTD_freeDCtx");
    L.decompressDCtx = (zstd_decompressdctx_fn)dlsym(z, "ZSTD_decompressDCtx");
    if (x) L.xxh64 = (xxh64_fn)dlsym(x, "XXH64");
    L.loaded = L.bound && L.createCCtx && L.setParam && L.compress2 && L.decompress && L.isError &&
               L.frameContentSize && L.findFrameCompressedSize;
}
static int libs(void) {
    pthread_once(&g_once, load_libs_once);
    return L.loaded ? 0 : ORC_E_NOLIB;
}

/* ------------------------------------------------------------------ growable buffer */
typedef struct {
    uint8_t *p;
    size_t n, cap;
} buf_t;
static void buf_reserve(buf_t *b, size_t extra) {
    if (b->n + extra <= b->cap) return;
    size_t c = b->cap ? b->cap : 4096;
    while (c < b->n + extra) c *= 2;
    b->p = (uint8_t *)realloc(b->p, c);
    b->cap = c;
}
static void buf_put(buf_t *b, const void *src, size_t n) {
    buf_reserve(b, n);
    if (n) memcpy(b->p + b->n, src, n);
    b->n += n;
}
static void buf_u16(buf_t *b, uint16_t v) {
    uint8_t t[2] = {(uint8_t)v, (uint8_t)(v >> 8)};
    buf_put(b, t, 2);
}
static void buf_u32(buf_t *b, uint32_t v) {
    uint8_t t[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    buf_put(b, t, 4);
}
static uint32_t rd_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint16_t rd_u16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ================================================================== encoder/quality.go */

/* quality.go:22-49 DetectEncoding.  quals = concatenated quality bytes of the first batch.
 * returns 0 = Phred33, 1 = Phred64. */
int orc_detect_encoding(const uint8_t *quals, size_t n) {
    unsigned minb = 255;
    for (size_t i = 0; i < n; i++) {
        if (quals[i] < minb) minb = quals[i];
        if (quals[i] < 59) return 0; /* quality.go:31 early exit */
    }
    if (minb == 255) return 0; /* quality.go:38 (no bytes, or only 0xFF bytes) */
    if (minb >= 64) return 1;  /* quality.go:43 */
    return 0;                  /* quality.go:48 ambiguous 59..63 */
}
/* quality.go:53-62 / 66-75 */
void orc_normalize_quality(uint8_t *q, size_t n, int phred64) {
    uint8_t off = phred64 ? 64 : 33;
    for (size_t i = 0; i < n; i++) q[i] = (uint8_t)(q[i] - off);
}
void orc_denormalize_quality(uint8_t *q, size_t n, int phred64) {
    uint8_t off = phred64 ? 64 : 33;
    for (size_t i = 0; i < n; i++) q[i] = (uint8_t)(q[i] + off);
}
/* quality.go:81-103 DeltaEncode (in place, backwards), quality.go:107-118 DeltaDecode */
void orc_delta_encode(uint8_t *q, size_t n) {
    if (n <= 1) return;
    for (size_t i = n - 1; i > 0; i--) q[i] = (uint8_t)(q[i] - q[i - 1]);
}
void orc_delta_decode(uint8_t *q, size_t n) {
    if (n <= 1) return;
    uint8_t acc = q[0];
    for (size_t i = 1; i < n; i++) {
        acc = (uint8_t)(acc + q[i]);
        q[i] = acc;
    }
}

/* ================================================================== encoder/sequence.go */

static inline int base_code(uint8_t b) { /* sequence.go:23-32 baseLookup */
    switch (b) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}
static inline int is_n_base(uint8_t b) { /* sequence.go:44-50 isNBase */
    switch (b) {
    case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 0;
    default: return 1;
    }
}
/* sequence.go:139-184 AppendPackedBases.  packed must hold (n+3)/4 bytes, npos up to
 * min(n,65536) entries.  Returns number of N positions. */
size_t orc_pack_bases(const uint8_t *seq, size_t n, uint8_t *packed, uint16_t *npos) {
    size_t plen = (n + 3) >> 2;
    memset(packed, 0, plen);
    for (size_t i = 0; i < n; i++) packed[i >> 2] |= (uint8_t)(base_code(seq[i]) << ((i & 3) << 1));
    size_t limit = n > ORC_MAX_SEQ_LEN ? ORC_MAX_SEQ_LEN : n; /* sequence.go:173-176 */
    size_t k = 0;
    for (size_t i = 0; i < limit; i++)
        if (is_n_base(seq[i])) npos[k++] = (uint16_t)i;
    return k;
}
/* sequence.go:188-223 AppendUnpackBases.  returns 0 or ORC_E_NPOS_RANGE (reference panics). */
int orc_unpack_bases(const uint8_t *packed, const uint16_t *npos, size_t nn, size_t seqlen, uint8_t *out) {
    static const char bases[4] = {'A', 'C', 'G', 'T'};
    for (size_t i = 0; i < seqlen; i++) out[i] = (uint8_t)bases[(packed[i >> 2] >> ((i & 3) << 1)) & 3];
    for (size_t k = 0; k < nn; k++) {
        if (npos[k] >= seqlen) return ORC_E_NPOS_RANGE;
        out[npos[k]] = 'N';
    }
    return 0;
}

/* ================================================================== fqparser/parser.go */

typedef struct {
    const uint8_t *hdr, *seq, *plus, *qual;
    size_t hlen, slen, plen, qlen;
} rec_t;

/* parser.go:209-220 readLine over a memory buffer: next line minus '\n' and one '\r';
 * returns 0 on EOF-before-newline (the partial line is discarded). */
static int read_line(const uint8_t *t, size_t n, size_t *pos, const uint8_t **line, size_t *len) {
    if (*pos >= n) return 0;
    const uint8_t *nl = (const uint8_t *)memchr(t + *pos, '\n', n - *pos);
    if (!nl) {
        *pos = n;
        return 0;
    }
    size_t l = (size_t)(nl - (t + *pos));
    *line = t + *pos;
    *pos += l + 1;
    if (l > 0 && (*line)[l - 1] == '\r') l--;
    *len = l;
    return 1;
}
/* parser.go:136-184 nextInto.  returns 1 record, 0 EOF, <0 error */
static int next_record(const uint8_t *t, size_t n, size_t *pos, rec_t *r) {
    const uint8_t *ln;
    size_t l;
    if (!read_line(t, n, pos, &ln, &l)) return 0;
    if (l == 0 || ln[0] != '@') return ORC_E_HEADER_AT;
    r->hdr = ln + 1;
    r->hlen = l - 1;
    if (!read_line(t, n, pos, &ln, &l)) return 0;
    r->seq = ln;
    r->slen = l;
    if (!read_line(t, n, pos, &ln, &l)) return 0;
    if (l == 0 || ln[0] != '+') return ORC_E_PLUS;
    r->plus = ln + 1;
    r->plen = l - 1;
    if (!read_line(t, n, pos, &ln, &l)) return 0;
    r->qual = ln;
    r->qlen = l;
    if (r->slen != r->qlen) return ORC_E_LEN_MISMATCH;
    return 1;
}

/* Parse up to max_records records starting at *pos (parser.go:188-205 ReadBatch).
 * recs must hold max_records entries.  *nrec = records read, *pos advanced past them.
 * returns 0, or the parse error (records before the error are still counted in *nrec,
 * but ReadBatch returns the error and the caller aborts: compress.go:141-144,341-344). */
static int read_batch(const uint8_t *t, size_t n, size_t *pos, rec_t *recs, size_t max_records, size_t *nrec) {
    *nrec = 0;
    for (size_t i = 0; i < max_records; i++) {
        size_t save = *pos;
        int rc = next_record(t, n, pos, &recs[i]);
        if (rc == 0) {
            *pos = n;
            (void)save;
            return 0;
        }
        if (rc < 0) return rc;
        *nrec = i + 1;
    }
    return 0;
}

/* Public: count + locate records (for tests of parser semantics).
 * out_fields: per record 8 x uint64: hdr_off,hdr_len,seq_off,seq_len,plus_off,plus_len,qual_off,qual_len */
int orc_parse(const uint8_t *text, size_t n, size_t max_records, uint64_t *out_fields, size_t *nrec, size_t *consumed) {
    rec_t *recs = (rec_t *)malloc(sizeof(rec_t) * (max_records ? max_records : 1));
    size_t pos = 0;
    int rc = read_batch(text, n, &pos, recs, max_records, nrec);
    if (out_fields)
        for (size_t i = 0; i < *nrec; i++) {
            out_fields[8 * i + 0] = (uint64_t)(recs[i].hdr - text);
            out_fields[8 * i + 1] = recs[i].hlen;
            out_fields[8 * i + 2] = (uint64_t)(recs[i].seq - text);
            out_fields[8 * i + 3] = recs[i].slen;
            out_fields[8 * i + 4] = (uint64_t)(recs[i].plus - text);
            out_fields[8 * i + 5] = recs[i].plen;
            out_fields[8 * i + 6] = (uint64_t)(recs[i].qual - text);
            out_fields[8 * i + 7] = recs[i].qlen;
        }
    if (consumed) *consumed = pos;
    free(recs);
    return rc;
}

/* ================================================================== compress.go block encode */

typedef struct {
    buf_t s[6]; /* 0 seqPacked 1 quality 2 headers 3 plusLines 4 nPositions 5 seqLengths */
    uint32_t orig_seq, orig_qual;
} streams_t;

static void streams_free(streams_t *st) {
    for (int i = 0; i < 6; i++) free(st->s[i].p);
    memset(st, 0, sizeof *st);
}

/* compress.go:474-520: per-record stream assembly */
static int encode_records(const rec_t *recs, size_t nrec, int phred64, streams_t *st, size_t *bad_record) {
    uint16_t *npos = (uint16_t *)malloc(sizeof(uint16_t) * ORC_MAX_SEQ_LEN);
    for (size_t i = 0; i < nrec; i++) {
        const rec_t *r = &recs[i];
        if (r->slen > ORC_MAX_SEQ_LEN) { /* compress.go:480-488 */
            for (size_t k = ORC_MAX_SEQ_LEN; k < r->slen; k++)
                if (is_n_base(r->seq[k])) {
                    if (bad_record) *bad_record = i;
                    free(npos);
                    return ORC_E_LONG_N;
                }
        }
        size_t plen = (r->slen + 3) >> 2;
        buf_reserve(&st->s[0], plen);
        size_t nn = 0;
        if (r->slen) { /* sequence.go:141-143: empty sequence appends nothing */
            nn = orc_pack_bases(r->seq, r->slen, st->s[0].p + st->s[0].n, npos);
            st->s[0].n += plen;
        }
        buf_u16(&st->s[4], (uint16_t)nn); /* compress.go:495 (u16 truncation is silent) */
        for (size_t k = 0; k < nn; k++) buf_u16(&st->s[4], npos[k]);
        buf_u32(&st->s[5], (uint32_t)r->slen); /* compress.go:501 */
        st->orig_seq += (uint32_t)r->slen;
        size_t q0 = st->s[1].n; /* compress.go:506-511 */
        buf_put(&st->s[1], r->qual, r->qlen);
        orc_normalize_quality(st->s[1].p + q0, r->qlen, phred64);
        orc_delta_encode(st->s[1].p + q0, r->qlen);
        st->orig_qual += (uint32_t)r->qlen;
        buf_u16(&st->s[2], (uint16_t)r->hlen); /* compress.go:514-515 */
        buf_put(&st->s[2], r->hdr, r->hlen);
        buf_u16(&st->s[3], (uint16_t)r->plen); /* compress.go:518-519 */
        buf_put(&st->s[3], r->plus, r->plen);
    }
    free(npos);
    return 0;
}

/* compress.go:146-154 + quality.go:22-49 over parsed records (order-independent result) */
static int detect_records(const rec_t *recs, size_t nrec) {
    unsigned minb = 255;
    for (size_t i = 0; i < nrec; i++)
        for (size_t k = 0; k < recs[i].qlen; k++) {
            if (recs[i].qual[k] < minb) minb = recs[i].qual[k];
            if (recs[i].qual[k] < 59) return 0;
        }
    return (minb != 255 && minb >= 64) ? 1 : 0;
}

/* Public: FASTQ chunk -> six pre-entropy streams of its first block (<= max_records records).
 * phred64: 0/1 forced, -1 = detect on these records (compress.go:146-154).
 * out[i]/cap[i] caller buffers; len[i] receives sizes (required size if ORC_E_NOSPACE).
 * info: [0]=nrec [1]=consumed bytes [2]=phred64 used [3]=orig_seq [4]=orig_qual [5]=bad record */
int orc_encode_streams(const uint8_t *text, size_t n, size_t max_records, int phred64, uint8_t *const out[6],
                       const size_t cap[6], size_t len[6], uint64_t info[6]) {
    if (max_records == 0) max_records = ORC_BLOCK_RECORDS;
    rec_t *recs = (rec_t *)malloc(sizeof(rec_t) * max_records);
    size_t pos = 0, nrec = 0;
    int rc = read_batch(text, n, &pos, recs, max_records, &nrec);
    if (rc < 0) {
        info[0] = nrec;
        free(recs);
        return rc;
    }
    if (phred64 < 0) phred64 = detect_records(recs, nrec); /* compress.go:146-154 */
    streams_t st;
    memset(&st, 0, sizeof st);
    size_t bad = 0;
    rc = encode_records(recs, nrec, phred64, &st, &bad);
    info[0] = nrec;
    info[1] = pos;
    info[2] = (uint64_t)phred64;
    info[3] = st.orig_seq;
    info[4] = st.orig_qual;
    info[5] = bad;
    if (rc == 0) {
        for (int i = 0; i < 6; i++) {
            len[i] = st.s[i].n;
            if (st.s[i].n > cap[i]) rc = ORC_E_NOSPACE;
        }
        if (rc == 0)
            for (int i = 0; i < 6; i++)
                if (st.s[i].n) memcpy(out[i], st.s[i].p, st.s[i].n);
    }
    streams_free(&st);
    free(recs);
    return rc;
}

/* ================================================================== zstd stand-in (libzstd 1.5.5) */

size_t orc_zstd_bound(size_t n) { return libs() ? 0 : L.bound(n); }

/* klauspost EncodeAll(nil-or-empty) yields zero bytes; mirror that. */
static int zstd_encode(void *cctx, const uint8_t *src, size_t n, buf_t *out) {
    if (n == 0) return 0;
    size_t b = L.bound(n);
    buf_reserve(out, b);
    size_t r = L.compress2(cctx, out->p + out->n, b, src, n);
    if (L.isError(r)) return ORC_E_ZSTD;
    out->n += r;
    return 0;
}
static void *new_cctx(int level) {
    void *c = L.createCCtx();
    L.setParam(c, 100, level); /* ZSTD_c_compressionLevel; reference: SpeedFastest (compress.go:116) */
    L.setParam(c, 201, 1);     /* ZSTD_c_checksumFlag: klauspost default CRC on */
    L.setParam(c, 200, 1);     /* ZSTD_c_contentSizeFlag */
    return c;
}
int orc_zstd_compress(const uint8_t *src, size_t n, int level, uint8_t *dst, size_t cap, size_t *out_len) {
    if (libs()) return ORC_E_NOLIB;
    void *c = new_cctx(level);
    buf_t b = {0};
    int rc = zstd_encode(c, src, n, &b);
    L.freeCCtx(c);
    if (rc == 0) {
        *out_len = b.n;
        if (b.n > cap) rc = ORC_E_NOSPACE;
        else if (b.n) memcpy(dst, b.p, b.n);
    }
    free(b.p);
    return rc;
}
/* TEST SUPPORT: frames shaped like other encoders' (klauspost's EncodeAll / streaming writer are not in the reference
 * tree): explicit window log (4 MiB windows with matches across > 100 blocks), with or without the frame content size
 * (a streaming encoder writes none) and the checksum.  window_log 0 / flags < 0 leave libzstd's defaults. */
int orc_zstd_compress_adv(const uint8_t *src, size_t n, int level, int window_log, int content_size, int checksum, uint8_t *dst, size_t cap,
                          size_t *out_len) {
    if (libs()) return ORC_E_NOLIB;
    void *c = L.createCCtx();
    L.setParam(c, 100, level);
    if (window_log > 0) L.setParam(c, 101, window_log); /* ZSTD_c_windowLog */
    if (content_size >= 0) L.setParam(c, 200, content_size);
    if (checksum >= 0) L.setParam(c, 201, checksum);
    buf_t b = {0};
    int rc = zstd_encode(c, src, n, &b);
    L.freeCCtx(c);
    if (rc == 0) {
        *out_len = b.n;
        if (b.n > cap) rc = ORC_E_NOSPACE;
        else if (b.n) memcpy(dst, b.p, b.n);
    }
    free(b.p);
    return rc;
}
/* DecodeAll: concatenated frames decode back to back; zero bytes -> empty. */
static int zstd_decode_all(const uint8_t *src, size_t n, buf_t *out) {
    size_t pos = 0;
    while (pos < n) {
        size_t fsz = L.findFrameCompressedSize(src + pos, n - pos);
        if (L.isError(fsz)) return ORC_E_ZSTD;
        unsigned long long cs = L.frameContentSize(src + pos, fsz);
        if (cs == (unsigned long long)-2) return ORC_E_ZSTD; /* CONTENTSIZE_ERROR */
        if (cs == (unsigned long long)-1) {                  /* unknown: bound by 128 KiB per block */
            cs = (unsigned long long)fsz * 200 + (1u << 17);
            if (cs > (1ull << 32)) cs = 1ull << 32;
        }
        buf_reserve(out, (size_t)cs + 1);
        size_t r = L.decompress(out->p + out->n, (size_t)cs + 1, src + pos, fsz);
        if (L.isError(r)) return ORC_E_ZSTD;
        out->n += r;
        pos += fsz;
    }
    return 0;
}
int orc_zstd_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    if (libs()) return ORC_E_NOLIB;
    buf_t b = {0};
    int rc = zstd_decode_all(src, n, &b);
    if (rc == 0) {
        *out_len = b.n;
        if (b.n > cap) rc = ORC_E_NOSPACE;
        else if (b.n) memcpy(dst, b.p, b.n);
    }
    free(b.p);
    return rc;
}
unsigned long long orc_xxh64(const uint8_t *p, size_t n, unsigned long long seed) {
    if (libs() || !L.xxh64) return 0;
    return L.xxh64(p, n, seed);
}

/* ================================================================== container.go + block write */

/* container.go:35-45 */
static void write_file_header(buf_t *o, uint8_t version, uint32_t block_size, uint8_t flags) {
    static const uint8_t magic[4] = {'F', 'Q', 'Z', 0};
    buf_put(o, magic, 4);
    buf_put(o, &version, 1);
    buf_u32(o, block_size);
    buf_put(o, &flags, 1);
}

/* compress.go:523-552: six EncodeAll + header + payloads (v2) ; v1 drops plus (container.go:85-96) */
static int write_block(void *cctx, streams_t *st, uint32_t nrec, int version, buf_t *o) {
    buf_t c[6];
    memset(c, 0, sizeof c);
    int rc = 0;
    for (int i = 0; i < 6 && rc == 0; i++) {
        if (version == 1 && i == 3) continue;
        rc = zstd_encode(cctx, st->s[i].p, st->s[i].n, &c[i]);
    }
    if (rc == 0) {
        buf_u32(o, nrec);
        buf_u32(o, (uint32_t)c[0].n);
        buf_u32(o, (uint32_t)c[1].n);
        buf_u32(o, (uint32_t)c[2].n);
        if (version >= 2) buf_u32(o, (uint32_t)c[3].n);
        buf_u32(o, (uint32_t)c[4].n);
        buf_u32(o, (uint32_t)c[5].n);
        buf_u32(o, st->orig_seq);
        buf_u32(o, st->orig_qual);
        for (int i = 0; i < 6; i++) {
            if (version == 1 && i == 3) continue;
            buf_put(o, c[i].p, c[i].n);
        }
    }
    for (int i = 0; i < 6; i++) free(c[i].p);
    return rc;
}

/* ---- whole-file compress: compress.go:125-238 (single worker order; the parallel path
 * produces the same bytes because blocks are written strictly by seqNum, compress.go:380-399).
 * threads > 1 runs block encoders on a pthread pool after a serial parse, like
 * compress.go:240-363 (one producer, W workers, ordered collector). */

typedef struct {
    const uint8_t *text;
    size_t n;
    rec_t *recs;       /* all records */
    size_t nrec;
    size_t nblocks;
    int phred64, level, version;
    buf_t *outs;       /* per block */
    int *rcs;
    size_t *bad;
    size_t next;       /* work counter */
    pthread_mutex_t mu;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    void *cctx = new_cctx(j->level); /* one encoder per worker: compress.go:281 */
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t b = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->nblocks) break;
        size_t r0 = b * ORC_BLOCK_RECORDS;
        size_t r1 = r0 + ORC_BLOCK_RECORDS;
        if (r1 > j->nrec) r1 = j->nrec;
        streams_t st;
        memset(&st, 0, sizeof st);
        size_t bad = 0;
        int rc = encode_records(j->recs + r0, r1 - r0, j->phred64, &st, &bad);
        if (rc == 0) rc = write_block(cctx, &st, (uint32_t)(r1 - r0), j->version, &j->outs[b]);
        j->rcs[b] = rc;
        j->bad[b] = r0 + bad;
        streams_free(&st);
    }
    L.freeCCtx(cctx);
    return NULL;
}

/* info: [0]=records [1]=blocks [2]=phred64 [3]=first bad record index (on error) */
int orc_compress(const uint8_t *text, size_t n, uint32_t header_block_size, int level, int threads, int version,
                 uint8_t *out, size_t cap, size_t *out_len, uint64_t info[4]) {
    if (libs()) return ORC_E_NOLIB;
    if (header_block_size == 0) header_block_size = ORC_BLOCK_RECORDS; /* compress.go:129-131 */
    if (version == 0) version = 2;
    if (threads < 1) threads = 1;
    /* serial parse of the whole input (the reference's single producer) */
    size_t cap_recs = n / 8 + 16, nrec = 0, pos = 0;
    rec_t *recs = (rec_t *)malloc(sizeof(rec_t) * cap_recs);
    int perr = 0;
    for (;;) {
        if (nrec == cap_recs) {
            cap_recs *= 2;
            recs = (rec_t *)realloc(recs, sizeof(rec_t) * cap_recs);
        }
        int rc = next_record(text, n, &pos, &recs[nrec]);
        if (rc == 0) break;
        if (rc < 0) {
            perr = rc;
            break;
        }
        nrec++;
    }
    if (info) {
        info[0] = nrec;
        info[3] = nrec;
    }
    if (perr) { /* "parsing FASTQ: ..." compress.go:143,222,343 */
        free(recs);
        return perr;
    }
    /* Phred detection on the first batch only (compress.go:146-154) */
    int phred64 = detect_records(recs, nrec < ORC_BLOCK_RECORDS ? nrec : ORC_BLOCK_RECORDS);
    buf_t o = {0};
    write_file_header(&o, (uint8_t)version, header_block_size, phred64 ? 2 : 0); /* compress.go:157-168 */
    job_t j;
    memset(&j, 0, sizeof j);
    j.text = text;
    j.n = n;
    j.recs = recs;
    j.nrec = nrec;
    j.nblocks = (nrec + ORC_BLOCK_RECORDS - 1) / ORC_BLOCK_RECORDS;
    j.phred64 = phred64;
    j.level = level;
    j.version = version;
    j.outs = (buf_t *)calloc(j.nblocks ? j.nblocks : 1, sizeof(buf_t));
    j.rcs = (int *)calloc(j.nblocks ? j.nblocks : 1, sizeof(int));
    j.bad = (size_t *)calloc(j.nblocks ? j.nblocks : 1, sizeof(size_t));
    pthread_mutex_init(&j.mu, NULL);
    if ((size_t)threads > j.nblocks) threads = (int)(j.nblocks ? j.nblocks : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    int rc = 0;
    for (size_t b = 0; b < j.nblocks; b++) {
        if (j.rcs[b] && !rc) {
            rc = j.rcs[b];
            if (info) info[3] = j.bad[b];
        }
        if (!rc) buf_put(&o, j.outs[b].p, j.outs[b].n);
        free(j.outs[b].p);
    }
    if (info) {
        info[1] = j.nblocks;
        info[2] = (uint64_t)phred64;
    }
    free(j.outs);
    free(j.rcs);
    free(j.bad);
    free(recs);
    pthread_mutex_destroy(&j.mu);
    if (rc == 0) {
        *out_len = o.n;
        if (o.n > cap) rc = ORC_E_NOSPACE;
        else memcpy(out, o.p, o.n);
    }
    free(o.p);
    return rc;
}

/* ================================================================== decode */

/* compress.go:944-1078 writeRecord and helpers, over six decoded streams.
 * have_plus = 0 reproduces "len(plusData)==0 -> '+\n'" (compress.go:995-999) */
static int decode_records(const buf_t s[6], uint32_t nrec, int phred64, buf_t *o) {
    size_t so = 0, qo = 0, ho = 0, po = 0, no = 0, lo = 0;
    uint16_t *npos = (uint16_t *)malloc(sizeof(uint16_t) * 65536);
    int rc = 0;
    for (uint32_t r = 0; r < nrec; r++) {
        if (lo + 4 > s[5].n) { rc = ORC_E_TRUNC_LEN; break; }
        size_t L0 = rd_u32(s[5].p + lo);
        lo += 4;
        if (no + 2 > s[4].n) { rc = ORC_E_TRUNC_NPOS; break; }
        size_t nn = rd_u16(s[4].p + no);
        no += 2;
        if (no + 2 * nn > s[4].n) { rc = ORC_E_TRUNC_NPOS; break; }
        for (size_t k = 0; k < nn; k++) npos[k] = rd_u16(s[4].p + no + 2 * k);
        no += 2 * nn;
        if (ho + 2 > s[2].n) { rc = ORC_E_TRUNC_HEADER; break; }
        size_t hl = rd_u16(s[2].p + ho);
        ho += 2;
        if (ho + hl > s[2].n) { rc = ORC_E_TRUNC_HEADER; break; }
        buf_reserve(o, hl + 2 * L0 + 70000);
        o->p[o->n++] = '@';
        memcpy(o->p + o->n, s[2].p + ho, hl);
        o->n += hl;
        o->p[o->n++] = '\n';
        ho += hl;
        size_t pl = (L0 + 3) / 4;
        if (so + pl > s[0].n) { rc = ORC_E_TRUNC_SEQ; break; }
        rc = orc_unpack_bases(s[0].p + so, npos, nn, L0, o->p + o->n);
        if (rc) break;
        o->n += L0;
        o->p[o->n++] = '\n';
        so += pl;
        if (s[3].n == 0) {
            o->p[o->n++] = '+';
            o->p[o->n++] = '\n';
        } else {
            if (po + 2 > s[3].n) { rc = ORC_E_TRUNC_PLUS; break; }
            size_t ql = rd_u16(s[3].p + po);
            po += 2;
            if (po + ql > s[3].n) { rc = ORC_E_TRUNC_PLUS; break; }
            buf_reserve(o, ql + L0 + 16);
            o->p[o->n++] = '+';
            memcpy(o->p + o->n, s[3].p + po, ql);
            o->n += ql;
            o->p[o->n++] = '\n';
            po += ql;
        }
        if (qo + L0 > s[1].n) { rc = ORC_E_TRUNC_QUAL; break; }
        buf_reserve(o, L0 + 2);
        memcpy(o->p + o->n, s[1].p + qo, L0);
        orc_delta_decode(o->p + o->n, L0);
        orc_denormalize_quality(o->p + o->n, L0, phred64);
        o->n += L0;
        o->p[o->n++] = '\n';
        qo += L0;
    }
    free(npos);
    return rc;
}

/* Public: six pre-entropy streams -> FASTQ text (back end only). */
int orc_decode_streams(const uint8_t *const in[6], const size_t len[6], uint32_t nrec, int phred64, uint8_t *out,
                       size_t cap, size_t *out_len) {
    buf_t s[6];
    for (int i = 0; i < 6; i++) {
        s[i].p = (uint8_t *)in[i];
        s[i].n = s[i].cap = len[i];
    }
    buf_t o = {0};
    int rc = decode_records(s, nrec, phred64, &o);
    if (rc == 0) {
        *out_len = o.n;
        if (o.n > cap) rc = ORC_E_NOSPACE;
        else if (o.n) memcpy(out, o.p, o.n);
    }
    free(o.p);
    return rc;
}

/* Whole-file decode: compress.go:558-628 (single worker order).
 * info: [0]=version [1]=flags [2]=blocks [3]=records [4]=header BlockSize */
int orc_decompress(const uint8_t *f, size_t n, uint8_t *out, size_t cap, size_t *out_len, uint64_t info[5]) {
    if (libs()) return ORC_E_NOLIB;
    if (n < 4) return ORC_E_TRUNC_FILE;
    if (!(f[0] == 'F' && f[1] == 'Q' && f[2] == 'Z' && f[3] == 0)) return ORC_E_MAGIC; /* container.go:53-55 */
    if (n < 10) return ORC_E_TRUNC_FILE;
    int version = f[4];
    uint8_t flags = f[9];
    if (info) {
        info[0] = (uint64_t)version;
        info[1] = flags;
        info[2] = info[3] = 0;
        info[4] = rd_u32(f + 5);
    }
    if (version != 1 && version != 2) return ORC_E_VERSION; /* compress.go:571-573 */
    int phred64 = (flags & 2) ? 1 : 0;                      /* compress.go:576-579 */
    size_t hsz = version == 1 ? 32 : 36;
    size_t pos = 10;
    buf_t o = {0};
    int rc = 0;
    while (pos < n) {
        if (pos + hsz > n) { rc = ORC_E_TRUNC_FILE; break; }
        const uint8_t *h = f + pos;
        uint32_t nrec = rd_u32(h);
        uint32_t sz[6];
        if (version == 1) {
            sz[0] = rd_u32(h + 4); sz[1] = rd_u32(h + 8); sz[2] = rd_u32(h + 12); sz[3] = 0;
            sz[4] = rd_u32(h + 16); sz[5] = rd_u32(h + 20);
        } else {
            for (int i = 0; i < 6; i++) sz[i] = rd_u32(h + 4 + 4 * i);
        }
        pos += hsz;
        buf_t s[6];
        memset(s, 0, sizeof s);
        for (int i = 0; i < 6 && rc == 0; i++) { /* payload order seq,qual,hdr,[plus],npos,len: compress.go:738-751 */
            if (version == 1 && i == 3) continue;
            if (pos + sz[i] > n) { rc = ORC_E_TRUNC_FILE; break; }
            rc = zstd_decode_all(f + pos, sz[i], &s[i]);
            pos += sz[i];
        }
        if (rc == 0) rc = decode_records(s, nrec, phred64, &o);
        for (int i = 0; i < 6; i++) free(s[i].p);
        if (rc) break;
        if (info) {
            info[2]++;
            info[3] += nrec;
        }
    }
    if (rc == 0) {
        *out_len = o.n;
        if (o.n > cap) rc = ORC_E_NOSPACE;
        else if (o.n) memcpy(out, o.p, o.n);
    }
    free(o.p);
    return rc;
}

/* Parallel whole-file decode: compress.go:630-719 (serial container walk = the single producer,
 * W workers each decoding whole blocks, output concatenated in block order = the ordered collector).
 * Used by bench.py as the multi-core CPU baseline; results equal orc_decompress. */
typedef struct {
    const uint8_t *f;
    size_t nblocks;
    size_t *off;      /* payload offset per block */
    uint32_t *nrec;
    uint32_t (*sz)[6];
    int version, phred64;
    buf_t *outs;
    int *rcs;
    size_t next;
    pthread_mutex_t mu;
} djob_t;

static void *dworker(void *arg) {
    djob_t *j = (djob_t *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t b = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->nblocks) break;
        buf_t s[6];
        memset(s, 0, sizeof s);
        size_t pos = j->off[b];
        int rc = 0;
        for (int i = 0; i < 6 && rc == 0; i++) {
            if (j->version == 1 && i == 3) continue;
            rc = zstd_decode_all(j->f + pos, j->sz[b][i], &s[i]);
            pos += j->sz[b][i];
        }
        if (rc == 0) rc = decode_records(s, j->nrec[b], j->phred64, &j->outs[b]);
        for (int i = 0; i < 6; i++) free(s[i].p);
        j->rcs[b] = rc;
    }
    return NULL;
}

int orc_decompress_mt(const uint8_t *f, size_t n, int threads, uint8_t *out, size_t cap, size_t *out_len) {
    if (libs()) return ORC_E_NOLIB;
    if (n < 4) return ORC_E_TRUNC_FILE;
    if (!(f[0] == 'F' && f[1] == 'Q' && f[2] == 'Z' && f[3] == 0)) return ORC_E_MAGIC;
    if (n < 10) return ORC_E_TRUNC_FILE;
    int version = f[4];
    if (version != 1 && version != 2) return ORC_E_VERSION;
    size_t hsz = version == 1 ? 32 : 36, pos = 10, nb = 0, capb = 64;
    djob_t j;
    memset(&j, 0, sizeof j);
    j.off = (size_t *)malloc(sizeof(size_t) * capb);
    j.nrec = (uint32_t *)malloc(sizeof(uint32_t) * capb);
    j.sz = (uint32_t(*)[6])malloc(sizeof(uint32_t[6]) * capb);
    int rc = 0;
    while (pos < n) {
        if (pos + hsz > n) { rc = ORC_E_TRUNC_FILE; break; }
        if (nb == capb) {
            capb *= 2;
            j.off = (size_t *)realloc(j.off, sizeof(size_t) * capb);
            j.nrec = (uint32_t *)realloc(j.nrec, sizeof(uint32_t) * capb);
            j.sz = (uint32_t(*)[6])realloc(j.sz, sizeof(uint32_t[6]) * capb);
        }
        const uint8_t *h = f + pos;
        j.nrec[nb] = rd_u32(h);
        if (version == 1) {
            j.sz[nb][0] = rd_u32(h + 4); j.sz[nb][1] = rd_u32(h + 8); j.sz[nb][2] = rd_u32(h + 12); j.sz[nb][3] = 0;
            j.sz[nb][4] = rd_u32(h + 16); j.sz[nb][5] = rd_u32(h + 20);
        } else
            for (int i = 0; i < 6; i++) j.sz[nb][i] = rd_u32(h + 4 + 4 * i);
        pos += hsz;
        j.off[nb] = pos;
        size_t pay = 0;
        for (int i = 0; i < 6; i++) pay += j.sz[nb][i];
        if (pos + pay > n) { rc = ORC_E_TRUNC_FILE; break; }
        pos += pay;
        nb++;
    }
    j.f = f;
    j.nblocks = nb;
    j.version = version;
    j.phred64 = (f[9] & 2) ? 1 : 0;
    j.outs = (buf_t *)calloc(nb ? nb : 1, sizeof(buf_t));
    j.rcs = (int *)calloc(nb ? nb : 1, sizeof(int));
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    if ((size_t)threads > nb) threads = (int)(nb ? nb : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dworker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    size_t total = 0;
    for (size_t b = 0; b < nb; b++) {
        if (j.rcs[b] && !rc) rc = j.rcs[b];
        total += j.outs[b].n;
    }
    if (rc == 0) {
        *out_len = total;
        if (total > cap) rc = ORC_E_NOSPACE;
        else {
            size_t o = 0;
            for (size_t b = 0; b < nb; b++) {
                if (j.outs[b].n) memcpy(out + o, j.outs[b].p, j.outs[b].n);
                o += j.outs[b].n;
            }
        }
    }
    for (size_t b = 0; b < nb; b++) free(j.outs[b].p);
    free(j.outs); free(j.rcs); free(j.off); free(j.nrec); free(j.sz);
    pthread_mutex_destroy(&j.mu);
    return rc;
}

/* Split one block of a .fqz into its six decoded streams (for checking GPU-written frames
 * with libzstd and for feeding reference-shaped streams to the GPU back end).
 * block_index counts from 0.  out[i] are malloc'd here; free with orc_free. */
int orc_block_streams(const uint8_t *f, size_t n, size_t block_index, uint8_t *out[6], size_t len[6], uint32_t *nrec_out) {
    if (libs()) return ORC_E_NOLIB;
    if (n < 10 || !(f[0] == 'F' && f[1] == 'Q' && f[2] == 'Z' && f[3] == 0)) return ORC_E_MAGIC;
    int version = f[4];
    if (version != 1 && version != 2) return ORC_E_VERSION;
    size_t hsz = version == 1 ? 32 : 36, pos = 10;
    for (size_t b = 0;; b++) {
        if (pos + hsz > n) return ORC_E_TRUNC_FILE;
        const uint8_t *h = f + pos;
        uint32_t sz[6];
        if (version == 1) {
            sz[0] = rd_u32(h + 4); sz[1] = rd_u32(h + 8); sz[2] = rd_u32(h + 12); sz[3] = 0;
            sz[4] = rd_u32(h + 16); sz[5] = rd_u32(h + 20);
        } else
            for (int i = 0; i < 6; i++) sz[i] = rd_u32(h + 4 + 4 * i);
        pos += hsz;
        if (b == block_index) {
            *nrec_out = rd_u32(h);
            for (int i = 0; i < 6; i++) {
                buf_t s = {0};
                int rc = 0;
                if (!(version == 1 && i == 3)) {
                    if (pos + sz[i] > n) return ORC_E_TRUNC_FILE;
                    rc = zstd_decode_all(f + pos, sz[i], &s);
                    pos += sz[i];
                }
                if (rc) return rc;
                out[i] = s.p;
                len[i] = s.n;
            }
            return 0;
        }
        for (int i = 0; i < 6; i++) pos += sz[i];
    }
}
void orc_free(void *p) { free(p); }
