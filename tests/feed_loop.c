/* feed_loop.c — TEST PROGRAM: the exact producer / feed / writer loop that INTEGRATION.md's cgo shim
 * (internal/compress/compress_cuda.go) runs, written in C against include/fqzgpu.h, so that the loop the Go
 * host would execute is compiled and tested even though this image has no Go toolchain.
 *
 *   feed_loop c <in.fastq> <out.fqz> [window_bytes]     compress.Compress   (compress.go:125-192)
 *   feed_loop d <in.fqz> <out.fastq> [window_bytes]     compress.Decompress (compress.go:558-604)
 *
 * Window buffers come from fqz_host_alloc (page-locked).  Prints "ok <bytes in> <bytes out> <feed calls> <seconds>".
 * Exit code: 0, or 1 with the library's error text on stderr. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fqzgpu.h"

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static int fail(fqz_ctx *ctx, int rc, const char *wrap) {
    const char *d = ctx ? fqz_last_error(ctx) : "";
    fprintf(stderr, "%s%s%s%s%s%s\n", wrap, wrap[0] ? ": " : "", fqz_strerror(rc), d && d[0] ? " [" : "", d && d[0] ? d : "", d && d[0] ? "]" : "");
    return 1;
}

static int do_compress(fqz_ctx *ctx, FILE *r, FILE *w, size_t window, size_t *nin, size_t *nout, unsigned *calls) {
    fqz_cstream *s = NULL;
    int rc = fqz_compress_begin(ctx, 0, &s);
    if (rc != FQZ_OK) return fail(ctx, rc, "");
    size_t cap = window + (64u << 20), len = 0;
    uint8_t *in = (uint8_t *)fqz_host_alloc(cap);
    size_t ocap = fqz_compress_bound(cap);
    uint8_t *out = (uint8_t *)fqz_host_alloc(ocap);
    if (!in || !out) return fail(ctx, FQZ_E_CUDA, "allocating window buffers");
    int eof = 0, ret = 0;
    while (!eof || len > 0) {
        /* top the window up from the reader (the unconsumed tail of the last call stays in front) */
        while (len < window && !eof) {
            size_t n = fread(in + len, 1, cap - len, r);
            len += n;
            *nin += n;
            if (n == 0) eof = 1;
        }
        size_t out_len = 0, used = 0;
        rc = fqz_compress_feed(s, len ? in : NULL, len, eof, out, ocap, &out_len, &used);
        (*calls)++;
        if (rc == FQZ_E_NEED_MORE) { /* no complete 100 000-record block yet: read more */
            if (cap - len < window / 2) {
                size_t ncap = 2 * cap;
                uint8_t *nb = (uint8_t *)fqz_host_alloc(ncap);
                if (!nb) { ret = fail(ctx, FQZ_E_CUDA, "growing the window"); break; }
                memcpy(nb, in, len);
                fqz_host_free(in);
                fqz_host_free(out);
                in = nb;
                cap = ncap;
                ocap = fqz_compress_bound(cap);
                out = (uint8_t *)fqz_host_alloc(ocap);
                if (!out) { ret = fail(ctx, FQZ_E_CUDA, "growing the window"); break; }
            }
            window = cap - (64u << 20);
            continue;
        }
        if (rc != FQZ_OK) { ret = fail(ctx, rc, "parsing FASTQ"); break; }
        if (out_len && fwrite(out, 1, out_len, w) != out_len) { ret = 1; break; }
        *nout += out_len;
        memmove(in, in + used, len - used);
        len -= used;
        if (eof) break;
    }
    fqz_compress_end(s);
    fqz_host_free(in);
    fqz_host_free(out);
    return ret;
}

static int do_decompress(fqz_ctx *ctx, FILE *r, FILE *w, size_t window, size_t room, size_t *nin, size_t *nout, unsigned *calls) {
    fqz_dstream *s = NULL;
    int rc = fqz_decompress_begin(ctx, &s);
    if (rc != FQZ_OK) return fail(ctx, rc, "");
    size_t cap = window, len = 0, ocap = room ? room : 6 * window + (1u << 20);  /* room: output bytes to start with (tests: too few) */
    uint8_t *in = (uint8_t *)fqz_host_alloc(cap), *out = (uint8_t *)fqz_host_alloc(ocap);
    if (!in || !out) return fail(ctx, FQZ_E_CUDA, "allocating window buffers");
    int eof = 0, ret = 0;
    for (;;) {
        while (len < cap && !eof) {
            size_t n = fread(in + len, 1, cap - len, r);
            len += n;
            *nin += n;
            if (n == 0) eof = 1;
        }
        size_t out_len = 0, used = 0;
        rc = fqz_decompress_feed(s, len ? in : NULL, len, eof, out, ocap, &out_len, &used);
        (*calls)++;
        if (rc == FQZ_E_NEED_MORE) {
            if (eof) { fprintf(stderr, "reading block header: unexpected EOF\n"); ret = 1; break; }
            uint8_t *nb = (uint8_t *)fqz_host_alloc(2 * cap);
            if (!nb) { ret = fail(ctx, FQZ_E_CUDA, "growing the window"); break; }
            memcpy(nb, in, len);
            fqz_host_free(in);
            in = nb;
            cap *= 2;
            continue;
        }
        if (rc == FQZ_E_NOSPACE) {
            fqz_host_free(out);
            ocap = 2 * (out_len > ocap ? out_len : ocap);
            out = (uint8_t *)fqz_host_alloc(ocap);
            if (!out) { ret = fail(ctx, FQZ_E_CUDA, "growing the output"); break; }
            continue;
        }
        if (rc != FQZ_OK) { ret = fail(ctx, rc, ""); break; }
        if (out_len && fwrite(out, 1, out_len, w) != out_len) { ret = 1; break; }
        *nout += out_len;
        memmove(in, in + used, len - used);
        len -= used;
        if (eof && len == 0) break;
        if (eof && used == 0 && out_len == 0) { fprintf(stderr, "reading block: unexpected EOF\n"); ret = 1; break; }
    }
    fqz_decompress_end(s);
    fqz_host_free(in);
    fqz_host_free(out);
    return ret;
}

int main(int argc, char **argv) {
    if (argc < 4 || (argv[1][0] != 'c' && argv[1][0] != 'd')) {
        fprintf(stderr, "usage: feed_loop c|d <in> <out> [window_bytes [output_bytes]]\n");
        return 2;
    }
    size_t window = argc > 4 ? (size_t)strtoull(argv[4], NULL, 10) : ((size_t)512 << 20);
    size_t room = argc > 5 ? (size_t)strtoull(argv[5], NULL, 10) : 0;
    FILE *r = fopen(argv[2], "rb"), *w = fopen(argv[3], "wb");
    if (!r || !w) {
        perror("open");
        return 2;
    }
    fqz_ctx *ctx = NULL;
    int rc = fqz_init(0, &ctx);
    if (rc != FQZ_OK) return fail(NULL, rc, "");
    size_t nin = 0, nout = 0;
    unsigned calls = 0;
    double t0 = now();
    int ret = argv[1][0] == 'c' ? do_compress(ctx, r, w, window, &nin, &nout, &calls) : do_decompress(ctx, r, w, window, room, &nin, &nout, &calls);
    double dt = now() - t0;
    fqz_destroy(ctx);
    fclose(r);
    fclose(w);
    if (ret == 0) printf("ok %zu %zu %u %.6f\n", nin, nout, calls, dt);
    return ret;
}
