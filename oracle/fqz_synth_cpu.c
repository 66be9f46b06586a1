/*
 * fqz_synth_cpu.c — CPU twin of the synthetic FASTQ generator (fastqpacker_b200/csrc/fqz_synth.cu,
 * Python twin tests/synth.py).  TEST / BENCH INFRASTRUCTURE ONLY: bench.py --impl reference uses it
 * to build the CPU arm's input without touching the GPU library; tests check it against the twins.
 * Workload shapes: SURVEY.md §8(d) (kind 0 = BASELINE config 2, kind 1 = config 4).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct {
    uint64_t s, cur;
    int left;
} rng_t;
static uint64_t next64(rng_t *g) {
    uint64_t s = g->s;
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    g->s = s;
    return s;
}
static uint32_t next16(rng_t *g) {
    if (g->left == 0) {
        g->cur = next64(g);
        g->left = 4;
    }
    uint32_t v = (uint32_t)(g->cur & 0xFFFF);
    g->cur >>= 16;
    g->left--;
    return v;
}

/* kind 2: kind 0 in which 35 % of the records copy bases and qualities of one of the 400 records in front */
static uint64_t dup_source(uint64_t seed, uint64_t rec) {
    uint64_t body = rec;
    for (int hop = 0; hop < 64 && body > 0; hop++) {
        uint64_t h = mix(seed ^ 0xD0B1E5ull ^ mix(body));
        if ((h & 0xFFFF) >= 22938) break;
        uint64_t d = 1 + (h >> 16) % 400;
        body = body > d ? body - d : 0;
    }
    return body;
}

/* one record into p (>= 1024 bytes free); returns its length */
static size_t synth_record(int kind, uint64_t seed, uint64_t rec, uint8_t *p) {
    rng_t g;
    g.s = mix(seed ^ mix(rec));
    if (!g.s) g.s = 0x1234567;
    g.left = 0;
    g.cur = 0;
    char hdr[160];
    int hl;
    uint32_t L;
    if (kind != 1) {
        const uint64_t T = 250000;
        uint64_t t = rec / T;
        unsigned tile = (unsigned)((1 + (t / 48) % 2) * 1000 + (1 + (t / 16) % 3) * 100 + (1 + t % 16));
        unsigned x = (unsigned)(1000 + next64(&g) % 20000);
        unsigned y = (unsigned)(1000 + ((rec % T) * 2) / 5 + next16(&g) % 40);
        hl = snprintf(hdr, sizeof hdr, "ERR532393.%llu HWI-ST571:218:C2DACACXX:5:%u:%u:%u/1", (unsigned long long)(rec + 1), tile, x, y);
        L = 150;
    } else {
        L = (uint32_t)(50 + next64(&g) % 251);
        hl = snprintf(hdr, sizeof hdr, "SRR_synth.%llu %llu length=%u", (unsigned long long)(rec + 1), (unsigned long long)(rec + 1), L);
    }
    uint8_t *o = p;
    *o++ = '@';
    memcpy(o, hdr, (size_t)hl);
    o += hl;
    *o++ = '\n';
    static const char B[4] = {'A', 'C', 'G', 'T'};
    if (kind == 2) {
        uint64_t body = dup_source(seed, rec);
        if (body != rec) {
            g.s = mix(seed ^ mix(body));
            if (!g.s) g.s = 0x1234567;
            g.left = 0;
            g.cur = 0;
            next64(&g);
            next16(&g);
        }
    }
    if (kind != 1) {
        int nread = next16(&g) < 655;
        for (uint32_t i = 0; i < L; i++) {
            uint32_t d = next16(&g);
            uint8_t b = (uint8_t)B[d & 3];
            if (nread && (d >> 2) < 1638) b = 'N';
            *o++ = b;
        }
    } else {
        int inrun = 0;
        for (uint32_t i = 0; i < L; i++) {
            uint32_t d = next16(&g);
            uint8_t b = (uint8_t)B[d & 3];
            uint32_t r = d >> 2;
            inrun = inrun ? (r < 14746) : (r < 82);
            if (inrun) b = 'N';
            *o++ = b;
        }
    }
    *o++ = '\n';
    *o++ = '+';
    if (kind == 1) {
        memcpy(o, hdr, (size_t)hl);
        o += hl;
    }
    *o++ = '\n';
    int qmax = kind != 1 ? 41 : 40, qmin = kind != 1 ? 2 : 0, base = kind != 1 ? 33 : 64;
    int q = qmax - 7 + (int)(next16(&g) % 8);
    int tail = 0;
    for (uint32_t i = 0; i < L; i++) {
        if (i) {
            uint32_t d = next16(&g);
            if (tail) {
            } else if (d < 66)
                tail = 1;
            else if (d >= 49218) {
                uint32_t r = d - 49218;
                uint32_t m = (r >> 1) % 10;
                int mag = m < 6 ? 1 : (m < 9 ? 2 : 3);
                q += (r & 1) ? mag : -mag;
                if (q < qmin) q = qmin;
                if (q > qmax) q = qmax;
            }
        }
        *o++ = (uint8_t)(base + (tail ? 2 : q));
    }
    *o++ = '\n';
    return (size_t)(o - p);
}

/* records [first, first+count) -> out; returns 0, or -15 when cap is too small (needs ~1 KiB slack) */
int orc_synth(int kind, uint64_t seed, uint64_t first, uint64_t count, uint8_t *out, size_t cap, size_t *out_len) {
    size_t n = 0;
    for (uint64_t i = 0; i < count; i++) {
        if (n + 1024 > cap) return -15;
        n += synth_record(kind, seed, first + i, out + n);
    }
    *out_len = n;
    return 0;
}
