// fqz_synth.cu — synthetic FASTQ generator (bench / test utility, never on the codec path).
//
// Counter-based: record i depends only on (seed, i), so any device — and the pure-Python twin in
// tests/synth.py — produces identical bytes, and a 64 GB data set can be generated in HBM shard by
// shard.  Integer arithmetic only.
//   kind 0: "ERR532393_1-shaped" Illumina HiSeq reads: 150 bp, Phred+33 (Q2..Q41), bare '+',
//           N in ~1 % of reads (10 % of their bases), first-order Markov qualities.
//   kind 1: read length 50..300, Phred+64, ~5 % N in clustered runs, '+' line repeats the header.
//   kind 2: kind 0 with duplicates: 35 % of the records copy the bases and qualities of one of the 400 records
//           in front of them (PCR / optical duplicates of a clustered run); read names stay unique.
#include "fqz_host.h"

__host__ __device__ static inline u64 synth_mix(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct SynthRng {
    u64 s;
    u64 cur;
    int left;
    __host__ __device__ void init(u64 seed, u64 rec) {
        s = synth_mix(seed ^ synth_mix(rec));
        if (s == 0) s = 0x1234567ull;
        left = 0;
        cur = 0;
    }
    __host__ __device__ u64 next64() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return s;
    }
    __host__ __device__ u32 next16() {  // four 16-bit draws per 64-bit step
        if (left == 0) {
            cur = next64();
            left = 4;
        }
        u32 v = (u32)(cur & 0xFFFFu);
        cur >>= 16;
        left--;
        return v;
    }
};

__host__ __device__ static inline int put_dec(u8 *p, u64 v) {  // returns digits written (p may be null: count only)
    char tmp[20];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + (v % 10));
        v /= 10;
    } while (v);
    if (p)
        for (int i = 0; i < n; i++) p[i] = (u8)tmp[n - 1 - i];
    return n;
}
__host__ __device__ static inline int put_str(u8 *p, const char *s) {
    int n = 0;
    while (s[n]) {
        if (p) p[n] = (u8)s[n];
        n++;
    }
    return n;
}

// kind 2: the record whose bases and qualities record `rec` carries (a copy may copy a copy)
__host__ __device__ static inline u64 synth_dup_source(u64 seed, u64 rec) {
    u64 body = rec;
    for (int hop = 0; hop < 64 && body > 0; hop++) {
        u64 h = synth_mix(seed ^ 0xD0B1E5ull ^ synth_mix(body));
        if ((h & 0xFFFFu) >= 22938u) break;  // 35 % copy
        u64 d = 1 + (h >> 16) % 400;
        body = body > d ? body - d : 0;
    }
    return body;
}

// Writes record `rec` at p (or only measures it when p == nullptr).  Returns its byte length.
__host__ __device__ static u32 synth_record(int kind, u64 seed, u64 rec, u8 *p) {
    SynthRng g;
    g.init(seed, rec);
    u32 n = 0;
#define EMIT(c)                    \
    do {                           \
        if (p) p[n] = (u8)(c);     \
        n++;                       \
    } while (0)
#define EMITF(call)                          \
    do {                                     \
        n += (u32)(call);                    \
    } while (0)
    u32 L;
    u32 hdr_start, hdr_len;
    if (kind != 1) {
        const u64 T = 250000;
        u64 t = rec / T;
        u32 tile = (u32)((1 + (t / 48) % 2) * 1000 + (1 + (t / 16) % 3) * 100 + (1 + t % 16));
        u32 x = 1000 + (u32)(g.next64() % 20000);
        u32 y = 1000 + (u32)(((rec % T) * 2) / 5) + (u32)(g.next16() % 40);
        EMIT('@');
        hdr_start = n;
        EMITF(put_str(p ? p + n : nullptr, "ERR532393."));
        EMITF(put_dec(p ? p + n : nullptr, rec + 1));
        EMITF(put_str(p ? p + n : nullptr, " HWI-ST571:218:C2DACACXX:5:"));
        EMITF(put_dec(p ? p + n : nullptr, tile));
        EMIT(':');
        EMITF(put_dec(p ? p + n : nullptr, x));
        EMIT(':');
        EMITF(put_dec(p ? p + n : nullptr, y));
        EMITF(put_str(p ? p + n : nullptr, "/1"));
        hdr_len = n - hdr_start;
        EMIT('\n');
        L = 150;
    } else {
        L = 50 + (u32)(g.next64() % 251);
        EMIT('@');
        hdr_start = n;
        EMITF(put_str(p ? p + n : nullptr, "SRR_synth."));
        EMITF(put_dec(p ? p + n : nullptr, rec + 1));
        EMIT(' ');
        EMITF(put_dec(p ? p + n : nullptr, rec + 1));
        EMITF(put_str(p ? p + n : nullptr, " length="));
        EMITF(put_dec(p ? p + n : nullptr, L));
        hdr_len = n - hdr_start;
        EMIT('\n');
    }
    if (!p) {  // length only: seq + '\n' + plus line + '\n' + qual + '\n'
        n += L + 1 + 1 + (kind == 1 ? hdr_len : 0) + 1 + L + 1;
        return n;
    }
    if (kind == 2) {
        u64 body = synth_dup_source(seed, rec);
        if (body != rec) {  // same draws as the source record made for its own header
            g.init(seed, body);
            g.next64();
            g.next16();
        }
    }
    // ---- sequence
    if (kind != 1) {
        bool nread = g.next16() < 655;  // ~1 % of reads carry Ns
        for (u32 i = 0; i < L; i++) {
            u32 d = g.next16();
            u8 b = (u8)("ACGT"[d & 3]);
            if (nread && (d >> 2) < 1638) b = 'N';  // 10 % of their bases
            EMIT(b);
        }
    } else {
        bool inrun = false;
        for (u32 i = 0; i < L; i++) {
            u32 d = g.next16();
            u8 b = (u8)("ACGT"[d & 3]);
            u32 r = d >> 2;  // 14 bits
            if (inrun) inrun = r < 14746;       // continue with p = 0.9
            else inrun = r < 82;                // start with p = 0.005
            if (inrun) b = 'N';
            EMIT(b);
        }
    }
    EMIT('\n');
    // ---- plus line
    EMIT('+');
    if (kind == 1)
        for (u32 i = 0; i < hdr_len; i++) EMIT(p[hdr_start + i]);
    EMIT('\n');
    // ---- qualities: first-order Markov chain per read
    {
        int qmax = (kind != 1) ? 41 : 40, qmin = (kind != 1) ? 2 : 0, base = (kind != 1) ? 33 : 64;
        int q = qmax - 7 + (int)(g.next16() % 8);
        bool tail = false;
        for (u32 i = 0; i < L; i++) {
            if (i) {
                u32 d = g.next16();
                if (tail) {
                } else if (d < 66) {
                    tail = true;  // Q2 tail, absorbing (~0.1 % per base)
                } else if (d >= 49218) {  // ~25 %: move by 1..3
                    u32 r = d - 49218;
                    int m = (int)((r >> 1) % 10);
                    int mag = m < 6 ? 1 : (m < 9 ? 2 : 3);
                    q += (r & 1) ? mag : -mag;
                    if (q > qmax) q = qmax;
                    if (q < qmin) q = qmin;
                }
            }
            EMIT(base + (tail ? 2 : q));
        }
    }
    EMIT('\n');
#undef EMIT
#undef EMITF
    return n;
}

__global__ void k_synth_len(int kind, u64 seed, u64 first, u64 count, u32 *lens) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) lens[i] = synth_record(kind, seed, first + i, nullptr);
}
__global__ void k_synth_write(int kind, u64 seed, u64 first, u64 count, const u32 *offs, u8 *out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) synth_record(kind, seed, first + i, out + offs[i]);
}

extern "C" int fqz_synth_device(fqz_ctx *c, int kind, uint64_t seed, uint64_t first, uint64_t count, void *d_out, size_t out_cap,
                                size_t *out_len) {
    if (!c || !out_len || kind < 0 || kind > 2) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->arena.reset();
    *out_len = 0;
    if (count == 0) return FQZ_OK;
    if (count > 9000000ull) return FQZ_E_TOO_LARGE;  // keeps u32 offsets far below 4 GiB
    u32 *d_lens = (u32 *)c->arena.alloc((size_t)(count + 1) * sizeof(u32));
    if (!d_lens) return FQZ_E_CUDA;
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_lens + count, 0, sizeof(u32), c->stream));
    u32 grid = (u32)((count + 127) / 128);
    FQZ_LAUNCH(k_synth_len, grid, 128, 0, c->stream, kind, seed, first, count, d_lens);
    FQZ_TRY(fqz_scan_excl_u32(c, d_lens, count + 1, count + 1, 1));
    u32 *h = (u32 *)c->h_pin;
    FQZ_CUDA_TRY(c, cudaMemcpyAsync(h, d_lens + count, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out_len = h[0];
    if (h[0] > out_cap) return FQZ_E_NOSPACE;
    FQZ_LAUNCH(k_synth_write, grid, 128, 0, c->stream, kind, seed, first, count, d_lens, (u8 *)d_out);
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    return FQZ_OK;
}
