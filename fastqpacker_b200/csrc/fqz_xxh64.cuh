// fqz_xxh64.cuh — XXH64 of many buffers at once (zstd frame content checksums, RFC 8878 §3.1.1:
// low 32 bits of XXH64 with seed 0; the reference's encoder writes it and its decoder verifies it,
// SURVEY.md F1).
//
// XXH64 is four serial multiply-rotate lanes per buffer and cannot be split inside a buffer, so it
// is parallelised ACROSS buffers: 4 threads (a quad) per buffer, 8 buffers per warp.  The serial
// chain is ~40 cycles per 32-byte stripe; what used to dominate was feeding it (every lane loading
// 8 bytes at a 32-byte stride).  Here each quad streams its buffer through shared memory with 1-D
// TMA bulk copies (cp.async.bulk, XX_STAGES chunks in flight per buffer, completion on one
// mbarrier per stage), so HBM sees large contiguous requests and the chain never waits on a load.
#pragma once
#include "fqz_common.cuh"
#include "fqz_zstd_tables.cuh"

#define XX_CH 1024u             // stripe bytes per chunk
#define XX_ROW (XX_CH + 32u)    // + 16 bytes of alignment spill, padded so that the 8 rows of a warp start 8 banks apart
#define XX_STAGES 3u
#define XX_WARPS 2u             // warps per CTA (8 buffers each)
#define XX_SMEM (XX_WARPS * 8u * XX_STAGES * (XX_ROW + 8u) + 128u)

__device__ __forceinline__ u64 xx_rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 xx_round(u64 acc, u64 in) { return xx_rotl64(acc + in * XXP2, 31) * XXP1; }
__device__ __forceinline__ u64 xx_merge(u64 h, u64 v) { return (h ^ xx_round(0, v)) * XXP1 + XXP4; }
__device__ __forceinline__ u64 xx_ld64(const u8 *p) { return (u64)ld_u32_unaligned(p) | ((u64)ld_u32_unaligned(p + 4) << 32); }

// XXH64(p[0..len)) computed by the quad of lanes q = 0..3 (gmask = their lane mask); every lane of
// the quad returns the hash.  MUST be called by all 32 lanes of the warp (eight buffers at once):
// control flow is kept warp-uniform — the eight quads step through their chunks in lockstep on
// shared per-stage mbarriers (8 arrivals each), shorter buffers idle by predication — because
// diverged quads would issue as eight separate instruction streams.
// rows: XX_STAGES staging rows of XX_ROW bytes (16-byte aligned) private to the quad;
// bars: XX_STAGES mbarriers private to the warp.  p may have any alignment; up to 15 bytes past
// p + len are read (every device buffer carries FQZ_PAD bytes of slack).
__device__ static u64 xxh64_quad_staged(const u8 *p, u64 len, u32 q, u32 gmask, u8 *rows, u64 *bars) {
    u64 acc = (q == 0) ? XXP1 + XXP2 : (q == 1) ? XXP2 : (q == 2) ? 0ull : 0ull - XXP1;
    const u64 nstripes = len >> 5, sbytes = nstripes << 5;
    const u32 sh = (u32)((uintptr_t)p & 15u);
    const u8 *ab = p - sh;
    const u64 span = (sh + sbytes + 15u) & ~15ull;
    const u32 nchunks = (u32)((sbytes + XX_CH - 1) / XX_CH);
    u32 maxchunks = nchunks;
    for (int d = 16; d >= 4; d >>= 1) maxchunks = max(maxchunks, __shfl_xor_sync(0xffffffffu, maxchunks, d));
    if (lane_id() == 0)
        for (u32 s = 0; s < XX_STAGES; s++) mbar_init(&bars[s], 8);
    __syncwarp();
    auto issue = [&](u32 k) {
        u32 s = k % XX_STAGES;
        if (q == 0 && k < maxchunks) {
            if (k < nchunks) {
                u64 off = (u64)k * XX_CH;
                u32 bytes = (u32)min((u64)(XX_CH + 16u), span - off);
                mbar_expect_tx(&bars[s], bytes);
                tma_load_1d(rows + s * XX_ROW, ab + off, bytes, &bars[s]);
            } else
                mbar_arrive(&bars[s]);
        }
#ifdef FQZ_EMU
        __syncwarp();
        if (lane_id() == 0 && k < maxchunks) bars[s] += 1;
#endif
    };
    for (u32 k = 0; k < XX_STAGES; k++) issue(k);
    const u32 o0 = sh + 8u * q;  // this lane's 8 bytes of stripe i start at row byte o0 + 32 i
    const u32 wsh = (o0 & 3u) * 8u;
    for (u32 k = 0; k < maxchunks; k++) {
        u32 s = k % XX_STAGES;
        mbar_wait(&bars[s], (k / XX_STAGES) & 1u);
        const u32 *w = (const u32 *)(rows + s * XX_ROW + (o0 & ~3u));
        u32 cs = (k < nchunks) ? (u32)(min((u64)XX_CH, sbytes - (u64)k * XX_CH) >> 5) : 0u;
        if (__all_sync(0xffffffffu, cs == XX_CH / 32u)) {  // the common case: eight full chunks
#pragma unroll 1
            for (u32 i = 0; i < XX_CH / 32u; i += 8) {
                u32 a[8], b[8], c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a[j] = w[8 * (i + j)];
                    b[j] = w[8 * (i + j) + 1];
                    c[j] = w[8 * (i + j) + 2];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    u64 v = (u64)__funnelshift_r(a[j], b[j], wsh) | ((u64)__funnelshift_r(b[j], c[j], wsh) << 32);
                    acc = xx_round(acc, v);
                }
            }
        } else {
            for (u32 i = 0; i < XX_CH / 32u; i++) {
                if (i < cs) {
                    u32 a = w[8 * i], b = w[8 * i + 1], c = w[8 * i + 2];
                    u64 v = (u64)__funnelshift_r(a, b, wsh) | ((u64)__funnelshift_r(b, c, wsh) << 32);
                    acc = xx_round(acc, v);
                }
            }
        }
        __syncwarp();  // every lane is done with this stage before it is refilled
        issue(k + XX_STAGES);
    }
    u64 a0 = __shfl_sync(gmask, acc, 0, 4), a1 = __shfl_sync(gmask, acc, 1, 4), a2 = __shfl_sync(gmask, acc, 2, 4), a3 = __shfl_sync(gmask, acc, 3, 4);
    u64 h;
    if (len >= 32) {
        h = xx_rotl64(a0, 1) + xx_rotl64(a1, 7) + xx_rotl64(a2, 12) + xx_rotl64(a3, 18);
        h = xx_merge(h, a0);
        h = xx_merge(h, a1);
        h = xx_merge(h, a2);
        h = xx_merge(h, a3);
    } else
        h = XXP5;
    h += len;
    const u8 *t = p + sbytes;  // < 32 bytes of tail, straight from global memory
    u32 rem = (u32)(len & 31);
    while (rem >= 8) {
        h ^= xx_round(0, xx_ld64(t));
        h = xx_rotl64(h, 27) * XXP1 + XXP4;
        t += 8;
        rem -= 8;
    }
    if (rem >= 4) {
        h ^= (u64)ld_u32_unaligned(t) * XXP1;
        h = xx_rotl64(h, 23) * XXP2 + XXP3;
        t += 4;
        rem -= 4;
    }
    while (rem) {
        h ^= (u64)(*t) * XXP5;
        h = xx_rotl64(h, 11) * XXP1;
        t++;
        rem--;
    }
    h ^= h >> 33;
    h *= XXP2;
    h ^= h >> 29;
    h *= XXP3;
    h ^= h >> 32;
    return h;
}

// carve the calling quad's staging rows and barriers out of the CTA's dynamic shared memory
__device__ __forceinline__ void xx_quad_smem(u8 *smem, u8 **rows, u64 **bars) {
    u32 quad = threadIdx.x >> 2;  // 0 .. XX_WARPS*8-1
    u8 *base = (u8 *)(((uintptr_t)smem + 127u) & ~(uintptr_t)127u);
    *rows = base + (size_t)quad * XX_STAGES * XX_ROW;
    *bars = (u64 *)(base + (size_t)XX_WARPS * 8u * XX_STAGES * XX_ROW) + (threadIdx.x >> 5) * XX_STAGES;
}
