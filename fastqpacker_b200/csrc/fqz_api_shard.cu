// fqz_api_shard.cu — planning calls for ONE input split across several GPUs.
// The reference's producer cuts blocks every 100 000 records while it parses (compress.go:240-300);
// with the text spread over GPUs every rank counts the newlines of its slice, the counts are exchanged
// through the host (the one tiny exchange), and the byte ranges are re-cut on block boundaries
// (fastqpacker_b200/sharding.py).  These two calls are the device side of that plan.
#include <algorithm>

#include "fqz_host.h"

static const u64 kPlanChunk = (u64)2 << 30;  // tile counts are scanned as u32

// newlines of d[0..n) chunk by chunk; when `want` (1-based) falls into a chunk, also its offset
static int count_or_find(fqz_ctx *c, const u8 *d, u64 n, u64 want, u64 *lines, u64 *offset) {
    cudaStream_t s = c->stream;
    u64 total = 0;
    bool found = false;
    for (u64 pos = 0; pos < n && !found; pos += kPlanChunk) {
        c->arena.reset();
        const u64 len = std::min(kPlanChunk, n - pos);
        const u32 ntiles = (u32)((len + FQZ_NL_TILE - 1) / FQZ_NL_TILE);
        u32 *d_tiles = (u32 *)c->arena.alloc(((size_t)ntiles + 1) * sizeof(u32));
        u64 *d_pos = (u64 *)c->arena.alloc(sizeof(u64));
        if (!d_tiles || !d_pos) {
            c->err = "arena: out of device memory (line count)";
            return FQZ_E_CUDA;
        }
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_tiles + ntiles, 0, sizeof(u32), s));
        {
            StageScope sc(c, ST_NL_COUNT, len);
            fqz_launch_newline_count(d + pos, len, 0, d_tiles, ntiles, s);
        }
        {
            StageScope sc(c, ST_SCAN, 0);
            FQZ_TRY(fqz_scan_excl_u32(c, d_tiles, (u64)ntiles + 1, (u64)ntiles + 1, 1));
        }
        u32 *h = (u32 *)c->h_pin;
        FQZ_TRY(fqz_pin_copy(c, h, d_tiles + ntiles, sizeof(u32)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        const u64 here = h[0];
        if (want && want > total && want <= total + here) {
            fqz_launch_find_newline(d + pos, len, d_tiles, ntiles, (u32)(want - total - 1), d_pos, s);
            u64 *hp = (u64 *)(c->h_pin + 64);
            FQZ_TRY(fqz_pin_copy(c, hp, d_pos, sizeof(u64)));
            FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
            *offset = pos + hp[0];
            found = true;
        }
        total += here;
    }
    FQZ_CUDA_TRY(c, cudaGetLastError());
    if (lines) *lines = total;
    if (want && !found) return FQZ_E_INVALID_ARG;
    return FQZ_OK;
}

extern "C" int fqz_count_lines_device(fqz_ctx *c, const void *d_text, size_t n, uint64_t *lines) {
    if (!c || !lines || (!d_text && n) || ((uintptr_t)d_text & 15u)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    u64 t = 0;
    FQZ_TRY(count_or_find(c, (const u8 *)d_text, n, 0, &t, nullptr));
    *lines = t;
    return FQZ_OK;
}

extern "C" int fqz_find_line_end_device(fqz_ctx *c, const void *d_text, size_t n, uint64_t k, uint64_t *offset) {
    if (!c || !offset || !k || (!d_text && n) || ((uintptr_t)d_text & 15u)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    u64 off = 0;
    FQZ_TRY(count_or_find(c, (const u8 *)d_text, n, k, nullptr, &off));
    *offset = off;
    return FQZ_OK;
}
