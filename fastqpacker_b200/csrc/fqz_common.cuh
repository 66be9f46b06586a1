// fqz_common.cuh — shared device helpers for the B200 (sm_100a) block codec.
//
// Everything here is byte/integer work bound by HBM bandwidth or by serial entropy-coding
// dependencies; there is no floating point and no tensor-core use anywhere in this library.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifndef FQZ_EMU
#include <cuda_runtime.h>
#define FQZ_LAUNCH(kernel, grid, block, smem, stream, ...) (++g_fqz_launches, kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#define FQZ_DYN_SMEM(type, name)                                      \
    extern __shared__ __align__(128) unsigned char fqz_dyn_smem_raw[]; \
    type *name = reinterpret_cast<type *>(fqz_dyn_smem_raw)
#endif

// kernels launched by this library on the calling thread (bench.py "gpu_launches"); a context is driven by one
// thread at a time, so per-thread counting keeps the contexts of a multi-GPU process apart
extern thread_local unsigned long long g_fqz_launches;

#define FQZ_BLOCK_RECORDS 100000u  // reference: internal/compress/compress.go:71 + batchPool :48-52
#define FQZ_MAX_SEQ_LEN 65536u     // reference: internal/encoder/sequence.go:11
#define FQZ_PAD 64                 // every device buffer is over-allocated by this many bytes

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

// error kinds, ordered as the reference meets them inside one record
enum { FQZ_K_HEADER_AT = 1, FQZ_K_PLUS = 2, FQZ_K_LEN = 3, FQZ_K_LONG_N = 4 };

// device-side status block of one compress window
struct FqzWinStatus {
    u64 err_key;   // min over (record << 8 | kind); ~0 = no error
    u32 qual_min;  // min quality byte over the file's first block (255 = none seen)
    u32 pad;
};

// ---------------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }

// inclusive scan inside a power-of-two sub-warp group of `width` lanes
__device__ __forceinline__ u32 group_incl_scan(u32 v, u32 mask, int width) {
    u32 l = threadIdx.x & (u32)(width - 1);
    for (int d = 1; d < width; d <<= 1) {
        u32 t = __shfl_up_sync(mask, v, (unsigned)d, width);
        if (l >= (u32)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u32 group_sum(u32 v, u32 mask, int width) {
    for (int d = width >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, width);
    return v;
}
__device__ __forceinline__ u32 group_min(u32 v, u32 mask, int width) {
    for (int d = width >> 1; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(mask, v, d, width));
    return v;
}
__device__ __forceinline__ u32 group_or(u32 v, u32 mask, int width) {
    for (int d = width >> 1; d > 0; d >>= 1) v |= __shfl_xor_sync(mask, v, d, width);
    return v;
}
// mask of the aligned `width`-lane group the calling lane belongs to
__device__ __forceinline__ u32 group_mask(int width) {
    u32 l = lane_id();
    u32 base = l & ~(u32)(width - 1);
    return (width == 32) ? 0xffffffffu : (((1u << width) - 1u) << base);
}

// CTA-wide exclusive scan of one u32 per thread (blockDim.x <= 1024, multiple of 32).
// `warp_sums` is a 33-entry shared array.  Returns exclusive prefix; *total = CTA sum.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *warp_sums, u32 *total) {
    u32 incl = group_incl_scan(v, 0xffffffffu, 32);
    u32 w = threadIdx.x >> 5, l = threadIdx.x & 31u, nw = (blockDim.x + 31u) >> 5;
    if (l == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        u32 s = (l < nw) ? warp_sums[l] : 0u;
        u32 si = group_incl_scan(s, 0xffffffffu, 32);
        warp_sums[l] = si - s;
        if (l == 31) warp_sums[32] = si;
    }
    __syncthreads();
    u32 r = warp_sums[w] + incl - v;
    if (total) *total = warp_sums[32];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------- byte helpers
// 32-bit little-endian load from an arbitrarily aligned address.  Touches the two aligned words
// that cover [p, p+4): every buffer carries FQZ_PAD bytes of slack so this never leaves it.
__device__ __forceinline__ u32 ld_u32_unaligned(const u8 *p) {
    uintptr_t a = (uintptr_t)p;
    const u32 *w = (const u32 *)(a & ~(uintptr_t)3);
    u32 sh = (u32)(a & 3) * 8u;
    u32 lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ u16 ld_u16_unaligned(const u8 *p) { return (u16)(p[0] | ((u32)p[1] << 8)); }
__device__ __forceinline__ void st_u16_unaligned(u8 *p, u32 v) {
    p[0] = (u8)v;
    p[1] = (u8)(v >> 8);
}
__device__ __forceinline__ void st_u32_unaligned(u8 *p, u32 v) {
    if (((uintptr_t)p & 3) == 0) {
        *(u32 *)p = v;
    } else {
        p[0] = (u8)v;
        p[1] = (u8)(v >> 8);
        p[2] = (u8)(v >> 16);
        p[3] = (u8)(v >> 24);
    }
}
// store the low `nbytes` (1..4) bytes of v
__device__ __forceinline__ void st_bytes(u8 *p, u32 v, u32 nbytes) {
    if (nbytes == 4) {
        st_u32_unaligned(p, v);
        return;
    }
    for (u32 i = 0; i < nbytes; i++) p[i] = (u8)(v >> (8 * i));
}

// 0xFF in every byte lane of w that is NOT one of ACGTacgt (reference: sequence.go:44-50 isNBase)
__device__ __forceinline__ u32 nonacgt_mask4(u32 w) {
    u32 up = w & 0xDFDFDFDFu;
    u32 ok = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) | __vcmpeq4(up, 0x47474747u) | __vcmpeq4(up, 0x54545454u);
    return ~ok;
}
// 2-bit codes of four bases (one per byte lane, in bits 0-1 of each lane); non-ACGT -> 0
// (reference: sequence.go:23-32 baseLookup).  A=0x41 C=0x43 G=0x47 T=0x54: t=(b>>1)&3 gives
// A0 C1 G3 T2, and t^(t>>1) maps that to A0 C1 G2 T3.
__device__ __forceinline__ u32 base_codes4(u32 w, u32 nmask) {
    u32 t = (w >> 1) & 0x03030303u;
    u32 c = t ^ ((t >> 1) & 0x01010101u);
    return c & ~nmask;
}
// pack the four 2-bit codes of base_codes4 into one byte (base i in bits 2i..2i+1)
__device__ __forceinline__ u32 pack_codes4(u32 c) { return (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xFFu; }
// byte mask: 0xFF for byte lanes k of the aligned word at `word_pos` with lo <= word_pos+k < hi
__device__ __forceinline__ u32 range_mask4(u32 word_pos, u32 lo, u32 hi) {
    u32 m = 0xFFFFFFFFu;
    if (word_pos < lo) {
        u32 s = lo - word_pos;
        m = (s >= 4) ? 0u : (m << (8 * s));
    }
    if (word_pos + 4 > hi) {
        u32 keep = (hi > word_pos) ? (hi - word_pos) : 0u;
        m = (keep == 0) ? 0u : (m & (0xFFFFFFFFu >> (8 * (4 - keep))));
    }
    return m;
}

// ---------------------------------------------------------------------------------- TMA (1-D bulk async copy) + mbarrier
// cp.async.bulk global->shared completes on an mbarrier; addresses and size are 16-byte multiples.
#if !defined(FQZ_EMU)
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
#else
// emulation: the barrier word holds the number of completed phases; copies are synchronous
__device__ __forceinline__ void mbar_init(u64 *bar, u32) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(u64 *, u32) {}
__device__ __forceinline__ void mbar_arrive(u64 *) {}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    while ((*bar & 1u) == parity) emu::yield();
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, u32 bytes, u64 *) {
    memcpy(smem_dst, gsrc, bytes);
}
#endif

// One elected thread stages [gsrc, gsrc+bytes) (16-byte aligned, 16-byte multiple) into shared
// memory through the TMA unit; the whole CTA then waits on the mbarrier phase.
// Splits into <=32 KiB bulk copies so that the mbarrier tx-count never overflows.
__device__ __forceinline__ void cta_stage_bulk(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar, u32 parity) {
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        u32 off = 0;
        while (off < bytes) {
            u32 c = min(bytes - off, 32768u);
            tma_load_1d((u8 *)smem_dst + off, (const u8 *)gsrc + off, c, bar);
            off += c;
        }
#ifdef FQZ_EMU
        *bar += 1;  // all (synchronous) copies landed: phase complete
#endif
    }
    mbar_wait(bar, parity);
}
