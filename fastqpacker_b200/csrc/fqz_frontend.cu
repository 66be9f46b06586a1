// fqz_frontend.cu — compress front end: FASTQ text -> six pre-entropy streams, bit-exact with
// the reference's fqparser + encoder + per-record stream assembly.
//
//   k_newline_count / k_newline_index   warp-ballot-free SIMD-in-register '\n' scan -> line_end[]
//                                       (replaces Parser.readLine, internal/fqparser/parser.go:209-243)
//   k_record_meta                       validation + per-record stream sizes + Phred min
//                                       (parser.go:136-184 nextInto, encoder/quality.go:22-49,
//                                        compress.go:477-488 long-read guard)
//   k_scan_*                            device-wide exclusive scans (stream offsets per record)
//   k_scatter_streams                   TMA-staged record chunks -> seqPacked / quality / headers /
//                                       plusLines / nPositions / seqLengths
//                                       (compress.go:490-519, sequence.go:139-184, quality.go:53-103)
//
// All kernels are HBM-bound byte shuffling: algorithmic bytes = F (text) + S (streams).
#include "fqz_common.cuh"
#include "fqz_kernels.h"

#define NL4 0x0a0a0a0au

// ---------------------------------------------------------------------------------- newline scan
// One CTA = one tile of FQZ_NL_TILE bytes; each thread owns 64 contiguous bytes (4 x 128-bit loads).
// lo: newlines at window offsets below lo do not count (a window that starts inside the 16-byte word holding
// the tail of the previous window's last record may see several short lines there; only the '\n' at
// skip - 1 is line -1)
__device__ __forceinline__ u32 nl_mask64(const u8 *text, u64 n, u64 pos, u64 lo, u64 &m_out) {
    u64 m = 0;  // bit k set <=> text[pos+k] == '\n'
    if (pos < n) {
        const uint4 *v = (const uint4 *)(text + pos);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            u64 p = pos + 16u * q;
            if (p >= n) break;
            uint4 x = __ldg(v + q);
            u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 e = __vcmpeq4(w[k], NL4) & 0x01010101u;
                u32 bits = (e * 0x01020408u) >> 24;  // 4 flags -> 4 bits
                m |= (u64)(bits & 0xFu) << (16 * q + 4 * k);
            }
        }
        if (pos + 64 > n) m &= (~0ull) >> (64 - (n - pos));  // bytes past the end do not count
        if (pos < lo) m = (lo - pos >= 64) ? 0ull : (m & ((~0ull) << (lo - pos)));
    }
    m_out = m;
    return (u32)__popcll(m);
}

__global__ void __launch_bounds__(FQZ_NL_THREADS) k_newline_count(const u8 *text, u64 n, u64 lo, u32 *tile_counts) {
    __shared__ u32 ws[33];
    u64 pos = (u64)blockIdx.x * FQZ_NL_TILE + (u64)threadIdx.x * 64u;
    u64 m;
    u32 c = nl_mask64(text, n, pos, lo, m);
    u32 total;
    block_excl_scan(c, ws, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// Offset of the newline number `target` (0-based) of the text: one CTA finds the tile in the scanned tile
// counts and rescans that tile (planning calls of the multi-GPU split, fqz_find_line_end_device).
__global__ void __launch_bounds__(FQZ_NL_THREADS) k_find_newline(const u8 *text, u64 n, const u32 *tile_prefix, u32 ntiles, u32 target, u64 *out_pos) {
    __shared__ u32 ws[33];
    __shared__ u32 s_tile;
    if (threadIdx.x == 0) {
        u32 lo = 0, hi = ntiles - 1;  // largest t with tile_prefix[t] <= target
        while (lo < hi) {
            u32 mid = (lo + hi + 1) >> 1;
            if (tile_prefix[mid] <= target) lo = mid; else hi = mid - 1;
        }
        s_tile = lo;
    }
    __syncthreads();
    const u32 t = s_tile;
    u64 pos = (u64)t * FQZ_NL_TILE + (u64)threadIdx.x * 64u;
    u64 m;
    u32 c = nl_mask64(text, n, pos, 0, m);
    u32 ex = block_excl_scan(c, ws, nullptr);
    u32 want = target - tile_prefix[t];
    if (want >= ex && want < ex + c) {
        for (u32 k = ex; k < want; k++) m &= m - 1;
        *out_pos = pos + (u64)(__ffsll((long long)m) - 1);
    }
}

// tile_prefix = exclusive scan of tile_counts.  line_end[i] = byte offset of the i-th '\n'.
__global__ void __launch_bounds__(FQZ_NL_THREADS) k_newline_index(const u8 *text, u64 n, u64 lo, const u32 *tile_prefix, u32 *line_end, u32 max_lines) {
    __shared__ u32 ws[33];
    u64 pos = (u64)blockIdx.x * FQZ_NL_TILE + (u64)threadIdx.x * 64u;
    u64 m;
    u32 c = nl_mask64(text, n, pos, lo, m);
    u32 ex = block_excl_scan(c, ws, nullptr);
    u32 idx = tile_prefix[blockIdx.x] + ex;
    while (m) {
        int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (idx < max_lines) line_end[idx] = (u32)(pos + (u64)b);
        idx++;
    }
}

// Count and index in ONE pass over the text: a tile counts its newlines, gets the number of newlines in front of it by
// a decoupled look-back over the tiles before it (tile numbers come from an atomic ticket, so a tile only waits for
// tiles that are already running), and writes its line ends straight from the masks it still holds in registers.
// look: one zeroed u64 per tile (+1: the ticket).  line_end has room for cap_lines entries; *total_lines = all newlines.
#define NS_FLAG_AGG (1ull << 62)
#define NS_FLAG_INC (2ull << 62)
__global__ void __launch_bounds__(FQZ_NL_THREADS) k_newline_scan(const u8 *text, u64 n, u64 lo, u32 *line_end, u32 cap_lines, unsigned long long *look,
                                                                 u32 ntiles, u32 *total_lines) {
    __shared__ u32 ws[33];
    __shared__ u32 s_tile, s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd((u32 *)(look + ntiles), 1u);
    __syncthreads();
    const u32 tile = s_tile;
    u64 pos = (u64)tile * FQZ_NL_TILE + (u64)threadIdx.x * 64u;
    u64 m;
    u32 c = nl_mask64(text, n, pos, lo, m);
    u32 total;
    u32 ex = block_excl_scan(c, ws, &total);
    if (threadIdx.x == 0) {
        volatile unsigned long long *L = look;
        unsigned long long base = 0;
        if (tile == 0) L[0] = NS_FLAG_INC | total;
        else {
            L[tile] = NS_FLAG_AGG | total;
            __threadfence();
            long long p = (long long)tile - 1;
            for (;;) {
                unsigned long long v = L[p];
                if ((v >> 62) == 0) {
#ifdef FQZ_EMU
                    emu::yield();
#endif
                    continue;
                }
                base += v & ((1ull << 62) - 1ull);
                if ((v >> 62) == 2) break;
                p--;
            }
            __threadfence();
            L[tile] = NS_FLAG_INC | (base + total);
        }
        s_base = (u32)base;
        if (tile == ntiles - 1) *total_lines = (u32)(base + total);
    }
    __syncthreads();
    u32 idx = s_base + ex;
    while (m) {
        int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (idx < cap_lines) line_end[idx] = (u32)(pos + (u64)b);
        idx++;
    }
}

// ---------------------------------------------------------------------------------- device-wide exclusive scan (u32, in place)
// data holds `narr` arrays of n elements, array a at data + a*stride.  2048 elements per CTA.
__device__ __forceinline__ u32 scan_tile_load(const u32 *d, u64 n, u64 base, u32 v[FQZ_SCAN_PER_THREAD]) {
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < FQZ_SCAN_PER_THREAD; i++) {
        u64 k = base + (u64)threadIdx.x * FQZ_SCAN_PER_THREAD + i;
        v[i] = (k < n) ? d[k] : 0u;
        s += v[i];
    }
    return s;
}
__global__ void __launch_bounds__(FQZ_SCAN_THREADS) k_scan_partial(const u32 *data, u64 n, u64 stride, u32 *sums, u32 ntiles) {
    __shared__ u32 ws[33];
    const u32 *d = data + (u64)blockIdx.y * stride;
    u32 v[FQZ_SCAN_PER_THREAD];
    u32 s = scan_tile_load(d, n, (u64)blockIdx.x * FQZ_SCAN_TILE, v);
    u32 total;
    block_excl_scan(s, ws, &total);
    if (threadIdx.x == 0) sums[(u64)blockIdx.y * ntiles + blockIdx.x] = total;
}
// sums == nullptr: single-tile scan (tile offset 0)
__global__ void __launch_bounds__(FQZ_SCAN_THREADS) k_scan_apply(u32 *data, u64 n, u64 stride, const u32 *sums, u32 ntiles) {
    __shared__ u32 ws[33];
    u32 *d = data + (u64)blockIdx.y * stride;
    u32 v[FQZ_SCAN_PER_THREAD];
    u64 base = (u64)blockIdx.x * FQZ_SCAN_TILE;
    u32 s = scan_tile_load(d, n, base, v);
    u32 ex = block_excl_scan(s, ws, nullptr);
    u32 run = ex + (sums ? sums[(u64)blockIdx.y * ntiles + blockIdx.x] : 0u);
#pragma unroll
    for (int i = 0; i < FQZ_SCAN_PER_THREAD; i++) {
        u64 k = base + (u64)threadIdx.x * FQZ_SCAN_PER_THREAD + i;
        if (k < n) d[k] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------------- per-record metadata
struct LineSpan {
    u32 hs, he, ss, se, ps, pe, qs, qe;  // [start, end) of the four lines, '\n' and one '\r' stripped
};
// reference: parser.go:209-220 (strip '\n', then one trailing '\r')
__device__ __forceinline__ u32 strip_cr(const u8 *text, u32 s, u32 e) { return (e > s && text[e - 1] == '\r') ? e - 1 : e; }
__device__ __forceinline__ LineSpan record_lines(const u8 *text, const u32 *line_end, u64 r) {
    LineSpan t;
    u64 l = 4 * r;
    u32 e0 = line_end[(long long)l - 1];  // entry -1 exists: 0xFFFFFFFF, or the end of the previous window's last line
    u32 e1 = line_end[l], e2 = line_end[l + 1], e3 = line_end[l + 2], e4 = line_end[l + 3];
    t.hs = e0 + 1u;
    t.he = strip_cr(text, t.hs, e1);
    t.ss = e1 + 1u;
    t.se = strip_cr(text, t.ss, e2);
    t.ps = e2 + 1u;
    t.pe = strip_cr(text, t.ps, e3);
    t.qs = e3 + 1u;
    t.qe = strip_cr(text, t.qs, e4);
    return t;
}

// Count bytes that are not ACGTacgt in text[a,b) with a group of `W` lanes (lane index g) using
// aligned 32-bit words.  Returns the lane's partial count.
// A byte is one of ACGTacgt exactly when, with bit 5 cleared, it equals the letter that its bits 1-2
// select from "ACTG" (the same two bits the packer uses as the base code): one byte_perm lookup and one
// XOR test four bytes.  Returns a word that is non-zero in exactly the byte lanes that are NOT ACGTacgt.
__device__ __forceinline__ u32 nonacgt_diff4(u32 x) {
    u32 t = (x >> 1) & 0x03030303u;
    u32 u2 = (t | (t >> 4)) & 0x00330033u;  // two selector nibbles per half word
    u32 sel = (u2 | (u2 >> 8)) & 0x3333u;   // four selector nibbles
    return (x & 0xDFDFDFDFu) ^ __byte_perm(0x47544341u, 0u, sel);
}
// text must be 16-byte aligned (window bases are): every lane of the group reads whole 16-byte vectors, so a
// 150-base line is two load round trips for the group instead of five.
template <int W>
__device__ __forceinline__ u32 partial_count_n(const u8 *text, u32 a, u32 b, u32 g) {
    u32 c = 0;
    if (b > a) {
        u32 v0 = a >> 4, v1 = (b - 1) >> 4;
        for (u32 v = v0 + g; v <= v1; v += W) {
            uint4 q = *(const uint4 *)(text + 16ull * v);
            u32 xs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (u32 k = 0; k < 4; k++) {
                u32 wp = 16u * v + 4u * k;
                u32 y = nonacgt_diff4(xs[k]);
                if (wp < a || wp + 4u > b) y &= range_mask4(wp, a, b);  // words at or beyond the ends of the line
                if (y) c += (u32)__popc((((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u);
            }
        }
    }
    return c;
}
template <int W>
__device__ __forceinline__ u32 partial_min_byte(const u8 *text, u32 a, u32 b, u32 g) {
    u32 mn = 0xFFFFFFFFu;
    if (b > a) {
        u32 w0 = a >> 2, w1 = (b - 1) >> 2;
        for (u32 w = w0 + g; w <= w1; w += W) {
            u32 x = *(const u32 *)(text + 4ull * w);
            x |= ~range_mask4(4u * w, a, b);
            mn = __vminu4(mn, x);
        }
    }
    u32 m2 = min(mn & 0xFFu, (mn >> 8) & 0xFFu);
    u32 m3 = min((mn >> 16) & 0xFFu, mn >> 24);
    return min(m2, m3);
}

// sizes: 5 arrays of (R+1) u32 at stride `stride`: seq, qual, hdr, plus, npos bytes per record.
// rec_base = index of this window's first record in the whole file (for error keys / Phred scope).
// tail_lines: complete lines after the last whole record (only checked when is_last).
__global__ void __launch_bounds__(FQZ_META_THREADS)
k_record_meta(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, u32 *sizes, u64 stride, FqzWinStatus *st,
              u64 phred_records /* records (file-global index <) taking part in Phred detection */) {
    const int W = FQZ_META_GROUP;
    u32 g = threadIdx.x & (W - 1);
    u32 gm = group_mask(W);
    u64 r = (u64)blockIdx.x * (FQZ_META_THREADS / W) + (threadIdx.x / W);
    if (r > R) return;  // whole group leaves together
    if (r == R) {
        // trailing partial record: the reference still validates the lines it can read before it
        // meets EOF (parser.go:138-166), e.g. a trailing blank line is a header error.
        if (g == 0 && tail_lines > 0) {
            u64 l = 4 * R;
            u32 hs = line_end[(long long)l - 1] + 1u;
            u32 he = strip_cr(text, hs, line_end[l]);
            u32 kind = 0;
            if (he == hs || text[hs] != '@') kind = FQZ_K_HEADER_AT;
            else if (tail_lines >= 3) {
                u32 ps = line_end[l + 1] + 1u;
                u32 pe = strip_cr(text, ps, line_end[l + 2]);
                if (pe == ps || text[ps] != '+') kind = FQZ_K_PLUS;
            }
            if (kind) atomicMin(&st->err_key, ((rec_base + R) << 8) | kind);
        }
        return;
    }
    LineSpan t = record_lines(text, line_end, r);
    u32 L = t.se - t.ss;
    u32 kind = 0;
    if (t.he == t.hs || text[t.hs] != '@') kind = FQZ_K_HEADER_AT;
    else if (t.pe == t.ps || text[t.ps] != '+') kind = FQZ_K_PLUS;
    else if (L != t.qe - t.qs) kind = FQZ_K_LEN;
    // N bases inside the tracked range, and the guard beyond it
    u32 lim = t.ss + min(L, FQZ_MAX_SEQ_LEN);
    u32 nn = group_sum(partial_count_n<W>(text, t.ss, lim, g), gm, W);
    if (L > FQZ_MAX_SEQ_LEN) {
        u32 beyond = group_sum(partial_count_n<W>(text, lim, t.se, g), gm, W);
        if (beyond && !kind) kind = FQZ_K_LONG_N;
    }
    if (rec_base + r < phred_records) {
        u32 mn = group_min(partial_min_byte<W>(text, t.qs, t.qe, g), gm, W);
        if (g == 0 && mn < 255u) atomicMin(&st->qual_min, mn);
    }
    if (g == 0) {
        if (kind) atomicMin(&st->err_key, ((rec_base + r) << 8) | kind);
        u32 H = (kind == FQZ_K_HEADER_AT) ? 0u : (t.he - t.hs - 1u);
        u32 P = (kind == FQZ_K_HEADER_AT || kind == FQZ_K_PLUS) ? 0u : (t.pe - t.ps - 1u);
        sizes[0 * stride + r] = (L + 3u) >> 2;
        sizes[1 * stride + r] = L;
        sizes[2 * stride + r] = 2u + H;
        sizes[3 * stride + r] = 2u + P;
        sizes[4 * stride + r] = 2u + 2u * nn;
    }
}

// Decide the file-global Phred flag from the min quality byte of block 0
// (reference: quality.go:22-49: any byte < 59 -> 33; min >= 64 -> 64; none / 59..63 -> 33).
__global__ void k_decide_phred(const FqzWinStatus *st, u32 *phred64) {
    u32 m = st->qual_min;
    *phred64 = (m != 255u && m >= 64u) ? 1u : 0u;
}

// ---------------------------------------------------------------------------------- stream scatter
struct Src {  // text bytes, possibly served from a shared-memory copy of [bias, bias+len)
    const u8 *p;
    u32 bias;
    __device__ __forceinline__ const u8 *at(u32 pos) const { return p + (pos - bias); }
};

// copy n bytes src->dst with a W-lane group: bytes up to dst alignment, aligned words, tail bytes
template <int W>
__device__ __forceinline__ void group_copy(u8 *dst, const u8 *src, u32 n, u32 g) {
    u32 head = (u32)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if (head > n) head = n;
    if (g < head) dst[g] = src[g];
    u32 nw = (n - head) >> 2;
    for (u32 w = g; w < nw; w += W) *(u32 *)(dst + head + 4u * w) = ld_u32_unaligned(src + head + 4u * w);
    u32 t0 = head + 4u * nw;
    if (g < n - t0) dst[t0 + g] = src[t0 + g];
}

// Writes one record into the six streams (a W-lane group).  t: its four lines; L, H, P: sequence / header / plus
// payload lengths as sized by the meta pass; o_*: the record's offsets in the streams; has_n: it holds non-ACGT bases.
struct ScStreams {
    u8 *seq, *qual, *hdr, *plus, *npos, *len;
};
template <int W>
__device__ __forceinline__ void scatter_record(const Src &src, const LineSpan &t, u64 r, u32 L, u32 H, u32 P, u32 o_seq, u32 o_qual, u32 o_hdr,
                                               u32 o_plus, u32 o_npos, bool has_n, u8 off, const ScStreams &S, u32 g, u32 gm) {
    u8 *s_seq = S.seq, *s_qual = S.qual, *s_hdr = S.hdr, *s_plus = S.plus, *s_npos = S.npos, *s_len = S.len;
    // seqLengths: u32 L (compress.go:501)
    if (g == 0) *(u32 *)(s_len + 4ull * r) = L;
    // headers / plusLines: u16 length prefix + bytes without the leading '@' / '+' (compress.go:514-519)
    if (g == 0) st_u16_unaligned(s_hdr + o_hdr, H & 0xFFFFu);
    group_copy<W>(s_hdr + o_hdr + 2, src.at(t.hs + 1u), H, g);
    if (g == 0) st_u16_unaligned(s_plus + o_plus, P & 0xFFFFu);
    group_copy<W>(s_plus + o_plus + 2, src.at(t.ps + 1u), P, g);
    // seqPacked + nPositions (sequence.go:139-184): 16 bases per lane per round.  k_record_meta has
    // already counted the record's non-ACGT bases (that count sized the N-position stream): the 99 % of
    // reads without any take a path with no N test, no position bookkeeping and no group scan.
    u32 units = (L + 15u) >> 4, rounds = (units + W - 1) / W;
    u32 nbase = 0;
    if (!has_n) {
        for (u32 u = g; u < units; u += W) {
            const u8 *sp = src.at(t.ss + 16u * u);
            u32 out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 bp = 16u * u + 4u * k;  // base index of this word's first byte
                if (bp < L) {
                    u32 x = ld_u32_unaligned(sp + 4 * k);
                    u32 codes = base_codes4(x, 0u);
                    if (bp + 4u > L) codes &= 0xFFFFFFFFu >> (8u * (bp + 4u - L));  // bytes past the read
                    out |= pack_codes4(codes) << (8 * k);
                }
            }
            u32 vb = min(16u, L - 16u * u);
            st_bytes(s_seq + o_seq + 4u * u, out, (vb + 3u) >> 2);
        }
    } else
    for (u32 j = 0; j < rounds; j++) {
        u32 u = j * W + g;
        u32 nmask16 = 0;
        if (u < units) {
            const u8 *sp = src.at(t.ss + 16u * u);
            u32 out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 bp = 16u * u + 4u * k;  // base index of this word's first byte
                if (bp < L) {
                    u32 x = ld_u32_unaligned(sp + 4 * k);
                    u32 rm = range_mask4(bp, 0u, L);
                    u32 nm = nonacgt_mask4(x) & rm;
                    u32 codes = base_codes4(x, nm) & rm;
                    out |= pack_codes4(codes) << (8 * k);
                    nmask16 |= ((((nm & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << (4 * k);
                }
            }
            u32 vb = min(16u, L - 16u * u);
            st_bytes(s_seq + o_seq + 4u * u, out, (vb + 3u) >> 2);
            if (16u * u >= FQZ_MAX_SEQ_LEN) nmask16 = 0;  // positions >= 65536 are not tracked
        }
        u32 cnt = (u32)__popc(nmask16);
        u32 incl = group_incl_scan(cnt, gm, W);
        u32 tot = __shfl_sync(gm, incl, W - 1, W);
        u32 k = nbase + incl - cnt;
        while (nmask16) {
            int b = __ffs((int)nmask16) - 1;
            nmask16 &= nmask16 - 1;
            st_u16_unaligned(s_npos + o_npos + 2u + 2u * k, 16u * u + (u32)b);
            k++;
        }
        nbase += tot;
    }
    if (g == 0) st_u16_unaligned(s_npos + o_npos, nbase & 0xFFFFu);  // compress.go:495 (silent u16 truncation)
    // quality: normalise + per-record delta (compress.go:506-510; quality.go:53-103)
    {
        u8 *dq = s_qual + o_qual;
        const u8 *sq = src.at(t.qs);
        u32 head = (u32)((4u - ((uintptr_t)dq & 3u)) & 3u);
        if (head == 0) head = 4;  // keeps every body word at index >= 1 so q[i-1] exists
        if (head > L) head = L;
        if (g < head) dq[g] = (u8)(sq[g] - (g ? sq[g - 1] : off));
        u32 nw = (L - head) >> 2;
        for (u32 w = g; w < nw; w += W) {
            u32 i = head + 4u * w;
            u32 cur = ld_u32_unaligned(sq + i);
            u32 prv = ld_u32_unaligned(sq + i - 1);
            *(u32 *)(dq + i) = __vsub4(cur, prv);
        }
        u32 t0 = head + 4u * nw;
        if (g < L - t0) {
            u32 i = t0 + g;
            dq[i] = (u8)(sq[i] - (i ? sq[i - 1] : off));
        }
    }
}

__global__ void __launch_bounds__(FQZ_SC_THREADS)
k_scatter_streams(const u8 *text, const u32 *line_end, u64 R, const u32 *offs, u64 stride, const u32 *phred64, u8 *s_seq, u8 *s_qual,
                  u8 *s_hdr, u8 *s_plus, u8 *s_npos, u8 *s_len, u32 stage_cap) {
    FQZ_DYN_SMEM(u8, smem);
    const int W = FQZ_SC_GROUP;
    u64 *bar = (u64 *)smem;
    u8 *stage = smem + 128;
    u64 r0 = (u64)blockIdx.x * FQZ_SC_RPC;
    u64 r1 = min(r0 + (u64)FQZ_SC_RPC, R);
    // contiguous text chunk of this CTA's records, widened to 16-byte boundaries for the bulk copy
    u32 c0 = line_end[4ll * (long long)r0 - 1] + 1u;
    u32 c1 = line_end[4 * r1 - 1] + 1u;
    u32 a0 = c0 & ~15u;
    u32 a1 = (c1 + 15u) & ~15u;
    Src src;
    if (a1 - a0 + 16u <= stage_cap) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        cta_stage_bulk(stage, text + a0, a1 - a0, bar, 0);
        src.p = stage;
        src.bias = a0;
    } else {  // chunk larger than the staging buffer (long reads): read HBM/L2 directly
        src.p = text;
        src.bias = 0;
    }
    const u8 off = *phred64 ? 64 : 33;
    ScStreams S = {s_seq, s_qual, s_hdr, s_plus, s_npos, s_len};
    u32 g = threadIdx.x & (W - 1);
    u32 gm = group_mask(W);
    for (u64 r = r0 + threadIdx.x / W; r < r1; r += FQZ_SC_THREADS / W) {
        // line_end holds window offsets; CR stripping must look at the same bytes the meta kernel saw
        LineSpan t;
        {
            u64 l = 4 * r;
            u32 e0 = line_end[(long long)l - 1];
            u32 e1 = line_end[l], e2 = line_end[l + 1], e3 = line_end[l + 2], e4 = line_end[l + 3];
            t.hs = e0 + 1u;
            t.he = (e1 > t.hs && *src.at(e1 - 1) == '\r') ? e1 - 1 : e1;
            t.ss = e1 + 1u;
            t.se = (e2 > t.ss && *src.at(e2 - 1) == '\r') ? e2 - 1 : e2;
            t.ps = e2 + 1u;
            t.pe = (e3 > t.ps && *src.at(e3 - 1) == '\r') ? e3 - 1 : e3;
            t.qs = e3 + 1u;
            t.qe = (e4 > t.qs && *src.at(e4 - 1) == '\r') ? e4 - 1 : e4;
        }
        u32 L = t.se - t.ss;
        u32 H = t.he - t.hs - 1u, P = t.pe - t.ps - 1u;
        u32 o_seq = offs[0 * stride + r], o_qual = offs[1 * stride + r], o_hdr = offs[2 * stride + r];
        u32 o_plus = offs[3 * stride + r], o_npos = offs[4 * stride + r];
        const bool has_n = (offs[4 * stride + r + 1] - o_npos) > 2u;
        scatter_record<W>(src, t, r, L, H, P, o_seq, o_qual, o_hdr, o_plus, o_npos, has_n, off, S, g, gm);
    }
}

// ---------------------------------------------------------------------------------- fused metadata + scatter (one pass over the text)
// k_record_meta + five device-wide scans + k_scatter_streams read the text twice and keep the per-record sizes in HBM
// in between.  Here a CTA stages its 64 records once (TMA), sizes them out of shared memory (validation, N count,
// long-read guard: same rules as k_record_meta), scans the 64 x 5 sizes locally, gets the five stream offsets of its
// first record by a decoupled look-back over the CTAs in front of it (each CTA publishes its five totals, then the
// running prefix; CTA numbers are handed out by an atomic ticket so that a CTA only ever waits for CTAs that are
// already running), writes the scanned offsets (the entropy stage wants them) and scatters.  The streams are
// allocated at fixed generous sizes before anything is known: a window whose streams do not fit says so in the
// status word and is redone by the separate kernels.  Phred detection (block 0 of the file) and the check of a
// trailing partial record run as two small kernels in front.
struct ScCaps {
    u32 cap[5];
};
#define SC_FLAG_AGG (1ull << 62)
#define SC_FLAG_INC (2ull << 62)
#define SC_VAL_MASK ((1ull << 62) - 1ull)
__global__ void __launch_bounds__(FQZ_META_THREADS) k_phred_min(const u8 *text, const u32 *line_end, u64 nrec, FqzWinStatus *st) {
    const int W = FQZ_META_GROUP;
    u32 g = threadIdx.x & (W - 1);
    u32 gm = group_mask(W);
    u64 r = (u64)blockIdx.x * (FQZ_META_THREADS / W) + (threadIdx.x / W);
    if (r >= nrec) return;
    LineSpan t = record_lines(text, line_end, r);
    u32 mn = group_min(partial_min_byte<W>(text, t.qs, t.qe, g), gm, W);
    if (g == 0 && mn < 255u) atomicMin(&st->qual_min, mn);
}
// trailing partial record: the reference still validates the lines it can read before it meets EOF (parser.go:138-166)
__global__ void k_tail_check(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, FqzWinStatus *st) {
    if (threadIdx.x || blockIdx.x || tail_lines == 0) return;
    u64 l = 4 * R;
    u32 hs = line_end[(long long)l - 1] + 1u;
    u32 he = strip_cr(text, hs, line_end[l]);
    u32 kind = 0;
    if (he == hs || text[hs] != '@') kind = FQZ_K_HEADER_AT;
    else if (tail_lines >= 3) {
        u32 ps = line_end[l + 1] + 1u;
        u32 pe = strip_cr(text, ps, line_end[l + 2]);
        if (pe == ps || text[ps] != '+') kind = FQZ_K_PLUS;
    }
    if (kind) atomicMin(&st->err_key, ((rec_base + R) << 8) | kind);
}
__global__ void __launch_bounds__(FQZ_SC_THREADS)
k_scatter_fused(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 *offs, u64 stride, const u32 *phred64, ScStreams S, ScCaps caps,
                FqzWinStatus *st, unsigned long long *look, u32 *ticket, u32 stage_cap) {
    FQZ_DYN_SMEM(u8, smem);
    const int W = FQZ_SC_GROUP;
    __shared__ u32 s_cta;
    __shared__ u32 s_sz[5][FQZ_SC_RPC + 1];  // sizes, then exclusive offsets inside the CTA
    __shared__ u32 s_own[FQZ_SC_RPC];        // bytes of the record's N-position item
    __shared__ u32 s_len[3][FQZ_SC_RPC];     // L, H, P as sized
    __shared__ u32 s_base[5], s_tot[5];
    __shared__ u32 s_skip;
    if (threadIdx.x == 0) {
        s_cta = atomicAdd(ticket, 1u);
        s_skip = 0;
    }
    __syncthreads();
    const u32 cta = s_cta;
    u64 *bar = (u64 *)smem;
    u8 *stage = smem + 128;
    u64 r0 = (u64)cta * FQZ_SC_RPC;
    u64 r1 = min(r0 + (u64)FQZ_SC_RPC, R);
    // contiguous text chunk of this CTA's records, widened to 16-byte boundaries for the bulk copy
    u32 c0 = line_end[4ll * (long long)r0 - 1] + 1u;
    u32 c1 = line_end[4 * r1 - 1] + 1u;
    u32 a0 = c0 & ~15u;
    u32 a1 = (c1 + 15u) & ~15u;
    Src src;
    if (a1 - a0 + 16u <= stage_cap) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        cta_stage_bulk(stage, text + a0, a1 - a0, bar, 0);
        src.p = stage;
        src.bias = a0;
    } else {  // chunk larger than the staging buffer (long reads): read HBM/L2 directly
        src.p = text;
        src.bias = 0;
    }
    u32 g = threadIdx.x & (W - 1);
    u32 gm = group_mask(W);
    // ---- 1. size the records (rules of k_record_meta)
    for (u64 r = r0 + threadIdx.x / W; r < r1; r += FQZ_SC_THREADS / W) {
        const u32 i = (u32)(r - r0);
        u64 l = 4 * r;
        u32 e0 = line_end[(long long)l - 1];
        u32 e1 = line_end[l], e2 = line_end[l + 1], e3 = line_end[l + 2], e4 = line_end[l + 3];
        u32 hs = e0 + 1u, he = (e1 > hs && *src.at(e1 - 1) == '\r') ? e1 - 1 : e1;
        u32 ss = e1 + 1u, se = (e2 > ss && *src.at(e2 - 1) == '\r') ? e2 - 1 : e2;
        u32 ps = e2 + 1u, pe = (e3 > ps && *src.at(e3 - 1) == '\r') ? e3 - 1 : e3;
        u32 qs = e3 + 1u, qe = (e4 > qs && *src.at(e4 - 1) == '\r') ? e4 - 1 : e4;
        u32 L = se - ss;
        u32 kind = 0;
        if (he == hs || *src.at(hs) != '@') kind = FQZ_K_HEADER_AT;
        else if (pe == ps || *src.at(ps) != '+') kind = FQZ_K_PLUS;
        else if (L != qe - qs) kind = FQZ_K_LEN;
        // non-ACGT bases inside the tracked range, and the guard beyond it: src.p - src.bias is 16-byte aligned like the text
        const u8 *tb = src.p - src.bias;
        u32 lim = ss + min(L, FQZ_MAX_SEQ_LEN);
        u32 nn = group_sum(partial_count_n<W>(tb, ss, lim, g), gm, W);
        if (L > FQZ_MAX_SEQ_LEN) {
            u32 beyond = group_sum(partial_count_n<W>(tb, lim, se, g), gm, W);
            if (beyond && !kind) kind = FQZ_K_LONG_N;
        }
        if (g == 0) {
            if (kind) atomicMin(&st->err_key, ((rec_base + r) << 8) | kind);
            u32 H = (kind == FQZ_K_HEADER_AT) ? 0u : (he - hs - 1u);
            u32 P = (kind == FQZ_K_HEADER_AT || kind == FQZ_K_PLUS) ? 0u : (pe - ps - 1u);
            s_sz[0][i] = (L + 3u) >> 2;
            s_sz[1][i] = L;
            s_sz[2][i] = 2u + H;
            s_sz[3][i] = 2u + P;
            s_sz[4][i] = 2u + 2u * nn;
            s_own[i] = 2u + 2u * nn;
            s_len[0][i] = L;
            s_len[1][i] = H;
            s_len[2][i] = P;
        }
    }
    __syncthreads();
    // ---- 2. local scan (five threads, 64 entries each) and 3. look-back
    if (threadIdx.x < 5) {
        const u32 a = threadIdx.x, cnt = (u32)(r1 - r0);
        u32 run = 0;
        for (u32 i = 0; i < cnt; i++) {
            u32 v = s_sz[a][i];
            s_sz[a][i] = run;
            run += v;
        }
        s_tot[a] = run;
        volatile unsigned long long *L5 = look;
        unsigned long long base = 0;
        if (cta == 0) {
            L5[(u64)cta * 5 + a] = SC_FLAG_INC | run;
        } else {
            L5[(u64)cta * 5 + a] = SC_FLAG_AGG | run;
            __threadfence();
            long long p = (long long)cta - 1;
            for (;;) {
                unsigned long long v = L5[(u64)p * 5 + a];
                if ((v >> 62) == 0) {
#ifdef FQZ_EMU
                    emu::yield();
#endif
                    continue;
                }
                base += v & SC_VAL_MASK;
                if ((v >> 62) == 2) break;
                p--;
            }
            __threadfence();
            L5[(u64)cta * 5 + a] = SC_FLAG_INC | (base + run);
        }
        s_base[a] = (u32)base;
        if (base + run > (unsigned long long)caps.cap[a]) {  // the fixed-size stream is too small: the window is redone
            s_skip = 1;
            atomicOr(&st->pad, 1u);
        }
    }
    __syncthreads();
    if (r1 == R && threadIdx.x < 5) offs[threadIdx.x * stride + R] = s_base[threadIdx.x] + s_tot[threadIdx.x];
    // ---- 4. offsets out, then the scatter proper
    const bool skip = s_skip != 0;
    const u8 off = *phred64 ? 64 : 33;
    for (u64 r = r0 + threadIdx.x / W; r < r1; r += FQZ_SC_THREADS / W) {
        const u32 i = (u32)(r - r0);
        u32 o_seq = s_base[0] + s_sz[0][i], o_qual = s_base[1] + s_sz[1][i], o_hdr = s_base[2] + s_sz[2][i];
        u32 o_plus = s_base[3] + s_sz[3][i], o_npos = s_base[4] + s_sz[4][i];
        if (g == 0) {
            offs[0 * stride + r] = o_seq;
            offs[1 * stride + r] = o_qual;
            offs[2 * stride + r] = o_hdr;
            offs[3 * stride + r] = o_plus;
            offs[4 * stride + r] = o_npos;
        }
        if (skip) continue;
        LineSpan t;
        {
            u64 l = 4 * r;
            u32 e0 = line_end[(long long)l - 1];
            u32 e1 = line_end[l], e2 = line_end[l + 1], e3 = line_end[l + 2], e4 = line_end[l + 3];
            t.hs = e0 + 1u;
            t.he = (e1 > t.hs && *src.at(e1 - 1) == '\r') ? e1 - 1 : e1;
            t.ss = e1 + 1u;
            t.se = (e2 > t.ss && *src.at(e2 - 1) == '\r') ? e2 - 1 : e2;
            t.ps = e2 + 1u;
            t.pe = (e3 > t.ps && *src.at(e3 - 1) == '\r') ? e3 - 1 : e3;
            t.qs = e3 + 1u;
            t.qe = (e4 > t.qs && *src.at(e4 - 1) == '\r') ? e4 - 1 : e4;
        }
        scatter_record<W>(src, t, r, s_len[0][i], s_len[1][i], s_len[2][i], o_seq, o_qual, o_hdr, o_plus, o_npos, s_own[i] > 2u, off, S, g, gm);
    }
}

// ---------------------------------------------------------------------------------- host launchers
void fqz_launch_newline_count(const u8 *text, u64 n, u64 lo, u32 *tile_counts, u32 ntiles, cudaStream_t s) {
    if (ntiles) FQZ_LAUNCH(k_newline_count, ntiles, FQZ_NL_THREADS, 0, s, text, n, lo, tile_counts);
}
void fqz_launch_newline_scan(const u8 *text, u64 n, u64 lo, u32 *line_end, u32 cap_lines, unsigned long long *look, u32 ntiles, u32 *total_lines,
                             cudaStream_t s) {
    if (ntiles) FQZ_LAUNCH(k_newline_scan, ntiles, FQZ_NL_THREADS, 0, s, text, n, lo, line_end, cap_lines, look, ntiles, total_lines);
}
void fqz_launch_find_newline(const u8 *text, u64 n, const u32 *tile_prefix, u32 ntiles, u32 target, u64 *out_pos, cudaStream_t s) {
    if (ntiles) FQZ_LAUNCH(k_find_newline, 1, FQZ_NL_THREADS, 0, s, text, n, tile_prefix, ntiles, target, out_pos);
}
void fqz_launch_newline_index(const u8 *text, u64 n, u64 lo, const u32 *tile_prefix, u32 ntiles, u32 *line_end, u32 max_lines, cudaStream_t s) {
    if (ntiles) FQZ_LAUNCH(k_newline_index, ntiles, FQZ_NL_THREADS, 0, s, text, n, lo, tile_prefix, line_end, max_lines);
}
void fqz_launch_scan_partial(const u32 *data, u64 n, u64 stride, u32 narr, u32 *sums, u32 ntiles, cudaStream_t s) {
    FQZ_LAUNCH(k_scan_partial, dim3(ntiles, narr), FQZ_SCAN_THREADS, 0, s, data, n, stride, sums, ntiles);
}
void fqz_launch_scan_apply(u32 *data, u64 n, u64 stride, u32 narr, const u32 *sums, u32 ntiles, cudaStream_t s) {
    FQZ_LAUNCH(k_scan_apply, dim3(ntiles, narr), FQZ_SCAN_THREADS, 0, s, data, n, stride, sums, ntiles);
}
void fqz_launch_record_meta(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, u32 *sizes, u64 stride,
                            FqzWinStatus *st, u64 phred_records, cudaStream_t s) {
    u32 per = FQZ_META_THREADS / FQZ_META_GROUP;
    u32 grid = (u32)((R + 1 + per - 1) / per);
    FQZ_LAUNCH(k_record_meta, grid, FQZ_META_THREADS, 0, s, text, line_end, R, rec_base, tail_lines, sizes, stride, st, phred_records);
}
void fqz_launch_decide_phred(const FqzWinStatus *st, u32 *phred64, cudaStream_t s) { FQZ_LAUNCH(k_decide_phred, 1, 1, 0, s, st, phred64); }
void fqz_launch_scatter(const u8 *text, const u32 *line_end, u64 R, const u32 *offs, u64 stride, const u32 *phred64, u8 *const streams[6],
                        cudaStream_t s) {
    if (!R) return;
    u32 grid = (u32)((R + FQZ_SC_RPC - 1) / FQZ_SC_RPC);
    FQZ_LAUNCH(k_scatter_streams, grid, FQZ_SC_THREADS, FQZ_SC_SMEM, s, text, line_end, R, offs, stride, phred64, streams[0], streams[1],
               streams[2], streams[3], streams[4], streams[5], (u32)(FQZ_SC_SMEM - 128));
}
// per-device kernel attributes: called once per context (the attribute is per device, and one process may drive several GPUs)
void fqz_launch_phred_min(const u8 *text, const u32 *line_end, u64 nrec, FqzWinStatus *st, cudaStream_t s) {
    if (!nrec) return;
    u32 per = FQZ_META_THREADS / FQZ_META_GROUP;
    FQZ_LAUNCH(k_phred_min, (u32)((nrec + per - 1) / per), FQZ_META_THREADS, 0, s, text, line_end, nrec, st);
}
void fqz_launch_scatter_fused(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, u32 *offs, u64 stride, const u32 *phred64,
                              u8 *const streams[6], const u32 caps[5], FqzWinStatus *st, unsigned long long *look, u32 *ticket, cudaStream_t s) {
    if (tail_lines) FQZ_LAUNCH(k_tail_check, 1, 32, 0, s, text, line_end, R, rec_base, tail_lines, st);
    if (!R) return;
    u32 grid = (u32)((R + FQZ_SC_RPC - 1) / FQZ_SC_RPC);
    ScStreams S = {streams[0], streams[1], streams[2], streams[3], streams[4], streams[5]};
    ScCaps C;
    for (int a = 0; a < 5; a++) C.cap[a] = caps[a];
    FQZ_LAUNCH(k_scatter_fused, grid, FQZ_SC_THREADS, FQZ_SC_SMEM, s, text, line_end, R, rec_base, offs, stride, phred64, S, C, st, look, ticket,
               (u32)(FQZ_SC_SMEM - 128));
}
int fqz_frontend_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_scatter_streams, cudaFuncAttributeMaxDynamicSharedMemorySize, FQZ_SC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scatter_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, FQZ_SC_SMEM);
    return (int)e;
}
