// fqz_zstd_enc.cu — GPU zstd (RFC 8878) frame encoder: the entropy stage of the block codec.
// Replaces zstd.Encoder.EncodeAll (reference call sites internal/compress/compress.go:523-528,
// options :115-118: SpeedFastest, checksum on, no dictionary).  The reference's encoder is the
// third-party klauspost/compress v1.19.1 (go.mod:8), absent from the tree; compressed bytes are not
// pinned by any reference test, so the contract is: valid RFC 8878 frames with content checksums
// that the reference's decoder accepts, ratio within 2 % of the CPU path.
//
// One WARP encodes one frame (<= 64 KiB of one stream) start to finish; thousands of frames are in
// flight per launch.  Two policies:
//   ENTROPY : literals-only compressed block (Huffman, 4 streams) / RLE block / raw block
//   AUTO    : greedy LZ77 (32 positions per step: ballot + match_any candidate search, hash table
//             in shared memory), Huffman literals, FSE sequences (predefined / RLE / dynamic)
#include <type_traits>

#include "fqz_zstd.h"
#include "fqz_zstd_tables.cuh"

#define FULL 0xffffffffu
#define HLOG 12
#define HEMPTY 0xFFFFu

// ---------------------------------------------------------------------------------- per-warp shared scratch
struct FseCT {       // encoding table of one FSE stream
    u16 tab[512];    // next-state table
    u32 dnb[64];     // deltaNbBits per symbol
    int dfs[64];     // deltaFindState per symbol
};
struct HufTmp {
    u32 node_w[512];
    u16 node_par[512];
    u8 node_dep[512];
    u8 ssym[256];
    u8 len[256];
    u16 wtab[64];  // FSE table of the Huffman weights
    u32 wdnb[16];
    int wdfs[16];
};
union EntropyScratch {
    HufTmp huf;
    FseCT fse[3];  // LL, OF, ML
};
struct WarpScratchLZ {
    u32 hist[256];
    u16 hlut[256];  // code << 4 | nbBits
    union {
        u16 htab[1 << HLOG];
        EntropyScratch e;
    } u;
    u8 tmpsym[512];  // FSE symbol spreading
    short norm[64];
};
struct WarpScratchEnt {
    u32 hist[256];
    u16 hlut[256];
    union {
        EntropyScratch e;
    } u;
    u8 tmpsym[512];
    short norm[64];
};

// ---------------------------------------------------------------------------------- bit writer (one lane)
struct BitW {
    u8 *base;
    u32 pos, cap;
    u64 acc;
    u32 nb;
    bool ovf;
    __device__ void init(u8 *b, u32 c) {
        base = b;
        pos = 0;
        cap = c;
        acc = 0;
        nb = 0;
        ovf = false;
    }
    __device__ __forceinline__ void add(u32 v, u32 n) {  // n <= 32, v < 2^n
        acc |= (u64)v << nb;
        nb += n;
        if (nb >= 32) {
            if (pos + 4 <= cap) st_u32_unaligned(base + pos, (u32)acc);
            else ovf = true;
            pos += 4;
            acc >>= 32;
            nb -= 32;
        }
    }
    __device__ u32 close() {  // end mark + flush; returns byte length
        add(1, 1);
        u32 nbytes = (nb + 7) >> 3;
        for (u32 i = 0; i < nbytes; i++) {
            if (pos < cap) base[pos] = (u8)(acc >> (8 * i));
            else ovf = true;
            pos++;
        }
        return pos;
    }
};

// ---------------------------------------------------------------------------------- FSE (one lane)
__device__ static u32 fse_optimal_log(u32 maxLog, u32 n, u32 maxSym) {
    u32 maxBitsSrc = (n > 1) ? hibit32(n - 1) : 0;
    maxBitsSrc = maxBitsSrc > 2 ? maxBitsSrc - 2 : 0;
    u32 minBitsSrc = hibit32(n) + 1, minBitsSym = hibit32(maxSym ? maxSym : 1) + 2;
    u32 minBits = min(minBitsSrc, minBitsSym);
    u32 t = min(maxLog, maxBitsSrc);
    t = max(t, minBits);
    t = max(t, 5u);
    return min(t, maxLog);
}
// counts[0..maxSym] summing to total -> norm[] summing to 1<<tlog, every present symbol >= 1
__device__ static void fse_normalize(const u32 *counts, u32 maxSym, u32 total, u32 tlog, short *norm) {
    u32 tsize = 1u << tlog;
    int sum = 0;
    u32 largest = 0;
    int largestN = -1;
    for (u32 s = 0; s <= maxSym; s++) {
        int nrm = 0;
        if (counts[s]) {
            nrm = (int)(((u64)counts[s] * tsize + total / 2) / total);
            if (nrm == 0) nrm = 1;
            if (nrm > largestN) {
                largestN = nrm;
                largest = s;
            }
        }
        norm[s] = (short)nrm;
        sum += nrm;
    }
    int diff = (int)tsize - sum;
    if (diff > 0) norm[largest] = (short)(norm[largest] + diff);
    while (diff < 0) {  // take the excess from the currently largest symbols, at most half of one at a time
        u32 b = 0;
        int bn = 0;
        for (u32 s = 0; s <= maxSym; s++)
            if (norm[s] > bn) {
                bn = norm[s];
                b = s;
            }
        if (bn <= 1) break;  // cannot happen: tsize >= number of present symbols
        int take = min(-diff, max(1, bn / 2));
        take = min(take, bn - 1);
        norm[b] = (short)(norm[b] - take);
        diff += take;
    }
}
// RFC 8878 §4.1.1 table description.  Returns bytes written.
__device__ static u32 fse_write_ncount(u8 *out, const short *norm, u32 maxSym, u32 tlog) {
    u32 tsize = 1u << tlog;
    u32 bitStream = tlog - 5, bitCount = 4;
    int remaining = (int)tsize + 1, threshold = (int)tsize;
    u32 nbBits = tlog + 1;
    u32 sym = 0, alphabet = maxSym + 1;
    bool prev0 = false;
    u32 o = 0;
    while (sym < alphabet && remaining > 1) {
        if (prev0) {
            u32 start = sym;
            while (sym < alphabet && !norm[sym]) sym++;
            if (sym == alphabet) break;
            while (sym >= start + 24) {
                start += 24;
                bitStream += 0xFFFFu << bitCount;
                out[o++] = (u8)bitStream;
                out[o++] = (u8)(bitStream >> 8);
                bitStream >>= 16;
            }
            while (sym >= start + 3) {
                start += 3;
                bitStream += 3u << bitCount;
                bitCount += 2;
            }
            bitStream += (sym - start) << bitCount;
            bitCount += 2;
            if (bitCount > 16) {
                out[o++] = (u8)bitStream;
                out[o++] = (u8)(bitStream >> 8);
                bitStream >>= 16;
                bitCount -= 16;
            }
        }
        {
            int count = norm[sym++];
            int mx = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += mx;
            bitStream += (u32)count << bitCount;
            bitCount += nbBits;
            bitCount -= (count < mx);
            prev0 = (count == 1);
            while (remaining < threshold) {
                nbBits--;
                threshold >>= 1;
            }
        }
        if (bitCount > 16) {
            out[o++] = (u8)bitStream;
            out[o++] = (u8)(bitStream >> 8);
            bitStream >>= 16;
            bitCount -= 16;
        }
    }
    out[o] = (u8)bitStream;
    out[o + 1] = (u8)(bitStream >> 8);
    o += (bitCount + 7) / 8;
    return o;
}
// Encoding table from a normalized distribution (symbols with norm -1 take one cell at the end).
__device__ static void fse_build_ctable(const short *norm, u32 maxSym, u32 tlog, u16 *tab, u32 *dnb, int *dfs, u8 *tsym) {
    u32 tsize = 1u << tlog, mask = tsize - 1, step = (tsize >> 1) + (tsize >> 3) + 3;
    u32 cumul[66];
    u32 high = tsize - 1;
    cumul[0] = 0;
    for (u32 u = 1; u <= maxSym + 1; u++) {
        if (norm[u - 1] == -1) {
            cumul[u] = cumul[u - 1] + 1;
            tsym[high--] = (u8)(u - 1);
        } else
            cumul[u] = cumul[u - 1] + (u32)norm[u - 1];
    }
    u32 pos = 0;
    for (u32 s = 0; s <= maxSym; s++)
        for (int i = 0; i < norm[s]; i++) {
            tsym[pos] = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (u32 u = 0; u < tsize; u++) {
        u32 s = tsym[u];
        tab[cumul[s]++] = (u16)(tsize + u);
    }
    int total = 0;
    for (u32 s = 0; s <= maxSym; s++) {
        int nrm = norm[s];
        if (nrm == 0) {
            dnb[s] = ((tlog + 1) << 16) - tsize;
            dfs[s] = 0;
        } else if (nrm == -1 || nrm == 1) {
            dnb[s] = (tlog << 16) - tsize;
            dfs[s] = total - 1;
            total++;
        } else {
            u32 maxBitsOut = tlog - hibit32((u32)nrm - 1);
            u32 minStatePlus = (u32)nrm << maxBitsOut;
            dnb[s] = (maxBitsOut << 16) - minStatePlus;
            dfs[s] = total - nrm;
            total += nrm;
        }
    }
}
struct FseState {
    u32 st;
    const u16 *tab;
    const u32 *dnb;
    const int *dfs;
    u32 tlog;  // 0 = RLE mode: no bits at all
    __device__ void init(const u16 *t, const u32 *d, const int *f, u32 lg, u32 sym) {
        tab = t;
        dnb = d;
        dfs = f;
        tlog = lg;
        st = 0;
        if (lg) {
            u32 nbo = (d[sym] + (1u << 15)) >> 16;
            u32 v = (nbo << 16) - d[sym];
            st = t[(int)(v >> nbo) + f[sym]];
        }
    }
    __device__ __forceinline__ void encode(BitW &bw, u32 sym) {
        if (!tlog) return;
        u32 nbo = (st + dnb[sym]) >> 16;
        bw.add(st & ((1u << nbo) - 1u), nbo);
        st = tab[(int)(st >> nbo) + dfs[sym]];
    }
    __device__ void flush(BitW &bw) {
        if (tlog) bw.add(st & ((1u << tlog) - 1u), tlog);
    }
};

// ---------------------------------------------------------------------------------- Huffman (warp)
// hist[256] -> code lengths (<= 11 bits) and canonical codes in hlut; weights in H.len (as
// nbBits).  Returns maxBits (0 if fewer than 2 distinct symbols).  *maxSymOut = last present symbol.
__device__ static u32 warp_huf_build(const u32 *hist, u32 total, u16 *hlut, HufTmp &H, u32 *maxSymOut) {
    u32 lane = lane_id();
    u32 present = 0, maxSym = 0;
    for (u32 i = 0; i < 8; i++) {
        u32 s = lane * 8 + i;
        if (hist[s]) {
            present++;
            maxSym = s;
        }
    }
    u32 m = __reduce_add_sync(FULL, present);
    maxSym = __reduce_max_sync(FULL, maxSym);
    *maxSymOut = maxSym;
    if (m < 2) return 0;
    u32 maxBits = 0;
    for (int iter = 0;; iter++) {
        u32 flo = 0;
        bool ones = false;
        if (iter > 0) {
            if (iter <= 7) flo = max(1u, total >> (11 - iter));
            else ones = true;
        }
        // rank sort of present symbols by (count', symbol)
        for (u32 i = 0; i < 8; i++) {
            u32 s = lane * 8 + i;
            u32 c = hist[s];
            if (!c) continue;
            u32 cs = ones ? 1u : max(c, flo);
            u32 rank = 0;
            for (u32 t = 0; t < 256; t++) {
                u32 ct = hist[t];
                if (!ct) continue;
                u32 cts = ones ? 1u : max(ct, flo);
                rank += (cts < cs || (cts == cs && t < s)) ? 1u : 0u;
            }
            H.ssym[rank] = (u8)s;
            H.node_w[rank] = cs;
        }
        __syncwarp();
        if (lane == 0) {
            u32 li = 0, ni = m, ne = m;
            for (u32 k = 0; k + 1 < m; k++) {
                u32 a, b;
                if (li < m && (ni >= ne || H.node_w[li] <= H.node_w[ni])) a = li++; else a = ni++;
                if (li < m && (ni >= ne || H.node_w[li] <= H.node_w[ni])) b = li++; else b = ni++;
                H.node_w[ne] = H.node_w[a] + H.node_w[b];
                H.node_par[a] = (u16)ne;
                H.node_par[b] = (u16)ne;
                ne++;
            }
            u32 root = ne - 1;
            H.node_dep[root] = 0;
            u32 md = 0;
            for (int i = (int)root - 1; i >= 0; i--) {
                u32 d = H.node_dep[H.node_par[i]] + 1u;
                H.node_dep[i] = (u8)min(d, 255u);
                if ((u32)i < m && d > md) md = d;
            }
            H.node_w[511] = md;
        }
        __syncwarp();
        maxBits = H.node_w[511];
        __syncwarp();
        if (maxBits <= HUF_MAXBITS) break;
    }
    // per-symbol lengths
    for (u32 i = 0; i < 8; i++) H.len[lane * 8 + i] = 0;
    __syncwarp();
    for (u32 r = lane; r < m; r += 32) H.len[H.ssym[r]] = H.node_dep[r];
    __syncwarp();
    // canonical codes, zstd order: weight 1 (longest codes) first, symbols ascending inside a weight
    // rankStart[w] (in cells of the 2^maxBits table) = sum_{w'<w} count[w'] << (w'-1)
    u32 *cnt_len = H.node_w;  // reuse: counts per code length 1..11
    if (lane < 16) cnt_len[lane] = 0;
    __syncwarp();
    for (u32 i = 0; i < 8; i++) {
        u32 l = H.len[lane * 8 + i];
        if (l) atomicAdd(&cnt_len[l], 1u);
    }
    __syncwarp();
    if (lane == 0) {
        u32 cells = 0;
        for (u32 w = 1; w <= maxBits; w++) {  // weight w <-> length maxBits+1-w
            u32 l = maxBits + 1 - w;
            u32 c = cnt_len[l];
            cnt_len[16 + l] = cells >> (w - 1);  // first code value of this length
            cells += c << (w - 1);
        }
    }
    __syncwarp();
    for (u32 i = 0; i < 8; i++) {
        u32 s = lane * 8 + i;
        u32 l = H.len[s];
        u32 e = 0;
        if (l) {
            u32 idx = 0;
            for (u32 t = 0; t < s; t++) idx += (H.len[t] == l) ? 1u : 0u;
            e = ((cnt_len[16 + l] + idx) << 4) | l;
        }
        hlut[s] = (u16)e;
    }
    __syncwarp();
    return maxBits;
}

// Huffman tree description (RFC 8878 §4.2.1) written by lane 0 into `out`; returns its size, or 0
// if it cannot be represented (then the caller stores the literals raw).
__device__ static u32 huf_write_tree(u8 *out, HufTmp &H, u32 maxBits, u32 maxSym, short *norm, u8 *tsym) {
    // weights of symbols 0..maxSym-1 (the last one is implied)
    u32 n = maxSym;
    u32 wc[13];
    for (u32 i = 0; i < 13; i++) wc[i] = 0;
    u32 maxW = 0;
    for (u32 s = 0; s < n; s++) {
        u32 l = H.len[s];
        u32 w = l ? maxBits + 1 - l : 0;
        wc[w]++;
        if (w > maxW) maxW = w;
    }
    u32 fse_size = 0;
    u32 mostc = 0;
    for (u32 i = 0; i <= maxW; i++) mostc = max(mostc, wc[i]);
    if (n > 2 && mostc != n && mostc > 1) {
        u32 tlog = fse_optimal_log(6, n, maxW);
        fse_normalize(wc, maxW, n, tlog, norm);
        u32 hs = fse_write_ncount(out + 1, norm, maxW, tlog);
        fse_build_ctable(norm, maxW, tlog, H.wtab, H.wdnb, H.wdfs, tsym);
        BitW bw;
        bw.init(out + 1 + hs, 300);
        FseState s1, s2;
        int ip = (int)n;
#define WGT(idx) (H.len[(idx)] ? maxBits + 1 - H.len[(idx)] : 0u)
        if (n & 1) {
            ip--; s1.init(H.wtab, H.wdnb, H.wdfs, tlog, WGT(ip));
            ip--; s2.init(H.wtab, H.wdnb, H.wdfs, tlog, WGT(ip));
            ip--; s1.encode(bw, WGT(ip));
        } else {
            ip--; s2.init(H.wtab, H.wdnb, H.wdfs, tlog, WGT(ip));
            ip--; s1.init(H.wtab, H.wdnb, H.wdfs, tlog, WGT(ip));
        }
        while (ip > 0) {
            ip--; s2.encode(bw, WGT(ip));
            ip--; s1.encode(bw, WGT(ip));
        }
        s2.flush(bw);
        s1.flush(bw);
        u32 bs = bw.close();
        fse_size = hs + bs;
        if (bw.ovf) fse_size = 0;
    }
    u32 raw_size = (n + 1) / 2;
    if (fse_size && fse_size < 128 && (fse_size < raw_size || n > 128)) {
        out[0] = (u8)fse_size;
        return 1 + fse_size;
    }
    if (n > 128) return 0;
    out[0] = (u8)(127 + n);
    for (u32 i = 0; i < n; i += 2) {
        u32 w0 = WGT(i), w1 = (i + 1 < n) ? WGT(i + 1) : 0u;
        out[1 + i / 2] = (u8)((w0 << 4) | w1);
    }
#undef WGT
    return 1 + raw_size;
}

// OR `nbits` (<= 32) of v into a zeroed bit array at absolute bit offset `bit` from word base wb
__device__ __forceinline__ void or_bits(u32 *wb, u64 bit, u32 v, u32 nbits) {
    if (!nbits) return;
    u64 w = bit >> 5;
    u32 sh = (u32)(bit & 31);
    atomicOr(&wb[w], v << sh);
    if (sh + nbits > 32) atomicOr(&wb[w + 1], v >> (32 - sh));
}

// Histogram of n bytes into hist[256] (zeroed here).  Byte value 0 — by far the most frequent
// symbol of delta-coded qualities and of the length / N streams — is counted in registers.
__device__ static void warp_histogram(const u8 *p, u32 n, u32 *hist) {
    u32 lane = lane_id();
    for (u32 i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    u32 zeros = 0;
    u32 head = (u32)((4u - ((uintptr_t)p & 3u)) & 3u);
    if (head > n) head = n;
    if (lane < head) {
        u32 b = p[lane];
        if (b) atomicAdd(&hist[b], 1u); else zeros++;
    }
    u32 nw = (n - head) >> 2;
    const u32 *w = (const u32 *)(p + head);
    for (u32 i = lane; i < nw; i += 32) {
        u32 x = w[i];
        u32 z = __vcmpeq4(x, 0u);
        zeros += (u32)__popc(z & 0x01010101u);
        if (z != 0xFFFFFFFFu) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 b = (x >> (8 * k)) & 0xFFu;
                if (b) atomicAdd(&hist[b], 1u);
            }
        }
    }
    u32 t0 = head + 4 * nw;
    if (lane < n - t0) {
        u32 b = p[t0 + lane];
        if (b) atomicAdd(&hist[b], 1u); else zeros++;
    }
    zeros = __reduce_add_sync(FULL, zeros);
    __syncwarp();
    if (lane == 0) hist[0] = zeros;
    __syncwarp();
}

// Huffman-encode lit[0..n) as 1 or 4 streams directly after the literals header/tree/jump table.
// Layout decisions and sizes were fixed by huf_plan(); this writes the stream bits.
// `wb` = 4-byte aligned base of the frame slot, stream k occupies bytes [sbyte[k], sbyte[k]+ssize[k]).
struct HufPlan {
    u32 nstreams;      // 1 or 4
    u32 seg;           // symbols per stream (first three)
    u32 sbits[4];      // payload bits per stream (without end mark)
    u32 ssize[4];      // bytes per stream
    u32 total;         // sum of ssize
};
// per-lane chunk of a stream: 8 lanes per stream
__device__ __forceinline__ void huf_lane_chunk(const HufPlan &P, u32 n, u32 lane, u32 &s0, u32 &s1, u32 &k, u32 &j) {
    k = lane >> 3;
    j = lane & 7;
    u32 a = 0, b = 0;
    if (k < P.nstreams) {
        a = k * P.seg;
        b = (k == P.nstreams - 1) ? n : min(n, a + P.seg);
        a = min(a, n);
    }
    u32 m = b - a, per = (m + 7) >> 3;
    s0 = min(b, a + j * per);
    s1 = min(b, s0 + per);
}
__device__ static void warp_huf_plan(const u8 *lit, u32 n, const u16 *hlut, HufPlan &P, u32 &mybits) {
    u32 lane = lane_id();
    P.nstreams = (n < 256) ? 1u : 4u;
    P.seg = (P.nstreams == 1) ? n : (n + 3) >> 2;
    u32 s0, s1, k, j;
    huf_lane_chunk(P, n, lane, s0, s1, k, j);
    u32 bits = 0;
    for (u32 i = s0; i < s1; i++) bits += hlut[lit[i]] & 15u;
    mybits = bits;
    u32 tot = group_sum(bits, group_mask(8), 8);
    P.total = 0;
    for (u32 q = 0; q < 4; q++) {
        u32 t = __shfl_sync(FULL, tot, (int)(q * 8));
        P.sbits[q] = (q < P.nstreams) ? t : 0;
        P.ssize[q] = (q < P.nstreams) ? (t + 1 + 7) >> 3 : 0;
        P.total += P.ssize[q];
    }
}
__device__ static void warp_huf_encode(const u8 *lit, u32 n, const u16 *hlut, const HufPlan &P, u32 mybits, u8 *dst /* first stream byte */) {
    u32 lane = lane_id();
    // zero the stream bytes (boundary words are merged with atomicOr)
    {
        u32 head = (u32)((4u - ((uintptr_t)dst & 3u)) & 3u);
        if (head > P.total) head = P.total;
        if (lane < head) dst[lane] = 0;
        u32 nw = (P.total - head) >> 2;
        u32 *zw = (u32 *)(dst + head);
        for (u32 i = lane; i < nw; i += 32) zw[i] = 0;
        u32 t0 = head + 4 * nw;
        if (lane < P.total - t0) dst[t0 + lane] = 0;
    }
    __syncwarp();
    u32 s0, s1, k, j;
    huf_lane_chunk(P, n, lane, s0, s1, k, j);
    // bit offset of this lane's chunk inside its stream: the LAST symbol is written first, so the
    // chunks of higher lanes come first
    u32 gm = group_mask(8);
    u32 incl = group_incl_scan(mybits, gm, 8);
    u32 tot = __shfl_sync(gm, incl, 7, 8);
    u32 after = tot - incl;  // bits of chunks j+1..7
    u32 sbyte = 0;
    for (u32 q = 0; q < k && q < 4; q++) sbyte += P.ssize[q];
    if (k < P.nstreams) {
        u32 *wb = (u32 *)((uintptr_t)dst & ~(uintptr_t)3);
        u64 bit = (u64)(((uintptr_t)dst & 3u) + sbyte) * 8u + after;
        u64 acc = 0;
        u32 nb = 0;
        for (u32 i = s1; i > s0; i--) {
            u32 e = hlut[lit[i - 1]];
            acc |= (u64)(e >> 4) << nb;
            nb += e & 15u;
            if (nb >= 32) {
                or_bits(wb, bit, (u32)acc, 32);
                bit += 32;
                acc >>= 32;
                nb -= 32;
            }
        }
        if (j == 0) {  // first chunk of the stream = last bits: append the end mark
            acc |= 1ull << nb;
            nb++;
        }
        if (nb > 32) {
            or_bits(wb, bit, (u32)acc, 32);
            bit += 32;
            acc >>= 32;
            nb -= 32;
        }
        or_bits(wb, bit, (u32)acc, nb);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------- literals section
// Writes the literals section for lit[0..n) at out; returns its size (always succeeds: raw fallback).
template <class WS>
__device__ static u32 warp_write_literals(const u8 *lit, u32 n, u8 *out, WS &S) {
    u32 lane = lane_id();
    u32 type = 0;  // 0 raw, 1 rle, 2 huffman
    u32 maxBits = 0, maxSym = 0, treeSize = 0;
    HufPlan P;
    u32 mybits = 0;
    P.nstreams = 1;
    P.total = 0;
    if (n >= 32) {
        warp_histogram(lit, n, S.hist);
        maxBits = warp_huf_build(S.hist, n, S.hlut, S.u.e.huf, &maxSym);
        if (maxBits == 0) {
            type = 1;  // a single distinct byte
        } else {
            warp_huf_plan(lit, n, S.hlut, P, mybits);
            // tree description goes right after the (3..5 byte) literals header
            u32 lh = 3 + (n >= 1024 ? 1 : 0) + (n >= 16384 ? 1 : 0);
            if (lane == 0) treeSize = huf_write_tree(out + lh, S.u.e.huf, maxBits, maxSym, S.norm, S.tmpsym);
            treeSize = __shfl_sync(FULL, treeSize, 0);
            u32 csize = treeSize + (P.nstreams == 4 ? 6u : 0u) + P.total;
            if (treeSize != 0 && csize + lh < n + 3 && csize < n) type = 2;
        }
    }
    if (type == 2) {
        u32 lh = 3 + (n >= 1024 ? 1 : 0) + (n >= 16384 ? 1 : 0);
        u32 csize = treeSize + (P.nstreams == 4 ? 6u : 0u) + P.total;
        if (lane == 0) {
            if (lh == 3) {
                u32 h = 2u | ((P.nstreams == 4 ? 1u : 0u) << 2) | (n << 4) | (csize << 14);
                out[0] = (u8)h; out[1] = (u8)(h >> 8); out[2] = (u8)(h >> 16);
            } else if (lh == 4) {
                u32 h = 2u | (2u << 2) | (n << 4) | (csize << 18);
                out[0] = (u8)h; out[1] = (u8)(h >> 8); out[2] = (u8)(h >> 16); out[3] = (u8)(h >> 24);
            } else {
                u32 h = 2u | (3u << 2) | (n << 4) | (csize << 22);
                out[0] = (u8)h; out[1] = (u8)(h >> 8); out[2] = (u8)(h >> 16); out[3] = (u8)(h >> 24);
                out[4] = (u8)(csize >> 10);
            }
            if (P.nstreams == 4) {
                u8 *jt = out + lh + treeSize;
                st_u16_unaligned(jt, P.ssize[0]);
                st_u16_unaligned(jt + 2, P.ssize[1]);
                st_u16_unaligned(jt + 4, P.ssize[2]);
            }
        }
        __syncwarp();
        warp_huf_encode(lit, n, S.hlut, P, mybits, out + lh + treeSize + (P.nstreams == 4 ? 6u : 0u));
        return lh + csize;
    }
    // raw / RLE literals
    u32 lh = 1 + (n > 31 ? 1 : 0) + (n > 4095 ? 1 : 0);
    if (lane == 0) {
        if (lh == 1) out[0] = (u8)(type | (n << 3));
        else if (lh == 2) {
            u32 h = type | (1u << 2) | (n << 4);
            out[0] = (u8)h; out[1] = (u8)(h >> 8);
        } else {
            u32 h = type | (3u << 2) | (n << 4);
            out[0] = (u8)h; out[1] = (u8)(h >> 8); out[2] = (u8)(h >> 16);
        }
        if (type == 1) out[lh] = lit[0];
    }
    if (type == 1) return lh + 1;
    for (u32 i = lane; i < n; i += 32) out[lh + i] = lit[i];
    return lh + n;
}

// ---------------------------------------------------------------------------------- sequences section
// seq arrays: ll[] (u16), ml[] (u16, matchLength-3), of[] (u32: raw offset on entry, offBase after).
// Returns section size; sets *ovf when the slot would overflow.
__device__ static u32 warp_write_sequences(u16 *sll, u16 *sml, u32 *sof, u32 nseq, u8 *out, u32 cap, WarpScratchLZ &S, bool *ovf) {
    u32 lane = lane_id();
    *ovf = false;
    if (nseq == 0) {
        if (lane == 0) out[0] = 0;
        return 1;
    }
    // 1. repeat-offset resolution (serial, RFC 8878 §3.1.1.5): offset -> offBase
    if (lane == 0) {
        u32 rep0 = 1, rep1 = 4, rep2 = 8;
        for (u32 i = 0; i < nseq; i++) {
            u32 off = sof[i];
            bool ll0 = (sll[i] == 0);
            u32 ob;
            if (!ll0) {
                if (off == rep0) ob = 1;
                else if (off == rep1) { ob = 2; rep1 = rep0; rep0 = off; }
                else if (off == rep2) { ob = 3; rep2 = rep1; rep1 = rep0; rep0 = off; }
                else { ob = off + 3; rep2 = rep1; rep1 = rep0; rep0 = off; }
            } else {
                if (off == rep1) { ob = 1; rep1 = rep0; rep0 = off; }
                else if (off == rep2) { ob = 2; rep2 = rep1; rep1 = rep0; rep0 = off; }
                else if (rep0 > 1 && off == rep0 - 1) { ob = 3; rep2 = rep1; rep1 = rep0; rep0 = off; }
                else { ob = off + 3; rep2 = rep1; rep1 = rep0; rep0 = off; }
            }
            sof[i] = ob;
        }
    }
    __syncwarp();
    // 2. code histograms (hist[0..63] LL, [64..127] OF, [128..191] ML)
    u32 *hLL = S.hist, *hOF = S.hist + 64, *hML = S.hist + 128;
    for (u32 i = lane; i < 192; i += 32) S.hist[i] = 0;
    __syncwarp();
    for (u32 i = lane; i < nseq; i += 32) {
        atomicAdd(&hLL[zstd_ll_code(sll[i])], 1u);
        atomicAdd(&hOF[hibit32(sof[i])], 1u);
        atomicAdd(&hML[zstd_ml_code(sml[i])], 1u);
    }
    __syncwarp();
    u32 total = 0;
    bool over = false;
    if (lane == 0) {
        u32 o = 0;
        if (nseq < 128) out[o++] = (u8)nseq;
        else if (nseq < 0x7F00) { out[o++] = (u8)((nseq >> 8) + 0x80); out[o++] = (u8)nseq; }
        else { out[o++] = 0xFF; out[o++] = (u8)(nseq - 0x7F00); out[o++] = (u8)((nseq - 0x7F00) >> 8); }
        u32 modes_at = o++;
        u32 mode[3], tlog[3];
        const u32 *hh[3] = {hLL, hOF, hML};
        const u32 maxLog[3] = {ZSTD_LL_MAXLOG, ZSTD_OF_MAXLOG, ZSTD_ML_MAXLOG};
        const u32 defLog[3] = {ZSTD_LL_DEFLOG, ZSTD_OF_DEFLOG, ZSTD_ML_DEFLOG};
        const u32 defMax[3] = {35, 28, 52};
        u32 rleSym[3] = {0, 0, 0};
        for (int t = 0; t < 3; t++) {
            u32 maxSym = 0, most = 0, mostSym = 0;
            for (u32 s = 0; s < 64; s++)
                if (hh[t][s]) {
                    maxSym = s;
                    if (hh[t][s] > most) { most = hh[t][s]; mostSym = s; }
                }
            // selection rule modelled on zstd's fast strategies: RLE if one symbol, predefined for
            // short or flat blocks, dynamic FSE otherwise
            u32 dynMin = ((1u << defLog[t]) * 9u) >> 3;
            u32 m;
            if (most == nseq) m = (nseq <= 2 && maxSym <= defMax[t]) ? 0u : 1u;
            else if (maxSym <= defMax[t] && (nseq < dynMin || most < (nseq >> (defLog[t] - 1)))) m = 0;
            else m = 2;
            mode[t] = m;
            FseCT &ct = S.u.e.fse[t];
            if (m == 1) {
                out[o++] = (u8)mostSym;
                rleSym[t] = mostSym;
                tlog[t] = 0;
            } else if (m == 0) {
                const short *dn = (t == 0) ? kLLDefNorm : (t == 1 ? kOFDefNorm : kMLDefNorm);
                for (u32 s = 0; s <= defMax[t]; s++) S.norm[s] = dn[s];
                tlog[t] = defLog[t];
                fse_build_ctable(S.norm, defMax[t], tlog[t], ct.tab, ct.dnb, ct.dfs, S.tmpsym);
            } else {
                tlog[t] = fse_optimal_log(maxLog[t], nseq, maxSym);
                fse_normalize(hh[t], maxSym, nseq, tlog[t], S.norm);
                if (o + 80 < cap) o += fse_write_ncount(out + o, S.norm, maxSym, tlog[t]);
                else over = true;
                fse_build_ctable(S.norm, maxSym, tlog[t], ct.tab, ct.dnb, ct.dfs, S.tmpsym);
            }
        }
        out[modes_at] = (u8)((mode[0] << 6) | (mode[1] << 4) | (mode[2] << 2));
        // 3. interleaved FSE bitstream, written backwards (last sequence first)
        if (!over) {
            BitW bw;
            bw.init(out + o, cap > o ? cap - o : 0);
            FseState sLL, sOF, sML;
            u32 i = nseq - 1;
            u32 llc = zstd_ll_code(sll[i]), ofc = hibit32(sof[i]), mlc = zstd_ml_code(sml[i]);
            sML.init(S.u.e.fse[2].tab, S.u.e.fse[2].dnb, S.u.e.fse[2].dfs, tlog[2], mlc);
            sOF.init(S.u.e.fse[1].tab, S.u.e.fse[1].dnb, S.u.e.fse[1].dfs, tlog[1], ofc);
            sLL.init(S.u.e.fse[0].tab, S.u.e.fse[0].dnb, S.u.e.fse[0].dfs, tlog[0], llc);
            bw.add(sll[i] & ((1u << kLLBits[llc]) - 1u), kLLBits[llc]);
            bw.add(sml[i] & ((1u << kMLBits[mlc]) - 1u), kMLBits[mlc]);
            bw.add(sof[i] & ((1u << ofc) - 1u), ofc);
            while (i-- > 0) {
                llc = zstd_ll_code(sll[i]);
                ofc = hibit32(sof[i]);
                mlc = zstd_ml_code(sml[i]);
                sOF.encode(bw, ofc);
                sML.encode(bw, mlc);
                sLL.encode(bw, llc);
                bw.add(sll[i] & ((1u << kLLBits[llc]) - 1u), kLLBits[llc]);
                bw.add(sml[i] & ((1u << kMLBits[mlc]) - 1u), kMLBits[mlc]);
                bw.add(sof[i] & ((1u << ofc) - 1u), ofc);
            }
            sML.flush(bw);
            sOF.flush(bw);
            sLL.flush(bw);
            u32 bs = bw.close();
            over = bw.ovf;
            total = o + bs;
        }
        (void)rleSym;
    }
    total = __shfl_sync(FULL, total, 0);
    *ovf = __shfl_sync(FULL, over ? 1 : 0, 0) != 0;
    return total;
}

// ---------------------------------------------------------------------------------- LZ77 parse (warp)
__device__ __forceinline__ u32 lz_hash(u32 v) { return (v * 2654435761u) >> (32 - HLOG); }

// Greedy parse of src[0..len): 32 consecutive positions are probed per step.  Emits sequences into
// sll/sml/sof and literals into lit.  Returns nseq; *nlit_out = literal bytes.
__device__ static u32 warp_lz_parse(const u8 *src, u32 len, u16 *htab, u8 *lit, u16 *sll, u16 *sml, u32 *sof, u32 *nlit_out) {
    u32 lane = lane_id();
    for (u32 i = lane; i < (1u << HLOG); i += 32) htab[i] = HEMPTY;
    __syncwarp();
    u32 anchor = 0, p = 0, nseq = 0, nlit = 0, lastoff = 0;
    while (p + 4 <= len) {
        u32 q = p + lane;
        bool valid = (q + 4 <= len);
        u32 v = valid ? ld_u32_unaligned(src + q) : 0u;
        u32 h = lz_hash(v);
        u32 cand = valid ? (u32)htab[h] : HEMPTY;
        u64 key = valid ? (u64)v : (0xFFFFFFFF00000000ull | lane);
        u32 peers = __match_any_sync(FULL, key);
        u32 lower = peers & ((1u << lane) - 1u);
        bool ok = false;
        if (lower) {
            cand = p + (31u - (u32)__clz((int)lower));
            ok = true;
        } else if (cand != HEMPTY && cand < q) {
            ok = (ld_u32_unaligned(src + cand) == v);
        }
        u32 ball = __ballot_sync(FULL, ok);
        if (ball == 0) {
            if (valid) htab[h] = (u16)q;
            __syncwarp();
            p += 32;
            continue;
        }
        int f = __ffs((int)ball) - 1;
        u32 ms = p + (u32)f;
        const u32 ms0 = ms;  // probe position, before any backward extension
        u32 c = __shfl_sync(FULL, cand, f);
        u32 mlen = 4;
        for (;;) {
            u32 o = mlen + 4u * lane;
            u32 x = 0xFFFFFFFFu;
            if (ms + o + 4 <= len) x = ld_u32_unaligned(src + ms + o) ^ ld_u32_unaligned(src + c + o);
            else if (ms + o < len) x = (ld_u32_unaligned(src + ms + o) ^ ld_u32_unaligned(src + c + o)) | (0xFFFFFFFFu << (8u * (len - ms - o)));
            u32 mm = __ballot_sync(FULL, x != 0u);
            if (mm) {
                int fl = __ffs((int)mm) - 1;
                u32 xx = __shfl_sync(FULL, x, fl);
                mlen += 4u * (u32)fl + (((u32)__ffs((int)xx) - 1u) >> 3);
                break;
            }
            mlen += 128;
        }
        while (ms > anchor && c > 0 && src[ms - 1] == src[c - 1]) {  // uniform backward extension
            ms--;
            c--;
            mlen++;
        }
        u32 off = ms - c;
        // a sequence costs roughly 3 bytes: short matches only pay off at a repeated / near offset
        bool take = (mlen >= 6) || (off == lastoff) || (mlen >= 5 && off < 256);
        if (!take) {  // skip the probe position and try again right after it
            u32 endp = min(ms0 + 1, len - 3);
            for (u32 q2 = p + lane; q2 < endp; q2 += 32) htab[lz_hash(ld_u32_unaligned(src + q2))] = (u16)q2;
            __syncwarp();
            p = ms0 + 1;
            continue;
        }
        u32 ll = ms - anchor;
        if (lane == 0) {
            sll[nseq] = (u16)ll;
            sml[nseq] = (u16)(mlen - 3);
            sof[nseq] = off;
        }
        for (u32 i = lane; i < ll; i += 32) lit[nlit + i] = src[anchor + i];
        nlit += ll;
        nseq++;
        lastoff = off;
        u32 endp = min(ms + mlen, len - 3);
        for (u32 q2 = p + lane; q2 < endp; q2 += 32) htab[lz_hash(ld_u32_unaligned(src + q2))] = (u16)q2;
        __syncwarp();
        anchor = p = ms + mlen;
    }
    for (u32 i = anchor + lane; i < len; i += 32) lit[nlit + i - anchor] = src[i];
    nlit += len - anchor;
    __syncwarp();
    *nlit_out = nlit;
    return nseq;
}

// ---------------------------------------------------------------------------------- frame writer
__device__ __forceinline__ void write_frame_header(u8 *o, u32 len) {
    o[0] = 0x28; o[1] = 0xB5; o[2] = 0x2F; o[3] = 0xFD;
    o[4] = 0x84;  // FCS 4 bytes, window descriptor present, content checksum, no dictionary
    o[5] = (u8)((FQZ_ZFRAME_LOG - 10) << 3);  // window = 64 KiB >= frame content
    o[6] = (u8)len; o[7] = (u8)(len >> 8); o[8] = (u8)(len >> 16); o[9] = (u8)(len >> 24);
}
__device__ __forceinline__ void write_block_header(u8 *o, u32 type, u32 size) {
    u32 h = 1u | (type << 1) | (size << 3);  // last block
    o[0] = (u8)h; o[1] = (u8)(h >> 8); o[2] = (u8)(h >> 16);
}

template <bool LZ>
__global__ void __launch_bounds__(ZENC_WARPS * 32) k_zenc(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots,
                                                          u8 *ws, u32 *out_sizes) {
    typedef typename std::conditional<LZ, WarpScratchLZ, WarpScratchEnt>::type WS;
    __shared__ WS scratch[ZENC_WARPS];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 wi = blockIdx.x * ZENC_WARPS + warp;
    if (wi >= nidx) return;
    u32 fi = index ? index[wi] : wi;
    WS &S = scratch[warp];
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    u8 *out = slots + fr.dst_off;
    u32 len = fr.src_len;
    u32 cap = (u32)FQZ_ZSLOT(len);
    if (lane == 0) write_frame_header(out, len);
    u8 *blk = out + 10;       // 3-byte block header, then content
    u8 *body = blk + 3;
    u32 body_cap = cap - 13 - 4;
    u32 bsize = 0, btype = 0;  // 0 raw, 1 rle, 2 compressed
    bool done = false;
    // RLE block: every byte equal
    {
        u32 b0 = src[0];
        bool same = true;
        for (u32 i = lane; i < len && same; i += 32) same = (src[i] == b0);
        if (__all_sync(FULL, same) && len > 1) {
            if (lane == 0) body[0] = (u8)b0;
            btype = 1;
            bsize = len;  // block header carries the regenerated size
            done = true;
        }
    }
    if (!done) {
        u32 total = 0;
        bool ovf = false;
        if (LZ && fr.policy == 0 && len >= 64) {
            if constexpr (LZ) {
                u8 *lit = ws + fr.ws_off;
                u32 maxseq = len / 4 + 1;
                u16 *sll = (u16 *)(lit + ((len + 15u) & ~15u));
                u16 *sml = sll + maxseq;
                u32 *sof = (u32 *)(sml + maxseq);  // 4*maxseq bytes after sll: 4-byte aligned
                u32 nlit = 0;
                u32 nseq = warp_lz_parse(src, len, S.u.htab, lit, sll, sml, sof, &nlit);
                __syncwarp();
                u32 lsz = warp_write_literals(nseq ? (const u8 *)lit : src, nlit, body, S);
                __syncwarp();
                u32 ssz = 0;
                if (lsz + 16 < body_cap) ssz = warp_write_sequences(sll, sml, sof, nseq, body + lsz, body_cap - lsz, S, &ovf);
                else ovf = true;
                total = lsz + ssz;
            }
        } else {
            u32 lsz = warp_write_literals(src, len, body, S);
            if (lane == 0) body[lsz] = 0;  // no sequences
            total = lsz + 1;
        }
        __syncwarp();
        if (!ovf && total < len) {
            btype = 2;
            bsize = total;
        } else {  // incompressible: raw block
            for (u32 i = lane; i < len; i += 32) body[i] = src[i];
            btype = 0;
            bsize = len;
        }
    }
    __syncwarp();
    if (lane == 0) {
        write_block_header(blk, btype, bsize);
        u32 payload = (btype == 1) ? 1u : bsize;
        u8 *ck = body + payload;
        u32 hsh = hashes[fi];
        ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
        out_sizes[fi] = 13 + payload + 4;
    }
}

// ---------------------------------------------------------------------------------- XXH64 (frame content checksum)
// XXH64 is four serial multiply-rotate lanes per frame and cannot be split inside a frame, so it is
// parallelised ACROSS frames: 4 threads per frame, 8 frames per warp.
__device__ __forceinline__ u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 xxh_round(u64 acc, u64 in) { return rotl64(acc + in * XXP2, 31) * XXP1; }
__device__ __forceinline__ u64 xxh_merge(u64 h, u64 v) { return (h ^ xxh_round(0, v)) * XXP1 + XXP4; }
__device__ __forceinline__ u64 ld_u64_unaligned(const u8 *p) { return (u64)ld_u32_unaligned(p) | ((u64)ld_u32_unaligned(p + 4) << 32); }

__device__ static u64 xxh64_quad(const u8 *p, u32 len, u32 q /*0..3*/, u32 gmask) {
    u64 acc = (q == 0) ? XXP1 + XXP2 : (q == 1) ? XXP2 : (q == 2) ? 0ull : 0ull - XXP1;
    u32 nstripes = len >> 5;
    const u8 *s = p + 8u * q;
    {
        // the rounds are a serial multiply-rotate chain, the loads are not: keep 8 stripes in flight
        // (three aligned words cover any 8 unaligned bytes, merged with funnel shifts)
        const u32 *w = (const u32 *)((uintptr_t)s & ~(uintptr_t)3);
        const u32 sh = (u32)((uintptr_t)s & 3u) * 8u;
        u32 i = 0;
        for (; i + 8 <= nstripes; i += 8) {
            u32 a[8], b[8], c[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                a[k] = w[8 * (i + k)];
                b[k] = w[8 * (i + k) + 1];
                c[k] = w[8 * (i + k) + 2];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                u64 v = (u64)__funnelshift_r(a[k], b[k], sh) | ((u64)__funnelshift_r(b[k], c[k], sh) << 32);
                acc = xxh_round(acc, v);
            }
        }
        for (; i < nstripes; i++) acc = xxh_round(acc, ld_u64_unaligned(s + 32ull * i));
    }
    u64 a0 = __shfl_sync(gmask, acc, 0, 4), a1 = __shfl_sync(gmask, acc, 1, 4), a2 = __shfl_sync(gmask, acc, 2, 4), a3 = __shfl_sync(gmask, acc, 3, 4);
    u64 h;
    if (len >= 32) {
        h = rotl64(a0, 1) + rotl64(a1, 7) + rotl64(a2, 12) + rotl64(a3, 18);
        h = xxh_merge(h, a0);
        h = xxh_merge(h, a1);
        h = xxh_merge(h, a2);
        h = xxh_merge(h, a3);
    } else
        h = XXP5;
    h += (u64)len;
    const u8 *t = p + 32ull * nstripes;
    u32 rem = len & 31;
    while (rem >= 8) {
        h ^= xxh_round(0, ld_u64_unaligned(t));
        h = rotl64(h, 27) * XXP1 + XXP4;
        t += 8;
        rem -= 8;
    }
    if (rem >= 4) {
        h ^= (u64)ld_u32_unaligned(t) * XXP1;
        h = rotl64(h, 23) * XXP2 + XXP3;
        t += 4;
        rem -= 4;
    }
    while (rem) {
        h ^= (u64)(*t) * XXP5;
        h = rotl64(h, 11) * XXP1;
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
__global__ void __launch_bounds__(128) k_xxh64_frames(const ZFrame *frames, u32 nframes, u32 *hashes) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 fi = t >> 2, q = t & 3;
    u32 gmask = group_mask(4);
    bool live = fi < nframes;
    u32 fj = live ? fi : nframes - 1;  // keep whole quads alive for the shuffles
    ZFrame fr = frames[fj];
    u64 h = xxh64_quad((const u8 *)(uintptr_t)fr.src, fr.src_len, q, gmask);
    if (live && q == 0) hashes[fi] = (u32)h;
}

void fqz_launch_xxh64(const ZFrame *frames, u32 nframes, u32 *hashes, cudaStream_t s) {
    if (!nframes) return;
    u32 threads = 128, grid = (nframes * 4 + threads - 1) / threads;
    FQZ_LAUNCH(k_xxh64_frames, grid, threads, 0, s, frames, nframes, hashes);
}
void fqz_launch_zenc(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u8 *ws, u32 *out_sizes, int lz,
                     cudaStream_t s) {
    if (!nidx) return;
    u32 grid = (nidx + ZENC_WARPS - 1) / ZENC_WARPS;
    if (lz) FQZ_LAUNCH(k_zenc<true>, grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, ws, out_sizes);
    else FQZ_LAUNCH(k_zenc<false>, grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, ws, out_sizes);
}
