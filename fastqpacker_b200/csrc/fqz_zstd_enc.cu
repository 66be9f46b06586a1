// fqz_zstd_enc.cu — GPU zstd (RFC 8878) frame encoder: the entropy stage of the block codec.
// Replaces zstd.Encoder.EncodeAll (reference call sites internal/compress/compress.go:523-528,
// options :115-118: SpeedFastest, checksum on, no dictionary).  The reference's encoder is the
// third-party klauspost/compress v1.19.1 (go.mod:8), absent from the tree; compressed bytes are not
// pinned by any reference test, so the contract is: valid RFC 8878 frames with content checksums
// that the reference's decoder accepts, ratio within 2 % of the CPU path.
//
// Three kernels, thousands of frames in flight per launch:
//   k_zenc_huf : literals-only frames (packed bases, qualities): one CTA per 128 KiB frame of eight
//                16 KiB blocks sharing one Huffman tree; RLE / raw fallbacks
//   k_zenc<2>  : item streams (headers, plus lines, N positions, lengths): one warp per 16 KiB frame:
//                item matcher, Huffman literals, FSE sequences (predefined / RLE / dynamic)
//   k_zenc<1>  : generic data (fqz_zstd_compress, policy AUTO): one warp per 64 KiB frame, greedy
//                LZ77 (32 positions per step: ballot + match_any candidate search, hash table in
//                shared memory), same entropy back end
//   k_zindex   : the skippable index frame in front of every stream of >= 4 frames (fqz_zstd.h)
#include <algorithm>
#include <type_traits>

#include "fqz_zstd.h"
#include "fqz_zstd_tables.cuh"
#include "fqz_xxh64.cuh"

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
    u8 tsym_b[512];    // second / third scratch so that the three sequence tables are built by three lanes at once
    short norm_b[128];
};
#define ZI_MAXM 8  // matches kept per item and lane (the rest of a pathological item stays literal)
struct WarpScratchItems {
    u32 hist[256];
    u16 hlut[256];
    union {
        u32 mbuf[32 * ZI_MAXM * 2];  // per lane: (pos | len << 16, offset)
        EntropyScratch e;
    } u;
    u8 tmpsym[512];
    short norm[64];
    u8 tsym_b[512];
    short norm_b[128];
};
struct WarpScratchEnt {
    u32 hist[256];
    u16 hlut[256];
    union {
        EntropyScratch e;
    } u;
    u8 tmpsym[512];
    short norm[64];
    u8 tsym_b[512];
    short norm_b[128];
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
        // rank sort of the present symbols by (count', symbol): the m keys count' << 8 | symbol are
        // compacted first (they are distinct), so ranking costs m compares per symbol instead of 256
        {
            u32 *keys = (u32 *)H.node_par;  // 256 words; node_par is only written by the merge below
            u32 at = group_incl_scan(present, FULL, 32) - present;
            for (u32 i = 0; i < 8; i++) {
                u32 s = lane * 8 + i;
                u32 c = hist[s];
                if (c) keys[at++] = ((ones ? 1u : max(c, flo)) << 8) | s;
            }
            __syncwarp();
            u32 myk[8], myr[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                u32 r = lane + 32u * j;
                myk[j] = (r < m) ? keys[r] : 0xFFFFFFFFu;
                myr[j] = 0;
            }
            const u32 nj = (m + 31u) >> 5;
            for (u32 t = 0; t < m; t++) {
                u32 kt = keys[t];
#pragma unroll
                for (u32 j = 0; j < 8; j++)
                    if (j < nj) myr[j] += (kt < myk[j]) ? 1u : 0u;
            }
            __syncwarp();  // every lane has read the keys: node_w / ssym may alias nothing, but keep the phases apart
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                u32 r = lane + 32u * j;
                if (r < m) {
                    H.ssym[myr[j]] = (u8)myk[j];
                    H.node_w[myr[j]] = myk[j] >> 8;
                }
            }
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
    {
        // index of a symbol among the symbols of its length, in symbol order: ranked over the compacted
        // (symbol-ordered) list of present symbols; the rank-sorted list in ssym is no longer needed
        u32 at = group_incl_scan(present, FULL, 32) - present;
        for (u32 i = 0; i < 8; i++) {
            u32 s = lane * 8 + i;
            hlut[s] = 0;
            if (hist[s]) H.ssym[at++] = (u8)s;
        }
        __syncwarp();
        for (u32 r = lane; r < m; r += 32) {
            u32 s = H.ssym[r], l = H.len[s], idx = 0;
            for (u32 t = 0; t < r; t++) idx += (H.len[H.ssym[t]] == l) ? 1u : 0u;
            hlut[s] = (u16)(((cnt_len[16 + l] + idx) << 4) | l);
        }
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
// offBase of one sequence given the repeat-offset history (RFC 8878 §3.1.1.5)
__device__ __forceinline__ u32 zstd_off_base(u32 off, bool ll0, u32 &rep0, u32 &rep1, u32 &rep2) {
    u32 ob;
#ifdef FQZ_REP0_ONLY
    ob = (!ll0 && off == rep0) ? 1u : off + 3u;
    rep0 = off;
    return ob;
#endif
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
    return ob;
}

// resolved: sof already holds offBase values (the item matcher resolves them in its clean-up pass)
template <class WS>
__device__ static u32 warp_write_sequences(u16 *sll, u16 *sml, u32 *sof, u32 nseq, u8 *out, u32 cap, WS &S, bool *ovf, bool resolved = false) {
    u32 lane = lane_id();
    *ovf = false;
    if (nseq == 0) {
        if (lane == 0) out[0] = 0;
        return 1;
    }
    // 1. repeat-offset resolution (serial, RFC 8878 §3.1.1.5): offset -> offBase
    if (lane == 0 && !resolved) {
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
    // the three tables (LL, OF, ML) are independent: lanes 0 / 1 / 2 choose the mode, normalise, build the
    // encoding table and write the table description of one each (into their scratch; lane 0 strings
    // the descriptions together afterwards)
    if (lane < 3) {
        const u32 t = lane;
        u8 *ts = (t == 0) ? S.tmpsym : (t == 1 ? (u8 *)S.hlut : S.tsym_b);  // the literals are written: the Huffman LUT is free
        short *nm = (t == 0) ? S.norm : S.norm_b + 64 * (t - 1);
        const u32 *h = S.hist + 64 * t;
        const u32 maxLog = (t == 0) ? ZSTD_LL_MAXLOG : (t == 1 ? ZSTD_OF_MAXLOG : ZSTD_ML_MAXLOG);
        const u32 defLog = (t == 0) ? ZSTD_LL_DEFLOG : (t == 1 ? ZSTD_OF_DEFLOG : ZSTD_ML_DEFLOG);
        const u32 defMax = (t == 0) ? 35u : (t == 1 ? 28u : 52u);
        u32 maxSym = 0, most = 0, mostSym = 0;
        for (u32 sy = 0; sy < 64; sy++) {
            u32 c = h[sy];
            if (c) {
                maxSym = sy;
                if (c > most) { most = c; mostSym = sy; }
            }
        }
        // selection rule modelled on zstd's fast strategies: RLE if one symbol, predefined for
        // short or flat blocks, dynamic FSE otherwise
        u32 dynMin = ((1u << defLog) * 9u) >> 3;
        u32 m;
        if (most == nseq) m = (nseq <= 2 && maxSym <= defMax) ? 0u : 1u;
        else if (maxSym <= defMax && (nseq < dynMin || most < (nseq >> (defLog - 1)))) m = 0;
        else m = 2;
        FseCT &ct = S.u.e.fse[t];
        u32 tl = 0, dlen = 0;
        if (m == 1) {
            ts[0] = (u8)mostSym;
            dlen = 1;
        } else if (m == 0) {
            const short *dn = (t == 0) ? kLLDefNorm : (t == 1 ? kOFDefNorm : kMLDefNorm);
            for (u32 sy = 0; sy <= defMax; sy++) nm[sy] = dn[sy];
            tl = defLog;
            fse_build_ctable(nm, defMax, tl, ct.tab, ct.dnb, ct.dfs, ts);
        } else {
            tl = fse_optimal_log(maxLog, nseq, maxSym);
            fse_normalize(h, maxSym, nseq, tl, nm);
            fse_build_ctable(nm, maxSym, tl, ct.tab, ct.dnb, ct.dfs, ts);
            dlen = fse_write_ncount(ts, nm, maxSym, tl);  // the spreading scratch is free again; a description is < 80 bytes
        }
        S.hist[194 + t] = tl;
        S.hist[200 + t] = m;
        S.hist[204 + t] = dlen;
    }
    __syncwarp();
    if (lane == 0) {
        u32 o = 0;
        if (nseq < 128) out[o++] = (u8)nseq;
        else if (nseq < 0x7F00) { out[o++] = (u8)((nseq >> 8) + 0x80); out[o++] = (u8)nseq; }
        else { out[o++] = 0xFF; out[o++] = (u8)(nseq - 0x7F00); out[o++] = (u8)((nseq - 0x7F00) >> 8); }
        out[o++] = (u8)((S.hist[200] << 6) | (S.hist[201] << 4) | (S.hist[202] << 2));
        for (u32 t = 0; t < 3; t++) {
            const u8 *ts = (t == 0) ? S.tmpsym : (t == 1 ? (const u8 *)S.hlut : S.tsym_b);
            u32 dlen = S.hist[204 + t];
            if (o + 80 < cap) {
                for (u32 k = 0; k < dlen; k++) out[o + k] = ts[k];
                o += dlen;
            } else
                over = true;
        }
        S.hist[192] = o;
        S.hist[193] = over ? 1u : 0u;
    }
    __syncwarp();
    // 3. interleaved FSE bitstream, written backwards (last sequence first), 32 sequences per round:
    //    A. the three state chains are the only serial part: lanes 0 / 1 / 2 walk the LL / OF / ML
    //       chain of the round in lockstep and leave (value, nbBits) of every state flush in shared
    //       memory — no bit writer on the chain;
    //    B. every lane then assembles the whole bit packet of ONE sequence (three state flushes,
    //       three extra-bit fields, <= 90 bits), a warp scan of the packet lengths gives its place,
    //       and the packets are OR-ed into a shared-memory stage that is flushed to the (aligned)
    //       output words with coalesced stores.
    {
        u32 o = S.hist[192];
        over = S.hist[193] != 0;
        const u32 tl0 = S.hist[194], tl1 = S.hist[195], tl2 = S.hist[196];
        u32 *codes = (u32 *)S.tmpsym;       // 32 x (llc | ofc << 8 | mlc << 16): the table builders are done with tmpsym
        u16 *sb = (u16 *)(codes + 32);      // [3][32] state flushes: value | nbBits << 12
        u32 *bst = (u32 *)S.hlut;           // 128-word bit stage (the literals are written: the Huffman LUT is free)
        u8 *gout = out + o;
        const u32 pre = (u32)((uintptr_t)gout & 3u) * 8u;  // bits of the aligned word in front of the bitstream
        u32 *gw = (u32 *)((uintptr_t)gout & ~(uintptr_t)3);
        const u32 capb = cap > o ? cap - o : 0u;
        for (u32 i = lane; i < 128; i += 32) bst[i] = 0;
        __syncwarp();
        if (lane == 0 && pre) bst[0] = gw[0] & ((1u << pre) - 1u);  // table descriptions written above by this warp
        __syncwarp();
        u32 carry = pre, wpos = 0;
        // chain state of lane t < 3 (0 LL, 1 OF, 2 ML)
        const u32 ct = min(lane, 2u);
        const u16 *ctab = S.u.e.fse[ct].tab;
        const u32 *cdnb = S.u.e.fse[ct].dnb;
        const int *cdfs = S.u.e.fse[ct].dfs;
        const u32 ctl = (ct == 0) ? tl0 : (ct == 1 ? tl1 : tl2);
        u32 st = 0;
        bool started = false;
        if (!over) {
            for (u32 cend = nseq; cend > 0;) {
                const u32 cs = cend >= 32 ? cend - 32 : 0, cn = cend - cs;
                const bool live = lane < cn;  // lane j codes sequence cend-1-j: lane order = emission order
                u32 vll = 0, vml = 0, vof = 1, llc = 0, ofc = 0, mlc = 0;
                if (live) {
                    u32 k = cend - 1u - lane;
                    vll = sll[k];
                    vml = sml[k];
                    vof = sof[k];
                    llc = zstd_ll_code(vll);
                    ofc = hibit32(vof);
                    mlc = zstd_ml_code(vml);
                    codes[lane] = llc | (ofc << 8) | (mlc << 16);
                }
                __syncwarp();
                if (lane < 3) {
                    if (ctl) {
                        u32 j = 0;
                        if (!started) {  // the last sequence only initialises the states
                            u32 sym = (codes[0] >> (8u * lane)) & 0xFFu;
                            u32 nbo = (cdnb[sym] + (1u << 15)) >> 16;
                            u32 v = (nbo << 16) - cdnb[sym];
                            st = ctab[(int)(v >> nbo) + cdfs[sym]];
                            sb[lane * 32] = 0;
                            j = 1;
                        }
                        for (; j < cn; j++) {
                            u32 sym = (codes[j] >> (8u * lane)) & 0xFFu;
                            u32 nbo = (st + cdnb[sym]) >> 16;
                            sb[lane * 32 + j] = (u16)((st & ((1u << nbo) - 1u)) | (nbo << 12));
                            st = ctab[(int)(st >> nbo) + cdfs[sym]];
                        }
                    } else
                        for (u32 j = 0; j < cn; j++) sb[lane * 32 + j] = 0;  // RLE table: the state never emits bits
                }
                started = true;
                __syncwarp();
                u64 lo = 0;
                u32 hi = 0, nb = 0;
                if (live) {
                    u32 fLL = sb[lane], fOF = sb[32 + lane], fML = sb[64 + lane];
                    lo = fOF & 0xFFFu;
                    nb = fOF >> 12;
                    lo |= (u64)(fML & 0xFFFu) << nb;
                    nb += fML >> 12;
                    lo |= (u64)(fLL & 0xFFFu) << nb;
                    nb += fLL >> 12;
                    u32 lb = kLLBits[llc], mb = kMLBits[mlc];
                    lo |= (u64)(vll & ((1u << lb) - 1u)) << nb;
                    nb += lb;
                    lo |= (u64)(vml & ((1u << mb) - 1u)) << nb;
                    nb += mb;  // <= 59
                    u32 ofx = vof & ((1u << ofc) - 1u);
                    lo |= (u64)ofx << nb;
                    if (nb + ofc > 64) hi = ofx >> (64u - nb);
                    nb += ofc;
                }
                u32 incl = group_incl_scan(nb, FULL, 32);
                u32 cbits = __shfl_sync(FULL, incl, 31);
                if ((u64)wpos * 4u + ((carry + cbits + 7u) >> 3) + 16u > (u64)capb + (pre >> 3)) {
                    over = true;
                    break;
                }
                if (nb) {
                    u32 off = carry + incl - nb, w = off >> 5, sh = off & 31u;
                    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32);
                    u32 x0 = l0 << sh, x1 = __funnelshift_l(l0, l1, sh), x2 = __funnelshift_l(l1, hi, sh), x3 = __funnelshift_l(hi, 0u, sh);
                    if (x0) atomicOr(&bst[w], x0);
                    if (x1) atomicOr(&bst[w + 1], x1);
                    if (x2) atomicOr(&bst[w + 2], x2);
                    if (x3) atomicOr(&bst[w + 3], x3);
                }
                __syncwarp();
                u32 T = carry + cbits, nw = T >> 5;
                for (u32 i = lane; i < nw; i += 32) gw[wpos + i] = bst[i];
                u32 lastw = bst[nw];
                __syncwarp();
                for (u32 i = lane; i <= nw + 3u; i += 32) bst[i] = 0;
                __syncwarp();
                if (lane == 0) bst[0] = lastw;
                __syncwarp();
                wpos += nw;
                carry = T & 31u;
                cend = cs;
            }
            if (!over) {
                // final state flush (ML, OF, LL) and the end mark
                u32 sLLv = __shfl_sync(FULL, st, 0), sOFv = __shfl_sync(FULL, st, 1), sMLv = __shfl_sync(FULL, st, 2);
                u32 pk = 0, nb = 0;
                if (tl2) { pk |= (sMLv & ((1u << tl2) - 1u)) << nb; nb += tl2; }
                if (tl1) { pk |= (sOFv & ((1u << tl1) - 1u)) << nb; nb += tl1; }
                if (tl0) { pk |= (sLLv & ((1u << tl0) - 1u)) << nb; nb += tl0; }
                pk |= 1u << nb;
                nb += 1;  // <= 27
                if (lane == 0) {
                    bst[0] |= pk << carry;
                    if (carry + nb > 32) bst[1] = pk >> (32u - carry);
                }
                __syncwarp();
                u32 T = carry + nb, nby = (T + 7u) >> 3;
                u8 *tail = (u8 *)(gw + wpos);
                for (u32 i = lane; i < nby; i += 32) tail[i] = (u8)(bst[i >> 2] >> (8u * (i & 3u)));
                u32 bs = wpos * 4u + nby - (pre >> 3);
                if (bs > capb) over = true;
                total = o + bs;
            }
        }
        __syncwarp();
    }
    over = __any_sync(FULL, over);
    *ovf = over;
    return __shfl_sync(FULL, total, 0);
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

// ---------------------------------------------------------------------------------- item matcher (warp)
// The four structured streams are chains of items (u16 length + bytes per header / plus line,
// u16 count + positions per read, u32 per length) and consecutive items resemble each other.
// Instead of a hash-table search, every lane takes one item and tries two candidate offsets:
//   d1 = length of the previous item  (item r aligned with item r-1 at their starts)
//   d2 = length of this item          (aligned at their ends: what follows a variable-width field)
// A match may run up to ZI_AHEAD bytes past its item, so that equal neighbours fuse into one long
// match in the serial clean-up pass.  All 32 lanes work on 32 consecutive items at a time.
#define ZR_MAXDIST ((8u << 20) - 1u)  // record matcher: matches stay inside the 8 MiB window every zstd decoder must support
#define ZR_MINMATCH 8u
#define ZR_MINRUN 12u
#define ZI_AHEAD 32u
#define ZI_MINMATCH 3u

// number of leading positions q in [p, lim) with S[q] == S[q - d]
__device__ __forceinline__ u32 zi_match_len(const u8 *S, u32 p, u32 d, u32 lim) {
    u32 l = 0;
    while (p + l < lim) {
        u32 x = ld_u32_unaligned(S + p + l) ^ ld_u32_unaligned(S + p + l - d);
        u32 rem = lim - (p + l);
        if (x == 0) {
            if (rem <= 4) return l + rem;
            l += 4;
        } else {
            u32 k = ((u32)__ffs((int)x) - 1u) >> 3;
            return l + min(k, rem);
        }
    }
    return l;
}

// ---- equality masks: the whole item is compared with its two candidate offsets up front — one pass of
//      independent loads (every source word is loaded once and reused by both offsets, a rolling funnel shift
//      serves the unaligned positions) instead of a compare loop whose every step waits for its own loads — and
//      the matches are then read off 128-bit masks with bit scans.  Items of more than ZM_SPAN positions
//      (with the look-ahead) keep the byte loop.
#define ZM_SPAN 128u
#ifdef FQZ_EMU
#define ZM_DEBUG_TOGGLE &&!getenv("FQZ_NOMASK")
#else
#define ZM_DEBUG_TOGGLE
#endif
struct Mask128 {
    u64 lo, hi;
};
__device__ __forceinline__ Mask128 m128_shr(Mask128 m, u32 k) {  // 0 < k < 64
    Mask128 r;
    r.lo = (m.lo >> k) | (m.hi << (64u - k));
    r.hi = m.hi >> k;
    return r;
}
__device__ __forceinline__ Mask128 m128_and(Mask128 a, Mask128 b) {
    Mask128 r;
    r.lo = a.lo & b.lo;
    r.hi = a.hi & b.hi;
    return r;
}
// positions from which at least `k` ones follow (k = 3, 8 or 12)
__device__ __forceinline__ Mask128 m128_runs(Mask128 m, u32 k) {
    Mask128 m2 = m128_and(m, m128_shr(m, 1));
    if (k == 3u) return m128_and(m2, m128_shr(m, 2));
    Mask128 m4 = m128_and(m2, m128_shr(m2, 2));
    Mask128 m8 = m128_and(m4, m128_shr(m4, 4));
    if (k == 8u) return m8;
    return m128_and(m8, m128_shr(m4, 8));  // 12
}
// number of ones from bit q upwards
__device__ __forceinline__ u32 m128_runlen(Mask128 m, u32 q) {
    if (q < 64u) {
        u64 x = ~(m.lo >> q);  // the q bits shifted in are zeros -> ones: the run ends at 64 - q at the latest
        u32 r = x ? (u32)__ffsll((long long)x) - 1u : 64u;
        if (r < 64u - q) return r;
        u64 y = ~m.hi;
        return (64u - q) + (y ? (u32)__ffsll((long long)y) - 1u : 64u);
    }
    u64 x = ~(m.hi >> (q - 64u));
    return x ? (u32)__ffsll((long long)x) - 1u : 64u;
}
// first set bit at or above q, ZM_SPAN if none
__device__ __forceinline__ u32 m128_next(Mask128 m, u32 q) {
    if (q < 64u) {
        u64 x = m.lo & (~0ull << q);
        if (x) return (u32)__ffsll((long long)x) - 1u;
        return m.hi ? 64u + (u32)__ffsll((long long)m.hi) - 1u : ZM_SPAN;
    }
    u64 x = m.hi & (~0ull << (q - 64u));
    return x ? 64u + (u32)__ffsll((long long)x) - 1u : ZM_SPAN;
}
// four flag bits of the byte lanes in which a and b agree
__device__ __forceinline__ u32 eq_bits4(u32 a, u32 b) { return (((__vcmpeq4(a, b) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }
// masks of S[p0 + i] == S[p0 + i - d] for i < n (n <= ZM_SPAN), d = d1 / d2 (0 = no such offset)
__device__ __forceinline__ void zi_eq_masks(const u8 *S, u32 p0, u32 n, u32 d1, u32 d2, Mask128 &A, Mask128 &B) {
    const u8 *c = S + p0, *a = c - d1, *b = c - d2;
    const u32 *cw = (const u32 *)((uintptr_t)c & ~(uintptr_t)3), *aw = (const u32 *)((uintptr_t)a & ~(uintptr_t)3),
              *bw = (const u32 *)((uintptr_t)b & ~(uintptr_t)3);
    const u32 cs = ((u32)(uintptr_t)c & 3u) * 8u, as = ((u32)(uintptr_t)a & 3u) * 8u, bs = ((u32)(uintptr_t)b & 3u) * 8u;
    u32 cp = cw[0], ap = d1 ? aw[0] : 0u, bp = d2 ? bw[0] : 0u;
    A.lo = A.hi = B.lo = B.hi = 0;
    const u32 nw = (n + 3u) >> 2;
#pragma unroll 4
    for (u32 j = 0; j < nw; j++) {
        u32 cn = cw[j + 1];
        u32 cv = __funnelshift_r(cp, cn, cs);
        cp = cn;
        u32 ea = 0, eb = 0;
        if (d1) {
            u32 an = aw[j + 1];
            ea = eq_bits4(cv, __funnelshift_r(ap, an, as));
            ap = an;
        }
        if (d2) {
            u32 bn = bw[j + 1];
            eb = eq_bits4(cv, __funnelshift_r(bp, bn, bs));
            bp = bn;
        }
        const u32 sh = (4u * j) & 63u;
        if (j < 16u) {
            A.lo |= (u64)ea << sh;
            B.lo |= (u64)eb << sh;
        } else {
            A.hi |= (u64)ea << sh;
            B.hi |= (u64)eb << sh;
        }
    }
    // positions at or beyond n do not count
    if (n < 64u) {
        const u64 keep = (1ull << n) - 1ull;
        A.lo &= keep; B.lo &= keep;
        A.hi = B.hi = 0;
    } else if (n < 128u) {
        const u64 keep = (n == 64u) ? 0ull : (1ull << (n - 64u)) - 1ull;
        A.hi &= keep; B.hi &= keep;
    }
}

// Parse of src[0..len) with item boundaries.  items != nullptr: items[i] - item_base is the frame
// position of item i (monotonic, nitems entries, the last one an end sentinel); items == nullptr:
// fixed stride `nitems` bytes.  Emits sequences into sll/sml/sof (raw offsets) and literals into lit.
// back: bytes of the same zstd frame in front of src (earlier blocks of a multi-block frame) that
// matches may reach into; cand: optional, per item 1 + the number of an earlier item with the same
// leading bytes (k_rec_match), whose offset is tried before the predecessor's;
// rep_known: false for a block that does not open its frame — the decoder's repeat-offset history is
// then whatever the block before left, which this warp does not know: it starts from three values no
// offset can equal, so repeat codes are only used for offsets this block has itself emitted.
__device__ static u32 warp_item_parse(const u8 *src, u32 len, const u32 *items, u32 item_base, u32 nitems, u32 *mbuf, u8 *lit, u16 *sll,
                                      u16 *sml, u32 *sof, u16 *slo, u32 *nlit_out, u32 back = 0, const u32 *cand = nullptr,
                                      bool rep_known = true) {
    u32 lane = lane_id();
    // first item that overlaps the frame
    long long i0 = 0, iend = 0;  // items [i0, iend) overlap [0, len)
    if (items) {
        // largest i with items[i] <= item_base (binary search, uniform)
        u32 lo = 0, hi = nitems - 1;  // invariant: items[lo] <= item_base (items[0] is the stream start)
        while (lo < hi) {
            u32 mid = (lo + hi + 1) >> 1;
            if (items[mid] <= item_base) lo = mid; else hi = mid - 1;
        }
        i0 = lo;
        iend = (long long)nitems - 1;  // the sentinel is not an item
    } else {
        i0 = 0;  // frames of fixed-stride streams start on an item boundary
        iend = ((long long)len + nitems - 1) / nitems;
    }
    u32 nm = 0;  // matches found so far (uncompacted list in sll = pos, sml = len, sof = offset)
    bool have_tail = false;  // last entry of the list, kept in registers so that the next round can extend it
    u32 tail_pos = 0, tail_end = 0, tail_off = 0, tail_idx = 0;
    for (long long ib = i0; ib < iend; ib += 32) {
        long long i = ib + lane;
        long long s = 0, e = 0;
        u32 d1 = 0, d2 = 0;
        bool live = i < iend;
        if (live) {
            if (items) {
                s = (long long)items[i] - item_base;
                e = (long long)items[i + 1] - item_base;
                d2 = (u32)(e - s);
                d1 = (i > 0) ? (u32)(items[i] - items[i - 1]) : 0u;
                if (cand) {
                    u32 ci = cand[i];  // 1 + number of the partner item
                    u32 dh = ci ? items[i] - items[ci - 1u] : 0u;
                    if (dh) d1 = dh;
                    d2 = 1;  // runs of one byte (equal qualities are zeros after the delta, poly-A / poly-G tails)
                }
            } else {
                s = i * nitems;
                e = s + nitems;
                d1 = d2 = nitems;
            }
            if (s >= (long long)len) live = false;
        }
        // uniform exit: item starts are monotonic, so once lane 0 is past the frame all lanes are
        if (__all_sync(FULL, !live)) break;
        u32 cnt = 0;
        if (live) {
            u32 p = (u32)max(s, 0ll), stop = (u32)min(e, (long long)len);
            u32 lim = (u32)min((long long)len, e + (long long)ZI_AHEAD);
            const u32 dmax = cand ? ZR_MAXDIST : 65535u;  // inside the frame's window
            if (d1 > dmax) d1 = 0;
            if (d2 > dmax || d2 == d1) d2 = 0;
            // record streams: unlike the item streams their literals are cheap (2 bits per base, a bit per zero),
            // so only matches that clearly beat them are taken
            const u32 min1 = cand ? ZR_MINMATCH : ZI_MINMATCH, min2 = cand ? ZR_MINRUN : ZI_MINMATCH;
            // (an offset that reaches in front of the frame at the item's start — the first items of a frame — takes the byte loop)
            // items of up to four spans go through the masks span by span (a match that runs into the next span
            // is found again there with the same offset and fused by the clean-up)
            if (lim - p <= 4u * ZM_SPAN && (!d1 || p + back >= d1) && (!d2 || p + back >= d2) ZM_DEBUG_TOGGLE) {
                const u32 e1 = d1, e2 = d2;
                if (e1 | e2) {
                    for (u32 cb = p; cb < stop && cnt < ZI_MAXM; cb += ZM_SPAN) {
                        const u32 n = min(ZM_SPAN, lim - cb);
                        Mask128 A, B;
                        zi_eq_masks(src, cb, n, e1, e2, A, B);
                        const Mask128 RA = m128_runs(A, min1), RB = m128_runs(B, min2);
                        Mask128 C;
                        C.lo = RA.lo | RB.lo;
                        C.hi = RA.hi | RB.hi;
                        const u32 qstop = min(stop - cb, n);
                        u32 q = 0;
                        while (cnt < ZI_MAXM) {
                            q = m128_next(C, q);
                            if (q >= qstop) break;
                            const bool ina = q < 64u ? (RA.lo >> q) & 1ull : (RA.hi >> (q - 64u)) & 1ull;
                            const bool inb = q < 64u ? (RB.lo >> q) & 1ull : (RB.hi >> (q - 64u)) & 1ull;
                            const u32 la = ina ? m128_runlen(A, q) : 0u, lb = inb ? m128_runlen(B, q) : 0u;
                            const u32 best = max(la, lb);
                            mbuf[(lane * ZI_MAXM + cnt) * 2] = (cb + q) | (best << 16);
                            mbuf[(lane * ZI_MAXM + cnt) * 2 + 1] = (la >= lb) ? e1 : e2;
                            cnt++;
                            q += best;
                            if (q >= ZM_SPAN) break;
                        }
                    }
                }
            } else
            while (p < stop && cnt < ZI_MAXM) {
                u32 la = (d1 && p + back >= d1) ? zi_match_len(src, p, d1, lim) : 0u;
                u32 lb = (d2 && p + back >= d2) ? zi_match_len(src, p, d2, lim) : 0u;
                if (la < min1) la = 0;
                if (lb < min2) lb = 0;
                u32 best = max(la, lb);
                if (best >= ZI_MINMATCH) {
                    mbuf[(lane * ZI_MAXM + cnt) * 2] = p | (best << 16);
                    mbuf[(lane * ZI_MAXM + cnt) * 2 + 1] = (la >= lb) ? d1 : d2;
                    cnt++;
                    p += best;
                } else
                    p += max(best, 1u);
            }
        }
        // ---- fuse chains inside the round: the first match of an item is absorbed by the last match
        //      of the item before it (or by the tail of the previous round) when both use the same
        //      offset and touch; runs of fully matched items (the all-zero plus / N streams, equal
        //      lengths) collapse into one entry instead of one per item
        u32 fpos = 0, fend = 0, foff = 0, lpos = 0, lend = 0, loff = 0;
        if (cnt) {
            u32 v = mbuf[(lane * ZI_MAXM) * 2];
            fpos = v & 0xFFFFu;
            fend = fpos + (v >> 16);
            foff = mbuf[(lane * ZI_MAXM) * 2 + 1];
            u32 w = mbuf[(lane * ZI_MAXM + cnt - 1) * 2];
            lpos = w & 0xFFFFu;
            lend = lpos + (w >> 16);
            loff = mbuf[(lane * ZI_MAXM + cnt - 1) * 2 + 1];
        }
        u32 pend_ = __shfl_up_sync(FULL, lend, 1), poff = __shfl_up_sync(FULL, loff, 1), pcnt = __shfl_up_sync(FULL, cnt, 1);
        if (lane == 0) {
            pend_ = tail_end;
            poff = tail_off;
            pcnt = have_tail ? 1u : 0u;
        }
        bool absorbed = cnt > 0 && pcnt > 0 && foff == poff && fpos <= pend_;
        u32 A = __ballot_sync(FULL, absorbed), S1 = __ballot_sync(FULL, cnt == 1);
        // chain that starts at lane h (h = -1: the tail): absorbed lanes h+1 .. h+t
        auto chain_len = [&](int h) -> u32 {
            if (h >= 31) return 0u;
            u32 a = A >> (h + 1), s1 = S1 >> (h + 1);
            u32 run = (u32)__ffs((int)~a) - 1u;   // consecutive absorbed lanes (ffs(0) = 0 -> 0xFFFFFFFF only if a == ~0)
            if (~a == 0u) run = 32u;
            u32 c1 = (~s1 == 0u) ? 32u : (u32)__ffs((int)~s1) - 1u;  // of which single-match items
            u32 t = (run <= c1) ? run : c1 + 1u;
            return min(t, 31u - (u32)h);
        };
        bool head = cnt > 0 && !(absorbed && cnt == 1);
        u32 t_me = head ? chain_len((int)lane) : 0u;
        u32 src_lane = min(lane + t_me, 31u);
        u32 ext_end = __shfl_sync(FULL, fend, (int)src_lane);
        if (t_me) lend = max(lend, ext_end);
        // the tail of the previous round may grow (lane 0 speaks for it)
        u32 t_tail = have_tail ? chain_len(-1) : 0u;
        u32 tail_new_end = __shfl_sync(FULL, fend, (int)(t_tail ? t_tail - 1u : 0u));
        if (t_tail) {
            tail_end = max(tail_end, tail_new_end);
            if (lane == 0) sml[tail_idx] = (u16)(tail_end - tail_pos);
        }
        u32 nwr = cnt - (absorbed ? 1u : 0u);
        u32 incl = group_incl_scan(nwr, FULL, 32);
        u32 tot = __shfl_sync(FULL, incl, 31);
        u32 at = nm + incl - nwr;
        for (u32 k = (absorbed ? 1u : 0u), o = 0; k < cnt; k++, o++) {
            u32 v = mbuf[(lane * ZI_MAXM + k) * 2];
            u32 pp = v & 0xFFFFu, ln = v >> 16;
            if (k == cnt - 1) ln = lend - lpos;  // possibly extended over the absorbed neighbours
            sll[at + o] = (u16)pp;
            sml[at + o] = (u16)ln;  // < 65536: a match cannot start at 0
            sof[at + o] = mbuf[(lane * ZI_MAXM + k) * 2 + 1];
        }
        // new tail = last entry written in this round
        u32 W = __ballot_sync(FULL, nwr > 0);
        if (W) {
            int lw = 31 - __clz((int)W);
            tail_pos = __shfl_sync(FULL, lpos, lw);
            tail_end = __shfl_sync(FULL, lend, lw);
            tail_off = __shfl_sync(FULL, loff, lw);
            tail_idx = __shfl_sync(FULL, at + nwr - 1u, lw);
            have_tail = true;
        }
        nm += tot;
        __syncwarp();
    }
    __syncwarp();
    // Clean-up: fuse / trim what the rounds left overlapping, turn positions into literal lengths,
    // resolve repeat offsets.  The fuse / trim state machine looks serial, but its state — the pending
    // match — is re-established by the first entry that starts a new match, so the list is cut into 32
    // segments that the lanes process in lockstep: pass A finds every segment's final pending match
    // from a cold start, pass B reruns the segment with the previous segment's pending as input and
    // PROVES that its own final state is what pass A promised the next lane (any mismatch — tiny lists,
    // runs that swallow a whole segment — falls back to the serial pass), pass C turns the survivors into
    // sequences with warp scans.  Only the repeat-offset history stays serial.
    u32 nseq = 0, nlit = 0;
    bool par_done = false;
    if (nm >= 256u && 8u * (nm + 1u) <= len) {
        u32 *toff = (u32 *)lit;  // the literal buffer is free until the literals are gathered below
        u16 *tpos = (u16 *)(toff + nm + 1u), *tlen = tpos + nm + 1u;
        const u32 seg = (nm + 31u) >> 5, ea = min(nm, lane * seg), eb = min(nm, ea + seg);
        // ---- pass A
        u32 kpos = 0, klen = 0, koff = 0;
        bool have = false;
        for (u32 e = ea; e < eb; e++) {
            u32 p = sll[e], l = sml[e], o = sof[e];
            if (have && o == koff && p <= kpos + klen) {
                klen = max(kpos + klen, p + l) - kpos;
                continue;
            }
            if (have && p < kpos + klen) {
                u32 cut = kpos + klen - p;
                if (l < cut + ZI_MINMATCH) continue;
                p += cut;
                l -= cut;
            }
            kpos = p; klen = l; koff = o; have = true;
        }
        const u32 a_pos = kpos, a_len = klen, a_off = koff;
        const bool a_have = have;
        // ---- pass B: incoming pending = what pass A left in the segment before
        {
            u32 ipos = __shfl_up_sync(FULL, a_pos, 1), ilen = __shfl_up_sync(FULL, a_len, 1), ioff = __shfl_up_sync(FULL, a_off, 1);
            bool ihave = __shfl_up_sync(FULL, a_have ? 1u : 0u, 1) != 0;
            if (lane == 0) ihave = false;
            kpos = ipos; klen = ilen; koff = ioff; have = ihave;
        }
        u32 w = ea;
        for (u32 e = ea; e < eb; e++) {
            u32 p = sll[e], l = sml[e], o = sof[e];
            if (have && o == koff && p <= kpos + klen) {
                klen = max(kpos + klen, p + l) - kpos;
                continue;
            }
            if (have && p < kpos + klen) {
                u32 cut = kpos + klen - p;
                if (l < cut + ZI_MINMATCH) continue;
                p += cut;
                l -= cut;
            }
            if (have) {
                tpos[w] = (u16)kpos; tlen[w] = (u16)klen; toff[w] = koff;
                w++;
            }
            kpos = p; klen = l; koff = o; have = true;
        }
        const bool last_seg = (ea < nm) && (eb == nm);
        bool bad = (ea < nm) && !last_seg && !(have == a_have && kpos == a_pos && klen == a_len && koff == a_off);
        if (last_seg && have) {  // nobody takes the last pending over
            tpos[w] = (u16)kpos; tlen[w] = (u16)klen; toff[w] = koff;
            w++;
        }
        if (!__any_sync(FULL, bad)) {
            // ---- pass C: sequence numbers, literal lengths and literal offsets
            const u32 cnt = w - ea;
            u32 cincl = group_incl_scan(cnt, FULL, 32);
            nseq = __shfl_sync(FULL, cincl, 31);
            const u32 g0 = cincl - cnt;
            u32 lastend = cnt ? (u32)tpos[w - 1] + tlen[w - 1] : 0u;
            for (int d = 1; d < 32; d <<= 1) {  // ends grow with the position in the list: prefix maximum
                u32 t = __shfl_up_sync(FULL, lastend, (unsigned)d);
                if ((int)lane >= d) lastend = max(lastend, t);
            }
            u32 prev0 = __shfl_up_sync(FULL, lastend, 1);
            if (lane == 0) prev0 = 0;
            u32 sumll = 0;
            {
                u32 pe = prev0;
                for (u32 k = ea; k < w; k++) {
                    sumll += (u32)tpos[k] - pe;
                    pe = (u32)tpos[k] + tlen[k];
                }
            }
            u32 lincl = group_incl_scan(sumll, FULL, 32);
            nlit = __shfl_sync(FULL, lincl, 31);
            u32 lo = lincl - sumll;
            __syncwarp();  // every lane has read the uncompacted entries of its segment; the arrays may be overwritten
            {
                u32 pe = prev0;
                for (u32 k = ea, g = g0; k < w; k++, g++) {
                    u32 p = tpos[k], l = tlen[k];
                    sll[g] = (u16)(p - pe);
                    sml[g] = (u16)(l - 3u);
                    sof[g] = toff[k];
                    slo[g] = (u16)lo;
                    lo += p - pe;
                    pe = p + l;
                }
            }
            const u32 pend_all = __shfl_sync(FULL, lastend, 31);
            __syncwarp();
            // ---- repeat offsets, 32 sequences per step.  The decoder's history is a move-to-front list as long as no
            //      offset is pushed that it already holds on top, so its first two entries are known without walking it:
            //      rep0 = the previous sequence's offset, rep1 = the offset of the run of equal offsets in front of
            //      that one.  A sequence that continues its run codes rep0 (its literal length is > 0: touching
            //      matches of one offset were fused above), the head of a run codes rep1 when it returns to the
            //      offset before the previous run (the start- / end-aligned offsets of an item alternate), else
            //      the offset itself; rep2 is not used.  The one case that would push a duplicate (ll == 0 with
            //      the previous offset) sends the list through the serial walk below instead.
            {
                u32 c0 = rep_known ? 1u : 0xFFFFFFF1u, c1 = rep_known ? 4u : 0xFFFFFFE1u;
                bool dup_push = false;
                for (u32 base = 0; base < nseq; base += 32) {
                    const u32 i = base + lane;
                    const bool live = i < nseq;
                    const u32 off = live ? sof[i] : 0u, ll = live ? sll[i] : 1u;
                    u32 prev = __shfl_up_sync(FULL, off, 1);
                    if (lane == 0) prev = c0;
                    const bool head = live && off != prev;
                    const u32 H = __ballot_sync(FULL, head), Lm = __ballot_sync(FULL, live);
                    // start of the run in front of this lane's run: the highest head below this lane
                    const u32 P = H & ((1u << lane) - 1u);
                    const int pl = P ? 31 - __clz((int)P) : -1;
                    u32 before = __shfl_sync(FULL, off, pl > 0 ? pl - 1 : 0);  // offset of the run in front of that one
                    if (pl < 0) before = c1;
                    else if (pl == 0) before = c0;
                    u32 ob;
                    if (!head) {
                        ob = 1u;
                        if (live && ll == 0) dup_push = true;
                    } else if (off == before) ob = ll ? 2u : 1u;
                    else ob = off + 3u;
                    if (live) toff[i] = ob;  // the scratch of the clean-up is free again; sof keeps the raw offsets for the fallback
                    // carry: last offset of the step and the offset of the run in front of its last run
                    const int last = 31 - __clz((int)Lm);
                    const int ql = H ? 31 - __clz((int)H) : -1;
                    u32 n1 = __shfl_sync(FULL, off, ql > 0 ? ql - 1 : 0);
                    if (ql < 0) n1 = c1;
                    else if (ql == 0) n1 = c0;
                    c1 = n1;
                    c0 = __shfl_sync(FULL, off, last);
                }
                __syncwarp();
                if (!__any_sync(FULL, dup_push)) {
                    for (u32 i = lane; i < nseq; i += 32) sof[i] = toff[i];
                } else {  // exact walk of the history (lane 0), staged through shared memory
                    u32 *inb = mbuf, *outb = mbuf + 64;
                    u32 rep0 = rep_known ? 1u : 0xFFFFFFF1u, rep1 = rep_known ? 4u : 0xFFFFFFE1u, rep2 = rep_known ? 8u : 0xFFFFFFD1u;
                    for (u32 base = 0; base < nseq; base += 32) {
                        u32 cn = min(32u, nseq - base);
                        __syncwarp();
                        if (lane < cn) {
                            inb[lane] = sof[base + lane];
                            inb[32 + lane] = sll[base + lane];
                        }
                        __syncwarp();
                        if (lane == 0)
                            for (u32 k = 0; k < cn; k++) outb[k] = zstd_off_base(inb[k], inb[32 + k] == 0, rep0, rep1, rep2);
                        __syncwarp();
                        if (lane < cn) sof[base + lane] = outb[lane];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbuf[0] = nseq;
                mbuf[1] = nlit;
                mbuf[2] = pend_all;
            }
            par_done = true;
        }
        __syncwarp();
    }
    if (!par_done) {
    // serial pass (lane 0): the list is read and written through shared-memory chunks of 32 entries so
    // that the serial chain never waits for global memory
    nseq = 0;
    nlit = 0;
    {
        u32 *inb = mbuf;             // 32 x (pos | len << 16), 32 x offset
        u32 *outb = mbuf + 64;       // 64 x (ll | (ml-3) << 16), 64 x offBase, 64 x literal offset
        u32 pend = 0;                // end of the last flushed match
        u32 kpos = 0, klen = 0, koff = 0;
        bool have = false;
        u32 rep0 = rep_known ? 1u : 0xFFFFFFF1u, rep1 = rep_known ? 4u : 0xFFFFFFE1u, rep2 = rep_known ? 8u : 0xFFFFFFD1u;
        for (u32 base = 0; base <= nm; base += 32) {  // one extra iteration flushes the pending match
            u32 cn = min(32u, nm - min(nm, base));
            __syncwarp();
            if (lane < cn) {
                inb[lane] = (u32)sll[base + lane] | ((u32)sml[base + lane] << 16);
                inb[32 + lane] = sof[base + lane];
            }
            __syncwarp();
            u32 produced = 0;
            if (lane == 0) {
                bool final_round = base + 32 > nm;
                for (u32 k = 0; k < cn + (final_round ? 1u : 0u); k++) {
                    bool last = (k == cn);
                    u32 p = 0, l = 0, o = 0;
                    if (!last) {
                        p = inb[k] & 0xFFFFu;
                        l = inb[k] >> 16;
                        o = inb[32 + k];
                        if (have && o == koff && p <= kpos + klen) {  // same offset, touching or overlapping: one match
                            klen = max(kpos + klen, p + l) - kpos;
                            continue;
                        }
                        if (have && p < kpos + klen) {  // overlap with a different offset: keep what is left
                            u32 cut = kpos + klen - p;
                            if (l < cut + ZI_MINMATCH) continue;
                            p += cut;
                            l -= cut;
                        }
                    }
                    if (have) {  // flush the pending match
                        u32 ll = kpos - pend;
                        u32 ob = zstd_off_base(koff, ll == 0, rep0, rep1, rep2);
                        outb[produced] = ll | ((klen - 3u) << 16);
                        outb[64 + produced] = ob;
                        outb[128 + produced] = nlit;
                        produced++;
                        nlit += ll;
                        pend = kpos + klen;
                    }
                    if (!last) {
                        kpos = p;
                        klen = l;
                        koff = o;
                        have = true;
                    } else
                        have = false;
                }
                outb[192] = produced;
                outb[193] = nlit;
                outb[194] = pend;
            }
            __syncwarp();
            produced = outb[192];
            // outputs never overtake inputs: entry nseq + j was read in this or an earlier chunk
            for (u32 q = lane; q < produced; q += 32) {
                sll[nseq + q] = (u16)outb[q];
                sml[nseq + q] = (u16)(outb[q] >> 16);
                sof[nseq + q] = outb[64 + q];
                slo[nseq + q] = (u16)outb[128 + q];
            }
            nseq += produced;
        }
        __syncwarp();
        if (lane == 0) {
            mbuf[0] = nseq;
            mbuf[1] = outb[193];
            mbuf[2] = outb[194];
        }
    }
    }
    __syncwarp();
    nseq = mbuf[0];
    nlit = mbuf[1];
    u32 pend = mbuf[2];
#ifdef FQZ_EMU
    if (lane == 0 && getenv("FQZ_DEBUG")) fprintf(stderr, "item_parse len %u nm %u nseq %u nlit(before tail) %u pend %u items %p stride/count %u\n", len, nm, nseq, nlit, pend, (const void *)items, nitems);
#endif
    __syncwarp();
    // literals: the bytes in front of every match, then the tail
    {
        u32 mstart = 0;  // running sum of ll + ml is not needed: the source of sequence i ends where its match starts
        (void)mstart;
        // position of match i = (sum of ll and ml of sequences < i) + ll_i; recomputed with a warp scan per 32 sequences
        u32 base = 0;
        for (u32 i0s = 0; i0s < nseq; i0s += 32) {
            u32 i = i0s + lane;
            u32 ll = 0, ml = 0;
            if (i < nseq) {
                ll = sll[i];
                ml = (u32)sml[i] + 3u;
            }
            u32 inc = group_incl_scan(ll + ml, FULL, 32);
            u32 startpos = base + inc - (ll + ml);  // where this sequence's literals begin
            if (i < nseq) {
                u32 dsto = slo[i];
                for (u32 k = 0; k < ll; k++) lit[dsto + k] = src[startpos + k];
            }
            base += __shfl_sync(FULL, inc, 31);
        }
    }
    for (u32 i = pend + lane; i < len; i += 32) lit[nlit + (i - pend)] = src[i];
    nlit += len - pend;
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

// MODE 1: hash-table LZ77, 2: item matcher (MODE 0, a literals-only variant, is no longer instantiated: k_zenc_huf and the
// three-kernel coder took its place)
// The item matcher runs as its own kernel in front of k_zenc<2> (which then only entropy-codes what it
// finds in the workspace): two kernels of half the code each instead of one whose 240 KB of instructions
// thrashed the instruction cache (15 % of its stalls were instruction fetch).
__global__ void __launch_bounds__(ZENC_WARPS * 32) k_zitems_parse(const ZFrame *frames, const u32 *index, u32 nidx, u8 *ws, u32 *parsed) {
    __shared__ u32 mbufs[ZENC_WARPS][32 * ZI_MAXM * 2];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 wi = blockIdx.x * ZENC_WARPS + warp;
    if (wi >= nidx) return;
    u32 fi = index ? index[wi] : wi;
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    u32 len = fr.src_len;
    u32 nseq = 0, nlit = 0;
    if (fr.policy != 1u /* FQZ_ZPOLICY_ENTROPY */ && len >= 64) {
        u8 *lit = ws + fr.ws_off;
        u32 maxseq = ((len / 2) + 2u) & ~1u;  // even: keeps sof 4-byte aligned (same layout as k_zenc<2>)
        u16 *sll = (u16 *)(lit + ((len + 15u) & ~15u));
        u16 *sml = sll + maxseq;
        u32 *sof = (u32 *)(sml + maxseq);
        u16 *slo = (u16 *)(sof + maxseq);
        nseq = warp_item_parse(src, len, (const u32 *)(uintptr_t)fr.items, fr.item_base, fr.item_count, mbufs[warp], lit, sll, sml, sof, slo, &nlit);
    }
    if (lane == 0) {
        parsed[2 * wi] = nseq;
        parsed[2 * wi + 1] = nlit;
    }
}

// PHASE 0: the whole frame in one kernel.  PHASE 1 / 2 (item frames): literals, then sequences and the
// frame's closing steps, as two kernels — `parsed` carries (sequences, literals) from the item matcher to
// phase 1 and (sequences, literals-section size | RLE mark) from phase 1 to phase 2.
#define ZENC_RLE_MARK 0xFFFFFFFFu
template <int MODE, int PHASE>
__global__ void __launch_bounds__(ZENC_WARPS * 32) k_zenc(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots,
                                                          u8 *ws, u32 *out_sizes, u32 *parsed) {
    constexpr bool LZ = MODE != 0;
    typedef typename std::conditional<MODE == 1, WarpScratchLZ, typename std::conditional<MODE == 2, WarpScratchItems, WarpScratchEnt>::type>::type WS;
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
    if (PHASE != 2 && lane == 0) write_frame_header(out, len);
    u8 *blk = out + 10;       // 3-byte block header, then content
    u8 *body = blk + 3;
    u32 body_cap = cap - 13 - 4;
    u32 bsize = 0, btype = 0;  // 0 raw, 1 rle, 2 compressed
    bool done = false;
    // RLE block: every byte equal
    if constexpr (PHASE != 2) {
        u32 b0 = src[0];
        bool same = true;
        for (u32 i = lane; i < len && same; i += 32) same = (src[i] == b0);
        if (__all_sync(FULL, same) && len > 1) {
            if (lane == 0) body[0] = (u8)b0;
            btype = 1;
            bsize = len;  // block header carries the regenerated size
            done = true;
        }
    } else if (parsed[2 * wi + 1] == ZENC_RLE_MARK) {
        btype = 1;
        bsize = len;
        done = true;
    }
    if constexpr (PHASE == 1) {
        if (done) {
            if (lane == 0) parsed[2 * wi + 1] = ZENC_RLE_MARK;
            return;
        }
    }
    if (!done) {
        u32 total = 0;
        bool ovf = false;
        if (LZ && fr.policy != 1u /* FQZ_ZPOLICY_ENTROPY */ && len >= 64) {
            if constexpr (LZ) {
                u8 *lit = ws + fr.ws_off;
                // item matcher: the uncompacted match list holds at most one entry per two bytes (items are >= 2 bytes)
                u32 maxseq = ((MODE == 2 ? len / 2 : len / 3) + 2u) & ~1u;  // even: keeps sof 4-byte aligned
                u16 *sll = (u16 *)(lit + ((len + 15u) & ~15u));
                u16 *sml = sll + maxseq;
                u32 *sof = (u32 *)(sml + maxseq);
                u32 nlit = 0;
                u32 nseq;
                if constexpr (MODE == 2) {  // parsed by k_zitems_parse
                    nseq = parsed[2 * wi];
                    nlit = parsed[2 * wi + 1];  // phase 2: the size of the literals section instead
                } else
                    nseq = warp_lz_parse(src, len, S.u.htab, lit, sll, sml, sof, &nlit);
                __syncwarp();
                u32 lsz;
                if constexpr (PHASE == 2) lsz = nlit;
                else lsz = warp_write_literals(nseq ? (const u8 *)lit : src, nlit, body, S);
                __syncwarp();
                if constexpr (PHASE == 1) {
                    if (lane == 0) parsed[2 * wi + 1] = lsz;
                    return;
                }
                u32 ssz = 0;
                if (lsz + 16 < body_cap) ssz = warp_write_sequences(sll, sml, sof, nseq, body + lsz, body_cap - lsz, S, &ovf, MODE == 2);
                else ovf = true;
                total = lsz + ssz;
            }
        } else {
            u32 lsz;
            if constexpr (PHASE == 2) lsz = parsed[2 * wi + 1];
            else lsz = warp_write_literals(src, len, body, S);
            if constexpr (PHASE == 1) {
                if (lane == 0) parsed[2 * wi + 1] = lsz;
                return;
            }
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

// ---------------------------------------------------------------------------------- literals-only frames: one CTA per frame
// Frames of the ENTROPY policy (packed bases, delta-coded qualities) hold up to eight 16 KiB blocks
// that share ONE Huffman tree: block 0 carries the tree description, blocks 1.. are "treeless"
// (RFC 8878 literals type 3).  One CTA (8 warps) codes one frame:
//   1. per-stream histograms (8 blocks x 4 streams, two u16 counters per word, coalesced 128-byte reads)
//   2. frame histogram -> length-limited Huffman code (CTA-parallel rank sort, serial merge)
//   3. exact size of every stream = <stream histogram, code lengths>  -> final layout, headers
//   4. every warp encodes the four streams of its block straight into their final position: 128
//      symbols per step, warp scan of the code lengths, bits OR-ed into a shared-memory stage that is
//      flushed with aligned 32-bit stores
// Decoders see ordinary multi-block frames; the GPU decoder decodes the 32 streams of a frame in parallel.
#define ZH_THREADS 256
#define ZH_WARPS 8
#define ZH_STAGE_WORDS 360  // a round of 32 lanes x ZH_CHUNK symbols x 11 bits, plus the carried partial word
#define ZH_CHUNK 32u        // symbols per lane and round
#define ZH_LANE_WORDS 12u
#define ZH_MIN_HUF 1024u  // blocks / frames below this are stored raw

struct ZhShared {
    u32 shist[32][128];  // per-stream histograms: symbol s in bits 16*(s&1).. of word s>>1 (<= 4096 symbols per stream)
    u32 hist[256];
    u32 keys[256];
    u16 hlut[256];  // code << 4 | nbBits
    HufTmp huf;
    u8 tmpsym[512];
    short norm[64];
    u8 tree[384];
    u32 sbits[32], sdst[32], ssize[32];
    u32 misc[8];  // 0 mode, 1 treeSize, 2 total bytes, 3 distinct, 4 maxSym, 5 maxBits
    u32 stage[ZH_WARPS][ZH_STAGE_WORDS];
};

// histogram of src[a, b) (one Huffman stream) by one warp.  The body is read as aligned 16-byte
// vectors, two per lane in flight, so that the counting never waits for a 4-byte load; zero — by far
// the most frequent byte of delta-coded qualities — is counted in a register.
__device__ __forceinline__ void hist_word(u32 x, u32 *h2, u32 &zeros) {
#pragma unroll
    for (u32 k = 0; k < 4; k++) {
        u32 sy = (x >> (8u * k)) & 0xFFu;
        if (sy == 0) zeros++;
        else atomicAdd(&h2[sy >> 1], 1u << (16u * (sy & 1u)));
    }
}
__device__ static void warp_stream_hist(const u8 *src, u32 a, u32 b, u32 *h2) {
    u32 lane = lane_id();
    u32 zeros = 0;
    // head: bytes in front of the first 16-byte boundary, tail: bytes behind the last one
    u32 a16 = a + (u32)((16u - ((uintptr_t)(src + a) & 15u)) & 15u);
    if (a16 > b) a16 = b;
    u32 nv = (b - a16) >> 4, b16 = a16 + 16u * nv;
    for (u32 p = a + lane; p < a16; p += 32) {
        u32 sy = src[p];
        if (sy == 0) zeros++;
        else atomicAdd(&h2[sy >> 1], 1u << (16u * (sy & 1u)));
    }
    const uint4 *v = (const uint4 *)(src + a16);
    u32 i = lane;
    for (; i + 32 < nv; i += 64) {
        uint4 x0 = v[i], x1 = v[i + 32];
        hist_word(x0.x, h2, zeros); hist_word(x0.y, h2, zeros); hist_word(x0.z, h2, zeros); hist_word(x0.w, h2, zeros);
        hist_word(x1.x, h2, zeros); hist_word(x1.y, h2, zeros); hist_word(x1.z, h2, zeros); hist_word(x1.w, h2, zeros);
    }
    if (i < nv) {
        uint4 x0 = v[i];
        hist_word(x0.x, h2, zeros); hist_word(x0.y, h2, zeros); hist_word(x0.z, h2, zeros); hist_word(x0.w, h2, zeros);
    }
    for (u32 p = b16 + lane; p < b; p += 32) {
        u32 sy = src[p];
        if (sy == 0) zeros++;
        else atomicAdd(&h2[sy >> 1], 1u << (16u * (sy & 1u)));
    }
    zeros = __reduce_add_sync(FULL, zeros);
    if (lane == 0 && zeros) atomicAdd(&h2[0], zeros);
}

// CTA-wide (256 threads) version of warp_huf_build: hist[256] -> H.len / hlut.  Returns maxBits
// (0 when fewer than two distinct symbols); misc[3] = distinct symbols, misc[4] = last present symbol.
__device__ static u32 cta_huf_build(const u32 *hist, u32 total, u16 *hlut, HufTmp &H, u32 *keys, u32 *misc) {
    u32 tid = threadIdx.x;
    u32 c = hist[tid];
    u32 m = (u32)__syncthreads_count(c != 0);
    if (tid == 0) misc[4] = 0;
    __syncthreads();
    if (c) atomicMax(&misc[4], tid);
    if (tid == 0) misc[3] = m;
    __syncthreads();
    if (m < 2) return 0;
    u32 maxBits = 0;
    for (int iter = 0;; iter++) {
        u32 flo = 0;
        bool ones = false;
        if (iter > 0) {
            if (iter <= 7) flo = max(1u, total >> (11 - iter));
            else ones = true;
        }
        u32 cs = c ? (ones ? 1u : max(c, flo)) : 0u;
        keys[tid] = cs;
        __syncthreads();
        if (c) {  // rank among the present symbols by (count', symbol)
            u32 rank = 0;
            for (u32 t = 0; t < 256; t++) {
                u32 kt = keys[t];
                rank += (kt != 0 && (kt < cs || (kt == cs && t < tid))) ? 1u : 0u;
            }
            H.ssym[rank] = (u8)tid;
            H.node_w[rank] = cs;
        }
        __syncthreads();
        if (tid == 0) {
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
            misc[5] = md;
        }
        __syncthreads();
        maxBits = misc[5];
        __syncthreads();
        if (maxBits <= HUF_MAXBITS) break;
    }
    H.len[tid] = 0;
    __syncthreads();
    if (tid < m) H.len[H.ssym[tid]] = H.node_dep[tid];
    u32 *cnt_len = H.node_w;  // reuse: [l] symbols of length l, [16 + l] first code value of length l
    __syncthreads();
    if (tid < 32) cnt_len[tid] = 0;
    __syncthreads();
    u32 l = H.len[tid];
    if (l) atomicAdd(&cnt_len[l], 1u);
    __syncthreads();
    if (tid == 0) {
        u32 cells = 0;
        for (u32 w = 1; w <= maxBits; w++) {  // weight w <-> length maxBits+1-w
            u32 ll = maxBits + 1 - w;
            u32 cc = cnt_len[ll];
            cnt_len[16 + ll] = cells >> (w - 1);
            cells += cc << (w - 1);
        }
    }
    __syncthreads();
    u32 e = 0;
    if (l) {
        u32 idx = 0;
        for (u32 t = 0; t < tid; t++) idx += (H.len[t] == l) ? 1u : 0u;
        e = ((cnt_len[16 + l] + idx) << 4) | l;
    }
    hlut[tid] = (u16)e;
    __syncthreads();
    return maxBits;
}

// One warp encodes the Huffman stream of src[a, b) to dst (exactly ((bits + 1) + 7) / 8 bytes).
// The LAST symbol goes to the lowest bits (the decoder reads the stream backwards).
// Rounds of ZH_CHUNK symbols per lane: every lane codes ITS run of symbols into a private bit buffer with a
// plain serial loop (table lookup, shift, or — no warp scan per handful of symbols), one warp scan of the lanes'
// bit counts then places the 32 buffers in the warp's stage (whole words by funnel shift, only the two words a
// lane shares with its neighbours by atomicOr), and the complete words go out with coalesced 32-bit stores.
// lbuf: ZH_LANE_WORDS x 32 words, word j of lane l at lbuf[32 j + l] (conflict free).
__device__ static void warp_stream_encode(const u8 *src, u32 a, u32 b, const u16 *hlut, u8 *dst, u32 *stage, u32 *lbuf) {
    u32 lane = lane_id();
    const u32 al = (u32)((uintptr_t)dst & 3u);  // stage byte i <-> global byte gbase[i]
    u8 *gbase = dst - al;
    for (u32 i = lane; i < ZH_STAGE_WORDS; i += 32) stage[i] = 0;
    __syncwarp();
    u32 P = 8u * al;  // bits in the stage (the first 8 al of the first word are not ours)
    bool first = true;
    const u32 m = b - a, T = (m + 32u * ZH_CHUNK - 1u) / (32u * ZH_CHUNK);
    for (u32 t = 0; t < T; t++) {
        // this lane's symbols: [lo, hi), coded from hi - 1 downwards (lane 0 owns the highest indices = lowest bits)
        const int top = (int)(b - 32u * ZH_CHUNK * t) - (int)(ZH_CHUNK * lane);
        const int hi = max(top, (int)a), lo = max(top - (int)ZH_CHUNK, (int)a);
        u64 acc = 0;
        u32 nb = 0, nwd = 0;
        if (hi > lo) {
            // whole words from the top through a rolling pair of aligned loads, then the bytes left at the bottom
            const u8 *pt = src + hi;
            const u32 *wp = (const u32 *)((uintptr_t)pt & ~(uintptr_t)3);
            const u32 sh = ((u32)(uintptr_t)pt & 3u) * 8u;
            u32 whi = sh ? wp[0] : 0u;  // bytes of the word that holds pt (only its low part is ours)
            int i = hi;
            for (; i - 4 >= lo; i -= 4) {
                wp--;
                const u32 wlo = wp[0];
                const u32 x = sh ? __funnelshift_r(wlo, whi, sh) : wlo;  // src[i-4 .. i-1]
                whi = wlo;
                // two symbols (<= 22 bits) on top of < 32 pending bits fit the 64-bit accumulator
#pragma unroll
                for (int k = 3; k >= 1; k -= 2) {
                    const u32 e1 = hlut[(x >> (8 * k)) & 0xFFu], e0 = hlut[(x >> (8 * (k - 1))) & 0xFFu];
                    const u32 l1 = e1 & 15u;
                    const u32 pair = (e1 >> 4) | ((e0 >> 4) << l1);
                    acc |= (u64)pair << nb;
                    nb += l1 + (e0 & 15u);
                    if (nb >= 32u) {
                        lbuf[32u * nwd + lane] = (u32)acc;
                        nwd++;
                        acc >>= 32;
                        nb -= 32u;
                    }
                }
            }
            for (; i > lo; i--) {  // at most three
                const u32 e = hlut[src[i - 1]];
                acc |= (u64)(e >> 4) << nb;
                nb += e & 15u;
                if (nb >= 32u) {
                    lbuf[32u * nwd + lane] = (u32)acc;
                    nwd++;
                    acc >>= 32;
                    nb -= 32u;
                }
            }
            if (nb) lbuf[32u * nwd + lane] = (u32)acc;
        }
        const u32 L = 32u * nwd + nb;  // bits of this lane
        const u32 incl = group_incl_scan(L, FULL, 32);
        const u32 tot = __shfl_sync(FULL, incl, 31);
        if (L) {
            const u32 bit = P + incl - L, w0 = bit >> 5, s2 = bit & 31u;
            const u32 nws = (L + 31u) >> 5;                // words of the lane buffer in use
            const u32 nst = ((s2 + L - 1u) >> 5) + 1u;     // stage words the lane's bits fall into
            // stage word w0 + k = low part of buffer word k | high part of buffer word k - 1; the first and the last
            // one may hold bits of the neighbouring lanes too
            u32 prev = 0;
            for (u32 k = 0; k < nst; k++) {
                const u32 cur = k < nws ? lbuf[32u * k + lane] : 0u;
                const u32 v = s2 ? __funnelshift_l(prev, cur, s2) : cur;
                if (k == 0 || k + 1u == nst) atomicOr(&stage[w0 + k], v);
                else stage[w0 + k] = v;
                prev = cur;
            }
        }
        __syncwarp();
        P += tot;
        // flush the complete words, keep the partial one
        const u32 nw = P >> 5;
        for (u32 i = lane; i < nw; i += 32) {
            u32 wv = stage[i];
            u8 *g = gbase + 4u * i;
            if (i == 0 && first && al) {
                for (u32 k = al; k < 4; k++) g[k] = (u8)(wv >> (8u * k));
            } else
                *(u32 *)g = wv;
        }
        __syncwarp();
        const u32 carry = stage[nw];
        __syncwarp();
        for (u32 i = lane; i < nw + 3u; i += 32) stage[i] = 0;
        __syncwarp();
        if (lane == 0) stage[0] = carry;
        __syncwarp();
        if (nw) {
            gbase += 4u * nw;
            P &= 31u;
            first = false;
        }
    }
    if (lane == 0) stage[P >> 5] |= 1u << (P & 31u);  // end mark
    P += 1;
    __syncwarp();
    u32 nbytes = (P + 7u) >> 3;  // stage bytes in use (the first `al` of the first word are not ours)
    u32 nw = nbytes >> 2;
    for (u32 i = lane; i < nw; i += 32) {
        u32 wv = stage[i];
        u8 *g = gbase + 4u * i;
        if (i == 0 && first && al) {
            for (u32 k = al; k < 4; k++) g[k] = (u8)(wv >> (8u * k));
        } else
            *(u32 *)g = wv;
    }
    u32 t0 = 4u * nw;
    if (lane < nbytes - t0) {
        u32 i = t0 + lane;
        if (!(first && i < al)) gbase[i] = (u8)(stage[i >> 2] >> (8u * (i & 3u)));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(ZH_THREADS) k_zenc_huf(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots,
                                                         u32 *out_sizes, const u32 *lzflags) {
    __shared__ ZhShared S;
    u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    u32 fi = index ? index[blockIdx.x] : blockIdx.x;
    if (lzflags && frames[fi].pad && lzflags[frames[fi].pad - 1u]) {
        // a stream with duplicated records is coded as ONE frame by the record matcher (k_lzrec_*), which writes
        // the size of the whole stream into its first frame's entry afterwards
        if (tid == 0) out_sizes[fi] = 0;
        return;
    }
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    u8 *out = slots + fr.dst_off;
    const u32 n = fr.src_len;
    const u32 nblk = (n + FQZ_ZBLOCK_ENT - 1) / FQZ_ZBLOCK_ENT;  // <= 8
    // ---- 1. per-stream histograms
    for (u32 i = tid; i < 32 * 128; i += ZH_THREADS) (&S.shist[0][0])[i] = 0;
    __syncthreads();
    if (warp < nblk) {
        u32 b0 = warp * FQZ_ZBLOCK_ENT, bn = min(FQZ_ZBLOCK_ENT, n - b0), seg = (bn + 3u) >> 2;
        for (u32 k = 0; k < 4; k++) {
            u32 a = b0 + min(k * seg, bn), b = (k == 3) ? b0 + bn : b0 + min((k + 1) * seg, bn);
            warp_stream_hist(src, a, b, S.shist[warp * 4 + k]);
        }
    }
    __syncthreads();
    {
        u32 c = 0;
        for (u32 st = 0; st < 32; st++) c += (S.shist[st][tid >> 1] >> (16u * (tid & 1u))) & 0xFFFFu;
        S.hist[tid] = c;
    }
    __syncthreads();
    // ---- 2. code
    u32 mode = 0;  // 0 raw block, 1 RLE block, 2 Huffman blocks
    u32 maxBits = 0;
    // all 256 byte values within 25 % of n / 256 (2-bit packed random bases): a Huffman code cannot gain a percent,
    // so the frame is stored without building one
    const u32 hc = S.hist[tid];
    const bool flat = !__syncthreads_or(4u * 256u * hc > 5u * n || 4u * 256u * hc < 3u * n);
    if (n >= ZH_MIN_HUF && !flat) maxBits = cta_huf_build(S.hist, n, S.hlut, S.huf, S.keys, S.misc);
    else {
        u32 distinct = (u32)__syncthreads_count(S.hist[tid] != 0);
        if (tid == 0) S.misc[3] = distinct;
        __syncthreads();
    }
    const u32 distinct = S.misc[3];
    if (distinct == 1 && n > 1) mode = 1;
    else if (maxBits) {
        if (tid == 0) S.misc[1] = huf_write_tree(S.tree, S.huf, maxBits, S.misc[4], S.norm, S.tmpsym);
        // ---- 3. exact stream sizes
        {
            u32 st = tid >> 3, j = tid & 7u, bits = 0;
            for (u32 sy = 32u * j; sy < 32u * j + 32u; sy++) bits += ((S.shist[st][sy >> 1] >> (16u * (sy & 1u))) & 0xFFFFu) * S.huf.len[sy];
            bits = group_sum(bits, group_mask(8), 8);
            if (j == 0) {
                S.sbits[st] = bits;
                S.ssize[st] = (bits + 8u) >> 3;  // + end mark, rounded up
            }
        }
        __syncthreads();
        if (tid == 0) {
            u32 treeSize = S.misc[1];
            u32 off = 10;
            for (u32 k = 0; k < nblk; k++) {
                u32 bn = min(FQZ_ZBLOCK_ENT, n - k * FQZ_ZBLOCK_ENT);
                if (bn < ZH_MIN_HUF) {  // short tail: raw block
                    S.sdst[4 * k] = off + 3;
                    off += 3 + bn;
                } else {
                    u32 pos = off + 3 + 5 + (k == 0 ? treeSize : 0u) + 6;
                    for (u32 q = 0; q < 4; q++) {
                        S.sdst[4 * k + q] = pos;
                        pos += S.ssize[4 * k + q];
                    }
                    off = pos + 1;
                }
            }
            S.misc[2] = off + 4;
            S.misc[0] = (treeSize != 0 && off + 4 < n + 17u) ? 2u : 0u;
        }
        __syncthreads();
        mode = S.misc[0];
    }
    const u32 hsh = hashes[fi];
    if (mode == 2) {
        if (tid == 0) {  // frame header, block headers, literals headers, tree, jump tables
            u32 treeSize = S.misc[1];
            out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
            out[4] = 0x84;  // FCS 4 bytes, window descriptor present, content checksum, no dictionary
            out[5] = (u8)((17 - 10) << 3);  // window = 128 KiB
            out[6] = (u8)n; out[7] = (u8)(n >> 8); out[8] = (u8)(n >> 16); out[9] = (u8)(n >> 24);
            for (u32 k = 0; k < nblk; k++) {
                u32 bn = min(FQZ_ZBLOCK_ENT, n - k * FQZ_ZBLOCK_ENT);
                u32 lastf = (k + 1 == nblk) ? 1u : 0u;
                if (bn < ZH_MIN_HUF) {
                    u8 *bh = out + S.sdst[4 * k] - 3;
                    u32 h = lastf | (0u << 1) | (bn << 3);
                    bh[0] = (u8)h; bh[1] = (u8)(h >> 8); bh[2] = (u8)(h >> 16);
                    continue;
                }
                u32 tsz = (k == 0) ? treeSize : 0u;
                u8 *bh = out + S.sdst[4 * k] - 6 - tsz - 5 - 3;
                u32 csize = tsz + 6 + S.ssize[4 * k] + S.ssize[4 * k + 1] + S.ssize[4 * k + 2] + S.ssize[4 * k + 3];
                u32 content = 5 + csize + 1;
                u32 h = lastf | (2u << 1) | (content << 3);
                bh[0] = (u8)h; bh[1] = (u8)(h >> 8); bh[2] = (u8)(h >> 16);
                u8 *lh = bh + 3;
                u32 lw = (k == 0 ? 2u : 3u) | (3u << 2) | (bn << 4) | (csize << 22);  // 4 streams, 18-bit sizes
                lh[0] = (u8)lw; lh[1] = (u8)(lw >> 8); lh[2] = (u8)(lw >> 16); lh[3] = (u8)(lw >> 24);
                lh[4] = (u8)(csize >> 10);
                u8 *tp = lh + 5;
                for (u32 i = 0; i < tsz; i++) tp[i] = S.tree[i];
                u8 *jt = tp + tsz;
                st_u16_unaligned(jt, S.ssize[4 * k]);
                st_u16_unaligned(jt + 2, S.ssize[4 * k + 1]);
                st_u16_unaligned(jt + 4, S.ssize[4 * k + 2]);
                out[S.sdst[4 * k + 3] + S.ssize[4 * k + 3]] = 0;  // sequences section: none
            }
            u8 *ck = out + S.misc[2] - 4;
            ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
            out_sizes[fi] = S.misc[2];
        }
        // ---- 4. encode.  The per-stream histograms (and the code builder's arrays behind them) are done with: their
        //      memory becomes the lanes' bit buffers
        __syncthreads();
        u32 *lbuf = &S.shist[0][0] + warp * (ZH_LANE_WORDS * 32u);
        static_assert(ZH_WARPS * ZH_LANE_WORDS * 32u <= 32u * 128u, "lane buffers must fit into the stream histograms");
        if (warp < nblk) {
            u32 b0 = warp * FQZ_ZBLOCK_ENT, bn = min(FQZ_ZBLOCK_ENT, n - b0), seg = (bn + 3u) >> 2;
            if (bn < ZH_MIN_HUF) {
                u8 *d = out + S.sdst[4 * warp];
                for (u32 i = lane; i < bn; i += 32) d[i] = src[b0 + i];
            } else {
                for (u32 k = 0; k < 4; k++) {
                    u32 a = b0 + min(k * seg, bn), b = (k == 3) ? b0 + bn : b0 + min((k + 1) * seg, bn);
                    warp_stream_encode(src, a, b, S.hlut, out + S.sdst[4 * warp + k], S.stage[warp], lbuf);
                }
            }
        }
        return;
    }
    // ---- raw / RLE: a single block
    if (tid == 0) {
        out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
        out[4] = 0x84;
        out[5] = (u8)((17 - 10) << 3);
        out[6] = (u8)n; out[7] = (u8)(n >> 8); out[8] = (u8)(n >> 16); out[9] = (u8)(n >> 24);
        u32 h = 1u | (mode << 1) | (n << 3);
        out[10] = (u8)h; out[11] = (u8)(h >> 8); out[12] = (u8)(h >> 16);
        u32 payload = (mode == 1) ? 1u : n;
        if (mode == 1) out[13] = src[0];
        u8 *ck = out + 13 + payload;
        ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
        out_sizes[fi] = 13 + payload + 4;
    }
    if (mode == 0) {  // dst-aligned word copy
        u8 *d = out + 13;
        u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
        if (head > n) head = n;
        if (tid < head) d[tid] = src[tid];
        u32 nw = (n - head) >> 2;
        for (u32 w = tid; w < nw; w += ZH_THREADS) *(u32 *)(d + head + 4u * w) = ld_u32_unaligned(src + head + 4u * w);
        u32 t0 = head + 4u * nw;
        if (tid < n - t0) d[t0 + tid] = src[t0 + tid];
    }
}

// ---------------------------------------------------------------------------------- literals-only frames in three kernels
// k_zenc_huf above keeps 255 threads of a CTA waiting at barriers while one thread merges the Huffman tree, FSE-codes
// its weights and lays the frame out (ncu: a third of its stall samples).  Split by the shape of the work instead:
//   k_zh_hist    one CTA per frame: the 32 per-stream histograms (all warps busy), written to HBM transposed
//   k_zh_plan    one WARP per frame: frame histogram, code (warp_huf_build), tree description, exact stream sizes,
//                layout, every header byte of the frame — the serial steps cost one warp, not eight, and
//                thousands of frames are planned side by side
//   k_zh_encode  one CTA per frame: the 32 streams coded to their final places (no barrier after the first)
// The plan of a frame (mode, stream positions, code) travels through HBM: 0.7 KB per 128 KiB frame.
struct ZhPlan {
    u32 mode;  // 0 raw block, 1 RLE block, 2 Huffman blocks
    u32 total;
    u32 sdst[32];
    u16 hlut[256];
};
#define ZH_HIST_WORDS (128u * 32u + 256u)  // per frame: shist[word][stream] (two u16 counters per word), then hist[256]
__global__ void __launch_bounds__(ZH_THREADS) k_zh_hist(const ZFrame *frames, const u32 *index, u32 nidx, u32 *out_sizes, const u32 *lzflags,
                                                        u32 *hist_g) {
    __shared__ u32 shist[32][128];
    u32 tid = threadIdx.x, warp = tid >> 5;
    u32 fi = index ? index[blockIdx.x] : blockIdx.x;
    if (lzflags && frames[fi].pad && lzflags[frames[fi].pad - 1u]) {
        // a segment with duplicated records is coded as ONE frame by the record matcher (k_lzrec_*), which writes
        // its size into its first frame's entry afterwards
        if (tid == 0) out_sizes[fi] = 0;
        return;
    }
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    const u32 n = fr.src_len;
    const u32 nblk = (n + FQZ_ZBLOCK_ENT - 1) / FQZ_ZBLOCK_ENT;  // <= 8
    for (u32 i = tid; i < 32 * 128; i += ZH_THREADS) (&shist[0][0])[i] = 0;
    __syncthreads();
    if (warp < nblk) {
        u32 b0 = warp * FQZ_ZBLOCK_ENT, bn = min(FQZ_ZBLOCK_ENT, n - b0), seg = (bn + 3u) >> 2;
        for (u32 k = 0; k < 4; k++) {
            u32 a = b0 + min(k * seg, bn), b = (k == 3) ? b0 + bn : b0 + min((k + 1) * seg, bn);
            warp_stream_hist(src, a, b, shist[warp * 4 + k]);
        }
    }
    __syncthreads();
    u32 *g = hist_g + (size_t)blockIdx.x * ZH_HIST_WORDS;
    for (u32 i = tid; i < 32 * 128; i += ZH_THREADS) g[i] = shist[i & 31u][i >> 5];  // [word][stream]: the planner's lane = stream
    {
        u32 c = 0;
        for (u32 st = 0; st < 32; st++) c += (shist[st][tid >> 1] >> (16u * (tid & 1u))) & 0xFFFFu;
        g[32 * 128 + tid] = c;
    }
}
#define ZHP_WARPS 4
struct ZhPlanScratch {
    u32 hist[256];
    u16 hlut[256];
    HufTmp huf;
    u8 tmpsym[512];
    short norm[64];
    u8 tree[384];
    u32 ssize[32], sdst[32];
    u32 misc[4];  // 0 mode, 1 treeSize, 2 total bytes
};
__global__ void __launch_bounds__(ZHP_WARPS * 32) k_zh_plan(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u32 *out_sizes,
                                                            const u32 *lzflags, const u32 *hist_g, ZhPlan *plans) {
    __shared__ ZhPlanScratch scratch[ZHP_WARPS];
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    const u32 wi = blockIdx.x * ZHP_WARPS + warp;
    if (wi >= nidx) return;
    const u32 fi = index ? index[wi] : wi;
    if (lzflags && frames[fi].pad && lzflags[frames[fi].pad - 1u]) return;
    ZhPlanScratch &S = scratch[warp];
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    u8 *out = slots + fr.dst_off;
    const u32 n = fr.src_len;
    const u32 nblk = (n + FQZ_ZBLOCK_ENT - 1) / FQZ_ZBLOCK_ENT;
    const u32 *g = hist_g + (size_t)wi * ZH_HIST_WORDS;
    bool notflat = false;
    u32 distinct = 0;
    for (u32 i = lane; i < 256; i += 32) {
        const u32 hc = g[32 * 128 + i];
        S.hist[i] = hc;
        // all 256 byte values within 25 % of n / 256 (2-bit packed random bases): a Huffman code cannot gain a percent
        notflat |= 4u * 256u * hc > 5u * n || 4u * 256u * hc < 3u * n;
        distinct += hc != 0;
    }
    __syncwarp();
    notflat = __any_sync(FULL, notflat);
    distinct = __reduce_add_sync(FULL, distinct);
    u32 mode = 0, maxBits = 0, maxSym = 0;
    if (n >= ZH_MIN_HUF && notflat) maxBits = warp_huf_build(S.hist, n, S.hlut, S.huf, &maxSym);
    if (distinct == 1 && n > 1) mode = 1;
    else if (maxBits) {
        if (lane == 0) S.misc[1] = huf_write_tree(S.tree, S.huf, maxBits, maxSym, S.norm, S.tmpsym);
        // exact size of stream `lane` = <its histogram, code lengths>
        {
            u32 bits = 0;
            for (u32 w = 0; w < 128; w++) {
                const u32 v = g[32u * w + lane];
                bits += (v & 0xFFFFu) * S.huf.len[2 * w] + (v >> 16) * S.huf.len[2 * w + 1];
            }
            S.ssize[lane] = (bits + 8u) >> 3;  // + end mark, rounded up
        }
        __syncwarp();
        if (lane == 0) {
            u32 treeSize = S.misc[1];
            u32 off = 10;
            for (u32 k = 0; k < nblk; k++) {
                u32 bn = min(FQZ_ZBLOCK_ENT, n - k * FQZ_ZBLOCK_ENT);
                if (bn < ZH_MIN_HUF) {  // short tail: raw block
                    S.sdst[4 * k] = off + 3;
                    off += 3 + bn;
                } else {
                    u32 pos = off + 3 + 5 + (k == 0 ? treeSize : 0u) + 6;
                    for (u32 q = 0; q < 4; q++) {
                        S.sdst[4 * k + q] = pos;
                        pos += S.ssize[4 * k + q];
                    }
                    off = pos + 1;
                }
            }
            S.misc[2] = off + 4;
            S.misc[0] = (treeSize != 0 && off + 4 < n + 17u) ? 2u : 0u;
        }
        __syncwarp();
        mode = S.misc[0];
    }
    const u32 hsh = hashes[fi];
    ZhPlan &P = plans[wi];
    if (mode == 2) {
        if (lane == 0) {  // frame header, block headers, literals headers, tree, jump tables
            u32 treeSize = S.misc[1];
            out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
            out[4] = 0x84;  // FCS 4 bytes, window descriptor present, content checksum, no dictionary
            out[5] = (u8)((17 - 10) << 3);  // window = 128 KiB
            out[6] = (u8)n; out[7] = (u8)(n >> 8); out[8] = (u8)(n >> 16); out[9] = (u8)(n >> 24);
            for (u32 k = 0; k < nblk; k++) {
                u32 bn = min(FQZ_ZBLOCK_ENT, n - k * FQZ_ZBLOCK_ENT);
                u32 lastf = (k + 1 == nblk) ? 1u : 0u;
                if (bn < ZH_MIN_HUF) {
                    u8 *bh = out + S.sdst[4 * k] - 3;
                    u32 h = lastf | (0u << 1) | (bn << 3);
                    bh[0] = (u8)h; bh[1] = (u8)(h >> 8); bh[2] = (u8)(h >> 16);
                    continue;
                }
                u32 tsz = (k == 0) ? treeSize : 0u;
                u8 *bh = out + S.sdst[4 * k] - 6 - tsz - 5 - 3;
                u32 csize = tsz + 6 + S.ssize[4 * k] + S.ssize[4 * k + 1] + S.ssize[4 * k + 2] + S.ssize[4 * k + 3];
                u32 content = 5 + csize + 1;
                u32 h = lastf | (2u << 1) | (content << 3);
                bh[0] = (u8)h; bh[1] = (u8)(h >> 8); bh[2] = (u8)(h >> 16);
                u8 *lh = bh + 3;
                u32 lw = (k == 0 ? 2u : 3u) | (3u << 2) | (bn << 4) | (csize << 22);  // 4 streams, 18-bit sizes
                lh[0] = (u8)lw; lh[1] = (u8)(lw >> 8); lh[2] = (u8)(lw >> 16); lh[3] = (u8)(lw >> 24);
                lh[4] = (u8)(csize >> 10);
                u8 *tp = lh + 5;
                for (u32 i = 0; i < tsz; i++) tp[i] = S.tree[i];
                u8 *jt = tp + tsz;
                st_u16_unaligned(jt, S.ssize[4 * k]);
                st_u16_unaligned(jt + 2, S.ssize[4 * k + 1]);
                st_u16_unaligned(jt + 4, S.ssize[4 * k + 2]);
                out[S.sdst[4 * k + 3] + S.ssize[4 * k + 3]] = 0;  // sequences section: none
            }
            u8 *ck = out + S.misc[2] - 4;
            ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
            out_sizes[fi] = S.misc[2];
            P.mode = 2;
            P.total = S.misc[2];
        }
        P.sdst[lane] = S.sdst[lane];
        for (u32 i = lane; i < 256; i += 32) P.hlut[i] = S.hlut[i];
        return;
    }
    // ---- raw / RLE: a single block (the raw payload is copied by k_zh_encode)
    if (lane == 0) {
        out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
        out[4] = 0x84;
        out[5] = (u8)((17 - 10) << 3);
        out[6] = (u8)n; out[7] = (u8)(n >> 8); out[8] = (u8)(n >> 16); out[9] = (u8)(n >> 24);
        u32 h = 1u | (mode << 1) | (n << 3);
        out[10] = (u8)h; out[11] = (u8)(h >> 8); out[12] = (u8)(h >> 16);
        u32 payload = (mode == 1) ? 1u : n;
        if (mode == 1) out[13] = src[0];
        u8 *ck = out + 13 + payload;
        ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
        out_sizes[fi] = 13 + payload + 4;
        P.mode = mode;
        P.total = 13 + payload + 4;
    }
}
struct ZhEncShared {
    u16 hlut[256];
    u32 sdst[32];
    u32 lbuf[ZH_WARPS * ZH_LANE_WORDS * 32];
    u32 stage[ZH_WARPS][ZH_STAGE_WORDS];
};
__global__ void __launch_bounds__(ZH_THREADS) k_zh_encode(const ZFrame *frames, const u32 *index, u32 nidx, u8 *slots, const u32 *lzflags,
                                                          const ZhPlan *plans) {
    __shared__ ZhEncShared S;
    u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    u32 fi = index ? index[blockIdx.x] : blockIdx.x;
    if (lzflags && frames[fi].pad && lzflags[frames[fi].pad - 1u]) return;
    ZFrame fr = frames[fi];
    const u8 *src = (const u8 *)(uintptr_t)fr.src;
    u8 *out = slots + fr.dst_off;
    const u32 n = fr.src_len;
    const u32 nblk = (n + FQZ_ZBLOCK_ENT - 1) / FQZ_ZBLOCK_ENT;
    const ZhPlan &P = plans[blockIdx.x];
    const u32 mode = P.mode;
    if (mode == 2) {
        S.hlut[tid] = P.hlut[tid];
        if (tid < 32) S.sdst[tid] = P.sdst[tid];
        __syncthreads();
        if (warp < nblk) {
            u32 b0 = warp * FQZ_ZBLOCK_ENT, bn = min(FQZ_ZBLOCK_ENT, n - b0), seg = (bn + 3u) >> 2;
            if (bn < ZH_MIN_HUF) {
                u8 *d = out + S.sdst[4 * warp];
                for (u32 i = lane; i < bn; i += 32) d[i] = src[b0 + i];
            } else {
                u32 *lbuf = S.lbuf + warp * (ZH_LANE_WORDS * 32u);
                for (u32 k = 0; k < 4; k++) {
                    u32 a = b0 + min(k * seg, bn), b = (k == 3) ? b0 + bn : b0 + min((k + 1) * seg, bn);
                    warp_stream_encode(src, a, b, S.hlut, out + S.sdst[4 * warp + k], S.stage[warp], lbuf);
                }
            }
        }
        return;
    }
    if (mode == 0) {  // dst-aligned word copy of the raw block
        u8 *d = out + 13;
        u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
        if (head > n) head = n;
        if (tid < head) d[tid] = src[tid];
        u32 nw = (n - head) >> 2;
        for (u32 w = tid; w < nw; w += ZH_THREADS) *(u32 *)(d + head + 4u * w) = ld_u32_unaligned(src + head + 4u * w);
        u32 t0 = head + 4u * nw;
        if (tid < n - t0) d[t0 + tid] = src[t0 + tid];
    }
}

// ---------------------------------------------------------------------------------- duplicated records in the literals-only streams
// Packed bases and qualities of i.i.d. reads hold no matches worth coding (k_zenc_huf), but real runs do:
// PCR / optical duplicates, identical test records, all-'F' quality lines.  They are looked for at RECORD
// granularity — one hash per record, not per byte:
//   k_rec_keys   one thread per record: key = hash of its length and of three 16-byte windows (head, middle, tail)
//   k_rec_match  one warp per segment (FQZ_ZSEG = 2 MiB of one stream of one block): walks the keys in order, 32
//                records per step, and pairs every record with the nearest earlier record of the same key
//                (two-probe table of the last ~8 K records in shared memory for the earlier steps, __match_any
//                inside the step; the last lane that wants a slot gets it, so the result does not depend on
//                scheduling).  A segment in which at least 1/16 of the records found a partner is flagged.
//   k_lzrec_*    a flagged segment is NOT cut into independent 128 KiB frames (every frame would have to send
//                the recurring reads again): it becomes ONE zstd frame of 16 KiB blocks, but all blocks are
//                parsed and entropy-coded in parallel, one warp each — the item matcher at the partner's offset
//                and at offset 1 (matches reach back into the earlier blocks of the segment), then the literals
//                / sequences writers of the item streams.  Only the repeat-offset history of the block in front
//                is unknown to a warp: see warp_item_parse.  2 MiB per frame keeps eight frames of a quality
//                stream decoding side by side (a frame with sequences is one serial chain for any decoder).
// Unflagged segments keep the literals-only frames; the cost there is the key pass and the parallel count.
#define ZR_HLOG 13
__device__ __forceinline__ u32 zr_lower_bound(const u32 *items, u32 n, u32 v) {  // first i in [0, n] with items[i] >= v (items[n] = sentinel)
    u32 a = 0, b = n;
    while (a < b) {
        u32 mid = (a + b) >> 1;
        if (items[mid] < v) a = mid + 1; else b = mid;
    }
    return a;
}
// records that START inside the segment: ranges[2 si] = number of the first one inside the block, ranges[2 si + 1] = count
__global__ void k_rec_ranges(const ZRStream *rs, u32 ns, u32 *ranges) {
    u32 si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= ns) return;
    const ZRStream &R = rs[si];
    const u32 *items = (const u32 *)(uintptr_t)R.items + R.rec0;
    u32 lo = zr_lower_bound(items, R.nrec, R.item_base), hi = zr_lower_bound(items, R.nrec, R.item_base + R.len);
    ranges[2 * si] = lo;
    ranges[2 * si + 1] = hi - lo;
}
__global__ void __launch_bounds__(256) k_rec_keys(const ZRStream *rs, const u32 *ranges, const u32 *offs_base, u32 *keys_base) {
    const ZRStream &R = rs[blockIdx.y];
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ranges[2 * blockIdx.y + 1]) return;
    r += ranges[2 * blockIdx.y];
    const u32 *items = (const u32 *)(uintptr_t)R.items + R.rec0;
    u32 s = items[r], n = items[r + 1] - s, key = 0;
    if (n >= 8u) {
        // three 16-byte windows (head, middle, tail; they overlap in short records) and the length: reads that merely
        // open alike — most quality lines of a run do — must not pair up, exact and near-exact copies must
        const u8 *P = (const u8 *)(uintptr_t)R.src + (s - R.item_base);  // the record starts inside the segment (it may end behind it)
        const u32 m = min(n, 16u) & ~3u;  // whole words of a window
        const u32 o1 = (n - m) >> 1, o2 = n - m;
        u32 h = n * 0x9E3779B1u;
        for (u32 k = 0; k < m; k += 4) {
            u32 w0 = ld_u32_unaligned(P + k), w1 = ld_u32_unaligned(P + o1 + k), w2 = ld_u32_unaligned(P + o2 + k);
            h = (h ^ w0) * 0x85EBCA77u;
            h = (h << 13 | h >> 19);
            h = (h ^ w1) * 0xC2B2AE3Du;
            h = (h << 11 | h >> 21);
            h = (h ^ w2) * 0x27D4EB2Fu;
            h = (h << 15 | h >> 17);
        }
        h ^= h >> 15;
        h *= 0x2C1B3C6Du;
        h ^= h >> 12;
        h *= 0x297A2D39u;
        h ^= h >> 15;
        key = h | 1u;
    }
    keys_base[(items - offs_base) + r] = key;
}
// Cheap, fully parallel look for repeated keys, so that the serial pairing pass only runs on streams that need it:
// one CTA per 2048 records counts the records whose key occurs earlier in the same chunk.  Order-independent
// (every slot keeps the LARGEST key that hashes to it, then counts its occurrences), hence the same on every run;
// keys that lose their slot are not counted, so this is a lower bound.  dupcnt[stream] += sum of (occurrences - 1).
#define ZR_DET_RECORDS 2048u
#define ZR_DET_SLOTS 4096u
__global__ void __launch_bounds__(256) k_rec_detect(const ZRStream *rs, const u32 *ranges, const u32 *offs_base, const u32 *keys_base, u32 *dupcnt) {
    __shared__ u32 tkey[ZR_DET_SLOTS], tcnt[ZR_DET_SLOTS];
    __shared__ u32 ws[33];
    const ZRStream &R = rs[blockIdx.y];
    const u32 nseg = ranges[2 * blockIdx.y + 1];
    const u32 r0 = blockIdx.x * ZR_DET_RECORDS;
    if (r0 >= nseg) return;
    const u32 r1 = min(nseg, r0 + ZR_DET_RECORDS);
    const u32 *keys = keys_base + ((const u32 *)(uintptr_t)R.items - offs_base) + R.rec0 + ranges[2 * blockIdx.y];
    for (u32 i = threadIdx.x; i < ZR_DET_SLOTS; i += blockDim.x) tkey[i] = 0, tcnt[i] = 0;
    __syncthreads();
    u32 k[ZR_DET_RECORDS / 256];
#pragma unroll
    for (u32 j = 0; j < ZR_DET_RECORDS / 256; j++) {
        u32 r = r0 + j * 256u + threadIdx.x;
        k[j] = r < r1 ? keys[r] : 0u;
        if (k[j]) atomicMax(&tkey[(k[j] >> 1) & (ZR_DET_SLOTS - 1u)], k[j]);
    }
    __syncthreads();
#pragma unroll
    for (u32 j = 0; j < ZR_DET_RECORDS / 256; j++)
        if (k[j] && tkey[(k[j] >> 1) & (ZR_DET_SLOTS - 1u)] == k[j]) atomicAdd(&tcnt[(k[j] >> 1) & (ZR_DET_SLOTS - 1u)], 1u);
    __syncthreads();
    u32 d = 0;
    for (u32 i = threadIdx.x; i < ZR_DET_SLOTS; i += blockDim.x) d += tcnt[i] > 1u ? tcnt[i] - 1u : 0u;
    u32 tot = 0;
    block_excl_scan(d, ws, &tot);
    if (threadIdx.x == 0 && tot) atomicAdd(&dupcnt[blockIdx.y], tot);
}
// cand[r] = 1 + number (inside the stream's block) of the partner record, 0 = none.  flags[ns] = any stream flagged.
__global__ void __launch_bounds__(32) k_rec_match(const ZRStream *rs, u32 ns, const u32 *offs_base, const u32 *keys_base, u32 *cand_base, u32 *flags,
                                                  const u32 *dupcnt, const u32 *ranges) {
    __shared__ u32 tab[1 << ZR_HLOG];  // key bits 17.. | 1 + record number (17 bits: a block holds 100 000 records)
    const u32 si = blockIdx.x, lane = lane_id();
    if (si >= ns) return;
    const ZRStream R = rs[si];
    const u32 first = ranges[2 * si];  // the segment's records inside the block
    const size_t at = ((const u32 *)(uintptr_t)R.items - offs_base) + R.rec0 + first;
    const u32 *keys = keys_base + at;
    u32 *cand = cand_base + at;
    const u32 nrec = ranges[2 * si + 1];
    // records of 20 bytes or more on average (the match list of a block holds one entry per two bytes), and enough
    // repeated keys for the 1/16 below to be in reach (the chunk-local count misses the far pairs: ask for half)
    if (nrec < 8u || (u64)nrec * 20u > R.len || nrec >= (1u << 17) || dupcnt[si] * 32u < nrec) {
        if (lane == 0) flags[si] = 0;
        return;
    }
    for (u32 i = lane; i < (1u << ZR_HLOG); i += 32) tab[i] = 0;
    __syncwarp();
    u32 hits = 0;
    u32 knext = lane < nrec ? keys[lane] : 0u;
    for (u32 ib = 0; ib < nrec; ib += 32) {
        const u32 i = ib + lane;
        const bool live = i < nrec;
        const u32 key = knext;
        knext = (i + 32 < nrec) ? keys[i + 32] : 0u;
        // two probes per key (a direct-mapped table loses ~4 % of the partners to slot collisions, and a lost
        // partner costs a whole record of literals)
        const u32 s1 = (key >> 1) & ((1u << ZR_HLOG) - 1u), s2 = ((key >> 14) ^ (key >> 25) ^ 0x555u) & ((1u << ZR_HLOG) - 1u);
        const u32 khi = key & 0xFFFE0000u;
        u32 c = 0, t1 = 0, t2 = 0;
        if (key) {
            t1 = tab[s1];
            t2 = tab[s2];
            if (t1 && (t1 & 0xFFFE0000u) == khi) c = t1 & 0x1FFFFu;
            else if (t2 && (t2 & 0xFFFE0000u) == khi) c = t2 & 0x1FFFFu;
        }
        u32 m = __match_any_sync(FULL, key);
        if (key) {
            u32 lower = m & ((1u << lane) - 1u);
            if (lower) c = ib + (31u - (u32)__clz((int)lower)) + 1u;
        }
        if (live) {
            cand[i] = c ? first + c : 0u;  // numbered inside the block, like items[]
            hits += c ? 1u : 0u;
        }
        // insert: the key's own slot if it has one, else a free one, else the slot whose record is older; all lanes
        // decide on the table as the step found it, the last lane that wants a slot gets it (same result on every run)
        u32 slot = s1;
        if (key) {
            if (t1 && (t1 & 0xFFFE0000u) == khi) slot = s1;
            else if (t2 && (t2 & 0xFFFE0000u) == khi) slot = s2;
            else if (!t1) slot = s1;
            else if (!t2) slot = s2;
            else slot = ((t1 & 0x1FFFFu) <= (t2 & 0x1FFFFu)) ? s1 : s2;
        }
        u32 ms = __match_any_sync(FULL, key ? slot : (0x10000u + lane));
        __syncwarp();
        if (key && (ms >> lane) == 1u) tab[slot] = khi | (i + 1u);
        __syncwarp();
    }
    hits = __reduce_add_sync(FULL, hits);
#ifdef FQZ_EMU
    if (lane == 0 && getenv("FQZ_DEBUG")) fprintf(stderr, "rec_match stream %u len %u records %u hits %u\n", si, R.len, nrec, hits);
#endif
    if (lane == 0) {
        const u32 f = (hits * 16u >= nrec) ? 1u : 0u;
        flags[si] = f;
        if (f) flags[ns] = 1u;
    }
}
// content checksum of the flagged streams (one quad per stream; unflagged streams cost nothing)
__global__ void __launch_bounds__(XX_WARPS * 32) k_xxh64_streams(const ZRStream *rs, u32 ns, const u32 *flags, u32 *hashes) {
    FQZ_DYN_SMEM(u8, smem);
    if (!flags[ns]) return;  // no stream flagged
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 si = t >> 2, q = t & 3;
    u32 gmask = group_mask(4);
    bool live = si < ns;
    u32 sj = live ? si : ns - 1;
    u8 *rows;
    u64 *bars;
    xx_quad_smem(smem, &rows, &bars);
    u64 h = xxh64_quad_staged((const u8 *)(uintptr_t)rs[sj].src, (live && flags[sj]) ? rs[sj].len : 0u, q, gmask, rows, bars);
    if (live && q == 0) hashes[si] = (u32)h;
}

// per-block staging of the flagged streams
#define ZR_OUT_STRIDE ((size_t)FQZ_ZBLOCK_ENT + 64u)
struct ZRBlock {
    const u8 *src;
    u32 len, back, k, nblk, si;
    u8 *lit, *out;
    u16 *sll, *sml, *slo;
    u32 *sof;
};
// block g (numbered through all streams of the batch) -> its stream and place; false: nothing to do
__device__ __forceinline__ bool zr_block(const ZRStream *rs, u32 ns, const u32 *flags, u8 *pool_ws, u8 *pool_out, u32 g0, u32 g, u32 gend, ZRStream &R,
                                         ZRBlock &B) {
    if (g >= gend || !flags[ns]) return false;  // flags[ns]: any stream flagged
    u32 lo = 0, hi = ns - 1;  // largest s with rs[s].blk0 <= g
    while (lo < hi) {
        u32 mid = (lo + hi + 1) >> 1;
        if (rs[mid].blk0 <= g) lo = mid; else hi = mid - 1;
    }
    if (!flags[lo]) return false;
    R = rs[lo];
    B.si = lo;
    B.k = g - R.blk0;
    B.nblk = R.nblk;
    if (B.k >= B.nblk) return false;
    B.back = B.k * FQZ_ZBLOCK_ENT;
    B.len = min(FQZ_ZBLOCK_ENT, R.len - B.back);
    B.src = (const u8 *)(uintptr_t)R.src + B.back;
    B.lit = pool_ws + (size_t)(g - g0) * FQZ_ZWS(FQZ_ZBLOCK_ENT);
    u32 maxseq = ((B.len / 2) + 2u) & ~1u;
    B.sll = (u16 *)(B.lit + ((B.len + 15u) & ~15u));
    B.sml = B.sll + maxseq;
    B.sof = (u32 *)(B.sml + maxseq);
    B.slo = (u16 *)(B.sof + maxseq);
    B.out = pool_out + (size_t)g * ZR_OUT_STRIDE;
    return true;
}
__global__ void __launch_bounds__(ZENC_WARPS * 32) k_lzrec_parse(const ZRStream *rs, u32 ns, const u32 *flags, const u32 *offs_base, const u32 *cand_base,
                                                                 u8 *pool_ws, u8 *pool_out, u32 g0, u32 gend, u32 *parsed) {
    __shared__ u32 mbufs[ZENC_WARPS][32 * ZI_MAXM * 2];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 g = g0 + blockIdx.x * ZENC_WARPS + warp;
    ZRStream R;
    ZRBlock B;
    if (!zr_block(rs, ns, flags, pool_ws, pool_out, g0, g, gend, R, B)) return;
    u32 *pr = parsed + 2 * (size_t)(g - g0);
    // the stream's own records: items[0 .. nrec] in stream coordinates, cand[] alongside
    const u32 *items = (const u32 *)(uintptr_t)R.items + R.rec0;
    const u32 *cand = cand_base + (items - offs_base);
    // one repeated byte: RLE block
    {
        u32 b0 = B.src[0];
        bool same = true;
        for (u32 i = lane; i < B.len && same; i += 32) same = (B.src[i] == b0);
        if (__all_sync(FULL, same) && B.len > 1) {
            if (lane == 0) {
                pr[0] = 0;
                pr[1] = ZENC_RLE_MARK;
            }
            return;
        }
    }
    u32 nseq = 0, nlit = 0;
    // the match list holds one entry per two bytes: only blocks whose records are >= 20 bytes on average are parsed
    const u32 b0 = R.item_base + B.back;
    const u32 ia = zr_lower_bound(items, R.nrec, b0), ib = zr_lower_bound(items, R.nrec, b0 + B.len);
    if (B.len >= 64 && (ib - ia + 1u) * 20u <= B.len + 256u)
        nseq = warp_item_parse(B.src, B.len, items, b0, R.nrec + 1u, mbufs[warp], B.lit, B.sll, B.sml, B.sof, B.slo, &nlit, B.back, cand, B.k == 0);
    if (lane == 0) {
        pr[0] = nseq;
        pr[1] = nseq ? nlit : B.len;
    }
}
// PHASE 1: literals section, PHASE 2: sequences section + block header (same split as the item frames)
template <int PHASE>
__global__ void __launch_bounds__(ZENC_WARPS * 32) k_lzrec_code(const ZRStream *rs, u32 ns, const u32 *flags, u8 *pool_ws, u8 *pool_out, u32 g0, u32 gend,
                                                                u32 *parsed, u32 *bsizes) {
    __shared__ WarpScratchItems scratch[ZENC_WARPS];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 g = g0 + blockIdx.x * ZENC_WARPS + warp;
    ZRStream R;
    ZRBlock B;
    if (!zr_block(rs, ns, flags, pool_ws, pool_out, g0, g, gend, R, B)) return;
    u32 *pr = parsed + 2 * (size_t)(g - g0);
    WarpScratchItems &S = scratch[warp];
    u8 *body = B.out + 3;
    const u32 body_cap = (u32)ZR_OUT_STRIDE - 3u - 8u;
    const u32 nseq = pr[0];
    const bool rle = pr[1] == ZENC_RLE_MARK;
    if constexpr (PHASE == 1) {
        if (rle) return;
        u32 nlit = pr[1];
        u32 lsz = warp_write_literals(nseq ? (const u8 *)B.lit : B.src, nlit, body, S);
        __syncwarp();
        if (lane == 0) pr[1] = lsz;
        return;
    } else {
        u32 btype, bsize;
        if (rle) {
            if (lane == 0) body[0] = B.src[0];
            btype = 1;
            bsize = B.len;
        } else {
            u32 lsz = pr[1], ssz = 0;
            bool ovf = false;
            if (nseq == 0) {
                if (lane == 0) body[lsz] = 0;
                ssz = 1;
            } else if (lsz + 16 < body_cap) ssz = warp_write_sequences(B.sll, B.sml, B.sof, nseq, body + lsz, body_cap - lsz, S, &ovf, true);
            else ovf = true;
            __syncwarp();
            u32 total = lsz + ssz;
            if (!ovf && total < B.len) {
                btype = 2;
                bsize = total;
            } else {
                for (u32 i = lane; i < B.len; i += 32) body[i] = B.src[i];
                btype = 0;
                bsize = B.len;
            }
        }
        if (lane == 0) {
            u32 h = ((B.k + 1u == B.nblk) ? 1u : 0u) | (btype << 1) | (bsize << 3);
            B.out[0] = (u8)h; B.out[1] = (u8)(h >> 8); B.out[2] = (u8)(h >> 16);
            bsizes[g] = 3u + ((btype == 1) ? 1u : bsize);
        }
    }
}
// strings the blocks of a flagged stream together in the output slots of its frames (one contiguous span):
// frame header, blocks, checksum.  grid (streams, chunks of ZR_CLOSE_BLOCKS blocks)
#define ZR_CLOSE_BLOCKS 32u
__global__ void __launch_bounds__(256) k_lzrec_close(const ZRStream *rs, const u32 *flags, const u32 *hashes, const u8 *pool_out, const u32 *bsizes,
                                                     const ZFrame *frames, u8 *slots, u32 *out_sizes) {
    __shared__ u32 warp_sums[33];
    const u32 si = blockIdx.x;
    if (!flags[si]) return;
    const ZRStream R = rs[si];
    const u32 k0 = blockIdx.y * ZR_CLOSE_BLOCKS;
    if (k0 >= R.nblk) return;
    u8 *out = slots + frames[R.first_frame].dst_off;
    // bytes of the blocks in front of this chunk
    u32 part = 0;
    for (u32 k = threadIdx.x; k < k0; k += blockDim.x) part += bsizes[R.blk0 + k];
    u32 tot = 0;
    block_excl_scan(part, warp_sums, &tot);
    u32 pos = 10u + tot;  // frame header: magic, descriptor, window, 4-byte content size
    const u32 kend = min(R.nblk, k0 + ZR_CLOSE_BLOCKS);
    for (u32 k = k0; k < kend; k++) {
        u32 sz = bsizes[R.blk0 + k];
        const u8 *src = pool_out + (size_t)(R.blk0 + k) * ZR_OUT_STRIDE;
        for (u32 i = threadIdx.x; i < sz; i += blockDim.x) out[pos + i] = src[i];
        pos += sz;
    }
    if (threadIdx.x == 0 && k0 == 0) {
        u32 wlog = 17;
        while (wlog < 23 && (1u << wlog) < R.len) wlog++;
        out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
        out[4] = 0x84;  // FCS 4 bytes, window descriptor present, content checksum, no dictionary
        out[5] = (u8)((wlog - 10) << 3);
        out[6] = (u8)R.len; out[7] = (u8)(R.len >> 8); out[8] = (u8)(R.len >> 16); out[9] = (u8)(R.len >> 24);
    }
    if (threadIdx.x == 0 && kend == R.nblk) {
        u32 hsh = hashes[si];
        u8 *ck = out + pos;
        ck[0] = (u8)hsh; ck[1] = (u8)(hsh >> 8); ck[2] = (u8)(hsh >> 16); ck[3] = (u8)(hsh >> 24);
        out_sizes[R.first_frame] = pos + 4;
    }
}

// ---------------------------------------------------------------------------------- XXH64 (frame content checksum)
// see fqz_xxh64.cuh: 4 threads per frame, 8 frames per warp, frames streamed through shared memory by TMA
__global__ void __launch_bounds__(XX_WARPS * 32) k_xxh64_frames(const ZFrame *frames, u32 nframes, u32 *hashes) {
    FQZ_DYN_SMEM(u8, smem);
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 fi = t >> 2, q = t & 3;
    u32 gmask = group_mask(4);
    bool live = fi < nframes;
    u32 fj = live ? fi : nframes - 1;  // keep whole quads alive for the shuffles
    ZFrame fr = frames[fj];
    u8 *rows;
    u64 *bars;
    xx_quad_smem(smem, &rows, &bars);
    u64 h = xxh64_quad_staged((const u8 *)(uintptr_t)fr.src, (live && fr.policy != FQZ_ZPOLICY_INDEX) ? fr.src_len : 0u, q, gmask, rows, bars);
    if (live && q == 0) hashes[fi] = (u32)h;
}

// One thread per frame (see FQZ_ZPOLICY_INDEX): an index frame writes its header, every frame that is
// listed in one writes its own entry.
__global__ void __launch_bounds__(128) k_zindex(const ZFrame *frames, u32 nframes, u8 *slots, u32 *out_sizes, const u32 *lzflags) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const ZFrame &F = frames[f];
    if (lzflags && F.pad) {
        // a stream of which any segment is coded as one frame carries no index (its frames are no longer one per entry)
        const ZFrame &X = (F.policy == FQZ_ZPOLICY_INDEX) ? F : (F.index_of ? frames[F.index_of - 1u] : F);
        bool any = false;
        if (X.policy == FQZ_ZPOLICY_INDEX)
            for (u32 k = 0; k < X.item_count; k++) any |= lzflags[X.pad - 1u + k] != 0;
        if (any) {
            if (F.policy == FQZ_ZPOLICY_INDEX) out_sizes[f] = 0;
            return;
        }
    }
    if (F.policy == FQZ_ZPOLICY_INDEX) {
        u32 n = F.src_len;
        u32 *o = (u32 *)(slots + F.dst_off);  // slots are 16-byte aligned
        o[0] = FQZ_ZINDEX_MAGIC;
        o[1] = FQZ_ZINDEX_BYTES(n) - 8u;
        o[2] = FQZ_ZINDEX_SIG;
        o[3] = n;
        o[4] = frames[f + 1].src_len;  // every frame but the last holds this many content bytes
        out_sizes[f] = FQZ_ZINDEX_BYTES(n);
        return;
    }
    if (!F.index_of) return;
    u32 fi = F.index_of - 1u, k = f - fi - 1u;
    u32 *o = (u32 *)(slots + frames[fi].dst_off);
    u32 first = 0, cnt = 0;
    const u32 *items = (const u32 *)(uintptr_t)F.items;
    if (F.policy == FQZ_ZPOLICY_ITEMS && items) {
        // items[] = start of every item in stream coordinates (ascending, last entry = end sentinel):
        // i0 = first item starting at or after the frame's first byte, i1 = same for the frame's end
        u32 nit = F.item_count - 1u;
        u32 lo0 = F.item_base, lo1 = F.item_base + F.src_len;
        u32 a = 0, b = nit;
        while (a < b) {
            u32 mid = (a + b) >> 1;
            if (items[mid] < lo0) a = mid + 1; else b = mid;
        }
        u32 i0 = a;
        b = nit;
        while (a < b) {
            u32 mid = (a + b) >> 1;
            if (items[mid] < lo1) a = mid + 1; else b = mid;
        }
        cnt = a - i0;
        if (cnt) first = items[i0] - lo0;
        if (cnt > 0xFFFFu || first > 0xFFFFu) cnt = 0xFFFFu, first = 0xFFFFu;  // unusable hint: the decoder falls back
    }
    o[5 + 2 * k] = out_sizes[f];
    o[6 + 2 * k] = first | (cnt << 16);
}
void fqz_launch_zindex(const ZFrame *frames, u32 nframes, u8 *slots, u32 *out_sizes, const u32 *lzflags, cudaStream_t s) {
    if (!nframes) return;
    FQZ_LAUNCH(k_zindex, (nframes + 127) / 128, 128, 0, s, frames, nframes, slots, out_sizes, lzflags);
}
void fqz_launch_xxh64(const ZFrame *frames, u32 nframes, u32 *hashes, cudaStream_t s) {
    if (!nframes) return;
    u32 threads = XX_WARPS * 32, grid = (nframes * 4 + threads - 1) / threads;
    FQZ_LAUNCH(k_xxh64_frames, grid, threads, XX_SMEM, s, frames, nframes, hashes);
}
// scratch == nullptr: the one-kernel version (k_zenc_huf); else fqz_zenc_huf_scratch(nidx) bytes for the histograms and plans
size_t fqz_zenc_huf_scratch(u32 nidx) { return (size_t)nidx * (ZH_HIST_WORDS * sizeof(u32) + sizeof(ZhPlan)) + 256; }
void fqz_launch_zenc_huf(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u32 *out_sizes, const u32 *lzflags,
                         u8 *scratch, cudaStream_t s) {
    if (!nidx) return;
    if (!scratch) {
        FQZ_LAUNCH(k_zenc_huf, nidx, ZH_THREADS, 0, s, frames, index, nidx, hashes, slots, out_sizes, lzflags);
        return;
    }
    u32 *hist_g = (u32 *)scratch;
    ZhPlan *plans = (ZhPlan *)(scratch + (((size_t)nidx * ZH_HIST_WORDS * sizeof(u32) + 127) & ~(size_t)127));
    FQZ_LAUNCH(k_zh_hist, nidx, ZH_THREADS, 0, s, frames, index, nidx, out_sizes, lzflags, hist_g);
    FQZ_LAUNCH(k_zh_plan, (nidx + ZHP_WARPS - 1) / ZHP_WARPS, ZHP_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, out_sizes, lzflags, hist_g, plans);
    FQZ_LAUNCH(k_zh_encode, nidx, ZH_THREADS, 0, s, frames, index, nidx, slots, lzflags, plans);
}
void fqz_launch_rec_match(const ZRStream *rs, u32 ns, u32 max_records, const u32 *offs_base, u32 *keys_base, u32 *cand_base, u32 *flags,
                          u32 *dupcnt, u32 *ranges, u32 *hashes, cudaStream_t s) {
    if (!ns) return;
    // a segment of FQZ_ZSEG bytes takes part with at most FQZ_ZSEG / 20 records (more, i.e. shorter ones, make it ineligible)
    const u32 per_seg = std::min<u32>(max_records, FQZ_ZSEG / 20u + 1u);
    FQZ_LAUNCH(k_rec_ranges, (ns + 127) / 128, 128, 0, s, rs, ns, ranges);
    FQZ_LAUNCH(k_rec_keys, dim3((per_seg + 255) / 256, ns), 256, 0, s, rs, ranges, offs_base, keys_base);
    FQZ_LAUNCH(k_rec_detect, dim3((per_seg + ZR_DET_RECORDS - 1) / ZR_DET_RECORDS, ns), 256, 0, s, rs, ranges, offs_base, keys_base, dupcnt);
    FQZ_LAUNCH(k_rec_match, ns, 32, 0, s, rs, ns, offs_base, keys_base, cand_base, flags, dupcnt, ranges);
    u32 threads = XX_WARPS * 32, grid = (ns * 4 + threads - 1) / threads;
    FQZ_LAUNCH(k_xxh64_streams, grid, threads, XX_SMEM, s, rs, ns, flags, hashes);
}
size_t fqz_lzrec_pool_ws(u32 nblocks) { return (size_t)nblocks * FQZ_ZWS(FQZ_ZBLOCK_ENT); }
size_t fqz_lzrec_pool_out(u32 nblocks) { return (size_t)nblocks * ZR_OUT_STRIDE; }
void fqz_launch_lzrec(const ZRStream *rs, u32 ns, const u32 *flags, const u32 *offs_base, const u32 *cand_base, u8 *pool_ws, u8 *pool_out, u32 g0,
                      u32 gend, u32 *parsed, u32 *bsizes, cudaStream_t s) {
    if (!ns || gend <= g0) return;
    u32 grid = (gend - g0 + ZENC_WARPS - 1) / ZENC_WARPS;
    FQZ_LAUNCH(k_lzrec_parse, grid, ZENC_WARPS * 32, 0, s, rs, ns, flags, offs_base, cand_base, pool_ws, pool_out, g0, gend, parsed);
    FQZ_LAUNCH(k_lzrec_code<1>, grid, ZENC_WARPS * 32, 0, s, rs, ns, flags, pool_ws, pool_out, g0, gend, parsed, bsizes);
    FQZ_LAUNCH(k_lzrec_code<2>, grid, ZENC_WARPS * 32, 0, s, rs, ns, flags, pool_ws, pool_out, g0, gend, parsed, bsizes);
}
void fqz_launch_lzrec_close(const ZRStream *rs, u32 ns, u32 max_blocks, const u32 *flags, const u32 *hashes, const u8 *pool_out, const u32 *bsizes,
                            const ZFrame *frames, u8 *slots, u32 *out_sizes, cudaStream_t s) {
    if (!ns) return;
    FQZ_LAUNCH(k_lzrec_close, dim3(ns, (max_blocks + ZR_CLOSE_BLOCKS - 1) / ZR_CLOSE_BLOCKS), 256, 0, s, rs, flags, hashes, pool_out, bsizes, frames,
               slots, out_sizes);
}
void fqz_launch_zenc(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u8 *ws, u32 *out_sizes, int lz,
                     u32 *parsed, cudaStream_t s, cudaEvent_t hashes_ready) {
    if (!nidx) {
        if (hashes_ready) cudaStreamWaitEvent(s, hashes_ready, 0);
        return;
    }
    u32 grid = (nidx + ZENC_WARPS - 1) / ZENC_WARPS;
    if (lz == 2) {
        FQZ_LAUNCH(k_zitems_parse, grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, ws, parsed);
        FQZ_LAUNCH((k_zenc<2, 1>), grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, ws, out_sizes, parsed);
        if (hashes_ready) cudaStreamWaitEvent(s, hashes_ready, 0);  // only the frame close needs the checksums
        FQZ_LAUNCH((k_zenc<2, 2>), grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, ws, out_sizes, parsed);
        return;
    }
    if (hashes_ready) cudaStreamWaitEvent(s, hashes_ready, 0);
    // lz == 1: generic data; literals-only frames go through k_zenc_huf / the three-kernel coder, never through here
    FQZ_LAUNCH((k_zenc<1, 0>), grid, ZENC_WARPS * 32, 0, s, frames, index, nidx, hashes, slots, ws, out_sizes, parsed);
}
int fqz_zstd_enc_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_xxh64_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, XX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_xxh64_streams, cudaFuncAttributeMaxDynamicSharedMemorySize, XX_SMEM);
    return (int)e;
}
