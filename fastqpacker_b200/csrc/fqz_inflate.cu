// fqz_inflate.cu — gzip / DEFLATE decoder kernels: the input side of compress mode for .gz files.
//
// Reference: cmd/fqpack/main.go:142-174 wraps a gzipped input in Go's compress/gzip reader (stdlib, multistream)
// before compress.Compress parses it.  A DEFLATE stream is one serial chain (Huffman codes of unknown length,
// matches into the previous 32 KiB), so the work is cut the way pugz / rapidgzip cut it on CPUs:
//
//   k_gz_find     the compressed bytes are cut into chunks; for every chunk one warp looks for the first place a
//                 decoder can restart: a BGZF member header (byte aligned, empty window) or the header of a
//                 non-final dynamic-Huffman block (any bit offset; the 32 KiB window in front of it is unknown).
//   k_gz_decode   one warp per chunk decodes from its restart point to the next chunk's.  The bit stream is walked by
//                 the whole warp (every lane decodes the token that would start at its bit, the chain through them
//                 is followed with shuffles), up to 32 tokens per step; the warp writes the literals and copies the
//                 matches.  Output is
//                 16-bit: a byte, or 256 + i for "byte i of the window I could not see".  One speculative pass into
//                 scratch room of six symbols per compressed byte gives the symbols, the sizes, the member
//                 boundaries and the proof that every chunk lands exactly on the next restart point (a restart point
//                 nobody lands on was a false positive and its chunk is merged into the one in front); only chunks
//                 that ran out of room are decoded again, at exact sizes.
//   k_gz_maps,    the window in front of chunk k is the resolved tail of chunk k-1: a serial chain, cut into
//   k_gz_bases    ~sqrt(chunks) groups (windows as maps of the group's base window, then the base windows in order).
//   k_gz_resolve  every symbol becomes a byte: out[p] = sym < 256 ? sym : window[sym - 256].  Parallel, HBM bound.
//   k_gz_crc      CRC-32 of every member from 2 KiB pieces: a piece's raw CRC is shifted to the member's end
//                 (multiplication by x^(8 * bytes behind it) mod P) and XORed into the member's accumulator.
//
// Integer / byte work only; bound by the serial Huffman chain of each chunk, not by HBM.
#include "fqz_inflate.cuh"

#define GZ_FULL 0xffffffffu
#define GZ_WARPS 4  // warps per CTA of k_gz_decode / k_gz_find
#define GZ_LL_BITS 10
#define GZ_D_BITS 8

static __constant__ u8 kGzClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---------------------------------------------------------------------------------- bit reader (LSB first)
// word i of the input; bytes behind n read as zero (the caller checks the consumed position against 8 n)
__device__ __forceinline__ u32 gz_word(const u32 *w, u64 n, u64 i) {
    u64 b = i * 4;
    if (b + 4 <= n) return w[i];
    if (b >= n) return 0u;
    return w[i] & (0xffffffffu >> (8u * (4u - (u32)(n - b))));
}
struct GzBits {
    const u32 *w;
    u64 n, wi, bb;
    u32 bc, nxt;  // nxt = word wi, loaded one refill ahead so that its latency hides behind the symbols in between
    __device__ __forceinline__ void seek(u64 bit) {
        wi = bit >> 5;
        u32 sh = (u32)(bit & 31u);
        bb = (u64)gz_word(w, n, wi) >> sh;
        bc = 32u - sh;
        wi++;
        nxt = gz_word(w, n, wi);
    }
    // at least 33 bits afterwards
    __device__ __forceinline__ void refill() {
        if (bc <= 32u) {
            bb |= (u64)nxt << bc;
            bc += 32u;
            wi++;
            nxt = gz_word(w, n, wi);
        }
    }
    __device__ __forceinline__ u32 peek(u32 k) const { return (u32)bb & ((1u << k) - 1u); }
    __device__ __forceinline__ void drop(u32 k) {
        bb >>= k;
        bc -= k;
    }
    __device__ __forceinline__ u32 take(u32 k) {
        u32 v = peek(k);
        drop(k);
        return v;
    }
    __device__ __forceinline__ u64 pos() const { return wi * 32u - bc; }
};

// canonical Huffman decode of one code taken LSB-first from `bits` (codes are packed starting with their most
// significant bit).  count[l] = codes of length l, symbol[] = symbols ordered by (length, value).  -1: no code.
__device__ __forceinline__ int gz_canon(const u32 *count, const u16 *symbol, u32 bits, u32 maxbits, u32 *len_out) {
    int code = 0, first = 0, index = 0;
    for (u32 len = 1; len <= maxbits; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        int cnt = (int)count[len];
        if (code - cnt < first) {
            *len_out = len;
            return (int)symbol[index + (code - first)];
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Warp-cooperative table build from code lengths.  Accepts what Go's huffmanDecoder.init accepts: an empty
// code, a complete code, or one single code of length 1 (compress/flate/inflate.go, "degenerate single-code").
// lut[e] = symbol << 4 | length for codes of at most lutbits bits, 0 otherwise.
__device__ bool gz_build(const u8 *lens, u32 n, u32 *count, u32 *start, u16 *symbol, u16 *lut, u32 lutbits) {
    const u32 lane = lane_id();
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (u32 i = lane; i < n; i += 32) {
        u32 l = lens[i];
        if (l) atomicAdd(&count[l], 1u);
    }
    __syncwarp();
    u32 ok = 1;
    if (lane == 0) {
        u32 kraft = 0, used = 0, off = 0;
        for (u32 l = 1; l <= 15; l++) {
            start[l] = off;
            off += count[l];
            used += count[l];
            kraft += count[l] << (15 - l);
        }
        start[0] = 0;
        ok = (used == 0 || kraft == 32768u || (used == 1 && count[1] == 1)) ? 1u : 0u;
    }
    ok = __shfl_sync(GZ_FULL, ok, 0);
    if (!ok) return false;
    __syncwarp();
    for (u32 base = 0; base < n; base += 32) {  // symbols of equal length keep their order
        u32 i = base + lane;
        u32 l = i < n ? lens[i] : 0u;
        u32 m = __match_any_sync(GZ_FULL, l);
        u32 rank = __popc(m & ((1u << lane) - 1u));
        if (l) symbol[start[l] + rank] = (u16)i;
        __syncwarp();
        if (l && rank == 0) start[l] += __popc(m);
        __syncwarp();
    }
    for (u32 e = lane; e < (1u << lutbits); e += 32) {
        u32 len = 0;
        int s = gz_canon(count, symbol, e, lutbits, &len);
        lut[e] = s < 0 ? (u16)0 : (u16)(((u32)s << 4) | len);
    }
    __syncwarp();
    return true;
}

// ---------------------------------------------------------------------------------- gzip member header
// Follows compress/gzip/gunzip.go readHeader: magic 1f 8b, method 8, FEXTRA / FNAME / FCOMMENT / FHCRC; strings of
// 512 bytes or more are ErrHeader; the header CRC (low 16 bits of the CRC-32 of everything before it) is checked.
__device__ u32 gz_crc_bytes(u32 crc, const u8 *p, u64 n) {  // bitwise CRC-32 update (pre / post inversion by the caller)
    for (u64 i = 0; i < n; i++) {
        crc ^= p[i];
        for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ (0xedb88320u & (0u - (crc & 1u)));
    }
    return crc;
}
// returns 0 and *data = offset of the deflate data, or a GZ_ST_ERR_* code with *data = offset of the error
__device__ u32 gz_member_header(const u8 *s, u64 n, u64 p, u64 *data) {
    *data = p;
    if (n - p < 10) {
        *data = n;
        return GZ_ST_ERR_TRUNC;
    }
    if (s[p] != 0x1f || s[p + 1] != 0x8b || s[p + 2] != 8) return GZ_ST_ERR_HEADER;
    const u32 flg = s[p + 3];
    u64 q = p + 10;
    if (flg & 4u) {  // FEXTRA
        if (n - q < 2) {
            *data = n;
            return GZ_ST_ERR_TRUNC;
        }
        u32 xlen = (u32)s[q] | ((u32)s[q + 1] << 8);
        q += 2;
        if (n - q < xlen) {
            *data = n;
            return GZ_ST_ERR_TRUNC;
        }
        q += xlen;
    }
    for (u32 bit = 8u; bit <= 16u; bit <<= 1) {  // FNAME, FCOMMENT: zero-terminated
        if (!(flg & bit)) continue;
        for (u32 i = 0;; i++) {
            if (i >= 512u) {
                *data = q;
                return GZ_ST_ERR_HEADER;
            }
            if (q >= n) {
                *data = n;
                return GZ_ST_ERR_TRUNC;
            }
            if (s[q++] == 0) break;
        }
    }
    if (flg & 2u) {  // FHCRC
        if (n - q < 2) {
            *data = n;
            return GZ_ST_ERR_TRUNC;
        }
        u32 want = (u32)s[q] | ((u32)s[q + 1] << 8);
        u32 got = ~gz_crc_bytes(0xffffffffu, s + p, q - p) & 0xffffu;
        if (want != got) {
            *data = q;
            return GZ_ST_ERR_HEADER;
        }
        q += 2;
    }
    *data = q;
    return 0;
}

// ---------------------------------------------------------------------------------- k_gz_find
// Is `bit` the header of a non-final dynamic-Huffman block?  Every test a real header passes: BFINAL 0, BTYPE 2,
// at most 286 / 30 codes, a complete code-length code, lengths that decode without overrun, a complete
// literal/length code with an end-of-block symbol, a complete (or empty / single) distance code.
__device__ bool gz_probe_dynamic(const u32 *w, u64 n, u64 bit) {
    const u64 nbits = n * 8;
    if (bit + 17 + 12 > nbits) return false;
    GzBits br;
    br.w = w;
    br.n = n;
    br.seek(bit);
    br.refill();
    u32 v = br.peek(17);
    if ((v & 7u) != 4u) return false;
    const u32 hlit = (v >> 3) & 31u, hdist = (v >> 8) & 31u, hclen = ((v >> 13) & 15u) + 4u;
    if (hlit > 29u || hdist > 29u) return false;
    br.drop(17);
    u8 cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    u32 kraft = 0;
    for (u32 i = 0; i < hclen; i++) {
        br.refill();
        u32 l = br.take(3);
        cl[kGzClOrder[i]] = (u8)l;
        if (l) kraft += 128u >> l;
        if (kraft > 128u) return false;  // over-subscribed already
    }
    if (kraft != 128u) return false;
    // 7-bit lookup of the code-length code: length in bits 0-2, symbol in bits 3-7
    u8 lut[128];
    {
        u32 code = 0;
        for (u32 l = 1; l <= 7; l++) {
            for (u32 s = 0; s < 19; s++) {
                if (cl[s] != l) continue;
                u32 r = __brev(code) >> (32 - l);
                for (u32 e = r; e < 128u; e += 1u << l) lut[e] = (u8)((s << 3) | l);
                code++;
            }
            code <<= 1;
        }
    }
    const u32 nlit = hlit + 257u, nsym = nlit + hdist + 1u;
    u32 i = 0, prev = 0, kraft_ll = 0, kraft_d = 0, nz_d = 0, one_d = 0;
    bool eob = false;
    while (i < nsym) {
        br.refill();
        u32 e = lut[br.peek(7)];
        br.drop(e & 7u);
        u32 s = e >> 3, len, rep;
        if (s < 16u) {
            len = s;
            rep = 1;
            prev = s;
        } else if (s == 16u) {
            if (i == 0) return false;
            len = prev;
            rep = 3u + br.take(2);
        } else if (s == 17u) {
            len = 0;
            rep = 3u + br.take(3);
            prev = 0;
        } else {
            len = 0;
            rep = 11u + br.take(7);
            prev = 0;
        }
        if (i + rep > nsym) return false;
        if (len) {
            u32 in_ll = (i < nlit) ? (min(i + rep, nlit) - i) : 0u;
            u32 in_d = rep - in_ll;
            kraft_ll += in_ll * (32768u >> len);
            kraft_d += in_d * (32768u >> len);
            nz_d += in_d;
            if (len == 1) one_d += in_d;
            if (i <= 256u && 256u < i + rep) eob = true;
        }
        i += rep;
    }
    if (br.pos() > nbits) return false;
    if (kraft_ll != 32768u || !eob) return false;
    return kraft_d == 32768u || nz_d == 0 || (nz_d == 1 && one_d == 1);
}

// BGZF member header (SAM spec 4.1): 1f 8b 08 04, XLEN 6, subfield 'B' 'C' of length 2
__device__ __forceinline__ bool gz_probe_bgzf(const u8 *s, u64 n, u64 p) {
    if (p + 18 > n) return false;
    return ld_u32_unaligned(s + p) == 0x04088b1fu && ld_u16_unaligned(s + p + 10) == 6u && ld_u32_unaligned(s + p + 12) == 0x00024342u;
}

__global__ void __launch_bounds__(GZ_WARPS * 32) k_gz_find(GzArgs a) {
    const u32 lane = lane_id();
    const u32 k = 1u + blockIdx.x * GZ_WARPS + (threadIdx.x >> 5);
    if (k >= a.nchunks) return;
    const u8 *s8 = (const u8 *)a.src;
    const u64 lo = (u64)k * a.chunk_bytes, hi = min(a.n, lo + a.chunk_bytes);
    u64 best = GZ_NO_TARGET;
    u32 type = GZ_AT_NONE;
    for (u64 base = lo; base < hi; base += 32) {  // member headers: byte aligned
        u64 p = base + lane;
        bool ok = p < hi && gz_probe_bgzf(s8, a.n, p);
        u32 m = __ballot_sync(GZ_FULL, ok);
        if (m) {
            best = (base + (u32)(__ffs((int)m) - 1)) * 8;
            type = GZ_AT_MEMBER;
            break;
        }
    }
    if (!a.bgzf_only) {
        // block headers: any bit.  1024 positions per step, 32 contiguous ones per lane: first the 13 header bits that
        // every lane can test without diverging (BFINAL 0, BTYPE 2, HLIT <= 29, HDIST <= 29: one position in nine
        // survives), then the full probe of the survivors
        const u64 lim = min(best, hi * 8);
        for (u64 base = lo * 8; base < lim; base += 1024) {
            const u64 p0 = base + (u64)lane * 32u;
            const u64 wi = p0 >> 5;
            const u32 sh = (u32)(p0 & 31u);
            const u32 w0 = gz_word(a.src, a.n, wi), w1 = gz_word(a.src, a.n, wi + 1), w2 = gz_word(a.src, a.n, wi + 2);
            const u64 v = ((u64)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);  // bits p0 .. p0 + 63
            u32 cand = 0;
#pragma unroll
            for (u32 j = 0; j < 32; j++) {
                const u32 x = (u32)(v >> j);
                const bool ok = (x & 7u) == 4u && ((x >> 3) & 31u) <= 29u && ((x >> 8) & 31u) <= 29u;
                cand |= (ok ? 1u : 0u) << j;
            }
            if (p0 >= lim)
                cand = 0;
            else if (lim - p0 < 32)
                cand &= (1u << (u32)(lim - p0)) - 1u;
            u32 hit = 32;
            while (cand) {
                const u32 j = (u32)__ffs((int)cand) - 1u;
                cand &= cand - 1u;
                if (gz_probe_dynamic(a.src, a.n, p0 + j)) {
                    hit = j;
                    break;
                }
            }
            const u32 m = __ballot_sync(GZ_FULL, hit < 32u);
            if (m) {  // lanes hold ascending positions: the first lane with a hit holds the first header
                const int l = __ffs((int)m) - 1;
                best = base + (u64)l * 32u + __shfl_sync(GZ_FULL, hit, l);
                type = GZ_AT_BLOCK;
                break;
            }
        }
    }
    if (lane == 0) {
        a.chunks[k].start_bit = best;
        a.chunks[k].start_type = type;
    }
}

// ---------------------------------------------------------------------------------- k_gz_decode
struct GzSm {
    u16 lut_ll[1u << GZ_LL_BITS];
    u16 lut_d[1u << GZ_D_BITS];
    u16 sym_ll[288];
    u16 sym_d[32];
    u32 cnt_ll[16], cnt_d[16], start[16];
    u32 q[32];
    u8 lens[320];
    u8 cl[32];
};

// lane 0: the header of a dynamic block up to the code lengths of the code-length code (S.cl)
__device__ u32 gz_dyn_counts(GzBits &br, GzSm &S, u32 *nlit, u32 *ndist) {
    br.refill();
    *nlit = br.take(5) + 257u;
    *ndist = br.take(5) + 1u;
    u32 hclen = br.take(4) + 4u;
    if (*nlit > 286u || *ndist > 30u) return GZ_ST_ERR_CORRUPT;
    for (u32 i = 0; i < 19; i++) S.cl[i] = 0;
    for (u32 i = 0; i < hclen; i++) {
        br.refill();
        S.cl[kGzClOrder[i]] = (u8)br.take(3);
    }
    return 0;
}
// lane 0: the nlit + ndist code lengths, coded with the code-length code whose tables sit in the distance slots
__device__ u32 gz_dyn_lengths(GzBits &br, GzSm &S, u32 total) {
    u32 i = 0;
    while (i < total) {
        br.refill();
        u32 e = S.lut_d[br.peek(7)];
        if (!e) return GZ_ST_ERR_CORRUPT;
        br.drop(e & 15u);
        u32 s = e >> 4;
        if (s < 16u) {
            S.lens[i++] = (u8)s;
            continue;
        }
        u32 rep, val = 0;
        if (s == 16u) {
            if (i == 0) return GZ_ST_ERR_CORRUPT;
            rep = 3u + br.take(2);
            val = S.lens[i - 1];
        } else if (s == 17u)
            rep = 3u + br.take(3);
        else
            rep = 11u + br.take(7);
        if (i + rep > total) return GZ_ST_ERR_CORRUPT;
        for (u32 k = 0; k < rep; k++) S.lens[i++] = (u8)val;
    }
    return 0;
}

// Whole warp: up to 32 tokens of a compressed block from bit *pos on into S.q (literal: value; match: length << 16 |
// distance).  A single thread walks a DEFLATE stream at ~400 cycles per token on this machine, so the walk is spread
// over the lanes: lane i decodes the complete token that WOULD start at bit *pos + i (literal/length code, extra
// bits, distance code, extra bits: at most 48 bits), then the real chain is followed through the 32 candidates with
// one shuffle per token — the lanes that were not token starts decoded garbage that nobody reads.
// Returns the count; *flags bit 0 = end of block, bits 8.. = error status; *pos is moved behind what was taken.
__device__ u32 gz_tokens(const u32 *w, u64 n, GzSm &S, u64 nbits, u64 *pos, u32 *flags) {
    const u32 lane = lane_id();
    u64 P = *pos;
    u32 cnt = 0, fl = 0;
    while (cnt < 32u && !fl) {
        const u64 off = P + lane;
        const u64 wi = off >> 5;
        const u32 sh = (u32)(off & 31u);
        u32 w0, w1, w2;
        if ((P >> 3) + 20 <= n) {  // the words of every lane lie inside the input (the same test in all lanes)
            w0 = w[wi];
            w1 = w[wi + 1];
            w2 = w[wi + 2];
        } else {
            w0 = gz_word(w, n, wi);
            w1 = gz_word(w, n, wi + 1);
            w2 = gz_word(w, n, wi + 2);
        }
        u64 v = ((u64)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);  // bits off .. off + 63
        const u64 left = nbits > P ? nbits - P : 0ull;
        const u32 room = left > 0xffffu ? 0xffffu : (u32)left;  // bits of input behind P, as far as this step can care
        u32 used, tok = 0, kind = 0;  // kind: 0 literal / match, 1 end of block, 2 no such code
        {
            u32 e = S.lut_ll[(u32)v & ((1u << GZ_LL_BITS) - 1u)], sym = 0, len = 0;
            if (e) {
                sym = e >> 4;
                len = e & 15u;
            } else {
                int c = gz_canon(S.cnt_ll, S.sym_ll, (u32)v & 0x7fffu, 15, &len);
                if (c < 0) {
                    kind = 2;
                    len = 1;
                } else
                    sym = (u32)c;
            }
            v >>= len;
            used = len;
            if (kind == 0) {
                if (sym < 256u)
                    tok = sym;
                else if (sym == 256u)
                    kind = 1;
                else if (sym >= 286u)
                    kind = 2;
                else {
                    // length: 257..264 -> 3..10; 265..284 -> groups of four with 1..5 extra bits; 285 -> 258
                    u32 eb = 0, ml = sym - 254u;
                    if (sym == 285u)
                        ml = 258u;
                    else if (sym >= 265u) {
                        eb = (sym - 261u) >> 2;
                        ml = ((4u + ((sym - 261u) & 3u)) << eb) + 3u + ((u32)v & ((1u << eb) - 1u));
                    }
                    v >>= eb;
                    used += eb;
                    u32 d = S.lut_d[(u32)v & ((1u << GZ_D_BITS) - 1u)], ds = 0, dl = 0;
                    if (d) {
                        ds = d >> 4;
                        dl = d & 15u;
                    } else {
                        int c = gz_canon(S.cnt_d, S.sym_d, (u32)v & 0x7fffu, 15, &dl);
                        if (c < 0) {
                            kind = 2;
                            dl = 1;
                        } else
                            ds = (u32)c;
                    }
                    v >>= dl;
                    used += dl;
                    if (ds >= 30u) kind = 2;
                    // distance: 0..3 -> 1..4; then pairs with 1..13 extra bits
                    u32 de = 0, dist = ds + 1u;
                    if (ds >= 4u && ds < 30u) {
                        de = (ds >> 1) - 1u;
                        dist = ((2u + (ds & 1u)) << de) + 1u + ((u32)v & ((1u << de) - 1u));
                    }
                    used += de;
                    tok = (ml << 16) | dist;
                }
            }
        }
        u32 cur = 0;  // the same in every lane
        while (cur < 32u && cnt < 32u) {
            const u32 u = __shfl_sync(GZ_FULL, used, (int)cur), k = __shfl_sync(GZ_FULL, kind, (int)cur);
            if (cur + u > room) {  // Go runs out of input before it can see what the zero padding decodes to
                fl = GZ_ST_ERR_TRUNC << 8;
                break;
            }
            if (k == 2u) {
                fl = GZ_ST_ERR_CORRUPT << 8;
                break;
            }
            if (k == 1u) {
                cur += u;
                fl = 1;
                break;
            }
            if (lane == cur) S.q[cnt] = tok;
            cnt++;
            cur += u;
        }
        P += cur;
    }
    *pos = P;
    *flags = fl;
    return cnt;
}

template <bool WRITE> __global__ void __launch_bounds__(GZ_WARPS * 32) k_gz_decode(GzArgs a) {
    __shared__ GzSm smem[GZ_WARPS];
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 li = blockIdx.x * GZ_WARPS + wid;
    if (li >= a.nlist) return;
    GzSm &S = smem[wid];
    GzChunk &C = a.chunks[a.list[li]];
    const u8 *s8 = (const u8 *)a.src;
    const u64 n = a.n, nbits = a.n * 8;
    const u64 target = C.target_bit;
    const u32 target_type = C.target_type;
    u16 *out = WRITE ? (u16 *)(uintptr_t)C.sym_ptr : nullptr;
    GzMember *mem = WRITE ? (GzMember *)(uintptr_t)C.mem_ptr : nullptr;
    const u64 sym_cap = C.sym_cap;
    const u32 mem_cap = C.mem_cap;
    bool over = false;  // out of room (speculative pass): counting goes on, writing stops

    GzBits br;
    br.w = a.src;
    br.n = n;
    br.wi = 0;
    br.bb = 0;
    br.bc = 0;
    br.nxt = 0;
    u64 pos = C.start_bit;  // bit position at member / block boundaries, the same in every lane
    u64 produced = 0;       // bytes decoded by this chunk
    bool known = false;     // the window in front of the current position lies inside this chunk's own output
    u64 member_from = 0;    // ... and starts here
    bool at_header = C.start_type == GZ_AT_MEMBER;
    u32 nmem = 0, status = 0;
    u64 errbit = 0;
    if (!at_header && lane == 0) br.seek(pos);

    for (;;) {
        if (pos > target) {
            status = GZ_ST_OVERSHOOT;
            break;
        }
        if (at_header) {
            if (pos == target && target_type == GZ_AT_MEMBER) {
                status = GZ_ST_REACHED;
                break;
            }
            if ((pos >> 3) == n) {
                status = GZ_ST_END;
                break;
            }
            u32 rc = 0;
            u64 data = 0;
            if (lane == 0) rc = gz_member_header(s8, n, pos >> 3, &data);
            rc = __shfl_sync(GZ_FULL, rc, 0);
            data = __shfl_sync(GZ_FULL, data, 0);
            if (rc) {
                status = rc;
                errbit = data * 8;
                break;
            }
            known = true;
            member_from = produced;
            pos = data * 8;
            if (lane == 0) br.seek(pos);
            at_header = false;
            continue;
        }
        if (pos == target && target_type == GZ_AT_BLOCK) {
            status = GZ_ST_REACHED;
            break;
        }
        // ---- block header
        u32 hdr = 0;
        if (lane == 0) {
            br.refill();
            hdr = br.take(3);
            if (br.pos() > nbits) hdr = 0x100u;
        }
        hdr = __shfl_sync(GZ_FULL, hdr, 0);
        if (hdr & 0x100u) {
            status = GZ_ST_ERR_TRUNC;
            errbit = nbits;
            break;
        }
        const u32 bfinal = hdr & 1u, btype = hdr >> 1;
        if (btype == 3u) {
            status = GZ_ST_ERR_CORRUPT;
            errbit = pos;
            break;
        }
        if (btype == 0u) {  // stored: LEN, ~LEN, bytes
            u64 at = 0;
            if (lane == 0) at = (br.pos() + 7u) >> 3;
            at = __shfl_sync(GZ_FULL, at, 0);
            if (at + 4 > n) {
                status = GZ_ST_ERR_TRUNC;
                errbit = nbits;
                break;
            }
            const u32 len = (u32)s8[at] | ((u32)s8[at + 1] << 8), nlen = (u32)s8[at + 2] | ((u32)s8[at + 3] << 8);
            if ((len ^ nlen) != 0xffffu) {
                status = GZ_ST_ERR_CORRUPT;
                errbit = at * 8;
                break;
            }
            if (at + 4 + len > n) {
                status = GZ_ST_ERR_TRUNC;
                errbit = nbits;
                break;
            }
            if (WRITE && produced + len > sym_cap) over = true;
            if (WRITE && !over)
                for (u32 t = lane; t < len; t += 32) out[produced + t] = (u16)s8[at + 4 + t];
            produced += len;
            pos = (at + 4 + len) * 8;
            if (lane == 0) br.seek(pos);
            __syncwarp();
        } else {
            // ---- code tables
            u32 rc = 0, nlit = 288, ndist = 32;
            if (btype == 1u) {
                for (u32 i = lane; i < 288u; i += 32) S.lens[i] = (u8)(i < 144u ? 8 : (i < 256u ? 9 : (i < 280u ? 7 : 8)));
                S.lens[288u + lane] = 5;
                __syncwarp();
            } else {
                if (lane == 0) {
                    rc = gz_dyn_counts(br, S, &nlit, &ndist);
                    if (br.pos() > nbits) rc = GZ_ST_ERR_TRUNC;
                }
                rc = __shfl_sync(GZ_FULL, rc, 0);
                nlit = __shfl_sync(GZ_FULL, nlit, 0);
                ndist = __shfl_sync(GZ_FULL, ndist, 0);
                if (!rc && !gz_build(S.cl, 19, S.cnt_d, S.start, S.sym_d, S.lut_d, 7)) rc = GZ_ST_ERR_CORRUPT;
                if (!rc) {
                    if (lane == 0) {
                        rc = gz_dyn_lengths(br, S, nlit + ndist);
                        if (br.pos() > nbits) rc = GZ_ST_ERR_TRUNC;
                    }
                    rc = __shfl_sync(GZ_FULL, rc, 0);
                }
            }
            if (!rc && !gz_build(S.lens, nlit, S.cnt_ll, S.start, S.sym_ll, S.lut_ll, GZ_LL_BITS)) rc = GZ_ST_ERR_CORRUPT;
            if (!rc && !gz_build(S.lens + nlit, ndist, S.cnt_d, S.start, S.sym_d, S.lut_d, GZ_D_BITS)) rc = GZ_ST_ERR_CORRUPT;
            if (rc) {
                status = rc;
                errbit = rc == GZ_ST_ERR_TRUNC ? nbits : pos;
                break;
            }
            // ---- tokens, up to 32 at a time
            {
                u64 p = 0;
                if (lane == 0) p = br.pos();
                pos = __shfl_sync(GZ_FULL, p, 0);
            }
            for (;;) {
                u32 fl = 0;
                u32 cnt = gz_tokens(a.src, n, S, nbits, &pos, &fl);
                __syncwarp();
                const u32 e = lane < cnt ? S.q[lane] : 0u;
                u32 mlen = e >> 16;
                const u32 dist = e & 0xffffu;
                u32 len = lane < cnt ? (mlen ? mlen : 1u) : 0u;
                u32 incl = group_incl_scan(len, GZ_FULL, 32);
                const u32 o = incl - len;
                // a match that reaches in front of its member (Go: dist > histSize): the step ends in front of it
                const u64 avail = known ? produced - member_from : produced + GZ_WINDOW;
                const u32 bad = __ballot_sync(GZ_FULL, mlen != 0u && (u64)dist > avail + o);
                if (bad) {
                    cnt = (u32)__ffs((int)bad) - 1u;
                    fl = GZ_ST_ERR_CORRUPT << 8;
                    if (lane >= cnt) {
                        len = 0;
                        mlen = 0;
                    }
                    incl = group_incl_scan(len, GZ_FULL, 32);
                }
                const u32 bytes = __shfl_sync(GZ_FULL, incl, 31);
                if (WRITE && produced + bytes > sym_cap) over = true;
                if (WRITE && !over) {
                    if (lane < cnt && !mlen) out[produced + o] = (u16)e;
                    // short matches whose source lies in front of this step do not depend on anything the step
                    // writes: every lane copies its own, four loads in flight per round trip to L2
                    const bool indep = mlen != 0u && mlen <= 32u && dist >= o + mlen;
                    if (indep) {
                        const long long s0 = (long long)(produced + o) - (long long)dist;
                        u16 *dp = out + produced + o;
                        for (u32 t = 0; t < mlen; t += 4) {
                            u16 v[4];
#pragma unroll
                            for (u32 k = 0; k < 4; k++) {
                                const long long s = s0 + t + k;
                                v[k] = (t + k < mlen) ? (s >= 0 ? out[s] : (u16)(256 + (long long)GZ_WINDOW + s)) : (u16)0;
                            }
#pragma unroll
                            for (u32 k = 0; k < 4; k++)
                                if (t + k < mlen) dp[t + k] = v[k];
                        }
                    }
                    __syncwarp();
                    u32 mm = __ballot_sync(GZ_FULL, mlen != 0u && !indep);
                    while (mm) {
                        const int j = __ffs((int)mm) - 1;
                        mm &= mm - 1u;
                        const u32 L = __shfl_sync(GZ_FULL, mlen, j), D = __shfl_sync(GZ_FULL, dist, j);
                        const u64 dst = produced + __shfl_sync(GZ_FULL, o, j);
                        for (u32 t = lane; t < L; t += 32) {
                            const u32 tt = t < D ? t : t % D;  // overlapping matches repeat with period D
                            const long long src = (long long)dst + (long long)tt - (long long)D;
                            out[dst + t] = src >= 0 ? out[src] : (u16)(256 + (long long)GZ_WINDOW + src);
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();  // S.q is rewritten by the next step
                produced += bytes;
                if (fl >> 8) {
                    status = fl >> 8;
                    break;
                }
                if (fl & 1u) break;
            }
            if (status) {
                errbit = status == GZ_ST_ERR_TRUNC ? nbits : pos;
                break;
            }
            if (lane == 0) br.seek(pos);
        }
        if (bfinal) {  // trailer: CRC-32, ISIZE
            const u64 at = (pos + 7u) >> 3;
            if (at + 8 > n) {
                status = GZ_ST_ERR_TRUNC;
                errbit = nbits;
                break;
            }
            if (WRITE && nmem >= mem_cap) over = true;
            if (WRITE && !over && lane == 0) {
                GzMember m;
                m.out_end = produced;  // relative to the chunk; k_gz_members adds its place
                m.trailer_byte = at;
                m.crc = (u32)s8[at] | ((u32)s8[at + 1] << 8) | ((u32)s8[at + 2] << 16) | ((u32)s8[at + 3] << 24);
                m.isize = (u32)s8[at + 4] | ((u32)s8[at + 5] << 8) | ((u32)s8[at + 6] << 16) | ((u32)s8[at + 7] << 24);
                m.crc_acc = 0;
                m.pad = 0;
                mem[nmem] = m;
            }
            nmem++;
            pos = (at + 8) * 8;
            at_header = true;
            known = false;
        }
    }
    if (lane == 0) {
        if (WRITE && !a.spec) {
            if (status < 16u && (over || produced != C.out_len || nmem != C.members || pos != C.end_bit)) {
                status = GZ_ST_ERR_INTERNAL;
                errbit = C.start_bit;
            }
            if (status >= 16u) atomicMin(a.err, (unsigned long long)((errbit << 8) | status));
        } else {
            C.overflow = over ? 1u : 0u;
            C.out_len = produced;
            C.end_bit = pos;
            C.err_bit = errbit;
            C.members = nmem;
            C.after_member = known ? produced - member_from : ~0ull;
            C.status = status;
        }
    }
}

// ---------------------------------------------------------------------------------- k_gz_maps, k_gz_bases, k_gz_resolve
// The window in front of a chunk that starts at a block header is the last 32 KiB in front of it: the tail of the
// previous chunk, with ITS markers resolved through ITS window, behind what is left of that window — one serial chain
// over all chunks.  It is cut into groups of ~sqrt(chunks): inside a group every window is kept as a MAP of the
// group's base window (entry < 256: a byte; else 256 + index into the base window), built chunk after chunk by one CTA
// per group (k_gz_maps); the base windows are then resolved group after group by one CTA (k_gz_bases); k_gz_resolve
// looks a marker up through its chunk's map and its group's base window.  2 sqrt(chunks) serial steps instead of chunks.
#define GZ_WIN_THREADS 1024u
#define GZ_WIN_BATCH 8u
struct GzWinStep {
    u32 type;
    u64 psym, plen;
};
__device__ __forceinline__ GzWinStep gz_win_step(const GzArgs &a, u32 i) {
    GzWinStep s;
    s.type = a.chunks[a.list[i]].start_type;
    const GzChunk &P = a.chunks[a.list[i ? i - 1 : 0]];
    s.psym = P.sym_ptr;
    s.plen = P.out_len;
    return s;
}
// element t of the window behind the previous chunk (tail = its last symbols, plen of them if fewer than 32 Ki):
// a symbol of that chunk, or — 256 + index — a byte of the previous chunk's own window
__device__ __forceinline__ u32 gz_win_src(const u16 *tail, u32 keep, u32 plen, u32 t) {
    return t < keep ? 256u + t + plen : (u32)tail[t - keep];
}
__global__ void __launch_bounds__(GZ_WIN_THREADS) k_gz_maps(GzArgs a) {
    const u32 first = blockIdx.x * a.gsize, last = min(a.nlist, first + a.gsize);
    GzWinStep nx = gz_win_step(a, first);
    for (u32 i = first; i < last; i++) {
        const GzWinStep cur = nx;
        if (i + 1 < last) nx = gz_win_step(a, i + 1);
        u16 *M = a.maps + (size_t)i * GZ_WINDOW;
        if (cur.type != GZ_AT_BLOCK) {  // starts at a member header: nothing in front of it can be referenced
            for (u32 t = threadIdx.x; t < GZ_WINDOW; t += GZ_WIN_THREADS) M[t] = 0;
        } else if (i == first) {  // its window IS the group's base window
            for (u32 t = threadIdx.x; t < GZ_WINDOW; t += GZ_WIN_THREADS) M[t] = (u16)(256u + t);
        } else {
            const u16 *Mp = M - GZ_WINDOW;
            const u32 keep = cur.plen >= GZ_WINDOW ? 0u : GZ_WINDOW - (u32)cur.plen;
            const u16 *tail = (const u16 *)(uintptr_t)cur.psym + (cur.plen >= GZ_WINDOW ? cur.plen - GZ_WINDOW : 0u);
            for (u32 b = 0; b < GZ_WINDOW; b += GZ_WIN_BATCH * GZ_WIN_THREADS) {
                u32 s[GZ_WIN_BATCH];
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++) s[k] = gz_win_src(tail, keep, (u32)min(cur.plen, (u64)GZ_WINDOW), b + k * GZ_WIN_THREADS + threadIdx.x);
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++)
                    if (s[k] >= 256u) s[k] = Mp[s[k] - 256u];
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++) M[b + k * GZ_WIN_THREADS + threadIdx.x] = (u16)s[k];
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(GZ_WIN_THREADS) k_gz_bases(GzArgs a) {
    const u32 ngroups = (a.nlist + a.gsize - 1) / a.gsize;
    for (u32 t = threadIdx.x; t < GZ_WINDOW; t += GZ_WIN_THREADS) a.win[t] = 0;  // group 0 starts the file
    __syncthreads();
    GzWinStep nx = gz_win_step(a, min(a.gsize, a.nlist - 1));
    for (u32 g = 1; g < ngroups; g++) {
        const u32 first = g * a.gsize;
        const GzWinStep cur = nx;
        if (g + 1 < ngroups) nx = gz_win_step(a, first + a.gsize);
        u8 *B = a.win + (size_t)g * GZ_WINDOW;
        if (cur.type != GZ_AT_BLOCK) {
            for (u32 t = threadIdx.x; t < GZ_WINDOW; t += GZ_WIN_THREADS) B[t] = 0;
        } else {
            const u8 *Bp = B - GZ_WINDOW;
            const u16 *Mp = a.maps + (size_t)(first - 1) * GZ_WINDOW;
            const u32 keep = cur.plen >= GZ_WINDOW ? 0u : GZ_WINDOW - (u32)cur.plen;
            const u16 *tail = (const u16 *)(uintptr_t)cur.psym + (cur.plen >= GZ_WINDOW ? cur.plen - GZ_WINDOW : 0u);
            for (u32 b = 0; b < GZ_WINDOW; b += GZ_WIN_BATCH * GZ_WIN_THREADS) {
                u32 s[GZ_WIN_BATCH];
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++) s[k] = gz_win_src(tail, keep, (u32)min(cur.plen, (u64)GZ_WINDOW), b + k * GZ_WIN_THREADS + threadIdx.x);
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++)
                    if (s[k] >= 256u) s[k] = Mp[s[k] - 256u];
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++)
                    if (s[k] >= 256u) s[k] = Bp[s[k] - 256u];
#pragma unroll
                for (u32 k = 0; k < GZ_WIN_BATCH; k++) B[b + k * GZ_WIN_THREADS + threadIdx.x] = (u8)s[k];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_gz_resolve(GzArgs a) {
    const u32 i = blockIdx.y;
    const GzChunk &C = a.chunks[a.list[i]];
    const u16 *sym = (const u16 *)(uintptr_t)C.sym_ptr;
    u8 *out = a.out + C.out_off;
    const bool has_window = C.start_type == GZ_AT_BLOCK;
    const u16 *M = has_window ? a.maps + (size_t)i * GZ_WINDOW : nullptr;
    const u8 *B = has_window ? a.win + (size_t)(i / a.gsize) * GZ_WINDOW : nullptr;
    const u32 first_valid = GZ_WINDOW - C.window_valid;
    bool bad = false;
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < C.out_len; p += (u64)gridDim.x * blockDim.x) {
        u32 s = sym[p];
        if (s >= 256u) {
            s -= 256u;
            if (!has_window || s < first_valid) {  // a match that reaches in front of its member
                bad = true;
                s = 0;
            } else {
                s = M[s];
                if (s >= 256u) s = B[s - 256u];
            }
        }
        out[p] = (u8)s;
    }
    if (bad) atomicMin(a.err, (unsigned long long)((C.start_bit << 8) | (has_window ? GZ_ST_ERR_CORRUPT : GZ_ST_ERR_INTERNAL)));
}

// ---------------------------------------------------------------------------------- k_gz_members
// the member records of every chunk, in stream order, into the file-wide table (out_end made absolute)
__global__ void __launch_bounds__(128) k_gz_members(GzArgs a) {
    const u32 i = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (i >= a.nlist) return;
    const GzChunk &C = a.chunks[a.list[i]];
    const GzMember *src = (const GzMember *)(uintptr_t)C.mem_ptr;
    for (u32 j = lane_id(); j < C.members; j += 32) {
        GzMember m = src[j];
        m.out_end += C.out_off;
        a.members[C.member_base + j] = m;
    }
}

// ---------------------------------------------------------------------------------- k_gz_crc
// reflected CRC-32 arithmetic: bit 31 is x^0 (zlib's multmodp)
__device__ __forceinline__ u32 gz_mulmod(u32 a, u32 b) {
    u32 p = 0;
    for (int i = 31; i >= 0; i--) {
        if (a & (1u << i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
// x^(8 * bytes) mod P
__device__ __forceinline__ u32 gz_xpow8(const u32 *pw, u64 bytes) {
    u32 r = 0x80000000u;
    for (u32 k = 0; bytes; k++, bytes >>= 1)
        if (bytes & 1u) r = gz_mulmod(pw[k], r);
    return r;
}
#define GZ_CRC_PIECE 2048u
__global__ void __launch_bounds__(128) k_gz_crc(GzArgs a, u64 out_len) {
    __shared__ u32 T[256];
    for (u32 i = threadIdx.x; i < 256u; i += blockDim.x) {
        u32 c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
        T[i] = c;
    }
    __syncthreads();
    const u64 piece = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 lo = piece * GZ_CRC_PIECE;
    if (lo >= out_len) return;
    const u64 hi = min(out_len, lo + GZ_CRC_PIECE);
    // first member that ends behind lo
    u32 m = 0, r = a.nmembers;
    while (m < r) {
        u32 mid = (m + r) >> 1;
        if (a.members[mid].out_end > lo)
            r = mid;
        else
            m = mid + 1;
    }
    while (lo < hi && m < a.nmembers) {
        const u64 mend = a.members[m].out_end;
        const u64 se = min(hi, mend);
        u32 c = 0;
        u64 p = lo;
        for (; p < se && (p & 3u); p++) c = T[(c ^ a.out[p]) & 0xffu] ^ (c >> 8);
        for (; p + 4 <= se; p += 4) {
            c ^= *(const u32 *)(a.out + p);
            c = T[c & 0xffu] ^ (c >> 8);
            c = T[c & 0xffu] ^ (c >> 8);
            c = T[c & 0xffu] ^ (c >> 8);
            c = T[c & 0xffu] ^ (c >> 8);
        }
        for (; p < se; p++) c = T[(c ^ a.out[p]) & 0xffu] ^ (c >> 8);
        if (c) atomicXor(&a.members[m].crc_acc, gz_mulmod(gz_xpow8(a.pw, mend - se), c));
        lo = se;
        if (lo < hi) {
            m++;
            while (m < a.nmembers && a.members[m].out_end <= lo) m++;
        }
    }
}
__global__ void __launch_bounds__(256) k_gz_verify(GzArgs a) {
    const u32 m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.nmembers) return;
    const GzMember &M = a.members[m];
    const u64 from = m ? a.members[m - 1].out_end : 0ull, len = M.out_end - from;
    const u32 crc = ~(gz_mulmod(gz_xpow8(a.pw, len), 0xffffffffu) ^ M.crc_acc);
    if (crc != M.crc || (u32)len != M.isize) atomicMin(a.err, (unsigned long long)(((M.trailer_byte * 8) << 8) | GZ_ST_ERR_CHECKSUM));
}

// ---------------------------------------------------------------------------------- launchers
void fqz_launch_gz_find(const GzArgs &a, cudaStream_t s) {
    if (a.nchunks < 2) return;
    u32 grid = (a.nchunks - 1 + GZ_WARPS - 1) / GZ_WARPS;
    FQZ_LAUNCH(k_gz_find, grid, GZ_WARPS * 32, 0, s, a);
}
void fqz_launch_gz_decode(const GzArgs &a, bool write, cudaStream_t s) {
    if (!a.nlist) return;
    u32 grid = (a.nlist + GZ_WARPS - 1) / GZ_WARPS;
    if (write)
        FQZ_LAUNCH(k_gz_decode<true>, grid, GZ_WARPS * 32, 0, s, a);
    else
        FQZ_LAUNCH(k_gz_decode<false>, grid, GZ_WARPS * 32, 0, s, a);
}
void fqz_launch_gz_members(const GzArgs &a, cudaStream_t s) {
    if (a.nlist && a.nmembers) FQZ_LAUNCH(k_gz_members, (a.nlist + 3) / 4, 128, 0, s, a);
}
void fqz_launch_gz_windows(const GzArgs &a, cudaStream_t s) {
    if (!a.maps) return;  // no chunk starts at a block header: nothing to resolve
    const u32 ngroups = (a.nlist + a.gsize - 1) / a.gsize;
    FQZ_LAUNCH(k_gz_maps, ngroups, GZ_WIN_THREADS, 0, s, a);
    FQZ_LAUNCH(k_gz_bases, 1, GZ_WIN_THREADS, 0, s, a);
}
void fqz_launch_gz_resolve(const GzArgs &a, cudaStream_t s) {
    if (!a.nlist) return;
    FQZ_LAUNCH(k_gz_resolve, dim3(16, a.nlist), 256, 0, s, a);
}
void fqz_launch_gz_crc(const GzArgs &a, u64 out_len, cudaStream_t s) {
    if (out_len) {
        u64 pieces = (out_len + GZ_CRC_PIECE - 1) / GZ_CRC_PIECE;
        FQZ_LAUNCH(k_gz_crc, (u32)((pieces + 127) / 128), 128, 0, s, a, out_len);
    }
    if (a.nmembers) FQZ_LAUNCH(k_gz_verify, (a.nmembers + 255) / 256, 256, 0, s, a);
}
