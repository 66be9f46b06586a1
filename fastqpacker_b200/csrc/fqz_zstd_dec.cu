// fqz_zstd_dec.cu — GPU zstd (RFC 8878) decoder: replaces zstd.Decoder.DecodeAll (reference call
// sites internal/compress/compress.go:785-814; decoder options :120-122).  Must accept everything
// the reference's encoder (klauspost/compress v1.19.1 SpeedFastest, go.mod:8, not in the tree) can
// emit — multi-block frames, raw / RLE / compressed blocks, 1- and 4-stream Huffman literals,
// treeless literals, predefined / RLE / FSE / repeat sequence tables, repeat offsets, matches that
// reach into earlier blocks, content checksums — as well as the small independent frames written by
// fqz_zstd_enc.cu.  No dictionaries (the reference uses none, compress.go:523-528).
//
// Pipeline over a batch of streams (each a chain of frames):
//   k_zd_hop        1 warp   / stream : frame + block positions (32 frames at a time behind an index frame)
//   k_zd_parse      1 thread / block  : literals header, sequence count, table positions
//   k_zd_link       1 thread / frame  : treeless / repeat provenance, output offsets known from headers
//   k_zd_offsets    1 thread / block  : arena offsets, table slots, literal groups (after the scans)
//   k_zd_literals   1 warp   / group  : <= 8 blocks of one frame, lane = (block, Huffman stream)
//   k_zd_seq_tables 1 warp   / block  : FSE decode tables into L2 (lanes 0-2)
//   k_zd_seq_decode 1 thread / block  : the serial FSE chain, 32 blocks per warp
//   k_zd_rawcopy    1 warp   / block  : raw / RLE blocks whose place is known from the headers
//   k_zd_execute    1 warp   / frame  : LZ77 copies (single-block frames staged in shared memory)
//   k_zd_checksum   4 lanes  / frame  : XXH64 of the regenerated content
#include "fqz_zstd.h"
#include "fqz_zstd_dec.h"
#include "fqz_zstd_tables.cuh"
#include "fqz_xxh64.cuh"

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------- backward bit reader
struct BitR {
    const u8 *start;
    const u8 *ptr;
    u64 cont;
    u32 consumed;  // bits consumed from the top of `cont`
    bool bad;
    __device__ static u64 load_le(const u8 *p, u32 n) {  // n <= 8 bytes
        if (n == 8) return (u64)ld_u32_unaligned(p) | ((u64)ld_u32_unaligned(p + 4) << 32);  // aligned words + funnel shift
        u64 v = 0;
        for (u32 i = 0; i < n; i++) v |= (u64)p[i] << (8 * i);
        return v;
    }
    __device__ void init(const u8 *s, u32 size) {
        start = s;
        bad = false;
        if (size == 0) {
            bad = true;
            ptr = s;
            cont = 0;
            consumed = 64;
            return;
        }
        u32 last = s[size - 1];
        if (last == 0) bad = true;  // the end mark is missing
        u32 hb = last ? hibit32(last) : 0;
        if (size >= 8) {
            ptr = s + size - 8;
            cont = load_le(ptr, 8);
            consumed = 8 - hb;
        } else {
            ptr = s;
            cont = load_le(s, size);
            consumed = 8 - hb + (8 - size) * 8;
        }
    }
    __device__ __forceinline__ u32 look(u32 n) const {  // n in 1..32; bits past the start read as 0
        if (consumed >= 64) return 0;
        return (u32)(((cont << consumed) >> 1) >> (63 - n));
    }
    __device__ __forceinline__ void skip(u32 n) { consumed += n; }
    __device__ __forceinline__ u32 read(u32 n) {
        if (n == 0) return 0;
        u32 v = look(n);
        skip(n);
        return v;
    }
    __device__ __forceinline__ void reload() {
        if (consumed > 64) {
            bad = true;  // read past the beginning of the stream
            consumed = 64;
            return;
        }
        if (ptr >= start + 8) {
            ptr -= consumed >> 3;
            consumed &= 7;
            cont = load_le(ptr, 8);
            return;
        }
        if (ptr == start) return;
        u32 nb = consumed >> 3;
        if (ptr - nb < start) nb = (u32)(ptr - start);
        ptr -= nb;
        consumed -= nb * 8;
        cont = load_le(ptr, 8);
    }
    __device__ bool finished() const { return ptr == start && consumed == 64; }
};

// ---------------------------------------------------------------------------------- FSE decode tables
struct FseDEnt {
    u16 base;  // new state base
    u8 sym;
    u8 nb;
};
// Reads an FSE table description (RFC 8878 §4.1.1).  Returns bytes consumed, 0 on error.
__device__ static u32 fse_read_ncount(const u8 *p, u32 avail, short *norm, u32 *maxSymIO, u32 *tlogOut, u32 maxLog) {
    if (avail < 1) return 0;
    u32 maxSym = *maxSymIO;
    u64 bits = 0;
    u32 have = 0, pos = 0;
#define NC_FILL()                                   \
    while (have <= 56 && pos < avail) {             \
        bits |= (u64)p[pos++] << have;              \
        have += 8;                                  \
    }
    NC_FILL();
    u32 tlog = (u32)(bits & 15) + 5;
    bits >>= 4;
    have -= 4;
    if (tlog > maxLog) return 0;
    int remaining = (1 << tlog) + 1, threshold = 1 << tlog;
    u32 nbBits = tlog + 1;
    u32 sym = 0;
    bool prev0 = false;
    while (remaining > 1 && sym <= maxSym) {
        NC_FILL();
        if (prev0) {
            u32 n0 = sym;
            for (;;) {
                NC_FILL();
                u32 r = (u32)(bits & 3);
                bits >>= 2;
                if (have < 2) return 0;
                have -= 2;
                n0 += r;
                if (r != 3) break;
            }
            if (n0 > maxSym + 1) return 0;
            while (sym < n0) norm[sym++] = 0;
            if (sym > maxSym) break;
            NC_FILL();
        }
        int mx = (2 * threshold - 1) - remaining;
        int count;
        if (have < nbBits) return 0;
        if ((int)(bits & (u32)(threshold - 1)) < mx) {
            count = (int)(bits & (u32)(threshold - 1));
            bits >>= (nbBits - 1);
            have -= nbBits - 1;
        } else {
            count = (int)(bits & (u32)(2 * threshold - 1));
            if (count >= threshold) count -= mx;
            bits >>= nbBits;
            have -= nbBits;
        }
        count--;
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (short)count;
        prev0 = (count == 0);
        while (remaining < threshold) {
            nbBits--;
            threshold >>= 1;
        }
    }
#undef NC_FILL
    if (remaining != 1) return 0;
    *maxSymIO = sym - 1;
    *tlogOut = tlog;
    // bytes consumed = everything pulled minus whole unused bytes
    u32 used = pos - (have >> 3);
    return used;
}
__device__ static void fse_build_dtable(const short *norm, u32 maxSym, u32 tlog, FseDEnt *dt, u16 *symNext /*>=maxSym+1*/) {
    u32 tsize = 1u << tlog, mask = tsize - 1, step = (tsize >> 1) + (tsize >> 3) + 3;
    u32 high = tsize - 1;
    for (u32 s = 0; s <= maxSym; s++) {
        if (norm[s] == -1) {
            dt[high--].sym = (u8)s;
            symNext[s] = 1;
        } else
            symNext[s] = (u16)norm[s];
    }
    u32 pos = 0;
    for (u32 s = 0; s <= maxSym; s++)
        for (int i = 0; i < norm[s]; i++) {
            dt[pos].sym = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (u32 u = 0; u < tsize; u++) {
        u32 s = dt[u].sym;
        u32 nx = symNext[s]++;
        u32 nb = tlog - hibit32(nx);
        dt[u].nb = (u8)nb;
        dt[u].base = (u16)((nx << nb) - tsize);
    }
}

// ---------------------------------------------------------------------------------- header parsing helpers
__device__ __forceinline__ u32 rd24(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16); }
__device__ __forceinline__ u32 rd32(const u8 *p) { return rd24(p) | ((u32)p[3] << 24); }

// Literals section header.  Returns header size (0 = corrupt).
__device__ static u32 parse_lit_header(const u8 *p, u32 avail, u32 *type, u32 *regen, u32 *csize, u32 *nstreams) {
    if (avail < 1) return 0;
    u32 b0 = p[0];
    u32 t = b0 & 3, fmt = (b0 >> 2) & 3;
    *type = t;
    *nstreams = 1;
    if (t < 2) {  // raw / RLE
        if (fmt == 0 || fmt == 2) {
            *regen = b0 >> 3;
            *csize = (t == 0) ? *regen : 1;
            return 1;
        }
        if (fmt == 1) {
            if (avail < 2) return 0;
            *regen = (b0 >> 4) | ((u32)p[1] << 4);
            *csize = (t == 0) ? *regen : 1;
            return 2;
        }
        if (avail < 3) return 0;
        *regen = (rd24(p) >> 4);
        *csize = (t == 0) ? *regen : 1;
        return 3;
    }
    if (fmt == 0 || fmt == 1) {
        if (avail < 3) return 0;
        u32 h = rd24(p);
        *regen = (h >> 4) & 0x3FF;
        *csize = (h >> 14) & 0x3FF;
        *nstreams = fmt == 0 ? 1 : 4;
        return 3;
    }
    if (fmt == 2) {
        if (avail < 4) return 0;
        u32 h = rd32(p);
        *regen = (h >> 4) & 0x3FFF;
        *csize = (h >> 18) & 0x3FFF;
        *nstreams = 4;
        return 4;
    }
    if (avail < 5) return 0;
    u32 h = rd32(p);
    *regen = (h >> 4) & 0x3FFFF;
    *csize = ((h >> 22) & 0x3FF) | ((u32)p[4] << 10);
    *nstreams = 4;
    return 5;
}

// ---------------------------------------------------------------------------------- walk: frames and blocks of each stream
// The container gives one compressed size per stream, zstd gives no index inside it: frames and
// blocks have to be hopped over serially.  The serial part is kept to the bare chain (one header
// load per hop, k_zd_hop, one thread per stream); everything that can be read once a block's
// position is known — literals header, sequence count, FSE table descriptions — is parsed by one
// thread per BLOCK (k_zd_parse), and the per-frame bookkeeping (treeless / repeat table provenance,
// output offsets known from headers alone) by one thread per frame (k_zd_link).
// k_zd_hop pass 0 (fill == 0): counts frames / blocks / output bytes per stream.
// pass 1: fills the frame table and the position part of the block table at the bases computed by the host.
// One WARP per stream.  Streams written by this library start with a skippable index frame
// (FQZ_ZPOLICY_INDEX: the compressed size of every frame behind it): the lanes then hop 32 frames at
// a time, and the index is only trusted as far as it is proven — every frame must start with the
// zstd magic and its block chain must end exactly where the index says; any mismatch sends the
// stream down the serial path (lane 0), which is also what streams of any other writer take.

// Hops over ONE zstd frame at p[pos..n).  Returns 0 ok / 1 corrupt / 2 dictionary.  fill: writes
// frames[fidx] and blocks[first_block ...].
__device__ static u32 zd_hop_frame(const u8 *p, u64 n, u64 pos, bool fill, u32 fidx, u32 first_block, u64 dst_off, u32 si, ZDFrame *frames,
                                   ZDBlock *blocks, u32 *nb_out, u64 *osz_out, u64 *end_out) {
    if (n - pos < 6) return 1;
    if (rd32(p + pos) != ZSTD_MAGIC) return 1;
    u32 fhd = p[pos + 4];
    u32 fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_ck = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 8) return 1;  // reserved bit
    if (did) return 2;      // dictionaries are not used by the reference
    u64 h = pos + 5;
    u64 window = 0;
    if (!single) {
        u32 wd = p[h++];
        u32 e = wd >> 3, m = wd & 7;
        window = ((u64)1 << (10 + e));
        window += (window >> 3) * m;
    }
    u32 fcs_size = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (n - h < fcs_size) return 1;
    u64 fcs = ~0ull;
    if (fcs_size == 1) fcs = p[h];
    else if (fcs_size == 2) fcs = (u64)(p[h] | (p[h + 1] << 8)) + 256;
    else if (fcs_size == 4) fcs = rd32(p + h);
    else if (fcs_size == 8) fcs = (u64)rd32(p + h) | ((u64)rd32(p + h + 4) << 32);
    h += fcs_size;
    if (single) window = fcs;
    u32 nb = 0;
    u64 bound = 0;
    for (;;) {
        if (n - h < 3) return 1;
        u32 bh = rd24(p + h);
        u32 last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
        h += 3;
        if (type == 3) return 1;
        u32 csz = (type == 1) ? 1 : bsz;
        if (n - h < csz || bsz > ZSTD_BLOCK_MAX) return 1;
        if (type == 2 && bsz < 2) return 1;
        if (fill) {
            ZDBlock B;
            B.src = (u64)(uintptr_t)(p + h);
            B.lit_off = 0;
            B.seq_off = 0;
            B.csize = csz;
            B.rsize = (type == 2) ? 0u : bsz;
            B.frame = fidx;
            B.lit_regen = 0;
            B.lit_csize = 0;
            B.nseq = 0;
            B.seq_pos = 0;
            B.bits_pos = 0;
            B.huf_block = 0xFFFFFFFFu;
            for (int t = 0; t < 3; t++) {
                B.tab_pos[t] = 0;
                B.fse_block[t] = 0xFFFFFFFFu;
            }
            B.out_off = 0xFFFFFFFFu;
            B.type = (u8)type;
            B.lit_type = 0;
            B.lit_streams = 0;
            B.lit_hdr = 0;
            B.modes = 0;
            B.err = 0;
            B.pad[0] = B.pad[1] = 0;
            blocks[first_block + nb] = B;
        }
        bound += (type == 2) ? ZSTD_BLOCK_MAX : bsz;
        h += csz;
        nb++;
        if (last) break;
    }
    u32 ck = 0;
    if (has_ck) {
        if (n - h < 4) return 1;
        ck = rd32(p + h);
        h += 4;
    }
    u64 osz = (fcs != ~0ull) ? fcs : bound;
    if (fcs != ~0ull && fcs > bound) return 1;
    if (fill) {
        ZDFrame F;
        F.dst_off = dst_off;
        F.content_size = fcs;
        F.first_block = first_block;
        F.nblocks = nb;
        F.has_ck = has_ck;
        F.ck = ck;
        F.stream = si;
        F.out_cap = osz;
        F.out_size = 0;
        F.err = 0;
        F.window = window;
        frames[fidx] = F;
    }
    *nb_out = nb;
    *osz_out = osz;
    *end_out = h;
    return 0;
}

__device__ __forceinline__ u64 warp_incl_scan_u64(u64 v) {
    u32 l = lane_id();
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, v, (unsigned)d);
        if (l >= (u32)d) v += t;
    }
    return v;
}

#define ZD_HOP_WARPS 4
__global__ void __launch_bounds__(ZD_HOP_WARPS * 32) k_zd_hop(const ZDStream *streams, u32 nstreams, ZDStreamInfo *info, ZDFrame *frames, ZDBlock *blocks,
                                                            int fill) {
    u32 si = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (si >= nstreams) return;
    ZDStream st = streams[si];
    const u8 *p = (const u8 *)(uintptr_t)st.src;
    const u64 n = st.csize;
    u32 fbase = 0, bbase = 0;
    u64 obase = 0;
    if (fill) {
        fbase = info[si].frame_base;
        bbase = info[si].block_base;
        obase = info[si].out_base;
    }
    u32 nframes = 0, nblocks = 0, status = 0;
    u64 out_bytes = 0;
    // ---- indexed path
    bool indexed = false;
    if (n >= FQZ_ZINDEX_HDR + 8u && rd32(p) == FQZ_ZINDEX_MAGIC && rd32(p + 8) == FQZ_ZINDEX_SIG) {
        u32 nf = rd32(p + 12);
        u64 isz = (u64)FQZ_ZINDEX_HDR + 8ull * nf;
        if (nf >= 1 && rd32(p + 4) == isz - 8ull && isz < n) {
            u64 sum = 0;
            for (u32 k = lane; k < nf; k += 32) sum += rd32(p + FQZ_ZINDEX_HDR + 8ull * k);
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            indexed = (sum == n - isz);
        }
        if (indexed) {
            u64 pos = isz;  // start of the first frame of the current chunk of 32
            for (u32 f0 = 0; f0 < nf; f0 += 32) {
                u32 k = f0 + lane;
                bool live = k < nf;
                u64 sz = live ? rd32(p + FQZ_ZINDEX_HDR + 8ull * k) : 0;
                u64 incl = warp_incl_scan_u64(sz);
                u64 start = pos + incl - sz;
                u32 nb = 0, rc = 0;
                u64 osz = 0, end = 0;
                if (live) {
                    // the frame is hopped inside its own bounds: a chain that leaves them is a mismatch
                    rc = zd_hop_frame(p, start + sz, start, false, 0, 0, 0, si, frames, blocks, &nb, &osz, &end);
                    if (rc == 0 && end != start + sz) rc = 1;
                    if (rc) nb = 0, osz = 0;
                }
                if (__any_sync(0xffffffffu, rc != 0)) {
                    indexed = false;
                    break;
                }
                u32 nbi = group_incl_scan(nb, 0xffffffffu, 32);
                u64 oi = warp_incl_scan_u64(osz);
                if (fill && live)
                    zd_hop_frame(p, start + sz, start, true, fbase + nframes + lane, bbase + nblocks + nbi - nb, obase + out_bytes + oi - osz, si, frames,
                                 blocks, &nb, &osz, &end);
                nframes += min(32u, nf - f0);
                nblocks += __shfl_sync(0xffffffffu, nbi, 31);
                out_bytes += __shfl_sync(0xffffffffu, oi, 31);
                pos += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    // ---- serial path (lane 0): any frame sequence
    if (!indexed) {
        nframes = 0;
        nblocks = 0;
        out_bytes = 0;
        if (lane == 0) {
            u64 pos = 0;
            while (pos < n) {
                if (n - pos < 4) { status = 1; break; }
                u32 magic = rd32(p + pos);
                if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
                    if (n - pos < 8) { status = 1; break; }
                    u64 sz = rd32(p + pos + 4);
                    if (n - pos - 8 < sz) { status = 1; break; }
                    pos += 8 + sz;
                    continue;
                }
                u32 nb = 0;
                u64 osz = 0, end = 0;
                u32 rc = zd_hop_frame(p, n, pos, fill != 0, fbase + nframes, bbase + nblocks, obase + out_bytes, si, frames, blocks, &nb, &osz, &end);
                if (rc) { status = rc; break; }
                out_bytes += osz;
                nframes++;
                nblocks += nb;
                pos = end;
            }
        }
    }
    if (lane == 0) {
        if (!fill) {
            info[si].nframes = nframes;
            info[si].nblocks = nblocks;
            info[si].out_bytes = out_bytes;
            info[si].status = status;
            info[si].lit_bytes = 0;
            info[si].nseq = 0;
        } else if (status)
            info[si].status = status;
    }
}

// One thread per block: literals header, sequence count, positions of the table descriptions.
// cnt[0][b] = regenerated literals, cnt[1][b] = sequences of block b, cnt[2][b] = 1 when it has
// sequences (all three scanned into arena offsets / table slots later).
__global__ void __launch_bounds__(128) k_zd_parse(ZDBlock *blocks, u32 nblocks, u32 *cnt, u32 cnt_stride) {
    u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nblocks) return;
    u32 lit_regen = 0, nseq = 0;
    const u32 type = blocks[bi].type;
    if (type == 2) {
        const u8 *c = (const u8 *)(uintptr_t)blocks[bi].src;
        const u32 csz = blocks[bi].csize;
        bool ok = true;
        u32 lt = 0, lr = 0, lc = 0, ls = 0, so = 0, sh = 1, modes = 0, bits_pos = 0;
        u32 tab_pos[3] = {0, 0, 0};
        u32 lh = parse_lit_header(c, csz, &lt, &lr, &lc, &ls);
        if (!lh || lh + lc > csz || lr > ZSTD_BLOCK_MAX) ok = false;
        if (ok) {
            so = lh + lc;  // sequences section
            if (so >= csz) ok = false;
        }
        if (ok) {
            u32 b0 = c[so];
            if (b0 == 0) nseq = 0;
            else if (b0 < 128) nseq = b0;
            else if (b0 < 255) {
                if (so + 2 > csz) ok = false;
                else nseq = ((b0 - 128) << 8) + c[so + 1];
                sh = 2;
            } else {
                if (so + 3 > csz) ok = false;
                else nseq = c[so + 1] + ((u32)c[so + 2] << 8) + 0x7F00;
                sh = 3;
            }
        }
        if (ok && nseq) {
            if (so + sh >= csz) ok = false;
            else {
                modes = c[so + sh];
                if (modes & 3) ok = false;
                u32 tp = so + sh + 1;
                // locate the three table descriptions (LL, OF, ML order) and the bitstream
                const u32 maxLogs[3] = {ZSTD_LL_MAXLOG, ZSTD_OF_MAXLOG, ZSTD_ML_MAXLOG};
                const u32 maxSyms[3] = {ZSTD_MAX_LL, ZSTD_MAX_OF, ZSTD_MAX_ML};
                for (int t = 0; t < 3 && ok; t++) {
                    u32 m = (modes >> (6 - 2 * t)) & 3;
                    tab_pos[t] = tp;
                    if (m == 1) {
                        if (tp >= csz) ok = false;
                        tp += 1;
                    } else if (m == 2) {
                        short nrm[64];
                        u32 ms = maxSyms[t], tl = 0;
                        u32 used = (tp < csz) ? fse_read_ncount(c + tp, csz - tp, nrm, &ms, &tl, maxLogs[t]) : 0;
                        if (!used) ok = false;
                        tp += used;
                    }
                }
                if (ok && tp > csz) ok = false;
                bits_pos = tp;
            }
        } else if (ok)
            bits_pos = so + sh;
        if (ok) {
            lit_regen = lr;
            ZDBlock &B = blocks[bi];
            B.lit_type = (u8)lt;
            B.lit_streams = (u8)ls;
            B.lit_hdr = (u8)lh;
            B.lit_csize = lc;
            B.lit_regen = lr;
            B.nseq = nseq;
            B.seq_pos = so + sh;
            B.bits_pos = bits_pos;
            B.modes = (u8)modes;
            for (int t = 0; t < 3; t++) B.tab_pos[t] = tab_pos[t];
        } else {
            nseq = 0;
            blocks[bi].err = 1;
        }
    }
    cnt[bi] = lit_regen;
    cnt[cnt_stride + bi] = nseq;
    cnt[2 * cnt_stride + bi] = nseq ? 1u : 0u;  // takes a slot of sequence decode tables
}

// One thread per frame, blocks in order: which block holds the Huffman tree of a treeless block and
// the FSE table of a repeat-mode block, where each block's output starts while that is computable
// from the headers alone, and sanity limits that keep the literal / sequence arenas within the
// output size (a frame that claims more is corrupt).
__global__ void __launch_bounds__(128) k_zd_link(ZDFrame *frames, u32 nframes, ZDBlock *blocks, u32 *cnt, u32 cnt_stride) {
    u32 fi = blockIdx.x * blockDim.x + threadIdx.x;
    if (fi >= nframes) return;
    const u32 b0 = frames[fi].first_block, nb = frames[fi].nblocks;
    const u64 cap = frames[fi].out_cap;
    u32 huf_src = 0xFFFFFFFFu, fse_src[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    u64 known_off = 0;  // ~0: not computable any more
    u64 lit_sum = 0, seq_sum = 0;
    bool bad = false;
    for (u32 k = 0; k < nb; k++) {
        ZDBlock &B = blocks[b0 + k];
        u32 bidx = b0 + k;
        if (B.err) { bad = true; break; }
        B.out_off = (known_off < 0xFFFFFFFFull) ? (u32)known_off : 0xFFFFFFFFu;
        if (B.type == 2) {
            if (B.lit_type == 3) {
                if (huf_src == 0xFFFFFFFFu) { bad = true; break; }
                B.huf_block = huf_src;
            } else {
                B.huf_block = bidx;
                if (B.lit_type == 2) huf_src = bidx;
            }
            if (B.nseq) {
                for (int t = 0; t < 3; t++) {
                    u32 m = (B.modes >> (6 - 2 * t)) & 3;
                    if (m == 3) {
                        if (fse_src[t] == 0xFFFFFFFFu) bad = true;
                        B.fse_block[t] = fse_src[t];
                    } else {
                        B.fse_block[t] = bidx;
                        fse_src[t] = bidx;
                    }
                }
                if (bad) break;
            }
            lit_sum += B.lit_regen;
            seq_sum += B.nseq;
        }
        if (known_off != ~0ull) {
            if (B.type != 2) known_off += B.rsize;
            else if (B.nseq == 0) known_off += B.lit_regen;
            else known_off = ~0ull;
        }
    }
    if (!bad && (lit_sum > cap || 3 * seq_sum > cap)) bad = true;
    // literal decode groups: up to ZD_GROUP consecutive blocks of ONE frame (they usually share a Huffman tree)
    if (!bad)
        for (u32 k = 0; k < nb; k += 8) cnt[3 * cnt_stride + b0 + k] = 1;
    if (bad) {  // the whole frame is rejected: its blocks take no arena space and are skipped by the decode kernels
        for (u32 k = 0; k < nb; k++) {
            blocks[b0 + k].err = 1;
            cnt[b0 + k] = 0;
            cnt[cnt_stride + b0 + k] = 0;
            cnt[2 * cnt_stride + b0 + k] = 0;
        }
    }
}

// arena offsets of every block = the scanned literal / sequence counts
__global__ void __launch_bounds__(128) k_zd_offsets(ZDBlock *blocks, u32 nblocks, const u32 *cnt, u32 cnt_stride, u32 *seqblk, u32 *litgrp) {
    u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nblocks) return;
    blocks[bi].lit_off = cnt[bi];
    blocks[bi].seq_off = cnt[cnt_stride + bi];
    u32 slot = cnt[2 * cnt_stride + bi];
    if (cnt[2 * cnt_stride + bi + 1] != slot) seqblk[slot] = bi;  // the scanned array has nblocks + 1 entries
    u32 g = cnt[3 * cnt_stride + bi];
    if (cnt[3 * cnt_stride + bi + 1] != g) litgrp[g] = bi;
}

// ---------------------------------------------------------------------------------- literals
#ifndef ZD_WARPS
#define ZD_WARPS 4
#endif
struct LitScratch {
    u16 dt[1 << HUF_MAXBITS];  // nbBits << 8 | symbol
    u8 weights[256];
    u32 rank[16];
    FseDEnt wdt[64];
    u16 wnext[16];
    short wnorm[16];
};
// Builds the Huffman decode table of block `hb` (its literals section carries the tree).  Returns
// tableLog (0 on error) and *tree_size = bytes of the tree description.
__device__ static u32 warp_huf_read_table(const ZDBlock &hb, LitScratch &S, u32 *tree_size) {
    u32 lane = lane_id();
    const u8 *gp = (const u8 *)(uintptr_t)hb.src + hb.lit_hdr;
    u32 avail = hb.lit_csize;
    u32 nsym = 0, tsz = 0, bad = 0;
    for (u32 i = lane; i < 256; i += 32) S.weights[i] = 0;
    // the tree description (at most 1 + 128 bytes) is parsed by one lane: stage it in shared memory (in
    // the decode table's space, which is only filled afterwards) so that the serial parse never waits
    // for global memory
    u8 *stage = (u8 *)S.dt + 16;  // 16 readable bytes in front: the bit reader loads aligned words around its window
    for (u32 i = lane; i < min(avail, 144u); i += 32) stage[i] = gp[i];
    const u8 *p = stage;
    __syncwarp();
    if (lane == 0) {
        if (avail < 1) bad = 1;
        else {
            u32 hbyte = p[0];
            if (hbyte >= 128) {
                nsym = hbyte - 127;
                tsz = 1 + (nsym + 1) / 2;
                if (tsz > avail) bad = 1;
                else
                    for (u32 i = 0; i < nsym; i++) {
                        u32 b = p[1 + i / 2];
                        S.weights[i] = (u8)((i & 1) ? (b & 15) : (b >> 4));
                    }
            } else {
                tsz = 1 + hbyte;
                if (hbyte == 0 || tsz > avail) bad = 1;
                else {
                    u32 ms = 12, tl = 0;
                    u32 used = fse_read_ncount(p + 1, hbyte, S.wnorm, &ms, &tl, 6);
                    if (!used || used >= hbyte) bad = 1;
                    else {
                        fse_build_dtable(S.wnorm, ms, tl, S.wdt, S.wnext);
                        BitR br;
                        br.init(p + 1 + used, hbyte - used);
                        u32 s1 = br.read(tl), s2 = br.read(tl);
                        br.reload();
                        // two interleaved states; the stream ends when the bits run out
                        for (;;) {
                            if (nsym >= 254) { bad = 1; break; }
                            S.weights[nsym++] = S.wdt[s1].sym;
                            {
                                u32 nb = S.wdt[s1].nb;
                                if (br.consumed + nb > 64 && br.ptr == br.start) { S.weights[nsym++] = S.wdt[s2].sym; break; }
                                s1 = S.wdt[s1].base + br.read(nb);
                                br.reload();
                            }
                            if (nsym >= 254) { bad = 1; break; }
                            S.weights[nsym++] = S.wdt[s2].sym;
                            {
                                u32 nb = S.wdt[s2].nb;
                                if (br.consumed + nb > 64 && br.ptr == br.start) { S.weights[nsym++] = S.wdt[s1].sym; break; }
                                s2 = S.wdt[s2].base + br.read(nb);
                                br.reload();
                            }
                        }
                        if (br.bad) bad = 1;
                    }
                }
            }
        }
        if (!bad) {  // implied last weight
            u32 total = 0;
            for (u32 i = 0; i < nsym; i++) {
                u32 w = S.weights[i];
                if (w > HUF_MAXBITS + 1) { bad = 1; break; }
                if (w) total += 1u << (w - 1);
            }
            if (total == 0) bad = 1;
            if (!bad) {
                u32 tl = hibit32(total) + 1;
                u32 rest = (1u << tl) - total;
                if (tl > HUF_MAXBITS + 1 || (rest & (rest - 1)) != 0) bad = 1;
                else {
                    S.weights[nsym] = (u8)(hibit32(rest) + 1);
                    nsym++;
                    S.rank[15] = tl;
                }
            }
        }
        S.rank[14] = bad;
        S.rank[13] = nsym;
        S.rank[12] = tsz;
    }
    __syncwarp();
    bad = S.rank[14];
    u32 tl = S.rank[15];
    nsym = S.rank[13];
    *tree_size = S.rank[12];
    __syncwarp();
    if (bad || tl > HUF_MAXBITS) return 0;
    // rank starts: weight 1 first
    if (lane < 12) S.rank[lane] = 0;
    __syncwarp();
    for (u32 i = lane; i < nsym; i += 32)
        if (S.weights[i]) atomicAdd(&S.rank[S.weights[i]], 1u);
    __syncwarp();
    if (lane == 0) {
        u32 start = 0;
        for (u32 w = 1; w <= tl; w++) {
            u32 c = S.rank[w];
            S.rank[w] = start;
            start += c << (w - 1);
        }
        S.rank[0] = start;  // must equal 1 << tl
    }
    __syncwarp();
    if (S.rank[0] != (1u << tl)) return 0;
    // Fill: a symbol of weight w owns 2^(w-1) consecutive cells.  Short runs are written by the lane that
    // owns the symbol; long runs (the few frequent symbols own most of the table — up to half of it each)
    // are handed to the whole warp, 32 cells per step, so that no lane writes a thousand cells alone.
    for (u32 s0 = 0; s0 < nsym; s0 += 32) {
        u32 s = s0 + lane;
        u32 w = (s < nsym) ? S.weights[s] : 0u;
        u32 len = 0, at = 0;
        u32 e = 0;
        if (w) {
            u32 idx = 0;
            for (u32 t = 0; t < s; t++) idx += (S.weights[t] == w) ? 1u : 0u;
            len = 1u << (w - 1);
            at = S.rank[w] + idx * len;
            e = ((tl + 1 - w) << 8) | s;
            if (len < 32u)
                for (u32 k = 0; k < len; k++) S.dt[at + k] = (u16)e;
        }
        u32 big = __ballot_sync(FULL, len >= 32u);
        while (big) {
            int l = __ffs((int)big) - 1;
            big &= big - 1u;
            u32 blen = __shfl_sync(FULL, len, l), bat = __shfl_sync(FULL, at, l), be = __shfl_sync(FULL, e, l);
            for (u32 k = lane; k < blen; k += 32) S.dt[bat + k] = (u16)be;
        }
    }
    __syncwarp();
    return tl;
}

// one Huffman stream, decoded by one lane; output assembled into aligned 32-bit stores.
// The stream is read backwards in ALIGNED 32-bit words that are loaded three refills ahead of
// their use, so the symbol chain (shift -> table lookup -> shift) never waits for global memory;
// the bit window lives left-aligned in a 64-bit register.  Instead of guarding every read against
// the start of the stream, the bits consumed are counted: a valid stream is consumed exactly
// (RFC 8878 §4.2.2), anything else is corrupt.  Word loads are clamped to the aligned word that
// holds src[0], so a corrupt stream never reads outside the block either.
struct HufBits {
    u64 bits;       // unread bits, left-aligned
    u32 nb;         // valid bits in `bits`
    const u32 *lo;  // aligned word that holds the first byte of the stream
    u32 wi;         // index (from lo) of the next word to prefetch; sticks at 0
    u32 w0, w1, w2, w3; // prefetched words, nearest first; w3 is the load in flight
    // src[csize - 1] != 0 (it holds the end mark).  Returns the payload bits of the stream.
    __device__ __forceinline__ u32 init(const u8 *src, u32 csize) {
        u32 hb = hibit32(src[csize - 1]);
        const u8 *endp = src + csize;
        lo = (const u32 *)((uintptr_t)src & ~(uintptr_t)3);
        const u32 *wa = (const u32 *)((uintptr_t)(endp - 1) & ~(uintptr_t)3);
        u32 k = (u32)((uintptr_t)(endp - 1) & 3u) + 1u;  // stream bytes in the top word
        u32 v = *wa;
        wi = (u32)(wa - lo);
        bits = (u64)v << (64u - 8u * k);
        bits <<= (8u - hb);  // padding and end mark
        nb = 8u * k - (8u - hb);
        wi = wi ? wi - 1u : 0u;
        w0 = lo[wi];
        wi = wi ? wi - 1u : 0u;
        w1 = lo[wi];
        wi = wi ? wi - 1u : 0u;
        w2 = lo[wi];
        wi = wi ? wi - 1u : 0u;
        w3 = lo[wi];
        wi = wi ? wi - 1u : 0u;
        return (csize - 1u) * 8u + hb;
    }
    // Afterwards nb >= 32 (bits past the stream start are whatever memory holds).  Branch-free on
    // purpose: the 32 lanes of a warp run 32 different streams and would otherwise take this path at
    // different times, every one of them paying for all the others.  The word loaded here is first
    // looked at by the NEXT refill that shifts the queue, several symbols later, so the chain never
    // waits for it; a refill that does not shift loads nothing (predicated load).
    __device__ __forceinline__ void refill() {
        const bool need = nb <= 32u;
        const u64 add = (u64)w0 << ((32u - nb) & 63u);
        bits |= need ? add : 0ull;
        nb += need ? 32u : 0u;
        w0 = need ? w1 : w0;
        w1 = need ? w2 : w1;
        w2 = need ? w3 : w2;
        if (need) w3 = lo[wi];
        wi = (need && wi) ? wi - 1u : wi;
    }
    __device__ __forceinline__ u32 peek(u32 n) const { return (u32)(bits >> 32) >> (32u - n); }  // 1 <= n <= 32
    __device__ __forceinline__ void drop(u32 n) {
        bits <<= n;
        nb -= n;
    }
    __device__ __forceinline__ u32 take(u32 n) {  // 0 <= n <= 32, n <= nb
        if (n == 0) return 0;
        u32 v = peek(n);
        drop(n);
        return v;
    }
};

__device__ static bool huf_decode_stream(const u8 *src, u32 csize, u8 *dst, u32 n, const u16 *dt, u32 tl) {
    if (csize == 0 || src[csize - 1] == 0) return false;  // empty, or the end mark is missing
    HufBits hb;
    const u32 total = hb.init(src, csize);
    u32 used = 0;
    u32 i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 3u)) {
        hb.refill();
        u32 e = dt[hb.peek(tl)];
        hb.drop(e >> 8);
        used += e >> 8;
        dst[i++] = (u8)e;
    }
    while (i + 4 <= n) {  // a refill leaves >= 33 bits: two symbols of <= 11 bits each, twice
        u32 w = 0;
        hb.refill();
        {
            u32 e = dt[hb.peek(tl)];
            hb.drop(e >> 8);
            used += e >> 8;
            w = e & 0xFFu;
            e = dt[hb.peek(tl)];
            hb.drop(e >> 8);
            used += e >> 8;
            w |= (e & 0xFFu) << 8;
        }
        hb.refill();
        {
            u32 e = dt[hb.peek(tl)];
            hb.drop(e >> 8);
            used += e >> 8;
            w |= (e & 0xFFu) << 16;
            e = dt[hb.peek(tl)];
            hb.drop(e >> 8);
            used += e >> 8;
            w |= (e & 0xFFu) << 24;
        }
        *(u32 *)(dst + i) = w;
        i += 4;
    }
    while (i < n) {
        hb.refill();
        u32 e = dt[hb.peek(tl)];
        hb.drop(e >> 8);
        used += e >> 8;
        dst[i++] = (u8)e;
    }
    return used == total;
}

// One warp takes eight consecutive blocks; lane = (block, Huffman stream).  Blocks that share a
// Huffman table (the treeless blocks of the frames written by k_zenc_huf: 8 blocks x 4 streams = all
// 32 lanes busy, one table build) are decoded together; blocks with their own trees (frames written
// by the reference's encoder) take one pass per table.  Sequence-free blocks whose position in the
// frame is known from the headers are decoded straight into the output.
#define ZD_GROUP 8
__global__ void __launch_bounds__(ZD_WARPS * 32) k_zd_literals(ZDBlock *blocks, u32 nblocks, const u32 *litgrp, u32 ngroups, const ZDFrame *frames,
                                                               u8 *litbuf, u8 *out) {
    __shared__ LitScratch scratch[ZD_WARPS];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 g = blockIdx.x * ZD_WARPS + warp;
    if (g >= ngroups) return;
    const u32 b0 = litgrp[g];  // first block of the group; the group never leaves its frame
    LitScratch &S = scratch[warp];
    u32 bl = lane >> 2, q = lane & 3u;
    u32 bi = b0 + bl;
    bool valid = bi < nblocks && blocks[bi].frame == blocks[b0].frame;
    ZDBlock B;
    if (valid) B = blocks[bi];
    else {
        B.type = 0;
        B.lit_type = 0;
        B.huf_block = 0xFFFFFFFFu;
    }
    bool comp = valid && B.type == 2 && !B.err;
    u8 *dst = nullptr;
    u32 n = 0;
    u32 err = 0;
    bool direct = false;
    if (comp) {
        const ZDFrame &F = frames[B.frame];
        direct = (B.nseq == 0 && B.out_off != 0xFFFFFFFFu);
        n = B.lit_regen;
        if (direct && (u64)B.out_off + n > F.out_cap) {
            err = 1;
            comp = false;
        } else
            dst = direct ? out + F.dst_off + B.out_off : litbuf + B.lit_off;
    }
    // raw / RLE literals: the block's four lanes copy / fill
    if (comp && B.lit_type < 2) {
        const u8 *c = (const u8 *)(uintptr_t)B.src + B.lit_hdr;
        if (B.lit_type == 0) for (u32 i = q; i < n; i += 4) dst[i] = c[i];
        else {
            u8 v = c[0];
            for (u32 i = q; i < n; i += 4) dst[i] = v;
        }
    }
    bool pending = comp && B.lit_type >= 2;
    for (;;) {
        u32 pm = __ballot_sync(FULL, pending);
        if (!pm) break;
        int fl = __ffs((int)pm) - 1;
        u32 tsrc = __shfl_sync(FULL, B.huf_block, fl);
        ZDBlock HB = blocks[tsrc];
        u32 tree = 0;
        u32 tl = warp_huf_read_table(HB, S, &tree);
        bool mine = pending && B.huf_block == tsrc;
        if (mine) {
            if (!tl) err = 1;
            else {
                const u8 *c = (const u8 *)(uintptr_t)B.src;
                u32 tsz = (B.lit_type == 3) ? 0u : tree;  // treeless: the streams start right after the header
                const u8 *sp = c + B.lit_hdr + tsz;
                if (tsz > B.lit_csize) err = 1;
                else {
                    u32 avail = B.lit_csize - tsz;
                    if (B.lit_streams == 1) {
                        if (q == 0 && !huf_decode_stream(sp, avail, dst, n, S.dt, tl)) err = 1;
                    } else if (avail < 10) err = 1;
                    else {
                        u32 s1 = sp[0] | (sp[1] << 8), s2 = sp[2] | (sp[3] << 8), s3 = sp[4] | (sp[5] << 8);
                        u32 seg = (n + 3) / 4;
                        if (6 + s1 + s2 + s3 >= avail || 3 * seg > n) err = 1;
                        else {
                            u32 off = 6 + (q > 0 ? s1 : 0) + (q > 1 ? s2 : 0) + (q > 2 ? s3 : 0);
                            u32 cs = q == 0 ? s1 : q == 1 ? s2 : q == 2 ? s3 : avail - 6 - s1 - s2 - s3;
                            u32 cnt = q < 3 ? seg : n - 3 * seg;
                            if (!huf_decode_stream(sp + off, cs, dst + q * seg, cnt, S.dt, tl)) err = 1;
                        }
                    }
                }
            }
            pending = false;
        }
        __syncwarp();
    }
    // per block verdict
    u32 e4 = err;
    e4 |= __shfl_xor_sync(FULL, e4, 1);
    e4 |= __shfl_xor_sync(FULL, e4, 2);
    if (valid && q == 0) {
        if (e4) blocks[bi].err = 1;
        else if (comp && direct) blocks[bi].rsize = n;
    }
}

// ---------------------------------------------------------------------------------- sequences
// Decoding the sequences of one block is a serial chain (three FSE states fed by one backward
// bitstream), so the parallelism is ACROSS blocks: k_zd_seq_tables builds the decode tables of every
// block that has sequences into global memory (one warp per block, lanes 0/1/2 one table each), and
// k_zd_seq_decode runs ONE THREAD per block with its tables in L2 — all 32 lanes of a warp busy on
// 32 different blocks, where one warp per block kept a single lane busy.
struct SeqScratch {
    FseDEnt dt[3][512];
    u16 next[3][64];
    short norm[3][64];
    u32 tl[3];
};
// Builds the decode table of stream type t (0 LL, 1 OF, 2 ML) of block SB (which carries its own
// description: modes 0 / 1 / 2).  Returns tableLog, 0xFF on error; RLE tables have log 0 and one entry.
__device__ static u32 build_seq_table(const ZDBlock &SB, int t, SeqScratch &S) {
    u32 m = (SB.modes >> (6 - 2 * t)) & 3;
    const u8 *c = (const u8 *)(uintptr_t)SB.src;
    const u32 maxLogs[3] = {ZSTD_LL_MAXLOG, ZSTD_OF_MAXLOG, ZSTD_ML_MAXLOG};
    const u32 maxSyms[3] = {ZSTD_MAX_LL, ZSTD_MAX_OF, ZSTD_MAX_ML};
    if (m == 0) {
        const short *dn = (t == 0) ? kLLDefNorm : (t == 1 ? kOFDefNorm : kMLDefNorm);
        u32 ms = (t == 0) ? 35 : (t == 1 ? 28 : 52);
        u32 tl = (t == 1) ? ZSTD_OF_DEFLOG : ZSTD_LL_DEFLOG;
        for (u32 s = 0; s <= ms; s++) S.norm[t][s] = dn[s];
        fse_build_dtable(S.norm[t], ms, tl, S.dt[t], S.next[t]);
        return tl;
    }
    if (m == 1) {
        u32 sym = c[SB.tab_pos[t]];
        if (sym > maxSyms[t]) return 0xFF;
        S.dt[t][0].sym = (u8)sym;
        S.dt[t][0].nb = 0;
        S.dt[t][0].base = 0;
        return 0;
    }
    if (m == 2) {
        u32 ms = maxSyms[t], tl = 0;
        u32 used = fse_read_ncount(c + SB.tab_pos[t], SB.csize - SB.tab_pos[t], S.norm[t], &ms, &tl, maxLogs[t]);
        if (!used) return 0xFF;
        fse_build_dtable(S.norm[t], ms, tl, S.dt[t], S.next[t]);
        return tl;
    }
    return 0xFF;  // repeat mode: the table lives with another block
}

// seqblk[slot] = block number of the slot-th block with sequences; tabs: ZD_TAB_WORDS u32 per slot:
// three tables of 512 entries, then the three table logs (0xFF = bad / not this block's own table).
#define ZD_TAB_WORDS (3u * 512u + 4u)
__global__ void __launch_bounds__(ZD_WARPS * 32) k_zd_seq_tables(const ZDBlock *blocks, const u32 *seqblk, u32 nsb, u32 *tabs) {
    __shared__ SeqScratch scratch[ZD_WARPS];
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 slot = blockIdx.x * ZD_WARPS + warp;
    if (slot >= nsb) return;
    const ZDBlock &B = blocks[seqblk[slot]];
    SeqScratch &S = scratch[warp];
    if (lane < 3) {
        bool own = ((B.modes >> (6 - 2 * lane)) & 3) != 3;
        S.tl[lane] = own ? build_seq_table(B, (int)lane, S) : 0xFFu;
    }
    __syncwarp();
    u32 *o = tabs + (size_t)slot * ZD_TAB_WORDS;
    for (int t = 0; t < 3; t++) {
        u32 tl = S.tl[t];
        if (tl == 0xFF) continue;
        u32 n = 1u << tl;
        const u32 *src = (const u32 *)S.dt[t];
        for (u32 i = lane; i < n; i += 32) o[512u * t + i] = src[i];
    }
    if (lane < 3) o[1536u + lane] = S.tl[lane];
}

// One thread per block with sequences.  slot_of[b] = table slot of block b (exclusive scan of the
// has-sequences flags).
__global__ void __launch_bounds__(128) k_zd_seq_decode(ZDBlock *blocks, const ZDFrame *frames, const u32 *seqblk, u32 nsb, const u32 *slot_of, const u32 *tabs,
                                                       u32 *seqbuf) {
    u32 slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nsb) return;
    const u32 bi = seqblk[slot];
    const ZDBlock &B = blocks[bi];
    u32 err = 0;
    u32 tl[3];
    const u32 *dt[3];
    for (int t = 0; t < 3; t++) {
        u32 sb = B.fse_block[t];  // == bi unless the table is repeated from an earlier block of the frame
        const u32 *tb = tabs + (size_t)slot_of[sb] * ZD_TAB_WORDS;
        dt[t] = tb + 512u * t;
        tl[t] = tb[1536u + t];
        if (tl[t] == 0xFF) err = 1;
    }
    const u32 nseq = B.nseq;
    u32 *sq = seqbuf + 3ull * B.seq_off;
    u64 total = 0;
    // the only block of its frame starts from the initial repeat history: resolve the offsets here, where
    // the chain is serial anyway, and spare k_zd_execute its per-sequence loop (RFC 8878 §3.1.1.5)
    const bool resolve = frames[B.frame].nblocks == 1;
    u32 rep0 = 1, rep1 = 4, rep2 = 8;
    if (!err) {
        const u8 *c = (const u8 *)(uintptr_t)B.src;
        const u32 bsz = B.csize - B.bits_pos;
        if (bsz == 0 || c[B.csize - 1] == 0) err = 1;  // no bitstream, or its end mark is missing
        else {
            // same reader as the Huffman streams: aligned words prefetched ahead of the state chain,
            // consumed bits counted and checked against the stream size at the end.  Every group of
            // reads between two refills takes at most 32 bits (offset extra bits <= 31; match + literal
            // length extra bits <= 16 + 16; three state updates <= 9 + 9 + 8).
            HufBits br;
            const u32 avail = br.init(c + B.bits_pos, bsz);
            u32 used = tl[0] + tl[1] + tl[2];
            br.refill();
            u32 sLL = br.take(tl[0]), sOF = br.take(tl[1]), sML = br.take(tl[2]);
            for (u32 i = 0; i < nseq; i++) {
                u32 eLL = dt[0][sLL], eOF = dt[1][sOF], eML = dt[2][sML];  // FseDEnt: base | sym << 16 | nb << 24
                u32 ofc = (eOF >> 16) & 0xFFu, mlc = (eML >> 16) & 0xFFu, llc = (eLL >> 16) & 0xFFu;
                if (ofc > 31 || mlc > 52 || llc > 35) { err = 1; break; }
                br.refill();
                u32 ofv = (1u << ofc) + br.take(ofc);
                br.refill();
                u32 mlb = kMLBits[mlc], llb = kLLBits[llc];
                u32 ml = kMLBase[mlc] + br.take(mlb);
                u32 ll = kLLBase[llc] + br.take(llb);
                used += ofc + mlb + llb;
                if (resolve) {
                    u32 f;
                    if (ofv > 3) {
                        f = ofv - 3;
                        rep2 = rep1; rep1 = rep0; rep0 = f;
                    } else {
                        u32 idx = ofv - 1 + (ll == 0 ? 1u : 0u);
                        if (idx == 0) f = rep0;
                        else {
                            f = (idx == 1) ? rep1 : (idx == 2) ? rep2 : rep0 - 1;
                            if (idx != 1) rep2 = rep1;
                            rep1 = rep0;
                            rep0 = f;
                        }
                    }
                    ofv = f;
                }
                sq[3 * i] = ll;
                sq[3 * i + 1] = ml;
                sq[3 * i + 2] = ofv;
                total += (u64)ll + ml;
                if (i + 1 < nseq) {
                    br.refill();
                    u32 nLL = eLL >> 24, nML = eML >> 24, nOF = eOF >> 24;
                    sLL = (eLL & 0xFFFFu) + br.take(nLL);
                    sML = (eML & 0xFFFFu) + br.take(nML);
                    sOF = (eOF & 0xFFFFu) + br.take(nOF);
                    used += nLL + nML + nOF;
                }
                if (used > avail) { err = 1; break; }
            }
            if (used != avail) err = 1;
        }
    }
    if (err || total > ZSTD_BLOCK_MAX) blocks[bi].err = 1;
    else if (resolve) blocks[bi].pad[0] = 1;  // sq[3i+2] holds real offsets
}

// ---------------------------------------------------------------------------------- execution
// One warp per frame: blocks in order, sequences in order; each copy is spread over the 32 lanes.
__device__ __forceinline__ void warp_copy(u8 *dst, const u8 *src, u32 n, u32 lane) {
    for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
}
// Raw and RLE blocks whose place in the frame is known from the headers alone (k_zd_link) are
// regenerated here, one warp per block with 16-byte stores, instead of by the frame's single warp in
// k_zd_execute (incompressible streams — 2-bit packed bases of random reads — are all raw blocks).
__global__ void __launch_bounds__(256) k_zd_rawcopy(const ZDBlock *blocks, u32 nblocks, const ZDFrame *frames, u8 *out) {
    u32 bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (bi >= nblocks) return;
    const ZDBlock &B = blocks[bi];
    if (B.type > 1 || B.out_off == 0xFFFFFFFFu || B.err) return;
    const ZDFrame &F = frames[B.frame];
    const u32 n = B.rsize;
    if ((u64)B.out_off + n > F.out_cap) return;  // k_zd_execute reports it
    u8 *dst = out + F.dst_off + B.out_off;
    const u8 *src = (const u8 *)(uintptr_t)B.src;
    u32 head = (u32)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    const u32 v0 = src[0];
    const u32 fill = v0 * 0x01010101u;
    const bool rle = B.type == 1;
    if (lane < head) dst[lane] = rle ? (u8)v0 : src[lane];
    u32 nv = (n - head) >> 4;
    uint4 *d16 = (uint4 *)(dst + head);
    const u8 *s16 = src + head;
    for (u32 v = lane; v < nv; v += 32) {
        uint4 x;
        if (rle) x = make_uint4(fill, fill, fill, fill);
        else {
            const u8 *sp = s16 + 16u * v;
            x = make_uint4(ld_u32_unaligned(sp), ld_u32_unaligned(sp + 4), ld_u32_unaligned(sp + 8), ld_u32_unaligned(sp + 12));
        }
        d16[v] = x;
    }
    u32 t0 = head + 16u * nv;
    if (lane < n - t0) dst[t0 + lane] = rle ? (u8)v0 : src[t0 + lane];
}

// n bytes from s to d by ONE lane, forward.  d and s may overlap when d - s >= 4 (an LZ match whose
// offset is at least 4): words are read before they are written.  Destination-aligned 32-bit stores,
// source words assembled from aligned loads.
__device__ __forceinline__ void lane_copy_fwd(u8 *d, const u8 *s, u32 n) {
    u32 k = 0;
    u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
    if (head > n) head = n;
    for (; k < head; k++) d[k] = s[k];
    for (; k + 4 <= n; k += 4) *(u32 *)(d + k) = ld_u32_unaligned(s + k);
    for (; k < n; k++) d[k] = s[k];
}

// Frames of one block with sequences and at most ZX_STAGE bytes of content (the 16 KiB item frames
// this library writes) are executed in shared memory — the dependency rounds then cost a shared-memory
// round trip instead of one through L2 — and written out with wide coalesced stores.
#define ZX_ENABLE 1
#ifndef ZX_STAGE
#define ZX_STAGE 16384u
#endif
#define ZX_SMEM (ZX_ENABLE ? ZD_WARPS * (ZX_STAGE + 64u) : 0u)
__global__ void __launch_bounds__(ZD_WARPS * 32) k_zd_execute(ZDFrame *frames, u32 nframes, ZDBlock *blocks, const u8 *litbuf, const u32 *seqbuf, u8 *out) {
    FQZ_DYN_SMEM(u8, smem);
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 fi = blockIdx.x * ZD_WARPS + warp;
    if (fi >= nframes) return;
    ZDFrame F = frames[fi];
    u8 *const gbase = out + F.dst_off;
    u8 *base = gbase;
    const bool staged = ZX_ENABLE && F.nblocks == 1 && F.out_cap <= ZX_STAGE && blocks[F.first_block].type == 2 && blocks[F.first_block].nseq > 0;
    if (staged) base = smem + (size_t)warp * (ZX_STAGE + 64u);
    u64 o = 0;  // bytes regenerated so far in this frame
    u32 rep0 = 1, rep1 = 4, rep2 = 8;
    u32 err = 0;
    for (u32 b = 0; b < F.nblocks && !err; b++) {
        ZDBlock B = blocks[F.first_block + b];
        if (B.err) { err = 1; break; }
        const u8 *c = (const u8 *)(uintptr_t)B.src;
        if (B.type <= 1 && B.out_off != 0xFFFFFFFFu) {  // regenerated by k_zd_rawcopy
            if (o + B.rsize > F.out_cap || (u64)B.out_off != o) { err = 1; break; }
            o += B.rsize;
        } else if (B.type == 0) {
            if (o + B.rsize > F.out_cap) { err = 1; break; }
            warp_copy(base + o, c, B.rsize, lane);
            o += B.rsize;
        } else if (B.type == 1) {
            if (o + B.rsize > F.out_cap) { err = 1; break; }
            u8 v = c[0];
            for (u32 i = lane; i < B.rsize; i += 32) base[o + i] = v;
            o += B.rsize;
        } else if (B.nseq == 0) {
            if (o + B.lit_regen > F.out_cap) { err = 1; break; }
            if (B.out_off == 0xFFFFFFFFu) warp_copy(base + o, litbuf + B.lit_off, B.lit_regen, lane);  // else: decoded in place
            else if ((u64)B.out_off != o) { err = 1; break; }
            o += B.lit_regen;
        } else {
            // Sequences are taken 32 at a time, one per lane.  Offsets are resolved against the repeat
            // history in a short uniform loop, a warp scan of the lengths gives every lane its place,
            // all literal runs are copied at once, and the matches complete in rounds: a lane copies as
            // soon as everything its source range needs is final — always true for the first unfinished
            // lane, and true for many at once when the offsets reach back beyond the group (round trip
            // through memory per round instead of per sequence).  Long matches are left to the whole warp.
            const u8 *lit = litbuf + B.lit_off;
            const u32 *sq = seqbuf + 3ull * B.seq_off;
            u32 lp = 0;
            u64 o0 = o;
            for (u32 sb = 0; sb < B.nseq && !err; sb += 32) {
                const u32 cn = min(32u, B.nseq - sb);
                const bool live = lane < cn;
                u32 ll = 0, ml = 0, ofv = 0;
                if (live) {
                    ll = sq[3 * (sb + lane)];
                    ml = sq[3 * (sb + lane) + 1];
                    ofv = sq[3 * (sb + lane) + 2];
                }
                // repeat offsets (RFC 8878 §3.1.1.5), uniform — unless k_zd_seq_decode resolved them already
                u32 off = ofv;
                if (!B.pad[0])
                for (u32 j = 0; j < cn; j++) {
                    u32 v = __shfl_sync(FULL, ofv, (int)j), l = __shfl_sync(FULL, ll, (int)j);
                    u32 f;
                    if (v > 3) {
                        f = v - 3;
                        rep2 = rep1; rep1 = rep0; rep0 = f;
                    } else {
                        u32 idx = v - 1 + (l == 0 ? 1u : 0u);
                        if (idx == 0) f = rep0;
                        else {
                            f = (idx == 1) ? rep1 : (idx == 2) ? rep2 : rep0 - 1;
                            if (idx != 1) rep2 = rep1;
                            rep1 = rep0;
                            rep0 = f;
                        }
                    }
                    if (lane == j) off = f;
                }
                u32 tot = ll + ml;
                u32 incl = group_incl_scan(tot, FULL, 32), lincl = group_incl_scan(ll, FULL, 32);
                u32 gtot = __shfl_sync(FULL, incl, 31), gll = __shfl_sync(FULL, lincl, 31);
                u64 lstart = o + incl - tot, mstart = lstart + ll;
                u32 lsrc = lp + lincl - ll;
                bool bad = live && (off == 0 || (u64)off > mstart);
                if (__any_sync(FULL, bad) || lp + gll > B.lit_regen || o + gtot > F.out_cap) { err = 1; break; }
                if (live) lane_copy_fwd(base + lstart, lit + lsrc, ll);
                __syncwarp();
                if (staged) {
                    // shared memory: one sequence after the other, every match copied by the whole warp
                    // (a shared-memory round trip per sequence, ~14 instructions)
                    for (u32 j = 0; j < cn; j++) {
                        u32 hm = (u32)__shfl_sync(FULL, mstart, (int)j), hoff = __shfl_sync(FULL, off, (int)j), hml = __shfl_sync(FULL, ml, (int)j);
                        u8 *d = base + hm;
                        const u8 *sp = d - hoff;
                        if (hoff >= hml) {
                            for (u32 k = lane; k < hml; k += 32) d[k] = sp[k];
                        } else if (hoff >= 32u) {
                            for (u32 k0 = 0; k0 < hml; k0 += 32) {
                                if (k0 + lane < hml) d[k0 + lane] = sp[k0 + lane];
                                __syncwarp();
                            }
                        } else {
                            for (u32 k = lane; k < hml; k += 32) d[k] = sp[k % hoff];
                        }
                        __syncwarp();
                    }
                } else {
                bool done = !live || ml == 0;
                    const u64 need = (off >= ml) ? (mstart - off + ml) : mstart;  // output that must be final before this lane copies
                    for (;;) {
                        u32 pend = __ballot_sync(FULL, !done);
                        if (!pend) break;
                        int hw = __ffs((int)pend) - 1;
                        u64 hwpos = __shfl_sync(FULL, mstart, hw);  // everything below is final
                        u32 hml = __shfl_sync(FULL, ml, hw);
                        if (hml > 64u) {  // long match at the front: the whole warp copies it
                            u32 hoff = __shfl_sync(FULL, off, hw);
                            u8 *d = base + hwpos;
                            const u8 *sp = d - hoff;
                            if (hoff >= hml) {
                                for (u32 k = lane; k < hml; k += 32) d[k] = sp[k];
                            } else if (hoff >= 32u) {  // overlapping, but 32 bytes at a time never overtake the source
                                for (u32 k0 = 0; k0 < hml; k0 += 32) {
                                    if (k0 + lane < hml) d[k0 + lane] = sp[k0 + lane];
                                    __syncwarp();
                                }
                            } else {
                                for (u32 k = lane; k < hml; k += 32) d[k] = sp[k % hoff];  // the source is the last `off` bytes, repeated
                            }
                            if ((int)lane == hw) done = true;
                        } else if (!done && ml <= 64u && ((int)lane == hw || need <= hwpos)) {
                            u8 *d = base + mstart;
                            const u8 *sp = d - off;
                            if (off >= 4u) lane_copy_fwd(d, sp, ml);
                            else
                                for (u32 k = 0; k < ml; k++) d[k] = sp[k];
                            done = true;
                        }
                        __syncwarp();
                    }
                }
                o += gtot;
                lp += gll;
            }
            if (err) break;
            u32 rest = B.lit_regen - lp;
            if (o + rest > F.out_cap) { err = 1; break; }
            warp_copy(base + o, lit + lp, rest, lane);
            o += rest;
            if (o - o0 > ZSTD_BLOCK_MAX) { err = 1; break; }
        }
        __syncwarp();
    }
    if (!err && F.content_size != ~0ull && o != F.content_size) err = 1;
    if (staged && !err) {  // shared memory -> output
        __syncwarp();
        u32 n = (u32)o;
        if ((((uintptr_t)gbase) & 15u) == 0) {
            u32 nv = n >> 4;
            for (u32 v = lane; v < nv; v += 32) ((uint4 *)gbase)[v] = ((const uint4 *)base)[v];
            for (u32 k = (nv << 4) + lane; k < n; k += 32) gbase[k] = base[k];
        } else
            for (u32 k = lane; k < n; k += 32) gbase[k] = base[k];
    }
    if (lane == 0) {
        frames[fi].out_size = o;
        frames[fi].err = err;
    }
}

// ---------------------------------------------------------------------------------- content checksum
// see fqz_xxh64.cuh: 4 threads per frame, 8 frames per warp, frames streamed through shared memory by TMA
__global__ void __launch_bounds__(XX_WARPS * 32) k_zd_checksum(ZDFrame *frames, u32 nframes, const u8 *out) {
    FQZ_DYN_SMEM(u8, smem);
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 fi = t >> 2, q = t & 3;
    u32 gm = group_mask(4);
    bool live = fi < nframes;
    ZDFrame F = frames[live ? fi : nframes - 1];
    u64 len = (!live || F.err || !F.has_ck) ? 0 : F.out_size;
    u8 *rows;
    u64 *bars;
    xx_quad_smem(smem, &rows, &bars);
    u64 h = xxh64_quad_staged(out + F.dst_off, len, q, gm, rows, bars);
    if (live && q == 0 && F.has_ck && !F.err && (u32)h != F.ck) frames[fi].err = 2;
}

// stream-level verdict: first failing frame wins; also gathers the decoded size of each stream
__global__ void k_zd_finish(const ZDFrame *frames, const ZDStreamInfo *info, u32 nstreams, ZDStreamResult *res) {
    u32 si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nstreams) return;
    ZDStreamInfo I = info[si];
    u64 total = 0;
    u32 err = I.status;
    bool contiguous = true;
    for (u32 f = 0; f < I.nframes; f++) {
        const ZDFrame &F = frames[I.frame_base + f];
        if (F.err && !err) err = 10 + F.err;
        if (F.out_size != F.out_cap) contiguous = false;
        total += F.out_size;
    }
    res[si].err = err;
    res[si].size = total;
    res[si].contiguous = contiguous ? 1u : 0u;
}

// Streams whose frames carried no content size were given upper-bound output regions: close the
// gaps so that every stream is contiguous again (one warp per stream, frames moved left in order).
__global__ void __launch_bounds__(128) k_zd_compact(ZDFrame *frames, const ZDStreamInfo *info, const ZDStreamResult *res, u32 nstreams, u8 *out) {
    u32 si = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (si >= nstreams) return;
    if (res[si].contiguous || res[si].err) return;
    ZDStreamInfo I = info[si];
    u64 run = I.out_base;
    for (u32 f = 0; f < I.nframes; f++) {
        ZDFrame F = frames[I.frame_base + f];
        if (F.dst_off != run) {
            u8 *d = out + run;
            const u8 *s = out + F.dst_off;
            for (u64 i = 0; i < F.out_size; i += 32) {
                u8 v = 0;
                if (i + lane < F.out_size) v = s[i + lane];
                __syncwarp();
                if (i + lane < F.out_size) d[i + lane] = v;
                __syncwarp();
            }
            if (lane == 0) frames[I.frame_base + f].dst_off = run;
        }
        run += F.out_size;
    }
}

// ---------------------------------------------------------------------------------- host launchers
void fqz_launch_zd_compact(ZDFrame *frames, const ZDStreamInfo *info, const ZDStreamResult *res, u32 nstreams, u8 *out, cudaStream_t s) {
    if (!nstreams) return;
    FQZ_LAUNCH(k_zd_compact, (nstreams * 32 + 127) / 128, 128, 0, s, frames, info, res, nstreams, out);
}
void fqz_launch_zd_hop(const ZDStream *streams, u32 nstreams, ZDStreamInfo *info, ZDFrame *frames, ZDBlock *blocks, int fill, cudaStream_t s) {
    if (!nstreams) return;
    FQZ_LAUNCH(k_zd_hop, (nstreams + ZD_HOP_WARPS - 1) / ZD_HOP_WARPS, ZD_HOP_WARPS * 32, 0, s, streams, nstreams, info, frames, blocks, fill);
}
void fqz_launch_zd_parse(ZDBlock *blocks, u32 nblocks, u32 *cnt, u32 cnt_stride, cudaStream_t s) {
    if (!nblocks) return;
    FQZ_LAUNCH(k_zd_parse, (nblocks + 127) / 128, 128, 0, s, blocks, nblocks, cnt, cnt_stride);
}
void fqz_launch_zd_link(ZDFrame *frames, u32 nframes, ZDBlock *blocks, u32 *cnt, u32 cnt_stride, cudaStream_t s) {
    if (!nframes) return;
    FQZ_LAUNCH(k_zd_link, (nframes + 127) / 128, 128, 0, s, frames, nframes, blocks, cnt, cnt_stride);
}
void fqz_launch_zd_offsets(ZDBlock *blocks, u32 nblocks, const u32 *cnt, u32 cnt_stride, u32 *seqblk, u32 *litgrp, cudaStream_t s) {
    if (!nblocks) return;
    FQZ_LAUNCH(k_zd_offsets, (nblocks + 127) / 128, 128, 0, s, blocks, nblocks, cnt, cnt_stride, seqblk, litgrp);
}
void fqz_launch_zd_literals(ZDBlock *blocks, u32 nblocks, const u32 *litgrp, u32 ngroups, const ZDFrame *frames, u8 *litbuf, u8 *out, cudaStream_t s) {
    if (!nblocks || !ngroups) return;
    FQZ_LAUNCH(k_zd_literals, (ngroups + ZD_WARPS - 1) / ZD_WARPS, ZD_WARPS * 32, 0, s, blocks, nblocks, litgrp, ngroups, frames, litbuf, out);
}
void fqz_launch_zd_sequences(ZDBlock *blocks, const ZDFrame *frames, const u32 *seqblk, u32 nsb, const u32 *slot_of, u32 *tabs, u32 *seqbuf, cudaStream_t s) {
    if (!nsb) return;
    FQZ_LAUNCH(k_zd_seq_tables, (nsb + ZD_WARPS - 1) / ZD_WARPS, ZD_WARPS * 32, 0, s, blocks, seqblk, nsb, tabs);
    FQZ_LAUNCH(k_zd_seq_decode, (nsb + 127) / 128, 128, 0, s, blocks, frames, seqblk, nsb, slot_of, tabs, seqbuf);
}
void fqz_launch_zd_execute(ZDFrame *frames, u32 nframes, ZDBlock *blocks, const u8 *litbuf, const u32 *seqbuf, u8 *out, cudaStream_t s) {
    if (!nframes) return;
    FQZ_LAUNCH(k_zd_execute, (nframes + ZD_WARPS - 1) / ZD_WARPS, ZD_WARPS * 32, ZX_SMEM, s, frames, nframes, blocks, litbuf, seqbuf, out);
}
void fqz_launch_zd_rawcopy(const ZDBlock *blocks, u32 nblocks, const ZDFrame *frames, u8 *out, cudaStream_t s) {
    if (!nblocks) return;
    FQZ_LAUNCH(k_zd_rawcopy, (nblocks * 32 + 255) / 256, 256, 0, s, blocks, nblocks, frames, out);
}
void fqz_launch_zd_checksum(ZDFrame *frames, u32 nframes, const u8 *out, cudaStream_t s) {
    if (!nframes) return;
    u32 threads = XX_WARPS * 32;
    FQZ_LAUNCH(k_zd_checksum, (nframes * 4 + threads - 1) / threads, threads, XX_SMEM, s, frames, nframes, out);
}
void fqz_launch_zd_finish(const ZDFrame *frames, const ZDStreamInfo *info, u32 nstreams, ZDStreamResult *res, cudaStream_t s) {
    if (!nstreams) return;
    FQZ_LAUNCH(k_zd_finish, (nstreams + 63) / 64, 64, 0, s, frames, info, nstreams, res);
}
int fqz_zstd_dec_init_device() {
    int e = 0;
    if (ZX_SMEM) e |= (int)cudaFuncSetAttribute(k_zd_execute, cudaFuncAttributeMaxDynamicSharedMemorySize, ZX_SMEM);
    e |= (int)cudaFuncSetAttribute(k_zd_checksum, cudaFuncAttributeMaxDynamicSharedMemorySize, XX_SMEM);
    return e;
}
