// fqz_backend.cu — decompress back end: six decoded streams per block -> FASTQ text, bit-exact
// with blockReader.writeRecord (internal/compress/compress.go:944-1078) and
// encoder.AppendUnpackBases / DeltaDecode / DenormalizeQuality (sequence.go:188-223,
// quality.go:66-75,107-118).
//
//   k_walk_prefixes   the header / plus / N-position streams are chains of u16-length-prefixed
//                     items with no index (compress.go:977-1015,1055-1078): one warp per
//                     (block, stream) stages the stream through shared memory and one lane hops
//                     from prefix to prefix, emitting one offset per record
//   k_record_sizes    per record: L, packed size, FASTQ bytes; validation ("truncated ... data")
//   k_emit_fastq      16 lanes per record: '@'+header, unpacked bases with N restored, '+'+payload,
//                     prefix-summed qualities; all wide stores aligned to the destination
#include "fqz_backend.h"
#include "fqz_zstd.h"

#define ZW_MAXC 8u     // candidate item starts kept per lane (128 bytes)
#define WALK_CH 4096u  // bytes per staged chunk (plus 16 bytes of overlap so a prefix never straddles)

// kind: 0 headers, 1 plus, 2 npos.  offs has nrec+1 entries per (block, kind): item r starts at offs[r].
// One warp per chain.  The stream is staged chunk by chunk into shared memory by the TMA unit (1-D
// bulk copies, double buffered: chunk c+1 lands while chunk c is walked); the warp then advances by
// verified runs of equal-length items (see below) instead of one serial hop per item.
struct WalkSmem {
    uint4 buf[4][2][(WALK_CH + 16) / 16];
    u64 bar[4][2];
    u16 cand[4][ZW_MAXC][32];
};
// Walks items r .. rend-1 of the chain p[0..size) starting at byte `pos`; writes offs[r'] for each and
// returns the position behind the last item walked in *pos_out, the items walked in *r_out, and a
// truncation flag in *err_out.  One warp; every lane returns the same values.
__device__ static void walk_chain(const u8 *p, u32 size, u32 kind, u32 pos, u32 r, const u32 nrec, u32 *offs, WalkSmem &SM, u32 *pos_out,
                                  u32 *r_out, u32 *err_out) {
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    uint4(*sm_buf)[2][(WALK_CH + 16) / 16] = SM.buf;
    u16(*sm_cand)[ZW_MAXC][32] = SM.cand;
    u64 *bar = SM.bar[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncwarp();
    u32 nchunks = (size + WALK_CH - 1) / WALK_CH;
    u32 ph[2] = {0, 0};
    bool pending[2] = {false, false};  // uniform across the warp; lane 0 talks to the TMA unit, every lane waits on the barrier
    auto issue = [&](u32 c, u32 slot) {
        if (c >= nchunks) return;
        if (lane == 0) {
            u32 off = c * WALK_CH;
            u32 bytes = min(WALK_CH + 16u, (size - off + 15u) & ~15u);  // the stream carries FQZ_PAD readable bytes of slack
            mbar_expect_tx(&bar[slot], bytes);
            tma_load_1d(sm_buf[warp][slot], p + off, bytes, &bar[slot]);
#ifdef FQZ_EMU
            bar[slot] += 1;
#endif
        }
        pending[slot] = true;
    };
    auto wait = [&](u32 slot) {
        if (!pending[slot]) return;
        __syncwarp();
        mbar_wait(&bar[slot], ph[slot]);
        __syncwarp();  // every lane has seen this phase complete before lane 0 may re-arm the barrier
        ph[slot] ^= 1u;
        pending[slot] = false;
    };
    u32 err = 0;
    u32 c = pos / WALK_CH, slot = 0;
    issue(c, 0);
    issue(c + 1, 1);
    while (r < nrec) {
        if (pos + 2 > size) { err = 1; break; }  // no room for the length prefix of item r
        u32 nc = pos / WALK_CH;
        if (nc != c) {
            if (nc == c + 1) {  // the prefetched chunk; refill the slot just left
                slot ^= 1u;
                issue(nc + 1, slot ^ 1u);
            } else {  // a long item jumped over the prefetched chunk
                wait(slot ^ 1u);
                issue(nc, slot);
                issue(nc + 1, slot ^ 1u);
            }
            c = nc;
        }
        wait(slot);
        const u8 *buf = (const u8 *)sm_buf[warp][slot];
        u32 cbase = c * WALK_CH;
        u32 lim = min(size, cbase + WALK_CH + 16u);  // bytes staged in this slot
        if (kind == 0) {
            // Headers: printable text behind a u16 length < 256, so the only zero bytes of the stream
            // are the high bytes of the length prefixes.  Every lane scans 128 staged bytes for them
            // (candidate item starts), the warp then PROVES the guess: the first candidate must be
            // pos and every candidate must end exactly where the next one starts.  A chunk that
            // fails the proof (binary headers, lengths >= 256, > ZW_MAXC items per 128 bytes) is
            // walked serially instead, so the result never depends on the guess.
            u32 rend = min(size, cbase + WALK_CH);
            u16 *cl = &sm_cand[warp][0][lane];  // candidate j of this lane at cl[32 * j]: conflict-free
            u32 k = 0;
            {
                // this lane owns the zero bytes at [q0, q0 + 128); it visits its 32 words in an order
                // rotated by the lane index so that the 32 lanes always hit 32 different banks
                u32 q0 = cbase + 128u * lane;
                const u32 *wbuf = (const u32 *)buf + 32u * lane;
                u64 zlo = 0, zhi = 0;  // one bit per byte of the 128
                if (q0 < lim) {
#pragma unroll 8
                    for (u32 w = 0; w < 32; w++) {
                        u32 idx = (w + lane) & 31u;
                        u32 x = wbuf[idx];
                        u32 nib = ((__vcmpeq4(x, 0u) & 0x01010101u) * 0x01020408u) >> 24;  // 4 flags -> 4 bits
                        u64 sh = (u64)(nib & 0xFu) << (4u * (idx & 15u));
                        if (idx < 16) zlo |= sh; else zhi |= sh;
                    }
                }
                for (int half = 0; half < 3; half++) {
                    // third round: the zero byte at offset WALK_CH (in the 16 staged overlap bytes) belongs to lane 31
                    u64 m = half == 0 ? zlo : (half == 1 ? zhi : ((lane == 31 && buf[WALK_CH] == 0) ? 1ull : 0ull));
                    while (m) {
                        u32 b = (u32)__ffsll((long long)m) - 1u;
                        m &= m - 1ull;
                        u32 zpos = q0 + 64u * half + b;  // position of the zero byte = candidate start + 1
                        if (zpos >= pos + 1u && zpos - 1u < rend && zpos + 1u <= size && zpos < lim) {
                            if (k < ZW_MAXC) cl[32u * k] = (u16)(zpos - 1u - cbase);
                            k++;
                        }
                    }
                }
            }
            u32 over = __ballot_sync(0xffffffffu, k > ZW_MAXC);
            u32 havem = __ballot_sync(0xffffffffu, k > 0);
            u32 incl = group_incl_scan(k, 0xffffffffu, 32);
            u32 total = __shfl_sync(0xffffffffu, incl, 31);
            u32 ex = incl - k;
            bool good = (over == 0) && total > 0;
            // first candidate of the next lane that has one
            u32 above = (lane == 31) ? 0u : (havem >> (lane + 1u));
            int nl = above ? (int)lane + __ffs((int)above) : -1;
            u32 c0 = k ? (u32)cl[0] : 0u;
            u32 succ_other = __shfl_sync(0xffffffffu, c0, nl < 0 ? 0 : nl);
            bool ok = true;
            u32 last_next = 0;
            if (good && k) {
                for (u32 j = 0; j < k; j++) {
                    u32 a = cl[32u * j];
                    u32 nx = a + 2u + (u32)buf[a];  // relative to cbase
                    if (j + 1 < k) ok = ok && (nx == (u32)cl[32u * (j + 1)]);
                    else if (nl >= 0) ok = ok && (nx == succ_other);
                    else {
                        ok = ok && (cbase + nx <= size);
                        last_next = cbase + nx;
                    }
                }
                if (ex == 0) ok = ok && (cbase + (u32)cl[0] == pos);
            }
            good = good && __all_sync(0xffffffffu, ok);
#ifdef FQZ_EMU
            if (lane == 0 && getenv("FQZ_DEBUG")) fprintf(stderr, "walk hdr chunk %u pos %u total %u good %d\n", c, pos, total, (int)good);
#endif
            if (good) {
                u32 acc = min(total, nrec - r);
                for (u32 j = 0; j < k; j++)
                    if (ex + j < acc) offs[r + ex + j] = cbase + (u32)cl[32u * j];
                if (acc < total) {
                    // the chain ends inside this chunk (NumRecords counts fewer items than the stream holds: the reference
                    // reads NumRecords of them and ignores the rest, compress.go:944-985): stop at the start of candidate acc
                    bool mine = acc >= ex && acc < ex + k;
                    u32 mm = __ballot_sync(0xffffffffu, mine);
                    u32 v = mine ? cbase + (u32)cl[32u * (acc - ex)] : 0u;
                    pos = __shfl_sync(0xffffffffu, v, __ffs((int)mm) - 1);
                } else {
                    int lastl = 31 - __clz((int)havem);
                    pos = __shfl_sync(0xffffffffu, last_next, lastl);
                }
                r += acc;
                continue;  // (when acc < total the loop ends: r == nrec)
            }
            // proof failed: lane 0 hops through the rest of this chunk
            if (lane == 0) {
                u32 cend = cbase + WALK_CH;
                while (r < nrec && pos < cend && pos + 2 <= size) {
                    u32 len = (u32)buf[pos - cbase] | ((u32)buf[pos - cbase + 1] << 8);
                    u32 next = pos + 2 + len;
                    if (next > size) { err = 1; break; }
                    offs[r++] = pos;
                    pos = next;
                }
            }
            pos = __shfl_sync(0xffffffffu, pos, 0);
            r = __shfl_sync(0xffffffffu, r, 0);
            err = __shfl_sync(0xffffffffu, err, 0);
            if (err) break;
            continue;
        }
        // The hop pos -> pos + 2 + len is serial, but consecutive items very often have the same
        // length (bare "+" lines, reads without N, headers of equal width): speculate that the next
        // 32 items are as long as this one, let every lane check one of them, accept the verified run.
        u32 len0 = (u32)buf[pos - cbase] | ((u32)buf[pos - cbase + 1] << 8);
        u32 isz = 2u + (kind == 2 ? 2u * len0 : len0);
        u32 cand = pos + lane * isz;
        bool ok = false;
        if (cand + 2u <= lim && cand < cbase + WALK_CH) {
            u32 l = (u32)buf[cand - cbase] | ((u32)buf[cand - cbase + 1] << 8);
            ok = (l == len0) && (cand + isz <= size);
        }
        u32 okm = __ballot_sync(0xffffffffu, ok);
        u32 run = (~okm == 0u) ? 32u : (u32)__ffs((int)~okm) - 1u;
        if (run == 0) {  // lane 0 failed: the item at pos runs past the end of the stream
            err = 1;
            break;
        }
        run = min(run, nrec - r);
        if (lane < run) offs[r + lane] = cand;
        r += run;
        pos += run * isz;
    }
    if (err && r < nrec && lane == 0) offs[r] = pos;  // the end of the last whole item (k_first_error reads item r - 1)
    wait(0);  // no bulk copy may still be in flight when the CTA's shared memory is released
    wait(1);
    __syncwarp();
    *pos_out = pos;
    *r_out = r;
    *err_out = err;
}

// index frame of a compressed item stream (FQZ_ZPOLICY_INDEX, fqz_zstd.h): usable for the segmented
// walk when it covers the decoded stream exactly.  Returns the frame count (0 = no usable hints).
__device__ __forceinline__ u32 bk_rd32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
__device__ static u32 walk_hint_frames(const u8 *h, u32 hsize, u32 size, u32 *fsz_out) {
    if (!h || hsize < FQZ_ZINDEX_HDR + 8u || size == 0) return 0;
    if (bk_rd32(h) != FQZ_ZINDEX_MAGIC || bk_rd32(h + 8) != FQZ_ZINDEX_SIG) return 0;
    u32 nf = bk_rd32(h + 12), fsz = bk_rd32(h + 16);
    if (nf == 0 || fsz == 0 || (u64)FQZ_ZINDEX_HDR + 8ull * nf > hsize || bk_rd32(h + 4) != FQZ_ZINDEX_HDR - 8u + 8u * nf) return 0;
    if ((u64)fsz * (nf - 1) >= size || (u64)fsz * nf < size) return 0;
    *fsz_out = fsz;
    return nf;
}

// Segmented walk: one warp per (block, chain, frame of the compressed stream).  The index frame says
// where the first item of every frame starts and how many items start in it; each warp walks its own
// items and PROVES its part: it must arrive exactly at the start the index gives for the next frame
// that has items (the end of the stream for the last one), and the counts must add up to nrec.  Every
// proven segment bumps ok[block * 3 + kind]; k_walk_prefixes redoes any chain whose proof is incomplete.
__global__ void __launch_bounds__(128) k_walk_segments(const BkBlock *blks, u32 nblocks, u32 *offs_base, u32 *ok) {
    __shared__ WalkSmem SM;
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 seg = blockIdx.x * 4 + warp;
    u32 b = blockIdx.y / 3, kind = blockIdx.y % 3;
    const BkBlock *Bp = blks + b;
    int sidx = kind == 0 ? 2 : (kind == 1 ? 3 : 4);
    const u8 *p = (const u8 *)(uintptr_t)Bp->stream[sidx];
    const u8 *h = (const u8 *)(uintptr_t)Bp->hint[kind];
    u32 size = Bp->size[sidx], nrec = Bp->nrec;
    u32 fsz = 0;
    u32 nf = walk_hint_frames(h, Bp->hint_size[kind], size, &fsz);
    if (seg >= nf) return;
    const u8 *ent = h + FQZ_ZINDEX_HDR;
    // records before this segment; first frame after it that has items
    u32 before = 0;
    for (u32 k = lane; k < seg; k += 32) before += bk_rd32(ent + 8ull * k + 4) >> 16;
    before = group_sum(before, 0xffffffffu, 32);
    u32 me = bk_rd32(ent + 8ull * seg + 4);
    u32 first = me & 0xFFFFu, cnt = me >> 16;
    bool good = !(first == 0xFFFFu && cnt == 0xFFFFu) && (cnt == 0 || first < fsz) && (u64)before + cnt <= nrec;
    u32 expect = size;  // where the walk must end
    bool last_items = true;
    if (good) {
        u32 nxt = 0xFFFFFFFFu;
        for (u32 k0 = seg + 1; k0 < nf && nxt == 0xFFFFFFFFu; k0 += 32) {
            u32 k = k0 + lane;
            u32 e = (k < nf) ? bk_rd32(ent + 8ull * k + 4) : 0u;
            u32 m = __ballot_sync(0xffffffffu, (e >> 16) != 0);
            if (m) {
                int l = __ffs((int)m) - 1;
                u32 el = __shfl_sync(0xffffffffu, e, l);
                nxt = (k0 + (u32)l) * fsz + (el & 0xFFFFu);
            }
        }
        if (nxt != 0xFFFFFFFFu) {
            expect = nxt;
            last_items = false;
        }
    }
    u32 *offs = offs_base + (3ull * Bp->rec_base + (u64)kind * nrec) + 3ull * b + kind;
    if (good && cnt) {
        u32 start = seg * fsz + first;
        if (before == 0 && start != 0) good = false;  // the chain starts at byte 0
        if (good) {
            u32 pos = 0, r = 0, err = 0;
#ifdef FQZ_EMU
            if (getenv("FQZ_DEBUG")) fprintf(stderr, "seg b %u kind %u seg %u lane %u start %u before %u cnt %u expect %u\n", b, kind, seg, lane, start, before, cnt, expect);
#endif
            walk_chain(p, size, kind, start, before, before + cnt, offs, SM, &pos, &r, &err);
            good = !err && r == before + cnt && pos == expect;
            if (good && last_items) {
                good = (r == nrec);
                if (good && lane == 0) offs[nrec] = pos;
            }
        }
    } else if (good && last_items && before != nrec)
        good = false;  // nothing starts here or later, yet records are missing
    if (good && lane == 0) atomicAdd(&ok[b * 3 + kind], 1u);
}

__global__ void __launch_bounds__(128) k_walk_prefixes(const BkBlock *blks, u32 nblocks, u32 *offs_base, const u32 *ok, FqzDecStatus *st) {
    __shared__ WalkSmem SM;
    u32 warp = threadIdx.x >> 5, lane = lane_id();
    u32 wi = blockIdx.x * 4 + warp;
    if (wi >= nblocks * 3) return;
    u32 b = wi / 3, kind = wi % 3;
    const BkBlock *Bp = blks + b;
    int sidx = kind == 0 ? 2 : (kind == 1 ? 3 : 4);
    const u8 *p = (const u8 *)(uintptr_t)Bp->stream[sidx];
    u32 size = Bp->size[sidx], nrec = Bp->nrec;
    u64 rec_base = Bp->rec_base;
    u32 *offs = offs_base + (3ull * rec_base + (u64)kind * nrec) + 3ull * b + kind;  // (nrec+1) entries per kind
    if (kind == 1 && size == 0) {  // v1 files / empty plus stream: every plus line is "+" (compress.go:995-999)
        for (u32 r = lane; r <= nrec; r += 32) offs[r] = 0;
        return;
    }
    if (ok) {  // every segment of this chain proved its part (k_walk_segments): nothing to redo
        u32 fsz = 0;
        u32 nf = walk_hint_frames((const u8 *)(uintptr_t)Bp->hint[kind], Bp->hint_size[kind], size, &fsz);
        if (nf && ok[b * 3 + kind] == nf) return;
    }
    u32 pos = 0, r = 0, err = 0;
    walk_chain(p, size, kind, 0, 0, nrec, offs, SM, &pos, &r, &err);
    if (lane == 0) {
        offs[nrec] = pos;
        if (err) {
            u32 code = kind == 0 ? BK_E_TRUNC_HEADER : (kind == 1 ? BK_E_TRUNC_PLUS : BK_E_TRUNC_NPOS);
            atomicMin(&st->err_key, ((u64)(rec_base + r) << 8) | code);
        }
    }
}

// Per record: sizes for the scans.  sz[0*stride + R] = packed bytes, sz[1*stride + R] = FASTQ bytes,
// sz[2*stride + R] = bases (R = window-global record index).  64-bit per-block totals let the host
// rule out u32 wrap-around before the scans are trusted.
__global__ void __launch_bounds__(256)
k_record_sizes(const BkBlock *blks, u32 nblocks, const u32 *offs_base, u32 *sz, u64 stride, BkTotals *tot, FqzDecStatus *st) {
    u32 b = blockIdx.y;
    BkBlock B = blks[b];
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = r < B.nrec;
    u64 R = B.rec_base + r;
    u32 L = 0, fq = 0, pk = 0;
    if (live) {
        const u32 *oh = offs_base + 3ull * B.rec_base + 3ull * b;
        const u32 *op = oh + (B.nrec + 1);
        u32 code = 0;
        if (4ull * r + 4 > B.size[5]) code = BK_E_TRUNC_LEN;  // compress.go:1047
        else L = *(const u32 *)((const u8 *)(uintptr_t)B.stream[5] + 4ull * r);
        u32 H = oh[r + 1] - oh[r] - 2;
        u32 P = B.size[3] ? op[r + 1] - op[r] - 2 : 0u;
        if (code) {
            atomicMin(&st->err_key, (R << 8) | code);
            L = 0;
        }
        pk = (u32)(((u64)L + 3u) >> 2);
        fq = H + P + 6u;
        sz[0 * stride + R] = pk;
        sz[1 * stride + R] = fq + 2u * L;
        sz[2 * stride + R] = L;
    }
    // warp totals -> one atomic per warp
    u64 tp = pk, tf = (u64)fq + 2ull * L, tb = L;
    for (int d = 16; d > 0; d >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, d);
        tf += __shfl_xor_sync(0xffffffffu, tf, d);
        tb += __shfl_xor_sync(0xffffffffu, tb, d);
    }
    if (lane_id() == 0 && tf) {
        atomicAdd(&tot[b].packed, tp);
        atomicAdd(&tot[b].fastq, tf);
        atomicAdd(&tot[b].bases, tb);
    }
}

// Error path only.  The walks and k_record_sizes have found a failure (a chain or the length stream runs short at some
// record); a sequence or quality stream that runs short, or an N position beyond its read, at an EARLIER point of the
// reference's record order must win (compress.go:944-1078 stops at the first).  One warp per block goes through the
// records in order, 32 at a time, with 64-bit running offsets (lengths of a damaged stream may add up to anything), up to
// the record of the failure known so far, and at that record only through the checks that come before the known one.
// limit_key: the failure known when the kernel was launched (~0: none; then every record is checked).
__global__ void __launch_bounds__(128) k_first_error(const BkBlock *blks, u32 nblocks, const u32 *offs_base, u64 limit_key, FqzDecStatus *st) {
    u32 b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
    if (b >= nblocks) return;
    BkBlock B = blks[b];
    const u64 lim_rec = limit_key >> 8;
    const u32 lim_code = limit_key == ~0ull ? 8u : (u32)(limit_key & 0xFF);
    if (B.rec_base > lim_rec) return;
    const u32 *on = offs_base + 3ull * B.rec_base + 3ull * b + 2ull * (B.nrec + 1);  // N-position items (kind 2)
    const u8 *lens = (const u8 *)(uintptr_t)B.stream[5];
    const u8 *nps = (const u8 *)(uintptr_t)B.stream[4];
    u64 so = 0, qo = 0;  // bytes of packed bases / qualities in front of the 32 records at hand
    for (u32 base = 0; base < B.nrec; base += 32) {
        u32 r = base + lane;
        u64 R = B.rec_base + r;
        bool live = r < B.nrec && R <= lim_rec;
        u32 cap = (live && R == lim_rec) ? lim_code : 8u;  // checks with a kind below cap come first in this record
        if (cap <= BK_E_TRUNC_SEQ) live = false;           // length, N-position or header failure: nothing of ours in front
        u64 L = live ? (u64) * (const u32 *)(lens + 4ull * r) : 0ull;
        u64 pk = (L + 3u) >> 2;
        u64 ipk = pk, iL = L;  // inclusive sums over the lanes
        for (int d = 1; d < 32; d <<= 1) {
            u64 t0 = __shfl_up_sync(0xffffffffu, ipk, (unsigned)d), t1 = __shfl_up_sync(0xffffffffu, iL, (unsigned)d);
            if (lane >= (u32)d) {
                ipk += t0;
                iL += t1;
            }
        }
        u32 code = 0;
        if (live) {
            if (so + ipk > B.size[0]) code = BK_E_TRUNC_SEQ;
            else if (cap > BK_E_NPOS_RANGE) {
                u32 a = on[r], nn = (on[r + 1] - a - 2) >> 1;
                for (u32 k = 0; k < nn && !code; k++)
                    if (((u32)nps[a + 2 + 2 * k] | ((u32)nps[a + 3 + 2 * k] << 8)) >= L) code = BK_E_NPOS_RANGE;
            }
            if (!code && cap > BK_E_TRUNC_QUAL && qo + iL > B.size[1]) code = BK_E_TRUNC_QUAL;
        }
        u32 hit = __ballot_sync(0xffffffffu, code != 0);
        if (hit) {  // the lowest record decides; a sequence stream short at a later record is short at every one behind it
            if (lane == (u32)__ffs((int)hit) - 1u) atomicMin(&st->err_key, (R << 8) | code);
            return;
        }
        so += __shfl_sync(0xffffffffu, ipk, 31);
        qo += __shfl_sync(0xffffffffu, iL, 31);
        if (B.rec_base + base + 31 >= lim_rec) return;
    }
}

// ---------------------------------------------------------------------------------- emit
#ifndef FQZ_EM_GROUP
#define FQZ_EM_GROUP 8  // lanes per record (8: 7.2 ms per 9.2 GB step, 16: 10.6 ms on B200)
#endif
template <int W>
__device__ __forceinline__ void bk_group_copy(u8 *dst, const u8 *src, u32 n, u32 g) {
    u32 head = (u32)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if (head > n) head = n;
    if (g < head) dst[g] = src[g];
    u32 nw = (n - head) >> 2;
    for (u32 w = g; w < nw; w += W) *(u32 *)(dst + head + 4u * w) = ld_u32_unaligned(src + head + 4u * w);
    u32 t0 = head + 4u * nw;
    if (g < n - t0) dst[t0 + g] = src[t0 + g];
}
// four consecutive bases starting at base index i of a packed record -> 4 ASCII bytes
__device__ __forceinline__ u32 unpack4(const u8 *packed, u32 i) {
    u32 bit = 2u * i;
    u32 x = ld_u32_unaligned(packed + (bit >> 3)) >> (bit & 7u);
    u32 b = x & 0xFFu;
    u32 sel = (b & 3u) | ((b & 0xCu) << 2) | ((b & 0x30u) << 4) | ((b & 0xC0u) << 6);
    return __byte_perm(0x54474341u, 0u, sel);  // "ACGT"
}
// inclusive byte-wise prefix sum (mod 256) inside a word
__device__ __forceinline__ u32 prefix4(u32 x) {
    x = __vadd4(x, x << 8);
    return __vadd4(x, x << 16);
}

__global__ void __launch_bounds__(256)
k_emit_fastq(const BkBlock *blks, u32 nblocks, const u32 *offs_base, const u32 *sc, u64 stride, u32 phred64, u8 *out, FqzDecStatus *st) {
    const int W = FQZ_EM_GROUP;
    u32 b = blockIdx.y;
    BkBlock B = blks[b];
    u32 g = threadIdx.x & (W - 1);
    u32 gm = group_mask(W);
    u32 r = blockIdx.x * (blockDim.x / W) + threadIdx.x / W;
    if (r >= B.nrec) return;
    const u32 *oh = offs_base + 3ull * B.rec_base + 3ull * b;
    const u32 *op = oh + (B.nrec + 1);
    const u32 *on = op + (B.nrec + 1);
    u64 R = B.rec_base + r;
    u32 L = sc[2 * stride + R + 1] - sc[2 * stride + R];
    u32 o_seq = sc[0 * stride + R] - sc[0 * stride + B.rec_base];
    u32 o_qual = sc[2 * stride + R] - sc[2 * stride + B.rec_base];
    u8 *d = out + sc[1 * stride + R];
    // "truncated sequence data" / "truncated quality data" (compress.go:1019,1032): the record is skipped
    if ((u64)o_seq + ((L + 3u) >> 2) > B.size[0]) {
        if (g == 0) atomicMin(&st->err_key, (R << 8) | BK_E_TRUNC_SEQ);
        return;
    }
    if ((u64)o_qual + L > B.size[1]) {
        if (g == 0) atomicMin(&st->err_key, (R << 8) | BK_E_TRUNC_QUAL);
        return;
    }
    const u8 *hs = (const u8 *)(uintptr_t)B.stream[2] + oh[r] + 2;
    u32 H = oh[r + 1] - oh[r] - 2;
    // '@' header '\n'
    if (g == 0) d[0] = '@';
    bk_group_copy<W>(d + 1, hs, H, g);
    if (g == 0) d[1 + H] = '\n';
    d += H + 2;
    // sequence: dst-aligned words of 4 bases
    {
        const u8 *pk = (const u8 *)(uintptr_t)B.stream[0] + o_seq;
        u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
        if (head > L) head = L;
        if (g < head) d[g] = (u8)(unpack4(pk, g) & 0xFFu);
        u32 nw = (L - head) >> 2;
        for (u32 w = g; w < nw; w += W) *(u32 *)(d + head + 4u * w) = unpack4(pk, head + 4u * w);
        u32 t0 = head + 4u * nw;
        if (g < L - t0) d[t0 + g] = (u8)(unpack4(pk, t0 + g) & 0xFFu);
        __syncwarp(gm);
        // restore N (sequence.go:217-220)
        const u8 *np = (const u8 *)(uintptr_t)B.stream[4] + on[r];
        u32 nn = (on[r + 1] - on[r] - 2) >> 1;
        for (u32 k = g; k < nn; k += W) {
            u32 pos = (u32)np[2 + 2 * k] | ((u32)np[3 + 2 * k] << 8);
            if (pos < L) d[pos] = 'N';
            else atomicMin(&st->err_key, (R << 8) | BK_E_NPOS_RANGE);  // the reference panics here
        }
        if (g == 0) d[L] = '\n';
        d += L + 1;
    }
    // '+' payload '\n'
    {
        u32 P = 0;
        const u8 *ps = nullptr;
        if (B.size[3]) {
            P = op[r + 1] - op[r] - 2;
            ps = (const u8 *)(uintptr_t)B.stream[3] + op[r] + 2;
        }
        if (g == 0) d[0] = '+';
        if (P) bk_group_copy<W>(d + 1, ps, P, g);
        if (g == 0) d[1 + P] = '\n';
        d += P + 2;
    }
    // quality: running sum mod 256 from the record start, plus the Phred offset
    {
        const u8 *q = (const u8 *)(uintptr_t)B.stream[1] + o_qual;
        u32 offw = (phred64 ? 64u : 33u) * 0x01010101u;
        u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
        if (head == 0) head = 4;
        if (head > L) head = L;
        u32 nunits = L ? 1u + ((L - head + 3u) >> 2) : 0u;
        u32 rounds = (nunits + W - 1) / W;
        u32 carry = 0;
        for (u32 j = 0; j < rounds; j++) {
            u32 u = j * W + g;
            u32 x = 0, i0 = 0, cnt = 0;
            if (u < nunits) {
                if (u == 0) { i0 = 0; cnt = head; }
                else { i0 = head + 4u * (u - 1); cnt = min(4u, L - i0); }
                x = ld_u32_unaligned(q + i0);
                if (cnt < 4) x &= 0xFFFFFFFFu >> (8u * (4u - cnt));
            }
            u32 pre = prefix4(x);
            u32 tot = (pre >> 24) & 0xFFu;            // sum of the unit's bytes (zero-padded)
            u32 incl = group_incl_scan(tot, gm, W);
            u32 before = (carry + incl - tot) & 0xFFu;
            u32 v = __vadd4(__vadd4(pre, before * 0x01010101u), offw);
            if (u < nunits) {
                if (cnt == 4 && u > 0) *(u32 *)(d + i0) = v;
                else
                    for (u32 k = 0; k < cnt; k++) d[i0 + k] = (u8)(v >> (8 * k));
            }
            carry = (carry + __shfl_sync(gm, incl, W - 1, W)) & 0xFFu;
        }
        if (g == 0) d[L] = '\n';
    }
}

// ---------------------------------------------------------------------------------- container walk
// Serial hop over the block headers of a .fqz resident in HBM (the format has no index:
// readNextDecompressJob, compress.go:721-758; header layouts container.go:116-152).  Stops after
// `cap` blocks or once the entered blocks span more than max_bytes.
__global__ void k_walk_container(const u8 *f, u64 n, u64 pos, u32 version, FqzBlockEntry *tab, u32 cap, u64 max_bytes, FqzWalkResult *res) {
    if (threadIdx.x || blockIdx.x) return;
    u32 hsz = version == 1 ? 32u : 36u;
    u32 nb = 0, status = 0;
    u64 start = pos;
    while (pos < n && nb < cap) {
        if (n - pos < hsz) { status = 1; break; }  // "reading block header: unexpected EOF"
        u32 v[9];
        for (u32 i = 0; i < hsz / 4; i++) {
            const u8 *q = f + pos + 4 * i;
            v[i] = (u32)q[0] | ((u32)q[1] << 8) | ((u32)q[2] << 16) | ((u32)q[3] << 24);
        }
        FqzBlockEntry e;
        e.nrec = v[0];
        e.pad = 0;
        if (version == 1) {  // seq, qual, headers, npos, lens (container.go:85-96)
            e.size[0] = v[1]; e.size[1] = v[2]; e.size[2] = v[3]; e.size[3] = 0; e.size[4] = v[4]; e.size[5] = v[5];
        } else {
            for (int i = 0; i < 6; i++) e.size[i] = v[1 + i];
        }
        u64 payload = 0;
        for (int i = 0; i < 6; i++) payload += e.size[i];
        if (n - pos - hsz < payload) { status = 1; break; }  // "reading compressed data: unexpected EOF"
        if (nb > 0 && pos + hsz + payload - start > max_bytes) break;
        e.payload_off = pos + hsz;
        tab[nb++] = e;
        pos += hsz + payload;
    }
    res->next_pos = pos;
    res->nblocks = nb;
    res->status = status;
    res->done = (pos >= n && !status) ? 1u : 0u;
    res->pad = 0;
}

// ---------------------------------------------------------------------------------- host launchers
void fqz_launch_walk_container(const u8 *fqz, u64 n, u64 pos, u32 version, FqzBlockEntry *table, u32 cap, u64 max_bytes,
                               FqzWalkResult *res, cudaStream_t s) {
    FQZ_LAUNCH(k_walk_container, 1, 32, 0, s, fqz, n, pos, version, table, cap, max_bytes, res);
}
void fqz_launch_walk_prefixes(const BkBlock *blks, u32 nblocks, u32 max_segments, u32 *offs, u32 *ok, FqzDecStatus *st, cudaStream_t s) {
    if (!nblocks) return;
    if (max_segments && ok) FQZ_LAUNCH(k_walk_segments, dim3((max_segments + 3) / 4, nblocks * 3), 128, 0, s, blks, nblocks, offs, ok);
    FQZ_LAUNCH(k_walk_prefixes, (nblocks * 3 + 3) / 4, 128, 0, s, blks, nblocks, offs, max_segments ? ok : nullptr, st);
}
void fqz_launch_record_sizes(const BkBlock *blks, u32 nblocks, u32 max_nrec, const u32 *offs, u32 *sz, u64 stride, BkTotals *tot,
                             FqzDecStatus *st, cudaStream_t s) {
    if (!nblocks || !max_nrec) return;
    FQZ_LAUNCH(k_record_sizes, dim3((max_nrec + 255) / 256, nblocks), 256, 0, s, blks, nblocks, offs, sz, stride, tot, st);
}
void fqz_launch_first_error(const BkBlock *blks, u32 nblocks, const u32 *offs, u64 limit_key, FqzDecStatus *st, cudaStream_t s) {
    if (!nblocks) return;
    FQZ_LAUNCH(k_first_error, (nblocks + 3) / 4, 128, 0, s, blks, nblocks, offs, limit_key, st);
}
void fqz_launch_emit(const BkBlock *blks, u32 nblocks, u32 max_nrec, const u32 *offs, const u32 *sc, u64 stride, u32 phred64, u8 *out,
                     FqzDecStatus *st, cudaStream_t s) {
    if (!nblocks || !max_nrec) return;
    const u32 per = 256 / FQZ_EM_GROUP;  // records per CTA
    FQZ_LAUNCH(k_emit_fastq, dim3((max_nrec + per - 1) / per, nblocks), 256, 0, s, blks, nblocks, offs, sc, stride, phred64, out, st);
}
