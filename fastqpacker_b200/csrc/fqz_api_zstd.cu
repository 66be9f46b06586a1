// fqz_api_zstd.cu — host orchestration of the zstd encode stage, container assembly
// (block headers + payload order of internal/compress/compress.go:530-552,
// internal/fqformat/container.go:83-113) and the compress entry points.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fqz_host.h"
#include "fqz_zstd.h"

// ---------------------------------------------------------------------------------- assembly kernels
// Copies every encoded frame from its slot to its final position.  final offset of frame f =
// fixed[f] (headers in front of it) + scan[f] (compressed bytes of the frames before it).
__global__ void __launch_bounds__(256) k_zassemble(const ZFrame *frames, u32 nframes, const u32 *scan, const u32 *fixed, const u8 *slots, u8 *out) {
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= nframes) return;
    u32 size = scan[warp + 1] - scan[warp];
    const u8 *src = slots + frames[warp].dst_off;  // 16-byte aligned slot
    u8 *dst = out + (u64)fixed[warp] + scan[warp];
    // 16-byte stores to the (arbitrarily aligned) destination; the slot is 16-byte aligned, so the
    // source of every vector is five aligned words merged with funnel shifts
    u32 head = (u32)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > size) head = size;
    if (lane < head) dst[lane] = src[lane];
    u32 nv = (size - head) >> 4;
    const u32 *sw = (const u32 *)src + (head >> 2);
    const u32 bs = (head & 3u) * 8u;
    uint4 *d16 = (uint4 *)(dst + head);
    for (u32 v = lane; v < nv; v += 32) {
        const u32 *sp = sw + 4u * v;
        u32 w0 = sp[0], w1 = sp[1], w2 = sp[2], w3 = sp[3], w4 = sp[4];  // w4 may lie in the slot's slack
        d16[v] = make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
    }
    u32 t0 = head + 16u * nv;
    if (lane < size - t0) dst[t0 + lane] = src[t0 + lane];
}

// 36-byte v2 block headers (container.go:97-109).  first[b*6+s] = index of the first frame of
// stream s of block b (first[nblocks*6] = nframes).
__global__ void k_write_block_headers(u32 nblocks, const u32 *first, const u32 *scan, const u32 *nrec, const u32 *orig, u32 base_off, u8 *out) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    u32 f0 = first[b * 6];
    u8 *h = out + (u64)base_off + 36ull * b + scan[f0];
    u32 v[9];
    v[0] = nrec[b];
    for (int s = 0; s < 6; s++) v[1 + s] = scan[first[b * 6 + s + 1]] - scan[first[b * 6 + s]];
    v[7] = orig[b];  // OriginalSeqSize
    v[8] = orig[b];  // OriginalQualSize (equal: every record has len(seq) == len(qual))
    for (int i = 0; i < 9; i++) {
        h[4 * i] = (u8)v[i];
        h[4 * i + 1] = (u8)(v[i] >> 8);
        h[4 * i + 2] = (u8)(v[i] >> 16);
        h[4 * i + 3] = (u8)(v[i] >> 24);
    }
}

// ---------------------------------------------------------------------------------- zstd batch
struct ZBatch {
    std::vector<ZFrame> frames;
    std::vector<u32> idx_ent, idx_lz, idx_items, idx_index;
    std::vector<ZRStream> rstreams;  // literals-only streams with record boundaries (duplicate-record search)
    u32 rblocks = 0, rmax_records = 0, rmax_blocks = 0;
    size_t slot_bytes = 0, ws_bytes = 0;
    // items / item_base / item_count: see ZFrame (policy FQZ_ZPOLICY_ITEMS only)
    // rec0 / nrec (policy ENTROPY with items): the block's records inside items[]
    void add_stream(const u8 *d_src, size_t len, int policy, const u32 *items = nullptr, u32 item_base = 0, u32 item_count = 0, u32 rec0 = 0,
                    u32 nrec = 0) {
        // literals-only stream with record boundaries: cut into segments of FQZ_ZSEG bytes, the units of the duplicate search
        const bool rseg = policy == FQZ_ZPOLICY_ENTROPY && items && len;
        const u32 rs_base = (u32)rstreams.size(), nseg = rseg ? (u32)((len + FQZ_ZSEG - 1) / FQZ_ZSEG) : 0u;
        const size_t fsz = (policy == FQZ_ZPOLICY_ENTROPY) ? FQZ_ZFRAME_ENT : (policy == FQZ_ZPOLICY_ITEMS ? FQZ_ZFRAME_ITEMS : FQZ_ZFRAME);
        const size_t nfr = (len + fsz - 1) / fsz;
        if (nfr >= FQZ_ZINDEX_MIN) {  // frame index in front of the stream (FQZ_ZPOLICY_INDEX)
            ZFrame f;
            f.src = 0;
            f.dst_off = slot_bytes;
            f.ws_off = 0;
            f.src_len = (u32)nfr;
            f.policy = FQZ_ZPOLICY_INDEX;
            f.items = 0;
            f.item_base = 0;
            f.item_count = 0;
            f.index_of = 0;
            f.pad = rseg ? rs_base + 1u : 0u;  // the stream's segments: rs_base .. rs_base + item_count - 1
            f.item_count = nseg;
            slot_bytes += (FQZ_ZINDEX_BYTES(nfr) + 15u) & ~(size_t)15;
            idx_index.push_back((u32)frames.size());
            frames.push_back(f);
        }
        const u32 index_of = (nfr >= FQZ_ZINDEX_MIN) ? (u32)frames.size() : 0u;  // 1 + number of the index frame just added
        for (u32 sg = 0; sg < nseg; sg++) {
            const size_t so = (size_t)sg * FQZ_ZSEG, sl = std::min<size_t>(FQZ_ZSEG, len - so);
            ZRStream r;
            r.src = (u64)(uintptr_t)(d_src + so);
            r.items = (u64)(uintptr_t)items;
            r.len = (u32)sl;
            r.item_base = item_base + (u32)so;
            r.rec0 = rec0;
            r.nrec = nrec;
            r.first_frame = (u32)frames.size() + (u32)(so / fsz);
            r.blk0 = rblocks;
            r.nblk = (u32)((sl + FQZ_ZBLOCK_ENT - 1) / FQZ_ZBLOCK_ENT);
            r.pad = 0;
            rblocks += r.nblk;
            rmax_records = std::max(rmax_records, nrec);
            rmax_blocks = std::max(rmax_blocks, r.nblk);
            rstreams.push_back(r);
        }
        for (size_t o = 0; o < len; o += fsz) {
            u32 l = (u32)std::min<size_t>(fsz, len - o);
            ZFrame f;
            f.src = (u64)(uintptr_t)(d_src + o);
            f.dst_off = slot_bytes;
            f.ws_off = 0;
            f.src_len = l;
            f.policy = (u32)policy;
            f.items = (u64)(uintptr_t)items;
            f.item_base = item_base + (u32)o;
            f.item_count = item_count;
            f.index_of = index_of;
            f.pad = rseg ? rs_base + (u32)(o / FQZ_ZSEG) + 1u : 0u;
            slot_bytes += FQZ_ZSLOT(l);
            if (policy == FQZ_ZPOLICY_AUTO || policy == FQZ_ZPOLICY_ITEMS) {
                f.ws_off = ws_bytes;
                ws_bytes += FQZ_ZWS(l);
                (policy == FQZ_ZPOLICY_ITEMS ? idx_items : idx_lz).push_back((u32)frames.size());
            } else
                idx_ent.push_back((u32)frames.size());
            frames.push_back(f);
        }
    }
};
struct ZEncoded {
    ZFrame *d_frames = nullptr;
    u8 *d_slots = nullptr;
    u32 *d_scan = nullptr;  // nframes+1, exclusive scan of frame sizes
    u32 nframes = 0;
};

// offs_base / offs_words: the scanned per-record offset arrays of the window (FrontOut::d_offs), when the
// literals-only frames carry item boundaries: enables the record matcher on them
static int zbatch_encode(fqz_ctx *c, ZBatch &zb, ZEncoded &ze, u64 src_bytes, const u32 *offs_base = nullptr, size_t offs_words = 0) {
    cudaStream_t s = c->stream;
    u32 nf = (u32)zb.frames.size();
    ze.nframes = nf;
    const u32 nrs = (offs_base && !c->opt_no_record_match) ? (u32)zb.rstreams.size() : 0u;
    const size_t idx_words = zb.idx_ent.size() + zb.idx_lz.size() + zb.idx_items.size() + zb.idx_index.size();
    const size_t rs_at = ((size_t)nf * sizeof(ZFrame) + idx_words * sizeof(u32) + 15u) & ~(size_t)15;
    size_t up = rs_at + (size_t)nrs * sizeof(ZRStream);
    FQZ_TRY(fqz_pin_reserve(c, 8192 + up));
    u8 *hp = c->h_pin + 4096;
    memcpy(hp, zb.frames.data(), (size_t)nf * sizeof(ZFrame));
    u32 *hidx = (u32 *)(hp + (size_t)nf * sizeof(ZFrame));
    if (!zb.idx_ent.empty()) memcpy(hidx, zb.idx_ent.data(), zb.idx_ent.size() * sizeof(u32));
    if (!zb.idx_lz.empty()) memcpy(hidx + zb.idx_ent.size(), zb.idx_lz.data(), zb.idx_lz.size() * sizeof(u32));
    if (!zb.idx_items.empty()) memcpy(hidx + zb.idx_ent.size() + zb.idx_lz.size(), zb.idx_items.data(), zb.idx_items.size() * sizeof(u32));
    const size_t index_at = zb.idx_ent.size() + zb.idx_lz.size() + zb.idx_items.size();
    if (!zb.idx_index.empty()) memcpy(hidx + index_at, zb.idx_index.data(), zb.idx_index.size() * sizeof(u32));
    if (nrs) memcpy(hp + rs_at, zb.rstreams.data(), (size_t)nrs * sizeof(ZRStream));
    u8 *d_up = (u8 *)c->arena.alloc(up + 16);
    ze.d_slots = (u8 *)c->arena.alloc(zb.slot_bytes + 16);
    u8 *d_ws = (u8 *)c->arena.alloc(zb.ws_bytes + 16);
    u32 *d_hash = (u32 *)c->arena.alloc((size_t)(nf + 1) * sizeof(u32));
    u32 *d_parsed = (u32 *)c->arena.alloc((zb.idx_items.size() + 1) * 2 * sizeof(u32));
    ze.d_scan = (u32 *)c->arena.alloc((size_t)(nf + 2) * sizeof(u32));
    if (!d_up || !ze.d_slots || !d_ws || !d_hash || !d_parsed || !ze.d_scan) {
        c->err = "arena: out of device memory (zstd stage)";
        return FQZ_E_CUDA;
    }
    FQZ_TRY(fqz_pin_copy(c, d_up, hp, up));
    ze.d_frames = (ZFrame *)d_up;
    u32 *d_idx = (u32 *)(d_up + (size_t)nf * sizeof(ZFrame));
    FQZ_CUDA_TRY(c, cudaMemsetAsync(ze.d_scan + nf, 0, sizeof(u32), s));
    u64 ent_bytes = 0, lz_bytes = 0;
    for (u32 i : zb.idx_ent) ent_bytes += zb.frames[i].src_len;
    for (u32 i : zb.idx_lz) lz_bytes += zb.frames[i].src_len;
    for (u32 i : zb.idx_items) lz_bytes += zb.frames[i].src_len;
    // The item-stream kernels (one warp per 16 KiB frame, ~1 ms of serial work each: every launch ends in a tail of
    // half-empty SMs) work on other frames than the literals-only coder, and only their last kernel needs the frame
    // checksums: they run on a second stream beside the checksum pass, the duplicate search and the literals-only coder,
    // the tail of one fills with CTAs of the other.  While profiling (per-stage CUDA events on one stream) all run in order.
    const bool fork = !c->prof.on && !c->opt_serial_entropy;
    cudaStream_t slz = fork ? c->stream_aux : s;
    if (fork) {
        FQZ_CUDA_TRY(c, cudaEventRecord(c->ev_fork, s));
        FQZ_CUDA_TRY(c, cudaStreamWaitEvent(slz, c->ev_fork, 0));
    }
    {
        StageScope sc(c, ST_XXH64, src_bytes);
        fqz_launch_xxh64(ze.d_frames, nf, d_hash, s);
    }
    if (fork) FQZ_CUDA_TRY(c, cudaEventRecord(c->ev_hash, s));
    {
        StageScope sc(c, ST_ZENC_LZ, lz_bytes);
        fqz_launch_zenc(ze.d_frames, d_idx + zb.idx_ent.size(), (u32)zb.idx_lz.size(), d_hash, ze.d_slots, d_ws, ze.d_scan, 1, nullptr, slz,
                        fork ? c->ev_hash : nullptr);
        fqz_launch_zenc(ze.d_frames, d_idx + zb.idx_ent.size() + zb.idx_lz.size(), (u32)zb.idx_items.size(), d_hash, ze.d_slots, d_ws, ze.d_scan,
                        2, d_parsed, slz, fork ? c->ev_hash : nullptr);
    }
    if (fork) FQZ_CUDA_TRY(c, cudaEventRecord(c->ev_join, slz));
    u32 *d_lzflags = nullptr;
    const u32 nent = (u32)zb.idx_ent.size();
    const ZRStream *d_rs = (const ZRStream *)(d_up + rs_at);
    u32 *d_rhash = nullptr, *d_bsizes = nullptr;
    u8 *pool_out = nullptr;
    if (nrs) {
        // duplicated records: pair them up, flag the streams that hold enough of them and code those with
        // sequences as one frame each; the blocks go through a workspace pool in batches (a launch over
        // unflagged streams returns at once)
        StageScope sc(c, ST_ZENC_DUP, ent_bytes);
        const u32 BB = std::min<u32>(zb.rblocks, 8192u);
        u32 *d_keys = (u32 *)c->arena.alloc(offs_words * sizeof(u32));
        u32 *d_cand = (u32 *)c->arena.alloc(offs_words * sizeof(u32));
        d_lzflags = (u32 *)c->arena.alloc((size_t)(4 * nrs + 4) * sizeof(u32));  // flags[nrs + 1], the duplicate counts, the record ranges
        d_rhash = (u32 *)c->arena.alloc((size_t)(nrs + 1) * sizeof(u32));
        d_bsizes = (u32 *)c->arena.alloc((size_t)(zb.rblocks + 1) * sizeof(u32));
        u8 *pool_ws = (u8 *)c->arena.alloc(fqz_lzrec_pool_ws(BB) + 64);
        pool_out = (u8 *)c->arena.alloc(fqz_lzrec_pool_out(zb.rblocks) + 64);
        u32 *d_rparsed = (u32 *)c->arena.alloc((size_t)BB * 2 * sizeof(u32));
        if (!d_keys || !d_cand || !d_lzflags || !d_rhash || !d_bsizes || !pool_ws || !pool_out || !d_rparsed) {
            c->err = "arena: out of device memory (record matcher)";
            return FQZ_E_CUDA;
        }
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_lzflags, 0, (size_t)(4 * nrs + 4) * sizeof(u32), s));
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_cand, 0, offs_words * sizeof(u32), s));  // a record that straddles into a segment is looked up there too
        fqz_launch_rec_match(d_rs, nrs, zb.rmax_records, offs_base, d_keys, d_cand, d_lzflags, d_lzflags + nrs + 1, d_lzflags + 2 * nrs + 2, d_rhash, s);
        for (u32 g0 = 0; g0 < zb.rblocks; g0 += BB)
            fqz_launch_lzrec(d_rs, nrs, d_lzflags, offs_base, d_cand, pool_ws, pool_out, g0, std::min<u32>(zb.rblocks, g0 + BB), d_rparsed, d_bsizes, s);
    }
    {
        StageScope sc(c, ST_ZENC_ENTROPY, ent_bytes);
        u8 *d_zh = (nent && !c->opt_huf_single) ? (u8 *)c->arena.alloc(fqz_zenc_huf_scratch(nent)) : nullptr;
        fqz_launch_zenc_huf(ze.d_frames, d_idx, nent, d_hash, ze.d_slots, ze.d_scan, d_lzflags, d_zh, s);
    }
    if (nrs) {
        StageScope sc(c, ST_ZENC_DUP, 0);
        fqz_launch_lzrec_close(d_rs, nrs, zb.rmax_blocks, d_lzflags, d_rhash, pool_out, d_bsizes, ze.d_frames, ze.d_slots, ze.d_scan, s);
    }
    if (fork) FQZ_CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_join, 0));
    {
        StageScope sc(c, ST_SCAN, 0);
        if (!zb.idx_index.empty()) fqz_launch_zindex(ze.d_frames, nf, ze.d_slots, ze.d_scan, d_lzflags, s);
        FQZ_TRY(fqz_scan_excl_u32(c, ze.d_scan, (u64)nf + 1, (u64)nf + 1, 1));
    }
    return FQZ_OK;
}

// ---------------------------------------------------------------------------------- fqz_zstd_compress (entropy stage alone)
extern "C" int fqz_zstd_compress(fqz_ctx *c, const uint8_t *src, size_t n, int policy, uint8_t *dst, size_t cap, size_t *out_len) {
    if (!c || !out_len || (!src && n)) return FQZ_E_INVALID_ARG;
    if (n > FQZ_MAX_WINDOW) return FQZ_E_TOO_LARGE;
    cudaSetDevice(c->device);
    c->arena.reset();
    c->err.clear();
    *out_len = 0;
    if (n == 0) return FQZ_OK;  // EncodeAll of an empty slice yields zero bytes
    cudaStream_t s = c->stream;
    u8 *d_src = (u8 *)c->arena.alloc(n + 64);
    if (!d_src) return FQZ_E_CUDA;
    FQZ_CUDA_TRY(c, cudaMemcpyAsync(d_src, src, n, cudaMemcpyHostToDevice, s));
    ZBatch zb;
    zb.add_stream(d_src, n, policy == FQZ_ZPOLICY_ENTROPY ? FQZ_ZPOLICY_ENTROPY : FQZ_ZPOLICY_AUTO);
    ZEncoded ze;
    FQZ_TRY(zbatch_encode(c, zb, ze, n));
    u32 nf = ze.nframes;
    u32 *h = (u32 *)c->h_pin;
    FQZ_TRY(fqz_pin_copy(c, h, ze.d_scan + nf, sizeof(u32)));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    size_t total = h[0];
    *out_len = total;
    if (total > cap) return FQZ_E_NOSPACE;
    u8 *d_out = (u8 *)c->arena.alloc(total + 16);
    u32 *d_fixed = (u32 *)c->arena.alloc((size_t)nf * sizeof(u32));
    if (!d_out || !d_fixed) return FQZ_E_CUDA;
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_fixed, 0, (size_t)nf * sizeof(u32), s));
    {
        StageScope sc(c, ST_ASSEMBLE, 2 * total);
        FQZ_LAUNCH(k_zassemble, (nf * 32 + 255) / 256, 256, 0, s, ze.d_frames, nf, ze.d_scan, d_fixed, ze.d_slots, d_out);
    }
    FQZ_CUDA_TRY(c, cudaMemcpyAsync(dst, d_out, total, cudaMemcpyDeviceToHost, s));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    return FQZ_OK;
}

// ---------------------------------------------------------------------------------- compress one window (device -> device)
// Stream policies: packed bases and delta-coded qualities go through the literals-only path
// (Huffman beats LZ+Huffman on them); the four structured streams get LZ77.
static const int kStreamPolicy[6] = {FQZ_ZPOLICY_ENTROPY, FQZ_ZPOLICY_ENTROPY, FQZ_ZPOLICY_ITEMS, FQZ_ZPOLICY_ITEMS, FQZ_ZPOLICY_ITEMS,
                                     FQZ_ZPOLICY_ITEMS};

// Compresses the whole blocks of d_text[0..n) into d_out.  Emits the 10-byte file header first
// when `with_file_header`.  Returns bytes written in *out_len, text consumed in *consumed.
int fqz_compress_window(fqz_ctx *c, const u8 *d_text, u64 n, bool is_last, u64 rec_base, int phred_mode, bool with_file_header,
                        u32 header_block_size, u8 *d_out, size_t out_cap, size_t *out_len, u64 *consumed, u64 *records, u32 *phred_out, u32 skip) {
    cudaStream_t s = c->stream;
    FrontOut fo;
    FQZ_TRY(fqz_run_frontend(c, d_text, n, is_last, rec_base, phred_mode, 0, fo, skip));
    *consumed = fo.consumed;
    *records = fo.R;
    *phred_out = fo.phred64;
    u32 nb = fo.nblocks;
    ZBatch zb;
    std::vector<u32> first((size_t)nb * 6 + 1), nrec(nb);
    u64 stream_bytes = 0;
    for (u32 b = 0; b < nb; b++) {
        nrec[b] = (u32)std::min<u64>(FQZ_BLOCK_RECORDS, fo.R - (u64)b * FQZ_BLOCK_RECORDS);
        for (int a = 0; a < 6; a++) {
            first[(size_t)b * 6 + a] = (u32)zb.frames.size();
            size_t o0 = fo.blk_off[a][b], o1 = fo.blk_off[a][b + 1];
            if (kStreamPolicy[a] == FQZ_ZPOLICY_ITEMS) {
                // headers / plus lines / N positions: item starts = the scanned per-record offsets; lengths: 4-byte items
                // (the block's own slice of the offsets: items[0] is the stream start, so that the first item has no predecessor and
                // a block is coded the same whether or not other blocks share its device window)
                if (a < 5)
                    zb.add_stream(fo.d_streams[a] + o0, o1 - o0, kStreamPolicy[a], fo.d_offs + a * fo.offs_stride + (size_t)b * FQZ_BLOCK_RECORDS, (u32)o0,
                                  nrec[b] + 1);
                else zb.add_stream(fo.d_streams[a] + o0, o1 - o0, kStreamPolicy[a], nullptr, 0, 4);
            } else  // packed bases / qualities: literals-only, with the record boundaries for the duplicate search
                zb.add_stream(fo.d_streams[a] + o0, o1 - o0, kStreamPolicy[a], fo.d_offs + a * fo.offs_stride, (u32)o0, (u32)fo.R + 1,
                              b * FQZ_BLOCK_RECORDS, nrec[b]);
            stream_bytes += o1 - o0;
        }
    }
    first[(size_t)nb * 6] = (u32)zb.frames.size();
    u32 nf = (u32)zb.frames.size();
    u32 hdr0 = with_file_header ? 10u : 0u;
    if (nf == 0) {  // no records: header-only output (compress.go:203-213)
        *out_len = hdr0;
        if (hdr0 > out_cap) return FQZ_E_NOSPACE;
    }
    ZEncoded ze;
    if (nf) FQZ_TRY(zbatch_encode(c, zb, ze, stream_bytes, fo.d_offs, (size_t)2 * fo.offs_stride));
    // fixed offsets (headers in front of each frame) + small tables, one upload
    std::vector<u32> fixed(nf);
    for (u32 b = 0; b < nb; b++)
        for (u32 f = first[(size_t)b * 6]; f < first[(size_t)(b + 1) * 6]; f++) fixed[f] = hdr0 + 36u * (b + 1);
    size_t up_words = (size_t)nf + first.size() + 2 * (size_t)nb;
    // the frame upload in zbatch_encode reads the same pinned buffer: wait before reusing it
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    FQZ_TRY(fqz_pin_reserve(c, 8192 + up_words * 4));
    u32 *hp = (u32 *)(c->h_pin + 4096);
    if (nf) memcpy(hp, fixed.data(), (size_t)nf * 4);
    memcpy(hp + nf, first.data(), first.size() * 4);
    if (nb) memcpy(hp + nf + first.size(), nrec.data(), (size_t)nb * 4);
    if (nb) memcpy(hp + nf + first.size() + nb, fo.orig.data(), (size_t)nb * 4);
    u32 *d_tab = (u32 *)c->arena.alloc(up_words * 4 + 16);
    if (!d_tab) return FQZ_E_CUDA;
    FQZ_TRY(fqz_pin_copy(c, d_tab, hp, up_words * 4));
    u32 *d_fixed = d_tab, *d_first = d_tab + nf, *d_nrec = d_first + first.size(), *d_orig = d_nrec + nb;
    size_t total = hdr0;
    if (nf) {
        u32 *h = (u32 *)c->h_pin;
        FQZ_TRY(fqz_pin_copy(c, h, ze.d_scan + nf, sizeof(u32)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        total = (size_t)hdr0 + 36ull * nb + h[0];
    }
    *out_len = total;
    if (total > out_cap) return FQZ_E_NOSPACE;
    if (with_file_header) {  // container.go:35-45
        u8 *fh = c->h_pin + 512;
        fh[0] = 'F'; fh[1] = 'Q'; fh[2] = 'Z'; fh[3] = 0;
        fh[4] = 2;
        u32 bs = header_block_size ? header_block_size : FQZ_BLOCK_RECORDS;  // compress.go:129-131
        fh[5] = (u8)bs; fh[6] = (u8)(bs >> 8); fh[7] = (u8)(bs >> 16); fh[8] = (u8)(bs >> 24);
        fh[9] = fo.phred64 ? 2 : 0;  // FlagPhred64, compress.go:162-164
        FQZ_TRY(fqz_pin_copy(c, d_out, fh, 10));
    }
    if (nf) {
        StageScope sc(c, ST_ASSEMBLE, 2 * (total - hdr0));
        FQZ_LAUNCH(k_write_block_headers, (nb + 63) / 64, 64, 0, s, nb, d_first, ze.d_scan, d_nrec, d_orig, hdr0, d_out);
        FQZ_LAUNCH(k_zassemble, (nf * 32 + 255) / 256, 256, 0, s, ze.d_frames, nf, ze.d_scan, d_fixed, ze.d_slots, d_out);
    }
    return FQZ_OK;
}

extern "C" size_t fqz_compress_bound(size_t n) { return n + n / 32 + 65536; }

// file-level state carried from window to window (Phred flag of the first block, record count)
struct CompState {
    bool first = true;
    u64 rec_base = 0;
    u32 phred64 = 0;
    int forced_phred = -1;    // >= 0: the flag was decided elsewhere (block 0 lives on another GPU)
    bool emit_header = true;  // false: a shard that is not the head of the file
};

// Compresses d_fastq[0..n) window by window (each window cut at a block boundary; the host only
// walks, all data stays in HBM).  !is_last: only whole 100 000-record blocks are taken and
// *consumed tells the caller where the unconsumed tail starts.
// h_out != nullptr: host pipeline mode — d_fastq is c->io.d_in still being uploaded (every window
// first gates the compute stream on the chunks it reads) and each window's output is downloaded to
// h_out on the copy stream while the next window is coded.
static int compress_device_impl(fqz_ctx *c, const u8 *d_fastq, u64 n, bool is_last, CompState &st, u32 header_block_size, u8 *d_out,
                                size_t out_cap, size_t *out_len, u64 *consumed, u8 *h_out = nullptr, size_t h_cap = 0) {
    // The entropy kernels are latency-bound per frame and every launch ends in a tail of half-empty SMs:
    // large windows amortise it (131 -> 146 GB/s from 1 GiB to 3 GB windows on config 2).  Window offsets
    // are u32: take + take / 4 must stay below 4 GiB.  The host pipeline keeps 1 GiB windows — its first
    // window cannot start before it has been uploaded.
    const u64 MAXWIN = (u64)3 << 30;
    u64 WIN = h_out ? (u64)1 << 30 : 3000000000ull;
    if (h_out && c->opt_host_window_bytes) WIN = c->opt_host_window_bytes;
    if (!h_out && c->opt_window_bytes) WIN = c->opt_window_bytes;
    u64 pos = 0;
    size_t written = 0;
    *out_len = 0;
    *consumed = 0;
    bool no_split = false;  // a halved tail window held no complete block (very long reads): take the tail whole
    while (st.first || pos < n) {
        c->arena.reset();
        u64 left = n - pos;
        u64 take = left;
        bool last = is_last;
        if (left > WIN + (WIN >> 2)) {
            take = WIN;
            last = false;
        } else if (h_out && !no_split && left > ((u64)384 << 20)) {
            // host pipeline: the compute and the download of the LAST window are not hidden behind any
            // upload, so the end of the input is taken in halving windows
            take = left / 2;
            last = false;
        }
        if (h_out) FQZ_TRY(fqz_io_gate(c, pos + take, nullptr));
        // windows must start 16-byte aligned for the vector loads: start at the aligned address below and
        // tell the front end how many bytes to skip (they end the previous window's last line)
        const u8 *wptr = d_fastq + pos;
        u32 skip = (u32)((uintptr_t)wptr & 15u);
        if (skip > pos) {  // only the caller's own pointer can be misaligned at pos == 0: copy (rare)
            u8 *tmp = (u8 *)c->arena.alloc(take + 64);
            if (!tmp) return FQZ_E_CUDA;
            StageScope sc(c, ST_COPY, 2 * take);
            FQZ_CUDA_TRY(c, cudaMemcpyAsync(tmp, wptr, take, cudaMemcpyDeviceToDevice, c->stream));
            FQZ_CUDA_TRY(c, cudaMemsetAsync(tmp + take, 0, 64, c->stream));
            wptr = tmp;
            skip = 0;
        }
        wptr -= skip;
        size_t wl = 0;
        u64 used = 0, recs = 0;
        u32 ph = 0;
        int pmode = st.first ? (st.forced_phred >= 0 ? st.forced_phred : -1) : (int)st.phred64;
        int rc = fqz_compress_window(c, wptr, take + skip, last, st.rec_base, pmode, st.first && st.emit_header, header_block_size,
                                     d_out + written, out_cap - written, &wl, &used, &recs, &ph, skip);
        if (rc == FQZ_OK) used -= skip;
        if (rc == FQZ_E_NEED_MORE && take < left && left <= WIN + (WIN >> 2) && !no_split) {
            no_split = true;
            continue;
        }
        if (rc == FQZ_E_NEED_MORE && take < left && WIN < MAXWIN) {
            // no complete 100 000-record block inside the window (long reads: a block of 10 kb reads is ~2 GB of
            // text): grow the window up to what u32 offsets allow before giving up
            WIN = std::min(MAXWIN, WIN * 2);
            no_split = true;
            continue;
        }
        if (rc == FQZ_E_NEED_MORE) {
            if (take < left) return FQZ_E_TOO_LARGE;  // no complete block inside the largest device window
            break;                                    // streaming: the tail waits for more data
        }
        if (rc != FQZ_OK) {
            if (rc == FQZ_E_NOSPACE) *out_len = written + wl;
            return rc;
        }
        if (st.first) st.phred64 = ph;
        if (h_out) {
            if (written + wl > h_cap) {
                *out_len = written + wl;
                return FQZ_E_NOSPACE;
            }
            FQZ_TRY(fqz_io_download(c, 0, h_out + written, d_out + written, wl));
        }
        written += wl;
        pos += (last && is_last) ? left : used;
        st.rec_base += recs;
        st.first = false;
        *consumed = pos;
        if (take == left) break;
    }
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    *out_len = written;
    return FQZ_OK;
}

extern "C" int fqz_compress_device(fqz_ctx *c, const void *d_fastq, size_t n, uint32_t header_block_size, void *d_out, size_t out_cap,
                                   size_t *out_len) {
    if (!c || !out_len || (!d_fastq && n) || !d_out) return FQZ_E_INVALID_ARG;
    if (((uintptr_t)d_fastq & 15u) != 0) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    CompState st;
    u64 used = 0;
    return compress_device_impl(c, (const u8 *)d_fastq, n, true, st, header_block_size, (u8 *)d_out, out_cap, out_len, &used);
}

// FASTQ text that is already in device memory (inflated there by fqz_compress_gz) -> complete .fqz in HOST memory
int fqz_compress_text_to_host(fqz_ctx *c, const u8 *d_text, u64 n, u32 header_block_size, u8 *out, size_t out_cap, size_t *out_len) {
    const size_t ocap = fqz_compress_bound(n);
    u8 *d_o = nullptr;
    FQZ_TRY(fqz_io_out_acquire(c, 0, ocap, &d_o));
    CompState st;
    u64 used = 0;
    size_t m = 0;
    FQZ_TRY(compress_device_impl(c, d_text, n, true, st, header_block_size, d_o, ocap, &m, &used));
    *out_len = m;
    if (m > out_cap) return FQZ_E_NOSPACE;
    if (m) FQZ_CUDA_TRY(c, cudaMemcpyAsync(out, d_o, m, cudaMemcpyDeviceToHost, c->stream));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return FQZ_OK;
}

// the device-memory twin of fqz_compress_shard: d_fastq may have any alignment (a shard starts where the plan cuts)
extern "C" int fqz_compress_shard_device(fqz_ctx *c, const void *d_fastq, size_t n, uint32_t header_block_size, int phred64, int emit_file_header,
                                         void *d_out, size_t out_cap, size_t *out_len, int *phred64_used) {
    if (!c || !out_len || (!d_fastq && n) || !d_out) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    CompState st;
    st.forced_phred = phred64 < 0 ? -1 : (phred64 ? 1 : 0);
    st.emit_header = emit_file_header != 0;
    u64 used = 0;
    int rc = compress_device_impl(c, (const u8 *)d_fastq, n, true, st, header_block_size, (u8 *)d_out, out_cap, out_len, &used);
    if (phred64_used) *phred64_used = (int)st.phred64;
    return rc;
}

// ---------------------------------------------------------------------------------- streaming (Seam B)
struct fqz_cstream {
    fqz_ctx *c;
    CompState st;
    u32 header_block_size = 0;
};
extern "C" int fqz_compress_begin(fqz_ctx *c, uint32_t header_block_size, fqz_cstream **out) {
    if (!c || !out) return FQZ_E_INVALID_ARG;
    *out = new fqz_cstream();
    (*out)->c = c;
    (*out)->header_block_size = header_block_size;
    return FQZ_OK;
}
extern "C" void fqz_compress_end(fqz_cstream *s) {
    if (!s) return;
    delete s;
}
// One fed window goes through the same copy pipeline as fqz_compress: chunked upload on its own stream, the
// compute stream gated chunk by chunk, every device window's output downloaded while the next one is coded
// (the calls the Go shim uses must overlap too).  Feed large windows (>= 1 GiB) from
// page-locked buffers (fqz_host_alloc): each call pays one pipeline fill and drain.
extern "C" int fqz_compress_feed(fqz_cstream *s, const uint8_t *fastq, size_t n, int is_last, uint8_t *out, size_t out_cap, size_t *out_len,
                                 size_t *consumed) {
    if (!s || !out_len || !consumed || (!fastq && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    fqz_ctx *c = s->c;
    cudaSetDevice(c->device);
    c->err.clear();
    c->arena.reset();
    *out_len = 0;
    *consumed = 0;
    if (n > FQZ_MAX_WINDOW) {  // take a window's worth; the caller re-presents the rest
        n = FQZ_MAX_WINDOW;
        is_last = 0;
    }
    size_t ocap = fqz_compress_bound(n);
    u8 *d_o = nullptr;
    static u8 dummy;
    int rc = fqz_io_upload(c, fastq, n);
    if (rc == FQZ_OK) rc = fqz_io_out_acquire(c, 0, ocap, &d_o);
    CompState trial = s->st;  // committed only when the call succeeds
    size_t m = 0;
    u64 used = 0;
    if (rc == FQZ_OK) rc = compress_device_impl(c, c->io.d_in, n, is_last != 0, trial, s->header_block_size, d_o, ocap, &m, &used, out ? out : &dummy, out_cap);
    int rc2 = fqz_io_finish(c);  // never return while a copy still reads or writes the caller's memory
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_E_NOSPACE) *out_len = m;  // nothing consumed; retry with a larger buffer
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    if (rc != FQZ_OK) return rc;
    if (used == 0 && m == 0 && !is_last) return FQZ_E_NEED_MORE;
    *out_len = m;
    s->st = trial;
    *consumed = (size_t)used;
    return FQZ_OK;
}

extern "C" int fqz_compress(fqz_ctx *c, const uint8_t *fastq, size_t n, uint32_t header_block_size, uint8_t *out, size_t out_cap,
                            size_t *out_len) {
    return fqz_compress_shard(c, fastq, n, header_block_size, -1, 1, out, out_cap, out_len, nullptr);
}

extern "C" int fqz_compress_shard(fqz_ctx *c, const uint8_t *fastq, size_t n, uint32_t header_block_size, int phred64, int emit_file_header,
                                  uint8_t *out, size_t out_cap, size_t *out_len, int *phred64_used) {
    if (!c || !out_len || (!fastq && n) || !out) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    c->arena.reset();
    *out_len = 0;
    size_t ocap = fqz_compress_bound(n);
    u8 *d_o = nullptr;
    int rc = fqz_io_upload(c, fastq, n);
    if (rc == FQZ_OK) rc = fqz_io_out_acquire(c, 0, ocap, &d_o);
    size_t m = 0;
    if (rc == FQZ_OK) {
        CompState st;
        st.forced_phred = phred64 < 0 ? -1 : (phred64 ? 1 : 0);
        st.emit_header = emit_file_header != 0;
        u64 used = 0;
        rc = compress_device_impl(c, c->io.d_in, n, true, st, header_block_size, d_o, ocap, &m, &used, out, out_cap);
        if (phred64_used) *phred64_used = (int)st.phred64;
    }
    int rc2 = fqz_io_finish(c);  // never return while a copy still reads or writes the caller's memory
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_OK || rc == FQZ_E_NOSPACE) *out_len = m;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}
