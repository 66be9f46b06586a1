// fqz_api_decompress.cu — host orchestration of the decompress path: container walk, zstd decode of
// the six (v1: five) streams of every block, back end, and the decompress entry points.
// Replaces compress.Decompress / decompressJobToPooledBuffer / blockReader.writeRecord
// (internal/compress/compress.go:558-604, 780-837, 944-1078).
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "fqz_backend.h"
#include "fqz_host.h"
#include "fqz_zstd.h"
#include "fqz_zstd_dec.h"

static const int kErrOfBk[8] = {0,
                                FQZ_E_TRUNC_LEN,
                                FQZ_E_TRUNC_NPOS,
                                FQZ_E_TRUNC_HEADER,
                                FQZ_E_TRUNC_SEQ,
                                FQZ_E_NPOS_RANGE,
                                FQZ_E_TRUNC_PLUS,
                                FQZ_E_TRUNC_QUAL};
// compress.go:787-813
static const char *kStreamWhat[6] = {"sequences", "quality", "headers", "plus-line payload", "N positions", "lengths"};

#define DEC_TABLE_CAP 1024u                  // fqz blocks per container-walk pass
#define DEC_WINDOW_BYTES ((u64)1280 << 20)   // compressed bytes per device window (device-resident calls; large windows amortise kernel tails,
                                             // and a reference-written file — one serial frame per stream — lasts as long as its windows are many)
#define DEC_WINDOW_HOST ((u64)256 << 20)     // ... of the host pipeline, whose first window cannot start before it has been uploaded
// decoded stream bytes per device window: FASTQ is at most 1.6 x the stream bytes (2 L text bytes per 1.25 L
// of packed bases + qualities), so FASTQ offsets stay below 2^32
#define DEC_WINDOW_OUT ((u64)2400 << 20)
// device-resident calls: the entropy stage takes as many blocks as its 32-bit arena offsets allow; the back end then
// runs over runs of blocks whose FASTQ stays below 2^32 bytes
#define DEC_WINDOW_OUT_DEV ((u64)3900 << 20)
#ifndef DEC_BACKEND_FASTQ
#define DEC_BACKEND_FASTQ ((u64)3900 << 20)
#endif

static int backend_error(fqz_ctx *c, const std::vector<BkBlock> &blks, u64 key, u64 block_base) {
    u32 kind = (u32)(key & 0xFF);
    u64 rec = key >> 8;
    size_t b = 0;
    while (b + 1 < blks.size() && blks[b + 1].rec_base <= rec) b++;
    int code = kErrOfBk[kind < 8 ? kind : 0];
    char msg[200];
    snprintf(msg, sizeof msg, "decompressing block %llu: %s (record %llu of the block)", (unsigned long long)(block_base + b), fqz_strerror(code),
             (unsigned long long)(rec - blks[b].rec_base));
    c->err = msg;
    return code;
}

// Back end over blocks whose six decoded streams are resident in HBM.  d_out == nullptr: the
// output is allocated from the arena and returned in *d_res.
// io_slot >= 0: the output goes to that staging slot of the copy pipeline instead of the arena.
static int backend_run(fqz_ctx *c, std::vector<BkBlock> &blks, u32 phred64, u8 *d_out, size_t out_cap, u8 **d_res, size_t *out_len,
                       u64 block_base, int io_slot = -1) {
    cudaStream_t s = c->stream;
    u32 nb = (u32)blks.size();
    u64 R = 0;
    u32 max_nrec = 0;
    for (auto &B : blks) {
        // a header claiming more records than the lengths stream holds fails at the first missing
        // length ("truncated length data", compress.go:1047): one record past the end is enough
        u64 have = (u64)B.size[5] / 4;
        if (B.nrec > have) B.nrec = (u32)(have + 1);
        B.rec_base = R;
        R += B.nrec;
        max_nrec = std::max(max_nrec, B.nrec);
    }
    *out_len = 0;
    if (d_res) *d_res = nullptr;
    if (R == 0) return FQZ_OK;
    if (R >= 0xFFFF0000ull || nb > 65535u) return FQZ_E_TOO_LARGE;
    size_t blk_b = (size_t)nb * sizeof(BkBlock), tot_b = (size_t)nb * sizeof(BkTotals);
    FQZ_TRY(fqz_pin_reserve(c, 8192 + blk_b + tot_b));
    u8 *hp = c->h_pin + 4096;
    memcpy(hp, blks.data(), blk_b);
    FqzDecStatus *hst = (FqzDecStatus *)c->h_pin;
    hst->err_key = ~0ull;
    u64 stride = ((R + 1 + 63) / 64) * 64;
    BkBlock *d_blks = (BkBlock *)c->arena.alloc(blk_b);
    BkTotals *d_tot = (BkTotals *)c->arena.alloc(tot_b);
    FqzDecStatus *d_st = (FqzDecStatus *)c->arena.alloc(64);
    u32 *d_offs = (u32 *)c->arena.alloc((size_t)(3 * (R + nb)) * sizeof(u32));
    u32 *d_ok = (u32 *)c->arena.alloc((size_t)nb * 3 * sizeof(u32));
    u32 max_segments = 0;  // frames of the longest hinted chain (the writer cuts item streams every FQZ_ZFRAME_ITEMS bytes)
    for (auto &B : blks)
        for (int k = 0; k < 3; k++)
            if (B.hint[k] && B.hint_size[k] >= 28) max_segments = std::max(max_segments, (B.size[2 + k] + FQZ_ZFRAME_ITEMS - 1) / FQZ_ZFRAME_ITEMS);
    u32 *d_sz = (u32 *)c->arena.alloc((size_t)(3 * stride) * sizeof(u32));
    if (!d_blks || !d_tot || !d_st || !d_offs || !d_sz || !d_ok) {
        c->err = "arena: out of device memory (back end tables)";
        return FQZ_E_CUDA;
    }
    FQZ_TRY(fqz_pin_copy(c, d_blks, hp, blk_b));
    FQZ_TRY(fqz_pin_copy(c, d_st, hst, sizeof(FqzDecStatus)));
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_tot, 0, tot_b, s));
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_ok, 0, (size_t)nb * 3 * sizeof(u32), s));
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_sz, 0, (size_t)(3 * stride) * sizeof(u32), s));
    u64 hdr_bytes = 0;
    for (auto &B : blks) hdr_bytes += (u64)B.size[2] + B.size[3] + B.size[4];
    {
        StageScope sc(c, ST_WALK, hdr_bytes);
        fqz_launch_walk_prefixes(d_blks, nb, max_segments, d_offs, d_ok, d_st, s);
    }
    {
        StageScope sc(c, ST_OFFSETS, 0);
        fqz_launch_record_sizes(d_blks, nb, max_nrec, d_offs, d_sz, stride, d_tot, d_st, s);
    }
    BkTotals *htot = (BkTotals *)(hp + blk_b);
    FqzDecStatus *hst2 = (FqzDecStatus *)(c->h_pin + 64);
    FQZ_TRY(fqz_pin_copy(c, hst2, d_st, sizeof(FqzDecStatus)));
    FQZ_TRY(fqz_pin_copy(c, htot, d_tot, tot_b));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    // Streams that cannot hold what the lengths add up to: the emit pass would find the record, but its 32-bit offsets are
    // only good when the sums are (a damaged length stream may add up to anything)
    bool short_any = false;
    for (u32 b = 0; b < nb; b++) short_any |= htot[b].packed > blks[b].size[0] || htot[b].bases > blks[b].size[1];
    if (hst2->err_key != ~0ull || short_any) {
        // a sequence / quality stream may run short, or an N position may lie beyond its read, at an earlier point of the
        // reference's record order than the failure seen so far
        fqz_launch_first_error(d_blks, nb, d_offs, hst2->err_key, d_st, s);
        FQZ_TRY(fqz_pin_copy(c, hst2, d_st, sizeof(FqzDecStatus)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        if (hst2->err_key != ~0ull) return backend_error(c, blks, hst2->err_key, block_base);
    }
    u64 total = 0, stream_bytes = 0;
    for (u32 b = 0; b < nb; b++) {
        total += htot[b].fastq;
        for (int a = 0; a < 6; a++) stream_bytes += blks[b].size[a];
    }
    if (total >= 0xFFFFFF00ull) return FQZ_E_TOO_LARGE;
    *out_len = (size_t)total;
    if (d_out) {
        if (total > out_cap) return FQZ_E_NOSPACE;
    } else {
        if (io_slot >= 0) FQZ_TRY(fqz_io_out_acquire(c, io_slot, (size_t)total + 64, &d_out));
        else d_out = (u8 *)c->arena.alloc((size_t)total + 64);
        if (!d_out) {
            c->err = "arena: out of device memory (FASTQ output)";
            return FQZ_E_CUDA;
        }
        *d_res = d_out;
    }
    {
        StageScope sc(c, ST_SCAN, 0);
        FQZ_TRY(fqz_scan_excl_u32(c, d_sz, R + 1, stride, 3));
    }
    {
        StageScope sc(c, ST_EMIT, stream_bytes + total);
        fqz_launch_emit(d_blks, nb, max_nrec, d_offs, d_sz, stride, phred64, d_out, d_st, s);
    }
    FQZ_TRY(fqz_pin_copy(c, hst2, d_st, sizeof(FqzDecStatus)));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    if (hst2->err_key != ~0ull) return backend_error(c, blks, hst2->err_key, block_base);
    return FQZ_OK;
}

// Entropy stage + back end for `nb` blocks of a container resident at d_fqz.
static int decode_blocks(fqz_ctx *c, const u8 *d_fqz, const FqzBlockEntry *ent, u32 nb, u32 phred64, u8 *d_out, size_t out_cap, u8 **d_res,
                         size_t *out_len, u64 block_base, int io_slot = -1) {
    std::vector<ZDStream> zs((size_t)nb * 6);
    for (u32 b = 0; b < nb; b++) {
        u64 off = ent[b].payload_off;
        for (int a = 0; a < 6; a++) {
            zs[(size_t)b * 6 + a].src = (u64)(uintptr_t)(d_fqz + off);
            zs[(size_t)b * 6 + a].csize = ent[b].size[a];
            off += ent[b].size[a];
        }
    }
    ZDecodeOut zo;
    const bool chunked = d_out != nullptr && io_slot < 0;  // device output: the back end may run in several passes
    int rc = fqz_zdecode_batch(c, zs, chunked ? DEC_WINDOW_OUT_DEV : DEC_WINDOW_OUT, zo);
    if (rc == FQZ_E_ZSTD && zo.err_stream >= 0) {
        char msg[240];
        snprintf(msg, sizeof msg, "decompressing block %llu: decompressing %s: %s", (unsigned long long)(block_base + zo.err_stream / 6),
                 kStreamWhat[zo.err_stream % 6], c->err.c_str());
        c->err = msg;
    }
    if (rc != FQZ_OK) return rc;
    std::vector<BkBlock> blks(nb);
    for (u32 b = 0; b < nb; b++) {
        BkBlock &B = blks[b];
        for (int a = 0; a < 6; a++) {
            size_t i = (size_t)b * 6 + a;
            if (zo.size[i] >= (1ull << 32)) return FQZ_E_TOO_LARGE;
            B.stream[a] = (u64)(uintptr_t)(zo.d_base + zo.off[i]);
            B.size[a] = (u32)zo.size[i];
        }
        B.nrec = ent[b].nrec;
        B.pad = 0;
        B.rec_base = 0;
        for (int k = 0; k < 3; k++) {  // compressed headers / plus / N-position streams: their index frames carry item hints
            B.hint[k] = zs[(size_t)b * 6 + 2 + k].src;
            B.hint_size[k] = (u32)zs[(size_t)b * 6 + 2 + k].csize;
        }
        B.pad2 = 0;
    }
    if (!chunked) return backend_run(c, blks, phred64, d_out, out_cap, d_res, out_len, block_base, io_slot);
    size_t written = 0;
    *out_len = 0;
    for (u32 b0 = 0; b0 < nb;) {
        u64 est = 0;
        u32 b1 = b0;
        while (b1 < nb) {  // FASTQ bytes of a block <= 2 L + names + plus payloads + 4 per record
            u64 e = 2ull * blks[b1].size[1] + blks[b1].size[2] + blks[b1].size[3] + 4ull * blks[b1].nrec;
            if (b1 > b0 && est + e > DEC_BACKEND_FASTQ) break;
            est += e;
            b1++;
        }
        std::vector<BkBlock> part(blks.begin() + b0, blks.begin() + b1);
        size_t wl = 0;
        rc = backend_run(c, part, phred64, d_out + written, out_cap - written, nullptr, &wl, block_base + b0, -1);
        if (rc == FQZ_E_NOSPACE) {  // report what the whole batch needs as far as it is known
            *out_len = written + wl;
            return rc;
        }
        if (rc != FQZ_OK) return rc;
        written += wl;
        b0 = b1;
    }
    *out_len = written;
    return FQZ_OK;
}

struct DecState {
    bool have_header = false;
    u32 version = 2, phred64 = 0;
    u64 block_base = 0;  // blocks decoded so far (error texts)
};

// Parses the 10-byte file header (container.go:48-67; version check compress.go:571-573).
static int parse_file_header(fqz_ctx *c, const u8 *h, u64 n, DecState &st);
static int read_file_header(fqz_ctx *c, const u8 *d_fqz, u64 n, DecState &st) {
    u8 *h = c->h_pin + 128;
    size_t take = (size_t)std::min<u64>(n, 10);
    FQZ_TRY(fqz_pin_copy(c, h, d_fqz, take));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return parse_file_header(c, h, n, st);
}
static int parse_file_header(fqz_ctx *c, const u8 *h, u64 n, DecState &st) {
    if (n < 4) return FQZ_E_TRUNC_FILE;
    if (!(h[0] == 'F' && h[1] == 'Q' && h[2] == 'Z' && h[3] == 0)) return FQZ_E_MAGIC;
    if (n < 10) return FQZ_E_TRUNC_FILE;
    st.version = h[4];
    if (st.version != 1 && st.version != 2) {
        char msg[64];
        snprintf(msg, sizeof msg, "unsupported file version: %u", st.version);
        c->err = msg;
        return FQZ_E_VERSION;
    }
    st.phred64 = (h[9] & 2) ? 1u : 0u;  // FlagPhred64, compress.go:576-579
    st.have_header = true;
    return FQZ_OK;
}

// Decodes the complete blocks of d_fqz[pos0, n).  d_out != nullptr: FASTQ goes to that device
// buffer; else each window is staged in the arena and copied to h_out.  *consumed = offset after
// the last block decoded.  A partial trailing block is an error only when is_last.
// host_io: d_fqz is c->io.d_in still being uploaded; windows gate on the chunks they read and
// every window's FASTQ is downloaded from an alternating staging slot while the next one is decoded.
// discard: the FASTQ of every window is produced in device memory and dropped (fqz_check); out_cap is ignored.
static int decompress_blocks(fqz_ctx *c, const u8 *d_fqz, u64 n, u64 pos0, bool is_last, bool partial_ok, DecState &st, u8 *d_out, u8 *h_out,
                             size_t out_cap, size_t *out_len, u64 *consumed, bool host_io = false, bool discard = false) {
    cudaStream_t s = c->stream;
    u64 pos = pos0;
    size_t written = 0;
    *out_len = 0;
    *consumed = pos0;
    std::vector<FqzBlockEntry> ent;
    u64 n_all = n;
    u64 want = 0;
    int win = 0;
    while (pos < n_all) {
        // host pipeline: nothing is downloaded before the first window has been uploaded and decoded, so the
        // pipeline is started with a small one
        const u64 win_bytes = host_io ? (win == 0 ? DEC_WINDOW_HOST / 4 : DEC_WINDOW_HOST) : DEC_WINDOW_BYTES;
        c->arena.reset();
        if (host_io) {  // the walk must not look at bytes that have not arrived yet
            size_t avail = 0;
            want = std::max<u64>(want, pos + win_bytes + ((u64)32 << 20));
            FQZ_TRY(fqz_io_gate(c, (size_t)std::min<u64>(n_all, want), &avail));
            n = avail;
        }
        FqzBlockEntry *d_tab = (FqzBlockEntry *)c->arena.alloc(DEC_TABLE_CAP * sizeof(FqzBlockEntry));
        FqzWalkResult *d_wr = (FqzWalkResult *)c->arena.alloc(sizeof(FqzWalkResult));
        if (!d_tab || !d_wr) return FQZ_E_CUDA;
        fqz_launch_walk_container(d_fqz, n, pos, st.version, d_tab, DEC_TABLE_CAP, win_bytes, d_wr, s);
        FQZ_TRY(fqz_pin_reserve(c, 8192 + DEC_TABLE_CAP * sizeof(FqzBlockEntry)));
        FqzWalkResult *hwr = (FqzWalkResult *)(c->h_pin + 256);
        FqzBlockEntry *htab = (FqzBlockEntry *)(c->h_pin + 4096);
        FQZ_TRY(fqz_pin_copy(c, hwr, d_wr, sizeof(FqzWalkResult)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        FqzWalkResult wr = *hwr;
        if (wr.nblocks) {
            FQZ_TRY(fqz_pin_copy(c, htab, d_tab, (size_t)wr.nblocks * sizeof(FqzBlockEntry)));
            FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
            ent.assign(htab, htab + wr.nblocks);
        } else
            ent.clear();
        u32 done = 0;
        u32 take = wr.nblocks;
        while (done < wr.nblocks) {
            take = std::min(take, wr.nblocks - done);
            c->arena.reset();  // the table lives on in `ent`
            size_t wl = 0;
            u8 *d_res = nullptr;
            int slot = (host_io && !discard) ? (win & 1) : -1;
            if (discard) out_cap = written + ((size_t)1 << 40);
            int rc = decode_blocks(c, d_fqz, ent.data() + done, take, st.phred64, d_out ? d_out + written : nullptr, out_cap - written,
                                   &d_res, &wl, st.block_base, slot);
            if (rc == FQZ_E_TOO_LARGE && take > 1) {  // the window decodes to more than one device pass holds
                take = (take + 1) / 2;
                continue;
            }
            if (rc == FQZ_OK && !d_out && wl > out_cap - written) rc = FQZ_E_NOSPACE;
            if (rc == FQZ_E_NOSPACE) {
                if (written && partial_ok) {  // streaming callers keep the progress made so far
                    *out_len = written;
                    return FQZ_OK;
                }
                *out_len = written + wl;
                return FQZ_E_NOSPACE;
            }
            if (rc != FQZ_OK) return rc;
            if (!d_out && wl && !discard) {
                if (host_io) {
                    FQZ_TRY(fqz_io_download(c, slot, h_out + written, d_res, wl));
                    win++;
                } else {
                    StageScope sc(c, ST_COPY, wl);
                    FQZ_CUDA_TRY(c, cudaMemcpyAsync(h_out + written, d_res, wl, cudaMemcpyDeviceToHost, s));
                    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
                }
            }
            written += wl;
            done += take;
            st.block_base += take;
            const FqzBlockEntry &L = ent[done - 1];
            u64 end = L.payload_off;
            for (int a = 0; a < 6; a++) end += L.size[a];
            pos = end;
            *consumed = pos;
            *out_len = written;
        }
        if (wr.status && n < n_all) {  // ran into bytes still in flight, not a truncated file: widen the gate
            if (wr.nblocks == 0) want = std::max<u64>(want, n + win_bytes);
            continue;
        }
        if (wr.status) {  // truncated block header / payload after the blocks just decoded
            if (is_last) {
                c->err = "reading block: unexpected EOF";
                return FQZ_E_TRUNC_FILE;
            }
            break;  // streaming: the caller re-presents the tail with more data
        }
        if (wr.nblocks == 0) break;
    }
    *out_len = written;
    return FQZ_OK;
}

extern "C" int fqz_decompress_device(fqz_ctx *c, const void *d_fqz, size_t n, void *d_out, size_t out_cap, size_t *out_len) {
    if (!c || !out_len || (!d_fqz && n) || (!d_out && out_cap)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    *out_len = 0;
    DecState st;
    FQZ_TRY(read_file_header(c, (const u8 *)d_fqz, n, st));
    u64 used = 0;
    static u8 dummy;
    int rc = decompress_blocks(c, (const u8 *)d_fqz, n, 10, true, false, st, d_out ? (u8 *)d_out : &dummy, nullptr, out_cap, out_len, &used);
    return rc;
}

extern "C" int fqz_decompress(fqz_ctx *c, const uint8_t *fqz, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!c || !out_len || (!fqz && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    *out_len = 0;
    DecState st;
    FQZ_TRY(parse_file_header(c, fqz, n, st));  // the header is read on the host; the payload streams to the device
    int rc = fqz_io_upload(c, fqz, n);
    size_t m = 0;
    if (rc == FQZ_OK) {
        u64 used = 0;
        static u8 dummy;
        rc = decompress_blocks(c, c->io.d_in, n, 10, true, false, st, nullptr, out ? out : &dummy, out_cap, &m, &used, true);
    }
    int rc2 = fqz_io_finish(c);  // never return while a copy still reads or writes the caller's memory
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_OK || rc == FQZ_E_NOSPACE) *out_len = m;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}

// ---------------------------------------------------------------------------------- info / check (ROADMAP.md PR-008, PR-009)
// fqz_info: what `fqpack info` would print — version, flags, block count, record count, stream sizes — from one hop
// over the block headers (container.go:48-152; the format has no index, compress.go:721-736).  Header arithmetic
// on the caller's host buffer: no device work.
extern "C" int fqz_info(fqz_ctx *c, const uint8_t *fqz, size_t n, fqz_file_info *out) {
    if (!c || !out || (!fqz && n)) return FQZ_E_INVALID_ARG;
    c->err.clear();
    memset(out, 0, sizeof *out);
    DecState st;
    FQZ_TRY(parse_file_header(c, fqz, n, st));
    out->version = st.version;
    out->flags = fqz[9];
    out->header_block_size = (u32)fqz[5] | ((u32)fqz[6] << 8) | ((u32)fqz[7] << 16) | ((u32)fqz[8] << 24);
    const size_t hsz = st.version == 1 ? 32 : 36;
    const int ns = st.version == 1 ? 5 : 6;
    size_t pos = 10;
    while (pos < n) {
        if (n - pos < hsz) {
            c->err = "reading block header: unexpected EOF";
            return FQZ_E_TRUNC_FILE;
        }
        u32 v[9];
        for (size_t i = 0; i < hsz / 4; i++) v[i] = (u32)fqz[pos + 4 * i] | ((u32)fqz[pos + 4 * i + 1] << 8) | ((u32)fqz[pos + 4 * i + 2] << 16) | ((u32)fqz[pos + 4 * i + 3] << 24);
        u64 payload = 0;
        // v1: seq, qual, headers, N positions, lengths (container.go:128-152); v2 adds the plus-line payload in front of N positions
        static const int k_v1_slot[5] = {0, 1, 2, 4, 5};
        for (int a = 0; a < ns; a++) {
            payload += v[1 + a];
            out->compressed[st.version == 1 ? k_v1_slot[a] : a] += v[1 + a];
        }
        if (n - pos - hsz < payload) {
            c->err = "reading compressed data: unexpected EOF";
            return FQZ_E_TRUNC_FILE;
        }
        out->blocks++;
        out->records += v[0];
        out->original_seq += v[1 + ns];
        out->original_qual += v[2 + ns];
        pos += hsz + payload;
    }
    return FQZ_OK;
}

// fqz_block_index: the index the container does not carry (compress.go:721-736 reads header after header), built as a
// side table from one hop over the block headers.  Host arithmetic, no context: usable without a GPU.
extern "C" int fqz_block_index(const uint8_t *fqz, size_t n, fqz_block_ref *out, size_t cap, size_t *count) {
    if (!count || (!fqz && n) || (!out && cap)) return FQZ_E_INVALID_ARG;
    *count = 0;
    if (n < 4) return FQZ_E_TRUNC_FILE;
    if (!(fqz[0] == 'F' && fqz[1] == 'Q' && fqz[2] == 'Z' && fqz[3] == 0)) return FQZ_E_MAGIC;
    if (n < 10) return FQZ_E_TRUNC_FILE;
    const u32 version = fqz[4];
    if (version != 1 && version != 2) return FQZ_E_VERSION;
    const size_t hsz = version == 1 ? 32 : 36;
    const int ns = version == 1 ? 5 : 6;
    size_t pos = 10, nb = 0;
    u64 first = 0;
    while (pos < n) {
        if (n - pos < hsz) return FQZ_E_TRUNC_FILE;
        u32 v[9];
        for (size_t i = 0; i < hsz / 4; i++) v[i] = (u32)fqz[pos + 4 * i] | ((u32)fqz[pos + 4 * i + 1] << 8) | ((u32)fqz[pos + 4 * i + 2] << 16) | ((u32)fqz[pos + 4 * i + 3] << 24);
        u64 payload = 0;
        for (int a = 0; a < ns; a++) payload += v[1 + a];
        if (n - pos - hsz < payload) return FQZ_E_TRUNC_FILE;
        if (nb < cap) {
            fqz_block_ref &r = out[nb];
            r.offset = pos;
            r.size = hsz + payload;
            r.first_record = first;
            r.records = v[0];
            r.reserved = 0;
            r.original_seq = v[1 + ns];
            r.original_qual = v[2 + ns];
        }
        nb++;
        *count = nb;
        first += v[0];
        pos += hsz + payload;
    }
    return nb > cap ? FQZ_E_NOSPACE : FQZ_OK;
}

// fqz_decompress_blocks: random access through the index — the bytes of blocks [first_block, first_block + num_blocks)
// are uploaded and decoded like a file that starts there (decompress_blocks with pos0 = 0); version and Phred flag come
// from the file header.
extern "C" int fqz_decompress_blocks(fqz_ctx *c, const uint8_t *fqz, size_t n, uint64_t first_block, uint64_t num_blocks, uint8_t *out,
                                     size_t out_cap, size_t *out_len) {
    if (!c || !out_len || (!fqz && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    *out_len = 0;
    DecState st;
    FQZ_TRY(parse_file_header(c, fqz, n, st));
    size_t nb = 0;
    int rc = fqz_block_index(fqz, n, nullptr, 0, &nb);  // a cut file: nb = the whole blocks in front of the cut, which still decode
    const bool cut = rc == FQZ_E_TRUNC_FILE;
    if (rc != FQZ_OK && rc != FQZ_E_NOSPACE && !cut) return rc;
    if (first_block > nb || num_blocks > nb - first_block) {
        c->err = cut ? "reading block: unexpected EOF" : "block range reaches past the last block of the file";
        return cut ? FQZ_E_TRUNC_FILE : FQZ_E_INVALID_ARG;
    }
    if (num_blocks == 0) return FQZ_OK;
    std::vector<fqz_block_ref> idx(nb);
    rc = fqz_block_index(fqz, n, idx.data(), nb, &nb);
    if (rc != FQZ_OK && rc != FQZ_E_TRUNC_FILE) return rc;
    const u64 lo = idx[first_block].offset;
    const u64 hi = idx[first_block + num_blocks - 1].offset + idx[first_block + num_blocks - 1].size;
    st.block_base = first_block;  // error texts name the block's index in the file
    rc = fqz_io_upload(c, fqz + lo, (size_t)(hi - lo));
    size_t m = 0;
    if (rc == FQZ_OK) {
        u64 used = 0;
        static u8 dummy;
        rc = decompress_blocks(c, c->io.d_in, hi - lo, 0, true, false, st, nullptr, out ? out : &dummy, out_cap, &m, &used, true);
    }
    int rc2 = fqz_io_finish(c);
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_OK || rc == FQZ_E_NOSPACE) *out_len = m;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}

// fqz_check: what `fqpack check` would do — walk the container, decode every stream of every block (frame checksums
// verified), rebuild every record, and report pass / fail without writing the FASTQ anywhere: the text of each
// device window is produced in HBM and dropped, so the check runs at the device-resident decompress rate plus
// the upload of the compressed file.  Returns the first error the decoder meets (same codes and texts as
// fqz_decompress); *records / *fastq_bytes (optional) = what the file holds.
extern "C" int fqz_check(fqz_ctx *c, const uint8_t *fqz, size_t n, uint64_t *records, uint64_t *fastq_bytes) {
    if (!c || (!fqz && n)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    if (records) *records = 0;
    if (fastq_bytes) *fastq_bytes = 0;
    fqz_file_info fi;
    // a cut file is reported by the decode below, which meets the damage in file order like fqz_decompress does (a corrupt
    // frame in an earlier block comes first)
    const int rc_info = fqz_info(c, fqz, n, &fi);
    if (rc_info != FQZ_OK && rc_info != FQZ_E_TRUNC_FILE) return rc_info;
    const std::string info_err = c->err;
    DecState st;
    FQZ_TRY(parse_file_header(c, fqz, n, st));
    int rc = fqz_io_upload(c, fqz, n);
    size_t m = 0;
    if (rc == FQZ_OK) {
        u64 used = 0;
        rc = decompress_blocks(c, c->io.d_in, n, 10, true, false, st, nullptr, nullptr, 0, &m, &used, true, true);
    }
    int rc2 = fqz_io_finish(c);
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    if (rc != FQZ_OK) return rc;
    if (rc_info != FQZ_OK) {
        c->err = info_err;
        return rc_info;
    }
    if (records) *records = fi.records;
    if (fastq_bytes) *fastq_bytes = m;
    return FQZ_OK;
}

// ---------------------------------------------------------------------------------- streaming (Seam B)
struct fqz_dstream {
    fqz_ctx *c;
    DecState st;
};
extern "C" int fqz_decompress_begin(fqz_ctx *c, fqz_dstream **out) {
    if (!c || !out) return FQZ_E_INVALID_ARG;
    *out = new fqz_dstream();
    (*out)->c = c;
    return FQZ_OK;
}
extern "C" void fqz_decompress_end(fqz_dstream *s) {
    if (!s) return;
    delete s;
}
// Like fqz_compress_feed, a fed window runs through the copy pipeline of the whole-buffer call.
extern "C" int fqz_decompress_feed(fqz_dstream *s, const uint8_t *fqz, size_t n, int is_last, uint8_t *out, size_t out_cap, size_t *out_len,
                                   size_t *consumed) {
    if (!s || !out_len || !consumed || (!fqz && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    fqz_ctx *c = s->c;
    cudaSetDevice(c->device);
    c->err.clear();
    *out_len = 0;
    *consumed = 0;
    if (n > FQZ_MAX_WINDOW) {  // take a window's worth; the caller re-presents the rest
        n = FQZ_MAX_WINDOW;
        is_last = 0;
    }
    u64 pos = 0;
    if (!s->st.have_header && n < 10 && !is_last) return FQZ_E_NEED_MORE;
    DecState trial = s->st;  // committed only when the call succeeds: FQZ_E_NOSPACE consumes nothing, not even the file header
    if (!trial.have_header) {
        FQZ_TRY(parse_file_header(c, fqz, n, trial));  // read on the host; the payload streams to the device
        pos = 10;
    }
    u64 used = pos;
    static u8 dummy;
    int rc = fqz_io_upload(c, fqz, n);
    if (rc == FQZ_OK) rc = decompress_blocks(c, c->io.d_in, n, pos, is_last != 0, true, trial, nullptr, out ? out : &dummy, out_cap, out_len, &used, true);
    int rc2 = fqz_io_finish(c);  // never return while a copy still reads or writes the caller's memory
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    if (rc == FQZ_OK) {
        s->st = trial;
        *consumed = (size_t)used;
    }
    if (rc == FQZ_OK && !is_last && used == 0 && n && *out_len == 0) return FQZ_E_NEED_MORE;
    return rc;
}

// ---------------------------------------------------------------------------------- block level
extern "C" int fqz_decode_streams(fqz_ctx *c, const uint8_t *const in[6], const size_t len[6], uint32_t num_records, int phred64, uint8_t *out,
                                  size_t out_cap, size_t *out_len) {
    if (!c || !in || !len || !out_len) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->arena.reset();
    c->err.clear();
    *out_len = 0;
    cudaStream_t s = c->stream;
    std::vector<BkBlock> blks(1);
    BkBlock &B = blks[0];
    for (int a = 0; a < 6; a++) {
        if (len[a] >= (1ull << 31)) return FQZ_E_TOO_LARGE;
        u8 *d = (u8 *)c->arena.alloc(len[a] + 64);
        if (!d) return FQZ_E_CUDA;
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d + (len[a] & ~(size_t)15), 0, 64, s));
        if (len[a]) FQZ_CUDA_TRY(c, cudaMemcpyAsync(d, in[a], len[a], cudaMemcpyHostToDevice, s));
        B.stream[a] = (u64)(uintptr_t)d;
        B.size[a] = (u32)len[a];
    }
    B.nrec = num_records;
    B.pad = 0;
    B.rec_base = 0;
    for (int k = 0; k < 3; k++) B.hint[k] = 0, B.hint_size[k] = 0;
    B.pad2 = 0;
    u8 *d_res = nullptr;
    size_t total = 0;
    FQZ_TRY(backend_run(c, blks, phred64 ? 1u : 0u, nullptr, 0, &d_res, &total, 0));
    *out_len = total;
    if (total > out_cap) return FQZ_E_NOSPACE;
    if (total) FQZ_CUDA_TRY(c, cudaMemcpyAsync(out, d_res, total, cudaMemcpyDeviceToHost, s));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    return FQZ_OK;
}
