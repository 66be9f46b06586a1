// fqz_api_front.cu — host orchestration of the compress front end + fqz_encode_streams.
#include <string.h>

#include <algorithm>
#include <string>

#include "fqz_host.h"

// boundaries of every fqz block inside the five scanned offset arrays
__global__ void k_gather_bounds(const u32 *offs, u64 stride, u64 R, u32 nblocks, u32 *out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 per = nblocks + 1;
    if (i >= 5 * per) return;
    u32 a = i / per, b = i % per;
    u64 r = min((u64)b * FQZ_BLOCK_RECORDS, R);
    out[i] = offs[a * stride + r];
}

static const int kErrOfKind[5] = {0, FQZ_E_HEADER_AT, FQZ_E_PLUS, FQZ_E_LEN_MISMATCH, FQZ_E_LONG_N};

// turns the window's error key into the reference's error text (and code)
static int frontend_error(fqz_ctx *c, const FqzWinStatus *hst2, const u8 *d_text, const u32 *d_line_end, u64 rec_base, u64 R, FrontOut &out) {
    cudaStream_t s = c->stream;
        u32 kind = (u32)(hst2->err_key & 0xFF);
        u64 rec = hst2->err_key >> 8;
        char msg[640];
        snprintf(msg, sizeof msg, "record %llu: %s", (unsigned long long)rec, fqz_strerror(kErrOfKind[kind < 5 ? kind : 0]));
        c->err = msg;
        if (kind == FQZ_K_LONG_N && rec >= rec_base && rec - rec_base < R) {
            // the reference names the record and its length (compress.go:484): fetch the header line
            u32 *le = (u32 *)(c->h_pin + 1792);
            u8 *hb = c->h_pin + 3072;
            u64 rr = rec - rec_base;
            if (fqz_pin_copy(c, le, d_line_end + 4 * rr - 1, 3 * sizeof(u32)) == FQZ_OK && cudaStreamSynchronize(s) == cudaSuccess) {
                u32 hs = le[0] + 1u, he = le[1], se = le[2];
                u32 hl = he > hs ? he - hs : 0u, show = hl > 256u ? 256u : hl;
                // lines are cut at '\n'; one trailing '\r' is not part of the field (parser.go:213-215)
                u8 *tmp = c->h_pin + 2048;
                if (fqz_pin_copy(c, hb, d_text + hs, show + 1u) == FQZ_OK && fqz_pin_copy(c, tmp, d_text + (se ? se - 1u : 0u), 1) == FQZ_OK &&
                    cudaStreamSynchronize(s) == cudaSuccess) {
                    u32 L = se - he - 1u;
                    if (L && tmp[0] == '\r') L--;
                    if (show == hl && show && hb[show - 1] == '\r') show--;
                    std::string q;
                    for (u32 i = 1; i < show; i++) {  // skip the '@'
                        u8 ch = hb[i];
                        if (ch == '"' || ch == '\\') { q += '\\'; q += (char)ch; }
                        else if (ch < 0x20 || ch >= 0x7f) { char e[8]; snprintf(e, sizeof e, "\\x%02x", ch); q += e; }
                        else q += (char)ch;
                    }
                    snprintf(msg, sizeof msg, "record \"%s\": sequence length %u has ambiguous bases beyond position %u; N-position tracking is limited to %u bp",
                             q.c_str(), L, FQZ_MAX_SEQ_LEN, FQZ_MAX_SEQ_LEN);
                    c->err = msg;
                }
            }
        }
        out.consumed = rec;  // callers read the offending record index here
        return kErrOfKind[kind < 5 ? kind : 0];
}

int fqz_run_frontend(fqz_ctx *c, const u8 *d_text, u64 n, bool is_last, u64 rec_base, int phred_mode, u64 max_records, FrontOut &out, u32 skip) {
    out = FrontOut();
    cudaStream_t s = c->stream;
    if (n >= ((u64)1 << 32) - 65536) return FQZ_E_TOO_LARGE;
    u32 ntiles = (u32)((n + FQZ_NL_TILE - 1) / FQZ_NL_TILE);
    u32 *h = (u32 *)c->h_pin;
    u64 nlines = 0;
    u32 *d_line_end = nullptr;
    u32 *d_tiles = nullptr;
    bool indexed = false;
    // ---- 1'. single pass: count and index together (look-back over the tiles), line_end[] sized for lines of 16 bytes
    //      or more on average; texts with shorter lines take the two passes below
    if (c->opt_frontend != 0 && ntiles) {  // 1 and 2
        const u64 cap_lines = n / 16 + 4096;
        u32 *d_alloc = (u32 *)c->arena.alloc((size_t)(cap_lines + 8) * sizeof(u32));
        unsigned long long *d_look = (unsigned long long *)c->arena.alloc(((size_t)ntiles + 2) * sizeof(unsigned long long));
        if (!d_alloc || !d_look) {
            c->err = "arena: out of device memory (line_end)";
            return FQZ_E_CUDA;
        }
        u32 *d_total = (u32 *)(d_look + ntiles + 1);
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_look, 0, ((size_t)ntiles + 2) * sizeof(unsigned long long), s));
        u32 *le = d_alloc + 4;
        if (!skip) FQZ_CUDA_TRY(c, cudaMemsetAsync(le - 1, 0xFF, sizeof(u32), s));
        {
            StageScope sc(c, ST_NL_INDEX, n);
            fqz_launch_newline_scan(d_text, n, skip ? skip - 1u : 0u, skip ? le - 1 : le, (u32)cap_lines, d_look, ntiles, d_total, s);
        }
        FQZ_TRY(fqz_pin_copy(c, h, d_total, sizeof(u32)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        if ((u64)h[0] <= cap_lines) {
            nlines = h[0];
            d_line_end = le;
            indexed = true;
        }
    }
    if (!indexed) {
    // ---- 1. count newlines per 16 KiB tile, scan, read the total
    d_tiles = (u32 *)c->arena.alloc(((size_t)ntiles + 1) * sizeof(u32));
    if (!d_tiles) {
        c->err = "arena: out of device memory (tiles)";
        return FQZ_E_CUDA;
    }
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_tiles + ntiles, 0, sizeof(u32), s));
    {
        StageScope sc(c, ST_NL_COUNT, n);
        fqz_launch_newline_count(d_text, n, skip ? skip - 1u : 0u, d_tiles, ntiles, s);
    }
    {
        StageScope sc(c, ST_SCAN, 0);
        FQZ_TRY(fqz_scan_excl_u32(c, d_tiles, (u64)ntiles + 1, (u64)ntiles + 1, 1));
    }
    FQZ_TRY(fqz_pin_copy(c, h, d_tiles + ntiles, sizeof(u32)));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    nlines = h[0];
    }
    // skip > 0: the window starts inside the 16-byte word that still holds the end of the previous
    // window's last line, '\n' included (windows are cut behind a newline): that line is line -1
    if (skip) {
        if (nlines == 0) return FQZ_E_INVALID_ARG;
        nlines -= 1;
    }
    u64 R_all = nlines / 4;
    u64 R = R_all;
    u32 tail_lines = 0;
    if (max_records && R > max_records) R = max_records;
    if (!is_last) {
        R = (R / FQZ_BLOCK_RECORDS) * FQZ_BLOCK_RECORDS;  // whole blocks only; the caller re-feeds the rest
        if (R == 0) return FQZ_E_NEED_MORE;
    } else if (R == R_all) {
        tail_lines = (u32)(nlines - 4 * R_all);
    }
    out.R = R;
    out.nblocks = (u32)((R + FQZ_BLOCK_RECORDS - 1) / FQZ_BLOCK_RECORDS);
    if (!indexed) {
    // ---- 2. line_end[]
    u64 want_lines = 4 * R + tail_lines;
    u32 *d_line_alloc = (u32 *)c->arena.alloc((size_t)(want_lines + 8) * sizeof(u32));
    if (!d_line_alloc) {
        c->err = "arena: out of device memory (line_end)";
        return FQZ_E_CUDA;
    }
    // entry -1 of line_end[]: end of the line in front of record 0 (0xFFFFFFFF + 1 = 0 when there is none)
    d_line_end = d_line_alloc + 4;
    if (!skip) FQZ_CUDA_TRY(c, cudaMemsetAsync(d_line_end - 1, 0xFF, sizeof(u32), s));
    {
        StageScope sc(c, ST_NL_INDEX, n + 4 * want_lines);
        fqz_launch_newline_index(d_text, n, skip ? skip - 1u : 0u, d_tiles, ntiles, skip ? d_line_end - 1 : d_line_end, (u32)want_lines + (skip ? 1u : 0u), s);
    }
    }
    // ---- 3'. fused path: one kernel sizes the records, scans the sizes by look-back and scatters (one pass over the
    //      text instead of two, no size arrays through HBM); the streams are allocated at fixed generous sizes and a
    //      window whose streams outgrow them is redone by the separate kernels below
    if (c->opt_frontend == 2 && R > 0) {
        const size_t arena_mark_unused = 0;
        (void)arena_mark_unused;
        u64 stride = ((R + 1 + 63) / 64) * 64;
        u32 nb5 = 5 * (out.nblocks + 1);
        u32 *d_offs = (u32 *)c->arena.alloc((size_t)(5 * stride) * sizeof(u32));
        u32 *d_bounds = (u32 *)c->arena.alloc((size_t)nb5 * sizeof(u32));
        const u32 ncta = (u32)((R + FQZ_SC_RPC - 1) / FQZ_SC_RPC);
        unsigned long long *d_look = (unsigned long long *)c->arena.alloc(((size_t)ncta * 5 + 2) * sizeof(unsigned long long));
        // what the streams of ordinary FASTQ need is ~0.7 n in all; every stream gets room for far more than its share
        // (seq: a base costs a text byte and a quality byte; qual: same; names, plus lines and N lists: half the text each)
        const u64 room[5] = {n / 8 + R + 64, n / 2 + 64, n / 2 + 2 * R + 64, n / 2 + 2 * R + 64, n / 4 + 2 * R + 64};
        u32 caps[5];
        u8 *d_str[6];
        bool ok = d_offs && d_bounds && d_look;
        for (int a = 0; a < 5 && ok; a++) {
            caps[a] = (u32)std::min<u64>(room[a], 0xFFFFFF00ull);
            d_str[a] = (u8 *)c->arena.alloc((size_t)caps[a] + 16);
            ok = d_str[a] != nullptr;
        }
        if (ok) {
            d_str[5] = (u8 *)c->arena.alloc((size_t)4 * R + 16);
            ok = d_str[5] != nullptr;
        }
        if (!ok) {
            c->err = "arena: out of device memory (fused front end)";
            return FQZ_E_CUDA;
        }
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_look, 0, ((size_t)ncta * 5 + 2) * sizeof(unsigned long long), s));
        u32 *d_ticket = (u32 *)(d_look + (size_t)ncta * 5);
        FqzWinStatus *hst = (FqzWinStatus *)(c->h_pin + 1024);
        hst->err_key = ~0ull;
        hst->qual_min = 255u;
        hst->pad = 0;
        FQZ_TRY(fqz_pin_copy(c, c->d_status, hst, sizeof(FqzWinStatus)));
        if (phred_mode == -1) {
            StageScope sc(c, ST_RECORD_META, 0);
            fqz_launch_phred_min(d_text, d_line_end, std::min<u64>(R, rec_base < FQZ_BLOCK_RECORDS ? FQZ_BLOCK_RECORDS - rec_base : 0), c->d_status, s);
            fqz_launch_decide_phred(c->d_status, c->d_phred, s);
        } else if (phred_mode >= 0) {
            u32 *hv = (u32 *)(c->h_pin + 1280);
            *hv = phred_mode ? 1u : 0u;
            FQZ_TRY(fqz_pin_copy(c, c->d_phred, hv, sizeof(u32)));
        }
        {
            StageScope sc(c, ST_SCATTER, n);
            fqz_launch_scatter_fused(d_text, d_line_end, R, rec_base, tail_lines, d_offs, stride, c->d_phred, d_str, caps, c->d_status, d_look, d_ticket, s);
        }
        FQZ_LAUNCH(k_gather_bounds, (nb5 + 127) / 128, 128, 0, s, d_offs, stride, R, out.nblocks, d_bounds);
        FQZ_TRY(fqz_pin_reserve(c, 4096 + (size_t)nb5 * sizeof(u32)));
        FqzWinStatus *hst2 = (FqzWinStatus *)(c->h_pin + 1536);
        u32 *hphred = (u32 *)(c->h_pin + 1600);
        u32 *hcons = (u32 *)(c->h_pin + 1664);
        u32 *hb = (u32 *)(c->h_pin + 2048);
        FQZ_TRY(fqz_pin_copy(c, hst2, c->d_status, sizeof(FqzWinStatus)));
        FQZ_TRY(fqz_pin_copy(c, hphred, c->d_phred, sizeof(u32)));
        FQZ_TRY(fqz_pin_copy(c, hcons, d_line_end + 4 * R - 1, sizeof(u32)));
        FQZ_TRY(fqz_pin_copy(c, hb, d_bounds, (size_t)nb5 * sizeof(u32)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        if (hst2->err_key != ~0ull) return frontend_error(c, hst2, d_text, d_line_end, rec_base, R, out);
        if (!(hst2->pad & 1u)) {
            out.phred64 = *hphred;
            out.d_offs = d_offs;
            out.offs_stride = stride;
            out.consumed = (u64)(*hcons) + 1;
            if (is_last && R == R_all) out.consumed = n;  // an unterminated / partial last record is dropped
            for (int a = 0; a < 5; a++) out.blk_off[a].assign(hb + a * (out.nblocks + 1), hb + (a + 1) * (out.nblocks + 1));
            out.blk_off[5].resize(out.nblocks + 1);
            out.orig.resize(out.nblocks);
            for (u32 b = 0; b <= out.nblocks; b++) out.blk_off[5][b] = (u32)(4 * std::min<u64>((u64)b * FQZ_BLOCK_RECORDS, R));
            for (u32 b = 0; b < out.nblocks; b++) out.orig[b] = out.blk_off[1][b + 1] - out.blk_off[1][b];
            for (int a = 0; a < 6; a++) out.d_streams[a] = d_str[a];
            c->fused_windows++;
            return FQZ_OK;
        }
        c->legacy_windows++;  // a stream outgrew its room: the separate kernels size everything exactly
    }
    // ---- 3. per-record sizes + validation + Phred min, then five scans
    u64 stride = ((R + 1 + 63) / 64) * 64;
    u32 *d_sizes = (u32 *)c->arena.alloc((size_t)(5 * stride) * sizeof(u32));
    u32 *d_bounds = (u32 *)c->arena.alloc((size_t)(5 * (out.nblocks + 1)) * sizeof(u32));
    if (!d_sizes || !d_bounds) {
        c->err = "arena: out of device memory (sizes)";
        return FQZ_E_CUDA;
    }
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_sizes, 0, (size_t)(5 * stride) * sizeof(u32), s));
    FqzWinStatus *hst = (FqzWinStatus *)(c->h_pin + 1024);
    hst->err_key = ~0ull;
    hst->qual_min = 255u;
    hst->pad = 0;
    FQZ_TRY(fqz_pin_copy(c, c->d_status, hst, sizeof(FqzWinStatus)));
    u64 phred_records = (phred_mode == -1) ? (u64)FQZ_BLOCK_RECORDS : 0;
    {
        StageScope sc(c, ST_RECORD_META, 0);
        fqz_launch_record_meta(d_text, d_line_end, R, rec_base, tail_lines, d_sizes, stride, c->d_status, phred_records, s);
    }
    if (phred_mode == -1) {
        fqz_launch_decide_phred(c->d_status, c->d_phred, s);
    } else if (phred_mode >= 0) {
        u32 *hv = (u32 *)(c->h_pin + 1280);
        *hv = phred_mode ? 1u : 0u;
        FQZ_TRY(fqz_pin_copy(c, c->d_phred, hv, sizeof(u32)));
    }
    {
        StageScope sc(c, ST_SCAN, 0);
        FQZ_TRY(fqz_scan_excl_u32(c, d_sizes, R + 1, stride, 5));
    }
    u32 nb = 5 * (out.nblocks + 1);
    FQZ_LAUNCH(k_gather_bounds, (nb + 127) / 128, 128, 0, s, d_sizes, stride, R, out.nblocks, d_bounds);
    // ---- 4. one small readback: status, Phred flag, block boundaries, consumed offset
    FQZ_TRY(fqz_pin_reserve(c, 4096 + (size_t)nb * sizeof(u32)));
    h = (u32 *)c->h_pin;
    FqzWinStatus *hst2 = (FqzWinStatus *)(c->h_pin + 1536);
    u32 *hphred = (u32 *)(c->h_pin + 1600);
    u32 *hcons = (u32 *)(c->h_pin + 1664);
    u32 *hb = (u32 *)(c->h_pin + 2048);
    FQZ_TRY(fqz_pin_copy(c, hst2, c->d_status, sizeof(FqzWinStatus)));
    FQZ_TRY(fqz_pin_copy(c, hphred, c->d_phred, sizeof(u32)));
    if (R) FQZ_TRY(fqz_pin_copy(c, hcons, d_line_end + 4 * R - 1, sizeof(u32)));
    FQZ_TRY(fqz_pin_copy(c, hb, d_bounds, (size_t)nb * sizeof(u32)));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    if (hst2->err_key != ~0ull) return frontend_error(c, hst2, d_text, d_line_end, rec_base, R, out);
    out.phred64 = *hphred;
    out.d_offs = d_sizes;
    out.offs_stride = stride;
    out.consumed = R ? (u64)(*hcons) + 1 : (is_last ? n : 0);
    if (is_last && R == R_all) out.consumed = n;  // an unterminated / partial last record is dropped
    for (int a = 0; a < 5; a++) out.blk_off[a].assign(hb + a * (out.nblocks + 1), hb + (a + 1) * (out.nblocks + 1));
    out.blk_off[5].resize(out.nblocks + 1);
    out.orig.resize(out.nblocks);
    for (u32 b = 0; b <= out.nblocks; b++) out.blk_off[5][b] = (u32)(4 * std::min<u64>((u64)b * FQZ_BLOCK_RECORDS, R));
    for (u32 b = 0; b < out.nblocks; b++) out.orig[b] = out.blk_off[1][b + 1] - out.blk_off[1][b];
    // ---- 5. carve the six streams and scatter
    u64 total = 0;
    for (int a = 0; a < 6; a++) {
        size_t sz = out.blk_off[a][out.nblocks];
        out.d_streams[a] = (u8 *)c->arena.alloc(sz + 16);
        if (!out.d_streams[a]) {
            c->err = "arena: out of device memory (streams)";
            return FQZ_E_CUDA;
        }
        total += sz;
    }
    {
        StageScope sc(c, ST_SCATTER, out.consumed + total);
        fqz_launch_scatter(d_text, d_line_end, R, d_sizes, stride, c->d_phred, out.d_streams, s);
    }
    return FQZ_OK;
}

extern "C" int fqz_encode_streams(fqz_ctx *c, const uint8_t *fastq, size_t n, int phred64, uint8_t *const outp[6], const size_t cap[6],
                                  size_t len[6], uint64_t info[6]) {
    if (!c || (!fastq && n) || !len || !info) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->arena.reset();
    c->err.clear();
    for (int i = 0; i < 6; i++) info[i] = 0, len[i] = 0;
    u8 *d_text = (u8 *)c->arena.alloc(n + 64);
    if (!d_text) return FQZ_E_CUDA;
    FQZ_CUDA_TRY(c, cudaMemsetAsync(d_text + (n & ~(size_t)15), 0, 64, c->stream));  // defined bytes in the slack
    if (n) FQZ_CUDA_TRY(c, cudaMemcpyAsync(d_text, fastq, n, cudaMemcpyHostToDevice, c->stream));
    FrontOut fo;
    int rc = fqz_run_frontend(c, d_text, n, true, 0, phred64 < 0 ? -1 : (phred64 ? 1 : 0), FQZ_BLOCK_RECORDS, fo, 0);
    if (rc != FQZ_OK) {
        info[5] = fo.consumed;
        return rc;
    }
    info[0] = fo.R;
    info[1] = fo.consumed;
    info[2] = fo.phred64;
    info[3] = fo.nblocks ? fo.orig[0] : 0;
    info[4] = info[3];
    for (int a = 0; a < 6; a++) len[a] = fo.nblocks ? fo.blk_off[a][1] : 0;
    for (int a = 0; a < 6; a++)
        if (len[a] > cap[a]) rc = FQZ_E_NOSPACE;
    if (rc == FQZ_OK)
        for (int a = 0; a < 6; a++)
            if (len[a]) FQZ_CUDA_TRY(c, cudaMemcpyAsync(outp[a], fo.d_streams[a], len[a], cudaMemcpyDeviceToHost, c->stream));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return rc;
}
