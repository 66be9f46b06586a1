// fqz_api_dec.cu — host orchestration of the zstd decode stage (+ fqz_zstd_decompress).
#include <string.h>

#include <algorithm>

#include "fqz_host.h"
#include "fqz_zstd_dec.h"

// Decodes a batch of streams.  On return out.d_base + out.off[i] is stream i (out.size[i] bytes).
int fqz_zdecode_batch(fqz_ctx *c, const std::vector<ZDStream> &streams, u64 max_out, ZDecodeOut &out) {
    cudaStream_t s = c->stream;
    u32 ns = (u32)streams.size();
    out.off.assign(ns, 0);
    out.size.assign(ns, 0);
    out.err_stream = -1;
    out.d_base = nullptr;
    if (ns == 0) return FQZ_OK;
    size_t in_b = (size_t)ns * sizeof(ZDStream), info_b = (size_t)ns * sizeof(ZDStreamInfo), res_b = (size_t)ns * sizeof(ZDStreamResult);
    FQZ_TRY(fqz_pin_reserve(c, 8192 + in_b + info_b + res_b));
    u8 *hp = c->h_pin + 4096;
    memcpy(hp, streams.data(), in_b);
    ZDStream *d_streams = (ZDStream *)c->arena.alloc(in_b);
    ZDStreamInfo *d_info = (ZDStreamInfo *)c->arena.alloc(info_b);
    ZDStreamResult *d_res = (ZDStreamResult *)c->arena.alloc(res_b);
    if (!d_streams || !d_info || !d_res) {
        c->err = "arena: out of device memory (zstd decode)";
        return FQZ_E_CUDA;
    }
    FQZ_TRY(fqz_pin_copy(c, d_streams, hp, in_b));
    u64 cbytes = 0;
    for (auto &st : streams) cbytes += st.csize;
    {
        StageScope sc(c, ST_ZDEC_SCAN, 0);
        fqz_launch_zd_hop(d_streams, ns, d_info, nullptr, nullptr, 0, s);
    }
    ZDStreamInfo *hinfo = (ZDStreamInfo *)(hp + in_b);
    FQZ_TRY(fqz_pin_copy(c, hinfo, d_info, info_b));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    u64 nframes = 0, nblocks = 0, obytes = 0;
    for (u32 i = 0; i < ns; i++) {
        if (hinfo[i].status) {
            out.err_stream = (int)i;
            c->err = hinfo[i].status == 2 ? "zstd: dictionary frames are not supported" : "zstd: corrupt frame or block header";
            return FQZ_E_ZSTD;
        }
        hinfo[i].frame_base = (u32)nframes;
        hinfo[i].block_base = (u32)nblocks;
        hinfo[i].out_base = obytes;
        hinfo[i].lit_base = 0;
        hinfo[i].seq_base = 0;
        out.off[i] = obytes;
        nframes += hinfo[i].nframes;
        nblocks += hinfo[i].nblocks;
        obytes += (hinfo[i].out_bytes + 63) & ~(u64)63;  // each stream starts 64-byte aligned
    }
    if (nframes >= (1ull << 31) || nblocks >= (1ull << 31) || obytes >= (1ull << 32) || (max_out && obytes > max_out)) return FQZ_E_TOO_LARGE;
    FQZ_TRY(fqz_pin_copy(c, d_info, hinfo, info_b));
    ZDFrame *d_frames = (ZDFrame *)c->arena.alloc((size_t)(nframes + 1) * sizeof(ZDFrame));
    ZDBlock *d_blocks = (ZDBlock *)c->arena.alloc((size_t)(nblocks + 1) * sizeof(ZDBlock));
    const u32 cstride = (u32)nblocks + 1;
    u32 *d_cnt = (u32 *)c->arena.alloc((size_t)cstride * 4 * sizeof(u32));  // literals, sequences, has-sequences, leads-a-literal-group per block (+ totals after the scan)
    u32 *d_seqblk = (u32 *)c->arena.alloc((size_t)cstride * sizeof(u32));
    u32 *d_litgrp = (u32 *)c->arena.alloc((size_t)cstride * sizeof(u32));
    u8 *d_out = (u8 *)c->arena.alloc(obytes + 64);
    if (!d_frames || !d_blocks || !d_cnt || !d_seqblk || !d_litgrp || !d_out) {
        c->err = "arena: out of device memory (zstd decode tables)";
        return FQZ_E_CUDA;
    }
    out.d_base = d_out;
    u64 lbytes = 0, nseq = 0;
    u32 nsb = 0;   // blocks with sequences
    u32 ngrp = 0;  // literal decode groups
    {
        StageScope sc(c, ST_ZDEC_SCAN, cbytes);
        FQZ_CUDA_TRY(c, cudaMemsetAsync(d_cnt, 0, (size_t)cstride * 4 * sizeof(u32), s));
        fqz_launch_zd_hop(d_streams, ns, d_info, d_frames, d_blocks, 1, s);
        fqz_launch_zd_parse(d_blocks, (u32)nblocks, d_cnt, cstride, s);
        fqz_launch_zd_link(d_frames, (u32)nframes, d_blocks, d_cnt, cstride, s);
        FQZ_TRY(fqz_scan_excl_u32(c, d_cnt, cstride, cstride, 4));
        fqz_launch_zd_offsets(d_blocks, (u32)nblocks, d_cnt, cstride, d_seqblk, d_litgrp, s);
        u32 *htot = (u32 *)c->h_pin;
        FQZ_TRY(fqz_pin_copy(c, htot, d_cnt + nblocks, sizeof(u32)));
        FQZ_TRY(fqz_pin_copy(c, htot + 1, d_cnt + cstride + nblocks, sizeof(u32)));
        FQZ_TRY(fqz_pin_copy(c, htot + 2, d_cnt + 2 * (size_t)cstride + nblocks, sizeof(u32)));
        FQZ_TRY(fqz_pin_copy(c, htot + 3, d_cnt + 3 * (size_t)cstride + nblocks, sizeof(u32)));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        lbytes = htot[0];  // <= obytes < 2^32 (k_zd_link rejects frames that claim more than their output)
        nseq = htot[1];
        nsb = htot[2];
        ngrp = htot[3];
    }
    u8 *d_lit = (u8 *)c->arena.alloc(lbytes + 64);
    u32 *d_seq = (u32 *)c->arena.alloc((size_t)nseq * 12 + 64);
    u32 *d_tabs = (u32 *)c->arena.alloc((size_t)nsb * FQZ_ZD_TAB_BYTES + 64);
    if (!d_lit || !d_seq || !d_tabs) {
        c->err = "arena: out of device memory (zstd decode arenas)";
        return FQZ_E_CUDA;
    }
    {
        StageScope sc(c, ST_ZDEC_LITERALS, lbytes);
        fqz_launch_zd_literals(d_blocks, (u32)nblocks, d_litgrp, ngrp, d_frames, d_lit, d_out, s);
    }
    {
        StageScope sc(c, ST_ZDEC_SEQUENCES, nseq * 12);
        fqz_launch_zd_sequences(d_blocks, d_frames, d_seqblk, nsb, d_cnt + 2 * (size_t)cstride, d_tabs, d_seq, s);
    }
    {
        StageScope sc(c, ST_ZDEC_EXECUTE, obytes);
        fqz_launch_zd_rawcopy(d_blocks, (u32)nblocks, d_frames, d_out, s);
        fqz_launch_zd_execute(d_frames, (u32)nframes, d_blocks, d_lit, d_seq, d_out, s);
    }
    {
        StageScope sc(c, ST_XXH64, obytes);
        fqz_launch_zd_checksum(d_frames, (u32)nframes, d_out, s);
    }
    fqz_launch_zd_finish(d_frames, d_info, ns, d_res, s);
    fqz_launch_zd_compact(d_frames, d_info, d_res, ns, d_out, s);
    ZDStreamResult *hres = (ZDStreamResult *)(hp + in_b + info_b);
    FQZ_TRY(fqz_pin_copy(c, hres, d_res, res_b));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    for (u32 i = 0; i < ns; i++) {
        if (hres[i].err) {
            out.err_stream = (int)i;
            c->err = hres[i].err == 12 ? "zstd: content checksum mismatch" : "zstd: corrupt block";
            return FQZ_E_ZSTD;
        }
        out.size[i] = hres[i].size;
    }
    return FQZ_OK;
}

extern "C" int fqz_zstd_decompress(fqz_ctx *c, const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    if (!c || !out_len || (!src && n)) return FQZ_E_INVALID_ARG;
    if (n > FQZ_MAX_WINDOW) return FQZ_E_TOO_LARGE;
    cudaSetDevice(c->device);
    c->arena.reset();
    c->err.clear();
    *out_len = 0;
    if (n == 0) return FQZ_OK;  // DecodeAll of zero bytes yields empty output
    cudaStream_t s = c->stream;
    u8 *d_src = (u8 *)c->arena.alloc(n + 64);
    if (!d_src) return FQZ_E_CUDA;
    FQZ_CUDA_TRY(c, cudaMemcpyAsync(d_src, src, n, cudaMemcpyHostToDevice, s));
    std::vector<ZDStream> st(1);
    st[0].src = (u64)(uintptr_t)d_src;
    st[0].csize = n;
    ZDecodeOut zo;
    FQZ_TRY(fqz_zdecode_batch(c, st, 0, zo));
    *out_len = zo.size[0];
    if (zo.size[0] > cap) return FQZ_E_NOSPACE;
    if (zo.size[0]) FQZ_CUDA_TRY(c, cudaMemcpyAsync(dst, zo.d_base + zo.off[0], zo.size[0], cudaMemcpyDeviceToHost, s));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    return FQZ_OK;
}
