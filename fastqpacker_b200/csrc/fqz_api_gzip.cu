// fqz_api_gzip.cu — host orchestration of the gzip input stage: fqz_is_gzip, fqz_gunzip, fqz_gunzip_device,
// fqz_compress_gz.  Reference: cmd/fqpack/main.go:142-174 (wrapInputMaybeGzip, inputHasGzipMagic) in front of
// compress.Compress; error behaviour follows Go's compress/gzip + compress/flate (stdlib).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fqz_host.h"
#include "fqz_inflate.cuh"

static u32 h_mulmod(u32 a, u32 b) {
    u32 p = 0;
    for (int i = 31; i >= 0; i--) {
        if (a & (1u << i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}

static int gz_fail(fqz_ctx *c, u32 status, u64 bit) {
    char msg[160];
    const unsigned long long at = (unsigned long long)(bit >> 3);
    switch (status) {
    case GZ_ST_ERR_HEADER:
        snprintf(msg, sizeof msg, "gzip: invalid header (at offset %llu)", at);
        c->err = msg;
        return FQZ_E_GZ_HEADER;
    case GZ_ST_ERR_CHECKSUM:
        snprintf(msg, sizeof msg, "gzip: invalid checksum (member trailer at offset %llu)", at);
        c->err = msg;
        return FQZ_E_GZ_CHECKSUM;
    case GZ_ST_ERR_CORRUPT:
        snprintf(msg, sizeof msg, "flate: corrupt input before offset %llu", at);
        c->err = msg;
        return FQZ_E_GZ_CORRUPT;
    case GZ_ST_ERR_TRUNC:
        c->err = "unexpected EOF";
        return FQZ_E_GZ_TRUNC;
    default:
        snprintf(msg, sizeof msg, "gzip stage: internal error %u at bit %llu", status, (unsigned long long)bit);
        c->err = msg;
        return FQZ_E_CUDA;
    }
}

static int gz_text_reserve(fqz_ctx *c, size_t need) {
    if (need <= c->gz_text_cap) return FQZ_OK;
    if (c->gz_text) cudaFree(c->gz_text);
    c->gz_text = nullptr;
    c->gz_text_cap = 0;
    size_t want = need + need / 16 + 4096;
    FQZ_CUDA_TRY(c, cudaMalloc((void **)&c->gz_text, want));
    c->gz_text_cap = want;
    return FQZ_OK;
}

// Inflates the gzip file d_gz[0..n) (device memory, 4-byte aligned, 64 readable bytes behind n).
// d_out != nullptr: the text goes there (FQZ_E_NOSPACE with *out_len = bytes needed when out_cap is short);
// d_out == nullptr: it goes to the context's own text buffer, returned in *d_text.  Uses the arena (not reset here).
static int gz_inflate(fqz_ctx *c, const u8 *d_gz, u64 n, u8 *d_out, u64 out_cap, u8 **d_text, u64 *out_len) {
    cudaStream_t s = c->stream;
    *out_len = 0;
    if (n == 0) {  // gzip.NewReader on an empty input: io.EOF
        c->err = "EOF";
        return FQZ_E_GZ_TRUNC;
    }
    u64 ch = c->opt_gz_chunk_bytes;
    if (!ch) {
        // a warp decodes one chunk; restart points only exist at block boundaries (every 20-60 KB of a zlib stream)
        ch = n / ((u64)c->sm_count * 36u);  // 36 warps of k_gz_decode are resident per SM (registers, shared memory): one wave
        ch = (ch + 4095u) & ~(u64)4095u;
        ch = std::min<u64>(std::max<u64>(ch, 32768u), (u64)4 << 20);
    }
    while ((n + ch - 1) / ch > 32768u) ch *= 2;
    if (ch >= ((u64)1 << 31)) return FQZ_E_TOO_LARGE;
    const u32 K = (u32)((n + ch - 1) / ch);

    const size_t chunks_b = (size_t)K * sizeof(GzChunk), list_b = (size_t)K * sizeof(u32);
    FQZ_TRY(fqz_pin_reserve(c, 8192 + chunks_b + list_b));
    GzChunk *h = (GzChunk *)(c->h_pin + 4096);
    u32 *hlist = (u32 *)(c->h_pin + 4096 + chunks_b);
    unsigned long long *herr = (unsigned long long *)c->h_pin;
    GzChunk *d_chunks = (GzChunk *)c->arena.alloc(chunks_b);
    u32 *d_list = (u32 *)c->arena.alloc(list_b);
    unsigned long long *d_err = (unsigned long long *)c->arena.alloc(sizeof(unsigned long long));
    if (!d_chunks || !d_list || !d_err) {
        c->err = "arena: out of device memory (gzip chunks)";
        return FQZ_E_CUDA;
    }
    GzArgs a;
    memset(&a, 0, sizeof a);
    a.src = (const u32 *)d_gz;
    a.n = n;
    a.chunks = d_chunks;
    a.nchunks = K;
    a.chunk_bytes = (u32)ch;
    a.list = d_list;
    a.err = d_err;
    {
        u32 p = 0x40000000u;  // x^1
        for (int k = 0; k < 3; k++) p = h_mulmod(p, p);
        a.pw[0] = p;  // x^8
        for (int k = 1; k < 48; k++) a.pw[k] = h_mulmod(a.pw[k - 1], a.pw[k - 1]);
    }

    // ---- restart points
    memset(h, 0, chunks_b);
    h[0].start_type = GZ_AT_MEMBER;
    FQZ_TRY(fqz_pin_copy(c, d_chunks, h, chunks_b));
    {
        StageScope sc(c, ST_GZ_FIND, n);
        fqz_launch_gz_find(a, s);
    }
    FQZ_TRY(fqz_pin_copy(c, h, d_chunks, chunks_b));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    std::vector<u32> chain;
    chain.push_back(0);
    for (u32 k = 1; k < K; k++)
        if (h[k].start_type != GZ_AT_NONE) chain.push_back(k);
    auto retarget = [&](size_t i) {
        GzChunk &C = h[chain[i]];
        if (i + 1 < chain.size()) {
            C.target_bit = h[chain[i + 1]].start_bit;
            C.target_type = h[chain[i + 1]].start_type;
        } else {
            C.target_bit = GZ_NO_TARGET;
            C.target_type = GZ_AT_NONE;
        }
    };
    for (size_t i = 0; i < chain.size(); i++) retarget(i);

    // ---- speculative pass: every chunk decodes into scratch room of GZ_SPEC_RATIO symbols per compressed byte (FASTQ
    //      inflates 3-5x) and a few member records; what it finds out — sizes, member ends, whether it lands exactly on
    //      the next restart point — is all the counting pass of a two-pass design would give, and the symbols come with it.
    //      A chunk that runs out of room keeps counting and is decoded again below, at its exact size.  Without the
    //      memory for the scratch the pass only counts (room 0) and every chunk is decoded again.
    const u64 GZ_SPEC_RATIO = 6;
    u64 spec_cap = GZ_SPEC_RATIO * ch;
    const u32 spec_mcap = (u32)(ch / 512u) + 8u;
    u16 *d_spec = (u16 *)c->arena.alloc((size_t)K * spec_cap * sizeof(u16));
    GzMember *d_smem = (GzMember *)c->arena.alloc((size_t)K * spec_mcap * sizeof(GzMember));
    if (!d_smem) {
        c->err = "arena: out of device memory (gzip members)";
        return FQZ_E_CUDA;
    }
    if (!d_spec) spec_cap = 0;
    for (u32 k = 0; k < K; k++) {
        h[k].sym_ptr = (u64)(uintptr_t)(d_spec + (size_t)k * spec_cap);
        h[k].sym_cap = spec_cap;
        h[k].mem_ptr = (u64)(uintptr_t)(d_smem + (size_t)k * spec_mcap);
        h[k].mem_cap = spec_mcap;
    }
    a.spec = 1;
    std::vector<u32> dirty = chain;
    size_t proven = 0;
    int reruns = 0;
    for (;;) {
        memcpy(hlist, dirty.data(), dirty.size() * sizeof(u32));
        FQZ_TRY(fqz_pin_copy(c, d_chunks, h, chunks_b));
        FQZ_TRY(fqz_pin_copy(c, d_list, hlist, dirty.size() * sizeof(u32)));
        a.nlist = (u32)dirty.size();
        {
            StageScope sc(c, ST_GZ_DECODE, n);
            fqz_launch_gz_decode(a, true, s);
        }
        FQZ_TRY(fqz_pin_copy(c, h, d_chunks, chunks_b));
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
        bool again = false;
        while (proven < chain.size()) {
            GzChunk &C = h[chain[proven]];
            if (C.status == GZ_ST_REACHED) {
                proven++;
                continue;
            }
            if (C.status == GZ_ST_END) {  // nothing behind the end of the input can be a restart point
                chain.resize(proven + 1);
                proven++;
                break;
            }
            if (C.status == GZ_ST_OVERSHOOT && proven + 1 < chain.size()) {
                // nobody lands on the next restart point: it was a false positive, this chunk decodes through it
                if (++reruns > 8)
                    chain.resize(proven + 1);  // a stream full of look-alikes: the rest goes to one warp
                else
                    chain.erase(chain.begin() + (long)proven + 1);
                retarget(proven);
                dirty.assign(1, chain[proven]);
                again = true;
                break;
            }
            if (C.status >= 16u) return gz_fail(c, C.status, C.err_bit);
            return gz_fail(c, GZ_ST_ERR_INTERNAL, C.start_bit);
        }
        if (!again) break;
    }

    // ---- places
    u64 total = 0, over_sym = 0;
    u32 nm = 0, over_mem = 0, nblock = 0;
    std::vector<u32> redo;
    for (size_t i = 0; i < chain.size(); i++) {
        GzChunk &C = h[chain[i]];
        C.out_off = total;
        C.member_base = nm;
        total += C.out_len;
        nm += C.members;
        if (C.start_type == GZ_AT_BLOCK) nblock++;
        if (C.overflow) {
            redo.push_back(chain[i]);
            over_sym += C.out_len;
            over_mem += C.members;
        }
        if (i == 0)
            C.window_valid = 0;
        else {
            const GzChunk &P = h[chain[i - 1]];
            u64 v = P.after_member != ~0ull ? P.after_member : (u64)P.window_valid + P.out_len;
            C.window_valid = (u32)std::min<u64>(v, GZ_WINDOW);
        }
    }
    c->gz_stats[0] = K;
    c->gz_stats[1] = chain.size();
    c->gz_stats[2] = (u64)reruns;
    c->gz_stats[3] = nm;
    *out_len = total;
    u8 *out = d_out;
    if (d_out) {
        if (total > out_cap) return FQZ_E_NOSPACE;
    } else {
        FQZ_TRY(gz_text_reserve(c, (size_t)total + 256));
        out = c->gz_text;
    }
    if (d_text) *d_text = out;
    // windows: groups of ~sqrt(chunks) chunks (fqz_inflate.cu, k_gz_maps)
    u32 gsize = 1;
    while ((u64)gsize * gsize < chain.size()) gsize++;
    const u32 ngroups = (u32)((chain.size() + gsize - 1) / gsize);
    u8 *d_win = (u8 *)c->arena.alloc((size_t)ngroups * GZ_WINDOW);
    u16 *d_maps = nblock ? (u16 *)c->arena.alloc((size_t)chain.size() * GZ_WINDOW * sizeof(u16)) : nullptr;
    GzMember *d_mem = (GzMember *)c->arena.alloc(((size_t)nm + 1) * sizeof(GzMember));
    u16 *d_osym = redo.empty() ? nullptr : (u16 *)c->arena.alloc((size_t)over_sym * sizeof(u16) + 64);
    GzMember *d_omem = redo.empty() ? nullptr : (GzMember *)c->arena.alloc(((size_t)over_mem + 1) * sizeof(GzMember));
    if (!d_win || !d_mem || (nblock && !d_maps) || (!redo.empty() && (!d_osym || !d_omem))) {
        c->err = "arena: out of device memory (gzip output)";
        return FQZ_E_CUDA;
    }
    {
        u64 so = 0;
        u32 mo = 0;
        for (u32 k : redo) {  // exact room for the chunks that ran out of it
            h[k].sym_ptr = (u64)(uintptr_t)(d_osym + so);
            h[k].sym_cap = h[k].out_len;
            h[k].mem_ptr = (u64)(uintptr_t)(d_omem + mo);
            h[k].mem_cap = h[k].members;
            so += h[k].out_len;
            mo += h[k].members;
        }
    }
    a.win = d_win;
    a.maps = d_maps;
    a.gsize = gsize;
    a.out = out;
    a.members = d_mem;
    a.nmembers = nm;
    *herr = ~0ull;
    FQZ_TRY(fqz_pin_copy(c, d_chunks, h, chunks_b));
    FQZ_TRY(fqz_pin_copy(c, d_err, herr, sizeof(unsigned long long)));
    c->gz_stats[4] = redo.size();
    if (!redo.empty()) {
        memcpy(hlist, redo.data(), redo.size() * sizeof(u32));
        FQZ_TRY(fqz_pin_copy(c, d_list, hlist, redo.size() * sizeof(u32)));
        a.nlist = (u32)redo.size();
        a.spec = 0;
        StageScope sc(c, ST_GZ_DECODE, n + 2 * over_sym);
        fqz_launch_gz_decode(a, true, s);
        // hlist is read by the copy kernel when it runs: it must have run before the chain is written over it
        FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    }
    memcpy(hlist, chain.data(), chain.size() * sizeof(u32));
    FQZ_TRY(fqz_pin_copy(c, d_list, hlist, chain.size() * sizeof(u32)));
    a.nlist = (u32)chain.size();
    {
        StageScope sc(c, ST_GZ_RESOLVE, 3 * total);
        fqz_launch_gz_members(a, s);
        fqz_launch_gz_windows(a, s);
        fqz_launch_gz_resolve(a, s);
    }
    {
        StageScope sc(c, ST_GZ_CRC, total);
        fqz_launch_gz_crc(a, total, s);
    }
    if (!d_out) FQZ_CUDA_TRY(c, cudaMemsetAsync(out + total, 0, 64, s));
    FQZ_TRY(fqz_pin_copy(c, herr, d_err, sizeof(unsigned long long)));
    FQZ_CUDA_TRY(c, cudaStreamSynchronize(s));
    FQZ_CUDA_TRY(c, cudaGetLastError());
    if (*herr != ~0ull) return gz_fail(c, (u32)(*herr & 0xffu), (u64)(*herr >> 8));
    return FQZ_OK;
}

extern "C" int fqz_gunzip_stats(fqz_ctx *c, uint64_t out[5]) {
    if (!c || !out) return FQZ_E_INVALID_ARG;
    for (int i = 0; i < 5; i++) out[i] = c->gz_stats[i];
    return FQZ_OK;
}

extern "C" int fqz_is_gzip(const uint8_t *buf, size_t n) { return buf && n >= 2 && buf[0] == 0x1f && buf[1] == 0x8b; }

extern "C" int fqz_gunzip_device(fqz_ctx *c, const void *d_gz, size_t n, void *d_out, size_t out_cap, size_t *out_len) {
    if (!c || !out_len || (!d_gz && n) || (!d_out && out_cap)) return FQZ_E_INVALID_ARG;
    if (((uintptr_t)d_gz & 15u) || ((uintptr_t)d_out & 15u)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    c->arena.reset();
    u64 len = 0;
    static u8 none;
    int rc = gz_inflate(c, (const u8 *)d_gz, n, d_out ? (u8 *)d_out : &none, out_cap, nullptr, &len);
    *out_len = (size_t)len;
    return rc;
}

// the compressed bytes go up through the copy pipeline's input buffer
static int gz_upload(fqz_ctx *c, const uint8_t *gz, size_t n) {
    FQZ_TRY(fqz_io_upload(c, gz, n));
    return fqz_io_gate(c, n, nullptr);
}

extern "C" int fqz_gunzip(fqz_ctx *c, const uint8_t *gz, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!c || !out_len || (!gz && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    c->arena.reset();
    *out_len = 0;
    u8 *d_text = nullptr;
    u64 len = 0;
    int rc = gz_upload(c, gz, n);
    if (rc == FQZ_OK) rc = gz_inflate(c, c->io.d_in, n, nullptr, 0, &d_text, &len);
    if (rc == FQZ_OK) {
        *out_len = (size_t)len;
        if (len > out_cap)
            rc = FQZ_E_NOSPACE;
        else if (len && cudaMemcpyAsync(out, d_text, len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            rc = FQZ_E_CUDA;
    }
    int rc2 = fqz_io_finish(c);  // never return while a copy still reads or writes the caller's memory
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}

extern "C" int fqz_compress_gz(fqz_ctx *c, const uint8_t *gz, size_t n, uint32_t header_block_size, uint8_t *out, size_t out_cap,
                               size_t *out_len, size_t *fastq_len) {
    if (!c || !out_len || (!gz && n) || (!out && out_cap)) return FQZ_E_INVALID_ARG;
    cudaSetDevice(c->device);
    c->err.clear();
    c->arena.reset();
    *out_len = 0;
    if (fastq_len) *fastq_len = 0;
    u8 *d_text = nullptr;
    u64 len = 0;
    int rc = gz_upload(c, gz, n);
    if (rc == FQZ_OK) rc = gz_inflate(c, c->io.d_in, n, nullptr, 0, &d_text, &len);
    if (rc == FQZ_OK) {
        if (fastq_len) *fastq_len = (size_t)len;
        rc = fqz_compress_text_to_host(c, d_text, len, header_block_size, out, out_cap, out_len);
    }
    int rc2 = fqz_io_finish(c);
    if (rc == FQZ_OK) rc = rc2;
    if (rc == FQZ_E_CUDA && c->err.empty()) c->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}
