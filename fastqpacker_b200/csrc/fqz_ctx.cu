// fqz_ctx.cu — context, device arena, stage profiler, error strings.
#include <stdio.h>
#include <string.h>

#include "fqz_host.h"

thread_local unsigned long long g_fqz_launches = 0;

// ---------------------------------------------------------------------------------- arena
void *Arena::alloc(size_t bytes) {
    size_t need = ((bytes + FQZ_PAD + 255) / 256) * 256;
    for (auto &ch : chunks) {
        if (ch.used + need <= ch.cap) {
            void *p = ch.p + ch.used;
            ch.used += need;
            return p;
        }
    }
    size_t cap = need > min_chunk ? need : min_chunk;
    void *p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    chunks.push_back(Chunk{(u8 *)p, cap, need});
    return p;
}
void Arena::reset() {
    if (chunks.size() > 1) {  // coalesce so that the next window finds one large chunk
        size_t total = 0;
        for (auto &ch : chunks) total += ch.cap;
        release();
        void *p = nullptr;
        size_t cap = total + total / 4;
        if (cudaMalloc(&p, cap) == cudaSuccess) chunks.push_back(Chunk{(u8 *)p, cap, 0});
        else cudaGetLastError();
    }
    for (auto &ch : chunks) ch.used = 0;
}
void Arena::release() {
    for (auto &ch : chunks) cudaFree(ch.p);
    chunks.clear();
}
size_t Arena::capacity() const {
    size_t t = 0;
    for (auto &ch : chunks) t += ch.cap;
    return t;
}

// ---------------------------------------------------------------------------------- profiler
static const char *k_stage_names[ST_COUNT_] = {
    "newline_count", "newline_index", "scan", "record_meta", "scatter_streams", "zstd_enc_entropy", "zstd_enc_lz", "zstd_enc_dup", "xxh64",
    "assemble", "zstd_dec_scan", "zstd_dec_literals", "zstd_dec_sequences", "zstd_dec_execute", "prefix_walk", "record_offsets",
    "emit_fastq", "copy", "gz_find", "gz_decode", "gz_resolve", "gz_crc"};

cudaEvent_t Profiler::get() {
    if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void Profiler::collect() {
    for (auto &r : open) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.stage] += t;
        pool.push_back(r.a);
        pool.push_back(r.b);
    }
    open.clear();
}
void Profiler::clear() {
    collect();
    for (int i = 0; i < ST_COUNT_; i++) {
        ms[i] = 0;
        launches[i] = 0;
        bytes[i] = 0;
    }
}
StageScope::StageScope(fqz_ctx *ctx, int st, u64 bytes) : c(ctx), stage(st), l0(g_fqz_launches) {
    c->prof.bytes[st] += bytes;
    if (c->prof.on) {
        a = c->prof.get();
        cudaEventRecord(a, c->stream);
    }
}
StageScope::~StageScope() {
    c->prof.launches[stage] += g_fqz_launches - l0;
    if (a) {
        cudaEvent_t b = c->prof.get();
        cudaEventRecord(b, c->stream);
        c->prof.open.push_back(Profiler::Rec{stage, a, b});
    }
}

// ---------------------------------------------------------------------------------- context
int fqz_pin_reserve(fqz_ctx *c, size_t bytes) {
    if (bytes <= c->h_pin_cap) return FQZ_OK;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr;
    c->h_pin_cap = 0;
    size_t cap = bytes < 65536 ? 65536 : bytes * 2;
    FQZ_CUDA_TRY(c, cudaMallocHost((void **)&c->h_pin, cap));
    c->h_pin_cap = cap;
    return FQZ_OK;
}
__global__ void __launch_bounds__(256) k_pin_copy(u8 *dst, const u8 *src, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        size_t nv = n >> 4;
        for (size_t k = i; k < nv; k += step) ((uint4 *)dst)[k] = ((const uint4 *)src)[k];
        for (size_t k = (nv << 4) + i; k < n; k += step) dst[k] = src[k];
    } else if ((((uintptr_t)dst | (uintptr_t)src) & 3u) == 0) {
        size_t nw = n >> 2;
        for (size_t k = i; k < nw; k += step) ((u32 *)dst)[k] = ((const u32 *)src)[k];
        for (size_t k = (nw << 2) + i; k < n; k += step) dst[k] = src[k];
    } else
        for (size_t k = i; k < n; k += step) dst[k] = src[k];
}
int fqz_pin_copy(fqz_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!bytes) return FQZ_OK;
    size_t units = (bytes + 15) / 16;
    u32 grid = (u32)((units + 255) / 256);
    if (grid > 128) grid = 128;
    FQZ_LAUNCH(k_pin_copy, grid, 256, 0, c->stream, (u8 *)dst, (const u8 *)src, bytes);
    return FQZ_OK;
}
int fqz_io_reserve(fqz_ctx *c, size_t bytes) {
    if (bytes <= c->h_io_cap) return FQZ_OK;
    if (c->h_io) cudaFreeHost(c->h_io);
    c->h_io = nullptr;
    c->h_io_cap = 0;
    size_t cap = bytes + bytes / 8 + 4096;
    FQZ_CUDA_TRY(c, cudaMallocHost((void **)&c->h_io, cap));
    c->h_io_cap = cap;
    return FQZ_OK;
}

// ---------------------------------------------------------------------------------- copy pipeline
static int io_init(fqz_ctx *c) {
    IoPipe &io = c->io;
    if (io.s_h2d) return FQZ_OK;
    FQZ_CUDA_TRY(c, cudaStreamCreateWithFlags(&io.s_h2d, cudaStreamNonBlocking));
    FQZ_CUDA_TRY(c, cudaStreamCreateWithFlags(&io.s_d2h, cudaStreamNonBlocking));
    FQZ_CUDA_TRY(c, cudaEventCreateWithFlags(&io.ev_done, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) FQZ_CUDA_TRY(c, cudaEventCreateWithFlags(&io.ev_out[k], cudaEventDisableTiming));
    return FQZ_OK;
}
static int io_dev_reserve(fqz_ctx *c, u8 **p, size_t *cap, size_t need) {
    if (need <= *cap) return FQZ_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 8 + 4096;
    FQZ_CUDA_TRY(c, cudaMalloc((void **)p, want));
    *cap = want;
    return FQZ_OK;
}
int fqz_io_upload(fqz_ctx *c, const u8 *host, size_t n) {
    FQZ_TRY(io_init(c));
    IoPipe &io = c->io;
    FQZ_TRY(io_dev_reserve(c, &io.d_in, &io.in_cap, n + 256));
    size_t nchunks = (n + FQZ_IO_CHUNK - 1) / FQZ_IO_CHUNK;
    while (io.ev_chunk.size() < nchunks) {
        cudaEvent_t e = nullptr;
        FQZ_CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        io.ev_chunk.push_back(e);
    }
    io.n = n;
    io.gated = 0;
    // the slack after the input must hold defined bytes (vector loads look a few bytes past the end)
    FQZ_CUDA_TRY(c, cudaMemsetAsync(io.d_in + (n & ~(size_t)15), 0, 64, io.s_h2d));
    for (size_t k = 0; k < nchunks; k++) {
        size_t off = k * FQZ_IO_CHUNK, len = n - off < FQZ_IO_CHUNK ? n - off : FQZ_IO_CHUNK;
        FQZ_CUDA_TRY(c, cudaMemcpyAsync(io.d_in + off, host + off, len, cudaMemcpyHostToDevice, io.s_h2d));
        FQZ_CUDA_TRY(c, cudaEventRecord(io.ev_chunk[k], io.s_h2d));
    }
    if (nchunks == 0) {  // empty input: still order the memset before the compute stream
        if (io.ev_chunk.empty()) {
            cudaEvent_t e = nullptr;
            FQZ_CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            io.ev_chunk.push_back(e);
        }
        FQZ_CUDA_TRY(c, cudaEventRecord(io.ev_chunk[0], io.s_h2d));
        FQZ_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, io.ev_chunk[0], 0));
    }
    return FQZ_OK;
}
int fqz_io_gate(fqz_ctx *c, size_t upto, size_t *avail) {
    IoPipe &io = c->io;
    if (upto > io.n) upto = io.n;
    if (upto > io.gated) {
        size_t k = (upto - 1) / FQZ_IO_CHUNK;  // chunks complete in order on one stream: waiting for the last one is enough
        FQZ_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, io.ev_chunk[k], 0));
        size_t end = (k + 1) * FQZ_IO_CHUNK;
        io.gated = end < io.n ? end : io.n;
    }
    if (avail) *avail = io.gated;
    return FQZ_OK;
}
int fqz_io_out_acquire(fqz_ctx *c, int slot, size_t bytes, u8 **p) {
    FQZ_TRY(io_init(c));
    IoPipe &io = c->io;
    if (io.out_busy[slot]) {  // the previous download from this slot must have drained
        FQZ_CUDA_TRY(c, cudaEventSynchronize(io.ev_out[slot]));
        io.out_busy[slot] = false;
    }
    FQZ_TRY(io_dev_reserve(c, &io.d_out[slot], &io.out_cap[slot], bytes + 256));
    *p = io.d_out[slot];
    return FQZ_OK;
}
int fqz_io_download(fqz_ctx *c, int slot, u8 *host_dst, const u8 *dev_src, size_t bytes) {
    IoPipe &io = c->io;
    if (!bytes) return FQZ_OK;
    FQZ_CUDA_TRY(c, cudaEventRecord(io.ev_done, c->stream));
    FQZ_CUDA_TRY(c, cudaStreamWaitEvent(io.s_d2h, io.ev_done, 0));
    FQZ_CUDA_TRY(c, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, io.s_d2h));
    FQZ_CUDA_TRY(c, cudaEventRecord(io.ev_out[slot], io.s_d2h));
    io.out_busy[slot] = true;
    return FQZ_OK;
}
int fqz_io_finish(fqz_ctx *c) {
    IoPipe &io = c->io;
    if (!io.s_h2d) return FQZ_OK;
    cudaError_t e1 = cudaStreamSynchronize(io.s_h2d), e2 = cudaStreamSynchronize(io.s_d2h), e3 = cudaStreamSynchronize(c->stream);
    io.out_busy[0] = io.out_busy[1] = false;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        c->err = std::string("copy pipeline: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
        return FQZ_E_CUDA;
    }
    return FQZ_OK;
}
void fqz_io_release(fqz_ctx *c) {
    IoPipe &io = c->io;
    if (io.s_h2d) cudaStreamSynchronize(io.s_h2d);
    if (io.s_d2h) cudaStreamSynchronize(io.s_d2h);
    if (io.d_in) cudaFree(io.d_in);
    for (int k = 0; k < 2; k++) {
        if (io.d_out[k]) cudaFree(io.d_out[k]);
        if (io.ev_out[k]) cudaEventDestroy(io.ev_out[k]);
    }
    if (io.ev_done) cudaEventDestroy(io.ev_done);
    for (auto e : io.ev_chunk) cudaEventDestroy(e);
    if (io.s_h2d) cudaStreamDestroy(io.s_h2d);
    if (io.s_d2h) cudaStreamDestroy(io.s_d2h);
    io = IoPipe();
}

extern "C" int fqz_abi_version(void) { return FQZ_ABI_VERSION; }

extern "C" int fqz_init(int device, fqz_ctx **out) {
    if (!out) return FQZ_E_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return FQZ_E_NO_DEVICE;  // no CPU fallback by design
    }
    if (device < 0 || device >= ndev) return FQZ_E_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return FQZ_E_NO_DEVICE;
    fqz_ctx *c = new fqz_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc((void **)&c->d_status, 256) != cudaSuccess ||
        cudaMalloc((void **)&c->d_phred, 256) != cudaSuccess) {
        delete c;
        return FQZ_E_CUDA;
    }
    cudaMemset(c->d_phred, 0, 256);
    if (cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_hash, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return FQZ_E_CUDA;
    }
    // dynamic shared-memory limits are per device: set them for every context (one process may drive several GPUs)
    if (fqz_frontend_init_device() | fqz_zstd_enc_init_device() | fqz_zstd_dec_init_device()) {
        cudaGetLastError();
        delete c;
        return FQZ_E_CUDA;
    }
    if (fqz_pin_reserve(c, 65536) != FQZ_OK) {
        delete c;
        return FQZ_E_CUDA;
    }
    *out = c;
    return FQZ_OK;
}

extern "C" void fqz_destroy(fqz_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->prof.collect();
    for (auto e : c->prof.pool) cudaEventDestroy(e);
    fqz_io_release(c);
    c->arena.release();
    if (c->d_status) cudaFree(c->d_status);
    if (c->d_phred) cudaFree(c->d_phred);
    if (c->gz_text) cudaFree(c->gz_text);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_io) cudaFreeHost(c->h_io);
    if (c->stream_aux) cudaStreamDestroy(c->stream_aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_hash) cudaEventDestroy(c->ev_hash);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char *fqz_last_error(const fqz_ctx *c) { return c ? c->err.c_str() : ""; }

extern "C" const char *fqz_strerror(int code) {
    switch (code) {
    case FQZ_OK: return "ok";
    case FQZ_E_HEADER_AT: return "invalid FASTQ: header line must start with @";
    case FQZ_E_PLUS: return "invalid FASTQ: separator line must start with +";
    case FQZ_E_LEN_MISMATCH: return "invalid FASTQ: sequence and quality lengths must match";
    case FQZ_E_LONG_N: return "sequence has ambiguous bases beyond position 65536; N-position tracking is limited to 65536 bp";
    case FQZ_E_MAGIC: return "invalid magic bytes: not an FQZ file";
    case FQZ_E_VERSION: return "unsupported file version";
    case FQZ_E_TRUNC_FILE: return "unexpected EOF";
    case FQZ_E_ZSTD: return "zstd: invalid or corrupt frame";
    case FQZ_E_TRUNC_HEADER: return "truncated header data";
    case FQZ_E_TRUNC_PLUS: return "truncated plus-line payload data";
    case FQZ_E_TRUNC_SEQ: return "truncated sequence data";
    case FQZ_E_TRUNC_QUAL: return "truncated quality data";
    case FQZ_E_TRUNC_LEN: return "truncated length data";
    case FQZ_E_TRUNC_NPOS: return "truncated N position data";
    case FQZ_E_NOSPACE: return "output buffer too small";
    case FQZ_E_NPOS_RANGE: return "N position beyond sequence length";
    case FQZ_E_GZ_HEADER: return "gzip: invalid header";
    case FQZ_E_GZ_CHECKSUM: return "gzip: invalid checksum";
    case FQZ_E_GZ_CORRUPT: return "flate: corrupt input";
    case FQZ_E_GZ_TRUNC: return "unexpected EOF";
    case FQZ_E_CUDA: return "CUDA error";
    case FQZ_E_NO_DEVICE: return "no CUDA device (libfqzgpu has no CPU fallback)";
    case FQZ_E_INVALID_ARG: return "invalid argument";
    case FQZ_E_NEED_MORE: return "window holds no complete block";
    case FQZ_E_TOO_LARGE: return "input larger than one device window; use the streaming calls";
    default: return "unknown error";
    }
}

// page-locked host buffers for the shim's window buffers: copies from / to them run at PCIe rate and overlap with
// kernels; pageable memory (a Go slice, malloc) is staged by the driver at a fraction of that
extern "C" void *fqz_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void fqz_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int fqz_set_option(fqz_ctx *c, int key, uint64_t value) {
    if (!c) return FQZ_E_INVALID_ARG;
    switch (key) {
    case FQZ_OPT_WINDOW_BYTES:
        if (value && value < ((u64)1 << 20)) return FQZ_E_INVALID_ARG;
        c->opt_window_bytes = value > ((u64)3 << 30) ? ((u64)3 << 30) : value;  // window offsets are u32
        return FQZ_OK;
    case FQZ_OPT_HOST_WINDOW_BYTES:
        if (value && value < ((u64)1 << 20)) return FQZ_E_INVALID_ARG;
        c->opt_host_window_bytes = value > ((u64)3 << 30) ? ((u64)3 << 30) : value;
        return FQZ_OK;
    case FQZ_OPT_FRONTEND:
        if (value > 2) return FQZ_E_INVALID_ARG;
        c->opt_frontend = (int)value;
        return FQZ_OK;
    case FQZ_OPT_SERIAL_ENTROPY:
        if (value > 1) return FQZ_E_INVALID_ARG;
        c->opt_serial_entropy = (int)value;
        return FQZ_OK;
    case FQZ_OPT_HUF_KERNELS:
        if (value > 1) return FQZ_E_INVALID_ARG;
        c->opt_huf_single = (int)value;
        return FQZ_OK;
    case FQZ_OPT_RECORD_MATCH:
        if (value > 1) return FQZ_E_INVALID_ARG;
        c->opt_no_record_match = value ? 0 : 1;
        return FQZ_OK;
    case FQZ_OPT_GZ_CHUNK_BYTES:
        if (value && (value < 256 || value > ((u64)1 << 30) || (value & 3u))) return FQZ_E_INVALID_ARG;
        c->opt_gz_chunk_bytes = value;
        return FQZ_OK;
    default: return FQZ_E_INVALID_ARG;
    }
}

extern "C" void fqz_stats_reset(fqz_ctx *c) {
    if (!c) return;
    c->prof.clear();
    c->launches_base = g_fqz_launches;
}
extern "C" void *fqz_get_stream(fqz_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" void fqz_profile_enable(fqz_ctx *c, int on) {
    if (c) c->prof.on = on != 0;
}
extern "C" int fqz_get_stats(fqz_ctx *c, fqz_stats *out) {
    if (!c || !out) return FQZ_E_INVALID_ARG;
    cudaStreamSynchronize(c->stream);
    c->prof.collect();
    memset(out, 0, sizeof *out);
    out->launches = g_fqz_launches - c->launches_base;
    out->n_stages = ST_COUNT_;
    for (int i = 0; i < ST_COUNT_; i++) {
        out->stage_name[i] = k_stage_names[i];
        out->stage_ms[i] = c->prof.ms[i];
        out->stage_launches[i] = c->prof.launches[i];
        out->stage_bytes[i] = c->prof.bytes[i];
    }
    return FQZ_OK;
}

// ---------------------------------------------------------------------------------- device-wide exclusive scan
int fqz_scan_excl_u32(fqz_ctx *c, u32 *d, u64 n, u64 stride, u32 narr) {
    if (n == 0) return FQZ_OK;
    u32 ntiles = (u32)((n + FQZ_SCAN_TILE - 1) / FQZ_SCAN_TILE);
    if (ntiles == 1) {
        fqz_launch_scan_apply(d, n, stride, narr, nullptr, 1, c->stream);
        return FQZ_OK;
    }
    u32 *sums = (u32 *)c->arena.alloc((size_t)narr * ntiles * sizeof(u32));
    if (!sums) {
        c->err = "arena: out of device memory (scan)";
        return FQZ_E_CUDA;
    }
    fqz_launch_scan_partial(d, n, stride, narr, sums, ntiles, c->stream);
    FQZ_TRY(fqz_scan_excl_u32(c, sums, ntiles, ntiles, narr));
    fqz_launch_scan_apply(d, n, stride, narr, sums, ntiles, c->stream);
    return FQZ_OK;
}
