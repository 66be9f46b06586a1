// fqz_host.h — host-side plumbing of libfqzgpu: context, device arena, stage profiler.
#pragma once
#include <string>
#include <vector>

#include "../../include/fqzgpu.h"
#include "fqz_kernels.h"

#define FQZ_CUDA_TRY(ctx, expr)                                                      \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);         \
            return FQZ_E_CUDA;                                                        \
        }                                                                             \
    } while (0)
#define FQZ_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != FQZ_OK) return rc__; \
    } while (0)

// Bump allocator over a few large cudaMalloc chunks; reset between windows.
struct Arena {
    struct Chunk {
        u8 *p;
        size_t cap, used;
    };
    std::vector<Chunk> chunks;
    size_t min_chunk = (size_t)64 << 20;
    void *alloc(size_t bytes);  // 256-byte aligned, FQZ_PAD slack; nullptr on failure
    void reset();               // keeps memory; coalesces several chunks into one
    void release();
    size_t capacity() const;
};

enum FqzStage {
    ST_NL_COUNT = 0,
    ST_NL_INDEX,
    ST_SCAN,
    ST_RECORD_META,
    ST_SCATTER,
    ST_ZENC_ENTROPY,
    ST_ZENC_LZ,
    ST_ZENC_DUP,
    ST_XXH64,
    ST_ASSEMBLE,
    ST_ZDEC_SCAN,
    ST_ZDEC_LITERALS,
    ST_ZDEC_SEQUENCES,
    ST_ZDEC_EXECUTE,
    ST_WALK,
    ST_OFFSETS,
    ST_EMIT,
    ST_COPY,
    ST_GZ_FIND,
    ST_GZ_DECODE,
    ST_GZ_RESOLVE,
    ST_GZ_CRC,
    ST_COUNT_
};

struct Profiler {
    bool on = false;
    struct Rec {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Rec> open;
    std::vector<cudaEvent_t> pool;
    double ms[ST_COUNT_] = {0};
    u64 launches[ST_COUNT_] = {0};
    u64 bytes[ST_COUNT_] = {0};
    cudaEvent_t get();
    void collect();
    void clear();
};

// Host-buffer entry points: copy engine pipeline.  Input is uploaded in chunks on its own stream
// (an event per chunk gates the compute stream), results are downloaded on a third stream while the
// next window is being coded.  Device staging buffers persist across calls (no cudaMalloc per call).
#define FQZ_IO_CHUNK ((size_t)32 << 20)
struct IoPipe {
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    u8 *d_in = nullptr;
    size_t in_cap = 0;
    u8 *d_out[2] = {nullptr, nullptr};
    size_t out_cap[2] = {0, 0};
    cudaEvent_t ev_out[2] = {nullptr, nullptr};  // download of slot k finished
    bool out_busy[2] = {false, false};
    cudaEvent_t ev_done = nullptr;                // compute of a window finished
    std::vector<cudaEvent_t> ev_chunk;
    size_t n = 0, gated = 0;  // input bytes of the current call / bytes the compute stream already waits for
};

struct fqz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_aux = nullptr;         // second compute stream: the item-stream kernels run beside the literals-only coder
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_hash = nullptr;
    Arena arena;
    Profiler prof;
    std::string err;
    FqzWinStatus *d_status = nullptr;  // device
    u32 *d_phred = nullptr;            // device: file-global Phred flag
    u8 *h_pin = nullptr;               // pinned host scratch for small readbacks / uploads
    size_t h_pin_cap = 0;
    u8 *h_io = nullptr;                // pinned staging for host-buffer entry points
    size_t h_io_cap = 0;
    int sm_count = 148;
    u64 launches_base = 0;
    IoPipe io;
    // fqz_set_option
    u64 opt_window_bytes = 0;       // device window of the compress calls (0 = default)
    u64 opt_host_window_bytes = 0;  // window of the host-buffer compress calls (0 = default)
    int opt_frontend = 0;           // 0: separate kernels (count, index, metadata, scans, scatter); 1: count + index in one pass (look-back); 2: metadata + scatter fused as well
    u64 fused_windows = 0, legacy_windows = 0;
    int opt_serial_entropy = 0;     // 1: item-stream kernels and literals-only coder one after the other on one stream
    int opt_huf_single = 0;         // 1: literals-only frames by the single kernel (k_zenc_huf) instead of histogram / plan / encode
    int opt_no_record_match = 0;    // 1: packed bases / qualities always literals-only (no duplicate-record search)
    u64 opt_gz_chunk_bytes = 0;     // compressed bytes per warp of the gzip input stage (0 = from the input size)
    // gzip input stage (fqz_api_gzip.cu): the inflated FASTQ text lives here between inflate and compress
    u8 *gz_text = nullptr;
    u64 gz_stats[5] = {0, 0, 0, 0, 0};  // last inflate: chunks cut, chunks decoded in parallel, restart points dropped as false, members, chunks decoded twice
    size_t gz_text_cap = 0;
};
int fqz_frontend_init_device();
int fqz_zstd_enc_init_device();
int fqz_zstd_dec_init_device();

// RAII stage marker: CUDA events around the launches of one pipeline stage when profiling is on
struct StageScope {
    fqz_ctx *c;
    int stage;
    u64 l0;
    cudaEvent_t a = nullptr;
    StageScope(fqz_ctx *ctx, int st, u64 bytes);
    ~StageScope();
};

int fqz_pin_reserve(fqz_ctx *c, size_t bytes);
// Small stream-ordered copy between device memory and the PINNED scratch (c->h_pin) done by a kernel
// over the mapped host memory instead of cudaMemcpyAsync: copy-engine queues are served in submission
// order, so a 4-byte readback queued behind gigabytes of pipelined upload would stall the compute
// stream until the whole upload has drained.  Never pass pageable memory.
int fqz_pin_copy(fqz_ctx *c, void *dst, const void *src, size_t bytes);
int fqz_io_reserve(fqz_ctx *c, size_t bytes);
// ---- copy pipeline (fqz_ctx.cu)
int fqz_io_upload(fqz_ctx *c, const u8 *host, size_t n);             // starts the chunked upload into c->io.d_in
int fqz_io_gate(fqz_ctx *c, size_t upto, size_t *avail);             // compute stream waits until [0, upto) has arrived
int fqz_io_out_acquire(fqz_ctx *c, int slot, size_t bytes, u8 **p);  // device staging for one window's output
int fqz_io_download(fqz_ctx *c, int slot, u8 *host_dst, const u8 *dev_src, size_t bytes);  // after the compute queued so far
int fqz_io_finish(fqz_ctx *c);
void fqz_io_release(fqz_ctx *c);
int fqz_scan_excl_u32(fqz_ctx *c, u32 *d, u64 n, u64 stride, u32 narr);

// ---- compress from a device-resident text into host memory (fqz_api_zstd.cu; used by fqz_compress_gz)
int fqz_compress_text_to_host(fqz_ctx *c, const u8 *d_text, u64 n, u32 header_block_size, u8 *out, size_t out_cap, size_t *out_len);

// ---- zstd decode stage (fqz_api_dec.cu)
struct ZDStream;
struct ZDecodeOut {
    u8 *d_base = nullptr;
    std::vector<u64> off, size;  // per stream
    int err_stream = -1;
};
// max_out: give up with FQZ_E_TOO_LARGE (before allocating) when the streams decode to more bytes; 0 = no limit
int fqz_zdecode_batch(fqz_ctx *c, const std::vector<ZDStream> &streams, u64 max_out, ZDecodeOut &out);

// ---- compress front end (fqz_api_front.cu)
struct FrontOut {
    u64 R = 0;          // records encoded from this window
    u32 nblocks = 0;    // fqz blocks (100 000 records each, last one may be short)
    u64 consumed = 0;   // bytes of text consumed
    u32 phred64 = 0;
    u8 *d_streams[6] = {0};
    std::vector<u32> blk_off[6];  // nblocks+1 boundaries of each stream (bytes)
    std::vector<u32> orig;        // per block: sum of sequence lengths (== quality lengths)
    const u32 *d_offs = nullptr;  // device: 5 arrays (seq, qual, hdr, plus, npos) of R+1 scanned stream offsets
    u64 offs_stride = 0;
};
// phred_mode: -1 detect on the first block of the file (rec_base must be 0), -2 use c->d_phred as
// already decided, 0/1 force.  max_records: cap on records taken (0 = all whole blocks / all).
// skip: bytes in front of the window's first record inside d_text (d_text is 16-byte aligned; the byte at skip - 1 is the '\n' that
// ended the previous window); out.consumed counts from d_text.
int fqz_run_frontend(fqz_ctx *c, const u8 *d_text, u64 n, bool is_last, u64 rec_base, int phred_mode, u64 max_records, FrontOut &out, u32 skip);
