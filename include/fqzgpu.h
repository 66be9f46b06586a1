/*
 * fqzgpu.h — C-ABI of libfqzgpu.so, the B200 (sm_100a) implementation of fqpack's block codec.
 *
 * This is the drop-in boundary for vertti/fastqpacker: the Go package internal/compress keeps its
 * exported API (Compress / Decompress / Options / DecompressOptions, compress.go:70-82,125,558)
 * and binds these entry points through cgo instead of running its CPU
 * workers.  Plain pointers and sizes only; the library never keeps a caller pointer after a call
 * returns (cgo rule), owns all device memory, pinned staging buffers and CUDA streams inside
 * fqz_ctx, and has NO CPU fallback: without a CUDA device every call fails with FQZ_E_NO_DEVICE.
 *
 * Reference citations are file:line in vertti/fastqpacker.
 */
#ifndef FQZGPU_H
#define FQZGPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FQZ_ABI_VERSION 1

/* ---- result codes.  Negative codes map 1:1 onto the reference's error texts so that the Go shim
 *      can return errors with the same wording. */
#define FQZ_OK 0
#define FQZ_E_HEADER_AT (-1)    /* "invalid FASTQ: header line must start with @"            internal/fqparser/parser.go:143 */
#define FQZ_E_PLUS (-2)         /* "invalid FASTQ: separator line must start with +"         parser.go:164 */
#define FQZ_E_LEN_MISMATCH (-3) /* "invalid FASTQ: sequence and quality lengths must match"  parser.go:180 */
#define FQZ_E_LONG_N (-4)       /* "record %q: sequence length %d has ambiguous bases beyond position 65536; ..." internal/compress/compress.go:484 */
#define FQZ_E_MAGIC (-5)        /* "invalid magic bytes: not an FQZ file"                    internal/fqformat/container.go:54 */
#define FQZ_E_VERSION (-6)      /* "unsupported file version: %d"                            compress.go:572 */
#define FQZ_E_TRUNC_FILE (-7)   /* "reading block header:" / "reading compressed data: unexpected EOF"  compress.go:727,732 */
#define FQZ_E_ZSTD (-8)         /* "decompressing sequences|quality|headers|plus-line payload|N positions|lengths: ..." compress.go:787-813 */
#define FQZ_E_TRUNC_HEADER (-9) /* "truncated header data"             compress.go:979 */
#define FQZ_E_TRUNC_PLUS (-10)  /* "truncated plus-line payload data"  compress.go:1002 */
#define FQZ_E_TRUNC_SEQ (-11)   /* "truncated sequence data"           compress.go:1020 */
#define FQZ_E_TRUNC_QUAL (-12)  /* "truncated quality data"            compress.go:1033 */
#define FQZ_E_TRUNC_LEN (-13)   /* "truncated length data"             compress.go:1048 */
#define FQZ_E_TRUNC_NPOS (-14)  /* "truncated N position data"         compress.go:1057 */
#define FQZ_E_NOSPACE (-15)     /* output buffer too small; *out_len = bytes required, nothing consumed */
#define FQZ_E_GZ_HEADER (-18)   /* "gzip: invalid header"    Go compress/gzip ErrHeader, reached through cmd/fqpack/main.go:151-156 */
#define FQZ_E_GZ_CHECKSUM (-19) /* "gzip: invalid checksum"  ErrChecksum: CRC-32 or ISIZE of a member */
#define FQZ_E_GZ_CORRUPT (-20)  /* "flate: corrupt input before offset %d"  compress/flate CorruptInputError */
#define FQZ_E_GZ_TRUNC (-21)    /* "unexpected EOF" inside a gzip member ("EOF" for an empty input) */
#define FQZ_E_NPOS_RANGE (-17)  /* N position >= sequence length (the reference panics: encoder/sequence.go:218-220) */
#define FQZ_E_CUDA (-32)        /* CUDA runtime failure; text in fqz_last_error() */
#define FQZ_E_NO_DEVICE (-33)   /* no usable CUDA device: there is no CPU fallback */
#define FQZ_E_INVALID_ARG (-34)
#define FQZ_E_NEED_MORE (-35)   /* streaming: window holds no complete block and is_last == 0 */
#define FQZ_E_TOO_LARGE (-36)   /* single call larger than FQZ_MAX_WINDOW; use the streaming calls */

#define FQZ_MAX_WINDOW ((size_t)3 << 30) /* bytes of FASTQ / .fqz handled by one device pass */
#define FQZ_DEVICE_SLACK 64               /* readable bytes required after a caller-supplied DEVICE buffer */

typedef struct fqz_ctx fqz_ctx;

/* One context per process and GPU (reference: one zstd encoder/decoder per worker goroutine,
 * compress.go:113-122,281,671).  Not thread-safe: one caller thread at a time per context. */
int fqz_init(int device, fqz_ctx **out);
void fqz_destroy(fqz_ctx *ctx);
const char *fqz_strerror(int code);
/* Detail of the last failure on this context (record index, CUDA error text, stream name). */
const char *fqz_last_error(const fqz_ctx *ctx);
int fqz_abi_version(void);

/* Tuning knobs (all optional; 0 restores the default).  The codec's results do not depend on them, only where
 * windows are cut and which front-end kernels run. */
#define FQZ_OPT_WINDOW_BYTES 1      /* FASTQ bytes per device pass of the compress calls (default 3e9, min 1 MiB) */
#define FQZ_OPT_HOST_WINDOW_BYTES 2 /* same for the host-buffer calls (default 1 GiB: their first window must be uploaded first) */
#define FQZ_OPT_RECORD_MATCH 3      /* 1 (default): search packed bases / qualities for duplicated records and code them as matches; 0: literals only */
#define FQZ_OPT_FRONTEND 4          /* front-end kernels (results identical): 0 (default) newline count, index, metadata, scans, scatter as
                                     * separate kernels; 1 count + index in one pass (look-back over the tiles); 2 metadata + scatter
                                     * fused as well.  1 and 2 read the text less often but measured slower on B200 */
#define FQZ_OPT_HUF_KERNELS 5       /* literals-only frames: 0 (default) histogram / plan / encode kernels, 1 one kernel per frame (same format, the
                                     * Huffman codes may differ) */
#define FQZ_OPT_SERIAL_ENTROPY 6    /* 0 (default): the item-stream kernels run on a second stream beside the literals-only coder (their launch
                                     * tails overlap); 1: one after the other (per-stage timings are only meaningful this way) */
#define FQZ_OPT_GZ_CHUNK_BYTES 7    /* gzip input: compressed bytes decoded by one warp (default: input size / (36 x SMs), 32 KiB .. 4 MiB;
                                     * multiple of 4, >= 256).  Results do not depend on it */
int fqz_set_option(fqz_ctx *ctx, int key, uint64_t value);

/* Page-locked host memory for the caller's window buffers (the Go shim reads the file into these instead of Go
 * slices): transfers from / to them run at PCIe rate and overlap with the kernels.  Any host memory is accepted
 * by the calls below; pageable memory is simply slower.  NULL when the allocation fails. */
void *fqz_host_alloc(size_t bytes);
void fqz_host_free(void *p);

/* ---- whole buffer, HOST memory: replaces the bodies of compress.Compress / compress.Decompress
 *      (compress.go:125-192, 558-604).  Output = complete .fqz file (10-byte header + blocks of
 *      100 000 records, F2) / complete FASTQ text.  header_block_size is echoed into the file
 *      header only (0 -> 100000, compress.go:129-131). */
size_t fqz_compress_bound(size_t fastq_bytes);
int fqz_compress(fqz_ctx *ctx, const uint8_t *fastq, size_t n, uint32_t header_block_size, uint8_t *out, size_t out_cap, size_t *out_len);
int fqz_decompress(fqz_ctx *ctx, const uint8_t *fqz, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);
/* ---- archive introspection and verification (the reference's ROADMAP.md PR-009 `fqpack info`, PR-008 `fqpack check`)
 * fqz_info: version, flags, block / record counts and per-stream compressed sizes from one hop over the block
 *   headers (container.go:48-152); header arithmetic on the host buffer, no device work.
 * fqz_check: decodes every stream of every block (frame checksums verified) and rebuilds every record on the GPU
 *   without writing the FASTQ anywhere; same error codes and texts as fqz_decompress. */
typedef struct fqz_file_info {
    uint32_t version;            /* 1 or 2 */
    uint32_t flags;              /* bit 1 = Phred+64 (container.go:28) */
    uint32_t header_block_size;  /* records per block as written in the file header */
    uint32_t reserved;
    uint64_t blocks, records;
    uint64_t compressed[6];      /* seqPacked, quality, headers, plusLines (0 in v1 files), nPositions, seqLengths */
    uint64_t original_seq, original_qual; /* sums of OriginalSeqSize / OriginalQualSize */
} fqz_file_info;
int fqz_info(fqz_ctx *ctx, const uint8_t *fqz, size_t n, fqz_file_info *out);
int fqz_check(fqz_ctx *ctx, const uint8_t *fqz, size_t n, uint64_t *records, uint64_t *fastq_bytes);

/* ---- block index and random access.  The container has no index (a reader must hop from block header to block header,
 * compress.go:690-758; ROADMAP.md lists one as future work), so this library builds it on demand as a SIDE TABLE that
 * costs one hop over the headers and can be stored beside the archive: no format change, every .fqz the reference
 * writes has one.
 * fqz_block_index: one entry per block, in file order.  Host arithmetic only, needs no context and no GPU.  *count = blocks
 *   in the file, also when cap is too small (FQZ_E_NOSPACE then; out may be NULL with cap 0 to size the table).  Errors
 *   as fqz_info: FQZ_E_MAGIC, FQZ_E_VERSION, FQZ_E_TRUNC_FILE (*count = whole blocks in front of the cut).
 * fqz_decompress_blocks: the FASTQ of blocks [first_block, first_block + num_blocks) alone = records
 *   [index[first_block].first_record, ...): only those blocks' bytes cross PCIe.  Blocks are independent
 *   (compress.go:523-528), which is also what lets several GPUs share a file (fastqpacker_b200/sharding.py).
 *   FQZ_E_INVALID_ARG when the range reaches past the last block.  Of a file that was cut, the whole blocks in front of the
 *   cut still decode (FQZ_E_TRUNC_FILE for a range that reaches into it). */
typedef struct fqz_block_ref {
    uint64_t offset;        /* of the block header from the start of the file */
    uint64_t size;          /* block header + its payloads */
    uint64_t first_record;  /* records in the blocks in front of this one */
    uint32_t records;       /* NumRecords (container.go:84) */
    uint32_t reserved;
    uint64_t original_seq, original_qual; /* OriginalSeqSize / OriginalQualSize: bases and quality bytes of the block */
} fqz_block_ref;
int fqz_block_index(const uint8_t *fqz, size_t n, fqz_block_ref *out, size_t cap, size_t *count);
int fqz_decompress_blocks(fqz_ctx *ctx, const uint8_t *fqz, size_t n, uint64_t first_block, uint64_t num_blocks, uint8_t *out,
                          size_t out_cap, size_t *out_len);

/* ---- gzip input (cmd/fqpack/main.go:123-174).  In compress mode fqpack reads an input whose name ends in ".gz" or that
 * starts with 1f 8b through Go's compress/gzip: concatenated members (BGZF, pigz -i, cat a.gz b.gz) decode as one
 * stream, every member's CRC-32 and ISIZE are verified, bytes behind the last member that are not a gzip header are
 * "gzip: invalid header".  fqz_is_gzip = inputHasGzipMagic (main.go:164-174).  fqz_gunzip: gzip file in host memory ->
 * text in host memory (FQZ_E_NOSPACE: *out_len = bytes needed).  fqz_gunzip_device: the same between device buffers
 * (16-byte aligned, FQZ_DEVICE_SLACK readable bytes behind the input).  fqz_compress_gz = gzip.NewReader + compress.Compress:
 * the COMPRESSED bytes cross PCIe, the text is inflated and coded in HBM; *fastq_len (optional) = bytes of FASTQ.
 * DEFLATE is decoded in parallel from restart points found in the compressed stream (BGZF member headers,
 * dynamic-block headers), proven by the decoder in front landing exactly on them. */
int fqz_is_gzip(const uint8_t *buf, size_t n);
/* how the last gzip input of this context was cut: [0] chunks, [1] chunks decoded side by side (restart points proven),
 * [2] restart points dropped as false positives, [3] gzip members, [4] chunks that outgrew the scratch room of the
 * speculative pass (more than 6 bytes of text per compressed byte) and were decoded a second time */
int fqz_gunzip_stats(fqz_ctx *ctx, uint64_t out[5]);
int fqz_gunzip(fqz_ctx *ctx, const uint8_t *gz, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);
int fqz_gunzip_device(fqz_ctx *ctx, const void *d_gz, size_t n, void *d_out, size_t out_cap, size_t *out_len);
int fqz_compress_gz(fqz_ctx *ctx, const uint8_t *gz, size_t n, uint32_t header_block_size, uint8_t *out, size_t out_cap, size_t *out_len,
                    size_t *fastq_len);

/* One shard of a file that is split across GPUs on block boundaries (blocks are independent,
 * compress.go:523-528).  phred64: -1 = detect on this shard's first block (the shard
 * holding block 0 of the file, compress.go:146-164), 0 / 1 = the file-global decision made there.
 * emit_file_header: 1 for the shard that starts the file, 0 for the others (blocks only), so that the
 * host gather is a plain concatenation in shard order (collectAndWriteResults, compress.go:365-403).
 * *phred64_used (optional) returns the flag in force. */
int fqz_compress_shard(fqz_ctx *ctx, const uint8_t *fastq, size_t n, uint32_t header_block_size, int phred64, int emit_file_header,
                       uint8_t *out, size_t out_cap, size_t *out_len, int *phred64_used);

/* ---- streaming, HOST memory (Seam B): the Go producer feeds raw windows read from io.Reader and
 *      writes what comes back to io.Writer in order (replaces produceCompressJobs / workers /
 *      collectAndWriteResults, compress.go:240-403, and their decompress twins :630-719).
 *      feed() consumes a whole number of blocks; the caller re-presents the unconsumed tail in
 *      front of the next window.  The first compress feed also emits the file header; Phred is
 *      decided on the first block only (compress.go:146-164).  A feed that does not return FQZ_OK
 *      (FQZ_E_NEED_MORE: no whole block in the window yet; FQZ_E_NOSPACE: *out_len = room the next
 *      block needs) leaves the stream as it was and consumes nothing — not even the file header —
 *      so the caller grows its buffer and presents the same window again. */
typedef struct fqz_cstream fqz_cstream;
typedef struct fqz_dstream fqz_dstream;
int fqz_compress_begin(fqz_ctx *ctx, uint32_t header_block_size, fqz_cstream **s);
int fqz_compress_feed(fqz_cstream *s, const uint8_t *fastq, size_t n, int is_last, uint8_t *out, size_t out_cap, size_t *out_len,
                      size_t *consumed);
void fqz_compress_end(fqz_cstream *s);
int fqz_decompress_begin(fqz_ctx *ctx, fqz_dstream **s);
int fqz_decompress_feed(fqz_dstream *s, const uint8_t *fqz, size_t n, int is_last, uint8_t *out, size_t out_cap, size_t *out_len,
                        size_t *consumed);
void fqz_decompress_end(fqz_dstream *s);

/* ---- whole buffer, DEVICE memory (benchmarks, GPU pipelines): same semantics, pointers are
 *      device pointers on the context's GPU, 16-byte aligned, with FQZ_DEVICE_SLACK readable bytes
 *      after the input.  *out_len is returned on the host. */
int fqz_compress_device(fqz_ctx *ctx, const void *d_fastq, size_t n, uint32_t header_block_size, void *d_out, size_t out_cap,
                        size_t *out_len);
int fqz_decompress_device(fqz_ctx *ctx, const void *d_fqz, size_t n, void *d_out, size_t out_cap, size_t *out_len);

/* One input split across GPUs with the text in DEVICE memory (bench.py --workload cfg5): the reference's
 * producer cuts a block every 100 000 records while parsing (compress.go:240-300); here every rank counts
 * the newlines of its byte slice, the counts are exchanged through the host, and fastqpacker_b200/sharding.py
 * re-cuts the slices on block boundaries.  fqz_count_lines_device: '\n' bytes in d_text[0..n).
 * fqz_find_line_end_device: offset of the k-th (1-based) '\n'.  fqz_compress_shard_device: fqz_compress_shard
 * on device memory (any alignment of d_fastq). */
int fqz_count_lines_device(fqz_ctx *ctx, const void *d_text, size_t n, uint64_t *lines);
int fqz_find_line_end_device(fqz_ctx *ctx, const void *d_text, size_t n, uint64_t k, uint64_t *offset);
int fqz_compress_shard_device(fqz_ctx *ctx, const void *d_fastq, size_t n, uint32_t header_block_size, int phred64, int emit_file_header,
                              void *d_out, size_t out_cap, size_t *out_len, int *phred64_used);

/* ---- block level (Seam A and the parity tests) -------------------------------------------------
 * fqz_encode_streams: FASTQ chunk -> the six pre-entropy streams of its first block
 *   (<= 100 000 records), byte-identical to what compressBlockWithBuffers builds before EncodeAll
 *   (compress.go:474-520).  Stream order: 0 seqPacked, 1 quality, 2 headers, 3 plusLines,
 *   4 nPositions, 5 seqLengths.  phred64: 0 / 1 forced, -1 = detect on these records
 *   (encoder/quality.go:22-49).  info: [0] records [1] bytes consumed [2] phred64 used
 *   [3] OriginalSeqSize [4] OriginalQualSize [5] index of the offending record on error.
 * fqz_decode_streams: the inverse back end = NumRecords x blockReader.writeRecord (compress.go:944-1078);
 *   an empty plusLines stream reproduces the v1 "+\n" rule (compress.go:995-999). */
int fqz_encode_streams(fqz_ctx *ctx, const uint8_t *fastq, size_t n, int phred64, uint8_t *const out[6], const size_t cap[6],
                       size_t len[6], uint64_t info[6]);
int fqz_decode_streams(fqz_ctx *ctx, const uint8_t *const in[6], const size_t len[6], uint32_t num_records, int phred64, uint8_t *out,
                       size_t out_cap, size_t *out_len);

/* Entropy stage alone = zstd.Encoder.EncodeAll / zstd.Decoder.DecodeAll (compress.go:523-528,
 * 785-814; klauspost/compress v1.19.1 is not in the reference tree).  Output of
 * fqz_zstd_compress is a sequence of RFC 8878 frames with content checksums; empty input yields
 * zero bytes.  fqz_zstd_decompress accepts any RFC 8878 frame sequence without dictionaries.
 * policy: FQZ_ZPOLICY_* below. */
#define FQZ_ZPOLICY_AUTO 0     /* LZ77 + Huffman literals + FSE sequences */
#define FQZ_ZPOLICY_ENTROPY 1  /* literals-only blocks: Huffman / RLE / raw, no match search */
int fqz_zstd_compress(fqz_ctx *ctx, const uint8_t *src, size_t n, int policy, uint8_t *dst, size_t cap, size_t *out_len);
int fqz_zstd_decompress(fqz_ctx *ctx, const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);

/* ---- measurement support (bench.py): kernel launch counter and per-stage device timings taken
 *      with CUDA events on the library's own stream. */
#define FQZ_MAX_STAGES 32
typedef struct {
    uint64_t launches;                 /* kernels launched by this library since fqz_stats_reset */
    uint32_t n_stages;
    const char *stage_name[FQZ_MAX_STAGES];
    double stage_ms[FQZ_MAX_STAGES];   /* accumulated device time, only while profiling is on */
    uint64_t stage_launches[FQZ_MAX_STAGES];
    uint64_t stage_bytes[FQZ_MAX_STAGES]; /* algorithmic bytes attributed to the stage */
} fqz_stats;
void fqz_stats_reset(fqz_ctx *ctx);
/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that a
 * caller can bracket calls with its own CUDA events (bench.py times on this stream). */
void *fqz_get_stream(fqz_ctx *ctx);
void fqz_profile_enable(fqz_ctx *ctx, int on);
int fqz_get_stats(fqz_ctx *ctx, fqz_stats *out);

/* ---- synthetic FASTQ generator used by the tests and bench (counter-based RNG, identical bytes
 *      on every device and in the CPU twin tests/synth.py).  kind: 0 = Illumina 150 bp Phred+33
 *      (BASELINE config 2), 1 = variable length 50-300, Phred+64, 5 % N, plus payloads (config 4).
 *      Writes records [first_record, first_record + num_records) to d_out. */
int fqz_synth_device(fqz_ctx *ctx, int kind, uint64_t seed, uint64_t first_record, uint64_t num_records, void *d_out, size_t out_cap,
                     size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif /* FQZGPU_H */
