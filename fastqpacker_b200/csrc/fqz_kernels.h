// fqz_kernels.h — launch geometry + host-callable launchers of the device code.
#pragma once
#include "fqz_common.cuh"

// newline scan: one CTA = 16 KiB of text, 64 contiguous bytes per thread
#define FQZ_NL_THREADS 256
#define FQZ_NL_TILE (FQZ_NL_THREADS * 64)
// device-wide scan: 2048 elements per CTA
#define FQZ_SCAN_THREADS 256
#define FQZ_SCAN_PER_THREAD 8
#define FQZ_SCAN_TILE (FQZ_SCAN_THREADS * FQZ_SCAN_PER_THREAD)
// record metadata: 8 lanes per record
#define FQZ_META_THREADS 256
#ifndef FQZ_META_GROUP
#define FQZ_META_GROUP 4  // lanes per record (4: 2.1 ms per 9.2 GB step, 8: 3.3 ms, 16: 5.2 ms on B200)
#endif
// stream scatter: 8 lanes per record, 64 records per CTA staged through 40 KiB of shared memory
#define FQZ_SC_THREADS 256
#ifndef FQZ_SC_GROUP
#define FQZ_SC_GROUP 4  // lanes per record (4: 5.1 ms per 9.2 GB step, 8: 6.1 ms, 16: 7.7 ms on B200)
#endif
#ifndef FQZ_SC_RPC
#define FQZ_SC_RPC 64
#endif
#ifndef FQZ_SC_SMEM
#define FQZ_SC_SMEM 40960
#endif

void fqz_launch_newline_count(const u8 *text, u64 n, u64 lo, u32 *tile_counts, u32 ntiles, cudaStream_t s);
// single pass count + index (look-back over the tiles); look: ntiles + 1 zeroed u64
void fqz_launch_newline_scan(const u8 *text, u64 n, u64 lo, u32 *line_end, u32 cap_lines, unsigned long long *look, u32 ntiles, u32 *total_lines,
                             cudaStream_t s);
void fqz_launch_find_newline(const u8 *text, u64 n, const u32 *tile_prefix, u32 ntiles, u32 target, u64 *out_pos, cudaStream_t s);
void fqz_launch_newline_index(const u8 *text, u64 n, u64 lo, const u32 *tile_prefix, u32 ntiles, u32 *line_end, u32 max_lines, cudaStream_t s);
void fqz_launch_scan_partial(const u32 *data, u64 n, u64 stride, u32 narr, u32 *sums, u32 ntiles, cudaStream_t s);
void fqz_launch_scan_apply(u32 *data, u64 n, u64 stride, u32 narr, const u32 *sums, u32 ntiles, cudaStream_t s);
void fqz_launch_record_meta(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, u32 *sizes, u64 stride,
                            FqzWinStatus *st, u64 phred_records, cudaStream_t s);
void fqz_launch_decide_phred(const FqzWinStatus *st, u32 *phred64, cudaStream_t s);
// fused single pass: sizes + look-back scan + scatter (fqz_frontend.cu); look: 5 zeroed u64 per CTA of FQZ_SC_RPC records, ticket: one zeroed u32;
// st->pad bit 0 is set when a stream outgrew caps[] (the window must then be redone by the separate kernels)
void fqz_launch_phred_min(const u8 *text, const u32 *line_end, u64 nrec, FqzWinStatus *st, cudaStream_t s);
void fqz_launch_scatter_fused(const u8 *text, const u32 *line_end, u64 R, u64 rec_base, u32 tail_lines, u32 *offs, u64 stride, const u32 *phred64,
                              u8 *const streams[6], const u32 caps[5], FqzWinStatus *st, unsigned long long *look, u32 *ticket, cudaStream_t s);
void fqz_launch_scatter(const u8 *text, const u32 *line_end, u64 R, const u32 *offs, u64 stride, const u32 *phred64, u8 *const streams[6],
                        cudaStream_t s);
