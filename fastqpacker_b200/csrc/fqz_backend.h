// fqz_backend.h — tables and launchers of the decompress back end (fqz_backend.cu).
#pragma once
#include "fqz_common.cuh"

// Error kinds in the order blockReader.writeRecord meets them inside one record
// (internal/compress/compress.go:944-975): length, N positions, header, sequence (+ N restore),
// plus line, quality.  The device keeps min over (record << 8 | kind), i.e. the first failure in
// the reference's sequential order.
enum {
    BK_E_TRUNC_LEN = 1,
    BK_E_TRUNC_NPOS = 2,
    BK_E_TRUNC_HEADER = 3,
    BK_E_TRUNC_SEQ = 4,
    BK_E_NPOS_RANGE = 5,
    BK_E_TRUNC_PLUS = 6,
    BK_E_TRUNC_QUAL = 7
};

// One fqz block after the entropy stage: six decoded streams resident in HBM.
// Stream order: 0 seqPacked, 1 quality, 2 headers, 3 plusLines, 4 nPositions, 5 seqLengths.
struct BkBlock {
    u64 stream[6];  // device addresses, each 64-byte aligned with FQZ_PAD readable slack
    u32 size[6];
    u32 nrec;
    u32 pad;
    u64 rec_base;   // window-global index of the block's first record
    // compressed form of the header / plus / N-position streams (device address, bytes) when it is at
    // hand: its index frame (FQZ_ZPOLICY_INDEX) lets the item chains be walked frame by frame in
    // parallel.  0 = none (block-level entry points that start from decoded streams).
    u64 hint[3];
    u32 hint_size[3];
    u32 pad2;
};
// per block, written by k_record_sizes: 64-bit totals so that the host can rule out u32 wrap
struct BkTotals {
    u64 packed, fastq, bases;
};
struct FqzDecStatus {
    u64 err_key;  // min over (record << 8 | BK_E_*); ~0 = none
};

// container walk (k_walk_container): one entry per fqz block
struct FqzBlockEntry {
    u64 payload_off;  // offset of the first payload byte (after the block header)
    u32 nrec;
    u32 size[6];      // compressed sizes in stream order (v1: size[3] = 0)
    u32 pad;
};
struct FqzWalkResult {
    u64 next_pos;   // offset of the first block not entered into the table
    u32 nblocks;    // entries written
    u32 status;     // 0 ok / more, 1 truncated block header or payload
    u32 done;       // 1 when next_pos reached the end of the file
    u32 pad;
};

void fqz_launch_walk_container(const u8 *fqz, u64 n, u64 pos, u32 version, FqzBlockEntry *table, u32 cap, u64 max_bytes,
                               FqzWalkResult *res, cudaStream_t s);
// max_segments: upper bound on the frames of any one chain (0 = no hints: serial walk only); ok: nblocks * 3 zeroed counters
void fqz_launch_walk_prefixes(const BkBlock *blks, u32 nblocks, u32 max_segments, u32 *offs, u32 *ok, FqzDecStatus *st, cudaStream_t s);
void fqz_launch_record_sizes(const BkBlock *blks, u32 nblocks, u32 max_nrec, const u32 *offs, u32 *sz, u64 stride, BkTotals *tot,
                             FqzDecStatus *st, cudaStream_t s);
// error path: first sequence / quality truncation or out-of-range N position in front of the failure `limit_key` (~0: none known)
void fqz_launch_first_error(const BkBlock *blks, u32 nblocks, const u32 *offs, u64 limit_key, FqzDecStatus *st, cudaStream_t s);
void fqz_launch_emit(const BkBlock *blks, u32 nblocks, u32 max_nrec, const u32 *offs, const u32 *sc, u64 stride, u32 phred64, u8 *out,
                     FqzDecStatus *st, cudaStream_t s);
