// fqz_zstd_dec.h — tables exchanged between the kernels of the GPU zstd decoder.
#pragma once
#include "fqz_common.cuh"

struct ZDStream {  // compressed bytes of one stream: a chain of zstd frames (possibly none)
    u64 src;       // device address
    u64 csize;
};
struct ZDStreamInfo {
    u32 nframes, nblocks;
    u64 out_bytes;   // sum of frame content sizes (upper bound for frames without a content size)
    u64 lit_bytes;   // sum of regenerated literal sizes
    u64 nseq;        // sum of sequence counts
    u32 status;      // 0 ok, 1 corrupt, 2 unsupported (dictionary)
    u32 frame_base, block_base;  // filled by the host before the second walk
    u64 out_base, lit_base, seq_base;
};
struct ZDFrame {
    u64 dst_off;        // into the output arena
    u64 content_size;   // ~0 when the frame header has none
    u64 out_cap;        // bytes reserved at dst_off
    u64 out_size;       // bytes regenerated (k_zd_execute)
    u64 window;
    u32 first_block, nblocks;
    u32 has_ck, ck;
    u32 stream;
    u32 err;            // 1 corrupt, 2 checksum mismatch
};
struct ZDBlock {
    u64 src;            // device address of the block content
    u64 lit_off;        // into the literal arena
    u64 seq_off;        // sequence index into the sequence arena (3 x u32 per sequence)
    u32 csize;          // stored size (1 for RLE)
    u32 rsize;          // regenerated size for raw / RLE blocks
    u32 frame;
    u32 lit_regen, lit_csize;
    u32 nseq;
    u32 seq_pos;        // offset of the symbol-compression-modes byte
    u32 bits_pos;       // offset of the sequence bitstream
    u32 tab_pos[3];     // offsets of the LL / OF / ML table descriptions
    u32 huf_block;      // block whose literals section holds the Huffman tree (treeless literals)
    u32 fse_block[3];   // block whose table description is reused (repeat mode)
    u32 out_off;        // offset of this block's output inside its frame when every block before it has a size
                        // known from the headers alone (raw, RLE, sequence-free); ~0 otherwise
    u8 type;            // 0 raw, 1 RLE, 2 compressed
    u8 lit_type, lit_streams, lit_hdr;
    u8 modes;
    u8 err;
    u8 pad[2];
};
struct ZDStreamResult {
    u64 size;
    u32 err;
    u32 contiguous;
};

void fqz_launch_zd_hop(const ZDStream *streams, u32 nstreams, ZDStreamInfo *info, ZDFrame *frames, ZDBlock *blocks, int fill, cudaStream_t s);
void fqz_launch_zd_parse(ZDBlock *blocks, u32 nblocks, u32 *cnt, u32 cnt_stride, cudaStream_t s);
void fqz_launch_zd_link(ZDFrame *frames, u32 nframes, ZDBlock *blocks, u32 *cnt, u32 cnt_stride, cudaStream_t s);
void fqz_launch_zd_offsets(ZDBlock *blocks, u32 nblocks, const u32 *cnt, u32 cnt_stride, u32 *seqblk, u32 *litgrp, cudaStream_t s);
void fqz_launch_zd_literals(ZDBlock *blocks, u32 nblocks, const u32 *litgrp, u32 ngroups, const ZDFrame *frames, u8 *litbuf, u8 *out, cudaStream_t s);
#define FQZ_ZD_TAB_BYTES ((3u * 512u + 4u) * 4u)  // decode tables of one block with sequences (k_zd_seq_tables)
void fqz_launch_zd_sequences(ZDBlock *blocks, const ZDFrame *frames, const u32 *seqblk, u32 nsb, const u32 *slot_of, u32 *tabs, u32 *seqbuf, cudaStream_t s);
void fqz_launch_zd_rawcopy(const ZDBlock *blocks, u32 nblocks, const ZDFrame *frames, u8 *out, cudaStream_t s);
void fqz_launch_zd_execute(ZDFrame *frames, u32 nframes, ZDBlock *blocks, const u8 *litbuf, const u32 *seqbuf, u8 *out, cudaStream_t s);
void fqz_launch_zd_checksum(ZDFrame *frames, u32 nframes, const u8 *out, cudaStream_t s);
void fqz_launch_zd_finish(const ZDFrame *frames, const ZDStreamInfo *info, u32 nstreams, ZDStreamResult *res, cudaStream_t s);
void fqz_launch_zd_compact(ZDFrame *frames, const ZDStreamInfo *info, const ZDStreamResult *res, u32 nstreams, u8 *out, cudaStream_t s);
