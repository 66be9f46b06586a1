// fqz_zstd.h — host-visible structures and launchers of the GPU zstd (RFC 8878) entropy stage.
#pragma once
#include "fqz_common.cuh"

// Every stream is cut into frames of at most FQZ_ZFRAME bytes; each frame is an independent zstd
// frame holding one block, so frames encode AND decode in parallel and each carries its own
// XXH64 content checksum.  Decoders (klauspost DecodeAll, libzstd) decode
// concatenated frames back to back.
#define FQZ_ZFRAME 65536u
#define FQZ_ZFRAME_LOG 16
// item-matcher policy: small frames, so that one 1 GiB window holds enough frames (one warp each) to fill 148 SMs
#ifndef FQZ_ZFRAME_ITEMS
#define FQZ_ZFRAME_ITEMS 16384u
#endif
// literals-only policy: frames of up to eight 16 KiB blocks sharing one Huffman tree (k_zenc_huf)
#define FQZ_ZFRAME_ENT 131072u
#define FQZ_ZBLOCK_ENT 16384u
// duplicate-record search: a stream is searched, and — when it holds duplicates — coded as ONE frame, per segment of this
// many bytes (16 literals-only frames): long enough that a recurring read is sent again only every 2 MiB, short enough
// that a 15 MB quality stream is eight frames that decode (and checksum) side by side instead of one serial chain
#define FQZ_ZSEG (16u * FQZ_ZFRAME_ENT)
#define FQZ_ZSLOT(len) ((((size_t)(len) + 512) + 15) & ~(size_t)15)  // output slot of one frame: raw fallback + table scratch always fit
#define FQZ_ZWS(len) (((size_t)(len)*6 + 1023) & ~(size_t)63)  // LZ workspace: literals + match / sequence arrays (<= len/2 + 2 entries, 10 B each)

struct ZFrame {
    u64 src;      // device address of the frame's content
    u64 dst_off;  // into the slot arena
    u64 ws_off;   // into the LZ workspace arena (policy AUTO only)
    u32 src_len;  // 1..FQZ_ZFRAME (policy AUTO) / 1..FQZ_ZFRAME_ENT (policy ENTROPY)
    u32 policy;   // FQZ_ZPOLICY_* or FQZ_ZPOLICY_ITEMS
    u64 items;      // item matcher: device address of the u32 item-start offsets of the stream (0 = fixed stride)
    u32 item_base;  // offset of this frame's first byte in the coordinates of items[]
    u32 item_count; // entries of items[] including the end sentinel; fixed stride: the stride in bytes
    u32 index_of;   // 1 + number of the FQZ_ZPOLICY_INDEX frame that lists this frame (0 = none)
    u32 pad;        // 1 + number of the frame's ZRStream (0 = none)
};
#define FQZ_ZPOLICY_ITEMS 2  // internal: LZ by item matcher (needs the item boundaries of the stream)
// internal: not a zstd frame but a SKIPPABLE frame (RFC 8878 §3.1.2) in front of a stream that was cut
// into many frames.  It lists, for each of the src_len frames that follow, the compressed size and —
// for the length-prefixed item streams — where the first item that starts inside the frame's content
// lies and how many items start there, so that the GPU decoder can find all frames at once instead
// of hopping header by header, and can walk the item chains of all frames in parallel.  Every
// decoder, the reference's included, skips it; ours trusts it only as far as it proves true.
// Layout: magic 0x184D2A5E, u32 payload size, then the payload: 'FQZI', u32 nframes, u32 content
// bytes per frame (the last one may be shorter), nframes x {u32 compressed size, u16 offset of the
// first item start, u16 item starts}.
#define FQZ_ZPOLICY_INDEX 3
#define FQZ_ZINDEX_MAGIC 0x184D2A5Eu
#define FQZ_ZINDEX_SIG 0x495A5146u  // "FQZI" little-endian
#define FQZ_ZINDEX_MIN 4u           // streams of fewer frames carry no index
#define FQZ_ZINDEX_HDR 20u
#define FQZ_ZINDEX_BYTES(n) (FQZ_ZINDEX_HDR + 8u * (u32)(n))

#ifndef ZENC_WARPS
#define ZENC_WARPS 4
#endif

void fqz_launch_xxh64(const ZFrame *frames, u32 nframes, u32 *hashes, cudaStream_t s);
// fills the index frames (FQZ_ZPOLICY_INDEX) from the sizes of the frames behind them
void fqz_launch_zindex(const ZFrame *frames, u32 nframes, u8 *slots, u32 *out_sizes, const u32 *lzflags, cudaStream_t s);
// index: optional list of frame numbers to encode (nullptr = frames 0..nidx-1)
// lzflags: optional per-stream flags of fqz_launch_rec_match (frames of flagged streams are left to fqz_launch_lzrec)
// scratch: fqz_zenc_huf_scratch(nidx) bytes -> the three-kernel version (histograms, one-warp-per-frame plan, encode);
// nullptr -> the single kernel
size_t fqz_zenc_huf_scratch(u32 nidx);
void fqz_launch_zenc_huf(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u32 *out_sizes, const u32 *lzflags,
                         u8 *scratch, cudaStream_t s);
// One segment (<= FQZ_ZSEG bytes) of a literals-only stream of one fqz block that carries its record boundaries
// (packed bases, qualities): the unit of the duplicate-record search (fqz_zstd_enc.cu, "duplicated records").
// Its frames carry pad = 1 + the segment's number in this table.
struct ZRStream {
    u64 src;           // device address of the segment's bytes
    u64 items;         // device address of the scanned record offsets of the whole window (u32)
    u32 len;           // bytes
    u32 item_base;     // stream coordinates of the segment's first byte
    u32 rec0, nrec;    // the BLOCK's records inside items[] (the segment's own are found by binary search: k_rec_ranges)
    u32 first_frame;   // number of the stream's first zstd frame in the frame table (its slots are contiguous)
    u32 blk0, nblk;    // 16 KiB blocks of the stream, numbered through all streams of the batch
    u32 pad;
};
// pairs duplicated records (cand, laid out like the scanned offset arrays `offs_base`; keys = scratch of the same
// shape), flags the streams that hold enough of them (flags[ns] = any) and hashes their content.  dupcnt: ns zeroed words.
// ranges: 2 * ns words of scratch (first record / record count of every segment)
void fqz_launch_rec_match(const ZRStream *rs, u32 ns, u32 max_records, const u32 *offs_base, u32 *keys_base, u32 *cand_base, u32 *flags,
                          u32 *dupcnt, u32 *ranges, u32 *hashes, cudaStream_t s);
// codes blocks [g0, gend) of the flagged streams (parse, literals, sequences) into their staging (pool_out: all
// blocks of the batch, pool_ws / parsed: gend - g0 blocks); _close strings the blocks of every flagged stream together
size_t fqz_lzrec_pool_ws(u32 nblocks);
size_t fqz_lzrec_pool_out(u32 nblocks);
void fqz_launch_lzrec(const ZRStream *rs, u32 ns, const u32 *flags, const u32 *offs_base, const u32 *cand_base, u8 *pool_ws, u8 *pool_out, u32 g0,
                      u32 gend, u32 *parsed, u32 *bsizes, cudaStream_t s);
void fqz_launch_lzrec_close(const ZRStream *rs, u32 ns, u32 max_blocks, const u32 *flags, const u32 *hashes, const u8 *pool_out, const u32 *bsizes,
                            const ZFrame *frames, u8 *slots, u32 *out_sizes, cudaStream_t s);
// parsed: 2 u32 per work item, scratch between the item matcher and the entropy kernel (lz == 2 only)
// hashes_ready: optional event the stream waits for before the kernels that read `hashes` (the checksums may still be
// computed on another stream while the matcher and the literals run)
void fqz_launch_zenc(const ZFrame *frames, const u32 *index, u32 nidx, const u32 *hashes, u8 *slots, u8 *ws, u32 *out_sizes, int lz,
                     u32 *parsed, cudaStream_t s, cudaEvent_t hashes_ready = nullptr);
