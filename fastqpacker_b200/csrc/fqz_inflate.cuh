// fqz_inflate.cuh — shared types of the gzip / DEFLATE input stage (fqz_inflate.cu, fqz_api_gzip.cu).
//
// Reference: cmd/fqpack/main.go:142-174 — in compress mode an input whose name ends in ".gz" or whose
// first two bytes are 1f 8b is read through Go's compress/gzip (multistream: concatenated members decode as
// one stream, CRC-32 and ISIZE of every member verified).  SURVEY.md §8 row f3.
#pragma once
#include "fqz_common.cuh"

// status of one chunk after a decode pass
enum {
    GZ_ST_NONE = 0,
    GZ_ST_REACHED = 1,    // stopped exactly on its target (the start of the next chunk)
    GZ_ST_END = 2,        // input ended cleanly behind a member trailer
    GZ_ST_OVERSHOOT = 3,  // walked past its target without landing on it: the target was not a real restart point
    GZ_ST_ERR_HEADER = 16,    // gzip: invalid header           (compress/gzip ErrHeader)
    GZ_ST_ERR_CORRUPT = 17,   // flate: corrupt input            (compress/flate CorruptInputError)
    GZ_ST_ERR_TRUNC = 18,     // unexpected EOF                  (io.ErrUnexpectedEOF)
    GZ_ST_ERR_CHECKSUM = 19,  // gzip: invalid checksum         (compress/gzip ErrChecksum: CRC-32 or ISIZE)
    GZ_ST_ERR_INTERNAL = 20   // the two decode passes disagree (a bug, never an input error)
};

// restart points
enum { GZ_AT_NONE = 0, GZ_AT_BLOCK = 1, GZ_AT_MEMBER = 2 };

#define GZ_WINDOW 32768u
#define GZ_NO_TARGET (~0ull)

// One chunk = the part of the compressed stream one warp decodes: from its restart point to the restart point
// of the next chunk.
struct GzChunk {
    u64 start_bit;     // GZ_AT_BLOCK: bit offset of a deflate block header; GZ_AT_MEMBER: 8 * byte offset of a gzip member header
    u64 target_bit;    // where to stop (start of the next chunk of the chain), GZ_NO_TARGET = end of input
    u32 start_type, target_type;
    // results of a decode pass
    u64 out_len;       // bytes decoded
    u64 end_bit;       // where the pass stopped
    u64 err_bit;       // position of the error (status >= 16)
    u64 after_member;  // bytes decoded since the last member header met inside the chunk; ~0 = none met
    u32 status;
    u32 members;       // members whose trailer lies inside the chunk
    u32 overflow;      // the speculative pass ran out of symbol or member room: the chunk is decoded again at exact sizes
    u32 pad0;
    // places: provisional ones for the speculative pass, exact ones for a chunk that is decoded again
    u64 sym_ptr;       // device address of the chunk's symbols (u16)
    u64 sym_cap;       // symbols there is room for
    u64 mem_ptr;       // device address of its member records (out_end relative to the chunk)
    u32 mem_cap;
    // filled by the host once every chunk's size is known
    u32 member_base;   // index of its first member in the file-wide member table
    u64 out_off;       // place of the chunk's text
    u32 window_valid;  // bytes of the 32 KiB window in front of the chunk that belong to the same member
    u32 pad1;
};

struct GzMember {
    u64 out_end;       // end of the member's data in the output
    u64 trailer_byte;  // offset of its 8-byte trailer
    u32 crc, isize;    // as stored in the trailer
    u32 crc_acc;       // XOR of the span CRCs, each shifted to the member's end (k_gz_crc)
    u32 pad;
};

struct GzArgs {
    const u32 *src;  // compressed input, 4-byte aligned, 64 readable bytes behind n
    u64 n;
    GzChunk *chunks;
    u32 nchunks;
    u32 chunk_bytes;
    const u32 *list;  // chunks to run
    u32 nlist;
    u32 bgzf_only;    // k_gz_find: look for BGZF member headers only
    u32 spec;         // k_gz_decode<write>: 1 = speculative pass (room may run out, results recorded), 0 = exact places (results verified)
    u8 *win;          // base window of every group of chunks, GZ_WINDOW bytes each
    u16 *maps;        // per chunk of the chain: its window as a map of the group's base window (nullptr: no chunk needs one)
    u32 gsize;        // chunks per group
    u8 *out;
    GzMember *members;
    u32 nmembers;
    unsigned long long *err;  // min over (bit position << 8 | status)
    u32 pw[48];       // x^(8 * 2^k) mod P (reflected CRC-32 polynomial), for shifting CRCs
};

void fqz_launch_gz_find(const GzArgs &a, cudaStream_t s);
void fqz_launch_gz_decode(const GzArgs &a, bool write, cudaStream_t s);
void fqz_launch_gz_members(const GzArgs &a, cudaStream_t s);
void fqz_launch_gz_windows(const GzArgs &a, cudaStream_t s);
void fqz_launch_gz_resolve(const GzArgs &a, cudaStream_t s);
void fqz_launch_gz_crc(const GzArgs &a, u64 out_len, cudaStream_t s);
