"""Host-side planning for one input split across several GPUs.

fqz blocks are independent (each stream of each block is its own chain of zstd frames, no
dictionary, no cross-block state: internal/compress/compress.go:523-528), so the codec shards with
no data-path collective.  What has to be agreed between ranks is only WHERE the blocks are:

* decompress: the container has no index (compress.go:613-625,721-736); the host walks the 36-byte
  (v1: 32-byte) block headers once and hands every rank a contiguous run of whole blocks.
* compress: blocks are cut every 100 000 records = 400 000 newlines (compress.go:48-52,71), so every
  rank counts the newlines of its byte slice, the counts are gathered (the one tiny exchange), and the
  byte ranges are re-cut on block boundaries.  Rank 0's first block decides the Phred flag
  (compress.go:146-164), which is broadcast with the plan.

The functions here are pure (bytes / integers in, plans out) and are exercised on CPU with
world_size-2 gloo in tests/test_sharding.py; the byte-level work they schedule runs on the GPUs
through libfqzgpu.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Sequence, Tuple

BLOCK_RECORDS = 100_000  # reference: compress.DefaultBlockSize + batchPool (F2)
LINES_PER_BLOCK = 4 * BLOCK_RECORDS
MAGIC = b"FQZ\x00"


@dataclass
class BlockRef:
    offset: int  # of the block header
    size: int  # header + payloads
    records: int


def walk_container(fqz: bytes) -> Tuple[int, int, List[BlockRef]]:
    """(version, flags, blocks) of a .fqz — the serial producer of compress.go:690-758."""
    if len(fqz) < 4:
        raise ValueError("unexpected EOF")
    if fqz[:4] != MAGIC:
        raise ValueError("invalid magic bytes: not an FQZ file")
    if len(fqz) < 10:
        raise ValueError("unexpected EOF")
    version, flags = fqz[4], fqz[9]
    if version not in (1, 2):
        raise ValueError(f"unsupported file version: {version}")
    hsz = 32 if version == 1 else 36
    pos, blocks = 10, []
    while pos < len(fqz):
        if pos + hsz > len(fqz):
            raise ValueError("reading block header: unexpected EOF")
        v = struct.unpack_from(f"<{hsz // 4}I", fqz, pos)
        payload = sum(v[1:6]) if version == 1 else sum(v[1:7])
        if pos + hsz + payload > len(fqz):
            raise ValueError("reading compressed data: unexpected EOF")
        blocks.append(BlockRef(pos, hsz + payload, v[0]))
        pos += hsz + payload
    return version, flags, blocks


def plan_decompress(blocks: Sequence[BlockRef], world: int) -> List[Tuple[int, int]]:
    """Contiguous [first, last) block ranges per rank, balanced by compressed bytes."""
    total = sum(b.size for b in blocks)
    plan, first, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        last = first
        while last < len(blocks) and (acc + blocks[last].size / 2 <= target or r == world - 1):
            acc += blocks[last].size
            last += 1
        plan.append((first, last))
        first = last
    return plan


def shard_container(fqz: bytes, world: int) -> List[bytes]:
    """Stand-alone .fqz per rank (file header + its run of blocks); outputs concatenate in rank order."""
    _, _, blocks = walk_container(fqz)
    out = []
    for first, last in plan_decompress(blocks, world):
        if first == last:
            out.append(fqz[:10])
            continue
        a, b = blocks[first].offset, blocks[last - 1].offset + blocks[last - 1].size
        out.append(fqz[:10] + fqz[a:b])
    return out


def slice_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Even byte slices [a, b) used for the newline count."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def plan_compress(newline_positions_per_slice: Sequence[Sequence[int]], n: int, world: int,
                  bounds: Sequence[Tuple[int, int]] | None = None) -> List[Tuple[int, int]]:
    """Block-aligned byte ranges per rank.  `bounds`: the slices the ranks hold when they are not the even ones
    of `slice_bounds` (a BGZF input is cut on member boundaries, `plan_bgzf`).

    newline_positions_per_slice[r] holds, for slice r, the absolute offsets of the newlines that end a
    block (every 400 000th newline of the file), which rank r can compute from its local count and the
    gathered counts of the slices before it (see `block_cut_candidates`).  Rank r takes the blocks
    whose first byte falls into its even slice; the last rank also takes the tail.
    """
    cuts = sorted(p + 1 for ps in newline_positions_per_slice for p in ps)  # first byte after each full block
    cuts = [c for c in cuts if c < n]
    starts = [0] + cuts  # block starts
    bounds = list(bounds) if bounds is not None else slice_bounds(n, world)
    plan = []
    for r, (a, b) in enumerate(bounds):
        mine = [s for s in starts if a <= s < b]
        if not mine:
            plan.append((0, 0))
            continue
        first = mine[0]
        nxt = [s for s in starts if s >= b]
        last = nxt[0] if nxt else n
        plan.append((first, last))
    return plan


def block_cut_candidates(local_newlines: Sequence[int], lines_before: int) -> List[int]:
    """Offsets (absolute) of the newlines in this slice that end a block, given the number of
    newlines in all earlier slices (the gathered counts)."""
    out = []
    k = LINES_PER_BLOCK - (lines_before % LINES_PER_BLOCK)  # this slice's k-th newline (1-based) ends a block
    while k <= len(local_newlines):
        out.append(local_newlines[k - 1])
        k += LINES_PER_BLOCK
    return out


# ---------------------------------------------------------------------------------- gzip input over several GPUs
# A plain gzip member is one serial chain and goes to ONE GPU whole (fqz_compress_gz).  A BGZF file (bgzip, the
# usual container of large FASTQ) is a sequence of independent members that each say how large they are (BSIZE in
# the 'BC' extra field) and how much text they hold (ISIZE in the trailer): it can be cut between members and the
# text offset of every cut is known without inflating anything.  Every rank inflates its run of members on its GPU
# (fqz_gunzip_device) and owns that slice of the text; from there on the flow is the one of a plain-text input
# (newline counts, block cuts, plan_compress with these slices as `bounds`).
@dataclass
class BgzfMember:
    offset: int  # of the member header in the file
    size: int  # BSIZE + 1
    text_offset: int  # of its first byte in the inflated text
    text_size: int  # ISIZE


def walk_bgzf(gz: bytes) -> List[BgzfMember]:
    """Members of a BGZF file, hopping from header to header (SAM spec 4.1).  ValueError for anything else."""
    out, pos, text = [], 0, 0
    n = len(gz)
    while pos < n:
        if n - pos < 18 or gz[pos : pos + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError(f"not a BGZF member at offset {pos}")
        xlen = gz[pos + 10] | (gz[pos + 11] << 8)
        q, end, bsize = pos + 12, pos + 12 + xlen, None
        while q + 4 <= end:
            slen = gz[q + 2] | (gz[q + 3] << 8)
            if gz[q : q + 2] == b"BC" and slen == 2:
                bsize = gz[q + 4] | (gz[q + 5] << 8)
            q += 4 + slen
        if bsize is None or pos + bsize + 1 > n or bsize + 1 < 12 + xlen + 8:
            raise ValueError(f"not a BGZF member at offset {pos}")
        size = bsize + 1
        isize = struct.unpack_from("<I", gz, pos + size - 4)[0]
        out.append(BgzfMember(pos, size, text, isize))
        pos += size
        text += isize
    return out


def plan_bgzf(members: Sequence[BgzfMember], world: int) -> List[Tuple[int, int]]:
    """Contiguous [first, last) member ranges per rank, balanced by bytes of TEXT."""
    total = sum(m.text_size for m in members)
    plan, first, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        last = first
        while last < len(members) and (acc + members[last].text_size / 2 <= target or r == world - 1):
            acc += members[last].text_size
            last += 1
        plan.append((first, last))
        first = last
    return plan


def bgzf_slices(members: Sequence[BgzfMember], plan: Sequence[Tuple[int, int]]) -> List[Tuple[Tuple[int, int], Tuple[int, int]]]:
    """Per rank: (byte range of its members in the file, byte range of their text)."""
    out = []
    end_file = members[-1].offset + members[-1].size if members else 0
    end_text = members[-1].text_offset + members[-1].text_size if members else 0
    for first, last in plan:
        fa = members[first].offset if first < len(members) else end_file
        ta = members[first].text_offset if first < len(members) else end_text
        fb = members[last].offset if last < len(members) else end_file
        tb = members[last].text_offset if last < len(members) else end_text
        out.append(((fa, fb), (ta, tb)))
    return out


def merge_compressed(parts: Sequence[bytes]) -> bytes:
    """Ordered host gather (collectAndWriteResults, compress.go:365-403): rank 0's part keeps the file
    header, the others contribute their blocks only.  Later parts may come with a file header of their
    own (a stand-alone .fqz) or without (`fqz_compress_shard(emit_file_header=0)`); a block header starts
    with NumRecords <= 100 000, which can never read as the magic, so the two are told apart safely."""
    out = bytearray(parts[0])
    for p in parts[1:]:
        out += p[10:] if p[:4] == MAGIC else p
    return bytes(out)
