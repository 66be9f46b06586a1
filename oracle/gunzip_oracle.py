"""CPU restatement of the gzip input path of compress mode.  TEST INFRASTRUCTURE ONLY (see oracle/fqz_oracle.py).

Reference: cmd/fqpack/main.go:142-174 — `wrapInputMaybeGzip` hands the input to Go's `compress/gzip`
reader when its name ends in ".gz" or it starts with 1f 8b (`inputHasGzipMagic`).  Go's standard library
is not part of /root/reference (it is the toolchain's), so this module restates the published behaviour of
`compress/gzip/gunzip.go` (go1.22: `Reader.readHeader`, `Reader.Read`, multistream mode) around a raw
DEFLATE decoder — Python's zlib stands in for `compress/flate`, both implement RFC 1951.

Pinned by: tests/test_gzip_oracle.py — the reference's own gzip cases (cmd/fqpack/main_test.go:12-161: plain,
gzip by extension, gzip by magic) and Python's independent `gzip` module on every generated file.

Errors are returned as the class the Go reader reports:
  HEADER   gzip.ErrHeader   "gzip: invalid header"
  CHECKSUM gzip.ErrChecksum "gzip: invalid checksum"
  CORRUPT  flate.CorruptInputError
  TRUNC    io.ErrUnexpectedEOF (io.EOF for an empty input)
"""
from __future__ import annotations

import zlib

HEADER, CHECKSUM, CORRUPT, TRUNC = "HEADER", "CHECKSUM", "CORRUPT", "TRUNC"
# the library's codes for the same classes (include/fqzgpu.h)
CODES = {HEADER: -18, CHECKSUM: -19, CORRUPT: -20, TRUNC: -21}


class GunzipError(Exception):
    def __init__(self, kind: str, offset: int):
        self.kind = kind
        self.offset = offset
        self.code = CODES[kind]
        super().__init__(f"{kind} at {offset}")


def has_gzip_magic(buf: bytes) -> bool:
    """inputHasGzipMagic, main.go:164-174"""
    return len(buf) >= 2 and buf[0] == 0x1F and buf[1] == 0x8B


def _read_header(b: bytes, p: int) -> int:
    """gunzip.go readHeader: returns the offset of the deflate data."""
    n = len(b)
    if n - p < 10:
        raise GunzipError(TRUNC, n)  # io.ReadFull: EOF at 0 bytes is handled by the caller, else unexpected EOF
    if b[p] != 0x1F or b[p + 1] != 0x8B or b[p + 2] != 8:
        raise GunzipError(HEADER, p)
    flg = b[p + 3]
    q = p + 10
    if flg & 4:  # flagExtra
        if n - q < 2:
            raise GunzipError(TRUNC, n)
        xlen = b[q] | (b[q + 1] << 8)
        q += 2
        if n - q < xlen:
            raise GunzipError(TRUNC, n)
        q += xlen
    for bit in (8, 16):  # flagName, flagComment: readString, at most 511 bytes + NUL
        if flg & bit:
            i = 0
            while True:
                if i >= 512:
                    raise GunzipError(HEADER, q)
                if q >= n:
                    raise GunzipError(TRUNC, n)
                c = b[q]
                q += 1
                if c == 0:
                    break
                i += 1
    if flg & 2:  # flagHdrCrc
        if n - q < 2:
            raise GunzipError(TRUNC, n)
        want = b[q] | (b[q + 1] << 8)
        if want != (zlib.crc32(b[p:q]) & 0xFFFF):
            raise GunzipError(HEADER, q)
        q += 2
    return q


def gunzip(b: bytes) -> bytes:
    """gzip.NewReader(b) + io.ReadAll in multistream mode."""
    b = bytes(b)
    n = len(b)
    if n == 0:
        raise GunzipError(TRUNC, 0)  # NewReader: io.EOF
    out = []
    p = 0
    while True:
        q = _read_header(b, p)
        d = zlib.decompressobj(-15)
        try:
            data = d.decompress(b[q:])
        except zlib.error:
            raise GunzipError(CORRUPT, q)
        if not d.eof:
            raise GunzipError(TRUNC, n)
        t = n - len(d.unused_data)  # trailer
        if n - t < 8:
            raise GunzipError(TRUNC, n)
        crc = int.from_bytes(b[t : t + 4], "little")
        isize = int.from_bytes(b[t + 4 : t + 8], "little")
        if crc != zlib.crc32(data) or isize != (len(data) & 0xFFFFFFFF):
            raise GunzipError(CHECKSUM, t)
        out.append(data)
        p = t + 8
        if p == n:  # readHeader: io.EOF at a member boundary ends the stream
            return b"".join(out)
