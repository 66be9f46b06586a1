"""ctypes binding of the CPU oracle (oracle/fqz_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  fastqpacker_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfqzoracle.so")

ERRORS = {
    -1: "invalid FASTQ: header line must start with @",
    -2: "invalid FASTQ: separator line must start with +",
    -3: "invalid FASTQ: sequence and quality lengths must match",
    -4: "ambiguous bases beyond position 65536",
    -5: "invalid magic bytes: not an FQZ file",
    -6: "unsupported file version",
    -7: "unexpected EOF",
    -8: "zstd decode error",
    -9: "truncated header data",
    -10: "truncated plus-line payload data",
    -11: "truncated sequence data",
    -12: "truncated quality data",
    -13: "truncated length data",
    -14: "truncated N position data",
    -15: "output buffer too small",
    -16: "libzstd not loadable",
    -17: "N position out of range",
}

STREAM_NAMES = ("seqPacked", "quality", "headers", "plusLines", "nPositions", "seqLengths")


class OracleError(Exception):
    def __init__(self, code: int, extra: str = ""):
        self.code = code
        super().__init__(ERRORS.get(code, f"oracle error {code}") + (f" ({extra})" if extra else ""))


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fqz_oracle.c", "fqz_synth_cpu.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libfqzoracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, sz, szp = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        L.orc_detect_encoding.argtypes = [u8p, sz]
        L.orc_detect_encoding.restype = C.c_int
        for f in (L.orc_normalize_quality, L.orc_denormalize_quality):
            f.argtypes = [u8p, sz, C.c_int]
            f.restype = None
        for f in (L.orc_delta_encode, L.orc_delta_decode):
            f.argtypes = [u8p, sz]
            f.restype = None
        L.orc_pack_bases.argtypes = [u8p, sz, u8p, C.c_void_p]
        L.orc_pack_bases.restype = sz
        L.orc_unpack_bases.argtypes = [u8p, C.c_void_p, sz, sz, u8p]
        L.orc_unpack_bases.restype = C.c_int
        L.orc_parse.argtypes = [u8p, sz, sz, C.c_void_p, szp, szp]
        L.orc_parse.restype = C.c_int
        L.orc_encode_streams.argtypes = [u8p, sz, sz, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_encode_streams.restype = C.c_int
        L.orc_zstd_bound.argtypes = [sz]
        L.orc_zstd_bound.restype = sz
        L.orc_zstd_compress.argtypes = [u8p, sz, C.c_int, u8p, sz, szp]
        L.orc_zstd_compress.restype = C.c_int
        L.orc_zstd_compress_adv.argtypes = [u8p, sz, C.c_int, C.c_int, C.c_int, C.c_int, u8p, sz, szp]
        L.orc_zstd_compress_adv.restype = C.c_int
        L.orc_zstd_decompress.argtypes = [u8p, sz, u8p, sz, szp]
        L.orc_zstd_decompress.restype = C.c_int
        L.orc_xxh64.argtypes = [u8p, sz, C.c_ulonglong]
        L.orc_xxh64.restype = C.c_ulonglong
        L.orc_compress.argtypes = [u8p, sz, C.c_uint32, C.c_int, C.c_int, C.c_int, u8p, sz, szp, C.c_void_p]
        L.orc_compress.restype = C.c_int
        L.orc_decompress.argtypes = [u8p, sz, u8p, sz, szp, C.c_void_p]
        L.orc_decompress.restype = C.c_int
        L.orc_decode_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, u8p, sz, szp]
        L.orc_decode_streams.restype = C.c_int
        L.orc_block_streams.argtypes = [u8p, sz, sz, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_block_streams.restype = C.c_int
        L.orc_decompress_mt.argtypes = [u8p, sz, C.c_int, u8p, sz, szp]
        L.orc_decompress_mt.restype = C.c_int
        L.orc_synth.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, u8p, sz, szp]
        L.orc_synth.restype = C.c_int
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        _lib = L
    return _lib


def _in(b) -> tuple:
    """bytes-like -> (pointer, length, keepalive)"""
    if isinstance(b, (bytes, bytearray)):
        buf = (C.c_ubyte * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")
        return C.cast(buf, C.c_void_p), len(b), buf
    import numpy as np

    a = np.ascontiguousarray(b, dtype=np.uint8)
    return C.c_void_p(a.ctypes.data), a.size, a


# ---------------------------------------------------------------- encoder KAT-level helpers
def detect_encoding(quals: list[bytes]) -> int:
    p, n, k = _in(b"".join(quals))
    return lib().orc_detect_encoding(p, n)


def _inplace(fn, data: bytes, *extra) -> bytes:
    buf = C.create_string_buffer(bytes(data), max(1, len(data)))
    fn(C.cast(buf, C.c_void_p), len(data), *extra)
    return buf.raw[: len(data)]


def normalize_quality(q: bytes, phred64: int) -> bytes:
    return _inplace(lib().orc_normalize_quality, q, phred64)


def denormalize_quality(q: bytes, phred64: int) -> bytes:
    return _inplace(lib().orc_denormalize_quality, q, phred64)


def delta_encode(q: bytes) -> bytes:
    return _inplace(lib().orc_delta_encode, q)


def delta_decode(q: bytes) -> bytes:
    return _inplace(lib().orc_delta_decode, q)


def pack_bases(seq: bytes):
    n = len(seq)
    packed = C.create_string_buffer(max(1, (n + 3) // 4))
    npos = (C.c_uint16 * max(1, min(n, 65536)))()
    p, _, k = _in(seq)
    nn = lib().orc_pack_bases(p, n, C.cast(packed, C.c_void_p), C.cast(npos, C.c_void_p))
    return packed.raw[: (n + 3) // 4], list(npos[:nn])


def unpack_bases(packed: bytes, npos: list[int], seqlen: int) -> bytes:
    out = C.create_string_buffer(max(1, seqlen))
    arr = (C.c_uint16 * max(1, len(npos)))(*npos)
    p, _, k = _in(packed)
    rc = lib().orc_unpack_bases(p, C.cast(arr, C.c_void_p), len(npos), seqlen, C.cast(out, C.c_void_p))
    if rc:
        raise OracleError(rc)
    return out.raw[:seqlen]


# ---------------------------------------------------------------- parser
def parse(text: bytes, max_records: int = 100000):
    """Returns (records, consumed); records = list of (header, sequence, plus, quality)."""
    p, n, k = _in(text)
    cap = min(max_records, n // 4 + 1)
    fields = (C.c_uint64 * (8 * max(1, cap)))()
    nrec, consumed = C.c_size_t(0), C.c_size_t(0)
    rc = lib().orc_parse(p, n, cap, C.cast(fields, C.c_void_p), C.byref(nrec), C.byref(consumed))
    if rc:
        raise OracleError(rc, f"after {nrec.value} records")
    t = bytes(text)
    recs = []
    for i in range(nrec.value):
        f = fields[8 * i : 8 * i + 8]
        recs.append((t[f[0] : f[0] + f[1]], t[f[2] : f[2] + f[3]], t[f[4] : f[4] + f[5]], t[f[6] : f[6] + f[7]]))
    return recs, consumed.value


# ---------------------------------------------------------------- block level
def encode_streams(text, max_records: int = 100000, phred64: int = -1):
    """FASTQ chunk -> dict(streams=[6 bytes], nrec, consumed, phred64, orig_seq, orig_qual)."""
    p, n, k = _in(text)
    caps = [n // 4 + n // 64 + 64, n + 64, n + n // 2 + 64, n + n // 2 + 64, 2 * n + 64, n + 64]
    while True:
        bufs = [C.create_string_buffer(c) for c in caps]
        outp = (C.c_void_p * 6)(*[C.cast(b, C.c_void_p) for b in bufs])
        capa = (C.c_size_t * 6)(*caps)
        lens = (C.c_size_t * 6)()
        info = (C.c_uint64 * 6)()
        rc = lib().orc_encode_streams(p, n, max_records, phred64, outp, capa, lens, info)
        if rc == -15:
            caps = [max(c, l) for c, l in zip(caps, lens)]
            continue
        if rc:
            raise OracleError(rc, f"record {info[5] if rc == -4 else info[0]}")
        return dict(
            streams=[bufs[i].raw[: lens[i]] for i in range(6)],
            nrec=info[0],
            consumed=info[1],
            phred64=info[2],
            orig_seq=info[3],
            orig_qual=info[4],
        )


def decode_streams(streams, nrec: int, phred64: int) -> bytes:
    keep = [_in(s) for s in streams]
    inp = (C.c_void_p * 6)(*[k[0] for k in keep])
    lens = (C.c_size_t * 6)(*[k[1] for k in keep])
    cap = sum(k[1] for k in keep) * 5 + 16 * nrec + 64
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    rc = lib().orc_decode_streams(inp, lens, nrec, phred64, C.cast(out, C.c_void_p), cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out.raw[: n.value]


# ---------------------------------------------------------------- zstd stand-in
def zstd_compress(data, level: int = 1) -> bytes:
    p, n, k = _in(data)
    cap = lib().orc_zstd_bound(n) + 64
    out = C.create_string_buffer(cap)
    m = C.c_size_t(0)
    rc = lib().orc_zstd_compress(p, n, level, C.cast(out, C.c_void_p), cap, C.byref(m))
    if rc:
        raise OracleError(rc)
    return out.raw[: m.value]


def zstd_compress_adv(data, level: int = 1, window_log: int = 0, content_size: int = -1, checksum: int = -1) -> bytes:
    """libzstd frame with an explicit window log and with / without content size and checksum (test support)."""
    p, n, k = _in(data)
    cap = lib().orc_zstd_bound(n) + 64
    out = C.create_string_buffer(cap)
    m = C.c_size_t(0)
    rc = lib().orc_zstd_compress_adv(p, n, level, window_log, content_size, checksum, C.cast(out, C.c_void_p), cap, C.byref(m))
    if rc:
        raise OracleError(rc)
    return out.raw[: m.value]


def zstd_decompress(data, cap: int | None = None) -> bytes:
    p, n, k = _in(data)
    cap = cap or max(1 << 20, n * 64)
    while True:
        out = C.create_string_buffer(cap)
        m = C.c_size_t(0)
        rc = lib().orc_zstd_decompress(p, n, C.cast(out, C.c_void_p), cap, C.byref(m))
        if rc == -15:
            cap = m.value
            continue
        if rc:
            raise OracleError(rc)
        return out.raw[: m.value]


def xxh64(data, seed: int = 0) -> int:
    p, n, k = _in(data)
    return lib().orc_xxh64(p, n, seed)


# ---------------------------------------------------------------- whole file
def compress(text, block_size: int = 0, level: int = 1, threads: int = 1, version: int = 2, return_info=False):
    import numpy as np

    p, n, k = _in(text)
    cap = n + n // 8 + (1 << 16)
    out = np.empty(cap, dtype=np.uint8)
    m = C.c_size_t(0)
    info = (C.c_uint64 * 4)()
    rc = lib().orc_compress(p, n, block_size, level, threads, version, C.c_void_p(out.ctypes.data), cap, C.byref(m), info)
    if rc:
        raise OracleError(rc, f"record {info[3]}")
    res = out[: m.value].tobytes()
    if return_info:
        return res, dict(records=info[0], blocks=info[1], phred64=info[2])
    return res


def decompress(fqz, cap: int | None = None, return_info=False):
    import numpy as np

    p, n, k = _in(fqz)
    cap = cap or max(1 << 20, n * 16)
    while True:
        out = np.empty(cap, dtype=np.uint8)
        m = C.c_size_t(0)
        info = (C.c_uint64 * 5)()
        rc = lib().orc_decompress(p, n, C.c_void_p(out.ctypes.data), cap, C.byref(m), info)
        if rc == -15:
            cap = m.value
            continue
        if rc:
            raise OracleError(rc)
        res = out[: m.value].tobytes()
        if return_info:
            return res, dict(version=info[0], flags=info[1], blocks=info[2], records=info[3], block_size=info[4])
        return res


def block_streams(fqz, block_index: int = 0):
    """Decoded six streams of one block of a .fqz (zstd decoded by libzstd)."""
    p, n, k = _in(fqz)
    outs = (C.c_void_p * 6)()
    lens = (C.c_size_t * 6)()
    nrec = C.c_uint32(0)
    rc = lib().orc_block_streams(p, n, block_index, outs, lens, C.byref(nrec))
    if rc:
        raise OracleError(rc)
    res = []
    for i in range(6):
        res.append(C.string_at(outs[i], lens[i]) if lens[i] else b"")
        if outs[i]:
            lib().orc_free(outs[i])
    return res, nrec.value


def decompress_mt(fqz, threads: int, cap: int):
    """Multi-threaded whole-file decode (block-parallel like compress.go:630-719) -> numpy uint8."""
    import numpy as np

    p, n, k = _in(fqz)
    out = np.empty(cap, dtype=np.uint8)
    m = C.c_size_t(0)
    rc = lib().orc_decompress_mt(p, n, threads, C.c_void_p(out.ctypes.data), cap, C.byref(m))
    if rc:
        raise OracleError(rc)
    return out[: m.value]


def compress_np(text, threads: int = 1, level: int = 1):
    """compress() returning a numpy view (no bytes copy) for the timed CPU baseline."""
    import numpy as np

    p, n, k = _in(text)
    cap = n + n // 8 + (1 << 16)
    out = np.empty(cap, dtype=np.uint8)
    m = C.c_size_t(0)
    info = (C.c_uint64 * 4)()
    rc = lib().orc_compress(p, n, 0, level, threads, 2, C.c_void_p(out.ctypes.data), cap, C.byref(m), info)
    if rc:
        raise OracleError(rc, f"record {info[3]}")
    return out[: m.value]


def synth(kind: int, seed: int, first: int, count: int):
    """CPU twin of the device FASTQ generator -> numpy uint8."""
    import numpy as np

    cap = count * (400 if kind == 0 else 760) + 4096
    out = np.empty(cap, dtype=np.uint8)
    m = C.c_size_t(0)
    rc = lib().orc_synth(kind, seed, first, count, C.c_void_p(out.ctypes.data), cap, C.byref(m))
    if rc:
        raise OracleError(rc)
    return out[: m.value]
