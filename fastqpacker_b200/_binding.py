"""ctypes binding of the C-ABI declared in include/fqzgpu.h.

`FqzLibrary(path)` wraps one shared object; `fastqpacker_b200._lib.library()` returns the product
instance bound to the in-tree `libfqzgpu.so` and raises if it is missing or no CUDA device
exists (there is no CPU fallback anywhere in this package).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

FQZ_OK = 0
FQZ_E_NOSPACE = -15
FQZ_E_NO_DEVICE = -33
FQZ_MAX_STAGES = 32

STREAM_NAMES = ("seqPacked", "quality", "headers", "plusLines", "nPositions", "seqLengths")


class FqzError(Exception):
    """Error raised by the library; `code` is the FQZ_E_* value, text follows the reference."""

    def __init__(self, code: int, text: str, detail: str = ""):
        self.code = code
        self.detail = detail
        super().__init__(f"{text}" + (f" [{detail}]" if detail else ""))


class _Stats(C.Structure):
    _fields_ = [
        ("launches", C.c_uint64),
        ("n_stages", C.c_uint32),
        ("stage_name", C.c_char_p * FQZ_MAX_STAGES),
        ("stage_ms", C.c_double * FQZ_MAX_STAGES),
        ("stage_launches", C.c_uint64 * FQZ_MAX_STAGES),
        ("stage_bytes", C.c_uint64 * FQZ_MAX_STAGES),
    ]


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        a = buf
        if a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a).view(np.uint8)
        return a
    return np.frombuffer(buf, dtype=np.uint8) if len(buf) else np.zeros(0, dtype=np.uint8)


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data if a.size else None)


class _FileInfo(C.Structure):
    _fields_ = [
        ("version", C.c_uint32),
        ("flags", C.c_uint32),
        ("header_block_size", C.c_uint32),
        ("reserved", C.c_uint32),
        ("blocks", C.c_uint64),
        ("records", C.c_uint64),
        ("compressed", C.c_uint64 * 6),
        ("original_seq", C.c_uint64),
        ("original_qual", C.c_uint64),
    ]


class _BlockRef(C.Structure):
    _fields_ = [
        ("offset", C.c_uint64),
        ("size", C.c_uint64),
        ("first_record", C.c_uint64),
        ("records", C.c_uint32),
        ("reserved", C.c_uint32),
        ("original_seq", C.c_uint64),
        ("original_qual", C.c_uint64),
    ]


class FqzLibrary:
    def __init__(self, path: str):
        self.path = path
        L = self.L = C.CDLL(path)
        vp, sz, szp, i32, u32 = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_uint32
        L.fqz_init.argtypes = [i32, C.POINTER(vp)]
        L.fqz_destroy.argtypes = [vp]
        L.fqz_destroy.restype = None
        L.fqz_strerror.argtypes = [i32]
        L.fqz_strerror.restype = C.c_char_p
        L.fqz_last_error.argtypes = [vp]
        L.fqz_last_error.restype = C.c_char_p
        L.fqz_abi_version.restype = i32
        L.fqz_encode_streams.argtypes = [vp, vp, sz, i32, vp, vp, vp, vp]
        L.fqz_stats_reset.argtypes = [vp]
        L.fqz_stats_reset.restype = None
        L.fqz_profile_enable.argtypes = [vp, i32]
        L.fqz_profile_enable.restype = None
        L.fqz_get_stats.argtypes = [vp, C.POINTER(_Stats)]
        L.fqz_get_stream.argtypes = [vp]
        L.fqz_get_stream.restype = vp
        self._opt(L, "fqz_set_option", [vp, i32, C.c_uint64])
        self._opt(L, "fqz_decode_streams", [vp, vp, vp, u32, i32, vp, sz, szp])
        self._opt(L, "fqz_zstd_compress", [vp, vp, sz, i32, vp, sz, szp])
        self._opt(L, "fqz_zstd_decompress", [vp, vp, sz, vp, sz, szp])
        self._opt(L, "fqz_compress", [vp, vp, sz, u32, vp, sz, szp])
        self._opt(L, "fqz_decompress", [vp, vp, sz, vp, sz, szp])
        self._opt(L, "fqz_compress_shard", [vp, vp, sz, u32, i32, i32, vp, sz, szp, C.POINTER(C.c_int)])
        self._opt(L, "fqz_compress_shard_device", [vp, vp, sz, u32, i32, i32, vp, sz, szp, C.POINTER(C.c_int)])
        self._opt(L, "fqz_count_lines_device", [vp, vp, sz, C.POINTER(C.c_uint64)])
        self._opt(L, "fqz_find_line_end_device", [vp, vp, sz, C.c_uint64, C.POINTER(C.c_uint64)])
        self._opt(L, "fqz_compress_device", [vp, vp, sz, u32, vp, sz, szp])
        self._opt(L, "fqz_decompress_device", [vp, vp, sz, vp, sz, szp])
        self._opt(L, "fqz_info", [vp, vp, sz, C.POINTER(_FileInfo)])
        self._opt(L, "fqz_check", [vp, vp, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
        self._opt(L, "fqz_block_index", [vp, sz, C.POINTER(_BlockRef), sz, szp])
        self._opt(L, "fqz_decompress_blocks", [vp, vp, sz, C.c_uint64, C.c_uint64, vp, sz, szp])
        self._opt(L, "fqz_compress_bound", [sz], restype=sz)
        self._opt(L, "fqz_host_alloc", [sz], restype=vp)
        self._opt(L, "fqz_host_free", [vp], restype=None)
        self._opt(L, "fqz_compress_begin", [vp, u32, C.POINTER(vp)])
        self._opt(L, "fqz_compress_feed", [vp, vp, sz, i32, vp, sz, szp, szp])
        self._opt(L, "fqz_compress_end", [vp], restype=None)
        self._opt(L, "fqz_decompress_begin", [vp, C.POINTER(vp)])
        self._opt(L, "fqz_decompress_feed", [vp, vp, sz, i32, vp, sz, szp, szp])
        self._opt(L, "fqz_decompress_end", [vp], restype=None)
        self._opt(L, "fqz_is_gzip", [vp, sz])
        self._opt(L, "fqz_gunzip_stats", [vp, C.POINTER(C.c_uint64 * 5)])
        self._opt(L, "fqz_gunzip", [vp, vp, sz, vp, sz, szp])
        self._opt(L, "fqz_gunzip_device", [vp, vp, sz, vp, sz, szp])
        self._opt(L, "fqz_compress_gz", [vp, vp, sz, u32, vp, sz, szp, szp])
        self._opt(L, "fqz_synth_device", [vp, i32, C.c_uint64, C.c_uint64, C.c_uint64, vp, sz, szp])

    @staticmethod
    def _opt(L, name, argtypes, restype=C.c_int):
        f = getattr(L, name, None)
        if f is not None:
            f.argtypes = argtypes
            f.restype = restype

    def strerror(self, code: int) -> str:
        return self.L.fqz_strerror(code).decode()

    def block_index(self, fqz) -> list:
        """Side-table index of a .fqz (one hop over the block headers, host only: needs no context and no GPU):
        a dict(offset, size, first_record, records, original_seq, original_qual) per block."""
        a = _as_u8(fqz)
        n = C.c_size_t(0)
        rc = self.L.fqz_block_index(_ptr(a), a.size, None, 0, C.byref(n))
        if rc not in (FQZ_OK, FQZ_E_NOSPACE):
            raise FqzError(rc, self.strerror(rc), f"after {n.value} whole blocks")
        tab = (_BlockRef * max(1, n.value))()
        rc = self.L.fqz_block_index(_ptr(a), a.size, tab, n.value, C.byref(n))
        if rc != FQZ_OK:
            raise FqzError(rc, self.strerror(rc))
        return [dict(offset=int(t.offset), size=int(t.size), first_record=int(t.first_record), records=int(t.records),
                     original_seq=int(t.original_seq), original_qual=int(t.original_qual)) for t in tab[: n.value]]

    def context(self, device: int = 0) -> "FqzContext":
        return FqzContext(self, device)


class FqzContext:
    """One GPU context (one per process and device); not thread-safe."""

    def __init__(self, lib: FqzLibrary, device: int = 0):
        self.lib = lib
        self.h = C.c_void_p()
        rc = lib.L.fqz_init(device, C.byref(self.h))
        if rc != FQZ_OK:
            raise FqzError(rc, lib.strerror(rc))

    def close(self):
        if self.h:
            self.lib.L.fqz_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != FQZ_OK:
            raise FqzError(rc, self.lib.strerror(rc), self.lib.L.fqz_last_error(self.h).decode())

    # ---- block level -------------------------------------------------------------------------
    def encode_streams(self, fastq, phred64: int = -1) -> dict:
        a = _as_u8(fastq)
        n = a.size
        caps = [n // 4 + n // 64 + 64, n + 64, n + n // 2 + 64, n + n // 2 + 64, 2 * n + 64, n + 64]
        while True:
            bufs = [np.empty(c, dtype=np.uint8) for c in caps]
            outp = (C.c_void_p * 6)(*[b.ctypes.data for b in bufs])
            capa = (C.c_size_t * 6)(*caps)
            lens = (C.c_size_t * 6)()
            info = (C.c_uint64 * 6)()
            rc = self.lib.L.fqz_encode_streams(self.h, _ptr(a), n, phred64, outp, capa, lens, info)
            if rc == FQZ_E_NOSPACE:
                caps = [max(c, int(l)) for c, l in zip(caps, lens)]
                continue
            if rc != FQZ_OK:
                e = FqzError(rc, self.lib.strerror(rc), self.lib.L.fqz_last_error(self.h).decode())
                e.record = int(info[5])
                raise e
            return dict(
                streams=[bufs[i][: lens[i]].tobytes() for i in range(6)],
                nrec=int(info[0]),
                consumed=int(info[1]),
                phred64=int(info[2]),
                orig_seq=int(info[3]),
                orig_qual=int(info[4]),
            )

    def decode_streams(self, streams, nrec: int, phred64: int) -> bytes:
        arrs = [_as_u8(s) for s in streams]
        inp = (C.c_void_p * 6)(*[(a.ctypes.data if a.size else None) for a in arrs])
        lens = (C.c_size_t * 6)(*[a.size for a in arrs])
        cap = sum(a.size for a in arrs) * 5 + 16 * nrec + 64
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m = C.c_size_t(0)
            rc = self.lib.L.fqz_decode_streams(self.h, inp, lens, nrec, phred64, _ptr(out), cap, C.byref(m))
            if rc == FQZ_E_NOSPACE:
                cap = m.value
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    def zstd_compress(self, data, policy: int = 0) -> bytes:
        a = _as_u8(data)
        cap = a.size + a.size // 64 + 4096
        out = np.empty(cap, dtype=np.uint8)
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_zstd_compress(self.h, _ptr(a), a.size, policy, _ptr(out), cap, C.byref(m)))
        return out[: m.value].tobytes()

    def zstd_decompress(self, data, cap: int | None = None) -> bytes:
        a = _as_u8(data)
        cap = cap or max(1 << 16, a.size * 8)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m = C.c_size_t(0)
            rc = self.lib.L.fqz_zstd_decompress(self.h, _ptr(a), a.size, _ptr(out), cap, C.byref(m))
            if rc == FQZ_E_NOSPACE:
                cap = max(m.value, cap * 2)
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    # ---- whole buffer, host memory -------------------------------------------------------------
    def compress(self, fastq, block_size: int = 0) -> bytes:
        a = _as_u8(fastq)
        cap = int(self.lib.L.fqz_compress_bound(a.size))
        out = np.empty(cap, dtype=np.uint8)
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_compress(self.h, _ptr(a), a.size, block_size, _ptr(out), cap, C.byref(m)))
        return out[: m.value].tobytes()

    def compress_shard(self, fastq, phred64: int = -1, file_header: bool = True, block_size: int = 0):
        """One shard of a block-sharded file -> (bytes, phred64 flag in force); see fqz_compress_shard."""
        a = _as_u8(fastq)
        cap = int(self.lib.L.fqz_compress_bound(a.size))
        out = np.empty(cap, dtype=np.uint8)
        m, used = C.c_size_t(0), C.c_int(0)
        self._check(
            self.lib.L.fqz_compress_shard(self.h, _ptr(a), a.size, block_size, phred64, 1 if file_header else 0, _ptr(out), cap, C.byref(m), C.byref(used))
        )
        return out[: m.value].tobytes(), used.value

    def compress_into(self, fastq: np.ndarray, out: np.ndarray, block_size: int = 0) -> int:
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_compress(self.h, _ptr(fastq), fastq.size, block_size, _ptr(out), out.size, C.byref(m)))
        return m.value

    def decompress(self, fqz, cap: int | None = None) -> bytes:
        a = _as_u8(fqz)
        cap = cap or max(1 << 16, a.size * 6)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m = C.c_size_t(0)
            rc = self.lib.L.fqz_decompress(self.h, _ptr(a), a.size, _ptr(out), cap, C.byref(m))
            if rc == FQZ_E_NOSPACE:
                cap = max(m.value, cap * 2)
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    def block_index(self, fqz) -> list:
        return self.lib.block_index(fqz)

    def decompress_blocks(self, fqz, first_block: int, num_blocks: int, cap: int | None = None) -> bytes:
        """FASTQ of blocks [first_block, first_block + num_blocks) alone (random access through the block index)."""
        a = _as_u8(fqz)
        cap = cap or max(1 << 16, a.size * 6)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m = C.c_size_t(0)
            rc = self.lib.L.fqz_decompress_blocks(self.h, _ptr(a), a.size, first_block, num_blocks, _ptr(out), cap, C.byref(m))
            if rc == FQZ_E_NOSPACE:
                cap = max(m.value, cap * 2)
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    # ---- gzip input (cmd/fqpack/main.go:142-174) -------------------------------------------------
    def is_gzip(self, data) -> bool:
        a = _as_u8(data)
        return bool(self.lib.L.fqz_is_gzip(_ptr(a), a.size))

    def gunzip(self, gz, cap: int | None = None) -> bytes:
        """gzip file (any number of members) -> its text, inflated on the GPU."""
        a = _as_u8(gz)
        cap = cap or max(1 << 16, a.size * 5)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m = C.c_size_t(0)
            rc = self.lib.L.fqz_gunzip(self.h, _ptr(a), a.size, _ptr(out), cap, C.byref(m))
            if rc == FQZ_E_NOSPACE:
                cap = m.value
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    def gunzip_stats(self) -> dict:
        v = (C.c_uint64 * 5)()
        self._check(self.lib.L.fqz_gunzip_stats(self.h, C.byref(v)))
        return dict(chunks=int(v[0]), parallel=int(v[1]), dropped=int(v[2]), members=int(v[3]), decoded_twice=int(v[4]))

    def gunzip_device(self, d_in: int, n: int, d_out: int, out_cap: int) -> int:
        """Device buffers; returns the text length (raises FqzError(-15) with .needed set when out_cap is short)."""
        m = C.c_size_t(0)
        rc = self.lib.L.fqz_gunzip_device(self.h, C.c_void_p(d_in), n, C.c_void_p(d_out), out_cap, C.byref(m))
        if rc == FQZ_E_NOSPACE:
            e = FqzError(rc, self.lib.strerror(rc))
            e.needed = m.value
            raise e
        self._check(rc)
        return m.value

    def compress_gz(self, gz, block_size: int = 0) -> bytes:
        """gzipped FASTQ -> .fqz: the compressed bytes are uploaded, inflated and coded on the GPU."""
        a = _as_u8(gz)
        cap = int(self.lib.L.fqz_compress_bound(a.size * 4))
        while True:
            out = np.empty(cap, dtype=np.uint8)
            m, f = C.c_size_t(0), C.c_size_t(0)
            rc = self.lib.L.fqz_compress_gz(self.h, _ptr(a), a.size, block_size, _ptr(out), cap, C.byref(m), C.byref(f))
            if rc == FQZ_E_NOSPACE:
                cap = m.value
                continue
            self._check(rc)
            return out[: m.value].tobytes()

    def info(self, fqz) -> dict:
        """`fqpack info`: version, flags, blocks, records, per-stream compressed sizes (header walk only)."""
        a = _as_u8(fqz)
        fi = _FileInfo()
        self._check(self.lib.L.fqz_info(self.h, _ptr(a), a.size, C.byref(fi)))
        return dict(version=fi.version, flags=fi.flags, phred64=bool(fi.flags & 2), header_block_size=fi.header_block_size, blocks=int(fi.blocks),
                    records=int(fi.records), compressed=[int(x) for x in fi.compressed], original_seq=int(fi.original_seq),
                    original_qual=int(fi.original_qual))

    def check(self, fqz):
        """`fqpack check`: full decode on the GPU without output -> (records, FASTQ bytes); raises FqzError on a bad file."""
        a = _as_u8(fqz)
        r, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.L.fqz_check(self.h, _ptr(a), a.size, C.byref(r), C.byref(b)))
        return r.value, b.value

    def decompress_into(self, fqz: np.ndarray, out: np.ndarray) -> int:
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_decompress(self.h, _ptr(fqz), fqz.size, _ptr(out), out.size, C.byref(m)))
        return m.value

    # ---- whole buffer, device memory (raw device pointers, e.g. torch.Tensor.data_ptr()) --------
    def compress_device(self, d_in: int, n: int, d_out: int, out_cap: int, block_size: int = 0) -> int:
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_compress_device(self.h, C.c_void_p(d_in), n, block_size, C.c_void_p(d_out), out_cap, C.byref(m)))
        return m.value

    def compress_shard_device(self, d_in: int, n: int, d_out: int, out_cap: int, phred64: int = -1, file_header: bool = True, block_size: int = 0):
        """fqz_compress_shard on device memory -> (bytes written, phred64 flag in force)."""
        m, used = C.c_size_t(0), C.c_int(0)
        self._check(
            self.lib.L.fqz_compress_shard_device(self.h, C.c_void_p(d_in), n, block_size, phred64, 1 if file_header else 0, C.c_void_p(d_out), out_cap,
                                                 C.byref(m), C.byref(used))
        )
        return m.value, used.value

    def count_lines_device(self, d_in: int, n: int) -> int:
        v = C.c_uint64(0)
        self._check(self.lib.L.fqz_count_lines_device(self.h, C.c_void_p(d_in), n, C.byref(v)))
        return v.value

    def find_line_end_device(self, d_in: int, n: int, k: int) -> int:
        """Offset of the k-th (1-based) newline of the device buffer."""
        v = C.c_uint64(0)
        self._check(self.lib.L.fqz_find_line_end_device(self.h, C.c_void_p(d_in), n, k, C.byref(v)))
        return v.value

    def decompress_device(self, d_in: int, n: int, d_out: int, out_cap: int) -> int:
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_decompress_device(self.h, C.c_void_p(d_in), n, C.c_void_p(d_out), out_cap, C.byref(m)))
        return m.value

    def synth_device(self, kind: int, seed: int, first: int, count: int, d_out: int, out_cap: int) -> int:
        m = C.c_size_t(0)
        self._check(self.lib.L.fqz_synth_device(self.h, kind, seed, first, count, C.c_void_p(d_out), out_cap, C.byref(m)))
        return m.value

    # ---- streaming ----------------------------------------------------------------------------
    def compress_stream(self, block_size: int = 0):
        return _CStream(self, block_size)

    def decompress_stream(self):
        return _DStream(self)

    # ---- tuning -------------------------------------------------------------------------------
    OPT_WINDOW_BYTES, OPT_HOST_WINDOW_BYTES, OPT_RECORD_MATCH, OPT_FRONTEND, OPT_HUF_KERNELS, OPT_SERIAL_ENTROPY = 1, 2, 3, 4, 5, 6
    OPT_GZ_CHUNK_BYTES = 7

    def set_option(self, key: int, value: int):
        self._check(self.lib.L.fqz_set_option(self.h, key, value))

    # ---- measurement --------------------------------------------------------------------------
    def stream_handle(self) -> int:
        """cudaStream_t of the context (for torch.cuda.ExternalStream / CUDA-event timing)."""
        return int(self.lib.L.fqz_get_stream(self.h) or 0)

    def stats_reset(self):
        self.lib.L.fqz_stats_reset(self.h)

    def profile(self, on: bool):
        self.lib.L.fqz_profile_enable(self.h, 1 if on else 0)

    def stats(self) -> dict:
        st = _Stats()
        self._check(self.lib.L.fqz_get_stats(self.h, C.byref(st)))
        stages = {}
        for i in range(st.n_stages):
            if st.stage_launches[i] or st.stage_ms[i]:
                stages[st.stage_name[i].decode()] = dict(
                    ms=st.stage_ms[i], launches=int(st.stage_launches[i]), bytes=int(st.stage_bytes[i])
                )
        return dict(launches=int(st.launches), stages=stages)


class _CStream:
    def __init__(self, ctx: FqzContext, block_size: int):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx._check(ctx.lib.L.fqz_compress_begin(ctx.h, block_size, C.byref(self.h)))

    def feed(self, data, is_last: bool, out: np.ndarray):
        a = _as_u8(data)
        m, used = C.c_size_t(0), C.c_size_t(0)
        rc = self.ctx.lib.L.fqz_compress_feed(self.h, _ptr(a), a.size, 1 if is_last else 0, _ptr(out), out.size, C.byref(m), C.byref(used))
        self.ctx._check(rc)
        return m.value, used.value

    def close(self):
        if self.h:
            self.ctx.lib.L.fqz_compress_end(self.h)
            self.h = C.c_void_p()


class _DStream:
    def __init__(self, ctx: FqzContext):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx._check(ctx.lib.L.fqz_decompress_begin(ctx.h, C.byref(self.h)))

    def feed(self, data, is_last: bool, out: np.ndarray):
        a = _as_u8(data)
        m, used = C.c_size_t(0), C.c_size_t(0)
        rc = self.ctx.lib.L.fqz_decompress_feed(self.h, _ptr(a), a.size, 1 if is_last else 0, _ptr(out), out.size, C.byref(m), C.byref(used))
        if rc == FQZ_E_NOSPACE:
            return -m.value, 0
        self.ctx._check(rc)
        return m.value, used.value

    def close(self):
        if self.h:
            self.ctx.lib.L.fqz_decompress_end(self.h)
            self.h = C.c_void_p()
