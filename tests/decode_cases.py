"""Decompress-side parity checks shared by the CPU-emulation tests (tests/test_emu_decode.py) and
the GPU tests (tests/test_gpu_decompress.py).  `ctx` is a fastqpacker_b200 FqzContext, `oracle`
the CPU oracle module (libzstd stands in for klauspost/compress)."""
import random
import struct

import pytest


def _zstd_data(name, scale=1.0):
    rnd = random.Random(sum(name.encode()) * 7919)
    n = lambda k: max(1, int(k * scale))
    if name == "tiny":
        return b"abc"
    if name == "one":
        return b"x"
    if name == "zeros":
        return bytes(n(300000))
    if name == "text":
        return b"hello world, hello zstd " * n(9000)
    if name == "quality_like":
        return bytes(rnd.choice([0, 0, 0, 0, 0, 0, 1, 255, 2, 254, 3]) for _ in range(n(400000)))
    if name == "random":
        return bytes(rnd.randrange(256) for _ in range(n(200000)))
    if name == "headers_like":
        return b"".join(b"\x21\x00ERR532393.%d HWI-ST571:218:C2DACACXX:5:1101:%d:%d/1" % (i, 1000 + i * 7 % 20000, 2000 + i) for i in range(n(12000)))
    if name == "runs":
        return b"".join(bytes([i & 255]) * (i % 700) for i in range(n(900)))
    if name == "mixed":
        return b"".join(bytes([rnd.randrange(4)]) * rnd.randrange(1, 40) + b"@ERR%d/1" % i for i in range(n(20000)))
    if name == "lengths_like":
        return struct.pack("<I", 150) * n(100000)
    if name == "small_alphabet":
        return bytes(rnd.randrange(7) for _ in range(n(140000)))
    raise KeyError(name)


ZSTD_DATA = ["tiny", "one", "zeros", "text", "quality_like", "random", "headers_like", "runs", "mixed", "lengths_like", "small_alphabet"]


def check_zstd_libzstd_frames(ctx, oracle, name, level, scale=1.0):
    """Frames written by libzstd (multi-block, FSE/repeat tables, treeless literals, repeat offsets,
    cross-block matches) must decode bit-exact on the device."""
    data = _zstd_data(name, scale)
    z = oracle.zstd_compress(data, level)
    assert ctx.zstd_decompress(z) == data
    # concatenated frames decode back to back (klauspost DecodeAll / libzstd behaviour)
    z2 = z + oracle.zstd_compress(b"second frame " * 10, level)
    assert ctx.zstd_decompress(z2) == data + b"second frame " * 10


def _far_match_data(nbytes, seed):
    """header-like lines in which every line copies a field of a line up to ~3 MiB back: matches that cross many
    128 KiB blocks inside a multi-megabyte window"""
    rnd = random.Random(seed)
    lines, out, size = [], bytearray(), 0
    while size < nbytes:
        if lines and rnd.random() < 0.6:
            src = lines[rnd.randrange(max(0, len(lines) - 60000), len(lines))]
            ln = src[: rnd.randrange(8, len(src))] + b":%d:%d" % (rnd.randrange(99999), rnd.randrange(99999))
        else:
            ln = b"@INSTR-%d:%d:FC%d:%d:%d:%d:%d" % tuple(rnd.randrange(1, 10 ** k) for k in (3, 3, 5, 1, 4, 5, 5))
        lines.append(ln)
        out += ln + b"\n"
        size += len(ln) + 1
    return bytes(out[:nbytes])


# (window log, content size flag, checksum flag, level): frame shapes other zstd encoders write and libzstd's defaults
# do not: no frame content size (a streaming encoder), no checksum, explicit 1 MiB .. 8 MiB
# windows, Single_Segment frames (content size known and not larger than the window: libzstd then drops the
# window descriptor)
FRAME_SHAPES = [(22, 0, 1, 1), (22, 1, 0, 3), (23, 0, 0, 1), (20, 0, 1, 5), (24, 1, 1, 1)]


def check_zstd_frame_shapes(ctx, oracle, nbytes, shapes=FRAME_SHAPES):
    data = _far_match_data(nbytes, 77)
    for wlog, cs, ck, level in shapes:
        z = oracle.zstd_compress_adv(data, level, wlog, cs, ck)
        desc = z[4]
        if not cs:
            assert (desc >> 6) == 0 and not (desc & 0x20), "frame carries a content size"
        if cs and (1 << wlog) >= nbytes:
            assert desc & 0x20, "expected a Single_Segment frame"
        assert ctx.zstd_decompress(z) == data, (wlog, cs, ck, level)
    # the same frames back to back, as DecodeAll sees a stream written in pieces
    a, b = oracle.zstd_compress_adv(data[: nbytes // 2], 1, 22, 0, 1), oracle.zstd_compress_adv(data[nbytes // 2 :], 1, 22, 0, 0)
    assert ctx.zstd_decompress(a + b) == data


def check_zstd_round_trip(ctx, oracle, name, policy, scale=1.0):
    data = _zstd_data(name, scale)
    z = ctx.zstd_compress(data, policy)
    assert oracle.zstd_decompress(z) == data
    assert ctx.zstd_decompress(z) == data


ENT_SIZES = [1, 2, 31, 1023, 1024, 1025, 4099, 16383, 16384, 16385, 2 * 16384 + 500, 2 * 16384 + 1024, 131071, 131072, 131073, 131072 + 1023,
             131072 + 1024, 3 * 131072 + 77]


def check_zstd_ent_sizes(ctx, oracle, n):
    """literals-only policy at block / frame boundaries (16 KiB blocks, 128 KiB frames, raw tails),
    for compressible, incompressible and single-symbol content."""
    rnd = random.Random(n)
    skew = bytes(rnd.choice(b"\x00\x00\x00\x00\x00\x01\xff\x02\xfe\x07") for _ in range(n))
    flat = bytes(rnd.randrange(256) for _ in range(n))
    wide = bytes(min(255, int(rnd.expovariate(0.05))) for _ in range(n))  # many symbols, long codes
    for data in (skew, flat, wide, b"\x07" * n):
        z = ctx.zstd_compress(data, 1)
        assert oracle.zstd_decompress(z) == data
        assert ctx.zstd_decompress(z) == data
    if n >= 4096:
        assert len(ctx.zstd_compress(skew, 1)) < 0.45 * n
        nfr = (n + 131071) // 131072
        index = 20 + 8 * nfr if nfr >= 4 else 0  # skippable frame index in front of streams of >= 4 frames
        assert len(ctx.zstd_compress(flat, 1)) <= n + 17 * nfr + index


def check_zstd_index(ctx, oracle, policies=(0, 1), n=5 * 131072 + 333):
    """Streams cut into >= 4 frames start with a skippable index frame (frame sizes).  Any decoder
    skips it (libzstd does here); the device decoder uses it only as far as it proves true: a
    tampered or foreign index must not change the result."""
    import struct

    rnd = random.Random(5)
    data = bytes(rnd.choice(b"ACGTTTGA\x00\x01") for _ in range(n))
    for policy in policies:
        z = ctx.zstd_compress(data, policy)
        magic, psize, sig, nfr, fsz = struct.unpack_from("<IIIII", z, 0)
        assert magic == 0x184D2A5E and sig == 0x495A5146 and psize == 12 + 8 * nfr and nfr >= 4
        assert fsz == (131072 if policy == 1 else 65536)
        ent = list(struct.unpack_from("<%dI" % (2 * nfr), z, 20))
        sizes, hints = ent[0::2], ent[1::2]
        body = 20 + 8 * nfr
        assert sum(sizes) == len(z) - body and not any(hints)
        assert oracle.zstd_decompress(z) == data
        assert ctx.zstd_decompress(z) == data

        def rebuild(sz):
            return z[:20] + struct.pack("<%dI" % (2 * nfr), *[v for pair in zip(sz, hints) for v in pair]) + z[body:]

        # entries swapped (the sum still matches), entries wrong, foreign payload behind the same magic
        sw = sizes[:]
        sw[0], sw[1] = sw[1] + 1, sw[0] - 1
        for bad in (sw, [v + 1 for v in sizes]):
            assert ctx.zstd_decompress(rebuild(bad)) == data
        foreign = struct.pack("<II", 0x184D2A5E, 12) + b"hello world!" + z[body:]
        assert ctx.zstd_decompress(foreign) == data
        assert ctx.zstd_decompress(z[body:]) == data  # and no index at all


def check_back_end(ctx, oracle, text):
    """six pre-entropy streams -> FASTQ == NumRecords x blockReader.writeRecord."""
    enc = oracle.encode_streams(text)
    want = oracle.decode_streams(enc["streams"], enc["nrec"], enc["phred64"])
    got = ctx.decode_streams(enc["streams"], enc["nrec"], enc["phred64"])
    assert got == want
    if enc["nrec"]:
        # v1 rule: an empty plus stream yields bare '+' lines (compress.go:995-999)
        s = list(enc["streams"])
        s[3] = b""
        assert ctx.decode_streams(s, enc["nrec"], enc["phred64"]) == oracle.decode_streams(s, enc["nrec"], enc["phred64"])


def check_decompress_reference_written(ctx, oracle, text):
    """reference-shaped .fqz (oracle container + libzstd frames) decodes bit-exact on the device."""
    for level in (1, 5):
        fqz = oracle.compress(text, level=level)
        assert ctx.decompress(fqz) == oracle.decompress(fqz)


def check_round_trip(ctx, oracle, text):
    fqz = ctx.compress(text)
    want = oracle.decompress(oracle.compress(text))  # the reference's own (lossy-normalising) round trip
    assert oracle.decompress(fqz) == want
    assert ctx.decompress(fqz) == want


def check_v1_file(ctx, oracle):
    """v1 container: 32-byte block headers, five streams, plus lines rebuilt as '+'
    (compress_test.go:502-592)."""
    from tests.fastq_cases import GOOD_CASES

    for name in ("three", "plus_payload", "rand_small"):
        text = GOOD_CASES[name]
        v1 = oracle.compress(text, version=1)
        assert v1[4] == 1
        want = oracle.decompress(v1)
        assert ctx.decompress(v1) == want


_TRUNC_CODES = {0: -11, 1: -12, 2: -9, 3: -10, 4: -14, 5: -13}


def check_decode_errors(ctx, oracle):
    """'truncated ... data' errors of blockReader (compress.go:979-1057): same error as the oracle,
    whichever stream runs short first in the reference's per-record order."""
    from fastqpacker_b200._binding import FqzError
    from tests.fastq_cases import GOOD_CASES

    text = GOOD_CASES["rand_plus"] + GOOD_CASES["nbases"].replace(b"IIII!!!!IIII", b"hhhhhhhhhhhh")
    enc = oracle.encode_streams(text, phred64=1)
    streams, nrec = enc["streams"], enc["nrec"]
    assert ctx.decode_streams(streams, nrec, 1) == oracle.decode_streams(streams, nrec, 1)
    rnd = random.Random(5)
    for which in range(6):
        for cut in (1, 2, 3, 7, len(streams[which]) // 2, len(streams[which]) - 1):
            s = list(streams)
            s[which] = s[which][: len(s[which]) - cut]
            if which == 3 and not s[which]:
                continue  # an empty plus stream is the v1 rule, not an error
            with pytest.raises(oracle.OracleError) as oe:
                oracle.decode_streams(s, nrec, 1)
            with pytest.raises(FqzError) as ge:
                ctx.decode_streams(s, nrec, 1)
            assert ge.value.code == oe.value.code, (which, cut)
    # two streams short at once: the reference's per-record order decides
    for _ in range(12):
        s = list(streams)
        for which in rnd.sample(range(6), 2):
            s[which] = s[which][: rnd.randrange(1, len(s[which]))]
        if not s[3]:
            continue
        with pytest.raises(oracle.OracleError) as oe:
            oracle.decode_streams(s, nrec, 1)
        with pytest.raises(FqzError) as ge:
            ctx.decode_streams(s, nrec, 1)
        assert ge.value.code == oe.value.code
    # more records claimed than present
    with pytest.raises(FqzError) as ge:
        ctx.decode_streams(streams, nrec + 5, 1)
    assert ge.value.code == -13
    # N position beyond the read (the reference panics; both sides report -17)
    s = list(streams)
    npos = bytearray(s[4])
    k = len(npos) - 2  # last position of the last record
    npos[k : k + 2] = struct.pack("<H", 5000)
    s[4] = bytes(npos)
    with pytest.raises(FqzError) as ge:
        ctx.decode_streams(s, nrec, 1)
    assert ge.value.code == -17


def check_file_errors(ctx, oracle):
    from fastqpacker_b200._binding import FqzError
    from tests.fastq_cases import GOOD_CASES

    fqz = oracle.compress(GOOD_CASES["rand_small"])
    cases = {
        "bad_magic": (b"FQX\x00" + fqz[4:], -5),
        "short_magic": (fqz[:3], -7),
        "short_header": (fqz[:8], -7),
        "bad_version": (fqz[:4] + b"\x03" + fqz[5:], -6),
        "cut_block_header": (fqz[:30], -7),
        "cut_payload": (fqz[:-9], -7),
    }
    for name, (blob, code) in cases.items():
        with pytest.raises(oracle.OracleError) as oe:
            oracle.decompress(blob)
        assert oe.value.code == code, name
        with pytest.raises(FqzError) as ge:
            ctx.decompress(blob)
        assert ge.value.code == code, name
        with pytest.raises(FqzError) as ce:  # `fqpack check` (fqz_check) rejects the same files the same way
            ctx.check(blob)
        assert ce.value.code == code, name
        with pytest.raises(FqzError) as ie:  # ... and so does the header walk of `fqpack info`
            ctx.info(blob)
        assert ie.value.code == code, name
    # corrupt frames: flip one byte inside each stream's payload in turn
    sizes = struct.unpack("<9I", fqz[10:46])[1:7]
    pos = 46
    for i, sz in enumerate(sizes):
        bad = bytearray(fqz)
        bad[pos + sz // 2] ^= 0x5A
        pos += sz
        try:
            want = oracle.decompress(bytes(bad))
        except oracle.OracleError as e:
            want = e.code
        try:
            got = ctx.decompress(bytes(bad))
        except FqzError as e:
            got = e.code
        if isinstance(want, int):
            assert isinstance(got, int) and got < 0, (i, want, got)  # both reject (the first check that trips may differ)
            with pytest.raises(FqzError):
                ctx.check(bytes(bad))
        else:
            assert got == want
    assert ctx.decompress(fqz[:10]) == b""  # header-only file (compress_test.go:160-173)
    assert ctx.check(fqz[:10]) == (0, 0)
    # info / check of good files, reference-written (v2, v1) and GPU-written
    text = GOOD_CASES["rand_small"]
    nrec = text.count(b"\n") // 4
    for blob, ver in ((fqz, 2), (oracle.compress(text, version=1), 1), (ctx.compress(text), 2)):
        fi = ctx.info(blob)
        hdr = struct.unpack("<8I" if ver == 1 else "<9I", blob[10 : 10 + (32 if ver == 1 else 36)])
        assert (fi["version"], fi["blocks"], fi["records"], fi["phred64"], fi["header_block_size"]) == (ver, 1, nrec, False, 100000)
        want_sizes = list(hdr[1:6]) if ver == 2 else [hdr[1], hdr[2], hdr[3], 0, hdr[4]]
        assert fi["compressed"][:5] == want_sizes and fi["compressed"][5] == hdr[6 if ver == 2 else 5]
        assert fi["original_seq"] == fi["original_qual"] == sum(len(l) for l in text.split(b"\n")[1::4])
        assert ctx.check(blob) == (nrec, len(oracle.decompress(blob)))


def check_item_hints(ctx, oracle, nrec=3000):
    """The index frames of the header / plus / N-position streams carry, per 16 KiB frame, where the
    first item starts and how many items start there; the device walks the frames in parallel and
    proves every hand-over.  The hints of a GPU-written file must describe the decoded streams
    exactly, and wrong hints must only cost the parallel walk, never the result."""
    import struct

    from tests import synth

    for kind in (0, 1):
        text = synth.fastq(kind, 21, 0, nrec)
        fqz = ctx.compress(text)
        assert oracle.decompress(fqz) == text
        assert ctx.decompress(fqz) == text
        hdr = struct.unpack_from("<9I", fqz, 10)
        streams = oracle.encode_streams(text)["streams"]
        pos = 10 + 36 + hdr[1] + hdr[2]
        tampered = 0
        for a in (2, 3, 4):  # headers, plus, npos
            z = fqz[pos : pos + hdr[1 + a]]
            raw = streams[a]
            if z[:4] == b"\x5e\x2a\x4d\x18":
                magic, psize, sig, nfr, fsz = struct.unpack_from("<IIIII", z, 0)
                assert sig == 0x495A5146 and fsz == 16384 and nfr == (len(raw) + fsz - 1) // fsz
                # item starts of the decoded stream
                starts, q = [], 0
                while q < len(raw):
                    starts.append(q)
                    n = raw[q] | (raw[q + 1] << 8)
                    q += 2 + (2 * n if a == 4 else n)
                for k in range(nfr):
                    first, cnt = struct.unpack_from("<HH", z, 20 + 8 * k + 4)
                    mine = [v for v in starts if k * fsz <= v < (k + 1) * fsz]
                    assert cnt == len(mine) and (not mine or first == mine[0] - k * fsz)
                for k, delta in ((0, (0, 1)), (1, (2, 0)), (nfr - 1, (0, -1))):
                    at = pos + 20 + 8 * k + 4
                    first, cnt = struct.unpack_from("<HH", fqz, at)
                    bad = fqz[:at] + struct.pack("<HH", (first + delta[0]) & 0xFFFF, (cnt + delta[1]) & 0xFFFF) + fqz[at + 4 :]
                    assert ctx.decompress(bad) == text
                    tampered += 1
            pos += hdr[1 + a]
        assert tampered >= 3  # the header stream of 3000 records spans >= 4 frames


def check_streaming(ctx, oracle, nrec=300, chunk=None):
    """Seam B feed() calls: windows cut at arbitrary byte positions give the same bytes as the
    whole-buffer calls."""
    import numpy as np

    from tests import synth

    text = synth.fastq(1, 11, 0, nrec)
    want_fqz = ctx.compress(text)
    out = np.empty(len(text) + (1 << 16), dtype=np.uint8)
    cs = ctx.compress_stream()
    got = b""
    pos = 0
    chunk = chunk or max(64, len(text) // 3)
    window = b""
    while True:
        window += text[pos : pos + chunk]
        pos += chunk
        last = pos >= len(text)
        try:
            m, used = cs.feed(window, last, out)
        except Exception as e:  # FQZ_E_NEED_MORE: no complete block yet
            assert getattr(e, "code", 0) == -35 and not last
            continue
        got += out[:m].tobytes()
        window = window[used:]
        if last:
            assert not window
            break
    cs.close()
    assert got == want_fqz
    # decompress side, fed in pieces that split headers and payloads
    want = oracle.decompress(want_fqz)
    ds = ctx.decompress_stream()
    big = np.empty(len(want) + 4096, dtype=np.uint8)
    res = b""
    pos = 0
    window = b""
    step = max(7, len(want_fqz) // 5)
    while True:
        window += want_fqz[pos : pos + step]
        pos += step
        last = pos >= len(want_fqz)
        try:
            m, used = ds.feed(window, last, big)
        except Exception as e:
            assert getattr(e, "code", 0) == -35 and not last
            continue
        assert m >= 0
        res += big[:m].tobytes()
        window = window[used:]
        if last:
            assert not window
            break
    ds.close()
    assert res == want


def multi_block_file(oracle, version=2, gpu_ctx=None):
    """A container of several small blocks (any NumRecords <= 100 000 per block is a valid file: the reader takes
    the count from each block header, compress.go:738-758) and the text of every block: reference-shaped blocks
    from the oracle, and one written by the GPU coder when a context is given."""
    from fastqpacker_b200.sharding import merge_compressed
    from tests.fastq_cases import GOOD_CASES, rand_fastq

    texts = [GOOD_CASES["rand_small"], GOOD_CASES["three"], rand_fastq(700, 21, lmin=1, lmax=40), rand_fastq(300, 23, plus_payload=True),
             GOOD_CASES["nbases"], rand_fastq(150, 22, lmin=100, lmax=400, n_rate=0.1)]
    # the text a block decodes to (the codec is lossy where the reference is: N bases keep quality 0 only, v1 drops
    # plus-line payloads)
    texts = [oracle.decompress(oracle.compress(t, version=version)) for t in texts]
    parts = [oracle.compress(t, version=version) for t in texts]
    if gpu_ctx is not None and version == 2:
        raw = rand_fastq(250, 24, lmin=30, lmax=90)
        texts.append(oracle.decompress(oracle.compress(raw)))
        parts.append(gpu_ctx.compress(raw))
        assert oracle.decompress(parts[-1]) == texts[-1]
    assert all(p[9] == parts[0][9] for p in parts)  # the Phred flag is file-global (compress.go:146-164): blocks of one file share it
    return merge_compressed(parts), texts


def check_block_index(lib, oracle):
    """fqz_block_index against the Python header walk (sharding.walk_container) and the oracle; needs no GPU."""
    from fastqpacker_b200._binding import FqzError
    from fastqpacker_b200.sharding import walk_container

    for version in (2, 1):
        fqz, texts = multi_block_file(oracle, version)
        assert oracle.decompress(fqz) == b"".join(texts)
        idx = lib.block_index(fqz)
        _, _, want = walk_container(fqz)
        assert [(e["offset"], e["size"], e["records"]) for e in idx] == [(b.offset, b.size, b.records) for b in want]
        first = 0
        for e, t in zip(idx, texts):
            assert e["first_record"] == first and e["records"] == t.count(b"\n") // 4
            assert e["original_seq"] == e["original_qual"] == sum(len(l) for l in t.split(b"\n")[1::4])
            first += e["records"]
        assert idx[-1]["offset"] + idx[-1]["size"] == len(fqz)
        assert lib.block_index(fqz[:10]) == []
        for blob, code in ((b"FQX\x00" + fqz[4:], -5), (fqz[:3], -7), (fqz[:8], -7), (fqz[:4] + b"\x07" + fqz[5:], -6),
                           (fqz[: idx[2]["offset"] + 11], -7), (fqz[:-1], -7)):
            with pytest.raises(FqzError) as e:
                lib.block_index(blob)
            assert e.value.code == code
    return fqz


def check_decompress_blocks(ctx, oracle, dense=True):
    """Random access: every block range of a multi-block file decodes to exactly those blocks' text."""
    from fastqpacker_b200._binding import FqzError

    for version in (2, 1):
        fqz, texts = multi_block_file(oracle, version, gpu_ctx=ctx)
        nb = len(texts)
        assert len(ctx.block_index(fqz)) == nb
        assert ctx.decompress(fqz) == b"".join(texts)
        for first in range(nb) if dense else (0, 2, nb - 1):
            for count in (1, 2, nb - first) if dense else (1, nb - first):
                if first + count <= nb:
                    assert ctx.decompress_blocks(fqz, first, count) == b"".join(texts[first : first + count]), (version, first, count)
        assert ctx.decompress_blocks(fqz, nb, 0) == b"" and ctx.decompress_blocks(fqz, 2, 0) == b""
        for first, count in ((0, nb + 1), (nb, 1), (nb + 5, 0)):
            with pytest.raises(FqzError) as e:
                ctx.decompress_blocks(fqz, first, count)
            assert e.value.code == -34
        # a cut file: the whole blocks in front of the cut still decode
        idx = ctx.block_index(fqz)
        cutf = fqz[: idx[4]["offset"] + 20]
        assert ctx.decompress_blocks(cutf, 1, 3) == b"".join(texts[1:4])
        for first, count in ((3, 2), (4, 1), (5, 1)):
            with pytest.raises(FqzError) as e:
                ctx.decompress_blocks(cutf, first, count)
            assert e.value.code == -7
        # a damaged block is only met by the ranges that hold it, and the error names its index in the file
        idx = ctx.block_index(fqz)
        bad = bytearray(fqz)
        hsz = 32 if version == 1 else 36
        bad[idx[3]["offset"] + hsz + 5] ^= 0x40
        bad = bytes(bad)
        assert ctx.decompress_blocks(bad, 0, 3) == b"".join(texts[:3])
        assert ctx.decompress_blocks(bad, 4, nb - 4) == b"".join(texts[4:])
        with pytest.raises(FqzError) as e:
            ctx.decompress_blocks(bad, 2, 3)
        assert e.value.code == -8 and "block 3" in str(e.value), str(e.value)


def check_block_header_fields(ctx, oracle, nrec=1500, quick=False):
    """Block headers that disagree with their streams (ROADMAP.md PR-005, `compress.go:738-758,944-1078`): NumRecords below
    what the streams hold decodes exactly that many records and ignores the rest, above it is `truncated ... data` at the
    first missing record however large the claim (no allocation follows the claim), and OriginalSeqSize / OriginalQualSize
    are not used by the reader at all.  Reference-shaped and GPU-written blocks, same verdict as the oracle."""
    from fastqpacker_b200._binding import FqzError
    from tests.fastq_cases import rand_fastq

    text = rand_fastq(nrec, 31, lmin=20, lmax=120, plus_payload=True)
    for fqz in (oracle.compress(text), ctx.compress(text)):
        hdr = list(struct.unpack("<9I", fqz[10:46]))
        assert hdr[0] == nrec

        def verdict(fn, blob):
            try:
                return fn(blob)
            except (FqzError, oracle.OracleError) as e:
                return e.code

        fields = ((0, (0, 1, 2, 31, 32, 33, 150, 151, nrec // 2, nrec - 1, nrec + 1, 100_001, 0x7FFFFFFF, 0xFFFFFFFF)),
                  (7, (0, 1, hdr[7] + 1, 0xFFFFFFFF)), (8, (0, hdr[8] - 1, 0xFFFFFFFF)))
        if quick:  # the CPU emulation of the kernels is slow
            fields = ((0, (1, 33, nrec - 1, nrec + 1, 0xFFFFFFFF)), (7, (0xFFFFFFFF,)), (8, (0,)))
        for field, values in fields:
            for v in values:
                h = list(hdr)
                h[field] = v
                blob = fqz[:10] + struct.pack("<9I", *h) + fqz[46:]
                want = verdict(oracle.decompress, blob)
                assert verdict(ctx.decompress, blob) == want, (field, v)
                chk = verdict(ctx.check, blob)
                assert chk == want if isinstance(want, int) else chk == (v if field == 0 else nrec, len(want)), (field, v)


def fuzz_fqz(ctx, oracle, seed: int):
    """A small .fqz (reference-shaped or GPU-written, v2 or v1, one or several blocks) with random damage: flipped bits,
    bytes replaced / dropped / inserted, the tail cut — anywhere: file header, block headers, frame headers, entropy
    tables, bit streams, checksums."""
    from fastqpacker_b200.sharding import merge_compressed
    from tests.fastq_cases import rand_fastq

    rnd = random.Random(0xF02 + seed)
    version = rnd.choice((2, 2, 1))
    parts = []
    for _ in range(rnd.choice((1, 1, 2, 3))):
        text = rand_fastq(rnd.randint(1, 120), rnd.randrange(1 << 20), lmin=1, lmax=rnd.choice((8, 60, 300)), plus_payload=rnd.random() < 0.3,
                          n_rate=rnd.choice((0.0, 0.05)))
        if rnd.random() < 0.3:
            text = text * rnd.randint(2, 6)  # repeated records: match sequences in every stream
        parts.append(ctx.compress(text) if version == 2 and rnd.random() < 0.5 else oracle.compress(text, version=version))
    b = bytearray(merge_compressed(parts))
    for _ in range(rnd.choice((0, 1, 1, 1, 2, 3))):
        op = rnd.randrange(5)
        p = rnd.randrange(len(b))
        if rnd.random() < 0.35:  # aim at the headers
            p = min(len(b) - 1, rnd.randrange(0, 10 + 36 + 24))
        if op == 0:
            b[p] ^= 1 << rnd.randrange(8)
        elif op == 1:
            b[p] = rnd.choice((0, 1, 0x7F, 0x80, 0xFF, rnd.randrange(256)))
        elif op == 2:
            del b[p]
        elif op == 3:
            b.insert(p, rnd.randrange(256))
        else:
            del b[p:]
    return bytes(b)


def check_fuzz_fqz(ctx, oracle, seed):
    """Damaged archives (ROADMAP.md PR-006): where the oracle decodes the file the GPU gives the same text, where it rejects
    it so does the GPU — with the same code unless the damage sits inside a zstd frame (libzstd and the GPU decoder may
    trip over different symptoms of it) — and `fqz_check` agrees with `fqz_decompress`."""
    from fastqpacker_b200._binding import FqzError

    blob = fuzz_fqz(ctx, oracle, seed)
    try:
        want = oracle.decompress(blob)
    except oracle.OracleError as e:
        want = e.code
    try:
        got = ctx.decompress(blob)
    except FqzError as e:
        got = e.code
    try:
        chk = ctx.check(blob)[1]
    except FqzError as e:
        chk = e.code
    if isinstance(want, int):
        assert isinstance(got, int) and got < 0, (seed, want, len(got) if not isinstance(got, int) else got)
        assert got == want or -8 in (got, want), (seed, want, got)
        assert chk == got, (seed, chk, got)
    else:
        assert got == want, (seed, got if isinstance(got, int) else len(got), len(want))
        assert chk == len(want), (seed, chk)


def fuzz_bytes(rnd, n):
    """n bytes built from pieces a compressor treats differently: noise, small alphabets, runs, repeats at random reach,
    near-copies, counters, text."""
    out = bytearray()
    while len(out) < n:
        kind = rnd.randrange(8)
        m = rnd.choice((1, 3, 17, 200, 3000, 40000))
        m = rnd.randint(1, m)
        if kind == 0:
            out += rnd.randbytes(m)
        elif kind == 1:
            k = rnd.randint(1, 9)
            out += bytes(rnd.randrange(k) for _ in range(min(m, 5000)))
        elif kind == 2:
            out += bytes([rnd.randrange(256)]) * m
        elif kind == 3 and out:  # copy from anywhere behind (overlapping when the reach is short)
            reach = rnd.randint(1, min(len(out), rnd.choice((1, 2, 3, 8, 300, 70000, 1 << 20))))
            for _ in range(min(m, 20000)):
                out.append(out[-reach])
        elif kind == 4 and out:  # a near-copy: a few bytes changed
            reach = rnd.randint(1, len(out))
            piece = bytearray(out[len(out) - reach : len(out) - reach + m])
            for _ in range(len(piece) // 50 + 1):
                piece[rnd.randrange(len(piece))] = rnd.randrange(256)
            out += piece
        elif kind == 5:
            out += b"".join(struct.pack("<I", 100 + (i % rnd.randint(1, 9))) for i in range(min(m, 4000)))
        elif kind == 6:
            out += b"".join(b"\x17\x00read%d/1 some:%d:text" % (i, rnd.randrange(99999)) for i in range(min(m, 500)))
        else:
            out += b"ACGTTGCA"[rnd.randrange(8) :] * (m // 4 + 1)
    return bytes(out[:n])


def check_fuzz_zstd(ctx, oracle, seed, max_bytes=400_000):
    """Entropy stage against libzstd both ways on random structured data: frames libzstd writes at a random level (fast
    negative levels to 19), window, checksum and content-size setting decode bit-exact on the device, several frames back
    to back; what the device writes under either policy decodes under libzstd."""
    rnd = random.Random(0x25D + seed)
    n = rnd.choice((0, 1, 5, 100, 3000, 70000, max_bytes))
    n = rnd.randint(0, n)
    data = fuzz_bytes(rnd, n)
    level = rnd.choice((-5, -1, 1, 1, 2, 3, 5, 7, 9, 12, 16, 19))
    if rnd.random() < 0.5:
        z = oracle.zstd_compress(data, level)
    else:
        z = oracle.zstd_compress_adv(data, level, window_log=rnd.choice((0, 10, 14, 17, 20, 22)), content_size=rnd.choice((-1, 0, 1)),
                                     checksum=rnd.choice((-1, 0, 1)))
    assert oracle.zstd_decompress(z, cap=len(data) + 64) == data
    assert ctx.zstd_decompress(z) == data, (seed, "decode", level, n)
    if rnd.random() < 0.3:
        more = fuzz_bytes(rnd, rnd.randint(0, 5000))
        assert ctx.zstd_decompress(z + oracle.zstd_compress(more, 3)) == data + more, (seed, "two frames")
    for policy in (0, 1):
        mine = ctx.zstd_compress(data, policy)
        assert oracle.zstd_decompress(mine, cap=len(data) + 64) == data, (seed, "encode", policy, n)
        assert ctx.zstd_decompress(mine) == data, (seed, "round trip", policy, n)


def check_fuzz_feed(ctx, oracle, seed):
    """fqz_decompress_feed with windows and output room of random size (a byte at a time near the file header, windows that
    end inside block headers and payloads, room for less than one block): every call consumes whole blocks only, the
    pieces concatenate to the text of the whole-buffer call.  Same for fqz_compress_feed on the text."""
    import numpy as np

    from fastqpacker_b200.sharding import merge_compressed
    from tests.fastq_cases import rand_fastq

    rnd = random.Random(0xFEED + seed)
    version = rnd.choice((2, 2, 1))
    texts = [rand_fastq(rnd.randint(1, 200), rnd.randrange(1 << 20), lmin=1, lmax=rnd.choice((8, 60, 300))) for _ in range(rnd.randint(1, 6))]
    parts = [ctx.compress(t) if version == 2 and rnd.random() < 0.5 else oracle.compress(t, version=version) for t in texts]
    if any(p[9] != parts[0][9] for p in parts):
        parts = parts[:1]
    fqz = merge_compressed(parts)
    want = oracle.decompress(fqz)
    assert ctx.decompress(fqz) == want
    ds = ctx.decompress_stream()
    room = np.empty(rnd.choice((len(want) + 64, max(1, len(want) // 3), 1000)), dtype=np.uint8)
    res, pos, window, calls = b"", 0, b"", 0
    while True:
        step = rnd.choice((1, 3, 9, 37, 1000, len(fqz) // 3 + 1, len(fqz)))
        window += fqz[pos : pos + step]
        pos += step
        last = pos >= len(fqz)
        while True:
            calls += 1
            assert calls < 100_000
            try:
                m, used = ds.feed(window, last, room)
            except Exception as e:
                assert getattr(e, "code", 0) == -35 and not last, (seed, getattr(e, "code", 0), str(e))
                break
            if m < 0:  # FQZ_E_NOSPACE: -m bytes are needed for the next block; nothing was consumed
                assert -m > room.size
                room = np.empty(-m, dtype=np.uint8)
                continue
            res += room[:m].tobytes()
            window = window[used:]
            if not (last and window):  # with is_last the caller keeps calling until the window is empty
                break
            assert used > 0
        if last:
            assert not window
            break
    ds.close()
    assert res == want, (seed, len(res), len(want))


def check_fuzz_streams(ctx, oracle, seed):
    """Back end alone (blockReader.writeRecord, compress.go:944-1078) on six decoded streams with random damage — lengths
    changed, N counts and positions changed, length prefixes of headers and plus lines changed, streams cut or extended,
    NumRecords off by a few: the text the oracle builds, or its error (-17 where the reference would panic)."""
    from fastqpacker_b200._binding import FqzError
    from tests.fastq_cases import rand_fastq

    rnd = random.Random(0x57E + seed)
    text = rand_fastq(rnd.randint(1, 300), rnd.randrange(1 << 20), lmin=rnd.choice((0, 1, 30)), lmax=rnd.choice((31, 150, 600)),
                      plus_payload=rnd.random() < 0.4, n_rate=rnd.choice((0.0, 0.02, 0.3)), phred=rnd.choice((33, 64)))
    enc = oracle.encode_streams(text)
    streams, nrec, phred = [bytearray(s) for s in enc["streams"]], enc["nrec"], enc["phred64"]
    for _ in range(rnd.choice((0, 1, 1, 2, 3))):
        which = rnd.choice((0, 1, 2, 2, 3, 4, 4, 5, 5))
        s = streams[which]
        op = rnd.randrange(5)
        if op == 0 and s:
            del s[rnd.randrange(len(s)) :]
        elif op == 1:
            s += rnd.randbytes(rnd.randint(1, 40))
        elif op == 2 and s:
            p = rnd.randrange(len(s))
            s[p] = (s[p] + rnd.choice((1, -1, 2, 16, 128))) & 255
        elif op == 3 and s:  # aim at the low byte of a length field: the stream still parses, shifted
            p = rnd.randrange(len(s)) & ~3 if which == 5 else rnd.randrange(len(s))
            s[p] = rnd.choice((0, 1, 2, 3, 4, 5, 8, 150))
        elif op == 4 and s:
            del s[rnd.randrange(len(s))]
    nrec = max(0, nrec + rnd.choice((0, 0, 0, -1, 1, -7, 3)))
    phred = phred if rnd.random() < 0.8 else 1 - phred
    streams = [bytes(s) for s in streams]
    try:
        want = oracle.decode_streams(streams, nrec, phred)
    except oracle.OracleError as e:
        want = e.code
    try:
        got = ctx.decode_streams(streams, nrec, phred)
    except FqzError as e:
        got = e.code
    assert got == want, (seed, got if isinstance(got, int) else len(got), want if isinstance(want, int) else len(want))


def check_fuzz_hints(ctx, oracle, seed, nrec=2500, _cache={}):
    """Damage inside the index frames (skippable frames in front of the header / plus / N-position streams of a GPU-written
    block): any zstd reader skips their content, so the file still decodes to the same text unless the
    damage hits the frame's own magic or size — the hints may only ever cost the parallel walk, never change the result."""
    import struct

    from fastqpacker_b200._binding import FqzError
    from tests import synth

    key = (id(ctx), nrec)
    if key not in _cache:
        text = synth.fastq(1, 77, 0, nrec)
        _cache[key] = (text, ctx.compress(text))
    text, fqz = _cache[key]
    rnd = random.Random(0x41D7 + seed)
    hdr = struct.unpack_from("<9I", fqz, 10)
    pos = 10 + 36 + hdr[1] + hdr[2]
    spans = []
    for a in (2, 3, 4):
        if fqz[pos : pos + 4] == b"\x5e\x2a\x4d\x18":
            spans.append((pos, 8 + struct.unpack_from("<I", fqz, pos + 4)[0]))
        pos += hdr[1 + a]
    assert spans
    b = bytearray(fqz)
    for _ in range(rnd.choice((1, 1, 2, 4))):
        at, ln = rnd.choice(spans)
        p = at + (rnd.randrange(ln) if rnd.random() < 0.8 else rnd.randrange(min(ln, 24)))
        op = rnd.randrange(3)
        if op == 0:
            b[p] ^= 1 << rnd.randrange(8)
        elif op == 1:
            b[p] = rnd.choice((0, 1, 0xFF, rnd.randrange(256)))
        else:
            b[p : p + 2] = struct.pack("<H", rnd.choice((0, 1, 0x3FFF, 0x4000, 0xFFFF, rnd.randrange(65536))))
    blob = bytes(b)
    try:
        want = oracle.decompress(blob)
    except oracle.OracleError as e:
        want = e.code
    try:
        got = ctx.decompress(blob)
    except FqzError as e:
        got = e.code
    if isinstance(want, int):
        assert isinstance(got, int) and got < 0, (seed, want)
    else:
        assert want == text and got == want, (seed, got if isinstance(got, int) else len(got))


def check_fuzz_zstd_index(ctx, oracle, seed, _cache={}):
    """Damage inside the index frame of a device-written zstd stream (a skippable frame any reader steps over): the same
    bytes as libzstd decodes, or an error where libzstd reports one — the index may only cost the parallel frame walk."""
    import struct

    from fastqpacker_b200._binding import FqzError

    rnd = random.Random(0x1DE + seed)
    policy = seed & 1
    if policy not in _cache:  # (one context per test session)
        r2 = random.Random(policy)
        data = fuzz_bytes(r2, 3 * 131072 + 333)
        _cache[policy] = (data, ctx.zstd_compress(data, policy))
    data, z = _cache[policy]
    assert z[:4] == b"\x5e\x2a\x4d\x18", "no index frame in front of a multi-frame stream"
    ln = 8 + struct.unpack_from("<I", z, 4)[0]
    b = bytearray(z)
    for _ in range(rnd.choice((1, 1, 2, 4))):
        p = rnd.randrange(ln) if rnd.random() < 0.8 else rnd.randrange(min(ln, 24))
        op = rnd.randrange(3)
        if op == 0:
            b[p] ^= 1 << rnd.randrange(8)
        elif op == 1:
            b[p] = rnd.choice((0, 1, 0xFF, rnd.randrange(256)))
        else:
            b[p : p + 4] = struct.pack("<I", rnd.choice((0, 1, 0xFFFF, 0x10000, len(z), 0xFFFFFFFF, rnd.randrange(1 << 32))))
    blob = bytes(b[: len(z)])
    try:
        want = oracle.zstd_decompress(blob, cap=len(data) + 64)
    except oracle.OracleError as e:
        want = e.code
    try:
        got = ctx.zstd_decompress(blob)
    except FqzError as e:
        got = e.code
    if isinstance(want, int):
        assert isinstance(got, int) and got < 0, (seed, want)
    else:
        assert got == want, (seed, got if isinstance(got, int) else len(got), len(want))
