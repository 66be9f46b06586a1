"""gzip inputs shared by the oracle, emulation and GPU parity tests of the gzip input stage
(reference: cmd/fqpack/main.go:142-174, cmd/fqpack/main_test.go:12-161).

Every file here is written by zlib (an implementation independent of both the GPU decoder and Go's
compress/flate) or by hand from RFC 1951 / 1952; the expected text is known by construction."""
import random
import struct
import zlib

from tests import synth
from tests.fastq_cases import bench_compress_input, rand_fastq


def _member(text: bytes, level=6, mem_level=8, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0, flush_mode=zlib.Z_SYNC_FLUSH, flg=0,
            extra=b"", name=b"", comment=b"", bad_hcrc=False) -> bytes:
    """One gzip member written by hand around a raw deflate stream (RFC 1952)."""
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    if flush_every:
        parts = []
        for i in range(0, len(text), flush_every):
            parts.append(c.compress(text[i : i + flush_every]))
            parts.append(c.flush(flush_mode))
        parts.append(c.flush())
        raw = b"".join(parts)
    else:
        raw = c.compress(text) + c.flush()
    hdr = bytearray(b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\xff")
    if flg & 4:
        hdr += struct.pack("<H", len(extra)) + extra
    if flg & 8:
        hdr += name + b"\0"
    if flg & 16:
        hdr += comment + b"\0"
    if flg & 2:
        h = zlib.crc32(bytes(hdr)) & 0xFFFF
        hdr += struct.pack("<H", h ^ (0x5555 if bad_hcrc else 0))
    return bytes(hdr) + raw + struct.pack("<II", zlib.crc32(text), len(text) & 0xFFFFFFFF)


def bgzf(text: bytes, block=3000, level=6, eof=True) -> bytes:
    """BGZF (SAM spec 4.1): independent members of at most `block` bytes of text, each with the BC extra field."""
    out = bytearray()
    pieces = [text[i : i + block] for i in range(0, len(text), block)] + ([b""] if eof else [])
    for p in pieces:
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = c.compress(p) + c.flush()
        bsize = 18 + len(raw) + 8 - 1
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + raw + struct.pack("<II", zlib.crc32(p), len(p))
    return bytes(out)


def _fastq(nrec=350, seed=7):
    return rand_fastq(nrec, seed, lmin=40, lmax=160)


def good_cases(scale=1.0):
    """name -> (gzip bytes, text)"""
    n = max(40, int(350 * scale))
    fq = _fastq(n)
    rnd = random.Random(11)
    far = bytes(rnd.randrange(256) for _ in range(32768))
    cases = {}

    def add(name, gz, text):
        cases[name] = (gz, text)

    for lvl in (1, 6, 9):
        add(f"level{lvl}", _member(fq, lvl), fq)
    add("stored_level0", _member(fq + fq, 0), fq + fq)  # stored blocks only, more than 65535 bytes
    add("small_blocks", _member(fq, 6, mem_level=1), fq)  # a block every 128 symbols: fixed and dynamic codes alternate
    add("mem_level3", _member(fq, 9, mem_level=3), fq)
    add("fixed_only", _member(fq[: len(fq) // 3], 6, strategy=zlib.Z_FIXED), fq[: len(fq) // 3])
    add("huffman_only", _member(fq[: len(fq) // 2], 6, strategy=zlib.Z_HUFFMAN_ONLY), fq[: len(fq) // 2])
    add("rle", _member(fq[: len(fq) // 2], 6, strategy=zlib.Z_RLE), fq[: len(fq) // 2])
    add("sync_flush", _member(fq, 6, flush_every=5000), fq)  # empty stored blocks between compressed ones
    add("full_flush", _member(fq, 6, flush_every=7000, flush_mode=zlib.Z_FULL_FLUSH), fq)
    add("empty", _member(b""), b"")
    add("one_byte", _member(b"A"), b"A")
    add("newline_only", _member(b"\n" * 3000), b"\n" * 3000)  # distance 1, length 258
    add("far_matches", _member(far * 3, 9), far * 3)  # distance 32768
    add("incompressible", _member(far + far[::-1][:20000], 6), far + far[::-1][:20000])
    add("identical_records", _member(bench_compress_input(300), 6), bench_compress_input(300))
    add("three_members", _member(fq[:9000]) + _member(fq[9000:20000], 1) + _member(fq[20000:], 9), fq)
    add("members_with_empty", _member(fq[:5000]) + _member(b"") + _member(fq[5000:12000]) + _member(b""), fq[:12000])
    add("bgzf", bgzf(fq), fq)
    add("bgzf_tiny_blocks", bgzf(fq[:20000], block=700), fq[:20000])
    add("bgzf_no_eof", bgzf(fq[:15000], eof=False), fq[:15000])
    add("bgzf_then_plain", bgzf(fq[:10000], eof=False) + _member(fq[10000:]), fq)
    add("hdr_name", _member(fq[:3000], flg=8, name=b"reads.fq"), fq[:3000])
    add("hdr_all", _member(fq[:3000], flg=2 | 4 | 8 | 16, extra=b"ab\x03\0xyz", name=b"n" * 511, comment=b"a comment"), fq[:3000])
    add("hdr_extra_empty", _member(fq[:3000], flg=4), fq[:3000])
    add("hdr_reserved_bits", _member(fq[:3000], flg=0xE0), fq[:3000])  # Go ignores the reserved flag bits
    # restart points that nobody lands on: a deflate stream / a BGZF file stored inside stored blocks
    inner = _member(fq, 6, mem_level=2)
    add("gz_in_stored", _member(inner, 0), inner)
    inner = bgzf(fq[:20000], block=900)
    add("bgzf_in_stored", _member(inner, 0), inner)
    add("synth_kind0", _member(synth.fastq(0, 3, 0, max(20, n // 2)), 6), synth.fastq(0, 3, 0, max(20, n // 2)))
    add("synth_kind1", _member(synth.fastq(1, 4, 0, max(20, n // 2)), 6), synth.fastq(1, 4, 0, max(20, n // 2)))
    return cases


def _bits(vals):
    """LSB-first bit packer: [(value, nbits)] -> bytes"""
    acc = nb = 0
    out = bytearray()
    for v, k in vals:
        acc |= v << nb
        nb += k
        while nb >= 8:
            out.append(acc & 0xFF)
            acc >>= 8
            nb -= 8
    if nb:
        out.append(acc & 0xFF)
    return bytes(out)


def _wrap(raw: bytes, text: bytes = b"") -> bytes:
    return b"\x1f\x8b\x08\0\0\0\0\0\0\xff" + raw + struct.pack("<II", zlib.crc32(text), len(text))


def bad_cases():
    """name -> (gzip bytes, error class of oracle.gunzip_oracle, or None = any error)"""
    fq = _fastq(120)
    good = _member(fq)
    two = _member(fq[:6000]) + _member(fq[6000:])
    cases = {}
    cases["empty_input"] = (b"", "TRUNC")
    cases["short_header"] = (good[:7], "TRUNC")
    cases["bad_magic"] = (b"\x1f\x8c" + good[2:], "HEADER")
    cases["bad_method"] = (good[:2] + b"\x07" + good[3:], "HEADER")
    cases["cut_in_data"] = (good[: len(good) // 2], "TRUNC")
    cases["cut_in_first_block_header"] = (good[:12], "TRUNC")
    cases["cut_in_trailer"] = (good[:-3], "TRUNC")
    cases["cut_whole_trailer"] = (good[:-8], "TRUNC")
    cases["bad_crc"] = (good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], "CHECKSUM")
    cases["bad_isize"] = (good[:-1] + bytes([good[-1] ^ 0x40]), "CHECKSUM")
    first = _member(fq[:6000])
    cases["bad_crc_first_member"] = (first[:-8] + bytes([first[-8] ^ 0x80]) + first[-7:] + _member(fq[6000:]), "CHECKSUM")
    cases["bad_crc_second_member"] = (two[:-8] + bytes([two[-8] ^ 1]) + two[-7:], "CHECKSUM")
    cases["trailing_garbage"] = (good + b"0123456789abcdef", "HEADER")
    cases["trailing_zeros"] = (good + b"\0" * 16, "HEADER")  # Go: any 10 bytes that are not a gzip header
    cases["trailing_short"] = (good + b"\x1f\x8b\x08", "TRUNC")
    cases["second_member_cut"] = (two[:-40], "TRUNC")
    cases["bad_hcrc"] = (_member(fq[:2000], flg=2 | 8, name=b"x", bad_hcrc=True), "HEADER")
    cases["name_too_long"] = (_member(fq[:2000], flg=8, name=b"n" * 512), "HEADER")
    cases["name_unterminated"] = (b"\x1f\x8b\x08\x08\0\0\0\0\0\xff" + b"abc", "TRUNC")
    cases["extra_cut"] = (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x20\0abc", "TRUNC")
    cases["btype3"] = (_wrap(_bits([(1, 1), (3, 2)]) + b"\0\0\0\0"), "CORRUPT")
    cases["stored_len_mismatch"] = (_wrap(_bits([(1, 1), (0, 2)]) + b"\x05\0\x00\xff" + b"hello"), "CORRUPT")
    # fixed block: literal 'A' (code 0x30 + 0x41 = 0x71, 8 bits, MSB first), match length 3 (code 257 = 0000001, 7 bits) distance 2
    # (code 1, 5 bits) with only one byte of history
    rev = lambda v, k: int(format(v, "0%db" % k)[::-1], 2)
    cases["distance_too_far"] = (_wrap(_bits([(1, 1), (1, 2), (rev(0x71, 8), 8), (rev(1, 7), 7), (rev(1, 5), 5), (0, 7)]) + b"\0\0\0\0"), "CORRUPT")
    cases["length_symbol_286"] = (_wrap(_bits([(1, 1), (1, 2), (rev(0xC6, 8), 8)]) + b"\0\0\0\0\0\0"), "CORRUPT")  # 286 = 11000110
    cases["distance_symbol_30"] = (_wrap(_bits([(1, 1), (1, 2), (rev(0x71, 8), 8), (rev(1, 7), 7), (rev(30, 5), 5)]) + b"\0\0\0\0\0\0"), "CORRUPT")
    # dynamic block with too many literal/length codes (HLIT = 30 -> 287)
    cases["too_many_codes"] = (_wrap(_bits([(1, 1), (2, 2), (30, 5), (0, 5), (0, 4)]) + b"\0" * 12), "CORRUPT")
    # dynamic block whose code-length code is over-subscribed (four codes of length 1)
    cases["oversubscribed"] = (_wrap(_bits([(1, 1), (2, 2), (0, 5), (0, 5), (0, 4), (1, 3), (1, 3), (1, 3), (1, 3)]) + b"\0" * 12), "CORRUPT")
    rnd = random.Random(5)
    big = _member(_fastq(300), 6, mem_level=2)
    for i in range(6):  # a flipped bit somewhere in the deflate data: corrupt stream or a checksum mismatch
        p = rnd.randrange(12, len(big) - 8)
        cases[f"bitflip_{i}"] = (big[:p] + bytes([big[p] ^ (1 << rnd.randrange(8))]) + big[p + 1 :], None)
    return cases


CHUNKS = (0, 512, 4096)  # FQZ_OPT_GZ_CHUNK_BYTES: one warp for the file, many small chunks, a few chunks


def check_good(ctx, gunzip_oracle, name, cases=None, chunks=CHUNKS):
    gz, text = (cases or good_cases())[name]
    assert gunzip_oracle.gunzip(gz) == text, "oracle disagrees with the construction"
    assert ctx.is_gzip(gz)
    stats = []
    for ch in chunks:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, ch)
        try:
            out = ctx.gunzip(gz)
        finally:
            ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
        assert out == text, (name, ch, len(out), len(text))
        stats.append(ctx.gunzip_stats())
    return stats


def check_bad(ctx, gunzip_oracle, name, chunks=CHUNKS):
    from fastqpacker_b200._binding import FqzError

    gz, kind = bad_cases()[name]
    try:
        gunzip_oracle.gunzip(gz)
        okind = "OK"
    except gunzip_oracle.GunzipError as e:
        okind = e.kind
    if kind is not None:
        assert okind == kind, (name, okind)
    else:
        assert okind != "OK"
    for ch in chunks:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, ch)
        try:
            ctx.gunzip(gz)
            got = 0
        except FqzError as e:
            got = e.code
        finally:
            ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
        if kind is not None:
            assert got == gunzip_oracle.CODES[kind], (name, ch, got)
        else:
            assert got in (-18, -19, -20, -21), (name, ch, got)


def check_compress_gz(ctx, oracle, text, level=6, chunk=4096, bgzf_block=0):
    """gzip.NewReader + compress.Compress in one call: same .fqz bytes as compressing the text, decodes under the oracle."""
    gz = bgzf(text, block=bgzf_block) if bgzf_block else _member(text, level)
    ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, chunk)
    try:
        fqz = ctx.compress_gz(gz)
    finally:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
    assert fqz == ctx.compress(text)
    assert oracle.decompress(fqz) == oracle.decompress(oracle.compress(text))
    return len(gz), len(fqz)


def fuzz_file(seed: int, max_text=20000):
    """A random gzip file (1-3 members, random zlib parameters, flush points, header fields) and, for odd seeds, a random
    corruption of it (bit flips or a cut).  Returns (bytes, chunk option)."""
    rnd = random.Random(0xF00D + seed)
    kinds = ("fastq", "random", "repeat", "runs", "mixed")
    gz = b""
    for _ in range(rnd.randint(1, 3)):
        n = rnd.randint(0, max_text)
        kind = rnd.choice(kinds)
        if kind == "fastq":
            t = rand_fastq(max(1, n // 200), rnd.randrange(1 << 20), lmin=20, lmax=150)[:n]
        elif kind == "random":
            t = bytes(rnd.randrange(256) for _ in range(n))
        elif kind == "repeat":
            unit = bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 600)))
            t = (unit * (n // len(unit) + 1))[:n]
        elif kind == "runs":
            t = b"".join(bytes([rnd.randrange(65, 70)]) * rnd.randint(1, 400) for _ in range(n // 100 + 1))[:n]
        else:
            t = (rand_fastq(max(1, n // 400), 3, lmin=20, lmax=150) + bytes(rnd.randrange(256) for _ in range(n // 2)))[:n]
        flg = rnd.choice((0, 0, 8, 4, 2, 2 | 8 | 16))
        gz += _member(t, level=rnd.randint(0, 9), mem_level=rnd.randint(1, 9),
                      strategy=rnd.choice((zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED)),
                      flush_every=rnd.choice((0, 0, 997, 5000)), flush_mode=rnd.choice((zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH)), flg=flg,
                      extra=b"xy\x02\0ab", name=b"f.fq", comment=b"c")
    if seed & 1:
        if rnd.random() < 0.3 and len(gz) > 1:
            gz = gz[: rnd.randrange(1, len(gz))]
        else:
            b = bytearray(gz)
            for _ in range(rnd.randint(1, 3)):
                p = rnd.randrange(len(b))
                b[p] ^= 1 << rnd.randrange(8)
            gz = bytes(b)
    return gz, rnd.choice((0, 256, 512, 1024, 4096))


def check_fuzz(ctx, gunzip_oracle, seed, max_text=20000):
    """Same verdict as the oracle: the same text, or an error where it reports one (the class must match for cuts:
    `unexpected EOF`; for flipped bits zlib and Go's flate may name different symptoms of the same damage)."""
    from fastqpacker_b200._binding import FqzError

    gz, chunk = fuzz_file(seed, max_text)
    try:
        want, kind = gunzip_oracle.gunzip(gz), None
    except gunzip_oracle.GunzipError as e:
        want, kind = None, e.kind
    ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, chunk)
    try:
        got, code = ctx.gunzip(gz), 0
    except FqzError as e:
        got, code = None, e.code
    finally:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
    if kind is None:
        assert code == 0 and got == want, (seed, chunk, code)
    else:
        assert code in (-18, -19, -20, -21), (seed, chunk, kind, code)


def check_fuzz_compress_gz(ctx, gunzip_oracle, seed, max_text=8000):
    """fqz_compress_gz = gzip.NewReader in front of compress.Compress (main.go:142-174): on any input the verdict of
    inflating first and compressing the text — the same .fqz bytes, the gzip error, or the FASTQ error."""
    from fastqpacker_b200._binding import FqzError

    gz, chunk = fuzz_file(seed, max_text)
    if seed % 4 >= 2:  # the generator's texts are rarely FASTQ: put a member of real records in front
        gz = _member(rand_fastq(30, seed, lmin=20, lmax=150), level=1 + seed % 9) + gz

    def verdict(fn):
        try:
            return fn()
        except FqzError as e:
            return e.code

    ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, chunk)
    try:
        got = verdict(lambda: ctx.compress_gz(gz))
        text = verdict(lambda: ctx.gunzip(gz))
    finally:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
    try:
        want_text = gunzip_oracle.gunzip(gz)
    except gunzip_oracle.GunzipError:
        want_text = None
    if want_text is None:
        assert isinstance(text, int) and got == text, (seed, got, text)
    else:
        assert text == want_text
        assert got == verdict(lambda: ctx.compress(want_text)), (seed, got if isinstance(got, int) else len(got))
