"""Pins the CPU oracle against every known-answer vector the reference's own unit tests hold for
the block-codec path and against the golden streams of testdata/sample.fq.
CPU only."""
import hashlib
import json
import os
import struct

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- internal/encoder/quality_test.go:9-55 (DeltaEncode), :57-103 (DeltaDecode)
DELTA_VECTORS = [
    ([40, 40, 40, 40], [40, 0, 0, 0]),
    ([30, 31, 32, 33], [30, 1, 1, 1]),
    ([40, 39, 38, 37], [40, 255, 255, 255]),
    ([40], [40]),
    ([], []),
]


@pytest.mark.parametrize("plain,enc", DELTA_VECTORS)
def test_delta_vectors(oracle, plain, enc):
    assert list(oracle.delta_encode(bytes(plain))) == enc
    assert list(oracle.delta_decode(bytes(enc))) == plain


def test_delta_roundtrip_152(oracle):  # quality_test.go:142-166
    q = bytes((30 + (i * 7) % 11) for i in range(152))
    assert oracle.delta_decode(oracle.delta_encode(q)) == q


# ---- quality_test.go:203-265 DetectEncoding table
DETECT = [
    ([b'!"#$'], 0),
    ([b"5III"], 0),
    ([b"!5I?"], 0),
    ([b"@ABh"], 1),
    ([b"efgh"], 1),
    ([b"hhh", b"@AB"], 1),
    ([], 0),
    ([b""], 0),
    ([b";<=>"], 0),
]


@pytest.mark.parametrize("quals,want", DETECT)
def test_detect_encoding(oracle, quals, want):
    assert oracle.detect_encoding(quals) == want


# ---- quality_test.go:267-345 Normalize / Denormalize
def test_normalize_vectors(oracle):
    assert list(oracle.normalize_quality(b"!5I", 0)) == [0, 20, 40]
    assert list(oracle.normalize_quality(b"@Th", 1)) == [0, 20, 40]
    assert oracle.denormalize_quality(bytes([0, 20, 40]), 0) == b"!5I"
    assert oracle.denormalize_quality(bytes([0, 20, 40]), 1) == b"@Th"
    assert oracle.normalize_quality(b"", 0) == b""


# ---- internal/encoder/sequence_test.go:13-53 N positions; :62-82 case normalisation
NPOS = [
    (b"ACGT", []),
    (b"AAAA", []),
    (b"TTTT", []),
    (b"acgt", []),
    (b"ACNGT", [2]),
    (b"NACTGN", [0, 5]),
    (b"NNNN", [0, 1, 2, 3]),
]


@pytest.mark.parametrize("seq,want", NPOS)
def test_pack_npos(oracle, seq, want):
    packed, npos = oracle.pack_bases(seq)
    assert npos == want
    expect = seq.upper().replace(b"n", b"N")
    assert oracle.unpack_bases(packed, npos, len(seq)) == expect


def test_pack_bit_layout(oracle):
    # base i -> bits 2(i&3) of byte i>>2, A=0 C=1 G=2 T=3 (sequence.go:153-159)
    packed, npos = oracle.pack_bases(b"ACGT")
    assert packed == bytes([0b11100100]) and npos == []
    packed, _ = oracle.pack_bases(b"TGCAT")
    assert packed == bytes([0b00011011, 0b00000011])
    packed, npos = oracle.pack_bases(b"RYKM.-*")  # every non-ACGTacgt byte packs as 0 and is listed
    assert packed == b"\0\0" and npos == list(range(7))


@pytest.mark.parametrize("n,size", [(1, 1), (2, 1), (3, 1), (4, 1), (5, 2), (8, 2), (9, 3), (100, 25), (152, 38)])
def test_packed_size(oracle, n, size):  # sequence_test.go:113-139
    packed, _ = oracle.pack_bases(bytes(b"ACGT"[i % 4] for i in range(n)))
    assert len(packed) == size


@pytest.mark.parametrize(
    "seq",
    [b"ACGTACGTACGTACGT", b"A" * 16, b"C" * 16, b"G" * 16, b"T" * 16, b"ACGTNNNNACGTNNNN", b"N", b"A", b"ACGT" * 25, b""],
)
def test_pack_roundtrip(oracle, seq):  # sequence_test.go:87-111,141-146
    packed, npos = oracle.pack_bases(seq)
    assert oracle.unpack_bases(packed, npos, len(seq)) == seq


# ---- internal/fqparser/parser_test.go
def test_parse_record(oracle):  # :13-27
    recs, _ = oracle.parse(b"@SEQ_ID\nGATTTGGGG\n+\n!''*((((*\n")
    assert recs == [(b"SEQ_ID", b"GATTTGGGG", b"", b"!''*((((*")]


def test_parse_multiple(oracle):  # :29-67
    recs, used = oracle.parse(b"@A\nAC\n+\nII\n@B\nGT\n+\nII\n@C\nNN\n+x\n!!\n")
    assert [r[0] for r in recs] == [b"A", b"B", b"C"]
    assert recs[2][2] == b"x"


def test_parse_errors(oracle):  # :69-95
    recs, _ = oracle.parse(b"")
    assert recs == []
    with pytest.raises(oracle.OracleError) as e:
        oracle.parse(b"SEQ\nAC\n+\nII\n")
    assert "must start with @" in str(e.value)
    with pytest.raises(oracle.OracleError) as e:
        oracle.parse(b"@SEQ\nACGT\n+\nII\n")
    assert "lengths must match" in str(e.value)
    with pytest.raises(oracle.OracleError) as e:
        oracle.parse(b"@SEQ\nACGT\n-\nIIII\n")
    assert "must start with +" in str(e.value)


def test_parse_quirks(oracle):
    # unterminated last line drops the record (parser.go:209-220)
    recs, _ = oracle.parse(b"@A\nAC\n+\nII\n@B\nGT\n+\nII")
    assert len(recs) == 1
    # CRLF is stripped (parser.go:213-215)
    recs, _ = oracle.parse(b"@A x\r\nAC\r\n+p\r\nII\r\n")
    assert recs == [(b"A x", b"AC", b"p", b"II")]
    # trailing blank line is a header error (parser.go:142)
    with pytest.raises(oracle.OracleError):
        oracle.parse(b"@A\nAC\n+\nII\n\n")
    # a quality line starting with '@' is fine (strict 4-line framing)
    recs, _ = oracle.parse(b"@A\nAC\n+\n@I\n@B\nGG\n+\n+I\n")
    assert [r[3] for r in recs] == [b"@I", b"+I"]


def test_read_batch_semantics(oracle):  # parser_test.go:149-182
    text = b"".join(b"@r%d\nACGT\n+\nIIII\n" % i for i in range(250))
    pos, got = 0, []
    while True:
        recs, used = oracle.parse(text[pos:], max_records=100)
        if not recs:
            break
        got.append(len(recs))
        pos += used
    assert got == [100, 100, 50]


# ---- golden streams of testdata/sample.fq
def test_sample_golden_streams(oracle, sample_fq):
    g = json.load(open(os.path.join(HERE, "golden", "sample_streams.json")))
    r = oracle.encode_streams(sample_fq)
    assert r["nrec"] == 3 and r["phred64"] == 0 and r["consumed"] == len(sample_fq)
    assert r["orig_seq"] == 180 and r["orig_qual"] == 180
    for name, got in zip(oracle.STREAM_NAMES, r["streams"]):
        assert got.hex() == g["streams"][name], name
    sha = {n: hashlib.sha256(s).hexdigest()[:16] for n, s in zip(oracle.STREAM_NAMES, r["streams"])}
    assert sha["seqPacked"] == "7f54336169a148cc" and sha["quality"] == "e006f193e5762c8b"
    assert sha["headers"] == "34fdbde9b4e6e888" and sha["nPositions"] == "2ad55adcc0cb838c"


def test_sample_file_layout(oracle, sample_fq):
    fqz = oracle.compress(sample_fq)
    assert fqz[:10].hex() == "46515a0002a086010000"  # SURVEY App. B file header
    assert fqz[10:14] == struct.pack("<I", 3)
    assert fqz[10 + 28 : 10 + 36] == struct.pack("<II", 180, 180)
    sizes = struct.unpack("<6I", fqz[14:38])
    assert 10 + 36 + sum(sizes) == len(fqz)
    out, info = oracle.decompress(fqz, return_info=True)
    assert out == sample_fq and info["blocks"] == 1 and info["records"] == 3


# ---- internal/compress/compress_test.go round trips
RT_INPUTS = [
    b"@SEQ_ID\nGATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT\n+\n!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65\n",
    b"@SEQ_1\nACGTACGT\n+\nIIIIIIII\n@SEQ_2\nTGCATGCA\n+\n!!!!!!!!\n@SEQ_3\nAAAACCCC\n+\n55555555\n",
    b"@SEQ_N\nACGTNNNNACGT\n+\nIIII!!!!IIII\n",
    b"@SEQ_1\nACGTACGT\n+SEQ_1 extra payload\nIIIIIIII\n@SEQ_2\nTGCATGCA\n+\nIIIIIIII\n",
    b"@HWI-ST123:4:1101:14346:1976#0/1\n" + b"ACGT" * 38 + b"\n+\n" + b"I" * 152 + b"\n",
    b"",
    b"@SEQ_64\nACGTACGT\n+\nhhhhhhhh\n",
    b"@S1\nACGT\n+\nhhhh\n@S2\nACGT\n+\n@ABC\n@S3\nTTTT\n+\nefgh\n",
]


@pytest.mark.parametrize("text", RT_INPUTS)
def test_roundtrip(oracle, text):
    fqz = oracle.compress(text)
    assert oracle.decompress(fqz) == text
    if not text:
        assert len(fqz) == 10  # header-only file (compress_test.go:160-173)


def test_roundtrip_1000_and_threads(oracle):  # compress_test.go:125-158,198-281
    text = b"".join(b"@SEQ_%d\n%s\n+\n%s\n" % (i, b"ACGTACGTAC" * 10, b"IIIIIIIIII" * 10) for i in range(1000))
    for th in (1, 4, 16):
        assert oracle.decompress(oracle.compress(text, threads=th, block_size=100)) == text
    fqz = oracle.compress(text, block_size=100)
    assert struct.unpack("<I", fqz[5:9])[0] == 100  # header echoes -b ...
    assert struct.unpack("<I", fqz[10:14])[0] == 1000  # ... but one block holds all records


def test_phred64_flag_and_mixed(oracle):  # compress_test.go:330-412,477-500
    t64 = b"".join(b"@S%d\nACGTACGT\n+\nhgfedcba\n" % i for i in range(50))
    fqz = oracle.compress(t64)
    assert fqz[9] == 2
    assert oracle.decompress(fqz) == t64
    mixed = b"@A\nACGT\n+\nhhhh\n@B\nACGT\n+\n!!!!\n"
    fqz = oracle.compress(mixed)
    assert fqz[9] == 0 and oracle.decompress(fqz) == mixed


def test_v1_compat(oracle):  # compress_test.go:502-592
    text = b"@SEQ_1\nACGTACGT\n+\nIIIIIIII\n"
    fqz = oracle.compress(text, version=1, block_size=1)
    assert fqz[4] == 1 and len(fqz) >= 10 + 32
    assert oracle.decompress(fqz) == text
    # v1 loses plus payloads by construction
    t2 = b"@SEQ_1\nACGT\n+payload\nIIII\n"
    assert oracle.decompress(oracle.compress(t2, version=1)) == b"@SEQ_1\nACGT\n+\nIIII\n"


def test_long_read_guard(oracle):  # compress_test.go:651-697
    seq = bytearray(b"ACGT" * 17500)
    seq[66000] = ord("N")
    bad = b"@SEQ_LONG\n" + bytes(seq) + b"\n+\n" + b"I" * 70000 + b"\n"
    with pytest.raises(oracle.OracleError) as e:
        oracle.compress(bad)
    assert "ambiguous bases beyond position" in str(e.value)
    seq[66000] = ord("A")
    seq[100] = ord("N")
    ok = b"@SEQ_LONG\n" + bytes(seq) + b"\n+\n" + b"I" * 70000 + b"\n"
    assert oracle.decompress(oracle.compress(ok)) == ok


def test_lossy_normalisation(oracle):
    text = b"@a\nacgtnRyk\r\n+\r\nIIIIIIII\r\n"
    assert oracle.decompress(oracle.compress(text)) == b"@a\nACGTNNNN\n+\nIIIIIIII\n"


def test_bad_container(oracle):  # container_test.go:36-43, compress.go:571-573
    with pytest.raises(oracle.OracleError) as e:
        oracle.decompress(b"NOPE\x02\xa0\x86\x01\x00\x00")
    assert "invalid magic" in str(e.value)
    with pytest.raises(oracle.OracleError) as e:
        oracle.decompress(b"FQZ\x00\x03\xa0\x86\x01\x00\x00")
    assert "unsupported file version" in str(e.value)


def test_zstd_frame_checksum_is_xxh64(oracle):
    data = bytes(range(256)) * 40
    frame = oracle.zstd_compress(data)
    assert frame[:4] == b"\x28\xb5\x2f\xfd" and frame[4] & 4
    assert struct.unpack("<I", frame[-4:])[0] == oracle.xxh64(data) & 0xFFFFFFFF
    assert oracle.zstd_decompress(frame + frame) == data + data  # concatenated frames
