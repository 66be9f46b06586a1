"""Regenerates tests/golden/sample_streams.json from tests/golden/sample.fq with an independent
pure-Python restatement of the reference's per-record stream assembly
(internal/compress/compress.go:474-520, internal/encoder/sequence.go:139-184,
internal/encoder/quality.go:22-103) and checks the SHA-256 prefixes of SURVEY.md App. B."""
import hashlib
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
EXPECT = {
    "seqPacked": (45, "7f54336169a148cc"),
    "quality": (180, "e006f193e5762c8b"),
    "headers": (60, "34fdbde9b4e6e888"),
    "plusLines": (6, "b0f66adc83641586"),
    "nPositions": (22, "2ad55adcc0cb838c"),
    "seqLengths": (12, "e39d8ebde04d5151"),
}


def streams_of(text: bytes):
    lines = text.split(b"\n")[:-1]
    recs = [lines[i : i + 4] for i in range(0, len(lines) - len(lines) % 4, 4)]
    quals = b"".join(r[3] for r in recs)
    phred64 = 0 if (not quals or min(quals) < 59 or min(quals) < 64) else 1
    off = 64 if phred64 else 33
    code = {c: v for v, pair in enumerate(("Aa", "Cc", "Gg", "Tt")) for c in pair.encode()}
    out = {k: bytearray() for k in EXPECT}
    for h, s, p, q in recs:
        packed = bytearray((len(s) + 3) // 4)
        npos = []
        for i, b in enumerate(s):
            packed[i >> 2] |= code.get(b, 0) << ((i & 3) * 2)
            if b not in code and i < 65536:
                npos.append(i)
        out["seqPacked"] += packed
        out["nPositions"] += struct.pack("<H", len(npos) & 0xFFFF) + b"".join(struct.pack("<H", x) for x in npos)
        out["seqLengths"] += struct.pack("<I", len(s))
        norm = [(b - off) & 255 for b in q]
        out["quality"] += bytes([norm[0]] + [(norm[i] - norm[i - 1]) & 255 for i in range(1, len(norm))]) if norm else b""
        out["headers"] += struct.pack("<H", (len(h) - 1) & 0xFFFF) + h[1:]
        out["plusLines"] += struct.pack("<H", (len(p) - 1) & 0xFFFF) + p[1:]
    return {k: bytes(v) for k, v in out.items()}, phred64


if __name__ == "__main__":
    text = open(os.path.join(HERE, "sample.fq"), "rb").read()
    st, phred64 = streams_of(text)
    for k, (n, sha) in EXPECT.items():
        assert len(st[k]) == n, (k, len(st[k]))
        assert hashlib.sha256(st[k]).hexdigest()[:16] == sha, (k, hashlib.sha256(st[k]).hexdigest()[:16])
    json.dump(
        {"phred64": phred64, "num_records": 3, "streams": {k: v.hex() for k, v in st.items()},
         "file_header_hex": "46515a0002a086010000", "orig_seq": 180, "orig_qual": 180},
        open(os.path.join(HERE, "sample_streams.json"), "w"), indent=1)
    print("ok")
