"""FASTQ inputs shared by the oracle, emulation and GPU parity tests (mirrors the inputs of the
reference's own tests, internal/compress/compress_test.go and internal/fqparser/parser_test.go)."""
import random


def rand_fastq(nrec, seed, lmin=1, lmax=300, phred=33, n_rate=0.02, plus_payload=False, crlf=False, lower=False):
    rnd = random.Random(seed)
    out = bytearray()
    nl = b"\r\n" if crlf else b"\n"
    for i in range(nrec):
        L = rnd.randint(lmin, lmax)
        alphabet = "ACGTacgt" if lower else "ACGT"
        seq = bytearray(rnd.choice(alphabet).encode()[0] for _ in range(L))
        for k in range(L):
            if rnd.random() < n_rate:
                seq[k] = rnd.choice(b"NnRY.-")
        q = rnd.randint(phred + 2, phred + 40)
        qual = bytearray()
        for _ in range(L):
            if rnd.random() < 0.3:
                q = min(phred + 41, max(phred, q + rnd.randint(-3, 3)))
            qual.append(q)
        hdr = b"@read%d/%d some:%d:text" % (i, seed, rnd.randint(0, 99999))
        out += hdr + nl + bytes(seq) + nl + (b"+" + (hdr[1:] if plus_payload else b"")) + nl + bytes(qual) + nl
    return bytes(out)


GOOD_CASES = {
    "single": b"@SEQ_ID\nGATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT\n+\n!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65\n",
    "three": b"@SEQ_1\nACGTACGT\n+\nIIIIIIII\n@SEQ_2\nTGCATGCA\n+\n!!!!!!!!\n@SEQ_3\nAAAACCCC\n+\n55555555\n",
    "nbases": b"@SEQ_N\nACGTNNNNACGT\n+\nIIII!!!!IIII\n",
    "plus_payload": b"@SEQ_1\nACGTACGT\n+SEQ_1 extra payload\nIIIIIIII\n@SEQ_2\nTGCATGCA\n+\nIIIIIIII\n",
    "illumina152": b"@HWI-ST123:4:1101:14346:1976#0/1\n" + b"ACGT" * 38 + b"\n+\n" + b"I" * 152 + b"\n",
    "empty": b"",
    "phred64": b"@SEQ_64\nACGTACGT\n+\nhhhhhhhh\n",
    "phred64_multi": b"@S1\nACGT\n+\nhhhh\n@S2\nACGT\n+\n@ABC\n@S3\nTTTT\n+\nefgh\n",
    "mixed_phred": b"@A\nACGT\n+\nhhhh\n@B\nACGT\n+\n!!!!\n",
    "ambiguous_phred": b"@A\nACGT\n+\n;<=>\n",
    "lossy": b"@a\nacgtnRyk\r\n+\r\nIIIIIIII\r\n",
    "unterminated": b"@A\nAC\n+\nII\n@B\nGT\n+\nII",
    "partial_tail_2lines": b"@A\nAC\n+\nII\n@B\nGT\n",
    "qual_starts_with_at": b"@A\nAC\n+\n@I\n@B\nGG\n+\n+I\n",
    "zero_length": b"@A\n\n+\n\n@B\nA\n+\nI\n",
    "lengths_1_to_40": b"".join(b"@r%d\n%s\n+\n%s\n" % (L, (b"ACGTN" * 9)[:L], (b"ABCDEFGHIJ" * 5)[:L]) for L in range(1, 41)),
    "all_n_300": b"@n\n" + b"N" * 300 + b"\n+\n" + b"#" * 300 + b"\n",
    "quality_0xff": b"@x\nACGT\n+\n\xff\xff\xff\xff\n",
    "rand_small": rand_fastq(200, 1),
    "rand_crlf": rand_fastq(50, 2, crlf=True),
    "rand_plus": rand_fastq(300, 3, plus_payload=True, phred=64),
    "rand_lower": rand_fastq(100, 4, lower=True, n_rate=0.2),
    "rand_long": rand_fastq(12, 5, lmin=900, lmax=5000),
}

BAD_CASES = {
    "no_at": (b"SEQ\nAC\n+\nII\n", -1, 0),
    "no_at_second": (b"@A\nAC\n+\nII\nB\nAC\n+\nII\n", -1, 1),
    "no_plus": (b"@SEQ\nACGT\n-\nIIII\n", -2, 0),
    "len_mismatch": (b"@SEQ\nACGT\n+\nII\n", -3, 0),
    "trailing_blank": (b"@A\nAC\n+\nII\n\n", -1, 1),
    "blank_first": (b"\n@A\nAC\n+\nII\n", -1, 0),
    "tail_bad_plus": (b"@A\nAC\n+\nII\n@B\nAC\nx\n", -2, 1),
}


def long_read(n_at=None, length=70000):
    seq = bytearray(b"ACGT" * (length // 4))
    if n_at is not None:
        seq[n_at] = ord("N")
    return b"@SEQ_LONG\n" + bytes(seq) + b"\n+\n" + b"I" * length + b"\n"


def short_read_handover(nrec=230_000, want_mod=12):
    """1-6 base reads; record 99 999 (the last of block 0) is a 1-base read that ends at an offset with
    (offset & 15) == want_mod, so that the next device window is entered `want_mod` bytes in front of its
    first record and those bytes hold more than one newline."""
    rnd = random.Random(99)
    recs = []
    for i in range(nrec):
        L = 1 if i == 99_999 else rnd.randint(1, 6)
        seq = bytes(rnd.choice(b"ACGT") for _ in range(L))
        qual = bytes(rnd.randint(35, 74) for _ in range(L))
        hdr = b"@r" if i == 99_999 else b"@r%d" % i
        recs.append(hdr + b"\n" + seq + b"\n+\n" + qual + b"\n")
    end = sum(len(r) for r in recs[:100_000])
    pad = (want_mod - end) % 16
    recs[0] = b"@" + b"p" * pad + recs[0][1:]
    text = b"".join(recs)
    assert sum(len(r) for r in recs[:100_000]) % 16 == want_mod
    return text
