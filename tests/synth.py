"""Pure-Python twin of fastqpacker_b200/csrc/fqz_synth.cu (test infrastructure): the same
counter-based generator, so the device generator can be checked byte for byte at small sizes."""
M = (1 << 64) - 1


def _mix(z):
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


class _Rng:
    def __init__(self, seed, rec):
        self.s = _mix(seed ^ _mix(rec)) or 0x1234567
        self.left = 0
        self.cur = 0

    def next64(self):
        s = self.s
        s ^= (s << 13) & M
        s ^= s >> 7
        s ^= (s << 17) & M
        self.s = s
        return s

    def next16(self):
        if self.left == 0:
            self.cur = self.next64()
            self.left = 4
        v = self.cur & 0xFFFF
        self.cur >>= 16
        self.left -= 1
        return v


def dup_source(seed: int, rec: int) -> int:
    """kind 2: the record whose bases and qualities record `rec` carries (35 % copy one of the 400 in front)."""
    body = rec
    for _ in range(64):
        if body == 0:
            break
        h = _mix(seed ^ 0xD0B1E5 ^ _mix(body))
        if (h & 0xFFFF) >= 22938:
            break
        d = 1 + (h >> 16) % 400
        body = body - d if body > d else 0
    return body


def record(kind: int, seed: int, rec: int) -> bytes:
    g = _Rng(seed, rec)
    if kind != 1:
        T = 250000
        t = rec // T
        tile = (1 + (t // 48) % 2) * 1000 + (1 + (t // 16) % 3) * 100 + (1 + t % 16)
        x = 1000 + g.next64() % 20000
        y = 1000 + ((rec % T) * 2) // 5 + g.next16() % 40
        hdr = b"ERR532393.%d HWI-ST571:218:C2DACACXX:5:%d:%d:%d/1" % (rec + 1, tile, x, y)
        L = 150
    else:
        L = 50 + g.next64() % 251
        hdr = b"SRR_synth.%d %d length=%d" % (rec + 1, rec + 1, L)
    seq = bytearray()
    if kind == 2:
        body = dup_source(seed, rec)
        if body != rec:
            g = _Rng(seed, body)
            g.next64()
            g.next16()
    if kind != 1:
        nread = g.next16() < 655
        for _ in range(L):
            d = g.next16()
            b = b"ACGT"[d & 3]
            if nread and (d >> 2) < 1638:
                b = ord("N")
            seq.append(b)
    else:
        inrun = False
        for _ in range(L):
            d = g.next16()
            b = b"ACGT"[d & 3]
            r = d >> 2
            inrun = (r < 14746) if inrun else (r < 82)
            if inrun:
                b = ord("N")
            seq.append(b)
    qmax, qmin, base = (41, 2, 33) if kind != 1 else (40, 0, 64)
    q = qmax - 7 + g.next16() % 8
    tail = False
    qual = bytearray()
    for i in range(L):
        if i:
            d = g.next16()
            if tail:
                pass
            elif d < 66:
                tail = True
            elif d >= 49218:
                r = d - 49218
                m = (r >> 1) % 10
                mag = 1 if m < 6 else (2 if m < 9 else 3)
                q += mag if (r & 1) else -mag
                q = max(qmin, min(qmax, q))
        qual.append(base + (2 if tail else q))
    plus = b"+" + (hdr if kind == 1 else b"")
    return b"@" + hdr + b"\n" + bytes(seq) + b"\n" + plus + b"\n" + bytes(qual) + b"\n"


def fastq(kind: int, seed: int, first: int, count: int) -> bytes:
    return b"".join(record(kind, seed, first + i) for i in range(count))
