"""BASELINE.json's full sizes on the device, checked through size-independent properties (the oracle
cannot follow at these sizes in seconds): the round trip reproduces the generator's bytes exactly, the
container walks block by block with the record counts adding up, every full block holds 100 000 records,
and a slice of the device generator equals its CPU twin (which the oracle tests tie to the reference)."""
import struct

import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED0001


@pytest.fixture(scope="module")
def ctx():
    import fastqpacker_b200 as fq

    return fq.context(0)


def _device_fastq(ctx, torch, kind, seed, nrec, per_record):
    cap = nrec * per_record + (1 << 20)
    d = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = 0
    for first in range(0, nrec, 4_000_000):
        c = min(4_000_000, nrec - first)
        n += ctx.synth_device(kind, seed, first, c, d.data_ptr() + n, cap - n)
    return d, n


def _walk(hdrs, m):
    pos, blocks = 10, []
    while pos < m:
        v = struct.unpack_from("<9I", hdrs, pos)
        blocks.append(v)
        pos += 36 + sum(v[1:7])
    assert pos == m
    return blocks


@pytest.mark.parametrize(
    "kind,seed,nrec,per_record,phred64",
    [
        (0, SEED, 25_000_000, 372, 0),        # config 2 / 3: 9 GB of 150 bp Phred+33 reads
        (1, 0x5EED0004, 2_000_000, 760, 1),   # config 4: 50-300 bp, Phred+64, N-heavy, '+' payloads
    ],
)
def test_full_size_round_trip(ctx, oracle, kind, seed, nrec, per_record, phred64):
    import torch

    d_in, n = _device_fastq(ctx, torch, kind, seed, nrec, per_record)
    # the device generator is the CPU generator (first and last 2000 records)
    head = oracle.synth(kind, seed, 0, 2000).tobytes()
    tail = oracle.synth(kind, seed, nrec - 2000, 2000).tobytes()
    assert d_in[: len(head)].cpu().numpy().tobytes() == head
    assert d_in[n - len(tail) : n].cpu().numpy().tobytes() == tail
    d_out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
    m = ctx.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())
    hdrs = d_out[:m].cpu().numpy().tobytes()
    assert hdrs[:4] == b"FQZ\x00" and hdrs[4] == 2 and hdrs[9] == (2 if phred64 else 0)
    blocks = _walk(hdrs, m)
    assert sum(b[0] for b in blocks) == nrec
    assert all(b[0] == 100_000 for b in blocks[:-1]) and len(blocks) == (nrec + 99_999) // 100_000
    assert all(b[7] == b[8] for b in blocks)  # OriginalSeqSize == OriginalQualSize
    assert n / m > (4.5 if kind == 0 else 3.5)
    d_back = torch.empty(n + (1 << 16), dtype=torch.uint8, device="cuda")
    k = ctx.decompress_device(d_out.data_ptr(), m, d_back.data_ptr(), d_back.numel())
    assert k == n
    assert bool(torch.equal(d_back[:n], d_in[:n]))
    # the first, a middle and the last block decode under the oracle too (it stands in for `fqpack -d`): each must be
    # the generator's records of that block (not only block 0)
    offs = [10]
    for b in blocks:
        offs.append(offs[-1] + 36 + sum(b[1:7]))
    for i in sorted({0, len(blocks) // 2, len(blocks) - 1}):
        got = oracle.decompress(hdrs[:10] + hdrs[offs[i] : offs[i + 1]])
        want = oracle.synth(kind, seed, i * 100_000, blocks[i][0]).tobytes()
        assert got == want, f"block {i} differs under the oracle"
