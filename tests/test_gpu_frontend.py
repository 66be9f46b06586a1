"""GPU parity: CUDA front end (through the C-ABI) vs the oracle, byte for byte."""
import json
import os

import numpy as np
import pytest

from tests import synth
from tests.fastq_cases import BAD_CASES, GOOD_CASES, long_read

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    import fastqpacker_b200 as fq

    return fq.context(0)


@pytest.mark.parametrize("frontend", [0, 1, 2])  # FQZ_OPT_FRONTEND: the three sets of front-end kernels give the same bytes
@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_streams_match_oracle(ctx, oracle, name, frontend):
    text = GOOD_CASES[name]
    want = oracle.encode_streams(text)
    try:
        ctx.set_option(ctx.OPT_FRONTEND, frontend)
        got = ctx.encode_streams(text)
    finally:
        ctx.set_option(ctx.OPT_FRONTEND, 0)
    for k in ("nrec", "phred64", "orig_seq", "orig_qual", "consumed"):
        assert got[k] == want[k], k
    for nm, a, b in zip(oracle.STREAM_NAMES, got["streams"], want["streams"]):
        assert a == b, nm


def test_sample_golden(ctx, sample_fq):
    g = json.load(open(os.path.join(HERE, "golden", "sample_streams.json")))
    got = ctx.encode_streams(sample_fq)
    from fastqpacker_b200._binding import STREAM_NAMES

    for nm, s in zip(STREAM_NAMES, got["streams"]):
        assert s.hex() == g["streams"][nm], nm
    assert got["nrec"] == 3 and got["phred64"] == 0


@pytest.mark.parametrize("name", sorted(BAD_CASES))
def test_errors_match_oracle(ctx, name):
    from fastqpacker_b200 import FqzError

    text, code, rec = BAD_CASES[name]
    with pytest.raises(FqzError) as ge:
        ctx.encode_streams(text)
    assert ge.value.code == code and ge.value.record == rec


def test_long_read_guard(ctx, oracle):
    from fastqpacker_b200 import FqzError

    with pytest.raises(FqzError) as ge:
        ctx.encode_streams(long_read(66000))
    assert ge.value.code == -4
    ok = long_read(100)
    assert ctx.encode_streams(ok)["streams"] == oracle.encode_streams(ok)["streams"]


@pytest.mark.parametrize("kind,count", [(0, 120000), (1, 60000)])
def test_synth_block_matches_oracle(ctx, oracle, kind, count):
    """One full 100 000-record block (and the cut after it) of BASELINE configs 2 and 4."""
    import torch

    cap = count * 800
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.synth_device(kind, 0x5EED0001 + kind * 3, 0, count, buf.data_ptr(), cap)
    text = buf[:n].cpu().numpy()
    # the device generator equals its Python twin on a prefix
    twin = synth.fastq(kind, 0x5EED0001 + kind * 3, 0, 50)
    assert text[: len(twin)].tobytes() == twin
    want = oracle.encode_streams(text)
    got = ctx.encode_streams(text)
    assert got["nrec"] == want["nrec"] == min(count, 100000)
    assert got["consumed"] == want["consumed"] and got["phred64"] == want["phred64"] == kind
    for nm, a, b in zip(oracle.STREAM_NAMES, got["streams"], want["streams"]):
        assert a == b, nm


@pytest.mark.parametrize("first", range(0, 400, 100))
def test_fastq_parser_fuzz(ctx, oracle, first):
    """Edited FASTQ texts (bytes dropped / inserted / replaced, lines dropped or doubled, the tail cut): the same six streams
    or the same error at the same record as the oracle (the fuzz target of the reference's ROADMAP.md PR-006)."""
    from tests.fastq_cases import check_fuzz_fastq

    try:
        ctx.set_option(ctx.OPT_FRONTEND, (first // 100) % 3)  # the three sets of front-end kernels give the same verdicts
        for seed in range(first, first + 100):
            check_fuzz_fastq(ctx, oracle, seed)
    finally:
        ctx.set_option(ctx.OPT_FRONTEND, 0)
