"""gzip input on the GPU through the C-ABI (fqz_gunzip, fqz_gunzip_device, fqz_compress_gz): the text must equal what
Go's compress/gzip reader hands to compress.Compress (cmd/fqpack/main.go:142-174), restated by oracle/gunzip_oracle.py;
errors carry the class the Go reader reports."""
import zlib

import numpy as np
import pytest

from tests import synth
from tests.gzip_cases import _member, bad_cases, bgzf, check_bad, check_compress_gz, check_good, good_cases

pytestmark = pytest.mark.gpu
CASES = good_cases(1.0)


@pytest.fixture(scope="module")
def ctx():
    import fastqpacker_b200 as fq

    return fq.context(0)


@pytest.fixture(scope="module")
def gunzip_oracle():
    from oracle import gunzip_oracle

    return gunzip_oracle


@pytest.mark.parametrize("name", sorted(CASES))
def test_gunzip(ctx, gunzip_oracle, name):
    check_good(ctx, gunzip_oracle, name, CASES)


@pytest.mark.parametrize("name", sorted(bad_cases()))
def test_gunzip_errors(ctx, gunzip_oracle, name):
    check_bad(ctx, gunzip_oracle, name)


def test_restart_points_are_used_and_false_ones_dropped(ctx, gunzip_oracle):
    stats = check_good(ctx, gunzip_oracle, "small_blocks", CASES, chunks=(512,))[0]
    assert stats["parallel"] > 4 and stats["members"] == 1
    stats = check_good(ctx, gunzip_oracle, "bgzf", CASES, chunks=(1024,))[0]
    assert stats["parallel"] > 4 and stats["members"] > 4
    stats = check_good(ctx, gunzip_oracle, "gz_in_stored", CASES, chunks=(512,))[0]
    assert stats["dropped"] > 0
    # the speculative pass has room for 6 bytes of text per compressed byte: FASTQ fits, runs of one byte do not
    st = check_good(ctx, gunzip_oracle, "small_blocks", CASES, chunks=(512,))[0]
    assert st["decoded_twice"] < st["parallel"] // 2, st
    assert check_good(ctx, gunzip_oracle, "newline_only", CASES, chunks=(256,))[0]["decoded_twice"] == 1
    assert check_good(ctx, gunzip_oracle, "identical_records", CASES, chunks=(512,))[0]["decoded_twice"] > 0


def test_compress_gz_configs(ctx, oracle, sample_fq):
    """.fq.gz of BASELINE configs 1, 2 and 4: same .fqz as from the plain text."""
    check_compress_gz(ctx, oracle, sample_fq)
    check_compress_gz(ctx, oracle, synth.fastq(0, 1, 0, 4000), chunk=0)
    check_compress_gz(ctx, oracle, synth.fastq(1, 2, 0, 4000), level=1, chunk=8192)
    check_compress_gz(ctx, oracle, synth.fastq(0, 3, 0, 4000), bgzf_block=60000, chunk=65536)


@pytest.mark.parametrize("kind,level,chunk", [(0, 6, 0), (0, 1, 65536), (1, 9, 0)])
def test_gunzip_large(ctx, gunzip_oracle, kind, level, chunk):
    """~37 MB of generator FASTQ (100 000 records): hundreds of deflate blocks, default and forced chunking."""
    import torch

    n = 100_000
    cap = 420 * n
    d = torch.empty(cap, dtype=torch.uint8, device="cuda")
    m = ctx.synth_device(kind, 17, 0, n, d.data_ptr(), cap)
    text = d[:m].cpu().numpy().tobytes()
    gz = _member(text, level)
    ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, chunk)
    try:
        out = ctx.gunzip(gz)
    finally:
        ctx.set_option(ctx.OPT_GZ_CHUNK_BYTES, 0)
    assert zlib.crc32(out) == zlib.crc32(text) and out == text
    st = ctx.gunzip_stats()
    assert st["parallel"] > 16, st
    # device buffers: size query, then the real call
    g = torch.from_numpy(np.frombuffer(gz, dtype=np.uint8).copy()).cuda()
    gp = torch.zeros(g.numel() + 64, dtype=torch.uint8, device="cuda")
    gp[: g.numel()] = g
    from fastqpacker_b200._binding import FqzError

    with pytest.raises(FqzError) as e:
        ctx.gunzip_device(gp.data_ptr(), g.numel(), 0, 0)
    assert e.value.code == -15 and e.value.needed == len(text)
    o = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    assert ctx.gunzip_device(gp.data_ptr(), g.numel(), o.data_ptr(), len(text)) == len(text)
    assert torch.equal(o[: len(text)].cpu(), torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()))


def test_gunzip_bgzf_large(ctx, gunzip_oracle):
    text = synth.fastq(0, 5, 0, 20000)
    gz = bgzf(text, block=65280)
    assert ctx.gunzip(gz) == text
    st = ctx.gunzip_stats()
    assert st["members"] == (len(text) + 65279) // 65280 + 1


@pytest.mark.parametrize("first", range(40, 160, 20))
def test_gunzip_fuzz(ctx, gunzip_oracle, first):
    """Random gzip files (1-3 members, random zlib level / memLevel / strategy, flush points, optional header fields, random
    chunk size) and random damage to every second one: the same text as the oracle, or an error where it reports one."""
    from tests.gzip_cases import check_fuzz

    for seed in range(first, first + 20):
        check_fuzz(ctx, gunzip_oracle, seed, 8000)


@pytest.mark.parametrize("first", range(0, 200, 50))
def test_compress_gz_fuzz(ctx, gunzip_oracle, first):
    """fqz_compress_gz on random and damaged gzip files: the verdict of inflating first and compressing the text."""
    from tests.gzip_cases import check_fuzz_compress_gz

    for seed in range(first, first + 50):
        check_fuzz_compress_gz(ctx, gunzip_oracle, seed)
