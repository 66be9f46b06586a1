"""GPU decompress path through the C-ABI: reference-shaped .fqz (oracle container + libzstd frames,
the stand-in for files written by fqpack) and GPU-written .fqz must decode bit-exact; error
behaviour follows internal/compress/compress.go:558-604,780-837,944-1078."""
import numpy as np
import pytest

from tests import synth
from tests.decode_cases import (
    ENT_SIZES,
    ZSTD_DATA,
    check_back_end,
    check_decode_errors,
    check_decompress_reference_written,
    check_block_header_fields,
    check_block_index,
    check_decompress_blocks,
    check_file_errors,
    check_round_trip,
    check_streaming,
    check_v1_file,
    check_zstd_ent_sizes,
    check_zstd_index,
    check_item_hints,
    check_zstd_libzstd_frames,
    check_zstd_round_trip,
    check_zstd_frame_shapes,
)
from tests.fastq_cases import GOOD_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fastqpacker_b200 as fq

    return fq.context(0)


@pytest.mark.parametrize("level", [1, 3, 9, 19])
@pytest.mark.parametrize("name", sorted(ZSTD_DATA))
def test_zstd_decodes_libzstd_frames(ctx, oracle, name, level):
    check_zstd_libzstd_frames(ctx, oracle, name, level, scale=4.0 if level < 19 else 1.0)


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("name", sorted(ZSTD_DATA))
def test_zstd_round_trip(ctx, oracle, name, policy):
    check_zstd_round_trip(ctx, oracle, name, policy, scale=4.0)


@pytest.mark.parametrize("n", ENT_SIZES + [5 * 131072 + 12345])
def test_zstd_entropy_policy_sizes(ctx, oracle, n):
    check_zstd_ent_sizes(ctx, oracle, n)


def test_item_hints(ctx, oracle):
    check_item_hints(ctx, oracle)


def test_zstd_index_frame(ctx, oracle):
    check_zstd_index(ctx, oracle)


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_back_end_matches_oracle(ctx, oracle, name):
    check_back_end(ctx, oracle, GOOD_CASES[name])


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_decompress_reference_written(ctx, oracle, name):
    check_decompress_reference_written(ctx, oracle, GOOD_CASES[name])


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_round_trip(ctx, oracle, name):
    check_round_trip(ctx, oracle, GOOD_CASES[name])


def test_v1_file(ctx, oracle):
    check_v1_file(ctx, oracle)


def test_decode_errors(ctx, oracle):
    check_decode_errors(ctx, oracle)


def test_file_errors(ctx, oracle):
    check_file_errors(ctx, oracle)


def test_block_header_fields(ctx, oracle):
    check_block_header_fields(ctx, oracle, nrec=1500)


def test_block_index_and_random_access(ctx, oracle):
    """fqz_block_index / fqz_decompress_blocks: the side-table index and the decode of any block range (v2, v1,
    reference-shaped and GPU-written blocks in one file)."""
    check_block_index(ctx.lib, oracle)
    check_decompress_blocks(ctx, oracle)


def test_streaming_small(ctx, oracle):
    check_streaming(ctx, oracle, nrec=300)


def test_streaming_multi_block(ctx, oracle):
    """250 000 records = 3 fqz blocks; feed windows that end inside blocks."""
    check_streaming(ctx, oracle, nrec=250000, chunk=30_000_000)


def _device_fastq(ctx, kind, seed, count):
    import torch

    cap = count * 800 + 4096
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.synth_device(kind, seed, 0, count, buf.data_ptr(), cap)
    return buf, n


@pytest.mark.parametrize("kind,count", [(0, 250000), (1, 130000)])
def test_decompress_multi_block(ctx, oracle, kind, count):
    """BASELINE config 3 shapes at test size: GPU-written and reference-written files, several blocks."""
    buf, n = _device_fastq(ctx, kind, 0x5EED0001 + kind * 3, count)
    text = buf[:n].cpu().numpy()
    ref = oracle.compress(text, threads=8)
    got = ctx.decompress(ref, cap=n + 4096)
    assert got == text.tobytes()
    fqz = ctx.compress(text)
    assert ctx.decompress(fqz, cap=n + 4096) == text.tobytes()


def test_device_round_trip_large(ctx):
    """Size-independent property at a larger size: decompress(compress(x)) == x entirely on the device
    (2 000 000 records, ~0.73 GB, 20 fqz blocks)."""
    import torch

    count = 2_000_000
    buf, n = _device_fastq(ctx, 0, 0x5EED0001, count)
    out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
    m = ctx.compress_device(buf.data_ptr(), n, out.data_ptr(), out.numel())
    back = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
    k = ctx.decompress_device(out.data_ptr(), m, back.data_ptr(), back.numel())
    assert k == n
    assert torch.equal(back[:n], buf[:n])
    # NOSPACE reports the size needed
    from fastqpacker_b200 import FqzError

    with pytest.raises(FqzError) as e:
        ctx.decompress_device(out.data_ptr(), m, back.data_ptr(), 1000)
    assert e.value.code == -15


def test_zstd_frame_shapes(ctx, oracle):
    """14 MB (> 100 blocks of 128 KiB) with matches up to ~3 MiB back, 4-16 MiB windows, frames without content size or
    checksum, Single_Segment frames."""
    check_zstd_frame_shapes(ctx, oracle, 14_000_000)


@pytest.mark.parametrize("first", range(0, 300, 100))
def test_archive_fuzz(ctx, oracle, first):
    """Damaged .fqz files (flipped bits, bytes replaced / dropped / inserted, cut tails; reference-shaped and GPU-written,
    v1 and v2): the text the oracle gives, or an error where it reports one (ROADMAP.md PR-006)."""
    from tests.decode_cases import check_fuzz_fqz

    for seed in range(first, first + 100):
        check_fuzz_fqz(ctx, oracle, seed)


@pytest.mark.parametrize("first", range(0, 40, 20))
def test_feed_fuzz(ctx, oracle, first):
    """fqz_decompress_feed with windows and output room of random size: whole blocks only, nothing consumed on FQZ_E_NOSPACE
    (not even the file header), the pieces concatenate to the whole-buffer result."""
    from tests.decode_cases import check_fuzz_feed

    for seed in range(first, first + 20):
        check_fuzz_feed(ctx, oracle, seed)


@pytest.mark.parametrize("first", range(0, 200, 100))
def test_zstd_fuzz(ctx, oracle, first):
    """Entropy stage against libzstd both ways on random structured data: libzstd frames of random level / window / checksum
    / content-size setting decode bit-exact, device-written frames of either policy decode under libzstd."""
    from tests.decode_cases import check_fuzz_zstd

    for seed in range(first, first + 100):
        check_fuzz_zstd(ctx, oracle, seed, 300_000 if seed >= 30 else 100_000)


@pytest.mark.parametrize("first", range(0, 1500, 500))
def test_streams_fuzz(ctx, oracle, first):
    """Back end alone on six decoded streams with random damage (lengths, N counts and positions, length prefixes, cuts,
    NumRecords off by a few): the text the oracle builds, or the FIRST error in the reference's record order
    (compress.go:944-1078) — also when several streams are damaged, when lengths add up past 2^32, and -17 where the
    reference would panic."""
    from tests.decode_cases import check_fuzz_streams

    for seed in range(first, first + 500):
        check_fuzz_streams(ctx, oracle, seed)


def test_index_frame_fuzz(ctx, oracle):
    """Damage inside the skippable index frames the GPU coder writes (item hints in front of the header / plus / N-position
    streams, the frame index of a multi-frame zstd stream): libzstd steps over them, so the result must not change — the
    index may only cost the parallel walk — unless the damage hits the skippable frame's own magic or size, where both
    readers report an error."""
    from tests.decode_cases import check_fuzz_hints, check_fuzz_zstd_index

    for seed in range(0, 200):
        check_fuzz_hints(ctx, oracle, seed)
    for seed in range(0, 200):
        check_fuzz_zstd_index(ctx, oracle, seed)
