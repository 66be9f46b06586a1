"""Runs the CUDA decompress kernels (zstd decoder, prefix walk, FASTQ emit) through the CPU
emulation (tests/emu) against the oracle.  Debug aid for the GPU kernels; the real parity tests are
the -m gpu ones (tests/test_gpu_decompress.py), which reuse these cases."""
import pytest

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


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu_lib import emu_context

    return emu_context()


@pytest.mark.parametrize("level", [1, 3, 9])
@pytest.mark.parametrize("name", sorted(ZSTD_DATA))
def test_zstd_decodes_libzstd_frames(emu, oracle, name, level):
    check_zstd_libzstd_frames(emu, oracle, name, level, scale=0.25)


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("name", sorted(ZSTD_DATA))
def test_zstd_round_trip(emu, oracle, name, policy):
    check_zstd_round_trip(emu, oracle, name, policy, scale=0.25)


@pytest.mark.parametrize("n", ENT_SIZES)
def test_zstd_entropy_policy_sizes(emu, oracle, n):
    check_zstd_ent_sizes(emu, oracle, n)


def test_zstd_frame_shapes(emu, oracle):
    check_zstd_frame_shapes(emu, oracle, 700_000, shapes=[(19, 0, 1, 1), (20, 1, 0, 3), (18, 0, 0, 1)])


def test_item_hints(emu, oracle):
    check_item_hints(emu, oracle, nrec=2200)  # the GPU suite runs 3000


def test_zstd_index_frame(emu, oracle):
    check_zstd_index(emu, oracle, policies=(1,), n=3 * 131072 + 333)


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_back_end_matches_oracle(emu, oracle, name):
    check_back_end(emu, oracle, GOOD_CASES[name])


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_decompress_reference_written(emu, oracle, name):
    check_decompress_reference_written(emu, oracle, GOOD_CASES[name])


@pytest.mark.parametrize("name", ["three", "rand_small", "rand_plus", "lossy", "empty"])
def test_round_trip(emu, oracle, name):
    check_round_trip(emu, oracle, GOOD_CASES[name])


def test_v1_file(emu, oracle):
    check_v1_file(emu, oracle)


def test_decode_errors(emu, oracle):
    check_decode_errors(emu, oracle)


def test_file_errors(emu, oracle):
    check_file_errors(emu, oracle)


def test_block_header_fields(emu, oracle):
    check_block_header_fields(emu, oracle, nrec=400, quick=True)


def test_block_index_and_random_access(emu, oracle):
    """fqz_block_index / fqz_decompress_blocks: the side-table index and the decode of any block range (v2, v1,
    reference-shaped and GPU-written blocks in one file)."""
    check_block_index(emu.lib, oracle)
    check_decompress_blocks(emu, oracle, dense=False)


def test_streaming(emu, oracle):
    check_streaming(emu, oracle, nrec=300)


def test_huf_kernel_variants(emu, oracle):
    """FQZ_OPT_HUF_KERNELS: the one-kernel and the three-kernel coder of the literals-only frames write frames of the same
    size that both decode (their Huffman codes may differ in tie breaks)."""
    from tests import synth

    text = synth.fastq(0, 33, 0, 4000)
    a = emu.compress(text)
    try:
        emu.set_option(emu.OPT_HUF_KERNELS, 1)
        b = emu.compress(text)
    finally:
        emu.set_option(emu.OPT_HUF_KERNELS, 0)
    assert oracle.decompress(a) == text and oracle.decompress(b) == text
    assert abs(len(a) - len(b)) <= 64


def test_archive_fuzz(emu, oracle):
    """Damaged .fqz files (tests/decode_cases.fuzz_fqz): the oracle's verdict; the GPU suite runs more seeds."""
    from tests.decode_cases import check_fuzz_fqz

    for seed in range(0, 40):
        check_fuzz_fqz(emu, oracle, seed)


def test_feed_fuzz(emu, oracle):
    """fqz_decompress_feed with random window and output sizes (tests/decode_cases.check_fuzz_feed)."""
    from tests.decode_cases import check_fuzz_feed

    for seed in (0, 13, 16, 22):  # 13, 16, 22: no room for the first block on the call that reads the file header
        check_fuzz_feed(emu, oracle, seed)


def test_zstd_fuzz(emu, oracle):
    """Entropy stage against libzstd both ways on random structured data (tests/decode_cases.check_fuzz_zstd)."""
    from tests.decode_cases import check_fuzz_zstd

    for seed in range(0, 25):
        check_fuzz_zstd(emu, oracle, seed, 100_000)


def test_streams_fuzz(emu, oracle):
    """Back end on damaged decoded streams (tests/decode_cases.check_fuzz_streams): the oracle's text or its first error."""
    from tests.decode_cases import check_fuzz_streams

    for seed in range(0, 150):
        check_fuzz_streams(emu, oracle, seed)


def test_index_frame_fuzz(emu, oracle):
    """Damage inside the index frames (item hints of a block, frame index of a zstd stream): never a different result."""
    from tests.decode_cases import check_fuzz_hints, check_fuzz_zstd_index

    for seed in range(0, 6):
        check_fuzz_hints(emu, oracle, seed, nrec=2200)
    for seed in range(0, 8):
        check_fuzz_zstd_index(emu, oracle, seed)
