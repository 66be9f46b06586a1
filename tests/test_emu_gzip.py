"""Runs the gzip input kernels (restart-point search, parallel DEFLATE decode, window resolution, CRC-32) through the
CPU emulation (tests/emu) against the gunzip oracle.  The parity tests proper are the -m gpu ones
(tests/test_gpu_gzip.py), which reuse these cases at full size."""
import pytest

from tests.gzip_cases import bad_cases, check_bad, check_compress_gz, check_good, good_cases

SCALE = 0.3
CASES = good_cases(SCALE)


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu_lib import emu_context

    return emu_context()


@pytest.fixture(scope="module")
def gunzip_oracle():
    from oracle import gunzip_oracle

    return gunzip_oracle


@pytest.mark.parametrize("name", sorted(CASES))
def test_gunzip(emu, gunzip_oracle, name):
    check_good(emu, gunzip_oracle, name, CASES, chunks=(1024,))


@pytest.mark.parametrize("name", ["level6", "small_blocks", "three_members", "bgzf", "stored_level0", "empty"])
def test_gunzip_one_chunk(emu, gunzip_oracle, name):
    check_good(emu, gunzip_oracle, name, CASES, chunks=(0,))


@pytest.mark.parametrize("name", sorted(bad_cases()))
def test_gunzip_errors(emu, gunzip_oracle, name):
    check_bad(emu, gunzip_oracle, name, chunks=(0, 1024) if name.startswith(("cut", "bad_crc", "bitflip")) else (1024,))


def test_restart_points_are_used_and_false_ones_dropped(emu, gunzip_oracle):
    stats = check_good(emu, gunzip_oracle, "small_blocks", CASES, chunks=(512,))[0]
    assert stats["parallel"] > 4 and stats["members"] == 1
    stats = check_good(emu, gunzip_oracle, "bgzf", CASES, chunks=(1024,))[0]
    assert stats["parallel"] > 4 and stats["members"] > 4
    stats = check_good(emu, gunzip_oracle, "gz_in_stored", CASES, chunks=(512,))[0]
    assert stats["dropped"] > 0
    # the speculative pass has room for 6 bytes of text per compressed byte: FASTQ fits, runs of one byte do not
    st = check_good(emu, gunzip_oracle, "small_blocks", CASES, chunks=(512,))[0]
    assert st["decoded_twice"] < st["parallel"] // 2, st
    assert check_good(emu, gunzip_oracle, "newline_only", CASES, chunks=(256,))[0]["decoded_twice"] == 1
    assert check_good(emu, gunzip_oracle, "identical_records", CASES, chunks=(512,))[0]["decoded_twice"] > 0


def test_compress_gz(emu, oracle, sample_fq):
    from tests.fastq_cases import rand_fastq

    check_compress_gz(emu, oracle, sample_fq)
    check_compress_gz(emu, oracle, rand_fastq(150, 9, lmin=30, lmax=120), chunk=1024)
    check_compress_gz(emu, oracle, rand_fastq(150, 10, lmin=30, lmax=120), bgzf_block=2000, chunk=1024)


def test_gunzip_fuzz(emu, gunzip_oracle):
    """Random gzip files and random damage (tests/gzip_cases.fuzz_file); the GPU suite runs more seeds."""
    from tests.gzip_cases import check_fuzz

    for seed in range(40, 70):
        check_fuzz(emu, gunzip_oracle, seed, 8000)


def test_compress_gz_fuzz(emu, gunzip_oracle):
    from tests.gzip_cases import check_fuzz_compress_gz

    for seed in range(0, 12):
        check_fuzz_compress_gz(emu, gunzip_oracle, seed)
