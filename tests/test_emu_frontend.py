"""Runs the CUDA front-end kernels through the CPU emulation (tests/emu) and compares the six
pre-entropy streams with the oracle byte for byte.  Debug aid for the GPU kernels; the real
parity tests are the -m gpu ones."""
import pytest

from tests.fastq_cases import BAD_CASES, GOOD_CASES, long_read


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu_lib import emu_context

    return emu_context()


@pytest.mark.parametrize("frontend", [0, 1, 2])  # FQZ_OPT_FRONTEND: the three sets of front-end kernels give the same bytes
@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_streams_match_oracle(emu, oracle, name, frontend):
    text = GOOD_CASES[name]
    want = oracle.encode_streams(text)
    try:
        emu.set_option(emu.OPT_FRONTEND, frontend)
        got = emu.encode_streams(text)
    finally:
        emu.set_option(emu.OPT_FRONTEND, 0)
    for k in ("nrec", "phred64", "orig_seq", "orig_qual", "consumed"):
        assert got[k] == want[k], k
    for nm, a, b in zip(oracle.STREAM_NAMES, got["streams"], want["streams"]):
        assert a == b, nm


@pytest.mark.parametrize("frontend", [0, 2])
@pytest.mark.parametrize("name", sorted(BAD_CASES))
def test_errors_match_oracle(emu, oracle, name, frontend):
    text, code, rec = BAD_CASES[name]
    with pytest.raises(oracle.OracleError) as oe:
        oracle.encode_streams(text)
    assert oe.value.code == code
    from fastqpacker_b200._binding import FqzError

    try:
        emu.set_option(emu.OPT_FRONTEND, frontend)
        with pytest.raises(FqzError) as ge:
            emu.encode_streams(text)
    finally:
        emu.set_option(emu.OPT_FRONTEND, 0)
    assert ge.value.code == code and ge.value.record == rec


def test_long_read_guard(emu, oracle):
    from fastqpacker_b200._binding import FqzError

    with pytest.raises(FqzError) as ge:
        emu.encode_streams(long_read(66000))
    assert ge.value.code == -4
    ok = long_read(100)
    want = oracle.encode_streams(ok)
    got = emu.encode_streams(ok)
    assert got["streams"] == want["streams"]


def test_forced_phred(emu, oracle):
    text = GOOD_CASES["three"]
    for p in (0, 1):
        assert emu.encode_streams(text, phred64=p)["streams"] == oracle.encode_streams(text, phred64=p)["streams"]


def test_window_hand_over_short_reads(emu, oracle):
    """a window cut behind a 1-base record; the next window is entered 12 bytes in front
    of its first record and those bytes hold three newlines.  Device path with 2 MiB windows."""
    from tests.fastq_cases import short_read_handover

    text = short_read_handover(nrec=205_000)
    try:
        emu.set_option(emu.OPT_WINDOW_BYTES, 2 << 20)
        emu.set_option(emu.OPT_HOST_WINDOW_BYTES, 2 << 20)
        cut = emu.compress(text)
    finally:
        emu.set_option(emu.OPT_WINDOW_BYTES, 0)
        emu.set_option(emu.OPT_HOST_WINDOW_BYTES, 0)
    assert oracle.decompress(cut) == text


@pytest.mark.parametrize("name", sorted(k for k, v in __import__("tests.fastq_cases", fromlist=["x"]).REPETITIVE_CASES.items() if v[1]))
def test_duplicated_records_ratio(emu, oracle, name):
    from tests.fastq_cases import check_repetitive

    check_repetitive(emu, oracle, name)


def test_shard_planning_calls(emu, oracle):
    from tests.fastq_cases import check_shard_planning

    check_shard_planning(emu, oracle, full=False)


@pytest.mark.parametrize("seed", range(6))
def test_duplicate_coder_fuzz(emu, oracle, seed):
    from tests.fastq_cases import check_fuzz_duplicates

    check_fuzz_duplicates(emu, oracle, seed, 1500)


def test_fastq_parser_fuzz(emu, oracle):
    """Edited FASTQ texts (tests/fastq_cases.fuzz_fastq): the oracle's verdict, record index included; the GPU suite runs
    more seeds."""
    from tests.fastq_cases import check_fuzz_fastq

    for seed in range(0, 60):
        check_fuzz_fastq(emu, oracle, seed)
