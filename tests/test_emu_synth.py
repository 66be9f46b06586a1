"""The device FASTQ generator (run through the CPU emulation) against its pure-Python twin."""
import ctypes as C

import numpy as np
import pytest

from tests import synth


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu_lib import emu_context

    return emu_context()


@pytest.mark.parametrize("kind,first,count", [(0, 0, 40), (0, 249990, 20), (1, 0, 40), (1, 99999, 7), (2, 0, 60), (2, 5000, 40)])
def test_synth_matches_twin(emu, kind, first, count):
    want = synth.fastq(kind, 0x5EED0001, first, count)
    out = np.zeros(len(want) + 128, dtype=np.uint8)
    n = emu.synth_device(kind, 0x5EED0001, first, count, out.ctypes.data, out.size)
    assert n == len(want)
    assert out[:n].tobytes() == want


def test_synth_is_valid_fastq(oracle):
    for kind in (0, 1):
        text = synth.fastq(kind, 7, 0, 30)
        recs, used = oracle.parse(text)
        assert len(recs) == 30 and used == len(text)
        assert oracle.decompress(oracle.compress(text)) == text
    assert oracle.compress(synth.fastq(1, 7, 0, 30))[9] == 2  # kind 1 is detected as Phred+64


def test_synth_cpu_twin_and_duplicates(oracle):
    """oracle/fqz_synth_cpu.c against the Python twin; kind 2 holds the promised share of copied reads."""
    for kind, first in ((0, 0), (1, 17), (2, 0), (2, 3000)):
        assert oracle.synth(kind, 0x5EED0001, first, 50).tobytes() == synth.fastq(kind, 0x5EED0001, first, 50)
    lines = oracle.synth(2, 0x5EED0001, 0, 4000).tobytes().split(b"\n")
    seqs = lines[1::4]
    dups = sum(1 for i, s in enumerate(seqs) if s in set(seqs[max(0, i - 400) : i]))
    assert 0.28 * len(seqs) < dups < 0.42 * len(seqs)
