"""tests/feed_loop.c — the producer / feed / writer loop of INTEGRATION.md's cgo shim, compiled in C against
include/fqzgpu.h (no Go toolchain in this image) — run against the CPU-emulated library here and against
libfqzgpu.so on the GPU (-m gpu)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp, libdir, libname):
    exe = os.path.join(tmp, "feed_loop")
    subprocess.check_call(
        ["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "feed_loop.c"), "-L", libdir,
         "-l" + libname, "-Wl,-rpath," + libdir]
    )
    return exe


def _run_loop(exe, tmp, text, oracle, window, dwindow):
    src, fqz, back = (os.path.join(tmp, n) for n in ("in.fq", "out.fqz", "back.fq"))
    open(src, "wb").write(text)
    r = subprocess.run([exe, "c", src, fqz, str(window)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr
    z = open(fqz, "rb").read()
    assert oracle.decompress(z) == text  # what `fqpack -d` would write
    r = subprocess.run([exe, "d", fqz, back, str(dwindow)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr
    assert open(back, "rb").read() == text
    # output room too small for the first block: FQZ_E_NOSPACE consumes nothing (not even the file header), the loop grows
    # the buffer and presents the same window again
    r = subprocess.run([exe, "d", fqz, back, str(dwindow), "4096"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr
    assert open(back, "rb").read() == text
    return z


def test_feed_loop_emulated(tmp_path, oracle):
    from tests.emu.emu_lib import emu_context
    from tests.fastq_cases import rand_fastq

    emu = emu_context()
    exe = _build(str(tmp_path), os.path.join(ROOT, "tests", "emu", "_build"), "fqzgpu_emu")
    text = rand_fastq(1500, 21, lmin=20, lmax=120, n_rate=0.0)
    z = _run_loop(exe, str(tmp_path), text, oracle, 1 << 20, 1 << 16)  # the decompress window has to grow to hold the block
    assert z == emu.compress(text)
    # errors come back with the reference's texts
    bad = os.path.join(str(tmp_path), "bad.fq")
    open(bad, "wb").write(b"@A\nACGT\n-\nIIII\n")
    r = subprocess.run([exe, "c", bad, os.path.join(str(tmp_path), "bad.fqz")], capture_output=True, text=True)
    assert r.returncode == 1 and "parsing FASTQ" in r.stderr and "record 0" in r.stderr


@pytest.mark.gpu
def test_feed_loop_gpu(tmp_path, oracle):
    import fastqpacker_b200 as fq

    exe = _build(str(tmp_path), os.path.join(ROOT, "fastqpacker_b200"), "fqzgpu")
    text = oracle.synth(0, 0x5EED0001, 0, 450_000).tobytes()  # 4.5 blocks, several feed calls at 64 MiB windows
    z = _run_loop(exe, str(tmp_path), text, oracle, 64 << 20, 8 << 20)
    assert z == fq.context(0).compress(text)
