#!/usr/bin/env python
"""TEST-ONLY: runs the fuzzers of tests/*_cases.py against the oracle over a range of seeds, on the CPU emulation of the
CUDA sources (tests/emu) — the long runs behind the seed counts in DESIGN.md 2; the test suites run a few seeds of each
(emulated) and a few hundred (-m gpu).

    python tools/fuzz_emu.py <target> <first seed> <last seed + 1> [--asan] [--frontend N]

targets: fastq (edited FASTQ), archive (damaged .fqz), streams (damaged decoded streams into the back end), zstd (entropy
stage against libzstd both ways), zstd-index / hints (damage inside the index frames), feed (fqz_decompress_feed with random
windows), gzip (random gzip files and damage), compress-gz.  --asan re-executes under the ASan + UBSan build of the emulation
(make -C tests/emu SAN=... B=_build_asan).  Exit code 1 when a seed fails."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SAN = "-fsanitize=address,undefined -fno-sanitize-recover=undefined"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 3:
        sys.exit(__doc__)
    target, lo, hi = args[0], int(args[1]), int(args[2])
    if "--asan" in sys.argv and not os.environ.get("FQZ_EMU_LIB"):
        emu = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-C", emu, "-s", "-j8", "SAN=" + SAN, "B=_build_asan"])
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, FQZ_EMU_LIB=os.path.join(emu, "_build_asan", "libfqzgpu_emu.so"), LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
        sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))

    from oracle import fqz_oracle as oracle
    from oracle import gunzip_oracle
    from tests import decode_cases as dc
    from tests import fastq_cases as fc
    from tests import gzip_cases as gc
    from tests.emu.emu_lib import emu_context

    ctx = emu_context()
    if "--frontend" in sys.argv:
        ctx.set_option(ctx.OPT_FRONTEND, int(sys.argv[sys.argv.index("--frontend") + 1]))
    run = {
        "fastq": lambda s: fc.check_fuzz_fastq(ctx, oracle, s),
        "archive": lambda s: dc.check_fuzz_fqz(ctx, oracle, s),
        "streams": lambda s: dc.check_fuzz_streams(ctx, oracle, s),
        "zstd": lambda s: dc.check_fuzz_zstd(ctx, oracle, s, 300_000 if s >= 30 else 100_000),
        "zstd-index": lambda s: dc.check_fuzz_zstd_index(ctx, oracle, s),
        "hints": lambda s: dc.check_fuzz_hints(ctx, oracle, s),
        "feed": lambda s: dc.check_fuzz_feed(ctx, oracle, s),
        "gzip": lambda s: gc.check_fuzz(ctx, gunzip_oracle, s, 8000),
        "compress-gz": lambda s: gc.check_fuzz_compress_gz(ctx, gunzip_oracle, s),
    }[target]
    t0, bad = time.time(), []
    for seed in range(lo, hi):
        try:
            run(seed)
        except AssertionError as e:
            bad.append(seed)
            print("FAIL", seed, str(e)[:300], flush=True)
    print(f"{target}: {hi - lo} seeds in {time.time() - t0:.0f} s, failures: {bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
