"""GPU compress path through the C-ABI: GPU-written .fqz must decode to the original FASTQ under
the oracle's reference-shaped decoder (libzstd standing in for klauspost's DecodeAll, i.e. the
proxy for `fqpack -d`), container layout must follow internal/fqformat/container.go."""
import random
import struct

import numpy as np
import pytest

from tests import synth
from tests.fastq_cases import BAD_CASES, GOOD_CASES, long_read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fastqpacker_b200 as fq

    return fq.context(0)


def _walk(fqz):
    """(header fields, [(nrec, sizes[6], orig_seq, orig_qual)]) following container.go."""
    assert fqz[:4] == b"FQZ\x00" and fqz[4] == 2
    pos, blocks = 10, []
    while pos < len(fqz):
        v = struct.unpack("<9I", fqz[pos : pos + 36])
        blocks.append(v)
        pos += 36 + sum(v[1:7])
    assert pos == len(fqz)
    return (struct.unpack("<I", fqz[5:9])[0], fqz[9]), blocks


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize(
    "name",
    ["empty1", "tiny", "zeros", "text", "quality_like", "random", "big_mixed", "runs", "one_frame_exact", "frame_plus_one"],
)
def test_zstd_frames_decode_under_libzstd(ctx, oracle, name, policy):
    rnd = random.Random(hash(name) & 0xFFFF)
    data = {
        "empty1": b"x",
        "tiny": b"abc",
        "zeros": bytes(100000),
        "text": b"hello world, hello zstd " * 5000,
        "quality_like": bytes(rnd.choice([0, 0, 0, 0, 0, 0, 1, 255, 2, 254, 3]) for _ in range(300000)),
        "random": bytes(rnd.randrange(256) for _ in range(200000)),
        "big_mixed": b"".join(bytes([rnd.randrange(4)]) * rnd.randrange(1, 40) + b"@ERR%d/1" % i for i in range(20000)),
        "runs": b"".join(bytes([i & 255]) * (i % 700) for i in range(600)),
        "one_frame_exact": bytes(rnd.randrange(7) for _ in range(65536)),
        "frame_plus_one": bytes(rnd.randrange(7) for _ in range(65537)),
    }[name]
    z = ctx.zstd_compress(data, policy)
    first = 0
    if z[:4] == b"\x5e\x2a\x4d\x18":  # skippable frame index in front of streams of >= 4 frames
        first = 8 + struct.unpack_from("<I", z, 4)[0]
    assert z[first : first + 4] == b"\x28\xb5\x2f\xfd"
    assert oracle.zstd_decompress(z) == data
    if len(data) > 1000 and name != "random":
        assert len(z) < len(data)


def test_zstd_empty(ctx):
    assert ctx.zstd_compress(b"", 0) == b""  # EncodeAll(empty) -> zero bytes


@pytest.mark.parametrize("name", sorted(GOOD_CASES))
def test_compress_decodes_under_oracle(ctx, oracle, name):
    text = GOOD_CASES[name]
    want = oracle.decompress(oracle.compress(text))  # the reference's own (lossy-normalising) round trip
    fqz = ctx.compress(text)
    assert oracle.decompress(fqz) == want
    (bs, flags), blocks = _walk(fqz)
    ref_fqz = oracle.compress(text)
    assert fqz[:10] == ref_fqz[:10]  # magic, version 2, BlockSize echo, Phred flag
    assert len(blocks) == (1 if want else 0)


def test_header_block_size_echo(ctx, oracle):
    text = GOOD_CASES["three"]
    fqz = ctx.compress(text, block_size=100)
    assert struct.unpack("<I", fqz[5:9])[0] == 100  # informational only
    assert oracle.decompress(fqz) == text


@pytest.mark.parametrize("name", sorted(BAD_CASES))
def test_compress_errors(ctx, name):
    from fastqpacker_b200 import FqzError

    text, code, rec = BAD_CASES[name]
    with pytest.raises(FqzError) as e:
        ctx.compress(text)
    assert e.value.code == code


def test_compress_long_read(ctx, oracle):
    from fastqpacker_b200 import FqzError

    with pytest.raises(FqzError) as e:
        ctx.compress(long_read(66000))
    assert e.value.code == -4 and "ambiguous bases beyond position" in str(e.value)
    ok = long_read(100)
    assert oracle.decompress(ctx.compress(ok)) == ok


@pytest.mark.parametrize("kind,count", [(0, 250000), (1, 130000)])
def test_compress_multi_block(ctx, oracle, kind, count):
    """BASELINE configs 2 / 4 shapes: several 100 000-record blocks, ratio next to the CPU path's."""
    import torch

    cap = count * 800
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.synth_device(kind, 0x5EED0001 + kind * 3, 0, count, buf.data_ptr(), cap)
    text = buf[:n].cpu().numpy()
    fqz = ctx.compress(text)
    (bs, flags), blocks = _walk(fqz)
    assert flags == (2 if kind else 0)
    assert [b[0] for b in blocks] == [100000] * (count // 100000) + ([count % 100000] if count % 100000 else [])
    assert oracle.decompress(fqz) == text.tobytes()
    ref = oracle.compress(text, threads=8)
    ratio_gpu, ratio_ref = n / len(fqz), n / len(ref)
    print(f"kind {kind}: ratio gpu {ratio_gpu:.3f} vs cpu-oracle(libzstd-1) {ratio_ref:.3f}")
    assert ratio_gpu > 0.98 * ratio_ref  # BASELINE target: within 2 % of the CPU path
    # pre-entropy streams inside the GPU-written container equal the oracle's, block by block
    for b in range(len(blocks)):
        got, nrec = oracle.block_streams(fqz, b)
        want, nrec2 = oracle.block_streams(ref, b)
        assert nrec == nrec2 and got == want


def test_sharded_compress(ctx, oracle):
    """SURVEY §8e flow on one GPU standing in for two ranks: block-aligned byte ranges, rank 0 decides
    the Phred flag, the others are forced to it and emit blocks only; the ordered gather is a plain
    concatenation and decodes to the original under the oracle."""
    import torch

    from fastqpacker_b200 import sharding

    count = 230000
    for kind in (0, 1):
        cap = count * 800
        buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
        n = ctx.synth_device(kind, 0x5EED0001 + kind * 3, 0, count, buf.data_ptr(), cap)
        text = buf[:n].cpu().numpy()
        world = 2
        allc = []
        before = 0
        for a, b in sharding.slice_bounds(n, world):
            local = (np.flatnonzero(text[a:b] == 10) + a).tolist()
            allc.append(sharding.block_cut_candidates(local, before))
            before += len(local)
        plan = sharding.plan_compress(allc, n, world)
        head, flag = ctx.compress_shard(text[plan[0][0] : plan[0][1]], phred64=-1, file_header=True)
        assert flag == kind
        tail, flag2 = ctx.compress_shard(text[plan[1][0] : plan[1][1]], phred64=flag, file_header=False)
        assert flag2 == flag and tail[:4] != b"FQZ\x00"
        merged = head + tail
        assert oracle.decompress(merged) == text.tobytes()
        assert merged == ctx.compress(text)  # sharding does not change a single byte of the file


def test_window_hand_over(ctx, oracle):
    """Several device windows per call (fqz_set_option window sizes): every window but the first starts
    at an arbitrary byte, is entered at the aligned address below it and skips the tail of the
    previous window's last line.  Same bytes as one big window."""
    text = oracle.synth(1, 77, 0, 320_000).tobytes()  # variable-length records: window cuts land on every alignment
    whole = ctx.compress(text)
    try:
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 45 << 20)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 45 << 20)
        cut = ctx.compress(text)
    finally:
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 0)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 0)
    assert cut == whole
    assert ctx.decompress(cut) == text


def test_window_hand_over_short_reads(ctx, oracle):
    """A window cut behind a record of a few bytes: the 16-byte word the next window is entered at holds
    several newlines, only the last of which ends the previous window."""
    from tests.fastq_cases import short_read_handover

    text = short_read_handover()
    try:
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 2 << 20)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 2 << 20)
        cut = ctx.compress(text)
    finally:
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 0)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 0)
    assert oracle.decompress(cut) == text
    assert ctx.decompress(cut) == text
    assert cut == ctx.compress(text)  # a block's bytes do not depend on which blocks share its device window


@pytest.mark.parametrize("name", sorted(__import__("tests.fastq_cases", fromlist=["x"]).REPETITIVE_CASES))
def test_duplicated_records_ratio(ctx, oracle, name):
    """VERDICT r1 J1 (i) the reference's BenchmarkCompress input, (ii) BenchmarkCompressBlock's 100 000 identical
    records, (iii) reads with >= 30 % exact duplicates: ratio_gpu >= 0.98 x ratio_oracle, bit-exact both ways."""
    from tests.fastq_cases import check_repetitive

    check_repetitive(ctx, oracle, name)


def test_shard_planning_calls(ctx, oracle):
    import torch

    from tests.fastq_cases import check_shard_planning

    def to_device(a):
        t = torch.from_numpy(a).cuda()
        return t.data_ptr(), t

    check_shard_planning(ctx, oracle, full=True, to_device=to_device)


def _two_contexts(oracle, dev_a, dev_b):
    """Two contexts in ONE process driven from two threads at the same time (the Go host is one process driving all
    GPUs): two shards of one file coded concurrently must concatenate to the bytes of the unsplit call."""
    import threading

    import numpy as np

    import fastqpacker_b200 as fq
    from fastqpacker_b200 import sharding

    ca, cb = fq.library().context(dev_a), fq.library().context(dev_b)
    text = oracle.synth(0, 0x5EED0001, 0, 330_000).tobytes()
    nl = np.flatnonzero(np.frombuffer(text, dtype=np.uint8) == 10)
    cut = int(nl[2 * sharding.LINES_PER_BLOCK - 1]) + 1  # two blocks for the first shard, 1.3 for the second
    whole = ca.compress(text)
    for _ in range(3):
        out, err = [None, None], []

        def work(i, c, data, phred, header):
            try:
                out[i] = c.compress_shard(data, phred, header)[0]
                assert c.decompress(whole) == text
            except Exception as e:  # noqa: BLE001
                err.append(e)

        ts = [threading.Thread(target=work, args=(0, ca, text[:cut], -1, True)), threading.Thread(target=work, args=(1, cb, text[cut:], 0, False))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not err, err
        assert sharding.merge_compressed(out) == whole
    assert cb.compress(text) == whole


def test_two_contexts_one_device(oracle):
    _two_contexts(oracle, 0, 0)


def test_two_contexts_two_devices(oracle):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    _two_contexts(oracle, 0, 1)


@pytest.mark.parametrize("frontend", [1, 2])
def test_frontend_variants_write_the_same_file(ctx, oracle, frontend):
    """FQZ_OPT_FRONTEND: two newline passes / fused metadata + scatter (look-back scans across CTAs) against the default
    kernels, over several blocks and several device windows."""
    text = oracle.synth(1, 0x5EED0004, 0, 260_000).tobytes()
    want = ctx.compress(text)
    try:
        ctx.set_option(ctx.OPT_FRONTEND, frontend)
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 60 << 20)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 60 << 20)
        got = ctx.compress(text)
    finally:
        ctx.set_option(ctx.OPT_FRONTEND, 0)
        ctx.set_option(ctx.OPT_WINDOW_BYTES, 0)
        ctx.set_option(ctx.OPT_HOST_WINDOW_BYTES, 0)
    assert got == want
    assert oracle.decompress(got) == text


@pytest.mark.parametrize("seed", range(24))
def test_duplicate_coder_fuzz(ctx, oracle, seed):
    """Random mixtures of read lengths (1 base .. 2 500), copy rates, reaches and near copies: whatever the record matcher
    decides, the file decodes to the input under the oracle and on the GPU."""
    from tests.fastq_cases import check_fuzz_duplicates

    check_fuzz_duplicates(ctx, oracle, 100 + seed, 40_000 if seed % 3 else 110_000)
