"""Multi-GPU host logic on CPU: world_size-2 gloo.  Every rank plans from its own
byte slice plus one tiny gather; the codec work itself is stood in for by the oracle here (the GPU
version of the same flow is tests/test_gpu_compress.py::test_sharded_compress)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fastqpacker_b200 import sharding  # noqa: E402

RECORDS = 230_000  # three fqz blocks: 100 000 + 100 000 + 30 000


def _text():
    from oracle import fqz_oracle as oracle

    return oracle.synth(0, 0x5EED0001, 0, RECORDS)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fqz_oracle as oracle

    text = _text()
    n = text.size
    a, b = sharding.slice_bounds(n, world)[rank]
    # 1. local newline positions (on the GPU this is k_newline_count / k_newline_index)
    local = (np.flatnonzero(text[a:b] == 10) + a).tolist()
    counts = [None] * world
    dist.all_gather_object(counts, len(local))  # the one tiny exchange
    before = sum(counts[:rank])
    cand = sharding.block_cut_candidates(local, before)
    allc = [None] * world
    dist.all_gather_object(allc, cand)
    plan = sharding.plan_compress(allc, n, world)
    lo, hi = plan[rank]
    # 2. every rank codes its own whole blocks; rank 0 owns block 0 and therefore the Phred decision,
    #    which is broadcast (the other ranks must not detect on their own first block, compress.go:146-164)
    flag = [None]
    if rank == 0:
        es = oracle.encode_streams(text[lo:hi].tobytes(), max_records=100000)
        flag[0] = int(es["phred64"])
    dist.broadcast_object_list(flag, src=0)
    part = oracle.compress(text[lo:hi].tobytes(), threads=2) if hi > lo else b"FQZ\x00\x02\xa0\x86\x01\x00\x00"
    assert part[9] == (2 if flag[0] else 0)  # the shard was coded with the broadcast flag
    if rank == 1:
        part = part[10:]  # headerless, as fqz_compress_shard(emit_file_header=0) emits it
    parts = [None] * world
    dist.all_gather_object(parts, part)  # ordered host gather
    if rank == 0:
        merged = sharding.merge_compressed(parts)
        whole = oracle.compress(text.tobytes(), threads=2)
        assert plan[0][0] == 0 and plan[-1][1] == n
        for r in range(world - 1):
            assert plan[r][1] == plan[r + 1][0]
            cut = plan[r][1]
            assert cut == 0 or (text[cut - 1] == 10 and int((text[:cut] == 10).sum()) % sharding.LINES_PER_BLOCK == 0)
        assert merged == whole  # same blocks, same order, same bytes
        # 3. decompress side: contiguous block runs per rank, outputs concatenate to the original
        shards = sharding.shard_container(merged, world)
        back = b"".join(oracle.decompress(s) for s in shards)
        assert back == text.tobytes()
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_plan_and_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_plan_decompress_balances_and_covers():
    from oracle import fqz_oracle as oracle

    fqz = oracle.compress(_text().tobytes(), threads=4)
    version, flags, blocks = sharding.walk_container(fqz)
    assert version == 2 and [b.records for b in blocks] == [100000, 100000, 30000]
    for world in (1, 2, 3, 4, 8):
        plan = sharding.plan_decompress(blocks, world)
        assert plan[0][0] == 0 and plan[-1][1] == len(blocks)
        assert all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.walk_container(b"FQX\x00" + fqz[4:])
    with pytest.raises(ValueError):
        sharding.walk_container(fqz[:-5])


def _bgzf_worker(rank, world, port, tmp):
    """One BGZF input over two ranks: members are cut between ranks from their headers alone, every rank inflates its
    own run (gunzip oracle here, fqz_gunzip_device on the GPU) and the block planning runs on those slices."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fqz_oracle as oracle
    from oracle import gunzip_oracle
    from tests.gzip_cases import bgzf

    text = _text().tobytes()
    n = len(text)
    gz = bgzf(text, block=65280, level=1)
    members = sharding.walk_bgzf(gz)
    assert sum(m.text_size for m in members) == n and members[-1].text_size == 0  # the EOF block
    plan_m = sharding.plan_bgzf(members, world)
    slices = sharding.bgzf_slices(members, plan_m)
    (fa, fb), (a, b) = slices[rank]
    mine = gunzip_oracle.gunzip(gz[fa:fb]) if fb > fa else b""  # this rank's members only
    assert mine == text[a:b]
    local = (np.flatnonzero(np.frombuffer(mine, dtype=np.uint8) == 10) + a).tolist()
    counts = [None] * world
    dist.all_gather_object(counts, len(local))
    cand = sharding.block_cut_candidates(local, sum(counts[:rank]))
    allc = [None] * world
    dist.all_gather_object(allc, cand)
    plan = sharding.plan_compress(allc, n, world, bounds=[s[1] for s in slices])
    lo, hi = plan[rank]
    # the bytes in front of a rank's first block finish the last block of the rank before it
    heads = [None] * world
    dist.all_gather_object(heads, mine[: max(lo, a) - a] if hi > lo else mine)
    if hi > lo:
        chunk = mine[lo - a :]
        r = rank + 1
        while len(chunk) < hi - lo and r < world:
            chunk += heads[r]
            r += 1
        assert len(chunk) == hi - lo and chunk == text[lo:hi]
    flag = [None]
    if rank == 0:
        flag[0] = int(oracle.encode_streams(text[lo:hi], max_records=100000)["phred64"])
    dist.broadcast_object_list(flag, src=0)
    part = oracle.compress(text[lo:hi], threads=2) if hi > lo else b""
    if rank > 0 and part:
        part = part[10:]
    parts = [None] * world
    dist.all_gather_object(parts, part)
    if rank == 0:
        assert sharding.merge_compressed(parts) == oracle.compress(text, threads=2)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bgzf_input(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_bgzf_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_walk_bgzf_rejects_plain_gzip():
    import gzip

    from tests.gzip_cases import bgzf

    gz = bgzf(b"@r\nACGT\n+\nIIII\n" * 5000, block=4000)
    members = sharding.walk_bgzf(gz)
    assert len(members) == (5000 * 15 + 3999) // 4000 + 1 and members[1].text_offset == 4000
    for world in (1, 2, 3, 8, 64):
        plan = sharding.plan_bgzf(members, world)
        assert plan[0][0] == 0 and plan[-1][1] == len(members) and all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.walk_bgzf(gzip.compress(b"plain gzip has no BC field"))
    with pytest.raises(ValueError):
        sharding.walk_bgzf(gz[:-5])
