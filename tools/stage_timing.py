"""Per-stage device times (CUDA events inside the library) of a device-resident compress."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fastqpacker_b200 as fq

ctx = fq.context(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
cap = count * 380
buf = torch.empty(cap + 4096, dtype=torch.uint8, device="cuda")
n = 0
step = 4000000
for first in range(0, count, step):
    c = min(step, count - first)
    n += ctx.synth_device(0, 0x5EED0001, first, c, buf.data_ptr() + n, cap - n)
out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
iters = 3
for it in range(2):
    m = ctx.compress_device(buf.data_ptr(), n, out.data_ptr(), out.numel())
torch.cuda.synchronize()
ctx.stats_reset()
ctx.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(iters):
    m = ctx.compress_device(buf.data_ptr(), n, out.data_ptr(), out.numel())
e1.record()
torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / iters
st = ctx.stats()
print(f"records {count}  fastq {n/1e6:.1f} MB -> fqz {m/1e6:.1f} MB  ratio {n/m:.3f}   {wall:.2f} ms/iter = {n/wall/1e6:.1f} GB/s FASTQ (profiling on)")
tot = 0
for k, v in st["stages"].items():
    ms = v["ms"] / iters
    tot += ms
    gbs = (v["bytes"] / iters) / (ms * 1e-3) / 1e9 if ms > 0 and v["bytes"] else 0
    print(f"{k:18s} {ms:9.3f} ms/iter  launches {v['launches']/iters:6.1f}  bytes {v['bytes']/iters/1e6:9.2f} MB  {gbs:8.1f} GB/s")
print(f"sum of stages {tot:.3f} ms")
ctx.profile(False)
torch.cuda.synchronize()
e0.record()
for it in range(iters):
    m = ctx.compress_device(buf.data_ptr(), n, out.data_ptr(), out.numel())
e1.record()
torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / iters
print(f"profiling off: {wall:.2f} ms/iter = {n/wall/1e6:.1f} GB/s FASTQ")
