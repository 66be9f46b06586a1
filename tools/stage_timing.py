"""Prints per-stage device times (CUDA events inside the library) for one block of synthetic data."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastqpacker_b200 as fq

ctx = fq.context(0)
count = 100000
buf = torch.empty(count * 400, dtype=torch.uint8, device="cuda")
n = ctx.synth_device(0, 0x5EED0001, 0, count, buf.data_ptr(), buf.numel())
text = buf[:n].cpu().numpy()
for it in range(3):
    ctx.encode_streams(text)
ctx.stats_reset()
ctx.profile(True)
for it in range(5):
    ctx.encode_streams(text)
st = ctx.stats()
for k, v in st["stages"].items():
    ms = v["ms"] / 5
    gbs = (v["bytes"] / 5) / (ms * 1e-3) / 1e9 if ms > 0 and v["bytes"] else 0
    print(f"{k:18s} {ms:8.3f} ms/iter  launches {v['launches']/5:5.1f}  bytes {v['bytes']/5/1e6:8.2f} MB  {gbs:8.1f} GB/s")
print(json.dumps(st))
