"""Decode throughput of a REFERENCE-SHAPED .fqz (oracle container: one libzstd level-1 frame per stream,
128 KiB blocks) on the device, next to a GPU-written file of the same records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fastqpacker_b200 as fq
from oracle import fqz_oracle as oracle

nrec = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ctx = fq.context(0)
text = oracle.synth(0, 0x5EED0001, 0, nrec)
n = text.size
t0 = time.perf_counter(); ref = oracle.compress_np(text, threads=os.cpu_count()); t_cpu = time.perf_counter() - t0
d_ref = torch.from_numpy(np.ascontiguousarray(ref)).cuda()
d_back = torch.empty(n + (1 << 16), dtype=torch.uint8, device="cuda")
d_text = torch.from_numpy(text).cuda()
d_out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
m = ctx.compress_device(d_text.data_ptr(), n, d_out.data_ptr(), d_out.numel())
for name, buf, size in (("reference-written", d_ref, ref.size), ("gpu-written", d_out, m)):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        k = ctx.decompress_device(buf.data_ptr(), size, d_back.data_ptr(), d_back.numel())
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert k == n and bool(torch.equal(d_back[:n], d_text))
    ctx.stats_reset(); ctx.profile(True)
    ctx.decompress_device(buf.data_ptr(), size, d_back.data_ptr(), d_back.numel())
    st = ctx.stats()["stages"]; ctx.profile(False)
    print(f"{name}: {size} bytes -> {n} bytes in {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s", {k2: round(v["ms"], 2) for k2, v in st.items() if v["ms"] > 0.05})
print(f"cpu oracle compress: {n/t_cpu/1e9:.2f} GB/s")
