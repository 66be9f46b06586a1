"""Probe of the host-buffer calls: wall time, per-stage device time, plain copy rates."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastqpacker_b200 as fq

ctx = fq.context(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
cap = count * 372 + (1 << 20)
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
n = 0
for first in range(0, count, 4000000):
    c = min(4000000, count - first)
    n += ctx.synth_device(0, 0x5EED0001, first, c, buf.data_ptr() + n, cap - n)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in.copy_(buf[:n])
h_out = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True)
h_back = torch.empty(n + 4096, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
# plain copy rates
for name, fn in (("h2d", lambda: buf[:n].copy_(h_in, non_blocking=True)), ("d2h", lambda: h_back[:n].copy_(buf[:n], non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {n/dt/1e9:.1f} GB/s")
a_in, a_out, a_back = h_in.numpy(), h_out.numpy(), h_back.numpy()
for it in range(3):
    m = ctx.compress_into(a_in, a_out)
for prof in (False, True):
    ctx.stats_reset(); ctx.profile(prof)
    t0 = time.perf_counter(); m = ctx.compress_into(a_in, a_out); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"compress e2e profile={prof}: {dt*1e3:.1f} ms = {n/dt/1e9:.1f} GB/s; launches {st['launches']}; stage sum {sum(v['ms'] for v in st['stages'].values()):.1f} ms")
    if prof:
        print({k: round(v['ms'], 2) for k, v in st['stages'].items()})
ctx.profile(False)
for it in range(2):
    k = ctx.decompress_into(a_out[:m], a_back)
for prof in (False, True):
    ctx.stats_reset(); ctx.profile(prof)
    t0 = time.perf_counter(); k = ctx.decompress_into(a_out[:m], a_back); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"decompress e2e profile={prof}: {dt*1e3:.1f} ms = {n/dt/1e9:.1f} GB/s; stage sum {sum(v['ms'] for v in st['stages'].values()):.1f} ms")
    if prof:
        print({k2: round(v['ms'], 2) for k2, v in st['stages'].items()})
