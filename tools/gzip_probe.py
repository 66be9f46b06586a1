"""One-GPU probe of the gzip input stage (bench.py's `gzip_input` side workload alone, with per-stage device times).
usage: python tools/gzip_probe.py [records] > gpurun_out/gzip_probe.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import bench
    import fastqpacker_b200 as fq

    nrec = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ctx = fq.context(0)
    lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=0)
    cap = nrec * 372 + (1 << 20)
    d_in = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ctx.synth_device(0, bench.SEED, 0, nrec, d_in.data_ptr(), cap)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(steps):
            fn()
        e1.record(lib_stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3

    ctx.profile(True)
    ctx.stats_reset()
    out = bench.gzip_workload(ctx, torch, d_in, nrec, timed)
    st = ctx.stats()["stages"]
    out["stage_ms_total"] = {k: round(v["ms"], 3) for k, v in st.items() if k.startswith("gz_")}
    out["stage_launches"] = {k: v["launches"] for k, v in st.items() if k.startswith("gz_")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
