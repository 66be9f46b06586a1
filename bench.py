#!/usr/bin/env python
"""bench.py — FASTQ GB/s of the fqz block codec on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--records R]

A "step" is one pass of the hot path over one batch of synthetic input: compress of R records of
synthetic Illumina-shaped FASTQ (BASELINE config 2: 25 000 000 records ~ 9 GB per GPU).  `value` is
device-resident compress throughput (input and output in HBM), `e2e` the same through the
host-buffer C-ABI call fqz_compress (pinned host buffers, H2D + D2H inside the timed region).
Decompress numbers (BASELINE config 3) ride along in "decompress".  Multi-GPU: one process per GPU,
blocks are independent so every rank codes its own records — weak scaling, no
data-path collective; timings are max over ranks.

--impl reference times the CPU restatement of the reference (oracle/, libzstd standing in for
klauspost/compress; the Go reference cannot be built in this image — SURVEY.md F7) on the box's
host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5EED0001
METRIC = "fastq_compress_throughput"
UNIT = "GB/s"
FULL_RECORDS = 25_000_000  # BASELINE config 2: ~9 GB of 150 bp Phred+33 reads
CPU_SAMPLE_RECORDS = 3_000_000  # ~1.1 GB: a few seconds on all host cores


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _bind_to_gpu_numa_node(index: int):
    """Pin this process (and the pinned host buffers it is about to allocate, by first touch) to the CPUs
    NVML reports as local to the GPU: with 8 ranks streaming ~50 GB/s each through host memory the
    end-to-end numbers are decided by which socket that memory sits on."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f,
                stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {
            "sm_mhz": statistics.median(busy),
            "sm_max_mhz": max(mx),
            "power_w_max": max(pw),
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: oracle (restated reference + libzstd level 1 + frame checksum) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the other ranks exit 0 without work
    from oracle import fqz_oracle as oracle

    cores = os.cpu_count() or 1
    nrec = min(args.records or FULL_RECORDS, args.cpu_records or CPU_SAMPLE_RECORDS)
    text = oracle.synth(0, SEED, 0, nrec)
    n = text.size
    fqz = None
    for _ in range(args.warmup if args.warmup < 2 else 1):  # one warm-up pass is plenty on the CPU
        fqz = oracle.compress_np(text, threads=cores)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fqz = oracle.compress_np(text, threads=cores)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    t0 = time.perf_counter()
    back = oracle.decompress_mt(fqz, cores, n + 4096)
    ddt = time.perf_counter() - t0
    assert back.size == n
    val = n / dt / 1e9
    sample = f"{nrec} records ({n / 1e9:.2f} GB) of the config-2 generator per step, oracle compress with {cores} threads"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": val,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "synthetic Illumina 150 bp Phred+33 reads shaped like ERR532393_1 (BASELINE config 2), compress", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "decompress_value": n / ddt / 1e9, "ratio": n / fqz.size,
                         "note": "restated CPU baseline (oracle + libzstd-1), not fqpack: no Go toolchain in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fastqpacker_b200 as fq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fq.context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=local)
    nrec = args.records or FULL_RECORDS
    peak, peak_src = _peaks()

    # ---- synthetic input, generated on the device (rank r owns records [r*nrec, (r+1)*nrec))
    cap = nrec * 372 + (1 << 20)
    d_in = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = 0
    step = 4_000_000
    for first in range(0, nrec, step):
        c = min(step, nrec - first)
        n += ctx.synth_device(0, SEED, rank * nrec + first, c, d_in.data_ptr() + n, cap - n)
    d_out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n + (1 << 16), dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """K calls bracketed by CUDA events on the library's stream; returns max-over-ranks seconds."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(steps):
            fn()
        e1.record(lib_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item()) / 1e3

    res = {}

    def do_compress():
        res["m"] = ctx.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())

    def do_decompress():
        res["k"] = ctx.decompress_device(d_out.data_ptr(), res["m"], d_back.data_ptr(), d_back.numel())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident compress (the headline `value`)
    ctx.stats_reset()
    t_c = timed(do_compress, args.steps, args.warmup)
    launches = ctx.stats()["launches"]  # includes warm-up
    launches_per_step = launches // (args.steps + args.warmup)
    # ---- device-resident decompress of the GPU-written file
    t_d = timed(do_decompress, args.steps, args.warmup)
    assert res["k"] == n
    ok = bool(torch.equal(d_back[:n], d_in[:n]))  # size-independent property: round trip at full size
    # ---- per-stage device times (CUDA events inside the library, same stream), one profiled pass each
    ctx.stats_reset()
    ctx.profile(True)
    do_compress()
    do_decompress()
    stages = ctx.stats()["stages"]
    ctx.profile(False)

    # ---- end to end through the host-buffer C-ABI calls (pinned host memory)
    e2e = None
    e2e_d = None
    m = res["m"]
    if not args.no_e2e:
        import numpy as np

        # pinned host memory is held one direction at a time (8 ranks share the box's RAM): input + output for
        # compress, then output + FASTQ for decompress; a head / tail sample of the input stays for the check
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in[:n])
        h_out = torch.empty(m + m // 8 + (1 << 20), dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        a_in, a_out = h_in.numpy(), h_out.numpy()
        head, tail = a_in[:4096].copy(), a_in[n - 4096 : n].copy()

        def e2e_compress():
            res["em"] = ctx.compress_into(a_in, a_out)

        e_steps = max(1, min(args.steps, args.e2e_steps))
        t_ec = timed(e2e_compress, e_steps, min(args.warmup, 3))
        del a_in, h_in
        h_back = torch.empty(n + (1 << 16), dtype=torch.uint8, pin_memory=True)
        a_back = h_back.numpy()

        def e2e_decompress():
            res["ek"] = ctx.decompress_into(a_out[: res["em"]], a_back)

        t_ed = timed(e2e_decompress, e_steps, min(args.warmup, 3))
        assert res["ek"] == n and np.array_equal(a_back[:4096], head) and np.array_equal(a_back[n - 4096 : n], tail)
        assert bool(torch.equal(torch.from_numpy(a_back[:n]).cuda(), d_in[:n])), "end-to-end round trip differs from the input"
        e2e = {"value": world * n * e_steps / t_ec / 1e9, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": res["em"],
               "steps": e_steps, "api": "fqz_compress (host buffers, pinned)"}
        e2e_d = {"value": world * n * e_steps / t_ed / 1e9, "unit": UNIT, "h2d_bytes_per_step": res["em"], "d2h_bytes_per_step": n,
                 "steps": e_steps, "api": "fqz_decompress (host buffers, pinned)"}
        del a_back, h_back, a_out, h_out
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- compressed bytes per stream class (block headers of the container, container.go:97-109): the
    #      algorithmic bytes of an entropy kernel are stream bytes read + compressed bytes written
    import struct

    hdrs = d_out[:m].cpu().numpy()
    z_ent = z_lz = 0
    pos = 10
    while pos < m:
        v = struct.unpack_from("<9I", hdrs, pos)
        z_ent += v[1] + v[2]
        z_lz += v[3] + v[4] + v[5] + v[6]
        pos += 36 + sum(v[1:7])
    del hdrs
    if "zstd_enc_lz" in stages:
        stages["zstd_enc_lz"]["bytes"] += z_lz
    if "zstd_enc_entropy" in stages:
        stages["zstd_enc_entropy"]["bytes"] += z_ent
    # ---- roofline of the dominant kernel (largest share of the compress step)
    comp_stages = ["newline_count", "newline_index", "scan", "record_meta", "scatter_streams", "zstd_enc_entropy", "zstd_enc_lz", "xxh64", "assemble", "copy"]
    cs = {k: v for k, v in stages.items() if k in comp_stages}
    top = max(cs, key=lambda k: cs[k]["ms"]) if cs else None
    roof = None
    if top:
        v = cs[top]
        # algorithmic bytes of the entropy kernels: stream bytes read + compressed bytes written (SURVEY §8d A2 = S + Z)
        ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0  # bytes per launch / time per launch
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f).get(top)
            if t:  # DRAM bytes of one launch over one full window, from the committed ncu --set full capture
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
                traffic_src = t["source"]
        except OSError:
            pass
        # a "launch" of the stage = its kernels over one device window
        nl = max(1, v["launches"] // {"zstd_enc_lz": 3, "assemble": 2}.get(top, 1))
        stage_kernels = {
            "zstd_enc_lz": "k_zitems_parse + k_zenc<2,1> + k_zenc<2,2> (item streams: matcher, Huffman literals, FSE sequences)",
            "zstd_enc_entropy": "k_zenc_huf (packed bases, qualities: Huffman frames)",
            "scatter_streams": "k_scatter_streams",
            "record_meta": "k_record_meta",
            "newline_count": "k_newline_count",
            "newline_index": "k_newline_index",
            "xxh64": "k_xxh64_frames",
            "assemble": "k_zassemble + k_write_block_headers",
            "scan": "k_scan_partial + k_scan_apply + k_zindex",
        }
        roof = {"bound": "hbm", "kernel": top, "kernels": stage_kernels.get(top, top), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": v["bytes"] / nl, "ms_per_launch": v["ms"] / nl,
                "peak_source": peak_src, "launches_per_step": nl, "kernel_launches_per_step": v["launches"], "kernel_ms_per_step": v["ms"],
                "share_of_step": v["ms"] / sum(x["ms"] for x in cs.values())}
    step_s = t_c / args.steps
    value = world * n * args.steps / t_c / 1e9
    pipeline = {"algorithmic_bytes_per_step": n + m, "achieved": (n + m) / step_s / 1e9, "peak": peak, "frac": (n + m) / step_s / 1e9 / peak,
                "note": "whole compress step: A = F + Z over device time (SURVEY §8d)"}
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": step_s * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": {
            "workload": "synthetic Illumina 150 bp Phred+33 reads shaped like ERR532393_1 (BASELINE config 2), compress at 1 B200 per rank",
            "records_per_gpu": nrec,
            "fastq_bytes_per_gpu": n,
            "fqz_bytes_per_gpu": m,
            "ratio": n / m,
            "blocks_per_gpu": (nrec + 99999) // 100000,
            "l2": "inputs (GBs) far larger than the 126 MB L2; no flush needed",
            "round_trip_ok": ok,
            "cpus_bound_to_gpu_numa_node": numa,
        },
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof,
        "pipeline_roofline": pipeline,
        "stages_ms_per_step": {k: round(v["ms"], 3) for k, v in stages.items()},
        "decompress": {"value": world * n * args.steps / t_d / 1e9, "unit": UNIT, "ms_per_step": t_d / args.steps * 1e3, "e2e": e2e_d,
                       "input": "GPU-written .fqz"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        cb, text, ref_fqz = cpu_baseline(args)
        line["cpu_baseline"] = cb
        # BASELINE config 3, second half: the REFERENCE-SHAPED file of the same sample (oracle container: one
        # libzstd level-1 frame per stream, 128 KiB blocks) decoded on the device.  Such frames are one serial
        # chain each: correctness path, reported for completeness.
        d_ref = torch.from_numpy(ref_fqz).cuda()
        d_txt = torch.empty(text.size + (1 << 16), dtype=torch.uint8, device="cuda")
        ctx.decompress_device(d_ref.data_ptr(), d_ref.numel(), d_txt.data_ptr(), d_txt.numel())
        t_r = timed(lambda: ctx.decompress_device(d_ref.data_ptr(), d_ref.numel(), d_txt.data_ptr(), d_txt.numel()), 1, 0)
        same = bool(torch.equal(d_txt[: text.size], torch.from_numpy(text).cuda()))
        line["decompress"]["reference_written"] = {"value": text.size / t_r / 1e9, "unit": UNIT, "bit_exact": same, "sample": cb["sample"],
                                                   "input": "oracle-written .fqz (reference-shaped: one libzstd-1 frame per stream)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    from oracle import fqz_oracle as oracle

    cores = os.cpu_count() or 1
    nrec = min(args.records or FULL_RECORDS, args.cpu_records or CPU_SAMPLE_RECORDS)
    text = oracle.synth(0, SEED, 0, nrec)
    n = text.size
    oracle.compress_np(text[: n // 8], threads=cores)  # warm the library
    t0 = time.perf_counter()
    fqz = oracle.compress_np(text, threads=cores)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    back = oracle.decompress_mt(fqz, cores, n + 4096)
    ddt = time.perf_counter() - t0
    assert back.size == n
    return {
        "value": n / dt / 1e9,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"first {nrec} records ({n / 1e9:.2f} GB) of the same generator, one pass, {cores} threads",
        "decompress_value": n / ddt / 1e9,
        "ratio": n / fqz.size,
        "note": "restated CPU baseline (oracle + libzstd level 1 + frame checksum), not fqpack: no Go toolchain in this image (SURVEY F7)",
    }, text, fqz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=0, help="records per GPU (default: BASELINE config 2, 25 000 000)")
    ap.add_argument("--cpu-records", type=int, default=0, help="records of the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, ...) write there too, so
    # everything but that line is sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
