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

Extra keys of the N=1 line: "cfg4" (BASELINE config 4: 2 M Phred+64 / N-heavy / '+' payload / 50-300 bp
records), "duplicates" (3 M reads of which 35 % are exact copies: GPU ratio next to the CPU path's),
"decompress.reference_written" (config 3: the FULL 25 M-record file written by the CPU path, decoded on
the GPUs — at N > 1 every rank decodes the block run sharding.plan_decompress gives it).

--workload cfg5 (BASELINE config 5, N >= 2 under torchrun): ONE logical 64 GB input spread over the ranks
as even slices that do not end on block boundaries; per-rank newline counts are exchanged, the slices are
re-cut on block boundaries (sharding.py), every rank codes its blocks, and the parts are gathered on the
host inside the timed region.

--impl reference times the CPU restatement of the reference (oracle/, libzstd standing in for
klauspost/compress; the Go reference cannot be built in this image — SURVEY.md F7) on the box's
host cores, on the SAME 25 M-record input (fewer timed passes than --steps asks for when a pass
takes seconds; "steps_timed" says how many).
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
CFG4_RECORDS = 2_000_000  # BASELINE config 4
CFG4_SEED = 0x5EED0004
GZIP_RECORDS = 1_000_000  # gzip input side workload (row f3)
DUP_RECORDS = 3_000_000
CFG5_BYTES = 64_000_000_000  # BASELINE config 5
WORKLOAD = "synthetic Illumina 150 bp Phred+33 reads shaped like ERR532393_1 (BASELINE config 2), compress at 1 B200 per rank"


def _synth_host(kind, seed, first, count, threads):
    """The CPU twin of the generator on `threads` threads (ctypes releases the GIL) -> one numpy array."""
    import concurrent.futures as cf

    import numpy as np

    from oracle import fqz_oracle as oracle

    step = 250_000
    jobs = [(f, min(step, first + count - f)) for f in range(first, first + count, step)]
    with cf.ThreadPoolExecutor(max(1, threads)) as ex:
        parts = list(ex.map(lambda j: oracle.synth(kind, seed, j[0], j[1]), jobs))
    return np.concatenate(parts) if len(parts) > 1 else parts[0]


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
    """CPU arm: oracle (restated reference + libzstd level 1 + frame checksum) on all host cores, on the same
    input as the GPU arm (config 2, 25 M records)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the other ranks exit 0 without work
    from oracle import fqz_oracle as oracle

    cores = os.cpu_count() or 1
    nrec = args.records or FULL_RECORDS
    text = _synth_host(0, SEED, 0, nrec, cores)
    n = text.size
    t0 = time.perf_counter()
    fqz = oracle.compress_np(text, threads=cores)  # warm-up pass (also sizes the timed ones)
    one = time.perf_counter() - t0
    # a pass over 9.2 GB takes seconds on the host: bound the whole run to about a minute of timed work
    steps = max(1, min(args.steps, int(60.0 / max(one, 1e-3))))
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fqz = oracle.compress_np(text, threads=cores)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    t0 = time.perf_counter()
    back = oracle.decompress_mt(fqz, cores, n + 4096)
    ddt = time.perf_counter() - t0
    assert back.size == n
    val = n / dt / 1e9
    sample = f"all {nrec} records ({n / 1e9:.2f} GB) of the config-2 generator per step, oracle compress with {cores} threads, {steps} timed passes"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": val,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "steps_timed": steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "records_per_gpu": nrec, "fastq_bytes_per_gpu": int(n), "fqz_bytes_per_gpu": int(fqz.size),
                   "ratio": n / fqz.size, "blocks_per_gpu": (nrec + 99999) // 100000},
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
    ctx.set_option(ctx.OPT_FRONTEND, args.frontend)
    ctx.set_option(ctx.OPT_HUF_KERNELS, args.huf_kernels)
    if args.window_bytes:
        ctx.set_option(ctx.OPT_WINDOW_BYTES, args.window_bytes)
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
    e2e_f = None
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

        # the same through the STREAMING calls the Go shim uses: 2 GiB windows fed one after the
        # other, the unconsumed tail of a call (less than one block) in front of the next window
        FEED = 2 << 30

        def e2e_feed():
            cs = ctx.compress_stream()
            pos = outpos = 0
            while True:
                end = min(n, pos + FEED)
                mm, used = cs.feed(a_in[pos:end], end == n, a_out[outpos:])
                outpos += mm
                pos += used
                if end == n:
                    break
            cs.close()
            res["fm"] = outpos

        t_ef = timed(e2e_feed, e_steps, 1)
        assert res["fm"] == res["em"], "the streaming calls wrote a different file"
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
        e2e_f = {"value": world * n * e_steps / t_ef / 1e9, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": res["fm"], "steps": e_steps,
                 "api": "fqz_compress_begin / fqz_compress_feed (2 GiB windows, pinned) / fqz_compress_end: the loop of INTEGRATION.md's shim"}
        del a_back, h_back, a_out, h_out
    # ---- BASELINE config 3, second half: the REFERENCE-WRITTEN file of the whole workload (rank 0's 25 M records,
    #      one libzstd level-1 frame per stream as the reference writes them), decoded by all ranks together
    refw, cb = None, None
    if not args.no_cpu:
        refw, cb = reference_written(args, ctx, torch, dist, rank, world, nrec, d_in, n, d_back, timed)
    extras = {}
    if world == 1 and not args.no_extras:
        extras["cfg4"] = side_workload(ctx, torch, 1, CFG4_SEED, min(CFG4_RECORDS, nrec), 760, timed, None,
                                       "synthetic Phred+64, N-heavy (5 % N), '+' payloads, 50-300 bp (BASELINE config 4)")
        extras["duplicates"] = side_workload(ctx, torch, 2, SEED, min(DUP_RECORDS, nrec), 372, timed, os.cpu_count() or 1,
                                             "config-2 reads of which 35 % are exact copies of one of the 400 reads in front (generator kind 2)")
        try:
            extras["gzip_input"] = gzip_workload(ctx, torch, d_in, min(GZIP_RECORDS, nrec), timed)
        except Exception as e:  # the side workload must never cost the headline line
            extras["gzip_input"] = {"error": repr(e)[:300]}
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
    comp_stages = ["newline_count", "newline_index", "scan", "record_meta", "scatter_streams", "zstd_enc_entropy", "zstd_enc_lz", "zstd_enc_dup", "xxh64", "assemble", "copy"]
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
            "zstd_enc_entropy": "k_zh_hist + k_zh_plan + k_zh_encode (packed bases, qualities: Huffman frames)",
            "zstd_enc_dup": "k_rec_ranges + k_rec_keys + k_rec_detect + k_rec_match + k_xxh64_streams + k_lzrec_* (duplicate-record search and coding)",
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
            "workload": WORKLOAD,
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
        "e2e_feed": e2e_f,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof,
        "pipeline_roofline": pipeline,
        "stages_ms_per_step": {k: round(v["ms"], 3) for k, v in stages.items()},
        "stages_note": "per-stage device times of ONE profiled pass in which the stages run one after the other; in the timed steps the "
                       "item-stream kernels (zstd_enc_lz) run on a second stream beside xxh64 / zstd_enc_dup / zstd_enc_entropy, so the "
                       "stage times add up to more than ms_per_step",
        "decompress": {"value": world * n * args.steps / t_d / 1e9, "unit": UNIT, "ms_per_step": t_d / args.steps * 1e3, "e2e": e2e_d,
                       "input": "GPU-written .fqz"},
        "clocks": clocks,
    }
    if cb is not None and world == 1:
        line["cpu_baseline"] = cb
    if refw is not None:
        line["decompress"]["reference_written"] = refw
    line.update(extras)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_cfg5(args):
    """BASELINE config 5: ONE logical input (default 64 GB) block-sharded over the ranks.

    Rank r holds the even slice of the file's records [R*r/N, R*(r+1)/N) — slices end on record but not on
    block boundaries — in pinned HOST memory.  A timed end-to-end step is the whole flow a multi-GPU host runs:
      H2D of the slice -> newline count on the GPU -> counts exchanged (the one tiny exchange) -> every rank finds
      the first block cut inside its slice -> the bytes in front of it go to the rank before (they finish ITS last
      block; NCCL send/recv, at most one block) -> fqz_compress_shard_device on the block-aligned range (rank 0
      decides the Phred flag on block 0 first and the flag travels with the cuts) -> D2H -> ordered gather of the
      parts into one host file (/dev/shm).
    "value" is the device-resident part (count ... compress), "e2e" everything.  Decompress: the gathered file is
    walked once, plan_decompress hands every rank a run of blocks, each rank decodes its run (H2D + D2H inside
    the e2e figure) and compares with its slice."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import fastqpacker_b200 as fq
    from fastqpacker_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fq.context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=local)
    peak, peak_src = _peaks()
    total_bytes = args.total_bytes or CFG5_BYTES
    R = int(total_bytes / 368.03)  # records of the logical file (config-2 records average 368.03 bytes)
    r0, r1 = R * rank // world, R * (rank + 1) // world
    nrec = r1 - r0
    LPB = sharding.LINES_PER_BLOCK

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_ints(vals):
        t = torch.tensor(vals, dtype=torch.int64, device="cuda")
        if world == 1:
            return [list(vals)]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    # ---- the rank's slice: generated on the device, parked in pinned host memory (one block of slack for the borrowed tail)
    slack = 48 << 20
    cap = nrec * 372 + (1 << 20)
    d_slice = torch.empty(cap + slack, dtype=torch.uint8, device="cuda")
    n = 0
    for first in range(0, nrec, 4_000_000):
        c = min(4_000_000, nrec - first)
        n += ctx.synth_device(0, SEED, r0 + first, c, d_slice.data_ptr() + n, cap - n)
    h_slice = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_slice.copy_(d_slice[:n])
    d_out = torch.empty(n // 2 + (64 << 20), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    out_path = _shared_path(f"fqz_bench_cfg5_{os.environ.get('MASTER_PORT', '0')}.fqz")
    st = {}

    def plan_and_compress():
        """device-resident part: count, exchange, borrow, compress.  Leaves the part in d_out[:st['m']]."""
        lines = ctx.count_lines_device(d_slice.data_ptr(), n)
        ph = -1
        if rank == 0:  # block 0 decides the Phred flag of the file (compress.go:146-164): code it alone first, read the flag
            e0 = ctx.find_line_end_device(d_slice.data_ptr(), n, min(LPB, lines)) + 1 if lines else 0
            _, ph = ctx.compress_shard_device(d_slice.data_ptr(), e0, d_out.data_ptr(), d_out.numel(), -1, True)
        allv = gather_ints([lines, ph])
        counts = [v[0] for v in allv]
        ph = allv[0][1]
        before = sum(counts[:rank])
        k = (LPB - before % LPB) % LPB  # this slice's k-th newline ends the block that is open at its start
        if k == 0:
            start = 0
        elif k <= lines:
            start = ctx.find_line_end_device(d_slice.data_ptr(), n, k) + 1
        else:
            start = n  # no cut inside the slice: all of it belongs to the rank before (tiny slices only)
        starts = [v[0] for v in gather_ints([start])]
        # the head of slice r+1 finishes the last block of rank r
        give = starts[rank] if rank > 0 else 0
        take = starts[rank + 1] if rank + 1 < world else 0
        assert take <= slack, "borrowed tail larger than one block of slack"
        if world > 1 and (give or take):
            ops = []
            if give:
                ops.append(dist.P2POp(dist.isend, d_slice[:give], rank - 1))
            if take:
                ops.append(dist.P2POp(dist.irecv, d_slice[n : n + take], rank + 1))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            torch.cuda.synchronize()
        st["start"], st["take"] = start, take
        m, _ = ctx.compress_shard_device(d_slice.data_ptr() + start, n - start + take, d_out.data_ptr(), d_out.numel(), ph, rank == 0)
        st["m"] = m

    def e2e_compress():
        with torch.cuda.stream(lib_stream):
            d_slice[:n].copy_(h_slice, non_blocking=True)
        plan_and_compress()
        m = st["m"]
        sizes = [v[0] for v in gather_ints([m])]  # where this part goes in the file
        assert sizes == st["sizes"], "part sizes changed between steps"
        off = sum(sizes[:rank])
        if direct:
            with torch.cuda.stream(lib_stream):
                h_file[off : off + m].copy_(d_out[:m], non_blocking=True)
            lib_stream.synchronize()
        else:  # page-locking the shared file failed: stage through this rank's pinned buffer
            with torch.cuda.stream(lib_stream):
                h_out[:m].copy_(d_out[:m], non_blocking=True)
            lib_stream.synchronize()
            mm[off : off + m] = h_out[:m].numpy()

    def timed_wall(fn, steps, warmup):
        """end-to-end steps hold host work and exchanges: wall clock between barriers, max over ranks"""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 2))
    ctx.stats_reset()
    t_dev = timed_wall(plan_and_compress, steps, warm)
    launches = ctx.stats()["launches"] // (steps + warm)
    # the ordered collector's output file (collectAndWriteResults, compress.go:365-403): one shared host file that every
    # rank maps; a rank page-locks the pages its own part lands in (the sizes are known from the passes above and do
    # not change), so that the part goes from HBM straight to its place in the file — no second host copy
    st["sizes"] = [v[0] for v in gather_ints([st["m"]])]
    total_out = sum(st["sizes"])
    if rank == 0:
        with open(out_path, "wb") as f:
            f.truncate(total_out)
    barrier()
    mm = np.memmap(out_path, dtype=np.uint8, mode="r+")
    h_file = torch.from_numpy(mm)
    my_off, my_m = sum(st["sizes"][:rank]), st["m"]
    page = 4096
    reg_lo = (mm.ctypes.data + my_off) & ~(page - 1)
    reg_hi = (mm.ctypes.data + my_off + my_m + page - 1) & ~(page - 1)
    direct = False
    try:
        direct = int(torch.cuda.cudart().cudaHostRegister(reg_lo, reg_hi - reg_lo, 0)) == 0
    except Exception:
        direct = False
    if not direct:
        try:
            torch.cuda.cudart().cudaGetLastError()  # a refused registration must not poison the next CUDA call
        except Exception:
            pass
    h_out = None if direct else torch.empty(my_m + (1 << 20), dtype=torch.uint8, pin_memory=True)
    t_e2e = timed_wall(e2e_compress, steps, 1)
    sizes = st["sizes"]
    fqz_total = sum(sizes)
    totals = gather_ints([n])
    file_bytes = sum(v[0] for v in totals)

    # ---- decompress: one walk of the gathered file, block runs per rank
    barrier()
    del h_slice, h_out
    fqz = mm[:fqz_total]
    _, _, blocks = sharding.walk_container(memoryview(fqz))
    first, last = sharding.plan_decompress(blocks, world)[rank]
    a = blocks[first].offset if last > first else 10
    b = blocks[last - 1].offset + blocks[last - 1].size if last > first else 10
    h_part = torch.empty(10 + b - a, dtype=torch.uint8, pin_memory=True)
    h_part.numpy()[:10] = fqz[:10]
    h_part.numpy()[10:] = fqz[a:b]
    nblocks, nrecords = len(blocks), sum(x.records for x in blocks)
    rec_lo = sum(x.records for x in blocks[:first])
    rec_n = sum(x.records for x in blocks[first:last])
    del fqz
    d_part = torch.empty(h_part.numel() + 64, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(rec_n * 372 + (1 << 20), dtype=torch.uint8, device="cuda")
    h_back = torch.empty(d_back.numel(), dtype=torch.uint8, pin_memory=True)
    dd = {}

    def dev_decompress():
        dd["k"] = ctx.decompress_device(d_part.data_ptr(), h_part.numel(), d_back.data_ptr(), d_back.numel())

    def e2e_decompress():
        with torch.cuda.stream(lib_stream):
            d_part[: h_part.numel()].copy_(h_part, non_blocking=True)
        dev_decompress()
        with torch.cuda.stream(lib_stream):
            h_back[: dd["k"]].copy_(d_back[: dd["k"]], non_blocking=True)
        lib_stream.synchronize()

    d_part[: h_part.numel()].copy_(h_part)
    t_ddev = timed_wall(dev_decompress, steps, 1)
    t_de2e = timed_wall(e2e_decompress, steps, 1)
    # what the run must decode to: records [rec_lo, rec_lo + rec_n) of the logical file
    del d_slice
    d_want = torch.empty(rec_n * 372 + (1 << 20), dtype=torch.uint8, device="cuda")
    want = 0
    for f0 in range(rec_lo, rec_lo + rec_n, 4_000_000):
        c = min(4_000_000, rec_lo + rec_n - f0)
        want += ctx.synth_device(0, SEED, f0, c, d_want.data_ptr() + want, d_want.numel() - want)
    same = dd["k"] == want and bool(torch.equal(d_back[:want], d_want[:want])) and bool(torch.equal(h_back[:want].cuda(), d_want[:want]))
    ok = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    if direct:
        try:
            torch.cuda.cudart().cudaHostUnregister(reg_lo)
        except Exception:
            pass
    if rank == 0:
        try:
            os.unlink(out_path)
        except OSError:
            pass
        gbs = lambda t: file_bytes * steps / t / 1e9
        line = {
            "metric": METRIC, "value": gbs(t_dev), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": t_dev / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "one synthetic config-2 FASTQ block-sharded over the GPUs (BASELINE config 5), compress + decompress",
                       "fastq_bytes": file_bytes, "records": nrecords, "blocks": nblocks, "fqz_bytes": fqz_total, "ratio": file_bytes / fqz_total,
                       "slices": "even record ranges per rank, re-cut on block boundaries from exchanged newline counts (sharding.py)",
                       "l2": "inputs (GBs) far larger than the 126 MB L2; no flush needed", "round_trip_ok": bool(ok.item() == 1.0),
                       "timing": "wall clock between barriers, max over ranks (the steps hold exchanges and host work)",
                       "cpus_bound_to_gpu_numa_node": numa},
            "e2e": {"value": gbs(t_e2e), "unit": UNIT, "h2d_bytes_per_step": file_bytes, "d2h_bytes_per_step": fqz_total, "steps": steps,
                    "api": "H2D slice, fqz_count_lines_device, exchange, fqz_compress_shard_device, D2H into the part's place in one shared host file",
                    "gather": "direct (file page-locked by every rank)" if direct else "staged through a pinned buffer"},
            "gpu_launches": launches * steps,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": (file_bytes + fqz_total) * steps / t_dev / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": (file_bytes + fqz_total) * steps / t_dev / 1e9 / world / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per GPU: (F + Z) / device time of the whole sharded step; per-kernel rooflines are in the default workload's line"},
            "decompress": {"value": gbs(t_ddev), "unit": UNIT, "ms_per_step": t_ddev / steps * 1e3,
                           "e2e": {"value": gbs(t_de2e), "unit": UNIT, "h2d_bytes_per_step": fqz_total, "d2h_bytes_per_step": file_bytes, "steps": steps,
                                   "api": "H2D block run, fqz_decompress_device, D2H (per rank; the ordered write of the FASTQ is the host's)"},
                           "input": "the gathered GPU-written .fqz, block runs from sharding.plan_decompress"},
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _shared_path(name):
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    return os.path.join(d, name)


def reference_written(args, ctx, torch, dist, rank, world, nrec, d_in, n, d_back, timed):
    """Rank 0 times the CPU path on the box's host cores over the WHOLE workload (= cpu_baseline) and leaves the
    file it wrote where all ranks can read it; every rank then decodes the run of blocks plan_decompress gives it
    and compares with the generator.  Returns (reference_written dict, cpu_baseline dict) on rank 0."""
    import numpy as np

    from fastqpacker_b200 import sharding
    from oracle import fqz_oracle as oracle

    cores = os.cpu_count() or 1
    path = _shared_path(f"fqz_bench_ref_{os.environ.get('MASTER_PORT', '0')}_{nrec}.fqz")
    cb = None
    if rank == 0:
        text = d_in[:n].cpu().numpy()  # rank 0 owns records [0, nrec): the device generator is the CPU twin
        oracle.compress_np(text[: max(1, n // 64)], threads=cores)  # warm the library
        t0 = time.perf_counter()
        fqz = oracle.compress_np(text, threads=cores)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        back = oracle.decompress_mt(fqz, cores, n + 4096)
        ddt = time.perf_counter() - t0
        assert back.size == n
        del back
        fqz.tofile(path)
        cb = {"value": n / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": f"all {nrec} records ({n / 1e9:.2f} GB) of the workload, one pass, {cores} threads",
              "decompress_value": n / ddt / 1e9, "ratio": n / fqz.size,
              "note": "restated CPU baseline (oracle + libzstd level 1 + frame checksum), not fqpack: no Go toolchain in this image (SURVEY F7)"}
        del text, fqz
    if world > 1:
        dist.barrier()
    fqz = np.fromfile(path, dtype=np.uint8)
    _, _, blocks = sharding.walk_container(memoryview(fqz))
    first, last = sharding.plan_decompress(blocks, world)[rank]
    rec_lo = sum(b.records for b in blocks[:first])
    rec_n = sum(b.records for b in blocks[first:last])
    a = blocks[first].offset if last > first else 10
    b = blocks[last - 1].offset + blocks[last - 1].size if last > first else 10
    shard = np.concatenate([fqz[:10], fqz[a:b]])
    total_records = sum(x.records for x in blocks)
    del fqz
    d_ref = torch.from_numpy(shard).cuda()
    # what the shard must decode to: records [rec_lo, rec_lo + rec_n) of rank 0's generator stream
    cap = d_in.numel()
    want = 0
    for f0 in range(rec_lo, rec_lo + rec_n, 4_000_000):
        c = min(4_000_000, rec_lo + rec_n - f0)
        want += ctx.synth_device(0, SEED, f0, c, d_in.data_ptr() + want, cap - want)
    res = {}

    def dec():
        res["k"] = ctx.decompress_device(d_ref.data_ptr(), d_ref.numel(), d_back.data_ptr(), d_back.numel())

    t_r = timed(dec, 1, 1)
    same = res["k"] == want and bool(torch.equal(d_back[:want], d_in[:want]))
    flag = torch.tensor([1.0 if same else 0.0, float(want)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(flag[1:], op=dist.ReduceOp.SUM)
        dist.barrier()
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass
    total = float(flag[1].item())
    refw = {"value": total / t_r / 1e9, "unit": UNIT, "bit_exact": bool(flag[0].item() == 1.0), "n_gpus": world,
            "records": total_records, "fastq_bytes": int(total), "blocks": len(blocks), "scaling": "strong (one file, block runs per rank)",
            "input": "oracle-written .fqz of the whole workload (reference-shaped: one libzstd-1 frame per stream and block)"}
    return (refw, cb) if rank == 0 else (None, None)


def gzip_workload(ctx, torch, d_in, nrec, timed):
    """SURVEY 8 row f3 (cmd/fqpack/main.go:142-174): the first `nrec` records of the workload as .fq.gz — one zlib level-1
    member (what `gzip -1` / pigz write) and BGZF (what bgzip writes) — inflated on the GPU.  Reports GB/s of TEXT:
    device-resident fqz_gunzip_device, end to end fqz_compress_gz (compressed bytes up, .fqz down, pinned), and zlib's
    inflate on one host core beside them."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    # cut at a record boundary: config-2 records are 4 lines
    probe = d_in[: nrec * 372 + 4096].cpu().numpy()
    nl = np.flatnonzero(probe == 10)
    n = int(nl[4 * nrec - 1]) + 1
    text = probe[:n].tobytes()
    del probe, nl

    def member(t):
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        return b"\x1f\x8b\x08\0\0\0\0\0\0\xff" + c.compress(t) + c.flush() + struct.pack("<II", zlib.crc32(t), len(t) & 0xFFFFFFFF)

    def bgzf_block(p):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = c.compress(p) + c.flush()
        return b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 25 + len(raw)) + raw + struct.pack("<II", zlib.crc32(p), len(p))

    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        fut = ex.submit(member, text)
        blocks = list(ex.map(bgzf_block, [text[i : i + 65280] for i in range(0, n, 65280)] + [b""]))
        files = {"gzip_level1": fut.result(), "bgzf": b"".join(blocks)}
    t_make = time.perf_counter() - t0
    out = {"workload": "first %d records of the config-2 input as .fq.gz" % nrec, "records": nrec, "fastq_bytes": n, "unit": UNIT,
           "made_in_s": t_make}
    d_text = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
    for name, gz in files.items():
        g = np.frombuffer(gz, dtype=np.uint8)
        d_gz = torch.zeros(g.size + 64, dtype=torch.uint8, device="cuda")
        d_gz[: g.size] = torch.from_numpy(g.copy()).cuda()
        res = {}

        def dev():
            res["n"] = ctx.gunzip_device(d_gz.data_ptr(), g.size, d_text.data_ptr(), n)

        dev()
        ctx.stats_reset()
        t = timed(dev, 3, 0)
        stages = {k: round(v["ms"] / 3, 3) for k, v in ctx.stats()["stages"].items() if k.startswith("gz_") and v["ms"]}
        ok = res["n"] == n and bool(torch.equal(d_text[:n].cpu(), torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy())))
        st = ctx.gunzip_stats()
        h_gz = torch.from_numpy(g.copy()).pin_memory()
        h_out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8).pin_memory()
        import ctypes as C

        def e2e():
            m, f = C.c_size_t(0), C.c_size_t(0)
            rc = ctx.lib.L.fqz_compress_gz(ctx.h, C.c_void_p(h_gz.data_ptr()), g.size, 0, C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(m), C.byref(f))
            assert rc == 0 and f.value == n, (rc, f.value)
            res["m"] = m.value

        e2e()
        t0 = time.perf_counter()
        for _ in range(2):
            e2e()
        t_e = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter()
        cpu_bytes, pos = 0, 0  # one host core over the first 64 MiB of the file, member after member
        while pos + 18 < min(len(gz), 1 << 26):
            # BGZF members say how long they are (BSIZE); a plain member runs to the end of the sample
            end = pos + struct.unpack_from("<H", gz, pos + 16)[0] + 1 if name == "bgzf" else min(len(gz), 1 << 26)
            d = zlib.decompressobj(31)
            cpu_bytes += len(d.decompress(gz[pos:end]))
            pos = end
        t_cpu = time.perf_counter() - t0
        out[name] = {"gz_bytes": g.size, "gunzip_device": {"value": 3 * n / t / 1e9, "unit": UNIT, "bit_exact": ok},
                     "compress_gz_e2e": {"value": n / t_e / 1e9, "unit": UNIT, "h2d_bytes": g.size, "d2h_bytes": res["m"], "fqz_bytes": res["m"]},
                     "cpu_zlib_inflate_1core": {"value": cpu_bytes / t_cpu / 1e9, "unit": UNIT},
                     "chunks": st}
        if stages:  # only while fqz_profile_enable is on (tools/gzip_probe.py)
            out[name]["stage_ms_per_call"] = stages
    return out


def side_workload(ctx, torch, kind, seed, nrec, per_record, timed, cpu_threads, what):
    """Device-resident compress + decompress GB/s of a second workload (N = 1); with cpu_threads also the CPU path's ratio."""
    cap = nrec * per_record + (1 << 20)
    d = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = 0
    for first in range(0, nrec, 4_000_000):
        c = min(4_000_000, nrec - first)
        n += ctx.synth_device(kind, seed, first, c, d.data_ptr() + n, cap - n)
    d_out = torch.empty(n // 2 + (1 << 20), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n + (1 << 16), dtype=torch.uint8, device="cuda")
    res = {}

    def comp():
        res["m"] = ctx.compress_device(d.data_ptr(), n, d_out.data_ptr(), d_out.numel())

    def dec():
        res["k"] = ctx.decompress_device(d_out.data_ptr(), res["m"], d_back.data_ptr(), d_back.numel())

    t_c = timed(comp, 3, 2)
    t_d = timed(dec, 3, 2)
    ok = res["k"] == n and bool(torch.equal(d_back[:n], d[:n]))
    out = {"workload": what, "records": nrec, "fastq_bytes": n, "fqz_bytes": res["m"], "ratio": n / res["m"],
           "compress": {"value": 3 * n / t_c / 1e9, "unit": UNIT}, "decompress": {"value": 3 * n / t_d / 1e9, "unit": UNIT}, "round_trip_ok": ok}
    if cpu_threads:
        from oracle import fqz_oracle as oracle

        text = d[:n].cpu().numpy()
        t0 = time.perf_counter()
        fqz = oracle.compress_np(text, threads=cpu_threads)
        dt = time.perf_counter() - t0
        out["cpu"] = {"ratio": n / fqz.size, "compress": {"value": n / dt / 1e9, "unit": UNIT}, "cores": cpu_threads, "kind": "port"}
        out["ratio_vs_cpu"] = (n / res["m"]) / (n / fqz.size)
        # the CPU-written file decodes on the GPU too
        d_ref = torch.from_numpy(fqz).cuda()
        k = ctx.decompress_device(d_ref.data_ptr(), d_ref.numel(), d_back.data_ptr(), d_back.numel())
        out["cpu_written_decodes_bit_exact"] = k == n and bool(torch.equal(d_back[:n], d[:n]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=0, help="records per GPU (default: BASELINE config 2, 25 000 000)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"], help="cfg5: one 64 GB input block-sharded over the ranks")
    ap.add_argument("--total-bytes", type=int, default=0, help="cfg5: size of the logical input (default 64e9)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-4 and duplicates side workloads")
    ap.add_argument("--frontend", type=int, default=0, choices=[0, 1, 2], help="A/B of the front-end kernels: see FQZ_OPT_FRONTEND")
    ap.add_argument("--window-bytes", type=int, default=0, help="FASTQ bytes per device pass of the compress calls (default 3e9): FQZ_OPT_WINDOW_BYTES")
    ap.add_argument("--huf-kernels", type=int, default=0, choices=[0, 1], help="A/B of the literals-only frame coder: see FQZ_OPT_HUF_KERNELS")
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
    elif args.workload == "cfg5":
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
