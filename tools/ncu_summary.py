"""Summarise ncu outputs for profiles/: launch list CSV -> per-kernel shares; .ncu-rep -> key metrics.
usage: python tools/ncu_summary.py launches <launches.csv> | top <file.ncu-rep>"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms total (ncu per-launch times: cold cache, serialised)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} n={n:4d} total={t / 1e3:9.3f} ms  avg={t / n:10.1f} us  share={t / tot * 100:5.1f}%")


def top(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in ["Kernel Name"] + WANT if w in hdr]
    for r in rows[2:]:
        print("---")
        for i in idx:
            print(f"  {hdr[i]:72s} {r[i][:90]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "top": top}[sys.argv[1]](sys.argv[2])
