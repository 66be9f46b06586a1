"""Per-source-line share of executed instructions / stall samples of one kernel in an .ncu-rep.
usage: python tools/ncu_hot_lines.py <file.ncu-rep> <kernel regex> <object.o> <mangled function substring> <source.cu>
Maps ncu's SASS page (instruction order) onto nvdisasm -g line markers of the same function."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kre, obj, fun, srcfile = sys.argv[1:6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}", "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((r[si].strip(), int(r[ie]), int(r[ss])))
    except (ValueError, IndexError):
        pass
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and fun in l][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//--------------------- .text.")), len(dis))
line, seq = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(line)
n = min(len(seq), len(sass))
agg, samp = collections.Counter(), collections.Counter()
for i in range(n):
    agg[seq[i]] += sass[i][1]
    samp[seq[i]] += sass[i][2]
tot, ts = sum(agg.values()) or 1, sum(samp.values()) or 1
src = open(srcfile).read().split("\n")
print(f"# {kre}: {len(sass)} SASS instructions (nvdisasm {len(seq)}), {tot} warp instructions executed, {ts} stall samples")
for k, v in agg.most_common(int(sys.argv[6]) if len(sys.argv) > 6 else 30):
    f, ln = k if k else ("?", 0)
    text = src[ln - 1].strip()[:100] if f == os.path.basename(srcfile) and ln > 0 else f
    print(f"{v / tot * 100:5.1f}% instr {samp[k] / ts * 100:5.1f}% samples  {f}:{ln}  {text}")
