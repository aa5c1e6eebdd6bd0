"""Aggregate an `ncu --page source --csv` export per CUDA source line.

ncu's CSV lists SASS instructions in program order without line numbers; nvdisasm -g lists the
same instructions in the same order with '//## File ..., line N' markers.  Zip the two.

    python tools/ncu_lines.py <source.csv> <libgsf.so> <kernel-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

csv_path, so_path, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
sass = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in out:
            sass = out
            break
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if ".section\t.text." in l and kname in l)
end = next((i for i, l in enumerate(lines) if i > start and l.startswith("//---------------------")), len(lines))
line_of = []          # per instruction: (file, line)
cur = ("?", 0)
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        line_of.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
print(f"# {len(data)} profiled instructions, {len(line_of)} disassembled instructions")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
total = 0
for k, r in enumerate(data):
    key = line_of[k] if k < len(line_of) else ("?", -1)
    try:
        s = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    total += s
    agg[key]["samples"] += s
    agg[key]["inst"] += int(r[idx["Instructions Executed"]] or 0)
    for st in stalls:
        agg[key][st] += int(r[idx[st]] or 0)
tot_inst = sum(v["inst"] for v in agg.values())
print(f"# total samples {total}, warp instructions {tot_inst}")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    top_st = sorted(((c[s], s.replace('stall_', '')) for s in stalls), reverse=True)[:3]
    desc = " ".join(f"{n}:{100 * v / max(c['samples'], 1):.0f}%" for v, n in top_st)
    print(f"{key[0]}:{key[1]:<5d} samples {100 * c['samples'] / total:5.1f}%  inst {100 * c['inst'] / tot_inst:5.1f}%   {desc}")
