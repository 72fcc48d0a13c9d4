"""Stall samples of one kernel per CUDA SOURCE LINE: the SASS page of an .ncu-rep joined with nvdisasm's line info.

ncu's own source view needs the GUI; its CSV export has the per-instruction sample counts but no line numbers, while
`nvdisasm --print-line-info` has the line of every instruction of the same cubin.  The two listings are in the same order.

Usage: python scripts/ncu_lines.py REPORT.ncu-rep MANGLED_KERNEL_NAME OBJECT.o [N_LINES]
  e.g. python scripts/ncu_lines.py gpurun_out/r02_prof_dopri5.ncu-rep _Z16sv_dopri5_kernelILi3ELi1ELb1EEv10Dopri5Args \
       casq_b200/csrc/_obj/sv_dopri5.o 40
(the object must be the build the report was taken from; compile with -lineinfo, which the product build does)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, fun, obj = sys.argv[1], sys.argv[2], os.path.abspath(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "casq_b200", "csrc")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(f"cd {tmp} && cuobjdump -xelf all {obj} > /dev/null && nvdisasm --print-line-info *.cubin > lines.txt 2>/dev/null", shell=True, check=True)
    subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass > {tmp}/sass.csv 2>/dev/null", shell=True, check=True)
    lines = open(os.path.join(tmp, "lines.txt")).read().split("\n")
    rows = list(csv.reader(open(os.path.join(tmp, "sass.csv"))))
start = next(i for i, l in enumerate(lines) if (".text." + fun) in l and l.strip().startswith(".section"))
cur, seq = None, []
for l in lines[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/\s+.*?;", l):
        seq.append(cur)
# the report lists every captured launch in turn: "Kernel Name" row, column header row, one row per instruction
demangled = None
agg, total, launches, k = collections.defaultdict(lambda: [0, 0, 0]), 0, 0, 0
hdr = i_s = i_e = None
for r in rows:
    if r and r[0] == "Kernel Name":
        demangled, k = r[1], -1
        continue
    if r and r[0] == "Address":
        hdr, k = r, 0
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        launches += 1
        continue
    if hdr is None or k < 0:
        continue
    try:
        s, e = int(r[i_s]), int(r[i_e])
    except (ValueError, IndexError):
        continue
    key = seq[k] if k < len(seq) and seq[k] else ("?", 0)
    k += 1
    agg[key][0] += s
    agg[key][1] += 1
    agg[key][2] += e
    total += s
print(f"{len(seq)} instructions in the cubin, {launches} launch(es) in the report (all must be {fun}), {total} stall samples")
cache = {}
for key, (s, n, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    path = os.path.join(src_dir, key[0])
    if key[0] not in cache:
        cache[key[0]] = open(path).read().split("\n") if os.path.exists(path) else None
    text = cache[key[0]][key[1] - 1].strip()[:100] if cache[key[0]] and key[1] > 0 else ""
    print(f"{s:8d} {100 * s / total:5.1f}%  sass={n:4d} executed={e:12d}  {key[0]}:{key[1]}  {text}")
