"""Attribute ONE stall reason of an ncu capture to CUDA source lines (ncu_lines.py does all samples).
usage: ncu_stall_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring> <stall column, e.g. stall_long_sb> [top n]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, obj, key, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and key in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section\t")), len(dis))
lines, cur = [], "?"
for l in dis[start:end]:
  m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
  if m:
    cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
    continue
  m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
  if m:
    lines.append((cur, m.group(2).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Source" in r and col in r)
ci, si = hdr.index(col), hdr.index("Source")
prof = []
for r in rows[rows.index(hdr) + 1:]:
  try: prof.append((r[si].strip(), int(r[ci])))
  except (ValueError, IndexError): pass
n = min(len(lines), len(prof))
by, byop = defaultdict(int), defaultdict(int)
for (loc, sass), (ps, c) in zip(lines[:n], prof[:n]):
  by[loc] += c
  byop[ps.split()[0] if not ps.startswith("@") else ps.split()[1]] += c
tot = sum(by.values())
print(f"{col}: {tot} samples over {n} instructions")
for loc, c in sorted(by.items(), key=lambda kv: -kv[1])[:top]:
  print(f"  {loc:28s} {100 * c / tot:5.1f}%")
print("by opcode of the stalled instruction:")
for op, c in sorted(byop.items(), key=lambda kv: -kv[1])[:12]:
  print(f"  {op:16s} {100 * c / tot:5.1f}%")
