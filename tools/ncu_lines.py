"""Attribute ncu stall samples / executed instructions of one kernel to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring>   (needs ncu, cuobjdump, nvdisasm)"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, obj, key = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# locate the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and key in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section\t")), len(dis))
lines = []  # (file:line, inline-chain, sass)
cur = "?"
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
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ii, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
prof = []
for r in rows:
    try: prof.append((r[si].strip(), int(r[ii]), int(r[sm])))
    except (ValueError, IndexError): pass
print(f"disasm instrs {len(lines)}  profile instrs {len(prof)}")
n = min(len(lines), len(prof))
by = defaultdict(lambda: [0, 0])
for (loc, sass), (psass, cnt, smp) in zip(lines[:n], prof[:n]):
    by[loc][0] += cnt; by[loc][1] += smp
ti = sum(v[0] for v in by.values()); ts = sum(v[1] for v in by.values())
byfile = defaultdict(lambda: [0, 0])
for loc, v in by.items():
    byfile[loc.split(":")[0]][0] += v[0]; byfile[loc.split(":")[0]][1] += v[1]
print("by file:")
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]): print(f"  {f:24s} insts {100*v[0]/ti:5.1f}%  samples {100*v[1]/ts:5.1f}%")
print("top lines by samples:")
for loc, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print(f"  {loc:28s} insts {100*v[0]/ti:5.1f}%  samples {100*v[1]/ts:5.1f}%")
