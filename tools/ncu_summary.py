"""Summarise an .ncu-rep: key metrics, stall reasons, instruction mix (needs ncu on PATH)."""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
for v in rows[2:]:
    name = v[h.index("Kernel Name")]
    print("== kernel:", name[:110])
    stalls = []
    for i, n in enumerate(h):
        if "average_warps_issue_stalled" in n:
            try: stalls.append((float(v[i]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("stalls/issue:", ", ".join(f"{n}={x:.2f}" for x, n in sorted(stalls, reverse=True)[:8]))
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    for i, n in enumerate(h):
        if n in want: print(f"  {n} = {v[i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next((r for r in rows if "Source" in r and "Instructions Executed" in r), None)
if hdr:
    si, ii, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp, tot = Counter(), Counter(), 0
    for r in rows:
        try: n = int(r[ii]); s = int(r[sm])
        except (ValueError, IndexError): continue
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[si])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += n; samp[op] += s; tot += n
    print("instruction mix (warp instructions, first kernel in report): total", tot)
    for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
        print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%  stall-samples {samp[op]}")
