"""Per-launch summary of an `ncu --set full` report: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor inst %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("launch__registers_per_thread", "regs"), ("smsp__warps_active.avg.per_cycle_active", "warps/SMSP")]
cols = [(c, n) for c, n in cols if c in hdr]
print(f"`ncu --set full --clock-control none` of {sys.argv[1].split('/')[-1]} (cold caches, serialised launches: shares and ratios, not absolute times).\n")
print("| " + " | ".join(n + (f" [{units[hdr.index(c)]}]" if units[hdr.index(c)] and n in ("DRAM read", "DRAM write") else "") for c, n in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    out = []
    for c, n in cols:
        v = r[hdr.index(c)]
        if n == "kernel":
            v = "`" + re.sub(r"\(.*", "", v).replace("void ", "") + "`"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
