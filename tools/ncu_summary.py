"""Markdown summary (one row per profiled launch) of an `ncu --set full` report.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_rNN.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "HMMA %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem")]
print(f"# ncu --set full summary: {rep}\n")
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for d in data:
    name = d[ix["Kernel Name"]].replace("void ", "").replace("extdm::", "")
    name = name.split("(")[0][:70]
    cells = []
    for key, _ in cols:
        if key in ix:
            v, u = d[ix[key]], units[ix[key]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        else:
            cells.append("-")
    print(f"| `{name}` | " + " | ".join(cells) + " |")
