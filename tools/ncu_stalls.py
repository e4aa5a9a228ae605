"""Aggregate warp-stall samples of one kernel from an ncu report, by CUDA source line.
   python tools/ncu_stalls.py report.ncu-rep kernel_regex [top_n]      (compile with -lineinfo)"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, cur_fn, hdr, ix = None, None, None, None
lines = {}          # (fn, file, line) -> [samples, text, {stall: n}]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    key = (cur_fn, cur_file, int(r[0]))
    e = lines.setdefault(key, [0, r[1], {h: 0 for h in stall_cols}])
    e[0] += n
    for h in stall_cols:
        e[2][h] += int(r[ix[h]] or 0)
fns = sorted({k[0] for k in lines})
for fn in fns:
    sel = {k: v for k, v in lines.items() if k[0] == fn}
    total = sum(v[0] for v in sel.values())
    print(f"== {fn[:150]}\n   total samples {total}")
    tot = {}
    for v in sel.values():
        for h, c in v[2].items():
            tot[h] = tot.get(h, 0) + c
    for h, c in sorted(tot.items(), key=lambda kv: -kv[1])[:7]:
        print(f"   {h:26s} {c:8d} {100 * c / max(total, 1):5.1f}%")
    for k, v in sorted(sel.items(), key=lambda kv: -kv[1][0])[:top]:
        rs = sorted(((c, h) for h, c in v[2].items()), reverse=True)[:2]
        print(f"   {v[0]:6d} {100 * v[0] / max(total, 1):5.1f}%  {k[1]}:{k[2]:<4d} {v[1].strip()[:80]:80s} "
              f"{rs[0][1][6:]}:{rs[0][0]} {rs[1][1][6:]}:{rs[1][0]}")
