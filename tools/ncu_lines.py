#!/usr/bin/env python3
"""Per-source-line digest of an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::N > f.csv`
then `python tools/ncu_lines.py f.csv [top]`: samples, warp instructions and thread instructions per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try:
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed")
        out.append((int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_t] or 0), cur_file, r[0], r[1].strip()))
    except (ValueError, IndexError):
        pass
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out); tt = sum(o[2] for o in out)
print(f"total samples {ts}  warp-inst {ti}  thread-inst {tt}  avg lanes {tt / max(ti, 1):.1f}")
byfile = {}
for o in out: byfile.setdefault(o[3], [0, 0]); byfile[o[3]][0] += o[0]; byfile[o[3]][1] += o[1]
for f, (s, i) in sorted(byfile.items(), key=lambda x: -x[1][0]): print(f"  {f:24s} samples {100*s/max(ts,1):5.1f}%  warp-inst {100*i/max(ti,1):5.1f}%")
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{100*o[0]/max(ts,1):5.1f}% smp {100*o[1]/max(ti,1):5.1f}% ins  {o[3]}:{o[4]:>4s}  {o[5][:100]}")
