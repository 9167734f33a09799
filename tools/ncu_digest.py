#!/usr/bin/env python3
"""Digest of an ncu report of k_pass: headline metrics per launch, and stall samples / executed
instructions per source line (ncu SASS page joined with nvdisasm --print-line-info of the same build).
usage: ncu_digest.py report.ncu-rep [launch_index] [lib.so]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "devastator_b200", "libdeva_b200.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
want = ["Grid Size", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers"]
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:62s} {units[i]:10s}", [r[i][:12] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src))); kname = srows[0][1]; sh = srows[1]; ix = {h: i for i, h in enumerate(sh)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
mang = "9PholdTestELi6" if "PholdTest" in kname else "10PholdBenchELi6"
fn = line = None; amap = {}
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if fn and "k_passIN9deva_b200" + mang in fn:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if m: amap[int(m.group(1), 16)] = line
data = []; base = None; seen = set()
for r in srows[2:]:
    if len(r) < len(sh): continue
    try: a = int(r[ix["Address"]], 16)
    except ValueError: continue
    if base is None: base = a
    if a in seen: continue
    seen.add(a); data.append((a - base, r))
tot_s = sum(int(r[ix["# Samples"]]) for _, r in data); tot_i = sum(int(r[ix["Instructions Executed"]]) for _, r in data)
print("SASS instr", len(data), "samples", tot_s, "warp-inst", tot_i)
for k in sh:
    if k.startswith("stall_") and "Not Issued" not in k:
        v = sum(int(r[ix[k]]) for _, r in data)
        if v * 100 > tot_s: print(f"  {k:24s} {v / tot_s * 100:5.1f}%")
bl = collections.Counter(); bi = collections.Counter(); lsb = collections.Counter()
for off, r in data:
    ln = amap.get(off); bl[ln] += int(r[ix["# Samples"]]); bi[ln] += int(r[ix["Instructions Executed"]]); lsb[ln] += int(r[ix["stall_long_sb"]])
print("--- source lines by stall samples (share of samples | long_sb share | share of executed warp-inst)")
for ln, c in bl.most_common(int(os.environ.get("TOP", 28))):
    print(f"{c / tot_s * 100:5.1f}%  lsb {lsb[ln] / tot_s * 100:5.1f}%  inst {bi[ln] / tot_i * 100:5.1f}%  {ln}")
