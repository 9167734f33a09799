#!/usr/bin/env python3
"""profiles/k_tile_step_traffic.json from an `ncu --set full` capture of tools/tile_probe.py (BASELINE
config 2, adaptive window - the configuration bench.py runs at N = 1): DRAM bytes, duration and executed
instructions per captured launch of k_tile_step, and the figures of the epoch's full launch.
usage: ncu_traffic.py report.ncu-rep [commits_per_epoch]"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[ix[k]].replace(",", "")); u = units[ix[k]]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3}.get(u, 1)
launches = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if "k_tile_step" not in name: continue
    launches.append({"kernel": "full" if ("true" in name or ", 1>" in name or "Lb1" in name) else "follow-up", "us": val(r, "gpu__time_duration.sum"),
                     "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                     "warp_inst": val(r, "smsp__inst_executed.sum"), "lanes": float(r[ix["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                     "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                     "warps_active_pct": float(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]])})
# an epoch = a full launch that evaluated tiles (not a rebin: those write as much as they read) + the follow-ups after it
full = [l for l in launches if l["kernel"] == "full" and l["dram_write"] < l["dram_read"]]
f = full[0]
i0 = launches.index(f)
epoch = [f]
for l in launches[i0 + 1:]:
    if l["kernel"] == "full": break
    epoch.append(l)
commits = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = {"source": os.path.basename(rep), "how": "ncu --set full --clock-control none, tools/tile_probe.py (1 048 576 LPs, p_remote 0.10, adaptive window), one steady-state epoch",
       "dram_bytes_per_full_launch": int(f["dram_read"] + f["dram_write"]), "full_launch_us_under_ncu": f["us"],
       "dram_bytes_per_epoch": int(sum(l["dram_read"] + l["dram_write"] for l in epoch)), "epoch_launches": len(epoch),
       "commits_per_epoch": commits, "dram_bytes_per_commit": (sum(l["dram_read"] + l["dram_write"] for l in epoch) / commits) if commits else None,
       "algorithmic_bytes_per_commit": 120, "launches": launches}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k_tile_step_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "launches"}, indent=1))
