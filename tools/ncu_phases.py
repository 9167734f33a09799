#!/usr/bin/env python3
"""Stall samples of k_tile_step by PHASE of eval_tile (line ranges of tile_core.cuh), from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --launch-skip N --launch-count 1 > f.csv`.
usage: ncu_phases.py f.csv"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
cur_file = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name"): continue
    try:
        d = {h: r[i] for i, h in enumerate(hdr) if i < len(r)}
        out.append((cur_file, int(r[0]), int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0), d))
    except ValueError: pass
# phases: (name, file, lo, hi); first match wins; line numbers are looked up from marker comments so they survive edits
src = open(__file__.replace("tools/ncu_phases.py", "devastator_b200/csrc/tile_core.cuh")).read().split("\n")
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
L = {k: find(p) for k, p in dict(append="DEVA_HD void append_local", route="DEVA_HD void route_append", walk="enum { MODE_FIRST", net="DEVA_HD void net_segment",
     scan="DEVA_HD void block_scan", runlp="DEVA_HD void run_lp_core", runrec="DEVA_HD void run_records", run="// ---- run", flush="// ---- flush",
     blk="DEVA_HD void blk_begin", eval="DEVA_HD void eval_tile", fast="// ---- fast path", slow="// ---- slow path", wb="// ---- write back",
     follow="struct WHdr", rebin="DEVA_HD void rebin_tile", sm="// ---- the epoch state machine").items()}
phases = [("append/route (flush atomics+stores)", L["append"], L["walk"]), ("walk (per-LP event loop)", L["walk"], L["net"]), ("net_segment", L["net"], L["scan"]),
          ("block_scan", L["scan"], L["runlp"]), ("run_lp_core (sort+chain)", L["runlp"], L["runrec"]), ("run_records: count/scatter/bins", L["runrec"], L["run"]),
          ("run_records: run loop", L["run"], L["flush"]), ("run_records: flush loop", L["flush"], L["blk"]), ("blk_begin/flush", L["blk"], L["eval"]),
          ("eval_tile: header", L["eval"], L["fast"]), ("eval_tile: TMA load + flags", L["fast"], L["slow"]), ("eval_tile: slow path", L["slow"], L["wb"]),
          ("eval_tile: write back", L["wb"], L["follow"]), ("follow_tile_warp", L["follow"], L["rebin"]), ("rebin_tile", L["rebin"], L["sm"]), ("state machine", L["sm"], 10**9)]
ts = sum(o[2] for o in out); ti = sum(o[3] for o in out)
acc = {}
stalls = {}
for f, ln, s, i, d in out:
    key = f
    if f == "tile_core.cuh":
        key = next((n for n, lo, hi in phases if lo <= ln < hi), "tile_core other")
    elif f in ("models.cuh", "deva_log.cuh"): key = "model (execute + log)"
    a = acc.setdefault(key, [0, 0]); a[0] += s; a[1] += i
    st = stalls.setdefault(key, {})
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            try: st[k] = st.get(k, 0) + int(v)
            except ValueError: pass
print(f"samples {ts} warp-inst {ti}")
for k, (s, i) in sorted(acc.items(), key=lambda x: -x[1][0]):
    top = sorted(stalls[k].items(), key=lambda x: -x[1])[:3]
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins  {k:40s} " + " ".join(f"{a[6:]}={100*b/max(s,1):.0f}%" for a, b in top))
tot = {}
for st in stalls.values():
    for k, v in st.items(): tot[k] = tot.get(k, 0) + v
print("all:", " ".join(f"{k[6:]}={100*v/ts:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
