#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libdeva_b200.so matching a name - instruction count, mnemonic
histogram of the memory / synchronisation / Blackwell-specific instructions, and every line that holds one of
the latter (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, ACQBULK/PREEXIT = programmatic launch).
usage: sass_listing.py [name-substring] > profiles/r2_sass_k_tile_step.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "devastator_b200", "libdeva_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else "k_tile_step"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
special = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|LDGDEPBAR|DEPBAR|FENCE|ACQBULK|PREEXIT|REDUX|ATOMG|ATOMS|RED|MEMBAR|ERRBAR|CCTL|BAR|MATCH|VOTE|SHFL|NANOSLEEP|CS2R)\b")
fn = None; body = collections.defaultdict(list)
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m: fn = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if fn and m: body[fn].append((m.group(1), m.group(2).strip()))
print("arch:", re.search(r"arch = (\S+)", out).group(1), " library:", os.path.relpath(lib, root))
for f, ins in body.items():
    if want not in f: continue
    dem = subprocess.run(["cu++filt", f], capture_output=True, text=True).stdout.strip() or f
    hist = collections.Counter()
    for _, t in ins:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0].split(".")[0]] += 1
    print("\n== %s\n   %d SASS instructions" % (dem, len(ins)))
    print("   " + "  ".join("%s %d" % kv for kv in sorted(hist.items(), key=lambda x: -x[1]) if special.search(kv[0]) or kv[0] in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "DFMA", "DMUL", "DADD", "IMAD", "BRA")))
    for a, t in ins:
        if re.search(r"\b(UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|ACQBULK|PREEXIT|FENCE)\b", t): print("   /*%s*/ %s" % (a, t))
