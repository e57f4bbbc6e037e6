#!/usr/bin/env python
"""Dynamic instruction counts of a profiled kernel per source call site: joins the per-instruction counters of an
.ncu-rep (source page) with the inlining chains of the cubin (nvdisasm -gi).
Usage: python scripts/ncu_by_source.py <rep> <cubin> <kernel-substring> [depth] [prefix-filter]"""
import csv
import re
import subprocess
import sys

rep, cubin, pat = sys.argv[1:4]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
flt = sys.argv[5] if len(sys.argv) > 5 else ""
txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if l.startswith("\t.section\t.text.") and pat in l)
rx = re.compile(r'//## File "([^"]+)", line (\d+)')
chain, pending, chains = (), [], {}
for l in txt[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = rx.search(l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        if pending:
            chain = tuple(reversed(pending)); pending = []
        chains[int(m.group(1), 16)] = chain
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
ia, iw, it, ismp = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base = int(data[0][ia], 16)
agg, order = {}, []
tw = tt = ts = 0
for r in data:
    off = int(r[ia], 16) - base
    ch = chains.get(off, ())
    key = ch[:depth]
    w, t, s = int(r[iw]), int(r[it]), int(r[ismp])
    tw += w; tt += t; ts += s
    if key not in agg:
        agg[key] = [0, 0, 0, 0]; order.append(key)
    a = agg[key]; a[0] += w; a[1] += t; a[2] += s; a[3] += 1
print(f"total warp instr {tw:.4g}, lanes/instr {tt / tw:.2f}")
print(" warp%  lanes stall%  n_sass  call site")
for k in order:
    name = " > ".join(f"{f}:{n}" for f, n in k)
    if flt and flt not in name:
        continue
    w, t, s, n = agg[k]
    if w > 0.002 * tw:
        print(f"{100 * w / tw:6.2f} {t / max(w, 1):6.1f} {100 * s / max(ts, 1):6.2f} {n:6d}  {name}")
