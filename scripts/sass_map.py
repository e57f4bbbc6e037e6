#!/usr/bin/env python
"""Attribute the SASS of one kernel to source lines THROUGH the inlining chain (nvdisasm -gi), so that code size can
be read per call site rather than per leaf wrapper (fadd/fmul/...).
Usage: python scripts/sass_map.py <cubin> <kernel-name-substring> [depth]
Prints instruction counts keyed by the outermost `depth` frames of the chain (file:line > file:line ...)."""
import re
import subprocess
import sys

cubin, pat = sys.argv[1], sys.argv[2]
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if l.startswith("\t.section\t.text.") and pat in l)
chain, pending, counts, order, total = (), [], {}, [], 0
rx = re.compile(r'//## File "([^"]+)", line (\d+)')
for l in txt[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = rx.search(l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        if pending:
            chain = tuple(reversed(pending))      # outermost first
            pending = []
        key = chain[:depth]
        if key not in counts:
            order.append(key)
        counts[key] = counts.get(key, 0) + 1
        total += 1
    elif pending and not l.strip().startswith("//"):
        pass
print("total instructions", total, "=", total * 16, "bytes")
for k in order:
    print(f"{counts[k]:6d}  " + " > ".join(f"{f}:{n}" for f, n in k))
