#!/usr/bin/env python
"""TEST / BENCH INPUT TOOL (lives under tests/: it uses the test-side scene generator, which borrows the oracle's Scene
container).  Writes the seeded synthetic scenes of BASELINE.json configs 4 and 5 as .rscn files, so that
bench.py / raydar-cuda load them through the product's own scene loader.
Usage: python tests/tools/make_synth_scenes.py <out_dir> [config4 [n]] [config5]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_scenes as ss  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
which = sys.argv[2:] or ["config4", "config5"]
if "config4" in which:
    i = which.index("config4")
    n = int(which[i + 1]) if i + 1 < len(which) and which[i + 1].isdigit() else 100_000
    ss.write_rscn(ss.config4(n), os.path.join(out, "config4.rscn"))
if "config5" in which:
    ss.write_rscn(ss.config5(), os.path.join(out, "config5.rscn"))
print("wrote", os.listdir(out))
