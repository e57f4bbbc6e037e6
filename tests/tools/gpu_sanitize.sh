#!/bin/bash
# TEST TOOL (checks against the oracle, hence under tests/).  compute-sanitizer (memcheck, racecheck, synccheck) over small renders with the fused scan, the generic fused scan and
# the cooperative hierarchy.  Usage: bash tests/tools/gpu_sanitize.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import raydar_b200 as rb
import synth_scenes as ss
from oracle import orc
which = sys.argv[1]
if which == "fused":
    scene = orc.load_rscn("scenes/benchmark.rscn").with_resolution(96, 54); accel = rb.ACCEL_FUSED; bounces = 12
elif which == "fused24":
    scene = ss.config5(64, 36); accel = rb.ACCEL_FUSED; bounces = 32
else:
    scene = ss.config4(20000, 64, 36); accel = rb.ACCEL_BVH_COOP; bounces = 12
r = rb.Renderer(rb.RendererConfig(2, bounces)); r.set_seed(5); r.set_accel(accel)
r.render_frame(scene); acc = r.read_accum()
ids, _ = r.first_hit()
want = orc.render(scene, 5, 0, 2, bounces, n_threads=8)
print(which, "match", np.array_equal(acc.view(np.uint32), want.view(np.uint32)))
r.close()
PY
for tool in memcheck racecheck synccheck; do
  for w in fused fused24 bvh2; do
    echo "== $tool $w"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py $w 2>&1 | grep -E "match|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -8
  done
done 2>&1 | tee $OUT/sanitizer_$TAG.log
