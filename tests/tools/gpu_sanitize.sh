#!/bin/bash
# TEST TOOL (checks against the oracle, hence under tests/).  compute-sanitizer (memcheck, racecheck, synccheck) over small renders with the fused scan, the generic fused scan and
# the cooperative hierarchy, and (on >= 2 devices) the multi-GPU handle with the peer-memory combine.  Usage: bash tests/tools/gpu_sanitize.sh <tag>
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
elif which == "multi":
    # the multi-GPU handle on two devices: peer_combine_kernel over NVLink peer memory, both partitions, progressive calls
    import torch
    if torch.cuda.device_count() < 2:
        print("multi skipped: one device"); sys.exit(0)
    scene = orc.load_rscn("scenes/benchmark.rscn").with_resolution(96, 54)
    one = rb.Renderer(rb.RendererConfig(4, 12)); one.set_seed(5); img1 = one.render_frame(scene)
    m = rb.Renderer(rb.RendererConfig(4, 12), devices=[0, 1]); m.set_seed(5)
    a = m.render_frame(scene)
    m.set_partition(rb.PARTITION_STRIPES, 16); b = m.render_frame(scene)
    m.new_frame(scene)
    while m.render_sample(scene) is not None: pass
    print(which, "match", bool(np.abs(a.astype(int) - img1.astype(int)).max() <= 1 and np.array_equal(b, img1)
                              and np.array_equal(m.read_accum().view(np.uint32), one.read_accum().view(np.uint32))))
    one.close(); m.close(); sys.exit(0)
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
  for w in fused fused24 bvh2 multi; do
    echo "== $tool $w"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py $w 2>&1 | grep -E "match|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -8
  done
done 2>&1 | tee $OUT/sanitizer_$TAG.log
