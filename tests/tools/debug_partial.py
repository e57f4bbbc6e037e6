import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raydar_b200 as rb
from oracle import orc
d = orc.load_rscn('scenes/default.rscn')
for (w, h, spp, cull) in [(427, 240, 16, True), (427, 240, 16, False), (512, 240, 16, True), (214, 120, 4, True), (100, 3, 8, True), (300, 1, 8, True)]:
    s = d.with_resolution(w, h)
    r = rb.Renderer(rb.RendererConfig(spp, 12)); r.set_seed(99); r.debug_set_cull(cull)
    r.render_frame(s); acc = r.read_accum().reshape(-1, 4)
    want = orc.render(s, 99, 0, spp, 12, n_threads=orc.max_threads()).reshape(-1, 4)
    bad = np.nonzero((acc.view(np.uint32) != want.view(np.uint32)).any(1))[0]
    print(f"{w}x{h} spp={spp} cull={cull}: n_pixels={w*h} bad={len(bad)}", "w histogram:", dict(zip(*np.unique(acc[:, 3], return_counts=True))))
    if len(bad):
        print("  bad range", bad.min(), bad.max(), "blocks", np.unique(bad // 256)[:20], "lanes-in-block", np.unique(bad % 256)[:40], "...")
        print("  first bad", bad[0], acc[bad[0]], want[bad[0]])
    r.close()
