"""world_size-2 gloo test of the N>1 path on the CPU: sample-range sharding + one reduce of the per-rank
accumulators.  The per-rank compute is the product's per-lane code run through tests/hostsim (no GPU here);
on the GPU box the same sharding helpers drive libraydar_cuda.so (bench.py, rdr_create_multi)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    import torch
    import torch.distributed as dist
    from oracle import orc
    import hostsim_py as hs
    from raydar_b200 import dist as rdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = orc.load_rscn(os.path.join(ROOT, "scenes", "default.rscn")).with_resolution(107, 60)
    spp = 6
    first, count = (rdist.weak_sample_range(rank, spp) if mode == "weak" else rdist.strong_sample_range(rank, world, spp))
    acc, _ = hs.render(scene, 42, first, count, 12)
    t = torch.from_numpy(acc.reshape(-1))
    rdist.reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy().reshape(acc.shape))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["weak", "strong"])
def test_sample_range_sharding_and_reduce(hs, orc, default_scene, tmp_path, mode):
    import torch.multiprocessing as mp
    world, spp = 2, 6
    out = str(tmp_path / f"acc_{mode}.npy")
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    got = np.load(out)
    scene = default_scene.with_resolution(107, 60)
    total = spp * world if mode == "weak" else spp
    want = orc.render(scene, 42, 0, total, 12, n_threads=2)
    assert np.array_equal(got[..., 3], want[..., 3])                 # every pixel got every sample exactly once
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)              # equal up to f32 summation order
    # the resolved 8-bit image differs by at most one level
    d = np.abs(orc.resolve(got, total).astype(int) - orc.resolve(want, total).astype(int))
    assert d.max() <= 1


def test_ranges_partition_the_samples():
    from raydar_b200 import dist as rdist
    for world in (1, 2, 4, 8):
        for total in (0, 1, 7, 1024, 4096):
            ranges = [rdist.strong_sample_range(r, world, total) for r in range(world)]
            covered = [s for b, n in ranges for s in range(b, b + n)]
            assert covered == list(range(total))
        assert [rdist.weak_sample_range(r, 1024)[0] for r in range(world)] == [1024 * r for r in range(world)]
