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
    first, count = (rdist.weak_sample_range(rank, spp) if mode == "weak" else rdist.strong_sample_range(rank, world, spp))   # "peer": strong
    acc, _ = hs.render(scene, 42, first, count, 12)
    t = torch.from_numpy(acc.reshape(-1))
    if mode == "peer":
        # CPU model of the fused reduce + resolve (peer_combine_kernel): every rank "reads" all accumulators (all_gather
        # stands in for the NVLink peer loads), sums ITS pixel slice in rank order, quantises it, and the RGBA8 slices
        # meet on rank 0
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        n_pixels = acc.shape[0] * acc.shape[1]
        first, n = rdist.peer_pixel_slice(rank, world, n_pixels)
        total = parts[0].numpy().reshape(-1, 4)[first:first + n].copy()
        for p in parts[1:]:
            total = total + p.numpy().reshape(-1, 4)[first:first + n]
        mine = torch.from_numpy(orc.resolve(total, spp).reshape(-1).copy())
        sizes = [rdist.peer_pixel_slice(r, world, n_pixels)[1] * 4 for r in range(world)]
        slices = [torch.empty(sz, dtype=torch.uint8) for sz in sizes] if rank == 0 else None
        if rank == 0:
            slices[0] = mine
            for r in range(1, world):
                dist.recv(slices[r], src=r)
            np.save(out_path, torch.cat(slices).numpy().reshape(acc.shape[0], acc.shape[1], 4))
        else:
            dist.send(mine, dst=0)
        dist.barrier()
        dist.destroy_process_group()
        return
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



def test_peer_combine_model(hs, orc, default_scene, tmp_path):
    """The fused reduce + resolve as a CPU model over gloo: pixel slices per rank, partial sums added in rank order,
    quantised per slice -- the image equals print_frame_buffer of the ordered sum of the ranks' accumulators."""
    import torch.multiprocessing as mp
    from raydar_b200 import dist as rdist
    world, spp = 2, 6
    out = str(tmp_path / "img_peer.npy")
    mp.spawn(_worker, args=(world, _free_port(), "peer", out), nprocs=world, join=True)
    got = np.load(out)
    scene = default_scene.with_resolution(107, 60)
    parts = []
    for r in range(world):
        first, count = rdist.strong_sample_range(r, world, spp)
        parts.append(hs.render(scene, 42, first, count, 12)[0])
    want = orc.resolve(parts[0] + parts[1], spp)
    assert np.array_equal(got, want)
    one = orc.resolve(orc.render(scene, 42, 0, spp, 12, n_threads=2), spp)
    assert np.abs(got.astype(int) - one.astype(int)).max() <= 1
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 6420, 1920 * 1080):
            sl = [rdist.peer_pixel_slice(r, world, n) for r in range(world)]
            assert sl[0][0] == 0 and all(sl[r][0] + sl[r][1] == (sl[r + 1][0] if r + 1 < world else n) for r in range(world))
