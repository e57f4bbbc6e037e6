"""Row-stripe (image-tile) sharding, SURVEY.md 8e "alternative": every shard renders ALL samples of its round-robin
row stripes, the accumulators are disjoint, so the reduce adds zeros and the result is bit-identical to one GPU.
CPU part: the product's stripe_pixel()/stripe_owned_pixels() (rdr_layout.h, compiled into tests/hostsim) against
the host-side rule in raydar_b200/dist.py, the per-lane render restricted to stripes against the oracle, and the
world_size-2 gloo reduce.  The GPU twins are in test_gpu_parity.py / test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (107, 60), (64, 33), (1920, 1080), (33, 16), (5, 17)])
@pytest.mark.parametrize("rows", [1, 3, 16])
@pytest.mark.parametrize("count", [1, 2, 3, 8])
def test_stripes_partition_the_image(hs, w, h, rows, count):
    from raydar_b200 import dist as rdist
    seen = np.zeros(w * h, np.int32)
    for index in range(count):
        px = hs.stripe_pixels(w, h, rows, index, count)
        assert np.all(np.diff(px.astype(np.int64)) > 0)                   # handed out in ascending (row-major) order
        want_rows = rdist.stripe_rows_owned(index, count, h, rows)
        want = (np.array(want_rows, np.int64)[:, None] * w + np.arange(w)[None, :]).reshape(-1)
        assert np.array_equal(px.astype(np.int64), want)
        seen[px] += 1
    assert np.all(seen == 1)                                              # every pixel belongs to exactly one shard


def test_more_shards_than_stripes_leaves_some_empty(hs):
    assert [len(hs.stripe_pixels(10, 20, 16, i, 4)) for i in range(4)] == [160, 40, 0, 0]
    assert len(hs.stripe_pixels(10, 20, 0, 0, 4)) == 200                  # rows == 0: whole image
    assert len(hs.stripe_pixels(10, 20, 16, 0, 1)) == 200                 # one shard: whole image


@pytest.mark.parametrize("count,rows", [(2, 16), (3, 7)])
def test_stripe_shards_sum_bit_identical(hs, orc, default_scene, count, rows):
    scene = default_scene.with_resolution(107, 61)
    spp, seed = 5, 99
    want = orc.render(scene, seed, 0, spp, 12, n_threads=2)
    total = np.zeros_like(want)
    for index in range(count):
        acc = hs.render_stripes(scene, seed, 0, spp, 12, rows, index, count)
        mine = np.zeros(61, bool); mine[[y for y in range(61) if (y // rows) % count == index]] = True
        assert np.all(u32(acc[~mine]) == 0)                               # rows of other shards stay untouched
        assert np.array_equal(u32(acc[mine]), u32(want[mine]))
        total += acc
    assert np.array_equal(u32(total), u32(want))                          # x + 0 == x: no summation-order noise


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    import torch
    import torch.distributed as dist
    from oracle import orc
    import hostsim_py as hs
    from raydar_b200 import dist as rdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = orc.load_rscn(os.path.join(ROOT, "scenes", "default.rscn")).with_resolution(107, 60)
    acc = hs.render_stripes(scene, 42, 0, 6, 12, rdist.STRIPE_ROWS, rank, world)
    t = torch.from_numpy(acc.reshape(-1))
    rdist.reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy().reshape(acc.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_stripe_sharding_and_reduce_gloo(orc, default_scene, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "acc_stripes.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    want = orc.render(default_scene.with_resolution(107, 60), 42, 0, 6, 12, n_threads=2)
    assert np.array_equal(u32(got), u32(want))                            # bit-identical, unlike sample-range sharding
