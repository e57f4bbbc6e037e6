"""-m gpu, needs >= 2 devices (skipped otherwise): the single-process multi-GPU handle (rdr_create_multi) against a
single-GPU render of the same samples, for both combines -- the fused reduce + resolve over NVLink peer memory (PEER,
the default) and the grouped ncclReduce onto device 0 (NCCL) -- and both partitions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


def ordered_sum(parts):
    """((a0 + a1) + a2) + ... in f32: the order of peer_combine_kernel (device 0 first)."""
    total = parts[0].copy()
    for p in parts[1:]:
        total = total + p
    return total


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_dev", [2, 4, 8])
@pytest.mark.parametrize("combine", ["peer", "nccl"])
def test_multi_gpu_matches_single(rb, orc, benchmark_scene, n_dev, combine):
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    scene = benchmark_scene.with_resolution(640, 360)
    spp = 16
    one = rb.Renderer(rb.RendererConfig(spp, 12)); one.set_seed(77)
    img1 = one.render_frame(scene); acc1 = one.read_accum()
    multi = rb.Renderer(rb.RendererConfig(spp, 12), devices=list(range(n_dev))); multi.set_seed(77)
    multi.set_combine(rb.COMBINE_PEER if combine == "peer" else rb.COMBINE_NCCL)
    assert multi.combine_in_use() == (rb.COMBINE_PEER if combine == "peer" else rb.COMBINE_NCCL)
    imgn = multi.render_frame(scene); accn = multi.read_accum()
    assert multi.sample_count() == spp
    assert np.array_equal(accn[..., 3], acc1[..., 3])                    # every pixel got all 16 samples once
    assert np.allclose(accn, acc1, rtol=1e-5, atol=1e-5)                 # equal up to f32 summation order
    assert np.abs(imgn.astype(int) - img1.astype(int)).max() <= 1
    if combine == "peer":
        # the PEER combine adds the devices' partial sums in device order: reproduce it exactly from single-GPU renders
        # of the same sample ranges, and the image from that sum
        parts = []
        for g in range(n_dev):
            b, e = spp * g // n_dev, spp * (g + 1) // n_dev
            one.set_max_sample_count(e - b); one.set_sample_offset(b); one.new_frame(scene); one.render_samples(e - b)
            parts.append(one.read_accum())
        one.set_max_sample_count(spp); one.set_sample_offset(0)
        want = ordered_sum(parts)
        assert np.array_equal(u32(accn), u32(want))
        assert np.array_equal(imgn, orc.resolve(want, spp))
        # a pinned image (the GPUs write it directly) and a pageable one (staging copy) hold the same bytes
        pinned = rb.HostImage(360, 640)
        multi.render_frame(scene, out=pinned.array)
        assert np.array_equal(pinned.array, imgn)
        pinned.close()
    # progressive path on the multi handle: exactly one sample per call (the devices take turns), None when exhausted
    multi.new_frame(scene)
    calls = 0
    last = None
    while True:
        img = multi.render_sample(scene)
        if img is None:
            break
        last = img
        calls += 1
        assert multi.sample_count() == calls
    assert calls == spp
    assert np.allclose(multi.read_accum(), acc1, rtol=1e-5, atol=1e-5)
    assert np.abs(last.astype(int) - img1.astype(int)).max() <= 1
    # render_samples(n) renders exactly n more samples while the frame has them, whatever the devices' shares look like
    multi.new_frame(scene)
    multi.render_samples(5); assert multi.sample_count() == 5
    multi.render_samples(3); assert multi.sample_count() == 8
    multi.render_samples(1000); assert multi.sample_count() == spp
    assert np.allclose(multi.read_accum(), acc1, rtol=1e-5, atol=1e-5)
    assert np.array_equal(multi.read_accum()[..., 3], acc1[..., 3])
    one.close(); multi.close()


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_dev", [2, 4, 8])
@pytest.mark.parametrize("combine", ["peer", "nccl"])
def test_multi_gpu_stripes_bit_identical(rb, benchmark_scene, n_dev, combine):
    """rdr_set_partition(STRIPES): round-robin 16-row stripes per device, all samples each.  PEER: every device resolves
    its own stripes from its own accumulator; NCCL: the reduce adds zeros.  Either way accumulator and image are
    bit-identical to one GPU (sample-range sharding is only equal up to summation order)."""
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    scene = benchmark_scene.with_resolution(640, 360)
    spp = 8
    one = rb.Renderer(rb.RendererConfig(spp, 12)); one.set_seed(77)
    img1 = one.render_frame(scene); acc1 = one.read_accum()
    multi = rb.Renderer(rb.RendererConfig(spp, 12), devices=list(range(n_dev))); multi.set_seed(77)
    multi.set_partition(rb.PARTITION_STRIPES, 16)
    multi.set_combine(rb.COMBINE_PEER if combine == "peer" else rb.COMBINE_NCCL)
    imgn = multi.render_frame(scene); accn = multi.read_accum()
    assert multi.sample_count() == spp
    assert np.array_equal(u32(accn), u32(acc1))
    assert np.array_equal(imgn, img1)
    multi.new_frame(scene)                                   # progressive: every call adds one sample on every device
    calls = 0
    last = None
    while True:
        img = multi.render_sample(scene)
        if img is None:
            break
        last = img
        calls += 1
    assert calls == spp and multi.sample_count() == spp
    assert np.array_equal(u32(multi.read_accum()), u32(acc1))
    assert np.array_equal(last, img1)
    # a resolution that does not divide into whole stripes, and a partial last stripe
    odd = benchmark_scene.with_resolution(333, 187)
    one.new_frame(odd); one.render_samples(spp)
    want = one.resolve()
    got = multi.render_frame(odd)
    assert np.array_equal(got, want)
    assert np.array_equal(u32(multi.read_accum()), u32(one.read_accum()))
    one.close(); multi.close()


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
def test_multi_gpu_settings_are_forwarded(rb, benchmark_scene):
    """rdr_set_accel on a multi-GPU handle reaches every device; a frame rendered with another search is the same frame."""
    scene = benchmark_scene.with_resolution(320, 180)
    multi = rb.Renderer(rb.RendererConfig(8, 12), devices=[0, 1]); multi.set_seed(5)
    multi.set_partition(rb.PARTITION_STRIPES, 16)
    ref = multi.render_frame(scene); ref_acc = multi.read_accum()
    multi.set_accel(rb.ACCEL_BVH_COOP)
    assert np.array_equal(multi.render_frame(scene), ref)
    assert np.array_equal(u32(multi.read_accum()), u32(ref_acc))
    multi.set_accel(rb.ACCEL_AUTO)
    multi.set_max_bounces(3)                                  # applies to the next launch
    a = multi.render_frame(scene)
    one = rb.Renderer(rb.RendererConfig(8, 3)); one.set_seed(5)
    assert np.array_equal(a, one.render_frame(scene))
    one.close(); multi.close()
