"""-m gpu, needs >= 2 devices (skipped otherwise): the single-process multi-GPU handle (rdr_create_multi:
sample-range sharding + one grouped ncclReduce onto device 0) against a single-GPU render of the same samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_multi_gpu_matches_single(rb, orc, benchmark_scene, n_dev):
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    scene = benchmark_scene.with_resolution(640, 360)
    spp = 16
    one = rb.Renderer(rb.RendererConfig(spp, 12)); one.set_seed(77)
    img1 = one.render_frame(scene); acc1 = one.read_accum()
    multi = rb.Renderer(rb.RendererConfig(spp, 12), devices=list(range(n_dev))); multi.set_seed(77)
    imgn = multi.render_frame(scene); accn = multi.read_accum()
    assert multi.sample_count() == spp
    assert np.array_equal(accn[..., 3], acc1[..., 3])                    # every pixel got all 16 samples once
    assert np.allclose(accn, acc1, rtol=1e-5, atol=1e-5)                 # equal up to f32 summation order
    assert np.abs(imgn.astype(int) - img1.astype(int)).max() <= 1
    # progressive path on the multi handle: one sample per device per call, None when exhausted
    multi.new_frame(scene)
    calls = 0
    while multi.render_sample(scene) is not None:
        calls += 1
    assert calls == spp // n_dev and multi.sample_count() == spp
    assert np.allclose(multi.read_accum(), acc1, rtol=1e-5, atol=1e-5)
    one.close(); multi.close()


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_multi_gpu_stripes_bit_identical(rb, benchmark_scene, n_dev):
    """rdr_set_partition(STRIPES): round-robin 16-row stripes per device, all samples each; the ncclReduce adds zeros,
    so accumulator and image are bit-identical to one GPU (sample-range sharding is only equal up to summation order)."""
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    scene = benchmark_scene.with_resolution(640, 360)
    spp = 8
    one = rb.Renderer(rb.RendererConfig(spp, 12)); one.set_seed(77)
    img1 = one.render_frame(scene); acc1 = one.read_accum()
    multi = rb.Renderer(rb.RendererConfig(spp, 12), devices=list(range(n_dev))); multi.set_seed(77)
    multi.set_partition(rb.PARTITION_STRIPES, 16)
    imgn = multi.render_frame(scene); accn = multi.read_accum()
    assert multi.sample_count() == spp
    assert np.array_equal(accn.view(np.uint32), acc1.view(np.uint32))
    assert np.array_equal(imgn, img1)
    multi.new_frame(scene)                                   # progressive: every call adds one sample on every device
    calls = 0
    while multi.render_sample(scene) is not None:
        calls += 1
    assert calls == spp and multi.sample_count() == spp
    assert np.array_equal(multi.read_accum().view(np.uint32), acc1.view(np.uint32))
    one.close(); multi.close()
