// rdr_launch.h -- host-callable launch wrappers implemented in rdr_kernels.cu
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "raydar_cuda.h"

#define RDR_BLOCK 256u

namespace rdr {

struct FrameParams;
struct SceneLayout;
struct f4;

// multi-GPU combine (peer_combine_kernel, rdr_kernels.cu): this GPU's share of the pixels, summed over n_src accumulators
constexpr uint32_t RDR_MAX_PEERS = 16u;
struct PeerCombine {
    const f4 *src[RDR_MAX_PEERS];    // accumulators in summation order (own and NVLink-mapped peers)
    uint32_t n_src;
    uint32_t first, count;           // contiguous share: pixels [first, first + count)
    uint32_t width, stripe_rows, stripe_index, stripe_count;   // stripe_count > 1: the count pixels of this shard's row stripes instead
    float divisor;
    uchar4 *rgba;                    // W*H image indexed by global pixel (pinned host memory or device memory), or NULL
    f4 *sum;                         // W*H f32 sums indexed by global pixel, or NULL
};
cudaError_t launch_peer_combine(const PeerCombine &C, cudaStream_t stream);

size_t scene_smem_bytes(const SceneLayout &L, bool staged, uint32_t block);
size_t fused_smem_bytes(const SceneLayout &L);       // dynamic shared memory of one CTA of the fused render kernel
cudaError_t launch_render(const FrameParams &P, int variant, int resident_ctas, cudaStream_t stream);
cudaError_t render_resident_ctas(const FrameParams &P, int variant, int *out);
cudaError_t launch_resolve(const f4 *accum, uchar4 *out, uint32_t n_pixels, float divisor, cudaStream_t stream);
cudaError_t launch_primary(const FrameParams &P, int variant, f4 *primary, int32_t *primary_idx, cudaStream_t stream);
cudaError_t launch_kat_trace(const FrameParams &P, int variant, uint32_t n, const float *rays, int32_t *ids, float *ts, cudaStream_t stream);
cudaError_t launch_trace_path(const FrameParams &P, int variant, uint32_t x, uint32_t y, uint32_t sample,
                              RdrPathStep *steps, uint32_t capacity, uint32_t *n_steps, float *rgba, cudaStream_t stream);
cudaError_t launch_kat_hit(bool sphere, uint32_t n, const float *rays, const float *prims, float *t, int32_t *hit, cudaStream_t stream);
cudaError_t launch_kat_vec(int op, uint32_t n, const float *in, float *out, cudaStream_t stream);
cudaError_t launch_kat_camera_rays(const FrameParams &P, uint32_t n, const uint32_t *xy, float *rays, cudaStream_t stream);
cudaError_t launch_kat_rng(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block,
                           uint32_t *out, cudaStream_t stream);

}  // namespace rdr
