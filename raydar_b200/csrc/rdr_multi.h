// rdr_multi.h -- single-process multi-GPU renderer: one sub-renderer per device, sample-range
// sharding (or round-robin row stripes, rdr_set_partition); the per-GPU accumulators are combined by a fused reduce + resolve
// kernel over NVLink peer memory, or by one ncclReduce(sum, f32) onto devices[0] (rdr_set_combine).
// The reference has no multi-device path; this is the B200 extension behind rdr_create_multi.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "raydar_cuda.h"

namespace rdr {

struct MultiGpu;
struct f4;

int multi_create(RdrRenderer *owner, const RdrConfig *config, int n_devices, const int *devices, MultiGpu **out);
void multi_destroy(MultiGpu *m);
int multi_new_frame(RdrRenderer *owner, MultiGpu *m, const RdrSceneFlat *scene);
int multi_render_samples(RdrRenderer *owner, MultiGpu *m, uint32_t n);
int multi_render_sample(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8, int *produced);
int multi_finish_frame(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8);
int multi_resolve(RdrRenderer *owner, MultiGpu *m, uint32_t divisor, uint8_t *rgba8);
int multi_read_accum(RdrRenderer *owner, MultiGpu *m, float *dst);
int multi_synchronize(RdrRenderer *owner, MultiGpu *m);
uint64_t multi_launch_count(const MultiGpu *m);
uint64_t multi_scene_device_bytes(const MultiGpu *m);      // per device
uint32_t multi_sample_count(const MultiGpu *m);
int multi_profiler(const MultiGpu *m, RdrProfiler *out);
void multi_set_config(MultiGpu *m, const RdrConfig &config);
void multi_set_seed(MultiGpu *m, uint64_t seed);
void multi_set_partition(MultiGpu *m, int partition, uint32_t stripe_rows);
void multi_set_accel(MultiGpu *m, int accel);
int multi_set_combine(RdrRenderer *owner, MultiGpu *m, int combine);
int multi_combine_in_use(const MultiGpu *m);               // RDR_COMBINE_PEER / RDR_COMBINE_NCCL (AUTO: one device)

// hooks implemented in rdr_api.cpp
int api_fail(RdrRenderer *r, int status, const char *msg);
f4 *api_accum(RdrRenderer *r);
cudaStream_t api_stream(RdrRenderer *r);
int api_device(RdrRenderer *r);
uint32_t api_pixels(RdrRenderer *r);
double api_device_ms(RdrRenderer *r);
uint32_t api_samples_left(RdrRenderer *r);
int api_render_launch(RdrRenderer *r, uint32_t n);
int api_render_finish(RdrRenderer *r, uint32_t n);
int api_resolve_from(RdrRenderer *r, const f4 *src, uint32_t divisor, uint8_t *rgba8);
void api_attach_multi(RdrRenderer *r, MultiGpu *m);
cudaEvent_t api_render_done_event(RdrRenderer *r);        // recorded behind the last render launch on the child's stream
uchar4 *api_rgba(RdrRenderer *r);
uint32_t api_width(RdrRenderer *r);
uint32_t api_height(RdrRenderer *r);
void api_count_launch(RdrRenderer *r);
bool is_host_image(const void *p, size_t bytes);          // inside a buffer from rdr_alloc_host_image

}  // namespace rdr
