// rdr_multi.cpp -- see rdr_multi.h.  NCCL is loaded with dlopen at first use so that the library has
// no link-time NCCL dependency (a host process that already loaded an NCCL, e.g. through
// torch.distributed, shares that copy).
#include "rdr_multi.h"
#include "rdr_core.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <string>
#include <vector>

namespace rdr {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    if (api.handle) return api;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) return api;
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))dlsym(api.handle, "ncclReduce");
    api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.CommInitAll && api.CommDestroy && api.Reduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api;
}

}  // namespace

struct MultiGpu {
    std::vector<RdrRenderer *> child;
    std::vector<ncclComm_t> comm;
    std::vector<uint32_t> begin, count;      // sample range per child for the current frame
    RdrConfig config{1024u, 12u};
    uint64_t seed = 0x5EEDull;
    f4 *root_sum = nullptr;              // on child[0]'s device
    size_t root_capacity = 0;
    bool reduced = false;                    // root_sum holds the sum of the current accumulators
    uint32_t sample_count = 0;
    int partition = RDR_PARTITION_SAMPLES, frame_partition = RDR_PARTITION_SAMPLES;   // requested / of the current frame
    uint32_t stripe_rows = 16;
};

static int child_fail(RdrRenderer *owner, RdrRenderer *c, int st)
{
    return api_fail(owner, st, rdr_last_error(c));
}

int multi_create(RdrRenderer *owner, const RdrConfig *config, int n_devices, const int *devices, MultiGpu **out)
{
    if (n_devices < 1 || !devices) return api_fail(owner, RDR_ERR_INVALID, "need at least one device");
    MultiGpu *m = new MultiGpu();
    if (config) m->config = *config;
    for (int g = 0; g < n_devices; ++g) {
        RdrRenderer *c = nullptr;
        int st = rdr_create(&m->config, devices[g], &c);
        if (st != RDR_OK) { multi_destroy(m); return api_fail(owner, st, rdr_last_error(nullptr)); }
        m->child.push_back(c);
    }
    if (n_devices > 1) {
        NcclApi &api = nccl();
        if (!api.ok) { multi_destroy(m); return api_fail(owner, RDR_ERR_NCCL, "cannot load libnccl.so.2"); }
        m->comm.resize(n_devices);
        ncclResult_t r = api.CommInitAll(m->comm.data(), n_devices, devices);
        if (r != ncclSuccess) {
            m->comm.clear();
            std::string msg = std::string("ncclCommInitAll: ") + api.GetErrorString(r);
            multi_destroy(m);
            return api_fail(owner, RDR_ERR_NCCL, msg.c_str());
        }
    }
    *out = m;
    return RDR_OK;
}

void multi_destroy(MultiGpu *m)
{
    if (!m) return;
    for (ncclComm_t c : m->comm) nccl().CommDestroy(c);
    if (m->root_sum && !m->child.empty()) { cudaSetDevice(api_device(m->child[0])); cudaFree(m->root_sum); }
    for (RdrRenderer *c : m->child) rdr_destroy(c);
    delete m;
}

void multi_set_config(MultiGpu *m, const RdrConfig &config) { m->config = config; for (RdrRenderer *c : m->child) rdr_set_max_bounces(c, config.max_bounces); }
void multi_set_seed(MultiGpu *m, uint64_t seed) { m->seed = seed; }
void multi_set_partition(MultiGpu *m, int partition, uint32_t stripe_rows) { m->partition = partition; m->stripe_rows = stripe_rows; }

int multi_new_frame(RdrRenderer *owner, MultiGpu *m, const RdrSceneFlat *scene)
{
    const uint32_t G = (uint32_t)m->child.size();
    const uint64_t S = m->config.max_sample_count;
    m->begin.assign(G, 0); m->count.assign(G, 0);
    m->frame_partition = m->partition;
    const bool stripes = m->frame_partition == RDR_PARTITION_STRIPES;
    for (uint32_t g = 0; g < G; ++g) {
        // SAMPLES: device g renders sample indices [S g / G, S (g + 1) / G) of every pixel;
        // STRIPES: device g renders every sample of its round-robin row stripes
        const uint32_t b = stripes ? 0u : (uint32_t)(S * g / G), e = stripes ? (uint32_t)S : (uint32_t)(S * (g + 1) / G);
        m->begin[g] = b; m->count[g] = e - b;
        RdrRenderer *c = m->child[g];
        rdr_set_max_sample_count(c, e - b);
        rdr_set_max_bounces(c, m->config.max_bounces);
        rdr_set_seed(c, m->seed);
        rdr_set_sample_offset(c, b);
        rdr_set_row_stripes(c, stripes ? m->stripe_rows : 0u, g, stripes ? G : 1u);
        int st = rdr_new_frame(c, scene);
        if (st != RDR_OK) return child_fail(owner, c, st);
    }
    const size_t n_pixels = api_pixels(m->child[0]);
    if (G > 1 && n_pixels > m->root_capacity) {
        cudaSetDevice(api_device(m->child[0]));
        if (m->root_sum) cudaFree(m->root_sum);
        m->root_sum = nullptr; m->root_capacity = 0;
        if (cudaMalloc(&m->root_sum, n_pixels * sizeof(f4)) != cudaSuccess) return api_fail(owner, RDR_ERR_NOMEM, "cudaMalloc(root_sum) failed");
        m->root_capacity = n_pixels;
    }
    m->reduced = false;
    m->sample_count = 0;
    return RDR_OK;
}

// starts every device, then waits for each: the per-GPU kernels overlap
static int render_shares(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> &share)
{
    const uint32_t G = (uint32_t)m->child.size();
    for (uint32_t g = 0; g < G; ++g) {
        int st = api_render_launch(m->child[g], share[g]);
        if (st != RDR_OK) return child_fail(owner, m->child[g], st);
    }
    for (uint32_t g = 0; g < G; ++g) {
        int st = api_render_finish(m->child[g], share[g]);
        if (st != RDR_OK) return child_fail(owner, m->child[g], st);
        if (m->frame_partition != RDR_PARTITION_STRIPES || g == 0u) m->sample_count += share[g];   // stripes: every device renders the same sample indices
    }
    m->reduced = false;
    return RDR_OK;
}

int multi_render_samples(RdrRenderer *owner, MultiGpu *m, uint32_t n)
{
    const uint32_t G = (uint32_t)m->child.size();
    std::vector<uint32_t> share(G, 0);
    const bool stripes = m->frame_partition == RDR_PARTITION_STRIPES;
    for (uint32_t g = 0; g < G; ++g) share[g] = std::min(stripes ? n : n / G + (g < n % G ? 1u : 0u), api_samples_left(m->child[g]));
    return render_shares(owner, m, share);
}

// one ncclReduce(sum, f32) of the per-GPU accumulators onto root_sum (device 0); the children keep their partial sums
static int reduce_to_root(RdrRenderer *owner, MultiGpu *m, const f4 **src)
{
    const uint32_t G = (uint32_t)m->child.size();
    if (G == 1) { *src = api_accum(m->child[0]); return RDR_OK; }
    if (!m->reduced) {
        NcclApi &api = nccl();
        const size_t count = (size_t)api_pixels(m->child[0]) * 4u;
        ncclResult_t r = api.GroupStart();
        for (uint32_t g = 0; g < G && r == ncclSuccess; ++g)
            r = api.Reduce(api_accum(m->child[g]), m->root_sum, count, ncclFloat32, ncclSum, 0, m->comm[g], api_stream(m->child[g]));
        ncclResult_t r2 = api.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) return api_fail(owner, RDR_ERR_NCCL, api.GetErrorString(r));
        m->reduced = true;
    }
    *src = m->root_sum;
    return RDR_OK;
}

int multi_resolve(RdrRenderer *owner, MultiGpu *m, uint32_t divisor, uint8_t *rgba8)
{
    const f4 *src = nullptr;
    int st = reduce_to_root(owner, m, &src);
    if (st != RDR_OK) return st;
    st = api_resolve_from(m->child[0], src, divisor ? divisor : m->sample_count, rgba8);
    return st == RDR_OK ? RDR_OK : child_fail(owner, m->child[0], st);
}

int multi_render_sample(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8, int *produced)
{
    const uint32_t G = (uint32_t)m->child.size();
    std::vector<uint32_t> share(G, 0);
    uint32_t total = 0;
    for (uint32_t g = 0; g < G; ++g) { share[g] = std::min(1u, api_samples_left(m->child[g])); total += share[g]; }
    if (total == 0u) return RDR_OK;                      // `None`
    int st = render_shares(owner, m, share);
    if (st != RDR_OK) return st;
    if ((st = multi_resolve(owner, m, 0u, rgba8)) != RDR_OK) return st;
    if (produced) *produced = 1;
    return RDR_OK;
}

int multi_render_frame(RdrRenderer *owner, MultiGpu *m, const RdrSceneFlat *scene, uint8_t *rgba8)
{
    int st = multi_new_frame(owner, m, scene);
    if (st != RDR_OK) return st;
    if ((st = render_shares(owner, m, m->count)) != RDR_OK) return st;
    return multi_resolve(owner, m, 0u, rgba8);
}

int multi_read_accum(RdrRenderer *owner, MultiGpu *m, float *dst)
{
    const f4 *src = nullptr;
    int st = reduce_to_root(owner, m, &src);
    if (st != RDR_OK) return st;
    cudaSetDevice(api_device(m->child[0]));
    const size_t bytes = (size_t)api_pixels(m->child[0]) * sizeof(f4);
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, api_stream(m->child[0])) != cudaSuccess ||
        cudaStreamSynchronize(api_stream(m->child[0])) != cudaSuccess)
        return api_fail(owner, RDR_ERR_CUDA, "accumulator read-back failed");
    return RDR_OK;
}

int multi_synchronize(RdrRenderer *owner, MultiGpu *m)
{
    for (RdrRenderer *c : m->child) { int st = rdr_synchronize(c); if (st != RDR_OK) return child_fail(owner, c, st); }
    return RDR_OK;
}

uint64_t multi_scene_device_bytes(const MultiGpu *m) { return m->child.empty() ? 0u : rdr_scene_device_bytes(m->child[0]); }

uint64_t multi_launch_count(const MultiGpu *m)
{
    uint64_t n = 0;
    for (RdrRenderer *c : m->child) n += rdr_launch_count(c);
    return n;
}

uint32_t multi_sample_count(const MultiGpu *m) { return m->sample_count; }

int multi_profiler(const MultiGpu *m, RdrProfiler *out)
{
    int st = rdr_profiler(m->child[0], out);
    for (RdrRenderer *c : m->child) out->device_render_ms = std::max(out->device_render_ms, api_device_ms(c));
    return st;
}

}  // namespace rdr

extern "C" int rdr_create_multi(const RdrConfig *config, int n_devices, const int *devices, RdrRenderer **out)
{
    if (!out) return RDR_ERR_INVALID;
    *out = nullptr;
    if (n_devices < 1 || !devices) return rdr::api_fail(nullptr, RDR_ERR_INVALID, "need at least one device");
    RdrRenderer *owner = nullptr;
    int st = rdr_create(config, devices[0], &owner);
    if (st != RDR_OK) return st;
    rdr::MultiGpu *m = nullptr;
    st = rdr::multi_create(owner, config, n_devices, devices, &m);
    if (st != RDR_OK) { rdr::api_fail(nullptr, st, rdr_last_error(owner)); rdr_destroy(owner); return st; }
    rdr::api_attach_multi(owner, m);
    *out = owner;
    return RDR_OK;
}
