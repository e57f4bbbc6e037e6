// rdr_multi.cpp -- see rdr_multi.h.
//
// Combining the per-GPU accumulators.  Two implementations behind rdr_set_combine:
//
//   PEER (default when every pair of devices has peer access -- NVLink / NVSwitch on a B200 box):
//     the devices map each other's accumulators (cudaDeviceEnablePeerAccess) and ONE kernel per device
//     (peer_combine_kernel, rdr_kernels.cu) sums that device's 1/G of the pixels over all accumulators, quantises
//     (print_frame_buffer, cpu.rs:221-230) and stores the RGBA8 words straight into the pinned host image.  The kernels
//     are ordered behind the render kernels of ALL devices with cross-device event waits, so a frame is: G render
//     launches, G combine launches, one wait -- no host round trip between rendering and the image, nothing funnels
//     through device 0, and 4x fewer bytes leave the GPUs than with an f32 reduce.  With the STRIPES partition a
//     device's pixels are its own stripes and it reads only its own accumulator.
//   NCCL: one grouped ncclReduce(sum, f32) onto devices[0], then resolve + copy there.  The fallback without peer
//     access, and the independent check of the PEER path in tests/test_gpu_multi.py.  NCCL is loaded with dlopen at
//     first use so that the library has no link-time NCCL dependency (a host process that already loaded an NCCL,
//     e.g. through torch.distributed, shares that copy).
#include "rdr_multi.h"
#include "rdr_layout.h"
#include "rdr_launch.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
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
    std::vector<int> devices;
    std::vector<ncclComm_t> comm;            // created at the first NCCL combine
    std::vector<uint32_t> begin, count;      // sample range per child for the current frame
    RdrConfig config{1024u, 12u};
    uint64_t seed = 0x5EEDull;
    int accel = RDR_ACCEL_AUTO;
    f4 *root_sum = nullptr;                  // on child[0]'s device: the summed accumulator (NCCL combine, rdr_read_accum)
    size_t root_capacity = 0;
    bool reduced = false;                    // root_sum holds the sum of the current accumulators
    uint32_t sample_count = 0;
    int partition = RDR_PARTITION_SAMPLES, frame_partition = RDR_PARTITION_SAMPLES;   // requested / of the current frame
    uint32_t stripe_rows = 16;
    bool peer_ok = false;                    // every pair of devices has peer access enabled
    int combine = RDR_COMBINE_AUTO;
    uint8_t *staging = nullptr;              // pinned image the combine kernels write when the caller's buffer is pageable
    size_t staging_bytes = 0;
    uint32_t next_turn = 0;                  // SAMPLES partition: whose turn the next progressive sample is
};

static int child_fail(RdrRenderer *owner, RdrRenderer *c, int st)
{
    return api_fail(owner, st, rdr_last_error(c));
}

static bool use_peer(const MultiGpu *m)
{
    return m->child.size() > 1u && m->peer_ok && m->combine != RDR_COMBINE_NCCL;
}

// maps every device's memory into every other device's address space (NVLink / NVSwitch peer access)
static bool enable_peer_access(const std::vector<int> &devices)
{
    for (int a : devices)
        for (int b : devices) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); return false; }
        }
    for (int a : devices) {
        if (cudaSetDevice(a) != cudaSuccess) { cudaGetLastError(); return false; }
        for (int b : devices) {
            if (a == b) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
            cudaGetLastError();      // clears cudaErrorPeerAccessAlreadyEnabled
        }
    }
    return true;
}

static int ensure_nccl(RdrRenderer *owner, MultiGpu *m)
{
    if (!m->comm.empty() || m->child.size() < 2u) return RDR_OK;
    NcclApi &api = nccl();
    if (!api.ok) return api_fail(owner, RDR_ERR_NCCL, "cannot load libnccl.so.2");
    m->comm.resize(m->devices.size());
    ncclResult_t r = api.CommInitAll(m->comm.data(), (int)m->devices.size(), m->devices.data());
    if (r != ncclSuccess) {
        m->comm.clear();
        const std::string msg = std::string("ncclCommInitAll: ") + api.GetErrorString(r);
        return api_fail(owner, RDR_ERR_NCCL, msg.c_str());
    }
    return RDR_OK;
}

int multi_create(RdrRenderer *owner, const RdrConfig *config, int n_devices, const int *devices, MultiGpu **out)
{
    if (n_devices < 1 || !devices) return api_fail(owner, RDR_ERR_INVALID, "need at least one device");
    if ((uint32_t)n_devices > RDR_MAX_PEERS) return api_fail(owner, RDR_ERR_INVALID, "too many devices");
    MultiGpu *m = new MultiGpu();
    if (config) m->config = *config;
    m->devices.assign(devices, devices + n_devices);
    for (int g = 0; g < n_devices; ++g) {
        RdrRenderer *c = nullptr;
        int st = rdr_create(&m->config, devices[g], &c);
        if (st != RDR_OK) { multi_destroy(m); return api_fail(owner, st, rdr_last_error(nullptr)); }
        m->child.push_back(c);
    }
    if (n_devices > 1) {
        m->peer_ok = enable_peer_access(m->devices);
        if (!m->peer_ok) {                       // no NVLink / PCIe peer path: the NCCL reduce is the only combine
            int st = ensure_nccl(owner, m);
            if (st != RDR_OK) { multi_destroy(m); return st; }
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
    if (m->staging) cudaFreeHost(m->staging);
    for (RdrRenderer *c : m->child) rdr_destroy(c);
    delete m;
}

void multi_set_config(MultiGpu *m, const RdrConfig &config) { m->config = config; for (RdrRenderer *c : m->child) rdr_set_max_bounces(c, config.max_bounces); }
void multi_set_seed(MultiGpu *m, uint64_t seed) { m->seed = seed; }
void multi_set_partition(MultiGpu *m, int partition, uint32_t stripe_rows) { m->partition = partition; m->stripe_rows = stripe_rows; }
void multi_set_accel(MultiGpu *m, int accel) { m->accel = accel; for (RdrRenderer *c : m->child) rdr_set_accel(c, accel); }

int multi_set_combine(RdrRenderer *owner, MultiGpu *m, int combine)
{
    if (combine == RDR_COMBINE_PEER && m->child.size() > 1u && !m->peer_ok)
        return api_fail(owner, RDR_ERR_UNSUPPORTED, "the devices of this handle have no peer access to each other");
    m->combine = combine;
    m->reduced = false;
    return RDR_OK;
}

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
    m->next_turn = 0;
    return RDR_OK;
}

// ---- rendering: every device is started before any is waited on, so the per-GPU kernels overlap ----------------
static int launch_shares(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> &share)
{
    for (size_t g = 0; g < m->child.size(); ++g) {
        int st = api_render_launch(m->child[g], share[g]);
        if (st != RDR_OK) return child_fail(owner, m->child[g], st);
    }
    m->reduced = false;
    return RDR_OK;
}

static int finish_shares(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> &share)
{
    const bool stripes = m->frame_partition == RDR_PARTITION_STRIPES;
    for (size_t g = 0; g < m->child.size(); ++g) {
        int st = api_render_finish(m->child[g], share[g]);
        if (st != RDR_OK) return child_fail(owner, m->child[g], st);
        if (!stripes || g == 0u) m->sample_count += share[g];     // stripes: every device renders the same sample indices
    }
    return RDR_OK;
}

static int render_shares(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> &share)
{
    int st = launch_shares(owner, m, share);
    return st != RDR_OK ? st : finish_shares(owner, m, share);
}

// n more samples of the frame over the devices (SAMPLES: n in total, as evenly as the devices' remaining shares allow;
// STRIPES: n on every device, each for its own rows)
static std::vector<uint32_t> split_samples(const MultiGpu *m, uint32_t n)
{
    const uint32_t G = (uint32_t)m->child.size();
    std::vector<uint32_t> share(G, 0), left(G, 0);
    if (m->frame_partition == RDR_PARTITION_STRIPES) {
        for (uint32_t g = 0; g < G; ++g) share[g] = std::min(n, api_samples_left(m->child[g]));
        return share;
    }
    uint64_t total_left = 0;
    for (uint32_t g = 0; g < G; ++g) { left[g] = api_samples_left(m->child[g]); total_left += left[g]; }
    uint64_t todo = std::min<uint64_t>(n, total_left);
    while (todo > 0u) {                              // every pass hands each device with room an equal part of what is still to do
        uint32_t open = 0;
        for (uint32_t g = 0; g < G; ++g) open += left[g] > share[g] ? 1u : 0u;
        const uint64_t part = std::max<uint64_t>(1u, todo / open);
        for (uint32_t g = 0; g < G && todo > 0u; ++g) {
            const uint64_t take = std::min<uint64_t>(std::min<uint64_t>(part, left[g] - share[g]), todo);
            share[g] += (uint32_t)take; todo -= take;
        }
    }
    return share;
}

int multi_render_samples(RdrRenderer *owner, MultiGpu *m, uint32_t n)
{
    return render_shares(owner, m, split_samples(m, n));
}

// ---- PEER combine ------------------------------------------------------------------------------------------------
// Queues one peer_combine_kernel per device.  wait_for: the shares just launched (NULL: nothing in flight) -- the kernel
// of device g waits, on the device, for the render kernels of every device that rendered.
// dst_rgba / dst_sum: where the pixels go (one of them), indexed by global pixel.
static int queue_peer_combine(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> *wait_for, float divisor, uchar4 *dst_rgba, f4 *dst_sum)
{
    const uint32_t G = (uint32_t)m->child.size();
    const uint32_t n_pixels = api_pixels(m->child[0]);
    const bool stripes = m->frame_partition == RDR_PARTITION_STRIPES;
    for (uint32_t g = 0; g < G; ++g) {
        RdrRenderer *c = m->child[g];
        if (cudaSetDevice(api_device(c)) != cudaSuccess) return api_fail(owner, RDR_ERR_CUDA, "cudaSetDevice failed");
        PeerCombine C{};
        C.width = api_width(c); C.divisor = divisor; C.rgba = dst_rgba; C.sum = dst_sum;
        if (stripes) {                       // the device's own stripes hold everything there is about its pixels
            C.src[0] = api_accum(c); C.n_src = 1u;
            C.stripe_rows = m->stripe_rows; C.stripe_index = g; C.stripe_count = G;
            C.count = stripe_owned_pixels(api_width(c), api_height(c), m->stripe_rows, g, G);
        } else {
            for (uint32_t h = 0; h < G; ++h) C.src[h] = api_accum(m->child[h]);
            C.n_src = G; C.stripe_count = 1u;
            C.first = (uint32_t)((uint64_t)n_pixels * g / G);
            C.count = (uint32_t)((uint64_t)n_pixels * (g + 1u) / G) - C.first;
            if (wait_for)
                for (uint32_t h = 0; h < G; ++h)
                    if (h != g && (*wait_for)[h] != 0u && cudaStreamWaitEvent(api_stream(c), api_render_done_event(m->child[h]), 0) != cudaSuccess)
                        return api_fail(owner, RDR_ERR_CUDA, "cudaStreamWaitEvent failed");
        }
        if (C.count == 0u) continue;
        if (launch_peer_combine(C, api_stream(c)) != cudaSuccess) return api_fail(owner, RDR_ERR_CUDA, "peer_combine_kernel launch failed");
        api_count_launch(c);
    }
    return RDR_OK;
}

static int sync_all(RdrRenderer *owner, MultiGpu *m)
{
    for (RdrRenderer *c : m->child) {
        if (cudaSetDevice(api_device(c)) != cudaSuccess || cudaStreamSynchronize(api_stream(c)) != cudaSuccess)
            return api_fail(owner, RDR_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    }
    return RDR_OK;
}

// the image the combine kernels write: the caller's buffer when it is pinned (rdr_alloc_host_image), else the staging image
static int combine_target(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8, size_t bytes, uint8_t **target)
{
    if (is_host_image(rgba8, bytes)) { *target = rgba8; return RDR_OK; }
    if (bytes > m->staging_bytes) {
        if (m->staging) cudaFreeHost(m->staging);
        m->staging = nullptr; m->staging_bytes = 0;
        if (cudaHostAlloc((void **)&m->staging, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return api_fail(owner, RDR_ERR_NOMEM, "cudaHostAlloc(staging image) failed");
        }
        m->staging_bytes = bytes;
    }
    *target = m->staging;
    return RDR_OK;
}

// shares launched (or NULL) -> combined image in rgba8.  Finishes the shares' bookkeeping on the way.
static int peer_resolve(RdrRenderer *owner, MultiGpu *m, const std::vector<uint32_t> *in_flight, uint32_t divisor, uint8_t *rgba8)
{
    if (!rgba8) return api_fail(owner, RDR_ERR_INVALID, "output image is NULL");
    const size_t bytes = (size_t)api_pixels(m->child[0]) * 4u;
    if (bytes == 0u) return in_flight ? finish_shares(owner, m, *in_flight) : RDR_OK;
    uint8_t *target = nullptr;
    int st = combine_target(owner, m, rgba8, bytes, &target);
    if (st != RDR_OK) return st;
    // the divisor is the frame's sample count AFTER the shares in flight (known before they finish)
    uint32_t div = divisor;
    if (div == 0u) {
        div = m->sample_count;
        if (in_flight) {
            const bool stripes = m->frame_partition == RDR_PARTITION_STRIPES;
            for (size_t g = 0; g < in_flight->size(); ++g) if (!stripes || g == 0u) div += (*in_flight)[g];
        }
    }
    if ((st = queue_peer_combine(owner, m, in_flight, (float)div, reinterpret_cast<uchar4 *>(target), nullptr)) != RDR_OK) return st;
    if (in_flight && (st = finish_shares(owner, m, *in_flight)) != RDR_OK) return st;
    if ((st = sync_all(owner, m)) != RDR_OK) return st;
    if (target != rgba8) memcpy(rgba8, target, bytes);
    return RDR_OK;
}

// ---- NCCL combine --------------------------------------------------------------------------------------------------
// one ncclReduce(sum, f32) of the per-GPU accumulators onto root_sum (device 0); the children keep their partial sums
static int nccl_reduce_to_root(RdrRenderer *owner, MultiGpu *m)
{
    int st = ensure_nccl(owner, m);
    if (st != RDR_OK) return st;
    const uint32_t G = (uint32_t)m->child.size();
    NcclApi &api = nccl();
    const size_t count = (size_t)api_pixels(m->child[0]) * 4u;
    ncclResult_t r = api.GroupStart();
    for (uint32_t g = 0; g < G && r == ncclSuccess; ++g)
        r = api.Reduce(api_accum(m->child[g]), m->root_sum, count, ncclFloat32, ncclSum, 0, m->comm[g], api_stream(m->child[g]));
    ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return api_fail(owner, RDR_ERR_NCCL, api.GetErrorString(r));
    return RDR_OK;
}

// the summed accumulator on device 0 (rdr_read_accum, NCCL resolve)
static int sum_to_root(RdrRenderer *owner, MultiGpu *m, const f4 **src)
{
    const uint32_t G = (uint32_t)m->child.size();
    if (G == 1) { *src = api_accum(m->child[0]); return RDR_OK; }
    if (!m->reduced) {
        int st;
        if (use_peer(m)) {                   // every device writes its share of the sums into device 0's buffer over NVLink
            if ((st = queue_peer_combine(owner, m, nullptr, 1.0f, nullptr, m->root_sum)) != RDR_OK) return st;
            if ((st = sync_all(owner, m)) != RDR_OK) return st;
        } else if ((st = nccl_reduce_to_root(owner, m)) != RDR_OK) return st;
        m->reduced = true;
    }
    *src = m->root_sum;
    return RDR_OK;
}

int multi_resolve(RdrRenderer *owner, MultiGpu *m, uint32_t divisor, uint8_t *rgba8)
{
    if (m->child.size() > 1u && use_peer(m)) return peer_resolve(owner, m, nullptr, divisor, rgba8);
    const f4 *src = nullptr;
    int st = sum_to_root(owner, m, &src);
    if (st != RDR_OK) return st;
    st = api_resolve_from(m->child[0], src, divisor ? divisor : m->sample_count, rgba8);
    return st == RDR_OK ? RDR_OK : child_fail(owner, m->child[0], st);
}

// render_sample (cpu.rs:142-158): exactly ONE more sample per call.  STRIPES: every device renders it for its own rows;
// SAMPLES: the devices take turns (each draws from its own sample range, so after all calls the frame is the one
// render_frame produces).
int multi_render_sample(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8, int *produced)
{
    const uint32_t G = (uint32_t)m->child.size();
    std::vector<uint32_t> share(G, 0);
    uint32_t total = 0;
    if (m->frame_partition == RDR_PARTITION_STRIPES) {
        for (uint32_t g = 0; g < G; ++g) { share[g] = std::min(1u, api_samples_left(m->child[g])); total += share[g]; }
    } else {
        for (uint32_t k = 0; k < G && total == 0u; ++k) {
            const uint32_t g = (m->next_turn + k) % G;
            if (api_samples_left(m->child[g]) != 0u) { share[g] = 1u; total = 1u; m->next_turn = (g + 1u) % G; }
        }
    }
    if (total == 0u) return RDR_OK;                      // `None`
    int st;
    if (G > 1u && use_peer(m)) {
        if ((st = launch_shares(owner, m, share)) != RDR_OK) return st;
        if ((st = peer_resolve(owner, m, &share, 0u, rgba8)) != RDR_OK) return st;
    } else {
        if ((st = render_shares(owner, m, share)) != RDR_OK) return st;
        if ((st = multi_resolve(owner, m, 0u, rgba8)) != RDR_OK) return st;
    }
    if (produced) *produced = 1;
    return RDR_OK;
}

// every sample the frame has left, then the image: one launch per device + one combine kernel per device, one wait
int multi_finish_frame(RdrRenderer *owner, MultiGpu *m, uint8_t *rgba8)
{
    std::vector<uint32_t> share(m->child.size(), 0);
    for (size_t g = 0; g < m->child.size(); ++g) share[g] = api_samples_left(m->child[g]);
    int st;
    if (m->child.size() > 1u && use_peer(m)) {
        if ((st = launch_shares(owner, m, share)) != RDR_OK) return st;
        return peer_resolve(owner, m, &share, 0u, rgba8);
    }
    if ((st = render_shares(owner, m, share)) != RDR_OK) return st;
    return multi_resolve(owner, m, 0u, rgba8);
}

int multi_read_accum(RdrRenderer *owner, MultiGpu *m, float *dst)
{
    const f4 *src = nullptr;
    int st = sum_to_root(owner, m, &src);
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
int multi_combine_in_use(const MultiGpu *m) { return m->child.size() < 2u ? RDR_COMBINE_AUTO : (use_peer(m) ? RDR_COMBINE_PEER : RDR_COMBINE_NCCL); }

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
