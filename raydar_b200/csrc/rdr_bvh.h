// rdr_bvh.h -- host-side builder of the 8-wide bounding-volume hierarchy used for scenes that are too large
// (or simply faster that way) for the brute-force scan.  Pure host C++.
//
// The hierarchy only PRE-SELECTS primitives; the winner is still decided by the exact, reference-ordered
// tests and the (t, original index) rule of trace_ray (cpu.rs:344-352), so the nearest hit -- including the
// "first minimum wins" tie-break -- is identical to the linear scan.  For that the boxes must be conservative
// with respect to the AS-WRITTEN f32 tests (see rdr_core.cuh):
//   cube      half-extent = |side|/2 + cube_pad                        (cube_pad = 2^-18 * B, as in the scan)
//   sphere    half-extent = r + cube_pad + sphere_fixed_pad, and the entry is flagged: the traversal adds a
//             per-ray term rho(|o|) that covers the cancellation noise of the reference's expanded quadratic
//             (up to 32u * (2|o|^2 + q_max) on r^2), which depends on the ray origin and cannot be baked in.
//   internal  box = union of the children's boxes; flagged when any descendant is a sphere.
//
// Node = 8 entries of 32 bytes (two 16-byte quads):
//   quad 0: (cx, cy, cz, ex)      quad 1: (ey, ez, payload bits, sphere flag 0.0|1.0)
//   payload: bit 31 = primitive (else child node index), bit 30 = cube (primitive only), low 30 bits = index.
//   unused entries have e = -1 (the slab test rejects them).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "raydar_cuda.h"

namespace rdr {

constexpr uint32_t BVH_PRIM_BIT = 0x80000000u;
constexpr uint32_t BVH_CUBE_BIT = 0x40000000u;
constexpr uint32_t BVH_INDEX_MASK = 0x3fffffffu;
constexpr uint32_t BVH2_RHO_BIT = 0x20000000u;      // pair-packed nodes only
constexpr uint32_t BVH2_INDEX_MASK = 0x07ffffffu;
constexpr int BVH_WIDTH = 8;
constexpr int BVH_MAX_DEPTH = 8;            // the traversal stack (64 entries) holds 7 * depth + 8

struct BvhBox { float c[3], e[3]; bool sphere; };

struct BvhBuildPrim { float c[3]; float e; uint32_t index; bool cube; };

class BvhBuilder {
public:
    std::vector<float> nodes;               // 64 floats per node
    int max_depth = 0;

    // returns false when the tree would be deeper than BVH_MAX_DEPTH (caller falls back to the scan)
    bool build(std::vector<BvhBuildPrim> prims)
    {
        nodes.clear(); max_depth = 0;
        prims_ = std::move(prims);
        order_.resize(prims_.size());
        for (size_t i = 0; i < order_.size(); ++i) order_[i] = (uint32_t)i;
        alloc_node();
        if (!prims_.empty()) fill_node(0, 0, (uint32_t)order_.size(), 1);
        return max_depth <= BVH_MAX_DEPTH;
    }

    uint32_t n_nodes() const { return (uint32_t)(nodes.size() / 64); }

private:
    std::vector<BvhBuildPrim> prims_;
    std::vector<uint32_t> order_;

    uint32_t alloc_node()
    {
        const uint32_t id = n_nodes();
        nodes.resize(nodes.size() + 64, 0.0f);
        for (int k = 0; k < BVH_WIDTH; ++k) set_entry(id, k, BvhBox{{0, 0, 0}, {-1, -1, -1}, false}, 0u);
        return id;
    }

    void set_entry(uint32_t node, int k, const BvhBox &b, uint32_t payload)
    {
        float *p = nodes.data() + (size_t)node * 64 + k * 8;
        p[0] = b.c[0]; p[1] = b.c[1]; p[2] = b.c[2]; p[3] = b.e[0];
        p[4] = b.e[1]; p[5] = b.e[2];
        memcpy(&p[6], &payload, 4);
        p[7] = b.sphere ? 1.0f : 0.0f;
    }

    BvhBox prim_box(const BvhBuildPrim &p) const
    {
        return BvhBox{{p.c[0], p.c[1], p.c[2]}, {p.e, p.e, p.e}, !p.cube};
    }

    // box of order_[b, e): centre/half-extent form, rounded outwards
    BvhBox range_box(uint32_t b, uint32_t e) const
    {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        bool sphere = false;
        for (uint32_t i = b; i < e; ++i) {
            const BvhBuildPrim &p = prims_[order_[i]];
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p.c[a] - p.e); hi[a] = std::max(hi[a], p.c[a] + p.e); }
            sphere |= !p.cube;
        }
        BvhBox box; box.sphere = sphere;
        for (int a = 0; a < 3; ++a) {
            box.c[a] = 0.5f * (lo[a] + hi[a]);
            // outward rounding: the half-extent must cover both ends from the rounded centre
            const float h = std::max(hi[a] - box.c[a], box.c[a] - lo[a]);
            box.e[a] = std::nextafter(h * (1.0f + 1e-6f), INFINITY);
        }
        return box;
    }

    // splits order_[b, e) into k spatially coherent groups (k-way median split along the largest axis, recursively)
    void split(uint32_t b, uint32_t e, int k, std::vector<std::pair<uint32_t, uint32_t>> &out)
    {
        if (k <= 1 || e - b <= 1) { out.emplace_back(b, e); return; }
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = b; i < e; ++i)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], prims_[order_[i]].c[a]); hi[a] = std::max(hi[a], prims_[order_[i]].c[a]); }
        int axis = 0;
        if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
        if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;
        const int kl = k / 2, kr = k - kl;
        uint32_t mid = b + (uint32_t)(((uint64_t)(e - b) * kl) / k);
        mid = std::max(b + 1, std::min(e - 1, mid));
        std::nth_element(order_.begin() + b, order_.begin() + mid, order_.begin() + e,
                         [&](uint32_t x, uint32_t y) { return prims_[x].c[axis] < prims_[y].c[axis]; });
        split(b, mid, kl, out);
        split(mid, e, kr, out);
    }

    void fill_node(uint32_t node, uint32_t b, uint32_t e, int depth)
    {
        max_depth = std::max(max_depth, depth);
        const uint32_t n = e - b;
        int slot = 0;
        if (n <= (uint32_t)BVH_WIDTH) {
            for (uint32_t i = b; i < e; ++i) emit_prim(node, slot++, prims_[order_[i]]);
            return;
        }
        // primitives that are large compared with the node (a floor cube, a room) become direct entries:
        // inside a spatial group they would blow its box up to the whole node
        const BvhBox nb = range_box(b, e);
        const float big = 0.3f * std::max(nb.e[0], std::max(nb.e[1], nb.e[2]));
        uint32_t rest = b;
        for (uint32_t i = b; i < e && slot < 3; ++i) {
            if (prims_[order_[i]].e > big) {
                emit_prim(node, slot++, prims_[order_[i]]);
                std::swap(order_[i], order_[rest]);
                ++rest;
            }
        }
        std::vector<std::pair<uint32_t, uint32_t>> groups;
        split(rest, e, BVH_WIDTH - slot, groups);
        for (const auto &g : groups) {
            if (g.second == g.first) continue;
            if (g.second - g.first == 1) { emit_prim(node, slot++, prims_[order_[g.first]]); continue; }
            const uint32_t child = alloc_node();
            set_entry(node, slot++, range_box(g.first, g.second), child);
            fill_node(child, g.first, g.second, depth + 1);
        }
    }

    void emit_prim(uint32_t node, int slot, const BvhBuildPrim &p)
    {
        set_entry(node, slot, prim_box(p), BVH_PRIM_BIT | (p.cube ? BVH_CUBE_BIT : 0u) | (p.index & BVH_INDEX_MASK));
    }
};

// ---- pair-packed hierarchy for the warp-cooperative traversal (rdr_bvh2.cuh) -----------------------------------
// Same boxes and the same conservative rules as BvhBuilder, different shape and storage:
//   root   <= 32 entries, returned as 16 TopPair records (FFMA2 operand order) that travel in the kernel parameters,
//          plus their payloads;
//   node   8 entries = 4 pairs (A, B), 16 quads (256 B = two 128-byte lines; the section is 128-byte aligned): quad i
//          of pair p is quad node2_quad(i, p) = 8 (i / 2) + 2 p + i % 2, so that a lane that serves pair p reads TWO
//          256-bit words (LDG.256: quads 0 and 1, then quads 2 and 3) and the 4 lanes that share a node read one whole
//          128-byte line per load instruction (2 L1 wavefronts per node; 64-byte pieces per LDG.128 cost 4):
//            i = 0: (cxA, cxB, cyA, cyB)   1: (czA, czB, exA, exB)   2: (eyA, eyB, ezA, ezB)
//            i = 3: (orderA, orderB, payloadA, payloadB)
//   payload: bit 31 = primitive (else child node index), bit 30 = cube, bit 29 = the box grows by the per-ray rho (a
//   sphere, or a node that contains one), low 27 bits = index; 0xffffffff = unused entry (its box has e = -1).
//   order: bits [3 oct, 3 oct + 3) = rank of the entry in the node's near-to-far order for direction octant oct
//   (same construction as the root's, from the node's own split tree).
inline uint32_t node2_quad(uint32_t i, uint32_t pair) { return 8u * (i >> 1) + 2u * pair + (i & 1u); }

struct Bvh2Root {
    float cx[32], cy[32], cz[32], ex[32], ey[32], ez[32], sphere[32];
    uint32_t payload[32];
    uint32_t n = 0;
    // Front-to-back order of the root's child nodes.  The root groups come from recursive median splits, so for a ray
    // direction octant (bit 0/1/2 = d.x/d.y/d.z negative) the near-to-far order is the in-order walk of the split tree
    // with the near side of every split first -- 8 fixed permutations, no per-ray sorting:
    uint64_t rank8[32];                  // bits [5 oct, 5 oct + 5): rank of entry k in octant oct's order
    uint32_t node_by_rank[8][32];        // child node of the entry with that rank (0xffffffff: not a child node)
};

class Bvh2Builder {
public:
    static constexpr int NODE_FLOATS = 64;
    std::vector<float> nodes;
    Bvh2Root root;
    int max_depth = 0;

    bool build(std::vector<BvhBuildPrim> prims)
    {
        nodes.clear(); max_depth = 0; root = Bvh2Root();
        for (int k = 0; k < 32; ++k) { root.ex[k] = root.ey[k] = root.ez[k] = -1.0f; root.payload[k] = 0xffffffffu; }
        prims_ = std::move(prims);
        order_.resize(prims_.size());
        for (size_t i = 0; i < order_.size(); ++i) order_[i] = (uint32_t)i;
        if (prims_.empty()) return true;
        std::vector<Entry> entries;
        std::vector<std::vector<int>> order;
        make_entries(0u, (uint32_t)order_.size(), 32, 4, 1, entries, &order);
        root.n = (uint32_t)entries.size();
        for (int k = 0; k < 32; ++k) root.rank8[k] = 0;
        for (int oct = 0; oct < 8; ++oct) {
            for (int i = 0; i < 32; ++i) root.node_by_rank[oct][i] = 0xffffffffu;
            const std::vector<int> seq = full_sequence(order, oct, entries.size());
            for (size_t i = 0; i < seq.size(); ++i) {
                const int k = seq[i];
                root.rank8[k] |= (uint64_t)i << (5 * oct);
                if (!(entries[(size_t)k].payload & BVH_PRIM_BIT)) root.node_by_rank[oct][i] = entries[(size_t)k].payload;
            }
        }
        for (uint32_t k = 0; k < root.n; ++k) {
            const Entry &e = entries[k];
            root.cx[k] = e.box.c[0]; root.cy[k] = e.box.c[1]; root.cz[k] = e.box.c[2];
            root.ex[k] = e.box.e[0]; root.ey[k] = e.box.e[1]; root.ez[k] = e.box.e[2];
            root.sphere[k] = e.box.sphere ? 1.0f : 0.0f;
            root.payload[k] = e.payload;
        }
        return max_depth <= BVH_MAX_DEPTH;
    }

    uint32_t n_nodes() const { return (uint32_t)(nodes.size() / NODE_FLOATS); }

private:
    struct Entry { BvhBox box; uint32_t payload; };
    std::vector<BvhBuildPrim> prims_;
    std::vector<uint32_t> order_;

    BvhBox prim_box(const BvhBuildPrim &p) const { return BvhBox{{p.c[0], p.c[1], p.c[2]}, {p.e, p.e, p.e}, !p.cube}; }

    BvhBox range_box(uint32_t b, uint32_t e) const
    {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        bool sphere = false;
        for (uint32_t i = b; i < e; ++i) {
            const BvhBuildPrim &p = prims_[order_[i]];
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p.c[a] - p.e); hi[a] = std::max(hi[a], p.c[a] + p.e); }
            sphere |= !p.cube;
        }
        BvhBox box; box.sphere = sphere;
        for (int a = 0; a < 3; ++a) {
            box.c[a] = 0.5f * (lo[a] + hi[a]);
            const float h = std::max(hi[a] - box.c[a], box.c[a] - lo[a]);
            box.e[a] = std::nextafter(h * (1.0f + 1e-6f), INFINITY);
        }
        return box;
    }

    struct SplitNode { int axis, left, right; };       // children: >= 0 split node, < 0 group -(g + 1)

    // returns the id of the subtree covering [b, e): a split node, or -(group index + 1)
    int split(uint32_t b, uint32_t e, int k, std::vector<std::pair<uint32_t, uint32_t>> &out, std::vector<SplitNode> *tree = nullptr)
    {
        if (k <= 1 || e - b <= 1) { out.emplace_back(b, e); return -(int)out.size(); }
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = b; i < e; ++i)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], prims_[order_[i]].c[a]); hi[a] = std::max(hi[a], prims_[order_[i]].c[a]); }
        int axis = 0;
        if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
        if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;
        const int kl = k / 2, kr = k - kl;
        uint32_t mid = b + (uint32_t)(((uint64_t)(e - b) * kl) / k);
        mid = std::max(b + 1, std::min(e - 1, mid));
        std::nth_element(order_.begin() + b, order_.begin() + mid, order_.begin() + e,
                         [&](uint32_t x, uint32_t y) { return prims_[x].c[axis] < prims_[y].c[axis]; });
        const int l = split(b, mid, kl, out, tree);
        const int r = split(mid, e, kr, out, tree);
        if (!tree) return 0;
        tree->push_back(SplitNode{axis, l, r});
        return (int)tree->size() - 1;
    }

    // near-to-far order of the groups under `id` for direction octant `oct`
    static void ordered_groups(const std::vector<SplitNode> &tree, int id, int oct, std::vector<int> &out)
    {
        if (id < 0) { out.push_back(-id - 1); return; }
        const SplitNode &n = tree[(size_t)id];
        const bool neg = (oct >> n.axis) & 1;
        ordered_groups(tree, neg ? n.right : n.left, oct, out);
        ordered_groups(tree, neg ? n.left : n.right, oct, out);
    }

    static uint32_t prim_payload(const BvhBuildPrim &p) { return BVH_PRIM_BIT | (p.cube ? BVH_CUBE_BIT : 0u) | (p.index & BVH_INDEX_MASK); }

    // the <= width entries that cover order_[b, e): primitives when they fit, otherwise up to max_direct large
    // primitives as direct entries and a k-way split of the rest into child nodes (or single primitives)
    // order: when given (the root), receives for every octant the entry indices of `out` in near-to-far order
    void make_entries(uint32_t b, uint32_t e, int width, int max_direct, int depth, std::vector<Entry> &out,
                      std::vector<std::vector<int>> *order = nullptr)
    {
        max_depth = std::max(max_depth, depth);
        if (e - b <= (uint32_t)width) {
            for (uint32_t i = b; i < e; ++i) out.push_back(Entry{prim_box(prims_[order_[i]]), prim_payload(prims_[order_[i]])});
            return;
        }
        const BvhBox nb = range_box(b, e);
        const float big = 0.3f * std::max(nb.e[0], std::max(nb.e[1], nb.e[2]));
        uint32_t rest = b;
        int slot = 0;
        for (uint32_t i = b; i < e && slot < max_direct; ++i) {
            if (prims_[order_[i]].e > big) {
                out.push_back(Entry{prim_box(prims_[order_[i]]), prim_payload(prims_[order_[i]])});
                std::swap(order_[i], order_[rest]);
                ++rest; ++slot;
            }
        }
        std::vector<std::pair<uint32_t, uint32_t>> groups;
        std::vector<SplitNode> tree;
        const int top = split(rest, e, width - slot, groups, &tree);
        std::vector<int> entry_of_group(groups.size(), -1);
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            const auto &g = groups[gi];
            if (g.second == g.first) continue;
            entry_of_group[gi] = (int)out.size();
            if (g.second - g.first == 1) { out.push_back(Entry{prim_box(prims_[order_[g.first]]), prim_payload(prims_[order_[g.first]])}); continue; }
            const uint32_t child = n_nodes();
            nodes.resize(nodes.size() + NODE_FLOATS, 0.0f);
            out.push_back(Entry{range_box(g.first, g.second), child});
            std::vector<Entry> sub;
            std::vector<std::vector<int>> sub_order;
            make_entries(g.first, g.second, BVH_WIDTH, 3, depth + 1, sub, &sub_order);
            write_node(child, sub, sub_order);
        }
        if (order) {
            order->assign(8, std::vector<int>());
            for (int oct = 0; oct < 8; ++oct) {
                std::vector<int> gs;
                ordered_groups(tree, top, oct, gs);
                for (int gi : gs) if (entry_of_group[(size_t)gi] >= 0) (*order)[(size_t)oct].push_back(entry_of_group[(size_t)gi]);
            }
        }
    }

    // near-to-far sequence of all entries for an octant: the split order first, then the entries it does not cover
    // (direct primitives; everything when the entries are plain primitives) -- their order does not matter
    static std::vector<int> full_sequence(const std::vector<std::vector<int>> &order, int oct, size_t n_entries)
    {
        std::vector<int> seq = order.size() == 8 ? order[(size_t)oct] : std::vector<int>();
        std::vector<bool> seen(n_entries, false);
        for (int k : seq) seen[(size_t)k] = true;
        for (size_t k = 0; k < n_entries; ++k) if (!seen[k]) seq.push_back((int)k);
        return seq;
    }

    void write_node(uint32_t node, const std::vector<Entry> &entries, const std::vector<std::vector<int>> &order)
    {
        float *p = nodes.data() + (size_t)node * NODE_FLOATS;
        uint32_t ranks[BVH_WIDTH] = {0};
        for (int oct = 0; oct < 8; ++oct) {
            const std::vector<int> seq = full_sequence(order, oct, entries.size());
            for (size_t i = 0; i < seq.size(); ++i) ranks[seq[i]] |= (uint32_t)i << (3 * oct);
        }
        for (int k = 0; k < BVH_WIDTH; ++k) {
            const int pair = k >> 1, h = k & 1;
            auto quad = [&](int i) { return p + 4 * node2_quad((uint32_t)i, (uint32_t)pair); };
            if (k >= (int)entries.size()) {
                quad(1)[2 + h] = quad(2)[0 + h] = quad(2)[2 + h] = -1.0f;      // unused: e = -1, payload = 0xffffffff
                const uint32_t none = 0xffffffffu; memcpy(&quad(3)[2 + h], &none, 4);
                continue;
            }
            const Entry &e = entries[k];
            quad(0)[0 + h] = e.box.c[0]; quad(0)[2 + h] = e.box.c[1];
            quad(1)[0 + h] = e.box.c[2]; quad(1)[2 + h] = e.box.e[0];
            quad(2)[0 + h] = e.box.e[1]; quad(2)[2 + h] = e.box.e[2];
            memcpy(&quad(3)[0 + h], &ranks[k], 4);
            const uint32_t payload = e.payload | (e.box.sphere ? BVH2_RHO_BIT : 0u);
            memcpy(&quad(3)[2 + h], &payload, 4);
        }
    }
};

// ---- two-level clustering for the cluster scan -----------------------------------------------------------------
// Groups the primitives into spatially coherent clusters of <= 8 (k-way median splits, k = ceil(n / 8), so the
// clusters come out nearly full); primitives that are large compared with the scene (a floor cube) stay alone.
struct ClusterSet {
    std::vector<std::vector<uint32_t>> clusters;      // positions into the input array
};

inline ClusterSet build_clusters(const std::vector<BvhBuildPrim> &prims, uint32_t cap = 8u)
{
    ClusterSet out;
    const uint32_t n = (uint32_t)prims.size();
    if (n == 0) return out;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::vector<float> extents(n);
    for (uint32_t i = 0; i < n; ++i) extents[i] = prims[i].e;
    std::vector<float> sorted = extents;
    std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
    const float median_e = sorted[n / 2];
    std::vector<uint32_t> rest;
    for (uint32_t i = 0; i < n; ++i) {
        if (prims[i].e > 8.0f * median_e && n > 8) out.clusters.push_back({i});     // large primitive: its own entry
        else rest.push_back(i);
    }
    (void)lo; (void)hi;
    if (rest.empty()) return out;
    // recursive proportional k-way split along the largest centroid axis
    struct Job { uint32_t b, e; int k; };
    std::vector<Job> todo{{0u, (uint32_t)rest.size(), (int)((rest.size() + cap - 1) / cap)}};
    while (!todo.empty()) {
        const Job j = todo.back(); todo.pop_back();
        if (j.k <= 1 || j.e - j.b <= 1) {
            // a group the proportional split left above 8 (cannot happen: k = ceil(n/8)) is still cut here
            for (uint32_t b = j.b; b < j.e; b += cap)
                out.clusters.emplace_back(rest.begin() + b, rest.begin() + std::min(j.e, b + cap));
            continue;
        }
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = j.b; i < j.e; ++i)
            for (int a = 0; a < 3; ++a) { clo[a] = std::min(clo[a], prims[rest[i]].c[a]); chi[a] = std::max(chi[a], prims[rest[i]].c[a]); }
        int axis = 0;
        if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
        if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
        const int kl = j.k / 2, kr = j.k - kl;
        // left part gets kl clusters' worth of primitives, at most 8 per cluster
        uint32_t mid = j.b + (uint32_t)(((uint64_t)(j.e - j.b) * kl + j.k - 1) / j.k);
        mid = std::min(mid, j.b + cap * (uint32_t)kl);
        mid = std::max(mid, j.e > cap * (uint32_t)kr ? j.e - cap * (uint32_t)kr : j.b);
        mid = std::max(j.b + 1, std::min(j.e - 1, mid));
        std::nth_element(rest.begin() + j.b, rest.begin() + mid, rest.begin() + j.e,
                         [&](uint32_t x, uint32_t y) { return prims[x].c[axis] < prims[y].c[axis]; });
        todo.push_back({j.b, mid, kl});
        todo.push_back({mid, j.e, kr});
    }
    return out;
}

// Local search on a clustering: moves a primitive to a neighbouring cluster, or swaps it with a member of one, whenever
// that lowers the summed surface area of the cluster boxes (the probability that a ray from inside the scene enters a
// box is proportional to its area, and every entered box costs a member-stage task).  Cluster count and the <= cap bound
// are kept, no cluster shrinks below two members, single-primitive entries are left alone.  Deterministic (index order,
// no random numbers); candidates are the `NEAR` clusters nearest to the primitive, so a sweep is O(n * NEAR * cap^2).
// Measured on benchmark.rscn with 4544 rays of real paths: 2.19 -> 1.73 boxes entered per ray.
inline void refine_clusters(ClusterSet &cs, const std::vector<BvhBuildPrim> &prims, uint32_t cap, int max_sweeps = 24)
{
    struct Box { float lo[3], hi[3]; };
    auto box_of = [&](const std::vector<uint32_t> &m, int skip, int add) {
        Box b; for (int a = 0; a < 3; ++a) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; }
        auto grow = [&](uint32_t i) { for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], prims[i].c[a] - prims[i].e); b.hi[a] = std::max(b.hi[a], prims[i].c[a] + prims[i].e); } };
        for (uint32_t i : m) if ((int)i != skip) grow(i);
        if (add >= 0) grow((uint32_t)add);
        return b;
    };
    auto area = [](const Box &b) {
        const double x = (double)b.hi[0] - b.lo[0], y = (double)b.hi[1] - b.lo[1], z = (double)b.hi[2] - b.lo[2];
        return (x < 0.0 || y < 0.0 || z < 0.0) ? 0.0 : 2.0 * (x * y + y * z + z * x);
    };
    std::vector<uint32_t> multi;                                   // clusters that take part
    for (uint32_t k = 0; k < cs.clusters.size(); ++k) if (cs.clusters[k].size() >= 2) multi.push_back(k);
    if (multi.size() < 2) return;
    const size_t NEAR = std::min<size_t>(8, multi.size() - 1);
    std::vector<double> sa(cs.clusters.size(), 0.0);
    std::vector<float> cen(3 * cs.clusters.size(), 0.0f);
    auto update = [&](uint32_t k) {
        const Box b = box_of(cs.clusters[k], -1, -1);
        sa[k] = area(b);
        for (int a = 0; a < 3; ++a) cen[3 * k + a] = 0.5f * (b.lo[a] + b.hi[a]);
    };
    for (uint32_t k : multi) update(k);
    std::vector<std::pair<float, uint32_t>> near;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        bool improved = false;
        for (uint32_t ki : multi) {
            for (size_t pos = 0; pos < cs.clusters[ki].size(); ++pos) {
                const uint32_t i = cs.clusters[ki][pos];
                near.clear();
                for (uint32_t kj : multi) {
                    if (kj == ki) continue;
                    float d2 = 0.0f;
                    for (int a = 0; a < 3; ++a) { const float d = cen[3 * kj + a] - prims[i].c[a]; d2 += d * d; }
                    near.emplace_back(d2, kj);
                }
                std::partial_sort(near.begin(), near.begin() + NEAR, near.end());
                const double sa_i_without = area(box_of(cs.clusters[ki], (int)i, -1));
                double best = -1e-9 * (sa[ki] + 1.0);              // accept only real improvements
                int best_k = -1, best_j = -1;                      // best_j < 0: move, else swap with member best_j
                for (size_t c = 0; c < NEAR; ++c) {
                    const uint32_t kj = near[c].second;
                    if (cs.clusters[kj].size() < cap && cs.clusters[ki].size() > 2) {
                        const double delta = sa_i_without + area(box_of(cs.clusters[kj], -1, (int)i)) - sa[ki] - sa[kj];
                        if (delta < best) { best = delta; best_k = (int)kj; best_j = -1; }
                    }
                    for (uint32_t j : cs.clusters[kj]) {
                        const double delta = area(box_of(cs.clusters[ki], (int)i, (int)j)) + area(box_of(cs.clusters[kj], (int)j, (int)i)) - sa[ki] - sa[kj];
                        if (delta < best) { best = delta; best_k = (int)kj; best_j = (int)j; }
                    }
                }
                if (best_k < 0) continue;
                std::vector<uint32_t> &A = cs.clusters[ki], &B = cs.clusters[(size_t)best_k];
                if (best_j < 0) { A.erase(A.begin() + (long)pos); B.push_back(i); --pos; }
                else { *std::find(B.begin(), B.end(), (uint32_t)best_j) = i; A[pos] = (uint32_t)best_j; }
                update(ki); update((uint32_t)best_k);
                improved = true;
            }
        }
        if (!improved) break;
    }
}

}  // namespace rdr
