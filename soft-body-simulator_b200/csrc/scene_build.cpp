#include "scene_build.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <numeric>

namespace sbsb200 {

// ---------------------------------------------------------------------------------------------
// boundary surface
// ---------------------------------------------------------------------------------------------
namespace {

struct FaceRecord
{
    uint32_t lo, mid, hi; // sorted key (triangle_t::operator== sorts, topology.cpp:234-243)
    uint32_t winding[3];  // winding of the first tet that introduced the face
    uint32_t incident;    // number of incident tets
};

inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

} // namespace

void extract_boundary(int64_t n_vertices, int64_t n_tets, uint32_t const* tets,
                      std::vector<uint32_t>& surf_to_tet, std::vector<uint32_t>* triangles)
{
    surf_to_tet.clear();
    if (triangles)
        triangles->clear();
    // Faces are discovered tet by tet in the order (v1,v2,v4) (v2,v3,v4) (v3,v1,v4) (v1,v3,v2)
    // (tetrahedron_t::faces_copy, topology.cpp:335-342); a face's index is its discovery rank
    // (triangle_set_t::add_triangle, topology.cpp:739-745).
    static constexpr int kFace[4][3] = {{0, 1, 3}, {1, 2, 3}, {2, 0, 3}, {0, 2, 1}};

    std::vector<FaceRecord> faces; // in discovery order
    faces.reserve(static_cast<size_t>(n_tets) * 2 + 4);
    size_t cap = 64;
    while (cap < static_cast<size_t>(n_tets) * 6 + 64)
        cap <<= 1;
    std::vector<uint32_t> slots(cap, 0xffffffffu); // open addressing -> index into faces

    for (int64_t t = 0; t < n_tets; ++t)
    {
        uint32_t const* v = tets + 4 * t;
        for (auto const& f : kFace)
        {
            uint32_t w[3] = {v[f[0]], v[f[1]], v[f[2]]};
            uint32_t k[3] = {w[0], w[1], w[2]};
            std::sort(k, k + 3);
            uint64_t h = mix64((static_cast<uint64_t>(k[0]) << 32) ^ k[1]);
            h          = mix64(h ^ (static_cast<uint64_t>(k[2]) * 0x9e3779b97f4a7c15ull));
            size_t s   = static_cast<size_t>(h) & (cap - 1);
            for (;;)
            {
                uint32_t const fi = slots[s];
                if (fi == 0xffffffffu)
                {
                    slots[s] = static_cast<uint32_t>(faces.size());
                    faces.push_back({k[0], k[1], k[2], {w[0], w[1], w[2]}, 1u});
                    break;
                }
                FaceRecord& r = faces[fi];
                if (r.lo == k[0] && r.mid == k[1] && r.hi == k[2])
                {
                    ++r.incident;
                    break;
                }
                s = (s + 1) & (cap - 1);
            }
        }
    }

    std::vector<uint32_t> tet_to_surf(static_cast<size_t>(std::max<int64_t>(n_vertices, 1)),
                                      0xffffffffu);
    for (FaceRecord const& r : faces)
    {
        if (r.incident == 2u)
            continue; // interior (tetrahedral_mesh_boundary.cpp:88)
        for (uint32_t vi : r.winding)
        {
            if (tet_to_surf[vi] == 0xffffffffu)
            {
                tet_to_surf[vi] = static_cast<uint32_t>(surf_to_tet.size());
                surf_to_tet.push_back(vi);
            }
            if (triangles)
                triangles->push_back(tet_to_surf[vi]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// spatial keys
// ---------------------------------------------------------------------------------------------
namespace {
inline uint64_t spread10(uint32_t v)
{
    uint64_t x = v & 0x3ffu;
    x          = (x | (x << 16)) & 0x30000ffull;
    x          = (x | (x << 8)) & 0x300f00full;
    x          = (x | (x << 4)) & 0x30c30c3ull;
    x          = (x | (x << 2)) & 0x9249249ull;
    return x;
}
} // namespace

void morton_keys(int64_t n, int k, uint32_t const* verts, double const* x0, int64_t n_vertices,
                 std::vector<uint64_t>& keys)
{
    keys.assign(static_cast<size_t>(n), 0);
    if (n == 0 || n_vertices == 0)
        return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < n_vertices; ++i)
        for (int c = 0; c < 3; ++c)
        {
            lo[c] = std::min(lo[c], x0[3 * i + c]);
            hi[c] = std::max(hi[c], x0[3 * i + c]);
        }
    double const ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-30});
    for (int64_t i = 0; i < n; ++i)
    {
        double c[3] = {0, 0, 0};
        for (int a = 0; a < k; ++a)
            for (int d = 0; d < 3; ++d)
                c[d] += x0[3 * static_cast<size_t>(verts[k * i + a]) + d];
        uint32_t q[3];
        for (int d = 0; d < 3; ++d)
        {
            double const u = (c[d] / k - lo[d]) / ext;
            q[d] = static_cast<uint32_t>(std::min(1023.0, std::max(0.0, u * 1024.0)));
        }
        keys[static_cast<size_t>(i)] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    }
}

// ---------------------------------------------------------------------------------------------
// colouring
// ---------------------------------------------------------------------------------------------
bool colour_constraints(int64_t n_vertices, int64_t n, int k, uint32_t const* verts,
                        uint64_t const* spatial_key, int32_t const* region, int max_colours,
                        ColourClass& out)
{
    constexpr int kWords = 4; // up to 256 colours
    if (max_colours > 64 * kWords)
        max_colours = 64 * kWords;
    std::vector<std::array<uint64_t, kWords>> used(static_cast<size_t>(n_vertices),
                                                   std::array<uint64_t, kWords>{});
    std::vector<int32_t> colour(static_cast<size_t>(n), -1);
    int32_t n_colours = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        std::array<uint64_t, kWords> mask{};
        for (int a = 0; a < k; ++a)
        {
            auto const& u = used[verts[k * i + a]];
            for (int w = 0; w < kWords; ++w)
                mask[w] |= u[w];
        }
        int32_t c = -1;
        for (int w = 0; w < kWords && c < 0; ++w)
            if (~mask[w])
                c = 64 * w + __builtin_ctzll(~mask[w]);
        if (c < 0 || c >= max_colours)
            return false;
        colour[static_cast<size_t>(i)] = c;
        n_colours                      = std::max(n_colours, c + 1);
        for (int a = 0; a < k; ++a)
            used[verts[k * i + a]][c >> 6] |= (1ull << (c & 63));
    }
    out.n_colours = n_colours;
    out.order.resize(static_cast<size_t>(n));
    std::iota(out.order.begin(), out.order.end(), 0u);
    std::stable_sort(out.order.begin(), out.order.end(), [&](uint32_t a, uint32_t b) {
        if (colour[a] != colour[b])
            return colour[a] < colour[b];
        if (region && region[a] != region[b])
            return region[a] < region[b];
        if (spatial_key && spatial_key[a] != spatial_key[b])
            return spatial_key[a] < spatial_key[b];
        return false;
    });
    out.offsets.assign(static_cast<size_t>(n_colours) + 1, 0);
    for (int64_t i = 0; i < n; ++i)
        ++out.offsets[static_cast<size_t>(colour[static_cast<size_t>(i)]) + 1];
    for (int32_t c = 0; c < n_colours; ++c)
        out.offsets[static_cast<size_t>(c) + 1] += out.offsets[static_cast<size_t>(c)];
    return true;
}

bool colouring_is_valid(int64_t n_vertices, int k, uint32_t const* verts, ColourClass const& cc)
{
    std::vector<int32_t> stamp(static_cast<size_t>(n_vertices), -1);
    for (int32_t c = 0; c < cc.n_colours; ++c)
        for (int64_t p = cc.offsets[c]; p < cc.offsets[c + 1]; ++p)
        {
            uint32_t const i = cc.order[static_cast<size_t>(p)];
            for (int a = 0; a < k; ++a)
            {
                uint32_t const v = verts[static_cast<size_t>(k) * i + a];
                if (stamp[v] == c)
                    return false;
                stamp[v] = c;
            }
        }
    return true;
}

// ---------------------------------------------------------------------------------------------
// regions
// ---------------------------------------------------------------------------------------------
void plan_regions(HostScene const& scene, std::vector<uint64_t> const& tet_keys, int32_t n_regions,
                  RegionPlan& plan)
{
    int64_t const T = scene.n_tets(), V = scene.n_vertices();
    n_regions       = std::max<int32_t>(1, n_regions);
    plan            = RegionPlan{};
    plan.n_regions  = n_regions;
    plan.tet_region.assign(static_cast<size_t>(T), 0);

    // Bodies are kept together when several whole bodies fit a region (ensembles): the key is
    // (body, morton) so that a region boundary never needlessly cuts through a small body.
    std::vector<uint32_t> by_key(static_cast<size_t>(T));
    std::iota(by_key.begin(), by_key.end(), 0u);
    std::vector<int32_t> tet_body(static_cast<size_t>(T), 0);
    for (size_t b = 0; b < scene.bodies.size(); ++b)
        if (scene.bodies[b].kind == BodyKind::tet)
            for (int64_t t = scene.bodies[b].t_offset; t < scene.bodies[b].t_offset + scene.bodies[b].n_tets; ++t)
                tet_body[static_cast<size_t>(t)] = static_cast<int32_t>(b);
    std::stable_sort(by_key.begin(), by_key.end(), [&](uint32_t a, uint32_t b) {
        if (tet_body[a] != tet_body[b])
            return tet_body[a] < tet_body[b];
        return tet_keys[a] < tet_keys[b];
    });
    for (int64_t p = 0; p < T; ++p)
        plan.tet_region[by_key[static_cast<size_t>(p)]] =
            static_cast<int32_t>((p * static_cast<int64_t>(n_regions)) / std::max<int64_t>(T, 1));

    // vertex ownership: -2 = untouched so far, r >= 0 = only region r so far, -1 = interface
    plan.vertex_region.assign(static_cast<size_t>(V), -2);
    for (int64_t t = 0; t < T; ++t)
        for (int a = 0; a < 4; ++a)
        {
            int32_t& o = plan.vertex_region[scene.tets[4 * static_cast<size_t>(t) + a]];
            int32_t const r = plan.tet_region[static_cast<size_t>(t)];
            if (o == -2)
                o = r;
            else if (o != r)
                o = -1;
        }
    // vertices touched by distance constraints stay global
    for (uint32_t v : scene.dist_pairs)
        plan.vertex_region[v] = -1;

    plan.region_vtx_offsets.assign(static_cast<size_t>(n_regions) + 1, 0);
    plan.vertex_slot.assign(static_cast<size_t>(V), 0);
    for (int64_t v = 0; v < V; ++v)
    {
        int32_t const r = plan.vertex_region[static_cast<size_t>(v)];
        if (r >= 0)
            ++plan.region_vtx_offsets[static_cast<size_t>(r) + 1];
        else
            ++plan.n_interface; // includes vertices no tet touches: they are integrated globally
    }
    for (int32_t r = 0; r < n_regions; ++r)
    {
        plan.max_region_vertices =
            std::max(plan.max_region_vertices, plan.region_vtx_offsets[static_cast<size_t>(r) + 1]);
        plan.region_vtx_offsets[static_cast<size_t>(r) + 1] += plan.region_vtx_offsets[static_cast<size_t>(r)];
    }
    plan.region_vtx.assign(static_cast<size_t>(plan.region_vtx_offsets.back()), 0);
    std::vector<int64_t> cursor(plan.region_vtx_offsets.begin(), plan.region_vtx_offsets.end() - 1);
    for (int64_t v = 0; v < V; ++v)
    {
        int32_t const r = plan.vertex_region[static_cast<size_t>(v)];
        if (r < 0)
        {
            plan.vertex_region[static_cast<size_t>(v)] = -1;
            continue;
        }
        int64_t const pos = cursor[static_cast<size_t>(r)]++;
        plan.region_vtx[static_cast<size_t>(pos)] = static_cast<uint32_t>(v);
        plan.vertex_slot[static_cast<size_t>(v)] =
            static_cast<uint32_t>(pos - plan.region_vtx_offsets[static_cast<size_t>(r)]);
    }

    // neighbour regions: regions sharing an interface vertex
    std::vector<std::pair<uint32_t, int32_t>> vr; // (interface vertex, region)
    for (int64_t t = 0; t < T; ++t)
        for (int a = 0; a < 4; ++a)
        {
            uint32_t const v = scene.tets[4 * static_cast<size_t>(t) + a];
            if (plan.vertex_region[v] == -1)
                vr.emplace_back(v, plan.tet_region[static_cast<size_t>(t)]);
        }
    std::sort(vr.begin(), vr.end());
    vr.erase(std::unique(vr.begin(), vr.end()), vr.end());
    std::vector<std::pair<int32_t, int32_t>> edges;
    for (size_t i = 0; i < vr.size();)
    {
        size_t j = i;
        while (j < vr.size() && vr[j].first == vr[i].first)
            ++j;
        for (size_t a = i; a < j; ++a)
            for (size_t b = i; b < j; ++b)
                if (a != b)
                    edges.emplace_back(vr[a].second, vr[b].second);
        i = j;
    }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    plan.nbr_offsets.assign(static_cast<size_t>(n_regions) + 1, 0);
    for (auto const& e : edges)
        ++plan.nbr_offsets[static_cast<size_t>(e.first) + 1];
    for (int32_t r = 0; r < n_regions; ++r)
        plan.nbr_offsets[static_cast<size_t>(r) + 1] += plan.nbr_offsets[static_cast<size_t>(r)];
    plan.nbr.resize(edges.size());
    for (size_t i = 0; i < edges.size(); ++i)
        plan.nbr[i] = edges[i].second; // edges sorted by (first, second) => grouped by region
}

} // namespace sbsb200
