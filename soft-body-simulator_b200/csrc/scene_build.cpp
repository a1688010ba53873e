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
                  RegionPlan& plan, bool one_region_per_body)
{
    int64_t const T = scene.n_tets();
    std::vector<int32_t> body_region(scene.bodies.size(), -1);
    if (one_region_per_body)
    {
        n_regions = 0;
        for (size_t b = 0; b < scene.bodies.size(); ++b)
            if (scene.bodies[b].kind == BodyKind::tet && scene.bodies[b].n_tets > 0)
                body_region[b] = n_regions++;
    }
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
    {
        uint32_t const t = by_key[static_cast<size_t>(p)];
        plan.tet_region[t] = one_region_per_body
                                 ? body_region[static_cast<size_t>(tet_body[t])]
                                 : static_cast<int32_t>((p * static_cast<int64_t>(n_regions)) / std::max<int64_t>(T, 1));
    }

    classify_regions(scene, plan.tet_region, n_regions, plan);
}

void classify_regions(HostScene const& scene, std::vector<int32_t> const& tet_region_in, int32_t n_regions,
                      RegionPlan& plan, int64_t capacity)
{
    int64_t const T = scene.n_tets(), V = scene.n_vertices();
    std::vector<int32_t> const tet_region_copy = tet_region_in;
    plan                                       = RegionPlan{};
    plan.n_regions                             = n_regions;
    plan.tet_region                            = tet_region_copy;
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

    // owner of every vertex: its region when interior, the lowest sharing region when on an
    // interface, round-robin for vertices no tet touches (they still fall under gravity)
    plan.vertex_owner.assign(static_cast<size_t>(V), -1);
    for (int64_t t = 0; t < T; ++t)
        for (int a = 0; a < 4; ++a)
        {
            int32_t& o      = plan.vertex_owner[scene.tets[4 * static_cast<size_t>(t) + a]];
            int32_t const r = plan.tet_region[static_cast<size_t>(t)];
            o               = o < 0 ? r : std::min(o, r);
        }
    for (int64_t v = 0; v < V; ++v)
        if (plan.vertex_owner[static_cast<size_t>(v)] < 0)
            plan.vertex_owner[static_cast<size_t>(v)] = static_cast<int32_t>(v % n_regions);

    plan.region_vtx_offsets.assign(static_cast<size_t>(n_regions) + 1, 0);
    plan.vertex_slot.assign(static_cast<size_t>(V), 0);
    for (int64_t v = 0; v < V; ++v)
    {
        int32_t& r = plan.vertex_region[static_cast<size_t>(v)];
        if (r >= 0 && capacity >= 0 && plan.region_vtx_offsets[static_cast<size_t>(r) + 1] >= capacity)
            r = -1; // the region's shared memory is full: private to the region, but kept in global memory
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

// ---------------------------------------------------------------------------------------------
// clustered colouring
// ---------------------------------------------------------------------------------------------
namespace {
struct Cluster
{
    int32_t body;
    uint32_t cx, cy, cz; // grid cell
    uint64_t morton;
    uint32_t first, count; // range in the (body, cell)-sorted tet list
    int32_t colour = -1, region = 0;
    int32_t part = 1;      // 0 = fetches vertices from global memory, 1 = all vertices resident
    uint32_t nv = 0;       // unique vertices
    uint32_t verts[4 * kMaxCluster];
};
} // namespace

void build_cluster_plan(HostScene const& scene, int32_t n_regions, bool one_region_per_body, ClusterPlan& out,
                        ResidentParams const* resident, RegionPlan* region_plan)
{
    int64_t const T = scene.n_tets(), V = scene.n_vertices();
    out             = ClusterPlan{};
    out.tet_region.assign(static_cast<size_t>(T), 0);
    if (T == 0)
    {
        out.n_regions = std::max<int32_t>(1, one_region_per_body ? 1 : n_regions);
        out.chunks.clear();
        return;
    }
    constexpr double kTargetTets = 5.0;

    // per tet: body, rest centroid, |rest volume|
    std::vector<int32_t> tet_body(static_cast<size_t>(T), 0);
    for (size_t b = 0; b < scene.bodies.size(); ++b)
        if (scene.bodies[b].kind == BodyKind::tet)
            for (int64_t t = scene.bodies[b].t_offset; t < scene.bodies[b].t_offset + scene.bodies[b].n_tets; ++t)
                tet_body[static_cast<size_t>(t)] = static_cast<int32_t>(b);

    std::vector<uint32_t> sorted(static_cast<size_t>(T));
    std::vector<std::array<uint32_t, 3>> cell(static_cast<size_t>(T));
    for (size_t b = 0; b < scene.bodies.size(); ++b)
    {
        HostBody const& hb = scene.bodies[b];
        if (hb.kind != BodyKind::tet || hb.n_tets == 0)
            continue;
        double lo[3] = {1e300, 1e300, 1e300};
        for (int64_t v = hb.v_offset; v < hb.v_offset + hb.n_vertices; ++v)
            for (int d = 0; d < 3; ++d)
                lo[d] = std::min(lo[d], scene.x0[3 * static_cast<size_t>(v) + d]);
        double vol = 0;
        for (int64_t t = hb.t_offset; t < hb.t_offset + hb.n_tets; ++t)
        {
            uint32_t const* v = &scene.tets[4 * static_cast<size_t>(t)];
            double e[3][3];
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 3; ++d)
                    e[k][d] = scene.x0[3 * static_cast<size_t>(v[k]) + d] - scene.x0[3 * static_cast<size_t>(v[3]) + d];
            vol += std::abs(e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) -
                            e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                            e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0])) / 6.0;
        }
        double h = std::cbrt(kTargetTets * vol / static_cast<double>(hb.n_tets));
        if (!(h > 0) || !std::isfinite(h))
            h = 1.0;
        // snap h so that tiny rounding of the cube root cannot drift the grid across a big lattice
        double const snapped = std::round(h * 1048576.0) / 1048576.0;
        h                    = snapped > 0 ? snapped : h;
        for (int64_t t = hb.t_offset; t < hb.t_offset + hb.n_tets; ++t)
        {
            uint32_t const* v = &scene.tets[4 * static_cast<size_t>(t)];
            for (int d = 0; d < 3; ++d)
            {
                double c = 0;
                for (int k = 0; k < 4; ++k)
                    c += scene.x0[3 * static_cast<size_t>(v[k]) + d];
                double const u = (c * 0.25 - lo[d]) / h;
                cell[static_cast<size_t>(t)][d] =
                    static_cast<uint32_t>(std::min(1048575.0, std::max(0.0, std::floor(u))));
            }
        }
    }
    std::iota(sorted.begin(), sorted.end(), 0u);
    std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) {
        if (tet_body[a] != tet_body[b])
            return tet_body[a] < tet_body[b];
        return cell[a] < cell[b]; // raster order (x, y, z); stable => insertion order inside a cell
    });

    // clusters: runs of equal (body, cell), split at kMaxCluster
    std::vector<Cluster> clusters;
    for (int64_t p = 0; p < T;)
    {
        int64_t q = p;
        uint32_t seen[kMaxClusterVertices];
        uint32_t n_seen = 0;
        while (q < T && q - p < kMaxCluster && tet_body[sorted[q]] == tet_body[sorted[p]] &&
               cell[sorted[q]] == cell[sorted[p]])
        { // at most kMaxClusterVertices distinct vertices per cluster (scratch slots of one thread)
            uint32_t fresh[4];
            uint32_t n_fresh = 0;
            for (int a = 0; a < 4; ++a)
            {
                uint32_t const v = scene.tets[4 * static_cast<size_t>(sorted[q]) + a];
                if (std::find(seen, seen + n_seen, v) == seen + n_seen &&
                    std::find(fresh, fresh + n_fresh, v) == fresh + n_fresh)
                    fresh[n_fresh++] = v;
            }
            if (n_seen + n_fresh > static_cast<uint32_t>(kMaxClusterVertices))
                break;
            for (uint32_t k = 0; k < n_fresh; ++k)
                seen[n_seen++] = fresh[k];
            ++q;
        }
        Cluster c;
        c.body   = tet_body[sorted[p]];
        c.cx     = cell[sorted[p]][0];
        c.cy     = cell[sorted[p]][1];
        c.cz     = cell[sorted[p]][2];
        c.morton = spread10(c.cx) | (spread10(c.cy) << 1) | (spread10(c.cz) << 2) |
                   ((static_cast<uint64_t>(c.cx >> 10) ^ (c.cy >> 10) ^ (c.cz >> 10)) << 30);
        c.first  = static_cast<uint32_t>(p);
        c.count  = static_cast<uint32_t>(q - p);
        clusters.push_back(c);
        p = q;
    }
    out.n_clusters = static_cast<int64_t>(clusters.size());

    // first-fit colouring in (body, raster) order; same-colour clusters must be vertex-disjoint
    {
        constexpr int kWords = 2; // up to 128 cluster colours
        std::vector<std::array<uint64_t, kWords>> used(static_cast<size_t>(V), std::array<uint64_t, kWords>{});
        for (Cluster& c : clusters)
        {
            std::array<uint64_t, kWords> mask{};
            for (uint32_t p = c.first; p < c.first + c.count; ++p)
                for (int a = 0; a < 4; ++a)
                {
                    auto const& u = used[scene.tets[4 * static_cast<size_t>(sorted[p]) + a]];
                    for (int w = 0; w < kWords; ++w)
                        mask[w] |= u[w];
                }
            int32_t col = -1;
            for (int w = 0; w < kWords && col < 0; ++w)
                if (~mask[w])
                    col = 64 * w + __builtin_ctzll(~mask[w]);
            if (col < 0)
                col = 64 * kWords - 1; // cannot happen for sane meshes; validity check will flag it
            c.colour      = col;
            out.n_colours = std::max(out.n_colours, col + 1);
            for (uint32_t p = c.first; p < c.first + c.count; ++p)
                for (int a = 0; a < 4; ++a)
                    used[scene.tets[4 * static_cast<size_t>(sorted[p]) + a]][col >> 6] |= (1ull << (col & 63));
        }
    }

    // regions
    if (one_region_per_body)
    {
        std::vector<int32_t> body_region(scene.bodies.size(), -1);
        int32_t nr = 0;
        for (size_t b = 0; b < scene.bodies.size(); ++b)
            if (scene.bodies[b].kind == BodyKind::tet && scene.bodies[b].n_tets > 0)
                body_region[b] = nr++;
        out.n_regions = std::max<int32_t>(1, nr);
        for (Cluster& c : clusters)
            c.region = body_region[static_cast<size_t>(c.body)];
    }
    else if (n_regions > 1 && resident && resident->slabs && !clusters.empty())
    { // layers of the cluster grid along its longest axis, consecutive layers per region, at most one region per layer
        uint32_t lo[3] = {~0u, ~0u, ~0u}, hi[3] = {0u, 0u, 0u};
        for (Cluster const& c : clusters)
        {
            uint32_t const g[3] = {c.cx, c.cy, c.cz};
            for (int d = 0; d < 3; ++d)
            {
                lo[d] = std::min(lo[d], g[d]);
                hi[d] = std::max(hi[d], g[d]);
            }
        }
        int axis = 0;
        for (int d = 1; d < 3; ++d)
            if (hi[d] - lo[d] > hi[axis] - lo[axis])
                axis = d;
        int64_t const layers = static_cast<int64_t>(hi[axis] - lo[axis]) + 1;
        out.n_regions        = static_cast<int32_t>(std::min<int64_t>(n_regions, layers));
        for (Cluster& c : clusters)
        {
            uint32_t const g[3] = {c.cx, c.cy, c.cz};
            c.region = static_cast<int32_t>((static_cast<int64_t>(g[axis] - lo[axis]) * out.n_regions) / layers);
        }
    }
    else if (n_regions > 1)
    {
        out.n_regions = n_regions;
        std::vector<uint32_t> idx(clusters.size());
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
            if (clusters[a].body != clusters[b].body)
                return clusters[a].body < clusters[b].body;
            return clusters[a].morton < clusters[b].morton;
        });
        int64_t seen = 0;
        for (uint32_t i : idx)
        { // balanced by tet count: cluster goes to the region its first tet falls into
            clusters[i].region = static_cast<int32_t>(std::min<int64_t>(n_regions - 1, (seen * n_regions) / T));
            seen += clusters[i].count;
        }
    }
    else
        out.n_regions = 1;

    // unique vertices per cluster
    uint32_t max_cluster_vertices = 0;
    for (Cluster& c : clusters)
    {
        c.nv = 0;
        for (uint32_t p = c.first; p < c.first + c.count; ++p)
            for (int a = 0; a < 4; ++a)
            {
                uint32_t const v = scene.tets[4 * static_cast<size_t>(sorted[p]) + a];
                if (std::find(c.verts, c.verts + c.nv, v) == c.verts + c.nv)
                    c.verts[c.nv++] = v;
            }
        max_cluster_vertices = std::max(max_cluster_vertices, c.nv);
    }
    for (Cluster const& c : clusters)
        for (uint32_t m = 0; m < c.count; ++m)
            out.tet_region[sorted[c.first + m]] = c.region;

    // Hand-off: renumber the colours so that consecutive colours hand a shared vertex over INSIDE a region as
    // often as possible.  D[a][b] = shared vertices whose clusters of colours a and b lie in different regions;
    // the order is the path of least total D (exhaustive up to 9 colours, nearest neighbour + 2-opt beyond).
    // On a lattice this is a Gray code of the cell parities: every step then depends on other regions across
    // one face orientation only.  Any order of the colours is a valid Gauss-Seidel order.
    if (resident && resident->handoff && out.n_regions > 1 && out.n_colours > 2)
    {
        int32_t const C = out.n_colours;
        struct Touch3
        {
            uint32_t v;
            int32_t colour, region;
        };
        std::vector<Touch3> touches;
        touches.reserve(clusters.size() * 8);
        for (Cluster const& c : clusters)
            for (uint32_t u = 0; u < c.nv; ++u)
                touches.push_back({c.verts[u], c.colour, c.region});
        std::sort(touches.begin(), touches.end(), [](Touch3 const& a, Touch3 const& b) { return a.v < b.v; });
        std::vector<int64_t> D(static_cast<size_t>(C) * C, 0);
        for (size_t i = 0; i < touches.size();)
        {
            size_t j = i;
            while (j < touches.size() && touches[j].v == touches[i].v)
                ++j;
            for (size_t a = i; a < j; ++a)
                for (size_t b = a + 1; b < j; ++b)
                    if (touches[a].region != touches[b].region)
                    {
                        ++D[static_cast<size_t>(touches[a].colour) * C + touches[b].colour];
                        ++D[static_cast<size_t>(touches[b].colour) * C + touches[a].colour];
                    }
            i = j;
        }
        auto const cost = [&](std::vector<int32_t> const& o) {
            int64_t s = 0;
            for (int32_t i = 0; i + 1 < C; ++i) // open path: the sweep's last colour never hands over to its first
                s += D[static_cast<size_t>(o[i]) * C + o[i + 1]];
            return s;
        };
        std::vector<int32_t> best(static_cast<size_t>(C));
        std::iota(best.begin(), best.end(), 0);
        int64_t best_cost = cost(best);
        if (C <= 9)
        {
            std::vector<int32_t> o(best);
            while (std::next_permutation(o.begin(), o.end()))
            {
                int64_t const s = cost(o);
                if (s < best_cost)
                {
                    best_cost = s;
                    best      = o;
                }
            }
        }
        else
        {
            std::vector<int32_t> o{0};
            std::vector<char> taken(static_cast<size_t>(C), 0);
            taken[0] = 1;
            while (static_cast<int32_t>(o.size()) < C)
            {
                int32_t pick = -1;
                for (int32_t c = 0; c < C; ++c)
                    if (!taken[c] && (pick < 0 || D[static_cast<size_t>(o.back()) * C + c] <
                                                      D[static_cast<size_t>(o.back()) * C + pick]))
                        pick = c;
                taken[pick] = 1;
                o.push_back(pick);
            }
            for (bool improved = true; improved;)
            {
                improved = false;
                for (int32_t i = 1; i + 1 < C; ++i)
                    for (int32_t j = i + 1; j < C; ++j)
                    {
                        std::vector<int32_t> t(o);
                        std::reverse(t.begin() + i, t.begin() + j + 1);
                        if (cost(t) < cost(o))
                        {
                            o        = t;
                            improved = true;
                        }
                    }
            }
            if (cost(o) < best_cost)
                best = o;
        }
        std::vector<int32_t> renumber(static_cast<size_t>(C));
        for (int32_t i = 0; i < C; ++i)
            renumber[static_cast<size_t>(best[i])] = i;
        for (Cluster& c : clusters)
            c.colour = renumber[static_cast<size_t>(c.colour)];
    }

    // resident schedule: vertex classification, cluster parts, launch shape
    if (resident)
    {
        out.nvc = static_cast<int32_t>((max_cluster_vertices + 3) / 4 * 4);
        out.nvc = std::max(out.nvc, 4);
        // threads per CTA: every cluster of a (colour, region) step gets its own thread when possible
        // (cluster i of the step, part A first, runs on thread i % nt)
        auto threads_needed = [&](std::vector<int64_t> const& a_cnt, std::vector<int64_t> const& b_cnt) {
            int64_t need = 0;
            for (size_t i = 0; i < a_cnt.size(); ++i)
                need = std::max(need, a_cnt[i] + b_cnt[i]);
            // more clusters than threads: as many rounds as the largest block needs, of equal width
            int64_t const rounds = std::max<int64_t>(1, (need + resident->max_threads - 1) / resident->max_threads);
            int64_t const width  = (need + rounds - 1) / rounds;
            return static_cast<int32_t>(
                std::min<int64_t>(resident->max_threads, std::max<int64_t>(64, (width + 31) / 32 * 32)));
        };
        size_t const n_steps = static_cast<size_t>(out.n_colours) * static_cast<size_t>(out.n_regions);
        std::vector<int64_t> a_cnt(n_steps, 0), b_cnt(n_steps, 0);
        for (Cluster const& c : clusters)
            ++b_cnt[static_cast<size_t>(c.colour) * out.n_regions + c.region];
        // The classification depends on the shared-memory capacity, which depends on the thread count
        // (scratch slots), which depends on the classification.  Start from a safe upper bound of the
        // thread count, then try once with the count that classification needs; keep it if it holds.
        std::vector<int64_t> total(b_cnt), none(n_steps, 0);
        auto classify_with = [&](int32_t nt) -> int32_t { // returns the thread count this classification needs
            int64_t const scratch  = static_cast<int64_t>(resident->handoff ? 2 : 1) * out.nvc * nt;
            int64_t const capacity = std::min<int64_t>(resident->smem_bytes / resident->vertex_bytes - scratch, 65535 - scratch);
            if (capacity < 0)
            {
                out.why_not = "shared memory cannot hold the per-thread scratch vertices";
                return -1;
            }
            classify_regions(scene, out.tet_region, out.n_regions, *region_plan, capacity);
            std::fill(a_cnt.begin(), a_cnt.end(), 0);
            std::fill(b_cnt.begin(), b_cnt.end(), 0);
            for (Cluster& c : clusters)
            {
                c.part = 1;
                for (uint32_t k = 0; k < c.nv; ++k)
                    if (region_plan->vertex_region[c.verts[k]] != c.region)
                        c.part = 0;
                ++(c.part == 0 ? a_cnt : b_cnt)[static_cast<size_t>(c.colour) * out.n_regions + c.region];
            }
            return threads_needed(a_cnt, b_cnt);
        };
        int32_t const upper = threads_needed(none, total);
        out.nt              = upper;
        if (region_plan)
        {
            int32_t const need = classify_with(upper);
            if (need > 0 && need < upper)
            {
                out.nt = need;
                if (classify_with(need) > need)
                { // more residency shifted clusters from A to B beyond the count: stay with the bound
                    out.nt = upper;
                    classify_with(upper);
                }
            }
        }
        out.rot = resident->rotate_items ? item_rotation(out.nt) : 0;
        {   // hand-off needs every step to be a single round (a thread's scratch slots belong to ONE cluster per step)
            int64_t widest = 0;
            for (size_t i = 0; i < a_cnt.size(); ++i)
                widest = std::max(widest, a_cnt[i] + b_cnt[i]);
            out.banks = resident->handoff && region_plan && widest <= out.nt ? 2 : 1;
        }
        if (out.nvc > kMaxClusterVertices)
            out.why_not = "a cluster has more than 16 distinct vertices";
        if (out.n_colours > 120)
            out.why_not = "more than 120 cluster colours (8-bit step distances)";
    }

    // final order: (colour, region, part, size descending, morton)
    std::vector<uint32_t> idx(clusters.size());
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
        Cluster const &x = clusters[a], &y = clusters[b];
        if (x.colour != y.colour)
            return x.colour < y.colour;
        if (x.region != y.region)
            return x.region < y.region;
        if (x.part != y.part)
            return x.part < y.part;
        if (x.count != y.count)
            return x.count > y.count;
        if (x.body != y.body)
            return x.body < y.body;
        return x.morton < y.morton;
    });

    size_t const n_chunks = static_cast<size_t>(out.n_colours) * static_cast<size_t>(out.n_regions) * 2;
    out.chunks.assign(n_chunks, ChunkDesc{});
    out.storage_order.resize(static_cast<size_t>(T));
    out.serial_order.reserve(static_cast<size_t>(T));
    bool const layout = resident && region_plan && out.why_not.empty();
    // colours that touch each vertex (the touch schedule the tag protocol of the resident kernel follows)
    std::vector<std::array<uint64_t, 2>> vcolours;
    if (layout)
    {
        out.tet_slots.assign(4 * static_cast<size_t>(T), 0);
        out.cl_fetch.assign(static_cast<size_t>(out.nvc) * clusters.size(), 0xffffffffu);
        out.cl_meta.assign(static_cast<size_t>(out.nvc) * clusters.size(), 0u);
        vcolours.assign(static_cast<size_t>(V), std::array<uint64_t, 2>{});
        for (Cluster const& c : clusters)
            for (uint32_t u = 0; u < c.nv; ++u)
                vcolours[c.verts[u]][c.colour >> 6] |= 1ull << (c.colour & 63);
        std::vector<char> is_surface(static_cast<size_t>(V), 0);
        for (HostBody const& hb : scene.bodies)
            if (hb.kind == BodyKind::tet)
                for (uint32_t lv : hb.surf_to_tet)
                    is_surface[static_cast<size_t>(hb.v_offset + lv)] = 1;
        out.vertex_meta.assign(static_cast<size_t>(V), 0);
        for (int64_t v = 0; v < V; ++v)
        {
            auto const& m = vcolours[static_cast<size_t>(v)];
            uint32_t last = 0xffu;
            if (m[1])
                last = 127u - static_cast<uint32_t>(__builtin_clzll(m[1]));
            else if (m[0])
                last = 63u - static_cast<uint32_t>(__builtin_clzll(m[0]));
            out.vertex_meta[static_cast<size_t>(v)] = last | (is_surface[static_cast<size_t>(v)] ? 0x100u : 0u);
        }
    }
    int64_t store = 0;
    size_t i      = 0;
    int64_t pair_clusters = 0;
    for (size_t ch = 0; ch < n_chunks; ++ch)
    {
        int32_t const col  = static_cast<int32_t>(ch / (2 * static_cast<size_t>(out.n_regions)));
        int32_t const reg  = static_cast<int32_t>((ch / 2) % static_cast<size_t>(out.n_regions));
        int32_t const part = static_cast<int32_t>(ch % 2);
        size_t j           = i;
        while (j < idx.size() && clusters[idx[j]].colour == col && clusters[idx[j]].region == reg &&
               clusters[idx[j]].part == part)
            ++j;
        ChunkDesc& d = out.chunks[ch];
        d.first      = static_cast<int32_t>(store);
        d.cfirst     = static_cast<int32_t>(i);
        for (size_t k = i; k < j; ++k)
        {
            Cluster const& c = clusters[idx[k]];
            for (uint32_t m = 0; m < c.count; ++m)
            {
                ++d.n[m];
                out.serial_order.push_back(sorted[c.first + m]);
            }
        }
        int64_t col_base = store;
        for (int m = 0; m < kMaxCluster; ++m)
        {
            for (size_t k = i; k < j; ++k)
            {
                Cluster const& c = clusters[idx[k]];
                if (c.count <= static_cast<uint32_t>(m))
                    continue;
                int64_t const pos = col_base + static_cast<int64_t>(k - i);
                uint32_t const t  = sorted[c.first + m];
                out.storage_order[static_cast<size_t>(pos)] = t;
                if (!layout)
                    continue;
                // vertex addresses in the CTA's shared array: [nvc * nt scratch | resident vertices]
                uint32_t const thread =
                    static_cast<uint32_t>((k - i + static_cast<size_t>(out.rot)) % static_cast<size_t>(out.nt));
                for (int a = 0; a < 4; ++a)
                {
                    uint32_t const v = scene.tets[4 * static_cast<size_t>(t) + a];
                    uint32_t slot;
                    if (region_plan->vertex_region[v] == reg)
                        slot = static_cast<uint32_t>(out.banks * out.nvc) * out.nt + region_plan->vertex_slot[v];
                    else
                    { // k-th fetched vertex of this cluster
                        uint32_t kf = 0;
                        for (uint32_t u = 0; u < c.nv && c.verts[u] != v; ++u)
                            kf += region_plan->vertex_region[c.verts[u]] != reg;
                        slot = kf * static_cast<uint32_t>(out.nt) + thread;
                    }
                    out.tet_slots[4 * static_cast<size_t>(pos) + a] = static_cast<uint16_t>(slot);
                }
            }
            col_base += d.n[m];
        }
        if (layout)
            for (size_t k = i; k < j; ++k)
            {
                Cluster const& c = clusters[idx[k]];
                uint32_t kf      = 0;
                for (uint32_t u = 0; u < c.nv; ++u)
                    if (region_plan->vertex_region[c.verts[u]] != reg)
                    {
                        uint32_t const v = c.verts[u];
                        // How many steps back the previous touch of v lies, for the four kinds of
                        // colour step: byte 2*(k>0) + cs, cs = 1 when a collision step precedes every
                        // sweep; 0xff = the predict step of the launch.  Touches are: predict, the
                        // collision steps if v is a surface vertex, the colours whose clusters contain v.
                        auto const& m    = vcolours[v];
                        int32_t prevc    = -1;
                        for (int32_t pc = col - 1; pc >= 0 && prevc < 0; --pc)
                            if (m[pc >> 6] >> (pc & 63) & 1ull)
                                prevc = pc;
                        int32_t const lastc = static_cast<int32_t>(out.vertex_meta[v] & 0xffu);
                        bool const surf     = (out.vertex_meta[v] & 0x100u) != 0u;
                        uint32_t word       = 0;
                        for (int later = 0; later < 2; ++later)
                            for (int cs = 0; cs < 2; ++cs)
                            {
                                int32_t d;
                                if (prevc >= 0)
                                    d = col - prevc;
                                else if (cs && surf)
                                    d = col + 1;
                                else if (later)
                                    d = out.n_colours + cs + col - lastc;
                                else
                                    d = 0xff;
                                word |= static_cast<uint32_t>(d) << (8 * (2 * later + cs));
                            }
                        out.cl_meta[static_cast<size_t>(kf) * clusters.size() + k] = word;
                        out.cl_fetch[static_cast<size_t>(kf++) * clusters.size() + k] = v;
                    }
            }
        store = col_base;
        pair_clusters = (part == 0 ? 0 : pair_clusters) + static_cast<int64_t>(j - i);
        out.max_chunk_clusters = std::max<int64_t>(out.max_chunk_clusters, pair_clusters);
        i = j;
    }
}

bool build_mailbox_routes(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& plan, int nvc, int world,
                          MailboxRoutes& out)
{
    out = MailboxRoutes{};
    int64_t const V = scene.n_vertices(), Q = cp.n_clusters;
    int32_t const Rn = plan.n_regions;
    if (world < 1 || Rn % world != 0 || nvc < cp.nvc)
    {
        out.why_not = "bad partition (world must divide the region count)";
        return false;
    }
    // owned non-resident vertices, grouped by owner region
    out.ifv_offsets.assign(static_cast<size_t>(Rn) + 1, 0);
    for (int64_t v = 0; v < V; ++v)
        if (plan.vertex_region[static_cast<size_t>(v)] < 0)
            ++out.ifv_offsets[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)]) + 1];
    for (int32_t r = 0; r < Rn; ++r)
        out.ifv_offsets[static_cast<size_t>(r) + 1] += out.ifv_offsets[static_cast<size_t>(r)];
    out.ifv.assign(static_cast<size_t>(out.ifv_offsets.back()), 0);
    out.ifv_meta.assign(out.ifv.size(), 0);
    out.ifv_pos.assign(static_cast<size_t>(V), kRouteNone);
    {
        std::vector<int32_t> cur(out.ifv_offsets.begin(), out.ifv_offsets.end() - 1);
        for (int64_t v = 0; v < V; ++v)
            if (plan.vertex_region[static_cast<size_t>(v)] < 0)
            {
                int32_t const at = cur[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)])]++;
                out.ifv[static_cast<size_t>(at)]      = static_cast<uint32_t>(v);
                out.ifv_meta[static_cast<size_t>(at)] = cp.vertex_meta.empty() ? 0u : cp.vertex_meta[static_cast<size_t>(v)];
                out.ifv_pos[static_cast<size_t>(v)]   = static_cast<uint32_t>(at);
            }
    }
    out.n_entries = static_cast<uint32_t>(static_cast<int64_t>(nvc) * Q);
    if (static_cast<uint64_t>(out.n_entries) + out.ifv.size() >= kRouteIndexMask)
    {
        out.why_not = "too many mailboxes for 28-bit routing words";
        return false;
    }
    std::vector<int32_t> cluster_colour(static_cast<size_t>(Q), 0), cluster_region(static_cast<size_t>(Q), 0),
        cluster_item(static_cast<size_t>(Q), 0); // position in its chunk = in its step for part A (first in a step)
    for (size_t ch = 0; ch < cp.chunks.size(); ++ch)
        for (int32_t i = 0; i < cp.chunks[ch].n[0]; ++i)
        {
            cluster_item[static_cast<size_t>(cp.chunks[ch].cfirst + i)] = i;
            cluster_colour[static_cast<size_t>(cp.chunks[ch].cfirst + i)] =
                static_cast<int32_t>(ch / (2 * static_cast<size_t>(Rn)));
            cluster_region[static_cast<size_t>(cp.chunks[ch].cfirst + i)] =
                static_cast<int32_t>((ch / 2) % static_cast<size_t>(Rn));
        }
    // routing word of a mailbox: its index and the rank whose memory holds it (the rank that reads it)
    auto const entry_route = [&](uint32_t box) {
        return box | (static_cast<uint32_t>(region_rank(cluster_region[box % static_cast<uint32_t>(Q)], Rn, world))
                      << kRouteRankShift);
    };
    auto const owner_route = [&](uint32_t pos) {
        return (out.n_entries + pos) |
               (static_cast<uint32_t>(region_rank(plan.vertex_owner[out.ifv[pos]], Rn, world)) << kRouteRankShift);
    };
    struct Touch
    {
        uint32_t vertex;
        int32_t colour;
        uint32_t box;
    };
    std::vector<Touch> touches;
    for (int j = 0; j < cp.nvc; ++j)
        for (int64_t q = 0; q < Q; ++q)
        {
            uint32_t const v = cp.cl_fetch[static_cast<size_t>(j) * Q + q];
            if (v != kRouteNone)
                touches.push_back({v, cluster_colour[static_cast<size_t>(q)],
                                   static_cast<uint32_t>(static_cast<int64_t>(j) * Q + q)});
        }
    std::sort(touches.begin(), touches.end(), [](Touch const& x, Touch const& y) {
        return x.vertex != y.vertex ? x.vertex < y.vertex : x.colour < y.colour;
    });
    out.to.assign(out.n_entries, kRouteNone);
    out.to_owner.assign(out.n_entries, kRouteNone);
    out.local_prev.assign(out.n_entries, 0);
    out.n_local = 0;
    if (cp.banks == 2 && static_cast<uint64_t>(out.n_entries) + out.ifv.size() >= kRouteLocalBit)
    {
        out.why_not = "too many mailboxes for 27-bit routing words (hand-off)";
        return false;
    }
    // hand-off: the next touch is by a cluster of the same region in the very next step -> the vertex goes
    // straight into that cluster's scratch slot (entry * nt + thread), no mailbox, no poll
    auto const local_route = [&](uint32_t box) {
        uint32_t const q = box % static_cast<uint32_t>(Q), entry = box / static_cast<uint32_t>(Q);
        uint32_t const thread = static_cast<uint32_t>((cluster_item[q] + cp.rot) % cp.nt);
        return kRouteLocalBit | (entry * static_cast<uint32_t>(cp.nt) + thread);
    };
    out.ifv_first.assign(out.ifv.size(), kRouteNone);
    for (size_t i = 0; i < touches.size();)
    {
        size_t j = i;
        while (j < touches.size() && touches[j].vertex == touches[i].vertex)
            ++j;
        uint32_t const v   = touches[i].vertex;
        uint32_t const pos = out.ifv_pos[v];
        if (pos == kRouteNone)
        {
            out.why_not = "a fetched vertex has no owner mailbox";
            return false;
        }
        out.ifv_first[pos] = entry_route(touches[i].box);
        for (size_t t = i; t < j; ++t)
        {
            if (t + 1 < j && touches[t + 1].colour == touches[t].colour)
            {
                out.why_not = "two clusters of one colour touch the same vertex";
                return false;
            }
            uint32_t const surface       = (cp.vertex_meta[v] & 0x100u) ? kRouteSurfaceBit : 0u;
            bool const local = cp.banks == 2 && t + 1 < j && touches[t + 1].colour == touches[t].colour + 1 &&
                               cluster_region[touches[t + 1].box % static_cast<uint32_t>(Q)] ==
                                   cluster_region[touches[t].box % static_cast<uint32_t>(Q)];
            if (local)
            {
                out.to[touches[t].box]             = local_route(touches[t + 1].box);
                out.to_owner[touches[t].box]       = local_route(touches[t + 1].box) | surface;
                out.local_prev[touches[t + 1].box] = 1;
                ++out.n_local;
                continue;
            }
            out.to[touches[t].box]       = entry_route(t + 1 < j ? touches[t + 1].box : touches[i].box);
            out.to_owner[touches[t].box] = (t + 1 < j ? entry_route(touches[t + 1].box) : owner_route(pos)) | surface;
        }
        i = j;
    }
    return true;
}

bool resident_layout_is_valid(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& rp)
{
    int64_t const T = scene.n_tets(), Q = cp.n_clusters;
    if (!cp.why_not.empty() || cp.nt <= 0 || cp.nvc <= 0 || cp.nvc % 4 != 0 || cp.nvc > kMaxClusterVertices ||
        cp.rot < 0 || cp.rot >= cp.nt)
        return false;
    if (static_cast<int64_t>(cp.tet_slots.size()) != 4 * T ||
        static_cast<int64_t>(cp.cl_fetch.size()) != static_cast<int64_t>(cp.nvc) * Q)
        return false;
    if (cp.banks < 1 || cp.banks > 2)
        return false;
    // tet slots name bank 0 of the scratch slots; resident vertices follow the last bank
    uint32_t const scratch = static_cast<uint32_t>(cp.banks * cp.nvc) * static_cast<uint32_t>(cp.nt);
    int64_t clusters_seen  = 0;
    for (size_t ch = 0; ch < cp.chunks.size(); ++ch)
    {
        ChunkDesc const& d = cp.chunks[ch];
        int32_t const reg  = static_cast<int32_t>((ch / 2) % static_cast<size_t>(cp.n_regions));
        int32_t const part = static_cast<int32_t>(ch % 2);
        if (d.cfirst != clusters_seen)
            return false;
        clusters_seen += d.n[0];
        for (int32_t i = 0; i < d.n[0]; ++i)
        {
            int64_t const q      = d.cfirst + i;
            uint32_t const thread = static_cast<uint32_t>((i + cp.rot) % cp.nt); // part A comes first in its step
            bool any_fetch        = false;
            for (int k = 0; k < cp.nvc; ++k)
            {
                uint32_t const v = cp.cl_fetch[static_cast<size_t>(k) * Q + q];
                if (v == 0xffffffffu)
                    continue;
                any_fetch = true;
                if (v >= rp.vertex_region.size() || rp.vertex_region[v] == reg)
                    return false; // resident vertices must not be fetched
            }
            if (any_fetch != (part == 0))
                return false;
            int64_t base = d.first;
            for (int m = 0; m < kMaxCluster && i < d.n[m]; ++m)
            {
                int64_t const pos = base + i;
                uint32_t const t  = cp.storage_order[static_cast<size_t>(pos)];
                for (int a = 0; a < 4; ++a)
                {
                    uint32_t const v    = scene.tets[4 * static_cast<size_t>(t) + a];
                    uint32_t const slot = cp.tet_slots[4 * static_cast<size_t>(pos) + a];
                    if (slot >= scratch)
                    { // resident slot of this region
                        int64_t const at = rp.region_vtx_offsets[static_cast<size_t>(reg)] + (slot - scratch);
                        if (rp.vertex_region[v] != reg || at >= rp.region_vtx_offsets[static_cast<size_t>(reg) + 1] ||
                            rp.region_vtx[static_cast<size_t>(at)] != v)
                            return false;
                    }
                    else
                    { // scratch entry k of the thread that runs this cluster
                        uint32_t const k = slot / static_cast<uint32_t>(cp.nt);
                        if (slot % static_cast<uint32_t>(cp.nt) != thread || k >= static_cast<uint32_t>(cp.nvc) ||
                            cp.cl_fetch[static_cast<size_t>(k) * Q + q] != v)
                            return false;
                    }
                }
                base += d.n[m];
            }
        }
    }
    return clusters_seen == Q;
}

bool cluster_plan_is_valid(HostScene const& scene, ClusterPlan const& plan)
{
    int64_t const T = scene.n_tets(), V = scene.n_vertices();
    if (static_cast<int64_t>(plan.storage_order.size()) != T || static_cast<int64_t>(plan.serial_order.size()) != T)
        return false;
    std::vector<char> seen(static_cast<size_t>(T), 0);
    for (uint32_t t : plan.storage_order)
    {
        if (t >= T || seen[t])
            return false;
        seen[t] = 1;
    }
    // per colour: a vertex may only be touched by ONE cluster (identified by chunk + cluster index)
    std::vector<int64_t> owner(static_cast<size_t>(V), -1);
    std::vector<int32_t> stamp(static_cast<size_t>(V), -1);
    int64_t cluster_id = 0;
    for (int32_t col = 0; col < plan.n_colours; ++col)
        for (int32_t rp = 0; rp < 2 * plan.n_regions; ++rp)
        {
            int32_t const reg  = rp / 2;
            ChunkDesc const& d = plan.chunks[static_cast<size_t>(col) * 2 * plan.n_regions + rp];
            for (int32_t i = 0; i < d.n[0]; ++i, ++cluster_id)
            {
                int64_t base = d.first;
                for (int m = 0; m < kMaxCluster && i < d.n[m]; ++m)
                {
                    uint32_t const t = plan.storage_order[static_cast<size_t>(base + i)];
                    if (plan.tet_region[t] != reg)
                        return false;
                    for (int a = 0; a < 4; ++a)
                    {
                        uint32_t const v = scene.tets[4 * static_cast<size_t>(t) + a];
                        if (stamp[v] == col && owner[v] != cluster_id)
                            return false;
                        stamp[v] = col;
                        owner[v] = cluster_id;
                    }
                    base += d.n[m];
                }
            }
        }
    return true;
}

} // namespace sbsb200
