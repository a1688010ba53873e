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
    int32_t part = 1;      // 0 = has a vertex that another region touches too, 1 = private to its region
    int32_t prio = 2;      // position inside its chunk: 0 = pushes a vertex to another region in its step, 1 = pulls
                           // one that was pushed in the very previous step, 2 = neither
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

    // unique vertices per cluster
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
    }

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

    // Ensembles (whole bodies per region): colour classes of equal size inside every body.  A CTA has a thread — and its
    // registers — for every cluster of its widest colour step, and first-fit on a lattice with odd sides is far from even
    // (5 x 5 x 16 cells: 72, 72, 48, 48, 48, 48, 32, 32), so a third of the register file would sit idle in most steps.
    // Kempe chains: inside a body, the clusters of two colours a, b form connected components (sharing a vertex), and
    // swapping a and b inside one component keeps the colouring valid; components with a surplus of the fuller colour are
    // swapped while that evens the two classes out.  Any valid colouring is a valid Gauss-Seidel order.  (A single body in one
    // region keeps its first-fit colours: its steps are as long as five tets in a row whatever their width.)
#ifdef SBSB200_NO_COLOUR_BALANCE
    if (false)
#else
    if (one_region_per_body && out.n_colours > 1)
#endif
    {
        int32_t const C = out.n_colours;
        std::vector<int64_t> v_off(static_cast<size_t>(V) + 1, 0);
        for (Cluster const& c : clusters)
            for (uint32_t k = 0; k < c.nv; ++k)
                ++v_off[static_cast<size_t>(c.verts[k]) + 1];
        for (int64_t v = 0; v < V; ++v)
            v_off[static_cast<size_t>(v) + 1] += v_off[static_cast<size_t>(v)];
        std::vector<uint32_t> v_cl(static_cast<size_t>(v_off.back()));
        {
            std::vector<int64_t> cur(v_off.begin(), v_off.end() - 1);
            for (size_t ci = 0; ci < clusters.size(); ++ci)
                for (uint32_t k = 0; k < clusters[ci].nv; ++k)
                    v_cl[static_cast<size_t>(cur[clusters[ci].verts[k]]++)] = static_cast<uint32_t>(ci);
        }
        std::vector<uint32_t> stamp(clusters.size(), 0), queue, comp;
        uint32_t epoch = 0;
        std::vector<int32_t> cnt(static_cast<size_t>(C));
        for (size_t b0 = 0; b0 < clusters.size();)
        { // clusters are sorted by body
            size_t b1 = b0;
            while (b1 < clusters.size() && clusters[b1].body == clusters[b0].body)
                ++b1;
            std::fill(cnt.begin(), cnt.end(), 0);
            for (size_t ci = b0; ci < b1; ++ci)
                ++cnt[static_cast<size_t>(clusters[ci].colour)];
            std::vector<int32_t> by_count(static_cast<size_t>(C));
            for (int round = 0; round < 8 * C; ++round)
            { // the fullest colour against the emptiest one first; when their components are too coarse, the next pairs
                std::iota(by_count.begin(), by_count.end(), 0);
                std::stable_sort(by_count.begin(), by_count.end(),
                                 [&](int32_t x, int32_t y) { return cnt[static_cast<size_t>(x)] > cnt[static_cast<size_t>(y)]; });
                bool changed = false;
                for (int32_t ia = 0; ia < C && !changed; ++ia)
                    for (int32_t ib = C - 1; ib > ia && !changed; --ib)
                    {
                        int32_t const a = by_count[static_cast<size_t>(ia)], b = by_count[static_cast<size_t>(ib)];
                        if (cnt[static_cast<size_t>(a)] - cnt[static_cast<size_t>(b)] < 2)
                            break;
                        ++epoch;
                        for (size_t seed = b0; seed < b1; ++seed)
                        {
                            if (stamp[seed] == epoch || (clusters[seed].colour != a && clusters[seed].colour != b))
                                continue;
                            comp.clear();
                            queue.assign(1, static_cast<uint32_t>(seed));
                            stamp[seed] = epoch;
                            int32_t surplus = 0; // clusters of colour a minus clusters of colour b in the component
                            while (!queue.empty())
                            {
                                uint32_t const ci = queue.back();
                                queue.pop_back();
                                comp.push_back(ci);
                                surplus += clusters[ci].colour == a ? 1 : -1;
                                for (uint32_t k = 0; k < clusters[ci].nv; ++k)
                                    for (int64_t e = v_off[clusters[ci].verts[k]];
                                         e < v_off[static_cast<size_t>(clusters[ci].verts[k]) + 1]; ++e)
                                    {
                                        uint32_t const cj = v_cl[static_cast<size_t>(e)];
                                        if (stamp[cj] != epoch && (clusters[cj].colour == a || clusters[cj].colour == b))
                                        {
                                            stamp[cj] = epoch;
                                            queue.push_back(cj);
                                        }
                                    }
                            }
                            int32_t const diff = cnt[static_cast<size_t>(a)] - cnt[static_cast<size_t>(b)];
                            if (surplus > 0 && std::abs(diff - 2 * surplus) < diff)
                            {
                                for (uint32_t ci : comp)
                                    clusters[ci].colour = clusters[ci].colour == a ? b : a;
                                cnt[static_cast<size_t>(a)] -= surplus;
                                cnt[static_cast<size_t>(b)] += surplus;
                                changed = true;
                            }
                        }
                    }
                if (!changed)
                    break;
            }
            b0 = b1;
        }
    }

    // regions
    if (one_region_per_body)
    { // ensembles: consecutive whole bodies per region, enough of them to fill the warps of a colour step
        std::vector<int32_t> body_index(scene.bodies.size(), -1);
        int32_t nb = 0;
        for (size_t b = 0; b < scene.bodies.size(); ++b)
            if (scene.bodies[b].kind == BodyKind::tet && scene.bodies[b].n_tets > 0)
                body_index[b] = nb++;
        int32_t group = 1;
        if (resident)
        {
            group = resident->bodies_per_region;
            if (group <= 0)
            { // about 160 clusters in the widest colour step of a region, and the vertices must fit shared memory
                std::vector<int32_t> per(static_cast<size_t>(std::max(nb, 1)) * static_cast<size_t>(out.n_colours), 0);
                int32_t widest = 1;
                for (Cluster const& c : clusters)
                    widest = std::max(widest, ++per[static_cast<size_t>(body_index[static_cast<size_t>(c.body)]) *
                                                        static_cast<size_t>(out.n_colours) + c.colour]);
                int64_t most_vertices = 1;
                for (HostBody const& hb : scene.bodies)
                    if (hb.kind == BodyKind::tet)
                        most_vertices = std::max<int64_t>(most_vertices, hb.n_vertices);
                int64_t const fit = resident->smem_bytes / (most_vertices * resident->vertex_bytes);
                group = static_cast<int32_t>(std::max<int64_t>(1, std::min<int64_t>((160 + widest / 2) / widest, fit)));
                if (resident->sm_count > 0)
                { // Whole regions are dealt to the resident CTAs in rounds, and the last round is as long as the others
                  // however few CTAs take part: what counts is the capacity an SM has to provide, rounds x CTAs per SM
                  // x bodies per region, against the bodies it really has.  Measured (profiles/r02_ab_bodies_per_region.txt):
                  // the frame time follows that capacity — 4 096 bodies of 50 clusters per step: five per region = 3 rounds
                  // x 2 CTAs x 5 = 30 against 36 with three per region, 17.3 ms against 19.4; 2 048 bodies: seven per region
                  // (one CTA of 352 threads per SM) = 2 x 1 x 7 = 14 against 16 with one per region, 9.0 ms against 10.6 —
                  // times a factor for the shapes that ran slower per body: 7, 10 warps (uneven over the four
                  // sub-partitions) and the 168-register instantiation above 256 threads.  Every SM must get a region;
                  // ties go to the smaller regions.
                    static double const warp_factor[13] = {1, 1, 1, 1, 1, 1, 1, 1.12, 1, 1.1, 1.2, 1.1, 1.1};
                    double best = 1e300;
                    group       = 1;
                    for (int32_t g = 1; g <= fit; ++g)
                    {
                        int64_t const nt = std::max<int64_t>(64, (static_cast<int64_t>(g) * widest + 31) / 32 * 32);
                        if (nt > 384)
                            break;
                        int64_t const ctas    = nt > 256 ? 1 : std::max<int64_t>(1, std::min<int64_t>(65536 / (128 * nt), 2048 / nt));
                        int64_t const regions = (nb + g - 1) / g;
                        if (g > 1 && regions < resident->sm_count)
                            break;
                        int64_t const slots   = ctas * resident->sm_count;
                        double const capacity = static_cast<double>((regions + slots - 1) / slots * ctas * g) * warp_factor[nt / 32];
                        if (capacity < best - 1e-9)
                        {
                            best  = capacity;
                            group = g;
                        }
                    }
                }
            }
        }
        out.n_regions = std::max<int32_t>(1, (nb + group - 1) / group);
        for (Cluster& c : clusters)
            c.region = body_index[static_cast<size_t>(c.body)] / group;
    }
    else if (n_regions > 1)
    {
        out.n_regions = n_regions;
        std::vector<uint32_t> idx(clusters.size());
        std::iota(idx.begin(), idx.end(), 0u);
        // Pencils: whole columns of clusters along the shortest axis of the cluster grid, bundled in the 2-D
        // Morton order of the two other coordinates.  No region boundary is then perpendicular to the pencil
        // axis, and with the colours ordered for it (below) half of the steps of a sweep on a lattice depend on
        // no other region.  Needs a few columns per region; otherwise compact Morton blocks.
        bool pencils = resident && resident->pencils;
        int axis     = 0;
        if (pencils)
        {
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
            for (int d = 1; d < 3; ++d)
                if (hi[d] - lo[d] < hi[axis] - lo[axis])
                    axis = d;
            int const b = (axis + 1) % 3, c = (axis + 2) % 3;
            int64_t const columns = (static_cast<int64_t>(hi[b] - lo[b]) + 1) * (static_cast<int64_t>(hi[c] - lo[c]) + 1);
            pencils               = columns >= 4 * static_cast<int64_t>(n_regions);
        }
        auto const column_key = [&](Cluster const& c) -> uint64_t {
            uint32_t const g[3] = {c.cx, c.cy, c.cz};
            uint32_t const u = g[(axis + 1) % 3], v = g[(axis + 2) % 3];
            // 2-D Morton code of (u, v): the bits of u on the even positions of a 3-D spread, v on the odd ones
            uint64_t key = 0;
            for (int bit = 0; bit < 20; ++bit)
                key |= (static_cast<uint64_t>(u >> bit & 1u) << (2 * bit)) | (static_cast<uint64_t>(v >> bit & 1u) << (2 * bit + 1));
            return key;
        };
        std::vector<uint64_t> key(clusters.size());
        for (size_t i = 0; i < clusters.size(); ++i)
            key[i] = pencils ? column_key(clusters[i]) : clusters[i].morton;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
            if (clusters[a].body != clusters[b].body)
                return clusters[a].body < clusters[b].body;
            return key[a] < key[b];
        });
        if (!pencils)
        { // Compact regions by recursive coordinate bisection: the set is cut across its longest axis into two parts
          // whose tet counts are in the ratio of the region counts they receive, recursively.  Boxes have fewer
          // shared vertices than ranges of a space-filling curve, and consecutive region numbers stay spatially
          // compact (a rank of a decomposed scene takes a block of them).  The unit that is dealt out is a block of
          // 2 x 2 x 2 cells of the cluster grid: on a lattice it holds one cluster of every colour, so every region
          // has the same number of clusters in every colour step (a box with odd sides would have 6 x 6 x 6 clusters
          // of one colour and 5 x 5 x 5 of another).
            struct Block
            {
                int32_t body;
                uint32_t bx, by, bz;
                int64_t weight = 0;
                double c[3]    = {0, 0, 0};
                int32_t region = 0;
            };
            std::vector<uint32_t> by_block(clusters.size());
            std::iota(by_block.begin(), by_block.end(), 0u);
            auto const block_less = [&](uint32_t x, uint32_t y) {
                Cluster const &p = clusters[x], &q = clusters[y];
                if (p.body != q.body)
                    return p.body < q.body;
                if (p.cx / 2 != q.cx / 2)
                    return p.cx / 2 < q.cx / 2;
                if (p.cy / 2 != q.cy / 2)
                    return p.cy / 2 < q.cy / 2;
                return p.cz / 2 < q.cz / 2;
            };
            std::stable_sort(by_block.begin(), by_block.end(), block_less);
            std::vector<Block> blocks;
            std::vector<uint32_t> block_of(clusters.size(), 0);
            for (size_t k = 0; k < by_block.size(); ++k)
            {
                Cluster const& cl = clusters[by_block[k]];
                if (k == 0 || block_less(by_block[k - 1], by_block[k]))
                {
                    Block nb;
                    nb.body = cl.body;
                    nb.bx   = cl.cx / 2;
                    nb.by   = cl.cy / 2;
                    nb.bz   = cl.cz / 2;
                    blocks.push_back(nb);
                }
                Block& bl = blocks.back();
                block_of[by_block[k]] = static_cast<uint32_t>(blocks.size() - 1);
                bl.weight += cl.count;
                for (uint32_t p = cl.first; p < cl.first + cl.count; ++p)
                    for (int a = 0; a < 4; ++a)
                        for (int d = 0; d < 3; ++d)
                            bl.c[d] += scene.x0[3 * static_cast<size_t>(scene.tets[4 * static_cast<size_t>(sorted[p]) + a]) + d];
            }
            for (Block& bl : blocks)
                for (int d = 0; d < 3; ++d)
                    bl.c[d] /= 4.0 * static_cast<double>(std::max<int64_t>(bl.weight, 1));
            std::vector<uint32_t> order(blocks.size());
            std::iota(order.begin(), order.end(), 0u);
            struct Job
            {
                size_t begin, end;
                int32_t r0, r1;
            };
            std::vector<Job> stack{{0, order.size(), 0, n_regions}};
            while (!stack.empty())
            {
                Job const j = stack.back();
                stack.pop_back();
                if (j.r1 - j.r0 <= 1 || j.end - j.begin <= 1)
                {
                    for (size_t k = j.begin; k < j.end; ++k)
                        blocks[order[k]].region = j.r0;
                    continue;
                }
                double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
                int64_t weight = 0;
                for (size_t k = j.begin; k < j.end; ++k)
                {
                    for (int d = 0; d < 3; ++d)
                    {
                        lo[d] = std::min(lo[d], blocks[order[k]].c[d]);
                        hi[d] = std::max(hi[d], blocks[order[k]].c[d]);
                    }
                    weight += blocks[order[k]].weight;
                }
                int ax = 0;
                for (int d = 1; d < 3; ++d)
                    if (hi[d] - lo[d] > hi[ax] - lo[ax])
                        ax = d;
                int const b = (ax + 1) % 3, c = (ax + 2) % 3;
                std::sort(order.begin() + static_cast<std::ptrdiff_t>(j.begin), order.begin() + static_cast<std::ptrdiff_t>(j.end),
                          [&](uint32_t x, uint32_t y) {
                              Block const &p = blocks[x], &q = blocks[y];
                              if (p.c[ax] != q.c[ax])
                                  return p.c[ax] < q.c[ax];
                              if (p.c[b] != q.c[b])
                                  return p.c[b] < q.c[b];
                              if (p.c[c] != q.c[c])
                                  return p.c[c] < q.c[c];
                              return x < y;
                          });
                int32_t const left = (j.r1 - j.r0) / 2;
                int64_t const want = weight * left / (j.r1 - j.r0);
                int64_t seen_w     = 0;
                size_t cut         = j.begin;
                while (cut < j.end - 1 && seen_w + blocks[order[cut]].weight / 2 < want)
                    seen_w += blocks[order[cut++]].weight;
                cut = std::max(cut, j.begin + 1);
                stack.push_back({j.begin, cut, j.r0, j.r0 + left});
                stack.push_back({cut, j.end, j.r0 + left, j.r1});
            }
            for (size_t i = 0; i < clusters.size(); ++i)
                clusters[i].region = blocks[block_of[i]].region;
        }
        int64_t seen   = 0;
        int32_t region = 0;
        for (size_t k = 0; pencils && k < idx.size(); ++k)
        { // balanced by tet count; a pencil column is never split
            uint32_t const i = idx[k];
            bool const new_unit =
                !pencils || k == 0 || clusters[idx[k - 1]].body != clusters[i].body || key[idx[k - 1]] != key[i];
            if (new_unit)
                region = static_cast<int32_t>(std::min<int64_t>(n_regions - 1, (seen * n_regions) / T));
            clusters[i].region = region;
            seen += clusters[i].count;
        }
    }
    else
        out.n_regions = 1;

    for (Cluster const& c : clusters)
        for (uint32_t m = 0; m < c.count; ++m)
            out.tet_region[sorted[c.first + m]] = c.region;

    // Renumber the colours so that consecutive colours hand a shared vertex over INSIDE a region as often as
    // possible.  D[a][b] = shared vertices whose clusters of colours a and b lie in different regions; the
    // order is the cycle of least total D (exhaustive up to 9 colours, nearest neighbour + 2-opt beyond).  On a
    // lattice this is a Gray code of the cell parities: every step then depends on other regions across one
    // face orientation only, and with pencil regions the steps that flip the parity along the pencil axis
    // depend on no other region at all.  Any order of the colours is a valid Gauss-Seidel order.
    if (resident && out.n_regions > 1 && out.n_colours > 2)
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
            for (int32_t i = 0; i < C; ++i) // a cycle: the last colour of a sweep hands over to the first of the next
                s += D[static_cast<size_t>(o[i]) * C + o[(i + 1) % C]];
            return s;
        };
        std::vector<int32_t> best(static_cast<size_t>(C));
        std::iota(best.begin(), best.end(), 0);
        int64_t best_cost = cost(best);
        if (C <= 9)
        {
            std::vector<int32_t> o(best);
            while (std::next_permutation(o.begin() + 1, o.end())) // a cycle: colour 0 stays first
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
        if (region_plan)
        {
            classify_regions(scene, out.tet_region, out.n_regions, *region_plan);
            for (Cluster& c : clusters)
            { // part 0: the cluster has a vertex that another region touches too
                c.part = 1;
                for (uint32_t k = 0; k < c.nv; ++k)
                    if (region_plan->vertex_region[c.verts[k]] != c.region)
                        c.part = 0;
            }
            // A vertex is OWNED (predicted, collided, committed, its surface copy written) by the region of the
            // first cluster of a sweep that touches it: after the owner's steps its next touch is then local, and the
            // contacts a rank of a decomposed scene detects for the vertices it owns are the ones it projects.
            {
                std::vector<int32_t> first_colour(static_cast<size_t>(V), 0x7fffffff);
                for (Cluster const& c : clusters)
                    for (uint32_t k = 0; k < c.nv; ++k)
                        if (c.colour < first_colour[c.verts[k]])
                        {
                            first_colour[c.verts[k]]              = c.colour;
                            region_plan->vertex_owner[c.verts[k]] = c.region;
                        }
            }
            // Inside a chunk the clusters that PUSH in their step come first, then those that wait for a record
            // pushed in the step before: the first clusters of a step run on the warps that have a sub-partition
            // to themselves (item_rotation) and finish early, so what a neighbour waits for leaves early, and who
            // waits does so on a warp the others do not wait for.
            {
                struct Touch4
                {
                    uint32_t v;
                    int32_t colour, region;
                    uint32_t cluster;
                };
                std::vector<Touch4> tl;
                for (size_t ci = 0; ci < clusters.size(); ++ci)
                    if (clusters[ci].part == 0)
                        for (uint32_t k = 0; k < clusters[ci].nv; ++k)
                            if (region_plan->vertex_region[clusters[ci].verts[k]] < 0)
                                tl.push_back({clusters[ci].verts[k], clusters[ci].colour, clusters[ci].region,
                                              static_cast<uint32_t>(ci)});
                std::sort(tl.begin(), tl.end(), [](Touch4 const& a, Touch4 const& b) {
                    return a.v != b.v ? a.v < b.v : a.colour < b.colour;
                });
                int32_t const C = out.n_colours;
                for (size_t i = 0; i < tl.size();)
                {
                    size_t j = i;
                    while (j < tl.size() && tl[j].v == tl[i].v)
                        ++j;
                    size_t const n = j - i;
                    for (size_t t = 0; t < n; ++t)
                    {
                        Touch4 const& me   = tl[i + t];
                        Touch4 const& next = tl[i + (t + 1) % n];
                        Touch4 const& prev = tl[i + (t + n - 1) % n];
                        Cluster& c         = clusters[me.cluster];
                        if (!resident->push_first)
                            continue;
                        if (next.region != me.region)
                            c.prio = 0;
                        else if (prev.region != me.region && (me.colour - prev.colour + C) % C == 1)
                            c.prio = std::min(c.prio, 1);
                    }
                    i = j;
                }
            }
        }
        // threads per CTA: every cluster of a (colour, region) step gets its own thread when possible
        // (cluster i of the step, part 0 first, runs on thread (i + rot) % nt); more clusters than threads:
        // as many rounds as the widest step needs, of equal width
        std::vector<int64_t> cnt(static_cast<size_t>(out.n_colours) * static_cast<size_t>(out.n_regions), 0);
        int64_t need = 0;
        for (Cluster const& c : clusters)
            need = std::max(need, ++cnt[static_cast<size_t>(c.colour) * out.n_regions + c.region]);
        int64_t const rounds = std::max<int64_t>(1, (need + resident->max_threads - 1) / resident->max_threads);
        int64_t const width  = (need + rounds - 1) / rounds;
        out.nt  = static_cast<int32_t>(std::min<int64_t>(resident->max_threads, std::max<int64_t>(64, (width + 31) / 32 * 32)));
        out.rot = resident->rotate_items ? item_rotation(out.nt) : 0;
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
        if (x.prio != y.prio)
            return x.prio < y.prio;
        if (x.body != y.body)
            return x.body < y.body;
        return x.morton < y.morton;
    });

    size_t const n_chunks = static_cast<size_t>(out.n_colours) * static_cast<size_t>(out.n_regions) * 2;
    out.chunks.assign(n_chunks, ChunkDesc{});
    out.storage_order.resize(static_cast<size_t>(T));
    out.serial_order.reserve(static_cast<size_t>(T));
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
                out.storage_order[static_cast<size_t>(col_base + static_cast<int64_t>(k - i))] = sorted[c.first + m];
            }
            col_base += d.n[m];
        }
        store = col_base;
        pair_clusters = (part == 0 ? 0 : pair_clusters) + static_cast<int64_t>(j - i);
        out.max_chunk_clusters = std::max<int64_t>(out.max_chunk_clusters, pair_clusters);
        i = j;
    }
}

bool build_exchange_plan(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& plan, int world,
                         ExchangePlan& out)
{
    out = ExchangePlan{};
    int64_t const V = scene.n_vertices(), T = scene.n_tets(), Q = cp.n_clusters;
    int32_t const Rn = cp.n_regions, C = cp.n_colours;
    out.n_regions = Rn;
    out.n_colours = C;
    out.world     = world;
    if (world < 1 || Rn % world != 0 || plan.n_regions != Rn || static_cast<int64_t>(plan.vertex_region.size()) != V)
    {
        out.why_not = "bad partition (world must divide the region count)";
        return false;
    }
    if (!cp.why_not.empty())
    {
        out.why_not = cp.why_not;
        return false;
    }
    // ---- clusters back from the chunks: colour, region, part, tets --------------------------------------
    struct Cl
    {
        int32_t colour, region, part;
        int64_t xq; // number among the exchange clusters, -1 for part 1
        uint32_t n_shared;
        uint32_t shared[kMaxClusterVertices];
    };
    std::vector<Cl> cl(static_cast<size_t>(Q));
    std::vector<char> touched(static_cast<size_t>(V), 0);
    for (uint32_t v : scene.tets)
        touched[v] = 1;
    auto const is_shared = [&](uint32_t v) { return touched[v] && plan.vertex_region[v] < 0; };
    out.chunk_xfirst.assign(static_cast<size_t>(C) * Rn, 0);
    int64_t nx = 0;
    uint32_t most_shared = 0;
    std::vector<int32_t> pos_cluster(static_cast<size_t>(T), 0); // storage position -> cluster
    for (size_t ch = 0; ch < cp.chunks.size(); ++ch)
    {
        ChunkDesc const& d = cp.chunks[ch];
        int32_t const col = static_cast<int32_t>(ch / (2 * static_cast<size_t>(Rn)));
        int32_t const reg = static_cast<int32_t>((ch / 2) % static_cast<size_t>(Rn));
        int32_t const part = static_cast<int32_t>(ch % 2);
        if (part == 0)
            out.chunk_xfirst[static_cast<size_t>(col) * Rn + reg] = static_cast<int32_t>(nx);
        for (int32_t i = 0; i < d.n[0]; ++i)
        {
            Cl& c      = cl[static_cast<size_t>(d.cfirst + i)];
            c.colour   = col;
            c.region   = reg;
            c.part     = part;
            c.xq       = part == 0 ? nx++ : -1;
            c.n_shared = 0;
            int64_t base = d.first;
            for (int m = 0; m < kMaxCluster && i < d.n[m]; ++m)
            {
                int64_t const pos = base + i;
                pos_cluster[static_cast<size_t>(pos)] = d.cfirst + i;
                uint32_t const t = cp.storage_order[static_cast<size_t>(pos)];
                for (int a = 0; a < 4; ++a)
                {
                    uint32_t const v = scene.tets[4 * static_cast<size_t>(t) + a];
                    if (is_shared(v) && std::find(c.shared, c.shared + c.n_shared, v) == c.shared + c.n_shared)
                    {
                        if (c.n_shared >= static_cast<uint32_t>(kMaxClusterVertices))
                        {
                            out.why_not = "a cluster has more than 16 shared vertices";
                            return false;
                        }
                        c.shared[c.n_shared++] = v;
                    }
                }
                base += d.n[m];
            }
            if ((c.n_shared > 0) != (part == 0))
            {
                out.why_not = "cluster parts and vertex classification disagree";
                return false;
            }
            most_shared = std::max(most_shared, c.n_shared);
        }
    }
    out.n_xclusters = nx;
    out.entries     = static_cast<int32_t>(std::max<uint32_t>(4u, (most_shared + 3u) / 4u * 4u));

    // ---- local vertex tables: owned vertices first, then guests -------------------------------------------
    out.vertex_owner = plan.vertex_owner;
    out.loc_off.assign(static_cast<size_t>(Rn) + 1, 0);
    out.n_owned.assign(static_cast<size_t>(Rn), 0);
    for (int64_t v = 0; v < V; ++v)
        ++out.n_owned[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)])];
    // guests: (region, shared vertex) pairs with region != owner
    std::vector<std::pair<int32_t, uint32_t>> guests;
    for (Cl const& c : cl)
        for (uint32_t k = 0; k < c.n_shared; ++k)
            if (plan.vertex_owner[c.shared[k]] != c.region)
                guests.emplace_back(c.region, c.shared[k]);
    std::sort(guests.begin(), guests.end());
    guests.erase(std::unique(guests.begin(), guests.end()), guests.end());
    {
        std::vector<int32_t> n_guests(static_cast<size_t>(Rn), 0);
        for (auto const& g : guests)
            ++n_guests[static_cast<size_t>(g.first)];
        for (int32_t r = 0; r < Rn; ++r)
        {
            int64_t const n = static_cast<int64_t>(out.n_owned[static_cast<size_t>(r)]) + n_guests[static_cast<size_t>(r)];
            out.max_local   = std::max(out.max_local, n);
            out.loc_off[static_cast<size_t>(r) + 1] = out.loc_off[static_cast<size_t>(r)] + static_cast<int32_t>(n);
        }
    }
    if (out.max_local > 65535)
    {
        out.why_not = "a region touches more than 65535 vertices (16-bit slots)";
        return false;
    }
    out.loc_vtx.assign(static_cast<size_t>(out.loc_off.back()), 0);
    std::vector<uint32_t> owner_slot(static_cast<size_t>(V), 0);
    std::vector<uint32_t> guest_slot(guests.size(), 0); // parallel to guests
    std::vector<int64_t> first_guest(static_cast<size_t>(Rn) + 1, 0);
    {
        std::vector<int32_t> cur(static_cast<size_t>(Rn), 0);
        for (int64_t v = 0; v < V; ++v)
        {
            int32_t const r = plan.vertex_owner[static_cast<size_t>(v)];
            owner_slot[static_cast<size_t>(v)] = static_cast<uint32_t>(cur[static_cast<size_t>(r)]);
            out.loc_vtx[static_cast<size_t>(out.loc_off[static_cast<size_t>(r)] + cur[static_cast<size_t>(r)]++)] =
                static_cast<uint32_t>(v);
        }
        for (size_t g = 0; g < guests.size(); ++g) // sorted by (region, vertex)
        {
            guest_slot[g] = static_cast<uint32_t>(cur[static_cast<size_t>(guests[g].first)]++);
            ++first_guest[static_cast<size_t>(guests[g].first) + 1];
        }
        for (int32_t r = 0; r < Rn; ++r)
            first_guest[static_cast<size_t>(r) + 1] += first_guest[static_cast<size_t>(r)];
    }
    auto const guest_index = [&](int32_t r, uint32_t v) -> size_t {
        return static_cast<size_t>(std::lower_bound(guests.begin() + first_guest[static_cast<size_t>(r)],
                                                    guests.begin() + first_guest[static_cast<size_t>(r) + 1],
                                                    std::make_pair(r, v)) -
                                   guests.begin());
    };
    auto const slot_of = [&](int32_t r, uint32_t v) -> uint32_t {
        return plan.vertex_owner[v] == r ? owner_slot[v] : guest_slot[guest_index(r, v)];
    };
    // ---- slots spread over the shared-memory banks ---------------------------------------------------------
    // A vertex slot is 16 bytes (fp32): the 32 banks form 8 columns, column = slot % 8, and a warp's 128-bit access
    // costs as many wavefronts as the fullest column holds distinct slots (4 when 32 lanes spread evenly).  Which
    // slots one warp instruction touches is static — lane = cluster of the step, instruction = (tet m of the cluster,
    // corner a) — so the numbering inside the owned range and inside the guest range of every region is chosen
    // to flatten those histograms (vertex order of a lattice put 12 wavefronts where 4 suffice: clusters of one
    // colour are two cells apart, so their vertices fell into every other column).  Deterministic: every rank of a
    // decomposed scene plans the same tables.
#ifndef SBSB200_NO_BANK_SPREAD
    if (cp.nt > 0 && cp.nt % 32 == 0)
#else
    if (false)
#endif
    {
        constexpr int kCols = 8;
        int32_t const nt = cp.nt, rot = cp.rot, warps = nt / 32;
        std::vector<int32_t> grp_off, grp_mem;      // groups of one region: local vertex numbers (current slots)
        std::vector<int32_t> v_off, v_grp;          // vertex -> groups
        std::vector<int32_t> col, hist, members[2][kCols];
        for (int32_t r = 0; r < Rn; ++r)
        {
            int32_t const no = out.n_owned[static_cast<size_t>(r)];
            int32_t const nl = out.loc_off[static_cast<size_t>(r) + 1] - out.loc_off[static_cast<size_t>(r)];
            if (nl <= kCols)
                continue;
            // (1) the access groups of the region
            grp_off.assign(1, 0);
            grp_mem.clear();
            for (int32_t c = 0; c < C; ++c)
            {
                ChunkDesc const* d = &cp.chunks[(static_cast<size_t>(c) * Rn + r) * 2];
                int32_t const nA = d[0].n[0], nB = d[1].n[0], rounds = (nA + nB + nt - 1) / nt;
                for (int32_t round = 0; round < rounds; ++round)
                    for (int32_t w = 0; w < warps; ++w)
                        for (int m = 0; m < kMaxCluster; ++m)
                            for (int a = 0; a < 4; ++a)
                            {
                                for (int32_t lane = 0; lane < 32; ++lane)
                                {
                                    int32_t const tid  = 32 * w + lane;
                                    int32_t const item = round * nt + (tid >= rot ? tid - rot : tid + nt - rot);
                                    if (item >= nA + nB)
                                        continue;
                                    ChunkDesc const& ch = item < nA ? d[0] : d[1];
                                    int32_t const ci    = item < nA ? item : item - nA;
                                    if (ci >= ch.n[m])
                                        continue;
                                    int64_t pos = ch.first + ci;
                                    for (int j = 0; j < m; ++j)
                                        pos += ch.n[j];
                                    uint32_t const v = scene.tets[4 * static_cast<size_t>(cp.storage_order[static_cast<size_t>(pos)]) + a];
                                    grp_mem.push_back(static_cast<int32_t>(slot_of(r, v)));
                                }
                                if (static_cast<int32_t>(grp_mem.size()) - grp_off.back() >= 2)
                                    grp_off.push_back(static_cast<int32_t>(grp_mem.size()));
                                else
                                    grp_mem.resize(static_cast<size_t>(grp_off.back()));
                            }
            }
            int32_t const G = static_cast<int32_t>(grp_off.size()) - 1;
            if (G == 0)
                continue;
            v_off.assign(static_cast<size_t>(nl) + 1, 0);
            for (int32_t v : grp_mem)
                ++v_off[static_cast<size_t>(v) + 1];
            for (int32_t v = 0; v < nl; ++v)
                v_off[static_cast<size_t>(v) + 1] += v_off[static_cast<size_t>(v)];
            v_grp.assign(grp_mem.size(), 0);
            {
                std::vector<int32_t> cur(v_off.begin(), v_off.end() - 1);
                for (int32_t g = 0; g < G; ++g)
                    for (int32_t i = grp_off[static_cast<size_t>(g)]; i < grp_off[static_cast<size_t>(g) + 1]; ++i)
                        v_grp[static_cast<size_t>(cur[static_cast<size_t>(grp_mem[static_cast<size_t>(i)])]++)] = g;
            }
            // (2) columns: start from the vertex order, improve by swaps inside the owned / guest range
            col.assign(static_cast<size_t>(nl), 0);
            hist.assign(static_cast<size_t>(G) * kCols, 0);
            for (int32_t v = 0; v < nl; ++v)
                col[static_cast<size_t>(v)] = v % kCols;
            for (int32_t g = 0; g < G; ++g)
                for (int32_t i = grp_off[static_cast<size_t>(g)]; i < grp_off[static_cast<size_t>(g) + 1]; ++i)
                    ++hist[static_cast<size_t>(g) * kCols + col[static_cast<size_t>(grp_mem[static_cast<size_t>(i)])]];
            auto const wavefronts = [&]() {
                int64_t n = 0;
                for (int32_t g = 0; g < G; ++g)
                    n += *std::max_element(&hist[static_cast<size_t>(g) * kCols], &hist[static_cast<size_t>(g) * kCols] + kCols);
                return n;
            };
            for (int32_t g = 0; g < G; ++g)
                out.bank_wavefronts_ideal += (grp_off[static_cast<size_t>(g) + 1] - grp_off[static_cast<size_t>(g)] + kCols - 1) / kCols;
            out.bank_wavefronts_before += wavefronts();
            // The search minimises the sum over the groups of the squared column counts (flat histograms minimise it, and
            // the change of a move is linear in the counts): moving v from column a to column k changes it by
            // 2 * (S[k] - S[a] + deg(v)) with S = sum of the histograms of v's groups.
            auto const column_sums = [&](int32_t v, int32_t* S) {
                for (int k = 0; k < kCols; ++k)
                    S[k] = 0;
                for (int32_t i = v_off[static_cast<size_t>(v)]; i < v_off[static_cast<size_t>(v) + 1]; ++i)
                {
                    int32_t const* h = &hist[static_cast<size_t>(v_grp[static_cast<size_t>(i)]) * kCols];
                    for (int k = 0; k < kCols; ++k)
                        S[k] += h[k];
                }
            };
            auto const move = [&](int32_t v, int k) {
                int32_t const from = col[static_cast<size_t>(v)];
                for (int32_t i = v_off[static_cast<size_t>(v)]; i < v_off[static_cast<size_t>(v) + 1]; ++i)
                {
                    int32_t* h = &hist[static_cast<size_t>(v_grp[static_cast<size_t>(i)]) * kCols];
                    --h[from];
                    ++h[k];
                }
                col[static_cast<size_t>(v)] = k;
            };
            for (int range = 0; range < 2; ++range)
                for (int k = 0; k < kCols; ++k)
                    members[range][k].clear();
            std::vector<int32_t> at_in_pool(static_cast<size_t>(nl), 0);
            for (int32_t v = 0; v < nl; ++v)
            {
                std::vector<int32_t>& pool = members[v < no ? 0 : 1][col[static_cast<size_t>(v)]];
                at_in_pool[static_cast<size_t>(v)] = static_cast<int32_t>(pool.size());
                pool.push_back(v);
            }
            uint64_t rng = 0x9e3779b97f4a7c15ull ^ (static_cast<uint64_t>(r) * 0xbf58476d1ce4e5b9ull);
            auto const next_random = [&]() {
                rng ^= rng << 13;
                rng ^= rng >> 7;
                rng ^= rng << 17;
                return rng;
            };
            constexpr int kPasses = 5, kCandidates = 4;
            for (int pass = 0; pass < kPasses; ++pass)
            {
                int64_t improved = 0;
                for (int32_t v = 0; v < nl; ++v)
                {
                    int32_t const deg = v_off[static_cast<size_t>(v) + 1] - v_off[static_cast<size_t>(v)];
                    if (deg == 0)
                        continue;
                    int const range = v < no ? 0 : 1, from = col[static_cast<size_t>(v)];
                    int32_t S[kCols];
                    column_sums(v, S);
                    int best_k = -1;
                    int32_t best_d = 0;
                    for (int k = 0; k < kCols; ++k)
                        if (k != from && !members[range][k].empty() && S[k] - S[from] + deg < best_d)
                        {
                            best_d = S[k] - S[from] + deg;
                            best_k = k;
                        }
                    if (best_k < 0)
                        continue;
                    // a partner out of column best_k takes v's column: the one that loses least
                    move(v, best_k);
                    std::vector<int32_t>& pool = members[range][best_k];
                    size_t best_u = pool.size();
                    int32_t best_total = 0;
                    for (int t = 0; t < kCandidates; ++t)
                    {
                        size_t const ui = static_cast<size_t>(next_random() % pool.size());
                        int32_t const u = pool[ui];
                        int32_t Su[kCols];
                        column_sums(u, Su);
                        int32_t const total =
                            best_d + Su[from] - Su[best_k] + (v_off[static_cast<size_t>(u) + 1] - v_off[static_cast<size_t>(u)]);
                        if (total < best_total)
                        {
                            best_total = total;
                            best_u     = ui;
                        }
                    }
                    if (best_u == pool.size())
                    {
                        move(v, from);
                        continue;
                    }
                    int32_t const u = pool[best_u];
                    move(u, from);
                    pool[best_u]                       = v;
                    std::vector<int32_t>& mine         = members[range][from];
                    mine[static_cast<size_t>(at_in_pool[static_cast<size_t>(v)])] = u;
                    std::swap(at_in_pool[static_cast<size_t>(v)], at_in_pool[static_cast<size_t>(u)]);
                    at_in_pool[static_cast<size_t>(v)] = static_cast<int32_t>(best_u);
                    ++improved;
                }
                if (improved * 200 < nl)
                    break;
            }
            out.bank_wavefronts += wavefronts();
            // (3) new numbers: the slots of a column in rising order go to its vertices in rising order
            {
                std::vector<int32_t> new_slot(static_cast<size_t>(nl), 0);
                for (int range = 0; range < 2; ++range)
                {
                    int32_t const lo = range == 0 ? 0 : no, hi = range == 0 ? no : nl;
                    std::vector<int32_t> next(kCols, 0);
                    for (int k = 0; k < kCols; ++k)
                    {
                        next[static_cast<size_t>(k)] = lo + ((k - lo) % kCols + kCols) % kCols;
                        std::sort(members[range][k].begin(), members[range][k].end());
                    }
                    for (int k = 0; k < kCols; ++k)
                        for (int32_t v : members[range][k])
                        {
                            new_slot[static_cast<size_t>(v)] = next[static_cast<size_t>(k)];
                            next[static_cast<size_t>(k)] += kCols;
                            if (new_slot[static_cast<size_t>(v)] >= hi)
                            {
                                out.why_not = "bank spreading lost a slot";
                                return false;
                            }
                        }
                }
                for (int32_t i = 0; i < no; ++i) // loc_vtx still holds the owned vertices in vertex order
                    owner_slot[out.loc_vtx[static_cast<size_t>(out.loc_off[static_cast<size_t>(r)] + i)]] =
                        static_cast<uint32_t>(new_slot[static_cast<size_t>(i)]);
                for (int64_t g = first_guest[static_cast<size_t>(r)]; g < first_guest[static_cast<size_t>(r) + 1]; ++g)
                    guest_slot[static_cast<size_t>(g)] = static_cast<uint32_t>(new_slot[guest_slot[static_cast<size_t>(g)]]);
            }
        }
    }
    for (int64_t v = 0; v < V; ++v)
        out.loc_vtx[static_cast<size_t>(out.loc_off[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)])]) +
                    owner_slot[static_cast<size_t>(v)]] = static_cast<uint32_t>(v);
    for (size_t g = 0; g < guests.size(); ++g)
        out.loc_vtx[static_cast<size_t>(out.loc_off[static_cast<size_t>(guests[g].first)]) + guest_slot[g]] = guests[g].second;
    out.tet_slots.assign(4 * static_cast<size_t>(T), 0);
    for (int64_t pos = 0; pos < T; ++pos)
    {
        uint32_t const t = cp.storage_order[static_cast<size_t>(pos)];
        int32_t const r  = cl[static_cast<size_t>(pos_cluster[static_cast<size_t>(pos)])].region;
        for (int a = 0; a < 4; ++a)
        {
            uint32_t const v = scene.tets[4 * static_cast<size_t>(t) + a];
            if (plan.vertex_owner[v] != r && !is_shared(v))
            {
                out.why_not = "a private vertex is owned by another region";
                return false;
            }
            out.tet_slots[4 * static_cast<size_t>(pos) + a] = static_cast<uint16_t>(slot_of(r, v));
        }
    }

    // ---- shared vertices: owner records and the touch chains ---------------------------------------------
    out.n_entries = static_cast<uint32_t>(static_cast<int64_t>(out.entries) * nx);
    std::vector<char> is_surface(static_cast<size_t>(V), 0);
    for (HostBody const& hb : scene.bodies)
        if (hb.kind == BodyKind::tet)
            for (uint32_t lv : hb.surf_to_tet)
                is_surface[static_cast<size_t>(hb.v_offset + lv)] = 1;
    out.osv_off.assign(static_cast<size_t>(Rn) + 1, 0);
    std::vector<uint32_t> osv_pos(static_cast<size_t>(V), kRouteNone);
    for (int64_t v = 0; v < V; ++v)
        if (is_shared(static_cast<uint32_t>(v)))
        {
            ++out.osv_off[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)]) + 1];
            ++out.n_shared;
        }
    for (int32_t r = 0; r < Rn; ++r)
        out.osv_off[static_cast<size_t>(r) + 1] += out.osv_off[static_cast<size_t>(r)];
    if (static_cast<uint64_t>(out.n_entries) + static_cast<uint64_t>(out.n_shared) >= kRouteIndexMask)
    {
        out.why_not = "too many mailboxes for 28-bit routing words";
        return false;
    }
    out.osv_slot.assign(static_cast<size_t>(out.n_shared), 0);
    out.osv_meta.assign(static_cast<size_t>(out.n_shared), 0);
    out.osv_first.assign(static_cast<size_t>(out.n_shared), kRouteNone);
    {
        std::vector<int32_t> cur(out.osv_off.begin(), out.osv_off.end() - 1);
        for (int64_t v = 0; v < V; ++v)
            if (is_shared(static_cast<uint32_t>(v)))
            {
                int32_t const at = cur[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)])]++;
                osv_pos[static_cast<size_t>(v)]      = static_cast<uint32_t>(at);
                out.osv_slot[static_cast<size_t>(at)] = owner_slot[static_cast<size_t>(v)];
            }
    }
    auto const rank_of = [&](int32_t region) { return static_cast<uint32_t>(region_rank(region, Rn, world)); };
    struct Touch
    {
        uint32_t vertex;
        int32_t colour, region;
        uint32_t entry;
        int64_t xq;
    };
    std::vector<Touch> touches;
    for (Cl const& c : cl)
        for (uint32_t k = 0; k < c.n_shared; ++k)
            touches.push_back({c.shared[k], c.colour, c.region, k, c.xq});
    std::sort(touches.begin(), touches.end(), [](Touch const& x, Touch const& y) {
        return x.vertex != y.vertex ? x.vertex < y.vertex : x.colour < y.colour;
    });
    auto const entry_route = [&](Touch const& t) {
        return static_cast<uint32_t>(static_cast<int64_t>(t.entry) * nx + t.xq) | (rank_of(t.region) << kRouteRankShift);
    };
    auto const owner_route = [&](uint32_t v) {
        return (out.n_entries + osv_pos[v]) | (rank_of(plan.vertex_owner[v]) << kRouteRankShift);
    };
    // per (variant, exchange cluster): the words, in the order of the cluster's shared vertices
    std::vector<std::vector<uint32_t>> pulls(4 * static_cast<size_t>(nx)), pushes(4 * static_cast<size_t>(nx));
    std::vector<char> colour_pulls(static_cast<size_t>(std::max(C, 1)), 0);
    out.pulls_by_colour.assign(static_cast<size_t>(std::max(C, 1)), 0);
    for (size_t i = 0; i < touches.size();)
    {
        size_t j = i;
        while (j < touches.size() && touches[j].vertex == touches[i].vertex)
            ++j;
        uint32_t const v    = touches[i].vertex;
        int32_t const owner = plan.vertex_owner[v];
        bool const surf     = is_surface[v] != 0;
        Touch const& first  = touches[i];
        Touch const& last   = touches[j - 1];
        for (size_t t = i; t + 1 < j; ++t)
            if (touches[t + 1].colour == touches[t].colour)
            {
                out.why_not = "two clusters of one colour touch the same vertex";
                return false;
            }
        out.osv_meta[osv_pos[v]]  = static_cast<uint32_t>(last.colour) | (surf ? kOsvSurface : 0u) |
                                   (last.region != owner ? kOsvLastRemote : 0u) |
                                   (first.region != owner ? kOsvFirstRemote : 0u);
        out.osv_first[osv_pos[v]] = entry_route(first);
        for (size_t t = i; t < j; ++t)
        {
            Touch const& me     = touches[t];
            uint32_t const slot = slot_of(me.region, v);
            for (int later = 0; later < 2; ++later)
                for (int cs = 0; cs < 2; ++cs)
                { // the previous touch: a colour of the same sweep, the owner's collision step, the last colour of
                  // the previous sweep, or the owner's predict step
                    int32_t from, d;
                    if (t > i)
                    {
                        from = touches[t - 1].region;
                        d    = me.colour - touches[t - 1].colour;
                    }
                    else if (cs && surf)
                    {
                        from = owner;
                        d    = me.colour + 1;
                    }
                    else if (later)
                    {
                        from = last.region;
                        d    = C + cs + me.colour - last.colour;
                    }
                    else
                    {
                        from = owner;
                        d    = static_cast<int32_t>(kPullPredict);
                    }
                    if (from != me.region)
                    {
                        pulls[static_cast<size_t>(2 * later + cs) * nx + me.xq].push_back(
                            kPullValid | (me.entry << 24) | (static_cast<uint32_t>(d) << 16) | slot);
                        ++out.n_pulls[2 * later + cs];
                        if (later && !cs)
                        {
                            colour_pulls[static_cast<size_t>(me.colour)] = 1;
                            ++out.pulls_by_colour[static_cast<size_t>(me.colour)];
                        }
                    }
                }
            for (int final_sweep = 0; final_sweep < 2; ++final_sweep)
                for (int cs = 0; cs < 2; ++cs)
                { // the next touch: a colour of the same sweep, the owner (collision step of the next sweep, or
                  // commit), or the first colour of the next sweep
                    int32_t to;
                    uint32_t route;
                    if (t + 1 < j)
                    {
                        to    = touches[t + 1].region;
                        route = entry_route(touches[t + 1]);
                    }
                    else if ((cs && surf) || final_sweep)
                    {
                        to    = owner;
                        route = owner_route(v);
                    }
                    else
                    {
                        to    = first.region;
                        route = entry_route(first);
                    }
                    if (to != me.region)
                    {
                        auto& w = pushes[static_cast<size_t>(2 * final_sweep + cs) * nx + me.xq];
                        w.push_back(kPullValid | slot);
                        w.push_back(route);
                        ++out.n_pushes[2 * final_sweep + cs];
                    }
                }
        }
        i = j;
    }
    for (int32_t c = 0; c < C; ++c)
        out.quiet_steps += !colour_pulls[static_cast<size_t>(c)];
    for (int64_t xq = 0; xq < nx; ++xq)
    { // later sweep, collision steps present: how many records a cluster pulls / pushes in its step
        ++out.pull_hist[std::min<size_t>(16, pulls[3 * static_cast<size_t>(nx) + xq].size())];
        ++out.push_hist[std::min<size_t>(16, pushes[1 * static_cast<size_t>(nx) + xq].size() / 2)];
    }
    size_t const E = static_cast<size_t>(out.entries);
    out.pull.assign(4 * E * static_cast<size_t>(nx), 0u);
    out.push.assign(4 * 2 * E * static_cast<size_t>(nx), 0u);
    for (size_t var = 0; var < 4; ++var)
        for (int64_t xq = 0; xq < nx; ++xq)
        {
            auto const& pl = pulls[var * static_cast<size_t>(nx) + xq];
            for (size_t k = 0; k < pl.size(); ++k) // [variant][k / 4][xq][k % 4]
                out.pull[((var * (E / 4) + k / 4) * static_cast<size_t>(nx) + xq) * 4 + k % 4] = pl[k];
            auto const& ps = pushes[var * static_cast<size_t>(nx) + xq];
            for (size_t k = 0; k < ps.size() / 2; ++k) // [variant][k / 2][xq][2 * (k % 2) ..]
            {
                size_t const at = ((var * (E / 2) + k / 2) * static_cast<size_t>(nx) + xq) * 4 + 2 * (k % 2);
                out.push[at]     = ps[2 * k];
                out.push[at + 1] = ps[2 * k + 1];
            }
        }

    // ---- surface vertices by owner -----------------------------------------------------------------------
    std::vector<uint32_t> sgv; // global vertex of surface vertex i (the order of DeviceScene::surf_v)
    for (HostBody const& hb : scene.bodies)
        if (hb.kind == BodyKind::tet)
            for (uint32_t lv : hb.surf_to_tet)
                sgv.push_back(static_cast<uint32_t>(hb.v_offset + lv));
    out.surf_off.assign(static_cast<size_t>(Rn) + 1, 0);
    for (uint32_t gv : sgv)
        ++out.surf_off[static_cast<size_t>(plan.vertex_owner[gv]) + 1];
    for (int32_t r = 0; r < Rn; ++r)
        out.surf_off[static_cast<size_t>(r) + 1] += out.surf_off[static_cast<size_t>(r)];
    out.surf_slot.assign(sgv.size(), 0);
    out.surf_index.assign(sgv.size(), 0);
    out.surf_osv.assign(sgv.size(), kRouteNone);
    {
        std::vector<int32_t> cur(out.surf_off.begin(), out.surf_off.end() - 1);
        for (size_t s = 0; s < sgv.size(); ++s)
        {
            uint32_t const gv = sgv[s];
            int32_t const at  = cur[static_cast<size_t>(plan.vertex_owner[gv])]++;
            out.surf_slot[static_cast<size_t>(at)]  = owner_slot[gv];
            out.surf_index[static_cast<size_t>(at)] = static_cast<uint32_t>(s);
            out.surf_osv[static_cast<size_t>(at)]   = osv_pos[gv];
        }
    }
    return true;
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
