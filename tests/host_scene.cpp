// TEST-ONLY harness: C entry points over the product's host-side scene compiler
// (soft-body-simulator_b200/csrc/scene_build.cpp) so that boundary extraction, colouring and
// the region plan can be tested on a CPU-only box.
#include "../soft-body-simulator_b200/csrc/scene_build.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace sbsb200;

extern "C" int hs_boundary(int64_t nV, int64_t nT, const uint32_t* tets, uint32_t* surf_to_tet, uint32_t* tris,
                           int64_t* n_tris)
{
    std::vector<uint32_t> s2t, tr;
    extract_boundary(nV, nT, tets, s2t, &tr);
    std::memcpy(surf_to_tet, s2t.data(), sizeof(uint32_t) * s2t.size());
    if (tris)
        std::memcpy(tris, tr.data(), sizeof(uint32_t) * tr.size());
    *n_tris = static_cast<int64_t>(tr.size() / 3);
    return static_cast<int>(s2t.size());
}

// colours k-vertex constraints; writes order[n], offsets[n_colours+1]; returns n_colours (<0 on failure)
extern "C" int hs_colour(int64_t nV, int64_t n, int k, const uint32_t* verts, const double* x0, uint32_t* order,
                         int64_t* offsets, int max_colours, int* valid)
{
    std::vector<uint64_t> keys;
    morton_keys(n, k, verts, x0, nV, keys);
    ColourClass cc;
    if (!colour_constraints(nV, n, k, verts, keys.data(), nullptr, max_colours, cc))
        return -1;
    std::memcpy(order, cc.order.data(), sizeof(uint32_t) * cc.order.size());
    std::memcpy(offsets, cc.offsets.data(), sizeof(int64_t) * cc.offsets.size());
    *valid = colouring_is_valid(nV, k, verts, cc) ? 1 : 0;
    return cc.n_colours;
}

// region plan of a single tet body; outputs tet_region[T], vertex_region[V]; returns n_interface
extern "C" int64_t hs_regions(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_regions,
                              int32_t* tet_region, int32_t* vertex_region, int32_t* n_neighbours)
{
    HostScene h;
    h.x0.assign(x0, x0 + 3 * nV);
    h.mass.assign(static_cast<size_t>(nV), 1.0);
    h.tets.assign(tets, tets + 4 * nT);
    h.tet_insertion.resize(static_cast<size_t>(nT));
    h.tet_material.assign(static_cast<size_t>(nT), 0);
    HostBody b;
    b.n_vertices = nV;
    b.n_tets     = nT;
    h.bodies.push_back(b);
    std::vector<uint64_t> keys;
    morton_keys(nT, 4, tets, x0, nV, keys);
    RegionPlan plan;
    plan_regions(h, keys, n_regions, plan);
    std::memcpy(tet_region, plan.tet_region.data(), sizeof(int32_t) * plan.tet_region.size());
    std::memcpy(vertex_region, plan.vertex_region.data(), sizeof(int32_t) * plan.vertex_region.size());
    for (int r = 0; r < plan.n_regions; ++r)
        n_neighbours[r] = plan.nbr_offsets[r + 1] - plan.nbr_offsets[r];
    return plan.n_interface;
}

// thread (i + rotation) % nt runs cluster i of a step of the resident schedule (scene_build.h)
extern "C" int hs_item_rotation(int nt) { return item_rotation(nt); }

// clustered colouring of a single tet body (or `n_bodies` copies laid out along x); outputs
// serial_order[T], storage_order[T], tet_region[T]; returns n_colours, or -1 if the plan is invalid
extern "C" int hs_cluster_plan(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_bodies,
                               int n_regions, int per_body, int64_t smem_bytes, uint32_t* serial_order,
                               uint32_t* storage_order,
                               int32_t* tet_region, int64_t* n_clusters, int64_t* max_chunk)
{
    HostScene h;
    for (int b = 0; b < n_bodies; ++b)
    {
        HostBody hb;
        hb.v_offset   = h.n_vertices();
        hb.n_vertices = nV;
        hb.t_offset   = h.n_tets();
        hb.n_tets     = nT;
        for (int64_t i = 0; i < nV; ++i)
        {
            h.x0.push_back(x0[3 * i] + 1000.0 * b);
            h.x0.push_back(x0[3 * i + 1]);
            h.x0.push_back(x0[3 * i + 2]);
            h.mass.push_back(1.0);
        }
        for (int64_t t = 0; t < nT; ++t)
        {
            for (int a = 0; a < 4; ++a)
                h.tets.push_back(static_cast<uint32_t>(hb.v_offset + tets[4 * t + a]));
            h.tet_insertion.push_back(h.n_constraints++);
            h.tet_material.push_back(0);
        }
        h.bodies.push_back(hb);
    }
    ClusterPlan plan;
    if (smem_bytes > 0)
    { // resident layout: the exchange plan must exist and the local vertex tables must fit shared memory
        ResidentParams rp;
        rp.smem_bytes   = smem_bytes;
        rp.vertex_bytes = 16;
        rp.max_threads  = 512;
        RegionPlan regions;
        build_cluster_plan(h, n_regions, per_body != 0, plan, &rp, &regions);
        if (!plan.why_not.empty())
            return -2;
        ExchangePlan xp;
        if (!build_exchange_plan(h, plan, regions, 1, xp))
            return -2;
        if (xp.max_local * rp.vertex_bytes > smem_bytes)
            return -3;
    }
    else
        build_cluster_plan(h, n_regions, per_body != 0, plan);
    if (!cluster_plan_is_valid(h, plan))
        return -1;
    std::memcpy(serial_order, plan.serial_order.data(), sizeof(uint32_t) * plan.serial_order.size());
    std::memcpy(storage_order, plan.storage_order.data(), sizeof(uint32_t) * plan.storage_order.size());
    std::memcpy(tet_region, plan.tet_region.data(), sizeof(int32_t) * plan.tet_region.size());
    *n_clusters = plan.n_clusters;
    *max_chunk  = plan.max_chunk_clusters;
    return plan.n_colours;
}

// decomposition over `world` ranks of a single lattice body: per-tet region and per-region rank, as
// the library plans it (regions_for / build_cluster_plan / region_rank); returns the region count
// bodies per region the planner picks for an ensemble of n_bodies copies of one mesh on sm_count SMs (ResidentParams::sm_count)
extern "C" int hs_ensemble_group(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_bodies, int sm_count,
                                 int32_t* n_regions, int32_t* threads)
{
    HostScene h;
    for (int b = 0; b < n_bodies; ++b)
    {
        HostBody hb;
        hb.v_offset   = h.n_vertices();
        hb.n_vertices = nV;
        hb.t_offset   = h.n_tets();
        hb.n_tets     = nT;
        for (int64_t i = 0; i < nV; ++i)
        {
            h.x0.push_back(x0[3 * i] + 1000.0 * b);
            h.x0.push_back(x0[3 * i + 1]);
            h.x0.push_back(x0[3 * i + 2]);
            h.mass.push_back(1.0);
        }
        for (int64_t t = 0; t < nT; ++t)
        {
            for (int a = 0; a < 4; ++a)
                h.tets.push_back(static_cast<uint32_t>(hb.v_offset + tets[4 * t + a]));
            h.tet_insertion.push_back(h.n_constraints++);
            h.tet_material.push_back(0);
        }
        h.bodies.push_back(hb);
    }
    ResidentParams rp;
    rp.smem_bytes  = 208 * 1024;
    rp.max_threads = 384;
    rp.sm_count    = sm_count;
    ClusterPlan cp;
    RegionPlan regions;
    build_cluster_plan(h, 0, true, cp, &rp, &regions);
    if (!cp.why_not.empty() || !cluster_plan_is_valid(h, cp))
        return -1;
    *n_regions = cp.n_regions;
    *threads   = cp.nt;
    return (n_bodies + cp.n_regions - 1) / cp.n_regions;
}

extern "C" int hs_partition(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int sm_count, int world,
                            int32_t* tet_region, int32_t* region_rank_out, int cap_regions)
{
    HostScene h;
    h.x0.assign(x0, x0 + 3 * nV);
    h.mass.assign(static_cast<size_t>(nV), 1.0);
    h.tets.assign(tets, tets + 4 * nT);
    for (int64_t t = 0; t < nT; ++t)
    {
        h.tet_insertion.push_back(h.n_constraints++);
        h.tet_material.push_back(0);
    }
    HostBody b;
    b.n_vertices = nV;
    b.n_tets     = nT;
    h.bodies.push_back(b);
    int32_t const n_regions = regions_for(sm_count, nT, world);
    if (n_regions > cap_regions || n_regions % world != 0)
        return -1;
    ResidentParams rp;
    rp.smem_bytes = 200 * 1024;
    ClusterPlan plan;
    RegionPlan regions;
    build_cluster_plan(h, n_regions, false, plan, &rp, &regions);
    if (!plan.why_not.empty() || !cluster_plan_is_valid(h, plan))
        return -2;
    ExchangePlan xp;
    if (!build_exchange_plan(h, plan, regions, world, xp))
        return -2;
    std::memcpy(tet_region, plan.tet_region.data(), sizeof(int32_t) * plan.tet_region.size());
    for (int32_t r = 0; r < n_regions; ++r)
        region_rank_out[r] = region_rank(r, n_regions, world);
    return n_regions;
}

// ---------------------------------------------------------------------------------------------------------
// Emulation of the resident schedule's exchange protocol (xpbd_resident.cuh) on the CPU, with an order-
// sensitive integer "projection": every region keeps its local vertex table, pulls and pushes exactly what
// the exchange plan says, phase by phase, and the result must equal the plain serial sweep in the exported
// order.  Checks on the way: a pull finds exactly the expected tag, a push never overwrites a record nobody
// consumed, a routing word names the rank that runs the reading region.
// `n_bodies` copies of the body (laid out along x); per_body: ensemble mode.  reverse: regions are visited
// in reverse order inside a phase (the protocol must not depend on it).
// Returns 0, or a negative code; stats = {regions, colours, exchange clusters, entries, shared vertices,
// pulls (later sweep, no collision steps), pushes (same), quiet steps, max_local, threads}.
// ---------------------------------------------------------------------------------------------------------
namespace {
inline uint64_t mix(uint64_t a, uint64_t b)
{
    uint64_t z = a * 0x9e3779b97f4a7c15ull + (b ^ (b >> 29)) + 0x7f4a7c15ull;
    z          = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z          = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
struct Box
{
    uint64_t value = 0;
    uint32_t tag   = 0;
    bool consumed  = true;
};
} // namespace

extern "C" int hs_exchange_emulate(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_bodies,
                                   int n_regions, int world, int per_body, int iterations, int collide, int reverse,
                                   int pencils, int64_t* stats)
{
    HostScene h;
    for (int b = 0; b < n_bodies; ++b)
    {
        HostBody hb;
        hb.v_offset   = h.n_vertices();
        hb.n_vertices = nV;
        hb.t_offset   = h.n_tets();
        hb.n_tets     = nT;
        for (int64_t i = 0; i < nV; ++i)
        {
            h.x0.push_back(x0[3 * i] + 1000.0 * b);
            h.x0.push_back(x0[3 * i + 1]);
            h.x0.push_back(x0[3 * i + 2]);
            h.mass.push_back(1.0);
        }
        for (int64_t t = 0; t < nT; ++t)
        {
            for (int a = 0; a < 4; ++a)
                h.tets.push_back(static_cast<uint32_t>(hb.v_offset + tets[4 * t + a]));
            h.tet_insertion.push_back(h.n_constraints++);
            h.tet_material.push_back(0);
        }
        extract_boundary(nV, nT, tets, hb.surf_to_tet, &hb.surf_triangles);
        h.bodies.push_back(hb);
    }
    ResidentParams rp;
    rp.smem_bytes = 200 * 1024;
    rp.pencils    = pencils != 0;
    ClusterPlan cp;
    RegionPlan regions;
    build_cluster_plan(h, n_regions, per_body != 0, cp, &rp, &regions);
    if (!cp.why_not.empty() || !cluster_plan_is_valid(h, cp))
        return -1;
    ExchangePlan xp;
    if (!build_exchange_plan(h, cp, regions, world, xp))
        return -2;
    int32_t const Rn = cp.n_regions, C = cp.n_colours, K = iterations, cs = collide ? 1 : 0;
    int64_t const V = h.n_vertices(), NX = xp.n_xclusters;
    if (Rn % world != 0)
        return -3;
    std::vector<char> surface(static_cast<size_t>(V), 0);
    for (HostBody const& hb : h.bodies)
        for (uint32_t lv : hb.surf_to_tet)
            surface[static_cast<size_t>(hb.v_offset + lv)] = 1;

    // ---- serial sweep in the exported order
    auto project = [](uint64_t* p[4], uint64_t t) {
        uint64_t const hsh = mix(mix(mix(*p[0], *p[1]), mix(*p[2], *p[3])), t);
        for (int a = 0; a < 4; ++a)
            *p[a] = mix(*p[a], hsh + static_cast<uint64_t>(a));
    };
    std::vector<uint64_t> ref(static_cast<size_t>(V));
    for (int64_t v = 0; v < V; ++v)
        ref[static_cast<size_t>(v)] = mix(static_cast<uint64_t>(v), 1);
    for (int k = 0; k < K; ++k)
    {
        if (cs)
            for (int64_t v = 0; v < V; ++v)
                if (surface[static_cast<size_t>(v)])
                    ref[static_cast<size_t>(v)] = mix(ref[static_cast<size_t>(v)], 77u + static_cast<uint64_t>(k));
        for (uint32_t t : cp.serial_order)
        {
            uint64_t* p[4];
            for (int a = 0; a < 4; ++a)
                p[a] = &ref[h.tets[4 * static_cast<size_t>(t) + a]];
            project(p, t);
        }
    }

    // ---- the regions, phase by phase
    uint32_t const base = 1;
    int32_t const per_iteration = C + cs, n_phases = 2 + K * per_iteration;
    std::vector<std::vector<Box>> box(static_cast<size_t>(world),
                                      std::vector<Box>(static_cast<size_t>(xp.n_entries) + static_cast<size_t>(xp.n_shared)));
    std::vector<std::vector<uint64_t>> sx(static_cast<size_t>(Rn));
    std::vector<uint64_t> result(static_cast<size_t>(V), 0);
    int error = 0;
    auto rank_of = [&](int32_t r) { return static_cast<uint32_t>(region_rank(r, Rn, world)); };
    auto push    = [&](int32_t /*from*/, uint32_t route, uint64_t value, uint32_t tag) {
        uint32_t const rank = route >> kRouteRankShift & 7u, index = route & kRouteIndexMask;
        if (rank >= static_cast<uint32_t>(world) || index >= box[rank].size())
        {
            error = -10;
            return;
        }
        Box& b = box[rank][index];
        if (!b.consumed)
            error = -11; // overwrites a record nobody read
        b = Box{value, tag, false};
    };
    auto pull = [&](int32_t region, uint32_t index, uint32_t expect) -> uint64_t {
        Box& b = box[rank_of(region)][index];
        if (b.tag != expect || b.consumed)
            error = -12; // the expected record is not there
        b.consumed = true;
        return b.value;
    };
    auto last_colour_tag = [&](int32_t k, uint32_t lastc) {
        return base + 1u + static_cast<uint32_t>(k * per_iteration + cs) + lastc;
    };
    for (int32_t r = 0; r < Rn; ++r)
        sx[static_cast<size_t>(r)].assign(static_cast<size_t>(xp.loc_off[r + 1] - xp.loc_off[r]), 0xdeadbeefdeadbeefull);
    for (int32_t p = 0; p < n_phases && !error; ++p)
    {
        uint32_t const tag = base + static_cast<uint32_t>(p);
        int32_t const k = p == 0 ? 0 : (p - 1) / per_iteration, q = p == 0 ? 0 : (p - 1) % per_iteration;
        for (int32_t rr = 0; rr < Rn && !error; ++rr)
        {
            int32_t const r = reverse ? Rn - 1 - rr : rr;
            std::vector<uint64_t>& s = sx[static_cast<size_t>(r)];
            uint32_t const* loc      = &xp.loc_vtx[static_cast<size_t>(xp.loc_off[r])];
            if (p == 0)
            { // predict
                for (int32_t i = 0; i < xp.n_owned[r]; ++i)
                    s[static_cast<size_t>(i)] = mix(loc[i], 1);
                for (int32_t i = xp.osv_off[r]; i < xp.osv_off[r + 1]; ++i)
                {
                    uint32_t const meta = xp.osv_meta[static_cast<size_t>(i)];
                    bool const to_me    = K == 0 || (cs && (meta & kOsvSurface));
                    if (!to_me && (meta & kOsvFirstRemote))
                        push(r, xp.osv_first[static_cast<size_t>(i)], s[xp.osv_slot[static_cast<size_t>(i)]], tag);
                }
            }
            else if (p == n_phases - 1)
            { // commit
                for (int32_t i = xp.osv_off[r]; i < xp.osv_off[r + 1]; ++i)
                {
                    uint32_t const meta = xp.osv_meta[static_cast<size_t>(i)];
                    if (K > 0 && (meta & kOsvLastRemote))
                        s[xp.osv_slot[static_cast<size_t>(i)]] =
                            pull(r, xp.n_entries + static_cast<uint32_t>(i), last_colour_tag(K - 1, meta & 0xffu));
                }
                for (int32_t i = 0; i < xp.n_owned[r]; ++i)
                    result[loc[i]] = s[static_cast<size_t>(i)];
            }
            else if (cs && q == 0)
            { // collision step: owned surface vertices
                for (int32_t i = xp.surf_off[r]; i < xp.surf_off[r + 1]; ++i)
                {
                    uint32_t const slot = xp.surf_slot[static_cast<size_t>(i)], o = xp.surf_osv[static_cast<size_t>(i)];
                    uint32_t const meta = o == kRouteNone ? 0u : xp.osv_meta[o];
                    if (o != kRouteNone && k > 0 && (meta & kOsvLastRemote))
                        s[slot] = pull(r, xp.n_entries + o, last_colour_tag(k - 1, meta & 0xffu));
                    if (loc[slot] >= static_cast<uint32_t>(V) || !surface[loc[slot]])
                        error = -13;
                    s[slot] = mix(s[slot], 77u + static_cast<uint64_t>(k));
                    if (o != kRouteNone && (meta & kOsvFirstRemote))
                        push(r, xp.osv_first[o], s[slot], tag);
                }
            }
            else
            { // colour step
                int32_t const c  = q - cs;
                int pv           = 2 * (k > 0 ? 1 : 0) + cs, qv = 2 * (k == K - 1 ? 1 : 0) + cs;
                size_t const E   = static_cast<size_t>(xp.entries);
                for (int part = 0; part < 2; ++part)
                {
                    ChunkDesc const& d = cp.chunks[(static_cast<size_t>(c) * Rn + r) * 2 + part];
                    for (int32_t i = 0; i < d.n[0] && !error; ++i)
                    {
                        int64_t const xq = part == 0 ? xp.chunk_xfirst[static_cast<size_t>(c) * Rn + r] + i : -1;
                        if (part == 0)
                            for (size_t e = 0; e < E; ++e)
                            {
                                uint32_t const w = xp.pull[((pv * (E / 4) + e / 4) * static_cast<size_t>(NX) + xq) * 4 + e % 4];
                                if (!(w & kPullValid))
                                    break;
                                uint32_t const dd = w >> 16 & 0xffu, entry = w >> 24 & 0xfu, slot = w & 0xffffu;
                                s[slot] = pull(r, static_cast<uint32_t>(entry * NX + xq), dd == kPullPredict ? base : tag - dd);
                            }
                        int64_t bse = d.first;
                        for (int m = 0; m < kMaxCluster && i < d.n[m]; ++m)
                        {
                            int64_t const pos = bse + i;
                            uint32_t const t  = cp.storage_order[static_cast<size_t>(pos)];
                            uint64_t* pp[4];
                            for (int a = 0; a < 4; ++a)
                            {
                                uint32_t const slot = xp.tet_slots[4 * static_cast<size_t>(pos) + a];
                                if (slot >= s.size() || loc[slot] != h.tets[4 * static_cast<size_t>(t) + a])
                                    error = -14;
                                pp[a] = &s[slot];
                            }
                            if (!error)
                                project(pp, t);
                            bse += d.n[m];
                        }
                        if (part == 0)
                            for (size_t e = 0; e < E; ++e)
                            {
                                size_t const at = ((qv * (E / 2) + e / 2) * static_cast<size_t>(NX) + xq) * 4 + 2 * (e % 2);
                                uint32_t const w = xp.push[at];
                                if (!(w & kPullValid))
                                    break;
                                push(r, xp.push[at + 1], s[w & 0xffffu], tag);
                            }
                    }
                }
            }
        }
    }
    if (error)
        return error;
    for (int64_t v = 0; v < V; ++v)
        if (result[static_cast<size_t>(v)] != ref[static_cast<size_t>(v)])
            return -20;
    if (stats)
    {
        stats[0] = Rn;
        stats[1] = C;
        stats[2] = NX;
        stats[3] = xp.entries;
        stats[4] = xp.n_shared;
        stats[5] = xp.n_pulls[2];
        stats[6] = xp.n_pushes[0];
        stats[7] = xp.quiet_steps;
        stats[8] = xp.max_local;
        stats[9] = cp.nt;
        for (int32_t c = 0; c < C && c < 13; ++c)
            stats[10 + c] = xp.pulls_by_colour[static_cast<size_t>(c)];
        if (getenv("HS_PRINT_HIST"))
        {
            printf("pulls per cluster:");
            for (int i = 0; i <= 16; ++i) printf(" %lld", static_cast<long long>(xp.pull_hist[i]));
            printf("\npushes per cluster:");
            for (int i = 0; i <= 16; ++i) printf(" %lld", static_cast<long long>(xp.push_hist[i]));
            printf("\n");
        }
        stats[23] = xp.bank_wavefronts_before;
        stats[24] = xp.bank_wavefronts;
        stats[25] = xp.bank_wavefronts_ideal;
    }
    return 0;
}
