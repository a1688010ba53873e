// TEST-ONLY harness: C entry points over the product's host-side scene compiler
// (soft-body-simulator_b200/csrc/scene_build.cpp) so that boundary extraction, colouring and
// the region plan can be tested on a CPU-only box.
#include "../soft-body-simulator_b200/csrc/scene_build.h"

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
    { // resident (persistent) layout: also check the shared-memory slots and fetch lists
        ResidentParams rp;
        rp.smem_bytes   = smem_bytes;
        rp.vertex_bytes = 16;
        rp.max_threads  = 512;
        RegionPlan regions;
        build_cluster_plan(h, n_regions, per_body != 0, plan, &rp, &regions);
        if (!plan.why_not.empty() || !resident_layout_is_valid(h, plan, regions))
            return -2;
        int64_t scratch = static_cast<int64_t>(plan.nvc) * plan.nt;
        if ((scratch + regions.max_region_vertices) * rp.vertex_bytes > smem_bytes)
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
    if (!plan.why_not.empty() || !cluster_plan_is_valid(h, plan) || !resident_layout_is_valid(h, plan, regions))
        return -2;
    std::memcpy(tet_region, plan.tet_region.data(), sizeof(int32_t) * plan.tet_region.size());
    for (int32_t r = 0; r < n_regions; ++r)
        region_rank_out[r] = region_rank(r, n_regions, world);
    return n_regions;
}

// mailbox routing of the resident schedule for a single body cut into `n_regions` regions run by
// `world` ranks.  Outputs (caller-sized with the capacities given): cl_fetch / to / to_owner
// [nvc * Q], cluster_region[Q], cluster_colour[Q], ifv / ifv_first [n_ifv], vertex_owner[V];
// dims = {nvc, Q, n_ifv, n_entries}.  Returns 0, -1 (no plan), -2 (no routes), -3 (capacity).
extern "C" int hs_mailbox_routes(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_regions,
                                 int world, int64_t cap_entries, int64_t cap_clusters, uint32_t* cl_fetch,
                                 uint32_t* to, uint32_t* to_owner, int32_t* cluster_region, int32_t* cluster_colour,
                                 uint32_t* ifv, uint32_t* ifv_first, int32_t* vertex_owner, int64_t* dims)
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
    extract_boundary(nV, nT, tets, b.surf_to_tet, &b.surf_triangles);
    h.bodies.push_back(b);
    ResidentParams rp;
    rp.smem_bytes = 200 * 1024;
    ClusterPlan plan;
    RegionPlan regions;
    build_cluster_plan(h, n_regions, false, plan, &rp, &regions);
    if (!plan.why_not.empty())
        return -1;
    MailboxRoutes routes;
    if (!build_mailbox_routes(h, plan, regions, plan.nvc, world, routes))
        return -2;
    int64_t const Q = plan.n_clusters;
    if (routes.n_entries > cap_entries || Q > cap_clusters || static_cast<int64_t>(routes.ifv.size()) > nV)
        return -3;
    std::memcpy(cl_fetch, plan.cl_fetch.data(), sizeof(uint32_t) * routes.n_entries);
    std::memcpy(to, routes.to.data(), sizeof(uint32_t) * routes.n_entries);
    std::memcpy(to_owner, routes.to_owner.data(), sizeof(uint32_t) * routes.n_entries);
    for (size_t ch = 0; ch < plan.chunks.size(); ++ch)
        for (int32_t i = 0; i < plan.chunks[ch].n[0]; ++i)
        {
            cluster_colour[plan.chunks[ch].cfirst + i] = static_cast<int32_t>(ch / (2 * static_cast<size_t>(n_regions)));
            cluster_region[plan.chunks[ch].cfirst + i] = static_cast<int32_t>((ch / 2) % static_cast<size_t>(n_regions));
        }
    std::memcpy(ifv, routes.ifv.data(), sizeof(uint32_t) * routes.ifv.size());
    std::memcpy(ifv_first, routes.ifv_first.data(), sizeof(uint32_t) * routes.ifv.size());
    std::memcpy(vertex_owner, regions.vertex_owner.data(), sizeof(int32_t) * static_cast<size_t>(nV));
    dims[0] = plan.nvc;
    dims[1] = Q;
    dims[2] = static_cast<int64_t>(routes.ifv.size());
    dims[3] = routes.n_entries;
    return 0;
}

// EXPERIMENTAL hand-off inside a region (ResidentParams::handoff, optionally with slab-shaped regions): plan a
// single lattice body, build the routes and check every local routing word: it must name the scratch slot
// (entry * nt + thread) of the cluster of the SAME region and the NEXT colour that fetches the same vertex, and
// exactly those entries must be marked "arrives by hand-off".  Returns the number of local hand-offs, or a
// negative error code; n_remote = touches that still go through a mailbox; steps_without_remote = colour steps
// (colour, region) none of whose clusters polls a mailbox filled by another region... counted as colours for
// which NO entry of ANY region has a remote previous touch inside the sweep.
extern "C" int64_t hs_handoff_check(int64_t nV, int64_t nT, const uint32_t* tets, const double* x0, int n_regions,
                                    int world, int slabs, int64_t* n_remote, int32_t* n_colours, int32_t* n_regions_out,
                                    int32_t* colours_without_remote)
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
    extract_boundary(nV, nT, tets, b.surf_to_tet, &b.surf_triangles);
    h.bodies.push_back(b);
    ResidentParams rp;
    rp.smem_bytes = 200 * 1024;
    rp.handoff    = true;
    rp.slabs      = slabs != 0;
    ClusterPlan plan;
    RegionPlan regions;
    build_cluster_plan(h, n_regions, false, plan, &rp, &regions);
    if (!plan.why_not.empty() || plan.banks != 2)
        return -1;
    if (!cluster_plan_is_valid(h, plan) || !resident_layout_is_valid(h, plan, regions))
        return -2;
    MailboxRoutes routes;
    if (!build_mailbox_routes(h, plan, regions, plan.nvc, world, routes))
        return -3;
    *n_colours     = plan.n_colours;
    *n_regions_out = plan.n_regions;
    int32_t const Rn = plan.n_regions;
    int64_t const Q = plan.n_clusters;
    std::vector<int32_t> colour(static_cast<size_t>(Q)), region(static_cast<size_t>(Q)), item(static_cast<size_t>(Q));
    for (size_t ch = 0; ch < plan.chunks.size(); ++ch)
        for (int32_t i = 0; i < plan.chunks[ch].n[0]; ++i)
        {
            size_t const q = static_cast<size_t>(plan.chunks[ch].cfirst + i);
            colour[q]      = static_cast<int32_t>(ch / (2 * static_cast<size_t>(Rn)));
            region[q]      = static_cast<int32_t>((ch / 2) % static_cast<size_t>(Rn));
            item[q]        = i;
        }
    // consumer lookup: (region, colour, entry, vertex) -> cluster
    int64_t local = 0, marked = 0, remote = 0;
    std::vector<char> colour_has_remote(static_cast<size_t>(plan.n_colours), 0);
    for (uint32_t box = 0; box < routes.n_entries; ++box)
    {
        marked += routes.local_prev[box];
        uint32_t const v = plan.cl_fetch[box];
        if (v == 0xffffffffu)
            continue;
        size_t const q = box % static_cast<size_t>(Q);
        if (!routes.local_prev[box] && colour[q] > 0)
            colour_has_remote[static_cast<size_t>(colour[q])] = 1; // polls a mailbox in the middle of a sweep
        uint32_t const word = routes.to[box];
        if (!(word & kRouteLocalBit))
        {
            ++remote;
            continue;
        }
        if ((routes.to_owner[box] & ~kRouteSurfaceBit) != word)
            return -5;
        ++local;
        uint32_t const idx = word & kRouteLocalIndexMask, thread = idx % plan.nt, entry = idx / plan.nt;
        if (entry >= static_cast<uint32_t>(plan.nvc))
            return -10;
        // the consumer runs on `thread` in the next colour step of the same region: its position in the step
        int32_t const want_item = static_cast<int32_t>((thread + plan.nt - plan.rot) % plan.nt);
        size_t const chA = (static_cast<size_t>(colour[q] + 1) * Rn + region[q]) * 2; // part A chunk of that step
        if (colour[q] + 1 >= plan.n_colours || want_item >= plan.chunks[chA].n[0])
            return -6;
        size_t const q2 = static_cast<size_t>(plan.chunks[chA].cfirst + want_item);
        if (plan.cl_fetch[static_cast<size_t>(entry) * Q + q2] != v)
            return -8;
        if (!routes.local_prev[static_cast<size_t>(entry) * Q + q2])
            return -7;
    }
    if (marked != local)
        return -9;
    *n_remote = remote;
    int32_t free_colours = 0;
    for (int32_t c = 1; c < plan.n_colours; ++c)
        free_colours += !colour_has_remote[static_cast<size_t>(c)];
    *colours_without_remote = free_colours;
    return local;
}
