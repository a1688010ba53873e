// Host-side scene compilation for the B200 XPBD path: boundary extraction, graph colouring,
// spatial ordering / region partition.  Pure C++ (no CUDA), so it is unit-testable on CPU.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace sbsb200 {

struct Material
{
    double mu, lambda, alpha, beta; // green_constraint.cpp:45-46 + constraint_t alpha/beta
    bool operator==(Material const& o) const
    {
        return mu == o.mu && lambda == o.lambda && alpha == o.alpha && beta == o.beta;
    }
};

enum class BodyKind : int { tet = 0, sdf = 1 };
enum class SdfKind : int { plane = 0, sphere = 1, box = 2, grid = 3 };

struct HostBody
{
    BodyKind kind = BodyKind::tet;
    bool collideable = true;                // handed to the cd system (sbsb200_set_body_collideable)
    // tet body
    int64_t v_offset = 0, n_vertices = 0;
    int64_t t_offset = 0, n_tets = 0;       // range in HostScene::tets (insertion order)
    int64_t s_offset = 0;                   // range in the global surface-vertex list
    std::vector<uint32_t> surf_to_tet;      // body-local tet-mesh vertex of surface vertex i
    std::vector<uint32_t> surf_triangles;   // boundary triangles as surface-vertex indices (3 per triangle)
    // sdf body
    SdfKind sdf_kind = SdfKind::plane;
    double a[3] = {0, 0, 0}, b[3] = {0, 0, 0}, r = 0; // plane: a=n, r=offset; sphere: a=c, r; box: a=min, b=max
    double volume[6] = {0, 0, 0, 0, 0, 0};
    // grid sdf (a..b = domain): cells per axis and node values in the order of grid_sdf.cuh
    uint32_t grid_n[3] = {0, 0, 0};
    std::vector<double> grid_nodes;
};

struct HostScene
{
    std::vector<HostBody> bodies;
    std::vector<double> x0;   // 3*V rest positions (global vertex order = bodies concatenated)
    std::vector<double> mass; // V
    // green constraints, insertion order
    std::vector<uint32_t> tets;          // 4*T global vertex ids
    std::vector<uint32_t> tet_insertion; // T: index in simulation_t::constraints_
    std::vector<uint16_t> tet_material;  // T
    std::vector<Material> materials;
    // distance constraints, insertion order
    std::vector<uint32_t> dist_pairs;     // 2*D global vertex ids
    std::vector<double> dist_rest;        // D
    std::vector<double> dist_alpha, dist_beta;
    std::vector<uint32_t> dist_insertion; // D
    uint32_t n_constraints = 0;           // running insertion counter

    int64_t n_vertices() const { return static_cast<int64_t>(mass.size()); }
    int64_t n_tets() const { return static_cast<int64_t>(tet_insertion.size()); }
    int64_t n_dist() const { return static_cast<int64_t>(dist_insertion.size()); }
};

// Surface of a tet mesh in the reference's numbering: boundary triangles are those with a
// number of incident tets != 2, visited in first-insertion order of the triangle; surface
// vertices are numbered in first-seen order (tetrahedral_mesh_boundary.cpp:84-119).
void extract_boundary(int64_t n_vertices, int64_t n_tets, uint32_t const* tets,
                      std::vector<uint32_t>& surf_to_tet, std::vector<uint32_t>* triangles);

struct ColourClass
{
    int32_t n_colours = 0;
    std::vector<uint32_t> order;   // sorted position -> constraint index (insertion-order index within its type)
    std::vector<int64_t> offsets;  // n_colours+1 prefix offsets into `order`
};

// Vertex-disjoint greedy colouring of k-vertex constraints (k = 4 tets, k = 2 edges), then a
// stable sort by (colour, region, spatial key).  region may be null.
// Returns false if more than max_colours would be needed.
bool colour_constraints(int64_t n_vertices, int64_t n, int k, uint32_t const* verts,
                        uint64_t const* spatial_key, int32_t const* region, int max_colours,
                        ColourClass& out);

// 30-bit Morton key per constraint from the centroid of its rest positions.
void morton_keys(int64_t n, int k, uint32_t const* verts, double const* x0, int64_t n_vertices,
                 std::vector<uint64_t>& keys);

// Region partition for the persistent schedule.  Tets are split into `n_regions` spatially
// compact, equal-count parts (contiguous ranges of the Morton order).  A vertex is interior
// to region r when every incident tet (and no other constraint) lies in r; otherwise it is
// an interface vertex and lives in global memory.
struct RegionPlan
{
    int32_t n_regions = 0;
    std::vector<int32_t> tet_region;        // T (insertion order)
    std::vector<int32_t> vertex_region;     // V: owning region of interior vertex, or -1 (interface / untouched)
    std::vector<int32_t> vertex_owner;      // V: region that predicts/commits the vertex (== vertex_region when interior)
    std::vector<uint32_t> vertex_slot;      // V: slot in the owning region's shared-memory array
    std::vector<int64_t> region_vtx_offsets;// R+1 into region_vtx
    std::vector<uint32_t> region_vtx;       // interior vertices grouped by region (slot order)
    std::vector<int32_t> nbr_offsets;       // R+1 into nbr
    std::vector<int32_t> nbr;               // neighbour regions (share >= 1 interface vertex)
    int64_t n_interface = 0;
    int64_t max_region_vertices = 0;
};

// one_region_per_body: ensembles — every tet body becomes its own region (no interface at all).
void plan_regions(HostScene const& scene, std::vector<uint64_t> const& tet_keys, int32_t n_regions,
                  RegionPlan& plan, bool one_region_per_body = false);

// ---------------------------------------------------------------------------------------------
// Clustered colouring of the Green constraints.
//
// Colouring single tets needs >= max-valence colours (32 on the 5-tets-per-cell lattice, 37-40
// with greedy), i.e. ~40 dependent passes per Gauss-Seidel sweep, each only T/40 wide.  Instead
// tets are grouped into small spatial CLUSTERS (the tets whose rest centroid falls in one cell
// of a uniform grid sized for ~5 tets; at most kMaxCluster).  One thread projects the tets of a
// cluster one after the other; clusters are coloured so that clusters of one colour share no
// vertex.  A sweep is then ~8 passes (exactly 8 on the lattice: the 2x2x2 parity pattern falls
// out of first-fit in raster order), each pass as wide as before, with 5x fewer barriers.
// The equivalent serial Gauss-Seidel order is: colour, cluster, tet-in-cluster.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxCluster = 8;
constexpr int kMaxClusterVertices = 16; // scratch entries per thread in the resident schedule

struct ChunkDesc
{
    int32_t first;          // storage index of the chunk's first tet
    int32_t cfirst;         // storage index of the chunk's first cluster
    int32_t n[kMaxCluster]; // n[j] = number of clusters in the chunk with more than j tets
};                          // tet j of cluster i (clusters sorted by size, descending) is stored at
                            // first + n[0] + ... + n[j-1] + i  ("column" layout: coalesced per j)

// Layout parameters of the resident (persistent) schedule; see xpbd_persistent.cuh.
struct ResidentParams
{
    int64_t smem_bytes   = 0; // shared memory available to one CTA
    int32_t vertex_bytes = 16; // sizeof(Real4<R>)
    int32_t max_threads  = 512;
    // EXPERIMENTAL, off by default (development knobs SBSB200_HANDOFF / SBSB200_SLABS):
    bool handoff         = false; // shared vertices go from one cluster to the next through shared memory when both
                                  // belong to the same region and run in consecutive steps (two banks of scratch
                                  // slots), and the colours are ordered so that most consecutive touches do
    bool slabs           = false; // regions = layers of clusters along the longest axis of the cluster grid (at most
                                  // one region per layer) instead of compact Morton blocks: with the hand-off and the
                                  // colour order that goes with it, most steps then depend on no other region at all
    bool rotate_items    = true;  // part A clusters on the warps with a sub-partition to themselves (item_rotation)
};

// Which thread runs cluster i of a (colour, region) step.  Warp w of a CTA issues on sub-partition w % 4,
// so with W = nt / 32 warps the sub-partitions W % 4 .. 3 hold one warp less than the others.  The
// clusters that exchange vertices with other regions (part A, first in a step) carry the polls and
// pushes on top of the projections: they go to the warps that have a sub-partition to themselves,
// i.e. the numbering of a step's clusters starts at warp W % 4 (measured: profiles/r01_summary.md).
inline int32_t item_rotation(int32_t nt)
{
    int32_t const warps = nt / 32;
    return nt > 0 && nt % 32 == 0 ? (32 * (warps % 4)) % nt : 0;
}

struct ClusterPlan
{
    int32_t n_colours = 0;
    int32_t n_regions = 1;
    int64_t n_clusters = 0;
    std::vector<uint32_t> storage_order; // storage position -> tet (index into HostScene::tets / 4)
    std::vector<uint32_t> serial_order;  // equivalent serial order of tets (colour, region, part, cluster, tet)
    // [(colour * n_regions + region) * 2 + part]; part 0 = clusters that fetch vertices from global
    // memory (shared with other regions, or not resident), part 1 = clusters whose vertices are all
    // resident in the region's shared memory.  Without a region plan everything is in part 1.
    std::vector<ChunkDesc> chunks;
    std::vector<int32_t> tet_region;     // T
    int64_t max_chunk_clusters = 0;      // max over (colour, region) of the clusters of both parts
    // resident schedule only
    int32_t nt  = 0;                     // threads per CTA the scratch slots were laid out for
    int32_t rot = 0;                     // cluster i of a step (part A first) runs on thread (i + rot) % nt, see item_rotation
    int32_t banks = 1;                   // banks of scratch slots: 2 = the cluster of step p uses bank p % 2, so that a
                                         // cluster may WRITE a shared vertex straight into the scratch slot of the
                                         // cluster of the same region that touches it in the next step (hand-off);
                                         // tet slots name bank 0, resident vertices start at banks * nvc * nt
    int32_t nvc = 0;                     // scratch entries per thread (multiple of 4, <= kMaxClusterVertices)
    std::vector<uint16_t> tet_slots;     // 4*T (storage order): index into the CTA's shared vertex array
    std::vector<uint32_t> cl_fetch;      // [nvc][n_clusters]: global vertex of scratch entry k, 0xffffffff = none
    // touch schedule for the tag protocol: per vertex  bits 0-7 = last colour touching it (0xff none),
    // bit 8 = surface vertex; per fetch entry four bytes "steps back to the previous touch" for
    // the four kinds of colour step (byte 2*(iteration > 0) + collision steps present), 0xff = the
    // predict step
    std::vector<uint32_t> vertex_meta;   // V
    std::vector<uint32_t> cl_meta;       // [nvc][n_clusters]
    std::string why_not;                 // non-empty when the resident layout could not be built
};

// n_regions <= 1 and no `resident`: no partition (graph schedule).  Otherwise clusters are dealt to
// regions in (body, Morton) order with balanced tet counts, or one region per body; with
// `resident` the vertex classification (RegionPlan) and the shared-memory layout are built too.
void build_cluster_plan(HostScene const& scene, int32_t n_regions, bool one_region_per_body,
                        ClusterPlan& out, ResidentParams const* resident = nullptr,
                        RegionPlan* region_plan = nullptr);

// ---- decomposition of one scene over several GPUs ---------------------------------------------
// How many regions to cut a scene into: one per SM of every rank, but at least ~64 clusters per
// colour step and region (small scenes use fewer SMs rather than synchronise regions of a few tets).
// Always a multiple of `world`: rank r runs the block of consecutive regions [r, r + 1) * n / world,
// which is spatially compact because regions follow the Morton order of the clusters.
inline int32_t regions_for(int sm_count, int64_t n_tets, int world = 1)
{
    int64_t per_rank = n_tets / (world > 0 ? world : 1) / 2560;
    per_rank         = per_rank < 1 ? 1 : per_rank > sm_count ? sm_count : per_rank;
    return static_cast<int32_t>(per_rank * (world > 0 ? world : 1));
}
inline int32_t region_rank(int32_t region, int32_t n_regions, int32_t world)
{
    return region / (n_regions / world);
}

// ---- mailboxes of the resident schedule (xpbd_persistent.cuh) -----------------------------------
// Entry (j, q) = scratch slot j of cluster q has mailbox j * Q + q (Q = number of clusters); owned
// non-resident vertex i (position in `ifv`) has mailbox n_entries + i.  Per vertex the entries are
// ordered by the colour of their cluster: each pushes to the next one; the last one pushes to the
// first one (next sweep) or to the owner (collision step, commit).  Routing words carry the mailbox
// index (bits 0-27) and the rank whose memory holds it (bits 28-30); bit 31 of `to_owner` marks
// surface vertices.
constexpr uint32_t kRouteNone       = 0xffffffffu;
constexpr uint32_t kRouteIndexMask  = 0x0fffffffu;
constexpr uint32_t kRouteSurfaceBit = 0x80000000u;
constexpr int kRouteRankShift       = 28;
// hand-off inside a region (ClusterPlan::banks == 2): bit 27 set, bits 0-26 = scratch slot (entry * nt + thread)
// of the cluster that touches the vertex in the next step; mailbox indices then stay below 2^27
constexpr uint32_t kRouteLocalBit       = 0x08000000u;
constexpr uint32_t kRouteLocalIndexMask = 0x07ffffffu;
constexpr uint32_t kMetaLocal           = 0xfefefefeu; // cl_meta of an entry that arrives by hand-off: nothing to poll

struct MailboxRoutes
{
    uint32_t n_entries = 0;             // nvc * Q
    std::vector<int32_t> ifv_offsets;   // [n_regions + 1] into ifv: vertices a region owns that are not resident
    std::vector<uint32_t> ifv;          // global vertex ids
    std::vector<uint32_t> ifv_meta;     // ClusterPlan::vertex_meta of those
    std::vector<uint32_t> ifv_first;    // routing word of the first entry touching the vertex in a sweep
    std::vector<uint32_t> ifv_pos;      // V: position in ifv, kRouteNone for resident vertices
    std::vector<uint32_t> to, to_owner; // [nvc][Q] routing words, kRouteNone for unused slots
    std::vector<uint8_t> local_prev;    // [nvc][Q] 1: the entry is handed over in shared memory by the previous touch
    int64_t n_local = 0;                // how many
    std::string why_not;
};

// nvc: scratch entries per thread of the kernel instantiation (>= cp.nvc); world: ranks the scene is cut over
bool build_mailbox_routes(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& plan, int nvc, int world,
                          MailboxRoutes& out);

// true when every tet slot of the resident layout resolves to the right vertex (resident slot of
// its region, or the scratch entry the running thread fetched) and parts are classified correctly
bool resident_layout_is_valid(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& rp);

// true when no two tets of different clusters of one colour share a vertex
bool cluster_plan_is_valid(HostScene const& scene, ClusterPlan const& plan);

// vertex classification / neighbour lists for a given tet->region map (see RegionPlan)
// capacity: at most that many resident vertices per region (the rest stay in global memory)
void classify_regions(HostScene const& scene, std::vector<int32_t> const& tet_region, int32_t n_regions,
                      RegionPlan& plan, int64_t capacity = -1);

// Validation helper: true when no two constraints of the same colour share a vertex.
bool colouring_is_valid(int64_t n_vertices, int k, uint32_t const* verts, ColourClass const& cc);

} // namespace sbsb200
