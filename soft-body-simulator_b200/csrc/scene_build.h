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

// Layout parameters of the resident schedule; see xpbd_resident.cuh.
struct ResidentParams
{
    int64_t smem_bytes   = 0;  // shared memory available to one CTA for vertices
    int32_t vertex_bytes = 16; // sizeof(Real4<R>)
    int32_t max_threads  = 512;
    bool rotate_items    = true; // clusters that exchange vertices on the warps with a sub-partition to themselves
    bool pencils         = false; // regions = bundles of whole cluster columns along the shortest axis of the cluster
                                 // grid (compact in the two other axes) instead of compact Morton blocks: with the
                                 // colour order that goes with it (a Gray code of the cell parities on a lattice)
                                 // every step that flips the parity along the pencil axis depends on no other region
    bool push_first      = true; // clusters that push in their step first inside a chunk (A/B knob)
    int32_t bodies_per_region = 0; // ensembles (one_region_per_body): consecutive bodies grouped into one region so
                                   // that a colour step fills its warps; 0 = choose (about 160 clusters per step, or,
                                   // when sm_count is given, the size that leaves the SMs the least idle capacity)
    int32_t sm_count = 0;          // SMs the regions are dealt to (0: unknown)
};

// Which thread runs cluster i of a (colour, region) step.  Warp w of a CTA issues on sub-partition w % 4,
// so with W = nt / 32 warps the sub-partitions W % 4 .. 3 hold one warp less than the others.  The
// clusters that exchange vertices with other regions (part 0, first in a step) carry the polls and
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
    // [(colour * n_regions + region) * 2 + part]; part 0 = clusters with a vertex that another region touches
    // too (they exchange it through mailboxes), part 1 = clusters private to the region.  Without a region
    // plan everything is in part 1.
    std::vector<ChunkDesc> chunks;
    std::vector<int32_t> tet_region;     // T
    int64_t max_chunk_clusters = 0;      // max over (colour, region) of the clusters of both parts
    // resident schedule only
    int32_t nt  = 0;                     // threads per CTA
    int32_t rot = 0;                     // cluster i of a step (part 0 first) runs on thread (i + rot) % nt, see item_rotation
    std::string why_not;                 // non-empty when the resident layout could not be built
};

// n_regions <= 1 and no `resident`: no partition (graph schedule).  Otherwise clusters are dealt to
// regions with balanced tet counts — pencils (ResidentParams::pencils) or (body, Morton) order — or, for
// ensembles, whole bodies per region; with `resident` the vertex classification (RegionPlan), the parts and
// the launch shape are built too, and the colours are renumbered so that consecutive colours hand a shared
// vertex over inside a region as often as possible.
void build_cluster_plan(HostScene const& scene, int32_t n_regions, bool one_region_per_body,
                        ClusterPlan& out, ResidentParams const* resident = nullptr,
                        RegionPlan* region_plan = nullptr);

// ---- decomposition of one scene over several GPUs ---------------------------------------------
// How many regions to cut a scene into: one per SM of every rank, but at least ~64 clusters per
// colour step and region (small scenes use fewer SMs rather than synchronise regions of a few tets).
// Always a multiple of `world`: rank r runs the block of consecutive regions [r, r + 1) * n / world,
// which is spatially compact because regions follow a space-filling order of the clusters.
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

// ---- the resident schedule's exchange plan (xpbd_resident.cuh) ------------------------------------
// Every region keeps ONE shared-memory slot for every vertex its tets touch: the vertices it owns (it predicts,
// collides and commits them) first, then its guests (owned by another region).  A vertex that only one region
// touches never leaves that region's shared memory during a substep.  A SHARED vertex (touched by several
// regions) travels: whoever touches it knows, statically, who touched it before and who touches it next —
// touches are: predict (owner), per iteration [the collision step (owner, surface vertices, when collision steps
// exist)] and the colours of the clusters that contain it, commit (owner).  When the next touch is by another
// region the position is PUSHED into a mailbox of that region, tagged with the step; when the previous touch was
// by another region the position is PULLED (polled for the expected tag) out of the own mailbox into the slot;
// when both touches are by the same region nothing happens at all — the value stays in the slot.
//
// Mailboxes: one per (exchange cluster, entry) — entry = position of the shared vertex among the shared vertices
// of its cluster — at index entry * n_xclusters + xq, and one per shared vertex for its owner at n_entries + i.
// Routing words carry the mailbox index (bits 0-27) and the rank whose memory holds it (bits 28-30).
constexpr uint32_t kRouteNone       = 0xffffffffu;
constexpr uint32_t kRouteIndexMask  = 0x0fffffffu;
constexpr int kRouteRankShift       = 28;
// pull word: bit 31 valid, bits 24-27 entry, bits 16-23 steps back to the previous touch (0xff = the predict
// step), bits 0-15 slot.  push words: {bit 31 valid | slot, routing word}
constexpr uint32_t kPullValid      = 0x80000000u;
constexpr uint32_t kPullPredict    = 0xffu;
// owned shared vertex (ExchangePlan::osv), meta word
constexpr uint32_t kOsvSurface     = 0x100u; // takes part in the collision steps
constexpr uint32_t kOsvLastRemote  = 0x200u; // the last colour touching it in a sweep belongs to another region
constexpr uint32_t kOsvFirstRemote = 0x400u; // the first one does

struct ExchangePlan
{
    int32_t n_regions = 0, n_colours = 0, world = 1;
    // local vertex tables
    std::vector<int32_t> loc_off;      // [R + 1] into loc_vtx
    std::vector<int32_t> n_owned;      // [R] the first n_owned slots of a region are the vertices it owns
    std::vector<uint32_t> loc_vtx;     // global vertex of slot i
    std::vector<int32_t> vertex_owner; // [V]
    int64_t max_local = 0;             // largest table
    int64_t n_shared  = 0;             // vertices touched by more than one region
    std::vector<uint16_t> tet_slots;   // 4 * T (storage order): slot in the table of the tet's region
    // exchange clusters (part 0 of every chunk), numbered in storage order
    int64_t n_xclusters = 0;
    std::vector<int32_t> chunk_xfirst; // [colour * R + region]: number of the chunk's first exchange cluster
    int32_t entries = 0;               // shared vertices per cluster at most, rounded up to a multiple of 4
    // per variant, COMPACTED (valid words first, then zeros):
    //   pull variant = 2 * (iteration > 0) + (collision steps present)
    //   push variant = 2 * (last iteration) + (collision steps present)
    std::vector<uint32_t> pull;        // [variant][entries / 4][n_xclusters][4]: four pull words per 16-byte record
    std::vector<uint32_t> push;        // [variant][entries / 2][n_xclusters][4]: {slot word, route, slot word, route}
    uint32_t n_entries = 0;            // entries * n_xclusters: mailboxes of the clusters
    // owner side
    std::vector<int32_t> osv_off;      // [R + 1] shared vertices a region owns
    std::vector<uint32_t> osv_slot, osv_meta, osv_first; // slot; last colour | kOsv*; routing word of the first
                                                         // cluster entry touching it in a sweep
    std::vector<int32_t> surf_off;     // [R + 1] surface vertices a region owns
    std::vector<uint32_t> surf_slot, surf_index, surf_osv; // slot; surface-vertex index; position in osv or kRouteNone
    // statistics
    int64_t n_pulls[4] = {0, 0, 0, 0}, n_pushes[4] = {0, 0, 0, 0};
    std::vector<int64_t> pulls_by_colour; // [colour], variant: later sweep, no collision steps
    // shared-memory wavefronts of the vertex accesses of one sweep (one warp instruction = one group of slots):
    // with slots in vertex order, with the order build_exchange_plan chooses, and the lower bound (8 columns)
    int64_t bank_wavefronts_before = 0, bank_wavefronts = 0, bank_wavefronts_ideal = 0;
    int64_t pull_hist[17] = {}, push_hist[17] = {}; // exchange clusters by records pulled / pushed per step (variant: later
                                                    // sweep / not the last sweep, collision steps present)
    int32_t quiet_steps = 0;           // colours whose clusters pull nothing in any region (variant: later sweep, no
                                       // collision steps): steps that wait for no other region
    std::string why_not;
};

bool build_exchange_plan(HostScene const& scene, ClusterPlan const& cp, RegionPlan const& plan, int world,
                         ExchangePlan& out);

// true when no two tets of different clusters of one colour share a vertex
bool cluster_plan_is_valid(HostScene const& scene, ClusterPlan const& plan);

// vertex classification / neighbour lists for a given tet->region map (see RegionPlan)
// capacity: at most that many resident vertices per region (the rest stay in global memory)
void classify_regions(HostScene const& scene, std::vector<int32_t> const& tet_region, int32_t n_regions,
                      RegionPlan& plan, int64_t capacity = -1);

// Validation helper: true when no two constraints of the same colour share a vertex.
bool colouring_is_valid(int64_t n_vertices, int k, uint32_t const* verts, ColourClass const& cc);

} // namespace sbsb200
