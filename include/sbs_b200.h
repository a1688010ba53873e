/*
 * sbs_b200.h — C ABI of the B200-native XPBD hot path (libsbsb200.so).
 *
 * The reference ("sbs", Q-Minh/soft-body-simulator) has no FFI: its plug-in points are C++
 * virtual interfaces and the predict/commit loops are inline in timestep_t::step, so the
 * replaceable unit is "a simulation_t plus timestep_t::step" (SURVEY.md §8b).  Each entry
 * point below names the reference interface it replaces (paths relative to the reference
 * root).  The C++ facade in soft-body-simulator_b200/cpp/sbs/ re-creates the reference's
 * class names on top of these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 (SBSB200_OK) or a negative error code, except the add_*
 *     functions, which return the new body index (>= 0) or a negative error code;
 *   - sbsb200_last_error() gives a message for the most recent failure on that context
 *     (pass NULL for failures of sbsb200_create);
 *   - host arrays are caller-owned and copied during the call; the context owns all device
 *     memory; no device pointer escapes;
 *   - one context per GPU; a context is not thread-safe;
 *   - there is NO CPU fallback: every call fails with SBSB200_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef SBS_B200_H
#define SBS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBSB200_OK 0
#define SBSB200_ERR_INVALID (-1)  /* bad argument / wrong call order */
#define SBSB200_ERR_CUDA (-2)     /* CUDA runtime failure, message in last_error */
#define SBSB200_ERR_STATE (-3)    /* e.g. step before finalize */
#define SBSB200_ERR_CAPACITY (-4) /* more colours / contacts than supported */

/* arithmetic of the device path: the production build is fp32, the validation build fp64
 * (include/sbs/aliases.h:9 uses double throughout). */
#define SBSB200_FP32 32
#define SBSB200_FP64 64

/* detect_mode of sbsb200_step */
#define SBSB200_DETECT_PER_FRAME 0   /* reference semantics: timestep.cpp:29-30 */
#define SBSB200_DETECT_PER_SUBSTEP 1 /* == `substeps` calls of step() with substeps=1 */

/* broadphase of the collision detection (sbsb200_set_broadphase) */
#define SBSB200_BROADPHASE_NONE 0 /* every surface vertex against every SDF (default: cheapest for analytic SDFs) */
#define SBSB200_BROADPHASE_BVH 1  /* linear BVH of bounding spheres per body, rebuilt on the device at every
                                   * detection, culled like point_bvh_model_t::collide (bvh_model.cpp:30-100) */

/* schedule selection (sbsb200_set_schedule) */
#define SBSB200_SCHED_AUTO 0
#define SBSB200_SCHED_GRAPH 1      /* one kernel per colour, whole frame in a CUDA graph */
#define SBSB200_SCHED_PERSISTENT 2 /* resident schedule: one kernel per substep, every region's vertices in shared memory */

/* region shape of the resident schedule (sbsb200_set_region_shape) */
#define SBSB200_REGIONS_PENCILS 0 /* bundles of cell columns along the shortest axis: half of the colour steps of a
                                   * sweep on a lattice then exchange nothing between regions */
#define SBSB200_REGIONS_COMPACT 1 /* compact blocks in Morton order (default: fewest shared vertices; measured faster
                                   * on B200 for the 1M-tet block, profiles/r02_summary.md) */

typedef struct sbsb200_ctx sbsb200_ctx;

typedef struct sbsb200_stats
{
    int32_t n_bodies;
    int32_t n_sdfs;
    int64_t n_vertices;
    int64_t n_tets;
    int64_t n_distance;
    int64_t n_surface_vertices;
    int32_t n_green_colours;
    int32_t n_distance_colours;
    int32_t schedule;           /* SBSB200_SCHED_* actually in use */
    int32_t n_regions;          /* persistent schedule: regions (= CTAs) */
    int64_t n_interface_vertices;
    int64_t kernels_launched;   /* kernels of this library launched since create */
    int64_t frames;             /* sbsb200_step calls since create */
    int64_t last_contact_count; /* contacts found by the most recent detection */
    double last_step_ms;        /* device time of the most recent sbsb200_step (CUDA events) */
    double kernel_ms;           /* persistent schedule: summed device time of the substep-kernel launches */
    int64_t kernel_launches;    /* ... and how many launches that sum covers (CUDA events around each) */
    int64_t n_shared_vertices;  /* resident schedule: vertices touched by more than one region */
    int64_t pulls_per_sweep;    /* ... mailbox pulls (= pushes) of one Gauss-Seidel sweep over all regions */
    int64_t pushes_per_sweep;
    int32_t quiet_colours;      /* ... colour steps of a sweep that (practically) wait for no other region */
    int32_t reserved_;
    int64_t green_general_calls; /* debug counter: projections that took the clamp / inversion route
                                  * (green_constraint.cpp:61-65, :91-102) on this device since finalize */
} sbsb200_stats;

/* ---- lifetime -------------------------------------------------------------------------- */

/* simulation_t{} (include/sbs/physics/simulation.h:15-45) bound to one GPU. */
int sbsb200_create(int device, int precision, sbsb200_ctx** out);
void sbsb200_destroy(sbsb200_ctx* ctx);
const char* sbsb200_last_error(const sbsb200_ctx* ctx);

/* Run on a caller-provided cudaStream_t (e.g. the framework's current stream); default is a
 * stream owned by the context.  Must be called before finalize. */
int sbsb200_set_stream(sbsb200_ctx* ctx, void* cuda_stream);
int sbsb200_set_schedule(sbsb200_ctx* ctx, int schedule);
/* How the resident schedule cuts a connected mesh into regions (a tuning knob; any shape gives a valid
 * Gauss-Seidel order, exported by sbsb200_get_constraint_order).  Before finalize.  No reference counterpart. */
int sbsb200_set_region_shape(sbsb200_ctx* ctx, int shape);

/* point_bvh_model_t (include/sbs/physics/collision/bvh_model.h:23-49): the broadphase.  Before finalize. */
int sbsb200_set_broadphase(sbsb200_ctx* ctx, int mode);

/* simulation_parameters_t::collision_compliance (include/sbs/physics/xpbd/simulation_parameters.h:24) */
int sbsb200_set_collision_compliance(sbsb200_ctx* ctx, double alpha);

/* ---- scene description (before finalize) ----------------------------------------------- */

/* tetrahedral_body_t(simulation, id, geometry) (src/physics/tetrahedral_body.cpp:29-83) plus
 * one green_constraint_t(alpha, beta, sim, body, v1..v4, E, nu) per tet in tet order
 * (src/physics/xpbd/green_constraint.cpp:11-47, main.cpp:37-54).
 * x0: 3*nV doubles; mass: nV doubles or NULL (= 1, particle.cpp:7; 0 pins the vertex,
 * particle.cpp:39-49); tets: 4*nT vertex indices local to the body. */
int sbsb200_add_tet_body(sbsb200_ctx* ctx, int64_t nV, const double* x0, const double* mass,
                         int64_t nT, const uint32_t* tets, double young_modulus,
                         double poisson_ratio, double alpha, double beta);

/* distance_constraint_t(alpha, beta, sim, b1, b2, v1, v2) (xpbd/distance_constraint.cpp:8-22),
 * n constraints, pairs = 2*n body-local vertex indices (v1 in b1, v2 in b2). */
int sbsb200_add_distance_constraints(sbsb200_ctx* ctx, int b1, int b2, int64_t n,
                                     const uint32_t* pairs, double alpha, double beta);

/* environment_body_t(sim, id, geometry, sdf_model_t) (src/physics/environment_body.cpp:80-88)
 * with sdf_model_t::from_plane (src/physics/collision/sdf_model.cpp:52-64) or an analytic
 * sdf_model_t (sdf_model.h:18-23).  volume = englobing AABB {min xyz, max xyz}
 * (collision_model.h:39-40).  Each takes one body slot, like simulation_t::add_body. */
int sbsb200_add_sdf_plane(sbsb200_ctx* ctx, const double normal[3], const double point[3],
                          const double volume[6]);
int sbsb200_add_sdf_sphere(sbsb200_ctx* ctx, const double centre[3], double radius,
                           const double volume[6]);
int sbsb200_add_sdf_box(sbsb200_ctx* ctx, const double box_min[3], const double box_max[3],
                        const double volume[6]);

/* brute_force_cd_system_t(objects) (src/physics/collision/brute_force_cd_system.cpp:8-12; main.cpp:77-86
 * builds the list): only the collision models handed to the cd system take part in detection.  Every
 * body does by default; flag = 0 takes a tet body's surface vertices, or an sdf body, out of detection
 * (a scene file's `"collideable": false`, src/io/load_scene.cpp:248-257).  Before sbsb200_finalize. */
int sbsb200_set_body_collideable(sbsb200_ctx* ctx, int body, int flag);

/* environment_body_t holding a discrete-grid sdf_model_t (sdf_model.cpp:18, evaluated by
 * Discregrid::CubicLagrangeDiscreteGrid::interpolate, sdf_model.cpp:71-74): 32-node cubic cells over
 * [domain_min, domain_max] with resolution[3] cells per axis.  node_values: one double per node in
 * Discregrid's node order (corners x-fastest, then two nodes per x-, y-, z-edge; see
 * csrc/grid_sdf.cuh); n_nodes must equal sbsb200_grid_node_count(resolution).  volume may be NULL:
 * the englobing volume is then the domain (environment_body.cpp:76). */
int64_t sbsb200_grid_node_count(const uint32_t resolution[3]);
int sbsb200_grid_node_position(const double domain_min[3], const double domain_max[3],
                               const uint32_t resolution[3], int64_t node, double position[3]);
int sbsb200_add_sdf_grid(sbsb200_ctx* ctx, const double domain_min[3], const double domain_max[3],
                         const uint32_t resolution[3], const double* node_values, int64_t n_nodes,
                         const double volume[6]);

/* environment_body_t(sim, id, geometry, domain, resolution) (environment_body.cpp:12-78): the domain
 * is extended to the triangle mesh and inflated as the reference does (:52-65), the signed distance
 * to the mesh (Discregrid::MeshDistance: closest triangle, sign from the angle-weighted pseudo-normal
 * of the closest feature) is sampled at every grid node ON THE DEVICE, and the body collides through
 * the resulting grid.  positions: 3*nV doubles; triangles: 3*nF vertex indices of a closed,
 * consistently oriented mesh; domain = {min xyz, max xyz}; resolution NULL = {10, 10, 10}
 * (environment_body.h:24). */
int sbsb200_add_sdf_mesh(sbsb200_ctx* ctx, int64_t nV, const double* positions, int64_t nF,
                         const uint32_t* triangles, const double domain[6], const uint32_t resolution[3]);
/* The extended domain alone (environment_body.cpp:52-65; host arithmetic, no context needed). */
int sbsb200_mesh_sdf_domain(int64_t nV, const double* positions, const double domain[6],
                            double extended[6]);

/* The grid of an sdf body added by sbsb200_add_sdf_grid / _mesh: domain {min xyz, max xyz},
 * resolution, node values (capacity cap; pass NULL to query).  Returns the node count, < 0 on error. */
int64_t sbsb200_get_sdf_grid(sbsb200_ctx* ctx, int body, double domain[6], uint32_t resolution[3],
                             double* node_values, int64_t cap);

/* sdf_model_t::evaluate (sdf_model.cpp:66-75) of an sdf body at n points, on the device, in the
 * context's precision: signed distance (DBL_MAX outside a grid's domain) and gradient (3*n).
 * Requires sbsb200_finalize. */
int sbsb200_eval_sdf(sbsb200_ctx* ctx, int body, int64_t n, const double* points, double* distance,
                     double* gradient);

/* Boundary extraction, graph colouring, region partition, SoA upload, BVH build.  Replaces
 * the incremental topology build (src/physics/topology.cpp:904-956) and
 * tetrahedral_mesh_boundary_t::extract_boundary_surface
 * (src/physics/tetrahedral_mesh_boundary.cpp:65-120). */
int sbsb200_finalize(sbsb200_ctx* ctx);

/* ---- introspection (after finalize) ---------------------------------------------------- */

/* Number of elastic constraints (green + distance), i.e. simulation_t::constraints().size(). */
int64_t sbsb200_constraint_count(const sbsb200_ctx* ctx);

/* simulation_t::remove_constraint (src/physics/simulation.cpp:34-39) for tetrahedron (Green) constraints, without
 * re-planning the scene: `constraints` are insertion indices as counted at finalize (they do NOT shift when other
 * constraints go, unlike the reference's swap-with-last positions; the C++ facade keeps the correspondence).  The
 * removed tets keep their place in the schedule with rest volume zero, so the projection leaves them alone
 * (gradient guard, green_constraint.cpp:67, :130-131); colours, regions and mailboxes stay as they are, the mesh
 * boundary too (the reference does not touch the mesh either).  sbsb200_constraint_count and
 * sbsb200_get_constraint_order afterwards describe the remaining constraints.  Distance constraints and constraint
 * insertion need a new scene (SBSB200_ERR_INVALID here).  One small copy, one kernel, one synchronisation. */
int sbsb200_remove_constraints(sbsb200_ctx* ctx, int64_t n, const uint32_t* constraints);

/* The serial Gauss-Seidel order equivalent to the GPU schedule: order[i] = insertion index
 * (position in simulation_t::constraints_, src/physics/simulation.cpp:29-32) of the i-th
 * projected constraint.  Feeding this permutation to the reference/oracle reproduces the
 * GPU result ("the reference run with constraints permuted into the same colour order"). */
int sbsb200_get_constraint_order(const sbsb200_ctx* ctx, uint32_t* order, int64_t n);

/* tetrahedral_mesh_boundary_t::surface_to_tetrahedral_mesh_index_map()
 * (tetrahedral_mesh_boundary.cpp:49-58).  Pass map=NULL to query the count. */
int64_t sbsb200_get_surface_map(const sbsb200_ctx* ctx, int body, uint32_t* map, int64_t cap);

/* tetrahedral_mesh_boundary_t::triangles() (tetrahedral_mesh_boundary.cpp:65-120): boundary triangles as
 * surface-vertex indices, 3 per triangle, in the reference's order.  Pass NULL to query the index count. */
int64_t sbsb200_get_surface_triangles(const sbsb200_ctx* ctx, int body, uint32_t* triangles, int64_t cap);

int sbsb200_get_stats(const sbsb200_ctx* ctx, sbsb200_stats* out);

/* Why the schedule in use differs from the requested one ("" when it does not). */
const char* sbsb200_schedule_note(const sbsb200_ctx* ctx);

/* ---- one scene decomposed over several GPUs (one context per GPU, any mix of processes) ---------
 * Every rank describes the SAME scene, calls set_partition(rank, world) before finalize, and then
 * runs its own block of regions; vertices shared between regions travel through mailboxes that
 * live on the reading rank, so a push to another rank is a peer store over NVLink issued by the
 * substep kernel itself (no host-side exchange, no collective).  After finalize every rank maps
 * the other ranks' mailbox arrays: across processes through CUDA IPC handles (64 bytes each,
 * exchanged by any means — bench.py uses torch.distributed.all_gather), inside one process
 * through connect_peer_context.  The reference has no counterpart (it is single-threaded).
 * The substep kernels of the ranks wait for each other's pushes: in a single process enqueue sbsb200_step on EVERY
 * rank before any call that blocks on one of them (synchronize, download, get_contacts, get_stats); a rank that
 * waits alone runs out of its poll budget, and the blocking call reports SBSB200_ERR_CUDA (the flag is cleared, the
 * state of that frame is lost). */
int sbsb200_set_partition(sbsb200_ctx* ctx, int rank, int world);
int sbsb200_get_mailbox_handle(sbsb200_ctx* ctx, void* handle64);
int sbsb200_connect_peers(sbsb200_ctx* ctx, const void* handles /* world x 64 bytes */, int world);
int sbsb200_connect_peer_context(sbsb200_ctx* ctx, int peer_rank, sbsb200_ctx* peer);
/* rank that owns (predicts, commits, holds the valid x and v of) every vertex of a body */
int sbsb200_get_vertex_ranks(const sbsb200_ctx* ctx, int body, int32_t* out, int64_t n);

/* ---- state ----------------------------------------------------------------------------- */

/* Overwrite x (and v, NULL = 0) of a body; xi = xn = x as tetrahedral_body_t::transform
 * leaves them (tetrahedral_body.cpp:121-132); refreshes the surface copy. 3*nV doubles. */
int sbsb200_upload(sbsb200_ctx* ctx, int body, const double* x, const double* v);

/* The same for n listed vertices of a body (body-local indices): what a UI does to a picked particle
 * between frames (main.cpp:158-165 writes simulation.particles()[b][v] directly; src/rendering/pick.cpp
 * finds the vertex).  x: 3*n doubles; v: 3*n doubles or NULL = leave the velocities alone. */
int sbsb200_set_vertices(sbsb200_ctx* ctx, int body, int64_t n, const uint32_t* vertices,
                         const double* x, const double* v);
/* particle_t::x(), v() of every vertex of a body (either may be NULL). */
int sbsb200_download(sbsb200_ctx* ctx, int body, double* x, double* v);
/* The boundary surface as a renderer consumes it, without downloading the whole state: 6 floats
 * (x, y, z, nx, ny, nz) per surface vertex from the surface copy of the last step
 * (tetrahedral_body_t::update_visual_model, tetrahedral_body.cpp:157-165; normals as
 * tetrahedral_mesh_boundary.cpp:122-145, but summed from zero at every call). */
int sbsb200_download_surface(sbsb200_ctx* ctx, int body, float* xyz_normal);
/* The same with the reference's 9-float vertex (prepare_vertices_for_surface_rendering,
 * tetrahedral_mesh_boundary.cpp:170-193): x, y, z, nx, ny, nz, r, g, b.  colours: 3 floats — the whole body in one
 * colour, as geometry_t::set_color leaves it (n_colours = 1) — or 3 per surface vertex (n_colours = surface vertices). */
int sbsb200_download_surface_rgb(sbsb200_ctx* ctx, int body, const float* colours, int64_t n_colours,
                                 float* xyz_normal_rgb);
/* particle_t::mass() = m (main.cpp:158-165 toggles 1 <-> 0 between frames). */
int sbsb200_set_mass(sbsb200_ctx* ctx, int body, int64_t vertex, double mass);
/* The same for n vertices of a body in one call (one copy, one synchronisation): what a host-side mirror of
 * simulation_t::particles() pushes after the caller edited masses (body-local vertex indices). */
int sbsb200_set_masses(sbsb200_ctx* ctx, int body, int64_t n, const uint32_t* vertices, const double* masses);

/* ---- the hot path ---------------------------------------------------------------------- */

/* timestep_t::step(simulation_t&) (src/physics/timestep.cpp:20-70): detection, then
 * `substeps` x (predict, `iterations` Gauss-Seidel sweeps, commit), then surface/BVH
 * update.  Asynchronous on the context's stream. */
int sbsb200_step(sbsb200_ctx* ctx, double dt, int substeps, int iterations, int detect_mode);

/* Same, with HOST buffers either side: uploads x_in/v_in (3*nV doubles each, v_in may be
 * NULL) for `body`, steps, downloads x_out/v_out and returns when they are valid. */
int sbsb200_step_host(sbsb200_ctx* ctx, int body, const double* x_in, const double* v_in,
                      double dt, int substeps, int iterations, int detect_mode, double* x_out,
                      double* v_out);

/* The same with float host buffers: half the bytes over PCIe when the caller keeps its particles in single
 * precision (a renderer-side mirror); the device state has the context's precision either way.
 * body = SBSB200_ALL_BODIES: every tetrahedral body at once, x / v concatenated in body order (an ensemble of
 * bodies crosses PCIe in one copy each way instead of one per body). */
#define SBSB200_ALL_BODIES (-1)
int sbsb200_step_host_f32(sbsb200_ctx* ctx, int body, const float* x_in, const float* v_in, double dt,
                          int substeps, int iterations, int detect_mode, float* x_out, float* v_out);

/* The same for n listed vertices of the body only (body-local indices; x_in NULL = keep the device state): what a
 * rank of a decomposed body exchanges with its host — the vertices it owns (sbsb200_get_vertex_ranks) — instead of
 * the whole body.  x, v: 3*n floats each, in the order of `vertices`. */
int sbsb200_step_host_vertices_f32(sbsb200_ctx* ctx, int body, int64_t n, const uint32_t* vertices,
                                   const float* x_in, const float* v_in, double dt, int substeps, int iterations,
                                   int detect_mode, float* x_out, float* v_out);

int sbsb200_synchronize(sbsb200_ctx* ctx);

/* Contacts of the most recent detection, as the arguments handed to
 * xpbd::contact_handler_t::handle (xpbd/contact_handler.cpp:14-54): tet body, tet-mesh vertex
 * (after from_surface_vertex), sdf body, contact point, unit normal.  Order is unspecified.
 * Pass NULL buffers to query the count.  Returns the count (>= 0) or an error. */
int64_t sbsb200_get_contacts(sbsb200_ctx* ctx, int64_t cap, int32_t* body, uint32_t* vertex,
                             int32_t* sdf_body, double* point, double* normal);

/* Validation aid: the number of vertices whose position or velocity is NaN / Inf (one kernel over the committed
 * state, synchronises).  The reference has no counterpart: it never checks.  Every sbsb200_step is also
 * wrapped in NVTX ranges ("sbsb200_step", "detection", "substep") for Nsight tools. */
int64_t sbsb200_count_non_finite(sbsb200_ctx* ctx);

/* Development aid (resident schedule): sbsb200_debug_trace_steps(ctx, N) before finalize selects a kernel
 * build that records %clock64 stamps of the first N colour steps of every launch, 16 per (region, step);
 * sbsb200_debug_read_trace returns the number of values available and copies min(cap, that) of them.
 * Not part of the reference API; N = 0 (default) runs the production kernels. */
int sbsb200_debug_trace_steps(sbsb200_ctx* ctx, int steps);
int64_t sbsb200_debug_read_trace(sbsb200_ctx* ctx, int64_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SBS_B200_H */
