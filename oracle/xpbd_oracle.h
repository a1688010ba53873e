/*
 * xpbd_oracle.h — CPU oracle for the XPBD hot path of Q-Minh/soft-body-simulator ("sbs").
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libsbsb200.so) never links, loads or calls anything in this directory.
 *
 * It is a plain-C, fp64, single-threaded restatement of the reference algorithm.  Every
 * function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Parity pinning: see the header of xpbd_oracle.c.
 */
#ifndef XPBD_ORACLE_H
#define XPBD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

/* ---- stateless helpers ------------------------------------------------------------------ */

/* get_simple_bar_model (src/geometry/get_simple_bar_model.cpp:6-122).
 * positions: 3*W*H*D floats, indices: 4*5*(W-1)(H-1)(D-1) ints. */
void orc_bar_model(int width, int height, int depth, float* positions, int32_t* indices);

/* 3x3 SVD, row-major in/out, sigma sorted descending and non-negative, full U and V
 * (the contract of Eigen::JacobiSVD<Matrix3d>(F, ComputeFullU|ComputeFullV) as used at
 * src/physics/xpbd/green_constraint.cpp:81-90). */
void orc_svd3(const double F[9], double U[9], double sigma[3], double V[9]);

/* Rest state of one Green constraint (green_constraint.cpp:33-46).
 * x0: 4 rest positions (12 doubles, vertex-major).  DmInv row-major. */
void orc_green_rest_state(const double x0[12], double DmInv[9], double* V0);

/* One call of green_constraint_t::project_positions (green_constraint.cpp:49-158) on a
 * stand-alone tet.  xi (12 doubles) and *lagrange are updated in place.
 * Returns 1 if the projection ran, 0 if it took the S < 1e-20 early-out (:130-131).
 * diag (optional, 8 doubles): sigma[3] after flip+clamp inputs (raw), C, S, delta_lambda,
 * inverted flag. */
int orc_green_project(double xi[12], const double xn[12], const double w[4], const double DmInv[9],
                      double V0, double young, double poisson, double alpha, double beta,
                      double dt, double* lagrange, double* diag);

/* Boundary surface of a tet mesh: restates tetrahedron_set_t::add_tetrahedron's triangle
 * numbering (src/physics/topology.cpp:904-956, faces order :335-342) and
 * tetrahedral_mesh_boundary_t::extract_boundary_surface
 * (src/physics/tetrahedral_mesh_boundary.cpp:65-120).
 * surf_to_tet: capacity nV.  tris: capacity 3*4*nT (surface-vertex indices).
 * Returns number of surface vertices; *n_tris receives the boundary triangle count. */
int orc_boundary_surface(int nV, int nT, const uint32_t* tets, uint32_t* surf_to_tet,
                         uint32_t* tris, int* n_tris);

/* ---- world API (mirrors the shape of include/sbs_b200.h so tests read symmetrically) ---- */

orc_world* orc_create(void);
void orc_destroy(orc_world* w);

/* simulation_parameters_t::collision_compliance (xpbd/simulation_parameters.h:24). */
void orc_set_collision_compliance(orc_world* w, double alpha);

/* tetrahedral_body_t(simulation, id, geometry) (src/physics/tetrahedral_body.cpp:29-83)
 * followed by one green_constraint_t per tet in tet order (main.cpp:37-54).
 * x0: 3*nV doubles; mass: nV doubles or NULL (=1, particle.cpp:7).
 * Returns the body index. */
int orc_add_tet_body(orc_world* w, int nV, const double* x0, const double* mass, int nT,
                     const uint32_t* tets, double young, double poisson, double alpha,
                     double beta);

/* distance_constraint_t(alpha, beta, sim, b1, b2, v1, v2) (xpbd/distance_constraint.cpp:8-22),
 * appended after everything added so far.  pairs: 2*n vertex indices. */
int orc_add_distance_constraints(orc_world* w, int b1, int b2, int n, const uint32_t* pairs,
                                 double alpha, double beta);

/* environment_body_t with an analytic sdf_model_t (sdf_model.cpp:25-30, :52-64).
 * Each takes one body slot.  volume = englobing AABB (min xyz, max xyz). */
int orc_add_sdf_plane(orc_world* w, const double normal[3], const double point[3],
                      const double volume[6]);
int orc_add_sdf_sphere(orc_world* w, const double centre[3], double radius,
                       const double volume[6]);
int orc_add_sdf_box(orc_world* w, const double bmin[3], const double bmax[3],
                    const double volume[6]);

/* ---- grid SDFs (Discregrid restated; PARITY UNPINNED, see xpbd_oracle.c) ---------------------- */
/* node count / node position / shape functions / interpolation of a CubicLagrangeDiscreteGrid with
 * res cells per axis over [dmin, dmax] */
int64_t orc_grid_node_count(const uint32_t res[3]);
void orc_grid_node_position(const double dmin[3], const double dmax[3], const uint32_t res[3], int64_t l,
                            double x[3]);
void orc_grid_shape(const double xi[3], double N[32], double dN[32][3]);
double orc_grid_interpolate(const double dmin[3], const double dmax[3], const uint32_t res[3],
                            const double* nodes, const double p[3], double grad[3]);
/* Discregrid::MeshDistance::signedDistanceCached at n points (closed, consistently oriented mesh) */
void orc_mesh_signed_distance(int nV, const double* x, int nF, const uint32_t* faces, int n, const double* points,
                              double* out);
/* environment_body.cpp:52-66 (domain extension) and :12-78 (bake); returns the node count */
void orc_mesh_sdf_domain(int nV, const double* x, const double domain[6], double out[6]);
int64_t orc_bake_mesh_sdf(int nV, const double* x, int nF, const uint32_t* faces, const double domain[6],
                          const uint32_t res[3], double out_domain[6], double* nodes, int64_t cap);
/* environment_body_t with a grid sdf_model_t; vol = NULL: volume() is the grid's domain
 * (environment_body.cpp:76) */
int orc_add_sdf_grid(orc_world* w, const double dmin[3], const double dmax[3], const uint32_t res[3],
                     const double* nodes, const double vol[6]);

/* brute_force_cd_system_t(objects) (collision/brute_force_cd_system.cpp:8-12, main.cpp:77-86): only the
 * collision models handed to the cd system take part in detection.  Default: every body does. */
int orc_set_body_collideable(orc_world* w, int body, int flag);

/* Number of elastic constraints (green + distance) in insertion order. */
int orc_constraint_count(const orc_world* w);

/* Reorder simulation_t::constraints_ : new_order[i] = insertion index of the constraint
 * that must run i-th ("the reference run with constraints permuted into the same colour
 * order").  Must be a permutation of 0..count-1.  Returns 0 on success. */
/* simulation_t::remove_constraint (src/physics/simulation.cpp:34-39): swap with the last, drop it */
int orc_remove_constraint(orc_world* w, uint32_t index);
int orc_set_constraint_order(orc_world* w, const uint32_t* new_order, int n);

/* Overwrite state (x, v) of a body; xi = xn = x (x0 untouched).  3*nV doubles each; v may
 * be NULL (zero). */
int orc_upload(orc_world* w, int body, const double* x, const double* v);
int orc_download(const orc_world* w, int body, double* x, double* v);
int orc_set_mass(orc_world* w, int body, int vertex, double mass);

/* timestep_t::step (src/physics/timestep.cpp:20-70).
 * detect_every_substep = 0: reference semantics (one detection per call, on the surface
 * copies taken at the end of the previous call).
 * detect_every_substep = 1: literally `substeps` calls of step() with substeps=1 and
 * dt/substeps. */
int orc_step(orc_world* w, double dt, int substeps, int iterations, int detect_every_substep);

/* Contacts produced by the most recent detection (before they were cleared at
 * timestep.cpp:68): body, tet-mesh vertex, sdf body, point(3), normal(3).
 * Pass NULL buffers to query the count. */
int orc_get_contacts(const orc_world* w, int cap, int32_t* body, uint32_t* vertex,
                     int32_t* sdf_body, double* point, double* normal);

/* Number of green projections executed / skipped by the early-out since creation. */
void orc_get_counters(const orc_world* w, uint64_t* projected, uint64_t* early_out);

#ifdef __cplusplus
}
#endif
#endif
