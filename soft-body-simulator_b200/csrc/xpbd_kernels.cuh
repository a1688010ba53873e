// Kernels of the "graph" schedule: one launch per stage / colour, thread per work item.
// (The persistent, region-resident schedule lives in xpbd_persistent.cuh.)
#pragma once

#include "grid_sdf.cuh"
#include "xpbd_math.cuh"

namespace sbsb200 {

// Device view of the scene.  Everything is SoA of 16-byte (fp32) / 32-byte (fp64) records so a
// thread's gather is one vector transaction per record.
constexpr int kShapeWords = 4; // Real4 records per entry of the rest-shape dictionary

template <typename R>
struct DeviceScene
{
    // vertices
    int64_t n_vertices;
    Real4<R>* pos;  // (xi.x, xi.y, xi.z, inverse mass)        particle_t::xi_, invmass()
    Real4<R>* prev; // (xn.x, xn.y, xn.z, unused)               particle_t::xn_ (== x_ between substeps)
    Real4<R>* vel;  // (v.x, v.y, v.z, unused)                  particle_t::v_
    // green constraints, colour-major
    int64_t n_tets;
    uint4 const* tet_v;     // 4 global vertex ids
    Real4<R> const* tet_r0; // DmInv row 0, d10
    Real4<R> const* tet_r1; // d11 d12 d20 d21
    Real4<R> const* tet_r2; // d22, signed V0, material id (as value), unused
    R* tet_lambda;
    // rest-shape dictionary: meshes with few distinct (DmInv, V0, material) records — every lattice has
    // ten — stream one byte per tet instead of 48 (n_shapes == 0: no dictionary)
    uint8_t const* tet_shape;
    Real4<R> const* shapes;    // [kShapeWords * n_shapes]: r0, r1, r2 of every distinct record and its material (mu, lambda, alpha, beta)
    int32_t n_shapes;
    Real4<R> const* materials; // (mu, lambda, alpha, beta)
    // distance constraints, colour-major
    int64_t n_dist;
    uint2 const* dist_v;
    Real4<R> const* dist_p; // (rest length, alpha, beta, unused)
    R* dist_lambda;
    // collision
    int64_t n_surface;
    uint32_t const* surf_v;    // global vertex id of surface vertex i
    int32_t const* surf_body;  // body of surface vertex i
    Real4<R>* surf_pos;        // visual-model copy the detection reads (tetrahedral_body.cpp:157-165)
    uint32_t* surf_first;      // index of the first contact of surface vertex i, 0xffffffff if none
    int32_t n_sdf;
    struct Sdf
    {
        int32_t kind, body;
        R a[3], b[3], r;    // plane: a = n, r = offset; sphere: a = centre, r; box / grid domain: a = min, b = max
        R vmin[3], vmax[3]; // englobing volume() box (collision_model.h:39-40), used by the BVH broadphase
        uint32_t grid_n[3]; // grid (kind 3): cells per axis and node values (grid_sdf.cuh)
        R const* grid_nodes;
    };
    Sdf const* sdf;
    int64_t contact_cap;
    uint32_t* contact_v;   // global vertex id | 0x80000000 when first contact of its vertex
    Real4<R>* contact_q;   // (qs.x, qs.y, qs.z, lambda)
    Real4<R>* contact_n;   // (n.x, n.y, n.z, sdf body as value)
    uint32_t* contact_count;
    R collision_alpha;
};

// timestep.cpp:35-43.  a = f * invmass with f = (0, -9.81, 0) accumulated once per substep
// (f is zeroed by the commit loop, :55).  Only the y component has a non-zero force.
template <typename R>
__device__ __forceinline__ void predict_vertex(Real4<R>& p, Real4<R> const& x, Real4<R>& v, R dt)
{
    R const fy = R(0) - R(9.81);
    v.y        = add_rn(v.y, mul_rn(mul_rn(fy, p.w), dt));
    p.x        = add_rn(x.x, mul_rn(v.x, dt));
    p.y        = add_rn(x.y, mul_rn(v.y, dt));
    p.z        = add_rn(x.z, mul_rn(v.z, dt));
}
// timestep.cpp:48-57
template <typename R>
__device__ __forceinline__ void commit_vertex(Real4<R> const& p, Real4<R>& xn, Real4<R>& v, R dt)
{
    v.x  = div_rn(sub_rn(p.x, xn.x), dt);
    v.y  = div_rn(sub_rn(p.y, xn.y), dt);
    v.z  = div_rn(sub_rn(p.z, xn.z), dt);
    xn.x = p.x;
    xn.y = p.y;
    xn.z = p.z;
}

template <typename R>
__global__ void __launch_bounds__(256) k_predict(DeviceScene<R> s, R dt)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= s.n_vertices)
        return;
    Real4<R> p       = ld4(&s.pos[i]);
    Real4<R> const x = ld4(&s.prev[i]);
    Real4<R> v       = ld4(&s.vel[i]);
    predict_vertex(p, x, v, dt);
    st4(&s.vel[i], v);
    st4(&s.pos[i], p);
}

// timestep.cpp:48-57
template <typename R>
__global__ void __launch_bounds__(256) k_integrate(DeviceScene<R> s, R dt)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= s.n_vertices)
        return;
    Real4<R> const p = ld4(&s.pos[i]);
    Real4<R> xn      = ld4(&s.prev[i]);
    Real4<R> v       = ld4(&s.vel[i]);
    commit_vertex(p, xn, v, dt);
    st4(&s.vel[i], v);
    st4(&s.prev[i], xn);
}

// One Green projection given gathered data; returns updated positions/lambda through refs.
// Steps 10-13 of green_constraint.cpp (:123-157).  `at` = alpha / dt^2 (:134).
template <typename R, bool kDamped>
SBS_HD void green_project_at(Real4<R>& p1, Real4<R>& p2, Real4<R>& p3, Real4<R>& p4, Vec3<R> xn1, Vec3<R> xn2,
                             Vec3<R> xn3, Vec3<R> xn4, Real4<R> r0, Real4<R> r1, Real4<R> r2, R mu, R lam, R at,
                             R beta, R dt, R& lambda)
{
    Vec3<R> const x1 = {p1.x, p1.y, p1.z}, x2 = {p2.x, p2.y, p2.z}, x3 = {p3.x, p3.y, p3.z},
                  x4 = {p4.x, p4.y, p4.z};
    GreenOut<R> const g =
        green_gradients<R>(x1, x2, x3, x4, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, mu, lam);
    Vec3<R> const f4 = {-(g.f1.x + g.f2.x + g.f3.x), -(g.f1.y + g.f2.y + g.f3.y), -(g.f1.z + g.f2.z + g.f3.z)};
    R const S = p1.w * dot(g.f1, g.f1) + p2.w * dot(g.f2, g.f2) + p3.w * dot(g.f3, g.f3) + p4.w * dot(f4, f4);
    if (S < R(1e-20)) // :67, :130-131
        return;
    R num = -(g.C + at * lambda);
    R den = S + at;
    if (kDamped)
    {
        R const bt  = beta * (dt * dt);
        R const gam = at * bt / dt;
        R const gd  = dot(g.f1, x1 - xn1) + dot(g.f2, x2 - xn2) + dot(g.f3, x3 - xn3) + dot(f4, x4 - xn4);
        num += gam * gd;
        den = (R(1) + gam) * S + at;
    }
    R const dl = div_(num, den);
    lambda += dl;
    R const k1 = -p1.w * dl, k2 = -p2.w * dl, k3 = -p3.w * dl, k4 = -p4.w * dl;
    p1.x += k1 * g.f1.x; p1.y += k1 * g.f1.y; p1.z += k1 * g.f1.z;
    p2.x += k2 * g.f2.x; p2.y += k2 * g.f2.y; p2.z += k2 * g.f2.z;
    p3.x += k3 * g.f3.x; p3.y += k3 * g.f3.y; p3.z += k3 * g.f3.z;
    p4.x += k4 * f4.x;   p4.y += k4 * f4.y;   p4.z += k4 * f4.z;
}

template <typename R, bool kDamped>
SBS_HD void green_project(Real4<R>& p1, Real4<R>& p2, Real4<R>& p3, Real4<R>& p4,
                                              Vec3<R> xn1, Vec3<R> xn2, Vec3<R> xn3, Vec3<R> xn4,
                                              Real4<R> r0, Real4<R> r1, Real4<R> r2, Real4<R> mat,
                                              R dt, R& lambda)
{
    green_project_at<R, kDamped>(p1, p2, p3, p4, xn1, xn2, xn3, xn4, r0, r1, r2, mat.x, mat.y, mat.z / (dt * dt),
                                 mat.w, dt, lambda);
}

__device__ __forceinline__ int mat_index(float v) { return __float_as_int(v); }
__device__ __forceinline__ int mat_index(double v) { return static_cast<int>(v); }

// Chunk descriptor of the clustered colouring (mirrors sbsb200::ChunkDesc in scene_build.h)
struct DevChunk
{
    int32_t first;  // storage index of the chunk's first tet
    int32_t cfirst; // storage index of the chunk's first cluster
    int32_t n[8];
};

// One colour of Green constraints: thread per CLUSTER, the cluster's tets one after the other
// (they share vertices; clusters of one colour do not).  Tet m of cluster i sits at
// first + n[0] + .. + n[m-1] + i, so every pass over m is a coalesced sweep.  first_iteration
// folds the lambda reset of constraint_t::prepare_for_projection (constraint.cpp:12-16) into the
// first sweep.
template <typename R, bool kDamped, bool kDict>
__global__ void __launch_bounds__(128)
k_project_green(DeviceScene<R> s, DevChunk chunk, R dt, int first_iteration)
{
    extern __shared__ __align__(16) unsigned char dict_raw[];
    Real4<R>* s_dict = reinterpret_cast<Real4<R>*>(dict_raw);
    if (kDict)
    {
        for (int w = threadIdx.x; w < kShapeWords * s.n_shapes; w += blockDim.x)
            s_dict[w] = s.shapes[w];
        __syncthreads();
    }
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= chunk.n[0])
        return;
    int64_t base = chunk.first;
#pragma unroll 1
    for (int m = 0; m < 8; ++m)
    {
        if (i >= chunk.n[m])
            break;
        int64_t const t   = base + i;
        base += chunk.n[m];
        uint4 const v     = __ldg(&s.tet_v[t]);
        Real4<R> r0, r1, r2;
        if (kDict)
        {
            int const sh = __ldg(&s.tet_shape[t]);
            r0           = s_dict[kShapeWords * sh];
            r1           = s_dict[kShapeWords * sh + 1];
            r2           = s_dict[kShapeWords * sh + 2];
        }
        else
        {
            r0 = ld4_ro(&s.tet_r0[t]);
            r1 = ld4_ro(&s.tet_r1[t]);
            r2 = ld4_ro(&s.tet_r2[t]);
        }
        Real4<R> p1 = ld4(&s.pos[v.x]), p2 = ld4(&s.pos[v.y]), p3 = ld4(&s.pos[v.z]), p4 = ld4(&s.pos[v.w]);
        Real4<R> const mat = ld4_ro(&s.materials[mat_index(r2.z)]);
        R lambda           = first_iteration ? R(0) : s.tet_lambda[t];
        Vec3<R> xn1{}, xn2{}, xn3{}, xn4{};
        if (kDamped)
        {
            Real4<R> const a = ld4(&s.prev[v.x]), b = ld4(&s.prev[v.y]), c = ld4(&s.prev[v.z]), d = ld4(&s.prev[v.w]);
            xn1 = {a.x, a.y, a.z};
            xn2 = {b.x, b.y, b.z};
            xn3 = {c.x, c.y, c.z};
            xn4 = {d.x, d.y, d.z};
        }
        R const lambda_in = lambda;
        green_project<R, kDamped>(p1, p2, p3, p4, xn1, xn2, xn3, xn4, r0, r1, r2, mat, dt, lambda);
        if (lambda != lambda_in || first_iteration)
            s.tet_lambda[t] = lambda;
        if (lambda != lambda_in)
        {
            st4(&s.pos[v.x], p1);
            st4(&s.pos[v.y], p2);
            st4(&s.pos[v.z], p3);
            st4(&s.pos[v.w], p4);
        }
    }
}

// distance_constraint.cpp:24-53, one colour, thread per constraint
template <typename R>
__global__ void __launch_bounds__(128)
k_project_distance(DeviceScene<R> s, int64_t first, int32_t count, R dt, int first_iteration)
{
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    int64_t const c  = first + i;
    uint2 const v    = __ldg(&s.dist_v[c]);
    Real4<R> const q = ld4_ro(&s.dist_p[c]);
    Real4<R> p1 = ld4(&s.pos[v.x]), p2 = ld4(&s.pos[v.y]);
    Real4<R> const n1 = ld4(&s.prev[v.x]), n2 = ld4(&s.prev[v.y]);
    R lambda          = first_iteration ? R(0) : s.dist_lambda[c];
    Vec3<R> const diff = {p1.x - p2.x, p1.y - p2.y, p1.z - p2.z};
    R const len        = sqrt_(dot(diff, diff));
    Vec3<R> const n    = {diff.x / len, diff.y / len, diff.z / len};
    R const C          = len - q.x;
    R const S          = p1.w + p2.w;
    R const dt2        = dt * dt;
    R const at = q.y / dt2, bt = q.z * dt2;
    Vec3<R> const d1 = {p1.x - n1.x, p1.y - n1.y, p1.z - n1.z}, d2 = {p2.x - n2.x, p2.y - n2.y, p2.z - n2.z};
    R const gd  = dot(n, d1) - dot(n, d2);
    R const gam = at * bt / dt;
    R const dl  = (-(C + at * lambda) + gam * gd) / ((R(1) + gam) * S + at);
    lambda += dl;
    s.dist_lambda[c] = lambda;
    R const k1 = p1.w * dl, k2 = -p2.w * dl;
    p1.x += k1 * n.x; p1.y += k1 * n.y; p1.z += k1 * n.z;
    p2.x += k2 * n.x; p2.y += k2 * n.y; p2.z += k2 * n.z;
    st4(&s.pos[v.x], p1);
    st4(&s.pos[v.y], p2);
}

// sdf_model_t::evaluate (sdf_model.cpp:66-75): signed distance + gradient; analytic kinds and the
// discrete grid (Discregrid's interpolate, grid_sdf.cuh)
template <typename R>
__device__ __forceinline__ R sdf_eval(typename DeviceScene<R>::Sdf const& f, Vec3<R> p, Vec3<R>& g)
{
    if (f.kind == 3)
    {
        R const q[3] = {p.x, p.y, p.z};
        R phi, gr[3];
        if (!grid_interpolate<R>(f.grid_n, f.a, f.b, f.grid_nodes, q, phi, gr))
        {
            g = {R(0), R(1), R(0)};
            return R(3.0e38); // outside the domain the reference sees numeric_limits<double>::max()
        }
        g = {gr[0], gr[1], gr[2]};
        return phi;
    }
    if (f.kind == 0)
    { // plane: Eigen::Hyperplane::signedDistance = n.p + offset (sdf_model.cpp:56-61)
        g = {f.a[0], f.a[1], f.a[2]};
        return f.a[0] * p.x + f.a[1] * p.y + f.a[2] * p.z + f.r;
    }
    if (f.kind == 1)
    { // sphere
        Vec3<R> const d = {p.x - f.a[0], p.y - f.a[1], p.z - f.a[2]};
        R const len     = sqrt_(dot(d, d));
        if (len > R(0))
            g = {d.x / len, d.y / len, d.z / len};
        else
            g = {R(0), R(1), R(0)};
        return len - f.r;
    }
    // box
    R const pc[3] = {p.x, p.y, p.z};
    R q[3], c[3], out2 = R(0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        c[i]      = R(0.5) * (f.a[i] + f.b[i]);
        R const h = R(0.5) * (f.b[i] - f.a[i]);
        q[i]      = abs_(pc[i] - c[i]) - h;
        if (q[i] > R(0))
            out2 += q[i] * q[i];
    }
    R gg[3] = {R(0), R(0), R(0)};
    R sd;
    if (out2 > R(0))
    {
        R const len = sqrt_(out2);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            gg[i] = q[i] > R(0) ? (pc[i] >= c[i] ? q[i] : -q[i]) / len : R(0);
        sd = len;
    }
    else
    {
        int ax = 0;
        if (q[1] > q[ax]) ax = 1;
        if (q[2] > q[ax]) ax = 2;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i == ax)
                gg[i] = pc[i] >= c[i] ? R(1) : R(-1);
        sd = q[ax];
    }
    g = {gg[0], gg[1], gg[2]};
    return sd;
}

// collision_constraint.cpp:21-48.  The thread owning the first contact of a vertex walks the
// vertex's contacts in order, so several SDFs touching one vertex are applied sequentially as
// in the reference's serial sweep.
template <typename R>
__global__ void __launch_bounds__(256) k_project_collision(DeviceScene<R> s, R dt, int first_iteration)
{
    int64_t const i   = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    uint32_t const n  = min(*s.contact_count, static_cast<uint32_t>(s.contact_cap));
    if (i >= n)
        return;
    uint32_t const tag = s.contact_v[i];
    if (!(tag & 0x80000000u))
        return;
    uint32_t const gv = tag & 0x7fffffffu;
    Real4<R> p        = ld4(&s.pos[gv]);
    R const at        = s.collision_alpha / (dt * dt);
    bool moved        = false;
    for (int64_t j = i; j < n; ++j)
    {
        if (j > i && s.contact_v[j] != gv)
            break;
        Real4<R> q       = ld4(&s.contact_q[j]);
        Real4<R> const m = ld4(&s.contact_n[j]);
        R lambda         = first_iteration ? R(0) : q.w;
        R const C        = (p.x - q.x) * m.x + (p.y - q.y) * m.y + (p.z - q.z) * m.z;
        if (C >= R(0))
        {
            if (first_iteration)
            {
                q.w = R(0);
                st4(&s.contact_q[j], q);
            }
            continue;
        }
        R const dl = -(C + at * lambda) / (p.w + at);
        lambda += dl;
        p.x += p.w * m.x * dl;
        p.y += p.w * m.y * dl;
        p.z += p.w * m.z * dl;
        q.w = lambda;
        st4(&s.contact_q[j], q);
        moved = true;
    }
    if (moved)
        st4(&s.pos[gv], p);
}

// tetrahedral_body_t::update_visual_model (tetrahedral_body.cpp:157-165): surface copy of x
template <typename R>
__global__ void __launch_bounds__(256) k_surface_gather(DeviceScene<R> s)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= s.n_surface)
        return;
    st4(&s.surf_pos[i], ld4(&s.prev[s.surf_v[i]]));
}

// Boundary surface for a renderer: per-vertex normals as the reference computes them
// (tetrahedral_mesh_boundary.cpp:122-145: sum of (p2 - p1) x (p3 - p1) over the incident boundary
// triangles, normalised) — except that the sum starts from zero at every call (the reference never
// resets it).  Positions come from the surface copy.
__device__ __forceinline__ void atomic_add3(Real4<float>* p, float x, float y, float z)
{
    atomicAdd(&p->x, x);
    atomicAdd(&p->y, y);
    atomicAdd(&p->z, z);
}
__device__ __forceinline__ void atomic_add3(Real4<double>* p, double x, double y, double z)
{
    atomicAdd(&p->x, x);
    atomicAdd(&p->y, y);
    atomicAdd(&p->z, z);
}
template <typename R>
__global__ void __launch_bounds__(256)
k_surface_normals(DeviceScene<R> s, int64_t first_triangle, int64_t n_triangles, uint32_t const* __restrict__ tri,
                  int64_t first_surface, Real4<R>* __restrict__ normal)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_triangles)
        return;
    uint32_t const a = tri[3 * (first_triangle + i)], b = tri[3 * (first_triangle + i) + 1],
                   c = tri[3 * (first_triangle + i) + 2];
    Real4<R> const p1 = ld4(&s.surf_pos[first_surface + a]), p2 = ld4(&s.surf_pos[first_surface + b]),
                   p3 = ld4(&s.surf_pos[first_surface + c]);
    Vec3<R> const n = cross(Vec3<R>{p2.x - p1.x, p2.y - p1.y, p2.z - p1.z}, Vec3<R>{p3.x - p1.x, p3.y - p1.y, p3.z - p1.z});
    atomic_add3(&normal[first_surface + a], n.x, n.y, n.z);
    atomic_add3(&normal[first_surface + b], n.x, n.y, n.z);
    atomic_add3(&normal[first_surface + c], n.x, n.y, n.z);
}
// (x, y, z, nx, ny, nz) as floats, the vertex layout of the reference's renderer (renderer.cpp:484-542)
// With `colour` (one rgb triple, or one per surface vertex) the record is the reference's 9-float vertex
// (prepare_vertices_for_surface_rendering, tetrahedral_mesh_boundary.cpp:170-193): position, normal, colour.
template <typename R>
__global__ void __launch_bounds__(256)
k_surface_pack(DeviceScene<R> s, int64_t first_surface, int64_t n, Real4<R> const* __restrict__ normal,
               float* __restrict__ out, float const* __restrict__ colour = nullptr, int64_t n_colours = 0)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    Real4<R> const p = ld4(&s.surf_pos[first_surface + i]);
    Real4<R> const m = ld4(&normal[first_surface + i]);
    R const len2     = m.x * m.x + m.y * m.y + m.z * m.z;
    R const inv      = len2 > R(0) ? R(1) / sqrt_(len2) : R(0);
    int64_t const w  = colour ? 9 : 6;
    out[w * i]       = float(p.x);
    out[w * i + 1]   = float(p.y);
    out[w * i + 2]   = float(p.z);
    out[w * i + 3]   = float(m.x * inv);
    out[w * i + 4]   = float(m.y * inv);
    out[w * i + 5]   = float(m.z * inv);
    if (colour)
    {
        float const* c = colour + (n_colours > 1 ? 3 * i : 0);
        out[w * i + 6] = c[0];
        out[w * i + 7] = c[1];
        out[w * i + 8] = c[2];
    }
}

// Host-format (3 doubles per vertex) <-> device layout.  Upload sets x = xi = xn and keeps the
// inverse mass (tetrahedral_body_t::transform semantics, tetrahedral_body.cpp:121-132).
template <typename R, typename H>
__global__ void __launch_bounds__(256)
k_unpack_state(DeviceScene<R> s, int64_t first, int64_t n, H const* __restrict__ x, H const* __restrict__ v)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    Real4<R> p = ld4(&s.pos[first + i]);
    p.x        = R(x[3 * i]);
    p.y        = R(x[3 * i + 1]);
    p.z        = R(x[3 * i + 2]);
    st4(&s.pos[first + i], p);
    st4(&s.prev[first + i], Real4<R>{p.x, p.y, p.z, R(0)});
    Real4<R> vel = {R(0), R(0), R(0), R(0)};
    if (v)
        vel = {R(v[3 * i]), R(v[3 * i + 1]), R(v[3 * i + 2]), R(0)};
    st4(&s.vel[first + i], vel);
}

// the same for a handful of vertices (dragging a picked vertex between frames): x = xi = xn <- x[i] and,
// when given, v <- v[i] for the listed vertices of one body; a null v leaves the velocity alone
template <typename R, typename H>
__global__ void __launch_bounds__(256)
k_scatter_state(DeviceScene<R> s, int64_t first, int64_t n, uint32_t const* __restrict__ which,
                H const* __restrict__ x, H const* __restrict__ v)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    int64_t const at = first + which[i];
    Real4<R> p       = ld4(&s.pos[at]);
    p.x              = R(x[3 * i]);
    p.y              = R(x[3 * i + 1]);
    p.z              = R(x[3 * i + 2]);
    st4(&s.pos[at], p);
    st4(&s.prev[at], Real4<R>{p.x, p.y, p.z, R(0)});
    if (v)
        st4(&s.vel[at], Real4<R>{R(v[3 * i]), R(v[3 * i + 1]), R(v[3 * i + 2]), R(0)});
}

// ... and back: x and v of the listed vertices (sbsb200_step_host_vertices_f32)
template <typename R, typename H>
__global__ void __launch_bounds__(256)
k_gather_state(DeviceScene<R> s, int64_t first, int64_t n, uint32_t const* __restrict__ which, H* __restrict__ x,
               H* __restrict__ v)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    int64_t const at = first + which[i];
    if (x)
    {
        Real4<R> const p = ld4(&s.prev[at]);
        x[3 * i]         = H(p.x);
        x[3 * i + 1]     = H(p.y);
        x[3 * i + 2]     = H(p.z);
    }
    if (v)
    {
        Real4<R> const q = ld4(&s.vel[at]);
        v[3 * i]         = H(q.x);
        v[3 * i + 1]     = H(q.y);
        v[3 * i + 2]     = H(q.z);
    }
}

// sbsb200_remove_constraints: rest volume zero for the tets at these storage positions (per-tet record, and the
// zero-volume twin of their dictionary record when there is a dictionary)
template <typename R>
__global__ void __launch_bounds__(256)
k_remove_tets(Real4<R>* __restrict__ tet_r2, uint8_t* __restrict__ tet_shape, int32_t n_base_shapes,
              uint32_t const* __restrict__ positions, int64_t n)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    uint32_t const p = positions[i];
    tet_r2[p].y      = R(0);
    if (tet_shape && tet_shape[p] < n_base_shapes)
        tet_shape[p] = static_cast<uint8_t>(tet_shape[p] + n_base_shapes);
}

// particle_t::mass() of a handful of vertices (sbsb200_set_masses): the inverse mass rides in pos[].w
template <typename R>
__global__ void __launch_bounds__(256)
k_scatter_inverse_mass(DeviceScene<R> s, int64_t first, int64_t n, uint32_t const* __restrict__ which,
                       R const* __restrict__ w)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n)
        s.pos[first + which[i]].w = w[i];
}

template <typename R, typename H>
__global__ void __launch_bounds__(256)
k_pack_state(DeviceScene<R> s, int64_t first, int64_t n, H* __restrict__ x, H* __restrict__ v)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    if (x)
    {
        Real4<R> const p = ld4(&s.prev[first + i]);
        x[3 * i]         = H(p.x);
        x[3 * i + 1]     = H(p.y);
        x[3 * i + 2]     = H(p.z);
    }
    if (v)
    {
        Real4<R> const q = ld4(&s.vel[first + i]);
        v[3 * i]         = H(q.x);
        v[3 * i + 1]     = H(q.y);
        v[3 * i + 2]     = H(q.z);
    }
}

// validation aid (sbsb200_count_non_finite): vertices whose committed position or velocity is not finite
template <typename R>
__global__ void __launch_bounds__(256) k_count_non_finite(DeviceScene<R> s, unsigned long long* __restrict__ count)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    bool bad        = false;
    if (i < s.n_vertices)
    {
        Real4<R> const p = ld4(&s.prev[i]);
        Real4<R> const v = ld4(&s.vel[i]);
        bad = !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && isfinite(v.x) && isfinite(v.y) && isfinite(v.z));
    }
    unsigned const n = __popc(__ballot_sync(0xffffffffu, bad));
    if ((threadIdx.x & 31) == 0 && n)
        atomicAdd(count, static_cast<unsigned long long>(n));
}

// sdf_model_t::evaluate at arbitrary points (sbsb200_eval_sdf): out = (distance, gradient) per point
template <typename R>
__global__ void __launch_bounds__(128)
k_eval_sdf(DeviceScene<R> s, int32_t k, int64_t n, double const* __restrict__ pts, double* __restrict__ out)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    Vec3<R> g;
    R const sd     = sdf_eval<R>(s.sdf[k], Vec3<R>{R(pts[3 * i]), R(pts[3 * i + 1]), R(pts[3 * i + 2])}, g);
    bool const out_of_domain = s.sdf[k].kind == 3 && sd >= R(3.0e38);
    out[4 * i]     = out_of_domain ? 1.7976931348623157e308 : static_cast<double>(sd);
    out[4 * i + 1] = out_of_domain ? 0. : static_cast<double>(g.x);
    out[4 * i + 2] = out_of_domain ? 0. : static_cast<double>(g.y);
    out[4 * i + 3] = out_of_domain ? 0. : static_cast<double>(g.z);
}

} // namespace sbsb200
