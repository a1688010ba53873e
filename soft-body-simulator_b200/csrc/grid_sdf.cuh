// grid_sdf.cuh — discrete signed-distance grids on the device.
//
// Replaces, for the non-analytic path of sdf_model_t::evaluate (src/physics/collision/sdf_model.cpp:71-74),
// Discregrid::CubicLagrangeDiscreteGrid::interpolate, and for environment_body_t's mesh constructor
// (src/physics/environment_body.cpp:12-78) Discregrid::MeshDistance sampled at the grid nodes.
// Discregrid is an un-vendored dependency of the reference (CMakeLists.txt:95-100, moving branch):
// its published algorithm is implemented here — 32-node serendipity cubic cells (8 corners + two
// nodes on each of the 12 edges), DBL_MAX outside the domain; exact point-triangle distance signed
// by the angle-weighted pseudo-normal of the closest feature.  Parity at this boundary is unpinned
// by the reference (no tests, no source); what pins it instead is listed in tests/test_grid_sdf.py.
//
// Node order of an (nx, ny, nz)-cell grid: the (nx+1)(ny+1)(nz+1) corners, x fastest; then two nodes
// per x-edge (edges x fastest, then y, then z), per y-edge (y fastest, then z, then x), per z-edge
// (z fastest, then x, then y).  Node values of all grids of a scene live in one array of R.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace sbsb200 {

struct GridDims
{
    uint32_t n[3];
    __host__ __device__ uint64_t corners() const { return uint64_t(n[0] + 1) * (n[1] + 1) * (n[2] + 1); }
    __host__ __device__ uint64_t edges(int axis) const
    {
        uint64_t e = 1;
        for (int d = 0; d < 3; ++d)
            e *= d == axis ? n[d] : n[d] + 1;
        return e;
    }
    __host__ __device__ uint64_t nodes() const { return corners() + 2 * (edges(0) + edges(1) + edges(2)); }
};

// position of node l (CubicLagrangeDiscreteGrid::indexToNodePosition)
__host__ __device__ inline void grid_node_position(GridDims const& g, double const lo[3], double const hi[3], uint64_t l,
                                                   double x[3])
{
    uint64_t ijk[3] = {0, 0, 0};
    int axis        = -1;
    uint64_t const nv = g.corners();
    if (l < nv)
    {
        ijk[0] = l % (g.n[0] + 1);
        ijk[1] = l / (g.n[0] + 1) % (g.n[1] + 1);
        ijk[2] = l / (uint64_t(g.n[0] + 1) * (g.n[1] + 1));
    }
    else
    {
        l -= nv;
        // fastest / middle / slowest axis of the edge numbering of each family
        int const order[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
        for (axis = 0; axis < 2 && l >= 2 * g.edges(axis); ++axis)
            l -= 2 * g.edges(axis);
        uint64_t e     = l / 2;
        int const a    = order[axis][0], b = order[axis][1], c = order[axis][2];
        uint64_t const na = g.n[a], nb = g.n[b] + 1;
        ijk[a] = e % na;
        ijk[b] = e / na % nb;
        ijk[c] = e / (na * nb);
    }
    for (int d = 0; d < 3; ++d)
        x[d] = lo[d] + (hi[d] - lo[d]) / double(g.n[d]) * double(ijk[d]);
    if (axis >= 0)
        x[axis] += (1.0 + double(l % 2)) / 3.0 * ((hi[axis] - lo[axis]) / double(g.n[axis]));
}

// value and gradient of the interpolant at p; returns false outside [lo, hi] (the reference then sees
// numeric_limits<double>::max(), i.e. "not penetrating")
template <typename R>
__host__ __device__ inline bool grid_interpolate(uint32_t const n[3], R const lo[3], R const hi[3], R const* nodes,
                                                 R const p[3], R& phi, R g[3])
{
    uint32_t m[3];
    R xi[3], c0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        if (!(lo[d] <= p[d] && p[d] <= hi[d]))
            return false;
        R const cell = (hi[d] - lo[d]) / R(n[d]);
        uint32_t k   = static_cast<uint32_t>((p[d] - lo[d]) * (R(1) / cell));
        k            = k >= n[d] ? n[d] - 1 : k;
        m[d]         = k;
        R const a = lo[d] + cell * R(k), b = a + cell;
        c0[d]     = R(2) / (b - a);
        xi[d]     = c0[d] * p[d] - (b + a) / (b - a);
    }
    uint32_t const nx = n[0], ny = n[1], nz = n[2];
    R const s[3][2] = {{R(1) - xi[0], R(1) + xi[0]}, {R(1) - xi[1], R(1) + xi[1]}, {R(1) - xi[2], R(1) + xi[2]}};
    R const r2      = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
    R const f       = R(9) * r2 - R(19);
    R v = R(0), gx = R(0), gy = R(0), gz = R(0);
    // corners: N = (1/64)(1 +- x)(1 +- y)(1 +- z)(9 r^2 - 19)
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
        int const bx = q & 1, by = q >> 1 & 1, bz = q >> 2 & 1;
        uint64_t const id = uint64_t(nx + 1) * (ny + 1) * (m[2] + bz) + uint64_t(nx + 1) * (m[1] + by) + m[0] + bx;
        R const w = nodes[id] * R(1.0 / 64.0);
        R const a = s[0][bx], b = s[1][by], c = s[2][bz];
        R const abc = a * b * c;
        v += w * abc * f;
        gx += w * ((bx ? b * c : -(b * c)) * f + abc * R(18) * xi[0]);
        gy += w * ((by ? a * c : -(a * c)) * f + abc * R(18) * xi[1]);
        gz += w * ((bz ? a * b : -(a * b)) * f + abc * R(18) * xi[2]);
    }
    // edge nodes at -+1/3 along axis t: N = (9/64)(1 - w^2)(1 -+ 3w)(1 +- u)(1 +- v)
    uint64_t off = uint64_t(nx + 1) * (ny + 1) * (nz + 1);
    R gr[3]      = {gx, gy, gz};
#pragma unroll
    for (int t = 0; t < 3; ++t)
    {
        int const u = t == 0 ? 1 : 0, w2 = t == 2 ? 1 : 2; // the two other axes, ascending
        // edge numbering: axis t fastest, then axis (t+1)%3, then (t+2)%3
        int const mid = (t + 1) % 3, slow = (t + 2) % 3;
        uint64_t const nt = n[t], nmid = n[mid] + 1;
        R const w   = xi[t];
        R const om  = R(1) - w * w;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            int const bu = q & 1, bv = q >> 1; // offsets along the axes u (lower) and w2 (higher)
            uint32_t idx[3];
            idx[t]  = m[t];
            idx[u]  = m[u] + bu;
            idx[w2] = m[w2] + bv;
            uint64_t const e = off + 2 * (uint64_t(idx[slow]) * nmid * nt + uint64_t(idx[mid]) * nt + idx[t]);
            R const fu = s[u][bu], fv = s[w2][bv];
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                R const sg  = h ? R(3) : R(-3);
                R const val = nodes[e + h] * R(9.0 / 64.0);
                R const gcu = om * (R(1) + sg * w);
                R const dg  = R(-2) * w * (R(1) + sg * w) + om * sg;
                v += val * gcu * fu * fv;
                gr[t] += val * dg * fu * fv;
                gr[u] += val * gcu * (bu ? fv : -fv);
                gr[w2] += val * gcu * (bv ? fu : -fu);
            }
        }
        off += 2 * (t == 0 ? uint64_t(nx) * (ny + 1) * (nz + 1) : t == 1 ? uint64_t(nx + 1) * ny * (nz + 1) : 0);
    }
    phi  = v;
    g[0] = gr[0] * c0[0];
    g[1] = gr[1] * c0[1];
    g[2] = gr[2] * c0[2];
    return true;
}

// ---- bake: Discregrid::MeshDistance at every grid node (environment_body.cpp:67-74) ----------------
// One record per triangle, precomputed on the host in fp64: corners, unit normal, the pseudo-normals of
// its three edges (own normal + the normal of the face across the edge) and of its three corners
// (angle-weighted sum over the incident faces).
struct BakeTriangle
{
    double p[3][3];
    double fn[3];
    double en[3][3]; // edge k: corner k -> corner k+1
    double vn[3][3];
};

// closest point of a triangle to p (Voronoi regions of the triangle's features); feature 0-2 corner,
// 3-5 edge k -> k+1, 6 interior
__device__ inline double closest_on_triangle(double const p[3], BakeTriangle const& t, double q[3], int& feature)
{
    double ab[3], ac[3], ap[3];
    for (int d = 0; d < 3; ++d)
    {
        ab[d] = t.p[1][d] - t.p[0][d];
        ac[d] = t.p[2][d] - t.p[0][d];
        ap[d] = p[d] - t.p[0][d];
    }
    auto const dot = [](double const* a, double const* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    double const d1 = dot(ab, ap), d2 = dot(ac, ap);
    double bp[3] = {p[0] - t.p[1][0], p[1] - t.p[1][1], p[2] - t.p[1][2]};
    double cp[3] = {p[0] - t.p[2][0], p[1] - t.p[2][1], p[2] - t.p[2][2]};
    double const d3 = dot(ab, bp), d4 = dot(ac, bp), d5 = dot(ab, cp), d6 = dot(ac, cp);
    double const vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    double s, u;
    if (d1 <= 0. && d2 <= 0.)
        feature = 0, s = 0., u = 0.;
    else if (d3 >= 0. && d4 <= d3)
        feature = 1, s = 1., u = 0.;
    else if (vc <= 0. && d1 >= 0. && d3 <= 0.)
        feature = 3, s = d1 / (d1 - d3), u = 0.;
    else if (d6 >= 0. && d5 <= d6)
        feature = 2, s = 0., u = 1.;
    else if (vb <= 0. && d2 >= 0. && d6 <= 0.)
        feature = 5, s = 0., u = d2 / (d2 - d6);
    else if (va <= 0. && (d4 - d3) >= 0. && (d5 - d6) >= 0.)
    {
        double const w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        feature = 4, s = 1. - w, u = w;
    }
    else
    {
        double const den = 1. / (va + vb + vc);
        feature = 6, s = vb * den, u = vc * den;
    }
    double r2 = 0.;
    for (int d = 0; d < 3; ++d)
    {
        q[d]           = t.p[0][d] + s * ab[d] + u * ac[d];
        double const r = p[d] - q[d];
        r2 += r * r;
    }
    return r2;
}

// thread per node; the triangles stream through shared memory in tiles of kBakeTile
constexpr int kBakeTile = 64;
__global__ void __launch_bounds__(128) k_bake_mesh_sdf(GridDims g, double lo0, double lo1, double lo2, double hi0,
                                                       double hi1, double hi2, BakeTriangle const* tris, int64_t n_tris,
                                                       double* nodes)
{
    __shared__ BakeTriangle tile[kBakeTile];
    uint64_t const l  = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    bool const active = l < g.nodes();
    double const lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
    double p[3] = {0., 0., 0.};
    if (active)
        grid_node_position(g, lo, hi, l, p);
    double best = 1.7976931348623157e308, dotn = 1.;
    for (int64_t first = 0; first < n_tris; first += kBakeTile)
    {
        int const count = static_cast<int>(n_tris - first < kBakeTile ? n_tris - first : kBakeTile);
        __syncthreads();
        {
            double const* src = reinterpret_cast<double const*>(tris + first);
            double* dst       = reinterpret_cast<double*>(tile);
            int const words   = count * static_cast<int>(sizeof(BakeTriangle) / sizeof(double));
            for (int i = threadIdx.x; i < words; i += blockDim.x)
                dst[i] = src[i];
        }
        __syncthreads();
        if (!active)
            continue;
        for (int i = 0; i < count; ++i)
        {
            double q[3];
            int feature;
            double const d2 = closest_on_triangle(p, tile[i], q, feature);
            if (d2 < best)
            { // strictly closer: the first triangle in mesh order wins ties, as a sequential search does
                best = d2;
                double const* n = feature < 3 ? tile[i].vn[feature] : feature < 6 ? tile[i].en[feature - 3] : tile[i].fn;
                dotn = (p[0] - q[0]) * n[0] + (p[1] - q[1]) * n[1] + (p[2] - q[2]) * n[2];
            }
        }
    }
    if (active)
    {
        double const dist = n_tris > 0 ? sqrt(best) : best;
        nodes[l]          = dotn < 0. ? -dist : dist;
    }
}

} // namespace sbsb200
