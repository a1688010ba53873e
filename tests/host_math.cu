// TEST-ONLY harness: compiles the product's __host__ __device__ math for the CPU so that the
// algebra of the Green projection (eigen-decomposition of F^T F instead of a general SVD) can be
// checked against the oracle without a GPU.  Not part of libsbsb200.so; never shipped.
#include "../soft-body-simulator_b200/csrc/xpbd_kernels.cuh"

#include <vector>

using namespace sbsb200;

template <typename R>
static int project(double* xi, const double* xn, const double* w, const double* DmInv, double V0, double mu,
                   double lam, double alpha, double beta, double dt, double* lagrange)
{
    Real4<R> p[4];
    Vec3<R> n[4];
    for (int a = 0; a < 4; ++a)
    {
        p[a] = {R(xi[3 * a]), R(xi[3 * a + 1]), R(xi[3 * a + 2]), R(w[a])};
        n[a] = {R(xn[3 * a]), R(xn[3 * a + 1]), R(xn[3 * a + 2])};
    }
    Real4<R> r0 = {R(DmInv[0]), R(DmInv[1]), R(DmInv[2]), R(DmInv[3])};
    Real4<R> r1 = {R(DmInv[4]), R(DmInv[5]), R(DmInv[6]), R(DmInv[7])};
    Real4<R> r2 = {R(DmInv[8]), R(V0), R(0), R(0)};
    Real4<R> mat = {R(mu), R(lam), R(alpha), R(beta)};
    R l = R(*lagrange);
    R const l0 = l;
    if (beta != 0.)
        green_project<R, true>(p[0], p[1], p[2], p[3], n[0], n[1], n[2], n[3], r0, r1, r2, mat, R(dt), l);
    else
        green_project<R, false>(p[0], p[1], p[2], p[3], n[0], n[1], n[2], n[3], r0, r1, r2, mat, R(dt), l);
    for (int a = 0; a < 4; ++a)
    {
        xi[3 * a]     = double(p[a].x);
        xi[3 * a + 1] = double(p[a].y);
        xi[3 * a + 2] = double(p[a].z);
    }
    *lagrange = double(l);
    return l != l0;
}

extern "C" int hostmath_green_project_f64(double* xi, const double* xn, const double* w, const double* DmInv,
                                          double V0, double mu, double lam, double alpha, double beta, double dt,
                                          double* lagrange)
{
    return project<double>(xi, xn, w, DmInv, V0, mu, lam, alpha, beta, dt, lagrange);
}
extern "C" int hostmath_green_project_f32(double* xi, const double* xn, const double* w, const double* DmInv,
                                          double V0, double mu, double lam, double alpha, double beta, double dt,
                                          double* lagrange)
{
    return project<float>(xi, xn, w, DmInv, V0, mu, lam, alpha, beta, dt, lagrange);
}

// grid SDF sampling (csrc/grid_sdf.cuh) compiled for the host: value and gradient at n points;
// returns the number of points inside the domain
template <typename R>
static int grid_sample(const uint32_t* res, const double* lo, const double* hi, const double* nodes, int64_t n_nodes,
                       int n, const double* pts, double* val, double* grad)
{
    std::vector<R> nd(nodes, nodes + n_nodes);
    R const l[3] = {R(lo[0]), R(lo[1]), R(lo[2])}, h[3] = {R(hi[0]), R(hi[1]), R(hi[2])};
    int inside = 0;
    for (int i = 0; i < n; ++i)
    {
        R const p[3] = {R(pts[3 * i]), R(pts[3 * i + 1]), R(pts[3 * i + 2])};
        R phi = R(0), g[3] = {R(0), R(0), R(0)};
        bool const in = grid_interpolate<R>(res, l, h, nd.data(), p, phi, g);
        inside += in;
        val[i] = in ? double(phi) : 1.7976931348623157e308;
        for (int k = 0; k < 3; ++k)
            grad[3 * i + k] = in ? double(g[k]) : 0.;
    }
    return inside;
}
extern "C" int hostmath_grid_sample_f64(const uint32_t* res, const double* lo, const double* hi, const double* nodes,
                                        int64_t n_nodes, int n, const double* pts, double* val, double* grad)
{
    return grid_sample<double>(res, lo, hi, nodes, n_nodes, n, pts, val, grad);
}
extern "C" int hostmath_grid_sample_f32(const uint32_t* res, const double* lo, const double* hi, const double* nodes,
                                        int64_t n_nodes, int n, const double* pts, double* val, double* grad)
{
    return grid_sample<float>(res, lo, hi, nodes, n_nodes, n, pts, val, grad);
}
