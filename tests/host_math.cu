// TEST-ONLY harness: compiles the product's __host__ __device__ math for the CPU so that the
// algebra of the Green projection (eigen-decomposition of F^T F instead of a general SVD) can be
// checked against the oracle without a GPU.  Not part of libsbsb200.so; never shipped.
#include "../soft-body-simulator_b200/csrc/xpbd_kernels.cuh"

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
