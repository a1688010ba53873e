// Device math of the XPBD hot path, templated on the arithmetic type (float = production,
// double = validation build).  No reference code here: the algebra is derived in DESIGN.md
// ("Green projection without a general SVD") from src/physics/xpbd/green_constraint.cpp:49-158.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// math that is also compiled for the host by the test-only harness tests/host_math.cu
#define SBS_HD __host__ __device__ __forceinline__

namespace sbsb200 {

template <typename R>
struct alignas(4 * sizeof(R) > 16 ? 16 : 4 * sizeof(R)) Real4
{
    R x, y, z, w;
};
static_assert(sizeof(Real4<float>) == 16 && alignof(Real4<float>) == 16, "float4 layout");
static_assert(sizeof(Real4<double>) == 32, "double4 layout");

template <typename R>
struct Vec3
{
    R x, y, z;
};

template <typename R>
SBS_HD Vec3<R> operator-(Vec3<R> a, Vec3<R> b)
{
    return {a.x - b.x, a.y - b.y, a.z - b.z};
}
template <typename R>
SBS_HD R dot(Vec3<R> a, Vec3<R> b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <typename R>
SBS_HD Vec3<R> cross(Vec3<R> a, Vec3<R> b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// fp32 uses the SFU approximations (MUFU.RSQ / MUFU.RCP, ~2 ulp) on the device: the IEEE
// sequences for 1/sqrt, sqrt and division are 20-40 dependent instructions each and sat on the
// critical path of every Jacobi rotation.  fp64 (validation build) stays IEEE.
SBS_HD float rsqrt_(float x)
{
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
SBS_HD double rsqrt_(double x) { return 1.0 / sqrt(x); }
SBS_HD float sqrt_(float x)
{
#ifdef __CUDA_ARCH__
    return x > 0.0f ? x * rsqrtf(x) : 0.0f;
#else
    return sqrtf(x);
#endif
}
SBS_HD double sqrt_(double x) { return sqrt(x); }
SBS_HD float div_(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
SBS_HD double div_(double a, double b) { return a / b; }
SBS_HD float abs_(float x) { return fabsf(x); }
SBS_HD double abs_(double x) { return fabs(x); }
SBS_HD float max_(float a, float b) { return fmaxf(a, b); }
SBS_HD double max_(double a, double b) { return fmax(a, b); }

// Rounding-exact arithmetic (never contracted into FMA): predict/commit are order-independent
// stages that must match the fp64 reference bit for bit in the validation build.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// vector loads/stores of (x, y, z, w) records
__device__ __forceinline__ Real4<float> ld4(Real4<float> const* p)
{
    float4 const v = *reinterpret_cast<float4 const*>(p);
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void st4(Real4<float>* p, Real4<float> v)
{
    *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ Real4<double> ld4(Real4<double> const* p)
{
    double2 const a = reinterpret_cast<double2 const*>(p)[0];
    double2 const b = reinterpret_cast<double2 const*>(p)[1];
    return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st4(Real4<double>* p, Real4<double> v)
{
    reinterpret_cast<double2*>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2*>(p)[1] = make_double2(v.z, v.w);
}
// read-only (non-coherent) path for per-constraint constants
__device__ __forceinline__ Real4<float> ld4_ro(Real4<float> const* p)
{
    float4 const v = __ldg(reinterpret_cast<float4 const*>(p));
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Real4<double> ld4_ro(Real4<double> const* p)
{
    double2 const a = __ldg(reinterpret_cast<double2 const*>(p));
    double2 const b = __ldg(reinterpret_cast<double2 const*>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}

template <typename R>
struct Eps;
template <>
struct Eps<float>
{
    static constexpr float off_rel = 1e-13f; // off^2 <= off_rel * diag^2  (~ (3e-7)^2)
    static constexpr int max_sweeps = 6;

};
template <>
struct Eps<double>
{
    static constexpr double off_rel = 1e-30;
    static constexpr int max_sweeps = 12;

};

// One Jacobi rotation on the symmetric matrix (app, aqq, apq, arp, arq) and the eigenvector
// columns p, q.  r is the third index.
template <typename R>
SBS_HD void jacobi_rotate(R& app, R& aqq, R& apq, R& arp, R& arq, Vec3<R>& vp,
                                              Vec3<R>& vq)
{
    if (apq == R(0))
        return;
    R const d   = aqq - app;
    R const den = abs_(d) + sqrt_(d * d + R(4) * apq * apq);
    R const t   = div_((d >= R(0) ? R(2) : R(-2)) * apq, den);
    R const c   = rsqrt_(R(1) + t * t);
    R const s   = t * c;
    app -= t * apq;
    aqq += t * apq;
    apq = R(0);
    R const rp = c * arp - s * arq;
    R const rq = s * arp + c * arq;
    arp        = rp;
    arq        = rq;
    Vec3<R> const np = {c * vp.x - s * vq.x, c * vp.y - s * vq.y, c * vp.z - s * vq.z};
    Vec3<R> const nq = {s * vp.x + c * vq.x, s * vp.y + c * vq.y, s * vp.z + c * vq.z};
    vp               = np;
    vq               = nq;
}

// Eigen-decomposition of the symmetric 3x3 A (a00 a11 a22 a01 a02 a12): cyclic Jacobi.
// Output: eigenvalues l0 >= l1 >= l2 and a PROPER rotation (v0 v1 v2) of eigenvectors.
template <typename R>
SBS_HD void sym_eig3(R a00, R a11, R a22, R a01, R a02, R a12, R& l0, R& l1,
                                         R& l2, Vec3<R>& v0, Vec3<R>& v1, Vec3<R>& v2)
{
    v0 = {R(1), R(0), R(0)};
    v1 = {R(0), R(1), R(0)};
    v2 = {R(0), R(0), R(1)};
#pragma unroll 1
    for (int sweep = 0; sweep < Eps<R>::max_sweeps; ++sweep)
    {
        R const off2  = a01 * a01 + a02 * a02 + a12 * a12;
        R const diag2 = a00 * a00 + a11 * a11 + a22 * a22;
        if (off2 <= Eps<R>::off_rel * diag2)
            break;
        jacobi_rotate(a00, a11, a01, a02, a12, v0, v1); // (p,q) = (0,1), r = 2
        jacobi_rotate(a00, a22, a02, a01, a12, v0, v2); // (0,2), r = 1
        jacobi_rotate(a11, a22, a12, a01, a02, v1, v2); // (1,2), r = 0
    }
    l0 = a00;
    l1 = a11;
    l2 = a22;
    // sort descending; a swap followed by one negation keeps det(V) = +1
    auto swap_cols = [](R& la, R& lb, Vec3<R>& va, Vec3<R>& vb) {
        R const tl = la;
        la         = lb;
        lb         = tl;
        Vec3<R> const tv = va;
        va               = vb;
        vb               = {-tv.x, -tv.y, -tv.z};
    };
    if (l0 < l1)
        swap_cols(l0, l1, v0, v1);
    if (l0 < l2)
        swap_cols(l0, l2, v0, v2);
    if (l1 < l2)
        swap_cols(l1, l2, v1, v2);
}

// 2^-n for small n >= 0, through the exponent field
SBS_HD float pow2_neg(int n, float)
{
#ifdef __CUDA_ARCH__
    return __int_as_float((127 - n) << 23);
#else
    return 1.0f / static_cast<float>(1 << n);
#endif
}
SBS_HD double pow2_neg(int n, double) { return 1.0 / static_cast<double>(1 << n); }

template <typename R>
struct GreenOut
{
    Vec3<R> f1, f2, f3; // negative gradients at vertices 1..3 (f4 = -(f1+f2+f3))
    R C;                // |V0| * psi
};

// Newton steps for the polar rotation, chosen a priori from the strain: with eps = |E_G|_F every
// singular value of F lies within delta0 = 1 - sqrt(1 - 2 eps) of 1 (eps for stretch), and one step
// of X <- (X + X^-T)/2 maps an error delta to delta^2 / (2 (1 + delta)).
template <typename R>
struct PolarSteps;
template <>
struct PolarSteps<float>
{ // final error <= ~3e-7
    static SBS_HD int of(float e2n) { return e2n <= 5.9e-7f ? 1 : e2n <= 1.5e-3f ? 2 : e2n <= 6.7e-2f ? 3 : 4; }
    static constexpr float max_e2n = 2.25f; // beyond: general route
};
template <>
struct PolarSteps<double>
{ // two more steps square 3e-7 twice
    static SBS_HD int of(double e2n) { return e2n <= 5.9e-7 ? 3 : e2n <= 1.5e-3 ? 4 : e2n <= 6.7e-2 ? 5 : 6; }
    static constexpr double max_e2n = 2.25;
};

// how often the general route ran on this device (debug counter behind sbsb200_stats.green_general_calls)
#ifdef __CUDACC__
__device__ unsigned long long g_general_route_calls = 0;
#endif

// GENERAL route of green_gradients (clamp or inversion active): columns of P and psi from the
// eigen-decomposition of A = F^T F.  Kept out of line: it is rare, long, and register hungry.
template <typename R>
__host__ __device__ __noinline__ void green_general(Vec3<R> c0, Vec3<R> c1, Vec3<R> c2, R a00, R a11, R a22, R a01,
                                                    R a02, R a12, bool inverted, R mu, R lam, Vec3<R>& pk0,
                                                    Vec3<R>& pk1, Vec3<R>& pk2, R& psi)
{
#ifdef __CUDA_ARCH__
    atomicAdd(&g_general_route_calls, 1ull);
#endif
    R const smin = R(0.577);
    R l0, l1, l2;
    Vec3<R> v0, v1, v2;
    sym_eig3(a00, a11, a22, a01, a02, a12, l0, l1, l2, v0, v1, v2);

    // U' columns
    auto Fv = [&](Vec3<R> v) -> Vec3<R> {
        return {c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z,
                c0.z * v.x + c1.z * v.y + c2.z * v.z};
    };
    Vec3<R> u0 = Fv(v0);
    R n0       = dot(u0, u0);
    if (n0 > R(0))
    {
        R const s = rsqrt_(n0);
        u0        = {u0.x * s, u0.y * s, u0.z * s};
    }
    else
        u0 = {R(1), R(0), R(0)};
    Vec3<R> u1 = Fv(v1);
    {
        R const p = dot(u1, u0);
        u1        = {u1.x - p * u0.x, u1.y - p * u0.y, u1.z - p * u0.z};
    }
    R const n1 = dot(u1, u1);
    if (n1 > l0 * R(sizeof(R) == 4 ? 1e-12 : 1e-28))
    {
        R const s = rsqrt_(n1);
        u1        = {u1.x * s, u1.y * s, u1.z * s};
    }
    else
    { // rank <= 1: any unit vector orthogonal to u0
        Vec3<R> const ax = abs_(u0.x) < R(0.6) ? Vec3<R>{R(1), R(0), R(0)} : Vec3<R>{R(0), R(1), R(0)};
        u1               = cross(u0, ax);
        R const s        = rsqrt_(dot(u1, u1));
        u1               = {u1.x * s, u1.y * s, u1.z * s};
    }
    Vec3<R> const u2 = cross(u0, u1);

    // clamped principal stretches (:92-102)
    R const s0 = max_(sqrt_(max_(l0, R(0))), smin);
    R const s1 = max_(sqrt_(max_(l1, R(0))), smin);
    R const s2 = inverted ? smin : max_(sqrt_(max_(l2, R(0))), smin);

    // Ehat, Piolahat (:104-106)
    R const eh0 = R(0.5) * (s0 * s0 - R(1)), eh1 = R(0.5) * (s1 * s1 - R(1)), eh2 = R(0.5) * (s2 * s2 - R(1));
    R const ehtr = eh0 + eh1 + eh2;
    R const ph0 = s0 * (R(2) * mu * eh0 + lam * ehtr), ph1 = s1 * (R(2) * mu * eh1 + lam * ehtr),
            ph2 = s2 * (R(2) * mu * eh2 + lam * ehtr);

    // psi from E = U Ehat V^T (:108-110): |E|_F^2 = sum ehat_i^2, tr E = sum ehat_i (u_i . v_i)
    R const Etr = eh0 * dot(u0, v0) + eh1 * dot(u1, v1) + eh2 * dot(u2, v2);
    psi         = mu * (eh0 * eh0 + eh1 * eh1 + eh2 * eh2) + R(0.5) * lam * Etr * Etr;

    // P = U Piolahat V^T (:112): P[r][k] = sum_i ph_i u_i[r] v_i[k]
    Vec3<R> const w0 = {ph0 * u0.x, ph0 * u0.y, ph0 * u0.z};
    Vec3<R> const w1 = {ph1 * u1.x, ph1 * u1.y, ph1 * u1.z};
    Vec3<R> const w2 = {ph2 * u2.x, ph2 * u2.y, ph2 * u2.z};
    pk0 = {w0.x * v0.x + w1.x * v1.x + w2.x * v2.x, w0.y * v0.x + w1.y * v1.x + w2.y * v2.x,
           w0.z * v0.x + w1.z * v1.x + w2.z * v2.x};
    pk1 = {w0.x * v0.y + w1.x * v1.y + w2.x * v2.y, w0.y * v0.y + w1.y * v1.y + w2.y * v2.y,
           w0.z * v0.y + w1.z * v1.y + w2.z * v2.y};
    pk2 = {w0.x * v0.z + w1.x * v1.z + w2.x * v2.z, w0.y * v0.z + w1.y * v1.z + w2.y * v2.z,
           w0.z * v0.z + w1.z * v1.z + w2.z * v2.z};
}

// Steps 2-9 of green_constraint_t::project_positions (green_constraint.cpp:61-120) for one
// tet: D = DmInv (row-major d00..d22), V0s = signed rest volume.
//
// Two algebraically equivalent routes to (P, psi), see DESIGN.md "Green projection without SVD":
//
//  FAST (tet not inverted and every singular value > 0.577, i.e. the clamp at :99-102 and the
//  flip at :92-96 are both inactive — the overwhelmingly common case).  With A = F^T F,
//  E_G = (A - I)/2 and R = polar rotation of F:
//      P    = U Phat V^T = F (2 mu E_G + lam tr(E_G) I)
//      |E|_F^2 = |E_G|_F^2,   tr E = tr(U Ehat V^T) = tr(R E_G)      (NOT tr(E_G): :108-110)
//  "every sigma > 0.577" <=> A - 0.577^2 I positive definite (three leading minors), and R
//  comes from the Newton iteration R <- (R + R^-T)/2 with the step count fixed a priori from
//  |E_G|_F (no convergence test in the loop).
//
//  "Inverted" (:61-65) is sign(det Ds) != sign(V0) with zero counted positive; since
//  det DmInv has the sign of V0 this is det F < 0, up to the zero cases — and those only matter
//  when sigma_3 = 0, where the general route clamps sigma_3 to 0.577 whatever the flag says.  det F
//  is the first cofactor expansion of the Newton iteration, so the test is free.
//
//  GENERAL: green_general above.  U' V^T is always a proper rotation there, which is exactly what
//  "if inverted: sigma3 -> -sigma3, U.col(2) -> -U.col(2)" produces; the flipped sigma3 is
//  negative and therefore always clamped to 0.577.
template <typename R>
SBS_HD GreenOut<R> green_gradients(Vec3<R> x1, Vec3<R> x2, Vec3<R> x3, Vec3<R> x4, R d00, R d01, R d02, R d10,
                                   R d11, R d12, R d20, R d21, R d22, R V0s, R mu, R lam)
{
    Vec3<R> const e1 = x1 - x4, e2 = x2 - x4, e3 = x3 - x4; // columns of Ds (:69-72)

    // F = Ds * DmInv, stored by columns c0 c1 c2 (:74)
    Vec3<R> const c0 = {e1.x * d00 + e2.x * d10 + e3.x * d20, e1.y * d00 + e2.y * d10 + e3.y * d20,
                        e1.z * d00 + e2.z * d10 + e3.z * d20};
    Vec3<R> const c1 = {e1.x * d01 + e2.x * d11 + e3.x * d21, e1.y * d01 + e2.y * d11 + e3.y * d21,
                        e1.z * d01 + e2.z * d11 + e3.z * d21};
    Vec3<R> const c2 = {e1.x * d02 + e2.x * d12 + e3.x * d22, e1.y * d02 + e2.y * d12 + e3.y * d22,
                        e1.z * d02 + e2.z * d12 + e3.z * d22};
    R const a00 = dot(c0, c0), a11 = dot(c1, c1), a22 = dot(c2, c2), a01 = dot(c0, c1), a02 = dot(c0, c2),
            a12 = dot(c1, c2);
    // first cofactors of F: det F, and the first Newton step
    Vec3<R> n0 = cross(c1, c2), n1 = cross(c2, c0), n2 = cross(c0, c1);
    R det      = dot(c0, n0);
    bool const inverted = det < R(0);

    R psi;
    R const nv = -abs_(V0s);
    GreenOut<R> o;

    R const g00 = R(0.5) * a00 - R(0.5), g11 = R(0.5) * a11 - R(0.5), g22 = R(0.5) * a22 - R(0.5);
    R const g01 = R(0.5) * a01, g02 = R(0.5) * a02, g12 = R(0.5) * a12;
    R const e2n = g00 * g00 + g11 * g11 + g22 * g22 + R(2) * (g01 * g01 + g02 * g02 + g12 * g12);
    // All singular values above the clamp <=> B = A - smin^2 I is positive definite.  Every eigenvalue of A is
    // at least 1 - 2 |E_G|_F, so |E_G|_F^2 < 0.1112 (strains below a third) settles it without the minors;
    // beyond that the three leading minors of B decide.
    bool fast = !inverted && e2n <= PolarSteps<R>::max_e2n;
    if (fast && e2n >= R(0.1112))
    {
        R const smin = R(0.577);
        R const b00 = a00 - smin * smin, b11 = a11 - smin * smin, b22 = a22 - smin * smin;
        R const m2  = b00 * b11 - a01 * a01;
        R const m3  = b22 * m2 - a02 * (a02 * b11 - a01 * a12) + a12 * (a02 * a01 - b00 * a12);
        fast        = b00 > R(0) && m2 > R(0) && m3 > R(0);
    }
    if (fast)
    {
        R const trg = g00 + g11 + g22;
        // H = -|V0| P DmInv^T = F N with N = (-|V0| M) DmInv^T, M = 2 mu E_G + lam tr(E_G) I  (:112-116)
        R const tm  = R(2) * mu * nv;
        R const lt  = lam * trg * nv;
        R const m00 = tm * g00 + lt, m11 = tm * g11 + lt, m22 = tm * g22 + lt;
        R const m01 = tm * g01, m02 = tm * g02, m12 = tm * g12;
        // N[k][c] = sum_j M[k][j] DmInv[c][j]
        R const n00 = m00 * d00 + m01 * d01 + m02 * d02, n01 = m00 * d10 + m01 * d11 + m02 * d12,
                n02 = m00 * d20 + m01 * d21 + m02 * d22;
        R const n10 = m01 * d00 + m11 * d01 + m12 * d02, n11 = m01 * d10 + m11 * d11 + m12 * d12,
                n12 = m01 * d20 + m11 * d21 + m12 * d22;
        R const n20 = m02 * d00 + m12 * d01 + m22 * d02, n21 = m02 * d10 + m12 * d11 + m22 * d12,
                n22 = m02 * d20 + m12 * d21 + m22 * d22;
        o.f1 = {c0.x * n00 + c1.x * n10 + c2.x * n20, c0.y * n00 + c1.y * n10 + c2.y * n20,
                c0.z * n00 + c1.z * n10 + c2.z * n20};
        o.f2 = {c0.x * n01 + c1.x * n11 + c2.x * n21, c0.y * n01 + c1.y * n11 + c2.y * n21,
                c0.z * n01 + c1.z * n11 + c2.z * n21};
        o.f3 = {c0.x * n02 + c1.x * n12 + c2.x * n22, c0.y * n02 + c1.y * n12 + c2.y * n22,
                c0.z * n02 + c1.z * n12 + c2.z * n22};
        // Polar rotation by Newton, X <- (X + X^-T) / 2, carried as X = sigma U with sigma = 2^-it:
        // U <- U + cof(U) / (sigma^2 det U) costs one fused multiply-add per entry and no scaling of U.
        Vec3<R> q0 = c0, q1 = c1, q2 = c2;
        int const steps = PolarSteps<R>::of(e2n);
        R inv_sigma2    = R(1); // 4^it
#pragma unroll 1
        for (int it = 0;;)
        {
            R const h = div_(inv_sigma2, det);
            q0        = {q0.x + h * n0.x, q0.y + h * n0.y, q0.z + h * n0.z};
            q1        = {q1.x + h * n1.x, q1.y + h * n1.y, q1.z + h * n1.z};
            q2        = {q2.x + h * n2.x, q2.y + h * n2.y, q2.z + h * n2.z};
            inv_sigma2 *= R(4);
            if (++it >= steps)
                break;
            n0  = cross(q1, q2);
            n1  = cross(q2, q0);
            n2  = cross(q0, q1);
            det = dot(q0, n0);
        }
        // tr(R E_G) with R[r][k] = sigma q_k[r], sigma = 2^-steps = 1 / sqrt(inv_sigma2)
        R const sigma = pow2_neg(steps, R(0));
        R const Etr = sigma * (q0.x * g00 + q0.y * g01 + q0.z * g02 + q1.x * g01 + q1.y * g11 + q1.z * g12 +
                               q2.x * g02 + q2.y * g12 + q2.z * g22);
        psi = mu * e2n + R(0.5) * lam * Etr * Etr;
    }
    else
    {
        Vec3<R> pk0, pk1, pk2; // columns of P
        green_general<R>(c0, c1, c2, a00, a11, a22, a01, a02, a12, inverted, mu, lam, pk0, pk1, pk2, psi);
        // H = -|V0| P DmInv^T (:115-116): column c of H = -|V0| * sum_k P[:,k] DmInv[c][k]
        o.f1 = {nv * (pk0.x * d00 + pk1.x * d01 + pk2.x * d02), nv * (pk0.y * d00 + pk1.y * d01 + pk2.y * d02),
                nv * (pk0.z * d00 + pk1.z * d01 + pk2.z * d02)};
        o.f2 = {nv * (pk0.x * d10 + pk1.x * d11 + pk2.x * d12), nv * (pk0.y * d10 + pk1.y * d11 + pk2.y * d12),
                nv * (pk0.z * d10 + pk1.z * d11 + pk2.z * d12)};
        o.f3 = {nv * (pk0.x * d20 + pk1.x * d21 + pk2.x * d22), nv * (pk0.y * d20 + pk1.y * d21 + pk2.y * d22),
                nv * (pk0.z * d20 + pk1.z * d21 + pk2.z * d22)};
    }
    o.C  = abs_(V0s) * psi; // :133
    return o;
}

} // namespace sbsb200
