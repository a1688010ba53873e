// Green projection of TWO independent tets at once on packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2).
//
// Why: a 3-register FFMA occupies the fma pipe of an SM sub-partition for two cycles, the packed form
// does two of them in the same slot.  The projection (xpbd_math.cuh) is ~300 fma-pipe instructions of
// ~390, and the resident kernel is bound by exactly that pipe on the sub-partitions that hold two
// warps (profiles/r01_summary.md).  Lane 0 and lane 1 of every pair belong to two tets of two different
// clusters of one colour (they share no vertex), so no swizzle is ever needed: the algebra below is the
// scalar fast route of green_gradients<float> / green_project_at<float, false> written once on pairs.
// Either lane leaving the fast route (inverted tet, clamp active, huge strain) sends both through the
// scalar code (by the caller).  Device only; measured by tools/bench_math.cu before it goes into the kernels.
#pragma once

#include "../soft-body-simulator_b200/csrc/xpbd_kernels.cuh"

namespace sbsb200 {

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000

struct P2
{
    float2 v;
    P2() = default;
    __device__ __forceinline__ P2(float2 a) : v(a) {}
    __device__ __forceinline__ P2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ P2(float a, float b) : v(make_float2(a, b)) {}
};
// a product not yet rounded: the next + or - turns it into one FFMA2 (the compiler never contracts
// the _rn intrinsics by itself)
struct P2Prod
{
    float2 a, b;
    __device__ __forceinline__ operator P2() const { return P2(__fmul2_rn(a, b)); }
};
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P2Prod operator*(P2 a, P2 b) { return {a.v, b.v}; }
__device__ __forceinline__ P2Prod operator*(P2Prod p, P2 c) { return {__fmul2_rn(p.a, p.b), c.v}; }
__device__ __forceinline__ P2Prod operator*(P2 c, P2Prod p) { return {c.v, __fmul2_rn(p.a, p.b)}; }
__device__ __forceinline__ P2 operator+(P2 a, P2 b) { return P2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ P2 operator-(P2 a, P2 b) { return P2(__fadd2_rn(a.v, neg2(b.v))); }
__device__ __forceinline__ P2 operator-(P2 a) { return P2(neg2(a.v)); }
__device__ __forceinline__ P2 operator+(P2Prod p, P2 c) { return P2(__ffma2_rn(p.a, p.b, c.v)); }
__device__ __forceinline__ P2 operator+(P2 c, P2Prod p) { return P2(__ffma2_rn(p.a, p.b, c.v)); }
__device__ __forceinline__ P2 operator+(P2Prod p, P2Prod q) { return P2(__ffma2_rn(p.a, p.b, __fmul2_rn(q.a, q.b))); }
__device__ __forceinline__ P2 operator-(P2Prod p, P2 c) { return P2(__ffma2_rn(p.a, p.b, neg2(c.v))); }
__device__ __forceinline__ P2 operator-(P2 c, P2Prod p) { return P2(__ffma2_rn(neg2(p.a), p.b, c.v)); }
__device__ __forceinline__ P2 operator-(P2Prod p, P2Prod q) { return P2(__ffma2_rn(neg2(q.a), q.b, __fmul2_rn(p.a, p.b))); }
__device__ __forceinline__ P2& operator+=(P2& a, P2Prod p) { return a = a + p; }
__device__ __forceinline__ P2& operator+=(P2& a, P2 b) { return a = a + b; }
__device__ __forceinline__ P2 dot(Vec3<P2> a, Vec3<P2> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3<P2> cross(Vec3<P2> a, Vec3<P2> b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ P2 abs2(P2 a) { return P2(fabsf(a.v.x), fabsf(a.v.y)); }

struct Real4Pair // two (x, y, z, w) records, component-wise
{
    P2 x, y, z, w;
};
__device__ __forceinline__ Real4Pair pair_of(Real4<float> a, Real4<float> b)
{
    return {P2(a.x, b.x), P2(a.y, b.y), P2(a.z, b.z), P2(a.w, b.w)};
}
__device__ __forceinline__ Real4<float> lane0(Real4Pair const& p) { return {p.x.v.x, p.y.v.x, p.z.v.x, p.w.v.x}; }
__device__ __forceinline__ Real4<float> lane1(Real4Pair const& p) { return {p.x.v.y, p.y.v.y, p.z.v.y, p.w.v.y}; }

// One Green projection of two tets.  p1..p4: their vertices (xi, w); r0..r2: their (DmInv, V0, material)
// records; mu, lam, at (= alpha / dt^2): per lane; lambda: per lane, updated.  Same steps as
// green_project_at<float, false> (green_constraint.cpp:61-157), no damping.  Returns false, with nothing
// changed, when a lane is off the fast route: the caller then runs green_project_at on each lane (it
// holds the unpacked operands anyway).
__device__ __forceinline__ bool green_project_pair(Real4Pair& p1, Real4Pair& p2, Real4Pair& p3, Real4Pair& p4,
                                                   Real4Pair const& r0, Real4Pair const& r1, Real4Pair const& r2, P2 mu,
                                                   P2 lam, P2 at, float dt, P2& lambda)
{
    typedef P2 R;
    Vec3<R> const x1 = {p1.x, p1.y, p1.z}, x2 = {p2.x, p2.y, p2.z}, x3 = {p3.x, p3.y, p3.z}, x4 = {p4.x, p4.y, p4.z};
    R const d00 = r0.x, d01 = r0.y, d02 = r0.z, d10 = r0.w, d11 = r1.x, d12 = r1.y, d20 = r1.z, d21 = r1.w, d22 = r2.x,
            V0s = r2.y;
    Vec3<R> const e1 = x1 - x4, e2 = x2 - x4, e3 = x3 - x4;
    Vec3<R> const c0 = {e1.x * d00 + e2.x * d10 + e3.x * d20, e1.y * d00 + e2.y * d10 + e3.y * d20,
                        e1.z * d00 + e2.z * d10 + e3.z * d20};
    Vec3<R> const c1 = {e1.x * d01 + e2.x * d11 + e3.x * d21, e1.y * d01 + e2.y * d11 + e3.y * d21,
                        e1.z * d01 + e2.z * d11 + e3.z * d21};
    Vec3<R> const c2 = {e1.x * d02 + e2.x * d12 + e3.x * d22, e1.y * d02 + e2.y * d12 + e3.y * d22,
                        e1.z * d02 + e2.z * d12 + e3.z * d22};
    R const a00 = dot(c0, c0), a11 = dot(c1, c1), a22 = dot(c2, c2), a01 = dot(c0, c1), a02 = dot(c0, c2),
            a12 = dot(c1, c2);
    Vec3<R> n0 = cross(c1, c2), n1 = cross(c2, c0), n2 = cross(c0, c1);
    R det      = dot(c0, n0);
    float const smin = 0.577f;
    R const b00 = a00 - R(smin * smin), b11 = a11 - R(smin * smin), b22 = a22 - R(smin * smin);
    R const m2  = b00 * b11 - a01 * a01;
    R const m3  = b22 * m2 - a02 * (a02 * b11 - a01 * a12) + a12 * (a02 * a01 - b00 * a12);
    R const g00 = R(0.5f) * a00 - R(0.5f), g11 = R(0.5f) * a11 - R(0.5f), g22 = R(0.5f) * a22 - R(0.5f);
    R const g01 = R(0.5f) * a01, g02 = R(0.5f) * a02, g12 = R(0.5f) * a12;
    R const e2n = g00 * g00 + g11 * g11 + g22 * g22 + R(2.f) * (g01 * g01 + g02 * g02 + g12 * g12);
    auto const fast = [&](int l) {
        float const de = l ? det.v.y : det.v.x, b = l ? b00.v.y : b00.v.x, mm2 = l ? m2.v.y : m2.v.x,
                    mm3 = l ? m3.v.y : m3.v.x, e = l ? e2n.v.y : e2n.v.x;
        return !(de < 0.f) && b > 0.f && mm2 > 0.f && mm3 > 0.f && e <= PolarSteps<float>::max_e2n;
    };
    if (!(fast(0) && fast(1)))
        return false;
    R const trg = g00 + g11 + g22;
    R const tm  = R(2.f) * mu;
    R const lt  = lam * trg;
    R const m00 = tm * g00 + lt, m11 = tm * g11 + lt, m22 = tm * g22 + lt;
    R const m01 = tm * g01, m02 = tm * g02, m12 = tm * g12;
    Vec3<R> const pk0 = {c0.x * m00 + c1.x * m01 + c2.x * m02, c0.y * m00 + c1.y * m01 + c2.y * m02,
                         c0.z * m00 + c1.z * m01 + c2.z * m02};
    Vec3<R> const pk1 = {c0.x * m01 + c1.x * m11 + c2.x * m12, c0.y * m01 + c1.y * m11 + c2.y * m12,
                         c0.z * m01 + c1.z * m11 + c2.z * m12};
    Vec3<R> const pk2 = {c0.x * m02 + c1.x * m12 + c2.x * m22, c0.y * m02 + c1.y * m12 + c2.y * m22,
                         c0.z * m02 + c1.z * m12 + c2.z * m22};
    // polar rotation: as many Newton steps as the more strained lane needs (a step on a converged
    // rotation leaves it where it is)
    Vec3<R> q0 = c0, q1 = c1, q2 = c2;
    int const steps = max(PolarSteps<float>::of(e2n.v.x), PolarSteps<float>::of(e2n.v.y));
#pragma unroll 1
    for (int it = 0;;)
    {
        R const h = P2(__fdividef(0.5f, det.v.x), __fdividef(0.5f, det.v.y));
        q0        = {R(0.5f) * q0.x + h * n0.x, R(0.5f) * q0.y + h * n0.y, R(0.5f) * q0.z + h * n0.z};
        q1        = {R(0.5f) * q1.x + h * n1.x, R(0.5f) * q1.y + h * n1.y, R(0.5f) * q1.z + h * n1.z};
        q2        = {R(0.5f) * q2.x + h * n2.x, R(0.5f) * q2.y + h * n2.y, R(0.5f) * q2.z + h * n2.z};
        if (++it >= steps)
            break;
        n0  = cross(q1, q2);
        n1  = cross(q2, q0);
        n2  = cross(q0, q1);
        det = dot(q0, n0);
    }
    R const Etr = q0.x * g00 + q0.y * g01 + q0.z * g02 + q1.x * g01 + q1.y * g11 + q1.z * g12 + q2.x * g02 +
                  q2.y * g12 + q2.z * g22;
    R const psi = mu * e2n + R(0.5f) * lam * Etr * Etr;
    R const av  = abs2(V0s);
    R const nv  = -av;
    Vec3<R> const f1 = {nv * (pk0.x * d00 + pk1.x * d01 + pk2.x * d02), nv * (pk0.y * d00 + pk1.y * d01 + pk2.y * d02),
                        nv * (pk0.z * d00 + pk1.z * d01 + pk2.z * d02)};
    Vec3<R> const f2 = {nv * (pk0.x * d10 + pk1.x * d11 + pk2.x * d12), nv * (pk0.y * d10 + pk1.y * d11 + pk2.y * d12),
                        nv * (pk0.z * d10 + pk1.z * d11 + pk2.z * d12)};
    Vec3<R> const f3 = {nv * (pk0.x * d20 + pk1.x * d21 + pk2.x * d22), nv * (pk0.y * d20 + pk1.y * d21 + pk2.y * d22),
                        nv * (pk0.z * d20 + pk1.z * d21 + pk2.z * d22)};
    R const C        = av * psi;
    Vec3<R> const f4 = {-(f1.x + f2.x + f3.x), -(f1.y + f2.y + f3.y), -(f1.z + f2.z + f3.z)};
    R const S = p1.w * dot(f1, f1) + p2.w * dot(f2, f2) + p3.w * dot(f3, f3) + p4.w * dot(f4, f4);
    R const num = -(C + at * lambda);
    R const den = S + at;
    // S < 1e-20: that lane's projection is skipped (green_constraint.cpp:67, :130-131)
    R const dl = P2(S.v.x < 1e-20f ? 0.f : __fdividef(num.v.x, den.v.x), S.v.y < 1e-20f ? 0.f : __fdividef(num.v.y, den.v.y));
    lambda     = lambda + dl;
    R const k1 = -(p1.w * dl), k2 = -(p2.w * dl), k3 = -(p3.w * dl), k4 = -(p4.w * dl);
    p1.x += k1 * f1.x; p1.y += k1 * f1.y; p1.z += k1 * f1.z;
    p2.x += k2 * f2.x; p2.y += k2 * f2.y; p2.z += k2 * f2.z;
    p3.x += k3 * f3.x; p3.y += k3 * f3.y; p3.z += k3 * f3.z;
    p4.x += k4 * f4.x; p4.y += k4 * f4.y; p4.z += k4 * f4.z;
    return true;
}

#endif // __CUDA_ARCH__ >= 1000

} // namespace sbsb200
