/*
 * xpbd_oracle.c — CPU oracle (fp64, serial) for the XPBD hot path of "sbs".
 *
 * TEST INFRASTRUCTURE ONLY (see xpbd_oracle.h).  Restates, in plain C, the algorithm of
 *   src/physics/timestep.cpp:20-70            frame / substep driver
 *   src/physics/gauss_seidel_solver.cpp:8-37  serial Gauss-Seidel sweep
 *   src/physics/constraint.cpp:12-16          lambda reset
 *   src/physics/xpbd/green_constraint.cpp:11-173
 *   src/physics/xpbd/distance_constraint.cpp:8-53
 *   src/physics/xpbd/collision_constraint.cpp:9-48
 *   src/physics/particle.cpp:6-53
 *   src/physics/collision/brute_force_cd_system.cpp:14-26, bvh_model.cpp:30-100,
 *   sdf_model.cpp:52-75, xpbd/contact_handler.cpp:14-54
 *   src/physics/tetrahedral_mesh_boundary.cpp:65-120, topology.cpp:335-342,:904-956
 *   src/geometry/get_simple_bar_model.cpp:6-122
 *
 * Parity pinning.  The reference has no tests and no numeric golden vectors for this path
 * (SURVEY.md §4, §8c).  What pins this file:
 *   (1) the mesh generator + boundary code are checked against the reference's own PLY
 *       fixtures (data/meshes/cube_tet.ply, tet_bar_5x2x2.ply, bar_tet.ply), copied as
 *       index/position fixtures under tests/golden/;
 *   (2) the whole step is checked against oracle/_ref — the reference's OWN physics
 *       translation units compiled unmodified against a linear-algebra shim (Eigen and
 *       Discregrid are absent from the reference tree and from this image), see
 *       oracle/build_ref.sh; golden vectors generated from it live in tests/golden/;
 *   (3) the 3x3 SVD is checked against numpy.linalg.svd.
 * Third-party arithmetic that is NOT in the reference tree: Eigen (>=3.3, unpinned) for
 * JacobiSVD/inverse/determinant and Discregrid (Q-Minh fork @master, unpinned) for the
 * KD-tree.  The Green projection depends on the SVD only through U*f(Sigma)*V^T, which is
 * independent of the SVD's sign/ordering freedom, so any accurate SVD reproduces it.  The
 * KD-tree only culls; the contact set equals {surface vertex : sdf < 0} whenever bodies
 * stay inside the SDF's englobing volume, which all graded scenes guarantee.
 */
#include "xpbd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small dense helpers (row-major 3x3)
 * ---------------------------------------------------------------------------------------- */
static void m3_mul(const double A[9], const double B[9], double C[9])
{
    double T[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            T[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
    memcpy(C, T, sizeof T);
}
static void m3_transpose(const double A[9], double T[9])
{
    double B[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(T, B, sizeof B);
}
static double m3_det(const double A[9])
{
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}
static void m3_inverse(const double A[9], double Inv[9])
{
    double const d  = m3_det(A);
    double const id = 1.0 / d;
    double B[9];
    B[0] = (A[4] * A[8] - A[5] * A[7]) * id;
    B[1] = (A[2] * A[7] - A[1] * A[8]) * id;
    B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = (A[5] * A[6] - A[3] * A[8]) * id;
    B[4] = (A[0] * A[8] - A[2] * A[6]) * id;
    B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = (A[3] * A[7] - A[4] * A[6]) * id;
    B[7] = (A[1] * A[6] - A[0] * A[7]) * id;
    B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    memcpy(Inv, B, sizeof B);
}

/* ------------------------------------------------------------------------------------------
 * 3x3 SVD: two-sided Jacobi (the published algorithm behind Eigen::JacobiSVD for square
 * inputs: per pivot pair, a rotation that symmetrises the 2x2 block followed by a symmetric
 * Jacobi rotation; sweeps until no pivot exceeds the threshold; then sign fix + sort).
 * ---------------------------------------------------------------------------------------- */
/* columns (p,q) of M <- columns (p,q) of M times G (2x2, row-major) */
static void cols_times(double M[9], int p, int q, const double G[4])
{
    for (int k = 0; k < 3; ++k)
    {
        double const a = M[3 * k + p], b = M[3 * k + q];
        M[3 * k + p] = a * G[0] + b * G[2];
        M[3 * k + q] = a * G[1] + b * G[3];
    }
}
/* rows (p,q) of M <- G^T times rows (p,q) of M */
static void rows_timesT(double M[9], int p, int q, const double G[4])
{
    for (int k = 0; k < 3; ++k)
    {
        double const a = M[3 * p + k], b = M[3 * q + k];
        M[3 * p + k] = G[0] * a + G[2] * b;
        M[3 * q + k] = G[1] * a + G[3] * b;
    }
}

void orc_svd3(const double F[9], double U[9], double sigma[3], double V[9])
{
    double W[9];
    double scale = 0.0;
    for (int i = 0; i < 9; ++i)
        if (fabs(F[i]) > scale)
            scale = fabs(F[i]);
    if (scale == 0.0)
        scale = 1.0;
    for (int i = 0; i < 9; ++i)
        W[i] = F[i] / scale;
    double const I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(U, I3, sizeof I3);
    memcpy(V, I3, sizeof I3);

    /* invariant: F/scale = U W V^T */
    double const precision = 2.0 * 2.220446049250313e-16;
    double const tiny      = 2.2250738585072014e-308;
    for (int sweep = 0; sweep < 64; ++sweep)
    {
        int finished = 1;
        for (int q = 1; q < 3; ++q)
            for (int p = 0; p < q; ++p)
            {
                double const maxdiag = fmax(fabs(W[4 * p]), fabs(W[4 * q]));
                double const thr     = fmax(tiny, precision * maxdiag);
                if (fabs(W[3 * p + q]) <= thr && fabs(W[3 * q + p]) <= thr)
                    continue;
                finished = 0;
                double const a = W[4 * p], b = W[3 * p + q], c = W[3 * q + p], d = W[4 * q];
                /* R = [[c1, s1],[-s1, c1]] with R^T B symmetric: tan = (b - c)/(a + d) */
                double c1 = 1.0, s1 = 0.0;
                double const h = hypot(a + d, b - c);
                if (h > tiny)
                {
                    c1 = (a + d) / h;
                    s1 = (b - c) / h;
                }
                double const x = c1 * a - s1 * c, y = c1 * b - s1 * d, z = s1 * b + c1 * d;
                /* J = [[cj, sj],[-sj, cj]] with J^T [[x,y],[y,z]] J diagonal */
                double cj = 1.0, sj = 0.0;
                if (fabs(y) > tiny)
                {
                    double const tau = (z - x) / (2.0 * y);
                    double const t =
                        tau >= 0 ? 1.0 / (tau + sqrt(1.0 + tau * tau)) : 1.0 / (tau - sqrt(1.0 + tau * tau));
                    cj = 1.0 / sqrt(1.0 + t * t);
                    sj = t * cj;
                }
                double const R[4] = {c1, s1, -s1, c1};
                double const J[4] = {cj, sj, -sj, cj};
                double const L[4] = {R[0] * J[0] + R[1] * J[2], R[0] * J[1] + R[1] * J[3],
                                     R[2] * J[0] + R[3] * J[2], R[2] * J[1] + R[3] * J[3]};
                rows_timesT(W, p, q, L); /* W <- L^T W */
                cols_times(W, p, q, J);  /* W <- W J   */
                cols_times(U, p, q, L);  /* U <- U L   */
                cols_times(V, p, q, J);  /* V <- V J   */
            }
        if (finished)
            break;
    }
    /* non-negative diagonal: flip U columns */
    for (int i = 0; i < 3; ++i)
    {
        double a = W[4 * i];
        if (a < 0)
        {
            a = -a;
            for (int k = 0; k < 3; ++k)
                U[3 * k + i] = -U[3 * k + i];
        }
        sigma[i] = a * scale;
    }
    /* sort descending, swapping columns of U and V */
    for (int i = 0; i < 3; ++i)
    {
        int best = i;
        for (int j = i + 1; j < 3; ++j)
            if (sigma[j] > sigma[best])
                best = j;
        if (best != i)
        {
            double t    = sigma[i];
            sigma[i]    = sigma[best];
            sigma[best] = t;
            for (int k = 0; k < 3; ++k)
            {
                t               = U[3 * k + i];
                U[3 * k + i]    = U[3 * k + best];
                U[3 * k + best] = t;
                t               = V[3 * k + i];
                V[3 * k + i]    = V[3 * k + best];
                V[3 * k + best] = t;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * get_simple_bar_model (src/geometry/get_simple_bar_model.cpp:6-122)
 * ---------------------------------------------------------------------------------------- */
void orc_bar_model(int W, int H, int D, float* pos, int32_t* idx)
{
    for (int i = 0; i < W; ++i)
        for (int j = 0; j < H; ++j)
            for (int k = 0; k < D; ++k)
            {
                size_t const row = ((size_t)i * H + j) * D + k; /* :17 */
                pos[3 * row]     = (float)i;
                pos[3 * row + 1] = (float)j;
                pos[3 * row + 2] = (float)k;
            }
    /* corner numbering p0..p7 of :44-51; odd cells (:61-87), even cells (:89-113) */
    static const int odd[5][4]  = {{1, 0, 5, 2}, {5, 2, 7, 6}, {7, 0, 5, 4}, {2, 0, 7, 3}, {5, 0, 7, 2}};
    static const int even[5][4] = {{3, 1, 4, 0}, {6, 1, 3, 2}, {4, 1, 6, 5}, {6, 3, 4, 7}, {3, 1, 6, 4}};
    for (int i = 0; i < W - 1; ++i)
        for (int j = 0; j < H - 1; ++j)
            for (int k = 0; k < D - 1; ++k)
            {
                int p[8];
                p[0] = (i * H + j) * D + k;
                p[1] = ((i + 1) * H + j) * D + k;
                p[2] = ((i + 1) * H + (j + 1)) * D + k;
                p[3] = (i * H + (j + 1)) * D + k;
                p[4] = (i * H + j) * D + (k + 1);
                p[5] = ((i + 1) * H + j) * D + (k + 1);
                p[6] = ((i + 1) * H + (j + 1)) * D + (k + 1);
                p[7] = (i * H + (j + 1)) * D + (k + 1);
                size_t const cell       = ((size_t)i * (H - 1) + j) * (D - 1) + k; /* :53 */
                const int(*tab)[4]      = ((i + j + k) % 2 == 1) ? odd : even;
                for (int t = 0; t < 5; ++t)
                    for (int a = 0; a < 4; ++a)
                        idx[(cell * 5 + t) * 4 + a] = p[tab[t][a]];
            }
}

/* ------------------------------------------------------------------------------------------
 * boundary surface
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
    uint32_t key[3]; /* sorted vertex ids: triangle_t::operator< / == (topology.cpp:234-254) */
    uint32_t v[3];   /* first-seen winding */
    uint32_t ntets;
    uint32_t used;
} tri_slot;

static uint64_t tri_hash(const uint32_t k[3])
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 3; ++i)
    {
        h ^= k[i];
        h *= 1099511628211ull;
        h ^= h >> 29;
    }
    return h;
}

int orc_boundary_surface(int nV, int nT, const uint32_t* tets, uint32_t* surf_to_tet,
                         uint32_t* tris, int* n_tris)
{
    size_t cap = 16;
    while (cap < (size_t)nT * 8 + 16)
        cap <<= 1;
    tri_slot* table   = (tri_slot*)calloc(cap, sizeof(tri_slot));
    uint32_t* order   = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nT * 4 + 4); /* fi -> slot */
    size_t n_triangles = 0;
    /* faces_copy order (topology.cpp:335-342): (v1,v2,v4) (v2,v3,v4) (v3,v1,v4) (v1,v3,v2) */
    static const int face[4][3] = {{0, 1, 3}, {1, 2, 3}, {2, 0, 3}, {0, 2, 1}};
    for (int t = 0; t < nT; ++t)
        for (int f = 0; f < 4; ++f)
        {
            uint32_t v[3], k[3];
            for (int a = 0; a < 3; ++a)
                k[a] = v[a] = tets[4 * t + face[f][a]];
            if (k[0] > k[1]) { uint32_t x = k[0]; k[0] = k[1]; k[1] = x; }
            if (k[1] > k[2]) { uint32_t x = k[1]; k[1] = k[2]; k[2] = x; }
            if (k[0] > k[1]) { uint32_t x = k[0]; k[0] = k[1]; k[1] = x; }
            size_t h = tri_hash(k) & (cap - 1);
            while (table[h].used &&
                   (table[h].key[0] != k[0] || table[h].key[1] != k[1] || table[h].key[2] != k[2]))
                h = (h + 1) & (cap - 1);
            if (!table[h].used)
            { /* add_triangle: new index = push_back position (topology.cpp:739-745) */
                table[h].used = 1;
                memcpy(table[h].key, k, sizeof k);
                memcpy(table[h].v, v, sizeof v);
                order[n_triangles++] = (uint32_t)h;
            }
            table[h].ntets++; /* create_triangle_to_tetrahedron_incidency */
        }
    /* extract_boundary_surface (tetrahedral_mesh_boundary.cpp:84-119) */
    uint32_t* tet_to_surf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nV > 0 ? nV : 1));
    for (int i = 0; i < nV; ++i)
        tet_to_surf[i] = 0xffffffffu;
    int nVs = 0, nTri = 0;
    for (size_t fi = 0; fi < n_triangles; ++fi)
    {
        tri_slot const* s = &table[order[fi]];
        if (s->ntets == 2u) /* :88 interior */
            continue;
        for (int j = 0; j < 3; ++j)
        {
            uint32_t const vi = s->v[j];
            if (tet_to_surf[vi] == 0xffffffffu)
            {
                tet_to_surf[vi]    = (uint32_t)nVs;
                surf_to_tet[nVs++] = vi;
            }
            if (tris)
                tris[3 * nTri + j] = tet_to_surf[vi];
        }
        ++nTri;
    }
    if (n_tris)
        *n_tris = nTri;
    free(tet_to_surf);
    free(order);
    free(table);
    return nVs;
}

/* ------------------------------------------------------------------------------------------
 * Green constraint
 * ---------------------------------------------------------------------------------------- */
void orc_green_rest_state(const double x0[12], double DmInv[9], double* V0)
{
    double Dm[9]; /* columns p1-p4, p2-p4, p3-p4 (green_constraint.cpp:38-41) */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            Dm[3 * r + c] = x0[3 * c + r] - x0[9 + r];
    m3_inverse(Dm, DmInv);           /* :43 */
    *V0 = (1. / 6.) * m3_det(Dm);    /* :44 */
}

static double signed_volume(const double* p1, const double* p2, const double* p3, const double* p4)
{ /* green_constraint.cpp:160-173 */
    double D[9];
    for (int r = 0; r < 3; ++r)
    {
        D[3 * r]     = p1[r] - p4[r];
        D[3 * r + 1] = p2[r] - p4[r];
        D[3 * r + 2] = p3[r] - p4[r];
    }
    return (1. / 6.) * m3_det(D);
}

/* core projection on 4 particle pointers */
static int green_project(double* xi[4], const double* xn[4], const double w[4],
                         const double DmInv[9], double V0s, double mu, double lam, double alpha,
                         double beta, double dt, double* lagrange, double* diag)
{
    double const Vsigned = signed_volume(xi[0], xi[1], xi[2], xi[3]); /* :61 */
    int const vpos = Vsigned >= 0., v0pos = V0s >= 0.;
    int const inverted = (vpos && !v0pos) || (!vpos && v0pos); /* :62-65 */

    double Ds[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            Ds[3 * r + c] = xi[c][r] - xi[3][r]; /* :69-72 */
    double F[9];
    m3_mul(Ds, DmInv, F); /* :74 */

    double U[9], V[9], sg[3];
    orc_svd3(F, U, sg, V); /* :81-90 */
    double fh[3] = {sg[0], sg[1], sg[2]};
    if (inverted)
    { /* :92-96 */
        fh[2] = -fh[2];
        for (int r = 0; r < 3; ++r)
            U[3 * r + 2] = -U[3 * r + 2];
    }
    double const smin = 0.577; /* :99-102 */
    for (int i = 0; i < 3; ++i)
        fh[i] = fh[i] > smin ? fh[i] : smin;

    double eh[3], ph[3], ehtr = 0.;
    for (int i = 0; i < 3; ++i)
    {
        eh[i] = 0.5 * (fh[i] * fh[i] - 1.); /* :104 */
        ehtr += eh[i];
    }
    for (int i = 0; i < 3; ++i)
        ph[i] = fh[i] * (2. * mu * eh[i] + lam * ehtr); /* :106 */

    double Vt[9], Eh[9] = {eh[0], 0, 0, 0, eh[1], 0, 0, 0, eh[2]};
    double Ph[9] = {ph[0], 0, 0, 0, ph[1], 0, 0, 0, ph[2]};
    m3_transpose(V, Vt);
    double E[9], P[9], T[9];
    m3_mul(U, Eh, T);
    m3_mul(T, Vt, E); /* :108  E = U Ehat V^T */
    double const Etr = E[0] + E[4] + E[8];
    double e2 = 0.;
    for (int i = 0; i < 9; ++i)
        e2 += E[i] * E[i];
    double const psi = mu * e2 + 0.5 * lam * Etr * Etr; /* :110 */
    m3_mul(U, Ph, T);
    m3_mul(T, Vt, P); /* :112 */

    double const V0 = fabs(V0s); /* :115 */
    double DmInvT[9], Hm[9];
    m3_transpose(DmInv, DmInvT);
    m3_mul(P, DmInvT, Hm);
    for (int i = 0; i < 9; ++i)
        Hm[i] *= -V0; /* :116 */
    double f[4][3];
    for (int r = 0; r < 3; ++r)
    {
        f[0][r] = Hm[3 * r];
        f[1][r] = Hm[3 * r + 1];
        f[2][r] = Hm[3 * r + 2];
        f[3][r] = -(f[0][r] + f[1][r] + f[2][r]); /* :117-120 */
    }
    double S = 0.;
    for (int a = 0; a < 4; ++a)
        S += w[a] * (f[a][0] * f[a][0] + f[a][1] * f[a][1] + f[a][2] * f[a][2]); /* :123-127 */
    if (diag)
    {
        diag[0] = sg[0];
        diag[1] = sg[1];
        diag[2] = sg[2];
        diag[3] = V0 * psi;
        diag[4] = S;
        diag[5] = 0.;
        diag[6] = (double)inverted;
    }
    if (S < 1e-20) /* :67, :130-131 */
        return 0;

    double const C   = V0 * psi; /* :133 */
    double const dt2 = dt * dt;
    double const at  = alpha / dt2;
    double const bt  = beta * dt2;
    double const gam = at * bt / dt; /* :134-137 */
    double g = 0.;
    for (int a = 0; a < 4; ++a)
        for (int r = 0; r < 3; ++r)
            g += f[a][r] * (xi[a][r] - xn[a][r]); /* :140-144 */
    double const num = -(C + at * (*lagrange)) + gam * g; /* :147-148 */
    double const den = (1. + gam) * S + at;               /* :149 */
    double const dl  = num / den;
    *lagrange += dl; /* :152 */
    for (int a = 0; a < 4; ++a)
        for (int r = 0; r < 3; ++r)
            xi[a][r] += w[a] * -f[a][r] * dl; /* :154-157 */
    if (diag)
        diag[5] = dl;
    return 1;
}

int orc_green_project(double xi[12], const double xn[12], const double w[4], const double DmInv[9],
                      double V0, double young, double poisson, double alpha, double beta,
                      double dt, double* lagrange, double* diag)
{
    double* pxi[4]       = {xi, xi + 3, xi + 6, xi + 9};
    const double* pxn[4] = {xn, xn + 3, xn + 6, xn + 9};
    double const mu  = young / (2. * (1 + poisson));                            /* :45 */
    double const lam = (young * poisson) / ((1 + poisson) * (1 - 2 * poisson)); /* :46 */
    return green_project(pxi, pxn, w, DmInv, V0, mu, lam, alpha, beta, dt, lagrange, diag);
}

/* ------------------------------------------------------------------------------------------
 * world
 * ---------------------------------------------------------------------------------------- */
typedef struct
{ /* particle_t (include/sbs/physics/particle.h:10-53) */
    double x0[3], xi[3], xn[3], x[3], v[3], f[3], m;
} particle;

static double invmass(const particle* p) { return p->m > 0. ? 1. / p->m : 0.; } /* particle.cpp:39-44 */

enum { C_GREEN = 0, C_DISTANCE = 1 };
typedef struct
{
    int type;
    double alpha, beta, lagrange; /* constraint_t (constraint.h:30-34) */
    int b1, b2;
    uint32_t v[4];
    double DmInv[9], V0, mu, lam; /* green */
    double d;                     /* distance rest length */
} constraint;

typedef struct
{ /* collision_constraint_t (xpbd/collision_constraint.h) */
    double alpha, lagrange;
    int b;
    uint32_t v;
    double qs[3], n[3];
    int sdf_body;
} coll_constraint;

enum { B_TET = 0, B_SDF = 1 };
enum { SDF_PLANE = 0, SDF_SPHERE = 1, SDF_BOX = 2, SDF_GRID = 3 };
typedef struct
{
    int kind;
    /* tet body */
    int nV;
    particle* p;
    int nVs;
    uint32_t* surf_to_tet; /* tetrahedral_mesh_boundary_t::vertex_index_map_ */
    double* surf_pos;      /* visual-model vertex positions: what detection sees */
    /* sdf body */
    int sdf_kind;
    double a[3], b[3], r; /* plane: a=normal, r=offset; sphere: a=centre, r; box: a=min,b=max; grid: domain a..b */
    double volume[6];
    int excluded;         /* not handed to the cd system (main.cpp:77-86 lists the models that take part) */
    uint32_t grid_n[3];   /* grid: cells per axis */
    double* grid_nodes;   /* grid: node values, orc_grid_node_count(grid_n) of them */
} body;

struct orc_world
{
    int nb, capb;
    body* bodies;
    int nc, capc;
    constraint* cons;
    int nk, capk;
    coll_constraint* coll;
    /* copy of last detection's contacts for orc_get_contacts */
    int nlast;
    coll_constraint* last;
    double collision_compliance;
    uint64_t projected, early_out;
};

orc_world* orc_create(void)
{
    orc_world* w            = (orc_world*)calloc(1, sizeof(orc_world));
    w->collision_compliance = 1e-8; /* simulation_parameters.h:24 */
    return w;
}
void orc_destroy(orc_world* w)
{
    if (!w)
        return;
    for (int i = 0; i < w->nb; ++i)
    {
        free(w->bodies[i].p);
        free(w->bodies[i].surf_to_tet);
        free(w->bodies[i].surf_pos);
        free(w->bodies[i].grid_nodes);
    }
    free(w->bodies);
    free(w->cons);
    free(w->coll);
    free(w->last);
    free(w);
}
void orc_set_collision_compliance(orc_world* w, double a) { w->collision_compliance = a; }

static body* new_body(orc_world* w)
{
    if (w->nb == w->capb)
    {
        w->capb   = w->capb ? 2 * w->capb : 8;
        w->bodies = (body*)realloc(w->bodies, sizeof(body) * (size_t)w->capb);
    }
    body* b = &w->bodies[w->nb++];
    memset(b, 0, sizeof *b);
    return b;
}
static constraint* new_constraint(orc_world* w)
{
    if (w->nc == w->capc)
    {
        w->capc = w->capc ? 2 * w->capc : 1024;
        w->cons = (constraint*)realloc(w->cons, sizeof(constraint) * (size_t)w->capc);
    }
    constraint* c = &w->cons[w->nc++];
    memset(c, 0, sizeof *c);
    return c;
}

static void refresh_surface(body* b)
{ /* tetrahedral_body_t::update_visual_model (tetrahedral_body.cpp:157-165) */
    for (int i = 0; i < b->nVs; ++i)
        memcpy(&b->surf_pos[3 * i], b->p[b->surf_to_tet[i]].x, 3 * sizeof(double));
}

int orc_add_tet_body(orc_world* w, int nV, const double* x0, const double* mass, int nT,
                     const uint32_t* tets, double young, double poisson, double alpha, double beta)
{
    int const id = w->nb;
    body* b      = new_body(w);
    b->kind      = B_TET;
    b->nV        = nV;
    b->p         = (particle*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(particle));
    for (int i = 0; i < nV; ++i)
    { /* particle_t(position) (particle.cpp:6-9) */
        particle* p = &b->p[i];
        for (int r = 0; r < 3; ++r)
            p->x0[r] = p->xi[r] = p->xn[r] = p->x[r] = x0[3 * i + r];
        p->m = mass ? mass[i] : 1.;
    }
    b->surf_to_tet = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nV > 0 ? nV : 1));
    b->nVs         = orc_boundary_surface(nV, nT, tets, b->surf_to_tet, NULL, NULL);
    b->surf_pos    = (double*)malloc(sizeof(double) * 3 * (size_t)(b->nVs > 0 ? b->nVs : 1));
    refresh_surface(b);
    double const mu  = young / (2. * (1 + poisson));
    double const lam = (young * poisson) / ((1 + poisson) * (1 - 2 * poisson));
    for (int t = 0; t < nT; ++t)
    {
        constraint* c = new_constraint(w);
        b             = &w->bodies[id];
        c->type       = C_GREEN;
        c->alpha      = alpha;
        c->beta       = beta;
        c->b1 = c->b2 = id;
        double r0[12];
        for (int a = 0; a < 4; ++a)
        {
            c->v[a] = tets[4 * t + a];
            memcpy(&r0[3 * a], b->p[c->v[a]].x0, 3 * sizeof(double));
        }
        orc_green_rest_state(r0, c->DmInv, &c->V0);
        c->mu  = mu;
        c->lam = lam;
    }
    return id;
}

int orc_add_distance_constraints(orc_world* w, int b1, int b2, int n, const uint32_t* pairs,
                                 double alpha, double beta)
{
    if (b1 < 0 || b2 < 0 || b1 >= w->nb || b2 >= w->nb || w->bodies[b1].kind != B_TET ||
        w->bodies[b2].kind != B_TET)
        return -1;
    for (int i = 0; i < n; ++i)
    {
        constraint* c = new_constraint(w);
        c->type       = C_DISTANCE;
        c->alpha      = alpha;
        c->beta       = beta;
        c->b1         = b1;
        c->b2         = b2;
        c->v[0]       = pairs[2 * i];
        c->v[1]       = pairs[2 * i + 1];
        const double* a = w->bodies[b1].p[c->v[0]].x0;
        const double* b = w->bodies[b2].p[c->v[1]].x0;
        c->d = sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) +
                    (a[2] - b[2]) * (a[2] - b[2])); /* distance_constraint.cpp:21 */
    }
    return 0;
}

static int add_sdf(orc_world* w, int kind, const double a[3], const double b[3], double r,
                   const double volume[6])
{
    int const id = w->nb;
    body* bd     = new_body(w);
    bd->kind     = B_SDF;
    bd->sdf_kind = kind;
    memcpy(bd->a, a, sizeof bd->a);
    if (b)
        memcpy(bd->b, b, sizeof bd->b);
    bd->r = r;
    memcpy(bd->volume, volume, sizeof bd->volume);
    return id;
}
int orc_set_body_collideable(orc_world* w, int b, int flag)
{
    if (b < 0 || b >= w->nb)
        return -1;
    w->bodies[b].excluded = !flag;
    return 0;
}
int orc_add_sdf_plane(orc_world* w, const double n[3], const double pt[3], const double vol[6])
{ /* Eigen::Hyperplane(n, e): offset = -n.e ; signedDistance(p) = n.p + offset (sdf_model.cpp:56-61) */
    return add_sdf(w, SDF_PLANE, n, NULL, -(n[0] * pt[0] + n[1] * pt[1] + n[2] * pt[2]), vol);
}
int orc_add_sdf_sphere(orc_world* w, const double c[3], double r, const double vol[6])
{
    return add_sdf(w, SDF_SPHERE, c, NULL, r, vol);
}
int orc_add_sdf_box(orc_world* w, const double bmin[3], const double bmax[3], const double vol[6])
{
    return add_sdf(w, SDF_BOX, bmin, bmax, 0., vol);
}


/* ---- grid SDF: Discregrid::CubicLagrangeDiscreteGrid restated ---------------------------------
 * Discregrid (github.com/Q-Minh/Discregrid @ master, CMakeLists.txt:95-100 — a moving branch, not
 * vendored, absent from this image) is restated from its published algorithm: PARITY UNPINNED at
 * this boundary.  What pins the restatement is mathematics, checked in
 * tests/test_oracle_cpu.py: the 32 shape functions are a nodal basis (N_i(x_j) = delta_ij, partition
 * of unity) and reproduce every polynomial of degree <= 3 exactly, values and gradients.
 *
 * Nodes of an (nx, ny, nz)-cell grid: first the (nx+1)(ny+1)(nz+1) cell corners, x fastest; then two
 * nodes at 1/3 and 2/3 of every x-edge (edge index x fastest, then y, then z), of every y-edge (y
 * fastest, then z, then x) and of every z-edge (z fastest, then x, then y). */
static void grid_counts(const uint32_t n[3], uint64_t* nv, uint64_t* nex, uint64_t* ney, uint64_t* nez)
{
    *nv  = (uint64_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
    *nex = (uint64_t)n[0] * (n[1] + 1) * (n[2] + 1);
    *ney = (uint64_t)(n[0] + 1) * n[1] * (n[2] + 1);
    *nez = (uint64_t)(n[0] + 1) * (n[1] + 1) * n[2];
}
int64_t orc_grid_node_count(const uint32_t n[3])
{
    uint64_t nv, nex, ney, nez;
    grid_counts(n, &nv, &nex, &ney, &nez);
    return (int64_t)(nv + 2 * (nex + ney + nez));
}
/* CubicLagrangeDiscreteGrid::indexToNodePosition */
void orc_grid_node_position(const double dmin[3], const double dmax[3], const uint32_t n[3], int64_t l,
                            double x[3])
{
    uint64_t nv, nex, ney, nez;
    grid_counts(n, &nv, &nex, &ney, &nez);
    double cell[3];
    for (int d = 0; d < 3; ++d)
        cell[d] = (dmax[d] - dmin[d]) / (double)n[d];
    uint64_t u = (uint64_t)l, ijk[3];
    int axis = -1;
    if (u < nv)
    {
        uint64_t const nxy = (uint64_t)(n[0] + 1) * (n[1] + 1), t = u % nxy;
        ijk[2] = u / nxy;
        ijk[1] = t / (n[0] + 1);
        ijk[0] = t % (n[0] + 1);
    }
    else if (u < nv + 2 * nex)
    {
        u -= nv;
        uint64_t const e = u / 2, t = e % ((uint64_t)(n[1] + 1) * n[0]);
        ijk[2] = e / ((uint64_t)(n[1] + 1) * n[0]);
        ijk[1] = t / n[0];
        ijk[0] = t % n[0];
        axis   = 0;
    }
    else if (u < nv + 2 * (nex + ney))
    {
        u -= nv + 2 * nex;
        uint64_t const e = u / 2, t = e % ((uint64_t)(n[2] + 1) * n[1]);
        ijk[0] = e / ((uint64_t)(n[2] + 1) * n[1]);
        ijk[2] = t / n[1];
        ijk[1] = t % n[1];
        axis   = 1;
    }
    else
    {
        u -= nv + 2 * (nex + ney);
        uint64_t const e = u / 2, t = e % ((uint64_t)(n[0] + 1) * n[2]);
        ijk[1] = e / ((uint64_t)(n[0] + 1) * n[2]);
        ijk[0] = t / n[2];
        ijk[2] = t % n[2];
        axis   = 2;
    }
    for (int d = 0; d < 3; ++d)
        x[d] = dmin[d] + cell[d] * (double)ijk[d];
    if (axis >= 0)
        x[axis] += (1.0 + (double)(u % 2)) / 3.0 * cell[axis];
}
/* node indices of cell (i, j, k) in the order of the shape functions */
static void grid_cell_nodes(const uint32_t n[3], uint64_t i, uint64_t j, uint64_t k, uint64_t c[32])
{
    uint64_t nv, nex, ney, nez;
    grid_counts(n, &nv, &nex, &ney, &nez);
    uint64_t const nx = n[0], ny = n[1], nz = n[2];
    for (int q = 0; q < 8; ++q)
        c[q] = (nx + 1) * (ny + 1) * (k + (uint64_t)(q >> 2 & 1)) + (nx + 1) * (j + (uint64_t)(q >> 1 & 1)) + i +
               (uint64_t)(q & 1);
    uint64_t off = nv;
    for (int q = 0; q < 4; ++q) /* x-edges: (j, k), (j, k+1), (j+1, k), (j+1, k+1) */
    {
        c[8 + 2 * q]     = off + 2 * (nx * (ny + 1) * (k + (uint64_t)(q & 1)) + nx * (j + (uint64_t)(q >> 1)) + i);
        c[8 + 2 * q + 1] = c[8 + 2 * q] + 1;
    }
    off += 2 * nex;
    for (int q = 0; q < 4; ++q) /* y-edges: (i, k), (i+1, k), (i, k+1), (i+1, k+1) */
    {
        c[16 + 2 * q]     = off + 2 * (ny * (nz + 1) * (i + (uint64_t)(q & 1)) + ny * (k + (uint64_t)(q >> 1)) + j);
        c[16 + 2 * q + 1] = c[16 + 2 * q] + 1;
    }
    off += 2 * ney;
    for (int q = 0; q < 4; ++q) /* z-edges: (i, j), (i, j+1), (i+1, j), (i+1, j+1) */
    {
        c[24 + 2 * q]     = off + 2 * (nz * (nx + 1) * (j + (uint64_t)(q & 1)) + nz * (i + (uint64_t)(q >> 1)) + k);
        c[24 + 2 * q + 1] = c[24 + 2 * q] + 1;
    }
}
/* 32-node serendipity cubic shape functions on [-1, 1]^3 and their derivatives
 * (CubicLagrangeDiscreteGrid's shape_function_): corners N = (1/64)(1 +- x)(1 +- y)(1 +- z)
 * (9(x^2+y^2+z^2) - 19); edge nodes at -+1/3 along x: (9/64)(1 - x^2)(1 -+ 3x)(1 +- y)(1 +- z). */
void orc_grid_shape(const double xi[3], double N[32], double dN[32][3])
{
    double const x = xi[0], y = xi[1], z = xi[2];
    double const r2 = x * x + y * y + z * z;
    double const s[3][2] = {{1 - x, 1 + x}, {1 - y, 1 + y}, {1 - z, 1 + z}};
    for (int q = 0; q < 8; ++q)
    {
        double const a = s[0][q & 1], b = s[1][q >> 1 & 1], c = s[2][q >> 2 & 1];
        double const sa = (q & 1) ? 1. : -1., sb = (q >> 1 & 1) ? 1. : -1., sc = (q >> 2 & 1) ? 1. : -1.;
        double const f = 9. * r2 - 19.;
        N[q]     = a * b * c * f / 64.;
        dN[q][0] = (sa * b * c * f + a * b * c * 18. * x) / 64.;
        dN[q][1] = (a * sb * c * f + a * b * c * 18. * y) / 64.;
        dN[q][2] = (a * b * sc * f + a * b * c * 18. * z) / 64.;
    }
    /* edge nodes: axis t carries the cubic, (u, v) select the edge */
    static const int uv[3][2]  = {{1, 2}, {0, 2}, {0, 1}};
    /* sign selectors of the 4 edges of each family, in the order of grid_cell_nodes */
    static const int sel[3][4][2] = {
        {{0, 0}, {0, 1}, {1, 0}, {1, 1}}, /* x-edges: (y, z) = (-,-), (-,+), (+,-), (+,+) */
        {{0, 0}, {1, 0}, {0, 1}, {1, 1}}, /* y-edges: (x, z) = (-,-), (+,-), (-,+), (+,+) */
        {{0, 0}, {0, 1}, {1, 0}, {1, 1}}, /* z-edges: (x, y) = (-,-), (-,+), (+,-), (+,+) */
    };
    for (int t = 0; t < 3; ++t)
    {
        double const w = xi[t];
        int const u = uv[t][0], v = uv[t][1];
        for (int q = 0; q < 4; ++q)
        {
            double const fu = s[u][sel[t][q][0]], fv = s[v][sel[t][q][1]];
            double const du = sel[t][q][0] ? 1. : -1., dv = sel[t][q][1] ? 1. : -1.;
            for (int h = 0; h < 2; ++h)
            {
                double const sg = h ? 3. : -3.;               /* node at w = -1/3 (h = 0) or +1/3 */
                double const g  = (1. - w * w) * (1. + sg * w); /* cubic along the edge */
                double const dg = -2. * w * (1. + sg * w) + (1. - w * w) * sg;
                int const id    = 8 + 8 * t + 2 * q + h;
                N[id]           = 9. / 64. * g * fu * fv;
                dN[id][t]       = 9. / 64. * dg * fu * fv;
                dN[id][u]       = 9. / 64. * g * du * fv;
                dN[id][v]       = 9. / 64. * g * fu * dv;
            }
        }
    }
}
/* CubicLagrangeDiscreteGrid::interpolate(field, x, &gradient) as sdf_model_t::evaluate calls it
 * (sdf_model.cpp:71-74): DBL_MAX outside the domain. */
double orc_grid_interpolate(const double dmin[3], const double dmax[3], const uint32_t n[3],
                            const double* nodes, const double p[3], double grad[3])
{
    uint64_t ijk[3];
    double xi[3], c0[3];
    for (int d = 0; d < 3; ++d)
    {
        if (!(dmin[d] <= p[d] && p[d] <= dmax[d]))
        {
            if (grad)
                grad[0] = grad[1] = grad[2] = 0.;
            return 1.7976931348623157e308;
        }
        double const cell = (dmax[d] - dmin[d]) / (double)n[d];
        uint64_t m        = (uint64_t)((p[d] - dmin[d]) * (1. / cell));
        if (m >= n[d])
            m = n[d] - 1;
        ijk[d]          = m;
        double const lo = dmin[d] + cell * (double)m, hi = lo + cell;
        c0[d]           = 2. / (hi - lo);
        xi[d]           = c0[d] * p[d] - (hi + lo) / (hi - lo);
    }
    uint64_t c[32];
    double N[32], dN[32][3];
    grid_cell_nodes(n, ijk[0], ijk[1], ijk[2], c);
    orc_grid_shape(xi, N, dN);
    double phi = 0., g[3] = {0., 0., 0.};
    for (int q = 0; q < 32; ++q)
    {
        double const v = nodes[c[q]];
        phi += N[q] * v;
        for (int d = 0; d < 3; ++d)
            g[d] += dN[q][d] * v;
    }
    if (grad)
        for (int d = 0; d < 3; ++d)
            grad[d] = g[d] * c0[d];
    return phi;
}

/* ---- Discregrid::MeshDistance restated (brute force over the triangles; Discregrid's BVH and cache
 * only accelerate the same query): unsigned distance to the closest triangle, sign from the
 * angle-weighted pseudo-normal of the closest feature (face, edge or vertex).  PARITY UNPINNED, as
 * above; pinned by analytic shapes in the tests (cube, octahedron). ---------------------------------- */
typedef struct
{
    int nV, nF;
    const double* x;
    const uint32_t* f;
    double* fn; /* 3*nF unit face normals */
    double* vn; /* 3*nV angle-weighted vertex normals */
    double* en; /* 9*nF: pseudo-normal of edge (i -> i+1) of face: own normal + the opposite face's */
} mesh_dist;
static void v3sub(const double* a, const double* b, double* r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3cross(const double* a, const double* b, double* r)
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static void mesh_dist_init(mesh_dist* m, int nV, const double* x, int nF, const uint32_t* f)
{
    m->nV = nV; m->nF = nF; m->x = x; m->f = f;
    m->fn = (double*)calloc((size_t)(3 * nF + 3), sizeof(double));
    m->vn = (double*)calloc((size_t)(3 * nV + 3), sizeof(double));
    m->en = (double*)calloc((size_t)(9 * nF + 9), sizeof(double));
    for (int t = 0; t < nF; ++t)
    {
        const double *p0 = &x[3 * f[3 * t]], *p1 = &x[3 * f[3 * t + 1]], *p2 = &x[3 * f[3 * t + 2]];
        double e[3][3], n[3];
        v3sub(p1, p0, e[0]); v3sub(p2, p1, e[1]); v3sub(p0, p2, e[2]);
        double a[3];
        v3sub(p2, p0, a);
        v3cross(e[0], a, n);
        double const ln = sqrt(v3dot(n, n));
        for (int d = 0; d < 3; ++d)
            m->fn[3 * t + d] = n[d] / ln;
        double len[3];
        for (int k = 0; k < 3; ++k)
            len[k] = sqrt(v3dot(e[k], e[k]));
        for (int k = 0; k < 3; ++k)
        { /* interior angle at corner k: between edge k and the reversed previous edge */
            int const pk = (k + 2) % 3;
            double cs    = -v3dot(e[k], e[pk]) / (len[k] * len[pk]);
            cs           = cs > 1. ? 1. : cs < -1. ? -1. : cs;
            double const al = acos(cs);
            for (int d = 0; d < 3; ++d)
                m->vn[3 * f[3 * t + k] + d] += al * m->fn[3 * t + d];
        }
    }
    for (int t = 0; t < nF; ++t)
        for (int k = 0; k < 3; ++k)
        {
            uint32_t const a = f[3 * t + k], b = f[3 * t + (k + 1) % 3];
            for (int d = 0; d < 3; ++d)
                m->en[9 * t + 3 * k + d] = m->fn[3 * t + d];
            for (int o = 0; o < nF; ++o) /* the face holding the opposite half-edge b -> a */
            {
                if (o == t)
                    continue;
                int hit = 0;
                for (int q = 0; q < 3; ++q)
                    hit |= f[3 * o + q] == b && f[3 * o + (q + 1) % 3] == a;
                if (hit)
                {
                    for (int d = 0; d < 3; ++d)
                        m->en[9 * t + 3 * k + d] += m->fn[3 * o + d];
                    break;
                }
            }
        }
}
static void mesh_dist_free(mesh_dist* m) { free(m->fn); free(m->vn); free(m->en); }
/* closest point of triangle (a, b, c) to p; feature: 0-2 vertex, 3-5 edge (k -> k+1), 6 face */
static double closest_on_triangle(const double* p, const double* a, const double* b, const double* c, double* q,
                                  int* feature)
{
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    v3sub(b, a, ab); v3sub(c, a, ac); v3sub(p, a, ap);
    double const d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
    double s = 0., t = 0.;
    if (d1 <= 0. && d2 <= 0.) { *feature = 0; s = 0.; t = 0.; goto done; }
    v3sub(p, b, bp);
    double const d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
    if (d3 >= 0. && d4 <= d3) { *feature = 1; s = 1.; t = 0.; goto done; }
    double const vc = d1 * d4 - d3 * d2;
    if (vc <= 0. && d1 >= 0. && d3 <= 0.) { *feature = 3; s = d1 / (d1 - d3); t = 0.; goto done; }
    v3sub(p, c, cp);
    double const d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
    if (d6 >= 0. && d5 <= d6) { *feature = 2; s = 0.; t = 1.; goto done; }
    double const vb = d5 * d2 - d1 * d6;
    if (vb <= 0. && d2 >= 0. && d6 <= 0.) { *feature = 5; s = 0.; t = d2 / (d2 - d6); goto done; }
    double const va = d3 * d6 - d5 * d4;
    if (va <= 0. && (d4 - d3) >= 0. && (d5 - d6) >= 0.)
    {
        double const w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *feature = 4; s = 1. - w; t = w; goto done;
    }
    {
        double const den = 1. / (va + vb + vc);
        *feature = 6; s = vb * den; t = vc * den;
    }
done:
    for (int d = 0; d < 3; ++d)
        q[d] = a[d] + s * ab[d] + t * ac[d];
    double r[3];
    v3sub(p, q, r);
    return v3dot(r, r);
}
static double mesh_signed_distance(const mesh_dist* m, const double p[3])
{
    double best = 1.7976931348623157e308, bq[3] = {0, 0, 0};
    int bf = -1, bfeat = 6;
    for (int t = 0; t < m->nF; ++t)
    {
        double q[3];
        int feat;
        double const d2 = closest_on_triangle(p, &m->x[3 * m->f[3 * t]], &m->x[3 * m->f[3 * t + 1]],
                                              &m->x[3 * m->f[3 * t + 2]], q, &feat);
        if (d2 < best)
        {
            best = d2; bf = t; bfeat = feat;
            memcpy(bq, q, sizeof bq);
        }
    }
    if (bf < 0)
        return best;
    const double* n = bfeat < 3 ? &m->vn[3 * m->f[3 * bf + bfeat]] : bfeat < 6 ? &m->en[9 * bf + 3 * (bfeat - 3)]
                                                                              : &m->fn[3 * bf];
    double r[3];
    v3sub(p, bq, r);
    double const dist = sqrt(best);
    return v3dot(r, n) < 0. ? -dist : dist;
}
void orc_mesh_signed_distance(int nV, const double* x, int nF, const uint32_t* faces, int n, const double* points,
                              double* out)
{
    mesh_dist m;
    mesh_dist_init(&m, nV, x, nF, faces);
    for (int i = 0; i < n; ++i)
        out[i] = mesh_signed_distance(&m, &points[3 * i]);
    mesh_dist_free(&m);
}
/* environment_body_t(simulation, id, geometry, domain, resolution) (environment_body.cpp:12-78): the
 * domain is extended to the mesh and inflated ONCE PER MESH VERTEX (the growth statements sit inside
 * the outer vertex loop, :55-65), the mesh distance is sampled at every grid node (:67-74). */
void orc_mesh_sdf_domain(int nV, const double* x, const double domain[6], double out[6])
{
    memcpy(out, domain, 6 * sizeof(double));
    for (int v = 0; v < nV; ++v)
    {
        if (v == 0) /* the inner loop's extension is idempotent after the first pass */
            for (int u = 0; u < nV; ++u)
                for (int d = 0; d < 3; ++d)
                {
                    if (x[3 * u + d] < out[d]) out[d] = x[3 * u + d];
                    if (x[3 * u + d] > out[3 + d]) out[3 + d] = x[3 * u + d];
                }
        for (int side = 1; side >= 0; --side)
        {
            double const dx = out[3] - out[0], dy = out[4] - out[1], dz = out[5] - out[2];
            double const g  = 1.0e-3 * sqrt(dx * dx + dy * dy + dz * dz);
            for (int d = 0; d < 3; ++d)
                out[3 * side + d] += side ? g : -g;
        }
    }
}
int64_t orc_bake_mesh_sdf(int nV, const double* x, int nF, const uint32_t* faces, const double domain[6],
                          const uint32_t res[3], double out_domain[6], double* nodes, int64_t cap)
{
    orc_mesh_sdf_domain(nV, x, domain, out_domain);
    int64_t const nn = orc_grid_node_count(res);
    if (!nodes || cap < nn)
        return nn;
    mesh_dist m;
    mesh_dist_init(&m, nV, x, nF, faces);
    for (int64_t l = 0; l < nn; ++l)
    {
        double p[3];
        orc_grid_node_position(out_domain, out_domain + 3, res, l, p);
        nodes[l] = mesh_signed_distance(&m, p);
    }
    mesh_dist_free(&m);
    return nn;
}
/* environment_body_t holding a grid sdf_model_t (sdf_model.cpp:18, environment_body.cpp:75-77):
 * volume() is the grid's domain unless given. */
int orc_add_sdf_grid(orc_world* w, const double dmin[3], const double dmax[3], const uint32_t res[3],
                     const double* nodes, const double vol[6])
{
    double v6[6] = {dmin[0], dmin[1], dmin[2], dmax[0], dmax[1], dmax[2]};
    int const id = add_sdf(w, SDF_GRID, dmin, dmax, 0., vol ? vol : v6);
    body* bd     = &w->bodies[id];
    memcpy(bd->grid_n, res, sizeof bd->grid_n);
    int64_t const nn = orc_grid_node_count(res);
    bd->grid_nodes   = (double*)malloc(sizeof(double) * (size_t)nn);
    memcpy(bd->grid_nodes, nodes, sizeof(double) * (size_t)nn);
    return id;
}

/* sdf_model_t::evaluate (sdf_model.cpp:66-75): (signed distance, gradient) */
static double sdf_eval(const body* s, const double p[3], double g[3])
{
    switch (s->sdf_kind)
    {
    case SDF_GRID:
        return orc_grid_interpolate(s->a, s->b, s->grid_n, s->grid_nodes, p, g);
    case SDF_PLANE:
        g[0] = s->a[0];
        g[1] = s->a[1];
        g[2] = s->a[2];
        return s->a[0] * p[0] + s->a[1] * p[1] + s->a[2] * p[2] + s->r;
    case SDF_SPHERE:
    {
        double d[3] = {p[0] - s->a[0], p[1] - s->a[1], p[2] - s->a[2]};
        double const len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (len > 0.)
        {
            g[0] = d[0] / len;
            g[1] = d[1] / len;
            g[2] = d[2] / len;
        }
        else
        {
            g[0] = 0.;
            g[1] = 1.;
            g[2] = 0.;
        }
        return len - s->r;
    }
    default:
    { /* box: exact signed distance, gradient of the active branch */
        double c[3], h[3], q[3], out2 = 0.;
        for (int i = 0; i < 3; ++i)
        {
            c[i] = 0.5 * (s->a[i] + s->b[i]);
            h[i] = 0.5 * (s->b[i] - s->a[i]);
            q[i] = fabs(p[i] - c[i]) - h[i];
            if (q[i] > 0.)
                out2 += q[i] * q[i];
        }
        if (out2 > 0.)
        {
            double const len = sqrt(out2);
            for (int i = 0; i < 3; ++i)
                g[i] = q[i] > 0. ? (p[i] >= c[i] ? q[i] : -q[i]) / len : 0.;
            return len;
        }
        int ax = 0;
        if (q[1] > q[ax]) ax = 1;
        if (q[2] > q[ax]) ax = 2;
        g[0] = g[1] = g[2] = 0.;
        g[ax]              = p[ax] >= c[ax] ? 1. : -1.;
        return q[ax];
    }
    }
}

/* point_bvh_model_t::collide narrowphase (bvh_model.cpp:66-96) + contact_handler_t::handle
 * (xpbd/contact_handler.cpp:14-54), for the pair (tet body tb, sdf body sb). */
static void collide_pair(orc_world* w, int tb, int sb)
{
    body* t       = &w->bodies[tb];
    body const* s = &w->bodies[sb];
    for (int vi = 0; vi < t->nVs; ++vi)
    {
        const double* pi = &t->surf_pos[3 * vi];
        double g[3];
        double const sd = sdf_eval(s, pi, g);
        if (!(sd < 0.)) /* :78-80 */
            continue;
        double const gl = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        if (w->nk == w->capk)
        {
            w->capk = w->capk ? 2 * w->capk : 256;
            w->coll = (coll_constraint*)realloc(w->coll, sizeof(coll_constraint) * (size_t)w->capk);
        }
        coll_constraint* c = &w->coll[w->nk++];
        c->alpha           = w->collision_compliance;
        c->lagrange        = 0.;
        c->b               = tb;
        c->sdf_body        = sb;
        c->v               = t->surf_to_tet[vi]; /* from_surface_vertex (contact_handler.cpp:39-40) */
        for (int r = 0; r < 3; ++r)
        {
            c->n[r]  = g[r] / gl;                    /* :82 */
            c->qs[r] = pi[r] + fabs(sd) * c->n[r];   /* :83-84 */
        }
    }
}

/* brute_force_cd_system_t::execute (brute_force_cd_system.cpp:14-26): all i<j pairs; only
 * bvh-sdf pairs act (bvh_model.cpp:33-34, sdf_model.cpp:38-43). */
static void detect(orc_world* w)
{
    for (int i = 0; i < w->nb; ++i)
        for (int j = i + 1; j < w->nb; ++j)
        {
            int const ki = w->bodies[i].kind, kj = w->bodies[j].kind;
            if (w->bodies[i].excluded || w->bodies[j].excluded)
                continue;
            if (ki == B_TET && kj == B_SDF)
                collide_pair(w, i, j);
            else if (ki == B_SDF && kj == B_TET)
                collide_pair(w, j, i);
        }
    free(w->last);
    w->nlast = w->nk;
    w->last  = (coll_constraint*)malloc(sizeof(coll_constraint) * (size_t)(w->nk > 0 ? w->nk : 1));
    memcpy(w->last, w->coll, sizeof(coll_constraint) * (size_t)w->nk);
}

static void project_collision(orc_world* w, coll_constraint* c, double dt)
{ /* collision_constraint.cpp:21-48 */
    particle* p    = &w->bodies[c->b].p[c->v];
    double const iw = invmass(p);
    double const C  = (p->xi[0] - c->qs[0]) * c->n[0] + (p->xi[1] - c->qs[1]) * c->n[1] +
                     (p->xi[2] - c->qs[2]) * c->n[2];
    if (C >= 0.)
        return;
    double const at = c->alpha / (dt * dt);
    double const dl = -(C + at * c->lagrange) / (iw + at);
    c->lagrange += dl;
    for (int r = 0; r < 3; ++r)
        p->xi[r] += iw * c->n[r] * dl;
}

static void project_distance(orc_world* w, constraint* c, double dt)
{ /* distance_constraint.cpp:24-53 */
    particle* p1 = &w->bodies[c->b1].p[c->v[0]];
    particle* p2 = &w->bodies[c->b2].p[c->v[1]];
    double const w1 = invmass(p1), w2 = invmass(p2);
    double diff[3] = {p1->xi[0] - p2->xi[0], p1->xi[1] - p2->xi[1], p1->xi[2] - p2->xi[2]};
    double const len = sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]);
    double n[3]      = {diff[0] / len, diff[1] / len, diff[2] / len};
    double const C   = len - c->d;
    double const S   = w1 + w2;
    double const dt2 = dt * dt;
    double const at = c->alpha / dt2, bt = c->beta * dt2;
    double g = 0.;
    for (int r = 0; r < 3; ++r)
        g += n[r] * (p1->xi[r] - p1->xn[r]) - n[r] * (p2->xi[r] - p2->xn[r]);
    double const gam = at * bt / dt;
    double const dl  = (-(C + at * c->lagrange) + gam * g) / ((1. + gam) * S + at);
    c->lagrange += dl;
    for (int r = 0; r < 3; ++r)
    {
        p1->xi[r] += w1 * n[r] * dl;
        p2->xi[r] += w2 * -n[r] * dl;
    }
}

static void project_green_c(orc_world* w, constraint* c, double dt)
{
    particle* p  = w->bodies[c->b1].p;
    double* xi[4];
    const double* xn[4];
    double iw[4];
    for (int a = 0; a < 4; ++a)
    {
        xi[a] = p[c->v[a]].xi;
        xn[a] = p[c->v[a]].xn;
        iw[a] = invmass(&p[c->v[a]]);
    }
    if (green_project(xi, xn, iw, c->DmInv, c->V0, c->mu, c->lam, c->alpha, c->beta, dt,
                      &c->lagrange, NULL))
        w->projected++;
    else
        w->early_out++;
}

/* gauss_seidel_solver_t::solve (gauss_seidel_solver.cpp:8-37) */
static void solve(orc_world* w, double dt, int iterations)
{
    for (int i = 0; i < w->nk; ++i)
        w->coll[i].lagrange = 0.; /* :15-18 */
    for (int i = 0; i < w->nc; ++i)
        w->cons[i].lagrange = 0.; /* :19-22 */
    for (int k = 0; k < iterations; ++k)
    {
        for (int i = 0; i < w->nk; ++i)
            project_collision(w, &w->coll[i], dt); /* :28-31 */
        for (int i = 0; i < w->nc; ++i)
        { /* :32-35 */
            if (w->cons[i].type == C_GREEN)
                project_green_c(w, &w->cons[i], dt);
            else
                project_distance(w, &w->cons[i], dt);
        }
    }
}

/* timestep_t::step (timestep.cpp:20-70) */
static void step_once(orc_world* w, double dt_frame, int substeps, int iterations)
{
    double const dt = dt_frame / (double)substeps; /* :22 */
    detect(w);                                     /* :29-30 */
    for (int s = 0; s < substeps; ++s)
    {
        for (int b = 0; b < w->nb; ++b)
            for (int i = 0; i < w->bodies[b].nV; ++i)
            { /* :35-43 */
                particle* p     = &w->bodies[b].p[i];
                p->f[1] -= 9.81;
                double const iw = invmass(p);
                for (int r = 0; r < 3; ++r)
                {
                    p->v[r]  = p->v[r] + (p->f[r] * iw) * dt;
                    p->xi[r] = p->x[r] + p->v[r] * dt;
                }
            }
        solve(w, dt, iterations); /* :45 */
        for (int b = 0; b < w->nb; ++b)
            for (int i = 0; i < w->bodies[b].nV; ++i)
            { /* :48-57 */
                particle* p = &w->bodies[b].p[i];
                for (int r = 0; r < 3; ++r)
                {
                    p->x[r]  = p->xi[r];
                    p->v[r]  = (p->x[r] - p->xn[r]) / dt;
                    p->xn[r] = p->x[r];
                    p->f[r]  = 0.;
                }
            }
    }
    for (int b = 0; b < w->nb; ++b) /* :60-66 */
        if (w->bodies[b].kind == B_TET)
            refresh_surface(&w->bodies[b]);
    w->nk = 0; /* :68 */
}

int orc_step(orc_world* w, double dt, int substeps, int iterations, int detect_every_substep)
{
    if (!w || substeps <= 0 || iterations < 0)
        return -1;
    if (!detect_every_substep)
        step_once(w, dt, substeps, iterations);
    else
        for (int s = 0; s < substeps; ++s)
            step_once(w, dt / (double)substeps, 1, iterations);
    return 0;
}

int orc_constraint_count(const orc_world* w) { return w->nc; }

/* simulation_t::remove_constraint (simulation.cpp:34-39): swap with the last constraint, drop it */
int orc_remove_constraint(orc_world* w, uint32_t index)
{
    if (index >= (uint32_t)w->nc)
        return -1;
    w->cons[index] = w->cons[w->nc - 1];
    --w->nc;
    return 0;
}

int orc_set_constraint_order(orc_world* w, const uint32_t* order, int n)
{
    if (n != w->nc)
        return -1;
    char* seen = (char*)calloc((size_t)n + 1, 1);
    for (int i = 0; i < n; ++i)
    {
        if (order[i] >= (uint32_t)n || seen[order[i]])
        {
            free(seen);
            return -2;
        }
        seen[order[i]] = 1;
    }
    free(seen);
    constraint* nc = (constraint*)malloc(sizeof(constraint) * (size_t)(w->capc > 0 ? w->capc : 1));
    for (int i = 0; i < n; ++i)
        nc[i] = w->cons[order[i]];
    free(w->cons);
    w->cons = nc;
    return 0;
}

int orc_upload(orc_world* w, int bi, const double* x, const double* v)
{
    if (bi < 0 || bi >= w->nb || w->bodies[bi].kind != B_TET)
        return -1;
    body* b = &w->bodies[bi];
    for (int i = 0; i < b->nV; ++i)
        for (int r = 0; r < 3; ++r)
        {
            b->p[i].x[r] = b->p[i].xi[r] = b->p[i].xn[r] = x[3 * i + r];
            b->p[i].v[r]                                 = v ? v[3 * i + r] : 0.;
        }
    refresh_surface(b); /* as tetrahedral_body_t::transform does (tetrahedral_body.cpp:131) */
    return 0;
}
int orc_download(const orc_world* w, int bi, double* x, double* v)
{
    if (bi < 0 || bi >= w->nb || w->bodies[bi].kind != B_TET)
        return -1;
    body const* b = &w->bodies[bi];
    for (int i = 0; i < b->nV; ++i)
        for (int r = 0; r < 3; ++r)
        {
            if (x)
                x[3 * i + r] = b->p[i].x[r];
            if (v)
                v[3 * i + r] = b->p[i].v[r];
        }
    return 0;
}
int orc_set_mass(orc_world* w, int bi, int vertex, double mass)
{
    if (bi < 0 || bi >= w->nb || w->bodies[bi].kind != B_TET || vertex < 0 ||
        vertex >= w->bodies[bi].nV)
        return -1;
    w->bodies[bi].p[vertex].m = mass;
    return 0;
}

int orc_get_contacts(const orc_world* w, int cap, int32_t* bodyv, uint32_t* vertex,
                     int32_t* sdf_body, double* point, double* normal)
{
    int const n = w->nlast < cap ? w->nlast : cap;
    for (int i = 0; i < n && bodyv; ++i)
    {
        bodyv[i]    = w->last[i].b;
        vertex[i]   = w->last[i].v;
        sdf_body[i] = w->last[i].sdf_body;
        memcpy(&point[3 * i], w->last[i].qs, 3 * sizeof(double));
        memcpy(&normal[3 * i], w->last[i].n, 3 * sizeof(double));
    }
    return w->nlast;
}

void orc_get_counters(const orc_world* w, uint64_t* projected, uint64_t* early_out)
{
    if (projected)
        *projected = w->projected;
    if (early_out)
        *early_out = w->early_out;
}
