// GPU-built BVH broadphase over the surface vertices (point_bvh_model_t, bvh_model.cpp:24-126).
//
// The reference keeps one KD-tree of bounding spheres per tetrahedral body, refits every sphere
// once per frame, and traverses it breadth-first against an SDF: a node's children are visited
// when the node's sphere "collides" — sdf(centre) < 0 or the sphere reaches the SDF's englobing
// volume() box (bvh_model.cpp:47-64) — and every vertex of a visited leaf goes to the narrowphase
// (:66-96).  So a vertex is examined iff EVERY proper ancestor of its leaf passes that test.
//
// Here: a linear BVH (Karras 2012) rebuilt on the device at every detection.  Keys are
// (body, 30-bit Morton code of the surface copy); one radix sort; the radix tree over the sorted
// keys contains one subtree per body (the body id is the key's prefix), nodes spanning several
// bodies always pass; bounding spheres are fitted bottom-up (sphere of two spheres); and the
// traversal is turned inside out: one thread per surface vertex walks from its leaf to the root
// and gives up at the first ancestor that fails the test — no queue, no stack, fully parallel.
// Tree shape and spheres differ from Discregrid's (which is not pinned), so the visited set can
// only be compared where it does not depend on them: bodies inside the volume box (everything
// penetrating is found) and bodies whose root sphere misses it (nothing is found).
#pragma once

#include "xpbd_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace sbsb200 {

template <typename R>
struct BvhView
{
    int64_t n;                // surface vertices = leaves
    uint64_t* keys;           // (body << 32) | morton, unsorted
    uint64_t* keys_sorted;
    uint32_t* leaf_in;        // 0..n-1
    uint32_t* leaf_surface;   // sorted position -> surface vertex index
    int32_t* parent;          // [2n-1]: internal nodes 0..n-2, leaves n-1..2n-2
    int32_t* child;           // [2(n-1)] left, right of internal node i (node ids as in parent[])
    uint32_t* range_first;    // [n-1] first / last sorted position covered by internal node i
    uint32_t* range_last;
    Real4<R>* sphere;         // [2n-1] (centre, radius)
    uint32_t* visits;         // [n-1] arrival counter of the bottom-up pass
    R lo[3], inv_extent[3];   // quantisation box of the Morton codes (tree quality only)
    typename DeviceScene<R>::Sdf const* sdf;
};

__device__ __forceinline__ Real4<float> ld4_l2(Real4<float> const* p)
{
    float4 const v = __ldcg(reinterpret_cast<float4 const*>(p));
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Real4<double> ld4_l2(Real4<double> const* p)
{
    double2 const a = __ldcg(reinterpret_cast<double2 const*>(p));
    double2 const b = __ldcg(reinterpret_cast<double2 const*>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

template <typename R>
__global__ void __launch_bounds__(256) k_bvh_keys(DeviceScene<R> s, BvhView<R> b)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= b.n)
        return;
    Real4<R> const p = ld4(&s.surf_pos[i]);
    R const c[3]     = {p.x, p.y, p.z};
    uint32_t q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        R const u = (c[d] - b.lo[d]) * b.inv_extent[d] * R(1024);
        q[d]      = static_cast<uint32_t>(u < R(0) ? R(0) : u > R(1023) ? R(1023) : u);
    }
    uint32_t const morton = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    b.keys[i]    = (static_cast<uint64_t>(static_cast<uint32_t>(s.surf_body[i])) << 32) | morton;
    b.leaf_in[i] = static_cast<uint32_t>(i);
}

// length of the common prefix of the keys at sorted positions i and j (ties broken by position)
__device__ __forceinline__ int bvh_delta(uint64_t const* keys, int64_t n, int64_t i, int64_t j)
{
    if (j < 0 || j >= n)
        return -1;
    uint64_t const a = keys[i], c = keys[j];
    if (a == c)
        return 64 + __clzll(static_cast<long long>(static_cast<uint64_t>(i) ^ static_cast<uint64_t>(j)));
    return __clzll(static_cast<long long>(a ^ c));
}

// Karras, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees" (2012): internal node i
template <typename R>
__global__ void __launch_bounds__(256) k_bvh_tree(BvhView<R> b)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    int64_t const n = b.n;
    if (i >= n - 1)
        return;
    uint64_t const* k = b.keys_sorted;
    int const d       = bvh_delta(k, n, i, i + 1) - bvh_delta(k, n, i, i - 1) >= 0 ? 1 : -1;
    int const dmin    = bvh_delta(k, n, i, i - d);
    int64_t lmax      = 2;
    while (bvh_delta(k, n, i, i + lmax * d) > dmin)
        lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (bvh_delta(k, n, i, i + (l + t) * d) > dmin)
            l += t;
    int64_t const j   = i + l * d;
    int const dnode   = bvh_delta(k, n, i, j);
    int64_t split     = 0;
    for (int64_t t = (l + 1) / 2;; t = (t + 1) / 2)
    {
        if (bvh_delta(k, n, i, i + (split + t) * d) > dnode)
            split += t;
        if (t <= 1)
            break;
    }
    int64_t const gamma = i + split * d + (d < 0 ? -1 : 0);
    int64_t const first = i < j ? i : j, last = i < j ? j : i;
    int32_t const left  = static_cast<int32_t>(first == gamma ? (n - 1) + gamma : gamma);
    int32_t const right = static_cast<int32_t>(last == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1);
    b.child[2 * i]      = left;
    b.child[2 * i + 1]  = right;
    b.parent[left]      = static_cast<int32_t>(i);
    b.parent[right]     = static_cast<int32_t>(i);
    b.range_first[i]    = static_cast<uint32_t>(first);
    b.range_last[i]     = static_cast<uint32_t>(last);
    b.visits[i]         = 0u;
    if (i == 0)
        b.parent[0] = -1;
}

template <typename R>
__device__ __forceinline__ Real4<R> sphere_of_two(Real4<R> a, Real4<R> c)
{
    R const dx = c.x - a.x, dy = c.y - a.y, dz = c.z - a.z;
    R const d  = sqrt_(dx * dx + dy * dy + dz * dz);
    if (d + c.w <= a.w)
        return a;
    if (d + a.w <= c.w)
        return c;
    R const r = R(0.5) * (d + a.w + c.w);
    R const t = (r - a.w) / d;
    return Real4<R>{a.x + dx * t, a.y + dy * t, a.z + dz * t, r};
}

// bottom-up fit: the second thread to arrive at a node fits it from its two children
template <typename R>
__global__ void __launch_bounds__(256) k_bvh_fit(DeviceScene<R> s, BvhView<R> b)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    int64_t const n = b.n;
    if (i >= n)
        return;
    Real4<R> const p = ld4(&s.surf_pos[b.leaf_surface[i]]);
    int32_t node     = static_cast<int32_t>(n - 1 + i);
    b.sphere[node]   = Real4<R>{p.x, p.y, p.z, R(0)};
    if (n == 1)
        return;
    for (;;)
    {
        __threadfence();
        node = b.parent[node];
        if (node < 0 || atomicAdd(&b.visits[node], 1u) == 0u)
            return; // the sibling subtree is not done yet: its thread continues upwards
        __threadfence();
        // written by another SM a moment ago: read from L2
        Real4<R> const a = ld4_l2(&b.sphere[b.child[2 * node]]);
        Real4<R> const c = ld4_l2(&b.sphere[b.child[2 * node + 1]]);
        st4(&b.sphere[node], sphere_of_two<R>(a, c));
    }
}

// is_sphere_colliding_with_sdf (bvh_model.cpp:47-64)
template <typename R>
__device__ __forceinline__ bool sphere_reaches_sdf(typename DeviceScene<R>::Sdf const& f, Real4<R> sph)
{
    Vec3<R> g;
    if (sdf_eval<R>(f, Vec3<R>{sph.x, sph.y, sph.z}, g) < R(0))
        return true;
    R const c[3] = {sph.x, sph.y, sph.z};
    R dist2      = R(0);
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        R const q    = c[d] < f.vmin[d] ? f.vmin[d] : c[d] > f.vmax[d] ? f.vmax[d] : c[d];
        R const diff = c[d] - q;
        dist2 += diff * diff;
    }
    return dist2 < sph.w * sph.w;
}

// Broadphase: bit k of cull[i] is set when the traversal against SDF k never reaches the leaf of
// surface vertex i.  One thread per surface vertex (leaf), walking its ancestors.
template <typename R>
__global__ void __launch_bounds__(256) k_bvh_cull(DeviceScene<R> s, BvhView<R> b, uint32_t* cull)
{
    int64_t const i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    int64_t const n = b.n;
    if (i >= n)
        return;
    uint32_t const surface = b.leaf_surface[i];
    uint32_t const body    = static_cast<uint32_t>(b.keys_sorted[i] >> 32);
    uint32_t mask          = 0u;
    for (int32_t k = 0; k < s.n_sdf && k < 32; ++k)
    {
        bool visited = true;
        for (int32_t node = n > 1 ? b.parent[n - 1 + i] : -1; node >= 0 && visited; node = b.parent[node])
        {
            // a node that spans several bodies is above every per-body tree of the reference
            if (static_cast<uint32_t>(b.keys_sorted[b.range_first[node]] >> 32) != body ||
                static_cast<uint32_t>(b.keys_sorted[b.range_last[node]] >> 32) != body)
                break;
            visited = sphere_reaches_sdf<R>(s.sdf[k], ld4(&b.sphere[node]));
        }
        if (!visited)
            mask |= 1u << k;
    }
    cull[surface] = mask;
}

} // namespace sbsb200
