// GPU-built BVH broadphase over the surface vertices (point_bvh_model_t, bvh_model.cpp:24-126).
//
// The reference keeps one KD-tree of bounding spheres per tetrahedral body, refits every sphere
// once per frame, and traverses it breadth-first against an SDF: a node's children are visited
// when the node's sphere "collides" — sdf(centre) < 0 or the sphere reaches the SDF's englobing
// volume() box (bvh_model.cpp:47-64) — and every vertex of a visited leaf goes to the narrowphase
// (:66-96).  So a vertex is examined iff EVERY proper ancestor of its leaf passes that test.
//
// Here: a linear BVH in its implicit form.  Once per frame the surface vertices are sorted by
// (body, 30-bit Morton code of the surface copy) with one radix sort; the hierarchy is the
// complete binary tree over that order (node j of level l covers the sorted leaves
// [j 2^l, (j+1) 2^l) — spatially compact because of the Morton order), so it needs no pointers.
// At every detection the bounding spheres are refitted bottom-up (sphere of two spheres; 256 leaves
// per CTA in shared memory, then the few upper levels), nodes that span several bodies are marked
// and always pass (the reference has one tree per body), and the traversal is turned inside out:
// one thread per surface vertex tests the ancestors of its leaf — all loads independent, no
// queue, no stack.  Tree shape and spheres differ from Discregrid's (which is not pinned), so
// the visited set can only be compared where it does not depend on them: bodies inside the
// volume box (everything penetrating is found) and bodies whose top sphere misses it (nothing).
#pragma once

#include "xpbd_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace sbsb200 {

constexpr int kBvhMaxLevels = 32;
constexpr int kBvhLeafBlock = 256; // leaves fitted per CTA (levels 1..8 in shared memory)
constexpr int kBvhTopNodes  = 1024; // level-8 nodes a CTA of k_detect_all can hold (262 144 surface vertices)

template <typename R>
struct BvhView
{
    int64_t n;                // surface vertices = leaves
    uint64_t* keys;           // (body << 32) | morton, unsorted
    uint64_t* keys_sorted;
    uint32_t* leaf_in;        // 0..n-1
    uint32_t* leaf_surface;   // sorted position -> surface vertex index
    uint32_t* leaf_of_surface; // and back
    uint32_t* done_counter;   // CTAs of the fit kernel that finished the lower levels
    Real4<R>* sphere;         // nodes of level l >= 1 at sphere[level_offset[l] + j]: (centre, radius);
                              // radius < 0: the node spans several bodies and always passes
    int32_t n_levels;         // levels 1 .. n_levels - 1 exist (level 0 = the leaves themselves)
    int32_t n_bodies;         // key prefix of a body that is not handed to the cd system: n_bodies + its index
    int32_t top_in_detect;    // 1: levels 9.. are built by every CTA of k_detect_all in shared memory out of level 8 (at most
                              // kBvhTopNodes nodes there): the refit is then free of its serial tail and of atomics
    int64_t level_offset[kBvhMaxLevels];
    int64_t level_count[kBvhMaxLevels];
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
    // key prefix = the body; a body that is not handed to the cd system (surf_body = ~index) gets a prefix of its own
    // beyond the others, so that the (partial) radix sort keeps every body's leaves together
    int32_t const sb    = s.surf_body[i];
    uint32_t const body = sb >= 0 ? static_cast<uint32_t>(sb) : static_cast<uint32_t>(b.n_bodies + ~sb);
    b.keys[i]    = (static_cast<uint64_t>(body) << 32) | morton;
    b.leaf_in[i] = static_cast<uint32_t>(i);
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

// sphere of a node from its two children; a child that does not exist (odd count) is skipped, a
// child or a pair that spans several bodies makes the node "always pass" (radius -1)
template <typename R>
__device__ __forceinline__ Real4<R> fit_pair(Real4<R> a, uint32_t body_a, bool has_b, Real4<R> c, uint32_t body_c,
                                             uint32_t& body_out)
{
    body_out = body_a;
    if (!has_b)
        return a;
    if (a.w < R(0) || c.w < R(0) || body_a != body_c)
    {
        body_out = 0xffffffffu;
        return Real4<R>{a.x, a.y, a.z, R(-1)};
    }
    return sphere_of_two<R>(a, c);
}

// Refit of every bounding sphere.  Levels 1 .. 8: every CTA fits the subtree over its 256 consecutive
// leaves in shared memory.  Levels 9 ..: the LAST CTA to finish does them, level by level (at most
// n / 256 nodes at level 8), so the refit is one launch.
template <typename R>
__global__ void __launch_bounds__(kBvhLeafBlock) k_bvh_fit(DeviceScene<R> s, BvhView<R> b)
{
    __shared__ Real4<R> sph[kBvhLeafBlock];
    __shared__ uint32_t body[kBvhLeafBlock];
    __shared__ uint32_t ticket;
    int64_t const base = blockIdx.x * static_cast<int64_t>(kBvhLeafBlock);
    int64_t const i    = base + threadIdx.x;
    if (i < b.n)
    {
        uint32_t const surface        = b.leaf_surface[i];
        Real4<R> const p              = ld4(&s.surf_pos[surface]);
        sph[threadIdx.x]              = Real4<R>{p.x, p.y, p.z, R(0)};
        body[threadIdx.x]             = static_cast<uint32_t>(b.keys_sorted[i] >> 32);
        b.leaf_of_surface[surface]    = static_cast<uint32_t>(i);
    }
    __syncthreads();
    int64_t count = b.n - base < kBvhLeafBlock ? b.n - base : kBvhLeafBlock; // nodes of the level below, in this CTA
    for (int l = 1; l <= 8 && l < b.n_levels; ++l)
    {
        int64_t const here = (count + 1) / 2;
        Real4<R> node;
        uint32_t nb     = 0;
        bool const mine = threadIdx.x < here;
        if (mine)
        {
            int const right = 2 * threadIdx.x + 1 < kBvhLeafBlock ? 2 * threadIdx.x + 1 : 0;
            node = fit_pair<R>(sph[2 * threadIdx.x], body[2 * threadIdx.x], 2 * threadIdx.x + 1 < count, sph[right],
                               body[right], nb);
        }
        __syncthreads();
        if (mine)
        {
            sph[threadIdx.x]  = node;
            body[threadIdx.x] = nb;
            st4(&b.sphere[b.level_offset[l] + (base >> l) + threadIdx.x], node);
        }
        __syncthreads();
        count = here;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *s.contact_count = 0u; // (the detection that follows appends to the list; the previous substep is done with it)
    if (b.n_levels <= 9 || b.top_in_detect)
        return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        ticket = atomicAdd(b.done_counter, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1)
        return;
    __threadfence();
    for (int l = 9; l < b.n_levels; ++l)
    {
        int64_t const below = b.level_count[l - 1];
        for (int64_t j = threadIdx.x; j < b.level_count[l]; j += blockDim.x)
        {
            Real4<R> const a = ld4_l2(&b.sphere[b.level_offset[l - 1] + 2 * j]);
            bool const has_c = 2 * j + 1 < below;
            Real4<R> const c = has_c ? ld4_l2(&b.sphere[b.level_offset[l - 1] + 2 * j + 1]) : a;
            // the body of a subtree: of its first leaf; several bodies are already marked by radius < 0,
            // two single-body children of different bodies are told apart by their first leaves
            uint32_t const body_a = static_cast<uint32_t>(b.keys_sorted[(2 * j) << (l - 1)] >> 32);
            uint32_t const body_c = has_c ? static_cast<uint32_t>(b.keys_sorted[(2 * j + 1) << (l - 1)] >> 32) : body_a;
            uint32_t nb;
            st4(&b.sphere[b.level_offset[l] + j], fit_pair<R>(a, body_a, has_c, c, body_c, nb));
        }
        __threadfence_block();
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *b.done_counter = 0u;
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

// Broadphase: bit k of the result is set when the traversal against SDF k never reaches the leaf.
// The ancestors of the leaf are tested from kBvhFirstLevel upwards: the reference's KD-tree stops
// splitting at about ten points per leaf, so its smallest spheres correspond to our level 3
// (8 leaves); the levels below only feed the fit.
constexpr int kBvhFirstLevel = 3;

// s_top: levels 9.. built by this CTA in shared memory (null: they are in b.sphere like the others);
// s_top_offset[l - 9]: where level l starts in it
template <typename R>
__device__ __forceinline__ uint32_t bvh_cull_mask(DeviceScene<R> const& s, BvhView<R> const& b, int64_t leaf,
                                                  Real4<R> const* s_top, int32_t const* s_top_offset)
{
    uint32_t mask = 0u;
    bool above    = false; // reached a node that spans several bodies: above the reference's per-body trees
    for (int l0 = kBvhFirstLevel; l0 < b.n_levels && !above; l0 += 8)
    {
        Real4<R> sph[8]; // eight ancestors at a time: their loads are independent
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (l0 + j < b.n_levels)
                sph[j] = s_top && l0 + j >= 9 ? s_top[s_top_offset[l0 + j - 9] + (leaf >> (l0 + j))]
                                              : ld4(&b.sphere[b.level_offset[l0 + j] + (leaf >> (l0 + j))]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (l0 + j < b.n_levels && !above)
            {
                if (sph[j].w < R(0))
                    above = true;
                else
                    for (int32_t k = 0; k < s.n_sdf && k < 32; ++k)
                        if (!(mask >> k & 1u) && !sphere_reaches_sdf<R>(s.sdf[k], sph[j]))
                            mask |= 1u << k;
            }
    }
    return mask;
}

__device__ __forceinline__ float as_real(float, int v) { return __int_as_float(v); }
__device__ __forceinline__ double as_real(double, int v) { return static_cast<double>(v); }

// Narrowphase + contact handling (bvh_model.cpp:66-96, xpbd/contact_handler.cpp:14-54):
// thread per surface vertex, every SDF in body order, warp-aggregated append so that the
// contacts of one vertex are contiguous and ordered by SDF body.
template <typename R>
__global__ void __launch_bounds__(256) k_detect_all(DeviceScene<R> s, BvhView<R> b)
{
    // Levels 9.. of the sphere tree (at most a few hundred nodes) built here, by every CTA for itself, out of the level-8
    // nodes the refit wrote: the refit kernel then has no serial tail, and the walk reads these levels from shared memory.
    extern __shared__ __align__(16) unsigned char top_raw[];
    __shared__ int32_t s_top_offset[kBvhMaxLevels];
    Real4<R>* s_top = nullptr;
    if (b.n > 0 && b.top_in_detect && b.n_levels > 9)
    {
        int32_t const n8 = static_cast<int32_t>(b.level_count[8]);
        Real4<R>* lvl8   = reinterpret_cast<Real4<R>*>(top_raw);          // [n8] level 8, then levels 9.. one after the other
        uint32_t* body8  = reinterpret_cast<uint32_t*>(lvl8 + 2 * n8 + 32); // body of the first leaf of every node, same layout
        for (int32_t j = threadIdx.x; j < n8; j += blockDim.x)
        {
            lvl8[j]  = ld4(&b.sphere[b.level_offset[8] + j]);
            body8[j] = static_cast<uint32_t>(b.keys_sorted[static_cast<int64_t>(j) << 8] >> 32);
        }
        __syncthreads();
        int32_t below_at = 0, below_n = n8, at = n8;
        for (int l = 9; l < b.n_levels; ++l)
        {
            int32_t const here = (below_n + 1) / 2;
            if (threadIdx.x == 0)
                s_top_offset[l - 9] = at;
            for (int32_t j = threadIdx.x; j < here; j += blockDim.x)
            {
                bool const has_c = 2 * j + 1 < below_n;
                uint32_t nb;
                lvl8[at + j]  = fit_pair<R>(lvl8[below_at + 2 * j], body8[below_at + 2 * j], has_c,
                                           lvl8[below_at + (has_c ? 2 * j + 1 : 2 * j)],
                                           body8[below_at + (has_c ? 2 * j + 1 : 2 * j)], nb);
                body8[at + j] = nb;
            }
            __syncthreads();
            below_at = at;
            below_n  = here;
            at += here;
        }
        s_top = lvl8;
    }
    int64_t const i   = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    bool const valid  = i < s.n_surface;
    int32_t n_mine    = 0;
    Vec3<R> p         = {R(0), R(0), R(0)};
    int32_t body      = -1;
    uint32_t culled   = 0u;
    if (valid)
    {
        Real4<R> const q = ld4(&s.surf_pos[i]);
        p                = {q.x, q.y, q.z};
        body             = s.surf_body[i]; // negative: the body was not handed to the cd system
        culled = b.n > 0 && body >= 0 ? bvh_cull_mask<R>(s, b, b.leaf_of_surface[i], s_top, s_top_offset) : 0u; // b.n == 0: no broadphase
        for (int32_t k = 0; body >= 0 && k < s.n_sdf; ++k)
        {
            Vec3<R> g;
            if (!(k < 32 && (culled >> k & 1u)) && sdf_eval<R>(s.sdf[k], p, g) < R(0))
                ++n_mine;
        }
    }
    // warp-aggregated reservation
    unsigned const lane = threadIdx.x & 31u;
    int32_t incl        = n_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int32_t const o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<unsigned>(d))
            incl += o;
    }
    int32_t const total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base       = 0;
    if (lane == 31 && total > 0)
        base = atomicAdd(s.contact_count, static_cast<uint32_t>(total));
    base = __shfl_sync(0xffffffffu, base, 31);
    if (valid)
        s.surf_first[i] = n_mine > 0 ? base + static_cast<uint32_t>(incl - n_mine) : 0xffffffffu;
    if (!valid || n_mine == 0)
        return;
    uint32_t slot      = base + static_cast<uint32_t>(incl - n_mine);
    uint32_t const gv  = s.surf_v[i];
    bool first         = true;
    for (int32_t k = 0; k < s.n_sdf; ++k)
    {
        Vec3<R> g;
        if (k < 32 && (culled >> k & 1u))
            continue;
        R const sd = sdf_eval<R>(s.sdf[k], p, g);
        if (!(sd < R(0)))
            continue;
        R const inv       = R(1) / sqrt_(dot(g, g)); // grad.normalized() (bvh_model.cpp:82)
        Vec3<R> const n   = {g.x * inv, g.y * inv, g.z * inv};
        R const a         = abs_(sd);
        if (slot < s.contact_cap)
        {
            s.contact_v[slot] = gv | (first ? 0x80000000u : 0u);
            st4(&s.contact_q[slot], Real4<R>{p.x + a * n.x, p.y + a * n.y, p.z + a * n.z, R(0)}); // :83-84
            st4(&s.contact_n[slot], Real4<R>{n.x, n.y, n.z, as_real(R(0), s.sdf[k].body)});
        }
        first = false;
        ++slot;
    }
    (void)body;
}


} // namespace sbsb200
