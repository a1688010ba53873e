// Persistent, region-resident schedule: ONE cooperative kernel per substep.
//
// The tet set is cut into spatially compact regions, one per CTA (one CTA per SM for a large
// connected mesh).  A vertex whose tets all lie in one region is INTERIOR to it: its (xi, w)
// record lives in that CTA's shared memory for the whole substep (predict, every iteration and
// colour, commit) and never touches L2/HBM in between.  Vertices shared by several regions are
// INTERFACE vertices and stay in global memory.  The Gauss-Seidel order is unchanged — colour
// major over the whole mesh — but instead of a grid-wide barrier per colour each region only
// waits for its NEIGHBOUR regions (those it shares an interface vertex with) through a
// monotone progress counter in global memory (release fence + relaxed store / relaxed poll +
// acquire fence, gpu scope).
//
//   step 0                      : predict   (timestep.cpp:35-43)
//   step 1 + k(1+C) + 0         : collision constraints of iteration k (gauss_seidel_solver.cpp:28-31)
//   step 1 + k(1+C) + 1 + c     : colour c of iteration k               (:32-35)
//   step 1 + K(1+C)             : commit    (timestep.cpp:48-57) + surface copy
//
// Before step j a region waits until every neighbour has published "steps < j done"; this
// orders both the read-after-write and the write-after-read hazards on interface vertices.
// Regions without any interface (independent bodies of an ensemble) skip all of it and are
// handed out round-robin, several per CTA.
#pragma once

#include "scene_build.h"
#include "xpbd_kernels.cuh"

#include <cooperative_groups.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace sbsb200 {

constexpr uint32_t kGlobalBit = 0x80000000u; // vertex address: global index when set, smem slot otherwise

template <typename R>
struct PersistentArgs
{
    DeviceScene<R> s;
    int32_t n_regions, n_colours, n_sync_regions, n_island_regions;
    int32_t const* sync_regions;   // regions with neighbours: one per CTA, CTA b takes sync_regions[b]
    int32_t const* island_regions; // regions without neighbours: CTA b takes island_regions[b + i*grid]
    uint4 const* tet_addr;         // per tet (schedule order): 4 vertex addresses
    DevChunk const* chunks;        // [n_colours * n_regions] clustered-colouring chunks, colour-major then region
    int32_t const* vtx_off;        // [n_regions + 1] interior vertex list
    uint32_t const* vtx;           // global vertex id of slot i
    int32_t const* ifv_off;        // [n_regions + 1] owned interface vertices
    uint32_t const* ifv;
    int32_t const* surf_off;       // [n_regions + 1] owned surface vertices
    uint32_t const* surf_index;    // surface vertex index (into s.surf_pos / s.surf_first)
    uint32_t const* surf_addr;     // its vertex address (slot or global)
    int32_t const* nbr_off;        // [n_regions + 1]
    int32_t const* nbr;
    uint32_t* progress;            // [n_regions] steps completed (monotone, wraps)
    uint32_t base;                 // progress value of every region when this launch starts
    int32_t iterations;
    int32_t collide;
    R dt;
};

// Progress flags: polled with relaxed loads, ONE acquire fence after the poll succeeds; published
// with one release fence followed by a relaxed store (a fence per poll iteration, or fence.sc via
// __threadfence(), costs several hundred cycles each on the per-step critical path).
__device__ __forceinline__ uint32_t ld_relaxed(uint32_t const* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// interface vertices are shared between SMs: bypass the (non-coherent) L1
__device__ __forceinline__ Real4<float> ld4_cg(Real4<float> const* p)
{
    float4 const v = __ldcg(reinterpret_cast<float4 const*>(p));
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Real4<double> ld4_cg(Real4<double> const* p)
{
    double2 const a = __ldcg(reinterpret_cast<double2 const*>(p));
    double2 const b = __ldcg(reinterpret_cast<double2 const*>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st4_cg(Real4<float>* p, Real4<float> v)
{
    __stcg(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));
}
__device__ __forceinline__ void st4_cg(Real4<double>* p, Real4<double> v)
{
    __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
    __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(v.z, v.w));
}

template <typename R>
__device__ __forceinline__ Real4<R> load_vertex(uint32_t addr, Real4<R> const* sx, Real4<R> const* pos)
{
    return (addr & kGlobalBit) ? ld4_cg(&pos[addr & ~kGlobalBit]) : sx[addr];
}
template <typename R>
__device__ __forceinline__ void store_vertex(uint32_t addr, Real4<R>* sx, Real4<R>* pos, Real4<R> v)
{
    if (addr & kGlobalBit)
        st4_cg(&pos[addr & ~kGlobalBit], v);
    else
        sx[addr] = v;
}

// collision constraints of one vertex, in list order (collision_constraint.cpp:21-48)
template <typename R>
__device__ __forceinline__ bool project_vertex_contacts(DeviceScene<R> const& s, uint32_t first, uint32_t n_contacts,
                                                        Real4<R>& p, R at, int first_iteration)
{
    bool moved        = false;
    uint32_t const gv = s.contact_v[first] & 0x7fffffffu;
    for (uint32_t j = first; j < n_contacts; ++j)
    {
        if (j > first && s.contact_v[j] != gv)
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
    return moved;
}

template <typename R>
__device__ void run_region(PersistentArgs<R> const& a, int32_t region, bool sync, Real4<R>* sx)
{
    DeviceScene<R> const& s = a.s;
    int const tid = threadIdx.x, nt = blockDim.x;
    R const dt              = a.dt;
    int32_t const v0 = a.vtx_off[region], nv = a.vtx_off[region + 1] - v0;
    int32_t const i0 = a.ifv_off[region], ni = a.ifv_off[region + 1] - i0;
    int32_t const n0 = a.nbr_off[region], nn = a.nbr_off[region + 1] - n0;
    uint32_t step = a.base;

    auto wait_neighbours = [&]() {
        if (sync)
        {
            if (tid < 32)
            {
                for (int32_t j = tid; j < nn; j += 32)
                {
                    uint32_t const* flag = &a.progress[a.nbr[n0 + j]];
                    while (static_cast<int32_t>(ld_relaxed(flag) - step) < 0)
                    {
                    }
                }
                fence_acq_rel_gpu();
            }
        }
        __syncthreads();
    };
    auto publish = [&]() {
        ++step;
        if (sync)
        {
            __syncthreads();
            if (tid == 0)
            {
                fence_acq_rel_gpu();
                st_relaxed(&a.progress[region], step);
            }
        }
    };

    // ---- step 0: predict --------------------------------------------------------------------
    for (int32_t i = tid; i < nv; i += nt)
    {
        uint32_t const gv = a.vtx[v0 + i];
        Real4<R> p        = ld4(&s.pos[gv]);
        Real4<R> const x  = ld4(&s.prev[gv]);
        Real4<R> v        = ld4(&s.vel[gv]);
        predict_vertex(p, x, v, dt);
        sx[i] = p;
    }
    for (int32_t i = tid; i < ni; i += nt)
    {
        uint32_t const gv = a.ifv[i0 + i];
        Real4<R> p        = ld4_cg(&s.pos[gv]);
        Real4<R> const x  = ld4(&s.prev[gv]);
        Real4<R> v        = ld4(&s.vel[gv]);
        predict_vertex(p, x, v, dt);
        st4_cg(&s.pos[gv], p);
    }
    publish();

    // ---- iterations ---------------------------------------------------------------------------
    uint32_t const n_contacts =
        a.collide ? min(*s.contact_count, static_cast<uint32_t>(s.contact_cap)) : 0u;
    R const at_c = s.collision_alpha / (dt * dt);
    int32_t const s0 = a.surf_off[region], ns = a.surf_off[region + 1] - s0;
    for (int32_t k = 0; k < a.iterations; ++k)
    {
        int const first_iteration = k == 0;
        if (a.collide)
        {
            wait_neighbours();
            if (n_contacts > 0)
                for (int32_t i = tid; i < ns; i += nt)
                {
                    uint32_t const first = s.surf_first[a.surf_index[s0 + i]];
                    if (first == 0xffffffffu)
                        continue;
                    uint32_t const addr = a.surf_addr[s0 + i];
                    Real4<R> p          = load_vertex(addr, sx, s.pos);
                    if (project_vertex_contacts(s, first, n_contacts, p, at_c, first_iteration))
                        store_vertex(addr, sx, s.pos, p);
                }
            publish();
        }
        for (int32_t c = 0; c < a.n_colours; ++c)
        {
            DevChunk const ch = a.chunks[c * a.n_regions + region];
            wait_neighbours();
            for (int32_t i = tid; i < ch.n[0]; i += nt)
            {
                int32_t base = ch.first;
#pragma unroll 1
                for (int m = 0; m < 8; ++m)
                {
                    if (i >= ch.n[m])
                        break;
                    int32_t const t = base + i;
                    base += ch.n[m];
                    uint4 const ad    = __ldg(&a.tet_addr[t]);
                    Real4<R> const r0 = ld4_ro(&s.tet_r0[t]);
                    Real4<R> const r1 = ld4_ro(&s.tet_r1[t]);
                    Real4<R> const r2 = ld4_ro(&s.tet_r2[t]);
                    Real4<R> p1 = load_vertex(ad.x, sx, s.pos), p2 = load_vertex(ad.y, sx, s.pos),
                             p3 = load_vertex(ad.z, sx, s.pos), p4 = load_vertex(ad.w, sx, s.pos);
                    Real4<R> const mat = ld4_ro(&s.materials[mat_index(r2.z)]);
                    R lambda           = first_iteration ? R(0) : s.tet_lambda[t];
                    R const lambda_in  = lambda;
                    Vec3<R> const z{};
                    green_project<R, false>(p1, p2, p3, p4, z, z, z, z, r0, r1, r2, mat, dt, lambda);
                    if (lambda != lambda_in || first_iteration)
                        s.tet_lambda[t] = lambda;
                    if (lambda != lambda_in)
                    {
                        store_vertex(ad.x, sx, s.pos, p1);
                        store_vertex(ad.y, sx, s.pos, p2);
                        store_vertex(ad.z, sx, s.pos, p3);
                        store_vertex(ad.w, sx, s.pos, p4);
                    }
                }
            }
            if (!sync)
                __syncthreads(); // publish() carries the barrier when syncing
            publish();
        }
    }

    // ---- last step: commit + surface copy -----------------------------------------------------
    wait_neighbours();
    for (int32_t i = tid; i < nv; i += nt)
    {
        uint32_t const gv = a.vtx[v0 + i];
        Real4<R> const p  = sx[i];
        Real4<R> xn       = ld4(&s.prev[gv]);
        Real4<R> v        = ld4(&s.vel[gv]);
        commit_vertex(p, xn, v, dt);
        st4(&s.vel[gv], v);
        st4(&s.prev[gv], xn);
    }
    for (int32_t i = tid; i < ni; i += nt)
    {
        uint32_t const gv = a.ifv[i0 + i];
        Real4<R> const p  = ld4_cg(&s.pos[gv]);
        Real4<R> xn       = ld4(&s.prev[gv]);
        Real4<R> v        = ld4(&s.vel[gv]);
        commit_vertex(p, xn, v, dt);
        st4(&s.vel[gv], v);
        st4(&s.prev[gv], xn);
    }
    // tetrahedral_body_t::update_visual_model (tetrahedral_body.cpp:157-165) for owned surface vertices
    for (int32_t i = tid; i < ns; i += nt)
    {
        uint32_t const addr = a.surf_addr[s0 + i];
        Real4<R> const p    = load_vertex(addr, sx, s.pos);
        st4(&s.surf_pos[a.surf_index[s0 + i]], Real4<R>{p.x, p.y, p.z, R(0)});
    }
    publish();
    __syncthreads(); // shared memory is reused by the next region of this CTA
}

template <typename R>
__global__ void __launch_bounds__(512) k_substep_persistent(PersistentArgs<R> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real4<R>* sx = reinterpret_cast<Real4<R>*>(smem_raw);
    if (static_cast<int32_t>(blockIdx.x) < a.n_sync_regions)
        run_region<R>(a, a.sync_regions[blockIdx.x], true, sx);
    for (int32_t i = blockIdx.x; i < a.n_island_regions; i += gridDim.x)
        run_region<R>(a, a.island_regions[i], false, sx);
}

template <typename T>
struct PBuf
{
    T* p = nullptr;
    ~PBuf()
    {
        if (p)
            cudaFree(p);
    }
    void upload(std::vector<T> const& h, cudaStream_t st)
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        if (cudaMalloc(&p, sizeof(T) * (h.empty() ? 1 : h.size())) != cudaSuccess)
            throw std::runtime_error("cudaMalloc failed in the persistent plan");
        if (!h.empty() &&
            cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st) != cudaSuccess)
            throw std::runtime_error("cudaMemcpyAsync failed in the persistent plan");
    }
};

template <typename R>
struct PersistentPlan
{
    PersistentArgs<R> args{};
    PBuf<int32_t> sync_regions, island_regions, vtx_off, ifv_off, surf_off, nbr_off, nbr;
    PBuf<DevChunk> chunks;
    PBuf<uint32_t> vtx, ifv, surf_index, surf_addr, progress;
    PBuf<uint4> tet_addr;
    int grid = 0, block = 0;
    size_t smem = 0;
    uint32_t base = 0;
    bool ready = false;
    std::string why_not;

    // how many regions to cut the scene into
    static int32_t regions_for(int sm_count, int64_t /*n_tets*/) { return sm_count; }
    // ensembles: many independent bodies -> one region per body
    static bool wants_region_per_body(HostScene const& h, int sm_count)
    {
        int64_t bodies = 0;
        for (auto const& b : h.bodies)
            bodies += (b.kind == BodyKind::tet && b.n_tets > 0);
        return bodies >= 2 * static_cast<int64_t>(sm_count);
    }

    // returns false (with why_not) when the scene does not fit this schedule
    bool build(HostScene const& h, ClusterPlan const& cp, RegionPlan const& plan, DeviceScene<R> const& d,
               cudaStream_t st, int sm_count)
    {
        int32_t const Rn = plan.n_regions;
        int64_t const T = h.n_tets(), V = h.n_vertices();
        if (T >= (int64_t{1} << 31) / 2)
        {
            why_not = "too many tets for 32-bit chunk offsets";
            return false;
        }
        // tets in storage order -> vertex addresses; chunk descriptors
        std::vector<uint4> addr(static_cast<size_t>(T));
        for (int64_t p = 0; p < T; ++p)
        {
            uint32_t const t = cp.storage_order[static_cast<size_t>(p)];
            int32_t const r  = plan.tet_region[t];
            uint32_t ad[4];
            for (int k = 0; k < 4; ++k)
            {
                uint32_t const gv = h.tets[4 * static_cast<size_t>(t) + k];
                ad[k] = plan.vertex_region[gv] == r ? plan.vertex_slot[gv] : (gv | kGlobalBit);
            }
            addr[static_cast<size_t>(p)] = make_uint4(ad[0], ad[1], ad[2], ad[3]);
        }
        std::vector<DevChunk> hchunks(cp.chunks.size());
        for (size_t i = 0; i < cp.chunks.size(); ++i)
        {
            hchunks[i].first = cp.chunks[i].first;
            for (int m = 0; m < 8; ++m)
                hchunks[i].n[m] = cp.chunks[i].n[m];
        }
        if (cp.n_regions != Rn)
        {
            why_not = "cluster plan and region plan disagree";
            return false;
        }
        // owned interface vertices, owned surface vertices
        std::vector<int32_t> ioff(static_cast<size_t>(Rn) + 1, 0), soff(static_cast<size_t>(Rn) + 1, 0);
        for (int64_t v = 0; v < V; ++v)
            if (plan.vertex_region[static_cast<size_t>(v)] < 0)
                ++ioff[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)]) + 1];
        for (int32_t r = 0; r < Rn; ++r)
            ioff[static_cast<size_t>(r) + 1] += ioff[static_cast<size_t>(r)];
        std::vector<uint32_t> ifv(static_cast<size_t>(ioff.back()));
        {
            std::vector<int32_t> cur(ioff.begin(), ioff.end() - 1);
            for (int64_t v = 0; v < V; ++v)
                if (plan.vertex_region[static_cast<size_t>(v)] < 0)
                    ifv[static_cast<size_t>(cur[static_cast<size_t>(plan.vertex_owner[static_cast<size_t>(v)])]++)] =
                        static_cast<uint32_t>(v);
        }
        std::vector<uint32_t> sgv; // global vertex of surface vertex i (same order as DeviceScene::surf_v)
        for (auto const& b : h.bodies)
            if (b.kind == BodyKind::tet)
                for (uint32_t lv : b.surf_to_tet)
                    sgv.push_back(static_cast<uint32_t>(b.v_offset + lv));
        for (uint32_t gv : sgv)
            ++soff[static_cast<size_t>(plan.vertex_owner[gv]) + 1];
        for (int32_t r = 0; r < Rn; ++r)
            soff[static_cast<size_t>(r) + 1] += soff[static_cast<size_t>(r)];
        std::vector<uint32_t> sidx(sgv.size()), sadr(sgv.size());
        {
            std::vector<int32_t> cur(soff.begin(), soff.end() - 1);
            for (size_t i = 0; i < sgv.size(); ++i)
            {
                uint32_t const gv = sgv[i];
                int32_t const r   = plan.vertex_owner[gv];
                int32_t const pos = cur[static_cast<size_t>(r)]++;
                sidx[static_cast<size_t>(pos)] = static_cast<uint32_t>(i);
                sadr[static_cast<size_t>(pos)] =
                    plan.vertex_region[gv] == r ? plan.vertex_slot[gv] : (gv | kGlobalBit);
            }
        }
        // sync regions (have neighbours) vs islands
        std::vector<int32_t> sync_r, island_r;
        for (int32_t r = 0; r < Rn; ++r)
            (plan.nbr_offsets[static_cast<size_t>(r) + 1] > plan.nbr_offsets[static_cast<size_t>(r)] ? sync_r : island_r)
                .push_back(r);
        std::vector<int32_t> voff(plan.region_vtx_offsets.begin(), plan.region_vtx_offsets.end());

        // launch shape
        int64_t const max_chunk = cp.max_chunk_clusters;
        smem  = static_cast<size_t>(std::max<int64_t>(plan.max_region_vertices, 1)) * sizeof(Real4<R>);
        block = static_cast<int>(std::min<int64_t>(512, std::max<int64_t>(64, (max_chunk + 31) / 32 * 32)));
        if (smem > 220 * 1024)
        {
            why_not = "a region's interior vertices do not fit shared memory";
            return false;
        }
        if (cudaFuncSetAttribute(k_substep_persistent<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)) != cudaSuccess)
        {
            why_not = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
            return false;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_substep_persistent<R>, block, smem) != cudaSuccess ||
            per_sm < 1)
        {
            why_not = "kernel does not fit an SM with this block/shared-memory size";
            return false;
        }
        int64_t const capacity = static_cast<int64_t>(per_sm) * sm_count;
        if (static_cast<int64_t>(sync_r.size()) > capacity)
        {
            why_not = "more synchronising regions than co-resident CTAs";
            return false;
        }
        grid = static_cast<int>(std::max<int64_t>(
            1, std::max<int64_t>(static_cast<int64_t>(sync_r.size()),
                                 std::min<int64_t>(capacity, static_cast<int64_t>(island_r.size())))));

        sync_regions.upload(sync_r, st);
        island_regions.upload(island_r, st);
        tet_addr.upload(addr, st);
        chunks.upload(hchunks, st);
        vtx_off.upload(voff, st);
        vtx.upload(plan.region_vtx, st);
        ifv_off.upload(ioff, st);
        this->ifv.upload(ifv, st);
        surf_off.upload(soff, st);
        surf_index.upload(sidx, st);
        surf_addr.upload(sadr, st);
        nbr_off.upload(plan.nbr_offsets, st);
        nbr.upload(plan.nbr, st);
        progress.upload(std::vector<uint32_t>(static_cast<size_t>(Rn), 0u), st);
        base = 0;

        args.s                = d;
        args.n_regions        = Rn;
        args.n_colours        = cp.n_colours;
        args.n_sync_regions   = static_cast<int32_t>(sync_r.size());
        args.n_island_regions = static_cast<int32_t>(island_r.size());
        args.sync_regions     = sync_regions.p;
        args.island_regions   = island_regions.p;
        args.tet_addr         = tet_addr.p;
        args.chunks           = chunks.p;
        args.vtx_off          = vtx_off.p;
        args.vtx              = vtx.p;
        args.ifv_off          = ifv_off.p;
        args.ifv              = this->ifv.p;
        args.surf_off         = surf_off.p;
        args.surf_index       = surf_index.p;
        args.surf_addr        = surf_addr.p;
        args.nbr_off          = nbr_off.p;
        args.nbr              = nbr.p;
        args.progress         = progress.p;
        ready                 = true;
        return true;
    }

    // steps published per substep launch
    uint32_t steps_per_substep(int iterations, bool collide) const
    {
        return 2u + static_cast<uint32_t>(iterations) * (static_cast<uint32_t>(args.n_colours) + (collide ? 1u : 0u));
    }

    // enqueue one substep; returns kernels launched
    int64_t substep(DeviceScene<R> const& d, R dt, int iterations, bool collide, cudaStream_t st)
    {
        args.s          = d;
        args.dt         = dt;
        args.iterations = iterations;
        args.collide    = collide ? 1 : 0;
        args.base       = base;
        void* params[]  = {&args};
        cudaError_t const e = cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(k_substep_persistent<R>),
                                                          dim3(static_cast<unsigned>(grid)), dim3(static_cast<unsigned>(block)),
                                                          params, smem, st);
        if (e != cudaSuccess)
            throw std::runtime_error(std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(e));
        base += steps_per_substep(iterations, collide);
        return 1;
    }
};

} // namespace sbsb200
