// Resident schedule: ONE cooperative kernel per substep, every region's vertices in shared memory.
//
// The tet set is cut into regions, one per CTA (one CTA per SM for a large connected mesh; ensembles put a
// few whole bodies into a region and hand regions out to the CTAs).  A region keeps ONE shared-memory slot
// for every vertex its tets touch — the vertices it owns first (it predicts, collides and commits them), then
// its guests — and the tet records hold 16-bit slots, so the projection loop reads and writes shared memory
// only.  Work unit = CLUSTER (<= 8 tets sharing vertices), one thread each; clusters of one colour share no
// vertex; a step = one colour of one region, a __syncthreads between steps.
//
// A vertex that only one region touches never leaves its shared memory during a substep.  A vertex that
// several regions touch TRAVELS, and only when it has to: which steps touch a vertex is static (predict, per
// iteration the collision step of its owner if it is a surface vertex and collision steps exist, the colours
// of the clusters that contain it, commit), so every touch knows who touched the vertex before and who
// touches it next (scene_build.h, ExchangePlan).  Next touch by another region: the position is PUSHED into a
// mailbox of that region, tagged with the step.  Previous touch by another region: it is PULLED — polled for
// the expected tag — from the own mailbox into the slot.  Both by the same region: nothing happens, the value
// stays in the slot.  A mailbox is one 16-byte word {x, y, z, tag} (fp32; three {value, tag} words in fp64)
// read and written with single 128-bit accesses, so position and flag arrive together: no fence, no separate
// flag, no barrier between regions.  Every touch is a read-modify-write that waits for the previous one, which
// orders read-after-write and write-after-read hazards alike.  A mailbox may sit in another GPU's memory: a
// push is then a system-scope store over NVLink issued by this kernel (compute and exchange are one kernel).
//
// With pencil-shaped regions and the colour order that goes with them (scene_build.cpp) the steps that flip
// the cell parity along the pencil axis — half of the steps of a sweep on a lattice — pull nothing at all.
//
//   step 0                      : predict   (timestep.cpp:35-43)
//   step 1 + k(1+C) + 0         : collision constraints of iteration k (gauss_seidel_solver.cpp:28-31)
//   step 1 + k(1+C) + 1 + c     : colour c of iteration k               (:32-35)
//   step 1 + K(1+C)             : commit    (timestep.cpp:48-57) + surface copy
//
// Regions that exchange are co-resident (cooperative launch), so waiting cannot deadlock: every wait is for a
// strictly earlier step.  A poll budget turns a lost update into an error code instead of a hang.
#pragma once

#include "scene_build.h"
#include "xpbd_kernels.cuh"

#include <algorithm>
#include <cooperative_groups.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace sbsb200 {

constexpr uint32_t kBoxIndexMask = kRouteIndexMask;
constexpr int kRankShift         = kRouteRankShift;
constexpr int kMaxWorld          = 8;
constexpr int kPollBudget     = 1 << 24;     // polls of one record before the kernel gives up
// Polls of one record in flight (fp32 pulls of a cluster), kPollGap ns apart.  Measured on config 3, one box, ms per frame
// (profiles/r02_ab_polls.txt): 1 generation 4.00, 3 generations 4.32 (gap 128 ns) / 4.22 (gap 40 ns), 1 generation with
// 300 ns of sleep between rounds 4.12, only the last record of a group polled while waiting 4.44 (5.43 with 3 generations).
// On another box 1 and 3 generations both gave 4.00: more polls in flight never help and often hurt (every poll is a sector
// request to the L2 slice that also has to take the neighbours' pushes), fewer polls see the record later.
#ifndef SBSB200_POLL_GENERATIONS
#define SBSB200_POLL_GENERATIONS 1
#endif
constexpr int kPollGenerations = SBSB200_POLL_GENERATIONS;
constexpr unsigned kPollGap    = 128;

// true when the wait should be abandoned: budget exhausted (sets the error flag) or another thread
// already gave up (checked every 1024 polls so that one lost update cannot stall the whole launch)
__device__ __forceinline__ bool poll_expired(uint32_t* error, int polls)
{
    if (polls > kPollBudget)
    {
        *reinterpret_cast<volatile uint32_t*>(error) = 1u;
        return true;
    }
    return (polls & 1023) == 0 && *reinterpret_cast<volatile uint32_t*>(error) != 0u;
}

// ---- 128-bit single-copy-atomic accesses (LDG/STG.E.128.STRONG.GPU) ----------------------------
struct Word128
{
    unsigned long long lo, hi;
};
// No "memory" clobber: a record carries its own flag and orders nothing else, and a clobber would
// make the compiler serialise the polls of a fetch list (each load waiting for the shared-memory
// store of the previous one) — measured at 8 x 600 cycles per cluster.
__device__ __forceinline__ Word128 ld_b128(void const* p)
{
    Word128 w;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(w.lo), "=l"(w.hi)
                 : "l"(p));
    return w;
}
__device__ __forceinline__ void st_b128(void* p, Word128 w)
{
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p),
                 "l"(w.lo), "l"(w.hi));
}

// system-scope variants: mailboxes in another GPU's memory (stores over NVLink), and polls of
// mailboxes a peer GPU writes
__device__ __forceinline__ Word128 ld_b128_sys(void const* p)
{
    Word128 w;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.sys.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(w.lo), "=l"(w.hi)
                 : "l"(p));
    return w;
}
__device__ __forceinline__ void st_b128_sys(void* p, Word128 w)
{
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.sys.global.b128 [%0], q;\n\t}" ::"l"(p),
                 "l"(w.lo), "l"(w.hi));
}
__device__ __forceinline__ Word128 ld_b128_scoped(void const* p, bool sys) { return sys ? ld_b128_sys(p) : ld_b128(p); }
__device__ __forceinline__ void st_b128_scoped(void* p, Word128 w, bool sys)
{
    if (sys)
        st_b128_sys(p, w);
    else
        st_b128(p, w);
}

// Exchange records.  fp32: one word {x, y, z, tag} per vertex.  fp64: three words {value, tag}.
template <typename R>
struct Xchg;
template <>
struct Xchg<float>
{
    static constexpr int kWords = 1;
    struct Raw
    {
        Word128 w;
    };
    static __device__ __forceinline__ Raw fetch(void const* base, uint32_t gv, bool sys = false)
    {
        return Raw{ld_b128_scoped(static_cast<char const*>(base) + 16ull * gv, sys)};
    }
    static __device__ __forceinline__ bool decode(Raw const& r, uint32_t expect, float& x, float& y, float& z)
    {
        x = __uint_as_float(static_cast<uint32_t>(r.w.lo));
        y = __uint_as_float(static_cast<uint32_t>(r.w.lo >> 32));
        z = __uint_as_float(static_cast<uint32_t>(r.w.hi));
        return static_cast<uint32_t>(r.w.hi >> 32) == expect;
    }
    static __device__ __forceinline__ bool load(void const* base, uint32_t gv, uint32_t expect, float& x, float& y,
                                                float& z, bool sys = false)
    {
        Word128 const w = ld_b128_scoped(static_cast<char const*>(base) + 16ull * gv, sys);
        x               = __uint_as_float(static_cast<uint32_t>(w.lo));
        y               = __uint_as_float(static_cast<uint32_t>(w.lo >> 32));
        z               = __uint_as_float(static_cast<uint32_t>(w.hi));
        return static_cast<uint32_t>(w.hi >> 32) == expect;
    }
    static __device__ __forceinline__ void store(void* base, uint32_t gv, float x, float y, float z, uint32_t tag,
                                                 bool sys = false)
    {
        Word128 w;
        w.lo = static_cast<unsigned long long>(__float_as_uint(x)) |
               (static_cast<unsigned long long>(__float_as_uint(y)) << 32);
        w.hi = static_cast<unsigned long long>(__float_as_uint(z)) | (static_cast<unsigned long long>(tag) << 32);
        st_b128_scoped(static_cast<char*>(base) + 16ull * gv, w, sys);
    }
};
template <>
struct Xchg<double>
{
    static constexpr int kWords = 3;
    struct Raw
    {
        Word128 w0, w1, w2;
    };
    static __device__ __forceinline__ Raw fetch(void const* base, uint32_t gv, bool sys = false)
    {
        char const* p = static_cast<char const*>(base) + 48ull * gv;
        return Raw{ld_b128_scoped(p, sys), ld_b128_scoped(p + 16, sys), ld_b128_scoped(p + 32, sys)};
    }
    static __device__ __forceinline__ bool decode(Raw const& r, uint32_t expect, double& x, double& y, double& z)
    {
        x = __longlong_as_double(static_cast<long long>(r.w0.lo));
        y = __longlong_as_double(static_cast<long long>(r.w1.lo));
        z = __longlong_as_double(static_cast<long long>(r.w2.lo));
        return static_cast<uint32_t>(r.w0.hi) == expect && static_cast<uint32_t>(r.w1.hi) == expect &&
               static_cast<uint32_t>(r.w2.hi) == expect;
    }
    static __device__ __forceinline__ bool load(void const* base, uint32_t gv, uint32_t expect, double& x, double& y,
                                                double& z, bool sys = false)
    {
        char const* p    = static_cast<char const*>(base) + 48ull * gv;
        Word128 const w0 = ld_b128_scoped(p, sys), w1 = ld_b128_scoped(p + 16, sys), w2 = ld_b128_scoped(p + 32, sys);
        x                = __longlong_as_double(static_cast<long long>(w0.lo));
        y                = __longlong_as_double(static_cast<long long>(w1.lo));
        z                = __longlong_as_double(static_cast<long long>(w2.lo));
        return static_cast<uint32_t>(w0.hi) == expect && static_cast<uint32_t>(w1.hi) == expect &&
               static_cast<uint32_t>(w2.hi) == expect;
    }
    static __device__ __forceinline__ void store(void* base, uint32_t gv, double x, double y, double z, uint32_t tag,
                                                 bool sys = false)
    {
        char* p = static_cast<char*>(base) + 48ull * gv;
        st_b128_scoped(p, Word128{static_cast<unsigned long long>(__double_as_longlong(x)), tag}, sys);
        st_b128_scoped(p + 16, Word128{static_cast<unsigned long long>(__double_as_longlong(y)), tag}, sys);
        st_b128_scoped(p + 32, Word128{static_cast<unsigned long long>(__double_as_longlong(z)), tag}, sys);
    }
};

template <typename R>
struct ResidentArgs
{
    DeviceScene<R> s;
    int32_t n_regions, n_colours;
    int32_t n_run;                 // regions this launch runs (entries of region_order)
    int32_t rot;                   // cluster i of a step runs on thread (i + rot) % nt (ClusterPlan::rot)
    int32_t const* region_order;   // CTA b runs region_order[b], [b + grid], ...; regions that exchange
                                   // vertices come first, one per CTA
    uint2 const* tet_slots;        // per tet (storage order): 4 x u16 slots into the region's vertex table
    // rest-shape dictionary (kDict kernels): meshes with few distinct rest shapes — every lattice has ten —
    // keep the (DmInv, V0, material) records in shared memory and stream one byte per tet instead of 48
    uint8_t const* tet_shape;      // per tet: index into shapes
    Real4<R> const* shapes;        // [kShapeWords * n_shapes]: r0, r1, r2 and the material of every distinct record
    int32_t n_shapes;
    DevChunk const* chunks;        // [(colour * n_regions + region) * 2 + part]
    // local vertex tables (ExchangePlan)
    int32_t const* loc_off;        // [n_regions + 1]
    int32_t const* n_owned;        // [n_regions]
    uint32_t const* loc_vtx;       // global vertex of slot i
    // exchange clusters: part 0 of every chunk, numbered chunk_xfirst[colour * n_regions + region] + i
    int32_t const* chunk_xfirst;
    int64_t n_xclusters;
    int32_t entries;               // mailboxes per exchange cluster (multiple of 4)
    uint4 const* pull;             // [variant][entries / 4][n_xclusters]: four pull words each, valid ones first
    uint4 const* push;             // [variant][entries / 2][n_xclusters]: {slot, route, slot, route}
    // owner side: shared vertices and surface vertices a region owns
    int32_t const* osv_off;        // [n_regions + 1]
    uint32_t const* osv_slot;
    uint32_t const* osv_meta;      // last colour | kOsvSurface | kOsvLastRemote | kOsvFirstRemote
    uint32_t const* osv_first;     // routing word of the first cluster entry touching the vertex in a sweep
    int32_t const* surf_off;       // [n_regions + 1]
    uint32_t const* surf_slot;
    uint32_t const* surf_index;    // surface vertex index (into s.surf_pos / s.surf_first)
    uint32_t const* surf_osv;      // position in osv, kRouteNone when only this region touches the vertex
    void* box;                     // mailboxes: n_entries of the clusters, then one per shared vertex for its owner
    uint32_t n_entries;            // entries * n_xclusters
    // decomposition over GPUs: every rank plans the same regions and runs its own block of them; a mailbox
    // lives on the rank that reads it, pushes to other ranks are peer stores
    int32_t rank, world;
    void* box_of_rank[kMaxWorld];  // mailbox arrays of all ranks (peer-mapped); [rank] == box
    uint32_t* error;               // set to 1 when a poll budget ran out
    uint32_t base;                 // tag of step 0 of this launch
    long long* trace;              // development aid: per-step clock stamps of one thread (nullptr = off)
    int32_t trace_steps;           // colour steps recorded per region
    int32_t iterations;
    int32_t collide;
    R dt;
};

// Push a position into the mailbox a routing word names (index + rank), tagged with the step.
template <typename R>
__device__ __forceinline__ void push(ResidentArgs<R> const& a, uint32_t route, R x, R y, R z, uint32_t tag)
{
    uint32_t const index = route & kBoxIndexMask;
    if (a.world > 1)
        Xchg<R>::store(a.box_of_rank[(route >> kRankShift) & 7u], index, x, y, z, tag, true);
    else
        Xchg<R>::store(a.box, index, x, y, z, tag);
}

// Wait for `expect` on (local) mailbox `b` and return the position in it.
template <typename R>
__device__ __forceinline__ void xchg_wait(ResidentArgs<R> const& a, uint32_t b, uint32_t expect, R& x, R& y, R& z)
{
    int polls = 0;
    while (!Xchg<R>::load(a.box, b, expect, x, y, z, a.world > 1))
    {
        if (poll_expired(a.error, ++polls))
            break;
        __nanosleep(20);
    }
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

constexpr int kMaxShapes = 256;
constexpr int kTraceWarps = 12; // warps per CTA the trace buffer has room for

// per-tet record as the projection consumes it
template <typename R, bool kDict>
struct TetRecord
{
    uint2 slots;
    Real4<R> r0, r1, r2;
    R lambda;
};
template <typename R>
struct TetRecord<R, true>
{
    uint2 slots;
    uint32_t shape; // r0, r1, r2 come out of the shared-memory dictionary when the tet runs
    R lambda;
};

template <typename R, bool kDict>
__device__ __forceinline__ TetRecord<R, kDict> load_tet(ResidentArgs<R> const& a, uint2 const* slots, int32_t t,
                                                        int first_iteration)
{
    TetRecord<R, kDict> q;
    q.slots = __ldg(&slots[t]);
    if constexpr (kDict)
        q.shape = __ldg(&a.tet_shape[t]);
    else
    {
        q.r0 = ld4_ro(&a.s.tet_r0[t]);
        q.r1 = ld4_ro(&a.s.tet_r1[t]);
        q.r2 = ld4_ro(&a.s.tet_r2[t]);
    }
    q.lambda = first_iteration ? R(0) : a.s.tet_lambda[t];
    return q;
}

__device__ __forceinline__ long long clock_stamp()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

// dynamic shared memory: [chunk descriptors of the region | first exchange cluster per colour | rest-shape dictionary | vertex table]
__host__ __device__ inline size_t chunk_area_bytes(int n_colours)
{
    return (static_cast<size_t>(n_colours) * 2 * sizeof(DevChunk) + 31) / 32 * 32;
}
__host__ __device__ inline size_t xfirst_area_bytes(int n_colours)
{
    return (static_cast<size_t>(n_colours) * sizeof(int32_t) + 31) / 32 * 32;
}
template <typename R>
__host__ __device__ inline size_t dict_area_bytes(int n_shapes)
{
    return static_cast<size_t>(kShapeWords * n_shapes) * sizeof(Real4<R>);
}

// What a thread keeps of a cluster between the moment it is prepared (head loaded, shared vertices pulled,
// normally one step ahead) and the moment it runs.
template <typename R, bool kDict>
struct ClusterHead
{
    TetRecord<R, kDict> tet0;
    R mu, lam, at; // material of the cluster's body; at = alpha / dt^2
};
template <typename R>
struct ClusterHead<R, true>
{
    TetRecord<R, true> tet0; // the material comes out of the dictionary with the rest shape
};

template <typename R, bool kDict>
__device__ __forceinline__ void load_cluster_head(ClusterHead<R, kDict>& h, ResidentArgs<R> const& a, DevChunk const& ch,
                                                  int32_t i, int first_iteration, Real4<R> const* s_dict)
{
    h.tet0 = load_tet<R, kDict>(a, a.tet_slots, ch.first + i, first_iteration);
    if constexpr (!kDict)
    {
        Real4<R> const mat = ld4_ro(&a.s.materials[mat_index(h.tet0.r2.z)]);
        h.mu               = mat.x;
        h.lam              = mat.y;
        h.at               = mat.z / (a.dt * a.dt);
    }
}

// Pull the shared vertices of a cluster whose previous touch was by another region: every pull word names
// the slot, the mailbox entry and how many steps back the previous touch lies; all polls of a group of four
// are in flight together.  `w` = the first group (loaded a step ahead), further groups are rare.
template <typename R, typename Stamp>
__device__ __forceinline__ void pull_cluster(ResidentArgs<R> const& a, uint4 const* pull_variant, int64_t xq, uint4 w,
                                             Real4<R>* sx, uint32_t tag, Stamp&& stamp)
{
    uint32_t const mine = static_cast<uint32_t>(xq); // mailbox of entry e: e * n_xclusters + xq
    uint32_t const nx   = static_cast<uint32_t>(a.n_xclusters);
    int const groups    = a.entries / 4;
    int polls           = 0;
    stamp(4);
    for (int g = 0;;)
    {
        uint32_t const word[4] = {w.x, w.y, w.z, w.w};
        uint32_t pending       = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (word[e] & kPullValid)
                pending |= 1u << e;
        bool const full = pending == 0xfu;
        // all polls of a group are in flight together (kPollGenerations of each, see there)
        constexpr int kGen = Xchg<R>::kWords == 1 ? kPollGenerations : 1;
        auto const poll_until_there = [&](uint32_t mask) {
            typename Xchg<R>::Raw raw[kGen][4];
#pragma unroll
            for (int g2 = 0; g2 < kGen; ++g2)
            {
                if (g2 > 0)
                    __nanosleep(kPollGap);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (mask >> e & 1u)
                        raw[g2][e] = Xchg<R>::fetch(a.box, (word[e] >> 24 & 0xfu) * nx + mine, a.world > 1);
            }
            while (mask)
            {
#pragma unroll
                for (int g2 = 0; g2 < kGen; ++g2)
                {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (mask >> e & 1u)
                        {
                            uint32_t const d = word[e] >> 16 & 0xffu;
                            R x, y, z;
                            if (Xchg<R>::decode(raw[g2][e], d == kPullPredict ? a.base : tag - d, x, y, z))
                            {
                                Real4<R>* dst = &sx[word[e] & 0xffffu];
                                dst->x        = x;
                                dst->y        = y;
                                dst->z        = z;
                                mask &= ~(1u << e);
                            }
                        }
                    if (polls == 0)
                        stamp(2);
                    if (!mask || poll_expired(a.error, ++polls))
                    {
                        mask = 0;
                        break;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (mask >> e & 1u)
                            raw[g2][e] = Xchg<R>::fetch(a.box, (word[e] >> 24 & 0xfu) * nx + mine, a.world > 1);
                }
            }
        };
        if (pending)
            poll_until_there(pending);
        if (!full || ++g >= groups)
            break;
        w = __ldg(&pull_variant[static_cast<int64_t>(g) * a.n_xclusters + xq]);
    }
    stamp(-(polls + 1)); // slot 1 <- number of poll rounds
}

// Push the shared vertices of the cluster that just ran whose next touch is by another region.
// w0, w1 = the first two groups of two (loaded before the tets ran).  The positions of a batch are read first
// and the stores issued back to back.
template <typename R>
__device__ __forceinline__ void push_cluster(ResidentArgs<R> const& a, uint4 const* push_variant, int64_t xq, uint4 w0,
                                             uint4 w1, Real4<R> const* sx, uint32_t tag)
{
    if (!(w0.x & kPullValid))
        return;
    {
        uint32_t const slot[4]  = {w0.x, w0.z, w1.x, w1.z};
        uint32_t const route[4] = {w0.y, w0.w, w1.y, w1.w};
        Real4<R> p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) // an empty entry re-reads the first one (it would name slot 0, which another thread may
                                    // be writing in this step); the loads stay unconditional, which the register
                                    // allocation of the 384-thread instantiation depends on (48 bytes of spills otherwise)
            p[e] = sx[((slot[e] & kPullValid) ? slot[e] : slot[0]) & 0xffffu];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (slot[e] & kPullValid)
                push<R>(a, route[e], p[e].x, p[e].y, p[e].z, tag);
    }
    if (!(w1.z & kPullValid))
        return;
    int const groups = a.entries / 2;
    for (int g = 2; g < groups; ++g)
    {
        uint4 const w = __ldg(&push_variant[static_cast<int64_t>(g) * a.n_xclusters + xq]);
        if (!(w.x & kPullValid))
            break;
        Real4<R> const p = sx[w.x & 0xffffu];
        push<R>(a, w.y, p.x, p.y, p.z, tag);
        if (!(w.z & kPullValid))
            break;
        Real4<R> const q = sx[w.z & 0xffffu];
        push<R>(a, w.w, q.x, q.y, q.z, tag);
    }
}

// One cluster: project its tets in order out of shared memory (slots of the region's vertex table).
template <typename R, bool kDict, typename Stamp>
__device__ __forceinline__ void run_cluster(ResidentArgs<R> const& a, DevChunk const& ch, int32_t i,
                                            ClusterHead<R, kDict> const& head, Real4<R>* sx, Real4<R> const* s_dict,
                                            int first_iteration, R inv_dt2, Stamp&& stamp)
{
    DeviceScene<R> const& s = a.s;
    // column layout: tet m of cluster i sits at first + n[0] + .. + n[m-1] + i
    int32_t n0 = ch.n[0], n1 = ch.n[1], n2 = ch.n[2], n3 = ch.n[3], n4 = ch.n[4], n5 = ch.n[5], n6 = ch.n[6],
            n7 = ch.n[7];
    int32_t t               = ch.first + i;
    TetRecord<R, kDict> cur = head.tet0;
    int tslot               = 8;
    stamp(tslot++);
#pragma unroll 1
    for (;;)
    {
        bool const more = i < n1;
        TetRecord<R, kDict> nxt;
        if (more)
            nxt = load_tet<R, kDict>(a, a.tet_slots, t + n0, first_iteration);
        uint32_t const a1 = cur.slots.x & 0xffffu, a2 = cur.slots.x >> 16, a3 = cur.slots.y & 0xffffu,
                       a4 = cur.slots.y >> 16;
        Real4<R> p1 = sx[a1], p2 = sx[a2], p3 = sx[a3], p4 = sx[a4];
        R lambda = cur.lambda;
        Vec3<R> const z{};
        Real4<R> r0, r1, r2;
        R mu, lam, at;
        if constexpr (kDict)
        {
            r0                 = s_dict[kShapeWords * cur.shape];
            r1                 = s_dict[kShapeWords * cur.shape + 1];
            r2                 = s_dict[kShapeWords * cur.shape + 2];
            Real4<R> const mat = s_dict[kShapeWords * cur.shape + 3];
            mu                 = mat.x;
            lam                = mat.y;
            at                 = mat.z * inv_dt2; // alpha / dt^2 (green_constraint.cpp:134)
        }
        else
        {
            r0  = cur.r0;
            r1  = cur.r1;
            r2  = cur.r2;
            mu  = head.mu;
            lam = head.lam;
            at  = head.at;
        }
        green_project_at<R, false>(p1, p2, p3, p4, z, z, z, z, r0, r1, r2, mu, lam, at, R(0), a.dt, lambda);
        if (lambda != cur.lambda || first_iteration)
            s.tet_lambda[t] = lambda;
        if (lambda != cur.lambda)
        {
            sx[a1] = p1;
            sx[a2] = p2;
            sx[a3] = p3;
            sx[a4] = p4;
        }
        stamp(tslot++);
        if (!more)
            break;
        t += n0;
        n0  = n1; n1 = n2; n2 = n3; n3 = n4; n4 = n5; n5 = n6; n6 = n7; n7 = 0;
        cur = nxt;
    }
}

// Predict (timestep.cpp:35-43) of the vertices a region owns, into its vertex table; guests only need their inverse
// mass.  Out of line: the loops below want many registers for a short time (four vertices per thread and trip: the
// index loads, then the state loads, are in flight together), which the colour steps must not pay for.
template <typename R, int kBatch>
__device__ __forceinline__ void predict_owned(Real4<R> const* pos, Real4<R> const* prev, Real4<R> const* vel,
                                           uint32_t const* loc_vtx, R dt, Real4<R>* sx, int32_t nl, int32_t no)
{
    struct
    {
        Real4<R> const *pos, *prev, *vel;
    } s{pos, prev, vel};
    struct
    {
        uint32_t const* loc_vtx;
    } a{loc_vtx};
    int32_t const l0 = 0;
    int const tid = threadIdx.x, nt = blockDim.x;
    // (four vertices per thread and trip: the index loads, then the state loads, are in flight together)
    for (int32_t i0 = tid; i0 < nl; i0 += kBatch * nt)
    {
        uint32_t gv[kBatch];
        Real4<R> pp[kBatch], x[kBatch], v[kBatch];
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            gv[e] = i0 + e * nt < nl ? a.loc_vtx[l0 + i0 + e * nt] : 0u;
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            if (i0 + e * nt < nl)
            {
                pp[e] = ld4(&s.pos[gv[e]]);
                if (i0 + e * nt < no)
                {
                    x[e] = ld4(&s.prev[gv[e]]);
                    v[e] = ld4(&s.vel[gv[e]]);
                }
            }
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            if (i0 + e * nt < nl)
            {
                if (i0 + e * nt < no)
                    predict_vertex(pp[e], x[e], v[e], dt);
                sx[i0 + e * nt] = pp[e];
            }
    }
}

// Commit (timestep.cpp:48-57) of the vertices a region owns, out of its vertex table.
template <typename R, int kBatch>
__device__ __forceinline__ void commit_owned(Real4<R>* prev, Real4<R>* vel, uint32_t const* loc_vtx, R dt,
                                          Real4<R> const* sx, int32_t no)
{
    struct
    {
        Real4<R> *prev, *vel;
    } s{prev, vel};
    struct
    {
        uint32_t const* loc_vtx;
    } a{loc_vtx};
    int32_t const l0 = 0;
    int const tid = threadIdx.x, nt = blockDim.x;
    for (int32_t i0 = tid; i0 < no; i0 += kBatch * nt)
    {
        uint32_t gv[kBatch];
        Real4<R> xn[kBatch], v[kBatch];
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            gv[e] = i0 + e * nt < no ? a.loc_vtx[l0 + i0 + e * nt] : 0u;
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            if (i0 + e * nt < no)
            {
                xn[e] = ld4(&s.prev[gv[e]]);
                v[e]  = ld4(&s.vel[gv[e]]);
            }
#pragma unroll
        for (int e = 0; e < kBatch; ++e)
            if (i0 + e * nt < no)
            {
                commit_vertex(sx[i0 + e * nt], xn[e], v[e], dt);
                st4(&s.vel[gv[e]], v[e]);
                st4(&s.prev[gv[e]], xn[e]);
            }
    }
}

// kExchange = false: the launch runs regions that share no vertex with any other (ensembles, a body in one
// region): no mailbox code at all.
template <typename R, bool kTrace, bool kDict, bool kExchange>
__device__ void run_region(ResidentArgs<R> const& a, int32_t region, Real4<R>* sx, DevChunk* s_chunks,
                           int32_t* s_xfirst, Real4<R>* s_dict)
{
    DeviceScene<R> const& s = a.s;
    int const tid = threadIdx.x, nt = blockDim.x;
    R const dt              = a.dt;
    int32_t const l0 = a.loc_off[region], nl = a.loc_off[region + 1] - l0, no = a.n_owned[region];
    int32_t const o0 = kExchange ? a.osv_off[region] : 0, n_osv = kExchange ? a.osv_off[region + 1] - o0 : 0;
    int32_t const s0 = a.surf_off[region], ns = a.surf_off[region + 1] - s0;
    // collision steps exist when detection is on — and, on a single GPU, only when it found something:
    // every CTA reads the same contact count, so they agree on the schedule (ranks of a decomposed
    // scene detect separately and could disagree: they always keep the steps)
    uint32_t const n_contacts =
        a.collide ? min(*s.contact_count, static_cast<uint32_t>(s.contact_cap)) : 0u;
    int32_t const C = a.n_colours, cs = (a.collide && (n_contacts > 0u || a.world > 1)) ? 1 : 0,
                  K = C > 0 ? a.iterations : 0;
    int32_t const per_iteration = C + cs;
    int32_t const n_phases      = 2 + K * per_iteration; // predict, K x ([collision] colours), commit

    // the region's chunk descriptors: [colour][part], read every step
    {
        constexpr int32_t W = static_cast<int32_t>(sizeof(DevChunk) / 4);
        int32_t const words = C * 2 * W;
        int32_t const* src  = reinterpret_cast<int32_t const*>(a.chunks);
        int32_t* dst        = reinterpret_cast<int32_t*>(s_chunks);
        for (int32_t w = tid; w < words; w += nt)
        {
            int32_t const c = w / (2 * W), rem = w % (2 * W);
            dst[w] = src[(static_cast<int64_t>(c) * a.n_regions + region) * 2 * W + rem];
        }
        if constexpr (kExchange)
            for (int32_t c = tid; c < C; c += nt)
                s_xfirst[c] = a.chunk_xfirst[static_cast<int64_t>(c) * a.n_regions + region];
        if constexpr (kDict)
            for (int32_t w = tid; w < kShapeWords * a.n_shapes; w += nt)
                s_dict[w] = a.shapes[w];
    }
    __syncthreads();
    int32_t traced = 0;
    // cluster i of a step runs on thread (i + rot) % nt (scene_build.h, item_rotation)
    int32_t const my_item = tid >= a.rot ? tid - a.rot : tid + nt - a.rot;
    // development aid: lane 0 of every warp records %clock64 stamps of its colour steps, 16 per (region, step, warp);
    // slot 3 holds %globaltimer at the start of the step (clocks of different SMs are not comparable)
    auto stamp = [&](int slot) { // slot < 0: record the value -slot in slot 1 instead of a clock stamp
        if (kTrace && (tid & 31) == 0 && a.trace && traced < a.trace_steps)
        {
            long long* row =
                &a.trace[((static_cast<int64_t>(region) * a.trace_steps + traced) * kTraceWarps + (tid >> 5)) * 16];
            if (slot < 0)
                row[1] = -slot;
            else
                row[slot] = clock_stamp();
            if (slot == 0)
            {
                long long g;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                row[3] = g;
            }
        }
    };
    // tag of the last colour step of iteration k that touches a vertex with this schedule
    auto last_colour_tag = [&](int32_t k, uint32_t lastc) -> uint32_t {
        return a.base + 1u + static_cast<uint32_t>(k * per_iteration + cs) + lastc;
    };
    R const at_c = s.collision_alpha / (dt * dt);
    R const inv_dt2 = R(1) / (dt * dt);
    int64_t const pull_stride = static_cast<int64_t>(a.entries / 4) * a.n_xclusters; // records per pull variant
    int64_t const push_stride = static_cast<int64_t>(a.entries / 2) * a.n_xclusters;

    // Phase p has tag base + p.  p = 0: predict; p = n_phases - 1: commit; in between iteration
    // k = (p - 1) / per_iteration, and q = (p - 1) % per_iteration is the collision step (q == 0 when
    // cs) or colour q - cs.  A colour phase with more clusters than threads takes several rounds.
    //
    // Every pass of the loop below ends at ONE place that prepares the cluster this thread runs
    // next (head loaded, shared vertices pulled) — before the barrier when that cluster belongs to the
    // next phase: a pulled vertex is touched by no cluster of this region in between (its previous touch
    // was by another region), and the records it waits for come from phases that do not wait for this thread.
    ClusterHead<R, kDict> head; // the prepared cluster of this thread ...
    int64_t cur_xq = -1;  // ... and its number among the exchange clusters (part 0), else -1
    int32_t item_i = -1;  // cluster index within its phase (part 0 first), -1: nothing prepared
    int32_t round  = 0;
    for (int32_t p = 0; p < n_phases;)
    {
        uint32_t const tag = a.base + static_cast<uint32_t>(p);
        int32_t const k    = p == 0 ? 0 : (p - 1) / per_iteration;
        int32_t const q    = p == 0 ? 0 : (p - 1) % per_iteration;
        bool const colour  = p > 0 && p < n_phases - 1 && !(cs && q == 0);
        int32_t const c    = q - cs;

        // ---- (1) what this thread runs after this pass: its loads (static data) go out first, so
        //          that they are in flight while the work of this pass runs
        bool advance = true;
        if (colour)
            advance = (round + 1) * nt >= s_chunks[2 * c].n[0] + s_chunks[2 * c + 1].n[0];
        int32_t const np      = advance ? p + 1 : p;
        int32_t const ni_next = advance ? my_item : (round + 1) * nt + my_item;
        bool has_next         = false;
        int32_t nk            = 0;
        ClusterHead<R, kDict> nhead;
        uint4 npull           = make_uint4(0u, 0u, 0u, 0u);
        int64_t next_xq       = -1;
        if (np > 0 && np < n_phases - 1)
        {
            nk               = (np - 1) / per_iteration;
            int32_t const nq = (np - 1) % per_iteration;
            if (!(cs && nq == 0))
            {
                int32_t const nc = nq - cs;
                int32_t const nA = s_chunks[2 * nc].n[0], nB = s_chunks[2 * nc + 1].n[0];
                if (ni_next < nA + nB)
                {
                    has_next           = true;
                    bool const in_x    = ni_next < nA;
                    DevChunk const& ch = s_chunks[2 * nc + (in_x ? 0 : 1)];
                    int32_t const ci   = in_x ? ni_next : ni_next - nA;
                    load_cluster_head<R, kDict>(nhead, a, ch, ci, nk == 0, s_dict);
                    if (kExchange && in_x)
                    {
                        next_xq = static_cast<int64_t>(s_xfirst[nc]) + ci;
                        npull   = __ldg(&a.pull[static_cast<int64_t>(2 * (nk > 0 ? 1 : 0) + cs) * pull_stride + next_xq]);
                    }
                }
            }
        }

        // ---- (2) the work of this pass
        if (p == 0)
        { // ---- predict (timestep.cpp:35-43): owned vertices; guests only need their inverse mass
            predict_owned<R, 2>(s.pos, s.prev, s.vel, a.loc_vtx + l0, dt, sx, nl, no);
            if constexpr (kExchange)
            {
                __syncthreads();
                // next touch of an owned shared vertex: the owner's collision step (surface vertex), else the
                // first cluster of the sweep that contains it — pushed when that cluster is another region's
                for (int32_t i = tid; i < n_osv; i += nt)
                {
                    uint32_t const meta = a.osv_meta[o0 + i];
                    bool const to_me    = K == 0 || (cs && (meta & kOsvSurface));
                    if (!to_me && (meta & kOsvFirstRemote))
                    {
                        Real4<R> const pp = sx[a.osv_slot[o0 + i]];
                        push<R>(a, a.osv_first[o0 + i], pp.x, pp.y, pp.z, tag);
                    }
                }
            }
        }
        else if (p == n_phases - 1)
        { // ---- commit (timestep.cpp:48-57) + surface copy
            if constexpr (kExchange)
            {
                for (int32_t i = tid; i < n_osv; i += nt)
                {
                    uint32_t const meta = a.osv_meta[o0 + i];
                    if (K > 0 && (meta & kOsvLastRemote))
                    {
                        Real4<R>* dst = &sx[a.osv_slot[o0 + i]];
                        R x, y, z;
                        xchg_wait<R>(a, a.n_entries + static_cast<uint32_t>(o0 + i), last_colour_tag(K - 1, meta & 0xffu), x, y, z);
                        dst->x = x;
                        dst->y = y;
                        dst->z = z;
                    }
                }
                __syncthreads();
            }
            commit_owned<R, 2>(s.prev, s.vel, a.loc_vtx + l0, dt, sx, no);
            // tetrahedral_body_t::update_visual_model (tetrahedral_body.cpp:157-165), owned surface vertices
            for (int32_t i = tid; i < ns; i += nt)
            {
                Real4<R> const pp = sx[a.surf_slot[s0 + i]];
                st4(&s.surf_pos[a.surf_index[s0 + i]], Real4<R>{pp.x, pp.y, pp.z, R(0)});
            }
        }
        else if (cs && q == 0)
        { // ---- collision constraints of the owned surface vertices (gauss_seidel_solver.cpp:28-31)
            for (int32_t i = tid; i < ns; i += nt)
            {
                uint32_t const slot  = a.surf_slot[s0 + i];
                uint32_t const first = n_contacts > 0 ? s.surf_first[a.surf_index[s0 + i]] : 0xffffffffu;
                uint32_t meta        = 0;
                uint32_t osv         = kRouteNone;
                if constexpr (kExchange)
                {
                    osv = a.surf_osv[s0 + i];
                    if (osv != kRouteNone)
                    { // shared: its last touch (last colour of the previous sweep) may have been another region's
                        meta = a.osv_meta[osv];
                        if (k > 0 && (meta & kOsvLastRemote))
                        {
                            R x, y, z;
                            xchg_wait<R>(a, a.n_entries + osv, last_colour_tag(k - 1, meta & 0xffu), x, y, z);
                            sx[slot].x = x;
                            sx[slot].y = y;
                            sx[slot].z = z;
                        }
                    }
                }
                if (first != 0xffffffffu)
                {
                    Real4<R> pp = sx[slot];
                    if (project_vertex_contacts(s, first, n_contacts, pp, at_c, k == 0))
                        sx[slot] = pp;
                }
                if constexpr (kExchange)
                    if (osv != kRouteNone && (meta & kOsvFirstRemote))
                    { // on to the first cluster of the sweep, contact or not: it waits for this step's tag
                        Real4<R> const pp = sx[slot];
                        push<R>(a, a.osv_first[osv], pp.x, pp.y, pp.z, tag);
                    }
            }
        }
        else
        { // ---- colour q - cs of iteration k (gauss_seidel_solver.cpp:32-35)
            int32_t const nA = s_chunks[2 * c].n[0];
            if (round == 0)
                stamp(0);
            if (item_i >= 0)
            { // cluster item_i of the phase runs on thread (item_i + rot) % nt; part 0 (clusters that exchange) first
                bool const in_x    = item_i < nA;
                DevChunk const& ch = s_chunks[2 * c + (in_x ? 0 : 1)];
                uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
                uint4 const* push_variant = nullptr;
                if (kExchange && cur_xq >= 0)
                { // where its shared vertices go afterwards (static): in flight while the tets run
                    push_variant = a.push + static_cast<int64_t>(2 * (k == K - 1 ? 1 : 0) + cs) * push_stride;
                    w0           = __ldg(&push_variant[cur_xq]);
                    w1           = __ldg(&push_variant[a.n_xclusters + cur_xq]);
                }
                run_cluster<R, kDict>(a, ch, in_x ? item_i : item_i - nA, head, sx, s_dict, k == 0, inv_dt2, stamp);
                // ---- (3) shared vertices whose next touch is another region's go there first: a neighbour's
                //          next step waits for them
                if (kExchange && cur_xq >= 0)
                    push_cluster<R>(a, push_variant, cur_xq, w0, w1, sx, tag);
                stamp(15);
            }
            stamp(6);
        }

        // ---- (4) the next cluster becomes the prepared one: its shared vertices are pulled now, before
        //          the barrier when it belongs to the next phase
        item_i = has_next ? ni_next : -1;
        cur_xq = -1;
        if (has_next)
        {
            head   = nhead;
            cur_xq = next_xq;
            if (kExchange && next_xq >= 0 && (npull.x & kPullValid))
                pull_cluster<R>(a, a.pull + static_cast<int64_t>(2 * (nk > 0 ? 1 : 0) + cs) * pull_stride, next_xq, npull,
                                sx, a.base + static_cast<uint32_t>(np), stamp);
        }
        if (advance)
        {
            __syncthreads();
            if (colour)
            {
                stamp(7);
                ++traced;
            }
            round = 0;
            p     = np;
        }
        else
            ++round;
    }
    __syncthreads(); // shared memory is reused by the next region of this CTA
}

template <typename R, bool kTrace, int kMaxThreads, int kMinBlocks, bool kDict, bool kExchange>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) k_substep_resident(ResidentArgs<R> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevChunk* s_chunks = reinterpret_cast<DevChunk*>(smem_raw);
    int32_t* s_xfirst  = reinterpret_cast<int32_t*>(smem_raw + chunk_area_bytes(a.n_colours));
    Real4<R>* s_dict   = reinterpret_cast<Real4<R>*>(smem_raw + chunk_area_bytes(a.n_colours) + xfirst_area_bytes(a.n_colours));
    Real4<R>* sx       = s_dict + (kDict ? kShapeWords * a.n_shapes : 0);
    // regions that exchange vertices come first in region_order (at most one per CTA: they must be
    // co-resident), the others follow and are handed out round-robin
    for (int32_t i = blockIdx.x; i < a.n_run; i += gridDim.x)
        run_region<R, kTrace, kDict, kExchange>(a, a.region_order[i], sx, s_chunks, s_xfirst, s_dict);
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
struct ResidentPlan
{
    ResidentArgs<R> args{};
    PBuf<int32_t> region_order, loc_off, n_owned, chunk_xfirst, osv_off, surf_off;
    PBuf<DevChunk> chunks;
    PBuf<uint32_t> loc_vtx, osv_slot, osv_meta, osv_first, surf_slot, surf_index, surf_osv, error;
    PBuf<uint2> tet_slots;
    PBuf<uint4> pull, push, box;
    PBuf<long long> trace;
    int64_t trace_len = 0;
    int grid = 0, block = 0;
    size_t smem = 0;
    uint32_t base = 1; // tags start at 1: a zero-initialised mailbox never matches an expected tag
    bool ready = false, cooperative = false;
    void const* kernel = nullptr;
    std::string why_not;

    static constexpr int64_t kSmemBudget = 224 * 1024;

    // ensembles: many independent bodies -> whole bodies per region
    static bool wants_region_per_body(HostScene const& h, int sm_count)
    {
        int64_t bodies = 0;
        for (auto const& b : h.bodies)
            bodies += (b.kind == BodyKind::tet && b.n_tets > 0);
        return bodies >= 2 * static_cast<int64_t>(sm_count);
    }
    static ResidentParams resident_params()
    {
        ResidentParams rp;
        rp.smem_bytes   = 208 * 1024; // vertex table (the rest: chunk descriptors, rest-shape dictionary)
        rp.vertex_bytes = static_cast<int32_t>(sizeof(Real4<R>));
        rp.max_threads  = 384; // 168 registers per thread; 512 threads (128 registers) spills in the tet loop
        return rp;
    }

    // the fewer threads a CTA has, the more registers each may use; regions that exchange nothing (ensembles)
    // run a leaner instantiation, several CTAs per SM
    template <bool kTrace>
    static void const* pick(int threads, bool dict, bool exchange, bool two_per_sm)
    {
#define SBS_K(T, B, D, X) reinterpret_cast<void const*>(k_substep_resident<R, kTrace, T, B, D, X>)
#ifndef SBS_ISLAND_MINBLOCKS
#define SBS_ISLAND_MINBLOCKS 2
#endif
        if (!exchange)
            return dict ? (threads <= 256 ? SBS_K(256, SBS_ISLAND_MINBLOCKS, true, false) : SBS_K(384, 1, true, false))
                        : (threads <= 256 ? SBS_K(256, SBS_ISLAND_MINBLOCKS, false, false) : SBS_K(384, 1, false, false));
        if (two_per_sm && threads <= 192) // more exchanging regions than SMs: two CTAs per SM must be co-resident
            return dict ? SBS_K(192, 2, true, true) : SBS_K(192, 2, false, true);
        return dict ? (threads <= 256 ? SBS_K(256, 1, true, true) : SBS_K(384, 1, true, true))
                    : (threads <= 256 ? SBS_K(256, 1, false, true) : SBS_K(384, 1, false, true));
#undef SBS_K
    }

    // rest-shape dictionary built by the engine (null / 0: none)
    uint8_t const* d_tet_shape = nullptr;
    Real4<R> const* d_shapes   = nullptr;
    int32_t n_shapes           = 0;

    // returns false (with why_not) when the scene does not fit this schedule
    // xp_in: the exchange plan of (cp, plan, world); its big arrays are moved to the device and released
    bool build(HostScene const& h, ClusterPlan const& cp, RegionPlan const& plan, ExchangePlan& xp_in,
               DeviceScene<R> const& d, cudaStream_t st, int sm_count, int rank = 0, int world = 1, int trace_n = 0)
    {
        int32_t const Rn = plan.n_regions;
        if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || Rn % world != 0)
        {
            why_not = "bad partition (world must divide the region count, at most 8 ranks)";
            return false;
        }
        int32_t const per_rank = Rn / world;
        int64_t const T        = h.n_tets();
        if (T >= (int64_t{1} << 31) / 2)
        {
            why_not = "too many tets for 32-bit chunk offsets";
            return false;
        }
        if (cp.n_regions != Rn || cp.nt <= 0)
        {
            why_not = "cluster plan and region plan disagree";
            return false;
        }
        if (!xp_in.why_not.empty() || xp_in.n_regions != Rn || xp_in.world != world ||
            static_cast<int64_t>(xp_in.tet_slots.size()) != 4 * T)
        {
            why_not = xp_in.why_not.empty() ? "no exchange plan for this partition" : xp_in.why_not;
            return false;
        }
        ExchangePlan& xp = xp_in;
        std::vector<uint2> slots(static_cast<size_t>(T));
        for (int64_t p = 0; p < T; ++p)
        {
            uint16_t const* q = &xp.tet_slots[4 * static_cast<size_t>(p)];
            slots[static_cast<size_t>(p)] = make_uint2(q[0] | (static_cast<uint32_t>(q[1]) << 16), q[2] | (static_cast<uint32_t>(q[3]) << 16));
        }
        std::vector<DevChunk> hchunks(cp.chunks.size());
        for (size_t i = 0; i < cp.chunks.size(); ++i)
        {
            hchunks[i].first  = cp.chunks[i].first;
            hchunks[i].cfirst = cp.chunks[i].cfirst;
            for (int m = 0; m < 8; ++m)
                hchunks[i].n[m] = cp.chunks[i].n[m];
        }
        auto const pack4 = [](std::vector<uint32_t> const& w) {
            std::vector<uint4> out(w.size() / 4);
            for (size_t i = 0; i < out.size(); ++i)
                out[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            return out;
        };
        // regions that share vertices must be co-resident; the others are handed out round-robin
        std::vector<int32_t> sync_r, island_r;
        for (int32_t r = rank * per_rank; r < (rank + 1) * per_rank; ++r)
            (plan.nbr_offsets[static_cast<size_t>(r) + 1] > plan.nbr_offsets[static_cast<size_t>(r)] ? sync_r : island_r)
                .push_back(r);
        bool const exchange   = xp.n_shared > 0;
        bool const two_per_sm = static_cast<int64_t>(sync_r.size()) > sm_count;

        // launch shape
        bool const dict = d_shapes != nullptr && n_shapes > 0 && n_shapes <= kMaxShapes;
        smem  = chunk_area_bytes(cp.n_colours) + xfirst_area_bytes(cp.n_colours) + (dict ? dict_area_bytes<R>(n_shapes) : 0) +
               static_cast<size_t>(std::max<int64_t>(xp.max_local, 1)) * sizeof(Real4<R>);
        block = cp.nt;
        if (static_cast<int64_t>(smem) > kSmemBudget)
        {
            why_not = "the vertices a region touches do not fit shared memory";
            return false;
        }
        if (trace_n > 0)
        { // development aid: clock stamps of the first trace_n colour steps of every launch (sbsb200_debug_read_trace)
            kernel = pick<true>(block, dict, exchange, two_per_sm);
            trace.upload(std::vector<long long>(static_cast<size_t>(Rn) * trace_n * kTraceWarps * 16, 0), st);
            args.trace       = trace.p;
            args.trace_steps = trace_n;
            trace_len        = static_cast<int64_t>(Rn) * trace_n * kTraceWarps * 16;
        }
        else
            kernel = pick<false>(block, dict, exchange, two_per_sm);
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
            cudaSuccess)
        {
            why_not = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
            return false;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1)
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
        cooperative = !sync_r.empty();

        std::vector<int32_t> order(sync_r);
        order.insert(order.end(), island_r.begin(), island_r.end());
        region_order.upload(order, st);
        tet_slots.upload(slots, st);
        chunks.upload(hchunks, st);
        loc_off.upload(xp.loc_off, st);
        n_owned.upload(xp.n_owned, st);
        loc_vtx.upload(xp.loc_vtx, st);
        chunk_xfirst.upload(xp.chunk_xfirst, st);
        pull.upload(pack4(xp.pull), st);
        push.upload(pack4(xp.push), st);
        osv_off.upload(xp.osv_off, st);
        osv_slot.upload(xp.osv_slot, st);
        osv_meta.upload(xp.osv_meta, st);
        osv_first.upload(xp.osv_first, st);
        surf_off.upload(xp.surf_off, st);
        surf_slot.upload(xp.surf_slot, st);
        surf_index.upload(xp.surf_index, st);
        surf_osv.upload(xp.surf_osv, st);
        size_t const n_boxes = static_cast<size_t>(xp.n_entries) + static_cast<size_t>(xp.n_shared) + 1;
        box.upload(std::vector<uint4>(n_boxes * Xchg<R>::kWords, make_uint4(0u, 0u, 0u, 0u)), st);
        box_bytes = n_boxes * Xchg<R>::kWords * sizeof(uint4);
        error.upload(std::vector<uint32_t>(1, 0u), st);
        base = 1;
        // the big host arrays are on the device now
        std::vector<uint32_t>().swap(xp.pull);
        std::vector<uint32_t>().swap(xp.push);
        std::vector<uint16_t>().swap(xp.tet_slots);

        args.s            = d;
        args.n_regions    = Rn;
        args.n_run        = static_cast<int32_t>(order.size());
        args.rank         = rank;
        args.world        = world;
        for (int r = 0; r < kMaxWorld; ++r)
            args.box_of_rank[r] = r == rank ? static_cast<void*>(box.p) : nullptr;
        args.n_colours    = cp.n_colours;
        args.rot          = cp.rot;
        args.region_order = region_order.p;
        args.tet_slots    = tet_slots.p;
        args.tet_shape    = dict ? d_tet_shape : nullptr;
        args.shapes       = dict ? d_shapes : nullptr;
        args.n_shapes     = dict ? n_shapes : 0;
        args.chunks       = chunks.p;
        args.loc_off      = loc_off.p;
        args.n_owned      = n_owned.p;
        args.loc_vtx      = loc_vtx.p;
        args.chunk_xfirst = chunk_xfirst.p;
        args.n_xclusters  = xp.n_xclusters;
        args.entries      = xp.entries;
        args.pull         = pull.p;
        args.push         = push.p;
        args.osv_off      = osv_off.p;
        args.osv_slot     = osv_slot.p;
        args.osv_meta     = osv_meta.p;
        args.osv_first    = osv_first.p;
        args.surf_off     = surf_off.p;
        args.surf_slot    = surf_slot.p;
        args.surf_index   = surf_index.p;
        args.surf_osv     = surf_osv.p;
        args.box          = box.p;
        args.n_entries    = xp.n_entries;
        args.error        = error.p;
        ready             = true;
        return true;
    }

    // tags consumed per substep launch
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
        cudaError_t const e =
            cooperative ? cudaLaunchCooperativeKernel(kernel, dim3(static_cast<unsigned>(grid)),
                                                      dim3(static_cast<unsigned>(block)), params, smem, st)
                        : cudaLaunchKernel(kernel, dim3(static_cast<unsigned>(grid)), dim3(static_cast<unsigned>(block)),
                                           params, smem, st);
        if (e != cudaSuccess)
            throw std::runtime_error(std::string("launch of the substep kernel: ") + cudaGetErrorString(e));
        base += steps_per_substep(iterations, collide);
        return 1;
    }

    // every rank's mailbox array must be mapped before the first step of a decomposed scene
    bool peers_connected() const
    {
        for (int r = 0; r < args.world; ++r)
            if (!args.box_of_rank[r])
                return false;
        return true;
    }
    size_t box_bytes = 0;

    // true when a launch ran out of its poll budget (call after synchronising the stream); the flag is cleared,
    // so that the next launch polls with a full budget again
    bool timed_out()
    {
        uint32_t e = 0;
        if (ready && error.p)
        {
            cudaMemcpy(&e, error.p, sizeof e, cudaMemcpyDeviceToHost);
            if (e != 0)
                cudaMemset(error.p, 0, sizeof e);
        }
        return e != 0;
    }
};

} // namespace sbsb200
