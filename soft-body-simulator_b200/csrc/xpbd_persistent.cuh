// Persistent, region-resident schedule: ONE cooperative kernel per substep.
//
// The tet set is cut into spatially compact regions, one per CTA (one CTA per SM for a large
// connected mesh).  A vertex whose tets all lie in one region is RESIDENT in it: its (xi, w)
// record lives in that CTA's shared memory for the whole substep (predict, every iteration and
// colour, commit) and never touches L2/HBM in between.  Vertices shared by several regions
// (and vertices that do not fit the shared memory) live in a global EXCHANGE array.
//
// Work unit = CLUSTER (<= 8 tets sharing vertices, <= 16 distinct vertices), one thread each.
// Every cluster carries a FETCH LIST: its vertices that are not resident.  A thread copies them
// from the exchange array into its private scratch slots of the shared vertex array, projects
// the cluster's tets one after the other entirely out of shared memory (the tet records hold
// 16-bit slots into that array, resident or scratch alike), and writes the scratch entries back.
// Per-tet constants are prefetched one tet ahead, a thread's first cluster of a colour step one
// step ahead.
//
// The Gauss-Seidel order is unchanged — colour major over the whole mesh.  There is no barrier
// between regions at all: synchronisation is per VERTEX and data-driven, through MAILBOXES.
// Every fetch-list entry (cluster, scratch slot) owns a mailbox; so does every non-resident
// vertex on behalf of its owner region (predict, collision constraints, commit).  A mailbox is one
// 16-byte word {x, y, z, tag} (fp32; three {value, tag} words in the fp64 build) read and written
// with single 128-bit accesses, so a reader sees position and tag together.  Which steps touch a
// vertex is static (the colours of the clusters that contain it, the collision steps if it is a
// surface vertex, predict, commit), so (a) whoever touches a vertex knows whose turn is next and
// PUSHES the new position into that mailbox, tagged with the current step, and (b) every reader
// knows the tag it has to wait for.  Every touch is a read-modify-write that waits for the
// previous touch, which orders read-after-write and write-after-read hazards alike without a
// fence (the record carries its own flag).  Mailboxes of the clusters of a step are laid out
// [slot][cluster]: the 32 lanes of a warp poll 32 adjacent records (4 cache lines per load
// instead of 32 — the scattered variant was bound by L1 wavefronts), and the pushes, which are
// fire-and-forget, carry the scatter.  A mailbox may just as well sit in another GPU's memory.
//
//   step 0                      : predict   (timestep.cpp:35-43)
//   step 1 + k(1+C) + 0         : collision constraints of iteration k (gauss_seidel_solver.cpp:28-31)
//   step 1 + k(1+C) + 1 + c     : colour c of iteration k               (:32-35)
//   step 1 + K(1+C)             : commit    (timestep.cpp:48-57) + surface copy
//
// Regions are co-resident (cooperative launch), so waiting on another region cannot deadlock:
// every wait is for a strictly earlier step.  A poll budget turns a lost update into an error
// code instead of a hang.
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

constexpr uint32_t kGlobalBit = 0x80000000u; // surface vertex address: global index when set, smem slot otherwise
constexpr uint32_t kNoVertex  = 0xffffffffu; // empty fetch-list entry
constexpr uint32_t kNoColour  = 0xffu;
constexpr uint32_t kNoBox     = 0xffffffffu; // "no such mailbox"
constexpr uint32_t kNoEntry   = 0xffffffffu; // cl_meta of an unused scratch slot
constexpr uint32_t kSurfaceBit = 0x80000000u; // cl_to_owner: the vertex is a surface vertex
// routing words: bits 0-27 mailbox index, bits 28-30 the rank (GPU) whose memory holds that mailbox
constexpr uint32_t kBoxIndexMask = kRouteIndexMask;
constexpr int kRankShift         = kRouteRankShift;
static_assert(kNoBox == kRouteNone && kSurfaceBit == kRouteSurfaceBit, "routing words: host and device agree");
constexpr int kMaxWorld          = 8;
constexpr int kPollBudget     = 1 << 24;     // polls of one record before the kernel gives up

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
struct PersistentArgs
{
    DeviceScene<R> s;
    int32_t n_regions, n_colours;
    int32_t n_run;                 // regions this launch runs (entries of region_order)
    int32_t nvc;                   // scratch entries per thread (4 * NVC4 of the instantiation)
    int32_t rot;                   // cluster i of a step runs on thread (i + rot) % nt (ClusterPlan::rot)
    int32_t banks;                 // banks of scratch slots (ClusterPlan::banks); 2 only in the kHandoff kernels:
                                   // the cluster of phase p keeps its fetched vertices in bank p % 2
    uint2 const* tet_slots_odd;    // kHandoff: tet_slots with the scratch slots moved to bank 1 (odd phases)
    int64_t n_clusters;
    int32_t const* region_order;   // CTA b runs region_order[b], [b + grid], ...; regions that share
                                   // vertices come first, one per CTA
    uint2 const* tet_slots;        // per tet (storage order): 4 x u16 slots into the shared vertex array
    // rest-shape dictionary (kDict kernels): meshes with few distinct rest shapes — every lattice has ten —
    // keep the (DmInv, V0, material) records in shared memory and stream one byte per tet instead of 48
    uint8_t const* tet_shape;      // per tet: index into shapes
    Real4<R> const* shapes;        // [3 * n_shapes]: r0, r1, r2 of every distinct record
    int32_t n_shapes;
    // per fetch-list entry, [nvc/4][n_clusters] each (entry j of cluster q: component j%4 of [j/4][q]);
    // the mailbox of that entry is box[j * n_clusters + q]
    uint4 const* cl_meta;          // touch schedule (ClusterPlan::cl_meta), 0xffffffff = no entry
    Real4<R> const* cl_w;          // inverse mass (patched by set_mass)
    uint4 const* cl_to;            // where the vertex goes next: the mailbox of the next entry touching it in
                                   // the sweep, or of the first one (next sweep) when this is the last
    uint4 const* cl_to_owner;      // the same, but the owner mailbox instead of "the first one": used when
                                   // the owner is next (commit; collision step if kSurfaceBit is set).
                                   // kNoBox = unused scratch slot
    DevChunk const* chunks;        // [(colour * n_regions + region) * 2 + part]
    int32_t const* vtx_off;        // [n_regions + 1] resident vertex list
    uint32_t const* vtx;           // global vertex id of resident slot i
    int32_t const* ifv_off;        // [n_regions + 1] owned non-resident vertices
    uint32_t const* ifv;
    uint32_t const* ifv_meta;      // their touch schedule (ClusterPlan::vertex_meta)
    uint32_t const* ifv_first;     // mailbox of the first entry touching them in a sweep (kNoBox: none);
                                   // the owner mailbox of owned vertex i is box[n_entries + i]
    int32_t const* surf_off;       // [n_regions + 1] owned surface vertices
    uint32_t const* surf_index;    // surface vertex index (into s.surf_pos / s.surf_first)
    uint32_t const* surf_addr;     // shared-memory slot, or (index into ifv) | kGlobalBit when not resident
    void* box;                     // mailboxes: n_entries of the fetch lists, then one per owned vertex
    uint32_t n_entries;            // nvc * n_clusters
    // decomposition over GPUs: every rank plans the same regions and runs its own block of them; the
    // mailbox of an entry lives on the rank that reads it, pushes to other ranks are peer stores
    int32_t rank, world;
    void* box_of_rank[kMaxWorld];  // mailbox arrays of all ranks (peer-mapped); [rank] == box
    uint32_t* error;               // set to 1 when a poll budget ran out
    uint32_t base;                 // tag of step 0 of this launch
    long long* trace;              // development aid: per-step clock stamps of thread 0 (nullptr = off)
    int32_t trace_steps;           // colour steps recorded per region
    int32_t iterations;
    int32_t collide;
    R dt;
};

// Push a position into the mailbox a routing word names (index + rank), tagged with the step.
template <typename R>
__device__ __forceinline__ void push(PersistentArgs<R> const& a, uint32_t route, R x, R y, R z, uint32_t tag)
{
    uint32_t const index = route & kBoxIndexMask;
    if (a.world > 1)
    {
        uint32_t const r = (route >> kRankShift) & 7u;
        Xchg<R>::store(a.box_of_rank[r], index, x, y, z, tag, true);
    }
    else
        Xchg<R>::store(a.box, index, x, y, z, tag);
}

// Wait for `expect` on (local) mailbox `b` and return the position in it.
template <typename R>
__device__ __forceinline__ void xchg_wait(PersistentArgs<R> const& a, uint32_t b, uint32_t expect, R& x, R& y, R& z)
{
    int polls = 0;
    while (!Xchg<R>::load(a.box, b, expect, x, y, z, a.world > 1))
    {
        if (poll_expired(a.error, ++polls))
            break;
        __nanosleep(20);
    }
}

template <typename R>
__device__ __forceinline__ Real4<R> load_vertex(uint32_t addr, Real4<R> const* sx, Real4<R> const* pos)
{
    if (addr & kGlobalBit)
    { // written by another thread of this CTA just before a barrier: read it from L2
        uint32_t const gv = addr & ~kGlobalBit;
        return Real4<R>{__ldcg(&pos[gv].x), __ldcg(&pos[gv].y), __ldcg(&pos[gv].z), R(0)};
    }
    return sx[addr];
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
__device__ __forceinline__ TetRecord<R, kDict> load_tet(PersistentArgs<R> const& a, uint2 const* slots, int32_t t,
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

// What a thread keeps of a cluster between the moment it is prepared (head loaded, vertices fetched,
// normally one step ahead) and the moment it runs.
template <typename R, bool kDict>
struct ClusterHead
{
    TetRecord<R, kDict> tet0;
    R mu, lam, at; // material of the cluster's body; at = alpha / dt^2
};

template <typename R, bool kDict>
__device__ __forceinline__ void load_cluster_head(ClusterHead<R, kDict>& h, PersistentArgs<R> const& a,
                                                  uint2 const* slots, DevChunk const& ch, int32_t i,
                                                  int first_iteration, Real4<R> const* s_dict)
{
    h.tet0 = load_tet<R, kDict>(a, slots, ch.first + i, first_iteration);
    R mat_id;
    if constexpr (kDict)
        mat_id = s_dict[3 * h.tet0.shape + 2].z; // waits for the shape id: one L2 round trip, a step ahead
    else
        mat_id = h.tet0.r2.z;
    Real4<R> const mat = ld4_ro(&a.s.materials[mat_index(mat_id)]);
    h.mu               = mat.x;
    h.lam              = mat.y;
    h.at               = mat.z / (a.dt * a.dt);
}

// Where a colour step sits in the substep: what the tag arithmetic needs.
struct StepInfo
{
    uint32_t step;  // tag of this step
    uint32_t base;  // tag of the predict step
    uint32_t shift; // which byte of a fetch entry's schedule word applies: 8 * (2 * (k > 0) + cs)
    bool to_owner;  // after the last entry of a sweep the owner is next (collision step or commit) ...
    bool surface_to_owner; // ... for surface vertices only (a collision step follows, not the commit)
};

// tag of the previous touch of a fetched vertex (ClusterPlan::cl_meta)
__device__ __forceinline__ uint32_t expected_tag(StepInfo const& si, uint32_t meta)
{
    uint32_t const d = (meta >> si.shift) & 0xffu;
    return d == 0xffu ? si.base : si.step - d;
}

// Fetch of one cluster: every entry of its fetch list waits for the tag of its previous touch and
// lands in the thread's scratch slots; all polls of a round are in flight together.
// kHandoff: entries marked kMetaLocal were written into the scratch slot by the previous touch (same region,
// previous step) and are not polled.
template <typename R, int NVC4, bool kHandoff, typename Stamp>
__device__ __forceinline__ void gather_cluster(PersistentArgs<R> const& a, int64_t q, uint4 const (&fmeta)[NVC4],
                                               Real4<R> const (&fw)[NVC4], Real4<R>* sx, StepInfo const& si,
                                               Stamp&& stamp)
{
    int const tid = threadIdx.x, nt = blockDim.x;
    uint32_t meta[4 * NVC4];
    R w[4 * NVC4];
    uint32_t pending = 0;
#pragma unroll
    for (int k = 0; k < NVC4; ++k)
    {
        meta[4 * k + 0] = fmeta[k].x;
        meta[4 * k + 1] = fmeta[k].y;
        meta[4 * k + 2] = fmeta[k].z;
        meta[4 * k + 3] = fmeta[k].w;
        w[4 * k + 0]    = fw[k].x;
        w[4 * k + 1]    = fw[k].y;
        w[4 * k + 2]    = fw[k].z;
        w[4 * k + 3]    = fw[k].w;
    }
#pragma unroll
    for (int j = 0; j < 4 * NVC4; ++j)
        if (meta[j] != kNoEntry && !(kHandoff && meta[j] == kMetaLocal))
            pending |= 1u << j;
    uint32_t const mine = static_cast<uint32_t>(q); // mailbox of entry j: j * n_clusters + q
    int polls = 0;
    stamp(4);
    while (pending)
    {
        // a round: every outstanding record is requested, then every answer is examined
        constexpr int kBatch = sizeof(R) == 4 ? 4 * NVC4 : 4;
#pragma unroll
        for (int h = 0; h < 4 * NVC4; h += kBatch)
        {
            typename Xchg<R>::Raw raw[kBatch];
#pragma unroll
            for (int e = 0; e < kBatch; ++e)
                if (pending >> (h + e) & 1u)
                    raw[e] = Xchg<R>::fetch(a.box, static_cast<uint32_t>(h + e) * static_cast<uint32_t>(a.n_clusters) + mine,
                                            a.world > 1);
#pragma unroll
            for (int e = 0; e < kBatch; ++e)
                if (pending >> (h + e) & 1u)
                {
                    R x, y, z;
                    if (Xchg<R>::decode(raw[e], expected_tag(si, meta[h + e]), x, y, z))
                    {
                        sx[(h + e) * nt + tid] = Real4<R>{x, y, z, w[h + e]};
                        pending &= ~(1u << (h + e));
                    }
                }
        }
        if (polls == 0)
            stamp(2);
        if (pending && poll_expired(a.error, ++polls))
            break;
    }
    stamp(-(polls + 1)); // slot 1 <- number of poll rounds
}

// One cluster, fetched already: project its tets in order out of shared memory -> write back.
// q = storage index of the cluster when it has a fetch list (part A), -1 otherwise.
// The fetched vertices are NOT written back here: push_cluster does that with the routing words
// loaded here (`to`, `to_owner`).
template <typename R, int NVC4, bool kDict, typename Stamp>
__device__ __forceinline__ void run_cluster(PersistentArgs<R> const& a, DevChunk const& ch, int32_t i, int64_t q,
                                            ClusterHead<R, kDict> const& head, Real4<R>* sx,
                                            Real4<R> const* s_dict, int first_iteration, uint2 const* slots,
                                            uint4 (&to)[NVC4], uint4 (&to_owner)[NVC4], Stamp&& stamp)
{
    DeviceScene<R> const& s = a.s;
    // where the fetched vertices go afterwards (static routing data): in flight while the tets run
#pragma unroll
    for (int k = 0; k < NVC4; ++k)
    {
        int64_t const at = static_cast<int64_t>(k) * a.n_clusters + (q >= 0 ? q : 0);
        to[k]            = q >= 0 ? __ldg(&a.cl_to[at]) : make_uint4(kNoBox, kNoBox, kNoBox, kNoBox);
        to_owner[k]      = q >= 0 ? __ldg(&a.cl_to_owner[at]) : make_uint4(kNoBox, kNoBox, kNoBox, kNoBox);
    }
    // column layout: tet m of cluster i sits at first + n[0] + .. + n[m-1] + i
    int32_t n0 = ch.n[0], n1 = ch.n[1], n2 = ch.n[2], n3 = ch.n[3], n4 = ch.n[4], n5 = ch.n[5], n6 = ch.n[6],
            n7 = ch.n[7];
    int32_t t        = ch.first + i;
    TetRecord<R, kDict> cur = head.tet0;
    int tslot        = 8;
    stamp(tslot++);
#pragma unroll 1
    for (;;)
    {
        bool const more = i < n1;
        TetRecord<R, kDict> nxt;
        if (more)
            nxt = load_tet<R, kDict>(a, slots, t + n0, first_iteration);
        uint32_t const a1 = cur.slots.x & 0xffffu, a2 = cur.slots.x >> 16, a3 = cur.slots.y & 0xffffu,
                       a4 = cur.slots.y >> 16;
        Real4<R> p1 = sx[a1], p2 = sx[a2], p3 = sx[a3], p4 = sx[a4];
        R lambda = cur.lambda;
        Vec3<R> const z{};
        Real4<R> r0, r1, r2;
        if constexpr (kDict)
        {
            r0 = s_dict[3 * cur.shape];
            r1 = s_dict[3 * cur.shape + 1];
            r2 = s_dict[3 * cur.shape + 2];
        }
        else
        {
            r0 = cur.r0;
            r1 = cur.r1;
            r2 = cur.r2;
        }
        green_project_at<R, false>(p1, p2, p3, p4, z, z, z, z, r0, r1, r2, head.mu, head.lam, head.at, R(0), a.dt,
                                   lambda);
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

// Push every fetched vertex of the cluster that just ran to whoever touches it next, always: the tag is
// what that touch waits for (select instead of branch per entry; the single-GPU and the peer-memory
// variants are unswitched).  Reads the thread's scratch slots: call it before they are refilled.
// kHandoff: sx is this step's bank of scratch slots, sx_next the next step's, where a hand-off inside the region
// deposits the vertex (routing word with kRouteLocalBit: the scratch slot of the cluster that runs next).
template <typename R, int NVC4, bool kHandoff, typename Stamp>
__device__ __forceinline__ void push_cluster(PersistentArgs<R> const& a, uint4 const (&to)[NVC4],
                                             uint4 const (&to_owner)[NVC4], Real4<R> const* sx, Real4<R>* sx_next,
                                             StepInfo const& si, Stamp&& stamp)
{
    int const tid = threadIdx.x, nt = blockDim.x;
    {
        bool const multi = a.world > 1;
#pragma unroll
        for (int h = 0; h < NVC4; ++h)
        {
            uint32_t const ta[4] = {to[h].x, to[h].y, to[h].z, to[h].w};
            uint32_t const tb[4] = {to_owner[h].x, to_owner[h].y, to_owner[h].z, to_owner[h].w};
            Real4<R> p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                p[e] = sx[(4 * h + e) * nt + tid];
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
                bool const owner_next = si.to_owner | (si.surface_to_owner & ((tb[e] & kSurfaceBit) != 0u));
                uint32_t const route  = owner_next ? tb[e] : ta[e];
                uint32_t const index  = route & kBoxIndexMask;
                bool const local = kHandoff && tb[e] != kNoBox && (route & kRouteLocalBit) != 0u;
                if (kHandoff && local)
                    sx_next[route & kRouteLocalIndexMask] = p[e];
                if (tb[e] != kNoBox && !local)
                {
                    if (!multi)
                        Xchg<R>::store(a.box, index, p[e].x, p[e].y, p[e].z, si.step);
                    else
                        Xchg<R>::store(a.box_of_rank[(route >> kRankShift) & 7u], index, p[e].x, p[e].y, p[e].z,
                                       si.step, true);
                }
            }
        }
    }
    stamp(15);
}

__device__ __forceinline__ long long clock_stamp()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

template <typename R, int NVC4, bool kTrace, bool kDict, bool kHandoff>
__device__ void run_region(PersistentArgs<R> const& a, int32_t region, Real4<R>* sx, DevChunk* s_chunks,
                           Real4<R>* s_dict)
{
    DeviceScene<R> const& s = a.s;
    int const tid = threadIdx.x, nt = blockDim.x;
    R const dt              = a.dt;
    Real4<R>* const sres    = sx + (kHandoff ? a.banks : 1) * a.nvc * nt; // resident vertices follow the scratch slots
    // kHandoff with two banks: the cluster of phase p keeps its fetched vertices in bank p % 2
    auto bank_of  = [&](int32_t phase) -> int32_t { return kHandoff && a.banks > 1 && (phase & 1) ? a.nvc * nt : 0; };
    auto slots_of = [&](int32_t phase) -> uint2 const* {
        return kHandoff && (phase & 1) ? a.tet_slots_odd : a.tet_slots;
    };
    int32_t const v0 = a.vtx_off[region], nv = a.vtx_off[region + 1] - v0;
    int32_t const i0 = a.ifv_off[region], ni = a.ifv_off[region + 1] - i0;
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
        if constexpr (kDict)
            for (int32_t w = tid; w < 3 * a.n_shapes; w += nt)
                s_dict[w] = a.shapes[w];
    }
    __syncthreads();
    int32_t traced = 0;
    // cluster i of a step runs on thread (i + rot) % nt (scene_build.h, item_rotation)
    int32_t const my_item = tid >= a.rot ? tid - a.rot : tid + nt - a.rot;
    auto stamp = [&](int slot) { // slot < 0: record the value -slot in slot 1 instead of a clock stamp
        if (kTrace && my_item == 0 && a.trace && traced < a.trace_steps)
        {
            long long* row = &a.trace[(static_cast<int64_t>(region) * a.trace_steps + traced) * 16];
            if (slot < 0)
                row[1] = -slot;
            else
                row[slot] = clock_stamp();
        }
    };
    // tag of the last colour step of iteration k that touches a vertex with this schedule
    auto last_colour_tag = [&](int32_t k, uint32_t lastc) -> uint32_t {
        return a.base + 1u + static_cast<uint32_t>(k * per_iteration + cs) + lastc;
    };
    R const at_c = s.collision_alpha / (dt * dt);

    // Phase p has tag base + p.  p = 0: predict; p = n_phases - 1: commit; in between iteration
    // k = (p - 1) / per_iteration, and q = (p - 1) % per_iteration is the collision step (q == 0 when
    // cs) or colour q - cs.  A colour phase with more clusters than threads takes several rounds.
    //
    // Every pass of the loop below ends at ONE place that prepares the cluster this thread runs
    // next (head loaded, vertices fetched) — before the barrier when that cluster belongs to the
    // next phase: the scratch slots are free once the thread's own cluster is written back, and
    // the records it waits for come from clusters of phases that do not wait for this thread.
    ClusterHead<R, kDict> head; // the prepared cluster of this thread ...
    int64_t cur_q = -1;   // ... and its storage index when it has a fetch list (part A), else -1
    int32_t item_i = -1; // cluster index within its phase (part A first), -1: nothing prepared
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
        bool has_next = false, next_in_a = false;
        int32_t nk = 0;
        ClusterHead<R, kDict> nhead;
        uint4 nmeta[NVC4];
        Real4<R> nw[NVC4];
        int64_t next_q = -1;
#pragma unroll
        for (int j = 0; j < NVC4; ++j)
        {
            nmeta[j] = make_uint4(kNoEntry, kNoEntry, kNoEntry, kNoEntry);
            nw[j]    = Real4<R>{R(0), R(0), R(0), R(0)};
        }
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
                    next_in_a          = ni_next < nA;
                    DevChunk const& ch = s_chunks[2 * nc + (next_in_a ? 0 : 1)];
                    int32_t const ci   = next_in_a ? ni_next : ni_next - nA;
                    load_cluster_head<R, kDict>(nhead, a, slots_of(np), ch, ci, nk == 0, s_dict);
                    if (next_in_a)
                    {
                        next_q = static_cast<int64_t>(ch.cfirst) + ci;
#pragma unroll
                        for (int j = 0; j < NVC4; ++j)
                        {
                            nmeta[j] = __ldg(&a.cl_meta[static_cast<int64_t>(j) * a.n_clusters + next_q]);
                            nw[j]    = ld4_ro(&a.cl_w[static_cast<int64_t>(j) * a.n_clusters + next_q]);
                        }
                    }
                }
            }
        }

        // ---- (2) the work of this pass
        uint4 to[NVC4], to_owner[NVC4]; // routing words of the cluster this pass runs (push_cluster)
        bool pushing = false;
        StepInfo const si{tag, a.base, static_cast<uint32_t>(8 * (2 * (k > 0 ? 1 : 0) + cs)), k == K - 1, cs != 0};
        if (p == 0)
        { // ---- predict (timestep.cpp:35-43)
            for (int32_t i = tid; i < nv; i += nt)
            {
                uint32_t const gv = a.vtx[v0 + i];
                Real4<R> pp       = ld4(&s.pos[gv]);
                Real4<R> const x  = ld4(&s.prev[gv]);
                Real4<R> v        = ld4(&s.vel[gv]);
                predict_vertex(pp, x, v, dt);
                sres[i] = pp;
            }
            for (int32_t i = tid; i < ni; i += nt)
            {
                uint32_t const gv = a.ifv[i0 + i];
                Real4<R> pp       = ld4(&s.pos[gv]);
                Real4<R> const x  = ld4(&s.prev[gv]);
                Real4<R> v        = ld4(&s.vel[gv]);
                predict_vertex(pp, x, v, dt);
                // next touch: the owner's collision step (surface vertex), else the first cluster of the
                // sweep that contains the vertex, else (no sweep touches it) the owner's commit
                uint32_t const first = a.ifv_first[i0 + i];
                bool const to_me     = K == 0 || first == kNoBox || (cs && (a.ifv_meta[i0 + i] & 0x100u));
                if (to_me)
                    Xchg<R>::store(a.box, a.n_entries + static_cast<uint32_t>(i0 + i), pp.x, pp.y, pp.z, tag, a.world > 1);
                else
                    push<R>(a, first, pp.x, pp.y, pp.z, tag);
            }
        }
        else if (p == n_phases - 1)
        { // ---- commit (timestep.cpp:48-57) + surface copy
            for (int32_t i = tid; i < nv; i += nt)
            {
                uint32_t const gv = a.vtx[v0 + i];
                Real4<R> const pp = sres[i];
                Real4<R> xn       = ld4(&s.prev[gv]);
                Real4<R> v        = ld4(&s.vel[gv]);
                commit_vertex(pp, xn, v, dt);
                st4(&s.vel[gv], v);
                st4(&s.prev[gv], xn);
            }
            for (int32_t i = tid; i < ni; i += nt)
            {
                uint32_t const gv    = a.ifv[i0 + i];
                uint32_t const meta  = a.ifv_meta[i0 + i];
                uint32_t const lastc = meta & 0xffu;
                uint32_t want        = a.base;
                if (K > 0)
                {
                    if (lastc != kNoColour)
                        want = last_colour_tag(K - 1, lastc);
                    else if (cs && (meta & 0x100u))
                        want = a.base + 1u + static_cast<uint32_t>((K - 1) * per_iteration);
                }
                Real4<R> pp;
                xchg_wait<R>(a, a.n_entries + static_cast<uint32_t>(i0 + i), want, pp.x, pp.y, pp.z);
                Real4<R> xn = ld4(&s.prev[gv]);
                Real4<R> v  = ld4(&s.vel[gv]);
                commit_vertex(pp, xn, v, dt);
                st4(&s.vel[gv], v);
                st4(&s.prev[gv], xn);
            }
            __syncthreads(); // prev[] of the owned vertices is final: the surface copy below reads it
            // tetrahedral_body_t::update_visual_model (tetrahedral_body.cpp:157-165), owned surface vertices
            for (int32_t i = tid; i < ns; i += nt)
            {
                uint32_t addr = a.surf_addr[s0 + i];
                if (addr & kGlobalBit)
                    addr = a.ifv[addr & ~kGlobalBit] | kGlobalBit;
                Real4<R> const pp = load_vertex(addr, sx, s.prev);
                st4(&s.surf_pos[a.surf_index[s0 + i]], Real4<R>{pp.x, pp.y, pp.z, R(0)});
            }
        }
        else if (cs && q == 0)
        { // ---- collision constraints of the owned surface vertices (gauss_seidel_solver.cpp:28-31);
          //      non-resident ones are re-tagged whether or not they have a contact
            for (int32_t i = tid; i < ns; i += nt)
            {
                uint32_t const addr  = a.surf_addr[s0 + i];
                uint32_t const first = n_contacts > 0 ? s.surf_first[a.surf_index[s0 + i]] : 0xffffffffu;
                if (addr & kGlobalBit)
                { // owned, not resident: out of the owner mailbox, on to the first cluster of the sweep
                    uint32_t const iv    = addr & ~kGlobalBit;
                    uint32_t const gv    = a.ifv[iv];
                    uint32_t const lastc = a.ifv_meta[iv] & 0xffu;
                    uint32_t const want  = (k > 0 && lastc != kNoColour) ? last_colour_tag(k - 1, lastc)
                                           : k > 0                       ? tag - static_cast<uint32_t>(per_iteration)
                                                                         : a.base;
                    Real4<R> pp;
                    xchg_wait<R>(a, a.n_entries + iv, want, pp.x, pp.y, pp.z);
                    if (first != 0xffffffffu)
                    {
                        pp.w = s.pos[gv].w;
                        project_vertex_contacts(s, first, n_contacts, pp, at_c, k == 0);
                    }
                    uint32_t const to = a.ifv_first[iv];
                    if (to != kNoBox)
                        push<R>(a, to, pp.x, pp.y, pp.z, tag);
                    else
                        Xchg<R>::store(a.box, a.n_entries + iv, pp.x, pp.y, pp.z, tag, a.world > 1);
                }
                else if (first != 0xffffffffu)
                {
                    Real4<R> pp = sx[addr];
                    if (project_vertex_contacts(s, first, n_contacts, pp, at_c, k == 0))
                        sx[addr] = pp;
                }
            }
        }
        else
        { // ---- colour q - cs of iteration k (gauss_seidel_solver.cpp:32-35)
            int32_t const nA = s_chunks[2 * c].n[0];
            if (round == 0)
                stamp(0);
            if (item_i >= 0)
            { // cluster item_i of the phase runs on thread (item_i + rot) % nt; part A (clusters that fetch) first
                bool const in_a    = item_i < nA;
                DevChunk const& ch = s_chunks[2 * c + (in_a ? 0 : 1)];
                run_cluster<R, NVC4, kDict>(a, ch, in_a ? item_i : item_i - nA, cur_q, head, sx, s_dict, k == 0,
                                            slots_of(p), to, to_owner, stamp);
                pushing = cur_q >= 0;
            }
            stamp(6);
        }

        // ---- (3) the fetched vertices of the cluster that ran go to whoever touches them next — first,
        //          because a neighbour's next step waits for them (issuing this thread's own polls ahead
        //          of the pushes hid their latency but cost 8 % of the frame: profiles/r01_summary.md) —
        //          and the next cluster becomes the prepared one: its shared vertices are fetched now,
        //          before the barrier when it belongs to the next phase
        if (pushing)
            push_cluster<R, NVC4, kHandoff>(a, to, to_owner, sx + bank_of(p), sx + bank_of(p + 1), si, stamp);
        item_i = has_next ? ni_next : -1;
        if (has_next)
        {
            head  = nhead;
            cur_q = next_q;
            if (next_in_a)
                gather_cluster<R, NVC4, kHandoff>(a, next_q, nmeta, nw, sx + bank_of(np),
                                        StepInfo{a.base + static_cast<uint32_t>(np), a.base,
                                                 static_cast<uint32_t>(8 * (2 * (nk > 0 ? 1 : 0) + cs)), false, false},
                                        stamp);
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

// dynamic shared memory: [chunk descriptors of the region | rest-shape dictionary | scratch slots | resident vertices]
__host__ __device__ inline size_t chunk_area_bytes(int n_colours)
{
    return (static_cast<size_t>(n_colours) * 2 * sizeof(DevChunk) + 31) / 32 * 32;
}
template <typename R>
__host__ __device__ inline size_t dict_area_bytes(int n_shapes)
{
    return static_cast<size_t>(3 * n_shapes) * sizeof(Real4<R>);
}

template <typename R, int NVC4, bool kTrace, int kMaxThreads, bool kDict, bool kHandoff = false>
__global__ void __launch_bounds__(kMaxThreads) k_substep_persistent(PersistentArgs<R> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevChunk* s_chunks = reinterpret_cast<DevChunk*>(smem_raw);
    Real4<R>* s_dict   = reinterpret_cast<Real4<R>*>(smem_raw + chunk_area_bytes(a.n_colours));
    Real4<R>* sx       = s_dict + (kDict ? 3 * a.n_shapes : 0);
    // regions that share vertices come first in region_order (at most one per CTA: they must be
    // co-resident), the others follow and are handed out round-robin
    for (int32_t i = blockIdx.x; i < a.n_run; i += gridDim.x)
        run_region<R, NVC4, kTrace, kDict, kHandoff>(a, a.region_order[i], sx, s_chunks, s_dict);
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
    PBuf<int32_t> region_order, vtx_off, ifv_off, surf_off;
    PBuf<DevChunk> chunks;
    PBuf<uint32_t> vtx, ifv, ifv_meta, ifv_first, surf_index, surf_addr, error;
    PBuf<uint2> tet_slots, tet_slots_odd;
    PBuf<uint4> cl_meta, cl_to, cl_to_owner, box;
    PBuf<Real4<R>> cl_w;
    std::vector<uint4> h_fetch; // host copy of cl_fetch, to patch cl_w when a mass changes
    PBuf<long long> trace;
    int64_t trace_len = 0;
    int grid = 0, block = 0;
    size_t smem = 0;
    uint32_t base = 1; // tags start at 1: the zero-initialised exchange array never matches an expected tag
    bool ready = false;
    void const* kernel = nullptr;
    std::string why_not;

    static constexpr int64_t kSmemBudget   = 224 * 1024;
    static constexpr int64_t kVertexBudget = 190 * 1024; // scratch + resident vertices (rest: chunk descriptors)

    static int32_t regions_for(int sm_count, int64_t n_tets, int world = 1)
    {
        return sbsb200::regions_for(sm_count, n_tets, world);
    }
    // ensembles: many independent bodies -> one region per body
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
        rp.smem_bytes   = kVertexBudget;
        rp.vertex_bytes = static_cast<int32_t>(sizeof(Real4<R>));
        rp.max_threads  = 384; // 168 registers per thread; 512 threads (128 registers) spills in the tet loop
        return rp;
    }

    // the fewer threads a CTA has, the more registers each may use (255 / 168)
    template <int NVC4, bool kTrace>
    static void const* pick(int threads, bool dict, bool handoff = false)
    {
        if constexpr (sizeof(R) == 4 && !kTrace)
            if (handoff) // experimental kernels (hand-off inside a region): fp32 only
            {
                if (dict)
                    return threads <= 256
                               ? reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, false, 256, true, true>)
                               : reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, false, 384, true, true>);
                return threads <= 256
                           ? reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, false, 256, false, true>)
                           : reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, false, 384, false, true>);
            }
        if (dict)
            return threads <= 256 ? reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, kTrace, 256, true>)
                                  : reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, kTrace, 384, true>);
        return threads <= 256 ? reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, kTrace, 256, false>)
                              : reinterpret_cast<void const*>(k_substep_persistent<R, NVC4, kTrace, 384, false>);
    }

    // returns false (with why_not) when the scene does not fit this schedule
    // rest-shape dictionary built by the engine (null / 0: none)
    uint8_t const* d_tet_shape  = nullptr;
    Real4<R> const* d_shapes    = nullptr;
    int32_t n_shapes            = 0;

    bool build(HostScene const& h, ClusterPlan const& cp, RegionPlan const& plan, DeviceScene<R> const& d,
               cudaStream_t st, int sm_count, int rank = 0, int world = 1)
    {
        int32_t const Rn = plan.n_regions;
        if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || Rn % world != 0)
        {
            why_not = "bad partition (world must divide the region count, at most 8 ranks)";
            return false;
        }
        int32_t const per_rank = Rn / world;
        int64_t const T = h.n_tets(), V = h.n_vertices();
        if (cp.banks > 1 && sizeof(R) != 4)
        {
            why_not = "the hand-off kernels exist in the fp32 build only";
            return false;
        }
        if (!cp.why_not.empty())
        {
            why_not = cp.why_not;
            return false;
        }
        if (T >= (int64_t{1} << 31) / 2)
        {
            why_not = "too many tets for 32-bit chunk offsets";
            return false;
        }
        if (cp.n_regions != Rn || cp.nt <= 0 || static_cast<int64_t>(cp.tet_slots.size()) != 4 * T)
        {
            why_not = "cluster plan and region plan disagree";
            return false;
        }
        int const nvc = cp.nvc <= 8 ? 8 : 16; // instantiations: 8 or 16 scratch entries per thread
        int64_t const scratch = static_cast<int64_t>(cp.banks) * cp.nvc * cp.nt; // where the plan's resident slots start
        std::vector<uint2> slots(static_cast<size_t>(T)), slots_odd(cp.banks > 1 ? static_cast<size_t>(T) : 0);
        for (int64_t p = 0; p < T; ++p)
        {
            uint32_t q[4], o[4];
            for (int k = 0; k < 4; ++k)
            { // the plan laid the slots out for cp.nvc scratch entries per bank; the kernel has `nvc`.  Scratch
              // slots name bank 0; with two banks the odd phases use the copy that names bank 1
                uint32_t const sl = cp.tet_slots[4 * static_cast<size_t>(p) + k];
                q[k]              = sl >= scratch ? sl + static_cast<uint32_t>(cp.banks * (nvc - cp.nvc) * cp.nt) : sl;
                o[k]              = sl >= scratch ? q[k] : sl + static_cast<uint32_t>(nvc * cp.nt);
                if (q[k] > 0xffffu || (cp.banks > 1 && o[k] > 0xffffu))
                {
                    why_not = "shared-memory slot does not fit 16 bits";
                    return false;
                }
            }
            slots[static_cast<size_t>(p)] = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
            if (cp.banks > 1)
                slots_odd[static_cast<size_t>(p)] = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        }
        int64_t const Q = cp.n_clusters;
        std::vector<uint4> fetch(static_cast<size_t>(nvc / 4) * static_cast<size_t>(Q),
                                 make_uint4(kNoVertex, kNoVertex, kNoVertex, kNoVertex));
        std::vector<uint4> meta(fetch.size(), make_uint4(kNoEntry, kNoEntry, kNoEntry, kNoEntry));
        for (int k4 = 0; k4 < cp.nvc / 4; ++k4)
            for (int64_t q = 0; q < Q; ++q)
            {
                size_t const at = static_cast<size_t>(k4) * Q + q;
                auto const f    = [&](int e) { return cp.cl_fetch[static_cast<size_t>(4 * k4 + e) * Q + q]; };
                auto const m    = [&](int e) {
                    return f(e) == kNoVertex ? kNoEntry : cp.cl_meta[static_cast<size_t>(4 * k4 + e) * Q + q];
                };
                fetch[at]       = make_uint4(f(0), f(1), f(2), f(3));
                meta[at]        = make_uint4(m(0), m(1), m(2), m(3));
            }
        auto const inv_mass = [&](uint32_t v) -> R {
            if (v == kNoVertex)
                return R(0);
            double const m = h.mass[v];
            return R(m > 0. ? 1. / m : 0.); // particle.cpp:39-44
        };
        std::vector<Real4<R>> fw(fetch.size());
        for (size_t i = 0; i < fetch.size(); ++i)
            fw[i] = Real4<R>{inv_mass(fetch[i].x), inv_mass(fetch[i].y), inv_mass(fetch[i].z), inv_mass(fetch[i].w)};
        h_fetch = fetch;
        std::vector<DevChunk> hchunks(cp.chunks.size());
        for (size_t i = 0; i < cp.chunks.size(); ++i)
        {
            hchunks[i].first  = cp.chunks[i].first;
            hchunks[i].cfirst = cp.chunks[i].cfirst;
            for (int m = 0; m < 8; ++m)
                hchunks[i].n[m] = cp.chunks[i].n[m];
        }
        // owned non-resident vertices and the routing of the mailboxes (scene_build.cpp, CPU-testable)
        MailboxRoutes routes;
        if (!build_mailbox_routes(h, cp, plan, nvc, world, routes))
        {
            why_not = routes.why_not;
            return false;
        }
        if (cp.banks > 1) // entries that arrive by hand-off are not polled
            for (int k4 = 0; k4 < cp.nvc / 4; ++k4)
                for (int64_t q = 0; q < Q; ++q)
                {
                    uint32_t* m = &meta[static_cast<size_t>(k4) * Q + q].x;
                    for (int e = 0; e < 4; ++e)
                        if (routes.local_prev[static_cast<size_t>(4 * k4 + e) * Q + q])
                            m[e] = kMetaLocal;
                }
        std::vector<int32_t> const& ioff  = routes.ifv_offsets;
        std::vector<uint32_t> const& ifv  = routes.ifv;
        std::vector<uint32_t> const& ifm  = routes.ifv_meta;
        std::vector<uint32_t> const& ifirst = routes.ifv_first;
        std::vector<uint32_t> const& ifv_pos = routes.ifv_pos;
        uint32_t const n_entries = routes.n_entries;
        std::vector<int32_t> soff(static_cast<size_t>(Rn) + 1, 0);
        auto const pack = [&](std::vector<uint32_t> const& r) {
            std::vector<uint4> out(fetch.size(), make_uint4(kNoBox, kNoBox, kNoBox, kNoBox));
            for (int k4 = 0; k4 < nvc / 4; ++k4)
                for (int64_t q = 0; q < Q; ++q)
                    out[static_cast<size_t>(k4) * Q + q] =
                        make_uint4(r[static_cast<size_t>(4 * k4 + 0) * Q + q], r[static_cast<size_t>(4 * k4 + 1) * Q + q],
                                   r[static_cast<size_t>(4 * k4 + 2) * Q + q], r[static_cast<size_t>(4 * k4 + 3) * Q + q]);
            return out;
        };
        std::vector<uint4> const h_to = pack(routes.to), h_to_owner = pack(routes.to_owner);

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
                sidx[static_cast<size_t>(pos)]  = static_cast<uint32_t>(i);
                sadr[static_cast<size_t>(pos)]  = plan.vertex_region[gv] == r
                                                      ? static_cast<uint32_t>(cp.banks * nvc * cp.nt) + plan.vertex_slot[gv]
                                                      : (ifv_pos[gv] | kGlobalBit);
            }
        }
        // regions that share vertices must be co-resident; the others are handed out round-robin
        std::vector<int32_t> sync_r, island_r;
        for (int32_t r = rank * per_rank; r < (rank + 1) * per_rank; ++r)
            (plan.nbr_offsets[static_cast<size_t>(r) + 1] > plan.nbr_offsets[static_cast<size_t>(r)] ? sync_r : island_r)
                .push_back(r);
        std::vector<int32_t> voff(plan.region_vtx_offsets.begin(), plan.region_vtx_offsets.end());

        // launch shape
        bool const dict = d_shapes != nullptr && n_shapes > 0 && n_shapes <= kMaxShapes;
        smem  = chunk_area_bytes(cp.n_colours) + (dict ? dict_area_bytes<R>(n_shapes) : 0) +
               static_cast<size_t>(static_cast<int64_t>(cp.banks) * nvc * cp.nt + std::max<int64_t>(plan.max_region_vertices, 1)) *
                   sizeof(Real4<R>);
        block = cp.nt;
        if (static_cast<int64_t>(smem) > kSmemBudget)
        {
            why_not = "resident vertices and scratch slots do not fit shared memory";
            return false;
        }
        // SBSB200_TRACE_STEPS=N (development aid): record clock stamps of the first N colour steps of
        // every launch, readable through sbsb200_debug_read_trace
        int trace_n = 0;
        if (char const* e = std::getenv("SBSB200_TRACE_STEPS"))
            trace_n = std::atoi(e);
        if (trace_n > 0 && nvc == 8)
        {
            kernel = pick<2, true>(block, dict);
            trace.upload(std::vector<long long>(static_cast<size_t>(Rn) * trace_n * 16, 0), st);
            args.trace       = trace.p;
            args.trace_steps = trace_n;
            trace_len        = static_cast<int64_t>(Rn) * trace_n * 16;
        }
        else
            kernel = nvc == 8 ? pick<2, false>(block, dict, cp.banks > 1) : pick<4, false>(block, dict, cp.banks > 1);
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

        std::vector<int32_t> order(sync_r);
        order.insert(order.end(), island_r.begin(), island_r.end());
        region_order.upload(order, st);
        tet_slots.upload(slots, st);
        if (cp.banks > 1)
            tet_slots_odd.upload(slots_odd, st);
        cl_to.upload(h_to, st);
        cl_to_owner.upload(h_to_owner, st);
        ifv_first.upload(ifirst, st);
        cl_meta.upload(meta, st);
        cl_w.upload(fw, st);
        chunks.upload(hchunks, st);
        vtx_off.upload(voff, st);
        vtx.upload(plan.region_vtx, st);
        ifv_off.upload(ioff, st);
        this->ifv.upload(ifv, st);
        ifv_meta.upload(ifm, st);
        surf_off.upload(soff, st);
        surf_index.upload(sidx, st);
        surf_addr.upload(sadr, st);
        box.upload(std::vector<uint4>((static_cast<size_t>(n_entries) + ifv.size() + 1) * Xchg<R>::kWords,
                                      make_uint4(0u, 0u, 0u, 0u)),
                   st);
        box_bytes = (static_cast<size_t>(n_entries) + ifv.size() + 1) * Xchg<R>::kWords * sizeof(uint4);
        error.upload(std::vector<uint32_t>(1, 0u), st);
        base = 1;

        args.s                = d;
        args.n_regions        = Rn;
        args.n_run            = static_cast<int32_t>(order.size());
        args.rank             = rank;
        args.world            = world;
        for (int r = 0; r < kMaxWorld; ++r)
            args.box_of_rank[r] = r == rank ? static_cast<void*>(box.p) : nullptr;
        args.n_colours        = cp.n_colours;

        args.nvc              = nvc;
        args.rot              = cp.rot;
        args.banks            = cp.banks;
        n_local_entries       = routes.n_local;
        args.n_clusters       = Q;
        args.region_order     = region_order.p;
        args.tet_slots        = tet_slots.p;
        args.tet_slots_odd    = cp.banks > 1 ? tet_slots_odd.p : tet_slots.p;
        args.tet_shape        = dict ? d_tet_shape : nullptr;
        args.shapes           = dict ? d_shapes : nullptr;
        args.n_shapes         = dict ? n_shapes : 0;
        args.cl_to            = cl_to.p;
        args.cl_to_owner      = cl_to_owner.p;
        args.ifv_first        = ifv_first.p;
        args.n_entries        = n_entries;
        args.cl_meta          = cl_meta.p;
        args.cl_w             = cl_w.p;
        args.chunks           = chunks.p;
        args.vtx_off          = vtx_off.p;
        args.vtx              = vtx.p;
        args.ifv_off          = ifv_off.p;
        args.ifv              = this->ifv.p;
        args.ifv_meta         = ifv_meta.p;
        args.surf_off         = surf_off.p;
        args.surf_index       = surf_index.p;
        args.surf_addr        = surf_addr.p;
        args.box              = box.p;
        args.error            = error.p;
        ready                 = true;
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
        cudaError_t const e = cudaLaunchCooperativeKernel(kernel, dim3(static_cast<unsigned>(grid)),
                                                          dim3(static_cast<unsigned>(block)), params, smem, st);
        if (e != cudaSuccess)
            throw std::runtime_error(std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(e));
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
    int64_t n_local_entries = 0; // fetch-list entries that arrive by hand-off inside their region

    // particle_t::mass() of a vertex changed: patch its inverse mass in every fetch list naming it
    void set_inverse_mass(uint32_t gv, R w, cudaStream_t st)
    {
        if (!ready)
            return;
        for (size_t i = 0; i < h_fetch.size(); ++i)
        {
            uint32_t const id[4] = {h_fetch[i].x, h_fetch[i].y, h_fetch[i].z, h_fetch[i].w};
            for (int e = 0; e < 4; ++e)
                if (id[e] == gv)
                    cudaMemcpyAsync(reinterpret_cast<R*>(cl_w.p + i) + e, &w, sizeof(R), cudaMemcpyHostToDevice, st);
        }
        cudaStreamSynchronize(st);
    }

    // true when a launch ran out of its poll budget (call after synchronising the stream)
    bool timed_out() const
    {
        uint32_t e = 0;
        if (ready && error.p)
            cudaMemcpy(&e, error.p, sizeof e, cudaMemcpyDeviceToHost);
        return e != 0;
    }
};

} // namespace sbsb200
