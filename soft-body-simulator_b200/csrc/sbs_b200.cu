// libsbsb200.so — C ABI (include/sbs_b200.h) over the sm_100a XPBD kernels.
//
// There is no CPU path in this file: every entry point that computes goes through CUDA and
// fails with SBSB200_ERR_CUDA when the device is unusable.
#include "../../include/sbs_b200.h"

#include "scene_build.h"
#include "xpbd_kernels.cuh"
#include "bvh.cuh"
#include "xpbd_resident.cuh"

#include <nvtx3/nvToolsExt.h> // header-only: ranges show up in Nsight tools, cost nothing otherwise

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

using namespace sbsb200;

namespace {

thread_local std::string g_create_error;

struct CudaError
{
    std::string msg;
};

#define CK(expr)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        cudaError_t const e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                                    \
            throw CudaError{std::string(#expr) + ": " + cudaGetErrorString(e__)};                  \
    } while (0)

template <typename T>
struct DevBuf
{
    T* p      = nullptr;
    size_t n  = 0;
    DevBuf()  = default;
    DevBuf(DevBuf const&) = delete;
    DevBuf& operator=(DevBuf const&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count)
    {
        release();
        n = count;
        CK(cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1)));
    }
    void upload(std::vector<T> const& h, cudaStream_t st)
    {
        alloc(h.size());
        if (!h.empty())
            CK(cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
    }
};

struct EngineBase
{
    virtual ~EngineBase() = default;
    virtual void build(sbsb200_ctx& c)                                                     = 0;
    virtual void step(sbsb200_ctx& c, double dt, int substeps, int iterations, int detect) = 0;
    virtual void upload(sbsb200_ctx& c, int body, double const* x, double const* v, bool sync = true) = 0;
    virtual void download(sbsb200_ctx& c, int body, double* x, double* v)                  = 0;
    virtual void set_vertices(sbsb200_ctx& c, int body, int64_t n, uint32_t const* which, double const* x,
                              double const* v)                                             = 0;
    virtual void set_masses(sbsb200_ctx& c, int64_t first, int64_t n, uint32_t const* which, double const* m) = 0;
    virtual bool remove_tets(sbsb200_ctx& c, std::vector<uint32_t> const& positions)       = 0;
    virtual uint64_t general_route_calls()                                                  = 0;
    virtual int64_t count_non_finite(sbsb200_ctx& c)                                        = 0;
    virtual void step_host_f32(sbsb200_ctx& c, int body, float const* x, float const* v, double dt, int substeps,
                               int iterations, int detect, float* x_out, float* v_out)      = 0;
    virtual void step_host_vertices_f32(sbsb200_ctx& c, int body, int64_t n, uint32_t const* which, float const* x,
                                        float const* v, double dt, int substeps, int iterations, int detect,
                                        float* x_out, float* v_out)                         = 0;
    virtual int64_t contacts(sbsb200_ctx& c, int64_t cap, int32_t* body, uint32_t* vertex,
                             int32_t* sdf_body, double* point, double* normal)             = 0;
    virtual void invalidate_graphs()                                                       = 0;
    virtual int64_t read_trace(sbsb200_ctx& c, int64_t* out, int64_t cap)                   = 0;
    virtual void kernel_times(double& ms, int64_t& launches)                                = 0;
    virtual void* mailbox_pointer(size_t& bytes)                                            = 0;
    virtual void set_peer_mailboxes(int rank, void* p)                                      = 0;
    virtual bool peers_connected()                                                          = 0;
    virtual void check_after_sync(sbsb200_ctx& c)                                           = 0;
    virtual void download_surface(sbsb200_ctx& c, int body, float* out, float const* colours = nullptr,
                                  int64_t n_colours = 0)                                   = 0;
    virtual void eval_sdf(sbsb200_ctx& c, int body, int64_t n, double const* pts, double* sd, double* grad) = 0;
};

} // namespace

struct sbsb200_ctx
{
    int device        = 0;
    int precision     = SBSB200_FP32;
    cudaStream_t stream = nullptr;
    bool own_stream   = false;
    std::string err;
    HostScene scene;
    bool finalized          = false;
    int schedule_request    = SBSB200_SCHED_AUTO;
    int schedule            = SBSB200_SCHED_GRAPH;
    double collision_alpha  = 1e-8; // simulation_parameters.h:24
    ClusterPlan green_plan;
    ColourClass dist_cc;
    RegionPlan plan;
    ExchangePlan xplan;          // resident schedule: who pulls and pushes which shared vertex when (host side)
    int region_shape  = SBSB200_REGIONS_COMPACT;
    int trace_steps   = 0;       // development aid (sbsb200_debug_trace_steps)
    uint64_t general_calls0 = 0; // device counter of green_general calls when the scene was finalized
    std::vector<uint32_t> order; // exported serial order (insertion indices)
    std::vector<int32_t> vertex_body;
    std::unique_ptr<EngineBase> engine;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed            = false;
    int64_t kernels       = 0;
    int64_t frames        = 0;
    int64_t last_contacts = 0;
    int sm_count          = 0;
    int broadphase = SBSB200_BROADPHASE_NONE;
    int rank = 0, world = 1; // decomposition of one scene over several GPUs (sbsb200_set_partition)
    std::vector<void*> ipc_opened;
    int64_t n_surface     = 0;
    bool any_damping      = false;
    std::string schedule_note;
    // sbsb200_remove_constraints: constraints taken out since finalize (by insertion index), and where a tet
    // constraint sits in the device arrays (built at the first removal)
    std::vector<char> removed;
    std::vector<uint32_t> position_of_insertion;
    int64_t n_removed = 0;
};

namespace {

// rounding-exact helpers: predict/commit must match the reference bit for bit in the fp64
// build, so they are kept free of FMA contraction there (see k_predict / k_integrate).

template <typename R>
struct Engine final : EngineBase
{
    DeviceScene<R> d{};
    DevBuf<Real4<R>> pos, prev, vel, tet_r0, tet_r1, tet_r2, materials, dist_p, surf_pos, contact_q, contact_n;
    DevBuf<uint4> tet_v;
    DevBuf<uint2> dist_v;
    DevBuf<R> tet_lambda, dist_lambda;
    DevBuf<uint32_t> surf_tri;       // boundary triangles of all bodies (surface-vertex indices local to the body)
    DevBuf<Real4<R>> surf_normal;
    DevBuf<float> surf_out;
    std::vector<int64_t> tri_offset; // per body: first triangle in surf_tri
    DevBuf<uint8_t> tet_shape;       // rest-shape dictionary (persistent schedule): index per tet ...
    DevBuf<Real4<R>> shape_records;  // ... into the distinct (r0, r1, r2) records
    DevBuf<uint32_t> surf_v, surf_first, contact_v, contact_count;
    DevBuf<int32_t> surf_body;
    DevBuf<typename DeviceScene<R>::Sdf> sdf;
    DevBuf<R> grid_nodes;            // node values of all grid SDFs, one after the other
    std::vector<int32_t> sdf_index;  // per body: its SDF record (detection uses the first n_active_sdf), -1 for tet bodies
    int32_t n_active_sdf = 0;
    DevBuf<double> stage_x, stage_v; // raw host-format staging for upload/download
    // BVH broadphase (bvh.cuh)
    BvhView<R> bvh{};
    DevBuf<uint64_t> bvh_keys, bvh_keys_sorted;
    DevBuf<uint32_t> bvh_leaf_in, bvh_leaf_surface, bvh_leaf_of_surface, bvh_counter;
    DevBuf<Real4<R>> bvh_sphere;
    DevBuf<unsigned char> bvh_temp;
    size_t bvh_temp_bytes = 0;
    // The order of the leaves (keys -> radix sort) is kept for kBvhTopologyFrames frames; the bounding
    // spheres are refitted at every detection, so an older order only costs culling quality, never a
    // contact.  (The reference builds its KD-tree once and only ever refits it, bvh_model.cpp:24-28, :102-126.)
    static constexpr int kBvhTopologyFrames = 8;
    int bvh_topology_valid_frames = 0; // 0: sort at the next detection
    int bvh_sort_end_bit          = 64; // key bits in use: 32 Morton bits + the bits of the body index
    size_t bvh_top_bytes          = 0;  // shared memory of k_detect_all for the top of the sphere tree (BvhView::top_in_detect)
    ResidentPlan<R> pp; // resident schedule resources (may be inactive)
    int32_t n_base_shapes = 0; // rest-shape dictionary: records [n_base_shapes, 2 n_base_shapes) are the zero-volume twins
    // CUDA-event pairs around the launches of the dominant kernel (persistent schedule: the substep
    // kernel), folded into a running sum when the statistics are read
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kev;
    size_t kev_used        = 0;
    double kernel_ms_sum   = 0;
    int64_t kernel_ms_count = 0;
    std::map<std::tuple<double, int, int, int>, cudaGraphExec_t> graphs;
    std::map<std::tuple<double, int, int, int>, int64_t> graph_kernels;

    ~Engine() override
    {
        invalidate_graphs();
        for (auto& e : kev)
        {
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
    }

    void* mailbox_pointer(size_t& bytes) override
    {
        bytes = pp.box_bytes;
        return pp.ready ? static_cast<void*>(pp.box.p) : nullptr;
    }
    void set_peer_mailboxes(int rank, void* p) override { pp.args.box_of_rank[rank] = p; }
    bool peers_connected() override { return !pp.ready || pp.peers_connected(); }

    void kernel_times(double& ms, int64_t& launches) override
    {
        for (size_t i = 0; i < kev_used; ++i)
        {
            float t = 0.f;
            if (cudaEventSynchronize(kev[i].second) == cudaSuccess &&
                cudaEventElapsedTime(&t, kev[i].first, kev[i].second) == cudaSuccess)
            {
                kernel_ms_sum += t;
                ++kernel_ms_count;
            }
        }
        kev_used = 0;
        ms       = kernel_ms_sum;
        launches = kernel_ms_count;
    }

    void invalidate_graphs() override
    {
        for (auto& kv : graphs)
            cudaGraphExecDestroy(kv.second);
        graphs.clear();
        graph_kernels.clear();
    }

    void build(sbsb200_ctx& c) override
    {
        HostScene const& h = c.scene;
        cudaStream_t st    = c.stream;
        int64_t const V = h.n_vertices(), T = h.n_tets(), D = h.n_dist();

        std::vector<Real4<R>> hpos(static_cast<size_t>(V)), hprev(static_cast<size_t>(V)),
            hvel(static_cast<size_t>(V), Real4<R>{R(0), R(0), R(0), R(0)});
        for (int64_t i = 0; i < V; ++i)
        {
            double const m  = h.mass[static_cast<size_t>(i)];
            double const iw = m > 0. ? 1. / m : 0.; // particle.cpp:39-44
            hpos[static_cast<size_t>(i)]  = {R(h.x0[3 * i]), R(h.x0[3 * i + 1]), R(h.x0[3 * i + 2]), R(iw)};
            hprev[static_cast<size_t>(i)] = {R(h.x0[3 * i]), R(h.x0[3 * i + 1]), R(h.x0[3 * i + 2]), R(0)};
        }
        pos.upload(hpos, st);
        prev.upload(hprev, st);
        vel.upload(hvel, st);

        // green constraints in schedule order
        std::vector<uint4> hv(static_cast<size_t>(T));
        std::vector<Real4<R>> r0(static_cast<size_t>(T)), r1(static_cast<size_t>(T)), r2(static_cast<size_t>(T));
        for (int64_t p = 0; p < T; ++p)
        {
            uint32_t const t  = c.green_plan.storage_order[static_cast<size_t>(p)];
            uint32_t const* v = &h.tets[4 * static_cast<size_t>(t)];
            hv[static_cast<size_t>(p)] = make_uint4(v[0], v[1], v[2], v[3]);
            // Dm columns x0_1-x0_4, x0_2-x0_4, x0_3-x0_4 (green_constraint.cpp:38-41)
            double m[9];
            for (int r = 0; r < 3; ++r)
                for (int col = 0; col < 3; ++col)
                    m[3 * r + col] = h.x0[3 * static_cast<size_t>(v[col]) + r] - h.x0[3 * static_cast<size_t>(v[3]) + r];
            double const det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                               m[2] * (m[3] * m[7] - m[4] * m[6]);
            double const id = 1.0 / det;
            double inv[9]   = {(m[4] * m[8] - m[5] * m[7]) * id, (m[2] * m[7] - m[1] * m[8]) * id,
                               (m[1] * m[5] - m[2] * m[4]) * id, (m[5] * m[6] - m[3] * m[8]) * id,
                               (m[0] * m[8] - m[2] * m[6]) * id, (m[2] * m[3] - m[0] * m[5]) * id,
                               (m[3] * m[7] - m[4] * m[6]) * id, (m[1] * m[6] - m[0] * m[7]) * id,
                               (m[0] * m[4] - m[1] * m[3]) * id};
            double const V0 = det / 6.0; // :44
            r0[static_cast<size_t>(p)] = {R(inv[0]), R(inv[1]), R(inv[2]), R(inv[3])};
            r1[static_cast<size_t>(p)] = {R(inv[4]), R(inv[5]), R(inv[6]), R(inv[7])};
            R matv;
            int const mi = h.tet_material[t];
            if (sizeof(R) == 4)
            {
                float f;
                std::memcpy(&f, &mi, 4);
                matv = R(f);
            }
            else
                matv = R(mi);
            r2[static_cast<size_t>(p)] = {R(inv[8]), R(V0), matv, R(0)};
        }
        { // rest-shape dictionary: distinct (DmInv, V0, material) records, bit for bit
            std::map<std::array<R, 12>, uint32_t> index;
            std::vector<uint8_t> shape(static_cast<size_t>(T), 0);
            std::vector<Real4<R>> records;
            bool small = true;
            for (int64_t p = 0; p < T && small; ++p)
            {
                auto const& a0 = r0[static_cast<size_t>(p)];
                auto const& a1 = r1[static_cast<size_t>(p)];
                auto const& a2 = r2[static_cast<size_t>(p)];
                std::array<R, 12> const key = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
                auto it = index.find(key);
                if (it == index.end())
                {
                    if (index.size() >= static_cast<size_t>(kMaxShapes))
                    {
                        small = false;
                        break;
                    }
                    it = index.emplace(key, static_cast<uint32_t>(index.size())).first;
                    records.push_back(a0);
                    records.push_back(a1);
                    records.push_back(a2);
                    Material const& mt = h.materials[h.tet_material[c.green_plan.storage_order[static_cast<size_t>(p)]]];
                    records.push_back(Real4<R>{R(mt.mu), R(mt.lambda), R(mt.alpha), R(mt.beta)});
                }
                shape[static_cast<size_t>(p)] = static_cast<uint8_t>(it->second);
            }
            n_base_shapes = 0;
            if (small && T > 0 && 2 * index.size() <= static_cast<size_t>(kMaxShapes))
            { // a twin of every record with rest volume zero: what a removed constraint points at (remove_tets)
                n_base_shapes      = static_cast<int32_t>(index.size());
                size_t const words = records.size();
                for (size_t w = 0; w < words; ++w)
                {
                    Real4<R> q = records[w];
                    if (w % kShapeWords == 2)
                        q.y = R(0);
                    records.push_back(q);
                }
            }
            int32_t const n_records = static_cast<int32_t>(records.size() / kShapeWords);
            if (small && T > 0)
            {
                tet_shape.upload(shape, st);
                shape_records.upload(records, st);
                pp.d_tet_shape = tet_shape.p;
                pp.d_shapes    = shape_records.p;
                pp.n_shapes    = n_records;
                d.tet_shape    = tet_shape.p;
                d.shapes       = shape_records.p;
                d.n_shapes     = n_records;
            }
        }
        tet_v.upload(hv, st);
        tet_r0.upload(r0, st);
        tet_r1.upload(r1, st);
        tet_r2.upload(r2, st);
        tet_lambda.alloc(static_cast<size_t>(T));
        CK(cudaMemsetAsync(tet_lambda.p, 0, sizeof(R) * std::max<int64_t>(T, 1), st));
        std::vector<Real4<R>> hm(h.materials.size());
        for (size_t i = 0; i < hm.size(); ++i)
            hm[i] = {R(h.materials[i].mu), R(h.materials[i].lambda), R(h.materials[i].alpha), R(h.materials[i].beta)};
        materials.upload(hm, st);

        // distance constraints in schedule order
        std::vector<uint2> dv(static_cast<size_t>(D));
        std::vector<Real4<R>> dp(static_cast<size_t>(D));
        for (int64_t p = 0; p < D; ++p)
        {
            uint32_t const i = c.dist_cc.order[static_cast<size_t>(p)];
            dv[static_cast<size_t>(p)] = make_uint2(h.dist_pairs[2 * static_cast<size_t>(i)], h.dist_pairs[2 * static_cast<size_t>(i) + 1]);
            dp[static_cast<size_t>(p)] = {R(h.dist_rest[i]), R(h.dist_alpha[i]), R(h.dist_beta[i]), R(0)};
        }
        dist_v.upload(dv, st);
        dist_p.upload(dp, st);
        dist_lambda.alloc(static_cast<size_t>(D));
        CK(cudaMemsetAsync(dist_lambda.p, 0, sizeof(R) * std::max<int64_t>(D, 1), st));

        // surfaces + SDFs
        std::vector<uint32_t> sv;
        std::vector<int32_t> sb;
        std::vector<typename DeviceScene<R>::Sdf> hs;
        {
            std::vector<R> pool;
            for (HostBody const& hb : h.bodies)
                if (hb.kind == BodyKind::sdf && hb.sdf_kind == SdfKind::grid)
                    for (double v : hb.grid_nodes)
                        pool.push_back(R(v));
            grid_nodes.upload(pool, st);
        }
        // SDF records: the bodies that take part in detection first (DeviceScene::n_sdf of them), the
        // others behind them so that sbsb200_eval_sdf can still evaluate them
        sdf_index.assign(h.bodies.size(), -1);
        std::vector<size_t> grid_at(h.bodies.size(), 0);
        {
            size_t at = 0;
            for (size_t b = 0; b < h.bodies.size(); ++b)
            {
                grid_at[b] = at;
                at += h.bodies[b].grid_nodes.size();
            }
        }
        n_active_sdf = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (size_t b = 0; b < h.bodies.size(); ++b)
            {
                HostBody const& hb = h.bodies[b];
                if (hb.kind != BodyKind::sdf || hb.collideable != (pass == 0))
                    continue;
                typename DeviceScene<R>::Sdf f{};
                f.kind = static_cast<int32_t>(hb.sdf_kind);
                f.body = static_cast<int32_t>(b);
                for (int k = 0; k < 3; ++k)
                {
                    f.a[k]      = R(hb.a[k]);
                    f.b[k]      = R(hb.b[k]);
                    f.vmin[k]   = R(hb.volume[k]);
                    f.vmax[k]   = R(hb.volume[3 + k]);
                    f.grid_n[k] = hb.grid_n[k];
                }
                f.r          = R(hb.r);
                f.grid_nodes = grid_nodes.p + grid_at[b];
                sdf_index[b] = static_cast<int32_t>(hs.size());
                hs.push_back(f);
                n_active_sdf += pass == 0;
            }
        // surface vertices of every tet body; a body that is not handed to the cd system keeps its
        // surface (the renderer reads it) but is marked with a negative body id, which detection skips
        for (size_t b = 0; b < h.bodies.size(); ++b)
        {
            HostBody const& hb = h.bodies[b];
            if (hb.kind == BodyKind::tet)
                for (uint32_t lv : hb.surf_to_tet)
                {
                    sv.push_back(static_cast<uint32_t>(hb.v_offset + lv));
                    sb.push_back(hb.collideable ? static_cast<int32_t>(b) : ~static_cast<int32_t>(b));
                }
        }
        int64_t const Vs = static_cast<int64_t>(sv.size());
        {
            std::vector<uint32_t> tri;
            tri_offset.assign(h.bodies.size() + 1, 0);
            for (size_t b = 0; b < h.bodies.size(); ++b)
            {
                tri_offset[b] = static_cast<int64_t>(tri.size() / 3);
                tri.insert(tri.end(), h.bodies[b].surf_triangles.begin(), h.bodies[b].surf_triangles.end());
            }
            tri_offset[h.bodies.size()] = static_cast<int64_t>(tri.size() / 3);
            surf_tri.upload(tri, st);
            surf_normal.alloc(static_cast<size_t>(Vs));
        }
        surf_v.upload(sv, st);
        surf_body.upload(sb, st);
        surf_pos.alloc(static_cast<size_t>(Vs));
        surf_first.alloc(static_cast<size_t>(Vs));
        CK(cudaMemsetAsync(surf_first.p, 0xff, sizeof(uint32_t) * std::max<int64_t>(Vs, 1), st));
        sdf.upload(hs, st);
        int64_t const cap = Vs * static_cast<int64_t>(hs.size());
        contact_v.alloc(static_cast<size_t>(cap));
        contact_q.alloc(static_cast<size_t>(cap));
        contact_n.alloc(static_cast<size_t>(cap));
        contact_count.alloc(1);
        CK(cudaMemsetAsync(contact_count.p, 0, sizeof(uint32_t), st));

        d.n_vertices      = V;
        d.pos             = pos.p;
        d.prev            = prev.p;
        d.vel             = vel.p;
        d.n_tets          = T;
        d.tet_v           = tet_v.p;
        d.tet_r0          = tet_r0.p;
        d.tet_r1          = tet_r1.p;
        d.tet_r2          = tet_r2.p;
        d.tet_lambda      = tet_lambda.p;
        d.materials       = materials.p;
        d.n_dist          = D;
        d.dist_v          = dist_v.p;
        d.dist_p          = dist_p.p;
        d.dist_lambda     = dist_lambda.p;
        d.n_surface       = Vs;
        d.surf_v          = surf_v.p;
        d.surf_body       = surf_body.p;
        d.surf_pos        = surf_pos.p;
        d.surf_first      = surf_first.p;
        d.n_sdf           = n_active_sdf;
        d.sdf             = sdf.p;
        d.contact_cap     = cap;
        d.contact_v       = contact_v.p;
        d.contact_q       = contact_q.p;
        d.contact_n       = contact_n.p;
        d.contact_count   = contact_count.p;
        d.collision_alpha = R(c.collision_alpha);
        c.n_surface       = Vs;
        if (c.broadphase == SBSB200_BROADPHASE_BVH && Vs > 0 && !hs.empty())
        { // linear BVH over the surface vertices, rebuilt at every detection
            size_t const n = static_cast<size_t>(Vs);
            bvh_keys.alloc(n);
            bvh_keys_sorted.alloc(n);
            bvh_leaf_in.alloc(n);
            bvh_leaf_surface.alloc(n);
            bvh_leaf_of_surface.alloc(n);
            bvh_counter.alloc(1);
            CK(cudaMemsetAsync(bvh_counter.p, 0, sizeof(uint32_t), st));
            // implicit complete binary tree over the sorted leaves: level l has ceil(n / 2^l) nodes
            int64_t offset = 0, count = Vs;
            bvh.n_levels   = 1;
            for (int l = 1; l < kBvhMaxLevels && count > 1; ++l)
            {
                count               = (count + 1) / 2;
                bvh.level_offset[l] = offset;
                bvh.level_count[l]  = count;
                offset += count;
                bvh.n_levels = l + 1;
            }
            bvh.level_offset[0] = 0;
            bvh.level_count[0]  = Vs;
            bvh_sphere.alloc(static_cast<size_t>(std::max<int64_t>(offset, 1)));
            bvh_top_bytes = (2 * static_cast<size_t>(bvh.level_count[8]) + 32) * (sizeof(Real4<R>) + sizeof(uint32_t));
            bvh.top_in_detect =
                bvh.n_levels > 9 && bvh.level_count[8] <= kBvhTopNodes && bvh_top_bytes <= 40 * 1024 ? 1 : 0;
            if (!bvh.top_in_detect)
                bvh_top_bytes = 0;
            bvh_sort_end_bit = 33;
            while (bvh_sort_end_bit < 64 && (uint64_t{1} << (bvh_sort_end_bit - 32)) < 2 * c.scene.bodies.size())
                ++bvh_sort_end_bit;
            bvh.n_bodies = static_cast<int32_t>(c.scene.bodies.size());
            CK(cub::DeviceRadixSort::SortPairs(nullptr, bvh_temp_bytes, bvh_keys.p, bvh_keys_sorted.p, bvh_leaf_in.p,
                                               bvh_leaf_surface.p, static_cast<int>(Vs), 0, 64, st));
            bvh_temp.alloc(bvh_temp_bytes);
            bvh.n            = Vs;
            bvh.keys         = bvh_keys.p;
            bvh.keys_sorted  = bvh_keys_sorted.p;
            bvh.leaf_in      = bvh_leaf_in.p;
            bvh.leaf_surface = bvh_leaf_surface.p;
            bvh.leaf_of_surface = bvh_leaf_of_surface.p;
            bvh.done_counter = bvh_counter.p;
            bvh.sphere       = bvh_sphere.p;
            bvh.sdf          = sdf.p;
            // quantisation box of the Morton codes: the rest shape, generously enlarged (tree quality only)
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int64_t i = 0; i < V; ++i)
                for (int k = 0; k < 3; ++k)
                {
                    lo[k] = std::min(lo[k], h.x0[3 * i + k]);
                    hi[k] = std::max(hi[k], h.x0[3 * i + k]);
                }
            for (int k = 0; k < 3; ++k)
            {
                double const ext  = std::max(hi[k] - lo[k], 1e-6);
                bvh.lo[k]         = R(lo[k] - ext);
                bvh.inv_extent[k] = R(1.0 / (3.0 * ext));
            }
        }

        if (Vs > 0)
        {
            k_surface_gather<R><<<static_cast<unsigned>((Vs + 255) / 256), 256, 0, st>>>(d);
            ++c.kernels;
        }
        if (c.schedule == SBSB200_SCHED_PERSISTENT)
        {
            bool ok = false;
            try
            {
                ok = pp.build(c.scene, c.green_plan, c.plan, c.xplan, d, st, c.sm_count, c.rank, c.world, c.trace_steps);
            }
            catch (std::exception const& e)
            {
                pp.why_not = e.what();
            }
            if (!ok)
            { // the (colour, region) order is still a valid colouring for the per-colour kernels
                c.schedule      = SBSB200_SCHED_GRAPH;
                c.schedule_note = "persistent schedule unavailable: " + pp.why_not;
            }
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
    }

    // enqueue one frame; returns number of kernels enqueued
    int64_t enqueue(sbsb200_ctx& c, double dt_frame, int substeps, int iterations, int detect, cudaStream_t st)
    {
        int64_t launched     = 0;
        R const dt           = R(dt_frame / static_cast<double>(substeps)); // timestep.cpp:22
        int64_t const V      = d.n_vertices;
        unsigned const gridV = static_cast<unsigned>((V + 255) / 256);
        bool const collide   = d.n_sdf > 0 && d.n_surface > 0;
        unsigned const gridS = static_cast<unsigned>((d.n_surface + 255) / 256);
        unsigned const gridC = static_cast<unsigned>((d.contact_cap + 255) / 256);
        d.collision_alpha    = R(c.collision_alpha);

        // leaves re-sorted once per frame at most (every kBvhTopologyFrames frames when the frame is enqueued
        // eagerly; a captured frame is replayed as it is, so it always sorts), refitted at every detection
        cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
        CK(cudaStreamIsCapturing(st, &capture));
        bool const eager        = capture == cudaStreamCaptureStatusNone;
        bool bvh_topology_built = false;
        if (eager && bvh_topology_valid_frames > 0)
        {
            bvh_topology_built = true;
            --bvh_topology_valid_frames;
        }
        auto detect_now = [&] {
            if (!collide)
                return;
            NvtxRange const range("detection");
            if (bvh.n <= 0) // (with the broadphase the refit kernel zeroes the count)
                CK(cudaMemsetAsync(d.contact_count, 0, sizeof(uint32_t), st));
            if (bvh.n > 0)
            { // broadphase: (once per frame: keys -> sort) -> refit of the bounding spheres; the cull itself
              // is the ancestor walk inside k_detect_all
              // (the reference builds its KD-tree once and refits every sphere per frame, bvh_model.cpp:102-126)
                int const n = static_cast<int>(bvh.n);
                if (!bvh_topology_built)
                {
                    k_bvh_keys<R><<<gridS, 256, 0, st>>>(d, bvh);
                    CK(cub::DeviceRadixSort::SortPairs(bvh_temp.p, bvh_temp_bytes, bvh.keys, bvh.keys_sorted,
                                                       bvh.leaf_in, bvh.leaf_surface, n, 0, bvh_sort_end_bit, st));
                    ++launched;
                    bvh_topology_built = true;
                    if (eager)
                        bvh_topology_valid_frames = kBvhTopologyFrames - 1;
                }
                k_bvh_fit<R><<<static_cast<unsigned>((n + kBvhLeafBlock - 1) / kBvhLeafBlock), kBvhLeafBlock, 0, st>>>(d, bvh);
                ++launched;
            }
            k_detect_all<R><<<gridS, 256, bvh_top_bytes, st>>>(d, bvh);
            ++launched;
        };
        if (detect == SBSB200_DETECT_PER_FRAME)
            detect_now(); // timestep.cpp:29-30
        for (int s = 0; s < substeps; ++s)
        {
            if (detect == SBSB200_DETECT_PER_SUBSTEP)
                detect_now();
            if (c.schedule == SBSB200_SCHED_PERSISTENT)
            {
                constexpr size_t kMaxTimedLaunches = 4096;
                bool const timed = kev_used < kMaxTimedLaunches;
                if (timed && kev_used == kev.size())
                {
                    cudaEvent_t e0 = nullptr, e1 = nullptr;
                    CK(cudaEventCreate(&e0));
                    CK(cudaEventCreate(&e1));
                    kev.emplace_back(e0, e1);
                }
                if (timed)
                    CK(cudaEventRecord(kev[kev_used].first, st));
                {
                    NvtxRange const range("substep (resident kernel)");
                    launched += pp.substep(d, dt, iterations, collide, st);
                }
                if (timed)
                    CK(cudaEventRecord(kev[kev_used++].second, st));
            }
            else
            {
                if (V > 0)
                {
                    k_predict<R><<<gridV, 256, 0, st>>>(d, dt);
                    ++launched;
                }
                for (int k = 0; k < iterations; ++k)
                {
                    int const first = k == 0;
                    if (collide)
                    { // gauss_seidel_solver.cpp:28-31
                        k_project_collision<R><<<gridC, 256, 0, st>>>(d, dt, first);
                        ++launched;
                    }
                    // gauss_seidel_solver.cpp:32-35 in the exported colour order
                    for (int32_t col = 0; col < c.green_plan.n_colours; ++col)
                    {
                        // regions and parts of one colour are adjacent in storage, but each has its own
                        // column layout -> one launch each (a single launch when there is no region plan)
                        ChunkDesc const* hd =
                            &c.green_plan.chunks[static_cast<size_t>(col) * 2 * c.green_plan.n_regions];
                        for (int32_t rp = 0; rp < 2 * c.green_plan.n_regions; ++rp)
                        {
                            ChunkDesc const& h2 = hd[rp];
                            if (h2.n[0] == 0)
                                continue;
                            DevChunk dc;
                            dc.first  = h2.first;
                            dc.cfirst = h2.cfirst;
                            for (int m = 0; m < 8; ++m)
                                dc.n[m] = h2.n[m];
                            unsigned const g = static_cast<unsigned>((h2.n[0] + 127) / 128);
                            size_t const dict_bytes = static_cast<size_t>(kShapeWords * d.n_shapes) * sizeof(Real4<R>);
                            if (c.any_damping && d.n_shapes > 0)
                                k_project_green<R, true, true><<<g, 128, dict_bytes, st>>>(d, dc, dt, first);
                            else if (c.any_damping)
                                k_project_green<R, true, false><<<g, 128, 0, st>>>(d, dc, dt, first);
                            else if (d.n_shapes > 0)
                                k_project_green<R, false, true><<<g, 128, dict_bytes, st>>>(d, dc, dt, first);
                            else
                                k_project_green<R, false, false><<<g, 128, 0, st>>>(d, dc, dt, first);
                            ++launched;
                        }
                    }
                    for (int32_t col = 0; col < c.dist_cc.n_colours; ++col)
                    {
                        int64_t const b = c.dist_cc.offsets[col], e = c.dist_cc.offsets[col + 1];
                        if (e == b)
                            continue;
                        k_project_distance<R><<<static_cast<unsigned>((e - b + 127) / 128), 128, 0, st>>>(
                            d, b, static_cast<int32_t>(e - b), dt, first);
                        ++launched;
                    }
                }
                if (V > 0)
                {
                    k_integrate<R><<<gridV, 256, 0, st>>>(d, dt);
                    ++launched;
                }
            }
            if (detect == SBSB200_DETECT_PER_SUBSTEP && d.n_surface > 0 && c.schedule != SBSB200_SCHED_PERSISTENT)
            {
                k_surface_gather<R><<<gridS, 256, 0, st>>>(d);
                ++launched;
            }
        }
        if (detect == SBSB200_DETECT_PER_FRAME && d.n_surface > 0 && c.schedule != SBSB200_SCHED_PERSISTENT)
        { // timestep.cpp:60-66 (the persistent kernel writes the surface copy itself)
            k_surface_gather<R><<<gridS, 256, 0, st>>>(d);
            ++launched;
        }
        CK(cudaGetLastError());
        return launched;
    }

    struct NvtxRange
    {
        explicit NvtxRange(char const* name) { nvtxRangePushA(name); }
        ~NvtxRange() { nvtxRangePop(); }
    };

    void step(sbsb200_ctx& c, double dt, int substeps, int iterations, int detect) override
    {
        NvtxRange const range("sbsb200_step (timestep_t::step)");
        cudaStream_t st = c.stream;
        CK(cudaEventRecord(c.ev0, st));
        bool const legacy_stream = st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
        if (c.schedule == SBSB200_SCHED_PERSISTENT || legacy_stream)
        { // persistent: a handful of launches per frame, and the progress base changes every launch.
          // legacy default stream: capture is not permitted on it, so the same kernels go out eagerly
            c.kernels += enqueue(c, dt, substeps, iterations, detect, st);
            CK(cudaEventRecord(c.ev1, st));
            c.timed = true;
            return;
        }
        auto const key = std::make_tuple(dt, substeps, iterations, detect);
        auto it        = graphs.find(key);
        if (it == graphs.end())
        {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int64_t n = 0;
            try
            {
                n = enqueue(c, dt, substeps, iterations, detect, st);
            }
            catch (...)
            {
                cudaStreamEndCapture(st, &graph);
                if (graph)
                    cudaGraphDestroy(graph);
                throw;
            }
            CK(cudaStreamEndCapture(st, &graph));
            cudaGraphExec_t exec = nullptr;
            cudaError_t const e  = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess)
                throw CudaError{std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)};
            it                 = graphs.emplace(key, exec).first;
            graph_kernels[key] = n;
        }
        CK(cudaGraphLaunch(it->second, st));
        c.kernels += graph_kernels[key];
        CK(cudaEventRecord(c.ev1, st));
        c.timed = true;
    }

    // Host arrays cross PCIe as raw doubles (a single DMA each when the caller's memory is
    // pinned); the conversion to the device layout runs on the GPU.
    void ensure_staging(int64_t n)
    {
        if (static_cast<int64_t>(stage_x.n) < 3 * n)
        {
            stage_x.alloc(static_cast<size_t>(3 * n));
            stage_v.alloc(static_cast<size_t>(3 * n));
        }
    }

    void upload(sbsb200_ctx& c, int body, double const* x, double const* v, bool sync) override
    {
        HostBody const& hb = c.scene.bodies[static_cast<size_t>(body)];
        int64_t const n    = hb.n_vertices;
        if (n == 0)
            return;
        ensure_staging(n);
        CK(cudaMemcpyAsync(stage_x.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
        if (v)
            CK(cudaMemcpyAsync(stage_v.p, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
        k_unpack_state<R, double><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(
            d, hb.v_offset, n, stage_x.p, v ? stage_v.p : nullptr);
        ++c.kernels;
        if (d.n_surface > 0)
        {
            k_surface_gather<R><<<static_cast<unsigned>((d.n_surface + 255) / 256), 256, 0, c.stream>>>(d);
            ++c.kernels;
        }
        CK(cudaGetLastError());
        if (sync) // the caller may reuse x / v as soon as we return (step_host: they stay valid until its download)
            CK(cudaStreamSynchronize(c.stream));
    }

    void set_vertices(sbsb200_ctx& c, int body, int64_t n, uint32_t const* which, double const* x,
                      double const* v) override
    {
        HostBody const& hb = c.scene.bodies[static_cast<size_t>(body)];
        ensure_staging(n);
        DevBuf<uint32_t> ids;
        ids.alloc(static_cast<size_t>(n));
        CK(cudaMemcpyAsync(ids.p, which, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c.stream));
        CK(cudaMemcpyAsync(stage_x.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
        if (v)
            CK(cudaMemcpyAsync(stage_v.p, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
        k_scatter_state<R, double><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(d, hb.v_offset, n, ids.p, stage_x.p,
                                                                                     v ? stage_v.p : nullptr);
        ++c.kernels;
        if (d.n_surface > 0)
        {
            k_surface_gather<R><<<static_cast<unsigned>((d.n_surface + 255) / 256), 256, 0, c.stream>>>(d);
            ++c.kernels;
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c.stream));
    }

    void download(sbsb200_ctx& c, int body, double* x, double* v) override
    {
        HostBody const& hb = c.scene.bodies[static_cast<size_t>(body)];
        int64_t const n    = hb.n_vertices;
        if (n == 0 || (!x && !v))
            return;
        ensure_staging(n);
        k_pack_state<R, double><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(
            d, hb.v_offset, n, x ? stage_x.p : nullptr, v ? stage_v.p : nullptr);
        ++c.kernels;
        CK(cudaGetLastError());
        if (x)
            CK(cudaMemcpyAsync(x, stage_x.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
        if (v)
            CK(cudaMemcpyAsync(v, stage_v.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
    }

    void check_after_sync(sbsb200_ctx& c) override { check_persistent(c); }

    void download_surface(sbsb200_ctx& c, int body, float* out, float const* colours, int64_t n_colours) override
    {
        HostBody const& hb = c.scene.bodies[static_cast<size_t>(body)];
        int64_t const n    = static_cast<int64_t>(hb.surf_to_tet.size());
        int64_t const nt   = tri_offset[static_cast<size_t>(body) + 1] - tri_offset[static_cast<size_t>(body)];
        if (n == 0)
            return;
        int64_t const width = colours ? 9 : 6; // floats per vertex
        if (surf_out.n < static_cast<size_t>(width * n))
            surf_out.alloc(static_cast<size_t>(width * n));
        DevBuf<float> dcol;
        if (colours)
        {
            dcol.alloc(static_cast<size_t>(3 * n_colours));
            CK(cudaMemcpyAsync(dcol.p, colours, sizeof(float) * 3 * static_cast<size_t>(n_colours), cudaMemcpyHostToDevice,
                               c.stream));
        }
        CK(cudaMemsetAsync(surf_normal.p + hb.s_offset, 0, sizeof(Real4<R>) * static_cast<size_t>(n), c.stream));
        if (nt > 0)
            k_surface_normals<R><<<static_cast<unsigned>((nt + 255) / 256), 256, 0, c.stream>>>(
                d, tri_offset[static_cast<size_t>(body)], nt, surf_tri.p, hb.s_offset, surf_normal.p);
        k_surface_pack<R><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(
            d, hb.s_offset, n, surf_normal.p, surf_out.p, colours ? dcol.p : nullptr, n_colours);
        c.kernels += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, surf_out.p, sizeof(float) * static_cast<size_t>(width * n), cudaMemcpyDeviceToHost,
                           c.stream));
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
    }

    void eval_sdf(sbsb200_ctx& c, int body, int64_t n, double const* pts, double* sd, double* grad) override
    {
        int const k = sdf_index[static_cast<size_t>(body)];
        DevBuf<double> in, out;
        in.alloc(static_cast<size_t>(3 * n));
        out.alloc(static_cast<size_t>(4 * n));
        CK(cudaMemcpyAsync(in.p, pts, sizeof(double) * 3 * static_cast<size_t>(n), cudaMemcpyHostToDevice, c.stream));
        k_eval_sdf<R><<<static_cast<unsigned>((n + 127) / 128), 128, 0, c.stream>>>(d, k, n, in.p, out.p);
        ++c.kernels;
        CK(cudaGetLastError());
        std::vector<double> h(static_cast<size_t>(4 * n));
        CK(cudaMemcpyAsync(h.data(), out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        for (int64_t i = 0; i < n; ++i)
        {
            sd[i] = h[static_cast<size_t>(4 * i)];
            for (int a = 0; a < 3; ++a)
                grad[3 * i + a] = h[static_cast<size_t>(4 * i + 1 + a)];
        }
    }

    void check_persistent(sbsb200_ctx& c)
    {
        if (c.schedule == SBSB200_SCHED_PERSISTENT && pp.timed_out())
            throw CudaError{"the persistent kernel ran out of its poll budget waiting for a vertex of another region"};
    }

    int64_t read_trace(sbsb200_ctx& c, int64_t* out, int64_t cap) override
    {
        int64_t const n = std::min<int64_t>(cap, pp.trace_len);
        if (n > 0 && out)
        {
            CK(cudaStreamSynchronize(c.stream));
            CK(cudaMemcpy(out, pp.trace.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
        }
        return pp.trace_len;
    }

    // particle_t::mass() of n vertices (first + which[i]): one staging copy, one kernel, one synchronisation.
    // The inverse mass rides in pos[].w; the resident kernel reads it at the start of every substep.
    // simulation_t::remove_constraint (simulation.cpp:34-39) without re-planning: the tets at these storage positions
    // get rest volume zero — per-tet record and, with a rest-shape dictionary, the zero-volume twin of their record.
    // Then H = -|V0| P DmInv^T vanishes, the gradient guard (green_constraint.cpp:67, :130-131) returns before the
    // multiplier or a position changes, and the schedule (colours, regions, mailboxes, captured graphs) stays valid.
    bool remove_tets(sbsb200_ctx& c, std::vector<uint32_t> const& positions) override
    {
        if (positions.empty())
            return true;
        if (d.n_shapes > 0 && n_base_shapes == 0)
            return false; // a dictionary without room for the twins: the caller rebuilds the scene
        DevBuf<uint32_t> ids;
        ids.alloc(positions.size());
        CK(cudaMemcpyAsync(ids.p, positions.data(), sizeof(uint32_t) * positions.size(), cudaMemcpyHostToDevice, c.stream));
        int64_t const n = static_cast<int64_t>(positions.size());
        k_remove_tets<R><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(
            tet_r2.p, d.n_shapes > 0 ? tet_shape.p : nullptr, n_base_shapes, ids.p, n);
        ++c.kernels;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c.stream));
        return true;
    }

    void set_masses(sbsb200_ctx& c, int64_t first, int64_t n, uint32_t const* which, double const* m) override
    {
        if (n <= 0)
            return;
        std::vector<R> iw(static_cast<size_t>(n));
        for (int64_t i = 0; i < n; ++i)
            iw[static_cast<size_t>(i)] = R(m[i] > 0. ? 1. / m[i] : 0.); // particle.cpp:39-44
        DevBuf<uint32_t> ids;
        DevBuf<R> w;
        ids.alloc(static_cast<size_t>(n));
        w.alloc(static_cast<size_t>(n));
        CK(cudaMemcpyAsync(ids.p, which, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c.stream));
        CK(cudaMemcpyAsync(w.p, iw.data(), sizeof(R) * n, cudaMemcpyHostToDevice, c.stream));
        k_scatter_inverse_mass<R><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(d, first, n, ids.p, w.p);
        ++c.kernels;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c.stream));
    }

    // validation aid: vertices whose position or velocity is NaN / Inf
    int64_t count_non_finite(sbsb200_ctx& c) override
    {
        DevBuf<unsigned long long> n;
        n.alloc(1);
        CK(cudaMemsetAsync(n.p, 0, sizeof(unsigned long long), c.stream));
        if (d.n_vertices > 0)
        {
            k_count_non_finite<R><<<static_cast<unsigned>((d.n_vertices + 255) / 256), 256, 0, c.stream>>>(d, n.p);
            ++c.kernels;
        }
        unsigned long long h = 0;
        CK(cudaMemcpyAsync(&h, n.p, sizeof h, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
        return static_cast<int64_t>(h);
    }

    uint64_t general_route_calls() override
    {
        unsigned long long n = 0;
        CK(cudaMemcpyFromSymbol(&n, g_general_route_calls, sizeof n));
        return n;
    }

    // sbsb200_step_host_vertices_f32: only the listed vertices cross PCIe (a rank of a decomposed body round-trips
    // the vertices it owns)
    DevBuf<uint32_t> host_vertex_ids;
    void step_host_vertices_f32(sbsb200_ctx& c, int body, int64_t n, uint32_t const* which, float const* x, float const* v,
                                double dt, int substeps, int iterations, int detect, float* x_out, float* v_out) override
    {
        HostBody const& hb = c.scene.bodies[static_cast<size_t>(body)];
        ensure_staging(n);
        if (host_vertex_ids.n < static_cast<size_t>(n))
            host_vertex_ids.alloc(static_cast<size_t>(n));
        float* sxf = reinterpret_cast<float*>(stage_x.p);
        float* svf = reinterpret_cast<float*>(stage_v.p);
        unsigned const grid = static_cast<unsigned>((n + 255) / 256);
        if (n > 0)
        {
            CK(cudaMemcpyAsync(host_vertex_ids.p, which, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c.stream));
            if (x)
            {
                CK(cudaMemcpyAsync(sxf, x, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c.stream));
                if (v)
                    CK(cudaMemcpyAsync(svf, v, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c.stream));
                k_scatter_state<R, float><<<grid, 256, 0, c.stream>>>(d, hb.v_offset, n, host_vertex_ids.p, sxf,
                                                                     v ? svf : nullptr);
                ++c.kernels;
                if (d.n_surface > 0)
                {
                    k_surface_gather<R><<<static_cast<unsigned>((d.n_surface + 255) / 256), 256, 0, c.stream>>>(d);
                    ++c.kernels;
                }
            }
        }
        step(c, dt, substeps, iterations, detect);
        ++c.frames;
        if (n > 0 && (x_out || v_out))
        {
            k_gather_state<R, float><<<grid, 256, 0, c.stream>>>(d, hb.v_offset, n, host_vertex_ids.p,
                                                                x_out ? sxf : nullptr, v_out ? svf : nullptr);
            ++c.kernels;
            CK(cudaGetLastError());
            if (x_out)
                CK(cudaMemcpyAsync(x_out, sxf, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
            if (v_out)
                CK(cudaMemcpyAsync(v_out, svf, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
        }
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
    }

    // sbsb200_step_host_f32: the state crosses PCIe as floats (half the bytes of the double interface)
    void step_host_f32(sbsb200_ctx& c, int body, float const* x, float const* v, double dt, int substeps, int iterations,
                       int detect, float* x_out, float* v_out) override
    {
        // body < 0: every tet body, concatenated in body order (the global vertex order)
        int64_t const first = body < 0 ? 0 : c.scene.bodies[static_cast<size_t>(body)].v_offset;
        int64_t const n     = body < 0 ? d.n_vertices : c.scene.bodies[static_cast<size_t>(body)].n_vertices;
        struct
        {
            int64_t v_offset;
        } const hb{first};
        ensure_staging(n);
        float* sxf = reinterpret_cast<float*>(stage_x.p);
        float* svf = reinterpret_cast<float*>(stage_v.p);
        if (n > 0)
        {
            CK(cudaMemcpyAsync(sxf, x, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c.stream));
            if (v)
                CK(cudaMemcpyAsync(svf, v, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c.stream));
            k_unpack_state<R, float><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(d, hb.v_offset, n, sxf,
                                                                                                 v ? svf : nullptr);
            ++c.kernels;
            if (d.n_surface > 0)
            {
                k_surface_gather<R><<<static_cast<unsigned>((d.n_surface + 255) / 256), 256, 0, c.stream>>>(d);
                ++c.kernels;
            }
        }
        step(c, dt, substeps, iterations, detect);
        ++c.frames;
        if (n > 0 && (x_out || v_out))
        {
            k_pack_state<R, float><<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(
                d, hb.v_offset, n, x_out ? sxf : nullptr, v_out ? svf : nullptr);
            ++c.kernels;
            CK(cudaGetLastError());
            if (x_out)
                CK(cudaMemcpyAsync(x_out, sxf, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
            if (v_out)
                CK(cudaMemcpyAsync(v_out, svf, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c.stream));
        }
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
    }

    int64_t contacts(sbsb200_ctx& c, int64_t cap, int32_t* body, uint32_t* vertex, int32_t* sdf_body,
                     double* point, double* normal) override
    {
        uint32_t n = 0;
        CK(cudaMemcpyAsync(&n, contact_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        check_persistent(c);
        n                 = static_cast<uint32_t>(std::min<int64_t>(n, d.contact_cap));
        c.last_contacts   = n;
        int64_t const m   = std::min<int64_t>(n, cap);
        if (m <= 0 || !vertex)
            return n;
        std::vector<uint32_t> hv(static_cast<size_t>(m));
        std::vector<Real4<R>> hq(static_cast<size_t>(m)), hn(static_cast<size_t>(m));
        CK(cudaMemcpyAsync(hv.data(), contact_v.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaMemcpyAsync(hq.data(), contact_q.p, sizeof(Real4<R>) * m, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaMemcpyAsync(hn.data(), contact_n.p, sizeof(Real4<R>) * m, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        for (int64_t i = 0; i < m; ++i)
        {
            uint32_t const gv = hv[static_cast<size_t>(i)] & 0x7fffffffu;
            int32_t const b   = c.vertex_body[gv];
            if (body)
                body[i] = b;
            vertex[i] = static_cast<uint32_t>(gv - c.scene.bodies[static_cast<size_t>(b)].v_offset);
            if (sdf_body)
            {
                if (sizeof(R) == 4)
                {
                    float f = float(hn[static_cast<size_t>(i)].w);
                    int32_t k;
                    std::memcpy(&k, &f, 4);
                    sdf_body[i] = k;
                }
                else
                    sdf_body[i] = static_cast<int32_t>(hn[static_cast<size_t>(i)].w);
            }
            if (point)
            {
                point[3 * i]     = double(hq[static_cast<size_t>(i)].x);
                point[3 * i + 1] = double(hq[static_cast<size_t>(i)].y);
                point[3 * i + 2] = double(hq[static_cast<size_t>(i)].z);
            }
            if (normal)
            {
                normal[3 * i]     = double(hn[static_cast<size_t>(i)].x);
                normal[3 * i + 1] = double(hn[static_cast<size_t>(i)].y);
                normal[3 * i + 2] = double(hn[static_cast<size_t>(i)].z);
            }
        }
        return n;
    }
};

int fail(sbsb200_ctx* c, int code, std::string const& msg)
{
    if (c)
        c->err = msg;
    else
        g_create_error = msg;
    return code;
}

template <typename F>
int guarded(sbsb200_ctx* c, F&& f)
{
    try
    {
        return f();
    }
    catch (CudaError const& e)
    {
        return fail(c, SBSB200_ERR_CUDA, e.msg);
    }
    catch (std::exception const& e)
    {
        return fail(c, SBSB200_ERR_INVALID, e.what());
    }
}

bool is_tet_body(sbsb200_ctx const* c, int b)
{
    return b >= 0 && b < static_cast<int>(c->scene.bodies.size()) &&
           c->scene.bodies[static_cast<size_t>(b)].kind == BodyKind::tet;
}

int add_sdf(sbsb200_ctx* c, SdfKind kind, double const a[3], double const b[3], double r, double const volume[6])
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    if (!a || !volume)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    HostBody hb;
    hb.kind     = BodyKind::sdf;
    hb.sdf_kind = kind;
    std::memcpy(hb.a, a, sizeof hb.a);
    if (b)
        std::memcpy(hb.b, b, sizeof hb.b);
    hb.r = r;
    std::memcpy(hb.volume, volume, sizeof hb.volume);
    c->scene.bodies.push_back(hb);
    return static_cast<int>(c->scene.bodies.size()) - 1;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sbsb200_create(int device, int precision, sbsb200_ctx** out)
{
    if (!out || (precision != SBSB200_FP32 && precision != SBSB200_FP64))
        return fail(nullptr, SBSB200_ERR_INVALID, "bad arguments to sbsb200_create");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count)
        return fail(nullptr, SBSB200_ERR_CUDA,
                    std::string("no usable CUDA device (there is no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    cudaDeviceProp prop{};
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, SBSB200_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SBSB200_ERR_CUDA,
                    "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                        "; this library contains sm_100a code only");
    auto* c      = new sbsb200_ctx();
    c->device    = device;
    c->precision = precision;
    c->sm_count  = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess)
    {
        delete c;
        return fail(nullptr, SBSB200_ERR_CUDA, "stream/event creation failed");
    }
    c->own_stream = true;
    *out          = c;
    return SBSB200_OK;
}

void sbsb200_destroy(sbsb200_ctx* c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    for (void* p : c->ipc_opened)
        cudaIpcCloseMemHandle(p);
    c->engine.reset();
    if (c->ev0)
        cudaEventDestroy(c->ev0);
    if (c->ev1)
        cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
}

const char* sbsb200_last_error(const sbsb200_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int sbsb200_set_stream(sbsb200_ctx* c, void* cuda_stream)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "set_stream must precede finalize");
    if (c->own_stream && c->stream)
        cudaStreamDestroy(c->stream);
    c->stream     = static_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return SBSB200_OK;
}

int sbsb200_set_schedule(sbsb200_ctx* c, int schedule)
{
    if (!c || schedule < SBSB200_SCHED_AUTO || schedule > SBSB200_SCHED_PERSISTENT)
        return fail(c, SBSB200_ERR_INVALID, "bad schedule");
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "set_schedule must precede finalize");
    c->schedule_request = schedule;
    return SBSB200_OK;
}

int sbsb200_set_broadphase(sbsb200_ctx* c, int mode)
{
    if (!c || (mode != SBSB200_BROADPHASE_NONE && mode != SBSB200_BROADPHASE_BVH))
        return fail(c, SBSB200_ERR_INVALID, "bad broadphase mode");
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "set_broadphase must precede finalize");
    c->broadphase = mode;
    return SBSB200_OK;
}

int sbsb200_set_region_shape(sbsb200_ctx* c, int shape)
{
    if (!c || (shape != SBSB200_REGIONS_PENCILS && shape != SBSB200_REGIONS_COMPACT))
        return fail(c, SBSB200_ERR_INVALID, "bad region shape");
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    c->region_shape = shape;
    return SBSB200_OK;
}

int sbsb200_debug_trace_steps(sbsb200_ctx* c, int steps)
{
    if (!c || steps < 0)
        return fail(c, SBSB200_ERR_INVALID, "bad argument");
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    c->trace_steps = steps;
    return SBSB200_OK;
}

int sbsb200_set_collision_compliance(sbsb200_ctx* c, double alpha)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    c->collision_alpha = alpha;
    if (c->engine)
        c->engine->invalidate_graphs();
    return SBSB200_OK;
}

int sbsb200_add_tet_body(sbsb200_ctx* c, int64_t nV, const double* x0, const double* mass, int64_t nT,
                         const uint32_t* tets, double E, double nu, double alpha, double beta)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    if (nV < 0 || nT < 0 || (nV > 0 && !x0) || (nT > 0 && !tets))
        return fail(c, SBSB200_ERR_INVALID, "null or negative-sized geometry");
    HostScene& h = c->scene;
    if (h.n_vertices() + nV >= (int64_t{1} << 31) || static_cast<int64_t>(h.n_constraints) + nT >= (int64_t{1} << 32))
        return fail(c, SBSB200_ERR_CAPACITY, "scene too large for 32-bit indices");
    for (int64_t i = 0; i < 4 * nT; ++i)
        if (tets[i] >= static_cast<uint64_t>(nV))
            return fail(c, SBSB200_ERR_INVALID, "tet vertex index out of range");
    HostBody hb;
    hb.kind       = BodyKind::tet;
    hb.v_offset   = h.n_vertices();
    hb.n_vertices = nV;
    hb.t_offset   = h.n_tets();
    hb.n_tets     = nT;
    extract_boundary(nV, nT, tets, hb.surf_to_tet, &hb.surf_triangles);
    h.x0.insert(h.x0.end(), x0, x0 + 3 * nV);
    if (mass)
        h.mass.insert(h.mass.end(), mass, mass + nV);
    else
        h.mass.insert(h.mass.end(), static_cast<size_t>(nV), 1.0); // particle.cpp:7
    Material const m{E / (2. * (1 + nu)), (E * nu) / ((1 + nu) * (1 - 2 * nu)), alpha, beta}; // :45-46
    size_t mi = 0;
    while (mi < h.materials.size() && !(h.materials[mi] == m))
        ++mi;
    if (mi == h.materials.size())
    {
        if (mi >= 65535)
            return fail(c, SBSB200_ERR_CAPACITY, "too many distinct materials");
        h.materials.push_back(m);
    }
    for (int64_t t = 0; t < nT; ++t)
    {
        for (int a = 0; a < 4; ++a)
            h.tets.push_back(static_cast<uint32_t>(hb.v_offset + tets[4 * t + a]));
        h.tet_insertion.push_back(h.n_constraints++);
        h.tet_material.push_back(static_cast<uint16_t>(mi));
    }
    h.bodies.push_back(std::move(hb));
    return static_cast<int>(h.bodies.size()) - 1;
}

int sbsb200_add_distance_constraints(sbsb200_ctx* c, int b1, int b2, int64_t n, const uint32_t* pairs,
                                     double alpha, double beta)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    if (!is_tet_body(c, b1) || !is_tet_body(c, b2) || n < 0 || (n > 0 && !pairs))
        return fail(c, SBSB200_ERR_INVALID, "bad body index or null pairs");
    HostScene& h        = c->scene;
    HostBody const& hb1 = h.bodies[static_cast<size_t>(b1)];
    HostBody const& hb2 = h.bodies[static_cast<size_t>(b2)];
    for (int64_t i = 0; i < n; ++i)
        if (pairs[2 * i] >= static_cast<uint64_t>(hb1.n_vertices) || pairs[2 * i + 1] >= static_cast<uint64_t>(hb2.n_vertices))
            return fail(c, SBSB200_ERR_INVALID, "distance constraint vertex out of range");
    for (int64_t i = 0; i < n; ++i)
    {
        uint32_t const g1 = static_cast<uint32_t>(hb1.v_offset + pairs[2 * i]);
        uint32_t const g2 = static_cast<uint32_t>(hb2.v_offset + pairs[2 * i + 1]);
        double s          = 0;
        for (int k = 0; k < 3; ++k)
        {
            double const dlt = h.x0[3 * static_cast<size_t>(g1) + k] - h.x0[3 * static_cast<size_t>(g2) + k];
            s += dlt * dlt;
        }
        h.dist_pairs.push_back(g1);
        h.dist_pairs.push_back(g2);
        h.dist_rest.push_back(std::sqrt(s)); // distance_constraint.cpp:21
        h.dist_alpha.push_back(alpha);
        h.dist_beta.push_back(beta);
        h.dist_insertion.push_back(h.n_constraints++);
    }
    return SBSB200_OK;
}

int sbsb200_set_body_collideable(sbsb200_ctx* c, int body, int flag)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    if (body < 0 || body >= static_cast<int>(c->scene.bodies.size()))
        return fail(c, SBSB200_ERR_INVALID, "no such body");
    c->scene.bodies[static_cast<size_t>(body)].collideable = flag != 0;
    return SBSB200_OK;
}

int sbsb200_add_sdf_plane(sbsb200_ctx* c, const double n[3], const double pt[3], const double volume[6])
{
    if (!n || !pt)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    // Eigen::Hyperplane(n, e): offset = -n.e (sdf_model.cpp:56-61 uses signedDistance / normal)
    return add_sdf(c, SdfKind::plane, n, nullptr, -(n[0] * pt[0] + n[1] * pt[1] + n[2] * pt[2]), volume);
}
int sbsb200_add_sdf_sphere(sbsb200_ctx* c, const double centre[3], double radius, const double volume[6])
{
    return add_sdf(c, SdfKind::sphere, centre, nullptr, radius, volume);
}
int sbsb200_add_sdf_box(sbsb200_ctx* c, const double bmin[3], const double bmax[3], const double volume[6])
{
    if (!bmax)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    return add_sdf(c, SdfKind::box, bmin, bmax, 0., volume);
}

int64_t sbsb200_grid_node_count(const uint32_t res[3])
{
    if (!res || !res[0] || !res[1] || !res[2])
        return SBSB200_ERR_INVALID;
    return static_cast<int64_t>(GridDims{{res[0], res[1], res[2]}}.nodes());
}

int sbsb200_grid_node_position(const double dmin[3], const double dmax[3], const uint32_t res[3], int64_t node,
                               double position[3])
{
    if (!dmin || !dmax || !position || node < 0 || node >= sbsb200_grid_node_count(res))
        return SBSB200_ERR_INVALID;
    grid_node_position(GridDims{{res[0], res[1], res[2]}}, dmin, dmax, static_cast<uint64_t>(node), position);
    return SBSB200_OK;
}

int sbsb200_add_sdf_grid(sbsb200_ctx* c, const double dmin[3], const double dmax[3], const uint32_t res[3],
                         const double* nodes, int64_t n_nodes, const double volume[6])
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!dmin || !dmax || !res || !nodes)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!res[0] || !res[1] || !res[2] || !(dmin[0] < dmax[0] && dmin[1] < dmax[1] && dmin[2] < dmax[2]))
        return fail(c, SBSB200_ERR_INVALID, "empty grid");
    if (n_nodes != sbsb200_grid_node_count(res))
        return fail(c, SBSB200_ERR_INVALID, "node count does not match the resolution (sbsb200_grid_node_count)");
    double const dom[6] = {dmin[0], dmin[1], dmin[2], dmax[0], dmax[1], dmax[2]};
    int const id        = add_sdf(c, SdfKind::grid, dmin, dmax, 0., volume ? volume : dom);
    if (id < 0)
        return id;
    HostBody& hb = c->scene.bodies[static_cast<size_t>(id)];
    std::memcpy(hb.grid_n, res, sizeof hb.grid_n);
    hb.grid_nodes.assign(nodes, nodes + n_nodes);
    return id;
}

int sbsb200_mesh_sdf_domain(int64_t nV, const double* x, const double domain[6], double dom[6])
{
    if (!x || !domain || !dom || nV < 0)
        return SBSB200_ERR_INVALID;
    // environment_body.cpp:52-65: the box is extended to the mesh and then inflated by 1e-3 of its
    // diagonal, max first, then min with the diagonal of the already grown box — once per mesh
    // vertex, because the growth statements sit inside the outer vertex loop
    std::memcpy(dom, domain, 6 * sizeof(double));
    for (int64_t v = 0; v < nV; ++v)
        for (int k = 0; k < 3; ++k)
        {
            dom[k]     = std::min(dom[k], x[3 * v + k]);
            dom[3 + k] = std::max(dom[3 + k], x[3 * v + k]);
        }
    for (int64_t v = 0; v < nV; ++v)
        for (int side = 1; side >= 0; --side)
        {
            double const dx = dom[3] - dom[0], dy = dom[4] - dom[1], dz = dom[5] - dom[2];
            double const g  = 1.0e-3 * std::sqrt(dx * dx + dy * dy + dz * dz);
            for (int k = 0; k < 3; ++k)
                dom[3 * side + k] += side ? g : -g;
        }
    return SBSB200_OK;
}

int sbsb200_add_sdf_mesh(sbsb200_ctx* c, int64_t nV, const double* x, int64_t nF, const uint32_t* tri,
                         const double domain[6], const uint32_t resolution[3])
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    if (!x || !tri || !domain || nV <= 0 || nF <= 0)
        return fail(c, SBSB200_ERR_INVALID, "null or empty mesh");
    uint32_t const dflt[3] = {10u, 10u, 10u}; // environment_body.h:24
    uint32_t const* res    = resolution ? resolution : dflt;
    if (!res[0] || !res[1] || !res[2])
        return fail(c, SBSB200_ERR_INVALID, "empty grid");
    for (int64_t i = 0; i < 3 * nF; ++i)
        if (tri[i] >= nV)
            return fail(c, SBSB200_ERR_INVALID, "triangle index out of range");
    return guarded(c, [&]() -> int {
        double dom[6];
        sbsb200_mesh_sdf_domain(nV, x, domain, dom);
        // per-triangle records with the pseudo-normals of faces, edges and corners (MeshDistance ctor)
        std::vector<BakeTriangle> recs(static_cast<size_t>(nF));
        std::vector<std::array<double, 3>> vn(static_cast<size_t>(nV), {0., 0., 0.});
        std::map<std::pair<uint32_t, uint32_t>, int64_t> half_edge; // (from, to) -> face
        for (int64_t f = 0; f < nF; ++f)
        {
            BakeTriangle& t = recs[static_cast<size_t>(f)];
            for (int k = 0; k < 3; ++k)
            {
                for (int a = 0; a < 3; ++a)
                    t.p[k][a] = x[3 * tri[3 * f + k] + a];
                half_edge[{tri[3 * f + k], tri[3 * f + (k + 1) % 3]}] = f;
            }
            double e[3][3], len[3];
            for (int k = 0; k < 3; ++k)
            {
                for (int a = 0; a < 3; ++a)
                    e[k][a] = t.p[(k + 1) % 3][a] - t.p[k][a];
                len[k] = std::sqrt(e[k][0] * e[k][0] + e[k][1] * e[k][1] + e[k][2] * e[k][2]);
            }
            double const ac[3] = {-e[2][0], -e[2][1], -e[2][2]};
            double n[3] = {e[0][1] * ac[2] - e[0][2] * ac[1], e[0][2] * ac[0] - e[0][0] * ac[2],
                           e[0][0] * ac[1] - e[0][1] * ac[0]};
            double const ln = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (!(ln > 0.) || !(len[0] > 0.) || !(len[1] > 0.) || !(len[2] > 0.))
                throw std::invalid_argument("degenerate triangle in the obstacle mesh");
            for (int a = 0; a < 3; ++a)
                t.fn[a] = n[a] / ln;
            for (int k = 0; k < 3; ++k)
            { // interior angle at corner k
                int const pk = (k + 2) % 3;
                double cs = -(e[k][0] * e[pk][0] + e[k][1] * e[pk][1] + e[k][2] * e[pk][2]) / (len[k] * len[pk]);
                cs        = std::min(1., std::max(-1., cs));
                double const al = std::acos(cs);
                for (int a = 0; a < 3; ++a)
                    vn[tri[3 * f + k]][static_cast<size_t>(a)] += al * t.fn[a];
            }
        }
        for (int64_t f = 0; f < nF; ++f)
        {
            BakeTriangle& t = recs[static_cast<size_t>(f)];
            for (int k = 0; k < 3; ++k)
            {
                auto const opp = half_edge.find({tri[3 * f + (k + 1) % 3], tri[3 * f + k]});
                for (int a = 0; a < 3; ++a)
                {
                    t.en[k][a] = t.fn[a] + (opp != half_edge.end() && opp->second != f
                                                ? recs[static_cast<size_t>(opp->second)].fn[a]
                                                : 0.);
                    t.vn[k][a] = vn[tri[3 * f + k]][static_cast<size_t>(a)];
                }
            }
        }
        GridDims const g{{res[0], res[1], res[2]}};
        int64_t const nn = static_cast<int64_t>(g.nodes());
        CK(cudaSetDevice(c->device));
        DevBuf<BakeTriangle> d_tris;
        DevBuf<double> d_nodes;
        d_tris.upload(recs, c->stream);
        d_nodes.alloc(static_cast<size_t>(nn));
        k_bake_mesh_sdf<<<static_cast<unsigned>((nn + 127) / 128), 128, 0, c->stream>>>(
            g, dom[0], dom[1], dom[2], dom[3], dom[4], dom[5], d_tris.p, nF, d_nodes.p);
        ++c->kernels;
        CK(cudaGetLastError());
        std::vector<double> nodes(static_cast<size_t>(nn));
        CK(cudaMemcpyAsync(nodes.data(), d_nodes.p, sizeof(double) * nodes.size(), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        // environment_body.cpp:75-77: the model's volume() is the extended domain
        return sbsb200_add_sdf_grid(c, dom, dom + 3, res, nodes.data(), nn, dom);
    });
}

int64_t sbsb200_get_sdf_grid(sbsb200_ctx* c, int body, double domain[6], uint32_t resolution[3], double* nodes,
                             int64_t cap)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (body < 0 || body >= static_cast<int>(c->scene.bodies.size()) ||
        c->scene.bodies[static_cast<size_t>(body)].kind != BodyKind::sdf ||
        c->scene.bodies[static_cast<size_t>(body)].sdf_kind != SdfKind::grid)
        return fail(c, SBSB200_ERR_INVALID, "not a grid sdf body");
    HostBody const& hb = c->scene.bodies[static_cast<size_t>(body)];
    if (domain)
        for (int k = 0; k < 3; ++k)
        {
            domain[k]     = hb.a[k];
            domain[3 + k] = hb.b[k];
        }
    if (resolution)
        std::memcpy(resolution, hb.grid_n, sizeof hb.grid_n);
    if (nodes)
        std::memcpy(nodes, hb.grid_nodes.data(),
                    sizeof(double) * static_cast<size_t>(std::min<int64_t>(cap, static_cast<int64_t>(hb.grid_nodes.size()))));
    return static_cast<int64_t>(hb.grid_nodes.size());
}

int sbsb200_eval_sdf(sbsb200_ctx* c, int body, int64_t n, const double* points, double* distance, double* gradient)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "eval_sdf before finalize");
    if (body < 0 || body >= static_cast<int>(c->scene.bodies.size()) ||
        c->scene.bodies[static_cast<size_t>(body)].kind != BodyKind::sdf)
        return fail(c, SBSB200_ERR_INVALID, "not an sdf body");
    if (n < 0 || (n > 0 && (!points || !distance || !gradient)))
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (n == 0)
        return SBSB200_OK;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->eval_sdf(*c, body, n, points, distance, gradient);
        return SBSB200_OK;
    });
}

int sbsb200_finalize(sbsb200_ctx* c)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "scene already finalized");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        HostScene& h    = c->scene;
        int64_t const V = h.n_vertices(), T = h.n_tets(), D = h.n_dist();
        c->vertex_body.assign(static_cast<size_t>(V), -1);
        int64_t s_off = 0;
        for (size_t b = 0; b < h.bodies.size(); ++b)
            if (h.bodies[b].kind == BodyKind::tet)
            {
                h.bodies[b].s_offset = s_off;
                s_off += static_cast<int64_t>(h.bodies[b].surf_to_tet.size());
                for (int64_t i = 0; i < h.bodies[b].n_vertices; ++i)
                    c->vertex_body[static_cast<size_t>(h.bodies[b].v_offset + i)] = static_cast<int32_t>(b);
            }
        c->any_damping = false;
        for (Material const& m : h.materials)
            c->any_damping = c->any_damping || m.beta != 0.;

        std::vector<uint64_t> dkeys;
        morton_keys(D, 2, h.dist_pairs.data(), h.x0.data(), V, dkeys);

        // schedule choice: the persistent kernel covers Green + collision constraints without
        // damping; distance constraints and beta != 0 (which needs xn in the projection) take
        // the per-colour kernels
        bool const persistent_ok = T > 0 && D == 0 && !c->any_damping;
        // AUTO: the resident kernel whenever it applies (connected meshes cut into one region per SM, ensembles
        // with a few whole bodies per region); the per-colour kernels in a CUDA graph otherwise
        bool const auto_persistent =
            persistent_ok;
        c->schedule = c->schedule_request == SBSB200_SCHED_PERSISTENT                     ? SBSB200_SCHED_PERSISTENT
                      : c->schedule_request == SBSB200_SCHED_AUTO && auto_persistent ? SBSB200_SCHED_PERSISTENT
                                                                                     : SBSB200_SCHED_GRAPH;
        if (c->world > 1)
        { // a decomposed scene exists only in the resident schedule (mailboxes in peer memory)
            if (!persistent_ok || c->schedule_request == SBSB200_SCHED_GRAPH)
                return fail(c, SBSB200_ERR_INVALID,
                            "a partitioned scene needs the persistent schedule (tets, no distance constraints, beta = 0)");
            c->schedule = SBSB200_SCHED_PERSISTENT;
        }
        if (c->schedule == SBSB200_SCHED_PERSISTENT && !persistent_ok)
        {
            c->schedule      = SBSB200_SCHED_GRAPH;
            c->schedule_note = "persistent schedule needs tets, no distance constraints and beta = 0";
        }

        if (c->schedule == SBSB200_SCHED_PERSISTENT)
        { // one region per SM; when the vertices a region touches do not fit its shared memory, two smaller
          // regions per SM (half the threads each); beyond that the per-colour kernels take over
            bool const ensemble = ResidentPlan<float>::wants_region_per_body(h, c->sm_count);
            int64_t const vertex_bytes = c->precision == SBSB200_FP32 ? 16 : 32;
            bool planned = false;
            std::string why;
            // attempts: pencil-shaped regions (unless compact ones were asked for), compact regions (fewer vertices
            // per region), two smaller compact regions per SM
            struct Attempt
            {
                bool pencils;
                int per_sm;
            };
            std::vector<Attempt> attempts;
            if (c->region_shape == SBSB200_REGIONS_PENCILS)
                attempts.push_back({true, 1});
            attempts.push_back({false, 1});
            attempts.push_back({false, 2});
            for (Attempt const& at : attempts)
            {
                if (planned)
                    break;
                int const per_sm  = at.per_sm;
                ResidentParams rp = c->precision == SBSB200_FP32 ? ResidentPlan<float>::resident_params()
                                                                 : ResidentPlan<double>::resident_params();
                rp.pencils     = at.pencils;
                if (ensemble)
                { // bodies per region: chosen by build_cluster_plan for the least idle capacity of the SMs (few bodies
                  // per SM come out as one body per region, so that every SM has independent regions to interleave)
                    rp.bodies_per_region = 0;
                    rp.sm_count          = c->sm_count;
                }
                rp.smem_bytes  = rp.smem_bytes / per_sm - (per_sm > 1 ? 4096 : 0);
                rp.max_threads = per_sm == 1 ? 384 : 192;
                int32_t n_regions = regions_for(c->sm_count, T, c->world);
                if (per_sm == 2)
                {
                    if (ensemble || n_regions != c->sm_count * c->world)
                        break; // smaller scenes already have the regions they need
                    n_regions *= 2;
                }
                build_cluster_plan(h, n_regions, ensemble, c->green_plan, &rp, &c->plan);
                if (!build_exchange_plan(h, c->green_plan, c->plan, c->world, c->xplan))
                    why = c->xplan.why_not;
                else if (c->xplan.max_local * vertex_bytes > rp.smem_bytes)
                    why = "the vertices a region touches do not fit shared memory";
                else
                    planned = true;
            }
            if (!planned)
            {
                if (c->world > 1)
                    return fail(c, SBSB200_ERR_CAPACITY, "partitioned scene: " + why);
                c->schedule      = SBSB200_SCHED_GRAPH;
                c->schedule_note = "resident schedule unavailable: " + why;
                c->plan          = RegionPlan{};
                c->xplan         = ExchangePlan{};
                build_cluster_plan(h, 1, false, c->green_plan);
            }
        }
        else
            build_cluster_plan(h, 1, false, c->green_plan);
        if (!cluster_plan_is_valid(h, c->green_plan))
            return fail(c, SBSB200_ERR_CAPACITY, "clustered colouring failed (more than 128 colours needed?)");
        if (!colour_constraints(V, D, 2, h.dist_pairs.data(), dkeys.data(), nullptr, 256, c->dist_cc))
            return fail(c, SBSB200_ERR_CAPACITY, "more than 256 colours needed for the distance constraints");

        // exported serial order: green colours, then distance colours
        c->order.clear();
        c->order.reserve(static_cast<size_t>(T + D));
        for (uint32_t t : c->green_plan.serial_order)
            c->order.push_back(h.tet_insertion[t]);
        for (uint32_t i : c->dist_cc.order)
            c->order.push_back(h.dist_insertion[i]);

        if (c->precision == SBSB200_FP32)
            c->engine.reset(new Engine<float>());
        else
            c->engine.reset(new Engine<double>());
        c->engine->build(*c);
        c->general_calls0 = c->engine->general_route_calls();
        if (c->world > 1 && c->schedule != SBSB200_SCHED_PERSISTENT)
            return fail(c, SBSB200_ERR_CAPACITY, "partitioned scene: " + c->schedule_note);
        c->finalized = true;
        return SBSB200_OK;
    });
}

int64_t sbsb200_constraint_count(const sbsb200_ctx* c)
{
    return c ? static_cast<int64_t>(c->scene.n_constraints) - c->n_removed : static_cast<int64_t>(SBSB200_ERR_INVALID);
}

int sbsb200_remove_constraints(sbsb200_ctx* c, int64_t n, const uint32_t* constraints)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "sbsb200_remove_constraints: call after sbsb200_finalize");
    if (n < 0 || (n > 0 && !constraints))
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    HostScene const& h = c->scene;
    int64_t const T = h.n_tets(), N = h.n_constraints;
    if (c->position_of_insertion.empty() && T > 0)
    { // insertion index -> tet -> storage position; 0xffffffff: not a tet constraint
        std::vector<uint32_t> position_of_tet(static_cast<size_t>(T), 0);
        for (int64_t p = 0; p < T; ++p)
            position_of_tet[c->green_plan.storage_order[static_cast<size_t>(p)]] = static_cast<uint32_t>(p);
        c->position_of_insertion.assign(static_cast<size_t>(N), 0xffffffffu);
        for (int64_t t = 0; t < T; ++t)
            c->position_of_insertion[h.tet_insertion[static_cast<size_t>(t)]] = position_of_tet[static_cast<size_t>(t)];
        c->removed.assign(static_cast<size_t>(N), 0);
    }
    std::vector<uint32_t> positions;
    std::vector<char> seen(c->removed);
    for (int64_t i = 0; i < n; ++i)
    {
        if (constraints[i] >= static_cast<uint64_t>(N) || c->position_of_insertion.empty() ||
            c->position_of_insertion[constraints[i]] == 0xffffffffu)
            return fail(c, SBSB200_ERR_INVALID,
                        "sbsb200_remove_constraints: not a tetrahedron constraint of this scene (distance constraints "
                        "need a rebuild)");
        if (seen[constraints[i]])
            return fail(c, SBSB200_ERR_INVALID, "sbsb200_remove_constraints: constraint removed twice");
        seen[constraints[i]] = 1;
        positions.push_back(c->position_of_insertion[constraints[i]]);
    }
    if (n == 0)
        return SBSB200_OK;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        if (!c->engine->remove_tets(*c, positions))
            return fail(c, SBSB200_ERR_CAPACITY,
                        "sbsb200_remove_constraints: the rest-shape dictionary has no room for zero-volume records; "
                        "rebuild the scene without the constraints");
        c->removed.swap(seen);
        c->n_removed += n;
        c->order.erase(std::remove_if(c->order.begin(), c->order.end(),
                                      [&](uint32_t id) { return id < c->removed.size() && c->removed[id]; }),
                       c->order.end());
        return SBSB200_OK;
    });
}

int sbsb200_get_constraint_order(const sbsb200_ctx* c, uint32_t* order, int64_t n)
{
    if (!c || !order)
        return SBSB200_ERR_INVALID;
    if (!c->finalized || n != static_cast<int64_t>(c->order.size()))
        return SBSB200_ERR_STATE;
    std::memcpy(order, c->order.data(), sizeof(uint32_t) * c->order.size());
    return SBSB200_OK;
}

int64_t sbsb200_get_surface_map(const sbsb200_ctx* c, int body, uint32_t* map, int64_t cap)
{
    if (!c || !is_tet_body(c, body))
        return SBSB200_ERR_INVALID;
    auto const& m = c->scene.bodies[static_cast<size_t>(body)].surf_to_tet;
    if (map)
        std::memcpy(map, m.data(), sizeof(uint32_t) * static_cast<size_t>(std::min<int64_t>(cap, static_cast<int64_t>(m.size()))));
    return static_cast<int64_t>(m.size());
}

int64_t sbsb200_get_surface_triangles(const sbsb200_ctx* c, int body, uint32_t* triangles, int64_t cap)
{
    if (!c || !is_tet_body(c, body))
        return SBSB200_ERR_INVALID;
    auto const& t = c->scene.bodies[static_cast<size_t>(body)].surf_triangles;
    if (triangles)
        std::memcpy(triangles, t.data(),
                    sizeof(uint32_t) * static_cast<size_t>(std::min<int64_t>(cap, static_cast<int64_t>(t.size()))));
    return static_cast<int64_t>(t.size());
}

int sbsb200_download_surface(sbsb200_ctx* c, int body, float* out)
{
    if (!c || !out)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "download_surface before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->download_surface(*c, body, out);
        return SBSB200_OK;
    });
}

int sbsb200_download_surface_rgb(sbsb200_ctx* c, int body, const float* colours, int64_t n_colours, float* out)
{
    if (!c || !out || !colours)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "download_surface_rgb before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    int64_t const n = static_cast<int64_t>(c->scene.bodies[static_cast<size_t>(body)].surf_to_tet.size());
    if (n_colours != 1 && n_colours != n)
        return fail(c, SBSB200_ERR_INVALID, "download_surface_rgb: one colour, or one per surface vertex");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->download_surface(*c, body, out, colours, n_colours);
        return SBSB200_OK;
    });
}

int sbsb200_get_stats(const sbsb200_ctx* cc, sbsb200_stats* out)
{
    if (!cc || !out)
        return SBSB200_ERR_INVALID;
    auto* c = const_cast<sbsb200_ctx*>(cc);
    std::memset(out, 0, sizeof *out);
    out->n_bodies = static_cast<int32_t>(c->scene.bodies.size());
    for (auto const& b : c->scene.bodies)
        out->n_sdfs += b.kind == BodyKind::sdf;
    out->n_vertices           = c->scene.n_vertices();
    out->n_tets               = c->scene.n_tets();
    out->n_distance           = c->scene.n_dist();
    out->n_surface_vertices   = c->n_surface;
    out->n_green_colours      = c->green_plan.n_colours;
    out->n_distance_colours   = c->dist_cc.n_colours;
    out->schedule             = c->schedule;
    out->n_regions            = c->plan.n_regions;
    out->n_interface_vertices = c->plan.n_interface;
    out->kernels_launched     = c->kernels;
    out->frames               = c->frames;
    out->last_contact_count   = c->last_contacts;
    out->n_shared_vertices    = c->xplan.n_shared;
    out->pulls_per_sweep      = c->xplan.n_pulls[2];
    out->pushes_per_sweep     = c->xplan.n_pushes[0];
    for (int64_t n : c->xplan.pulls_by_colour)
        out->quiet_colours += c->schedule == SBSB200_SCHED_PERSISTENT && 50 * n <= c->xplan.n_pulls[2];
    if (c->engine)
    {
        cudaSetDevice(c->device);
        c->engine->kernel_times(out->kernel_ms, out->kernel_launches);
        try
        {
            out->green_general_calls = static_cast<int64_t>(c->engine->general_route_calls() - c->general_calls0);
        }
        catch (...)
        {
            out->green_general_calls = -1;
        }
    }
    if (c->timed)
    {
        cudaSetDevice(c->device);
        float ms = 0.f;
        if (cudaEventSynchronize(c->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess)
            out->last_step_ms = ms;
    }
    return SBSB200_OK;
}

const char* sbsb200_schedule_note(const sbsb200_ctx* c) { return c ? c->schedule_note.c_str() : ""; }

int sbsb200_upload(sbsb200_ctx* c, int body, const double* x, const double* v)
{
    if (!c || !x)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "upload before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->upload(*c, body, x, v);
        return SBSB200_OK;
    });
}

int sbsb200_set_vertices(sbsb200_ctx* c, int body, int64_t n, const uint32_t* vertices, const double* x,
                         const double* v)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "set_vertices before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    if (n < 0 || (n > 0 && (!vertices || !x)))
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    for (int64_t i = 0; i < n; ++i)
        if (vertices[i] >= c->scene.bodies[static_cast<size_t>(body)].n_vertices)
            return fail(c, SBSB200_ERR_INVALID, "vertex index out of range");
    if (n == 0)
        return SBSB200_OK;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->set_vertices(*c, body, n, vertices, x, v);
        return SBSB200_OK;
    });
}

int sbsb200_download(sbsb200_ctx* c, int body, double* x, double* v)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "download before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->download(*c, body, x, v);
        return SBSB200_OK;
    });
}

int sbsb200_set_masses(sbsb200_ctx* c, int body, int64_t n, const uint32_t* vertices, const double* masses)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!is_tet_body(c, body) || n < 0 || (n > 0 && (!vertices || !masses)))
        return fail(c, SBSB200_ERR_INVALID, "bad body or arguments");
    HostBody const& hb = c->scene.bodies[static_cast<size_t>(body)];
    for (int64_t i = 0; i < n; ++i)
        if (vertices[i] >= static_cast<uint64_t>(hb.n_vertices))
            return fail(c, SBSB200_ERR_INVALID, "vertex index out of range");
    for (int64_t i = 0; i < n; ++i)
        c->scene.mass[static_cast<size_t>(hb.v_offset + vertices[i])] = masses[i];
    if (!c->finalized || n == 0)
        return SBSB200_OK;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->set_masses(*c, hb.v_offset, n, vertices, masses);
        return SBSB200_OK;
    });
}

int sbsb200_set_mass(sbsb200_ctx* c, int body, int64_t vertex, double mass)
{
    if (vertex < 0 || vertex > 0xffffffffll)
        return c ? fail(c, SBSB200_ERR_INVALID, "bad body or vertex") : SBSB200_ERR_INVALID;
    uint32_t const v = static_cast<uint32_t>(vertex);
    return sbsb200_set_masses(c, body, 1, &v, &mass);
}

int sbsb200_set_partition(sbsb200_ctx* c, int rank, int world)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (c->finalized)
        return fail(c, SBSB200_ERR_STATE, "set_partition must precede finalize");
    if (world < 1 || world > 8 || rank < 0 || rank >= world)
        return fail(c, SBSB200_ERR_INVALID, "bad rank/world (at most 8 ranks)");
    c->rank  = rank;
    c->world = world;
    return SBSB200_OK;
}

int sbsb200_get_mailbox_handle(sbsb200_ctx* c, void* handle64)
{
    if (!c || !handle64)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "get_mailbox_handle before finalize");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        size_t bytes = 0;
        void* p      = c->engine->mailbox_pointer(bytes);
        if (!p)
            return fail(c, SBSB200_ERR_STATE, "this scene has no mailboxes (not the persistent schedule)");
        cudaIpcMemHandle_t hnd;
        CK(cudaIpcGetMemHandle(&hnd, p));
        std::memcpy(handle64, &hnd, 64);
        return SBSB200_OK;
    });
}

int sbsb200_connect_peers(sbsb200_ctx* c, const void* handles, int world)
{
    if (!c || !handles)
        return SBSB200_ERR_INVALID;
    if (!c->finalized || world != c->world)
        return fail(c, SBSB200_ERR_STATE, "connect_peers: finalize first, with the same world size");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        for (int r = 0; r < world; ++r)
        {
            if (r == c->rank)
                continue;
            cudaIpcMemHandle_t hnd;
            std::memcpy(&hnd, static_cast<char const*>(handles) + 64 * r, 64);
            void* p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(p);
            c->engine->set_peer_mailboxes(r, p);
        }
        return SBSB200_OK;
    });
}

int sbsb200_connect_peer_context(sbsb200_ctx* c, int peer_rank, sbsb200_ctx* peer)
{
    if (!c || !peer || !c->finalized || !peer->finalized || peer_rank < 0 || peer_rank >= c->world || peer_rank == c->rank)
        return fail(c, SBSB200_ERR_INVALID, "connect_peer_context: both contexts finalized, peer_rank another rank");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        if (peer->device != c->device)
        {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, peer->device));
            if (!can)
                return fail(c, SBSB200_ERR_CUDA, "no peer access between the two devices");
            cudaError_t const e = cudaDeviceEnablePeerAccess(peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                throw CudaError{std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)};
            cudaGetLastError();
        }
        size_t bytes = 0;
        void* p      = peer->engine->mailbox_pointer(bytes);
        if (!p)
            return fail(c, SBSB200_ERR_STATE, "the peer has no mailboxes");
        c->engine->set_peer_mailboxes(peer_rank, p);
        return SBSB200_OK;
    });
}

int sbsb200_get_vertex_ranks(const sbsb200_ctx* c, int body, int32_t* out, int64_t n)
{
    if (!c || !out || !c->finalized || !is_tet_body(c, body))
        return SBSB200_ERR_INVALID;
    HostBody const& hb = c->scene.bodies[static_cast<size_t>(body)];
    if (n != hb.n_vertices)
        return SBSB200_ERR_INVALID;
    int32_t const per_rank = std::max<int32_t>(1, c->plan.n_regions / std::max(1, c->world));
    for (int64_t i = 0; i < n; ++i)
        out[i] = c->world > 1 && !c->plan.vertex_owner.empty()
                     ? c->plan.vertex_owner[static_cast<size_t>(hb.v_offset + i)] / per_rank
                     : 0;
    return SBSB200_OK;
}

int sbsb200_step(sbsb200_ctx* c, double dt, int substeps, int iterations, int detect_mode)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "step before finalize");
    if (c->world > 1 && !c->engine->peers_connected())
        return fail(c, SBSB200_ERR_STATE, "partitioned scene: connect the peers' mailboxes before stepping");
    if (!(dt > 0.) || substeps <= 0 || iterations < 0 ||
        (detect_mode != SBSB200_DETECT_PER_FRAME && detect_mode != SBSB200_DETECT_PER_SUBSTEP))
        return fail(c, SBSB200_ERR_INVALID, "bad step arguments");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->step(*c, dt, substeps, iterations, detect_mode);
        ++c->frames;
        return SBSB200_OK;
    });
}

int sbsb200_step_host(sbsb200_ctx* c, int body, const double* x_in, const double* v_in, double dt, int substeps,
                      int iterations, int detect_mode, double* x_out, double* v_out)
{
    if (!c || !x_in)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "step_host before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    // upload and step are only enqueued; the one synchronisation is the download's
    int rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->upload(*c, body, x_in, v_in, false);
        return SBSB200_OK;
    });
    if (rc)
        return rc;
    rc = sbsb200_step(c, dt, substeps, iterations, detect_mode);
    if (rc)
        return rc;
    if (!x_out && !v_out) // nothing to download: still return only when x_in / v_in may be reused and the step is done
        return sbsb200_synchronize(c);
    return sbsb200_download(c, body, x_out, v_out);
}

int sbsb200_step_host_f32(sbsb200_ctx* c, int body, const float* x_in, const float* v_in, double dt, int substeps,
                          int iterations, int detect_mode, float* x_out, float* v_out)
{
    if (!c || !x_in)
        return fail(c, SBSB200_ERR_INVALID, "null argument");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "step_host before finalize");
    if (body != SBSB200_ALL_BODIES && !is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    if (c->world > 1 && !c->engine->peers_connected())
        return fail(c, SBSB200_ERR_STATE, "partitioned scene: connect the peers' mailboxes before stepping");
    if (!(dt > 0.) || substeps <= 0 || iterations < 0 ||
        (detect_mode != SBSB200_DETECT_PER_FRAME && detect_mode != SBSB200_DETECT_PER_SUBSTEP))
        return fail(c, SBSB200_ERR_INVALID, "bad step arguments");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->step_host_f32(*c, body, x_in, v_in, dt, substeps, iterations, detect_mode, x_out, v_out);
        return SBSB200_OK;
    });
}


int sbsb200_step_host_vertices_f32(sbsb200_ctx* c, int body, int64_t n, const uint32_t* vertices, const float* x_in,
                                   const float* v_in, double dt, int substeps, int iterations, int detect_mode,
                                   float* x_out, float* v_out)
{
    if (!c || n < 0 || (n > 0 && !vertices))
        return fail(c, SBSB200_ERR_INVALID, "bad arguments");
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "step_host before finalize");
    if (!is_tet_body(c, body))
        return fail(c, SBSB200_ERR_INVALID, "not a tetrahedral body");
    for (int64_t i = 0; i < n; ++i)
        if (vertices[i] >= static_cast<uint64_t>(c->scene.bodies[static_cast<size_t>(body)].n_vertices))
            return fail(c, SBSB200_ERR_INVALID, "vertex index out of range");
    if (c->world > 1 && !c->engine->peers_connected())
        return fail(c, SBSB200_ERR_STATE, "partitioned scene: connect the peers' mailboxes before stepping");
    if (!(dt > 0.) || substeps <= 0 || iterations < 0 ||
        (detect_mode != SBSB200_DETECT_PER_FRAME && detect_mode != SBSB200_DETECT_PER_SUBSTEP))
        return fail(c, SBSB200_ERR_INVALID, "bad step arguments");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        c->engine->step_host_vertices_f32(*c, body, n, vertices, x_in, v_in, dt, substeps, iterations, detect_mode, x_out,
                                          v_out);
        return SBSB200_OK;
    });
}

int64_t sbsb200_count_non_finite(sbsb200_ctx* c)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "count_non_finite before finalize");
    int64_t n    = 0;
    int const rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        n = c->engine->count_non_finite(*c);
        return SBSB200_OK;
    });
    return rc < 0 ? rc : n;
}

int64_t sbsb200_debug_read_trace(sbsb200_ctx* c, int64_t* out, int64_t cap)
{
    if (!c || !c->finalized)
        return SBSB200_ERR_STATE;
    int64_t n    = 0;
    int const rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        n = c->engine->read_trace(*c, out, cap);
        return SBSB200_OK;
    });
    return rc ? rc : n;
}

int sbsb200_synchronize(sbsb200_ctx* c)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        if (c->engine)
            c->engine->check_after_sync(*c);
        return SBSB200_OK;
    });
}

int64_t sbsb200_get_contacts(sbsb200_ctx* c, int64_t cap, int32_t* body, uint32_t* vertex, int32_t* sdf_body,
                             double* point, double* normal)
{
    if (!c)
        return SBSB200_ERR_INVALID;
    if (!c->finalized)
        return fail(c, SBSB200_ERR_STATE, "get_contacts before finalize");
    int64_t n = 0;
    int const rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        n = c->engine->contacts(*c, cap, body, vertex, sdf_body, point, normal);
        return SBSB200_OK;
    });
    return rc ? rc : n;
}

} // extern "C"
