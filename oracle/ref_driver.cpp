// ref_driver.cpp — C entry points over the REFERENCE'S OWN classes (sbs::physics::*), compiled
// from /root/reference/src by oracle/build_ref.sh against oracle/ref_shim.  TEST INFRASTRUCTURE.
//
// Everything that computes here is reference code: simulation_t, tetrahedral_body_t,
// environment_body_t, green_constraint_t, distance_constraint_t, collision_constraint_t,
// xpbd::contact_handler_t, brute_force_cd_system_t, gauss_seidel_solver_t, timestep_t::step.
// This file only builds the scene the way main.cpp:22-86,120-124 does and copies state in/out.
// The exported functions mirror oracle/xpbd_oracle.h (prefix ref_ instead of orc_).
#include <cstring>
#include <memory>
#include <vector>

#include <sbs/common/geometry.h>
#include <sbs/physics/collision/brute_force_cd_system.h>
#include <sbs/physics/environment_body.h>
#include <sbs/physics/gauss_seidel_solver.h>
#include <sbs/physics/simulation.h>
#include <sbs/physics/tetrahedral_body.h>
#include <sbs/physics/timestep.h>
#include <sbs/physics/xpbd/contact_handler.h>
#include <sbs/physics/xpbd/distance_constraint.h>
#include <sbs/physics/xpbd/green_constraint.h>

using namespace sbs;
using namespace sbs::physics;

namespace {

struct recorded_contact
{
    index_type b1, b2, surface_vertex;
    Eigen::Vector3d p, n;
};

// forwards to the reference's xpbd::contact_handler_t and keeps a copy for ref_get_contacts
class recording_handler_t : public collision::contact_handler_t
{
  public:
    recording_handler_t(simulation_t& s, std::vector<recorded_contact>& log) : inner_(s), log_(log) {}
    void handle(collision::contact_t const& c) override
    {
        auto const& sc = reinterpret_cast<collision::surface_mesh_particle_to_sdf_contact_t const&>(c);
        log_.push_back({c.b1(), c.b2(), sc.vi(), c.point(), c.normal()});
        inner_.handle(c);
    }

  private:
    xpbd::contact_handler_t inner_;
    std::vector<recorded_contact>& log_;
};

struct world
{
    simulation_t sim;
    std::vector<int> is_tet; // per body
    std::vector<int> excluded; // bodies whose collision model is not handed to the cd system
    std::vector<recorded_contact> log;
    bool cd_ready = false;
};

common::geometry_t dummy_triangle()
{
    common::geometry_t g;
    g.positions     = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    g.indices       = {0, 1, 2};
    g.geometry_type = common::geometry_t::geometry_type_t::triangle;
    g.set_color(100, 100, 100);
    return g;
}

int add_env(world* w, collision::sdf_model_t const& model)
{
    auto const idx = static_cast<index_type>(w->sim.bodies().size());
    w->sim.add_body(std::make_unique<environment_body_t>(w->sim, idx, dummy_triangle(), model)); // main.cpp:68-72
    w->is_tet.push_back(0);
    w->cd_ready = false;
    return static_cast<int>(idx);
}

void ensure_cd(world* w)
{
    if (w->cd_ready)
        return;
    std::vector<collision::collision_model_t*> objects; // main.cpp:77-86
    for (std::size_t i = 0; i < w->sim.bodies().size(); ++i)
        if (i >= w->excluded.size() || !w->excluded[i])
            objects.push_back(&(w->sim.bodies()[i]->collision_model()));
    w->sim.use_collision_detection_system(std::make_unique<collision::brute_force_cd_system_t>(objects));
    w->sim.collision_detection_system()->use_contact_handler(
        std::make_unique<recording_handler_t>(w->sim, w->log));
    w->cd_ready = true;
}

} // namespace

extern "C" {

void* ref_create() { return new world(); }
void ref_destroy(void* h) { delete static_cast<world*>(h); }

void ref_set_collision_compliance(void* h, double a)
{
    static_cast<world*>(h)->sim.simulation_parameters().collision_compliance = a;
}

int ref_add_tet_body(void* h, int nV, const double* x0, const double* mass, int nT, const uint32_t* tets,
                     double young, double poisson, double alpha, double beta)
{
    world* w       = static_cast<world*>(h);
    auto const idx = static_cast<index_type>(w->sim.bodies().size());
    std::vector<Eigen::Vector3d> positions;
    positions.reserve(static_cast<std::size_t>(nV));
    for (int i = 0; i < nV; ++i)
        positions.emplace_back(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]);
    tetrahedron_set_t topology;
    topology.reserve_vertices(static_cast<std::size_t>(nV));
    for (int t = 0; t < nT; ++t)
        topology.add_tetrahedron(tetrahedron_t{tets[4 * t], tets[4 * t + 1], tets[4 * t + 2], tets[4 * t + 3]});
    w->sim.add_body(); // main.cpp:26 — the slot must exist before the body constructor runs
    w->sim.bodies()[idx] = std::make_unique<tetrahedral_body_t>(w->sim, idx, positions, topology);
    w->is_tet.push_back(1);
    if (mass)
        for (int i = 0; i < nV; ++i)
            w->sim.particles()[idx][static_cast<std::size_t>(i)].mass() = mass[i];
    auto const& body = *dynamic_cast<tetrahedral_body_t*>(w->sim.bodies()[idx].get());
    for (auto const& tet : body.physical_model().tetrahedra()) // main.cpp:37-54
        w->sim.add_constraint(std::make_unique<xpbd::green_constraint_t>(
            alpha, beta, w->sim, idx, tet.v1(), tet.v2(), tet.v3(), tet.v4(), young, poisson));
    w->cd_ready = false;
    return static_cast<int>(idx);
}

int ref_add_distance_constraints(void* h, int b1, int b2, int n, const uint32_t* pairs, double alpha, double beta)
{
    world* w = static_cast<world*>(h);
    for (int i = 0; i < n; ++i)
        w->sim.add_constraint(std::make_unique<xpbd::distance_constraint_t>(
            alpha, beta, w->sim, static_cast<index_type>(b1), static_cast<index_type>(b2), pairs[2 * i],
            pairs[2 * i + 1]));
    return 0;
}

int ref_add_sdf_plane(void* h, const double n[3], const double pt[3], const double vol[6])
{
    Eigen::AlignedBox3d const volume{Eigen::Vector3d{vol[0], vol[1], vol[2]}, Eigen::Vector3d{vol[3], vol[4], vol[5]}};
    auto const model = collision::sdf_model_t::from_plane( // main.cpp:63-67
        Eigen::Hyperplane<scalar_type, 3>(Eigen::Vector3d{n[0], n[1], n[2]}, Eigen::Vector3d{pt[0], pt[1], pt[2]}),
        volume);
    return add_env(static_cast<world*>(h), model);
}

int ref_add_sdf_sphere(void* h, const double c[3], double r, const double vol[6])
{
    Eigen::AlignedBox3d const volume{Eigen::Vector3d{vol[0], vol[1], vol[2]}, Eigen::Vector3d{vol[3], vol[4], vol[5]}};
    Eigen::Vector3d const centre{c[0], c[1], c[2]};
    collision::sdf_model_t::analytic_sdf_type const f =
        [centre, r](Eigen::Vector3d const& p) -> std::pair<scalar_type, Eigen::Vector3d> {
        Eigen::Vector3d const d = p - centre;
        scalar_type const len   = d.norm();
        Eigen::Vector3d const g = len > 0. ? Eigen::Vector3d(d / len) : Eigen::Vector3d{0., 1., 0.};
        return {len - r, g};
    };
    return add_env(static_cast<world*>(h), collision::sdf_model_t{f, volume}); // sdf_model.h:23
}

int ref_add_sdf_box(void* h, const double bmin[3], const double bmax[3], const double vol[6])
{
    Eigen::AlignedBox3d const volume{Eigen::Vector3d{vol[0], vol[1], vol[2]}, Eigen::Vector3d{vol[3], vol[4], vol[5]}};
    Eigen::Vector3d const lo{bmin[0], bmin[1], bmin[2]}, hi{bmax[0], bmax[1], bmax[2]};
    collision::sdf_model_t::analytic_sdf_type const f =
        [lo, hi](Eigen::Vector3d const& p) -> std::pair<scalar_type, Eigen::Vector3d> {
        Eigen::Vector3d const c = (lo + hi) * 0.5, hf = (hi - lo) * 0.5;
        double q[3], out2 = 0.;
        for (int i = 0; i < 3; ++i)
        {
            q[i] = std::abs(p(i) - c(i)) - hf(i);
            if (q[i] > 0.)
                out2 += q[i] * q[i];
        }
        Eigen::Vector3d g{0., 0., 0.};
        if (out2 > 0.)
        {
            double const len = std::sqrt(out2);
            for (int i = 0; i < 3; ++i)
                g(i) = q[i] > 0. ? (p(i) >= c(i) ? q[i] : -q[i]) / len : 0.;
            return {len, g};
        }
        int ax = 0;
        if (q[1] > q[ax]) ax = 1;
        if (q[2] > q[ax]) ax = 2;
        g(ax) = p(ax) >= c(ax) ? 1. : -1.;
        return {q[ax], g};
    };
    return add_env(static_cast<world*>(h), collision::sdf_model_t{f, volume});
}

// environment_body_t(simulation, id, geometry, domain, resolution) — the reference's own bake
// (environment_body.cpp:12-78) over the shim's Discregrid stand-ins.  geometry_t holds float positions.
int ref_add_sdf_mesh(void* h, int nV, const double* x, int nF, const uint32_t* faces, const double dom[6],
                     const uint32_t res[3])
{
    world* w = static_cast<world*>(h);
    common::geometry_t g;
    for (int i = 0; i < 3 * nV; ++i)
        g.positions.push_back(static_cast<float>(x[i]));
    for (int i = 0; i < 3 * nF; ++i)
        g.indices.push_back(static_cast<int>(faces[i]));
    g.geometry_type = common::geometry_t::geometry_type_t::triangle;
    g.set_color(100, 100, 100);
    Eigen::AlignedBox3d const domain{Eigen::Vector3d{dom[0], dom[1], dom[2]}, Eigen::Vector3d{dom[3], dom[4], dom[5]}};
    auto const idx = static_cast<index_type>(w->sim.bodies().size());
    w->sim.add_body(std::make_unique<environment_body_t>(w->sim, idx, g, domain,
                                                         std::array<unsigned int, 3u>{res[0], res[1], res[2]}));
    w->is_tet.push_back(0);
    w->cd_ready = false;
    return static_cast<int>(idx);
}

// sdf_model_t(Discregrid::CubicLagrangeDiscreteGrid const&) (sdf_model.cpp:18) with given node values:
// addFunction samples its argument at the nodes in node order, so a counting functor installs them.
int ref_add_sdf_grid(void* h, const double dmin[3], const double dmax[3], const uint32_t res[3], const double* nodes,
                     const double vol[6])
{
    Eigen::AlignedBox3d const domain{Eigen::Vector3d{dmin[0], dmin[1], dmin[2]}, Eigen::Vector3d{dmax[0], dmax[1], dmax[2]}};
    Discregrid::CubicLagrangeDiscreteGrid grid(domain, {res[0], res[1], res[2]});
    std::size_t next = 0;
    grid.addFunction([&](Eigen::Vector3d const&) { return nodes[next++]; });
    collision::sdf_model_t model(grid);
    model.volume() = vol ? Eigen::AlignedBox3d{Eigen::Vector3d{vol[0], vol[1], vol[2]}, Eigen::Vector3d{vol[3], vol[4], vol[5]}}
                         : domain;
    return add_env(static_cast<world*>(h), model);
}

// sdf_model_t::evaluate (sdf_model.cpp:66-75) of an environment body at n points
int ref_sdf_evaluate(void* h, int body, int n, const double* pts, double* sd, double* grad)
{
    world* w = static_cast<world*>(h);
    auto const* env = dynamic_cast<environment_body_t const*>(w->sim.bodies()[static_cast<std::size_t>(body)].get());
    if (!env)
        return -1;
    for (int i = 0; i < n; ++i)
    {
        auto const [d, g] = env->sdf().evaluate(Eigen::Vector3d{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
        sd[i]             = d;
        for (int k = 0; k < 3; ++k)
            grad[3 * i + k] = g(k);
    }
    return 0;
}

int ref_get_volume(void* h, int body, double out[6])
{
    world* w       = static_cast<world*>(h);
    auto const& vb = w->sim.bodies()[static_cast<std::size_t>(body)]->collision_model().volume();
    for (int k = 0; k < 3; ++k)
    {
        out[k]     = vb.min()(k);
        out[3 + k] = vb.max()(k);
    }
    return 0;
}

int ref_set_body_collideable(void* h, int body, int flag)
{
    world* w = static_cast<world*>(h);
    if (body < 0 || static_cast<std::size_t>(body) >= w->sim.bodies().size())
        return -1;
    w->excluded.resize(w->sim.bodies().size(), 0);
    w->excluded[static_cast<std::size_t>(body)] = !flag;
    w->cd_ready = false;
    return 0;
}

int ref_constraint_count(void* h) { return static_cast<int>(static_cast<world*>(h)->sim.constraints().size()); }

// simulation_t::remove_constraint (simulation.cpp:34-39), the reference's own code: swap with the last, drop it
int ref_remove_constraint(void* h, uint32_t index)
{
    world* w = static_cast<world*>(h);
    if (index >= w->sim.constraints().size())
        return -1;
    w->sim.remove_constraint(index);
    return 0;
}

// the reference run "with constraints permuted": reorder simulation_t::constraints_
int ref_set_constraint_order(void* h, const uint32_t* order, int n)
{
    world* w = static_cast<world*>(h);
    auto& cs = w->sim.constraints();
    if (n != static_cast<int>(cs.size()))
        return -1;
    std::vector<std::unique_ptr<constraint_t>> next(cs.size());
    for (int i = 0; i < n; ++i)
    {
        if (order[i] >= cs.size() || !cs[order[i]])
            return -2;
        next[static_cast<std::size_t>(i)] = std::move(cs[order[i]]);
    }
    cs = std::move(next);
    return 0;
}

int ref_upload(void* h, int b, const double* x, const double* v)
{
    world* w = static_cast<world*>(h);
    auto& ps = w->sim.particles().at(static_cast<std::size_t>(b));
    for (std::size_t i = 0; i < ps.size(); ++i)
    {
        Eigen::Vector3d const p{x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        ps[i].x() = p;
        ps[i].xi() = p;
        ps[i].xn() = p;
        ps[i].v() = v ? Eigen::Vector3d{v[3 * i], v[3 * i + 1], v[3 * i + 2]} : Eigen::Vector3d{0., 0., 0.};
    }
    w->sim.bodies()[static_cast<std::size_t>(b)]->update_visual_model(); // as transform() does (tetrahedral_body.cpp:131)
    return 0;
}

int ref_download(void* h, int b, double* x, double* v)
{
    world* w       = static_cast<world*>(h);
    auto const& ps = w->sim.particles().at(static_cast<std::size_t>(b));
    for (std::size_t i = 0; i < ps.size(); ++i)
        for (int r = 0; r < 3; ++r)
        {
            if (x)
                x[3 * i + r] = ps[i].x()(r);
            if (v)
                v[3 * i + r] = ps[i].v()(r);
        }
    return 0;
}

int ref_set_mass(void* h, int b, int vertex, double m)
{
    static_cast<world*>(h)->sim.particles().at(static_cast<std::size_t>(b)).at(static_cast<std::size_t>(vertex)).mass() = m;
    return 0;
}

int ref_step(void* h, double dt, int substeps, int iterations, int detect_every_substep)
{
    world* w = static_cast<world*>(h);
    ensure_cd(w);
    auto run = [&](double dtf, int s) {
        timestep_t ts{}; // main.cpp:120-124
        ts.dt()         = dtf;
        ts.iterations() = static_cast<std::size_t>(iterations);
        ts.substeps()   = static_cast<std::size_t>(s);
        ts.solver()     = std::make_unique<gauss_seidel_solver_t>();
        w->log.clear();
        ts.step(w->sim);
    };
    if (!detect_every_substep)
        run(dt, substeps);
    else
        for (int s = 0; s < substeps; ++s)
            run(dt / static_cast<double>(substeps), 1);
    return 0;
}

int ref_get_contacts(void* h, int cap, int32_t* body, uint32_t* vertex, int32_t* sdf_body, double* point, double* normal)
{
    world* w    = static_cast<world*>(h);
    int const n = static_cast<int>(w->log.size());
    for (int i = 0; i < n && i < cap && body; ++i)
    {
        recorded_contact const& c = w->log[static_cast<std::size_t>(i)];
        auto const* tb = dynamic_cast<tetrahedral_body_t const*>(w->sim.bodies()[c.b1].get());
        body[i]        = static_cast<int32_t>(c.b1);
        vertex[i]      = tb ? tb->surface_mesh().from_surface_vertex(c.surface_vertex) : c.surface_vertex;
        sdf_body[i]    = static_cast<int32_t>(c.b2);
        for (int r = 0; r < 3; ++r)
        {
            point[3 * i + r]  = c.p(r);
            normal[3 * i + r] = c.n(r);
        }
    }
    return n;
}

// tetrahedral_mesh_boundary_t::surface_to_tetrahedral_mesh_index_map of a body
int ref_get_surface_map(void* h, int b, uint32_t* map, int cap)
{
    world* w       = static_cast<world*>(h);
    auto const* tb = dynamic_cast<tetrahedral_body_t const*>(w->sim.bodies()[static_cast<std::size_t>(b)].get());
    if (!tb)
        return -1;
    auto const& m = tb->surface_mesh().surface_to_tetrahedral_mesh_index_map();
    for (std::size_t i = 0; i < m.size() && static_cast<int>(i) < cap && map; ++i)
        map[i] = m[i];
    return static_cast<int>(m.size());
}

} // extern "C"
