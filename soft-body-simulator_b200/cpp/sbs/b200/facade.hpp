// sbs/b200/facade.hpp — the reference's C++ entry points for the XPBD hot path, re-created over the
// C ABI of libsbsb200.so (include/sbs_b200.h).  Header only; link with -lsbsb200.
//
// A program written against the reference's physics API (main.cpp:17-130 is the canonical one)
// compiles against these headers with the same include paths (<sbs/physics/simulation.h>, ...),
// class names, constructor signatures and call order.  What differs, and why:
//   * Eigen is not a dependency here.  sbs::vec3 / sbs::affine3 / sbs::aligned_box3 / sbs::hyperplane3
//     stand where the reference uses Eigen::Vector3d / Affine3d / AlignedBox3d / Hyperplane (same
//     member names for what the path uses: x() y() z() operator[] min() max() normal() offset()).
//   * The replaceable unit is timestep_t::step(simulation_t&) as a whole (predict and commit are
//     inline in it, src/physics/timestep.cpp:35-57).  constraint_t::project_positions and
//     solver_t::solve exist for source compatibility and throw: there is no CPU path.
//   * sdf_model_t takes analytic shapes (plane, sphere, box) instead of a std::function
//     (include/sbs/physics/collision/sdf_model.h:18-23): a host callback cannot run on the device.
//     Discrete grids are there: from_grid (node values) and environment_body_t's triangle-mesh
//     constructor (src/physics/environment_body.cpp:12-78), whose mesh distance is baked on the device.
//   * particles() is a host mirror: it is refreshed from the device when read after a step, and
//     written back (x, v, mass) before the next step when it was handed out mutable — the way
//     main.cpp:158-165 pins a picked vertex by setting its mass to 0 between frames.
// Every class cites the reference declaration it stands for.
#ifndef SBS_B200_FACADE_HPP
#define SBS_B200_FACADE_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <unordered_map>
#include <vector>

#include "../../../../include/sbs_b200.h"
#include "../common/node.h"

namespace sbs {

using scalar_type = double;        // include/sbs/aliases.h:9
using index_type  = std::uint32_t; // include/sbs/aliases.h:8

// ---- minimal stand-ins for the Eigen types on the path ------------------------------------------
struct vec3
{
    double v[3]{0., 0., 0.};
    vec3() = default;
    vec3(double x, double y, double z) : v{x, y, z} {}
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double& x() { return v[0]; }
    double& y() { return v[1]; }
    double& z() { return v[2]; }
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    vec3 operator+(vec3 const& o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
    vec3 operator-(vec3 const& o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
    vec3 operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
    double dot(vec3 const& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double norm() const { return std::sqrt(dot(*this)); }
    vec3 normalized() const
    {
        double const n = norm();
        return {v[0] / n, v[1] / n, v[2] / n};
    }
    void setZero() { v[0] = v[1] = v[2] = 0.; }
};

// rows of [A | t]: p -> A p + t (Eigen::Affine3d)
struct affine3
{
    double m[3][4]{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
    static affine3 translation(double x, double y, double z)
    {
        affine3 a;
        a.m[0][3] = x;
        a.m[1][3] = y;
        a.m[2][3] = z;
        return a;
    }
    // this <- this * rotation(angle, unit axis)   (Eigen: transform.rotate(AngleAxisd))
    affine3& rotate(double angle, vec3 axis)
    {
        axis           = axis.normalized();
        double const c = std::cos(angle), s = std::sin(angle), t = 1. - c;
        double const r[3][3] = {
            {t * axis[0] * axis[0] + c, t * axis[0] * axis[1] - s * axis[2], t * axis[0] * axis[2] + s * axis[1]},
            {t * axis[0] * axis[1] + s * axis[2], t * axis[1] * axis[1] + c, t * axis[1] * axis[2] - s * axis[0]},
            {t * axis[0] * axis[2] - s * axis[1], t * axis[1] * axis[2] + s * axis[0], t * axis[2] * axis[2] + c}};
        double n[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                n[i][j] = m[i][0] * r[0][j] + m[i][1] * r[1][j] + m[i][2] * r[2][j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                m[i][j] = n[i][j];
        return *this;
    }
    // this <- this * diag(s)   (Eigen: transform.scale(Vector3d))
    affine3& scale(vec3 const& s)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                m[i][j] *= s[j];
        return *this;
    }
    vec3 operator*(vec3 const& p) const
    {
        return {m[0][0] * p[0] + m[0][1] * p[1] + m[0][2] * p[2] + m[0][3],
                m[1][0] * p[0] + m[1][1] * p[1] + m[1][2] * p[2] + m[1][3],
                m[2][0] * p[0] + m[2][1] * p[1] + m[2][2] * p[2] + m[2][3]};
    }
};

struct aligned_box3
{
    vec3 lo, hi;
    aligned_box3() = default;
    aligned_box3(vec3 const& mn, vec3 const& mx) : lo(mn), hi(mx) {}
    vec3 const& min() const { return lo; }
    vec3 const& max() const { return hi; }
};

// Eigen::Hyperplane<double, 3>(normal, point): signedDistance(p) = n.p + offset
struct hyperplane3
{
    vec3 n;
    double d = 0.;
    hyperplane3() = default;
    hyperplane3(vec3 const& normal, vec3 const& point) : n(normal), d(-normal.dot(point)) {}
    vec3 const& normal() const { return n; }
    double offset() const { return d; }
    double signedDistance(vec3 const& p) const { return n.dot(p) + d; }
};

namespace b200 {
struct error : std::runtime_error
{
    int code;
    error(int c, std::string const& what) : std::runtime_error(what), code(c) {}
};
} // namespace b200

namespace common {

// include/sbs/common/geometry.h:9-29
struct geometry_t
{
    std::vector<float> positions;
    std::vector<int> indices;
    std::vector<float> normals;
    std::vector<float> uvs;
    std::vector<std::uint8_t> colors;
    enum class geometry_type_t { triangle, tetrahedron };
    geometry_type_t geometry_type = geometry_type_t::triangle;
    bool has_colors() const { return !colors.empty(); }
    bool has_positions() const { return !positions.empty(); }
    bool has_indices() const { return !indices.empty(); }
    bool has_normals() const { return !normals.empty(); }
    bool has_uvs() const { return !uvs.empty(); }
    bool is_triangle_mesh() const { return geometry_type == geometry_type_t::triangle; }
    bool is_tetrahedral_mesh() const { return geometry_type == geometry_type_t::tetrahedron; }
    void set_color(std::uint8_t r, std::uint8_t g, std::uint8_t b)
    {
        colors.clear();
        for (std::size_t i = 0; i < positions.size() / 3; ++i)
        {
            colors.push_back(r);
            colors.push_back(g);
            colors.push_back(b);
        }
    }
};

} // namespace common

namespace geometry {

// include/sbs/geometry/get_simple_bar_model.h:9 — unit lattice, vertex id i*H*D + j*D + k, five tets
// per cell with the orientation alternating on (i+j+k) % 2 (src/geometry/get_simple_bar_model.cpp:44-113).
// Pinned by the reference's data/meshes/cube_tet.ply, tet_bar_5x2x2.ply, bar_tet.ply (tests/golden).
inline common::geometry_t get_simple_bar_model(std::size_t width, std::size_t height, std::size_t depth)
{
    common::geometry_t g;
    g.geometry_type = common::geometry_t::geometry_type_t::tetrahedron;
    for (std::size_t i = 0; i < width; ++i)
        for (std::size_t j = 0; j < height; ++j)
            for (std::size_t k = 0; k < depth; ++k)
            {
                g.positions.push_back(static_cast<float>(i));
                g.positions.push_back(static_cast<float>(j));
                g.positions.push_back(static_cast<float>(k));
            }
    static int const corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                     {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static int const odd[5][4]    = {{1, 0, 5, 2}, {5, 2, 7, 6}, {7, 0, 5, 4}, {2, 0, 7, 3}, {5, 0, 7, 2}};
    static int const even[5][4]   = {{3, 1, 4, 0}, {6, 1, 3, 2}, {4, 1, 6, 5}, {6, 3, 4, 7}, {3, 1, 6, 4}};
    for (std::size_t i = 0; i + 1 < width; ++i)
        for (std::size_t j = 0; j + 1 < height; ++j)
            for (std::size_t k = 0; k + 1 < depth; ++k)
            {
                int p[8];
                for (int c = 0; c < 8; ++c)
                    p[c] = static_cast<int>(((i + corner[c][0]) * height + (j + corner[c][1])) * depth + (k + corner[c][2]));
                auto const& table = ((i + j + k) % 2 == 1) ? odd : even;
                for (auto const& t : table)
                    for (int a = 0; a < 4; ++a)
                        g.indices.push_back(p[t[a]]);
            }
    return g;
}

} // namespace geometry

namespace physics {

class simulation_t;

// include/sbs/physics/particle.h:10-53, src/physics/particle.cpp:6-53
class particle_t
{
  public:
    using position_type     = vec3;
    using velocity_type     = vec3;
    using acceleration_type = vec3;
    using force_type        = vec3;
    particle_t() = default;
    particle_t(position_type const& p) : x0_(p), xi_(p), xn_(p), x_(p) {}
    position_type const& x0() const { return x0_; }
    position_type const& x() const { return x_; }
    position_type const& xi() const { return xi_; }
    position_type const& xn() const { return xn_; }
    velocity_type const& v() const { return v_; }
    force_type const& f() const { return f_; }
    scalar_type const& mass() const { return m_; }
    scalar_type invmass() const { return m_ > 0. ? 1. / m_ : 0.; }
    bool fixed() const { return m_ == 0.; }
    acceleration_type a() const { return f_ * invmass(); }
    position_type& x0() { return x0_; }
    position_type& x() { return x_; }
    position_type& xi() { return xi_; }
    position_type& xn() { return xn_; }
    velocity_type& v() { return v_; }
    force_type& f() { return f_; }
    scalar_type& mass() { return m_; }

  private:
    position_type x0_, xi_, xn_, x_;
    velocity_type v_;
    force_type f_;
    scalar_type m_{1.}; // particle.cpp:7
};

namespace xpbd {
// include/sbs/physics/xpbd/simulation_parameters.h:10-28
struct simulation_parameters_t
{
    scalar_type compliance           = 1e-4;
    scalar_type damping              = 0.;
    scalar_type young_modulus        = 1e6;
    scalar_type poisson_ratio        = 0.30;
    scalar_type hooke_coefficient    = 1.;
    scalar_type collision_compliance = 1e-8;
    scalar_type collision_damping    = 0.;
};
} // namespace xpbd

// include/sbs/physics/constraint.h:12-35.  Constraints are descriptions here: the device owns the
// multipliers and does the projecting.
class constraint_t
{
  public:
    constraint_t(scalar_type alpha, scalar_type beta) : alpha_(alpha), beta_(beta) {}
    virtual ~constraint_t() = default;
    void prepare_for_projection(simulation_t&) {}
    virtual void project_positions(simulation_t&, scalar_type)
    {
        throw std::logic_error("sbs-b200: constraints are projected on the device by timestep_t::step; there is no "
                               "CPU path");
    }
    scalar_type alpha() const { return alpha_; }
    scalar_type beta() const { return beta_; }

  private:
    scalar_type alpha_, beta_;
};

namespace collision {

class contact_handler_t;

// include/sbs/physics/collision/collision_model.h:22-51
class collision_model_t
{
  public:
    enum class model_type_t { bvh, sdf };
    virtual ~collision_model_t() = default;
    virtual model_type_t model_type() const = 0;
    aligned_box3 const& volume() const { return volume_; }
    aligned_box3& volume() { return volume_; }
    index_type id() const { return id_; }
    index_type& id() { return id_; }

  private:
    aligned_box3 volume_;
    index_type id_ = 0;
};

// include/sbs/physics/collision/sdf_model.h:15-44 (analytic shapes and discrete grids, see the file comment)
class sdf_model_t : public collision_model_t
{
  public:
    enum class shape_t { plane, sphere, box, grid, mesh };
    model_type_t model_type() const override { return model_type_t::sdf; }
    // sdf_model.cpp:52-64
    static sdf_model_t from_plane(hyperplane3 const& plane, aligned_box3 const& volume)
    {
        sdf_model_t m;
        m.shape_    = shape_t::plane;
        m.a_        = plane.normal();
        m.d_        = plane.offset();
        m.volume()  = volume;
        return m;
    }
    static sdf_model_t from_sphere(vec3 const& centre, scalar_type radius, aligned_box3 const& volume)
    {
        sdf_model_t m;
        m.shape_   = shape_t::sphere;
        m.a_       = centre;
        m.d_       = radius;
        m.volume() = volume;
        return m;
    }
    static sdf_model_t from_box(vec3 const& box_min, vec3 const& box_max, aligned_box3 const& volume)
    {
        sdf_model_t m;
        m.shape_   = shape_t::box;
        m.a_       = box_min;
        m.b_       = box_max;
        m.volume() = volume;
        return m;
    }
    // sdf_model_t(Discregrid::CubicLagrangeDiscreteGrid const&) (sdf_model.cpp:18): node values in Discregrid's
    // node order over `domain` (positions: sbsb200_grid_node_position); volume() = the domain
    static sdf_model_t from_grid(aligned_box3 const& domain, std::array<unsigned int, 3u> const& resolution,
                                 std::vector<scalar_type> node_values)
    {
        sdf_model_t m;
        m.shape_      = shape_t::grid;
        m.a_          = domain.min();
        m.b_          = domain.max();
        m.resolution_ = resolution;
        m.values_     = std::move(node_values);
        m.volume()    = domain;
        return m;
    }
    // the model environment_body_t's mesh constructor builds (environment_body.cpp:12-78): the grid is
    // sampled on the device when the scene is built; a..b is the extended domain
    static sdf_model_t from_mesh(std::vector<scalar_type> positions, std::vector<std::uint32_t> triangles,
                                 aligned_box3 const& domain, std::array<unsigned int, 3u> const& resolution)
    {
        sdf_model_t m;
        m.shape_            = shape_t::mesh;
        m.resolution_       = resolution;
        double const in[6]  = {domain.min()[0], domain.min()[1], domain.min()[2],
                               domain.max()[0], domain.max()[1], domain.max()[2]};
        double out[6];
        if (sbsb200_mesh_sdf_domain(static_cast<std::int64_t>(positions.size() / 3), positions.data(), in, out) != 0)
            throw std::invalid_argument("sbs-b200: bad obstacle mesh");
        m.given_domain_ = domain;
        m.a_            = vec3(out[0], out[1], out[2]);
        m.b_            = vec3(out[3], out[4], out[5]);
        m.volume()      = aligned_box3{m.a_, m.b_}; // environment_body.cpp:76
        m.values_       = std::move(positions);
        m.triangles_    = std::move(triangles);
        return m;
    }
    // sdf_model.cpp:52-75.  On the host only the plane is evaluated (signed distance and gradient as
    // Eigen::Hyperplane gives them, :57-61); the other shapes are sampled where the detection samples them:
    // simulation_t::evaluate_sdf(body, p) asks the device (sbsb200_eval_sdf).
    std::pair<scalar_type, vec3> evaluate(vec3 const& p) const
    {
        if (shape_ != shape_t::plane)
            throw std::logic_error("sbs-b200: sdf_model_t::evaluate on the host covers planes; use "
                                   "simulation_t::evaluate_sdf for the shapes the device samples");
        return {a_[0] * p[0] + a_[1] * p[1] + a_[2] * p[2] + d_, a_};
    }
    shape_t shape() const { return shape_; }
    vec3 const& a() const { return a_; }
    vec3 const& b() const { return b_; }
    scalar_type d() const { return d_; }
    std::array<unsigned int, 3u> const& resolution() const { return resolution_; }
    std::vector<scalar_type> const& values() const { return values_; }        // grid: nodes; mesh: positions
    std::vector<std::uint32_t> const& triangles() const { return triangles_; }
    aligned_box3 const& given_domain() const { return given_domain_; }

  private:
    shape_t shape_ = shape_t::plane;
    vec3 a_, b_;
    scalar_type d_ = 0.;
    std::array<unsigned int, 3u> resolution_{{10u, 10u, 10u}};
    std::vector<scalar_type> values_;
    std::vector<std::uint32_t> triangles_;
    aligned_box3 given_domain_;
};

// include/sbs/physics/collision/bvh_model.h:23-49: the surface vertices of a tetrahedral body.
// The device finds contacts itself; the object exists so that cd_system_t can be set up as in main.cpp:77-84.
class point_bvh_model_t : public collision_model_t
{
  public:
    model_type_t model_type() const override { return model_type_t::bvh; }
};

// include/sbs/physics/collision/contact.h:11-58
class contact_t
{
  public:
    enum class type_t { surface_particle_to_sdf };
    contact_t(type_t contact_type, index_type body1, index_type body2, vec3 const& contact_point, vec3 const& contact_normal)
        : type_(contact_type), bodies_{body1, body2}, point_(contact_point), normal_(contact_normal)
    {
    }
    type_t type() const { return type_; }
    vec3 const& point() const { return point_; }
    vec3 const& normal() const { return normal_; }
    index_type b1() const { return bodies_[0]; }
    index_type b2() const { return bodies_[1]; }

  private:
    type_t type_;
    index_type bodies_[2];
    vec3 point_, normal_;
};
class surface_mesh_particle_to_sdf_contact_t : public contact_t
{
  public:
    surface_mesh_particle_to_sdf_contact_t(contact_t::type_t contact_type, index_type body1, index_type body2,
                                           vec3 const& contact_point, vec3 const& contact_normal, index_type vi)
        : contact_t(contact_type, body1, body2, contact_point, contact_normal), vi_(vi)
    {
    }
    index_type vi() const { return vi_; } // surface vertex of body b1 (tetrahedral_mesh_boundary_t numbering)
    index_type& vi() { return vi_; }

  private:
    index_type vi_;
};

// include/sbs/physics/collision/contact.h:60-64.  The device turns contacts into collision constraints itself
// (xpbd/contact_handler.cpp:14-54); simulation_t::contacts() lists what it found at the last detection.
class contact_handler_t
{
  public:
    virtual ~contact_handler_t() = default;
    virtual void handle(contact_t const&) {}
};

// include/sbs/physics/collision/cd_system.h:19-47
class cd_system_t
{
  public:
    cd_system_t(std::vector<collision_model_t*> const& collision_objects) : collision_objects_(collision_objects) {}
    virtual ~cd_system_t() = default;
    std::vector<collision_model_t*> const& collision_objects() const { return collision_objects_; }
    std::unique_ptr<contact_handler_t> const& contact_handler() const { return contact_handler_; }
    std::unique_ptr<contact_handler_t>& contact_handler() { return contact_handler_; }
    void use_contact_handler(std::unique_ptr<contact_handler_t> h) { contact_handler_ = std::move(h); }
    // cd_system.h:36-38: refit and pair loop — both happen on the device inside timestep_t::step
    // (k_bvh_fit, k_detect_all); calling them on the facade changes nothing
    virtual void update(simulation_t const&) {}
    virtual void execute() {}

  private:
    std::vector<collision_model_t*> collision_objects_;
    std::unique_ptr<contact_handler_t> contact_handler_;
};

// include/sbs/physics/collision/brute_force_cd_system.h — all body pairs; only body-SDF pairs act
// (bvh_model.cpp:30-100, sdf_model.cpp:36-44).  On the device: every surface vertex against every SDF.
class brute_force_cd_system_t : public cd_system_t
{
  public:
    using cd_system_t::cd_system_t;
};

} // namespace collision

namespace xpbd {
// include/sbs/physics/xpbd/contact_handler.h — contacts become collision_constraint_t with
// simulation_parameters().collision_compliance (xpbd/contact_handler.cpp:42-52); on the device.
class contact_handler_t : public collision::contact_handler_t
{
  public:
    explicit contact_handler_t(simulation_t&) {}
};
} // namespace xpbd

// include/sbs/physics/body.h:16-46
} // namespace physics

namespace common {

// include/sbs/common/mesh.h:13-40: what a renderer draws
class shared_vertex_surface_mesh_i : public renderable_node_t
{
  public:
    struct vertex_type
    {
        vec3 position, normal;
        std::array<float, 3> color{{0.f, 0.f, 0.f}};
    };
    struct triangle_type
    {
        std::array<std::uint32_t, 3> vertices;
    };
    virtual std::size_t triangle_count() const         = 0;
    virtual std::size_t vertex_count() const           = 0;
    virtual vertex_type vertex(std::size_t vi) const   = 0;
    virtual triangle_type triangle(std::size_t f) const = 0;
};

} // namespace common

namespace physics {

// include/sbs/physics/tetrahedral_mesh_boundary.h: the boundary surface of a tetrahedral body.  Triangles, the
// surface -> tetrahedral-mesh vertex map (tetrahedral_mesh_boundary.cpp:49-58, :65-120) and, after every
// tetrahedral_body_t::update_visual_model, positions and normals come from the device (sbsb200_get_surface_map,
// _get_surface_triangles, _download_surface: boundary gather and normals are kernels, nothing is recomputed here).
class tetrahedral_mesh_boundary_t : public common::shared_vertex_surface_mesh_i
{
  public:
    using vertex_type   = common::shared_vertex_surface_mesh_i::vertex_type;
    using triangle_type = common::shared_vertex_surface_mesh_i::triangle_type;
    std::size_t triangle_count() const override { return triangles_.size(); }
    std::size_t vertex_count() const override { return vertices_.size(); }
    vertex_type vertex(std::size_t vi) const override { return vertices_.at(vi); }
    triangle_type triangle(std::size_t f) const override { return triangles_.at(f); }
    vertex_type& mutable_vertex(std::size_t vi) { return vertices_.at(vi); }
    triangle_type& mutable_triangle(std::size_t f) { return triangles_.at(f); }
    std::vector<index_type> const& surface_to_tetrahedral_mesh_index_map() const { return to_tet_; }
    index_type from_surface_vertex(std::size_t vi) const { return to_tet_.at(vi); }
    // tetrahedral_mesh_boundary.cpp:170-208: (x, y, z, nx, ny, nz, r, g, b) per vertex, three indices per triangle
    void prepare_vertices_for_rendering() override
    {
        std::vector<float> buffer;
        buffer.reserve(9 * vertices_.size());
        for (vertex_type const& v : vertices_)
        {
            for (int d = 0; d < 3; ++d)
                buffer.push_back(static_cast<float>(v.position[d]));
            for (int d = 0; d < 3; ++d)
                buffer.push_back(static_cast<float>(v.normal[d]));
            for (int d = 0; d < 3; ++d)
                buffer.push_back(v.color[static_cast<std::size_t>(d)]);
        }
        transfer_vertices_for_rendering(std::move(buffer));
    }
    void prepare_indices_for_rendering() override
    {
        std::vector<std::uint32_t> buffer;
        buffer.reserve(3 * triangles_.size());
        for (triangle_type const& t : triangles_)
            buffer.insert(buffer.end(), t.vertices.begin(), t.vertices.end());
        transfer_indices_for_rendering(std::move(buffer));
    }

  private:
    friend class tetrahedral_body_t;
    std::vector<vertex_type> vertices_;
    std::vector<triangle_type> triangles_;
    std::vector<index_type> to_tet_;
};

class body_t
{
  public:
    using collision_model_type = collision::collision_model_t;
    using visual_model_type    = common::shared_vertex_surface_mesh_i;
    body_t(simulation_t& simulation, index_type id) : id_(id), simulation_(simulation) {}
    virtual ~body_t() = default;
    virtual collision_model_type const& collision_model() const = 0;
    virtual collision_model_type& collision_model()             = 0;
    virtual void transform(affine3 const& affine)               = 0;
    index_type id() const { return id_; }
    simulation_t const& simulation() const { return simulation_; }

  protected:
    simulation_t& simulation() { return simulation_; }

  private:
    index_type id_;
    simulation_t& simulation_;
};

struct tetrahedron_t
{
    index_type v[4];
    index_type v1() const { return v[0]; }
    index_type v2() const { return v[1]; }
    index_type v3() const { return v[2]; }
    index_type v4() const { return v[3]; }
};

// the part of tetrahedron_set_t (include/sbs/physics/topology.h) main.cpp:37 iterates
class tetrahedron_set_t
{
  public:
    std::vector<tetrahedron_t> const& tetrahedra() const { return tets_; }
    std::vector<tetrahedron_t>& tetrahedra() { return tets_; }
    std::size_t tetrahedron_count() const { return tets_.size(); }

  private:
    std::vector<tetrahedron_t> tets_;
};

// include/sbs/physics/simulation.h:15-45
class simulation_t
{
  public:
    simulation_t() = default;
    simulation_t(simulation_t const&) = delete;
    simulation_t& operator=(simulation_t const&) = delete;
    ~simulation_t()
    {
        if (ctx_)
            sbsb200_destroy(ctx_);
    }

    void use_collision_detection_system(std::unique_ptr<collision::cd_system_t> cd_system)
    {
        cd_system_ = std::move(cd_system);
        invalidate();
    }
    void add_particle(particle_t const& p, index_type const body_idx)
    {
        particles_.at(body_idx).push_back(p);
        invalidate();
    }
    void add_body(std::unique_ptr<body_t> body)
    {
        particles_.emplace_back();
        bodies_.push_back(std::move(body));
        invalidate();
    }
    void add_body() // simulation.cpp:12-15: the slot must exist before a body constructor fills it
    {
        particles_.emplace_back();
        bodies_.emplace_back();
        invalidate();
    }
    void add_constraint(std::unique_ptr<constraint_t> constraint)
    {
        constraints_.push_back(std::move(constraint));
        invalidate();
    }
    // simulation.cpp:34-39: swap with the last.  A Green constraint of a built device scene is taken out in place
    // (sbsb200_remove_constraints: no re-planning); anything else rebuilds the device scene at the next step.
    void remove_constraint(index_type constraint_idx);

    std::vector<std::vector<particle_t>> const& particles() const
    {
        const_cast<simulation_t*>(this)->refresh_host();
        return particles_;
    }
    std::vector<std::vector<particle_t>>& particles()
    {
        refresh_host();
        host_written_ = true;
        return particles_;
    }
    std::vector<std::unique_ptr<body_t>> const& bodies() const { return bodies_; }
    std::vector<std::unique_ptr<body_t>>& bodies() { return bodies_; } // replaced bodies are noticed at the next step
    std::vector<std::unique_ptr<constraint_t>> const& constraints() const { return constraints_; }
    // simulation.h:31: handed out mutable, the list may change behind the facade's back — if it did, the device scene
    // is rebuilt from it at the next step (add_constraint / remove_constraint are the cheap ways)
    std::vector<std::unique_ptr<constraint_t>>& constraints()
    {
        constraints_handed_out_ = true; // compared with what the device scene was built from at the next step
        return constraints_;
    }
    // simulation.h:24, :32-33.  The reference's contact handler fills this list at every detection and
    // timestep_t::step clears it before returning (timestep.cpp:68), so between steps it is empty there too; on the
    // device the contacts are a compacted list the step consumes (read them with sbsb200_get_contacts).  Adding a
    // collision constraint by hand has no device counterpart.
    std::vector<std::unique_ptr<constraint_t>> const& collision_constraints() const { return collision_constraints_; }
    void add_collision_constraint(std::unique_ptr<constraint_t>)
    {
        throw std::logic_error("sbs-b200: collision constraints are created by the detection on the device "
                               "(timestep_t::step); there is no CPU path");
    }
    std::unique_ptr<collision::cd_system_t> const& collision_detection_system() const { return cd_system_; }
    xpbd::simulation_parameters_t const& simulation_parameters() const { return simulation_parameters_; }
    xpbd::simulation_parameters_t& simulation_parameters() { return simulation_parameters_; }

    // ---- device side (used by timestep_t; not part of the reference API) --------------------
    int device = 0;                    // CUDA device of this simulation
    int precision = SBSB200_FP32;      // SBSB200_FP64 = validation build
    int detect_mode = SBSB200_DETECT_PER_FRAME; // timestep.cpp:29-30
    sbsb200_ctx* context() { return ctx_; }
    // the device scene (built now if the description changed) and the device index of a body (-1: not on the device)
    sbsb200_ctx* ensure_device()
    {
        if (dirty_ || !ctx_)
            build_device();
        return ctx_;
    }
    int device_body(index_type body) const { return device_body_.at(body); }
    // what the reference hands to contact_handler_t::handle at a detection (bvh_model.cpp:66-96): the contacts of the
    // most recent detection on the device (sbsb200_get_contacts), bodies as simulation body indices, vi as surface vertex
    std::vector<collision::surface_mesh_particle_to_sdf_contact_t> contacts();
    // sdf_model_t::evaluate of an environment body as the detection samples it (sbsb200_eval_sdf)
    std::pair<scalar_type, vec3> evaluate_sdf(index_type body, vec3 const& p);
    void device_step(scalar_type dt, std::size_t substeps, std::size_t iterations);
    // the serial constraint order equivalent to the device schedule (sbsb200_get_constraint_order)
    std::vector<index_type> device_constraint_order();

  private:
    void invalidate() { dirty_ = true; }
    void check(int rc, char const* what)
    {
        if (rc < 0)
            throw b200::error(rc, std::string(what) + ": " + sbsb200_last_error(ctx_));
    }
    void build_device();
    void refresh_host();
    void push_host();

    std::vector<std::vector<particle_t>> particles_;
    std::vector<std::unique_ptr<body_t>> bodies_;
    std::vector<std::unique_ptr<constraint_t>> constraints_;
    std::vector<std::unique_ptr<constraint_t>> collision_constraints_; // always empty between steps (timestep.cpp:68)
    bool constraints_handed_out_ = false; // constraints() was handed out mutable since the last step
    std::unique_ptr<collision::cd_system_t> cd_system_;
    xpbd::simulation_parameters_t simulation_parameters_;
    sbsb200_ctx* ctx_ = nullptr;
    std::vector<int> device_body_;  // simulation body index -> device body index (-1: not on the device)
    std::vector<body_t const*> built_bodies_; // the bodies the device scene was built from
    std::vector<constraint_t const*> device_constraint_; // device insertion index -> constraint (null: removed in place)
    std::vector<std::vector<double>> device_mass_; // per body: the masses the device scene holds
    scalar_type device_collision_compliance_ = -1; // the value last pushed to the device
    bool dirty_        = true;      // scene description changed since the device scene was built
    bool host_stale_   = false;     // the device stepped since particles_ was last refreshed
    bool host_written_ = false;     // particles_ was handed out mutable since the last upload
};

// include/sbs/physics/tetrahedral_body.h, src/physics/tetrahedral_body.cpp:29-83,121-132
class tetrahedral_body_t : public body_t
{
  public:
    tetrahedral_body_t(simulation_t& simulation, index_type id, common::geometry_t const& geometry)
        : body_t(simulation, id)
    {
        if (!geometry.is_tetrahedral_mesh() || !geometry.has_positions() || !geometry.has_indices())
            throw std::invalid_argument("tetrahedral_body_t needs a tetrahedral geometry with positions and indices");
        for (std::size_t i = 0; i + 2 < geometry.positions.size(); i += 3)
            simulation.add_particle(
                particle_t(vec3(geometry.positions[i], geometry.positions[i + 1], geometry.positions[i + 2])), id);
        for (std::size_t i = 0; i + 3 < geometry.indices.size(); i += 4)
            physical_model_.tetrahedra().push_back(tetrahedron_t{
                {static_cast<index_type>(geometry.indices[i]), static_cast<index_type>(geometry.indices[i + 1]),
                 static_cast<index_type>(geometry.indices[i + 2]), static_cast<index_type>(geometry.indices[i + 3])}});
        collision_model_.id() = id;
        colors_               = geometry.colors;
    }
    collision_model_type const& collision_model() const override { return collision_model_; }
    collision_model_type& collision_model() override { return collision_model_; }
    // tetrahedral_body.cpp:85-119, :157-165.  The boundary lives on the device: it is fetched when first asked for
    // (which builds the device scene if need be), update_visual_model() refreshes positions and normals from the
    // surface copy of the last step.  Colours follow the reference: surface vertex i takes colour i of the geometry
    // (tetrahedral_body.cpp:66-78).
    tetrahedral_mesh_boundary_t const& surface_mesh()
    {
        fetch_boundary();
        return visual_model_;
    }
    visual_model_type& visual_model()
    {
        fetch_boundary();
        return visual_model_;
    }
    void update_visual_model()
    {
        fetch_boundary();
        sbsb200_ctx* ctx = simulation().ensure_device();
        int const db     = simulation().device_body(id());
        std::vector<float> out(6 * std::max<std::size_t>(1, visual_model_.vertices_.size()));
        int const rc = sbsb200_download_surface(ctx, db, out.data());
        if (rc < 0)
            throw b200::error(rc, std::string("sbsb200_download_surface: ") + sbsb200_last_error(ctx));
        for (std::size_t i = 0; i < visual_model_.vertices_.size(); ++i)
        {
            visual_model_.vertices_[i].position = vec3(out[6 * i], out[6 * i + 1], out[6 * i + 2]);
            visual_model_.vertices_[i].normal   = vec3(out[6 * i + 3], out[6 * i + 4], out[6 * i + 5]);
        }
        visual_model_.mark_vertices_dirty();
    }
    void update_collision_model() {} // the sphere tree is refitted on the device at every detection (bvh.cuh)
    void update_physical_model() {}  // tetrahedral_body.cpp:116-119: no-op
    void transform(affine3 const& affine) override // tetrahedral_body.cpp:121-132: x0, xi, xn and x
    {
        for (particle_t& p : simulation().particles().at(id()))
        {
            p.x0() = affine * p.x0();
            p.xi() = affine * p.xi();
            p.xn() = affine * p.xn();
            p.x()  = affine * p.x();
        }
    }
    tetrahedron_set_t const& physical_model() const { return physical_model_; }

  private:
    void fetch_boundary()
    {
        if (fetched_)
            return;
        sbsb200_ctx* ctx = simulation().ensure_device();
        int const db     = simulation().device_body(id());
        std::int64_t const nv = sbsb200_get_surface_map(ctx, db, nullptr, 0);
        std::int64_t const ni = sbsb200_get_surface_triangles(ctx, db, nullptr, 0);
        if (nv < 0 || ni < 0)
            throw b200::error(static_cast<int>(nv < 0 ? nv : ni), "sbs-b200: no boundary for this body on the device");
        visual_model_.to_tet_.resize(static_cast<std::size_t>(nv));
        std::vector<std::uint32_t> tri(static_cast<std::size_t>(ni));
        sbsb200_get_surface_map(ctx, db, visual_model_.to_tet_.data(), nv);
        sbsb200_get_surface_triangles(ctx, db, tri.data(), ni);
        visual_model_.triangles_.resize(tri.size() / 3);
        for (std::size_t f = 0; f < visual_model_.triangles_.size(); ++f)
            visual_model_.triangles_[f].vertices = {{tri[3 * f], tri[3 * f + 1], tri[3 * f + 2]}};
        visual_model_.vertices_.assign(static_cast<std::size_t>(nv), tetrahedral_mesh_boundary_t::vertex_type{});
        for (std::size_t i = 0; i < visual_model_.vertices_.size() && 3 * i + 2 < colors_.size(); ++i)
            visual_model_.vertices_[i].color = {{colors_[3 * i] / 255.f, colors_[3 * i + 1] / 255.f, colors_[3 * i + 2] / 255.f}};
        visual_model_.set_as_physically_simulated_body();
        fetched_ = true;
    }
    tetrahedron_set_t physical_model_;
    tetrahedral_mesh_boundary_t visual_model_;
    std::vector<std::uint8_t> colors_;
    bool fetched_ = false;
    collision::point_bvh_model_t collision_model_;
};

// include/sbs/physics/environment_body.h, src/physics/environment_body.cpp:12-88
class environment_body_t : public body_t
{
  public:
    // environment_body.cpp:12-78: triangle mesh -> grid SDF over the extended domain
    environment_body_t(simulation_t& simulation, index_type id, common::geometry_t const& geometry,
                       aligned_box3 const& domain, std::array<unsigned int, 3u> const& resolution = {10u, 10u, 10u})
        : body_t(simulation, id),
          collision_model_(collision::sdf_model_t::from_mesh(
              std::vector<scalar_type>(geometry.positions.begin(), geometry.positions.end()),
              std::vector<std::uint32_t>(geometry.indices.begin(), geometry.indices.end()), domain, resolution))
    {
        if (geometry.geometry_type != common::geometry_t::geometry_type_t::triangle || !geometry.has_positions() ||
            !geometry.has_indices())
            throw std::invalid_argument("sbs-b200: environment_body_t needs an indexed triangle mesh"); // :27-29
        collision_model_.id() = id;
    }
    environment_body_t(simulation_t& simulation, index_type id, common::geometry_t const& /*visual*/,
                       collision::sdf_model_t const& sdf_model)
        : body_t(simulation, id), collision_model_(sdf_model)
    {
        collision_model_.id() = id;
    }
    collision_model_type const& collision_model() const override { return collision_model_; }
    collision_model_type& collision_model() override { return collision_model_; }
    void transform(affine3 const&) override {}
    collision::sdf_model_t const& sdf() const { return collision_model_; }

  private:
    collision::sdf_model_t collision_model_;
};

namespace xpbd {

// include/sbs/physics/xpbd/green_constraint.h, src/physics/xpbd/green_constraint.cpp:11-47
class green_constraint_t : public constraint_t
{
  public:
    green_constraint_t(scalar_type const alpha, scalar_type const beta, simulation_t const&, index_type bi,
                       index_type v1, index_type v2, index_type v3, index_type v4, scalar_type young_modulus,
                       scalar_type poisson_ratio)
        : constraint_t(alpha, beta), bi_(bi), v_{v1, v2, v3, v4}, E_(young_modulus), nu_(poisson_ratio)
    {
    }
    index_type body() const { return bi_; }
    index_type const* vertices() const { return v_; }
    scalar_type young_modulus() const { return E_; }
    scalar_type poisson_ratio() const { return nu_; }

  private:
    index_type bi_, v_[4];
    scalar_type E_, nu_;
};

// include/sbs/physics/xpbd/distance_constraint.h, src/physics/xpbd/distance_constraint.cpp:8-22
class distance_constraint_t : public constraint_t
{
  public:
    distance_constraint_t(scalar_type const alpha, scalar_type const beta, simulation_t const&, index_type b1,
                          index_type b2, index_type v1, index_type v2)
        : constraint_t(alpha, beta), b1_(b1), b2_(b2), v1_(v1), v2_(v2)
    {
    }
    index_type b1() const { return b1_; }
    index_type b2() const { return b2_; }
    index_type v1() const { return v1_; }
    index_type v2() const { return v2_; }

  private:
    index_type b1_, b2_, v1_, v2_;
};

} // namespace xpbd

// include/sbs/physics/solver.h:13-17, gauss_seidel_solver.h.  The sweep runs on the device inside
// timestep_t::step (colour-major Gauss-Seidel, collisions first: gauss_seidel_solver.cpp:25-36).
class solver_t
{
  public:
    virtual ~solver_t() = default;
    virtual void solve(simulation_t& simulation, scalar_type dt, std::size_t iterations) = 0;
};
class gauss_seidel_solver_t : public solver_t
{
  public:
    void solve(simulation_t&, scalar_type, std::size_t) override
    {
        throw std::logic_error("sbs-b200: the Gauss-Seidel sweep runs on the device inside timestep_t::step "
                               "(predict and commit are inline in it); there is no CPU path");
    }
};

// include/sbs/physics/timestep.h:14-39, src/physics/timestep.cpp:20-70
class timestep_t
{
  public:
    timestep_t() = default;
    timestep_t(scalar_type const dt, std::size_t const iterations, std::size_t const substeps)
        : dt_(dt), iterations_(iterations), substeps_(substeps)
    {
    }
    void step(simulation_t& simulation) { simulation.device_step(dt_, substeps_, iterations_); }
    scalar_type dt() const { return dt_; }
    scalar_type& dt() { return dt_; }
    std::size_t iterations() const { return iterations_; }
    std::size_t& iterations() { return iterations_; }
    std::size_t substeps() const { return substeps_; }
    std::size_t& substeps() { return substeps_; }
    std::unique_ptr<solver_t> const& solver() const { return solver_; }
    std::unique_ptr<solver_t>& solver() { return solver_; }

  private:
    scalar_type dt_{0.};
    std::size_t iterations_{0u};
    std::size_t substeps_{0u};
    std::unique_ptr<solver_t> solver_{};
};

// ---- simulation_t: device plumbing ---------------------------------------------------------------

inline void simulation_t::remove_constraint(index_type const constraint_idx)
{
    constraint_t const* gone = constraints_.at(constraint_idx).get();
    bool in_place            = false;
    if (!dirty_ && ctx_ && dynamic_cast<xpbd::green_constraint_t const*>(gone))
    {
        auto const it = std::find(device_constraint_.begin(), device_constraint_.end(), gone);
        if (it != device_constraint_.end())
        {
            std::uint32_t const id = static_cast<std::uint32_t>(it - device_constraint_.begin());
            in_place               = sbsb200_remove_constraints(ctx_, 1, &id) == SBSB200_OK;
            if (in_place)
                *it = nullptr;
        }
    }
    std::swap(constraints_.at(constraint_idx), constraints_.back());
    constraints_.pop_back();
    if (!in_place)
        invalidate();
}

inline void simulation_t::build_device()
{
    if (ctx_)
    {
        sbsb200_destroy(ctx_);
        ctx_ = nullptr;
    }
    int const rc = sbsb200_create(device, precision, &ctx_);
    if (rc < 0)
        throw b200::error(rc, std::string("sbsb200_create: ") + sbsb200_last_error(nullptr));
    check(sbsb200_set_collision_compliance(ctx_, simulation_parameters_.collision_compliance), "collision compliance");
    bool const collide = static_cast<bool>(cd_system_);
    device_body_.assign(bodies_.size(), -1);
    for (std::size_t b = 0; b < bodies_.size(); ++b)
    {
        if (auto const* tb = dynamic_cast<tetrahedral_body_t const*>(bodies_[b].get()))
        {
            std::vector<particle_t> const& ps = particles_[b];
            std::vector<double> x0(3 * ps.size()), mass(ps.size());
            for (std::size_t i = 0; i < ps.size(); ++i)
            {
                for (int d = 0; d < 3; ++d)
                    x0[3 * i + d] = ps[i].x0()[d];
                mass[i] = ps[i].mass();
            }
            // the green constraints of this body, in insertion order; one material per body
            std::vector<std::uint32_t> tets;
            double E = simulation_parameters_.young_modulus, nu = simulation_parameters_.poisson_ratio,
                   alpha = simulation_parameters_.compliance, beta = simulation_parameters_.damping;
            bool first = true;
            for (auto const& c : constraints_)
                if (auto const* g = dynamic_cast<xpbd::green_constraint_t const*>(c.get()))
                    if (g->body() == b)
                    {
                        if (!first && (g->young_modulus() != E || g->poisson_ratio() != nu || g->alpha() != alpha ||
                                       g->beta() != beta))
                            throw std::invalid_argument("sbs-b200: one material per tetrahedral body");
                        first = false;
                        E     = g->young_modulus();
                        nu    = g->poisson_ratio();
                        alpha = g->alpha();
                        beta  = g->beta();
                        tets.insert(tets.end(), g->vertices(), g->vertices() + 4);
                    }
            (void)tb;
            int const db = sbsb200_add_tet_body(ctx_, static_cast<std::int64_t>(ps.size()), x0.data(), mass.data(),
                                                static_cast<std::int64_t>(tets.size() / 4), tets.data(), E, nu, alpha,
                                                beta);
            check(db, "sbsb200_add_tet_body");
            device_body_[b] = db;
        }
        else if (auto const* eb = dynamic_cast<environment_body_t const*>(bodies_[b].get()))
        {
            if (!collide)
                continue;
            collision::sdf_model_t const& m = eb->sdf();
            double const vol[6] = {m.volume().min()[0], m.volume().min()[1], m.volume().min()[2],
                                   m.volume().max()[0], m.volume().max()[1], m.volume().max()[2]};
            int db              = -1;
            if (m.shape() == collision::sdf_model_t::shape_t::plane)
            { // point on the plane: -offset * n / |n|^2
                double const nn    = m.a().dot(m.a());
                double const n[3]  = {m.a()[0], m.a()[1], m.a()[2]};
                double const pt[3] = {-m.d() * n[0] / nn, -m.d() * n[1] / nn, -m.d() * n[2] / nn};
                db                 = sbsb200_add_sdf_plane(ctx_, n, pt, vol);
            }
            else if (m.shape() == collision::sdf_model_t::shape_t::sphere)
            {
                double const c[3] = {m.a()[0], m.a()[1], m.a()[2]};
                db                = sbsb200_add_sdf_sphere(ctx_, c, m.d(), vol);
            }
            else if (m.shape() == collision::sdf_model_t::shape_t::grid)
            {
                double const lo[3] = {m.a()[0], m.a()[1], m.a()[2]}, hi[3] = {m.b()[0], m.b()[1], m.b()[2]};
                std::uint32_t const res[3] = {m.resolution()[0], m.resolution()[1], m.resolution()[2]};
                db = sbsb200_add_sdf_grid(ctx_, lo, hi, res, m.values().data(),
                                          static_cast<std::int64_t>(m.values().size()), vol);
            }
            else if (m.shape() == collision::sdf_model_t::shape_t::mesh)
            {
                aligned_box3 const& g = m.given_domain();
                double const dom[6]   = {g.min()[0], g.min()[1], g.min()[2], g.max()[0], g.max()[1], g.max()[2]};
                std::uint32_t const res[3] = {m.resolution()[0], m.resolution()[1], m.resolution()[2]};
                db = sbsb200_add_sdf_mesh(ctx_, static_cast<std::int64_t>(m.values().size() / 3), m.values().data(),
                                          static_cast<std::int64_t>(m.triangles().size() / 3), m.triangles().data(),
                                          dom, res);
            }
            else
            {
                double const lo[3] = {m.a()[0], m.a()[1], m.a()[2]}, hi[3] = {m.b()[0], m.b()[1], m.b()[2]};
                db = sbsb200_add_sdf_box(ctx_, lo, hi, vol);
            }
            check(db, "sbsb200_add_sdf_*");
            device_body_[b] = db;
        }
    }
    // only the collision models handed to the cd system take part in detection (brute_force_cd_system.cpp:8-12)
    if (collide)
        for (std::size_t b = 0; b < bodies_.size(); ++b)
        {
            if (device_body_[b] < 0)
                continue;
            auto const& listed = cd_system_->collision_objects();
            bool const takes_part =
                std::find(listed.begin(), listed.end(), &bodies_[b]->collision_model()) != listed.end();
            if (!takes_part)
                check(sbsb200_set_body_collideable(ctx_, device_body_[b], 0), "sbsb200_set_body_collideable");
        }
    // Note: the C ABI numbers constraints per call (all tets of a body, then each batch of distance
    // constraints); device_constraint_order() maps back to positions in constraints().
    for (auto const& c : constraints_)
        if (auto const* d = dynamic_cast<xpbd::distance_constraint_t const*>(c.get()))
        {
            std::uint32_t const pair[2] = {d->v1(), d->v2()};
            check(sbsb200_add_distance_constraints(ctx_, device_body_.at(d->b1()), device_body_.at(d->b2()), 1, pair,
                                                   d->alpha(), d->beta()),
                  "sbsb200_add_distance_constraints");
        }
    check(sbsb200_finalize(ctx_), "sbsb200_finalize");
    // device numbering: the green constraints body by body, then the distance constraints
    device_constraint_.clear();
    for (std::size_t b = 0; b < bodies_.size(); ++b)
        if (device_body_[b] >= 0 && dynamic_cast<tetrahedral_body_t const*>(bodies_[b].get()))
            for (auto const& c : constraints_)
                if (auto const* g = dynamic_cast<xpbd::green_constraint_t const*>(c.get()))
                    if (g->body() == b)
                        device_constraint_.push_back(c.get());
    for (auto const& c : constraints_)
        if (dynamic_cast<xpbd::distance_constraint_t const*>(c.get()))
            device_constraint_.push_back(c.get());
    // the device scene was built from the particles' masses: nothing to push until the caller changes one
    device_mass_.assign(bodies_.size(), {});
    for (std::size_t b = 0; b < bodies_.size(); ++b)
        if (device_body_[b] >= 0 && dynamic_cast<tetrahedral_body_t const*>(bodies_[b].get()))
            for (particle_t const& p : particles_[b])
                device_mass_[b].push_back(p.mass());
    device_collision_compliance_ = simulation_parameters_.collision_compliance;
    dirty_        = false;
    host_stale_   = false;
    host_written_ = true; // positions/velocities of the host mirror go to the device before the first step
}

inline void simulation_t::push_host()
{
    for (std::size_t b = 0; b < bodies_.size(); ++b)
    {
        if (device_body_[b] < 0 || !dynamic_cast<tetrahedral_body_t const*>(bodies_[b].get()))
            continue;
        std::vector<particle_t> const& ps = particles_[b];
        std::vector<double> x(3 * ps.size()), v(3 * ps.size());
        for (std::size_t i = 0; i < ps.size(); ++i)
            for (int d = 0; d < 3; ++d)
            {
                x[3 * i + d] = ps[i].x()[d];
                v[3 * i + d] = ps[i].v()[d];
            }
        if (!ps.empty())
            check(sbsb200_upload(ctx_, device_body_[b], x.data(), v.data()), "sbsb200_upload");
        // masses: only the ones the caller changed since they last went to the device (main.cpp:158-165 toggles a
        // picked particle between 1 and 0), in one call
        std::vector<double>& known = device_mass_[b];
        std::vector<std::uint32_t> which;
        std::vector<double> mass;
        for (std::size_t i = 0; i < ps.size(); ++i)
            if (i >= known.size() || known[i] != ps[i].mass())
            {
                which.push_back(static_cast<std::uint32_t>(i));
                mass.push_back(ps[i].mass());
            }
        if (!which.empty())
            check(sbsb200_set_masses(ctx_, device_body_[b], static_cast<std::int64_t>(which.size()), which.data(),
                                     mass.data()),
                  "sbsb200_set_masses");
        known.resize(ps.size());
        for (std::size_t i = 0; i < ps.size(); ++i)
            known[i] = ps[i].mass();
    }
    host_written_ = false;
}

inline void simulation_t::refresh_host()
{
    if (!host_stale_ || !ctx_)
        return;
    for (std::size_t b = 0; b < bodies_.size(); ++b)
    {
        if (device_body_[b] < 0 || !dynamic_cast<tetrahedral_body_t const*>(bodies_[b].get()))
            continue;
        std::vector<particle_t>& ps = particles_[b];
        std::vector<double> x(3 * ps.size()), v(3 * ps.size());
        if (ps.empty())
            continue;
        check(sbsb200_download(ctx_, device_body_[b], x.data(), v.data()), "sbsb200_download");
        for (std::size_t i = 0; i < ps.size(); ++i)
        {
            vec3 const p(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
            ps[i].x()  = p; // timestep.cpp:48-57 leaves x = xi = xn, f = 0
            ps[i].xi() = p;
            ps[i].xn() = p;
            ps[i].v()  = vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
            ps[i].f().setZero();
        }
    }
    host_stale_ = false;
}

inline void simulation_t::device_step(scalar_type dt, std::size_t substeps, std::size_t iterations)
{
    std::vector<body_t const*> now;
    for (auto const& b : bodies_)
        now.push_back(b.get());
    if (constraints_handed_out_ && !dirty_)
    { // the caller had the list in hand: still the constraints the device scene holds, in any order?
        std::vector<constraint_t const*> have, built;
        for (auto const& c : constraints_)
            have.push_back(c.get());
        for (constraint_t const* c : device_constraint_)
            if (c)
                built.push_back(c);
        std::sort(have.begin(), have.end());
        std::sort(built.begin(), built.end());
        if (have != built)
            invalidate();
    }
    constraints_handed_out_ = false;
    if (dirty_ || !ctx_ || now != built_bodies_)
    {
        refresh_host();
        build_device();
        built_bodies_ = now;
    }
    if (host_written_)
        push_host();
    // xpbd/contact_handler.cpp:42-52 reads the compliance whenever it creates a collision constraint, i.e. at every
    // detection: a value changed through simulation_parameters() applies from the next step on
    if (device_collision_compliance_ != simulation_parameters_.collision_compliance)
    {
        check(sbsb200_set_collision_compliance(ctx_, simulation_parameters_.collision_compliance), "collision compliance");
        device_collision_compliance_ = simulation_parameters_.collision_compliance;
    }
    check(sbsb200_step(ctx_, dt, static_cast<int>(substeps), static_cast<int>(iterations), detect_mode),
          "sbsb200_step");
    host_stale_ = true;
}

inline std::vector<collision::surface_mesh_particle_to_sdf_contact_t> simulation_t::contacts()
{
    std::vector<collision::surface_mesh_particle_to_sdf_contact_t> out;
    if (dirty_ || !ctx_)
        return out; // nothing has been detected on a device scene yet
    std::int64_t const n = sbsb200_get_contacts(ctx_, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (n < 0)
        throw b200::error(static_cast<int>(n), std::string("sbsb200_get_contacts: ") + sbsb200_last_error(ctx_));
    std::vector<std::int32_t> body(static_cast<std::size_t>(n)), sdf(static_cast<std::size_t>(n));
    std::vector<std::uint32_t> vertex(static_cast<std::size_t>(n));
    std::vector<double> point(3 * static_cast<std::size_t>(n)), normal(3 * static_cast<std::size_t>(n));
    if (n > 0)
        check(static_cast<int>(sbsb200_get_contacts(ctx_, n, body.data(), vertex.data(), sdf.data(), point.data(), normal.data())),
              "sbsb200_get_contacts");
    std::vector<index_type> host_of_device;
    for (std::size_t b = 0; b < device_body_.size(); ++b)
        if (device_body_[b] >= 0)
        {
            host_of_device.resize(std::max<std::size_t>(host_of_device.size(), static_cast<std::size_t>(device_body_[b]) + 1), 0);
            host_of_device[static_cast<std::size_t>(device_body_[b])] = static_cast<index_type>(b);
        }
    std::unordered_map<index_type, std::unordered_map<index_type, index_type>> surface_of; // body -> tet vertex -> surface vertex
    for (std::int64_t i = 0; i < n; ++i)
    {
        index_type const b1 = host_of_device.at(static_cast<std::size_t>(body[static_cast<std::size_t>(i)]));
        index_type const b2 = host_of_device.at(static_cast<std::size_t>(sdf[static_cast<std::size_t>(i)]));
        auto it = surface_of.find(b1);
        if (it == surface_of.end())
        {
            it = surface_of.emplace(b1, std::unordered_map<index_type, index_type>{}).first;
            std::int64_t const ns = sbsb200_get_surface_map(ctx_, device_body_[b1], nullptr, 0);
            std::vector<std::uint32_t> map(static_cast<std::size_t>(std::max<std::int64_t>(ns, 0)));
            if (ns > 0)
                sbsb200_get_surface_map(ctx_, device_body_[b1], map.data(), ns);
            for (std::size_t sv = 0; sv < map.size(); ++sv)
                it->second.emplace(map[sv], static_cast<index_type>(sv));
        }
        std::size_t const k = 3 * static_cast<std::size_t>(i);
        out.emplace_back(collision::contact_t::type_t::surface_particle_to_sdf, b1, b2, vec3(point[k], point[k + 1], point[k + 2]),
                         vec3(normal[k], normal[k + 1], normal[k + 2]), it->second.at(vertex[static_cast<std::size_t>(i)]));
    }
    return out;
}

inline std::pair<scalar_type, vec3> simulation_t::evaluate_sdf(index_type body, vec3 const& p)
{
    sbsb200_ctx* ctx = ensure_device();
    double const pt[3] = {p[0], p[1], p[2]};
    double sd = 0., grad[3] = {0., 0., 0.};
    check(sbsb200_eval_sdf(ctx, device_body_.at(body), 1, pt, &sd, grad), "sbsb200_eval_sdf");
    return {sd, vec3(grad[0], grad[1], grad[2])};
}

inline std::vector<index_type> simulation_t::device_constraint_order()
{
    if (dirty_ || !ctx_)
        build_device();
    std::int64_t const n = sbsb200_constraint_count(ctx_);
    std::vector<index_type> order(static_cast<std::size_t>(std::max<std::int64_t>(n, 0)));
    if (n > 0)
        check(sbsb200_get_constraint_order(ctx_, order.data(), n), "sbsb200_get_constraint_order");
    // device insertion index -> constraint -> its position in constraints() now
    std::unordered_map<constraint_t const*, index_type> position;
    for (std::size_t i = 0; i < constraints_.size(); ++i)
        position.emplace(constraints_[i].get(), static_cast<index_type>(i));
    std::vector<index_type> device_to_host(device_constraint_.size(), 0);
    for (std::size_t i = 0; i < device_constraint_.size(); ++i)
        if (device_constraint_[i])
            device_to_host[i] = position.at(device_constraint_[i]);
    for (index_type& o : order)
        o = device_to_host.at(o);
    return order;
}

} // namespace physics
} // namespace sbs

#endif // SBS_B200_FACADE_HPP
