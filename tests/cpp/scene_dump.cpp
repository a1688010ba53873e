// scene_dump.cpp — loads a scene file through sbs::io::load_scene and prints everything it produced in a
// canonical text form.  ONE source, compiled twice: against the reference's own headers and sources
// (oracle/build_ref.sh -> oracle/_ref/ref_scene_dump, the checker) and against this repo's host side
// (soft-body-simulator_b200/cpp, the product).  tests/test_load_scene.py compares the two outputs; the
// reference's outputs on tests/golden/scenes/*.json are committed as tests/golden/scenes/*.dump.
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include <sbs/common/geometry.h>
#include <sbs/common/scene.h>
#include <sbs/io/load_scene.h>

namespace {

struct dump_node_t : sbs::common::renderable_node_t
{
    sbs::common::geometry_t geometry;
    bool physical       = false;
    double mass_density = 0., v[3] = {0., 0., 0.};
    void prepare_vertices_for_rendering() override {}
    void prepare_indices_for_rendering() override {}
};

} // namespace

int main(int argc, char** argv)
{
    if (argc < 2)
        return 2;
    std::vector<std::shared_ptr<dump_node_t>> made;
    sbs::common::scene_t const scene = sbs::io::load_scene(
        argv[1],
        [&](sbs::io::scene::scene_body_info const& info) -> std::shared_ptr<sbs::common::renderable_node_t> {
            auto n      = std::make_shared<dump_node_t>();
            n->geometry = info.geometry;
            made.push_back(n);
            return n;
        },
        [&](sbs::io::scene::physics_body_info const& info) -> std::shared_ptr<sbs::common::renderable_node_t> {
            auto n          = std::make_shared<dump_node_t>();
            n->geometry     = info.geometry;
            n->physical     = true;
            n->mass_density = info.mass_density;
            n->v[0]         = info.velocity.vx;
            n->v[1]         = info.velocity.vy;
            n->v[2]         = info.velocity.vz;
            made.push_back(n);
            return n;
        });
    auto const& d = scene.directional_light;
    auto const& p = scene.point_light;
    std::printf("nodes %zu\n", scene.nodes.size());
    if (!scene.nodes.empty() || argc > 2)
    {
        std::printf("directional %.9g %.9g %.9g | %.9g %.9g %.9g | %.9g %.9g %.9g | %.9g %.9g %.9g %.9g\n", d.dx, d.dy, d.dz,
                    d.ambient.r, d.ambient.g, d.ambient.b, d.diffuse.r, d.diffuse.g, d.diffuse.b, d.specular.r,
                    d.specular.g, d.specular.b, d.specular.exp);
        std::printf("point %.9g %.9g %.9g | %.9g %.9g %.9g | %.9g %.9g %.9g | %.9g %.9g %.9g %.9g | %.9g %.9g %.9g\n", p.x, p.y,
                    p.z, p.ambient.r, p.ambient.g, p.ambient.b, p.diffuse.r, p.diffuse.g, p.diffuse.b, p.specular.r,
                    p.specular.g, p.specular.b, p.specular.exp, p.attenuation.constant, p.attenuation.linear,
                    p.attenuation.quadratic);
    }
    for (std::size_t i = 0; i < scene.nodes.size(); ++i)
    {
        auto const& n = *scene.nodes[i];
        auto const& m = *made[i];
        std::printf("node %s %s %s %s\n", n.id().c_str(), n.is_environment_body() ? "environment" : "-",
                    n.is_physically_simulated_body() ? "physical" : "-", n.is_collideable_body() ? "collideable" : "-");
        std::printf("  geometry %s positions %zu indices %zu colors %zu normals %zu uvs %zu\n",
                    m.geometry.geometry_type == sbs::common::geometry_t::geometry_type_t::triangle ? "triangle" : "tetrahedron",
                    m.geometry.positions.size(), m.geometry.indices.size(), m.geometry.colors.size(),
                    m.geometry.normals.size(), m.geometry.uvs.size());
        if (m.physical)
            std::printf("  physics %.17g %.17g %.17g %.17g\n", m.mass_density, m.v[0], m.v[1], m.v[2]);
        std::printf("  positions");
        for (float x : m.geometry.positions)
            std::printf(" %a", static_cast<double>(x));
        std::printf("\n  indices");
        for (auto x : m.geometry.indices)
            std::printf(" %d", static_cast<int>(x));
        std::printf("\n  colors");
        for (auto x : m.geometry.colors)
            std::printf(" %d", static_cast<int>(x));
        std::printf("\n");
    }
    return 0;
}
