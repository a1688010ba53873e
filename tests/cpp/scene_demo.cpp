// scene_demo.cpp — a scene file as the front door of the GPU solver: io::load_scene reads the JSON and
// the PLY assets, the two factories create the bodies on a simulation_t (tetrahedral_body_t + Green
// constraints for `objects`; environment_body_t through its triangle-mesh constructor, i.e. a grid SDF
// baked on the device, for collideable `environment` bodies), then timestep_t::step runs on the GPU.
// Output: per simulated body, rows of (x0, x, v) doubles.  usage: scene_demo scene.json frames S K out.bin [32|64]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iterator>
#include <memory>
#include <vector>

#include <sbs/common/scene.h>
#include <sbs/io/load_scene.h>
#include <sbs/physics/collision/brute_force_cd_system.h>
#include <sbs/physics/environment_body.h>
#include <sbs/physics/gauss_seidel_solver.h>
#include <sbs/physics/simulation.h>
#include <sbs/physics/tetrahedral_body.h>
#include <sbs/physics/timestep.h>
#include <sbs/physics/xpbd/contact_handler.h>
#include <sbs/physics/xpbd/green_constraint.h>

namespace {

struct body_node_t : sbs::common::renderable_node_t
{
    sbs::index_type body = 0;
    void prepare_vertices_for_rendering() override {}
    void prepare_indices_for_rendering() override {}
};

} // namespace

int main(int argc, char** argv)
{
    if (argc < 6)
    {
        std::fprintf(stderr, "usage: %s scene.json frames substeps iterations out.bin [32|64]\n", argv[0]);
        return 2;
    }
    try
    {
        sbs::physics::simulation_t simulation{};
        if (argc > 6)
            simulation.precision = std::atoi(argv[6]);
        std::vector<sbs::index_type> simulated;

        auto const environment_factory = [&](sbs::io::scene::scene_body_info const& info) {
            auto node  = std::make_shared<body_node_t>();
            node->body = static_cast<sbs::index_type>(simulation.bodies().size());
            // the grid covers the mesh's bounding box grown by 1 on every side
            sbs::vec3 lo{1e30, 1e30, 1e30}, hi{-1e30, -1e30, -1e30};
            for (std::size_t i = 0; i + 2 < info.geometry.positions.size(); i += 3)
                for (int d = 0; d < 3; ++d)
                {
                    lo[d] = std::min<double>(lo[d], info.geometry.positions[i + d] - 1.);
                    hi[d] = std::max<double>(hi[d], info.geometry.positions[i + d] + 1.);
                }
            simulation.add_body(std::make_unique<sbs::physics::environment_body_t>(
                simulation, node->body, info.geometry, sbs::aligned_box3{lo, hi}, std::array<unsigned int, 3u>{8u, 8u, 8u}));
            return std::static_pointer_cast<sbs::common::renderable_node_t>(node);
        };
        auto const physics_factory = [&](sbs::io::scene::physics_body_info const& info) {
            auto node  = std::make_shared<body_node_t>();
            node->body = static_cast<sbs::index_type>(simulation.bodies().size());
            simulation.add_body();
            simulation.bodies()[node->body] =
                std::make_unique<sbs::physics::tetrahedral_body_t>(simulation, node->body, info.geometry);
            auto const& body =
                *dynamic_cast<sbs::physics::tetrahedral_body_t*>(simulation.bodies()[node->body].get());
            for (auto& p : simulation.particles()[node->body])
            {
                p.mass() = info.mass_density;
                p.v()    = sbs::vec3{info.velocity.vx, info.velocity.vy, info.velocity.vz};
            }
            auto const& prm = simulation.simulation_parameters();
            for (auto const& t : body.physical_model().tetrahedra())
                simulation.add_constraint(std::make_unique<sbs::physics::xpbd::green_constraint_t>(
                    prm.compliance, prm.damping, simulation, node->body, t.v1(), t.v2(), t.v3(), t.v4(),
                    prm.young_modulus, prm.poisson_ratio));
            simulated.push_back(node->body);
            return std::static_pointer_cast<sbs::common::renderable_node_t>(node);
        };
        sbs::common::scene_t const scene = sbs::io::load_scene(argv[1], environment_factory, physics_factory);
        if (scene.nodes.empty())
        {
            std::fprintf(stderr, "empty scene\n");
            return 4;
        }

        std::vector<sbs::physics::collision::collision_model_t*> collision_objects{};
        for (auto const& node : scene.nodes)
            if (node->is_collideable_body())
                collision_objects.push_back(
                    &simulation.bodies()[static_cast<body_node_t const&>(*node).body]->collision_model());
        simulation.use_collision_detection_system(
            std::make_unique<sbs::physics::collision::brute_force_cd_system_t>(collision_objects));
        simulation.collision_detection_system()->use_contact_handler(
            std::make_unique<sbs::physics::xpbd::contact_handler_t>(simulation));

        sbs::physics::timestep_t timestep{};
        timestep.dt()         = 0.016;
        timestep.substeps()   = std::atoi(argv[3]);
        timestep.iterations() = std::atoi(argv[4]);
        timestep.solver()     = std::make_unique<sbs::physics::gauss_seidel_solver_t>();
        for (int f = 0; f < std::atoi(argv[2]); ++f)
            timestep.step(simulation);

        std::FILE* out = std::fopen(argv[5], "wb");
        if (!out)
            return 3;
        for (sbs::index_type b : simulated)
            for (auto const& p : static_cast<sbs::physics::simulation_t const&>(simulation).particles()[b])
            {
                double const row[9] = {p.x0().x(), p.x0().y(), p.x0().z(), p.x().x(), p.x().y(),
                                       p.x().z(),  p.v().x(),  p.v().y(),  p.v().z()};
                std::fwrite(row, sizeof(double), 9, out);
            }
        std::fclose(out);
        std::printf("%zu nodes, %zu simulated bodies, %zu constraints\n", scene.nodes.size(), simulated.size(),
                    simulation.constraints().size());
    }
    catch (sbs::b200::error const& e)
    {
        std::fprintf(stderr, "sbs-b200 error %d: %s\n", e.code, e.what());
        return 10;
    }
    return 0;
}
