// The reference's demo set-up (main.cpp:17-130) written against the B200 facade headers: a
// tetrahedralised bar, transformed, one green constraint per tet, a plane-SDF floor, brute-force
// collision detection, timestep_t::step.  Writes x0, x, v of the bar (9 doubles per particle) to argv[7].
//   facade_demo W H D frames substeps iterations out.bin [precision]
#include <cstdio>
#include <cstdlib>
#include <iterator>

#include <sbs/geometry/get_simple_bar_model.h>
#include <sbs/physics/collision/brute_force_cd_system.h>
#include <sbs/physics/environment_body.h>
#include <sbs/physics/gauss_seidel_solver.h>
#include <sbs/physics/simulation.h>
#include <string>

#include <sbs/physics/tetrahedral_body.h>
#include <sbs/physics/timestep.h>
#include <sbs/physics/xpbd/contact_handler.h>
#include <sbs/physics/xpbd/distance_constraint.h>
#include <sbs/physics/xpbd/green_constraint.h>

int main(int argc, char** argv)
{
    if (argc < 8)
    {
        std::fprintf(stderr, "usage: %s W H D frames substeps iterations out.bin [32|64] [mesh|dynamic|remove|surface]\n", argv[0]);
        return 2;
    }
    std::size_t const W = std::atoi(argv[1]), H = std::atoi(argv[2]), D = std::atoi(argv[3]);
    int const frames = std::atoi(argv[4]);
    try
    {
        sbs::physics::simulation_t simulation{};
        if (argc > 8)
            simulation.precision = std::atoi(argv[8]);

        sbs::common::geometry_t beam_geometry = sbs::geometry::get_simple_bar_model(W, H, D);
        beam_geometry.set_color(255, 255, 0);
        auto const beam_idx = static_cast<sbs::index_type>(simulation.bodies().size());
        simulation.add_body();
        simulation.bodies()[beam_idx] =
            std::make_unique<sbs::physics::tetrahedral_body_t>(simulation, beam_idx, beam_geometry);
        sbs::physics::tetrahedral_body_t& beam =
            *dynamic_cast<sbs::physics::tetrahedral_body_t*>(simulation.bodies()[beam_idx].get());
        sbs::affine3 beam_transform = sbs::affine3::translation(-1., 0.4, -1.);
        beam_transform.rotate(0.3, sbs::vec3{0., 1., 0.2}.normalized());
        beam_transform.scale(sbs::vec3{1.0, 0.8, 2.});
        beam.transform(beam_transform);
        for (auto const& tetrahedron : beam.physical_model().tetrahedra())
        {
            auto const alpha = simulation.simulation_parameters().compliance;
            auto const beta  = simulation.simulation_parameters().damping;
            auto const nu    = simulation.simulation_parameters().poisson_ratio;
            auto const E     = simulation.simulation_parameters().young_modulus;
            simulation.add_constraint(std::make_unique<sbs::physics::xpbd::green_constraint_t>(
                alpha, beta, simulation, beam_idx, tetrahedron.v1(), tetrahedron.v2(), tetrahedron.v3(),
                tetrahedron.v4(), E, nu));
        }

        sbs::common::geometry_t floor_geometry; // visual only
        auto const floor_idx = static_cast<sbs::index_type>(simulation.bodies().size());
        sbs::aligned_box3 const floor_volume{sbs::vec3{-200., -5., -200.}, sbs::vec3{200., 5., 200.}};
        auto const floor_collision_model = sbs::physics::collision::sdf_model_t::from_plane(
            sbs::hyperplane3(sbs::vec3{0., 1., 0.}, sbs::vec3{0., 0., 0.}), floor_volume);
        simulation.add_body(std::make_unique<sbs::physics::environment_body_t>(simulation, floor_idx, floor_geometry,
                                                                               floor_collision_model));

        if (argc > 9 && std::string(argv[9]) == "mesh")
        { // a triangle-mesh obstacle: environment_body_t(simulation, id, geometry, domain, resolution)
            sbs::common::geometry_t rock; // octahedron poking up through the floor under the beam
            rock.geometry_type = sbs::common::geometry_t::geometry_type_t::triangle;
            rock.positions     = {1.5f, 0.f, 3.f, -1.5f, 0.f, 3.f, 0.f, 0.75f, 3.f, 0.f, -0.75f, 3.f, 0.f, 0.f, 6.f, 0.f, 0.f, 0.f};
            rock.indices       = {0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5};
            auto const rock_idx = static_cast<sbs::index_type>(simulation.bodies().size());
            sbs::aligned_box3 const rock_domain{sbs::vec3{-6., -3., -6.}, sbs::vec3{8., 8., 30.}};
            simulation.add_body(std::make_unique<sbs::physics::environment_body_t>(
                simulation, rock_idx, rock, rock_domain, std::array<unsigned int, 3u>{8u, 6u, 12u}));
        }

        std::vector<sbs::physics::collision::collision_model_t*> collision_objects{};
        std::transform(simulation.bodies().begin(), simulation.bodies().end(), std::back_inserter(collision_objects),
                       [](std::unique_ptr<sbs::physics::body_t>& b) { return &(b->collision_model()); });
        simulation.use_collision_detection_system(
            std::make_unique<sbs::physics::collision::brute_force_cd_system_t>(collision_objects));
        simulation.collision_detection_system()->use_contact_handler(
            std::make_unique<sbs::physics::xpbd::contact_handler_t>(simulation));

        sbs::physics::timestep_t timestep{};
        timestep.dt()         = 0.016;
        timestep.iterations() = std::atoi(argv[6]);
        timestep.substeps()   = std::atoi(argv[5]);
        timestep.solver()     = std::make_unique<sbs::physics::gauss_seidel_solver_t>();

        for (int f = 0; f < frames; ++f)
        {
            timestep.step(simulation);
            if (f == 0) // main.cpp:158-165: pin a picked vertex by setting its mass to 0 between frames
                simulation.particles()[beam_idx][0].mass() = 0.;
            if (f == 0 && argc > 9 && std::string(argv[9]) == "remove")
            { // Green constraints go between frames and nothing comes: the device scene is patched in place
              // (sbsb200_remove_constraints); the serial order before and after goes to <out>.order
                std::vector<sbs::index_type> const before = simulation.device_constraint_order();
                for (sbs::index_type const gone : {5u, 17u, 40u, 17u})
                    simulation.remove_constraint(gone); // simulation.cpp:34-39: the last one takes the place
                std::vector<sbs::index_type> const after = simulation.device_constraint_order();
                std::FILE* o = std::fopen((std::string(argv[7]) + ".order").c_str(), "wb");
                if (!o)
                    return 3;
                std::uint32_t const counts[2] = {static_cast<std::uint32_t>(before.size()),
                                                 static_cast<std::uint32_t>(after.size())};
                std::fwrite(counts, sizeof(std::uint32_t), 2, o);
                std::fwrite(before.data(), sizeof(sbs::index_type), before.size(), o);
                std::fwrite(after.data(), sizeof(sbs::index_type), after.size(), o);
                std::fclose(o);
            }
            if (f == 0 && argc > 9 && std::string(argv[9]) == "dynamic")
            { // constraints come and go between frames: simulation.cpp:29-39 (remove = swap with the last)
                simulation.remove_constraint(5);
                auto const last = static_cast<sbs::index_type>(simulation.particles()[beam_idx].size() - 1);
                simulation.add_constraint(std::make_unique<sbs::physics::xpbd::distance_constraint_t>(
                    1e-4, 0., simulation, beam_idx, beam_idx, 1, last));
            }
        }

        if (argc > 9 && std::string(argv[9]) == "surface")
        { // what the reference's renderer consumes (renderer.cpp:484-542): the body's visual model, refreshed from the
          // device, as the 9-float vertex buffer and the index buffer of tetrahedral_mesh_boundary.cpp:170-208
            beam.update_visual_model();
            sbs::common::shared_vertex_surface_mesh_i& mesh = beam.visual_model();
            mesh.prepare_vertices_for_rendering();
            mesh.prepare_indices_for_rendering();
            std::FILE* o = std::fopen((std::string(argv[7]) + ".surface").c_str(), "wb");
            if (!o)
                return 3;
            std::uint32_t const counts[2] = {static_cast<std::uint32_t>(mesh.vertex_count()),
                                             static_cast<std::uint32_t>(mesh.triangle_count())};
            std::fwrite(counts, sizeof(std::uint32_t), 2, o);
            std::fwrite(mesh.get_cpu_vertex_buffer().data(), sizeof(float), mesh.get_cpu_vertex_buffer().size(), o);
            std::fwrite(mesh.get_cpu_index_buffer().data(), sizeof(std::uint32_t), mesh.get_cpu_index_buffer().size(), o);
            std::fwrite(beam.surface_mesh().surface_to_tetrahedral_mesh_index_map().data(), sizeof(sbs::index_type),
                        mesh.vertex_count(), o);
            std::fclose(o);
            // the contacts of the last detection as contact_handler_t::handle would have seen them (contact.h:11-58),
            // and the floor's signed distance at a point, on the host (sdf_model.cpp:57-61) and as the device samples it
            auto const contacts = simulation.contacts();
            std::FILE* c = std::fopen((std::string(argv[7]) + ".contacts").c_str(), "wb");
            if (!c)
                return 3;
            for (auto const& k : contacts)
            {
                double const row[9] = {static_cast<double>(k.b1()), static_cast<double>(k.b2()), static_cast<double>(k.vi()),
                                       k.point().x(), k.point().y(), k.point().z(), k.normal().x(), k.normal().y(), k.normal().z()};
                std::fwrite(row, sizeof(double), 9, c);
            }
            sbs::vec3 const probe{0.3, 1.25, -2.};
            auto const& floor = *dynamic_cast<sbs::physics::environment_body_t*>(simulation.bodies()[floor_idx].get());
            auto const on_host   = floor.sdf().evaluate(probe);
            auto const on_device = simulation.evaluate_sdf(floor_idx, probe);
            double const tail[9] = {-1., on_host.first, on_device.first, on_host.second.x(), on_host.second.y(), on_host.second.z(),
                                    on_device.second.x(), on_device.second.y(), on_device.second.z()};
            std::fwrite(tail, sizeof(double), 9, c);
            std::fclose(c);
        }

        auto const& ps = static_cast<sbs::physics::simulation_t const&>(simulation).particles()[beam_idx];
        std::FILE* out = std::fopen(argv[7], "wb");
        if (!out)
            return 3;
        for (auto const& p : ps)
        {
            double const row[9] = {p.x0().x(), p.x0().y(), p.x0().z(), p.x().x(), p.x().y(),
                                   p.x().z(),  p.v().x(),  p.v().y(),  p.v().z()};
            std::fwrite(row, sizeof(double), 9, out);
        }
        std::fclose(out);
        std::printf("%zu particles, %zu constraints, %d frames\n", ps.size(), simulation.constraints().size(), frames);
    }
    catch (sbs::b200::error const& e)
    {
        std::fprintf(stderr, "sbs-b200 error %d: %s\n", e.code, e.what());
        return 10;
    }
    return 0;
}
