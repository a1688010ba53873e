// sbs/io/load_scene.h — the reference's scene-file front door (include/sbs/io/load_scene.h:15-36,
// src/io/load_scene.cpp:83-363) with the same signature: a JSON file naming two lights, `environment`
// bodies and simulated `objects`; every body's PLY asset (path relative to the scene file) is read, given
// the file's colour when it has none, rescaled into its `box` and moved by its `translation` (float
// arithmetic as in the reference, :11-82), then handed to the caller's factory, which returns the node
// (the factories are where a program creates its tetrahedral_body_t / environment_body_t on the GPU
// solver; see tests/cpp/scene_demo.cpp).  Bodies whose asset is missing or not a .ply are skipped; a path
// that does not exist or does not end in .json yields an empty scene.  Host-side only.
#ifndef SBS_IO_LOAD_SCENE_H
#define SBS_IO_LOAD_SCENE_H

#include <cmath>
#include <filesystem>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <string>

#include <sbs/common/geometry.h>
#include <sbs/common/scene.h>
#include <sbs/io/json.h>
#include <sbs/io/ply.h>

namespace sbs {
namespace io {
namespace scene {

struct scene_body_info
{
    std::string id;
    common::geometry_t geometry;
};

struct physics_body_info : scene_body_info
{
    double mass_density = 0.;
    struct velocity_t
    {
        double vx = 0., vy = 0., vz = 0.;
    } velocity;
};

} // namespace scene

namespace detail {

// load_scene.cpp:11-68: map the mesh's bounding box onto the given box, axis by axis; an axis along
// which the mesh is flat is left alone
inline void fit_into_box(common::geometry_t& g, double const lo[3], double const hi[3])
{
    double mn[3], mx[3];
    for (int d = 0; d < 3; ++d)
    {
        mn[d] = std::numeric_limits<float>::max();
        mx[d] = std::numeric_limits<float>::lowest();
    }
    for (std::size_t i = 0; i + 2 < g.positions.size(); i += 3)
        for (int d = 0; d < 3; ++d)
        {
            float const c = g.positions[i + static_cast<std::size_t>(d)];
            if (c < mn[d])
                mn[d] = c;
            if (c > mx[d])
                mx[d] = c;
        }
    for (std::size_t i = 0; i + 2 < g.positions.size(); i += 3)
        for (int d = 0; d < 3; ++d)
        {
            double const extent = mx[d] - mn[d];
            float& c            = g.positions[i + static_cast<std::size_t>(d)];
            if (!(std::abs(extent) < 1e-8))
                c = static_cast<float>(lo[d] + (hi[d] - lo[d]) * (c - mn[d]) / extent);
        }
}

// load_scene.cpp:70-82: the offset is rounded to float before it is added
inline void shift(common::geometry_t& g, double const t[3])
{
    for (std::size_t i = 0; i + 2 < g.positions.size(); i += 3)
        for (int d = 0; d < 3; ++d)
            g.positions[i + static_cast<std::size_t>(d)] += static_cast<float>(t[d]);
}

inline void read_xyz(json_value const& v, double out[3])
{
    out[0] = v["x"].get<double>();
    out[1] = v["y"].get<double>();
    out[2] = v["z"].get<double>();
}

template <typename L>
void read_phong(json_value const& spec, L& light)
{
    light.ambient.r    = spec["ambient"]["r"].get<float>();
    light.ambient.g    = spec["ambient"]["g"].get<float>();
    light.ambient.b    = spec["ambient"]["b"].get<float>();
    light.diffuse.r    = spec["diffuse"]["r"].get<float>();
    light.diffuse.g    = spec["diffuse"]["g"].get<float>();
    light.diffuse.b    = spec["diffuse"]["b"].get<float>();
    light.specular.r   = spec["specular"]["r"].get<float>();
    light.specular.g   = spec["specular"]["g"].get<float>();
    light.specular.b   = spec["specular"]["b"].get<float>();
    light.specular.exp = spec["specular"]["exp"].get<float>();
}

// what environment bodies and objects share (load_scene.cpp:182-236 and :267-323): asset, colour, box,
// translation.  nullopt = skip this body.
inline std::optional<common::geometry_t> read_body_geometry(std::filesystem::path const& scene_file,
                                                            json_value const& spec)
{
    std::filesystem::path const asset = scene_file.parent_path() / spec["geometry"]["path"].get<std::string>();
    (void)spec["geometry"]["type"].get<std::string>(); // required by the reference too, value unused (:184)
    if (!(asset.has_extension() && asset.extension() == ".ply" && std::filesystem::exists(asset) && asset.has_filename()))
        return std::nullopt;
    std::optional<common::geometry_t> g = read_ply(asset);
    if (!g.has_value())
        return std::nullopt;
    if (!g->has_colors())
        g->set_color(spec["color"]["r"].get<std::uint8_t>(), spec["color"]["g"].get<std::uint8_t>(),
                     spec["color"]["b"].get<std::uint8_t>());
    if (spec.contains("box"))
    {
        double lo[3], hi[3];
        read_xyz(spec["box"]["min"], lo);
        read_xyz(spec["box"]["max"], hi);
        fit_into_box(*g, lo, hi);
    }
    if (spec.contains("translation"))
    {
        double t[3];
        read_xyz(spec["translation"], t);
        shift(*g, t);
    }
    return g;
}

} // namespace detail

inline common::scene_t load_scene(
    std::filesystem::path const& path,
    std::function<std::shared_ptr<common::renderable_node_t>(scene::scene_body_info const&)> environment_body_factory,
    std::function<std::shared_ptr<common::renderable_node_t>(scene::physics_body_info const&)> physics_body_factory)
{
    if (!(std::filesystem::exists(path) && path.has_filename() && path.has_extension() &&
          path.extension().string() == ".json"))
        return {};
    std::ifstream ifs{path.string()};
    json_value const spec = parse_json(ifs);

    common::scene_t scene{};
    {
        json_value const& l = spec["lights"]["directional"];
        double d[3];
        detail::read_xyz(l["direction"], d);
        scene.directional_light.dx = static_cast<float>(d[0]);
        scene.directional_light.dy = static_cast<float>(d[1]);
        scene.directional_light.dz = static_cast<float>(d[2]);
        detail::read_phong(l, scene.directional_light);
    }
    {
        json_value const& l = spec["lights"]["point"];
        double p[3];
        detail::read_xyz(l["position"], p);
        scene.point_light.x = static_cast<float>(p[0]);
        scene.point_light.y = static_cast<float>(p[1]);
        scene.point_light.z = static_cast<float>(p[2]);
        detail::read_phong(l, scene.point_light);
        scene.point_light.attenuation.constant  = l["attenuation"]["constant"].get<float>();
        scene.point_light.attenuation.linear    = l["attenuation"]["linear"].get<float>();
        scene.point_light.attenuation.quadratic = l["attenuation"]["quadratic"].get<float>();
    }
    auto const finish = [&](std::shared_ptr<common::renderable_node_t> const& node, std::string const& id,
                            json_value const& body_spec) {
        node->set_id(id);
        if (body_spec.contains("collideable") && body_spec["collideable"].get<bool>())
            node->set_as_collideable_body();
        else
            node->set_as_non_collideable_body();
        scene.nodes.push_back(node);
    };
    for (json_value const& body_spec : spec["environment"])
    {
        std::optional<common::geometry_t> g = detail::read_body_geometry(path, body_spec);
        if (!g.has_value())
            continue;
        scene::scene_body_info info;
        info.id       = body_spec["id"].get<std::string>();
        info.geometry = std::move(*g);
        auto node     = environment_body_factory(info);
        node->set_as_environment_body();
        finish(node, info.id, body_spec);
    }
    for (json_value const& body_spec : spec["objects"])
    {
        // the physics block is read before the asset (load_scene.cpp:287-293): a body without one is an error
        // even when its asset would have been skipped... after the path checks, as in the reference
        std::filesystem::path const asset = path.parent_path() / body_spec["geometry"]["path"].get<std::string>();
        if (!(asset.has_extension() && asset.extension() == ".ply" && std::filesystem::exists(asset) && asset.has_filename()))
            continue;
        scene::physics_body_info info;
        (void)body_spec["physics"]["type"].get<std::string>();
        info.mass_density = body_spec["physics"]["mass"].get<double>();
        double v[3];
        detail::read_xyz(body_spec["physics"]["velocity"], v);
        info.velocity.vx = v[0];
        info.velocity.vy = v[1];
        info.velocity.vz = v[2];
        std::optional<common::geometry_t> g = detail::read_body_geometry(path, body_spec);
        if (!g.has_value())
            continue;
        info.id       = body_spec["id"].get<std::string>();
        info.geometry = std::move(*g);
        auto node     = physics_body_factory(info);
        node->set_as_physically_simulated_body();
        finish(node, info.id, body_spec);
    }
    return scene;
}

} // namespace io
} // namespace sbs

#endif // SBS_IO_LOAD_SCENE_H
