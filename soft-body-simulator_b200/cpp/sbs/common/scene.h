// sbs/common/scene.h — the reference's scene_t (include/sbs/common/scene.h:11-53): the nodes
// io::load_scene created and the two lights of the scene file.
#ifndef SBS_COMMON_SCENE_H
#define SBS_COMMON_SCENE_H

#include <memory>
#include <vector>

#include <sbs/common/node.h>

namespace sbs {
namespace common {

struct scene_t
{
    std::vector<std::shared_ptr<renderable_node_t>> nodes;
    struct ambient_t { float r = 0, g = 0, b = 0; };
    struct diffuse_t { float r = 0, g = 0, b = 0; };
    struct specular_t { float r = 0, g = 0, b = 0, exp = 0; };
    struct point_light_t
    {
        float x = 0, y = 0, z = 0;
        struct attenuation_t { float constant = 0, linear = 0, quadratic = 0; };
        ambient_t ambient;
        diffuse_t diffuse;
        specular_t specular;
        attenuation_t attenuation;
    };
    struct directional_light_t
    {
        float dx = 0, dy = 0, dz = 0;
        ambient_t ambient;
        diffuse_t diffuse;
        specular_t specular;
    };
    directional_light_t directional_light;
    point_light_t point_light;
};

} // namespace common
} // namespace sbs

#endif // SBS_COMMON_SCENE_H
