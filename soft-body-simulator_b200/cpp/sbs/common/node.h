// sbs/common/node.h — the reference's renderable_node_t (include/sbs/common/node.h:13-84,
// src/common/node.cpp): what io::load_scene hands back, one per body.  The GL object names are kept
// as plain integers for source compatibility; nothing here touches OpenGL.
#ifndef SBS_COMMON_NODE_H
#define SBS_COMMON_NODE_H

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace sbs {
namespace common {

class renderable_node_t
{
  public:
    virtual ~renderable_node_t() = default;
    void set_id(std::string const& id) { id_ = id; }
    std::string const& id() const { return id_; }
    void set_vao(unsigned int vao) { VAO_ = vao; }
    void set_vbo(unsigned int vbo) { VBO_ = vbo; }
    void set_ebo(unsigned int ebo) { EBO_ = ebo; }
    unsigned int const& VAO() const { return VAO_; }
    unsigned int const& VBO() const { return VBO_; }
    unsigned int const& EBO() const { return EBO_; }
    unsigned int& VAO() { return VAO_; }
    unsigned int& VBO() { return VBO_; }
    unsigned int& EBO() { return EBO_; }
    void mark_vertices_dirty() { transfer_vertices_ = true; }
    void mark_indices_dirty() { transfer_indices_ = true; }
    void mark_should_render_wireframe() { wireframe_ = true; }
    void mark_vertices_clean() { transfer_vertices_ = false; }
    void mark_indices_clean() { transfer_indices_ = false; }
    void mark_should_render_triangles() { wireframe_ = false; }
    bool should_transfer_vertices() const { return transfer_vertices_; }
    bool should_transfer_indices() const { return transfer_indices_; }
    bool should_render_triangles() const { return !wireframe_; }
    bool should_render_wireframe() const { return wireframe_; }
    bool is_environment_body() const { return body_type_ == body_type_t::environment; }
    bool is_physically_simulated_body() const { return body_type_ == body_type_t::physical; }
    void set_as_environment_body() { body_type_ = body_type_t::environment; }
    void set_as_physically_simulated_body() { body_type_ = body_type_t::physical; }
    void set_as_collideable_body() { is_collideable_ = true; }
    void set_as_non_collideable_body() { is_collideable_ = false; }
    bool is_collideable_body() const { return is_collideable_; }
    std::vector<float> const& get_cpu_vertex_buffer() const { return cpu_vertex_buffer_; }
    std::vector<std::uint32_t> const& get_cpu_index_buffer() const { return cpu_index_buffer_; }
    virtual void prepare_vertices_for_rendering() = 0;
    virtual void prepare_indices_for_rendering()  = 0;

  protected:
    void transfer_vertices_for_rendering(std::vector<float>&& vertices)
    {
        cpu_vertex_buffer_ = std::move(vertices);
        mark_vertices_dirty();
    }
    void transfer_indices_for_rendering(std::vector<std::uint32_t>&& indices)
    {
        cpu_index_buffer_ = std::move(indices);
        mark_indices_dirty();
    }

  private:
    std::string id_;
    bool transfer_vertices_ = true, transfer_indices_ = true, wireframe_ = false;
    bool is_collideable_ = false;
    enum class body_type_t { environment, physical } body_type_ = body_type_t::environment;
    std::vector<float> cpu_vertex_buffer_;        // (x, y, z, nx, ny, nz, r, g, b)
    std::vector<std::uint32_t> cpu_index_buffer_; // (v1, v2, v3)
    unsigned int VBO_ = 0, VAO_ = 0, EBO_ = 0;
};

} // namespace common
} // namespace sbs

#endif // SBS_COMMON_NODE_H
