// sbs/io/ply.h — PLY reader/writer with the reference's interface (include/sbs/io/ply.h:60-98) and
// file conventions (src/io/ply.cpp): elements `vertex` (x y z [nx ny nz] [r g b] [u v]), `face`
// (triangles) and the non-standard `tet`; the index list property is called `indices`
// (`vertex_indices` is accepted too when reading); ascii, binary little- and big-endian; paths must
// end in ".ply"; header lines that are not understood (comment, obj_info) are skipped.  Own
// implementation; host-side only (this is the front door of the GPU solver, not part of the hot path).
//
// Known-answer tests: write_ply(get_simple_bar_model(2,2,2)) reproduces the reference's
// data/meshes/cube_tet.ply byte for byte, (5,2,2) its tet_bar_5x2x2.ply (tests/test_ply_io.py).
// One deliberate difference: the reference's BINARY writer emits only three indices per element even
// for tets (ply.cpp, the `j < 3` loop); this writer emits all of them.
#ifndef SBS_IO_PLY_H
#define SBS_IO_PLY_H

#include <array>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <istream>
#include <optional>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include <sbs/common/geometry.h>

namespace sbs {
namespace io {

enum class ply_format_t { ascii, binary_little_endian, binary_big_endian };

struct ply_element_property_t
{
    std::string name;
    bool is_list = false;
    std::array<std::string, 2> type; // scalar: type[0]; list: type[0] = count type, type[1] = item type
};
struct ply_element_t
{
    std::string name;
    std::size_t count = 0;
    std::vector<ply_element_property_t> properties;
};
struct ply_header_description_t
{
    ply_format_t format = ply_format_t::ascii;
    std::vector<ply_element_t> elements;
};

inline ply_format_t string_to_format(std::string const& s)
{
    if (s == "binary_little_endian")
        return ply_format_t::binary_little_endian;
    if (s == "binary_big_endian")
        return ply_format_t::binary_big_endian;
    return ply_format_t::ascii;
}

namespace detail {

inline bool machine_is_little_endian()
{
    std::uint16_t const probe = 1;
    unsigned char first;
    std::memcpy(&first, &probe, 1);
    return first == 1;
}

inline int type_size(std::string const& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8")
        return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16")
        return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32")
        return 4;
    if (t == "double" || t == "float64")
        return 8;
    return 0;
}

// one scalar of a binary stream as a double (exact for every PLY type but 64-bit integers, which PLY lacks)
inline bool read_binary_scalar(std::istream& is, std::string const& t, bool swap, double& out)
{
    int const n = type_size(t);
    if (n == 0)
        return false;
    unsigned char b[8];
    is.read(reinterpret_cast<char*>(b), n);
    if (!is)
        return false;
    if (swap)
        for (int i = 0; i < n / 2; ++i)
            std::swap(b[i], b[n - 1 - i]);
    if (t == "char" || t == "int8") { std::int8_t v; std::memcpy(&v, b, 1); out = v; }
    else if (t == "uchar" || t == "uint8") { std::uint8_t v; std::memcpy(&v, b, 1); out = v; }
    else if (t == "short" || t == "int16") { std::int16_t v; std::memcpy(&v, b, 2); out = v; }
    else if (t == "ushort" || t == "uint16") { std::uint16_t v; std::memcpy(&v, b, 2); out = v; }
    else if (t == "int" || t == "int32") { std::int32_t v; std::memcpy(&v, b, 4); out = v; }
    else if (t == "uint" || t == "uint32") { std::uint32_t v; std::memcpy(&v, b, 4); out = v; }
    else if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); out = v; }
    else { double v; std::memcpy(&v, b, 8); out = v; }
    return true;
}

template <typename T>
inline void write_binary(std::ostream& os, T v, bool swap)
{
    unsigned char b[sizeof(T)];
    std::memcpy(b, &v, sizeof(T));
    if (swap)
        for (std::size_t i = 0; i < sizeof(T) / 2; ++i)
            std::swap(b[i], b[sizeof(T) - 1 - i]);
    os.write(reinterpret_cast<char const*>(b), static_cast<std::streamsize>(sizeof(T)));
}

inline bool parse_header(std::istream& is, ply_header_description_t& d)
{
    std::string line;
    if (!std::getline(is, line))
        return false;
    if (!line.empty() && line.back() == '\r')
        line.pop_back();
    if (line != "ply")
        return false;
    while (std::getline(is, line))
    {
        if (!line.empty() && line.back() == '\r')
            line.pop_back();
        std::istringstream ls(line);
        std::string word;
        ls >> word;
        if (word == "end_header")
            return true;
        if (word == "format")
        {
            std::string f;
            ls >> f;
            d.format = string_to_format(f);
        }
        else if (word == "element")
        {
            ply_element_t e;
            ls >> e.name >> e.count;
            d.elements.push_back(e);
        }
        else if (word == "property" && !d.elements.empty())
        {
            ply_element_property_t p;
            std::string t;
            ls >> t;
            if (t == "list")
            {
                p.is_list = true;
                ls >> p.type[0] >> p.type[1] >> p.name;
            }
            else
            {
                p.type[0] = t;
                ls >> p.name;
            }
            d.elements.back().properties.push_back(p);
        }
        // anything else (comment, obj_info, ...) is skipped
    }
    return false; // no end_header
}

// the values of one element row, scalars and list items alike, read with `next`
template <typename Next>
inline bool consume_row(ply_element_t const& e, Next&& next, common::geometry_t& g, std::size_t list_arity)
{
    float xyz[3] = {0, 0, 0}, nrm[3] = {0, 0, 0}, uv[2] = {0, 0};
    std::uint8_t rgb[3] = {0, 0, 0};
    bool has_n = false, has_c = false, has_uv = false;
    for (ply_element_property_t const& p : e.properties)
    {
        if (p.is_list)
        {
            double count;
            if (!next(p.type[0], count))
                return false;
            bool const is_indices = p.name == "indices" || p.name == "vertex_indices" || p.name == "vertex_index";
            if (is_indices && static_cast<std::size_t>(count) != list_arity)
                return false; // triangles and tets only
            for (int i = 0; i < static_cast<int>(count); ++i)
            {
                double v;
                if (!next(p.type[1], v))
                    return false;
                if (is_indices)
                    g.indices.push_back(static_cast<int>(v));
            }
            continue;
        }
        double v;
        if (!next(p.type[0], v))
            return false;
        if (e.name != "vertex")
            continue;
        std::string const& n = p.name;
        if (n == "x") xyz[0] = static_cast<float>(v);
        else if (n == "y") xyz[1] = static_cast<float>(v);
        else if (n == "z") xyz[2] = static_cast<float>(v);
        else if (n == "nx") { nrm[0] = static_cast<float>(v); has_n = true; }
        else if (n == "ny") { nrm[1] = static_cast<float>(v); has_n = true; }
        else if (n == "nz") { nrm[2] = static_cast<float>(v); has_n = true; }
        else if (n == "r" || n == "red") { rgb[0] = static_cast<std::uint8_t>(v); has_c = true; }
        else if (n == "g" || n == "green") { rgb[1] = static_cast<std::uint8_t>(v); has_c = true; }
        else if (n == "b" || n == "blue") { rgb[2] = static_cast<std::uint8_t>(v); has_c = true; }
        else if (n == "u" || n == "s") { uv[0] = static_cast<float>(v); has_uv = true; }
        else if (n == "v" || n == "t") { uv[1] = static_cast<float>(v); has_uv = true; }
    }
    if (e.name == "vertex")
    {
        g.positions.insert(g.positions.end(), xyz, xyz + 3);
        if (has_n)
            g.normals.insert(g.normals.end(), nrm, nrm + 3);
        if (has_c)
            g.colors.insert(g.colors.end(), rgb, rgb + 3);
        if (has_uv)
            g.uvs.insert(g.uvs.end(), uv, uv + 2);
    }
    return true;
}

inline std::optional<common::geometry_t> read_body(std::istream& is, ply_header_description_t const& d)
{
    common::geometry_t g;
    g.geometry_type = common::geometry_t::geometry_type_t::triangle;
    for (ply_element_t const& e : d.elements)
        if (e.name == "tet")
            g.geometry_type = common::geometry_t::geometry_type_t::tetrahedron;
    bool const binary = d.format != ply_format_t::ascii;
    bool const swap   = binary && (machine_is_little_endian() != (d.format == ply_format_t::binary_little_endian));
    for (ply_element_t const& e : d.elements)
    {
        std::size_t const arity = e.name == "tet" ? 4u : 3u;
        for (std::size_t row = 0; row < e.count; ++row)
        {
            bool ok;
            if (binary)
                ok = consume_row(e, [&](std::string const& t, double& v) { return read_binary_scalar(is, t, swap, v); },
                                 g, arity);
            else
                ok = consume_row(e, [&](std::string const& t, double& v) {
                    if (type_size(t) == 0)
                        return false;
                    is >> v;
                    return static_cast<bool>(is);
                }, g, arity);
            if (!ok)
                return std::nullopt;
        }
    }
    return g;
}

} // namespace detail

// include/sbs/io/ply.h:87-98
inline std::optional<common::geometry_t> read_ply_ascii(std::istream& is, ply_header_description_t const& description)
{
    return detail::read_body(is, description);
}
inline std::optional<common::geometry_t> read_ply_binary(std::istream& is, ply_header_description_t const& description)
{
    return detail::read_body(is, description);
}

// include/sbs/io/ply.h:80-86
inline std::optional<common::geometry_t> read_ply(std::istream& is)
{
    ply_header_description_t d;
    if (!detail::parse_header(is, d))
        return std::nullopt;
    return detail::read_body(is, d);
}
inline std::optional<common::geometry_t> read_ply(std::filesystem::path const& path)
{
    if (!path.has_extension() || path.extension() != ".ply")
        return std::nullopt; // src/io/ply.cpp:303-320
    std::ifstream ifs(path.c_str(), std::ios::binary);
    if (!ifs.is_open())
        return std::nullopt;
    return read_ply(ifs);
}

// include/sbs/io/ply.h:60-79, src/io/ply.cpp:25-301
inline void write_ply(std::ostream& os, common::geometry_t const& geometry, ply_format_t format = ply_format_t::ascii)
{
    bool const is_tet_mesh       = geometry.is_tetrahedral_mesh();
    std::uint32_t const vertices = static_cast<std::uint32_t>(geometry.positions.size() / 3u);
    bool const has_normals       = geometry.normals.size() == geometry.positions.size();
    bool const has_colors        = geometry.colors.size() == geometry.positions.size();
    bool const has_uvs           = geometry.uvs.size() == static_cast<std::size_t>(vertices) * 2u;
    std::uint8_t const arity     = is_tet_mesh ? 4u : 3u;
    std::uint32_t const elements = static_cast<std::uint32_t>(geometry.indices.size() / arity);

    os << "ply\n"
       << (format == ply_format_t::ascii                  ? "format ascii 1.0\n"
           : format == ply_format_t::binary_little_endian ? "format binary_little_endian 1.0\n"
                                                          : "format binary_big_endian 1.0\n")
       << "element vertex " << vertices << "\n"
       << "property float x\nproperty float y\nproperty float z\n";
    if (has_normals)
        os << "property float nx\nproperty float ny\nproperty float nz\n";
    if (has_colors)
        os << "property uchar r\nproperty uchar g\nproperty uchar b\n";
    if (has_uvs)
        os << "property float u\nproperty float v\n";
    os << "element " << (is_tet_mesh ? "tet " : "face ") << elements << "\n"
       << "property list uchar int indices\nend_header\n";

    if (format == ply_format_t::ascii)
    {
        for (std::uint32_t i = 0; i < vertices; ++i)
        {
            // std::to_string(float): "%f", the formatting of the reference's fixtures
            os << std::to_string(geometry.positions[3 * i]) << " " << std::to_string(geometry.positions[3 * i + 1]) << " "
               << std::to_string(geometry.positions[3 * i + 2]);
            if (has_normals)
                os << " " << std::to_string(geometry.normals[3 * i]) << " " << std::to_string(geometry.normals[3 * i + 1])
                   << " " << std::to_string(geometry.normals[3 * i + 2]);
            if (has_colors)
                os << " " << std::to_string(geometry.colors[3 * i]) << " " << std::to_string(geometry.colors[3 * i + 1])
                   << " " << std::to_string(geometry.colors[3 * i + 2]);
            if (has_uvs)
                os << " " << std::to_string(geometry.uvs[2 * i]) << " " << std::to_string(geometry.uvs[2 * i + 1]);
            os << "\n";
        }
        for (std::uint32_t e = 0; e < elements; ++e)
        {
            os << std::to_string(arity);
            for (std::uint8_t j = 0; j < arity; ++j)
                os << " " << std::to_string(geometry.indices[static_cast<std::size_t>(e) * arity + j]);
            os << "\n";
        }
        return;
    }
    bool const swap = detail::machine_is_little_endian() != (format == ply_format_t::binary_little_endian);
    for (std::uint32_t i = 0; i < vertices; ++i)
    {
        for (int k = 0; k < 3; ++k)
            detail::write_binary(os, geometry.positions[3 * i + k], swap);
        if (has_normals)
            for (int k = 0; k < 3; ++k)
                detail::write_binary(os, geometry.normals[3 * i + k], swap);
        if (has_colors)
            for (int k = 0; k < 3; ++k)
                detail::write_binary(os, geometry.colors[3 * i + k], swap);
        if (has_uvs)
            for (int k = 0; k < 2; ++k)
                detail::write_binary(os, geometry.uvs[2 * i + k], swap);
    }
    for (std::uint32_t e = 0; e < elements; ++e)
    {
        detail::write_binary(os, arity, swap);
        for (std::uint8_t j = 0; j < arity; ++j)
            detail::write_binary(os, static_cast<std::int32_t>(geometry.indices[static_cast<std::size_t>(e) * arity + j]), swap);
    }
}
inline void write_ply(std::filesystem::path const& filepath, common::geometry_t const& geometry,
                      ply_format_t format = ply_format_t::ascii)
{
    if (!filepath.has_extension() || filepath.extension() != ".ply")
        return;
    std::ofstream ofs(filepath.c_str(), std::ios::binary);
    if (!ofs.is_open())
        return;
    write_ply(ofs, geometry, format);
}

} // namespace io
} // namespace sbs

#endif // SBS_IO_PLY_H
