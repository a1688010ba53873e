// Test driver for sbs/io/ply.h:
//   ply_tool bar W H D out.ply [ascii|binary_little_endian|binary_big_endian]   write get_simple_bar_model
//   ply_tool convert in.ply out.ply FORMAT                                       read, then write
//   ply_tool dump in.ply                                                         print counts and raw values
#include <cstdio>
#include <cstdlib>
#include <string>

#include <sbs/geometry/get_simple_bar_model.h>
#include <sbs/io/ply.h>

int main(int argc, char** argv)
{
    std::string const cmd = argc > 1 ? argv[1] : "";
    if (cmd == "bar" && argc >= 6)
    {
        auto const g = sbs::geometry::get_simple_bar_model(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]));
        sbs::io::write_ply(std::filesystem::path(argv[5]), g, sbs::io::string_to_format(argc > 6 ? argv[6] : "ascii"));
        return 0;
    }
    if (cmd == "convert" && argc >= 5)
    {
        auto const g = sbs::io::read_ply(std::filesystem::path(argv[2]));
        if (!g)
            return 4;
        sbs::io::write_ply(std::filesystem::path(argv[3]), *g, sbs::io::string_to_format(argv[4]));
        return 0;
    }
    if (cmd == "dump" && argc >= 3)
    {
        auto const g = sbs::io::read_ply(std::filesystem::path(argv[2]));
        if (!g)
        {
            std::printf("none\n");
            return 0;
        }
        std::printf("%s %zu %zu %zu %zu\n", g->is_tetrahedral_mesh() ? "tet" : "tri", g->positions.size() / 3,
                    g->indices.size(), g->colors.size(), g->normals.size());
        for (float p : g->positions)
            std::printf("%.9g ", p);
        std::printf("\n");
        for (int i : g->indices)
            std::printf("%d ", i);
        std::printf("\n");
        for (unsigned c : g->colors)
            std::printf("%u ", c);
        std::printf("\n");
        return 0;
    }
    return 2;
}
