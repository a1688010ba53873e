#!/bin/sh
# Compile the reference's OWN physics translation units, unmodified and where they lie under
# /root/reference, against oracle/ref_shim (Eigen and Discregrid are absent from the reference
# tree and from this image) plus oracle/ref_driver.cpp.  Output only into oracle/_ref/.
# Three reference sources use Windows-style '..\..\include\...' includes; they are satisfied by
# files literally so named, generated under oracle/_ref/bs/, that forward to the real headers.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SBS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/src/physics" ] || { echo "reference tree not found at $REF" >&2; exit 3; }
mkdir -p "$OUT/bs" "$OUT/obj"
printf '#include <sbs/physics/collision/sdf_model.h>\n' > "$OUT/bs/..\\..\\..\\include\\sbs\\physics\\collision\\sdf_model.h"
printf '#include <sbs/physics/environment_body.h>\n' > "$OUT/bs/..\\..\\include\\sbs\\physics\\environment_body.h"
printf '#include <sbs/physics/tetrahedral_mesh_boundary.h>\n' > "$OUT/bs/..\\..\\include\\sbs\\physics\\tetrahedral_mesh_boundary.h"
SRCS="
src/physics/body.cpp src/physics/constraint.cpp src/physics/environment_body.cpp
src/physics/gauss_seidel_solver.cpp src/physics/particle.cpp src/physics/simulation.cpp
src/physics/solver.cpp src/physics/tetrahedral_body.cpp src/physics/tetrahedral_mesh_boundary.cpp
src/physics/timestep.cpp src/physics/topology.cpp
src/physics/xpbd/collision_constraint.cpp src/physics/xpbd/contact_handler.cpp
src/physics/xpbd/distance_constraint.cpp src/physics/xpbd/green_constraint.cpp
src/physics/xpbd/simulation_parameters.cpp
src/physics/collision/brute_force_cd_system.cpp src/physics/collision/bvh_model.cpp
src/physics/collision/cd_system.cpp src/physics/collision/collision_model.cpp
src/physics/collision/contact.cpp src/physics/collision/sdf_model.cpp
src/common/mesh.cpp src/common/node.cpp src/common/geometry.cpp src/common/primitive.cpp
src/geometry/get_simple_bar_model.cpp
"
CXXFLAGS="-std=c++17 -O3 -fPIC -DNDEBUG -w -include cstdint -include cassert -include memory -include optional -include functional -include string"
INC="-I$REF/include -I$HERE/ref_shim -I$OUT/bs"
OBJS=""
for s in $SRCS; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_').o"
  g++ $CXXFLAGS $INC -c "$REF/$s" -o "$o" &
  OBJS="$OBJS $o"
done
g++ $CXXFLAGS $INC -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o" &
wait
g++ -shared -o "$OUT/libsbsref.so" $OBJS "$OUT/obj/ref_driver.o"
echo "$OUT/libsbsref.so"
# The reference's own scene loader (src/io/load_scene.cpp, ply.cpp) under tests/cpp/scene_dump.cpp, against
# the nlohmann/json.hpp that happens to be bundled with this image's python packages (the reference fetches it
# from the network, CMakeLists.txt:19-25).  Skipped when that header is not found.
NL="$(python3 -c 'import sysconfig,os;print(os.path.join(sysconfig.get_paths()["purelib"],"include","cudnn_frontend","thirdparty"))' 2>/dev/null || true)"
if [ -f "$NL/nlohmann/json.hpp" ]; then
  g++ -std=c++17 -O1 -w -include cstdint -include cassert -include optional -I"$REF/include" -I"$NL" \
    "$HERE/../tests/cpp/scene_dump.cpp" "$REF/src/io/load_scene.cpp" "$REF/src/io/ply.cpp" \
    "$REF/src/common/node.cpp" "$REF/src/common/geometry.cpp" -o "$OUT/ref_scene_dump"
  echo "$OUT/ref_scene_dump"
fi
