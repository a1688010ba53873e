// Same include path as the reference's include/sbs/common/mesh.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
