// Same include path as the reference's include/sbs/physics/xpbd/distance_constraint.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
