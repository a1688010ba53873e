// Same include path as the reference's include/sbs/physics/xpbd/green_constraint.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
