// Same include path as the reference's include/sbs/physics/collision/collision_model.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
