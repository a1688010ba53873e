// Same include path as the reference's include/sbs/physics/xpbd/simulation_parameters.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
