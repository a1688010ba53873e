// Same include path as the reference's include/sbs/geometry/get_simple_bar_model.h; the declarations live in sbs/b200/facade.hpp.
#include <sbs/b200/facade.hpp>
