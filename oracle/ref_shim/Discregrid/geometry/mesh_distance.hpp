#include "../../shim_discregrid.h"
