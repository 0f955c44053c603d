/* forwards to the shim: see sf_ref_shim.h */
#include "../../sf_ref_shim.h"
