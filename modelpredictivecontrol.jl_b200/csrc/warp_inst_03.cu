// step_warp specialisations for NT = 3 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_03(std::vector<WarpEntry>& v) { warp_register_nt<3>(v); }
}  // namespace bmpc
