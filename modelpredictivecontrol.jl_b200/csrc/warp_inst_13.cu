// step_warp specialisations for NT = 13 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_13(std::vector<WarpEntry>& v) { warp_register_nt<13>(v); }
}  // namespace bmpc
