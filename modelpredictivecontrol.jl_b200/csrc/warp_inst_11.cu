// step_warp specialisations for NT = 11 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_11(std::vector<WarpEntry>& v) { warp_register_nt<11>(v); }
}  // namespace bmpc
