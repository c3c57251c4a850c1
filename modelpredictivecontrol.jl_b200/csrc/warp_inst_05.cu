// step_warp specialisations for NT = 5 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_05(std::vector<WarpEntry>& v) { warp_register_nt<5>(v); }
}  // namespace bmpc
