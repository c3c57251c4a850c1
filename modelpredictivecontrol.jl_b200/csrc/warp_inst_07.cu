// step_warp specialisations for NT = 7 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_07(std::vector<WarpEntry>& v) { warp_register_nt<7>(v); }
}  // namespace bmpc
