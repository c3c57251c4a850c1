// step_warp specialisations for NT = 16 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_16(std::vector<WarpEntry>& v) { warp_register_nt<16>(v); }
}  // namespace bmpc
