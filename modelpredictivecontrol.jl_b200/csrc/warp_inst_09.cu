// step_warp specialisations for NT = 9 variables (see bmpc_warp_registry.h).
#include "bmpc_warp_registry.h"

namespace bmpc {
void warp_register_09(std::vector<WarpEntry>& v) { warp_register_nt<9>(v); }
}  // namespace bmpc
