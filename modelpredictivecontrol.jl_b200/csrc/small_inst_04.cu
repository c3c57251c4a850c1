// step_small specialisations for NZT = 4 move variables (see bmpc_small_registry.h).
#include "bmpc_small_registry.h"

namespace bmpc {
void small_register_04(std::vector<SmallEntry>& v) {
    v.push_back(small_entry<4, 0, 3, 2>());
    v.push_back(small_entry<4, 1, 3, 2>());
}
}  // namespace bmpc
