// step_small specialisations for NZT = 12 move variables (see bmpc_small_registry.h).
#include "bmpc_small_registry.h"

namespace bmpc {
void small_register_12(std::vector<SmallEntry>& v) {
    v.push_back(small_entry<12, 0, 3, 2>());
    v.push_back(small_entry<12, 1, 3, 2>());
}
}  // namespace bmpc
