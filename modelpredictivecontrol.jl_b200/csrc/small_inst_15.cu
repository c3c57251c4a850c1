// step_small specialisations for NZT = 15 move variables (see bmpc_small_registry.h).
#include "bmpc_small_registry.h"

namespace bmpc {
void small_register_15(std::vector<SmallEntry>& v) {
    v.push_back(small_entry<15, 0, 3, 2>());
    v.push_back(small_entry<15, 1, 3, 2>());
    v.push_back(small_entry<15, 0, 6, 4>());
    v.push_back(small_entry<15, 1, 6, 4>());
}
}  // namespace bmpc
