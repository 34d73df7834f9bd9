// TEST INFRASTRUCTURE ONLY (oracle/): gtl::FlatMap is used by the reference only through
// emplace(key, nullptr) -> pair<iterator,bool> (ctc_beam_entry.h:115); std::unordered_map has
// identical semantics for that call (no iteration order is ever observed).
#ifndef CTCX_ORACLE_SHIM_FLATMAP_H_
#define CTCX_ORACLE_SHIM_FLATMAP_H_
#include <unordered_map>
namespace tensorflow {
namespace gtl {
template <typename K, typename V>
using FlatMap = std::unordered_map<K, V>;
}  // namespace gtl
}  // namespace tensorflow
#endif
