// TEST INFRASTRUCTURE ONLY (oracle/): integer typedefs the reference headers expect from TF.
#ifndef CTCX_ORACLE_SHIM_TYPES_H_
#define CTCX_ORACLE_SHIM_TYPES_H_
#include <cstdint>
namespace tensorflow {
typedef std::int32_t int32;
typedef std::int64_t int64;
}  // namespace tensorflow
#endif
