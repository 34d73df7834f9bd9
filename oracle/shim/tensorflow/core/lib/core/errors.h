// TEST INFRASTRUCTURE ONLY (oracle/): errors::InvalidArgument (ctc_ext_beam_search_decoder.h:238,241).
#ifndef CTCX_ORACLE_SHIM_ERRORS_H_
#define CTCX_ORACLE_SHIM_ERRORS_H_
#include <string>
#include "tensorflow/core/lib/core/status.h"
namespace tensorflow {
namespace errors {
inline Status InvalidArgument(const std::string& msg) { return Status(msg); }
}  // namespace errors
}  // namespace tensorflow
#endif
