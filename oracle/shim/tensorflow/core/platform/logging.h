// TEST INFRASTRUCTURE ONLY (oracle/): CHECK macros (ctc_ext_beam_search_decoder.h:32,83,234-236).
#ifndef CTCX_ORACLE_SHIM_LOGGING_H_
#define CTCX_ORACLE_SHIM_LOGGING_H_
#include <cstdio>
#include <cstdlib>
namespace tensorflow {
namespace shim_internal {
inline void CheckFail(const char* what, const char* file, int line) {
  std::fprintf(stderr, "CHECK failed: %s at %s:%d\n", what, file, line);
  std::abort();
}
template <typename T>
T&& CheckNotNull(const char* what, const char* file, int line, T&& t) {
  if (t == nullptr) CheckFail(what, file, line);
  return static_cast<T&&>(t);
}
}  // namespace shim_internal
}  // namespace tensorflow
#define CHECK(c) \
  do { if (!(c)) ::tensorflow::shim_internal::CheckFail(#c, __FILE__, __LINE__); } while (0)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NOTNULL(p) \
  ::tensorflow::shim_internal::CheckNotNull(#p " != nullptr", __FILE__, __LINE__, (p))
#endif
