// TEST INFRASTRUCTURE ONLY (oracle/): the slice of tensorflow::Status the decoder header uses
// (ctc_ext_beam_search_decoder.h:229-260: Status::OK(), returned by value, ok(), message).
#ifndef CTCX_ORACLE_SHIM_STATUS_H_
#define CTCX_ORACLE_SHIM_STATUS_H_
#include <string>
namespace tensorflow {
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};
}  // namespace tensorflow
#endif
