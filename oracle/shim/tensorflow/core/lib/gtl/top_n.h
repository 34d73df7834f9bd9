// TEST INFRASTRUCTURE ONLY (oracle/).
//
// Restatement of tensorflow::gtl::TopN<T, Cmp> -- a THIRD-PARTY dependency of the reference that is
// NOT vendored under /root/reference (it comes from whatever `pip install tensorflow` resolved to,
// configure.sh:91-125; version unpinned, TF 1.14-2.x era). The reference's call sites are
// ctc_ext_beam_search_decoder.h:59 (leaves_), :84-85 (Extract/Reset), :142 (push), :153-154
// (size/peek_bottom), :195-199 (peek_bottom/push), :245-252 (TopPaths).
//
// Published algorithm (tensorflow/core/lib/gtl/top_n.h), restated:
//   * elements_ is a vector with three states. UNORDERED: plain vector. BOTTOM_KNOWN: plain vector
//     whose element 0 is the minimum (w.r.t. Cmp, "a ranks above b"). HEAP_SORTED: limit+1 slots, the
//     first `limit` form a heap whose front is the minimum, the last slot is scratch.
//   * push: in the vector states append (keeping the minimum in front when BOTTOM_KNOWN); on reaching
//     limit+1 elements heapify and pop the minimum into the scratch slot. In HEAP_SORTED a value is
//     admitted only if it ranks STRICTLY above the current minimum.
//   * peek_bottom: the minimum (linear scan for the first minimum when UNORDERED).
//   * Extract: all kept elements sorted best-first; leaves the container empty and UNORDERED.
// Set semantics are exact; only the order among exactly-equal keys depends on libstdc++'s
// sort/heap mechanics.
#ifndef CTCX_ORACLE_SHIM_TOP_N_H_
#define CTCX_ORACLE_SHIM_TOP_N_H_
#include <algorithm>
#include <cstddef>
#include <functional>
#include <utility>
#include <vector>
#include "tensorflow/core/platform/logging.h"

namespace tensorflow {
namespace gtl {

template <class T, class Cmp = std::greater<T> >
class TopN {
 public:
  enum State { UNORDERED, BOTTOM_KNOWN, HEAP_SORTED };

  explicit TopN(size_t limit) : limit_(limit), cmp_(), state_(UNORDERED) {}
  TopN(size_t limit, const Cmp& cmp) : limit_(limit), cmp_(cmp), state_(UNORDERED) {}

  size_t limit() const { return limit_; }
  size_t size() const { return std::min(elements_.size(), limit_); }
  bool empty() const { return size() == 0; }

  void push(const T& v) {
    if (limit_ == 0) return;
    if (state_ != HEAP_SORTED) {
      elements_.push_back(v);
      if (!(state_ == UNORDERED || cmp_(elements_.back(), elements_.front()))) {
        // BOTTOM_KNOWN and the newcomer does not rank above the known bottom: it is the new bottom.
        std::swap(elements_.front(), elements_.back());
      }
      if (elements_.size() == limit_ + 1) {
        std::make_heap(elements_.begin(), elements_.end(), cmp_);
        std::pop_heap(elements_.begin(), elements_.end(), cmp_);
        state_ = HEAP_SORTED;
      }
    } else {
      if (cmp_(v, elements_.front())) {
        elements_.back() = v;  // scratch slot
        std::pop_heap(elements_.begin(), elements_.end(), cmp_);
      }
    }
  }

  const T& peek_bottom() {
    CHECK(!empty());
    if (state_ == UNORDERED) {
      size_t min_candidate = 0;
      for (size_t i = 1; i < elements_.size(); ++i) {
        if (cmp_(elements_[min_candidate], elements_[i])) min_candidate = i;
      }
      if (min_candidate != 0) std::swap(elements_[0], elements_[min_candidate]);
      state_ = BOTTOM_KNOWN;
    }
    return elements_.front();
  }

  // Caller owns the returned vector (the reference wraps it in a unique_ptr).
  std::vector<T>* Extract() {
    std::vector<T>* out = new std::vector<T>;
    out->swap(elements_);
    if (state_ != HEAP_SORTED) {
      std::sort(out->begin(), out->end(), cmp_);
    } else {
      out->pop_back();  // scratch slot
      std::sort_heap(out->begin(), out->end(), cmp_);
    }
    state_ = UNORDERED;
    return out;
  }

  typename std::vector<T>::const_iterator unsorted_begin() const { return elements_.begin(); }
  typename std::vector<T>::const_iterator unsorted_end() const {
    return elements_.begin() + size();
  }

  void Reset() {
    elements_.clear();
    state_ = UNORDERED;
  }

 private:
  std::vector<T> elements_;
  size_t limit_;
  Cmp cmp_;
  State state_;
};

}  // namespace gtl
}  // namespace tensorflow
#endif
