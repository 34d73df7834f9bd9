// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path: only tests/, bench.py's CPU
// baseline leg and __graft_entry__.smoke() may load the library built from this file.
//
// C driver around the reference's UNMODIFIED header-only decoder. The headers are compiled where they
// lie under /root/reference (nothing is copied into this repo); the TensorFlow/Eigen pieces they
// include are replaced by the small stand-ins in oracle/shim/. Output goes to oracle/_ref/ only.
//
// This file restates the OpKernel glue, which cannot be compiled without TensorFlow:
//   * Compute's batch/time loops         tensorflow_ctc_ext_beam_search_decoder/cc/kernels/
//                                        ctc_ext_beam_search_decoder_kernels.cc:55-90
//   * ValidateInputsGenerateOutputs      same file :97-160 (the checks that depend on values)
// Sparse packing (StoreAllDecodedSequences, :163-257) is restated in tests/ctcx_testlib.py because
// it is pure index bookkeeping over the dense rows returned here.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "tensorflow_ctc_ext_beam_search_decoder/cc/util/ctc_ext_beam_search_decoder.h"

namespace {

// Stand-in for the Eigen::Map<const Eigen::Array<T,Dynamic,1>> that kernels.cc:76-78 hands to Step():
// Step only calls maxCoeff(), size() and operator()(int) (ctc_ext_beam_search_decoder.h:67-83).
template <typename T>
struct RowView {
  const T* p;
  int n;
  T maxCoeff() const {
    T m = p[0];
    for (int i = 1; i < n; ++i) m = (p[i] > m) ? p[i] : m;
    return m;
  }
  int size() const { return n; }
  T operator()(int i) const { return p[i]; }
};

void SetErr(char* err, int errcap, const std::string& msg) {
  if (err != nullptr && errcap > 0) {
    std::snprintf(err, static_cast<size_t>(errcap), "%s", msg.c_str());
  }
}

// A scorer for the reference's extension point (util/ctc_beam_scorer.h:14-20, "a thin layer for
// integrating language model scoring"). Only STATELESS scorers compile against the reference as it
// is: BeamEntry::AddAlignmentCandidate takes a BeamEntry<T>* (util/ctc_beam_entry.h:190), i.e. the
// default EmptyBeamState, so a decoder instantiated with any other state type is ill-formed. A
// stateless scorer sees no labels; what it can express is a constant expansion score per new label
// (a label insertion penalty): GetStateExpansionScore(state, previous) = previous + penalty.
template <typename T>
class ConstScorer : public tensorflow::ctc::BaseBeamScorer<T, tensorflow::ctc::ctc_beam_search::EmptyBeamState> {
 public:
  explicit ConstScorer(T penalty) : penalty_(penalty) {}
  T GetStateExpansionScore(const tensorflow::ctc::ctc_beam_search::EmptyBeamState& state,
                           T previous_score) const override {
    (void)state;
    return previous_score + penalty_;
  }

 private:
  T penalty_;
};

// Decodes utterances [b_begin, b_end) of a time-major [T,B,C] tensor. Outputs are dense rows with
// stride T per (b, p): dec[(b*P+p)*T + i], ali[(b*P+p)*T + i]; logp[b*P+p].
template <typename T, typename Decoder>
int RunBatch(Decoder& decoder, const T* logits, int max_time, int batch, int num_classes, const int* seq_len,
             int b_begin, int b_end, int top_paths, bool merge_repeated, int* dec_len, int* dec,
             int* ali_len, int* ali, T* logp, char* err, int errcap);

template <typename T>
int Decode(const T* logits, int max_time, int batch, int num_classes, const int* seq_len,
           int b_begin, int b_end, int beam_width, int top_paths, bool merge_repeated,
           int blank_index, int blank_label, int* dec_len, int* dec, int* ali_len, int* ali,
           T* logp, char* err, int errcap, const T* penalty = nullptr) {
  using tensorflow::ctc::CTCExtBeamSearchDecoder;
  // kernels.cc:118-120
  if (max_time == 0) {
    SetErr(err, errcap, "max_time is 0");
    return 2;
  }
  // kernels.cc:134-138
  for (int b = 0; b < batch; ++b) {
    if (!(seq_len[b] <= max_time)) {
      SetErr(err, errcap, "sequence_length(" + std::to_string(b) + ") <= " + std::to_string(max_time));
      return 5;
    }
  }
  if (penalty != nullptr) {  // the same loops with a scorer plugged into the extension point
    ConstScorer<T> scorer(*penalty);
    CTCExtBeamSearchDecoder<T> decoder(num_classes, blank_index, beam_width, &scorer, blank_label, 1,
                                       merge_repeated);
    return RunBatch<T>(decoder, logits, max_time, batch, num_classes, seq_len, b_begin, b_end, top_paths,
                       merge_repeated, dec_len, dec, ali_len, ali, logp, err, errcap);
  }
  typename CTCExtBeamSearchDecoder<T>::DefaultBeamScorer scorer;
  // kernels.cc:55-57: ONE decoder with batch_size=1, re-used across the batch after Reset().
  CTCExtBeamSearchDecoder<T> decoder(num_classes, blank_index, beam_width, &scorer, blank_label, 1,
                                     merge_repeated);
  return RunBatch<T>(decoder, logits, max_time, batch, num_classes, seq_len, b_begin, b_end, top_paths,
                     merge_repeated, dec_len, dec, ali_len, ali, logp, err, errcap);
}

template <typename T, typename Decoder>
int RunBatch(Decoder& decoder, const T* logits, int max_time, int batch, int num_classes, const int* seq_len,
             int b_begin, int b_end, int top_paths, bool merge_repeated, int* dec_len, int* dec,
             int* ali_len, int* ali, T* logp, char* err, int errcap) {
  std::vector<std::vector<int> > paths, alignments;
  std::vector<T> log_probs;
  for (int b = b_begin; b < b_end; ++b) {
    // kernels.cc:74-79
    for (int t = 0; t < seq_len[b]; ++t) {
      RowView<T> row = {logits + (static_cast<size_t>(t) * batch + b) * num_classes, num_classes};
      decoder.Step(row);
    }
    // kernels.cc:81-83
    tensorflow::Status s = decoder.TopPaths(top_paths, &paths, &alignments, &log_probs, merge_repeated);
    if (!s.ok()) {
      SetErr(err, errcap, s.error_message());
      return s.error_message().find("more paths") != std::string::npos ? 6 : 7;
    }
    decoder.Reset();  // kernels.cc:85
    for (int p = 0; p < top_paths; ++p) {
      const size_t row = static_cast<size_t>(b) * top_paths + p;
      logp[row] = log_probs[p];  // kernels.cc:87-89
      dec_len[row] = static_cast<int>(paths[p].size());
      ali_len[row] = static_cast<int>(alignments[p].size());
      for (size_t i = 0; i < paths[p].size(); ++i) dec[row * max_time + i] = paths[p][i];
      for (size_t i = 0; i < alignments[p].size(); ++i) ali[row * max_time + i] = alignments[p][i];
    }
  }
  return 0;
}

// After every frame, the full beam (best first): used by tests to pin per-frame state
// (SURVEY.md Appendix C trace). n_out[t] = number of leaves after frame t.
template <typename T>
int Trace(const T* logits, int max_time, int num_classes, int beam_width, bool merge_repeated,
          int blank_index, int blank_label, int* n_out, T* logp, int* dec_len, int* dec, int* ali) {
  using tensorflow::ctc::CTCExtBeamSearchDecoder;
  typename CTCExtBeamSearchDecoder<T>::DefaultBeamScorer scorer;
  CTCExtBeamSearchDecoder<T> decoder(num_classes, blank_index, beam_width, &scorer, blank_label, 1,
                                     merge_repeated);
  std::vector<std::vector<int> > paths, alignments;
  std::vector<T> log_probs;
  for (int t = 0; t < max_time; ++t) {
    RowView<T> row = {logits + static_cast<size_t>(t) * num_classes, num_classes};
    decoder.Step(row);
    int n = beam_width;
    while (n > 0 && !decoder.TopPaths(n, &paths, &alignments, &log_probs, merge_repeated).ok()) --n;
    n_out[t] = n;
    for (int p = 0; p < n; ++p) {
      const size_t r = static_cast<size_t>(t) * beam_width + p;
      logp[r] = log_probs[p];
      dec_len[r] = static_cast<int>(paths[p].size());
      for (size_t i = 0; i < paths[p].size(); ++i) dec[r * max_time + i] = paths[p][i];
      for (size_t i = 0; i < alignments[p].size(); ++i) ali[r * max_time + i] = alignments[p][i];
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int ctcx_ref_decode_f32(const float* logits, int T, int B, int C, const int* seq_len, int b_begin,
                        int b_end, int W, int P, int merge_repeated, int blank_index,
                        int blank_label, int* dec_len, int* dec, int* ali_len, int* ali, float* logp,
                        char* err, int errcap) {
  return Decode<float>(logits, T, B, C, seq_len, b_begin, b_end, W, P, merge_repeated != 0,
                       blank_index, blank_label, dec_len, dec, ali_len, ali, logp, err, errcap);
}

int ctcx_ref_decode_f64(const double* logits, int T, int B, int C, const int* seq_len, int b_begin,
                        int b_end, int W, int P, int merge_repeated, int blank_index,
                        int blank_label, int* dec_len, int* dec, int* ali_len, int* ali,
                        double* logp, char* err, int errcap) {
  return Decode<double>(logits, T, B, C, seq_len, b_begin, b_end, W, P, merge_repeated != 0,
                        blank_index, blank_label, dec_len, dec, ali_len, ali, logp, err, errcap);
}

int ctcx_ref_trace_f32(const float* logits, int T, int C, int W, int merge_repeated,
                       int blank_index, int blank_label, int* n_out, float* logp, int* dec_len,
                       int* dec, int* ali) {
  return Trace<float>(logits, T, C, W, merge_repeated != 0, blank_index, blank_label, n_out, logp,
                      dec_len, dec, ali);
}


/* the reference decoder with a stateless scorer (constant expansion score) plugged into its
 * BaseBeamScorer extension point */
int ctcx_ref_decode_penalty_f32(const float* logits, int T, int B, int C, const int* seq_len, int b_begin,
                                int b_end, int W, int P, int merge_repeated, int blank_index,
                                int blank_label, float penalty, int* dec_len, int* dec, int* ali_len,
                                int* ali, float* logp, char* err, int errcap) {
  return Decode<float>(logits, T, B, C, seq_len, b_begin, b_end, W, P, merge_repeated != 0,
                       blank_index, blank_label, dec_len, dec, ali_len, ali, logp, err, errcap, &penalty);
}

}  // extern "C"
