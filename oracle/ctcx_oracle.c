/* TEST INFRASTRUCTURE ONLY. This is the CPU oracle: a plain-C restatement of the reference's
 * algorithm used by tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() as the CHECKER.
 * The product path (ctc-beam-search-op_b200/) never links, loads or calls it.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this restatement against
 *   (1) the reference's own known-answer test
 *       (python/ops/ctc_ext_beam_search_decoder_ops_test.py:25-64, all seven outputs),
 *   (2) the edge-case table measured on the compiled reference,
 *   (3) oracle/_ref/libctcx_ref.so -- the reference's unmodified headers compiled by
 *       oracle/Makefile -- on seeded random inputs (labels, alignments and log-probs identical),
 *   (4) committed fixtures under tests/golden/ generated from (3) by tests/golden/make_golden.py.
 *
 * The algorithm itself is in ctcx_oracle_impl.h (instantiated for float and double below).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Decision-margin classes recorded per utterance (minimum absolute gap of every comparison of that
 * class). A margin of 0 is an exact tie, which the reference resolves through libstdc++ heap/sort
 * mechanics (unspecified order); BASELINE.json's north star excuses utterances whose competing
 * scores differ by <= 1e-5. */
enum {
  CTCX_MARGIN_ACCEPT = 0, /* child score vs beam bottom (decoder.h:152-154,189) */
  CTCX_MARGIN_BOTTOM = 1, /* bottom vs runner-up when evicting (decoder.h:192-198) */
  CTCX_MARGIN_ORDER = 2,  /* adjacent totals in the sorted `branches` (decoder.h:84) */
  CTCX_MARGIN_ALIGN = 3,  /* best vs second alignment candidate (entry.h:69-72,140) */
  CTCX_MARGIN_FINAL = 4,  /* adjacent totals among the returned paths (decoder.h:245-252) */
  CTCX_N_MARGINS = 5
};

/* Design statistics (summed over all frames of all utterances of one call). */
typedef struct {
  long long frames;
  long long turns_passed_gate;          /* members passing the gate (decoder.h:157) */
  long long child_evals;                /* inactive children scored (decoder.h:168-187) */
  long long accepted;                   /* children pushed into the beam (decoder.h:199) */
  long long wipes;                      /* evicted members re-scored and reset by their parent */
  long long wipes_before_turn;          /* ... whose own turn had not come yet */
  long long frames_with_effective_wipe; /* frames where such a member lost an acceptable child */
  long long relevant_children;          /* children above the W-th member total, per frame */
  long long surviving_children;         /* fresh children still in the beam at frame end */
  long long revisit_accepts;            /* evicted member re-accepted on revisit (rounding case) */
  long long max_relevant_children;
  long long effective_wipes;
  long long revisit_above_former;       /* re-scored evicted member scoring above its former total */
} ctcx_oracle_stats;

#define CTCX_CAT_(a, b) a##b
#define CTCX_CAT(a, b) CTCX_CAT_(a, b)

#define REAL float
#define SUF(name) CTCX_CAT(name, _f32)
#define REAL_EXP(v) expf(v) /* Eigen::numext::exp<float> -> std::exp(float) (decoder.h:76) */
#define REAL_LOG(v) logf(v)
#include "ctcx_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_EXP
#undef REAL_LOG

#define REAL double
#define SUF(name) CTCX_CAT(name, _f64)
#define REAL_EXP(v) exp(v)
#define REAL_LOG(v) log(v)
#include "ctcx_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_EXP
#undef REAL_LOG

/* logits: time-major [T,B,C]; outputs: dense rows with stride T per (b,p); margins: [B,5] or NULL;
 * stats: accumulates, or NULL. Returns 0 or the error code with its message in err. */
int ctcx_oracle_decode_f32(const float* logits, int T, int B, int C, const int* seq_len, int W,
                           int P, int merge_repeated, int blank_index, int blank_label,
                           int* dec_len, int* dec, int* ali_len, int* ali, float* logp,
                           double* margins, ctcx_oracle_stats* stats, char* err, int errcap) {
  return decode_f32(logits, T, B, C, seq_len, W, P, merge_repeated, blank_index, blank_label,
                    dec_len, dec, ali_len, ali, logp, margins, stats, err, errcap, NULL);
}

/* the same with a scorer: lm = [C+1, C] expansion-score table (see ctcx_oracle_impl.h), or NULL */
int ctcx_oracle_decode_lm_f32(const float* logits, int T, int B, int C, const int* seq_len, int W,
                              int P, int merge_repeated, int blank_index, int blank_label,
                              const float* lm, int* dec_len, int* dec, int* ali_len, int* ali,
                              float* logp, double* margins, ctcx_oracle_stats* stats, char* err,
                              int errcap) {
  return decode_f32(logits, T, B, C, seq_len, W, P, merge_repeated, blank_index, blank_label,
                    dec_len, dec, ali_len, ali, logp, margins, stats, err, errcap, lm);
}

int ctcx_oracle_decode_f64(const double* logits, int T, int B, int C, const int* seq_len, int W,
                           int P, int merge_repeated, int blank_index, int blank_label,
                           int* dec_len, int* dec, int* ali_len, int* ali, double* logp,
                           double* margins, ctcx_oracle_stats* stats, char* err, int errcap) {
  return decode_f64(logits, T, B, C, seq_len, W, P, merge_repeated, blank_index, blank_label,
                    dec_len, dec, ali_len, ali, logp, margins, stats, err, errcap, NULL);
}
