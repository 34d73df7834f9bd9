/* TEST INFRASTRUCTURE ONLY (oracle/). Sweeps libm_port.h against the host libm.
 *   ctcx_port_mismatches(stride): every stride-th float of the three domains; returns the number of
 *   inputs whose results differ in any bit (out[0..2] = per-function counts). stride=1 is the
 *   exhaustive sweep quoted in libm_port.h (about 2 CPU-minutes). */
#include "libm_port.h"

long long ctcx_port_mismatches(unsigned stride, long long* out) {
  if (stride == 0) stride = 1;
  long long be = 0, bl = 0, bg = 0;
  for (uint64_t u = 0x80000000ull; u <= (uint64_t)ctcx_asuint(-200.0f); u += stride) {
    volatile float x = ctcx_asfloat((uint32_t)u);
    if (ctcx_asuint(expf(x)) != ctcx_asuint(ctcx_port_expf(x))) ++be;
  }
  for (uint64_t u = 0; u <= 0x3f800000ull; u += stride) {
    volatile float x = ctcx_asfloat((uint32_t)u);
    if (ctcx_asuint(log1pf(x)) != ctcx_asuint(ctcx_port_log1pf(x))) ++bl;
  }
  for (uint64_t u = 0x3f800000ull; u <= (uint64_t)ctcx_asuint(65536.0f); u += stride) {
    volatile float x = ctcx_asfloat((uint32_t)u);
    if (ctcx_asuint(logf(x)) != ctcx_asuint(ctcx_port_logf(x))) ++bg;
  }
  if (out) {
    out[0] = be;
    out[1] = bl;
    out[2] = bg;
  }
  return be + bl + bg;
}

/* element-wise helpers for the Python side of the tests */
void ctcx_port_expf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = ctcx_port_expf(x[i]); }
void ctcx_port_log1pf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = ctcx_port_log1pf(x[i]); }
void ctcx_port_logf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = ctcx_port_logf(x[i]); }
void ctcx_libm_expf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = expf(x[i]); }
void ctcx_libm_log1pf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = log1pf(x[i]); }
void ctcx_libm_logf_v(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = logf(x[i]); }
