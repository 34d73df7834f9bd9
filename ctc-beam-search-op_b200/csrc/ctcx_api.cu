// C-ABI shim (include/ctcx.h) over the kernels in ctcx_kernels.cuh: validation, workspace carve-up,
// launches and the host-buffer convenience entry. No global state; no CPU fallback.
//
// Reference counterparts (tensorflow_ctc_ext_beam_search_decoder/cc/kernels/
// ctc_ext_beam_search_decoder_kernels.cc): ValidateInputsGenerateOutputs :97-160, Compute :20-95,
// StoreAllDecodedSequences :163-257.
#include "../../include/ctcx.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctcx_kernels.cuh"
#include "ctcx_beam_v2.cuh"
#include "ctcx_beam_v3.cuh"
#include "ctcx_beam_wide.cuh"

namespace {

thread_local char g_cuda_err[256] = "";

bool Check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  std::snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return false;
}
#define CTCX_CUDA(call)                                 \
  do {                                                  \
    if (!Check((call), #call)) return CTCX_ERR_CUDA;    \
  } while (0)

constexpr int kMaxBeamWidth = 1024;
constexpr int kMaxClasses = 65535;
constexpr int kListCapMax = 4608;  // candidate-list entries kept in shared memory (8 B each)
constexpr uint32_t kMagic = 0x43544358u;  // "CTCX"

size_t Align256(size_t v) { return (v + 255) / 256 * 256; }

struct Tier {
  int wmax, nt;
};
Tier PickTier(int W) {
  if (W <= 32) return {32, 128};
  if (W <= 128) return {128, 256};
  if (W <= 256) return {256, 256};
  return {1024, 1024};
}

// Wide-vocabulary fast path (ctcx_beam_wide.cuh): 32 < C <= 2048 and the candidate list of the worst
// frame (beam_width rows x the Kc = min(C-1, 2*beam_width+2) best classes of the frame) fits in
// shared memory. WideKc = sorted classes the beam kernel uses, WideKs = row stride of the sorted
// arrays (Kc + one sentinel when classes are left out, rounded up to 8).
int WideKc(int W, int C) { return std::min(C - 1, 2 * W + 2); }
int WideKe(int W, int C) { return std::min(C - 1, WideKc(W, C) + 1); }
int WideKs(int W, int C) { return (WideKe(W, C) + 7) / 8 * 8; }
bool UseWide(int W, int C) {
  if (C <= 32 || C > 2048 || W > 256) return false;
  const char* impl = std::getenv("CTCX_BEAM_IMPL");
  if (impl != nullptr && std::strcmp(impl, "generic") == 0) return false;
  ctcx::BeamSmemWide lay;
  lay.Init(PickTier(W).wmax, W * WideKc(W, C), C, WideKs(W, C));
  return lay.bytes <= 200 * 1024;
}

// Everything a decode leaves behind for pack, at fixed offsets inside the caller's workspace.
struct Workspace {
  struct Header {
    uint32_t magic;
    int T, B, C, W, P;
    int real_bytes;  // 4: float32 decode, 8: float64 decode
  };
  size_t header, off, bp, fin_total, fin_kind, fin_n, flags, dec_len, ali_len, dec, ali, dec_off,
      ali_off, sizes, ptrs, stats, t_done, state, srt_pl, srt_cls, bytes;
  int Cs;  // row stride of the sorted-class arrays (0 when the wide fast path does not apply)
  void Init(int T, int B, int C, int W, int P) {
    size_t o = 0;
    const size_t b = (size_t)B, t = (size_t)T, w = (size_t)W, pp = (size_t)P;
    (void)C;
    header = o; o += Align256(sizeof(Header));
    off = o; o += Align256(t * b * 8);                           // float or double
    bp = o; o += Align256(b * t * w * 8);
    fin_total = o; o += Align256(b * pp * 8);                    // float or double
    fin_kind = o; o += Align256(b * pp * 4);
    fin_n = o; o += Align256(b * 4);
    flags = o; o += Align256(b * 4);
    dec_len = o; o += Align256(b * pp * 4);
    ali_len = o; o += Align256(b * pp * 4);
    dec = o; o += Align256(b * pp * t * 4);
    ali = o; o += Align256(b * pp * t * 4);
    dec_off = o; o += Align256(pp * b * 8);
    ali_off = o; o += Align256(pp * b * 8);
    sizes = o; o += Align256(4 * pp * 8);
    ptrs = o; o += Align256(6 * pp * 8);
    stats = o; o += Align256(16 * 4);
    t_done = o; o += Align256(b * 4);                            // streaming: frames consumed so far
    state = o; o += Align256(b * ctcx::StreamStateBytes(W));     // streaming: beam between chunks
    Cs = UseWide(W, C) ? WideKs(W, C) : 0;                       // wide fast path: best classes per frame, sorted
    srt_pl = o; o += Align256(t * b * (size_t)Cs * 4);
    srt_cls = o; o += Align256(t * b * (size_t)Cs * 2);
    bytes = o;
  }
};

template <typename R, int WMAX, int NT>
cudaError_t LaunchBeam(const ctcx::BeamParamsT<R>& p, size_t smem, cudaStream_t stream) {
  auto kern = ctcx::BeamKernelT<R, WMAX, NT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<p.B, NT, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int WMAX, int NT>
cudaError_t LaunchBeamV2(const ctcx::BeamParams& p, size_t smem, cudaStream_t stream) {
  auto kern = (p.dbg_cycles != nullptr) ? ctcx::BeamKernelV2<WMAX, NT, true> : ctcx::BeamKernelV2<WMAX, NT, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<p.B, NT, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int WMAX, int NT>
cudaError_t LaunchBeamV3(const ctcx::BeamParams& p, size_t smem, cudaStream_t stream) {
  auto kern = (p.dbg_cycles != nullptr) ? ctcx::BeamKernelV3<WMAX, NT, true> : ctcx::BeamKernelV3<WMAX, NT, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<p.B, NT, smem, stream>>>(p);
  return cudaGetLastError();
}

// Exact upcast of half-precision logits to the float32 the decoder computes in.
template <typename H>
__global__ void UpcastKernel(const H* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

template <int WMAX, int NT>
cudaError_t LaunchBeamWide(const ctcx::BeamParams& p, size_t smem, cudaStream_t stream) {
  auto kern = (p.dbg_cycles != nullptr) ? ctcx::BeamKernelWide<WMAX, NT, true> : ctcx::BeamKernelWide<WMAX, NT, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<p.B, NT, smem, stream>>>(p);
  return cudaGetLastError();
}

// flag = 1 if any entry is positive or NaN (scorer tables hold log-probabilities)
__global__ void PositiveKernel(const float* v, long long n, int* flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(v[i] <= 0.0f)) *flag = 1;
}

// Reduces the per-utterance flags to {anomaly, too_few_leaves, first bad utterance}.
__global__ void FlagsKernel(const int* flags, const int* seq_len, int B, int T, int* out) {
  // out[0] = OR of anomaly bits, out[1] = first b with too few leaves (or B), out[2] = first b with
  // sequence_length out of range (or B), out[3] = first b with negative sequence_length (or B)
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const int f = flags ? flags[b] : 0;
    if (f & 1) atomicOr(&out[0], 1);
    if (f & 2) atomicMin(&out[1], b);
    if (seq_len[b] > T) atomicMin(&out[2], b);
    if (seq_len[b] < 0) atomicMin(&out[3], b);
    if (f & 4) atomicMin(&out[4], b);  // streaming: more frames fed than the stream was sized for
  }
}

// optional per-kernel timing (ctcx_profile_enable): CUDA events on the launching stream
thread_local int g_profile = 0;
thread_local cudaEvent_t g_ev[6];
thread_local bool g_ev_ready = false;
thread_local float g_ms[5] = {0, 0, 0, 0, 0};  // lognorm, beam, trace, scan+flags, total
void ProfRecord(int i, cudaStream_t s) {
  if (!g_profile) return;
  if (!g_ev_ready) {
    for (int k = 0; k < 6; ++k) cudaEventCreate(&g_ev[k]);
    g_ev_ready = true;
  }
  cudaEventRecord(g_ev[i], s);
}

thread_local long long* g_dbg_cycles = nullptr;
int DeviceSmCount() {
  int sm_count = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  return sm_count;
}

// kernel 1: softmax normalisers of `rows` consecutive logit rows
cudaError_t LaunchLogNorm(const float* logits_dev, float* off_dev, long long rows, int C, cudaStream_t stream) {
  const int sm_count = DeviceSmCount();
  if (C <= 64) {  // thread per row, rows staged through shared memory
    long long blocks = (rows + ctcx::kLogNormRows - 1) / ctcx::kLogNormRows;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    const size_t lsm = (size_t)ctcx::kLogNormRows * (C | 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(ctcx::LogNormRowKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm);
    if (e != cudaSuccess) return e;
    ctcx::LogNormRowKernel<<<(unsigned)blocks, ctcx::kLogNormRows, lsm, stream>>>(logits_dev, off_dev, rows, C);
  } else {  // warp per row
    long long blocks = (rows + 7) / 8;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    ctcx::LogNormKernel<<<(unsigned)blocks, 256, 0, stream>>>(logits_dev, off_dev, rows, C);
  }
  return cudaGetLastError();
}

cudaError_t LaunchLogNorm(const double* logits_dev, double* off_dev, long long rows, int C, cudaStream_t stream) {
  long long blocks = (rows + 7) / 8;
  if (blocks > (long long)DeviceSmCount() * 16) blocks = (long long)DeviceSmCount() * 16;
  ctcx::LogNormKernelF64<<<(unsigned)blocks, 256, 0, stream>>>(logits_dev, off_dev, rows, C);
  return cudaGetLastError();
}

// kernel 1 for wide vocabularies: fused normaliser + the best classes of every row ordered by log-prob
template <int NI>
cudaError_t LaunchNormTopClassesNI(const float* logits_dev, float* off_dev, long long rows, int C, int blank,
                                   int Ke, int Ks, float* srt_pl, unsigned short* srt_cls, cudaStream_t stream) {
  auto kern = ctcx::NormTopClassesKernel<NI>;
  const size_t smem = (size_t)8 * NI * 32 * sizeof(float) + (size_t)8 * Ke * 8;  // exp terms + selection buffer per warp
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long blocks = std::min<long long>((rows + 7) / 8, (long long)DeviceSmCount() * 16);
  kern<<<(unsigned)blocks, 256, smem, stream>>>(logits_dev, off_dev, rows, C, blank, Ke, Ks, srt_pl, srt_cls);
  return cudaGetLastError();
}
cudaError_t LaunchNormTopClasses(const float* logits_dev, float* off_dev, long long rows, int C, int blank,
                                 int W, float* srt_pl, unsigned short* srt_cls, cudaStream_t stream) {
  const int Ke = WideKe(W, C), Ks = WideKs(W, C), ni = (C + 31) / 32;
#define CTCX_TOPC(N) LaunchNormTopClassesNI<N>(logits_dev, off_dev, rows, C, blank, Ke, Ks, srt_pl, srt_cls, stream)
  if (ni <= 2) return CTCX_TOPC(2);
  if (ni <= 4) return CTCX_TOPC(4);
  if (ni <= 8) return CTCX_TOPC(8);
  if (ni <= 16) return CTCX_TOPC(16);
  if (ni <= 32) return CTCX_TOPC(32);
  return CTCX_TOPC(64);
#undef CTCX_TOPC
}

// kernel 2: picks the beam kernel for the shape (fast path for narrow vocabularies, generic
// otherwise; CTCX_BEAM_IMPL=generic | v2 forces the generic / the previous fast kernel for A/B
// tests). Returns CTCX_OK, CTCX_ERR_UNSUPPORTED or CTCX_ERR_CUDA.
int LaunchBeamFor(ctcx::BeamParams& bp, cudaStream_t stream) {
  const int W = bp.W, C = bp.C;
  bp.kid_words = (C + 31) / 32;
  const long long full_list = (long long)W * C;
  bp.cand_cap = (full_list <= kListCapMax) ? (int)full_list : 0;
  bp.dbg_totals = nullptr;
  bp.dbg_n = nullptr;
  bp.dbg_cycles = g_dbg_cycles;  // test/measurement hook (ctcx_debug_set_cycles_buffer)
  const Tier tier = PickTier(W);
  const char* impl = std::getenv("CTCX_BEAM_IMPL");
  // a scorer table breaks the fast kernels' monotone-prefix argument: generic kernel only
  const bool want_generic = (impl != nullptr && std::strcmp(impl, "generic") == 0) || bp.lm != nullptr;
  const bool want_v2 = impl != nullptr && std::strcmp(impl, "v2") == 0 && bp.state == nullptr;
  cudaError_t e;
  if (bp.srt_pl != nullptr && bp.lm == nullptr) {  // wide-vocabulary fast path (the caller ran NormTopClassesKernel)
    bp.Kc = WideKc(W, C);
    bp.cand_cap = W * bp.Kc;
    ctcx::BeamSmemWide layw;
    layw.Init(tier.wmax, bp.cand_cap, C, bp.Cs);
    switch (tier.wmax) {
      case 32: e = LaunchBeamWide<32, 256>(bp, layw.bytes, stream); break;
      case 128: e = LaunchBeamWide<128, 256>(bp, layw.bytes, stream); break;
      default: e = LaunchBeamWide<256, 256>(bp, layw.bytes, stream); break;
    }
  } else if (!want_generic && C <= 32 && bp.cand_cap > 0 && tier.wmax <= 256) {
    if (want_v2) {
      ctcx::BeamSmemV2 lay2;
      lay2.Init(tier.wmax, bp.cand_cap);
      if (lay2.bytes > 220 * 1024) return CTCX_ERR_UNSUPPORTED;
      switch (tier.wmax) {
        case 32: e = LaunchBeamV2<32, 256>(bp, lay2.bytes, stream); break;
        case 128: e = LaunchBeamV2<128, 256>(bp, lay2.bytes, stream); break;
        default: e = LaunchBeamV2<256, 256>(bp, lay2.bytes, stream); break;
      }
    } else {
      ctcx::BeamSmemV3 lay3;
      lay3.Init(tier.wmax, bp.cand_cap);
      if (lay3.bytes > 220 * 1024) return CTCX_ERR_UNSUPPORTED;
      switch (tier.wmax) {
        case 32: e = LaunchBeamV3<32, 256>(bp, lay3.bytes, stream); break;
        case 128: e = LaunchBeamV3<128, 256>(bp, lay3.bytes, stream); break;
        default: e = LaunchBeamV3<256, 256>(bp, lay3.bytes, stream); break;
      }
    }
  } else {
    ctcx::BeamSmem lay;
    lay.Init(tier.wmax, tier.nt, C, bp.kid_words, bp.cand_cap, 4, W);
    if (lay.bytes > 220 * 1024) return CTCX_ERR_UNSUPPORTED;
    switch (tier.wmax) {
      case 32: e = LaunchBeam<float, 32, 128>(bp, lay.bytes, stream); break;
      case 128: e = LaunchBeam<float, 128, 256>(bp, lay.bytes, stream); break;
      case 256: e = LaunchBeam<float, 256, 256>(bp, lay.bytes, stream); break;
      default: e = LaunchBeam<float, 1024, 1024>(bp, lay.bytes, stream); break;
    }
  }
  return Check(e, "beam kernel launch") ? CTCX_OK : CTCX_ERR_CUDA;
}

// T = double: the generic kernel instantiated for double scores (64-bit keys)
int LaunchBeamFor(ctcx::BeamParamsT<double>& bp, cudaStream_t stream) {
  const int W = bp.W, C = bp.C;
  bp.kid_words = (C + 31) / 32;
  const long long full_list = (long long)W * C;
  bp.cand_cap = (full_list <= kListCapMax) ? (int)full_list : 0;
  bp.dbg_totals = nullptr;
  bp.dbg_n = nullptr;
  bp.dbg_cycles = nullptr;
  // double state is twice as wide: the largest tier that fits in shared memory holds 512 slots
  Tier tier = PickTier(W);
  if (tier.wmax > 256) tier = (W <= 512) ? Tier{512, 512} : Tier{1024, 1024};
  ctcx::BeamSmem lay;
  lay.Init(tier.wmax, tier.nt, C, bp.kid_words, bp.cand_cap, 8, W);
  if (lay.bytes > 220 * 1024) return CTCX_ERR_UNSUPPORTED;  // beam_width > 512, or a very wide vocabulary
  cudaError_t e;
  switch (tier.wmax) {
    case 32: e = LaunchBeam<double, 32, 128>(bp, lay.bytes, stream); break;
    case 128: e = LaunchBeam<double, 128, 256>(bp, lay.bytes, stream); break;
    case 256: e = LaunchBeam<double, 256, 256>(bp, lay.bytes, stream); break;
    case 512: e = LaunchBeam<double, 512, 512>(bp, lay.bytes, stream); break;
    default: return CTCX_ERR_UNSUPPORTED;
  }
  return Check(e, "beam kernel launch") ? CTCX_OK : CTCX_ERR_CUDA;
}

// kernels 3 + 4: trace-back of the top paths, per-path offsets and sizes
cudaError_t LaunchTraceAndScan(const ctcx::TraceParams& tp, const ctcx::ScanParams& sp, cudaStream_t stream,
                               bool profile) {
  const int W = tp.W;
  const long long walks = (long long)tp.B * tp.P;
  if (walks >= 4096) {
    // thousands of independent walks hide the latency of the dependent loads by themselves, and
    // touch one record per frame instead of whole rows
    ctcx::TraceKernel<<<(unsigned)((walks + 127) / 128), 128, 0, stream>>>(tp);
  } else {
    // one warp per (utterance, path); two blocks of 2^rows_log2 back-pointer rows per warp in
    // shared memory (about 26 KB per block)
    constexpr int kTraceWarps = 2;
    int rows_log2 = 5;
    while (rows_log2 > 0 && ((size_t)W << rows_log2) * sizeof(uint2) > 26 * 1024) --rows_log2;
    const size_t tsm = (size_t)kTraceWarps * 2 * ((size_t)W << rows_log2) * sizeof(uint2);
    auto tk = ctcx::TraceWarpKernel<kTraceWarps>;
    cudaError_t e = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
    if (e != cudaSuccess) return e;
    tk<<<(unsigned)((walks + kTraceWarps - 1) / kTraceWarps), kTraceWarps * 32, tsm, stream>>>(tp, rows_log2);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (profile) ProfRecord(3, stream);
  ctcx::ScanKernel<<<tp.P, 1024, 0, stream>>>(sp);
  return cudaGetLastError();
}

thread_local int g_err_batch = -1;
thread_local int g_err_max_time = 0;
thread_local char g_msg[160];

}  // namespace

extern "C" {

const char* ctcx_strerror(int code) {
  switch (code) {
    case CTCX_OK: return "ok";
    case CTCX_ERR_INPUTS_NOT_3D: return "inputs is not a 3-Tensor";
    case CTCX_ERR_MAX_TIME_ZERO: return "max_time is 0";
    case CTCX_ERR_SEQ_LEN_NOT_VECTOR: return "sequence_length is not a vector";
    case CTCX_ERR_SEQ_LEN_BATCH: return "len(sequence_length) != batch_size.  ";
    case CTCX_ERR_SEQ_LEN_RANGE:
      std::snprintf(g_msg, sizeof(g_msg), "sequence_length(%d) <= %d", g_err_batch, g_err_max_time);
      return g_msg;
    case CTCX_ERR_TOO_MANY_PATHS: return "requested more paths than the beam width.";
    case CTCX_ERR_TOO_FEW_LEAVES: return "Less leaves in the beam search than requested.";
    case CTCX_ERR_BAD_ARGUMENT: return "bad argument (blank_index outside [0, num_classes), negative sequence_length, or beam_width/top_paths < 1)";
    case CTCX_ERR_UNSUPPORTED: return "shape not supported by this build (see ctcx_get_limits)";
    case CTCX_ERR_WORKSPACE: return "workspace missing, misaligned, too small or not produced by ctcx_decode_f32";
    case CTCX_ERR_CUDA: return g_cuda_err;
    default: return "unknown error";
  }
}

const char* ctcx_last_cuda_error(void) { return g_cuda_err; }

int ctcx_get_limits(ctcx_limits* out) {
  if (out) {
    out->max_beam_width = kMaxBeamWidth;
    out->max_classes = kMaxClasses;
    out->max_top_paths = kMaxBeamWidth;
  }
  return 100;
}

size_t ctcx_workspace_bytes(int T, int B, int C, int W, int P) {
  if (T <= 0 || B < 0 || C <= 0 || W <= 0 || P <= 0) return 0;
  Workspace ws;
  ws.Init(T, B, C, W, P);
  return ws.bytes;
}
}  // extern "C"

namespace {
template <typename R>
int DecodeImpl(const R* logits_dev, int T, int B, int C, const int32_t* seq_len_dev, int W,
               int P, int merge_repeated, int blank_index, int blank_label, void* workspace,
               size_t workspace_bytes, void* stream_v, ctcx_sizes* sizes, int32_t* flags_out,
               const R* lm_dev = nullptr) {
  constexpr bool kF32 = (sizeof(R) == 4);
  cudaStream_t stream = (cudaStream_t)stream_v;
  // --- validation, in the reference's order (kernels.cc:111-138), then TopPaths' (decoder.h:237) ---
  if (T == 0) return CTCX_ERR_MAX_TIME_ZERO;
  if (T < 0 || B < 0 || C <= 0 || W < 1 || P < 1 || blank_index < 0 || blank_index >= C)
    return CTCX_ERR_BAD_ARGUMENT;
  if (W > kMaxBeamWidth || C > kMaxClasses) return CTCX_ERR_UNSUPPORTED;
  if (sizes == nullptr) return CTCX_ERR_BAD_ARGUMENT;
  Workspace ws;
  ws.Init(T, B, C, W, P);
  if (workspace == nullptr || workspace_bytes < ws.bytes || ((uintptr_t)workspace & 255u))
    return CTCX_ERR_WORKSPACE;
  unsigned char* base = (unsigned char*)workspace;
  int* d_stats = (int*)(base + ws.stats);

  // sequence_length lives on the device: its range check (kernels.cc:134-138) is folded into the
  // flags reduction at the end -- every kernel clamps the lengths it walks -- so that a decode has
  // ONE host round trip. Only when top_paths > beam_width (an error either way) is it evaluated
  // first, because the reference reports a bad length before TopPaths' own error (decoder.h:237).
  {
    const int init[5] = {0, B, B, B, B};
    CTCX_CUDA(cudaMemcpyAsync(d_stats, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  }
  if (B > 0 && P > W) {
    FlagsKernel<<<(B + 255) / 256, 256, 0, stream>>>(nullptr, seq_len_dev, B, T, d_stats);
    CTCX_CUDA(cudaGetLastError());
    int h[4];
    CTCX_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CTCX_CUDA(cudaStreamSynchronize(stream));
    if (h[2] < B) {
      g_err_batch = h[2];
      g_err_max_time = T;
      return CTCX_ERR_SEQ_LEN_RANGE;
    }
    if (h[3] < B) return CTCX_ERR_BAD_ARGUMENT;
    // TopPaths is reached for utterance 0 only after its frames; for B == 0 it is never reached
    return CTCX_ERR_TOO_MANY_PATHS;
  }

  Workspace::Header hdr = {kMagic, T, B, C, W, P, (int)sizeof(R)};
  CTCX_CUDA(cudaMemcpyAsync(base + ws.header, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, stream));

  if (B > 0) {
    ProfRecord(0, stream);
    bool fused_prepass = false;
    if constexpr (kF32) fused_prepass = (ws.Cs > 0 && lm_dev == nullptr);
    if (!fused_prepass) CTCX_CUDA(LaunchLogNorm(logits_dev, (R*)(base + ws.off), (long long)T * B, C, stream));

    ctcx::BeamParamsT<R> bp;
    bp.logits = logits_dev;
    bp.off = (const R*)(base + ws.off);
    bp.seq_len = seq_len_dev;
    bp.T = T; bp.B = B; bp.C = C; bp.W = W; bp.P = P;
    bp.blank_index = blank_index;
    bp.bp = (uint2*)(base + ws.bp);
    bp.fin_total = (R*)(base + ws.fin_total);
    bp.fin_kind = (int*)(base + ws.fin_kind);
    bp.fin_n = (int*)(base + ws.fin_n);
    bp.flags = (int*)(base + ws.flags);
    bp.Tcap = T; bp.t_done = nullptr; bp.state = nullptr;  // one-shot decode
    bp.srt_pl = nullptr; bp.srt_cls = nullptr; bp.Cs = ws.Cs; bp.Kc = 0;
    bp.lm = lm_dev;
    if constexpr (kF32) {
      if (fused_prepass) {  // wide vocabulary: normaliser and candidate classes in one pass over the logits
        bp.srt_pl = (const float*)(base + ws.srt_pl);
        bp.srt_cls = (const unsigned short*)(base + ws.srt_cls);
        CTCX_CUDA(LaunchNormTopClasses(logits_dev, (float*)(base + ws.off), (long long)T * B, C, blank_index, W,
                                       (float*)(base + ws.srt_pl), (unsigned short*)(base + ws.srt_cls), stream));
      }
    }
    ProfRecord(1, stream);
    const int brc = LaunchBeamFor(bp, stream);
    if (brc != CTCX_OK) return brc;
    ProfRecord(2, stream);

    ctcx::TraceParams tp;
    tp.bp = bp.bp; tp.seq_len = seq_len_dev; tp.fin_kind = bp.fin_kind;
    tp.fin_n = bp.fin_n; tp.T = T; tp.B = B; tp.W = W; tp.P = P;
    tp.merge_repeated = merge_repeated ? 1 : 0; tp.blank_label = blank_label;
    tp.dec_len = (int*)(base + ws.dec_len); tp.dec = (int*)(base + ws.dec);
    tp.ali_len = (int*)(base + ws.ali_len); tp.ali = (int*)(base + ws.ali);
    ctcx::ScanParams sp;
    sp.dec_len = tp.dec_len; sp.ali_len = tp.ali_len; sp.B = B; sp.P = P;
    sp.dec_off = (long long*)(base + ws.dec_off); sp.ali_off = (long long*)(base + ws.ali_off);
    sp.sizes = (long long*)(base + ws.sizes);
    CTCX_CUDA(LaunchTraceAndScan(tp, sp, stream, true));

    FlagsKernel<<<(B + 255) / 256, 256, 0, stream>>>(bp.flags, seq_len_dev, B, T, d_stats);
    CTCX_CUDA(cudaGetLastError());
    ProfRecord(4, stream);
  } else {
    CTCX_CUDA(cudaMemsetAsync(base + ws.sizes, 0, 4 * (size_t)P * 8, stream));
  }

  std::vector<long long> h_sizes(4 * (size_t)P);
  int h_stats[4];
  CTCX_CUDA(cudaMemcpyAsync(h_sizes.data(), base + ws.sizes, h_sizes.size() * 8, cudaMemcpyDeviceToHost, stream));
  CTCX_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
  CTCX_CUDA(cudaStreamSynchronize(stream));
  if (g_profile && B > 0) {
    for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&g_ms[k], g_ev[k], g_ev[k + 1]);
    cudaEventElapsedTime(&g_ms[4], g_ev[0], g_ev[4]);
  }
  if (h_stats[2] < B) {  // kernels.cc:134-138
    g_err_batch = h_stats[2];
    g_err_max_time = T;
    return CTCX_ERR_SEQ_LEN_RANGE;
  }
  if (h_stats[3] < B) return CTCX_ERR_BAD_ARGUMENT;  // negative sequence_length
  if (h_stats[1] < B) return CTCX_ERR_TOO_FEW_LEAVES;
  for (int p = 0; p < P; ++p) {
    if (sizes->n_decoded) sizes->n_decoded[p] = h_sizes[0 * (size_t)P + p];
    if (sizes->max_decoded) sizes->max_decoded[p] = h_sizes[1 * (size_t)P + p];
    if (sizes->n_alignment) sizes->n_alignment[p] = h_sizes[2 * (size_t)P + p];
    if (sizes->max_alignment) sizes->max_alignment[p] = h_sizes[3 * (size_t)P + p];
  }
  if (flags_out) *flags_out = h_stats[0];
  return CTCX_OK;
}

int PackImpl(const void* workspace, int T, int B, int P, int64_t* const* decoded_indices,
             int64_t* const* decoded_values, int64_t* const* decoded_shape,
             int64_t* const* alignment_indices, int64_t* const* alignment_values,
             int64_t* const* alignment_shape, void* log_probability, int real_bytes, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (workspace == nullptr || T <= 0 || B < 0 || P < 1) return CTCX_ERR_WORKSPACE;
  const unsigned char* base = (const unsigned char*)workspace;
  Workspace::Header hdr;
  CTCX_CUDA(cudaMemcpyAsync(&hdr, base, sizeof(hdr), cudaMemcpyDeviceToHost, stream));
  CTCX_CUDA(cudaStreamSynchronize(stream));
  if (hdr.magic != kMagic || hdr.T != T || hdr.B != B || hdr.P != P || hdr.real_bytes != real_bytes)
    return CTCX_ERR_WORKSPACE;
  Workspace ws;
  ws.Init(hdr.T, hdr.B, hdr.C, hdr.W, hdr.P);
  // device copy of the 6*P output pointers
  std::vector<long long*> table(6 * (size_t)P);
  for (int p = 0; p < P; ++p) {
    table[0 * (size_t)P + p] = (long long*)decoded_indices[p];
    table[1 * (size_t)P + p] = (long long*)decoded_values[p];
    table[2 * (size_t)P + p] = (long long*)decoded_shape[p];
    table[3 * (size_t)P + p] = (long long*)alignment_indices[p];
    table[4 * (size_t)P + p] = (long long*)alignment_values[p];
    table[5 * (size_t)P + p] = (long long*)alignment_shape[p];
  }
  unsigned char* wbase = (unsigned char*)workspace;
  CTCX_CUDA(cudaMemcpyAsync(wbase + ws.ptrs, table.data(), table.size() * sizeof(void*),
                            cudaMemcpyHostToDevice, stream));
  if (B > 0) {
    ctcx::PackParams pp;
    pp.dec_len = (const int*)(base + ws.dec_len); pp.dec = (const int*)(base + ws.dec);
    pp.ali_len = (const int*)(base + ws.ali_len); pp.ali = (const int*)(base + ws.ali);
    pp.dec_off = (const long long*)(base + ws.dec_off); pp.ali_off = (const long long*)(base + ws.ali_off);
    pp.sizes = (const long long*)(base + ws.sizes);
    pp.fin_total = (const void*)(base + ws.fin_total);
    pp.ptrs = (long long* const*)(base + ws.ptrs);
    pp.log_prob = log_probability;
    pp.real_bytes = real_bytes;
    pp.T = T; pp.B = B; pp.P = P;
    dim3 grid((unsigned)B, (unsigned)P);
    ctcx::PackKernel<<<grid, 128, 0, stream>>>(pp);
    CTCX_CUDA(cudaGetLastError());
  } else {
    // empty batch: shapes [0, 0]
    const long long zeros[2] = {0, 0};
    for (int p = 0; p < P; ++p) {
      CTCX_CUDA(cudaMemcpyAsync(decoded_shape[p], zeros, 16, cudaMemcpyHostToDevice, stream));
      CTCX_CUDA(cudaMemcpyAsync(alignment_shape[p], zeros, 16, cudaMemcpyHostToDevice, stream));
    }
  }
  // the pointer table lives in a host vector: a cudaMemcpyAsync from pageable memory returns only
  // after the source has been copied to the driver's staging buffer, so no synchronisation is needed
  return CTCX_OK;
}
}  // namespace

extern "C" {

int ctcx_decode_f32(const float* logits_dev, int T, int B, int C, const int32_t* seq_len_dev, int W,
                    int P, int merge_repeated, int blank_index, int blank_label, void* workspace,
                    size_t workspace_bytes, void* stream_v, ctcx_sizes* sizes, int32_t* flags_out) {
  return DecodeImpl<float>(logits_dev, T, B, C, seq_len_dev, W, P, merge_repeated, blank_index, blank_label,
                           workspace, workspace_bytes, stream_v, sizes, flags_out);
}

/* Decode with a scorer plugged into the reference's extension point (util/ctc_beam_scorer.h:31-65). */
int ctcx_decode_scorer_f32(const float* logits_dev, int T, int B, int C, const int32_t* seq_len_dev, int W,
                           int P, int merge_repeated, int blank_index, int blank_label,
                           const float* expansion_scores_dev, void* workspace, size_t workspace_bytes,
                           void* stream_v, ctcx_sizes* sizes, int32_t* flags_out) {
  if (expansion_scores_dev != nullptr && C > 0) {  // expansion scores are log-probabilities: <= 0
    cudaStream_t stream = (cudaStream_t)stream_v;
    int* d_flag = nullptr;
    if (workspace == nullptr || ((uintptr_t)workspace & 255u)) return CTCX_ERR_WORKSPACE;
    Workspace ws;
    ws.Init(T > 0 ? T : 1, B > 0 ? B : 0, C, W > 0 ? W : 1, P > 0 ? P : 1);
    if (workspace_bytes < ws.bytes) return CTCX_ERR_WORKSPACE;
    d_flag = (int*)((unsigned char*)workspace + ws.stats) + 8;
    CTCX_CUDA(cudaMemsetAsync(d_flag, 0, 4, stream));
    const long long n = (long long)(C + 1) * C;
    PositiveKernel<<<(unsigned)std::min<long long>((n + 255) / 256, 1024), 256, 0, stream>>>(expansion_scores_dev, n, d_flag);
    CTCX_CUDA(cudaGetLastError());
    int h_flag = 0;
    CTCX_CUDA(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, stream));
    CTCX_CUDA(cudaStreamSynchronize(stream));
    if (h_flag) return CTCX_ERR_BAD_ARGUMENT;
  }
  return DecodeImpl<float>(logits_dev, T, B, C, seq_len_dev, W, P, merge_repeated, blank_index, blank_label,
                           workspace, workspace_bytes, stream_v, sizes, flags_out, expansion_scores_dev);
}

int ctcx_decode_f64(const double* logits_dev, int T, int B, int C, const int32_t* seq_len_dev, int W,
                    int P, int merge_repeated, int blank_index, int blank_label, void* workspace,
                    size_t workspace_bytes, void* stream_v, ctcx_sizes* sizes, int32_t* flags_out) {
  return DecodeImpl<double>(logits_dev, T, B, C, seq_len_dev, W, P, merge_repeated, blank_index, blank_label,
                            workspace, workspace_bytes, stream_v, sizes, flags_out);
}

int ctcx_pack_f32(const void* workspace, int T, int B, int P, int64_t* const* decoded_indices,
                  int64_t* const* decoded_values, int64_t* const* decoded_shape,
                  int64_t* const* alignment_indices, int64_t* const* alignment_values,
                  int64_t* const* alignment_shape, float* log_probability, void* stream_v) {
  return PackImpl(workspace, T, B, P, decoded_indices, decoded_values, decoded_shape, alignment_indices,
                  alignment_values, alignment_shape, log_probability, 4, stream_v);
}

int ctcx_pack_f64(const void* workspace, int T, int B, int P, int64_t* const* decoded_indices,
                  int64_t* const* decoded_values, int64_t* const* decoded_shape,
                  int64_t* const* alignment_indices, int64_t* const* alignment_values,
                  int64_t* const* alignment_shape, double* log_probability, void* stream_v) {
  return PackImpl(workspace, T, B, P, decoded_indices, decoded_values, decoded_shape, alignment_indices,
                  alignment_values, alignment_shape, log_probability, 8, stream_v);
}

int ctcx_workspace_views(const void* workspace, int T, int B, int P, const int32_t** dec_len,
                         const int32_t** dec, const int32_t** ali_len, const int32_t** ali,
                         const float** logp) {
  if (workspace == nullptr) return CTCX_ERR_WORKSPACE;
  const unsigned char* base = (const unsigned char*)workspace;
  Workspace::Header hdr;
  CTCX_CUDA(cudaMemcpy(&hdr, base, sizeof(hdr), cudaMemcpyDeviceToHost));
  if (hdr.magic != kMagic || hdr.T != T || hdr.B != B || hdr.P != P) return CTCX_ERR_WORKSPACE;
  if (logp != nullptr && hdr.real_bytes != 4) return CTCX_ERR_BAD_ARGUMENT;  // float32 decodes only
  Workspace ws;
  ws.Init(hdr.T, hdr.B, hdr.C, hdr.W, hdr.P);
  if (dec_len) *dec_len = (const int32_t*)(base + ws.dec_len);
  if (dec) *dec = (const int32_t*)(base + ws.dec);
  if (ali_len) *ali_len = (const int32_t*)(base + ws.ali_len);
  if (ali) *ali = (const int32_t*)(base + ws.ali);
  if (logp) *logp = (const float*)(base + ws.fin_total);
  return CTCX_OK;
}

void ctcx_free_host(ctcx_host_result* r) {
  if (!r) return;
  auto free_list = [&](int64_t** l) {
    if (!l) return;
    for (int p = 0; p < r->top_paths; ++p) std::free(l[p]);
    std::free(l);
  };
  free_list(r->decoded_indices);
  free_list(r->decoded_values);
  free_list(r->decoded_shape);
  free_list(r->alignment_indices);
  free_list(r->alignment_values);
  free_list(r->alignment_shape);
  std::free(r->n_decoded);
  std::free(r->n_alignment);
  std::free(r->log_probability);
  std::free(r->log_probability_f64);
  std::free(r);
}

static int DecodeHostImpl(const void* logits_host, int rb, int T, int B, int C, const int32_t* seq_len_host,
                          int W, int P, int merge_repeated, int blank_index, int blank_label,
                          int device, ctcx_host_result** result) {
  if (result == nullptr) return CTCX_ERR_BAD_ARGUMENT;
  *result = nullptr;
  if (T == 0) return CTCX_ERR_MAX_TIME_ZERO;
  if (T < 0 || B < 0 || C <= 0 || W < 1 || P < 1) return CTCX_ERR_BAD_ARGUMENT;
  CTCX_CUDA(cudaSetDevice(device));
  cudaStream_t stream;
  CTCX_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t n_logits = (size_t)T * B * C;
  const size_t ws_bytes = ctcx_workspace_bytes(T, B, C, W, P);
  void* d_logits = nullptr;
  int32_t* d_seq = nullptr;
  void* d_ws = nullptr;
  unsigned char* d_out = nullptr;
  int rc = CTCX_OK;
  std::vector<int64_t> n_dec(P), max_dec(P), n_ali(P), max_ali(P);
  ctcx_sizes sizes = {n_dec.data(), max_dec.data(), n_ali.data(), max_ali.data()};
  int32_t flags = 0;
  ctcx_host_result* r = nullptr;
  std::vector<int64_t*> p_di(P), p_dv(P), p_ds(P), p_ai(P), p_av(P), p_as(P);
  size_t out_bytes = 0;
#define CTCX_TRY(call)                                        \
  do {                                                        \
    if (!Check((call), #call)) { rc = CTCX_ERR_CUDA; goto done; } \
  } while (0)
  CTCX_TRY(cudaMalloc(&d_logits, n_logits * rb + 8));
  CTCX_TRY(cudaMalloc(&d_seq, (size_t)B * 4 + 4));
  CTCX_TRY(cudaMalloc(&d_ws, ws_bytes));
  CTCX_TRY(cudaMemcpyAsync(d_logits, logits_host, n_logits * rb, cudaMemcpyHostToDevice, stream));
  CTCX_TRY(cudaMemcpyAsync(d_seq, seq_len_host, (size_t)B * 4, cudaMemcpyHostToDevice, stream));
  rc = (rb == 8) ? ctcx_decode_f64((const double*)d_logits, T, B, C, d_seq, W, P, merge_repeated, blank_index,
                                   blank_label, d_ws, ws_bytes, stream, &sizes, &flags)
                 : ctcx_decode_f32((const float*)d_logits, T, B, C, d_seq, W, P, merge_repeated, blank_index,
                                   blank_label, d_ws, ws_bytes, stream, &sizes, &flags);
  if (rc != CTCX_OK) goto done;
  {
    // one device block for all outputs, then one D2H copy
    std::vector<size_t> offs;
    size_t o = 0;
    auto take = [&](size_t n64) { size_t at = o; o += Align256(n64 * 8); return at; };
    std::vector<size_t> o_di(P), o_dv(P), o_ds(P), o_ai(P), o_av(P), o_as(P);
    for (int p = 0; p < P; ++p) {
      o_di[p] = take((size_t)n_dec[p] * 2); o_dv[p] = take((size_t)n_dec[p]); o_ds[p] = take(2);
      o_ai[p] = take((size_t)n_ali[p] * 2); o_av[p] = take((size_t)n_ali[p]); o_as[p] = take(2);
    }
    const size_t o_lp = o;
    o += Align256((size_t)B * P * rb);
    out_bytes = o;
    CTCX_TRY(cudaMalloc(&d_out, out_bytes + 256));
    for (int p = 0; p < P; ++p) {
      p_di[p] = (int64_t*)(d_out + o_di[p]); p_dv[p] = (int64_t*)(d_out + o_dv[p]); p_ds[p] = (int64_t*)(d_out + o_ds[p]);
      p_ai[p] = (int64_t*)(d_out + o_ai[p]); p_av[p] = (int64_t*)(d_out + o_av[p]); p_as[p] = (int64_t*)(d_out + o_as[p]);
    }
    rc = (rb == 8) ? ctcx_pack_f64(d_ws, T, B, P, p_di.data(), p_dv.data(), p_ds.data(), p_ai.data(), p_av.data(),
                                   p_as.data(), (double*)(d_out + o_lp), stream)
                   : ctcx_pack_f32(d_ws, T, B, P, p_di.data(), p_dv.data(), p_ds.data(), p_ai.data(), p_av.data(),
                                   p_as.data(), (float*)(d_out + o_lp), stream);
    if (rc != CTCX_OK) goto done;
    std::vector<unsigned char> h_out(out_bytes);
    CTCX_TRY(cudaMemcpyAsync(h_out.data(), d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
    CTCX_TRY(cudaStreamSynchronize(stream));
    r = (ctcx_host_result*)std::calloc(1, sizeof(ctcx_host_result));
    r->top_paths = P;
    r->flags = flags;
    r->n_decoded = (int64_t*)std::malloc(sizeof(int64_t) * P);
    r->n_alignment = (int64_t*)std::malloc(sizeof(int64_t) * P);
    auto mk = [&]() { return (int64_t**)std::calloc((size_t)P, sizeof(int64_t*)); };
    r->decoded_indices = mk(); r->decoded_values = mk(); r->decoded_shape = mk();
    r->alignment_indices = mk(); r->alignment_values = mk(); r->alignment_shape = mk();
    auto dup = [&](size_t at, size_t n64) {
      int64_t* m = (int64_t*)std::malloc(n64 * 8 + 8);
      std::memcpy(m, h_out.data() + at, n64 * 8);
      return m;
    };
    for (int p = 0; p < P; ++p) {
      r->n_decoded[p] = n_dec[p];
      r->n_alignment[p] = n_ali[p];
      r->decoded_indices[p] = dup(o_di[p], (size_t)n_dec[p] * 2);
      r->decoded_values[p] = dup(o_dv[p], (size_t)n_dec[p]);
      r->decoded_shape[p] = dup(o_ds[p], 2);
      r->alignment_indices[p] = dup(o_ai[p], (size_t)n_ali[p] * 2);
      r->alignment_values[p] = dup(o_av[p], (size_t)n_ali[p]);
      r->alignment_shape[p] = dup(o_as[p], 2);
    }
    void* lp = std::malloc((size_t)B * P * rb + 8);
    std::memcpy(lp, h_out.data() + o_lp, (size_t)B * P * rb);
    if (rb == 8) r->log_probability_f64 = (double*)lp; else r->log_probability = (float*)lp;
    *result = r;
  }
done:
#undef CTCX_TRY
  cudaFree(d_logits);
  cudaFree(d_seq);
  cudaFree(d_ws);
  cudaFree(d_out);
  cudaStreamDestroy(stream);
  return rc;
}

int ctcx_decode_host_f32(const float* logits_host, int T, int B, int C, const int32_t* seq_len_host,
                         int W, int P, int merge_repeated, int blank_index, int blank_label,
                         int device, ctcx_host_result** result) {
  return DecodeHostImpl(logits_host, 4, T, B, C, seq_len_host, W, P, merge_repeated, blank_index, blank_label,
                        device, result);
}

int ctcx_decode_host_f64(const double* logits_host, int T, int B, int C, const int32_t* seq_len_host,
                         int W, int P, int merge_repeated, int blank_index, int blank_label,
                         int device, ctcx_host_result** result) {
  return DecodeHostImpl(logits_host, 8, T, B, C, seq_len_host, W, P, merge_repeated, blank_index, blank_label,
                        device, result);
}

/* fp16 / bf16 logits (what an acoustic model's projection typically emits): upcast exactly to
 * float32 into `scratch_dev` ([T*B*C] float32, caller-allocated device memory), then decode as
 * ctcx_decode_f32 does. dtype: 0 = IEEE half, 1 = bfloat16. */
int ctcx_decode_half(const void* logits_dev, int dtype, float* scratch_dev, int T, int B, int C,
                     const int32_t* seq_len_dev, int W, int P, int merge_repeated, int blank_index,
                     int blank_label, void* workspace, size_t workspace_bytes, void* stream_v,
                     ctcx_sizes* sizes, int32_t* flags_out) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (T == 0) return CTCX_ERR_MAX_TIME_ZERO;
  if (T < 0 || B < 0 || C <= 0 || (dtype != 0 && dtype != 1) || (scratch_dev == nullptr && B > 0))
    return CTCX_ERR_BAD_ARGUMENT;
  const long long n = (long long)T * B * C;
  if (n > 0) {
    const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32);
    if (dtype == 0)
      UpcastKernel<__half><<<blocks, 256, 0, stream>>>((const __half*)logits_dev, scratch_dev, n);
    else
      UpcastKernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits_dev, scratch_dev, n);
    CTCX_CUDA(cudaGetLastError());
  }
  return ctcx_decode_f32(scratch_dev, T, B, C, seq_len_dev, W, P, merge_repeated, blank_index, blank_label,
                         workspace, workspace_bytes, stream_v, sizes, flags_out);
}

/* ---- streaming: Step / TopPaths / Reset of the reference decoder (decoder.h:39-53) ---- */

size_t ctcx_stream_workspace_bytes(int max_time_total, int B, int C, int W, int P) {
  return ctcx_workspace_bytes(max_time_total, B, C, W, P);
}

int ctcx_stream_reset(void* workspace, size_t workspace_bytes, int T_total, int B, int C, int W, int P,
                      void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (T_total <= 0 || B < 0 || C <= 0 || W < 1 || P < 1) return CTCX_ERR_BAD_ARGUMENT;
  if (W > kMaxBeamWidth || C > kMaxClasses) return CTCX_ERR_UNSUPPORTED;
  if (P > W) return CTCX_ERR_TOO_MANY_PATHS;
  Workspace ws;
  ws.Init(T_total, B, C, W, P);
  if (workspace == nullptr || workspace_bytes < ws.bytes || ((uintptr_t)workspace & 255u))
    return CTCX_ERR_WORKSPACE;
  unsigned char* base = (unsigned char*)workspace;
  Workspace::Header hdr = {kMagic, T_total, B, C, W, P, 4};
  CTCX_CUDA(cudaMemcpyAsync(base + ws.header, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, stream));
  if (B > 0) {
    // decoder.h:212-227: with zero frames consumed the kernels start from the root
    CTCX_CUDA(cudaMemsetAsync(base + ws.t_done, 0, (size_t)B * 4, stream));
    CTCX_CUDA(cudaMemsetAsync(base + ws.flags, 0, (size_t)B * 4, stream));
    CTCX_CUDA(cudaMemsetAsync(base + ws.state, 0, (size_t)B * ctcx::StreamStateBytes(W), stream));
  }
  CTCX_CUDA(cudaStreamSynchronize(stream));  // the header was staged from the stack
  return CTCX_OK;
}

int ctcx_stream_step_f32(void* workspace, int T_total, int B, int C, int W, int P, const float* logits_dev,
                         int chunk_time, const int32_t* chunk_len_dev, int blank_index, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (workspace == nullptr) return CTCX_ERR_WORKSPACE;
  if (chunk_time <= 0 || chunk_time > T_total || chunk_len_dev == nullptr || blank_index < 0 || blank_index >= C)
    return CTCX_ERR_BAD_ARGUMENT;
  if (B == 0) return CTCX_OK;
  Workspace ws;
  ws.Init(T_total, B, C, W, P);
  unsigned char* base = (unsigned char*)workspace;
  if (ws.Cs == 0) CTCX_CUDA(LaunchLogNorm(logits_dev, (float*)(base + ws.off), (long long)chunk_time * B, C, stream));
  ctcx::BeamParams bp;
  bp.logits = logits_dev;
  bp.off = (const float*)(base + ws.off);
  bp.seq_len = chunk_len_dev;  // frames of THIS chunk to consume, per utterance
  bp.T = chunk_time; bp.B = B; bp.C = C; bp.W = W; bp.P = P;
  bp.blank_index = blank_index;
  bp.bp = (uint2*)(base + ws.bp);
  bp.fin_total = (float*)(base + ws.fin_total);
  bp.fin_kind = (int*)(base + ws.fin_kind);
  bp.fin_n = (int*)(base + ws.fin_n);
  bp.flags = (int*)(base + ws.flags);
  bp.Tcap = T_total;
  bp.t_done = (int*)(base + ws.t_done);
  bp.state = base + ws.state;
  bp.srt_pl = nullptr; bp.srt_cls = nullptr; bp.Cs = ws.Cs; bp.Kc = 0; bp.lm = nullptr;
  if (ws.Cs > 0) {
    bp.srt_pl = (const float*)(base + ws.srt_pl);
    bp.srt_cls = (const unsigned short*)(base + ws.srt_cls);
    CTCX_CUDA(LaunchNormTopClasses(logits_dev, (float*)(base + ws.off), (long long)chunk_time * B, C, blank_index, W,
                                   (float*)(base + ws.srt_pl), (unsigned short*)(base + ws.srt_cls), stream));
  }
  return LaunchBeamFor(bp, stream);
}

int ctcx_stream_top_paths(void* workspace, int T_total, int B, int C, int W, int P, int merge_repeated,
                          int blank_label, void* stream_v, ctcx_sizes* sizes, int32_t* flags_out) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (workspace == nullptr || sizes == nullptr) return CTCX_ERR_WORKSPACE;
  Workspace ws;
  ws.Init(T_total, B, C, W, P);
  unsigned char* base = (unsigned char*)workspace;
  int* d_stats = (int*)(base + ws.stats);
  std::vector<long long> h_sizes(4 * (size_t)P, 0);
  int h_stats[5] = {0, B, B, B, B};
  if (B > 0) {
    CTCX_CUDA(cudaMemcpyAsync(d_stats, h_stats, sizeof(h_stats), cudaMemcpyHostToDevice, stream));
    ctcx::TraceParams tp;
    tp.bp = (const uint2*)(base + ws.bp);
    tp.seq_len = (const int*)(base + ws.t_done);  // frames consumed so far
    tp.fin_kind = (const int*)(base + ws.fin_kind);
    tp.fin_n = (const int*)(base + ws.fin_n);
    tp.T = T_total; tp.B = B; tp.W = W; tp.P = P;
    tp.merge_repeated = merge_repeated ? 1 : 0; tp.blank_label = blank_label;
    tp.dec_len = (int*)(base + ws.dec_len); tp.dec = (int*)(base + ws.dec);
    tp.ali_len = (int*)(base + ws.ali_len); tp.ali = (int*)(base + ws.ali);
    ctcx::ScanParams sp;
    sp.dec_len = tp.dec_len; sp.ali_len = tp.ali_len; sp.B = B; sp.P = P;
    sp.dec_off = (long long*)(base + ws.dec_off); sp.ali_off = (long long*)(base + ws.ali_off);
    sp.sizes = (long long*)(base + ws.sizes);
    CTCX_CUDA(LaunchTraceAndScan(tp, sp, stream, false));
    FlagsKernel<<<(B + 255) / 256, 256, 0, stream>>>((const int*)(base + ws.flags), tp.seq_len, B, T_total, d_stats);
    CTCX_CUDA(cudaGetLastError());
    CTCX_CUDA(cudaMemcpyAsync(h_sizes.data(), base + ws.sizes, h_sizes.size() * 8, cudaMemcpyDeviceToHost, stream));
    CTCX_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
    CTCX_CUDA(cudaStreamSynchronize(stream));
  }
  if (h_stats[4] < B) {  // an utterance was fed more frames than the stream holds
    g_err_batch = h_stats[4];
    g_err_max_time = T_total;
    return CTCX_ERR_SEQ_LEN_RANGE;
  }
  if (h_stats[1] < B) return CTCX_ERR_TOO_FEW_LEAVES;
  for (int p = 0; p < P; ++p) {
    if (sizes->n_decoded) sizes->n_decoded[p] = h_sizes[0 * (size_t)P + p];
    if (sizes->max_decoded) sizes->max_decoded[p] = h_sizes[1 * (size_t)P + p];
    if (sizes->n_alignment) sizes->n_alignment[p] = h_sizes[2 * (size_t)P + p];
    if (sizes->max_alignment) sizes->max_alignment[p] = h_sizes[3 * (size_t)P + p];
  }
  if (flags_out) *flags_out = h_stats[0];
  return CTCX_OK;
}

/* measurement hook (bench.py): per-kernel device times of this thread's last ctcx_decode_f32, from
 * CUDA events recorded on the launching stream. out_ms = {lognorm, beam, trace, scan, total}. */
void ctcx_profile_enable(int on) { g_profile = on; }
void ctcx_profile_get(float* out_ms) {
  for (int k = 0; k < 5; ++k) out_ms[k] = g_ms[k];
}

/* measurement hook: device buffer [B,16] int64 receiving per-phase clock64 cycles of the fast beam
 * kernel (thread 0 of every CTA, summed over frames); NULL switches it off. */
void ctcx_debug_set_cycles_buffer(long long* dev_buf) { g_dbg_cycles = dev_buf; }

/* test hook: y = f(x) element-wise with the exact device math; op 0 expf, 1 log1pf, 2 logf */
int ctcx_debug_math_f32(int op, const float* x_dev, float* y_dev, int n, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (n <= 0) return CTCX_OK;
  ctcx::MathTestKernel<<<(n + 255) / 256, 256, 0, stream>>>(op, x_dev, y_dev, n);
  CTCX_CUDA(cudaGetLastError());
  return CTCX_OK;
}


/* the double-precision twin: op 0 exp (x <= 0), 1 log (x >= 1), 2 LogSumExp(x, 0) */
int ctcx_debug_math_f64(int op, const double* x_dev, double* y_dev, int n, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (n <= 0) return CTCX_OK;
  ctcx::MathTestKernelF64<<<(n + 255) / 256, 256, 0, stream>>>(op, x_dev, y_dev, n);
  CTCX_CUDA(cudaGetLastError());
  return CTCX_OK;
}

}  // extern "C"
