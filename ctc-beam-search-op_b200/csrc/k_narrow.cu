// Translation unit of the narrow-vocabulary fast beam kernel (ctcx_beam_v4.cuh).
#include "ctcx_beam_v4.cuh"
#include "ctcx_launch.h"

namespace ctcx {

namespace {
constexpr int kListCapMax = 4608;  // candidate-list entries kept in shared memory (8 B each)

int TierOf(int W) { return (W <= 32) ? 32 : (W <= 128) ? 128 : 256; }

template <typename IN, int WMAX, bool TIMING>
LaunchStatus LaunchOne(const BeamParams& p, size_t smem, cudaStream_t stream) {
  auto kern = BeamKernelV4<IN, WMAX, 256, TIMING>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return LaunchFrom(e, "cudaFuncSetAttribute(BeamKernelV4)");
  // persistent CTAs: as many as the device keeps resident, each pulls utterances from p.queue
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
  if (e != cudaSuccess) return LaunchFrom(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(BeamKernelV4)");
  const long long resident = (long long)sms * (per_sm > 0 ? per_sm : 1);
  const unsigned grid = (unsigned)(p.B < resident ? p.B : resident);
  kern<<<grid, 256, smem, stream>>>(p);
  return LaunchFrom(cudaGetLastError(), "BeamKernelV4 launch");
}

template <typename IN>
LaunchStatus LaunchTyped(const BeamParams& p, int wmax, size_t smem, cudaStream_t stream) {
  if constexpr (sizeof(IN) == 4) {
    if (p.dbg_cycles != nullptr) {  // timing build: float32 inputs only
      switch (wmax) {
        case 32: return LaunchOne<IN, 32, true>(p, smem, stream);
        case 128: return LaunchOne<IN, 128, true>(p, smem, stream);
        default: return LaunchOne<IN, 256, true>(p, smem, stream);
      }
    }
  }
  switch (wmax) {
    case 32: return LaunchOne<IN, 32, false>(p, smem, stream);
    case 128: return LaunchOne<IN, 128, false>(p, smem, stream);
    default: return LaunchOne<IN, 256, false>(p, smem, stream);
  }
}
}  // namespace

bool NarrowFastShape(int W, int C) {
  if (C > 32 || W > 256 || (long long)W * C > kListCapMax) return false;
  BeamSmemV4 lay;
  lay.Init(TierOf(W), W * C);
  return lay.bytes <= 220 * 1024;
}

LaunchStatus LaunchBeamNarrow(BeamParams& p, int in_dtype, cudaStream_t stream) {
  if (!NarrowFastShape(p.W, p.C) || p.queue == nullptr || p.bp32 == nullptr) return {kLaunchUnsupported, cudaSuccess, ""};
  p.cand_cap = p.W * p.C;
  p.kid_words = 1;
  const int wmax = TierOf(p.W);
  BeamSmemV4 lay;
  lay.Init(wmax, p.cand_cap);
  switch (in_dtype) {
    case kInF32: return LaunchTyped<float>(p, wmax, lay.bytes, stream);
    case kInF16: return LaunchTyped<__half>(p, wmax, lay.bytes, stream);
    case kInBF16: return LaunchTyped<__nv_bfloat16>(p, wmax, lay.bytes, stream);
    default: return {kLaunchUnsupported, cudaSuccess, ""};
  }
}

}  // namespace ctcx
