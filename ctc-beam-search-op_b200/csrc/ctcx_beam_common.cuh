// Constants shared by the fast beam kernels (ctcx_beam_v4.cuh for num_classes <= 32,
// ctcx_beam_wide.cuh for wide vocabularies): the score histogram of the next-beam selection and the
// scalar slots of the per-CTA scratch block in shared memory.
#pragma once
#include "ctcx_kernels.cuh"

namespace ctcx {

constexpr int kBinsV2 = 256;   // score-histogram bins (256 measured best: 512 costs scan work, 128 crowds the boundary bin)
constexpr int kBinsLog2V2 = 8;
constexpr int kBndFast = 32;  // boundary items handled by one warp

enum {
  kV2NCand = 0, kV2NRisk, kV2MinKey, kV2MaxKey, kV2Changed, kV2NBnd, kV2Bstar, kV2KRem, kV2E,
  kV2NNew, kV2Off0, kV2Off1, kV2Anomaly, kV2MinBase, kV2LpMin, kV2Prefix, kV2PrefixHi, kV2K,
  kV2LpMax, kV2Gap, kV2TopBin
};
enum { kV3Found = 21, kV3Cv = 22 };

}  // namespace ctcx
