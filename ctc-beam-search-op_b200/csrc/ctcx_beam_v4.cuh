// Beam kernel, fourth generation: the fast path for num_classes <= 32 (all of BASELINE's char-CTC
// shapes). Bit-identical results to the generic BeamKernel (ctcx_kernels.cuh); what changed against the
// third generation (round 1) is how the kernel is fed and scheduled:
//
//   * persistent CTAs: the grid is sized to the machine (resident CTAs per SM x SM count) and every
//     CTA pulls utterances from a queue until it is empty -- no tail wave, ragged lengths balance
//   * the kernel reads the caller's RAW logits -- float32, float16 or bfloat16, any time stride (a
//     batch shard of a larger tensor is decoded in place) -- and the softmax normaliser of
//     decoder.h:71-80 is computed here: the last warp works ONE FRAME AHEAD of the other seven (row
//     t+2 in flight in a register, frame t+1's normaliser / class order / prefix masks being built
//     while the others update frame t), so there is no normaliser kernel, no `off` array, no fp32
//     scratch for half inputs, and nothing of it on the frame's critical path
//   * an optional device word `ready` = number of leading frames that have landed lets the kernel
//     run WHILE the logits are still being copied host->device in time slabs (ctcx_decode_hostin_*)
//   * 4-byte back-pointer records (beam_width <= 256, labels < 256)
//   * the histogram scan (PD) is one bin per thread over all eight warps instead of one warp
//   * PG: survivors are scattered to their final slot first, then slot r is written by thread r
//     (state) and thread WMAX + r (hash, row info, parent table) -- linear stores, half the chain
//
// Per frame:
//   S   (last warp, for frame t+1) raw row -> x, off, per-class log-probs, classes ranked by log-prob,
//       prefix masks pref[j] = set of the j best classes
//   PA  member update, one thread per member                          decoder.h:95-143
//   PC  revisit-wipe fixed point. A row's candidates above ANY threshold v are a
//       bitmask: pref[L(v)] & ~member-children, L found by an exact search over the sorted class
//       scores (fp add is monotone); a wipe query is a sum of popcounts -- no list.
//   PB  only candidates inside the predicted score range (previous top-to-threshold gap x1.25) are
//       listed and histogrammed (256 bins); wiped rows are skipped                 decoder.h:146-187
//   PD  suffix scan -> boundary bin of the W-th item; if the prediction missed (fewer than
//       W items in range) PB/PD run again over the full admissible range
//   PE  items above the boundary bin are scattered into score groups
//   PF  the boundary bin is cut exactly (warp rank, or radix select for pathological ties)
//   PG  rank inside the score group = new slot; next beam, parent table, back-pointer records
//
// Total order of the beam: (score desc, members before children, children by (row, label)) -- the
// order the reference's sequential strict-'>' admission yields with a stable tie policy.
#pragma once
#include "ctcx_beam_common.cuh"

namespace ctcx {

// Shared-memory layout. Every offset is a compile-time constant of the tier (WMAX): the candidate list,
// the only array whose size depends on the shape, comes last -- so the kernel addresses all arrays as
// "base + immediate" and spends no registers on array pointers.
template <int WMAX, bool LM = false>
struct BeamSmemV4 {
  static constexpr size_t w = (size_t)WMAX;
  static constexpr size_t hash = 0;                       // u64 [2][WMAX]
  static constexpr size_t phash = hash + 2 * w * 8;       // u64 [2][WMAX]
  static constexpr size_t sorted = phash + 2 * w * 8;     // u64 [WMAX]   score-grouped survivors
  static constexpr size_t fin = sorted + w * 8;           // u64 [WMAX]   survivors at their final slot
  static constexpr size_t bnd = fin + w * 8;              // u64 [32]     boundary-bin items (fast path)
  static constexpr size_t exptab = bnd + kBndFast * 8;    // u64 [32]
  static constexpr size_t row = exptab + 32 * 8;          // uint4 [WMAX] {old total, old blank, label, member-children mask}
  static constexpr size_t total = row + w * 16;           // f32 [2][WMAX]
  static constexpr size_t blk = total + 2 * w * 4;
  static constexpr size_t lab = blk + 2 * w * 4;
  static constexpr size_t ab = lab + 2 * w * 4;
  static constexpr size_t an = ab + 2 * w * 4;
  static constexpr size_t label = an + 2 * w * 4;         // i32 [2][WMAX]
  static constexpr size_t m_nt = label + 2 * w * 4;       // f32 [WMAX] x 5: the members' new values
  static constexpr size_t m_nb = m_nt + w * 4;
  static constexpr size_t m_nl = m_nb + w * 4;
  static constexpr size_t m_nab = m_nl + w * 4;
  static constexpr size_t m_nan = m_nab + w * 4;
  static constexpr size_t m_key = m_nan + w * 4;          // u32 [WMAX]
  static constexpr size_t m_rec = m_key + w * 4;          // u32 [WMAX]
  static constexpr size_t m_pslot = m_rec + w * 4;        // i32 [WMAX]
  static constexpr size_t risk = m_pslot + w * 4;         // i32 [WMAX]
  static constexpr size_t risk_new = risk + w * 4;        // i32 [WMAX]
  static constexpr size_t wiped = risk_new + w * 4;       // u32 [WMAX]
  static constexpr size_t htab = wiped + w * 4;           // u32 [8*WMAX]  (hash tag << 10 | slot), 0xffffffff = empty
  static constexpr size_t hist = htab + 8 * w * 4;        // u32 [kBinsV2]
  static constexpr size_t offs = hist + kBinsV2 * 4;      // u32 [kBinsV2]
  static constexpr size_t bins2 = offs + kBinsV2 * 4;     // u32 [256]
  static constexpr size_t wtot = bins2 + 256 * 4;         // u32 [16]     per-warp histogram totals + top bins (PD)
  static constexpr size_t x = wtot + 16 * 4;              // f32 [2][32]  raw logits of the frame
  static constexpr size_t pl = x + 2 * 32 * 4;            // f32 [2][32]  x[l] - off
  static constexpr size_t pls = pl + 2 * 32 * 4;          // f32 [2][32]  class log-probs sorted descending (-inf padding)
  static constexpr size_t plh = pls + 2 * 32 * 4;         // f32 [2][8]   pls[0,4,8,...]: heads of the groups of four
  static constexpr size_t pref = plh + 2 * 8 * 4;         // u32 [2][36]  pref[j] = classes at sorted positions < j
  static constexpr size_t fsc = pref + 2 * 36 * 4;        // f32 [2][4]   {off, lp_max, lp_min, -}
  static constexpr size_t bits = fsc + 2 * 4 * 4;         // u32 [32]     S warp scratch: sort keys, then class bits
  static constexpr size_t e = bits + 32 * 4;              // f32 [32]     S warp scratch: exp terms of the normaliser
  static constexpr size_t scal = e + 32 * 4;              // 32 x 4 B
  static constexpr size_t lm = (scal + 32 * 4 + 15) / 16 * 16;  // f32 [33][32]  scorer table (LM kernels only), row = previous label + 1
  static constexpr size_t list = lm + (LM ? 33 * 32 * 4 : 0);     // uint2 [cand_cap] {score key, (row<<16)|label}
  static constexpr size_t Bytes(int cand_cap) { return (list + (size_t)cand_cap * 8 + 15) / 16 * 16; }
};

enum { kV4Utt = 23, kV4Abort = 24 };  // scalar slots in addition to the kV2* / kV3* ones

// MINB = resident CTAs per SM the register allocation is tuned for: 4 (64 registers) for batches that
// fill the machine, 2 (128 registers: more loads in flight, no re-materialisation) for the latency
// regime of at most two utterances per SM.
// LM = a scorer table is plugged in (util/ctc_beam_scorer.h:31-65 as a [C+1, C] table of expansion
// log-probabilities <= 0): a child's base is old total (or old blank) + lm[previous label + 1][label],
// which breaks the "prefix of the sorted classes" shortcut -- the candidate mask of a row is then built
// by testing all 32 classes.
template <typename IN, int WMAX, int NT, bool TIMING, int MINB, bool LM = false>
__global__ void __launch_bounds__(NT, (TIMING ? 1 : MINB)) BeamKernelV4(BeamParams p) {
  static_assert(NT >= WMAX && NT >= kBinsV2, "one thread per beam slot and per histogram bin");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NWARP = NT / 32;
  constexpr int TS = 8 * WMAX;  // parent look-up table slots (load factor <= 1/8: ~97% of the look-ups miss)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.W, C = p.C, T = p.T, B = p.B, blank = p.blank_index;
  const bool s_warp = (warp == NWARP - 1);

  using lay = BeamSmemV4<WMAX, LM>;
  unsigned long long* s_hash = (unsigned long long*)(smem + lay::hash);
  unsigned long long* s_phash = (unsigned long long*)(smem + lay::phash);
  unsigned long long* s_sorted = (unsigned long long*)(smem + lay::sorted);
  unsigned long long* s_fin = (unsigned long long*)(smem + lay::fin);
  unsigned long long* s_bnd = (unsigned long long*)(smem + lay::bnd);
  unsigned long long* s_exptab = (unsigned long long*)(smem + lay::exptab);
  uint4* s_row = (uint4*)(smem + lay::row);
  uint2* c_list = (uint2*)(smem + lay::list);
  float* s_total = (float*)(smem + lay::total);
  float* s_blk = (float*)(smem + lay::blk);
  float* s_lab = (float*)(smem + lay::lab);
  float* s_ab = (float*)(smem + lay::ab);
  float* s_an = (float*)(smem + lay::an);
  int* s_label = (int*)(smem + lay::label);
  float* m_nt = (float*)(smem + lay::m_nt);
  float* m_nb = (float*)(smem + lay::m_nb);
  float* m_nl = (float*)(smem + lay::m_nl);
  float* m_nab = (float*)(smem + lay::m_nab);
  float* m_nan = (float*)(smem + lay::m_nan);
  unsigned* m_key = (unsigned*)(smem + lay::m_key);
  unsigned* m_rec = (unsigned*)(smem + lay::m_rec);
  int* m_pslot = (int*)(smem + lay::m_pslot);
  int* s_risk = (int*)(smem + lay::risk);
  int* s_risk_new = (int*)(smem + lay::risk_new);
  unsigned* s_wiped = (unsigned*)(smem + lay::wiped);
  unsigned* s_htab = (unsigned*)(smem + lay::htab);
  unsigned* s_hist = (unsigned*)(smem + lay::hist);
  unsigned* s_offs = (unsigned*)(smem + lay::offs);
  unsigned* s_bins2 = (unsigned*)(smem + lay::bins2);
  unsigned* s_wtot = (unsigned*)(smem + lay::wtot);
  float* s_xb = (float*)(smem + lay::x);
  float* s_plb = (float*)(smem + lay::pl);
  float* s_plSb = (float*)(smem + lay::pls);
  float* s_plHb = (float*)(smem + lay::plh);
  unsigned* s_prefb = (unsigned*)(smem + lay::pref);
  float* s_fsc = (float*)(smem + lay::fsc);
  unsigned* s_bits = (unsigned*)(smem + lay::bits);
  float* s_e = (float*)(smem + lay::e);
  volatile int* sc = (volatile int*)(smem + lay::scal);
  int* sci = (int*)(smem + lay::scal);
  unsigned* scu = (unsigned*)(smem + lay::scal);

  LoadExpTable(s_exptab, tid, NT);
  const float* s_lm = (const float*)(smem + lay::lm);
  unsigned valid_mask = 0u;  // LM: the non-blank classes
  if constexpr (LM) {
    float* w_lm = (float*)(smem + lay::lm);
    for (int i = tid; i < 33 * 32; i += NT) {
      const int r = i >> 5, l = i & 31;
      w_lm[i] = (r <= C && l < C) ? p.lm[(size_t)r * C + l] : 0.0f;
    }
    valid_mask = ((C >= 32) ? 0xffffffffu : ((1u << C) - 1u)) & ~(1u << blank);
  }

  // thread -> (row, class slice) mapping of the candidate pass
  // (latency regime: all threads share the rows, two per row at the 128-slot tier; throughput regime:
  // one thread per row -- the candidate search runs on half as many warps, fewer instructions in total)
  constexpr int PARTS = (MINB >= 4 && !TIMING) ? 1 : NT / WMAX;  // threads per row
  constexpr int CP = 32 / PARTS;    // classes per thread
  static_assert(NT % WMAX == 0 && 32 % PARTS == 0, "row/class tiling");
  const int prow = tid / PARTS, pbase = (tid % PARTS) * CP;
  // thread -> (slot, role) mapping of the beam write-out (PG): role 0 = state + record, role 1 = hash side
  constexpr int PGS = (NT >= 2 * WMAX) ? 2 : 1;
  const int pg_slot = tid % WMAX, pg_role = tid / WMAX;

  // optional per-phase clock64 instrumentation (thread 0), compiled out of the production kernel
  long long cyc[TIMING ? 24 : 1] = {0};
  long long tprev = 0;
  const bool timing = TIMING && (p.dbg_cycles != nullptr) && tid == 0;
#define CTCX_TICK(i)                      \
  if (TIMING && timing) {                 \
    const long long now_ = clock64();     \
    cyc[TIMING ? (i) : 0] += now_ - tprev; \
    tprev = now_;                         \
  }

  int ready_known = (p.ready != nullptr) ? 0 : 0x7fffffff;  // S warp: frames known to have landed

  // Tasks: (time slice, utterance), slice-major -- task q is slice q / B of utterance q % B. With one slice
  // per utterance (batches that fit the resident CTAs, streaming) a task is a whole utterance. Larger
  // batches are cut into p.n_slices slices of p.slice_frames frames so that the CTAs stay evenly loaded
  // to the end (no tail wave, ragged lengths balance): the beam is handed from slice to slice through
  // the per-utterance state block in HBM, and a slice waits for its predecessor's release of
  // p.progress[b]. Every predecessor has a smaller task number, i.e. it is already owned by a resident
  // CTA that waits for nothing later: the wait cannot deadlock.
  const int n_tasks = B * p.n_slices;
  for (;;) {  // ---- persistent loop: one task per iteration ----
    if (tid == 0) sci[kV4Utt] = atomicAdd(p.queue, 1);
    __syncthreads();
    const int task = sci[kV4Utt];
    if (task >= n_tasks) break;
    const int slice = task / B, b = task - slice * B;

    // streaming: frames already consumed by earlier calls; this call contributes up to Lall more
    const int t_done = (p.t_done != nullptr) ? p.t_done[b] : 0;
    const int Lall = max(0, min(p.seq_len[b], p.Tcap - t_done));
    const int t0 = slice * p.slice_frames;  // first frame of this task (within this call's logits)
    if (slice > 0 && t0 >= Lall) {  // the utterance ended in an earlier slice
      __syncthreads();
      continue;
    }
    const int L = min(Lall, t0 + p.slice_frames) - t0;  // frames of this task
    const bool last_slice = (t0 + L == Lall);
    const bool resume = (p.state != nullptr) && (t_done > 0 || slice > 0);
    const size_t row0 = (size_t)b * C;  // element offset of (t = 0, b) in the logits tensor
    if (slice > 0) {  // wait for the previous slice of this utterance (acquire), then read its state
      if (tid == 0) {
        int done = 0;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(done) : "l"(p.progress + b) : "memory");
          if (done >= slice) break;
          __nanosleep(100);
        }
      }
      __syncthreads();
    }

    // S warp: raw logit (lane = class) of frame t, as float. Frames arrive in order; wait for the copy
    // that is still in flight (p.ready counts the frames that have landed).
    auto load_row = [&](int t) -> float {
      if (t >= ready_known) {
        int r = 0;
        if (lane == 0) {
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(r) : "l"(p.ready) : "memory");
            if (r > t) break;
            if (clock64() - t0 > 8000000000ll) { sci[kV4Abort] = 1; r = 0x7fffffff; break; }  // ~4 s: copy lost
            __nanosleep(200);
          }
        }
        ready_known = __shfl_sync(kFull, r, 0);
      }
      return (lane < C) ? LoadLogit<IN>(p.logits, (size_t)t * (size_t)p.tstride + row0 + lane) : 0.0f;
    };
    // The S warp prepares frame t+1 in three stages, each placed where the warp has nothing else to do:
    // S1 (while the other warps are in PA): max and the exp-sum of the softmax normaliser (decoder.h:71-80)
    auto prepare1 = [&](float xr, float& mx_out) -> float {
      const bool in_row = lane < C;
      const float mx = UnKey(__reduce_max_sync(kFull, in_row ? KeyOf(xr) : 0u));
      s_e[lane] = in_row ? ExpfExact(__fsub_rn(xr, mx), s_exptab) : 0.0f;
      __syncwarp();
      float sum = 0.0f;  // index order, as the reference sums (trailing +0 terms leave it unchanged)
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(s_e + i);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, v.x), v.y), v.z), v.w);
      }
      mx_out = mx;
      return sum;
    };
    // S2 (while the other warps list the candidates, PB): the normaliser, per-class log-probs, classes
    // ranked by log-prob. Equal keys keep lane order; the order inside a tie never matters (a prefix of
    // the sorted classes never ends inside a group of equal scores). Returns the lane's rank.
    auto prepare2 = [&](float xr, float mx, float sum, int buf) -> int {
      const float off = __fadd_rn(mx, LogfExact(sum));
      const bool lane_ok = (lane < C) && (lane != blank);
      const float pl_lane = lane_ok ? __fsub_rn(xr, off) : 0.0f;
      s_xb[buf * 32 + lane] = xr;
      s_plb[buf * 32 + lane] = pl_lane;
      const unsigned key = lane_ok ? KeyOf(pl_lane) : 0u;  // blank / padding sort last
      s_bits[lane] = key;  // the keys of the row, read back as broadcasts
      if (lane == 0) s_fsc[buf * 4 + 0] = off;
      __syncwarp();
      int rank = 0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint4 k = *reinterpret_cast<const uint4*>(s_bits + i);
        rank += (k.x > key) ? 1 : 0;
        rank += (k.y > key) ? 1 : 0;
        rank += (k.z > key) ? 1 : 0;
        rank += (k.w > key) ? 1 : 0;
      }
      rank += __popc(__match_any_sync(kFull, key) & ((1u << lane) - 1u));
      s_plSb[buf * 32 + rank] = lane_ok ? pl_lane : NegInf();
      if ((rank & 3) == 0) s_plHb[buf * 8 + (rank >> 2)] = lane_ok ? pl_lane : NegInf();
      return rank;
    };
    // S3 (while the other warps write the next beam, PG): prefix masks of the sorted classes
    auto prepare3 = [&](int rank, int buf) {
      const bool lane_ok = (lane < C) && (lane != blank);
      unsigned* bpref = s_prefb + buf * 36;
      __syncwarp();
      s_bits[rank] = lane_ok ? (1u << lane) : 0u;
      __syncwarp();
      unsigned incl = s_bits[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl |= v;
      }
      bpref[lane + 1] = incl;
      const int cv = __popc(__ballot_sync(kFull, lane_ok));
      __syncwarp();
      if (lane == 0) {
        bpref[0] = 0u;
        s_fsc[buf * 4 + 1] = (cv > 0) ? s_plSb[buf * 32] : NegInf();
        s_fsc[buf * 4 + 2] = (cv > 0) ? s_plSb[buf * 32 + cv - 1] : 0.0f;
      }
    };

    // ---- initial state: the root (decoder.h:212-227) ----
    for (int i = tid; i < TS; i += NT) s_htab[i] = 0xffffffffu;
    for (int i = tid; i < WMAX; i += NT) {
      s_wiped[i] = 0u;
      s_row[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < kBinsV2; i += NT) s_hist[i] = 0u;
    if (tid == 0) {
      if (!resume) {
        s_total[0] = 0.0f;
        s_blk[0] = 0.0f;
        s_lab[0] = NegInf();
        s_ab[0] = 0.0f;  // empty alignment with probability 1 (entry.h:204-209)
        s_an[0] = NegInf();
        s_label[0] = -1;
        s_hash[0] = kRootHash;
        s_phash[0] = 0ull;
      }
      sci[kV2Anomaly] = 0;
      sci[kV2NCand] = 0;
      sci[kV2NRisk] = 0;
      scu[kV2MinKey] = 0xffffffffu;
      scu[kV2MaxKey] = 0u;
      sci[kV2NBnd] = 0;
      scu[kV2MinBase] = 0xffffffffu;
      scu[kV2Gap] = 0u;
      sci[kV3Found] = 0;
      sci[kV4Abort] = 0;
    }
    int n = 1;
    float xr_next = 0.0f;  // S warp: raw row of the frame after the one being prepared
    float s_xr = 0.0f, s_mx = 0.0f, s_sum = 0.0f;  // S warp: the row being prepared, between the stages
    int s_rank = 0;
    if (s_warp && L > 0) {
      const float x0 = load_row(t0);
      if (L > 1) xr_next = load_row(t0 + 1);
      const float sum0 = prepare1(x0, s_mx);
      prepare3(prepare2(x0, s_mx, sum0, 0), 0);
    }
    __syncthreads();
    int carried = 0;  // flag bits handed on from the previous slice / call
    if (resume) {  // beam as the previous slice left it (buffer 0: local frame 0 reads buffer 0)
      StreamView sv(p.state + (size_t)b * StreamStateBytes(W), W);
      n = __ldcg(&sv.hdr->n);
      carried = __ldcg(&sv.hdr->flags);
      for (int i = tid; i < n; i += NT) {  // (L2 loads: another SM wrote the block)
        s_total[i] = __ldcg(sv.total + i); s_blk[i] = __ldcg(sv.blk + i); s_lab[i] = __ldcg(sv.lab + i);
        s_ab[i] = __ldcg(sv.ab + i); s_an[i] = __ldcg(sv.an + i); s_label[i] = __ldcg(sv.label + i);
        s_hash[i] = __ldcg(sv.hash + i); s_phash[i] = __ldcg(sv.phash + i);
      }
      if (tid == 0) {
        scu[kV2Gap] = __ldcg(&sv.hdr->gap);
        sci[kV2Anomaly] = carried & 1;
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += NT) {  // row info + parent look-up table of the initial beam
      s_row[i] = make_uint4(__float_as_uint(s_total[i]), __float_as_uint(s_blk[i]), (unsigned)s_label[i], 0u);
      const unsigned long long hsh = s_hash[i];
      unsigned h = (unsigned)hsh & (TS - 1);
      const unsigned entry = ((unsigned)(hsh >> 42) << 10) | (unsigned)i;
      while (atomicCAS(&s_htab[h], 0xffffffffu, entry) != 0xffffffffu) h = (h + 1) & (TS - 1);
    }
    __syncthreads();

    if (timing) tprev = clock64();
    int t_end = L;  // frames actually consumed (smaller only if the input copy was lost)
    for (int t = 0; t < L; ++t) {
      const int cur = t & 1, nxt = cur ^ 1;
      const float* x = s_xb + cur * 32;
      const float* s_pl = s_plb + cur * 32;
      const float* s_plS = s_plSb + cur * 32;
      const float* s_plH = s_plHb + cur * 8;
      const unsigned* s_pref = s_prefb + cur * 36;
      const float off = s_fsc[cur * 4 + 0];
      const float* o_total = s_total + cur * WMAX;
      const float* o_blk = s_blk + cur * WMAX;
      const float* o_lab = s_lab + cur * WMAX;
      const float* o_ab = s_ab + cur * WMAX;
      const float* o_an = s_an + cur * WMAX;
      const int* o_label = s_label + cur * WMAX;
      const unsigned long long* o_hash = s_hash + cur * WMAX;
      const unsigned long long* o_phash = s_phash + cur * WMAX;

      // ---- S: the last warp prepares frame t+1 (and puts row t+2 in flight) while the others run PA ----
      if (s_warp && t + 1 < L) {
        s_xr = xr_next;
        if (t + 2 < L) xr_next = load_row(t0 + t + 2);
        s_sum = prepare1(s_xr, s_mx);
      }
      const float xb = x[blank];
      const float pb = __fsub_rn(xb, off);
      CTCX_TICK(7)  // frame setup
      // ---- PA: update the existing members (decoder.h:95-143) ----
      unsigned my_key = 0u;
      bool suspect = false;
      if (tid < n) {
        const int i = tid;
        const int lbl = o_label[i];
        int pslot = -1;
        float v_nl = o_lab[i], v_an = NegInf();
        float rescore = NegInf();  // what the parent's re-score of this member would be (decoder.h:172-182)
        unsigned an_kind = kAnNone, an_src = 0xffu;
        if (lbl >= 0) {
          const unsigned long long ph = o_phash[i];
          unsigned h = (unsigned)ph & (TS - 1);
          const unsigned tag = (unsigned)(ph >> 42);  // 22 hash bits disjoint from the table index
          for (;;) {  // parent->Active() <=> the parent prefix is in the beam (decoder.h:97)
            const unsigned e0 = s_htab[h], e1 = s_htab[(h + 1) & (TS - 1)];  // two probes in flight
            if (e0 == 0xffffffffu) break;
            if ((e0 >> 10) == tag && o_hash[e0 & 1023u] == ph) { pslot = (int)(e0 & 1023u); break; }
            if (e1 == 0xffffffffu) break;
            if ((e1 >> 10) == tag && o_hash[e1 & 1023u] == ph) { pslot = (int)(e1 & 1023u); break; }
            h = (h + 2) & (TS - 1);
          }
          CTCX_TICK(8)  // parent look-up
          const float xl = x[lbl];
          const float pl = __fsub_rn(xl, off);
          const float self_an = __fadd_rn(o_an[i], pl);
          if (pslot >= 0) {
            const bool same = (lbl == o_label[pslot]);
            float base = same ? o_blk[pslot] : o_total[pslot];
            if constexpr (LM) base = __fadd_rn(base, s_lm[(o_label[pslot] + 1) * 32 + lbl]);  // decoder.h:103,114
            v_nl = __fsub_rn(__fadd_rn(LogSumExp(o_lab[i], base, s_exptab), xl), off);
            rescore = __fadd_rn(pl, base);
            v_an = __fadd_rn(o_ab[pslot], pl);
            an_kind = kAnParAb;
            an_src = (unsigned)pslot;
            if (!same) {
              const float c2 = __fadd_rn(o_an[pslot], pl);
              if (c2 > v_an) { v_an = c2; an_kind = kAnParAn; }
            }
            if (self_an > v_an) { v_an = self_an; an_kind = kAnSelfAn; an_src = (unsigned)i; }
          } else {
            v_nl = __fadd_rn(o_lab[i], pl);
            v_an = self_an;
            an_kind = kAnSelfAn;
            an_src = (unsigned)i;
          }
        }
        CTCX_TICK(9)  // first LSE + alignment candidates
        const float v_nb = __fsub_rn(__fadd_rn(o_total[i], xb), off);
        const float c1 = __fadd_rn(o_ab[i], pb), c2 = __fadd_rn(o_an[i], pb);
        const unsigned ab_kind = (c2 > c1) ? kAbFromAn : kAbFromAb;
        const float v_nt = LogSumExp(v_nb, v_nl, s_exptab);
        CTCX_TICK(10)  // second LSE
        m_nt[i] = v_nt;
        m_nb[i] = v_nb;
        m_nl[i] = v_nl;
        m_nab[i] = (c2 > c1) ? c2 : c1;
        m_nan[i] = v_an;
        my_key = KeyOf(v_nt);
        m_key[i] = my_key;
        m_rec[i] = PackRec32((unsigned)i, an_src, ab_kind, an_kind, (unsigned)(lbl & 0xff));
        m_pslot[i] = pslot;
        // Precondition of the one event that is reported instead of modelled (DESIGN.md "Known deviation"):
        // were this member evicted and then re-scored by its parent, rounding would put the re-score ABOVE
        // the member's own total, so the reference could accept it again (decoder.h:189-199; it does so only
        // when the beam bottom ties with that total). Mathematically total >= re-score always. The utterance
        // is flagged if such a member then drops out of the beam (after the selection, below).
        suspect = KeyOf(rescore) > my_key;
        if (pslot >= 0) {
          atomicOr(&s_row[pslot].w, 1u << lbl);
          if (pslot < i) {
            const int q = atomicAdd(&sci[kV2NRisk], 1);
            s_risk[q] = i;
          }
        }
      }
      CTCX_TICK(11)  // stores + atomics
      {
        const unsigned kmin = __reduce_min_sync(kFull, (tid < n) ? my_key : 0xffffffffu);
        const unsigned kmax = __reduce_max_sync(kFull, (tid < n) ? my_key : 0u);
        if (lane == 0 && warp * 32 < n) {
          atomicMin(&scu[kV2MinKey], kmin);
          atomicMax(&scu[kV2MaxKey], kmax);
        }
      }
      CTCX_TICK(12)  // min/max reduction
      if (__builtin_expect(n < W, 0)) {  // beam not full: every finite child is admissible; bound the score range
        unsigned kb = 0xffffffffu;
        if (tid < n) {
          const float ob = o_blk[tid], ot = o_total[tid];
          if (ot > NegInf()) kb = KeyOf((ob > NegInf()) ? fminf(ot, ob) : ot);
        }
        kb = __reduce_min_sync(kFull, kb);
        if (lane == 0 && warp * 32 < n) atomicMin(&scu[kV2MinBase], kb);
      }
      __syncthreads();
      CTCX_TICK(0)  // PA
      const int n_risk = sci[kV2NRisk];

      // Candidates of one row above a threshold, as a class bitmask. The classes are sorted by
      // log-prob and fp addition is monotone, so "(x_l - off) + old total > thr" holds exactly for a
      // prefix of the sorted order: 2-round exact search, then drop the classes that are already
      // members (decoder.h:168) and re-test the repeated label, whose base is the old blank
      // probability (decoder.h:172-177).
      auto cand_mask = [&](const uint4 ri, const float thr, const int c_lo = 0, const int c_n = 32) -> unsigned {
        const float ot = __uint_as_float(ri.x);
        if constexpr (LM) {  // every class of [c_lo, c_lo + c_n) on its own: score = pl + (base + lm)  (decoder.h:171-182)
          const float ob = __uint_as_float(ri.y);
          const int lb = (int)ri.z;
          const float* lmrow = s_lm + (lb + 1) * 32;
          unsigned m = 0u;
#pragma unroll 2
          for (int g = c_lo; g < c_lo + c_n; g += 4) {
            const float4 q = *reinterpret_cast<const float4*>(s_pl + g);
            const float4 e = *reinterpret_cast<const float4*>(lmrow + g);
            m |= (__fadd_rn(q.x, __fadd_rn((g == lb) ? ob : ot, e.x)) > thr) ? (1u << g) : 0u;
            m |= (__fadd_rn(q.y, __fadd_rn((g + 1 == lb) ? ob : ot, e.y)) > thr) ? (2u << g) : 0u;
            m |= (__fadd_rn(q.z, __fadd_rn((g + 2 == lb) ? ob : ot, e.z)) > thr) ? (4u << g) : 0u;
            m |= (__fadd_rn(q.w, __fadd_rn((g + 3 == lb) ? ob : ot, e.w)) > thr) ? (8u << g) : 0u;
          }
          return m & valid_mask & ~ri.w;
        }
        // prefix length = number of sorted scores above thr (the predicate is monotone): first the
        // heads of the 8 groups of 4, then the group itself. -inf padding never passes.
        const float4 ha = *reinterpret_cast<const float4*>(s_plH), hb = *reinterpret_cast<const float4*>(s_plH + 4);
        int g = 0;
        g += (__fadd_rn(ha.x, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(ha.y, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(ha.z, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(ha.w, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(hb.x, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(hb.y, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(hb.z, ot) > thr) ? 1 : 0;
        g += (__fadd_rn(hb.w, ot) > thr) ? 1 : 0;
        int pos = 0;
        if (g > 0) {  // group g-1 is the last one whose head passes
          const float4 q = *reinterpret_cast<const float4*>(s_plS + 4 * (g - 1));
          pos = 4 * (g - 1) + 1;
          pos += (__fadd_rn(q.y, ot) > thr) ? 1 : 0;
          pos += (__fadd_rn(q.z, ot) > thr) ? 1 : 0;
          pos += (__fadd_rn(q.w, ot) > thr) ? 1 : 0;
        }
        unsigned m = s_pref[pos] & ~ri.w;
        const int lb = (int)ri.z;
        if (lb >= 0 && ((m >> lb) & 1u) && !(__fadd_rn(s_pl[lb], __uint_as_float(ri.y)) > thr)) m &= ~(1u << lb);
        return m;
      };

      // ---- PC: revisit-wipe fixed point ----
      if (__builtin_expect(n_risk > 0, 0)) {
        for (;;) {
          for (int q = warp; q < n_risk; q += NWARP) {  // one warp per at-risk member
            const int m = s_risk[q];
            const int pslot = m_pslot[m];
            int verdict = 0;
            if (!s_wiped[pslot]) {
              const unsigned vkey = m_key[m];
              const float v = m_nt[m];
              int cnt = 0;
              for (int j = lane; j < n; j += 32) {  // members ranking before m
                const unsigned kj = m_key[j];
                cnt += (kj > vkey || (kj == vkey && j < m)) ? 1 : 0;
              }
              // children visited before the parent reaches label(m), from rows that are not wiped
              const unsigned below = (1u << o_label[m]) - 1u;
#pragma unroll 1
              for (int r0 = 0; r0 <= pslot; r0 += 32) {
                const int r = r0 + lane;
                if (r <= pslot && !s_wiped[r]) {
                  unsigned mk = cand_mask(s_row[r], v);
                  if (r == pslot) mk &= below;
                  cnt += __popc(mk);
                }
              }
              cnt = __reduce_add_sync(kFull, cnt);
              verdict = (cnt >= W) ? 1 : 0;
            }
            if (lane == 0) s_risk_new[q] = verdict;
          }
          __syncthreads();
          // every thread inspects the (few) verdicts itself: no flag, no extra barrier when nothing
          // changes -- the common case
          bool changed = false;
          int min_wiped = 0x7fffffff, max_parent = -1;
          for (int q = 0; q < n_risk; ++q) {
            const int m = s_risk[q];
            const unsigned v = (unsigned)s_risk_new[q];
            changed |= (s_wiped[m] != v);
            if (v) min_wiped = min(min_wiped, m);
            max_parent = max(max_parent, m_pslot[m]);
          }
          if (!changed) break;
          __syncthreads();  // all reads of s_wiped are done
          for (int q = tid; q < n_risk; q += NT) s_wiped[s_risk[q]] = (unsigned)s_risk_new[q];
          __syncthreads();
          // a query only counts rows up to its parent's: if every wiped row lies beyond every parent
          // row, no count (and no parent) is affected and the verdicts are final
          if (min_wiped > max_parent) break;
        }
      }
      CTCX_TICK(2)  // PC

      // ---- S, second and third part: normaliser, log-probs, class ranks and prefix masks of frame t+1,
      // while the others list the candidates ----
      if (s_warp && t + 1 < L) {
        s_rank = prepare2(s_xr, s_mx, s_sum, nxt);
        prepare3(s_rank, nxt);
      }

      // ---- PB / PD: list + histogram of the items in the score range, boundary bin ----
      const unsigned minkey_m = scu[kV2MinKey];
      const float lp_max = s_fsc[cur * 4 + 1];
      unsigned lo_true;  // no item lies below this key
      if (n == W) {
        lo_true = minkey_m;  // decoder.h:151-155: nothing at or below the W-th member total is admitted
      } else {
        const unsigned kb = scu[kV2MinBase];
        unsigned lo_c = minkey_m;
        if (kb != 0xffffffffu) lo_c = KeyOf(__fadd_rn(UnKey(kb), s_fsc[cur * 4 + 2]));
        lo_true = max(min(minkey_m, lo_c), kKeyNegInf);
      }
      // every item is <= max(best member, best possible child); old totals are sorted, slot 0 is the max
      const unsigned hi_key = max(scu[kV2MaxKey], KeyOf(__fadd_rn(lp_max, o_total[0])));
      unsigned lo_key = lo_true;
      int shift = 0;
      bool clamped = false;
      int n_cand = 0;
      auto bucket_of = [&](unsigned key) -> int { return (key > lo_key) ? (int)((key - lo_key) >> shift) : 0; };
      for (int attempt = 0; attempt < 2; ++attempt) {
        // Score range of the histogram. Survivors crowd near the top while the admissible range reaches
        // far below, so the first attempt only looks at [hi - 1.25*gap - 64, hi], gap = the previous
        // frame's top-to-threshold distance; if fewer than W items live there the second attempt
        // takes the whole admissible range. The prediction affects speed only.
        lo_key = lo_true;
        if (attempt == 0 && n == W) {
          const unsigned gap = scu[kV2Gap];
          const unsigned long long reach = (5ull * gap) / 4ull + 64ull;  // measured: 1.0-1.25 x gap is best, below 1.0 it always misses
          if (gap != 0u && reach < (unsigned long long)(hi_key - lo_true)) lo_key = hi_key - (unsigned)reach;
        }
        clamped = (lo_key != lo_true);
        const unsigned span = hi_key - lo_key;
        shift = max(0, (32 - __clz(span | 1u)) - kBinsLog2V2);  // (key - lo) >> shift < kBinsV2
        const float thr = (n == W) ? UnKey(lo_key) : NegInf();  // listed children: score > thr
        const bool member_in = !clamped || my_key > lo_key;
        CTCX_TICK(16)  // PB: range

        // PB pass 1: admissible classes of this thread's (row, class slice). A warp whose rows all lie
        // beyond the beam (the S warp at beam widths up to 112, for one) has nothing to list.
        const bool warp_has_rows = ((warp * 32) / PARTS) < n;
        unsigned mymask = 0u;
        float r_ot = 0.0f, r_ob = 0.0f;
        int r_label = -1;
        if (prow < n && !s_wiped[prow]) {
          const uint4 ri = s_row[prow];
          r_ot = __uint_as_float(ri.x);
          r_ob = __uint_as_float(ri.y);
          r_label = (int)ri.z;
          if (__fadd_rn(lp_max, r_ot) > thr) {
            const unsigned m = cand_mask(ri, thr, pbase, CP) >> pbase;
            mymask = (CP == 32) ? m : (m & ((1u << (CP & 31)) - 1u));
          }
        }
        CTCX_TICK(17)  // PB: masks
        int pos0 = 0;
        if (warp_has_rows && __any_sync(kFull, mymask != 0u)) {  // (warp-uniform) somebody has something to list
          const int cnt = __popc(mymask);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
          }
          // list positions: one shared-memory atomic per warp (the order of the list does not matter)
          int base = 0;
          if (lane == 31 && incl) base = atomicAdd(&sci[kV2NCand], incl);
          base = __shfl_sync(kFull, base, 31);
          pos0 = base + incl - cnt;
        }
        CTCX_TICK(18)  // PB: scan
        // PB pass 2: list + histogram
        {
          unsigned m = mymask;
          int pos = pos0;
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1u;
            const int l = pbase + k;
            float cbase = (l == r_label) ? r_ob : r_ot;
            if constexpr (LM) cbase = __fadd_rn(cbase, s_lm[(r_label + 1) * 32 + l]);
            const unsigned key = KeyOf(__fadd_rn(s_pl[l], cbase));  // :172-182
            c_list[pos++] = make_uint2(key, ((unsigned)prow << 16) | (unsigned)l);
            atomicAdd(&s_hist[bucket_of(key)], 1u);
          }
          if (tid < n && member_in) atomicAdd(&s_hist[bucket_of(my_key)], 1u);
        }
        CTCX_TICK(19)  // PB: list
        __syncthreads();
        n_cand = sci[kV2NCand];
        CTCX_TICK(1)  // PB

        // ---- PD: boundary bin of the W-th item and group offsets: one bin per thread ----
        // Thread d (0..255) owns bin 255 - d, so an inclusive PREFIX scan in thread order is a SUFFIX
        // scan in bin order: warp scan, per-warp totals through shared memory, one barrier.
        static_assert(kBinsV2 == 256 && NT >= 256, "one bin per thread");
        unsigned pd_h = 0u, pd_incl = 0u;
        const int pd_bin = kBinsV2 - 1 - tid;
        if (tid < kBinsV2) {
          pd_h = s_hist[pd_bin];
          pd_incl = pd_h;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, pd_incl, o);
            if (lane >= o) pd_incl += v;
          }
          const unsigned nz = __ballot_sync(kFull, pd_h != 0u);
          if (lane == 31) {
            s_wtot[warp] = pd_incl;
            s_wtot[8 + warp] = nz ? (unsigned)(kBinsV2 - 1 - (warp * 32 + (__ffs(nz) - 1))) : 0xffffffffu;  // highest non-empty bin of the warp
          }
        }
        __syncthreads();
        if (tid < kBinsV2) {
          const uint4 wa = *reinterpret_cast<const uint4*>(s_wtot), wb = *reinterpret_cast<const uint4*>(s_wtot + 4);
          const unsigned wt[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
          unsigned before = 0u, total = 0u;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            before += (q < warp) ? wt[q] : 0u;
            total += wt[q];
          }
          const unsigned above = before + pd_incl - pd_h;  // items in bins above this thread's
          s_offs[pd_bin] = above;
          s_hist[pd_bin] = 0u;  // counters in PE / next attempt
          // with a clamped range the cut is valid only if the W-th item lies inside the range
          const bool usable = !clamped || (int)total >= W;
          const int K = min(W, (int)total);
          if (usable && (int)(above + pd_h) >= K && (int)above < K) {  // the boundary bin
            const uint4 ta = *reinterpret_cast<const uint4*>(s_wtot + 8), tb = *reinterpret_cast<const uint4*>(s_wtot + 12);
            const unsigned tops[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
            int topbin = -1;
#pragma unroll
            for (int q = 7; q >= 0; --q)
              if (tops[q] != 0xffffffffu) topbin = (int)tops[q];  // warp 0 owns the highest bins
            sci[kV2Bstar] = pd_bin;
            sci[kV2KRem] = K - (int)above;
            sci[kV2E] = (int)pd_h;
            sci[kV2NNew] = K;
            sci[kV2TopBin] = topbin;
            sci[kV3Found] = 1;
          }
        }
        __syncthreads();
        CTCX_TICK(3)  // PD
        if (__builtin_expect(sc[kV3Found] != 0, 1)) break;  // otherwise the prediction missed: run again over the full range
        if (tid == 0) sci[kV2NCand] = 0;  // the list is rebuilt
        __syncthreads();
      }
      const bool member_in = !clamped || my_key > lo_key;
      const int bstar = sc[kV2Bstar], k_rem = sc[kV2KRem], e_b = sc[kV2E], n_new = sc[kV2NNew];
      if (TIMING && timing) {  // event counters (slots 20-23): how selective would a "below the cut" filter be?
        int below = 0, wiped = 0;
        for (int q = 0; q < n_risk; ++q) {
          const int m = s_risk[q];
          below += (bucket_of(m_key[m]) <= bstar || (clamped && m_key[m] <= lo_key)) ? 1 : 0;
          wiped += s_wiped[m] ? 1 : 0;
        }
        cyc[TIMING ? 20 : 0] += n_risk;
        cyc[TIMING ? 21 : 0] += below;
        cyc[TIMING ? 22 : 0] += wiped;
        cyc[TIMING ? 23 : 0] += (n_risk > 0) ? 1 : 0;
      }
      // a suspect member (see PA) that is not certain to stay in the beam: report the utterance
      if (__builtin_expect(suspect, 0) && (!member_in || bucket_of(my_key) <= bstar)) sci[kV2Anomaly] = 1;
      const bool bnd_all = (e_b == k_rem);
      // next frame's range prediction: the measured top-to-threshold gap
      const unsigned gap_next = (unsigned)(sc[kV2TopBin] - bstar + 1) << shift;

      // ---- PE: scatter every item at or above the boundary bin into its score group ----
      auto place = [&](unsigned key, unsigned okey) {
        const int bucket = bucket_of(key);
        const unsigned long long comp = ((unsigned long long)key << 32) | (unsigned long long)(~okey);
        if (bucket > bstar || (bucket == bstar && bnd_all)) {
          const unsigned pos = s_offs[bucket] + atomicAdd(&s_hist[bucket], 1u);
          if (pos < (unsigned)WMAX) s_sorted[pos] = comp;
        } else if (bucket == bstar && e_b <= kBndFast) {
          const int pos = atomicAdd(&sci[kV2NBnd], 1);
          if (pos < kBndFast) s_bnd[pos] = comp;
        }
      };
      if (tid < n && member_in) place(my_key, (unsigned)tid);
      for (int c0 = tid; c0 < n_cand; c0 += 4 * NT) {  // four independent entries in flight
        uint2 e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * NT;
          e[u] = (c < n_cand) ? c_list[c] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (e[u].x) place(e[u].x, 0x80000000u | e[u].y);
      }
      __syncthreads();
      CTCX_TICK(4)  // PE

      // ---- PF: cut the boundary bin exactly ----
      // Usual case (at most 32 items in the boundary bin): the LAST warp ranks them with shuffles and
      // writes the k_rem best straight into their final positions, while the other warps already rank
      // the groups above the boundary (PG) -- no barrier between the two phases.
      const bool pf_fast = !bnd_all && e_b <= kBndFast;
      const int bnd_start = pf_fast ? (int)s_offs[bstar] : 0x7fffffff;  // slots from here on are set by PF
      if (!bnd_all) {
        if (__builtin_expect(e_b <= kBndFast, 1)) {
          if (s_warp) {
            const unsigned long long mine = (lane < e_b) ? s_bnd[lane] : 0ull;
            const unsigned mlo = (unsigned)mine, mhi = (unsigned)(mine >> 32);
            int rank = 0;
            for (int j = 0; j < e_b; ++j) {
              const unsigned olo = __shfl_sync(kFull, mlo, j), ohi = __shfl_sync(kFull, mhi, j);
              const unsigned long long other = ((unsigned long long)ohi << 32) | olo;
              rank += (other > mine) ? 1 : 0;
            }
            if (lane < e_b && rank < k_rem) s_fin[s_offs[bstar] + rank] = mine;
          }
        } else {
          // many items in the boundary bin (coarse bins after a missed prediction, or pathological
          // ties such as constant logits): radix select of the k_rem largest (key, ~order) composites
          auto for_each_bnd = [&](auto&& f) {
            if (tid < n && member_in && bucket_of(my_key) == bstar)
              f(((unsigned long long)my_key << 32) | (unsigned long long)(~(unsigned)tid), my_key, (unsigned)tid);
            for (int c = tid; c < n_cand; c += NT) {
              const uint2 e = c_list[c];
              if (e.x && bucket_of(e.x) == bstar)
                f(((unsigned long long)e.x << 32) | (unsigned long long)(~(0x80000000u | e.y)), e.x,
                  0x80000000u | e.y);
            }
          };
          const int npass = 8;
          if (tid == 0) { scu[kV2Prefix] = 0u; scu[kV2PrefixHi] = 0u; sci[kV2K] = k_rem; }
          __syncthreads();
          for (int pass = npass - 1; pass >= 0; --pass) {
            const int sh = pass * 8;
            unsigned* bins = s_bins2;
            for (int i = tid; i < 256; i += NT) bins[i] = 0u;
            __syncthreads();
            const unsigned long long prefix =
                ((unsigned long long)scu[kV2PrefixHi] << 32) | (unsigned long long)scu[kV2Prefix];
            for_each_bnd([&](unsigned long long v, unsigned, unsigned) {
              const unsigned long long hi = (sh + 8 >= 64) ? 0ull : (v >> (sh + 8));
              if (hi == prefix) atomicAdd(&bins[(unsigned)(v >> sh) & 255u], 1u);
            });
            __syncthreads();
            if (warp == 0) {
              const int k = sci[kV2K];
              unsigned h[8];
              unsigned loc = 0;
#pragma unroll
              for (int q = 0; q < 8; ++q) { h[q] = bins[lane * 8 + q]; loc += h[q]; }
              unsigned suf = loc;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_down_sync(kFull, suf, o);
                if (lane + o < 32) suf += v;
              }
              unsigned acc = suf - loc;
              if ((int)suf >= k && (int)acc < k) {
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                  if ((int)(acc + h[q]) >= k && (int)acc < k) {
                    const unsigned long long np = (prefix << 8) | (unsigned long long)(lane * 8 + q);
                    scu[kV2Prefix] = (unsigned)np;
                    scu[kV2PrefixHi] = (unsigned)(np >> 32);
                    sci[kV2K] = k - (int)acc;
                  }
                  acc += h[q];
                }
              }
            }
            __syncthreads();
          }
          const unsigned long long cut =
              ((unsigned long long)scu[kV2PrefixHi] << 32) | (unsigned long long)scu[kV2Prefix];
          for_each_bnd([&](unsigned long long v, unsigned key, unsigned okey) {
            if (v >= cut) {
              const unsigned pos = s_offs[bstar] + atomicAdd(&s_hist[bstar], 1u);
              if (pos < (unsigned)WMAX)
                s_sorted[pos] = ((unsigned long long)key << 32) | (unsigned long long)(~okey);
            }
          });
          __syncthreads();
        }
      }

      CTCX_TICK(5)  // PF
      // ---- PG: rank inside the score group = new slot; write the next beam + back-pointers ----
      {
        float* w_total = s_total + nxt * WMAX;
        float* w_blk = s_blk + nxt * WMAX;
        float* w_lab = s_lab + nxt * WMAX;
        float* w_ab = s_ab + nxt * WMAX;
        float* w_an = s_an + nxt * WMAX;
        int* w_label = s_label + nxt * WMAX;
        unsigned long long* w_hash = s_hash + nxt * WMAX;
        unsigned long long* w_phash = s_phash + nxt * WMAX;
        // clear the parent look-up table (this frame's look-ups happened in PA) before re-filling it
        for (int i = tid; i < TS / 4; i += NT)  // 16-byte stores
          reinterpret_cast<uint4*>(s_htab)[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        // survivors above the boundary (and the radix-selected boundary items): scatter to the final slot
        const int n_grouped = min(n_new, bnd_start);
        if (tid < n_grouped) {
          const unsigned long long comp = s_sorted[tid];
          const int bucket = bucket_of((unsigned)(comp >> 32));
          const int g0 = (int)s_offs[bucket];
          const int g1 = (bucket == bstar && !bnd_all) ? n_new : g0 + (int)s_hist[bucket];
          // Groups usually hold 1-3 items; quantised logits (bfloat16 inputs) make exact ties, i.e. groups of
          // dozens: four independent loads per round keep the count off the shared-memory latency.
          int rank = 0, j = g0;
#pragma unroll 1
          for (; j + 4 <= g1; j += 4) {
            const unsigned long long a = s_sorted[j], b2 = s_sorted[j + 1], c2 = s_sorted[j + 2], d2 = s_sorted[j + 3];
            rank += ((a > comp) ? 1 : 0) + ((b2 > comp) ? 1 : 0) + ((c2 > comp) ? 1 : 0) + ((d2 > comp) ? 1 : 0);
          }
#pragma unroll 1
          for (; j < g1; ++j) rank += (s_sorted[j] > comp) ? 1 : 0;
          s_fin[g0 + rank] = comp;
        }
        CTCX_TICK(13)  // PG: rank in group
        __syncthreads();  // table cleared, final order known; s_hist / scalars no longer needed this frame
        CTCX_TICK(14)  // PG: barrier
        for (int i = tid; i < kBinsV2; i += NT) s_hist[i] = 0u;
        if (tid == 0) {
          sci[kV2NCand] = 0;
          sci[kV2NRisk] = 0;
              scu[kV2MinKey] = 0xffffffffu;
          scu[kV2MaxKey] = 0u;
          sci[kV2NBnd] = 0;
          scu[kV2MinBase] = 0xffffffffu;
          scu[kV2Gap] = gap_next;
          sci[kV3Found] = 0;
        }
        if (tid < n) s_wiped[tid] = 0u;
        if (pg_role < PGS && pg_slot < n_new) {
          const int r = pg_slot;
          const unsigned long long comp = s_fin[r];
          const unsigned okey = ~(unsigned)(comp & 0xffffffffull);
          const bool fresh = (okey & 0x80000000u) != 0u;
          const int src = fresh ? (int)((okey & 0x7fffffffu) >> 16) : (int)okey;  // parent row / old slot
          const int lbl = fresh ? (int)(okey & 0xffffu) : o_label[src];
          const float s = UnKey((unsigned)(comp >> 32));
          if (pg_role == 0) {  // state arrays + back-pointer record
            unsigned rec;
            if (!fresh) {  // surviving member
              w_total[r] = m_nt[src];
              w_blk[r] = m_nb[src];
              w_lab[r] = m_nl[src];
              w_ab[r] = m_nab[src];
              w_an[r] = m_nan[src];
              rec = m_rec[src];
            } else {  // fresh child (decoder.h:170-187)
              const float pl = __fsub_rn(x[lbl], off);
              float v_an = __fadd_rn(o_ab[src], pl);
              unsigned an_kind = kAnParAb;
              if (lbl != o_label[src]) {
                const float c2 = __fadd_rn(o_an[src], pl);
                if (c2 > v_an) { v_an = c2; an_kind = kAnParAn; }
              }
              w_total[r] = s;
              w_blk[r] = NegInf();
              w_lab[r] = s;
              w_ab[r] = NegInf();
              w_an[r] = v_an;
              rec = PackRec32(0xffu, (unsigned)src, kAbFromAb, an_kind, (unsigned)lbl);
            }
            w_label[r] = lbl;
            p.bp32[((size_t)b * p.Tcap + (t_done + t0 + t)) * W + r] = rec;
            if (p.dbg_totals) p.dbg_totals[((size_t)b * T + t0 + t) * W + r] = w_total[r];
          }
          if (PGS == 1 || pg_role == 1) {  // prefix hash, row info of the next frame, parent look-up table
            unsigned long long hsh;
            float nt_, nb_;
            if (!fresh) {
              hsh = o_hash[src];
              w_phash[r] = o_phash[src];
              nt_ = m_nt[src];
              nb_ = m_nb[src];
            } else {
              hsh = HashChild(o_hash[src], lbl);
              w_phash[r] = o_hash[src];
              nt_ = s;
              nb_ = NegInf();
            }
            w_hash[r] = hsh;
            s_row[r] = make_uint4(__float_as_uint(nt_), __float_as_uint(nb_), (unsigned)lbl, 0u);
            unsigned h = (unsigned)hsh & (TS - 1);
            const unsigned entry = ((unsigned)(hsh >> 42) << 10) | (unsigned)r;
            while (atomicCAS(&s_htab[h], 0xffffffffu, entry) != 0xffffffffu) h = (h + 1) & (TS - 1);
          }
        }
        if (p.dbg_n && tid == 0) p.dbg_n[(size_t)b * T + t0 + t] = n_new;
        CTCX_TICK(15)  // PG: state write
      }
      __syncthreads();
      CTCX_TICK(6)  // PG
      n = n_new;
      if (__builtin_expect(p.ready != nullptr && sc[kV4Abort] != 0, 0)) {  // the input copy never arrived
        t_end = t + 1;
        break;
      }
    }
    if (TIMING && timing)
      for (int i = 0; i < (TIMING ? 24 : 1); ++i) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg_cycles + (size_t)b * 24 + i), (unsigned long long)cyc[i]);
        cyc[TIMING ? i : 0] = 0;
      }

    {
      const int cur = t_end & 1;
      const int overflow = (p.seq_len[b] > p.Tcap - t_done) ? 4 : 0;
      const int lost = ((t_end != L) ? 8 : 0) | (carried & 8);  // an input frame never arrived (this slice or before)
      // ---- final beam (decoder.h:229-261): sorted, the first P slots are the top paths ----
      if (last_slice || lost) {
        for (int q = tid; q < p.P; q += NT) {
          if (q < n) {
            p.fin_total[(size_t)b * p.P + q] = s_total[cur * WMAX + q];
            p.fin_kind[(size_t)b * p.P + q] = (s_ab[cur * WMAX + q] > s_an[cur * WMAX + q]) ? 1 : 0;
          } else {
            p.fin_total[(size_t)b * p.P + q] = 0.0f;
            p.fin_kind[(size_t)b * p.P + q] = 0;
          }
        }
        if (tid == 0) {
          p.fin_n[b] = n;
          p.flags[b] = (sci[kV2Anomaly] ? 1 : 0) | ((p.P > n) ? 2 : 0) | overflow | lost;
        }
      }
      // ---- carry the beam (and the score-range prediction) to the next slice / the next call ----
      if (p.state != nullptr && (!last_slice || p.t_done != nullptr)) {
        StreamView sv(p.state + (size_t)b * StreamStateBytes(W), W);
        for (int i = tid; i < n; i += NT) {
          sv.total[i] = s_total[cur * WMAX + i]; sv.blk[i] = s_blk[cur * WMAX + i];
          sv.lab[i] = s_lab[cur * WMAX + i]; sv.ab[i] = s_ab[cur * WMAX + i];
          sv.an[i] = s_an[cur * WMAX + i]; sv.label[i] = s_label[cur * WMAX + i];
          sv.hash[i] = s_hash[cur * WMAX + i]; sv.phash[i] = s_phash[cur * WMAX + i];
        }
        if (tid == 0) {
          sv.hdr->n = n;
          sv.hdr->gap = scu[kV2Gap];
          sv.hdr->flags = (sci[kV2Anomaly] ? 1 : 0) | overflow | lost | (carried & 4);
        }
      }
      if (p.t_done != nullptr && tid == 0) p.t_done[b] = t_done + t_end;
      if (p.n_slices > 1 && !last_slice) {  // release the next slice of this utterance
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p.progress + b), "r"(slice + 1) : "memory");
      }
    }
    __syncthreads();  // the next task re-initialises the shared state
  }
#undef CTCX_TICK
}

}  // namespace ctcx
