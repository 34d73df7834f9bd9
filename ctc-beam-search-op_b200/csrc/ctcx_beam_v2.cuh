// Beam kernel, second generation (the fast path for num_classes <= 32 with the candidate list in
// shared memory). Same parallel formulation and the same bit-exact results as BeamKernel in
// ctcx_kernels.cuh (which remains the generic path: wide vocabularies / streaming mode), but
// organised around ONE pass over the candidates instead of several:
//
//   PA  member update (one thread per member)                         decoder.h:95-143
//   PB  fresh children: lane = class, warp = row; each accepted child is appended to the list and
//       counted in a 256-bin histogram of its score key over [threshold, upper bound]
//                                                                     decoder.h:146-187
//   PC  revisit-wipe fixed point (only when some member's parent precedes it)   SURVEY A.4
//   PD  warp 0: suffix scan of the histogram -> boundary bin of the W-th item, group offsets
//   PE  every item above the boundary bin is scattered into its score group (descending bins)
//   PF  the boundary bin is cut exactly (warp-level rank; radix select for pathological ties)
//   PG  rank inside the (tiny) score group = final beam slot; write the next beam, its parent
//       look-up table and the back-pointer records                    decoder.h:189-199
//
// This replaces the reference's gtl::TopN push/evict sequence by an order-preserving selection:
// the beam is totally ordered by (score desc, members before children, visiting order).
#pragma once
#include "ctcx_kernels.cuh"

namespace ctcx {

constexpr int kBinsV2 = 256;   // score-histogram bins (256 measured best: 512 costs scan work, 128 crowds the boundary bin)
constexpr int kBinsLog2V2 = 8;
constexpr int kBndFast = 32;  // boundary items handled by one warp

struct BeamSmemV2 {
  size_t hash, phash;              // u64 [2][WMAX]
  size_t sorted;                   // u64 [WMAX]   score-grouped survivors
  size_t bnd;                      // u64 [32]     boundary-bin items (fast path)
  size_t exptab;                   // u64 [32]
  size_t row;                      // uint4 [WMAX] {old total, old blank, label, member-children mask}
  size_t list;                     // uint2 [cand_cap] {score key, (row<<16)|label}
  size_t total, blk, lab, ab, an;  // f32 [2][WMAX]
  size_t label;                    // i32 [2][WMAX]
  size_t m_nt, m_nb, m_nl, m_nab, m_nan;  // f32 [WMAX]
  size_t m_key, m_rec;             // u32 [WMAX]
  size_t m_pslot;                  // i32 [WMAX]
  size_t risk, risk_new;           // i32 [WMAX]
  size_t wiped;                    // u32 [WMAX]
  size_t htab;                     // u32 [4*WMAX]  (hash tag << 10 | slot), 0xffffffff = empty
  size_t hist, offs;               // u32 [kBinsV2] each
  size_t bins2;                    // u32 [256]
  size_t rowstart;                 // i32 [WMAX+1] list offset of each row's first candidate
  size_t pl;                       // f32 [32]     x[l] - off of the current frame
  size_t wsum;                     // i32 [32]     per-warp candidate counts (block scan)
  size_t x;                        // f32 [2][32]
  size_t scal;                     // 32 x 4 B
  size_t bytes;
  __host__ __device__ void Init(int wmax, int cand_cap) {
    size_t o = 0;
    const size_t w = (size_t)wmax;
    hash = o; o += 2 * w * 8;
    phash = o; o += 2 * w * 8;
    sorted = o; o += w * 8;
    bnd = o; o += kBndFast * 8;
    exptab = o; o += 32 * 8;
    row = o; o += w * 16;
    list = o; o += (((size_t)cand_cap + 1024) * 8 + 15) / 16 * 16;  // + one scratch slot per thread
    total = o; o += 2 * w * 4;
    blk = o; o += 2 * w * 4;
    lab = o; o += 2 * w * 4;
    ab = o; o += 2 * w * 4;
    an = o; o += 2 * w * 4;
    label = o; o += 2 * w * 4;
    m_nt = o; o += w * 4;
    m_nb = o; o += w * 4;
    m_nl = o; o += w * 4;
    m_nab = o; o += w * 4;
    m_nan = o; o += w * 4;
    m_key = o; o += w * 4;
    m_rec = o; o += w * 4;
    m_pslot = o; o += w * 4;
    risk = o; o += w * 4;
    risk_new = o; o += w * 4;
    wiped = o; o += w * 4;
    htab = o; o += 4 * w * 4;
    hist = o; o += kBinsV2 * 4;
    offs = o; o += kBinsV2 * 4;
    bins2 = o; o += 256 * 4;
    rowstart = o; o += (w + 1 + 3) / 4 * 4 * 4;
    pl = o; o += 32 * 4;
    wsum = o; o += 32 * 4;
    x = o; o += 2 * 32 * 4;
    scal = o; o += 32 * 4;
    bytes = (o + 15) / 16 * 16;
  }
};

enum {
  kV2NCand = 0, kV2NRisk, kV2MinKey, kV2MaxKey, kV2Changed, kV2NBnd, kV2Bstar, kV2KRem, kV2E,
  kV2NNew, kV2Off0, kV2Off1, kV2Anomaly, kV2MinBase, kV2LpMin, kV2Prefix, kV2PrefixHi, kV2K,
  kV2LpMax, kV2Gap, kV2TopBin
};

template <int WMAX, int NT, bool TIMING>
__global__ void __launch_bounds__(NT, ((NT <= 256 && !TIMING) ? 4 : 1)) BeamKernelV2(BeamParams p) {
  static_assert(NT >= WMAX && 2 * NT >= kBinsV2, "one thread per beam slot and per two histogram bins");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NWARP = NT / 32;
  constexpr int TS = 4 * WMAX;  // parent look-up table slots (load factor <= 1/4)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int W = p.W, C = p.C, T = p.T, B = p.B, blank = p.blank_index;
  const int L = max(0, min(p.seq_len[b], p.T));

  BeamSmemV2 lay;
  lay.Init(WMAX, p.cand_cap);
  unsigned long long* s_hash = (unsigned long long*)(smem + lay.hash);
  unsigned long long* s_phash = (unsigned long long*)(smem + lay.phash);
  unsigned long long* s_sorted = (unsigned long long*)(smem + lay.sorted);
  unsigned long long* s_bnd = (unsigned long long*)(smem + lay.bnd);
  unsigned long long* s_exptab = (unsigned long long*)(smem + lay.exptab);
  uint4* s_row = (uint4*)(smem + lay.row);
  uint2* c_list = (uint2*)(smem + lay.list);
  float* s_total = (float*)(smem + lay.total);
  float* s_blk = (float*)(smem + lay.blk);
  float* s_lab = (float*)(smem + lay.lab);
  float* s_ab = (float*)(smem + lay.ab);
  float* s_an = (float*)(smem + lay.an);
  int* s_label = (int*)(smem + lay.label);
  float* m_nt = (float*)(smem + lay.m_nt);
  float* m_nb = (float*)(smem + lay.m_nb);
  float* m_nl = (float*)(smem + lay.m_nl);
  float* m_nab = (float*)(smem + lay.m_nab);
  float* m_nan = (float*)(smem + lay.m_nan);
  unsigned* m_key = (unsigned*)(smem + lay.m_key);
  unsigned* m_rec = (unsigned*)(smem + lay.m_rec);
  int* m_pslot = (int*)(smem + lay.m_pslot);
  int* s_risk = (int*)(smem + lay.risk);
  int* s_risk_new = (int*)(smem + lay.risk_new);
  unsigned* s_wiped = (unsigned*)(smem + lay.wiped);
  unsigned* s_htab = (unsigned*)(smem + lay.htab);
  unsigned* s_hist = (unsigned*)(smem + lay.hist);
  unsigned* s_offs = (unsigned*)(smem + lay.offs);
  unsigned* s_bins2 = (unsigned*)(smem + lay.bins2);
  int* s_rowstart = (int*)(smem + lay.rowstart);
  float* s_pl = (float*)(smem + lay.pl);
  int* s_wsum = (int*)(smem + lay.wsum);
  float* s_x = (float*)(smem + lay.x);
  volatile int* sc = (volatile int*)(smem + lay.scal);
  int* sci = (int*)(smem + lay.scal);
  unsigned* scu = (unsigned*)(smem + lay.scal);

  // ---- initial state: the root (decoder.h:212-227) ----
  LoadExpTable(s_exptab, tid, NT);
  for (int i = tid; i < TS; i += NT) s_htab[i] = 0xffffffffu;
  for (int i = tid; i < WMAX; i += NT) {
    s_wiped[i] = 0u;
    s_row[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < kBinsV2; i += NT) s_hist[i] = 0u;
  if (tid == 0) {
    s_total[0] = 0.0f;
    s_blk[0] = 0.0f;
    s_lab[0] = NegInf();
    s_ab[0] = 0.0f;  // empty alignment with probability 1 (entry.h:204-209)
    s_an[0] = NegInf();
    s_label[0] = -1;
    s_hash[0] = kRootHash;
    s_phash[0] = 0ull;
    sci[kV2Anomaly] = 0;
    sci[kV2NCand] = 0;
    sci[kV2NRisk] = 0;
    scu[kV2MinKey] = 0xffffffffu;
    scu[kV2MaxKey] = 0u;
    sci[kV2NBnd] = 0;
    scu[kV2MinBase] = 0xffffffffu;
    scu[kV2Gap] = 0u;
  }
  int n = 1;
  // thread -> (row, class slice) mapping of the candidate pass
  constexpr int PARTS = NT / WMAX;  // threads per row
  constexpr int CP = 32 / PARTS;    // classes per thread
  static_assert(NT % WMAX == 0 && 32 % PARTS == 0, "row/class tiling");
  const int prow = tid / PARTS, pbase = (tid % PARTS) * CP;
  const unsigned class_mask = ((C >= 32) ? 0xffffffffu : ((1u << C) - 1u)) & ~(1u << blank);
  if (L > 0) {
    const float* g = p.logits + (size_t)b * C;
    if (tid < C) s_x[tid] = g[tid];
    if (tid == 0) ((float*)sci)[kV2Off0] = p.off[b];
  }
  __syncthreads();
  if (tid == 0) {
    s_htab[(unsigned)kRootHash & (TS - 1)] = ((unsigned)(kRootHash >> 42) << 10) | 0u;
    s_row[0] = make_uint4(__float_as_uint(0.0f), __float_as_uint(0.0f), 0xffffffffu, 0u);
  }
  __syncthreads();

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
  if (timing) tprev = clock64();
  for (int t = 0; t < L; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    const float* x = s_x + cur * 32;
    const float off = ((const float*)sci)[kV2Off0 + cur];
    const float* o_total = s_total + cur * WMAX;
    const float* o_blk = s_blk + cur * WMAX;
    const float* o_lab = s_lab + cur * WMAX;
    const float* o_ab = s_ab + cur * WMAX;
    const float* o_an = s_an + cur * WMAX;
    const int* o_label = s_label + cur * WMAX;
    const unsigned long long* o_hash = s_hash + cur * WMAX;
    const unsigned long long* o_phash = s_phash + cur * WMAX;

    // prefetch the next frame's row (last warp, idle during PA); consumed after the barrier that
    // ends this frame
    if (warp == NWARP - 1 && t + 1 < L) {
      if (lane < C) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(s_x + nxt * 32 + lane);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa),
                     "l"(p.logits + ((size_t)(t + 1) * B + b) * C + lane));
      }
      if (lane == 31) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared((float*)sci + kV2Off0 + nxt);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa),
                     "l"(p.off + (size_t)(t + 1) * B + b));
      }
      asm volatile("cp.async.commit_group;\n" ::);
    }

    // per-class log-probabilities of this frame (lane = class index), their max and min
    const bool lane_ok = (lane < C) && (lane != blank);
    const float pl_lane = lane_ok ? __fsub_rn(x[lane], off) : 0.0f;
    const float xb = x[blank];
    const float pb = __fsub_rn(xb, off);
    if (warp == NWARP - 1) {  // idle during PA
      s_pl[lane] = pl_lane;
      const unsigned kmx = __reduce_max_sync(kFull, lane_ok ? KeyOf(pl_lane) : 0u);
      const unsigned kmn = __reduce_min_sync(kFull, lane_ok ? KeyOf(pl_lane) : 0xffffffffu);
      if (lane == 0) {
        ((float*)sci)[kV2LpMax] = (kmx != 0u) ? UnKey(kmx) : NegInf();
        ((float*)sci)[kV2LpMin] = (kmn != 0xffffffffu) ? UnKey(kmn) : 0.0f;
      }
    }
    CTCX_TICK(7)  // frame setup
    // ---- PA: update the existing members (decoder.h:95-143) ----
    unsigned my_key = 0u;
    if (tid < n) {
      const int i = tid;
      const int lbl = o_label[i];
      int pslot = -1;
      float v_nl = o_lab[i], v_an = NegInf();
      unsigned an_kind = kAnNone, an_src = kInvalidSlot;
      if (lbl >= 0) {
        const unsigned long long ph = o_phash[i];
        unsigned h = (unsigned)ph & (TS - 1);
        const unsigned tag = (unsigned)(ph >> 42);  // 22 hash bits disjoint from the table index
        for (;;) {  // parent->Active() <=> the parent prefix is in the beam (decoder.h:97)
          const unsigned e = s_htab[h];
          if (e == 0xffffffffu) break;
          if ((e >> 10) == tag && o_hash[e & 1023u] == ph) { pslot = (int)(e & 1023u); break; }
          h = (h + 1) & (TS - 1);
        }
        CTCX_TICK(8)  // parent look-up
        const float xl = x[lbl];
        const float pl = __fsub_rn(xl, off);
        const float self_an = __fadd_rn(o_an[i], pl);
        if (pslot >= 0) {
          const bool same = (lbl == o_label[pslot]);
          const float base = same ? o_blk[pslot] : o_total[pslot];
          v_nl = __fsub_rn(__fadd_rn(LogSumExp(o_lab[i], base, s_exptab), xl), off);
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
      m_rec[i] = PackRec((unsigned)i, an_src, ab_kind, an_kind);
      m_pslot[i] = pslot;
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
    if (n < W) {  // beam not full: every finite child is admissible; bound the score range
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

    // ---- PB: fresh children (decoder.h:146-187) + score histogram ----
    const unsigned minkey_m = scu[kV2MinKey];
    const float th0f = (n == W) ? UnKey(minkey_m) : NegInf();
    const float lp_max = ((const float*)sci)[kV2LpMax];
    unsigned lo_true;  // no item lies below this key
    if (n == W) {
      lo_true = minkey_m;
    } else {
      const unsigned kb = scu[kV2MinBase];
      unsigned lo_c = minkey_m;
      if (kb != 0xffffffffu) lo_c = KeyOf(__fadd_rn(UnKey(kb), ((const float*)sci)[kV2LpMin]));
      lo_true = max(min(minkey_m, lo_c), kKeyNegInf);
    }
    // every item is <= max(best member, best possible child); old totals are sorted, slot 0 is the max
    const unsigned hi_key = max(scu[kV2MaxKey], KeyOf(__fadd_rn(lp_max, o_total[0])));
    // Histogram range. Survivors crowd near the top while the admissible range reaches far below,
    // so the 256 bins are centred on a PREDICTION of the W-th score: the previous frame's
    // top-to-threshold gap (x3 + slack). Keys below the range are clamped into bin 0. This only
    // affects speed: whatever bin the W-th item falls in is cut exactly in PF.
    unsigned lo_key = lo_true;
    {
      const unsigned gap = scu[kV2Gap];
      if (gap != 0u && n == W) {
        const unsigned long long reach = 3ull * gap + 64ull;
        if (reach < (unsigned long long)(hi_key - lo_true)) lo_key = hi_key - (unsigned)reach;
      }
    }
    const bool clamped = (lo_key != lo_true);
    const unsigned span = hi_key - lo_key;
    const int shift = max(0, (32 - __clz(span | 1u)) - kBinsLog2V2);  // (key - lo) >> shift < kBinsV2
    auto bucket_of = [&](unsigned key) -> int { return (key > lo_key) ? (int)((key - lo_key) >> shift) : 0; };
    CTCX_TICK(16)  // PB: range
    // pass 1: each thread scores its (row, class slice) and keeps a bitmask of admissible children
    unsigned okmask = 0u;
    float sv[CP];
    if (prow < n) {
      const uint4 ri = s_row[prow];
      const float r_ot = __uint_as_float(ri.x), r_ob = __uint_as_float(ri.y);
      const int r_label = (int)ri.z;
      const unsigned vm = ((~ri.w) & class_mask) >> pbase;  // not blank, not already a member
      if ((CP == 32 ? vm : (vm & ((1u << (CP & 31)) - 1u))) && __fadd_rn(lp_max, r_ot) > th0f) {
#pragma unroll
        for (int k = 0; k < CP; ++k) {
          const int l = pbase + k;
          sv[k] = __fadd_rn(s_pl[l], (l == r_label) ? r_ob : r_ot);  // :172-182
          if (((vm >> k) & 1u) && sv[k] > th0f) okmask |= 1u << k;
        }
      }
    }
    CTCX_TICK(17)  // PB: pass 1
    int pos0;
    {
      const int cnt = __popc(okmask);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w2 = 0; w2 < NWARP; ++w2) {
        const int v = s_wsum[w2];
        total += v;
        if (w2 < warp) before += v;
      }
      pos0 = before + incl - cnt;
      if (tid == 0) {
        sci[kV2NCand] = total;
        s_rowstart[n] = total;
      }
      if (prow < n && pbase == 0) s_rowstart[prow] = pos0;
    }
    CTCX_TICK(18)  // PB: scan
    // pass 2: write them in visiting order (row, then class). Branch-free: a rejected class stores
    // into this thread's scratch slot behind the list, so the 16 stores overlap instead of forming 16
    // divergent regions.
    {
      uint2* scratch = c_list + p.cand_cap + tid;
#pragma unroll
      for (int k = 0; k < CP; ++k) {
        const bool ok = (okmask >> k) & 1u;
        uint2* dst = ok ? (c_list + pos0 + __popc(okmask & ((1u << k) - 1u))) : scratch;
        *dst = make_uint2(KeyOf(sv[k]), ((unsigned)prow << 16) | (unsigned)(pbase + k));
      }
    }
    __syncthreads();
    CTCX_TICK(19)  // PB: pass 2
    // pass 3: histogram of members and listed children, balanced over the threads, four entries in
    // flight. Everything below the predicted range lands in bin 0 and is counted per warp (one atomic
    // instead of hundreds on the same address).
    {
      const int n_cand_ = sci[kV2NCand];
      int n_clamped = 0;
      if (tid < n) {
        const int bk = bucket_of(my_key);
        if (bk == 0) ++n_clamped; else atomicAdd(&s_hist[bk], 1u);
      }
      for (int c0 = tid; c0 < n_cand_; c0 += 4 * NT) {
        int bk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * NT;
          bk[u] = (c < n_cand_) ? bucket_of(c_list[c].x) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (bk[u] > 0) atomicAdd(&s_hist[bk[u]], 1u);
          n_clamped += (bk[u] == 0) ? 1 : 0;
        }
      }
      n_clamped = __reduce_add_sync(kFull, n_clamped);
      if (lane == 0 && n_clamped) atomicAdd(&s_hist[0], (unsigned)n_clamped);
    }
    __syncthreads();
    CTCX_TICK(1)  // PB
    const int n_cand = sci[kV2NCand];
    const int n_risk = sci[kV2NRisk];

    // ---- PC: revisit-wipe fixed point ----
    if (n_risk > 0) {
      for (;;) {
        if (tid == 0) sci[kV2Changed] = 0;
        for (int q = warp; q < n_risk; q += NWARP) {
          const int m = s_risk[q];
          const int pslot = m_pslot[m];
          int verdict = 0;
          if (!s_wiped[pslot]) {
            const unsigned vkey = m_key[m];
            const unsigned idm = ((unsigned)pslot << 16) | (unsigned)o_label[m];
            int cnt = 0;
            for (int j = lane; j < n; j += 32) {
              const unsigned kj = m_key[j];
              cnt += (kj > vkey || (kj == vkey && j < m)) ? 1 : 0;
            }
            const int c_end = s_rowstart[pslot + 1];  // children visited up to the parent's turn
            for (int c = lane; c < c_end; c += 32) {
              const uint2 e = c_list[c];
              cnt += (e.x > vkey && e.y < idm && !s_wiped[e.y >> 16]) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
            verdict = (cnt >= W) ? 1 : 0;
          }
          if (lane == 0) s_risk_new[q] = verdict;
        }
        __syncthreads();
        for (int q = tid; q < n_risk; q += NT) {
          const int m = s_risk[q];
          const unsigned v = (unsigned)s_risk_new[q];
          if (s_wiped[m] != v) {
            s_wiped[m] = v;
            sci[kV2Changed] = 1;
          }
        }
        __syncthreads();
        const int changed = sc[kV2Changed];
        __syncthreads();
        if (!changed) break;
      }
      for (int q = tid; q < n_risk; q += NT) {
        const int m = s_risk[q];
        if (s_wiped[m]) {
          const int pslot = m_pslot[m];
          const int lbl = o_label[m];
          const float base = (lbl == o_label[pslot]) ? o_blk[pslot] : o_total[pslot];
          if (KeyOf(__fadd_rn(__fsub_rn(x[lbl], off), base)) > m_key[m]) sci[kV2Anomaly] = 1;
          sci[kV2Changed] = 2;
        }
      }
      __syncthreads();
      if (sc[kV2Changed] == 2) {  // drop the children of wiped members from list and histogram
        for (int c = tid; c < n_cand; c += NT) {
          const uint2 e = c_list[c];
          if (s_wiped[e.y >> 16]) {
            atomicSub(&s_hist[bucket_of(e.x)], 1u);
            c_list[c].x = 0u;
          }
        }
      }
      __syncthreads();
    }

    CTCX_TICK(2)  // PC
    // ---- PD: boundary bin of the W-th item and group offsets (two bins per thread) ----
    {
      // suffix sums over bins kBinsV2-1..0: thread `tid` owns bins hi = kBinsV2-1-2*tid and lo = hi-1, so an
      // inclusive PREFIX scan in thread order is an inclusive SUFFIX scan in bin order
      const int bin_hi = kBinsV2 - 1 - 2 * tid;
      unsigned h_hi = 0u, h_lo = 0u;
      if (bin_hi >= 1) {
        const uint2 hh = *reinterpret_cast<const uint2*>(&s_hist[bin_hi - 1]);
        h_lo = hh.x;
        h_hi = hh.y;
      }
      const unsigned h2 = h_hi + h_lo;
      unsigned incl = h2;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned nz = __ballot_sync(kFull, h2 != 0u);
      if (lane == 31) s_wsum[warp] = (int)incl;
      {
        const int l0 = nz ? (__ffs(nz) - 1) : 0;  // first lane (highest bins) holding anything
        const unsigned hh = __shfl_sync(kFull, h_hi, l0);
        const int tb = nz ? ((kBinsV2 - 1 - 2 * (warp * 32 + l0)) - (hh ? 0 : 1)) : -1;
        if (lane == 0) s_wsum[16 + warp] = tb;
      }
      __syncthreads();
      unsigned before = 0u, total = 0u;
      int topbin = -1;
#pragma unroll
      for (int w2 = 0; w2 < NWARP; ++w2) {
        const unsigned v = (unsigned)s_wsum[w2];
        total += v;
        if (w2 < warp) before += v;
        topbin = max(topbin, s_wsum[16 + w2]);
      }
      if (bin_hi >= 1) {
        const int K = min(W, (int)total);
        const unsigned above_hi = before + incl - h2;  // items in bins above bin_hi
        const unsigned above_lo = above_hi + h_hi;
        *reinterpret_cast<uint2*>(&s_offs[bin_hi - 1]) = make_uint2(above_lo, above_hi);
        *reinterpret_cast<uint2*>(&s_hist[bin_hi - 1]) = make_uint2(0u, 0u);  // re-used as counters in PE
        if ((int)(above_hi + h_hi) >= K && (int)above_hi < K) {
          sci[kV2Bstar] = bin_hi;
          sci[kV2KRem] = K - (int)above_hi;
          sci[kV2E] = (int)h_hi;
          sci[kV2NNew] = K;
          sci[kV2TopBin] = topbin;
        } else if ((int)(above_lo + h_lo) >= K && (int)above_lo < K) {
          sci[kV2Bstar] = bin_hi - 1;
          sci[kV2KRem] = K - (int)above_lo;
          sci[kV2E] = (int)h_lo;
          sci[kV2NNew] = K;
          sci[kV2TopBin] = topbin;
        }
      }
    }
    __syncthreads();
    CTCX_TICK(3)  // PD
    const int bstar = sc[kV2Bstar], k_rem = sc[kV2KRem], e_b = sc[kV2E], n_new = sc[kV2NNew];
    const bool bnd_all = (e_b == k_rem);
    // next frame's range prediction: measured top-to-threshold gap; if the prediction missed (the
    // threshold fell into the clamped bin) widen to the whole range just used
    const unsigned gap_next = (clamped && bstar == 0) ? span
                                                      : ((unsigned)(sc[kV2TopBin] - bstar + 1) << shift);

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
    if (tid < n) place(my_key, (unsigned)tid);
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
    if (!bnd_all) {
      if (e_b <= kBndFast) {
        if (warp == 0) {
          const unsigned long long mine = (lane < e_b) ? s_bnd[lane] : 0ull;
          const unsigned mlo = (unsigned)mine, mhi = (unsigned)(mine >> 32);
          int rank = 0;
          for (int j = 0; j < e_b; ++j) {
            const unsigned olo = __shfl_sync(kFull, mlo, j), ohi = __shfl_sync(kFull, mhi, j);
            const unsigned long long other = ((unsigned long long)ohi << 32) | olo;
            rank += (other > mine) ? 1 : 0;
          }
          if (lane < e_b && rank < k_rem) s_sorted[s_offs[bstar] + rank] = mine;
          if (lane == 0) s_hist[bstar] = (unsigned)k_rem;
        }
      } else {
        // many items in the boundary bin (coarse bins after a missed prediction, or pathological
        // ties such as constant logits): radix select of the k_rem largest (key, ~order) composites
        auto for_each_bnd = [&](auto&& f) {
          if (tid < n && bucket_of(my_key) == bstar)
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
      }
      __syncthreads();
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
      for (int i = tid; i < TS; i += NT) s_htab[i] = 0xffffffffu;
      unsigned long long comp = 0ull;
      int r = -1;
      if (tid < n_new) {
        comp = s_sorted[tid];
        const int bucket = bucket_of((unsigned)(comp >> 32));
        const int g0 = (int)s_offs[bucket], g1 = g0 + (int)s_hist[bucket];
        int rank = 0;
        for (int j = g0; j < g1; ++j) rank += (s_sorted[j] > comp) ? 1 : 0;
        r = g0 + rank;
      }
      CTCX_TICK(13)  // PG: rank in group
      __syncthreads();  // table cleared, ranks known; s_hist / scalars no longer needed this frame
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
      }
      if (tid < n) s_wiped[tid] = 0u;
      if (r >= 0) {
        const unsigned okey = ~(unsigned)(comp & 0xffffffffull);
        unsigned rec;
        int lbl;
        unsigned long long hsh;
        float nt_, nb_;
        if (!(okey & 0x80000000u)) {  // surviving member
          const int i = (int)okey;
          nt_ = m_nt[i];
          nb_ = m_nb[i];
          w_lab[r] = m_nl[i];
          w_ab[r] = m_nab[i];
          w_an[r] = m_nan[i];
          lbl = o_label[i];
          hsh = o_hash[i];
          w_phash[r] = o_phash[i];
          rec = m_rec[i];
        } else {  // fresh child (decoder.h:170-187)
          const int row = (int)((okey & 0x7fffffffu) >> 16);
          lbl = (int)(okey & 0xffffu);
          const float s = UnKey((unsigned)(comp >> 32));
          const float pl = __fsub_rn(x[lbl], off);
          float v_an = __fadd_rn(o_ab[row], pl);
          unsigned an_kind = kAnParAb;
          if (lbl != o_label[row]) {
            const float c2 = __fadd_rn(o_an[row], pl);
            if (c2 > v_an) { v_an = c2; an_kind = kAnParAn; }
          }
          nt_ = s;
          nb_ = NegInf();
          w_lab[r] = s;
          w_ab[r] = NegInf();
          w_an[r] = v_an;
          hsh = HashChild(o_hash[row], lbl);
          w_phash[r] = o_hash[row];
          rec = PackRec(kInvalidSlot, (unsigned)row, kAbFromAb, an_kind);
        }
        w_total[r] = nt_;
        w_blk[r] = nb_;
        w_label[r] = lbl;
        w_hash[r] = hsh;
        s_row[r] = make_uint4(__float_as_uint(nt_), __float_as_uint(nb_), (unsigned)lbl, 0u);
        p.bp[((size_t)b * p.Tcap + t) * W + r] = make_uint2(rec, (unsigned)lbl);
        if (p.dbg_totals) p.dbg_totals[((size_t)b * T + t) * W + r] = nt_;
        unsigned h = (unsigned)hsh & (TS - 1);
        const unsigned entry = ((unsigned)(hsh >> 42) << 10) | (unsigned)r;
        while (atomicCAS(&s_htab[h], 0xffffffffu, entry) != 0xffffffffu) h = (h + 1) & (TS - 1);
      }
      if (p.dbg_n && tid == 0) p.dbg_n[(size_t)b * T + t] = n_new;
      CTCX_TICK(15)  // PG: state write
    }
    asm volatile("cp.async.wait_all;\n" ::);
    __syncthreads();
    CTCX_TICK(6)  // PG
    n = n_new;
  }
  if (TIMING && timing)
    for (int i = 0; i < (TIMING ? 24 : 1); ++i) p.dbg_cycles[(size_t)b * 24 + i] = cyc[i];
#undef CTCX_TICK

  // ---- final beam (decoder.h:229-261): sorted, the first P slots are the top paths ----
  {
    const int cur = L & 1;
    if (tid < p.P) {
      if (tid < n) {
        p.fin_total[(size_t)b * p.P + tid] = s_total[cur * WMAX + tid];
        p.fin_kind[(size_t)b * p.P + tid] = (s_ab[cur * WMAX + tid] > s_an[cur * WMAX + tid]) ? 1 : 0;
      } else {
        p.fin_total[(size_t)b * p.P + tid] = 0.0f;
        p.fin_kind[(size_t)b * p.P + tid] = 0;
      }
    }
    if (tid == 0) {
      p.fin_n[b] = n;
      p.flags[b] = (sci[kV2Anomaly] ? 1 : 0) | ((p.P > n) ? 2 : 0);
    }
  }
}

}  // namespace ctcx
