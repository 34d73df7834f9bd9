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

constexpr int kBinsV2 = 256;
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
  size_t htab;                     // i32 [2*WMAX]
  size_t hist, offs, bins2;        // u32 [256] each
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
    list = o; o += (size_t)cand_cap * 8;
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
    htab = o; o += 2 * w * 4;
    hist = o; o += kBinsV2 * 4;
    offs = o; o += kBinsV2 * 4;
    bins2 = o; o += kBinsV2 * 4;
    x = o; o += 2 * 32 * 4;
    scal = o; o += 32 * 4;
    bytes = (o + 15) / 16 * 16;
  }
};

enum {
  kV2NCand = 0, kV2NRisk, kV2MinKey, kV2MaxKey, kV2Changed, kV2NBnd, kV2Bstar, kV2KRem, kV2E,
  kV2NNew, kV2Off0, kV2Off1, kV2Anomaly, kV2MinBase, kV2LpMin, kV2Prefix, kV2PrefixHi, kV2K
};

template <int WMAX, int NT>
__global__ void __launch_bounds__(NT) BeamKernelV2(BeamParams p) {
  static_assert(NT >= WMAX && NT >= kBinsV2, "one thread per beam slot and per histogram bin");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NWARP = NT / 32;
  constexpr int TS = 2 * WMAX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int W = p.W, C = p.C, T = p.T, B = p.B, blank = p.blank_index;
  const int L = p.seq_len[b];

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
  int* s_htab = (int*)(smem + lay.htab);
  unsigned* s_hist = (unsigned*)(smem + lay.hist);
  unsigned* s_offs = (unsigned*)(smem + lay.offs);
  unsigned* s_bins2 = (unsigned*)(smem + lay.bins2);
  float* s_x = (float*)(smem + lay.x);
  volatile int* sc = (volatile int*)(smem + lay.scal);
  int* sci = (int*)(smem + lay.scal);
  unsigned* scu = (unsigned*)(smem + lay.scal);

  // ---- initial state: the root (decoder.h:212-227) ----
  LoadExpTable(s_exptab, tid, NT);
  for (int i = tid; i < TS; i += NT) s_htab[i] = -1;
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
  }
  int n = 1;
  if (L > 0) {
    const float* g = p.logits + (size_t)b * C;
    if (tid < C) s_x[tid] = g[tid];
    if (tid == 0) ((float*)sci)[kV2Off0] = p.off[b];
  }
  __syncthreads();
  if (tid == 0) {
    s_htab[(unsigned)kRootHash & (TS - 1)] = 0;
    s_row[0] = make_uint4(__float_as_uint(0.0f), __float_as_uint(0.0f), 0xffffffffu, 0u);
  }
  __syncthreads();

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

    // prefetch the next frame's row; consumed after the barrier that ends this frame
    if (t + 1 < L) {
      if (tid < C) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(s_x + nxt * 32 + tid);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa),
                     "l"(p.logits + ((size_t)(t + 1) * B + b) * C + tid));
      } else if (tid == 32) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared((float*)sci + kV2Off0 + nxt);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa),
                     "l"(p.off + (size_t)(t + 1) * B + b));
      }
      asm volatile("cp.async.commit_group;\n" ::);
    }

    // per-lane class constants (lane = class index)
    const bool lane_ok = (lane < C) && (lane != blank);
    const float pl_lane = lane_ok ? __fsub_rn(x[lane], off) : 0.0f;
    const float xb = x[blank];
    const float pb = __fsub_rn(xb, off);

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
        for (;;) {  // parent->Active() <=> the parent prefix is in the beam (decoder.h:97)
          const int s = s_htab[h];
          if (s < 0) break;
          if (o_hash[s] == ph) { pslot = s; break; }
          h = (h + 1) & (TS - 1);
        }
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
      const float v_nb = __fsub_rn(__fadd_rn(o_total[i], xb), off);
      const float c1 = __fadd_rn(o_ab[i], pb), c2 = __fadd_rn(o_an[i], pb);
      const unsigned ab_kind = (c2 > c1) ? kAbFromAn : kAbFromAb;
      const float v_nt = LogSumExp(v_nb, v_nl, s_exptab);
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
    {
      unsigned kmin = (tid < n) ? my_key : 0xffffffffu, kmax = (tid < n) ? my_key : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(kFull, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(kFull, kmax, o));
      }
      if (lane == 0 && warp * 32 < n) {
        atomicMin(&scu[kV2MinKey], kmin);
        atomicMax(&scu[kV2MaxKey], kmax);
      }
    }
    if (n < W) {  // beam not full: every finite child is admissible; bound the score range
      unsigned kb = 0xffffffffu;
      if (tid < n) {
        const float ob = o_blk[tid], ot = o_total[tid];
        if (ot > NegInf()) kb = KeyOf((ob > NegInf()) ? fminf(ot, ob) : ot);
      }
      float lpm = lane_ok ? pl_lane : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        kb = min(kb, __shfl_xor_sync(kFull, kb, o));
        lpm = fminf(lpm, __shfl_xor_sync(kFull, lpm, o));
      }
      if (lane == 0) {
        if (warp * 32 < n) atomicMin(&scu[kV2MinBase], kb);
        if (warp == 0) ((float*)sci)[kV2LpMin] = lpm;
      }
    }
    __syncthreads();

    // ---- PB: fresh children (decoder.h:146-187) + score histogram ----
    const unsigned minkey_m = scu[kV2MinKey];
    const float th0f = (n == W) ? UnKey(minkey_m) : NegInf();
    unsigned lo_key;
    if (n == W) {
      lo_key = minkey_m;
    } else {
      const unsigned kb = scu[kV2MinBase];
      unsigned lo_c = minkey_m;
      if (kb != 0xffffffffu) lo_c = KeyOf(__fadd_rn(UnKey(kb), ((const float*)sci)[kV2LpMin]));
      lo_key = min(minkey_m, lo_c);
      lo_key = max(lo_key, kKeyNegInf);
    }
    const unsigned hi_key = max(scu[kV2MaxKey], KeyOf(o_total[0]));
    const unsigned span = hi_key - lo_key;
    const int shift = max(0, (32 - __clz(span | 1u)) - 8);  // (key - lo) >> shift < 256
    if (tid < n) atomicAdd(&s_hist[(my_key - lo_key) >> shift], 1u);
    for (int row = warp; row < n; row += NWARP) {
      const uint4 ri = s_row[row];
      const float base = ((int)ri.z == lane) ? __uint_as_float(ri.y) : __uint_as_float(ri.x);
      const float s = __fadd_rn(pl_lane, base);  // decoder.h:172-182: (x - off) + old blank/total
      const bool ok = lane_ok && !((ri.w >> lane) & 1u) && (s > th0f);
      const unsigned m = __ballot_sync(kFull, ok);
      if (m) {
        int basepos = 0;
        if (lane == 0) basepos = atomicAdd(&sci[kV2NCand], __popc(m));
        basepos = __shfl_sync(kFull, basepos, 0);
        if (ok) {
          const unsigned key = KeyOf(s);
          c_list[basepos + __popc(m & ((1u << lane) - 1u))] =
              make_uint2(key, ((unsigned)row << 16) | (unsigned)lane);
          atomicAdd(&s_hist[(key - lo_key) >> shift], 1u);
        }
      }
    }
    __syncthreads();
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
            for (int c = lane; c < n_cand; c += 32) {
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
            atomicSub(&s_hist[(e.x - lo_key) >> shift], 1u);
            c_list[c].x = 0u;
          }
        }
      }
      __syncthreads();
    }

    // ---- PD: boundary bin of the W-th item and group offsets (warp 0) ----
    if (warp == 0) {
      unsigned h[8];
      unsigned loc = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { h[q] = s_hist[lane * 8 + q]; loc += h[q]; }
      unsigned suf = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(kFull, suf, o);
        if (lane + o < 32) suf += v;
      }
      const unsigned above = suf - loc;
      const int total = (int)__shfl_sync(kFull, suf, 0);
      const int K = min(W, total);
      unsigned acc = above;
#pragma unroll
      for (int q = 7; q >= 0; --q) {
        s_offs[lane * 8 + q] = acc;
        if ((int)(acc + h[q]) >= K && (int)acc < K) {
          sci[kV2Bstar] = lane * 8 + q;
          sci[kV2KRem] = K - (int)acc;
          sci[kV2E] = (int)h[q];
          sci[kV2NNew] = K;
        }
        acc += h[q];
        s_hist[lane * 8 + q] = 0u;  // re-used as per-group position counters in PE
      }
    }
    __syncthreads();
    const int bstar = sc[kV2Bstar], k_rem = sc[kV2KRem], e_b = sc[kV2E], n_new = sc[kV2NNew];
    const bool bnd_all = (e_b == k_rem);

    // ---- PE: scatter every item at or above the boundary bin into its score group ----
    auto place = [&](unsigned key, unsigned okey) {
      const int bucket = (int)((key - lo_key) >> shift);
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
    for (int c = tid; c < n_cand; c += NT) {
      const uint2 e = c_list[c];
      if (e.x) place(e.x, 0x80000000u | e.y);
    }
    __syncthreads();

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
        // pathological ties (e.g. constant logits): radix select of the k_rem largest
        // (key, ~order) composites among the items of the boundary bin
        const unsigned lo_b = lo_key + ((unsigned)bstar << shift);
        auto for_each_bnd = [&](auto&& f) {
          if (tid < n && (int)((my_key - lo_key) >> shift) == bstar)
            f(((unsigned long long)(my_key - lo_b) << 32) | (unsigned long long)(~(unsigned)tid), my_key,
              (unsigned)tid);
          for (int c = tid; c < n_cand; c += NT) {
            const uint2 e = c_list[c];
            if (e.x && (int)((e.x - lo_key) >> shift) == bstar)
              f(((unsigned long long)(e.x - lo_b) << 32) | (unsigned long long)(~(0x80000000u | e.y)), e.x,
                0x80000000u | e.y);
          }
        };
        const int npass = (32 + shift + 7) / 8;
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
      for (int i = tid; i < TS; i += NT) s_htab[i] = -1;
      unsigned long long comp = 0ull;
      int r = -1;
      if (tid < n_new) {
        comp = s_sorted[tid];
        const int bucket = (int)(((unsigned)(comp >> 32) - lo_key) >> shift);
        const int g0 = (int)s_offs[bucket], g1 = g0 + (int)s_hist[bucket];
        int rank = 0;
        for (int j = g0; j < g1; ++j) rank += (s_sorted[j] > comp) ? 1 : 0;
        r = g0 + rank;
      }
      __syncthreads();  // table cleared, ranks known; s_hist / scalars no longer needed this frame
      if (tid < kBinsV2) s_hist[tid] = 0u;
      if (tid == 0) {
        sci[kV2NCand] = 0;
        sci[kV2NRisk] = 0;
        scu[kV2MinKey] = 0xffffffffu;
        scu[kV2MaxKey] = 0u;
        sci[kV2NBnd] = 0;
        scu[kV2MinBase] = 0xffffffffu;
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
        p.bp[((size_t)b * T + t) * W + r] = make_uint2(rec, (unsigned)lbl);
        if (p.dbg_totals) p.dbg_totals[((size_t)b * T + t) * W + r] = nt_;
        unsigned h = (unsigned)hsh & (TS - 1);
        while (atomicCAS(&s_htab[h], -1, r) != -1) h = (h + 1) & (TS - 1);
      }
      if (p.dbg_n && tid == 0) p.dbg_n[(size_t)b * T + t] = n_new;
    }
    asm volatile("cp.async.wait_all;\n" ::);
    __syncthreads();
    n = n_new;
  }

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
