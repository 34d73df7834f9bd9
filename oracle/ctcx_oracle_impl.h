/* TEST INFRASTRUCTURE ONLY (oracle/). Included twice by ctcx_oracle.c with REAL = float / double.
 *
 * Plain-C restatement of the reference's CTC "extended" beam search (sequential semantics).
 * Citations are relative to /root/reference/tensorflow_ctc_ext_beam_search_decoder/cc/:
 *   decoder.h = util/ctc_ext_beam_search_decoder.h, entry.h = util/ctc_beam_entry.h,
 *   loss_util.h = util/ctc_loss_util.h, kernels.cc = kernels/ctc_ext_beam_search_decoder_kernels.cc.
 *
 * Differences in representation (not in behaviour):
 *   * trie nodes live in one growable array; children are found through an open-addressing table
 *     keyed by (parent id, label) instead of a per-node FlatMap (entry.h:114-122);
 *   * alignment label sequences are persistent linked lists (symbol, previous cell) instead of
 *     std::vector copies (entry.h:215-221); a candidate only materialises a cell when its owner
 *     survives into the next frame. Only priority_queue::top() is ever read by the reference
 *     (entry.h:93-100), so each queue is kept as "first-pushed maximum";
 *   * the beam container (gtl::TopN, un-vendored TensorFlow) is an unordered array with exact set
 *     semantics; among exactly equal keys the choice differs from libstdc++'s heap order, which is
 *     why the decision margins (ctcx_oracle.c: CTCX_MARGIN_*) are recorded.
 */

typedef struct {
  REAL prob;
  int has;  /* queue non-empty */
  int cell; /* linked-list cell of the label sequence (-1 = empty sequence) */
} SUF(OldCand);

typedef struct {
  REAL prob;
  int has;
  int prev_cell; /* sequence it extends (-1 = starts a new sequence, entry.h:204-213) */
  int symbol;
  REAL second;   /* best prob among the other pushes (alignment margin only) */
} SUF(NewCand);

typedef struct {
  int parent; /* node id; -1 for the root (decoder.h:218) */
  int label;  /* -1 for the root */
  REAL o_total, o_blank, o_label; /* oldp (entry.h:238) */
  REAL n_total, n_blank, n_label; /* newp */
  SUF(OldCand) o_ab, o_an;        /* tops of old_cands.cands_blank / cands_nblank */
  SUF(NewCand) n_ab, n_an;        /* tops of new_cands */
  /* instrumentation only */
  int stamp;      /* frame id in which the node was in `branches` */
  int turn;       /* its index in `branches` in that frame */
  int shadow_valid;
  REAL shadow_total, shadow_blank;
  REAL evicted_total; /* n_total at the moment of eviction (instrumentation) */
} SUF(Node);

typedef struct {
  SUF(Node)* nodes;
  int n_nodes, cap_nodes;
  int* tab_node; /* child table: open addressing over (parent,label) -> node id */
  long long* tab_key;
  int tab_cap, tab_used;
  int* cell_sym; /* alignment cells */
  int* cell_prev;
  int n_cells, cap_cells;
  int* leaves;   /* the beam (leaves_, decoder.h:59) */
  int n_leaves;
  int* branches;
  int W, C, blank_index, blank_label;
  int frame;
  double margin[CTCX_N_MARGINS];
  ctcx_oracle_stats* stats;
  /* scorer extension point (util/ctc_beam_scorer.h:31-65): NULL = the default scorer
   * (GetStateExpansionScore returns previous_score unchanged); otherwise a [C+1, C] table of
   * expansion scores, row = label of the expanded entry + 1 (row 0: the root), column = new label:
   * ExpandState caches table[from_label + 1][to_label] in the child (decoder.h:171) and
   * GetStateExpansionScore returns previous_score + that value (decoder.h:103,114,176,182). */
  const REAL* lm;
} SUF(Dec);

static REAL SUF(expansion)(const SUF(Dec)* d, int from_label, int to_label, REAL previous_score) {
  if (!d->lm) return previous_score;
  return previous_score + d->lm[(size_t)(from_label + 1) * (size_t)d->C + (size_t)to_label];
}

#define NEG_INF ((REAL)(-INFINITY))

static void SUF(reset_prob_new)(SUF(Node)* n) { n->n_total = n->n_blank = n->n_label = NEG_INF; }
static void SUF(reset_prob_old)(SUF(Node)* n) { n->o_total = n->o_blank = n->o_label = NEG_INF; }
static void SUF(reset_cands_new)(SUF(Node)* n) {
  n->n_ab.has = n->n_an.has = 0;
  n->n_ab.prob = n->n_an.prob = NEG_INF;
  n->n_ab.second = n->n_an.second = NEG_INF;
}
static void SUF(reset_cands_old)(SUF(Node)* n) {
  n->o_ab.has = n->o_an.has = 0;
  n->o_ab.prob = n->o_an.prob = NEG_INF;
  n->o_ab.cell = n->o_an.cell = -1;
}

static void SUF(note)(SUF(Dec)* d, int which, double m) {
  if (m < 0) m = -m;
  if (m < d->margin[which]) d->margin[which] = m;
}

static int SUF(add_node)(SUF(Dec)* d, int parent, int label) {
  if (d->n_nodes == d->cap_nodes) {
    d->cap_nodes = d->cap_nodes ? d->cap_nodes * 2 : 1024;
    d->nodes = (SUF(Node)*)realloc(d->nodes, sizeof(SUF(Node)) * (size_t)d->cap_nodes);
  }
  SUF(Node)* n = &d->nodes[d->n_nodes];
  n->parent = parent;
  n->label = label;
  SUF(reset_prob_old)(n);
  SUF(reset_prob_new)(n);
  SUF(reset_cands_old)(n);
  SUF(reset_cands_new)(n);
  n->stamp = -1;
  n->turn = 0;
  n->shadow_valid = 0;
  return d->n_nodes++;
}

static unsigned SUF(tab_hash)(long long key, int mask) {
  unsigned long long h = (unsigned long long)key * 0x9E3779B97F4A7C15ull;
  return (unsigned)(h >> 36) & (unsigned)mask;
}

static void SUF(tab_grow)(SUF(Dec)* d) {
  int old_cap = d->tab_cap;
  int* old_node = d->tab_node;
  long long* old_key = d->tab_key;
  d->tab_cap = old_cap ? old_cap * 2 : 4096;
  d->tab_node = (int*)malloc(sizeof(int) * (size_t)d->tab_cap);
  d->tab_key = (long long*)malloc(sizeof(long long) * (size_t)d->tab_cap);
  for (int i = 0; i < d->tab_cap; ++i) d->tab_node[i] = -1;
  int mask = d->tab_cap - 1;
  for (int j = 0; j < old_cap; ++j) {
    if (old_node[j] < 0) continue;
    int i = (int)SUF(tab_hash)(old_key[j], mask);
    while (d->tab_node[i] >= 0) i = (i + 1) & mask;
    d->tab_node[i] = old_node[j];
    d->tab_key[i] = old_key[j];
  }
  free(old_node);
  free(old_key);
}

/* entry.h:114-122 GetChild (create != 0) or a pure look-up (create == 0, instrumentation). */
static int SUF(get_child)(SUF(Dec)* d, int parent, int label, int create) {
  if ((d->tab_used + 1) * 2 > d->tab_cap) SUF(tab_grow)(d);
  long long key = (long long)(((unsigned long long)(unsigned)parent << 32) | (unsigned)label);
  int mask = d->tab_cap - 1;
  int i = (int)SUF(tab_hash)(key, mask);
  while (d->tab_node[i] >= 0) {
    if (d->tab_key[i] == key) return d->tab_node[i];
    i = (i + 1) & mask;
  }
  if (!create) return -1;
  int id = SUF(add_node)(d, parent, label);
  d->tab_node[i] = id;
  d->tab_key[i] = key;
  d->tab_used++;
  return id;
}

static int SUF(new_cell)(SUF(Dec)* d, int prev, int sym) {
  if (d->n_cells == d->cap_cells) {
    d->cap_cells = d->cap_cells ? d->cap_cells * 2 : 4096;
    d->cell_sym = (int*)realloc(d->cell_sym, sizeof(int) * (size_t)d->cap_cells);
    d->cell_prev = (int*)realloc(d->cell_prev, sizeof(int) * (size_t)d->cap_cells);
  }
  d->cell_sym[d->n_cells] = sym;
  d->cell_prev[d->n_cells] = prev;
  return d->n_cells++;
}

/* loss_util.h:33-41. NOTE: the float functions log1pf/expf even when REAL is double (:39-40). */
static REAL SUF(lse)(REAL a, REAL b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  return (a > b) ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

/* entry.h:190-228 AddAlignmentCandidate, called on node `self` with source node `src`. */
static void SUF(add_cand)(SUF(Dec)* d, int self, int src, int from_blank, int to_blank, int sym,
                          REAL p) {
  SUF(Node)* me = &d->nodes[self];
  const SUF(Node)* s = &d->nodes[src];
  REAL old_prob;
  int prev_cell;
  if (from_blank && s->o_ab.has) { /* entry.h:194-198 */
    old_prob = s->o_ab.prob;
    prev_cell = s->o_ab.cell;
  } else if (!from_blank && s->o_an.has) { /* entry.h:199-203 */
    old_prob = s->o_an.prob;
    prev_cell = s->o_an.cell;
  } else {
    /* entry.h:204-213: start a new sequence; probability 0 only from-blank, and only when `self`
     * is the root or a New() child of the root. */
    int zero_ok = (me->parent < 0) || (d->nodes[me->parent].parent < 0 && me->o_total == NEG_INF);
    old_prob = zero_ok ? (from_blank ? (REAL)0 : NEG_INF) : NEG_INF;
    prev_cell = -1;
  }
  REAL np = old_prob + p; /* entry.h:218 */
  SUF(NewCand)* q = to_blank ? &me->n_ab : &me->n_an;
  /* priority_queue with (a.prob < b.prob) (entry.h:69-72): the earlier push stays on top among
   * equal maxima, so replace only on strictly greater. */
  if (!q->has) {
    q->has = 1;
    q->prob = np;
    q->prev_cell = prev_cell;
    q->symbol = sym;
    q->second = NEG_INF;
  } else if (np > q->prob) {
    q->second = q->prob;
    q->prob = np;
    q->prev_cell = prev_cell;
    q->symbol = sym;
  } else if (np > q->second) {
    q->second = np;
  }
}

static void SUF(dec_reset)(SUF(Dec)* d) {
  /* decoder.h:212-227 */
  d->n_nodes = 0;
  d->n_cells = 0;
  d->tab_used = 0;
  for (int i = 0; i < d->tab_cap; ++i) d->tab_node[i] = -1;
  int root = SUF(add_node)(d, -1, -1);
  d->nodes[root].n_total = (REAL)0; /* ln 1 */
  d->nodes[root].n_blank = (REAL)0;
  d->leaves[0] = root;
  d->n_leaves = 1;
  d->frame = 0;
}

/* index within leaves of the beam bottom; *second = the runner-up's value.
 * TIE POLICY (this repo's, documented in DESIGN.md): `leaves` is kept in arrival order (members in
 * `branches` order, then accepted children in visiting order). Among exactly equal totals the LAST
 * arrival is the bottom, and sorting is stable, i.e. the beam is totally ordered by
 * (total descending, arrival ascending). The reference's order among equal keys is whatever
 * libstdc++'s heap leaves (unspecified); the CTCX_MARGIN_* records tell where that can matter. */
static int SUF(bottom)(const SUF(Dec)* d, REAL* second) {
  int bi = 0;
  REAL sec = (REAL)INFINITY;
  for (int i = 1; i < d->n_leaves; ++i) {
    REAL v = d->nodes[d->leaves[i]].n_total;
    REAL bv = d->nodes[d->leaves[bi]].n_total;
    if (v <= bv) {
      sec = bv;
      bi = i;
    } else if (v < sec) {
      sec = v;
    }
  }
  if (second) *second = sec;
  return bi;
}

/* decoder.h:151-155 is_candidate */
static int SUF(is_candidate)(SUF(Dec)* d, REAL total, int record_margin) {
  if (!(total > NEG_INF)) return 0;
  if (d->n_leaves < d->W) return 1;
  REAL th = d->nodes[d->leaves[SUF(bottom)(d, NULL)]].n_total;
  if (record_margin) SUF(note)(d, CTCX_MARGIN_ACCEPT, (double)total - (double)th);
  return total > th;
}

/* stable descending insertion sort of node ids by n_total (TopN::Extract, decoder.h:84 / :252). */
static void SUF(sort_desc)(SUF(Dec)* d, int* ids, int n) {
  for (int i = 1; i < n; ++i) {
    int id = ids[i];
    REAL v = d->nodes[id].n_total;
    int j = i - 1;
    while (j >= 0 && d->nodes[ids[j]].n_total < v) {
      ids[j + 1] = ids[j];
      --j;
    }
    ids[j + 1] = id;
  }
}

/* One frame: decoder.h:67-210 */
static void SUF(step)(SUF(Dec)* d, const REAL* x) {
  const int C = d->C, blank = d->blank_index;
  ctcx_oracle_stats* st = d->stats;
  /* decoder.h:71-80: softmax normaliser, sum in index order */
  REAL mx = x[0];
  for (int j = 1; j < C; ++j) mx = (x[j] > mx) ? x[j] : mx;
  REAL sum = (REAL)0;
  for (int j = 0; j < C; ++j) sum += REAL_EXP(x[j] - mx);
  REAL off = mx + REAL_LOG(sum);

  /* decoder.h:84-85: branches = beam sorted by decreasing total; beam emptied */
  const int n = d->n_leaves;
  int* br = d->branches;
  for (int i = 0; i < n; ++i) br[i] = d->leaves[i];
  SUF(sort_desc)(d, br, n);
  for (int i = 1; i < n; ++i)
    SUF(note)(d, CTCX_MARGIN_ORDER,
              (double)d->nodes[br[i - 1]].n_total - (double)d->nodes[br[i]].n_total);
  d->n_leaves = 0;

  /* decoder.h:87-92: new -> old (the only place a sequence cell is materialised) */
  for (int i = 0; i < n; ++i) {
    SUF(Node)* b = &d->nodes[br[i]];
    b->o_total = b->n_total;
    b->o_blank = b->n_blank;
    b->o_label = b->n_label;
    b->o_ab.has = b->n_ab.has;
    b->o_ab.prob = b->n_ab.prob;
    b->o_ab.cell = b->n_ab.has ? SUF(new_cell)(d, b->n_ab.prev_cell, b->n_ab.symbol) : -1;
    b->o_an.has = b->n_an.has;
    b->o_an.prob = b->n_an.prob;
    b->o_an.cell = b->n_an.has ? SUF(new_cell)(d, b->n_an.prev_cell, b->n_an.symbol) : -1;
    SUF(reset_cands_new)(b);
    b->stamp = d->frame;
    b->turn = i;
    b->shadow_valid = 0;
  }

  /* decoder.h:95-143: update the existing members */
  for (int i = 0; i < n; ++i) {
    const int bi = br[i];
    SUF(Node)* b = &d->nodes[bi];
    if (b->parent >= 0) {
      const SUF(Node)* par = &d->nodes[b->parent];
      const REAL p = x[b->label] - off;
      if (par->n_total != NEG_INF) { /* parent->Active(), decoder.h:97 */
        if (b->label == par->label) { /* :98-108 */
          b->n_label = SUF(lse)(b->n_label, SUF(expansion)(d, par->label, b->label, par->o_blank)) + x[b->label] - off;
          SUF(add_cand)(d, bi, b->parent, 1, 0, b->label, p);
          SUF(add_cand)(d, bi, bi, 0, 0, b->label, p);
        } else { /* :109-121 */
          b->n_label = SUF(lse)(b->n_label, SUF(expansion)(d, par->label, b->label, par->o_total)) + x[b->label] - off;
          SUF(add_cand)(d, bi, b->parent, 1, 0, b->label, p);
          SUF(add_cand)(d, bi, b->parent, 0, 0, b->label, p);
          SUF(add_cand)(d, bi, bi, 0, 0, b->label, p);
        }
      } else { /* :123-128 */
        b->n_label += x[b->label] - off;
        SUF(add_cand)(d, bi, bi, 0, 0, b->label, p);
      }
    }
    /* :129-142 */
    b->n_blank = b->o_total + x[blank] - off;
    SUF(add_cand)(d, bi, bi, 1, 1, d->blank_label, x[blank] - off);
    SUF(add_cand)(d, bi, bi, 0, 1, d->blank_label, x[blank] - off);
    b->n_total = SUF(lse)(b->n_blank, b->n_label);
    if (b->n_an.has) SUF(note)(d, CTCX_MARGIN_ALIGN, (double)b->n_an.prob - (double)b->n_an.second);
    SUF(note)(d, CTCX_MARGIN_ALIGN, (double)b->n_ab.prob - (double)b->n_ab.second);
    d->leaves[d->n_leaves++] = bi;
  }

  /* instrumentation: children above the weakest a-priori threshold (design statistics only) */
  if (st) {
    REAL th0 = NEG_INF;
    if (n == d->W) th0 = d->nodes[d->leaves[SUF(bottom)(d, NULL)]].n_total;
    long long rel = 0;
    for (int i = 0; i < n; ++i) {
      const SUF(Node)* b = &d->nodes[br[i]];
      for (int l = 0; l < C; ++l) {
        if (l == blank) continue;
        int c = SUF(get_child)(d, br[i], l, 0);
        if (c >= 0 && d->nodes[c].stamp == d->frame) continue; /* a member */
        REAL s = x[l] - off + SUF(expansion)(d, b->label, l, (l == b->label) ? b->o_blank : b->o_total);
        if (s > th0) ++rel;
      }
    }
    st->relevant_children += rel;
    if (rel > st->max_relevant_children) st->max_relevant_children = rel;
    st->frames += 1;
  }

  /* decoder.h:146-209: grow new leaves */
  int frame_effective = 0;
  for (int i = 0; i < n; ++i) {
    const int bi = br[i];
    if (st && d->nodes[bi].shadow_valid) {
      /* instrumentation: this member was wiped before its turn. Would it have had an
       * acceptable child right now? */
      const SUF(Node)* b = &d->nodes[bi];
      int eff = 0;
      for (int l = 0; l < C && !eff; ++l) {
        if (l == blank) continue;
        int c = SUF(get_child)(d, bi, l, 0);
        if (c >= 0 && d->nodes[c].n_total != NEG_INF) continue;
        REAL s = x[l] - off + SUF(expansion)(d, b->label, l, (l == b->label) ? b->shadow_blank : b->shadow_total);
        if (SUF(is_candidate)(d, s, 0)) eff = 1;
      }
      if (eff) {
        st->effective_wipes += 1;
        frame_effective = 1;
      }
    }
    if (!SUF(is_candidate)(d, d->nodes[bi].o_total, 0)) continue; /* gate, :157 */
    if (st) st->turns_passed_gate += 1;
    for (int ind = 0; ind < C; ++ind) {
      if (ind == blank) continue; /* :163 */
      const int ci = SUF(get_child)(d, bi, ind, 1); /* may move d->nodes */
      SUF(Node)* c = &d->nodes[ci];
      const SUF(Node)* b = &d->nodes[bi];
      if (c->n_total != NEG_INF) continue; /* c.Active(), :168 */
      const REAL p = x[ind] - off;
      c->n_blank = NEG_INF; /* :170 */
      if (c->label == b->label) { /* :172-177 */
        c->n_label = p + SUF(expansion)(d, b->label, ind, b->o_blank);
        SUF(add_cand)(d, ci, bi, 1, 0, ind, p);
      } else { /* :178-185 */
        c->n_label = p + SUF(expansion)(d, b->label, ind, b->o_total);
        SUF(add_cand)(d, ci, bi, 1, 0, ind, p);
        SUF(add_cand)(d, ci, bi, 0, 0, ind, p);
      }
      c->n_total = c->n_label; /* :187 */
      if (st) st->child_evals += 1;
      const int revisit = (c->stamp == d->frame); /* was a member at the start of this frame */
      if (st && revisit && c->n_total > c->evicted_total) st->revisit_above_former += 1;
      if (SUF(is_candidate)(d, c->n_total, 1)) { /* :189 */
        if (d->n_leaves == d->W) {               /* :192-198 evict the bottom */
          REAL second;
          int k = SUF(bottom)(d, &second);
          SUF(Node)* bot = &d->nodes[d->leaves[k]];
          SUF(note)(d, CTCX_MARGIN_BOTTOM, (double)second - (double)bot->n_total);
          bot->evicted_total = bot->n_total;
          SUF(reset_prob_new)(bot);
          SUF(reset_cands_new)(bot);
          for (int q = k + 1; q < d->n_leaves; ++q) d->leaves[q - 1] = d->leaves[q];
          --d->n_leaves;
        }
        d->leaves[d->n_leaves++] = ci; /* :199 */
        SUF(note)(d, CTCX_MARGIN_ALIGN, (double)c->n_an.prob - (double)c->n_an.second);
        if (st) {
          st->accepted += 1;
          if (revisit) st->revisit_accepts += 1;
        }
      } else { /* :200-206 deactivate ("wipe") the child */
        if (st && revisit) {
          st->wipes += 1;
          if (c->turn > i) {
            st->wipes_before_turn += 1;
            c->shadow_valid = 1;
            c->shadow_total = c->o_total;
            c->shadow_blank = c->o_blank;
          }
        }
        SUF(reset_prob_old)(c);
        SUF(reset_prob_new)(c);
        SUF(reset_cands_old)(c);
        SUF(reset_cands_new)(c);
      }
    }
  }
  if (st) {
    st->frames_with_effective_wipe += frame_effective;
    for (int i = 0; i < d->n_leaves; ++i)
      if (d->nodes[d->leaves[i]].stamp != d->frame) st->surviving_children += 1;
  }
  d->frame += 1;
}

/* decoder.h:229-261 TopPaths + entry.h:123-152 LabelSeq / AlignmentLabelSeq.
 * Returns 0, or 6 / 7 for the two InvalidArgument cases. */
static int SUF(top_paths)(SUF(Dec)* d, int P, int merge_repeated, int max_time, int* dec_len,
                          int* dec, int* ali_len, int* ali, REAL* logp) {
  if (P > d->W) return 6;        /* :237-239 */
  if (P > d->n_leaves) return 7; /* :240-243 */
  int* ids = d->branches;
  for (int i = 0; i < d->n_leaves; ++i) ids[i] = d->leaves[i];
  SUF(sort_desc)(d, ids, d->n_leaves);
  for (int i = 1; i < d->n_leaves && i <= P; ++i)
    SUF(note)(d, CTCX_MARGIN_FINAL,
              (double)d->nodes[ids[i - 1]].n_total - (double)d->nodes[ids[i]].n_total);
  for (int p = 0; p < P; ++p) {
    const SUF(Node)* e = &d->nodes[ids[p]];
    /* entry.h:137-152 */
    const SUF(NewCand)* pick = NULL;
    if (e->n_ab.has && e->n_an.has) {
      SUF(note)(d, CTCX_MARGIN_ALIGN, (double)e->n_ab.prob - (double)e->n_an.prob);
      pick = (e->n_ab.prob > e->n_an.prob) ? &e->n_ab : &e->n_an;
    } else if (e->n_ab.has) {
      pick = &e->n_ab;
    } else if (e->n_an.has) {
      pick = &e->n_an;
    }
    int n = 0;
    if (pick) {
      for (int c = pick->prev_cell; c >= 0; c = d->cell_prev[c]) ++n;
      n += 1;
      int* out = ali + (size_t)p * max_time;
      if (n <= max_time) {
        int k = n - 1;
        out[k--] = pick->symbol;
        for (int c = pick->prev_cell; c >= 0; c = d->cell_prev[c]) out[k--] = d->cell_sym[c];
      }
    }
    ali_len[p] = n;
    /* entry.h:123-136 */
    int m = 0, prev_label = -1;
    int* o = dec + (size_t)p * max_time;
    for (int c = ids[p]; d->nodes[c].parent >= 0; c = d->nodes[c].parent) {
      if (!merge_repeated || d->nodes[c].label != prev_label) {
        if (m < max_time) o[m] = d->nodes[c].label;
        ++m;
      }
      prev_label = d->nodes[c].label;
    }
    for (int a = 0, z = m - 1; a < z; ++a, --z) {
      int tmp = o[a];
      o[a] = o[z];
      o[z] = tmp;
    }
    dec_len[p] = m;
    logp[p] = e->n_total; /* :258 */
  }
  return 0;
}

/* kernels.cc:20-95 Compute (validation :97-160 for the value-dependent checks). */
static int SUF(decode)(const REAL* logits, int T, int B, int C, const int* seq_len, int W, int P,
                       int merge_repeated, int blank_index, int blank_label, int* dec_len, int* dec,
                       int* ali_len, int* ali, REAL* logp, double* margins,
                       ctcx_oracle_stats* stats, char* err, int errcap, const REAL* lm) {
  if (T == 0) { /* kernels.cc:118-120 */
    snprintf(err, (size_t)errcap, "max_time is 0");
    return 2;
  }
  for (int b = 0; b < B; ++b) { /* kernels.cc:134-138 */
    if (!(seq_len[b] <= T)) {
      snprintf(err, (size_t)errcap, "sequence_length(%d) <= %d", b, T);
      return 5;
    }
  }
  SUF(Dec) d;
  memset(&d, 0, sizeof(d));
  d.W = W;
  d.C = C;
  d.blank_index = blank_index;
  d.blank_label = blank_label;
  d.stats = stats;
  d.lm = lm;
  d.leaves = (int*)malloc(sizeof(int) * (size_t)(W + 1));
  d.branches = (int*)malloc(sizeof(int) * (size_t)(W + 1));
  SUF(tab_grow)(&d);
  int rc = 0;
  for (int b = 0; b < B && rc == 0; ++b) {
    SUF(dec_reset)(&d); /* kernels.cc:55-57 / :85 */
    for (int k = 0; k < CTCX_N_MARGINS; ++k) d.margin[k] = INFINITY;
    for (int t = 0; t < seq_len[b]; ++t) /* kernels.cc:74-79 */
      SUF(step)(&d, logits + ((size_t)t * B + b) * C);
    const size_t row = (size_t)b * P;
    rc = SUF(top_paths)(&d, P, merge_repeated, T, dec_len + row, dec + row * T, ali_len + row,
                        ali + row * T, logp + row);
    if (rc == 6) snprintf(err, (size_t)errcap, "requested more paths than the beam width.");
    if (rc == 7) snprintf(err, (size_t)errcap, "Less leaves in the beam search than requested.");
    if (margins)
      for (int k = 0; k < CTCX_N_MARGINS; ++k) margins[(size_t)b * CTCX_N_MARGINS + k] = d.margin[k];
  }
  free(d.nodes);
  free(d.tab_node);
  free(d.tab_key);
  free(d.cell_sym);
  free(d.cell_prev);
  free(d.leaves);
  free(d.branches);
  return rc;
}

#undef NEG_INF
