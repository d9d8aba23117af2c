/* CPU oracle for CUTTING / PRUNING / SPLITTING + SCC labelling, plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * The product (libmpn_b200.so) never links or loads this file; tests/ load it through ctypes as the checker for graphs that
 * are too large for the Python restatement (oracle/postproc_oracle.py), which it follows function by function and against
 * which — and against the golden outputs of the unmodified reference in tests/golden/post_*.npz — the CPU suite pins it
 * (tests/test_c_oracle.py).  Integer / index work: the bar is bit-exact decisions and label integers.
 *
 * Reference lines restated (all under /root/reference):
 *   compute_SCC_and_Clusters            utils.py:30-52    networkx SCC emission order, sorted(key=len), isolated nodes last
 *   remove_edges_single_direction       utils.py:125-142  CUTTING
 *   pruning                             utils.py:144-339  (live lines 161-188, 277-317) picks taken from one snapshot per round
 *   splitting                           utils.py:54-123   global float equality on the minimum probability (utils.py:96-98)
 *   post_processing                     inference.py:70-169   CUT -> PRUNE -> CUT -> SPLIT -> labels
 *   networkx.strongly_connected_components (networkx 2.5.1, env_gnn.yml:76; not vendored): non-recursive Tarjan with
 *   Nuutila's modifications; node order = first appearance in the edge list, successor order = insertion order.
 *
 * PRUNE and SPLIT are written as the "parallel rounds" of SURVEY.md appendix B (one snapshot per round), the formulation the
 * Python oracle validates against the statement-by-statement mirror and against the reference's outputs.
 *
 * Build (also done by __graft_entry__.build()):  gcc -O2 -shared -fPIC oracle/postproc_oracle.c -o oracle/_build/libpostproc_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

#define PO_OK 0
#define PO_ERR_ALLOC 1
#define PO_ERR_ARG 2

/* ------------------------------------------------------------------------------------------------------------------
 * Strongly connected components of the active-edge digraph, in networkx's emission order.
 * comp[v] = emission index of v's component, -1 for nodes without any active edge (they are not in the DiGraph:
 * utils.py:31 builds it from the active edge list).  Returns the number of components or -1 (allocation failure).
 * sizes_out (malloc'ed, caller frees): size of each emitted component.
 * ---------------------------------------------------------------------------------------------------------------- */
static i64 scc_emission_order(const i64* src, const i64* dst, const uint8_t* act, i64 E, i64 N, i64* comp, i64** sizes_out) {
  i64 A = 0;
  for (i64 e = 0; e < E; ++e) A += act[e] != 0;
  i64* order = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));      /* nodes in first-appearance order */
  i64* beg = (i64*)calloc((size_t)N + 2, sizeof(i64));                    /* adjacency offsets (active out-edges, edge order) */
  i64* adj = (i64*)malloc(sizeof(i64) * (size_t)(A > 0 ? A : 1));
  i64* it = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));          /* resumable successor iterator of every node */
  i64* pre = (i64*)calloc((size_t)(N > 0 ? N : 1), sizeof(i64));          /* preorder number, 0 = not visited */
  i64* low = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  uint8_t* found = (uint8_t*)calloc((size_t)(N > 0 ? N : 1), 1);
  uint8_t* seen = (uint8_t*)calloc((size_t)(N > 0 ? N : 1), 1);
  i64* queue = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));       /* DFS stack */
  i64* sccq = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));        /* Tarjan's component stack */
  i64* sizes = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  if (!order || !beg || !adj || !it || !pre || !low || !found || !seen || !queue || !sccq || !sizes) {
    free(order); free(beg); free(adj); free(it); free(pre); free(low); free(found); free(seen); free(queue); free(sccq); free(sizes);
    return -1;
  }
  i64 n_order = 0;
  for (i64 e = 0; e < E; ++e) {
    if (!act[e]) continue;
    const i64 u = src[e], v = dst[e];
    if (!seen[u]) { seen[u] = 1; order[n_order++] = u; }
    if (!seen[v]) { seen[v] = 1; order[n_order++] = v; }
    beg[u + 2]++;
  }
  for (i64 v = 0; v < N; ++v) beg[v + 2] += beg[v + 1];                   /* beg[v+1] = start of v, filled below to its end */
  for (i64 e = 0; e < E; ++e)
    if (act[e]) adj[beg[src[e] + 1]++] = dst[e];                          /* stable: successors in edge (= insertion) order */
  /* now beg[v] .. beg[v+1] is v's successor list */
  for (i64 v = 0; v < N; ++v) { it[v] = beg[v]; comp[v] = -1; }
  i64 counter = 0, n_comp = 0, nq = 0, ns = 0;
  for (i64 oi = 0; oi < n_order; ++oi) {
    const i64 source = order[oi];
    if (found[source]) continue;
    nq = 0;
    queue[nq++] = source;
    while (nq > 0) {
      const i64 v = queue[nq - 1];
      if (pre[v] == 0) pre[v] = ++counter;
      int done = 1;
      while (it[v] < beg[v + 1]) {
        const i64 w = adj[it[v]++];
        if (pre[w] == 0) { queue[nq++] = w; done = 0; break; }
      }
      if (!done) continue;
      low[v] = pre[v];
      for (i64 k = beg[v]; k < beg[v + 1]; ++k) {
        const i64 w = adj[k];
        if (found[w]) continue;
        if (pre[w] > pre[v]) { if (low[w] < low[v]) low[v] = low[w]; }
        else if (pre[w] < low[v]) low[v] = pre[w];
      }
      --nq;
      if (low[v] == pre[v]) {
        i64 size = 1;
        comp[v] = n_comp; found[v] = 1;
        while (ns > 0 && pre[sccq[ns - 1]] > pre[v]) {
          const i64 k = sccq[--ns];
          comp[k] = n_comp; found[k] = 1; ++size;
        }
        sizes[n_comp++] = size;
      } else {
        sccq[ns++] = v;
      }
    }
  }
  free(order); free(beg); free(adj); free(it); free(pre); free(low); free(found); free(seen); free(queue); free(sccq);
  *sizes_out = sizes;
  return n_comp;
}

/* numbering 1: compute_SCC_and_Clusters (utils.py:30-52): components sorted by size (stable, ascending), then the nodes without
 * an active edge as singletons in index order.  numbering 0: canonical, label = smallest node id of the component. */
int po_scc_labels(const i64* src, const i64* dst, const i64* act64, i64 E, i64 N, int numbering, i64* labels, i64* n_comp_out) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !act64)) || (N > 0 && !labels)) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  i64* comp = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  if (!act || !comp) { free(act); free(comp); return PO_ERR_ALLOC; }
  for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
  i64* sizes = NULL;
  const i64 nc = scc_emission_order(src, dst, act, E, N, comp, &sizes);
  free(act);
  if (nc < 0) { free(comp); return PO_ERR_ALLOC; }
  i64 total = 0;
  if (numbering == 1) {
    /* stable counting sort of the components by size */
    i64* cnt = (i64*)calloc((size_t)N + 2, sizeof(i64));
    i64* rank = (i64*)malloc(sizeof(i64) * (size_t)(nc > 0 ? nc : 1));
    if (!cnt || !rank) { free(cnt); free(rank); free(comp); free(sizes); return PO_ERR_ALLOC; }
    for (i64 c = 0; c < nc; ++c) cnt[sizes[c] + 1]++;
    for (i64 s = 0; s <= N; ++s) cnt[s + 1] += cnt[s];
    for (i64 c = 0; c < nc; ++c) rank[c] = cnt[sizes[c]]++;
    total = nc;
    for (i64 v = 0; v < N; ++v) labels[v] = comp[v] >= 0 ? rank[comp[v]] : total++;
    free(cnt); free(rank);
  } else {
    i64* first = (i64*)malloc(sizeof(i64) * (size_t)(nc > 0 ? nc : 1));
    if (!first) { free(comp); free(sizes); return PO_ERR_ALLOC; }
    for (i64 c = 0; c < nc; ++c) first[c] = N;
    for (i64 v = 0; v < N; ++v) if (comp[v] >= 0 && v < first[comp[v]]) first[comp[v]] = v;
    total = nc;
    for (i64 v = 0; v < N; ++v) { if (comp[v] >= 0) labels[v] = first[comp[v]]; else { labels[v] = v; ++total; } }
    free(first);
  }
  if (n_comp_out) *n_comp_out = total;
  free(comp); free(sizes);
  return PO_OK;
}

/* ------------------------------------------------------------------------------------------------------------------
 * CUTTING (utils.py:125-142): an active (u,v) survives only if (v,u) is active as well.  rev[e] = index of (dst[e], src[e])
 * or -1; edges are unique.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { i64 key, idx; } keyed;
static int keyed_cmp(const void* a, const void* b) {
  const i64 x = ((const keyed*)a)->key, y = ((const keyed*)b)->key;
  return x < y ? -1 : x > y;
}
int po_reverse_map(const i64* src, const i64* dst, i64 E, i64 N, i64* rev) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !rev))) return PO_ERR_ARG;
  keyed* k = (keyed*)malloc(sizeof(keyed) * (size_t)(E > 0 ? E : 1));
  if (!k) return PO_ERR_ALLOC;
  int sorted = 1;
  for (i64 e = 0; e < E; ++e) {
    k[e].key = src[e] * N + dst[e];
    k[e].idx = e;
    if (e > 0 && k[e].key < k[e - 1].key) sorted = 0;
  }
  if (!sorted) qsort(k, (size_t)E, sizeof(keyed), keyed_cmp);
  /* first sorted position of every row: the search for (v, u) only walks row v */
  i64* row_beg = (i64*)malloc(sizeof(i64) * ((size_t)N + 1));
  if (!row_beg) { free(k); return PO_ERR_ALLOC; }
  {
    i64 pos = 0;
    for (i64 v = 0; v <= N; ++v) {
      while (pos < E && k[pos].key < v * N) ++pos;
      row_beg[v] = pos;
    }
  }
  for (i64 e = 0; e < E; ++e) {
    const i64 want = dst[e] * N + src[e];
    i64 lo = row_beg[dst[e]], hi = row_beg[dst[e] + 1];   /* first position with key >= want */
    while (lo < hi) {
      const i64 mid = (lo + hi) >> 1;
      if (k[mid].key < want) lo = mid + 1; else hi = mid;
    }
    rev[e] = (lo < E && k[lo].key == want) ? k[lo].idx : -1;
  }
  free(k); free(row_beg);
  return PO_OK;
}
static void cut_round(uint8_t* act, const i64* rev, i64 E, uint8_t* tmp) {
  for (i64 e = 0; e < E; ++e) tmp[e] = act[e] && rev[e] >= 0 && act[rev[e]];
  memcpy(act, tmp, (size_t)E);
}

/* ------------------------------------------------------------------------------------------------------------------
 * PRUNING (utils.py:161-188, 277-317) as rounds: while a node has more than C-1 active out- (in-) edges, every violating
 * node drops its active out- (in-) edge of minimum probability — first index on ties (torch.argmin) — all picks of a round
 * taken from the same snapshot.  Returns through *changed whether anything violated at all (the reference returns [] if not).
 * ---------------------------------------------------------------------------------------------------------------- */
static int prune_rounds(const i64* src, const i64* dst, uint8_t* act, const float* prob, i64 E, i64 N, int C, int* changed) {
  i64* fo = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* fi = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* bo = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* bi = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  if (!fo || !fi || !bo || !bi) { free(fo); free(fi); free(bo); free(bi); return PO_ERR_ALLOC; }
  *changed = 0;
  for (;;) {
    memset(fo, 0, sizeof(i64) * (size_t)N);
    memset(fi, 0, sizeof(i64) * (size_t)N);
    for (i64 e = 0; e < E; ++e) if (act[e]) { fo[src[e]]++; fi[dst[e]]++; }
    int any = 0;
    for (i64 v = 0; v < N; ++v) { bo[v] = bi[v] = -1; any |= fo[v] > C - 1 || fi[v] > C - 1; }
    if (!any) break;
    *changed = 1;
    for (i64 e = 0; e < E; ++e) {                    /* increasing edge id: a later equal probability does not replace the pick */
      if (!act[e]) continue;
      const i64 u = src[e], v = dst[e];
      if (fo[u] > C - 1 && (bo[u] < 0 || prob[e] < prob[bo[u]])) bo[u] = e;
      if (fi[v] > C - 1 && (bi[v] < 0 || prob[e] < prob[bi[v]])) bi[v] = e;
    }
    for (i64 v = 0; v < N; ++v) {
      if (bo[v] >= 0) act[bo[v]] = 0;
      if (bi[v] >= 0) act[bi[v]] = 0;
    }
  }
  free(fo); free(fi); free(bo); free(bi);
  return PO_OK;
}

/* ------------------------------------------------------------------------------------------------------------------
 * SPLITTING (utils.py:54-123) as rounds: every cluster larger than C finds the minimum probability among the active edges
 * with either endpoint in it; every edge ANYWHERE whose probability equals one of those minima is switched off (the
 * reference compares floats with ==, utils.py:96-98); recompute the clusters; repeat.
 * ---------------------------------------------------------------------------------------------------------------- */
static int float_cmp(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return x < y ? -1 : x > y;
}
static int split_rounds(const i64* src, const i64* dst, uint8_t* act, const float* prob, i64 E, i64 N, int C) {
  i64* comp = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  float* m = NULL;
  uint8_t* big = NULL;
  if (!comp) return PO_ERR_ALLOC;
  for (;;) {
    i64* sizes = NULL;
    const i64 nc = scc_emission_order(src, dst, act, E, N, comp, &sizes);
    if (nc < 0) { free(comp); return PO_ERR_ALLOC; }
    i64 n_big = 0;
    for (i64 c = 0; c < nc; ++c) n_big += sizes[c] > C;
    if (n_big == 0) { free(sizes); break; }
    m = (float*)malloc(sizeof(float) * (size_t)nc);
    big = (uint8_t*)malloc((size_t)nc);
    if (!m || !big) { free(m); free(big); free(sizes); free(comp); return PO_ERR_ALLOC; }
    for (i64 c = 0; c < nc; ++c) { big[c] = sizes[c] > C; m[c] = INFINITY; }
    free(sizes);
    for (i64 e = 0; e < E; ++e) {
      if (!act[e]) continue;
      const i64 cs = comp[src[e]], cd = comp[dst[e]];          /* both >= 0: the endpoints of an active edge are in the digraph */
      if (big[cs] && prob[e] < m[cs]) m[cs] = prob[e];
      if (big[cd] && prob[e] < m[cd]) m[cd] = prob[e];
    }
    i64 nv = 0;
    for (i64 c = 0; c < nc; ++c) if (big[c] && isfinite(m[c])) m[nv++] = m[c];
    qsort(m, (size_t)nv, sizeof(float), float_cmp);
    for (i64 e = 0; e < E; ++e) {
      if (!act[e]) continue;                         /* (switching an inactive edge off changes nothing) */
      i64 lo = 0, hi = nv;
      while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (m[mid] < prob[e]) lo = mid + 1; else hi = mid;
      }
      if (lo < nv && m[lo] == prob[e]) act[e] = 0;
    }
    free(m); free(big);
    m = NULL; big = NULL;
  }
  free(comp);
  return PO_OK;
}

/* ------------------------------------------------------------------------------------------------------------------
 * SPLITTING exactly as the reference orders it (utils.py:54-123), one cluster at a time: l = lowest label with more than C
 * nodes in the reference numbering; drop (globally, float ==) the minimum-probability value among the active edges touching
 * cluster l; relabel; continue on l AS RE-READ IN THE NEW NUMBERING (utils.py:112) while that cluster is oversized, else
 * look for the next oversized cluster.  One SCC pass per dropped value: for graphs up to ~10^5 active edges.  Differs from
 * split_rounds only under cross-cluster probability ties (oracle/postproc_oracle.py::split_rounds).
 * ---------------------------------------------------------------------------------------------------------------- */
static int reference_labels(const i64* src, const i64* dst, const uint8_t* act, i64 E, i64 N, i64* labels, i64* count /*[N]*/) {
  i64* comp = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* sizes = NULL;
  if (!comp) return PO_ERR_ALLOC;
  const i64 nc = scc_emission_order(src, dst, act, E, N, comp, &sizes);
  if (nc < 0) { free(comp); return PO_ERR_ALLOC; }
  i64* cnt = (i64*)calloc((size_t)N + 2, sizeof(i64));
  i64* rank = (i64*)malloc(sizeof(i64) * (size_t)(nc > 0 ? nc : 1));
  if (!cnt || !rank) { free(cnt); free(rank); free(comp); free(sizes); return PO_ERR_ALLOC; }
  for (i64 c = 0; c < nc; ++c) cnt[sizes[c] + 1]++;
  for (i64 k = 0; k <= N; ++k) cnt[k + 1] += cnt[k];
  for (i64 c = 0; c < nc; ++c) rank[c] = cnt[sizes[c]]++;
  i64 total = nc;
  memset(count, 0, sizeof(i64) * (size_t)N);
  for (i64 v = 0; v < N; ++v) { labels[v] = comp[v] >= 0 ? rank[comp[v]] : total++; count[labels[v]]++; }
  free(cnt); free(rank); free(comp); free(sizes);
  return PO_OK;
}
static int split_sequential(const i64* src, const i64* dst, uint8_t* act, const float* prob, i64 E, i64 N, int C) {
  i64* labels = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* count = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  if (!labels || !count) { free(labels); free(count); return PO_ERR_ALLOC; }
  int rc = reference_labels(src, dst, act, E, N, labels, count);
  while (rc == PO_OK) {
    i64 l = -1;
    for (i64 k = 0; k < N; ++k) if (count[k] > C) { l = k; break; }       /* lowest oversized label (utils.py:60-64) */
    if (l < 0) break;
    for (;;) {
      float m = INFINITY;
      for (i64 e = 0; e < E; ++e)                                          /* either endpoint in the cluster (utils.py:71) */
        if (act[e] && (labels[src[e]] == l || labels[dst[e]] == l) && prob[e] < m) m = prob[e];
      for (i64 e = 0; e < E; ++e) if (prob[e] == m) act[e] = 0;            /* global float equality (utils.py:96-98) */
      rc = reference_labels(src, dst, act, E, N, labels, count);
      if (rc != PO_OK || !(count[l] > C)) break;                           /* l re-read in the NEW numbering (utils.py:112) */
    }
  }
  free(labels); free(count);
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * SPLITTING, exact AND scalable (design study for the CUDA path): the reference's order only
 * matters between clusters that can meet a probability tie.  T = the probability values carried by two or more active edges
 * that touch oversized clusters when SPLITTING starts (edges never gain clusters and clusters never grow, so no tie outside T
 * can ever arise).  An oversized cluster is TAINTED while an active edge touching it has a value in T.
 *   - untainted oversized clusters: all drop their minimum in the same iteration (their values are not in T, so they cannot
 *     touch another oversized cluster: order-independent, as in split_rounds);
 *   - tainted oversized clusters: one per iteration, the lowest label of the reference numbering, exactly as the reference
 *     picks it (utils.py:60-64,112).
 * Equal to split_sequential (tests/test_c_oracle.py) except for one third-order effect that is not modelled: the emission
 * order of two tainted clusters of EQUAL size could depend on whether an untainted cluster that connects them through
 * one-directional edges has already been processed.
 * ---------------------------------------------------------------------------------------------------------------- */
static int in_sorted(const float* v, i64 n, float x) {
  i64 lo = 0, hi = n;
  while (lo < hi) {
    const i64 mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && v[lo] == x;
}
static int split_hybrid(const i64* src, const i64* dst, uint8_t* act, const float* prob, i64 E, i64 N, int C, i64* stats /*[3]*/) {
  i64* labels = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  i64* count = (i64*)malloc(sizeof(i64) * (size_t)(N > 0 ? N : 1));
  uint8_t* tainted = (uint8_t*)malloc((size_t)(N > 0 ? N : 1));
  float* mins = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  float* tie = NULL;
  float* vals = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  i64 n_tie = 0;
  if (!labels || !count || !tainted || !mins || !vals) { free(labels); free(count); free(tainted); free(mins); free(vals); return PO_ERR_ALLOC; }
  int rc = reference_labels(src, dst, act, E, N, labels, count);
  if (rc == PO_OK) {                                   /* T: duplicated values among the active edges touching oversized clusters */
    i64 n = 0;
    for (i64 e = 0; e < E; ++e) n += act[e] && (count[labels[src[e]]] > C || count[labels[dst[e]]] > C);
    tie = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    if (!tie) rc = PO_ERR_ALLOC;
    else {
      i64 k = 0;
      for (i64 e = 0; e < E; ++e)
        if (act[e] && (count[labels[src[e]]] > C || count[labels[dst[e]]] > C)) tie[k++] = prob[e];
      qsort(tie, (size_t)n, sizeof(float), float_cmp);
      for (i64 i = 0; i + 1 < n; ++i)
        if (tie[i] == tie[i + 1] && (n_tie == 0 || tie[n_tie - 1] != tie[i])) tie[n_tie++] = tie[i];
    }
  }
  stats[0] = n_tie; stats[1] = stats[2] = 0;           /* tie values, iterations, tainted steps */
  i64 cur = -1;                                        /* label the reference's inner loop is on */
  while (rc == PO_OK) {
    int any = 0;
    for (i64 k = 0; k < N; ++k) { tainted[k] = 0; mins[k] = INFINITY; any |= count[k] > C; }
    if (!any) break;
    ++stats[1];
    for (i64 e = 0; e < E; ++e) {
      if (!act[e]) continue;
      const i64 ls = labels[src[e]], ld = labels[dst[e]];
      const int hit = n_tie > 0 && in_sorted(tie, n_tie, prob[e]);
      if (count[ls] > C) { if (prob[e] < mins[ls]) mins[ls] = prob[e]; tainted[ls] |= hit; }
      if (count[ld] > C) { if (prob[e] < mins[ld]) mins[ld] = prob[e]; tainted[ld] |= hit; }
    }
    i64 nv = 0;
    /* the reference stays on label l — RE-READ in the new numbering (utils.py:112; nodes that lost every active edge move to
     * the end of the numbering, so this need not be the lowest oversized label) — while that cluster is oversized */
    if (!(cur >= 0 && cur < N && count[cur] > C && tainted[cur])) {
      cur = -1;
      for (i64 k = 0; k < N; ++k) if (count[k] > C && tainted[k]) { cur = k; break; }
    }
    for (i64 k = 0; k < N; ++k)
      if (count[k] > C && !tainted[k]) vals[nv++] = mins[k];
    if (cur >= 0) { vals[nv++] = mins[cur]; ++stats[2]; }
    qsort(vals, (size_t)nv, sizeof(float), float_cmp);
    for (i64 e = 0; e < E; ++e) if (act[e] && in_sorted(vals, nv, prob[e])) act[e] = 0;
    rc = reference_labels(src, dst, act, E, N, labels, count);
  }
  free(labels); free(count); free(tainted); free(mins); free(tie); free(vals);
  return rc;
}

/* stage entry points (each takes and returns int64 activity vectors like the reference's `predictions`) */
int po_cut(const i64* src, const i64* dst, i64* act64, i64 E, i64 N) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !act64))) return PO_ERR_ARG;
  i64* rev = (i64*)malloc(sizeof(i64) * (size_t)(E > 0 ? E : 1));
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  uint8_t* tmp = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  int rc = (!rev || !act || !tmp) ? PO_ERR_ALLOC : po_reverse_map(src, dst, E, N, rev);
  if (rc == PO_OK) {
    for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
    cut_round(act, rev, E, tmp);
    for (i64 e = 0; e < E; ++e) act64[e] = act[e];
  }
  free(rev); free(act); free(tmp);
  return rc;
}
int po_prune(const i64* src, const i64* dst, i64* act64, const float* prob, i64 E, i64 N, int num_cameras, int* changed) {
  if (E < 0 || N < 0 || !changed || (E > 0 && (!src || !dst || !act64 || !prob))) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  if (!act) return PO_ERR_ALLOC;
  for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
  const int rc = prune_rounds(src, dst, act, prob, E, N, num_cameras, changed);
  if (rc == PO_OK) for (i64 e = 0; e < E; ++e) act64[e] = act[e];
  free(act);
  return rc;
}
int po_split(const i64* src, const i64* dst, i64* act64, const float* prob, i64 E, i64 N, int num_cameras) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !act64 || !prob))) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  if (!act) return PO_ERR_ALLOC;
  for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
  const int rc = split_rounds(src, dst, act, prob, E, N, num_cameras);
  if (rc == PO_OK) for (i64 e = 0; e < E; ++e) act64[e] = act[e];
  free(act);
  return rc;
}

int po_split_sequential(const i64* src, const i64* dst, i64* act64, const float* prob, i64 E, i64 N, int num_cameras) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !act64 || !prob))) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  if (!act) return PO_ERR_ALLOC;
  for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
  const int rc = split_sequential(src, dst, act, prob, E, N, num_cameras);
  if (rc == PO_OK) for (i64 e = 0; e < E; ++e) act64[e] = act[e];
  free(act);
  return rc;
}

int po_split_hybrid(const i64* src, const i64* dst, i64* act64, const float* prob, i64 E, i64 N, int num_cameras, i64* stats) {
  if (E < 0 || N < 0 || !stats || (E > 0 && (!src || !dst || !act64 || !prob))) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  if (!act) return PO_ERR_ALLOC;
  for (i64 e = 0; e < E; ++e) act[e] = act64[e] != 0;
  const int rc = split_hybrid(src, dst, act, prob, E, N, num_cameras, stats);
  if (rc == PO_OK) for (i64 e = 0; e < E; ++e) act64[e] = act[e];
  free(act);
  return rc;
}

/* inference.post_processing (inference.py:70-169): CUT -> PRUNE -> CUT -> SPLIT -> labels.  pred: the reference's
 * `predictions` (argmax of the logits), prob: softmax(logits)[:, 1].  numbering as in po_scc_labels. */
int po_post_processing(const i64* src, const i64* dst, const i64* pred, const float* prob, i64 E, i64 N, int num_cameras,
                       int cutting, int pruning, int splitting, int numbering, i64* labels_out, i64* act_out) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst || !pred || !prob || !act_out)) || (N > 0 && !labels_out)) return PO_ERR_ARG;
  for (i64 e = 0; e < E; ++e)
    if (src[e] < 0 || src[e] >= N || dst[e] < 0 || dst[e] >= N) return PO_ERR_ARG;
  uint8_t* act = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  uint8_t* tmp = (uint8_t*)malloc((size_t)(E > 0 ? E : 1));
  i64* rev = cutting ? (i64*)malloc(sizeof(i64) * (size_t)(E > 0 ? E : 1)) : NULL;
  int rc = (!act || !tmp || (cutting && !rev)) ? PO_ERR_ALLOC : PO_OK;
  if (rc == PO_OK) for (i64 e = 0; e < E; ++e) act[e] = pred[e] != 0;
  if (rc == PO_OK && cutting) rc = po_reverse_map(src, dst, E, N, rev);
  if (rc == PO_OK && cutting) cut_round(act, rev, E, tmp);
  if (rc == PO_OK && pruning) { int changed; rc = prune_rounds(src, dst, act, prob, E, N, num_cameras, &changed); }
  if (rc == PO_OK && cutting) cut_round(act, rev, E, tmp);
  if (rc == PO_OK && splitting) rc = split_rounds(src, dst, act, prob, E, N, num_cameras);
  if (rc == PO_OK) {
    for (i64 e = 0; e < E; ++e) act_out[e] = act[e];
    rc = po_scc_labels(src, dst, act_out, E, N, numbering, labels_out, NULL);
  }
  free(act); free(tmp); free(rev);
  return rc;
}
