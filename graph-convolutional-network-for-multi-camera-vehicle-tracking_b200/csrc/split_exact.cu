// SPLITTING (utils.py:54-123) in the reference's own one-cluster-at-a-time order, for the clusters whose outcome depends on
// that order (probability ties).  Host code: the reference's order is sequential by definition; what this file removes is the
// cost of following it — the reference recomputes every strongly connected component of the whole graph after each dropped
// value, here a step only recomputes the weakly connected component(s) it touched.
//
// What the reference does (utils.py:54-123, compute_SCC_and_Clusters utils.py:30-52):
//   labels = position in  sorted(SCCs of the active digraph, key=len)  (stable; networkx emission order within one size), nodes
//   without an active edge appended last;  l = lowest label with more than C nodes;  loop: m = min probability over the active
//   edges with an endpoint in cluster l;  every edge of the graph with probability == m is switched off (float ==, :96-98);
//   relabel;  stay on the INTEGER l (re-read in the new numbering, :112) while that label is oversized, else start over.
//
// Facts used (each checked against the unmodified reference's outputs in tests/):
//   * networkx emits the SCCs source by source (sources in first-appearance order of the nodes in the active edge list, u
//     before v), each source's DFS in post-order; a DFS never leaves the weakly connected component (WCC) of its source.  So
//     the emission order is the lexicographic order of (first-appearance key of the emitting source, index within that DFS),
//     and both parts are functions of the WCC alone: a step only invalidates the keys of the WCCs that lost an edge.
//   * every small (<= C nodes) cluster sorts before every oversized one, so the label of the i-th oversized cluster is
//     n_small + i and "label l is still oversized" reads  0 <= l - n_small' < n_big'  after the step: only the CHANGE of the
//     number of small clusters matters, never its value.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <iterator>
#include <set>
#include <thread>
#include <tuple>
#include <vector>

#include <cstring>

#include "common.cuh"

namespace mpn {

namespace {

struct SplitExact {
  long long m;
  const float* prob;
  int C, n;                                       // camera bound, number of (local) nodes
  const int *ls = nullptr, *ld = nullptr;         // endpoints of every edge: the caller's arrays (global ids) or ls_own / ld_own
  std::vector<int> ls_own, ld_own;                // (ids relabelled in order of first use, scc_emission_keys_host)
  std::vector<int> out_ptr, out_adj, in_ptr, in_adj;     // incident edge ids per node, ascending (= edge order)
  std::vector<int> out_first, in_first;           // first entry of the lists that may still be alive
  std::vector<uint8_t> alive;
  std::vector<int> comp, wcc;                     // cluster / WCC id of a node (-1: no active edge)
  std::vector<int> pre, low, it, mark;            // traversal scratch
  struct Cluster { int size; long long src_key; int idx; std::vector<int> nodes; };
  std::vector<Cluster> clusters;
  std::vector<std::vector<int>> wcc_nodes;
  typedef std::tuple<int, long long, int, int> BigKey;      // (size, key of the emitting source, index in its DFS, cluster id)
  std::set<BigKey> big;
  long long n_small = 0;
  int mark_gen = 0;
  std::vector<int> sc_stack, sc_wn, sc_dfs, sc_scc;          // recompute() scratch (a step recomputes a handful of nodes: no
  std::vector<std::pair<long long, int>> sc_keyed;           // allocation per call)

  long long first_key(int v) {                    // first-appearance key of v in the current active edge list: 2*edge + side
    int& a = out_first[v];
    while (a < out_ptr[v + 1] && !alive[out_adj[a]]) ++a;
    int& b = in_first[v];
    while (b < in_ptr[v + 1] && !alive[in_adj[b]]) ++b;
    long long k = INT64_MAX;
    if (a < out_ptr[v + 1]) k = 2ll * out_adj[a];
    if (b < in_ptr[v + 1]) k = std::min(k, 2ll * in_adj[b] + 1);
    return k;
  }

  void unregister_wcc(int w) {
    for (int v : wcc_nodes[w]) {
      const int c = comp[v];
      if (c < 0 || clusters[c].size == 0) continue;
      if (clusters[c].size > C) big.erase(BigKey(clusters[c].size, clusters[c].src_key, clusters[c].idx, c));
      else --n_small;
      clusters[c].size = 0;                       // dead
      std::vector<int>().swap(clusters[c].nodes);
    }
  }

  // nodes: a set closed under the alive edges.  Splits it into WCCs, runs networkx's SCC generator on each, registers clusters.
  void recompute(const std::vector<int>& nodes, bool lazy_only = false) {
    ++mark_gen;
    std::vector<int>&stack = sc_stack, &wn = sc_wn, &dfs = sc_dfs, &scc_stack = sc_scc;
    std::vector<std::pair<long long, int>>& keyed = sc_keyed;
    scc_stack.clear();
    for (int seed : nodes) {
      if (mark[seed] == mark_gen || (lazy_only && wcc[seed] != -2)) continue;
      mark[seed] = mark_gen;
      wn.clear();
      stack.assign(1, seed);
      while (!stack.empty()) {                    // undirected traversal over the alive edges
        const int v = stack.back();
        stack.pop_back();
        wn.push_back(v);
        for (int k = out_first[v]; k < out_ptr[v + 1]; ++k) {
          const int e = out_adj[k];
          if (alive[e] && mark[ld[e]] != mark_gen) { mark[ld[e]] = mark_gen; stack.push_back(ld[e]); }
        }
        for (int k = in_first[v]; k < in_ptr[v + 1]; ++k) {
          const int e = in_adj[k];
          if (alive[e] && mark[ls[e]] != mark_gen) { mark[ls[e]] = mark_gen; stack.push_back(ls[e]); }
        }
      }
      if (wn.size() == 1 && first_key(wn[0]) == INT64_MAX) {      // no active edge left: not in the digraph (utils.py:34-42)
        comp[wn[0]] = -1;
        wcc[wn[0]] = -1;
        continue;
      }
      const int w = (int)wcc_nodes.size();
      keyed.clear();
      for (int v : wn) { keyed.emplace_back(first_key(v), v); wcc[v] = w; comp[v] = -2; pre[v] = 0; it[v] = out_first[v]; }
      std::sort(keyed.begin(), keyed.end());
      wcc_nodes.emplace_back(wn);
      int counter = 0;
      for (const auto& kv : keyed) {              // sources in node insertion order of nx.DiGraph(active edge list)
        const int source = kv.second;
        if (comp[source] != -2) continue;
        int emitted = 0;
        dfs.assign(1, source);
        while (!dfs.empty()) {
          const int v = dfs.back();
          if (pre[v] == 0) pre[v] = ++counter;
          bool done = true;
          while (it[v] < out_ptr[v + 1]) {
            const int e = out_adj[it[v]++];
            if (!alive[e]) continue;
            const int u = ld[e];
            if (pre[u] == 0) { dfs.push_back(u); done = false; break; }
          }
          if (!done) continue;
          int lo = pre[v];
          for (int k = out_first[v]; k < out_ptr[v + 1]; ++k) {
            const int e = out_adj[k];
            if (!alive[e]) continue;
            const int u = ld[e];
            if (comp[u] == -2) lo = std::min(lo, pre[u] > pre[v] ? low[u] : pre[u]);
          }
          low[v] = lo;
          dfs.pop_back();
          if (lo == pre[v]) {
            const int c = (int)clusters.size();
            clusters.emplace_back();
            Cluster& cl = clusters.back();
            cl.nodes.push_back(v);
            comp[v] = c;
            while (!scc_stack.empty() && pre[scc_stack.back()] > pre[v]) {
              comp[scc_stack.back()] = c;
              cl.nodes.push_back(scc_stack.back());
              scc_stack.pop_back();
            }
            cl.size = (int)cl.nodes.size();
            cl.src_key = kv.first;
            cl.idx = emitted++;
            if (cl.size > C) big.insert(BigKey(cl.size, cl.src_key, cl.idx, c));
            else ++n_small;
          } else {
            scc_stack.push_back(v);
          }
        }
      }
    }
  }
};

}  // namespace

// keep_out[i] = 0 for the edges SPLITTING switches off.  stats_out[4]: dropped values, steps taken on a label that was not the
// lowest oversized one (utils.py:112 re-reads the integer), clusters examined, 0.
// builds the adjacency of the sub-problem (local node ids in order of first use) and registers every cluster
// sel (optional): entries with sel[i] == 0 are not part of the sub-problem (they keep their edge id: ids only need to be ordered)
static void split_exact_init(SplitExact& S, const int* src, const int* dst, const float* prob, long long m, int n_nodes, int C,
                             std::vector<int>* global_of_local, const int* seeds = nullptr, long long n_seeds = 0,
                             const uint8_t* sel = nullptr) {
  const auto t_init0 = std::chrono::steady_clock::now();
  S.m = m; S.prob = prob; S.C = C;
  std::vector<int> local;
  int n = 0;
  if (global_of_local != nullptr) {               // relabel in order of first use (the caller wants the list of nodes that appear)
    local.assign((size_t)n_nodes, -1);
    S.ls_own.resize(m); S.ld_own.resize(m);
    for (long long i = 0; i < m; ++i) {
      if (sel && !sel[i]) continue;
      if (local[src[i]] < 0) { local[src[i]] = n++; global_of_local->push_back(src[i]); }
      if (local[dst[i]] < 0) { local[dst[i]] = n++; global_of_local->push_back(dst[i]); }
      S.ls_own[i] = local[src[i]];
      S.ld_own[i] = local[dst[i]];
    }
    S.ls = S.ls_own.data(); S.ld = S.ld_own.data();
  } else {                                        // global ids as they are: no pass over the edges, no copy
    S.ls = src; S.ld = dst;
    n = n_nodes;
  }
  S.n = n;
  // incident-edge lists per node (ascending edge id = edge order).  The in-lists are built by a helper thread while this one
  // builds the out-lists and the per-node state (independent arrays; at 4 M edges each pass is ~10 ms of cache misses).
  auto build_lists = [&](const int* end_of, std::vector<int>& ptr, std::vector<int>& adj, std::vector<int>& first, long long* count) {
    ptr.assign(n + 1, 0);
    long long c = 0;
    for (long long i = 0; i < m; ++i) {
      if (sel && !sel[i]) continue;
      ptr[end_of[i] + 1]++;
      ++c;
    }
    for (int v = 0; v < n; ++v) ptr[v + 1] += ptr[v];
    adj.resize(c);
    first.assign(ptr.begin(), ptr.end() - 1);              // (doubles as the fill cursor, restored below)
    for (long long i = 0; i < m; ++i) {
      if (sel && !sel[i]) continue;
      adj[first[end_of[i]]++] = (int)i;
    }
    for (int v = 0; v < n; ++v) first[v] = ptr[v];
    if (count) *count = c;
  };
  long long m_sel = 0;
  std::thread in_builder([&] { build_lists(S.ld, S.in_ptr, S.in_adj, S.in_first, nullptr); });
  build_lists(S.ls, S.out_ptr, S.out_adj, S.out_first, &m_sel);
  if (sel) S.alive.assign(sel, sel + m); else S.alive.assign(m, 1);
  S.comp.assign(n, -1); S.wcc.assign(n, -2);               // -2: component not examined yet (registered on first use)
  S.pre.assign(n, 0); S.low.assign(n, 0); S.it.assign(n, 0); S.mark.assign(n, 0);
  in_builder.join();
  if (getenv("MPN_POST_DEBUG") != nullptr)
    fprintf(stderr, "[split engine] adjacency of %lld edges / %d nodes built at +%.2f ms\n", m_sel, n,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_init0).count());
  if (seeds == nullptr) {
    std::vector<int> all(n);
    for (int v = 0; v < n; ++v) all[v] = v;
    S.recompute(all);
  } else {
    // only the weakly connected components that hold an oversized cluster are examined now; a component that merely carries a
    // tied probability value is examined when (if) that value is dropped
    std::vector<int> start;
    start.reserve((size_t)n_seeds);
    for (long long i = 0; i < n_seeds; ++i) {
      int v = seeds[i] >= 0 && seeds[i] < n_nodes ? seeds[i] : -1;
      if (v >= 0 && !local.empty()) v = local[(size_t)v];
      if (v >= 0) start.push_back(v);
    }
    S.recompute(start);
  }
}

// networkx's emission order for a sub-graph that is a union of whole weakly connected components, given as its active edges in
// edge order: for every node that appears, the key of its strongly connected component = (first-appearance key of the source
// whose DFS emitted it, in units of 2 * local edge index + side; index within that DFS; size).  Used by the reference label
// numbering for the components that one-directional edges tie together (postproc.cu).
void scc_emission_keys_host(const int* src, const int* dst, long long m, int n_nodes, std::vector<int>& node_out,
                            std::vector<long long>& key_out, std::vector<int>& idx_out, std::vector<int>& size_out) {
  SplitExact S;
  node_out.clear();
  split_exact_init(S, src, dst, nullptr, m, n_nodes, 0x7fffffff, &node_out);
  key_out.resize(S.n); idx_out.resize(S.n); size_out.resize(S.n);
  for (int v = 0; v < S.n; ++v) {
    const SplitExact::Cluster& c = S.clusters[S.comp[v]];
    key_out[v] = c.src_key; idx_out[v] = c.idx; size_out[v] = c.size;
  }
}

// tied (optional, [m]): 1 for an edge whose probability value is carried by at least two active edges of the graph (only those can
// be switched off by another cluster's step); nullptr: found here by sorting.  seeds (optional): global ids of the nodes of the
// oversized clusters; nullptr: every component is examined up front.
int split_exact_host_impl(const int* src, const int* dst, const float* prob, long long m, int n_nodes, int C, uint8_t* keep,
                          int64_t* stats_out, const uint8_t* tied, const int* seeds, long long n_seeds, const uint8_t* sel) {
  SplitExact S;
  static const bool dbg = getenv("MPN_POST_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // probability value (bits) -> the edges that carry it, for the values carried by more than one edge: open-addressing table of
  // chain heads over the list of tied edges (no allocation per value; built in one pass)
  auto bits_of = [&](int e) { uint32_t b; float f = prob[e] == 0.0f ? 0.0f : prob[e]; memcpy(&b, &f, 4); return b; };
  std::vector<uint8_t> tied_local;
  std::vector<int> tied_ids;
  size_t tcap = 16;
  const uint32_t TIE_EMPTY = 0xFFFFFFFFu;
  std::vector<uint32_t> tkeys;
  std::vector<int> thead, tnext;
  auto tie_slot = [&](uint32_t b) {
    size_t h = ((size_t)b * 2654435761u) & (tcap - 1);
    while (tkeys[h] != TIE_EMPTY && tkeys[h] != b) h = (h + 1) & (tcap - 1);
    return h;
  };
  auto build_tie_index = [&] {
    if (tied == nullptr) {                          // stand-alone call: find the shared values by sorting
      std::vector<int> order;
      order.reserve(m);
      for (long long i = 0; i < m; ++i)
        if (!sel || sel[i]) order.push_back((int)i);
      std::sort(order.begin(), order.end(), [&](int a, int b) { return prob[a] < prob[b] || (prob[a] == prob[b] && a < b); });
      tied_local.assign(m, 0);
      const long long mo = (long long)order.size();
      for (long long i = 0; i < mo;) {
        long long j = i + 1;
        while (j < mo && prob[order[j]] == prob[order[i]]) ++j;
        if (j - i > 1)
          for (long long k = i; k < j; ++k) tied_local[order[k]] = 1;
        i = j;
      }
      tied = tied_local.data();
    }
    for (long long i = 0; i < m; ++i)
      if (tied[i] && (!sel || sel[i])) tied_ids.push_back((int)i);
    while (tcap < 2 * tied_ids.size() + 2) tcap <<= 1;
    tkeys.assign(tcap, TIE_EMPTY);
    thead.assign(tcap, -1);
    tnext.assign(tied_ids.size(), -1);
    for (size_t t = 0; t < tied_ids.size(); ++t) {
      const uint32_t b = bits_of(tied_ids[t]);
      if (b == TIE_EMPTY) continue;                 // (a NaN pattern: never equal to anything)
      const size_t h = tie_slot(b);
      tkeys[h] = b;
      tnext[t] = thead[h];
      thead[h] = (int)t;
    }
  };
  // the tie index only reads the caller's arrays: it is built by a helper thread while this one builds the adjacency
  std::thread tie_builder(build_tie_index);
  for (long long i = 0; i < m; ++i) keep[i] = 1;
  split_exact_init(S, src, dst, prob, m, n_nodes, C, nullptr, seeds, n_seeds, sel);
  const double t1 = now();
  tie_builder.join();
  const double t2 = now();
  long long steps = 0, off_lowest = 0;
  long long sticky = -1;                          // index (in the order of `big`) of the cluster the reference's inner loop is on
  std::vector<int> affected, kill, lazy;
  std::vector<long long> wcc_stamp;
  while (!S.big.empty()) {
    const long long idx = (sticky >= 0) ? sticky : 0;
    if (idx > 0) ++off_lowest;
    auto itb = S.big.begin();
    std::advance(itb, idx);
    const int c = std::get<3>(*itb);
    // minimum probability over the active edges with an endpoint in the cluster (utils.py:69-95)
    float mn = INFINITY;
    int e_min = -1;
    for (int v : S.clusters[c].nodes) {
      for (int k = S.out_first[v]; k < S.out_ptr[v + 1]; ++k) { const int e = S.out_adj[k]; if (S.alive[e] && prob[e] < mn) { mn = prob[e]; e_min = e; } }
      for (int k = S.in_first[v]; k < S.in_ptr[v + 1]; ++k) { const int e = S.in_adj[k]; if (S.alive[e] && prob[e] < mn) { mn = prob[e]; e_min = e; } }
    }
    if (e_min < 0) { set_error("split: an oversized cluster without a finite edge probability"); return MPN_ERR_INVALID; }
    // every edge of the graph with that probability (float ==, utils.py:96-98): the minimum edge itself and, when the value is
    // carried by more than one edge, the others
    kill.clear();
    kill.push_back(e_min);
    if (tied[e_min]) {
      const uint32_t b = bits_of(e_min);
      if (b != TIE_EMPTY)
        for (int t = thead[tie_slot(b)]; t >= 0; t = tnext[t]) {
          const int e = tied_ids[t];
          if (e != e_min && S.alive[e]) kill.push_back(e);
        }
    }
    // components that have not been examined yet (they hold no oversized cluster) enter the books before the step is counted
    lazy.clear();
    for (int e : kill) {
      if (S.wcc[S.ls[e]] == -2) lazy.push_back(S.ls[e]);
      if (S.wcc[S.ld[e]] == -2) lazy.push_back(S.ld[e]);
    }
    if (!lazy.empty()) S.recompute(lazy, true);
    const long long small_before = S.n_small;
    affected.clear();
    for (int e : kill) {
      S.alive[e] = 0;
      keep[e] = 0;
      const int w = S.wcc[S.ls[e]];
      if ((size_t)w >= wcc_stamp.size()) wcc_stamp.resize(S.wcc_nodes.size(), -1);
      if (wcc_stamp[w] != steps) { wcc_stamp[w] = steps; affected.push_back(w); }
    }
    ++steps;
    for (int w : affected) S.unregister_wcc(w);
    for (int w : affected) {
      std::vector<int> nodes;
      nodes.swap(S.wcc_nodes[w]);
      S.recompute(nodes);
    }
    // np.bincount(ID_pred)[l] > num_cameras with l the INTEGER of the cluster just handled (utils.py:112)
    const long long j = idx - (S.n_small - small_before);
    sticky = (j >= 0 && j < (long long)S.big.size()) ? j : -1;
  }
  if (dbg) fprintf(stderr, "[split engine] m=%lld: init %.2f ms, tie map %.2f ms, %lld steps %.2f ms\n", m, t1 - t0, t2 - t1, steps, now() - t2);
  if (stats_out) { stats_out[0] = steps; stats_out[1] = off_lowest; stats_out[2] = (long long)S.clusters.size(); stats_out[3] = 0; }
  return MPN_OK;
}

}  // namespace mpn

extern "C" int mpn_split_exact_host(const int32_t* src, const int32_t* dst, const float* prob, int64_t n_active, int32_t n_nodes,
                                    int32_t num_cameras, uint8_t* keep_out, int64_t* stats_out) {
  MPN_REQUIRE(n_nodes > 0 && n_active >= 0 && n_active < (1ll << 31) && num_cameras >= 1 &&
                  (n_active == 0 || (src && dst && prob && keep_out)),
              "split_exact_host: bad arguments");
  for (int64_t i = 0; i < n_active; ++i)
    MPN_REQUIRE(src[i] >= 0 && src[i] < n_nodes && dst[i] >= 0 && dst[i] < n_nodes, "split_exact_host: node id out of range");
  if (n_active == 0) {
    if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = stats_out[3] = 0;
    return MPN_OK;
  }
  return mpn::split_exact_host_impl(src, dst, prob, n_active, n_nodes, num_cameras, keep_out, stats_out, nullptr, nullptr, 0, nullptr);
}
