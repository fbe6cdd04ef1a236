// Sparse primal-dual blossom matcher with pricing, for large matching graphs.
//
// The matching graphs of this path are complete graphs on the K odd cubes (reference:
// decoders/mwpm/matching.py:125-173), but their weights are lattice path lengths: the optimal
// matching almost only uses edges between near neighbours.  So the optimum is computed on a
// sparse candidate graph (the k nearest partners of every cube) and then PROVED optimal for the
// complete graph by pricing: the dual solution of the sparse problem is checked against every
// one of the K(K-1)/2 edges, edges with negative reduced cost are added, and the solver is
// warm-started from the repaired duals until no edge is violated.  By LP duality (Edmonds'
// perfect-matching polytope: vertex duals free, blossom duals >= 0) the final matching is a
// minimum-weight perfect matching of the complete graph - the same optimum as handing the solver
// every edge (solve_one in mwpm.cpp), only the choice among ties may differ.
//
// The solver is the classical primal-dual formulation with adjacency lists (labels S/T on
// top-level blossoms, dual adjustment by the minimum of the free-vertex / S-S / T-blossom slacks,
// blossom expansion when a T-blossom's dual reaches zero), on weights -2*len so that vertex duals
// stay integral and S-S slacks stay even - organised so that the work follows the part of the
// forest that changes: every free vertex roots a tree once, an augmentation retires its two trees
// and keeps the others, dual adjustments are lazy (one running sum), and the next adjustment
// comes from three heaps instead of a pass over the forest (see SparseBlossom::solve).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <utility>
#include <vector>

namespace fpb_sparse {

typedef long long ll;

class SparseBlossom {
 public:
  int n = 0;
  struct EdgeRec {  // one cache line holds four edges: endpoints and weight are read together
    int u, v;
    ll w;
  };
  std::vector<EdgeRec> E;
  std::vector<ll> dual;   // [0, n): vertices (2 x LP dual); [n, 2n): blossoms
  std::vector<int> mate;  // remote endpoint of the matched edge, -1 = free
  long long n_stages = 0, n_deltas = 0, n_augmentations = 0;

  void reset(int n_) {
    n = n_;
    E.clear();
    n_stages = n_deltas = n_augmentations = 0;
  }
  int n_edges() const { return (int)E.size(); }
  void add_edge(int u, int v, ll w) { E.push_back(EdgeRec{u, v, w}); }
  int endpoint(int p) const { return (p & 1) ? E[p >> 1].v : E[p >> 1].u; }
  // the dual of vertex x as of now (see solve(): adjustments are applied lazily)
  ll act(int x) const { return dual[x] + voff[x] * Delta; }
  ll slack(int k) const {
    const EdgeRec& e = E[k];
    return act(e.u) + act(e.v) - 2 * e.w;
  }
  int mate_vertex(int v) const { return mate[v] < 0 ? -1 : endpoint(mate[v]); }

  void build_adjacency() {
    const int ne = n_edges();
    adj_off.assign(n + 1, 0);
    for (int k = 0; k < ne; ++k) {
      ++adj_off[E[k].u + 1];
      ++adj_off[E[k].v + 1];
    }
    for (int v = 0; v < n; ++v) adj_off[v + 1] += adj_off[v];
    adj.resize(2 * (size_t)ne);
    fill_tmp.assign(adj_off.begin(), adj_off.end() - 1);
    for (int k = 0; k < ne; ++k) {
      adj[fill_tmp[E[k].u]++] = AdjRec{2 * k + 1, E[k].v, E[k].w};
      adj[fill_tmp[E[k].v]++] = AdjRec{2 * k, E[k].u, E[k].w};
    }
  }

  bool is_bare(int v) const { return inblossom[v] == v; }  // top-level, not inside a blossom
  int top_blossom(int v) const { return inblossom[v]; }    // v itself, or the outermost blossom around it (>= n)
  void double_all() {
    for (size_t i = 0; i < dual.size(); ++i) dual[i] *= 2;
    for (size_t k = 0; k < E.size(); ++k) E[k].w *= 2;
    for (size_t i = 0; i < adj.size(); ++i) adj[i].wt *= 2;
  }
  // Dissolve the blossoms around v: a blossom's dual z moves onto its vertices (z / 2 each in LP
  // units), which keeps every inner edge's reduced cost, adds slack to the edges leaving it and
  // breaks only the tightness of the base's matched edge - so the base and its partner become free.
  void make_bare(int v) {
    while (inblossom[v] != v) {
      const int b = inblossom[v];
      const ll z = dual[b];
      if (z != 0) {
        leaves(b, [&](int x) { dual[x] += z; });
        dual[b] = 0;
        unmatch(blossombase[b]);
      }
      expand_blossom(b, true);
    }
  }
  void unmatch(int v) {
    if (mate[v] < 0) return;
    const int m = endpoint(mate[v]);
    mate[v] = -1;
    mate[m] = -1;
  }

  template <class F>
  void for_neighbours(int u, F&& f) const {
    for (int i = adj_off[u]; i < adj_off[u + 1]; ++i) f(adj[i].w, adj[i].wt);
  }

  // Start state: vertex duals y (feasible on every edge) and a matching on tight edges given as
  // vertex mates (-1 = free).  Free vertices must share the parity of their duals.
  void start(const std::vector<ll>& y, const std::vector<int>& vmate) {
    const int m = 2 * n;
    dual.assign(m, 0);
    for (int v = 0; v < n; ++v) dual[v] = y[v];
    mate.assign(n, -1);
    for (int v = 0; v < n; ++v) {
      if (vmate[v] < 0 || mate[v] >= 0) continue;
      for (int i = adj_off[v]; i < adj_off[v + 1]; ++i)
        if (adj[i].w == vmate[v]) {
          mate[v] = adj[i].p;
          mate[vmate[v]] = adj[i].p ^ 1;
          break;
        }
    }
    label.assign(m, 0);
    labelend.assign(m, -1);
    inblossom.resize(n);
    for (int v = 0; v < n; ++v) inblossom[v] = v;
    blossomparent.assign(m, -1);
    blossombase.assign(m, -1);
    for (int v = 0; v < n; ++v) blossombase[v] = v;
    if ((int)childs.size() < m) {
      childs.resize(m);
      endps.resize(m);
    }
    for (int b = 0; b < m; ++b) {
      childs[b].clear();
      endps[b].clear();
    }
    unused.clear();
    for (int b = m - 1; b >= n; --b) unused.push_back(b);
    voff.assign(n, 0);
    boff.assign(m, 0);
    minkey.assign(n, KEY_NONE);
    Delta = 0;
    stamp.assign(m, 0);
    stamp_t = 0;
    treeroot.assign(m, -1);
    dead.assign(n, 0);
    dead_touched.clear();
  }

  // One search with persistent trees: every free vertex becomes the root of an alternating tree,
  // the forest grows along tight edges and the duals are adjusted when it is stuck.  An augmenting
  // path between two trees retires just those two trees (retire_trees): their vertices become
  // unlabelled, the rest of the forest stays, so no tree is ever grown twice.
  //
  // Dual adjustments are lazy: only the running sum Delta of all adjustments is kept; a vertex under
  // an S (T) top-level blossom has the dual  dual[x] - Delta  (+ Delta), an S (T) blossom the dual
  // dual[b] + Delta  (- Delta), and the stored values are converted whenever a label changes
  // (relabel).  The next adjustment is the smallest key of three heaps, each key being
  // "value + Delta at insertion", which stays exact for as long as the labels at both ends stay:
  // slack of an edge from an S-vertex to an unlabelled vertex, half the slack of an edge between two
  // S-blossoms, dual of a T-blossom.  Entries are never removed when labels change; an entry is
  // checked when it reaches the top, and whoever changes a label pushes the entries that are exact
  // from then on (scan of a new S-vertex, retire_trees, expansion of a T-blossom).
  // Returns true when the matching is perfect.
  bool solve() {
    allow.assign(n_edges(), 0);
    lab_touched.clear();
    allow_touched.clear();
    heap_free.clear();
    heap_ss.clear();
    heap_t.clear();
    queue.clear();
    Delta = 0;
    std::fill(minkey.begin(), minkey.end(), KEY_NONE);
    edge_queue.clear();
    n_free = 0;
    for (int v = 0; v < n; ++v)
      if (mate[v] < 0 && label[inblossom[v]] == 0) {
        assign_label(v, 1, -1);
        ++n_free;
      }
    if (n_free) ++n_stages;
    while (n_free > 0) {
      for (;;) {
        if (!edge_queue.empty()) {  // single edges that became tight
          const int v = edge_queue.back().first, k = edge_queue.back().second;
          edge_queue.pop_back();
          if (label[inblossom[v]] == 1) scan_edge(v, E[k].u == v ? 2 * k + 1 : 2 * k, E[k].u == v ? E[k].v : E[k].u, E[k].w);
#ifdef FPB_MWPM_CHECK
          check_forest("after edge event");
#endif
          continue;
        }
        if (queue.empty()) break;
        const int v = queue.back();  // a new S-vertex: all its edges
        queue.pop_back();
        if (label[inblossom[v]] != 1) continue;  // its tree was retired since it was queued
        for (int ai = adj_off[v]; ai < adj_off[v + 1]; ++ai) {
          const AdjRec& ar = adj[ai];  // the neighbour and the weight travel with the adjacency entry
          const bool retired_v = scan_edge(v, ar.p, ar.w, ar.wt);
#ifdef FPB_MWPM_CHECK
          check_forest("after scan_edge");
#endif
          if (retired_v) break;  // v is unlabelled now
        }
      }
      if (n_free <= 0) break;
      // the next event: the smallest exact key of the three heaps
      int deltatype = -1;
      ll key = 0;
      while (!heap_free.empty()) {
        const int k = heap_free.front().second;
        const int lu = label[inblossom[E[k].u]], lv = label[inblossom[E[k].v]];
        const ll kk = heap_free.front().first;
        if (((lu == 1 && lv == 0) || (lu == 0 && lv == 1)) && slack(k) + Delta == kk) break;
        heap_pop(heap_free);
        // a stale entry may have kept better-looking edges of its unlabelled end out of the heap
        if (lu == 0 && minkey[E[k].u] == kk) push_free_edges(E[k].u, false);
        if (lv == 0 && minkey[E[k].v] == kk) push_free_edges(E[k].v, false);
      }
      if (!edge_queue.empty()) continue;  // rebuilding a stale entry found edges that are tight already
      if (!heap_free.empty()) {
        deltatype = 2;
        key = heap_free.front().first;
      }
      while (!heap_ss.empty()) {
        const int k = heap_ss.front().second;
        const int bu = inblossom[E[k].u], bv = inblossom[E[k].v];
        if (bu != bv && label[bu] == 1 && label[bv] == 1 && slack(k) / 2 + Delta == heap_ss.front().first) break;
        heap_pop(heap_ss);
      }
      if (!heap_ss.empty() && (deltatype < 0 || heap_ss.front().first < key)) {
        deltatype = 3;
        key = heap_ss.front().first;
      }
      while (!heap_t.empty()) {
        const int b = heap_t.front().second;
        if (blossomparent[b] < 0 && blossombase[b] >= 0 && label[b] == 2 && dual[b] == heap_t.front().first) break;
        heap_pop(heap_t);
      }
      if (!heap_t.empty() && (deltatype < 0 || heap_t.front().first < key)) {
        deltatype = 4;
        key = heap_t.front().first;
      }
      if (deltatype < 0) break;  // the candidate graph has no perfect matching
      ++n_deltas;
#ifdef FPB_MWPM_CHECK
      if (key < Delta) check_failed("event key below Delta", "adjustment", deltatype, 0, key);
#endif
      if (key > Delta) Delta = key;
#ifdef FPB_MWPM_CHECK
      check_duals("after adjustment");
#endif
      if (deltatype == 2) {
        const int k = heap_free.front().second;
        heap_pop(heap_free);
        edge_queue.push_back(std::make_pair(label[inblossom[E[k].u]] == 1 ? E[k].u : E[k].v, k));
      } else if (deltatype == 3) {
        const int k = heap_ss.front().second;
        heap_pop(heap_ss);
        edge_queue.push_back(std::make_pair(E[k].u, k));
      } else {
        const int b = heap_t.front().second;
        heap_pop(heap_t);
        expand_blossom(b, false);
      }
    }
    // back to plain duals, then what the end of a stage does: S-blossoms with zero dual are expanded
    for (int x = 0; x < n; ++x) {
      dual[x] += voff[x] * Delta;
      voff[x] = 0;
    }
    for (int b = n; b < 2 * n; ++b) {
      dual[b] += boff[b] * Delta;
      boff[b] = 0;
    }
    Delta = 0;
    {
      const size_t n_lab = lab_touched.size();  // expansions at the end add no labels
      for (size_t q = 0; q < n_lab; ++q) {
        const int b = lab_touched[q];
        if (b >= n && blossomparent[b] < 0 && blossombase[b] >= 0 && label[b] == 1 && dual[b] == 0) expand_blossom(b, true);
      }
    }
    end_stage();
    for (int v = 0; v < n; ++v)
      if (mate[v] < 0) return false;
    return true;
  }

  // 2 x (sum of the duals of the blossoms that contain both u and v)
  ll common_blossom_dual(int u, int v) {
    if (inblossom[u] != inblossom[v] || inblossom[u] < n) return 0;
    ++stamp_t;
    for (int b = blossomparent[u]; b >= 0; b = blossomparent[b]) stamp[b] = stamp_t;
    ll s = 0;
    bool in = false;
    for (int b = blossomparent[v]; b >= 0; b = blossomparent[b]) {
      in = in || stamp[b] == stamp_t;
      if (in) s += dual[b];
    }
    return 2 * s;
  }

 private:
  struct AdjRec {
    int p;   // remote endpoint (2 * edge + side)
    int w;   // the neighbour
    ll wt;   // the edge weight
  };
  std::vector<int> adj_off, fill_tmp;
  std::vector<AdjRec> adj;
  std::vector<int> label, labelend, inblossom, blossomparent, blossombase;
  std::vector<std::vector<int>> childs, endps;
  std::vector<char> allow;
  std::vector<int> unused, queue, stamp;
  std::vector<std::pair<int, int>> edge_queue;  // (S-vertex, edge) pairs to look at
  int n_free = 0;
  int stamp_t = 0;

  // lazy dual adjustments (see solve()): sign of Delta in the dual of a vertex / of a blossom
  ll Delta = 0;
  std::vector<signed char> voff, boff;
  typedef std::pair<ll, int> HeapRec;  // (value + Delta at insertion, edge or blossom)
  std::vector<HeapRec> heap_free, heap_ss, heap_t;
  static constexpr ll KEY_NONE = INT64_MAX;
  std::vector<ll> minkey;  // per unlabelled vertex: key of its entry in heap_free
  static void heap_push(std::vector<HeapRec>& h, ll key, int id) {
    h.push_back(HeapRec(key, id));
    std::push_heap(h.begin(), h.end(), std::greater<HeapRec>());
  }
  static void heap_pop(std::vector<HeapRec>& h) {
    std::pop_heap(h.begin(), h.end(), std::greater<HeapRec>());
    h.pop_back();
  }

  std::vector<int> lab_touched, allow_touched;
  std::vector<int> treeroot;  // per labelled top-level blossom: the root vertex of its tree
  std::vector<char> dead;     // per root vertex: its tree was retired
  std::vector<int> dead_touched, retired, to_expand;

  static int wrap(int j, int len) { return j < 0 ? j + len : j; }

  // per-stage state is undone through these lists instead of clearing whole arrays
  void set_label(int i, int t) {
    lab_touched.push_back(i);
    label[i] = t;
  }
  void set_allow(int k) {
    if (!allow[k]) {
      allow[k] = 1;
      allow_touched.push_back(k);
    }
  }
  void end_stage() {
    for (size_t q = 0; q < lab_touched.size(); ++q) label[lab_touched[q]] = 0;
    for (size_t q = 0; q < allow_touched.size(); ++q) allow[allow_touched[q]] = 0;
    for (size_t q = 0; q < dead_touched.size(); ++q) dead[dead_touched[q]] = 0;
    lab_touched.clear();
    allow_touched.clear();
    dead_touched.clear();
  }

  // The stored duals of top-level blossom b follow its label: t = 0 unlabelled, 1 = S, 2 = T.
  void relabel(int b, int t) {
    const int sv = t == 1 ? -1 : t == 2 ? 1 : 0;
    leaves(b, [&](int x) {
      dual[x] += (voff[x] - sv) * Delta;
      voff[x] = (signed char)sv;
    });
    if (b >= n) {
      dual[b] += (boff[b] + sv) * Delta;
      boff[b] = (signed char)-sv;
    }
  }

  // One edge (v, w) seen from the S-vertex v; p is the endpoint number at w.  Returns true when it
  // closed an augmenting path (v's tree is retired then).
  bool scan_edge(int v, int p, int w, ll wt) {
    const int k = p >> 1;
    if (inblossom[v] == inblossom[w]) return false;
    ll kslack = 0;
    if (!allow[k]) {
      kslack = act(v) + act(w) - 2 * wt;
      if (kslack <= 0) set_allow(k);
    }
    const int lw = label[inblossom[w]];
    if (allow[k]) {
      if (lw == 0) {
        assign_label(w, 2, p ^ 1);
      } else if (lw == 1) {
        const int base = scan_blossom(v, w);
        if (base >= 0) {
          add_blossom(base, k);
        } else {
          const int r1 = treeroot[inblossom[v]], r2 = treeroot[inblossom[w]];
          augment_matching(k);
          retire_trees(r1, r2);
          ++n_augmentations;
          n_free -= 2;
          return true;
        }
      } else if (label[w] == 0) {
        set_label(w, 2);
        labelend[w] = p ^ 1;
      }
    } else if (lw == 1) {
      heap_push(heap_ss, kslack / 2 + Delta, k);
    } else if (lw == 0 && kslack + Delta < minkey[w]) {  // only w's best edge needs an entry
      minkey[w] = kslack + Delta;
      heap_push(heap_free, kslack + Delta, k);
    }
    return false;
  }

  // The two trees rooted at the (just matched) vertices r1 and r2 leave the forest.  What the end
  // of a stage does for the whole forest is done for their members only: labels are dropped, the
  // duals become plain numbers again, S-blossoms with zero dual are expanded.  Then every edge from
  // a retired vertex to a vertex that is still S is looked at again: a tight one re-queues the
  // S-vertex (it will pull the retired vertex into its tree), the others enter the heap of
  // S-to-unlabelled edges.  Allowed-edge flags of retired vertices are dropped, since such an edge
  // may lose its tightness once one end is labelled again.  The list of labelled entries is
  // compacted on the way.
  void retire_trees(int r1, int r2) {
    dead[r1] = dead[r2] = 1;
    dead_touched.push_back(r1);
    dead_touched.push_back(r2);
    retired.clear();
    to_expand.clear();
    size_t keep = 0;
    for (size_t q = 0; q < lab_touched.size(); ++q) {
      const int i = lab_touched[q];
      if (label[i] == 0) continue;
      if (i >= n && blossombase[i] < 0) {  // a blossom number that was freed
        label[i] = 0;
        continue;
      }
      int top = i < n ? inblossom[i] : i;
      while (blossomparent[top] >= 0) top = blossomparent[top];
      if (top == i ? !dead[treeroot[i]] : (label[top] != 0 && !dead[treeroot[top]])) {
        lab_touched[keep++] = i;
        continue;
      }
      if (top == i) {
        relabel(i, 0);
        leaves(i, [&](int x) { retired.push_back(x); });
        if (i >= n && label[i] == 1 && dual[i] == 0) to_expand.push_back(i);
      }
      label[i] = 0;
    }
    lab_touched.resize(keep);
    for (size_t q = 0; q < retired.size(); ++q) label[retired[q]] = 0;
    for (size_t q = 0; q < to_expand.size(); ++q) expand_blossom(to_expand[q], true);
    for (size_t q = 0; q < retired.size(); ++q) push_free_edges(retired[q], true);
  }

  // Edges from the unlabelled vertex w to S-vertices: a tight one is queued (its S-end will pull w
  // into a tree), and the one with the least slack enters the heap of S-to-unlabelled edges as w's
  // entry.  Called whenever w becomes unlabelled, and when w's entry turns out to be stale.
  void push_free_edges(int w, bool drop_flags) {
    ll best = KEY_NONE;
    int bk = -1;
    for (int ai = adj_off[w]; ai < adj_off[w + 1]; ++ai) {
      const AdjRec& ar = adj[ai];
      if (drop_flags) allow[ar.p >> 1] = 0;
      const int lx = label[inblossom[ar.w]];
      if (drop_flags && lx == 2 && inblossom[ar.w] != ar.w && label[ar.w] == 2 && endpoint(labelend[ar.w]) == w)
        label[ar.w] = 0;  // a vertex inside a T-blossom was marked as reached from w
      if (lx != 1) continue;
      const ll sl = act(w) + act(ar.w) - 2 * ar.wt;
      if (sl <= 0) {
        edge_queue.push_back(std::make_pair(ar.w, ar.p >> 1));
      } else if (sl + Delta < best) {
        best = sl + Delta;
        bk = ar.p >> 1;
      }
    }
    minkey[w] = best;
    if (bk >= 0) heap_push(heap_free, best, bk);
  }

  template <class F>
  void leaves(int b, F&& f) {
    if (b < n) {
      f(b);
    } else {
      for (size_t i = 0; i < childs[b].size(); ++i) leaves(childs[b][i], f);
    }
  }

#ifdef FPB_MWPM_CHECK
  // Invariant checks of the development build (g++ -DFPB_MWPM_CHECK, see tests/test_mwpm_native.py):
  // dual feasibility on every edge between different top-level blossoms (S-S slacks even, blossom
  // duals non-negative), and every labelled blossom hanging, through alternating labels of one
  // tree, under the free vertex its treeroot names.
  [[noreturn]] void check_failed(const char* what, const char* where, int a, int b, ll v) {
    fprintf(stderr, "fpb_mwpm check failed (%s, %s): %d %d %lld, Delta %lld\n", what, where, a, b, v, Delta);
    abort();
  }
  void check_duals(const char* where) {
    for (int k = 0; k < n_edges(); ++k) {
      const int bu = inblossom[E[k].u], bv = inblossom[E[k].v];
      if (bu == bv) continue;
      const ll sl = slack(k);
      if (sl < 0 || (label[bu] == 1 && label[bv] == 1 && (sl & 1))) check_failed("edge slack", where, E[k].u, E[k].v, sl);
    }
    for (int b = n; b < 2 * n; ++b)
      if (blossombase[b] >= 0 && dual[b] + boff[b] * Delta < 0) check_failed("blossom dual", where, b, 0, dual[b] + boff[b] * Delta);
  }
  void check_forest(const char* where) {
    for (int v = 0; v < n; ++v) {
      const int b = inblossom[v];
      if (label[b] != 1 && label[b] != 2) continue;
      int x = b, steps = 0;
      bool bad = false;
      while (labelend[x] >= 0 && steps++ < 2 * n) {
        const int nx = inblossom[endpoint(labelend[x])];
        bad = bad || label[nx] != 3 - label[x] || treeroot[nx] != treeroot[x];
        x = nx;
      }
      int rootv = -1;
      leaves(x, [&](int y) {
        if (mate[y] < 0) rootv = y;
      });
      if (bad || rootv < 0 || rootv != treeroot[b]) check_failed("tree of a labelled blossom", where, b, rootv, treeroot[b]);
    }
  }
#endif
  void assign_label(int w, int t, int p) {
    const int b = inblossom[w];
    set_label(w, t);
    set_label(b, t);
    labelend[w] = labelend[b] = p;
    treeroot[b] = p < 0 ? w : treeroot[inblossom[endpoint(p)]];  // the free vertex this tree grows from
    relabel(b, t);
    if (t == 1) {
      leaves(b, [&](int v) { queue.push_back(v); });
    } else {
      if (b >= n) heap_push(heap_t, dual[b], b);  // a T-blossom's stored dual is its key
      const int base = blossombase[b];
      assign_label(endpoint(mate[base]), 1, mate[base] ^ 1);
    }
  }

  // Trace back from v and w towards the roots; returns the base vertex of the first common
  // blossom (a new blossom closes) or -1 (the two trees differ: an augmenting path).
  int scan_blossom(int v, int w) {
    path.clear();
    int base = -1;
    while (v != -1 || w != -1) {
      int b = inblossom[v];
      if (label[b] & 4) {
        base = blossombase[b];
        break;
      }
      path.push_back(b);
      label[b] = 5;
      if (labelend[b] == -1) {
        v = -1;
      } else {
        v = endpoint(labelend[b]);
        b = inblossom[v];
        v = endpoint(labelend[b]);
      }
      if (w != -1) std::swap(v, w);
    }
    for (size_t i = 0; i < path.size(); ++i) label[path[i]] = 1;
    return base;
  }
  std::vector<int> path;

  void add_blossom(int base, int k) {
    int v = E[k].u, w = E[k].v;
    const int bb = inblossom[base];
    int bv = inblossom[v], bw = inblossom[w];
    const int b = unused.back();
    unused.pop_back();
    blossombase[b] = base;
    blossomparent[b] = -1;
    blossomparent[bb] = b;
    std::vector<int>& pth = childs[b];
    std::vector<int>& eps = endps[b];
    pth.clear();
    eps.clear();
    while (bv != bb) {
      blossomparent[bv] = b;
      pth.push_back(bv);
      eps.push_back(labelend[bv]);
      v = endpoint(labelend[bv]);
      bv = inblossom[v];
    }
    pth.push_back(bb);
    std::reverse(pth.begin(), pth.end());
    std::reverse(eps.begin(), eps.end());
    eps.push_back(2 * k);
    while (bw != bb) {
      blossomparent[bw] = b;
      pth.push_back(bw);
      eps.push_back(labelend[bw] ^ 1);
      w = endpoint(labelend[bw]);
      bw = inblossom[w];
    }
    set_label(b, 1);
    labelend[b] = labelend[bb];
    treeroot[b] = treeroot[bb];
    for (size_t i = 0; i < pth.size(); ++i) {  // the children's duals stop moving
      const int c = pth[i];
      if (c >= n) {
        dual[c] += boff[c] * Delta;
        boff[c] = 0;
      }
    }
    boff[b] = 1;
    dual[b] = -Delta;  // zero as of now
    leaves(b, [&](int x) {
      if (label[inblossom[x]] == 2) queue.push_back(x);
      if (voff[x] > 0) {  // was under a T-blossom
        dual[x] += 2 * Delta;
        voff[x] = -1;
      }
      inblossom[x] = b;
    });
  }

  void expand_blossom(int b, bool endstage) {
    const bool t_node = !endstage && label[b] == 2;
    if (t_node) relabel(b, 0);  // plain duals while the children are sorted out
    for (size_t i = 0; i < childs[b].size(); ++i) {
      const int s = childs[b][i];
      blossomparent[s] = -1;
      if (s < n) {
        inblossom[s] = s;
      } else if (endstage && dual[s] == 0) {
        expand_blossom(s, endstage);
      } else {
        leaves(s, [&](int x) { inblossom[x] = s; });
      }
    }
    if (t_node) {
      // the blossom was a T-node of a tree: relabel the even-length half of its cycle that
      // leads from the entry child to the base, and re-examine the rest
      const int len = (int)childs[b].size();
      const int entrychild = inblossom[endpoint(labelend[b] ^ 1)];
      int j = (int)(std::find(childs[b].begin(), childs[b].end(), entrychild) - childs[b].begin());
      int jstep, endptrick;
      if (j & 1) {
        j -= len; jstep = 1; endptrick = 0;
      } else {
        jstep = -1; endptrick = 1;
      }
      int p = labelend[b];
      while (j != 0) {
        label[endpoint(p ^ 1)] = 0;
        label[endpoint(endps[b][wrap(j - endptrick, len)] ^ endptrick ^ 1)] = 0;
        assign_label(endpoint(p ^ 1), 2, p);
        set_allow(endps[b][wrap(j - endptrick, len)] >> 1);
        j += jstep;
        p = endps[b][wrap(j - endptrick, len)] ^ endptrick;
        set_allow(p >> 1);
        j += jstep;
      }
      int bv = childs[b][wrap(j, len)];
      set_label(endpoint(p ^ 1), 2);
      set_label(bv, 2);
      labelend[endpoint(p ^ 1)] = labelend[bv] = p;
      treeroot[bv] = treeroot[b];
      relabel(bv, 2);
      if (bv >= n) heap_push(heap_t, dual[bv], bv);
      j += jstep;
      while (childs[b][wrap(j, len)] != entrychild) {
        bv = childs[b][wrap(j, len)];
        if (label[bv] == 1) {
          j += jstep;
          continue;
        }
        int found = -1;
        leaves(bv, [&](int x) {
          if (found < 0 && label[x] != 0) found = x;
        });
        if (found >= 0) {
          label[found] = 0;
          label[endpoint(mate[blossombase[bv]])] = 0;
          assign_label(found, 2, labelend[found]);
        }
        j += jstep;
      }
      // children that stay unlabelled: their edges to S-vertices count for the next adjustment
      for (size_t i = 0; i < childs[b].size(); ++i)
        if (label[childs[b][i]] == 0) leaves(childs[b][i], [&](int x) { push_free_edges(x, false); });
    }
    set_label(b, -1);
    labelend[b] = -1;
    childs[b].clear();
    endps[b].clear();
    blossombase[b] = -1;
    unused.push_back(b);
  }

  // Swap matched / unmatched edges over the alternating path through blossom b between
  // vertex v and the base, making v the new base.
  void augment_blossom(int b, int v) {
    int t = v;
    while (blossomparent[t] != b) t = blossomparent[t];
    if (t >= n) augment_blossom(t, v);
    const int len = (int)childs[b].size();
    const int i = (int)(std::find(childs[b].begin(), childs[b].end(), t) - childs[b].begin());
    int j = i, jstep, endptrick;
    if (i & 1) {
      j -= len; jstep = 1; endptrick = 0;
    } else {
      jstep = -1; endptrick = 1;
    }
    while (j != 0) {
      j += jstep;
      t = childs[b][wrap(j, len)];
      const int p = endps[b][wrap(j - endptrick, len)] ^ endptrick;
      if (t >= n) augment_blossom(t, endpoint(p));
      j += jstep;
      t = childs[b][wrap(j, len)];
      if (t >= n) augment_blossom(t, endpoint(p ^ 1));
      mate[endpoint(p)] = p ^ 1;
      mate[endpoint(p ^ 1)] = p;
    }
    std::rotate(childs[b].begin(), childs[b].begin() + i, childs[b].end());
    std::rotate(endps[b].begin(), endps[b].begin() + i, endps[b].end());
    blossombase[b] = blossombase[childs[b][0]];
  }

  void augment_matching(int k) {
    for (int side = 0; side < 2; ++side) {
      int s = side ? E[k].v : E[k].u;
      int p = side ? 2 * k : 2 * k + 1;
      for (;;) {
        const int bs = inblossom[s];
        if (bs >= n) augment_blossom(bs, s);
        mate[s] = p;
        if (labelend[bs] == -1) break;
        const int t = endpoint(labelend[bs]);
        const int bt = inblossom[t];
        s = endpoint(labelend[bt]);
        const int j = endpoint(labelend[bt] ^ 1);
        if (bt >= n) augment_blossom(bt, j);
        mate[j] = labelend[bt];
        p = labelend[bt] ^ 1;
      }
    }
  }
};

}  // namespace fpb_sparse
