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
// minimum-weight perfect matching of the complete graph - the same optimum the dense solver in
// mwpm.cpp finds, only the choice among ties may differ.
//
// The solver is the classical primal-dual formulation with adjacency lists (labels S/T on
// top-level blossoms, least-slack edge tracking per blossom, dual adjustment by the minimum of
// the free-vertex / S-S / T-blossom slacks, blossom expansion during and at the end of a stage),
// one augmentation per stage with every free vertex as a root, on weights -2*len so that vertex
// duals stay integral and S-S slacks stay even.
#pragma once
#include <algorithm>
#include <cstdint>
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
  long long n_stages = 0, n_deltas = 0;

  void reset(int n_) {
    n = n_;
    E.clear();
    n_stages = n_deltas = 0;
  }
  int n_edges() const { return (int)E.size(); }
  void add_edge(int u, int v, ll w) { E.push_back(EdgeRec{u, v, w}); }
  int endpoint(int p) const { return (p & 1) ? E[p >> 1].v : E[p >> 1].u; }
  ll slack(int k) const {
    const EdgeRec& e = E[k];
    return dual[e.u] + dual[e.v] - 2 * e.w;
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
    bestedge.assign(m, -1);
    if ((int)childs.size() < m) {
      childs.resize(m);
      endps.resize(m);
      bestedges.resize(m);
    }
    for (int b = 0; b < m; ++b) {
      childs[b].clear();
      endps[b].clear();
      bestedges[b].clear();
    }
    has_bestedges.assign(m, 0);
    unused.clear();
    for (int b = m - 1; b >= n; --b) unused.push_back(b);
    bestedgeto.assign(m, -1);
    stamp.assign(m, 0);
    stamp_t = 0;
  }

  // Stages: label every free vertex S, grow the alternating forest along tight edges, adjust the
  // duals when it is stuck, until one augmenting path is found; repeat until none exists.
  // Labels, least-slack records and allowed-edge flags are tracked in lists, so that a dual
  // adjustment and the clean-up of a stage cost the size of the forest, not of the graph.
  // Returns true when the matching is perfect.
  bool solve() {
    allow.assign(n_edges(), 0);
    upd_stamp.assign(2 * (size_t)n, 0);
    upd_t = 0;
    lab_touched.clear();
    be_touched.clear();
    allow_touched.clear();
    for (;;) {
      queue.clear();
      for (int v = 0; v < n; ++v)
        if (mate[v] < 0 && label[inblossom[v]] == 0) assign_label(v, 1, -1);
      if (queue.empty()) break;
      ++n_stages;
      bool augmented = false;
      for (;;) {
        while (!queue.empty() && !augmented) {
          const int v = queue.back();
          queue.pop_back();
          for (int ai = adj_off[v]; ai < adj_off[v + 1]; ++ai) {
            const AdjRec& ar = adj[ai];  // the neighbour and the weight travel with the adjacency entry
            const int p = ar.p, k = p >> 1, w = ar.w;
            if (inblossom[v] == inblossom[w]) continue;
            ll kslack = 0;
            if (!allow[k]) {
              kslack = dual[v] + dual[w] - 2 * ar.wt;
              if (kslack <= 0) set_allow(k);
            }
            if (allow[k]) {
              if (label[inblossom[w]] == 0) {
                assign_label(w, 2, p ^ 1);
              } else if (label[inblossom[w]] == 1) {
                const int base = scan_blossom(v, w);
                if (base >= 0) {
                  add_blossom(base, k);
                } else {
                  augment_matching(k);
                  augmented = true;
                  break;
                }
              } else if (label[w] == 0) {
                set_label(w, 2);
                labelend[w] = p ^ 1;
              }
            } else if (label[inblossom[w]] == 1) {
              const int b = inblossom[v];
              if (bestedge[b] < 0 || kslack < slack(bestedge[b])) set_bestedge(b, k);
            } else if (label[w] == 0) {
              if (bestedge[w] < 0 || kslack < slack(bestedge[w])) set_bestedge(w, k);
            }
          }
        }
        if (augmented) break;
        int deltatype = -1, deltaedge = -1, deltablossom = -1;
        ll delta = 0;
        for (size_t q = 0; q < be_touched.size(); ++q) {
          const int i = be_touched[q], k = bestedge[i];
          if (k < 0) continue;
          if (i < n && label[inblossom[i]] == 0) {
            const ll d = slack(k);
            if (deltatype < 0 || d < delta) {
              delta = d; deltatype = 2; deltaedge = k;
            }
          } else if (blossomparent[i] < 0 && label[i] == 1) {
            const ll d = slack(k) / 2;
            if (deltatype < 0 || d < delta) {
              delta = d; deltatype = 3; deltaedge = k;
            }
          }
        }
        for (size_t q = 0; q < lab_touched.size(); ++q) {
          const int b = lab_touched[q];
          if (b >= n && blossombase[b] >= 0 && blossomparent[b] < 0 && label[b] == 2 && (deltatype < 0 || dual[b] < delta)) {
            delta = dual[b]; deltatype = 4; deltablossom = b;
          }
        }
        if (deltatype < 0) break;  // the candidate graph has no perfect matching
        ++n_deltas;
        ++upd_t;
        for (size_t q = 0; q < lab_touched.size(); ++q) {
          const int i = lab_touched[q];
          if (blossomparent[i] >= 0 || (i >= n && blossombase[i] < 0)) continue;  // not top-level
          const int l = label[i];
          if ((l != 1 && l != 2) || upd_stamp[i] == upd_t) continue;
          upd_stamp[i] = upd_t;
          const ll dv = l == 1 ? -delta : delta;
          leaves(i, [&](int x) { dual[x] += dv; });
          if (i >= n) dual[i] -= dv;
        }
        if (deltatype == 2) {
          set_allow(deltaedge);
          int i = E[deltaedge].u;
          if (label[inblossom[i]] == 0) i = E[deltaedge].v;
          queue.push_back(i);
        } else if (deltatype == 3) {
          set_allow(deltaedge);
          queue.push_back(E[deltaedge].u);
        } else {
          expand_blossom(deltablossom, false);
        }
      }
      if (augmented) {
        const size_t n_lab = lab_touched.size();  // expansions at the end of a stage add no labels
        for (size_t q = 0; q < n_lab; ++q) {
          const int b = lab_touched[q];
          if (b >= n && blossomparent[b] < 0 && blossombase[b] >= 0 && label[b] == 1 && dual[b] == 0) expand_blossom(b, true);
        }
      }
      end_stage();
      if (!augmented) break;
    }
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
  std::vector<int> label, labelend, inblossom, blossomparent, blossombase, bestedge;
  std::vector<std::vector<int>> childs, endps, bestedges;
  std::vector<char> has_bestedges, allow;
  std::vector<int> unused, queue, bestedgeto, touched, stamp;
  int stamp_t = 0;

  std::vector<int> lab_touched, be_touched, allow_touched, upd_stamp;
  int upd_t = 0;

  static int wrap(int j, int len) { return j < 0 ? j + len : j; }

  // per-stage state is undone through these lists instead of clearing whole arrays
  void set_label(int i, int t) {
    lab_touched.push_back(i);
    label[i] = t;
  }
  void set_bestedge(int i, int k) {
    if (bestedge[i] < 0) be_touched.push_back(i);
    bestedge[i] = k;
  }
  void set_allow(int k) {
    if (!allow[k]) {
      allow[k] = 1;
      allow_touched.push_back(k);
    }
  }
  void end_stage() {
    for (size_t q = 0; q < lab_touched.size(); ++q) {
      const int i = lab_touched[q];
      label[i] = 0;
      if (i >= n) {
        has_bestedges[i] = 0;
        bestedges[i].clear();
      }
    }
    for (size_t q = 0; q < be_touched.size(); ++q) bestedge[be_touched[q]] = -1;
    for (size_t q = 0; q < allow_touched.size(); ++q) allow[allow_touched[q]] = 0;
    lab_touched.clear();
    be_touched.clear();
    allow_touched.clear();
  }

  template <class F>
  void leaves(int b, F&& f) {
    if (b < n) {
      f(b);
    } else {
      for (size_t i = 0; i < childs[b].size(); ++i) leaves(childs[b][i], f);
    }
  }

  void assign_label(int w, int t, int p) {
    const int b = inblossom[w];
    set_label(w, t);
    set_label(b, t);
    labelend[w] = labelend[b] = p;
    bestedge[w] = bestedge[b] = -1;
    if (t == 1) {
      leaves(b, [&](int v) { queue.push_back(v); });
    } else {
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
    dual[b] = 0;
    leaves(b, [&](int x) {
      if (label[inblossom[x]] == 2) queue.push_back(x);
      inblossom[x] = b;
    });
    // least-slack edges from the new blossom to every neighbouring S-blossom
    touched.clear();
    auto consider = [&](int e) {
      int j = E[e].v;
      if (inblossom[j] == b) j = E[e].u;
      const int bj = inblossom[j];
      if (bj != b && label[bj] == 1) {
        if (bestedgeto[bj] < 0) {
          bestedgeto[bj] = e;
          touched.push_back(bj);
        } else if (slack(e) < slack(bestedgeto[bj])) {
          bestedgeto[bj] = e;
        }
      }
    };
    for (size_t i = 0; i < pth.size(); ++i) {
      const int c = pth[i];
      if (c >= n && has_bestedges[c]) {
        for (size_t q = 0; q < bestedges[c].size(); ++q) consider(bestedges[c][q]);
      } else {
        leaves(c, [&](int x) {
          for (int ai = adj_off[x]; ai < adj_off[x + 1]; ++ai) consider(adj[ai].p >> 1);
        });
      }
      if (c >= n) {
        has_bestedges[c] = 0;
        bestedges[c].clear();
      }
      bestedge[c] = -1;
    }
    bestedges[b].clear();
    bestedge[b] = -1;
    for (size_t i = 0; i < touched.size(); ++i) {
      const int e = bestedgeto[touched[i]];
      bestedgeto[touched[i]] = -1;
      bestedges[b].push_back(e);
      if (bestedge[b] < 0 || slack(e) < slack(bestedge[b])) set_bestedge(b, e);
    }
    has_bestedges[b] = 1;
  }

  void expand_blossom(int b, bool endstage) {
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
    if (!endstage && label[b] == 2) {
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
      bestedge[bv] = -1;
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
    }
    set_label(b, -1);
    labelend[b] = -1;
    childs[b].clear();
    endps[b].clear();
    blossombase[b] = -1;
    has_bestedges[b] = 0;
    bestedges[b].clear();
    bestedge[b] = -1;
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
