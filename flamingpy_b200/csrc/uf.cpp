// Host-side Union-Find + peeling decoder for a batch of trials (C ABI: include/fpb_uf.h).
//
// What it computes follows the reference's decoder (flamingpy/decoders/unionfind/algos.py and
// uf_classes.py; Delfosse-Nickerson, arXiv:1709.06218, with erasures as in arXiv:1703.01517):
//
//  1. Clusters start as the connected components of the erased edges; every other node is a cluster
//     of its own (algos.py:58-129).  A cluster's parity is the XOR of its cubes' parities; a cluster
//     holding a boundary point counts as even (uf_classes.py:92-98, algos.py:186-191).
//  2. While odd clusters remain, every node of every odd cluster grows each of its incident edges by
//     half an edge (Support.grow, uf_classes.py:134-142: empty -> half-grown -> grown; an edge between
//     two growing nodes gets both halves in one round); edges that became fully grown merge the
//     clusters at their ends (algos.py:132-193).  The set of grown edges after each round does not
//     depend on the order in which clusters, nodes and edges are visited, so it is reproduced exactly
//     (tests/golden/uf_*.npz hold the reference's own Support tables).
//  3. A spanning forest of the grown edges is peeled from the leaves (algos.py:196-311): an edge enters
//     the recovery when the subtree it cuts off holds an odd number of unsatisfied cubes; in a tree
//     with an odd number of them one boundary point absorbs the excess (algos.py:207-225).
//     The reference takes rustworkx's minimum_spanning_tree and the boundary point its set iteration
//     meets last; here the forest is a breadth-first one rooted at the lowest-numbered boundary point.
//     Both choices give recoveries that clear the same syndrome and differ by cycles of grown edges.
//
// Written for this data layout: a CSR adjacency over cubes + boundary points shared by all trials,
// per-trial state in flat arrays reused across trials (one workspace per OpenMP thread), cluster
// node lists as intrusive singly linked lists so that a merge is O(1), pruning of fully grown nodes
// done while a list is walked.  A trial touches only its own footprint (odd cubes, erased edges and
// whatever the clusters grow over) and restores it afterwards, so a quiet trial costs next to nothing.
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/fpb_uf.h"

namespace {

struct Graph {
  int n_nodes = 0, n_cubes = 0, n_edges = 0;
  std::vector<int> off, adj_edge, adj_node;
  const int32_t *ea = nullptr, *eb = nullptr;
};

struct Workspace {
  std::vector<int> parent, csize, head, tail, next, odd, odd_next, fusion, order, up_edge, up_node, stamp;
  std::vector<int> touched_nodes, touched_edges;
  std::vector<uint8_t> cpar, cbnd, support, npar, seen;
  std::vector<uint64_t> touched_bits;  // one bit per node: in touched_nodes
  int stamp_t = 0;

  // The per-node / per-edge state is kept in its "nothing happened" form between trials (every node a
  // cluster of its own, every edge empty); a trial records what it touches and puts exactly that back,
  // so that its cost follows the size of its syndrome instead of the size of the lattice.
  void resize(const Graph& g) {
    const int n = g.n_nodes;
    parent.resize(n), csize.assign(n, 1), head.resize(n), tail.resize(n);
    next.assign(n, -1), up_edge.resize(n), up_node.resize(n), stamp.assign(n, 0);
    cpar.assign(n, 0), cbnd.resize(n), npar.assign(n, 0), seen.assign(n, 0), touched_bits.assign((n + 63) / 64, 0);
    for (int u = 0; u < n; ++u) {
      parent[u] = head[u] = tail[u] = u;
      cbnd[u] = u >= g.n_cubes;
    }
    support.assign(g.n_edges, 0);
    stamp_t = 0;
  }
  void touch(int u) {
    uint64_t& wd = touched_bits[u >> 6];
    const uint64_t bit = 1ull << (u & 63);
    if (!(wd & bit)) {
      wd |= bit;
      touched_nodes.push_back(u);
    }
  }
  void restore(const Graph& g) {
    for (size_t i = 0; i < touched_nodes.size(); ++i) {
      const int u = touched_nodes[i];
      parent[u] = head[u] = tail[u] = u;
      csize[u] = 1;
      next[u] = -1;
      cpar[u] = npar[u] = seen[u] = 0;
      touched_bits[u >> 6] = 0;
      cbnd[u] = u >= g.n_cubes;
      stamp[u] = 0;
    }
    for (size_t i = 0; i < touched_edges.size(); ++i) support[touched_edges[i]] = 0;
    touched_nodes.clear();
    touched_edges.clear();
    stamp_t = 0;
  }
  int find(int u) {
    int r = u;
    while (parent[r] != r) r = parent[r];
    while (parent[u] != r) {
      const int nx = parent[u];
      parent[u] = r;
      u = nx;
    }
    return r;
  }
  void unite(int a, int b) {  // roots
    if (csize[a] < csize[b]) {
      const int t = a;
      a = b;
      b = t;
    }
    parent[b] = a;
    csize[a] += csize[b];
    cpar[a] ^= cpar[b];
    cbnd[a] |= cbnd[b];
    if (head[b] >= 0) {  // append b's node list
      if (head[a] < 0) {
        head[a] = head[b];
      } else {
        next[tail[a]] = head[b];
      }
      tail[a] = tail[b];
    }
  }
};

// Returns 0, or -1 when an odd cluster is stuck.
int decode_one(const Graph& g, Workspace& w, const uint8_t* parity, const uint8_t* erased, uint8_t* flips, uint8_t* grown) {
  std::memset(flips, 0, (size_t)g.n_edges);
  if (grown) std::memset(grown, 0, (size_t)g.n_edges);
  // The trial's footprint: every edge whose support leaves 0 (touched_edges), and as nodes the odd cubes
  // and both ends of every fully grown edge - only those take part in merges, lists and the forest.
  for (int u = 0; u < g.n_cubes; ++u)
    if (parity[u] & 1) {
      w.cpar[u] = 1;
      w.touch(u);
    }
  if (erased)
    for (int q = 0; q < g.n_edges; ++q) {
      if (g.ea[q] < 0 || !erased[q]) continue;
      w.support[q] = 2;
      w.touched_edges.push_back(q);
      w.touch(g.ea[q]);
      w.touch(g.eb[q]);
      const int a = w.find(g.ea[q]), b = w.find(g.eb[q]);
      if (a != b) w.unite(a, b);
    }
  // the first round grows every cluster of odd parity, also one that an erasure already ties to a
  // boundary point (algos.py:124-126 fills the list by parity alone; the boundary test comes after
  // the first growth, algos.py:186-191)
  w.odd.clear();
  for (size_t i = 0; i < w.touched_nodes.size(); ++i) {
    const int u = w.touched_nodes[i];
    if (w.parent[u] == u && w.cpar[u]) w.odd.push_back(u);
  }

  int rc = 0;
  while (!w.odd.empty()) {
    w.fusion.clear();
    long changed = 0;
    for (size_t i = 0; i < w.odd.size(); ++i) {
      const int r = w.odd[i];
      int prev = -1, u = w.head[r];
      while (u >= 0) {
        bool open = false;  // does u keep an edge that is not fully grown?
        for (int ai = g.off[u]; ai < g.off[u + 1]; ++ai) {
          const int q = g.adj_edge[ai];
          uint8_t& s = w.support[q];
          if (s == 0) {
            s = 1;
            w.touched_edges.push_back(q);
            open = true;
            ++changed;
          } else if (s == 1) {
            s = 2;
            w.fusion.push_back(q);
            ++changed;
          }
        }
        const int nx = w.next[u];
        if (!open) {  // prune: nothing left to grow from u
          if (prev < 0) {
            w.head[r] = nx;
          } else {
            w.next[prev] = nx;
          }
          if (w.tail[r] == u) w.tail[r] = prev;
        } else {
          prev = u;
        }
        u = nx;
      }
    }
    for (size_t i = 0; i < w.fusion.size(); ++i) {
      const int q = w.fusion[i];
      w.touch(g.ea[q]);
      w.touch(g.eb[q]);
      const int a = w.find(g.ea[q]), b = w.find(g.eb[q]);
      if (a != b) w.unite(a, b);
    }
    ++w.stamp_t;
    w.odd_next.clear();
    for (size_t i = 0; i < w.odd.size(); ++i) {
      const int r = w.find(w.odd[i]);
      if (w.stamp[r] == w.stamp_t) continue;
      w.stamp[r] = w.stamp_t;
      if (w.cpar[r] && !w.cbnd[r]) w.odd_next.push_back(r);
    }
    w.odd.swap(w.odd_next);
    // nothing grew and an odd cluster without a boundary point remains: it never will (the reference loops
    // forever here).  An odd cluster of the first round that already holds a boundary point just drops out,
    // also when every edge around it is erased (p_swap = 1) and there is nothing to grow.
    if (!changed && !w.odd.empty()) {
      rc = -1;
      break;
    }
  }
  if (rc) {  // stuck: no recovery for this trial
    w.restore(g);
    return rc;
  }

  if (grown)
    for (size_t i = 0; i < w.touched_edges.size(); ++i) grown[w.touched_edges[i]] = w.support[w.touched_edges[i]] == 2;

  // spanning forest of the grown edges, breadth first, boundary points first as roots (in ascending
  // order, then the cubes in ascending order); then peeling.  Nodes outside the trial's footprint have
  // no grown edge and even parity: trees of one node that contribute nothing.
  for (size_t i = 0; i < w.touched_nodes.size(); ++i) {
    const int u = w.touched_nodes[i];
    w.npar[u] = u < g.n_cubes ? (parity[u] & 1) : 0;
  }
  w.order.clear();
  // one tree from `root`: breadth first over the grown edges, then peeled from the leaves
  auto tree = [&](int root) {
    const size_t first = w.order.size();
    w.seen[root] = 1;
    w.up_edge[root] = -1;
    w.order.push_back(root);
    for (size_t i = first; i < w.order.size(); ++i) {
      const int u = w.order[i];
      for (int ai = g.off[u]; ai < g.off[u + 1]; ++ai) {
        const int q = g.adj_edge[ai], v = g.adj_node[ai];
        if (w.support[q] != 2 || w.seen[v]) continue;
        w.seen[v] = 1;
        w.up_edge[v] = q;
        w.up_node[v] = u;
        w.order.push_back(v);
      }
    }
    for (size_t i = w.order.size(); i-- > first + 1;) {
      const int u = w.order[i];
      if (w.npar[u]) {
        flips[w.up_edge[u]] ^= 1;
        w.npar[w.up_node[u]] ^= 1;
      }
    }
    if (root < g.n_cubes && w.npar[root]) rc = -1;  // an odd tree without a boundary point
  };
  for (int pass = 0; pass < 2; ++pass) {
    const int lo = pass == 0 ? g.n_cubes : 0, hi = pass == 0 ? g.n_nodes : g.n_cubes;
    if (lo >= hi) continue;
    for (int wi = lo >> 6; wi <= (hi - 1) >> 6; ++wi)
      for (uint64_t m = w.touched_bits[wi]; m; m &= m - 1) {  // touched nodes in ascending order
        const int root = (wi << 6) + __builtin_ctzll(m);
        if (root >= lo && root < hi && !w.seen[root]) tree(root);
      }
  }
  w.restore(g);
  return rc;
}

}  // namespace

extern "C" int fpb_uf_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

extern "C" long long fpb_uf_erasures(int n_lattice_nodes, int n_edges, const int32_t* qubit_node, const int32_t* adj_indptr,
                                     const int32_t* adj_indices, int n_trials, const uint8_t* state, uint8_t* erased,
                                     int threads) {
  long long total = 0;
#ifdef _OPENMP
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) reduction(+ : total) schedule(static)
#endif
  for (int t = 0; t < n_trials; ++t) {
    const uint8_t* st = state + (size_t)t * n_lattice_nodes;
    uint8_t* er = erased + (size_t)t * n_edges;
    for (int q = 0; q < n_edges; ++q) {
      const int v = qubit_node[q];
      int p_count = 0;
      for (int i = adj_indptr[v]; i < adj_indptr[v + 1]; ++i) p_count += st[adj_indices[i]] == 2;
      er[q] = p_count >= 2;
      total += er[q];
    }
  }
  (void)threads;
  return total;
}

extern "C" long long fpb_uf_check_correction(int n_trials, int n_qubits, const uint8_t* bits, const uint32_t* bit_words,
                                             const uint8_t* flips, const uint32_t* flip_words, int n_planes,
                                             const int32_t* plane_off, const int32_t* plane_qubits, uint8_t* fail,
                                             int threads) {
  const int words = (n_qubits + 31) / 32;
  long long total = 0;
#ifdef _OPENMP
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) reduction(+ : total) schedule(static)
#endif
  for (int t = 0; t < n_trials; ++t) {
    const uint8_t* b8 = bits ? bits + (size_t)t * n_qubits : nullptr;
    const uint32_t* bw = bit_words ? bit_words + (size_t)t * words : nullptr;
    const uint8_t* f8 = flips ? flips + (size_t)t * n_qubits : nullptr;
    const uint32_t* fw = flip_words ? flip_words + (size_t)t * words : nullptr;
    uint8_t failed = 0;
    for (int pl = 0; pl < n_planes; ++pl) {
      unsigned par = 0;
      for (int i = plane_off[pl]; i < plane_off[pl + 1]; ++i) {
        const int g = plane_qubits[i];
        const unsigned b = b8 ? b8[g] : (bw[g >> 5] >> (g & 31)) & 1u;
        const unsigned f = f8 ? f8[g] : (fw[g >> 5] >> (g & 31)) & 1u;
        par ^= b ^ f;
      }
      failed |= (uint8_t)(par & 1u);
    }
    fail[t] = failed;
    total += failed;
  }
  (void)threads;
  return total;
}

extern "C" int fpb_uf_batch(int n_nodes, int n_cubes, int n_edges, const int32_t* edge_a, const int32_t* edge_b,
                            int n_trials, const uint8_t* parity, const uint8_t* erased, uint8_t* flips, uint8_t* grown,
                            int threads) {
  Graph g;
  g.n_nodes = n_nodes, g.n_cubes = n_cubes, g.n_edges = n_edges, g.ea = edge_a, g.eb = edge_b;
  g.off.assign(n_nodes + 1, 0);
  for (int q = 0; q < n_edges; ++q) {
    if (edge_a[q] < 0) continue;
    ++g.off[edge_a[q] + 1];
    ++g.off[edge_b[q] + 1];
  }
  for (int u = 0; u < n_nodes; ++u) g.off[u + 1] += g.off[u];
  g.adj_edge.resize(g.off[n_nodes]);
  g.adj_node.resize(g.off[n_nodes]);
  {
    std::vector<int> fill(g.off.begin(), g.off.end() - 1);
    for (int q = 0; q < n_edges; ++q) {
      if (edge_a[q] < 0) continue;
      g.adj_edge[fill[edge_a[q]]] = q, g.adj_node[fill[edge_a[q]]++] = edge_b[q];
      g.adj_edge[fill[edge_b[q]]] = q, g.adj_node[fill[edge_b[q]]++] = edge_a[q];
    }
  }
  int first_bad = -1;
#ifdef _OPENMP
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel num_threads(nt)
#endif
  {
    Workspace w;
    w.resize(g);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (int t = 0; t < n_trials; ++t) {
      const int rc = decode_one(g, w, parity + (size_t)t * n_cubes, erased ? erased + (size_t)t * n_edges : nullptr,
                                flips + (size_t)t * n_edges, grown ? grown + (size_t)t * n_edges : nullptr);
      if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
        if (first_bad < 0 || t < first_bad) first_bad = t;
      }
    }
  }
  (void)threads;
  return first_bad < 0 ? 0 : -(1 + first_bad);
}
