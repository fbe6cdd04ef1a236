// Native minimum-weight perfect matching for the matching graphs the CUDA pipeline builds.
//
// Replaces the reference's only native component, the LEMON wrapper
// flamingpy/cpp/lemonpy.cpp:13-41 (and rx.max_weight_matching, decoders/mwpm/matching.py:286-292),
// neither of which can be installed offline.  Host-side C++, C ABI, OpenMP over trials.
//
// Algorithm: Edmonds' blossom algorithm in its primal-dual form on adjacency lists
// (mwpm_sparse.h: persistent alternating trees, lazy dual adjustments, event heaps).  Lengths are
// scaled to 64-bit integers (2^-32 resolution) and enter as weights -2 * length, so that all dual
// arithmetic is exact.  Small graphs hand the solver every edge of the complete graph
// (solve_one); from 48 cubes on it runs on a sparse candidate graph and the result is proved
// optimal for the complete graph by pricing (solve_one_sparse, the fpb_mwpm_sparse_* batch form).
//
// Matching-graph structure used (decoders/mwpm/matching.py:97-173): K odd cubes, pairwise
// lengths d_ab; with boundary points one private virtual partner per cube at length b_a and a
// zero-weight clique among virtual nodes.  That graph is equivalent to a complete graph on the K
// real nodes with weights min(d_ab, b_a + b_b) (plus one dummy node at weight b_a when K is
// odd): a pair whose minimum is b_a + b_b means "both go to their connector".  This halves n.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <queue>
#include <vector>

#include "mwpm_sparse.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef long long ll;

// non-negative scaled length -> integer, round half up (one cvttsd2si instead of a libm call)
inline ll to_fixed(double v) { return (ll)(v + 0.5); }

// Greedy start of the primal-dual solver (all drivers below): feasible duals y_u = best incident
// weight; then every free vertex lowers its dual until its first edge becomes tight and takes that
// partner if it is still free; then augmenting paths of three tight edges (free - matched =
// matched - free), which need no search.  All weights are even, so all duals stay even.  False
// when some vertex has no edge at all.
bool greedy_start(fpb_sparse::SparseBlossom& sb, int n, std::vector<ll>& y, std::vector<int>& vmate) {
  y.assign(n, INT64_MIN);
  vmate.assign(n, -1);
  for (int k = 0; k < sb.n_edges(); ++k) {
    y[sb.E[k].u] = std::max(y[sb.E[k].u], sb.E[k].w);
    y[sb.E[k].v] = std::max(y[sb.E[k].v], sb.E[k].w);
  }
  for (int u = 0; u < n; ++u)
    if (y[u] == INT64_MIN) return false;
  for (int u = 0; u < n; ++u) {
    if (vmate[u] >= 0) continue;
    ll smin = INT64_MAX;
    int pick = -1;
    sb.for_neighbours(u, [&](int v, ll w) {
      const ll sl = y[u] + y[v] - 2 * w;
      if (sl < smin || (sl == smin && pick >= 0 && vmate[pick] >= 0 && vmate[v] < 0)) {
        smin = sl;
        pick = v;
      }
    });
    y[u] -= smin;
    if (vmate[pick] < 0) {
      vmate[u] = pick;
      vmate[pick] = u;
    }
  }
  for (int u = 0; u < n; ++u) {
    if (vmate[u] >= 0) continue;
    bool found = false;
    sb.for_neighbours(u, [&](int v, ll w) {
      if (found || vmate[v] < 0 || y[u] + y[v] != 2 * w) return;
      const int v2 = vmate[v];
      int x = -1;
      sb.for_neighbours(v2, [&](int c, ll w2) {
        if (x < 0 && c != u && vmate[c] < 0 && y[v2] + y[c] == 2 * w2) x = c;
      });
      if (x >= 0) {
        vmate[u] = v; vmate[v] = u;
        vmate[v2] = x; vmate[x] = v2;
        found = true;
      }
    });
  }
  sb.start(y, vmate);
  return true;
}

// One matching graph, every edge of the reduced complete graph handed to the solver (no candidate
// selection, no pricing): the small-graph path and the last resort of the sparse drivers.
// pair_len: packed upper triangle (row-major, a < b); bnd_len: K lengths or null.  Writes out_src /
// out_dst (dst = -1: matched to its connector); returns the number of queries written (<= K) or a
// negative code.
int solve_one(int K, const double* pair_len, const double* bnd_len, int* out_src, int* out_dst) {
  if (K <= 0) return 0;
  if (!bnd_len && (K & 1)) return -1;  // no perfect matching
  if (K == 1) {
    out_src[0] = 0;
    out_dst[0] = -1;
    return 1;
  }
  const int n = (bnd_len && (K & 1)) ? K + 1 : K;
  // scale to ~2^-32 resolution; the solver's weights are -2 * length, so S-S slacks stay even
  double wmax = 0;
  const long long npairs = (long long)K * (K - 1) / 2;
  for (long long i = 0; i < npairs; ++i) wmax = std::max(wmax, pair_len[i]);
  if (bnd_len)
    for (int a = 0; a < K; ++a) wmax = std::max(wmax, 2 * bnd_len[a]);
  if (!(wmax < 1e15)) return -2;
  int shift = 32;
  while (shift > 0 && std::ldexp(wmax + 1.0, shift) * (double)(n + 2) > 1e17) --shift;
  const double scale = std::ldexp(1.0, shift);
  static thread_local fpb_sparse::SparseBlossom sb_tls;
  static thread_local std::vector<ll> y_tls;
  static thread_local std::vector<int> vmate_tls;
  fpb_sparse::SparseBlossom& sb = sb_tls;
  sb.reset(n);
  long long pos = 0;
  for (int a = 0; a < K - 1; ++a)
    for (int b = a + 1; b < K; ++b, ++pos) {
      double w = pair_len[pos];
      if (bnd_len) w = std::min(w, bnd_len[a] + bnd_len[b]);
      sb.add_edge(a, b, -(2 * to_fixed(w * scale)));
    }
  if (n > K)
    for (int a = 0; a < K; ++a) sb.add_edge(a, K, -(2 * to_fixed(bnd_len[a] * scale)));
  sb.build_adjacency();
  if (!greedy_start(sb, n, y_tls, vmate_tls) || !sb.solve()) return -3;
  int nq = 0;
  for (int a = 0; a < K; ++a) {
    const int b = sb.mate_vertex(a);
    if (b < 0) return -3;
    if (b >= K) {  // dummy
      out_src[nq] = a;
      out_dst[nq++] = -1;
    } else if (a < b) {
      const long long at = (ll)a * K - (ll)a * (a + 1) / 2 + (b - a - 1);
      if (bnd_len && bnd_len[a] + bnd_len[b] < pair_len[at]) {  // the pair weighs b_a + b_b: both end at their connector
        out_src[nq] = a; out_dst[nq++] = -1;
        out_src[nq] = b; out_dst[nq++] = -1;
      } else {
        out_src[nq] = a;
        out_dst[nq++] = b;
      }
    }
  }
  return nq;
}

// Tuning / statistics of the sparse solver (process-wide).
std::atomic<int> g_mode(0);   // 0: sparse from SPARSE_MIN_K cubes on, 1: always the complete graph, 2: always sparse
std::atomic<int> g_knn(12);    // nearest partners per cube in the first candidate graph
// sparse solves, retries with a denser candidate graph, pricing rounds, priced-in edges, then ns
// spent in: candidate scan, graph build + jump start, solve stages, pricing (summed over threads)
std::atomic<long long> g_stats[8];
inline long long now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
const int SPARSE_MIN_K = 48;

// Same problem as solve_one, solved on a sparse candidate graph and proved optimal for the
// complete graph by pricing (mwpm_sparse.h).  Returns -100 when the caller should fall back to
// the complete graph (candidate graph without a perfect matching, or too many pricing rounds).
int solve_one_sparse(int K, const double* pair_len, const double* bnd_len, int* out_src, int* out_dst, int knn,
                     const int* cand = nullptr, int cand_k = 0) {
  using fpb_sparse::SparseBlossom;
  const int n = (bnd_len && (K & 1)) ? K + 1 : K;
  const long long t_begin = now_ns();
  double wmax = 0;
  const long long npairs = (long long)K * (K - 1) / 2;
  for (long long i = 0; i < npairs; ++i) wmax = std::max(wmax, pair_len[i]);
  if (bnd_len)
    for (int a = 0; a < K; ++a) wmax = std::max(wmax, 2 * bnd_len[a]);
  if (!(wmax < 1e15)) return -2;
  int shift = 32;
  while (shift > 0 && std::ldexp(wmax + 1.0, shift) * (double)(n + 2) > 1e17) --shift;
  const double scale = std::ldexp(1.0, shift);
  // One pass over the packed triangle selects, per vertex, its knn nearest partners in the reduced
  // complete graph (see the header comment of this file).  Length and partner are packed in one
  // 64-bit key (len * (n + 2) < 1e17 by the choice of the scale); thr[u] is the current knn-th
  // smallest key of u and thrd[u] a double-precision bound on the lengths that can still beat it,
  // so that almost every pair is rejected by two floating-point compares in a vectorisable loop.
  int vbits = 1;
  while ((1 << vbits) < n) ++vbits;
  // per vertex a buffer of 2 * knn keys: offers are appended, a full buffer is cut back to its knn
  // smallest keys (nth_element), which also tightens the threshold - O(1) amortised per offer.
  // Scratch is kept per thread (fresh megabyte-sized vectors would be mmap'ed and page-faulted per
  // graph) and used through local pointers (thread_local objects cost a TLS lookup per access).
  const int cap2 = 2 * knn;
  static thread_local std::vector<unsigned long long> top_tls, thr_tls, keys_tls;
  static thread_local std::vector<int> cnt_tls;
  static thread_local std::vector<double> thrd_tls, wrow_tls;
  static thread_local std::vector<unsigned char> flag_tls;
  top_tls.resize((size_t)n * cap2);
  thr_tls.assign((size_t)n, ~0ull);
  cnt_tls.assign((size_t)n, 0);
  thrd_tls.assign((size_t)n, HUGE_VAL);
  wrow_tls.resize((size_t)K + 8);
  flag_tls.resize((size_t)K + 8);
  unsigned long long* const top = top_tls.data();
  unsigned long long* const thr = thr_tls.data();
  int* const cnt = cnt_tls.data();
  double* const thrd = thrd_tls.data();
  double* const wrow = wrow_tls.data();
  unsigned char* const flag = flag_tls.data();
  std::vector<unsigned long long>& keys = keys_tls;
  const double inv_scale = 1.0 / scale;
  auto offer = [&](int u, unsigned long long key) {
    unsigned long long* t = top + (size_t)u * cap2;
    t[cnt[u]++] = key;
    if (cnt[u] == cap2) {
      std::nth_element(t, t + (knn - 1), t + cap2);
      thr[u] = t[knn - 1];
      // key < thr[u] needs len <= len(thr[u]), i.e. w * scale < len(thr[u]) + 0.5
      thrd[u] = ((double)(thr[u] >> vbits) + 1.0) * inv_scale;
      cnt[u] = knn;
    }
  };
  auto reduced = [&](int u, int v) -> double {  // u < v < K
    const double w = pair_len[(size_t)((ll)u * K - (ll)u * (u + 1) / 2 + (v - u - 1))];
    return bnd_len ? std::min(w, bnd_len[u] + bnd_len[v]) : w;
  };
  if (cand) {
    // candidate partners computed on the GPU (fpb_knn): the first knn of each cube's list
    const int take = std::min(knn, cand_k);
    for (int u = 0; u < K; ++u)
      for (int i = 0; i < take; ++i) {
        const int v = cand[(size_t)u * cand_k + i];
        if (v < 0 || v >= K || v == u) break;
        const unsigned long long L = (unsigned long long)to_fixed((v < u ? reduced(v, u) : reduced(u, v)) * scale);
        top[(size_t)u * cap2 + cnt[u]++] = (L << vbits) | (unsigned)v;
      }
    for (int u = 0; u < K; ++u)
      if (cnt[u] == knn) {  // worst candidate = threshold, as the scan would have left it
        unsigned long long worst = 0;
        for (int i = 0; i < cnt[u]; ++i) worst = std::max(worst, top[(size_t)u * cap2 + i]);
        thr[u] = worst;
      }
  } else {
  // Start every threshold at 1.3 x the median knn-th reduced length of a few sampled vertices:
    // a vertex whose true knn-th length is larger ends the scan with fewer than knn offers and is
    // redone exactly below; for all others the filter cannot drop one of their knn nearest.
    auto gather_row = [&](int u, double* out) {  // reduced lengths from u to every v != u (v < K)
      int c = 0;
      for (int v = 0; v < u; ++v) out[c++] = reduced(v, u);
      for (int v = u + 1; v < K; ++v) out[c++] = reduced(u, v);
      return c;
    };
    if (K - 1 > 2 * knn) {
      const int n_s = std::min(K, 16);
      double d_s[16];
      for (int i = 0; i < n_s; ++i) {
        const int c = gather_row((int)((long long)i * K / n_s), wrow);
        std::nth_element(wrow, wrow + (knn - 1), wrow + c);
        d_s[i] = wrow[knn - 1];
      }
      std::nth_element(d_s, d_s + n_s / 2, d_s + n_s);
      const double t0 = 1.3 * d_s[n_s / 2] + inv_scale;
      for (int u = 0; u < K; ++u) thrd[u] = t0;
    }
    {
      long long pos = 0;
      for (int a = 0; a < K - 1; ++a) {
        const int nb = K - 1 - a;
        const double* row = pair_len + pos;
        pos += nb;
        if (bnd_len) {
          const double ba = bnd_len[a];
          const double* bb = bnd_len + a + 1;
          for (int j = 0; j < nb; ++j) wrow[j] = std::min(row[j], ba + bb[j]);
          row = wrow;
        }
        const double ta = thrd[a];  // may go stale (too large) during the row: only costs exact checks
        const double* tb = thrd + a + 1;
        for (int j = 0; j < nb; ++j) flag[j] = (unsigned char)((row[j] < ta) | (row[j] < tb[j]));
        for (int j = 0; j < nb; ++j) {
          if (!flag[j]) continue;
          const int b = a + 1 + j;
          const unsigned long long L = (unsigned long long)to_fixed(row[j] * scale);
          const unsigned long long kb = (L << vbits) | (unsigned)b, ka = (L << vbits) | (unsigned)a;
          if (kb < thr[a]) offer(a, kb);
          if (ka < thr[b]) offer(b, ka);
        }
      }
    }
    for (int u = 0; u < K; ++u) {
      if (cnt[u] >= knn || cnt[u] >= K - 1) continue;
      // fewer than knn partners under the starting threshold: select from the whole row
      cnt[u] = 0;
      thr[u] = ~0ull;
      for (int v = 0; v < K; ++v) {
        if (v == u) continue;
        const unsigned long long L = (unsigned long long)to_fixed((v < u ? reduced(v, u) : reduced(u, v)) * scale);
        const unsigned long long key = (L << vbits) | (unsigned)v;
        if (key < thr[u]) offer(u, key);
      }
    }
  }
  std::vector<ll> dlen;
  if (n > K) {
    dlen.resize(K);
    for (int a = 0; a < K; ++a) {
      dlen[a] = to_fixed(bnd_len[a] * scale);
      const unsigned long long kd = ((unsigned long long)dlen[a] << vbits) | (unsigned)K;
      const unsigned long long ka = ((unsigned long long)dlen[a] << vbits) | (unsigned)a;
      if (kd < thr[a]) offer(a, kd);
      if (ka < thr[K]) offer(K, ka);
    }
  }
  int wshift = 0;  // weights and duals are doubled at every warm restart (see below)
  auto wt = [&](int u, int v) -> ll {  // matcher weight (maximised) of edge u < v
    if (v >= K) return -((2 * dlen[u]) << wshift);
    return -((2 * to_fixed(reduced(u, v) * scale)) << wshift);
  };
  const long long t_scan = now_ns();
  g_stats[4].fetch_add(t_scan - t_begin);
  (void)keys;
  // cut every list back to its knn nearest, then emit each candidate edge once: the pair (u, v)
  // comes from the smaller endpoint's list, or from the larger one's if the smaller does not list it
  const unsigned long long vmask = (1ull << vbits) - 1;
  for (int u = 0; u < n; ++u) {
    unsigned long long* t = top + (size_t)u * cap2;
    if (cnt[u] > knn) {
      std::nth_element(t, t + (knn - 1), t + cnt[u]);
      cnt[u] = knn;
    }
  }
  auto lists = [&](int u, int v) {  // is v in u's list?
    const unsigned long long* t = top + (size_t)u * cap2;
    for (int i = 0; i < cnt[u]; ++i)
      if ((int)(t[i] & vmask) == v) return true;
    return false;
  };
  static thread_local SparseBlossom sb_tls;
  SparseBlossom& sb = sb_tls;
  sb.reset(n);
  for (int u = 0; u < n; ++u) {
    const unsigned long long* t = top + (size_t)u * cap2;
    for (int i = 0; i < cnt[u]; ++i) {
      const int v = (int)(t[i] & vmask);
      if (u < v || !lists(v, u)) sb.add_edge(std::min(u, v), std::max(u, v), -(ll)(2 * (t[i] >> vbits)));
    }
  }
  if (bnd_len) {
    // Cubes that end at their connector are paired with EACH OTHER in the reduced graph, at
    // b_a + b_b whatever the pairing.  Nearest partners alone give that group a star (everyone
    // points at the few cubes closest to the boundary), which has no perfect matching; a chain
    // through the cubes in order of their boundary length lets neighbours in that order pair up.
    static thread_local std::vector<int> order_tls;
    order_tls.resize((size_t)K);
    int* const order = order_tls.data();
    for (int a = 0; a < K; ++a) order[a] = a;
    std::sort(order, order + K, [&](int a, int b) { return bnd_len[a] < bnd_len[b] || (bnd_len[a] == bnd_len[b] && a < b); });
    for (int i = 0; i < K; ++i)
      for (int j = 1; j <= 3 && i + j < K; ++j) {
        const int a = std::min(order[i], order[i + j]), b2 = std::max(order[i], order[i + j]);
        if (!lists(a, b2) && !lists(b2, a)) sb.add_edge(a, b2, wt(a, b2));
      }
  }
  sb.build_adjacency();
  static thread_local std::vector<ll> y_tls;
  static thread_local std::vector<int> vmate_tls;
  std::vector<ll>& y = y_tls;
  std::vector<int>& vmate = vmate_tls;
  auto cold_start = [&]() -> bool { return greedy_start(sb, n, y, vmate); };
  if (!cold_start()) return -100;
  g_stats[0].fetch_add(1);
  g_stats[5].fetch_add(now_ns() - t_scan);
  struct Viol { ll s; int u, v; };
  std::vector<Viol> viol;
  std::vector<int> marked;
  bool done = false;
  for (int round = 0; round < 40 && !done; ++round) {
    const long long t_a = now_ns();
    if (!sb.solve()) return -100;
    const long long t_b = now_ns();
    g_stats[6].fetch_add(t_b - t_a);
    // pricing: reduced cost of every edge of the complete graph under the sparse optimum's duals.
    // A double-precision estimate with a safety margin clears almost every edge in a vectorisable
    // loop; the few near or below zero are recomputed exactly in integers.
    viol.clear();
    {
      static thread_local std::vector<double> duald_tls;
      duald_tls.resize((size_t)n);
      double* const duald = duald_tls.data();
      double dmax = 0;
      for (int v = 0; v < n; ++v) {
        duald[v] = (double)sb.dual[v];
        dmax = std::max(dmax, std::fabs(duald[v]));
      }
      const double c4 = std::ldexp(4.0 * scale, wshift);  // reduced cost = y_u + y_v + 4 * 2^wshift * len
      const double margin = std::ldexp(8.0, wshift) + 1e-13 * (2 * dmax + c4 * wmax) + 64.0;
      auto exact = [&](int u, int v) {
        ll sl = sb.dual[u] + sb.dual[v] - 2 * wt(u, v);
        if (sl < 0) {
          sl += sb.common_blossom_dual(u, v);
          if (sl < 0) viol.push_back(Viol{sl, u, v});
        }
      };
      long long pos = 0;
      for (int u = 0; u < K - 1; ++u) {
        const int nb = K - 1 - u;
        const double* row = pair_len + pos;
        pos += nb;
        if (bnd_len) {
          const double bu = bnd_len[u];
          const double* bb = bnd_len + u + 1;
          for (int j = 0; j < nb; ++j) wrow[j] = std::min(row[j], bu + bb[j]);
          row = wrow;
        }
        const double du = duald[u] - margin;
        const double* dv = duald + u + 1;
        for (int j = 0; j < nb; ++j) flag[j] = (unsigned char)(du + dv[j] + c4 * row[j] < 0.0);
        for (int j = 0; j < nb; j += 8) {  // flags are sparse: test eight at a time
          unsigned long long eight;
          std::memcpy(&eight, flag + j, 8);  // flag[] has 8 bytes of slack past K
          if (!eight) continue;
          for (int q = j; q < std::min(nb, j + 8); ++q)
            if (flag[q]) exact(u, u + 1 + q);
        }
      }
      if (n > K)
        for (int u = 0; u < K; ++u) exact(u, K);
    }
    g_stats[7].fetch_add(now_ns() - t_b);
    if (viol.empty()) {
      done = true;
      break;
    }
    g_stats[2].fetch_add(1);
    const size_t cap = (size_t)4 * n;
    if (viol.size() > cap) {
      std::nth_element(viol.begin(), viol.begin() + cap, viol.end(), [](const Viol& a, const Viol& b) { return a.s < b.s; });
      viol.resize(cap);
    }
    g_stats[3].fetch_add((long long)viol.size());
    // Warm restart: keep matching, duals and blossoms.  Everything is doubled first, so that all
    // duals are even again (the stages need the free vertices' duals to share one parity).  A
    // violated edge becomes feasible by raising the dual of one endpoint (after dissolving the
    // blossoms around it) - that only adds slack to its other edges - and that vertex and its
    // partner become free: the next solve() runs a few stages instead of starting over.
    bool warm = wshift < 6;
    if (warm) {
      sb.double_all();
      ++wshift;
      marked.clear();
      for (size_t i = 0; i < viol.size() && warm; ++i) {
        const int u = viol[i].u, v = viol[i].v;
        if (sb.dual[u] + sb.dual[v] - 2 * wt(u, v) + sb.common_blossom_dual(u, v) >= 0) continue;  // covered by an earlier raise
        const int x = sb.is_bare(u) ? u : (sb.is_bare(v) ? v : u);
        sb.make_bare(x);  // dissolves the blossoms around x (their bases become free)
        const ll rc = sb.dual[u] + sb.dual[v] - 2 * wt(u, v) + sb.common_blossom_dual(u, v);
        if (rc >= 0) continue;
        sb.dual[x] -= rc;
        marked.push_back(x);
      }
    }
    if (warm) {
      for (size_t i = 0; i < viol.size(); ++i) sb.add_edge(viol[i].u, viol[i].v, wt(viol[i].u, viol[i].v));
      sb.build_adjacency();
      for (size_t i = 0; i < marked.size(); ++i) sb.unmatch(marked[i]);
    } else {
      // too many doublings: start over on the enlarged graph
      wshift = 0;
      for (int k = 0; k < sb.n_edges(); ++k) sb.E[k].w = wt(sb.E[k].u, sb.E[k].v);
      for (size_t i = 0; i < viol.size(); ++i) sb.add_edge(viol[i].u, viol[i].v, wt(viol[i].u, viol[i].v));
      sb.build_adjacency();
      if (!cold_start()) return -100;
    }
  }
  if (!done) return -100;
  int nq = 0;
  for (int a = 0; a < K; ++a) {
    const int b = sb.mate_vertex(a);
    if (b < 0) return -3;
    if (b >= K) {
      out_src[nq] = a;
      out_dst[nq++] = -1;
    } else if (a < b) {
      const long long pos = (ll)a * K - (ll)a * (a + 1) / 2 + (b - a - 1);
      if (bnd_len && bnd_len[a] + bnd_len[b] < pair_len[pos]) {
        out_src[nq] = a; out_dst[nq++] = -1;
        out_src[nq] = b; out_dst[nq++] = -1;
      } else {
        out_src[nq] = a;
        out_dst[nq++] = b;
      }
    }
  }
  return nq;
}

int solve_auto(int K, const double* pair_len, const double* bnd_len, int* out_src, int* out_dst,
               const int* cand = nullptr, int cand_k = 0) {
  const int mode = g_mode.load();
  if (K >= 4 && (bnd_len || !(K & 1)) && (mode == 2 || (mode == 0 && K >= SPARSE_MIN_K))) {
    // a candidate graph without a perfect matching (or one that needs too many pricing rounds)
    // is retried with four times the partners; with K - 1 partners it is the complete graph
    for (int knn = std::max(2, g_knn.load());; knn *= 4) {
      const bool use_cand = cand && knn <= cand_k;  // a denser retry needs the host's own scan
      const int r = solve_one_sparse(K, pair_len, bnd_len, out_src, out_dst, std::min(knn, K), use_cand ? cand : nullptr,
                                     cand_k);
      if (r != -100) return r;
      g_stats[1].fetch_add(1);
      if (knn >= K) break;
    }
  }
  return solve_one(K, pair_len, bnd_len, out_src, out_dst);
}


// ---------------------------------------------------------------------------------------------
// Sparse hand-off (include/fpb_mwpm.h, fpb_mwpm_sparse_*): the same sparse-candidate-graph +
// pricing solver as solve_one_sparse, for a whole batch of trials whose K(K-1)/2 lengths stay on
// the GPU.  A job sees only the candidate lists with their reduced lengths (fpb_knn_len), the
// lengths of the few pairs it asks for (chain edges; every pair of a small graph: fpb_pairs_at) and
// the pairs the device-side pricing pass flags under its duals (fpb_price).  The solver state of
// every trial is kept across the pricing rounds of the batch, so a violated edge costs the same
// warm restart as in solve_one_sparse.
// ---------------------------------------------------------------------------------------------
struct SparseJob {
  enum State { EMPTY, DENSE_WAIT, NEED_START, NEED_SOLVE, PRICING, DONE, FAILED };
  State state = EMPTY;
  int K = 0, n = 0, wshift = 0, rounds = 0, knn = 0;
  bool bnd = false;
  double scale = 0, wmax = 0;
  std::vector<double> blen;   // boundary lengths
  std::vector<ll> dlen;       // fixed-point boundary lengths (edges to the dummy vertex K)
  fpb_sparse::SparseBlossom sb;
  std::vector<ll> w0;         // matcher weight of every edge of sb.E at wshift = 0
  std::vector<double> elen;   // reduced length of every edge of sb.E (dummy edges: boundary length)
  std::vector<ll> y;
  std::vector<int> vmate;
  std::vector<int> req_a, req_b;  // pairs whose lengths were asked for
  struct Viol { ll s; int u, v; double len; };
  std::vector<Viol> pending;  // violated edges to the dummy vertex, found on the host
  std::vector<int> res_src, res_dst;
  int error = 0;

  ll fixed_w(double len) const { return -(2 * to_fixed(len * scale)); }
  ll wt_now(ll base) const { return base * ((ll)1 << wshift); }
  void add_edge(int u, int v, double len) {
    const ll b = (v >= K) ? -(2 * dlen[u]) : fixed_w(len);
    sb.add_edge(u, v, wt_now(b));
    w0.push_back(b);
    elen.push_back(len);
  }

  void init(int K_, const int* cand, const double* cand_len, int cand_k, const double* bnd_len, double wmax_, int knn_) {
    // jobs are recycled between batches (their vectors keep their capacity): reset every field
    state = EMPTY;
    n = wshift = rounds = error = 0;
    scale = 0;
    blen.clear(); dlen.clear(); w0.clear(); elen.clear(); req_a.clear(); req_b.clear(); pending.clear();
    res_src.clear(); res_dst.clear();
    K = K_;
    bnd = bnd_len != nullptr;
    knn = knn_;
    wmax = wmax_;
    if (K == 0) { state = DONE; return; }
    if (!bnd && (K & 1)) { state = FAILED; error = -1; return; }
    n = (bnd && (K & 1)) ? K + 1 : K;
    if (bnd) blen.assign(bnd_len, bnd_len + K);
    if (K < SPARSE_MIN_K || K < 4) {  // small graph: ask for every pair and solve the complete graph
      for (int a = 0; a < K - 1; ++a)
        for (int b = a + 1; b < K; ++b) { req_a.push_back(a); req_b.push_back(b); }
      state = DENSE_WAIT;
      return;
    }
    if (!(wmax < 1e15)) { state = FAILED; error = -2; return; }
    int shift = 32;
    while (shift > 0 && std::ldexp(wmax + 1.0, shift) * (double)(n + 2) > 1e17) --shift;
    scale = std::ldexp(1.0, shift);
    sb.reset(n);
    w0.clear();
    elen.clear();
    if (n > K) {
      dlen.resize(K);
      for (int a = 0; a < K; ++a) dlen[a] = to_fixed(blen[a] * scale);
    }
    // candidate edges, each pair once: from the smaller endpoint's list, or from the larger one's if the
    // smaller does not list it (as in solve_one_sparse)
    const int take = std::min(knn, cand_k);
    std::vector<double> worst(K, 0.0);
    auto lists = [&](int u, int v) {  // is v among the first `take` candidates of u?
      const int* c = cand + (size_t)u * cand_k;
      for (int i = 0; i < take; ++i) {
        if (c[i] == v) return true;
        if (c[i] < 0) break;
      }
      return false;
    };
    for (int u = 0; u < K; ++u)
      for (int i = 0; i < take; ++i) {
        const int v = cand[(size_t)u * cand_k + i];
        if (v < 0 || v >= K || v == u) break;
        const double len = cand_len[(size_t)u * cand_k + i];
        worst[u] = std::max(worst[u], len);
        if (u < v || !lists(v, u)) add_edge(std::min(u, v), std::max(u, v), len);
      }
    auto listed = [&](int a, int b) { return lists(a, b) || lists(b, a); };
    if (n > K) {
      // dummy vertex: the cubes nearest to the boundary, and every cube whose boundary length beats
      // its worst candidate
      std::vector<int> order(K);
      for (int a = 0; a < K; ++a) order[a] = a;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return blen[a] < blen[b] || (blen[a] == blen[b] && a < b); });
      std::vector<char> in(K, 0);
      for (int i = 0; i < std::min(K, knn); ++i) in[order[i]] = 1;
      for (int a = 0; a < K; ++a)
        if (blen[a] < worst[a]) in[a] = 1;
      for (int a = 0; a < K; ++a)
        if (in[a]) add_edge(a, K, blen[a]);
    }
    if (bnd) {
      // chain through the cubes in order of boundary length (see solve_one_sparse): ask for the lengths
      std::vector<int> order(K);
      for (int a = 0; a < K; ++a) order[a] = a;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return blen[a] < blen[b] || (blen[a] == blen[b] && a < b); });
      for (int i = 0; i < K; ++i)
        for (int j = 1; j <= 3 && i + j < K; ++j) {
          const int a = std::min(order[i], order[i + j]), b2 = std::max(order[i], order[i + j]);
          if (!listed(a, b2)) { req_a.push_back(a); req_b.push_back(b2); }
        }
    }
    state = NEED_START;
  }

  // lengths of the requested pairs have arrived
  void start(const double* req_len) {
    if (state == DENSE_WAIT) {
      std::vector<double> tri(req_len, req_len + req_a.size());
      res_src.resize(K);
      res_dst.resize(K);
      const int r = solve_one(K, tri.data(), bnd ? blen.data() : nullptr, res_src.data(), res_dst.data());
      if (r < 0) { state = FAILED; error = r; return; }
      res_src.resize(r);
      res_dst.resize(r);
      state = DONE;
      return;
    }
    if (state != NEED_START) return;
    for (size_t i = 0; i < req_a.size(); ++i) {
      const int a = req_a[i], b = req_b[i];
      add_edge(a, b, bnd ? std::min(req_len[i], blen[a] + blen[b]) : req_len[i]);
    }
    sb.build_adjacency();
    state = cold_start() ? NEED_SOLVE : FAILED;
    if (state == FAILED) error = -100;
  }

  bool cold_start() { return greedy_start(sb, n, y, vmate); }

  // one solve; on success the duals for the device-side pricing pass: an edge of reduced length w is
  // violated iff y_u + y_v + ((4 * to_fixed(w * scale)) << wshift) + blossom duals < 0
  void solve(ll* y_out, int* top_out, ll* ztop_out, double* scale_out, int* wshift_out) {
    if (state != NEED_SOLVE) return;
    if (!sb.solve()) { state = FAILED; error = -100; return; }
    for (int v = 0; v < K; ++v) y_out[v] = sb.dual[v];
    if (top_out)
      for (int v = 0; v < K; ++v) {
        const int b = sb.top_blossom(v);
        top_out[v] = b >= n ? b : -1;
        ztop_out[v] = b >= n ? 2 * sb.dual[b] : 0;
      }
    *scale_out = scale;
    *wshift_out = wshift;
    pending.clear();
    if (n > K)
      for (int u = 0; u < K; ++u) {
        ll sl = sb.dual[u] + sb.dual[K] - 2 * wt_now(-(2 * dlen[u]));
        if (sl < 0) {
          sl += sb.common_blossom_dual(u, K);
          if (sl < 0) pending.push_back(Viol{sl, u, K, blen[u]});
        }
      }
    state = PRICING;
  }

  // pairs the pricing pass flagged for this trial (u < v < K, reduced lengths); returns true if the
  // job needs another solve
  bool apply(const int* vu, const int* vv, const double* vlen, size_t cnt, int max_rounds) {
    if (state != PRICING) return false;
    std::vector<Viol> viol(pending);
    for (size_t i = 0; i < cnt; ++i) {
      const int u = vu[i], v = vv[i];
      ll sl = sb.dual[u] + sb.dual[v] - 2 * wt_now(fixed_w(vlen[i]));
      if (sl < 0) {
        sl += sb.common_blossom_dual(u, v);
        if (sl < 0) viol.push_back(Viol{sl, u, v, vlen[i]});
      }
    }
    if (viol.empty()) {
      finish();
      return false;
    }
    if (++rounds > max_rounds) { state = FAILED; error = -100; return false; }  // the caller solves it from the full triangle
    g_stats[2].fetch_add(1);
    const size_t cap = (size_t)4 * n;
    if (viol.size() > cap) {
      std::nth_element(viol.begin(), viol.begin() + cap, viol.end(), [](const Viol& a, const Viol& b) { return a.s < b.s; });
      viol.resize(cap);
    }
    g_stats[3].fetch_add((long long)viol.size());
    auto wt = [&](const Viol& e) { return wt_now(e.v >= K ? -(2 * dlen[e.u]) : fixed_w(e.len)); };
    if (wshift < 6) {  // warm restart, as in solve_one_sparse
      sb.double_all();
      ++wshift;
      std::vector<int> marked;
      for (size_t i = 0; i < viol.size(); ++i) {
        const int u = viol[i].u, v = viol[i].v;
        if (sb.dual[u] + sb.dual[v] - 2 * wt(viol[i]) + sb.common_blossom_dual(u, v) >= 0) continue;
        const int x = sb.is_bare(u) ? u : (sb.is_bare(v) ? v : u);
        sb.make_bare(x);
        const ll rc = sb.dual[u] + sb.dual[v] - 2 * wt(viol[i]) + sb.common_blossom_dual(u, v);
        if (rc >= 0) continue;
        sb.dual[x] -= rc;
        marked.push_back(x);
      }
      for (size_t i = 0; i < viol.size(); ++i) add_edge(viol[i].u, viol[i].v, viol[i].len);
      sb.build_adjacency();
      for (size_t i = 0; i < marked.size(); ++i) sb.unmatch(marked[i]);
      state = NEED_SOLVE;
    } else {  // too many doublings: start over on the enlarged graph
      wshift = 0;
      for (int k = 0; k < sb.n_edges(); ++k) sb.E[k].w = w0[k];
      for (size_t i = 0; i < viol.size(); ++i) add_edge(viol[i].u, viol[i].v, viol[i].len);
      sb.build_adjacency();
      state = cold_start() ? NEED_SOLVE : FAILED;
      if (state == FAILED) error = -100;
    }
    return state == NEED_SOLVE;
  }

  void finish() {
    res_src.clear();
    res_dst.clear();
    for (int a = 0; a < K; ++a) {
      const int b = sb.mate_vertex(a);
      if (b < 0) { state = FAILED; error = -3; return; }
      if (b >= K) {
        res_src.push_back(a);
        res_dst.push_back(-1);
      } else if (a < b) {
        const double len = elen[sb.mate[a] >> 1];
        if (bnd && !(len < blen[a] + blen[b])) {  // the pair weighs b_a + b_b: both end at their connector
          res_src.push_back(a); res_dst.push_back(-1);
          res_src.push_back(b); res_dst.push_back(-1);
        } else {
          res_src.push_back(a);
          res_dst.push_back(b);
        }
      }
    }
    state = DONE;
  }
};

struct SparseBatch {
  int T = 0, threads = 0;
  std::vector<long long> q_off;  // [T + 1]
  std::vector<long long> req_off;  // [T + 1] offsets of each trial's requests
  std::vector<SparseJob> jobs;   // grows to the largest batch seen; only the first T are in use
};
// finished batches are kept for reuse: a job owns some thirty vectors (the solver's forests, blossom
// lists, adjacency), and allocating them afresh for every trial costs more than the pricing pass saves
std::mutex g_pool_mutex;
std::vector<SparseBatch*> g_pool;

}  // namespace

extern "C" {

// mode 0: automatic (sparse + pricing for large graphs), 1: complete graph only, 2: sparse only
// (the complete-graph last resort stays).  knn > 0 sets the nearest-partner count of the first candidate graph.
void fpb_mwpm_configure(int mode, int knn) {
  g_mode.store(mode);
  if (knn > 0) g_knn.store(knn);
}

// out[0..7] = sparse solves, retries with a denser candidate graph, pricing rounds, priced-in edges since load / last reset.
void fpb_mwpm_stats(long long* out, int reset) {
  for (int i = 0; i < 8; ++i) {
    out[i] = g_stats[i].load();
    if (reset) g_stats[i].store(0);
  }
}

// Single graph (see solve_one).  out_src / out_dst need room for K entries.
int fpb_mwpm(int K, const double* pair_len, const double* bnd_len, int* out_src, int* out_dst) {
  return solve_auto(K, pair_len, bnd_len, out_src, out_dst);
}

// T graphs in the ragged layout of fpb_outputs (odd_count[T], pair_len / bnd_len concatenated).
// q_off[T+1] must hold the exclusive prefix of odd_count (each trial writes at most K queries);
// n_queries[T] receives the counts.  OpenMP over trials.
int fpb_mwpm_batch(int T, const int* odd_count, const double* pair_len, const double* bnd_len,
                   const long long* q_off, int* out_src, int* out_dst, int* n_queries, int n_threads) {
  std::vector<long long> poff((size_t)T + 1, 0);
  for (int t = 0; t < T; ++t) poff[t + 1] = poff[t] + (long long)odd_count[t] * (odd_count[t] - 1) / 2;
  int bad = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
  for (int t = 0; t < T; ++t) {
    const int K = odd_count[t];
    const int r = solve_auto(K, pair_len + poff[t], bnd_len ? bnd_len + q_off[t] : nullptr, out_src + q_off[t],
                            out_dst + q_off[t]);
    if (r < 0) bad |= 1;
    n_queries[t] = r < 0 ? 0 : r;
  }
  return bad ? -1 : 0;
}

// Same as fpb_mwpm_batch with the candidate partners of every odd cube supplied (fpb_knn in fpb.h:
// cand[q_off[t] + a][cand_k], indices into trial t's odd list, nearest first, -1 padded), which saves
// the solver's own selection pass over the K(K-1)/2 lengths.  The result is the same exact optimum.
int fpb_mwpm_batch_cand(int T, const int* odd_count, const double* pair_len, const double* bnd_len,
                        const long long* q_off, const int* cand, int cand_k, int* out_src, int* out_dst,
                        int* n_queries, int n_threads) {
  std::vector<long long> poff((size_t)T + 1, 0);
  for (int t = 0; t < T; ++t) poff[t + 1] = poff[t] + (long long)odd_count[t] * (odd_count[t] - 1) / 2;
  int bad = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
  for (int t = 0; t < T; ++t) {
    const int K = odd_count[t];
    const int r = solve_auto(K, pair_len + poff[t], bnd_len ? bnd_len + q_off[t] : nullptr, out_src + q_off[t],
                             out_dst + q_off[t], cand ? cand + q_off[t] * cand_k : nullptr, cand_k);
    if (r < 0) bad |= 1;
    n_queries[t] = r < 0 ? 0 : r;
  }
  return bad ? -1 : 0;
}

// ---- sparse hand-off (see SparseJob above and include/fpb_mwpm.h) ----
void* fpb_mwpm_sparse_begin(int T, const int* odd_count, const long long* q_off, const int* cand,
                            const double* cand_len, int cand_k, const double* bnd_len, const double* wmax,
                            int n_threads) {
  SparseBatch* B = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool.empty()) {
      B = g_pool.back();
      g_pool.pop_back();
    }
  }
  if (!B) B = new SparseBatch();
  B->T = T;
  B->threads = n_threads;
  B->q_off.assign(q_off, q_off + T + 1);
  if ((int)B->jobs.size() < T) B->jobs.resize((size_t)T);
  const int knn = std::max(2, g_knn.load());
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < T; ++t)
    B->jobs[t].init(odd_count[t], cand + q_off[t] * cand_k, cand_len + q_off[t] * cand_k, cand_k,
                    bnd_len ? bnd_len + q_off[t] : nullptr, wmax[t], knn);
  B->req_off.assign((size_t)T + 1, 0);
  for (int t = 0; t < T; ++t) B->req_off[t + 1] = B->req_off[t] + (long long)B->jobs[t].req_a.size();
  return B;
}

long long fpb_mwpm_sparse_requests(void* ctx, int* trial, int* a, int* b, long long cap) {
  SparseBatch* B = static_cast<SparseBatch*>(ctx);
  const long long total = B->req_off[B->T];
  if (cap < total || !trial || !a || !b) return total;
  for (int t = 0; t < B->T; ++t) {
    const SparseJob& J = B->jobs[t];
    long long o = B->req_off[t];
    for (size_t i = 0; i < J.req_a.size(); ++i, ++o) {
      trial[o] = t;
      a[o] = J.req_a[i];
      b[o] = J.req_b[i];
    }
  }
  return total;
}

int fpb_mwpm_sparse_start(void* ctx, const double* req_len) {
  SparseBatch* B = static_cast<SparseBatch*>(ctx);
#ifdef _OPENMP
  if (B->threads > 0) omp_set_num_threads(B->threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < B->T; ++t) B->jobs[t].start(req_len ? req_len + B->req_off[t] : nullptr);
  return 0;
}

// Solve every job that needs it; y [sum K], scale / wshift / active [T] receive what fpb_price takes.
// Returns the number of jobs awaiting a pricing pass.
int fpb_mwpm_sparse_solve(void* ctx, long long* y, int* top, long long* ztop, double* scale, int* wshift,
                          unsigned char* active) {
  SparseBatch* B = static_cast<SparseBatch*>(ctx);
  int n_active = 0;
#ifdef _OPENMP
  if (B->threads > 0) omp_set_num_threads(B->threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_active)
  for (int t = 0; t < B->T; ++t) {
    SparseJob& J = B->jobs[t];
    active[t] = 0;
    if (J.state == SparseJob::NEED_SOLVE) {
      g_stats[0].fetch_add(J.rounds == 0 ? 1 : 0);
      J.solve(y + B->q_off[t], top ? top + B->q_off[t] : nullptr, ztop ? ztop + B->q_off[t] : nullptr, scale + t,
              wshift + t);
    }
    if (J.state == SparseJob::PRICING) {
      active[t] = 1;
      ++n_active;
    }
  }
  return n_active;
}

// The flagged pairs of fpb_price (any order).  Returns the number of jobs that need another solve.
int fpb_mwpm_sparse_apply(void* ctx, long long n, const int* vt, const int* va, const int* vb, const double* vlen,
                          int max_rounds) {
  SparseBatch* B = static_cast<SparseBatch*>(ctx);
  // bucket the pairs by trial
  std::vector<long long> off((size_t)B->T + 1, 0);
  for (long long i = 0; i < n; ++i) ++off[vt[i] + 1];
  for (int t = 0; t < B->T; ++t) off[t + 1] += off[t];
  std::vector<int> ua((size_t)n), ub((size_t)n);
  std::vector<double> ul((size_t)n);
  {
    std::vector<long long> fill(off.begin(), off.end() - 1);
    for (long long i = 0; i < n; ++i) {
      const long long o = fill[vt[i]]++;
      ua[o] = va[i];
      ub[o] = vb[i];
      ul[o] = vlen[i];
    }
  }
  int again = 0;
#ifdef _OPENMP
  if (B->threads > 0) omp_set_num_threads(B->threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : again)
  for (int t = 0; t < B->T; ++t)
    if (B->jobs[t].apply(ua.data() + off[t], ub.data() + off[t], ul.data() + off[t], (size_t)(off[t + 1] - off[t]),
                         max_rounds > 0 ? max_rounds : 40))
      ++again;
  return again;
}

// Queries of every finished trial (layout of fpb_mwpm_batch); failed [T] flags the trials the caller
// must solve from the full triangle instead (no perfect matching in the candidate graph, ...).
int fpb_mwpm_sparse_finish(void* ctx, int* out_src, int* out_dst, int* n_queries, unsigned char* failed) {
  SparseBatch* B = static_cast<SparseBatch*>(ctx);
  int n_failed = 0;
  for (int t = 0; t < B->T; ++t) {
    const SparseJob& J = B->jobs[t];
    failed[t] = J.state != SparseJob::DONE;
    n_queries[t] = 0;
    if (failed[t]) {
      ++n_failed;
      continue;
    }
    n_queries[t] = (int)J.res_src.size();
    std::copy(J.res_src.begin(), J.res_src.end(), out_src + B->q_off[t]);
    std::copy(J.res_dst.begin(), J.res_dst.end(), out_dst + B->q_off[t]);
  }
  return n_failed;
}

void fpb_mwpm_sparse_free(void* ctx) {
  if (!ctx) return;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (g_pool.size() < 8) g_pool.push_back(static_cast<SparseBatch*>(ctx));
  else delete static_cast<SparseBatch*>(ctx);
}

int fpb_mwpm_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
