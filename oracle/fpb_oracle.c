/*
 * TEST INFRASTRUCTURE ONLY - C restatement of the reference's Monte-Carlo EC trial.
 *
 * Same algorithm as oracle/restate.py (which is pinned bit-for-bit against the live
 * reference and the vectors in tests/golden/), written in plain C so that the parity
 * tests can check full-size lattices in seconds and bench.py can time a CPU baseline
 * ("port") on all host cores.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * It consumes the same static description as the CUDA library (fpb_config in
 * include/fpb.h).  Reference lines followed (relative to the reference root):
 *   entangle      flamingpy/cv/ops.py:65,84 (sparse [[I,0],[A,I]] product, ascending j, + p_i)
 *   bin           flamingpy/cv/gkp.py:19-62 (np.divmod semantics)
 *   Z_err_cond    flamingpy/cv/gkp.py:116-133 (n = -10..9 in both sums)
 *   weights       flamingpy/decoders/decoder.py:58-116
 *   parity        flamingpy/codes/stabilizer.py:43-55
 *   matching      flamingpy/decoders/mwpm/matching.py:125-173 with networkx-style
 *                 label-setting Dijkstra (codes/graphs/stabilizer_graph.py:392-410)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/fpb.h"

#define SQRT_PI 1.7724538509055159
#define FLOAT_MIN 2.2250738585072014e-308

typedef struct fpo_result {
  int64_t n_trials;
  int32_t n_complex;
  int64_t total_qubits;
  double* hom;      /* [T][Qtot] */
  uint8_t* bits;    /* [T][Qtot] */
  double* weights;  /* [T][Qtot] */
  uint8_t* parity[FPB_MAX_COMPLEX];   /* [T][S] */
  int32_t* odd_count[FPB_MAX_COMPLEX]; /* [T] */
  int64_t* odd_off[FPB_MAX_COMPLEX];   /* [T+1] */
  int64_t* pair_off[FPB_MAX_COMPLEX];  /* [T+1] */
  int32_t* odd_ids[FPB_MAX_COMPLEX];
  double* pair_len[FPB_MAX_COMPLEX];
  double* bnd_len[FPB_MAX_COMPLEX];
  uint8_t* bnd_side[FPB_MAX_COMPLEX];
} fpo_result;

static void gkp_bin(double x, int* bit, double* frac) {
  const double a = SQRT_PI;
  double mod = fmod(x, a);
  double div = (x - mod) / a;
  if (mod != 0.0) {
    if (mod < 0.0) { mod += a; div -= 1.0; }
  } else {
    mod = 0.0;
  }
  double fl;
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = 0.0;
  }
  long long n = (long long)fl;
  if (mod > a / 2) { n += 1; *frac = mod - a; } else { *frac = mod; }
  *bit = (int)(n & 1);
}

static double z_err_cond(double var, double f) {
  double num = 0.0, den = 0.0;
  for (int n = -10; n < 10; ++n) {
    double a = f - (double)(2 * n + 1) * SQRT_PI;
    double b = f - (double)n * SQRT_PI;
    num += exp(-(a * a) / var);
    den += exp(-(b * b) / var);
  }
  return den != 0.0 ? num / den : 0.0;
}

/* binary heap of (key, cube) with lazy deletion */
typedef struct { double k; int v; } hent;
typedef struct { hent* a; int n; } heap_t;
static void hpush(heap_t* h, double k, int v) {
  int i = h->n++;
  while (i > 0) {
    int p = (i - 1) >> 1;
    if (h->a[p].k <= k) break;
    h->a[i] = h->a[p];
    i = p;
  }
  h->a[i].k = k; h->a[i].v = v;
}
static hent hpop(heap_t* h) {
  hent top = h->a[0];
  hent last = h->a[--h->n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && h->a[c + 1].k < h->a[c].k) ++c;
    if (h->a[c].k >= last.k) break;
    h->a[i] = h->a[c];
    i = c;
  }
  h->a[i] = last;
  return top;
}

typedef struct {
  int S;
  int* nbr;       /* [S][6] neighbour cube or -1 */
  const int* cq;  /* [S][6] local qubit */
} cgraph;

static void dijkstra(const cgraph* g, const double* w, const int* seeds, const double* seed_d, int n_seed,
                     double* dist, uint8_t* done, heap_t* h) {
  for (int i = 0; i < g->S; ++i) { dist[i] = INFINITY; done[i] = 0; }
  h->n = 0;
  for (int i = 0; i < n_seed; ++i)
    if (seed_d[i] < dist[seeds[i]]) { dist[seeds[i]] = seed_d[i]; hpush(h, seed_d[i], seeds[i]); }
  while (h->n) {
    hent e = hpop(h);
    int u = e.v;
    if (done[u]) continue;
    done[u] = 1;
    for (int k = 0; k < 6; ++k) {
      int v = g->nbr[u * 6 + k];
      if (v < 0) continue;
      double nd = e.k + w[g->cq[u * 6 + k]];
      if (nd < dist[v]) { dist[v] = nd; hpush(h, nd, v); }
    }
  }
}

void fpo_free(fpo_result* r) {
  if (!r) return;
  free(r->hom); free(r->bits); free(r->weights);
  for (int c = 0; c < FPB_MAX_COMPLEX; ++c) {
    free(r->parity[c]); free(r->odd_count[c]); free(r->odd_off[c]); free(r->pair_off[c]);
    free(r->odd_ids[c]); free(r->pair_len[c]); free(r->bnd_len[c]); free(r->bnd_side[c]);
  }
  memset(r, 0, sizeof *r);
}

int fpo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Run T injected-noise trials.  stage as in fpb.h.  n_threads <= 0: all cores. */
int fpo_run(const fpb_config* cfg, int64_t T, const double* x, const uint8_t* state, int stage,
            int n_threads, fpo_result* res) {
  memset(res, 0, sizeof *res);
  const int N = cfg->n_nodes, ncx = cfg->n_complex;
  int Qtot = 0, qoff[FPB_MAX_COMPLEX];
  for (int c = 0; c < ncx; ++c) { qoff[c] = Qtot; Qtot += cfg->complex_[c].n_qubits; }
  res->n_trials = T; res->n_complex = ncx; res->total_qubits = Qtot;
  res->hom = malloc(sizeof(double) * T * Qtot);
  res->bits = malloc((size_t)T * Qtot);
  res->weights = malloc(sizeof(double) * T * Qtot);
  const double wdelta = cfg->weight_delta > 0 ? cfg->weight_delta : cfg->delta;
  const double mult = cfg->weight_multiplier != 0 ? cfg->weight_multiplier : 1.0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif

  /* ---- noise: entangle, bin, weights ---- */
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < T; ++t) {
    const double* xr = x + t * 2 * (int64_t)N;
    const uint8_t* sr = state + t * (int64_t)N;
    for (int c = 0; c < ncx; ++c) {
      const fpb_complex_desc* d = &cfg->complex_[c];
      for (int q = 0; q < d->n_qubits; ++q) {
        const int i = d->qubit_node[q];
        const int e0 = cfg->adj_indptr[i], e1 = cfg->adj_indptr[i + 1];
        double acc = 0.0;
        int pc = 0;
        for (int e = e0; e < e1; ++e) {
          const int j = cfg->adj_indices[e];
          acc = acc + (double)cfg->adj_vals[e] * xr[j];
          pc += sr[j] == 2;
        }
        const double hom = acc + xr[N + i];
        int bit; double frac;
        gkp_bin(hom, &bit, &frac);
        double w;
        if (cfg->weight_method == FPB_WEIGHT_UNIFORM) {
          w = 1.0;
        } else {
          double err;
          if (pc <= 1) {
            err = z_err_cond((double)(e1 - e0 + 1) * wdelta, frac);
            if (err > 0.5) err = 0.5;
            if (err == 0.0) err = FLOAT_MIN;
          } else {
            err = pc == 2 ? 0.25 : (pc == 3 ? 1.0 / 3.0 : 0.4);
          }
          w = cfg->weight_integer ? rint(-mult * log(err)) : -log(err);
        }
        const int64_t o = t * Qtot + qoff[c] + q;
        res->hom[o] = hom; res->bits[o] = (uint8_t)bit; res->weights[o] = w;
      }
    }
  }
  if (stage < FPB_STAGE_SYNDROME) return 0;

  /* ---- syndrome ---- */
  for (int c = 0; c < ncx; ++c) {
    const fpb_complex_desc* d = &cfg->complex_[c];
    const int S = d->n_cubes;
    res->parity[c] = malloc((size_t)T * S);
    res->odd_count[c] = malloc(sizeof(int32_t) * T);
    res->odd_off[c] = malloc(sizeof(int64_t) * (T + 1));
    res->pair_off[c] = malloc(sizeof(int64_t) * (T + 1));
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
      const uint8_t* b = res->bits + t * Qtot + qoff[c];
      int K = 0;
      for (int s = 0; s < S; ++s) {
        int par = 0;
        for (int f = 0; f < 6; ++f) {
          const int q = d->cube_qubit[s * 6 + f];
          if (q >= 0) par ^= b[q];
        }
        res->parity[c][t * S + s] = (uint8_t)par;
        K += par;
      }
      res->odd_count[c][t] = K;
    }
    res->odd_off[c][0] = 0; res->pair_off[c][0] = 0;
    for (int64_t t = 0; t < T; ++t) {
      const int64_t K = res->odd_count[c][t];
      res->odd_off[c][t + 1] = res->odd_off[c][t] + K;
      res->pair_off[c][t + 1] = res->pair_off[c][t] + K * (K - 1) / 2;
    }
    const int64_t tot = res->odd_off[c][T];
    res->odd_ids[c] = malloc(sizeof(int32_t) * (tot + 1));
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
      int64_t o = res->odd_off[c][t];
      for (int s = 0; s < S; ++s)
        if (res->parity[c][t * S + s]) res->odd_ids[c][o++] = s;
    }
  }
  if (stage < FPB_STAGE_GRAPH) return 0;

  /* ---- matching graph ---- */
  for (int c = 0; c < ncx; ++c) {
    const fpb_complex_desc* d = &cfg->complex_[c];
    const int S = d->n_cubes;
    cgraph g;
    g.S = S; g.cq = d->cube_qubit;
    g.nbr = malloc(sizeof(int) * S * 6);
    for (int s = 0; s < S; ++s)
      for (int k = 0; k < 6; ++k) {
        const unsigned kind = (d->ninfo[s] >> (2 * k)) & 3u;
        g.nbr[s * 6 + k] = kind < 2 ? s + d->d_off[2 * k + kind] : -1;
      }
    const int has_bnd = d->n_low > 0;
    res->pair_len[c] = malloc(sizeof(double) * (res->pair_off[c][T] + 1));
    res->bnd_len[c] = malloc(sizeof(double) * (res->odd_off[c][T] + 1));
    res->bnd_side[c] = malloc((size_t)(res->odd_off[c][T] + 1));
    /* one task per (trial, source) so threads stay busy when trials differ in size */
#pragma omp parallel
    {
      double* dist = malloc(sizeof(double) * S);
      double* dist2 = malloc(sizeof(double) * S);
      uint8_t* done = malloc(S);
      heap_t h; h.a = malloc(sizeof(hent) * (6 * (size_t)S + 16)); h.n = 0;
      int nseed_max = d->n_low > d->n_high ? d->n_low : d->n_high;
      double* sd = malloc(sizeof(double) * (nseed_max + 1));
#pragma omp for schedule(dynamic, 1)
      for (int64_t t = 0; t < T; ++t) {
        const double* w = res->weights + t * Qtot + qoff[c];
        const int K = res->odd_count[c][t];
        const int32_t* odd = res->odd_ids[c] + res->odd_off[c][t];
        double* pl = res->pair_len[c] + res->pair_off[c][t];
        int64_t pos = 0;
        for (int a = 0; a + 1 < K; ++a) {
          const int src = odd[a];
          const double zero = 0.0;
          dijkstra(&g, w, &src, &zero, 1, dist, done, &h);
          for (int b = a + 1; b < K; ++b) pl[pos++] = dist[odd[b]];
        }
        if (has_bnd && K > 0) {
          for (int i = 0; i < d->n_low; ++i) sd[i] = w[d->slot_qubit[d->low_slot[i]]];
          dijkstra(&g, w, d->low_cube, sd, d->n_low, dist, done, &h);
          for (int i = 0; i < d->n_high; ++i) sd[i] = w[d->slot_qubit[d->high_slot[i]]];
          dijkstra(&g, w, d->high_cube, sd, d->n_high, dist2, done, &h);
          double* bl = res->bnd_len[c] + res->odd_off[c][t];
          uint8_t* bs = res->bnd_side[c] + res->odd_off[c][t];
          for (int b = 0; b < K; ++b) {
            const double lo = dist[odd[b]], hi = dist2[odd[b]];
            const int high = hi < lo; /* np.argmin((low, high)): tie -> low */
            bl[b] = high ? hi : lo;
            bs[b] = (uint8_t)high;
          }
        }
      }
      free(dist); free(dist2); free(done); free(h.a); free(sd);
    }
    free(g.nbr);
  }
  return 0;
}
