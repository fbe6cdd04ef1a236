// Device kernels of the macronode noise layer, CVMacroLayer (flamingpy/noise/cv.py:352-625):
//
//   k_macro_gen      Philox labels, initial means and measurement normals of the 4N micronodes
//                    (cv.py:121-140,323-349,205)
//   k_macro_measure  per macronode: stars-first permutation, CZ across the dumbbells, four-splitter,
//                    p-measurement of the star and q-measurement of the planets (cv.py:421-489,529-546)
//   k_macro_reduce   per macronode: processed neighbour outcomes, effective homodyne value, bit value
//                    and conditional phase error probability of the reduced node (cv.py:491-527,548-625)
//   k_macro_weights  per syndrome qubit: assign_weights(prob_precomputed=True) (decoders/decoder.py:58-93)
//                    + bit packing, after which the syndrome / matching-graph stages run as for CVLayer
//
// The reference computes the means in float32 (cv.py:329: np.zeros(2N, float32); cv/ops.py:127: the
// four-splitter .astype(np.single)) and NumPy's float32 matrix-vector product sums a row of four as
// (t0 + t1) + (t2 + t3); both are reproduced so that outcomes, hence bits, are bit-exact.
#pragma once
#include "fpb_kernels.cuh"

namespace fpb {

struct MacroParams {
  int N;                    // canonical nodes (macronodes); micronodes = 4N
  const int* partner;       // [4N] micronode at the other end of the dumbbell, -1 none (padding)
  const int8_t* sign;       // [4N] polarity weight of that edge
  float sd;                 // float32((delta / 2) ** 0.5), cv.py:296
  double delta;
  const uint8_t* state;     // [T][4N] 2 = p-squeezed, else "GKP"
  const float* mean;        // [T][4N] initial q means
  const double* z;          // [T][4N] standard normals: N stars, then 3N planets
  uint8_t* body;            // [T][4N] body index 1..4 of every micronode
  double* val;              // [T][4N] star: p outcome; planet: q outcome
  uint8_t* red_state;       // [T][N] reduced label: 2 = p, 1 = GKP
  uint8_t* bit_val;         // [T][N]
  double* p_phase;          // [T][N]
  double* outcome;          // [T][N] effective homodyne value of the reduced node
};

struct MacroGenParams {
  int M;                    // micronodes
  uint32_t seed_lo, seed_hi;
  long long first_trial, chain_len;
  int fresh;
  double p_swap;
  uint32_t p_thresh;
  uint8_t* state;           // [T][M] 0 silent, 1 GKP, 2 p
  float* mean;              // [T][M]
  double* z;                // [T][M]
};

__device__ __forceinline__ bool macro_label_is_p(const MacroGenParams& g, long long trial, uint32_t node) {
  const Philox4 r = philox4x32_10((uint32_t)trial, (uint32_t)((unsigned long long)trial >> 32), node, SLOT_MACRO_LABEL,
                                  g.seed_lo, g.seed_hi);
  return r.x < g.p_thresh;
}

__global__ void __launch_bounds__(256) k_macro_gen(MacroGenParams g, long long n_trials) {
  const long long total = n_trials * (long long)g.M;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long tl = idx / g.M;
    const uint32_t node = (uint32_t)(idx - tl * g.M);
    const long long trial = g.first_trial + tl;
    const long long chain_start = (g.chain_len > 0) ? (trial / g.chain_len) * g.chain_len : 0;
    int st;
    if (g.p_swap >= 1.0) {
      st = 2;
    } else if (g.p_swap <= 0.0) {
      st = (g.fresh || (((trial - chain_start) & 1) == 0)) ? 1 : 0;
    } else if (macro_label_is_p(g, trial, node)) {
      st = 2;
    } else if (g.fresh) {
      st = 1;
    } else {  // reused layer: GKP_t = V \ (p_t U GKP_{t-1}), as in k_gen
      int run = 1;
      for (long long s = trial - 1; s >= chain_start; --s) {
        if (macro_label_is_p(g, s, node)) break;
        ++run;
      }
      st = (run & 1) ? 1 : 0;
    }
    const Philox4 r = philox4x32_10((uint32_t)trial, (uint32_t)((unsigned long long)trial >> 32), node, SLOT_MACRO_DRAW,
                                    g.seed_lo, g.seed_hi);
    // initial mean (cv.py:332-345): uniform in [0, 2 sqrt(pi)) for p, a fair bit times sqrt(pi) for GKP
    float m = 0.0f;
    if (st == 2) m = (float)(u01_53(r.x, r.y) * (2.0 * SQRT_PI));
    else if (st == 1) m = (r.x & 1u) ? (float)SQRT_PI : 0.0f;
    // measurement normal: Box-Muller from the other 64 bits and a second block
    const Philox4 r2 = philox4x32_10((uint32_t)trial, (uint32_t)((unsigned long long)trial >> 32), node,
                                     SLOT_MACRO_DRAW + 16u, g.seed_lo, g.seed_hi);
    const double u1 = u01_53(r.z, r.w), u2 = u01_53(r2.x, r2.y);
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    g.state[idx] = (uint8_t)st;
    g.mean[idx] = m;
    g.z[idx] = sqrt(-2.0 * log(u1)) * cs;
  }
}

// y = H x with H = +-1/2 (cv/ops.py:120-127), float32, summed (t0 + t1) + (t2 + t3)
__device__ __forceinline__ void four_splitter(const float x[4], float y[4]) {
  const float h0 = __fmul_rn(0.5f, x[0]), h1 = __fmul_rn(0.5f, x[1]), h2 = __fmul_rn(0.5f, x[2]), h3 = __fmul_rn(0.5f, x[3]);
  y[0] = __fadd_rn(__fadd_rn(h0, h1), __fadd_rn(h2, h3));
  y[1] = __fadd_rn(__fadd_rn(-h0, h1), __fadd_rn(-h2, h3));
  y[2] = __fadd_rn(__fadd_rn(-h0, -h1), __fadd_rn(h2, h3));
  y[3] = __fadd_rn(__fadd_rn(h0, -h1), __fadd_rn(-h2, h3));
}

__global__ void __launch_bounds__(256) k_macro_measure(MacroParams p, long long n_trials) {
  const long long total = n_trials * (long long)p.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / p.N;
    const int m = (int)(idx - t * p.N);
    const long long mo = t * 4ll * p.N;  // micronode row of this trial
    const uint8_t* st = p.state + mo + 4 * m;
    // stars first: GKP micronodes in index order, then the p-squeezed ones (cv.py:529-546)
    int perm[4];
    int k = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (st[j] != 2) perm[k++] = j;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (st[j] == 2) perm[k++] = j;
    // means after the CZ gates (cv/ops.py:65,84): q unchanged, p_i = sign * q_partner (float32, exact)
    float xq[4], xp[4], yq[4], yp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = 4 * m + perm[j];
      xq[j] = p.mean[mo + u];
      const int nb = p.partner[u];
      xp[j] = nb >= 0 ? __fmul_rn((float)p.sign[u], p.mean[mo + nb]) : 0.0f;
    }
    four_splitter(xq, yq);
    four_splitter(xp, yp);
    const double sd = (double)p.sd;
    const double* z = p.z + mo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = 4 * m + perm[j];
      // rng.normal(means, covs) = means + covs * z (cv.py:205): star in p, planets in q
      const double v = j == 0 ? __dadd_rn((double)yp[0], __dmul_rn(sd, z[m]))
                              : __dadd_rn((double)yq[j], __dmul_rn(sd, z[p.N + 3 * m + (j - 1)]));
      p.val[mo + u] = v;
      p.body[mo + u] = (uint8_t)(j + 1);
    }
    p.red_state[idx] = st[perm[0]] == 2 ? 2 : 1;
  }
}

// Z_err_cond(var, val, use_hom_val=True) (cv/gkp.py:116-133): both sums over n = -10 .. 9, the
// numerator over the multiples 2n + (1 - bit) of sqrt(pi)
__device__ __forceinline__ double z_err_cond_hom(double var, double val) {
  int bit;
  double frac;
  gkp_bin(val, bit, frac);
  const int factor = 1 - bit;
  double num = 0.0, den = 0.0;
#pragma unroll 1
  for (int n = -10; n < 10; ++n) {
    const double a = val - (double)(2 * n + factor) * SQRT_PI;
    const double b = val - (double)n * SQRT_PI;
    num += exp(-(a * a) / var);
    den += exp(-(b * b) / var);
  }
  return den != 0.0 ? num / den : 0.0;
}

__global__ void __launch_bounds__(128) k_macro_reduce(MacroParams p, long long n_trials) {
  const long long total = n_trials * (long long)p.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / p.N;
    const int m = (int)(idx - t * p.N);
    const long long mo = t * 4ll * p.N;
    const uint8_t* body = p.body + mo;
    const double* val = p.val + mo;
    const uint8_t* st = p.state + mo;
    // micronodes of this macronode by body index
    int at[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) at[body[4 * m + j] - 1] = 4 * m + j;
    double outcome = __dmul_rn(2.0, val[at[0]]);
    double zg[4];
    int n_gkp = 0, num_p = 0;
    for (int i = 0; i < 4; ++i) {  // micronodes i = 1..4 of the manuscript, in body order (cv.py:566-571)
      const int nb = p.partner[at[i]];
      if (nb < 0) continue;  // padding micronode: no neighbour
      const int nbody = body[nb];
      const int mb = nb & ~3;
      double h2 = 0.0, h3 = 0.0, h4 = 0.0;  // planet outcomes of the neighbouring macronode (cv.py:447-465)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = body[mb + j];
        const double v = val[mb + j];
        if (b == 2) h2 = v;
        else if (b == 3) h3 = v;
        else if (b == 4) h4 = v;
      }
      // processed outcome by the neighbour's body index (cv.py:548-564)
      const double zv = nbody == 1 ? 0.0 : nbody == 2 ? __dsub_rn(h2, h4) : nbody == 3 ? __dsub_rn(h3, h4) : __dadd_rn(h2, h3);
      if (st[nb] == 2) {
        ++num_p;
        outcome = __dsub_rn(outcome, zv);
      } else {
        zg[n_gkp++] = zv;
      }
    }
    double p_err = 0.0;
    int bits = 0;
    for (int i = 0; i < n_gkp; ++i) {
      p_err += z_err_cond_hom(2.0 * p.delta, zg[i]);
      int b;
      double f;
      gkp_bin(zg[i], b, f);
      bits += b;
    }
    p_err += z_err_cond_hom(2.0 * (double)(2 + num_p) * p.delta, outcome);
    p_err = fmax(fmin(p_err, 0.5), 0.0);
    int bp;
    double fp;
    gkp_bin(outcome, bp, fp);
    p.bit_val[idx] = (uint8_t)((bp + bits) & 1);
    p.p_phase[idx] = p_err;
    p.outcome[idx] = outcome;
  }
}

// weights from the precomputed probabilities + the bits of the syndrome qubits, packed
struct MacroWeightParams {
  int N, Qtot, bit_words;
  const int* adj_indptr;
  const int* adj_indices;
  const int* qubit_node;
  int method, integer;
  double multiplier;
  const uint8_t* red_state;  // [T][N]
  const uint8_t* bit_val;    // [T][N]
  const double* p_phase;     // [T][N]
  const double* outcome;     // [T][N]
  double* hom;               // [T][Qtot] <- effective homodyne value
  double* weights;           // [T][Qtot]
  uint32_t* bits;            // [T][bit_words]
};

__global__ void __launch_bounds__(256) k_macro_weights(MacroWeightParams p) {
  const int bpt = (p.Qtot + blockDim.x - 1) / blockDim.x;
  const int t = blockIdx.x / bpt;
  const int k = (blockIdx.x - t * bpt) * blockDim.x + threadIdx.x;
  const bool valid = k < p.Qtot;
  int bit = 0;
  if (valid) {
    const int i = p.qubit_node[k];
    const long long no = (long long)t * p.N;
    double w = 1.0;
    if (p.method != 0) {
      int pc = 0;
      for (int e = p.adj_indptr[i]; e < p.adj_indptr[i + 1]; ++e) pc += p.red_state[no + p.adj_indices[e]] == 2;
      double err;
      if (pc <= 1) {
        err = fmin(p.p_phase[no + i], 0.5);
        if (err == 0.0) err = FLOAT_MIN;
      } else {
        err = (pc == 2) ? 0.25 : ((pc == 3) ? (1.0 / 3.0) : 0.4);
      }
      w = p.integer ? rint(-p.multiplier * log(err)) : -log(err);
    }
    bit = p.bit_val[no + i];
    p.hom[(long long)t * p.Qtot + k] = p.outcome[no + i];
    p.weights[(long long)t * p.Qtot + k] = w;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, valid && bit);
  if ((threadIdx.x & 31) == 0 && (k >> 5) < p.bit_words) p.bits[(long long)t * p.bit_words + (k >> 5)] = word;
}

}  // namespace fpb
