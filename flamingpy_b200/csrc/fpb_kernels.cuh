// Device kernels of the Monte-Carlo EC trial path (sm_100a).
//
//   k_gen       Philox labels + Gaussian quadrature draws      (noise/cv.py:121-140,205,277-290)
//   k_noise     CZ entangle + GKP bin + blueprint weights       (cv/ops.py:65,84; cv/gkp.py:19-62,91-133;
//                                                               decoders/decoder.py:58-116)
//   k_syndrome  stabilizer parity + ordered odd-cube compaction (codes/stabilizer.py:43-55;
//                                                               codes/graphs/stabilizer_graph.py:304-306)
//   k_scan      per-trial offsets of the ragged outputs
//   k_paths     batched shortest paths between odd cubes and to the LOW/HIGH
//               connectors = matching-graph lengths            (decoders/mwpm/matching.py:125-173)
//
// File paths in parentheses are the reference code each kernel replaces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "philox.cuh"

namespace fpb {

constexpr double SQRT_PI = 1.7724538509055159;       // np.sqrt(np.pi)
constexpr double FLOAT_MIN = 2.2250738585072014e-308;  // sys.float_info.min (decoder.py:27)
constexpr int MAXC = 2;
constexpr int CURCAP = 64;  // per-warp capacity of the "current bucket" list

struct ComplexDev {
  int S, Q, qoff;      // cubes, qubits, offset of this complex's qubits in the global qubit list
  int par_words;       // ceil(S / 32)
  int n_slots;
  int n_low, n_high;
  const int* cube_qubit;      // [S][6] local qubit id or -1
  const uint16_t* ninfo;      // [S]
  const uint16_t* sbase;      // [S]
  const int* slot_qubit;      // [n_slots]
  const int* low_cube;
  const int* low_slot;
  const int* high_cube;
  const int* high_slot;
  int d_off[12];
  int s_off[12];
  // per-run outputs
  uint32_t* parity;      // [T][par_words]
  int* odd_count;        // [T]
  int* odd_fixed;        // [T][S]
  long long* odd_off;    // [T+1]
  long long* pair_off;   // [T+1]
  int* odd_ids;          // [total_odd]
  double* pair_len;      // [total_pairs]
  double* bnd_len;       // [total_odd]
  uint8_t* bnd_side;     // [total_odd]
};

struct NoiseParams {
  int N, Qtot, bit_words;
  const int* adj_indptr;
  const int* adj_indices;
  const int8_t* adj_vals;
  const int* qubit_node;  // [Qtot] node index per global syndrome qubit
  double delta_w;         // weight_options["delta"]
  int method, integer;
  double multiplier;
  const double* x;        // [T][2N]
  const uint8_t* state;   // [T][N]
  double* hom;            // [T][Qtot]
  double* weights;        // [T][Qtot]
  uint32_t* bits;         // [T][bit_words]
};

struct GenParams {
  int N;
  uint32_t seed_lo, seed_hi;
  long long first_trial;
  long long chain_len;  // <= 0: single chain starting at trial 0
  int fresh;
  double p_swap;
  uint32_t p_thresh;    // label draw < p_thresh  <=>  p-squeezed
  double sigma_q_p, sigma_q_gkp, sigma_p;  // float32-rounded standard deviations
  double* x;            // [T][2N]
  uint8_t* state;       // [T][N]
};

// ------------------------------------------------------------------------------------------
// k_gen: labels and draws
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool label_is_p(const GenParams& g, long long trial, uint32_t node) {
  Philox4 r = philox4x32_10((uint32_t)trial, (uint32_t)((unsigned long long)trial >> 32), node, SLOT_LABEL,
                            g.seed_lo, g.seed_hi);
  return r.x < g.p_thresh;
}

__global__ void __launch_bounds__(256) k_gen(GenParams g, long long n_trials) {
  long long total = n_trials * (long long)g.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long tl = idx / g.N;
    uint32_t node = (uint32_t)(idx - tl * g.N);
    long long trial = g.first_trial + tl;
    long long chain_start = (g.chain_len > 0) ? (trial / g.chain_len) * g.chain_len : 0;
    int st;
    if (g.p_swap >= 1.0) {
      st = 2;
    } else if (g.p_swap <= 0.0) {
      // no label draws at all; a reused layer alternates noisy / silent trials
      st = (g.fresh || (((trial - chain_start) & 1) == 0)) ? 1 : 0;
    } else if (label_is_p(g, trial, node)) {
      st = 2;
    } else if (g.fresh) {
      st = 1;
    } else {
      // run length of consecutive non-p draws ending at this trial, counted from the chain start;
      // GKP_t = V \ (p_t U GKP_{t-1})  =>  noisy iff the run length is odd
      int run = 1;
      for (long long s = trial - 1; s >= chain_start; --s) {
        if (label_is_p(g, s, node)) break;
        ++run;
      }
      st = (run & 1) ? 1 : 0;
    }
    Philox4 r = philox4x32_10((uint32_t)trial, (uint32_t)((unsigned long long)trial >> 32), node,
                              SLOT_QUADRATURE, g.seed_lo, g.seed_hi);
    double u1 = u01_53(r.x, r.y), u2 = u01_53(r.z, r.w);
    double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    double sq = (st == 2) ? g.sigma_q_p : (st == 1 ? g.sigma_q_gkp : 0.0);
    double sp = (st == 0) ? 0.0 : g.sigma_p;
    g.state[idx] = (uint8_t)st;
    double* xr = g.x + tl * 2ll * g.N;
    xr[node] = sq * (rad * cs);
    xr[g.N + node] = sp * (rad * sn);
  }
}

// ------------------------------------------------------------------------------------------
// k_noise: one thread per (trial, syndrome qubit)
// ------------------------------------------------------------------------------------------
// np.divmod(x, alpha) followed by the remainder > alpha/2 adjustment (cv/gkp.py:30-34).
// fmod is exact; the division is IEEE round-to-nearest like NumPy's.
__device__ __forceinline__ void gkp_bin(double x, int& bit, double& frac) {
  const double a = SQRT_PI;
  double mod = fmod(x, a);
  double div = __ddiv_rn(__dsub_rn(x, mod), a);
  if (mod != 0.0) {
    if (mod < 0.0) {
      mod = __dadd_rn(mod, a);
      div = __dsub_rn(div, 1.0);
    }
  } else {
    mod = 0.0;
  }
  double fl;
  if (div != 0.0) {
    fl = floor(div);
    if (__dsub_rn(div, fl) > 0.5) fl += 1.0;
  } else {
    fl = 0.0;
  }
  long long n = (long long)fl;
  if (mod > a * 0.5) {
    n += 1;
    frac = __dsub_rn(mod, a);
  } else {
    frac = mod;
  }
  bit = (int)(n & 1);  // Python's n % 2 for negative n as well
}

// Z_err_cond(var, hom) scalar branch with var_num = 10 (cv/gkp.py:116-133): both sums run
// over n = -10 .. 9.  Terms more than e^-80 below the largest one of their sum cannot change
// a double-precision sum and are skipped; the largest terms (|n| small) are always evaluated,
// so underflow of a whole sum to 0 happens exactly when it does in the reference.
__device__ __forceinline__ double z_err_cond(double var, double f) {
  double num = 0.0, den = 0.0;
  const double inv = 1.0 / var;
  // exponents of the dominant terms: numerator n in {-1, 0}, denominator n = 0
  double a0 = f - SQRT_PI, a1 = f + SQRT_PI;
  double num_top = fmin(a0 * a0, a1 * a1) * inv;
  double den_top = f * f * inv;
#pragma unroll 1
  for (int n = -10; n < 10; ++n) {
    double a = f - (double)(2 * n + 1) * SQRT_PI;
    double b = f - (double)n * SQRT_PI;
    double ea = (a * a) / var, eb = (b * b) / var;
    if (ea - num_top < 80.0) num += exp(-ea);
    if (eb - den_top < 80.0) den += exp(-eb);
  }
  return (den != 0.0) ? num / den : 0.0;
}

__global__ void __launch_bounds__(256) k_noise(NoiseParams p) {
  const int bpt = (p.Qtot + blockDim.x - 1) / blockDim.x;  // blocks per trial
  const int t = blockIdx.x / bpt;
  const int k = (blockIdx.x - t * bpt) * blockDim.x + threadIdx.x;
  const bool valid = k < p.Qtot;
  int bit = 0;
  if (valid) {
    const int i = p.qubit_node[k];
    const double* xr = p.x + (long long)t * 2 * p.N;
    const uint8_t* sr = p.state + (long long)t * p.N;
    const int e0 = p.adj_indptr[i], e1 = p.adj_indptr[i + 1];
    // p'_i = sum_j A_ij q_j (ascending j) + p_i, left to right like the sparse product
    double acc = 0.0;
    int pc = 0;
    for (int e = e0; e < e1; ++e) {
      const int j = p.adj_indices[e];
      acc = __dadd_rn(acc, __dmul_rn((double)p.adj_vals[e], xr[j]));
      pc += (sr[j] == 2);
    }
    const double hom = __dadd_rn(acc, xr[p.N + i]);
    double frac;
    gkp_bin(hom, bit, frac);
    double w;
    if (p.method == 0) {
      w = 1.0;
    } else {
      double err;
      if (pc <= 1) {
        err = z_err_cond((double)(e1 - e0 + 1) * p.delta_w, frac);
        err = fmin(err, 0.5);
        if (err == 0.0) err = FLOAT_MIN;
      } else {
        err = (pc == 2) ? 0.25 : ((pc == 3) ? (1.0 / 3.0) : 0.4);
      }
      w = p.integer ? rint(-p.multiplier * log(err)) : -log(err);
    }
    p.hom[(long long)t * p.Qtot + k] = hom;
    p.weights[(long long)t * p.Qtot + k] = w;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, valid && bit);
  if ((threadIdx.x & 31) == 0 && (k >> 5) < p.bit_words) p.bits[(long long)t * p.bit_words + (k >> 5)] = word;
}

// ------------------------------------------------------------------------------------------
// k_syndrome: one CTA per (trial, complex)
// ------------------------------------------------------------------------------------------
struct SyndromeParams {
  int n_complex, bit_words;
  const uint32_t* bits;
  ComplexDev cx[MAXC];
};

__global__ void __launch_bounds__(256) k_syndrome(SyndromeParams p) {
  const int t = blockIdx.x;
  const ComplexDev& c = p.cx[blockIdx.y];
  const uint32_t* bits = p.bits + (long long)t * p.bit_words;
  __shared__ int warp_cnt[8];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int s0 = 0; s0 < c.S; s0 += blockDim.x) {
    const int s = s0 + threadIdx.x;
    int par = 0;
    if (s < c.S) {
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const int q = c.cube_qubit[s * 6 + f];
        if (q >= 0) {
          const int g = c.qoff + q;
          par ^= (bits[g >> 5] >> (g & 31)) & 1u;
        }
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, par);
    if (lane == 0) {
      const int ws = s0 + wid * 32;  // first cube of this warp's 32-cube word
      if (ws < c.S) c.parity[(long long)t * c.par_words + (ws >> 5)] = m;
      warp_cnt[wid] = __popc(m);
    }
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < wid; ++w) before += warp_cnt[w];
    if (par) c.odd_fixed[(long long)t * c.S + before + __popc(m & ((1u << lane) - 1u))] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) c.odd_count[t] = base_s;
}

// ------------------------------------------------------------------------------------------
// k_scan: exclusive prefix sums over trials (single CTA)
// ------------------------------------------------------------------------------------------
struct ScanParams {
  int n_complex;
  int tasks_per_item;
  int has_boundary[MAXC];
  ComplexDev cx[MAXC];
  int* item_off;         // [T * n_complex + 1]
  long long* totals;     // [2 * MAXC + 1]: total_odd[c], total_pairs[c], total_items
};

__device__ __forceinline__ int n_tasks_of(int K, int has_boundary) {
  // K-1 single-source searches (the last odd cube needs none, matching.py:134) plus one
  // combined LOW+HIGH search when boundary points exist and the syndrome is non-trivial
  return K > 0 ? (K - 1) + (has_boundary ? 1 : 0) : 0;
}

__global__ void __launch_bounds__(1024) k_scan(ScanParams p, int T) {
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // three scans per complex (odd, pairs) and one over (trial, complex) items
  for (int which = 0; which < 2 * p.n_complex + 1; ++which) {
    const bool items = which == 2 * p.n_complex;
    const int c = which >> 1;
    const int n = items ? T * p.n_complex : T;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      long long v = 0;
      if (i < n) {
        if (items) {
          const int tt = i / p.n_complex, cc = i - tt * p.n_complex;
          const int nt = n_tasks_of(p.cx[cc].odd_count[tt], p.has_boundary[cc]);
          v = (nt + p.tasks_per_item - 1) / p.tasks_per_item;
        } else {
          const long long K = p.cx[c].odd_count[i];
          v = (which & 1) ? K * (K - 1) / 2 : K;
        }
      }
      long long inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        long long ws = wsum[lane];
        long long winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          long long y = __shfl_up_sync(0xffffffffu, winc, o);
          if (lane >= o) winc += y;
        }
        wsum[lane] = winc - ws;  // exclusive warp offsets
      }
      __syncthreads();
      const long long excl = carry_s + wsum[wid] + inc - v;
      if (i < n) {
        if (items) p.item_off[i] = (int)excl;
        else if (which & 1) p.cx[c].pair_off[i] = excl;
        else p.cx[c].odd_off[i] = excl;
      }
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const long long tot = carry_s;
      if (items) { p.item_off[n] = (int)tot; p.totals[2 * MAXC] = tot; }
      else if (which & 1) { p.cx[c].pair_off[n] = tot; p.totals[MAXC + c] = tot; }
      else { p.cx[c].odd_off[n] = tot; p.totals[c] = tot; }
    }
    __syncthreads();
  }
}

// compact the fixed-stride odd lists into the ragged contract layout
__global__ void __launch_bounds__(256) k_compact_odd(ComplexDev c, int T) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int K = c.odd_count[t];
  const long long off = c.odd_off[t];
  for (int i = threadIdx.x; i < K; i += blockDim.x) c.odd_ids[off + i] = c.odd_fixed[(long long)t * c.S + i];
}

// ------------------------------------------------------------------------------------------
// k_paths: persistent CTAs, one warp per shortest-path search
// ------------------------------------------------------------------------------------------
struct PathsParams {
  int n_complex, T, Qtot;
  int tasks_per_item;
  int S_max, slots_max;       // smem sizing (max over complexes)
  int warps;                  // warps per CTA
  double step_factor;         // bucket width in units of the trial's minimum weight
  const double* weights;      // [T][Qtot]
  const int* item_off;        // [T * n_complex + 1]
  int total_items;
  int* counter;               // work-queue head
  ComplexDev cx[MAXC];
  // path-toggle mode (fpb_paths_toggle); n_queries == 0 for the matching-graph run
  int n_queries;
  const int* q_trial;
  const int* q_cx;
  const int* q_src;
  const int* q_dst;
  uint32_t* flips;            // [T][bit_words]
  int bit_words;
};

struct WarpMem {
  double* dist;     // [S]
  uint16_t* list;   // [S]
  uint8_t* flag;    // [S]  bit0 dirty, bits1-3 incoming direction, 7<<1 = seeded from a connector
  uint16_t* cur;    // [CURCAP]
};

struct CtaMem {
  double* wslot;    // [slots_max]
  uint16_t* ninfo;  // [S_max]
  uint16_t* sbase;  // [S_max]
  int* offs;        // [24] d_off then s_off
};

__host__ __device__ inline size_t paths_cta_bytes(int S_max, int slots_max) {
  size_t b = (size_t)slots_max * 8;
  b += ((size_t)S_max * 2 + 7) / 8 * 8;
  b += ((size_t)S_max * 2 + 7) / 8 * 8;
  b += 24 * 4 + 32;
  return b;
}
__host__ __device__ inline size_t paths_warp_bytes(int S_max) {
  size_t b = (size_t)S_max * 8;
  b += ((size_t)S_max * 2 + 7) / 8 * 8;
  b += ((size_t)S_max + 7) / 8 * 8;
  b += CURCAP * 2;
  return b;
}

#define FPB_INF __longlong_as_double(0x7ff0000000000000ll)

// One bucketed label-correcting search on the cube grid, run by one warp out of shared memory.
// Buckets of width `step` are processed in increasing order; with step <= the minimum weight a
// node's distance is final when its bucket is reached (label setting), with a wider bucket the
// dirty flag re-queues nodes improved inside the bucket.  Pass A compacts the open list (drops
// settled nodes) and collects the dirty nodes of the current bucket; pass B relaxes them, one
// direction at a time so that the 32 targets of one instruction are distinct cubes and no
// atomics are needed.  If `stop_at` >= 0 the search ends once that cube is settled.
__device__ __forceinline__ void warp_search(const CtaMem& cm, const WarpMem& wm, int S, int n_seed,
                                            double step, int lane, int stop_at) {
  int Lc = n_seed;
  double t_lo = 0.0, t_hi = step;
  const uint32_t lt = (1u << lane) - 1u;
  while (true) {
    // ---------------- pass A ----------------
    int out = 0, ncur = 0;
    bool more = false;
    double minv = FPB_INF;
    for (int base = 0; base < Lc; base += 32) {
      const int i = base + lane;
      const bool valid = i < Lc;
      const int v = valid ? wm.list[i] : 0;
      const double dv = valid ? wm.dist[v] : FPB_INF;
      const bool keep = valid && !(dv < t_lo);
      const bool inb = keep && dv < t_hi;
      const bool dirty = inb && (wm.flag[v] & 1);
      if (keep && !inb) minv = fmin(minv, dv);
      __syncwarp();
      const uint32_t km = __ballot_sync(0xffffffffu, keep);
      if (keep) wm.list[out + __popc(km & lt)] = (uint16_t)v;
      out += __popc(km);
      const uint32_t tm = __ballot_sync(0xffffffffu, dirty);
      const int pos = ncur + __popc(tm & lt);
      if (dirty) {
        if (pos < CURCAP) {
          wm.cur[pos] = (uint16_t)v;
          wm.flag[v] &= 0xfe;
        }
      }
      const int tot = ncur + __popc(tm);
      if (tot > CURCAP) more = true;
      ncur = tot < CURCAP ? tot : CURCAP;
    }
    Lc = out;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) minv = fmin(minv, __shfl_xor_sync(0xffffffffu, minv, o));
    __syncwarp();
    if (ncur == 0) {
      if (!(minv < FPB_INF)) break;
      if (stop_at >= 0 && wm.dist[stop_at] < t_hi) break;  // target settled
      t_lo = t_hi;
      t_hi = (floor(minv / step) + 1.0) * step;
      if (!(t_hi > minv)) t_hi = minv + step;
      continue;
    }
    // ---------------- pass B ----------------
    bool again = more;
    double min_b = FPB_INF;
    for (int cb = 0; cb < ncur; cb += 32) {
      const bool valid = cb + lane < ncur;
      const int v = valid ? wm.cur[cb + lane] : 0;
      const double dv = valid ? wm.dist[v] : 0.0;
      const uint32_t info = valid ? cm.ninfo[v] : 0xaaau;  // all faces "none"
      const int sb = cm.sbase[v];
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        const uint32_t kind = (info >> (2 * d)) & 3u;
        const bool act = kind < 2u;
        const int u = v + cm.offs[2 * d + (kind & 1u)];
        const int slot = sb + cm.offs[12 + 2 * d + (kind & 1u)];
        bool isnew = false;
        if (act) {
          const double nd = dv + cm.wslot[slot];
          const double old = wm.dist[u];
          if (nd < old) {
            wm.dist[u] = nd;
            wm.flag[u] = (uint8_t)(1 | (d << 1));
            isnew = !(old < FPB_INF);
            if (nd < t_hi) again = true;
            else min_b = fmin(min_b, nd);
          }
        }
        const uint32_t nm = __ballot_sync(0xffffffffu, isnew);
        if (isnew) wm.list[Lc + __popc(nm & lt)] = (uint16_t)u;
        Lc += __popc(nm);
        __syncwarp();
      }
    }
    again = __any_sync(0xffffffffu, again);
    if (!again) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) min_b = fmin(min_b, __shfl_xor_sync(0xffffffffu, min_b, o));
      const double mn = fmin(minv, min_b);
      if (!(mn < FPB_INF)) break;
      if (stop_at >= 0 && wm.dist[stop_at] < t_hi) break;
      t_lo = t_hi;
      t_hi = (floor(mn / step) + 1.0) * step;
      if (!(t_hi > mn)) t_hi = mn + step;
    }
  }
}

__device__ __forceinline__ void warp_clear(const WarpMem& wm, int S, int lane) {
  for (int v = lane; v < S; v += 32) {
    wm.dist[v] = FPB_INF;
    wm.flag[v] = 0;
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_seed_connector(const CtaMem& cm, const WarpMem& wm, int n,
                                                    const int* cubes, const int* slots, int lane) {
  for (int i = lane; i < n; i += 32) {
    const int c = cubes[i];
    wm.dist[c] = cm.wslot[slots[i]];
    wm.flag[c] = (uint8_t)(1 | (7 << 1));
    wm.list[i] = (uint16_t)c;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(512) k_paths(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int item_s;
  __shared__ double red_s[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nthreads = blockDim.x;

  CtaMem cm;
  unsigned char* ptr = smem;
  cm.wslot = (double*)ptr; ptr += (size_t)p.slots_max * 8;
  cm.ninfo = (uint16_t*)ptr; ptr += ((size_t)p.S_max * 2 + 7) / 8 * 8;
  cm.sbase = (uint16_t*)ptr; ptr += ((size_t)p.S_max * 2 + 7) / 8 * 8;
  cm.offs = (int*)ptr; ptr += 24 * 4 + 32;
  const size_t wbytes = paths_warp_bytes(p.S_max);
  unsigned char* wp = ptr + (size_t)wid * wbytes;
  WarpMem wm;
  wm.dist = (double*)wp; wp += (size_t)p.S_max * 8;
  wm.list = (uint16_t*)wp; wp += ((size_t)p.S_max * 2 + 7) / 8 * 8;
  wm.flag = (uint8_t*)wp; wp += ((size_t)p.S_max + 7) / 8 * 8;
  wm.cur = (uint16_t*)wp;

  int staged_tc = -1, staged_c = -1;
  double step = 1.0;

  while (true) {
    if (threadIdx.x == 0) item_s = atomicAdd(p.counter, 1);
    __syncthreads();
    const int item = item_s;
    __syncthreads();
    if (item >= p.total_items) break;
    // (trial, complex) of this item: last tc with item_off[tc] <= item
    int lo = 0, hi = p.T * p.n_complex;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.item_off[mid] <= item) lo = mid; else hi = mid;
    }
    const int tc = lo;
    const int t = tc / p.n_complex, ci = tc - t * p.n_complex;
    const ComplexDev& c = p.cx[ci];
    const int chunk = item - p.item_off[tc];
    const int S = c.S;

    if (staged_c != ci) {
      for (int v = threadIdx.x; v < S; v += nthreads) {
        cm.ninfo[v] = c.ninfo[v];
        cm.sbase[v] = c.sbase[v];
      }
      if (threadIdx.x < 12) cm.offs[threadIdx.x] = c.d_off[threadIdx.x];
      else if (threadIdx.x < 24) cm.offs[threadIdx.x] = c.s_off[threadIdx.x - 12];
      staged_c = ci;
    }
    if (staged_tc != tc) {
      const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
      double mn = FPB_INF;
      for (int s = threadIdx.x; s < c.n_slots; s += nthreads) {
        const int q = c.slot_qubit[s];
        const double ws = (q >= 0) ? w[q] : FPB_INF;
        cm.wslot[s] = ws;
        mn = fmin(mn, ws);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) red_s[wid] = mn;
      __syncthreads();
      mn = FPB_INF;
      for (int w2 = 0; w2 < (nthreads >> 5); ++w2) mn = fmin(mn, red_s[w2]);
      step = (mn > 0.0 && mn < FPB_INF) ? mn * p.step_factor : 1.0;
      staged_tc = tc;
    }
    __syncthreads();

    const int K = c.odd_count[t];
    const int* odd = c.odd_fixed + (long long)t * S;
    const int has_bnd = c.n_low > 0;
    const int n_tasks = n_tasks_of(K, has_bnd);
    const long long poff = c.pair_off[t];
    const long long ooff = c.odd_off[t];
    const int task_end = min(n_tasks, (chunk + 1) * p.tasks_per_item);
    for (int task = chunk * p.tasks_per_item + wid; task < task_end; task += (nthreads >> 5)) {
      warp_clear(wm, S, lane);
      if (task < K - 1) {
        // single-source search from odd cube `task`; row `task` of the packed upper triangle
        const int src = odd[task];
        if (lane == 0) {
          wm.dist[src] = 0.0;
          wm.flag[src] = 1;
          wm.list[0] = (uint16_t)src;
        }
        __syncwarp();
        warp_search(cm, wm, S, 1, step, lane, -1);
        const long long row = poff + (long long)task * K - (long long)task * (task + 1) / 2;
        for (int b = task + 1 + lane; b < K; b += 32) c.pair_len[row + (b - task - 1)] = wm.dist[odd[b]];
      } else {
        // LOW then HIGH connector searches; keep the nearer one (ties -> low, matching.py:149-150)
        warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, lane);
        warp_search(cm, wm, S, c.n_low, step, lane, -1);
        for (int b = lane; b < K; b += 32) c.bnd_len[ooff + b] = wm.dist[odd[b]];
        __syncwarp();
        warp_clear(wm, S, lane);
        warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, lane);
        warp_search(cm, wm, S, c.n_high, step, lane, -1);
        for (int b = lane; b < K; b += 32) {
          const double lo_d = c.bnd_len[ooff + b];
          const double hi_d = wm.dist[odd[b]];
          const bool high = hi_d < lo_d;
          c.bnd_len[ooff + b] = high ? hi_d : lo_d;
          c.bnd_side[ooff + b] = high ? 1 : 0;
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// k_path_toggle: one warp per query; search with early stop, then walk the tree back and
// XOR-toggle the crossed qubits (mwpm/algos.py:88-95)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_path_toggle(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int item_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  // every warp keeps its own copy of the weights here (queries of one CTA mix trials), so the
  // layout is per warp: wslot, dist, list, flag, cur; ninfo/sbase/offs are read from global.
  const size_t per_warp = (size_t)p.slots_max * 8 + paths_warp_bytes(p.S_max);
  unsigned char* wp = smem + (size_t)wid * per_warp;
  double* wslot = (double*)wp; wp += (size_t)p.slots_max * 8;
  WarpMem wm;
  wm.dist = (double*)wp; wp += (size_t)p.S_max * 8;
  wm.list = (uint16_t*)wp; wp += ((size_t)p.S_max * 2 + 7) / 8 * 8;
  wm.flag = (uint8_t*)wp; wp += ((size_t)p.S_max + 7) / 8 * 8;
  wm.cur = (uint16_t*)wp;
  __shared__ int offs_s[MAXC][24];
  if (threadIdx.x < 24 * p.n_complex) {
    const int ci = threadIdx.x / 24, j = threadIdx.x % 24;
    offs_s[ci][j] = j < 12 ? p.cx[ci].d_off[j] : p.cx[ci].s_off[j - 12];
  }
  __syncthreads();
  (void)item_s;
  for (int qi = blockIdx.x * nwarps + wid; qi < p.n_queries; qi += gridDim.x * nwarps) {
    const int t = p.q_trial[qi], ci = p.q_cx[qi];
    const ComplexDev& c = p.cx[ci];
    const int S = c.S;
    const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
    double mn = FPB_INF;
    for (int s = lane; s < c.n_slots; s += 32) {
      const int q = c.slot_qubit[s];
      const double ws = (q >= 0) ? w[q] : FPB_INF;
      wslot[s] = ws;
      mn = fmin(mn, ws);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    const double step = (mn > 0.0 && mn < FPB_INF) ? mn * p.step_factor : 1.0;
    CtaMem cm;
    cm.wslot = wslot;
    cm.ninfo = const_cast<uint16_t*>(c.ninfo);
    cm.sbase = const_cast<uint16_t*>(c.sbase);
    cm.offs = offs_s[ci];
    const int* odd = c.odd_fixed + (long long)t * S;
    const int src = odd[p.q_src[qi]];
    uint32_t* flips = p.flips + (long long)t * p.bit_words;
    warp_clear(wm, S, lane);
    int walk_from;
    if (p.q_dst[qi] >= 0) {
      const int dst = odd[p.q_dst[qi]];
      if (lane == 0) {
        wm.dist[src] = 0.0;
        wm.flag[src] = 1;
        wm.list[0] = (uint16_t)src;
      }
      __syncwarp();
      warp_search(cm, wm, S, 1, step, lane, dst);
      walk_from = dst;
    } else {
      // to the nearer connector, as recorded by the matching-graph run
      const int side = c.bnd_side[c.odd_off[t] + p.q_src[qi]];
      if (side == 0) {
        warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, lane);
        warp_search(cm, wm, S, c.n_low, step, lane, src);
      } else {
        warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, lane);
        warp_search(cm, wm, S, c.n_high, step, lane, src);
      }
      walk_from = src;
    }
    __syncwarp();
    if (lane == 0) {
      int u = walk_from;
      const bool pair_query = p.q_dst[qi] >= 0;
      for (int guard = 0; guard < S + 2; ++guard) {
        if (pair_query && u == src) break;
        const int din = (wm.flag[u] >> 1) & 7;
        if (din == 7) {
          // seeded from a connector: toggle the boundary qubit and stop
          const int side = c.bnd_side[c.odd_off[t] + p.q_src[qi]];
          const int n = side ? c.n_high : c.n_low;
          const int* cubes = side ? c.high_cube : c.low_cube;
          const int* slots = side ? c.high_slot : c.low_slot;
          for (int i = 0; i < n; ++i)
            if (cubes[i] == u) {
              const int g = c.qoff + c.slot_qubit[slots[i]];
              atomicXor(&flips[g >> 5], 1u << (g & 31));
              break;
            }
          break;
        }
        // u was reached from v by direction din; step back along din^1
        const int back = din ^ 1;
        const uint32_t kb = (c.ninfo[u] >> (2 * back)) & 3u;
        const int v = u + offs_s[ci][2 * back + (kb & 1u)];
        const uint32_t kf = (c.ninfo[v] >> (2 * din)) & 3u;
        const int slot = c.sbase[v] + offs_s[ci][12 + 2 * din + (kf & 1u)];
        const int g = c.qoff + c.slot_qubit[slot];
        atomicXor(&flips[g >> 5], 1u << (g & 31));
        u = v;
      }
    }
    __syncwarp();
  }
}

}  // namespace fpb
