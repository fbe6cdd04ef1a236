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
  uint16_t* word_rank;   // [T][par_words] odd cubes before each 32-cube parity word (static-weight gather), or nullptr
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
// k_iid: i.i.d. Z errors (flamingpy/noise/iid.py:37-50): bit_val = (uniform < p) per node, drawn
// from the Philox stream of (trial, node, SLOT_IID) for the syndrome qubits, packed LSB-first;
// k_fill: constant weights for the trials of a bits-only run (uniform weights, decoder.py:115-116)
// ------------------------------------------------------------------------------------------
struct IidParams {
  int Qtot, bit_words;
  uint32_t seed_lo, seed_hi;
  long long first_trial;
  double p_err;
  const int* qubit_node;  // [Qtot]
  uint32_t* bits;         // [T][bit_words]
};
__global__ void __launch_bounds__(256) k_iid(IidParams g, long long n_trials) {
  const long long total = n_trials * g.bit_words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / g.bit_words;
    const int wd = (int)(i - t * g.bit_words);
    const unsigned long long trial = (unsigned long long)(g.first_trial + t);
    uint32_t word = 0;
    const int q0 = wd * 32;
    const int nq = min(32, g.Qtot - q0);
    for (int b = 0; b < nq; ++b) {
      const Philox4 r = philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)g.qubit_node[q0 + b], SLOT_IID,
                                      g.seed_lo, g.seed_hi);
      word |= (u01_53(r.x, r.y) < g.p_err ? 1u : 0u) << b;
    }
    g.bits[i] = word;
  }
}
__global__ void __launch_bounds__(256) k_fill(double* dst, long long n, double value) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = value;
}

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
    if (c.word_rank && lane == 0 && s0 + wid * 32 < c.S)
      c.word_rank[(long long)t * c.par_words + ((s0 + wid * 32) >> 5)] = (uint16_t)before;
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
  __shared__ int kmax_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // three scans per complex (odd, pairs) and one over (trial, complex) items
  for (int which = 0; which < 2 * p.n_complex + 1; ++which) {
    const bool items = which == 2 * p.n_complex;
    const int c = which >> 1;
    const int n = items ? T * p.n_complex : T;
    if (threadIdx.x == 0) {
      carry_s = 0;
      kmax_s = 0;
    }
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
          if (!(which & 1)) atomicMax(&kmax_s, (int)K);
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
      else { p.cx[c].odd_off[n] = tot; p.totals[c] = tot; p.totals[2 * MAXC + 1 + c] = kmax_s; }
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
  double step_factor;         // bucket width in units of the trial's mean weight
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
  int* path_qubits;           // fpb_paths_list: [n_queries][path_cap] local qubit ids in walk order (else nullptr)
  int* path_len;              // [n_queries]
  int path_cap;
  double* wscratch;           // [grid][slots_max] slot-ordered weights (global-weights variant only)
};

constexpr int NRING = 32;       // ring slots: bucket index mod 32 is kept in one byte per cube
constexpr int STAGE_CAP = 512;  // per-warp staging list of the cubes of the current bucket
constexpr int MAX_CHUNKS = 32;  // 16-cube chunks per lane in one scan: S <= 16 * 32 * 32 = 16384 (template NCH)

struct WarpMem {
  double* dist;     // [S]
  uint8_t* bkt;     // [ceil16(S)] ring slot of the cube's pending entry, 0xFF none; 16-byte aligned
  uint16_t* stage;  // [STAGE_CAP]
  uint8_t* pred;    // [S] incoming direction (path-toggle kernel only)
  int* ilen;        // [S] sum of int(weight) along the shortest path (FPB_LENGTH_RX_TRUNC only)
  int sb;           // 0: bkt[c] is cube c's byte; > 0: residue-major rows, byte (c & 15) * sb + (c >> 4)
};

struct CtaMem {
  const double* wslot;    // [n_slots] weights of the trial in face-slot order
  const uint16_t* ninfo;  // [S]
  const uint16_t* sbase;  // [S]
};

// direction tables of one complex held in registers (uniform across the warp)
struct DirTab {
  int dreg[6], dwrap[6], sreg[6], swrap[6];
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) / 16 * 16; }
__host__ __device__ inline size_t paths_cta_bytes(int S_max, int slots_max) {
  return align16((size_t)slots_max * 8) + align16((size_t)S_max * 2) + align16((size_t)S_max * 2);
}
// staging-list entries per search: the wide-team kernels of very large lattices stage more cubes per scan
__host__ __device__ inline int paths_stage_entries(int S_max) { return S_max > 6144 ? 4 * STAGE_CAP : STAGE_CAP; }
// bucket bytes of one search: room for the residue-major layout of the team kernel (16 rows of
// paths_bkt_row(S) bytes, one row per cube index mod 16; >= the linear layout's align16(S))
__host__ __device__ inline int paths_bkt_row(int S) { return (((S + 15) >> 4) + 15) & ~15; }
__host__ __device__ inline size_t paths_bkt_bytes(int S_max) { return (size_t)16 * paths_bkt_row(S_max); }
__host__ __device__ inline size_t paths_warp_bytes(int S_max, bool pred) {
  return align16((size_t)S_max * 8) + paths_bkt_bytes(S_max) + align16((size_t)paths_stage_entries(S_max) * 2) +
         (pred ? align16((size_t)S_max) : 0);
}

#define FPB_INF __longlong_as_double(0x7ff0000000000000ll)

__device__ __forceinline__ WarpMem carve_warp(unsigned char* wp, int S_max, bool pred) {
  WarpMem wm;
  wm.dist = (double*)wp; wp += align16((size_t)S_max * 8);
  wm.bkt = (uint8_t*)wp; wp += paths_bkt_bytes(S_max);
  wm.sb = 0;
  wm.stage = (uint16_t*)wp; wp += align16((size_t)paths_stage_entries(S_max) * 2);
  wm.pred = pred ? (uint8_t*)wp : nullptr;
  wm.ilen = nullptr;
  return wm;
}

__device__ __forceinline__ DirTab load_dirtab(const ComplexDev& c) {
  DirTab t;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    t.dreg[d] = c.d_off[2 * d];
    t.dwrap[d] = c.d_off[2 * d + 1];
    t.sreg[d] = c.s_off[2 * d];
    t.swrap[d] = c.s_off[2 * d + 1];
  }
  return t;
}

__device__ __forceinline__ int bucket_of(double x, double inv) { return __double2int_rd(x * inv); }

__device__ __forceinline__ void warp_clear(const WarpMem& wm, int S, int lane) {
  for (int v = lane; v < S; v += 32) wm.dist[v] = FPB_INF;
  const int n16 = (S + 15) >> 4;
  uint4* b4 = reinterpret_cast<uint4*>(wm.bkt);
  for (int i = lane; i < n16; i += 32) b4[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  __syncwarp();
}

// File cube c with tentative distance d0 before a search starts (bucket 0 is the ring head).
// Returns the lane's contribution to the `present` mask.
__device__ __forceinline__ uint32_t seed_cube(const WarpMem& wm, int c, double d0, double inv, uint8_t pred_code) {
  wm.dist[c] = d0;
  if (wm.ilen) wm.ilen[c] = (int)d0;  // int(weight) of the connector's boundary edge (0 for a source cube)
  int bi = bucket_of(d0, inv);
  bi = bi < NRING - 1 ? bi : NRING - 1;
  wm.bkt[wm.sb ? (c & 15) * wm.sb + (c >> 4) : c] = (uint8_t)bi;
  if (wm.pred) wm.pred[c] = pred_code;
  return 1u << bi;
}

// One bucketed shortest-path search on the cube grid, run by one warp out of shared memory.
//
// Every cube with a pending (tentative, not yet relaxed) distance is filed under the ring slot
// (bucket index mod 32) of that distance in one byte per cube, so a cube has exactly one entry
// and a later improvement simply overwrites it.  Buckets of width `step` are processed in
// increasing order: a dense 16-bytes-per-load scan stages the cubes of the current slot, then
// they are relaxed 32 at a time.  All six neighbour distances and weights of a batch are loaded
// up front; improvements are stored direction by direction (within one direction the 32
// targets are distinct cubes, across directions the last store wins) and a verify round
// re-stores any smaller value that was overwritten, so no atomics are needed.  With
// step <= the minimum weight a cube is final when its bucket comes up (label setting); a wider
// bucket re-files cubes improved inside it.  Cubes more than 31 buckets ahead are parked in the
// farthest slot and moved on when that slot comes up.  If `stop_at` >= 0 the search ends as
// soon as that cube is settled.
// ---- explicit shared-memory accessors (32-bit shared-window addresses; volatile keeps the
// program order of the search's loads and stores, which the conflict resolution relies on) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#ifdef FPB_EMU  // host build of tests/emu (CPU SIMT emulation for the parity tests): plain volatile accesses
__device__ __forceinline__ double lds_f64(uint32_t a) { return *fpb_emu::shared_ptr<volatile double>(a); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { *fpb_emu::shared_ptr<volatile double>(a) = v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { return *fpb_emu::shared_ptr<volatile uint16_t>(a); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { *fpb_emu::shared_ptr<volatile uint16_t>(a) = (uint16_t)v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { *fpb_emu::shared_ptr<volatile uint8_t>(a) = (uint8_t)v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  volatile uint32_t* q = fpb_emu::shared_ptr<volatile uint32_t>(a);
  return make_uint4(q[0], q[1], q[2], q[3]);
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
  volatile uint32_t* q = fpb_emu::shared_ptr<volatile uint32_t>(a);
  q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { return *fpb_emu::shared_ptr<volatile uint32_t>(a); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { *fpb_emu::shared_ptr<volatile uint32_t>(a) = v; }
#else
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v)); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
#endif

// bytes of `word` equal to the byte replicated in r4: returns 0x80 in each matching byte
// (exact: no carries cross a byte boundary)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t word, uint32_t r4) {
  const uint32_t x = word ^ r4;
  const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(t | x | 0x7f7f7f7fu);
}
// gather bits 7,15,23,31 into bits 0..3
__device__ __forceinline__ uint32_t movemask4(uint32_t z) { return (z * 0x00204081u) >> 28; }

// ILEN (FPB_LENGTH_RX_TRUNC): next to the float distance every cube carries the sum of int(weight) along
// the path that gave it that distance - what the reference's rustworkx backend reports as the path
// weight (codes/graphs/stabilizer_graph.py:486-500: Dijkstra on the float weights, then
// weight += int(edge weight) along the returned path).  The owner of a target's final distance writes it.
template <bool PRED, int NCH, int NV, bool WG, bool LEAN = false, bool ILEN = false>
__device__ __forceinline__ void warp_search(const CtaMem& cm, const WarpMem& wm, const DirTab& dt, int S,
                                            double step, uint32_t present, int lane, int stop_at) {
  const double inv = 1.0 / step;
  const int n16 = (S + 15) >> 4;
  const uint32_t dist_a = smem_addr(wm.dist), bkt_a = smem_addr(wm.bkt), stage_a = smem_addr(wm.stage);
  const uint32_t pred_a = PRED ? smem_addr(wm.pred) : 0u;
  const uint32_t ilen_a = ILEN ? smem_addr(wm.ilen) : 0u;
  // WG: the trial's weights (and the face tables) live in global memory instead of shared memory
  const uint32_t wslot_a = WG ? 0u : smem_addr(cm.wslot);
  int cur = 0;
  while (present) {
    {
      const uint32_t rot = __funnelshift_r(present, present, cur & (NRING - 1));
      cur += __ffs(rot) - 1;
    }
    const int r = cur & (NRING - 1);
    const int far = cur + NRING - 1;
    while ((present >> r) & 1u) {
      present &= ~(1u << r);
      // ---------------- scan: stage the cubes filed under slot r ----------------
      // matching bytes are cleared (0xFF) in the same pass, one 16-byte store per chunk
      uint32_t masks[NCH];
      int cnt = 0;
      const uint32_t r4 = (uint32_t)r * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int ch = j * 32 + lane;
        uint32_t m = 0;
        if (ch < n16) {
          uint4 q = lds_v4(bkt_a + (ch << 4));
          const uint32_t z0 = eq_bytes(q.x, r4), z1 = eq_bytes(q.y, r4), z2 = eq_bytes(q.z, r4), z3 = eq_bytes(q.w, r4);
          m = movemask4(z0) | (movemask4(z1) << 4) | (movemask4(z2) << 8) | (movemask4(z3) << 12);
          if (m) {
            q.x |= (z0 >> 7) * 0xffu;
            q.y |= (z1 >> 7) * 0xffu;
            q.z |= (z2 >> 7) * 0xffu;
            q.w |= (z3 >> 7) * 0xffu;
            sts_v4(bkt_a + (ch << 4), q);
          }
        }
        masks[j] = m;
        cnt += __popc(m);
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = incl - cnt;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        uint32_t m = masks[j];
        const int base = (j * 32 + lane) << 4;
        while (m) {
          const int cube = base + __ffs(m) - 1;
          m &= m - 1;
          if (pos < STAGE_CAP) sts_u16(stage_a + 2 * pos, cube);
          else sts_u8(bkt_a + cube, r);  // staging list full: leave it filed for the next scan
          ++pos;
        }
      }
      if (total > STAGE_CAP) present |= 1u << r;
      const int nst = total < STAGE_CAP ? total : STAGE_CAP;
      __syncwarp();
      // ---------------- relax the staged cubes, 32 * NV at a time ----------------
      uint32_t lp = 0;  // slots this lane filed cubes under
      for (int cb = 0; cb < nst; cb += 32 * NV) {
        uint32_t ua[NV][6];  // shared address of the neighbour's distance
        double nd[NV][6];
        bool imp[NV][6];
        int il[NV][6];  // ILEN: int length of the candidate path
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int idx = cb + k * 32 + lane;
          const bool valid = idx < nst;
          const int v = valid ? (int)lds_u16(stage_a + 2 * idx) : 0;
          const uint32_t va = dist_a + 8u * v;
          const double dv = lds_f64(va);
          const int iv = ILEN ? (int)lds_u32(ilen_a + 4u * v) : 0;
          const int bt = bucket_of(dv, inv);
          if (valid && bt > cur) {  // a parked cube: move it towards its real bucket
            const int s2 = (bt < far ? bt : far) & (NRING - 1);
            sts_u8(bkt_a + v, s2);
            lp |= 1u << s2;
          }
          const bool act = valid && bt <= cur;
          const uint32_t info = act ? (uint32_t)cm.ninfo[v] : 0xaaau;  // inactive lanes: every face "none"
          const int sb = cm.sbase[v];
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const uint32_t kind = (info >> (2 * d)) & 3u;
            const bool a = kind < 2u;
            // absent faces read harmless in-range addresses (their own cube, slot 0)
            ua[k][d] = a ? va + 8 * (kind ? dt.dwrap[d] : dt.dreg[d]) : va;
            const int slot = a ? sb + (kind ? dt.swrap[d] : dt.sreg[d]) : 0;
            if (LEAN) {
              // many resident warps, shared-memory pipe is the limiter: only existing faces load, and
              // the weight is fetched only if the neighbour is farther from the source than this cube
              const double old = a ? lds_f64(ua[k][d]) : 0.0;
              const bool cand = a && old > dv;
              nd[k][d] = cand ? dv + (WG ? cm.wslot[slot] : lds_f64(wslot_a + 8u * slot)) : 0.0;
              imp[k][d] = cand && nd[k][d] < old;
            } else {
              const double wgt = WG ? cm.wslot[slot] : lds_f64(wslot_a + 8u * slot);
              nd[k][d] = dv + wgt;
              imp[k][d] = a && nd[k][d] < lds_f64(ua[k][d]);
              if (ILEN) il[k][d] = iv + (a ? __double2int_rz(fmin(wgt, 2.0e9)) : 0);
            }
          }
        }
        // ILEN: a staged cube's (distance, int length) pair must be read before any lane of this round stores,
        // or a candidate could combine a neighbour's new distance with the old int length
        if (ILEN) __syncwarp();
#pragma unroll
        for (int d = 0; d < 6; ++d)
#pragma unroll
          for (int k = 0; k < NV; ++k)
            if (imp[k][d]) sts_f64(ua[k][d], nd[k][d]);
        __syncwarp();
        double fin[NV][6];
        bool redo;
        do {
          redo = false;
#pragma unroll
          for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int d = 0; d < 6; ++d) fin[k][d] = (!LEAN || imp[k][d]) ? lds_f64(ua[k][d]) : 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d)
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const bool lost = imp[k][d] && fin[k][d] > nd[k][d];
              if (lost) sts_f64(ua[k][d], nd[k][d]);
              redo |= lost;
            }
          redo = __any_sync(0xffffffffu, redo);
        } while (redo);
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const int bi = bucket_of(fin[k][d], inv);
            const int s2 = (bi < far ? bi : far) & (NRING - 1);
            const uint32_t ui = (ua[k][d] - dist_a) >> 3;
            if (imp[k][d]) sts_u8(bkt_a + ui, s2);
            lp |= imp[k][d] ? (1u << s2) : 0u;
            if (PRED) {
              if (imp[k][d] && fin[k][d] == nd[k][d]) sts_u8(pred_a + ui, d);
            }
            if (ILEN) {
              if (imp[k][d] && fin[k][d] == nd[k][d]) sts_u32(ilen_a + 4u * ui, (uint32_t)il[k][d]);
            }
          }
        __syncwarp();
      }
      present |= __reduce_or_sync(0xffffffffu, lp);
    }
#ifndef FPB_NO_EARLY_STOP
    if (stop_at >= 0) {
      if (bucket_of(wm.dist[stop_at], inv) <= cur) break;  // the target is settled
    }
#endif
  }
}

__device__ __forceinline__ uint32_t warp_seed_connector(const CtaMem& cm, const WarpMem& wm, int n,
                                                        const int* cubes, const int* slots, double inv,
                                                        int lane) {
  uint32_t lp = 0;
  for (int i = lane; i < n; i += 32) lp |= seed_cube(wm, cubes[i], cm.wslot[slots[i]], inv, 7);
  __syncwarp();
  return __reduce_or_sync(0xffffffffu, lp);
}

// bucket width = factor x the trial's mean edge weight (any positive width is correct; about
// 1.5 mean weights measured fastest on B200 across the BASELINE configurations)
__device__ __forceinline__ double bucket_step(double wmean, double factor) {
  return (wmean > 0.0 && wmean < FPB_INF) ? wmean * factor : 1.0;
}

template <int NCH, int NV, int MAXT, int MINB, bool WG, bool ILEN = false>
__global__ void __launch_bounds__(MAXT, MINB) k_paths(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int item_s;
  __shared__ double red_s[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nthreads = blockDim.x;

  // shared variant: weights + face tables per CTA in shared memory, then the per-warp state;
  // WG variant (large lattices): weights in a per-CTA global scratch row, tables read in place
  unsigned char* ptr = smem;
  double* wslot;
  uint16_t* ninfo_s = nullptr;
  uint16_t* sbase_s = nullptr;
  if (WG) {
    wslot = p.wscratch + (size_t)blockIdx.x * p.slots_max;
  } else {
    wslot = (double*)ptr; ptr += align16((size_t)p.slots_max * 8);
    ninfo_s = (uint16_t*)ptr; ptr += align16((size_t)p.S_max * 2);
    sbase_s = (uint16_t*)ptr; ptr += align16((size_t)p.S_max * 2);
  }
  CtaMem cm;
  cm.wslot = wslot; cm.ninfo = ninfo_s; cm.sbase = sbase_s;
  const size_t warp_b = paths_warp_bytes(p.S_max, false) + (ILEN ? align16((size_t)p.S_max * 4) : 0);
  WarpMem wm = carve_warp(ptr + (size_t)wid * warp_b, p.S_max, false);
  if (ILEN) wm.ilen = (int*)(ptr + (size_t)wid * warp_b + paths_warp_bytes(p.S_max, false));

  int staged_tc = -1, staged_c = -1;
  double step = 1.0;
  DirTab dt = load_dirtab(p.cx[0]);

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
      if (WG) {
        cm.ninfo = c.ninfo;
        cm.sbase = c.sbase;
      } else {
        for (int v = threadIdx.x; v < S; v += nthreads) {
          ninfo_s[v] = c.ninfo[v];
          sbase_s[v] = c.sbase[v];
        }
      }
      dt = load_dirtab(c);
      staged_c = ci;
    }
    if (staged_tc != tc) {
      const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
      double sum = 0.0;
      for (int s = threadIdx.x; s < c.n_slots; s += nthreads) {
        const int q = c.slot_qubit[s];
        const double ws = (q >= 0) ? w[q] : FPB_INF;
        wslot[s] = ws;
        sum += (q >= 0) ? ws : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) red_s[wid] = sum;
      __syncthreads();
      sum = 0.0;
      for (int w2 = 0; w2 < (nthreads >> 5); ++w2) sum += red_s[w2];
      step = bucket_step(sum / (double)c.Q, p.step_factor);
      staged_tc = tc;
      if (WG) __threadfence_block();
    }
    __syncthreads();

    const double inv = 1.0 / step;
    const int K = c.odd_count[t];
    const int* odd = c.odd_fixed + (long long)t * S;
    const int has_bnd = c.n_low > 0;
    const int n_tasks = n_tasks_of(K, has_bnd);
    const long long poff = c.pair_off[t];
    const long long ooff = c.odd_off[t];
    const int task_end = min(n_tasks, (chunk + 1) * p.tasks_per_item);
    for (int task = chunk * p.tasks_per_item + wid; task < task_end; task += (nthreads >> 5)) {
      // task < K-1: single-source search from odd cube `task` (row `task` of the packed upper
      // triangle); task == K-1: LOW then HIGH connector searches, keeping the nearer one
      // (ties -> low, matching.py:149-150)
      const bool pair_task = task < K - 1;
      const int n_phase = pair_task ? 1 : 2;
      for (int ph = 0; ph < n_phase; ++ph) {
        warp_clear(wm, S, lane);
        uint32_t present = 0;
        if (pair_task) {
          if (lane == 0) present = seed_cube(wm, odd[task], 0.0, inv, 0);
          present = __shfl_sync(0xffffffffu, present, 0);
          __syncwarp();
        } else if (ph == 0) {
          present = warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, inv, lane);
        } else {
          present = warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, inv, lane);
        }
        warp_search<false, NCH, NV, WG, (MINB > 1) && !ILEN, ILEN>(cm, wm, dt, S, step, present, lane, -1);
        // ILEN: the reported length is the int sum; the nearer connector is chosen on the int sums, too
        // (mwpm/matching.py:149-150 on the rustworkx lengths)
        if (pair_task) {
          const long long row = poff + (long long)task * K - (long long)task * (task + 1) / 2;
          for (int b = task + 1 + lane; b < K; b += 32)
            c.pair_len[row + (b - task - 1)] = ILEN ? (double)wm.ilen[odd[b]] : wm.dist[odd[b]];
        } else if (ph == 0) {
          for (int b = lane; b < K; b += 32) c.bnd_len[ooff + b] = ILEN ? (double)wm.ilen[odd[b]] : wm.dist[odd[b]];
        } else {
          for (int b = lane; b < K; b += 32) {
            const double lo_d = c.bnd_len[ooff + b];
            const double hi_d = ILEN ? (double)wm.ilen[odd[b]] : wm.dist[odd[b]];
            const bool high = hi_d < lo_d;
            c.bnd_len[ooff + b] = high ? hi_d : lo_d;
            c.bnd_side[ooff + b] = high ? 1 : 0;
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Two-warp teams: the same search as warp_search, run by a pair of warps that share one
// distance / bucket / staging state.  Warp `half` scans its half of the bucket chunks and
// relaxes directions 3*half .. 3*half+2 of every staged cube, so each warp executes about half
// of the instruction stream and an SM holds twice as many warps for the same shared memory.
// The store-then-verify protocol extends across the pair with named barriers (bar.sync id, 64):
// all stores of a batch precede every verify load, and the redo flag, staging counts and filed
// slots are exchanged through a few shared words.  Control flow is identical in both warps.
// ------------------------------------------------------------------------------------------
#ifdef FPB_EMU
__device__ __forceinline__ void team_bar(int id) { fpb_emu::named_barrier(id, 64, 0); }
__device__ __forceinline__ bool team_bar_or(int id, bool pred) { return fpb_emu::named_barrier(id, 64, pred) != 0; }
#else
__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// the pair's barrier with an OR of one predicate per thread folded in (bar.red)
__device__ __forceinline__ bool team_bar_or(int id, bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %2, 0;\n\tbar.red.or.pred p, %1, 64, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(id), "r"((int)pred)
      : "memory");
  return r != 0;
}
#endif

struct TeamX {      // exchange words of one team (shared memory)
  int cnt[2];       // staged cubes per half
  int lp[2];        // slots filed by each half during a sub-round
  int redo[2][2];   // [round parity][half]
};

// cinfo_a: shared address of one 32-bit word per cube = face kinds (low 12 bits) | weight-slot
// base << 16.  Each warp of the pair relaxes half of the staged cubes (all six directions).
template <int NCH2, int NV, bool WG>
__device__ __forceinline__ void team_search(const CtaMem& cm, uint32_t cinfo_a, const WarpMem& wm, const DirTab& dt,
                                            int S, double step, uint32_t present, int lane, int half, int bar_id,
                                            volatile TeamX* tx) {
  const double inv = 1.0 / step;
  const int n16 = (S + 15) >> 4;
  const uint32_t dist_a = smem_addr(wm.dist), bkt_a = smem_addr(wm.bkt);
  const uint32_t stage_a = smem_addr(wm.stage) + (uint32_t)half * STAGE_CAP;  // own half of the staging list
  const uint32_t stage0_a = smem_addr(wm.stage), stage1_a = stage0_a + STAGE_CAP;
  const uint32_t wslot_a = WG ? 0u : smem_addr(cm.wslot);  // WG: weights and face tables in global memory
  constexpr int HCAP = STAGE_CAP / 2;  // entries per half
  int cur = 0;
  while (present) {
    {
      const uint32_t rot = __funnelshift_r(present, present, cur & (NRING - 1));
      cur += __ffs(rot) - 1;
    }
    const int r = cur & (NRING - 1);
    const int far = cur + NRING - 1;
    while ((present >> r) & 1u) {
      present &= ~(1u << r);
      // ---------------- scan (this warp's chunks): stage cubes filed under slot r ----------------
      uint32_t masks[NCH2];
      int cnt = 0;
      const uint32_t r4 = (uint32_t)r * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NCH2; ++j) {
        const int ch = (2 * j + half) * 32 + lane;
        uint32_t m = 0;
        if (ch < n16) {
          uint4 q = lds_v4(bkt_a + (ch << 4));
          const uint32_t z0 = eq_bytes(q.x, r4), z1 = eq_bytes(q.y, r4), z2 = eq_bytes(q.z, r4), z3 = eq_bytes(q.w, r4);
          m = movemask4(z0) | (movemask4(z1) << 4) | (movemask4(z2) << 8) | (movemask4(z3) << 12);
          if (m) {
            q.x |= (z0 >> 7) * 0xffu;
            q.y |= (z1 >> 7) * 0xffu;
            q.z |= (z2 >> 7) * 0xffu;
            q.w |= (z3 >> 7) * 0xffu;
            sts_v4(bkt_a + (ch << 4), q);
          }
        }
        masks[j] = m;
        cnt += __popc(m);
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = incl - cnt;
      uint32_t lp = 0;  // slots this lane filed cubes under
#pragma unroll
      for (int j = 0; j < NCH2; ++j) {
        uint32_t m = masks[j];
        const int base = ((2 * j + half) * 32 + lane) << 4;
        while (m) {
          const int cube = base + __ffs(m) - 1;
          m &= m - 1;
          if (pos < HCAP) sts_u16(stage_a + 2 * pos, cube);
          else sts_u8(bkt_a + cube, r);  // staging half full: leave it filed for the next scan
          ++pos;
        }
      }
      if (total > HCAP) lp |= 1u << r;
      if (lane == 0) tx->cnt[half] = total < HCAP ? total : HCAP;
      team_bar(bar_id);  // (a) both halves of the staging list are complete
      const int n0 = tx->cnt[0], n1 = tx->cnt[1];
      const int nst = n0 + n1;
      // ---------------- relax: 64 staged cubes per round, 32 per warp, all six directions ----------------
      for (int cb = 0; cb < nst; cb += 64) {
        uint32_t ua[6], wa[6];
        double nd[6], old[6];
        bool imp[6];
        const int idx = cb + half * 32 + lane;
        const bool valid = idx < nst;
        const uint32_t sa = idx < n0 ? stage0_a + 2 * idx : stage1_a + 2 * (idx - n0);
        const int v = valid ? (int)lds_u16(sa) : 0;
        const uint32_t va = dist_a + 8u * v;
        const double dv = lds_f64(va);
        const int bt = bucket_of(dv, inv);
        if (valid && bt > cur) {  // a parked cube: move it towards its real bucket
          const int s2 = (bt < far ? bt : far) & (NRING - 1);
          sts_u8(bkt_a + v, s2);
          lp |= 1u << s2;
        }
        const bool act = valid && bt <= cur;
        uint32_t ci = 0xaaau;  // inactive lanes: every face "none"
        if (act) ci = WG ? ((uint32_t)cm.ninfo[v] | ((uint32_t)cm.sbase[v] << 16)) : lds_u32(cinfo_a + 4u * v);
        const uint32_t sb = WG ? (ci >> 16) : wslot_a + 8u * (ci >> 16);  // slot index (WG) or shared address
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          const uint32_t kind = (ci >> (2 * d)) & 3u;
          const bool a = kind < 2u;
          ua[d] = va + 8 * (kind ? dt.dwrap[d] : dt.dreg[d]);
          wa[d] = sb + (WG ? 1 : 8) * (kind ? dt.swrap[d] : dt.sreg[d]);
          // the shared-memory pipe is the limiter: only existing faces load, and the weight is
          // fetched only if the neighbour is farther from the source than this cube
          old[d] = a ? lds_f64(ua[d]) : 0.0;
          imp[d] = a && old[d] > dv;
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          nd[d] = imp[d] ? dv + (WG ? cm.wslot[wa[d]] : lds_f64(wa[d])) : 0.0;
          imp[d] = imp[d] && nd[d] < old[d];
        }
#pragma unroll
        for (int d = 0; d < 6; ++d)
          if (imp[d]) sts_f64(ua[d], nd[d]);
        team_bar(bar_id);  // (b) all stores of the round, from both warps, precede the verify loads
        // verify: a store may have been overwritten by a later, larger one; every improving
        // candidate re-reads its target until no store was lost, so that the bucket each cube is
        // filed under is the bucket of its final distance (filing late costs cascades of rework)
        double fin[6];
        bool chk[6];  // still a possible owner of the target's final value
#pragma unroll
        for (int d = 0; d < 6; ++d) chk[d] = imp[d];
        bool redo;
        do {
          bool mine = false;
#pragma unroll
          for (int d = 0; d < 6; ++d) fin[d] = chk[d] ? lds_f64(ua[d]) : 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const bool lost = chk[d] && fin[d] > nd[d];
            if (lost) sts_f64(ua[d], nd[d]);
            mine |= lost;
            chk[d] = chk[d] && !(fin[d] < nd[d]);  // beaten by a smaller value: its owner files the cube
          }
          redo = team_bar_or(bar_id, mine);  // (c) re-stores visible; did either warp lose a store?
        } while (redo);
        // after the last round every surviving candidate saw its own value in place: it owns the
        // target's final distance and files the cube under that distance's bucket
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          const bool own = chk[d] && fin[d] == nd[d];
          const int bi = bucket_of(nd[d], inv);
          const int s2 = (bi < far ? bi : far) & (NRING - 1);
          const uint32_t ui = (ua[d] - dist_a) >> 3;
          if (own) sts_u8(bkt_a + ui, s2);
          lp |= own ? (1u << s2) : 0u;
        }
      }
      lp = __reduce_or_sync(0xffffffffu, lp);
      if (lane == 0) tx->lp[half] = (int)lp;
      team_bar(bar_id);  // (d) filed slots exchanged, bucket bytes visible to the next scan
      present |= (uint32_t)tx->lp[0] | (uint32_t)tx->lp[1];
    }
  }
}

// ------------------------------------------------------------------------------------------
// team_search_rm: the two-warp search with (1) the bucket bytes in residue-major order and the
// staged cubes dealt to the lanes of a round at a stride, and (2) a verify round that only
// re-checks what it re-stored.
//
// (1) ncu on the round-1 kernel: 47 % of its shared-memory wavefronts were bank-conflict replays
// of the random 8-byte gathers (3.7 wavefronts per 21-lane load where 1.5 would do).  A 64-bit
// access is conflict-free when its lanes' addresses fall into different bank pairs, i.e. when the
// cubes of a half-warp differ mod 16 - and every neighbour gather (a uniform index offset per
// direction) and every weight gather (slot = 3 * cube + const, 3 coprime to 16) then is, too.  So
// byte row r of the bucket array holds the cubes with index = r (mod 16): the scan emits each warp's
// cubes sorted by residue, and warp-round rho (of 2R per pass) deals entry k * 2R + rho - a stride
// sample through both warps' lists with about two cubes of every residue - to lane (k & 1) * 16 + k / 2.
// (2) After the stores of a round (barrier b) every improving candidate re-reads its target once.
// It either sees its own value (it owns the target for now and files it), a smaller one (beaten:
// done), or a larger one (its store was overwritten: it re-stores and must look again next pass).
// Memory only moves down across passes, so a candidate that did not re-store has nothing left to
// lose: later passes load for the re-storers only (round 1 spent 27 % of its wavefronts re-loading
// every candidate 2.25 times per round).  A later, smaller owner files after an earlier one (the
// passes are barrier-separated), so the bucket byte ends up being that of the final distance.
// ------------------------------------------------------------------------------------------
template <int NCHT, bool WG>
__device__ __forceinline__ void team_search_rm(const CtaMem& cm, uint32_t cinfo_a, const WarpMem& wm, const DirTab& dt,
                                               int S, double step, uint32_t present, int lane, int half, int bar_id,
                                               volatile TeamX* tx) {
  const double inv = 1.0 / step;
  const int SB = wm.sb;                 // bytes per residue row = 16-byte chunks in the whole array
  const uint32_t dist_a = smem_addr(wm.dist), bkt_a = smem_addr(wm.bkt);
  const uint32_t stage_a = smem_addr(wm.stage) + (uint32_t)half * STAGE_CAP;  // own half of the staging list
  const uint32_t stage0_a = smem_addr(wm.stage), stage1_a = stage0_a + STAGE_CAP;
  const uint32_t wslot_a = WG ? 0u : smem_addr(cm.wslot);
  constexpr int HCAP = STAGE_CAP / 2;
  // this lane's chunks are consecutive (list order = byte order = residue order); NCHT is odd, so
  // the lanes' 16-byte loads fall into distinct bank quads
  const int ch0 = (half * 32 + lane) * NCHT;
  int cube0[NCHT];  // cube of byte 0 of each chunk: (column << 4) | row
  {
    const uint32_t cr = (uint32_t)SB >> 4;          // chunks per row
    const uint32_t rcp = 0xffffffffu / cr + 1u;     // ch / cr == umulhi(ch, rcp) for ch < 65536
#pragma unroll
    for (int j = 0; j < NCHT; ++j) {
      const uint32_t ch = (uint32_t)(ch0 + j);
      const uint32_t row = __umulhi(ch, rcp);
      const uint32_t col = (ch - row * cr) << 4;
      cube0[j] = (int)((col << 4) | row);
    }
  }
  int cur = 0;
  while (present) {
    {
      const uint32_t rot = __funnelshift_r(present, present, cur & (NRING - 1));
      cur += __ffs(rot) - 1;
    }
    const int r = cur & (NRING - 1);
    const int far = cur + NRING - 1;
    while ((present >> r) & 1u) {
      present &= ~(1u << r);
      // ---------------- scan (this warp's chunks): stage cubes filed under slot r ----------------
      uint32_t masks[NCHT];
      int cnt = 0;
      const uint32_t r4 = (uint32_t)r * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NCHT; ++j) {
        const int ch = ch0 + j;
        uint32_t m = 0;
        if (ch < SB) {
          uint4 q = lds_v4(bkt_a + (ch << 4));
          const uint32_t z0 = eq_bytes(q.x, r4), z1 = eq_bytes(q.y, r4), z2 = eq_bytes(q.z, r4), z3 = eq_bytes(q.w, r4);
          m = movemask4(z0) | (movemask4(z1) << 4) | (movemask4(z2) << 8) | (movemask4(z3) << 12);
          if (m) {
            q.x |= (z0 >> 7) * 0xffu;
            q.y |= (z1 >> 7) * 0xffu;
            q.z |= (z2 >> 7) * 0xffu;
            q.w |= (z3 >> 7) * 0xffu;
            sts_v4(bkt_a + (ch << 4), q);
          }
        }
        masks[j] = m;
        cnt += __popc(m);
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = incl - cnt;
      uint32_t lp = 0;  // slots this lane filed cubes under
#pragma unroll
      for (int j = 0; j < NCHT; ++j) {
        uint32_t m = masks[j];
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          if (pos < HCAP) sts_u16(stage_a + 2 * pos, cube0[j] + (bit << 4));
          else sts_u8(bkt_a + ((ch0 + j) << 4) + bit, r);  // staging half full: leave it filed for the next scan
          ++pos;
        }
      }
      if (total > HCAP) lp |= 1u << r;
      if (lane == 0) tx->cnt[half] = total < HCAP ? total : HCAP;
      team_bar(bar_id);  // (a) both halves of the staging list are complete
      const int n0 = tx->cnt[0], n1 = tx->cnt[1];
      const int nst = n0 + n1;
      // ---------------- relax: W warp-rounds, entry k * W + rho of the lists goes to round rho ----------------
      // a 64-bit access is served one half-warp at a time: the 16 lanes of a half-warp take every second
      // entry of the stride sample, so that each half-warp holds (about) one cube of every residue
      const int W = (nst + 31) >> 5;
      const int deal = (((lane & 15) << 1) | (lane >> 4)) * W;
      for (int rho = half; rho < ((W + 1) & ~1); rho += 2) {
        // W odd: the last round is warp 0's alone; warp 1 only takes part in its barriers
        const bool live = rho < W;
        uint32_t ua[6], fa[6], fs[6];
        double nd[6];
        bool chk[6];  // stored in the last pass: must see what became of it
#pragma unroll
        for (int d = 0; d < 6; ++d) chk[d] = false;
        if (live) {
          uint32_t wa[6];
          double old[6];
          bool imp[6];
          const int idx = deal + rho;
          const bool valid = idx < nst;
          const uint32_t sa = idx < n0 ? stage0_a + 2 * idx : stage1_a + 2 * (idx - n0);
          const int v = valid ? (int)lds_u16(sa) : 0;
          const uint32_t va = dist_a + 8u * v;
          const double dv = lds_f64(va);
          const int bt = bucket_of(dv, inv);
          if (valid && bt > cur) {  // a parked cube: move it towards its real bucket
            const int s2 = (bt < far ? bt : far) & (NRING - 1);
            sts_u8(bkt_a + (v & 15) * SB + (v >> 4), s2);
            lp |= 1u << s2;
          }
          const bool act = valid && bt <= cur;
          uint32_t ci = 0xaaau;  // inactive lanes: every face "none"
          if (act) ci = WG ? ((uint32_t)cm.ninfo[v] | ((uint32_t)cm.sbase[v] << 16)) : lds_u32(cinfo_a + 4u * v);
          const uint32_t sb = WG ? (ci >> 16) : wslot_a + 8u * (ci >> 16);  // slot index (WG) or shared address
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const uint32_t kind = (ci >> (2 * d)) & 3u;
            const bool a = kind < 2u;
            ua[d] = va + 8 * (kind ? dt.dwrap[d] : dt.dreg[d]);
            wa[d] = sb + (WG ? 1 : 8) * (kind ? dt.swrap[d] : dt.sreg[d]);
            old[d] = a ? lds_f64(ua[d]) : 0.0;
            imp[d] = a && old[d] > dv;
          }
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            nd[d] = imp[d] ? dv + (WG ? cm.wslot[wa[d]] : lds_f64(wa[d])) : 0.0;
            imp[d] = imp[d] && nd[d] < old[d];
          }
#pragma unroll
          for (int d = 0; d < 6; ++d)
            if (imp[d]) sts_f64(ua[d], nd[d]);
          // where an improved target is filed: byte address and ring slot of its new distance
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const int bi = bucket_of(nd[d], inv);
            fs[d] = (uint32_t)((bi < far ? bi : far) & (NRING - 1));
            const uint32_t ui = (ua[d] - dist_a) >> 3;
            fa[d] = bkt_a + (ui & 15u) * (uint32_t)SB + (ui >> 4);
            chk[d] = imp[d];
          }
        }
        team_bar(bar_id);  // (b) all stores of the round, from both warps, precede the verify loads
        bool redo;
        do {
          bool mine = false;
          const bool any = chk[0] | chk[1] | chk[2] | chk[3] | chk[4] | chk[5];
          if (__any_sync(0xffffffffu, any)) {
            double fin[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) fin[d] = chk[d] ? lds_f64(ua[d]) : 0.0;
#pragma unroll
            for (int d = 0; d < 6; ++d) {
              const bool lost = chk[d] && fin[d] > nd[d];
              const bool own = chk[d] && fin[d] == nd[d];  // in place: it files the target (a later, smaller owner re-files)
              if (lost) sts_f64(ua[d], nd[d]);
              if (own) sts_u8(fa[d], fs[d]);
              lp |= own ? (1u << fs[d]) : 0u;
              mine |= lost;
              chk[d] = lost;
            }
          }
          redo = team_bar_or(bar_id, mine);  // (c) re-stores and bytes visible; did either warp re-store?
        } while (redo);
      }
      lp = __reduce_or_sync(0xffffffffu, lp);
      if (lane == 0) tx->lp[half] = (int)lp;
      team_bar(bar_id);  // (d) filed slots exchanged, bucket bytes visible to the next scan
      present |= (uint32_t)tx->lp[0] | (uint32_t)tx->lp[1];
    }
  }
}

// RM: residue-major bucket bytes + stride-dealt rounds + lean verify (team_search_rm; NCH2 is then the
// odd number of consecutive 16-byte chunks per lane, 64 * NCH2 >= paths_bkt_row(S_max))
template <int NCH2, int NV, int MAXT, int MINB, bool WG, bool RM = false>
__global__ void __launch_bounds__(MAXT, MINB) k_paths_team(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int item_s;
  __shared__ double red_s[32];
  __shared__ TeamX tx_s[16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int team = wid >> 1, half = wid & 1, n_teams = blockDim.x >> 6;
  const int bar_id = 1 + team;
  const int nthreads = blockDim.x;
  volatile TeamX* tx = &tx_s[team];

  unsigned char* ptr = smem;
  double* wslot;
  uint32_t* cinfo_s = nullptr;
  if (WG) {
    wslot = p.wscratch + (size_t)blockIdx.x * p.slots_max;
  } else {
    wslot = (double*)ptr; ptr += align16((size_t)p.slots_max * 8);
    cinfo_s = (uint32_t*)ptr; ptr += 2 * align16((size_t)p.S_max * 2);  // ninfo | sbase << 16
  }
  CtaMem cm;
  cm.wslot = wslot; cm.ninfo = nullptr; cm.sbase = nullptr;
  const uint32_t cinfo_a = WG ? 0u : smem_addr(cinfo_s);
  WarpMem wm = carve_warp(ptr + (size_t)team * paths_warp_bytes(p.S_max, false), p.S_max, false);
  DirTab dt = load_dirtab(p.cx[0]);

  int staged_tc = -1, staged_c = -1;
  double step = 1.0;

  while (true) {
    if (threadIdx.x == 0) item_s = atomicAdd(p.counter, 1);
    __syncthreads();
    const int item = item_s;
    __syncthreads();
    if (item >= p.total_items) break;
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
    wm.sb = RM ? paths_bkt_row(S) : 0;

    if (staged_c != ci) {
      if (WG) {
        cm.ninfo = c.ninfo;
        cm.sbase = c.sbase;
      } else {
        for (int v = threadIdx.x; v < S; v += nthreads) cinfo_s[v] = (uint32_t)c.ninfo[v] | ((uint32_t)c.sbase[v] << 16);
      }
      dt = load_dirtab(c);
      staged_c = ci;
    }
    if (staged_tc != tc) {
      const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
      double sum = 0.0;
      for (int s = threadIdx.x; s < c.n_slots; s += nthreads) {
        const int q = c.slot_qubit[s];
        const double ws = (q >= 0) ? w[q] : FPB_INF;
        wslot[s] = ws;
        sum += (q >= 0) ? ws : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) red_s[wid] = sum;
      __syncthreads();
      sum = 0.0;
      for (int w2 = 0; w2 < (nthreads >> 5); ++w2) sum += red_s[w2];
      step = bucket_step(sum / (double)c.Q, p.step_factor);
      staged_tc = tc;
      if (WG) __threadfence_block();
    }
    __syncthreads();

    const double inv = 1.0 / step;
    const int K = c.odd_count[t];
    const int* odd = c.odd_fixed + (long long)t * S;
    const int has_bnd = c.n_low > 0;
    const int n_tasks = n_tasks_of(K, has_bnd);
    const long long poff = c.pair_off[t];
    const long long ooff = c.odd_off[t];
    const int task_end = min(n_tasks, (chunk + 1) * p.tasks_per_item);
    for (int task = chunk * p.tasks_per_item + team; task < task_end; task += n_teams) {
      const bool pair_task = task < K - 1;
      const int n_phase = pair_task ? 1 : 2;
      for (int ph = 0; ph < n_phase; ++ph) {
        // clear the shared state, half each
        for (int v = half * 32 + lane; v < S; v += 64) wm.dist[v] = FPB_INF;
        {
          const int n16 = RM ? wm.sb : (S + 15) >> 4;
          uint4* b4 = reinterpret_cast<uint4*>(wm.bkt);
          for (int i = half * 32 + lane; i < n16; i += 64) b4[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
        team_bar(bar_id);
        uint32_t present = 0;
        if (half == 0) {
          if (pair_task) {
            if (lane == 0) present = seed_cube(wm, odd[task], 0.0, inv, 0);
            present = __shfl_sync(0xffffffffu, present, 0);
          } else if (ph == 0) {
            present = warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, inv, lane);
          } else {
            present = warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, inv, lane);
          }
          if (lane == 0) tx->lp[0] = (int)present;
        }
        team_bar(bar_id);
        present = (uint32_t)tx->lp[0];
        team_bar(bar_id);  // lp[0] is rewritten inside the search
        if (RM) team_search_rm<NCH2, WG>(cm, cinfo_a, wm, dt, S, step, present, lane, half, bar_id, tx);
        else team_search<NCH2, NV, WG>(cm, cinfo_a, wm, dt, S, step, present, lane, half, bar_id, tx);
        if (pair_task) {
          const long long row = poff + (long long)task * K - (long long)task * (task + 1) / 2;
          for (int b = task + 1 + half * 32 + lane; b < K; b += 64) c.pair_len[row + (b - task - 1)] = wm.dist[odd[b]];
        } else if (ph == 0) {
          for (int b = half * 32 + lane; b < K; b += 64) c.bnd_len[ooff + b] = wm.dist[odd[b]];
        } else {
          for (int b = half * 32 + lane; b < K; b += 64) {
            const double lo_d = c.bnd_len[ooff + b];
            const double hi_d = wm.dist[odd[b]];
            const bool high = hi_d < lo_d;
            c.bnd_len[ooff + b] = high ? hi_d : lo_d;
            c.bnd_side[ooff + b] = high ? 1 : 0;
          }
        }
        team_bar(bar_id);  // outputs read before the state is cleared again
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// W-warp teams (W = 4, 8, 16) for lattices so large that only one to four searches fit an SM's
// shared memory: the same protocol as the two-warp team, with the scan chunks, the staging list
// and every relaxation round split W ways (32 * W cubes per round), so that the few searches in
// flight still occupy the SM.  Weights and face tables stay in global memory (WG layout).
// ------------------------------------------------------------------------------------------
template <int W>
struct TeamXW {
  int cnt[W];
  int lp[W];
  int redo[2][W];
};
#ifdef FPB_EMU
template <int W>
__device__ __forceinline__ void teamw_bar(int id) { fpb_emu::named_barrier(id, 32 * W, 0); }
template <int W>
__device__ __forceinline__ bool teamw_bar_or(int id, bool pred) { return fpb_emu::named_barrier(id, 32 * W, pred) != 0; }
#else
template <int W>
__device__ __forceinline__ void teamw_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * W) : "memory"); }
// barrier over the team that also ORs one predicate of every thread (one instruction instead of
// flag store + barrier + flag loads)
template <int W>
__device__ __forceinline__ bool teamw_bar_or(int id, bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %2, 0;\n\tbar.red.or.pred p, %1, %3, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(id), "r"((int)pred), "n"(32 * W)
      : "memory");
  return r != 0;
}
#endif

template <int NCHW, int W>
__device__ __forceinline__ void teamw_search(const CtaMem& cm, const WarpMem& wm, const DirTab& dt, int S, double step,
                                             uint32_t present, int lane, int widx, int bar_id, int seg,
                                             volatile TeamXW<W>* tx) {
  const double inv = 1.0 / step;
  const int n16 = (S + 15) >> 4;
  const uint32_t dist_a = smem_addr(wm.dist), bkt_a = smem_addr(wm.bkt);
  const uint32_t stage0_a = smem_addr(wm.stage);
  const uint32_t stage_a = stage0_a + 2u * (uint32_t)(widx * seg);  // this warp's segment of the staging list
  int cur = 0;
  while (present) {
    {
      const uint32_t rot = __funnelshift_r(present, present, cur & (NRING - 1));
      cur += __ffs(rot) - 1;
    }
    const int r = cur & (NRING - 1);
    const int far = cur + NRING - 1;
    while ((present >> r) & 1u) {
      present &= ~(1u << r);
      uint32_t masks[NCHW];
      int cnt = 0;
      const uint32_t r4 = (uint32_t)r * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NCHW; ++j) {
        const int ch = (W * j + widx) * 32 + lane;
        uint32_t m = 0;
        if (ch < n16) {
          uint4 q = lds_v4(bkt_a + (ch << 4));
          const uint32_t z0 = eq_bytes(q.x, r4), z1 = eq_bytes(q.y, r4), z2 = eq_bytes(q.z, r4), z3 = eq_bytes(q.w, r4);
          m = movemask4(z0) | (movemask4(z1) << 4) | (movemask4(z2) << 8) | (movemask4(z3) << 12);
          if (m) {
            q.x |= (z0 >> 7) * 0xffu;
            q.y |= (z1 >> 7) * 0xffu;
            q.z |= (z2 >> 7) * 0xffu;
            q.w |= (z3 >> 7) * 0xffu;
            sts_v4(bkt_a + (ch << 4), q);
          }
        }
        masks[j] = m;
        cnt += __popc(m);
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = incl - cnt;
      uint32_t lp = 0;
#pragma unroll
      for (int j = 0; j < NCHW; ++j) {
        uint32_t m = masks[j];
        const int base = ((W * j + widx) * 32 + lane) << 4;
        while (m) {
          const int cube = base + __ffs(m) - 1;
          m &= m - 1;
          if (pos < seg) sts_u16(stage_a + 2 * pos, cube);
          else sts_u8(bkt_a + cube, r);  // segment full: leave it filed for the next scan
          ++pos;
        }
      }
      if (total > seg) lp |= 1u << r;
      if (lane == 0) tx->cnt[widx] = total < seg ? total : seg;
      teamw_bar<W>(bar_id);  // (a) every segment of the staging list is complete
      // lane w < W holds segment w's count and the number of cubes staged before it
      const int my_cnt = lane < W ? tx->cnt[lane] : 0;
      int my_end = my_cnt;
#pragma unroll
      for (int o = 1; o < W; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, my_end, o);
        if (lane >= o) my_end += y;
      }
      const int nst = __shfl_sync(0xffffffffu, my_end, W - 1);
      for (int cb = 0; cb < nst; cb += 32 * W) {
        uint32_t ua[6], wa[6];
        double nd[6], old[6];
        bool imp[6];
        const int idx = cb + widx * 32 + lane;
        const bool valid = idx < nst;
        // segment of idx: the first w with idx < end[w]
        int sgm = 0, sbeg = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int e = __shfl_sync(0xffffffffu, my_end, w);
          if (idx >= e) {
            sgm = w + 1;
            sbeg = e;
          }
        }
        const uint32_t sa = stage0_a + 2u * (uint32_t)((valid ? sgm : 0) * seg + (valid ? idx - sbeg : 0));
        const int v = valid ? (int)lds_u16(sa) : 0;
        const uint32_t va = dist_a + 8u * v;
        const double dv = lds_f64(va);
        const int bt = bucket_of(dv, inv);
        if (valid && bt > cur) {  // a parked cube: move it towards its real bucket
          const int s2 = (bt < far ? bt : far) & (NRING - 1);
          sts_u8(bkt_a + v, s2);
          lp |= 1u << s2;
        }
        const bool act = valid && bt <= cur;
        uint32_t ci = 0xaaau;  // inactive lanes: every face "none"
        if (act) ci = (uint32_t)cm.ninfo[v] | ((uint32_t)cm.sbase[v] << 16);
        const uint32_t sb = ci >> 16;
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          const uint32_t kind = (ci >> (2 * d)) & 3u;
          const bool a = kind < 2u;
          ua[d] = va + 8 * (kind ? dt.dwrap[d] : dt.dreg[d]);
          wa[d] = sb + (kind ? dt.swrap[d] : dt.sreg[d]);
          old[d] = a ? lds_f64(ua[d]) : 0.0;
          imp[d] = a && old[d] > dv;
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          nd[d] = imp[d] ? dv + cm.wslot[wa[d]] : 0.0;
          imp[d] = imp[d] && nd[d] < old[d];
        }
#pragma unroll
        for (int d = 0; d < 6; ++d)
          if (imp[d]) sts_f64(ua[d], nd[d]);
        teamw_bar<W>(bar_id);  // (b) all stores of the round precede the verify loads
        double fin[6];
        bool chk[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) chk[d] = imp[d];
        bool redo;
        do {
          bool mine = false;
#pragma unroll
          for (int d = 0; d < 6; ++d) fin[d] = chk[d] ? lds_f64(ua[d]) : 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const bool lost = chk[d] && fin[d] > nd[d];
            if (lost) sts_f64(ua[d], nd[d]);
            mine |= lost;
            chk[d] = chk[d] && !(fin[d] < nd[d]);
          }
          redo = teamw_bar_or<W>(bar_id, mine);  // (c) re-stores visible; any store lost anywhere in the team?
        } while (redo);
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          const bool own = chk[d] && fin[d] == nd[d];
          const int bi = bucket_of(nd[d], inv);
          const int s2 = (bi < far ? bi : far) & (NRING - 1);
          const uint32_t ui = (ua[d] - dist_a) >> 3;
          if (own) sts_u8(bkt_a + ui, s2);
          lp |= own ? (1u << s2) : 0u;
        }
      }
      lp = __reduce_or_sync(0xffffffffu, lp);
      if (lane == 0) tx->lp[widx] = (int)lp;
      teamw_bar<W>(bar_id);  // (d) filed slots exchanged, bucket bytes visible to the next scan
      present |= __reduce_or_sync(0xffffffffu, lane < W ? (uint32_t)tx->lp[lane] : 0u);
    }
  }
}

template <int NCHW, int W, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_paths_teamw(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int item_s;
  __shared__ double red_s[32];
  __shared__ TeamXW<W> tx_s[16 / W];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int team = wid / W, widx = wid % W, n_teams = blockDim.x / (32 * W);
  const int bar_id = 1 + team;
  const int nthreads = blockDim.x;
  volatile TeamXW<W>* tx = &tx_s[team];
  const int seg = paths_stage_entries(p.S_max) / W;

  double* wslot = p.wscratch + (size_t)blockIdx.x * p.slots_max;
  CtaMem cm;
  cm.wslot = wslot; cm.ninfo = nullptr; cm.sbase = nullptr;
  const WarpMem wm = carve_warp(smem + (size_t)team * paths_warp_bytes(p.S_max, false), p.S_max, false);
  DirTab dt = load_dirtab(p.cx[0]);

  int staged_tc = -1, staged_c = -1;
  double step = 1.0;

  while (true) {
    if (threadIdx.x == 0) item_s = atomicAdd(p.counter, 1);
    __syncthreads();
    const int item = item_s;
    __syncthreads();
    if (item >= p.total_items) break;
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
      cm.ninfo = c.ninfo;
      cm.sbase = c.sbase;
      dt = load_dirtab(c);
      staged_c = ci;
    }
    if (staged_tc != tc) {
      const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
      double sum = 0.0;
      for (int s = threadIdx.x; s < c.n_slots; s += nthreads) {
        const int q = c.slot_qubit[s];
        const double ws = (q >= 0) ? w[q] : FPB_INF;
        wslot[s] = ws;
        sum += (q >= 0) ? ws : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) red_s[wid] = sum;
      __syncthreads();
      sum = 0.0;
      for (int w2 = 0; w2 < (nthreads >> 5); ++w2) sum += red_s[w2];
      step = bucket_step(sum / (double)c.Q, p.step_factor);
      staged_tc = tc;
      __threadfence_block();
    }
    __syncthreads();

    const double inv = 1.0 / step;
    const int K = c.odd_count[t];
    const int* odd = c.odd_fixed + (long long)t * S;
    const int has_bnd = c.n_low > 0;
    const int n_tasks = n_tasks_of(K, has_bnd);
    const long long poff = c.pair_off[t];
    const long long ooff = c.odd_off[t];
    const int task_end = min(n_tasks, (chunk + 1) * p.tasks_per_item);
    for (int task = chunk * p.tasks_per_item + team; task < task_end; task += n_teams) {
      const bool pair_task = task < K - 1;
      const int n_phase = pair_task ? 1 : 2;
      for (int ph = 0; ph < n_phase; ++ph) {
        for (int v = widx * 32 + lane; v < S; v += 32 * W) wm.dist[v] = FPB_INF;
        {
          const int n16 = (S + 15) >> 4;
          uint4* b4 = reinterpret_cast<uint4*>(wm.bkt);
          for (int i = widx * 32 + lane; i < n16; i += 32 * W) b4[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
        teamw_bar<W>(bar_id);
        uint32_t present = 0;
        if (widx == 0) {
          if (pair_task) {
            if (lane == 0) present = seed_cube(wm, odd[task], 0.0, inv, 0);
            present = __shfl_sync(0xffffffffu, present, 0);
          } else if (ph == 0) {
            present = warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, inv, lane);
          } else {
            present = warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, inv, lane);
          }
          if (lane == 0) tx->lp[0] = (int)present;
        }
        teamw_bar<W>(bar_id);
        present = (uint32_t)tx->lp[0];
        teamw_bar<W>(bar_id);  // lp[0] is rewritten inside the search
        teamw_search<NCHW, W>(cm, wm, dt, S, step, present, lane, widx, bar_id, seg, tx);
        if (pair_task) {
          const long long row = poff + (long long)task * K - (long long)task * (task + 1) / 2;
          for (int b = task + 1 + widx * 32 + lane; b < K; b += 32 * W) c.pair_len[row + (b - task - 1)] = wm.dist[odd[b]];
        } else if (ph == 0) {
          for (int b = widx * 32 + lane; b < K; b += 32 * W) c.bnd_len[ooff + b] = wm.dist[odd[b]];
        } else {
          for (int b = widx * 32 + lane; b < K; b += 32 * W) {
            const double lo_d = c.bnd_len[ooff + b];
            const double hi_d = wm.dist[odd[b]];
            const bool high = hi_d < lo_d;
            c.bnd_len[ooff + b] = high ? hi_d : lo_d;
            c.bnd_side[ooff + b] = high ? 1 : 0;
          }
        }
        teamw_bar<W>(bar_id);  // outputs read before the state is cleared again
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// k_path_toggle: one warp per query; search with early stop, then walk the tree back and
// XOR-toggle the crossed qubits (mwpm/algos.py:88-95)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t toggle_warp_bytes(int S_max, int slots_max) {
  return align16((size_t)slots_max * 8) + paths_warp_bytes(S_max, true);
}

template <int NCH, bool WG>
__global__ void __launch_bounds__(256) k_path_toggle(const __grid_constant__ PathsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  // queries of one CTA mix trials, so every warp keeps its own copy of the trial's weights;
  // the static face tables are read from global memory (L1/L2 resident)
  unsigned char* wp;
  double* wslot;
  if (WG) {
    wp = smem + (size_t)wid * paths_warp_bytes(p.S_max, true);
    wslot = p.wscratch + ((size_t)blockIdx.x * nwarps + wid) * p.slots_max;
  } else {
    wp = smem + (size_t)wid * toggle_warp_bytes(p.S_max, p.slots_max);
    wslot = (double*)wp; wp += align16((size_t)p.slots_max * 8);
  }
  const WarpMem wm = carve_warp(wp, p.S_max, true);
  for (int qi = blockIdx.x * nwarps + wid; qi < p.n_queries; qi += gridDim.x * nwarps) {
    const int t = p.q_trial[qi], ci = p.q_cx[qi];
    const ComplexDev& c = p.cx[ci];
    const int S = c.S;
    const DirTab dt = load_dirtab(c);
    const double* w = p.weights + (long long)t * p.Qtot + c.qoff;
    double sum = 0.0;
    for (int s = lane; s < c.n_slots; s += 32) {
      const int q = c.slot_qubit[s];
      const double ws = (q >= 0) ? w[q] : FPB_INF;
      wslot[s] = ws;
      sum += (q >= 0) ? ws : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (WG) __threadfence_block();
    __syncwarp();
    const double step = bucket_step(sum / (double)c.Q, p.step_factor);
    const double inv = 1.0 / step;
    CtaMem cm;
    cm.wslot = wslot; cm.ninfo = c.ninfo; cm.sbase = c.sbase;
    const int* odd = c.odd_fixed + (long long)t * S;
    const int src = odd[p.q_src[qi]];
    const bool pair_query = p.q_dst[qi] >= 0;
    uint32_t* flips = p.flips + (long long)t * p.bit_words;
    int* plist = p.path_qubits ? p.path_qubits + (long long)qi * p.path_cap : nullptr;
    warp_clear(wm, S, lane);
    int walk_from, stop, side = 0;
    uint32_t present = 0;
    if (pair_query) {
      if (lane == 0) present = seed_cube(wm, src, 0.0, inv, 6);
      present = __shfl_sync(0xffffffffu, present, 0);
      __syncwarp();
      stop = walk_from = odd[p.q_dst[qi]];
    } else {
      // to the nearer connector, as recorded by the matching-graph run
      side = c.bnd_side[c.odd_off[t] + p.q_src[qi]];
      present = side ? warp_seed_connector(cm, wm, c.n_high, c.high_cube, c.high_slot, inv, lane)
                     : warp_seed_connector(cm, wm, c.n_low, c.low_cube, c.low_slot, inv, lane);
      stop = walk_from = src;
    }
    warp_search<true, NCH, 1, WG>(cm, wm, dt, S, step, present, lane, stop);
    __syncwarp();
    if (lane == 0) {
      int u = walk_from;
      int steps = 0;
      // one crossed qubit: XOR-toggle it (fpb_paths_toggle) or append it to the query's list (fpb_paths_list)
      auto cross = [&](int local_q) {
        if (plist) {
          if (steps < p.path_cap) plist[steps] = local_q;
          ++steps;
        } else {
          const int g = c.qoff + local_q;
          atomicXor(&flips[g >> 5], 1u << (g & 31));
        }
      };
      for (int guard = 0; guard < S + 2; ++guard) {
        const int din = wm.pred[u];
        if (din == 6) break;  // reached the source cube
        if (din == 7) {
          // seeded from a connector: toggle the boundary qubit and stop
          const int n = side ? c.n_high : c.n_low;
          const int* cubes = side ? c.high_cube : c.low_cube;
          const int* slots = side ? c.high_slot : c.low_slot;
          for (int i = 0; i < n; ++i)
            if (cubes[i] == u) {
              cross(c.slot_qubit[slots[i]]);
              break;
            }
          break;
        }
        // u was reached from v by direction din; step back along din^1
        const int back = din ^ 1;
        const uint32_t kb = (c.ninfo[u] >> (2 * back)) & 3u;
        const int v = u + (kb ? dt.dwrap[back] : dt.dreg[back]);
        const uint32_t kf = (c.ninfo[v] >> (2 * din)) & 3u;
        const int slot = c.sbase[v] + (kf ? dt.swrap[din] : dt.sreg[din]);
        cross(c.slot_qubit[slot]);
        u = v;
      }
      if (plist) p.path_len[qi] = steps <= p.path_cap ? steps : -1;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Static-weight path (uniform weights, or blueprint weights with every node p-squeezed:
// decoder.py:86-93,115-116 make the weights trial-independent).  The all-pairs distances of the
// cube grid are computed once with k_paths (odd list = every cube) into a packed upper triangle;
// a trial's matching graph is then a pure gather - the HBM-bound case of SURVEY.md section 7.
// ------------------------------------------------------------------------------------------
__global__ void k_fill_static(double* weights, const double* static_w, int Qtot, long long T) {
  const long long n = T * Qtot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    weights[i] = static_w[i % Qtot];
}

// flag != 0 if any trial's weights differ from the static ones (labels were not all "p")
__global__ void k_check_static(const double* weights, const double* static_w, int Qtot, long long T, int* flag) {
  const long long n = T * Qtot;
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= weights[i] != static_w[i % Qtot];
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void k_iota_odd(ComplexDev c) {  // the "every cube is odd" pseudo-trial of the table build
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.S; i += gridDim.x * blockDim.x) c.odd_fixed[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) c.odd_count[0] = c.S;
}

constexpr int KNN_NONE = 0x7fffffff;

// (length, index) order of the candidate lists: shorter first, ties by the lower index.
__device__ __forceinline__ bool knn_less(double d, int b, double e, int f) { return d < e || (d == e && b < f); }

// Keeps (d0,b0) <= (d1,b1) <= (d2,b2), the three smallest seen so far.
__device__ __forceinline__ void knn_insert(double d, int b, double& d0, int& b0, double& d1, int& b1, double& d2,
                                           int& b2) {
  if (!knn_less(d, b, d2, b2)) return;
  if (knn_less(d, b, d1, b1)) {
    d2 = d1, b2 = b1;
    if (knn_less(d, b, d0, b0)) {
      d1 = d0, b1 = b0, d0 = d, b0 = b;
    } else {
      d1 = d, b1 = b;
    }
  } else {
    d2 = d, b2 = b;
  }
}

// ------------------------------------------------------------------------------------------
// k_knn: candidate partners for the host matcher.  One warp per odd cube a of trial t: its row of
// the reduced complete graph, min(d_ab, b_a + b_b) (decoders/mwpm/matching.py:144-173 folded into
// the pair lengths, see csrc/mwpm.cpp), is staged in shared memory and the kn smallest entries are
// extracted in kn rounds of a warp-wide lexicographic (length, index) minimum.  The host solver
// starts from this sparse graph and proves optimality against the complete graph by pricing.
// ------------------------------------------------------------------------------------------
__global__ void k_knn(ComplexDev c, int kn, int row_cap, int* out /* [total_odd][kn] */,
                      double* out_len /* [total_odd][kn] reduced lengths, or nullptr */,
                      unsigned long long* trial_wmax /* [T] bits of the largest reduced length / 2 x boundary length, or nullptr */) {
  extern __shared__ double knn_rows[];  // [warps][row_cap]
  const int t = blockIdx.x;
  const int K = c.odd_count[t];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int a = blockIdx.y * nw + wid;
  if (a >= K) return;
  double* row = knn_rows + (size_t)wid * row_cap;
  const long long poff = c.pair_off[t], ooff = c.odd_off[t];
  const double* pl = c.pair_len + poff;
  const bool bnd = c.n_low > 0;
  const double ba = bnd ? c.bnd_len[ooff + a] : 0.0;
  double rmax = 2.0 * ba;
  double d0 = FPB_INF, d1 = FPB_INF, d2 = FPB_INF;
  int b0 = KNN_NONE, b1 = KNN_NONE, b2 = KNN_NONE, mine = 0;
  for (int b = lane; b < K; b += 32) {
    double d = FPB_INF;
    if (b < a) d = pl[(long long)b * K - (long long)b * (b + 1) / 2 + (a - b - 1)];
    else if (b > a) d = pl[(long long)a * K - (long long)a * (a + 1) / 2 + (b - a - 1)];
    if (bnd && b != a) d = fmin(d, ba + c.bnd_len[ooff + b]);
    row[b] = d;
    if (b != a) {
      rmax = fmax(rmax, d);
      ++mine;
      knn_insert(d, b, d0, b0, d1, b1, d2, b2);
    }
  }
  if (trial_wmax) {  // non-negative doubles order like their bit patterns
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, off));
    if (lane == 0) atomicMax(&trial_wmax[t], (unsigned long long)__double_as_longlong(rmax));
  }
  // Selection: every lane keeps the three smallest (length, index) of its own stride in registers;
  // a round is then one warp argmin over the lane heads.  A lane whose three are used up rescans
  // its stride for the next three (rare: twelve winners seldom sit in one residue class mod 32).
  int used = 0;
  int* o = out + (ooff + a) * kn;
  for (int r = 0; r < kn; ++r) {
    double bd = d0;
    int bb = b0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const int ob = __shfl_xor_sync(0xffffffffu, bb, off);
      if (knn_less(od, ob, bd, bb)) {
        bd = od;
        bb = ob;
      }
    }
    if (lane == 0) {
      o[r] = bb == KNN_NONE ? -1 : bb;
      if (out_len) out_len[(ooff + a) * kn + r] = bd;
    }
    if (bb != KNN_NONE && bb == b0) {  // this lane's head was taken
      ++used;
      const double ld = d0;
      const int lb = b0;
      d0 = d1, b0 = b1, d1 = d2, b1 = b2, d2 = FPB_INF, b2 = KNN_NONE;
      if (b0 == KNN_NONE && used < mine) {
        for (int b = lane; b < K; b += 32) {
          const double d = row[b];
          if (b != a && knn_less(ld, lb, d, b)) knn_insert(d, b, d0, b0, d1, b1, d2, b2);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Sparse hand-off to the host matcher: the triangle of lengths stays in HBM.  k_pairs_at returns the
// lengths of requested pairs; k_price tests every pair of the active trials against the matcher's
// vertex duals in the matcher's own integer arithmetic, y_a + y_b + ((4 * fixed(w_ab * scale_t)) <<
// wshift_t) < 0 with w_ab the reduced length min(d_ab, b_a + b_b) and fixed(v) = (long long)(v + 0.5)
// (csrc/mwpm.cpp: to_fixed, wt), and appends the violators - the pricing pass of csrc/mwpm.cpp moved
// to where the lengths are.  Exact, so tight pairs (slack 0: every two cubes that end at their
// connectors) are not reported.  Blossom duals are non-negative and only add slack, so the vertex
// test never misses a violated edge; the host adds them when it re-checks the returned pairs.
// ------------------------------------------------------------------------------------------
__global__ void k_pairs_at(ComplexDev c, long long n, const int* trial, const int* qa, const int* qb, double* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = trial[i];
    const long long K = c.odd_count[t];
    const long long a = min(qa[i], qb[i]), b = max(qa[i], qb[i]);
    out[i] = (a == b) ? 0.0 : c.pair_len[c.pair_off[t] + a * K - a * (a + 1) / 2 + (b - a - 1)];
  }
}

constexpr int PRICE_ROWS = 8;  // rows (= warps) per CTA

struct PriceParams {
  const long long* y;          // [total_odd] vertex duals of the host matcher
  const int* top;              // [total_odd] top-level blossom of each vertex (-1: none), or nullptr
  const long long* ztop;       // [total_odd] 2 x the dual of that blossom
  const double* scale;         // [T] fixed-point scale of the trial's lengths (a power of two)
  const int* wshift;           // [T] doublings applied to the trial's weights and duals
  const uint8_t* active;       // [T] trials to price
  long long cap;
  int* out_trial;
  int* out_a;
  int* out_b;
  double* out_len;             // reduced length of the violating pair
  unsigned long long* n_out;   // appended entries (may exceed cap: the caller retries with more room)
};

__global__ void __launch_bounds__(PRICE_ROWS * 32) k_price(ComplexDev c, PriceParams p) {
  const int t = blockIdx.x;
  if (!p.active[t]) return;
  const int K = c.odd_count[t];
  const int lane = threadIdx.x & 31;
  const long long poff = c.pair_off[t], ooff = c.odd_off[t];
  const bool bnd = c.n_low > 0;
  const double scale = p.scale[t];
  const int wshift = p.wshift[t] + 2;  // 4 * fixed << wshift
  const long long* y = p.y + ooff;
  const int* top = p.top ? p.top + ooff : nullptr;
  const long long* ztop = p.top ? p.ztop + ooff : nullptr;
  for (int a = blockIdx.y * PRICE_ROWS + (threadIdx.x >> 5); a < K - 1; a += gridDim.y * PRICE_ROWS) {
    const double* row = c.pair_len + poff + ((long long)a * K - (long long)a * (a + 1) / 2 - a - 1);
    const long long ya = y[a];
    const int ta = top ? top[a] : -1;
    const long long za = ta >= 0 ? ztop[a] : 0;
    const double ba = bnd ? c.bnd_len[ooff + a] : 0.0;
    for (int b = a + 1 + lane; b < K; b += 32) {
      double w = row[b];
      if (bnd) w = fmin(w, ba + c.bnd_len[ooff + b]);
      // two vertices of one top-level blossom: its dual is part of their slack (a lower bound of
      // the sum over all their common blossoms, so still no violated pair is missed)
      const long long zc = (ta >= 0 && top[b] == ta) ? za : 0;
      if (ya + y[b] + zc + (__double2ll_rz(w * scale + 0.5) << wshift) < 0) {
        const unsigned long long pos = atomicAdd(p.n_out, 1ull);
        if ((long long)pos < p.cap) {
          p.out_trial[pos] = t;
          p.out_a[pos] = a;
          p.out_b[pos] = b;
          p.out_len[pos] = w;
        }
      }
    }
  }
}

constexpr int GATHER_ROWS = 8;  // rows (= warps) per CTA

__global__ void __launch_bounds__(GATHER_ROWS * 32) k_gather(ComplexDev c, const double* tab_pair,
                                                              const double* tab_bnd, const uint8_t* tab_side) {
  const int t = blockIdx.x;
  const int K = c.odd_count[t];
  const int a0 = blockIdx.y * GATHER_ROWS;
  if (a0 >= K) return;
  extern __shared__ uint16_t odd_s[];  // the trial's odd list
  const int* odd = c.odd_fixed + (long long)t * c.S;
  for (int i = threadIdx.x; i < K; i += blockDim.x) odd_s[i] = (uint16_t)odd[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, a = a0 + (threadIdx.x >> 5);
  if (a >= K) return;
  const long long S = c.S;
  const long long src = odd_s[a];
  if (a < K - 1) {
    // row `src` of the table triangle holds (src, j), j > src, contiguously
    const double* trow = tab_pair + (src * S - src * (src + 1) / 2 - src - 1);
    double* orow = c.pair_len + c.pair_off[t] + ((long long)a * K - (long long)a * (a + 1) / 2 - a - 1);
    for (int b = a + 1 + lane; b < K; b += 32) orow[b] = trow[odd_s[b]];
  }
  if (tab_bnd && lane == 0) {
    const long long o = c.odd_off[t] + a;
    c.bnd_len[o] = tab_bnd[src];
    c.bnd_side[o] = tab_side[src];
  }
}

// k_gather_rows: the same gather, driven by the table instead of by the trials.  One CTA stages row `src`
// of the all-pairs table in shared memory once and serves every trial of its group in which cube `src`
// is odd, one warp per trial.  The warp walks the trial's parity words from `src` on: lane l of word w
// stands for cube 32 w + l, and if that cube is odd its length row_s[cube - src - 1] goes to position
// word_rank[w] + popc(word below l) of the trial's row - consecutive shared-memory reads, compacted
// coalesced writes, and no index loads at all.  The 8-byte table entries are read from HBM once per
// batch instead of once per trial (k_gather: one 32-byte sector per ~2 useful entries), so the DRAM
// traffic is the written lengths alone.
constexpr int GATHER_MAXK = 16;  // 32 parity words per register: S <= 16 * 32 * 32 = 16384 cubes

__global__ void __launch_bounds__(256) k_gather_rows(ComplexDev c, const double* tab_pair, const double* tab_bnd,
                                                     const uint8_t* tab_side, int T) {
  extern __shared__ double row_s[];  // entries (src, src + 1 ..)
  const int src = blockIdx.x;
  const long long S = c.S;
  const int n = (int)(S - src - 1);
  const double* trow = tab_pair + ((long long)src * S - (long long)src * (src + 1) / 2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) row_s[i] = trow[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int w0 = src >> 5, pwn = c.par_words;
  const uint32_t below = (1u << (src & 31)) - 1u, lt = (1u << lane) - 1u;
  const double* rs = row_s - src - 1;  // rs[cube] for cube > src
  for (int t = blockIdx.y * nwarp + warp; t < T; t += gridDim.y * nwarp) {
    const uint32_t* pw = c.parity + (long long)t * pwn;
    const uint16_t* wr = c.word_rank + (long long)t * pwn;
    const uint32_t word0 = pw[w0];
    if (!((word0 >> (src & 31)) & 1u)) continue;  // cube `src` is even in this trial (uniform over the warp)
    const long long a = wr[w0] + __popc(word0 & below);
    const long long K = c.odd_count[t];
    if (tab_bnd && lane == 0) {
      const long long o = c.odd_off[t] + a;
      c.bnd_len[o] = tab_bnd[src];
      c.bnd_side[o] = tab_side[src];
    }
    double* orow = c.pair_len + c.pair_off[t] + (a * K - a * (a + 1) / 2 - a - 1);  // indexed by position b
    // the trial's parity words and word ranks from w0 on, 32 words per register: one round trip to L2
    uint32_t pwr[GATHER_MAXK];
    uint32_t wrr[GATHER_MAXK];
    const int k0 = w0 >> 5;
#pragma unroll
    for (int k = 0; k < GATHER_MAXK; ++k) {
      const int idx = (k << 5) + lane;
      const bool in = k >= k0 && idx < pwn;
      pwr[k] = in ? pw[idx] : 0u;
      wrr[k] = in ? (uint32_t)wr[idx] : 0u;
    }
    const uint32_t after = ~(below | (1u << (src & 31)));  // word w0: cubes after src only
#pragma unroll
    for (int k = 0; k < GATHER_MAXK; ++k) {
      if (k < k0 || (k << 5) >= pwn) continue;  // uniform
      const int i0 = k == k0 ? (w0 & 31) : 0;
      const int i1 = min(32, pwn - (k << 5));
      for (int i = i0; i < i1; ++i) {
        const uint32_t word = __shfl_sync(0xffffffffu, pwr[k], i);
        const uint32_t base = __shfl_sync(0xffffffffu, wrr[k], i);
        const int w = (k << 5) + i;
        if (((w == w0 ? word & after : word) >> lane) & 1u) orow[base + __popc(word & lt)] = rs[(w << 5) + lane];
      }
    }
  }
}

}  // namespace fpb
