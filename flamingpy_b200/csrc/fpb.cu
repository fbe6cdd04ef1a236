// C ABI of the B200 Monte-Carlo EC trial backend (see include/fpb.h).
//
// Host side only: owns the device tables and result buffers of a handle, launches the
// kernels of fpb_kernels.cuh on the handle's stream and times every stage with CUDA events.
// There is no CPU fallback anywhere in this file: without a usable device fpb_create fails.
#include "../../include/fpb.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX: named ranges per pipeline stage for nsys / ncu --nvtx

#include "fpb_kernels.cuh"
#include "fpb_macro.cuh"
#include "fpb_uf.cuh"

using namespace fpb;

namespace {

// one NVTX range per pipeline stage (the reference has no tracing at all)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
               cudaGetErrorString(e_));                                                       \
      return fail(e_ == cudaErrorMemoryAllocation ? FPB_ERR_NOMEM : FPB_ERR_CUDA, buf_);      \
    }                                                                                         \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int ensure(size_t n) {
    if (n <= cap) return FPB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    CK(cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return FPB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
int upload(const T* host, size_t n, T** dev) {
  *dev = nullptr;
  if (n == 0) return FPB_OK;
  CK(cudaMalloc(dev, n * sizeof(T)));
  CK(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return FPB_OK;
}

struct ComplexHost {
  int S = 0, Q = 0, qoff = 0, par_words = 0, n_slots = 0, n_low = 0, n_high = 0;
  int* cube_qubit = nullptr;
  uint16_t* ninfo = nullptr;
  uint16_t* sbase = nullptr;
  int* slot_qubit = nullptr;
  int *low_cube = nullptr, *low_slot = nullptr, *high_cube = nullptr, *high_slot = nullptr;
  int d_off[12], s_off[12];
  DevBuf<uint32_t> parity;
  DevBuf<int> odd_count, odd_fixed, odd_ids;
  DevBuf<uint16_t> word_rank;
  DevBuf<long long> odd_off, pair_off;
  DevBuf<double> pair_len, bnd_len;
  DevBuf<uint8_t> bnd_side;
  long long total_odd = 0, total_pairs = 0;
  int max_odd = 0;  // largest odd count of the last batch
  DevBuf<int> knn;  // candidate partners of the last fpb_knn call
  DevBuf<double> knn_len;
  DevBuf<double> tab_pair, tab_bnd;  // all-pairs table of the static-weight path
  DevBuf<uint8_t> tab_side;
};

}  // namespace

struct fpb_handle {
  int device = 0;
  int sm_count = 0;
  int N = 0, Qtot = 0, bit_words = 0, ncx = 0;
  int S_max = 0, slots_max = 0;
  double delta = 0, p_swap = 0, weight_delta = 0, weight_multiplier = 1, step_factor = 1;
  int weight_method = 0, weight_integer = 0, label_mode = 0, reserve_sms = 0, length_mode = 0;
  long long chain_len = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t mark[2] = {nullptr, nullptr};
  int *adj_indptr = nullptr, *adj_indices = nullptr, *qubit_node = nullptr;
  int8_t* adj_vals = nullptr;
  ComplexHost cx[FPB_MAX_COMPLEX];
  DevBuf<double> x, hom, weights;
  DevBuf<uint8_t> state;
  DevBuf<uint32_t> bits, flips;
  DevBuf<int> item_off, counter, q_trial, q_cx, q_src, q_dst, path_qubits, path_len;
  DevBuf<long long> totals;
  // macronode layer (fpb_set_macro / fpb_run_macro_*)
  int* macro_partner = nullptr;
  int8_t* macro_sign = nullptr;
  DevBuf<uint8_t> m_state, m_body, m_red_state, m_bit_val;
  DevBuf<float> m_mean;
  DevBuf<double> m_z, m_val, m_p_phase, m_outcome;
  bool last_macro = false;
  bool last_labels = false;  // the last run had state labels (injected, generated or macronode): erasures are defined
  // device Union-Find decoder (fpb_uf_set_graph / fpb_uf_decode)
  int* uf_edge_a[FPB_MAX_COMPLEX] = {nullptr, nullptr};
  int* uf_edge_b[FPB_MAX_COMPLEX] = {nullptr, nullptr};
  int uf_nodes[FPB_MAX_COMPLEX] = {0, 0};
  // logical check on the device (fpb_set_planes / fpb_logical_check)
  int n_planes = 0;
  int *plane_off = nullptr, *plane_qubits = nullptr;
  long long flips_T = 0;  // trials the flips buffer holds a recovery for (0: none)
  DevBuf<uint8_t> fail;
  // sparse hand-off (fpb_knn_len / fpb_pairs_at / fpb_price)
  DevBuf<unsigned long long> trial_wmax, price_n;
  DevBuf<long long> price_y;
  DevBuf<double> price_c, pair_out;
  DevBuf<int> price_ws, price_top;
  DevBuf<uint8_t> price_active;
  DevBuf<int> price_idx;
  long long* totals_host = nullptr;  // pinned + mapped: kernels write it directly (no copy-engine hop)
  long long* totals_map = nullptr;   // device alias of totals_host
  long long T = 0;
  int last_stage = 0;
  bool have_rng = false;
  fpb_timings tm;
  // shortest-path launch configuration
  int warps = 0, ctas_per_sm = 0, tasks_per_item = 0;
  size_t paths_smem = 0;
  bool paths_wg = false;  // weights / tables in global memory (large lattices)
  bool paths_team = false;  // two warps per search (k_paths_team)
  int team_w = 0;           // > 0: W warps per search (k_paths_teamw)
  DevBuf<double> wscratch;
  // static-weight fast path (fpb_build_table)
  bool use_table = false;
  std::vector<int> qubit_degree;  // host copy: lattice degree of every syndrome qubit
  DevBuf<double> static_w;        // [Qtot]
  DevBuf<int> static_flag;
  int* static_flag_host = nullptr;  // pinned + mapped
  int* static_flag_map = nullptr;
};

namespace {

ComplexDev make_dev(const ComplexHost& c) {
  ComplexDev d;
  memset(&d, 0, sizeof d);
  d.S = c.S; d.Q = c.Q; d.qoff = c.qoff; d.par_words = c.par_words; d.n_slots = c.n_slots;
  d.n_low = c.n_low; d.n_high = c.n_high;
  d.cube_qubit = c.cube_qubit; d.ninfo = c.ninfo; d.sbase = c.sbase; d.slot_qubit = c.slot_qubit;
  d.low_cube = c.low_cube; d.low_slot = c.low_slot; d.high_cube = c.high_cube; d.high_slot = c.high_slot;
  memcpy(d.d_off, c.d_off, sizeof d.d_off);
  memcpy(d.s_off, c.s_off, sizeof d.s_off);
  d.parity = c.parity.p; d.odd_count = c.odd_count.p; d.odd_fixed = c.odd_fixed.p; d.word_rank = c.word_rank.p;
  d.odd_off = c.odd_off.p; d.pair_off = c.pair_off.p; d.odd_ids = c.odd_ids.p;
  d.pair_len = c.pair_len.p; d.bnd_len = c.bnd_len.p; d.bnd_side = c.bnd_side.p;
  return d;
}

typedef void (*paths_fn)(const PathsParams);
typedef void (*toggle_fn)(const PathsParams);

// kernel variant by lattice size: NCH = 16-cube chunks each lane scans, NV = cubes relaxed per
// lane per batch, MINB = CTAs per SM the register budget is sized for
// (NV = 2 variants are compiled for at most 384 threads so that they get 170 registers)
// FPB_LENGTH_RX_TRUNC: one warp per search, weights in global memory, int lengths carried along
paths_fn pick_ilen_kernel(int S) {
  if (S <= 1024) return k_paths<2, 1, 256, 1, true, true>;
  if (S <= 2560) return k_paths<5, 1, 256, 1, true, true>;
  if (S <= 4096) return k_paths<8, 1, 256, 1, true, true>;
  if (S <= 8192) return k_paths<16, 1, 256, 1, true, true>;
  return k_paths<32, 1, 256, 1, true, true>;
}
paths_fn pick_paths_kernel(int S, bool wg, int* max_warps) {
  if (wg) {  // weights and face tables in global memory: lattices that do not fit shared memory
    *max_warps = 8;
    if (S <= 4096) return k_paths<8, 1, 256, 1, true>;
    if (S <= 8192) return k_paths<16, 1, 256, 1, true>;
    return k_paths<32, 1, 256, 1, true>;
  }
  *max_warps = 16;
  if (S <= 1024) return k_paths<2, 1, 512, 2, false>;
  if (S <= 1536) return k_paths<3, 1, 512, 1, false>;
  *max_warps = 12;
  if (S <= 2560) return k_paths<5, 2, 384, 1, false>;
  if (S <= 4096) return k_paths<8, 2, 384, 1, false>;
  return k_paths<16, 2, 384, 1, false>;
}
// two-warp-team variant (shared-memory lattices with more than 1024 cubes): NCH2 = chunks per warp
paths_fn pick_team_kernel(int S, bool wg) {
  // FPB_TEAM_RM=1: residue-major bucket bytes + stride-dealt rounds + lean verify (team_search_rm).  Measured on
  // B200 (profiles/r2_k_paths_rm_experiment.md): 45 % fewer bank-conflict replays and 10 % fewer instructions,
  // but 5-20 % slower than the round-1 search on every lattice - the kernel is bound by the dependent
  // latency of a round at 16 warps per SM, not by the shared-memory pipe - so it stays opt-in.
  static const bool rm = getenv("FPB_TEAM_RM") != nullptr;
  if (rm) {
    const int rows = paths_bkt_row(S);  // chunks to scan; 64 lanes x NCHT (odd) chunks each
    if (wg) {
      if (rows <= 64 * 3) return k_paths_team<3, 2, 512, 1, true, true>;
      if (rows <= 64 * 5) return k_paths_team<5, 2, 512, 1, true, true>;
      return k_paths_team<9, 2, 512, 1, true, true>;
    }
    if (rows <= 64 * 3) return k_paths_team<3, 2, 512, 1, false, true>;
    if (rows <= 64 * 5) return k_paths_team<5, 2, 512, 1, false, true>;
    return k_paths_team<9, 2, 512, 1, false, true>;
  }
  if (wg) {
    if (S <= 4096) return k_paths_team<4, 2, 512, 1, true>;
    return k_paths_team<8, 2, 512, 1, true>;
  }
  if (S <= 2048) return k_paths_team<2, 2, 512, 1, false>;
  if (S <= 3072) return k_paths_team<3, 2, 512, 1, false>;
  if (S <= 4096) return k_paths_team<4, 2, 512, 1, false>;
  return k_paths_team<8, 2, 512, 1, false>;
}
// W-warp teams (global-memory weights) for lattices of which at most four searches fit an SM
paths_fn pick_teamw_kernel(int S, int W) {
  if (W == 4) {
    if (S <= 4096) return k_paths_teamw<2, 4, 512>;
    if (S <= 6144) return k_paths_teamw<3, 4, 512>;
    if (S <= 8192) return k_paths_teamw<4, 4, 512>;
  } else if (W == 8) {
    if (S <= 8192) return k_paths_teamw<2, 8, 512>;
    if (S <= 12288) return k_paths_teamw<3, 8, 512>;
    if (S <= 16384) return k_paths_teamw<4, 8, 512>;
  } else if (W == 16) {
    if (S <= 8192) return k_paths_teamw<1, 16, 512>;
    if (S <= 16384) return k_paths_teamw<2, 16, 512>;
  }
  return nullptr;
}
toggle_fn pick_toggle_kernel(int S, bool wg) {
  if (wg) {
    if (S <= 4096) return k_path_toggle<8, true>;
    if (S <= 8192) return k_path_toggle<16, true>;
    return k_path_toggle<32, true>;
  }
  if (S <= 1024) return k_path_toggle<2, false>;
  if (S <= 1536) return k_path_toggle<3, false>;
  if (S <= 2560) return k_path_toggle<5, false>;
  if (S <= 4096) return k_path_toggle<8, false>;
  return k_path_toggle<16, false>;
}

int configure_variant(fpb_handle* h, bool wg) {
  int max_optin = 0;
  CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  int smem_per_sm = 0;
  CK(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
  const size_t cta_b = wg ? 0 : paths_cta_bytes(h->S_max, h->slots_max);
  const size_t warp_b = paths_warp_bytes(h->S_max, false);
  int max_warps = 16;
  paths_fn fn = pick_paths_kernel(h->S_max, wg, &max_warps);
  cudaFuncAttributes fattr;
  CK(cudaFuncGetAttributes(&fattr, fn));
  max_optin -= (int)fattr.sharedSizeBytes;  // static shared memory counts against the opt-in limit
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  int best = 0;
  for (int ctas = 1; ctas <= 4; ++ctas) {
    // per-CTA budget when `ctas` CTAs share an SM (1 KB reserved per CTA by the runtime)
    size_t budget = (size_t)smem_per_sm / ctas - 1024;
    if (budget > (size_t)max_optin) budget = max_optin;
    if (budget < cta_b + warp_b) break;
    int w = (int)((budget - cta_b) / warp_b);
    if (w > max_warps) w = max_warps;
    int resident = 0;  // what registers + shared memory really allow
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fn, w * 32, cta_b + (size_t)w * warp_b));
    if (resident > ctas) resident = ctas;
    if (w * resident > best) {
      best = w * resident;
      h->warps = w;
      h->ctas_per_sm = resident;
    }
  }
  if (best == 0) return FPB_OK;  // does not fit; caller decides
  h->paths_wg = wg;
  h->paths_smem = cta_b + (size_t)h->warps * warp_b;
  h->tasks_per_item = h->warps * 4;
  return FPB_OK;
}

int configure_paths(fpb_handle* h) {
  // shared-memory budget: one CTA-wide copy of the trial's weights and the face tables, plus a
  // distance array, bucket bytes and a staging list per warp (= per concurrent search).  Lattices
  // whose weights do not fit next to at least one search fall back to the variant that keeps
  // weights and tables in global memory (L2) and only the per-search state in shared memory.
  h->warps = 0;
  h->ctas_per_sm = 0;
  if (h->S_max > 16 * 32 * MAX_CHUNKS || h->S_max > 65535)
    return fail(FPB_ERR_UNSUPPORTED, "lattice too large for the shortest-path kernel's bucket scan");
  if (h->length_mode == FPB_LENGTH_RX_TRUNC) {
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    paths_fn fn = pick_ilen_kernel(h->S_max);
    cudaFuncAttributes fattr;
    CK(cudaFuncGetAttributes(&fattr, fn));
    const size_t room = (size_t)max_optin - fattr.sharedSizeBytes;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room));
    const size_t warp_b = paths_warp_bytes(h->S_max, false) + align16((size_t)h->S_max * 4);
    int w = (int)(room / warp_b);
    if (w > 8) w = 8;
    if (w < 1) return fail(FPB_ERR_UNSUPPORTED, "lattice too large for the int-length shortest-path kernel");
    h->warps = w;
    h->ctas_per_sm = 1;
    h->paths_wg = true;
    h->paths_team = false;
    h->team_w = 0;
    h->paths_smem = (size_t)w * warp_b;
    h->tasks_per_item = w * 4;
    return FPB_OK;
  }
  int rc = FPB_OK;
  if (h->S_max <= 16 * 32 * 16 && !getenv("FPB_FORCE_WG")) rc = configure_variant(h, false);
  if (rc) return rc;
  h->paths_team = false;
  if (h->warps > 0 && !h->paths_wg && h->S_max > 1024 && h->S_max <= 8192 && h->ctas_per_sm == 1 &&
      !getenv("FPB_NO_TEAM")) {
    // two warps per search on the same shared-memory states (at most 8 teams = 512 threads per
    // CTA).  If the CTA-wide weight copy leaves room for fewer than 8 searches, the team variant
    // that keeps weights and face tables in global memory holds more of them.
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const size_t warp_b = paths_warp_bytes(h->S_max, false);
    for (int wg = 0; wg <= 1; ++wg) {
      paths_fn tf = pick_team_kernel(h->S_max, wg != 0);
      cudaFuncAttributes fattr;
      CK(cudaFuncGetAttributes(&fattr, tf));
      const size_t cta_b = wg ? 0 : paths_cta_bytes(h->S_max, h->slots_max);
      const size_t room = (size_t)max_optin - fattr.sharedSizeBytes;
      if (room < cta_b + warp_b) continue;
      int teams = (int)((room - cta_b) / warp_b);
      if (teams > 8) teams = 8;
      if (h->paths_team && teams <= h->warps) continue;  // the shared-weights variant already holds as many
      const size_t need = cta_b + (size_t)teams * warp_b;
      // the attribute is per function and process-wide: always the full opt-in room, so that a handle with
      // a smaller lattice in the same template bucket cannot lower another live handle's limit
      CK(cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room));
      int resident = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, tf, teams * 64, need));
      if (resident < 1) continue;
      h->paths_team = true;
      h->paths_wg = wg != 0;
      h->warps = teams;  // searches in flight per CTA
      h->paths_smem = need;
      h->tasks_per_item = teams * 4;
    }
  }
  if (h->warps == 0) rc = configure_variant(h, true);
  if (rc) return rc;
  if (h->warps == 0)
    return fail(FPB_ERR_UNSUPPORTED, "lattice too large for the shared-memory shortest-path kernel");
  // Very large lattices: one to five searches fit an SM, so each search gets 16 / 8 / 4 warps
  // (FPB_TEAM_W=0 disables, FPB_TEAM_W=4|8|16 forces a width).
  h->team_w = 0;
  {
    const char* ev = getenv("FPB_TEAM_W");
    const int force_w = ev ? atoi(ev) : -1;
    if (force_w != 0 && h->S_max > 1024) {
      int max_optin = 0;
      CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
      const size_t warp_b = paths_warp_bytes(h->S_max, false);
      const int fit0 = (int)(((size_t)max_optin - 2048) / warp_b);
      const int W = force_w > 0 ? force_w : (fit0 <= 1 ? 16 : fit0 == 2 ? 8 : fit0 <= 5 ? 4 : 0);
      paths_fn tf = W ? pick_teamw_kernel(h->S_max, W) : nullptr;
      if (tf) {
        cudaFuncAttributes fattr;
        CK(cudaFuncGetAttributes(&fattr, tf));
        int teams = (int)(((size_t)max_optin - fattr.sharedSizeBytes) / warp_b);
        if (teams > 16 / W) teams = 16 / W;
        if (teams >= 1) {
          const size_t need = (size_t)teams * warp_b;
          CK(cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((size_t)max_optin - fattr.sharedSizeBytes)));
          int resident = 0;
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, tf, teams * W * 32, need));
          if (resident >= 1) {
            h->team_w = W;
            h->paths_team = true;
            h->paths_wg = true;
            h->warps = teams;
            h->ctas_per_sm = 1;
            h->paths_smem = need;
            h->tasks_per_item = teams * 4;
          }
        }
      }
    }
  }
  return FPB_OK;
}

int ensure_batch(fpb_handle* h, long long T) {
  int rc;
  if ((rc = h->x.ensure((size_t)T * 2 * h->N))) return rc;
  if ((rc = h->state.ensure((size_t)T * h->N))) return rc;
  if ((rc = h->hom.ensure((size_t)T * h->Qtot))) return rc;
  if ((rc = h->weights.ensure((size_t)T * h->Qtot))) return rc;
  if ((rc = h->bits.ensure((size_t)T * h->bit_words))) return rc;
  if ((rc = h->item_off.ensure((size_t)T * h->ncx + 1))) return rc;
  if ((rc = h->counter.ensure(1))) return rc;
  if ((rc = h->totals.ensure(3 * MAXC + 1))) return rc;
  for (int c = 0; c < h->ncx; ++c) {
    ComplexHost& cx = h->cx[c];
    if ((rc = cx.parity.ensure((size_t)T * cx.par_words))) return rc;
    if ((rc = cx.odd_count.ensure((size_t)T))) return rc;
    if ((rc = cx.odd_fixed.ensure((size_t)T * cx.S))) return rc;
    if (h->use_table && getenv("FPB_GATHER_ROWS") && (rc = cx.word_rank.ensure((size_t)T * cx.par_words))) return rc;
    if ((rc = cx.odd_off.ensure((size_t)T + 1))) return rc;
    if ((rc = cx.pair_off.ensure((size_t)T + 1))) return rc;
  }
  return FPB_OK;
}

int launch_paths(fpb_handle* h, long long T, int* launches) {
  cudaStream_t st = h->stream;
  const long long total_items = h->totals_host[2 * MAXC];
  if (total_items <= 0) return FPB_OK;
  PathsParams pp;
  memset(&pp, 0, sizeof pp);
  pp.n_complex = h->ncx; pp.T = (int)T; pp.Qtot = h->Qtot;
  pp.tasks_per_item = h->tasks_per_item;
  pp.S_max = h->S_max; pp.slots_max = h->slots_max; pp.warps = h->warps;
  pp.step_factor = h->step_factor;
  pp.weights = h->weights.p;
  pp.item_off = h->item_off.p;
  pp.total_items = (int)total_items;
  pp.counter = h->counter.p;
  for (int c = 0; c < h->ncx; ++c) pp.cx[c] = make_dev(h->cx[c]);
  CK(cudaMemsetAsync(h->counter.p, 0, sizeof(int), st));
  long long grid = (long long)(h->sm_count - h->reserve_sms) * h->ctas_per_sm;
  if (grid > total_items) grid = total_items;
  if (h->paths_wg) {
    int rc2 = h->wscratch.ensure((size_t)grid * h->slots_max);
    if (rc2) return rc2;
    pp.wscratch = h->wscratch.p;
  }
  int mw_unused = 0;
  if (h->length_mode == FPB_LENGTH_RX_TRUNC)
    pick_ilen_kernel(h->S_max)<<<(unsigned)grid, h->warps * 32, h->paths_smem, st>>>(pp);
  else if (h->team_w)
    pick_teamw_kernel(h->S_max, h->team_w)<<<(unsigned)grid, h->warps * h->team_w * 32, h->paths_smem, st>>>(pp);
  else if (h->paths_team)
    pick_team_kernel(h->S_max, h->paths_wg)<<<(unsigned)grid, h->warps * 64, h->paths_smem, st>>>(pp);
  else
    pick_paths_kernel(h->S_max, h->paths_wg, &mw_unused)<<<(unsigned)grid, h->warps * 32, h->paths_smem, st>>>(pp);
  ++*launches;
  h->tm.graph_ctas = (int)grid;
  return FPB_OK;
}

// scan the odd counts, read the totals back, size the ragged outputs, compact the odd lists
int post_syndrome(fpb_handle* h, long long T, int stage, int* launches) {
  cudaStream_t st = h->stream;
  ScanParams sc;
  sc.n_complex = h->ncx;
  sc.tasks_per_item = h->tasks_per_item > 0 ? h->tasks_per_item : 1;
  for (int c = 0; c < h->ncx; ++c) {
    sc.cx[c] = make_dev(h->cx[c]);
    sc.has_boundary[c] = h->cx[c].n_low > 0;
  }
  sc.item_off = h->item_off.p;
  // the totals go straight to mapped host memory: a D2H copy would queue on the copy engine behind
  // another handle's bulk result transfer and stall this handle's next launch
  sc.totals = h->totals_map;
  k_scan<<<1, 1024, 0, st>>>(sc, (int)T);
  ++*launches;
  CK(cudaStreamSynchronize(st));
  int rc;
  for (int c = 0; c < h->ncx; ++c) {
    ComplexHost& cx = h->cx[c];
    cx.total_odd = h->totals_host[c];
    cx.total_pairs = h->totals_host[MAXC + c];
    cx.max_odd = (int)h->totals_host[2 * MAXC + 1 + c];
    if ((rc = cx.odd_ids.ensure((size_t)cx.total_odd + 1))) return rc;
    if ((rc = cx.bnd_len.ensure((size_t)cx.total_odd + 1))) return rc;
    if ((rc = cx.bnd_side.ensure((size_t)cx.total_odd + 1))) return rc;
    if (stage >= FPB_STAGE_GRAPH)
      if ((rc = cx.pair_len.ensure((size_t)cx.total_pairs + 1))) return rc;
    k_compact_odd<<<(unsigned)T, 256, 0, st>>>(make_dev(cx), (int)T);
    ++*launches;
  }
  return FPB_OK;
}

// source of a batch: injected draws + labels, Philox draws + labels, injected bits, Philox i.i.d. bits
enum { SRC_INJECTED = 0, SRC_GEN = 1, SRC_BITS = 2, SRC_IID = 3, SRC_WEIGHTED = 4, SRC_MACRO = 5, SRC_MACRO_GEN = 6 };

int run_pipeline(fpb_handle* h, long long T, int stage, int source, uint64_t seed, long long first_trial,
                 double p_err = 0.0) {
  const bool gen = source == SRC_GEN;
  if (T <= 0) return fail(FPB_ERR_ARG, "n_trials must be positive");
  if (T > 0x7fffffffll / (h->ncx > 0 ? h->ncx : 1) / 2) return fail(FPB_ERR_ARG, "n_trials too large for one batch");
  if (stage < FPB_STAGE_NOISE || stage > FPB_STAGE_GRAPH) return fail(FPB_ERR_ARG, "bad stage");
  cudaStream_t st = h->stream;
  memset(&h->tm, 0, sizeof h->tm);
  h->T = 0;  // no valid result until every stage below has completed
  h->last_stage = 0;
  h->flips_T = 0;
  for (int c = 0; c < h->ncx; ++c) h->cx[c].total_odd = h->cx[c].total_pairs = 0;
  int launches = 0;

  NvtxRange nvtx_run("fpb:run_pipeline");
  CK(cudaEventRecord(h->ev[0], st));
  nvtxRangePushA("fpb:gen");
  if (gen) {
    GenParams g;
    g.N = h->N;
    g.seed_lo = (uint32_t)seed;
    g.seed_hi = (uint32_t)(seed >> 32);
    g.first_trial = first_trial;
    g.chain_len = h->chain_len;
    g.fresh = h->label_mode == FPB_MODE_FRESH;
    g.p_swap = h->p_swap;
    double th = h->p_swap * 4294967296.0;
    g.p_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
    // float32-rounded standard deviations, noise/cv.py:280-289
    g.sigma_q_p = (double)(float)(1.0 / pow(2.0 * h->delta, 0.5));
    g.sigma_q_gkp = (double)(float)pow(h->delta / 2.0, 0.5);
    g.sigma_p = (double)(float)pow(h->delta / 2.0, 0.5);
    g.x = h->x.p;
    g.state = h->state.p;
    long long total = T * (long long)h->N;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)h->sm_count * 32;
    k_gen<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(g, T);
    ++launches;
  }
  if (source == SRC_IID) {
    IidParams g;
    g.Qtot = h->Qtot; g.bit_words = h->bit_words;
    g.seed_lo = (uint32_t)seed; g.seed_hi = (uint32_t)(seed >> 32);
    g.first_trial = first_trial;
    g.p_err = p_err;
    g.qubit_node = h->qubit_node;
    g.bits = h->bits.p;
    const long long total = T * (long long)h->bit_words;
    const long long blocks = (total + 255) / 256, cap = (long long)h->sm_count * 32;
    k_iid<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(g, T);
    ++launches;
  }
  const bool macro = source == SRC_MACRO || source == SRC_MACRO_GEN;
  h->last_macro = macro;
  h->last_labels = macro || source == SRC_INJECTED || source == SRC_GEN;
  if (source == SRC_MACRO_GEN) {
    MacroGenParams g;
    g.M = 4 * h->N;
    g.seed_lo = (uint32_t)seed; g.seed_hi = (uint32_t)(seed >> 32);
    g.first_trial = first_trial; g.chain_len = h->chain_len;
    g.fresh = h->label_mode == FPB_MODE_FRESH;
    g.p_swap = h->p_swap;
    const double th = h->p_swap * 4294967296.0;
    g.p_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
    g.state = h->m_state.p; g.mean = h->m_mean.p; g.z = h->m_z.p;
    const long long total = T * 4ll * h->N, blocks = (total + 255) / 256, cap = (long long)h->sm_count * 32;
    k_macro_gen<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(g, T);
    ++launches;
  }
  nvtxRangePop();
  CK(cudaEventRecord(h->ev[1], st));
  nvtxRangePushA("fpb:noise+weights");
  if (macro) {
    MacroParams mp;
    mp.N = h->N; mp.partner = h->macro_partner; mp.sign = h->macro_sign;
    mp.sd = (float)pow(h->delta / 2.0, 0.5);
    mp.delta = h->delta;
    mp.state = h->m_state.p; mp.mean = h->m_mean.p; mp.z = h->m_z.p;
    mp.body = h->m_body.p; mp.val = h->m_val.p;
    mp.red_state = h->m_red_state.p; mp.bit_val = h->m_bit_val.p; mp.p_phase = h->m_p_phase.p; mp.outcome = h->m_outcome.p;
    const long long total = T * (long long)h->N, cap = (long long)h->sm_count * 32;
    const long long b256 = (total + 255) / 256, b128 = (total + 127) / 128;
    k_macro_measure<<<(unsigned)(b256 < cap ? b256 : cap), 256, 0, st>>>(mp, T);
    k_macro_reduce<<<(unsigned)(b128 < 2 * cap ? b128 : 2 * cap), 128, 0, st>>>(mp, T);
    MacroWeightParams wp;
    wp.N = h->N; wp.Qtot = h->Qtot; wp.bit_words = h->bit_words;
    wp.adj_indptr = h->adj_indptr; wp.adj_indices = h->adj_indices; wp.qubit_node = h->qubit_node;
    wp.method = h->weight_method; wp.integer = h->weight_integer; wp.multiplier = h->weight_multiplier;
    wp.red_state = h->m_red_state.p; wp.bit_val = h->m_bit_val.p; wp.p_phase = h->m_p_phase.p; wp.outcome = h->m_outcome.p;
    wp.hom = h->hom.p; wp.weights = h->weights.p; wp.bits = h->bits.p;
    const long long bpt = (h->Qtot + 255) / 256;
    k_macro_weights<<<(unsigned)(bpt * T), 256, 0, st>>>(wp);
    launches += 3;
  } else if (source == SRC_WEIGHTED) {
    // weights and bits were given: nothing to compute before the syndrome
    CK(cudaMemsetAsync(h->hom.p, 0, (size_t)T * h->Qtot * sizeof(double), st));
  } else if (source == SRC_BITS || source == SRC_IID) {
    // bits-only trials: no homodyne outcomes; every syndrome qubit carries the uniform weight 1
    const long long nw = T * (long long)h->Qtot;
    const long long blocks = (nw + 255) / 256, cap = (long long)h->sm_count * 32;
    k_fill<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(h->weights.p, nw, 1.0);
    CK(cudaMemsetAsync(h->hom.p, 0, (size_t)nw * sizeof(double), st));
    ++launches;
  } else {
    NoiseParams p;
    p.N = h->N; p.Qtot = h->Qtot; p.bit_words = h->bit_words;
    p.adj_indptr = h->adj_indptr; p.adj_indices = h->adj_indices; p.adj_vals = h->adj_vals;
    p.qubit_node = h->qubit_node;
    p.delta_w = h->weight_delta; p.method = h->weight_method; p.integer = h->weight_integer;
    p.multiplier = h->weight_multiplier;
    p.x = h->x.p; p.state = h->state.p; p.hom = h->hom.p; p.weights = h->weights.p; p.bits = h->bits.p;
    long long bpt = (h->Qtot + 255) / 256;
    k_noise<<<(unsigned)(bpt * T), 256, 0, st>>>(p);
    ++launches;
  }
  nvtxRangePop();
  CK(cudaEventRecord(h->ev[2], st));
  if (stage >= FPB_STAGE_SYNDROME) {
    NvtxRange nvtx_s("fpb:syndrome");
    SyndromeParams sp;
    sp.n_complex = h->ncx; sp.bit_words = h->bit_words; sp.bits = h->bits.p;
    for (int c = 0; c < h->ncx; ++c) sp.cx[c] = make_dev(h->cx[c]);
    k_syndrome<<<dim3((unsigned)T, h->ncx), 256, 0, st>>>(sp);
    ++launches;
    int rcs = post_syndrome(h, T, stage, &launches);
    if (rcs) return rcs;
  }
  CK(cudaEventRecord(h->ev[3], st));
  if (stage >= FPB_STAGE_GRAPH) {
    NvtxRange nvtx_g("fpb:matching_graph");
    if (h->warps == 0) return fail(FPB_ERR_UNSUPPORTED, "shortest-path kernel not configured for this lattice");
    if (h->use_table) {
      // static weights: gather the matching graph from the all-pairs table
      if (h->weight_method == FPB_WEIGHT_BLUEPRINT) {
        *h->static_flag_host = 0;
        k_check_static<<<h->sm_count * 4, 256, 0, st>>>(h->weights.p, h->static_w.p, h->Qtot, T, h->static_flag_map);
        ++launches;
      }
      for (int c = 0; c < h->ncx; ++c) {
        ComplexHost& cx = h->cx[c];
        if (cx.total_odd == 0) continue;
        // FPB_GATHER_ROWS=1: the table-driven gather.  It reads the table once per batch (DRAM reads 2.5 GB ->
        // 0.3 GB per 256 d=21 trials) but writes short unaligned segments: 2.76 TB/s against 3.4-3.7 TB/s of
        // the trial-driven kernel below (profiles/r2_gather_experiment.md), so it stays opt-in.
        static const bool by_rows = getenv("FPB_GATHER_ROWS") != nullptr;
        if (by_rows && T >= 8) {
          // table-driven: every table row is staged once per batch (k_gather_rows)
          const size_t smem = (size_t)cx.S * sizeof(double);
          CK(cudaFuncSetAttribute(k_gather_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          // one CTA per table row serves all trials (8 warps = 8 trials in flight), so that staging the row
          // (<= 8 S bytes) is paid once per batch; FPB_GATHER_GROUPS splits the trials over more CTAs per row
          static const int env_groups = getenv("FPB_GATHER_GROUPS") ? atoi(getenv("FPB_GATHER_GROUPS")) : 0;
          int groups = env_groups > 0 ? env_groups : (int)(T / 512);
          if (groups < 1) groups = 1;
          if (groups > 16) groups = 16;
          k_gather_rows<<<dim3((unsigned)cx.S, (unsigned)groups), 256, smem, st>>>(
              make_dev(cx), cx.tab_pair.p, cx.n_low > 0 ? cx.tab_bnd.p : nullptr, cx.tab_side.p, (int)T);
        } else {
          // rows = odd cubes: no CTA beyond the largest odd count of the batch
          dim3 grid((unsigned)T, (unsigned)((cx.max_odd + GATHER_ROWS - 1) / GATHER_ROWS));
          k_gather<<<grid, GATHER_ROWS * 32, (size_t)cx.S * 2, st>>>(make_dev(cx), cx.tab_pair.p,
                                                                    cx.n_low > 0 ? cx.tab_bnd.p : nullptr, cx.tab_side.p);
        }
        ++launches;
      }
    } else {
      int rc2 = launch_paths(h, T, &launches);
      if (rc2) return rc2;
    }
  }
  CK(cudaEventRecord(h->ev[4], st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&h->tm.gen_ms, h->ev[0], h->ev[1]));
  CK(cudaEventElapsedTime(&h->tm.noise_ms, h->ev[1], h->ev[2]));
  CK(cudaEventElapsedTime(&h->tm.syndrome_ms, h->ev[2], h->ev[3]));
  CK(cudaEventElapsedTime(&h->tm.graph_ms, h->ev[3], h->ev[4]));
  CK(cudaEventElapsedTime(&h->tm.total_ms, h->ev[0], h->ev[4]));
  if (stage >= FPB_STAGE_GRAPH && h->use_table && h->weight_method == FPB_WEIGHT_BLUEPRINT && *h->static_flag_host)
    return fail(FPB_ERR_STATE, "static-weight table in use but a trial's labels are not all p-squeezed");
  h->T = T;
  h->last_stage = stage;
  h->tm.launches = launches;
  h->tm.graph_warps_per_cta = h->warps;
  h->tm.graph_smem_bytes = (int)h->paths_smem;
  return FPB_OK;
}

}  // namespace

extern "C" {

const char* fpb_last_error(void) { return g_err.c_str(); }
int fpb_abi_version(void) { return FPB_ABI_VERSION; }

int fpb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(FPB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  return n;
}

int fpb_create(const fpb_config* cfg, fpb_handle** out) {
  if (!cfg || !out) return fail(FPB_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != FPB_ABI_VERSION) return fail(FPB_ERR_ARG, "ABI version mismatch");
  if (cfg->n_complex < 1 || cfg->n_complex > FPB_MAX_COMPLEX) return fail(FPB_ERR_ARG, "n_complex must be 1 or 2");
  if (cfg->n_nodes <= 0 || !cfg->adj_indptr || !cfg->adj_indices || !cfg->adj_vals)
    return fail(FPB_ERR_ARG, "missing lattice adjacency");
  if (!(cfg->delta > 0)) return fail(FPB_ERR_ARG, "delta must be positive");
  if (cfg->p_swap < 0 || cfg->p_swap > 1) return fail(FPB_ERR_ARG, "p_swap must be in [0, 1]");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(FPB_ERR_CUDA, std::string("no usable CUDA device (no CPU fallback exists): ") +
                                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(FPB_ERR_ARG, "device ordinal out of range");
  CK(cudaSetDevice(cfg->device));
  fpb_handle* h = new fpb_handle();
  h->device = cfg->device;
  int rc = FPB_OK;
  auto bail = [&](int code) {
    fpb_destroy(h);
    return code;
  };
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return bail(fail(FPB_ERR_CUDA, "cudaGetDeviceProperties failed"));
  h->sm_count = prop.multiProcessorCount;
  h->N = cfg->n_nodes;
  h->ncx = cfg->n_complex;
  h->delta = cfg->delta;
  h->p_swap = cfg->p_swap;
  h->weight_method = cfg->weight_method;
  h->weight_delta = cfg->weight_delta > 0 ? cfg->weight_delta : cfg->delta;
  h->weight_integer = cfg->weight_integer;
  h->weight_multiplier = cfg->weight_multiplier != 0 ? cfg->weight_multiplier : 1.0;
  h->label_mode = cfg->label_mode;
  h->chain_len = cfg->chain_len;
  h->length_mode = cfg->length_mode;
  if (h->length_mode != FPB_LENGTH_FLOAT && h->length_mode != FPB_LENGTH_RX_TRUNC) {
    fpb_destroy(h);
    return fail(FPB_ERR_ARG, "unknown length_mode");
  }
  h->reserve_sms = cfg->reserve_sms < 0 ? 0 : cfg->reserve_sms;
  if (h->reserve_sms >= h->sm_count) h->reserve_sms = h->sm_count - 1;
  h->step_factor = cfg->delta_step > 0 ? cfg->delta_step : 2.0;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(FPB_ERR_CUDA, "stream creation failed"));
  for (auto& ev : h->ev)
    if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(FPB_ERR_CUDA, "event creation failed"));
  for (auto& ev : h->mark)
    if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(FPB_ERR_CUDA, "event creation failed"));
  if (cudaHostAlloc(&h->totals_host, (3 * MAXC + 1) * sizeof(long long), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&h->totals_map, h->totals_host, 0) != cudaSuccess)
    return bail(fail(FPB_ERR_CUDA, "mapped pinned allocation failed"));
  memset(h->totals_host, 0, (3 * MAXC + 1) * sizeof(long long));
  if (cudaHostAlloc(&h->static_flag_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&h->static_flag_map, h->static_flag_host, 0) != cudaSuccess)
    return bail(fail(FPB_ERR_CUDA, "mapped pinned allocation failed"));
  *h->static_flag_host = 0;

  const int N = h->N;
  const int nnz = cfg->adj_indptr[N];
  if ((rc = upload(cfg->adj_indptr, (size_t)N + 1, &h->adj_indptr))) return bail(rc);
  if ((rc = upload(cfg->adj_indices, (size_t)nnz, &h->adj_indices))) return bail(rc);
  if ((rc = upload(cfg->adj_vals, (size_t)nnz, &h->adj_vals))) return bail(rc);

  std::vector<int> qnode;
  int qoff = 0;
  for (int c = 0; c < h->ncx; ++c) {
    const fpb_complex_desc& d = cfg->complex_[c];
    ComplexHost& cx = h->cx[c];
    if (d.n_cubes <= 0 || d.n_qubits <= 0 || !d.qubit_node || !d.cube_qubit || !d.ninfo || !d.sbase || !d.slot_qubit)
      return bail(fail(FPB_ERR_ARG, "incomplete complex description"));
    if (d.n_cubes > 65535 || d.n_slots > 65535)
      return bail(fail(FPB_ERR_UNSUPPORTED, "complex exceeds the 16-bit cube / slot index range"));
    cx.S = d.n_cubes; cx.Q = d.n_qubits; cx.qoff = qoff; cx.par_words = (d.n_cubes + 31) / 32;
    cx.n_slots = d.n_slots; cx.n_low = d.n_low; cx.n_high = d.n_high;
    if ((cx.n_low > 0) != (cx.n_high > 0)) return bail(fail(FPB_ERR_ARG, "LOW and HIGH connectors must come together"));
    memcpy(cx.d_off, d.d_off, sizeof cx.d_off);
    memcpy(cx.s_off, d.s_off, sizeof cx.s_off);
    for (int q = 0; q < d.n_qubits; ++q) {
      if (d.qubit_node[q] < 0 || d.qubit_node[q] >= N) return bail(fail(FPB_ERR_ARG, "qubit node index out of range"));
      qnode.push_back(d.qubit_node[q]);
      h->qubit_degree.push_back(cfg->adj_indptr[d.qubit_node[q] + 1] - cfg->adj_indptr[d.qubit_node[q]]);
    }
    std::vector<uint16_t> sb16(d.n_cubes);
    for (int s = 0; s < d.n_cubes; ++s) {
      if (d.sbase[s] < 0 || d.sbase[s] > 65535) return bail(fail(FPB_ERR_ARG, "sbase out of range"));
      sb16[s] = (uint16_t)d.sbase[s];
    }
    if ((rc = upload(d.cube_qubit, (size_t)d.n_cubes * 6, &cx.cube_qubit))) return bail(rc);
    if ((rc = upload(d.ninfo, (size_t)d.n_cubes, &cx.ninfo))) return bail(rc);
    if ((rc = upload(sb16.data(), (size_t)d.n_cubes, &cx.sbase))) return bail(rc);
    if ((rc = upload(d.slot_qubit, (size_t)d.n_slots, &cx.slot_qubit))) return bail(rc);
    if ((rc = upload(d.low_cube, (size_t)d.n_low, &cx.low_cube))) return bail(rc);
    if ((rc = upload(d.low_slot, (size_t)d.n_low, &cx.low_slot))) return bail(rc);
    if ((rc = upload(d.high_cube, (size_t)d.n_high, &cx.high_cube))) return bail(rc);
    if ((rc = upload(d.high_slot, (size_t)d.n_high, &cx.high_slot))) return bail(rc);
    qoff += d.n_qubits;
    if (d.n_cubes > h->S_max) h->S_max = d.n_cubes;
    if (d.n_slots > h->slots_max) h->slots_max = d.n_slots;
  }
  h->Qtot = qoff;
  h->bit_words = (qoff + 31) / 32;
  if ((rc = upload(qnode.data(), qnode.size(), &h->qubit_node))) return bail(rc);
  rc = configure_paths(h);
  if (rc != FPB_OK && rc != FPB_ERR_UNSUPPORTED) return bail(rc);
  // FPB_ERR_UNSUPPORTED: the handle still serves the noise / syndrome stages
  *out = h;
  return FPB_OK;
}

int fpb_destroy(fpb_handle* h) {
  if (!h) return FPB_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->adj_indptr); cudaFree(h->adj_indices); cudaFree(h->adj_vals); cudaFree(h->qubit_node);
  for (auto& cx : h->cx) {
    cudaFree(cx.cube_qubit); cudaFree(cx.ninfo); cudaFree(cx.sbase); cudaFree(cx.slot_qubit);
    cudaFree(cx.low_cube); cudaFree(cx.low_slot); cudaFree(cx.high_cube); cudaFree(cx.high_slot);
    cx.parity.release(); cx.odd_count.release(); cx.odd_fixed.release(); cx.word_rank.release(); cx.odd_ids.release();
    cx.odd_off.release(); cx.pair_off.release(); cx.pair_len.release(); cx.bnd_len.release();
    cx.bnd_side.release(); cx.knn.release(); cx.knn_len.release(); cx.tab_pair.release(); cx.tab_bnd.release(); cx.tab_side.release();
  }
  h->x.release(); h->hom.release(); h->weights.release(); h->state.release(); h->bits.release();
  h->flips.release(); h->item_off.release(); h->counter.release(); h->totals.release();
  h->path_qubits.release(); h->path_len.release();
  cudaFree(h->macro_partner); cudaFree(h->macro_sign);
  for (int c = 0; c < FPB_MAX_COMPLEX; ++c) { cudaFree(h->uf_edge_a[c]); cudaFree(h->uf_edge_b[c]); }
  cudaFree(h->plane_off); cudaFree(h->plane_qubits); h->fail.release();
  h->m_state.release(); h->m_body.release(); h->m_red_state.release(); h->m_bit_val.release(); h->m_mean.release();
  h->m_z.release(); h->m_val.release(); h->m_p_phase.release(); h->m_outcome.release();
  h->trial_wmax.release(); h->price_n.release(); h->price_y.release(); h->price_c.release(); h->price_ws.release(); h->price_top.release();
  h->pair_out.release(); h->price_active.release(); h->price_idx.release();
  h->q_trial.release(); h->q_cx.release(); h->q_src.release(); h->q_dst.release(); h->wscratch.release();
  if (h->totals_host) cudaFreeHost(h->totals_host);
  if (h->static_flag_host) cudaFreeHost(h->static_flag_host);
  h->static_w.release(); h->static_flag.release();
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->mark)
    if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return FPB_OK;
}

int fpb_run_injected(fpb_handle* h, int64_t n_trials, const double* x, const uint8_t* state, int on_device,
                     int stage) {
  if (!h || !x || !state) return fail(FPB_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  int rc = ensure_batch(h, n_trials);
  if (rc) return rc;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(cudaMemcpyAsync(h->x.p, x, (size_t)n_trials * 2 * h->N * sizeof(double), kind, h->stream));
  CK(cudaMemcpyAsync(h->state.p, state, (size_t)n_trials * h->N, kind, h->stream));
  h->have_rng = false;
  return run_pipeline(h, n_trials, stage, SRC_INJECTED, 0, 0);
}

int fpb_run_rng(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, int stage) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  if (first_trial < 0) return fail(FPB_ERR_ARG, "first_trial must be >= 0");
  CK(cudaSetDevice(h->device));
  int rc = ensure_batch(h, n_trials);
  if (rc) return rc;
  h->have_rng = true;
  return run_pipeline(h, n_trials, stage, SRC_GEN, seed, first_trial);
}

static int check_bits_run(fpb_handle* h) {
  if (h->weight_method != FPB_WEIGHT_UNIFORM)
    return fail(FPB_ERR_ARG, "bits-only trials carry no homodyne outcomes: the handle must use uniform weights");
  return FPB_OK;
}

int fpb_run_bits(fpb_handle* h, int64_t n_trials, const uint32_t* bits, int on_device, int stage) {
  if (!h || !bits) return fail(FPB_ERR_ARG, "null argument");
  int rc = check_bits_run(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  if ((rc = ensure_batch(h, n_trials))) return rc;
  CK(cudaMemcpyAsync(h->bits.p, bits, (size_t)n_trials * h->bit_words * sizeof(uint32_t),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  h->have_rng = false;
  return run_pipeline(h, n_trials, stage, SRC_BITS, 0, 0);
}

int fpb_run_weighted(fpb_handle* h, int64_t n_trials, const double* weights, const uint32_t* bits, int on_device,
                     int stage) {
  if (!h || !weights || !bits) return fail(FPB_ERR_ARG, "null argument");
  if (stage < FPB_STAGE_SYNDROME) return fail(FPB_ERR_ARG, "a weights + bits run starts at the syndrome stage");
  if (h->use_table) return fail(FPB_ERR_STATE, "the static-weight table is in use: given weights would be ignored");
  CK(cudaSetDevice(h->device));
  int rc = ensure_batch(h, n_trials);
  if (rc) return rc;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(cudaMemcpyAsync(h->weights.p, weights, (size_t)n_trials * h->Qtot * sizeof(double), kind, h->stream));
  CK(cudaMemcpyAsync(h->bits.p, bits, (size_t)n_trials * h->bit_words * sizeof(uint32_t), kind, h->stream));
  h->have_rng = false;
  return run_pipeline(h, n_trials, stage, SRC_WEIGHTED, 0, 0);
}

static int ensure_macro(fpb_handle* h, long long T) {
  if (!h->macro_partner) return fail(FPB_ERR_STATE, "no macronode tables: call fpb_set_macro first");
  if (h->use_table) return fail(FPB_ERR_STATE, "the static-weight table is in use: macronode weights are trial-dependent");
  int rc;
  const size_t M = (size_t)T * 4 * h->N, Nn = (size_t)T * h->N;
  if ((rc = h->m_state.ensure(M))) return rc;
  if ((rc = h->m_mean.ensure(M))) return rc;
  if ((rc = h->m_z.ensure(M))) return rc;
  if ((rc = h->m_body.ensure(M))) return rc;
  if ((rc = h->m_val.ensure(M))) return rc;
  if ((rc = h->m_red_state.ensure(Nn))) return rc;
  if ((rc = h->m_bit_val.ensure(Nn))) return rc;
  if ((rc = h->m_p_phase.ensure(Nn))) return rc;
  if ((rc = h->m_outcome.ensure(Nn))) return rc;
  return FPB_OK;
}

int fpb_set_macro(fpb_handle* h, const int32_t* partner, const int8_t* sign) {
  if (!h || !partner || !sign) return fail(FPB_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  const int M = 4 * h->N;
  for (int i = 0; i < M; ++i) {
    if (partner[i] < -1 || partner[i] >= M) return fail(FPB_ERR_ARG, "macronode partner out of range");
    if (partner[i] >= 0 && (partner[partner[i]] != i || partner[i] / 4 == i / 4))
      return fail(FPB_ERR_ARG, "macronode partners must pair micronodes of different macronodes");
  }
  if (h->macro_partner) cudaFree(h->macro_partner);
  if (h->macro_sign) cudaFree(h->macro_sign);
  h->macro_partner = nullptr; h->macro_sign = nullptr;
  int rc;
  if ((rc = upload(partner, (size_t)M, &h->macro_partner))) return rc;
  if ((rc = upload(sign, (size_t)M, &h->macro_sign))) return rc;
  return FPB_OK;
}

int fpb_run_macro_injected(fpb_handle* h, int64_t n_trials, const uint8_t* state, const float* mean, const double* z,
                           int on_device, int stage) {
  if (!h || !state || !mean || !z) return fail(FPB_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  int rc = ensure_batch(h, n_trials);
  if (rc) return rc;
  if ((rc = ensure_macro(h, n_trials))) return rc;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const size_t M = (size_t)n_trials * 4 * h->N;
  CK(cudaMemcpyAsync(h->m_state.p, state, M, kind, h->stream));
  CK(cudaMemcpyAsync(h->m_mean.p, mean, M * sizeof(float), kind, h->stream));
  CK(cudaMemcpyAsync(h->m_z.p, z, M * sizeof(double), kind, h->stream));
  h->have_rng = false;
  return run_pipeline(h, n_trials, stage, SRC_MACRO, 0, 0);
}

int fpb_run_macro_rng(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, int stage) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  if (first_trial < 0) return fail(FPB_ERR_ARG, "first_trial must be >= 0");
  CK(cudaSetDevice(h->device));
  int rc = ensure_batch(h, n_trials);
  if (rc) return rc;
  if ((rc = ensure_macro(h, n_trials))) return rc;
  h->have_rng = true;
  return run_pipeline(h, n_trials, stage, SRC_MACRO_GEN, seed, first_trial);
}

int fpb_fetch_macro(fpb_handle* h, uint8_t* red_state, uint8_t* bit_val, double* p_phase_cond, double* outcome,
                    int on_device) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  if (h->T <= 0 || !h->last_macro) return fail(FPB_ERR_STATE, "the last run was not a macronode run");
  CK(cudaSetDevice(h->device));
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const size_t n = (size_t)h->T * h->N;
  if (red_state) CK(cudaMemcpyAsync(red_state, h->m_red_state.p, n, kind, h->stream));
  if (bit_val) CK(cudaMemcpyAsync(bit_val, h->m_bit_val.p, n, kind, h->stream));
  if (p_phase_cond) CK(cudaMemcpyAsync(p_phase_cond, h->m_p_phase.p, n * sizeof(double), kind, h->stream));
  if (outcome) CK(cudaMemcpyAsync(outcome, h->m_outcome.p, n * sizeof(double), kind, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return FPB_OK;
}

int fpb_run_iid(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, double p_err, int stage) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  if (first_trial < 0) return fail(FPB_ERR_ARG, "first_trial must be >= 0");
  if (!(p_err >= 0.0 && p_err <= 1.0)) return fail(FPB_ERR_ARG, "Probability is not between 0 and 1.");
  int rc = check_bits_run(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  if ((rc = ensure_batch(h, n_trials))) return rc;
  h->have_rng = false;
  return run_pipeline(h, n_trials, stage, SRC_IID, seed, first_trial, p_err);
}

int fpb_get_sizes(fpb_handle* h, fpb_sizes* out) {
  if (!h || !out) return fail(FPB_ERR_ARG, "null argument");
  memset(out, 0, sizeof *out);
  out->n_trials = h->T;
  out->n_complex = h->ncx;
  out->total_qubits = h->Qtot;
  for (int c = 0; c < h->ncx; ++c) {
    out->total_odd[c] = h->cx[c].total_odd;
    out->total_pairs[c] = h->last_stage >= FPB_STAGE_GRAPH ? h->cx[c].total_pairs : 0;
  }
  return FPB_OK;
}

int fpb_device_outputs(fpb_handle* h, fpb_outputs* out) {
  if (!h || !out) return fail(FPB_ERR_ARG, "null argument");
  if (h->T <= 0) return fail(FPB_ERR_STATE, "no run yet");
  memset(out, 0, sizeof *out);
  out->hom = h->hom.p; out->bits = h->bits.p; out->weights = h->weights.p;
  out->state = h->state.p; out->x = h->x.p;
  for (int c = 0; c < h->ncx; ++c) {
    ComplexHost& cx = h->cx[c];
    if (h->last_stage >= FPB_STAGE_SYNDROME) {
      out->parity[c] = cx.parity.p; out->odd_count[c] = cx.odd_count.p; out->odd_ids[c] = cx.odd_ids.p;
    }
    if (h->last_stage >= FPB_STAGE_GRAPH) {
      out->pair_len[c] = cx.pair_len.p;
      if (cx.n_low > 0) { out->bnd_len[c] = cx.bnd_len.p; out->bnd_side[c] = cx.bnd_side.p; }
    }
  }
  return FPB_OK;
}

int fpb_fetch(fpb_handle* h, const fpb_outputs* dst, int on_device) {
  if (!h || !dst) return fail(FPB_ERR_ARG, "null argument");
  if (h->T <= 0) return fail(FPB_ERR_STATE, "no run yet");
  CK(cudaSetDevice(h->device));
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  cudaStream_t st = h->stream;
  const size_t T = (size_t)h->T;
#define COPY(dstp, srcp, count, type)                                                             \
  if ((dstp) && (count) > 0) CK(cudaMemcpyAsync((dstp), (srcp), (size_t)(count) * sizeof(type), kind, st))
  COPY(dst->hom, h->hom.p, T * h->Qtot, double);
  COPY(dst->bits, h->bits.p, T * h->bit_words, uint32_t);
  COPY(dst->weights, h->weights.p, T * h->Qtot, double);
  COPY(dst->state, h->state.p, T * h->N, uint8_t);
  COPY(dst->x, h->x.p, T * 2 * h->N, double);
  for (int c = 0; c < h->ncx; ++c) {
    ComplexHost& cx = h->cx[c];
    const bool syn = h->last_stage >= FPB_STAGE_SYNDROME, gr = h->last_stage >= FPB_STAGE_GRAPH;
    if ((dst->parity[c] || dst->odd_count[c] || dst->odd_ids[c]) && !syn)
      return fail(FPB_ERR_STATE, "syndrome outputs requested but the last run stopped before that stage");
    if ((dst->pair_len[c] || dst->bnd_len[c] || dst->bnd_side[c]) && !gr)
      return fail(FPB_ERR_STATE, "graph outputs requested but the last run stopped before that stage");
    COPY(dst->parity[c], cx.parity.p, T * cx.par_words, uint32_t);
    COPY(dst->odd_count[c], cx.odd_count.p, T, int32_t);
    COPY(dst->odd_ids[c], cx.odd_ids.p, cx.total_odd, int32_t);
    COPY(dst->pair_len[c], cx.pair_len.p, cx.total_pairs, double);
    if (cx.n_low > 0) {
      COPY(dst->bnd_len[c], cx.bnd_len.p, cx.total_odd, double);
      COPY(dst->bnd_side[c], cx.bnd_side.p, cx.total_odd, uint8_t);
    }
  }
#undef COPY
  CK(cudaStreamSynchronize(st));
  return FPB_OK;
}

int fpb_get_timings(fpb_handle* h, fpb_timings* out) {
  if (!h || !out) return fail(FPB_ERR_ARG, "null argument");
  *out = h->tm;
  return FPB_OK;
}

int fpb_synchronize(fpb_handle* h) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return FPB_OK;
}

int fpb_build_table(fpb_handle* h) {
  if (!h) return fail(FPB_ERR_ARG, "null handle");
  const bool uniform = h->weight_method == FPB_WEIGHT_UNIFORM;
  if (h->length_mode != FPB_LENGTH_FLOAT && !uniform && !h->weight_integer)
    return fail(FPB_ERR_UNSUPPORTED, "the static table holds float path sums; int-truncated lengths differ from them unless the weights are integers");
  if (!uniform && !(h->weight_method == FPB_WEIGHT_BLUEPRINT && h->p_swap >= 1.0))
    return fail(FPB_ERR_UNSUPPORTED, "weights are trial-dependent: the static table needs uniform weights or p_swap == 1");
  if (h->warps == 0) return fail(FPB_ERR_UNSUPPORTED, "shortest-path kernel not configured for this lattice");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // trial-independent weights (decoders/decoder.py:86-93 with p_count = degree, :115-116)
  std::vector<double> w((size_t)h->Qtot, 1.0);
  if (!uniform)
    for (int q = 0; q < h->Qtot; ++q) {
      const int deg = h->qubit_degree[q];
      if (deg < 2 || deg > 4) return fail(FPB_ERR_UNSUPPORTED, "a syndrome qubit with fewer than 2 neighbours has hom-dependent weight");
      const double err = deg == 2 ? 0.25 : (deg == 3 ? 1.0 / 3.0 : 0.4);
      w[q] = h->weight_integer ? std::rint(-h->weight_multiplier * std::log(err)) : -std::log(err);
    }
  int rc;
  if ((rc = ensure_batch(h, 1))) return rc;
  if ((rc = h->static_w.ensure((size_t)h->Qtot))) return rc;
  if ((rc = h->static_flag.ensure(1))) return rc;
  CK(cudaMemcpyAsync(h->static_w.p, w.data(), (size_t)h->Qtot * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->weights.p, h->static_w.p, (size_t)h->Qtot * sizeof(double), cudaMemcpyDeviceToDevice, st));
  // pseudo-trial in which every cube is odd: its matching graph is the all-pairs table
  h->use_table = false;
  for (int c = 0; c < h->ncx; ++c) k_iota_odd<<<64, 256, 0, st>>>(make_dev(h->cx[c]));
  int launches = 0;
  h->T = 1;
  h->last_stage = FPB_STAGE_GRAPH;
  if ((rc = post_syndrome(h, 1, FPB_STAGE_GRAPH, &launches))) return rc;
  if ((rc = launch_paths(h, 1, &launches))) return rc;
  for (int c = 0; c < h->ncx; ++c) {
    ComplexHost& cx = h->cx[c];
    if ((rc = cx.tab_pair.ensure((size_t)cx.total_pairs + 1))) return rc;
    if ((rc = cx.tab_bnd.ensure((size_t)cx.S + 1))) return rc;
    if ((rc = cx.tab_side.ensure((size_t)cx.S + 1))) return rc;
    CK(cudaMemcpyAsync(cx.tab_pair.p, cx.pair_len.p, (size_t)cx.total_pairs * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (cx.n_low > 0) {
      CK(cudaMemcpyAsync(cx.tab_bnd.p, cx.bnd_len.p, (size_t)cx.S * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(cx.tab_side.p, cx.bnd_side.p, (size_t)cx.S, cudaMemcpyDeviceToDevice, st));
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  h->use_table = true;
  h->T = 0;  // the pseudo-trial is not a result
  h->last_stage = 0;
  return FPB_OK;
}

int fpb_mark(fpb_handle* h, int which) {
  if (!h || which < 0 || which > 1) return fail(FPB_ERR_ARG, "bad mark");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->mark[which], h->stream));
  return FPB_OK;
}

int fpb_elapsed_ms(fpb_handle* h, float* ms) {
  if (!h || !ms) return fail(FPB_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->mark[1]));
  CK(cudaEventElapsedTime(ms, h->mark[0], h->mark[1]));
  return FPB_OK;
}

static int run_knn(fpb_handle* h, int complex_index, int k, int32_t* ids, double* lens, double* wmax, int on_device) {
  if (!h || !ids) return fail(FPB_ERR_ARG, "null argument");
  if (complex_index < 0 || complex_index >= h->ncx) return fail(FPB_ERR_ARG, "bad complex index");
  if (k < 1 || k > 64) return fail(FPB_ERR_ARG, "k must be in 1..64");
  if (h->last_stage < FPB_STAGE_GRAPH) return fail(FPB_ERR_STATE, "fpb_knn needs a completed graph-stage run");
  CK(cudaSetDevice(h->device));
  ComplexHost& cx = h->cx[complex_index];
  cudaStream_t st = h->stream;
  int rc;
  if (wmax) {
    if ((rc = h->trial_wmax.ensure((size_t)h->T))) return rc;
    CK(cudaMemsetAsync(h->trial_wmax.p, 0, (size_t)h->T * sizeof(unsigned long long), st));
  }
  if (cx.total_odd > 0) {
    if ((rc = cx.knn.ensure((size_t)cx.total_odd * k))) return rc;
    if (lens && (rc = cx.knn_len.ensure((size_t)cx.total_odd * k))) return rc;
    // one warp per odd cube; its row of cx.max_odd doubles lives in shared memory
    const int row_cap = cx.max_odd > 0 ? cx.max_odd : 1;
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    int warps = (int)(((size_t)max_optin - 1024) / ((size_t)row_cap * sizeof(double)));
    if (warps > 8) warps = 8;
    if (warps < 1) return fail(FPB_ERR_UNSUPPORTED, "too many odd cubes in one trial for the candidate kernel");
    const size_t smem = (size_t)warps * row_cap * sizeof(double);
    CK(cudaFuncSetAttribute(k_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024));
    dim3 grid((unsigned)h->T, (unsigned)((cx.max_odd + warps - 1) / warps));
    k_knn<<<grid, warps * 32, smem, st>>>(make_dev(cx), k, row_cap, cx.knn.p, lens ? cx.knn_len.p : nullptr,
                                          wmax ? h->trial_wmax.p : nullptr);
    CK(cudaGetLastError());
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpyAsync(ids, cx.knn.p, (size_t)cx.total_odd * k * sizeof(int32_t), kind, st));
    if (lens) CK(cudaMemcpyAsync(lens, cx.knn_len.p, (size_t)cx.total_odd * k * sizeof(double), kind, st));
  }
  if (wmax)  // the bits of a non-negative double, written by atomicMax
    CK(cudaMemcpyAsync(wmax, h->trial_wmax.p, (size_t)h->T * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return FPB_OK;
}

int fpb_knn(fpb_handle* h, int complex_index, int k, int32_t* ids, int on_device) {
  return run_knn(h, complex_index, k, ids, nullptr, nullptr, on_device);
}

int fpb_knn_len(fpb_handle* h, int complex_index, int k, int32_t* ids, double* lens, double* wmax) {
  if (!lens || !wmax) return fail(FPB_ERR_ARG, "null argument");
  return run_knn(h, complex_index, k, ids, lens, wmax, 0);
}

int fpb_pairs_at(fpb_handle* h, int complex_index, int64_t n, const int32_t* trial, const int32_t* a, const int32_t* b,
                 double* out) {
  if (!h || (n > 0 && (!trial || !a || !b || !out))) return fail(FPB_ERR_ARG, "null argument");
  if (complex_index < 0 || complex_index >= h->ncx) return fail(FPB_ERR_ARG, "bad complex index");
  if (h->last_stage < FPB_STAGE_GRAPH) return fail(FPB_ERR_STATE, "fpb_pairs_at needs a completed graph-stage run");
  if (n <= 0) return FPB_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  ComplexHost& cx = h->cx[complex_index];
  // validate on the host against the odd counts of the last run
  std::vector<int> counts((size_t)h->T);
  CK(cudaMemcpyAsync(counts.data(), cx.odd_count.p, (size_t)h->T * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; ++i) {
    if (trial[i] < 0 || trial[i] >= h->T) return fail(FPB_ERR_ARG, "pair trial out of range");
    const int K = counts[trial[i]];
    if (a[i] < 0 || a[i] >= K || b[i] < 0 || b[i] >= K) return fail(FPB_ERR_ARG, "pair endpoint out of range");
  }
  int rc;
  if ((rc = h->q_trial.ensure((size_t)n))) return rc;
  if ((rc = h->q_src.ensure((size_t)n))) return rc;
  if ((rc = h->q_dst.ensure((size_t)n))) return rc;
  if ((rc = h->pair_out.ensure((size_t)n))) return rc;
  CK(cudaMemcpyAsync(h->q_trial.p, trial, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->q_src.p, a, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->q_dst.p, b, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  const long long blocks = (n + 255) / 256, cap = (long long)h->sm_count * 8;
  k_pairs_at<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(make_dev(cx), n, h->q_trial.p, h->q_src.p,
                                                                     h->q_dst.p, h->pair_out.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->pair_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return FPB_OK;
}

int fpb_pairs_of_trial(fpb_handle* h, int complex_index, int64_t trial, double* out, int64_t cap) {
  if (!h || !out) return fail(FPB_ERR_ARG, "null argument");
  if (complex_index < 0 || complex_index >= h->ncx) return fail(FPB_ERR_ARG, "bad complex index");
  if (h->last_stage < FPB_STAGE_GRAPH) return fail(FPB_ERR_STATE, "fpb_pairs_of_trial needs a completed graph-stage run");
  if (trial < 0 || trial >= h->T) return fail(FPB_ERR_ARG, "trial out of range");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  ComplexHost& cx = h->cx[complex_index];
  long long off[2];
  CK(cudaMemcpyAsync(off, cx.pair_off.p + trial, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const long long n = off[1] - off[0];
  if (n > cap) return fail(FPB_ERR_ARG, "output buffer smaller than K(K-1)/2");
  if (n > 0) CK(cudaMemcpyAsync(out, cx.pair_len.p + off[0], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return FPB_OK;
}

int fpb_price(fpb_handle* h, int complex_index, const int64_t* y, const int32_t* top, const int64_t* ztop,
              const double* scale, const int32_t* wshift, const uint8_t* active, int64_t cap, int32_t* out_trial, int32_t* out_a, int32_t* out_b, double* out_len,
              int64_t* n_out) {
  if (!h || !y || !scale || !wshift || !active || !n_out) return fail(FPB_ERR_ARG, "null argument");
  if (cap > 0 && (!out_trial || !out_a || !out_b || !out_len)) return fail(FPB_ERR_ARG, "null output array");
  if (complex_index < 0 || complex_index >= h->ncx) return fail(FPB_ERR_ARG, "bad complex index");
  if (h->last_stage < FPB_STAGE_GRAPH) return fail(FPB_ERR_STATE, "fpb_price needs a completed graph-stage run");
  *n_out = 0;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  ComplexHost& cx = h->cx[complex_index];
  if (cx.total_odd == 0) return FPB_OK;
  if (cap < 0) cap = 0;
  int rc;
  const size_t T = (size_t)h->T;
  if ((top != nullptr) != (ztop != nullptr)) return fail(FPB_ERR_ARG, "top and ztop come together");
  if ((rc = h->price_y.ensure(2 * (size_t)cx.total_odd))) return rc;
  if ((rc = h->price_top.ensure((size_t)cx.total_odd))) return rc;
  if ((rc = h->price_c.ensure(T))) return rc;
  if ((rc = h->price_ws.ensure(T))) return rc;
  if ((rc = h->price_active.ensure(T))) return rc;
  if ((rc = h->price_idx.ensure(3 * (size_t)cap + 1))) return rc;
  if ((rc = h->pair_out.ensure((size_t)cap + 1))) return rc;
  if ((rc = h->price_n.ensure(1))) return rc;
  CK(cudaMemcpyAsync(h->price_y.p, y, (size_t)cx.total_odd * sizeof(long long), cudaMemcpyHostToDevice, st));
  if (top) {
    CK(cudaMemcpyAsync(h->price_y.p + cx.total_odd, ztop, (size_t)cx.total_odd * sizeof(long long), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->price_top.p, top, (size_t)cx.total_odd * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(h->price_c.p, scale, T * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->price_ws.p, wshift, T * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->price_active.p, active, T, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(h->price_n.p, 0, sizeof(unsigned long long), st));
  PriceParams pp;
  pp.top = top ? h->price_top.p : nullptr; pp.ztop = h->price_y.p + cx.total_odd;
  pp.y = h->price_y.p; pp.scale = h->price_c.p; pp.wshift = h->price_ws.p; pp.active = h->price_active.p;
  pp.cap = cap;
  pp.out_trial = h->price_idx.p; pp.out_a = h->price_idx.p + cap; pp.out_b = h->price_idx.p + 2 * cap;
  pp.out_len = h->pair_out.p; pp.n_out = h->price_n.p;
  int rows = (cx.max_odd + PRICE_ROWS - 1) / PRICE_ROWS;
  if (rows < 1) rows = 1;
  if (rows > 64) rows = 64;
  k_price<<<dim3((unsigned)T, (unsigned)rows), PRICE_ROWS * 32, 0, st>>>(make_dev(cx), pp);
  CK(cudaGetLastError());
  unsigned long long found = 0;
  CK(cudaMemcpyAsync(&found, h->price_n.p, sizeof found, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *n_out = (int64_t)found;
  const size_t take = (size_t)(found < (unsigned long long)cap ? found : (unsigned long long)cap);
  if (take > 0) {
    CK(cudaMemcpyAsync(out_trial, pp.out_trial, take * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_a, pp.out_a, take * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_b, pp.out_b, take * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_len, pp.out_len, take * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return FPB_OK;
}

// fpb_paths_toggle (flips_host) and fpb_paths_list (qubits_host / length_host) share the query machinery
static int run_path_queries(fpb_handle* h, int64_t n_queries, const int32_t* trial, const int32_t* cx_id,
                            const int32_t* src, const int32_t* dst, uint32_t* flips_host, int max_len,
                            int32_t* qubits_host, int32_t* length_host) {
  if (h->last_stage < FPB_STAGE_GRAPH) return fail(FPB_ERR_STATE, "path queries need a completed graph-stage run");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t nflip = (size_t)h->T * h->bit_words;
  if (flips_host) {
    if ((rc = h->flips.ensure(nflip))) return rc;
    CK(cudaMemsetAsync(h->flips.p, 0, nflip * sizeof(uint32_t), st));
  }
  if (n_queries > 0) {
    if (!trial || !cx_id || !src || !dst) return fail(FPB_ERR_ARG, "null query array");
    // validate on the host: indices must address odd cubes of the last run
    std::vector<int> counts[FPB_MAX_COMPLEX];
    for (int c = 0; c < h->ncx; ++c) {
      counts[c].resize((size_t)h->T);
      CK(cudaMemcpyAsync(counts[c].data(), h->cx[c].odd_count.p, (size_t)h->T * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_queries; ++i) {
      if (trial[i] < 0 || trial[i] >= h->T || cx_id[i] < 0 || cx_id[i] >= h->ncx) return fail(FPB_ERR_ARG, "query trial / complex out of range");
      const int K = counts[cx_id[i]][trial[i]];
      if (src[i] < 0 || src[i] >= K || dst[i] >= K || dst[i] < -1 || dst[i] == src[i]) return fail(FPB_ERR_ARG, "query endpoint out of range");
      if (dst[i] == -1 && h->cx[cx_id[i]].n_low == 0) return fail(FPB_ERR_ARG, "connector query on a complex without boundary points");
    }
    if ((rc = h->q_trial.ensure((size_t)n_queries))) return rc;
    if ((rc = h->q_cx.ensure((size_t)n_queries))) return rc;
    if ((rc = h->q_src.ensure((size_t)n_queries))) return rc;
    if ((rc = h->q_dst.ensure((size_t)n_queries))) return rc;
    CK(cudaMemcpyAsync(h->q_trial.p, trial, (size_t)n_queries * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->q_cx.p, cx_id, (size_t)n_queries * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->q_src.p, src, (size_t)n_queries * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->q_dst.p, dst, (size_t)n_queries * sizeof(int), cudaMemcpyHostToDevice, st));
    PathsParams pp;
    memset(&pp, 0, sizeof pp);
    pp.n_complex = h->ncx; pp.T = (int)h->T; pp.Qtot = h->Qtot;
    pp.S_max = h->S_max; pp.slots_max = h->slots_max;
    pp.step_factor = h->step_factor;
    pp.weights = h->weights.p;
    for (int c = 0; c < h->ncx; ++c) pp.cx[c] = make_dev(h->cx[c]);
    pp.n_queries = (int)n_queries;
    pp.q_trial = h->q_trial.p; pp.q_cx = h->q_cx.p; pp.q_src = h->q_src.p; pp.q_dst = h->q_dst.p;
    pp.flips = h->flips.p; pp.bit_words = h->bit_words;
    if (qubits_host) {
      if ((rc = h->path_qubits.ensure((size_t)n_queries * max_len))) return rc;
      if ((rc = h->path_len.ensure((size_t)n_queries))) return rc;
      pp.path_qubits = h->path_qubits.p; pp.path_len = h->path_len.p; pp.path_cap = max_len;
    }
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const bool wg = h->paths_wg;
    const size_t per_warp = wg ? paths_warp_bytes(h->S_max, true) : toggle_warp_bytes(h->S_max, h->slots_max);
    int warps = (int)((size_t)max_optin / per_warp);
    if (warps < 1) return fail(FPB_ERR_UNSUPPORTED, "lattice too large for the path-toggle kernel");
    if (warps > 8) warps = 8;
    const size_t smem = (size_t)warps * per_warp;
    toggle_fn tfn = pick_toggle_kernel(h->S_max, wg);
    {
      cudaFuncAttributes fattr;
      CK(cudaFuncGetAttributes(&fattr, tfn));
      CK(cudaFuncSetAttribute(tfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)((size_t)max_optin - fattr.sharedSizeBytes)));
    }
    long long grid = (n_queries + warps - 1) / warps;
    const long long cap = (long long)h->sm_count * 4;
    if (grid > cap) grid = cap;
    if (wg) {
      if ((rc = h->wscratch.ensure((size_t)grid * warps * h->slots_max))) return rc;
      pp.wscratch = h->wscratch.p;
    }
    tfn<<<(unsigned)grid, warps * 32, smem, st>>>(pp);
    CK(cudaGetLastError());
    if (qubits_host) {
      CK(cudaMemcpyAsync(qubits_host, h->path_qubits.p, (size_t)n_queries * max_len * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(length_host, h->path_len.p, (size_t)n_queries * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
  }
  if (flips_host) CK(cudaMemcpyAsync(flips_host, h->flips.p, nflip * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (flips_host) h->flips_T = h->T;  // the recovery stays on the device for fpb_logical_check
  return FPB_OK;
}

int fpb_paths_toggle(fpb_handle* h, int64_t n_queries, const int32_t* trial, const int32_t* cx_id,
                     const int32_t* src, const int32_t* dst, uint32_t* flips_host) {
  if (!h || !flips_host) return fail(FPB_ERR_ARG, "null argument");
  return run_path_queries(h, n_queries, trial, cx_id, src, dst, flips_host, 0, nullptr, nullptr);
}

int fpb_paths_list(fpb_handle* h, int64_t n_queries, const int32_t* trial, const int32_t* cx_id, const int32_t* src,
                   const int32_t* dst, int32_t max_len, int32_t* qubits, int32_t* length) {
  if (!h || !qubits || !length) return fail(FPB_ERR_ARG, "null argument");
  if (max_len < 1) return fail(FPB_ERR_ARG, "max_len must be positive");
  return run_path_queries(h, n_queries, trial, cx_id, src, dst, nullptr, max_len, qubits, length);
}

int fpb_uf_set_graph(fpb_handle* h, int complex_index, int32_t n_nodes, const int32_t* edge_a, const int32_t* edge_b) {
  if (!h || !edge_a || !edge_b) return fail(FPB_ERR_ARG, "null argument");
  if (complex_index < 0 || complex_index >= h->ncx) return fail(FPB_ERR_ARG, "complex index out of range");
  const ComplexHost& cx = h->cx[complex_index];
  if (n_nodes < cx.S || n_nodes > 65000) return fail(FPB_ERR_ARG, "n_nodes must be in [n_cubes, 65000]");
  for (int q = 0; q < cx.Q; ++q) {
    if (edge_a[q] < 0) continue;
    if (edge_a[q] >= cx.S || edge_b[q] < 0 || edge_b[q] >= n_nodes || edge_b[q] == edge_a[q])
      return fail(FPB_ERR_ARG, "edge end out of range (edge_a: a cube, edge_b: another cube or a boundary point)");
  }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(h->uf_edge_a[complex_index]);
  cudaFree(h->uf_edge_b[complex_index]);
  h->uf_edge_a[complex_index] = h->uf_edge_b[complex_index] = nullptr;
  h->uf_nodes[complex_index] = 0;
  int rc;
  if ((rc = upload(edge_a, (size_t)cx.Q, &h->uf_edge_a[complex_index]))) return rc;
  if ((rc = upload(edge_b, (size_t)cx.Q, &h->uf_edge_b[complex_index]))) return rc;
  h->uf_nodes[complex_index] = n_nodes;
  return FPB_OK;
}

int fpb_uf_decode(fpb_handle* h, int use_erasures, uint32_t* flips_host, uint32_t* grown_host, int32_t* n_stuck) {
  if (!h || !n_stuck) return fail(FPB_ERR_ARG, "null argument");
  if (h->T <= 0 || h->last_stage < FPB_STAGE_SYNDROME) return fail(FPB_ERR_STATE, "the Union-Find decoder needs a completed syndrome-stage run");
  if (use_erasures && !h->last_labels) return fail(FPB_ERR_STATE, "the last run carried no state labels: no erasures to take");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  UfParams up;
  memset(&up, 0, sizeof up);
  up.n_complex = h->ncx; up.N = h->N; up.bit_words = h->bit_words;
  size_t smem = 0;
  for (int c = 0; c < h->ncx; ++c) {
    if (!h->uf_edge_a[c]) return fail(FPB_ERR_STATE, "no decoder graph: call fpb_uf_set_graph for every complex first");
    up.cx[c] = make_dev(h->cx[c]);
    up.edge_a[c] = h->uf_edge_a[c];
    up.edge_b[c] = h->uf_edge_b[c];
    up.n_nodes[c] = h->uf_nodes[c];
    const size_t b = uf_smem_bytes(h->uf_nodes[c], h->cx[c].S, h->cx[c].Q);
    if (b > smem) smem = b;
  }
  int max_optin = 0;
  CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  cudaFuncAttributes fattr;
  CK(cudaFuncGetAttributes(&fattr, k_uf));
  const size_t room = (size_t)max_optin - fattr.sharedSizeBytes;
  if (smem > room) return fail(FPB_ERR_UNSUPPORTED, "lattice too large for the shared-memory Union-Find kernel");
  CK(cudaFuncSetAttribute(k_uf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room));
  up.state = use_erasures ? (h->last_macro ? h->m_red_state.p : h->state.p) : nullptr;
  up.adj_indptr = h->adj_indptr; up.adj_indices = h->adj_indices; up.qubit_node = h->qubit_node;
  int rc;
  const size_t nflip = (size_t)h->T * h->bit_words;
  if ((rc = h->flips.ensure(nflip))) return rc;
  if ((rc = h->counter.ensure(1))) return rc;
  CK(cudaMemsetAsync(h->flips.p, 0, nflip * sizeof(uint32_t), st));
  CK(cudaMemsetAsync(h->counter.p, 0, sizeof(int), st));
  up.flips = h->flips.p;
  up.stuck = h->counter.p;
  if (grown_host) {  // second half of the flips buffer
    if ((rc = h->flips.ensure(2 * nflip))) return rc;
    up.flips = h->flips.p;
    up.grown = h->flips.p + nflip;
    CK(cudaMemsetAsync(h->flips.p, 0, 2 * nflip * sizeof(uint32_t), st));
  }
  NvtxRange nvtx_uf("fpb:union_find");
  k_uf<<<dim3((unsigned)h->T, (unsigned)h->ncx), 256, smem, st>>>(up);
  CK(cudaGetLastError());
  if (flips_host) CK(cudaMemcpyAsync(flips_host, h->flips.p, nflip * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (grown_host) CK(cudaMemcpyAsync(grown_host, h->flips.p + nflip, nflip * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  int stuck = 0;
  CK(cudaMemcpyAsync(&stuck, h->counter.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *n_stuck = stuck;
  h->flips_T = h->T;
  return FPB_OK;
}

int fpb_set_planes(fpb_handle* h, int32_t n_planes, const int32_t* plane_off, const int32_t* plane_qubits) {
  if (!h || !plane_off || (n_planes > 0 && !plane_qubits)) return fail(FPB_ERR_ARG, "null argument");
  if (n_planes < 0 || plane_off[0] != 0) return fail(FPB_ERR_ARG, "plane offsets must start at 0");
  for (int i = 0; i < n_planes; ++i)
    if (plane_off[i + 1] < plane_off[i]) return fail(FPB_ERR_ARG, "plane offsets must not decrease");
  const int total = plane_off[n_planes];
  for (int i = 0; i < total; ++i)
    if (plane_qubits[i] < 0 || plane_qubits[i] >= h->Qtot) return fail(FPB_ERR_ARG, "plane qubit out of range");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(h->plane_off);
  cudaFree(h->plane_qubits);
  h->plane_off = h->plane_qubits = nullptr;
  h->n_planes = 0;
  int rc;
  if ((rc = upload(plane_off, (size_t)n_planes + 1, &h->plane_off))) return rc;
  if ((rc = upload(plane_qubits, (size_t)total, &h->plane_qubits))) return rc;
  h->n_planes = n_planes;
  return FPB_OK;
}

int fpb_logical_check(fpb_handle* h, uint8_t* fail_host, int64_t* n_fail) {
  if (!h || !n_fail) return fail(FPB_ERR_ARG, "null argument");
  if (!h->plane_off) return fail(FPB_ERR_STATE, "no correlation surfaces: call fpb_set_planes first");
  if (h->T <= 0 || h->flips_T != h->T) return fail(FPB_ERR_STATE, "no recovery on the device for the last run (fpb_uf_decode / fpb_paths_toggle)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  if ((rc = h->fail.ensure((size_t)h->T))) return rc;
  if ((rc = h->counter.ensure(1))) return rc;
  CK(cudaMemsetAsync(h->counter.p, 0, sizeof(int), st));
  CheckParams cp;
  cp.T = (int)h->T; cp.bit_words = h->bit_words; cp.n_planes = h->n_planes;
  cp.plane_off = h->plane_off; cp.plane_qubits = h->plane_qubits;
  cp.bits = h->bits.p; cp.flips = h->flips.p; cp.fail = h->fail.p; cp.n_fail = h->counter.p;
  long long blocks = (h->T + 7) / 8;
  const long long cap = (long long)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_logical_check<<<(unsigned)blocks, 256, 0, st>>>(cp);
  CK(cudaGetLastError());
  if (fail_host) CK(cudaMemcpyAsync(fail_host, h->fail.p, (size_t)h->T, cudaMemcpyDeviceToHost, st));
  int nf = 0;
  CK(cudaMemcpyAsync(&nf, h->counter.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *n_fail = nf;
  return FPB_OK;
}

}  // extern "C"
