/*
 * fpb.h - C ABI of the B200 Monte-Carlo error-correction trial backend.
 *
 * The reference (XanaduAI/flamingpy) has no C ABI for this path: its plug-in
 * surface is Python (`CVLayer.apply_noise`, `assign_weights`,
 * `MatchingGraph.with_edges_from`) plus one pybind11 pattern
 * (`flamingpy/cpp/lemonpy.cpp:50-84`: C-contiguous arrays in, freshly owned
 * arrays out, errors as exceptions).  This header is the boundary a maintainer
 * binds instead (INTEGRATION.md shows the ctypes / pybind11 stubs).  Each entry
 * point names the reference interface it replaces; paths are relative to the
 * reference root.
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, every function
 * returns 0 on success or a negative error code with the message available
 * from fpb_last_error().  A handle owns one CUDA device, one stream and its
 * device buffers; it is thread-compatible (one thread at a time per handle).
 * There is no CPU fallback: fpb_create fails if no CUDA device is usable.
 */
#ifndef FPB_H
#define FPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPB_ABI_VERSION 2
#define FPB_MAX_COMPLEX 2

/* error codes */
#define FPB_OK 0
#define FPB_ERR_ARG (-1)
#define FPB_ERR_CUDA (-2)
#define FPB_ERR_NOMEM (-3)
#define FPB_ERR_STATE (-4)
#define FPB_ERR_UNSUPPORTED (-5)

/* weight methods: assign_weights(method=...) flamingpy/decoders/decoder.py:58-116 */
#define FPB_WEIGHT_UNIFORM 0
#define FPB_WEIGHT_BLUEPRINT 1

/* label modes for the device RNG (SURVEY.md finding 0.3) */
#define FPB_MODE_COMPAT 0 /* one reused CVLayer per chain: GKP_t = V \ (p_t U GKP_{t-1}) */
#define FPB_MODE_FRESH 1  /* a new CVLayer per trial: every non-p node is a noisy GKP state */

/* node states in `state` arrays */
#define FPB_STATE_SILENT 0
#define FPB_STATE_GKP 1
#define FPB_STATE_P 2

/* what a path "length" is (fpb_config.length_mode) */
#define FPB_LENGTH_FLOAT 0    /* sum of the float weights along the shortest path: the networkx backend
                                 (codes/graphs/stabilizer_graph.py:405-410) */
#define FPB_LENGTH_RX_TRUNC 1 /* sum of int(weight) along the float-shortest path: what the reference's default
                                 rustworkx backend reports (stabilizer_graph.py:486-500, _path_weight) */

/* stages to run */
#define FPB_STAGE_NOISE 1    /* homodyne values, bits, weights */
#define FPB_STAGE_SYNDROME 2 /* + stabilizer parities and odd-cube lists */
#define FPB_STAGE_GRAPH 3    /* + matching-graph lengths */

/* One error complex = one stabilizer graph: a 3-D box of cubes in the order of
 * `SurfaceCode.{ec}_stabilizers` (flamingpy/codes/surface_code.py:396-428), each
 * with six face slots (-x,+x,-y,+y,-z,+z).  Replaces StabilizerGraph.__init__ /
 * connect_nodes (codes/graphs/stabilizer_graph.py:72-88,289-302). */
typedef struct fpb_complex_desc {
  int32_t n_cubes;           /* S */
  int32_t n_qubits;          /* Q: syndrome qubits of this complex */
  const int32_t* qubit_node; /* [Q] lattice node index of each syndrome qubit, ascending */
  const int32_t* cube_qubit; /* [S*6] local qubit id on each face, -1 if absent */
  const uint16_t* ninfo;     /* [S] six 2-bit face kinds: 0 regular, 1 periodic wrap, 2 none, 3 LOW/HIGH */
  const int32_t* sbase;      /* [S] weight-slot base of each cube */
  int32_t d_off[12];         /* [6][2] cube-index offset per direction for kind 0 / 1 */
  int32_t s_off[12];         /* [6][2] weight-slot offset per direction for kind 0 / 1 */
  int32_t n_slots;
  const int32_t* slot_qubit; /* [n_slots] local qubit id held by each weight slot, -1 unused */
  int32_t n_low;             /* cubes adjacent to the LOW connector (0 without boundary points) */
  const int32_t* low_cube;   /* [n_low] */
  const int32_t* low_slot;   /* [n_low] weight slot of the boundary qubit */
  int32_t n_high;
  const int32_t* high_cube;
  const int32_t* high_slot;
} fpb_complex_desc;

/* Static description of SurfaceCode + CVLayer + weight options.
 * Replaces SurfaceCode.__init__ tables (codes/surface_code.py:309-371),
 * CVLayer.__init__ (noise/cv.py:70-105) and the weight_options dict of
 * correct() (decoders/decoder.py:227-273). */
typedef struct fpb_config {
  int32_t abi_version; /* FPB_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal */
  int32_t n_nodes;     /* N */
  const int32_t* adj_indptr;  /* [N+1] CSR of EGraph.adj_generator (codes/graphs/egraph.py:157-178) */
  const int32_t* adj_indices; /* ascending within a row */
  const int8_t* adj_vals;     /* edge weight (+-1 polarity) */
  int32_t n_complex;          /* 1, or 2 for ec="both" */
  fpb_complex_desc complex_[FPB_MAX_COMPLEX];
  double delta;         /* CVLayer(delta=...) */
  double p_swap;        /* CVLayer(p_swap=...); 0 = none */
  int32_t weight_method; /* FPB_WEIGHT_* */
  double weight_delta;  /* weight_options["delta"] */
  int32_t weight_integer;   /* weight_options["integer"] */
  double weight_multiplier; /* weight_options["multiplier"] */
  int32_t label_mode;   /* FPB_MODE_* (device RNG only) */
  int64_t chain_len;    /* trials per reused-layer chain in compat mode; <=0: one chain */
  double delta_step;    /* shortest-path bucket width in units of the trial's mean weight; <=0: default (2.0) */
  int32_t length_mode;  /* FPB_LENGTH_* */
  int32_t reserve_sms;  /* SMs the persistent shortest-path kernel leaves free, so that short kernels of other
                           handles (fpb_price, fpb_knn_len of the previous batch) need not wait for its end */
  int32_t reserved_i[6]; /* must be 0 */
  double reserved_d[4];  /* must be 0 */
} fpb_config;

typedef struct fpb_handle fpb_handle;

/* Result sizes of the last run (per complex c). */
typedef struct fpb_sizes {
  int64_t n_trials;
  int32_t n_complex;
  int64_t total_qubits;            /* sum of Q over complexes = row length of hom / weights */
  int64_t total_odd[FPB_MAX_COMPLEX];   /* sum over trials of K */
  int64_t total_pairs[FPB_MAX_COMPLEX]; /* sum over trials of K(K-1)/2 */
} fpb_sizes;

/* Host (or device) destinations for fpb_fetch; any pointer may be NULL to skip it.
 * Row-major, trial-major.  Qubit columns are complex 0's qubits then complex 1's. */
typedef struct fpb_outputs {
  double* hom;        /* [T][Qtot]  CVLayer.measure_hom hom_val_p (noise/cv.py:161-217) */
  uint32_t* bits;     /* [T][ceil(Qtot/32)] bit-packed bit_val, LSB first (noise/cv.py:147-158) */
  double* weights;    /* [T][Qtot]  node attribute "weight" (decoders/decoder.py:58-116) */
  uint8_t* state;     /* [T][N]     FPB_STATE_* labels used (noise/cv.py:121-140) */
  double* x;          /* [T][2N]    raw quadrature draws (noise/cv.py:205) */
  uint32_t* parity[FPB_MAX_COMPLEX];  /* [T][ceil(S/32)] Stabilizer.parity (codes/stabilizer.py:43-55) */
  int32_t* odd_count[FPB_MAX_COMPLEX]; /* [T] K */
  int32_t* odd_ids[FPB_MAX_COMPLEX];   /* [total_odd] odd cubes, stabilizer order, trials concatenated */
  double* pair_len[FPB_MAX_COMPLEX];   /* [total_pairs] per trial the packed upper triangle (a<b, row-major)
                                          of shortest-path lengths (decoders/mwpm/matching.py:125-142) */
  double* bnd_len[FPB_MAX_COMPLEX];    /* [total_odd] length to the nearer connector (matching.py:144-173) */
  uint8_t* bnd_side[FPB_MAX_COMPLEX];  /* [total_odd] 0 low, 1 high */
} fpb_outputs;

/* Per-stage device times of the last run, CUDA events on the handle's stream. */
typedef struct fpb_timings {
  float gen_ms;      /* Philox label + quadrature generation (device RNG runs only) */
  float noise_ms;    /* entangle + bin + weights */
  float syndrome_ms; /* parity + odd-list compaction + offsets */
  float graph_ms;    /* shortest paths + matching-graph assembly */
  float total_ms;
  int32_t launches;  /* kernels launched by the last run */
  int32_t graph_ctas;
  int32_t graph_warps_per_cta;
  int32_t graph_smem_bytes;
} fpb_timings;

const char* fpb_last_error(void);
int fpb_abi_version(void);
/* number of visible CUDA devices, or a negative error code */
int fpb_device_count(void);

/* Build a handle: uploads the static tables.  Replaces constructing
 * SurfaceCode + CVLayer (codes/surface_code.py:309, noise/cv.py:70). */
int fpb_create(const fpb_config* cfg, fpb_handle** out);
int fpb_destroy(fpb_handle* h);

/* Injected-noise trials: `x` [T][2N] are the draws rng.normal(0, sigma) would
 * return (q block then p block) and `state` [T][N] the FPB_STATE_* labels.
 * Replaces CVLayer.apply_noise after sampling + assign_weights + MatchingGraph
 * construction (noise/cv.py:205-217,147-158; decoders/decoder.py:31-119;
 * decoders/mwpm/matching.py:97-173).  `on_device` != 0: pointers are device
 * pointers on the handle's device.  `stage` = FPB_STAGE_*. */
int fpb_run_injected(fpb_handle* h, int64_t n_trials, const double* x, const uint8_t* state,
                     int on_device, int stage);

/* Device-RNG trials [first_trial, first_trial + n_trials): Philox4x32-10 keyed by
 * `seed` with counter (trial, node, slot), so results do not depend on batching
 * or on the number of GPUs.  Replaces the body of ec_mc_trial up to the
 * matching graph (flamingpy/simulations.py:46-59). */
int fpb_run_rng(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, int stage);

/* Trials of the i.i.d. Z-error model (flamingpy/noise/iid.py:19-50: bit_val = rng.random() < p
 * per node), which has no homodyne outcomes: the handle must have been created with
 * weight_method = FPB_WEIGHT_UNIFORM (decoders/decoder.py:115-116); `hom` reads back as zeros and
 * `weights` as ones.  fpb_run_bits takes the bits of the syndrome qubits, [n_trials][bit_words]
 * uint32 words, LSB first, in the qubit order of fpb_outputs.bits; fpb_run_iid draws them on the
 * device from the Philox stream (seed; trial, node), independent of batching and GPU count.
 * With static weights every later graph stage can come from fpb_build_table's gather. */
int fpb_run_bits(fpb_handle* h, int64_t n_trials, const uint32_t* bits, int on_device, int stage);
int fpb_run_iid(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, double p_err, int stage);

/* Trials given as the two node attributes the reference's matching-graph builder reads
 * (decoders/mwpm/matching.py:97-173 through codes/graphs/stabilizer_graph.py:304-329): the
 * `weight` of every syndrome qubit, [n_trials][Qtot] doubles, and its `bit_val`,
 * [n_trials][bit_words] packed like fpb_outputs.bits.  Runs parity -> odd cubes -> matching graph
 * (stage >= FPB_STAGE_SYNDROME) whatever noise model and weight assignment produced them;
 * `hom` reads back as zeros. */
int fpb_run_weighted(fpb_handle* h, int64_t n_trials, const double* weights, const uint32_t* bits, int on_device,
                     int stage);

/* Macronode noise layer, CVMacroLayer (flamingpy/noise/cv.py:352-625; the reference CLI's default noise,
 * simulations.py:224).  Every lattice node is a macronode of four micronodes (micronode 4 i + j of node i),
 * one per incident edge plus partner-less padding (codes/graphs/egraph.py:26-103).
 *
 * fpb_set_macro: the macronode lattice - partner [4N] = micronode at the other end of each edge 'dumbbell'
 *   (-1: padding), sign [4N] = that edge's weight (+-1 polarity).  Replaces EGraph.macronize + the
 *   adjacency matrix of CVMacroLayer.__init__ (cv.py:367-375).
 * fpb_run_macro_injected: trials on caller-supplied draws - state [T][4N] micronode labels (FPB_STATE_*),
 *   mean [T][4N] float32 initial q means of the two-step sampling (cv.py:323-349), z [T][4N] standard normals
 *   of the N star p-measurements then the 3N planet q-measurements (cv.py:474-489) - through CZ gates,
 *   four-splitters (cv/ops.py:104-127), homodyne measurement and the reduction of every macronode to the
 *   canonical lattice (cv.py:548-625), then the weights of assign_weights(prob_precomputed=True)
 *   (decoders/decoder.py:58-93; uniform weights if the handle says so), parities and matching graph.
 *   fpb_outputs.hom holds the effective homodyne value of each syndrome qubit.
 * fpb_run_macro_rng: the same with labels, means and normals from the Philox stream of (seed; trial, micronode).
 * fpb_fetch_macro: per lattice node [T][N] the reduced state label (FPB_STATE_GKP / FPB_STATE_P), bit_val,
 *   p_phase_cond and effective homodyne value of the last macronode run (any pointer may be NULL). */
int fpb_set_macro(fpb_handle* h, const int32_t* partner, const int8_t* sign);
int fpb_run_macro_injected(fpb_handle* h, int64_t n_trials, const uint8_t* state, const float* mean, const double* z,
                           int on_device, int stage);
int fpb_run_macro_rng(fpb_handle* h, uint64_t seed, int64_t first_trial, int64_t n_trials, int stage);
int fpb_fetch_macro(fpb_handle* h, uint8_t* red_state, uint8_t* bit_val, double* p_phase_cond, double* outcome,
                    int on_device);

int fpb_get_sizes(fpb_handle* h, fpb_sizes* out);
/* Copy results of the last run into caller buffers (host if on_device == 0). */
int fpb_fetch(fpb_handle* h, const fpb_outputs* dst, int on_device);
/* Raw device pointers of the last run's result buffers (valid until the next run). */
int fpb_device_outputs(fpb_handle* h, fpb_outputs* out);
int fpb_get_timings(fpb_handle* h, fpb_timings* out);
int fpb_synchronize(fpb_handle* h);
/* User timing marks on the handle's stream (the stream every kernel of the handle is
 * launched on): fpb_mark records CUDA event `which` (0 or 1); fpb_elapsed_ms synchronises
 * and returns the device time between mark 0 and mark 1. */
int fpb_mark(fpb_handle* h, int which);
int fpb_elapsed_ms(fpb_handle* h, float* ms);

/* Static-weight fast path.  With uniform weights, or blueprint weights and p_swap == 1 (every
 * node p-squeezed), assign_weights gives trial-independent weights (decoders/decoder.py:86-93,
 * 115-116), so the shortest-path lengths between all cubes are computed once here and every later
 * graph-stage run gathers its matching graph from that table instead of searching (results are
 * bit-identical to the search path).  Fails with FPB_ERR_UNSUPPORTED for other weight options;
 * a later injected run whose labels are not all "p" fails with FPB_ERR_STATE. */
int fpb_build_table(fpb_handle* h);

/* Candidate partners for the host matcher, from the last graph-stage run: for every odd cube a of
 * every trial of complex `complex_index`, the k nearest other odd cubes of the same trial in the
 * matching graph (shortest-path length, or both boundary lengths if that is shorter - the pair
 * weight of the reduced complete graph that stands for decoders/mwpm/matching.py:125-173), as
 * indices into the trial's odd list, nearest first, ties by index, -1 padded.  ids: int32
 * [total_odd][k] in the ragged order of odd_ids.  A matcher that starts from this sparse graph and
 * prices the remaining pairs (fpb_mwpm.h) saves its own pass over the K(K-1)/2 lengths. */
int fpb_knn(fpb_handle* h, int complex_index, int k, int32_t* ids, int on_device);

/* The sparse hand-off to a host matcher: the K(K-1)/2 lengths stay in HBM and only what the matcher
 * touches crosses the bus (1.5 MB per d=13 trial otherwise).  All three calls refer to the last
 * graph-stage run and to the reduced pair weight w_ab = min(d_ab, b_a + b_b) (without boundary points:
 * d_ab) that stands for decoders/mwpm/matching.py:125-173 in csrc/mwpm.cpp.
 *
 * fpb_knn_len: fpb_knn plus the reduced weights of the candidates, lens [total_odd][k] (inf padded), and
 *   per trial the largest reduced weight / twice the largest boundary length, wmax [n_trials] (the
 *   matcher's fixed-point scale).
 * fpb_pairs_at: the shortest-path lengths d_ab of n requested pairs (trial, a, b), a, b indices into the
 *   trial's odd list.
 * fpb_price: pricing of the complete graph under the matcher's duals, in the matcher's integer
 *   arithmetic (fpb_mwpm.h).  y [total_odd] are its vertex duals (ragged like odd_ids), scale / wshift
 *   [n_trials] the fixed-point scale (a power of two) and the number of doublings of each trial's
 *   weights, active [n_trials] selects the trials; every pair with
 *   y_a + y_b + ((4 * (int64)(w_ab * scale + 0.5)) << wshift) < 0 is appended to (out_trial, out_a,
 *   out_b, out_len = w_ab), at most `cap` entries; *n_out receives the number found (> cap: call again
 *   with more room).  Blossom duals only add slack, so no violated edge is missed; optionally
 *   top / ztop [total_odd] (both or neither) name each vertex's outermost blossom (-1: none) and
 *   twice its dual, which is added for two vertices of the same blossom - a lower bound of their
 *   true blossom term that spares the host most of the pairs it would clear anyway.
 * fpb_pairs_of_trial: all K(K-1)/2 lengths of one trial (the packed triangle of fpb_outputs.pair_len),
 *   for the rare graph a sparse solver gives up on. */
int fpb_knn_len(fpb_handle* h, int complex_index, int k, int32_t* ids, double* lens, double* wmax);
int fpb_pairs_at(fpb_handle* h, int complex_index, int64_t n, const int32_t* trial, const int32_t* a, const int32_t* b,
                 double* out);
int fpb_pairs_of_trial(fpb_handle* h, int complex_index, int64_t trial, double* out, int64_t cap);
int fpb_price(fpb_handle* h, int complex_index, const int64_t* y, const int32_t* top, const int64_t* ztop,
              const double* scale, const int32_t* wshift, const uint8_t* active, int64_t cap, int32_t* out_trial, int32_t* out_a, int32_t* out_b, double* out_len,
              int64_t* n_out);

/* Correction paths for matched pairs of the last run.  Query i asks, in trial
 * trial[i] and complex cx[i], for the shortest path from odd cube src[i] to
 * odd cube dst[i] (dst >= 0), or to its nearer connector (dst == -1).  The local
 * qubit ids crossed are XOR-toggled into flips [T][ceil(Qtot/32)] (bit-packed,
 * same layout as bits).  Replaces edge_path + the common_vertex loop of
 * mwpm_decoder (decoders/mwpm/algos.py:88-95). */
int fpb_paths_toggle(fpb_handle* h, int64_t n_queries, const int32_t* trial, const int32_t* cx,
                     const int32_t* src, const int32_t* dst, uint32_t* flips_host);

/* The same queries as fpb_paths_toggle, answered as ordered paths: for query i the local qubit
 * ids (qubit order of the query's complex) crossed by the shortest path, walking from dst back to
 * src (dst >= 0), or from src to its nearer connector (dst == -1; the last entry is the boundary
 * qubit), into qubits[i * max_len ...], and the number of qubits into length[i] (-1 if the path has
 * more than max_len qubits).  Together with the cube-qubit incidence this is
 * MatchingGraph.edge_path (decoders/mwpm/matching.py:71-74; consumed by algos.py:88-95). */
int fpb_paths_list(fpb_handle* h, int64_t n_queries, const int32_t* trial, const int32_t* cx, const int32_t* src,
                   const int32_t* dst, int32_t max_len, int32_t* qubits, int32_t* length);

/* Union-Find decoder on the device (csrc/fpb_uf.cuh), the alternative to the host decoder of
 * include/fpb_uf.h: same cluster growth as uf_decode / union_find / peeling
 * (flamingpy/decoders/unionfind/algos.py:58-334; Support.grow, uf_classes.py:134-142), one CTA per
 * (trial, complex) on the syndromes the last run left in HBM.
 *
 * fpb_uf_set_graph: the decoder graph of one complex, once per handle - nodes [0, n_cubes) are the
 * cubes, [n_cubes, n_nodes) one boundary point per boundary qubit; edge q (a syndrome qubit, local id)
 * joins edge_a[q] (a cube, or -1: the qubit has no edge, stabilizer_graph.py:295-298) and edge_b[q].
 * fpb_uf_decode: needs a run with stage >= FPB_STAGE_SYNDROME.  use_erasures != 0 starts from the
 * erased edges of assign_weights(code, "UF") (two or more p-squeezed neighbours,
 * decoders/decoder.py:96-114), taken from the state labels of the last run (the reduced labels after a
 * macronode run).  Writes the recovery as flips [T][ceil(Qtot/32)] (bit-packed, layout of bits; NULL:
 * it stays on the device for fpb_logical_check),
 * optionally (grown_host != NULL, same layout) the fully grown edges when the growth stops - the
 * reference's Support table, which does not depend on visiting order - and the number of
 * (trial, complex) items left with an odd cluster that can neither grow nor reach a boundary point
 * (their flips stay zero). */
int fpb_uf_set_graph(fpb_handle* h, int complex_index, int32_t n_nodes, const int32_t* edge_a, const int32_t* edge_b);
int fpb_uf_decode(fpb_handle* h, int use_erasures, uint32_t* flips_host, uint32_t* grown_host, int32_t* n_stuck);

/* Logical verdict on the device: check_correction (flamingpy/decoders/decoder.py:151-224) for every
 * trial of the last run, on the bit values in HBM XOR the recovery the last fpb_uf_decode or
 * fpb_paths_toggle call left there (flips_host of fpb_uf_decode may be NULL: the recovery then never
 * crosses the bus).  fpb_set_planes, once per handle: the checked correlation surfaces as index lists
 * into the global qubit order (plane i = plane_qubits[plane_off[i] .. plane_off[i+1])) - for the
 * reference's surface codes the syndrome qubits of plane 0 along x [open primal] / y [open dual] /
 * x, y [toric] / x, y, z [periodic] of every complex (decoder.py:197-214).  fpb_logical_check:
 * fail_host[t] = 1 where some plane has odd parity (NULL: only the count), *n_fail = their number. */
int fpb_set_planes(fpb_handle* h, int32_t n_planes, const int32_t* plane_off, const int32_t* plane_qubits);
int fpb_logical_check(fpb_handle* h, uint8_t* fail_host, int64_t* n_fail);

#ifdef __cplusplus
}
#endif
#endif /* FPB_H */
