/* fpb_mwpm.h - C ABI of the host-side minimum-weight perfect matcher (libfpb_mwpm.so).
 *
 * Replaces the reference's native matcher plug-in, the LEMON wrapper
 *   flamingpy/cpp/lemonpy.cpp:13-41,50-84  (mwpm(adjacency) -> matched pairs; bound from
 *   flamingpy/decoders/mwpm/lemon.py:25-40), and rx.max_weight_matching
 *   (flamingpy/decoders/mwpm/matching.py:286-292),
 * for the matching graphs libfpb.so builds (fpb.h: fpb_outputs.pair_len / bnd_len).  Plain C,
 * caller-owned buffers, no GPU.  The matching graph of one trial and complex
 * (decoders/mwpm/matching.py:97-173) is given as
 *   K          number of odd cubes,
 *   pair_len   K(K-1)/2 shortest-path lengths, packed upper triangle, row-major (a < b),
 *   bnd_len    K lengths to the nearer boundary connector, or NULL when the code has none.
 * A result is a list of queries (src, dst): dst >= 0 pairs two odd cubes, dst = -1 sends src to
 * its connector - exactly what fpb_paths_toggle (fpb.h) takes.
 */
#ifndef FPB_MWPM_H
#define FPB_MWPM_H

#ifdef __cplusplus
extern "C" {
#endif

/* One graph.  out_src / out_dst need room for K entries.  Returns the number of queries written,
 * or < 0: -1 no perfect matching (odd K without boundary), -2 lengths out of range, -3 internal. */
int fpb_mwpm(int K, const double* pair_len, const double* bnd_len, int* out_src, int* out_dst);

/* T graphs in the ragged layout of fpb_outputs: odd_count[T]; pair_len / bnd_len concatenated over
 * trials; q_off[T + 1] = exclusive prefix sum of odd_count (trial t writes its queries at
 * q_off[t], at most odd_count[t] of them); n_queries[T] receives the counts.  OpenMP over trials,
 * n_threads <= 0 keeps the runtime's default.  Returns 0, or -1 if any trial failed. */
int fpb_mwpm_batch(int T, const int* odd_count, const double* pair_len, const double* bnd_len,
                   const long long* q_off, int* out_src, int* out_dst, int* n_queries, int n_threads);

/* fpb_mwpm_batch with candidate partners supplied by the GPU (fpb_knn in fpb.h): cand is int32
 * [sum of odd_count][cand_k], row q_off[t] + a = the cand_k nearest partners of odd cube a of trial
 * t (indices into that trial's odd list, nearest first, -1 padded); NULL = select on the host.  The
 * candidates only seed the sparse solver; pricing against all K(K-1)/2 lengths keeps the result exact. */
int fpb_mwpm_batch_cand(int T, const int* odd_count, const double* pair_len, const double* bnd_len,
                        const long long* q_off, const int* cand, int cand_k, int* out_src, int* out_dst,
                        int* n_queries, int n_threads);

int fpb_mwpm_max_threads(void);

/* mode 0: sparse candidate graph + pricing from 48 cubes on, dense O(n^3) below (default);
 * 1: dense only; 2: sparse for every K >= 4.  knn > 0: nearest partners per cube in the first
 * candidate graph (default 12).  Both solvers return a minimum-weight perfect matching of the same
 * complete graph; only the choice among ties can differ. */
void fpb_mwpm_configure(int mode, int knn);

/* out[0..7]: sparse solves, retries with a denser candidate graph, pricing rounds, priced-in edges,
 * then nanoseconds (summed over threads) in candidate scan, graph build + jump start, solve stages,
 * pricing.  reset != 0 zeroes the counters. */
void fpb_mwpm_stats(long long* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* FPB_MWPM_H */
