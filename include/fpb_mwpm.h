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

/* Sparse hand-off: the same exact solver for a batch whose K(K-1)/2 lengths stay on the GPU.  The
 * caller feeds each step from libfpb.so (fpb.h):
 *   ctx = fpb_mwpm_sparse_begin(T, odd_count, q_off, cand, cand_len, cand_k, bnd_len, wmax, threads)
 *           cand / cand_len / wmax from fpb_knn_len; bnd_len [sum K] or NULL; q_off as in fpb_mwpm_batch
 *   n   = fpb_mwpm_sparse_requests(ctx, trial, a, b, cap)   pairs whose lengths the jobs need (chain edges
 *           between cubes of neighbouring boundary length; every pair of a graph below 48 cubes); returns
 *           the count, fills the arrays when cap >= count
 *   fpb_mwpm_sparse_start(ctx, req_len)                      req_len [n] from fpb_pairs_at
 *   loop: fpb_mwpm_sparse_solve(ctx, y, top, ztop, scale, wshift, active)  -> number of trials awaiting pricing (0: done);
 *           top / ztop [sum K] (or NULL): outermost blossom of each vertex and twice its dual, for fpb_price
 *         fpb_mwpm_sparse_apply(ctx, n, vt, va, vb, vlen, max_rounds)   the pairs fpb_price flagged; a trial still
 *           violated after max_rounds (<= 0: 40) warm restarts is given up (reported by finish)
 *   fpb_mwpm_sparse_finish(ctx, out_src, out_dst, n_queries, failed)   queries in the layout of fpb_mwpm_batch;
 *           failed [T] marks trials to be solved from their full triangle instead (returns their number)
 *   fpb_mwpm_sparse_free(ctx)
 * Exactness: a trial is finished only when no pair of the complete graph violates dual feasibility
 * (exact integer vertex-dual test on the device, re-check with blossom duals here), as in fpb_mwpm. */
void* fpb_mwpm_sparse_begin(int T, const int* odd_count, const long long* q_off, const int* cand,
                            const double* cand_len, int cand_k, const double* bnd_len, const double* wmax,
                            int n_threads);
long long fpb_mwpm_sparse_requests(void* ctx, int* trial, int* a, int* b, long long cap);
int fpb_mwpm_sparse_start(void* ctx, const double* req_len);
int fpb_mwpm_sparse_solve(void* ctx, long long* y, int* top, long long* ztop, double* scale, int* wshift,
                          unsigned char* active);
int fpb_mwpm_sparse_apply(void* ctx, long long n, const int* vt, const int* va, const int* vb, const double* vlen,
                          int max_rounds);
int fpb_mwpm_sparse_finish(void* ctx, int* out_src, int* out_dst, int* n_queries, unsigned char* failed);
void fpb_mwpm_sparse_free(void* ctx);

int fpb_mwpm_max_threads(void);

/* mode 0: sparse candidate graph + pricing from 48 cubes on, every edge of the complete graph
 * below (default); 1 ("dense"): the complete graph for every K; 2: sparse for every K >= 4.
 * knn > 0: nearest partners per cube in the first candidate graph (default 12).  It is one
 * primal-dual solver in all modes (csrc/mwpm_sparse.h); each mode returns a minimum-weight
 * perfect matching of the same complete graph, only the choice among ties can differ. */
void fpb_mwpm_configure(int mode, int knn);

/* out[0..7]: sparse solves, retries with a denser candidate graph, pricing rounds, priced-in edges,
 * then nanoseconds (summed over threads) in candidate scan, graph build + jump start, solve stages,
 * pricing.  reset != 0 zeroes the counters. */
void fpb_mwpm_stats(long long* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* FPB_MWPM_H */
