/* fpb_uf.h - C ABI of the host-side Union-Find + peeling decoder (libfpb_uf.so).
 *
 * Replaces, for a batch of trials, the reference's pure-Python decoder
 *   flamingpy/decoders/unionfind/algos.py:58-334  (initialize_cluster_trees, union_find,
 *                                                  obtain_spanning_forest, peeling, uf_decoder)
 *   flamingpy/decoders/unionfind/uf_classes.py    (Node / Root / Support / Boundary)
 * which `decoder.correct(code, decoder="UF", ...)` reaches through `decoder_dict["UF"]`
 * (decoders/decoder.py:279).  The decoder works on the stabilizer graph of one error complex without
 * its 'low' / 'high' connectors (algos.py:77,157; uf_classes.py:128): nodes are the stabilizer cubes
 * and the boundary points, every syndrome qubit is one edge (its `common_vertex`).
 *
 * The GPU pipeline (include/fpb.h, stage FPB_STAGE_SYNDROME) delivers what it consumes: the parity of
 * every cube and, through the state labels, the erased edges (weight -1 of `assign_weights(code, "UF")`,
 * decoders/decoder.py:96-114: two or more p-squeezed neighbours).
 */
#ifndef FPB_UF_H
#define FPB_UF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Decode n_trials syndromes on one graph.
 *   n_nodes, n_cubes : nodes [0, n_cubes) are stabilizer cubes, [n_cubes, n_nodes) boundary points
 *   n_edges          : syndrome qubits; edge q joins edge_a[q] and edge_b[q] (node ids), or is
 *                      absent from the graph when edge_a[q] < 0 (a qubit shared twice by one pair
 *                      of cubes keeps one edge only, stabilizer_graph.py:295-298)
 *   parity  [T][n_cubes] : 1 = unsatisfied stabilizer (Stabilizer.parity, codes/stabilizer.py:43-55)
 *   erased  [T][n_edges] : 1 = erased edge (starts fully grown), or NULL for none
 *   flips   [T][n_edges] : out, 1 = qubit in the recovery set (algos.py:286-311)
 *   grown   [T][n_edges] : out or NULL, 1 = edge fully grown when the clusters stopped growing
 *                          (the Support table handed to span_forest, uf_classes.py:144-169)
 *   threads              : OpenMP threads over trials, <= 0 = all
 * Returns 0, or -(1 + t) when trial t has an odd cluster that can neither grow nor reach a boundary
 * point (an odd number of unsatisfied stabilizers on a complex without boundary; the reference loops
 * forever there). */
int fpb_uf_batch(int n_nodes, int n_cubes, int n_edges, const int32_t* edge_a, const int32_t* edge_b, int n_trials,
                 const uint8_t* parity, const uint8_t* erased, uint8_t* flips, uint8_t* grown, int threads);

/* Erased edges of a batch from the state labels: erased[t][q] = 1 when syndrome qubit q (lattice node
 * qubit_node[q]) has two or more neighbours labelled p-squeezed (state == 2) - the weight -1 of
 * `assign_weights(code, "UF")`, decoders/decoder.py:96-114.  adj_indptr / adj_indices: the lattice
 * adjacency in CSR form (the same arrays fpb_config carries).  state [T][n_lattice_nodes], erased [T][n_edges].
 * Returns the number of erased edges in the batch. */
long long fpb_uf_erasures(int n_lattice_nodes, int n_edges, const int32_t* qubit_node, const int32_t* adj_indptr,
                          const int32_t* adj_indices, int n_trials, const uint8_t* state, uint8_t* erased, int threads);

/* check_correction (flamingpy/decoders/decoder.py:151-224) for a batch on the host: fail[t] = 1 when
 * some checked correlation surface - plane i = global qubit ids plane_qubits[plane_off[i] ..
 * plane_off[i+1]) - has odd parity of bit value XOR recovery.  Bit values and recovery each either
 * unpacked ([T][n_qubits] bytes; the word pointer NULL) or as the device delivers them
 * ([T][ceil(n_qubits/32)] words, LSB first; the byte pointer NULL).  Returns the number of failures. */
long long fpb_uf_check_correction(int n_trials, int n_qubits, const uint8_t* bits, const uint32_t* bit_words,
                                  const uint8_t* flips, const uint32_t* flip_words, int n_planes,
                                  const int32_t* plane_off, const int32_t* plane_qubits, uint8_t* fail, int threads);

int fpb_uf_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
