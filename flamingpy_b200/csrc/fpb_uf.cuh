// Union-Find + peeling decoder on the device: one CTA per (trial, complex), all state in shared memory.
//
// Computes what flamingpy/decoders/unionfind/algos.py does (Delfosse-Nickerson growth with erasures,
// then peeling), in the synchronous formulation of csrc/uf.cpp (the host decoder, which is pinned edge
// for edge to the reference's Support tables):
//
//   1. erased edges (two or more p-squeezed neighbours of the qubit, decoders/decoder.py:96-114) start
//      fully grown; clusters = connected components of the grown edges (algos.py:58-129);
//   2. per round every edge gains one half per end that lies in a growing cluster - odd parity, and from
//      the second round on no boundary point (uf_classes.py:92-98,134-142, algos.py:132-193) - then the
//      newly grown edges merge clusters.  The grown set after each round is independent of any visiting
//      order, so it equals the reference's and the host decoder's;
//   3. a breadth-first spanning forest of the grown edges, rooted at the lowest-numbered boundary point of
//      each component (else its lowest cube), is built one level per pass from a node-centric frontier and
//      peeled level by level from the leaves (algos.py:196-311).
//      Among the edges that can attach a node to the previous level the lowest-numbered one is its tree
//      edge - a different forest than the host decoder's FIFO order, so the two recoveries may differ by
//      cycles of grown edges; both clear the syndrome inside the same grown set.
//
// Clusters are kept as a parent forest over node ids with min-id hooking (atomicMin on roots, repeated
// until an edge pass finds nothing to hook) and full path compression between passes; cluster parity and
// boundary contact are recomputed per round with shared-memory atomics.  Nodes [0, S) are cubes, [S, n_nodes)
// boundary points (leaves), edges are the complex's syndrome qubits (edge_a < 0: the qubit has no edge).
#pragma once
#include "fpb_kernels.cuh"

namespace fpb {

struct UfParams {
  int n_complex, N, bit_words;
  ComplexDev cx[MAXC];
  const int* edge_a[MAXC];    // [Q] first end (a cube) or -1
  const int* edge_b[MAXC];    // [Q] second end (cube or boundary point)
  int n_nodes[MAXC];
  const uint8_t* state;       // [T][N] labels (2 = p-squeezed) for the erasures, or nullptr: none
  const int* adj_indptr;      // lattice CSR
  const int* adj_indices;
  const int* qubit_node;      // [Qtot] lattice node of every syndrome qubit
  uint32_t* flips;            // [T][bit_words], zeroed before the launch
  uint32_t* grown;            // [T][bit_words] fully grown edges when growth stops (zeroed), or nullptr
  int* stuck;                 // number of (trial, complex) items whose odd cluster could not be resolved
};

__host__ __device__ inline size_t uf_smem_bytes(int n_nodes, int n_cubes, int Q) {
  return align16((size_t)Q) + 3 * align16((size_t)n_nodes * 4) + align16((size_t)n_nodes * 2) +
         align16((size_t)(n_nodes - n_cubes) * 4);
}

__device__ __forceinline__ int uf_find(const volatile int* par, int u) {
  for (;;) {
    const int p = par[u];
    if (p == u) return u;
    u = p;
  }
}

// hook the clusters joined by fully grown edges until no edge crosses two clusters; leaves par[] fully compressed
__device__ __forceinline__ void uf_merge(volatile int* par, const volatile uint8_t* sup, const int* ea, const int* eb, int Q,
                                         int NN, volatile int* flag) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int pass = 0; pass <= NN; ++pass) {  // every pass that hooks lowers a parent, so NN passes always suffice
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int q = tid; q < Q; q += nthr) {
      if (sup[q] < 2) continue;
      const int a = uf_find(par, ea[q]), b = uf_find(par, eb[q]);
      if (a != b) {
        atomicMin((int*)&par[a > b ? a : b], a > b ? b : a);  // only ever lowers a parent: no cycles
        *flag = 1;
      }
    }
    __syncthreads();
    for (int u = tid; u < NN; u += nthr) par[u] = uf_find(par, u);  // roots do not move in this pass
    __syncthreads();
    const int again = *flag;
    __syncthreads();
    if (!again) break;
  }
}

__global__ void __launch_bounds__(256) k_uf(const __grid_constant__ UfParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int flag_s[3];
  const int t = blockIdx.x, ci = blockIdx.y;
  const ComplexDev& c = p.cx[ci];
  const int S = c.S, Q = c.Q, NN = p.n_nodes[ci];
  const int* ea = p.edge_a[ci];
  const int* eb = p.edge_b[ci];
  const int tid = threadIdx.x, nthr = blockDim.x;
  unsigned char* ptr = smem;
  volatile uint8_t* sup = ptr; ptr += align16((size_t)Q);          // 0 empty, 1 half grown, 2 grown, 3 grown + in the recovery
  volatile int* par = (int*)ptr; ptr += align16((size_t)NN * 4);
  volatile int* cl = (int*)ptr; ptr += align16((size_t)NN * 4);   // growth: per root, bit 0 parity, bit 1 boundary; peeling: running parity
  volatile int* upe = (int*)ptr; ptr += align16((size_t)NN * 4);  // peeling: tree edge towards the root
  volatile uint16_t* depth = (uint16_t*)ptr; ptr += align16((size_t)NN * 2);
  int* bedge = (int*)ptr;                                          // the one edge of every boundary point
  const uint32_t* parw = c.parity + (long long)t * c.par_words;
  auto cube_parity = [&](int u) -> int { return (int)((parw[u >> 5] >> (u & 31)) & 1u); };

  for (int u = tid; u < NN; u += nthr) par[u] = u;
  const uint8_t* st = p.state ? p.state + (long long)t * p.N : nullptr;
  for (int q = tid; q < Q; q += nthr) {
    uint8_t s = 0;
    if (st && ea[q] >= 0) {
      const int node = p.qubit_node[c.qoff + q];
      int pc = 0;
      for (int e = p.adj_indptr[node]; e < p.adj_indptr[node + 1]; ++e) pc += st[p.adj_indices[e]] == 2;
      if (pc >= 2) s = 2;
    }
    sup[q] = s;
    if (ea[q] >= 0 && eb[q] >= S) bedge[eb[q] - S] = q;
  }
  __syncthreads();
  if (st) uf_merge(par, sup, ea, eb, Q, NN, &flag_s[0]);

  bool ok = true;
  for (int round = 0;; ++round) {
    if (round > 2 * Q + 2) {  // cannot happen: every round grows at least one half edge
      ok = false;
      break;
    }
    // parity and boundary contact of every cluster, kept at its root
    for (int u = tid; u < NN; u += nthr) cl[u] = 0;
    if (tid == 0) flag_s[0] = flag_s[1] = flag_s[2] = 0;
    __syncthreads();
    for (int u = tid; u < NN; u += nthr) {
      const int r = par[u];
      if (u >= S) atomicOr((int*)&cl[r], 2);
      else if (cube_parity(u)) atomicXor((int*)&cl[r], 1);
    }
    __syncthreads();
    // the first round grows every odd cluster, later ones only those without a boundary point (algos.py:124-126,186-191)
    for (int u = tid; u < NN; u += nthr) {
      const int f = cl[par[u]];
      if (round == 0 ? (f & 1) : (f == 1)) flag_s[1] = 1;
      if (f == 1) flag_s[2] = 1;  // an odd cluster without a boundary point: it has to grow
    }
    for (int q = tid; q < Q; q += nthr) {
      const int a = ea[q];
      if (a < 0) continue;
      const int s = sup[q];
      if (s >= 2) continue;
      const int fa = cl[par[a]], fb = cl[par[eb[q]]];
      const int inc = (round == 0 ? (fa & 1) + (fb & 1) : (fa == 1) + (fb == 1));
      if (inc) {
        sup[q] = (uint8_t)(s + inc > 2 ? 2 : s + inc);
        flag_s[0] = 1;
      }
    }
    __syncthreads();
    const int any_growing = flag_s[1], changed = flag_s[0], must_grow = flag_s[2];
    __syncthreads();
    if (!any_growing) break;
    if (!changed) {
      // nothing left to grow: fine if the odd clusters of the first round all hold a boundary point (every edge
      // around them erased, p_swap = 1); an odd cluster without one can never be resolved
      ok = !must_grow;
      break;
    }
    uf_merge(par, sup, ea, eb, Q, NN, &flag_s[0]);
  }
  if (!ok) {
    if (tid == 0) atomicAdd(p.stuck, 1);
    return;
  }
  if (p.grown) {
    uint32_t* gr = p.grown + (long long)t * p.bit_words;
    for (int q = tid; q < Q; q += nthr)
      if (sup[q] >= 2) {
        const int g = c.qoff + q;
        atomicOr(&gr[g >> 5], 1u << (g & 31));
      }
  }

  // ---- spanning forest: roots at depth 0, then one breadth-first level per pass over the grown edges ----
  for (int u = tid; u < NN; u += nthr) {
    cl[u] = 0x7fffffff;
    upe[u] = 0x7fffffff;
    depth[u] = 0xffff;
  }
  __syncthreads();
  for (int u = S + tid; u < NN; u += nthr) atomicMin((int*)&cl[par[u]], u);  // lowest boundary point of the component
  __syncthreads();
  for (int u = tid; u < NN; u += nthr)
    if (par[u] == u) depth[cl[u] != 0x7fffffff ? cl[u] : u] = 0;
  __syncthreads();
  for (int u = tid; u < NN; u += nthr) cl[u] = u < S ? cube_parity(u) : 0;
  // One level per pass over the nodes: a node of the current level offers each of its grown edges to the unreached
  // node at the other end, which keeps the lowest-numbered offer as its tree edge (cubes have up to six edges,
  // listed by cube_qubit; a boundary point has one).
  int levels = 0;
  for (; levels < NN; ++levels) {
    if (tid == 0) flag_s[2] = 0;
    __syncthreads();
    for (int u = tid; u < NN; u += nthr) {
      if (depth[u] != levels) continue;
      if (u >= S) {
        const int q = bedge[u - S];
        if (sup[q] >= 2 && depth[ea[q]] == 0xffff) {
          atomicMin((int*)&upe[ea[q]], q);
          flag_s[2] = 1;
        }
        continue;
      }
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const int q = c.cube_qubit[u * 6 + f];
        if (q < 0 || sup[q] < 2) continue;
        const int a = ea[q];
        if (a < 0) continue;
        const int v = a == u ? eb[q] : a;
        if (depth[v] == 0xffff) {
          atomicMin((int*)&upe[v], q);
          flag_s[2] = 1;
        }
      }
    }
    __syncthreads();
    const int found = flag_s[2];
    for (int u = tid; u < NN; u += nthr)
      if (depth[u] == 0xffff && upe[u] != 0x7fffffff) depth[u] = (uint16_t)(levels + 1);
    __syncthreads();
    if (!found) break;
  }
  // ---- peeling, deepest level first: an odd node flips its tree edge and hands its parity to its parent ----
  for (int L = levels; L >= 1; --L) {
    for (int u = tid; u < NN; u += nthr) {
      if (depth[u] != L || !(cl[u] & 1)) continue;
      const int q = upe[u];
      sup[q] = 3;
      atomicXor((int*)&cl[ea[q] == u ? eb[q] : ea[q]], 1);
    }
    __syncthreads();
  }
  if (tid == 0) flag_s[0] = 0;
  __syncthreads();
  for (int u = tid; u < S; u += nthr)
    if (depth[u] == 0 && (cl[u] & 1)) flag_s[0] = 1;  // an odd tree without a boundary point
  __syncthreads();
  if (tid == 0 && flag_s[0]) atomicAdd(p.stuck, 1);
  uint32_t* fl = p.flips + (long long)t * p.bit_words;
  for (int q = tid; q < Q; q += nthr)
    if (sup[q] == 3) {
      const int g = c.qoff + q;
      atomicOr(&fl[g >> 5], 1u << (g & 31));
    }
}

// ------------------------------------------------------------------------------------------
// k_logical_check: the logical verdict of every trial after a recovery (check_correction,
// decoders/decoder.py:151-224): a trial fails when any checked correlation surface - a plane of
// syndrome qubits - has odd parity of bit_val XOR recovery.  One warp per trial; the planes are
// index lists into the global qubit order (the layout of bits / flips).
// ------------------------------------------------------------------------------------------
struct CheckParams {
  int T, bit_words, n_planes;
  const int* plane_off;      // [n_planes + 1]
  const int* plane_qubits;   // global qubit ids
  const uint32_t* bits;      // [T][bit_words]
  const uint32_t* flips;     // [T][bit_words]
  uint8_t* fail;             // [T]
  int* n_fail;
};

__global__ void __launch_bounds__(256) k_logical_check(CheckParams p) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < p.T; t += warps) {
    const uint32_t* b = p.bits + (long long)t * p.bit_words;
    const uint32_t* f = p.flips + (long long)t * p.bit_words;
    int failed = 0;
    for (int pl = 0; pl < p.n_planes; ++pl) {
      int par = 0;
      for (int i = p.plane_off[pl] + lane; i < p.plane_off[pl + 1]; i += 32) {
        const int g = p.plane_qubits[i];
        par ^= (int)(((b[g >> 5] ^ f[g >> 5]) >> (g & 31)) & 1u);
      }
      failed |= __popc(__ballot_sync(0xffffffffu, par)) & 1;
    }
    if (lane == 0) {
      p.fail[t] = (uint8_t)failed;
      if (failed) atomicAdd(p.n_fail, 1);
    }
  }
}

}  // namespace fpb
