// TEST INFRASTRUCTURE - kernels that pin the semantics of the SIMT machine in tests/emu (run by
// tests/test_emulated_kernels.py::test_simt_machine_semantics through selftest_run): warp shuffles and votes,
// __syncthreads with exited threads, named barriers with a predicate OR, shared-memory atomics, dynamic shared
// memory addressed through 32-bit shared-window offsets, saturating float-to-int conversions.
#include <cstdint>
#include <cuda_runtime.h>
#ifdef FPB_EMU  // defined by the stand-in header of tests/emu/include; under a real toolkit this file compiles to nothing

__global__ void k_warp(int* out) {  // one warp
  const int lane = threadIdx.x & 31;
  int v = lane * 3 + 1;
  out[lane] = __shfl_sync(0xffffffffu, v, 5);                       // 16 everywhere
  out[32 + lane] = __shfl_up_sync(0xffffffffu, v, 2);               // lanes 0, 1 keep their own
  out[64 + lane] = __shfl_xor_sync(0xffffffffu, v, 16);
  out[96 + lane] = (int)__ballot_sync(0xffffffffu, lane % 3 == 0);
  out[128 + lane] = __any_sync(0xffffffffu, lane == 31) + 2 * (int)__reduce_or_sync(0xffffffffu, 1u << (lane & 7));
  int inc = 1;                                                       // inclusive scan
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  out[160 + lane] = inc;
}

__global__ void k_block(int* out, int n_exit) {  // 128 threads; the first n_exit leave before the barriers
  __shared__ int acc;
  __shared__ int slot[128];
  if (threadIdx.x == 127) acc = 0;
  if ((int)threadIdx.x < n_exit) return;
  __syncthreads();
  atomicAdd(&acc, (int)threadIdx.x);
  slot[threadIdx.x] = (int)threadIdx.x;
  __syncthreads();
  // every thread reads what another warp wrote before the barrier
  out[threadIdx.x] = acc + slot[127 - (threadIdx.x - n_exit) % (128 - n_exit)];
}

__global__ void k_named(int* out) {  // 128 threads = two teams of 64 with their own barriers and predicate ORs
  extern __shared__ __align__(16) unsigned char smem[];
  double* vals = (double*)smem;
  const int team = threadIdx.x >> 6, r = threadIdx.x & 63;
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(vals + threadIdx.x);
  *fpb_emu::shared_ptr<volatile double>(a) = 0.5 * threadIdx.x;
  const int any = fpb_emu::named_barrier(1 + team, 64, team == 1 && r == 7);
  const double other = vals[(team << 6) + (63 - r)];
  const int none = fpb_emu::named_barrier(1 + team, 64, 0);
  out[threadIdx.x] = (int)(2.0 * other) * 4 + any * 2 + none;
}

__global__ void k_convert(int* out) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  out[0] = __double2int_rd(inf);
  out[1] = __double2int_rd(-inf);
  out[2] = __double2int_rd(-2.5);
  out[3] = __double2int_rz(-2.5);
  out[4] = __double2int_rn(inf - inf);
  out[5] = (int)__funnelshift_r(0x80000001u, 0x80000001u, 1);
  out[6] = __ffs(0x50) + 100 * __popc(0xf0f0u);
}

extern "C" int selftest_run(int* out /* [1024] */) {
  int* d;
  cudaMalloc(&d, 1024 * sizeof(int));
  cudaMemset(d, 0, 1024 * sizeof(int));
  k_warp<<<1, 32, 0, nullptr>>>(d);
  k_block<<<1, 128, 0, nullptr>>>(d + 192, 0);
  k_block<<<1, 128, 0, nullptr>>>(d + 320, 40);
  k_named<<<1, 128, 128 * sizeof(double), nullptr>>>(d + 448);
  k_convert<<<1, 1, 0, nullptr>>>(d + 576);
  cudaMemcpy(out, d, 1024 * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return 0;
}
#endif  // FPB_EMU
