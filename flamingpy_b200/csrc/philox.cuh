// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the
// transforms used by the device-RNG trial path.  Counter = (trial lo, trial hi,
// node, slot), key = run seed, so any (trial, node) draw can be recomputed
// statelessly - that is what lets the carry-over recurrence of a reused
// CVLayer (SURVEY.md finding 0.3) be evaluated by looking back over earlier
// trials' label draws without storing any state.
#pragma once
#include <cstdint>

namespace fpb {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  Philox4 out = {c0, c1, c2, c3};
  return out;
}

// slot ids of the per-(trial, node) streams
constexpr uint32_t SLOT_QUADRATURE = 0;  // 128 bits -> one Box-Muller pair (q, p)
constexpr uint32_t SLOT_LABEL = 1;       // 32 bits  -> Bernoulli(p_swap)
constexpr uint32_t SLOT_IID = 2;         // 53 bits  -> Bernoulli(p_err) of the i.i.d. noise model
// macronode layer: the "node" of the counter is the micronode index
constexpr uint32_t SLOT_MACRO_LABEL = 3;  // 32 bits  -> Bernoulli(p_swap) per micronode
constexpr uint32_t SLOT_MACRO_DRAW = 4;   // 53 bits -> initial mean (uniform or bit), 64 bits -> measurement normal

// uniform in (0, 1) from 53 random bits (never 0, never 1)
__host__ __device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace fpb
