// TEST INFRASTRUCTURE - not part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the UNMODIFIED kernel sources of
// flamingpy_b200/csrc (fpb_kernels.cuh, fpb_macro.cuh, fpb.cu) into a host library and run them on
// the CPU with SIMT semantics, so that the parity tests can execute the real kernel code in a
// container without a GPU (tests/emu/build_emu.py builds it, tests/test_emulated_kernels.py uses it).
// Nothing under flamingpy_b200/ loads this library by itself; the product's libfpb.so is always the
// nvcc build, and it fails without a CUDA device.
//
// Execution model: one CTA at a time; every CUDA thread of the CTA is a fiber (own stack, cooperative
// switch); warp collectives (__shfl*_sync, __ballot_sync, __any_sync, __reduce_or_sync, __syncwarp),
// __syncthreads and named barriers (bar.sync / bar.red.or through fpb_emu::named_barrier) block the
// calling fiber until every live participant has arrived.  Fibers of a warp run round-robin, warps
// run round-robin; a pass over the whole CTA without any progress is reported as a deadlock.
// Shared memory: `__shared__` variables become statics, the dynamic window is one static buffer; both
// lie in this library's data segment, so 32-bit "shared addresses" are offsets from one anchor.
// What this does NOT model: timing, bank conflicts, memory-ordering races between lanes of a warp
// inside one unsynchronised stretch (each fiber runs alone until its next collective).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <sys/mman.h>

#define FPB_EMU 1

// ------------------------------------------------------------------ language keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __constant__ static
#define __align__(n) alignas(n)

// ------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

// ------------------------------------------------------------------ the SIMT machine
namespace fpb_emu {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 << 10;
constexpr size_t DYN_SMEM_BYTES = 256 << 10;
constexpr int N_BARRIERS = 16;

struct Warp {
  unsigned long long val[2][32];
  int count[2];
  unsigned part[2];            // participants of the generation released last in each slot
  long long release_gen;       // highest generation released so far
  long long lane_gen[32];      // next generation each lane will enter
  unsigned alive;              // lanes that have not returned
};
struct NamedBarrier {
  int count;
  int pred_or;
  long long gen;               // completed generations
  int result;                  // OR of the last completed generation
};
struct Fiber {
  void* sp;
  bool done;
};

struct Machine {
  Fiber fiber[MAX_THREADS];
  Warp warp[MAX_THREADS / 32];
  NamedBarrier bar[N_BARRIERS];
  void* sched_sp = nullptr;
  int cur = 0, n_threads = 0, live = 0;
  long long progress = 0;
  unsigned char* stacks = nullptr;
  const std::function<void()>* body = nullptr;
  dim3 block_dim, grid_dim;
  std::recursive_mutex lock;   // kernels of all host threads run one at a time
};
extern Machine M;
alignas(128) extern unsigned char dyn_smem_buf[DYN_SMEM_BYTES];
extern unsigned char smem_anchor;

extern "C" void fpb_emu_switch(void** save_sp, void* load_sp);
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void yield();

inline unsigned char* dyn_smem() { return dyn_smem_buf; }
inline uint32_t shared_addr(const void* p) { return (uint32_t)((uintptr_t)p - (uintptr_t)&smem_anchor); }
template <class T>
inline T* shared_ptr(uint32_t a) { return (T*)((uintptr_t)&smem_anchor + (intptr_t)(int32_t)a); }

// one warp-wide rendezvous; returns the slot holding every lane's value
unsigned long long* warp_exchange(unsigned mask, unsigned long long mine, unsigned* participants);
int named_barrier(int id, int n_threads, int pred);  // bar.sync / bar.red.or: OR of the predicates

template <class... P, class... A>
inline void launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, void* /*stream*/, A&&... args) {
  std::function<void()> body = [&]() { kernel(args...); };
  run_grid(grid, block, smem, body);
}

template <class T>
inline unsigned long long to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payloads of up to 8 bytes");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

}  // namespace fpb_emu

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ------------------------------------------------------------------ synchronisation and warp collectives
static inline int fpb_emu_lane() { return (int)(threadIdx.x & 31u); }
static inline void __syncthreads() { fpb_emu::named_barrier(0, -1, 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  unsigned part;
  fpb_emu::warp_exchange(mask, 0, &part);
}
static inline void __threadfence_block() {}
static inline void __threadfence() {}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, fpb_emu::to_bits(v), &part);
  const int lane = fpb_emu_lane();
  const int idx = (lane & ~(width - 1)) | (src & (width - 1));
  return ((part >> idx) & 1u) ? fpb_emu::from_bits<T>(s[idx]) : v;
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, fpb_emu::to_bits(v), &part);
  const int lane = fpb_emu_lane();
  const int idx = lane - (int)delta;
  return (idx >= (lane & ~(width - 1)) && ((part >> idx) & 1u)) ? fpb_emu::from_bits<T>(s[idx]) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, fpb_emu::to_bits(v), &part);
  const int lane = fpb_emu_lane();
  const int idx = lane + (int)delta;
  return (idx < (lane & ~(width - 1)) + width && ((part >> idx) & 1u)) ? fpb_emu::from_bits<T>(s[idx]) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, fpb_emu::to_bits(v), &part);
  const int idx = fpb_emu_lane() ^ lane_mask;
  (void)width;
  return ((part >> idx) & 1u) ? fpb_emu::from_bits<T>(s[idx]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, pred ? 1ull : 0ull, &part);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if (((part >> i) & 1u) && s[i]) r |= 1u << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, v, &part);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if ((part >> i) & 1u) r |= (unsigned)s[i];
  return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  unsigned part;
  const unsigned long long* s = fpb_emu::warp_exchange(mask, v, &part);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if ((part >> i) & 1u) r += (unsigned)s[i];
  return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

// ------------------------------------------------------------------ atomics (fibers never pre-empt each other)
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicXor(T* p, U v) { T o = *p; *p = (T)(o ^ (T)v); return o; }
template <class T, class U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// ------------------------------------------------------------------ arithmetic intrinsics (compile with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
// float -> integer conversions saturate (and map NaN to 0) like the device instructions
static inline int fpb_emu_sat_int(double r) {
  if (r != r) return 0;
  return r >= 2147483647.0 ? 2147483647 : (r <= -2147483648.0 ? (int)(-2147483647 - 1) : (int)r);
}
static inline long long fpb_emu_sat_ll(double r) {
  if (r != r) return 0;
  return r >= 9223372036854775808.0 ? 9223372036854775807ll : (r <= -9223372036854775808.0 ? (-9223372036854775807ll - 1) : (long long)r);
}
static inline int __double2int_rz(double x) { return fpb_emu_sat_int(std::trunc(x)); }
static inline int __double2int_rd(double x) { return fpb_emu_sat_int(std::floor(x)); }
static inline int __double2int_ru(double x) { return fpb_emu_sat_int(std::ceil(x)); }
static inline int __double2int_rn(double x) { return fpb_emu_sat_int(std::nearbyint(x)); }
static inline long long __double2ll_rz(double x) { return fpb_emu_sat_ll(std::trunc(x)); }
static inline long long __double2ll_rd(double x) { return fpb_emu_sat_ll(std::floor(x)); }
static inline unsigned long long __double2ull_rz(double x) {
  if (!(x > 0.0)) return 0;
  return x >= 18446744073709551616.0 ? ~0ull : (unsigned long long)x;
}
static inline double __longlong_as_double(long long v) { return fpb_emu::from_bits<double>((unsigned long long)v); }
static inline long long __double_as_longlong(double v) { return (long long)fpb_emu::to_bits(v); }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  const unsigned long long x = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(x >> (shift & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
  const unsigned long long x = ((unsigned long long)hi << 32) | lo;
  return (unsigned)((x << (shift & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const unsigned long long x = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((x >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return fpb_emu::shared_addr(p); }
static inline void sincospi(double x, double* s, double* c) {
  // exact argument reduction to [-1/4, 1/4] quarter turns, as the device function does
  double r = x - 2.0 * std::floor(x * 0.5 + 0.25) ;  // r in [-0.5, 1.5)
  const double pi = 3.14159265358979323846;
  int q = (int)std::floor(r * 2.0 + 0.5);           // nearest multiple of 1/2
  const double t = (r - 0.5 * q) * pi;
  const double st = std::sin(t), ct = std::cos(t);
  switch (q & 3) {
    case 0: *s = st; *c = ct; break;
    case 1: *s = ct; *c = -st; break;
    case 2: *s = -st; *c = -ct; break;
    default: *s = -ct; *c = st; break;
  }
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::min;
using std::max;
static inline int min(int a, unsigned b) { return a < (long long)b ? a : (int)b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline size_t min(size_t a, int b) { return a < (size_t)b ? a : (size_t)b; }
static inline size_t max(size_t a, int b) { return a > (size_t)b ? a : (size_t)b; }

// ------------------------------------------------------------------ runtime API (host memory, one "device")
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;
struct fpb_emu_event { double t_ms; };
typedef fpb_emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocMapped = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaDeviceAttr {
  cudaDevAttrMaxSharedMemoryPerBlockOptin, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cudaDevAttrMultiProcessorCount,
  cudaDevAttrMaxThreadsPerBlock
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
struct cudaFuncAttributes { size_t sharedSizeBytes = 2048; int numRegs = 64; int maxThreadsPerBlock = 1024; size_t localSizeBytes = 0; };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount, major, minor;
  size_t totalGlobalMem, sharedMemPerBlockOptin, sharedMemPerMultiprocessor;
  int l2CacheSize;
};
// the emulated device: B200's shared-memory limits (they select the kernel variants), two "SMs" (grids stay small)
constexpr int FPB_EMU_SMS = 2;
constexpr int FPB_EMU_SMEM_OPTIN = 227 * 1024, FPB_EMU_SMEM_SM = 228 * 1024;

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulator error"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  switch (a) {
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = FPB_EMU_SMEM_OPTIN; break;
    case cudaDevAttrMaxSharedMemoryPerMultiprocessor: *v = FPB_EMU_SMEM_SM; break;
    case cudaDevAttrMultiProcessorCount: *v = FPB_EMU_SMS; break;
    case cudaDevAttrMaxThreadsPerBlock: *v = 1024; break;
  }
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "fpb SIMT emulator (CPU)");
  p->multiProcessorCount = FPB_EMU_SMS;
  p->major = 10; p->minor = 0;
  p->totalGlobalMem = (size_t)8 << 30;
  p->sharedMemPerBlockOptin = FPB_EMU_SMEM_OPTIN;
  p->sharedMemPerMultiprocessor = FPB_EMU_SMEM_SM;
  p->l2CacheSize = 126 << 20;
  return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
  *p = (T*)malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) {
  *p = (T*)calloc(n ? n : 1, 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaHostAlloc(p, n, 0); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
template <class T, class U> static inline cudaError_t cudaHostGetDevicePointer(T** d, U* h, unsigned) { *d = (T*)h; return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new fpb_emu_event{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  e->t_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { *a = cudaFuncAttributes(); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int threads, size_t smem) {
  // shared memory and the 2048-thread limit; registers are not modelled
  const int by_smem = (int)((size_t)FPB_EMU_SMEM_SM / (smem + 1024 + 2048));
  const int by_thr = threads > 0 ? 2048 / threads : 0;
  *n = std::min(by_smem, by_thr);
  return cudaSuccess;
}
