// TEST INFRASTRUCTURE - the SIMT machine behind tests/emu/include/cuda_runtime.h (see the header
// comment there): fibers, the CTA scheduler, warp rendezvous and barriers.  x86-64 System V only.
#include <cuda_runtime.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace fpb_emu {

Machine M;
alignas(128) unsigned char dyn_smem_buf[DYN_SMEM_BYTES];
unsigned char smem_anchor;

// save the callee-saved registers on the current stack, swap stack pointers, restore
asm(R"(
.text
.globl fpb_emu_switch
.type fpb_emu_switch,@function
fpb_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size fpb_emu_switch,.-fpb_emu_switch
)");

static void set_thread_index(int tid) {
  const dim3& b = M.block_dim;
  threadIdx.x = (unsigned)tid % b.x;
  threadIdx.y = ((unsigned)tid / b.x) % b.y;
  threadIdx.z = (unsigned)tid / (b.x * b.y);
}

void yield() {
  const int me = M.cur;
  fpb_emu_switch(&M.fiber[me].sp, M.sched_sp);
  // resumed: the scheduler has set M.cur and threadIdx for this fiber again
}

static void fiber_main() {
  (*M.body)();
  const int me = M.cur;
  M.fiber[me].done = true;
  M.warp[me >> 5].alive &= ~(1u << (me & 31));
  --M.live;
  ++M.progress;
  for (;;) yield();  // never resumed
}

static void prepare_fiber(int tid) {
  unsigned char* top = M.stacks + (size_t)(tid + 1) * STACK_BYTES;
  void** sp = (void**)((uintptr_t)top & ~(uintptr_t)15);
  *--sp = nullptr;               // keeps the entry's stack 16-byte aligned after `ret`
  *--sp = (void*)fiber_main;     // return address popped by fpb_emu_switch
  for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
  M.fiber[tid].sp = sp;
  M.fiber[tid].done = false;
}

unsigned long long* warp_exchange(unsigned mask, unsigned long long mine, unsigned* participants) {
  const int tid = M.cur;
  Warp& w = M.warp[tid >> 5];
  const int lane = tid & 31;
  const long long g = w.lane_gen[lane]++;
  const int s = (int)(g & 1);
  w.val[s][lane] = mine;
  ++w.count[s];
  for (;;) {
    if (w.release_gen >= g) break;
    const unsigned expect = mask & w.alive;
    if (w.count[s] >= __builtin_popcount(expect)) {  // everyone who can still arrive has arrived
      w.count[s] = 0;
      w.part[s] = expect;  // the lanes whose values are valid (a lane may exit before the others read them)
      w.release_gen = g;
      ++M.progress;
      break;
    }
    yield();
  }
  *participants = w.part[s];
  return w.val[s];
}

int named_barrier(int id, int n_threads, int pred) {
  NamedBarrier& b = M.bar[id];
  const long long g = b.gen;
  ++b.count;
  b.pred_or |= pred ? 1 : 0;
  for (;;) {
    if (b.gen > g) break;
    const int expect = n_threads < 0 ? M.live : n_threads;
    if (b.count >= expect) {
      b.result = b.pred_or;
      b.count = 0;
      b.pred_or = 0;
      ++b.gen;
      ++M.progress;
      break;
    }
    yield();
  }
  return b.result;
}

// FPB_EMU_SHUFFLE=<seed>: visit warps and lanes in a pseudo-random order that changes from pass to pass, to run the
// kernels under other interleavings than the fixed round-robin one
static unsigned long long shuffle_state = 0;
static bool shuffle_on = false;
static unsigned next_random() {
  shuffle_state = shuffle_state * 6364136223846793005ull + 1442695040888963407ull;
  return (unsigned)(shuffle_state >> 33);
}
static void fill_order(int* order, int n) {
  for (int i = 0; i < n; ++i) order[i] = i;
  if (shuffle_on)
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[next_random() % (unsigned)(i + 1)]);
}

static void run_cta() {
  const int n = M.n_threads;
  const int n_warps = (n + 31) / 32;
  memset(M.warp, 0, sizeof(Warp) * (size_t)n_warps);
  memset(M.bar, 0, sizeof M.bar);
  for (int w = 0; w < n_warps; ++w) {
    M.warp[w].release_gen = -1;
    const int lanes = std::min(32, n - 32 * w);
    M.warp[w].alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
  }
  for (int t = 0; t < n; ++t) prepare_fiber(t);
  M.live = n;
  int warp_order[MAX_THREADS / 32], lane_order[32];
  while (M.live > 0) {
    const long long before_cta = M.progress;
    fill_order(warp_order, n_warps);
    for (int wi = 0; wi < n_warps && M.live > 0; ++wi) {
      const int w = warp_order[wi];
      // lanes of one warp round-robin for as long as the warp gets anywhere on its own
      // (shuffled runs: one pass per visit, so that warps interleave at every collective)
      for (;;) {
        const long long before = M.progress;
        fill_order(lane_order, 32);
        for (int li = 0; li < 32; ++li) {
          const int l = lane_order[li];
          const int t = 32 * w + l;
          if (t >= n || M.fiber[t].done) continue;
          M.cur = t;
          set_thread_index(t);
          fpb_emu_switch(&M.sched_sp, M.fiber[t].sp);
        }
        if (M.progress == before || shuffle_on) break;
      }
    }
    if (M.progress == before_cta && M.live > 0) {
      fprintf(stderr, "fpb_emu: deadlock - %d live threads of block (%u,%u,%u) wait for a barrier or collective nobody completes\n",
              M.live, blockIdx.x, blockIdx.y, blockIdx.z);
      abort();
    }
  }
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  std::lock_guard<std::recursive_mutex> guard(M.lock);
  const int n = (int)(block.x * block.y * block.z);
  if (n <= 0 || n > MAX_THREADS || smem > DYN_SMEM_BYTES) {
    fprintf(stderr, "fpb_emu: launch with %d threads, %zu bytes of dynamic shared memory\n", n, smem);
    abort();
  }
  if (!M.stacks) {
    if (const char* sv = getenv("FPB_EMU_SHUFFLE")) {
      shuffle_on = true;
      shuffle_state = strtoull(sv, nullptr, 10) * 2654435761ull + 88172645463325252ull;
    }
    M.stacks = (unsigned char*)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE,
                                    MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (M.stacks == (unsigned char*)MAP_FAILED) abort();
  }
  M.body = &body;
  M.n_threads = n;
  M.block_dim = block;
  M.grid_dim = grid;
  blockDim = block;
  gridDim = grid;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
        run_cta();
      }
  M.body = nullptr;
}

}  // namespace fpb_emu
