// TEST INFRASTRUCTURE -- fiber scheduler of the warp emulator (see include/cuda_runtime.h).
//
// One OS thread.  Every CUDA thread of the current block is a ucontext fiber; fibers run until they
// reach a rendezvous (warp collective or block barrier), deposit their value and hand control back to
// the scheduler, which resumes the next unfinished fiber round-robin.  A rendezvous completes when
// all lanes of the warp (threads of the block) have arrived; lanes can be at most one rendezvous
// apart, so two slot buffers indexed by the parity of the lane's rendezvous counter suffice.
// Blocks of a grid run one after the other.
#include <cuda_runtime.h>
#include <stdio.h>
#include <ucontext.h>

#include <unordered_map>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

// The dynamic shared memory arrays the kernels declare with `extern __shared__` (block-scope extern
// declarations inside namespace madrl resolve to these).
namespace madrl {
constexpr size_t kEmuSmemBytes = 256 * 1024;
__attribute__((aligned(16))) unsigned char ww_smem[kEmuSmemBytes];
__attribute__((aligned(16))) uint32_t smem_u32[kEmuSmemBytes / 4];
}  // namespace madrl

namespace madrl_emu {
namespace {

constexpr size_t kStackBytes = 256 * 1024;

struct Rendezvous {          // per warp (collectives) or per block (barrier)
  uint64_t slots[2][32];
  long epoch[2] = {-1, -1};
  int count[2] = {0, 0};
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  long warp_phase = 0, block_phase = 0;
  long syncs = 0;              // __syncwarp / __syncthreads passed: the memory-ordering epoch
  unsigned tid = 0;
};

// per shared-memory byte: who touched it in which epoch (see smem_access)
struct Shadow {
  long w_epoch = -1, r_epoch = -1, a_epoch = -1;
  int w_tid = -1;
  uint64_t r_tids_lo = 0, r_tids_hi = 0, a_tids_lo = 0, a_tids_hi = 0;   // up to 128 threads per block
};
std::unordered_map<uint32_t, Shadow> g_shadow;
long g_races = 0;

inline void bit_set(uint64_t& lo, uint64_t& hi, unsigned t) { (t < 64 ? lo : hi) |= 1ull << (t & 63); }
inline bool others(uint64_t lo, uint64_t hi, unsigned t) {
  uint64_t l = lo, h = hi;
  (t < 64 ? l : h) &= ~(1ull << (t & 63));
  return (l | h) != 0;
}

void report(const char* what, uint32_t addr, unsigned tid, long epoch) {
  if (++g_races <= 8)
    fprintf(stderr, "madrl_emu: shared-memory %s hazard at offset %u: thread %u (block %u), no __syncwarp/__syncthreads since the conflicting access (epoch %ld)\n",
            what, addr, tid, blockIdx.x, epoch);
}

std::vector<Fiber> g_fibers;
std::vector<Rendezvous> g_warps;
Rendezvous g_block;
ucontext_t g_sched;
int g_cur = -1;
long g_progress = 0;          // bumped whenever a rendezvous completes or a fiber finishes
const std::function<void()>* g_body = nullptr;

void yield_to_scheduler() {
  Fiber& f = g_fibers[g_cur];
  swapcontext(&f.ctx, &g_sched);
}

void fiber_entry() {
  (*g_body)();
  g_fibers[g_cur].done = true;
  ++g_progress;
  yield_to_scheduler();
  abort();   // a finished fiber is never resumed
}

}  // namespace

unsigned char* smem_anchor() { return madrl::ww_smem; }

void note_sync() { ++g_fibers[g_cur].syncs; }

extern "C" long madrl_emu_race_count() { return g_races; }
extern "C" void madrl_emu_race_reset() { g_races = 0; }

void smem_access(uint32_t addr, uint32_t bytes, SmemKind kind) {
  const Fiber& f = g_fibers[g_cur];
  const long e = f.syncs;
  const unsigned t = f.tid;
  for (uint32_t b = 0; b < bytes; ++b) {
    Shadow& s = g_shadow[addr + b];
    const bool w_conf = s.w_epoch == e && s.w_tid != (int)t;
    const bool r_conf = s.r_epoch == e && others(s.r_tids_lo, s.r_tids_hi, t);
    const bool a_conf = s.a_epoch == e && others(s.a_tids_lo, s.a_tids_hi, t);
    if (kind == SMEM_READ) {
      if (w_conf) { report("read-after-write", addr + b, t, e); }
      else if (a_conf) { report("read-vs-atomic", addr + b, t, e); }
      if (s.r_epoch != e) { s.r_epoch = e; s.r_tids_lo = s.r_tids_hi = 0; }
      bit_set(s.r_tids_lo, s.r_tids_hi, t);
    } else if (kind == SMEM_WRITE) {
      if (w_conf) { report("write-after-write", addr + b, t, e); }
      else if (r_conf) { report("write-after-read", addr + b, t, e); }
      else if (a_conf) { report("write-vs-atomic", addr + b, t, e); }
      s.w_epoch = e; s.w_tid = (int)t;
    } else {
      if (w_conf) { report("atomic-vs-write", addr + b, t, e); }
      else if (r_conf) { report("atomic-vs-read", addr + b, t, e); }
      if (s.a_epoch != e) { s.a_epoch = e; s.a_tids_lo = s.a_tids_hi = 0; }
      bit_set(s.a_tids_lo, s.a_tids_hi, t);
    }
    if (w_conf || r_conf || a_conf) break;   // one report per access
  }
}

void check_smem(size_t bytes) {
  if (bytes > madrl::kEmuSmemBytes) {
    fprintf(stderr, "madrl_emu: %zu bytes of dynamic shared memory requested, emulator has %zu\n", bytes,
            madrl::kEmuSmemBytes);
    abort();
  }
}

const uint64_t* warp_exchange(uint64_t v) {
  Fiber& f = g_fibers[g_cur];
  Rendezvous& r = g_warps[f.tid >> 5];
  const long ph = f.warp_phase++;
  const int p = (int)(ph & 1);
  if (r.epoch[p] != ph) { r.epoch[p] = ph; r.count[p] = 0; }
  r.slots[p][f.tid & 31] = v;
  if (++r.count[p] == 32) ++g_progress;
  while (r.count[p] < 32 || r.epoch[p] != ph) yield_to_scheduler();
  return r.slots[p];
}

void block_barrier() {
  Fiber& f = g_fibers[g_cur];
  const long ph = f.block_phase++;
  const int p = (int)(ph & 1);
  const int n = (int)g_fibers.size();
  if (g_block.epoch[p] != ph) { g_block.epoch[p] = ph; g_block.count[p] = 0; }
  if (++g_block.count[p] == n) ++g_progress;
  while (g_block.count[p] < n || g_block.epoch[p] != ph) yield_to_scheduler();
}

void run_grid(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
  if (block == 0 || block % 32 != 0) {
    fprintf(stderr, "madrl_emu: block size %u is not a multiple of 32\n", block);
    abort();
  }
  check_smem(smem);
  g_body = &body;
  gridDim = dim3(grid);
  blockDim = dim3(block);
  std::vector<char*> stacks(block);   // fiber stacks, reused by every block of the grid
  for (auto& st : stacks) st = static_cast<char*>(malloc(kStackBytes));
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx = uint3{b, 0, 0};
    // uninitialised shared memory must not be relied upon: poison it
    memset(madrl::ww_smem, 0xCD, smem);      // only what the launch asked for is addressable
    memset(madrl::smem_u32, 0xCD, smem);
    g_shadow.clear();
    g_fibers.assign(block, Fiber());
    g_warps.assign(block / 32, Rendezvous());
    g_block = Rendezvous();
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = g_fibers[t];
      f.tid = t;
      f.stack = stacks[t];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStackBytes;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, fiber_entry, 0);
    }
    unsigned remaining = block;
    long seen = g_progress;
    unsigned idle_rounds = 0;
    while (remaining > 0) {
      unsigned ran = 0;
      for (unsigned t = 0; t < block; ++t) {
        Fiber& f = g_fibers[t];
        if (f.done) continue;
        g_cur = (int)t;
        threadIdx = uint3{t, 0, 0};
        swapcontext(&g_sched, &f.ctx);
        ++ran;
        if (f.done) --remaining;
      }
      if (g_progress == seen) {
        // a full round in which no rendezvous completed and no fiber finished: some lanes wait for
        // lanes that exited or took a different path -- a divergent collective, i.e. a kernel bug
        if (++idle_rounds > 2) {
          fprintf(stderr, "madrl_emu: deadlock in block %u (%u fibers waiting at a rendezvous the others never reach)\n", b, ran);
          abort();
        }
      } else {
        idle_rounds = 0;
        seen = g_progress;
      }
    }
  }
  for (char* st : stacks) free(st);
  g_fibers.clear();
  g_cur = -1;
  g_body = nullptr;
}

}  // namespace madrl_emu
