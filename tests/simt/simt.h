// simt.h -- a small SIMT interpreter: runs a CUDA kernel's threads as cooperative fibers on ONE host
// thread, with real warp and block collectives.  TEST INFRASTRUCTURE (tests/simt): it lets the unit
// tests execute the actual kernel sources of torchmd_b200/csrc on the CPU to check their LOGIC
// (indexing, masks, list building, band handling, double buffering ...) without a GPU.  It says
// nothing about performance and is never part of the product.
//
// Model: blocks of a grid run one after the other; the threads of a block are ucontext fibers
// scheduled round-robin; a fiber runs until it reaches a collective (__shfl*/__ballot/... or
// __syncthreads) where it parks until every live participant has arrived.  Threads that have
// returned do not take part (CUDA semantics).  __activemask() first lets every other fiber advance
// to its next parking point, then reports the lanes of the warp that are still alive.
#pragma once
#include <setjmp.h>
#include <ucontext.h>

#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

namespace simt {

struct uint3_ { unsigned x, y, z; };

enum Kind { K_NONE, K_BALLOT, K_ANY, K_SHFL, K_SHFL_XOR, K_SHFL_UP, K_RMIN, K_RMAX, K_SYNCWARP };

struct Warp {
  unsigned live = 0;       // lanes that have not returned
  unsigned arrived = 0;    // lanes parked at the current collective
  unsigned req_mask = 0;   // mask argument of the current collective
  unsigned gen = 0;        // bumped when a collective completes
  int kind = K_NONE;
  unsigned long long in[32], aux[32], out[32];
};

// Context switches: a fiber is ENTERED through ucontext (makecontext/swapcontext), every later switch is a
// _setjmp/_longjmp pair -- glibc's swapcontext saves and restores the signal mask with a system call per switch,
// which was a third of the suite's run time.  -DSIMT_UCONTEXT_ONLY keeps swapcontext throughout (AddressSanitizer
// build: it follows swapcontext but not a longjmp between stacks).  The build passes -U_FORTIFY_SOURCE: the
// fortified longjmp refuses jumps to a lower stack address.
struct Fiber {
  ucontext_t ctx;
  jmp_buf jb;
  bool done = false;
  bool started = false;
};

struct Block {
  dim3 grid, block, bidx;
  int nthreads = 0;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  ucontext_t sched;
  jmp_buf sched_jb;
  int cur = -1;           // running fiber
  int live = 0;           // threads that have not returned
  int bar_arrived = 0;    // __syncthreads
  unsigned bar_gen = 0;
  int bar_or = 0, bar_or_result = 0;
  const std::function<void()>* body = nullptr;
  long long clock = 0;
};

inline Block*& current() {
  static Block* b = nullptr;
  return b;
}
inline int tid() { return current()->cur; }
inline int lane() { return current()->cur & 31; }
inline Warp& warp() { return current()->warps[current()->cur >> 5]; }
inline void to_scheduler(Block* b, Fiber& f) {
#if defined(SIMT_UCONTEXT_ONLY)
  swapcontext(&f.ctx, &b->sched);
#else
  if (_setjmp(f.jb) == 0) _longjmp(b->sched_jb, 1);
#endif
}
inline void to_fiber(Block& b, Fiber& f) {
#if defined(SIMT_UCONTEXT_ONLY)
  swapcontext(&b.sched, &f.ctx);
#else
  if (_setjmp(b.sched_jb) == 0) {
    if (f.started) _longjmp(f.jb, 1);
    f.started = true;
    swapcontext(&b.sched, &f.ctx);  // first entry; the fiber comes back through sched_jb
  }
#endif
}
inline void yield() {
  Block* b = current();
  to_scheduler(b, b->fibers[b->cur]);
}
inline long long fake_clock() { return current()->clock += 1000; }

inline void complete(Warp& w) {
  const unsigned part = w.arrived;
  switch (w.kind) {
    case K_BALLOT: {
      unsigned r = 0;
      for (int l = 0; l < 32; ++l)
        if ((part >> l & 1u) && w.in[l]) r |= 1u << l;
      for (int l = 0; l < 32; ++l) w.out[l] = r;
      break;
    }
    case K_ANY: {
      unsigned long long r = 0;
      for (int l = 0; l < 32; ++l)
        if ((part >> l & 1u) && w.in[l]) r = 1;
      for (int l = 0; l < 32; ++l) w.out[l] = r;
      break;
    }
    case K_SHFL:
      for (int l = 0; l < 32; ++l) w.out[l] = w.in[w.aux[l] & 31];
      break;
    case K_SHFL_XOR:
      for (int l = 0; l < 32; ++l) w.out[l] = w.in[(l ^ (int)w.aux[l]) & 31];
      break;
    case K_SHFL_UP:
      for (int l = 0; l < 32; ++l) w.out[l] = (l >= (int)w.aux[l]) ? w.in[l - (int)w.aux[l]] : w.in[l];
      break;
    case K_RMIN:
    case K_RMAX: {
      bool first = true;
      long long r = 0;
      for (int l = 0; l < 32; ++l)
        if (part >> l & 1u) {
          const long long v = (long long)w.in[l];
          if (first || (w.kind == K_RMIN ? v < r : v > r)) r = v;
          first = false;
        }
      for (int l = 0; l < 32; ++l) w.out[l] = (unsigned long long)r;
      break;
    }
    default:
      break;
  }
  w.arrived = 0;
  w.kind = K_NONE;
  w.gen++;
}

// One warp collective: every live lane of `mask` deposits (value, aux) and gets out[lane].
inline unsigned long long collective(int kind, unsigned mask, unsigned long long value, unsigned long long aux) {
  Warp& w = warp();
  const int l = lane();
  const unsigned mygen = w.gen;
  w.kind = kind;
  w.req_mask = mask;
  w.in[l] = value;
  w.aux[l] = aux;
  w.arrived |= 1u << l;
  if (w.arrived == (mask & w.live)) complete(w);
  else
    while (w.gen == mygen) yield();
  return w.out[l];
}

inline void block_barrier(int pred) {
  Block* b = current();
  const unsigned mygen = b->bar_gen;
  b->bar_or |= (pred != 0);
  if (++b->bar_arrived == b->live) {
    b->bar_or_result = b->bar_or;
    b->bar_or = 0;
    b->bar_arrived = 0;
    b->bar_gen++;
  } else {
    while (b->bar_gen == mygen) yield();
  }
}

inline void fiber_main() {
  Block* b = current();
  (*b->body)();
  // the thread returns: it leaves its warp and the block; collectives the others wait in may complete now
  const int t = b->cur;
  Warp& w = b->warps[t >> 5];
  w.live &= ~(1u << (t & 31));
  b->live--;
  b->fibers[t].done = true;
  if (w.arrived && w.arrived == (w.req_mask & w.live)) complete(w);
  if (b->live > 0 && b->bar_arrived == b->live) {
    b->bar_or_result = b->bar_or;
    b->bar_or = 0;
    b->bar_arrived = 0;
    b->bar_gen++;
  }
  to_scheduler(b, b->fibers[t]);
}

// Scheduling order, from the environment (read once): SIMT_SCHEDULE = "forward" (default), "reverse", or
// "random:<seed>" (a new permutation of the threads every scheduler round; blocks of a grid in a shuffled order too).
// A kernel without races between its synchronisation points gives the same results under every order (up to the
// order of floating-point atomics); one that relies on lane 3 running before lane 5 does not.
struct Schedule {
  int mode = 0;  // 0 forward, 1 reverse, 2 random
  unsigned long long state = 0x9E3779B97F4A7C15ull;
  Schedule() {
    const char* e = getenv("SIMT_SCHEDULE");
    if (!e) return;
    if (!strcmp(e, "reverse")) mode = 1;
    else if (!strncmp(e, "random", 6)) {
      mode = 2;
      if (e[6] == ':') state ^= strtoull(e + 7, nullptr, 10) * 0xD1342543DE82EF95ull;
    }
  }
  unsigned next() {  // xorshift64*
    state ^= state >> 12;
    state ^= state << 25;
    state ^= state >> 27;
    return (unsigned)((state * 0x2545F4914F6CDD1Dull) >> 33);
  }
  void order(std::vector<int>& idx) {
    const int n = (int)idx.size();
    if (mode == 0) for (int i = 0; i < n; ++i) idx[i] = i;
    else if (mode == 1) for (int i = 0; i < n; ++i) idx[i] = n - 1 - i;
    else {
      for (int i = 0; i < n; ++i) idx[i] = i;
      for (int i = n - 1; i > 0; --i) {
        const int j = (int)(next() % (unsigned)(i + 1));
        const int t = idx[i];
        idx[i] = idx[j];
        idx[j] = t;
      }
    }
  }
};
inline Schedule& schedule() {
  static Schedule s;
  return s;
}

inline void run_block(Block& b) {
  constexpr size_t STACK = 192 * 1024;
  b.nthreads = (int)(b.block.x * b.block.y * b.block.z);
  b.fibers.clear();
  b.fibers.resize(b.nthreads);
  b.warps.assign((b.nthreads + 31) / 32, Warp());
  b.live = b.nthreads;
  b.bar_arrived = 0;
  b.bar_or = 0;
  for (int t = 0; t < b.nthreads; ++t) {
    Fiber& f = b.fibers[t];
    // stacks are kept for the next block (blocks run one after the other): no mmap/munmap and page faults per fiber
    static std::vector<std::unique_ptr<char[]>> pool;
    if ((int)pool.size() <= t) pool.resize(t + 1);
    if (!pool[t]) pool[t].reset(new char[STACK]);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = pool[t].get();
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = &b.sched;
    makecontext(&f.ctx, (void (*)())fiber_main, 0);
    b.warps[t >> 5].live |= 1u << (t & 31);
  }
  Block* saved = current();
  current() = &b;
  int remaining = b.nthreads;
  long long idle_rounds = 0;
  std::vector<int> turn(b.nthreads);
  schedule().order(turn);
  while (remaining > 0) {
    int progressed = 0;
    if (schedule().mode == 2) schedule().order(turn);
    for (int k = 0; k < b.nthreads; ++k) {
      const int t = turn[k];
      if (b.fibers[t].done) continue;
      b.cur = t;
      to_fiber(b, b.fibers[t]);
      if (b.fibers[t].done) {
        --remaining;
        ++progressed;
      }
    }
    // a full round in which nobody finished is normal (threads parked at collectives make progress
    // without finishing); a very long streak means a deadlock in the kernel under test
    idle_rounds = progressed ? 0 : idle_rounds + 1;
    if (idle_rounds > 50000000) {
      fprintf(stderr, "simt: no thread finished for 5e7 scheduler rounds -- deadlock in the kernel under test?\n");
      abort();
    }
  }
  current() = saved;
}

template <typename F>
inline void run_grid(dim3 grid, dim3 block, F&& f) {
  const std::function<void()> body(f);
  const unsigned nblocks = grid.x * grid.y * grid.z;
  std::vector<int> turn(nblocks);
  schedule().order(turn);
  for (unsigned k = 0; k < nblocks; ++k) {
    const unsigned id = (unsigned)turn[k];
    Block b;
    b.grid = grid;
    b.block = block;
    b.bidx = dim3(id % grid.x, (id / grid.x) % grid.y, id / (grid.x * grid.y));
    b.body = &body;
    run_block(b);
  }
}

struct ThreadIdxProxy {
  struct X { operator unsigned() const { return (unsigned)tid() % current()->block.x; } } x;
  struct Y { operator unsigned() const { return ((unsigned)tid() / current()->block.x) % current()->block.y; } } y;
  struct Z { operator unsigned() const { return (unsigned)tid() / (current()->block.x * current()->block.y); } } z;
};
struct BlockIdxProxy {
  struct X { operator unsigned() const { return current()->bidx.x; } } x;
  struct Y { operator unsigned() const { return current()->bidx.y; } } y;
  struct Z { operator unsigned() const { return current()->bidx.z; } } z;
};
struct BlockDimProxy {
  struct X { operator unsigned() const { return current()->block.x; } } x;
  struct Y { operator unsigned() const { return current()->block.y; } } y;
  struct Z { operator unsigned() const { return current()->block.z; } } z;
};
struct GridDimProxy {
  struct X { operator unsigned() const { return current()->grid.x; } } x;
  struct Y { operator unsigned() const { return current()->grid.y; } } y;
  struct Z { operator unsigned() const { return current()->grid.z; } } z;
};

}  // namespace simt

static simt::ThreadIdxProxy threadIdx;
static simt::BlockIdxProxy blockIdx;
static simt::BlockDimProxy blockDim;
static simt::GridDimProxy gridDim;

// ---- collectives in CUDA spelling -----------------------------------------------------------------
inline void __syncthreads() { simt::block_barrier(0); }
inline int __syncthreads_or(int p) {
  simt::block_barrier(p);
  return simt::current()->bar_or_result;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::collective(simt::K_SYNCWARP, mask, 0, 0); }
inline unsigned __ballot_sync(unsigned mask, int p) { return (unsigned)simt::collective(simt::K_BALLOT, mask, p != 0, 0); }
inline int __any_sync(unsigned mask, int p) { return (int)simt::collective(simt::K_ANY, mask, p != 0, 0); }
inline unsigned __activemask() {
  simt::yield();  // let every other thread reach its next parking point (or return) first
  if (simt::schedule().mode == 2) simt::yield();  // orders change between rounds: two rounds reach everyone
  return simt::warp().live;
}
template <typename T>
inline unsigned long long simt_bits(T v) {
  unsigned long long b = 0;
  static_assert(sizeof(T) <= 8, "shuffle payload");
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
inline T simt_unbits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src) { return simt_unbits<T>(simt::collective(simt::K_SHFL, mask, simt_bits(v), (unsigned)src)); }
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int m) { return simt_unbits<T>(simt::collective(simt::K_SHFL_XOR, mask, simt_bits(v), (unsigned)m)); }
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d) { return simt_unbits<T>(simt::collective(simt::K_SHFL_UP, mask, simt_bits(v), d)); }
inline int __reduce_min_sync(unsigned mask, int v) { return (int)(long long)simt::collective(simt::K_RMIN, mask, (unsigned long long)(long long)v, 0); }
inline int __reduce_max_sync(unsigned mask, int v) { return (int)(long long)simt::collective(simt::K_RMAX, mask, (unsigned long long)(long long)v, 0); }
