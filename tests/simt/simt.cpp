// simt.cpp - the fiber scheduler of tests/simt/simt.h (TEST INFRASTRUCTURE ONLY).
#include "simt.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

static int env_int(const char *name, int fallback) {
  const char *v = std::getenv(name);
  return v && *v ? std::atoi(v) : fallback;
}
// ---- "device" memory: host memory, or named shared-memory segments when SIMT_IPC=1 (see simt_runtime.h) -------------
namespace {
struct Segment { std::string name; size_t bytes; bool owner; };
std::mutex g_mem_mu;
std::map<void *, Segment> g_segments;        // mapped segments of this process: own allocations and opened peers
unsigned long g_segment_counter = 0;
bool ipc_mode() { static const bool on = env_int("SIMT_IPC", 0) != 0; return on; }
void unlink_all() {
  for (auto &kv : g_segments)
    if (kv.second.owner) shm_unlink(kv.second.name.c_str());
}
}  // namespace

void *simt_device_alloc(size_t bytes) {
  bytes = (bytes + 4095) / 4096 * 4096;
  if (bytes == 0) bytes = 4096;
  if (!ipc_mode()) return std::aligned_alloc(4096, bytes);
  std::lock_guard<std::mutex> lock(g_mem_mu);
  static bool registered = false;
  if (!registered) { registered = true; std::atexit(unlink_all); }
  char name[64];
  std::snprintf(name, sizeof name, "/nfh_simt_%ld_%lu", (long) getpid(), g_segment_counter++);
  const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (ftruncate(fd, (off_t) bytes) != 0) { close(fd); shm_unlink(name); return nullptr; }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(name); return nullptr; }
  g_segments[p] = Segment{name, bytes, true};
  return p;
}
void simt_device_free(void *p) {
  if (!p) return;
  if (!ipc_mode()) { std::free(p); return; }
  std::lock_guard<std::mutex> lock(g_mem_mu);
  auto it = g_segments.find(p);
  if (it == g_segments.end()) return;
  munmap(p, it->second.bytes);
  if (it->second.owner) shm_unlink(it->second.name.c_str());
  g_segments.erase(it);
}
int simt_ipc_export(void *p, char name_out[64]) {
  if (!ipc_mode()) { std::memcpy(name_out, &p, sizeof p); return 0; }          // same process: the pointer itself
  std::lock_guard<std::mutex> lock(g_mem_mu);
  auto it = g_segments.find(p);
  if (it == g_segments.end() || it->second.name.size() >= 56) return 1;        // only whole allocations are exported
  std::memset(name_out, 0, 64);
  std::memcpy(name_out, it->second.name.c_str(), it->second.name.size());
  std::memcpy(name_out + 56, &it->second.bytes, 8);
  return 0;
}
void *simt_ipc_open(const char name[64]) {
  if (!ipc_mode()) { void *p; std::memcpy(&p, name, sizeof p); return p; }
  size_t bytes;
  std::memcpy(&bytes, name + 56, 8);
  const int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return nullptr;
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return nullptr;
  std::lock_guard<std::mutex> lock(g_mem_mu);
  g_segments[p] = Segment{std::string(name), bytes, false};
  return p;
}
void simt_ipc_close(void *p) {
  if (!ipc_mode()) return;
  std::lock_guard<std::mutex> lock(g_mem_mu);
  auto it = g_segments.find(p);
  if (it == g_segments.end()) return;
  munmap(p, it->second.bytes);
  g_segments.erase(it);
}

int simt_device_count() { return env_int("SIMT_DEVICES", 1); }
int simt_sm_count() { return env_int("SIMT_SMS", 4); }

// Fiber switch.  swapcontext() makes a sigprocmask system call per switch, and an emulated EM run switches fibers a
// few hundred million times; on x86-64 the switch is done by hand instead: push the callee-saved registers, swap the
// stack pointers, pop, return.  Everything else (and any other architecture) uses ucontext.
#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_ASM_SWITCH 1
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");
#endif

namespace simt {

// all emulator state is per host thread: several host threads may each run kernels (one "rank" each)
static thread_local Cta g_cta;
static thread_local unsigned long long g_launches = 0, g_switches = 0;
alignas(1024) static thread_local unsigned char g_dyn_smem[256 * 1024];
constexpr size_t kStackBytes = 256 * 1024;

Cta &cta() { return g_cta; }
unsigned char *dyn_smem() { return g_dyn_smem; }
unsigned long long launches() { return g_launches; }
unsigned long long switches() { return g_switches; }

static inline void to_scheduler() {
  Cta &c = g_cta;
#ifdef SIMT_ASM_SWITCH
  simt_switch(&c.fibers[c.cur].sp, c.sched_sp);
#else
  swapcontext(&c.fibers[c.cur].ctx, &c.sched);
#endif
}
static inline void to_fiber(unsigned t) {
  Cta &c = g_cta;
#ifdef SIMT_ASM_SWITCH
  simt_switch(&c.sched_sp, c.fibers[t].sp);
#else
  swapcontext(&c.sched, &c.fibers[t].ctx);
#endif
}

static thread_local std::vector<PendingCopy> g_loads, g_stores;
std::vector<PendingCopy> &pending_loads() { return g_loads; }
std::vector<PendingCopy> &pending_stores() { return g_stores; }
void flush_loads(uint64_t *bar) {
  for (size_t i = 0; i < g_loads.size();) {
    if (g_loads[i].bar == bar) {
      const std::function<void()> run = g_loads[i].run;
      g_loads.erase(g_loads.begin() + (long) i);
      run();
    } else {
      i++;
    }
  }
}
void flush_stores() {
  std::vector<PendingCopy> todo;
  todo.swap(g_stores);
  for (auto &c : todo) c.run();
}
void poison(void *p, size_t bytes) {
  const uint64_t nan = 0x7ff8dead0000beefull;      // a quiet NaN: any arithmetic on it stays NaN and raises the kernels' flag
  for (size_t i = 0; i + 8 <= bytes; i += 8) std::memcpy((char *) p + i, &nan, 8);
}

void yield() {
  g_switches++;
  to_scheduler();
}

static void fiber_main() {
  Cta &c = g_cta;
  (*c.body)();
  c.fibers[c.cur].done = true;
  // an exited thread no longer takes part in barriers: release whoever waits for it
  c.live--;
  if (c.live && c.bar_arrived >= c.live) { c.bar_arrived = 0; c.bar_gen++; }
  Cta::Warp &w = c.warps[c.cur >> 5];
  w.live--;
  if (w.live && w.arrived >= w.live) { w.arrived = 0; w.gen++; }
  for (;;) to_scheduler();                 // never resumed; never returns
}

static void set_thread(unsigned t) {
  Cta &c = g_cta;
  c.cur = t;
  threadIdx.x = t % c.block.x;
  threadIdx.y = (t / c.block.x) % c.block.y;
  threadIdx.z = t / (c.block.x * c.block.y);
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &kernel_call) {
  Cta &c = g_cta;
  const unsigned n = block.x * block.y * block.z;
  if (n == 0 || n > 1024 || dyn_smem_bytes > sizeof(g_dyn_smem)) { std::fprintf(stderr, "simt: bad launch\n"); std::abort(); }
  g_launches++;
  // SIMT_ORDER: the order in which the scheduler resumes the fibers of a CTA in every round - "forward" (thread 0
  // first, the default), "reverse", or "random:<seed>".  Results must not depend on it: a fiber runs undisturbed
  // between two rendezvous points, so a different order is a different legal interleaving of the CTA's threads at
  // barrier granularity, and a missing barrier (a read that only works because thread 0 happened to run first) shows
  // up as a changed result or a deadlock.
  int order_mode = 0;
  static thread_local uint64_t rng = 0;
  if (const char *o = std::getenv("SIMT_ORDER")) {
    if (!std::strcmp(o, "reverse")) order_mode = 1;
    else if (!std::strncmp(o, "random:", 7)) { order_mode = 2; if (!rng) rng = 0x9e3779b97f4a7c15ull ^ std::strtoull(o + 7, nullptr, 10); }
  }
  auto next_random = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  std::vector<unsigned> order(n);
  for (unsigned t = 0; t < n; t++) order[t] = order_mode == 1 ? n - 1 - t : t;
  if (c.stacks.size() < (size_t) n * kStackBytes) c.stacks.resize((size_t) n * kStackBytes);
  c.fibers.resize(n);
  c.n_threads = n;
  c.block = block;
  c.body = &kernel_call;
  blockDim = block;
  gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        c.live = n; c.bar_arrived = 0;
        for (unsigned w = 0; w < (n + 31) / 32; w++) {
          c.warps[w].live = std::min(32u, n - 32 * w);
          c.warps[w].arrived = 0;
        }
        for (unsigned t = 0; t < n; t++) {
          Fiber &f = c.fibers[t];
          f.done = false;
#ifdef SIMT_ASM_SWITCH
          // initial frame: six callee-saved registers, the entry point as return address; after the pops and the
          // `ret` the stack pointer is 8 modulo 16, as at any function entry
          uintptr_t top = (uintptr_t) (c.stacks.data() + (size_t) (t + 1) * kStackBytes) & ~(uintptr_t) 15;
          void **frame = reinterpret_cast<void **>(top - 64);
          for (int k = 0; k < 6; k++) frame[k] = nullptr;
          frame[6] = (void *) &fiber_main;
          frame[7] = nullptr;
          f.sp = frame;
#else
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = c.stacks.data() + (size_t) t * kStackBytes;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_main, 0);
#endif
        }
        unsigned remaining = n;
        unsigned long long idle_rounds = 0;
        while (remaining) {
          const unsigned long long before = g_switches;
          unsigned finished_now = 0;
          if (order_mode == 2)                          // a fresh permutation every round
            for (unsigned k = n - 1; k > 0; k--) std::swap(order[k], order[next_random() % (k + 1)]);
          for (unsigned k = 0; k < n; k++) {
            const unsigned t = order[k];
            if (c.fibers[t].done) continue;
            set_thread(t);
            to_fiber(t);
            if (c.fibers[t].done) { remaining--; finished_now++; }
          }
          // every live fiber yielded and nobody finished for a very long time: a deadlock (missed barrier) in the kernel
          idle_rounds = finished_now ? 0 : idle_rounds + 1;
          if (idle_rounds > 2000000ull) { std::fprintf(stderr, "simt: no progress in block (%u,%u) - deadlock?\n", bx, by); std::abort(); }
          (void) before;
        }
        // a CTA may not end with a load nobody waited for; stores are drained as the hardware drains them at exit
        if (!g_loads.empty()) { std::fprintf(stderr, "simt: block (%u,%u) ended with %zu bulk loads nobody waited for\n", bx, by, g_loads.size()); std::abort(); }
        flush_stores();
      }
}

}  // namespace simt
