// tools/cusim/cusim.cpp -- DEVELOPMENT TOOL, NOT PRODUCT (see cuda_runtime.h in this directory).
//
// Execution model of the simulator:
//   * a launch runs to completion before cusim::launch() returns (all "streams" are synchronous);
//   * min(grid, CUSIM_WORKERS) OS threads take CTAs in blockIdx order, so a CTA only ever waits for CTAs with a lower
//     index that are running or finished -- the same forward-progress guarantee the kernels rely on on the GPU
//     (decoupled look-back in k_encode_fused);
//   * inside a CTA every CUDA thread is a fiber (own stack, hand-written x86-64 context switch); the fibers are
//     scheduled round-robin and only switch at __syncthreads / warp collectives / __nanosleep, so code between two
//     such points runs without interleaving (a legal schedule; it does not explore races);
//   * a warp collective is a rendezvous keyed by its mask: every lane named in the mask must arrive with the same
//     mask (exited lanes count as arrived); a CTA barrier likewise.  A CTA that makes no progress for
//     CUSIM_DEADLOCK_S seconds (mismatched masks, divergent barriers) aborts with a dump of what every fiber waits on.
#include "cuda_runtime.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/mman.h>

#if !defined(__x86_64__)
#error "cusim's context switch is written for x86-64"
#endif

extern "C" void cusim_switch(void** saveSp, void* loadSp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
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
.size cusim_switch,.-cusim_switch
)");

namespace cusim {
namespace {

constexpr size_t kStackBytes = 512 << 10;

struct Rendezvous {
  unsigned mask = 0, arrived = 0;
  unsigned gen = 0;
  uint64_t vals[32];
  uint64_t snap[2][32];
  unsigned snapPresent[2];
};

struct WarpState {
  unsigned exited = 0;                 // lanes whose thread has returned (or that do not exist)
  std::vector<Rendezvous> rv;
};

struct Fiber {
  void* sp = nullptr;
  bool done = false;
  ThreadCtx ctx;
  const char* waitingOn = "";
  unsigned waitMask = 0;
};

struct Cta {
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  int nThreads = 0, nExited = 0;
  int barCount = 0; unsigned barGen = 0;
  int nbarCount[16] = {0}; unsigned nbarGen[16] = {0};     // named barriers (bar.sync id, nThreads)
  void* schedSp = nullptr;
  Fiber* cur = nullptr;
  const std::function<void()>* body = nullptr;
  unsigned long long progress = 0;     // bumped whenever any rendezvous completes or a fiber exits
};

thread_local Cta* tlCta = nullptr;
thread_local ThreadCtx tlHostCtx;      // threadIdx etc. read outside a kernel (never meaningful)

struct WorkerStacks {                  // per OS thread, reused across launches
  std::vector<void*> stacks;
  ~WorkerStacks() { for (void* s : stacks) munmap(s, kStackBytes); }
  void* get(size_t i) {
    while (stacks.size() <= i) {
      void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (p == MAP_FAILED) { std::perror("cusim: mmap stack"); std::abort(); }
      mprotect(p, 4096, PROT_NONE);    // guard page at the low end
      stacks.push_back(p);
    }
    return stacks[i];
  }
};
thread_local WorkerStacks tlStacks;

void fiberMain() {
  Cta* cta = tlCta;
  Fiber* f = cta->cur;
  (*cta->body)();
  f->done = true;
  cta->nExited++;
  cta->warps[f->ctx.warp].exited |= 1u << f->ctx.lane;
  cta->progress++;
  void* dummy;
  cusim_switch(&dummy, cta->schedSp);
  std::abort();   // never resumed
}

void switchToScheduler() {
  Cta* cta = tlCta;
  Fiber* f = cta->cur;
  cusim_switch(&f->sp, cta->schedSp);
}

int envInt(const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; }

void dumpCta(Cta& cta) {
  std::fprintf(stderr, "[cusim] CTA (%u,%u,%u) made no progress: %d threads, %d exited, barrier count %d\n",
               cta.fibers[0].ctx.bid.x, cta.fibers[0].ctx.bid.y, cta.fibers[0].ctx.bid.z, cta.nThreads, cta.nExited, cta.barCount);
  int shown = 0;
  for (auto& f : cta.fibers)
    if (!f.done && shown++ < 64)
      std::fprintf(stderr, "   thread %d (warp %d lane %d) waits on %s mask %08x\n", f.ctx.linear, f.ctx.warp, f.ctx.lane, f.waitingOn, f.waitMask);
}

void runCta(Cta& cta, dim3 grid, dim3 block, unsigned bx, unsigned by, unsigned bz, void* dyn, const std::function<void()>& body) {
  const int n = (int)(block.x * block.y * block.z);
  cta.nThreads = n; cta.nExited = 0; cta.barCount = 0; cta.barGen = 0; cta.body = &body; cta.progress = 0;
  cta.fibers.assign((size_t)n, Fiber());
  cta.warps.assign((size_t)(n + 31) / 32, WarpState());
  if (n % 32) cta.warps.back().exited = ~0u << (n % 32);
  for (int t = 0; t < n; t++) {
    Fiber& f = cta.fibers[(size_t)t];
    f.ctx.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
    f.ctx.bid = uint3{bx, by, bz};
    f.ctx.bdim = block; f.ctx.gdim = grid;
    f.ctx.linear = t; f.ctx.lane = t & 31; f.ctx.warp = t >> 5; f.ctx.dynSmem = dyn;
    // initial frame: six callee-saved registers, then the entry address in a 16-byte aligned slot
    uintptr_t top = ((uintptr_t)tlStacks.get((size_t)t) + kStackBytes - 64) & ~(uintptr_t)15;
    void** slot = (void**)top;
    slot[0] = (void*)&fiberMain;
    slot[1] = nullptr;
    void** sp = slot - 6;
    for (int i = 0; i < 6; i++) sp[i] = nullptr;
    f.sp = sp;
  }
  tlCta = &cta;
  const int deadlockS = envInt("CUSIM_DEADLOCK_S", 30);
  unsigned long long lastProgress = ~0ull;
  auto lastChange = std::chrono::steady_clock::now();
  // CUSIM_SHUFFLE=<seed>: the fibers of a CTA are resumed in a different pseudo-random order on every pass (warps and lanes alike), to
  // shake out code that only works because lower-numbered threads happen to run first
  const int shuffleSeed = envInt("CUSIM_SHUFFLE", 0);
  std::vector<int> order((size_t)n);
  for (int t = 0; t < n; t++) order[(size_t)t] = t;
  unsigned long long rng = 0x9E3779B97F4A7C15ull * (unsigned long long)(shuffleSeed + 1) + bx * 7919ull + by * 104729ull;
  while (cta.nExited < n) {
    if (shuffleSeed) {
      for (int t = n - 1; t > 0; t--) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[(size_t)t], order[(size_t)(rng % (unsigned long long)(t + 1))]);
      }
    }
    for (int q = 0; q < n; q++) {
      const int t = order[(size_t)q];
      Fiber& f = cta.fibers[(size_t)t];
      if (f.done) continue;
      cta.cur = &f;
      cusim_switch(&cta.schedSp, f.sp);
    }
    if (cta.progress != lastProgress) { lastProgress = cta.progress; lastChange = std::chrono::steady_clock::now(); }
    else {
      std::this_thread::yield();
      if (std::chrono::steady_clock::now() - lastChange > std::chrono::seconds(deadlockS)) { dumpCta(cta); std::abort(); }
    }
  }
  cta.cur = nullptr;
  tlCta = nullptr;
}

}  // namespace

ThreadCtx& self() { return tlCta && tlCta->cur ? tlCta->cur->ctx : tlHostCtx; }

void yield() {
  if (tlCta && tlCta->cur) { tlCta->cur->waitingOn = "yield"; tlCta->progress++; switchToScheduler(); }
  else std::this_thread::yield();
}

void syncthreads() {
  Cta* cta = tlCta;
  Fiber* f = cta->cur;
  const unsigned myGen = cta->barGen;
  cta->barCount++;
  f->waitingOn = "__syncthreads";
  for (;;) {
    if (cta->barGen != myGen) break;
    if (cta->barCount + cta->nExited >= cta->nThreads) { cta->barCount = 0; cta->barGen++; cta->progress++; break; }
    switchToScheduler();
  }
  f->waitingOn = "";
}

void namedBarrier(int id, int nThreads) {
  Cta* cta = tlCta;
  Fiber* f = cta->cur;
  if (id < 1 || id > 15) { std::fprintf(stderr, "[cusim] named barrier id %d out of range\n", id); std::abort(); }
  const unsigned myGen = cta->nbarGen[id];
  cta->nbarCount[id]++;
  f->waitingOn = "named barrier";
  for (;;) {
    if (cta->nbarGen[id] != myGen) break;
    if (cta->nbarCount[id] >= nThreads) { cta->nbarCount[id] = 0; cta->nbarGen[id]++; cta->progress++; break; }
    switchToScheduler();
  }
  f->waitingOn = "";
}

void warpExchange(unsigned mask, uint64_t mine, uint64_t out[32], unsigned* present) {
  Cta* cta = tlCta;
  Fiber* f = cta->cur;
  WarpState& w = cta->warps[(size_t)f->ctx.warp];
  const int lane = f->ctx.lane;
  if (!((mask >> lane) & 1)) {
    std::fprintf(stderr, "[cusim] thread %d (lane %d) calls a warp collective with mask %08x that does not name it\n", f->ctx.linear, lane, mask);
    std::abort();
  }
  size_t idx = w.rv.size();
  for (size_t i = 0; i < w.rv.size(); i++) if (w.rv[i].mask == mask) { idx = i; break; }
  if (idx == w.rv.size()) { w.rv.emplace_back(); w.rv.back().mask = mask; }
  const unsigned myGen = w.rv[idx].gen;
  {
    Rendezvous& r = w.rv[idx];
    if ((r.arrived >> lane) & 1) { std::fprintf(stderr, "[cusim] lane %d arrived twice at the collective with mask %08x\n", lane, mask); std::abort(); }
    r.vals[lane] = mine; r.arrived |= 1u << lane;
  }
  f->waitingOn = "warp collective"; f->waitMask = mask;
  for (;;) {
    Rendezvous& r = w.rv[idx];          // re-fetch: the vector may have grown while this fiber slept
    if (r.gen != myGen) break;
    if (((r.arrived | w.exited) & mask) == mask) {
      std::memcpy(r.snap[myGen & 1], r.vals, sizeof r.vals);
      r.snapPresent[myGen & 1] = r.arrived;
      r.arrived = 0; r.gen++; cta->progress++;
      break;
    }
    switchToScheduler();
  }
  Rendezvous& r = w.rv[idx];
  std::memcpy(out, r.snap[myGen & 1], sizeof r.vals);
  *present = r.snapPresent[myGen & 1];
  f->waitingOn = "";
}

void launch(dim3 grid, dim3 block, size_t smem, void* /*stream*/, const std::function<void()>& body) {
  const unsigned long long nCta = (unsigned long long)grid.x * grid.y * grid.z;
  if (nCta == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024) {
    std::fprintf(stderr, "[cusim] invalid launch configuration grid (%u,%u,%u) block (%u,%u,%u)\n", grid.x, grid.y, grid.z, block.x, block.y, block.z);
    std::abort();
  }
  if (smem > (227u << 10)) { std::fprintf(stderr, "[cusim] %zu bytes of dynamic shared memory exceed 227 KB\n", smem); std::abort(); }
  const int maxWorkers = envInt("CUSIM_WORKERS", 16);
  const int nWorkers = (int)std::min<unsigned long long>(nCta, (unsigned long long)maxWorkers);
  std::atomic<unsigned long long> next{0};
  auto worker = [&]() {
    Cta cta;
    std::vector<uint8_t> dyn(smem + 64);
    void* dynAligned = (void*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    for (;;) {
      const unsigned long long b = next.fetch_add(1);
      if (b >= nCta) break;
      std::memset(dyn.data(), 0xA5, dyn.size());   // shared memory starts with garbage, as on the device
      runCta(cta, grid, block, (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)), dynAligned, body);
    }
  };
  if (nWorkers == 1) { std::thread t(worker); t.join(); }     // always a fresh thread: __shared__ statics are thread_local
  else {
    std::vector<std::thread> pool;
    for (int i = 0; i < nWorkers; i++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
}

}  // namespace cusim

// ---- host API --------------------------------------------------------------------------------------
namespace {
std::mutex gRegMutex;
std::map<uintptr_t, std::pair<size_t, int>> gRegions;   // base -> (size, type)
void registerRegion(void* p, size_t n, int type) { std::lock_guard<std::mutex> l(gRegMutex); gRegions[(uintptr_t)p] = {n, type}; }
}  // namespace

cudaError_t cusimMalloc(void** p, size_t n) {
  void* q = nullptr;
  if (posix_memalign(&q, 512, n ? n : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
  if (n <= ((size_t)256 << 20) && !std::getenv("CUSIM_NO_POISON")) std::memset(q, 0xCD, n);   // device memory is not zero-initialised
  registerRegion(q, n, cudaMemoryTypeDevice);
  *p = q;
  return cudaSuccess;
}
cudaError_t cusimMallocHost(void** p, size_t n) {
  void* q = nullptr;
  if (posix_memalign(&q, 512, n ? n : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
  registerRegion(q, n, cudaMemoryTypeHost);
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  { std::lock_guard<std::mutex> l(gRegMutex); gRegions.erase((uintptr_t)p); }
  std::free(p);
  return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; r++) std::memmove((uint8_t*)dst + r * dpitch, (const uint8_t*)src + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaMemset(void* dst, int v, size_t n) { std::memset(dst, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) { std::memset(dst, v, n); return cudaSuccess; }
struct CusimStream { int id; };
struct CusimEvent { std::chrono::steady_clock::time_point t; };
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new CusimStream{0}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new CusimStream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CusimEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  if (a == cudaDevAttrMultiProcessorCount) { const char* e = std::getenv("CUSIM_SMS"); *v = e ? std::atoi(e) : 3; }
  else if (a == cudaDevAttrMaxSharedMemoryPerBlockOptin) *v = 227 << 10;
  else *v = 0;
  return cudaSuccess;
}
int cusimOccupancy() { const char* e = std::getenv("CUSIM_OCCUPANCY"); return e ? std::atoi(e) : 2; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  std::lock_guard<std::mutex> l(gRegMutex);
  a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = (void*)p;
  auto it = gRegions.upper_bound((uintptr_t)p);
  if (it != gRegions.begin()) {
    --it;
    if ((uintptr_t)p < it->first + std::max<size_t>(it->second.first, 1)) { a->type = (cudaMemoryType)it->second.second; a->devicePointer = (void*)p; }
  }
  return cudaSuccess;
}

// test hooks: "device" buffers for the ctypes harness (device-pointer code paths of the C ABI)
extern "C" __attribute__((visibility("default"))) void* cusim_device_alloc(size_t n) { void* p = nullptr; cusimMalloc(&p, n); return p; }
extern "C" __attribute__((visibility("default"))) void cusim_device_free(void* p) { cudaFree(p); }
