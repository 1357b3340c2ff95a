// lerc_context.cpp -- CUDA plumbing behind the stateless C API: a pool of contexts (one stream, one
// growable device arena, pinned staging) so that concurrent callers never share scratch memory
// (the reference is re-entrant with all state on the caller's stack, SURVEY.md 8b "Threading").
#include "lerc_internal.h"
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <map>
#include <string>
#include <cstring>

namespace lerc {

bool cudaOk(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  if (std::getenv("LERC_B200_VERBOSE"))
    std::fprintf(stderr, "[lerc_b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
  cudaGetLastError();   // clear the sticky-free error state
  return false;
}

// ---- arena ----------------------------------------------------------------------------------
void* Arena::alloc(size_t bytes, size_t align) {
#ifdef LERC_CUSIM
  // simulator builds only (tools/cusim): with CUSIM_GUARD_ARENA every scratch allocation is its own heap block, so that an
  // AddressSanitizer build puts redzones around each of them
  if (std::getenv("CUSIM_GUARD_ARENA")) {
    uint8_t* p = nullptr;
    if (!cudaOk(cudaMalloc(&p, ((bytes ? bytes : 1) + 15) / 16 * 16), "cudaMalloc(guarded scratch)")) return nullptr;   // whole 16-byte chunks, like the real arena
    retired.push_back(p);
    return p;
  }
#endif
  size_t off = (used + align - 1) / align * align;
  if (off + bytes > cap) {
    // Grow by replacing the block.  Pointers handed out earlier in this call stay valid because the
    // old block is only retired (freed at releaseContext), never reused.
    size_t want = cap ? cap * 2 : ((size_t)64 << 20);
    while (want < bytes + align) want *= 2;
    uint8_t* nb = nullptr;
    if (!cudaOk(cudaMalloc(&nb, want), "cudaMalloc(arena)")) return nullptr;
    if (base) retired.push_back(base);
    base = nb; cap = want; used = 0;
    off = 0;
  }
  used = off + bytes;
  return base + off;
}

void Arena::reset() {
  used = 0;
  for (uint8_t* p : retired) cudaFree(p);
  retired.clear();
}

void Context::forkSide() {
  cudaEventRecord(evFork, stream);
  cudaStreamWaitEvent(side, evFork, 0);
  mainSaved = stream; stream = side;
}
void Context::backToMain() {
  cudaEventRecord(evJoin, side);
  stream = mainSaved; mainSaved = nullptr;
  sidePending = true;
}
void Context::joinSide() {
  if (sidePending) { cudaStreamWaitEvent(stream, evJoin, 0); sidePending = false; }
}

bool Context::pipeStreams() {
  if (copyIn) return true;
  if (!cudaOk(cudaStreamCreateWithFlags(&copyIn, cudaStreamNonBlocking), "cudaStreamCreate") ||
      !cudaOk(cudaStreamCreateWithFlags(&copyOut, cudaStreamNonBlocking), "cudaStreamCreate")) { copyIn = copyOut = nullptr; return false; }
  for (int d = 0; d < 2; d++)
    for (int i = 0; i < 16; i++)
      if (!cudaOk(cudaEventCreateWithFlags(&evStrip[d][i], cudaEventDisableTiming), "cudaEventCreate")) return false;
  return true;
}

void* Context::pinnedAlloc(size_t bytes) {
  size_t off = (pinnedUsed + 63) / 64 * 64;
  if (off + bytes > pinnedCap) return nullptr;
  pinnedUsed = off + bytes;
  return pinned + off;
}

// ---- pool -----------------------------------------------------------------------------------
namespace {
std::mutex gMutex;
std::vector<Context*> gFree;
Stats gStats = {0, 0, 0, 0, 0};
bool gNoDevice = false;
}  // namespace

Stats& globalStats() { return gStats; }

bool gProfileEnabled = false;
namespace { std::map<std::string, std::pair<double, unsigned long long>> gProfile; }

cudaEvent_t Context::takeEvent() {
  if (!eventPool.empty()) { cudaEvent_t e = eventPool.back(); eventPool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

void profileReport(char* buf, int bufLen, bool reset) {
  std::lock_guard<std::mutex> lock(gMutex);
  std::string out;
  for (auto& kv : gProfile) { char line[256]; std::snprintf(line, sizeof line, "%s\t%llu\t%.6f\n", kv.first.c_str(), kv.second.second, kv.second.first); out += line; }
  if (buf && bufLen > 0) { std::strncpy(buf, out.c_str(), (size_t)bufLen - 1); buf[bufLen - 1] = 0; }
  if (reset) gProfile.clear();
}

Context* acquireContext() {
  int dev = 0;
  const cudaError_t devErr = cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lock(gMutex);
    if (gNoDevice) return nullptr;
    if (devErr == cudaSuccess)                      // contexts are bound to the device that was current when they were created
      for (size_t i = gFree.size(); i-- > 0;)
        if (gFree[i]->device == dev) { Context* c = gFree[i]; gFree.erase(gFree.begin() + (long)i); return c; }
  }
  if (devErr != cudaSuccess) {
    // No usable CUDA device: the product has no CPU fallback by design; every call fails loudly.
    std::lock_guard<std::mutex> lock(gMutex);
    gNoDevice = true;
    std::fprintf(stderr, "[lerc_b200] no CUDA device available: this library has no CPU fallback\n");
    cudaGetLastError();
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    std::lock_guard<std::mutex> lock(gMutex);
    gNoDevice = true;
    std::fprintf(stderr, "[lerc_b200] no CUDA device available: this library has no CPU fallback\n");
    cudaGetLastError();
    return nullptr;
  }
  Context* c = new Context();
  c->device = dev;
  if (!cudaOk(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete c; return nullptr; }
  if (!cudaOk(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking), "cudaStreamCreate") ||
      !cudaOk(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming), "cudaEventCreate") ||
      !cudaOk(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming), "cudaEventCreate")) { delete c; return nullptr; }
  c->pinnedCap = (size_t)1 << 20;
  if (!cudaOk(cudaMallocHost(&c->pinned, c->pinnedCap), "cudaMallocHost")) { cudaStreamDestroy(c->stream); delete c; return nullptr; }
  return c;
}

void releaseContext(Context* c) {
  if (!c) return;
  if (c->sidePending) { cudaStreamSynchronize(c->side); c->sidePending = false; }
  // early error returns can leave kernels / copies in flight that use the arena or the pinned staging: drain the call's stream before
  // they go back to the pool (free on the success paths, which have synchronised already)
  if (c->drainOnRelease) {
    cudaStreamSynchronize(c->stream);
    if (c->copyIn) { cudaStreamSynchronize(c->copyIn); cudaStreamSynchronize(c->copyOut); }     // (strip copies of a pipelined host-resident call)
  }
  c->drainOnRelease = true;
  c->arena.reset();
  c->pinnedUsed = 0;
  std::lock_guard<std::mutex> lock(gMutex);
  if (!c->profRecs.empty()) {
    cudaEventSynchronize(c->profRecs.back().b);
    for (auto& r : c->profRecs) {
      float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
      auto& slot = gProfile[r.name]; slot.first += ms; slot.second += 1;
      c->eventPool.push_back(r.a); c->eventPool.push_back(r.b);
    }
    c->profRecs.clear();
  }
  gStats.kernelLaunches += c->kernelLaunches;
  c->kernelLaunches = 0;
  gFree.push_back(c);
}

PtrKind classifyPointer(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return PTR_HOST_PAGEABLE; }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return PTR_DEVICE;
  if (attr.type == cudaMemoryTypeHost) return PTR_HOST_PINNED;
  return PTR_HOST_PAGEABLE;
}

}  // namespace lerc
