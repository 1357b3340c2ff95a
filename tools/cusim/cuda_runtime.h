// tools/cusim/cuda_runtime.h -- DEVELOPMENT TOOL, NOT PRODUCT.
//
// A CUDA-semantics shim that lets g++ compile the product's .cu / .cuh sources unchanged (after the launch-syntax
// rewrite of tools/cusim/prep.py) into tools/cusim/_build/libLerc_sim.so, in which every kernel runs on the host:
// one OS thread per resident CTA, one fiber per CUDA thread, warp collectives and barriers as rendezvous points.
// It exists so that kernel LOGIC (indexing, scans, bit packing, speculation/fallback decisions) can be debugged in
// the GPU-less development container before GPU minutes are spent; it says nothing about performance, memory
// coalescing or the GPU memory model.  The product library (lerc_b200/libLerc.so.4) never includes, links or loads
// anything from this directory and still has no CPU fallback; `-m gpu` tests, smoke() and bench.py never touch it.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <functional>

#define LERC_CUSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define CUSIM_NOINLINE __attribute__((noinline))   // prep.py rewrites __noinline__ (libstdc++ uses that spelling inside attributes)
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static
#define __grid_constant__

// ---- vector types ------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8)))  uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8)))  int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8)))  float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
struct __attribute__((aligned(16))) longlong2 { long long x, y; };
struct ushort2 { unsigned short x, y; };
struct __attribute__((aligned(8))) ushort4 { unsigned short x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline ushort2 make_ushort2(unsigned short x, unsigned short y) { return ushort2{x, y}; }
static inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) { return ushort4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

// ---- runtime (cusim.cpp) -------------------------------------------------------------------------
namespace cusim {
struct ThreadCtx {          // what a CUDA thread sees
  uint3 tid, bid;
  dim3 bdim, gdim;
  int lane, warp, linear;
  void* dynSmem;
};
ThreadCtx& self();
void yield();                                                     // give the other fibers of the CTA a turn
void syncthreads();
void namedBarrier(int id, int nThreads);
void warpExchange(unsigned mask, uint64_t mine, uint64_t out[32], unsigned* present);   // rendezvous of the lanes in mask
void launch(dim3 grid, dim3 block, size_t smem, void* stream, const std::function<void()>& body);
inline void* dynSmem() { return self().dynSmem; }
}  // namespace cusim

#define threadIdx (cusim::self().tid)
#define blockIdx  (cusim::self().bid)
#define blockDim  (cusim::self().bdim)
#define gridDim   (cusim::self().gdim)
#define warpSize  32

// ---- host API ------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
typedef struct CusimStream* cudaStream_t;
typedef struct CusimEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cusimMalloc(void** p, size_t n);
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cusimMalloc((void**)p, n); }
cudaError_t cusimMallocHost(void** p, size_t n);
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cusimMallocHost((void**)p, n); }
cudaError_t cudaFree(void* p);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* dst, int v, size_t n);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
int cusimOccupancy();
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = cusimOccupancy(); return cudaSuccess; }

// ---- device intrinsics ---------------------------------------------------------------------------
static inline void __syncthreads() { cusim::syncthreads(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { cusim::yield(); }

namespace cusim {
template <class T> inline uint64_t toBits(T v) { static_assert(sizeof(T) <= 8, "shuffle of a wide type"); uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T fromBits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T shflFrom(unsigned mask, T v, int src) {
  uint64_t all[32]; unsigned present;
  warpExchange(mask, toBits(v), all, &present);
  if (src < 0 || src > 31 || !((present >> src) & 1)) return v;
  return fromBits<T>(all[src]);
}
}  // namespace cusim

static inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t all[32]; unsigned p; cusim::warpExchange(mask, 0, all, &p); }
static inline unsigned __activemask() { return 0xffffffffu; }   // only meaningful for converged code; the product does not use it
template <class T> inline T __shfl_sync(unsigned mask, T v, int srcLane, int width = 32) {
  const int lane = cusim::self().lane;
  return cusim::shflFrom(mask, v, (lane & ~(width - 1)) | (srcLane & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int laneMask, int width = 32) {
  const int lane = cusim::self().lane, src = lane ^ laneMask;
  return cusim::shflFrom(mask, v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : -1);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = cusim::self().lane, src = lane - (int)delta;
  return cusim::shflFrom(mask, v, src >= (lane & ~(width - 1)) ? src : -1);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = cusim::self().lane, src = lane + (int)delta;
  return cusim::shflFrom(mask, v, src < (lane & ~(width - 1)) + width ? src : -1);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t all[32]; unsigned present, r = 0;
  cusim::warpExchange(mask, pred ? 1 : 0, all, &present);
  for (int i = 0; i < 32; i++) if (((present >> i) & 1) && all[i]) r |= 1u << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
  uint64_t all[32]; unsigned present; int r = 1;
  cusim::warpExchange(mask, pred ? 1 : 0, all, &present);
  for (int i = 0; i < 32; i++) if (((present >> i) & 1) && !all[i]) r = 0;
  return r;
}
#define CUSIM_REDUCE(NAME, T, INIT, OP)                                                     \
  static inline T NAME(unsigned mask, T v) {                                                \
    uint64_t all[32]; unsigned present; T r = INIT;                                         \
    cusim::warpExchange(mask, cusim::toBits(v), all, &present);                             \
    for (int i = 0; i < 32; i++) if ((present >> i) & 1) { T o = cusim::fromBits<T>(all[i]); r = OP; } \
    return r;                                                                               \
  }
CUSIM_REDUCE(__reduce_or_sync, unsigned, 0u, (r | o))
CUSIM_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, (r & o))
CUSIM_REDUCE(__reduce_xor_sync, unsigned, 0u, (r ^ o))
CUSIM_REDUCE(__reduce_add_sync, unsigned, 0u, (r + o))
CUSIM_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (o < r ? o : r))
CUSIM_REDUCE(__reduce_max_sync, unsigned, 0u, (o > r ? o : r))
#undef CUSIM_REDUCE
static inline unsigned __match_any_sync(unsigned mask, unsigned long long v) {
  uint64_t all[32]; unsigned present, r = 0;
  cusim::warpExchange(mask, v, all, &present);
  for (int i = 0; i < 32; i++) if (((present >> i) & 1) && all[i] == v) r |= 1u << i;
  return r;
}

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return (unsigned)(((((unsigned long long)hi << 32) | lo) << s) >> 32); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return (unsigned)((((unsigned long long)hi << 32) | lo) >> s); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { if (s >= 32) return lo; return __funnelshift_l(lo, hi, s); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {   // unsigned 4 x 8-bit dot product + c
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { if (s >= 32) return hi; return __funnelshift_r(lo, hi, s); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const unsigned long long src = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned b = (unsigned)(src >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) b = (b & 0x80) ? 0xff : 0;
    r |= b << (8 * i);
  }
  return r;
}
static inline unsigned __vadd4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) + (b >> (8 * i))) & 0xffu) << (8 * i); return r; }
static inline unsigned __vsub4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) - (b >> (8 * i))) & 0xffu) << (8 * i); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
// fp intrinsics: the sim build is compiled with -ffp-contract=off, so plain operators are round-to-nearest, unfused
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline unsigned __float_as_uint(float v) { return cusim::fromBits<unsigned>(cusim::toBits(v)); }
static inline int __float_as_int(float v) { return cusim::fromBits<int>(cusim::toBits(v)); }
static inline float __uint_as_float(unsigned v) { return cusim::fromBits<float>(v); }
static inline float __int_as_float(int v) { return cusim::fromBits<float>((unsigned)v); }
static inline long long __double_as_longlong(double v) { return cusim::fromBits<long long>(cusim::toBits(v)); }
static inline double __longlong_as_double(long long v) { return cusim::fromBits<double>((uint64_t)v); }
static inline double __hiloint2double(int hi, int lo) { return cusim::fromBits<double>(((uint64_t)(unsigned)hi << 32) | (unsigned)lo); }
static inline int __double2hiint(double v) { return (int)(cusim::toBits(v) >> 32); }
static inline int __double2loint(double v) { return (int)(unsigned)cusim::toBits(v); }
// CUDA's float -> integer casts saturate and send NaN to 0; x86's do not.  The kernels guard the cases that matter to
// the format themselves; these helpers are here for intrinsics that promise the CUDA behaviour.
static inline unsigned __double2uint_rz(double v) { return v != v ? 0u : (v <= 0 ? 0u : (v >= 4294967295.0 ? 0xffffffffu : (unsigned)v)); }
static inline int __double2int_rz(double v) { return v != v ? 0 : (v <= -2147483648.0 ? INT32_MIN : (v >= 2147483647.0 ? INT32_MAX : (int)v)); }
static inline int __float2int_rz(float v) { return __double2int_rz((double)v); }
static inline unsigned __float2uint_rz(float v) { return __double2uint_rz((double)v); }
static inline double __int2double_rn(int v) { return (double)v; }
static inline double __uint2double_rn(unsigned v) { return (double)v; }

// min / max as CUDA's global overloads (mixed signedness follows the usual arithmetic conversions, like CUDA's)
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }

// atomics (shared or global: both are plain host memory here)
template <class T, class U> inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> inline T atomicSub(T* p, U v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> inline T atomicAnd(T* p, U v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> inline T atomicXor(T* p, U v) { return __atomic_fetch_xor(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U, class V> inline T atomicCAS(T* p, U cmp, V v) { T e = (T)cmp; __atomic_compare_exchange_n(p, &e, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return e; }
template <class T, class U> inline T atomicMin(T* p, U v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while ((T)v < old && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T, class U> inline T atomicMax(T* p, U v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while ((T)v > old && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
