// tools/cusim/cub/device/device_scan.cuh -- DEVELOPMENT TOOL, NOT PRODUCT: stands in for CUB's DeviceScan in the sim build.
#pragma once
#include <cuda_runtime.h>
namespace cub {
struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void* tmp, size_t& tmpBytes, In in, Out out, int n, cudaStream_t = nullptr) {
    if (!tmp) { tmpBytes = 16; return cudaSuccess; }
    using V = std::remove_reference_t<decltype(out[0])>;
    V run = 0;
    for (int i = 0; i < n; i++) { V v = (V)in[i]; out[i] = run; run = (V)(run + v); }
    return cudaSuccess;
  }
  template <class In, class Out>
  static cudaError_t InclusiveSum(void* tmp, size_t& tmpBytes, In in, Out out, int n, cudaStream_t = nullptr) {
    if (!tmp) { tmpBytes = 16; return cudaSuccess; }
    using V = std::remove_reference_t<decltype(out[0])>;
    V run = 0;
    for (int i = 0; i < n; i++) { run = (V)(run + (V)in[i]); out[i] = run; }
    return cudaSuccess;
  }
  template <class In, class Out, class Op>
  static cudaError_t InclusiveScan(void* tmp, size_t& tmpBytes, In in, Out out, Op op, int n, cudaStream_t = nullptr) {
    if (!tmp) { tmpBytes = 16; return cudaSuccess; }
    using V = std::remove_reference_t<decltype(out[0])>;
    V run{};
    for (int i = 0; i < n; i++) { run = i ? (V)op(run, (V)in[i]) : (V)in[i]; out[i] = run; }
    return cudaSuccess;
  }
};
}  // namespace cub
