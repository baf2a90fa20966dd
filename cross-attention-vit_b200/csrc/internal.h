// Host-side internals shared by the translation units of libcavit_sm100a.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cavit.h"

namespace cavit {

// thread-local last-error message; returns `code` so callers can `return fail(...)`.
int fail(int code, const char* fmt, ...);
// device status word (kernel-side time-outs) of the CURRENT device (one per device, cached per device id).
int* status_word();
void count_launch(int n = 1);
// Checks cudaGetLastError() after a launch.
int check_launch(const char* what);

// 3-D bf16 tensor map {inner, rows, groups} with 128-byte swizzle. Cached by value.
// Returns nullptr on failure (fail() already called).
const CUtensorMap* tensor_map_bf16_3d(const void* base, uint64_t inner, uint64_t rows, uint64_t groups,
                                      uint64_t row_stride_elems, uint64_t group_stride_elems,
                                      uint32_t box_inner, uint32_t box_rows);
// 4-D bf16 tensor map {d0, d1, d2, d3} (strides in elements for dims 1..3), 128-byte swizzle.
const CUtensorMap* tensor_map_bf16_4d(const void* base, const uint64_t dims[4], const uint64_t strides_elems[3],
                                      const uint32_t box[4]);

// Generic tiled map: rank <= 5, fp32 (dtype_f32 = 1) or bf16 elements, swizzle_bytes 0 (none) / 32 / 64 / 128 (1 = 128),
// zero fill out of bounds.
// strides_bytes holds the strides of dims 1 .. rank-1. Cached by value; nullptr on failure.
const CUtensorMap* tensor_map_nd(int dtype_f32, uint32_t rank, const void* base, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// p * 2^32 threshold for the counter-based dropout masks (common.cuh); p in [0, 1)
inline uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}
int sm_count();

// "done once" flag per CUDA device: kernel function attributes (dynamic shared memory limits) belong to a device, so a
// process that uses a second GPU has to set them again there.
struct PerDeviceFlag {
  unsigned char f[64] = {};
  static int dev() { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : -1; }
  bool unset() const { const int d = dev(); return d < 0 || !f[d]; }
  void set() { const int d = dev(); if (d >= 0) f[d] = 1; }
};
// largest dynamic shared-memory size a kernel has been opted into on the current device
struct PerDeviceMax {
  size_t v[64] = {};
  size_t get() const { const int d = PerDeviceFlag::dev(); return d < 0 ? 0 : v[d]; }
  void set(size_t x) { const int d = PerDeviceFlag::dev(); if (d >= 0) v[d] = x; }
};

// Short-sequence (N <= 256) attention backward: one persistent CTA per SM walks whole heads (attn_short.cu).
int launch_attn_bwd_short(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int G, int B,
                          int N, int H, float scale, cudaStream_t stream);

}  // namespace cavit
