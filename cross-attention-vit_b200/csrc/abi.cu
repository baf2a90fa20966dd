// cavit-sm100 — library plumbing: error reporting, device status word, TMA descriptor cache.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "internal.h"

namespace cavit {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return CAVIT_OK;
}

__device__ int g_status_word[4];

// Both are properties of the CURRENT device (a __device__ symbol has one instance per device): cached per device id so
// that a process driving several GPUs (model on cuda:1 while cuda:0 is also in use) never gets another device's address.
constexpr int kMaxDevices = 64;

int* status_word() {
  static std::atomic<int*> cache[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  int* p = cache[dev].load(std::memory_order_acquire);
  if (!p) {
    void* q = nullptr;
    if (cudaGetSymbolAddress(&q, g_status_word) != cudaSuccess) return nullptr;
    p = static_cast<int*>(q);
    cache[dev].store(p, std::memory_order_release);
  }
  return p;
}

int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[20];
  bool operator<(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::mutex g_map_mu;
static std::map<MapKey, CUtensorMap*> g_maps;

static const CUtensorMap* lookup_or_encode(const MapKey& key, uint32_t rank, const void* base,
                                           const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                                           const cuuint32_t* box) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    fail(CAVIT_E_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  // 64-byte aligned storage that is never freed (descriptors are referenced by in-flight launches).
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) {
    fail(CAVIT_E_DEVICE, "posix_memalign failed");
    return nullptr;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fail(CAVIT_E_BADARG,
         "cuTensorMapEncodeTiled failed (%d): rank %u base %p dims [%llu,%llu,%llu] strideB [%llu,%llu] box [%u,%u,%u]",
         (int)r, rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
         (unsigned long long)dims[2], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
         box[0], box[1], box[2]);
    free(m);
    return nullptr;
  }
  if (g_maps.size() > 65536) g_maps.clear();  // descriptors stay alive (leaked) — bounded by churn
  g_maps[key] = m;
  return m;
}

// Generic tiled map (any rank <= 5, fp32 or bf16, any swizzle; zero fill outside the tensor). strides_bytes: dims 1..rank-1.
const CUtensorMap* tensor_map_nd(int dtype_f32, uint32_t rank, const void* base, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  if (rank < 1 || rank > 5) {
    fail(CAVIT_E_BADARG, "tensor_map_nd: rank %u", rank);
    return nullptr;
  }
  MapKey key{};
  key.v[0] = 100 + rank;
  key.v[1] = reinterpret_cast<uint64_t>(base);
  key.v[2] = (uint64_t)dtype_f32 | ((uint64_t)swizzle_bytes << 8);
  for (uint32_t i = 0; i < rank; ++i) key.v[3 + i] = dims[i];
  for (uint32_t i = 0; i + 1 < rank; ++i) key.v[8 + i] = strides_bytes[i];
  for (uint32_t i = 0; i < rank; ++i) key.v[12 + i] = box[i];
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    fail(CAVIT_E_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) {
    fail(CAVIT_E_DEVICE, "posix_memalign failed");
    return nullptr;
  }
  cuuint64_t d[5] = {1, 1, 1, 1, 1}, st[4] = {0, 0, 0, 0};
  cuuint32_t bx[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  for (uint32_t i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (uint32_t i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  // swizzle_bytes: 0 = none, 1 or 128 = SWIZZLE_128B, 64 = SWIZZLE_64B, 32 = SWIZZLE_32B
  const CUtensorMapSwizzle swz = (swizzle_bytes == 1 || swizzle_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(m, dtype_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                  const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fail(CAVIT_E_BADARG,
         "cuTensorMapEncodeTiled failed (%d): rank %u base %p dims [%llu,%llu,%llu,%llu,%llu] strideB [%llu,%llu,%llu,%llu] "
         "box [%u,%u,%u,%u,%u]", (int)r, rank, base, (unsigned long long)d[0], (unsigned long long)d[1],
         (unsigned long long)d[2], (unsigned long long)d[3], (unsigned long long)d[4], (unsigned long long)st[0],
         (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)st[3], bx[0], bx[1], bx[2], bx[3], bx[4]);
    free(m);
    return nullptr;
  }
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps[key] = m;
  return m;
}

const CUtensorMap* tensor_map_bf16_3d(const void* base, uint64_t inner, uint64_t rows, uint64_t groups,
                                      uint64_t row_stride_elems, uint64_t group_stride_elems, uint32_t box_inner,
                                      uint32_t box_rows) {
  MapKey key{};
  key.v[0] = 3;
  key.v[1] = reinterpret_cast<uint64_t>(base);
  key.v[2] = inner; key.v[3] = rows; key.v[4] = groups;
  key.v[5] = row_stride_elems; key.v[6] = group_stride_elems;
  key.v[7] = box_inner; key.v[8] = box_rows;
  cuuint64_t dims[3] = {inner, rows, groups};
  if (groups <= 1 || group_stride_elems == 0) group_stride_elems = rows * row_stride_elems;
  cuuint64_t strides[2] = {row_stride_elems * 2, group_stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  return lookup_or_encode(key, 3, base, dims, strides, box);
}

const CUtensorMap* tensor_map_bf16_4d(const void* base, const uint64_t dims_[4], const uint64_t strides_elems[3],
                                      const uint32_t box_[4]) {
  MapKey key{};
  key.v[0] = 4;
  key.v[1] = reinterpret_cast<uint64_t>(base);
  for (int i = 0; i < 4; ++i) key.v[2 + i] = dims_[i];
  for (int i = 0; i < 3; ++i) key.v[6 + i] = strides_elems[i];
  for (int i = 0; i < 4; ++i) key.v[9 + i] = box_[i];
  cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
  cuuint64_t strides[3] = {strides_elems[0] * 2, strides_elems[1] * 2, strides_elems[2] * 2};
  cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
  return lookup_or_encode(key, 4, base, dims, strides, box);
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_abi_version(void) { return CAVIT_ABI_VERSION; }
const char* cavit_last_error(void) { return g_err; }

int cavit_device_ok(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

int cavit_device_status(int reset) {
  int* p = status_word();
  if (!p) return fail(CAVIT_E_DEVICE, "status word unavailable");
  int h[4] = {0, 0, 0, 0};
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(CAVIT_E_DEVICE, "device error: %s", cudaGetErrorString(e));
  e = cudaMemcpy(h, p, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(CAVIT_E_DEVICE, "status read: %s", cudaGetErrorString(e));
  if (reset && h[0] != 0) cudaMemset(p, 0, sizeof(h));
  if (h[0] != 0) fail(h[0], "kernel-side time-out code %d", h[0]);
  return h[0];
}

long long cavit_launch_count(void) { return g_launches.load(); }

}  // extern "C"
