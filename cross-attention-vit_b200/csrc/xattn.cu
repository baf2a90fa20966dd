// cavit-sm100 — K-XATTN: single-query cross attention (L_q = 1), head_dim 64, HBM-bound.
//
// CrossAttention (model_cross.py:88-99) projects ONE query (the CLS token) per sample, so the
// score tensor is [B, H, 1, N]: this is a memory-bound pass over K and V (each read exactly once),
// not a tensor-core problem. One CTA per (fusion, sample, head):
//   phase 1  one thread per key row: 64-wide dot with the query (8 x 16-byte loads of a full 128 B
//            line), scores kept in shared memory, block max / sum by warp shuffles;
//   phase 2  softmax probabilities written out (saved for backward);
//   phase 3  out = sum_n p_n V[n]: each warp streams whole 128-byte V rows (coalesced), lanes own 2
//            channels, partial sums combined across warps through shared memory.
// Algorithmic bytes per (sample, head): 2 * N * 64 * 2 (K and V read once) + 4N (probs).
#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int XA_THREADS = 128;
constexpr int XA_D = 64;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < XA_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__device__ __forceinline__ float dot64(const float* qs, const bf16* row) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < XA_D; i += 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + i));
    float2 f;
    f = unpack_bf16(v.x); s += qs[i] * f.x + qs[i + 1] * f.y;
    f = unpack_bf16(v.y); s += qs[i + 2] * f.x + qs[i + 3] * f.y;
    f = unpack_bf16(v.z); s += qs[i + 4] * f.x + qs[i + 5] * f.y;
    f = unpack_bf16(v.w); s += qs[i + 6] * f.x + qs[i + 7] * f.y;
  }
  return s;
}

// grid = (H, B, K)
__global__ void __launch_bounds__(XA_THREADS)
xattn_fwd_kernel(const float* __restrict__ q, const bf16* __restrict__ kv, float* __restrict__ out, float* __restrict__ probs,
                 int B, int N, int H, float scale, DropCfg dcfg, int use_drop) {
  extern __shared__ float sm[];
  float* s_p = sm;           // [N]
  float* s_q = sm + N;       // [64]
  float* s_red = s_q + XA_D; // [4]
  float* s_acc = s_red + 4;  // [4][64]
  const int h = blockIdx.x, b = blockIdx.y, k = blockIdx.z;
  const int C = H * XA_D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* kvb = kv + ((long long)k * B + b) * (long long)N * 2 * C;
  if (tid < XA_D) s_q[tid] = q[((long long)k * B + b) * C + h * XA_D + tid] * scale;
  __syncthreads();
  float qs[XA_D];
#pragma unroll
  for (int i = 0; i < XA_D; ++i) qs[i] = s_q[i];
  float mx = -INFINITY;
  for (int n = tid; n < N; n += XA_THREADS) {
    const float s = dot64(qs, kvb + (long long)n * 2 * C + h * XA_D);
    s_p[n] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce(mx, s_red, true);
  float sum = 0.f;
  for (int n = tid; n < N; n += XA_THREADS) {
    const float e = __expf(s_p[n] - mx);
    s_p[n] = e;
    sum += e;
  }
  sum = block_reduce(sum, s_red, false);
  const float inv = 1.0f / sum;
  float* pr = probs + (((long long)k * B + b) * H + h) * N;
  const unsigned long long seed = use_drop ? *dcfg.seed : 0ull;
  const unsigned long long pbase = (((unsigned long long)k * B + b) * H + h) * (unsigned long long)N;
  for (int n = tid; n < N; n += XA_THREADS) {
    const float pn = s_p[n] * inv;
    pr[n] = pn;  // saved pre-dropout (softmax backward needs it); attn_drop (model_cross.py:97) applies to the weights used below
    s_p[n] = use_drop ? pn * drop_mult(dcfg, seed, pbase + n) : pn;
  }
  __syncthreads();
  float a0 = 0.f, a1 = 0.f;
  const bf16* vb = kvb + C + h * XA_D + lane * 2;
  for (int n = warp; n < N; n += XA_THREADS / 32) {
    const float2 v = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(vb + (long long)n * 2 * C)));
    const float pn = s_p[n];
    a0 += pn * v.x;
    a1 += pn * v.y;
  }
  s_acc[warp * XA_D + lane * 2] = a0;
  s_acc[warp * XA_D + lane * 2 + 1] = a1;
  __syncthreads();
  if (tid < XA_D) {
    float s = 0.f;
    for (int w = 0; w < XA_THREADS / 32; ++w) s += s_acc[w * XA_D + tid];
    out[((long long)k * B + b) * C + h * XA_D + tid] = s;
  }
}

// grid = (H, B, K)
__global__ void __launch_bounds__(XA_THREADS)
xattn_bwd_kernel(const float* __restrict__ q, const bf16* __restrict__ kv, const float* __restrict__ probs,
                 const float* __restrict__ dout, float* __restrict__ dq, bf16* __restrict__ dkv, int B, int N, int H,
                 float scale, DropCfg dcfg, int use_drop) {
  extern __shared__ float sm[];
  float* s_ds = sm;           // [N]
  float* s_q = sm + N;        // [64] (unscaled q)
  float* s_do = s_q + XA_D;   // [64]
  float* s_red = s_do + XA_D; // [4]
  float* s_acc = s_red + 4;   // [4][64]
  const int h = blockIdx.x, b = blockIdx.y, k = blockIdx.z;
  const int C = H * XA_D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long kb = (long long)k * B + b;
  const bf16* kvb = kv + kb * (long long)N * 2 * C;
  bf16* dkvb = dkv + kb * (long long)N * 2 * C;
  const float* pr = probs + (kb * H + h) * N;
  if (tid < XA_D) {
    s_q[tid] = q[kb * C + h * XA_D + tid];
    s_do[tid] = dout[kb * C + h * XA_D + tid];
  }
  __syncthreads();
  float dov[XA_D];
#pragma unroll
  for (int i = 0; i < XA_D; ++i) dov[i] = s_do[i];
  // dp_n = m_n <dout, V[n]> (m_n = attention-dropout multiplier);  dot_pd = sum_n p_n dp_n
  const unsigned long long seed = use_drop ? *dcfg.seed : 0ull;
  const unsigned long long pbase = (kb * H + h) * (unsigned long long)N;
  float part = 0.f;
  for (int n = tid; n < N; n += XA_THREADS) {
    float dp = dot64(dov, kvb + (long long)n * 2 * C + C + h * XA_D);
    if (use_drop) dp *= drop_mult(dcfg, seed, pbase + n);
    s_ds[n] = dp;
    part += pr[n] * dp;
  }
  const float dot_pd = block_reduce(part, s_red, false);
  // ds_n = p_n (dp_n - dot_pd); dK[n] = scale ds_n q; dV[n] = p_n dout
  for (int n = tid; n < N; n += XA_THREADS) {
    const float pn = pr[n];
    const float ds = pn * (s_ds[n] - dot_pd);
    s_ds[n] = ds;
    const float pv = use_drop ? pn * drop_mult(dcfg, seed, pbase + n) : pn;  // dropped weight that multiplied V[n]
    bf16* dk = dkvb + (long long)n * 2 * C + h * XA_D;
    bf16* dv = dk + C;
    const float dss = ds * scale;
#pragma unroll
    for (int i = 0; i < XA_D; i += 8) {
      uint4 w;
      w.x = pack_bf16(dss * s_q[i], dss * s_q[i + 1]);
      w.y = pack_bf16(dss * s_q[i + 2], dss * s_q[i + 3]);
      w.z = pack_bf16(dss * s_q[i + 4], dss * s_q[i + 5]);
      w.w = pack_bf16(dss * s_q[i + 6], dss * s_q[i + 7]);
      *reinterpret_cast<uint4*>(dk + i) = w;
      w.x = pack_bf16(pv * dov[i], pv * dov[i + 1]);
      w.y = pack_bf16(pv * dov[i + 2], pv * dov[i + 3]);
      w.z = pack_bf16(pv * dov[i + 4], pv * dov[i + 5]);
      w.w = pack_bf16(pv * dov[i + 6], pv * dov[i + 7]);
      *reinterpret_cast<uint4*>(dv + i) = w;
    }
  }
  __syncthreads();
  // dq[d] = scale * sum_n ds_n K[n][d]
  float a0 = 0.f, a1 = 0.f;
  const bf16* kbp = kvb + h * XA_D + lane * 2;
  for (int n = warp; n < N; n += XA_THREADS / 32) {
    const float2 v = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(kbp + (long long)n * 2 * C)));
    const float ds = s_ds[n];
    a0 += ds * v.x;
    a1 += ds * v.y;
  }
  s_acc[warp * XA_D + lane * 2] = a0;
  s_acc[warp * XA_D + lane * 2 + 1] = a1;
  __syncthreads();
  if (tid < XA_D) {
    float s = 0.f;
    for (int w = 0; w < XA_THREADS / 32; ++w) s += s_acc[w * XA_D + tid];
    dq[kb * C + h * XA_D + tid] = s * scale;
  }
}

}  // namespace cavit

using namespace cavit;

extern "C" {

static int xa_drop(float p, const uint64_t* seed_dev, uint32_t site, DropCfg* d) {
  d->seed = reinterpret_cast<const unsigned long long*>(seed_dev);
  d->site = site;
  d->thresh = 0;
  d->inv_keep = 1.f;
  if (p <= 0.f) return 0;
  if (p >= 1.f || !seed_dev) return -1;
  d->thresh = drop_threshold(p);
  d->inv_keep = 1.0f / (1.0f - p);
  return 1;
}

int cavit_xattn_fwd(const float* q, const void* kv, float* out, float* probs, int32_t K, int32_t B, int32_t N, int32_t H,
                    float scale, float p_drop, const uint64_t* seed_dev, uint32_t site, void* stream) {
  if (!q || !kv || !out || !probs) return fail(CAVIT_E_BADARG, "cavit_xattn_fwd: null pointer");
  if (K <= 0 || B <= 0 || N <= 0 || H <= 0) return fail(CAVIT_E_BADARG, "cavit_xattn_fwd: bad extents");
  const size_t smem = sizeof(float) * ((size_t)N + XA_D + 4 + 4 * XA_D);
  if (smem > 200 * 1024) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_xattn_fwd: N=%d too long for one CTA", N);
  static PerDeviceMax cur;
  if (smem > 48 * 1024 && smem > cur.get()) {
    cudaFuncSetAttribute(xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cur.set(smem);
  }
  DropCfg d;
  const int use = xa_drop(p_drop, seed_dev, site, &d);
  if (use < 0) return fail(CAVIT_E_BADARG, "cavit_xattn_fwd: bad dropout arguments");
  xattn_fwd_kernel<<<dim3(H, B, K), XA_THREADS, smem, as_stream(stream)>>>(q, reinterpret_cast<const bf16*>(kv), out, probs,
                                                                           B, N, H, scale, d, use);
  count_launch();
  return check_launch("cavit_xattn_fwd");
}

int cavit_xattn_bwd(const float* q, const void* kv, const float* probs, const float* dout, float* dq, void* dkv, int32_t K,
                    int32_t B, int32_t N, int32_t H, float scale, float p_drop, const uint64_t* seed_dev, uint32_t site,
                    void* stream) {
  if (!q || !kv || !probs || !dout || !dq || !dkv) return fail(CAVIT_E_BADARG, "cavit_xattn_bwd: null pointer");
  if (K <= 0 || B <= 0 || N <= 0 || H <= 0) return fail(CAVIT_E_BADARG, "cavit_xattn_bwd: bad extents");
  const size_t smem = sizeof(float) * ((size_t)N + 2 * XA_D + 4 + 4 * XA_D);
  if (smem > 200 * 1024) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_xattn_bwd: N=%d too long for one CTA", N);
  static PerDeviceMax cur;
  if (smem > 48 * 1024 && smem > cur.get()) {
    cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cur.set(smem);
  }
  DropCfg d;
  const int use = xa_drop(p_drop, seed_dev, site, &d);
  if (use < 0) return fail(CAVIT_E_BADARG, "cavit_xattn_bwd: bad dropout arguments");
  xattn_bwd_kernel<<<dim3(H, B, K), XA_THREADS, smem, as_stream(stream)>>>(q, reinterpret_cast<const bf16*>(kv), probs, dout,
                                                                           dq, reinterpret_cast<bf16*>(dkv), B, N, H, scale, d, use);
  count_launch();
  return check_launch("cavit_xattn_bwd");
}

}  // extern "C"
