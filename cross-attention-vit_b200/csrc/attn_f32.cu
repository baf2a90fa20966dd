// cavit-sm100 — K-ATTN for the fp32-tolerance mode: fused online-softmax self-attention, forward and backward, in fp32
// on the CUDA cores (head_dim 64).
//
// The reference trains in fp32 (L.Trainer without precision=, /root/reference/main_mist.py:211-218) and north_star asks
// for ~1e-3 on logits and attention outputs in that mode. The tcgen05 attention kernels (attn.cu, attn_short.cu) round
// Q, K, V and the probabilities to bf16 (8 mantissa bits): good for 2e-2, not for 1e-3. In the fp32 mode the projections
// run as 3-term bf16 split products on the tensor cores (gemm.cu) and attention — 7 % of the FLOPs at cfg2 — runs here:
// fp32 operands straight from the fp32 packed QKV activation, fp32 scores / softmax / accumulation, expf from libdevice.
// Like the tensor-core kernels it never writes the N x N score matrix anywhere: scores live in registers.
//
// Mapping: two threads per row (each owns 32 of the 64 head-dim values of its row in registers; the two partial dot
// products are combined with one shuffle), 64 rows per 128-thread CTA, the other side's rows staged through shared memory
// in tiles of 64 and read as warp-wide broadcasts:
//   forward   row = query:  s_j = q . k_j, online softmax (8 keys at a time), o += p_j v_j
//   backward  dq kernel     row = query:  p_j = exp(scale s_j - lse), ds_j = p_j (do . v_j - delta), dq += ds_j k_j;
//                           also writes delta = rowsum(dO o O) for the second kernel
//             dkdv kernel   row = key:    dv += p_i do_i, dk += ds_i q_i over all queries
// (S and dP are recomputed by both backward kernels: 7 instead of 5 tile products, no atomics, deterministic.)
// Replaces `matmul / softmax / matmul` of Attention.forward and their autograd (/root/reference/model_cross.py:50-61).
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int AF_ROWS = 64;      // rows per CTA (two threads each)
constexpr int AF_THREADS = 128;
constexpr int AF_TILE = 64;      // staged rows of the other side per step

struct AttnF32Params {
  const float* qkv;    // [G][B*N][3C]  (q | k | v thirds, each (h d))
  const float* o;      // [G][B*N][C]   (backward)
  const float* d_o;    // [G][B*N][C]   (backward)
  float* out;          // forward: [G][B*N][C]; backward: dqkv [G][B*N][3C]
  float* lse;          // [G][B][H][N]  natural-log LSE of the scaled scores (written by fwd, read by bwd)
  float* delta;        // [G][B][H][N]  (backward)
  int B, N, H;
  float scale;
};

// dst[r][0..63] = src[(row0 + r) * ld + 0..63] for r < AF_TILE (zeros past `rows`); 128 threads, 16 per row.
__device__ __forceinline__ void stage_tile(float (*dst)[AF_TILE], const float* src, long long ld, int row0, int rows) {
  const int c4 = threadIdx.x & 15, r0 = threadIdx.x >> 4;
#pragma unroll
  for (int it = 0; it < AF_TILE / 8; ++it) {
    const int r = r0 + it * 8;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows) v = __ldg(reinterpret_cast<const float4*>(src + (long long)(row0 + r) * ld) + c4);
    reinterpret_cast<float4*>(dst[r])[c4] = v;
  }
}

// partial dot product of this thread's 32 values with row[half*32 .. +32) (shared memory, broadcast reads), 4 chains
__device__ __forceinline__ float dot32(const float (&a)[32], const float* row) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float4 k = reinterpret_cast<const float4*>(row)[t];
    s0 = fmaf(a[4 * t + 0], k.x, s0);
    s1 = fmaf(a[4 * t + 1], k.y, s1);
    s2 = fmaf(a[4 * t + 2], k.z, s2);
    s3 = fmaf(a[4 * t + 3], k.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void axpy32(float (&acc)[32], float w, const float* row) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float4 v = reinterpret_cast<const float4*>(row)[t];
    acc[4 * t + 0] = fmaf(w, v.x, acc[4 * t + 0]);
    acc[4 * t + 1] = fmaf(w, v.y, acc[4 * t + 1]);
    acc[4 * t + 2] = fmaf(w, v.z, acc[4 * t + 2]);
    acc[4 * t + 3] = fmaf(w, v.w, acc[4 * t + 3]);
  }
}
__device__ __forceinline__ void load32(float (&a)[32], const float* src, bool active) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) v = __ldg(reinterpret_cast<const float4*>(src) + t);
    a[4 * t + 0] = v.x; a[4 * t + 1] = v.y; a[4 * t + 2] = v.z; a[4 * t + 3] = v.w;
  }
}
__device__ __forceinline__ void store32(float* dst, const float (&a)[32], float mul) {
#pragma unroll
  for (int t = 0; t < 8; ++t)
    reinterpret_cast<float4*>(dst)[t] = make_float4(a[4 * t] * mul, a[4 * t + 1] * mul, a[4 * t + 2] * mul, a[4 * t + 3] * mul);
}

// grid = (ceil(N / 64), B * H, G)
__global__ void __launch_bounds__(AF_THREADS) attn_fwd_f32_kernel(const AttnF32Params p) {
  __shared__ __align__(16) float Ks[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Vs[AF_TILE][AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int half = threadIdx.x & 1, r = blockIdx.x * AF_ROWS + (threadIdx.x >> 1);
  const bool active = r < N;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  float q[32], o[32];
  load32(q, base + (long long)r * 3 * C + half * 32, active);
#pragma unroll
  for (int t = 0; t < 32; ++t) { q[t] *= p.scale; o[t] = 0.f; }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < N; j0 += AF_TILE) {
    __syncthreads();
    stage_tile(Ks, base + C, 3LL * C, j0, N);
    stage_tile(Vs, base + 2 * C, 3LL * C, j0, N);
    __syncthreads();
    const int jn = min(AF_TILE, N - j0);
    for (int jc = 0; jc < jn; jc += 8) {
      float s[8];
      float mx = m;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float part = dot32(q, &Ks[jc + jj][half * 32]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        s[jj] = (jc + jj < jn) ? part : -INFINITY;
        mx = fmaxf(mx, s[jj]);
      }
      const float corr = expf(m - mx);      // first chunk: exp(-inf) = 0 (key jc is always valid, so mx is finite)
      l *= corr;
#pragma unroll
      for (int t = 0; t < 32; ++t) o[t] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float pj = expf(s[jj] - mx);
        l += pj;
        axpy32(o, pj, &Vs[jc + jj][half * 32]);
      }
      m = mx;
    }
  }
  if (active) {
    store32(p.out + (tok0 + r) * C + h * 64 + half * 32, o, 1.0f / l);
    if (half == 0) p.lse[(((long long)g * p.B + b) * p.H + h) * N + r] = m + logf(l);
  }
}

// grid = (ceil(N / 64), B * H, G): dQ (and delta) of 64 query rows
__global__ void __launch_bounds__(AF_THREADS) attn_bwd_dq_f32_kernel(const AttnF32Params p) {
  __shared__ __align__(16) float Ks[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Vs[AF_TILE][AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int half = threadIdx.x & 1, r = blockIdx.x * AF_ROWS + (threadIdx.x >> 1);
  const bool active = r < N;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  const long long stat = (((long long)g * p.B + b) * p.H + h) * N;
  float q[32], d_o[32], dq[32];
  load32(q, base + (long long)r * 3 * C + half * 32, active);
  load32(d_o, p.d_o + (tok0 + r) * C + h * 64 + half * 32, active);
  float delta;
  {
    load32(dq, p.o + (tok0 + r) * C + h * 64 + half * 32, active);   // dq as scratch for the output row
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 32; ++t) s = fmaf(d_o[t], dq[t], s);
    delta = s + __shfl_xor_sync(0xffffffffu, s, 1);
  }
  const float lse = active ? p.lse[stat + r] : 0.f;
  if (active && half == 0) p.delta[stat + r] = delta;
#pragma unroll
  for (int t = 0; t < 32; ++t) dq[t] = 0.f;
  for (int j0 = 0; j0 < N; j0 += AF_TILE) {
    __syncthreads();
    stage_tile(Ks, base + C, 3LL * C, j0, N);
    stage_tile(Vs, base + 2 * C, 3LL * C, j0, N);
    __syncthreads();
    const int jn = min(AF_TILE, N - j0);
    for (int j = 0; j < jn; ++j) {
      float s = dot32(q, &Ks[j][half * 32]);
      float dp = dot32(d_o, &Vs[j][half * 32]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float pj = expf(fmaf(s, p.scale, -lse));
      axpy32(dq, pj * (dp - delta), &Ks[j][half * 32]);
    }
  }
  if (active) store32(p.out + (tok0 + r) * 3 * C + h * 64 + half * 32, dq, p.scale);
}

// grid = (ceil(N / 64), B * H, G): dK, dV of 64 key rows
__global__ void __launch_bounds__(AF_THREADS) attn_bwd_dkdv_f32_kernel(const AttnF32Params p) {
  __shared__ __align__(16) float Qs[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Ds[AF_TILE][AF_TILE];
  __shared__ float lse_s[AF_TILE], delta_s[AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int half = threadIdx.x & 1, r = blockIdx.x * AF_ROWS + (threadIdx.x >> 1);
  const bool active = r < N;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  const long long stat = (((long long)g * p.B + b) * p.H + h) * N;
  float k[32], v[32], dk[32], dv[32];
  load32(k, base + (long long)r * 3 * C + C + half * 32, active);
  load32(v, base + (long long)r * 3 * C + 2 * C + half * 32, active);
#pragma unroll
  for (int t = 0; t < 32; ++t) { dk[t] = 0.f; dv[t] = 0.f; }
  for (int i0 = 0; i0 < N; i0 += AF_TILE) {
    __syncthreads();
    stage_tile(Qs, base, 3LL * C, i0, N);
    stage_tile(Ds, p.d_o + tok0 * C + h * 64, (long long)C, i0, N);
    if (threadIdx.x < AF_TILE) {   // padded query rows: lse = +inf makes their probabilities exactly 0
      const int i = i0 + threadIdx.x;
      lse_s[threadIdx.x] = (i < N) ? p.lse[stat + i] : INFINITY;
      delta_s[threadIdx.x] = (i < N) ? p.delta[stat + i] : 0.f;
    }
    __syncthreads();
    const int in = min(AF_TILE, N - i0);
    for (int i = 0; i < in; ++i) {
      float s = dot32(k, &Qs[i][half * 32]);
      float dp = dot32(v, &Ds[i][half * 32]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float pi = expf(fmaf(s, p.scale, -lse_s[i]));
      axpy32(dv, pi, &Ds[i][half * 32]);
      axpy32(dk, pi * (dp - delta_s[i]), &Qs[i][half * 32]);
    }
  }
  if (active) {
    float* dst = p.out + (tok0 + r) * 3 * C + h * 64 + half * 32;
    store32(dst + C, dk, p.scale);
    store32(dst + 2 * C, dv, 1.0f);
  }
}

static int check_attn_f32(const char* what, int G, int B, int N, int H) {
  if (G <= 0 || B <= 0 || N <= 0 || H <= 0) return fail(CAVIT_E_BADARG, "%s: non-positive extent", what);
  if ((long long)B * H > 65535 || G > 65535) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "%s: B*H = %lld, G = %d exceed the grid limits", what, (long long)B * H, G);
  return CAVIT_OK;
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_attn_fwd_f32(const float* qkv, float* out, float* lse, int32_t G, int32_t B, int32_t N, int32_t H, float scale,
                       void* stream) {
  if (!qkv || !out || !lse) return fail(CAVIT_E_BADARG, "cavit_attn_fwd_f32: null pointer");
  if (int rc = check_attn_f32("cavit_attn_fwd_f32", G, B, N, H)) return rc;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(CAVIT_E_BADARG, "cavit_attn_fwd_f32: 16-byte aligned buffers expected");
  AttnF32Params p{};
  p.qkv = qkv; p.out = out; p.lse = lse; p.B = B; p.N = N; p.H = H; p.scale = scale;
  attn_fwd_f32_kernel<<<dim3((N + AF_ROWS - 1) / AF_ROWS, B * H, G), AF_THREADS, 0, as_stream(stream)>>>(p);
  count_launch();
  return check_launch("cavit_attn_fwd_f32");
}

int cavit_attn_bwd_f32(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, float* delta,
                       int32_t G, int32_t B, int32_t N, int32_t H, float scale, void* stream) {
  if (!qkv || !out || !dout || !lse || !dqkv || !delta) return fail(CAVIT_E_BADARG, "cavit_attn_bwd_f32: null pointer");
  if (int rc = check_attn_f32("cavit_attn_bwd_f32", G, B, N, H)) return rc;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
       reinterpret_cast<uintptr_t>(dqkv)) & 15)
    return fail(CAVIT_E_BADARG, "cavit_attn_bwd_f32: 16-byte aligned buffers expected");
  AttnF32Params p{};
  p.qkv = qkv; p.o = out; p.d_o = dout; p.out = dqkv; p.lse = const_cast<float*>(lse); p.delta = delta;
  p.B = B; p.N = N; p.H = H; p.scale = scale;
  const dim3 grid((N + AF_ROWS - 1) / AF_ROWS, B * H, G);
  attn_bwd_dq_f32_kernel<<<grid, AF_THREADS, 0, as_stream(stream)>>>(p);      // also writes delta for the next kernel
  attn_bwd_dkdv_f32_kernel<<<grid, AF_THREADS, 0, as_stream(stream)>>>(p);
  count_launch(2);
  return check_launch("cavit_attn_bwd_f32");
}

}  // extern "C"
