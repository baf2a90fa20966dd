// cavit-sm100 — K-ATTN for the fp32-tolerance mode: fused online-softmax self-attention, forward and backward, in fp32
// on the CUDA cores (head_dim 64).
//
// The reference trains in fp32 (L.Trainer without precision=, /root/reference/main_mist.py:211-218) and north_star asks
// for ~1e-3 on logits and attention outputs in that mode. The tcgen05 attention kernels (attn.cu, attn_short.cu) round
// Q, K, V and the probabilities to bf16 (8 mantissa bits): good for 2e-2, not for 1e-3. In the fp32 mode the projections
// run as 3-term bf16 split products on the tensor cores (gemm.cu) and attention — 7 % of the FLOPs at cfg2 — runs here:
// fp32 operands straight from the fp32 packed QKV activation, fp32 scores / softmax / accumulation (exponentials as 2^x on the MUFU, rel. error 2^-22).
// Like the tensor-core kernels it never writes the N x N score matrix anywhere: scores live in registers.
//
// Mapping: a QUAD of threads per group of rows — each thread owns a 16-wide slice of the 64 head-dim values of RQ rows in
// registers (RQ = 4 query rows forward, 2 rows backward), partial dot products are combined inside the quad with two
// shuffles. The other side's rows are staged through shared memory in tiles of 64 and read as broadcasts; every value
// fetched from shared memory feeds RQ (x 2 or 3 in the backward kernels) FMAs, which is what keeps the loops bound by
// the FMA pipe instead of the shared-memory port (the first version, one row per thread pair, read one operand per FMA
// and ran at 13 TFLOP/s).
//   forward   rows = queries:  s_j = q . k_j, online softmax (8 keys at a time), o += p_j v_j
//   backward  dq kernel     rows = queries:  p_j = exp(scale s_j - lse), ds_j = p_j (do . v_j - delta), dq += ds_j k_j;
//                           also writes delta = rowsum(dO o O) for the second kernel
//             dkdv kernel   rows = keys:     dv += p_i do_i, dk += ds_i q_i over all queries
// (S and dP are recomputed by both backward kernels: 7 instead of 5 tile products, no atomics, deterministic.)
// Replaces `matmul / softmax / matmul` of Attention.forward and their autograd (/root/reference/model_cross.py:50-61).
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int AF_THREADS = 128;
constexpr int AF_TILE = 64;      // staged rows of the other side per step
constexpr int AF_RQ_FWD = 4;     // rows per quad, forward  (128 rows per CTA)
constexpr int AF_RQ_BWD = 2;     // rows per quad, backward (64 rows per CTA)

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

// 2^x on the MUFU (rel. error 2^-22): the exponentials work in the log2 domain, scale * log2(e) folded into q / the scores
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float AF_LOG2E = 1.4426950408889634f, AF_LN2 = 0.6931471805599453f;

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ void load16(float (&a)[16], const float* src, bool active) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) v = __ldg(reinterpret_cast<const float4*>(src) + t);
    a[4 * t + 0] = v.x; a[4 * t + 1] = v.y; a[4 * t + 2] = v.z; a[4 * t + 3] = v.w;
  }
}
__device__ __forceinline__ void store16(float* dst, const float (&a)[16], float mul) {
#pragma unroll
  for (int t = 0; t < 4; ++t)
    reinterpret_cast<float4*>(dst)[t] = make_float4(a[4 * t] * mul, a[4 * t + 1] * mul, a[4 * t + 2] * mul, a[4 * t + 3] * mul);
}
// RQ partial dot products of this thread's 16-wide slices with ONE staged row slice (4 broadcast loads, RQ * 16 FMAs)
template <int RQ>
__device__ __forceinline__ void dots16(const float (&a)[RQ][16], const float* row, float (&out)[RQ]) {
  float4 k[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) k[t] = reinterpret_cast<const float4*>(row)[t];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; t += 2) {
      s0 = fmaf(a[i][4 * t + 0], k[t].x, s0); s0 = fmaf(a[i][4 * t + 1], k[t].y, s0);
      s0 = fmaf(a[i][4 * t + 2], k[t].z, s0); s0 = fmaf(a[i][4 * t + 3], k[t].w, s0);
      s1 = fmaf(a[i][4 * t + 4], k[t + 1].x, s1); s1 = fmaf(a[i][4 * t + 5], k[t + 1].y, s1);
      s1 = fmaf(a[i][4 * t + 6], k[t + 1].z, s1); s1 = fmaf(a[i][4 * t + 7], k[t + 1].w, s1);
    }
    out[i] = s0 + s1;
  }
}
// acc[i] += w[i] * row slice for RQ rows (4 broadcast loads, RQ * 16 FMAs)
template <int RQ>
__device__ __forceinline__ void axpys16(float (&acc)[RQ][16], const float (&w)[RQ], const float* row) {
  float4 v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) v[t] = reinterpret_cast<const float4*>(row)[t];
#pragma unroll
  for (int i = 0; i < RQ; ++i)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      acc[i][4 * t + 0] = fmaf(w[i], v[t].x, acc[i][4 * t + 0]);
      acc[i][4 * t + 1] = fmaf(w[i], v[t].y, acc[i][4 * t + 1]);
      acc[i][4 * t + 2] = fmaf(w[i], v[t].z, acc[i][4 * t + 2]);
      acc[i][4 * t + 3] = fmaf(w[i], v[t].w, acc[i][4 * t + 3]);
    }
}

// grid = (ceil(N / 128), B * H, G)
__global__ void __launch_bounds__(AF_THREADS) attn_fwd_f32_kernel(const AttnF32Params p) {
  constexpr int RQ = AF_RQ_FWD;
  __shared__ __align__(16) float Ks[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Vs[AF_TILE][AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int ql = threadIdx.x & 3, r0 = blockIdx.x * (32 * RQ) + (threadIdx.x >> 2) * RQ;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  float q[RQ][16], o[RQ][16], m[RQ], l[RQ];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    load16(q[i], base + (long long)(r0 + i) * 3 * C + ql * 16, r0 + i < N);
#pragma unroll
    for (int t = 0; t < 16; ++t) { q[i][t] *= p.scale * AF_LOG2E; o[i][t] = 0.f; }   // scores in the log2 domain
    m[i] = -INFINITY;
    l[i] = 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += AF_TILE) {
    __syncthreads();
    stage_tile(Ks, base + C, 3LL * C, j0, N);
    stage_tile(Vs, base + 2 * C, 3LL * C, j0, N);
    __syncthreads();
    const int jn = min(AF_TILE, N - j0);
    for (int jc = 0; jc < jn; jc += 8) {
      float s[8][RQ];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) dots16<RQ>(q, &Ks[jc + jj][ql * 16], s[jj]);
      float mx[RQ], corr[RQ];
#pragma unroll
      for (int i = 0; i < RQ; ++i) mx[i] = m[i];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          const float v = quad_sum(s[jj][i]);
          s[jj][i] = (jc + jj < jn) ? v : -INFINITY;
          mx[i] = fmaxf(mx[i], s[jj][i]);
        }
#pragma unroll
      for (int i = 0; i < RQ; ++i) {
        corr[i] = ex2f(m[i] - mx[i]);        // first chunk: 2^(-inf) = 0 (key jc is always valid, so mx is finite)
        l[i] *= corr[i];
        m[i] = mx[i];
#pragma unroll
        for (int t = 0; t < 16; ++t) o[i][t] *= corr[i];
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float pj[RQ];
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          pj[i] = ex2f(s[jj][i] - mx[i]);
          l[i] += pj[i];
        }
        axpys16<RQ>(o, pj, &Vs[jc + jj][ql * 16]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RQ; ++i)
    if (r0 + i < N) {
      store16(p.out + (tok0 + r0 + i) * C + h * 64 + ql * 16, o[i], 1.0f / l[i]);
      if (ql == 0) p.lse[(((long long)g * p.B + b) * p.H + h) * N + r0 + i] = m[i] * AF_LN2 + logf(l[i]);
    }
}

// grid = (ceil(N / 64), B * H, G): dQ (and delta) of 64 query rows
__global__ void __launch_bounds__(AF_THREADS) attn_bwd_dq_f32_kernel(const AttnF32Params p) {
  constexpr int RQ = AF_RQ_BWD;
  __shared__ __align__(16) float Ks[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Vs[AF_TILE][AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int ql = threadIdx.x & 3, r0 = blockIdx.x * (32 * RQ) + (threadIdx.x >> 2) * RQ;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  const long long stat = (((long long)g * p.B + b) * p.H + h) * N;
  float q[RQ][16], d_o[RQ][16], dq[RQ][16], delta[RQ], lse[RQ];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    const bool active = r0 + i < N;
    load16(q[i], base + (long long)(r0 + i) * 3 * C + ql * 16, active);
    load16(d_o[i], p.d_o + (tok0 + r0 + i) * C + h * 64 + ql * 16, active);
    load16(dq[i], p.o + (tok0 + r0 + i) * C + h * 64 + ql * 16, active);     // dq as scratch for the output row
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s = fmaf(d_o[i][t], dq[i][t], s);
    delta[i] = quad_sum(s);
    lse[i] = active ? p.lse[stat + r0 + i] * AF_LOG2E : 0.f;
    if (active && ql == 0) p.delta[stat + r0 + i] = delta[i];
#pragma unroll
    for (int t = 0; t < 16; ++t) dq[i][t] = 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += AF_TILE) {
    __syncthreads();
    stage_tile(Ks, base + C, 3LL * C, j0, N);
    stage_tile(Vs, base + 2 * C, 3LL * C, j0, N);
    __syncthreads();
    const int jn = min(AF_TILE, N - j0);
#pragma unroll 2
    for (int j = 0; j < jn; ++j) {
      float s[RQ], dp[RQ], ds[RQ];
      dots16<RQ>(q, &Ks[j][ql * 16], s);
      dots16<RQ>(d_o, &Vs[j][ql * 16], dp);
#pragma unroll
      for (int i = 0; i < RQ; ++i) {
        const float pj = ex2f(fmaf(quad_sum(s[i]), p.scale * AF_LOG2E, -lse[i]));
        ds[i] = pj * (quad_sum(dp[i]) - delta[i]);
      }
      axpys16<RQ>(dq, ds, &Ks[j][ql * 16]);
    }
  }
#pragma unroll
  for (int i = 0; i < RQ; ++i)
    if (r0 + i < N) store16(p.out + (tok0 + r0 + i) * 3 * C + h * 64 + ql * 16, dq[i], p.scale);
}

// grid = (ceil(N / 64), B * H, G): dK, dV of 64 key rows
__global__ void __launch_bounds__(AF_THREADS) attn_bwd_dkdv_f32_kernel(const AttnF32Params p) {
  constexpr int RQ = AF_RQ_BWD;
  __shared__ __align__(16) float Qs[AF_TILE][AF_TILE];
  __shared__ __align__(16) float Ds[AF_TILE][AF_TILE];
  __shared__ float lse_s[AF_TILE], delta_s[AF_TILE];
  const int g = blockIdx.z, b = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int C = p.H * 64, N = p.N;
  const int ql = threadIdx.x & 3, r0 = blockIdx.x * (32 * RQ) + (threadIdx.x >> 2) * RQ;
  const long long tok0 = ((long long)g * p.B + b) * N;
  const float* base = p.qkv + tok0 * 3 * C + h * 64;
  const long long stat = (((long long)g * p.B + b) * p.H + h) * N;
  float k[RQ][16], v[RQ][16], dk[RQ][16], dv[RQ][16];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    const bool active = r0 + i < N;
    load16(k[i], base + (long long)(r0 + i) * 3 * C + C + ql * 16, active);
    load16(v[i], base + (long long)(r0 + i) * 3 * C + 2 * C + ql * 16, active);
#pragma unroll
    for (int t = 0; t < 16; ++t) { dk[i][t] = 0.f; dv[i][t] = 0.f; }
  }
  for (int i0 = 0; i0 < N; i0 += AF_TILE) {
    __syncthreads();
    stage_tile(Qs, base, 3LL * C, i0, N);
    stage_tile(Ds, p.d_o + tok0 * C + h * 64, (long long)C, i0, N);
    if (threadIdx.x < AF_TILE) {   // padded query rows: lse = +inf makes their probabilities exactly 0
      const int i = i0 + threadIdx.x;
      lse_s[threadIdx.x] = (i < N) ? p.lse[stat + i] * AF_LOG2E : INFINITY;
      delta_s[threadIdx.x] = (i < N) ? p.delta[stat + i] : 0.f;
    }
    __syncthreads();
    const int in = min(AF_TILE, N - i0);
#pragma unroll 2
    for (int i = 0; i < in; ++i) {
      float s[RQ], dp[RQ], pi[RQ], ds[RQ];
      dots16<RQ>(k, &Qs[i][ql * 16], s);
      dots16<RQ>(v, &Ds[i][ql * 16], dp);
#pragma unroll
      for (int r = 0; r < RQ; ++r) {
        pi[r] = ex2f(fmaf(quad_sum(s[r]), p.scale * AF_LOG2E, -lse_s[i]));
        ds[r] = pi[r] * (quad_sum(dp[r]) - delta_s[i]);
      }
      axpys16<RQ>(dv, pi, &Ds[i][ql * 16]);
      axpys16<RQ>(dk, ds, &Qs[i][ql * 16]);
    }
  }
#pragma unroll
  for (int i = 0; i < RQ; ++i)
    if (r0 + i < N) {
      float* dst = p.out + (tok0 + r0 + i) * 3 * C + h * 64 + ql * 16;
      store16(dst + C, dk[i], p.scale);
      store16(dst + 2 * C, dv[i], 1.0f);
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
  attn_fwd_f32_kernel<<<dim3((N + 32 * AF_RQ_FWD - 1) / (32 * AF_RQ_FWD), B * H, G), AF_THREADS, 0, as_stream(stream)>>>(p);
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
  const dim3 grid((N + 32 * AF_RQ_BWD - 1) / (32 * AF_RQ_BWD), B * H, G);
  attn_bwd_dq_f32_kernel<<<grid, AF_THREADS, 0, as_stream(stream)>>>(p);      // also writes delta for the next kernel
  attn_bwd_dkdv_f32_kernel<<<grid, AF_THREADS, 0, as_stream(stream)>>>(p);
  count_launch(2);
  return check_launch("cavit_attn_bwd_f32");
}

}  // extern "C"
