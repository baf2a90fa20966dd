// cavit-sm100 — K-XFOLD: the single-query cross attention of CrossAttentionBlock with the key / value projections
// and the LayerNorm FOLDED AWAY (exact in real arithmetic, /root/reference/model_cross.py:88-114).
//
// CrossAttention projects ONE query (the CLS token) per sample against all N tokens of the fused sequence
// cat(cls_i, patches_j). Because the query is a single vector per head,
//     s_n = q_h . k_n = q_h . (Wk_h xn_n + bk_h) = (Wk_h^T q_h) . xn_n + const
//     o_h = sum_n p_n v_n = Wv_h (sum_n p_n xn_n) + bv_h          (sum_n p_n = 1)
// the [T, C] x [C, 2C] K/V GEMM (and its dgrad / wgrad), the materialised LayerNorm output of the fused sequence and
// the [T, 2C] K/V activation and gradient are never needed: with q'_h = Wk_h^T q_h (a [B, C] x [C, H*C] GEMM),
// a_h = q'_h o gamma and xhat_n = (x_n - mu_n) rstd_n,
//     s_n   = scale * a_h . xhat_n                         (terms constant in n cancel in the softmax)
//     zhat_h = sum_n p_n xhat_n,   z_h = gamma o zhat_h + beta,   o_h = Wv_h z_h + bv_h  (GEMM on [B, H*C])
// so one pass over the fp32 token streams (read from HBM once, a second time out of L2) replaces
// LayerNorm + K/V GEMM + attention; LayerNorm statistics are computed in the same pass. The backward kernel is the
// exact adjoint: with w_h = (Wv_h^T do_h) o gamma,
//     dp_n = w_h . xhat_n,  ds_n = p_n (dp_n - sum_j p_j dp_j),  dxhat_n = sum_h scale ds_hn a_h + p_hn w_h,
//     dx_n = rstd_n (dxhat_n - mean(dxhat_n) - xhat_n mean(dxhat_n o xhat_n))     (both means follow from s_n, dp_n)
//     da_h = scale sum_n ds_hn xhat_n  ->  dq'_h = da_h o gamma,  dgamma += da_h o q'_h + gz_h o zhat_h,  dbeta += gz_h.
// Two implementations of this algebra live here:
//   * xfold_fwd_kernel / xfold_bwd_kernel: everything fp32 on the CUDA cores (the token streams are fp32), MORE accurate than the
//     bf16 K/V route; one CTA per (fusion k, sample b), one thread per channel pair (C <= 1024). Used by the fp32-tolerance
//     mode and for shapes outside the tensor-core tile. Algorithmic traffic 4*C bytes per token forward, 12*C backward (read
//     x, read-modify-write the stream gradient), but the kernels are bound by instruction issue (~55 thread instructions
//     per token element), not by HBM;
//   * xfold_tc_fwd_kernel / xfold_tc_bwd_kernel (further down): the contractions on tcgen05 from one bf16 xhat tile per
//     sample — the bf16 mode's default where the shape fits.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int XF_MAX_FUSIONS = 16;

struct XfoldParams {
  const float* x;       // token streams [M][B*N][C] fp32
  const float* cls;     // CLS rows of the fused sequences [K][B][C] fp32 (row 0)
  const float* qp;      // q' [K][B][H][C] fp32
  const float* gamma;   // [K][C]
  const float* beta;    // [K][C]
  float* zhat;          // [K][B][H][C] fp32
  bf16* z;              // [K][B][H][C] bf16 (forward only)
  bf16* z_lo;           // fp32-tolerance mode: second bf16 plane, z ~ z + z_lo (NULL otherwise)
  float* probs;         // [K][B][H][N] softmax (pre-dropout)
  float* mean;          // [K][B][N]
  float* rstd;          // [K][B][N]
  float* scratch;       // forward: [K][B][N][HP]; backward: [K][B][N][2*HP + 2]
  // backward only
  const float* gz;      // [K][B][H][C] fp32 = Wv_h^T do_h
  float* dx;            // stream gradients [M][B*N][C] fp32, atomically accumulated: rows n >= 1 of stream tok_src[k],
                        // row 0 (the CLS row of the fused sequence) of stream cls_src[k]
  float* dqp;           // [K][B][H][C] fp32
  float* dgamma;        // [K][C] (atomically accumulated)
  float* dbeta;         // [K][C] (atomically accumulated)
  int B, N, C, H, K;
  int cls_src[XF_MAX_FUSIONS], tok_src[XF_MAX_FUSIONS];
  float scale, eps;
  DropCfg drop;
  int use_drop;
  int disjoint;   // every token stream is read by at most one fusion: plain read-modify-write instead of atomics
  int* status_word;   // device status word (bounded mbarrier waits of the tcgen05 variant)
  int precise;        // backward: keep the all-fp32 CUDA-core kernel (fp32-tolerance mode)
};

__device__ __forceinline__ const float* xf_row(const XfoldParams& p, int k, int b, int n) {
  if (n == 0) return p.cls + ((long long)k * p.B + b) * p.C;
  return p.x + ((long long)p.tok_src[k] * p.B * p.N + (long long)b * p.N + n) * p.C;
}

__device__ __forceinline__ float xf_block_sum(float v, float* red, int nwarps) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < nwarps; ++w) r += red[w];
  return r;
}

constexpr int XF_CHUNK = 64;  // token rows whose per-row coefficients are staged in shared memory at a time
template <int H> struct XfCfg {
  static constexpr int THREADS = 32 * H;                  // one thread per PAIR of channels (C = 64 H)
  static constexpr int NVL = (H + 1) / 2;                 // float4 per lane of a warp that owns a token row
  static constexpr int MINB = H <= 2 ? 8 : (H <= 4 ? 5 : (H <= 6 ? 4 : (H <= 8 ? 3 : 2)));
  static constexpr int HP = (H + 3) & ~3;
};

// Sum V per-lane values across the 32 lanes with ~V + log2(32/V) shuffles instead of 5 V: at every halving step a lane
// keeps one half of its values and trades the other half with its partner. Afterwards lane l holds the total of value
// warp_multi_index<V>(l) (every index is held by 32 / V lanes).
template <int V>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[V], int lane) {
  int off = 16;
#pragma unroll
  for (int cnt = V; cnt > 1; cnt >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) {
      const float send = up ? v[i] : v[i + cnt / 2];
      const float keep = up ? v[i + cnt / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o <= off) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}
template <int V>
__device__ __forceinline__ int warp_multi_index(int lane) {
  int idx = 0, bit = 4;
#pragma unroll
  for (int cnt = V; cnt > 1; cnt >>= 1, --bit) idx = (idx << 1) | ((lane >> bit) & 1);
  return idx;
}
constexpr int xf_pow2(int h) { return h <= 1 ? 1 : (h <= 2 ? 2 : (h <= 4 ? 4 : (h <= 8 ? 8 : 16))); }

// Dot products of NV row vectors (this lane's float4 slices of up to two token rows) with the H (or 2H) folded vectors
// in shared memory. Processing two rows per sweep halves the shared-memory traffic, which bounds pass 1.
template <int H, int NVL, int NVEC>
__device__ __forceinline__ void xf_dots(const float* s_vec, int C, int nv, int lane, const float4 (&v0)[NVL],
                                        const float4 (&v1)[NVL], float (&d0)[NVEC * H], float (&d1)[NVEC * H]) {
#pragma unroll
  for (int j = 0; j < NVEC * H; ++j) {
    const float4* vec = reinterpret_cast<const float4*>(s_vec + j * C);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nv) {
        const float4 a = vec[c4];
        a0 += (a.x * v0[i].x + a.y * v0[i].y) + (a.z * v0[i].z + a.w * v0[i].w);
        a1 += (a.x * v1[i].x + a.y * v1[i].y) + (a.z * v1[i].z + a.w * v1[i].w);
      }
    }
    d0[j] = a0;   // per-lane partial sums; reduced across the warp by warp_multi_reduce
    d1[j] = a1;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// smem: a[H][C] | A[H] | m[H] | chunk[XF_CHUNK][HP] | mu[XF_CHUNK] | rstd[XF_CHUNK] | f[H]
template <int H>
__global__ void __launch_bounds__(XfCfg<H>::THREADS, XfCfg<H>::MINB)
xfold_fwd_kernel(const XfoldParams p) {
  constexpr int HP = XfCfg<H>::HP, NVL = XfCfg<H>::NVL;
  extern __shared__ float sm[];
  float* s_a = sm;                 // [H][C]
  float* s_A = s_a + H * p.C;      // [H]
  float* s_m = s_A + H;            // [H]
  float* s_ch = s_m + H + ((4 - ((2 * H) & 3)) & 3);   // [XF_CHUNK][HP], 16-byte aligned: scores, then weights of the chunk
  float* s_mu = s_ch + XF_CHUNK * HP;                  // [XF_CHUNK] mean of the chunk's rows
  float* s_rs = s_mu + XF_CHUNK;                       // [XF_CHUNK] rstd
  float* s_f = s_rs + XF_CHUNK;                        // [H] rescale factor of the running accumulators (then 1 / L_h)
  const int b = blockIdx.x, k = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int C = p.C, N = p.N;
  const long long kb = (long long)k * p.B + b;
  const float* qp = p.qp + kb * H * C;
  const float* gamma = p.gamma + (long long)k * C;
  // a_h = q'_h o gamma; A_h = sum_c a_h[c]
  for (int h = 0; h < H; ++h)
    for (int c = tid; c < C; c += blockDim.x) s_a[h * C + c] = qp[h * C + c] * gamma[c];
  __syncthreads();
  for (int h = warp; h < H; h += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += s_a[h * C + c];
    s = warp_sum(s);
    if (lane == 0) s_A[h] = s;
  }
  __syncthreads();
  // The token rows are walked ONCE from HBM, in chunks of XF_CHUNK rows (round 2; before, pass 2 re-read all N rows after
  // pass 1 had streamed them and 2x the algorithmic bytes came from DRAM: 629 MB for 310 MB, L2 hit rate 4 %, ncu):
  //   (A) one warp per PAIR of rows of the chunk: LayerNorm statistics and the H scores of each row (into shared memory);
  //   (B) one warp per head: ONLINE softmax update — running max M_h, running sum L_h, rescale factor f_h for what has
  //       been accumulated so far, weights w_hn = exp(s_hn - M_h) rstd_n of the chunk's rows;
  //   (C) one thread per channel pair: acc_h = acc_h f_h + sum_n w_hn x_n — the second read of the chunk's rows comes out of
  //       L1 / L2, they were fetched a few microseconds earlier by this same CTA.
  // Afterwards zhat_h = acc_h / L_h - m_h and the probabilities are normalised from the raw scores kept in `probs`.
  const int nv = C >> 2;  // float4 per row
  const int c = 2 * tid;
  float2 acc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) acc[h] = make_float2(0.f, 0.f);
  const float* xs = p.x + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C + c;
  const float* x0 = p.cls + kb * C + c;
  float run_M = -INFINITY, run_L = 0.f, run_m = 0.f;      // per head, kept by lane 0.. of warp (h % nwarps); see (B)
  for (int n0 = 0; n0 < N; n0 += XF_CHUNK) {
    const int rows = min(XF_CHUNK, N - n0);
    // ---- (A)
    for (int rr = 2 * warp; rr < rows; rr += 2 * nwarps) {
      const int n = n0 + rr;
      const bool two = rr + 1 < rows;
      const float4* row0 = reinterpret_cast<const float4*>(xf_row(p, k, b, n));
      const float4* row1 = reinterpret_cast<const float4*>(xf_row(p, k, b, two ? n + 1 : n));
      float4 v0[NVL], v1[NVL];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < NVL; ++i) {
        const int c4 = lane + 32 * i;
        if (c4 < nv) {
          v0[i] = __ldg(row0 + c4);
          v1[i] = __ldg(row1 + c4);
          s0 += (v0[i].x + v0[i].y) + (v0[i].z + v0[i].w);
          s1 += (v1[i].x + v1[i].y) + (v1[i].z + v1[i].w);
        }
      }
      const float mu0 = warp_sum(s0) / (float)C, mu1 = warp_sum(s1) / (float)C;
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int i = 0; i < NVL; ++i) {
        const int c4 = lane + 32 * i;
        if (c4 < nv) {
          float d0 = v0[i].x - mu0, d1 = v0[i].y - mu0, d2 = v0[i].z - mu0, d3 = v0[i].w - mu0;
          q0 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          d0 = v1[i].x - mu1; d1 = v1[i].y - mu1; d2 = v1[i].z - mu1; d3 = v1[i].w - mu1;
          q1 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
      }
      const float rs0 = rsqrtf(warp_sum(q0) / (float)C + p.eps), rs1 = rsqrtf(warp_sum(q1) / (float)C + p.eps);
      float d0[H], d1[H];
      xf_dots<H, NVL, 1>(s_a, C, nv, lane, v0, v1, d0, d1);
      constexpr int HQ = xf_pow2(H);
      float red[2 * HQ];
#pragma unroll
      for (int h = 0; h < HQ; ++h) {
        red[h] = h < H ? d0[h] : 0.f;
        red[HQ + h] = h < H ? d1[h] : 0.f;
      }
      const float tot = warp_multi_reduce<2 * HQ>(red, lane);
      const int idx = warp_multi_index<2 * HQ>(lane), r = idx / HQ, h = idx % HQ;
      if (h < H && (r == 0 || two) && (lane & (32 / (2 * HQ) - 1)) == 0) {  // one lane per (row, head)
        const float mu = r ? mu1 : mu0, rs = r ? rs1 : rs0;
        s_ch[(rr + r) * HP + h] = p.scale * rs * (tot - mu * s_A[h]);        // score s_hn
      }
      if (lane == 0) {
        s_mu[rr] = mu0; s_rs[rr] = rs0;
        p.mean[kb * N + n] = mu0;
        p.rstd[kb * N + n] = rs0;
        if (two) {
          s_mu[rr + 1] = mu1; s_rs[rr + 1] = rs1;
          p.mean[kb * N + n + 1] = mu1;
          p.rstd[kb * N + n + 1] = rs1;
        }
      }
    }
    __syncthreads();
    // ---- (B) heads h = warp, warp + nwarps, ...: each warp owns the running statistics of its heads in registers; a warp
    // owns at most ONE head when nwarps >= H, which holds for every instantiation (threads = 32 H)
    if (warp < H) {
      const int h = warp;
      float mx = run_M;
      for (int r = lane; r < rows; r += 32) mx = fmaxf(mx, s_ch[r * HP + h]);
      mx = warp_max(mx);
      const float f = __expf(run_M - mx);       // first chunk: exp(-inf) = 0
      float sum = 0.f, m = 0.f;
      float* pr = p.probs + (kb * H + h) * N + n0;
      for (int r = lane; r < rows; r += 32) {
        const float sc = s_ch[r * HP + h];
        pr[r] = sc;                              // raw score; normalised after the last chunk
        const float e = __expf(sc - mx);
        sum += e;
        const float w = e * s_rs[r];
        s_ch[r * HP + h] = w;
        m += w * s_mu[r];
      }
      run_L = run_L * f + warp_sum(sum);
      run_m = run_m * f + warp_sum(m);
      run_M = mx;
      if (lane == 0) s_f[h] = f;
    }
    __syncthreads();
    // ---- (C)
    {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float f = s_f[h];
        acc[h].x *= f;
        acc[h].y *= f;
      }
      for (int r0 = 0; r0 < rows; r0 += 8) {
        float2 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // 8 independent 8-byte loads in flight per thread
          const int n = n0 + r0 + j;
          xv[j] = (r0 + j < rows) ? __ldg(reinterpret_cast<const float2*>(n == 0 ? x0 : xs + (long long)n * C)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (r0 + j < rows) {
            const float4* w4 = reinterpret_cast<const float4*>(s_ch + (r0 + j) * HP);
#pragma unroll
            for (int i = 0; i < HP / 4; ++i) {
              const float4 t = w4[i];
              const float wv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (4 * i + e < H) {
                  acc[4 * i + e].x = fmaf(wv[e], xv[j].x, acc[4 * i + e].x);
                  acc[4 * i + e].y = fmaf(wv[e], xv[j].y, acc[4 * i + e].y);
                }
            }
          }
        }
      }
    }
    __syncthreads();     // s_ch / s_mu / s_rs / s_f are rewritten by the next chunk
  }
  // ---- final normalisation: 1 / L_h, m_h / L_h, probabilities
  if (warp < H) {
    const int h = warp;
    const float inv = 1.0f / run_L;
    if (lane == 0) {
      s_f[h] = inv;
      s_m[h] = run_m * inv;
    }
    float* pr = p.probs + (kb * H + h) * N;
    for (int n = lane; n < N; n += 32) pr[n] = __expf(pr[n] - run_M) * inv;
  }
  __syncthreads();
  {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      acc[h].x *= s_f[h];
      acc[h].y *= s_f[h];
    }
    const float2 g = *reinterpret_cast<const float2*>(gamma + c);
    const float2 be = *reinterpret_cast<const float2*>(p.beta + (long long)k * C + c);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float zx = acc[h].x - s_m[h], zy = acc[h].y - s_m[h];
      *reinterpret_cast<float2*>(p.zhat + (kb * H + h) * C + c) = make_float2(zx, zy);
      const float z0 = fmaf(g.x, zx, be.x), z1 = fmaf(g.y, zy, be.y);
      const uint32_t zh = pack_bf16(z0, z1);
      *reinterpret_cast<uint32_t*>(p.z + (kb * H + h) * C + c) = zh;
      if (p.z_lo) {
        const float2 r = unpack_bf16_fast(zh);
        *reinterpret_cast<uint32_t*>(p.z_lo + (kb * H + h) * C + c) = pack_bf16(z0 - r.x, z1 - r.y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// scratch per token row (floats): sdot[HP] | dpraw[HP] | ds[HP] | peff[HP] | c1 | c2 | mu | rstd
// smem: a[H][C] | w[H][C] | A[H] | W[H] | chunk[XF_CHUNK][2*HP + 4]
template <int H>
__global__ void __launch_bounds__(XfCfg<H>::THREADS, XfCfg<H>::MINB)
xfold_bwd_kernel(const XfoldParams p) {
  constexpr int HP = XfCfg<H>::HP, NVL = XfCfg<H>::NVL;
  constexpr int RS = 4 * HP + 4;
  constexpr int CS = 2 * HP + 4;   // staged part of a scratch row: ds[HP] | peff[HP] | c1 c2 mu rstd
  extern __shared__ float sm[];
  const int C = p.C, N = p.N;
  float* s_a = sm;               // [H][C]  a_h = q'_h o gamma
  float* s_w = s_a + H * C;      // [H][C]  w_h = gz_h o gamma   (contiguous after a: 2H vectors for xf_dots)
  float* s_A = s_w + H * C;      // [H]
  float* s_W = s_A + H;          // [H]
  float* s_ch = s_W + H + ((4 - ((2 * H) & 3)) & 3);   // 16-byte aligned
  const int b = blockIdx.x, k = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const long long kb = (long long)k * p.B + b;
  const float* qp = p.qp + kb * H * C;
  const float* gz = p.gz + kb * H * C;
  const float* gamma = p.gamma + (long long)k * C;
  float* rowc = p.scratch + kb * N * RS;
  for (int h = 0; h < H; ++h)
    for (int c = tid; c < C; c += blockDim.x) {
      const float g = gamma[c];
      s_a[h * C + c] = qp[h * C + c] * g;
      s_w[h * C + c] = gz[h * C + c] * g;
    }
  __syncthreads();
  for (int h = warp; h < 2 * H; h += nwarps) {
    const float* v = s_a + h * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += v[c];
    s = warp_sum(s);
    if (lane == 0) { if (h < H) s_A[h] = s; else s_W[h - H] = s; }
  }
  __syncthreads();
  // ---- pass 1: one warp per PAIR of token rows: sdot_h = scale a_h . xhat_n, dpraw_h = w_h . xhat_n
  const int nv = C >> 2;
  for (int n = 2 * warp; n < N; n += 2 * nwarps) {
    const bool two = n + 1 < N;
    const int n1 = two ? n + 1 : n;
    const float4* row0 = reinterpret_cast<const float4*>(xf_row(p, k, b, n));
    const float4* row1 = reinterpret_cast<const float4*>(xf_row(p, k, b, n1));
    const float mu0 = p.mean[kb * N + n], rs0 = p.rstd[kb * N + n];
    const float mu1 = p.mean[kb * N + n1], rs1 = p.rstd[kb * N + n1];
    float4 v0[NVL], v1[NVL];
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nv) {
        v0[i] = __ldg(row0 + c4);
        v1[i] = __ldg(row1 + c4);
      }
    }
    float d0[2 * H], d1[2 * H];
    xf_dots<H, NVL, 2>(s_a, C, nv, lane, v0, v1, d0, d1);
    constexpr int HQ = xf_pow2(H);
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // a-dots, then w-dots: [2 rows][HQ] values per reduction
      float red[2 * HQ];
#pragma unroll
      for (int h = 0; h < HQ; ++h) {
        red[h] = h < H ? d0[which * H + h] : 0.f;
        red[HQ + h] = h < H ? d1[which * H + h] : 0.f;
      }
      const float tot = warp_multi_reduce<2 * HQ>(red, lane);
      const int idx = warp_multi_index<2 * HQ>(lane), r = idx / HQ, h = idx % HQ;
      if (h < H && (r == 0 || two) && (lane & (32 / (2 * HQ) - 1)) == 0) {
        const float mu = r ? mu1 : mu0, rs = r ? rs1 : rs0;
        float* row = rowc + (long long)(n + r) * RS;
        if (which == 0) row[h] = p.scale * rs * (tot - mu * s_A[h]);
        else row[HP + h] = rs * (tot - mu * s_W[h]);
      }
    }
    if (lane == 0) {
      float* r = rowc + (long long)n * RS;
      r[4 * HP + 2] = mu0;
      r[4 * HP + 3] = rs0;
      if (two) {
        r[RS + 4 * HP + 2] = mu1;
        r[RS + 4 * HP + 3] = rs1;
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- softmax backward, one warp per head: D_h = sum_n p dpraw; ds_hn = p_hn (dpraw_hn - D_h)
  for (int h = warp; h < H; h += nwarps) {
    const float* pr = p.probs + (kb * H + h) * N;
    float D = 0.f;
    for (int n = lane; n < N; n += 32) {
      const float pe = pr[n];
      rowc[(long long)n * RS + 3 * HP + h] = pe;
      D += pe * rowc[(long long)n * RS + HP + h];
    }
    D = warp_sum(D);
    for (int n = lane; n < N; n += 32) rowc[(long long)n * RS + 2 * HP + h] = pr[n] * (rowc[(long long)n * RS + HP + h] - D);
  }
  __threadfence_block();
  __syncthreads();
  // ---- per row: c1 = mean_c(dxhat), c2 = mean_c(dxhat o xhat)
  const float invC = 1.0f / (float)C;
  for (int n = tid; n < N; n += blockDim.x) {
    float* r = rowc + (long long)n * RS;
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float ds = r[2 * HP + h], pe = r[3 * HP + h];
      c1 += p.scale * ds * s_A[h] + pe * s_W[h];
      c2 += ds * r[h] + pe * r[HP + h];
    }
    r[4 * HP] = c1 * invC;
    r[4 * HP + 1] = c2 * invC;
  }
  __threadfence_block();
  __syncthreads();
  // ---- pass 2: one thread per channel pair; the per-row coefficients are staged per chunk of rows in shared memory
  {
    const int c = 2 * tid;
    float2 av[H], wv[H], dacc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      av[h] = make_float2(s_a[h * C + c] * p.scale, s_a[h * C + c + 1] * p.scale);
      wv[h] = make_float2(s_w[h * C + c], s_w[h * C + c + 1]);
      dacc[h] = make_float2(0.f, 0.f);
    }
    const float* xs = p.x + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C + c;
    const float* x0 = p.cls + kb * C + c;
    float* dxs = p.dx + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C + c;
    float* dx0 = p.dx + ((long long)p.cls_src[k] * p.B * N + (long long)b * N) * C + c;
    for (int n0 = 0; n0 < N; n0 += XF_CHUNK) {
      const int rows = min(XF_CHUNK, N - n0);
      __syncthreads();
      for (int i = tid; i < rows * CS; i += blockDim.x) s_ch[i] = rowc[(long long)(n0 + i / CS) * RS + 2 * HP + i % CS];
      __syncthreads();
      for (int r0 = 0; r0 < rows; r0 += 4) {
        float2 xv[4], old[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // all loads first: the stores below may alias as far as the compiler knows
          const int n = n0 + r0 + j;
          const bool ok = r0 + j < rows;
          xv[j] = ok ? __ldg(reinterpret_cast<const float2*>(n == 0 ? x0 : xs + (long long)n * C)) : make_float2(0.f, 0.f);
          old[j] = (ok && p.disjoint) ? *reinterpret_cast<const float2*>(n == 0 ? dx0 : dxs + (long long)n * C)
                                      : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + r0 + j;
          if (r0 + j < rows) {
            const float4* r4 = reinterpret_cast<const float4*>(s_ch + (r0 + j) * CS);
            float ds[HP], pe[HP];
#pragma unroll
            for (int i = 0; i < HP / 4; ++i) {
              const float4 t = r4[i];
              ds[4 * i] = t.x; ds[4 * i + 1] = t.y; ds[4 * i + 2] = t.z; ds[4 * i + 3] = t.w;
              const float4 u = r4[HP / 4 + i];
              pe[4 * i] = u.x; pe[4 * i + 1] = u.y; pe[4 * i + 2] = u.z; pe[4 * i + 3] = u.w;
            }
            const float4 tl = r4[2 * (HP / 4)];  // c1 | c2 | mu | rstd
            const float xhx = (xv[j].x - tl.z) * tl.w, xhy = (xv[j].y - tl.z) * tl.w;
            float dx_ = 0.f, dy_ = 0.f;
#pragma unroll
            for (int h = 0; h < H; ++h) {
              dx_ = fmaf(ds[h], av[h].x, dx_);
              dy_ = fmaf(ds[h], av[h].y, dy_);
              dx_ = fmaf(pe[h], wv[h].x, dx_);
              dy_ = fmaf(pe[h], wv[h].y, dy_);
              dacc[h].x = fmaf(ds[h], xhx, dacc[h].x);
              dacc[h].y = fmaf(ds[h], xhy, dacc[h].y);
            }
            const float ox = tl.w * (dx_ - tl.x - xhx * tl.y), oy = tl.w * (dy_ - tl.x - xhy * tl.y);
            float* dst = (n == 0) ? dx0 : dxs + (long long)n * C;
            if (p.disjoint) {
              *reinterpret_cast<float2*>(dst) = make_float2(old[j].x + ox, old[j].y + oy);
            } else {
              atomicAdd(dst, ox);
              atomicAdd(dst + 1, oy);
            }
          }
        }
      }
    }
    const float2 g = *reinterpret_cast<const float2*>(gamma + c);
    float2 dg = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float dax = dacc[h].x * p.scale, day = dacc[h].y * p.scale;   // da_h[c] = scale sum_n ds_hn xhat_n[c]
      *reinterpret_cast<float2*>(p.dqp + (kb * H + h) * C + c) = make_float2(dax * g.x, day * g.y);
      const float2 gzh = *reinterpret_cast<const float2*>(gz + h * C + c);
      const float2 qph = *reinterpret_cast<const float2*>(qp + h * C + c);
      const float2 zh = *reinterpret_cast<const float2*>(p.zhat + (kb * H + h) * C + c);
      dg.x += dax * qph.x + gzh.x * zh.x;
      dg.y += day * qph.y + gzh.y * zh.y;
      db.x += gzh.x;
      db.y += gzh.y;
    }
    atomicAdd(p.dgamma + (long long)k * C + c, dg.x);
    atomicAdd(p.dgamma + (long long)k * C + c + 1, dg.y);
    atomicAdd(p.dbeta + (long long)k * C + c, db.x);
    atomicAdd(p.dbeta + (long long)k * C + c + 1, db.y);
  }
}

template <int H>
static int xfold_launch(bool bwd, const XfoldParams& p, cudaStream_t st) {
  constexpr int HP = XfCfg<H>::HP;
  const int threads = XfCfg<H>::THREADS;
  size_t smem = sizeof(float) * (bwd ? (2 * (size_t)H * p.C + 2 * H + 4 + XF_CHUNK * (2 * HP + 4))
                                     : ((size_t)H * p.C + 2 * H + 4 + XF_CHUNK * HP + 2 * XF_CHUNK + H + 4));
  // Residency cap (experiment knob): the token rows of a (fusion, sample) pair are read twice; with every SM holding MINB
  // CTAs the rows in flight between the two passes exceed L2. Padding the dynamic shared memory request lowers residency.
  static const int pad_kb = [] { const char* e = getenv("CAVIT_XFOLD_PAD_KB"); return e ? atoi(e) : 0; }();
  if (pad_kb > 0 && smem < (size_t)pad_kb * 1024) smem = (size_t)pad_kb * 1024;
  if (bwd) {
    static PerDeviceMax cur;
    if (smem > cur.get()) {
      if (cudaFuncSetAttribute(xfold_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(CAVIT_E_LAUNCH, "xfold bwd smem attribute");
      cur.set(smem);
    }
    xfold_bwd_kernel<H><<<dim3(p.B, p.K), threads, smem, st>>>(p);
  } else {
    static PerDeviceMax cur;
    if (smem > cur.get()) {
      if (cudaFuncSetAttribute(xfold_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(CAVIT_E_LAUNCH, "xfold fwd smem attribute");
      cur.set(smem);
    }
    xfold_fwd_kernel<H><<<dim3(p.B, p.K), threads, smem, st>>>(p);
  }
  count_launch();
  return check_launch(bwd ? "cavit_xfold_bwd" : "cavit_xfold_fwd");
}

// ------------------------------------------------------------------------------------------------ forward on tcgen05
// The same fold with both contractions on the tensor cores (bf16 mode; N <= 256, C % 128 == 0, the tile below fits shared
// memory). The SIMT kernel above spends ~55 thread instructions per token element (ncu: 58 % issue slots busy at 20 % of the
// HBM peak): its two skinny contractions — H scores per token, H weighted sums per channel — are 12 FMA per element plus
// the shared-memory traffic and index arithmetic to feed them. Here, per (fusion, sample):
//   1. sixteen warps stream the fp32 token rows ONCE, compute the LayerNorm statistics and write xhat = (x - mu) rstd as bf16
//      into ONE 128-byte-swizzled tile [n][c] (C/64 column chunks of [RA rows][128 B]);
//   2. S[n][h]  = sum_c xhat[n][c] a_h[c]    tcgen05, A = the tile read K-major,  B = a (bf16, [16][C]);   D in TMEM [n][16]
//   3. softmax over n per head (one thread per token row, block reductions), probabilities to HBM (backward needs them) and
//      as bf16 into a K-major tile P[h][n];
//   4. zhat[c][h] = sum_n xhat[n][c] P[h][n]  tcgen05, A = THE SAME tile read MN-major (M = channels), B = P;  D in TMEM [c][16]
//   5. epilogue: zhat (fp32) and z = gamma o zhat + beta (bf16), one thread per channel.
// xhat in bf16 is what the unfolded route feeds its K / V GEMMs too; accumulation is fp32. About 7 thread instructions per
// token element remain (step 1), the rest is HBM time.
constexpr int XT_THREADS = 512;                 // 16 warps: 64 token rows (4 per warp) in flight while streaming; rows / channels use the first 8
constexpr int XT_WARPS = XT_THREADS / 32;
constexpr int XT_HP = 16;                      // heads padded to the smallest UMMA N for M = 128

struct XtLayout {
  int RA, pitch, chunks, kchunks, off_b1, off_p, off_red, off_bar, bytes;
};
static XtLayout xt_layout(int N, int C) {
  XtLayout L;
  L.RA = (N + 15) & ~15;
  L.pitch = L.RA * 128;
  L.chunks = C / 64;
  L.kchunks = (L.RA + 63) / 64;
  L.off_b1 = L.chunks * L.pitch;                 // a tile: [chunks][16 rows][128 B]
  L.off_p = L.off_b1 + L.chunks * 2048;          // P tile: [kchunks][16 rows][128 B]
  L.off_red = L.off_p + 4 * 2048;                // always 4 k-chunks: the M = 128 reads of step 2 may run over the tile end
  L.off_bar = L.off_red + 2 * XT_WARPS * XT_HP * 4;     // [max | sum][warp][head]
  L.bytes = L.off_bar + 64 + 1024;               // + alignment slack
  return L;
}

// L2 prefetch of the token rows of a warp's NEXT streaming iteration (4 rows, XT_WARPS apart, 4 NV lines of 128 bytes each)
template <int NV>
__device__ __forceinline__ void xt_prefetch_rows(const float* xs, int r_next, int N, int lane) {
  constexpr int LINES = 4 * NV;                    // per row
#ifdef CAVIT_NO_XF_PREFETCH
  return;
#endif
#pragma unroll
  for (int q = lane; q < 4 * LINES; q += 32) {
    const int nn = r_next + XT_WARPS * (q / LINES);
    if (nn >= 1 && nn < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + (long long)nn * (128 * NV) + (q % LINES) * 32));
  }
}

template <int H, int NV>     // NV = C / 128 float4 per lane and token row
__global__ void __launch_bounds__(XT_THREADS, 1)
xfold_tc_fwd_kernel(const XfoldParams p, const XtLayout L) {
  extern __shared__ uint8_t xt_raw[];
  const uint32_t base = (smem_u32(xt_raw) + 1023u) & ~1023u;
  uint8_t* gen = xt_raw + (base - smem_u32(xt_raw));
  uint8_t* g_b1 = gen + L.off_b1;
  uint8_t* g_p = gen + L.off_p;
  float* s_red = reinterpret_cast<float*>(gen + L.off_red);
  const uint32_t bar1 = base + L.off_bar, bar2 = bar1 + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + L.off_bar + 16);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int C = 128 * NV, ctiles = NV;
  const int N = p.N, mtiles = (N + 127) >> 7;

  if (tid == 0) {
    *abort_flag = 0;
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  // rows h >= H of the a tile and of P are never written: zero both tiles once
  for (int i = tid; i < (L.off_red - L.off_b1) / 16; i += XT_THREADS) reinterpret_cast<uint4*>(g_b1)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;              // S tiles: columns [0, 32); zhat tiles: [32 + 16 j, ...)
  const uint32_t idesc1 = umma_idesc_bf16(XT_HP, 0, 0, 128), idesc2 = umma_idesc_bf16(XT_HP, 1, 0, 128);
  const float inv_c = 1.0f / (float)C;
  const int items = p.K * p.B;
  uint32_t phase = 0;

  for (int item = blockIdx.x; item < items; item += gridDim.x, phase ^= 1u) {
    const int k = item / p.B, b = item - k * p.B;
    const long long kb = item;
    const float* xs = p.x + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C;   // token rows n >= 1 of the fused sequence
    const float* x0 = p.cls + kb * C;                                                    // row 0: the CLS row
    // ---- a_h = q'_h o gamma as bf16 into the K-major B tile of step 2
    {
      const float* qp = p.qp + kb * H * C;
      const float* gamma = p.gamma + (long long)k * C;
      for (int i = tid; i < H * (C >> 3); i += XT_THREADS) {
        const int h = i / (C >> 3), c8 = i - h * (C >> 3);          // 8 channels = one 16-byte unit
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(qp + h * C + c8 * 8)), q1 = __ldg(reinterpret_cast<const float4*>(qp + h * C + c8 * 8) + 1);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8) + 1);
        uint4 w;
        w.x = pack_bf16(q0.x * g0.x, q0.y * g0.y); w.y = pack_bf16(q0.z * g0.z, q0.w * g0.w);
        w.z = pack_bf16(q1.x * g1.x, q1.y * g1.y); w.w = pack_bf16(q1.z * g1.z, q1.w * g1.w);
        *reinterpret_cast<uint4*>(g_b1 + (c8 >> 3) * 2048 + h * 128 + (((c8 & 7) ^ (h & 7)) << 4)) = w;
      }
    }
    // ---- step 1: token rows -> LayerNorm statistics -> xhat (bf16) tile; four rows per warp in flight. Rows are written
    // in order, so the S MMAs of the first 128 rows (step 2, tile 0) are issued as soon as those rows are in shared memory
    // and run under the streaming of the remaining rows.
    auto issue_s_tile = [&](int m) {
      for (int cc = 0; cc < L.chunks; ++cc) {
        const uint64_t ad = umma_desc_sw128(base + cc * L.pitch + m * 16384, 16, 1024);
        const uint64_t bd = umma_desc_sw128(base + L.off_b1 + cc * 2048, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem + m * XT_HP, ad + 2 * kk, bd + 2 * kk, idesc1, (cc | kk) != 0 ? 1u : 0u);
      }
    };
    bool tile0_issued = false;
    for (int r0 = warp; r0 < L.RA; r0 += 4 * XT_WARPS) {
      if (mtiles == 2 && !tile0_issued && r0 - warp >= 128) {   // block-uniform: rows 0 .. 127 are complete
        tile0_issued = true;
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) issue_s_tile(0);
          __syncwarp();
        }
      }
      xt_prefetch_rows<NV>(xs, r0 + 4 * XT_WARPS, N, lane);
      float4 v[4][NV];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = r0 + XT_WARPS * j;
        if (n < N) {
          const float4* row = reinterpret_cast<const float4*>(n == 0 ? x0 : xs + (long long)n * C) + lane;
#pragma unroll
          for (int i = 0; i < NV; ++i) v[j][i] = __ldg(row + 32 * i);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = r0 + XT_WARPS * j;
        if (n >= L.RA) continue;                 // warp-uniform
        uint8_t* dst = gen + n * 128;
        const uint32_t sw = static_cast<uint32_t>(n & 7);
        // this lane's float4 i sits in column chunk (lane >> 4) + 2 i, 16-byte unit ((lane & 15) >> 1) ^ (n & 7), half lane & 1
        uint8_t* d0 = dst + (lane >> 4) * L.pitch + (((static_cast<uint32_t>(lane & 15) >> 1) ^ sw) << 4) + (lane & 1) * 8;
        if (n >= N) {                            // rows of the last 16-row group beyond N: zero (step 4 reads them)
#pragma unroll
          for (int i = 0; i < NV; ++i) *reinterpret_cast<uint2*>(d0 + 2 * i * L.pitch) = make_uint2(0u, 0u);
          continue;
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) sum += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
        const float mu = warp_sum(sum) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[j][i].x -= mu; v[j][i].y -= mu; v[j][i].z -= mu; v[j][i].w -= mu;
          q += (v[j][i].x * v[j][i].x + v[j][i].y * v[j][i].y) + (v[j][i].z * v[j][i].z + v[j][i].w * v[j][i].w);
        }
        const float rs = rsqrtf(warp_sum(q) * inv_c + p.eps);
        if (lane == 0) {
          p.mean[kb * N + n] = mu;
          p.rstd[kb * N + n] = rs;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint2 w;
          w.x = pack_bf16(v[j][i].x * rs, v[j][i].y * rs);
          w.y = pack_bf16(v[j][i].z * rs, v[j][i].w * rs);
          *reinterpret_cast<uint2*>(d0 + 2 * i * L.pitch) = w;
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- step 2: S = xhat a^T
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        for (int m = tile0_issued ? 1 : 0; m < mtiles; ++m) issue_s_tile(m);
        umma_commit(bar1);
      }
      __syncwarp();
    }
    mbar_wait(bar1, phase, abort_flag, p.status_word, ERR_TIMEOUT_ATTN);
    tc_fence_after();
    // ---- step 3: softmax over the token axis, one thread per token row (TMEM lane)
    {
      const int n = tid, m = tid >> 7;
      float sc[XT_HP];
      if (m < mtiles) {
        uint32_t r[16];
        tmem_ld16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + m * XT_HP, r);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < XT_HP; ++h) sc[h] = (n < N) ? __uint_as_float(r[h]) * p.scale : -INFINITY;
      } else {
#pragma unroll
        for (int h = 0; h < XT_HP; ++h) sc[h] = -INFINITY;
      }
      float* s_max = s_red;                       // [warp][head]
      float* s_sum = s_red + XT_WARPS * XT_HP;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float mx = warp_max(sc[h]);
        if (lane == 0) s_max[warp * XT_HP + h] = mx;
      }
      __syncthreads();
      float e[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float mx = s_max[h];
#pragma unroll
        for (int w = 1; w < XT_WARPS; ++w) mx = fmaxf(mx, s_max[w * XT_HP + h]);
        e[h] = __expf(sc[h] - mx);               // rows beyond N: exp(-inf) = 0
        const float sm = warp_sum(e[h]);
        if (lane == 0) s_sum[warp * XT_HP + h] = sm;
      }
      __syncthreads();
      float* pr = p.probs + kb * H * N;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float tot = s_sum[h];
#pragma unroll
        for (int w = 1; w < XT_WARPS; ++w) tot += s_sum[w * XT_HP + h];
        const float pv = __fdividef(e[h], tot);
        if (n < N) pr[h * N + n] = pv;
        if (n < 4 * 64)                           // P[h][n], K-major: k-chunk n / 64, 16-byte unit (n % 64) / 8
          *reinterpret_cast<bf16*>(g_p + (n >> 6) * 2048 + h * 128 + ((((n & 63) >> 3) ^ (h & 7)) << 4) + (n & 7) * 2) = __float2bfloat16(pv);
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- step 4: zhat^T = xhat^T P^T   (A = the xhat tile read MN-major: M = 128 channels = two 64-column chunks)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const int ksteps = L.RA >> 4;
        for (int j = 0; j < ctiles; ++j) {
          const uint64_t ad = umma_desc_sw128(base + (2 * j) * L.pitch, L.pitch, 1024);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t bd = umma_desc_sw128(base + L.off_p + (ks >> 2) * 2048, 16, 1024) + 2 * (ks & 3);
            umma_bf16_ss(tmem + 32 + j * XT_HP, ad + ks * 128, bd, idesc2, ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar2);
      }
      __syncwarp();
    }
    mbar_wait(bar2, phase, abort_flag, p.status_word, ERR_TIMEOUT_ATTN);
    tc_fence_after();
    // ---- step 5: one thread per channel; channel tile j is read by the warps whose TMEM lane quadrants cover it
    for (int j = warp >> 2; j < ctiles; j += XT_WARPS / 4) {
      const int c = j * 128 + (warp & 3) * 32 + lane;
      uint32_t r[16];
      tmem_ld16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 32 + j * XT_HP, r);
      tmem_ld_wait();
      const float g = __ldg(p.gamma + (long long)k * C + c), be = __ldg(p.beta + (long long)k * C + c);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float zh = __uint_as_float(r[h]);
        p.zhat[(kb * H + h) * C + c] = zh;
        p.z[(kb * H + h) * C + c] = __float2bfloat16(fmaf(g, zh, be));
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM and the tiles are rewritten by the next item
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ------------------------------------------------------------------------------------------------ backward on tcgen05
// The exact adjoint with its three contractions on the tensor cores, same xhat tile (bf16 mode, same shape conditions, C <= 384
// for the TMEM budget). Per (fusion, sample), with B1 = [a_h ; w_h] (bf16, [32][C]; a_h = q'_h o gamma, w_h = gz_h o gamma):
//   1. stream the token rows once, xhat = (x - mu) rstd with the SAVED statistics -> bf16 tile [n][c];
//   2. [sa | dp][n][32] = xhat B1^T                      (A = tile K-major)            sa_hn = a_h . xhat_n, dp_hn = w_h . xhat_n
//   3. one thread per token row: D_h = sum_n p_hn dp_hn (block reduction), ds_hn = p_hn (dp_hn - D_h),
//      c1_n = mean_c(dxhat_n), c2_n = mean_c(dxhat_n o xhat_n) in closed form; Q[n][32] = [scale ds | p] as bf16, 64-byte rows;
//   4. da^T[c][16] = sum_n xhat[n][c] Q[n][0..15]        (A = tile MN-major, B = Q MN-major)     da_h = scale sum_n ds_hn xhat_n
//      dxhat^T[c][n] = sum_j B1[j][c] Q[n][j]            (A = B1 MN-major, B = Q K-major), 128 channels at a time, two TMEM buffers
//   5. one thread per channel (so a warp covers 128 contiguous bytes of a token row): dx_n = rstd_n (dxhat_n - c1_n - xhat_n c2_n)
//      added to the stream gradient (read-modify-write); dq'_h = da_h o gamma; dgamma, dbeta by atomics.
constexpr int XB_COLS = 32;
constexpr int XB_D3_STRIDE = 224;     // TMEM columns per dxhat^T buffer: token rows (RA <= 224)

struct XbLayout {
  int RA, pitch, chunks, off_b1, off_q, off_row, off_sums, off_red, off_bar, bytes;
};
static XbLayout xb_layout(int N, int C) {
  XbLayout L;
  L.RA = (N + 15) & ~15;
  L.pitch = L.RA * 128;
  L.chunks = C / 64;
  L.off_b1 = L.chunks * L.pitch;                  // [chunks][32 rows][128 B]
  L.off_q = L.off_b1 + L.chunks * 4096;           // [256 rows][64 B]
  L.off_row = L.off_q + 256 * 64;                 // per token row: c1 | c2 | rstd | -  (float4)
  L.off_sums = L.off_row + 256 * 16;              // A_h[16] | W_h[16]
  L.off_red = L.off_sums + 256;                   // [warp][16]
  L.off_bar = L.off_red + XT_WARPS * 16 * 4;
  L.bytes = L.off_bar + 64 + 1024;
  return L;
}

template <int H, int NV>
__global__ void __launch_bounds__(XT_THREADS, 1)
xfold_tc_bwd_kernel(const XfoldParams p, const XbLayout L) {
  extern __shared__ uint8_t xt_raw[];
  const uint32_t base = (smem_u32(xt_raw) + 1023u) & ~1023u;
  uint8_t* gen = xt_raw + (base - smem_u32(xt_raw));
  uint8_t* g_b1 = gen + L.off_b1;
  uint8_t* g_q = gen + L.off_q;
  float* s_c1 = reinterpret_cast<float*>(gen + L.off_row);     // float4 per token row
  float* s_sums = reinterpret_cast<float*>(gen + L.off_sums);
  float* s_red = reinterpret_cast<float*>(gen + L.off_red);
  const uint32_t bar1 = base + L.off_bar, bar2 = bar1 + 8, bar3 = bar1 + 16;   // bar3, bar3 + 8: the two dxhat buffers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + L.off_bar + 32);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int C = 128 * NV;
  const int N = p.N, mtiles = (N + 127) >> 7;

  if (tid == 0) {
    *abort_flag = 0;
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    mbar_init(bar3, 1);
    mbar_init(bar3 + 8, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < (L.off_row - L.off_b1) / 16; i += XT_THREADS) reinterpret_cast<uint4*>(g_b1)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // TMEM: [sa | dp] tiles in [0, 64), reused by the da tiles once step 3 has read them; dxhat^T buffers at 64 and 64 + 224
  const uint32_t tmem = *tmem_slot;
  const uint32_t tD2 = tmem, tD3 = tmem + 64;
  const uint32_t idesc1 = umma_idesc_bf16(XB_COLS, 0, 0, 128), idesc2 = umma_idesc_bf16(16, 1, 1, 128);
  const uint32_t idesc3 = umma_idesc_bf16(L.RA, 1, 0, 128);
  const float inv_c = 1.0f / (float)C;
  const int items = p.K * p.B;
  uint32_t phase = 0, ph3 = 0;

  for (int item = blockIdx.x; item < items; item += gridDim.x, phase ^= 1u) {
    const int k = item / p.B, b = item - k * p.B;
    const long long kb = item;
    const float* xs = p.x + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C;
    const float* x0 = p.cls + kb * C;
    float* dxs = p.dx + ((long long)p.tok_src[k] * p.B * N + (long long)b * N) * C;
    float* dx0 = p.dx + ((long long)p.cls_src[k] * p.B * N + (long long)b * N) * C;
    const float* qp = p.qp + kb * H * C;
    const float* gz = p.gz + kb * H * C;
    const float* gamma = p.gamma + (long long)k * C;
    if (tid < 32) s_sums[tid] = 0.f;
    __syncthreads();
    // ---- B1 = [a ; w] as bf16 (rows h and 16 + h), and A_h = sum_c a_h[c], W_h = sum_c w_h[c] of the rounded values
    for (int i = tid; i < 2 * H * (C >> 3); i += XT_THREADS) {
      const int which = i / (H * (C >> 3)), r = i - which * (H * (C >> 3));
      const int h = r / (C >> 3), c8 = r - h * (C >> 3);
      const float* src = (which ? gz : qp) + h * C + c8 * 8;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8) + 1);
      uint4 w;
      w.x = pack_bf16(q0.x * g0.x, q0.y * g0.y); w.y = pack_bf16(q0.z * g0.z, q0.w * g0.w);
      w.z = pack_bf16(q1.x * g1.x, q1.y * g1.y); w.w = pack_bf16(q1.z * g1.z, q1.w * g1.w);
      const int j = which * 16 + h;
      *reinterpret_cast<uint4*>(g_b1 + (c8 >> 3) * 4096 + j * 128 + (((c8 & 7) ^ (j & 7)) << 4)) = w;
      const float2 a0 = unpack_bf16_fast(w.x), a1 = unpack_bf16_fast(w.y), a2 = unpack_bf16_fast(w.z), a3 = unpack_bf16_fast(w.w);
      atomicAdd(s_sums + j, ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y)));
    }
    // ---- step 1: token rows -> xhat (bf16) tile with the saved statistics
    for (int r0 = warp; r0 < L.RA; r0 += 4 * XT_WARPS) {
      xt_prefetch_rows<NV>(xs, r0 + 4 * XT_WARPS, N, lane);
      float4 v[4][NV];
      float mu[4], rs[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = r0 + XT_WARPS * j;
        if (n < N) {
          const float4* row = reinterpret_cast<const float4*>(n == 0 ? x0 : xs + (long long)n * C) + lane;
#pragma unroll
          for (int i = 0; i < NV; ++i) v[j][i] = __ldg(row + 32 * i);
          mu[j] = __ldg(p.mean + kb * N + n);
          rs[j] = __ldg(p.rstd + kb * N + n);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = r0 + XT_WARPS * j;
        if (n >= L.RA) continue;                 // warp-uniform
        const uint32_t sw = static_cast<uint32_t>(n & 7);
        uint8_t* d0 = gen + n * 128 + (lane >> 4) * L.pitch + (((static_cast<uint32_t>(lane & 15) >> 1) ^ sw) << 4) + (lane & 1) * 8;
        if (n >= N) {
#pragma unroll
          for (int i = 0; i < NV; ++i) *reinterpret_cast<uint2*>(d0 + 2 * i * L.pitch) = make_uint2(0u, 0u);
          continue;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint2 w;
          w.x = pack_bf16((v[j][i].x - mu[j]) * rs[j], (v[j][i].y - mu[j]) * rs[j]);
          w.y = pack_bf16((v[j][i].z - mu[j]) * rs[j], (v[j][i].w - mu[j]) * rs[j]);
          *reinterpret_cast<uint2*>(d0 + 2 * i * L.pitch) = w;
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- step 2: [sa | dp] = xhat B1^T
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        for (int m = 0; m < mtiles; ++m)
          for (int cc = 0; cc < L.chunks; ++cc) {
            const uint64_t ad = umma_desc_sw128(base + cc * L.pitch + m * 16384, 16, 1024);
            const uint64_t bd = umma_desc_sw128(base + L.off_b1 + cc * 4096, 16, 1024);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem + m * XB_COLS, ad + 2 * kk, bd + 2 * kk, idesc1, (cc | kk) != 0 ? 1u : 0u);
          }
        umma_commit(bar1);
      }
      __syncwarp();
    }
    mbar_wait(bar1, phase, abort_flag, p.status_word, ERR_TIMEOUT_ATTN);
    tc_fence_after();
    // ---- step 3: softmax backward and the per-row coefficients, one thread per token row
    {
      const int n = tid, m = tid >> 7;
      float sa[H], dp[H], pe[H];
      if (m < mtiles) {
        uint32_t r[32];
        tmem_ld32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + m * XB_COLS, r);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < H; ++h) {
          sa[h] = __uint_as_float(r[h]);
          dp[h] = __uint_as_float(r[16 + h]);
          pe[h] = (n < N) ? __ldg(p.probs + (kb * H + h) * N + n) : 0.f;
        }
      } else {
#pragma unroll
        for (int h = 0; h < H; ++h) sa[h] = dp[h] = pe[h] = 0.f;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float part = warp_sum(pe[h] * dp[h]);
        if (lane == 0) s_red[warp * 16 + h] = part;
      }
      __syncthreads();
      if (tid < 256) {
        float c1 = 0.f, c2 = 0.f;
        float sds[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float D = s_red[h];
#pragma unroll
          for (int w = 1; w < XT_WARPS; ++w) D += s_red[w * 16 + h];
          sds[h] = p.scale * pe[h] * (dp[h] - D);
          c1 += sds[h] * s_sums[h] + pe[h] * s_sums[16 + h];
          c2 += sds[h] * sa[h] + pe[h] * dp[h];
        }
        *reinterpret_cast<float4*>(s_c1 + 4 * n) = make_float4(c1 * inv_c, c2 * inv_c, (n < N) ? __ldg(p.rstd + kb * N + n) : 0.f, 0.f);
        // Q[n][0..15] = scale ds, Q[n][16..31] = p: 64-byte rows, SWIZZLE_64B (16-byte unit u of row n at u ^ ((n >> 1) & 3))
        float qv[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          qv[j] = j < H ? sds[j < H ? j : 0] : 0.f;
          qv[16 + j] = j < H ? pe[j < H ? j : 0] : 0.f;
        }
        const uint32_t sw = static_cast<uint32_t>((n >> 1) & 3);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_bf16(qv[8 * u], qv[8 * u + 1]); w.y = pack_bf16(qv[8 * u + 2], qv[8 * u + 3]);
          w.z = pack_bf16(qv[8 * u + 4], qv[8 * u + 5]); w.w = pack_bf16(qv[8 * u + 6], qv[8 * u + 7]);
          *reinterpret_cast<uint4*>(g_q + n * 64 + ((static_cast<uint32_t>(u) ^ sw) << 4)) = w;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- step 4: da^T[c][16] = xhat^T Q[:, 0..15]  and  dxhat^T[c][n] = B1^T Q^T, 128 channels at a time, two TMEM buffers
    auto issue_dxhat = [&](int j) {     // channel tile j -> buffer j & 1
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint64_t ad = umma_desc_sw128(base + L.off_b1 + (2 * j) * 4096, 4096, 1024) + kk * 128;   // B1 read MN-major, K = 32 rows
        const uint64_t bd = umma_desc_layout(base + L.off_q, 16, 512, 4) + 2 * kk;                       // Q read K-major, N = RA rows
        umma_bf16_ss(tD3 + (j & 1) * XB_D3_STRIDE, ad, bd, idesc3, kk != 0 ? 1u : 0u);
      }
      umma_commit(bar3 + 8u * (j & 1));
    };
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const int ksteps = L.RA >> 4;
        for (int j = 0; j < NV; ++j) {
          const uint64_t ad = umma_desc_sw128(base + (2 * j) * L.pitch, L.pitch, 1024);
          const uint64_t bd = umma_desc_layout(base + L.off_q, 16, 512, 4);
          for (int ks = 0; ks < ksteps; ++ks) umma_bf16_ss(tD2 + j * 16, ad + ks * 128, bd + ks * 64, idesc2, ks != 0 ? 1u : 0u);
        }
        umma_commit(bar2);
        issue_dxhat(0);
        if (NV > 1) issue_dxhat(1);
      }
      __syncwarp();
    }
    mbar_wait(bar2, phase, abort_flag, p.status_word, ERR_TIMEOUT_ATTN);
    tc_fence_after();
    // ---- step 5a: dq' = da o gamma, dgamma, dbeta; one thread per channel
    {
      const int j = warp >> 2;
      if (j < NV) {
        const int c = j * 128 + (warp & 3) * 32 + lane;
        uint32_t r[16];
        tmem_ld16(tD2 + (static_cast<uint32_t>((warp & 3) * 32) << 16) + j * 16, r);
        tmem_ld_wait();
        const float g = __ldg(gamma + c);
        float dg = 0.f, db = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float da = __uint_as_float(r[h]);
          p.dqp[(kb * H + h) * C + c] = da * g;
          const float gzh = __ldg(gz + h * C + c);
          dg += da * __ldg(qp + h * C + c) + gzh * __ldg(p.zhat + (kb * H + h) * C + c);
          db += gzh;
        }
        atomicAdd(p.dgamma + (long long)k * C + c, dg);
        atomicAdd(p.dbeta + (long long)k * C + c, db);
      }
    }
    // ---- step 5b: dx, one thread per channel (a warp instruction covers 128 contiguous bytes of a token row); the four
    // warps of a TMEM lane quadrant take every fourth group of 16 token rows
    for (int j = 0; j < NV; ++j) {
      mbar_wait(bar3 + 8u * (j & 1), (ph3 >> (j & 1)) & 1u, abort_flag, p.status_word, ERR_TIMEOUT_ATTN);
      ph3 ^= 1u << (j & 1);
      tc_fence_after();
      const int c = j * 128 + (warp & 3) * 32 + lane;
      const uint8_t* xcol = gen + (c >> 6) * L.pitch + (c & 7) * 2;
      const uint32_t cu = static_cast<uint32_t>(c & 63) >> 3;
      uint32_t xo[8];                              // byte offset of this channel's 16-byte unit in a row with n % 8 == k
#pragma unroll
      for (int kx = 0; kx < 8; ++kx) xo[kx] = (cu ^ static_cast<uint32_t>(kx)) << 4;
      for (int u = warp >> 2; u < (L.RA >> 4); u += XT_WARPS / 4) {
        uint32_t r[16];
        tmem_ld16(tD3 + (j & 1) * XB_D3_STRIDE + (static_cast<uint32_t>((warp & 3) * 32) << 16) + u * 16, r);
        const int n0 = u * 16;
        const uint8_t* xb = xcol + n0 * 128;
        const float4* rcb = reinterpret_cast<const float4*>(s_c1) + n0;
        {   // L2 prefetch of the gradient rows this warp reads in its NEXT round: one 128-byte line per row, one lane each
          const int nn = n0 + 16 * (XT_WARPS / 4) * ((lane >> 4) + 1) + (lane & 15);   // lanes 0-15: next round, 16-31: the one after
#ifndef CAVIT_NO_XF_PREFETCH
          if (nn < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(dxs + (long long)nn * C + (c - lane)));
#endif
          // (the first two rounds of a channel tile are not prefetched: they overlap the previous tile's stores)
        }
        if (p.disjoint && n0 > 0 && n0 + 16 <= N) {
          // sixteen whole token rows of the donor stream: constant row offsets, no predicates (warp-uniform branch)
          float* d = dxs + (long long)n0 * C + c;
          float old[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) old[e] = d[e * C];
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 rc = rcb[e];               // c1 | c2 | rstd | -
            const float xh = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(xb + e * 128 + xo[e & 7])) << 16);
            d[e * C] = old[e] + rc.z * (__uint_as_float(r[e]) - rc.x - xh * rc.y);
          }
        } else {
          float old[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;
            if (p.disjoint && n < N) old[e] = (n == 0 ? dx0 : dxs + (long long)n * C)[c];
          }
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;
            if (n < N) {
              const float4 rc = rcb[e];
              const float xh = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(xb + e * 128 + xo[e & 7])) << 16);
              const float o = rc.z * (__uint_as_float(r[e]) - rc.x - xh * rc.y);
              float* d = (n == 0 ? dx0 : dxs + (long long)n * C) + c;
              if (p.disjoint) *d = old[e] + o;
              else atomicAdd(d, o);
            }
          }
        }
      }
      // Only a tile that is followed by another one in the same TMEM buffer (j + 2 < NV) needs the block to meet here; the
      // other tiles run into the next one without a barrier (ncu: 3.7 barrier-stall cycles per issued instruction before)
      if (j + 2 < NV) {
        tc_fence_before();
        __syncthreads();                         // this dxhat buffer is free again
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) issue_dxhat(j + 2);
          __syncwarp();
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM and the tiles are rewritten by the next item
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// -1 = not decided yet (environment CAVIT_XFOLD_TC, default on); 0 / 1 set by cavit_xfold_tensor_cores()
static std::atomic<int> g_xfold_tc{-1};
static bool xfold_tc_enabled() {
  int v = g_xfold_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("CAVIT_XFOLD_TC");
    v = (e && e[0] == '0') ? 0 : 1;
    g_xfold_tc.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
static bool xfold_tc_fwd_ok(const XfoldParams& p) {
  if (!xfold_tc_enabled() || p.z_lo || p.N > 256 || (p.C % 128) || p.C > 512 || p.H > XT_HP) return false;
  return xt_layout(p.N, p.C).bytes <= 227 * 1024;
}
static bool xfold_tc_bwd_ok(const XfoldParams& p) {
  // the forward must have been the tensor-core one too? No: both variants save the same probabilities / statistics.
  if (!xfold_tc_enabled() || p.use_drop || p.N > XB_D3_STRIDE || (p.C % 128) || p.C > 384 || p.H > 16) return false;
  return xb_layout(p.N, p.C).bytes <= 227 * 1024;
}
template <int H>
static int xfold_tc_bwd_launch(const XfoldParams& p, cudaStream_t st) {
  static_assert(H % 2 == 0 && H <= 6, "C = 64 H must be a multiple of 128 and at most 384");
  constexpr int NV = H / 2;
  const XbLayout L = xb_layout(p.N, p.C);
  static PerDeviceMax cur;
  if (L.bytes > cur.get()) {
    if (cudaFuncSetAttribute(xfold_tc_bwd_kernel<H, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes) != cudaSuccess)
      return fail(CAVIT_E_LAUNCH, "xfold tc bwd smem attribute");
    cur.set(L.bytes);
  }
  const int items = p.K * p.B;
  xfold_tc_bwd_kernel<H, NV><<<items < sm_count() ? items : sm_count(), XT_THREADS, L.bytes, st>>>(p, L);
  count_launch();
  return check_launch("cavit_xfold_bwd");
}

template <int H, int NV>
static int xfold_tc_fwd_launch_nv(const XfoldParams& p, cudaStream_t st) {
  const XtLayout L = xt_layout(p.N, p.C);
  static PerDeviceMax cur;
  if (L.bytes > cur.get()) {
    if (cudaFuncSetAttribute(xfold_tc_fwd_kernel<H, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes) != cudaSuccess)
      return fail(CAVIT_E_LAUNCH, "xfold tc fwd smem attribute");
    cur.set(L.bytes);
  }
  const int items = p.K * p.B;
  xfold_tc_fwd_kernel<H, NV><<<items < sm_count() ? items : sm_count(), XT_THREADS, L.bytes, st>>>(p, L);
  count_launch();
  return check_launch("cavit_xfold_fwd");
}
template <int H>
static int xfold_tc_fwd_launch(const XfoldParams& p, cudaStream_t st) {
  static_assert(H % 2 == 0 && H <= 8, "C = 64 H must be a multiple of 128 and at most 512");
  return xfold_tc_fwd_launch_nv<H, H / 2>(p, st);
}

static int xfold_dispatch(bool bwd, const XfoldParams& p, cudaStream_t st) {
  if (!bwd && xfold_tc_fwd_ok(p)) {
    switch (p.H) {     // C = 64 H and C % 128 == 0: even head counts
      case 2: return xfold_tc_fwd_launch<2>(p, st);
      case 4: return xfold_tc_fwd_launch<4>(p, st);
      case 6: return xfold_tc_fwd_launch<6>(p, st);
      case 8: return xfold_tc_fwd_launch<8>(p, st);
      default: break;
    }
  }
  if (bwd && p.precise == 0 && xfold_tc_bwd_ok(p)) {
    switch (p.H) {
      case 2: return xfold_tc_bwd_launch<2>(p, st);
      case 4: return xfold_tc_bwd_launch<4>(p, st);
      case 6: return xfold_tc_bwd_launch<6>(p, st);
      default: break;
    }
  }
  switch (p.H) {
    case 1: return xfold_launch<1>(bwd, p, st);
    case 2: return xfold_launch<2>(bwd, p, st);
    case 3: return xfold_launch<3>(bwd, p, st);
    case 4: return xfold_launch<4>(bwd, p, st);
    case 6: return xfold_launch<6>(bwd, p, st);
    case 8: return xfold_launch<8>(bwd, p, st);
    case 12: return xfold_launch<12>(bwd, p, st);
    case 16: return xfold_launch<16>(bwd, p, st);
    default: return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_xfold: num_heads %d is not instantiated (1,2,3,4,6,8,12,16)", p.H);
  }
}

// 0/1 expansion of a per-head projection: E[c_out][h*C + c_in] = W[c_out][c_in] if c_out belongs to head h else 0
__global__ void expand_heads_kernel(const bf16* __restrict__ W, bf16* __restrict__ E, int C, int H, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cin = (int)(i % C);
    const long long t = i / C;
    const int h = (int)(t % H);
    const long long row = t / H;                 // g * C + c_out
    const int cout = (int)(row % C);
    E[i] = (cout / 64 == h) ? W[row * C + cin] : __float2bfloat16(0.f);
  }
}
// dW[c_out][c_in] = dE[c_out][head(c_out)*C + c_in]
__global__ void fold_heads_kernel(const float* __restrict__ dE, float* __restrict__ dW, int C, int H, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cin = (int)(i % C);
    const long long row = i / C;
    const int cout = (int)(row % C);
    dW[i] = dE[(row * H + cout / 64) * C + cin];
  }
}

}  // namespace cavit

using namespace cavit;

extern "C" {

static int xfold_fill(XfoldParams& p, const float* x, const float* cls, const float* qp, const float* gamma, const float* beta,
                      int32_t K, int32_t B, int32_t N, int32_t C, int32_t H, const int32_t* cls_src, const int32_t* tok_src,
                      float scale, float eps, float p_drop, const uint64_t* seed_dev, uint32_t site) {
  if (K <= 0 || K > XF_MAX_FUSIONS || B <= 0 || N <= 0 || H <= 0 || C != H * 64 || C > 1024)
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_xfold: K=%d B=%d N=%d C=%d H=%d", K, B, N, C, H);
  p.x = x; p.cls = cls; p.qp = qp; p.gamma = gamma; p.beta = beta;
  p.B = B; p.N = N; p.C = C; p.H = H; p.K = K;
  for (int i = 0; i < K; ++i) { p.cls_src[i] = cls_src[i]; p.tok_src[i] = tok_src[i]; }
  p.scale = scale; p.eps = eps;
  p.drop.seed = reinterpret_cast<const unsigned long long*>(seed_dev);
  p.drop.site = site; p.drop.thresh = 0; p.drop.inv_keep = 1.f;
  p.use_drop = 0;
  p.status_word = status_word();
  if (!p.status_word) return fail(CAVIT_E_DEVICE, "cavit_xfold: no device status word");
  p.disjoint = 1;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < i; ++j)
      if (tok_src[i] == tok_src[j]) p.disjoint = 0;
  // With attention dropout the weights no longer sum to one, so beta / bv stop being constants of the fold:
  // callers use the unfolded path (cavit_xattn_*) when dropout is active.
  if (p_drop > 0.f) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_xfold: attention dropout needs the unfolded path");
  return CAVIT_OK;
}

int cavit_xfold_tensor_cores(int on) {
  const int prev = xfold_tc_enabled() ? 1 : 0;
  if (on >= 0) g_xfold_tc.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

int64_t cavit_xfold_scratch_floats(int32_t K, int32_t B, int32_t N, int32_t H) {
  const int HP = (H + 3) & ~3;
  return (int64_t)K * B * N * (4 * HP + 4);
}

int cavit_xfold_fwd(const float* x, const float* cls, const float* qp, const float* gamma, const float* beta, float* zhat,
                    void* z, void* z_lo, float* probs, float* mean, float* rstd, float* scratch, int32_t K, int32_t B, int32_t N,
                    int32_t C, int32_t H, const int32_t* cls_src, const int32_t* tok_src, float scale, float eps, float p_drop,
                    const uint64_t* seed_dev, uint32_t site, void* stream) {
  if (!x || !cls || !qp || !gamma || !beta || !zhat || !z || !probs || !mean || !rstd || !scratch || !cls_src || !tok_src)
    return fail(CAVIT_E_BADARG, "cavit_xfold_fwd: null pointer");
  XfoldParams p = {};
  int rc = xfold_fill(p, x, cls, qp, gamma, beta, K, B, N, C, H, cls_src, tok_src, scale, eps, p_drop, seed_dev, site);
  if (rc) return rc;
  p.zhat = zhat; p.z = reinterpret_cast<bf16*>(z); p.z_lo = reinterpret_cast<bf16*>(z_lo); p.probs = probs; p.mean = mean; p.rstd = rstd; p.scratch = scratch;
  return xfold_dispatch(false, p, as_stream(stream));
}

int cavit_xfold_bwd(const float* x, const float* cls, const float* qp, const float* gamma, const float* zhat,
                    const float* probs, const float* mean, const float* rstd, const float* gz, float* scratch, float* dx,
                    float* dqp, float* dgamma, float* dbeta, int32_t K, int32_t B, int32_t N, int32_t C, int32_t H,
                    const int32_t* cls_src, const int32_t* tok_src, float scale, float p_drop, const uint64_t* seed_dev,
                    uint32_t site, int32_t exact_fp32, void* stream) {
  if (!x || !cls || !qp || !gamma || !zhat || !probs || !mean || !rstd || !gz || !scratch || !dx || !dqp || !dgamma ||
      !dbeta || !cls_src || !tok_src)
    return fail(CAVIT_E_BADARG, "cavit_xfold_bwd: null pointer");
  XfoldParams p = {};
  int rc = xfold_fill(p, x, cls, qp, gamma, gamma, K, B, N, C, H, cls_src, tok_src, scale, 0.f, p_drop, seed_dev, site);
  if (rc) return rc;
  p.zhat = const_cast<float*>(zhat); p.probs = const_cast<float*>(probs);
  p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.gz = gz; p.scratch = scratch; p.dx = dx; p.dqp = dqp; p.dgamma = dgamma; p.dbeta = dbeta;
  p.precise = exact_fp32;
  return xfold_dispatch(true, p, as_stream(stream));
}

int cavit_expand_heads(const void* W, void* E, int32_t groups, int32_t C, int32_t H, void* stream) {
  if (!W || !E || groups <= 0 || C != H * 64) return fail(CAVIT_E_BADARG, "cavit_expand_heads: bad args");
  const long long total = (long long)groups * C * H * C;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  expand_heads_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const bf16*>(W), reinterpret_cast<bf16*>(E), C, H,
                                                                  total);
  count_launch();
  return check_launch("cavit_expand_heads");
}

int cavit_fold_heads(const float* dE, float* dW, int32_t groups, int32_t C, int32_t H, void* stream) {
  if (!dE || !dW || groups <= 0 || C != H * 64) return fail(CAVIT_E_BADARG, "cavit_fold_heads: bad args");
  const long long total = (long long)groups * C * C;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  fold_heads_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(dE, dW, C, H, total);
  count_launch();
  return check_launch("cavit_fold_heads");
}

}  // extern "C"
