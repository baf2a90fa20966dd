// cavit-sm100 — token assembly and classification tails of the CNN-stem encoders
// (`ViT`, /root/reference/model.py:79-105,218-286 and `ViT3D`, /root/reference/modelv2.py:187-241).
//
// The CNN stems stay outside the hot path; what they hand over is an fp32 feature
// map, and everything from there to the loss runs here:
//   * channel-major feature maps -> token rows (flatten + cat + transpose + CLS + positional add of
//     modelv2.py:203-224) as one tiled transpose, and its adjoint;
//   * Conv3d(kernel = stride = grid) patch embedding of model.py:84,95-100 as a pure permutation
//     into bf16 patch rows (token order (a b c) of the conv output, feature order (cin i0 i1 i2) of
//     the flattened conv weight) feeding the embedding GEMM, and its adjoint for d(feature map);
//   * the single-logit BCE-with-logits tail of model.py:224,279-286.
// All of it is HBM-bound index work: one pass over the bytes, coalesced on the fp32 side.
#include "common.cuh"
#include "internal.h"

namespace cavit {

// ------------------------------------------------------------------------------ channels <-> tokens
// X[b][off + s][c] = feat[b][c][s] + pos[off + s][c]          (32 x 32 tiles through shared memory)
// grid = (ceil(S/32), ceil(C/32), B), block = (32, 8)
__global__ void tokens_from_channels_kernel(const float* __restrict__ feat, const float* __restrict__ pos,
                                            float* __restrict__ X, int C, int S, int N, int off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* f = feat + (long long)b * C * S;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, s = s0 + threadIdx.x;
    if (c < C && s < S) tile[j][threadIdx.x] = __ldg(f + (long long)c * S + s);
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int s = s0 + j, c = c0 + threadIdx.x;
    if (c < C && s < S) {
      const long long n = off + s;
      X[((long long)b * N + n) * C + c] = tile[threadIdx.x][j] + __ldg(pos + n * C + c);
    }
  }
}

// dfeat[b][c][s] = dX[b][off + s][c]
__global__ void tokens_to_channels_kernel(const float* __restrict__ dX, float* __restrict__ dfeat, int C, int S, int N,
                                          int off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int s = s0 + j, c = c0 + threadIdx.x;
    if (c < C && s < S) tile[j][threadIdx.x] = __ldg(dX + ((long long)b * N + off + s) * C + c);
  }
  __syncthreads();
  float* f = dfeat + (long long)b * C * S;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, s = s0 + threadIdx.x;
    if (c < C && s < S) f[(long long)c * S + s] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------------------ token mean (ViT3D without a CLS token)
// out[b][c] = mean_n x[b][n][c]   (modelv2.py:233-235: `x.mean(dim=1)`).  grid = (ceil(C/128), B), block 128: thread = column.
__global__ void token_mean_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = x + (long long)b * N * C + c;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int n = 0;
  for (; n + 4 <= N; n += 4) {
    a0 += __ldg(p + (long long)n * C);
    a1 += __ldg(p + (long long)(n + 1) * C);
    a2 += __ldg(p + (long long)(n + 2) * C);
    a3 += __ldg(p + (long long)(n + 3) * C);
  }
  for (; n < N; ++n) a0 += __ldg(p + (long long)n * C);
  out[(long long)b * C + c] = ((a0 + a1) + (a2 + a3)) / (float)N;
}
// dx[b][n][c] = dmean[b][c] / N
__global__ void token_mean_bwd_kernel(const float* __restrict__ dmean, float* __restrict__ dx, int B, int N, int C) {
  const int C4 = C >> 2;
  const long long total = (long long)B * N * C4;
  const float inv = 1.0f / (float)N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long b = i / ((long long)N * C4);
    float4 v = __ldg(reinterpret_cast<const float4*>(dmean + b * C) + c4);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    reinterpret_cast<float4*>(dx)[i] = v;
  }
}

// ------------------------------------------------------------------------------ Conv3d(k = stride) patch rows
// feat: fp32 [M*B][Cin][A][Bd][Cd] (sample index m*B + b: the stem runs modality-major);
// rows: bf16 [(b*M + m)*Np + t][f],  t = (a'*Bn + b')*Cn + c',  f = ((cin*g0 + i0)*g1 + i1)*g2 + i2.
// Thread = one fp32 input element (coalesced along the innermost axis); FWD gathers into rows, else scatters back.
struct ConvPatchGeom {
  int M, B, Cin, A, Bd, Cd, g0, g1, g2;
};
template <bool FWD>
__global__ void conv_patch_rows_kernel(float* __restrict__ feat, bf16* __restrict__ rows, ConvPatchGeom q) {
  const int An = q.A / q.g0, Bn = q.Bd / q.g1, Cn = q.Cd / q.g2;
  const long long Np = (long long)An * Bn * Cn;
  const int P = q.Cin * q.g0 * q.g1 * q.g2;
  const long long total = (long long)q.M * q.B * q.Cin * q.A * q.Bd * q.Cd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long e = i;
    const int z = (int)(e % q.Cd); e /= q.Cd;
    const int y = (int)(e % q.Bd); e /= q.Bd;
    const int x = (int)(e % q.A); e /= q.A;
    const int cin = (int)(e % q.Cin); e /= q.Cin;
    const int b = (int)(e % q.B);
    const int m = (int)(e / q.B);
    const int a1 = x / q.g0, i0 = x - a1 * q.g0;
    const int b1 = y / q.g1, i1 = y - b1 * q.g1;
    const int c1 = z / q.g2, i2 = z - c1 * q.g2;
    if (a1 >= An || b1 >= Bn || c1 >= Cn) {   // remainder voxels a strided conv never reads
      if (!FWD) feat[i] = 0.f;
      continue;
    }
    const long long t = ((long long)a1 * Bn + b1) * Cn + c1;
    const int f = ((cin * q.g0 + i0) * q.g1 + i1) * q.g2 + i2;
    const long long r = ((long long)b * q.M + m) * Np + t;
    if (FWD) rows[r * P + f] = __float2bfloat16(feat[i]);
    else feat[i] = __bfloat162float(rows[r * P + f]);
  }
}

// ------------------------------------------------------------------------------ BCE-with-logits tail
// z[b] = x[b].w + b0;  loss = mean_b( max(z,0) - z*y + log1p(exp(-|z|)) )       (model.py:224,279-286)
// single block of 256 threads (B is a batch of volumes: at most a few thousand)
__global__ void bce_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b0,
                                    const float* __restrict__ y, float* __restrict__ logits, float* __restrict__ loss, int B,
                                    int C) {
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int b = warp; b < B; b += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[(long long)b * C + c] * __ldg(w + c);
    s = warp_sum(s) + b0[0];
    if (lane == 0) {
      logits[b] = s;
      if (y) acc += fmaxf(s, 0.f) - s * y[b] + log1pf(expf(-fabsf(s)));
    }
  }
  if (!y) return;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    loss[0] = t / (float)B;
  }
}

// dz[b] = scale * (sigmoid(z[b]) - y[b]) / B;  dx[b][c] = dz[b] w[c];  dw[c] = sum_b dz[b] x[b][c];  db = sum_b dz[b]
// grid = ceil(C/256), thread = one column
__global__ void bce_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y,
                                    const float* __restrict__ logits, float scale, const float* __restrict__ scale_dev,
                                    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int B, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (scale_dev) scale *= __ldg(scale_dev);
  scale /= (float)B;
  const float wc = c < C ? w[c] : 0.f;
  float aw = 0.f, ab = 0.f;
  for (int b = 0; b < B; ++b) {
    const float z = logits[b];
    const float dz = scale * (1.f / (1.f + expf(-z)) - y[b]);
    ab += dz;
    if (c < C) {
      aw += dz * x[(long long)b * C + c];
      dx[(long long)b * C + c] = dz * wc;
    }
  }
  if (c < C) dw[c] = aw;
  if (c == 0) db[0] = ab;
}

static int grid_cap(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_tokens_from_channels(const float* feat, const float* cls, const float* pos, float* tokens, int32_t B, int32_t C,
                               int32_t S, int32_t has_cls, void* stream) {
  if (!feat || !pos || !tokens || (has_cls && !cls)) return fail(CAVIT_E_BADARG, "cavit_tokens_from_channels: null pointer");
  if (B <= 0 || C <= 0 || S <= 0) return fail(CAVIT_E_BADARG, "cavit_tokens_from_channels: empty");
  if (B > 65535) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_tokens_from_channels: B=%d", B);
  const int off = has_cls ? 1 : 0, N = S + off;
  cudaStream_t st = as_stream(stream);
  tokens_from_channels_kernel<<<dim3((S + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st>>>(feat, pos, tokens, C, S, N, off);
  count_launch();
  if (has_cls) {
    int rc = cavit_cls_rows(cls, pos, tokens, 1, B, N, C, stream);
    if (rc) return rc;
  }
  return check_launch("cavit_tokens_from_channels");
}

int cavit_tokens_to_channels(const float* dtokens, float* dfeat, int32_t B, int32_t C, int32_t S, int32_t has_cls,
                             void* stream) {
  if (!dtokens || !dfeat) return fail(CAVIT_E_BADARG, "cavit_tokens_to_channels: null pointer");
  if (B <= 0 || C <= 0 || S <= 0 || B > 65535) return fail(CAVIT_E_BADARG, "cavit_tokens_to_channels: bad extents");
  const int off = has_cls ? 1 : 0, N = S + off;
  tokens_to_channels_kernel<<<dim3((S + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, as_stream(stream)>>>(dtokens, dfeat, C, S,
                                                                                                        N, off);
  count_launch();
  return check_launch("cavit_tokens_to_channels");
}

static int conv_patch_check(const char* who, const void* a, const void* b, int M, int B, int Cin, int A, int Bd, int Cd, int g0,
                            int g1, int g2) {
  if (!a || !b) return fail(CAVIT_E_BADARG, "%s: null pointer", who);
  if (M <= 0 || B <= 0 || Cin <= 0 || g0 <= 0 || g1 <= 0 || g2 <= 0 || A < g0 || Bd < g1 || Cd < g2)
    return fail(CAVIT_E_BADARG, "%s: bad geometry", who);
  return CAVIT_OK;
}

int cavit_conv_patch_rows(const float* feat, void* rows_bf16, int32_t M, int32_t B, int32_t Cin, int32_t A, int32_t Bd,
                          int32_t Cd, int32_t g0, int32_t g1, int32_t g2, void* stream) {
  int rc = conv_patch_check("cavit_conv_patch_rows", feat, rows_bf16, M, B, Cin, A, Bd, Cd, g0, g1, g2);
  if (rc) return rc;
  ConvPatchGeom q{M, B, Cin, A, Bd, Cd, g0, g1, g2};
  const long long total = (long long)M * B * Cin * A * Bd * Cd;
  conv_patch_rows_kernel<true><<<grid_cap(total, 256), 256, 0, as_stream(stream)>>>(const_cast<float*>(feat),
                                                                                   reinterpret_cast<bf16*>(rows_bf16), q);
  count_launch();
  return check_launch("cavit_conv_patch_rows");
}

int cavit_conv_patch_rows_bwd(const void* drows_bf16, float* dfeat, int32_t M, int32_t B, int32_t Cin, int32_t A, int32_t Bd,
                              int32_t Cd, int32_t g0, int32_t g1, int32_t g2, void* stream) {
  int rc = conv_patch_check("cavit_conv_patch_rows_bwd", drows_bf16, dfeat, M, B, Cin, A, Bd, Cd, g0, g1, g2);
  if (rc) return rc;
  ConvPatchGeom q{M, B, Cin, A, Bd, Cd, g0, g1, g2};
  const long long total = (long long)M * B * Cin * A * Bd * Cd;
  conv_patch_rows_kernel<false><<<grid_cap(total, 256), 256, 0, as_stream(stream)>>>(
      dfeat, reinterpret_cast<bf16*>(const_cast<void*>(drows_bf16)), q);
  count_launch();
  return check_launch("cavit_conv_patch_rows_bwd");
}

int cavit_token_mean_fwd(const float* x, float* out, int32_t B, int32_t N, int32_t C, void* stream) {
  if (!x || !out || B <= 0 || N <= 0 || C <= 0 || B > 65535) return fail(CAVIT_E_BADARG, "cavit_token_mean_fwd: bad args");
  token_mean_fwd_kernel<<<dim3((C + 127) / 128, B), 128, 0, as_stream(stream)>>>(x, out, N, C);
  count_launch();
  return check_launch("cavit_token_mean_fwd");
}

int cavit_token_mean_bwd(const float* dmean, float* dx, int32_t B, int32_t N, int32_t C, void* stream) {
  if (!dmean || !dx || B <= 0 || N <= 0 || C <= 0 || (C % 4)) return fail(CAVIT_E_BADARG, "cavit_token_mean_bwd: bad args");
  token_mean_bwd_kernel<<<grid_cap((long long)B * N * C / 4, 256), 256, 0, as_stream(stream)>>>(dmean, dx, B, N, C);
  count_launch();
  return check_launch("cavit_token_mean_bwd");
}

int cavit_bce_head_fwd(const float* x, const float* w, const float* b0, const float* targets, float* logits, float* loss,
                       int32_t B, int32_t C, void* stream) {
  if (!x || !w || !b0 || !logits || (targets && !loss)) return fail(CAVIT_E_BADARG, "cavit_bce_head_fwd: null pointer");
  if (B <= 0 || C <= 0) return fail(CAVIT_E_BADARG, "cavit_bce_head_fwd: empty");
  bce_head_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(x, w, b0, targets, logits, loss, B, C);
  count_launch();
  return check_launch("cavit_bce_head_fwd");
}

int cavit_bce_head_bwd(const float* x, const float* w, const float* targets, const float* logits, float loss_scale,
                       const float* loss_scale_dev, float* dx, float* dw, float* db, int32_t B, int32_t C, void* stream) {
  if (!x || !w || !targets || !logits || !dx || !dw || !db) return fail(CAVIT_E_BADARG, "cavit_bce_head_bwd: null pointer");
  if (B <= 0 || C <= 0) return fail(CAVIT_E_BADARG, "cavit_bce_head_bwd: empty");
  bce_head_bwd_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(x, w, targets, logits, loss_scale, loss_scale_dev, dx, dw,
                                                                      db, B, C);
  count_launch();
  return check_launch("cavit_bce_head_bwd");
}

}  // extern "C"
