// cavit-sm100 — small memory-bound kernels around the GEMM / attention cores: patch gather,
// CLS row, embedding parameter gradients, casts, bias-gradient column sums, row gathers and the
// classification tail (last head Linear + mean over streams + cross-entropy, fwd and bwd).
#include "common.cuh"
#include "internal.h"

namespace cavit {

// ------------------------------------------------------------------------------------------ cast
__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16(src[(n4 << 2) + threadIdx.x]);
}

// fp32-tolerance mode: x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits between them). The two planes are
// what the tensor-core GEMM consumes as (A_hi, A_lo) / (B_hi, B_lo) in its 3-pass split-operand mode (gemm.cu).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  const float2 h = unpack_bf16_fast(hi);
  lo = pack_bf16(a - h.x, b - h.y);
}
__global__ void cast_split_kernel(const float* __restrict__ src, bf16* __restrict__ hi, bf16* __restrict__ lo, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    uint2 h, l;
    split_bf16x2(v.x, v.y, h.x, l.x);
    split_bf16x2(v.z, v.w, h.y, l.y);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long j = (n4 << 2) + threadIdx.x;
    const bf16 h = __float2bfloat16(src[j]);
    hi[j] = h;
    lo[j] = __float2bfloat16(src[j] - __bfloat162float(h));
  }
}

// fp32-tolerance mode GELU (exact erf form, nn.GELU default, model_cross.py:24) on fp32 pre-activations:
//   forward   h = gelu(u)          -> split planes (the fc2 / head operand) and, optionally, an fp32 copy
//   backward  du = dh * gelu'(u)   -> split planes (the fc1 dgrad / wgrad operand)
__global__ void gelu_split_kernel(const float* __restrict__ u, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                  float* __restrict__ h32, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(u) + i);
    float4 g;
    g.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f));
    g.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f));
    g.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752f));
    g.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752f));
    if (hi) {
      uint2 h, l;
      split_bf16x2(g.x, g.y, h.x, l.x);
      split_bf16x2(g.z, g.w, h.y, l.y);
      reinterpret_cast<uint2*>(hi)[i] = h;
      reinterpret_cast<uint2*>(lo)[i] = l;
    }
    if (h32) reinterpret_cast<float4*>(h32)[i] = g;
  }
}
__device__ __forceinline__ float gelu_grad_exact(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
__global__ void gelu_bwd_split_kernel(const float* __restrict__ dh, const float* __restrict__ u, bf16* __restrict__ hi,
                                      bf16* __restrict__ lo, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(dh) + i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(u) + i);
    uint2 h, l;
    split_bf16x2(d.x * gelu_grad_exact(v.x), d.y * gelu_grad_exact(v.y), h.x, l.x);
    split_bf16x2(d.z * gelu_grad_exact(v.z), d.w * gelu_grad_exact(v.w), h.y, l.y);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// ------------------------------------------------------------------------------------------ patchify
// out[m][b*Np + t][f] = img[b][m][0][di*dp + a][hi*hp + bb][wi*wp + c]
//   t = (hi*Wn + wi)*Dn + di   (d fastest),   f = (a*hp + bb)*wp + c      
// One thread produces two consecutive features (one 4-byte bf16x2 store).
__global__ void patchify_kernel(const float* __restrict__ img, bf16* __restrict__ out, bf16* __restrict__ out_lo, int B, int M,
                                int D, int H, int W, int dp, int hp, int wp, int sample_major) {
  const int Dn = D / dp, Hn = H / hp, Wn = W / wp;
  const long long Np = (long long)Dn * Hn * Wn;
  const int P = dp * hp * wp;
  const long long pairs = (long long)M * B * Np * P / 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    long long e = i * 2;
    const int f = (int)(e % P);
    e /= P;
    const long long t = e % Np;
    e /= Np;
    int b, m;
    if (sample_major) { m = (int)(e % M); b = (int)(e / M); }
    else { b = (int)(e % B); m = (int)(e / B); }
    const int di = (int)(t % Dn), wi = (int)((t / Dn) % Wn), hi = (int)(t / ((long long)Dn * Wn));
    const float* vol = img + ((long long)b * M + m) * ((long long)D * H * W);
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ff = f + j;
      const int c = ff % wp, bb = (ff / wp) % hp, a = ff / (wp * hp);
      v[j] = __ldg(vol + ((long long)(di * dp + a) * H + (hi * hp + bb)) * W + (wi * wp + c));
    }
    if (out_lo) {   // fp32-tolerance mode: hi + lo planes of the patch rows
      uint32_t h, l;
      split_bf16x2(v[0], v[1], h, l);
      reinterpret_cast<uint32_t*>(out)[i] = h;
      reinterpret_cast<uint32_t*>(out_lo)[i] = l;
    } else {
      reinterpret_cast<uint32_t*>(out)[i] = pack_bf16(v[0], v[1]);
    }
  }
}

// Vector variant: when the run of features that is contiguous in BOTH the patch row and the volume (wp, times hp when the
// patch spans the whole W axis, times dp when it also spans H) is a multiple of 4, one thread moves 4 features: one index
// decode, one 16-byte load (when aligned), one 8-byte store. cfg2 (224x224x1 slices, 16x16x1 patches): runs of 16 floats.
__global__ void patchify_vec4_kernel(const float* __restrict__ img, bf16* __restrict__ out, bf16* __restrict__ out_lo, int B,
                                     int M, int D, int H, int W, int dp, int hp, int wp, int sample_major) {
  const int Dn = D / dp, Hn = H / hp, Wn = W / wp;
  const long long Np = (long long)Dn * Hn * Wn;
  const int P = dp * hp * wp;
  const long long quads = (long long)M * B * Np * P / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += stride) {
    long long e = i * 4;
    const int f = (int)(e % P);
    e /= P;
    const long long t = e % Np;
    e /= Np;
    int b, m;
    if (sample_major) { m = (int)(e % M); b = (int)(e / M); }
    else { b = (int)(e % B); m = (int)(e / B); }
    const int di = (int)(t % Dn), wi = (int)((t / Dn) % Wn), hi = (int)(t / ((long long)Dn * Wn));
    const int c = f % wp, bb = (f / wp) % hp, a = f / (wp * hp);
    const float* src = img + ((long long)b * M + m) * ((long long)D * H * W) +
                       ((long long)(di * dp + a) * H + (hi * hp + bb)) * W + (wi * wp + c);
    float4 v;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
    }
    uint2 o;
    if (out_lo) {
      uint2 l;
      split_bf16x2(v.x, v.y, o.x, l.x);
      split_bf16x2(v.z, v.w, o.y, l.y);
      reinterpret_cast<uint2*>(out_lo)[i] = l;
    } else {
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
    }
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// tokens[m][b*N][c] = cls[c] + pos[0][c]
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ tokens,
                                int MB, int N, int C) {
  const long long total = (long long)MB * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long mb = i / C;
    tokens[mb * (long long)N * C + c] = cls[c] + pos[c];
  }
}

// dpos[n][c] = sum over (m,b) of dtokens[(m*B+b)*N + n][c]; dcls = dpos[0] (cls and pos[0] feed the same row).
// grid = (column chunks, slices of the (m, b) axis): every slice adds its partial sum atomically into the zeroed outputs, so
// the 310 MB read at cfg2 is spread over the whole GPU instead of 148 blocks that each walk 1024 rows serially.
__global__ void embed_param_grads_kernel(const float* __restrict__ dtok, float* __restrict__ dpos, float* __restrict__ dcls,
                                         int MB, int N, int C) {
  const int C4 = C >> 2;
  const long long total = (long long)N * C4;
  const int per = (MB + gridDim.y - 1) / gridDim.y;
  const int mb0 = blockIdx.y * per, mb1 = min(MB, mb0 + per);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* p = reinterpret_cast<const float4*>(dtok) + i;
#pragma unroll 4
    for (int mb = mb0; mb < mb1; ++mb) {
      const float4 v = __ldg(p + (long long)mb * total);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* o = dpos + i * 4;
    atomicAdd(o + 0, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    if (i < C4) {
      float* q = dcls + i * 4;
      atomicAdd(q + 0, acc.x); atomicAdd(q + 1, acc.y); atomicAdd(q + 2, acc.z); atomicAdd(q + 3, acc.w);
    }
  }
}

// ------------------------------------------------------------------------------------------ colsum
// out[g][c] = sum_r x[g][r][c]. grid = (ceil(C/256), groups, row_slices); block = 256 (8 warps). Each lane
// owns 8 consecutive columns (one 16-byte load per row), each warp strides over the rows of its slice;
// the 8 warps are combined through shared memory and the row slices with fp32 atomics into a
// zero-initialised out.
template <bool LO>   // LO: rows are hi + lo pairs (fp32-tolerance mode); compile-time so that the plain loop keeps its unrolling
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x_lo, long long ldx, long long gs, int rows, int C,
                   float* __restrict__ out, long long out_gs) {
  __shared__ float sm[8][256 + 8];
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 256 + lane * 8;
  const int slices = gridDim.z;
  const int rows_per = (rows + slices - 1) / slices;
  const int r0 = blockIdx.z * rows_per, r1 = min(rows, r0 + rows_per);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c < C) {  // C % 8 == 0 is checked by the launcher
    const bf16* p = x + (long long)g * gs + c;
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + (long long)r * ldx));
      float2 f;
      f = unpack_bf16(v.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack_bf16(v.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack_bf16(v.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack_bf16(v.w); acc[6] += f.x; acc[7] += f.y;
      if (LO) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(x_lo + (long long)g * gs + c + (long long)r * ldx));
        f = unpack_bf16(w.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16(w.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16(w.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16(w.w); acc[6] += f.x; acc[7] += f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < C) {
    float s_ = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s_ += sm[w][threadIdx.x];
    atomicAdd(out + (long long)g * out_gs + cc, s_);
  }
}

// ------------------------------------------------------------------------------------------ gather rows
struct GatherIdx {
  int use;            // 0: group g uses group strides directly
  int src[16], dst[16];   // else: group g reads from src group src[g] and writes dst group dst[g]
};
__global__ void gather_rows_f32_kernel(float* __restrict__ src, long long srs, long long sgs, float* __restrict__ dst,
                                       long long drs, long long dgs, int rows, int C, int groups, int accumulate,
                                       int zero_src, const GatherIdx gi) {
  const int C4 = C >> 2;
  const long long total = (long long)groups * rows * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long e = i / C4;
    const int r = (int)(e % rows);
    const int g = (int)(e / rows);
    const int gs_ = gi.use ? gi.src[g] : g, gd = gi.use ? gi.dst[g] : g;
    float4* sp = reinterpret_cast<float4*>(src + (long long)gs_ * sgs + (long long)r * srs) + c4;
    const float4 v = *sp;
    if (zero_src) *sp = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* d = reinterpret_cast<float4*>(dst + (long long)gd * dgs + (long long)r * drs) + c4;
    if (accumulate) {
      float4 o = *d;
      o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      *d = o;
    } else {
      *d = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ elementwise
__global__ void add_bf16_f32_kernel(const float* a, const bf16* __restrict__ b, float* out, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(a)[i];
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(b) + i);
    const float2 lo = unpack_bf16(q.x), hi = unpack_bf16(q.y);
    v.x += lo.x; v.y += lo.y; v.z += hi.x; v.w += hi.y;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

__global__ void gelu_bwd_bf16_kernel(const bf16* __restrict__ dh, const bf16* __restrict__ u, bf16* __restrict__ du, long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 d = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(dh) + i));
    const float2 x = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(u) + i));
    reinterpret_cast<uint32_t*>(du)[i] = pack_bf16(d.x * gelu_erf_grad(x.x), d.y * gelu_erf_grad(x.y));
  }
}

// out[s*Np + t][:] = in[s*(Np+1) + 1 + t][:]   (16-byte chunks)
__global__ void compact_patch_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int S, int Np, int C8) {
  const long long total = (long long)S * Np * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const long long row = i / C8;
    const long long s_ = row / Np, t = row % Np;
    out[i] = __ldg(in + (s_ * (Np + 1) + 1 + t) * C8 + c);
  }
}

// ------------------------------------------------------------------------------------------ dropout
// mode 0: out_f32 = m*a_f32            mode 1: out_bf16 = m*a_bf16
// mode 2: out_f32 = a_f32 + m*b_bf16   mode 3: out_bf16 = bf16(m*a_f32)      mode 4: out_u8 = keep
// m = keep / (1 - p); two elements per thread (one hash).
__global__ void dropout_kernel(int mode, const void* a, const void* b, void* out, long long n, DropCfg d) {
  const unsigned long long seed = *d.seed;
  const long long pairs = (n + 1) >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
    const float2 m = drop_mult2(d, seed, (uint64_t)i * 2);
    const long long e = i * 2;
    const bool two = e + 1 < n;
    if (mode == 0) {
      const float* x = reinterpret_cast<const float*>(a);
      float* o = reinterpret_cast<float*>(out);
      if (two) {
        const float2 v = *reinterpret_cast<const float2*>(x + e);
        *reinterpret_cast<float2*>(o + e) = make_float2(v.x * m.x, v.y * m.y);
      } else {
        o[e] = x[e] * m.x;
      }
    } else if (mode == 1) {
      const bf16* x = reinterpret_cast<const bf16*>(a);
      bf16* o = reinterpret_cast<bf16*>(out);
      if (two) {
        const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + e));
        *reinterpret_cast<uint32_t*>(o + e) = pack_bf16(v.x * m.x, v.y * m.y);
      } else {
        o[e] = __float2bfloat16(__bfloat162float(x[e]) * m.x);
      }
    } else if (mode == 2) {
      const float* r = reinterpret_cast<const float*>(a);
      const bf16* x = reinterpret_cast<const bf16*>(b);
      float* o = reinterpret_cast<float*>(out);
      if (two) {
        const float2 rv = *reinterpret_cast<const float2*>(r + e);
        const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + e));
        *reinterpret_cast<float2*>(o + e) = make_float2(rv.x + v.x * m.x, rv.y + v.y * m.y);
      } else {
        o[e] = r[e] + __bfloat162float(x[e]) * m.x;
      }
    } else if (mode == 3) {
      const float* x = reinterpret_cast<const float*>(a);
      bf16* o = reinterpret_cast<bf16*>(out);
      if (two) {
        const float2 v = *reinterpret_cast<const float2*>(x + e);
        *reinterpret_cast<uint32_t*>(o + e) = pack_bf16(v.x * m.x, v.y * m.y);
      } else {
        o[e] = __float2bfloat16(x[e] * m.x);
      }
    } else {
      uint8_t* o = reinterpret_cast<uint8_t*>(out);
      o[e] = m.x != 0.f;
      if (two) o[e + 1] = m.y != 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------ classification tail
constexpr int HEAD_MAX_CLASSES = 8;

// grid = B, block = 256. logits[b][k] = (1/M) sum_m D_m( b2[m][k] + sum_f h[m][b][f] W2[m][k][f] ), where D_m is the
// (optional) dropout the reference applies to every head's logits (model_cross.py:182) before the mean.
__device__ __forceinline__ float head_ld(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float head_ld(const float* p) { return *p; }
__device__ __forceinline__ void head_st(bf16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void head_st(float* p, float v) { *p = v; }

// T = bf16 (bf16 mode: the hidden activations are the bf16 output of the head GEMM) or float (fp32-tolerance mode)
template <typename T>
__global__ void head_logits_kernel(const T* __restrict__ h, const float* __restrict__ W2, const float* __restrict__ b2,
                                   float* __restrict__ logits, int M, int B, int F, int classes, DropCfg d, int use_drop) {
  __shared__ float red[HEAD_MAX_CLASSES][8];
  __shared__ float total[HEAD_MAX_CLASSES];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < HEAD_MAX_CLASSES) total[threadIdx.x] = 0.f;
  const unsigned long long seed = use_drop ? *d.seed : 0ull;
  for (int m = 0; m < M; ++m) {
    float acc[HEAD_MAX_CLASSES];
#pragma unroll
    for (int k = 0; k < HEAD_MAX_CLASSES; ++k) acc[k] = 0.f;
    const T* hr = h + ((long long)m * B + b) * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const float hv = head_ld(hr + f);
#pragma unroll
      for (int k = 0; k < HEAD_MAX_CLASSES; ++k)
        if (k < classes) acc[k] += hv * __ldg(W2 + ((long long)m * classes + k) * F + f);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < HEAD_MAX_CLASSES; ++k) {
      const float s_ = warp_sum(acc[k]);
      if (lane == 0) red[k][warp] = s_;
    }
    __syncthreads();
    if (threadIdx.x < classes) {
      float s_ = b2[m * classes + threadIdx.x];
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s_ += red[threadIdx.x][w];
      if (use_drop) s_ *= drop_mult(d, seed, ((uint64_t)m * B + b) * classes + threadIdx.x);
      total[threadIdx.x] += s_;
    }
  }
  __syncthreads();
  if (threadIdx.x < classes) logits[b * classes + threadIdx.x] = total[threadIdx.x] / (float)M;
}

// single block: loss = mean_b CE(logits[b], label[b]) with label smoothing
__global__ void ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ loss,
                               int B, int classes, float smoothing) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* z = logits + b * classes;
    float mx = z[0];
    for (int k = 1; k < classes; ++k) mx = fmaxf(mx, z[k]);
    float se = 0.f, sz = 0.f;
    for (int k = 0; k < classes; ++k) {
      se += expf(z[k] - mx);
      sz += z[k];
    }
    const float lse = mx + logf(se);
    const float nll = lse - z[labels[b]];
    const float smooth = lse - sz / (float)classes;
    acc += (1.f - smoothing) * nll + smoothing * smooth;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    loss[0] = s / (float)B;
  }
}

__device__ __forceinline__ void dlogits_row(const float* z, long long label, int classes, float smoothing, float scale,
                                            float* dz) {
  float mx = z[0];
  for (int k = 1; k < classes; ++k) mx = fmaxf(mx, z[k]);
  float se = 0.f;
  for (int k = 0; k < classes; ++k) se += expf(z[k] - mx);
  for (int k = 0; k < classes; ++k) {
    const float p = expf(z[k] - mx) / se;
    const float tgt = (1.f - smoothing) * (k == label ? 1.f : 0.f) + smoothing / (float)classes;
    dz[k] = (p - tgt) * scale;
  }
}

// grid = (B, M): dh[m][b][f] = sum_k dz[b][k] W2[m][k][f],  dz = dlogits / M
template <typename T>
__global__ void head_dh_kernel(const float* __restrict__ W2, const long long* __restrict__ labels,
                               const float* __restrict__ logits, float scale, const float* __restrict__ scale_dev,
                               T* __restrict__ dh, int M, int B, int F, int classes, float smoothing, DropCfg d,
                               int use_drop) {
  const int b = blockIdx.x, m = blockIdx.y;
  if (scale_dev) scale *= __ldg(scale_dev);
  float dz[HEAD_MAX_CLASSES];
  dlogits_row(logits + b * classes, labels[b], classes, smoothing, scale / ((float)B * (float)M), dz);
  if (use_drop) {
    const unsigned long long seed = *d.seed;
    for (int k = 0; k < classes; ++k) dz[k] *= drop_mult(d, seed, ((uint64_t)m * B + b) * classes + k);
  }
  T* o = dh + ((long long)m * B + b) * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < classes; ++k) s += dz[k] * __ldg(W2 + ((long long)m * classes + k) * F + f);
    head_st(o + f, s);
  }
}

// grid = (ceil(F/256), M): dW2[m][k][f] = sum_b dz[b][k] h[m][b][f];  block (0, m) also writes db2.
template <typename T>
__global__ void head_dw_kernel(const T* __restrict__ h, const long long* __restrict__ labels,
                               const float* __restrict__ logits, float scale, const float* __restrict__ scale_dev,
                               float* __restrict__ dW2, float* __restrict__ db2, int M, int B, int F, int classes,
                               float smoothing, DropCfg d, int use_drop) {
  const int m = blockIdx.y;
  if (scale_dev) scale *= __ldg(scale_dev);
  const unsigned long long seed = use_drop ? *d.seed : 0ull;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[HEAD_MAX_CLASSES], accb[HEAD_MAX_CLASSES];
#pragma unroll
  for (int k = 0; k < HEAD_MAX_CLASSES; ++k) acc[k] = accb[k] = 0.f;
  // d(logits) of every sample is the same for all columns f: the block computes the B rows once into shared memory
  // (each thread used to redo the softmax of every sample: B dependent exp chains per thread, 0.24 ms for a [256, 2] problem)
  extern __shared__ float s_dz[];   // [B][classes]
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float dz[HEAD_MAX_CLASSES];
    dlogits_row(logits + b * classes, labels[b], classes, smoothing, scale / ((float)B * (float)M), dz);
    for (int k = 0; k < classes; ++k) {
      float v = dz[k];
      if (use_drop) v *= drop_mult(d, seed, ((uint64_t)m * B + b) * classes + k);
      s_dz[b * classes + k] = v;
    }
  }
  __syncthreads();
#pragma unroll 4
  for (int b = 0; b < B; ++b) {
    const float hv = (f < F) ? head_ld(h + ((long long)m * B + b) * F + f) : 0.f;
#pragma unroll
    for (int k = 0; k < HEAD_MAX_CLASSES; ++k)
      if (k < classes) {
        const float dzk = s_dz[b * classes + k];
        acc[k] += dzk * hv;
        accb[k] += dzk;
      }
  }
  if (f < F)
    for (int k = 0; k < classes; ++k) dW2[((long long)m * classes + k) * F + f] = acc[k];
  if (blockIdx.x == 0 && threadIdx.x < classes) db2[m * classes + threadIdx.x] = accb[threadIdx.x];
}

// ------------------------------------------------------------------------------------------ fused Adam
// torch.optim.Adam(lr, betas, eps, weight_decay) semantics (L2-style decay folded into the gradient, bias-corrected
// moments; /root/reference/model_cross.py:276-279) over the flat parameter / gradient slabs, one pass: 16 B read +
// 12 B written per parameter (+ 2 B for the refreshed bf16 GEMM operand copy, which removes the per-step cast kernel).
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, bf16* __restrict__ pb, long long n, float lr_c1, float b1, float b2,
                                 float eps, float wd, float rsqrt_c2, float gscale) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define CAVIT_ADAM1(F)                                              \
    {                                                               \
      const float gg = gv.F * gscale + wd * pv.F;                   \
      mv.F = b1 * mv.F + (1.f - b1) * gg;                           \
      vv.F = b2 * vv.F + (1.f - b2) * gg * gg;                      \
      pv.F -= lr_c1 * mv.F / (sqrtf(vv.F) * rsqrt_c2 + eps);        \
    }
    CAVIT_ADAM1(x) CAVIT_ADAM1(y) CAVIT_ADAM1(z) CAVIT_ADAM1(w)
#undef CAVIT_ADAM1
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb) {
      uint2 o;
      o.x = pack_bf16(pv.x, pv.y);
      o.y = pack_bf16(pv.z, pv.w);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
}

static int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cavit

using namespace cavit;

static int make_drop(float p, const uint64_t* seed_dev, uint32_t site, DropCfg* d) {
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

template <typename T>
static int head_fwd_launch(const T* h, const float* W2, const float* b2, const int64_t* labels, float* logits, float* loss,
                           int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing, float p_drop,
                           const uint64_t* seed_dev, uint32_t site, void* stream) {
  if (!h || !W2 || !b2 || !labels || !logits || !loss) return fail(CAVIT_E_BADARG, "cavit_head_loss_fwd: null pointer");
  if (classes < 1 || classes > HEAD_MAX_CLASSES) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "num_classes=%d (max %d)", classes, HEAD_MAX_CLASSES);
  cudaStream_t st = as_stream(stream);
  DropCfg d;
  const int use = make_drop(p_drop, seed_dev, site, &d);
  if (use < 0) return fail(CAVIT_E_BADARG, "cavit_head_loss_fwd: bad dropout arguments");
  head_logits_kernel<T><<<B, 256, 0, st>>>(h, W2, b2, logits, M, B, F, classes, d, use);
  ce_loss_kernel<<<1, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), loss, B, classes, smoothing);
  count_launch(2);
  return check_launch("cavit_head_loss_fwd");
}

template <typename T>
static int head_bwd_launch(const T* h, const float* W2, const int64_t* labels, const float* logits, float loss_scale,
                           const float* loss_scale_dev, T* dh, float* dW2, float* db2, int32_t M, int32_t B, int32_t F,
                           int32_t classes, float smoothing, float p_drop, const uint64_t* seed_dev, uint32_t site,
                           void* stream) {
  if (!h || !W2 || !labels || !logits || !dh || !dW2 || !db2) return fail(CAVIT_E_BADARG, "cavit_head_loss_bwd: null pointer");
  if (classes < 1 || classes > HEAD_MAX_CLASSES) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "num_classes=%d", classes);
  cudaStream_t st = as_stream(stream);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  DropCfg d;
  const int use = make_drop(p_drop, seed_dev, site, &d);
  if (use < 0) return fail(CAVIT_E_BADARG, "cavit_head_loss_bwd: bad dropout arguments");
  head_dh_kernel<T><<<dim3(B, M), 256, 0, st>>>(W2, lab, logits, loss_scale, loss_scale_dev, dh, M, B, F, classes, smoothing, d, use);
  if ((size_t)B * classes * sizeof(float) > 96 * 1024) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_head_loss_bwd: batch %d too large", B);
  static PerDeviceFlag dw_attr;
  if (dw_attr.unset()) {
    cudaFuncSetAttribute(head_dw_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    dw_attr.set();
  }
  head_dw_kernel<T><<<dim3((F + 255) / 256, M), 256, (size_t)B * classes * sizeof(float), st>>>(
      h, lab, logits, loss_scale, loss_scale_dev, dW2, db2, M, B, F, classes, smoothing, d, use);
  count_launch(2);
  return check_launch("cavit_head_loss_bwd");
}

extern "C" {

int cavit_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return fail(CAVIT_E_BADARG, "cavit_cast_bf16: bad args");
  if (n == 0) return CAVIT_OK;
  cast_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(src, reinterpret_cast<bf16*>(dst), n);
  count_launch();
  return check_launch("cavit_cast_bf16");
}

static int patchify_launch(const float* img, void* patches, void* patches_lo, int32_t B, int32_t M, int32_t D, int32_t H,
                           int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t sample_major, void* stream) {
  if (!img || !patches) return fail(CAVIT_E_BADARG, "cavit_patchify: null pointer");
  if (dp <= 0 || hp <= 0 || wp <= 0 || D % dp || H % hp || W % wp)
    return fail(CAVIT_E_BADARG, "image dimensions must be divisible by the patch size");
  const long long P = (long long)dp * hp * wp;
  if (P % 2) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_patchify: patch_dim must be even");
  const long long pairs = (long long)M * B * (D / dp) * (H / hp) * (W / wp) * P / 2;
  long long run = wp;          // features contiguous in both the patch row and the volume
  if (wp == W) { run *= hp; if (hp == H) run *= dp; }
  if (run % 4 == 0) {
    patchify_vec4_kernel<<<grid_for(pairs / 2, 256), 256, 0, as_stream(stream)>>>(
        img, reinterpret_cast<bf16*>(patches), reinterpret_cast<bf16*>(patches_lo), B, M, D, H, W, dp, hp, wp, sample_major);
    count_launch();
    return check_launch("cavit_patchify");
  }
  patchify_kernel<<<grid_for(pairs, 256), 256, 0, as_stream(stream)>>>(
      img, reinterpret_cast<bf16*>(patches), reinterpret_cast<bf16*>(patches_lo), B, M, D, H, W, dp, hp, wp, sample_major);
  count_launch();
  return check_launch("cavit_patchify");
}

int cavit_patchify(const float* img, void* patches, int32_t B, int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp,
                   int32_t hp, int32_t wp, int32_t sample_major, void* stream) {
  return patchify_launch(img, patches, nullptr, B, M, D, H, W, dp, hp, wp, sample_major, stream);
}

int cavit_patchify_split(const float* img, void* patches_hi, void* patches_lo, int32_t B, int32_t M, int32_t D, int32_t H,
                         int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t sample_major, void* stream) {
  if (!patches_lo) return fail(CAVIT_E_BADARG, "cavit_patchify_split: null lo plane");
  return patchify_launch(img, patches_hi, patches_lo, B, M, D, H, W, dp, hp, wp, sample_major, stream);
}

int cavit_cast_split(const float* src, void* hi, void* lo, int64_t n, void* stream) {
  if (!src || !hi || !lo || n < 0) return fail(CAVIT_E_BADARG, "cavit_cast_split: bad args");
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7))
    return fail(CAVIT_E_BADARG, "cavit_cast_split: misaligned buffer");
  if (n == 0) return CAVIT_OK;
  cast_split_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(src, reinterpret_cast<bf16*>(hi),
                                                                              reinterpret_cast<bf16*>(lo), n);
  count_launch();
  return check_launch("cavit_cast_split");
}

int cavit_gelu_split(const float* u, void* h_hi, void* h_lo, float* h_f32, int64_t n, void* stream) {
  if (!u || n <= 0 || (n % 4) || (!h_hi && !h_f32) || ((h_hi == nullptr) != (h_lo == nullptr)))
    return fail(CAVIT_E_BADARG, "cavit_gelu_split: bad args (n % 4 == 0; hi and lo together, or h_f32)");
  gelu_split_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(u, reinterpret_cast<bf16*>(h_hi),
                                                                          reinterpret_cast<bf16*>(h_lo), h_f32, n / 4);
  count_launch();
  return check_launch("cavit_gelu_split");
}

int cavit_gelu_bwd_split(const float* dh, const float* u, void* du_hi, void* du_lo, int64_t n, void* stream) {
  if (!dh || !u || !du_hi || !du_lo || n <= 0 || (n % 4)) return fail(CAVIT_E_BADARG, "cavit_gelu_bwd_split: bad args");
  gelu_bwd_split_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(dh, u, reinterpret_cast<bf16*>(du_hi),
                                                                              reinterpret_cast<bf16*>(du_lo), n / 4);
  count_launch();
  return check_launch("cavit_gelu_bwd_split");
}

int cavit_cls_rows(const float* cls, const float* pos, float* tokens, int32_t M, int32_t B, int32_t N, int32_t C,
                   void* stream) {
  if (!cls || !pos || !tokens) return fail(CAVIT_E_BADARG, "cavit_cls_rows: null pointer");
  cls_rows_kernel<<<grid_for((long long)M * B * C, 256), 256, 0, as_stream(stream)>>>(cls, pos, tokens, M * B, N, C);
  count_launch();
  return check_launch("cavit_cls_rows");
}

int cavit_embed_param_grads(const float* dtokens, float* dpos, float* dcls, int32_t M, int32_t B, int32_t N, int32_t C,
                            void* stream) {
  if (!dtokens || !dpos || !dcls || (C % 4)) return fail(CAVIT_E_BADARG, "cavit_embed_param_grads: bad args");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(dpos, 0, sizeof(float) * (size_t)N * C, st);
  cudaMemsetAsync(dcls, 0, sizeof(float) * (size_t)C, st);
  const int gx = grid_for((long long)N * C / 4, 128);
  int gy = (sm_count() * 8 + gx - 1) / gx;     // ~8 blocks per SM in total
  if (gy > M * B) gy = M * B;
  if (gy < 1) gy = 1;
  embed_param_grads_kernel<<<dim3(gx, gy), 128, 0, st>>>(dtokens, dpos, dcls, M * B, N, C);
  count_launch();
  return check_launch("cavit_embed_param_grads");
}

static int colsum_launch(const void* x, const void* x_lo, int64_t ldx, int64_t x_gs, int32_t rows, int32_t C, int32_t groups,
                         float* out, int64_t out_gs, void* stream) {
  if (!x || !out || rows <= 0 || C <= 0 || (C % 8) || (ldx % 8) || (x_gs % 8) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(x_lo) & 15))
    return fail(CAVIT_E_BADARG, "cavit_colsum_bf16: bad args (C, ld must be multiples of 8; 16-byte aligned base)");
  cudaStream_t st = as_stream(stream);
  if (out_gs == C) {
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C * groups, st);
  } else {
    for (int g = 0; g < groups; ++g) cudaMemsetAsync(out + (long long)g * out_gs, 0, sizeof(float) * C, st);
  }
  const int cblocks = (C + 255) / 256;
  int slices = (sm_count() * 4) / (cblocks * groups);  // ~4 blocks per SM in total
  const int max_slices = (rows + 63) / 64;             // at least 8 rows per warp
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  dim3 grid(cblocks, groups, slices);
  if (x_lo)
    colsum_bf16_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(x_lo), ldx, x_gs,
                                                   rows, C, out, out_gs);
  else
    colsum_bf16_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), nullptr, ldx, x_gs, rows, C, out, out_gs);
  count_launch();
  return check_launch("cavit_colsum_bf16");
}

int cavit_colsum_bf16(const void* x, int64_t ldx, int64_t x_gs, int32_t rows, int32_t C, int32_t groups, float* out,
                      int64_t out_gs, void* stream) {
  return colsum_launch(x, nullptr, ldx, x_gs, rows, C, groups, out, out_gs, stream);
}

int cavit_colsum_split(const void* x_hi, const void* x_lo, int64_t ldx, int64_t x_gs, int32_t rows, int32_t C, int32_t groups,
                       float* out, int64_t out_gs, void* stream) {
  if (!x_lo) return fail(CAVIT_E_BADARG, "cavit_colsum_split: null lo plane");
  return colsum_launch(x_hi, x_lo, ldx, x_gs, rows, C, groups, out, out_gs, stream);
}

int cavit_gather_rows_f32(float* src, int64_t srs, int64_t sgs, float* dst, int64_t drs, int64_t dgs, int32_t rows,
                          int32_t C, int32_t groups, int32_t accumulate, int32_t zero_src, void* stream) {
  if (!src || !dst || (C % 4) || (srs % 4) || (sgs % 4) || (drs % 4) || (dgs % 4))
    return fail(CAVIT_E_BADARG, "cavit_gather_rows_f32: bad args");
  GatherIdx gi{};
  gather_rows_f32_kernel<<<grid_for((long long)groups * rows * C / 4, 256), 256, 0, as_stream(stream)>>>(
      src, srs, sgs, dst, drs, dgs, rows, C, groups, accumulate, zero_src, gi);
  count_launch();
  return check_launch("cavit_gather_rows_f32");
}

int cavit_gather_rows_f32_indexed(float* src, int64_t srs, int64_t sgs, const int32_t* src_group, float* dst, int64_t drs,
                                  int64_t dgs, const int32_t* dst_group, int32_t rows, int32_t C, int32_t groups,
                                  int32_t accumulate, int32_t zero_src, void* stream) {
  if (!src || !dst || (C % 4) || (srs % 4) || (sgs % 4) || (drs % 4) || (dgs % 4))
    return fail(CAVIT_E_BADARG, "cavit_gather_rows_f32_indexed: bad args");
  if (groups < 1 || groups > 16) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gather_rows_f32_indexed: groups=%d (max 16)", groups);
  GatherIdx gi{};
  gi.use = 1;
  for (int g = 0; g < groups; ++g) {
    gi.src[g] = src_group ? src_group[g] : g;
    gi.dst[g] = dst_group ? dst_group[g] : g;
    if (gi.src[g] < 0 || gi.dst[g] < 0) return fail(CAVIT_E_BADARG, "cavit_gather_rows_f32_indexed: negative group index");
  }
  for (int g = 0; g < groups; ++g)      // read-modify-write / move needs distinct targets inside one launch
    for (int h = g + 1; h < groups; ++h)
      if (gi.dst[g] == gi.dst[h] || (zero_src && gi.src[g] == gi.src[h]))
        return fail(CAVIT_E_BADARG, "cavit_gather_rows_f32_indexed: duplicate group index");
  gather_rows_f32_kernel<<<grid_for((long long)groups * rows * C / 4, 256), 256, 0, as_stream(stream)>>>(
      src, srs, sgs, dst, drs, dgs, rows, C, groups, accumulate, zero_src, gi);
  count_launch();
  return check_launch("cavit_gather_rows_f32_indexed");
}

int cavit_add_bf16_f32(const float* a, const void* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n <= 0 || (n % 4)) return fail(CAVIT_E_BADARG, "cavit_add_bf16_f32: bad args");
  add_bf16_f32_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(a, reinterpret_cast<const bf16*>(b), out, n / 4);
  count_launch();
  return check_launch("cavit_add_bf16_f32");
}

int cavit_gelu_bwd_bf16(const void* dh, const void* u, void* du, int64_t n, void* stream) {
  if (!dh || !u || !du || n <= 0 || (n % 2)) return fail(CAVIT_E_BADARG, "cavit_gelu_bwd_bf16: bad args");
  gelu_bwd_bf16_kernel<<<grid_for(n / 2, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const bf16*>(dh), reinterpret_cast<const bf16*>(u), reinterpret_cast<bf16*>(du), n / 2);
  count_launch();
  return check_launch("cavit_gelu_bwd_bf16");
}

int cavit_compact_patch_rows_bf16(const void* in, void* out, int32_t S, int32_t Np, int32_t C, void* stream) {
  if (!in || !out || S <= 0 || Np <= 0 || C <= 0 || (C % 8)) return fail(CAVIT_E_BADARG, "cavit_compact_patch_rows_bf16: bad args");
  compact_patch_rows_kernel<<<grid_for((long long)S * Np * (C / 8), 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), S, Np, C / 8);
  count_launch();
  return check_launch("cavit_compact_patch_rows_bf16");
}

int cavit_dropout(int32_t mode, const void* a, const void* b, void* out, int64_t n, float p, const uint64_t* seed_dev,
                  uint32_t site, void* stream) {
  if (mode < 0 || mode > 4 || !out || n <= 0 || (mode != 4 && !a) || (mode == 2 && !b))
    return fail(CAVIT_E_BADARG, "cavit_dropout: bad args");
  DropCfg d;
  const int use = make_drop(p, seed_dev, site, &d);
  if (use < 0 || !seed_dev) return fail(CAVIT_E_BADARG, "cavit_dropout: p must be in [0, 1) and seed_dev non-null");
  dropout_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, as_stream(stream)>>>(mode, a, b, out, n, d);
  count_launch();
  return check_launch("cavit_dropout");
}

int cavit_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(CAVIT_E_BADARG, "cavit_adam_step: null pointer");
  if (n <= 0 || (n & 3) || step < 1) return fail(CAVIT_E_BADARG, "cavit_adam_step: n must be a positive multiple of 4, step >= 1");
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(CAVIT_E_BADARG, "cavit_adam_step: buffers must be 16-byte aligned");
  const double c1 = 1.0 - pow((double)beta1, (double)step), c2 = 1.0 - pow((double)beta2, (double)step);
  adam_step_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, reinterpret_cast<bf16*>(params_bf16), n, (float)((double)lr / c1), beta1, beta2, eps,
      weight_decay, (float)(1.0 / sqrt(c2)), grad_scale);
  count_launch();
  return check_launch("cavit_adam_step");
}

int cavit_head_loss_fwd(const void* h, const float* W2, const float* b2, const int64_t* labels, float* logits, float* loss,
                        int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing, float p_drop,
                        const uint64_t* seed_dev, uint32_t site, void* stream) {
  return head_fwd_launch<bf16>(reinterpret_cast<const bf16*>(h), W2, b2, labels, logits, loss, M, B, F, classes, smoothing,
                               p_drop, seed_dev, site, stream);
}

int cavit_head_loss_bwd(const void* h, const float* W2, const int64_t* labels, const float* logits, float loss_scale,
                        const float* loss_scale_dev, void* dh, float* dW2, float* db2, int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing,
                        float p_drop, const uint64_t* seed_dev, uint32_t site, void* stream) {
  return head_bwd_launch<bf16>(reinterpret_cast<const bf16*>(h), W2, labels, logits, loss_scale, loss_scale_dev,
                               reinterpret_cast<bf16*>(dh), dW2, db2, M, B, F, classes, smoothing, p_drop, seed_dev, site, stream);
}

/* fp32-tolerance mode: the hidden activations of the heads and their gradient are fp32 */
int cavit_head_loss_fwd_f32(const float* h, const float* W2, const float* b2, const int64_t* labels, float* logits,
                            float* loss, int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing, void* stream) {
  return head_fwd_launch<float>(h, W2, b2, labels, logits, loss, M, B, F, classes, smoothing, 0.f, nullptr, 0, stream);
}

int cavit_head_loss_bwd_f32(const float* h, const float* W2, const int64_t* labels, const float* logits, float loss_scale,
                            const float* loss_scale_dev, float* dh, float* dW2, float* db2, int32_t M, int32_t B, int32_t F,
                            int32_t classes, float smoothing, void* stream) {
  return head_bwd_launch<float>(h, W2, labels, logits, loss_scale, loss_scale_dev, dh, dW2, db2, M, B, F, classes, smoothing,
                                0.f, nullptr, 0, stream);
}

}  // extern "C"
