// cavit-sm100 — K-GEMM: persistent, warp-specialised tcgen05 GEMM with fused epilogues.
//
//   D[g][m][n] = sum_k A_g(m,k) * B_g(n,k)      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM, 320 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, 128 x BN x 16 per instruction) + TMEM alloc
//   warps 2..9  epilogue       (tcgen05.ld -> smem transpose -> bias / GELU / residual / ... -> coalesced global)
// Two TMEM accumulator stages (2 x BN fp32 columns) let the epilogue of tile i overlap the
// mainloop of tile i+1. Tiles are walked in a static persistent schedule (tile = cta + i*grid).
// Operands may be K-major or MN-major (the latter is what dgrad's W and wgrad's dY / X are), so
// no transposed copies of weights or activations are ever materialised.
//
// Replaces the reference's cuBLAS `addmm`/`mm` call sites (SURVEY.md §2.2 G1-G4, X2, B*).
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

struct GemmDev {
  int M, N, K, groups;
  int a_mn, b_mn, epi, out_fp32, accumulate, embed_np, split_k, kb_per_split;
  int passes;   // 1: D = A B^T.  3 (fp32-tolerance mode): operands are bf16 hi + lo pairs, D = Ah Bh^T + Al Bh^T + Ah Bl^T
  void* out; long long ldo, out_gs;
  const float* bias; long long bias_gs;
  const float* resid; long long ldr, resid_gs;
  void* aux; long long ldaux, aux_gs;
  int tiles_m, tiles_n;
  int vec_ok;  // all epilogue pointers / leading dimensions allow 16-byte (fp32) / 8-byte (bf16) vector access
  int* status;
};

// CTAS = 1: one CTA computes a 128 x BN tile. CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a
// 256 x BN tile; each CTA stages its own 128 rows of A and HALF of the B tile, so the L2 -> SM traffic per FLOP drops
// by a third at BN = 256 (the K = 384 GEMMs of cfg2 run into the L2 fabric limit with 128 x 256 tiles) and the
// shared-memory ring gets deeper for the same footprint.
// EW = epilogue warps per TMEM lane quadrant. With two (each taking half of the tile's columns) the bias / GELU epilogues
// of the 256-wide pair tiles are bounded by the dependent-instruction latency of 2 warps per scheduler (ncu: 52 % issue
// utilisation at 31 % tensor activity); four warps per quadrant (64 columns each, no register prefetch) double the
// latency hiding, but measured no gain once the bias vector is prefetched ahead of the accumulator wait, and it costs
// an operand stage (shared memory) and registers (96 / thread at 576 threads): all epilogues use two.
template <int BN, int CTAS, int EPI>
struct GemmCfg {
// The GELU / GELU' epilogues run ~660 warp instructions per 32-column chunk (one MUFU + a degree-7 polynomial per element) and
// are bounded by issue + dependent-instruction latency: FOUR warps per quadrant (16 epilogue warps, four per scheduler, 64
// columns each) hide that latency better than two — same-box A/B round 2: fc1+GELU 0.388 -> 0.359 ms, fc2-dgrad.GELU' 0.418 ->
// 0.389 ms, cfg2 step 29.78 -> 29.17 ms. (Round 1 measured no gain from this; the difference is the branch-free batched fast
// path, which fits 90 registers per thread.) The plain epilogues keep two (CAVIT_EPI_EW_OTHER for experiments).
#ifndef CAVIT_GELU_EW
#define CAVIT_GELU_EW 4
#endif
#ifndef CAVIT_EPI_EW_OTHER
#define CAVIT_EPI_EW_OTHER 2
#endif
  // (every warp takes whole 32-column chunks: 192-wide tiles split three ways, 128 / 256-wide tiles four ways)
  static constexpr int EW_WANT = (EPI == CAVIT_EPI_BIAS_GELU || EPI == CAVIT_EPI_GELU_BWD) ? CAVIT_GELU_EW : CAVIT_EPI_EW_OTHER;
  static constexpr int EW = (EW_WANT == 4 && BN == 192) ? 3 : EW_WANT;
  static_assert(BN % (EW * 32) == 0, "epilogue warps must own whole 32-column chunks");
  static constexpr int EPI_WARPS = 4 * EW;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;   // TMA warp, MMA warp, epilogue warps
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // as many operand stages as fit beside the epilogue's transposition stages in 227 KB (EW = 2: 8 / 6 / 6 stages for pair tiles
  // of 128 / 192 / 256 columns, 6 / 4 / 4 for single-CTA tiles; EW = 4: 6 / 5 / 5 and 5 / 4 / 3)
  static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 - 256 - EPI_WARPS * 4096;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 4096 /*epilogue transpose stages*/;
};

// ---------------------------------------------------------------------------------------- epilogue
// tcgen05.ld hands every thread one ROW of a 32x32 fp32 accumulator chunk; storing that directly
// would make each warp store touch 32 different rows (32 sectors per instruction). Each chunk is
// therefore transposed through a 4 KB XOR-swizzled shared-memory stage (conflict-free both ways) so
// that afterwards each group of 8 lanes owns one 128-byte row segment: residual / aux loads and all
// stores are fully coalesced (4 rows x 128 B per warp instruction). The loads a chunk needs
// (residual, positional embedding, GELU pre-activation) are issued one chunk AHEAD into registers
// (for the first chunk of a tile even before the accumulator is ready), so that the epilogue
// streams HBM instead of waiting on it. Eight epilogue warps (two per TMEM lane quadrant, each
// taking half of the tile's columns) keep this off the critical path of the short-K mainloops.
// Output modes of the fast path (compile-time): bf16 store, fp32 store, fp32 split-K reduction.
enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_RED = 2 };

__device__ __forceinline__ void stg_v2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void stg_v4f(void* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(void* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(a) : "memory");
}
__device__ __forceinline__ float4 ldg_v4f(const void* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(__cvta_generic_to_global(p))
               : "memory");
  return v;
}
__device__ __forceinline__ uint2 ldg_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(__cvta_generic_to_global(p)) : "memory");
  return v;
}

// Per-lane addressing of the transposed domain, computed once per tile: lane (rsub = lane/8,
// c4 = lane%8) owns columns [c4*4, c4*4+4) of rows rsub, rsub+4, ... of the warp's 32-row slab.
struct EpiLane {
  char* out;          // &out[g][row0 + rsub][0]
  const char* resid;  // &resid[g][row0 + rsub][0]
  char* aux;          // &aux[g][row0 + rsub][0]
  const float* bias;  // &bias[g][0]
  long long out_step, resid_step, aux_step;  // byte strides for 4 rows
  int rows_left;      // rows of this lane's 8 that are inside M: iteration `it` is valid iff it*4 < rows_left
  // EPI_EMBED: GEMM row r = (sequence s, patch t) is written to token row s*(np+1) + 1 + t (row 0 of every sequence is
  // the CLS token, written by cavit_cls_rows) and gets positional row 1 + t added (model_cross.py:194-197)
  long long embed_row;   // first GEMM row of this lane (row0 + rsub)
  int embed_np;
  char* out_base;        // &out[g][0][0]
  const char* pos_base;  // &pos[0][0]
  long long ldo_bytes, ldr_bytes;
};

template <int EPI>
struct EpiPre {
  float4 r[(EPI == CAVIT_EPI_BIAS_RESID) ? 8 : 1];
  uint2 a[(EPI == CAVIT_EPI_GELU_BWD || EPI == CAVIT_EPI_RELU_BWD) ? 8 : 1];
};

template <int EPI>
__device__ __forceinline__ void epi_prefetch(const EpiLane& L, int col, EpiPre<EPI>& pre) {
  if (EPI != CAVIT_EPI_BIAS_RESID && EPI != CAVIT_EPI_GELU_BWD && EPI != CAVIT_EPI_RELU_BWD) return;
#pragma unroll
  for (int it = 0; it < 8; ++it) {   // full 32-row slabs only (partial slabs take the generic path): no guards, no branches
    if (EPI == CAVIT_EPI_BIAS_RESID) pre.r[it] = ldg_v4f(L.resid + it * L.resid_step + col * 4);
    else pre.a[it] = ldg_v2(L.aux + it * L.aux_step + col * 2);
  }
}

// GELU / GELU' of NP pairs evaluated as ONE interleaved stream: every polynomial step runs over all pairs before the next
// step, so each warp carries NP independent dependency chains. The epilogue warps (two per scheduler) were stalled on
// fixed-latency dependencies ('wait' 0.92 and 'short scoreboard' 0.77 cycles per issued instruction, ncu round 1) because
// every row group was a separate predicated block: one chain of 7 dependent FFMA2 + MUFU at a time.
// NEG: r = -R (all coefficients negated at compile time), which lets gelu() finish with one FFMA2 (see gelu_batch).
template <int NP, bool NEG>
__device__ __forceinline__ void gelu_terms_batch(const float2 (&u)[NP], float2 (&e)[NP], float2 (&r)[NP], float2 (&ac)[NP]) {
  constexpr float sg = NEG ? -1.0f : 1.0f;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    ac[j] = make_float2(fminf(fabsf(u[j].x), 5.0f), fminf(fabsf(u[j].y), 5.0f));
    const float2 earg = mul2(mul2(u[j], u[j]), splat2(-0.72134752044448170f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j].x) : "f"(earg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j].y) : "f"(earg.y));
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(ac[j], splat2(sg * -2.501152098e-05f), splat2(sg * 5.606001891e-04f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * -5.354749036e-03f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * 2.881915092e-02f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * -9.833444611e-02f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * 2.302476772e-01f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * -3.942042539e-01f));
#pragma unroll
  for (int j = 0; j < NP; ++j) r[j] = fma2(r[j], ac[j], splat2(sg * 4.997907545e-01f));
}
template <int NP>
__device__ __forceinline__ void gelu_batch(float2 (&u)[NP]) {   // u -> gelu(u), same formula as gelu_pair (common.cuh)
  float2 e[NP], r[NP], ac[NP];
  gelu_terms_batch<NP, true>(u, e, r, ac);
#pragma unroll
  for (int j = 0; j < NP; ++j)   // max(u, 0) - a e R with r = -R; a = min(|u|, 5) instead of |u|: beyond 5 the term is < 2e-6 either way
    u[j] = fma2(ac[j], mul2(e[j], r[j]), make_float2(fmaxf(u[j].x, 0.f), fmaxf(u[j].y, 0.f)));
}
template <int NP>
__device__ __forceinline__ void gelu_grad_batch(float2 (&u)[NP]) {   // u -> gelu'(u), same formula as gelu_grad_pair
  float2 e[NP], r[NP], ac[NP];
  gelu_terms_batch<NP, false>(u, e, r, ac);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float2 t = mul2(e[j], fma2(ac[j], splat2(-0.3989422804014327f), r[j]));
    u[j] = make_float2(u[j].x < 0.f ? t.x : 1.0f - t.x, u[j].y < 0.f ? t.y : 1.0f - t.y);
  }
}

// Fast path: all pointers / leading dimensions vector-aligned, the chunk's 32 columns inside N and all 32 rows of the
// warp's slab inside M. Straight-line code in phases (loads, bias, activation over all 16 pairs, stores).
template <int EPI, int OUT>
__device__ __forceinline__ void epi_chunk_fast(const EpiLane& L, int col, const float* stage, int lane, const EpiPre<EPI>& pre,
                                               const float4 bias4) {
  const int c4 = lane & 7, rsub = lane >> 3;
  // v[2 * it], v[2 * it + 1]: columns (col, col+1) and (col+2, col+3) of row 4 * it + rsub
  float2 v[16];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    const float4 t = reinterpret_cast<const float4*>(stage + r * 32)[c4 ^ (r & 7)];
    v[2 * it] = make_float2(t.x, t.y);
    v[2 * it + 1] = make_float2(t.z, t.w);
  }
  constexpr bool kBias = (EPI == CAVIT_EPI_BIAS || EPI == CAVIT_EPI_BIAS_GELU || EPI == CAVIT_EPI_BIAS_RESID ||
                          EPI == CAVIT_EPI_EMBED || EPI == CAVIT_EPI_BIAS_RELU);
  if (kBias) {   // loaded by the caller before the accumulator wait (an L2 round trip per chunk otherwise)
    const float2 b01 = make_float2(bias4.x, bias4.y), b23 = make_float2(bias4.z, bias4.w);
#pragma unroll
    for (int it = 0; it < 8; ++it) { v[2 * it] = add2(v[2 * it], b01); v[2 * it + 1] = add2(v[2 * it + 1], b23); }
  }
  char* o = L.out + col * ((OUT == OUT_BF16) ? 2 : 4);
  if (EPI == CAVIT_EPI_BIAS_GELU) {
    // pre-activation u (bf16) out for the backward pass; GELU of the fp32 u, as the reference evaluates it (the backward's
    // GELU'(bf16 u) differs from GELU'(u) by less than the bf16 rounding of the gradient it multiplies)
    char* ax = L.aux + col * 2;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      stg_v2(ax + it * L.aux_step, pack_bf16(v[2 * it].x, v[2 * it].y), pack_bf16(v[2 * it + 1].x, v[2 * it + 1].y));
#if CAVIT_GELU_EW == 4
    {   // 16 epilogue warps: 113 registers per thread, two half batches
      float2 (&va)[8] = *reinterpret_cast<float2(*)[8]>(&v[0]);
      float2 (&vb)[8] = *reinterpret_cast<float2(*)[8]>(&v[8]);
      gelu_batch<8>(va);
      gelu_batch<8>(vb);
    }
#else
    gelu_batch<16>(v);
#endif
  } else if (EPI == CAVIT_EPI_GELU_BWD) {
    float2 g[16];
#pragma unroll
    for (int it = 0; it < 8; ++it) { g[2 * it] = unpack_bf16_fast(pre.a[it].x); g[2 * it + 1] = unpack_bf16_fast(pre.a[it].y); }
#if CAVIT_GELU_EW == 4
    {
      float2 (&ga)[8] = *reinterpret_cast<float2(*)[8]>(&g[0]);
      float2 (&gb)[8] = *reinterpret_cast<float2(*)[8]>(&g[8]);
      gelu_grad_batch<8>(ga);
      gelu_grad_batch<8>(gb);
    }
#else
    gelu_grad_batch<16>(g);
#endif
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = mul2(v[j], g[j]);
  } else if (EPI == CAVIT_EPI_BIAS_RELU) {   // nn.TransformerEncoderLayer's default activation (modelv2.py:72-78)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = make_float2(fmaxf(v[j].x, 0.f), fmaxf(v[j].y, 0.f));
  } else if (EPI == CAVIT_EPI_RELU_BWD) {    // dY * [h > 0] with h = relu(u) as stored by the forward
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float2 h01 = unpack_bf16_fast(pre.a[it].x), h23 = unpack_bf16_fast(pre.a[it].y);
      v[2 * it] = make_float2(h01.x > 0.f ? v[2 * it].x : 0.f, h01.y > 0.f ? v[2 * it].y : 0.f);
      v[2 * it + 1] = make_float2(h23.x > 0.f ? v[2 * it + 1].x : 0.f, h23.y > 0.f ? v[2 * it + 1].y : 0.f);
    }
  } else if (EPI == CAVIT_EPI_BIAS_RESID) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      v[2 * it] = add2(v[2 * it], make_float2(pre.r[it].x, pre.r[it].y));
      v[2 * it + 1] = add2(v[2 * it + 1], make_float2(pre.r[it].z, pre.r[it].w));
    }
  } else if (EPI == CAVIT_EPI_EMBED) {
    float4 pe[8];
    long long orow[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const long long r = L.embed_row + 4 * it;
      const long long sq = r / L.embed_np;
      const int t = (int)(r - sq * L.embed_np);
      pe[it] = ldg_v4f(L.pos_base + (long long)(1 + t) * L.ldr_bytes + col * 4);
      orow[it] = sq * (L.embed_np + 1) + 1 + t;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float2 a01 = add2(v[2 * it], make_float2(pe[it].x, pe[it].y)), a23 = add2(v[2 * it + 1], make_float2(pe[it].z, pe[it].w));
      stg_v4f(L.out_base + orow[it] * L.ldo_bytes + col * 4, a01.x, a01.y, a23.x, a23.y);
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    char* oi = o + it * L.out_step;
    if (OUT == OUT_RED) {
      float* of = reinterpret_cast<float*>(oi);
      red_add_f32(of, v[2 * it].x); red_add_f32(of + 1, v[2 * it].y); red_add_f32(of + 2, v[2 * it + 1].x); red_add_f32(of + 3, v[2 * it + 1].y);
    } else if (OUT == OUT_F32) {
      stg_v4f(oi, v[2 * it].x, v[2 * it].y, v[2 * it + 1].x, v[2 * it + 1].y);
    } else {
      stg_v2(oi, pack_bf16(v[2 * it].x, v[2 * it].y), pack_bf16(v[2 * it + 1].x, v[2 * it + 1].y));
    }
  }
}

// Generic path (ragged N, unaligned leading dimensions, accumulate, the CLS-skipping row map of the embedding).
template <int EPI>
__device__ __noinline__ void epi_chunk_slow(const GemmDev* pp, int g, long long row0, int col0, const float* stage, int lane) {
  const GemmDev& p = *pp;
  const int c4 = lane & 7, rsub = lane >> 3;
  const int col = col0 + c4 * 4;
  const int ncols = min(4, p.N - col);
  if (ncols <= 0) return;
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias)
    for (int i = 0; i < ncols; ++i) b[i] = __ldg(p.bias + (long long)g * p.bias_gs + col + i);
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    const long long row = row0 + r;
    if (row >= p.M) break;
    const float4 t = reinterpret_cast<const float4*>(stage + r * 32)[c4 ^ (r & 7)];
    float v[4] = {t.x + b[0], t.y + b[1], t.z + b[2], t.w + b[3]};
    long long out_row = row;
    if (EPI == CAVIT_EPI_EMBED) out_row = (row / p.embed_np) * (p.embed_np + 1) + 1 + row % p.embed_np;
    if (EPI == CAVIT_EPI_BIAS_GELU) {
      bf16* aux = reinterpret_cast<bf16*>(p.aux) + (long long)g * p.aux_gs + row * p.ldaux + col;
      for (int i = 0; i < ncols; ++i) {
        aux[i] = __float2bfloat16(v[i]);
        v[i] = gelu_poly(v[i]);   // bit-identical to the fast path (common.cuh)
      }
    } else if (EPI == CAVIT_EPI_GELU_BWD) {
      const bf16* aux = reinterpret_cast<const bf16*>(p.aux) + (long long)g * p.aux_gs + row * p.ldaux + col;
      for (int i = 0; i < ncols; ++i) v[i] *= gelu_poly_grad(__bfloat162float(aux[i]));
    } else if (EPI == CAVIT_EPI_BIAS_RELU) {
      for (int i = 0; i < ncols; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (EPI == CAVIT_EPI_RELU_BWD) {
      const bf16* aux = reinterpret_cast<const bf16*>(p.aux) + (long long)g * p.aux_gs + row * p.ldaux + col;
      for (int i = 0; i < ncols; ++i) v[i] = __bfloat162float(aux[i]) > 0.f ? v[i] : 0.f;
    } else if (EPI == CAVIT_EPI_BIAS_RESID || EPI == CAVIT_EPI_EMBED) {
      const float* rp = (EPI == CAVIT_EPI_BIAS_RESID) ? p.resid + (long long)g * p.resid_gs + row * p.ldr + col
                                                      : p.resid + (1 + row % p.embed_np) * p.ldr + col;
      for (int i = 0; i < ncols; ++i) v[i] += rp[i];
    }
    if (p.out_fp32) {
      float* o = reinterpret_cast<float*>(p.out) + (long long)g * p.out_gs + out_row * p.ldo + col;
      for (int i = 0; i < ncols; ++i) {
        if (p.split_k > 1) red_add_f32(o + i, v[i]);
        else o[i] = p.accumulate ? o[i] + v[i] : v[i];
      }
    } else {
      bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)g * p.out_gs + out_row * p.ldo + col;
      for (int i = 0; i < ncols; ++i) o[i] = __float2bfloat16(v[i]);
    }
  }
}

// Body of one epilogue warp: q = TMEM lane quadrant (rows q*32..), half = which half of the tile's columns.
template <int BN, int EPI, int OUT, int CTAS, int EW>
__device__ __forceinline__ void epilogue_role(const GemmDev& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              float* stage, volatile int* abort_flag, int q, int half, int lane,
                                              int total_tiles, int splits, int tiles_per_group, int unit, int num_units,
                                              int cta_rank) {
  constexpr int CH = BN / EW / 32;  // 32-column chunks per warp (`half` = which 1/EW of the tile's columns)
  constexpr bool kPrefetch = (EW == 2);
  const int c4 = lane & 7, rsub = lane >> 3;
  const bool fast_kind = p.vec_ok;
  constexpr int osz = (OUT == OUT_BF16) ? 2 : 4;
  int as = 0;
  uint32_t aphase = 0;
  for (int work = unit; work < total_tiles; work += num_units) {
    const int tile = work / splits;
    const int g = tile / tiles_per_group;
    const int rem = tile - g * tiles_per_group;
    const int m0 = (rem / p.tiles_n) * (GEMM_BM * CTAS) + cta_rank * GEMM_BM;
    const int n0 = (rem % p.tiles_n) * BN + half * (BN / EW);
    const long long row0 = (long long)m0 + q * 32;
    EpiLane L;
    {
      const long long r = row0 + rsub;
      L.out = reinterpret_cast<char*>(p.out) + ((long long)g * p.out_gs + r * p.ldo) * osz;
      L.out_step = 4 * p.ldo * osz;
      L.resid = reinterpret_cast<const char*>(p.resid) + ((long long)g * p.resid_gs + r * p.ldr) * 4;
      L.resid_step = 16 * p.ldr;
      L.aux = reinterpret_cast<char*>(p.aux) + ((long long)g * p.aux_gs + r * p.ldaux) * 2;
      L.aux_step = 8 * p.ldaux;
      L.bias = p.bias + (long long)g * p.bias_gs;
      const long long left = (long long)p.M - r;
      L.rows_left = left > 32 ? 32 : (left < 0 ? 0 : (int)left);
      L.embed_row = r;
      L.embed_np = p.embed_np > 0 ? p.embed_np : 1;
      L.out_base = reinterpret_cast<char*>(p.out) + (long long)g * p.out_gs * osz;
      L.pos_base = reinterpret_cast<const char*>(p.resid);
      L.ldo_bytes = p.ldo * osz;
      L.ldr_bytes = p.ldr * 4;
    }
    // L2 prefetch of what this warp's slab of the NEXT tile reads from HBM in its epilogue (the fp32 residual rows / the bf16
    // pre-activations of GELU'): those loads sit between the accumulator and the stores, and the bias + residual GEMMs are
    // HBM-bound. One 128-byte line per lane and instruction.
    // (measured: GELU' 0.385 -> 0.364 ms, out-proj 0.160 -> 0.156 ms; with a long mainloop between the prefetch and its use —
    // fc2 forward, K = 1536 — the kernel got 6 % SLOWER, also when only the current tile's slab was prefetched: hence the K limit)
#ifndef CAVIT_NO_EPI_PREFETCH   // A/B switch (NVCC_EXTRA=-DCAVIT_NO_EPI_PREFETCH with CAVIT_BUILD_TAG)
    if ((EPI == CAVIT_EPI_GELU_BWD || (EPI == CAVIT_EPI_BIAS_RESID && p.K <= 512)) && work + num_units < total_tiles) {
      const int ntile = (work + num_units) / splits;
      const int ng = ntile / tiles_per_group;
      const int nrem = ntile - ng * tiles_per_group;
      const long long nrow0 = (long long)(nrem / p.tiles_n) * (GEMM_BM * CTAS) + cta_rank * GEMM_BM + q * 32;
      const int nn0 = (nrem % p.tiles_n) * BN + half * (BN / EW);
      constexpr int esz = (EPI == CAVIT_EPI_BIAS_RESID) ? 4 : 2;
      constexpr int LPR = ((BN / EW) * esz + 127) / 128;       // lines per row of the slab
      const char* src = (EPI == CAVIT_EPI_BIAS_RESID)
                            ? reinterpret_cast<const char*>(p.resid) + ((long long)ng * p.resid_gs + nrow0 * p.ldr + nn0) * 4
                            : reinterpret_cast<const char*>(p.aux) + ((long long)ng * p.aux_gs + nrow0 * p.ldaux + nn0) * 2;
      const long long pitch = (EPI == CAVIT_EPI_BIAS_RESID) ? (long long)p.ldr * 4 : (long long)p.ldaux * 2;
#pragma unroll
      for (int i = lane; i < 32 * LPR; i += 32) {
        const int rr = i / LPR, ll = i - rr * LPR;
        if (nrow0 + rr < p.M && nn0 + ll * (128 / esz) < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + rr * pitch + ll * 128));
      }
    }
#endif
    // the fast path needs the warp's whole 32-row slab inside M; the (last) partial slab of a group takes the generic path
    const bool fast_tile = fast_kind && ((long long)p.M - row0 >= 32);   // warp-uniform
    EpiPre<EPI> pre;
    if (kPrefetch && fast_tile && n0 + 32 <= p.N) epi_prefetch<EPI>(L, n0 + c4 * 4, pre);  // overlaps the tile's mainloop
    constexpr bool kHasBias = (EPI == CAVIT_EPI_BIAS || EPI == CAVIT_EPI_BIAS_GELU || EPI == CAVIT_EPI_BIAS_RESID ||
                               EPI == CAVIT_EPI_EMBED || EPI == CAVIT_EPI_BIAS_RELU);
    float4 bias4[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      bias4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kHasBias && fast_tile && n0 + c * 32 + 32 <= p.N) bias4[c] = __ldg(reinterpret_cast<const float4*>(L.bias + n0 + c * 32 + c4 * 4));
    }
    mbar_wait(tfull0 + 8u * as, aphase, abort_flag, p.status, ERR_TIMEOUT_TMEM_FULL);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * (BN / EW);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;  // warp-uniform
      uint32_t acc[32];
      tmem_ld32(t_row + c * 32, acc);
      tmem_ld_wait();
      float4* srow = reinterpret_cast<float4*>(stage + lane * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        srow[j ^ (lane & 7)] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                           __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
      __syncwarp();
      if (fast_tile && (col0 + 32 <= p.N)) {  // warp-uniform
        if (!kPrefetch) epi_prefetch<EPI>(L, col0 + c4 * 4, pre);  // many warps: latency is hidden by the other warps
        const EpiPre<EPI> cur = pre;
        if (kPrefetch && c + 1 < CH && col0 + 64 <= p.N) epi_prefetch<EPI>(L, col0 + 32 + c4 * 4, pre);
        epi_chunk_fast<EPI, OUT>(L, col0 + c4 * 4, stage, lane, cur, bias4[c]);
      } else {
        epi_chunk_slow<EPI>(&p, g, row0, col0, stage, lane);
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {  // the MMA issuer (leader CTA) waits for the epilogue warps of BOTH CTAs of a pair
      if (CTAS == 2 && cta_rank != 0) mbar_arrive_remote(tempty0 + 8u * as, 0);
      else mbar_arrive(tempty0 + 8u * as);
    }
    as ^= 1;
    if (as == 0) aphase ^= 1u;
  }
}

template <int BN, int EPI, int OUT, int CTAS>
__global__ void __launch_bounds__((GemmCfg<BN, CTAS, EPI>::THREADS), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                         const GemmDev p) {
  using Cfg = GemmCfg<BN, CTAS, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // barrier block after the operand stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES);
  const uint32_t bar0 = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* epi_stage = reinterpret_cast<float*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (CTAS == 2) ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CTAS, num_units = gridDim.x / CTAS;  // a "unit" owns a tile: a CTA or a CTA pair

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), Cfg::EPI_WARPS * CTAS);
    }
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.passes > 1) {
      prefetch_tmap(&tmAlo);
      prefetch_tmap(&tmBlo);
    }
  }
  if (warp == 1) {
    if (CTAS == 2) {
      tmem_alloc_pair(smem_u32(tmem_slot), Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int splits = p.split_k > 1 ? p.split_k : 1;
  const int tiles_per_group = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_per_group * p.groups * splits;  // work item = (group, m tile, n tile, k split)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (each CTA stages its own rows)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = unit; work < total_tiles; work += num_units) {
        const int split = work % splits;
        const int tile = work / splits;
        const int g = tile / tiles_per_group;
        const int rem = tile - g * tiles_per_group;
        const int m0 = (rem / p.tiles_n) * (GEMM_BM * CTAS) + cta_rank * GEMM_BM;
        const int n0 = (rem % p.tiles_n) * BN + cta_rank * (BN / CTAS);  // this CTA's share of the B tile
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        // split-operand mode: the k range is walked three times with (A, B) = (hi, hi), (lo, hi), (hi, lo) — to the MMA
        // issuer it is one three-times-longer accumulation, the missing lo x lo term is below 2^-16 relative
        for (int ps = 0; ps < p.passes; ++ps) {
        const CUtensorMap* mapA = (ps == 1) ? &tmAlo : &tmA;
        const CUtensorMap* mapB = (ps == 2) ? &tmBlo : &tmB;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_EMPTY);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          // the pair's transactions are all credited to the leader's barrier (see tma_load_3d_pair)
          if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CTAS);
          const int k0 = kb * GEMM_BK;
          auto load = [&](const CUtensorMap* m, uint32_t dst, int c0, int c1) {
            if (CTAS == 2) tma_load_3d_pair(m, full_bar(stage), dst, c0, c1, g);
            else tma_load_3d(m, full_bar(stage), dst, c0, c1, g);
          };
          if (!p.a_mn) {
            load(mapA, sA, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c) load(mapA, sA + c * (GEMM_BK * 128), m0 + c * 64, k0);
          }
          if (!p.b_mn) {
            load(mapB, sB, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / CTAS / 64; ++c) load(mapB, sB + c * (GEMM_BK * 128), n0 + c * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of a pair only)
    // The whole warp walks the schedule (converged waits); one elected lane issues, so that the compiler emits
    // ELECT + predicated UTCHMMA instead of a per-instruction election loop.
    if (cta_rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN, p.a_mn, p.b_mn, GEMM_BM * CTAS);
      // descriptor strides: K-major: SBO = 1024 (8 rows x 128 B); MN-major: LBO = 64-wide chunk
      // pitch (BK rows x 128 B), SBO = 1024 (8 k-rows x 128 B).
      const uint32_t a_lbo = p.a_mn ? GEMM_BK * 128 : 16, b_lbo = p.b_mn ? GEMM_BK * 128 : 16;
      const uint32_t a_kstep = (p.a_mn ? 16 * 128 : 32) >> 4, b_kstep = (p.b_mn ? 16 * 128 : 32) >> 4;  // in 16-byte units
      const uint64_t adesc0 = umma_desc_sw128(smem_base, a_lbo, 1024);
      const uint64_t bdesc0 = umma_desc_sw128(smem_base + Cfg::A_BYTES, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int work = unit; work < total_tiles; work += num_units) {
        const int split = work % splits;
        const int kb0 = split * p.kb_per_split;
        const int num_kb = (min(num_kb_total, kb0 + p.kb_per_split) - kb0) * p.passes;
        mbar_wait(tempty_bar(as), aphase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_TMEM_EMPTY);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, abort_flag, p.status, ERR_TIMEOUT_FULL);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint32_t>(stage * (Cfg::STAGE_BYTES >> 4));
            const uint64_t bd = bdesc0 + static_cast<uint32_t>(stage * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              if (CTAS == 2) umma_bf16_ss_pair(d_tmem, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16_ss(d_tmem, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (CTAS == 2) {
              umma_commit_pair(empty_bar(stage));  // frees the stage in BOTH CTAs once these MMAs retire
              if (kb == num_kb - 1) umma_commit_pair(tfull_bar(as));
            } else {
              umma_commit(empty_bar(stage));
              if (kb == num_kb - 1) umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2..9)
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which 1/EW of the tile's columns
    float* stage = epi_stage + (warp - 2) * 1024;
    epilogue_role<BN, EPI, OUT, CTAS, Cfg::EW>(p, tmem_base, tfull_bar(0), tempty_bar(0), stage, abort_flag, q, half, lane, total_tiles,
                                      splits, tiles_per_group, unit, num_units, cta_rank);
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();  // the peer may still read this CTA's smem / barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, int OUT, int CTAS>
static int launch_gemm_epi(const CUtensorMap* const* tm, GemmDev& d, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTAS, EPI>;
  static PerDeviceFlag attr_set;
  if (attr_set.unset()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, OUT, CTAS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "gemm smem attribute: %s", cudaGetErrorString(e));
    attr_set.set();
  }
  const long long total = (long long)d.tiles_m * d.tiles_n * d.groups * d.split_k;
  const int max_units = sm_count() / CTAS;
  const int units = (int)(total < max_units ? total : max_units);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CTAS, 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, EPI, OUT, CTAS>, *tm[0], *tm[1], *tm[2], *tm[3], d);
  if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "cavit_gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("cavit_gemm");
}

template <int BN, int CTAS>
static int launch_gemm(const CUtensorMap* const* tm, GemmDev& d, cudaStream_t stream) {
  d.tiles_m = (d.M + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS);
  d.tiles_n = (d.N + BN - 1) / BN;
  const int num_kb = (d.K + GEMM_BK - 1) / GEMM_BK;
  if (d.split_k > num_kb) d.split_k = num_kb;
  if (d.split_k < 1) d.split_k = 1;
  d.kb_per_split = (num_kb + d.split_k - 1) / d.split_k;
  d.split_k = (num_kb + d.kb_per_split - 1) / d.kb_per_split;  // no empty splits
  if (d.accumulate) d.vec_ok = 0;  // read-modify-write outputs take the generic path
  const int out = d.split_k > 1 ? OUT_RED : (d.out_fp32 ? OUT_F32 : OUT_BF16);
#define CAVIT_GEMM_CASE(E, O) \
  if (d.epi == E && out == O) return launch_gemm_epi<BN, E, O, CTAS>(tm, d, stream);
  CAVIT_GEMM_CASE(CAVIT_EPI_NONE, OUT_BF16)
  CAVIT_GEMM_CASE(CAVIT_EPI_NONE, OUT_F32)
  CAVIT_GEMM_CASE(CAVIT_EPI_NONE, OUT_RED)
  CAVIT_GEMM_CASE(CAVIT_EPI_BIAS, OUT_BF16)
  CAVIT_GEMM_CASE(CAVIT_EPI_BIAS, OUT_F32)
  CAVIT_GEMM_CASE(CAVIT_EPI_BIAS_GELU, OUT_BF16)
  CAVIT_GEMM_CASE(CAVIT_EPI_BIAS_RESID, OUT_F32)
  CAVIT_GEMM_CASE(CAVIT_EPI_GELU_BWD, OUT_BF16)
  CAVIT_GEMM_CASE(CAVIT_EPI_EMBED, OUT_F32)
  CAVIT_GEMM_CASE(CAVIT_EPI_BIAS_RELU, OUT_BF16)
  CAVIT_GEMM_CASE(CAVIT_EPI_RELU_BWD, OUT_BF16)
#undef CAVIT_GEMM_CASE
  return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: epilogue %d with output mode %d is not instantiated", d.epi, out);
}

}  // namespace cavit

using namespace cavit;

extern "C" int cavit_gemm(const cavit_gemm_args* a, void* stream) {
  if (!a) return fail(CAVIT_E_BADARG, "cavit_gemm: null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->groups <= 0)
    return fail(CAVIT_E_BADARG, "cavit_gemm: non-positive extent M=%d N=%d K=%d groups=%d", a->M, a->N, a->K, a->groups);
  if (!a->A || !a->B || !a->out) return fail(CAVIT_E_BADARG, "cavit_gemm: null operand");
  if ((a->A_lo == nullptr) != (a->B_lo == nullptr))
    return fail(CAVIT_E_BADARG, "cavit_gemm: split operands need both A_lo and B_lo");
  if (a->A_lo && ((reinterpret_cast<uintptr_t>(a->A_lo) & 15) || (reinterpret_cast<uintptr_t>(a->B_lo) & 15)))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: lo planes need 16-byte aligned bases");
  if ((a->lda % 8) || (a->ldb % 8) || (a->a_gs % 8) || (a->b_gs % 8) ||
      (reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: operands need 16-byte aligned rows (ld %% 8 == 0)");
  // inner (contiguous) extents must also keep TMA's 16-byte rule
  const int a_inner = a->a_mn ? a->M : a->K, b_inner = a->b_mn ? a->N : a->K;
  if ((a_inner % 8) || (b_inner % 8))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: contiguous extent must be a multiple of 8 (A %d, B %d)", a_inner, b_inner);
  if (a->epi < CAVIT_EPI_NONE || a->epi > CAVIT_EPI_RELU_BWD) return fail(CAVIT_E_BADARG, "cavit_gemm: bad epilogue %d", a->epi);
  if ((a->epi == CAVIT_EPI_BIAS_GELU || a->epi == CAVIT_EPI_GELU_BWD || a->epi == CAVIT_EPI_RELU_BWD) && !a->aux)
    return fail(CAVIT_E_BADARG, "cavit_gemm: epilogue %d needs aux", a->epi);
  if ((a->epi == CAVIT_EPI_BIAS_RESID || a->epi == CAVIT_EPI_EMBED) && (!a->resid || !a->out_fp32))
    return fail(CAVIT_E_BADARG, "cavit_gemm: residual epilogues need resid and fp32 out");
  if (a->epi == CAVIT_EPI_EMBED && a->embed_np <= 0) return fail(CAVIT_E_BADARG, "cavit_gemm: embed_np");
  if ((a->epi == CAVIT_EPI_BIAS || a->epi == CAVIT_EPI_BIAS_GELU || a->epi == CAVIT_EPI_BIAS_RESID || a->epi == CAVIT_EPI_EMBED ||
       a->epi == CAVIT_EPI_BIAS_RELU) && !a->bias)
    return fail(CAVIT_E_BADARG, "cavit_gemm: epilogue %d needs bias", a->epi);
  if (a->accumulate && !a->out_fp32) return fail(CAVIT_E_BADARG, "cavit_gemm: accumulate needs fp32 out");
  if (a->split_k > 1 && (!a->out_fp32 || a->epi != CAVIT_EPI_NONE))
    return fail(CAVIT_E_BADARG, "cavit_gemm: split_k needs fp32 out and EPI_NONE");
  int* status = status_word();
  if (!status) return fail(CAVIT_E_DEVICE, "cavit_gemm: no device status word");

  // N tile: 256 when it divides the work well, else 128 (short N). CTA pairs (256-row tiles, each CTA staging half
  // of the B tile) whenever there is more than one 128-row tile; CAVIT_GEMM_1CTA=1 forces single-CTA tiles.
  // wgrad (MN-major A: few, short tiles with a long split-K reduction) keeps single-CTA tiles. With pairs and a K-major
  // B operand, N tiles of 192 remove the 10-25 % padding of N = 1152 (QKV) and N = 384 (out-proj, fc2).
  static const bool force_1cta = [] { const char* e = getenv("CAVIT_GEMM_1CTA"); return e && e[0] == '1'; }();
  static const bool wgrad_pair = [] { const char* e = getenv("CAVIT_WGRAD_PAIR"); return e && e[0] == '1'; }();
  // dgrad (K-major A, MN-major B) with N = 384 runs as single-CTA 128 x 192 tiles instead of pair 256 x 128 tiles: a pair
  // cannot split a 192-wide MN-major B tile on a 64-column chunk boundary, and its 128-wide MMAs (32 tensor cycles per SM)
  // are shorter than the issue time of one MMA; measured 5 % faster (fc1 dgrad 0.251 -> 0.239 ms). CAVIT_DGRAD_BN192=0 disables.
  static const bool dgrad192 = [] { const char* e = getenv("CAVIT_DGRAD_BN192"); return !(e && e[0] == '0'); }();
  const bool dgrad_single = dgrad192 && !a->a_mn && a->b_mn && a->N % 192 == 0 && a->N % 256 != 0;
  const int CTAS = (!force_1cta && !dgrad_single && a->M > GEMM_BM && (!a->a_mn || wgrad_pair)) ? 2 : 1;
  int BN = (a->N >= 256 && (a->N % 256 == 0 || a->N > 1024)) ? 256 : 128;
  if (CTAS == 2 && !a->b_mn && a->N % 192 == 0 && a->N % 256 != 0) BN = 192;
  // wgrad with N = 384 (dW[., C = 384]): two 192-wide tiles instead of three 128-wide ones. A single issuing thread needs
  // ~40-75 cycles per tcgen05.mma while a 128x128x16 MMA occupies the tensor pipe for only 32: wider MMAs (48 cycles) are
  // what keeps the split-K wgrad loop from being issue-bound.
  static const bool wgrad192 = [] { const char* e = getenv("CAVIT_WGRAD_BN192"); return !(e && e[0] == '0'); }();
  if (wgrad192 && CTAS == 1 && a->a_mn && a->b_mn && a->N % 192 == 0 && a->N % 256 != 0) BN = 192;
  if (dgrad_single) BN = 192;
  const bool split_ops = a->A_lo != nullptr;
  auto map_a = [&](const void* base) {
    return !a->a_mn ? tensor_map_bf16_3d(base, a->K, a->M, a->groups, a->lda, a->a_gs, 64, GEMM_BM)
                    : tensor_map_bf16_3d(base, a->M, a->K, a->groups, a->lda, a->a_gs, 64, GEMM_BK);
  };
  auto map_b = [&](const void* base) {
    return !a->b_mn ? tensor_map_bf16_3d(base, a->K, a->N, a->groups, a->ldb, a->b_gs, 64, BN / CTAS)
                    : tensor_map_bf16_3d(base, a->N, a->K, a->groups, a->ldb, a->b_gs, 64, GEMM_BK);
  };
  const CUtensorMap* tm[4];
  tm[0] = map_a(a->A);
  tm[1] = map_b(a->B);
  tm[2] = split_ops ? map_a(a->A_lo) : tm[0];
  tm[3] = split_ops ? map_b(a->B_lo) : tm[1];
  if (!tm[0] || !tm[1] || !tm[2] || !tm[3]) return CAVIT_E_BADARG;

  GemmDev d;
  d.M = a->M; d.N = a->N; d.K = a->K; d.groups = a->groups;
  d.a_mn = a->a_mn ? 1 : 0; d.b_mn = a->b_mn ? 1 : 0;
  d.epi = a->epi; d.out_fp32 = a->out_fp32; d.accumulate = a->accumulate; d.embed_np = a->embed_np;
  d.out = a->out; d.ldo = a->ldo; d.out_gs = a->out_gs;
  d.bias = a->bias; d.bias_gs = a->bias_gs;
  d.resid = a->resid; d.ldr = a->ldr; d.resid_gs = a->resid_gs;
  d.aux = a->aux; d.ldaux = a->ldaux; d.aux_gs = a->aux_gs;
  d.tiles_m = d.tiles_n = 0;
  {
    auto al = [](const void* q, uintptr_t m) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & (m - 1)) == 0; };
    const bool out_ok = a->out_fp32 ? (al(a->out, 16) && a->ldo % 4 == 0 && a->out_gs % 4 == 0)
                                    : (al(a->out, 8) && a->ldo % 4 == 0 && a->out_gs % 4 == 0);
    d.vec_ok = out_ok && al(a->bias, 16) && a->bias_gs % 4 == 0 && al(a->resid, 16) && a->ldr % 4 == 0 &&
               a->resid_gs % 4 == 0 && al(a->aux, 8) && a->ldaux % 4 == 0 && a->aux_gs % 4 == 0;
  }
  d.split_k = a->split_k > 1 ? a->split_k : 1;
  d.passes = split_ops ? 3 : 1;
  d.kb_per_split = 0;
  d.status = status;
  if (d.split_k > 1 && !a->accumulate) {  // atomics combine into a zeroed output
    cudaStream_t st = as_stream(stream);
    if (a->ldo == a->N && (a->groups == 1 || a->out_gs == (int64_t)a->M * a->N)) {
      cudaMemsetAsync(a->out, 0, sizeof(float) * (size_t)a->groups * a->M * a->N, st);
    } else {
      for (int g = 0; g < a->groups; ++g)
        cudaMemset2DAsync(reinterpret_cast<float*>(a->out) + (size_t)g * a->out_gs, sizeof(float) * a->ldo, 0,
                          sizeof(float) * a->N, a->M, st);
    }
  }
  if (CTAS == 2) {
    if (BN == 256) return launch_gemm<256, 2>(tm, d, as_stream(stream));
    if (BN == 192) return launch_gemm<192, 2>(tm, d, as_stream(stream));
    return launch_gemm<128, 2>(tm, d, as_stream(stream));
  }
  if (BN == 256) return launch_gemm<256, 1>(tm, d, as_stream(stream));
  if (BN == 192) return launch_gemm<192, 1>(tm, d, as_stream(stream));
  return launch_gemm<128, 1>(tm, d, as_stream(stream));
}
