// cavit-sm100 — K-LN: LayerNorm forward / backward, warp-per-row, HBM-bound.
//
// Each warp owns one row: every lane issues up to NV independent 16-byte loads (coalesced 512 B
// per warp instruction), statistics are reduced with warp shuffles in fp32 (two-pass: mean, then
// centred variance, like ATen), and the normalised row is written as bf16 (the GEMM operand
// format) with 8-byte stores. Backward fuses the residual-gradient add and the bf16 copy of dx
// that the next dgrad/wgrad GEMM consumes; d(gamma)/d(beta) are reduced deterministically in two
// stages (per-warp partial rows, then a column reduction).
//
// The "fusion" row map implements `cat(cls_i, patches_j)` (model_cross.py:140) without the copy:
// row 0 of every sample is read from the CLS-donor stream, rows 1.. from the patch-donor stream.
//
// Algorithmic bytes per row (C columns): fwd 4C (x) + 2C (y) + 8; bwd 2C (dy) + 4C (x) + 4C (dresid)
// + 4C (dx) + 2C (dx bf16).
#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;
constexpr int LN_MAX_BLOCKS_PER_GROUP = 256;
constexpr int LN_MAX_FUSIONS = 16;

struct RowMap {
  int fusion;  // 0: plain rows; 1: fusion gather
  int N;       // tokens per sample (fusion mode)
  int B;       // samples (fusion mode)
  const float* x_cls;   // fusion: CLS inputs [K][B][C] (row 0 of every sample is read from here)
  const float* dy_cls;  // fusion bwd: optional extra fp32 gradient [K][B][C] added to dy on row 0
  int cls_src[LN_MAX_FUSIONS];
  int tok_src[LN_MAX_FUSIONS];
};

// Pointer to the source row (x) of logical row r of group g.
__device__ __forceinline__ const float* src_row(const RowMap& rm, const float* x, int g, long long r, long long row_stride,
                                                long long gs, int C) {
  if (!rm.fusion) return x + (long long)g * gs + r * row_stride;
  const int n = (int)(r % rm.N);
  if (n == 0) return rm.x_cls + ((long long)g * rm.B + r / rm.N) * C;
  return x + (long long)rm.tok_src[g] * gs + r * row_stride;
}
// Offset of the destination row (dx) in fusion mode: scatter back into the donor streams.
__device__ __forceinline__ long long fusion_dst_offset(const RowMap& rm, int g, long long r, long long row_stride,
                                                       long long gs) {
  const int n = (int)(r % rm.N);
  const int s = (n == 0) ? rm.cls_src[g] : rm.tok_src[g];
  return (long long)s * gs + r * row_stride;
}

template <int NV>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, long long row_stride, long long gs, int rows_per_group, int groups, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, bf16* __restrict__ y,
              bf16* __restrict__ y_lo, float* __restrict__ y32, float* __restrict__ mean, float* __restrict__ rstd,
              const RowMap rm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)rows_per_group * groups;
  const int C4 = C >> 2;
  const float inv_c = 1.0f / (float)C;
  for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < total; row += (long long)gridDim.x * LN_WARPS) {
    const int g = (int)(row / rows_per_group);
    const long long r = row - (long long)g * rows_per_group;
    const float4* xr = reinterpret_cast<const float4*>(src_row(rm, x, g, r, row_stride, gs, C));
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        v[i] = __ldg(xr + c4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mu = warp_sum(s) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        ss += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rs = rsqrtf(warp_sum(ss) * inv_c + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma + (long long)g * C);
    const float4* b4 = reinterpret_cast<const float4*>(beta + (long long)g * C);
    uint2* yr = reinterpret_cast<uint2*>(y + row * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        const float4 gm = __ldg(g4 + c4), bt = __ldg(b4 + c4);
        const float4 f = make_float4((v[i].x - mu) * rs * gm.x + bt.x, (v[i].y - mu) * rs * gm.y + bt.y,
                                     (v[i].z - mu) * rs * gm.z + bt.z, (v[i].w - mu) * rs * gm.w + bt.w);
        if (y) {
          uint2 o;
          o.x = pack_bf16(f.x, f.y);
          o.y = pack_bf16(f.z, f.w);
          yr[c4] = o;
          if (y_lo) {   // fp32-tolerance mode: second bf16 plane with the rounding residual (split GEMM operand)
            const float2 h01 = unpack_bf16_fast(o.x), h23 = unpack_bf16_fast(o.y);
            uint2 l;
            l.x = pack_bf16(f.x - h01.x, f.y - h01.y);
            l.y = pack_bf16(f.z - h23.x, f.w - h23.y);
            reinterpret_cast<uint2*>(y_lo + row * C)[c4] = l;
          }
        }
        // post-norm encoders (modelv2.py:72-78): the normalised row IS the next residual stream, kept in fp32
        if (y32) reinterpret_cast<float4*>(y32 + row * C)[c4] = f;
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// grid = (blocks_per_group, groups). partials: [groups][blocks_per_group][3][C] followed by one ticket counter per group.
// COL: also accumulate the column sums of the OUTPUT gradient (dx incl. the residual term): that is the bias gradient
// of the Linear layer that produced this LayerNorm's input (out-proj / fc2), so no separate pass over dx is needed.
// The block that takes the last ticket of its group reduces the partial rows (no separate finalize launch).
// Two resident blocks per SM up to C = 384 (NV <= 3: fits 128 registers); wider rows keep every load of the row in registers
// (up to 5 float4 arrays of NV entries), which spills under a 128-register cap (C = 1024: 1.4 KB of spills per thread and a
// 2.8x slower kernel), so they run one block per SM with the full register file.
template <int NV, bool COL, bool DYF>
__global__ void __launch_bounds__(LN_THREADS, (NV <= 3 ? 2 : 1))
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, long long row_stride, long long gs,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              int rows_per_group, int C, const float* dresid, float* dx, long long dx_row_stride, long long dx_gs,
              bf16* __restrict__ dx_bf16, bf16* __restrict__ dx_lo, float* __restrict__ partials,
              unsigned int* __restrict__ tickets,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcol, const RowMap rm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y;
  const int C4 = C >> 2;
  const float inv_c = 1.0f / (float)C;
  float4 gm[NV], dg[NV], db[NV], dc[COL ? NV : 1];
  const float4* g4 = reinterpret_cast<const float4*>(gamma + (long long)g * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    gm[i] = (c4 < C4) ? __ldg(g4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (COL) dc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = (long long)blockIdx.x * LN_WARPS + warp; r < rows_per_group; r += (long long)gridDim.x * LN_WARPS) {
    const long long row = (long long)g * rows_per_group + r;
    const float4* xr = reinterpret_cast<const float4*>(src_row(rm, x, g, r, row_stride, gs, C));
    // DYF: the incoming gradient is the fp32 residual-stream gradient itself (post-norm encoders), else a bf16 dgrad output
    const uint2* dyr = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_) + row * C);
    const float4* dyf = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * C);
    const float4* dyc = nullptr;  // extra fp32 gradient on the CLS row of a fusion
    if (rm.fusion && rm.dy_cls && (r % rm.N) == 0)
      dyc = reinterpret_cast<const float4*>(rm.dy_cls + ((long long)g * rm.B + r / rm.N) * C);
    long long doff;
    if (rm.fusion) doff = fusion_dst_offset(rm, g, r, row_stride, gs);  // scatter back into the donor stream
    else doff = (long long)g * dx_gs + r * dx_row_stride;
    // issue every load of the row up front (x, dy, residual gradient) before the reductions
    float4 xv[NV], dr[NV], d4[DYF ? NV : 1];
    uint2 d2[DYF ? 1 : NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        xv[i] = __ldg(xr + c4);
        if (DYF) d4[DYF ? i : 0] = __ldg(dyf + c4);
        else d2[DYF ? 0 : i] = __ldg(dyr + c4);
        if (!rm.fusion && dresid) dr[i] = *(reinterpret_cast<const float4*>(dresid + doff) + c4);
        else dr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], gy[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        float2 d01, d23;
        if (DYF) { d01 = make_float2(d4[DYF ? i : 0].x, d4[DYF ? i : 0].y); d23 = make_float2(d4[DYF ? i : 0].z, d4[DYF ? i : 0].w); }
        else { d01 = unpack_bf16(d2[DYF ? 0 : i].x); d23 = unpack_bf16(d2[DYF ? 0 : i].y); }
        if (dyc) {
          const float4 e = __ldg(dyc + c4);
          d01.x += e.x; d01.y += e.y; d23.x += e.z; d23.y += e.w;
        }
        xh[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
        dg[i].x += d01.x * xh[i].x; dg[i].y += d01.y * xh[i].y; dg[i].z += d23.x * xh[i].z; dg[i].w += d23.y * xh[i].w;
        db[i].x += d01.x; db[i].y += d01.y; db[i].z += d23.x; db[i].w += d23.y;
        gy[i] = make_float4(d01.x * gm[i].x, d01.y * gm[i].y, d23.x * gm[i].z, d23.y * gm[i].w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      }
    }
    const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < C4) {
        float4 o;
        o.x = rs * (gy[i].x - m1 - xh[i].x * m2);
        o.y = rs * (gy[i].y - m1 - xh[i].y * m2);
        o.z = rs * (gy[i].z - m1 - xh[i].z * m2);
        o.w = rs * (gy[i].w - m1 - xh[i].w * m2);
        if (rm.fusion) {
          float* d = dx + doff + c4 * 4;
          atomicAdd(d + 0, o.x); atomicAdd(d + 1, o.y); atomicAdd(d + 2, o.z); atomicAdd(d + 3, o.w);
        } else {
          o.x += dr[i].x; o.y += dr[i].y; o.z += dr[i].z; o.w += dr[i].w;
          *(reinterpret_cast<float4*>(dx + doff) + c4) = o;
          if (COL) { dc[i].x += o.x; dc[i].y += o.y; dc[i].z += o.z; dc[i].w += o.w; }
          if (dx_bf16) {
            uint2 q;
            q.x = pack_bf16(o.x, o.y);
            q.y = pack_bf16(o.z, o.w);
            *(reinterpret_cast<uint2*>(dx_bf16 + row * C) + c4) = q;
            if (dx_lo) {
              const float2 h01 = unpack_bf16_fast(q.x), h23 = unpack_bf16_fast(q.y);
              uint2 l;
              l.x = pack_bf16(o.x - h01.x, o.y - h01.y);
              l.y = pack_bf16(o.z - h23.x, o.w - h23.y);
              *(reinterpret_cast<uint2*>(dx_lo + row * C) + c4) = l;
            }
          }
        }
      }
    }
  }
  // d(gamma), d(beta) (, column sums): combine the block's 8 warps in shared memory, then one partial row per block
  constexpr int NP = COL ? 3 : 2;
  __shared__ float s_part[NP * 1024];
  __shared__ unsigned int s_ticket;
  for (int i = threadIdx.x; i < NP * C; i += LN_THREADS) s_part[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    if (c4 < C4) {
      float* pg = s_part + c4 * 4;
      float* pb = s_part + C + c4 * 4;
      atomicAdd(pg + 0, dg[i].x); atomicAdd(pg + 1, dg[i].y); atomicAdd(pg + 2, dg[i].z); atomicAdd(pg + 3, dg[i].w);
      atomicAdd(pb + 0, db[i].x); atomicAdd(pb + 1, db[i].y); atomicAdd(pb + 2, db[i].z); atomicAdd(pb + 3, db[i].w);
      if (COL) {
        float* pc = s_part + 2 * C + c4 * 4;
        atomicAdd(pc + 0, dc[i].x); atomicAdd(pc + 1, dc[i].y); atomicAdd(pc + 2, dc[i].z); atomicAdd(pc + 3, dc[i].w);
      }
    }
  }
  __syncthreads();
  const long long prow = (long long)g * gridDim.x + blockIdx.x;
  float* po = partials + prow * 3 * C;
  for (int i = threadIdx.x; i < NP * C; i += LN_THREADS) po[i] = s_part[i];
  // last block of the group reduces the partial rows
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(tickets + g, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  const float* pbase = partials + (long long)g * gridDim.x * 3 * C;
  for (int i = threadIdx.x; i < NP * C; i += LN_THREADS) {
    float acc = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;   // fixed summation order: deterministic
    unsigned int r = 0;
    for (; r + 4 <= gridDim.x; r += 4) {
      acc += __ldcg(pbase + (long long)r * 3 * C + i);
      acc1 += __ldcg(pbase + (long long)(r + 1) * 3 * C + i);
      acc2 += __ldcg(pbase + (long long)(r + 2) * 3 * C + i);
      acc3 += __ldcg(pbase + (long long)(r + 3) * 3 * C + i);
    }
    for (; r < gridDim.x; ++r) acc += __ldcg(pbase + (long long)r * 3 * C + i);
    acc = (acc + acc1) + (acc2 + acc3);
    const int which = i / C, c = i - which * C;
    float* out = which == 0 ? dgamma : (which == 1 ? dbeta : dcol);
    out[(long long)g * C + c] = acc;
  }
  if (threadIdx.x == 0) tickets[g] = 0;   // ready for the next launch (stream order)
}

static int blocks_per_group(long long rows, int groups) {
  long long b = (rows + LN_WARPS - 1) / LN_WARPS;
  // two resident blocks per SM (128 registers / thread): one grid-stride wave, and few partial rows for the
  // in-kernel reduction of the parameter gradients
  long long want = ((long long)sm_count() * 2 + groups - 1) / groups;
  if (want > LN_MAX_BLOCKS_PER_GROUP) want = LN_MAX_BLOCKS_PER_GROUP;
  if (b > want) b = want;
  if (b < 1) b = 1;
  return (int)b;
}

static int ln_fwd_launch(const float* x, long long row_stride, long long gs, int rpg, int groups, int C,
                         const float* gamma, const float* beta, float eps, void* y, float* y32, float* mean, float* rstd,
                         const RowMap& rm, cudaStream_t st, void* y_lo = nullptr) {
  if (C % 4 || C <= 0 || C > 1024) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm: C=%d (need C %% 4 == 0, C <= 1024)", C);
  if ((row_stride % 4) || (gs % 4)) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm: strides must be multiples of 4");
  const long long total = (long long)rpg * groups;
  if (total <= 0) return fail(CAVIT_E_BADARG, "layernorm: no rows");
  long long blocks = (total + LN_WARPS - 1) / LN_WARPS;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const int nv = (C / 4 + 31) / 32;
#define LN_FWD_CASE(NVV)                                                                                      \
  case NVV:                                                                                                   \
    ln_fwd_kernel<NVV><<<(int)blocks, LN_THREADS, 0, st>>>(x, row_stride, gs, rpg, groups, C, gamma, beta, eps, \
                                                           reinterpret_cast<bf16*>(y), reinterpret_cast<bf16*>(y_lo), y32, mean, rstd, rm);  \
    break;
  switch (nv) {
    LN_FWD_CASE(1) LN_FWD_CASE(2) LN_FWD_CASE(3) LN_FWD_CASE(4) LN_FWD_CASE(5) LN_FWD_CASE(6) LN_FWD_CASE(7) LN_FWD_CASE(8)
    default: return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm: C=%d", C);
  }
#undef LN_FWD_CASE
  count_launch();
  return check_launch("cavit_ln_fwd");
}

static int ln_bwd_launch(const void* dy, const float* x, long long row_stride, long long gs, const float* mean,
                         const float* rstd, const float* gamma, int rpg, int groups, int C, const float* dresid,
                         float* dx, long long dx_rs, long long dx_gs, void* dx_bf16, float* dgamma, float* dbeta,
                         float* dcol, float* partials, const RowMap& rm, cudaStream_t st, bool dy_f32 = false,
                         void* dx_lo = nullptr) {
  if (C % 4 || C <= 0 || C > 1024) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm bwd: C=%d", C);
  if ((row_stride % 4) || (gs % 4) || (dx_rs % 4) || (dx_gs % 4))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm bwd: strides must be multiples of 4");
  if (!partials || !dgamma || !dbeta) return fail(CAVIT_E_BADARG, "layernorm bwd: null workspace / outputs");
  const int bpg = blocks_per_group(rpg, groups);
  dim3 grid(bpg, groups);
  const int nv = (C / 4 + 31) / 32;
  if (dcol && rm.fusion) return fail(CAVIT_E_BADARG, "layernorm bwd: dcol is not available on the fusion row map");
  unsigned int* tickets = reinterpret_cast<unsigned int*>(partials + (size_t)groups * LN_MAX_BLOCKS_PER_GROUP * 3 * C);
#define LN_BWD_LAUNCH(NVV, COLV, DYFV)                                                                              \
  ln_bwd_kernel<NVV, COLV, DYFV><<<grid, LN_THREADS, 0, st>>>(dy, x, row_stride, gs, mean, rstd, gamma, rpg, C, dresid, dx, \
                                                              dx_rs, dx_gs, reinterpret_cast<bf16*>(dx_bf16),             \
                                                              reinterpret_cast<bf16*>(dx_lo), partials,                   \
                                                              tickets, dgamma, dbeta, dcol, rm)
#define LN_BWD_CASE(NVV)                                                    \
  case NVV:                                                                 \
    if (dy_f32) { if (dcol) LN_BWD_LAUNCH(NVV, true, true); else LN_BWD_LAUNCH(NVV, false, true); }   \
    else { if (dcol) LN_BWD_LAUNCH(NVV, true, false); else LN_BWD_LAUNCH(NVV, false, false); }        \
    break;
  switch (nv) {
    LN_BWD_CASE(1) LN_BWD_CASE(2) LN_BWD_CASE(3) LN_BWD_CASE(4) LN_BWD_CASE(5) LN_BWD_CASE(6) LN_BWD_CASE(7) LN_BWD_CASE(8)
    default: return fail(CAVIT_E_UNSUPPORTED_SHAPE, "layernorm bwd: C=%d", C);
  }
#undef LN_BWD_CASE
#undef LN_BWD_LAUNCH
  count_launch();
  return check_launch("cavit_ln_bwd");
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_ln_fwd(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups,
                 int32_t C, const float* gamma, const float* beta, float eps, void* y, float* mean, float* rstd,
                 void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return fail(CAVIT_E_BADARG, "cavit_ln_fwd: null pointer");
  RowMap rm{};
  rm.fusion = 0;
  return ln_fwd_launch(x, x_row_stride, x_gs, rows_per_group, groups, C, gamma, beta, eps, y, nullptr, mean, rstd, rm,
                       as_stream(stream));
}

int cavit_ln_fwd_dual(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups,
                      int32_t C, const float* gamma, const float* beta, float eps, void* y_bf16, float* y_f32, float* mean,
                      float* rstd, void* stream) {
  if (!x || !gamma || !beta || (!y_bf16 && !y_f32) || !mean || !rstd) return fail(CAVIT_E_BADARG, "cavit_ln_fwd_dual: null pointer");
  RowMap rm{};
  rm.fusion = 0;
  return ln_fwd_launch(x, x_row_stride, x_gs, rows_per_group, groups, C, gamma, beta, eps, y_bf16, y_f32, mean, rstd, rm,
                       as_stream(stream));
}

/* fp32-tolerance mode: the normalised rows as bf16 hi + lo planes (the split operand of the next GEMM) */
int cavit_ln_fwd_split(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups, int32_t C,
                       const float* gamma, const float* beta, float eps, void* y_hi, void* y_lo, float* mean, float* rstd,
                       void* stream) {
  if (!x || !gamma || !beta || !y_hi || !y_lo || !mean || !rstd) return fail(CAVIT_E_BADARG, "cavit_ln_fwd_split: null pointer");
  RowMap rm{};
  rm.fusion = 0;
  return ln_fwd_launch(x, x_row_stride, x_gs, rows_per_group, groups, C, gamma, beta, eps, y_hi, nullptr, mean, rstd, rm,
                       as_stream(stream), y_lo);
}

/* fp32-tolerance mode: fp32 incoming gradient (an fp32 dgrad output), dx additionally as bf16 hi + lo planes (nullable pair) */
int cavit_ln_bwd_split(const float* dy_f32, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                       const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                       const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, float* dcol, float* partials, void* stream) {
  if (!dy_f32 || !x || !mean || !rstd || !gamma || !dx) return fail(CAVIT_E_BADARG, "cavit_ln_bwd_split: null pointer");
  if ((dx_hi == nullptr) != (dx_lo == nullptr)) return fail(CAVIT_E_BADARG, "cavit_ln_bwd_split: dx_hi and dx_lo go together");
  RowMap rm{};
  rm.fusion = 0;
  return ln_bwd_launch(dy_f32, x, x_row_stride, x_gs, mean, rstd, gamma, rows_per_group, groups, C, dresid, dx,
                       dx_row_stride, dx_gs, dx_hi, dgamma, dbeta, dcol, partials, rm, as_stream(stream), true, dx_lo);
}

size_t cavit_ln_bwd_workspace_floats(int32_t groups, int32_t C) {
  return (size_t)groups * LN_MAX_BLOCKS_PER_GROUP * 3 * (size_t)C + (size_t)groups;   // partial rows + ticket counters
}

int cavit_ln_bwd(const void* dy, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                 const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                 const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_bf16, float* dgamma,
                 float* dbeta, float* dcol, float* partials, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dx) return fail(CAVIT_E_BADARG, "cavit_ln_bwd: null pointer");
  RowMap rm{};
  rm.fusion = 0;
  return ln_bwd_launch(dy, x, x_row_stride, x_gs, mean, rstd, gamma, rows_per_group, groups, C, dresid, dx,
                       dx_row_stride, dx_gs, dx_bf16, dgamma, dbeta, dcol, partials, rm, as_stream(stream));
}

int cavit_ln_bwd_f32(const float* dy_f32, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                     const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                     const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_bf16, float* dgamma,
                     float* dbeta, float* dcol, float* partials, void* stream) {
  if (!dy_f32 || !x || !mean || !rstd || !gamma || !dx) return fail(CAVIT_E_BADARG, "cavit_ln_bwd_f32: null pointer");
  RowMap rm{};
  rm.fusion = 0;
  return ln_bwd_launch(dy_f32, x, x_row_stride, x_gs, mean, rstd, gamma, rows_per_group, groups, C, dresid, dx,
                       dx_row_stride, dx_gs, dx_bf16, dgamma, dbeta, dcol, partials, rm, as_stream(stream), true);
}

int cavit_ln_fusion_fwd(const float* streams, int64_t stream_gs, const float* x_cls, int32_t B, int32_t N, int32_t C,
                        int32_t K, const int32_t* cls_src, const int32_t* tok_src, const float* gamma, const float* beta,
                        float eps, void* y, float* mean, float* rstd, void* stream) {
  if (!streams || !x_cls || !cls_src || !tok_src || !gamma || !beta || !y) return fail(CAVIT_E_BADARG, "cavit_ln_fusion_fwd: null pointer");
  if (K <= 0 || K > LN_MAX_FUSIONS) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_ln_fusion_fwd: K=%d (max %d)", K, LN_MAX_FUSIONS);
  RowMap rm{};
  rm.fusion = 1;
  rm.N = N;
  rm.B = B;
  rm.x_cls = x_cls;
  for (int k = 0; k < K; ++k) { rm.cls_src[k] = cls_src[k]; rm.tok_src[k] = tok_src[k]; }
  return ln_fwd_launch(streams, C, stream_gs, B * N, K, C, gamma, beta, eps, y, nullptr, mean, rstd, rm, as_stream(stream));
}

int cavit_ln_fusion_bwd(const void* dy, const float* dy_cls, const float* streams, int64_t stream_gs, const float* x_cls,
                        const float* mean, const float* rstd, const float* gamma, int32_t B, int32_t N, int32_t C,
                        int32_t K, const int32_t* cls_src, const int32_t* tok_src, float* dstreams, float* dgamma,
                        float* dbeta, float* partials, void* stream) {
  if (!dy || !streams || !x_cls || !cls_src || !tok_src || !dstreams) return fail(CAVIT_E_BADARG, "cavit_ln_fusion_bwd: null pointer");
  if (K <= 0 || K > LN_MAX_FUSIONS) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_ln_fusion_bwd: K=%d", K);
  RowMap rm{};
  rm.fusion = 1;
  rm.N = N;
  rm.B = B;
  rm.x_cls = x_cls;
  rm.dy_cls = dy_cls;
  for (int k = 0; k < K; ++k) { rm.cls_src[k] = cls_src[k]; rm.tok_src[k] = tok_src[k]; }
  return ln_bwd_launch(dy, streams, C, stream_gs, mean, rstd, gamma, B * N, K, C, nullptr, dstreams, C, stream_gs,
                       nullptr, dgamma, dbeta, nullptr, partials, rm, as_stream(stream));
}

}  // extern "C"
