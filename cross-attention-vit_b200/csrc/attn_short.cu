// cavit-sm100 — K-ATTN, short-sequence backward (N <= 256 tokens: every 2-D slice config, e.g. N = 197).
//
// The generic backward in attn.cu launches one CTA per 128-key tile; for N = 197 each CTA lives for two
// query tiles only, so TMEM allocation, barrier set-up, the first TMA round trip and the final drains
// are all exposed, the second query tile is 46 % padding, and dQ needs fp32 atomics plus two helper
// kernels (delta, dQ conversion). Here ONE persistent CTA per SM walks whole heads (stream, sample, head):
//
//   * the query axis is cut into two balanced halves of ceil16(N/2) and the rest (112 + 96 for N = 197)
//     instead of 128 + 128, the key axis into <= 2 tiles of 128: a head is <= 4 "steps" (key tile j, half);
//   * per step:  S^T = K_j Q_half^T, dP^T = V_j dO_half^T  (tcgen05, TMEM)  ->  16 warps form
//     P^T = exp2(S^T*scale*log2e - LSE), dS^T = P^T o (dP^T - delta) in registers -> swizzled smem (bf16)
//     ->  dV_j += P^T dO_half,  dK_j += dS^T Q_half,  dQ_half += dS K_j  (TMEM accumulators);
//   * dQ accumulates over the key tiles INSIDE TMEM, so it is written once, as bf16, straight into the packed
//     dQKV activation: no fp32 atomics, no memset, no conversion kernel;
//   * delta = rowsum(dO o O) and LSE*log2e of the NEXT head are staged by two helper warps while the
//     current head is being processed (no separate delta kernel); padded query columns get LSE = +inf,
//     which makes their probabilities exactly 0 without any per-element masking;
//   * the control thread prefetches the next head's K/V/Q/dO tiles by TMA as soon as the last MMA reading
//     a buffer has retired, and issues S^T,dP^T of step t+1 as soon as step t's scores are in registers,
//     so the tensor pipe, the TMA engine and the 16 elementwise warps overlap across steps and heads.
// The kernel is specialised at compile time on (key tiles, query halves) so that the step loop unrolls.
//
// Reference semantics: autograd of Attention.forward (/root/reference/model_cross.py:50-61), SURVEY.md §A.9.
#include <stdio.h>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int SB_EW_WARPS = 16;
constexpr int SB_AUX_WARPS = 2;
constexpr int SB_THREADS = (SB_EW_WARPS + 1 + SB_AUX_WARPS) * 32;  // 608: warps 0-15 elementwise, 16 control, 17-18 helpers
constexpr int SB_TILE = 16384;                                     // [128 rows][64 bf16], 128B-swizzled
constexpr int SB_OFF_K = 0;                                        // K_0 | K_1
constexpr int SB_OFF_V = 2 * SB_TILE;                              // V_0 | V_1
constexpr int SB_OFF_Q = 4 * SB_TILE;                              // Q rows 0..255
constexpr int SB_OFF_DO = 6 * SB_TILE;                             // dO rows 0..255
constexpr int SB_OFF_PT = 8 * SB_TILE;                             // P^T  [128 keys][2 chunks of 64 queries]
constexpr int SB_OFF_DST = 10 * SB_TILE;                           // dS^T
constexpr int SB_OFF_AUX = 12 * SB_TILE;                           // [2 buffers][lse2 | delta][256] fp32
constexpr int SB_OFF_BAR = SB_OFF_AUX + 4096;
constexpr int SB_SMEM = SB_OFF_BAR + 256 + 1024;

struct AttnBwdShortParams {
  const float* lse;
  const bf16* o;
  const bf16* dout;
  bf16* dqkv;
  int N, H, C, B, G;
  int hN0, hN1;  // query columns per half (multiples of 16, <= 128; hN1 = 0 with a single half)
  float scale, scale_log2;
  int* status;
};

// Optional timeline trace (build with NVCC_EXTRA=-DCAVIT_SB_TRACE): CTA 0 records clock64() at protocol points.
#ifdef CAVIT_SB_TRACE
__device__ unsigned long long g_sb_trace[2][4096];
__device__ __forceinline__ void sb_trace(int role, uint32_t& n, int tag) {
  if (blockIdx.x == 0 && n < 2047) {
    g_sb_trace[role][2 * n] = clock64();
    g_sb_trace[role][2 * n + 1] = tag;
    ++n;
  }
}
#define SB_TRACE(role, n, tag) sb_trace(role, n, tag)
#else
#define SB_TRACE(role, n, tag)
#endif

__device__ __forceinline__ float sb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t sb_desc_add(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// 8 score columns of one key row -> packed bf16 P^T and dS^T
__device__ __forceinline__ void sb_group(const uint32_t (&rs)[8], const uint32_t (&rp)[8], const float* lse2, const float* dlt,
                                         float sl2, uint4& wp, uint4& wd) {
  const float4 la = *reinterpret_cast<const float4*>(lse2), lb = *reinterpret_cast<const float4*>(lse2 + 4);
  const float4 da = *reinterpret_cast<const float4*>(dlt), db = *reinterpret_cast<const float4*>(dlt + 4);
  const float p0 = sb_ex2(fmaf(__uint_as_float(rs[0]), sl2, -la.x)), p1 = sb_ex2(fmaf(__uint_as_float(rs[1]), sl2, -la.y));
  const float p2 = sb_ex2(fmaf(__uint_as_float(rs[2]), sl2, -la.z)), p3 = sb_ex2(fmaf(__uint_as_float(rs[3]), sl2, -la.w));
  const float p4 = sb_ex2(fmaf(__uint_as_float(rs[4]), sl2, -lb.x)), p5 = sb_ex2(fmaf(__uint_as_float(rs[5]), sl2, -lb.y));
  const float p6 = sb_ex2(fmaf(__uint_as_float(rs[6]), sl2, -lb.z)), p7 = sb_ex2(fmaf(__uint_as_float(rs[7]), sl2, -lb.w));
  wp = make_uint4(pack_bf16(p0, p1), pack_bf16(p2, p3), pack_bf16(p4, p5), pack_bf16(p6, p7));
  wd = make_uint4(pack_bf16(p0 * (__uint_as_float(rp[0]) - da.x), p1 * (__uint_as_float(rp[1]) - da.y)),
                  pack_bf16(p2 * (__uint_as_float(rp[2]) - da.z), p3 * (__uint_as_float(rp[3]) - da.w)),
                  pack_bf16(p4 * (__uint_as_float(rp[4]) - db.x), p5 * (__uint_as_float(rp[5]) - db.y)),
                  pack_bf16(p6 * (__uint_as_float(rp[6]) - db.z), p7 * (__uint_as_float(rp[7]) - db.w)));
}

template <int NKV, int NH>
__global__ void __launch_bounds__(SB_THREADS, 1)
attn_bwd_short_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ16,
                      const __grid_constant__ CUtensorMap tmDO16, const AttnBwdShortParams p) {
  constexpr int NS = NKV * NH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base + SB_OFF_K, sV = base + SB_OFF_V, sQ = base + SB_OFF_Q, sDO = base + SB_OFF_DO;
  const uint32_t sPT = base + SB_OFF_PT, sDST = base + SB_OFF_DST;
  float* s_aux = reinterpret_cast<float*>(gen + SB_OFF_AUX);
  const uint32_t bar0 = base + SB_OFF_BAR;
  auto bar_kv = [&](int j) { return bar0 + 8u * j; };
  auto bar_q = [&](int hf) { return bar0 + 16 + 8u * hf; };
  const uint32_t bar_s = bar0 + 32, bar_sfree = bar0 + 40, bar_p = bar0 + 48, bar_d = bar0 + 56;
  auto bar_auxfull = [&](int b) { return bar0 + 64 + 8u * b; };
  auto bar_auxfree = [&](int b) { return bar0 + 80 + 8u * b; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SB_OFF_BAR + 128);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BH = p.B * p.H;
  const int items = p.G * BH;

  if (tid == 0) {
    *abort_flag = 0;
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_kv(i), 1);
      mbar_init(bar_q(i), 1);
      mbar_init(bar_auxfull(i), SB_AUX_WARPS);
      mbar_init(bar_auxfree(i), SB_EW_WARPS);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_sfree, SB_EW_WARPS);
    mbar_init(bar_p, SB_EW_WARPS);
    mbar_init(bar_d, 1);
    fence_barrier_init();
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmQ16);
    prefetch_tmap(&tmDO16);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384) dQ_half0 [384,448) dQ_half1 [448,512)
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;

  if (warp == SB_EW_WARPS) {
    // ================================================================= control warp: TMA + MMA issue
    // The whole warp runs the protocol (waits are executed by all lanes, converged); one elected lane issues.
    {
      const uint32_t idesc_kmn = umma_idesc_bf16(64, 0, 1);   // dV, dK: A K-major (P^T / dS^T), B MN-major, N = 64
      const uint32_t idesc_mnmn = umma_idesc_bf16(64, 1, 1);  // dQ:     A = dS^T viewed MN-major, B = K MN-major
      const uint32_t idesc_s0 = umma_idesc_bf16(p.hN0, 0, 0), idesc_s1 = umma_idesc_bf16(NH == 2 ? p.hN1 : p.hN0, 0, 0);
      // descriptor bases (byte offsets are added in units of 16 B)
      const uint64_t dK_k = umma_desc_sw128(sK, 16, 1024), dV_k = umma_desc_sw128(sV, 16, 1024);       // K-major reads
      const uint64_t dQ_k = umma_desc_sw128(sQ, 16, 1024), dDO_k = umma_desc_sw128(sDO, 16, 1024);
      const uint64_t dPT_k = umma_desc_sw128(sPT, 16, 1024), dDST_k = umma_desc_sw128(sDST, 16, 1024);
      const uint64_t dK_mn = umma_desc_sw128(sK, SB_TILE, 1024), dQ_mn = umma_desc_sw128(sQ, SB_TILE, 1024);  // MN-major reads
      const uint64_t dDO_mn = umma_desc_sw128(sDO, SB_TILE, 1024), dDST_mn = umma_desc_sw128(sDST, SB_TILE, 1024);
      const int nq16_0 = p.hN0 >> 4, nq16_1 = p.hN1 >> 4;
      const int nk16_last = (min(128, p.N - (NKV - 1) * 128) + 15) >> 4;  // key rows beyond ceil16(N) are never read
      const uint32_t rowb1 = p.hN0 * 128;
      auto decode = [&](int item, int& g, int& h, int& row_base) {
        const int bh = item % BH;
        g = item / BH;
        h = bh % p.H;
        row_base = (bh / p.H) * p.N;
      };
      auto load_kv = [&](int j, int g, int h, int row_base) {
        mbar_arrive_expect_tx(bar_kv(j), 2 * SB_TILE);
        tma_load_3d(&tmKV, bar_kv(j), sK + j * SB_TILE, p.C + h * 64, row_base + j * 128, g);
        tma_load_3d(&tmKV, bar_kv(j), sV + j * SB_TILE, 2 * p.C + h * 64, row_base + j * 128, g);
      };
      auto load_q = [&](int hf, int g, int h, int row_base) {
        const int row0 = hf ? p.hN0 : 0, n = hf ? p.hN1 : p.hN0;
        mbar_arrive_expect_tx(bar_q(hf), 2 * n * 128);
        for (int r = 0; r < n; r += 16) {
          tma_load_3d(&tmQ16, bar_q(hf), sQ + (row0 + r) * 128, h * 64, row_base + row0 + r, g);
          tma_load_3d(&tmDO16, bar_q(hf), sDO + (row0 + r) * 128, h * 64, row_base + row0 + r, g);
        }
      };
      auto issue_s = [&](int j, int hf) {  // S^T = K_j Q_half^T ; dP^T = V_j dO_half^T
        const uint32_t rowb = hf ? rowb1 : 0u;
        const uint32_t idesc = hf ? idesc_s1 : idesc_s0;
        const uint64_t a0 = sb_desc_add(dK_k, j * SB_TILE), b0 = sb_desc_add(dQ_k, rowb);
        const uint64_t a1 = sb_desc_add(dV_k, j * SB_TILE), b1 = sb_desc_add(dDO_k, rowb);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tST, a0 + 2 * k, b0 + 2 * k, idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tDPT, a1 + 2 * k, b1 + 2 * k, idesc, k != 0);
        umma_commit(bar_s);
      };
      auto issue_d = [&](int j, int hf) {
        const uint32_t rowb = hf ? rowb1 : 0u;
        const int nq16 = hf ? nq16_1 : nq16_0;
        const uint64_t bdo = sb_desc_add(dDO_mn, rowb), bq = sb_desc_add(dQ_mn, rowb), bk = sb_desc_add(dK_mn, j * SB_TILE);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV_j[kv][d] += P^T[kv][q] dO[q][d]
          if (k < nq16)
            umma_bf16_ss(tDV, dPT_k + (((k >> 2) * SB_TILE + (k & 3) * 32) >> 4), bdo + k * 128, idesc_kmn, (hf | k) != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK_j[kv][d] += dS^T[kv][q] Q[q][d]
          if (k < nq16)
            umma_bf16_ss(tDK, dDST_k + (((k >> 2) * SB_TILE + (k & 3) * 32) >> 4), bq + k * 128, idesc_kmn, (hf | k) != 0);
        const int nk16 = (j == NKV - 1) ? nk16_last : 8;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ_half[q][d] += dS[q][kv] K_j[kv][d]   (A = dS^T viewed MN-major)
          if (k < nk16) umma_bf16_ss(tDQ + hf * 64, dDST_mn + k * 128, bk + k * 128, idesc_mnmn, (j | k) != 0);
        umma_commit(bar_d);
      };
      // With two key tiles and two halves every buffer of the next head is refilled at least one step before
      // the step that issues the next head's first S^T; otherwise that issue waits for the last step's prefetch.
      constexpr bool kEarlyNext = (NKV == 2 && NH == 2);
      int item = blockIdx.x;
      int g, h, row_base;
      if (item < items) {
        decode(item, g, h, row_base);
        if (elect_one()) {
          for (int j = 0; j < NKV; ++j) load_kv(j, g, h, row_base);
          for (int hf = 0; hf < NH; ++hf) load_q(hf, g, h, row_base);
        }
        __syncwarp();
        mbar_wait(bar_kv(0), 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        mbar_wait(bar_q(0), 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        if (elect_one()) issue_s(0, 0);
        __syncwarp();
      }
      uint32_t t = 0;
      [[maybe_unused]] uint32_t ntr = (lane == 0) ? 0u : 4096u;
      for (int it = 0; item < items; item += gridDim.x, ++it) {
        const int next_item = item + gridDim.x;
        const bool has_next = next_item < items;
        if (has_next) decode(next_item, g, h, row_base);
        const uint32_t hp = it & 1, hpn = hp ^ 1u;
#pragma unroll
        for (int s = 0; s < NS; ++s, ++t) {
          const int j = s / NH, hf = s % NH;
          // step t's scores are in registers everywhere -> the tensor pipe may overwrite S^T / dP^T
          SB_TRACE(0, ntr, 0);
          mbar_wait(bar_sfree, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          SB_TRACE(0, ntr, 1);
          if (s + 1 < NS) {
            const int j2 = (s + 1) / NH, h2 = (s + 1) % NH;
            mbar_wait(bar_kv(j2), hp, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            mbar_wait(bar_q(h2), hp, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            tc_fence_after();
            if (elect_one()) issue_s(j2, h2);
            __syncwarp();
          } else if (kEarlyNext && has_next) {
            mbar_wait(bar_kv(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            mbar_wait(bar_q(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            tc_fence_after();
            if (elect_one()) issue_s(0, 0);
            __syncwarp();
          }
          // P^T, dS^T of step t are in shared memory (and the accumulators they overwrite have been drained)
          SB_TRACE(0, ntr, 2);
          mbar_wait(bar_p, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          SB_TRACE(0, ntr, 3);
          tc_fence_after();
          if (elect_one()) issue_d(j, hf);
          __syncwarp();
          SB_TRACE(0, ntr, 4);
#ifdef CAVIT_SB_TRACE
          mbar_wait(bar_d, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          SB_TRACE(0, ntr, 5);
#endif
          if (has_next) {
            const bool kv_free = (hf == NH - 1), q_free = (j == NKV - 1);
            if (kv_free || q_free) {  // refill buffers whose last reader (an MMA of this step) has retired
              mbar_wait(bar_d, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              if (elect_one()) {
                if (kv_free) load_kv(j, g, h, row_base);
                if (q_free) load_q(hf, g, h, row_base);
              }
              __syncwarp();
              SB_TRACE(0, ntr, 6);
            }
            if (!kEarlyNext && s + 1 == NS) {
              mbar_wait(bar_kv(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              mbar_wait(bar_q(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              tc_fence_after();
              if (elect_one()) issue_s(0, 0);
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp > SB_EW_WARPS) {
    // ================================================================= helper warps: LSE*log2e and delta of the next head
    const int atid = tid - (SB_EW_WARPS + 1) * 32;  // 0..63
    int ait = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++ait) {
      const int buf = ait & 1;
      if (ait >= 2) mbar_wait(bar_auxfree(buf), ((ait >> 1) - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      const int bh = item % BH, g = item / BH;
      const int b = bh / p.H, h = bh % p.H;
      const long long lse_base = (((long long)g * p.B + b) * p.H + h) * p.N;
      const long long row0 = ((long long)g * p.B + b) * p.N;
      float* o_lse = s_aux + buf * 512;
      float* o_del = o_lse + 256;
      for (int q = atid; q < 256; q += SB_AUX_WARPS * 32) {
        float l2 = INFINITY, dl = 0.f;  // padded query columns: exp2(s - inf) = 0 exactly, so P = dS = 0 there
        if (q < p.N) {
          l2 = p.lse[lse_base + q] * 1.4426950408889634f;
          const uint4* po = reinterpret_cast<const uint4*>(p.o + (row0 + q) * p.C + h * 64);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (row0 + q) * p.C + h * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 a = __ldg(po + i), d = __ldg(pd + i);
            const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
            const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
            dl += a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y;
          }
        }
        o_lse[q] = l2;
        o_del[q] = dl;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_auxfull(buf));
    }
  } else {
    // ================================================================= elementwise warps
    const int quad = warp & 3, part = warp >> 2;
    const int trow = quad * 32 + lane;  // key row inside the tile = TMEM lane (scores) / query row (dQ)
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_col = t_lane + part * 16;                 // this warp's 16-column slice of a 64-wide accumulator
    uint8_t* const myPT = gen + SB_OFF_PT + trow * 128;        // this key row inside chunk 0 of P^T
    const uint32_t sw = static_cast<uint32_t>(trow & 7);
    const float sl2 = p.scale_log2, scale = p.scale;
    const int N = p.N;
    // last key tile: rows >= N are masked to zero; warps entirely beyond ceil16(valid rows) are never read by an MMA
    const int last_valid = N - (NKV - 1) * 128;
    const bool row_ok_last = trow < last_valid;
    const bool warp_active_last = quad * 32 < ((last_valid + 15) & ~15);
    // Drains the accumulators finalised by step (pj, phf) of the head whose dQKV rows start at `hb`. A 128 x 64 accumulator
    // block goes through a shared-memory staging tile so that the global stores are whole 128-byte rows (8 lanes x 16 B):
    // storing each thread's 32-byte slice directly made every warp instruction touch 32 different rows of the packed
    // [T][3C] gradient (2304-byte stride) — 50 % excess sectors under ncu and 2 - 3.7 k cycles per drain in the clock64
    // timeline, a third of a head's 23 k. Staging tiles: the P^T / dS^T buffers, which are free exactly here (the MMAs
    // of the previous step have retired — bar_d — and this step's P^T / dS^T are stored after the drain). The four warps of a
    // TMEM lane quadrant (parts 0 .. 3) own the rows of that quadrant in every tile — staging rows and P^T rows alike — and
    // meet at a named barrier before and after the copy-out, so no warp overwrites rows another one is still copying.
    auto stage_block = [&](uint32_t taddr, float sc, uint8_t* stg) {
      uint32_t r[16];
      tmem_ld16(taddr + t_col, r);
      tmem_ld_wait();
      uint4 w0, w1;
      w0.x = pack_bf16(__uint_as_float(r[0]) * sc, __uint_as_float(r[1]) * sc);
      w0.y = pack_bf16(__uint_as_float(r[2]) * sc, __uint_as_float(r[3]) * sc);
      w0.z = pack_bf16(__uint_as_float(r[4]) * sc, __uint_as_float(r[5]) * sc);
      w0.w = pack_bf16(__uint_as_float(r[6]) * sc, __uint_as_float(r[7]) * sc);
      w1.x = pack_bf16(__uint_as_float(r[8]) * sc, __uint_as_float(r[9]) * sc);
      w1.y = pack_bf16(__uint_as_float(r[10]) * sc, __uint_as_float(r[11]) * sc);
      w1.z = pack_bf16(__uint_as_float(r[12]) * sc, __uint_as_float(r[13]) * sc);
      w1.w = pack_bf16(__uint_as_float(r[14]) * sc, __uint_as_float(r[15]) * sc);
      uint8_t* srow = stg + trow * 128;
      *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(2 * part) ^ sw) << 4)) = w0;
      *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(2 * part + 1) ^ sw) << 4)) = w1;
    };
    auto copy_block = [&](const uint8_t* stg, bf16* out0, int rows_valid) {   // this warp: rows quad*32 + part*8 .. + 7
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int rr = quad * 32 + part * 8 + i * 4 + (lane >> 3);
        if (rr < rows_valid) {
          const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((static_cast<uint32_t>(lane & 7) ^ static_cast<uint32_t>(rr & 7)) << 4));
          *reinterpret_cast<uint4*>(out0 + (long long)rr * (3 * p.C) + (lane & 7) * 8) = v;
        }
      }
    };
    auto drain = [&](bf16* hb, int pj, int phf) {
      const bool kv_done = (phf == NH - 1);   // dK_j (scaled) and dV_j of the key tile
      const bool q_done = (pj == NKV - 1);    // dQ of this half (accumulated over the key tiles in TMEM)
      if (!kv_done && !q_done) return;
      uint8_t* const stg_k = gen + SB_OFF_PT;
      uint8_t* const stg_v = gen + SB_OFF_PT + SB_TILE;
      uint8_t* const stg_q = gen + SB_OFF_DST;
      if (kv_done) {
        stage_block(tDK, scale, stg_k);
        stage_block(tDV, 1.0f, stg_v);
      }
      if (q_done) stage_block(tDQ + phf * 64, scale, stg_q);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
      if (kv_done) {
        bf16* blk = hb + (long long)(pj * 128) * (3 * p.C);
        const int rows_valid = min(128, N - pj * 128);
        copy_block(stg_k, blk + p.C, rows_valid);
        copy_block(stg_v, blk + 2 * p.C, rows_valid);
      }
      if (q_done) {
        const int hn = phf ? p.hN1 : p.hN0, q0 = phf ? p.hN0 : 0;
        copy_block(stg_q, hb + (long long)q0 * (3 * p.C), min(hn, N - q0));
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");   // the staging rows become P^T / dS^T rows again
    };

    uint32_t t = 0;
    [[maybe_unused]] uint32_t ntr = (tid == 0) ? 0u : 4096u;
    bf16* prev_hb = nullptr;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int bh = item % BH, g = item / BH;
      const int b = bh / p.H, h = bh % p.H;
      bf16* const hb = p.dqkv + (((long long)g * p.B + b) * N) * (3 * p.C) + h * 64;
      const float* lse2 = s_aux + (it & 1) * 512;
      mbar_wait(bar_auxfull(it & 1), (it >> 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
#pragma unroll
      for (int s = 0; s < NS; ++s, ++t) {
        const int j = s / NH, hf = s % NH;
        const int row0 = hf ? p.hN0 : 0;
        const int ngrp = (hf ? p.hN1 : p.hN0) >> 3;  // groups of 8 query columns in this half
        const int g0 = (part * ngrp) >> 2;
        const int cnt = (((part + 1) * ngrp) >> 2) - g0;  // <= 4 groups for this warp
        const bool active = (j < NKV - 1) || warp_active_last;
        SB_TRACE(1, ntr, 10);
        mbar_wait(bar_s, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        SB_TRACE(1, ntr, 11);
        tc_fence_after();
        uint4 wp[4], wd[4];
        // Two passes of <= 2 column groups keep the live register set small; S^T / dP^T are released to the tensor
        // pipe right after the last read.
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          uint32_t rs[2][8], rp[2][8];
          if (active) {
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              const int gi = pass * 2 + gg;
              if (gi < cnt) {
                tmem_ld8(tST + t_lane + (g0 + gi) * 8, rs[gg]);
                tmem_ld8(tDPT + t_lane + (g0 + gi) * 8, rp[gg]);
              }
            }
            tmem_ld_wait();
          }
          if (pass == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sfree);
          }
          if (active) {
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              const int gi = pass * 2 + gg;
              if (gi < cnt) {
                const float* l = lse2 + row0 + (g0 + gi) * 8;
                sb_group(rs[gg], rp[gg], l, l + 256, sl2, wp[gi], wd[gi]);
                if (j == NKV - 1 && !row_ok_last) {
                  wp[gi] = make_uint4(0u, 0u, 0u, 0u);
                  wd[gi] = make_uint4(0u, 0u, 0u, 0u);
                }
              }
            }
          }
        }
        SB_TRACE(1, ntr, 12);
        // MMAs of the previous step done: P^T / dS^T buffers are free, accumulators it finalised can be drained
        if (s > 0) {
          mbar_wait(bar_d, (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          tc_fence_after();
          SB_TRACE(1, ntr, 13);
          drain(hb, (s - 1) / NH, (s - 1) % NH);
        } else if (t > 0) {
          mbar_wait(bar_d, (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          tc_fence_after();
          SB_TRACE(1, ntr, 13);
          drain(prev_hb, NKV - 1, NH - 1);
        }
        SB_TRACE(1, ntr, 14);
        if (active) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            if (gi < cnt) {
              const uint32_t cg = g0 + gi;
              uint8_t* dst = myPT + (cg >> 3) * SB_TILE + (((cg & 7) ^ sw) << 4);
              *reinterpret_cast<uint4*>(dst) = wp[gi];
              *reinterpret_cast<uint4*>(dst + (SB_OFF_DST - SB_OFF_PT)) = wd[gi];
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        SB_TRACE(1, ntr, 15);
      }
      prev_hb = hb;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_auxfree(it & 1));
    }
    if (t > 0) {
      mbar_wait(bar_d, (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      tc_fence_after();
      drain(prev_hb, NKV - 1, NH - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int NKV, int NH>
static int launch_short(const CUtensorMap* tkv, const CUtensorMap* tq, const CUtensorMap* td, const AttnBwdShortParams& p,
                        int grid, cudaStream_t stream) {
  static PerDeviceFlag attr;
  if (attr.unset()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_short_kernel<NKV, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "attn bwd (short) smem attribute: %s", cudaGetErrorString(e));
    attr.set();
  }
  attn_bwd_short_kernel<NKV, NH><<<grid, SB_THREADS, SB_SMEM, stream>>>(*tkv, *tq, *td, p);
  return CAVIT_OK;
}

int launch_attn_bwd_short(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int G, int B,
                          int N, int H, float scale, cudaStream_t stream) {
  const int C = H * 64;
  const long long T = (long long)B * N;
  if ((long long)G * B * H > 0x7fffffffLL) return fail(CAVIT_E_BADARG, "cavit_attn_bwd: too many heads");
  const CUtensorMap* tkv = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, 128);
  const CUtensorMap* tq = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, 16);
  const CUtensorMap* td = tensor_map_bf16_3d(dout, C, T, G, C, T * C, 64, 16);
  if (!tkv || !tq || !td) return CAVIT_E_BADARG;
  AttnBwdShortParams p;
  p.lse = lse;
  p.o = reinterpret_cast<const bf16*>(out);
  p.dout = reinterpret_cast<const bf16*>(dout);
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.N = N; p.H = H; p.C = C; p.B = B; p.G = G;
  const int NP = (N + 15) & ~15;
  const int nkv = (N + 127) / 128;
  p.hN0 = NP >= 32 ? ((NP / 2 + 15) & ~15) : NP;
  p.hN1 = NP - p.hN0;
  const int nh = p.hN1 > 0 ? 2 : 1;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "no status word");
  const int items = G * B * H;
  const int grid = items < sm_count() ? items : sm_count();
  int rc;
  if (nkv == 2 && nh == 2) rc = launch_short<2, 2>(tkv, tq, td, p, grid, stream);
  else if (nkv == 1 && nh == 2) rc = launch_short<1, 2>(tkv, tq, td, p, grid, stream);
  else if (nkv == 1 && nh == 1) rc = launch_short<1, 1>(tkv, tq, td, p, grid, stream);
  else return fail(CAVIT_E_UNSUPPORTED_SHAPE, "attn bwd (short): N=%d", N);
  if (rc) return rc;
  count_launch();
#ifdef CAVIT_SB_TRACE
  {
    static int dumps = 0;
    if (++dumps == 3) {
      cudaStreamSynchronize(stream);
      static unsigned long long h[2][4096];
      cudaMemcpyFromSymbol(h, g_sb_trace, sizeof(h));
      for (int role = 0; role < 2; ++role)
        for (int i = 0; i < 400; ++i) fprintf(stderr, "SBTRACE %d %d %llu %llu\n", role, i, h[role][2 * i] - h[0][0], h[role][2 * i + 1]);
    }
  }
#endif
  return check_launch("cavit_attn_bwd(short)");
}

}  // namespace cavit
