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
//     current head is being processed (no separate delta kernel);
//   * the control thread prefetches the next head's K/V/Q/dO tiles by TMA as soon as the last MMA reading
//     a buffer has retired, and issues S^T,dP^T of step t+1 as soon as step t's scores are in registers,
//     so the tensor pipe, the TMA engine and the 16 elementwise warps overlap across steps and heads.
//
// Reference semantics: autograd of Attention.forward (/root/reference/model_cross.py:50-61), SURVEY.md §A.9.
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
  int nkv, nh, hN0, hN1;  // key tiles, query halves, columns per half (multiples of 16, <= 128)
  float scale, scale_log2;
  int* status;
};

__device__ __forceinline__ float sb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(SB_THREADS, 1)
attn_bwd_short_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ16,
                      const __grid_constant__ CUtensorMap tmDO16, const AttnBwdShortParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base + SB_OFF_K, sV = base + SB_OFF_V, sQ = base + SB_OFF_Q, sDO = base + SB_OFF_DO;
  const uint32_t sPT = base + SB_OFF_PT, sDST = base + SB_OFF_DST;
  uint8_t* genPT = gen + SB_OFF_PT;
  uint8_t* genDST = gen + SB_OFF_DST;
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
  const long long items = (long long)p.G * BH;
  const int ns = p.nkv * p.nh;

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
    // ================================================================= control warp (lane 0): TMA + MMA issue
    if (lane == 0) {
      const uint32_t idesc_kmn = umma_idesc_bf16(64, 0, 1);   // dV, dK: A K-major (P^T / dS^T), B MN-major, N = 64
      const uint32_t idesc_mnmn = umma_idesc_bf16(64, 1, 1);  // dQ:     A = dS^T viewed MN-major, B = K MN-major
      auto decode = [&](long long item, int& g, int& h, int& row_base) {
        const int bh = (int)(item % BH);
        g = (int)(item / BH);
        h = bh % p.H;
        row_base = (bh / p.H) * p.N;
      };
      auto load_kv = [&](int j, long long item) {
        int g, h, row_base;
        decode(item, g, h, row_base);
        mbar_arrive_expect_tx(bar_kv(j), 2 * SB_TILE);
        tma_load_3d(&tmKV, bar_kv(j), sK + j * SB_TILE, p.C + h * 64, row_base + j * 128, g);
        tma_load_3d(&tmKV, bar_kv(j), sV + j * SB_TILE, 2 * p.C + h * 64, row_base + j * 128, g);
      };
      auto load_q = [&](int hf, long long item) {
        int g, h, row_base;
        decode(item, g, h, row_base);
        const int row0 = hf ? p.hN0 : 0, n = hf ? p.hN1 : p.hN0;
        mbar_arrive_expect_tx(bar_q(hf), 2 * n * 128);
        for (int r = 0; r < n; r += 16) {
          tma_load_3d(&tmQ16, bar_q(hf), sQ + (row0 + r) * 128, h * 64, row_base + row0 + r, g);
          tma_load_3d(&tmDO16, bar_q(hf), sDO + (row0 + r) * 128, h * 64, row_base + row0 + r, g);
        }
      };
      auto issue_s = [&](int j, int hf) {  // S^T = K_j Q_half^T ; dP^T = V_j dO_half^T
        const uint32_t row0 = hf ? p.hN0 : 0;
        const uint32_t idesc = umma_idesc_bf16(hf ? p.hN1 : p.hN0, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t kd = umma_desc_sw128(sK + j * SB_TILE + k * 32, 16, 1024);
          const uint64_t qd = umma_desc_sw128(sQ + row0 * 128 + k * 32, 16, 1024);
          umma_bf16_ss(tST, kd, qd, idesc, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t vd = umma_desc_sw128(sV + j * SB_TILE + k * 32, 16, 1024);
          const uint64_t dd = umma_desc_sw128(sDO + row0 * 128 + k * 32, 16, 1024);
          umma_bf16_ss(tDPT, vd, dd, idesc, k != 0);
        }
        umma_commit(bar_s);
      };
      auto issue_d = [&](int j, int hf) {
        const uint32_t row0 = hf ? p.hN0 : 0;
        const int nq16 = (hf ? p.hN1 : p.hN0) >> 4;
        for (int k = 0; k < nq16; ++k) {  // dV_j[kv][d] += P^T[kv][q] dO[q][d]
          const uint64_t ad = umma_desc_sw128(sPT + (k >> 2) * SB_TILE + (k & 3) * 32, 16, 1024);
          const uint64_t bd = umma_desc_sw128(sDO + (row0 + 16 * k) * 128, SB_TILE, 1024);
          umma_bf16_ss(tDV, ad, bd, idesc_kmn, (hf | k) != 0);
        }
        for (int k = 0; k < nq16; ++k) {  // dK_j[kv][d] += dS^T[kv][q] Q[q][d]
          const uint64_t ad = umma_desc_sw128(sDST + (k >> 2) * SB_TILE + (k & 3) * 32, 16, 1024);
          const uint64_t bd = umma_desc_sw128(sQ + (row0 + 16 * k) * 128, SB_TILE, 1024);
          umma_bf16_ss(tDK, ad, bd, idesc_kmn, (hf | k) != 0);
        }
        const int nk16 = (min(128, p.N - j * 128) + 15) >> 4;  // key rows beyond N hold dS^T = 0: skip their k-steps
        for (int k = 0; k < nk16; ++k) {  // dQ_half[q][d] += dS[q][kv] K_j[kv][d]   (A = dS^T viewed MN-major)
          const uint64_t ad = umma_desc_sw128(sDST + k * 2048, SB_TILE, 1024);
          const uint64_t bd = umma_desc_sw128(sK + j * SB_TILE + k * 2048, SB_TILE, 1024);
          umma_bf16_ss(tDQ + hf * 64, ad, bd, idesc_mnmn, (j | k) != 0);
        }
        umma_commit(bar_d);
      };
      // With two key tiles and two halves every buffer of the next head is refilled at least one step before
      // the step that issues the next head's first S^T; otherwise that issue waits for the last step's prefetch.
      const bool early_next = (p.nkv == 2 && p.nh == 2);
      long long item = blockIdx.x;
      if (item < items) {
        for (int j = 0; j < p.nkv; ++j) load_kv(j, item);
        for (int hf = 0; hf < p.nh; ++hf) load_q(hf, item);
        mbar_wait(bar_kv(0), 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        mbar_wait(bar_q(0), 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        issue_s(0, 0);
      }
      uint32_t t = 0;
      for (int it = 0; item < items; item += gridDim.x, ++it) {
        const long long next_item = item + gridDim.x;
        const bool has_next = next_item < items;
        const uint32_t hp = it & 1, hpn = hp ^ 1u;
        for (int s = 0; s < ns; ++s, ++t) {
          const int j = s / p.nh, hf = s % p.nh;
          // step t's scores are in registers everywhere -> the tensor pipe may overwrite S^T / dP^T
          mbar_wait(bar_sfree, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          bool next_s_pending = false;
          if (s + 1 < ns) {
            const int j2 = (s + 1) / p.nh, h2 = (s + 1) % p.nh;
            mbar_wait(bar_kv(j2), hp, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            mbar_wait(bar_q(h2), hp, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            tc_fence_after();
            issue_s(j2, h2);
          } else if (has_next) {
            if (early_next) {
              mbar_wait(bar_kv(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              mbar_wait(bar_q(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              tc_fence_after();
              issue_s(0, 0);
            } else {
              next_s_pending = true;
            }
          }
          // P^T, dS^T of step t are in shared memory (and the accumulators they overwrite have been drained)
          mbar_wait(bar_p, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          tc_fence_after();
          issue_d(j, hf);
          if (has_next) {
            const bool kv_free = (hf == p.nh - 1), q_free = (j == p.nkv - 1);
            if (kv_free || q_free) {  // refill buffers whose last reader (an MMA of this step) has retired
              mbar_wait(bar_d, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              if (kv_free) load_kv(j, next_item);
              if (q_free) load_q(hf, next_item);
            }
            if (next_s_pending) {
              mbar_wait(bar_kv(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              mbar_wait(bar_q(0), hpn, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              tc_fence_after();
              issue_s(0, 0);
            }
          }
        }
      }
    }
  } else if (warp > SB_EW_WARPS) {
    // ================================================================= helper warps: LSE*log2e and delta of the next head
    const int atid = tid - (SB_EW_WARPS + 1) * 32;  // 0..63
    int ait = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++ait) {
      const int buf = ait & 1;
      if (ait >= 2) mbar_wait(bar_auxfree(buf), ((ait >> 1) - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      const int bh = (int)(item % BH), g = (int)(item / BH);
      const int b = bh / p.H, h = bh % p.H;
      const long long lse_base = (((long long)g * p.B + b) * p.H + h) * p.N;
      const long long row0 = ((long long)g * p.B + b) * p.N;
      float* o_lse = s_aux + buf * 512;
      float* o_del = o_lse + 256;
      for (int q = atid; q < 256; q += SB_AUX_WARPS * 32) {
        float l2 = 0.f, dl = 0.f;
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
    // drains the accumulators that step (pj, phf) of head (pg, pb, ph) finalised
    auto drain = [&](int pg, int pb, int ph, int pj, int phf) {
      const long long rowb = ((long long)pg * p.B + pb) * p.N;
      if (phf == p.nh - 1) {  // dK_j (scaled) and dV_j: each warp stores a 16-column slice of both
        const int kv = pj * 128 + trow;
        bf16* drow = p.dqkv + (rowb + kv) * (3 * p.C) + ph * 64 + part * 16;
#pragma unroll
        for (int which = 0; which < 2; ++which) {  // one 16-column slice at a time (register pressure)
          uint32_t r[16];
          tmem_ld16((which == 0 ? tDK : tDV) + t_lane + part * 16, r);
          tmem_ld_wait();
          const float sc = which == 0 ? p.scale : 1.0f;
          if (kv < p.N) {
#pragma unroll
            for (int c = 0; c < 16; c += 8) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(r[c]) * sc, __uint_as_float(r[c + 1]) * sc);
              w.y = pack_bf16(__uint_as_float(r[c + 2]) * sc, __uint_as_float(r[c + 3]) * sc);
              w.z = pack_bf16(__uint_as_float(r[c + 4]) * sc, __uint_as_float(r[c + 5]) * sc);
              w.w = pack_bf16(__uint_as_float(r[c + 6]) * sc, __uint_as_float(r[c + 7]) * sc);
              *reinterpret_cast<uint4*>(drow + (which + 1) * p.C + c) = w;
            }
          }
        }
      }
      if (pj == p.nkv - 1) {  // dQ of this half is complete (accumulated over the key tiles in TMEM)
        const int hn = phf ? p.hN1 : p.hN0;
        const int q = (phf ? p.hN0 : 0) + trow;
        uint32_t r[16];
        tmem_ld16(tDQ + phf * 64 + t_lane + part * 16, r);
        tmem_ld_wait();
        if (trow < hn && q < p.N) {
          bf16* drow = p.dqkv + (rowb + q) * (3 * p.C) + ph * 64 + part * 16;
#pragma unroll
          for (int c = 0; c < 16; c += 8) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(r[c]) * p.scale, __uint_as_float(r[c + 1]) * p.scale);
            w.y = pack_bf16(__uint_as_float(r[c + 2]) * p.scale, __uint_as_float(r[c + 3]) * p.scale);
            w.z = pack_bf16(__uint_as_float(r[c + 4]) * p.scale, __uint_as_float(r[c + 5]) * p.scale);
            w.w = pack_bf16(__uint_as_float(r[c + 6]) * p.scale, __uint_as_float(r[c + 7]) * p.scale);
            *reinterpret_cast<uint4*>(drow + c) = w;
          }
        }
      }
    };

    uint32_t t = 0;
    int pg = 0, pb = 0, ph = 0, pj = 0, phf = 0;
    int it = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int bh = (int)(item % BH), g = (int)(item / BH);
      const int b = bh / p.H, h = bh % p.H;
      const float* lse2 = s_aux + (it & 1) * 512;
      const float* dlt = lse2 + 256;
      mbar_wait(bar_auxfull(it & 1), (it >> 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      for (int s = 0; s < ns; ++s, ++t) {
        const int j = s / p.nh, hf = s % p.nh;
        const int row0 = hf ? p.hN0 : 0;
        const int ngrp = (hf ? p.hN1 : p.hN0) >> 3;  // groups of 8 query columns in this half
        const int g0 = (part * ngrp) >> 2;
        const int cnt = (((part + 1) * ngrp) >> 2) - g0;  // <= 4 groups for this warp
        mbar_wait(bar_s, t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        // Two passes of <= 2 column groups keep the live register set small: raw scores of a pass are folded into
        // packed bf16 before the next pass is read; S^T / dP^T are released to the tensor pipe after the last read.
        const bool kv_ok = (j * 128 + trow) < p.N;
        uint4 wp[4], wd[4];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          uint32_t rs[2][8], rp[2][8];
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int gi = pass * 2 + gg;
            if (gi < cnt) {
              tmem_ld8(tST + t_lane + (g0 + gi) * 8, rs[gg]);
              tmem_ld8(tDPT + t_lane + (g0 + gi) * 8, rp[gg]);
            }
          }
          tmem_ld_wait();
          if (pass == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sfree);
          }
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int gi = pass * 2 + gg;
            if (gi < cnt) {
              const int q0 = row0 + (g0 + gi) * 8;
              const float4 la = *reinterpret_cast<const float4*>(lse2 + q0), lb = *reinterpret_cast<const float4*>(lse2 + q0 + 4);
              const float4 da = *reinterpret_cast<const float4*>(dlt + q0), db = *reinterpret_cast<const float4*>(dlt + q0 + 4);
              const float l[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
              const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
              float pr[8], ds[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const bool ok = kv_ok && (q0 + c < p.N);
                pr[c] = ok ? sb_ex2(fmaf(__uint_as_float(rs[gg][c]), p.scale_log2, -l[c])) : 0.f;
                ds[c] = pr[c] * (__uint_as_float(rp[gg][c]) - d[c]);
              }
              wp[gi] = make_uint4(pack_bf16(pr[0], pr[1]), pack_bf16(pr[2], pr[3]), pack_bf16(pr[4], pr[5]), pack_bf16(pr[6], pr[7]));
              wd[gi] = make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
            }
          }
        }
        if (t > 0) {  // MMAs of the previous step done: P^T / dS^T buffers are free, its finished accumulators can be drained
          mbar_wait(bar_d, (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          tc_fence_after();
          drain(pg, pb, ph, pj, phf);
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (gi < cnt) {
            const int cg = g0 + gi;
            const uint32_t off = (cg >> 3) * SB_TILE + trow * 128 + (((cg & 7) ^ (trow & 7)) << 4);
            *reinterpret_cast<uint4*>(genPT + off) = wp[gi];
            *reinterpret_cast<uint4*>(genDST + off) = wd[gi];
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        pg = g; pb = b; ph = h; pj = j; phf = hf;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_auxfree(it & 1));
    }
    if (t > 0) {
      mbar_wait(bar_d, (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      tc_fence_after();
      drain(pg, pb, ph, pj, phf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_attn_bwd_short(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int G, int B,
                          int N, int H, float scale, cudaStream_t stream) {
  const int C = H * 64;
  const long long T = (long long)B * N;
  const CUtensorMap* tkv = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, 128);
  const CUtensorMap* tq = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, 16);
  const CUtensorMap* td = tensor_map_bf16_3d(dout, C, T, G, C, T * C, 64, 16);
  if (!tkv || !tq || !td) return CAVIT_E_BADARG;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "attn bwd (short) smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  AttnBwdShortParams p;
  p.lse = lse;
  p.o = reinterpret_cast<const bf16*>(out);
  p.dout = reinterpret_cast<const bf16*>(dout);
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.N = N; p.H = H; p.C = C; p.B = B; p.G = G;
  const int NP = (N + 15) & ~15;
  p.nkv = (N + 127) / 128;
  p.hN0 = NP >= 32 ? ((NP / 2 + 15) & ~15) : NP;
  p.hN1 = NP - p.hN0;
  p.nh = p.hN1 > 0 ? 2 : 1;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "no status word");
  const long long items = (long long)G * B * H;
  const int grid = (int)(items < sm_count() ? items : sm_count());
  attn_bwd_short_kernel<<<grid, SB_THREADS, SB_SMEM, stream>>>(*tkv, *tq, *td, p);
  count_launch();
  return check_launch("cavit_attn_bwd(short)");
}

}  // namespace cavit
