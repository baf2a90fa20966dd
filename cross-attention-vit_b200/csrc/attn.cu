// cavit-sm100 — K-ATTN: fused self-attention softmax(Q K^T * scale) V for head_dim 64 on tcgen05.
//
// Forward (one CTA per 128 query rows of one (stream, sample, head)):
//   S = Q K_j^T            tcgen05.mma 128x128x64 into TMEM          (K-major A and B)
//   online softmax         one thread per query row reads its S row with tcgen05.ld, keeps the
//                          running max / sum in registers, writes P (bf16) into a 128B-swizzled
//                          shared tile
//   O_j = P V_j            tcgen05.mma 128x64x128 (A = P from smem, B = V as MN-major operand)
//   O = O*alpha + O_j      accumulated in registers, normalised and stored as merged heads.
// The N x N score matrix exists only in TMEM / shared memory. K/V tiles are double-buffered TMA
// loads straight out of the packed QKV activation ([T][3C], exactly what the QKV GEMM wrote), so
// no head-split permute is ever materialised ('b n (h d) -> b h n d', model_cross.py:53).
//
// Backward (one CTA per 128 key/value rows; loops over query tiles), SURVEY.md §A.9:
//   S^T = K Q_i^T, dP^T = V dO_i^T          (TMEM)
//   P^T = exp(S^T*scale - LSE), dS^T = P^T o (dP^T - D)    (registers -> swizzled smem, bf16)
//   dV += P^T dO_i, dK += dS^T Q_i          (TMEM accumulators, B operands MN-major)
//   dQ_i = dS K                             (A = dS^T tile viewed MN-major) -> fp32 red.global.add
//
// Ragged tails (N = tile + 1): out-of-range keys get P = 0; out-of-range query rows are not stored.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int ATT_THREADS = 128;
constexpr int ATT_TILE = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_D * 2;  // 16 KB: [128 rows][64 bf16]

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Store 32 consecutive bf16 values (cols c0..c0+31 of row r) into a [128][128] bf16 tile kept as two
// K-major SWIZZLE_128B chunks of 64 columns (chunk pitch 16 KB, row pitch 128 B).
__device__ __forceinline__ void store_row_chunk_sw128(uint8_t* tile, int r, int c0, const float (&v)[32]) {
  const int half = c0 >> 6;
  const int c16_0 = (c0 & 63) >> 3;  // first 16-byte chunk inside the 128-byte row
  uint8_t* row = tile + half * ATT_TILE_BYTES + r * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q;
    q.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    q.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    q.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    q.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(row + (((c16_0 + i) ^ (r & 7)) << 4)) = q;
  }
}

// Same for 16 consecutive values (c0 multiple of 16).
__device__ __forceinline__ void store_row_16_sw128(uint8_t* tile, int r, int c0, const float (&v)[16]) {
  const int half = c0 >> 6;
  const int c16_0 = (c0 & 63) >> 3;
  uint8_t* row = tile + half * ATT_TILE_BYTES + r * 128;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 q;
    q.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    q.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    q.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    q.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(row + (((c16_0 + i) ^ (r & 7)) << 4)) = q;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Optional timeline trace (build with NVCC_EXTRA=-DCAVIT_FWD_TRACE): CTA 0 records clock64() at protocol points of the
// TMA producer (role 0), the MMA issuer (1) and lane 0 of the first softmax warp of each slot (2, 3).
#ifdef CAVIT_FWD_TRACE
__device__ unsigned long long g_fwd_trace[4][2048];
__device__ __forceinline__ void fwd_trace(int role, uint32_t& n, int tag) {
  if (blockIdx.x == 0 && n < 1024) {
    g_fwd_trace[role][2 * n] = clock64();
    g_fwd_trace[role][2 * n + 1] = tag;
    ++n;
  }
}
#define FWD_TRACE(role, n, tag) fwd_trace(role, n, tag)
#else
#define FWD_TRACE(role, n, tag)
#endif

struct AttnFwdParams {
  bf16* out;
  float* lse;
  int N, H, C, B, G;
  long long out_gs;  // elements between groups in out
  float scale_log2;  // scale * log2(e)
  int* status;
};

// Forward, persistent + pipelined ("ping-pong"). One CTA per SM walks work items = 256 query rows (two
// 128-row tiles, "slots") of one (stream, sample, head). 18 warps:
//   warps 0-7   softmax group of slot 0     (quad = TMEM lane quadrant, part = 64-column half of the key tile)
//   warps 8-15  softmax group of slot 1
//   warp 16     TMA producer: Q0,Q1 per item; K/V tiles through a 3-stage ring shared by both slots
//   warp 17     MMA issuer:   S[s] = Q[s] K_j^T, O[s] = P[s] V_j  (issue order PV0(j), S0(j+1), PV1(j), S1(j+1))
// While one slot's warps run the softmax of tile j the tensor pipe works for the other slot, and the
// producer runs up to 3 K/V tiles ahead (also across work items), so neither TMA latency nor MMA
// latency sits on a softmax warp's critical path. O_j is produced fresh in TMEM and folded into
// register accumulators one tile later (o = o*alpha + O_j), so there is no TMEM read-modify-write.
// smem: Q0 | Q1 | (K,V) x 3 | P0(2 chunks) | P1(2 chunks) | max/sum exchange [2 slots][2 parts][128] x2 | barriers
constexpr int ATT_FWD_THREADS = 18 * 32;
constexpr int ATT_FWD_KV_STAGES = 3;
constexpr int ATT_FWD_SMEM = (2 + 2 * ATT_FWD_KV_STAGES + 4) * ATT_TILE_BYTES + 4096 + 1024 + 256;

__global__ void __launch_bounds__(ATT_FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;                                       // 2 tiles
  const uint32_t sKV = base + 2 * ATT_TILE_BYTES;                 // stage: K then V
  const uint32_t sP = base + (2 + 2 * ATT_FWD_KV_STAGES) * ATT_TILE_BYTES;  // slot s at + s * 2 tiles
  uint8_t* genP = gen + (2 + 2 * ATT_FWD_KV_STAGES) * ATT_TILE_BYTES;
  float* s_xchg = reinterpret_cast<float*>(gen + (6 + 2 * ATT_FWD_KV_STAGES) * ATT_TILE_BYTES);  // [slot][max|sum][part][128]
  const uint32_t bar0 = base + (6 + 2 * ATT_FWD_KV_STAGES) * ATT_TILE_BYTES + 4096;
  const uint32_t bar_qfull = bar0, bar_qempty = bar0 + 8;
  auto bar_kvfull = [&](int st) { return bar0 + 16 + 8u * st; };
  auto bar_kvempty = [&](int st) { return bar0 + 16 + 8u * (ATT_FWD_KV_STAGES + st); };
  const uint32_t bar_x = bar0 + 16 + 16 * ATT_FWD_KV_STAGES;  // then per slot: s_full, s_free, p_full, o_full
  auto bar_sfull = [&](int sl) { return bar_x + 32u * sl; };
  auto bar_sfree = [&](int sl) { return bar_x + 32u * sl + 8; };
  auto bar_pfull = [&](int sl) { return bar_x + 32u * sl + 16; };
  auto bar_ofull = [&](int sl) { return bar_x + 32u * sl + 24; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (6 + 2 * ATT_FWD_KV_STAGES) * ATT_TILE_BYTES + 4096 + 192);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkv = (p.N + ATT_TILE - 1) / ATT_TILE;
  const int nqp = (p.N + 2 * ATT_TILE - 1) / (2 * ATT_TILE);
  const int BH = p.B * p.H;
  const long long items = (long long)p.G * BH * nqp;

  if (tid == 0) {
    *abort_flag = 0;
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    for (int st = 0; st < ATT_FWD_KV_STAGES; ++st) {
      mbar_init(bar_kvfull(st), 1);
      mbar_init(bar_kvempty(st), 1);
    }
    for (int sl = 0; sl < 2; ++sl) {
      mbar_init(bar_sfull(sl), 1);
      mbar_init(bar_sfree(sl), 8);
      mbar_init(bar_pfull(sl), 8);
      mbar_init(bar_ofull(sl), 1);
    }
    fence_barrier_init();
    prefetch_tmap(&tmQKV);
  }
  if (warp == 16) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

  if (warp == 16) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      int it = 0;
      uint32_t ntr = 0;
      (void)ntr;
      for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int qp = (int)(item % nqp);
        const long long gbh = item / nqp;
        const int bh = (int)(gbh % BH), g = (int)(gbh / BH);
        const int b = bh / p.H, h = bh % p.H;
        const int row_base = b * p.N;
        FWD_TRACE(0, ntr, 0);
        {   // L2 prefetch of the NEXT item's Q and first key / value tiles: one whole item ahead of their smem loads
          const long long nitem = item + gridDim.x;
          if (nitem < items) {
            const int nqp_i = (int)(nitem % nqp);
            const long long ngbh = nitem / nqp;
            const int nbh = (int)(ngbh % BH), ng = (int)(ngbh / BH);
            const int nrow = (nbh / p.H) * p.N, nh = nbh % p.H;
            tma_prefetch_3d(&tmQKV, nh * ATT_D, nrow + nqp_i * 2 * ATT_TILE, ng);
            tma_prefetch_3d(&tmQKV, nh * ATT_D, nrow + nqp_i * 2 * ATT_TILE + ATT_TILE, ng);
            const int npf = nkv < 3 ? nkv : 3;
            for (int j = 0; j < npf; ++j) {
              tma_prefetch_3d(&tmQKV, p.C + nh * ATT_D, nrow + j * ATT_TILE, ng);
              tma_prefetch_3d(&tmQKV, 2 * p.C + nh * ATT_D, nrow + j * ATT_TILE, ng);
            }
          }
        }
        mbar_wait(bar_qempty, (it & 1) ^ 1u, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        FWD_TRACE(0, ntr, 1);
        mbar_arrive_expect_tx(bar_qfull, 2 * ATT_TILE_BYTES);
        tma_load_3d(&tmQKV, bar_qfull, sQ, h * ATT_D, row_base + qp * 2 * ATT_TILE, g);
        tma_load_3d(&tmQKV, bar_qfull, sQ + ATT_TILE_BYTES, h * ATT_D, row_base + qp * 2 * ATT_TILE + ATT_TILE, g);
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(bar_kvempty(st), ph ^ 1u, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          FWD_TRACE(0, ntr, 2);
          mbar_arrive_expect_tx(bar_kvfull(st), 2 * ATT_TILE_BYTES);
          tma_load_3d(&tmQKV, bar_kvfull(st), sKV + st * 2 * ATT_TILE_BYTES, p.C + h * ATT_D, row_base + j * ATT_TILE, g);
          tma_load_3d(&tmQKV, bar_kvfull(st), sKV + st * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES, 2 * p.C + h * ATT_D,
                      row_base + j * ATT_TILE, g);
          if (++st == ATT_FWD_KV_STAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 17) {
    // ================================================================= MMA issuer
    // The whole warp runs the schedule converged (every lane executes the waits); one elected lane issues, so that the
    // compiler emits ELECT + back-to-back predicated UTCHMMA instead of re-deriving a single active thread per instruction.
    {
      const uint32_t idesc_s = umma_idesc_bf16(128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(64, 0, 1);
      const uint64_t dQ[2] = {umma_desc_sw128(sQ, 16, 1024), umma_desc_sw128(sQ + ATT_TILE_BYTES, 16, 1024)};
      const uint64_t dP[2] = {umma_desc_sw128(sP, 16, 1024), umma_desc_sw128(sP + 2 * ATT_TILE_BYTES, 16, 1024)};
      const uint64_t dK0 = umma_desc_sw128(sKV, 16, 1024);
      const uint64_t dV0 = umma_desc_sw128(sKV + ATT_TILE_BYTES, ATT_TILE_BYTES, 1024);
      int st = 0;
      uint32_t ph = 0;
      uint32_t t = 0;  // global tile counter of this CTA (same for both slots)
      int it = 0;
      uint32_t ntr = (lane == 0) ? 0u : 4096u;   // trace: lane 0 only
      (void)ntr;
      auto issue_s = [&](int sl, int stage) {
        // S[sl] may be overwritten once the slot's softmax warps have read the previous tile
        mbar_wait(bar_sfree(sl), (t & 1) ^ 1u, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        FWD_TRACE(1, ntr, 10 + sl);
        tc_fence_after();
        // descriptors advance by constant adds on the 16-byte-granular start-address field (all of shared memory fits it)
        const uint64_t ad0 = dQ[sl], bd0 = dK0 + static_cast<uint32_t>(stage * (2 * ATT_TILE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tmem + sl * 128, ad0 + 2 * k, bd0 + 2 * k, idesc_s, k != 0);
          umma_commit(bar_sfull(sl));
        }
        __syncwarp();
      };
      auto issue_pv = [&](int sl, int stage, uint32_t tt) {
        mbar_wait(bar_pfull(sl), tt & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        FWD_TRACE(1, ntr, 12 + sl);
        tc_fence_after();
        const uint64_t ad0 = dP[sl], bd0 = dV0 + static_cast<uint32_t>(stage * (2 * ATT_TILE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_TILE / 16; ++k)
            umma_bf16_ss(tmem + 256 + sl * 64, ad0 + ((k >> 2) * (ATT_TILE_BYTES >> 4) + (k & 3) * 2), bd0 + k * (2048 >> 4),
                         idesc_o, k != 0);
          umma_commit(bar_ofull(sl));
        }
        __syncwarp();
      };
      auto commit = [&](uint32_t bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      // The schedule is software-pipelined ACROSS work items: after the P.V of a slot's last key tile the next item's first
      // S for that slot is issued at once (its Q arrived while this item was running), so a slot never waits for the other
      // slot's last tile at an item boundary and the two slots stay half a period apart.
      bool first = true;
      for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const bool has_next = item + gridDim.x < items;
        FWD_TRACE(1, ntr, 14);
        if (first) {
          mbar_wait(bar_qfull, it & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          mbar_wait(bar_kvfull(st), ph, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          issue_s(0, st);
          issue_s(1, st);
          // Q is only read by the S MMAs: it is released as soon as the item's LAST S MMAs are issued, so that the producer
          // loads the next item's Q (and runs ahead on its K/V) while this item's softmax / P.V still run.
          if (nkv == 1) commit(bar_qempty);
          first = false;
        }
        for (int j = 0; j < nkv; ++j) {
          int nst = st + 1;
          uint32_t nph = ph;
          if (nst == ATT_FWD_KV_STAGES) { nst = 0; nph ^= 1u; }
          const bool more = (j + 1 < nkv);
          const bool nxt = more || has_next;   // the next S of each slot: next key tile, or tile 0 of the next item
          if (more && nkv > 2) {
            // inside an item of a long sequence the NEXT tile's S goes out before this tile's P.V: S only needs the score buffer back (released
            // when the slot's pass 2 has read it), and issuing the 8 P.V MMAs first kept ~700 cycles of single-thread issue
            // time on the softmax warps' critical path S(t+1) <- P(t)
            mbar_wait(bar_kvfull(nst), nph, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            ++t; issue_s(0, nst); --t;
            issue_pv(0, st, t);
            ++t; issue_s(1, nst); --t;
            issue_pv(1, st, t);
          } else {   // short heads (N <= 256); item boundaries wait for the next Q here
            // S of the next tile first (its score buffer is released early, see pass 2), then this tile's P.V: 0.260 -> 0.256 ms
            if (more) mbar_wait(bar_kvfull(nst), nph, abort_flag, p.status, ERR_TIMEOUT_ATTN);
            if (!more && has_next) {
              mbar_wait(bar_qfull, (it + 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              FWD_TRACE(1, ntr, 15);
              mbar_wait(bar_kvfull(nst), nph, abort_flag, p.status, ERR_TIMEOUT_ATTN);
              FWD_TRACE(1, ntr, 16);
            }
            if (nxt) { ++t; issue_s(0, nst); --t; }
            issue_pv(0, st, t);
            if (nxt) { ++t; issue_s(1, nst); --t; }
            issue_pv(1, st, t);
          }
          if (more && j + 2 == nkv) commit(bar_qempty);          // this item's last S MMAs are out
          if (!more && has_next && nkv == 1) commit(bar_qempty);  // single-tile items: the next item's only S MMAs are out
          commit(bar_kvempty(st));  // K_j / V_j no longer needed once everything issued so far retires
          ++t;
          st = nst;
          ph = nph;
        }
      }
    }
  } else {
    // ================================================================= softmax warps
    const int sl = warp >> 3;                       // slot
    const int quad = warp & 3, part = (warp >> 2) & 1;
    const int row = quad * 32 + lane;               // query row inside the slot's tile (= TMEM lane)
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem + sl * 128, tO = tmem + 256 + sl * 64;
    uint8_t* myP = genP + sl * 2 * ATT_TILE_BYTES;
    float* x_max = s_xchg + sl * 512;               // [part][128]
    float* x_sum = x_max + 256;
    const int bar_id = 1 + sl * 4 + quad;           // the two warps (parts) of this slot's lane quadrant
    const int cbase = part * 64;
    uint32_t t = 0;
    uint32_t ntr = 0;
    (void)ntr;
    const bool tracer = (quad == 0 && part == 0 && lane == 0);
    (void)tracer;
    // Item coordinates (qp, h, b, g) are advanced incrementally — item += gridDim.x is a mixed-radix add with carries — because
    // div / mod by run-time values cost ~150 dependent instructions per item on the softmax warps' critical path (the launcher
    // rejects problems with >= 2^31 items, so everything is 32-bit).
    int qp, h, b, g;
    int d_qp, d_h, d_b, d_g;   // gridDim.x in the same radix
    {
      unsigned v = blockIdx.x;
      qp = (int)(v % (unsigned)nqp); v /= (unsigned)nqp;
      h = (int)(v % (unsigned)p.H); v /= (unsigned)p.H;
      b = (int)(v % (unsigned)p.B); g = (int)(v / (unsigned)p.B);
      v = gridDim.x;
      d_qp = (int)(v % (unsigned)nqp); v /= (unsigned)nqp;
      d_h = (int)(v % (unsigned)p.H); v /= (unsigned)p.H;
      d_b = (int)(v % (unsigned)p.B); d_g = (int)(v / (unsigned)p.B);
    }
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const int row_base = b * p.N;
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha = 0.f;
      for (int j = 0; j < nkv; ++j, ++t) {
        if (tracer) FWD_TRACE(2 + sl, ntr, 20);
        mbar_wait(bar_sfull(sl), t & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        if (tracer) FWD_TRACE(2 + sl, ntr, 21);
        tc_fence_after();
        const int kv_valid = min(ATT_TILE, p.N - j * ATT_TILE);
        // pass 1: row maximum of this warp's 64 columns
        float m_part = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // a 32-column chunk entirely beyond the last key (ragged last tile, e.g. N = 197: 59 of its 128 columns) holds no
          // scores: skip its TMEM read and, below, its exponentials (warp-uniform; only the straddling chunk is masked)
          if (cbase + c * 32 >= kv_valid) continue;
          uint32_t r[32];
          tmem_ld32(tS + t_lane + cbase + c * 32, r);
          tmem_ld_wait();
          if (cbase + c * 32 + 32 <= kv_valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) m_part = fmaxf(m_part, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cbase + c * 32 + i < kv_valid) m_part = fmaxf(m_part, __uint_as_float(r[i]));
          }
        }
        x_max[part * ATT_TILE + row] = m_part;
        if (tracer) FWD_TRACE(2 + sl, ntr, 22);
        named_bar_sync(bar_id, 64);
        if (tracer) FWD_TRACE(2 + sl, ntr, 23);
        const float m_tile = fmaxf(m_part, x_max[(part ^ 1) * ATT_TILE + row]);
        const float m_new = fmaxf(m_run, m_tile * p.scale_log2);
        // fold the previous tile's O into the accumulators (its P.V finished long ago; this also
        // guarantees that the MMA no longer reads the P buffer we are about to overwrite)
        if (j > 0) {
          mbar_wait(bar_ofull(sl), (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          if (tracer) FWD_TRACE(2 + sl, ntr, 24);
          tc_fence_after();
          uint32_t r[32];
          tmem_ld32(tO + t_lane + part * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = o[i] * alpha + __uint_as_float(r[i]);
        }
        const float alpha_new = ex2_approx(m_run - m_new);  // m_run = -inf on the first tile -> 0
        // pass 2: probabilities -> bf16 -> swizzled smem. S[sl] is handed back to the MMA warp as soon as this warp's LAST
        // read of it has landed in registers (before the exponentials of that chunk), so the next S of this slot is computed
        // under the rest of pass 2 instead of after it (the clock64 timeline showed the slot waiting ~1.1 k cycles for it).
        const int last_read = (cbase + 32 < kv_valid) ? 1 : ((cbase < kv_valid) ? 0 : -1);
        if (last_read < 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sfree(sl));
        }
        float l_part = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float pv[32];
          if (cbase + c * 32 >= kv_valid) {   // no keys here: P = 0 (the P.V MMA still reads these columns)
#pragma unroll
            for (int i = 0; i < 32; ++i) pv[i] = 0.f;
            store_row_chunk_sw128(myP, row, cbase + c * 32, pv);
            continue;
          }
          uint32_t r[32];
          tmem_ld32(tS + t_lane + cbase + c * 32, r);
          tmem_ld_wait();
          if (c == last_read) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sfree(sl));   // S[sl] fully read by this warp
          }
          if (cbase + c * 32 + 32 <= kv_valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              pv[i] = ex2_approx(__uint_as_float(r[i]) * p.scale_log2 - m_new);
              l_part += pv[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = ex2_approx(__uint_as_float(r[i]) * p.scale_log2 - m_new);
              pv[i] = (cbase + c * 32 + i < kv_valid) ? e : 0.f;
              l_part += pv[i];
            }
          }
          store_row_chunk_sw128(myP, row, cbase + c * 32, pv);
        }
        if (tracer) FWD_TRACE(2 + sl, ntr, 25);
        x_sum[part * ATT_TILE + row] = l_part;
        fence_proxy_async_smem();                     // P stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pfull(sl));
        named_bar_sync(bar_id, 64);
        l_run = l_run * alpha_new + (l_part + x_sum[(part ^ 1) * ATT_TILE + row]);
        m_run = m_new;
        alpha = alpha_new;
      }
      // last tile's O
      if (tracer) FWD_TRACE(2 + sl, ntr, 26);
      mbar_wait(bar_ofull(sl), (t - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      if (tracer) FWD_TRACE(2 + sl, ntr, 27);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tO + t_lane + part * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = o[i] * alpha + __uint_as_float(r[i]);
      }
      if (tracer) FWD_TRACE(2 + sl, ntr, 28);
      tc_fence_before();
      // Output rows go through the slot's P buffer (free: the last P.V has retired) so that the global stores are whole
      // 128-byte rows (8 lanes x 16 B) instead of 32 scattered 16-byte pieces per warp instruction, which kept the LSU busy
      // for ~2000 cycles per item. Staging layout: [128 rows][8 x 16 B], 16-byte chunk index XOR (row & 7).
      {
        const float inv_l = __fdividef(1.0f, l_run);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(o[i] * inv_l, o[i + 1] * inv_l);
          w.y = pack_bf16(o[i + 2] * inv_l, o[i + 3] * inv_l);
          w.z = pack_bf16(o[i + 4] * inv_l, o[i + 5] * inv_l);
          w.w = pack_bf16(o[i + 6] * inv_l, o[i + 7] * inv_l);
          *reinterpret_cast<uint4*>(myP + row * 128 + (((part * 4 + (i >> 3)) ^ (row & 7)) << 4)) = w;
        }
        const int q = qp * 2 * ATT_TILE + sl * ATT_TILE + row;
        if (part == 0 && q < p.N)
          p.lse[(((long long)g * p.B + b) * p.H + h) * p.N + q] = (m_run + __log2f(l_run)) * 0.6931471805599453f;
        if (tracer) FWD_TRACE(2 + sl, ntr, 29);
        named_bar_sync(bar_id, 64);   // both column halves of this quadrant's 32 rows are staged
        if (tracer) FWD_TRACE(2 + sl, ntr, 30);
        // the two warps of the quadrant store 16 rows each; the next use of this P region (pass 2 of the next item) comes
        // after the max-exchange barrier of the same two warps, so the staged rows are not overwritten while being read
        const int c16 = lane & 7;
        const int r0 = quad * 32 + part * 16 + (lane >> 3), qr0 = qp * 2 * ATT_TILE + sl * ATT_TILE + r0;
        bf16* orow = p.out + (long long)g * p.out_gs + (long long)(row_base + qr0) * p.C + h * ATT_D + c16 * 8;
        const long long ostep = 4ll * p.C;
#pragma unroll
        for (int it2 = 0; it2 < 4; ++it2) {
          const int r = r0 + it2 * 4;
          if (qr0 + it2 * 4 < p.N) {
            const uint4 v = *reinterpret_cast<const uint4*>(myP + r * 128 + ((c16 ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(orow + it2 * ostep) = v;
          }
        }
        if (tracer) FWD_TRACE(2 + sl, ntr, 31);
      }
      qp += d_qp; h += d_h; b += d_b; g += d_g;
      if (qp >= nqp) { qp -= nqp; ++h; }
      if (h >= p.H) { h -= p.H; ++b; }
      if (b >= p.B) { b -= p.B; ++g; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------ backward
// delta[g][b][h][q] = sum_d dO[row][h*64+d] * O[row][h*64+d]; one warp per (row, head).
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                  long long rows_total /*G*B*N*/, int N, int H, int C) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w = wid; w < rows_total * H; w += nw) {
    const long long row = w / H;
    const int h = (int)(w % H);
    const uint32_t a = *reinterpret_cast<const uint32_t*>(o + row * C + h * ATT_D + lane * 2);
    const uint32_t d = *reinterpret_cast<const uint32_t*>(dout + row * C + h * ATT_D + lane * 2);
    const float2 af = unpack_bf16(a), df = unpack_bf16(d);
    const float s = warp_sum(af.x * df.x + af.y * df.y);
    if (lane == 0) {
      const long long gb = row / N;
      const int q = (int)(row % N);
      delta[(gb * H + h) * N + q] = s;
    }
  }
}

struct AttnBwdParams {
  const float* lse;
  const float* delta;
  bf16* dqkv;
  float* dq_acc;
  int N, H, C, B;
  long long qkv_gs, acc_gs;
  float scale, scale_log2;
  int* status;
};

// Backward, pipelined. 17 warps: warps 0..15 compute (quad = TMEM lane quadrant, part = 32-column slice of the
// query tile), warp 16 is the control warp whose lane 0 issues every TMA load and tcgen05.mma.
// Per query tile i the tensor pipe runs  S^T,dP^T(i+1)  right after the compute warps have pulled
// S^T,dP^T(i) into registers, and  dV,dK,dQ(i)  once P^T,dS^T(i) are in shared memory, so the
// elementwise work of tile i+1 overlaps the five MMAs of tile i. dQ partial tiles are transposed in
// shared memory and accumulated with 16-byte vector reductions (red.global.add.v4.f32).
// smem: K | V | Q0 | Q1 | dO0 | dO1 | PT(2 chunks) | dST(2 chunks) | dQ stages 16 x 2 KB | lse/delta [3][128] x2 | barriers
constexpr int ATT_BWD_COMPUTE_WARPS = 16;
constexpr int ATT_BWD_THREADS = (ATT_BWD_COMPUTE_WARPS + 3) * 32;   // + control warp + two more MMA issuers (dV, dK)
constexpr int ATT_BWD_SMEM = 10 * ATT_TILE_BYTES + 16 * 2048 + 3072 + 1024 + 128;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base, sV = base + ATT_TILE_BYTES, sQ = base + 2 * ATT_TILE_BYTES, sDO = base + 4 * ATT_TILE_BYTES;
  const uint32_t sPT = base + 6 * ATT_TILE_BYTES, sDST = base + 8 * ATT_TILE_BYTES;
  uint8_t* genPT = gen + 6 * ATT_TILE_BYTES;
  uint8_t* genDST = gen + 8 * ATT_TILE_BYTES;
  float* dq_stage = reinterpret_cast<float*>(gen + 10 * ATT_TILE_BYTES);           // [16 warps][32 rows][16 cols]
  // LSE / delta of a query tile are staged one tile ahead into a ring of 3 buffers: buffer (i+1)%3 is written at
  // the top of iteration i (before this warp's bar_sfree arrival, which orders it before S^T(i+1) becomes
  // visible) and its previous contents (tile i-2) were last read before bar_p(i-2), which every warp has passed.
  float* s_lse = reinterpret_cast<float*>(gen + 10 * ATT_TILE_BYTES + 16 * 2048);  // [3][128]
  float* s_delta = s_lse + 3 * ATT_TILE;                                           // [3][128]
  const uint32_t bar0 = base + 10 * ATT_TILE_BYTES + 16 * 2048 + 3072;
  const uint32_t bar_kv = bar0, bar_q0 = bar0 + 8, bar_s = bar0 + 24, bar_sfree = bar0 + 32, bar_p = bar0 + 40,
                 bar_d = bar0 + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 10 * ATT_TILE_BYTES + 16 * 2048 + 3072 + 64);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = (warp >> 2) & 3;
  const int trow = quad * 32 + lane;  // row inside the tile (= TMEM lane)
  const int g = blockIdx.z, bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int kv0 = blockIdx.x * ATT_TILE;
  const int row_base = b * p.N;
  const int nq = (p.N + ATT_TILE - 1) / ATT_TILE;
  const long long lse_base = (((long long)g * p.B + b) * p.H + h) * p.N;

  if (tid == 0) {
    *abort_flag = 0;
    mbar_init(bar_kv, 1);
    mbar_init(bar_q0, 1);
    mbar_init(bar_q0 + 8, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_sfree, ATT_BWD_COMPUTE_WARPS);
    mbar_init(bar_p, ATT_BWD_COMPUTE_WARPS);
    mbar_init(bar_d, 3);   // three issuers commit per tile: control warp (dQ), warp 17 (dV), warp 18 (dK)
    fence_barrier_init();
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (tid < ATT_TILE) {  // LSE (log2 domain) / delta of query tile 0
    s_lse[tid] = (tid < p.N) ? p.lse[lse_base + tid] * 1.4426950408889634f : 0.f;
    s_delta[tid] = (tid < p.N) ? p.delta[lse_base + tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;
  const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;

  if (warp == ATT_BWD_COMPUTE_WARPS) {
    // ================================================================= control warp
    // (runs converged: every lane executes the waits, one elected lane issues the TMA loads and the MMAs back to back)
    {
      const uint32_t idesc_kk = umma_idesc_bf16(128, 0, 0);   // S^T, dP^T : both operands K-major, N = 128
      const uint32_t idesc_mnmn = umma_idesc_bf16(64, 1, 1);  // dQ        : A = dS^T viewed MN-major, B = K MN-major
      auto load_q = [&](int i) {
        if (elect_one()) {
          const uint32_t bar = bar_q0 + 8 * (i & 1);
          mbar_arrive_expect_tx(bar, 2 * ATT_TILE_BYTES);
          tma_load_3d(&tmQKV, bar, sQ + (i & 1) * ATT_TILE_BYTES, h * ATT_D, row_base + i * ATT_TILE, g);
          tma_load_3d(&tmDO, bar, sDO + (i & 1) * ATT_TILE_BYTES, h * ATT_D, row_base + i * ATT_TILE, g);
        }
        __syncwarp();
      };
      // One thread issues 32 MMAs per query tile, 24 of them only 32 tensor-pipe cycles long (N = 64): the issue path must
      // not rebuild descriptors. All of them are constant adds on these bases (16-byte units of the start-address field).
      const uint64_t dK_k = umma_desc_sw128(sK, 16, 1024), dV_k = umma_desc_sw128(sV, 16, 1024);
      // (second Q / dO buffer = + one tile; selected arithmetically: indexing a local array would go through local memory)
      const uint64_t dQ_k0 = umma_desc_sw128(sQ, 16, 1024), dDO_k0 = umma_desc_sw128(sDO, 16, 1024);
      const uint64_t dDST_mn = umma_desc_sw128(sDST, ATT_TILE_BYTES, 1024), dK_mn = umma_desc_sw128(sK, ATT_TILE_BYTES, 1024);
      auto issue_s = [&](int i) {  // S^T = K Q_i^T ; dP^T = V dO_i^T
        const uint32_t boff = static_cast<uint32_t>(i & 1) * (ATT_TILE_BYTES >> 4);
        const uint64_t qd = dQ_k0 + boff, dd = dDO_k0 + boff;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tST, dK_k + 2 * k, qd + 2 * k, idesc_kk, k != 0);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tDPT, dV_k + 2 * k, dd + 2 * k, idesc_kk, k != 0);
          umma_commit(bar_s);
        }
        __syncwarp();
      };
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kv, 2 * ATT_TILE_BYTES);
        tma_load_3d(&tmQKV, bar_kv, sK, p.C + h * ATT_D, row_base + kv0, g);
        tma_load_3d(&tmQKV, bar_kv, sV, 2 * p.C + h * ATT_D, row_base + kv0, g);
      }
      __syncwarp();
      load_q(0);
      if (nq > 1) load_q(1);
      mbar_wait(bar_kv, 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      mbar_wait(bar_q0, 0, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      tc_fence_after();
      issue_s(0);
      for (int i = 0; i < nq; ++i) {
        const int buf = i & 1;
        // S^T,dP^T(i) are in registers everywhere -> the tensor pipe may start on tile i+1
        mbar_wait(bar_sfree, i & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        if (i + 1 < nq) {
          mbar_wait(bar_q0 + 8 * ((i + 1) & 1), ((i + 1) >> 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          tc_fence_after();
          issue_s(i + 1);
        }
        // P^T,dS^T(i) are in shared memory (and dQ(i-1) has been drained) -> dV, dK, dQ of tile i
        mbar_wait(bar_p, i & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_TILE / 16; ++k)   // dQ[q][d] = dS[q][kv] K[kv][d]  (A = dS^T viewed MN-major)
            umma_bf16_ss(tDQ, dDST_mn + k * 128, dK_mn + k * 128, idesc_mnmn, k != 0);
          umma_commit(bar_d);   // (dV / dK of this tile are issued and committed by the two other issuer warps)
        }
        __syncwarp();
        if (i + 2 < nq) {  // refill this Q/dO buffer once the MMAs that read it have retired
          mbar_wait(bar_d, i & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
          load_q(i + 2);
        }
      }
    }
  } else if (warp > ATT_BWD_COMPUTE_WARPS) {
    // ================================================================= MMA issuers for dV (warp 17) and dK (warp 18)
    // A single thread needs ~75 cycles per tcgen05.mma (descriptor arithmetic, R2UR, election), i.e. ~2400 cycles for the
    // 32 MMAs of a query tile whose tensor-pipe time is 1280 cycles: the issue stream is split over three warps
    // (control: S^T, dP^T of tile i+1 and dQ of tile i; these two: dV and dK, which accumulate in their own TMEM columns).
    {
      const bool is_dk = (warp == ATT_BWD_COMPUTE_WARPS + 2);
      const uint32_t idesc_kmn = umma_idesc_bf16(64, 0, 1);
      const uint64_t a0 = umma_desc_sw128(is_dk ? sDST : sPT, 16, 1024);                   // dS^T (dK) or P^T (dV), K-major
      const uint64_t b0 = umma_desc_sw128(is_dk ? sQ : sDO, ATT_TILE_BYTES, 1024);          // Q (dK) or dO (dV), MN-major
      const uint32_t acc = is_dk ? tDK : tDV;
      for (int i = 0; i < nq; ++i) {
        mbar_wait(bar_p, i & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        const uint64_t b = b0 + static_cast<uint32_t>(i & 1) * (ATT_TILE_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_TILE / 16; ++k)   // dV[kv][d] += P^T[kv][q] dO[q][d]   /   dK[kv][d] += dS^T[kv][q] Q[q][d]
            umma_bf16_ss(acc, a0 + ((k >> 2) * (ATT_TILE_BYTES >> 4) + (k & 3) * 2), b + k * 128, idesc_kmn, (i | k) != 0);
          umma_commit(bar_d);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= compute warps
    const bool kv_ok = (kv0 + trow) < p.N;
    float* my_stage = dq_stage + warp * 512;
    // drains dQ of query tile `qi` from TMEM: transpose through smem, 16-byte vector reductions
    auto drain_dq = [&](int qi) {
      uint32_t r[16];
      tmem_ld16(tDQ + t_lane + part * 16, r);
      tmem_ld_wait();
      float4* srow = reinterpret_cast<float4*>(my_stage + lane * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        srow[j ^ ((lane >> 1) & 3)] = make_float4(__uint_as_float(r[4 * j]) * p.scale, __uint_as_float(r[4 * j + 1]) * p.scale,
                                                  __uint_as_float(r[4 * j + 2]) * p.scale, __uint_as_float(r[4 * j + 3]) * p.scale);
      __syncwarp();
      const int j = lane & 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + (lane >> 2);
        const float4 v = reinterpret_cast<const float4*>(my_stage + rr * 16)[j ^ ((rr >> 1) & 3)];
        const int q = qi * ATT_TILE + quad * 32 + rr;
#ifdef CAVIT_BWD_NO_DQ   // timing experiment only: how much of the kernel is the fp32 dQ reduction traffic?
        if (q < 0)
#else
        if (q < p.N)
#endif
          red_add_v4(p.dq_acc + (long long)g * p.acc_gs + (long long)(row_base + q) * p.C + h * ATT_D + part * 16 + j * 4, v.x, v.y,
                     v.z, v.w);
      }
      __syncwarp();
    };

    for (int i = 0; i < nq; ++i) {
      const int q0 = i * ATT_TILE;
      const float* lse_i = s_lse + (i % 3) * ATT_TILE;
      const float* delta_i = s_delta + (i % 3) * ATT_TILE;
      if (i + 1 < nq && tid < ATT_TILE) {  // stage LSE / delta of the next query tile
        const int q = q0 + ATT_TILE + tid;
        s_lse[((i + 1) % 3) * ATT_TILE + tid] = (q < p.N) ? p.lse[lse_base + q] * 1.4426950408889634f : 0.f;
        s_delta[((i + 1) % 3) * ATT_TILE + tid] = (q < p.N) ? p.delta[lse_base + q] : 0.f;
      }
      mbar_wait(bar_s, i & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
      tc_fence_after();
      const int cb = part * 32;
      uint32_t rs[32], rp[32];
      tmem_ld32(tST + t_lane + cb, rs);
      tmem_ld32(tDPT + t_lane + cb, rp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree);
      const int q_valid = min(ATT_TILE, p.N - q0);
      float pt[32], dst[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = cb + j;
        const bool ok = kv_ok && (col < q_valid);
        const float pr = ok ? ex2_approx(__uint_as_float(rs[j]) * p.scale_log2 - lse_i[col]) : 0.f;
        pt[j] = pr;
        dst[j] = pr * (__uint_as_float(rp[j]) - delta_i[col]);
      }
      if (i > 0) {  // MMAs of tile i-1 done: P^T/dS^T buffers are free and dQ(i-1) is complete
        mbar_wait(bar_d, (i - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
        tc_fence_after();
        drain_dq(i - 1);
      }
      store_row_chunk_sw128(genPT, trow, cb, pt);
      store_row_chunk_sw128(genDST, trow, cb, dst);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    mbar_wait(bar_d, (nq - 1) & 1, abort_flag, p.status, ERR_TIMEOUT_ATTN);
    tc_fence_after();
    drain_dq(nq - 1);
    // write dK (scaled) and dV for this key/value tile: each warp stores a 16-column slice of both
    {
      const int kv = kv0 + trow;
      bf16* drow = p.dqkv + (long long)g * p.qkv_gs + (long long)(row_base + kv) * (3 * p.C) + h * ATT_D + part * 16;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t t = which == 0 ? tDK : tDV;
        const float sc = which == 0 ? p.scale : 1.0f;
        bf16* dstp = drow + (which == 0 ? p.C : 2 * p.C);
        uint32_t r[16];
        tmem_ld16(t + t_lane + part * 16, r);
        tmem_ld_wait();
        if (kv < p.N) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(r[j]) * sc, __uint_as_float(r[j + 1]) * sc);
            w.y = pack_bf16(__uint_as_float(r[j + 2]) * sc, __uint_as_float(r[j + 3]) * sc);
            w.z = pack_bf16(__uint_as_float(r[j + 4]) * sc, __uint_as_float(r[j + 5]) * sc);
            w.w = pack_bf16(__uint_as_float(r[j + 6]) * sc, __uint_as_float(r[j + 7]) * sc);
            *reinterpret_cast<uint4*>(dstp + j) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// dqkv[row][0:C] = bf16(dq_acc[row][0:C])
__global__ void attn_dq_store_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv, long long rows, int C) {
  const int C4 = C >> 2;
  const long long total = rows * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C4;
    const int c4 = (int)(i % C4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(acc) + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *(reinterpret_cast<uint2*>(dqkv + row * 3 * C) + c4) = o;
  }
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_attn_fwd(const void* qkv, void* out, float* lse, int32_t G, int32_t B, int32_t N, int32_t H, float scale,
                   void* stream) {
  if (!qkv || !out || !lse) return fail(CAVIT_E_BADARG, "cavit_attn_fwd: null pointer");
  if (G <= 0 || B <= 0 || N <= 0 || H <= 0) return fail(CAVIT_E_BADARG, "cavit_attn_fwd: bad extents");
  const int C = H * ATT_D;
  const long long T = (long long)B * N;
  const CUtensorMap* tm = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, ATT_TILE);
  if (!tm) return CAVIT_E_BADARG;
  static PerDeviceFlag attr;
  if (attr.unset()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "attn fwd smem attribute: %s", cudaGetErrorString(e));
    attr.set();
  }
  AttnFwdParams p;
  p.out = reinterpret_cast<bf16*>(out);
  p.lse = lse;
  p.N = N; p.H = H; p.C = C; p.B = B; p.G = G;
  p.out_gs = T * C;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "no status word");
  const long long items = (long long)G * B * H * ((N + 2 * ATT_TILE - 1) / (2 * ATT_TILE));
  if (items >= (1ll << 31)) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_attn_fwd: too many work items");
  const int grid = (int)(items < sm_count() ? items : sm_count());
  attn_fwd_kernel<<<grid, ATT_FWD_THREADS, ATT_FWD_SMEM, as_stream(stream)>>>(*tm, p);
#ifdef CAVIT_FWD_TRACE
  {
    static int calls = 0;
    if (++calls == 3) {
      cudaDeviceSynchronize();
      static unsigned long long h[4][2048];
      cudaMemcpyFromSymbol(h, g_fwd_trace, sizeof(h));
      for (int role = 0; role < 4; ++role)
        for (int i = 0; i < 400; ++i)
          if (h[role][2 * i]) fprintf(stderr, "FWDTRACE %d %d %llu %llu\n", role, i, h[role][2 * i] - h[1][0], h[role][2 * i + 1]);
    }
  }
#endif
  count_launch();
  return check_launch("cavit_attn_fwd");
}

int cavit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                   float* dq_acc, int32_t G, int32_t B, int32_t N, int32_t H, float scale, void* stream) {
  if (!qkv || !out || !dout || !lse || !dqkv || !delta || !dq_acc) return fail(CAVIT_E_BADARG, "cavit_attn_bwd: null pointer");
  if (G <= 0 || B <= 0 || N <= 0 || H <= 0) return fail(CAVIT_E_BADARG, "cavit_attn_bwd: bad extents");
  const int C = H * ATT_D;
  const long long T = (long long)B * N;
  cudaStream_t st = as_stream(stream);
  // N <= 256 (all 2-D slice configs): whole-head persistent kernel, no delta / dQ helper kernels (attn_short.cu).
  // CAVIT_ATTN_GENERIC=1 forces the generic tile kernel below (A/B measurements, tests of both paths).
  static const bool force_generic = [] { const char* e = getenv("CAVIT_ATTN_GENERIC"); return e && e[0] == '1'; }();
  if (N <= 256 && !force_generic) return launch_attn_bwd_short(qkv, out, dout, lse, dqkv, G, B, N, H, scale, st);
  const CUtensorMap* tq = tensor_map_bf16_3d(qkv, 3 * C, T, G, 3 * C, T * 3 * C, 64, ATT_TILE);
  const CUtensorMap* td = tensor_map_bf16_3d(dout, C, T, G, C, T * C, 64, ATT_TILE);
  if (!tq || !td) return CAVIT_E_BADARG;
  static PerDeviceFlag attr;
  if (attr.unset()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "attn bwd smem attribute: %s", cudaGetErrorString(e));
    attr.set();
  }
  {
    const long long warps = (long long)G * T * H;
    long long blocks = (warps + 7) / 8;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    attn_delta_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(out), reinterpret_cast<const bf16*>(dout),
                                                   delta, (long long)G * T, N, H, C);
    count_launch();
  }
  cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)G * T * C, st);
  AttnBwdParams p;
  p.lse = lse; p.delta = delta;
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.dq_acc = dq_acc;
  p.N = N; p.H = H; p.C = C; p.B = B;
  p.qkv_gs = T * 3 * C;
  p.acc_gs = T * C;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "no status word");
  dim3 grid((N + ATT_TILE - 1) / ATT_TILE, B * H, G);
  attn_bwd_kernel<<<grid, ATT_BWD_THREADS, ATT_BWD_SMEM, st>>>(*tq, *td, p);
  count_launch();
  int rc = check_launch("cavit_attn_bwd");
  if (rc) return rc;
  {
    const long long total = (long long)G * T * C / 4;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    attn_dq_store_kernel<<<(int)blocks, 256, 0, st>>>(dq_acc, reinterpret_cast<bf16*>(dqkv), (long long)G * T, C);
    count_launch();
  }
  return check_launch("cavit_attn_bwd(dq)");
}

}  // extern "C"
