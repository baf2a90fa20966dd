// cavit-sm100 — K-EMBED: patch extraction fused with the embedding GEMM and the positional add; fused weight gradient.
//
//   tokens[stream][b*N + 1 + t][:] = unfold(img)[b, m, t, :] . W^T + bias + pos[1 + t]
//   t = (hi*Wn + wi)*Dn + di (d fastest),  feature f = (a*hp + b)*wp + c          (bit-exact index map)
//
// Replaces einops.rearrange + patch_to_embedding + cat(cls) + pos add of /root/reference/model_cross.py:189-198 and
// modelv3.py:125-140, and the weight gradient their autograd computes. The unfolded [B, Np, P] tensor never exists and
// nothing is converted on the way: TMA fetches BRICKS of the fp32 volume straight into the 128-byte-swizzled operand stage
// and the forward's tcgen05.mma.kind::tf32 consumes the fp32 words as they are (TF32 keeps 10 mantissa bits, three more
// than the bf16 operand copy the separate patchify kernel makes), against the fp32 master weights. The weight gradient
// (second half of this file) needs MN-major operands and goes through a bf16 conversion stage.
//
// Brick = for one patch row hi and depth index di the Wn patches of nv consecutive volumes (v = sample*M + modality) =
// nv*Wn token SLOTS. One 5-D TMA box {c: cw, wi: Wn, h: 1, z: 1, v: nv} of the volume tensor map fetches one PIECE of a
// brick: cw = min(wp, 32) consecutive features — row (a, b) of the reference's feature order (p1 p2 p3), or a 32-wide part
// of it — of every slot. It lands as [slot][cw fp32], one 4*cw-byte row per token, swizzled with the TMA mode whose span
// equals that row (SWIZZLE_32B / 64B / 128B for wp = 8 / 16 / >= 32): exactly the canonical K-major operand tile of that
// swizzle mode, and 32 / cw pieces make the 32-feature k-block. (One box with two or four rows h packed into a 128-byte
// swizzled row does not work: with an inner box extent below the swizzle span TMA does not pack rows densely — measured.)
// Volumes past the batch are zero-filled by TMA.
//   forward  D[slot][n] += A[slot][k32] W[n][k32]^T    work item = (brick, 128/192/256 channels), K loop over P/32 k-blocks
//   wgrad    dW[c][f]   += dY[slot][c]^T X[slot][f]    work item = (128 channels, 128 features, a slice of the bricks),
//                                                      K loop over bricks of <= 64 slots; fp32 red.add into dW
// 2-D slices stored as (D, H, 1) with (dp, hp, 1) patches (BASELINE.json configs[1]) have no contiguous run along W; they are
// addressed in the frame (1, D, H) x (1, dp, hp): same memory, same feature order, token index with the frame's strides.
//
// Forward warp roles: 0 TMA producer, 1 MMA issuer + TMEM owner, 2-9 epilogue (two warps per TMEM lane quadrant, alternating
// 32-column chunks). TMA ring of 4-6 stages, two TMEM accumulator stages.
#include "common.cuh"
#include "internal.h"

namespace cavit {

#ifndef CAVIT_EMBED_EPI_WARPS
#define CAVIT_EMBED_EPI_WARPS 8
#endif
constexpr int EM_BM = 128;                // token slots per forward tile
constexpr int EM_EPI_WARPS = CAVIT_EMBED_EPI_WARPS;   // 4 or 8: one or two warps per TMEM lane quadrant
constexpr int EM_THREADS = 64 + 32 * EM_EPI_WARPS;
static_assert(EM_EPI_WARPS == 4 || EM_EPI_WARPS == 8, "epilogue warps");
constexpr int EM_A_BYTES = EM_BM * 128;   // [128 slots][32 fp32]

// Geometry in the kernel's FRAME (fD, fH, fW) x (dp, hp, wp): the volume axes as the bricks see them, fW contiguous.
struct EmbedParams {
  int V, M, Bs, Dn, Hn, Wn, dp, hp, wp, P, C, Np, Ntok, sample_major;
  int fD, fH, fW;        // frame extents
  int tz, ty, tx;        // token index of frame patch (di, hi, wi) = di*tz + hi*ty + wi*tx
  int nv, vgroups, tiles_m, tiles_n, a_bytes;
  int cw, rows32;        // a piece = cw features (4*cw-byte rows); rows32 = 32 / cw pieces per forward k-block
  int layout;            // UMMA layout type of the piece tiles: 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B
  float* out;
  const float* bias;
  const float* pos;
  int* status;
};

// tcgen05.mma kind::tf32: fp32 words in shared memory read as TF32, fp32 accumulation in TMEM; K = 8 per instruction.
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (format 2 at bits 7-9 / 10-12), transpose bits 15 / 16
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n, int a_mn, int b_mn, int m = 128) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Volume-map coordinates {c0, h, z} of piece `q` (features q*cw .. q*cw + cw - 1) of the patches in patch row hi / depth
// index di.
__device__ __forceinline__ void piece_coords(const EmbedParams& p, int q, int hi, int di, int& c0, int& h, int& z) {
  int row;                       // (a, b) row of the feature axis, row = a*hp + b
  if (p.wp <= 32) { row = q; c0 = 0; }
  else { const int per = p.wp >> 5; row = q / per; c0 = (q - row * per) << 5; }
  const int a = row / p.hp, b = row - a * p.hp;
  h = hi * p.hp + b;
  z = di * p.dp + a;
}

template <int BN>
struct EmbedCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = EM_A_BYTES + B_BYTES;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 256 - EM_EPI_WARPS * 4096;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EM_EPI_WARPS * 4096;
  static_assert(STAGES >= 3 && 8 * (2 * STAGES + 4) + 8 <= 256, "barrier block");
};

// out row / positional row of token slot `slot` of brick (vg, hi, di); out_row < 0: the slot holds no token
__device__ __forceinline__ void embed_slot_rows(const EmbedParams& p, int vg, int hi, int di, int slot, long long& out_row,
                                                int& pos_row) {
  const int vl = slot / p.Wn, wi = slot - vl * p.Wn;
  const int v = vg * p.nv + vl;
  out_row = -1;
  pos_row = 0;
  if (vl >= p.nv || v >= p.V) return;
  const int t = hi * p.ty + wi * p.tx + di * p.tz;
  const int m = v % p.M, b = v / p.M;
  if (p.sample_major) {   // ModelVIT: streams concatenated on the token axis, one positional table over all of them
    out_row = (long long)b * p.Ntok + 1 + m * p.Np + t;
    pos_row = 1 + m * p.Np + t;
  } else {                // ModelCross: one token stream per modality, shared positional table
    out_row = ((long long)m * p.Bs + b) * p.Ntok + 1 + t;
    pos_row = 1 + t;
  }
}

template <int BN>
__global__ void __launch_bounds__(EM_THREADS, 1)
embed_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmVol, const __grid_constant__ CUtensorMap tmW,
                       const EmbedParams p) {
  using Cfg = EmbedCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* epi_stage = reinterpret_cast<float*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // slots past nv*Wn are never written by TMA; their accumulator rows are never stored, but keep them finite
  const int sub_bytes = EM_BM * p.cw * 4;             // one piece tile [128 slots][cw fp32]
  if (p.a_bytes < EM_A_BYTES) {
    const int land = p.a_bytes / p.rows32;
    for (int s = 0; s < STAGES; ++s)
      for (int r = 0; r < p.rows32; ++r) {
        uint4* st = reinterpret_cast<uint4*>(smem_gen + s * Cfg::STAGE_BYTES + r * sub_bytes + land);
        for (int i = threadIdx.x; i < (sub_bytes - land) / 16; i += EM_THREADS) st[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EM_EPI_WARPS);
    }
    fence_barrier_init();
    prefetch_tmap(&tmVol);
    prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = p.P >> 5;
  const int total = p.tiles_m * p.tiles_n;
  const int per_vg = p.Hn * p.Dn;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total; work += gridDim.x) {
        const int tm = work / p.tiles_n, tn = work - tm * p.tiles_n;
        const int vg = tm / per_vg, rem = tm - vg * per_vg, hi = rem / p.Dn, di = rem - hi * p.Dn;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_EMPTY);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES, sB = sA + EM_A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), p.a_bytes + Cfg::B_BYTES);
          for (int r = 0; r < p.rows32; ++r) {
            int c0, h, z;
            piece_coords(p, kb * p.rows32 + r, hi, di, c0, h, z);
            tma_load_5d(&tmVol, full_bar(stage), sA + r * sub_bytes, c0, 0, h, z, vg * p.nv);
          }
          tma_load_3d(&tmW, full_bar(stage), sB, kb << 5, tn * BN, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(BN, 0, 0, EM_BM);
    const uint64_t adesc0 = umma_desc_layout(smem_base, 16, 32 * p.cw, p.layout);     // 8 rows of 4*cw bytes per atom
    const uint64_t bdesc0 = umma_desc_sw128(smem_base + EM_A_BYTES, 16, 1024);
    const int ksub = p.cw >> 3;                                                    // MMAs (8 features) per piece tile
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
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
          for (int k = 0; k < 4; ++k) {   // 8 features = 32 bytes of a row per MMA
            const int sub = k / ksub, within = k - sub * ksub;
            umma_tf32_ss(d_tmem, ad + static_cast<uint32_t>((sub * sub_bytes) >> 4) + 2 * within, bd + 2 * k, idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else {
    // ---------------------------------------------------------------- epilogue: + bias + pos, scatter rows to tokens
    const int q = warp & 3;                                  // TMEM lane quadrant of this warp
    float* stage_t = epi_stage + (warp - 2) * 1024;
    const int chunk0 = (warp - 2) >> 2;                      // with 8 warps the two of a quadrant alternate column chunks
    const int c4 = lane & 7, rsub = lane >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      const int tm = work / p.tiles_n, tn = work - tm * p.tiles_n;
      const int vg = tm / per_vg, rem = tm - vg * per_vg, hi = rem / p.Dn, di = rem - hi * p.Dn;
      long long orow[8];
      int prow[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) embed_slot_rows(p, vg, hi, di, q * 32 + it * 4 + rsub, orow[it], prow[it]);
      mbar_wait(tfull_bar(as), aphase, abort_flag, p.status, ERR_TIMEOUT_TMEM_FULL);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EM_EPI_WARPS / 4) {
        const int col0 = tn * BN + c * 32;
        if (col0 >= p.C) break;                              // warp-uniform (C % 32 == 0)
        // positional rows and bias first: their loads are in flight while the accumulator chunk comes out of TMEM
        float4 pe[8];
#pragma unroll
        for (int it = 0; it < 8; ++it)
          pe[it] = (orow[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(p.pos + (long long)prow[it] * p.C + col0 + c4 * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c4 * 4));
        uint32_t acc[32];
        tmem_ld32(t_row + c * 32, acc);
        tmem_ld_wait();
        float4* srow = reinterpret_cast<float4*>(stage_t + lane * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          srow[j ^ (lane & 7)] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                             __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (orow[it] >= 0) {
            const int r = it * 4 + rsub;
            const float4 t = reinterpret_cast<const float4*>(stage_t + r * 32)[c4 ^ (r & 7)];
            float4 o;
            o.x = t.x + bias4.x + pe[it].x; o.y = t.y + bias4.y + pe[it].y;
            o.z = t.z + bias4.z + pe[it].z; o.w = t.w + bias4.w + pe[it].w;
            *reinterpret_cast<float4*>(p.out + orow[it] * p.C + col0 + c4 * 4) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
//   dW[c][f] = sum over patch tokens of dY[token][c] * unfold(img)[token][f]      (input volumes need no gradient)
// A GEMM whose reduction axis is the token axis: both operands MN-major (k = token slot), bf16 — tcgen05 reads MN-major
// TF32 operands only in the SWIZZLE_128B_BASE32B layout with 128-byte rows, which the 32- / 64-byte feature runs of
// wp = 8 / 16 patches cannot form, so this kernel stages the volume brick in fp32 and converts. Per k-block one BRICK of
// <= 64 token slots (fixed hi, di; the Wn patches of nbq samples x all M modalities):
//   A  dY rows straight from the bf16 token-gradient tensor by one 5-D TMA box per 64 channels
//      {c: 64, di: 1, hw: Wn, m: M, b: nbq} (the CLS rows are skipped by the base pointer) — already the swizzled tile;
//   B  the same volume brick as forward, 128 features wide: fp32 landing -> converter warps -> bf16 MN-major tile.
// Work item = (128 channels, 128 features, a slice of the bricks); partial sums meet in dW with fp32 red.add (dW zeroed
// by the caller side of the ABI). Slots of a brick that hold no token (Wn * M * nbq < 64) are zero rows of both operands.
constexpr int EW_STAGES = 3;
constexpr int EW_THREADS = 320;
constexpr int EW_BK = 64;            // token slots per k-block
constexpr int EW_BN = 128;
constexpr int EW_A_BYTES = 2 * EW_BK * 128;     // two 64-channel chunks of 64 k-rows
constexpr int EW_B_BYTES = 2 * EW_BK * 128;     // two 64-feature chunks of 64 k-rows
constexpr int EW_LAND_BYTES = EW_BK * EW_BN * 4;
constexpr int EW_STAGE_BYTES = EW_LAND_BYTES + EW_A_BYTES + EW_B_BYTES;
constexpr int EW_SMEM_BYTES = EW_STAGES * EW_STAGE_BYTES + 1024 + 256 + 4 * 4096;

struct EmbedWgradParams {
  int Wn, wp, nh, nz, nvw, slots, M, nbq, Dn, Hn, hp, dp, C, P;
  int bgroups, bricks, tiles_c, tiles_p, splits, bricks_per_split, land_bytes, a_bytes, slice_frame;
  float* dW;
  int* status;
};

__global__ void __launch_bounds__(EW_THREADS, 1)
embed_wgrad_kernel(const __grid_constant__ CUtensorMap tmVol, const __grid_constant__ CUtensorMap tmDY,
                   const EmbedWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + EW_STAGES * EW_STAGE_BYTES;
  auto land_bar = [&](int s) { return bar0 + 8u * s; };
  auto bfull_bar = [&](int s) { return bar0 + 8u * (EW_STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * EW_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * EW_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * EW_STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + EW_STAGES * EW_STAGE_BYTES + 8 * (3 * EW_STAGES + 4));
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* epi_stage = reinterpret_cast<float*>(smem_gen + EW_STAGES * EW_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // k-rows no TMA box / converter ever writes (slots .. 63) must be zero in both operand tiles
  for (int s = 0; s < EW_STAGES; ++s) {
    uint4* ab = reinterpret_cast<uint4*>(smem_gen + s * EW_STAGE_BYTES + EW_LAND_BYTES);
    for (int i = threadIdx.x; i < (EW_A_BYTES + EW_B_BYTES) / 16; i += EW_THREADS) {
      const int row = (i % (EW_BK * 8)) / 8;    // 8 chunks of 16 B per 128-byte k-row, 64 rows per 8 KB chunk block
      if (row >= p.slots) ab[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < EW_STAGES; ++s) {
      mbar_init(land_bar(s), 1);
      mbar_init(bfull_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tmVol);
    prefetch_tmap(&tmDY);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tile fastest: the CTAs running side by side work on the SAME slice of bricks for different (channel, feature) tiles, so
  // a brick comes from HBM once and from L2 for the other tiles
  const int tiles = p.tiles_c * p.tiles_p;
  const int total = tiles * p.splits;
  const int rows_per_tile = EW_BN / p.wp;      // (a, b) rows of the feature axis per 128-feature tile

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total; work += gridDim.x) {
        const int tile = work % tiles, split = work / tiles;
        const int tc = tile / p.tiles_p, tp = tile - tc * p.tiles_p;
        const int row0 = tp * rows_per_tile, a0 = row0 / p.hp, h0 = row0 - a0 * p.hp;
        const int k0 = split * p.bricks_per_split, k1 = min(p.bricks, k0 + p.bricks_per_split);
        for (int kbi = k0; kbi < k1; ++kbi) {
          const int bg = kbi % p.bgroups, r = kbi / p.bgroups, di = r % p.Dn, hi = r / p.Dn;
          mbar_wait(empty_bar(stage), phase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_EMPTY);
          const uint32_t sL = smem_base + stage * EW_STAGE_BYTES, sA = sL + EW_LAND_BYTES;
          mbar_arrive_expect_tx(land_bar(stage), p.land_bytes + 2 * p.a_bytes);
          tma_load_5d(&tmVol, land_bar(stage), sL, 0, 0, hi * p.hp + h0, di * p.dp + a0, bg * p.nvw);
          const int u = p.slice_frame ? hi : di, w0 = p.slice_frame ? 0 : hi * p.Wn;
          tma_load_5d(&tmDY, land_bar(stage), sA, tc * 128, u, w0, 0, bg * p.nbq);
          tma_load_5d(&tmDY, land_bar(stage), sA + EW_BK * 128, tc * 128 + 64, u, w0, 0, bg * p.nbq);
          if (++stage == EW_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(EW_BN, 1, 1, 128);
    const uint64_t adesc0 = umma_desc_sw128(smem_base + EW_LAND_BYTES, EW_BK * 128, 1024);
    const uint64_t bdesc0 = umma_desc_sw128(smem_base + EW_LAND_BYTES + EW_A_BYTES, EW_BK * 128, 1024);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      const int split = work / tiles;
      const int k0 = split * p.bricks_per_split, k1 = min(p.bricks, k0 + p.bricks_per_split);
      mbar_wait(tempty_bar(as), aphase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_TMEM_EMPTY);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * EW_BN;
      for (int kbi = k0; kbi < k1; ++kbi) {
        mbar_wait(land_bar(stage), phase, abort_flag, p.status, ERR_TIMEOUT_FULL);
        mbar_wait(bfull_bar(stage), phase, abort_flag, p.status, ERR_TIMEOUT_FULL);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint32_t>(stage * (EW_STAGE_BYTES >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint32_t>(stage * (EW_STAGE_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < EW_BK / 16; ++k)
            umma_bf16_ss(d_tmem, ad + k * 128, bd + k * 128, idesc, ((kbi - k0) | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kbi == k1 - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == EW_STAGES) { stage = 0; phase ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp < 6) {
    const int tid = threadIdx.x - 64;
    const int slot = tid >> 1, half = tid & 1;
    const int vl = slot / p.Wn, wi = slot - vl * p.Wn;
    const bool has = slot < p.slots;
    int stage = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      const int split = work / tiles;
      const int k0 = split * p.bricks_per_split, k1 = min(p.bricks, k0 + p.bricks_per_split);
      for (int kbi = k0; kbi < k1; ++kbi) {
        mbar_wait(land_bar(stage), phase, abort_flag, p.status, ERR_TIMEOUT_FULL);
        if (has) {
          const float* land = reinterpret_cast<const float*>(smem_gen + stage * EW_STAGE_BYTES);
          uint8_t* brow = smem_gen + stage * EW_STAGE_BYTES + EW_LAND_BYTES + EW_A_BYTES + half * (EW_BK * 128) + slot * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = half * 64 + 8 * j, row = f / p.wp, c0 = f - row * p.wp;
            const int zl = row / p.nh, hl = row - zl * p.nh;
            const float4* src = reinterpret_cast<const float4*>(land + ((((vl * p.nz + zl) * p.nh + hl) * p.Wn + wi) * p.wp + c0));
            const float4 x0 = src[0], x1 = src[1];
            uint4 q;
            q.x = pack_bf16(x0.x, x0.y); q.y = pack_bf16(x0.z, x0.w);
            q.z = pack_bf16(x1.x, x1.y); q.w = pack_bf16(x1.z, x1.w);
            *reinterpret_cast<uint4*>(brow + ((j ^ (slot & 7)) << 4)) = q;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bfull_bar(stage));
        if (++stage == EW_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    float* stage_t = epi_stage + (warp - 6) * 1024;
    const int c4 = lane & 7, rsub = lane >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      const int tile = work % tiles;
      const int tc = tile / p.tiles_p, tp = tile - tc * p.tiles_p;
      mbar_wait(tfull_bar(as), aphase, abort_flag, p.status, ERR_TIMEOUT_TMEM_FULL);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * EW_BN;
#pragma unroll 1
      for (int c = 0; c < EW_BN / 32; ++c) {
        const int col0 = tp * EW_BN + c * 32;
        if (col0 >= p.P) break;                 // P % 64 == 0: chunks are whole or absent
        uint32_t acc[32];
        tmem_ld32(t_row + c * 32, acc);
        tmem_ld_wait();
        float4* srow = reinterpret_cast<float4*>(stage_t + lane * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          srow[j ^ (lane & 7)] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                             __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rsub, crow = tc * 128 + q * 32 + r;
          if (crow < p.C) {
            const float4 t = reinterpret_cast<const float4*>(stage_t + r * 32)[c4 ^ (r & 7)];
            float* o = p.dW + (long long)crow * p.P + col0 + c4 * 4;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(o)), "f"(t.x) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(o + 1)), "f"(t.y) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(o + 2)), "f"(t.z) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(o + 3)), "f"(t.w) : "memory");
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// db[c] = sum_{n >= 1} dpos[n][c]: the embedding bias sees exactly the patch-token gradients, and
// dpos[n] = sum over streams / samples of the token gradient at position n is already there (cavit_embed_param_grads).
__global__ void embed_bias_grad_kernel(const float* __restrict__ dpos, float* __restrict__ db, int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int n = 1;
  for (; n + 4 <= N; n += 4) {
    a0 += dpos[(long long)n * C + c];
    a1 += dpos[(long long)(n + 1) * C + c];
    a2 += dpos[(long long)(n + 2) * C + c];
    a3 += dpos[(long long)(n + 3) * C + c];
  }
  for (; n < N; ++n) a0 += dpos[(long long)n * C + c];
  db[c] = (a0 + a1) + (a2 + a3);
}

// Geometry shared by forward and weight gradient; returns 0 when this shape is outside the fused kernels' reach.
static int embed_geometry(EmbedParams& p, int B, int M, int D, int H, int W, int dp, int hp, int wp, int C, int sample_major) {
  if (B <= 0 || M <= 0 || dp <= 0 || hp <= 0 || wp <= 0 || D % dp || H % hp || W % wp) return 0;
  const int Dn0 = D / dp;
  if (W == 1 && wp == 1) {        // (D, H, 1) slices: frame (1, D, H), token t = hi*Dn + di = wi'*Dn + hi'
    p.tz = 0; p.ty = 1; p.tx = Dn0;
    W = H; wp = hp; H = D; hp = dp; D = 1; dp = 1;
  } else {                        // t = (hi*Wn + wi)*Dn + di
    p.tz = 1; p.ty = (W / wp) * Dn0; p.tx = Dn0;
  }
  // a 32-feature k-block is a whole number of patch rows (wp = 8, 16, 32) or a 32-wide part of one (wp = 64, 96, ...)
  if (wp < 8 || (wp <= 32 && (32 % wp || hp % (32 / wp))) || (wp > 32 && wp % 32) || (W % 4) || (C % 32)) return 0;
  p.V = B * M; p.M = M; p.Bs = B;
  p.fD = D; p.fH = H; p.fW = W;
  p.Dn = D / dp; p.Hn = H / hp; p.Wn = W / wp;
  p.dp = dp; p.hp = hp; p.wp = wp;
  p.cw = wp <= 32 ? wp : 32;
  p.rows32 = 32 / p.cw;
  p.layout = p.cw == 32 ? 2 : (p.cw == 16 ? 4 : 6);
  p.P = dp * hp * wp; p.C = C;
  p.Np = p.Dn * p.Hn * p.Wn;
  p.Ntok = sample_major ? M * p.Np + 1 : p.Np + 1;
  p.sample_major = sample_major;
  if (p.Wn > EM_BM || p.P % 32) return 0;
  p.nv = EM_BM / p.Wn;
  if (p.nv > p.V) p.nv = p.V;
  if (p.nv > 256) p.nv = 256;
  p.vgroups = (p.V + p.nv - 1) / p.nv;
  p.tiles_m = p.vgroups * p.Hn * p.Dn;
  p.a_bytes = p.nv * p.Wn * 128;
  return 1;
}

// fp32 volume map, dims {c, wi, h, z, v}; a box {cw, Wn, 1, 1, nv} lands as [v][wi][c] = one 4*cw-byte row per token slot,
// swizzled with the mode whose span is that row.
static const CUtensorMap* volume_map(const float* img, const EmbedParams& p, int nv) {
  const int D = p.fD, H = p.fH, W = p.fW;
  const uint64_t dims[5] = {(uint64_t)p.wp, (uint64_t)p.Wn, (uint64_t)H, (uint64_t)D, (uint64_t)p.V};
  const uint64_t strides[4] = {(uint64_t)p.wp * 4, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)D * H * W * 4};
  const uint32_t box[5] = {(uint32_t)p.cw, (uint32_t)p.Wn, 1u, 1u, (uint32_t)nv};
  return tensor_map_nd(1, 5, img, dims, strides, box, p.cw * 4);
}

// fp32 brick map of the weight gradient: box {wp, Wn, rows b, planes a, nv}, no swizzle (landing buffer for the converters)
static const CUtensorMap* volume_brick_map(const float* img, const EmbedParams& p, int nb_rows, int nz, int nv) {
  const int D = p.fD, H = p.fH, W = p.fW;
  const uint64_t dims[5] = {(uint64_t)p.wp, (uint64_t)p.Wn, (uint64_t)H, (uint64_t)D, (uint64_t)p.V};
  const uint64_t strides[4] = {(uint64_t)p.wp * 4, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)D * H * W * 4};
  const uint32_t box[5] = {(uint32_t)p.wp, (uint32_t)p.Wn, (uint32_t)nb_rows, (uint32_t)nz, (uint32_t)nv};
  return tensor_map_nd(1, 5, img, dims, strides, box, 0);
}

template <int BN>
static int embed_fwd_launch(const CUtensorMap* tv, const CUtensorMap* tw, EmbedParams& p, cudaStream_t st) {
  using Cfg = EmbedCfg<BN>;
  static PerDeviceFlag attr;
  if (attr.unset()) {
    cudaError_t e = cudaFuncSetAttribute(embed_fused_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "embed fwd smem attribute: %s", cudaGetErrorString(e));
    attr.set();
  }
  p.tiles_n = (p.C + BN - 1) / BN;
  const long long total = (long long)p.tiles_m * p.tiles_n;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  embed_fused_fwd_kernel<BN><<<grid, EM_THREADS, Cfg::SMEM_BYTES, st>>>(*tv, *tw, p);
  count_launch();
  return check_launch("cavit_embed_fused_fwd");
}

}  // namespace cavit

using namespace cavit;

extern "C" {

int cavit_embed_fused_supported(int32_t B, int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp, int32_t wp,
                                int32_t C) {
  EmbedParams p{};
  return embed_geometry(p, B, M, D, H, W, dp, hp, wp, C, 0);
}

int cavit_embed_fused_wgrad_supported(int32_t B, int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp,
                                      int32_t wp, int32_t C) {
  EmbedParams g{};
  if (!embed_geometry(g, B, M, D, H, W, dp, hp, wp, C, 0) || 64 % g.wp) return 0;
  if (M * g.Wn > EW_BK || (C % 8)) return 0;                 // a brick holds all modalities of >= 1 sample
  const int rows_per_tile = EW_BN / g.wp;
  if (rows_per_tile <= g.hp) return g.hp % rows_per_tile == 0;
  return rows_per_tile % g.hp == 0 && g.dp % (rows_per_tile / g.hp) == 0 && (g.P % EW_BN) == 0;
}

int cavit_embed_fused_fwd(const float* img, const float* W_f32, const float* bias, const float* pos, float* tokens, int32_t B,
                          int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t C,
                          int32_t sample_major, void* stream) {
  if (!img || !W_f32 || !bias || !pos || !tokens) return fail(CAVIT_E_BADARG, "cavit_embed_fused_fwd: null pointer");
  EmbedParams p{};
  if (!embed_geometry(p, B, M, D, H, W, dp, hp, wp, C, sample_major))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_embed_fused_fwd: patch (%d,%d,%d) of (%d,%d,%d), C=%d is outside the fused "
                "kernel's reach (wp in {8,16,32,64,96,...}, hp %% (32/wp) == 0, W %% 4 == 0, W/wp <= 128, C %% 32 == 0)",
                dp, hp, wp, D, H, W, C);
  if ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(W_f32) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(tokens)) & 15)
    return fail(CAVIT_E_BADARG, "cavit_embed_fused_fwd: 16-byte aligned buffers expected");
  p.out = tokens; p.bias = bias; p.pos = pos;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "cavit_embed_fused_fwd: no device status word");
  const int BN = (C % 256 == 0) ? 256 : ((C % 192 == 0) ? 192 : 128);
  const CUtensorMap* tv = volume_map(img, p, p.nv);
  const uint64_t wd[3] = {(uint64_t)p.P, (uint64_t)C, 1};
  const uint64_t ws[2] = {(uint64_t)p.P * 4, (uint64_t)p.P * 4 * (uint64_t)C};
  const uint32_t wbox[3] = {32u, (uint32_t)BN, 1u};
  const CUtensorMap* tw = tensor_map_nd(1, 3, W_f32, wd, ws, wbox, 1);
  if (!tv || !tw) return CAVIT_E_BADARG;
  cudaStream_t st = as_stream(stream);
  if (BN == 256) return embed_fwd_launch<256>(tv, tw, p, st);
  if (BN == 192) return embed_fwd_launch<192>(tv, tw, p, st);
  return embed_fwd_launch<128>(tv, tw, p, st);
}

int cavit_embed_fused_wgrad(const float* img, const void* dtokens_bf16, float* dW, int32_t B, int32_t M, int32_t D, int32_t H,
                            int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t C, int32_t sample_major, void* stream) {
  if (!img || !dtokens_bf16 || !dW) return fail(CAVIT_E_BADARG, "cavit_embed_fused_wgrad: null pointer");
  EmbedParams g{};
  if (!embed_geometry(g, B, M, D, H, W, dp, hp, wp, C, sample_major) || !cavit_embed_fused_wgrad_supported(B, M, D, H, W, dp, hp, wp, C))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_embed_fused_wgrad: shape outside the fused kernel's reach");
  if ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(dtokens_bf16) | reinterpret_cast<uintptr_t>(dW)) & 15)
    return fail(CAVIT_E_BADARG, "cavit_embed_fused_wgrad: 16-byte aligned buffers expected");
  EmbedWgradParams p{};
  p.Wn = g.Wn; p.wp = g.wp; p.M = M; p.Dn = g.Dn; p.Hn = g.Hn; p.hp = g.hp; p.dp = g.dp; p.C = C; p.P = g.P;
  const int rows_per_tile = EW_BN / g.wp;
  if (rows_per_tile <= g.hp) { p.nh = rows_per_tile; p.nz = 1; } else { p.nh = g.hp; p.nz = rows_per_tile / g.hp; }
  p.nbq = EW_BK / (M * g.Wn);
  if (p.nbq > B) p.nbq = B;
  p.nvw = p.nbq * M;
  p.slots = p.nvw * g.Wn;
  p.bgroups = (B + p.nbq - 1) / p.nbq;
  p.bricks = g.Hn * g.Dn * p.bgroups;
  p.tiles_c = (C + 127) / 128;
  p.tiles_p = (g.P + EW_BN - 1) / EW_BN;
  {
    const int tiles = p.tiles_c * p.tiles_p, sms = sm_count();
    int splits = (2 * sms + tiles - 1) / tiles;           // about two waves of work items, each with >= 8 bricks
    if (splits > p.bricks / 8) splits = p.bricks / 8;
    if (splits < 1) splits = 1;
    p.bricks_per_split = (p.bricks + splits - 1) / splits;
    p.splits = (p.bricks + p.bricks_per_split - 1) / p.bricks_per_split;
  }
  p.land_bytes = p.nvw * p.nz * p.nh * g.Wn * g.wp * 4;
  p.a_bytes = p.slots * 128;
  p.dW = dW;
  p.status = status_word();
  if (!p.status) return fail(CAVIT_E_DEVICE, "cavit_embed_fused_wgrad: no device status word");
  const CUtensorMap* tv = volume_brick_map(img, g, p.nh, p.nz, p.nvw);
  // token gradients, CLS rows skipped by the base pointer: dims {c, u, w, m, b} with token index = u + U * w, where a
  // brick's tokens are w0 .. w0 + Wn - 1 at fixed u: u = di, w = hi*Wn + wi in the volume frame; u = hi', w = wi' in the
  // slice frame (embed_geometry)
  const long long Ntok = g.Ntok, Np = g.Np;
  const uint64_t m_stride = sample_major ? (uint64_t)Np * C : (uint64_t)B * Ntok * C;
  const uint64_t b_stride = (uint64_t)Ntok * C;
  const uint64_t U = g.tz ? (uint64_t)g.Dn : (uint64_t)g.Hn, Wext = g.tz ? (uint64_t)g.Hn * g.Wn : (uint64_t)g.Wn;
  p.slice_frame = g.tz ? 0 : 1;
  const uint64_t dims[5] = {(uint64_t)C, U, Wext, (uint64_t)M, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)C * 2, U * C * 2, m_stride * 2, b_stride * 2};
  const uint32_t box[5] = {64u, 1u, (uint32_t)g.Wn, (uint32_t)M, (uint32_t)p.nbq};
  const CUtensorMap* td = tensor_map_nd(0, 5, reinterpret_cast<const uint8_t*>(dtokens_bf16) + (size_t)C * 2, dims, strides, box, 1);
  if (!tv || !td) return CAVIT_E_BADARG;
  cudaStream_t st = as_stream(stream);
  static PerDeviceFlag attr;
  if (attr.unset()) {
    cudaError_t e = cudaFuncSetAttribute(embed_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM_BYTES);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "embed wgrad smem attribute: %s", cudaGetErrorString(e));
    attr.set();
  }
  cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)C * g.P, st);
  const long long total = (long long)p.tiles_c * p.tiles_p * p.splits;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  embed_wgrad_kernel<<<grid, EW_THREADS, EW_SMEM_BYTES, st>>>(*tv, *td, p);
  count_launch();
  return check_launch("cavit_embed_fused_wgrad");
}

int cavit_embed_bias_grad(const float* dpos, float* db, int32_t N, int32_t C, void* stream) {
  if (!dpos || !db || N < 1 || C < 1) return fail(CAVIT_E_BADARG, "cavit_embed_bias_grad: bad args");
  embed_bias_grad_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(dpos, db, N, C);
  count_launch();
  return check_launch("cavit_embed_bias_grad");
}

}  // extern "C"
