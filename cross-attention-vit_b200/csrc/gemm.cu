// cavit-sm100 — K-GEMM: persistent, warp-specialised tcgen05 GEMM with fused epilogues.
//
//   D[g][m][n] = sum_k A_g(m,k) * B_g(n,k)      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM, 192 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, 128 x BN x 16 per instruction) + TMEM alloc
//   warps 2..5  epilogue       (tcgen05.ld -> bias / GELU / residual / ... -> global)
// Two TMEM accumulator stages (2 x BN fp32 columns) let the epilogue of tile i overlap the
// mainloop of tile i+1. Tiles are walked in a static persistent schedule (tile = cta + i*grid).
// Operands may be K-major or MN-major (the latter is what dgrad's W and wgrad's dY / X are), so
// no transposed copies of weights or activations are ever materialised.
//
// Replaces the reference's cuBLAS `addmm`/`mm` call sites (SURVEY.md §2.2 G1-G4, X2, B*).
#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmDev {
  int M, N, K, groups;
  int a_mn, b_mn, epi, out_fp32, accumulate, embed_np, split_k, kb_per_split;
  void* out; long long ldo, out_gs;
  const float* bias; long long bias_gs;
  const float* resid; long long ldr, resid_gs;
  void* aux; long long ldaux, aux_gs;
  int tiles_m, tiles_n;
  int* status;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// Epilogue for one 32-column chunk held by one thread (one output row).
__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, int g, long long row, int col0, bool row_ok,
                                               uint32_t (&acc)[32]) {
  if (!row_ok) return;
  const int ncols = min(32, p.N - col0);
  if (ncols <= 0) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);

  const float* bias = p.bias ? p.bias + (long long)g * p.bias_gs + col0 : nullptr;
  long long out_row = row;
  if (p.epi == CAVIT_EPI_EMBED) {
    const long long b = row / p.embed_np, t = row % p.embed_np;
    out_row = b * (p.embed_np + 1) + 1 + t;
  }
  const bool full = (ncols == 32);

  if (bias) {
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + i));
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
      }
    } else {
      for (int i = 0; i < ncols; ++i) v[i] += __ldg(bias + i);
    }
  }

  if (p.epi == CAVIT_EPI_BIAS_GELU) {
    bf16* aux = reinterpret_cast<bf16*>(p.aux) + (long long)g * p.aux_gs + row * p.ldaux + col0;
    if (full && (p.ldaux % 8 == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q;
        q.x = pack_bf16(v[i], v[i + 1]); q.y = pack_bf16(v[i + 2], v[i + 3]);
        q.z = pack_bf16(v[i + 4], v[i + 5]); q.w = pack_bf16(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(aux + i) = q;
      }
    } else {
      for (int i = 0; i < ncols; ++i) aux[i] = __float2bfloat16(v[i]);
    }
    // GELU is applied to the bf16-rounded pre-activation so that backward (which reads u as
    // bf16) differentiates exactly the function the forward evaluated.
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(__bfloat162float(__float2bfloat16(v[i])));
  } else if (p.epi == CAVIT_EPI_GELU_BWD) {
    const bf16* aux = reinterpret_cast<const bf16*>(p.aux) + (long long)g * p.aux_gs + row * p.ldaux + col0;
    if (full && (p.ldaux % 8 == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 q = *reinterpret_cast<const uint4*>(aux + i);
        float2 f;
        f = unpack_bf16(q.x); v[i] *= gelu_erf_grad(f.x); v[i + 1] *= gelu_erf_grad(f.y);
        f = unpack_bf16(q.y); v[i + 2] *= gelu_erf_grad(f.x); v[i + 3] *= gelu_erf_grad(f.y);
        f = unpack_bf16(q.z); v[i + 4] *= gelu_erf_grad(f.x); v[i + 5] *= gelu_erf_grad(f.y);
        f = unpack_bf16(q.w); v[i + 6] *= gelu_erf_grad(f.x); v[i + 7] *= gelu_erf_grad(f.y);
      }
    } else {
      for (int i = 0; i < ncols; ++i) v[i] *= gelu_erf_grad(__bfloat162float(aux[i]));
    }
  } else if (p.epi == CAVIT_EPI_BIAS_RESID || p.epi == CAVIT_EPI_EMBED) {
    const float* r;
    if (p.epi == CAVIT_EPI_BIAS_RESID)
      r = p.resid + (long long)g * p.resid_gs + row * p.ldr + col0;
    else
      r = p.resid + (1 + row % p.embed_np) * p.ldr + col0;  // positional embedding row 1 + t
    if (full && (p.ldr % 4 == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(r + i);
        v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
      }
    } else {
      for (int i = 0; i < ncols; ++i) v[i] += r[i];
    }
  }

  if (p.split_k > 1) {  // partial tile of a split-K reduction: combine with fp32 atomics
    float* o = reinterpret_cast<float*>(p.out) + (long long)g * p.out_gs + out_row * p.ldo + col0;
    for (int i = 0; i < ncols; ++i) atomicAdd(o + i, v[i]);
    return;
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + (long long)g * p.out_gs + out_row * p.ldo + col0;
    if (full && (p.ldo % 4 == 0)) {
      if (p.accumulate) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 o4 = *reinterpret_cast<float4*>(o + i);
          o4.x += v[i]; o4.y += v[i + 1]; o4.z += v[i + 2]; o4.w += v[i + 3];
          *reinterpret_cast<float4*>(o + i) = o4;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
      for (int i = 0; i < ncols; ++i) o[i] = p.accumulate ? o[i] + v[i] : v[i];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)g * p.out_gs + out_row * p.ldo + col0;
    if (full && (p.ldo % 8 == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q;
        q.x = pack_bf16(v[i], v[i + 1]); q.y = pack_bf16(v[i + 2], v[i + 3]);
        q.z = pack_bf16(v[i + 4], v[i + 5]); q.w = pack_bf16(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = q;
      }
    } else {
      for (int i = 0; i < ncols; ++i) o[i] = __float2bfloat16(v[i]);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmDev p) {
  using Cfg = GemmCfg<BN>;
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int splits = p.split_k > 1 ? p.split_k : 1;
  const int tiles_per_group = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_per_group * p.groups * splits;  // work item = (group, m tile, n tile, k split)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
        const int split = work % splits;
        const int tile = work / splits;
        const int g = tile / tiles_per_group;
        const int rem = tile - g * tiles_per_group;
        const int m0 = (rem / p.tiles_n) * GEMM_BM;
        const int n0 = (rem % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_EMPTY);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const int k0 = kb * GEMM_BK;
          if (!p.a_mn) {
            tma_load_3d(&tmA, full_bar(stage), sA, k0, m0, g);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c)
              tma_load_3d(&tmA, full_bar(stage), sA + c * (GEMM_BK * 128), m0 + c * 64, k0, g);
          }
          if (!p.b_mn) {
            tma_load_3d(&tmB, full_bar(stage), sB, k0, n0, g);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_3d(&tmB, full_bar(stage), sB + c * (GEMM_BK * 128), n0 + c * 64, k0, g);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN, p.a_mn, p.b_mn);
      // descriptor strides: K-major: SBO = 1024 (8 rows x 128 B); MN-major: LBO = 64-wide chunk
      // pitch (BK rows x 128 B), SBO = 1024 (8 k-rows x 128 B).
      const uint32_t a_lbo = p.a_mn ? GEMM_BK * 128 : 16, b_lbo = p.b_mn ? GEMM_BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? 16 * 128 : 32, b_kstep = p.b_mn ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
        const int split = work % splits;
        const int kb0 = split * p.kb_per_split;
        const int num_kb = min(num_kb_total, kb0 + p.kb_per_split) - kb0;
        mbar_wait(tempty_bar(as), aphase ^ 1u, abort_flag, p.status, ERR_TIMEOUT_TMEM_EMPTY);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, abort_flag, p.status, ERR_TIMEOUT_FULL);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = umma_desc_sw128(sA + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = umma_desc_sw128(sB + k * b_kstep, b_lbo, 1024);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
      const int tile = work / splits;
      const int g = tile / tiles_per_group;
      const int rem = tile - g * tiles_per_group;
      const int m0 = (rem / p.tiles_n) * GEMM_BM;
      const int n0 = (rem % p.tiles_n) * BN;
      mbar_wait(tfull_bar(as), aphase, abort_flag, p.status, ERR_TIMEOUT_TMEM_FULL);
      tc_fence_after();
      const long long row = (long long)m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n0 + c * 32 >= p.N) break;  // warp-uniform
        uint32_t acc[32];
        tmem_ld32(t_row + c * 32, acc);
        tmem_ld_wait();
        epilogue_chunk(p, g, row, n0 + c * 32, row_ok, acc);
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

template <int BN>
static int launch_gemm(const CUtensorMap* ta, const CUtensorMap* tb, GemmDev& d, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(CAVIT_E_LAUNCH, "gemm smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  d.tiles_m = (d.M + GEMM_BM - 1) / GEMM_BM;
  d.tiles_n = (d.N + BN - 1) / BN;
  const int num_kb = (d.K + GEMM_BK - 1) / GEMM_BK;
  if (d.split_k > num_kb) d.split_k = num_kb;
  if (d.split_k < 1) d.split_k = 1;
  d.kb_per_split = (num_kb + d.split_k - 1) / d.split_k;
  d.split_k = (num_kb + d.kb_per_split - 1) / d.kb_per_split;  // no empty splits
  const long long total = (long long)d.tiles_m * d.tiles_n * d.groups * d.split_k;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  gemm_bf16_tcgen05_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(*ta, *tb, d);
  count_launch();
  return check_launch("cavit_gemm");
}

}  // namespace cavit

using namespace cavit;

extern "C" int cavit_gemm(const cavit_gemm_args* a, void* stream) {
  if (!a) return fail(CAVIT_E_BADARG, "cavit_gemm: null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->groups <= 0)
    return fail(CAVIT_E_BADARG, "cavit_gemm: non-positive extent M=%d N=%d K=%d groups=%d", a->M, a->N, a->K, a->groups);
  if (!a->A || !a->B || !a->out) return fail(CAVIT_E_BADARG, "cavit_gemm: null operand");
  if ((a->lda % 8) || (a->ldb % 8) || (a->a_gs % 8) || (a->b_gs % 8) ||
      (reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: operands need 16-byte aligned rows (ld %% 8 == 0)");
  // inner (contiguous) extents must also keep TMA's 16-byte rule
  const int a_inner = a->a_mn ? a->M : a->K, b_inner = a->b_mn ? a->N : a->K;
  if ((a_inner % 8) || (b_inner % 8))
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_gemm: contiguous extent must be a multiple of 8 (A %d, B %d)", a_inner, b_inner);
  if (a->epi < CAVIT_EPI_NONE || a->epi > CAVIT_EPI_EMBED) return fail(CAVIT_E_BADARG, "cavit_gemm: bad epilogue %d", a->epi);
  if ((a->epi == CAVIT_EPI_BIAS_GELU || a->epi == CAVIT_EPI_GELU_BWD) && !a->aux)
    return fail(CAVIT_E_BADARG, "cavit_gemm: epilogue %d needs aux", a->epi);
  if ((a->epi == CAVIT_EPI_BIAS_RESID || a->epi == CAVIT_EPI_EMBED) && (!a->resid || !a->out_fp32))
    return fail(CAVIT_E_BADARG, "cavit_gemm: residual epilogues need resid and fp32 out");
  if (a->epi == CAVIT_EPI_EMBED && a->embed_np <= 0) return fail(CAVIT_E_BADARG, "cavit_gemm: embed_np");
  if (a->accumulate && !a->out_fp32) return fail(CAVIT_E_BADARG, "cavit_gemm: accumulate needs fp32 out");
  if (a->split_k > 1 && (!a->out_fp32 || a->epi != CAVIT_EPI_NONE))
    return fail(CAVIT_E_BADARG, "cavit_gemm: split_k needs fp32 out and EPI_NONE");
  int* status = status_word();
  if (!status) return fail(CAVIT_E_DEVICE, "cavit_gemm: no device status word");

  // N tile: 256 when it divides the work well, else 128 (short N) — both keep 128-row M tiles.
  const int BN = (a->N >= 256 && (a->N % 256 == 0 || a->N > 1024)) ? 256 : 128;
  const CUtensorMap *ta, *tb;
  if (!a->a_mn)
    ta = tensor_map_bf16_3d(a->A, a->K, a->M, a->groups, a->lda, a->a_gs, 64, GEMM_BM);
  else
    ta = tensor_map_bf16_3d(a->A, a->M, a->K, a->groups, a->lda, a->a_gs, 64, GEMM_BK);
  if (!ta) return CAVIT_E_BADARG;
  if (!a->b_mn)
    tb = tensor_map_bf16_3d(a->B, a->K, a->N, a->groups, a->ldb, a->b_gs, 64, BN);
  else
    tb = tensor_map_bf16_3d(a->B, a->N, a->K, a->groups, a->ldb, a->b_gs, 64, GEMM_BK);
  if (!tb) return CAVIT_E_BADARG;

  GemmDev d;
  d.M = a->M; d.N = a->N; d.K = a->K; d.groups = a->groups;
  d.a_mn = a->a_mn ? 1 : 0; d.b_mn = a->b_mn ? 1 : 0;
  d.epi = a->epi; d.out_fp32 = a->out_fp32; d.accumulate = a->accumulate; d.embed_np = a->embed_np;
  d.out = a->out; d.ldo = a->ldo; d.out_gs = a->out_gs;
  d.bias = a->bias; d.bias_gs = a->bias_gs;
  d.resid = a->resid; d.ldr = a->ldr; d.resid_gs = a->resid_gs;
  d.aux = a->aux; d.ldaux = a->ldaux; d.aux_gs = a->aux_gs;
  d.tiles_m = d.tiles_n = 0;
  d.split_k = a->split_k > 1 ? a->split_k : 1;
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
  if (BN == 256) return launch_gemm<256>(ta, tb, d, as_stream(stream));
  return launch_gemm<128>(ta, tb, d, as_stream(stream));
}
