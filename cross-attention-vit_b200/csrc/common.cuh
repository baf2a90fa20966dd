// cavit-sm100 — shared device helpers: PTX wrappers for mbarrier / TMA / tcgen05 (sm_100a only).
// Every mbarrier wait is BOUNDED: a protocol bug must never hang the GPU box. On time-out the
// waiting thread records a code in the library's device status word and raises a CTA-wide abort
// flag so that the kernel drains in bounded time with garbage results; cavit_device_status()
// reports it to the host.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cavit {

typedef __nv_bfloat16 bf16;

enum : int {
  ERR_NONE = 0,
  ERR_TIMEOUT_FULL = 101,
  ERR_TIMEOUT_EMPTY = 102,
  ERR_TIMEOUT_TMEM_FULL = 103,
  ERR_TIMEOUT_TMEM_EMPTY = 104,
  ERR_TIMEOUT_ATTN = 105,
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}

// Bounded wait. `abort_flag` lives in shared memory (one per CTA); `status` is the global word.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag,
                                          int* status, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 255u) == 0u) {
      if (*abort_flag) return;
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) {  // 2 s
        *abort_flag = 1;
        atomicCAS(status, 0, code);
        return;
      }
    }
  }
}

// One elected lane of a fully converged warp (the compiler turns `if (elect_one())` into ELECT + predicated
// issue; with `if (lane == 0)` it cannot prove a single active thread and wraps every tcgen05.mma in a loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a tensor tile (no shared memory, no barrier): issued one work item ahead so that the later
// cp.async.bulk.tensor load of the same box is an L2 hit instead of a ~3000-cycle DRAM round trip.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
// TMA load issued by either CTA of a pair into its OWN shared memory; the transaction bytes are credited to the
// LEADER CTA's mbarrier (peer bit of the barrier address cleared), which is the one the MMA issuer waits on.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: 256 x N x 16, issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once all previously issued MMAs of the pair have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05 ops of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (base+i), columns c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors (SWIZZLE_128B)
// Shared-memory matrix descriptor (64 bit): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout type [61,64) (2 = SWIZZLE_128B).
// K-major tile  [rows][64 bf16]: 128 B per row, 8-row swizzle atoms 1024 B apart  -> SBO = 1024.
// MN-major tile [k rows][64 bf16 of M/N]: 8-k-row atoms 1024 B apart (SBO), next 64-wide M/N chunk
// `mn_chunk_bytes` further (LBO).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same with an explicit layout type: 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B (rows of 128 / 64 / 32 bytes; SBO = 8 rows).
__device__ __forceinline__ uint64_t umma_desc_layout(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n, int a_mn, int b_mn, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------- math
// Exact-erf GELU (nn.GELU default) evaluated with the Abramowitz-Stegun 7.1.26 rational form of
// erfc (|abs err| < 1.5e-7, far below bf16 resolution) so that the fused GEMM epilogues need one
// MUFU.EX2, one MUFU.RCP and ~10 FMAs per element instead of libdevice erff + expf:
//   Phi(u) = 0.5 erfc(-u / sqrt 2);  erfc(z) = poly(t) e^{-z^2}, t = 1 / (1 + p z), z >= 0
//   gelu(u) = u Phi(u);  gelu'(u) = Phi(u) + u e^{-u^2/2} / sqrt(2 pi)   (same exponential)
__device__ __forceinline__ void gelu_phi_terms(float u, float& phi, float& e) {
  const float z = fabsf(u) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));  // e = exp(-u^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * e;  // 0.5 erfc(|u| / sqrt 2)
  phi = (u >= 0.f) ? 1.0f - half_erfc : half_erfc;
}
__device__ __forceinline__ float gelu_erf(float u) {
  float phi, e;
  gelu_phi_terms(u, phi, e);
  return u * phi;
}
__device__ __forceinline__ float gelu_erf_grad(float u) {
  float phi, e;
  gelu_phi_terms(u, phi, e);
  return fmaf(u * 0.3989422804014327f, e, phi);
}
// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issued instruction)
__device__ __forceinline__ unsigned long long f2_as_u64(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 u64_as_f2(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)), "l"(f2_as_u64(c)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

// Fused-epilogue GELU for TWO pre-activations per call, built for issue slots: the GEMM epilogues are bounded by
// instruction issue and by the 16-lane MUFU, so this form uses ONE MUFU per element and packed FFMA2 math:
//   e = exp(-u^2/2);  R(a) = Phi(-a) / e = erfcx(a / sqrt 2) / 2  ~  degree-7 minimax polynomial on a = min(|u|, 5)
//   gelu(u)  = max(u, 0) - |u| e R                     (= u Phi(u), exact-erf nn.GELU, model_cross.py:24)
//   gelu'(u) = [u >= 0] - sign(u) e (R - a / sqrt(2 pi))
// |rel err| of R <= 4.2e-4 on [0, 5] (for |u| > 5 the term is < 2e-6 absolute), i.e. |gelu err| <= 7e-5 and
// |gelu' err| <= 2.1e-4 absolute: an order of magnitude below the bf16 rounding of the stored result.
__device__ __forceinline__ void gelu_pair_terms(float2 u, float2& e, float2& r, float2& ac) {
  ac = make_float2(fminf(fabsf(u.x), 5.0f), fminf(fabsf(u.y), 5.0f));
  const float2 earg = mul2(mul2(u, u), splat2(-0.72134752044448170f));  // -u^2/2 * log2(e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(earg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(earg.y));
  r = fma2(ac, splat2(-2.501152098e-05f), splat2(5.606001891e-04f));
  r = fma2(r, ac, splat2(-5.354749036e-03f));
  r = fma2(r, ac, splat2(2.881915092e-02f));
  r = fma2(r, ac, splat2(-9.833444611e-02f));
  r = fma2(r, ac, splat2(2.302476772e-01f));
  r = fma2(r, ac, splat2(-3.942042539e-01f));
  r = fma2(r, ac, splat2(4.997907545e-01f));
}
__device__ __forceinline__ float2 gelu_pair(float2 u) {
  float2 e, r, ac;
  gelu_pair_terms(u, e, r, ac);
  const float2 w = mul2(u, mul2(e, r));  // |w| = |u| Phi(-|u|)
  return make_float2(fmaxf(u.x, 0.f) - fabsf(w.x), fmaxf(u.y, 0.f) - fabsf(w.y));
}
__device__ __forceinline__ float2 gelu_grad_pair(float2 u) {
  float2 e, r, ac;
  gelu_pair_terms(u, e, r, ac);
  const float2 t = mul2(e, fma2(ac, splat2(-0.3989422804014327f), r));  // e (R - a / sqrt(2 pi)): gelu'(-|u|)
  // u < 0: t ; u >= 0: 1 - t
  return make_float2(u.x < 0.f ? t.x : 1.0f - t.x, u.y < 0.f ? t.y : 1.0f - t.y);
}

// Scalar twins of gelu_pair / gelu_grad_pair as the GEMM epilogue's fast path evaluates them (gemm.cu: gelu_batch,
// gelu_grad_batch): the SAME operations in the same order, so that a row gets bit-identical results whether its 32-row slab
// takes the vectorised fast path or the generic one — which slab is partial depends on the batch size, and a sample's logits
// must not (tests/test_gpu_ddp_nccl.py: a shard of the batch reproduces the global-batch run).
__device__ __forceinline__ void gelu_poly_terms(float u, float sg, float& e, float& r, float& ac) {
  ac = fminf(fabsf(u), 5.0f);
  const float earg = (u * u) * -0.72134752044448170f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(earg));
  r = fmaf(ac, sg * -2.501152098e-05f, sg * 5.606001891e-04f);
  r = fmaf(r, ac, sg * -5.354749036e-03f);
  r = fmaf(r, ac, sg * 2.881915092e-02f);
  r = fmaf(r, ac, sg * -9.833444611e-02f);
  r = fmaf(r, ac, sg * 2.302476772e-01f);
  r = fmaf(r, ac, sg * -3.942042539e-01f);
  r = fmaf(r, ac, sg * 4.997907545e-01f);
}
__device__ __forceinline__ float gelu_poly(float u) {
  float e, r, ac;
  gelu_poly_terms(u, -1.0f, e, r, ac);
  return fmaf(ac, __fmul_rn(e, r), fmaxf(u, 0.f));
}
__device__ __forceinline__ float gelu_poly_grad(float u) {
  float e, r, ac;
  gelu_poly_terms(u, 1.0f, e, r, ac);
  const float t = __fmul_rn(e, fmaf(ac, -0.3989422804014327f, r));
  return u < 0.f ? t : 1.0f - t;
}

// ---------------------------------------------------------------- dropout
// Counter-based Bernoulli masks: keep(i) is a pure function of (seed, site, element index), so the
// backward pass regenerates exactly the mask the forward pass used without storing it. One
// splitmix64 evaluation yields two 32-bit uniforms (elements 2j and 2j+1 of a site).
struct DropCfg {
  const unsigned long long* seed;  // device scalar (changes every step; graph-replay safe)
  uint32_t site;                   // which dropout module
  uint32_t thresh;                 // p * 2^32: keep iff uniform >= thresh
  float inv_keep;                  // 1 / (1 - p)
};
__device__ __forceinline__ uint64_t drop_hash(uint64_t seed, uint32_t site, uint64_t pair_idx) {
  uint64_t x = pair_idx * 0x9E3779B97F4A7C15ull + (seed ^ (static_cast<uint64_t>(site) * 0xD1B54A32D192ED03ull));
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
// multiplier (0 or 1/(1-p)) for element idx
__device__ __forceinline__ float drop_mult(const DropCfg& d, uint64_t seed, uint64_t idx) {
  const uint64_t h = drop_hash(seed, d.site, idx >> 1);
  const uint32_t u = (idx & 1) ? static_cast<uint32_t>(h >> 32) : static_cast<uint32_t>(h);
  return u >= d.thresh ? d.inv_keep : 0.f;
}
// multipliers for elements idx (even) and idx+1
__device__ __forceinline__ float2 drop_mult2(const DropCfg& d, uint64_t seed, uint64_t even_idx) {
  const uint64_t h = drop_hash(seed, d.site, even_idx >> 1);
  return make_float2(static_cast<uint32_t>(h) >= d.thresh ? d.inv_keep : 0.f,
                     static_cast<uint32_t>(h >> 32) >= d.thresh ? d.inv_keep : 0.f);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}
// two bit operations per pair (the intrinsic path goes through PRMT + shift per element)
__device__ __forceinline__ float2 unpack_bf16_fast(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cavit
