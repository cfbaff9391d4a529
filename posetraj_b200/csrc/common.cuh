// posetraj_b200 — shared device helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers for mbarrier / TMA / tcgen05 (TMEM) plus a few vector
// load/store and bf16 helpers.  Everything here is sm_100a-only on purpose: there is
// no fallback path ("no CPU fallback, no multi-backend dispatch").
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define PT_DEVICE __device__ __forceinline__

namespace pt {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------
PT_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch (see launch.h).  `griddep_wait` returns once every grid this one depends on has
// completed and its memory is visible; `griddep_launch` lets the next grid in the stream start being scheduled.
// Both are no-ops for a grid launched without the attribute.
// Compiled in only with -DPT_ENABLE_PDL (measured slower inside the captured step, profiles/r1h_pdl_gn.md, and the
// instructions are not free even without the launch attribute: PREEXIT shows up in GroupNorm's stall samples).
#ifdef PT_ENABLE_PDL
PT_DEVICE void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PT_DEVICE void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
PT_DEVICE void griddep_wait() {}
PT_DEVICE void griddep_launch() {}
#endif

PT_DEVICE uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// One lane of a CONVERGED warp (deterministic: the same lane for the same mask).  The producer / MMA-issue warps run
// their loops with all 32 lanes and branch on this only around the asynchronous instruction itself: every operand of
// the TMA / tcgen05 instruction is then computed in warp-uniform control flow and lives in uniform registers.  Inside
// an `if (lane == 0)` region the same operands sit in vector registers and ptxas wraps EVERY UTMALDG / UTCHMMA in a
// "waterfall" (ELECT + 5-7 R2UR.BROADCAST + loop branch), ~120 cycles per instruction — more than a 128 x 128 x 16
// MMA takes to execute (profiles/r2b_mlp.md).
PT_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// a value every lane holds identically, made provably warp-uniform for the compiler (uniform-register allocation)
PT_DEVICE uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

PT_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

PT_DEVICE float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// x * sigmoid(x) with MUFU ex2 / rcp (relative error ~2e-7): the IEEE division of x / (1 + exp(-x)) costs ~10
// instructions and the GroupNorm+SiLU pass is issue-bound with it (profiles/r1g_groupnorm.md)
PT_DEVICE float silu_f(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// SiLU of x = 2h as h + h tanh(h) with ONE MUFU (tanh.approx.f32, max relative error 2^-11 on the tanh).  Relative error
// of the result <= 2.5e-4 for x >= -1 and below the bf16 rounding (2^-9) down to x = -3; further out in the negative tail
// (0.13 % of a unit normal) the cancellation in 1 + tanh(h) leaves an ABSOLUTE error <= |h| 4.9e-4 ~ 1e-3 on values
// |silu| < 0.14, i.e. the size of the bf16 rounding of an ordinary O(0.5) activation.  For the GroupNorm + SiLU apply phase,
// which is MUFU-queue bound with ex2 + rcp per element (ncu stall_mio on the MUFU instructions, profiles/r3_glue_kernels.md).
PT_DEVICE float silu_half_tanh(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// GEGLU gate  value * gelu_erf(g)  = value * g * Phi(g)  for the tensor-core epilogues (gemm.cu, mlp.cu), where it is
// evaluated 128 x 64 times per hidden chunk and bounds the kernel (profiles/r1b, r2).  Phi(g) is evaluated as
//     sigmoid(2 g (c0 + c1 s + c2 s^2)),  s = min(g^2, 81)
// — the tanh-form with a quintic argument re-fitted to the EXACT erf GELU (tests/test_gate_approx_cpu.py: max
// |g Phi~(g) - g Phi(g)| = 2.6e-5 over all g, 150x below the bf16 rounding of the result; the textbook tanh constants
// give 4.7e-4).  8 FP32 + 2 MUFU per gate instead of 26 + 2 for the Abramowitz-Stegun 7.1.26 evaluation used before.
// The constants carry the factor -2 log2(e) of the exponential.
PT_DEVICE float geglu_gate_fast(float value, float g) {
  const float s = fminf(g * g, 81.0f);
  float poly = fmaf(-3.53110710e-04f * -2.885390081777927f, s, 3.70155061e-02f * -2.885390081777927f);
  poly = fmaf(poly, s, 7.97496957e-01f * -2.885390081777927f);
  const float e = ex2_approx(poly * g);          // exp(-2u)
  return value * g * rcp_approx(1.0f + e);       // g * sigmoid(2u); e = +inf for very negative g gives exactly 0
}

// The same gate with ONE MUFU: sigmoid(2u) = (1 + tanh(u)) / 2, so value * g * Phi(g) = hv + hv tanh(u) with
// hv = value * g / 2 and u = g (c0 + c1 s + c2 s^2) (the same re-fitted quintic, without the exponential's -2 log2(e)).
// tanh.approx.f32 is good to 2^-11 relative: the result is within 2.5e-4 relative for g >= -1; in the negative tail the
// cancellation 1 + tanh(u) leaves an absolute error <= |hv| 4.9e-4, below the bf16 rounding of an ordinary gated value
// (same argument as silu_half_tanh).  11.5 issue slots and one MUFU per gate instead of 12.5 and two: the gate evaluation
// bounds the fused feed-forward (mlp.cu) and its MUFU queue is what the gate warps stall on.
PT_DEVICE float geglu_gate_tanh(float value, float g) {
  const float s = fminf(g * g, 81.0f);
  float poly = fmaf(-3.53110710e-04f, s, 3.70155061e-02f);
  poly = fmaf(poly, s, 7.97496957e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(poly * g));
  const float hv = (0.5f * value) * g;
  return fmaf(hv, t, hv);
}

// exact (erf) GELU, as diffusers' GEGLU uses F.gelu(gate) with approximate="none"
PT_DEVICE float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

PT_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

PT_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

PT_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

PT_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------
PT_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

PT_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

PT_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

PT_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

PT_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

PT_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a pipeline bug becomes a trap (reported as a CUDA error through the C ABI)
// instead of a hung GPU.  try_wait is a HW-suspended wait, so the bound is generous.
PT_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      asm volatile("trap;");
    }
  }
}

// ----------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, shared::cta destination of the issuing CTA
// ----------------------------------------------------------------------------------
PT_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

PT_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

PT_DEVICE void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                           int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------------
// thread-block clusters / CTA pairs (cta_group::2)
// ----------------------------------------------------------------------------------
PT_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

PT_DEVICE uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}

PT_DEVICE uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}

// all threads of all CTAs of the cluster
PT_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
PT_DEVICE uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}

PT_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Remote arrive WITHOUT the cluster-scope release: `mbarrier.arrive.release.cluster` above makes ptxas emit
// MEMBAR.ALL.GPU + ERRBAR in front of the arrive (12 % of all stall samples of the fused feed-forward, whose gate warps
// signal the leader CTA twice per hidden chunk: profiles/r2b_mlp.md).  What those signals order is either a finished
// tcgen05.ld (tcgen05.fence::before_thread_sync) or shared-memory writes of the signalling CTA that its own tensor core
// will read (fence.proxy.async.shared::cta), so the default CTA-scope release is enough — the form CUTLASS uses for
// the accumulator-empty barriers of its 2-SM kernels.
PT_DEVICE void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Remote arrive with RELAXED semantics: for signals that order nothing but finished tcgen05.ld reads (already ordered
// by tcgen05.wait::ld + tcgen05.fence::before_thread_sync).  A releasing arrive has to wait for every earlier load of
// the thread, and ptxas therefore sinks it below the code that consumes those loads — in the fused feed-forward that
// delayed "accumulator drained" by the whole gate evaluation of the chunk (profiles/r2b_mlp.md).
PT_DEVICE void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMA loads of a CTA pair: data lands in THIS CTA's smem, the transaction bytes are credited to an mbarrier that
// may live in the peer (leader) CTA — `bar_cluster_addr` is a shared::cluster address.
PT_DEVICE void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

PT_DEVICE void tma_load_3d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0, int32_t c1,
                                int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------
PT_DEVICE void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one converged warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

PT_DEVICE void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// arrive on the mbarrier at this smem offset in every CTA of `cta_mask` once all prior cta_group::2 MMAs retire
PT_DEVICE void tc_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA): 256 x N x 16, leader CTA only
PT_DEVICE void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

PT_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, converged
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

PT_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp, converged
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

PT_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PT_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PT_DEVICE void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
PT_DEVICE void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire
PT_DEVICE void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, single CTA
PT_DEVICE void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1").
//   bits [ 0,14) start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) layout: 0 none, 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B
PT_DEVICE uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

// K-major bf16 tile whose rows are 128 B (64 elements) wide, written by TMA with SWIZZLE_128B:
// 8-row swizzle atoms of 1024 B, stacked along M/N -> SBO = 1024, LBO unused (1).
PT_DEVICE uint64_t make_desc_kmajor_sw128(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024, 2); }

// Instruction descriptor for kind::f16, A/B = bf16, D = fp32.  a_major/b_major: 0 = K-major, 1 = MN-major.
PT_DEVICE uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_major, uint32_t b_major) {
  uint32_t d = 0;
  d |= 1u << 4;            // D format: F32
  d |= 1u << 7;            // A format: BF16
  d |= 1u << 10;           // B format: BF16
  d |= (a_major & 1u) << 15;
  d |= (b_major & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (base_lane + i), columns [c, c+32)
PT_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// 32 lanes x 32 columns store: thread i writes lane (base_lane + i), columns [c, c+32)
PT_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

PT_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

PT_DEVICE void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

PT_DEVICE void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// ----------------------------------------------------------------------------------
// vector global access
// ----------------------------------------------------------------------------------
PT_DEVICE uint4 ldg_u4(const void* p) { return *reinterpret_cast<const uint4*>(p); }
PT_DEVICE void stg_u4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

PT_DEVICE uint4 ldg_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace pt
