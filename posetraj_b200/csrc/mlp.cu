// posetraj_b200 — fused GEGLU feed-forward (diffusers FeedForward: net.0 = GEGLU proj, net.2 = Linear) on tcgen05.
//
//   out[m, :] = acc_scale * (GEGLU(x[m, :] W1^T + b1) W2^T + b2) + res1_scale * res1[m, :] + res2_scale * res2[m, :]
//   GEGLU(h) = h[:, :H] * gelu_erf(h[:, H:])                                   (H = hidden = 4C)
//
// Replaces the pair  pt_gemm(geglu) -> [rows, 4C] round trip through HBM -> pt_gemm(ff.out)  of every
// BasicTransformerBlock.ff / TemporalBasicTransformerBlock.ff_in / .ff at the widths whose output accumulator fits
// TMEM next to the hidden one (C <= 320, i.e. level 0 of the SVD UNet: 21 feed-forwards per denoise step, each of
// which wrote and re-read a 206 MB hidden tensor — profiles/r1j_gemm_full.md).  Reference semantics: diffusers 0.24.0
// `FeedForward` / `GEGLU` as wired by models/modified_svd.py:70-74,100-107.
//
// One CTA PAIR (cta_group::2) owns a 256-row tile (128 rows per CTA) and walks the hidden dimension in chunks of 64:
//     GEMM1(j): acc1[128 x 128 fp32 per CTA] = X[256 x C] . W1[value rows j | gate rows j]^T        (TMEM cols 0..127)
//     gate   : P_j = bf16( (v + b1v) * gelu(g + b1g) )   -> swizzled smem tile, the A operand of
//     GEMM2(j): acc2[128 x C fp32 per CTA] += P_j[256 x 64] . W2[:, chunk j]^T                      (TMEM cols 128..128+C)
// so the hidden activations never leave the SM.  X stays resident in shared memory for the whole tile; W1 / W2 stream
// through two TMA rings, each CTA staging half of every weight tile (the pair halves the L2->SM weight traffic, which
// is what bounds the unfused level-0 GEMMs).  Roles per CTA: warp 0 TMA (X, W1), warp 1 MMA issue (leader CTA only),
// warp 2 TMA (W2), warps 3..10 gate / output epilogue (two warps per TMEM lane quarter).
// Issue order G1(j+1) before G2(j): the tensor pipe works on the next hidden chunk while the gate of chunk j runs.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kMlpThreads = 32 * 11;
constexpr int kMlpXChunkBytes = 128 * 64 * 2;   // one K chunk of the X tile: 128 rows x 64 channels (SWIZZLE_128B)
constexpr int kMlpW1SlotBytes = 64 * 64 * 2;    // 64 weight rows x 64 channels
constexpr int kMlpPBytes = 128 * 64 * 2;        // gated hidden chunk: 128 rows x 64
constexpr int kMlpW1Slots = 6;
constexpr int kMlpCtl = 1024;
constexpr uint32_t kMlpTmemCols = 512;
constexpr uint32_t kAcc2Col = 128;

struct MlpParams {
  int rows, C, hidden, k_chunks, n_chunks, num_tiles, w2_slots, w2_slot_bytes;
  const float* bias1;
  const float* bias2;
  float acc_scale, res1_scale, res2_scale;
  const bf16* res1;
  const bf16* res2;
  int res_ld;
  bf16* out;
  int out_ld;
  long long* trace;  // debug: clock64 stamps of CTA 0 (tools/mlp_trace.py); nullptr in production
};

struct alignas(64) MlpTmap {
  uint64_t opaque[16];
};

PT_DEVICE float mlp_gate(float value, float g) { return geglu_gate_tanh(value, g); }

// debug trace: slot = event id, up to 64 chunks per role
PT_DEVICE void mlp_stamp(long long* trace, int role, int ev, uint32_t chunk) {
  if (trace != nullptr && blockIdx.x == 0 && chunk < 64u && (threadIdx.x & 31) == 0) trace[(role * 8 + ev) * 64 + chunk] = clock64();
}

PT_DEVICE void unpack8(uint4 u, float* f) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

PT_DEVICE uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_geglu_kernel(const __grid_constant__ MlpTmap tmap_x, const __grid_constant__ MlpTmap tmap_w1,
                 const __grid_constant__ MlpTmap tmap_w2, const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  // barriers (identical offsets in both CTAs of the pair)
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem);  // leader: X tile of both CTAs landed
  uint64_t* x_empty = x_full + 1;                        // both: every GEMM1 of the tile retired
  uint64_t* w1_full = x_empty + 1;                       // [6] leader
  uint64_t* w1_empty = w1_full + kMlpW1Slots;            // [6] both
  uint64_t* w2_full = w1_empty + kMlpW1Slots;            // [4] leader
  uint64_t* w2_empty = w2_full + 4;                      // [4] both
  uint64_t* acc1_full = w2_empty + 4;                    // both: GEMM1(j) retired
  uint64_t* acc1_empty = acc1_full + 1;                  // leader: 16 warps hold acc1(j) in registers
  uint64_t* p_full = acc1_empty + 1;                     // [2] leader: 16 warps wrote P(j)
  uint64_t* p_empty = p_full + 2;                        // [2] both: GEMM2 reading P[b] retired
  uint64_t* acc2_full = p_empty + 2;                     // both: last GEMM2 of the tile retired
  uint64_t* acc2_empty = acc2_full + 1;                  // leader: 16 warps finished reading acc2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  volatile uint32_t* zero_word = reinterpret_cast<volatile uint32_t*>(smem + 512);  // always 0 (see the gate warps)
  uint8_t* sX = smem + kMlpCtl;
  uint8_t* sP = sX + (size_t)p.k_chunks * kMlpXChunkBytes;
  uint8_t* sW1 = sP + 2 * kMlpPBytes;
  uint8_t* sW2 = sW1 + kMlpW1Slots * kMlpW1SlotBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int pair_first = (int)cluster_id_x();
  const int pair_step = (int)num_clusters_x();
  const int quarter_rows = p.C >> 2;                       // W2 rows per box (each CTA stages 2 boxes per chunk)
  const uint32_t w2_half_bytes = (uint32_t)quarter_rows * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < kMlpW1Slots; ++s) {
      mbar_init(&w1_full[s], 1);
      mbar_init(&w1_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w2_full[s], 1);
      mbar_init(&w2_empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 16);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&p_full[b], 16);
      mbar_init(&p_empty[b], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    *zero_word = 0u;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, kMlpTmemCols);
  tc_fence_before();
  cluster_sync_all();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------ TMA: X tile + W1 ring (whole warp, one elected lane issues) -----------------
    {
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair_first; t < p.num_tiles; t += pair_step, ++it) {
        const int r0 = t * 256 + (int)cta_rank * 128;
        mbar_wait(x_empty, ((uint32_t)it & 1u) ^ 1u);
        const uint32_t xbar = map_to_cta(smem_u32(x_full), 0u);
        if (elect_one()) {
          if (cta_rank == 0) mbar_arrive_expect_tx(x_full, 2u * (uint32_t)p.k_chunks * kMlpXChunkBytes);
          for (int kc = 0; kc < p.k_chunks; ++kc)
            tma_load_2d_pair(sX + (size_t)kc * kMlpXChunkBytes, &tmap_x, xbar, kc * 64, r0);
        }
        __syncwarp();
        for (int j = 0; j < p.n_chunks; ++j) {
          // accumulator columns [value 64 | gate 64]: rank 0 stages the value rows, rank 1 the gate rows
          const int wrow = (cta_rank == 0 ? 0 : p.hidden) + j * 64;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&w1_empty[slot], phase ^ 1u);
            const uint32_t bar = map_to_cta(smem_u32(&w1_full[slot]), 0u);
            if (elect_one()) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&w1_full[slot], 2u * kMlpW1SlotBytes);
              tma_load_2d_pair(sW1 + (size_t)slot * kMlpW1SlotBytes, &tmap_w1, bar, kc * 64, wrow);
            }
            __syncwarp();
            if (++slot == kMlpW1Slots) {
              slot = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------ TMA: W2 ring (whole warp, one elected lane issues) --------------------------
    {
      int slot = 0;
      uint32_t phase = 0;
      for (int t = pair_first; t < p.num_tiles; t += pair_step) {
        for (int j = 0; j < p.n_chunks; ++j) {
          mbar_wait(&w2_empty[slot], phase ^ 1u);
          const uint32_t bar = map_to_cta(smem_u32(&w2_full[slot]), 0u);
          uint8_t* dst = sW2 + (size_t)slot * p.w2_slot_bytes;
          if (elect_one()) {
            if (cta_rank == 0) mbar_arrive_expect_tx(&w2_full[slot], 4u * w2_half_bytes);
            // MMA a covers output columns [0, C/2): rank r supplies W2 rows [r*C/4, +C/4); MMA b the upper half
            tma_load_2d_pair(dst, &tmap_w2, bar, j * 64, (int)cta_rank * quarter_rows);
            tma_load_2d_pair(dst + w2_half_bytes, &tmap_w2, bar, j * 64, (p.C >> 1) + (int)cta_rank * quarter_rows);
          }
          __syncwarp();
          if (++slot == p.w2_slots) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp of the leader CTA, one elected lane issues) ----------
    if (cta_rank == 0) {
      const uint32_t idesc1 = make_idesc_bf16(256, 128, 0, 0);
      const uint32_t idesc2 = make_idesc_bf16(256, (uint32_t)(p.C >> 1), 0, 0);
      const uint32_t acc1 = uniform_u32(tmem_base);
      const uint32_t acc2a = acc1 + kAcc2Col;
      const uint32_t acc2b = acc2a + (uint32_t)(p.C >> 1);
      int s1 = 0, s2 = 0;
      uint32_t ph1 = 0, ph2 = 0;
      uint32_t g1 = 0;   // global GEMM1 chunk counter
      uint32_t g2 = 0;   // global GEMM2 chunk counter
      int it = 0;
      auto issue_gemm2 = [&](int jj) {
        const uint32_t b = g2 & 1u;
        mlp_stamp(p.trace, 0, 3, g2);
        mbar_wait(&p_full[b], (g2 >> 1) & 1u);
        mlp_stamp(p.trace, 0, 4, g2);
        mbar_wait(&w2_full[s2], ph2);
        mlp_stamp(p.trace, 0, 5, g2);
        if (jj == 0) mbar_wait(acc2_empty, ((uint32_t)it & 1u) ^ 1u);   // the previous tile's output has been read
        tc_fence_after();
        const uint32_t sPb = smem_u32(sP + (size_t)b * kMlpPBytes);
        const uint32_t sW = smem_u32(sW2 + (size_t)s2 * p.w2_slot_bytes);
        const uint64_t pdesc = make_desc_kmajor_sw128(sPb);
        const uint64_t wa = make_desc_kmajor_sw128(sW);
        const uint64_t wb = make_desc_kmajor_sw128(sW + w2_half_bytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (jj | k) != 0 ? 1u : 0u;
            tc_mma_bf16_pair(acc2a, pdesc + (uint64_t)(2 * k), wa + (uint64_t)(2 * k), idesc2, accum);
            tc_mma_bf16_pair(acc2b, pdesc + (uint64_t)(2 * k), wb + (uint64_t)(2 * k), idesc2, accum);
          }
          tc_commit_pair(&w2_empty[s2], 3);
          tc_commit_pair(&p_empty[b], 3);
        }
        __syncwarp();
        mlp_stamp(p.trace, 0, 6, g2);
        if (++s2 == p.w2_slots) {
          s2 = 0;
          ph2 ^= 1u;
        }
        ++g2;
      };
      for (int t = pair_first; t < p.num_tiles; t += pair_step, ++it) {
        mbar_wait(x_full, (uint32_t)it & 1u);
        for (int j = 0; j < p.n_chunks; ++j) {
          mlp_stamp(p.trace, 0, 0, g1);
          mbar_wait(acc1_empty, (g1 & 1u) ^ 1u);   // the gate warps hold acc1 of the previous chunk in registers
          tc_fence_after();
          mlp_stamp(p.trace, 0, 1, g1);
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&w1_full[s1], ph1);
            tc_fence_after();
            const uint64_t adesc = make_desc_kmajor_sw128(smem_u32(sX + (size_t)kc * kMlpXChunkBytes));
            const uint64_t bdesc = make_desc_kmajor_sw128(smem_u32(sW1 + (size_t)s1 * kMlpW1SlotBytes));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_bf16_pair(acc1, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, (kc | k) != 0 ? 1u : 0u);
              tc_commit_pair(&w1_empty[s1], 3);
            }
            __syncwarp();
            if (++s1 == kMlpW1Slots) {
              s1 = 0;
              ph1 ^= 1u;
            }
          }
          if (elect_one()) {
            tc_commit_pair(acc1_full, 3);
            if (j == p.n_chunks - 1) tc_commit_pair(x_empty, 3);   // X may be overwritten once these retire
          }
          __syncwarp();
          mlp_stamp(p.trace, 0, 2, g1);
          ++g1;
          if (j > 0) issue_gemm2(j - 1);
        }
        issue_gemm2(p.n_chunks - 1);
        if (elect_one()) tc_commit_pair(acc2_full, 3);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ gate / output warps ------------------------------
    const int q = warp & 3;                 // TMEM lane quarter (hardware: warp id mod 4)
    const int hsel = (warp - 3) >> 2;       // which half of the columns
    const int row = q * 32 + lane;          // row inside this CTA's 128-row half == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc1_empty_l = map_to_cta(smem_u32(acc1_empty), 0u);
    const uint32_t acc2_empty_l = map_to_cta(smem_u32(acc2_empty), 0u);
    const uint32_t p_full_l = map_to_cta(smem_u32(&p_full[0]), 0u);
    uint32_t g = 0;
    int it = 0;
    for (int t = pair_first; t < p.num_tiles; t += pair_step, ++it) {
      for (int j = 0; j < p.n_chunks; ++j, ++g) {
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 0, g);
        mbar_wait(acc1_full, g & 1u);
        tc_fence_after();
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 1, g);
        uint32_t v[32], gt[32];
        tmem_ld_32x32(t_lane + (uint32_t)(hsel * 32), v);
        tmem_ld_32x32(t_lane + 64u + (uint32_t)(hsel * 32), gt);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(acc1_empty_l);
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 2, g);
        // The arrive has no result, so the instruction scheduler would sink it below the whole gate evaluation
        // (measured: "accumulator drained" was signalled ~1.5k cycles late and GEMM1 of the next chunk waited for
        // it).  A volatile shared-memory load issued after it in program order cannot pass it; folding the (always
        // zero) word into a few gate inputs puts the load — and with it the arrive — on the critical path.
        {
          const uint32_t z = *zero_word;
          v[0] ^= z; v[8] ^= z; v[16] ^= z; v[24] ^= z;
          gt[0] ^= z; gt[8] ^= z; gt[16] ^= z; gt[24] ^= z;
        }
        const int hcol = j * 64 + hsel * 32;
        const float4* bv = reinterpret_cast<const float4*>(p.bias1 + hcol);
        const float4* bg = reinterpret_cast<const float4*>(p.bias1 + p.hidden + hcol);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 x = __ldg(bv + (i >> 2)), y = __ldg(bg + (i >> 2));
          pk[i >> 1] = pack_bf16x2(mlp_gate(__uint_as_float(v[i]) + x.x, __uint_as_float(gt[i]) + y.x),
                                   mlp_gate(__uint_as_float(v[i + 1]) + x.y, __uint_as_float(gt[i + 1]) + y.y));
          pk[(i >> 1) + 1] = pack_bf16x2(mlp_gate(__uint_as_float(v[i + 2]) + x.z, __uint_as_float(gt[i + 2]) + y.z),
                                         mlp_gate(__uint_as_float(v[i + 3]) + x.w, __uint_as_float(gt[i + 3]) + y.w));
        }
        const uint32_t b = g & 1u;
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 3, g);
        mbar_wait(&p_empty[b], ((g >> 1) & 1u) ^ 1u);   // the GEMM2 that read this P buffer two chunks ago retired
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 4, g);
        uint8_t* prow = sP + (size_t)b * kMlpPBytes + (size_t)row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (hsel * 4 + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(p_full_l + b * 8u);
        if (warp == 3 && lane == 0) mlp_stamp(p.trace, 1, 5, g);
      }
      // ---- output of the tile: acc2 -> bias, scale, residuals -> bf16 -> global.  tcgen05.ld hands every thread one
      // ROW; each warp transposes its 32 x 32 chunk through an XOR-swizzled fp32 tile (the two P buffers are idle now:
      // 8 warps x 4 KB) so that a lane owns 8 consecutive columns of 4 rows and 4 neighbouring lanes cover 64
      // contiguous bytes of a row: residual loads and output stores are coalesced row segments (as in gemm.cu) ----
      mbar_wait(acc2_full, (uint32_t)it & 1u);
      tc_fence_after();
      const int sub_row = lane >> 2;
      const int seg = lane & 3;
      const uint32_t stage_u32 = smem_u32(sP + (size_t)(warp - 3) * 4096);
      const int row0 = t * 256 + (int)cta_rank * 128 + q * 32;
      const int half_cols = p.C >> 1;
      const int nch = half_cols >> 5;
      for (int c = 0; c < nch; ++c) {
        const int col = hsel * half_cols + c * 32 + seg * 8;
        uint32_t a[32];
        tmem_ld_32x32(t_lane + kAcc2Col + (uint32_t)(hsel * half_cols + c * 32), a);
        uint4 r1[4], r2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gr = row0 + i * 8 + sub_row;
          const bool ok = gr < p.rows;
          r1[i] = (ok && p.res1 != nullptr) ? ldg_nc_u4(p.res1 + (size_t)gr * p.res_ld + col) : make_uint4(0, 0, 0, 0);
          r2[i] = (ok && p.res2 != nullptr) ? ldg_nc_u4(p.res2 + (size_t)gr * p.res_ld + col) : make_uint4(0, 0, 0, 0);
        }
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + col) + 1);
        tmem_wait_ld();
        if (c == nch - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_relaxed(acc2_empty_l);
        }
        __syncwarp();  // the previous chunk's transposed reads are done
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t addr = stage_u32 + (uint32_t)lane * 128u + (uint32_t)((jj ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[4 * jj]), "r"(a[4 * jj + 1]),
                       "r"(a[4 * jj + 2]), "r"(a[4 * jj + 3]) : "memory");
        }
        __syncwarp();
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + sub_row;
          const uint32_t a0 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg) ^ (rr & 7)) << 4);
          const uint32_t a1 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg + 1) ^ (rr & 7)) << 4);
          float f[8];
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a0));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a1));
          const int gr = row0 + rr;
          if (gr >= p.rows) continue;
          float x1[8], x2[8];
          unpack8(r1[i], x1);
          unpack8(r2[i], x2);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float val = (f[e] + bb[e]) * p.acc_scale;
            val = fmaf(p.res1_scale, x1[e], val);
            f[e] = fmaf(p.res2_scale, x2[e], val);
          }
          stg_u4(p.out + (size_t)gr * p.out_ld + col, pack8(f));
        }
      }
      // the staging tiles alias the P buffers: no warp may start the next tile's gate (which writes P) before every
      // warp of this CTA has finished reading its transposed rows
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kMlpTmemCols);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_mlp_geglu(const PtMlpArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->tmap_x && a->tmap_w1 && a->tmap_w2 && a->out && a->bias1 && a->bias2, "pt_mlp_geglu: null argument");
  PT_CHECK_ARG(a->rows > 0 && a->C >= 64 && a->C <= 320 && (a->C % 64) == 0, "pt_mlp_geglu: C must be a multiple of 64 in [64, 320]");
  PT_CHECK_ARG(a->hidden == 4 * a->C, "pt_mlp_geglu: hidden must be 4*C");
  PT_CHECK_ARG((a->out_ld % 8) == 0 && (a->res_ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->res1) |
               reinterpret_cast<uintptr_t>(a->res2) | reinterpret_cast<uintptr_t>(a->bias1) | reinterpret_cast<uintptr_t>(a->bias2)) & 15u) == 0,
               "pt_mlp_geglu: operands must be 16-byte aligned with strides multiple of 8");
  MlpParams p;
  p.rows = a->rows;
  p.C = a->C;
  p.hidden = a->hidden;
  p.k_chunks = a->C / 64;
  p.n_chunks = a->hidden / 64;
  p.num_tiles = (a->rows + 255) / 256;
  p.w2_slot_bytes = 2 * (a->C / 4) * 128;
  const int fixed = kMlpCtl + p.k_chunks * kMlpXChunkBytes + 2 * kMlpPBytes + kMlpW1Slots * kMlpW1SlotBytes;
  int w2_slots = (226 * 1024 - 1024 - fixed) / p.w2_slot_bytes;
  if (w2_slots > 4) w2_slots = 4;
  PT_CHECK_ARG(w2_slots >= 2, "pt_mlp_geglu: shared memory budget");
  p.w2_slots = w2_slots;
  p.bias1 = a->bias1;
  p.bias2 = a->bias2;
  p.acc_scale = a->acc_scale;
  p.res1_scale = a->res1_scale;
  p.res2_scale = a->res2_scale;
  p.res1 = reinterpret_cast<const bf16*>(a->res1);
  p.res2 = reinterpret_cast<const bf16*>(a->res2);
  p.res_ld = a->res_ld;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  p.trace = reinterpret_cast<long long*>(a->trace);
  const size_t smem_bytes = (size_t)fixed + (size_t)w2_slots * p.w2_slot_bytes + 1024;
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_geglu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pt_fail(e, "pt_mlp_geglu: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  MlpTmap tx, t1, t2;
  memcpy(&tx, a->tmap_x, sizeof(tx));
  memcpy(&t1, a->tmap_w1, sizeof(t1));
  memcpy(&t2, a->tmap_w2, sizeof(t2));
  const int sms = pt_num_sms();
  const int pairs = p.num_tiles < sms / 2 ? p.num_tiles : sms / 2;
  cudaError_t e = pt_launch(mlp_geglu_kernel, dim3(2 * pairs), dim3(kMlpThreads), smem_bytes, stream, 2, tx, t1, t2, p);
  if (e != cudaSuccess) return pt_fail(e, "pt_mlp_geglu: cluster launch");
  return pt_launched("pt_mlp_geglu");
}
