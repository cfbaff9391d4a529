// posetraj_b200 — persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution (sm_100a).
//
//   D[m, n] = epilogue( sum_t sum_k A[m + shift_t, k] * Wt[n, t*K + k] )
//
// One kernel serves every dense contraction on the PoseTraj denoise path:
//   * Linear layers (proj_in/out, to_q/k/v, to_out, GEGLU feed-forwards)        num_taps = 1
//   * 3x3 spatial convs of ResnetBlock2D / Down/Upsample2D / conv_in / conv_out   num_taps = 9, the A rows
//     are the zero-haloed NHWC image space, so a tap is just a row shift and the halo supplies the padding
//   * temporal (3,1,1) convs of TemporalResnetBlock                               num_taps = 3, shift = +-H*W
//   * 1x1 shortcut / ControlNet zero-convs                                        num_taps = 1 (+ 2 K sources)
//
// Structure (per CTA, one CTA per SM, tiles 128 x block_n — or 256 x block_n shared by a CTA pair, cta_group::2 —
// K step 64):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> SWIZZLE_128B smem ring, mbarrier expect_tx)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma kind::f16, accumulators in TMEM, 2 accumulator stages)
//   warps 2..9/13 : epilogue       (tcgen05.ld -> smem transpose -> bias / row-vector / residual / blend / GEGLU ->
//                                   coalesced global stores; 8 warps, 12 for GEGLU)
// The three pipelines (smem full/empty, TMEM full/empty, static persistent tile schedule) follow the
// canonical Blackwell GEMM anatomy; block_n, stage count and tap table are runtime values, the epilogue flavour
// (generic / GEGLU / 8 lean variants) and the pair mode are template parameters.
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
// warp 0 TMA, warp 1 MMA, then the epilogue warps: 8 for the memory-bound epilogues, 12 for GEGLU whose 128 x block_n/2
// erf evaluations per tile are ALU work that more resident warps hide better (4 schedulers either way)
// kEpi selects the epilogue instantiation: 0 generic (every feature, runtime flags), 1 GEGLU, 2 + bits = lean
// variants for bf16 outputs with n_out % 8 == 0 and no SiLU (bit 0 per-row vector, bit 1 residual operands,
// bit 2 second output) — the compiler does not if-convert what is not instantiated.
constexpr int kEpiGeneric = 0, kEpiGeglu = 1, kEpiFast = 2;
// kEpi 10..17: the lean variants again, with SIXTEEN epilogue warps and no operand lookahead.  With <= 10 k-steps per
// tile the tile time is the epilogue's, and 8 warps (2 per scheduler) issue only ~15 % of the time there
// (profiles/r1j_gemm_full.md): four warps per scheduler hide the dependency latency instead of registers.
constexpr int kEpiWide = kEpiFast + 8;
__host__ __device__ constexpr int epi_warps(int epi) { return epi == kEpiGeglu ? 12 : (epi >= kEpiWide ? 16 : 8); }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + 32 * epi_warps(epi); }
__host__ __device__ constexpr int stage_warp_bytes(int epi) { return epi == kEpiGeglu ? 32 * 32 * 2 : 32 * 32 * 4; }
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr uint32_t kTmemCols = 512;             // 2 accumulator stages x 256 fp32 columns
constexpr int kMaxStages = 8;
constexpr int kSmemCtl = 3072;                  // barriers + tmem pointer (256 B) + staged bias (2 KiB)

struct GemmParams {
  int rows_per_batch, batches, n_out;
  int k0_chunks, k1_chunks, num_taps;
  int tap_shift[9];
  int block_n, geglu, gate_row_offset;
  int stages, stage_bytes;
  int tiles_per_batch, num_m_tiles, num_n_tiles;
  const float* bias;
  const float* rowvec;
  int rowvec_ld, rowvec_mode, rv_a, rv_b, rv_c, rv_mod, rv_off;
  float acc_scale;
  const float* acc_scale_ptr;  // optional device scalar multiplied into acc_scale (graph-replay safe)
  const bf16* res1;
  const bf16* res2;
  float res1_scale, res2_scale;
  int res_ld;
  void* out;
  int out_ld, out_dtype;
  void* out2;
  const bf16* aux;
  float aux_scale;
  int map_mode, pW1, pH1, ostride, oW, oH, out_halo, act_silu;
  // fused all-to-all (PtGemmArgs.scatter_mode): rows leave through NVLink peer memory
  int scatter_mode, sc_world, sc_J, sc_S, sc_kept_off, sc_kept_total;
  int sc_start[8], sc_count[8];
  bf16* sc_peer[8];
  int epi_depth;  // chunks of operand lookahead in the lean epilogue (1 or 2)
  long long* trace;  // debug (pt_gemm_set_trace): clock64 stamps of CTA 0's epilogue warp 2 / MMA warp, or nullptr
};

PT_DEVICE void gemm_stamp(long long* trace, int ev, int it) {
  if (trace != nullptr && blockIdx.x == 0 && it < 64 && (threadIdx.x & 31) == 0) trace[ev * 64 + it] = clock64();
}

struct alignas(64) TmapParam {
  uint64_t opaque[16];
};

// ---------------------------------------------------------------------------------------------------------
// Epilogue building blocks.
//
// tcgen05.ld hands every thread ONE accumulator row (32 consecutive fp32 columns).  Doing the epilogue in that
// layout makes every global access touch 32 different rows per instruction.  Instead each epilogue warp transposes
// its 32x32 chunk through a private, XOR-swizzled smem tile so that afterwards a lane owns 8 consecutive columns
// of a row and 4 neighbouring lanes cover 64 contiguous bytes: residual loads and output stores are coalesced
// row segments, and each lane only needs 16 bytes per operand in flight.
// ---------------------------------------------------------------------------------------------------------

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

// GEGLU gate: see geglu_gate_fast in common.cuh (value * g * Phi(g), Phi re-fitted to the exact erf GELU, 2.6e-5 abs).
PT_DEVICE float geglu_gate(float value, float g) { return geglu_gate_tanh(value, g); }

// activation codes of PtGemmArgs.act_silu: 1 SiLU, 2 GELU (exact erf), 3 quick-GELU x*sigmoid(1.702 x)
PT_DEVICE float apply_act(int act, float v) {
  if (act == 1) return silu_f(v);
  if (act == 2) return gelu_erf_f(v);
  return v * rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * v));
}

// slow path of the store side: fewer than 8 valid columns in this lane's segment (only conv_out: n_out = 4)
__device__ __noinline__ void epilogue_tail(const GemmParams& p, const float* acc, const float* sb_seg, int ncol,
                                           int nvalid, long long orow, int grp, float acc_scale) {
  for (int j = 0; j < nvalid; ++j) {
    float v = acc[j] + sb_seg[j];
    if (p.rowvec_mode != 0) v += __ldg(p.rowvec + (size_t)grp * p.rowvec_ld + ncol + j);
    v *= acc_scale;
    if (p.act_silu) v = apply_act(p.act_silu, v);
    if (p.res1 != nullptr) v = fmaf(p.res1_scale, __bfloat162float(p.res1[(size_t)orow * p.res_ld + ncol + j]), v);
    if (p.res2 != nullptr) v = fmaf(p.res2_scale, __bfloat162float(p.res2[(size_t)orow * p.res_ld + ncol + j]), v);
    const size_t off = (size_t)orow * p.out_ld + ncol + j;
    if (p.out_dtype == PT_DT_BF16) reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16(v);
    else reinterpret_cast<float*>(p.out)[off] = v;
    if (p.out2 != nullptr) {
      if (p.aux != nullptr) v = fmaf(p.aux_scale, __bfloat162float(p.aux[off]), v);
      if (p.out_dtype == PT_DT_BF16) reinterpret_cast<bf16*>(p.out2)[off] = __float2bfloat16(v);
      else reinterpret_cast<float*>(p.out2)[off] = v;
    }
  }
}

struct EpiRows {
  bf16* out_ptr[4];      // first element of the 4 output rows this lane finishes (local, or a peer's in scatter mode)
  long long out_off[4];  // their element offset in the local row space (aux / out2)
  long long res_off[4];
  int grp[4];
  uint32_t vmask;
};

// Lean epilogue of one accumulator tile for bf16 outputs (see kEpiFast).  val = acc_scale*(acc + bias + rowvec) +
// s1*res1 + s2*res2 ; out = val ; out2 = val + aux_scale*aux.
template <bool kRowvec, bool kRes, bool kOut2, int kStride, typename ArriveFn>
PT_DEVICE void epilogue_fast(const GemmParams& p, const EpiRows& R, uint32_t t_acc, uint32_t stage_u32, int n0,
                             int chunks, int hsel, int lane, float acc_scale, const ArriveFn& arrive_drained) {
  const int sub_row = lane >> 2;
  const int seg = lane & 3;
  bf16* out2 = reinterpret_cast<bf16*>(p.out2);
  const bool has_res2 = kRes && p.res2 != nullptr;
  const bool has_res1 = kRes && p.res1 != nullptr;
  // Operands (bias, residual, aux) are fetched one chunk AHEAD of the chunk that uses them, in rotating register
  // sets: the global-load latency of chunk c+2 overlaps the TMEM read, transpose and math of chunk c.  A second chunk
  // of lookahead (PT_EPI_DEPTH=2) measured no gain (profiles/r1e_experiments.md).  (res2 — rare — is fetched in the
  // chunk that uses it.)
  struct Pre {
    uint4 r1[4], ax[4];
    float4 b0, b1;
  };
  auto fetch = [&](int c, Pre& P) {
    const int ncol = n0 + c * 32 + seg * 8;
    P.b0 = make_float4(0.f, 0.f, 0.f, 0.f);
    P.b1 = P.b0;
    if (c < chunks && ncol < p.n_out) {
      if (p.bias != nullptr) {
        P.b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol));
        P.b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + 1);
      }
      if constexpr (kStride != 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if constexpr (kRes) P.r1[i] = has_res1 ? ldg_nc_u4(p.res1 + R.res_off[i] + ncol) : make_uint4(0, 0, 0, 0);
          if constexpr (kOut2) P.ax[i] = p.aux != nullptr ? ldg_nc_u4(p.aux + R.out_off[i] + ncol) : make_uint4(0, 0, 0, 0);
        }
      }
    }
  };
  // one 32-column chunk: TMEM -> registers -> swizzled smem transpose -> 4 coalesced row segments per lane
  auto process = [&](int c, const Pre& P) {
    const int ncol = n0 + c * 32 + seg * 8;
    const bool act = ncol < p.n_out;  // n_out % 8 == 0 on this path: a segment is full or empty
    uint32_t v[32];
    tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
    uint4 r2[4];
    if constexpr (kStride != 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (kRes) r2[i] = (act && has_res2) ? ldg_nc_u4(p.res2 + R.res_off[i] + ncol) : make_uint4(0, 0, 0, 0);
      }
    }
    tmem_wait_ld();
    if (c + kStride >= chunks) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_drained();
    }
    __syncwarp();  // the previous chunk's transposed reads are done
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = stage_u32 + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                   "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
    }
    __syncwarp();
    const float4 b0 = P.b0, b1 = P.b1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = i * 8 + sub_row;
      const uint32_t a0 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg) ^ (rr & 7)) << 4);
      const uint32_t a1 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg + 1) ^ (rr & 7)) << 4);
      float f[8];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a0));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a1));
      if (!act || !((R.vmask >> i) & 1u)) continue;
      // sixteen-warp flavour: operands are loaded where they are used (96 registers per thread, no lookahead sets)
      uint4 r1w = make_uint4(0, 0, 0, 0), r2w = r1w, axw = r1w;
      if constexpr (kStride == 4) {
        if constexpr (kRes) {
          if (has_res1) r1w = ldg_nc_u4(p.res1 + R.res_off[i] + ncol);
          if (has_res2) r2w = ldg_nc_u4(p.res2 + R.res_off[i] + ncol);
        }
        if constexpr (kOut2) axw = p.aux != nullptr ? ldg_nc_u4(p.aux + R.out_off[i] + ncol) : make_uint4(0, 0, 0, 0);
      }
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
      f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      if constexpr (kRowvec) {
        const float4* rv = reinterpret_cast<const float4*>(p.rowvec + (size_t)R.grp[i] * p.rowvec_ld + ncol);
        const float4 v0 = __ldg(rv), v1 = __ldg(rv + 1);
        f[0] += v0.x; f[1] += v0.y; f[2] += v0.z; f[3] += v0.w;
        f[4] += v1.x; f[5] += v1.y; f[6] += v1.z; f[7] += v1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= acc_scale;
      if constexpr (kRes) {
        float r[8];
        unpack8(kStride == 4 ? r1w : P.r1[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(p.res1_scale, r[j], f[j]);
        if (has_res2) {
          unpack8(kStride == 4 ? r2w : r2[i], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(p.res2_scale, r[j], f[j]);
        }
      }
      stg_u4(R.out_ptr[i] + ncol, pack8(f));
      if constexpr (kOut2) {
        float r[8];
        unpack8(kStride == 4 ? axw : P.ax[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(p.aux_scale, r[j], f[j]);
        stg_u4(out2 + R.out_off[i] + ncol, pack8(f));
      }
    }
  };
  Pre P0, P1;
  if constexpr (kStride == 4) {
    // sixteen epilogue warps: no lookahead registers, the other three warps of the scheduler cover the load latency
    for (int c = hsel; c < chunks; c += 4) {
      fetch(c, P0);
      process(c, P0);
    }
    return;
  }
  fetch(hsel, P0);
  // optional second chunk of lookahead (three register sets; not for the second-output variants, which already
  // carry 16 more registers per set)
  if constexpr ((kRes || kOut2) && !(kRes && kOut2)) {
    if (p.epi_depth >= 2) {
      Pre P2;
      fetch(hsel + 2, P1);
      for (int c = hsel; c < chunks; c += 6) {
        fetch(c + 4, P2);
        process(c, P0);
        if (c + 2 < chunks) {
          fetch(c + 6, P0);
          process(c + 2, P1);
        }
        if (c + 4 < chunks) {
          fetch(c + 8, P1);
          process(c + 4, P2);
        }
      }
      return;
    }
  }
  for (int c = hsel; c < chunks; c += 4) {
    fetch(c + 2, P1);
    process(c, P0);
    if (c + 2 < chunks) {
      fetch(c + 4, P0);
      process(c + 2, P1);
    }
  }
}

// kPair: the kernel runs as clusters of two CTAs (one SM pair) that share one 256 x block_n tile through
// tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of A and HALF of the B tile, so a k-step costs
// 16 KiB + block_n*64 B of L2->SM traffic per SM instead of 16 KiB + block_n*128 B: the large-M layers are
// L2-bandwidth bound with single-CTA tiles (profiles/r1b).
template <int kEpi, bool kPair>
__global__ void __launch_bounds__(gemm_threads(kEpi), 1)
gemm_tcgen05_kernel(const __grid_constant__ TmapParam tmap_a0, const __grid_constant__ TmapParam tmap_a1,
                    const __grid_constant__ TmapParam tmap_b, const __grid_constant__ GemmParams p) {
  constexpr bool kGeglu = kEpi == kEpiGeglu;
  constexpr int kEpiWarps = epi_warps(kEpi);
  constexpr int kStageWarpBytes = stage_warp_bytes(kEpi);
  constexpr int kChunkStride = kEpiWarps / 4;  // epilogue warps per TMEM lane quarter
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;             // [kMaxStages]
  uint64_t* tfull_bar = empty_bar + kMaxStages;            // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(smem + 256);     // [2][256] bias of the tile, per accumulator stage
  uint8_t* stage_base = smem + kSmemCtl;                   // kEpiWarps x kStageWarpBytes transpose tiles
  uint8_t* tiles = stage_base + ((kEpiWarps * kStageWarpBytes + 1023) & ~1023);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int k_chunks = p.k0_chunks + p.k1_chunks;
  const int k_iters = p.num_taps * k_chunks;
  const int half = p.block_n >> 1;
  const uint32_t b_half_bytes = (uint32_t)half * kBlockK * 2;
  // pair mode: cta_rank 0 is the leader (issues the MMAs, owns the full / tmem-empty barriers both CTAs signal)
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const int tile_first = kPair ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = kPair ? (int)num_clusters_x() : (int)gridDim.x;
  constexpr int kTileM = kPair ? 2 * kBlockM : kBlockM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    if (p.k1_chunks > 0) tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair(tmem_ptr, kTmemCols);
    else tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();  // the peer's barriers must be initialised before anything remote
  __syncthreads();  // (also in pair mode: the CTA-level barrier is what race checkers track for the smem handshake)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // PDL: barrier init, TMEM allocation and descriptor prefetch above overlap the tail of the previous kernel;
  // operands, residuals and output buffers are only touched after its completion
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer (whole warp, one elected lane issues) ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile_first; t < num_tiles; t += tile_step) {
        const int m_tile = t / p.num_n_tiles;
        const int n_tile = t - m_tile * p.num_n_tiles;
        const int batch = m_tile / p.tiles_per_batch;
        const int r0 = (m_tile - batch * p.tiles_per_batch) * kTileM + (int)cta_rank * kBlockM;
        const int n0 = kGeglu ? n_tile * half : n_tile * p.block_n;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int arow = r0 + p.tap_shift[tap];
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sA = tiles + (size_t)stage * p.stage_bytes;
            uint8_t* sB = sA + kABytes;
            const int kcol = (tap * k_chunks + kc) * kBlockK;
            const int ak = (kc < p.k0_chunks ? kc : kc - p.k0_chunks) * kBlockK;
            const TmapParam* ta = kc < p.k0_chunks ? &tmap_a0 : &tmap_a1;
            if constexpr (kPair) {
              // both CTAs credit the LEADER's full barrier; the leader expects the bytes of the whole pair
              const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0u);
              // B rows of this CTA: the second half of the tile's N range for rank 1 (GEGLU: rank 0 = value rows,
              // rank 1 = gate rows, which is exactly the accumulator column order [value | gate])
              const int brow = kGeglu ? (cta_rank == 0 ? n0 : p.gate_row_offset + n0) : n0 + (int)cta_rank * half;
              if (elect_one()) {
                if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * ((uint32_t)kABytes + b_half_bytes));
                tma_load_3d_pair(sA, ta, bar, ak, arow, batch);
                tma_load_2d_pair(sB, &tmap_b, bar, kcol, brow);
              }
            } else {
              const int brow1 = kGeglu ? p.gate_row_offset + n0 : n0 + half;
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kABytes + 2u * b_half_bytes);
                tma_load_3d(sA, ta, &full_bar[stage], ak, arow, batch);
                tma_load_2d(sB, &tmap_b, &full_bar[stage], kcol, n0);
                tma_load_2d(sB + b_half_bytes, &tmap_b, &full_bar[stage], kcol, brow1);
              }
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp of the leader CTA, one elected lane issues) ----------
    if (cta_rank == 0) {
      const uint32_t tmem_u = uniform_u32(tmem_base);
      const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = tile_first; t < num_tiles; t += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        gemm_stamp(p.trace, 0, it);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        gemm_stamp(p.trace, 1, it);
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * 256u;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(tiles + (size_t)stage * p.stage_bytes);
          const uint64_t adesc = make_desc_kmajor_sw128(sA);
          const uint64_t bdesc = make_desc_kmajor_sw128(sA + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // +32 bytes (16 bf16) along K inside the 128-byte swizzle row: +2 in the >>4 address field
              if constexpr (kPair)
                tc_mma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) != 0 ? 1u : 0u);
              else
                tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if constexpr (kPair) tc_commit_pair(&empty_bar[stage], 3);
            else tc_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        gemm_stamp(p.trace, 2, it);
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one()) {
          if constexpr (kPair) tc_commit_pair(&tfull_bar[acc], 3);
          else tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue warps ----------------------------
    // 8 warps: warp % 4 selects the TMEM lane quarter (hardware restriction), (warp - 2) / 4 the chunk parity.
    const int q = warp & 3;
    const int ew = warp - 2;
    const int hsel = ew >> 2;
    const int etid = threadIdx.x - 64;  // 0..255
    const int sub_row = lane >> 2;      // row inside a group of 8 after the transpose
    const int seg = lane & 3;           // which 8 columns of the 32-column chunk
    uint8_t* stage = stage_base + ew * kStageWarpBytes;
    const uint32_t stage_u32 = smem_u32(stage);
    // conditioning_scale lives in device memory so that a captured graph picks up a new value at replay
    const float acc_scale = p.acc_scale_ptr != nullptr ? p.acc_scale * __ldg(p.acc_scale_ptr) : p.acc_scale;
    // "accumulator drained" goes to the LEADER's barrier (remote arrive for rank 1 of a pair)
    const uint32_t tempty_addr0 = kPair ? map_to_cta(smem_u32(&tempty_bar[0]), 0u) : smem_u32(&tempty_bar[0]);
    auto arrive_tempty = [&](int acc) {
      if constexpr (kPair) mbar_arrive_cluster(tempty_addr0 + (uint32_t)acc * 8u);
      else mbar_arrive(&tempty_bar[acc]);
    };
    int it = 0;
    for (int t = tile_first; t < num_tiles; t += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m_tile = t / p.num_n_tiles;
      const int n_tile = t - m_tile * p.num_n_tiles;
      const int batch = m_tile / p.tiles_per_batch;
      const int row_base = (m_tile - batch * p.tiles_per_batch) * kTileM + (int)cta_rank * kBlockM + q * 32;
      // the 4 rows this lane finishes (after the transpose): row_base + i*8 + sub_row
      int orow4[4], grp4[4];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row_base + i * 8 + sub_row;
        bool valid = r < p.rows_per_batch;
        long long orow = (long long)batch * p.rows_per_batch + r;
        if (p.map_mode == 1) {
          const int per_img = p.pW1 * p.pH1;
          const int img = r / per_img;
          const int rem = r - img * per_img;
          const int y = rem / p.pW1;
          const int x = rem - y * p.pW1;
          valid = valid && (y < p.pH1 - 1) && (x < p.pW1 - 1) && (y % p.ostride == 0) && (x % p.ostride == 0);
          const long long oimg = (long long)batch * (p.rows_per_batch / per_img) + img;
          if (p.out_halo)
            orow = (oimg * (p.oH + 1) + y / p.ostride) * (p.oW + 1) + x / p.ostride;
          else
            orow = (oimg * p.oH + y / p.ostride) * p.oW + x / p.ostride;
        }
        int grp = 0;
        if (p.rowvec_mode == 1) {
          grp = (int)(orow / p.rv_a);
        } else if (p.rowvec_mode == 2) {
          grp = (int)(((orow / p.rv_a) * p.rv_b + orow % p.rv_mod + p.rv_off) % p.rv_c);
        }
        if (!valid) { orow = 0; grp = 0; }
        orow4[i] = (int)orow;
        grp4[i] = grp;
        vmask |= (valid ? 1u : 0u) << i;
      }
      if constexpr (kEpi >= kEpiFast) {
        EpiRows R;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          R.out_off[i] = (long long)orow4[i] * p.out_ld;
          R.res_off[i] = (long long)orow4[i] * p.res_ld;
          R.grp[i] = grp4[i];
          if (p.scatter_mode == 0) {
            R.out_ptr[i] = reinterpret_cast<bf16*>(p.out) + R.out_off[i];
          } else {
            // (b, j, s) of the local row, owner q of its split-axis index, row in q's layout (see PtGemmArgs)
            const int per_b = p.sc_J * p.sc_S;
            const int b = orow4[i] / per_b;
            const int rem = orow4[i] - b * per_b;
            const int j = rem / p.sc_S;
            const int s = rem - j * p.sc_S;
            const int u = p.scatter_mode == 1 ? s : j;
            int q = 0;
            while (q + 1 < p.sc_world && u >= p.sc_start[q] + p.sc_count[q]) ++q;
            const long long drow = p.scatter_mode == 1
                ? ((long long)b * p.sc_kept_total + p.sc_kept_off + j) * p.sc_count[q] + (s - p.sc_start[q])
                : ((long long)b * p.sc_count[q] + (j - p.sc_start[q])) * p.sc_kept_total + p.sc_kept_off + s;
            R.out_ptr[i] = p.sc_peer[q] + drow * p.out_ld;
          }
        }
        R.vmask = vmask;
        if (warp == 2) gemm_stamp(p.trace, 3, it);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 2) gemm_stamp(p.trace, 4, it);
        const uint32_t t_acc_f = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
        const int chunks_f = p.block_n >> 5;
        if (hsel >= chunks_f) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty(acc);
        }
        constexpr int kBits = (kEpi - kEpiFast) & 7;
        epilogue_fast<(kBits & 1) != 0, (kBits & 2) != 0, (kBits & 4) != 0, kChunkStride>(
            p, R, t_acc_f, stage_u32, n_tile * p.block_n, chunks_f, hsel, lane, acc_scale, [&]() { arrive_tempty(acc); });
        if (warp == 2) gemm_stamp(p.trace, 5, it);
        continue;
      }

      long long gout_off[4];
      bf16* gout = reinterpret_cast<bf16*>(p.out);
#pragma unroll
      for (int i = 0; i < 4; ++i) gout_off[i] = (long long)orow4[i] * p.out_ld;
      // stage this tile's bias in smem, indexed like the accumulator columns
      const int n0 = kGeglu ? n_tile * half : n_tile * p.block_n;
      float* sb = sbias + acc * 256;
      if (etid < p.block_n) {
        float bv = 0.f;
        if (p.bias != nullptr) {
          if constexpr (kGeglu) {
            const int nn = n0 + (etid < half ? etid : etid - half);
            if (nn < p.n_out) bv = __ldg(p.bias + (etid < half ? nn : p.gate_row_offset + nn));
          } else if (n0 + etid < p.n_out) {
            bv = __ldg(p.bias + n0 + etid);
          }
        }
        sb[etid] = bv;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      const int chunks = (kGeglu ? half : p.block_n) >> 5;
      if (hsel >= chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty(acc);
      }
      for (int c = hsel; c < chunks; c += kChunkStride) {
        const int ncol = n0 + c * 32 + seg * 8;      // first of this lane's 8 output columns
        const int nvalid = min(8, p.n_out - ncol);   // <= 0: nothing to store
        if constexpr (kGeglu) {
          // GEGLU tiles carry bias only (host-checked): out = (x + bx) * gelu(g + bg), computed row-per-thread,
          // then transposed as bf16 (64-byte rows, 16-byte chunks XOR-swizzled by (row >> 1) & 3)
          uint32_t v[32];
          uint32_t g[32];
          tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
          tmem_ld_32x32(t_acc + (uint32_t)(half + c * 32), g);
          tmem_wait_ld();
          if (c + kChunkStride >= chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc);
          }
          uint32_t pk[16];
          const float4* sbv = reinterpret_cast<const float4*>(sb + c * 32);
          const float4* sbg = reinterpret_cast<const float4*>(sb + half + c * 32);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bx = sbv[j >> 2], bg = sbg[j >> 2];  // broadcast 128-bit smem reads
            const float x0 = __uint_as_float(v[j]) + bx.x, x1 = __uint_as_float(v[j + 1]) + bx.y;
            const float x2 = __uint_as_float(v[j + 2]) + bx.z, x3 = __uint_as_float(v[j + 3]) + bx.w;
            const float g0 = __uint_as_float(g[j]) + bg.x, g1 = __uint_as_float(g[j + 1]) + bg.y;
            const float g2 = __uint_as_float(g[j + 2]) + bg.z, g3 = __uint_as_float(g[j + 3]) + bg.w;
            pk[j >> 1] = pack_bf16x2(geglu_gate(x0, g0), geglu_gate(x1, g1));
            pk[(j >> 1) + 1] = pack_bf16x2(geglu_gate(x2, g2), geglu_gate(x3, g3));
          }
          __syncwarp();  // the previous chunk's transposed reads are done
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = stage_u32 + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                         "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + sub_row;
            const uint32_t addr = stage_u32 + (uint32_t)rr * 64u + (uint32_t)((seg ^ ((rr >> 1) & 3)) << 4);
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
            // GEGLU outputs are bf16 with n_out % 8 == 0 (host-checked): a segment is full or empty
            if (((vmask >> i) & 1u) && nvalid > 0) stg_u4(gout + gout_off[i] + ncol, u);
          }
        } else {
          uint32_t v[32];
          tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
          // operands of this lane's 4 row segments, fetched before the TMEM wait (coalesced: 4 lanes = 64 B of a row)
          const bool full = nvalid == 8;
          uint4 r1[4], r2[4], ax[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = full && ((vmask >> i) & 1u);
            r1[i] = (ok && p.res1 != nullptr) ? ldg_u4(p.res1 + (size_t)orow4[i] * p.res_ld + ncol) : make_uint4(0, 0, 0, 0);
            r2[i] = (ok && p.res2 != nullptr) ? ldg_u4(p.res2 + (size_t)orow4[i] * p.res_ld + ncol) : make_uint4(0, 0, 0, 0);
            ax[i] = (ok && p.out2 != nullptr && p.aux != nullptr) ? ldg_u4(p.aux + (size_t)orow4[i] * p.out_ld + ncol) : make_uint4(0, 0, 0, 0);
          }
          tmem_wait_ld();
          if (c + kChunkStride >= chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc);
          }
          __syncwarp();  // the previous chunk's transposed reads are done
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = stage_u32 + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                         "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
          }
          __syncwarp();
          const float* sb_seg = sb + c * 32 + seg * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(sb_seg);
          const float4 b1 = *reinterpret_cast<const float4*>(sb_seg + 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + sub_row;
            const uint32_t a0 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg) ^ (rr & 7)) << 4);
            const uint32_t a1 = stage_u32 + (uint32_t)rr * 128u + (uint32_t)(((2 * seg + 1) ^ (rr & 7)) << 4);
            float f[8];
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a0));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a1));
            if (!((vmask >> i) & 1u) || nvalid <= 0) continue;
            if (!full) {
              float tmp[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tmp[j] = f[j];
              epilogue_tail(p, tmp, sb_seg, ncol, nvalid, orow4[i], grp4[i], acc_scale);
              continue;
            }
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            if (p.rowvec_mode != 0) {
              const float4* rv = reinterpret_cast<const float4*>(p.rowvec + (size_t)grp4[i] * p.rowvec_ld + ncol);
              const float4 v0 = __ldg(rv), v1 = __ldg(rv + 1);
              f[0] += v0.x; f[1] += v0.y; f[2] += v0.z; f[3] += v0.w;
              f[4] += v1.x; f[5] += v1.y; f[6] += v1.z; f[7] += v1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] *= acc_scale;
            if (p.act_silu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = apply_act(p.act_silu, f[j]);
            }
            if (p.res1 != nullptr) {
              float r[8];
              unpack8(r1[i], r);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(p.res1_scale, r[j], f[j]);
            }
            if (p.res2 != nullptr) {
              float r[8];
              unpack8(r2[i], r);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(p.res2_scale, r[j], f[j]);
            }
            const size_t off = (size_t)orow4[i] * p.out_ld + ncol;
            if (p.out_dtype == PT_DT_BF16) {
              stg_u4(reinterpret_cast<bf16*>(p.out) + off, pack8(f));
            } else {
              float* o = reinterpret_cast<float*>(p.out) + off;
              *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
            }
            if (p.out2 != nullptr) {
              float r[8];
              unpack8(ax[i], r);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(p.aux_scale, r[j], f[j]);
              if (p.out_dtype == PT_DT_BF16) {
                stg_u4(reinterpret_cast<bf16*>(p.out2) + off, pack8(f));
              } else {
                float* o = reinterpret_cast<float*>(p.out2) + off;
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              }
            }
          }
        }
      }
    }
  }

  // ------------------------------ teardown -------------------------------------
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();  // the peer may still read this CTA's smem / signal its barriers
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace pt

using namespace pt;

static long long* g_gemm_trace = nullptr;
// debug hook (tools/gemm_trace.py): int64[8*64] device buffer receiving clock64 stamps of CTA 0, or NULL to switch off
extern "C" void pt_gemm_set_trace(void* buf) { g_gemm_trace = reinterpret_cast<long long*>(buf); }

extern "C" int pt_gemm(const PtGemmArgs* a, void* stream) {
  if (a == nullptr || a->tmap_a0 == nullptr || a->tmap_b == nullptr || a->out == nullptr)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: null argument");
  if (a->block_n < 32 || a->block_n > 256 || (a->block_n % 32) != 0)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: block_n must be a multiple of 32 in [32,256]");
  if (a->cta_pair && a->block_n < 64)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: cta_pair needs block_n >= 64");
  if (a->geglu && (a->block_n % 64) != 0)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: GEGLU needs block_n multiple of 64");
  if (a->num_taps < 1 || a->num_taps > 9 || a->k0_chunks < 1 || a->k1_chunks < 0)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: bad tap / K configuration");
  if (a->k1_chunks > 0 && a->tmap_a1 == nullptr)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: k1_chunks > 0 without tmap_a1");
  if (a->rows_per_batch < 1 || a->batches < 1 || a->n_out < 1)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: empty problem");
  if (a->map_mode == 1 && (a->pW1 < 2 || a->pH1 < 2 || a->ostride < 1 || (a->rows_per_batch % (a->pW1 * a->pH1)) != 0))
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: bad haloed-image mapping");
  if (a->rowvec_mode != 0 && (a->rowvec == nullptr || a->rv_a < 1))
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: rowvec_mode without rowvec");
  if (a->geglu && (a->rowvec_mode != 0 || a->res1 != nullptr || a->res2 != nullptr || a->out2 != nullptr || a->acc_scale != 1.0f || a->acc_scale_ptr != nullptr))
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: GEGLU tiles support a bias-only epilogue");
  if (a->scatter_mode != 0) {
    if (a->scatter_mode < 0 || a->scatter_mode > 2 || a->sc_world < 1 || a->sc_world > 8 || a->sc_J < 1 || a->sc_S < 1 ||
        a->sc_kept_total < 1 || a->sc_kept_off < 0)
      return pt_fail(cudaErrorInvalidValue, "pt_gemm: bad scatter geometry");
    for (int q = 0; q < a->sc_world; ++q)
      if (a->sc_peer[q] == nullptr || a->sc_count[q] < 1 || (reinterpret_cast<uintptr_t>(a->sc_peer[q]) & 15u) != 0)
        return pt_fail(cudaErrorInvalidValue, "pt_gemm: scatter needs a 16-byte aligned peer buffer and >= 1 item per rank");
    if (a->geglu || a->out2 != nullptr || a->out_dtype != PT_DT_BF16 || (a->n_out % 8) != 0 || a->act_silu || a->out_halo)
      return pt_fail(cudaErrorInvalidValue, "pt_gemm: scatter supports plain bf16 outputs with n_out % 8 == 0 only");
  }
  if (a->geglu && (a->out_dtype != PT_DT_BF16 || (a->n_out % 8) != 0 || (a->out_ld % 8) != 0 ||
                   (reinterpret_cast<uintptr_t>(a->out) & 15u) != 0))
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: GEGLU output must be bf16, 16-byte aligned, with n_out and out_ld multiples of 8");

  GemmParams p;
  p.rows_per_batch = a->rows_per_batch;
  p.batches = a->batches;
  p.n_out = a->n_out;
  p.k0_chunks = a->k0_chunks;
  p.k1_chunks = a->k1_chunks;
  p.num_taps = a->num_taps;
  for (int i = 0; i < 9; ++i) p.tap_shift[i] = a->tap_shift[i];
  p.block_n = a->block_n;
  p.geglu = a->geglu ? 1 : 0;
  p.gate_row_offset = a->gate_row_offset;
  p.stage_bytes = kABytes + (a->cta_pair ? a->block_n / 2 : a->block_n) * kBlockK * 2;
  // epilogue flavour (needed for the shared-memory budget below)
  int epi = kEpiGeneric;
  if (a->geglu) {
    epi = kEpiGeglu;
  } else if (a->out_dtype == PT_DT_BF16 && (a->n_out % 8) == 0 && !a->act_silu && (a->out_ld % 8) == 0 &&
             (a->res_ld % 8) == 0 && (a->rowvec_mode == 0 || (a->rowvec_ld % 4) == 0) &&
             ((reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->out2) | reinterpret_cast<uintptr_t>(a->res1) |
               reinterpret_cast<uintptr_t>(a->res2) | reinterpret_cast<uintptr_t>(a->aux) | reinterpret_cast<uintptr_t>(a->bias) |
               reinterpret_cast<uintptr_t>(a->rowvec)) & 15u) == 0) {
    epi = kEpiFast + (a->rowvec_mode != 0 ? 1 : 0) + ((a->res1 != nullptr || a->res2 != nullptr) ? 2 : 0) +
          (a->out2 != nullptr ? 4 : 0);
    static int env_wide = -2;   // PT_EPI16: max k-steps per tile for the 16-warp epilogue (0 disables; default 0)
    if (env_wide == -2) {
      const char* e = getenv("PT_EPI16");
      env_wide = e ? atoi(e) : 0;
    }
    if (a->num_taps * (a->k0_chunks + a->k1_chunks) <= env_wide) epi += 8;
  }
  const int stage_area = (epi_warps(epi) * stage_warp_bytes(epi) + 1023) & ~1023;
  const int smem_limit = 227 * 1024 - kSmemCtl - 1024 - stage_area;
  int stages = smem_limit / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  {
    static int env_stages = -2;   // experiment knob: PT_GEMM_STAGES caps the ring depth
    if (env_stages == -2) {
      const char* e = getenv("PT_GEMM_STAGES");
      env_stages = e ? atoi(e) : -1;
    }
    if (env_stages > 0 && stages > env_stages) stages = env_stages;
  }
  p.stages = stages;
  const bool pair = a->cta_pair != 0;
  const int tile_m = pair ? 2 * kBlockM : kBlockM;
  p.tiles_per_batch = (a->rows_per_batch + tile_m - 1) / tile_m;
  p.num_m_tiles = p.tiles_per_batch * a->batches;
  const int n_per_tile = a->geglu ? a->block_n / 2 : a->block_n;
  p.num_n_tiles = (a->n_out + n_per_tile - 1) / n_per_tile;
  p.bias = a->bias;
  p.rowvec = a->rowvec;
  p.rowvec_ld = a->rowvec_ld;
  p.rowvec_mode = a->rowvec_mode;
  p.rv_a = a->rv_a > 0 ? a->rv_a : 1;
  p.rv_b = a->rv_b > 0 ? a->rv_b : 1;
  p.rv_c = a->rv_c > 0 ? a->rv_c : 1;
  p.rv_mod = a->rv_mod > 0 ? a->rv_mod : p.rv_b;
  p.rv_off = a->rv_off > 0 ? a->rv_off : 0;
  p.acc_scale = a->acc_scale;
  p.acc_scale_ptr = a->acc_scale_ptr;
  p.res1 = reinterpret_cast<const bf16*>(a->res1);
  p.res2 = reinterpret_cast<const bf16*>(a->res2);
  p.res1_scale = a->res1_scale;
  p.res2_scale = a->res2_scale;
  p.res_ld = a->res_ld;
  p.out = a->out;
  p.out_ld = a->out_ld;
  p.out_dtype = a->out_dtype;
  p.out2 = a->out2;
  p.aux = reinterpret_cast<const bf16*>(a->aux);
  p.aux_scale = a->aux_scale;
  p.map_mode = a->map_mode;
  p.pW1 = a->pW1 > 0 ? a->pW1 : 1;
  p.pH1 = a->pH1 > 0 ? a->pH1 : 1;
  p.ostride = a->ostride > 0 ? a->ostride : 1;
  p.oW = a->oW;
  p.oH = a->oH;
  p.out_halo = a->out_halo;
  p.act_silu = a->act_silu;
  {
    static int env_depth = -2;   // experiment knob: PT_EPI_DEPTH=2 = two chunks of operand lookahead (no gain: r1e)
    if (env_depth == -2) {
      const char* e = getenv("PT_EPI_DEPTH");
      env_depth = e ? atoi(e) : -1;
    }
    p.epi_depth = env_depth > 0 ? env_depth : 1;
  }
  p.trace = g_gemm_trace;
  p.scatter_mode = a->scatter_mode;
  p.sc_world = a->sc_world; p.sc_J = a->sc_J > 0 ? a->sc_J : 1; p.sc_S = a->sc_S > 0 ? a->sc_S : 1;
  p.sc_kept_off = a->sc_kept_off; p.sc_kept_total = a->sc_kept_total;
  for (int q = 0; q < 8; ++q) {
    p.sc_start[q] = a->sc_start[q];
    p.sc_count[q] = a->sc_count[q];
    p.sc_peer[q] = reinterpret_cast<bf16*>(a->sc_peer[q]);
  }

  const size_t smem_bytes = (size_t)kSmemCtl + (size_t)stage_area + (size_t)p.stages * p.stage_bytes + 1024;
  typedef void (*KernelFn)(TmapParam, TmapParam, TmapParam, GemmParams);
#define PT_GEMM_ROW(P)                                                                                          \
  {gemm_tcgen05_kernel<0, P>,  gemm_tcgen05_kernel<1, P>,  gemm_tcgen05_kernel<2, P>,  gemm_tcgen05_kernel<3, P>,  \
   gemm_tcgen05_kernel<4, P>,  gemm_tcgen05_kernel<5, P>,  gemm_tcgen05_kernel<6, P>,  gemm_tcgen05_kernel<7, P>,  \
   gemm_tcgen05_kernel<8, P>,  gemm_tcgen05_kernel<9, P>,  gemm_tcgen05_kernel<10, P>, gemm_tcgen05_kernel<11, P>, \
   gemm_tcgen05_kernel<12, P>, gemm_tcgen05_kernel<13, P>, gemm_tcgen05_kernel<14, P>, gemm_tcgen05_kernel<15, P>, \
   gemm_tcgen05_kernel<16, P>, gemm_tcgen05_kernel<17, P>}
  constexpr int kNumEpi = 18;
  static const KernelFn kernels[2][kNumEpi] = {PT_GEMM_ROW(false), PT_GEMM_ROW(true)};
#undef PT_GEMM_ROW
  static bool attr_set[PT_MAX_DEVICES] = {false};  // cudaFuncSetAttribute is per device
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    for (int m = 0; m < 2; ++m)
      for (int i = 0; i < kNumEpi; ++i) {
        cudaError_t e = cudaFuncSetAttribute(kernels[m][i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return pt_fail(e, "pt_gemm: cudaFuncSetAttribute");
      }
    attr_set[dev_slot] = true;
  }
  if (p.scatter_mode != 0 && epi < kEpiFast)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: scatter needs 16-byte aligned operands and strides (lean epilogue)");
  const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  const int sms = pt_num_sms();

  TmapParam ta0, ta1, tb;
  memcpy(&ta0, a->tmap_a0, sizeof(TmapParam));
  memcpy(&ta1, a->tmap_a1 ? a->tmap_a1 : a->tmap_a0, sizeof(TmapParam));
  memcpy(&tb, a->tmap_b, sizeof(TmapParam));
  if (pair) {
    const int clusters = (int)(tiles < sms / 2 ? tiles : sms / 2);
    cudaError_t e = pt_launch(kernels[1][epi], dim3(2 * clusters), dim3(gemm_threads(epi)), smem_bytes, stream, 2, ta0, ta1, tb, p);
    if (e != cudaSuccess) return pt_fail(e, "pt_gemm: cluster launch");
  } else {
    const int grid = (int)(tiles < sms ? tiles : sms);
    pt_launch(kernels[0][epi], dim3(grid), dim3(gemm_threads(epi)), smem_bytes, stream, 1, ta0, ta1, tb, p);
  }
  return pt_launched("pt_gemm");
}
