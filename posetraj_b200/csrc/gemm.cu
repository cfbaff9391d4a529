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
// Structure (per CTA, one CTA per SM, tiles 128 x block_n, K step 64):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> SWIZZLE_128B smem ring, mbarrier expect_tx)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma kind::f16, accumulators in TMEM, 2 accumulator stages)
//   warps 2..9    : epilogue       (tcgen05.ld -> bias / row-vector / residual / blend / GEGLU -> global)
// The three pipelines (smem full/empty, TMEM full/empty, static persistent tile schedule) follow the
// canonical Blackwell GEMM anatomy; block_n, stage count and tap table are runtime values so the single
// instantiation covers all shapes.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
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
  int rowvec_ld, rowvec_mode, rv_a, rv_b, rv_c;
  float acc_scale;
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
};

struct alignas(64) TmapParam {
  uint64_t opaque[16];
};

PT_DEVICE void store8(int out_dtype, void* base, size_t off, const float* v, int nvalid) {
  if (out_dtype == PT_DT_BF16) {
    bf16* o = reinterpret_cast<bf16*>(base) + off;
    if (nvalid == 8) {
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]);
      u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]);
      u.w = pack_bf16x2(v[6], v[7]);
      stg_u4(o, u);
    } else {
      for (int j = 0; j < nvalid; ++j) o[j] = __float2bfloat16(v[j]);
    }
  } else {
    float* o = reinterpret_cast<float*>(base) + off;
    if (nvalid == 8) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int j = 0; j < nvalid; ++j) o[j] = v[j];
    }
  }
}

// 8 bf16 of a residual-like operand -> raw registers (zero when absent); vector path needs 16-byte alignment,
// which holds whenever ld % 8 == 0 and n % 8 == 0 (checked on the host for the vector case).
PT_DEVICE uint4 load8_raw(const bf16* src, int nvalid) {
  if (nvalid == 8) return ldg_u4(src);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  for (int j = 0; j < nvalid; ++j) {
    const uint32_t b = (uint32_t)__bfloat16_as_ushort(src[j]);
    w[j >> 1] |= (j & 1) ? (b << 16) : b;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

PT_DEVICE void fma8(float* v, uint4 u, float s) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = fmaf(s, a.x, v[0]); v[1] = fmaf(s, a.y, v[1]); v[2] = fmaf(s, b.x, v[2]); v[3] = fmaf(s, b.y, v[3]);
  v[4] = fmaf(s, c.x, v[4]); v[5] = fmaf(s, c.y, v[5]); v[6] = fmaf(s, d.x, v[6]); v[7] = fmaf(s, d.y, v[7]);
}

// Operands of one 32-column chunk of one accumulator row, fetched BEFORE the TMEM wait so that their
// global-load latency overlaps the tcgen05.ld and is paid once per chunk rather than once per 8 columns.
struct ChunkOperands {
  uint4 r1[4], r2[4], ax[4];
  float rv[32];
};

PT_DEVICE void prefetch_chunk(const GemmParams& p, ChunkOperands& o, int n, long long orow, int grp, bool valid) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    const int nn = n + g8 * 8;
    const int nvalid = valid ? max(0, min(8, p.n_out - nn)) : 0;
    o.r1[g8] = (p.res1 != nullptr && nvalid > 0) ? load8_raw(p.res1 + (size_t)orow * p.res_ld + nn, nvalid) : make_uint4(0, 0, 0, 0);
    o.r2[g8] = (p.res2 != nullptr && nvalid > 0) ? load8_raw(p.res2 + (size_t)orow * p.res_ld + nn, nvalid) : make_uint4(0, 0, 0, 0);
    o.ax[g8] = (p.out2 != nullptr && nvalid > 0) ? load8_raw(p.aux + (size_t)orow * p.out_ld + nn, nvalid) : make_uint4(0, 0, 0, 0);
  }
  if (p.rowvec_mode != 0 && valid) {
    const float* rv = p.rowvec + (size_t)grp * p.rowvec_ld + n;
    if (n + 32 <= p.n_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rv) + j);
        o.rv[4 * j] = t.x; o.rv[4 * j + 1] = t.y; o.rv[4 * j + 2] = t.z; o.rv[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) o.rv[j] = (n + j < p.n_out) ? __ldg(rv + j) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) o.rv[j] = 0.f;
  }
}

// f[32] already holds acc (+bias, GEGLU applied); finish and store the chunk.
PT_DEVICE void finish_chunk(const GemmParams& p, float* f, const ChunkOperands& o, int n, long long orow) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    const int nn = n + g8 * 8;
    const int nvalid = min(8, p.n_out - nn);
    if (nvalid <= 0) break;
    float* v = f + g8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] + o.rv[g8 * 8 + j]) * p.acc_scale;
    if (p.act_silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
    }
    if (p.res1 != nullptr) fma8(v, o.r1[g8], p.res1_scale);
    if (p.res2 != nullptr) fma8(v, o.r2[g8], p.res2_scale);
    store8(p.out_dtype, p.out, (size_t)orow * p.out_ld + nn, v, nvalid);
    if (p.out2 != nullptr) {
      fma8(v, o.ax[g8], p.aux_scale);
      store8(p.out_dtype, p.out2, (size_t)orow * p.out_ld + nn, v, nvalid);
    }
  }
}

template <bool kGeglu>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ TmapParam tmap_a0, const __grid_constant__ TmapParam tmap_a1,
                    const __grid_constant__ TmapParam tmap_b, const __grid_constant__ GemmParams p) {
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
  uint8_t* tiles = smem + kSmemCtl;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int k_chunks = p.k0_chunks + p.k1_chunks;
  const int k_iters = p.num_taps * k_chunks;
  const uint32_t b_bytes = (uint32_t)p.block_n * kBlockK * 2;

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
      mbar_init(&tempty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_tile = t / p.num_n_tiles;
        const int n_tile = t - m_tile * p.num_n_tiles;
        const int batch = m_tile / p.tiles_per_batch;
        const int r0 = (m_tile - batch * p.tiles_per_batch) * kBlockM;
        const int half = p.block_n >> 1;
        const int n0 = kGeglu ? n_tile * half : n_tile * p.block_n;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int arow = r0 + p.tap_shift[tap];
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sA = tiles + (size_t)stage * p.stage_bytes;
            uint8_t* sB = sA + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kABytes + b_bytes);
            if (kc < p.k0_chunks) {
              tma_load_3d(sA, &tmap_a0, &full_bar[stage], kc * kBlockK, arow, batch);
            } else {
              tma_load_3d(sA, &tmap_a1, &full_bar[stage], (kc - p.k0_chunks) * kBlockK, arow, batch);
            }
            const int kcol = (tap * k_chunks + kc) * kBlockK;
            if constexpr (kGeglu) {
              tma_load_2d(sB, &tmap_b, &full_bar[stage], kcol, n0);
              tma_load_2d(sB + (size_t)half * kBlockK * 2, &tmap_b, &full_bar[stage], kcol,
                          p.gate_row_offset + n0);
            } else {
              tma_load_2d(sB, &tmap_b, &full_bar[stage], kcol, n0);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(tiles + (size_t)stage * p.stage_bytes);
          const uint64_t adesc = make_desc_kmajor_sw128(sA);
          const uint64_t bdesc = make_desc_kmajor_sw128(sA + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes (16 bf16) along K inside the 128-byte swizzle row: +2 in the >>4 address field
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                        (ki | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------ epilogue warps ----------------------------
    // 8 warps: warp % 4 selects the TMEM lane quarter (hardware restriction), (warp - 2) / 4 the chunk parity.
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;  // 0..255
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m_tile = t / p.num_n_tiles;
      const int n_tile = t - m_tile * p.num_n_tiles;
      const int batch = m_tile / p.tiles_per_batch;
      const int r = (m_tile - batch * p.tiles_per_batch) * kBlockM + q * 32 + lane;
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
        grp = (int)(((orow / p.rv_a) * p.rv_b + orow % p.rv_b) % p.rv_c);
      }
      if (!valid) { orow = 0; grp = 0; }

      // stage this tile's bias in smem, indexed like the accumulator columns
      const int half = p.block_n >> 1;
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
      asm volatile("bar.sync 1, 256;" ::: "memory");

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      const int chunks = (kGeglu ? half : p.block_n) >> 5;
      if (hsel >= chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      for (int c = hsel; c < chunks; c += 2) {
        const int n = n0 + c * 32;
        float f[32];
        if constexpr (kGeglu) {
          // GEGLU tiles carry bias only (host-checked): out = (x + bx) * gelu(g + bg)
          uint32_t v[32];
          uint32_t g[32];
          tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
          tmem_ld_32x32(t_acc + (uint32_t)(half + c * 32), g);
          tmem_wait_ld();
          if (c + 2 >= chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xv = __uint_as_float(v[j]) + sb[c * 32 + j];
            const float gv = __uint_as_float(g[j]) + sb[half + c * 32 + j];
            f[j] = xv * gelu_erf_f(gv);
          }
          if (valid) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int nvalid = min(8, p.n_out - (n + g8 * 8));
              if (nvalid > 0) store8(p.out_dtype, p.out, (size_t)orow * p.out_ld + n + g8 * 8, f + g8 * 8, nvalid);
            }
          }
        } else {
          uint32_t v[32];
          tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
          ChunkOperands ops;
          prefetch_chunk(p, ops, n, orow, grp, valid);
          tmem_wait_ld();
          if (c + 2 >= chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + sb[c * 32 + j];
          if (valid) finish_chunk(p, f, ops, n, orow);
        }
      }
    }
  }

  // ------------------------------ teardown -------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_gemm(const PtGemmArgs* a, void* stream) {
  if (a == nullptr || a->tmap_a0 == nullptr || a->tmap_b == nullptr || a->out == nullptr)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: null argument");
  if (a->block_n < 32 || a->block_n > 256 || (a->block_n % 32) != 0)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: block_n must be a multiple of 32 in [32,256]");
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
  if (a->out2 != nullptr && a->aux == nullptr)
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: out2 without aux");
  if (a->geglu && (a->rowvec_mode != 0 || a->res1 != nullptr || a->res2 != nullptr || a->out2 != nullptr || a->acc_scale != 1.0f))
    return pt_fail(cudaErrorInvalidValue, "pt_gemm: GEGLU tiles support a bias-only epilogue");

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
  p.stage_bytes = kABytes + a->block_n * kBlockK * 2;
  const int smem_limit = 227 * 1024 - kSmemCtl - 1024;
  int stages = smem_limit / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  p.tiles_per_batch = (a->rows_per_batch + kBlockM - 1) / kBlockM;
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
  p.acc_scale = a->acc_scale;
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

  const size_t smem_bytes = (size_t)kSmemCtl + (size_t)p.stages * p.stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pt_fail(e, "pt_gemm: cudaFuncSetAttribute");
    attr_set = true;
  }
  const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  const int sms = pt_num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);

  TmapParam ta0, ta1, tb;
  memcpy(&ta0, a->tmap_a0, sizeof(TmapParam));
  memcpy(&ta1, a->tmap_a1 ? a->tmap_a1 : a->tmap_a0, sizeof(TmapParam));
  memcpy(&tb, a->tmap_b, sizeof(TmapParam));
  if (p.geglu)
    gemm_tcgen05_kernel<true><<<grid, kGemmThreads, smem_bytes, (cudaStream_t)stream>>>(ta0, ta1, tb, p);
  else
    gemm_tcgen05_kernel<false><<<grid, kGemmThreads, smem_bytes, (cudaStream_t)stream>>>(ta0, ta1, tb, p);
  return pt_launched("pt_gemm");
}
