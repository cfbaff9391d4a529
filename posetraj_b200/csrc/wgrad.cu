// posetraj_b200 — weight gradient of the linear / implicit-GEMM convolution layers on tcgen05 (configs[3] training).
//
//   dW[n, t*K + k] = sum_r dD[r, n] * A[r + shift_t, k]          (oracle/backward.py conv_rows_wgrad)
//
// The contraction runs over ROWS (pixels x frames x batch: 80 640 at level 0), the output is the small [N, taps*K]
// weight matrix.  Per CTA: one 128 (n) x up-to-256 (k) tile of one tap, over one slice of the rows.  BOTH operands are
// read in the layout the tensors already have (rows outermost), i.e. MN-major for the tensor core:
//   A operand  = gradient tile  [64 rows x 128 n], n contiguous: two TMA boxes {64 n, 64 rows} of dD itself
//   B operand  = input tile     [64 rows x kw k],  k contiguous: kw/64 TMA boxes {64 k, 64 rows} of the layer's input,
//                row coordinate shifted by the tap (zero halo / TMA out-of-bounds fill = the padding, as in the forward)
//   SWIZZLE_128B atoms of 8 rows x 64 elements; LBO = 8 KiB between 64-element blocks along n / k, SBO = 1 KiB between
//   8-row groups; a K step of 16 rows advances both descriptors by 2 KiB.
//   accumulator 128 x kw fp32 in TMEM, written as an fp32 partial tile; pt_reduce_partials folds the row slices in a
//   fixed order (deterministic, no floating-point atomics).
// Round-2 first version: 128 x 64 tiles with a K-major A operand that needed dD^T from pt_transpose_bf16 — 24 KiB of
// L2->SM traffic per 1 MFLOP (L2-bound at ~330 TFLOP/s) plus 8 ms of transposes per step; the 256-wide tile halves the
// bytes per FLOP and the transposes are gone.
// Roles: warp 0 TMA producer, warp 1 MMA issue (both run with all lanes, one elected lane issues), warps 2..5 epilogue.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kWgStages = 4;
constexpr int kWgBlkBytes = 64 * 64 * 2;           // one TMA box: 64 rows x 64 elements
constexpr int kWgABytes = 2 * kWgBlkBytes;         // gradient tile: 64 rows x 128 n
constexpr int kWgBBytes = 4 * kWgBlkBytes;         // input tile: 64 rows x (up to) 256 k
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgThreads = 32 * 6;
constexpr int kWgTmemCols = 256;

struct WgradParams {
  int rows, N, K, num_taps;
  int tap_shift[9];
  int splits, rows_per_split;   // rows_per_split is a multiple of 64
  int n_tiles, k_tiles;          // k_tiles CTAs share the K / 64 column blocks of a tap as evenly as possible
  float* partials;
};

struct alignas(64) WgTmap {
  uint64_t opaque[16];
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ WgTmap tmap_dt, const __grid_constant__ WgTmap tmap_a, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* acc_bar = empty_bar + kWgStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
  uint8_t* tiles = smem + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // blockIdx.x = (tap * n_tiles + n_tile) * k_tiles + k_tile ; blockIdx.y = row slice
  const int k_tile = blockIdx.x % p.k_tiles;
  const int n_tile = (blockIdx.x / p.k_tiles) % p.n_tiles;
  const int tap = blockIdx.x / (p.k_tiles * p.n_tiles);
  const int r_begin = blockIdx.y * p.rows_per_split;
  const int r_end = min(p.rows, r_begin + p.rows_per_split);
  const int steps = r_end > r_begin ? (r_end - r_begin + 63) / 64 : 0;
  // this CTA's column blocks of the tap: blocks [kb0, kb0 + kbn) of K / 64
  const int kblocks = p.K / 64;
  const int kb_base = kblocks / p.k_tiles, kb_rem = kblocks % p.k_tiles;
  const int kbn = kb_base + (k_tile < kb_rem ? 1 : 0);
  const int kb0 = k_tile * kb_base + min(k_tile, kb_rem);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dt);
    tma_prefetch_desc(&tmap_a);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kWgTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < steps; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* sA = tiles + (size_t)stage * kWgStageBytes;
      const int r0 = r_begin + i * 64;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(kWgABytes + kbn * kWgBlkBytes));
        tma_load_2d(sA, &tmap_dt, &full_bar[stage], n_tile * 128, r0);                                // {n, rows}
        tma_load_2d(sA + kWgBlkBytes, &tmap_dt, &full_bar[stage], n_tile * 128 + 64, r0);
        for (int j = 0; j < kbn; ++j)
          tma_load_2d(sA + kWgABytes + j * kWgBlkBytes, &tmap_a, &full_bar[stage], (kb0 + j) * 64, r0 + p.tap_shift[tap]);  // {k, rows}
      }
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)(kbn * 64), 1, 1);   // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < steps; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sA = smem_u32(tiles + (size_t)stage * kWgStageBytes);
      // 64 rows (the contraction dimension) of 128 B per 64-element block, 8-row swizzle atoms of 1024 B (SBO), blocks
      // 8 KiB apart (LBO); a K step of 16 rows = 2048 B
      const uint64_t adesc = make_smem_desc(sA, kWgBlkBytes, 1024, 2);
      const uint64_t bdesc = make_smem_desc(sA + kWgABytes, kWgBlkBytes, 1024, 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tmem_u, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (i | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // epilogue: thread = one n row of the tile, kbn * 64 fp32 columns
    const int q = warp & 3;
    const int n = n_tile * 128 + q * 32 + lane;
    float* dst = p.partials + ((size_t)blockIdx.y * p.N + n) * ((size_t)p.num_taps * p.K) + (size_t)tap * p.K + kb0 * 64;
    if (steps > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < 2 * kbn; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_acc + (uint32_t)c * 32u, v);
        tmem_wait_ld();
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    } else if (n < p.N) {
      for (int j = 0; j < kbn * 64; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_wgrad(const PtWgradArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->tmap_dt && a->tmap_a && a->partials, "pt_wgrad: null argument");
  PT_CHECK_ARG(a->rows > 0 && a->N > 0 && a->K > 0 && a->K % 64 == 0 && a->num_taps >= 1 && a->num_taps <= 9 && a->splits >= 1,
               "pt_wgrad: K must be a multiple of 64, 1..9 taps, splits >= 1");
  PT_CHECK_ARG(((size_t)a->num_taps * a->K) % 4 == 0 && (reinterpret_cast<uintptr_t>(a->partials) & 15u) == 0, "pt_wgrad: partials must be 16-byte aligned");
  WgradParams p;
  p.rows = a->rows; p.N = a->N; p.K = a->K; p.num_taps = a->num_taps;
  for (int i = 0; i < 9; ++i) p.tap_shift[i] = a->tap_shift[i];
  p.splits = a->splits;
  const int steps = (a->rows + 63) / 64;
  p.rows_per_split = ((steps + a->splits - 1) / a->splits) * 64;
  p.n_tiles = (a->N + 127) / 128;
  p.k_tiles = (a->K / 64 + 3) / 4;   // <= 4 column blocks (256 k) per CTA
  p.partials = a->partials;
  const size_t smem_bytes = 1024 + (size_t)kWgStages * kWgStageBytes + 1024;
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return pt_fail(e, "pt_wgrad: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  const long long gx = (long long)a->num_taps * p.n_tiles * p.k_tiles;
  PT_CHECK_ARG(gx <= 0x7fffffffLL && a->splits <= 65535, "pt_wgrad: grid too large");
  WgTmap td, ta;
  memcpy(&td, a->tmap_dt, sizeof(td));
  memcpy(&ta, a->tmap_a, sizeof(ta));
  pt_launch(wgrad_kernel, dim3((unsigned)gx, (unsigned)a->splits), dim3(kWgThreads), smem_bytes, stream, 1, td, ta, p);
  return pt_launched("pt_wgrad");
}
