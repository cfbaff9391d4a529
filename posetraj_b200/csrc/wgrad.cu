// posetraj_b200 — weight gradient of the linear / implicit-GEMM convolution layers on tcgen05 (configs[3] training).
//
//   dW[n, t*K + k] = sum_r dD[r, n] * A[r + shift_t, k]          (oracle/backward.py conv_rows_wgrad)
//
// The contraction runs over ROWS (pixels x frames x batch: 80 640 at level 0), the output is the small [N, taps*K]
// weight matrix.  Per CTA: one 128 (n) x 64 (k) tile of one tap, over one slice of the rows:
//   A operand  = dD^T tile [128 n x 64 rows], K-major (rows contiguous): TMA box of the transposed gradient
//   B operand  = input tile [64 rows x 64 k], MN-major (k contiguous): TMA box of the layer's input, row coordinate
//                shifted by the tap (the zero halo / TMA out-of-bounds fill supply the padding, as in the forward)
//   accumulator 128 x 64 fp32 in TMEM, written as an fp32 partial tile; pt_reduce_partials folds the row slices in a
//   fixed order (deterministic, no floating-point atomics).
// Roles: warp 0 TMA producer, warp 1 MMA issue (both run with all lanes, one elected lane issues), warps 2..5 epilogue.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kWgStages = 6;
constexpr int kWgABytes = 128 * 64 * 2;   // dD^T tile: 128 n x 64 rows
constexpr int kWgBBytes = 64 * 64 * 2;    // input tile: 64 rows x 64 k
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgThreads = 32 * 6;

struct WgradParams {
  int rows, N, K, num_taps;
  int tap_shift[9];
  int splits, rows_per_split;   // rows_per_split is a multiple of 64
  int n_tiles, k_tiles;
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
  if (warp == 1) tmem_alloc(tmem_ptr, 64);
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
        mbar_arrive_expect_tx(&full_bar[stage], kWgStageBytes);
        tma_load_2d(sA, &tmap_dt, &full_bar[stage], r0, n_tile * 128);                               // {rows, n}
        tma_load_2d(sA + kWgABytes, &tmap_a, &full_bar[stage], k_tile * 64, r0 + p.tap_shift[tap]);  // {k, rows}
      }
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major (as P x V in attn_spatial.cu)
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < steps; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sA = smem_u32(tiles + (size_t)stage * kWgStageBytes);
      const uint64_t adesc = make_desc_kmajor_sw128(sA);
      // 64 rows (the contraction dimension) of 128 B each, 8-row swizzle atoms of 1024 B; a K step of 16 rows = 2048 B
      const uint64_t bdesc = make_smem_desc(sA + kWgABytes, 1024, 1024, 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tmem_u, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(k * 128), idesc, (i | k) != 0 ? 1u : 0u);
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
    // epilogue: thread = one n row of the tile, 64 fp32 columns
    const int q = warp & 3;
    const int n = n_tile * 128 + q * 32 + lane;
    float* dst = p.partials + ((size_t)blockIdx.y * p.N + n) * ((size_t)p.num_taps * p.K) + (size_t)tap * p.K + k_tile * 64;
    if (steps > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
      for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
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
  p.k_tiles = a->K / 64;
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
