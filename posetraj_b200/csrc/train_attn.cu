// posetraj_b200 — attention backward for the training step of BASELINE configs[3].
//
// Reference: autograd through BasicTransformerBlock.attn1 / TemporalBasicTransformerBlock.attn1
// (F.scaled_dot_product_attention inside diffusers' Attention, run by models/modified_svd.py:64-107) during
// `accelerator.backward(loss)`, scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1470.  Formulas: oracle/backward.py
// (attention_backward), the FlashAttention-2 recomputation scheme:
//     P = exp2(scale*log2e * Q K^T - lse),  D_i = sum_d dO_id O_id,
//     dV = P^T dO,   dP = dO V^T,   dS = P o (dP - D),   dQ = scale * dS K,   dK = scale * dS^T Q.
//
//   pt_attention_delta           D = rowsum(dO o O) per (row, head), fp32
//   pt_attention_spatial_bwd     two kernels, no atomics (deterministic): dK/dV per 64-key block looping over the query
//                                blocks, dQ per 64-query block looping over the key blocks; each recomputes S and dP.
//                                Warp-level mma.sync (m16n8k16 bf16, fp32 accumulate) on cp.async-staged, XOR-swizzled
//                                64 x 64 shared-memory tiles.  7 matmuls of S^2*64 MACs per (image, head): 2.5x + 1x
//                                (recomputation) the forward's FLOPs.
//   pt_attention_temporal_bwd    one warp per (batch, pixel, head): the F x F problem (F <= 32) in fp32 through shared
//                                memory; HBM-bound (reads qkv + dO, writes dqkv once).
// These are first, correct versions on the legacy tensor-core path; the forward kernels are tcgen05 (attn_spatial.cu).
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr float kLog2e = 1.4426950408889634f;

PT_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

PT_DEVICE void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

PT_DEVICE void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

PT_DEVICE void cp_async16(uint32_t saddr, const void* g, bool ok) {
  const int sz = ok ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
PT_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PT_DEVICE void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// 64 rows x 64 bf16 tile (128 B per row), 16-byte chunks XOR-swizzled by (row & 7): conflict-free ldmatrix.
constexpr int kBwdTile = 64;
constexpr int kBwdTileBytes = 64 * 128;

PT_DEVICE uint32_t tile_addr(uint32_t base, int row, int chunk) { return base + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

// rows [row0, row0 + 64) of one (image, head, part) column block -> smem; rows >= S are zero-filled
PT_DEVICE void load_tile(uint32_t sbase, const bf16* g, int ld, int row0, int S, int tid) {
  for (int i = tid; i < 512; i += 128) {
    const int r = i >> 3, ch = i & 7;
    const bool ok = row0 + r < S;
    cp_async16(tile_addr(sbase, r, ch), g + (size_t)(ok ? row0 + r : 0) * ld + ch * 8, ok);
  }
}

// A fragments (16 rows x 64 k) of rows [r0, r0+16) of a tile: 4 k-steps
PT_DEVICE void load_a_frags(uint32_t (&a)[4][4], uint32_t sbase, int r0, int lane) {
  const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], tile_addr(sbase, row, ks * 2 + (lane >> 4)));
}

// acc[nt] (16 x 8 each, nt over the 64 tile rows) += A(16 x 64) * T^T where the tile T is stored [n rows][k cols]
PT_DEVICE void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int np = 0; np < 4; ++np) {      // pairs of n-tiles
    const int nrow = np * 16 + (lane & 7) + (mi >> 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      ldsm_x4(b, tile_addr(sbase, nrow, ks * 2 + (mi & 1)));
      mma_bf16_16816(acc[np * 2], a[ks], b[0], b[1]);
      mma_bf16_16816(acc[np * 2 + 1], a[ks], b[2], b[3]);
    }
  }
}

// acc[nt] (16 x 8 each, nt over the 64 tile COLUMNS) += A(16 x 64) * T where the tile T is stored [k rows][n cols]
PT_DEVICE void mma_a_tile(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int krow = ks * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_trans(b, tile_addr(sbase, krow, np * 2 + (mi >> 1)));
      mma_bf16_16816(acc[np * 2], a[ks], b[0], b[1]);
      mma_bf16_16816(acc[np * 2 + 1], a[ks], b[2], b[3]);
    }
  }
}

// C fragments of a 16 x 64 block -> A fragments (bf16) of the same block
PT_DEVICE void c_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

struct AttnBwdParams {
  const bf16* qkv;   // [n_img*S, ld] = Q | K | V
  int ld;
  const bf16* dout;  // [n_img*S, dout_ld]
  int dout_ld;
  const float* lse;    // [n_img, heads, S] log2 domain
  const float* delta;  // [n_img, heads, S]
  bf16* dqkv;          // [n_img*S, dld] = dQ | dK | dV
  int dld;
  int S, heads, C;
  float scale, scale_log2;
};

// ------------------------------------------------------------------------------------------------------------
// delta = rowsum(dO o O): one warp per (row, head)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* o, int o_ld, const bf16* dout, int d_ld, float* delta, long long rows,
                                                         int S, int heads) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const long long row = w / heads;
  const int head = (int)(w - row * heads);
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + (size_t)row * o_ld + head * 64 + 2 * lane));
  const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + (size_t)row * d_ld + head * 64 + 2 * lane));
  const float t = warp_sum(a.x * b.x + a.y * b.y);
  if (lane == 0) {
    const long long img = row / S;
    const int s = (int)(row - img * S);
    delta[((size_t)img * heads + head) * S + s] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------
// dK, dV of one 64-key block: warp w owns keys [16w, 16w+16); loop over the query blocks
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sK = smem_u32(smem), sV = sK + kBwdTileBytes;
  const uint32_t sQ0 = sV + kBwdTileBytes;           // 2 stages x (Q, dO)
  float* s_lse = reinterpret_cast<float*>(smem + 6 * kBwdTileBytes);  // [2][64]
  float* s_delta = s_lse + 128;                                        // [2][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int key0 = blockIdx.x * kBwdTile, head = blockIdx.y, img = blockIdx.z;
  const size_t img_row = (size_t)img * p.S;
  const bf16* qg = p.qkv + img_row * p.ld + head * 64;
  const bf16* kg = qg + p.C;
  const bf16* vg = qg + 2 * p.C;
  const bf16* dog = p.dout + img_row * p.dout_ld + head * 64;
  const float* lse_g = p.lse + ((size_t)img * p.heads + head) * p.S;
  const float* delta_g = p.delta + ((size_t)img * p.heads + head) * p.S;
  const int n_q = (p.S + kBwdTile - 1) / kBwdTile;

  auto load_q = [&](int st, int it) {
    const int q0 = it * kBwdTile;
    load_tile(sQ0 + (uint32_t)st * 2 * kBwdTileBytes, qg, p.ld, q0, p.S, tid);
    load_tile(sQ0 + (uint32_t)st * 2 * kBwdTileBytes + kBwdTileBytes, dog, p.dout_ld, q0, p.S, tid);
    if (tid < 64) {
      const bool ok = q0 + tid < p.S;
      s_lse[st * 64 + tid] = ok ? lse_g[q0 + tid] : INFINITY;   // exp2(x - inf) = 0: padded queries contribute nothing
      s_delta[st * 64 + tid] = ok ? delta_g[q0 + tid] : 0.f;
    }
  };
  load_tile(sK, kg, p.ld, key0, p.S, tid);
  load_tile(sV, vg, p.ld, key0, p.S, tid);
  load_q(0, 0);
  cp_async_commit();

  uint32_t ka[4][4], va[4][4];
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
  const bool key_ok0 = key0 + warp * 16 + g < p.S, key_ok1 = key0 + warp * 16 + g + 8 < p.S;

  for (int it = 0; it < n_q; ++it) {
    const int st = it & 1;
    if (it + 1 < n_q) load_q(st ^ 1, it + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (it == 0) {
      load_a_frags(ka, sK, warp * 16, lane);
      load_a_frags(va, sV, warp * 16, lane);
    }
    const uint32_t sQ = sQ0 + (uint32_t)st * 2 * kBwdTileBytes, sdO = sQ + kBwdTileBytes;
    float sc[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = dp[i][j] = 0.f;
    mma_a_tileT(sc, ka, sQ, lane);    // S^T  = K Q^T   [16 keys x 64 queries]
    mma_a_tileT(dp, va, sdO, lane);   // dP^T = V dO^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int qc = nt * 8 + 2 * t;
      const float l0 = s_lse[st * 64 + qc], l1 = s_lse[st * 64 + qc + 1];
      const float d0 = s_delta[st * 64 + qc], d1 = s_delta[st * 64 + qc + 1];
      float p0 = key_ok0 ? exp2f(sc[nt][0] * p.scale_log2 - l0) : 0.f;
      float p1 = key_ok0 ? exp2f(sc[nt][1] * p.scale_log2 - l1) : 0.f;
      float p2 = key_ok1 ? exp2f(sc[nt][2] * p.scale_log2 - l0) : 0.f;
      float p3 = key_ok1 ? exp2f(sc[nt][3] * p.scale_log2 - l1) : 0.f;
      sc[nt][0] = p0; sc[nt][1] = p1; sc[nt][2] = p2; sc[nt][3] = p3;
      dp[nt][0] = p0 * (dp[nt][0] - d0);
      dp[nt][1] = p1 * (dp[nt][1] - d1);
      dp[nt][2] = p2 * (dp[nt][2] - d0);
      dp[nt][3] = p3 * (dp[nt][3] - d1);
    }
    uint32_t pa[4][4], dsa[4][4];
    c_to_a(pa, sc);
    c_to_a(dsa, dp);
    mma_a_tile(dv, pa, sdO, lane);    // dV += P^T dO
    mma_a_tile(dk, dsa, sQ, lane);    // dK += dS^T Q
    __syncthreads();
  }
  cp_async_wait<0>();
  // write dK (x scale) and dV
  bf16* dkg = p.dqkv + img_row * p.dld + p.C + head * 64;
  bf16* dvg = p.dqkv + img_row * p.dld + 2 * p.C + head * 64;
  const int r0 = key0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (r0 < p.S) {
      *reinterpret_cast<uint32_t*>(dkg + (size_t)r0 * p.dld + c) = pack_bf16x2(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
      *reinterpret_cast<uint32_t*>(dvg + (size_t)r0 * p.dld + c) = pack_bf16x2(dv[nt][0], dv[nt][1]);
    }
    if (r1 < p.S) {
      *reinterpret_cast<uint32_t*>(dkg + (size_t)r1 * p.dld + c) = pack_bf16x2(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
      *reinterpret_cast<uint32_t*>(dvg + (size_t)r1 * p.dld + c) = pack_bf16x2(dv[nt][2], dv[nt][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// dQ of one 64-query block: warp w owns queries [16w, 16w+16); loop over the key blocks
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sdO = sQ + kBwdTileBytes;
  const uint32_t sK0 = sdO + kBwdTileBytes;          // 2 stages x (K, V)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * kBwdTile, head = blockIdx.y, img = blockIdx.z;
  const size_t img_row = (size_t)img * p.S;
  const bf16* qg = p.qkv + img_row * p.ld + head * 64;
  const bf16* kg = qg + p.C;
  const bf16* vg = qg + 2 * p.C;
  const bf16* dog = p.dout + img_row * p.dout_ld + head * 64;
  const int n_k = (p.S + kBwdTile - 1) / kBwdTile;

  auto load_kv = [&](int st, int it) {
    load_tile(sK0 + (uint32_t)st * 2 * kBwdTileBytes, kg, p.ld, it * kBwdTile, p.S, tid);
    load_tile(sK0 + (uint32_t)st * 2 * kBwdTileBytes + kBwdTileBytes, vg, p.ld, it * kBwdTile, p.S, tid);
  };
  load_tile(sQ, qg, p.ld, q0, p.S, tid);
  load_tile(sdO, dog, p.dout_ld, q0, p.S, tid);
  load_kv(0, 0);
  cp_async_commit();

  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float* lse_g = p.lse + ((size_t)img * p.heads + head) * p.S;
  const float* delta_g = p.delta + ((size_t)img * p.heads + head) * p.S;
  const float l0 = r0 < p.S ? lse_g[r0] : INFINITY, l1 = r1 < p.S ? lse_g[r1] : INFINITY;
  const float d0 = r0 < p.S ? delta_g[r0] : 0.f, d1 = r1 < p.S ? delta_g[r1] : 0.f;

  uint32_t qa[4][4], doa[4][4];
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

  for (int it = 0; it < n_k; ++it) {
    const int st = it & 1;
    if (it + 1 < n_k) load_kv(st ^ 1, it + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (it == 0) {
      load_a_frags(qa, sQ, warp * 16, lane);
      load_a_frags(doa, sdO, warp * 16, lane);
    }
    const uint32_t sK = sK0 + (uint32_t)st * 2 * kBwdTileBytes, sV = sK + kBwdTileBytes;
    float sc[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = dp[i][j] = 0.f;
    mma_a_tileT(sc, qa, sK, lane);    // S  = Q K^T   [16 queries x 64 keys]
    mma_a_tileT(dp, doa, sV, lane);   // dP = dO V^T
    const int kbase = it * kBwdTile;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = kbase + nt * 8 + 2 * t;
      const bool k0ok = kc < p.S, k1ok = kc + 1 < p.S;
      const float p0 = k0ok ? exp2f(sc[nt][0] * p.scale_log2 - l0) : 0.f;
      const float p1 = k1ok ? exp2f(sc[nt][1] * p.scale_log2 - l0) : 0.f;
      const float p2 = k0ok ? exp2f(sc[nt][2] * p.scale_log2 - l1) : 0.f;
      const float p3 = k1ok ? exp2f(sc[nt][3] * p.scale_log2 - l1) : 0.f;
      dp[nt][0] = p0 * (dp[nt][0] - d0);
      dp[nt][1] = p1 * (dp[nt][1] - d0);
      dp[nt][2] = p2 * (dp[nt][2] - d1);
      dp[nt][3] = p3 * (dp[nt][3] - d1);
    }
    uint32_t dsa[4][4];
    c_to_a(dsa, dp);
    mma_a_tile(dq, dsa, sK, lane);    // dQ += dS K
    __syncthreads();
  }
  cp_async_wait<0>();
  bf16* dqg = p.dqkv + img_row * p.dld + head * 64;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (r0 < p.S) *reinterpret_cast<uint32_t*>(dqg + (size_t)r0 * p.dld + c) = pack_bf16x2(dq[nt][0] * p.scale, dq[nt][1] * p.scale);
    if (r1 < p.S) *reinterpret_cast<uint32_t*>(dqg + (size_t)r1 * p.dld + c) = pack_bf16x2(dq[nt][2] * p.scale, dq[nt][3] * p.scale);
  }
}

// ------------------------------------------------------------------------------------------------------------
// dK, dV on tcgen05 (S >= 256): one CTA per (128-key block, head, image), looping over 128-query blocks.
//   S  = Q_i K_j^T   and   dP = dO_i V_j^T              128 x 128 fp32 each, in TMEM (rows = queries = TMEM lanes, so the
//                                                       softmax statistics lse / delta are per-thread scalars)
//   P  = exp2(scale S - lse),  dS = P o (dP - delta)    -> bf16, SWIZZLE_128B smem tiles [queries x keys]
//   dV += P^T dO_i,   dK += dS^T Q_i                    128 x 64 fp32 each, in TMEM (rows = keys): P / dS are read as the
//                                                       MN-major A operand (= transposed for free), Q_i / dO_i as MN-major B
// warp 0: TMA producer (K_j, V_j once; ring of {Q_i, dO_i}); warp 1: MMA issue; warps 2..9: two warps per TMEM lane quarter,
// one per 64-key half of the tile — a thread owns one query row and 64 of its 128 score columns.
// Same anatomy as the forward kernel (attn_spatial.cu); no atomics, dQ keeps its own kernel.
// ------------------------------------------------------------------------------------------------------------
constexpr int kTcTile = 128;
constexpr int kTcTileBytes = 128 * 64 * 2;   // 16 KiB: one [128 x 64] bf16 tile (K, V, Q, dO, each half of P^T / dS^T)
constexpr uint32_t kTcTmemCols = 512;         // S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384)

struct alignas(64) BwdTmap {
  uint64_t opaque[16];
};

__global__ void __launch_bounds__(320, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ BwdTmap tmap_qkv, const __grid_constant__ BwdTmap tmap_do, const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* qd_full = kv_full + 1;      // [2]
  uint64_t* qd_empty = qd_full + 2;     // [2]
  uint64_t* s_full = qd_empty + 2;      // S^T(i), dP^T(i) complete
  uint64_t* s_read = s_full + 1;        // all 8 softmax warps hold S^T(i), dP^T(i) in registers
  uint64_t* p_full = s_read + 1;        // P^T(i), dS^T(i) staged in shared memory
  uint64_t* pv_done = p_full + 1;       // dV / dK MMAs of iteration i retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);
  uint8_t* sK = smem + 1024;
  uint8_t* sV = sK + kTcTileBytes;
  uint8_t* sQD = sV + kTcTileBytes;                 // 2 stages x (Q 16 KiB | dO 16 KiB)
  uint8_t* sP = sQD + 4 * kTcTileBytes;             // P^T: two 64-query halves
  uint8_t* sDS = sP + 2 * kTcTileBytes;             // dS^T: two 64-query halves

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int key0 = blockIdx.x * kTcTile;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int n_q = (p.S + kTcTile - 1) / kTcTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_read, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * kTcTileBytes);
      tma_load_3d(sK, &tmap_qkv, kv_full, p.C + head * 64, key0, img);
      tma_load_3d(sV, &tmap_qkv, kv_full, 2 * p.C + head * 64, key0, img);
    }
    __syncwarp();
    for (int i = 0; i < n_q; ++i) {
      const int st = i & 1;
      mbar_wait(&qd_empty[st], (uint32_t)((i >> 1) & 1) ^ 1u);
      uint8_t* sQ = sQD + (size_t)st * 2 * kTcTileBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&qd_full[st], 2 * kTcTileBytes);
        tma_load_3d(sQ, &tmap_qkv, &qd_full[st], head * 64, i * kTcTile, img);
        tma_load_3d(sQ + kTcTileBytes, &tmap_do, &qd_full[st], head * 64, i * kTcTile, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // A (K-major) x B (K-major)
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 1, 1);    // A = P^T / dS^T (MN-major: keys contiguous) x B (MN-major: head dim contiguous)
    const uint64_t kdesc = make_desc_kmajor_sw128(smem_u32(sK));
    const uint64_t vdesc = make_desc_kmajor_sw128(smem_u32(sV));
    auto issue_s = [&](int i) {
      const int st = i & 1;
      mbar_wait(&qd_full[st], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t sQ = smem_u32(sQD + (size_t)st * 2 * kTcTileBytes);
      const uint64_t qdesc = make_desc_kmajor_sw128(sQ);
      const uint64_t ddesc = make_desc_kmajor_sw128(sQ + kTcTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_u, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_u + 128u, ddesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    tc_fence_after();
    issue_s(0);
    for (int i = 0; i < n_q; ++i) {
      const uint32_t par = (uint32_t)i & 1u;
      if (i + 1 < n_q) {
        mbar_wait(s_read, par);          // S^T(i), dP^T(i) are in registers: their TMEM columns are free
        tc_fence_after();
        issue_s(i + 1);
      }
      mbar_wait(p_full, par);
      tc_fence_after();
      const int st = i & 1;
      const uint32_t sQ = smem_u32(sQD + (size_t)st * 2 * kTcTileBytes);
      const uint64_t qmn = make_smem_desc(sQ, 1024, 1024, 2);                 // Q_i as [K = queries][N = d], MN-major
      const uint64_t dmn = make_smem_desc(sQ + kTcTileBytes, 1024, 1024, 2);  // dO_i likewise
      // P / dS tiles: [128 queries (K) x 2 x 64 keys (M)], 128 B per row and half: M blocks 16 KiB apart (LBO), 8-row groups
      // 1 KiB apart (SBO); a K step of 16 queries = 2 KiB
      const uint64_t pmn = make_smem_desc(smem_u32(sP), kTcTileBytes, 1024, 2);
      const uint64_t smn = make_smem_desc(smem_u32(sDS), kTcTileBytes, 1024, 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tc_mma_bf16(tmem_u + 256u, pmn + (uint64_t)(k * 128), dmn + (uint64_t)(k * 128), idesc_o, (i | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) tc_mma_bf16(tmem_u + 320u, smn + (uint64_t)(k * 128), qmn + (uint64_t)(k * 128), idesc_o, (i | k) != 0 ? 1u : 0u);
        tc_commit(pv_done);
        tc_commit(&qd_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ softmax warps ----------------------------
    const int q = warp & 3;               // TMEM lane quarter
    const int half = (warp - 2) >> 2;     // 64-key half of the tile
    const int row = q * 32 + lane;        // S / dP row (query of the current block) == TMEM lane; later: key row of dK / dV
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* lse_g = p.lse + ((size_t)img * p.heads + head) * p.S;
    const float* delta_g = p.delta + ((size_t)img * p.heads + head) * p.S;
    const int kbase = key0 + half * 64;                      // first key of this warp's columns
    const bool kfull = kbase + 64 <= p.S;
    for (int i = 0; i < n_q; ++i) {
      const uint32_t par = (uint32_t)i & 1u;
      const int qrow = i * kTcTile + row;
      const float l = qrow < p.S ? __ldg(lse_g + qrow) : INFINITY;   // exp2(x - inf) = 0: padded queries contribute nothing
      const float dl = qrow < p.S ? __ldg(delta_g + qrow) : 0.f;
      mbar_wait(s_full, par);
      tc_fence_after();
      uint32_t sv[64], dv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(t_lane + (uint32_t)(half * 64 + c * 32), *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_ld_32x32(t_lane + 128u + (uint32_t)(half * 64 + c * 32), *reinterpret_cast<uint32_t(*)[32]>(&dv[c * 32]));
      }
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_read);
      // P and dS of this thread's 64 columns, packed to bf16 in registers: the exponentials run while the tensor core still
      // works on dV / dK of the previous query block
      uint4 pk[8], dk[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float pr[8], ds[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pr[j] = ex2_approx(fmaf(__uint_as_float(sv[ch * 8 + j]), p.scale_log2, -l));
          if (!kfull && kbase + ch * 8 + j >= p.S) pr[j] = 0.f;      // keys beyond S are TMA zero-fill
          ds[j] = pr[j] * (__uint_as_float(dv[ch * 8 + j]) - dl);
        }
        pk[ch] = make_uint4(pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]), pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
        dk[ch] = make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
      }
      if (i > 0) {                                           // the staging tiles are free once dV / dK of i-1 retired
        mbar_wait(pv_done, (uint32_t)(i - 1) & 1u);
        tc_fence_after();
      }
      uint8_t* prow = sP + (size_t)half * kTcTileBytes + (size_t)row * 128;
      uint8_t* drow = sDS + (size_t)half * kTcTileBytes + (size_t)row * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const int sw = (ch ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + sw) = pk[ch];
        *reinterpret_cast<uint4*>(drow + sw) = dk[ch];
      }
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // dV (half 1) / dK (half 0) of this key row once the last MMAs retired
    mbar_wait(pv_done, (uint32_t)(n_q - 1) & 1u);
    tc_fence_after();
    const int krow = key0 + row;
    const float osc = half == 0 ? p.scale : 1.0f;
    bf16* dst = p.dqkv + ((size_t)img * p.S + krow) * p.dld + (half == 0 ? p.C : 2 * p.C) + head * 64;
    const uint32_t t_o = t_lane + (half == 0 ? 320u : 256u);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + (uint32_t)c * 32u, o);
      tmem_wait_ld();
      if (krow < p.S) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[d]) * osc, __uint_as_float(o[d + 1]) * osc);
          u.y = pack_bf16x2(__uint_as_float(o[d + 2]) * osc, __uint_as_float(o[d + 3]) * osc);
          u.z = pack_bf16x2(__uint_as_float(o[d + 4]) * osc, __uint_as_float(o[d + 5]) * osc);
          u.w = pack_bf16x2(__uint_as_float(o[d + 6]) * osc, __uint_as_float(o[d + 7]) * osc);
          stg_u4(dst + c * 32 + d, u);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// dQ on tcgen05 (S >= 256): one CTA per (128-query block, head, image), looping over 128-key blocks.
//   S = Q_i K_j^T,  dP = dO_i V_j^T   (rows = queries = TMEM lanes: lse / delta are per-thread scalars)
//   dS = P o (dP - delta) -> bf16, SWIZZLE_128B smem tiles;   dQ += dS K_j   (K_j re-read MN-major), scaled at the end
// Roles as in the dK / dV kernel; the ring carries {K_j, V_j}, Q_i / dO_i stay resident.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ BwdTmap tmap_qkv, const __grid_constant__ BwdTmap tmap_do, const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* qd_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = qd_full + 1;      // [2]
  uint64_t* kv_empty = kv_full + 2;     // [2]
  uint64_t* s_full = kv_empty + 2;
  uint64_t* s_read = s_full + 1;
  uint64_t* p_full = s_read + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);
  uint8_t* sQ = smem + 1024;
  uint8_t* sdO = sQ + kTcTileBytes;
  uint8_t* sKV = sdO + kTcTileBytes;                // 2 stages x (K 16 KiB | V 16 KiB)
  uint8_t* sDS = sKV + 4 * kTcTileBytes;            // dS: two 64-key halves

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTcTile;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int n_k = (p.S + kTcTile - 1) / kTcTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(qd_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_read, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(qd_full, 2 * kTcTileBytes);
      tma_load_3d(sQ, &tmap_qkv, qd_full, head * 64, q0, img);
      tma_load_3d(sdO, &tmap_do, qd_full, head * 64, q0, img);
    }
    __syncwarp();
    for (int j = 0; j < n_k; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_empty[st], (uint32_t)((j >> 1) & 1) ^ 1u);
      uint8_t* sK = sKV + (size_t)st * 2 * kTcTileBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTcTileBytes);
        tma_load_3d(sK, &tmap_qkv, &kv_full[st], p.C + head * 64, j * kTcTile, img);
        tma_load_3d(sK + kTcTileBytes, &tmap_qkv, &kv_full[st], 2 * p.C + head * 64, j * kTcTile, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ));
    const uint64_t ddesc = make_desc_kmajor_sw128(smem_u32(sdO));
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t sK = smem_u32(sKV + (size_t)st * 2 * kTcTileBytes);
      const uint64_t kdesc = make_desc_kmajor_sw128(sK);
      const uint64_t vdesc = make_desc_kmajor_sw128(sK + kTcTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_u, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_u + 128u, ddesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(qd_full, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < n_k; ++j) {
      const uint32_t par = (uint32_t)j & 1u;
      if (j + 1 < n_k) {
        mbar_wait(s_read, par);
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(p_full, par);
      tc_fence_after();
      const int st = j & 1;
      const uint64_t kmn = make_smem_desc(smem_u32(sKV + (size_t)st * 2 * kTcTileBytes), 1024, 1024, 2);   // K_j as [K = keys][N = d]
      const uint32_t sDa = smem_u32(sDS);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t sdesc = make_desc_kmajor_sw128(sDa + (uint32_t)(k >> 2) * kTcTileBytes) + (uint64_t)(2 * (k & 3));
          tc_mma_bf16(tmem_u + 256u, sdesc, kmn + (uint64_t)(k * 128), idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        tc_commit(pv_done);
        tc_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // 64-key half of the tile
    const int row = q * 32 + lane;        // query row inside the block == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int qrow = q0 + row;
    const size_t sidx = ((size_t)img * p.heads + head) * p.S + (qrow < p.S ? qrow : 0);
    const float l = qrow < p.S ? p.lse[sidx] : INFINITY;
    const float dl = qrow < p.S ? p.delta[sidx] : 0.f;
    for (int j = 0; j < n_k; ++j) {
      const uint32_t par = (uint32_t)j & 1u;
      const int kbase = j * kTcTile + half * 64;
      mbar_wait(s_full, par);
      tc_fence_after();
      uint32_t sv[64], dv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(t_lane + (uint32_t)(half * 64 + c * 32), *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_ld_32x32(t_lane + 128u + (uint32_t)(half * 64 + c * 32), *reinterpret_cast<uint32_t(*)[32]>(&dv[c * 32]));
      }
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_read);
      const bool full = kbase + 64 <= p.S;
      uint4 dk[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float ds[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float pr = ex2_approx(fmaf(__uint_as_float(sv[ch * 8 + jj]), p.scale_log2, -l));
          if (!full && kbase + ch * 8 + jj >= p.S) pr = 0.f;     // keys beyond S are TMA zero-fill
          ds[jj] = pr * (__uint_as_float(dv[ch * 8 + jj]) - dl);
        }
        dk[ch] = make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
      }
      if (j > 0) {                                           // the staging tile is free once dQ += dS K of j-1 retired
        mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      uint8_t* drow = sDS + (size_t)half * kTcTileBytes + (size_t)row * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(drow + ((ch ^ (row & 7)) << 4)) = dk[ch];
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (uint32_t)(n_k - 1) & 1u);
    tc_fence_after();
    // dQ: this warp writes 32 of the row's 64 head dims
    uint32_t o[32];
    tmem_ld_32x32(t_lane + 256u + (uint32_t)half * 32u, o);
    tmem_wait_ld();
    if (qrow < p.S) {
      bf16* dst = p.dqkv + ((size_t)img * p.S + qrow) * p.dld + head * 64 + half * 32;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o[d]) * p.scale, __uint_as_float(o[d + 1]) * p.scale);
        u.y = pack_bf16x2(__uint_as_float(o[d + 2]) * p.scale, __uint_as_float(o[d + 3]) * p.scale);
        u.z = pack_bf16x2(__uint_as_float(o[d + 4]) * p.scale, __uint_as_float(o[d + 5]) * p.scale);
        u.w = pack_bf16x2(__uint_as_float(o[d + 6]) * p.scale, __uint_as_float(o[d + 7]) * p.scale);
        stg_u4(dst + d, u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// temporal attention backward: one warp per (batch, pixel, head); rows (b*F + f)*HW + s
// ------------------------------------------------------------------------------------------------------------
struct TAttnBwdParams {
  const bf16* qkv;
  int ld;
  const bf16* dout;
  int dout_ld;
  bf16* dqkv;
  int dld;
  int B, F, HW, heads, C;
  float scale, scale_log2;
  int warps_per_cta;
};

constexpr int kTStride = 68;  // fp32 row stride of the per-warp q/k/v/dO copies: 16-byte aligned rows whose float4 reads of 8
                              // consecutive rows cover all 32 banks

__global__ void attn_temporal_bwd_kernel(const TAttnBwdParams p) {
  extern __shared__ __align__(16) float tsm[];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const long long w = (long long)blockIdx.x * p.warps_per_cta + wl;
  const long long total = (long long)p.B * p.HW * p.heads;
  if (w >= total) return;   // whole warps only; no block-wide barriers below
  const int F = p.F;
  const int head = (int)(w % p.heads);
  const long long bs = w / p.heads;
  const int s = (int)(bs % p.HW);
  const int b = (int)(bs / p.HW);
  float* base = tsm + (size_t)wl * (size_t)((4 * F * kTStride + 2 * F * F + 3) & ~3);   // 16-byte aligned per-warp slabs
  float* sq = base;
  float* sk = sq + F * kTStride;
  float* sv = sk + F * kTStride;
  float* sdo = sv + F * kTStride;
  float* sp = sdo + F * kTStride;   // [F][F] probabilities
  float* sds = sp + F * F;          // [F][F] dP, then dS
  const size_t row0 = (size_t)b * F * p.HW + s;
  for (int f = 0; f < F; ++f) {
    const size_t row = row0 + (size_t)f * p.HW;
    const bf16* r = p.qkv + row * p.ld + head * 64 + 2 * lane;
    const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(r));
    const float2 k = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(r + p.C));
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(r + 2 * p.C));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.dout + row * p.dout_ld + head * 64 + 2 * lane));
    *reinterpret_cast<float2*>(sq + f * kTStride + 2 * lane) = q;
    *reinterpret_cast<float2*>(sk + f * kTStride + 2 * lane) = k;
    *reinterpret_cast<float2*>(sv + f * kTStride + 2 * lane) = v;
    *reinterpret_cast<float2*>(sdo + f * kTStride + 2 * lane) = d;
  }
  __syncwarp();
  // S = Q K^T and dP = dO V^T
  for (int e = lane; e < F * F; e += 32) {
    const int i = e / F, j = e - i * F;
    float a = 0.f, c = 0.f;
#pragma unroll 4
    for (int d = 0; d < 64; d += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(sq + i * kTStride + d), k4 = *reinterpret_cast<const float4*>(sk + j * kTStride + d);
      const float4 o4 = *reinterpret_cast<const float4*>(sdo + i * kTStride + d), v4 = *reinterpret_cast<const float4*>(sv + j * kTStride + d);
      a = fmaf(q4.x, k4.x, a); a = fmaf(q4.y, k4.y, a); a = fmaf(q4.z, k4.z, a); a = fmaf(q4.w, k4.w, a);
      c = fmaf(o4.x, v4.x, c); c = fmaf(o4.y, v4.y, c); c = fmaf(o4.z, v4.z, c); c = fmaf(o4.w, v4.w, c);
    }
    sp[e] = a;
    sds[e] = c;
  }
  __syncwarp();
  // row softmax, then dS = P o (dP - sum_j P dP) * scale
  for (int i = 0; i < F; ++i) {
    const float sv_ = lane < F ? sp[i * F + lane] : -INFINITY;
    const float m = warp_max(sv_);
    const float e = lane < F ? exp2f((sv_ - m) * p.scale_log2) : 0.f;
    const float l = warp_sum(e);
    const float pr = e / l;
    const float dpv = lane < F ? sds[i * F + lane] : 0.f;
    const float dl = warp_sum(pr * dpv);
    if (lane < F) {
      sp[i * F + lane] = pr;
      sds[i * F + lane] = pr * (dpv - dl) * p.scale;
    }
  }
  __syncwarp();
  // dQ_i = sum_j dS_ij K_j ; dK_j = sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i   (lane owns head dims 2*lane, 2*lane+1)
  const int c0 = 2 * lane;
  for (int f = 0; f < F; ++f) {
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int j = 0; j < F; ++j) {
      const float ds_fj = sds[f * F + j], ds_jf = sds[j * F + f], p_jf = sp[j * F + f];
      const float2 kk = *reinterpret_cast<const float2*>(sk + j * kTStride + c0);
      const float2 qq = *reinterpret_cast<const float2*>(sq + j * kTStride + c0);
      const float2 oo = *reinterpret_cast<const float2*>(sdo + j * kTStride + c0);
      q0 = fmaf(ds_fj, kk.x, q0);
      q1 = fmaf(ds_fj, kk.y, q1);
      k0 = fmaf(ds_jf, qq.x, k0);
      k1 = fmaf(ds_jf, qq.y, k1);
      v0 = fmaf(p_jf, oo.x, v0);
      v1 = fmaf(p_jf, oo.y, v1);
    }
    bf16* o = p.dqkv + (row0 + (size_t)f * p.HW) * p.dld + head * 64 + c0;
    *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(q0, q1);
    *reinterpret_cast<uint32_t*>(o + p.C) = pack_bf16x2(k0, k1);
    *reinterpret_cast<uint32_t*>(o + 2 * p.C) = pack_bf16x2(v0, v1);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_attention_delta(const void* out, int32_t out_ld, const void* dout, int32_t dout_ld, float* delta, int64_t rows,
                                  int32_t S, int32_t heads, void* stream) {
  PT_CHECK_ARG(out && dout && delta && rows > 0 && S > 0 && heads > 0 && rows % S == 0, "pt_attention_delta: bad argument");
  PT_CHECK_ARG(out_ld % 2 == 0 && dout_ld % 2 == 0, "pt_attention_delta: row strides must be even");
  const long long warps = (long long)rows * heads;
  pt_launch(attn_delta_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(out), (int)out_ld,
            reinterpret_cast<const bf16*>(dout), (int)dout_ld, delta, (long long)rows, (int)S, (int)heads);
  return pt_launched("pt_attention_delta");
}

extern "C" int pt_attention_spatial_bwd(const PtAttnSpatialBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->qkv && a->dout && a->lse && a->delta && a->dqkv, "pt_attention_spatial_bwd: null argument");
  PT_CHECK_ARG(a->S > 0 && a->heads > 0 && a->n_img > 0 && a->C == a->heads * 64, "pt_attention_spatial_bwd: need C == heads*64");
  PT_CHECK_ARG(a->ld % 8 == 0 && a->dout_ld % 8 == 0 && a->dld % 2 == 0, "pt_attention_spatial_bwd: row strides must be multiples of 8");
  AttnBwdParams p;
  p.qkv = reinterpret_cast<const bf16*>(a->qkv); p.ld = a->ld;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.dout_ld = a->dout_ld;
  p.lse = a->lse; p.delta = a->delta;
  p.dqkv = reinterpret_cast<bf16*>(a->dqkv); p.dld = a->dld;
  p.S = a->S; p.heads = a->heads; p.C = a->C;
  p.scale = 0.125f; p.scale_log2 = 0.125f * kLog2e;
  const size_t smem = 6 * kBwdTileBytes + 4 * 64 * sizeof(float);
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pt_fail(e, "pt_attention_spatial_bwd: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  dim3 grid((a->S + kBwdTile - 1) / kBwdTile, a->heads, a->n_img);
  int rc;
  static int env_tc = -2;   // PT_ATTN_BWD_TC=0 keeps the mma.sync dK / dV kernel at every size
  if (env_tc == -2) {
    const char* e = getenv("PT_ATTN_BWD_TC");
    env_tc = e ? atoi(e) : 1;
  }
  if (env_tc != 0 && a->tmap_qkv != nullptr && a->tmap_dout != nullptr && a->S >= 256 && a->dld % 8 == 0) {
    const size_t smem_tc = 1024 + (size_t)kTcTileBytes * (2 + 4 + 2 + 2) + 1024;
    static bool attr_tc[PT_MAX_DEVICES] = {false};
    if (!attr_tc[dev_slot]) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
      if (e != cudaSuccess) return pt_fail(e, "pt_attention_spatial_bwd: cudaFuncSetAttribute (tcgen05)");
      attr_tc[dev_slot] = true;
    }
    BwdTmap tq, td;
    memcpy(&tq, a->tmap_qkv, sizeof(tq));
    memcpy(&td, a->tmap_dout, sizeof(td));
    dim3 grid_tc((a->S + kTcTile - 1) / kTcTile, a->heads, a->n_img);
    pt_launch(attn_bwd_dkv_tc_kernel, grid_tc, dim3(320), smem_tc, stream, 1, tq, td, p);
    rc = pt_launched("pt_attention_spatial_bwd (dK, dV: tcgen05)");
  } else {
    pt_launch(attn_bwd_dkv_kernel, grid, dim3(128), smem, stream, 1, p);
    rc = pt_launched("pt_attention_spatial_bwd (dK, dV)");
  }
  if (rc != 0) return rc;
  if (env_tc != 0 && a->tmap_qkv != nullptr && a->tmap_dout != nullptr && a->S >= 256 && a->dld % 8 == 0) {
    const size_t smem_tc = 1024 + (size_t)kTcTileBytes * (2 + 4 + 2) + 1024;
    static bool attr_tq[PT_MAX_DEVICES] = {false};
    if (!attr_tq[dev_slot]) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
      if (e != cudaSuccess) return pt_fail(e, "pt_attention_spatial_bwd: cudaFuncSetAttribute (tcgen05 dQ)");
      attr_tq[dev_slot] = true;
    }
    BwdTmap tq, td;
    memcpy(&tq, a->tmap_qkv, sizeof(tq));
    memcpy(&td, a->tmap_dout, sizeof(td));
    dim3 grid_tc((a->S + kTcTile - 1) / kTcTile, a->heads, a->n_img);
    pt_launch(attn_bwd_dq_tc_kernel, grid_tc, dim3(320), smem_tc, stream, 1, tq, td, p);
    return pt_launched("pt_attention_spatial_bwd (dQ: tcgen05)");
  }
  pt_launch(attn_bwd_dq_kernel, grid, dim3(128), smem, stream, 1, p);
  return pt_launched("pt_attention_spatial_bwd (dQ)");
}

extern "C" int pt_attention_temporal_bwd(const PtAttnTemporalBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->qkv && a->dout && a->dqkv, "pt_attention_temporal_bwd: null argument");
  PT_CHECK_ARG(a->B > 0 && a->F > 0 && a->F <= 32 && a->HW > 0 && a->heads > 0 && a->C == a->heads * 64,
               "pt_attention_temporal_bwd: need C == heads*64 and 1 <= F <= 32");
  PT_CHECK_ARG(a->ld % 2 == 0 && a->dout_ld % 2 == 0 && a->dld % 2 == 0, "pt_attention_temporal_bwd: row strides must be even");
  TAttnBwdParams p;
  p.qkv = reinterpret_cast<const bf16*>(a->qkv); p.ld = a->ld;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.dout_ld = a->dout_ld;
  p.dqkv = reinterpret_cast<bf16*>(a->dqkv); p.dld = a->dld;
  p.B = a->B; p.F = a->F; p.HW = a->HW; p.heads = a->heads; p.C = a->C;
  p.scale = 0.125f; p.scale_log2 = 0.125f * kLog2e;
  const size_t per_warp = (size_t)((4 * a->F * kTStride + 2 * a->F * a->F + 3) & ~3) * sizeof(float);
  int wpc = (int)((size_t)96 * 1024 / per_warp);
  if (wpc > 8) wpc = 8;
  if (wpc < 1) wpc = 1;
  p.warps_per_cta = wpc;
  const size_t smem = per_warp * wpc;
  static int attr_smem[PT_MAX_DEVICES] = {0};
  const int dev_slot = pt_device_slot();
  if ((int)smem > attr_smem[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pt_fail(e, "pt_attention_temporal_bwd: cudaFuncSetAttribute");
    attr_smem[dev_slot] = (int)smem;
  }
  const long long warps = (long long)a->B * a->HW * a->heads;
  pt_launch(attn_temporal_bwd_kernel, dim3((unsigned)((warps + wpc - 1) / wpc)), dim3(32 * wpc), smem, stream, 1, p);
  return pt_launched("pt_attention_temporal_bwd");
}
