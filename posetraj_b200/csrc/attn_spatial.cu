// posetraj_b200 — spatial self-attention (per frame, non-causal, head_dim 64) on tcgen05 / TMEM / TMA.
//
// Replaces BasicTransformerBlock.attn1 -> F.scaled_dot_product_attention (batch B*F, heads
// 5/10/20, sequence H*W in {2880, 720, 180, 45}).  Input is the fused QKV projection [rows, 3C] (Q | K | V,
// head h at columns h*64 of each part); output [rows, C].
//
// One CTA per (128-query tile, head, image):
//   warp 0 lane 0 : TMA producer  — Q tile once, then a ring of {K_j, V_j} tiles (128 keys each)
//   warp 1 lane 0 : MMA issuer    — S_j = Q K_j^T (128x128, fp32 in TMEM, double-buffered), O_j = P_j V_j (128x64)
//   warps 2..5    : softmax       — one thread per query row: tcgen05.ld S, online max/sum with exp2, P_j as
//                                   bf16 into a SWIZZLE_128B smem tile (A operand of the PV MMA), running output
//                                   kept in registers and rescaled there (O_j is read back from TMEM per tile).
// The QK^T of tile j+1 is issued before the softmax of tile j finishes, so the tensor pipe overlaps the
// MUFU-bound softmax.  FLOPs: 4*S*64 per query row per head.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kAttnThreads = 192;
constexpr int kQTile = 128;
constexpr int kKTile = 128;
constexpr int kHd = 64;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: Q, K, V tiles and each 64-key half of P
constexpr int kKvStages = 3;
constexpr uint32_t kAttnTmemCols = 512;   // S0 [0,128) S1 [128,256) O [256,320)

struct AttnParams {
  int S, heads, C;
  float scale_log2;  // head_dim^-0.5 * log2(e)
  bf16* out;
  int out_ld;
};

struct alignas(64) AttnTmap {
  uint64_t opaque[16];
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_spatial_kernel(const __grid_constant__ AttnTmap tmap_qkv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = q_full + 1;              // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;    // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;     // [2]
  uint64_t* p_full = s_full + 2;               // [1]
  uint64_t* o_full = p_full + 1;               // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  uint8_t* sQ = smem + 1024;
  uint8_t* sP = sQ + kTileBytes;               // 2 x 16 KiB (keys 0-63 | keys 64-127)
  uint8_t* sKV = sP + 2 * kTileBytes;          // kKvStages x (K 16 KiB + V 16 KiB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int n_kv = (p.S + kKTile - 1) / kKTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kAttnTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tmap_qkv, q_full, head * kHd, q0, img);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        uint8_t* sK = sKV + (size_t)stage * 2 * kTileBytes;
        mbar_arrive_expect_tx(&kv_full[stage], 2 * kTileBytes);
        tma_load_3d(sK, &tmap_qkv, &kv_full[stage], p.C + head * kHd, j * kKTile, img);
        tma_load_3d(sK + kTileBytes, &tmap_qkv, &kv_full[stage], 2 * p.C + head * kHd, j * kKTile, img);
        if (++stage == kKvStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // Q (K-major) x K (K-major)
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // P (K-major) x V (MN-major: hd contiguous)
      const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ));
      auto issue_s = [&](int j, int stage) {
        const uint64_t kdesc = make_desc_kmajor_sw128(smem_u32(sKV + (size_t)stage * 2 * kTileBytes));
        const uint32_t d = tmem_base + (uint32_t)(j & 1) * 128u;
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          tc_mma_bf16(d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        tc_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      int stage_s = 0;  // stage of the next S to issue
      uint32_t phase_s = 0;
      int stage_o = 0;  // stage of the next PV to issue
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      if (++stage_s == kKvStages) { stage_s = 0; phase_s ^= 1u; }
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(&kv_full[stage_s], phase_s);
          tc_fence_after();
          issue_s(j + 1, stage_s);
          if (++stage_s == kKvStages) { stage_s = 0; phase_s ^= 1u; }
        }
        mbar_wait(p_full, (uint32_t)j & 1u);  // P_j in smem, O TMEM drained, S_j consumed
        tc_fence_after();
        const uint32_t sV = smem_u32(sKV + (size_t)stage_o * 2 * kTileBytes + kTileBytes);
        // V tile: rows = keys (128 B each, 8-row swizzle atoms of 1024 B): MN-major B operand, K step of 16 keys
        // = 2048 B; LBO (stride between 64-wide N blocks) is unused for N = 64.
        const uint64_t vdesc = make_smem_desc(sV, 1024, 1024, 2);
#pragma unroll
        for (int k = 0; k < kKTile / 16; ++k) {
          const uint64_t pdesc = make_desc_kmajor_sw128(smem_u32(sP) + (uint32_t)(k >> 2) * kTileBytes) +
                                 (uint64_t)(2 * (k & 3));
          tc_mma_bf16(tmem_base + 256u, pdesc, vdesc + (uint64_t)(k * 128), idesc_o, k != 0 ? 1u : 0u);
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[stage_o]);
        if (++stage_o == kKvStages) stage_o = 0;
      }
    }
  } else {
    // ------------------------------ softmax warps ----------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float o_acc[kHd];
#pragma unroll
    for (int d = 0; d < kHd; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY;  // running max of raw scores
    float l_run = 0.f;
    float alpha_prev = 1.f;   // rescale owed to o_acc before adding the previous tile's O
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_s = t_lane + (uint32_t)(j & 1) * 128u;
      const int kvalid = min(kKTile, p.S - j * kKTile);  // keys beyond S are TMA zero-fill: mask them
      // pass 1: row max
      float m_new = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + (uint32_t)c * 32u, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kvalid) m_new = fmaxf(m_new, __uint_as_float(v[i]));
      }
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);  // 0 on the first tile (m_run = -inf)
      // fold the previous tile's O into the register accumulator (also frees the O columns of TMEM)
      if (j > 0) {
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + 256u + (uint32_t)c * 32u, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      alpha_prev = alpha;
      // pass 2: P = exp2((s - m) * scale), row sum, bf16 P into the swizzled smem tile
      const float mb = m_new * p.scale_log2;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + (uint32_t)c * 32u, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c * 32 + i < kvalid) ? exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -mb)) : 0.f;
          float p1 = (c * 32 + i + 1 < kvalid) ? exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mb)) : 0.f;
          pk[i >> 1] = pack_bf16x2(p0, p1);
          // sum what the tensor core will actually multiply (bf16-rounded probabilities)
          const float2 r = unpack_bf16x2(pk[i >> 1]);
          lsum += r.x + r.y;
        }
        // keys [c*32, c*32+32) -> half (c >> 1), 16-byte chunks (c & 1)*4 .. +3 of this row, XOR-swizzled
        uint8_t* prow = sP + (size_t)(c >> 1) * kTileBytes + (size_t)row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c & 1) * 4 + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) =
              make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l_run = fmaf(l_run, alpha, lsum);
      m_run = m_new;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last tile's O
    mbar_wait(o_full, (uint32_t)(n_kv - 1) & 1u);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_lane + 256u + (uint32_t)c * 32u, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
    }
    if (q0 + row < p.S) {
      const float inv = 1.0f / l_run;
      bf16* dst = p.out + ((size_t)img * p.S + q0 + row) * p.out_ld + head * kHd;
#pragma unroll
      for (int d = 0; d < kHd; d += 8) {
        uint4 u;
        u.x = pack_bf16x2(o_acc[d] * inv, o_acc[d + 1] * inv);
        u.y = pack_bf16x2(o_acc[d + 2] * inv, o_acc[d + 3] * inv);
        u.z = pack_bf16x2(o_acc[d + 4] * inv, o_acc[d + 5] * inv);
        u.w = pack_bf16x2(o_acc[d + 6] * inv, o_acc[d + 7] * inv);
        stg_u4(dst + d, u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_attention_spatial(const PtAttnSpatialArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->tmap_qkv != nullptr && a->out != nullptr, "pt_attention_spatial: null argument");
  PT_CHECK_ARG(a->S > 0 && a->heads > 0 && a->n_img > 0 && a->C == a->heads * kHd,
               "pt_attention_spatial: need C == heads*64 and a non-empty problem");
  AttnParams p;
  p.S = a->S;
  p.heads = a->heads;
  p.C = a->C;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  const size_t smem_bytes = 1024 + (size_t)kTileBytes * (1 + 2 + 2 * kKvStages) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return pt_fail(e, "pt_attention_spatial: cudaFuncSetAttribute");
    attr_set = true;
  }
  AttnTmap tm;
  memcpy(&tm, a->tmap_qkv, sizeof(tm));
  dim3 grid((a->S + kQTile - 1) / kQTile, a->heads, a->n_img);
  attn_spatial_kernel<<<grid, kAttnThreads, smem_bytes, (cudaStream_t)stream>>>(tm, p);
  return pt_launched("pt_attention_spatial");
}
