// posetraj_b200 — spatial self-attention (per frame, non-causal, head_dim 64) on tcgen05 / TMEM / TMA.
//
// Replaces BasicTransformerBlock.attn1 -> F.scaled_dot_product_attention (batch B*F, heads
// 5/10/20, sequence H*W in {2880, 720, 180, 45}).  Input is the fused QKV projection [rows, 3C] (Q | K | V,
// head h at columns h*64 of each part); output [rows, C].
//
// One CTA per (NQ x 128 queries, head, image), NQ = 2 (two query tiles share every K/V tile) or 1 (S <= 128):
//   warp 0 lane 0   : TMA producer — the NQ Q tiles once, then a ring of {K_j, V_j} tiles (128 keys each)
//   warp 1 lane 0   : MMA issuer   — S_g = Q_g K_j^T (128x128 fp32 in TMEM), O_g += P_g V_j (128x64 in TMEM)
//   warps 2..5 (+6..9) : softmax group g — one thread per query row: ONE tcgen05.ld of S per tile (128 registers),
//                     running max with LAZY rescaling (O and l are only rescaled when the max grew by more than 2^8,
//                     so O stays in TMEM and is read back once at the end), P_g as bf16 into a SWIZZLE_128B smem
//                     tile (A operand of the PV MMA).
// While group 0 runs its softmax the tensor pipe works for group 1 and vice versa; the exponentials (MUFU ex2,
// 16/clk/SM) are the floor: 256x128 of them per K/V tile against 1024 MMA cycles.
// FLOPs: 4*S*64 per query row per head.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

constexpr int kQTile = 128;
constexpr int kKTile = 128;
constexpr int kHd = 64;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: Q, K, V tiles and each 64-key half of P
constexpr int kKvStages = 3;
constexpr uint32_t kAttnTmemCols = 512;   // S_0 [0,128) S_1 [128,256) O_0 [256,320) O_1 [320,384)
constexpr float kRescaleThreshold = 8.0f; // in log2 units: P <= 2^8 between rescales

struct AttnParams {
  int S, heads, C;
  float scale_log2;  // head_dim^-0.5 * log2(e)
  bf16* out;
  int out_ld;
  float* lse;  // optional [n_img, heads, S]: log2-domain log-sum-exp of the scaled scores (training: pt_attention_spatial_bwd)
};

template <bool B>
struct FullTile {
  static constexpr bool value = B;
};

struct alignas(64) AttnTmap {
  uint64_t opaque[16];
};

template <int NQ>
__global__ void __launch_bounds__(64 + 128 * NQ, 1)
attn_spatial_kernel(const __grid_constant__ AttnTmap tmap_qkv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = q_full + 1;              // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;    // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;     // [2]  S_g(j) complete
  uint64_t* p_full = s_full + 2;               // [2]  P_g(j) in smem, S_g(j) consumed, O_g rescaled
  uint64_t* o_done = p_full + 2;               // [2]  PV_g(j) retired
  uint64_t* s_read = o_done + 2;               // [2]  S_g(j) is in the softmax warps' registers: its TMEM columns are free
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_read + 2);
  uint8_t* sQ = smem + 1024;                   // NQ x 16 KiB
  uint8_t* sP = sQ + NQ * kTileBytes;          // NQ x 2 x 16 KiB (keys 0-63 | keys 64-127)
  uint8_t* sKV = sP + NQ * 2 * kTileBytes;     // kKvStages x (K 16 KiB + V 16 KiB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (kQTile * NQ);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int n_kv = (p.S + kKTile - 1) / kKTile;
  // query groups of this CTA that hold at least one real query: the last CTA of a sequence whose length is not a multiple
  // of NQ x 128 (2880 = 11 x 256 + 64) runs ONE group instead of spending a full group on TMA zero-fill
  const int nq_act = (NQ == 2 && q0 + kQTile >= p.S) ? 1 : NQ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_done[g], 1);
      mbar_init(&s_read[g], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kAttnTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // PDL: everything above is private to this CTA; the qkv rows are the previous kernel's output
  griddep_launch();
  griddep_wait();

  // Producer and MMA-issue warps run with all 32 lanes; one elected lane issues the asynchronous instruction, whose
  // operands then live in uniform registers (no per-instruction R2UR waterfall: see elect_one in common.cuh).
  if (warp == 0) {
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, nq_act * kTileBytes);
        for (int g = 0; g < nq_act; ++g) tma_load_3d(sQ + g * kTileBytes, &tmap_qkv, q_full, head * kHd, q0 + g * kQTile, img);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        uint8_t* sK = sKV + (size_t)stage * 2 * kTileBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[stage], 2 * kTileBytes);
          tma_load_3d(sK, &tmap_qkv, &kv_full[stage], p.C + head * kHd, j * kKTile, img);
          tma_load_3d(sK + kTileBytes, &tmap_qkv, &kv_full[stage], 2 * p.C + head * kHd, j * kKTile, img);
        }
        __syncwarp();
        if (++stage == kKvStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t tmem_u = uniform_u32(tmem_base);
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // Q (K-major) x K (K-major)
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // P (K-major) x V (MN-major: hd contiguous)
      auto issue_s = [&](int g, int stage) {
        const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ + g * kTileBytes));
        const uint64_t kdesc = make_desc_kmajor_sw128(smem_u32(sKV + (size_t)stage * 2 * kTileBytes));
        const uint32_t d = tmem_u + (uint32_t)g * 128u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k)
            tc_mma_bf16(d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
          tc_commit(&s_full[g]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      for (int g = 0; g < nq_act; ++g) issue_s(g, 0);
      int stage = 0;       // stage of K_j / V_j
      int stage_n = 1 % kKvStages;  // stage of K_{j+1}
      uint32_t phase_n = (kKvStages == 1) ? 1u : 0u;
      for (int j = 0; j < n_kv; ++j) {
        const bool has_next = j + 1 < n_kv;
        if (has_next) {
          mbar_wait(&kv_full[stage_n], phase_n);
          tc_fence_after();
        }
        const uint32_t sV = smem_u32(sKV + (size_t)stage * 2 * kTileBytes + kTileBytes);
        // V tile: rows = keys (128 B each, 8-row swizzle atoms of 1024 B): MN-major B operand, K step of 16 keys
        // = 2048 B; LBO (stride between 64-wide N blocks) is unused for N = 64.
        const uint64_t vdesc = make_smem_desc(sV, 1024, 1024, 2);
        // Event-driven issue: S_g(j+1) goes out as soon as the softmax warps hold S_g(j) in registers (its latency
        // hides behind their exponentials), PV_g(j) as soon as P_g(j) is in shared memory — whichever comes first.
        uint32_t pend_s = has_next ? ((1u << nq_act) - 1u) : 0u;
        uint32_t pend_pv = (1u << nq_act) - 1u;
        const uint32_t par = (uint32_t)j & 1u;
        uint32_t spins = 0;
        while (pend_s | pend_pv) {
          bool progressed = false;
#pragma unroll
          for (int g = 0; g < NQ; ++g) {
            // (the polls are made warp-uniform: every lane must take the same branch around elect_one)
            if (((pend_s >> g) & 1u) && __all_sync(0xffffffffu, mbar_try_wait(&s_read[g], par))) {
              tc_fence_after();
              issue_s(g, stage_n);
              pend_s &= ~(1u << g);
              progressed = true;
            }
            if (((pend_pv >> g) & 1u) && __all_sync(0xffffffffu, mbar_try_wait(&p_full[g], par))) {  // P_g(j) in smem, O_g rescaled
              tc_fence_after();
              const uint32_t sPg = smem_u32(sP + (size_t)g * 2 * kTileBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kKTile / 16; ++k) {
                  const uint64_t pdesc = make_desc_kmajor_sw128(sPg + (uint32_t)(k >> 2) * kTileBytes) + (uint64_t)(2 * (k & 3));
                  tc_mma_bf16(tmem_u + 256u + (uint32_t)g * 64u, pdesc, vdesc + (uint64_t)(k * 128), idesc_o,
                              (j | k) != 0 ? 1u : 0u);
                }
                tc_commit(&o_done[g]);
              }
              __syncwarp();
              pend_pv &= ~(1u << g);
              progressed = true;
            }
          }
          if (!progressed && ++spins > (1u << 26)) asm volatile("trap;");
        }
        if (elect_one()) tc_commit(&kv_empty[stage]);
        __syncwarp();
        stage = stage_n;
        if (++stage_n == kKvStages) {
          stage_n = 0;
          phase_n ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------ softmax warps ----------------------------
    const int g = (warp - 2) >> 2;      // query tile of this warp
    const int q = warp & 3;             // TMEM lane quarter
    const int row = q * 32 + lane;      // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = t_lane + (uint32_t)g * 128u;
    const uint32_t t_o = t_lane + 256u + (uint32_t)g * 64u;
    uint8_t* sPg = sP + (size_t)g * 2 * kTileBytes;
    float m_ref = -INFINITY;  // the max the exponent offsets currently refer to (raw score units)
    float l_run = 0.f;
    const bool active = g < nq_act;   // (an inactive group's warps go straight to the final barrier)
    for (int j = 0; active && j < n_kv; ++j) {
      mbar_wait(&s_full[g], (uint32_t)j & 1u);
      tc_fence_after();
      const int kvalid = min(kKTile, p.S - j * kKTile);  // keys beyond S are TMA zero-fill: mask them
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + (uint32_t)c * 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_wait_ld();
      // S_g(j) now lives in registers: hand its TMEM columns back so that Q_g K_{j+1}^T can start right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_read[g]);
      // row max with 4 independent chains (a single FMNMX chain is 128 dependent ops)
      float m_tile;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (kvalid == kKTile) {
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i + u < kvalid) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
          }
        }
        m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      // PV_g(j-1) must have retired before P_g is overwritten / O_g is rescaled
      if (j > 0) {
        mbar_wait(&o_done[g], (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      const bool grow = (m_tile - m_ref) * p.scale_log2 > kRescaleThreshold;  // also true on the first tile
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_tile : m_ref;
        const float alpha = ex2_approx((m_ref - m_new) * p.scale_log2);  // 0 on the first tile, 1 for rows that keep m_ref
        l_run *= alpha;
        m_ref = m_new;
        if (j > 0) {
          // rare path (the running max grew by > 2^8): 8 columns at a time keeps the live S registers intact
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint32_t o[8];
            tmem_ld_32x8(t_o + (uint32_t)c * 8u, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x8(t_o + (uint32_t)c * 8u, o);
          }
          tmem_wait_st();
        }
      }
      // P = exp2((s - m_ref) * scale) as bf16 into the swizzled smem tile; row sum of what the MMA will multiply.
      // Two instantiations: only the LAST key tile of a sequence can be partial, and the per-element masking
      // (ISETP + FSEL) was a third of the instructions of this loop when it ran on every tile (profiles/r2d_attn.md).
      const float mb = m_ref * p.scale_log2;
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent row-sum chains
      auto exp_tile = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int k0 = c * 32 + i;
            float p0 = ex2_approx(fmaf(__uint_as_float(v[k0]), p.scale_log2, -mb));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[k0 + 1]), p.scale_log2, -mb));
            if constexpr (!kFull) {
              if (k0 >= kvalid) p0 = 0.f;
              if (k0 + 1 >= kvalid) p1 = 0.f;
            }
            pk[i >> 1] = pack_bf16x2(p0, p1);
            // row sum in fp32 of the un-rounded probabilities (the bf16 rounding of P is unbiased; re-deriving the
            // rounded values costs 3 extra ALU ops per pair in a loop that is issue-bound)
            ls4[(i >> 1) & 3] += p0 + p1;
          }
          // keys [c*32, c*32+32) -> half (c >> 1), 16-byte chunks (c & 1)*4 .. +3 of this row, XOR-swizzled
          uint8_t* prow = sPg + (size_t)(c >> 1) * kTileBytes + (size_t)row * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int chunk = ((c & 1) * 4 + ch) ^ (row & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
          }
        }
      };
      if (kvalid == kKTile) exp_tile(FullTile<true>{});
      else exp_tile(FullTile<false>{});
      const float lsum = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      l_run += lsum;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
    // O_g = sum_j P_g(j) V_j is complete once the last PV retired
    if (active) {
    mbar_wait(&o_done[g], (uint32_t)(n_kv - 1) & 1u);
    tc_fence_after();
    const int qrow = q0 + g * kQTile + row;
    const float inv = 1.0f / l_run;
    bf16* dst = p.out + ((size_t)img * p.S + qrow) * p.out_ld + head * kHd;
    if (p.lse != nullptr && qrow < p.S)
      p.lse[((size_t)img * p.heads + head) * p.S + qrow] = m_ref * p.scale_log2 + log2f(l_run);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + (uint32_t)c * 32u, o);
      tmem_wait_ld();
      if (qrow < p.S) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[d]) * inv, __uint_as_float(o[d + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[d + 2]) * inv, __uint_as_float(o[d + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[d + 4]) * inv, __uint_as_float(o[d + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[d + 6]) * inv, __uint_as_float(o[d + 7]) * inv);
          stg_u4(dst + c * 32 + d, u);
        }
      }
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

template <int NQ>
static int launch_attn(const PtAttnSpatialArgs* a, const AttnParams& p, cudaStream_t st) {
  const size_t smem_bytes = 1024 + (size_t)kTileBytes * (NQ + 2 * NQ + 2 * kKvStages) + 1024;
  static bool attr_set[PT_MAX_DEVICES] = {false};  // cudaFuncSetAttribute is per device
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return pt_fail(e, "pt_attention_spatial: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  AttnTmap tm;
  memcpy(&tm, a->tmap_qkv, sizeof(tm));
  dim3 grid((a->S + kQTile * NQ - 1) / (kQTile * NQ), a->heads, a->n_img);
  pt_launch(attn_spatial_kernel<NQ>, dim3(grid), dim3(64 + 128 * NQ), smem_bytes, (void*)st, 1, tm, p);
  return pt_launched("pt_attention_spatial");
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
  p.lse = a->lse;
  if (a->S <= kQTile) return launch_attn<1>(a, p, (cudaStream_t)stream);
  return launch_attn<2>(a, p, (cudaStream_t)stream);
}
