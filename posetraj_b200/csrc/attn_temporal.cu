// posetraj_b200 — temporal self-attention over frames (sequence F <= 32, head_dim 64, non-causal).
//
// Replaces TemporalBasicTransformerBlock.attn1 (models/modified_svd.py:79-81): for every (batch, pixel, head)
// a tiny F x F attention.  4*F*64 FLOPs per token (0.018 TFLOP/step) against 4 x 128 B of traffic per token
// and head: the op is HBM-bound, so it is a one-warp-per-problem register kernel (mma.sync m16n8k16 on
// fragments loaded straight from global memory); no shared memory, no TMEM.
// Input: fused QKV [rows, 3C] with row = (b*F + f)*HW + s; output [rows, C].
// Algorithmic bytes per launch: rows * (3C + C) * 2.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

struct TAttnParams {
  const bf16* qkv;
  int ld;
  bf16* out;
  int out_ld;
  int B, F, HW, heads, C;
  float scale_log2;
};

PT_DEVICE void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

PT_DEVICE uint32_t ld_pair(const bf16* p, bool ok) {  // two consecutive bf16 (4-byte aligned)
  return ok ? __ldg(reinterpret_cast<const unsigned int*>(p)) : 0u;
}

PT_DEVICE uint32_t ld_2rows(const bf16* p0, bool ok0, const bf16* p1, bool ok1) {  // one bf16 from each of two rows
  const uint32_t lo = ok0 ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p0)) : 0u;
  const uint32_t hi = ok1 ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p1)) : 0u;
  return lo | (hi << 16);
}

// MT = number of 16-row query tiles (F <= 16*MT); keys are padded to 16*MT as well.
template <int MT>
__global__ void __launch_bounds__(256) attn_temporal_kernel(const TAttnParams p) {
  griddep_launch();
  griddep_wait();
  constexpr int NT = 2 * MT;  // 8-key tiles
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.B * p.HW * p.heads;
  if (warp_global >= total) return;
  const int head = warp_global % p.heads;
  const int bs = warp_global / p.heads;
  const int s = bs % p.HW;
  const int b = bs / p.HW;
  const int g = lane >> 2;
  const int t = lane & 3;
  const size_t frame_stride = (size_t)p.HW * p.ld;
  const bf16* base = p.qkv + ((size_t)b * p.F * p.HW + s) * p.ld + head * 64;
  const bf16* qb = base;
  const bf16* kb = base + p.C;
  const bf16* vb = base + 2 * p.C;

  // S = Q K^T
  float sc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[mt][nt][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {  // 16 head-dims per step
    const int d0 = ks * 16 + 2 * t;
    uint32_t bk[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int key = nt * 8 + g;
      const bool ok = key < p.F;
      bk[nt][0] = ld_pair(kb + key * frame_stride + d0, ok);
      bk[nt][1] = ld_pair(kb + key * frame_stride + d0 + 8, ok);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      uint32_t a[4];
      a[0] = ld_pair(qb + r0 * frame_stride + d0, r0 < p.F);
      a[1] = ld_pair(qb + r1 * frame_stride + d0, r1 < p.F);
      a[2] = ld_pair(qb + r0 * frame_stride + d0 + 8, r0 < p.F);
      a[3] = ld_pair(qb + r1 * frame_stride + d0 + 8, r1 < p.F);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma16816(sc[mt][nt], a, bk[nt][0], bk[nt][1]);
    }
  }

  // softmax over keys (rows g and g+8 of each m-tile; a row lives in the 4 lanes sharing g)
  uint32_t pa[MT][MT][4];  // P as A fragments: [m-tile][16-key step][4]
  float inv_l[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int k0 = nt * 8 + 2 * t;
      if (k0 < p.F) { m0 = fmaxf(m0, sc[mt][nt][0]); m1 = fmaxf(m1, sc[mt][nt][2]); }
      if (k0 + 1 < p.F) { m0 = fmaxf(m0, sc[mt][nt][1]); m1 = fmaxf(m1, sc[mt][nt][3]); }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int k0 = nt * 8 + 2 * t;
      const float e0 = (k0 < p.F) ? exp2f((sc[mt][nt][0] - m0) * p.scale_log2) : 0.f;
      const float e1 = (k0 + 1 < p.F) ? exp2f((sc[mt][nt][1] - m0) * p.scale_log2) : 0.f;
      const float e2 = (k0 < p.F) ? exp2f((sc[mt][nt][2] - m1) * p.scale_log2) : 0.f;
      const float e3 = (k0 + 1 < p.F) ? exp2f((sc[mt][nt][3] - m1) * p.scale_log2) : 0.f;
      const uint32_t u01 = pack_bf16x2(e0, e1), u23 = pack_bf16x2(e2, e3);
      const float2 r01 = unpack_bf16x2(u01), r23 = unpack_bf16x2(u23);
      l0 += r01.x + r01.y;
      l1 += r23.x + r23.y;
      // C fragment of key tile nt -> A fragment of 16-key step nt/2: even tile -> a0/a1, odd tile -> a2/a3
      pa[mt][nt >> 1][(nt & 1) * 2 + 0] = u01;
      pa[mt][nt >> 1][(nt & 1) * 2 + 1] = u23;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    inv_l[mt][0] = 1.0f / l0;
    inv_l[mt][1] = 1.0f / l1;
  }

  // O = P V, 8 head-dims per n-tile
  bf16* ob = p.out + ((size_t)b * p.F * p.HW + s) * p.out_ld + head * 64;
  const size_t out_frame_stride = (size_t)p.HW * p.out_ld;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    float oc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) oc[mt][i] = 0.f;
    const int d = dn * 8 + g;
#pragma unroll
    for (int ks = 0; ks < MT; ++ks) {  // 16 keys per step
      const int k0 = ks * 16 + 2 * t;
      const uint32_t b0 = ld_2rows(vb + k0 * frame_stride + d, k0 < p.F, vb + (k0 + 1) * frame_stride + d, k0 + 1 < p.F);
      const uint32_t b1 = ld_2rows(vb + (k0 + 8) * frame_stride + d, k0 + 8 < p.F,
                                   vb + (k0 + 9) * frame_stride + d, k0 + 9 < p.F);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma16816(oc[mt], pa[mt][ks], b0, b1);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const int dc = dn * 8 + 2 * t;
      if (r0 < p.F)
        *reinterpret_cast<uint32_t*>(ob + r0 * out_frame_stride + dc) =
            pack_bf16x2(oc[mt][0] * inv_l[mt][0], oc[mt][1] * inv_l[mt][0]);
      if (r1 < p.F)
        *reinterpret_cast<uint32_t*>(ob + r1 * out_frame_stride + dc) =
            pack_bf16x2(oc[mt][2] * inv_l[mt][1], oc[mt][3] * inv_l[mt][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Staged variant: the same one-warp-per-problem arithmetic, but Q, K and V (F rows of 128 bytes each) are brought in with
// 16-byte cp.async into a warp-private, XOR-swizzled shared-memory tile and the fragments are read with ldmatrix; O leaves
// through the same tile as 16-byte stores.  The register-fragment kernel above issues 4-byte (Q, K, O) and 2-byte (V)
// global accesses, 64 load instructions per lane for 5.4 KB per warp, and measured 0.42 of the HBM copy rate; this one
// issues 11 loads and 4 stores per lane on whole 128-byte lines.  Needs 16-byte aligned rows (ld % 8 == 0).
// Shared memory: 3 tiles x 16*MT rows x 128 B per warp = 48 KB per CTA (8 warps at MT = 1, 4 warps at MT = 2).
// ---------------------------------------------------------------------------------------------------------
PT_DEVICE void ta_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
PT_DEVICE void ta_ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// byte address of 16-byte chunk `chunk` (0..7) of row `row` in a tile of 128-byte rows, chunks swizzled by (row & 7)
PT_DEVICE uint32_t ta_addr(uint32_t tile, int row, int chunk) {
  return tile + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

template <int MT>
__global__ void __launch_bounds__(MT == 1 ? 256 : 128) attn_temporal_staged_kernel(const TAttnParams p) {
  constexpr int NT = 2 * MT;
  constexpr int kRows = 16 * MT;
  constexpr int kWarps = MT == 1 ? 8 : 4;
  __shared__ __align__(128) uint8_t s_tiles[kWarps * 3 * kRows * 128];
  griddep_launch();
  griddep_wait();
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.B * p.HW * p.heads;
  if (warp_global >= total) return;
  const int head = warp_global % p.heads;
  const int bs = warp_global / p.heads;
  const int s = bs % p.HW;
  const int b = bs / p.HW;
  const int g = lane >> 2;
  const int t = lane & 3;
  const size_t frame_stride = (size_t)p.HW * p.ld;
  const bf16* base = p.qkv + ((size_t)b * p.F * p.HW + s) * p.ld + head * 64;
  const uint32_t tq = smem_u32(s_tiles) + (uint32_t)(threadIdx.x >> 5) * (3u * kRows * 128u);
  const uint32_t tk = tq + kRows * 128u;
  const uint32_t tv = tk + kRows * 128u;

  // ---- stage Q, K, V: chunk c = row * 8 + col, 32 chunks (4 rows) per warp instruction ----
  const int n_chunks = p.F * 8;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const bf16* src = base + m * p.C;
    const uint32_t tile = tq + (uint32_t)m * (kRows * 128u);
#pragma unroll
    for (int i = 0; i < kRows / 4; ++i) {
      const int c = lane + 32 * i;
      const int row = c >> 3, col = c & 7;
      if (c < n_chunks) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ta_addr(tile, row, col)),
                     "l"(src + (size_t)row * frame_stride + col * 8)
                     : "memory");
      } else {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ta_addr(tile, row, col)), "r"(0u) : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- S = Q K^T ----
  float sc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[mt][nt][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t bk[MT][4];  // per 16 keys: {b0, b1} of key tile 2*j, {b0, b1} of key tile 2*j + 1
#pragma unroll
    for (int j = 0; j < MT; ++j)
      ta_ldsm_x4(bk[j], ta_addr(tk, j * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t a[4];
      ta_ldsm_x4(a, ta_addr(tq, mt * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)));
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma16816(sc[mt][nt], a, bk[nt >> 1][(nt & 1) * 2], bk[nt >> 1][(nt & 1) * 2 + 1]);
    }
  }

  // ---- softmax over keys (identical to the register kernel) ----
  uint32_t pa[MT][MT][4];
  float inv_l[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int k0 = nt * 8 + 2 * t;
      if (k0 < p.F) { m0 = fmaxf(m0, sc[mt][nt][0]); m1 = fmaxf(m1, sc[mt][nt][2]); }
      if (k0 + 1 < p.F) { m0 = fmaxf(m0, sc[mt][nt][1]); m1 = fmaxf(m1, sc[mt][nt][3]); }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int k0 = nt * 8 + 2 * t;
      const float e0 = (k0 < p.F) ? exp2f((sc[mt][nt][0] - m0) * p.scale_log2) : 0.f;
      const float e1 = (k0 + 1 < p.F) ? exp2f((sc[mt][nt][1] - m0) * p.scale_log2) : 0.f;
      const float e2 = (k0 < p.F) ? exp2f((sc[mt][nt][2] - m1) * p.scale_log2) : 0.f;
      const float e3 = (k0 + 1 < p.F) ? exp2f((sc[mt][nt][3] - m1) * p.scale_log2) : 0.f;
      const uint32_t u01 = pack_bf16x2(e0, e1), u23 = pack_bf16x2(e2, e3);
      const float2 r01 = unpack_bf16x2(u01), r23 = unpack_bf16x2(u23);
      l0 += r01.x + r01.y;
      l1 += r23.x + r23.y;
      pa[mt][nt >> 1][(nt & 1) * 2 + 0] = u01;
      pa[mt][nt >> 1][(nt & 1) * 2 + 1] = u23;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    inv_l[mt][0] = 1.0f / l0;
    inv_l[mt][1] = 1.0f / l1;
  }

  // ---- O = P V, two 8-wide head-dim tiles per ldmatrix.trans; O staged over the Q tile ----
  __syncwarp();  // every lane is done reading Q before it is overwritten
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    float oc[2][MT][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) oc[h][mt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < MT; ++ks) {
      uint32_t bv[4];  // {b0, b1} of head-dim tile 2*dp, {b0, b1} of tile 2*dp + 1
      ta_ldsm_x4_trans(bv, ta_addr(tv, ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma16816(oc[0][mt], pa[mt][ks], bv[0], bv[1]);
        mma16816(oc[1][mt], pa[mt][ks], bv[2], bv[3]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const int dn = dp * 2 + h;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ta_addr(tq, r0, dn) + 4u * t),
                     "r"(pack_bf16x2(oc[h][mt][0] * inv_l[mt][0], oc[h][mt][1] * inv_l[mt][0]))
                     : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ta_addr(tq, r1, dn) + 4u * t),
                     "r"(pack_bf16x2(oc[h][mt][2] * inv_l[mt][1], oc[h][mt][3] * inv_l[mt][1]))
                     : "memory");
      }
  }
  __syncwarp();
  bf16* ob = p.out + ((size_t)b * p.F * p.HW + s) * p.out_ld + head * 64;
  const size_t out_frame_stride = (size_t)p.HW * p.out_ld;
#pragma unroll
  for (int i = 0; i < kRows / 4; ++i) {
    const int c = lane + 32 * i;
    const int row = c >> 3, col = c & 7;
    if (c < n_chunks) {
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(ta_addr(tq, row, col)));
      *reinterpret_cast<uint4*>(ob + (size_t)row * out_frame_stride + col * 8) = u;
    }
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_attention_temporal(const PtAttnTemporalArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->qkv != nullptr && a->out != nullptr, "pt_attention_temporal: null argument");
  PT_CHECK_ARG(a->B > 0 && a->F > 0 && a->HW > 0 && a->heads > 0 && a->C == a->heads * 64,
               "pt_attention_temporal: need C == heads*64 and a non-empty problem");
  PT_CHECK_ARG(a->F <= 32, "pt_attention_temporal: at most 32 frames");
  PT_CHECK_ARG(a->ld % 2 == 0 && a->out_ld % 2 == 0, "pt_attention_temporal: row strides must be even");
  TAttnParams p;
  p.qkv = reinterpret_cast<const bf16*>(a->qkv);
  p.ld = a->ld;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  p.B = a->B; p.F = a->F; p.HW = a->HW; p.heads = a->heads; p.C = a->C;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  const long long warps = (long long)a->B * a->HW * a->heads;
  static int staged_env = -1;  // PT_TATTN_STAGED=0: the register-fragment kernel
  if (staged_env < 0) {
    const char* e = getenv("PT_TATTN_STAGED");
    staged_env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool aligned = a->ld % 8 == 0 && a->out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->out) & 15) == 0;
  if (staged_env != 0 && aligned) {
    if (a->F <= 16)
      pt_launch(attn_temporal_staged_kernel<1>, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, 1, p);
    else
      pt_launch(attn_temporal_staged_kernel<2>, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, stream, 1, p);
    return pt_launched("pt_attention_temporal");
  }
  const int blocks = (int)((warps + 7) / 8);
  if (a->F <= 16)
    pt_launch(attn_temporal_kernel<1>, dim3(blocks), dim3(256), 0, stream, 1, p);
  else
    pt_launch(attn_temporal_kernel<2>, dim3(blocks), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_attention_temporal");
}
