// posetraj_b200 — GroupNorm(32)(+SiLU) and LayerNorm on token-major (NHWC) bf16 activations.
//
// HBM-bound glue of SURVEY.md §8a/§2.2: 152 GroupNorm and 161 LayerNorm calls per denoise step.
//   * GroupNorm statistics are per (image, group) for the spatial blocks (ResnetBlock2D, transformer `norm`,
//     conv_norm_out) and per (batch, group) across all F frames for TemporalResnetBlock (5-D input) — the same
//     kernels with a different "rows per statistics group".
//   * The apply kernel can read a channel concat of two tensors (up-block `cat([h, skip], 1)` is never
//     materialised) and can write the zero-haloed image layout the implicit-GEMM 3x3 conv consumes.
//   * LayerNorm optionally adds the per-frame position embedding first and also emits that sum
//     (`hidden_states_mix = hidden_states + emb`, models/modified_svd.py:196-197).
// Algorithmic bytes: GN stats pass reads numel*2 B, apply pass reads numel*2 B and writes numel*2 B (the second
// read is served from shared memory for the rows the CTA could keep resident and from the 126 MB L2 for the rest);
// LayerNorm reads and writes numel*2 B.
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

// ---------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) in ONE launch: every CTA reduces its slab of rows to per-group partial sums, publishes them,
// waits until the other CTAs of the same statistics group have published theirs (the grid is sized to be fully
// co-resident, so the wait cannot deadlock), folds all partials in a fixed order and normalises the slab it has just
// read.  The first `res_slots` rows of every thread are fetched with cp.async into shared memory (all of them in
// flight at once: ~90 KB per CTA without a register held) and stay there for the second pass; the rows that do not
// fit are streamed through registers and re-read from the L2.  HBM traffic: numel*2 B in + out rows*C*2 B out.
// Deterministic on purpose: no floating-point atomics, every reduction in a fixed order, so two runs of the same
// step are bit-identical (a 1-ulp wobble in a mean flips bf16 roundings downstream and decorrelates whole runs).
// Workspace: uint32 arrive[1024] | uint32 depart[1024] (zero before the first launch; the kernel re-arms them)
//            | double partials[num_stat*splits*64].
// ---------------------------------------------------------------------------------------------------------
struct GnParams {
  const bf16* x0;
  const bf16* x1;
  int c0, c1, ld0, ld1;
  int rows_per_stat;  // rows sharing statistics (H*W, or F*H*W for the temporal 5-D norm)
  int num_stat;       // number of statistics groups (B*F or B)
  int splits;         // CTAs per statistics group
  double* partials;   // [num_stat, splits, 32, 2]
  unsigned int* arrive;  // [num_stat]
  unsigned int* depart;  // [num_stat]
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  bf16* out;
  int out_ld;
  int halo;  // 1: out rows are the zero-haloed image space; an image is H x W with H*W dividing rows_per_stat
  int H, W;
  int mode;       // 0 fused, 1 statistics only (-> sums), 2 normalise only (<- sums, count)
  double* sums;   // [num_stat, 32, 2]
  double count;   // mode 2: elements per (statistics, norm) group over ALL ranks
  int n_peers;    // mode 2: > 0 = sum the per-rank sums of these peers (rank order) instead of reading `sums`
  const double* sums_peers[8];
  int res_slots;  // rows per thread kept resident in shared memory between the two passes
  int res_off;    // byte offset of the resident area in dynamic shared memory
  unsigned int mW, mH;  // 2^32 / W + 1, 2^32 / H + 1: row -> (image row, image) of the haloed output by multiply-high
};

PT_DEVICE void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
PT_DEVICE void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
PT_DEVICE uint4 lds_u4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}

// Phase 2 of gn_fused_kernel for one thread: normalise (+SiLU) its rows and write the output layout.  Slot i of the thread
// is row r0 + i * rpar of the statistics group; the first n_res slots are read back from shared memory, the rest from
// global memory (L2 hits).  Four rows per iteration without predicates, then a one-row tail.  The haloed output row of
// group row rk is rk + yg + img * (W + 1) with yg = rk / W, img = yg / H (one extra column per image row, one extra row
// per image), the two divisions as multiply-high by 2^32 / W, 2^32 / H (exact for rk * W < 2^32, checked by the host).
// The first version of this phase walked (x, y, img) with `while` loops and carried live / resident predicates per row:
// 157 instructions per 16-byte vector where the arithmetic needs ~60, on a kernel with 3.75 warps per scheduler
// (ncu: 61 % of the level-0 launch's samples, spread thin over fixed-latency stalls; profiles/r3_glue_kernels.md).
// kSilu: 0 none, 1 x * sigmoid(x) with ex2 + rcp, 2 h + h tanh(h) (sc / sh then carry the factor 1/2: h = x / 2)
template <int kSilu, bool kHalo>
PT_DEVICE void gn_apply_rows(const GnParams& p, const float (&sc)[8], const float (&sh)[8], const bf16* base, int ld,
                             bf16* out_base, int r0, int rpar, int n_my, int n_res, uint32_t res_u32, uint32_t res_stride) {
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int W1 = p.W + 1;
  auto emit = [&](const uint4 u, const int rk) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = fmaf(v[j], sc[j], sh[j]);
      if (kSilu == 1) v[j] = silu_f(v[j]);
      if (kSilu == 2) v[j] = silu_half_tanh(v[j]);
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    if (kHalo) {
      const int yg = p.W > 1 ? (int)__umulhi((unsigned)rk, p.mW) : rk;
      const int img = p.H > 1 ? (int)__umulhi((unsigned)yg, p.mH) : yg;
      bf16* dst = out_base + (size_t)(rk + yg + img * W1) * p.out_ld;
      stg_u4(dst, o);
      // the zero halo: one extra column right of every image row, one extra row below every image
      const bool last_x = rk - yg * p.W == p.W - 1;
      if (last_x) stg_u4(dst + p.out_ld, zero4);
      if (yg - img * p.H == p.H - 1) {
        stg_u4(dst + (size_t)W1 * p.out_ld, zero4);
        if (last_x) stg_u4(dst + (size_t)(W1 + 1) * p.out_ld, zero4);
      }
    } else {
      stg_u4(out_base + (size_t)rk * p.out_ld, o);
    }
  };
  int i = 0;
  for (; i + 4 <= n_res; i += 4) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = lds_u4(res_u32 + (uint32_t)(i + k) * res_stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(u[k], r0 + (i + k) * rpar);
  }
  for (; i < n_res; ++i) emit(lds_u4(res_u32 + (uint32_t)i * res_stride), r0 + i * rpar);
  for (; i + 4 <= n_my; i += 4) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = ldg_u4(base + (size_t)(r0 + (i + k) * rpar) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(u[k], r0 + (i + k) * rpar);
  }
  for (; i < n_my; ++i) emit(ldg_u4(base + (size_t)(r0 + i * rpar) * ld), r0 + i * rpar);
}

// kCluster: the CTAs of one statistics group form ONE thread-block cluster (<= 16 CTAs, every row resident in shared
// memory): the partial sums are exchanged through distributed shared memory behind the hardware cluster barrier instead of
// the global-memory rendezvous (atomic counter + polling), whose ~10 us of latency is the floor of the 107 GroupNorm
// launches on the small tensors of levels 1-3.  Opt-in (PT_GN_CLUSTER=1): measured slower, profiles/r2i_groupnorm_cluster.md.  Co-residency is guaranteed by the cluster
// launch, so this path makes no assumption about the rest of the GPU being empty.
template <bool kCluster>
__global__ void __launch_bounds__(512) gn_fused_kernel(const GnParams p) {
  extern __shared__ __align__(16) float s_red[];  // [rpar][C] sums, [rpar][C] squares, then [C] + [C] per-channel totals; then resident rows
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ double s_cl[64];  // kCluster: this CTA's per-group {sum, sum of squares}, read by the peers
  const int C = p.c0 + p.c1;
  const int cvec = C >> 3;            // threads along channels (8 channels each)
  const int rpar = blockDim.x / cvec; // row lanes
  const int tc = threadIdx.x % cvec;
  const int tr = threadIdx.x / cvec;
  const int stat = blockIdx.x / p.splits;
  const int split = blockIdx.x - stat * p.splits;
  const int cg = C >> 5;
  const int c = tc * 8;
  const bf16* src;
  int ld;
  if (c < p.c0) {
    src = p.x0 + c;
    ld = p.ld0;
  } else {
    src = p.x1 + (c - p.c0);
    ld = p.ld1;
  }
  const int rows_per_split = (p.rows_per_stat + p.splits - 1) / p.splits;
  const int r_begin = split * rows_per_split;
  const int r_end = min(p.rows_per_stat, r_begin + rows_per_split);
  const bf16* base = src + (size_t)stat * p.rows_per_stat * ld;

  // affine parameters of this thread's 8 channels: fetched now, used after the rendezvous
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c) + 1);
  // PDL: the activations are the previous kernel's output.  The NEXT grid may only be released once every CTA of
  // this one is resident (they wait for each other below): after the rendezvous, or at once when there is none.
  griddep_wait();
  if (p.mode == 2) griddep_launch();

  // ---------------- phase 1: partial statistics of this CTA's rows ----------------
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  // resident rows: slot i of this thread holds row r_begin + tr + i*rpar (a thread only ever reads its own slots, so
  // cp.async.wait_group is all the synchronisation needed)
  const int n_my = (r_end - r_begin - tr + rpar - 1) / rpar;  // rows of this thread (<= 0: none)
  const int n_res = p.mode == 0 ? min(n_my, p.res_slots) : 0;
  const uint32_t res_u32 = smem_u32(reinterpret_cast<uint8_t*>(s_red) + p.res_off) + threadIdx.x * 16u;
  const uint32_t res_stride = blockDim.x * 16u;
  for (int i = 0; i < n_res; ++i) cp_async_16(res_u32 + (uint32_t)i * res_stride, base + (size_t)(r_begin + tr + i * rpar) * ld);
  if (p.mode != 2) {
    int r = r_begin + tr + n_res * rpar;
    for (; r + 7 * rpar < r_end; r += 8 * rpar) {  // 8 independent 16-byte loads in flight per thread
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = ldg_u4(base + (size_t)(r + k * rpar) * ld);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y), cc = unpack_bf16x2(u[k].z), d = unpack_bf16x2(u[k].w);
        const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sum[j] += v[j];
          sq[j] = fmaf(v[j], v[j], sq[j]);
        }
      }
    }
    for (; r < r_end; r += rpar) {
      const uint4 u = ldg_u4(base + (size_t)r * ld);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum[j] += v[j];
        sq[j] = fmaf(v[j], v[j], sq[j]);
      }
    }
    if (n_res > 0) {
      cp_async_wait_all();
      for (int i = 0; i < n_res; ++i) {
        const uint4 u = lds_u4(res_u32 + (uint32_t)i * res_stride);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sum[j] += v[j];
          sq[j] = fmaf(v[j], v[j], sq[j]);
        }
      }
    }
  }
  if (p.mode != 2) {
  float* s_sum = s_red;
  float* s_sq = s_red + (size_t)rpar * C;
  float* c_sum = s_sq + (size_t)rpar * C;
  float* c_sq = c_sum + C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_sum[tr * C + c + j] = sum[j];
    s_sq[tr * C + c + j] = sq[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < rpar; ++i) {
      a += s_sum[i * C + ch];
      b += s_sq[i * C + ch];
    }
    c_sum[ch] = a;
    c_sq[ch] = b;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // this CTA's channels of a group in fp32 (a few thousand elements), fp64 only across CTAs: a chain of cg fp64 adds
    // on 32 threads was ~1 us of pure latency on the 1/64-rate fp64 pipe
    float a = 0.f, b = 0.f;
    for (int i = 0; i < cg; ++i) {
      a += c_sum[threadIdx.x * cg + i];
      b += c_sq[threadIdx.x * cg + i];
    }
    if constexpr (kCluster) {
      s_cl[2 * threadIdx.x] = (double)a;
      s_cl[2 * threadIdx.x + 1] = (double)b;
    } else {
      double* dst = p.partials + (((size_t)stat * p.splits + split) * 32 + threadIdx.x) * 2;
      dst[0] = (double)a;
      dst[1] = (double)b;
      __threadfence();
    }
  }
  __syncthreads();

  // ---------------- rendezvous of the CTAs of this statistics group ----------------
  if constexpr (kCluster) {
    cluster_sync_all();   // release / acquire at cluster scope: every peer's s_cl is visible
  } else {
    if (threadIdx.x == 0) {
      atomicAdd(&p.arrive[stat], 1u);
      unsigned int spins = 0;
      while (*reinterpret_cast<volatile unsigned int*>(&p.arrive[stat]) < (unsigned)p.splits) {
        __nanosleep(20);
        if (++spins > (1u << 25)) asm volatile("trap;");  // a sizing bug becomes an error, not a hung GPU
      }
      __threadfence();
    }
    __syncthreads();
  }
  griddep_launch();

  // ---------------- fixed-order fp64 fold of all partials (identical in every CTA of the group) ----------------
  double* s_part = reinterpret_cast<double*>(s_red);  // [nsl][64], reuses the reduction scratch (>= 4 KiB)
  int nsl = blockDim.x >> 6;
  if (nsl > 8) nsl = 8;
  if constexpr (kCluster) {
    nsl = 1;
    if (threadIdx.x < 64) {
      const uint32_t mine = smem_u32(&s_cl[threadIdx.x]);
      double acc = 0.0;
      for (int q = 0; q < p.splits; ++q) {   // rank order: identical bits in every CTA of the cluster
        double v;
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(map_to_cta(mine, (uint32_t)q)));
        acc += v;
      }
      s_part[threadIdx.x] = acc;
    }
    // nobody may leave (or overwrite s_cl in a later launch) while a peer still reads its partials
    cluster_sync_all();
  } else {
  if ((int)threadIdx.x < nsl * 64) {
    const int vi = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const double* srcp = p.partials + (size_t)stat * p.splits * 64 + vi;
    double acc = 0.0;
    int i = sl;
    for (; i + 3 * nsl < p.splits; i += 4 * nsl) {
      const double d0 = __ldcg(srcp + (size_t)i * 64), d1 = __ldcg(srcp + (size_t)(i + nsl) * 64);
      const double d2 = __ldcg(srcp + (size_t)(i + 2 * nsl) * 64), d3 = __ldcg(srcp + (size_t)(i + 3 * nsl) * 64);
      acc += d0; acc += d1; acc += d2; acc += d3;
    }
    for (; i < p.splits; i += nsl) acc += __ldcg(srcp + (size_t)i * 64);
    s_part[sl * 64 + vi] = acc;
  }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < nsl; ++sl) {
      a += s_part[sl * 64 + 2 * threadIdx.x];
      b += s_part[sl * 64 + 2 * threadIdx.x + 1];
    }
    if (p.mode == 1) {
      // statistics only: the folded sums of this rank's rows (one CTA per group writes them); the caller all-reduces
      if (split == 0) {
        p.sums[((size_t)stat * 32 + threadIdx.x) * 2] = a;
        p.sums[((size_t)stat * 32 + threadIdx.x) * 2 + 1] = b;
      }
    }
    const double cnt = (double)p.rows_per_stat * cg;
    const double m = a / cnt;
    double var = b / cnt - m * m;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = rsqrtf((float)var + p.eps);
  }
  __syncthreads();
  // every CTA has read the partials it needs: the last one to get here re-arms the counters for the next launch
  if constexpr (!kCluster) {
    if (threadIdx.x == 0) {
      if (atomicAdd(&p.depart[stat], 1u) == (unsigned)(p.splits - 1)) {
        p.arrive[stat] = 0u;
        p.depart[stat] = 0u;
      }
    }
  }

  if (p.mode == 1) return;
  } else {
    // normalise only: statistics were reduced across ranks by the caller
    if (threadIdx.x < 32) {
      double a = 0.0, b = 0.0;
      if (p.n_peers > 0) {
        for (int q = 0; q < p.n_peers; ++q) {   // fixed order: every rank computes bit-identical statistics
          a += __ldcg(p.sums_peers[q] + ((size_t)stat * 32 + threadIdx.x) * 2);       // L1 bypass: peer memory
          b += __ldcg(p.sums_peers[q] + ((size_t)stat * 32 + threadIdx.x) * 2 + 1);
        }
      } else {
        a = p.sums[((size_t)stat * 32 + threadIdx.x) * 2];
        b = p.sums[((size_t)stat * 32 + threadIdx.x) * 2 + 1];
      }
      const double m = a / p.count;
      double var = b / p.count - m * m;
      if (var < 0) var = 0;
      s_mean[threadIdx.x] = (float)m;
      s_rstd[threadIdx.x] = rsqrtf((float)var + p.eps);
    }
    __syncthreads();
  }
  // ---------------- phase 2: normalise (+SiLU) the same rows (L2 hits), write the output layout ----------------
  float sc[8], sh[8];
  {
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cg;
      sc[j] = ga[j] * s_rstd[g];
      sh[j] = be[j] - s_mean[g] * sc[j];
    }
  }
  const size_t out_rows_per_stat =
      p.halo ? (size_t)(p.rows_per_stat / (p.H * p.W)) * (size_t)((p.H + 1) * (p.W + 1)) : (size_t)p.rows_per_stat;
  bf16* out_base = p.out + (size_t)stat * out_rows_per_stat * p.out_ld + c;
  const int r0 = r_begin + tr;
  if (p.silu == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] *= 0.5f;
      sh[j] *= 0.5f;
    }
  }
#define PT_GN_APPLY(SILU, HALO) gn_apply_rows<SILU, HALO>(p, sc, sh, base, ld, out_base, r0, rpar, n_my, n_res, res_u32, res_stride)
  if (p.halo) {
    if (p.silu == 2) PT_GN_APPLY(2, true);
    else if (p.silu) PT_GN_APPLY(1, true);
    else PT_GN_APPLY(0, true);
  } else {
    if (p.silu == 2) PT_GN_APPLY(2, false);
    else if (p.silu) PT_GN_APPLY(1, false);
    else PT_GN_APPLY(0, false);
  }
#undef PT_GN_APPLY
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: G lanes per row (G = 8/16/32 so that every lane owns nvec/G 16-byte vectors), row kept in registers
// ---------------------------------------------------------------------------------------------------------
struct LnParams {
  const bf16* x;
  int ld;
  const float* gamma;
  const float* beta;
  float eps;
  bf16* out;
  int out_ld;
  int rows, C;
  const float* addvec;  // optional [F, C] fp32 added before normalising: frame = (row / hw) % F
  int hw, F;
  bf16* sum_out;        // optional: x + addvec (bf16), same ld as out
};

PT_DEVICE void unpack_row8(uint4 u, float* f) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

constexpr int kLnMaxVec = 8;  // vectors per lane: 8 x 8 channels x 32 lanes = 2048 channels at G = 32

// V = 16-byte vectors per lane (compile-time so that the row really lives in V*8 registers, not kLnMaxVec*8)
template <int G, int V>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  constexpr int kRowsPerWarp = 32 / G;
  const int lane = threadIdx.x & 31;
  const int l = lane % G;
  const int nvec = p.C >> 3;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.rows + kRowsPerWarp - 1) / kRowsPerWarp;
  int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  griddep_launch();
  griddep_wait();
  if (grp >= n_groups) return;
  // persistent warps: the loads of the NEXT row group are in flight while the current one is reduced and stored
  auto row_of = [&](int g) {
    const int r = g * kRowsPerWarp + lane / G;
    return r < p.rows ? r : p.rows - 1;  // keep the lanes in the shuffles; they just do not store
  };
  uint4 raw[V];
  {
    const bf16* src = p.x + (size_t)row_of(grp) * p.ld;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      raw[i] = (vi < nvec) ? ldg_nc_u4(src + vi * 8) : make_uint4(0, 0, 0, 0);
    }
  }
  for (; grp < n_groups; grp += warps_total) {
    const int row = row_of(grp);
    const bool live = grp * kRowsPerWarp + lane / G < p.rows;
    const float* av = nullptr;
    if (p.addvec != nullptr) av = p.addvec + (size_t)((row / p.hw) % p.F) * p.C;
    float v[V][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      if (vi < nvec) {
        const float2 a = unpack_bf16x2(raw[i].x), b = unpack_bf16x2(raw[i].y), c = unpack_bf16x2(raw[i].z), d = unpack_bf16x2(raw[i].w);
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
        v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
      }
    }
    // prefetch the next row group of this warp
    if (grp + warps_total < n_groups) {
      const bf16* nsrc = p.x + (size_t)row_of(grp + warps_total) * p.ld;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int vi = l + i * G;
        raw[i] = (vi < nvec) ? ldg_nc_u4(nsrc + vi * 8) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      if (vi < nvec) {
        if (av != nullptr) {
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(av + vi * 8));
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(av + vi * 8) + 1);
          v[i][0] += e0.x; v[i][1] += e0.y; v[i][2] += e0.z; v[i][3] += e0.w;
          v[i][4] += e1.x; v[i][5] += e1.y; v[i][6] += e1.z; v[i][7] += e1.w;
          if (p.sum_out != nullptr) {
            uint4 o;
            o.x = pack_bf16x2(v[i][0], v[i][1]);
            o.y = pack_bf16x2(v[i][2], v[i][3]);
            o.z = pack_bf16x2(v[i][4], v[i][5]);
            o.w = pack_bf16x2(v[i][6], v[i][7]);
            if (live) stg_u4(p.sum_out + (size_t)row * p.out_ld + vi * 8, o);
            // normalise exactly what the consumer of sum_out will see (bf16-rounded), like the reference does
            const float2 ra = unpack_bf16x2(o.x), rb = unpack_bf16x2(o.y), rc = unpack_bf16x2(o.z), rd = unpack_bf16x2(o.w);
            v[i][0] = ra.x; v[i][1] = ra.y; v[i][2] = rb.x; v[i][3] = rb.y;
            v[i][4] = rc.x; v[i][5] = rc.y; v[i][6] = rd.x; v[i][7] = rd.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)p.C + p.eps);
    if (live) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int vi = l + i * G;
        if (vi < nvec) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8) + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, gg[j], bb[j]);
          uint4 u;
          u.x = pack_bf16x2(o[0], o[1]);
          u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]);
          u.w = pack_bf16x2(o[6], o[7]);
          stg_u4(p.out + (size_t)row * p.out_ld + vi * 8, u);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm, packed variant: the row stays in registers as the bf16 it was loaded as (V x 16 bytes per lane instead of
// V x 8 floats) and is unpacked once per pass (sum, centred squares, output).  Same arithmetic and the same order of every
// reduction as layernorm_kernel, so the two are bit-identical; what changes is the register budget: 80 instead of 117,
// i.e. 3 resident CTAs per SM instead of 2 (level-0 launches 28.7 -> 27.6 us, with the position add 46.2 -> 43.0 us,
// class 2.49 -> 2.35 ms per step).  Measured and NOT kept (profiles/r3_glue_kernels.md): a cp.async prefetch ring with
// 120 KB in flight per SM (slower: 32.8 us), 4 CTAs per SM at 64 registers (slower), one-pass statistics (26.7 us, not
// worth a second set of numerics), predicate-free instantiations (spills at the 80-register cap: 34.8 us).
// Handles the plain form and addvec + sum_out (the temporal norm_in: the row is re-packed after the add, exactly the
// bf16 rounding the consumer of sum_out sees); addvec without sum_out keeps the fp32 kernel above.
// ---------------------------------------------------------------------------------------------------------
// kFull (needs C == 8 * G * V, every width of this model): no lane has a vector slot to predicate off (-24 % instructions)
// and gamma | beta are staged in shared memory once per CTA.  Pays on the position-add form (45.1 -> 36.9 us at level 0);
// on the plain form it changes nothing at level 0 and its prologue costs 2 us on the small tensors, so the host uses it
// for the position-add launches only.
template <int G, int V, bool kFull>
__global__ void __launch_bounds__(256, (V <= 5 ? 3 : 2)) layernorm_packed_kernel(const LnParams p) {
  constexpr int kRowsPerWarp = 32 / G;
  const int lane = threadIdx.x & 31;
  const int l = lane % G;
  const int nvec = p.C >> 3;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.rows + kRowsPerWarp - 1) / kRowsPerWarp;
  int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // gamma | beta staged once per CTA: as global loads in the output pass they sat behind the statistics in program order
  // (no registers left to hoist them) and their latency was exposed once per vector (ncu: 44 % long-scoreboard stalls)
  extern __shared__ __align__(16) float s_gb[];
  if (kFull) {
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      s_gb[i] = __ldg(p.gamma + i);
      s_gb[p.C + i] = __ldg(p.beta + i);
    }
    __syncthreads();
  }
  griddep_launch();
  griddep_wait();
  if (grp >= n_groups) return;
  auto row_of = [&](int g) {
    const int r = g * kRowsPerWarp + lane / G;
    return r < p.rows ? r : p.rows - 1;
  };
  uint4 nxt[V];
  {
    const bf16* src = p.x + (size_t)row_of(grp) * p.ld;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      nxt[i] = (kFull || vi < nvec) ? ldg_nc_u4(src + vi * 8) : make_uint4(0, 0, 0, 0);
    }
  }
  for (; grp < n_groups; grp += warps_total) {
    const int row = row_of(grp);
    const bool live = grp * kRowsPerWarp + lane / G < p.rows;
    uint4 cur[V];
#pragma unroll
    for (int i = 0; i < V; ++i) cur[i] = nxt[i];
    if (grp + warps_total < n_groups) {
      const bf16* nsrc = p.x + (size_t)row_of(grp + warps_total) * p.ld;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int vi = l + i * G;
        nxt[i] = (kFull || vi < nvec) ? ldg_nc_u4(nsrc + vi * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    float s = 0.f;
    if (p.addvec != nullptr) {
      const float* av = p.addvec + (size_t)((row / p.hw) % p.F) * p.C;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int vi = l + i * G;
        if (kFull || vi < nvec) {
          float v[8];
          unpack_row8(cur[i], v);
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(av + vi * 8));
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(av + vi * 8) + 1);
          v[0] += e0.x; v[1] += e0.y; v[2] += e0.z; v[3] += e0.w;
          v[4] += e1.x; v[5] += e1.y; v[6] += e1.z; v[7] += e1.w;
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]);
          o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]);
          o.w = pack_bf16x2(v[6], v[7]);
          if (live) stg_u4(p.sum_out + (size_t)row * p.out_ld + vi * 8, o);
          cur[i] = o;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      if (kFull || vi < nvec) {
        float v[8];
        unpack_row8(cur[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int vi = l + i * G;
      if (kFull || vi < nvec) {
        float v[8];
        unpack_row8(cur[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)p.C + p.eps);
    if (live) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int vi = l + i * G;
        if (kFull || vi < nvec) {
          float4 g0, g1, b0, b1;
          if (kFull) {
            g0 = *reinterpret_cast<const float4*>(s_gb + vi * 8);
            g1 = *reinterpret_cast<const float4*>(s_gb + vi * 8 + 4);
            b0 = *reinterpret_cast<const float4*>(s_gb + p.C + vi * 8);
            b1 = *reinterpret_cast<const float4*>(s_gb + p.C + vi * 8 + 4);
          } else {
            g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8));
            g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8) + 1);
            b0 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8));
            b1 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8) + 1);
          }
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float v[8], o[8];
          unpack_row8(cur[i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf((v[j] - mean) * rstd, gg[j], bb[j]);
          uint4 u;
          u.x = pack_bf16x2(o[0], o[1]);
          u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]);
          u.w = pack_bf16x2(o[6], o[7]);
          stg_u4(p.out + (size_t)row * p.out_ld + vi * 8, u);
        }
      }
    }
  }
}

}  // namespace pt

using namespace pt;

static int gn_block_threads(int C) {
  const int cvec = C / 8;
  int rpar = 256 / cvec;
  if (rpar < 1) rpar = 1;
  return cvec * rpar;
}

static int gn_red_bytes(int C) {
  const int threads = gn_block_threads(C);
  const int rpar = threads / (C / 8);
  size_t b = sizeof(float) * ((size_t)2 * rpar * C + 2 * C);
  if (b < 4096) b = 4096;
  return (int)b;
}

// CTAs per statistics group.  The whole grid must be co-resident (the CTAs of a group wait for each other), so it
// is capped by what the occupancy calculator says fits on the device at once (registers / threads; the shared
// memory left over is then divided between the resident CTAs, see gn_res_slots).
static int gn_ctas_per_sm(int C) {
  const int threads = gn_block_threads(C);
  const int smem = gn_red_bytes(C);
  static int cached_threads[PT_MAX_DEVICES] = {0}, cached_smem[PT_MAX_DEVICES] = {0}, cached_blocks[PT_MAX_DEVICES] = {0};
  const int d = pt_device_slot();
  if (cached_threads[d] != threads || cached_smem[d] != smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gn_fused_kernel<false>, threads, smem) != cudaSuccess || nb < 1) nb = 1;
    cached_threads[d] = threads;
    cached_smem[d] = smem;
    cached_blocks[d] = nb;
  }
  return cached_blocks[d] < 4 ? cached_blocks[d] : 4;
}

constexpr int kGnSmemPerSm = 216 * 1024;  // of 228 KB: 1 KB per CTA is reserved by the system, static smem, margin

// rows per thread that can stay in shared memory between the statistics and the normalisation pass
static int gn_res_slots(int C, int per_sm, int rows_per_split) {
  static int env_res = -2;  // experiment knob: PT_GN_RESIDENT=0 disables the resident rows
  if (env_res == -2) {
    const char* e = getenv("PT_GN_RESIDENT");
    env_res = e ? atoi(e) : -1;
  }
  if (env_res == 0) return 0;
  const int threads = gn_block_threads(C);
  const int rpar = threads / (C / 8);
  const int avail = kGnSmemPerSm / per_sm - gn_red_bytes(C);
  int slots = avail > 0 ? avail / (threads * 16) : 0;
  const int need = (rows_per_split + rpar - 1) / rpar;
  return slots < need ? slots : need;
}

static int gn_splits(int num_stat, int rows_per_stat, int C) {
  const int threads = gn_block_threads(C);
  const int per_sm = gn_ctas_per_sm(C);
  const int capacity = pt_num_sms() * per_sm;
  int splits = capacity / num_stat;
  const int rpar = threads / (C / 8);
  const int max_splits = (rows_per_stat + rpar * 8 - 1) / (rpar * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

extern "C" int64_t pt_groupnorm_workspace_bytes(int32_t num_stat, int32_t rows_per_stat, int32_t channels) {
  if (num_stat < 1 || num_stat > 1024 || rows_per_stat < 1 || channels < 32 || channels % 8) return -1;
  // splits <= 4 CTAs per SM / num_stat, independent of the occupancy query (which needs a device)
  const int64_t max_ctas = (int64_t)pt_num_sms() * 4 + num_stat;
  return 8192 + (int64_t)sizeof(double) * 64 * max_ctas;
}

extern "C" int pt_groupnorm(const PtGroupNormArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x0 != nullptr && (a->out != nullptr || a->mode == 1) && a->stats != nullptr && a->gamma && a->beta,
               "pt_groupnorm: null argument");
  const int C = a->c0 + a->c1;
  PT_CHECK_ARG(a->c0 > 0 && a->c0 % 8 == 0 && a->c1 % 8 == 0 && C % 32 == 0 && C / 8 <= 512,
               "pt_groupnorm: channels must be multiples of 8 (sources) and 32 (total), <= 4096");
  PT_CHECK_ARG(a->c1 == 0 || a->x1 != nullptr, "pt_groupnorm: c1 > 0 without x1");
  PT_CHECK_ARG(a->rows_per_stat > 0 && a->num_stat > 0 && a->num_stat <= 1024, "pt_groupnorm: need 1..1024 statistics groups");
  PT_CHECK_ARG(!a->halo || (a->H > 0 && a->W > 0 && a->rows_per_stat % (a->H * a->W) == 0),
               "pt_groupnorm: halo output needs H*W dividing rows_per_stat");
  PT_CHECK_ARG(a->num_stat <= pt_num_sms(), "pt_groupnorm: more statistics groups than SMs (co-resident grid)");
  const int threads = gn_block_threads(C);
  const int splits = gn_splits(a->num_stat, a->rows_per_stat, C);
  GnParams p;
  p.x0 = reinterpret_cast<const bf16*>(a->x0);
  p.x1 = reinterpret_cast<const bf16*>(a->x1);
  p.c0 = a->c0; p.c1 = a->c1; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.rows_per_stat = a->rows_per_stat;
  p.num_stat = a->num_stat;
  p.splits = splits;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->stats);
  p.arrive = reinterpret_cast<unsigned int*>(ws);
  p.depart = reinterpret_cast<unsigned int*>(ws + 4096);
  p.partials = reinterpret_cast<double*>(ws + 8192);
  p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
  {
    static int silu_tanh_env = -1;  // PT_GN_SILU_TANH=0: x * sigmoid(x) with ex2 + rcp (two MUFU per element) instead of h + h tanh(h)
    if (silu_tanh_env < 0) {
      const char* e = getenv("PT_GN_SILU_TANH");
      silu_tanh_env = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (p.silu != 0 && silu_tanh_env != 0) p.silu = 2;
  }
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  p.halo = a->halo; p.H = a->H > 0 ? a->H : 1; p.W = a->W > 0 ? a->W : 1;
  p.mW = p.W > 1 ? 0xFFFFFFFFu / (unsigned)p.W + 1u : 0u;
  p.mH = p.H > 1 ? 0xFFFFFFFFu / (unsigned)p.H + 1u : 0u;
  PT_CHECK_ARG(!a->halo || (long long)a->rows_per_stat * p.W < (1ll << 32), "pt_groupnorm: rows_per_stat * W must stay below 2^32");
  p.mode = a->mode; p.sums = a->sums; p.count = a->count;
  p.n_peers = a->mode == 2 ? a->n_peers : 0;
  PT_CHECK_ARG(p.n_peers >= 0 && p.n_peers <= 8, "pt_groupnorm: at most 8 peers");
  for (int q = 0; q < 8; ++q) p.sums_peers[q] = q < p.n_peers ? a->sums_peers[q] : nullptr;
  for (int q = 0; q < p.n_peers; ++q) PT_CHECK_ARG(a->sums_peers[q] != nullptr, "pt_groupnorm: null peer sums");
  PT_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "pt_groupnorm: mode must be 0, 1 or 2");
  PT_CHECK_ARG(a->mode == 0 || a->sums != nullptr || (a->mode == 2 && a->n_peers > 0), "pt_groupnorm: modes 1/2 need `sums`");
  PT_CHECK_ARG(a->mode != 2 || a->count > 0, "pt_groupnorm: mode 2 needs `count`");
  static bool attr_set[PT_MAX_DEVICES] = {false};  // cudaFuncSetAttribute is per device
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(gn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGnSmemPerSm);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGnSmemPerSm);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_fused_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return pt_fail(e, "pt_groupnorm: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  // Cluster path (mode 0): the smallest cluster size in {1, 2, 4, 8, 16} that keeps EVERY row of its CTAs resident in
  // shared memory, provided the whole grid fits the GPU in one wave (otherwise — the 25-100 MB tensors of level 0 and
  // the 5-D temporal statistics — the co-resident global rendezvous below is the better schedule).
  {
    // Measured on B200 (profiles/r2i_groupnorm_cluster.md): parity green, but the class got SLOWER in the captured step
    // (4.59 -> 5.11 ms for the 152 launches): cluster scheduling costs more than the polled counter it replaces.
    // Kept behind PT_GN_CLUSTER=1 as an experiment; the default is the co-resident rendezvous below.
    static int env_cl = -2;
    if (env_cl == -2) {
      const char* e = getenv("PT_GN_CLUSTER");
      env_cl = e ? atoi(e) : 0;
    }
    const int rpar = threads / (C / 8);
    if (a->mode == 0 && env_cl != 0) {
      for (int cs = 1; cs <= 16; cs *= 2) {
        const int rps = (a->rows_per_stat + cs - 1) / cs;
        const int need = (rps + rpar - 1) / rpar;
        const size_t smem = (size_t)gn_red_bytes(C) + (size_t)need * threads * 16;
        if (smem > (size_t)kGnSmemPerSm) continue;
        const int per_sm = (int)((size_t)kGnSmemPerSm / smem) < 4 ? (int)((size_t)kGnSmemPerSm / smem) : 4;
        if ((long long)a->num_stat * cs > (long long)pt_num_sms() * per_sm) break;
        if (cs > 1 && (a->num_stat * cs) % cs != 0) break;
        GnParams pc = p;
        pc.splits = cs;
        pc.res_slots = need;
        pc.res_off = gn_red_bytes(C);
        cudaError_t e = pt_launch(gn_fused_kernel<true>, dim3(a->num_stat * cs), dim3(threads), smem, (void*)stream, cs, pc);
        if (e != cudaSuccess) return pt_fail(e, "pt_groupnorm: cluster launch");
        return pt_launched("pt_groupnorm");
      }
    }
  }
  const int rows_per_split = (a->rows_per_stat + splits - 1) / splits;
  p.res_slots = a->mode == 0 ? gn_res_slots(C, gn_ctas_per_sm(C), rows_per_split) : 0;
  p.res_off = gn_red_bytes(C);
  const size_t smem_bytes = (size_t)p.res_off + (size_t)p.res_slots * threads * 16;
  // The CTAs of a statistics group wait for each other inside the kernel.  Launched COOPERATIVELY the driver either places
  // the whole grid at once or fails the launch (e.g. a second stream / process holding SMs): a clean error instead of CTAs
  // spinning on peers that cannot be scheduled.  PT_GN_COOP=0 restores the plain launch (profiles/r2n_gn_cooperative.md).
  static int env_coop = -2;
  if (env_coop == -2) {
    const char* e = getenv("PT_GN_COOP");
    env_coop = e ? atoi(e) : 1;
  }
  if (env_coop != 0 && a->mode != 2 && splits > 1) pt_next_launch_cooperative() = true;
  cudaError_t le = pt_launch(gn_fused_kernel<false>, dim3(a->num_stat * splits), dim3(threads), smem_bytes, (void*)stream, 1, p);
  if (le != cudaSuccess) return pt_fail(le, "pt_groupnorm: launch (cooperative: the grid must be co-resident)");
  return pt_launched("pt_groupnorm");
}

extern "C" int pt_layernorm(const PtLayerNormArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->out && a->gamma && a->beta, "pt_layernorm: null argument");
  PT_CHECK_ARG(a->C % 8 == 0 && a->C >= 8 && a->C <= kLnMaxVec * 256, "pt_layernorm: C must be a multiple of 8, <= 2048");
  PT_CHECK_ARG(a->rows > 0, "pt_layernorm: empty problem");
  PT_CHECK_ARG(a->addvec == nullptr || (a->hw > 0 && a->F > 0), "pt_layernorm: addvec needs hw and F");
  LnParams p;
  p.x = reinterpret_cast<const bf16*>(a->x);
  p.ld = a->ld;
  p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  p.rows = a->rows; p.C = a->C;
  p.addvec = a->addvec; p.hw = a->hw > 0 ? a->hw : 1; p.F = a->F > 0 ? a->F : 1;
  p.sum_out = reinterpret_cast<bf16*>(a->sum_out);
  // lanes per row: the widest sub-warp group that divides the vector count evenly with <= 8 vectors per lane
  const int nvec = a->C / 8;
  int G = 32;
  for (int g : {32, 16, 8}) {
    if (nvec % g == 0 && nvec / g <= kLnMaxVec) { G = g; break; }
  }
  if (nvec < 8) G = 8;
  const int V = (nvec + G - 1) / G;
  const int rows_per_block = 8 * (32 / G);
  int blocks = (a->rows + rows_per_block - 1) / rows_per_block;
  cudaStream_t st = (cudaStream_t)stream;
  // packed variant (3 resident CTAs per SM at V <= 5) unless PT_LN_PACKED=0 or the fp32-add form (addvec without sum_out)
  static int packed_env = -1;
  if (packed_env < 0) {
    const char* e = getenv("PT_LN_PACKED");
    packed_env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool packed = packed_env != 0 && (a->addvec == nullptr || a->sum_out != nullptr);
  // persistent grid: as many CTAs as are co-resident (2 per SM for the fp32-row kernel at 117 registers, 3 for the
  // packed one at V <= 5); each warp strides over row groups with the next group's loads in flight
  static int full_env = -1;  // PT_LN_FULL=0: keep the per-vector predicates even when C == 8 * G * V
  if (full_env < 0) {
    const char* e = getenv("PT_LN_FULL");
    full_env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool full = full_env != 0 && nvec == G * V && a->addvec != nullptr && a->rows >= 16384;  // levels 0 / 1
  const int per_sm = packed ? (V <= 5 ? 3 : 2) : 2;
  const int max_blocks = pt_num_sms() * per_sm;
  if (blocks > max_blocks) blocks = max_blocks;
#define PT_LN_LAUNCH(GG, VV)                                                                        \
  do {                                                                                              \
    if (packed && full) pt_launch(layernorm_packed_kernel<GG, VV, true>, dim3(blocks), dim3(256), (size_t)a->C * 8, (void*)st, 1, p); \
    else if (packed) pt_launch(layernorm_packed_kernel<GG, VV, false>, dim3(blocks), dim3(256), 0, (void*)st, 1, p); \
    else pt_launch(layernorm_kernel<GG, VV>, dim3(blocks), dim3(256), 0, (void*)st, 1, p);          \
  } while (0)
#define PT_LN_G(GG)                                   \
  switch (V) {                                        \
    case 1: PT_LN_LAUNCH(GG, 1); break;               \
    case 2: PT_LN_LAUNCH(GG, 2); break;               \
    case 3: PT_LN_LAUNCH(GG, 3); break;               \
    case 4: PT_LN_LAUNCH(GG, 4); break;               \
    case 5: PT_LN_LAUNCH(GG, 5); break;               \
    case 6: PT_LN_LAUNCH(GG, 6); break;               \
    case 7: PT_LN_LAUNCH(GG, 7); break;               \
    default: PT_LN_LAUNCH(GG, 8); break;              \
  }
  if (G == 32) { PT_LN_G(32) } else if (G == 16) { PT_LN_G(16) } else { PT_LN_G(8) }
#undef PT_LN_G
#undef PT_LN_LAUNCH
  return pt_launched("pt_layernorm");
}
