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
// read normally hits the 126 MB L2); LayerNorm reads and writes numel*2 B.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

// ---------------------------------------------------------------------------------------------------------
// GroupNorm statistics: per-CTA partial sum / sum of squares per (stat group s, split, norm group g); the CTA that
// finishes last for a statistics group (ticket counter) folds the partials IN A FIXED ORDER into mean / rstd.
// Deterministic on purpose: no floating-point atomics anywhere, so two runs of the same step are bit-identical
// (a 1-ulp wobble in a mean flips bf16 roundings downstream and decorrelates whole runs).
// Workspace layout: uint32 tickets[<= 1024] in a reserved first 4 KiB (zero before the first launch; the kernel
// re-arms them, and no other region ever overlaps them even when differently shaped problems share one workspace)
// | float mean_rstd[num_stat*64] | double partials[num_stat*splits*64].
// ---------------------------------------------------------------------------------------------------------
struct GnStatsParams {
  const bf16* x0;
  const bf16* x1;
  int c0, c1, ld0, ld1;
  int rows_per_stat;  // rows sharing statistics (H*W, or F*H*W for the temporal 5-D norm)
  int num_stat;       // number of statistics groups (B*F or B)
  int splits;         // CTAs per statistics group
  double* partials;   // [num_stat, splits, 32, 2]
  float* mean_rstd;   // [num_stat, 32, 2]
  unsigned int* tickets;  // [num_stat]
  float eps;
};

__global__ void __launch_bounds__(512) gn_stats_kernel(const GnStatsParams p) {
  extern __shared__ float s_red[];  // [rpar][C] sums, [rpar][C] squares, then [C] + [C] per-channel totals
  __shared__ unsigned int s_ticket;
  const int C = p.c0 + p.c1;
  const int cvec = C >> 3;            // threads along channels (8 channels each)
  const int rpar = blockDim.x / cvec; // row lanes
  const int tc = threadIdx.x % cvec;
  const int tr = threadIdx.x / cvec;
  const int stat = blockIdx.x / p.splits;
  const int split = blockIdx.x - stat * p.splits;
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
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  const bf16* base = src + (size_t)stat * p.rows_per_stat * ld;
  int r = r_begin + tr;
  // 4 independent 16-byte loads in flight per thread
  for (; r + 3 * rpar < r_end; r += 4 * rpar) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = ldg_nc_u4(base + (size_t)(r + k * rpar) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
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
    const uint4 u = ldg_nc_u4(base + (size_t)r * ld);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sum[j] += v[j];
      sq[j] = fmaf(v[j], v[j], sq[j]);
    }
  }
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
  const int cg = C >> 5;
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < cg; ++i) {
      a += (double)c_sum[threadIdx.x * cg + i];
      b += (double)c_sq[threadIdx.x * cg + i];
    }
    double* dst = p.partials + (((size_t)stat * p.splits + split) * 32 + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = b;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[stat], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)(p.splits - 1)) return;
  // last CTA of this statistics group: fixed-order fp64 fold of all partials -> mean, rstd.  The fold is spread
  // over blockDim/64 slices with 4 loads in flight each (a serial chain of up to ~300 L2 round trips otherwise).
  __threadfence();
  double* s_part = reinterpret_cast<double*>(s_red);  // [nsl][64], reuses the reduction scratch (>= 4 KiB)
  int nsl = blockDim.x >> 6;
  if (nsl > 8) nsl = 8;
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
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < nsl; ++sl) {
      a += s_part[sl * 64 + 2 * threadIdx.x];
      b += s_part[sl * 64 + 2 * threadIdx.x + 1];
    }
    const double cnt = (double)p.rows_per_stat * cg;
    const double m = a / cnt;
    double var = b / cnt - m * m;
    if (var < 0) var = 0;
    p.mean_rstd[((size_t)stat * 32 + threadIdx.x) * 2] = (float)m;
    p.mean_rstd[((size_t)stat * 32 + threadIdx.x) * 2 + 1] = rsqrtf((float)var + p.eps);
    if (threadIdx.x == 0) p.tickets[stat] = 0u;  // re-arm for the next launch
  }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU), optional concat input, optional zero-haloed output
// ---------------------------------------------------------------------------------------------------------
struct GnApplyParams {
  const bf16* x0;
  const bf16* x1;
  int c0, c1, ld0, ld1;
  int rows_per_stat, num_stat, splits;
  const float* mean_rstd;  // [num_stat, 32, 2] from gn_stats_kernel
  const float* gamma;
  const float* beta;
  int silu;
  bf16* out;
  int out_ld;
  int halo;  // 1: out rows are the zero-haloed image space; an image is H x W with H*W dividing rows_per_stat
  int H, W;
};

__global__ void __launch_bounds__(512) gn_apply_kernel(const GnApplyParams p) {
  const int C = p.c0 + p.c1;
  const int stat = blockIdx.x / p.splits;
  const int split = blockIdx.x - stat * p.splits;
  const int cg = C >> 5;
  const int cvec = C >> 3;
  const int rpar = blockDim.x / cvec;
  const int tc = threadIdx.x % cvec;
  const int tr = threadIdx.x / cvec;
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
  // this thread's 8 channels: scale / shift straight from the folded statistics
  float sc[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c) + 1);
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cg;
      const float2 mr = __ldg(reinterpret_cast<const float2*>(p.mean_rstd) + (size_t)stat * 32 + g);
      sc[j] = ga[j] * mr.y;
      sh[j] = be[j] - mr.x * sc[j];
    }
  }
  // iterate over OUTPUT rows of this statistics group (haloed space if requested), 4 rows in flight per thread
  const int HW = p.H * p.W;
  const int W1 = p.W + 1, H1 = p.H + 1;
  const int imgs_per_stat = p.halo ? p.rows_per_stat / HW : 1;
  const int P = H1 * W1;
  const int out_rows = p.halo ? imgs_per_stat * P : p.rows_per_stat;
  const int rows_per_split = (out_rows + p.splits - 1) / p.splits;
  const int r_begin = split * rows_per_split;
  const int r_end = min(out_rows, r_begin + rows_per_split);
  const bf16* in_base = src + (size_t)stat * p.rows_per_stat * ld;
  bf16* out_base = p.out + (size_t)stat * out_rows * p.out_ld + c;
  int r = r_begin + tr;
  int img = 0, y = 0, x = 0;
  if (p.halo) {
    img = r / P;
    const int rem = r - img * P;
    y = rem / W1;
    x = rem - y * W1;
  }
  for (; r < r_end; r += 4 * rpar) {
    uint4 u[4];
    bool live[4], pad[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rk = r + k * rpar;
      live[k] = rk < r_end;
      pad[k] = false;
      long long in_row = rk;
      if (p.halo) {
        pad[k] = (y == p.H) || (x == p.W);
        in_row = (long long)img * HW + y * p.W + x;
        x += rpar;
        while (x >= W1) { x -= W1; ++y; }
        while (y >= H1) { y -= H1; ++img; }
      }
      u[k] = (live[k] && !pad[k]) ? ldg_nc_u4(in_base + (size_t)in_row * ld) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!live[k]) continue;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (!pad[k]) {
        const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y), cc = unpack_bf16x2(u[k].z), d = unpack_bf16x2(u[k].w);
        float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], sc[j], sh[j]);
          if (p.silu) v[j] = silu_f(v[j]);
        }
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
      }
      stg_u4(out_base + (size_t)(r + k * rpar) * p.out_ld, o);
    }
  }
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

constexpr int kLnMaxVec = 8;  // vectors per lane: 8 x 8 channels x 32 lanes = 2048 channels at G = 32

template <int G>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  constexpr int kRowsPerWarp = 32 / G;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int l = lane % G;
  int row = gwarp * kRowsPerWarp + lane / G;
  const bool live = row < p.rows;
  if (!live) row = p.rows - 1;  // keep the lanes in the shuffles; they just do not store
  const int nvec = p.C >> 3;
  const bf16* src = p.x + (size_t)row * p.ld;
  const float* av = nullptr;
  if (p.addvec != nullptr) av = p.addvec + (size_t)((row / p.hw) % p.F) * p.C;
  uint4 raw[kLnMaxVec];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vi = l + i * G;
    raw[i] = (vi < nvec) ? ldg_nc_u4(src + vi * 8) : make_uint4(0, 0, 0, 0);
  }
  float v[kLnMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vi = l + i * G;
    if (vi < nvec) {
      const float2 a = unpack_bf16x2(raw[i].x), b = unpack_bf16x2(raw[i].y), c = unpack_bf16x2(raw[i].z), d = unpack_bf16x2(raw[i].w);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
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
  for (int i = 0; i < kLnMaxVec; ++i) {
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
  if (!live) return;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
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

}  // namespace pt

using namespace pt;

static int gn_block_threads(int C) {
  const int cvec = C / 8;
  int rpar = 256 / cvec;
  if (rpar < 1) rpar = 1;
  return cvec * rpar;
}

static int gn_splits(int num_stat, int rows_per_stat, int C) {
  const int threads = gn_block_threads(C);
  int splits = (pt_num_sms() * 4 + num_stat - 1) / num_stat;
  const int rpar = threads / (C / 8);
  const int max_splits = (rows_per_stat + rpar * 8 - 1) / (rpar * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

static size_t gn_partials_bytes(int num_stat, int splits) { return sizeof(double) * 64 * (size_t)num_stat * splits; }

extern "C" int64_t pt_groupnorm_workspace_bytes(int32_t num_stat, int32_t rows_per_stat, int32_t channels) {
  if (num_stat < 1 || rows_per_stat < 1 || channels < 32 || channels % 8) return -1;
  if (num_stat > 1024) return -1;
  return (int64_t)(4096 + sizeof(float) * 64 * (size_t)num_stat +
                   gn_partials_bytes(num_stat, gn_splits(num_stat, rows_per_stat, channels)));
}

extern "C" int pt_groupnorm(const PtGroupNormArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x0 != nullptr && a->out != nullptr && a->stats != nullptr && a->gamma && a->beta,
               "pt_groupnorm: null argument");
  const int C = a->c0 + a->c1;
  PT_CHECK_ARG(a->c0 > 0 && a->c0 % 8 == 0 && a->c1 % 8 == 0 && C % 32 == 0 && C / 8 <= 512,
               "pt_groupnorm: channels must be multiples of 8 (sources) and 32 (total), <= 4096");
  PT_CHECK_ARG(a->c1 == 0 || a->x1 != nullptr, "pt_groupnorm: c1 > 0 without x1");
  PT_CHECK_ARG(a->rows_per_stat > 0 && a->num_stat > 0 && a->num_stat <= 1024, "pt_groupnorm: need 1..1024 statistics groups");
  PT_CHECK_ARG(!a->halo || (a->H > 0 && a->W > 0 && a->rows_per_stat % (a->H * a->W) == 0),
               "pt_groupnorm: halo output needs H*W dividing rows_per_stat");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = gn_block_threads(C);
  const int splits = gn_splits(a->num_stat, a->rows_per_stat, C);
  const int rpar = threads / (C / 8);

  GnStatsParams s;
  s.x0 = reinterpret_cast<const bf16*>(a->x0);
  s.x1 = reinterpret_cast<const bf16*>(a->x1);
  s.c0 = a->c0; s.c1 = a->c1; s.ld0 = a->ld0; s.ld1 = a->ld1;
  s.rows_per_stat = a->rows_per_stat;
  s.num_stat = a->num_stat;
  s.splits = splits;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->stats);
  s.tickets = reinterpret_cast<unsigned int*>(ws);
  s.mean_rstd = reinterpret_cast<float*>(ws + 4096);
  s.partials = reinterpret_cast<double*>(ws + 4096 + sizeof(float) * 64 * (size_t)a->num_stat);
  s.eps = a->eps;
  size_t stats_smem = sizeof(float) * ((size_t)2 * rpar * C + 2 * C);
  if (stats_smem < 4096) stats_smem = 4096;
  gn_stats_kernel<<<a->num_stat * splits, threads, stats_smem, st>>>(s);
  int rc = pt_launched("pt_groupnorm(stats)");
  if (rc) return rc;

  GnApplyParams p;
  p.x0 = s.x0; p.x1 = s.x1; p.c0 = a->c0; p.c1 = a->c1; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.rows_per_stat = a->rows_per_stat; p.num_stat = a->num_stat;
  p.mean_rstd = s.mean_rstd;
  p.gamma = a->gamma; p.beta = a->beta; p.silu = a->silu;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out_ld = a->out_ld;
  p.halo = a->halo; p.H = a->H > 0 ? a->H : 1; p.W = a->W > 0 ? a->W : 1;
  // the apply pass has its own split: enough CTAs to fill the machine, >= 16 rows per row lane
  const int out_rows = a->halo ? (a->rows_per_stat / (p.H * p.W)) * (p.H + 1) * (p.W + 1) : a->rows_per_stat;
  int asplits = (pt_num_sms() * 6 + a->num_stat - 1) / a->num_stat;
  const int amax = (out_rows + rpar * 16 - 1) / (rpar * 16);
  if (asplits > amax) asplits = amax;
  if (asplits < 1) asplits = 1;
  p.splits = asplits;
  gn_apply_kernel<<<a->num_stat * asplits, threads, 0, st>>>(p);
  return pt_launched("pt_groupnorm(apply)");
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
  const int rows_per_block = 8 * (32 / G);
  const int blocks = (a->rows + rows_per_block - 1) / rows_per_block;
  cudaStream_t st = (cudaStream_t)stream;
  if (G == 32) layernorm_kernel<32><<<blocks, 256, 0, st>>>(p);
  else if (G == 16) layernorm_kernel<16><<<blocks, 256, 0, st>>>(p);
  else layernorm_kernel<8><<<blocks, 256, 0, st>>>(p);
  return pt_launched("pt_layernorm");
}
