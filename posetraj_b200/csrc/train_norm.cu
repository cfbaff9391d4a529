// posetraj_b200 — normalisation / activation backward and the operand transpose of the training step (BASELINE
// configs[3]), rewritten for bandwidth: every pass reads / writes whole rows, 16 bytes per thread,
// reductions in two stages with a fixed order (deterministic).  Formulas: oracle/backward.py (checked against autograd).
//
//   pt_groupnorm_bwd   GroupNorm(32)(+SiLU) backward (diffusers ResnetBlock2D / TemporalResnetBlock norms, 4-D and 5-D
//                      statistics, optional un-materialised channel concat, zero-haloed dout).  Five launches:
//                        P0  per-(statistics group, row chunk, channel) sum x, sum x^2          (reads x)
//                        F0  -> mean, rstd per (statistics group, norm group)
//                        P1  per-(.., channel) A = sum dy', B = sum dy' xh, dy' = dy silu'(y)     (reads x, dout)
//                        F1  -> m1 = mean(gamma dy'), m2 = mean(gamma dy' xh) per norm group; dgamma = sum B, dbeta = sum A
//                        P2  dx = rstd (gamma dy' - m1 - xh m2)                                   (reads x, dout, writes dx)
//                      6 tensor passes of traffic; the first version ran one CTA per (statistics group, norm group) with
//                      2-byte strided loads (0.74 ms per call on average, 121 ms of the 443 ms step: profiles/r2j).
//   pt_layernorm_bwd   one warp per row, the row held in registers (one pass over x / dout), per-CTA dgamma / dbeta partials
//   pt_geglu_bwd       8 channels per thread
//   pt_transpose_bf16  64 x 64 tiles, 16-byte loads and stores (wgrad's dD^T operand)
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

PT_DEVICE void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
PT_DEVICE uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

static int grid_cap(long long n, int per_block, int max_blocks) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// out[i] (+)= scale * sum_b partials[b*n + i]: 32 elements x 8 slices of b per CTA, slices folded in order (deterministic)
__global__ void __launch_bounds__(256) fold_partials_kernel(const float* partials, int nb, long long n, float scale, float* out, int accumulate) {
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (long long i0 = (long long)blockIdx.x * 32; i0 < n; i0 += (long long)gridDim.x * 32) {
    const long long i = i0 + cl;
    float t = 0.f;
    if (i < n)
      for (int b = sl; b < nb; b += 8) t += partials[(size_t)b * n + i];
    sm[sl][cl] = t;
    __syncthreads();
    if (sl == 0 && i < n) {
      float tt = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tt += sm[k][cl];
      out[i] = accumulate ? out[i] + scale * tt : scale * tt;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm backward
// ---------------------------------------------------------------------------------------------------------------
struct GnB {
  const bf16* x0;
  const bf16* x1;
  int c0, c1, ld0, ld1;
  const bf16* dout;
  int dout_ld, halo, H, W;
  const float* gamma;
  const float* beta;
  float eps;
  int silu, rows_per_stat, num_stat;
  bf16* dx0;
  bf16* dx1;
  int dld0, dld1;
  int chunks, rows_per_chunk, cvec, rpar;
  float* part0;    // [num_stat*chunks][2][C]: sum x | sum x^2
  float* part1;    // [num_stat*chunks][2][C]: B (dgamma partial) | A (dbeta partial)
  float* meanrstd; // [num_stat][32][2]
  float* m12;      // [num_stat][32][2]
};

PT_DEVICE long long gnb_dout_row(const GnB& p, long long row) {
  if (!p.halo) return row;
  const int hw = p.H * p.W;
  const long long img = row / hw;
  const int rem = (int)(row - img * hw);
  const int y = rem / p.W, x = rem - y * p.W;
  return (img * (p.H + 1) + y) * (p.W + 1) + x;
}

// reduce the per-thread 2 x 8 accumulators over the row lanes and write the per-channel partials of this CTA
PT_DEVICE void gnb_write_partials(const GnB& p, float (&a)[8], float (&b)[8], float* sm, float* dst, int cl, int rl) {
  const int C = p.c0 + p.c1;
  float* sa = sm;
  float* sb = sm + (size_t)p.rpar * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sa[(size_t)rl * C + cl * 8 + i] = a[i];
    sb[(size_t)rl * C + cl * 8 + i] = b[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ta = 0.f, tb = 0.f;
    for (int l = 0; l < p.rpar; ++l) {
      ta += sa[(size_t)l * C + c];
      tb += sb[(size_t)l * C + c];
    }
    dst[c] = ta;
    dst[C + c] = tb;
  }
}

template <int PHASE>
__global__ void __launch_bounds__(512) gn_bwd_pass_kernel(const GnB p) {
  extern __shared__ float gnb_sm[];  // [2][rpar][C]
  const int C = p.c0 + p.c1;
  const int cg = C / 32;
  const int stat = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
  const int cl = threadIdx.x % p.cvec, rl = threadIdx.x / p.cvec;
  const int c = cl * 8;
  const long long r_begin = (long long)stat * p.rows_per_stat + (long long)chunk * p.rows_per_chunk;
  long long r_end = r_begin + p.rows_per_chunk;
  const long long stat_end = (long long)(stat + 1) * p.rows_per_stat;
  if (r_end > stat_end) r_end = stat_end;
  const bool from0 = c < p.c0;
  const bf16* xs = from0 ? p.x0 + c : p.x1 + (c - p.c0);
  const int xld = from0 ? p.ld0 : p.ld1;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.f;
  if (PHASE == 0) {
    for (long long r = r_begin + rl; r < r_end; r += p.rpar) {
      float v[8];
      unpack8(ldg_u4(xs + (size_t)r * xld), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += v[i];
        b[i] = fmaf(v[i], v[i], b[i]);
      }
    }
    gnb_write_partials(p, a, b, gnb_sm, p.part0 + (size_t)blockIdx.x * 2 * C, cl, rl);
    return;
  }
  float ga[8], be[8], mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ga[i] = p.gamma[c + i];
    be[i] = p.beta[c + i];
    const int g = (c + i) / cg;
    mean[i] = p.meanrstd[((size_t)stat * 32 + g) * 2];
    rstd[i] = p.meanrstd[((size_t)stat * 32 + g) * 2 + 1];
    if (PHASE == 2) {
      m1[i] = p.m12[((size_t)stat * 32 + g) * 2];
      m2[i] = p.m12[((size_t)stat * 32 + g) * 2 + 1];
    }
  }
  bf16* dxs = from0 ? p.dx0 + c : p.dx1 + (c - p.c0);
  const int dld = from0 ? p.dld0 : p.dld1;
  for (long long r = r_begin + rl; r < r_end; r += p.rpar) {
    float v[8], d[8];
    unpack8(ldg_u4(xs + (size_t)r * xld), v);
    unpack8(ldg_u4(p.dout + (size_t)gnb_dout_row(p, r) * p.dout_ld + c), d);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (v[i] - mean[i]) * rstd[i];
      float dy = d[i];
      if (p.silu) {
        const float y = fmaf(xh, ga[i], be[i]);
        const float sg = 1.0f / (1.0f + __expf(-y));
        dy *= sg * (1.0f + y * (1.0f - sg));
      }
      if (PHASE == 1) {
        a[i] = fmaf(dy, xh, a[i]);   // B: dgamma partial
        b[i] += dy;                  // A: dbeta partial
      } else {
        o[i] = rstd[i] * (dy * ga[i] - m1[i] - xh * m2[i]);
      }
    }
    if (PHASE == 2) stg_u4(dxs + (size_t)r * dld, pack8(o));
  }
  if (PHASE == 1) gnb_write_partials(p, a, b, gnb_sm, p.part1 + (size_t)blockIdx.x * 2 * C, cl, rl);
}

// one CTA per (statistics group, norm group): fold the chunk x channel partials of the group in fp64 (fixed order)
template <int PHASE>
__global__ void __launch_bounds__(256) gn_bwd_fold_kernel(const GnB p) {
  __shared__ double red0[8], red1[8];
  const int C = p.c0 + p.c1;
  const int cg = C / 32;
  const int stat = blockIdx.x >> 5, g = blockIdx.x & 31;
  const float* part = (PHASE == 0 ? p.part0 : p.part1) + (size_t)stat * p.chunks * 2 * C + g * cg;
  double s0 = 0.0, s1 = 0.0;
  const int total = p.chunks * cg;
  for (int e = threadIdx.x; e < total; e += 256) {
    const int k = e / cg, i = e - k * cg;
    double v0 = (double)part[(size_t)k * 2 * C + i];
    double v1 = (double)part[(size_t)k * 2 * C + C + i];
    if (PHASE == 1) {   // weight by gamma: sum_c gamma_c B_c, sum_c gamma_c A_c
      const double gm = (double)p.gamma[g * cg + i];
      v0 *= gm;
      v1 *= gm;
    }
    s0 += v0;
    s1 += v1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red0[threadIdx.x >> 5] = s0;
    red1[threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s0 = s1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      s0 += red0[w];
      s1 += red1[w];
    }
    const double n = (double)p.rows_per_stat * cg;
    if (PHASE == 0) {
      const double mean = s0 / n;
      double var = s1 / n - mean * mean;
      if (var < 0.0) var = 0.0;
      p.meanrstd[((size_t)stat * 32 + g) * 2] = (float)mean;
      p.meanrstd[((size_t)stat * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
    } else {
      p.m12[((size_t)stat * 32 + g) * 2] = (float)(s1 / n);      // mean(gamma dy')
      p.m12[((size_t)stat * 32 + g) * 2 + 1] = (float)(s0 / n);  // mean(gamma dy' xh)
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row, NO octets (8 channels) per lane held in registers
// ---------------------------------------------------------------------------------------------------------------
struct LnB {
  const bf16* x;
  int ld;
  const bf16* dout;
  int dout_ld;
  const float* gamma;
  float eps;
  int rows, C;
  const float* addvec;
  int hw, F;
  bf16* dx;
  int dx_ld, accumulate_dx;
  float* partials;  // [gridDim.x][2][C] or nullptr
};

template <int NO>
__global__ void __launch_bounds__(256) ln_bwd2_kernel(const LnB p) {
  extern __shared__ float lnb_sm[];   // [8 warps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int noct = p.C >> 3;
  float dga[NO][8], dbe[NO][8], gam[NO][8];
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    const int o = lane + 32 * i;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dga[i][k] = dbe[i][k] = 0.f;
      gam[i][k] = o < noct ? p.gamma[o * 8 + k] : 0.f;
    }
  }
  const float inv_c = 1.0f / (float)p.C;
  for (int row = blockIdx.x * nw + warp; row < p.rows; row += gridDim.x * nw) {
    const bf16* xr = p.x + (size_t)row * p.ld;
    const bf16* dr = p.dout + (size_t)row * p.dout_ld;
    const float* av = p.addvec != nullptr ? p.addvec + (size_t)((row / p.hw) % p.F) * p.C : nullptr;
    float v[NO][8], d[NO][8];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      const int o = lane + 32 * i;
      if (o < noct) {
        unpack8(ldg_u4(xr + o * 8), v[i]);
        unpack8(ldg_u4(dr + o * 8), d[i]);
        if (av != nullptr) {
          const float4 a0 = *reinterpret_cast<const float4*>(av + o * 8), a1 = *reinterpret_cast<const float4*>(av + o * 8 + 4);
          v[i][0] += a0.x; v[i][1] += a0.y; v[i][2] += a0.z; v[i][3] += a0.w;
          v[i][4] += a1.x; v[i][5] += a1.y; v[i][6] += a1.z; v[i][7] += a1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          sum += v[i][k];
          sq = fmaf(v[i][k], v[i][k], sq);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = d[i][k] = 0.f;
      }
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    const float mean = sum * inv_c;
    const float rstd = rsqrtf(fmaxf(sq * inv_c - mean * mean, 0.f) + p.eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      if (lane + 32 * i < noct) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[i][k] = (v[i][k] - mean) * rstd;    // xh
          const float g = d[i][k] * gam[i][k];
          s1 += g;
          s2 = fmaf(g, v[i][k], s2);
        }
      }
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      const int o = lane + 32 * i;
      if (o < noct) {
        float out[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          out[k] = rstd * (d[i][k] * gam[i][k] - s1 - v[i][k] * s2);
          dga[i][k] = fmaf(d[i][k], v[i][k], dga[i][k]);
          dbe[i][k] += d[i][k];
        }
        bf16* optr = p.dx + (size_t)row * p.dx_ld + o * 8;
        if (p.accumulate_dx) {
          float prev[8];
          unpack8(ldg_u4(optr), prev);
#pragma unroll
          for (int k = 0; k < 8; ++k) out[k] += prev[k];
        }
        stg_u4(optr, pack8(out));
      }
    }
  }
  if (p.partials == nullptr) return;
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    const int o = lane + 32 * i;
    if (o < noct) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lnb_sm[(size_t)(warp * 2 + 0) * p.C + o * 8 + k] = dga[i][k];
        lnb_sm[(size_t)(warp * 2 + 1) * p.C + o * 8 + k] = dbe[i][k];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * p.C; c += blockDim.x) {
    const int which = c / p.C, cc = c - which * p.C;
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += lnb_sm[(size_t)(w * 2 + which) * p.C + cc];
    p.partials[((size_t)blockIdx.x * 2 + which) * p.C + cc] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEGLU backward (exact erf derivative), 8 channels per thread
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) geglu_bwd2_kernel(const bf16* h, int ld, const bf16* dout, int dout_ld, bf16* dh, int dh_ld, long long rows,
                                                         int H) {
  const int hv = H >> 3;
  const long long total = rows * hv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / hv;
    const int c = (int)(i - r * hv) * 8;
    float v[8], g[8], d[8], ov[8], og[8];
    unpack8(ldg_u4(h + (size_t)r * ld + c), v);
    unpack8(ldg_u4(h + (size_t)r * ld + H + c), g);
    unpack8(ldg_u4(dout + (size_t)r * dout_ld + c), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float Phi = 0.5f * (1.0f + erff(g[k] * 0.70710678118654752440f));
      const float phi = 0.3989422804014327f * __expf(-0.5f * g[k] * g[k]);
      ov[k] = d[k] * g[k] * Phi;
      og[k] = d[k] * v[k] * (Phi + g[k] * phi);
    }
    stg_u4(dh + (size_t)r * dh_ld + c, pack8(ov));
    stg_u4(dh + (size_t)r * dh_ld + H + c, pack8(og));
  }
}

__global__ void __launch_bounds__(256) geglu_bwd_scalar_kernel(const bf16* h, int ld, const bf16* dout, int dout_ld, bf16* dh, int dh_ld,
                                                               long long rows, int H) {
  const long long total = rows * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int c = (int)(i - r * H);
    const float v = __bfloat162float(h[(size_t)r * ld + c]);
    const float g = __bfloat162float(h[(size_t)r * ld + H + c]);
    const float d = __bfloat162float(dout[(size_t)r * dout_ld + c]);
    const float Phi = 0.5f * (1.0f + erff(g * 0.70710678118654752440f));
    const float phi = 0.3989422804014327f * __expf(-0.5f * g * g);
    dh[(size_t)r * dh_ld + c] = __float2bfloat16(d * g * Phi);
    dh[(size_t)r * dh_ld + H + c] = __float2bfloat16(d * v * (Phi + g * phi));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 transpose [rows, cols] -> [cols, rows]
// ---------------------------------------------------------------------------------------------------------------
// 64 x 64 tiles, 16-byte loads and stores (cols, ld_in, ld_out multiples of 8)
__global__ void __launch_bounds__(256) transpose64_kernel(const bf16* in, int ld_in, bf16* out, int ld_out, int rows, int cols) {
  __shared__ __align__(16) unsigned short tile[64][72];   // row stride 144 B: 16-byte aligned rows
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, ch = i & 7;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < rows && c0 + ch * 8 < cols) u = ldg_u4(in + (size_t)(r0 + r) * ld_in + c0 + ch * 8);
    *reinterpret_cast<uint4*>(&tile[r][ch * 8]) = u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int c = i & 63, rch = i >> 6;       // output row c (input column), 8 input rows rch*8 .. +7
    if (c0 + c < cols && r0 + rch * 8 < rows) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k] = (uint32_t)tile[rch * 8 + 2 * k][c] | ((uint32_t)tile[rch * 8 + 2 * k + 1][c] << 16);
      stg_u4(out + (size_t)(c0 + c) * ld_out + r0 + rch * 8, make_uint4(w[0], w[1], w[2], w[3]));
    }
  }
}

__global__ void __launch_bounds__(256) transpose32_kernel(const bf16* in, int ld_in, bf16* out, int ld_out, int rows, int cols) {
  __shared__ bf16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) out[(size_t)c * ld_out + r] = tile[tx][i];
  }
}

}  // namespace pt

using namespace pt;

static int gnb_chunks(int num_stat) {
  int c = (592 + num_stat - 1) / num_stat;
  return c < 1 ? 1 : c;
}

extern "C" int64_t pt_groupnorm_bwd_workspace_bytes(int32_t num_stat, int32_t channels) {
  if (num_stat < 1 || channels < 32) return -1;
  const int64_t blocks = (int64_t)num_stat * gnb_chunks(num_stat);
  return 2 * blocks * 2 * channels * (int64_t)sizeof(float) + (int64_t)num_stat * 32 * 4 * (int64_t)sizeof(float);
}

extern "C" int pt_groupnorm_bwd(const PtGroupNormBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x0 && a->dout && a->gamma && a->beta && a->dx0 && a->workspace, "pt_groupnorm_bwd: null argument");
  const int C = a->c0 + a->c1;
  PT_CHECK_ARG(a->c0 > 0 && C % 32 == 0 && C % 8 == 0 && a->c0 % 8 == 0 && C <= 4096, "pt_groupnorm_bwd: channels must be multiples of 32 (concat boundary: 8), <= 4096");
  PT_CHECK_ARG(a->c1 == 0 || (a->x1 != nullptr && a->dx1 != nullptr), "pt_groupnorm_bwd: c1 > 0 without x1 / dx1");
  PT_CHECK_ARG(a->rows_per_stat > 0 && a->num_stat > 0, "pt_groupnorm_bwd: empty problem");
  PT_CHECK_ARG(!a->halo || (a->H > 0 && a->W > 0 && a->rows_per_stat % (a->H * a->W) == 0), "pt_groupnorm_bwd: bad halo geometry");
  PT_CHECK_ARG(a->ld0 % 8 == 0 && a->dout_ld % 8 == 0 && a->dld0 % 8 == 0 && (a->c1 == 0 || (a->ld1 % 8 == 0 && a->dld1 % 8 == 0)),
               "pt_groupnorm_bwd: row strides must be multiples of 8");
  GnB p;
  p.x0 = reinterpret_cast<const bf16*>(a->x0);
  p.x1 = reinterpret_cast<const bf16*>(a->x1);
  p.c0 = a->c0; p.c1 = a->c1; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.dout = reinterpret_cast<const bf16*>(a->dout);
  p.dout_ld = a->dout_ld; p.halo = a->halo; p.H = a->H > 0 ? a->H : 1; p.W = a->W > 0 ? a->W : 1;
  p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
  p.rows_per_stat = a->rows_per_stat; p.num_stat = a->num_stat;
  p.dx0 = reinterpret_cast<bf16*>(a->dx0);
  p.dx1 = reinterpret_cast<bf16*>(a->dx1);
  p.dld0 = a->dld0; p.dld1 = a->dld1;
  p.cvec = C / 8;
  p.rpar = 512 / p.cvec;
  if (p.rpar < 1) p.rpar = 1;
  int chunks = gnb_chunks(a->num_stat);
  const int max_chunks = (a->rows_per_stat + 4 * p.rpar - 1) / (4 * p.rpar);   // at least ~4 rows per row lane
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.chunks = chunks;
  p.rows_per_chunk = (a->rows_per_stat + chunks - 1) / chunks;
  const size_t blocks_ws = (size_t)a->num_stat * gnb_chunks(a->num_stat);
  p.part0 = reinterpret_cast<float*>(a->workspace);
  p.part1 = p.part0 + blocks_ws * 2 * C;
  p.meanrstd = p.part1 + blocks_ws * 2 * C;
  p.m12 = p.meanrstd + (size_t)a->num_stat * 64;
  const int threads = p.cvec * p.rpar;
  const size_t smem = (size_t)2 * p.rpar * C * sizeof(float);
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(gn_bwd_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_bwd_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return pt_fail(e, "pt_groupnorm_bwd: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  PT_CHECK_ARG(smem <= 64 * 1024, "pt_groupnorm_bwd: too many channels for the reduction scratch");
  const dim3 grid(a->num_stat * chunks);
  int rc;
  pt_launch(gn_bwd_pass_kernel<0>, grid, dim3(threads), smem, stream, 1, p);
  if ((rc = pt_launched("pt_groupnorm_bwd (P0)")) != 0) return rc;
  pt_launch(gn_bwd_fold_kernel<0>, dim3(a->num_stat * 32), dim3(256), 0, stream, 1, p);
  if ((rc = pt_launched("pt_groupnorm_bwd (F0)")) != 0) return rc;
  pt_launch(gn_bwd_pass_kernel<1>, grid, dim3(threads), smem, stream, 1, p);
  if ((rc = pt_launched("pt_groupnorm_bwd (P1)")) != 0) return rc;
  pt_launch(gn_bwd_fold_kernel<1>, dim3(a->num_stat * 32), dim3(256), 0, stream, 1, p);
  if ((rc = pt_launched("pt_groupnorm_bwd (F1)")) != 0) return rc;
  pt_launch(gn_bwd_pass_kernel<2>, grid, dim3(threads), 0, stream, 1, p);
  if ((rc = pt_launched("pt_groupnorm_bwd (P2)")) != 0) return rc;
  if (a->dgb_out != nullptr) {
    // dgamma | dbeta = the per-channel B | A partials summed over every (statistics group, chunk) in order
    pt_launch(fold_partials_kernel, dim3(grid_cap(2LL * C, 32, 512)), dim3(256), 0, stream, 1, (const float*)p.part1, a->num_stat * chunks,
              (long long)(2 * C), 1.0f, a->dgb_out, (int)a->accumulate_dgb);
    return pt_launched("pt_groupnorm_bwd (dgamma)");
  }
  return 0;
}

template <int NO>
static int launch_ln_bwd(const LnB& p, int blocks, size_t smem, void* stream) {
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(ln_bwd2_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4);
    if (e != cudaSuccess) return pt_fail(e, "pt_layernorm_bwd: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  pt_launch(ln_bwd2_kernel<NO>, dim3(blocks), dim3(256), smem, stream, 1, p);
  return pt_launched("pt_layernorm_bwd");
}

extern "C" int pt_layernorm_bwd(const PtLayerNormBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->dout && a->gamma && a->dx, "pt_layernorm_bwd: null argument");
  PT_CHECK_ARG(a->C >= 32 && a->C <= 2048 && a->C % 8 == 0 && a->rows > 0, "pt_layernorm_bwd: C must be a multiple of 8 in [32, 2048]");
  PT_CHECK_ARG(a->ld % 8 == 0 && a->dout_ld % 8 == 0 && a->dx_ld % 8 == 0, "pt_layernorm_bwd: row strides must be multiples of 8");
  LnB p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.dout_ld = a->dout_ld;
  p.gamma = a->gamma; p.eps = a->eps; p.rows = a->rows; p.C = a->C;
  p.addvec = a->addvec; p.hw = a->hw > 0 ? a->hw : 1; p.F = a->F > 0 ? a->F : 1;
  p.dx = reinterpret_cast<bf16*>(a->dx); p.dx_ld = a->dx_ld;
  p.accumulate_dx = a->accumulate_dx;
  p.partials = a->partials;
  const int blocks = a->n_blocks;
  PT_CHECK_ARG(blocks >= 1 && blocks <= 4096, "pt_layernorm_bwd: n_blocks out of range");
  const size_t smem = a->partials != nullptr ? (size_t)8 * 2 * a->C * sizeof(float) : 0;
  const int no = (a->C / 8 + 31) / 32;
  int rc;
  if (no <= 1) rc = launch_ln_bwd<1>(p, blocks, smem, stream);
  else if (no == 2) rc = launch_ln_bwd<2>(p, blocks, smem, stream);
  else if (no == 3) rc = launch_ln_bwd<3>(p, blocks, smem, stream);
  else if (no <= 5) rc = launch_ln_bwd<5>(p, blocks, smem, stream);
  else rc = launch_ln_bwd<8>(p, blocks, smem, stream);
  if (rc) return rc;
  if (a->partials != nullptr && a->dgb_out != nullptr) {
    pt_launch(fold_partials_kernel, dim3(grid_cap(2LL * a->C, 32, 512)), dim3(256), 0, stream, 1, (const float*)a->partials, blocks,
              (long long)(2 * a->C), 1.0f, a->dgb_out, (int)a->accumulate_dgb);
    return pt_launched("pt_layernorm_bwd (dgamma)");
  }
  return 0;
}

extern "C" int pt_geglu_bwd(const void* h, int32_t ld, const void* dout, int32_t dout_ld, void* dh, int32_t dh_ld, int64_t rows,
                            int32_t hidden, void* stream) {
  PT_CHECK_ARG(h && dout && dh && rows > 0 && hidden > 0, "pt_geglu_bwd: bad argument");
  if (hidden % 8 == 0 && ld % 8 == 0 && dout_ld % 8 == 0 && dh_ld % 8 == 0) {
    pt_launch(geglu_bwd2_kernel, dim3(grid_cap(rows * (hidden / 8), 256, 148 * 16)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h),
              (int)ld, reinterpret_cast<const bf16*>(dout), (int)dout_ld, reinterpret_cast<bf16*>(dh), (int)dh_ld, (long long)rows, (int)hidden);
  } else {
    pt_launch(geglu_bwd_scalar_kernel, dim3(grid_cap(rows * hidden, 256, 148 * 16)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h),
              (int)ld, reinterpret_cast<const bf16*>(dout), (int)dout_ld, reinterpret_cast<bf16*>(dh), (int)dh_ld, (long long)rows, (int)hidden);
  }
  return pt_launched("pt_geglu_bwd");
}

extern "C" int pt_transpose_bf16(const void* in, int32_t ld_in, void* out, int32_t ld_out, int32_t rows, int32_t cols, void* stream) {
  PT_CHECK_ARG(in && out && rows > 0 && cols > 0, "pt_transpose_bf16: bad argument");
  PT_CHECK_ARG((rows + 31) / 32 <= 65535, "pt_transpose_bf16: too many rows for one launch");
  // vector path: the 8-row groups of the last tile may spill into the padding of an output row (zeros are written there)
  const bool vec = cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ld_out >= (rows + 7) / 8 * 8 && ((uintptr_t)in % 16 == 0) &&
                   ((uintptr_t)out % 16 == 0);
  if (vec)
    pt_launch(transpose64_kernel, dim3((cols + 63) / 64, (rows + 63) / 64), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(in), (int)ld_in,
              reinterpret_cast<bf16*>(out), (int)ld_out, (int)rows, (int)cols);
  else
    pt_launch(transpose32_kernel, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(in), (int)ld_in,
              reinterpret_cast<bf16*>(out), (int)ld_out, (int)rows, (int)cols);
  return pt_launched("pt_transpose_bf16");
}
