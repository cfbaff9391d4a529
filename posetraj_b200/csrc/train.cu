// posetraj_b200 — backward / optimizer kernels of the configs[3] training step.
//
// The reference trains the ControlNet by autograd through ControlNet + frozen UNet
// (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1404-1475: EDM-weighted MSE :1423-1436, accelerator.backward :1470,
// AdamW :1472).  These are the hand-written counterparts of the operators autograd differentiates there, in the data
// layouts of the forward library (token-major bf16, zero-haloed conv inputs) and following the formulas pinned against
// autograd in oracle/backward.py (tests/test_backward_oracle_cpu.py):
//   pt_edm_loss          loss :1423-1436 and d loss / d model_pred
//   pt_groupnorm_bwd     GroupNorm(32)(+SiLU) backward: 4-D and 5-D statistics, un-materialised channel concat, haloed dOut
//   pt_layernorm_bwd     LayerNorm backward
//   pt_geglu_fwd / _bwd  GEGLU as a separate elementwise pass (training keeps the pre-activations)
//   pt_colsum            bias / per-batch-row (time embedding) gradients: deterministic column sums
//   pt_reduce_partials   fixed-order fold of per-block partial gradients
//   pt_transpose_bf16    operand transposes for dgrad (W^T through pt_gemm with negated taps) and wgrad
//   pt_adamw             fused AdamW step on fp32 master weights + bf16 working copy
// dgrad is pt_gemm itself (oracle/backward.py: the forward kernel with negated shifts and W_t as [K, N]); wgrad is
// pt_wgrad in wgrad.cu (tcgen05).  All reductions are two-stage with a fixed order: two runs give identical bits.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

// ---------------------------------------------------------------------------------------------------------------
// block reduction helpers (fixed order: deterministic)
// ---------------------------------------------------------------------------------------------------------------
PT_DEVICE double block_sum_d(double v, double* red) {  // red: >= 32 doubles of shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ---------------------------------------------------------------------------------------------------------------
// EDM loss (train...cam_concat.py:1417-1436)
// ---------------------------------------------------------------------------------------------------------------
struct EdmParams {
  const bf16* pred;
  int pred_ld;
  const float* noisy;
  const float* target;
  long long sample_stride, frame_stride;
  const float* sigmas;
  int B, F, C, HW;
  float weight;
  bf16* dpred;
  int dpred_ld;
  double* partials;
  float* loss;
  int accumulate;
};

__global__ void __launch_bounds__(256) edm_loss_kernel(const EdmParams p) {
  __shared__ double red[32];
  const long long per_sample = (long long)p.F * p.HW;
  const long long total = (long long)p.B * per_sample;
  double acc = 0.0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per_sample);
    const long long rem = idx - (long long)b * per_sample;
    const int f = (int)(rem / p.HW);
    const int pix = (int)(rem - (long long)f * p.HW);
    const float s = p.sigmas[b];
    const float s2p1 = s * s + 1.0f;
    const float c_out = -s / sqrtf(s2p1);
    const float c_skip = 1.0f / s2p1;
    const float w = s2p1 / (s * s);
    // d mean(w (den - tgt)^2) / d pred = 2 w (den - tgt) c_out / (F C HW) / B
    const float gscale = p.weight * 2.0f * w * c_out / ((float)per_sample * p.C * p.B);
    for (int c = 0; c < p.C; ++c) {
      const long long li = (long long)b * p.sample_stride + (long long)f * p.frame_stride + (long long)c * p.HW + pix;
      const float pr = __bfloat162float(p.pred[(size_t)idx * p.pred_ld + c]);
      const float den = pr * c_out + c_skip * p.noisy[li];
      const float d = den - p.target[li];
      acc += (double)(w * d * d);
      if (p.dpred != nullptr) p.dpred[(size_t)idx * p.dpred_ld + c] = __float2bfloat16(gscale * d);
    }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) p.partials[blockIdx.x] = t;
}

__global__ void edm_loss_final_kernel(const EdmParams p, int nblocks) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nblocks; ++i) t += p.partials[i];
  // mean over (F C HW) per sample, then mean over the batch
  const double mean = t / ((double)p.F * p.HW * p.C * p.B);
  const float v = (float)(mean * p.weight);
  p.loss[0] = p.accumulate ? p.loss[0] + v : v;
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm(32) (+SiLU) backward
// ---------------------------------------------------------------------------------------------------------------
struct GnBwdParams {
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
  float* stats;   // [num_stat*32][4]: mean, rstd, S1/n, S2/n
  float* dgb;     // [num_stat][2][C] per-statistics-group partials of dgamma / dbeta (nullptr: not wanted)
};

PT_DEVICE float gn_load_x(const GnBwdParams& p, long long row, int c) {
  return c < p.c0 ? __bfloat162float(p.x0[(size_t)row * p.ld0 + c]) : __bfloat162float(p.x1[(size_t)row * p.ld1 + (c - p.c0)]);
}

PT_DEVICE long long gn_dout_row(const GnBwdParams& p, long long row) {
  if (!p.halo) return row;
  const int hw = p.H * p.W;
  const long long img = row / hw;
  const int rem = (int)(row - img * hw);
  const int y = rem / p.W, x = rem - y * p.W;
  return (img * (p.H + 1) + y) * (p.W + 1) + x;
}

// one block per (statistics group, norm group): phase 0 statistics, phase 1 the two backward sums (+ dgamma/dbeta
// partials of this statistics group), phase 2 dx
__global__ void __launch_bounds__(256) gn_bwd_kernel(const GnBwdParams p, int phase) {
  __shared__ double red[32];
  const int C = p.c0 + p.c1;
  const int cg = C / 32;
  const int s = blockIdx.x / 32, grp = blockIdx.x % 32;
  const long long n = (long long)p.rows_per_stat * cg;
  float* st = p.stats + (size_t)blockIdx.x * 4;
  if (phase == 0) {
    double a = 0.0, b = 0.0;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
      const long long r = e / cg;
      const int c = grp * cg + (int)(e - r * cg);
      const double v = gn_load_x(p, (long long)s * p.rows_per_stat + r, c);
      a += v;
      b += v * v;
    }
    const double sa = block_sum_d(a, red);
    const double sb = block_sum_d(b, red);
    if (threadIdx.x == 0) {
      const double mean = sa / (double)n;
      double var = sb / (double)n - mean * mean;
      if (var < 0.0) var = 0.0;
      st[0] = (float)mean;
      st[1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    return;
  }
  const float mean = st[0], rstd = st[1];
  if (phase == 1) {
    double a = 0.0, b = 0.0;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
      const long long r = e / cg;
      const int c = grp * cg + (int)(e - r * cg);
      const long long row = (long long)s * p.rows_per_stat + r;
      const float xh = (gn_load_x(p, row, c) - mean) * rstd;
      float dy = __bfloat162float(p.dout[(size_t)gn_dout_row(p, row) * p.dout_ld + c]);
      if (p.silu) {
        const float y = xh * p.gamma[c] + p.beta[c];
        const float sg = 1.0f / (1.0f + __expf(-y));
        dy *= sg * (1.0f + y * (1.0f - sg));
      }
      const float g = dy * p.gamma[c];
      a += (double)g;
      b += (double)g * xh;
    }
    const double sa = block_sum_d(a, red);
    const double sb = block_sum_d(b, red);
    if (threadIdx.x == 0) {
      st[2] = (float)(sa / (double)n);
      st[3] = (float)(sb / (double)n);
    }
    if (p.dgb != nullptr) {
      // per-channel sums over the rows of this statistics group: thread t -> channel (t % cg), row lanes t / cg
      const int lanes = blockDim.x / cg;
      if (lanes > 0) {
        __shared__ float cs[2][256];
        const int cl = threadIdx.x % cg, rl = threadIdx.x / cg;
        float ga = 0.f, be = 0.f;
        if (rl < lanes) {
          const int c = grp * cg + cl;
          for (long long r = rl; r < p.rows_per_stat; r += lanes) {
            const long long row = (long long)s * p.rows_per_stat + r;
            const float xh = (gn_load_x(p, row, c) - mean) * rstd;
            float dy = __bfloat162float(p.dout[(size_t)gn_dout_row(p, row) * p.dout_ld + c]);
            if (p.silu) {
              const float y = xh * p.gamma[c] + p.beta[c];
              const float sg = 1.0f / (1.0f + __expf(-y));
              dy *= sg * (1.0f + y * (1.0f - sg));
            }
            ga += dy * xh;
            be += dy;
          }
        }
        __syncthreads();
        cs[0][threadIdx.x] = ga;
        cs[1][threadIdx.x] = be;
        __syncthreads();
        if (threadIdx.x < cg) {
          float sg = 0.f, sb2 = 0.f;
          for (int l = 0; l < lanes; ++l) {
            sg += cs[0][l * cg + threadIdx.x];
            sb2 += cs[1][l * cg + threadIdx.x];
          }
          p.dgb[((size_t)s * 2 + 0) * C + grp * cg + threadIdx.x] = sg;
          p.dgb[((size_t)s * 2 + 1) * C + grp * cg + threadIdx.x] = sb2;
        }
      }
    }
    return;
  }
  // phase 2: dx = rstd (g - mean(g) - xh mean(g xh))
  const float m1 = st[2], m2 = st[3];
  for (long long e = threadIdx.x; e < n; e += blockDim.x) {
    const long long r = e / cg;
    const int c = grp * cg + (int)(e - r * cg);
    const long long row = (long long)s * p.rows_per_stat + r;
    const float xh = (gn_load_x(p, row, c) - mean) * rstd;
    float dy = __bfloat162float(p.dout[(size_t)gn_dout_row(p, row) * p.dout_ld + c]);
    if (p.silu) {
      const float y = xh * p.gamma[c] + p.beta[c];
      const float sg = 1.0f / (1.0f + __expf(-y));
      dy *= sg * (1.0f + y * (1.0f - sg));
    }
    const float g = dy * p.gamma[c];
    const float dx = rstd * (g - m1 - xh * m2);
    if (c < p.c0) p.dx0[(size_t)row * p.dld0 + c] = __float2bfloat16(dx);
    else p.dx1[(size_t)row * p.dld1 + (c - p.c0)] = __float2bfloat16(dx);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row; per-block partial dgamma / dbeta
// ---------------------------------------------------------------------------------------------------------------
struct LnBwdParams {
  const bf16* x;
  int ld;
  const bf16* dout;
  int dout_ld;
  const float* gamma;
  float eps;
  int rows, C;
  const float* addvec;  // optional: the forward normalised x + addvec[frame] (modified_svd.py:196-197)
  int hw, F;
  bf16* dx;
  int dx_ld;
  int accumulate_dx;    // dx += (the tensor also feeds a residual branch whose gradient is already there)
  float* partials;      // [gridDim.x][2][C] or nullptr
};

constexpr int kLnBwdMaxPerLane = 64;  // C <= 2048

__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p) {
  extern __shared__ float ln_sm[];   // [8 warps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = (p.C + 31) / 32;
  float dga[kLnBwdMaxPerLane], dbe[kLnBwdMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnBwdMaxPerLane; ++i) { dga[i] = 0.f; dbe[i] = 0.f; }
  for (int row = blockIdx.x * nw + warp; row < p.rows; row += gridDim.x * nw) {
    const bf16* xr = p.x + (size_t)row * p.ld;
    const bf16* dr = p.dout + (size_t)row * p.dout_ld;
    const float* av = p.addvec != nullptr ? p.addvec + (size_t)((row / p.hw) % p.F) * p.C : nullptr;
    float sum = 0.f, sq = 0.f;
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      if (c < p.C) {
        const float v = __bfloat162float(xr[c]) + (av ? av[c] : 0.f);
        sum += v;
        sq += v * v;
      }
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    const float mean = sum / p.C;
    const float rstd = rsqrtf(fmaxf(sq / p.C - mean * mean, 0.f) + p.eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      if (c < p.C) {
        const float xh = (__bfloat162float(xr[c]) + (av ? av[c] : 0.f) - mean) * rstd;
        const float g = __bfloat162float(dr[c]) * p.gamma[c];
        s1 += g;
        s2 += g * xh;
      }
    }
    s1 = warp_sum(s1) / p.C;
    s2 = warp_sum(s2) / p.C;
#pragma unroll
    for (int i = 0; i < kLnBwdMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < p.C) {
        const float xh = (__bfloat162float(xr[c]) + (av ? av[c] : 0.f) - mean) * rstd;
        const float d = __bfloat162float(dr[c]);
        float dx = rstd * (d * p.gamma[c] - s1 - xh * s2);
        bf16* o = p.dx + (size_t)row * p.dx_ld + c;
        if (p.accumulate_dx) dx += __bfloat162float(*o);
        *o = __float2bfloat16(dx);
        dga[i] += d * xh;
        dbe[i] += d;
      }
    }
  }
  if (p.partials == nullptr) return;
#pragma unroll
  for (int i = 0; i < kLnBwdMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (i < per && c < p.C) {
      ln_sm[(warp * 2 + 0) * p.C + c] = dga[i];
      ln_sm[(warp * 2 + 1) * p.C + c] = dbe[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * p.C; c += blockDim.x) {
    const int which = c / p.C, cc = c - which * p.C;
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += ln_sm[(w * 2 + which) * p.C + cc];
    p.partials[((size_t)blockIdx.x * 2 + which) * p.C + cc] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEGLU as a separate pass (training): out = v * gelu(g); backward with the exact erf derivative
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const bf16* h, int ld, bf16* out, int out_ld, long long rows, int H) {
  const long long total = rows * (H / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (H / 2);
    const int c = (int)(i - r * (H / 2)) * 2;
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h + (size_t)r * ld + c));
    const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h + (size_t)r * ld + H + c));
    *reinterpret_cast<uint32_t*>(out + (size_t)r * out_ld + c) = pack_bf16x2(geglu_gate_fast(v.x, g.x), geglu_gate_fast(v.y, g.y));
  }
}

__global__ void __launch_bounds__(256) geglu_bwd_kernel(const bf16* h, int ld, const bf16* dout, int dout_ld, bf16* dh, int dh_ld,
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
// column sums per row group (bias gradients: 1 group; time-embedding row-vector gradients: 1 group per batch row)
// ---------------------------------------------------------------------------------------------------------------
struct ColsumParams {
  const bf16* x;
  int ld, halo, H, W;
  long long rows_per_group;
  int groups, C;
  float scale;
  float* out;   // [groups][C]
  int accumulate;
};

// block = 32 columns x 8 row lanes; grid = (C/32 rounded up, groups)
__global__ void __launch_bounds__(256) colsum_kernel(const ColsumParams p) {
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int g = blockIdx.y;
  float acc = 0.f;
  if (c < p.C) {
    for (long long r = rl; r < p.rows_per_group; r += 8) {
      long long row = (long long)g * p.rows_per_group + r;
      if (p.halo) {
        const int hw = p.H * p.W;
        const long long img = row / hw;
        const int rem = (int)(row - img * hw);
        const int y = rem / p.W, x = rem - y * p.W;
        row = (img * (p.H + 1) + y) * (p.W + 1) + x;
      }
      acc += __bfloat162float(p.x[(size_t)row * p.ld + c]);
    }
  }
  sm[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && c < p.C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i][cl];
    float* o = p.out + (size_t)g * p.C + c;
    *o = p.accumulate ? *o + p.scale * t : p.scale * t;
  }
}

// out[i] (+)= scale * sum_b partials[b][i], b in order
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* partials, int nb, long long n, float scale, float* out, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int b = 0; b < nb; ++b) t += partials[(size_t)b * n + i];
    out[i] = accumulate ? out[i] + scale * t : scale * t;
  }
}

// sum over [rows, C] of a*b (bf16): the AlphaBlender mix_factor gradient is such a dot product
__global__ void __launch_bounds__(256) dot_kernel(const bf16* a, int lda, const bf16* b, int ldb, long long rows, int C, double* partials) {
  __shared__ double red[32];
  double acc = 0.0;
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    acc += (double)(__bfloat162float(a[(size_t)r * lda + c]) * __bfloat162float(b[(size_t)r * ldb + c]));
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void dot_final_kernel(const double* partials, int nb, float scale, float* out, int accumulate) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nb; ++i) t += partials[i];
  const float v = (float)(t * scale);
  out[0] = accumulate ? out[0] + v : v;
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 transpose [rows, cols] -> [cols, rows] (32 x 32 tiles through shared memory)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const bf16* in, int ld_in, bf16* out, int ld_out, int rows, int cols) {
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

// ---------------------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments)
// ---------------------------------------------------------------------------------------------------------------
struct AdamParams {
  float* master;
  const float* grad;
  float* m;
  float* v;
  bf16* work;   // optional bf16 working copy of the parameter
  long long n;
  float lr, beta1, beta2, eps, wd, bc1, bc2, grad_scale;
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams p) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const float g = p.grad[i] * p.grad_scale;
    float w = p.master[i];
    w -= p.lr * p.wd * w;
    const float m = p.beta1 * p.m[i] + (1.0f - p.beta1) * g;
    const float v = p.beta2 * p.v[i] + (1.0f - p.beta2) * g * g;
    p.m[i] = m;
    p.v[i] = v;
    w -= p.lr * (m / p.bc1) / (sqrtf(v / p.bc2) + p.eps);
    p.master[i] = w;
    if (p.work != nullptr) p.work[i] = __float2bfloat16(w);
  }
}

}  // namespace pt

using namespace pt;

static int grid_for(long long n, int per_block = 256, int max_blocks = 148 * 8) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

extern "C" int64_t pt_edm_loss_workspace_bytes(void) { return 8 * 1024; }

extern "C" int pt_edm_loss(const PtEdmLossArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->pred && a->noisy && a->target && a->sigmas && a->loss && a->workspace, "pt_edm_loss: null argument");
  PT_CHECK_ARG(a->B > 0 && a->F > 0 && a->C > 0 && a->HW > 0, "pt_edm_loss: empty problem");
  EdmParams p;
  p.pred = reinterpret_cast<const bf16*>(a->pred);
  p.pred_ld = a->pred_ld;
  p.noisy = a->noisy;
  p.target = a->target;
  p.sample_stride = a->sample_stride;
  p.frame_stride = a->frame_stride;
  p.sigmas = a->sigmas;
  p.B = a->B; p.F = a->F; p.C = a->C; p.HW = a->HW;
  p.weight = a->weight;
  p.dpred = reinterpret_cast<bf16*>(a->dpred);
  p.dpred_ld = a->dpred_ld;
  p.partials = reinterpret_cast<double*>(a->workspace);
  p.loss = a->loss;
  p.accumulate = a->accumulate;
  const int blocks = grid_for((long long)a->B * a->F * a->HW, 256, 1024);
  pt_launch(edm_loss_kernel, dim3(blocks), dim3(256), 0, stream, 1, p);
  int rc = pt_launched("pt_edm_loss");
  if (rc) return rc;
  pt_launch(edm_loss_final_kernel, dim3(1), dim3(32), 0, stream, 1, p, blocks);
  return pt_launched("pt_edm_loss(final)");
}

extern "C" int64_t pt_groupnorm_bwd_workspace_bytes(int32_t num_stat, int32_t channels) {
  if (num_stat < 1 || channels < 32) return -1;
  return (int64_t)num_stat * 32 * 4 * 4 + (int64_t)num_stat * 2 * channels * 4;
}

extern "C" int pt_groupnorm_bwd(const PtGroupNormBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x0 && a->dout && a->gamma && a->beta && a->dx0 && a->workspace, "pt_groupnorm_bwd: null argument");
  const int C = a->c0 + a->c1;
  PT_CHECK_ARG(a->c0 > 0 && C % 32 == 0 && C / 32 <= 256, "pt_groupnorm_bwd: channels must be a multiple of 32 (<= 8192)");
  PT_CHECK_ARG(a->c1 == 0 || (a->x1 != nullptr && a->dx1 != nullptr), "pt_groupnorm_bwd: c1 > 0 without x1 / dx1");
  PT_CHECK_ARG(a->rows_per_stat > 0 && a->num_stat > 0, "pt_groupnorm_bwd: empty problem");
  PT_CHECK_ARG(!a->halo || (a->H > 0 && a->W > 0 && a->rows_per_stat % (a->H * a->W) == 0), "pt_groupnorm_bwd: bad halo geometry");
  GnBwdParams p;
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
  p.stats = reinterpret_cast<float*>(a->workspace);
  p.dgb = (a->dgb_out != nullptr) ? p.stats + (size_t)a->num_stat * 32 * 4 : nullptr;
  for (int phase = 0; phase < 3; ++phase) {
    pt_launch(gn_bwd_kernel, dim3(a->num_stat * 32), dim3(256), 0, stream, 1, p, phase);
    int rc = pt_launched("pt_groupnorm_bwd");
    if (rc) return rc;
  }
  if (a->dgb_out != nullptr) {
    // fold the per-statistics-group partials [num_stat][2*C] (dgamma | dbeta) in order
    pt_launch(reduce_partials_kernel, dim3(grid_for(2LL * C)), dim3(256), 0, stream, 1, (const float*)p.dgb, a->num_stat, (long long)(2 * C),
              1.0f, a->dgb_out, a->accumulate_dgb);
    return pt_launched("pt_groupnorm_bwd(dgamma)");
  }
  return 0;
}

extern "C" int pt_layernorm_bwd(const PtLayerNormBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->dout && a->gamma && a->dx, "pt_layernorm_bwd: null argument");
  PT_CHECK_ARG(a->C >= 32 && a->C <= 32 * kLnBwdMaxPerLane && a->rows > 0, "pt_layernorm_bwd: C must be in [32, 2048]");
  LnBwdParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.dout_ld = a->dout_ld;
  p.gamma = a->gamma; p.eps = a->eps; p.rows = a->rows; p.C = a->C;
  p.addvec = a->addvec; p.hw = a->hw > 0 ? a->hw : 1; p.F = a->F > 0 ? a->F : 1;
  p.dx = reinterpret_cast<bf16*>(a->dx); p.dx_ld = a->dx_ld;
  p.accumulate_dx = a->accumulate_dx;
  p.partials = a->partials;
  int blocks = a->n_blocks;
  PT_CHECK_ARG(blocks >= 1 && blocks <= 4096, "pt_layernorm_bwd: n_blocks out of range");
  const size_t smem = (size_t)8 * 2 * a->C * sizeof(float);
  static bool attr_set[PT_MAX_DEVICES] = {false};
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4);
    if (e != cudaSuccess) return pt_fail(e, "pt_layernorm_bwd: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  pt_launch(ln_bwd_kernel, dim3(blocks), dim3(256), smem, stream, 1, p);
  int rc = pt_launched("pt_layernorm_bwd");
  if (rc) return rc;
  if (a->partials != nullptr && a->dgb_out != nullptr) {
    pt_launch(reduce_partials_kernel, dim3(grid_for(2LL * a->C)), dim3(256), 0, stream, 1, (const float*)a->partials, blocks,
              (long long)(2 * a->C), 1.0f, a->dgb_out, a->accumulate_dgb);
    return pt_launched("pt_layernorm_bwd(dgamma)");
  }
  return 0;
}

extern "C" int pt_geglu_fwd(const void* h, int32_t ld, void* out, int32_t out_ld, int64_t rows, int32_t hidden, void* stream) {
  PT_CHECK_ARG(h && out && rows > 0 && hidden > 0 && hidden % 2 == 0 && ld % 2 == 0 && out_ld % 2 == 0, "pt_geglu_fwd: bad argument");
  pt_launch(geglu_fwd_kernel, dim3(grid_for(rows * (hidden / 2))), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h), (int)ld,
            reinterpret_cast<bf16*>(out), (int)out_ld, (long long)rows, (int)hidden);
  return pt_launched("pt_geglu_fwd");
}

extern "C" int pt_geglu_bwd(const void* h, int32_t ld, const void* dout, int32_t dout_ld, void* dh, int32_t dh_ld, int64_t rows,
                            int32_t hidden, void* stream) {
  PT_CHECK_ARG(h && dout && dh && rows > 0 && hidden > 0, "pt_geglu_bwd: bad argument");
  pt_launch(geglu_bwd_kernel, dim3(grid_for(rows * hidden)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h), (int)ld,
            reinterpret_cast<const bf16*>(dout), (int)dout_ld, reinterpret_cast<bf16*>(dh), (int)dh_ld, (long long)rows, (int)hidden);
  return pt_launched("pt_geglu_bwd");
}

extern "C" int pt_colsum(const PtColsumArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->out && a->rows_per_group > 0 && a->groups > 0 && a->C > 0, "pt_colsum: bad argument");
  PT_CHECK_ARG(a->groups <= 65535, "pt_colsum: too many groups");
  ColsumParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld; p.halo = a->halo; p.H = a->H > 0 ? a->H : 1; p.W = a->W > 0 ? a->W : 1;
  p.rows_per_group = a->rows_per_group; p.groups = a->groups; p.C = a->C; p.scale = a->scale; p.out = a->out; p.accumulate = a->accumulate;
  pt_launch(colsum_kernel, dim3((a->C + 31) / 32, a->groups), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_colsum");
}

extern "C" int pt_reduce_partials(const float* partials, int32_t nb, int64_t n, float scale, float* out, int32_t accumulate, void* stream) {
  PT_CHECK_ARG(partials && out && nb > 0 && n > 0, "pt_reduce_partials: bad argument");
  pt_launch(reduce_partials_kernel, dim3(grid_for(n)), dim3(256), 0, stream, 1, partials, (int)nb, (long long)n, scale, out, (int)accumulate);
  return pt_launched("pt_reduce_partials");
}

extern "C" int pt_dot_bf16(const void* a, int32_t lda, const void* b, int32_t ldb, int64_t rows, int32_t cols, float scale, float* out,
                           int32_t accumulate, void* workspace, void* stream) {
  PT_CHECK_ARG(a && b && out && workspace && rows > 0 && cols > 0, "pt_dot_bf16: bad argument");
  const int blocks = grid_for(rows * cols, 256, 1024);
  pt_launch(dot_kernel, dim3(blocks), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(a), (int)lda, reinterpret_cast<const bf16*>(b),
            (int)ldb, (long long)rows, (int)cols, reinterpret_cast<double*>(workspace));
  int rc = pt_launched("pt_dot_bf16");
  if (rc) return rc;
  pt_launch(dot_final_kernel, dim3(1), dim3(32), 0, stream, 1, (const double*)workspace, blocks, scale, out, (int)accumulate);
  return pt_launched("pt_dot_bf16(final)");
}

extern "C" int pt_transpose_bf16(const void* in, int32_t ld_in, void* out, int32_t ld_out, int32_t rows, int32_t cols, void* stream) {
  PT_CHECK_ARG(in && out && rows > 0 && cols > 0, "pt_transpose_bf16: bad argument");
  PT_CHECK_ARG((rows + 31) / 32 <= 65535, "pt_transpose_bf16: too many rows for one launch");
  pt_launch(transpose_kernel, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(in), (int)ld_in,
            reinterpret_cast<bf16*>(out), (int)ld_out, (int)rows, (int)cols);
  return pt_launched("pt_transpose_bf16");
}

extern "C" int pt_adamw(const PtAdamWArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->master && a->grad && a->m && a->v && a->n > 0 && a->step >= 1, "pt_adamw: bad argument");
  AdamParams p;
  p.master = a->master; p.grad = a->grad; p.m = a->m; p.v = a->v; p.work = reinterpret_cast<bf16*>(a->work); p.n = a->n;
  p.lr = a->lr; p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps; p.wd = a->weight_decay; p.grad_scale = a->grad_scale;
  p.bc1 = 1.0f - powf(a->beta1, (float)a->step);
  p.bc2 = 1.0f - powf(a->beta2, (float)a->step);
  pt_launch(adamw_kernel, dim3(grid_for(a->n)), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_adamw");
}
