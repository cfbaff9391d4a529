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

// 8 channels (16 bytes) per thread
__global__ void __launch_bounds__(256) geglu_fwd8_kernel(const bf16* h, int ld, bf16* out, int out_ld, long long rows, int H) {
  const int hv = H >> 3;
  const long long total = rows * hv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / hv;
    const int c = (int)(i - r * hv) * 8;
    const uint4 uv = ldg_u4(h + (size_t)r * ld + c), ug = ldg_u4(h + (size_t)r * ld + H + c);
    const uint32_t* pv = reinterpret_cast<const uint32_t*>(&uv);
    const uint32_t* pg = reinterpret_cast<const uint32_t*>(&ug);
    uint4 o;
    uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v = unpack_bf16x2(pv[k]), g = unpack_bf16x2(pg[k]);
      po[k] = pack_bf16x2(geglu_gate_fast(v.x, g.x), geglu_gate_fast(v.y, g.y));
    }
    stg_u4(out + (size_t)r * out_ld + c, o);
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

// out[i] (+)= scale * sum_b partials[b][i], b in order.  Two shapes occur: (a) FEW partials of a LARGE array (wgrad: 2-20 row
// slices of a whole weight matrix) -> one thread per 4 consecutive elements, 16-byte loads, the b loop in registers;
// (b) MANY partials of a SMALL array (norm scale / shift gradients: hundreds of CTAs x 2C) -> 32 elements x 8 slices of b per
// CTA, slices folded in order.  Both are deterministic.
__global__ void __launch_bounds__(256) reduce_partials_vec4_kernel(const float* partials, int nb, long long n4, float scale, float* out,
                                                                   int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b2 = 0; b2 < nb; ++b2) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partials + (size_t)b2 * n4 * 4) + i);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out) + i;
    if (accumulate) {
      const float4 c = *o;
      *o = make_float4(c.x + scale * t.x, c.y + scale * t.y, c.z + scale * t.z, c.w + scale * t.w);
    } else {
      *o = make_float4(scale * t.x, scale * t.y, scale * t.z, scale * t.w);
    }
  }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* partials, int nb, long long n, float scale, float* out, int accumulate) {
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

// sum over [rows, C] of a*b (bf16): the AlphaBlender mix_factor gradient is such a dot product
__global__ void __launch_bounds__(256) dot_kernel(const bf16* a, int lda, const bf16* b, int ldb, long long rows, int C, double* partials) {
  __shared__ double red[32];
  double acc = 0.0;
  if ((C & 7) == 0 && (lda & 7) == 0 && (ldb & 7) == 0) {   // 8 elements (16 bytes) per thread and operand
    const int cv = C >> 3;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / cv;
      const int c = (int)(i - r * cv) * 8;
      const uint4 ua = ldg_u4(a + (size_t)r * lda + c), ub = ldg_u4(b + (size_t)r * ldb + c);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&ua);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(&ub);
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = unpack_bf16x2(pa[k]), y = unpack_bf16x2(pb[k]);
        t = fmaf(x.x, y.x, t);
        t = fmaf(x.y, y.y, t);
      }
      acc += (double)t;
    }
  } else {
    const long long total = rows * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / C;
      const int c = (int)(i - r * C);
      acc += (double)(__bfloat162float(a[(size_t)r * lda + c]) * __bfloat162float(b[(size_t)r * ldb + c]));
    }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) dot_final_kernel(const double* partials, int nb, float scale, float* out, int accumulate) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += partials[i];   // fixed thread -> element map: deterministic
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) {
    const float v = (float)(t * scale);
    out[0] = accumulate ? out[0] + v : v;
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

extern "C" int pt_geglu_fwd(const void* h, int32_t ld, void* out, int32_t out_ld, int64_t rows, int32_t hidden, void* stream) {
  PT_CHECK_ARG(h && out && rows > 0 && hidden > 0 && hidden % 2 == 0 && ld % 2 == 0 && out_ld % 2 == 0, "pt_geglu_fwd: bad argument");
  if (hidden % 8 == 0 && ld % 8 == 0 && out_ld % 8 == 0)
    pt_launch(geglu_fwd8_kernel, dim3(grid_for(rows * (hidden / 8), 256, 148 * 16)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h),
              (int)ld, reinterpret_cast<bf16*>(out), (int)out_ld, (long long)rows, (int)hidden);
  else
    pt_launch(geglu_fwd_kernel, dim3(grid_for(rows * (hidden / 2))), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(h), (int)ld,
              reinterpret_cast<bf16*>(out), (int)out_ld, (long long)rows, (int)hidden);
  return pt_launched("pt_geglu_fwd");
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
  if (nb <= 32 && n % 4 == 0 && ((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0)
    pt_launch(reduce_partials_vec4_kernel, dim3(grid_for(n / 4, 256, 148 * 16)), dim3(256), 0, stream, 1, partials, (int)nb, (long long)(n / 4), scale,
              out, (int)accumulate);
  else
    pt_launch(reduce_partials_kernel, dim3(grid_for(n, 32, 148 * 8)), dim3(256), 0, stream, 1, partials, (int)nb, (long long)n, scale, out,
              (int)accumulate);
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
  pt_launch(dot_final_kernel, dim3(1), dim3(256), 0, stream, 1, (const double*)workspace, blocks, scale, out, (int)accumulate);
  return pt_launched("pt_dot_bf16(final)");
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
