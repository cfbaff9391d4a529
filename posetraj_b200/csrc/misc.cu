// posetraj_b200 — small latency-bound / layout kernels around the GEMM core.
//
//   pt_small_linear      tiny-M linear layers (time / aug / frame-position embeddings, the batched time_emb_proj of
//                        all resnets, the degenerate 1-token cross-attention vectors, the camera columns of
//                        cc_projection): weight-bandwidth bound GEMV, one warp per output feature.
//   pt_timestep_sincos   Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0), optionally
//                        reading t = 0.25*ln(sigma[step]) from the device sigma table so a captured step replays.
//   pt_upsample2x        nearest 2x (Upsample2D) written straight into the zero-haloed conv input layout.
//   pt_conv3x3_direct    3x3 conv (+SiLU) for the <= 32-channel head of the ControlNet conditioning embedding
//                        (models/controlnet_sdv.py:84-109): bandwidth-bound, kept off the tensor cores.
//   pt_nchw_to_tokens / pt_tokens_to_nchw   boundary layout conversion (reference tensors are NCHW).
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

// ---------------------------------------------------------------------------------------------------------
struct SmallLinearParams {
  const float* in;   // [M, K] fp32
  int in_ld;
  const bf16* w;     // [N, K] bf16 (row stride w_ld)
  int w_ld;
  const float* bias; // [N] or null
  float* out;        // [M, N] fp32 (row stride out_ld)
  int out_ld;
  int M, N, K;
  int act_in_silu, act_out_silu, accumulate;
};

// kStage: the (activated) input rows are staged ONCE per CTA in shared memory.  Without it every warp re-evaluated
// silu(x) for its own output row: 2 MUFU per element x N rows — for the batched time_emb_proj GEMV (N ~ 55 000, K = 1280)
// that was 176 us per launch, 3x the time of streaming its 141 MB of weights (profiles/r2v_train_launches.md).
template <bool kStage>
__global__ void __launch_bounds__(256) small_linear_kernel(const SmallLinearParams p) {
  extern __shared__ __align__(16) float sl_x[];   // kStage: [rows of this pass (<= 8)][K]
  griddep_launch();
  griddep_wait();
  // a warp walks output rows n_first, n_first + warps_total, ...: with the staged input a CTA amortises its prologue (the
  // M x K activated inputs) over many rows — one row per warp made the 40 320-row time_emb_proj GEMV of the UNet 5 040 CTAs
  // that each staged 10 KB to stream 20 KB of weights (72.6 us for 103 MB = 1.4 TB/s)
  const int n_first = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (!kStage && n_first >= p.N) return;
  for (int m0 = 0; m0 < p.M; m0 += 8) {
    const int mrows = min(8, p.M - m0);
    if (kStage) {
      if (m0 > 0) __syncthreads();
      for (int idx = threadIdx.x; idx < mrows * p.K; idx += 256) {
        const int i = idx / p.K, k = idx - i * p.K;
        float x = p.in[(size_t)(m0 + i) * p.in_ld + k];
        if (p.act_in_silu) x = silu_f(x);
        sl_x[idx] = x;
      }
      __syncthreads();
    }
    for (int n = n_first; n < p.N; n += warps_total) {
    const bf16* wr = p.w + (size_t)n * p.w_ld;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if ((p.K & 7) == 0 && (p.w_ld & 7) == 0) {
#pragma unroll 5
      for (int k = lane * 8; k < p.K; k += 256) {
        const uint4 u = ldg_u4(wr + k);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        const float wv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < mrows) {
            if (kStage) {
              const float4 x0 = *reinterpret_cast<const float4*>(sl_x + (size_t)i * p.K + k);
              const float4 x1 = *reinterpret_cast<const float4*>(sl_x + (size_t)i * p.K + k + 4);
              acc[i] = fmaf(x0.x, wv[0], acc[i]); acc[i] = fmaf(x0.y, wv[1], acc[i]);
              acc[i] = fmaf(x0.z, wv[2], acc[i]); acc[i] = fmaf(x0.w, wv[3], acc[i]);
              acc[i] = fmaf(x1.x, wv[4], acc[i]); acc[i] = fmaf(x1.y, wv[5], acc[i]);
              acc[i] = fmaf(x1.z, wv[6], acc[i]); acc[i] = fmaf(x1.w, wv[7], acc[i]);
            } else {
              const float* xr = p.in + (size_t)(m0 + i) * p.in_ld + k;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float x = xr[j];
                if (p.act_in_silu) x = silu_f(x);
                acc[i] = fmaf(x, wv[j], acc[i]);
              }
            }
          }
        }
      }
    } else {
      for (int k = lane; k < p.K; k += 32) {
        const float wv = __bfloat162float(wr[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < mrows) {
            float x;
            if (kStage) {
              x = sl_x[(size_t)i * p.K + k];
            } else {
              x = p.in[(size_t)(m0 + i) * p.in_ld + k];
              if (p.act_in_silu) x = silu_f(x);
            }
            acc[i] = fmaf(x, wv, acc[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < mrows) {
          float v = acc[i] + (p.bias != nullptr ? p.bias[n] : 0.f);
          if (p.act_out_silu) v = silu_f(v);
          float* o = p.out + (size_t)(m0 + i) * p.out_ld + n;
          *o = p.accumulate ? (*o + v) : v;
        }
      }
    }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct SinCosParams {
  const float* t;       // [M] or null
  const float* sigmas;  // used when t is null: t = 0.25 * ln(sigmas[*step_index]) for every row
  const int* step_index;
  float* out;           // [M, dim]
  int out_ld;
  int M, dim;
};

__global__ void sincos_kernel(const SinCosParams p) {
  griddep_launch();
  griddep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = p.dim >> 1;
  if (idx >= p.M * half) return;
  const int m = idx / half;
  const int k = idx - m * half;
  float t;
  if (p.t != nullptr) {
    t = p.t[m];
  } else {
    t = 0.25f * logf(p.sigmas[*p.step_index]);
  }
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
  const float arg = t * freq;
  p.out[(size_t)m * p.out_ld + k] = cosf(arg);         // flip_sin_to_cos: cos first
  p.out[(size_t)m * p.out_ld + half + k] = sinf(arg);
}

// ---------------------------------------------------------------------------------------------------------
struct UpsampleParams {
  const bf16* x;  // compact [n, H, W, C]
  int ld;
  bf16* out;      // haloed [n, 2H+1, 2W+1, C] (or compact [n, 2H, 2W, C])
  int out_ld;
  int n, H, W, C, halo, scale;
};

// One thread per INPUT vector (16 bytes): one load, scale x scale stores plus the zero halo it is responsible for (the
// column right of the last pixel of a row, the row below the last row of an image).  blockIdx.y walks the input image
// rows, so the only integer division is x = i / cvec.  (The first version ran one thread per OUTPUT vector with six 64-bit
// divisions each: ~300 instructions per 16 bytes, 0.40 of the HBM copy rate.)
__global__ void __launch_bounds__(256) upsample2x_kernel(const UpsampleParams p) {
  griddep_launch();
  griddep_wait();
  const int cvec = p.C >> 3;
  const int sc = p.scale;
  const int oW = sc * p.W + (p.halo ? 1 : 0), oH = sc * p.H + (p.halo ? 1 : 0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.W * cvec) return;
  const int x = i / cvec, cv = i - x * cvec;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int yrow = blockIdx.y; yrow < p.n * p.H; yrow += gridDim.y) {
    const int img = yrow / p.H, y = yrow - img * p.H;
    const uint4 v = ldg_u4(p.x + ((size_t)yrow * p.W + x) * p.ld + cv * 8);
    bf16* o = p.out + (((size_t)img * oH + (size_t)(sc * y)) * oW + (size_t)(sc * x)) * p.out_ld + cv * 8;
    for (int dy = 0; dy < sc; ++dy)
      for (int dx = 0; dx < sc; ++dx) stg_u4(o + ((size_t)dy * oW + dx) * p.out_ld, v);
    if (p.halo) {
      if (x == p.W - 1)
        for (int dy = 0; dy < sc; ++dy) stg_u4(o + ((size_t)dy * oW + sc) * p.out_ld, zero4);
      if (y == p.H - 1) {
        for (int dx = 0; dx < sc; ++dx) stg_u4(o + ((size_t)sc * oW + dx) * p.out_ld, zero4);
        if (x == p.W - 1) stg_u4(o + ((size_t)sc * oW + sc) * p.out_ld, zero4);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct ConvDirectParams {
  const void* x;     // NCHW fp32 [n, Cin, H, W] (in_nchw_f32) or NHWC bf16 compact [n, H, W, ld]
  int in_nchw_f32, in_ld;
  const float* w;    // [3][3][Cin][Cout] fp32
  const float* bias; // [Cout]
  bf16* out;         // NHWC bf16, compact or haloed, row stride out_ld
  int out_ld, out_halo;
  int n, H, W, Cin, Cout, stride, silu;
};

template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_direct_kernel(const ConvDirectParams p) {
  extern __shared__ float s_w[];  // 9*Cin*COUT + COUT
  const int wcount = 9 * p.Cin * COUT;
  for (int i = threadIdx.x; i < wcount; i += blockDim.x) s_w[i] = p.w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_w[wcount + i] = p.bias[i];
  griddep_launch();
  griddep_wait();  // the weights above are constants; the activations below are not
  __syncthreads();
  const int oH = (p.H + p.stride - 1) / p.stride, oW = (p.W + p.stride - 1) / p.stride;
  const long long total = (long long)p.n * oH * oW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % oW);
    const long long t2 = idx / oW;
    const int oy = (int)(t2 % oH);
    const int img = (int)(t2 / oH);
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = s_w[wcount + co];
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * p.stride + ky - 1;
      if (iy < 0 || iy >= p.H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * p.stride + kx - 1;
        if (ix < 0 || ix >= p.W) continue;
        const float* wt = s_w + (size_t)(ky * 3 + kx) * p.Cin * COUT;
        for (int ci = 0; ci < p.Cin; ++ci) {
          float xv;
          if (p.in_nchw_f32) {
            xv = reinterpret_cast<const float*>(p.x)[(((size_t)img * p.Cin + ci) * p.H + iy) * p.W + ix];
          } else {
            xv = __bfloat162float(reinterpret_cast<const bf16*>(p.x)[(((size_t)img * p.H + iy) * p.W + ix) * p.in_ld + ci]);
          }
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv, wt[ci * COUT + co], acc[co]);
        }
      }
    }
    size_t orow;
    if (p.out_halo)
      orow = ((size_t)img * (oH + 1) + oy) * (oW + 1) + ox;
    else
      orow = ((size_t)img * oH + oy) * oW + ox;
    bf16* o = p.out + orow * p.out_ld;
#pragma unroll
    for (int co = 0; co < COUT; co += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = p.silu ? silu_f(acc[co + j]) : acc[co + j];
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]);
      u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]);
      u.w = pack_bf16x2(v[6], v[7]);
      stg_u4(o + co, u);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct LayoutParams {
  const void* src;
  void* dst;
  int n, C, H, W;
  int ld;        // token row stride (elements)
  int halo;      // token side uses the zero-haloed layout
  int f32;       // NCHW side is fp32 (else bf16)
};

// 32 pixels x 32 channels transposed through smem; block (32, 8)
__global__ void nchw_to_tokens_kernel(const LayoutParams p) {
  griddep_launch();
  griddep_wait();
  __shared__ float tile[32][33];
  const int HW = p.H * p.W;
  const int img = blockIdx.z;
  const int pix0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, pix = pix0 + threadIdx.x;
    float v = 0.f;
    if (c < p.C && pix < HW) {
      const size_t si = ((size_t)img * p.C + c) * HW + pix;
      v = p.f32 ? reinterpret_cast<const float*>(p.src)[si] : __bfloat162float(reinterpret_cast<const bf16*>(p.src)[si]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int pix = pix0 + i, c = c0 + threadIdx.x;
    if (c < p.C && pix < HW) {
      size_t row;
      if (p.halo) {
        const int y = pix / p.W, x = pix - y * p.W;
        row = ((size_t)img * (p.H + 1) + y) * (p.W + 1) + x;
      } else {
        row = (size_t)img * HW + pix;
      }
      reinterpret_cast<bf16*>(p.dst)[row * p.ld + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
  }
}

__global__ void tokens_to_nchw_kernel(const LayoutParams p) {
  griddep_launch();
  griddep_wait();
  __shared__ float tile[32][33];
  const int HW = p.H * p.W;
  const int img = blockIdx.z;
  const int pix0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int pix = pix0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (c < p.C && pix < HW) {
      size_t row;
      if (p.halo) {
        const int y = pix / p.W, x = pix - y * p.W;
        row = ((size_t)img * (p.H + 1) + y) * (p.W + 1) + x;
      } else {
        row = (size_t)img * HW + pix;
      }
      v = __bfloat162float(reinterpret_cast<const bf16*>(p.src)[row * p.ld + c]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, pix = pix0 + threadIdx.x;
    if (c < p.C && pix < HW) {
      const size_t di = ((size_t)img * p.C + c) * HW + pix;
      const float v = tile[threadIdx.x][i];
      if (p.f32)
        reinterpret_cast<float*>(p.dst)[di] = v;
      else
        reinterpret_cast<bf16*>(p.dst)[di] = __float2bfloat16(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// row-block copy (layout exchange around a sharded temporal sub-block) and bf16 axpy
// ---------------------------------------------------------------------------------------------------------
struct RowBlockCopyParams {
  const bf16* src;
  bf16* dst;
  int src_ld, dst_ld, cvec;
  const int* src_row;
  const int* dst_row;
  const int* rows;
};

// grid (x: slices of a block, y: block); each thread moves 16 bytes at a time
__global__ void __launch_bounds__(256) row_block_copy_kernel(const RowBlockCopyParams p) {
  griddep_launch();
  griddep_wait();
  const int blk = blockIdx.y;
  const long long total = (long long)p.rows[blk] * p.cvec;
  const bf16* s = p.src + (size_t)p.src_row[blk] * p.src_ld;
  bf16* d = p.dst + (size_t)p.dst_row[blk] * p.dst_ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / p.cvec;
    const int cv = (int)(i - r * p.cvec);
    stg_u4(d + r * p.dst_ld + cv * 8, ldg_nc_u4(s + r * p.src_ld + cv * 8));
  }
}

struct AxpyParams {
  const bf16* x;
  const bf16* y;
  bf16* out;
  int ld_x, ld_y, ld_out, rows, cvec;
  float scale;
};

__global__ void __launch_bounds__(256) axpy_bf16_kernel(const AxpyParams p) {
  griddep_launch();
  griddep_wait();
  const long long total = (long long)p.rows * p.cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / p.cvec;
    const int cv = (int)(i - r * p.cvec);
    const uint4 a = ldg_nc_u4(p.x + r * p.ld_x + cv * 8), b = ldg_nc_u4(p.y + r * p.ld_y + cv * 8);
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
    const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
    uint4 o;
    o.x = pack_bf16x2(fmaf(p.scale, b0.x, a0.x), fmaf(p.scale, b0.y, a0.y));
    o.y = pack_bf16x2(fmaf(p.scale, b1.x, a1.x), fmaf(p.scale, b1.y, a1.y));
    o.z = pack_bf16x2(fmaf(p.scale, b2.x, a2.x), fmaf(p.scale, b2.y, a2.y));
    o.w = pack_bf16x2(fmaf(p.scale, b3.x, a3.x), fmaf(p.scale, b3.y, a3.y));
    stg_u4(p.out + r * p.ld_out + cv * 8, o);
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_row_block_copy(const PtRowBlockCopyArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->src && a->dst && a->src_row && a->dst_row && a->rows, "pt_row_block_copy: null argument");
  PT_CHECK_ARG(a->n_blocks > 0 && a->n_blocks <= 65535 && a->cols > 0 && a->cols % 8 == 0 && a->src_ld % 8 == 0 && a->dst_ld % 8 == 0,
               "pt_row_block_copy: cols / strides must be multiples of 8, 1..65535 blocks");
  RowBlockCopyParams p;
  p.src = reinterpret_cast<const bf16*>(a->src);
  p.dst = reinterpret_cast<bf16*>(a->dst);
  p.src_ld = a->src_ld; p.dst_ld = a->dst_ld; p.cvec = a->cols / 8;
  p.src_row = a->src_row; p.dst_row = a->dst_row; p.rows = a->rows;
  int slices = (pt_num_sms() * 8 + a->n_blocks - 1) / a->n_blocks;
  if (slices < 1) slices = 1;
  if (slices > 64) slices = 64;
  dim3 grid(slices, a->n_blocks);
  pt_launch(row_block_copy_kernel, dim3(grid), dim3(256), 0, (void*)stream, 1, p);
  return pt_launched("pt_row_block_copy");
}

extern "C" int pt_axpy_bf16(const PtAxpyArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->y && a->out, "pt_axpy_bf16: null argument");
  PT_CHECK_ARG(a->rows > 0 && a->cols > 0 && a->cols % 8 == 0 && a->ld_x % 8 == 0 && a->ld_y % 8 == 0 && a->ld_out % 8 == 0,
               "pt_axpy_bf16: cols / strides must be multiples of 8");
  AxpyParams p;
  p.x = reinterpret_cast<const bf16*>(a->x);
  p.y = reinterpret_cast<const bf16*>(a->y);
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ld_x = a->ld_x; p.ld_y = a->ld_y; p.ld_out = a->ld_out; p.rows = a->rows; p.cvec = a->cols / 8;
  p.scale = a->scale;
  const long long total = (long long)a->rows * p.cvec;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)pt_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pt_launch(axpy_bf16_kernel, dim3((int)blocks), dim3(256), 0, (void*)stream, 1, p);
  return pt_launched("pt_axpy_bf16");
}


extern "C" int pt_small_linear(const PtSmallLinearArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->in && a->w && a->out, "pt_small_linear: null argument");
  PT_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "pt_small_linear: empty problem");
  SmallLinearParams p;
  p.in = a->in; p.in_ld = a->in_ld;
  p.w = reinterpret_cast<const bf16*>(a->w); p.w_ld = a->w_ld;
  p.bias = a->bias;
  p.out = a->out; p.out_ld = a->out_ld;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.act_in_silu = a->act_in_silu; p.act_out_silu = a->act_out_silu; p.accumulate = a->accumulate;
  int blocks = (a->N + 7) / 8;
  const size_t stage_bytes = (size_t)(a->M < 8 ? a->M : 8) * a->K * sizeof(float);
  if (stage_bytes <= 48 * 1024 && a->N >= 64) {
    const int cap = pt_num_sms() * 4;   // staged: 4 CTAs per SM, every warp strides over the remaining rows
    if (blocks > cap) blocks = cap;
    pt_launch(small_linear_kernel<true>, dim3(blocks), dim3(256), stage_bytes, (void*)stream, 1, p);
  }
  else
    pt_launch(small_linear_kernel<false>, dim3(blocks), dim3(256), 0, (void*)stream, 1, p);
  return pt_launched("pt_small_linear");
}

extern "C" int pt_timestep_sincos(const PtSinCosArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->out != nullptr, "pt_timestep_sincos: null argument");
  PT_CHECK_ARG(a->t != nullptr || (a->sigmas != nullptr && a->step_index != nullptr), "pt_timestep_sincos: need t or (sigmas, step_index)");
  PT_CHECK_ARG(a->M > 0 && a->dim > 0 && a->dim % 2 == 0, "pt_timestep_sincos: dim must be even");
  SinCosParams p;
  p.t = a->t; p.sigmas = a->sigmas; p.step_index = a->step_index;
  p.out = a->out; p.out_ld = a->out_ld; p.M = a->M; p.dim = a->dim;
  const int total = a->M * (a->dim / 2);
  pt_launch(sincos_kernel, dim3((total + 127) / 128), dim3(128), 0, (void*)stream, 1, p);
  return pt_launched("pt_timestep_sincos");
}

extern "C" int pt_upsample2x(const PtUpsampleArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->out, "pt_upsample2x: null argument");
  PT_CHECK_ARG(a->n > 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->C % 8 == 0, "pt_upsample2x: C must be a multiple of 8");
  UpsampleParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld;
  p.out = reinterpret_cast<bf16*>(a->out); p.out_ld = a->out_ld;
  p.n = a->n; p.H = a->H; p.W = a->W; p.C = a->C; p.halo = a->halo;
  p.scale = (a->scale == 1) ? 1 : 2;
  const int wc = a->W * (a->C / 8);
  long long rows = (long long)a->n * a->H;
  if (rows > 65535) rows = 65535;   // the kernel strides over the remaining image rows
  pt_launch(upsample2x_kernel, dim3((unsigned)((wc + 255) / 256), (unsigned)rows), dim3(256), 0, (void*)stream, 1, p);
  return pt_launched("pt_upsample2x");
}

extern "C" int pt_conv3x3_direct(const PtConvDirectArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->w && a->bias && a->out, "pt_conv3x3_direct: null argument");
  PT_CHECK_ARG(a->Cout == 16 || a->Cout == 32, "pt_conv3x3_direct: Cout must be 16 or 32");
  PT_CHECK_ARG(a->Cin >= 1 && a->Cin <= 32 && (a->stride == 1 || a->stride == 2), "pt_conv3x3_direct: Cin <= 32, stride 1|2");
  PT_CHECK_ARG(a->n > 0 && a->H > 0 && a->W > 0, "pt_conv3x3_direct: empty problem");
  ConvDirectParams p;
  p.x = a->x; p.in_nchw_f32 = a->in_nchw_f32; p.in_ld = a->in_ld;
  p.w = a->w; p.bias = a->bias;
  p.out = reinterpret_cast<bf16*>(a->out); p.out_ld = a->out_ld; p.out_halo = a->out_halo;
  p.n = a->n; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.stride = a->stride; p.silu = a->silu;
  const int oH = (a->H + a->stride - 1) / a->stride, oW = (a->W + a->stride - 1) / a->stride;
  const long long total = (long long)a->n * oH * oW;
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)pt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const size_t smem = sizeof(float) * (9 * (size_t)a->Cin * a->Cout + a->Cout);
  if (a->Cout == 16)
    pt_launch(conv3x3_direct_kernel<16>, dim3((int)blocks), dim3(128), smem, (void*)stream, 1, p);
  else
    pt_launch(conv3x3_direct_kernel<32>, dim3((int)blocks), dim3(128), smem, (void*)stream, 1, p);
  return pt_launched("pt_conv3x3_direct");
}

static int layout_common(const PtLayoutArgs* a, void* stream, bool to_tokens) {
  PT_CHECK_ARG(a != nullptr && a->nchw && a->tokens, "pt layout: null argument");
  PT_CHECK_ARG(a->n > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->ld >= a->C, "pt layout: bad shape");
  PT_CHECK_ARG(a->n <= 65535 && (a->C + 31) / 32 <= 65535, "pt layout: too many images/channels");
  LayoutParams p;
  p.src = to_tokens ? a->nchw : a->tokens;
  p.dst = to_tokens ? a->tokens : a->nchw;
  p.n = a->n; p.C = a->C; p.H = a->H; p.W = a->W; p.ld = a->ld; p.halo = a->halo; p.f32 = a->nchw_f32;
  dim3 grid((a->H * a->W + 31) / 32, (a->C + 31) / 32, a->n);
  dim3 block(32, 8);
  if (to_tokens)
    pt_launch(nchw_to_tokens_kernel, dim3(grid), dim3(block), 0, (void*)stream, 1, p);
  else
    pt_launch(tokens_to_nchw_kernel, dim3(grid), dim3(block), 0, (void*)stream, 1, p);
  return pt_launched(to_tokens ? "pt_nchw_to_tokens" : "pt_tokens_to_nchw");
}

extern "C" int pt_nchw_to_tokens(const PtLayoutArgs* a, void* stream) { return layout_common(a, stream, true); }
extern "C" int pt_tokens_to_nchw(const PtLayoutArgs* a, void* stream) { return layout_common(a, stream, false); }
