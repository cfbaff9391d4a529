// posetraj_b200 — the small backward kernels of the whole-network training step (BASELINE configs[3])
// that pt_gemm / pt_wgrad / train.cu do not cover.  Reference: autograd through the layers named below during
// `accelerator.backward(loss)`, scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1470.  All HBM-bound, vectorised 16 bytes
// per thread where the layout allows; reductions are two-stage with a fixed order (deterministic).
//
//   pt_upsample2x_bwd     Upsample2D nearest x2 / the zero-halo copy (Downsample2D input): sum of the 2x2 children / strip
//   pt_dilate2x           stride-2 conv (Downsample2D, cond-embedding blocks) output gradient -> the zero-haloed
//                         full-resolution row space its dgrad / wgrad run in (value at even pixels, zero elsewhere)
//   pt_zero_halo          zero the halo rows of a zero-haloed buffer (dgrad writes garbage there)
//   pt_silu_fwd / _bwd    SiLU as its own pass (the conditioning embedding keeps its pre-activations in training)
//   pt_small_linear_bwd   time / frame-position / image-embedding MLPs (fp32 [M <= 64, K] rows): dx, dW, db
//   pt_colsum_grouped     bias, time-embedding and cross-attention-constant gradients: per-group column sums where the
//                         group of a row follows PtGemmArgs.rowvec_mode (incl. the reference's mis-aligned temporal
//                         context broadcast) or the frame index (frame position embedding)
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

static inline unsigned grid_1d(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

PT_DEVICE uint4 add_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = unpack_bf16x2(pa[i]), y = unpack_bf16x2(pb[i]);
    pr[i] = pack_bf16x2(x.x + y.x, x.y + y.y);
  }
  return r;
}

// ------------------------------------------------------------------------------------------------------------
struct UpBwdParams {
  const bf16* dout;
  int dout_ld;
  bf16* dx;
  int dx_ld;
  int n, H, W, C, halo, scale;
};

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const UpBwdParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.n * p.H * p.W * cv;
  const int oW = p.scale * p.W + (p.halo ? 1 : 0), oH = p.scale * p.H + (p.halo ? 1 : 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long pix = i / cv;
    const int x = (int)(pix % p.W);
    const int y = (int)((pix / p.W) % p.H);
    const long long img = pix / ((long long)p.W * p.H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < p.scale; ++dy)
      for (int dx = 0; dx < p.scale; ++dx) {
        const long long orow = (img * oH + (p.scale * y + dy)) * oW + (p.scale * x + dx);
        const uint4 u = ldg_u4(p.dout + (size_t)orow * p.dout_ld + c * 8);
        const uint32_t* pu = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(pu[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    stg_u4(p.dx + (size_t)pix * p.dx_ld + c * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
struct DilateParams {
  const bf16* src;   // [n, oH(+1), oW(+1)] rows
  int src_ld, src_halo;
  bf16* dst;         // [n, H+1, W+1] rows, fully written
  int dst_ld;
  int n, H, W, oH, oW, C;
};

__global__ void __launch_bounds__(256) dilate2x_kernel(const DilateParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.n * (p.H + 1) * (p.W + 1) * cv;
  const int sW = p.oW + (p.src_halo ? 1 : 0), sH = p.oH + (p.src_halo ? 1 : 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long row = i / cv;
    const int x = (int)(row % (p.W + 1));
    const int y = (int)((row / (p.W + 1)) % (p.H + 1));
    const long long img = row / ((long long)(p.W + 1) * (p.H + 1));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y < p.H && x < p.W && !(y & 1) && !(x & 1) && (y >> 1) < p.oH && (x >> 1) < p.oW)
      v = ldg_u4(p.src + (size_t)((img * sH + (y >> 1)) * sW + (x >> 1)) * p.src_ld + c * 8);
    stg_u4(p.dst + (size_t)row * p.dst_ld + c * 8, v);
  }
}

__global__ void __launch_bounds__(256) zero_halo_kernel(bf16* x, int ld, int n, int H, int W, int C) {
  const int cv = C >> 3;
  const int per_img = H + W + 1;   // column W of rows 0..H-1, then the whole row H
  const long long total = (long long)n * per_img * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long e = i / cv;
    const int k = (int)(e % per_img);
    const long long img = e / per_img;
    const int y = k < H ? k : H;
    const int xx = k < H ? W : k - H;
    stg_u4(x + (size_t)((img * (H + 1) + y) * (W + 1) + xx) * ld + c * 8, make_uint4(0u, 0u, 0u, 0u));
  }
}

// ------------------------------------------------------------------------------------------------------------
PT_DEVICE float silu_exact(float x) { return x / (1.0f + __expf(-x)); }
PT_DEVICE float silu_grad(float x) {
  const float s = 1.0f / (1.0f + __expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

// mode 0: out = silu(x); mode 1: out = dy * silu'(x)
__global__ void __launch_bounds__(256) silu_kernel(const bf16* x, int ld, const bf16* dy, int dy_ld, bf16* out, int out_ld, long long rows,
                                                   int cols, int mode) {
  const int cv = cols >> 3;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long r = i / cv;
    const uint4 u = ldg_u4(x + (size_t)r * ld + c * 8);
    const uint32_t* pu = reinterpret_cast<const uint32_t*>(&u);
    uint4 o;
    uint32_t* po = reinterpret_cast<uint32_t*>(&o);
    if (mode == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(pu[k]);
        po[k] = pack_bf16x2(silu_exact(f.x), silu_exact(f.y));
      }
    } else {
      const uint4 g = ldg_u4(dy + (size_t)r * dy_ld + c * 8);
      const uint32_t* pg = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(pu[k]), d = unpack_bf16x2(pg[k]);
        po[k] = pack_bf16x2(d.x * silu_grad(f.x), d.y * silu_grad(f.y));
      }
    }
    stg_u4(out + (size_t)r * out_ld + c * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// y = W act(x) + b (PtSmallLinearArgs without act_out): dW[n,k] (+)= sum_m dy[m,n] act(x[m,k]); db[n] (+)= sum_m dy[m,n];
// dx[m,k] (+)= act'(x[m,k]) * sum_n dy[m,n] W[n,k]
// ------------------------------------------------------------------------------------------------------------
struct SlBwdParams {
  const float* x;
  int x_ld;
  const bf16* w;
  int w_ld;
  const float* dy;
  int dy_ld;
  int M, N, K, act_in_silu;
  float* dx;
  int dx_ld, accumulate_dx;
  float* dw;   // [N, K] contiguous
  float* db;
  int accumulate_w;
};

__global__ void __launch_bounds__(256) small_linear_dw_kernel(const SlBwdParams p) {
  const long long total = (long long)p.N * p.K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % p.K);
    const int n = (int)(i / p.K);
    float acc = 0.f;
    for (int m = 0; m < p.M; ++m) {
      float u = p.x[(size_t)m * p.x_ld + k];
      if (p.act_in_silu) u = silu_exact(u);
      acc = fmaf(p.dy[(size_t)m * p.dy_ld + n], u, acc);
    }
    p.dw[i] = p.accumulate_w ? p.dw[i] + acc : acc;
    if (k == 0 && p.db != nullptr) {
      float s = 0.f;
      for (int m = 0; m < p.M; ++m) s += p.dy[(size_t)m * p.dy_ld + n];
      p.db[n] = p.accumulate_w ? p.db[n] + s : s;
    }
  }
}

// grid (K / 256, M, N slabs of 1024): each CTA sums its slab of n for 256 k; slabs are folded by small_linear_dx_fold_kernel
__global__ void __launch_bounds__(256) small_linear_dx_kernel(const SlBwdParams p, float* partial) {
  __shared__ float sdy[1024];
  const int m = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int n0 = blockIdx.z * 1024;
  const int nn = min(1024, p.N - n0);
  for (int j = threadIdx.x; j < nn; j += 256) sdy[j] = p.dy[(size_t)m * p.dy_ld + n0 + j];
  __syncthreads();
  if (k >= p.K) return;
  float acc = 0.f;
  const bf16* w = p.w + (size_t)n0 * p.w_ld + k;
#pragma unroll 4
  for (int j = 0; j < nn; ++j) acc = fmaf(sdy[j], __bfloat162float(w[(size_t)j * p.w_ld]), acc);
  partial[((size_t)blockIdx.z * p.M + m) * p.K + k] = acc;
}

__global__ void __launch_bounds__(256) small_linear_dx_fold_kernel(const SlBwdParams p, const float* partial, int slabs) {
  const long long total = (long long)p.M * p.K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / p.K), k = (int)(i - (long long)m * p.K);
    float acc = 0.f;
    for (int s2 = 0; s2 < slabs; ++s2) acc += partial[(size_t)s2 * total + i];
    if (p.act_in_silu) acc *= silu_grad(p.x[(size_t)m * p.x_ld + k]);
    float* o = p.dx + (size_t)m * p.dx_ld + k;
    *o = p.accumulate_dx ? *o + acc : acc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// grouped column sums, two stages (per-CTA partials, then a fixed-order fold).  group(row):
//   mode 1: row / ga                                  (PtGemmArgs.rowvec_mode 1; bias: ga = rows)
//   mode 2: ((row / ga) * gb + row % gb) % gc         (PtGemmArgs.rowvec_mode 2; gc <= 4)
//   mode 3: (row / ga) % gc                           (frame index of row (b*F + f)*HW + s: ga = HW, gc = F)
// Modes 1 and 3 are RUNS of ga consecutive rows with one group: a CTA sums a slice of one run.  Mode 2 interleaves the
// groups row by row: a CTA keeps one accumulator set per group.  Threads read whole rows, 8 channels (16 bytes) each.
// ------------------------------------------------------------------------------------------------------------
struct ColsumGParams {
  const bf16* x;
  int ld;
  long long rows;
  int C, groups, mode, ga, gb, gc;
  int sub;            // modes 1 / 3: CTAs per run; mode 2: number of row chunks
  int rows_per_cta;
  int cvec, rpar;
  float* partials;    // [blocks][C]
};

constexpr int kCsMaxBlocks = 2048;

template <int NG>   // accumulator sets per thread: 1 (modes 1 / 3) or gc (mode 2)
__global__ void __launch_bounds__(512) colsum_grouped_kernel(const ColsumGParams p) {
  extern __shared__ float cs_sm[];  // [rpar][C]
  const int cl = threadIdx.x % p.cvec, rl = threadIdx.x / p.cvec;
  long long r_begin, r_end;
  if (p.mode == 2) {
    r_begin = (long long)blockIdx.x * p.rows_per_cta;
    r_end = r_begin + p.rows_per_cta;
    if (r_end > p.rows) r_end = p.rows;
  } else {
    const long long run = blockIdx.x / p.sub;
    const int s = blockIdx.x % p.sub;
    r_begin = run * p.ga + (long long)s * p.rows_per_cta;
    r_end = r_begin + p.rows_per_cta;
    if (r_end > (run + 1) * p.ga) r_end = (run + 1) * p.ga;
  }
  float acc[NG][8];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
  for (long long r = r_begin + rl; r < r_end; r += p.rpar) {
    const uint4 u = ldg_u4(p.x + (size_t)r * p.ld + cl * 8);
    const uint32_t* pu = reinterpret_cast<const uint32_t*>(&u);
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(pu[k]);
      v[2 * k] = f.x;
      v[2 * k + 1] = f.y;
    }
    if (NG == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[0][i] += v[i];
    } else {
      const int g = (int)((((r / p.ga) * p.gb) + (r % p.gb)) % p.gc);
#pragma unroll
      for (int gg = 0; gg < NG; ++gg)
        if (gg == g) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[gg][i] += v[i];
        }
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g > 0) __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) cs_sm[(size_t)rl * p.C + cl * 8 + i] = acc[g][i];
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      float t = 0.f;
      for (int l = 0; l < p.rpar; ++l) t += cs_sm[(size_t)l * p.C + c];
      p.partials[((size_t)blockIdx.x * NG + g) * p.C + c] = t;
    }
  }
}

// 32 columns x 8 slices of the partial blocks per CTA; blockIdx.y = group
__global__ void __launch_bounds__(256) colsum_fold_kernel(const ColsumGParams p, int blocks, float scale, float* out, int out_ld, int accumulate) {
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const int c = blockIdx.x * 32 + cl;
  float t = 0.f;
  if (c < p.C) {
    if (p.mode == 1) {
      for (int b = g * p.sub + sl; b < (g + 1) * p.sub; b += 8) t += p.partials[(size_t)b * p.C + c];
    } else if (p.mode == 3) {
      const int runs = blocks / p.sub;
      const int per_g = (runs - g + p.gc - 1) / p.gc * p.sub;   // partial blocks of this group: run = g + gc*j, s
      for (int e = sl; e < per_g; e += 8) {
        const int run = g + p.gc * (e / p.sub), s2 = e % p.sub;
        t += p.partials[((size_t)run * p.sub + s2) * p.C + c];
      }
    } else {
      for (int b = sl; b < blocks; b += 8) t += p.partials[((size_t)b * p.gc + g) * p.C + c];
    }
  }
  sm[sl][cl] = t;
  __syncthreads();
  if (sl == 0 && c < p.C) {
    float tt = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tt += sm[k][cl];
    float* o = out + (size_t)g * out_ld + c;
    *o = accumulate ? *o + scale * tt : scale * tt;
  }
}

// scalar fallback for widths that are not multiples of 8 (one thread per column)
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const ColsumGParams p, float scale, float* out, int out_ld, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  for (int g = 0; g < p.groups; ++g) {
    float t = 0.f;
    for (long long r = 0; r < p.rows; ++r) {
      int gr;
      if (p.mode == 1) gr = (int)(r / p.ga);
      else if (p.mode == 2) gr = (int)((((r / p.ga) * p.gb) + (r % p.gb)) % p.gc);
      else gr = (int)((r / p.ga) % p.gc);
      if (gr == g) t += __bfloat162float(p.x[(size_t)r * p.ld + c]);
    }
    float* o = out + (size_t)g * out_ld + c;
    *o = accumulate ? *o + scale * t : scale * t;
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_upsample2x_bwd(const PtUpsampleArgs* a, void* stream) {
  // same struct as the forward: `x` / `ld` = the gradient to WRITE (compact [n*H*W, C]), `out` / `out_ld` = the gradient
  // of the forward output to READ
  PT_CHECK_ARG(a != nullptr && a->x && a->out && a->n > 0 && a->H > 0 && a->W > 0, "pt_upsample2x_bwd: bad argument");
  PT_CHECK_ARG(a->C % 8 == 0 && a->ld % 8 == 0 && a->out_ld % 8 == 0 && (a->scale == 1 || a->scale == 2), "pt_upsample2x_bwd: C, strides % 8, scale 1|2");
  UpBwdParams p;
  p.dout = reinterpret_cast<const bf16*>(a->out); p.dout_ld = a->out_ld;
  p.dx = const_cast<bf16*>(reinterpret_cast<const bf16*>(a->x)); p.dx_ld = a->ld;
  p.n = a->n; p.H = a->H; p.W = a->W; p.C = a->C; p.halo = a->halo; p.scale = a->scale;
  const long long total = (long long)a->n * a->H * a->W * (a->C / 8);
  pt_launch(upsample2x_bwd_kernel, dim3(grid_1d(total, 256)), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_upsample2x_bwd");
}

extern "C" int pt_dilate2x(const void* src, int32_t src_ld, int32_t src_halo, void* dst, int32_t dst_ld, int32_t n, int32_t H, int32_t W,
                           int32_t C, void* stream) {
  PT_CHECK_ARG(src && dst && n > 0 && H > 0 && W > 0 && C % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "pt_dilate2x: bad argument");
  DilateParams p;
  p.src = reinterpret_cast<const bf16*>(src); p.src_ld = src_ld; p.src_halo = src_halo;
  p.dst = reinterpret_cast<bf16*>(dst); p.dst_ld = dst_ld;
  p.n = n; p.H = H; p.W = W; p.oH = (H + 1) / 2; p.oW = (W + 1) / 2; p.C = C;
  const long long total = (long long)n * (H + 1) * (W + 1) * (C / 8);
  pt_launch(dilate2x_kernel, dim3(grid_1d(total, 256)), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_dilate2x");
}

extern "C" int pt_zero_halo(void* x, int32_t ld, int32_t n, int32_t H, int32_t W, int32_t C, void* stream) {
  PT_CHECK_ARG(x && n > 0 && H > 0 && W > 0 && C % 8 == 0 && ld % 8 == 0, "pt_zero_halo: bad argument");
  const long long total = (long long)n * (H + W + 1) * (C / 8);
  pt_launch(zero_halo_kernel, dim3(grid_1d(total, 256)), dim3(256), 0, stream, 1, reinterpret_cast<bf16*>(x), (int)ld, (int)n, (int)H, (int)W,
            (int)C);
  return pt_launched("pt_zero_halo");
}

extern "C" int pt_silu_fwd(const void* x, int32_t ld, void* out, int32_t out_ld, int64_t rows, int32_t cols, void* stream) {
  PT_CHECK_ARG(x && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && out_ld % 8 == 0, "pt_silu_fwd: bad argument");
  pt_launch(silu_kernel, dim3(grid_1d(rows * (cols / 8), 256)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(x), (int)ld,
            (const bf16*)nullptr, 0, reinterpret_cast<bf16*>(out), (int)out_ld, (long long)rows, (int)cols, 0);
  return pt_launched("pt_silu_fwd");
}

extern "C" int pt_silu_bwd(const void* x, int32_t ld, const void* dy, int32_t dy_ld, void* dx, int32_t dx_ld, int64_t rows, int32_t cols,
                           void* stream) {
  PT_CHECK_ARG(x && dy && dx && rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0, "pt_silu_bwd: bad argument");
  pt_launch(silu_kernel, dim3(grid_1d(rows * (cols / 8), 256)), dim3(256), 0, stream, 1, reinterpret_cast<const bf16*>(x), (int)ld,
            reinterpret_cast<const bf16*>(dy), (int)dy_ld, reinterpret_cast<bf16*>(dx), (int)dx_ld, (long long)rows, (int)cols, 1);
  return pt_launched("pt_silu_bwd");
}

extern "C" int64_t pt_small_linear_bwd_workspace_bytes(int32_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  return (int64_t)((N + 1023) / 1024) * M * K * (int64_t)sizeof(float);
}

extern "C" int pt_small_linear_bwd(const PtSmallLinearBwdArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->w && a->dy && a->M > 0 && a->N > 0 && a->K > 0, "pt_small_linear_bwd: bad argument");
  SlBwdParams p;
  p.x = a->x; p.x_ld = a->x_ld; p.w = reinterpret_cast<const bf16*>(a->w); p.w_ld = a->w_ld; p.dy = a->dy; p.dy_ld = a->dy_ld;
  p.M = a->M; p.N = a->N; p.K = a->K; p.act_in_silu = a->act_in_silu;
  p.dx = a->dx; p.dx_ld = a->dx_ld; p.accumulate_dx = a->accumulate_dx;
  p.dw = a->dw; p.db = a->db; p.accumulate_w = a->accumulate_w;
  if (a->dw != nullptr) {
    pt_launch(small_linear_dw_kernel, dim3(grid_1d((long long)a->N * a->K, 256, 148 * 32)), dim3(256), 0, stream, 1, p);
    int rc = pt_launched("pt_small_linear_bwd (dW)");
    if (rc != 0) return rc;
  }
  if (a->dx != nullptr) {
    PT_CHECK_ARG(a->dx_workspace != nullptr, "pt_small_linear_bwd: dx needs dx_workspace (pt_small_linear_bwd_workspace_bytes)");
    const int slabs = (a->N + 1023) / 1024;
    float* partial = reinterpret_cast<float*>(a->dx_workspace);
    pt_launch(small_linear_dx_kernel, dim3((a->K + 255) / 256, a->M, slabs), dim3(256), 0, stream, 1, p, partial);
    int rc = pt_launched("pt_small_linear_bwd (dx)");
    if (rc != 0) return rc;
    pt_launch(small_linear_dx_fold_kernel, dim3(grid_1d((long long)a->M * a->K, 256)), dim3(256), 0, stream, 1, p, (const float*)partial, slabs);
    return pt_launched("pt_small_linear_bwd (dx fold)");
  }
  return 0;
}

extern "C" int64_t pt_colsum_grouped_workspace_bytes(int64_t rows, int32_t groups, int32_t C) {
  if (rows <= 0 || groups <= 0 || C <= 0) return -1;
  return (int64_t)kCsMaxBlocks * C * (int64_t)sizeof(float);
}

extern "C" int pt_colsum_grouped(const PtColsumGroupedArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->out && a->workspace && a->rows > 0 && a->C > 0, "pt_colsum_grouped: bad argument");
  PT_CHECK_ARG(a->groups > 0, "pt_colsum_grouped: groups must be positive");
  PT_CHECK_ARG(a->mode >= 1 && a->mode <= 3 && a->ga > 0 && (a->mode == 1 || a->gc > 0) && (a->mode != 2 || a->gb > 0), "pt_colsum_grouped: bad grouping");
  if (a->C > 2048 && a->C % 8 == 0) {
    // wide rows (the GEGLU projections of levels 1-3: 5120 / 10240 columns): column blocks of 2048, one after the other
    // on the stream (they share the workspace)
    for (int c0 = 0; c0 < a->C; c0 += 2048) {
      PtColsumGroupedArgs b = *a;
      b.x = reinterpret_cast<const bf16*>(a->x) + c0;
      b.C = a->C - c0 < 2048 ? a->C - c0 : 2048;
      b.out = a->out + c0;
      const int rc = pt_colsum_grouped(&b, stream);
      if (rc != 0) return rc;
    }
    return 0;
  }
  ColsumGParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld; p.rows = a->rows; p.C = a->C; p.groups = a->groups;
  p.mode = a->mode; p.ga = a->ga; p.gb = a->gb > 0 ? a->gb : 1; p.gc = a->gc > 0 ? a->gc : 1;
  p.partials = reinterpret_cast<float*>(a->workspace);
  const bool vec = a->C % 8 == 0 && a->ld % 8 == 0 && ((uintptr_t)a->x % 16 == 0) && a->C <= 4096;
  const long long runs = (a->rows + a->ga - 1) / a->ga;
  if (!vec || (a->mode != 2 && (runs > kCsMaxBlocks || a->rows % a->ga != 0)) || (a->mode == 2 && p.gc > 4)) {
    p.sub = 1; p.rows_per_cta = 0; p.cvec = 0; p.rpar = 0;
    pt_launch(colsum_scalar_kernel, dim3((a->C + 255) / 256), dim3(256), 0, stream, 1, p, a->scale, a->out, (int)a->out_ld, (int)a->accumulate);
    return pt_launched("pt_colsum_grouped (scalar)");
  }
  p.cvec = a->C / 8;
  p.rpar = 512 / p.cvec;
  if (p.rpar < 1) p.rpar = 1;
  const int threads = p.cvec * p.rpar;
  const size_t smem = (size_t)p.rpar * a->C * sizeof(float);
  int blocks;
  if (a->mode == 2) {
    long long chunks = (a->rows + 4LL * p.rpar - 1) / (4LL * p.rpar);
    if (chunks > 512) chunks = 512;
    if (chunks < 1) chunks = 1;
    p.sub = (int)chunks;
    p.rows_per_cta = (int)((a->rows + chunks - 1) / chunks);
    blocks = (int)chunks;
    switch (p.gc) {
      case 1: pt_launch(colsum_grouped_kernel<1>, dim3(blocks), dim3(threads), smem, stream, 1, p); break;
      case 2: pt_launch(colsum_grouped_kernel<2>, dim3(blocks), dim3(threads), smem, stream, 1, p); break;
      case 3: pt_launch(colsum_grouped_kernel<3>, dim3(blocks), dim3(threads), smem, stream, 1, p); break;
      default: pt_launch(colsum_grouped_kernel<4>, dim3(blocks), dim3(threads), smem, stream, 1, p); break;
    }
  } else {
    long long sub = 592 / runs;
    const long long max_sub = ((long long)a->ga + 4LL * p.rpar - 1) / (4LL * p.rpar);
    if (sub > max_sub) sub = max_sub;
    if (sub < 1) sub = 1;
    p.sub = (int)sub;
    p.rows_per_cta = (int)(((long long)a->ga + sub - 1) / sub);
    blocks = (int)(runs * sub);
    pt_launch(colsum_grouped_kernel<1>, dim3(blocks), dim3(threads), smem, stream, 1, p);
  }
  int rc = pt_launched("pt_colsum_grouped");
  if (rc != 0) return rc;
  pt_launch(colsum_fold_kernel, dim3((a->C + 31) / 32, a->groups), dim3(256), 0, stream, 1, p, blocks, a->scale, a->out, (int)a->out_ld,
            (int)a->accumulate);
  return pt_launched("pt_colsum_grouped (fold)");
}
