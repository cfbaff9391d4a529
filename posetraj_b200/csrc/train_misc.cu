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

__global__ void __launch_bounds__(256) small_linear_dx_kernel(const SlBwdParams p) {
  // one CTA per (m, block of 256 k); the dy row is staged through shared memory in chunks
  __shared__ float sdy[1024];
  const int m = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  float acc = 0.f;
  for (int n0 = 0; n0 < p.N; n0 += 1024) {
    const int nn = min(1024, p.N - n0);
    __syncthreads();
    for (int j = threadIdx.x; j < nn; j += 256) sdy[j] = p.dy[(size_t)m * p.dy_ld + n0 + j];
    __syncthreads();
    if (k < p.K) {
      for (int j = 0; j < nn; ++j) acc = fmaf(sdy[j], __bfloat162float(p.w[(size_t)(n0 + j) * p.w_ld + k]), acc);
    }
  }
  if (k < p.K) {
    if (p.act_in_silu) acc *= silu_grad(p.x[(size_t)m * p.x_ld + k]);
    float* o = p.dx + (size_t)m * p.dx_ld + k;
    *o = p.accumulate_dx ? *o + acc : acc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// grouped column sums, two stages.  group(row):
//   mode 1: row / ga                                  (PtGemmArgs.rowvec_mode 1; bias: ga = rows)
//   mode 2: ((row / ga) * gb + row % gb) % gc         (PtGemmArgs.rowvec_mode 2)
//   mode 3: (row / ga) % gc                           (frame index of row (b*F + f)*HW + s: ga = HW, gc = F)
// ------------------------------------------------------------------------------------------------------------
struct ColsumGParams {
  const bf16* x;
  int ld;
  long long rows;
  int C, groups, mode, ga, gb, gc;
  int rows_per_cta;
  float* partials;   // [chunks][groups][C]
};

constexpr int kCsMaxGroups = 40;

PT_DEVICE int cs_group(const ColsumGParams& p, long long row) {
  if (p.mode == 1) return (int)(row / p.ga);
  if (p.mode == 2) return (int)((((row / p.ga) * p.gb) + (row % p.gb)) % p.gc);
  return (int)((row / p.ga) % p.gc);
}

__global__ void __launch_bounds__(256) colsum_grouped_kernel(const ColsumGParams p) {
  extern __shared__ float cs_acc[];  // [groups][8][32]
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  for (int i = threadIdx.x; i < p.groups * 256; i += 256) cs_acc[i] = 0.f;
  __syncthreads();
  const long long r_begin = (long long)blockIdx.y * p.rows_per_cta;
  long long r_end = r_begin + p.rows_per_cta;
  if (r_end > p.rows) r_end = p.rows;
  if (c < p.C) {
    int cur = -1;
    float acc = 0.f;
    for (long long r = r_begin + rl; r < r_end; r += 8) {
      const int g = cs_group(p, r);
      if (g != cur) {
        if (cur >= 0) cs_acc[(cur * 8 + rl) * 32 + cl] += acc;   // (group, rl, cl) belongs to this thread alone
        cur = g;
        acc = 0.f;
      }
      acc += __bfloat162float(p.x[(size_t)r * p.ld + c]);
    }
    if (cur >= 0) cs_acc[(cur * 8 + rl) * 32 + cl] += acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.groups * 32; i += 256) {
    const int g = i >> 5, cc = i & 31;
    if (blockIdx.x * 32 + cc < p.C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += cs_acc[(g * 8 + k) * 32 + cc];
      p.partials[((size_t)blockIdx.y * p.groups + g) * p.C + blockIdx.x * 32 + cc] = t;
    }
  }
}

__global__ void __launch_bounds__(256) colsum_fold_kernel(const float* partials, int chunks, long long n, float scale, float* out, int out_ld,
                                                          int C, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int b = 0; b < chunks; ++b) t += partials[(size_t)b * n + i];
    const long long g = i / C;
    const int c = (int)(i - g * C);
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
    pt_launch(small_linear_dx_kernel, dim3((a->K + 255) / 256, a->M), dim3(256), 0, stream, 1, p);
    return pt_launched("pt_small_linear_bwd (dx)");
  }
  return 0;
}

extern "C" int64_t pt_colsum_grouped_workspace_bytes(int64_t rows, int32_t groups, int32_t C) {
  if (rows <= 0 || groups <= 0 || C <= 0) return -1;
  long long chunks = (rows + 1023) / 1024;
  if (chunks > 592) chunks = 592;
  return (int64_t)chunks * groups * C * (int64_t)sizeof(float);
}

extern "C" int pt_colsum_grouped(const PtColsumGroupedArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->x && a->out && a->workspace && a->rows > 0 && a->C > 0, "pt_colsum_grouped: bad argument");
  PT_CHECK_ARG(a->groups > 0 && a->groups <= kCsMaxGroups, "pt_colsum_grouped: 1..40 groups");
  PT_CHECK_ARG(a->mode >= 1 && a->mode <= 3 && a->ga > 0 && (a->mode == 1 || a->gc > 0) && (a->mode != 2 || a->gb > 0), "pt_colsum_grouped: bad grouping");
  ColsumGParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ld = a->ld; p.rows = a->rows; p.C = a->C; p.groups = a->groups;
  p.mode = a->mode; p.ga = a->ga; p.gb = a->gb; p.gc = a->gc;
  long long chunks = (a->rows + 1023) / 1024;
  if (chunks > 592) chunks = 592;
  p.rows_per_cta = (int)((a->rows + chunks - 1) / chunks);
  p.partials = reinterpret_cast<float*>(a->workspace);
  const size_t smem = (size_t)a->groups * 256 * sizeof(float);
  pt_launch(colsum_grouped_kernel, dim3((a->C + 31) / 32, (unsigned)chunks), dim3(256), smem, stream, 1, p);
  int rc = pt_launched("pt_colsum_grouped");
  if (rc != 0) return rc;
  const long long n = (long long)a->groups * a->C;
  pt_launch(colsum_fold_kernel, dim3(grid_1d(n, 256)), dim3(256), 0, stream, 1, (const float*)p.partials, (int)chunks, n, a->scale, a->out,
            (int)a->out_ld, (int)a->C, (int)a->accumulate);
  return pt_launched("pt_colsum_grouped (fold)");
}
