// posetraj_b200 — the two small kernels the VAE needs beyond the denoise-step library:
//   * row softmax for the single-head, head_dim = C attention of the VAE mid blocks (diffusers `Attention` with
//     heads = 1: logits are a plain [S, S] GEMM of the 512-wide q and k, so they are materialised in fp32 by
//     pt_gemm and normalised here into the bf16 A operand of the P V GEMM);
//   * `time_conv_out`, the Conv3d (3,1,1) over frames on the 3 decoded image channels, fused with the
//     token-major -> NCHW fp32 hand-off to the caller.
// Both are HBM/L2-bound and run once per video, not per denoise step.
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

struct SoftmaxParams {
  const float* in;
  bf16* out;
  int rows, cols, ld_in, ld_out;
};

// one CTA per row; the row (<= 40 KB) is read three times, the 2nd and 3rd time from L1/L2
__global__ void __launch_bounds__(256) softmax_rows_kernel(const SoftmaxParams p) {
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const float* src = p.in + (size_t)row * p.ld_in;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < p.cols; c += blockDim.x) m = fmaxf(m, src[c]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = s_red[0];
      for (int i = 1; i < 8; ++i) v = fmaxf(v, s_red[i]);
      s_bcast = v;
    }
    __syncthreads();
    m = s_bcast;
    float s = 0.f;
    for (int c = threadIdx.x; c < p.cols; c += blockDim.x) s += __expf(src[c] - m);
    s = warp_sum(s);
    __syncthreads();  // s_red / s_bcast of the max pass have been consumed
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
      for (int i = 0; i < 8; ++i) v += s_red[i];
      s_bcast = 1.0f / v;
    }
    __syncthreads();
    const float inv = s_bcast;
    bf16* dst = p.out + (size_t)row * p.ld_out;
    for (int c = threadIdx.x; c < p.cols; c += blockDim.x) dst[c] = __float2bfloat16(__expf(src[c] - m) * inv);
    __syncthreads();
  }
}

struct TimeConvParams {
  const float* in;  // [B*F*HW, ld] token-major, C channels used
  int ld;
  const float* w;   // [C, C, 3] (out, in, kt)
  const float* bias;
  float* out;       // [B*F, C, H, W]
  int B, F, HW, C;
};

__global__ void __launch_bounds__(256) time_conv3_kernel(const TimeConvParams p) {
  griddep_launch();
  griddep_wait();
  const long long total = (long long)p.B * p.F * p.HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % p.HW);
    const long long bf = idx / p.HW;
    const int f = (int)(bf % p.F);
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = c < p.C ? p.bias[c] : 0.f;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int ff = f + kt - 1;
      if (ff < 0 || ff >= p.F) continue;  // zero padding in time
      const float* src = p.in + (size_t)(idx + (long long)(kt - 1) * p.HW) * p.ld;
      for (int ci = 0; ci < p.C; ++ci) {
        const float v = src[ci];
        for (int c = 0; c < p.C; ++c) acc[c] = fmaf(p.w[(c * p.C + ci) * 3 + kt], v, acc[c]);
      }
    }
    for (int c = 0; c < p.C; ++c) p.out[((size_t)bf * p.C + c) * p.HW + pix] = acc[c];
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_softmax_rows(const PtSoftmaxArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->in && a->out, "pt_softmax_rows: null argument");
  PT_CHECK_ARG(a->rows > 0 && a->cols > 0 && a->ld_in >= a->cols && a->ld_out >= a->cols, "pt_softmax_rows: bad shape");
  SoftmaxParams p;
  p.in = a->in;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.rows = a->rows; p.cols = a->cols; p.ld_in = a->ld_in; p.ld_out = a->ld_out;
  const int cap = pt_num_sms() * 8;
  const int blocks = a->rows < cap ? a->rows : cap;
  pt_launch(softmax_rows_kernel, dim3(blocks), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_softmax_rows");
}

extern "C" int pt_time_conv3(const PtTimeConvArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->in && a->w && a->bias && a->out, "pt_time_conv3: null argument");
  PT_CHECK_ARG(a->C >= 1 && a->C <= 4 && a->ld >= a->C, "pt_time_conv3: 1..4 channels");
  PT_CHECK_ARG(a->B > 0 && a->F > 0 && a->HW > 0, "pt_time_conv3: empty problem");
  TimeConvParams p;
  p.in = a->in; p.ld = a->ld; p.w = a->w; p.bias = a->bias; p.out = a->out;
  p.B = a->B; p.F = a->F; p.HW = a->HW; p.C = a->C;
  const long long total = (long long)a->B * a->F * a->HW;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)pt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  pt_launch(time_conv3_kernel, dim3((int)blocks), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_time_conv3");
}
