// posetraj_b200 — kernels of the image-conditioning branch, all once per video:
//   * the reference's anti-aliased resize (pipeline/pipeline_stable_video_diffusion_controlnet.py:602-712): separable
//     Gaussian blur with reflect padding, then bicubic interpolation with align_corners=True (PyTorch's A = -0.75
//     cubic convolution), written straight into the im2col rows of CLIP's 14x14 / stride-14 patch embedding;
//   * self-attention for the CLIP vision tower's shape (257 tokens, 16 heads of 80 — neither the 64-wide tcgen05
//     tiles of the UNet attention nor anything a tensor-core pipeline would pay for: 0.17 GFLOP per layer).
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

struct BlurParams {
  const float* in;
  float* out;
  const float* w;
  int planes, H, W, k, axis;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // torch 'reflect' padding (no edge repeat): -1 -> 1, n -> n-2; pad < n is checked by the host
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) blur_reflect_kernel(const BlurParams p) {
  griddep_launch();
  griddep_wait();
  const long long total = (long long)p.planes * p.H * p.W;
  const int front = (p.k - 1) / 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % p.W);
    const long long t = idx / p.W;
    const int y = (int)(t % p.H);
    const float* plane = p.in + (t / p.H) * (long long)p.H * p.W;
    float acc = 0.f;
    if (p.axis == 0) {
      const float* row = plane + (size_t)y * p.W;
      for (int i = 0; i < p.k; ++i) acc = fmaf(__ldg(p.w + i), row[reflect_idx(x + i - front, p.W)], acc);
    } else {
      for (int i = 0; i < p.k; ++i) acc = fmaf(__ldg(p.w + i), plane[(size_t)reflect_idx(y + i - front, p.H) * p.W + x], acc);
    }
    p.out[idx] = acc;
  }
}

struct BicubicParams {
  const float* in;
  int C, H, W, S, P;
  float* out_f32;
  bf16* out_patches;
  int ld;
};

// PyTorch's cubic convolution coefficients (A = -0.75) for the 4 taps around floor(x), t = frac(x)
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  float x = t + 1.0f;
  c[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  x = t;
  c[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 1.0f - t;
  c[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 2.0f - t;
  c[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

__global__ void __launch_bounds__(256) bicubic_kernel(const BicubicParams p) {
  griddep_launch();
  griddep_wait();
  const int total = p.C * p.S * p.S;
  const float sy = p.S > 1 ? (float)(p.H - 1) / (float)(p.S - 1) : 0.f;
  const float sx = p.S > 1 ? (float)(p.W - 1) / (float)(p.S - 1) : 0.f;
  const int pp = p.S / (p.P > 0 ? p.P : 1);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ox = idx % p.S;
    const int oy = (idx / p.S) % p.S;
    const int c = idx / (p.S * p.S);
    const float ry = sy * (float)oy, rx = sx * (float)ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float cy[4], cx[4];
    cubic_coeffs(ry - (float)iy, cy);
    cubic_coeffs(rx - (float)ix, cx);
    const float* plane = p.in + (size_t)c * p.H * p.W;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int yy = min(max(iy - 1 + i, 0), p.H - 1);
      const float* row = plane + (size_t)yy * p.W;
      float r = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) r = fmaf(cx[j], row[min(max(ix - 1 + j, 0), p.W - 1)], r);
      acc = fmaf(cy[i], r, acc);
    }
    if (p.out_f32 != nullptr) p.out_f32[idx] = acc;
    if (p.out_patches != nullptr) {
      const int py = oy / p.P, ky = oy - py * p.P, px = ox / p.P, kx = ox - px * p.P;
      p.out_patches[(size_t)(py * pp + px) * p.ld + (c * p.P + ky) * p.P + kx] = __float2bfloat16(acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// attention for short sequences: one CTA per (head, block of 32 queries); K^T and V of the head in shared memory
// (bf16), one warp per query at a time: lanes own keys for q.k and the softmax, head dims for p.v
// ---------------------------------------------------------------------------------------------------------
struct AttnSmallParams {
  const bf16* qkv;
  int ld;
  bf16* out;
  int out_ld;
  int S, heads, hd, Sp;  // Sp: S rounded up to 32
  float scale;
};

constexpr int kSmallQPerCta = 32;

__global__ void __launch_bounds__(256) attn_small_kernel(const AttnSmallParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  // layout: Kt [hd][Sp] bf16 | V [S][hd] bf16 | per-warp probabilities [8][Sp] fp32 | per-warp query [8][hd] fp32
  bf16* sKt = reinterpret_cast<bf16*>(smem_attn);
  bf16* sV = sKt + (size_t)p.hd * p.Sp;
  float* sP = reinterpret_cast<float*>(sV + (size_t)p.Sp * p.hd);
  float* sQ = sP + 8 * p.Sp;
  const int head = blockIdx.x;
  const int q0 = blockIdx.y * kSmallQPerCta;
  const int C = p.heads * p.hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();
  griddep_wait();
  for (int i = threadIdx.x; i < p.Sp * p.hd; i += blockDim.x) {
    const int key = i / p.hd, d = i - key * p.hd;
    bf16 kv = __float2bfloat16(0.f), vv = kv;
    if (key < p.S) {
      kv = p.qkv[(size_t)key * p.ld + C + head * p.hd + d];
      vv = p.qkv[(size_t)key * p.ld + 2 * C + head * p.hd + d];
    }
    sKt[(size_t)d * p.Sp + key] = kv;
    sV[(size_t)key * p.hd + d] = vv;
  }
  __syncthreads();
  const int nk = p.Sp >> 5;  // keys per lane
  float* myP = sP + warp * p.Sp;
  float* myQ = sQ + warp * p.hd;
  for (int qi = warp; qi < kSmallQPerCta; qi += 8) {
    const int q = q0 + qi;
    if (q >= p.S) break;
    for (int d = lane; d < p.hd; d += 32) myQ[d] = __bfloat162float(p.qkv[(size_t)q * p.ld + head * p.hd + d]) * p.scale;
    __syncwarp();
    float m = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      const int key = lane + 32 * j;
      float s = 0.f;
      for (int d = 0; d < p.hd; ++d) s = fmaf(myQ[d], __bfloat162float(sKt[(size_t)d * p.Sp + key]), s);
      if (key >= p.S) s = -INFINITY;
      myP[key] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = 0; j < nk; ++j) {
      const int key = lane + 32 * j;
      const float e = __expf(myP[key] - m);   // exp(-inf) = 0 for the padded keys
      myP[key] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < p.hd; d += 32) {
      float o = 0.f;
      for (int key = 0; key < p.S; ++key) o = fmaf(myP[key], __bfloat162float(sV[(size_t)key * p.hd + d]), o);
      p.out[(size_t)q * p.out_ld + head * p.hd + d] = __float2bfloat16(o * inv);
    }
    __syncwarp();
  }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_blur_reflect(const PtBlurArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->in && a->out && a->w, "pt_blur_reflect: null argument");
  PT_CHECK_ARG(a->planes > 0 && a->H > 0 && a->W > 0 && a->k >= 1 && (a->axis == 0 || a->axis == 1), "pt_blur_reflect: bad shape");
  PT_CHECK_ARG(a->k - 1 - (a->k - 1) / 2 < (a->axis == 0 ? a->W : a->H), "pt_blur_reflect: reflect padding needs pad < size");
  BlurParams p;
  p.in = a->in; p.out = a->out; p.w = a->w; p.planes = a->planes; p.H = a->H; p.W = a->W; p.k = a->k; p.axis = a->axis;
  const long long total = (long long)a->planes * a->H * a->W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)pt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  pt_launch(blur_reflect_kernel, dim3((int)blocks), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_blur_reflect");
}

extern "C" int pt_bicubic_resize(const PtBicubicArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->in && (a->out_f32 || a->out_patches), "pt_bicubic_resize: null argument");
  PT_CHECK_ARG(a->C > 0 && a->H > 0 && a->W > 0 && a->S > 0, "pt_bicubic_resize: bad shape");
  PT_CHECK_ARG(a->out_patches == nullptr || (a->P > 0 && a->S % a->P == 0 && a->ld >= a->C * a->P * a->P),
               "pt_bicubic_resize: patch rows need P dividing S and ld >= C*P*P");
  BicubicParams p;
  p.in = a->in; p.C = a->C; p.H = a->H; p.W = a->W; p.S = a->S; p.P = a->P > 0 ? a->P : 1;
  p.out_f32 = a->out_f32; p.out_patches = reinterpret_cast<bf16*>(a->out_patches); p.ld = a->ld;
  const int total = a->C * a->S * a->S;
  pt_launch(bicubic_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_bicubic_resize");
}

extern "C" int pt_attention_small(const PtAttnSmallArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->qkv && a->out, "pt_attention_small: null argument");
  PT_CHECK_ARG(a->S > 0 && a->heads > 0 && a->head_dim > 0 && a->head_dim <= 128, "pt_attention_small: head_dim must be 1..128");
  AttnSmallParams p;
  p.qkv = reinterpret_cast<const bf16*>(a->qkv); p.ld = a->ld;
  p.out = reinterpret_cast<bf16*>(a->out); p.out_ld = a->out_ld;
  p.S = a->S; p.heads = a->heads; p.hd = a->head_dim; p.Sp = (a->S + 31) / 32 * 32;
  p.scale = 1.0f / sqrtf((float)a->head_dim);
  const size_t smem = (size_t)2 * p.hd * p.Sp * 2 + (size_t)8 * p.Sp * 4 + (size_t)8 * p.hd * 4;
  PT_CHECK_ARG(smem <= 200 * 1024, "pt_attention_small: sequence too long for the shared-memory K/V (use pt_attention_spatial)");
  static bool attr_set[PT_MAX_DEVICES] = {false};  // cudaFuncSetAttribute is per device
  const int dev_slot = pt_device_slot();
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return pt_fail(e, "pt_attention_small: cudaFuncSetAttribute");
    attr_set[dev_slot] = true;
  }
  dim3 grid(a->heads, (a->S + kSmallQPerCta - 1) / kSmallQPerCta);
  pt_launch(attn_small_kernel, grid, dim3(256), smem, stream, 1, p);
  return pt_launched("pt_attention_small");
}
