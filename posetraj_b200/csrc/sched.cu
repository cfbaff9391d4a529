// posetraj_b200 — fused CFG combine + v-prediction Euler step + next-step model input (fp32 math).
//
// Replaces, per denoise step:
//   pipeline/pipeline_stable_video_diffusion_controlnet.py:567-569   noise_pred = u + g_f * (c - u)
//   utils/scheduling_euler_discrete_karras_fix.py:481-520            v-prediction Euler update
//   pipeline/...:532-537 + scheduling...:284-285                     next latent_model_input = cat(x/sqrt(s^2+1), image_latents)
// HBM-bound: per latent element it reads 2 predictions + 1 fp32 latent and writes 1 fp32 latent (+ the bf16
// model input of the next step), i.e. F*C*H*W*(2*2+4+4) algorithmic bytes (+ the fused next-input write).
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

struct CfgEulerParams {
  const void* pred;
  int pred_ld, pred_nchw_f32;
  float* latents;
  const float* guidance;
  const float* sigmas;
  const int* step_index;
  int F, C, H, W;
  bf16* next_in;
  const float* image_latents;
  int next_ld, next_padded, mode, single_pred, row_begin, row_count;
};

// one thread per (f, y, x); C (= 4) channels handled in a short loop
__global__ void __launch_bounds__(256) cfg_euler_kernel(const CfgEulerParams p) {
  griddep_launch();
  griddep_wait();
  const int HW = p.H * p.W;
  const int total = p.F * HW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = idx / HW;
  const int pix = idx - f * HW;
  const int step = *p.step_index;
  const float sigma = p.sigmas[step];
  const float sigma_next = p.sigmas[step + 1];
  const float g = p.guidance[f];

  float newx[8];
  for (int c = 0; c < p.C; ++c) {
    const size_t li = ((size_t)f * p.C + c) * HW + pix;
    float x = p.latents[li];
    if (p.mode == 0) {
      float u, cnd;
      if (p.pred_nchw_f32) {
        const float* pr = reinterpret_cast<const float*>(p.pred);
        u = pr[li];
        cnd = p.single_pred ? u : pr[(size_t)p.F * p.C * HW + li];
      } else {
        const bf16* pr = reinterpret_cast<const bf16*>(p.pred);
        u = __bfloat162float(pr[(size_t)idx * p.pred_ld + c]);
        cnd = p.single_pred ? u : __bfloat162float(pr[((size_t)total + idx) * p.pred_ld + c]);
      }
      // same operation order as the reference scheduler (fp32)
      const float v = u + g * (cnd - u);
      const float s2p1 = sigma * sigma + 1.0f;
      const float x0 = v * (-sigma / sqrtf(s2p1)) + x / s2p1;
      const float d = (x - x0) / sigma;
      const float dt = sigma_next - sigma;
      x = x + d * dt;
      p.latents[li] = x;
    }
    newx[c] = x;
  }

  if (p.next_in != nullptr) {
    const float s_in = (p.mode == 0) ? sigma_next : sigma;
    const float inv = 1.0f / sqrtf(s_in * s_in + 1.0f);
    const int y = pix / p.W;
    const int xq = pix - y * p.W;
    for (int b = p.row_begin; b < p.row_begin + p.row_count; ++b) {
      const int lb = b - p.row_begin;  // row inside next_in (a shard holds only its own rows)
      size_t row;
      if (p.next_padded) {
        row = (size_t)(lb * p.F + f) * ((p.H + 1) * (p.W + 1)) + (size_t)y * (p.W + 1) + xq;
      } else {
        row = (size_t)(lb * p.F + f) * HW + pix;
      }
      bf16* o = p.next_in + row * p.next_ld;
      for (int c = 0; c < p.C; ++c) {
        // the reference divides: sample / ((sigma**2 + 1) ** 0.5)
        o[c] = __float2bfloat16(newx[c] * inv);
        const size_t ii = (((size_t)b * p.F + f) * p.C + c) * HW + pix;
        o[p.C + c] = __float2bfloat16(p.image_latents[ii]);
      }
    }
  }
}

__global__ void step_advance_kernel(int* step_index) {
  griddep_launch();
  griddep_wait();
  *step_index += 1;
}

}  // namespace pt

using namespace pt;

extern "C" int pt_cfg_euler_step(const PtCfgEulerArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr, "pt_cfg_euler_step: null args");
  PT_CHECK_ARG(a->latents && a->guidance && a->sigmas && a->step_index, "pt_cfg_euler_step: null pointer");
  PT_CHECK_ARG(a->mode == 1 || a->noise_pred != nullptr, "pt_cfg_euler_step: null noise_pred");
  PT_CHECK_ARG(a->F > 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->C <= 8, "pt_cfg_euler_step: bad shape (C must be 1..8)");
  PT_CHECK_ARG(a->next_in == nullptr || (a->image_latents != nullptr && a->next_ld >= 2 * a->C),
               "pt_cfg_euler_step: next_in needs image_latents and next_ld >= 2C");
  CfgEulerParams p;
  p.pred = a->noise_pred;
  p.pred_ld = a->pred_ld;
  p.pred_nchw_f32 = a->pred_nchw_f32;
  p.latents = a->latents;
  p.guidance = a->guidance;
  p.sigmas = a->sigmas;
  p.step_index = a->step_index;
  p.F = a->F; p.C = a->C; p.H = a->H; p.W = a->W;
  p.next_in = reinterpret_cast<bf16*>(a->next_in);
  p.image_latents = a->image_latents;
  p.next_ld = a->next_ld;
  p.next_padded = a->next_padded;
  p.mode = a->mode;
  p.single_pred = a->single_pred;
  p.row_begin = a->row_count > 0 ? a->row_begin : 0;
  p.row_count = a->row_count > 0 ? a->row_count : 2;
  PT_CHECK_ARG(p.row_begin >= 0 && p.row_begin + p.row_count <= 2, "pt_cfg_euler_step: row window must lie inside the CFG pair");
  const int total = a->F * a->H * a->W;
  pt_launch(cfg_euler_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, 1, p);
  return pt_launched("pt_cfg_euler_step");
}

extern "C" int pt_step_advance(int32_t* step_index, void* stream) {
  PT_CHECK_ARG(step_index != nullptr, "pt_step_advance: null");
  pt_launch(step_advance_kernel, dim3(1), dim3(1), 0, stream, 1, step_index);
  return pt_launched("pt_step_advance");
}
