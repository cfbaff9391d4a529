// posetraj_b200 — trajectory-map rasterisation, bit-exact against OpenCV.
//
// Replaces the CPU loop of scripts/run_inference_vipseg_json_repro.py:438-449 (same drawing in utils/dataset.py:
// 741-766): per frame transition k, on a black H x W canvas, for every track in order
//     cv2.line(img, p_k, p_{k+1}, (0,0,255), 3);  cv2.circle(img, p_{k+1}, 3, (0,255,0), -1)
// then BGR->RGB, a black last frame, and VaeImageProcessor.preprocess (x/255*2-1) in the pipeline (:500).
//
// OpenCV's 8-connected thick line is integer / 16.16 fixed-point work (imgproc/src/drawing.cpp: cv::line ->
// ThickLine -> FillConvexPoly + Line2 + Circle), restated here operation by operation:
//   * the centre line is clipped to the image grown by `thickness` pixels (cv::clipLine, truncating double math);
//   * the body is the convex quad p +- dp, dp = round(2.0 * normal) in 16.16, scan-filled with rounded DDA edges,
//     its outline drawn by the fixed-point DDA Line2 (clipped to the image in 16.16);
//   * both ends get a filled midpoint circle of radius 2; the track head gets a filled circle of radius 3.
// Painter's order (later tracks overwrite earlier ones) is kept WITHOUT serialising the tracks: every primitive
// writes its sequence number with atomicMax into an order map and a second pass turns "last writer" into colours.
// HBM-bound: algorithmic bytes = the F*3*H*W outputs written once (+ the 4-byte order map).
#include "common.cuh"
#include "launch.h"
#include "../../include/posetraj_b200.h"

namespace pt {

typedef long long i64;
constexpr int kXyShift = 16;
constexpr i64 kXyOne = 1ll << kXyShift;
constexpr int kThickness = 3;

struct RasterParams {
  const int* tracks;  // [K, F, 2] (x, y)
  int K, F, H, W;
  unsigned int* order;  // [F-1, H, W]
  float* out_f32;       // [F, 3, H, W] or null
  uint8_t* out_u8;      // [F, H, W, 3] or null
  int swap_per_track;   // dataset variant: channels swap after every track
};

// cv::clipLine(Size2l, Point2l&, Point2l&)
__device__ bool clip_line(i64 width, i64 height, i64& x1, i64& y1, i64& x2, i64& y2) {
  const i64 right = width - 1, bottom = height - 1;
  if (width <= 0 || height <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    i64 a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (i64)(__dmul_rn((double)(a - y1), (double)(x2 - x1)) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (i64)(__dmul_rn((double)(a - y2), (double)(x2 - x1)) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (i64)(__dmul_rn((double)(a - x1), (double)(y2 - y1)) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (i64)(__dmul_rn((double)(a - x2), (double)(y2 - y1)) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

struct Canvas {
  unsigned int* ord;
  int H, W;
  unsigned int seq;
  __device__ void put(i64 x, i64 y) const {
    if (x >= 0 && x < W && y >= 0 && y < H) atomicMax(ord + (size_t)y * W + x, seq);
  }
  __device__ void hline(i64 x1, i64 x2, i64 y) const {  // inclusive span, clipped
    if (y < 0 || y >= H) return;
    if (x1 < 0) x1 = 0;
    if (x2 >= W) x2 = W - 1;
    for (i64 x = x1; x <= x2; ++x) atomicMax(ord + (size_t)y * W + x, seq);
  }
};

// Circle(img, center, radius, color, fill = 1): midpoint circle, horizontal spans (one thread)
__device__ void circle_fill(const Canvas& cv, i64 cx, i64 cy, int radius) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    const i64 y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
    const i64 x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
    cv.hline(x11, x12, y11);
    cv.hline(x11, x12, y12);
    cv.hline(x21, x22, y21);
    cv.hline(x21, x22, y22);
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

// Line2: fixed-point DDA between two 16.16 points; the points of the walk are closed-form, so the CTA's threads
// take them round-robin.
__device__ void line2(const Canvas& cv, i64 x1, i64 y1, i64 x2, i64 y2, int tid, int nthreads) {
  if (!clip_line((i64)cv.W << kXyShift, (i64)cv.H << kXyShift, x1, y1, x2, y2)) return;
  i64 dx = x2 - x1, dy = y2 - y1;
  const i64 j = dx < 0 ? -1 : 0, ax = (dx ^ j) - j;
  const i64 i = dy < 0 ? -1 : 0, ay = (dy ^ i) - i;
  i64 x_step, y_step;
  int ecount;
  if (ax > ay) {
    dy = (dy ^ j) - j;
    if (j) { i64 t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    x_step = kXyOne;
    y_step = (dy << kXyShift) / (ax | 1);
    ecount = (int)((x2 - x1) >> kXyShift);
  } else {
    dx = (dx ^ i) - i;
    if (i) { i64 t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    x_step = (dx << kXyShift) / (ay | 1);
    y_step = kXyOne;
    ecount = (int)((y2 - y1) >> kXyShift);
  }
  x1 += kXyOne >> 1;
  y1 += kXyOne >> 1;
  if (tid == 0) cv.put((x2 + (kXyOne >> 1)) >> kXyShift, (y2 + (kXyOne >> 1)) >> kXyShift);
  if (ax > ay) {
    const i64 xs = x1 >> kXyShift;
    for (int n = tid; n <= ecount; n += nthreads) cv.put(xs + n, (y1 + (i64)n * y_step) >> kXyShift);
  } else {
    const i64 ys = y1 >> kXyShift;
    for (int n = tid; n <= ecount; n += nthreads) cv.put((x1 + (i64)n * x_step) >> kXyShift, ys + n);
  }
}

// FillConvexPoly(img, v, 4, color, LINE_8, XY_SHIFT): thread 0 walks the two DDA edges and records one span per
// scanline in shared memory, then all threads paint the spans.
__device__ void fill_convex_quad(const Canvas& cv, const i64 (&vx)[4], const i64 (&vy)[4], int* s_x1, int* s_x2,
                                 int* s_rows, int tid, int nthreads) {
  constexpr int npts = 4;
  const i64 delta = kXyOne >> 1;
  {
    i64 px = vx[npts - 1], py = vy[npts - 1];
    for (int i = 0; i < npts; ++i) {
      line2(cv, px, py, vx[i], vy[i], tid, nthreads);
      px = vx[i];
      py = vy[i];
    }
  }
  if (tid == 0) {
    int imin = 0;
    i64 ymin = vy[0], ymax = vy[0], xmin = vx[0], xmax = vx[0];
    for (int i = 0; i < npts; ++i) {
      if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
      ymax = max(ymax, vy[i]);
      xmax = max(xmax, vx[i]);
      xmin = min(xmin, vx[i]);
    }
    xmin = (xmin + delta) >> kXyShift;
    xmax = (xmax + delta) >> kXyShift;
    ymin = (ymin + delta) >> kXyShift;
    ymax = (ymax + delta) >> kXyShift;
    int y_first = 0, y_count = 0;
    if (!(xmax < 0 || ymax < 0 || xmin >= cv.W || ymin >= cv.H)) {
      ymax = min(ymax, (i64)cv.H - 1);
      int e_idx[2] = {imin, imin}, e_di[2] = {1, npts - 1};
      i64 e_x[2] = {-kXyOne, -kXyOne}, e_dx[2] = {0, 0}, e_ye[2] = {ymin, ymin};
      i64 y = ymin;
      int edges = npts;
      y_first = (int)max(ymin, (i64)0);
      do {
        for (int i = 0; i < 2; ++i) {
          if (y >= e_ye[i]) {
            int idx0 = e_idx[i], di = e_di[i];
            int idx = idx0 + di;
            if (idx >= npts) idx -= npts;
            for (; edges-- > 0;) {
              const i64 ty = (vy[idx] + delta) >> kXyShift;
              if (ty > y) {
                const i64 xs = vx[idx0], xe = vx[idx];
                e_ye[i] = ty;
                e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                e_x[i] = xs;
                e_idx[i] = idx;
                break;
              }
              idx0 = idx;
              idx += di;
              if (idx >= npts) idx -= npts;
            }
          }
        }
        if (edges < 0) break;
        if (y >= 0) {
          int left = 0, right = 1;
          if (e_x[0] > e_x[1]) { left = 1; right = 0; }
          i64 xx1 = (e_x[left] + delta) >> kXyShift;
          i64 xx2 = (e_x[right] + delta) >> kXyShift;
          int a = 0, b = -1;  // empty span
          if (xx2 >= 0 && xx1 < cv.W) {
            if (xx1 < 0) xx1 = 0;
            if (xx2 >= cv.W) xx2 = cv.W - 1;
            a = (int)xx1;
            b = (int)xx2;
          }
          s_x1[y - y_first] = a;
          s_x2[y - y_first] = b;
          y_count = (int)(y - y_first) + 1;
        }
        e_x[0] += e_dx[0];
        e_x[1] += e_dx[1];
      } while (++y <= ymax);
    }
    s_rows[0] = y_first;
    s_rows[1] = y_count;
  }
  __syncthreads();
  const int y_first = s_rows[0], y_count = s_rows[1];
  // paint: threads stride over the pixels of the spans row by row (rows are short: <= ~W pixels)
  for (int r = 0; r < y_count; ++r) {
    const int a = s_x1[r], b = s_x2[r];
    unsigned int* row = cv.ord + (size_t)(y_first + r) * cv.W;
    for (int x = a + tid; x <= b; x += nthreads) atomicMax(row + x, cv.seq);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) raster_prims_kernel(const RasterParams p) {
  extern __shared__ int s_spans[];  // [H] x1, [H] x2, [2] rows
  int* s_x1 = s_spans;
  int* s_x2 = s_spans + p.H;
  int* s_rows = s_spans + 2 * p.H;
  const int k = blockIdx.x, f = blockIdx.y;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int* t0 = p.tracks + ((size_t)k * p.F + f) * 2;
  const i64 ax0 = t0[0], ay0 = t0[1], ax1 = t0[2], ay1 = t0[3];
  Canvas cv;
  cv.ord = p.order + (size_t)f * p.H * p.W;
  cv.H = p.H;
  cv.W = p.W;
  cv.seq = 2u * (unsigned)k + 1u;  // the line of track k; its head circle gets 2k + 2

  // cv::line: clip the centre line to the image grown by `thickness` on every side
  i64 x0 = ax0 + kThickness, y0 = ay0 + kThickness, x1 = ax1 + kThickness, y1 = ay1 + kThickness;
  if (clip_line((i64)p.W + 2 * kThickness, (i64)p.H + 2 * kThickness, x0, y0, x1, y1)) {
    x0 = (x0 - kThickness) << kXyShift;
    y0 = (y0 - kThickness) << kXyShift;
    x1 = (x1 - kThickness) << kXyShift;
    y1 = (y1 - kThickness) << kXyShift;
    // ThickLine
    const double inv = 1.0 / (double)kXyOne;
    const double dx = (double)(x0 - x1) * inv, dy = (double)(y1 - y0) * inv;
    double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const i64 th = (i64)kThickness << (kXyShift - 1);
    if (fabs(r) > 2.220446049250313e-16) {
      r = ((double)th + (double)kXyOne * 0.5) / sqrt(r);
      const i64 dpx = (i64)rint(__dmul_rn(dy, r)), dpy = (i64)rint(__dmul_rn(dx, r));
      const i64 vx[4] = {x0 + dpx, x0 - dpx, x1 - dpx, x1 + dpx};
      const i64 vy[4] = {y0 + dpy, y0 - dpy, y1 - dpy, y1 + dpy};
      fill_convex_quad(cv, vx, vy, s_x1, s_x2, s_rows, tid, nthreads);
    }
    const int cap_r = (int)((th + (kXyOne >> 1)) >> kXyShift);
    if (tid == 0) circle_fill(cv, (x0 + (kXyOne >> 1)) >> kXyShift, (y0 + (kXyOne >> 1)) >> kXyShift, cap_r);
    if (tid == 32) circle_fill(cv, (x1 + (kXyOne >> 1)) >> kXyShift, (y1 + (kXyOne >> 1)) >> kXyShift, cap_r);
  }
  // cv2.circle(img, p_{k+1}, 3, green, -1) — drawn after the line of the same track, before the next track
  if (tid == 64) {
    cv.seq = 2u * (unsigned)k + 2u;
    circle_fill(cv, ax1, ay1, 3);
  }
}

// last writer -> colour.  Order value 0: background; odd: a line (BGR (0,0,255) -> RGB red); even: a head disc (green).
__global__ void __launch_bounds__(256) raster_resolve_kernel(const RasterParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = (size_t)p.F * hw;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx / hw);
    const size_t pix = idx - (size_t)f * hw;
    unsigned int v = 0;
    if (f < p.F - 1) v = p.order[(size_t)f * hw + pix];  // the last frame is the reference's black padding image
    const bool line = v != 0 && (v & 1u);
    const bool disc = v != 0 && !(v & 1u);
    // inference scripts: one BGR->RGB at the end -> lines red.  Dataset variant: the conversion runs after EVERY track,
    // so the pixels of track k are swapped (K - k) times: red iff that count is odd.
    bool first = line, last = false;  // first / last channel of the stored array
    if (line && p.swap_per_track) {
      const int k = (int)((v - 1u) >> 1);
      const bool odd = ((p.K - k) & 1) != 0;
      first = odd;
      last = !odd;
    }
    if (p.out_u8 != nullptr) {
      uint8_t* o = p.out_u8 + idx * 3;
      o[0] = first ? 255 : 0;
      o[1] = disc ? 255 : 0;
      o[2] = last ? 255 : 0;
    }
    if (p.out_f32 != nullptr) {
      float* o = p.out_f32 + (size_t)f * 3 * hw + pix;
      o[0] = first ? 1.0f : -1.0f;  // x / 255 * 2 - 1
      o[hw] = disc ? 1.0f : -1.0f;
      o[2 * hw] = last ? 1.0f : -1.0f;
    }
  }
}

}  // namespace pt

using namespace pt;

extern "C" int64_t pt_rasterize_workspace_bytes(int32_t F, int32_t H, int32_t W) {
  if (F < 1 || H < 1 || W < 1) return -1;
  return (int64_t)sizeof(unsigned int) * (F > 1 ? F - 1 : 1) * H * W;
}

extern "C" int pt_rasterize_tracks(const PtRasterArgs* a, void* stream) {
  PT_CHECK_ARG(a != nullptr && a->order != nullptr, "pt_rasterize_tracks: null argument");
  PT_CHECK_ARG(a->K == 0 || a->tracks != nullptr, "pt_rasterize_tracks: K > 0 without tracks");
  PT_CHECK_ARG(a->K >= 0 && a->K < (1 << 30) && a->F >= 1 && a->H >= 1 && a->W >= 1 && a->H <= 4096,
               "pt_rasterize_tracks: bad shape (K >= 0, F >= 1, 1 <= H <= 4096, W >= 1)");
  PT_CHECK_ARG(a->out_f32 != nullptr || a->out_u8 != nullptr, "pt_rasterize_tracks: no output requested");
  cudaStream_t st = (cudaStream_t)stream;
  RasterParams p;
  p.tracks = a->tracks;
  p.K = a->K; p.F = a->F; p.H = a->H; p.W = a->W;
  p.order = reinterpret_cast<unsigned int*>(a->order);
  p.out_f32 = a->out_f32;
  p.out_u8 = a->out_u8;
  p.swap_per_track = a->swap_per_track;
  if (a->F > 1) {
    cudaError_t e = cudaMemsetAsync(p.order, 0, sizeof(unsigned int) * (size_t)(a->F - 1) * a->H * a->W, st);
    if (e != cudaSuccess) return pt_fail(e, "pt_rasterize_tracks: memset");
    if (a->K > 0) {
      PT_CHECK_ARG(a->K <= 65535 * 32, "pt_rasterize_tracks: too many tracks");
      dim3 grid(a->K, a->F - 1);
      raster_prims_kernel<<<grid, 128, sizeof(int) * (2 * (size_t)a->H + 2), st>>>(p);
      int rc = pt_launched("pt_rasterize_tracks(primitives)");
      if (rc) return rc;
    }
  }
  const size_t total = (size_t)a->F * a->H * a->W;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)pt_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  raster_resolve_kernel<<<(int)blocks, 256, 0, st>>>(p);
  return pt_launched("pt_rasterize_tracks(resolve)");
}
