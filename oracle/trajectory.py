"""ORACLE (test infrastructure) — trajectory-map rasterisation, SURVEY.md §8a row R1.

Two checkers:
  * `trajectory_maps_cv2`: the reference's own algorithm, line for line
    (/root/reference/scripts/run_inference_vipseg_json_repro.py:438-449), calling OpenCV like the reference does
    (`opencv-python==4.8.0.74` in the reference's requirements.txt:9; 4.13.0 in this image).  PINNED: it *is* the
    reference's rasteriser.
  * `thick_line` / `circle_fill` / `trajectory_maps_restated`: a pure-Python restatement of what `cv2.line(...,
    thickness=3)` and `cv2.circle(..., 3, ..., -1)` do (OpenCV imgproc/src/drawing.cpp: cv::line -> clipLine to the
    image grown by `thickness` -> ThickLine -> FillConvexPoly + Line2 + Circle; cv::circle -> Circle).  It documents
    the integer algorithm the CUDA kernel (posetraj_b200/csrc/raster.cu) follows and is pinned against cv2 on random
    and real tracks by tests/test_trajectory_cpu.py.
"""
from __future__ import annotations

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << 16


def trajectory_maps_cv2(tracks, num_frames, height, width, start=0):
    """[F, H, W, 3] uint8 RGB: num_frames-1 drawn transitions + the black padding image."""
    import cv2
    out = np.zeros((num_frames, height, width, 3), dtype=np.uint8)
    for k in range(num_frames - 1):
        mask_img = np.zeros((height, width, 3), dtype=np.uint8)
        for tr in tracks:
            a, b = tr[start + k], tr[start + k + 1]
            cv2.line(mask_img, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), (0, 0, 255), 3)
            cv2.circle(mask_img, (int(b[0]), int(b[1])), 3, (0, 255, 0), -1)
        out[k] = cv2.cvtColor(mask_img, cv2.COLOR_BGR2RGB)
    return out


def trajectory_maps_cv2_dataset(tracks, num_frames, height, width, start=0):
    """utils/dataset.py:741-766 line for line: the cvtColor sits INSIDE the track loop (channels swap once per track)."""
    import cv2
    out = np.zeros((num_frames, height, width, 3), dtype=np.uint8)
    for k in range(num_frames - 1):
        mask_img = np.zeros((height, width, 3), dtype=np.uint8)
        for tr in tracks:
            a, b = tr[start + k], tr[start + k + 1]
            cv2.line(mask_img, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), (0, 0, 255), 3)
            cv2.circle(mask_img, (int(b[0]), int(b[1])), 3, (0, 255, 0), -1)
            mask_img = cv2.cvtColor(mask_img, cv2.COLOR_BGR2RGB)
        out[k] = mask_img
    return out


def preprocess(images_u8):
    """VaeImageProcessor.preprocess on the RGB images: [F, H, W, 3] uint8 -> [F, 3, H, W] float32 in [-1, 1]."""
    x = images_u8.astype(np.float32) / 255.0
    return (2.0 * x - 1.0).transpose(0, 3, 1, 2)


# ---------------------------------------------------------------------------------------------------------------
# restatement of OpenCV's integer rasterisation
# ---------------------------------------------------------------------------------------------------------------
def _trunc(v):
    return int(v)  # C cast double -> int64 truncates toward zero


def _cdiv(num, den):
    q = abs(num) // abs(den)
    return q if (num >= 0) == (den > 0) else -q  # C integer division truncates toward zero


def clip_line(width, height, x1, y1, x2, y2):
    """cv::clipLine(Size2l, Point2l&, Point2l&)."""
    right, bottom = width - 1, height - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _trunc(float(a - y1) * float(x2 - x1) / float(y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _trunc(float(a - y2) * float(x2 - x1) / float(y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _trunc(float(a - x1) * float(y2 - y1) / float(x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _trunc(float(a - x2) * float(y2 - y1) / float(x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def _hline(img, x1, x2, y, color):
    H, W = img.shape[:2]
    if 0 <= y < H:
        x1, x2 = max(x1, 0), min(x2, W - 1)
        if x1 <= x2:
            img[y, x1:x2 + 1] = color


def circle_fill(img, cx, cy, radius, color):
    """Circle(img, center, radius, color, fill=1): midpoint circle painted as horizontal spans."""
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        _hline(img, cx - dx, cx + dx, cy - dy, color)
        _hline(img, cx - dx, cx + dx, cy + dy, color)
        _hline(img, cx - dy, cx + dy, cy - dx, color)
        _hline(img, cx - dy, cx + dy, cy + dx, color)
        dy += 1
        err += plus
        plus += 2
        mask = (err <= 0) - 1
        err -= minus & mask
        dx += mask
        minus -= mask & 2


def line2(img, p1, p2, color):
    """Line2: DDA between two 16.16 fixed-point points, clipped to the image in fixed point."""
    H, W = img.shape[:2]
    ok, x1, y1, x2, y2 = clip_line(W << XY_SHIFT, H << XY_SHIFT, p1[0], p1[1], p2[0], p2[1])
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    j = -1 if dx < 0 else 0
    ax = (dx ^ j) - j
    i = -1 if dy < 0 else 0
    ay = (dy ^ i) - i

    def put(x, y):
        if 0 <= x < W and 0 <= y < H:
            img[y, x] = color

    if ax > ay:
        dy = (dy ^ j) - j
        if j:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _cdiv(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        dx = (dx ^ i) - i
        if i:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _cdiv(dx << XY_SHIFT, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x1 >>= XY_SHIFT
        while ecount >= 0:
            put(x1, y1 >> XY_SHIFT)
            x1 += 1
            y1 += y_step
            ecount -= 1
    else:
        y1 >>= XY_SHIFT
        while ecount >= 0:
            put(x1 >> XY_SHIFT, y1)
            x1 += x_step
            y1 += 1
            ecount -= 1


def fill_convex_poly(img, v, color):
    """FillConvexPoly(img, v, npts, color, LINE_8, XY_SHIFT): outline by Line2, body by two rounded DDA edges."""
    H, W = img.shape[:2]
    npts = len(v)
    delta = XY_ONE >> 1
    p0 = v[-1]
    for p in v:
        line2(img, p0, p, color)
        p0 = p
    ys = [p[1] for p in v]
    xs = [p[0] for p in v]
    imin = min(range(npts), key=lambda i: (ys[i], i))
    xmin, xmax = (min(xs) + delta) >> XY_SHIFT, (max(xs) + delta) >> XY_SHIFT
    ymin, ymax = (min(ys) + delta) >> XY_SHIFT, (max(ys) + delta) >> XY_SHIFT
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= W or ymin >= H:
        return
    ymax = min(ymax, H - 1)
    edge = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=npts - 1, x=-XY_ONE, dx=0, ye=ymin)]
    y, edges = ymin, npts
    while True:
        for e in edge:
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = (idx0 + di) % npts
                while True:
                    edges -= 1
                    if edges + 1 <= 0:  # for (; edges-- > 0; )
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs_, xe = v[idx0][0], v[idx][0]
                        e["ye"] = ty
                        e["dx"] = _cdiv((xe - xs_) * 2 + (ty - y), 2 * (ty - y))
                        e["x"] = xs_
                        e["idx"] = idx
                        break
                    idx0 = idx
                    idx = (idx + di) % npts
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if edge[0]["x"] > edge[1]["x"] else (0, 1)
            xx1 = (edge[left]["x"] + delta) >> XY_SHIFT
            xx2 = (edge[right]["x"] + delta) >> XY_SHIFT
            if xx2 >= 0 and xx1 < W:
                _hline(img, xx1, xx2, y, color)
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


def thick_line(img, p0, p1, color, thickness=3):
    """cv2.line(img, p0, p1, color, thickness) for thickness > 1, LINE_8, integer points."""
    H, W = img.shape[:2]
    ok, a0, b0, a1, b1 = clip_line(W + 2 * thickness, H + 2 * thickness, p0[0] + thickness, p0[1] + thickness,
                                   p1[0] + thickness, p1[1] + thickness)
    if not ok:
        return
    x0, y0 = (a0 - thickness) << XY_SHIFT, (b0 - thickness) << XY_SHIFT
    x1, y1 = (a1 - thickness) << XY_SHIFT, (b1 - thickness) << XY_SHIFT
    inv = 1.0 / XY_ONE
    dx, dy = (x0 - x1) * inv, (y1 - y0) * inv
    r = dx * dx + dy * dy
    odd = thickness & 1
    th = thickness << (XY_SHIFT - 1)
    if abs(r) > 2.220446049250313e-16:
        r = (th + odd * XY_ONE * 0.5) / np.sqrt(r)
        dpx, dpy = int(np.rint(dy * r)), int(np.rint(dx * r))  # cvRound: half to even
        fill_convex_poly(img, [(x0 + dpx, y0 + dpy), (x0 - dpx, y0 - dpy), (x1 - dpx, y1 - dpy), (x1 + dpx, y1 + dpy)], color)
    for px, py in ((x0, y0), (x1, y1)):
        circle_fill(img, (px + (XY_ONE >> 1)) >> XY_SHIFT, (py + (XY_ONE >> 1)) >> XY_SHIFT,
                    (th + (XY_ONE >> 1)) >> XY_SHIFT, color)


def trajectory_maps_restated(tracks, num_frames, height, width, start=0):
    out = np.zeros((num_frames, height, width, 3), dtype=np.uint8)
    for k in range(num_frames - 1):
        img = out[k]
        for tr in tracks:
            a, b = tr[start + k], tr[start + k + 1]
            thick_line(img, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), (255, 0, 0))   # RGB red == BGR (0,0,255)
            circle_fill(img, int(b[0]), int(b[1]), 3, (0, 255, 0))
    return out
