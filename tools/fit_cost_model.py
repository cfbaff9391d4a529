"""Fits the per-tile cost model of posetraj_b200.ops.tile_cost_ns to a tools/gemm_sweep.py log and prints the constants,
the rms log error and how far the model's pick is from the measured optimum per shape.
Usage: python tools/fit_cost_model.py gpurun_out/r2f_sweep.log"""
import math
import sys

import numpy as np
from scipy.optimize import least_squares

NUM_SMS = 148
cols = [(False, bn) for bn in (64, 96, 128, 160, 192, 224, 256)] + [(True, bn) for bn in (64, 96, 128, 160, 192, 224, 256)]
data = []
for line in open(sys.argv[1]):
    if "|" not in line or line.startswith("kind"):
        continue
    head, cells = line.split("|")[0].split(), line.split("|")[1].split()
    kind, rows, N, K, taps, geglu, res = head[0], int(head[1]), int(head[2]), int(head[3]), int(head[4]), int(head[5]), int(head[6])
    batches = 2 if kind == "tconv" else 1
    for (pair, bn), c in zip(cols, cells):
        if c in ("-", "err"):
            continue
        data.append((rows, batches, N, taps * K // 64, bool(geglu), bool(res), pair, bn, float(c) * 1e3, (kind, rows, N, K, taps, geglu)))


def cost(x, rows, batches, n_out, k_iters, geglu, has_res, pair, bn):
    c_mma, floor_s, b_s, e0, e1, floor_p, b_p, launch = x
    tm = 256 if pair else 128
    m_tiles = batches * math.ceil(rows / batches / tm)
    per = bn // 2 if geglu else bn
    n_tiles = math.ceil(n_out / per)
    units = NUM_SMS // 2 if pair else NUM_SMS
    t_max = math.ceil(m_tiles * n_tiles / units)
    kstep = max(bn * c_mma, (floor_p + b_p * bn) if pair else (floor_s + b_s * bn))
    epi = e0 + e1 * per * (2.0 if geglu else 1.0) * (1.3 if has_res else 1.0)
    return launch + t_max * (k_iters * kstep + max(epi, 0.0))


def resid(x):
    return [math.log(cost(x, *d[:8]) / d[8]) for d in data]


x0 = [0.55, 100.0, 0.25, 100.0, 8.0, 100.0, 0.1, 3000.0]
r = least_squares(resid, x0, bounds=([0.3, 0, 0, -2000, 0, 0, 0, 0], [2, 1000, 3, 3000, 50, 1000, 3, 20000]))
x = r.x
print("constants c_mma, floor_s, b_s, e0, e1, floor_p, b_p, launch_ns =", [round(float(v), 4) for v in x])
print("rms log error", float(np.sqrt(np.mean(np.square(resid(x))))))
shapes = {}
for d in data:
    shapes.setdefault(d[9], []).append(d)
worse = []
for key, ds in shapes.items():
    best = min(ds, key=lambda d: d[8])
    pick = min(ds, key=lambda d: cost(x, *d[:8]))
    worse.append(pick[8] / best[8])
    print(key, "best", ("p" if best[6] else "s") + str(best[7]), round(best[8] / 1e3, 1), "model picks", ("p" if pick[6] else "s") + str(pick[7]),
          round(pick[8] / 1e3, 1), f"(+{100 * (pick[8] / best[8] - 1):.1f}%)")
print("mean loss of the model's pick vs the measured optimum: %.1f%%, max %.1f%%" % (100 * (np.mean(worse) - 1), 100 * (max(worse) - 1)))
