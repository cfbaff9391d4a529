"""Full-size timing of the image-conditioning branch on one B200: 320x576 image -> anti-aliased resize -> CLIP ViT-H/14
vision tower (632 M random-init parameters) -> image_embeds; CUDA events, median of 5, per-class split."""
import json
import sys

import torch

sys.path.insert(0, ".")
from posetraj_b200.clip import CLIPVisionConfig, CLIPVisionModelWithProjection  # noqa: E402

dev = torch.device("cuda", 0)
m = CLIPVisionModelWithProjection.from_random(CLIPVisionConfig(), dev, seed=0)
img = torch.rand(1, 3, 320, 576, device=dev)


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


out = m.encode_image(img)
assert torch.isfinite(out).all() and out.shape == (1, 1, 1024)
ms = timed(lambda: m.encode_image(img))
plan = m._plan
sp = torch.cuda.current_stream().cuda_stream
cls = {}
for op in list(m._resize.values())[0].ops + plan.ops:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); op.launch(sp); e.record(); torch.cuda.synchronize()
    c = cls.setdefault(getattr(op, "kind", "misc"), {"ms": 0.0, "launches": 0})
    c["ms"] = round(c["ms"] + s.elapsed_time(e), 3); c["launches"] += 1
flops = sum(getattr(op, "alg_flops", 0.0) for op in plan.ops)
print(json.dumps({"what": "clip_encode_image", "ms": round(ms, 3), "gflop": round(flops / 1e9, 1), "launches": len(plan.ops) + 3,
                  "classes": cls}))
