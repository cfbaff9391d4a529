"""Generates tests/golden/resize_golden.pt by EXECUTING the reference's own resize functions
(/root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:602-712).  The module itself cannot be imported
(it imports diffusers at the top), so the six pure-torch function definitions are cut out of the file by AST and
exec'd unchanged.  Run in the build container (the reference tree does not travel to the GPU box)."""
import ast
import os

import torch

SRC = "/root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py"
WANT = {"_resize_with_antialiasing", "_compute_padding", "_filter2d", "_gaussian", "_gaussian_blur2d"}

tree = ast.parse(open(SRC).read())
ns = {"torch": torch}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in WANT:
        exec(compile(ast.Module([node], []), SRC, "exec"), ns)
resize = ns["_resize_with_antialiasing"]

g = torch.Generator().manual_seed(0)
cases = {}
for name, (h, w) in {"320x576": (320, 576), "224x224": (224, 224), "250x300": (250, 300), "576x1024": (576, 1024),
                     "97x131": (97, 131)}.items():
    x = torch.rand(1, 3, h, w, generator=g)
    y = resize(x, (224, 224))
    # keep the fixture small: full output for one case, a strided sample + checksum for the others
    cases[name] = {"input_seed_order": list(cases).__len__(), "shape": (h, w), "sample": y[:, :, ::7, ::5].clone(),
                   "sum": float(y.double().sum()), "abs_sum": float(y.double().abs().sum())}
out = os.path.join(os.path.dirname(__file__), "resize_golden.pt")
torch.save(cases, out)
print(out, {k: v["sum"] for k, v in cases.items()})
