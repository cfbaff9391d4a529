"""Generates tests/golden/train_sigma_golden.json by EXECUTING the reference's own `stratified_uniform` and
`rand_cosine_interpolated` (/root/reference/scripts/train_svd_traj_VIPSeg_14_cam_concat.py:289-336), cut out of the
script by AST (the script itself imports diffusers/accelerate and cannot be imported).  Run in the build container."""
import ast
import json
import math
import os

import torch

SRC = "/root/reference/scripts/train_svd_traj_VIPSeg_14_cam_concat.py"
tree = ast.parse(open(SRC).read())
ns = {"torch": torch, "math": math}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("stratified_uniform", "rand_cosine_interpolated"):
        exec(compile(ast.Module([node], []), SRC, "exec"), ns)
cases = []
for seed, n in [(0, 1), (1, 2), (2, 16), (3, 7)]:
    torch.manual_seed(seed)
    u = torch.rand([n])           # what stratified_uniform draws first under this seed
    torch.manual_seed(seed)
    sig = ns["rand_cosine_interpolated"](shape=[n, ], image_d=64, noise_d_low=32, noise_d_high=64, sigma_data=0.5,
                                         min_value=0.002, max_value=700)
    cases.append({"seed": seed, "n": n, "u": u.tolist(), "sigmas": sig.tolist()})
out = os.path.join(os.path.dirname(__file__), "train_sigma_golden.json")
json.dump(cases, open(out, "w"), indent=1)
print(out, cases[1])
