"""Mints tests/golden/model_golden.safetensors: seeded inputs and the fp32 CPU ORACLE's outputs for one denoise step of
the small SVD-shaped config (the reference ships no golden vectors for the blocks, so the build
pins its own).  The GPU tests compare the kernels with these committed vectors (and with the live oracle); the CPU
tests check that the oracle still reproduces them, so an accidental change of the oracle's semantics is caught.
Run:  python tests/golden/gen_model_golden.py"""
import math
import os
import sys

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import make_small_inputs, oracle_pair, small_cfg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.safetensors")


def main():
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=0, cam=True)
    inp = make_small_inputs(cfg)
    sigma = 10.0
    x = torch.cat([torch.cat([inp["latents"]] * 2) / (sigma ** 2 + 1) ** 0.5, inp["image_latents"]], dim=2)
    t = torch.tensor(0.25 * math.log(sigma))
    with torch.no_grad():
        down, mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                           camera_cond=inp["camera_cond"], conditioning_scale=0.8)
        pred = o_unet(x, t, inp["image_embeddings"], down_block_additional_residuals=down,
                      mid_block_additional_residual=mid, added_time_ids=inp["added_time_ids"])
    out = {"sample": x, "timestep": t.reshape(1), "noise_pred": pred, "mid_residual": mid,
           "down_residual_0": down[0], "down_residual_11": down[11]}
    save_file({k: v.contiguous() for k, v in out.items()}, OUT,
              metadata={"config": "parity_util.small_cfg()", "weights": "oracle_pair(seed=0, cam=True), bf16-valued",
                        "inputs": "make_small_inputs(cfg) (seed 1234)", "sigma": "10.0", "conditioning_scale": "0.8"})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
