"""Mints tests/golden/train_golden.safetensors: the fp32 CPU ORACLE's losses and ControlNet gradients for one training
step of the small SVD-shaped config (the reference ships no golden vectors, so the build pins its
own).  The backward kernels of the next round are held against these; the CPU suite checks that the oracle still
reproduces them.  Stored: both losses, the noise prediction, per-parameter gradient norms and the full gradient of a few
small tensors.  Run:  python tests/golden/gen_train_golden.py"""
import os
import sys

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import make_small_inputs, oracle_pair, small_cfg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_golden.safetensors")
FULL = ("controlnet_mid_block.weight", "controlnet_down_blocks.0.bias", "time_embedding.linear_2.bias",
        "controlnet_cond_embedding.conv_in.weight", "down_blocks.0.resnets.0.time_mixer.mix_factor")


def step_inputs(cfg):
    inp = make_small_inputs(cfg)
    g = torch.Generator().manual_seed(77)
    F, h, w = cfg.num_frames, inp["latents"].shape[-2], inp["latents"].shape[-1]
    return dict(latents=torch.randn(1, F, 4, h, w, generator=g) * 0.9, noise=torch.randn(1, F, 4, h, w, generator=g),
                sigmas=torch.tensor([1.7]), image_embeddings=inp["image_embeddings"][1:2],
                trajectories=inp["controlnet_condition"][:1], motion_values=torch.tensor([127.0]),
                camera_cond=inp["camera_cond"][:1])


def run():
    from oracle.train import training_step
    torch.set_num_threads(4)
    cfg = small_cfg()
    unet, cnet = oracle_pair(cfg, seed=0, cam=True)
    unet.requires_grad_(False)
    cnet.requires_grad_(True)
    out = training_step(unet, cnet, ran_idx=1, **step_inputs(cfg))
    out["loss"].backward()
    names = [n for n, _ in cnet.named_parameters()]
    norms = torch.tensor([float(p.grad.norm()) for _, p in cnet.named_parameters()])
    res = {"loss": out["loss"].detach().reshape(1), "loss_main": out["loss_main"].detach().reshape(1),
           "loss_spatial": out["loss_spatial"].detach().reshape(1), "model_pred": out["model_pred"].detach(),
           "grad_norms": norms}
    grads = dict(cnet.named_parameters())
    for n in FULL:
        res["grad." + n] = grads[n].grad.detach().clone()
    return res, names


if __name__ == "__main__":
    res, names = run()
    save_file({k: v.contiguous() for k, v in res.items()}, OUT,
              metadata={"config": "parity_util.small_cfg()", "weights": "oracle_pair(seed=0, cam=True), bf16-valued",
                        "inputs": "gen_train_golden.step_inputs (seed 77), sigma 1.7, ran_idx 1",
                        "grad_norms_order": "ControlNetSDVModel.named_parameters()", "n_params": str(len(names))})
    print("wrote", OUT, os.path.getsize(OUT), "bytes; loss", float(res["loss"]), "params", len(names))
