"""The two example scripts (the reference's inference script core and training loop core on posetraj_b200) run end to end
on small random-init models."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inference_example(cuda_dev, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import infer_trajectory
    out = str(tmp_path / "frames.npz")
    video = infer_trajectory.main(["--json", os.path.join(ROOT, "tests", "golden", "traj_9_E0zfiF4DCt8.json"), "--small", "--steps", "3",
                                   "--height", "128", "--width", "192", "--tracks", "4", "--out", out])
    z = np.load(out)
    assert z["frames"].shape == (14, 128, 192, 3) and np.isfinite(z["frames"]).all()
    assert z["trajectory_maps"].shape == (14, 128, 192, 3) and z["trajectory_maps"][:13].any() and not z["trajectory_maps"][13].any()
    assert video.shape == z["frames"].shape


def test_training_example(cuda_dev, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_controlnet
    losses = train_controlnet.main(["--small", "--steps", "4", "--frames", "3", "--height", "16", "--width", "24", "--bbox", "--lr", "1e-4",
                                    "--out", str(tmp_path)])
    assert len(losses) == 4 and all(torch.isfinite(x).all() for x in losses)
    from posetraj_b200.models import ControlNetSDVModel
    again = ControlNetSDVModel.from_pretrained(str(tmp_path), subfolder="controlnet", device=cuda_dev)
    assert again.flags["bbox"]
