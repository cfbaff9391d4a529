"""configs[3] whole training step at full size on one GPU (bench.train_step_leg), with a per-phase split."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from posetraj_b200.config import SVDConfig
from posetraj_b200.models import UNetSpatioTemporalConditionControlNetModel

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = SVDConfig()
unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
res = bench.train_step_leg(cfg, unet, dev, 0, 1, torch.cuda.synchronize, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
print(json.dumps(res))
