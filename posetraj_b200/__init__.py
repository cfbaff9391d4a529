"""posetraj_b200 — B200-native (sm_100a) implementation of PoseTraj's denoising hot path behind the reference's
own Python signatures.  See DESIGN.md; the compute lives in libposetraj_b200.so (csrc/, C ABI in include/)."""
from .config import SVDConfig  # noqa: F401
from .models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel  # noqa: F401
from .pipeline import StableVideoDiffusionPipelineControlNet  # noqa: F401
from .scheduler import EulerDiscreteScheduler  # noqa: F401
from .vae import AutoencoderKLTemporalDecoder  # noqa: F401
from .clip import CLIPVisionModelWithProjection, resize_with_antialiasing  # noqa: F401
from .train_engine import ControlNetTrainer  # noqa: F401
