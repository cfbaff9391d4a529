"""ORACLE (test infrastructure, never the product path) — CPU restatement of the image-conditioning branch of the
pipeline:

  `_encode_image`                /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:145-172
  `_resize_with_antialiasing`    /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:602-632
  `_compute_padding`/`_filter2d`/`_gaussian`/`_gaussian_blur2d`                                   :635-712
  the CLIP vision tower          `transformers.CLIPVisionModelWithProjection` (third-party; the SVD checkpoint's
                                 image_encoder is ViT-H/14: 32 layers, width 1280, 16 heads of 80, MLP 5120, GELU,
                                 projection 1024)

Pinning: the resize functions are checked against tests/golden/resize_golden.pt, produced by EXECUTING the reference's
own function sources (tests/golden/gen_resize_golden.py); the vision tower is checked against the installed
`transformers` implementation on identical weights (tests/test_clip_cpu.py).  Parameter names equal the HF state-dict
key tree, so the SVD `image_encoder` checkpoint loads unchanged.

Only tests/ and __graft_entry__.smoke() may import this.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# anti-aliased resize (reference :602-712)
# ----------------------------------------------------------------------------------------------
def blur_params(h: int, w: int, size):
    """Kernel sizes and sigmas of the Gaussian pre-filter (:608-628)."""
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = [int(max(2.0 * 2 * sigmas[0], 3)), int(max(2.0 * 2 * sigmas[1], 3))]
    ks = [k + 1 if k % 2 == 0 else k for k in ks]
    return tuple(ks), sigmas


def gaussian_window(window_size: int, sigma: float) -> torch.Tensor:
    """`_gaussian` (:686-697) for one sigma, fp32."""
    s = torch.tensor([[sigma]], dtype=torch.float32)
    x = (torch.arange(window_size, dtype=torch.float32) - window_size // 2).expand(1, -1)
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * s.pow(2.0)))
    return (g / g.sum(-1, keepdim=True))[0]


def resize_with_antialiasing(x: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """Separable Gaussian blur with reflect padding (x pass, then y pass), then bicubic, align_corners=True."""
    if x.ndim == 3:
        x = x.unsqueeze(0)
    b, c, h, w = x.shape
    (ky, kx), (sy, sx) = blur_params(h, w, size)
    wx, wy = gaussian_window(kx, sx), gaussian_window(ky, sy)
    px, py = (kx - 1) // 2, (ky - 1) // 2
    xp = F.pad(x, (px, kx - 1 - px, 0, 0), mode="reflect")
    xb = F.conv2d(xp.reshape(b * c, 1, h, -1), wx.view(1, 1, 1, kx)).reshape(b, c, h, w)
    yp = F.pad(xb, (0, 0, py, ky - 1 - py), mode="reflect")
    yb = F.conv2d(yp.reshape(b * c, 1, -1, w), wy.view(1, 1, ky, 1)).reshape(b, c, h, w)
    return F.interpolate(yb, size=size, mode="bicubic", align_corners=True)


# ----------------------------------------------------------------------------------------------
# CLIP vision tower with projection (HF naming)
# ----------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.q_proj, self.k_proj, self.v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, s, d = x.shape
        sh = lambda t: t.view(b, s, self.heads, d // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sh(self.q_proj(x)), sh(self.k_proj(x)), sh(self.v_proj(x)))
        return self.out_proj(o.transpose(1, 2).reshape(b, s, d))


class _Mlp(nn.Module):
    def __init__(self, dim, inner, act):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(dim, inner), nn.Linear(inner, dim), act

    def forward(self, x):
        h = self.fc1(x)
        h = F.gelu(h) if self.act == "gelu" else h * torch.sigmoid(1.702 * h)
        return self.fc2(h)


class _Layer(nn.Module):
    def __init__(self, dim, heads, inner, act, eps):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(dim, eps=eps)
        self.self_attn = _Attn(dim, heads)
        self.layer_norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, inner, act)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, dim, image_size, patch):
        super().__init__()
        n = (image_size // patch) ** 2
        self.class_embedding = nn.Parameter(torch.randn(dim))
        self.patch_embedding = nn.Conv2d(3, dim, patch, stride=patch, bias=False)
        self.position_embedding = nn.Embedding(n + 1, dim)

    def forward(self, pixel_values):
        p = self.patch_embedding(pixel_values).flatten(2).transpose(1, 2)
        cls = self.class_embedding.expand(p.shape[0], 1, -1)
        return torch.cat([cls, p], 1) + self.position_embedding.weight[None]


class _Encoder(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class _VisionModel(nn.Module):
    def __init__(self, dim, heads, inner, layers, image_size, patch, act, eps):
        super().__init__()
        self.embeddings = _Embeddings(dim, image_size, patch)
        self.pre_layrnorm = nn.LayerNorm(dim, eps=eps)   # (sic) the HF attribute name
        self.encoder = _Encoder([_Layer(dim, heads, inner, act, eps) for _ in range(layers)])
        self.post_layernorm = nn.LayerNorm(dim, eps=eps)

    def forward(self, pixel_values):
        x = self.pre_layrnorm(self.embeddings(pixel_values))
        for layer in self.encoder.layers:
            x = layer(x)
        return self.post_layernorm(x[:, 0])


class CLIPVisionModelWithProjection(nn.Module):
    def __init__(self, hidden_size=1280, num_attention_heads=16, intermediate_size=5120, num_hidden_layers=32,
                 image_size=224, patch_size=14, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5):
        super().__init__()
        self.vision_model = _VisionModel(hidden_size, num_attention_heads, intermediate_size, num_hidden_layers,
                                         image_size, patch_size, hidden_act, layer_norm_eps)
        self.visual_projection = nn.Linear(hidden_size, projection_dim, bias=False)

    def forward(self, pixel_values):
        return self.visual_projection(self.vision_model(pixel_values))


def encode_image(model: CLIPVisionModelWithProjection, image01: torch.Tensor) -> torch.Tensor:
    """`_encode_image` (:145-172) up to the CFG duplication: image in [0, 1] -> anti-aliased 224x224 -> CLIP
    `image_embeds` -> [N, 1, D] (the reference applies neither the [-1,1] mapping nor the CLIP mean/std here)."""
    size = model.vision_model.embeddings.position_embedding.weight.shape[0] - 1
    side = int(round(math.sqrt(size))) * model.vision_model.embeddings.patch_embedding.kernel_size[0]
    x = resize_with_antialiasing(image01, (side, side))
    return model(x).unsqueeze(1)
