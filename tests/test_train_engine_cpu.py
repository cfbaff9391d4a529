"""Host logic of the whole-network reverse pass (posetraj_b200/train_engine.py) that needs no GPU: the gradient
re-layouts are the exact inverses of the WeightStore layouts (a gradient written in the kernel layout lands on the right
parameter element), and the requires-grad sweep prunes the frozen UNet's encoder.  The kernels themselves are checked on
the GPU (tests/test_train_step_gpu.py)."""
import math
from types import SimpleNamespace

import pytest
import torch


def _store(sd):
    from posetraj_b200.engine import WeightStore
    return WeightStore(sd, torch.device("cpu"))


def _roundtrip(kind_key, kernel_tensor, shapes, sd):
    """d/dp sum(kernel_layout(p) * G) = to_param_grads(G): compare with autograd through the same re-layout."""
    from posetraj_b200.train_engine import to_param_grads
    G = torch.randn(kernel_tensor.shape, generator=torch.Generator().manual_seed(5))
    got = to_param_grads(kind_key, G, shapes)
    return G, got


def test_gradient_relayouts_invert_the_weight_store_layouts():
    from posetraj_b200.train_engine import to_param_grads
    g = torch.Generator().manual_seed(0)
    sd = {"c.weight": torch.randn(6, 5, 3, 3, generator=g), "c.bias": torch.randn(6, generator=g),
          "t.weight": torch.randn(4, 4, 3, 1, 1, generator=g), "l.weight": torch.randn(7, 9, generator=g),
          "p.weight": torch.randn(8, 8, 1, 1, generator=g),
          "a.to_q.weight": torch.randn(8, 8, generator=g), "a.to_k.weight": torch.randn(8, 8, generator=g),
          "a.to_v.weight": torch.randn(8, 8, generator=g),
          "x.time_emb_proj.weight": torch.randn(3, 4, generator=g), "y.time_emb_proj.weight": torch.randn(5, 4, generator=g),
          "x.time_emb_proj.bias": torch.randn(3, generator=g), "y.time_emb_proj.bias": torch.randn(5, generator=g),
          "cc.weight": torch.randn(6, 6 + 12, generator=g)}
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    w = _store(sd)
    cases = [
        (w.conv3("c.weight", 64), lambda p: torch.nn.functional.pad(p["c.weight"].permute(0, 2, 3, 1), (0, 64 - 5)).reshape(6, -1)),
        (w.conv3("c.weight"), lambda p: p["c.weight"].permute(0, 2, 3, 1).reshape(6, -1)),
        (w.tconv("t.weight"), lambda p: p["t.weight"].reshape(4, 4, 3).permute(0, 2, 1).reshape(4, 12)),
        (w.linear("l.weight"), lambda p: p["l.weight"]),
        (w.linear("p.weight"), lambda p: p["p.weight"].reshape(8, 8)),
        (w.qkv("a."), lambda p: torch.cat([p["a.to_q.weight"], p["a.to_k.weight"], p["a.to_v.weight"]], 0)),
        (w.cat_rows(["x.time_emb_proj.weight", "y.time_emb_proj.weight"], "bf16"),
         lambda p: torch.cat([p["x.time_emb_proj.weight"], p["y.time_emb_proj.weight"]], 0)),
        (w.cat_rows(["x.time_emb_proj.bias", "y.time_emb_proj.bias"], "f32"),
         lambda p: torch.cat([p["x.time_emb_proj.bias"], p["y.time_emb_proj.bias"]], 0)),
        (w.f32("c.bias"), lambda p: p["c.bias"]),
    ]
    for kt, layout in cases:
        kind_key = w.origin(kt)
        G = torch.randn(kt.shape, generator=g)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        assert torch.allclose(layout(params).detach().to(torch.bfloat16).float(), kt.float(), atol=0, rtol=0) or kt.dtype == torch.float32
        (layout(params) * G).sum().backward()
        got = to_param_grads(kind_key, G, shapes)
        for name, pg in got.items():
            assert tuple(pg.shape) == shapes[name]
            assert torch.equal(pg, params[name].grad), (kind_key, name)
        touched = {k for k, v in params.items() if v.grad is not None}
        assert set(got) == touched, kind_key
    # padded output channels (the conditioning embedding's 16 / 32 / 96-wide layers live in 64-column buffers): extra rows dropped
    kt = w.conv3("c.weight", 64)
    G = torch.randn(64, kt.shape[1], generator=g)
    pg = to_param_grads(w.origin(kt), G, shapes)["c.weight"]
    assert torch.equal(pg, G[:6].view(6, 3, 3, 64)[..., :5].permute(0, 3, 1, 2))
    # cc_projection: two column blocks of ONE parameter
    feat, cam = w.cc_split("cc.weight", 6)
    assert torch.equal(feat.float(), sd["cc.weight"][:, :6].to(torch.bfloat16).float()) and cam.shape == (6, 12)
    gf, gc = torch.randn(6, 6, generator=g), torch.randn(6, 12, generator=g)
    (pf, sf), = to_param_grads(w.origin(feat), gf, shapes).values()
    (pc, sc), = to_param_grads(w.origin(cam), gc, shapes).values()
    full = torch.zeros(6, 18)
    full[sf] = pf
    full[sc] = pc
    assert torch.equal(full, torch.cat([gf, gc], 1))
    # mix factors arrive as d mix_factor already
    assert to_param_grads(("mix", "m.mix_factor"), torch.tensor([0.25]), {"m.mix_factor": (1,)})["m.mix_factor"].tolist() == [0.25]
    with pytest.raises(KeyError):
        to_param_grads(("nope", "c.weight"), G, shapes)


def test_weight_store_refresh_follows_the_state_dict_in_place():
    sd = {"l.weight": torch.randn(4, 8), "l.bias": torch.randn(4)}
    w = _store(sd)
    kt, kb = w.linear("l.weight"), w.f32("l.bias")
    p0 = kt.data_ptr()
    sd["l.weight"].mul_(2.0)
    sd["l.bias"].add_(1.0)
    w.refresh()
    assert kt.data_ptr() == p0                                    # pointers (and TMA descriptors built on them) stay valid
    assert torch.equal(kt.float(), sd["l.weight"].to(torch.bfloat16).float()) and torch.equal(kb, sd["l.bias"])


def test_requires_grad_sweep_prunes_what_does_not_depend_on_the_seeds():
    """A frozen plan is only differentiated downstream of the injected residuals: `out` of a GEMM with a second output
    (out2 = out + m * aux) does not depend on aux, so nothing upstream of it is marked."""
    from posetraj_b200 import ops
    from posetraj_b200.train_engine import Tape, _key
    t = lambda: torch.zeros(4, 8, dtype=torch.bfloat16)
    x, w_, h, skip, aux, y, z = t(), t(), t(), t(), t(), t(), t()

    def fake(cls, **io):
        op = cls.__new__(cls)
        base = dict(a0=None, a1=None, w=w_, out=None, bias=None, rowvec=None, res1=None, res2=None, out2=None, aux=None)
        base.update(io)
        op.io = SimpleNamespace(**base)
        return op

    enc = fake(ops.Gemm, a0=x, out=h, out2=skip, aux=aux)        # encoder GEMM: h continues down, skip = h + m * aux
    down = fake(ops.Gemm, a0=h, out=y)                            # rest of the encoder: independent of aux
    dec = fake(ops.Gemm, a0=y, a1=skip, out=z)                    # decoder consumes the skip
    plan = SimpleNamespace(w=None)
    tape = Tape(plan, trainable=False)
    tape.mark([[enc, down, dec]], seeds=[aux])
    assert tape.needs(skip) and tape.needs(z)
    assert not tape.needs(h) and not tape.needs(y) and not tape.needs(x)
