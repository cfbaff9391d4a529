"""The GEGLU gate of the tensor-core epilogues (posetraj_b200/csrc/common.cuh `geglu_gate_fast`) restated in numpy:
value * g * sigmoid(2 g (c0 + c1 s + c2 s^2)), s = min(g^2, 81), against the EXACT erf GELU diffusers' GEGLU uses
(F.gelu(gate), approximate="none").  Bound: 3e-5 absolute over the whole real line — two orders
of magnitude below the bf16 rounding (2^-9 relative) of the gated value, so it does not show in any parity number."""
import math
import re
from pathlib import Path

import numpy as np


def _constants():
    src = (Path(__file__).resolve().parents[1] / "posetraj_b200" / "csrc" / "common.cuh").read_text()
    body = src[src.index("PT_DEVICE float geglu_gate_fast"):]
    body = body[: body.index("}")]
    c = [float(x) for x in re.findall(r"(-?\d\.\d+e[+-]\d+)f \* -2\.885390081777927f", body)]
    assert len(c) == 3, c
    return c[2], c[1], c[0]      # c0, c1, c2 in the order of the formula


def gate_np(g):
    c0, c1, c2 = _constants()
    g = g.astype(np.float32)
    s = np.minimum(g * g, np.float32(81.0))
    k = np.float32(-2.885390081777927)
    poly = (np.float32(c2) * k) * s + np.float32(c1) * k
    poly = poly * s + np.float32(c0) * k
    with np.errstate(over="ignore"):
        e = np.exp2((poly * g).astype(np.float32))
    return (g / (np.float32(1.0) + e)).astype(np.float32)


def test_gate_matches_exact_erf_gelu():
    g = np.concatenate([np.linspace(-30, 30, 600001), np.linspace(-1e3, 1e3, 2001), [-1e30, 1e30, 0.0]])
    exact = np.array([x * 0.5 * (1.0 + math.erf(x / math.sqrt(2.0))) for x in g])
    got = gate_np(g).astype(np.float64)
    err = np.abs(got - exact)
    finite = np.abs(g) < 1e20
    assert np.isfinite(got).all()
    assert err[finite].max() < 3e-5, (err[finite].max(), g[finite][err[finite].argmax()])
    assert got[-3] == 0.0 and got[-2] == np.float32(1e30)      # saturates cleanly at both ends


def test_gate_is_much_closer_than_the_textbook_tanh_form():
    g = np.linspace(-6, 6, 120001)
    exact = np.array([x * 0.5 * (1.0 + math.erf(x / math.sqrt(2.0))) for x in g])
    tanh_form = 0.5 * g * (1 + np.tanh(math.sqrt(2 / math.pi) * (g + 0.044715 * g ** 3)))
    assert np.abs(gate_np(g) - exact).max() * 10 < np.abs(tanh_form - exact).max()


def test_tanh_form_is_the_same_function():
    """`geglu_gate_tanh` (one MUFU: hv + hv tanh(u)) is `geglu_gate_fast` rewritten with sigmoid(2u) = (1 + tanh(u)) / 2: the
    same three constants, and with an exact tanh the same 3e-5 bound against the erf GELU; what the hardware tanh.approx
    (2^-11 relative) adds is bounded on the GPU in tests/test_gemm_gpu.py."""
    src = (Path(__file__).resolve().parents[1] / "posetraj_b200" / "csrc" / "common.cuh").read_text()
    body = src[src.index("PT_DEVICE float geglu_gate_tanh"):]
    body = body[: body.index("}")]
    c = [float(x) for x in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body)]
    assert (c[2], c[1], c[0]) == _constants()
    c0, c1, c2 = _constants()
    g = np.concatenate([np.linspace(-30, 30, 600001), np.linspace(-1e3, 1e3, 2001)]).astype(np.float32)
    s = np.minimum(g * g, np.float32(81.0))
    poly = (np.float32(c2) * s + np.float32(c1)).astype(np.float32)
    poly = (poly * s + np.float32(c0)).astype(np.float32)
    u = (poly * g).astype(np.float32)
    hv = (np.float32(0.5) * g).astype(np.float32)
    got = (hv * np.tanh(u.astype(np.float64)).astype(np.float32) + hv).astype(np.float64)
    exact = np.array([x * 0.5 * (1.0 + math.erf(x / math.sqrt(2.0))) for x in g.astype(np.float64)])
    assert np.abs(got - exact).max() < 3e-5
