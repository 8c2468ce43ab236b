"""Synthetic weights and inputs for benchmarks (BASELINE.json: random-init U-Net of the named width, synthetic y).

The recipe is the one SURVEY.md §8d / BASELINE.md §5 state: variance-scaling uniform init with the reference's
fan quirk (var = gain / fan_out, pnpflow/models.py:165-216), zero biases, unit GroupNorm scale, and the layers the
reference creates with init_scale=0 re-drawn with gain 1 (ResBlock conv2, attention proj_out) and 1e-3
(end_conv.2) so that the loop is well-posed.  Throughput does not depend on the values.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict

import torch

from . import _lib

NETS = {
    "celeba128": dict(input_channels=3, input_height=128, ch=32, ch_mult=(1, 2, 4, 8), num_res_blocks=6, attn_resolutions=(16, 8)),
    "afhq256": dict(input_channels=3, input_height=256, ch=32, ch_mult=(1, 2, 4, 8), num_res_blocks=6, attn_resolutions=(16, 8)),
}


def expected_shapes(cfg: Dict) -> Dict[str, tuple]:
    """state_dict key -> shape, asked from the engine's own plan (no GPU needed)."""
    lib = _lib.load()
    c = _lib.UNetConfigC()
    c.input_channels, c.input_height, c.ch, c.num_levels = cfg["input_channels"], cfg["input_height"], cfg["ch"], len(cfg["ch_mult"])
    for i, m in enumerate(cfg["ch_mult"]):
        c.ch_mult[i] = m
    c.num_res_blocks, c.num_attn_resolutions = cfg["num_res_blocks"], len(cfg["attn_resolutions"])
    for i, m in enumerate(cfg["attn_resolutions"]):
        c.attn_resolutions[i] = m
    h = C.c_void_p()
    _lib.check(lib.pnpf_create(C.byref(c), C.byref(h)))
    out = {}
    for i in range(lib.pnpf_num_weights(h)):
        shp, nd = (C.c_int64 * 4)(), C.c_int()
        _lib.check(lib.pnpf_weight_shape(h, i, C.byref(shp), C.byref(nd)))
        out[lib.pnpf_weight_name(h, i).decode()] = tuple(shp[k] for k in range(nd.value))
    lib.pnpf_destroy(h)
    return out


def random_state_dict(cfg: Dict, seed: int = 0, inner_gain: float = 1.0, end_gain: float = 1e-3) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in expected_shapes(cfg).items():
        if len(shape) == 1:
            is_gn_scale = name.endswith(".weight")
            sd[name] = torch.ones(shape) if is_gn_scale else torch.zeros(shape)
            continue
        gain = 1.0
        if name.endswith("conv2.weight") or name.endswith("proj_out.weight"):
            gain = inner_gain
        elif name == "end_conv.2.weight":
            gain = end_gain
        fan_out = shape[0] * (shape[2] * shape[3] if len(shape) == 4 else 1)
        bound = math.sqrt(3.0 * gain / max(1.0, fan_out))
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def synthetic_clean(batch: int, channels: int, side: int, seed: int) -> torch.Tensor:
    """clean = 2*U[0,1) - 1 (data range of the reference's Normalize(0.5, 0.5), dataloaders.py:30,83), box-filtered
    to look image-like."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, channels, side, side, generator=g) * 2 - 1
    k = 9
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (k // 2,) * 4, mode="reflect"), k, stride=1)
    return (x / x.abs().amax(dim=(1, 2, 3), keepdim=True)).contiguous()
