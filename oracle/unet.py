"""Oracle restatement of the Flow-Matching velocity U-Net (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/pnpflow/models.py:
  Swish                                  :24-30
  group_norm (32 groups, eps 1e-6)       :33-38
  upsample / downsample                  :41-55
  ResidualBlock.forward                  :94-113
  SelfAttention.forward                  :145-162
  variance_scaling_init_ (fan quirk)     :165-216
  get_sinusoidal_positional_embedding    :253-279
  TimestepEmbedding.forward              :296-299
  UNet.__init__ (key scheme) / forward   :302-440 / :442-495
and the hyper-parameters of pnpflow/utils.py:172-179 (CelebA/AFHQ nets) and
demo/dirichlet/Diri_PnP.ipynb:33-40 (MNIST net).

It is written functionally over a flat ``state_dict`` (the reference's checkpoint
format, pnpflow/utils.py:225) instead of as nn.Modules, so that the same walk
(``unet_layer_spec``) drives the oracle forward, the synthetic-weight recipe and
the tests of the engine's weight repacker.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    input_channels: int
    input_height: int
    ch: int = 32
    ch_mult: Tuple[int, ...] = (1, 2, 4, 8)
    num_res_blocks: int = 6
    attn_resolutions: Tuple[int, ...] = (16, 8)

    @property
    def temb_ch(self) -> int:
        return self.ch * 4


CELEBA_128 = UNetConfig(3, 128)                                   # pnpflow/utils.py:172-179, dim_image 128
AFHQ_256 = UNetConfig(3, 256)                                     # same factory, dim_image 256
MNIST_28 = UNetConfig(1, 28, 32, (1, 2), 2, (16,))                # demo/dirichlet/Diri_PnP.ipynb:33-40


# ----------------------------------------------------------------------------------------------
# Layer walk: one entry per module in *execution order*, with the reference state_dict prefix.
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Layer:
    kind: str          # 'conv' | 'res' | 'attn' | 'down' | 'up' | 'end'
    prefix: str        # state_dict key prefix
    in_ch: int
    out_ch: int
    side: int          # spatial side of the *input* of this layer
    skip_ch: int = 0   # channels popped from the skip stack and concatenated (up path)
    push: bool = False  # output is pushed on the skip stack


def unet_layer_spec(cfg: UNetConfig) -> List[Layer]:
    """Execution-ordered layer list (models.py:360-436 for the naming, :442-495 for the order)."""
    L: List[Layer] = []
    nres = len(cfg.ch_mult)
    side = cfg.input_height
    assert side % 2 ** (nres - 1) == 0, "input_height doesn't satisfy the condition"
    L.append(Layer('conv', 'begin_conv', cfg.input_channels, cfg.ch, side, push=True))
    skip = [cfg.ch]
    in_ch = cfg.ch
    for lvl in range(nres):
        out_ch = cfg.ch * cfg.ch_mult[lvl]
        for blk in range(cfg.num_res_blocks):
            has_attn = side in cfg.attn_resolutions
            L.append(Layer('res', f'down_modules.{lvl}.{lvl}a_{blk}a_block', in_ch, out_ch, side,
                           push=not has_attn))
            if has_attn:
                L.append(Layer('attn', f'down_modules.{lvl}.{lvl}a_{blk}b_attn', out_ch, out_ch, side, push=True))
            skip.append(out_ch)
            in_ch = out_ch
        if lvl != nres - 1:
            L.append(Layer('down', f'down_modules.{lvl}.{lvl}b_downsample', in_ch, in_ch, side, push=True))
            side //= 2
            skip.append(in_ch)
    L.append(Layer('res', 'mid_modules.0', in_ch, in_ch, side))
    L.append(Layer('attn', 'mid_modules.1', in_ch, in_ch, side))
    L.append(Layer('res', 'mid_modules.2', in_ch, in_ch, side))
    for idx, lvl in enumerate(reversed(range(nres))):
        out_ch = cfg.ch * cfg.ch_mult[lvl]
        for blk in range(cfg.num_res_blocks + 1):
            sc = skip.pop()
            L.append(Layer('res', f'up_modules.{idx}.{lvl}a_{blk}a_block', in_ch + sc, out_ch, side, skip_ch=sc))
            if side in cfg.attn_resolutions:
                L.append(Layer('attn', f'up_modules.{idx}.{lvl}a_{blk}b_attn', out_ch, out_ch, side))
            in_ch = out_ch
        if lvl != 0:
            L.append(Layer('up', f'up_modules.{idx}.{lvl}b_upsample.up_conv', in_ch, in_ch, side))
            side *= 2
    assert not skip
    L.append(Layer('end', 'end_conv', in_ch, cfg.input_channels, side))
    return L


# ----------------------------------------------------------------------------------------------
# Functional forward
# ----------------------------------------------------------------------------------------------
def swish(x: torch.Tensor) -> torch.Tensor:
    return torch.sigmoid(x) * x                       # models.py:29-30 (sigmoid first)


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], eps=1e-6)   # models.py:33-38


def _conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """models.py:253-279.  t is used raw (in [0,1)), layout [sin(half) | cos(half)]."""
    assert t.dim() == 1
    t = t.to(torch.get_default_dtype())
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float, device=t.device) * -k)
    ang = t[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1), "constant", 0)
    return emb


def time_embedding(sd, t: torch.Tensor, cfg: UNetConfig) -> torch.Tensor:
    """models.py:296-299: Linear(ch->4ch) -> Swish -> Linear(4ch->4ch)."""
    e = sinusoidal_embedding(t, cfg.ch)
    e = F.linear(e, sd['temb_net.main.0.weight'], sd['temb_net.main.0.bias'])
    e = swish(e)
    return F.linear(e, sd['temb_net.main.2.weight'], sd['temb_net.main.2.bias'])


def res_block(sd, p: str, x: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
    """models.py:94-113."""
    h = _conv(swish(_gn(x, sd, p + '.norm1')), sd, p + '.conv1')
    h = h + F.linear(swish(temb), sd[p + '.temb_proj.weight'], sd[p + '.temb_proj.bias'])[:, :, None, None]
    h = _conv(swish(_gn(h, sd, p + '.norm2')), sd, p + '.conv2')
    if (p + '.shortcut.weight') in sd:                 # 1x1 conv iff in_ch != out_ch (models.py:85-92)
        x = _conv(x, sd, p + '.shortcut', padding=0)
    assert x.shape == h.shape
    return x + h


def self_attention(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """models.py:145-162 (single head, d = C, scale C^-1/2, softmax over keys)."""
    _, C, H, W = x.shape
    h = _gn(x, sd, p + '.norm')
    q = _conv(h, sd, p + '.attn_q', padding=0).view(-1, C, H * W)
    k = _conv(h, sd, p + '.attn_k', padding=0).view(-1, C, H * W)
    v = _conv(h, sd, p + '.attn_v', padding=0).view(-1, C, H * W)
    attn = torch.bmm(q.permute(0, 2, 1), k) * (int(C) ** (-0.5))
    attn = torch.softmax(attn, dim=-1)
    h = torch.bmm(v, attn.permute(0, 2, 1)).view(-1, C, H, W)
    h = _conv(h, sd, p + '.proj_out', padding=0)
    return x + h


def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, x: torch.Tensor, t: torch.Tensor,
                 tap: Optional[Callable[[int, Layer, torch.Tensor], None]] = None) -> torch.Tensor:
    """v_theta(x, t)  (models.py:442-495).  ``tap(i, layer, out)`` observes every layer output."""
    B = x.shape[0]
    temb = time_embedding(sd, t, cfg)
    assert list(temb.shape) == [B, cfg.temb_ch]
    hs: List[torch.Tensor] = []
    h = x
    for i, L in enumerate(unet_layer_spec(cfg)):
        if L.kind == 'conv':
            h = _conv(h, sd, L.prefix)
        elif L.kind == 'res':
            inp = torch.cat([h, hs.pop()], dim=1) if L.skip_ch else h      # down path: h is hs[-1] (models.py:459)
            h = res_block(sd, L.prefix, inp, temb)
        elif L.kind == 'attn':
            h = self_attention(sd, L.prefix, h)
        elif L.kind == 'down':
            h = _conv(h, sd, L.prefix, stride=2)
        elif L.kind == 'up':
            h = _conv(F.interpolate(h, scale_factor=2, mode='nearest'), sd, L.prefix)
        elif L.kind == 'end':
            h = _conv(swish(_gn(h, sd, L.prefix + '.0')), sd, L.prefix + '.2')
        if L.push:
            hs.append(h)
        if tap is not None:
            tap(i, L, h)
    assert not hs
    assert list(h.shape) == [B, cfg.input_channels, x.shape[2], x.shape[3]]
    return h


# ----------------------------------------------------------------------------------------------
# Synthetic weights (SURVEY.md §8d / BASELINE.md §5 recipe, restated with an explicit generator)
# ----------------------------------------------------------------------------------------------
def _vs_uniform(shape: Sequence[int], gain: float, gen: torch.Generator) -> torch.Tensor:
    """variance_scaling_init_(mode='fan_avg') as the reference really behaves: models.py:165-176
    returns fan_out for any mode other than 'fan_in', so var = gain / fan_out; uniform(+-sqrt(3 var))."""
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_out = shape[0] * rf
    bound = math.sqrt(3.0 * gain / max(1.0, fan_out))
    return (torch.rand(tuple(shape), generator=gen, dtype=torch.float32) * 2 - 1) * bound


def init_state_dict(cfg: UNetConfig, seed: int = 0, inner_gain: float = 1.0, end_gain: float = 1e-3,
                    perturb: float = 0.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights with the reference's key scheme.

    Layers the reference creates with init_scale=0 (every ResBlock conv2, every attention proj_out,
    end_conv.2; models.py:84,131-137,432) would make v ~ 1e-4 and parity vacuous, so they are drawn
    with variance gain ``inner_gain`` (conv2/proj_out) and ``end_gain`` (end_conv.2): the recipe
    SURVEY Appendix C shows to be stable and precision-insensitive.  ``perturb`` > 0 additionally
    randomises biases and GroupNorm affine parameters (used by unit tests so that those code paths
    are exercised; the reference initialises them to 0 / 1).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bias(n):
        if perturb:
            return (torch.rand(n, generator=g) * 2 - 1) * perturb
        return torch.zeros(n)

    def conv(p, cin, cout, k, gain=1.0):
        sd[p + '.weight'] = _vs_uniform((cout, cin, k, k), gain, g)
        sd[p + '.bias'] = bias(cout)

    def dense(p, cin, cout):
        sd[p + '.weight'] = _vs_uniform((cout, cin), 1.0, g)
        sd[p + '.bias'] = bias(cout)

    def gn(p, c):
        sd[p + '.weight'] = torch.ones(c) + ((torch.rand(c, generator=g) * 2 - 1) * perturb if perturb else 0)
        sd[p + '.bias'] = bias(c)

    dense('temb_net.main.0', cfg.ch, cfg.temb_ch)
    dense('temb_net.main.2', cfg.temb_ch, cfg.temb_ch)
    for L in unet_layer_spec(cfg):
        p = L.prefix
        if L.kind == 'conv' or L.kind == 'down' or L.kind == 'up':
            conv(p, L.in_ch, L.out_ch, 3)
        elif L.kind == 'res':
            dense(p + '.temb_proj', cfg.temb_ch, L.out_ch)
            gn(p + '.norm1', L.in_ch)
            conv(p + '.conv1', L.in_ch, L.out_ch, 3)
            gn(p + '.norm2', L.out_ch)
            conv(p + '.conv2', L.out_ch, L.out_ch, 3, inner_gain)
            if L.in_ch != L.out_ch:
                conv(p + '.shortcut', L.in_ch, L.out_ch, 1)
        elif L.kind == 'attn':
            for n in ('attn_q', 'attn_k', 'attn_v'):
                conv(p + '.' + n, L.in_ch, L.in_ch, 1)
            conv(p + '.proj_out', L.in_ch, L.in_ch, 1, inner_gain)
            gn(p + '.norm', L.in_ch)
        elif L.kind == 'end':
            gn(p + '.0', L.in_ch)
            conv(p + '.2', L.in_ch, L.out_ch, 3, end_gain)
    return sd
