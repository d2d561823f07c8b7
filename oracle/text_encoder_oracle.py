"""CPU oracle of the step BEFORE the hot path: the reference TextEncoder + duration predictor
(`/root/reference/model.py:148-535`), restated as plain fp32 torch functions over a flat state-dict.

TEST INFRASTRUCTURE ONLY -- like cfm_oracle.py it may be imported by `tests/` (and a future bench leg), never by the
package.  No native implementation of this component exists yet: this file and `tests/golden/text_golden.npz` (outputs of
the live reference, `tests/golden/make_text_golden.py`) are step (a) -- oracle and boundary -- of the next scope row.

Boundary the native version has to keep (model.py:500-535): `encoder(x, x_lengths, spks=None) -> (mu, logw, x_mask)` with
x (B, T_x) int64 token ids, x_lengths (B,), spks (B, spk_emb_dim) already embedded; mu (B, n_feats, T_x), logw (B, 1, T_x),
x_mask (B, 1, T_x) float.  State-dict keys are the reference's (`encoder.*` of MatchaTTS): see `param_shapes`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class TextEncCfg:
    """Architecture literals of main.py:63-79 (LJSpeech single speaker)."""
    n_vocab: int = 178
    n_feats: int = 80
    n_channels: int = 192
    filter_channels: int = 768
    n_heads: int = 2
    n_layers: int = 6
    kernel_size: int = 3
    prenet: bool = True
    filter_channels_dp: int = 256
    kernel_size_dp: int = 3
    n_spks: int = 1
    spk_emb_dim: int = 64

    @property
    def width(self) -> int:            # encoder width: the speaker embedding is concatenated after the prenet (model.py:523-524)
        return self.n_channels + (self.spk_emb_dim if self.n_spks > 1 else 0)


def param_shapes(cfg: TextEncCfg) -> Dict[str, tuple]:
    """Key -> shape of the reference TextEncoder's state-dict (module tree of model.py:441-498)."""
    C, W, Fc, k = cfg.n_channels, cfg.width, cfg.filter_channels, cfg.kernel_size
    s: Dict[str, tuple] = {"emb.weight": (cfg.n_vocab, C)}
    if cfg.prenet:                     # ConvReluNorm(C, C, C, kernel_size=5, n_layers=3), model.py:463-471
        for i in range(3):
            s[f"prenet.conv_layers.{i}.weight"] = (C, C, 5)
            s[f"prenet.conv_layers.{i}.bias"] = (C,)
            s[f"prenet.norm_layers.{i}.gamma"] = (C,)
            s[f"prenet.norm_layers.{i}.beta"] = (C,)
        s["prenet.proj.weight"] = (C, C, 1)
        s["prenet.proj.bias"] = (C,)
    for i in range(cfg.n_layers):      # Encoder, model.py:412-425
        for n in ("q", "k", "v", "o"):
            s[f"encoder.attn_layers.{i}.conv_{n}.weight"] = (W, W, 1)
            s[f"encoder.attn_layers.{i}.conv_{n}.bias"] = (W,)
        s[f"encoder.norm_layers_1.{i}.gamma"] = (W,)
        s[f"encoder.norm_layers_1.{i}.beta"] = (W,)
        s[f"encoder.ffn_layers.{i}.conv_1.weight"] = (Fc, W, k)
        s[f"encoder.ffn_layers.{i}.conv_1.bias"] = (Fc,)
        s[f"encoder.ffn_layers.{i}.conv_2.weight"] = (W, Fc, k)
        s[f"encoder.ffn_layers.{i}.conv_2.bias"] = (W,)
        s[f"encoder.norm_layers_2.{i}.gamma"] = (W,)
        s[f"encoder.norm_layers_2.{i}.beta"] = (W,)
    s["proj_m.weight"] = (cfg.n_feats, W, 1)
    s["proj_m.bias"] = (cfg.n_feats,)
    D, kd = cfg.filter_channels_dp, cfg.kernel_size_dp
    s.update({"proj_w.conv_1.weight": (D, W, kd), "proj_w.conv_1.bias": (D,), "proj_w.norm_1.gamma": (D,), "proj_w.norm_1.beta": (D,),
              "proj_w.conv_2.weight": (D, D, kd), "proj_w.conv_2.bias": (D,), "proj_w.norm_2.gamma": (D,), "proj_w.norm_2.beta": (D,),
              "proj_w.proj.weight": (1, D, 1), "proj_w.proj.bias": (1,)})
    return s


def make_state_dict(cfg: TextEncCfg, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded weights with every affine term away from its identity value (the reference zero-initialises prenet.proj,
    model.py:197-198, which would hide that branch).  Deterministic for a given torch CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape in param_shapes(cfg).items():
        if key.endswith("gamma"):
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("beta") or key.endswith("bias"):
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        elif key == "emb.weight":
            sd[key] = torch.randn(shape, generator=g) * cfg.n_channels ** -0.5       # model.py:459
        else:
            fan_in = shape[1] * (shape[2] if len(shape) > 2 else 1)
            sd[key] = torch.randn(shape, generator=g) * fan_in ** -0.5
    return sd


def sequence_mask(length: Tensor, max_length: int) -> Tensor:
    return torch.arange(max_length, device=length.device)[None, :] < length[:, None]      # model.py:42-46


def channel_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-4) -> Tensor:
    """model.py:148-166 -- normalisation over the CHANNEL axis of (B, C, T), biased variance, eps 1e-4."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma[None, :, None] + beta[None, :, None]


def prenet(sd, x: Tensor, m: Tensor) -> Tensor:
    """model.py:200-207 -- 3 x [conv k5 on x*m -> channel norm -> ReLU], residual 1x1 projection, mask."""
    h = x
    for i in range(3):
        h = F.conv1d(h * m, sd[f"prenet.conv_layers.{i}.weight"], sd[f"prenet.conv_layers.{i}.bias"], padding=2)
        h = torch.relu(channel_norm(h, sd[f"prenet.norm_layers.{i}.gamma"], sd[f"prenet.norm_layers.{i}.beta"]))
    return (x + F.conv1d(h, sd["prenet.proj.weight"], sd["prenet.proj.bias"])) * m


def rope(x: Tensor, d: int, base: float = 10_000.0) -> Tensor:
    """model.py:244-289 -- rotary embedding on the first d features of (B, H, T, c); the rest passes through.
    angle[t, j] = t * base^(-2 (j mod d/2) / d); out = x cos + rot(x) sin with rot(x) = [-x[d/2:], x[:d/2]]."""
    T = x.shape[2]
    theta = 1.0 / (base ** (torch.arange(0, d, 2).float() / d))
    ang = torch.arange(T).float()[:, None] * theta[None, :]
    ang = torch.cat([ang, ang], dim=1)                                    # (T, d)
    xr, xp = x[..., :d], x[..., d:]
    rot = torch.cat([-xr[..., d // 2:], xr[..., :d // 2]], dim=-1)
    return torch.cat([xr * ang.cos() + rot * ang.sin(), xp], dim=-1)


def attention(sd, pfx: str, x: Tensor, amask: Tensor, cfg: TextEncCfg) -> Tensor:
    """model.py:335-365 -- 1x1-conv q/k/v (with bias), RoPE on half of each head, scores / sqrt(c), masked_fill(-1e4),
    softmax, 1x1-conv output."""
    B, W, T = x.shape
    H = cfg.n_heads
    c = W // H
    q, k, v = (F.conv1d(x, sd[f"{pfx}.conv_{n}.weight"], sd[f"{pfx}.conv_{n}.bias"]).reshape(B, H, c, T).transpose(2, 3)
               for n in ("q", "k", "v"))                                  # (B, H, T, c)
    d = int(c * 0.5)                                                       # model.py:319-320
    q, k = rope(q, d), rope(k, d)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(c)
    s = s.masked_fill(amask == 0, -1e4)
    o = torch.matmul(torch.softmax(s, dim=-1), v)                          # (B, H, T, c)
    o = o.transpose(2, 3).reshape(B, W, T)
    return F.conv1d(o, sd[f"{pfx}.conv_o.weight"], sd[f"{pfx}.conv_o.bias"])


def encoder(sd, x: Tensor, m: Tensor, cfg: TextEncCfg, trace: Optional[dict] = None) -> Tensor:
    """model.py:427-438 -- post-norm transformer: x = LN1(x + MHA(x)); x = LN2(x + FFN(x)); conv FFN of kernel k."""
    amask = m.unsqueeze(2) * m.unsqueeze(-1)                               # (B, 1, T, T)
    pad = cfg.kernel_size // 2
    for i in range(cfg.n_layers):
        x = x * m
        y = attention(sd, f"encoder.attn_layers.{i}", x, amask, cfg)
        x = channel_norm(x + y, sd[f"encoder.norm_layers_1.{i}.gamma"], sd[f"encoder.norm_layers_1.{i}.beta"])
        h = torch.relu(F.conv1d(x * m, sd[f"encoder.ffn_layers.{i}.conv_1.weight"], sd[f"encoder.ffn_layers.{i}.conv_1.bias"], padding=pad))
        y = F.conv1d(h * m, sd[f"encoder.ffn_layers.{i}.conv_2.weight"], sd[f"encoder.ffn_layers.{i}.conv_2.bias"], padding=pad) * m
        x = channel_norm(x + y, sd[f"encoder.norm_layers_2.{i}.gamma"], sd[f"encoder.norm_layers_2.{i}.beta"])
        if trace is not None:
            trace[f"layer{i}"] = x.clone()
    return x * m


def duration_predictor(sd, x: Tensor, m: Tensor, cfg: TextEncCfg) -> Tensor:
    """model.py:224-235 -- conv -> ReLU -> channel norm (note the order), twice; 1x1 projection; masked."""
    pad = cfg.kernel_size_dp // 2
    h = torch.relu(F.conv1d(x * m, sd["proj_w.conv_1.weight"], sd["proj_w.conv_1.bias"], padding=pad))
    h = channel_norm(h, sd["proj_w.norm_1.gamma"], sd["proj_w.norm_1.beta"])
    h = torch.relu(F.conv1d(h * m, sd["proj_w.conv_2.weight"], sd["proj_w.conv_2.bias"], padding=pad))
    h = channel_norm(h, sd["proj_w.norm_2.gamma"], sd["proj_w.norm_2.beta"])
    return F.conv1d(h * m, sd["proj_w.proj.weight"], sd["proj_w.proj.bias"]) * m


def text_encoder_forward(sd, cfg: TextEncCfg, x: Tensor, x_lengths: Tensor, spks: Optional[Tensor] = None,
                         trace: Optional[dict] = None):
    """model.py:500-535 -> (mu, logw, x_mask)."""
    h = F.embedding(x, sd["emb.weight"]) * math.sqrt(cfg.n_channels)
    h = h.transpose(1, 2)                                                  # (B, C, T)
    m = sequence_mask(x_lengths, h.shape[2]).unsqueeze(1).to(h.dtype)
    if cfg.prenet:
        h = prenet(sd, h, m)
    if trace is not None:
        trace["prenet"] = h.clone()
    if cfg.n_spks > 1:
        h = torch.cat([h, spks.unsqueeze(-1).expand(-1, -1, h.shape[-1])], dim=1)
    h = encoder(sd, h, m, cfg, trace)
    mu = F.conv1d(h, sd["proj_m.weight"], sd["proj_m.bias"]) * m
    logw = duration_predictor(sd, h, m, cfg)
    return mu, logw, m
