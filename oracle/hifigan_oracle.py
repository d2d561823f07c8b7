"""CPU oracle of the step AFTER the hot path: the vendored HiFi-GAN `Generator`
(`/root/reference/hifigan/models.py:148-206`, ResBlock1 `:14-98`, config v1 `hifigan/config.py:1-28`) and the waveglow-style
`Denoiser` (`/root/reference/hifigan/denoiser.py:12-68`), restated as plain fp32 torch functions over a flat state-dict.

TEST INFRASTRUCTURE ONLY -- like cfm_oracle.py it may be imported by `tests/`, `__graft_entry__.smoke()` and bench.py's
CPU legs, never by the package.  Pinned by `tests/golden/hifigan_golden.npz`: outputs of the LIVE reference modules on
seeded weights and mels (`tests/golden/make_hifigan_golden.py`).

Boundary the native version keeps: `Generator(h)(mel) -> wav` with mel (B, 80, T) float, wav (B, 1, 256 T) in (-1, 1)
(main.py:196-198 calls `vocoder(mel).clamp(-1, 1)`); `Denoiser(vocoder)(audio, strength) -> audio` with audio (B, n).
State-dict keys are the reference's, either weight-normed (`*.weight_g`, `*.weight_v`, as in the published checkpoint that
main.py:146-147 loads) or plain (`*.weight`, after `remove_weight_norm()`, main.py:149).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU_SLOPE = 0.1          # hifigan/models.py:11


@dataclass(frozen=True)
class HifiganCfg:
    """hifigan/config.py v1 (the only configuration main.py:140 instantiates)."""
    num_mels: int = 80
    upsample_initial_channel: int = 512
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))

    @property
    def hop(self) -> int:
        return int(math.prod(self.upsample_rates))

    def channels(self, level: int) -> int:     # level 0 = conv_pre output, level i = after ups[i-1]
        return self.upsample_initial_channel // (2 ** level)


def get_padding(kernel_size: int, dilation: int = 1) -> int:     # hifigan/xutils.py:37-38
    return int((kernel_size * dilation - dilation) / 2)


def param_shapes(cfg: HifiganCfg = HifiganCfg()) -> Dict[str, tuple]:
    """Key -> shape of the Generator's state-dict after remove_weight_norm() (module tree of models.py:149-179)."""
    s: Dict[str, tuple] = {"conv_pre.weight": (cfg.upsample_initial_channel, cfg.num_mels, 7),
                           "conv_pre.bias": (cfg.upsample_initial_channel,)}
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        ci, co = cfg.channels(i), cfg.channels(i + 1)
        s[f"ups.{i}.weight"] = (ci, co, k)       # ConvTranspose1d: (in, out, k)
        s[f"ups.{i}.bias"] = (co,)
        for j, rk in enumerate(cfg.resblock_kernel_sizes):
            for which in ("convs1", "convs2"):
                for m in range(len(cfg.resblock_dilation_sizes[j])):
                    s[f"resblocks.{i * nk + j}.{which}.{m}.weight"] = (co, co, rk)
                    s[f"resblocks.{i * nk + j}.{which}.{m}.bias"] = (co,)
    cl = cfg.channels(len(cfg.upsample_rates))
    s["conv_post.weight"] = (1, cl, 7)
    s["conv_post.bias"] = (1,)
    return s


def make_state_dict(cfg: HifiganCfg = HifiganCfg(), seed: int = 0, gain: float = 1.0) -> Dict[str, Tensor]:
    """Seeded plain state-dict with O(1) activations at every level (the reference's init, std 0.01, lets the signal die out
    after two layers, which would test nothing): conv weights N(0, gain^2 / fan_in) with fan_in the number of inputs that
    reach one output, biases N(0, 0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
            continue
        if name.startswith("ups."):
            i = int(name.split(".")[1])
            fan = shape[0] * shape[2] / cfg.upsample_rates[i]
            std = gain / math.sqrt(fan)
        elif name.startswith("resblocks."):
            std = 0.7 * gain / math.sqrt(shape[1] * shape[2])
        elif name.startswith("conv_post."):
            std = 0.25 * gain / math.sqrt(shape[1] * shape[2])
        else:
            std = gain / math.sqrt(shape[1] * shape[2])
        sd[name] = std * torch.randn(shape, generator=g)
    return sd


def to_weight_norm(sd: Dict[str, Tensor], seed: int = 1) -> Dict[str, Tensor]:
    """The same weights in the checkpoint's weight-normed form (torch.nn.utils.weight_norm, dim 0: `weight = g * v / |v|`
    with the norm over every dimension but the first): v is the weight times a random positive factor per slice."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, w in sd.items():
        if not name.endswith(".weight"):
            out[name] = w.clone()
            continue
        f = 0.5 + torch.rand((w.shape[0], 1, 1), generator=g)
        out[name + "_v"] = w * f
        out[name + "_g"] = w.flatten(1).norm(dim=1).view(-1, 1, 1)
    return out


def fold_weight_norm(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """remove_weight_norm() (models.py:197-205) on a state-dict: `*.weight_g`, `*.weight_v` -> `*.weight`."""
    out: Dict[str, Tensor] = {}
    for name, t in sd.items():
        if name.endswith(".weight_v"):
            base = name[: -len("_v")]
            v, g = t.float(), sd[base + "_g"].float()
            out[base] = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))
        elif name.endswith(".weight_g"):
            continue
        else:
            out[name] = t.float()
    return out


def resblock1(sd: Dict[str, Tensor], prefix: str, x: Tensor, k: int, dilations, trace: Optional[dict] = None) -> Tensor:
    """ResBlock1.forward, models.py:84-91."""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{prefix}.convs1.{m}.weight"], sd[f"{prefix}.convs1.{m}.bias"], dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{prefix}.convs2.{m}.weight"], sd[f"{prefix}.convs2.{m}.bias"], dilation=1, padding=get_padding(k, 1))
        x = xt + x
        if trace is not None:
            trace[f"{prefix}.pair{m}"] = x
    return x


def generator_forward(sd: Dict[str, Tensor], mel: Tensor, cfg: HifiganCfg = HifiganCfg(), trace: Optional[dict] = None) -> Tensor:
    """Generator.forward, models.py:181-195.  `sd` is a plain state-dict (fold_weight_norm first if needed)."""
    nk = len(cfg.resblock_kernel_sizes)
    x = F.conv1d(mel.float(), sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    if trace is not None:
        trace["conv_pre"] = x
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if trace is not None:
            trace[f"ups.{i}"] = x
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            r = resblock1(sd, f"resblocks.{i * nk + j}", x, rk, rd, trace)
            xs = r if xs is None else xs + r
        x = xs / nk
        if trace is not None:
            trace[f"level.{i}"] = x
    x = F.leaky_relu(x)            # default slope 0.01 (models.py:191)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def flops_per_frame(cfg: HifiganCfg = HifiganCfg()) -> float:
    """Multiply-add FLOPs (2 per MAC) of Generator.forward per mel frame, dense count at the real channel widths."""
    f = 2.0 * cfg.num_mels * cfg.upsample_initial_channel * 7
    rate = 1
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        ci, co = cfg.channels(i), cfg.channels(i + 1)
        f += 2.0 * ci * co * k * rate          # every input sample meets every tap once
        rate *= u
        for rk, rd in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
            f += 2.0 * co * co * rk * 2 * len(rd) * rate
    f += 2.0 * cfg.channels(len(cfg.upsample_rates)) * 7 * rate
    return f


# ----------------------------------------------------------------------------------------------- denoiser
def stft_mag_phase(audio: Tensor, n_fft: int = 1024, hop: int = 256, win: int = 1024) -> Tuple[Tensor, Tensor]:
    """denoiser.py:29-39: torch.stft (centered, reflect padding, periodic Hann window) -> magnitude, phase."""
    spec = torch.stft(audio, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), return_complex=True)
    spec = torch.view_as_real(spec)
    return torch.sqrt(spec.pow(2).sum(-1)), torch.atan2(spec[..., -1], spec[..., 0])


def denoiser_bias_spec(sd: Dict[str, Tensor], cfg: HifiganCfg = HifiganCfg(), n_fft: int = 1024, hop: int = 256, win: int = 1024) -> Tensor:
    """denoiser.py:21-22, 56-60 (mode "zeros"): magnitude spectrum of the vocoder's output on an all-zero mel, first frame."""
    bias_audio = generator_forward(sd, torch.zeros(1, cfg.num_mels, 88), cfg).float().squeeze(0)
    bias_spec, _ = stft_mag_phase(bias_audio, n_fft, hop, win)
    return bias_spec[:, :, 0][:, :, None]


def denoiser_forward(audio: Tensor, bias_spec: Tensor, strength: float = 0.0005, n_fft: int = 1024, hop: int = 256,
                     win: int = 1024) -> Tensor:
    """Denoiser.forward, denoiser.py:62-68."""
    mag, ang = stft_mag_phase(audio, n_fft, hop, win)
    mag = torch.clamp(mag - bias_spec * strength, 0.0)
    return torch.istft(torch.complex(mag * torch.cos(ang), mag * torch.sin(ang)), n_fft=n_fft, hop_length=hop, win_length=win,
                       window=torch.hann_window(win))
