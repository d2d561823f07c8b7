"""Drop-in host side of the step after the hot path: the vendored HiFi-GAN vocoder of the reference
(`hifigan/models.py:148-206` Generator, `hifigan/config.py` v1, `hifigan/env.py` AttrDict, `hifigan/denoiser.py` Denoiser),
run by the sm_100a kernels behind `mtts_voc_*` of include/mtts.h.

    Generator(h)                      hifigan/models.py:149-179      h = AttrDict(v1)
    Generator.forward(mel) -> wav     hifigan/models.py:181-195      mel (B, 80, T) -> wav (B, 1, 256 T)
    Generator.remove_weight_norm()    hifigan/models.py:197-205
    Denoiser(vocoder, ...)(audio, strength)                          hifigan/denoiser.py:12-68

Same constructor / forward signatures and state-dict keys as the reference modules: a freshly constructed Generator holds the
weight-normed parameters (`conv_pre.weight_g`, `conv_pre.weight_v`, `conv_pre.bias`, ...), so the published checkpoint loads
with `vocoder.load_state_dict(state["generator"])` exactly as in main.py:146-147, and `remove_weight_norm()` turns them into
plain `*.weight` (main.py:149).  PyTorch tensors are only buffers; there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib
from .text_encoder import _Node, _aligned_buffer

# hifigan/config.py:1-28 (the architecture keys; the training keys of the reference dict are irrelevant at inference)
v1 = {
    "resblock": "1",
    "upsample_rates": [8, 8, 2, 2],
    "upsample_kernel_sizes": [16, 16, 4, 4],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "num_mels": 80,
    "n_fft": 1024,
    "hop_size": 256,
    "win_size": 1024,
    "sampling_rate": 22050,
}


class AttrDict(dict):
    """hifigan/env.py:7-10."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def generator_param_spec(h) -> list:
    """[(state-dict key after remove_weight_norm(), shape)] of the reference Generator, in the order libmtts lists them."""
    C0 = h.upsample_initial_channel
    nk = len(h.resblock_kernel_sizes)
    spec = [("conv_pre.weight", (C0, getattr(h, "num_mels", 80), 7)), ("conv_pre.bias", (C0,))]
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        ci, co = C0 // (2 ** i), C0 // (2 ** (i + 1))
        spec += [(f"ups.{i}.weight", (ci, co, k)), (f"ups.{i}.bias", (co,))]
        for j, rk in enumerate(h.resblock_kernel_sizes):
            for which in ("convs1", "convs2"):
                for m in range(len(h.resblock_dilation_sizes[j])):
                    spec += [(f"resblocks.{i * nk + j}.{which}.{m}.weight", (co, co, rk)), (f"resblocks.{i * nk + j}.{which}.{m}.bias", (co,))]
    cl = C0 // (2 ** len(h.upsample_rates))
    spec += [("conv_post.weight", (1, cl, 7)), ("conv_post.bias", (1,))]
    return spec


def _voc_config(h) -> _lib.MttsVocConfig:
    if str(getattr(h, "resblock", "1")) != "1":
        raise NotImplementedError("the native vocoder implements ResBlock1 (config v1, what main.py:140 instantiates); resblock='2' is not built")
    nu, nk = len(h.upsample_rates), len(h.resblock_kernel_sizes)
    nd = len(h.resblock_dilation_sizes[0])
    if nu > 4 or nk > 3 or nd > 3 or any(len(d) != nd for d in h.resblock_dilation_sizes) or len(h.upsample_kernel_sizes) != nu:
        raise NotImplementedError("the native vocoder supports up to 4 upsampling stages, 3 resblocks per stage and 3 dilations per resblock")
    cfg = _lib.MttsVocConfig()
    cfg.num_mels = int(getattr(h, "num_mels", 80))
    cfg.upsample_initial_channel = int(h.upsample_initial_channel)
    cfg.n_ups, cfg.n_resblocks, cfg.n_dilations = nu, nk, nd
    for i in range(nu):
        cfg.upsample_rates[i] = int(h.upsample_rates[i])
        cfg.upsample_kernel_sizes[i] = int(h.upsample_kernel_sizes[i])
    for j in range(nk):
        cfg.resblock_kernel_sizes[j] = int(h.resblock_kernel_sizes[j])
        for m in range(nd):
            cfg.resblock_dilation_sizes[j][m] = int(h.resblock_dilation_sizes[j][m])
    return cfg


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`*.weight_g`, `*.weight_v` -> `*.weight` = g * v / |v| (torch.nn.utils.weight_norm, dim 0: the norm runs over every axis
    but the first -- also for ConvTranspose1d, whose first axis is the input channel); other entries pass through."""
    out: Dict[str, torch.Tensor] = {}
    for name, t in sd.items():
        if name.endswith(".weight_v"):
            base = name[: -len("_v")]
            v, g = t.detach().float(), sd[base + "_g"].detach().float()
            out[base] = v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.ndim - 1))))
        elif name.endswith(".weight_g"):
            continue
        else:
            out[name] = t.detach().float()
    return out


class _VocEngine:
    """One libmtts vocoder handle per (Generator, device): packed weights + per-shape workspaces."""
    MAX_SHAPES = 4

    def __init__(self, cfg: _lib.MttsVocConfig, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        _lib.check(self.lib.mtts_voc_create(C.byref(cfg), device.index or 0, C.byref(h)))
        self.h = h
        self.names = [self.lib.mtts_voc_weight_name(h, i).decode() for i in range(self.lib.mtts_voc_num_weights(h))]
        self.hop = self.lib.mtts_voc_hop_length(h)
        with torch.cuda.device(device):
            self.arena, self.arena_ptr = _aligned_buffer(self.lib.mtts_voc_weight_arena_bytes(h), device, 1024)
        self.ws: Dict[Tuple[int, int], Tuple[torch.Tensor, int, int]] = {}
        self.static: Dict[Tuple[int, int], dict] = {}
        self.side_stream = None
        self.packed_version = None

    def __del__(self):
        try:
            torch.cuda.synchronize(self.device)
            self.lib.mtts_voc_destroy(self.h)
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def load_weights(self, tensors: Dict[str, torch.Tensor]):
        st = self._stream()
        _lib.check(self.lib.mtts_voc_set_weight_arena(self.h, self.arena_ptr, self.arena.numel() - 1024, st))
        keep = []
        for i, name in enumerate(self.names):
            src = tensors[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(src)
            _lib.check(self.lib.mtts_voc_load_weight(self.h, i, src.data_ptr(), src.numel(), st))
        torch.cuda.current_stream(self.device).synchronize()   # `keep` may be freed afterwards
        self.ws.clear()                                        # set_weight_arena dropped the plans and graphs
        self.static.clear()

    def workspace(self, B: int, T: int):
        key = (B, T)
        if key not in self.ws:
            n = self.lib.mtts_voc_workspace_bytes(self.h, B, T)
            if n == 0:
                raise _lib.MttsError(f"unsupported shape B={B}, T={T}")
            while len(self.ws) >= self.MAX_SHAPES:
                old = next(iter(self.ws))
                buf, ptr, nb = self.ws.pop(old)
                torch.cuda.synchronize(self.device)
                _lib.check(self.lib.mtts_voc_release_workspace(self.h, ptr, nb))
                self.static.pop(old, None)
            buf, ptr = _aligned_buffer(n, self.device, 1024)
            self.ws[key] = (buf, ptr, n)
        else:
            self.ws[key] = self.ws.pop(key)
        return self.ws[key]

    def forward(self, mel: torch.Tensor, use_graph: bool = True) -> torch.Tensor:
        B, _, T = mel.shape
        _, ptr, n = self.workspace(B, T)
        dev = mel.device
        if not use_graph:
            wav = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=dev)
            _lib.check(self.lib.mtts_voc_generator_forward(self.h, mel.data_ptr(), wav.data_ptr(), ptr, n, B, T, 0, self._stream()))
            return wav
        key = (B, T)
        st = self.static.get(key)
        if st is None:
            with torch.inference_mode(False):
                st = {"mel": torch.empty(B, mel.shape[1], T, dtype=torch.float32, device=dev),
                      "wav": torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=dev)}
            self.static[key] = st
        st["mel"].copy_(mel)
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:                 # the legacy default stream cannot be captured
            if self.side_stream is None:
                self.side_stream = torch.cuda.Stream(self.device)
            self.side_stream.wait_stream(cur)
            run = self.side_stream
        else:
            run = cur
        _lib.check(self.lib.mtts_voc_generator_forward(self.h, st["mel"].data_ptr(), st["wav"].data_ptr(), ptr, n, B, T, 1, run.cuda_stream))
        if run is not cur:
            cur.wait_stream(run)
        return st["wav"].clone()

    def launch_count(self) -> int:
        return self.lib.mtts_voc_last_launch_count(self.h)


class Generator(nn.Module):
    """HiFi-GAN generator; constructor / forward / remove_weight_norm as reference hifigan/models.py:148-206."""

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self._vcfg = _voc_config(h)                       # raises for architectures the native engine does not implement
        self._spec = generator_param_spec(h)
        for key, shape in self._spec:
            node = self
            *path, leaf = key.split(".")
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            if leaf == "bias":
                wshape = dict(self._spec)[key[:-4] + "weight"]
                bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
                node.register_parameter("bias", nn.Parameter((torch.rand(shape) * 2.0 - 1.0) * bound, requires_grad=False))
                continue
            # weight_norm(Conv) (models.py:153, :159-170) with init_weights N(0, 0.01) everywhere but conv_pre (xutils.py:25-28)
            if key.startswith("conv_pre."):
                bound = 1.0 / math.sqrt(int(math.prod(shape[1:])))
                v = (torch.rand(shape) * 2.0 - 1.0) * bound
            else:
                v = torch.randn(shape) * 0.01
            node.register_parameter("weight_g", nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1), requires_grad=False))
            node.register_parameter("weight_v", nn.Parameter(v, requires_grad=False))
        self._engines: Dict[torch.device, _VocEngine] = {}
        self.use_cuda_graph = True

    def remove_weight_norm(self):
        """models.py:197-205: every `weight_g` / `weight_v` pair becomes a plain `weight`."""
        print("Removing weight norm...")
        for mod in self.modules():
            if isinstance(mod, _Node) and "weight_v" in mod._parameters:
                w = fold_weight_norm({"x.weight_g": mod.weight_g, "x.weight_v": mod.weight_v})["x.weight"]
                del mod._parameters["weight_g"], mod._parameters["weight_v"]
                mod.register_parameter("weight", nn.Parameter(w.to(mod.bias.device), requires_grad=False))
        self.__dict__.pop("_plist", None)
        self._storage_epoch = getattr(self, "_storage_epoch", 0) + 1

    def _apply(self, fn, *args, **kwargs):
        self._storage_epoch = getattr(self, "_storage_epoch", 0) + 1      # .to() / .cuda() swap storage without bumping versions
        return super()._apply(fn, *args, **kwargs)

    def _weights_version(self):
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        return (getattr(self, "_storage_epoch", 0),) + tuple(p._version for p in plist)

    def _engine(self, device: torch.device) -> _VocEngine:
        if device.type != "cuda":
            raise RuntimeError("matcha_tts_b200.hifigan.Generator runs on CUDA (sm_100a) only; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engines.get(device)
        if eng is None:
            eng = _VocEngine(self._vcfg, device)
            self._engines[device] = eng
        ver = self._weights_version()
        if eng.packed_version != ver:
            sd = fold_weight_norm(dict(self.state_dict()))
            missing = [n for n in eng.names if n not in sd]
            if missing:
                raise KeyError(f"vocoder weights missing for the native engine: {missing[:4]}...")
            eng.load_weights(sd)
            eng.packed_version = ver
        return eng

    def forward(self, x):
        """x: mel (B, num_mels, T) -> waveform (B, 1, T * hop) in (-1, 1)."""
        if x.ndim != 3 or x.shape[1] != self._vcfg.num_mels:
            raise ValueError(f"mel must be (B, {self._vcfg.num_mels}, T); got {tuple(x.shape)}")
        eng = self._engine(x.device)
        mel = x.detach().to(torch.float32).contiguous()
        return eng.forward(mel, use_graph=self.use_cuda_graph)

    def last_launch_count(self) -> int:
        return next(reversed(self._engines.values())).launch_count()


class ModeException(Exception):
    """hifigan/denoiser.py:7-8."""


class Denoiser(nn.Module):
    """Removes the vocoder's bias from generated audio; constructor / forward as reference hifigan/denoiser.py:12-68.

    `vocoder` is any module mapping a (1, 80, 88) mel to a (1, 1, n) waveform on a CUDA device (normally the native Generator).
    The STFT / inverse STFT run on libmtts' own FFT kernels (mtts_stft_magnitude, mtts_denoiser_forward)."""

    def __init__(self, vocoder, filter_length=1024, n_overlap=4, win_length=1024, mode="zeros"):
        super().__init__()
        if filter_length != 1024 or n_overlap != 4 or win_length != 1024:
            raise NotImplementedError("the native Denoiser implements the reference's configuration only: filter_length=1024, n_overlap=4, "
                                      "win_length=1024")
        self.filter_length = filter_length
        self.hop_length = int(filter_length / n_overlap)
        self.win_length = win_length
        prm = next(vocoder.parameters())
        dtype, device = prm.dtype, prm.device
        if device.type != "cuda":
            raise RuntimeError("matcha_tts_b200.hifigan.Denoiser runs on CUDA (sm_100a) only: move the vocoder to the GPU first")
        self.device = device
        self._lib = _lib.load()
        if mode == "zeros":
            mel_input = torch.zeros((1, 80, 88), dtype=dtype, device=device)
        elif mode == "normal":
            mel_input = torch.randn((1, 80, 88), dtype=dtype, device=device)
        else:
            raise ModeException(f"Mode {mode} if not supported")
        with torch.no_grad():
            bias_audio = vocoder(mel_input).float().squeeze(0)                    # (1, n)
            bias_spec = self.stft_magnitude(bias_audio)                           # (1, 513, F)
        self.register_buffer("bias_spec", bias_spec[:, :, 0][:, :, None].contiguous())
        self._ws = None

    def stft_magnitude(self, audio: torch.Tensor) -> torch.Tensor:
        """|torch.stft(audio, 1024, hop 256, win 1024, periodic Hann, centred)| : (B, n) -> (B, 513, 1 + n // 256)."""
        audio = audio.detach().to(torch.float32).contiguous()
        B, n = audio.shape
        F = self._lib.mtts_stft_frames(n)
        mag = torch.empty(B, 513, F, dtype=torch.float32, device=audio.device)
        _lib.check(self._lib.mtts_stft_magnitude(audio.device.index or 0, audio.data_ptr(), mag.data_ptr(), B, n,
                                                 torch.cuda.current_stream(audio.device).cuda_stream))
        return mag

    @torch.inference_mode()
    def forward(self, audio, strength=0.0005):
        """audio (B, n) -> denoised audio (B, 256 * (n // 256))."""
        if audio.ndim != 2:
            raise ValueError(f"audio must be (B, n); got {tuple(audio.shape)}")
        if audio.device.type != "cuda":
            raise RuntimeError("matcha_tts_b200.hifigan.Denoiser runs on CUDA (sm_100a) only; there is no CPU path")
        a = audio.detach().to(torch.float32).contiguous()
        B, n = a.shape
        nbytes = self._lib.mtts_stft_workspace_bytes(B, n)
        if nbytes == 0:
            raise _lib.MttsError(f"unsupported shape B={B}, n={n}: the centred frames need more than 512 samples")
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != a.device:
            with torch.inference_mode(False):
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        out = torch.empty(B, 256 * (self._lib.mtts_stft_frames(n) - 1), dtype=torch.float32, device=a.device)
        bias = self.bias_spec.to(device=a.device, dtype=torch.float32).contiguous()
        _lib.check(self._lib.mtts_denoiser_forward(a.device.index or 0, a.data_ptr(), bias.data_ptr(), float(strength), out.data_ptr(),
                                                   self._ws.data_ptr(), self._ws.numel(), B, n,
                                                   torch.cuda.current_stream(a.device).cuda_stream))
        return out


def load_vocoder(path: str, device) -> Generator:
    """main.py:134-150 without the download: Generator(AttrDict(v1)), `state["generator"]` loaded strictly, eval mode, weight norm
    removed."""
    vocoder = Generator(AttrDict(v1)).to(device)
    state = torch.load(path, map_location=device)
    vocoder.load_state_dict(state["generator"])
    vocoder.eval()
    vocoder.remove_weight_norm()
    return vocoder
