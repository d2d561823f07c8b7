"""Drop-in host side of the step before the hot path: the reference TextEncoder + duration predictor
(reference model.py:441-535; blocks :148-438), run by the sm_100a kernels behind `mtts_text_*` of include/mtts.h.

    TextEncoder(encoder_type, encoder_params, duration_predictor_params, n_vocab, n_spks=1, spk_emb_dim=128)   model.py:441-498
    TextEncoder.forward(x, x_lengths, spks=None) -> (mu, logw, x_mask)                                         model.py:500-535

Same constructor / forward signature and state-dict keys (`emb.weight`, `prenet.conv_layers.0.weight`, ...,
`encoder.attn_layers.5.conv_q.weight`, ..., `proj_w.proj.bias`) as the reference module, so a reference checkpoint loads
with strict=True (matcha_tts_b200.checkpoint.load_lightning_checkpoint).  PyTorch tensors are only buffers; there is no CPU
implementation here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


def text_encoder_param_spec(n_vocab: int, n_channels: int, filter_channels: int, n_layers: int, kernel_size: int, prenet: bool,
                            n_feats: int, filter_channels_dp: int, kernel_size_dp: int, width: int):
    """[(state-dict key, shape)] of the reference TextEncoder, in the order libmtts lists them (model.py:441-498)."""
    Cc, W, Fc, k, D, kd = n_channels, width, filter_channels, kernel_size, filter_channels_dp, kernel_size_dp
    spec = [("emb.weight", (n_vocab, Cc))]
    if prenet:                                            # ConvReluNorm(C, C, C, kernel_size=5, n_layers=3), model.py:463-471
        for i in range(3):
            spec += [(f"prenet.conv_layers.{i}.weight", (Cc, Cc, 5)), (f"prenet.conv_layers.{i}.bias", (Cc,)),
                     (f"prenet.norm_layers.{i}.gamma", (Cc,)), (f"prenet.norm_layers.{i}.beta", (Cc,))]
        spec += [("prenet.proj.weight", (Cc, Cc, 1)), ("prenet.proj.bias", (Cc,))]
    for i in range(n_layers):                             # Encoder, model.py:412-425
        for n in ("q", "k", "v", "o"):
            spec += [(f"encoder.attn_layers.{i}.conv_{n}.weight", (W, W, 1)), (f"encoder.attn_layers.{i}.conv_{n}.bias", (W,))]
        spec += [(f"encoder.norm_layers_1.{i}.gamma", (W,)), (f"encoder.norm_layers_1.{i}.beta", (W,)),
                 (f"encoder.ffn_layers.{i}.conv_1.weight", (Fc, W, k)), (f"encoder.ffn_layers.{i}.conv_1.bias", (Fc,)),
                 (f"encoder.ffn_layers.{i}.conv_2.weight", (W, Fc, k)), (f"encoder.ffn_layers.{i}.conv_2.bias", (W,)),
                 (f"encoder.norm_layers_2.{i}.gamma", (W,)), (f"encoder.norm_layers_2.{i}.beta", (W,))]
    spec += [("proj_m.weight", (n_feats, W, 1)), ("proj_m.bias", (n_feats,)),
             ("proj_w.conv_1.weight", (D, W, kd)), ("proj_w.conv_1.bias", (D,)), ("proj_w.norm_1.gamma", (D,)), ("proj_w.norm_1.beta", (D,)),
             ("proj_w.conv_2.weight", (D, D, kd)), ("proj_w.conv_2.bias", (D,)), ("proj_w.norm_2.gamma", (D,)), ("proj_w.norm_2.beta", (D,)),
             ("proj_w.proj.weight", (1, D, 1)), ("proj_w.proj.bias", (1,))]
    return spec


class _Node(nn.Module):
    """Parameter container; only exists so state-dict keys equal the reference's module paths."""


def _init(key: str, shape, n_channels: int) -> torch.Tensor:
    """The reference's initialisation: N(0, C^-1/2) embedding (model.py:459), xavier-uniform q/k/v (:327-332), zero prenet.proj
    (:197-198), gamma 1 / beta 0 (:149-150), PyTorch's conv default elsewhere."""
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "gamma":
        return torch.ones(shape)
    if leaf == "beta" or key.startswith("prenet.proj."):
        return torch.zeros(shape)
    if key == "emb.weight":
        return torch.randn(shape) * n_channels ** -0.5
    if leaf == "weight" and any(f".conv_{n}." in key for n in "qkv"):
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape) * 2.0 - 1.0) * bound
    return None   # filled from the fan-in of the matching weight by the caller


def _aligned_buffer(nbytes: int, device, align: int = 1024) -> Tuple[torch.Tensor, int]:
    with torch.inference_mode(False):
        buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + align - 1) // align * align
    return buf, ptr


class _TextEngine:
    """One libmtts text-encoder handle per (TextEncoder, device): packed weights + per-shape workspaces."""
    MAX_SHAPES = 32

    def __init__(self, cfg: _lib.MttsTextConfig, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        _lib.check(self.lib.mtts_text_create(C.byref(cfg), device.index or 0, C.byref(h)))
        self.h = h
        self.names = [self.lib.mtts_text_weight_name(h, i).decode() for i in range(self.lib.mtts_text_num_weights(h))]
        with torch.cuda.device(device):
            self.arena, self.arena_ptr = _aligned_buffer(self.lib.mtts_text_weight_arena_bytes(h), device, 256)
        self.ws: Dict[Tuple[int, int], Tuple[torch.Tensor, int, int]] = {}
        self.static: Dict[Tuple[int, int], dict] = {}
        self.side_stream = None
        self.packed_version = None
        self.n_feats = cfg.n_feats

    def __del__(self):
        try:
            torch.cuda.synchronize(self.device)
            self.lib.mtts_text_destroy(self.h)
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def load_weights(self, tensors: Dict[str, torch.Tensor]):
        st = self._stream()
        _lib.check(self.lib.mtts_text_set_weight_arena(self.h, self.arena_ptr, self.arena.numel() - 256, st))
        keep = []
        for i, name in enumerate(self.names):
            src = tensors[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(src)
            _lib.check(self.lib.mtts_text_load_weight(self.h, i, src.data_ptr(), src.numel(), st))
        torch.cuda.current_stream(self.device).synchronize()   # `keep` may be freed afterwards
        self.ws.clear()                                        # set_weight_arena dropped the plans and graphs
        self.static.clear()

    def workspace(self, B: int, Tx: int):
        key = (B, Tx)
        if key not in self.ws:
            n = self.lib.mtts_text_workspace_bytes(self.h, B, Tx)
            if n == 0:
                raise _lib.MttsError(f"unsupported shape B={B}, T_x={Tx}")
            while len(self.ws) >= self.MAX_SHAPES:
                old = next(iter(self.ws))
                buf, ptr, nb = self.ws.pop(old)
                torch.cuda.synchronize(self.device)
                _lib.check(self.lib.mtts_text_release_workspace(self.h, ptr, nb))
                self.static.pop(old, None)
            buf, ptr = _aligned_buffer(n, self.device, 1024)
            self.ws[key] = (buf, ptr, n)
        else:
            self.ws[key] = self.ws.pop(key)
        return self.ws[key]

    def forward(self, tokens, lengths, spks, use_graph: bool = True):
        """-> (mu, logw, x_mask).  With use_graph the inputs are staged into per-shape static buffers and the call's launches
        replay as one CUDA graph (keyed on pointers inside the library); the outputs are fresh tensors either way."""
        B, Tx = tokens.shape
        _, ptr, n = self.workspace(B, Tx)
        dev = tokens.device
        if not use_graph:
            mu = torch.empty(B, self.n_feats, Tx, dtype=torch.float32, device=dev)
            logw = torch.empty(B, 1, Tx, dtype=torch.float32, device=dev)
            x_mask = torch.empty(B, 1, Tx, dtype=torch.float32, device=dev)
            _lib.check(self.lib.mtts_text_encoder_forward(
                self.h, tokens.data_ptr(), lengths.data_ptr(), spks.data_ptr() if spks is not None else None, mu.data_ptr(),
                logw.data_ptr(), x_mask.data_ptr(), ptr, n, B, Tx, 0, self._stream()))
            return mu, logw, x_mask
        key = (B, Tx)
        st = self.static.get(key)
        if st is None:
            with torch.inference_mode(False):
                st = {"tok": torch.empty(B, Tx, dtype=torch.int64, device=dev), "len": torch.empty(B, dtype=torch.int64, device=dev),
                      "spk": None if spks is None else torch.empty(spks.shape, dtype=torch.float32, device=dev),
                      "mu": torch.empty(B, self.n_feats, Tx, dtype=torch.float32, device=dev),
                      "logw": torch.empty(B, 1, Tx, dtype=torch.float32, device=dev),
                      "mask": torch.empty(B, 1, Tx, dtype=torch.float32, device=dev)}
            self.static[key] = st
        st["tok"].copy_(tokens); st["len"].copy_(lengths)
        if spks is not None:
            st["spk"].copy_(spks)
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:                 # the legacy default stream cannot be captured
            if self.side_stream is None:
                self.side_stream = torch.cuda.Stream(self.device)
            self.side_stream.wait_stream(cur)
            run = self.side_stream
        else:
            run = cur
        _lib.check(self.lib.mtts_text_encoder_forward(
            self.h, st["tok"].data_ptr(), st["len"].data_ptr(), st["spk"].data_ptr() if spks is not None else None, st["mu"].data_ptr(),
            st["logw"].data_ptr(), st["mask"].data_ptr(), ptr, n, B, Tx, 1, run.cuda_stream))
        if run is not cur:
            cur.wait_stream(run)
        return st["mu"].clone(), st["logw"].clone(), st["mask"].clone()

    def launch_count(self) -> int:
        return self.lib.mtts_text_last_launch_count(self.h)


class TextEncoder(nn.Module):
    """Text encoder + duration predictor; constructor / forward as reference model.py:441-535."""

    def __init__(self, encoder_type, encoder_params, duration_predictor_params, n_vocab, n_spks=1, spk_emb_dim=128):
        super().__init__()
        self.encoder_type = encoder_type
        self.encoder_params = encoder_params
        self.n_vocab = n_vocab
        self.n_feats = encoder_params.n_feats
        self.n_channels = encoder_params.n_channels
        self.spk_emb_dim = spk_emb_dim
        self.n_spks = n_spks
        self.width = self.n_channels + (spk_emb_dim if n_spks > 1 else 0)
        self._cfg_tuple = (n_vocab, self.n_feats, self.n_channels, encoder_params.filter_channels, encoder_params.n_heads,
                           encoder_params.n_layers, encoder_params.kernel_size, 1 if encoder_params.prenet else 0,
                           duration_predictor_params.filter_channels_dp, duration_predictor_params.kernel_size,
                           n_spks, spk_emb_dim)
        self._spec = text_encoder_param_spec(n_vocab, self.n_channels, encoder_params.filter_channels, encoder_params.n_layers,
                                             encoder_params.kernel_size, bool(encoder_params.prenet), self.n_feats,
                                             duration_predictor_params.filter_channels_dp, duration_predictor_params.kernel_size,
                                             self.width)
        shapes = dict(self._spec)
        for key, shape in self._spec:
            v = _init(key, shape, self.n_channels)
            if v is None:
                wshape = shape if key.endswith("weight") else shapes[key[:-4] + "weight"]
                bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
                v = (torch.rand(shape) * 2.0 - 1.0) * bound
            node = self
            *path, leaf = key.split(".")
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            node.register_parameter(leaf, nn.Parameter(v, requires_grad=False))
        self._engines: Dict[torch.device, _TextEngine] = {}
        self.use_cuda_graph = True

    def _apply(self, fn, *args, **kwargs):
        self._storage_epoch = getattr(self, "_storage_epoch", 0) + 1      # .to() / .cuda() swap storage without bumping versions
        return super()._apply(fn, *args, **kwargs)

    def _weights_version(self):
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        return (getattr(self, "_storage_epoch", 0),) + tuple(p._version for p in plist)

    def _engine(self, device: torch.device) -> _TextEngine:
        if device.type != "cuda":
            raise RuntimeError("matcha_tts_b200.TextEncoder runs on CUDA (sm_100a) only; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engines.get(device)
        if eng is None:
            eng = _TextEngine(_lib.MttsTextConfig(*self._cfg_tuple), device)
            self._engines[device] = eng
        ver = self._weights_version()
        if eng.packed_version != ver:
            sd = dict(self.state_dict())
            d = int((self.width // self.encoder_params.n_heads) * 0.5)                 # model.py:319-320
            sd["@rope_theta"] = 1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))   # model.py:262, computed like the reference
            missing = [n for n in eng.names if n not in sd]
            if missing:
                raise KeyError(f"text encoder weights missing for the native engine: {missing[:4]}...")
            eng.load_weights(sd)
            eng.packed_version = ver
        return eng

    def forward(self, x, x_lengths, spks=None):
        """x: (B, T_x) token ids; x_lengths: (B,); spks: (B, spk_emb_dim) already embedded (model.py:523) or None
        -> mu (B, n_feats, T_x), logw (B, 1, T_x), x_mask (B, 1, T_x)."""
        if x.ndim != 2 or x_lengths.shape != (x.shape[0],):
            raise ValueError(f"x must be (B, T_x) and x_lengths (B,); got {tuple(x.shape)}, {tuple(x_lengths.shape)}")
        if (self.n_spks > 1) != (spks is not None) or (spks is not None and tuple(spks.shape) != (x.shape[0], self.spk_emb_dim)):
            raise ValueError(f"spks must be {'(B, %d)' % self.spk_emb_dim if self.n_spks > 1 else 'None'} for n_spks={self.n_spks}")
        eng = self._engine(x.device)
        B, Tx = x.shape
        tok = x.detach().to(torch.int64).contiguous()
        lens = x_lengths.detach().to(device=x.device, dtype=torch.int64).contiguous()
        s32 = None if spks is None else spks.detach().to(device=x.device, dtype=torch.float32).contiguous()
        mu, logw, x_mask = eng.forward(tok, lens, s32, use_graph=self.use_cuda_graph)
        return mu, logw, x_mask

    def last_launch_count(self) -> int:
        return next(reversed(self._engines.values())).launch_count()
