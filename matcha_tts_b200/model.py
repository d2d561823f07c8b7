"""Drop-in host side of the CFM decoder hot path.

Mirrors the reference's module surface for this path -- same class names, constructor and
forward signatures, state-dict keys and error behaviour -- and routes the compute to the
sm_100a kernels behind the C ABI of include/mtts.h (ctypes; PyTorch tensors are only buffers).

    Decoder.forward(x, mask, mu, t, spks=None, cond=None)             reference model.py:964-1048
    BASECFM / CFM .forward(mu, mask, n_timesteps, temperature, ...)    reference model.py:1084-1145
    MatchaTTS.synthesize / synthesise(x, x_lengths, n_timesteps, ...)  reference model.py:1264-1300
    sequence_mask, fix_len_compatibility, generate_path, denormalize   reference model.py:42-125

There is no CPU implementation here: tensors must live on a CUDA device and the extension must
be built (`matcha_tts_b200.build.build()`); anything else raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


# ----------------------------------------------------------------------------------------------
# small host helpers (reference model.py:42-125)
# ----------------------------------------------------------------------------------------------
def sequence_mask(length: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """Prefix mask (B, max_length) from lengths (reference model.py:42-46)."""
    if max_length is None:
        max_length = int(length.max())
    steps = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return steps[None, :] < length[:, None]


def fix_len_compatibility(length, num_downsamplings_in_unet: int = 2) -> int:
    """Round a length up to a multiple of 2**num_downsamplings (reference model.py:49-55)."""
    factor = 2 ** num_downsamplings_in_unet
    return int(math.ceil(float(length) / factor) * factor)


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Hard monotonic alignment from integer durations (reference model.py:64-76).

    duration: (B, T_x); mask: (B, T_x, T_y) -> path (B, T_x, T_y) with path[b, i, j] = 1 iff
    frame j belongs to token i.
    """
    b, t_x, t_y = mask.shape
    ends = torch.cumsum(duration, 1)                       # (B, T_x) exclusive end frame of token i
    starts = ends - duration
    frames = torch.arange(t_y, device=duration.device, dtype=ends.dtype)[None, None, :]
    path = ((frames < ends[:, :, None]) & (frames >= starts[:, :, None])).to(mask.dtype)
    return path * mask


def expand_by_duration(mu: torch.Tensor, duration: torch.Tensor, x_mask: torch.Tensor, y_mask: torch.Tensor) -> torch.Tensor:
    """mu_y = path^T . mu of reference model.py:1284-1288 as a gather: with a hard monotonic path (one token per
    frame) the dense (B, T_y, T_x) x (B, T_x, n_feats) product only copies column token(j) of mu to frame j, so the
    result is bit-identical.  mu: (B, n_feats, T_x); duration: (B, T_x) integer-valued frames per token;
    x_mask: (B, 1, T_x); y_mask: (B, 1, T_y) -> (B, n_feats, T_y).  Frames that no unmasked token covers are zero."""
    t_y = y_mask.shape[-1]
    ends = torch.cumsum(duration, 1)                                      # exclusive end frame of token i
    frames = torch.arange(t_y, device=mu.device, dtype=ends.dtype)[None, :].expand(mu.shape[0], t_y).contiguous()
    tok = torch.searchsorted(ends.contiguous(), frames, right=True)      # first token whose end lies beyond frame j
    covered = tok < duration.shape[1]
    tok = tok.clamp_max(duration.shape[1] - 1)
    keep = (covered & (torch.gather(x_mask[:, 0] != 0, 1, tok)) & (y_mask[:, 0] != 0)).to(mu.dtype)
    # like the reference's result -- a transposed view of a (B, T_y, n_feats) product (model.py:1288-1289) -- so that
    # randn_like(mu_y) in the sampler (model.py:1085) lays the noise out in the same memory order for the same seed
    out = torch.gather(mu.transpose(1, 2), 1, tok[:, :, None].expand(-1, -1, mu.shape[1])) * keep[:, :, None]
    return out.transpose(1, 2)


def denormalize(data: torch.Tensor, mu, std) -> torch.Tensor:
    """mel = data * std + mu (reference model.py:108-125); mu/std scalars or per-channel."""
    def prep(v):
        if isinstance(v, (float, int)):
            return v
        v = torch.as_tensor(v, dtype=data.dtype, device=data.device)
        return v.unsqueeze(-1) if v.ndim > 0 else v
    return data * prep(std) + prep(mu)


# ----------------------------------------------------------------------------------------------
# estimator parameter tree (state-dict layout of reference `Decoder`)
# ----------------------------------------------------------------------------------------------
def estimator_param_spec(in_channels: int, out_channels: int, channels: int, heads: int, head_dim: int,
                         n_mid: int):
    """[(state-dict key, shape)] of the reference Decoder, in the order libmtts lists them."""
    c, td, ad, fd = channels, channels * 4, heads * head_dim, channels * 4
    stages = [("down_blocks.0", in_channels), ("down_blocks.1", c)]
    stages += [(f"mid_blocks.{i}", c) for i in range(n_mid)]
    stages += [("up_blocks.0", 2 * c), ("up_blocks.1", 2 * c)]
    spec = [("time_mlp.linear_1.weight", (td, in_channels)), ("time_mlp.linear_1.bias", (td,)),
            ("time_mlp.linear_2.weight", (td, td)), ("time_mlp.linear_2.bias", (td,))]
    for name, ci in stages:
        r, t = f"{name}.0", f"{name}.1.0"
        spec += [(f"{r}.mlp.1.weight", (c, td)), (f"{r}.mlp.1.bias", (c,))]
        for blk, cin in (("block1", ci), ("block2", c)):
            spec += [(f"{r}.{blk}.block.0.weight", (c, cin, 3)), (f"{r}.{blk}.block.0.bias", (c,)),
                     (f"{r}.{blk}.block.1.weight", (c,)), (f"{r}.{blk}.block.1.bias", (c,))]
        spec += [(f"{r}.res_conv.weight", (c, ci, 1)), (f"{r}.res_conv.bias", (c,)),
                 (f"{t}.norm1.weight", (c,)), (f"{t}.norm1.bias", (c,)),
                 (f"{t}.attn1.to_q.weight", (ad, c)), (f"{t}.attn1.to_k.weight", (ad, c)),
                 (f"{t}.attn1.to_v.weight", (ad, c)),
                 (f"{t}.attn1.to_out.0.weight", (c, ad)), (f"{t}.attn1.to_out.0.bias", (c,)),
                 (f"{t}.norm3.weight", (c,)), (f"{t}.norm3.bias", (c,)),
                 (f"{t}.ff.net.0.alpha", (fd,)), (f"{t}.ff.net.0.beta", (fd,)),
                 (f"{t}.ff.net.0.proj.weight", (fd, c)), (f"{t}.ff.net.0.proj.bias", (fd,)),
                 (f"{t}.ff.net.2.weight", (c, fd)), (f"{t}.ff.net.2.bias", (c,))]
    spec += [("down_blocks.0.2.conv.weight", (c, c, 3)), ("down_blocks.0.2.conv.bias", (c,)),
             ("down_blocks.1.2.weight", (c, c, 3)), ("down_blocks.1.2.bias", (c,)),
             ("up_blocks.0.2.conv.weight", (c, c, 4)), ("up_blocks.0.2.conv.bias", (c,)),
             ("up_blocks.1.2.weight", (c, c, 3)), ("up_blocks.1.2.bias", (c,)),
             ("final_block.block.0.weight", (c, c, 3)), ("final_block.block.0.bias", (c,)),
             ("final_block.block.1.weight", (c,)), ("final_block.block.1.bias", (c,)),
             ("final_proj.weight", (out_channels, c, 1)), ("final_proj.bias", (out_channels,))]
    return spec


class _Node(nn.Module):
    """Parameter container; only exists so state-dict keys equal the reference's module paths."""


def _init_param(key: str, shape, spec: Dict[str, tuple]) -> torch.Tensor:
    """PyTorch-default-like initialisation (uniform +-1/sqrt(fan_in); norms 1/0; SnakeBeta 0)."""
    leaf = key.rsplit(".", 1)[-1]
    if leaf in ("alpha", "beta"):
        return torch.zeros(shape)
    if ".block.1." in key or ".norm1." in key or ".norm3." in key:
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    wshape = shape if leaf == "weight" else spec[key[:-4] + "weight"]
    if key.startswith("up_blocks.0.2.conv."):            # ConvTranspose1d weight is (in, out, k)
        fan_in = wshape[1] * wshape[2]
    else:
        fan_in = int(math.prod(wshape[1:]))
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape) * 2.0 - 1.0) * bound


# ----------------------------------------------------------------------------------------------
# native engine: one libmtts handle per (Decoder, device)
# ----------------------------------------------------------------------------------------------
def _aligned_buffer(nbytes: int, device, align: int = 1024) -> Tuple[torch.Tensor, int]:
    with torch.inference_mode(False):
        buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + align - 1) // align * align
    return buf, ptr


class _Engine:
    def __init__(self, cfg: _lib.MttsConfig, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.cfg = cfg
        h = C.c_void_p()
        # the library binds everything to the handle's device and restores the caller's current device (DeviceGuard in
        # mtts_api.cu); the torch-side allocations below are made under the same device for symmetry
        _lib.check(self.lib.mtts_create(C.byref(cfg), device.index or 0, C.byref(h)))
        self.h = h
        self.n_weights = self.lib.mtts_num_weights(h)
        self.names = [self.lib.mtts_weight_name(h, i).decode() for i in range(self.n_weights)]
        with torch.cuda.device(device):
            self.arena, self.arena_ptr = _aligned_buffer(self.lib.mtts_weight_arena_bytes(h), device, 256)
        self.ws: Dict[Tuple[int, int], Tuple[torch.Tensor, int, int]] = {}
        self.static: Dict[tuple, dict] = {}
        self.side_stream = None
        self.packed_version = None

    def __del__(self):
        try:
            torch.cuda.synchronize(self.device)     # nothing enqueued may still use the handle's graphs / arena
            self.lib.mtts_destroy(self.h)
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def load_weights(self, tensors: Dict[str, torch.Tensor]):
        st = self._stream()
        _lib.check(self.lib.mtts_set_weight_arena(self.h, self.arena_ptr, self.arena.numel() - 256 + 0, st))
        keep = []
        for i, name in enumerate(self.names):
            src = tensors[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(src)
            _lib.check(self.lib.mtts_load_weight(self.h, i, src.data_ptr(), src.numel(), st))
        torch.cuda.current_stream(self.device).synchronize()   # `keep` may be freed afterwards
        for key in list(self.ws):                              # plans and graphs captured over the old weights are gone
            self._drop_shape(key)
        self.static.clear()

    MAX_SHAPES = 24     # workspaces (and static graph buffers) kept per engine; least recently used shapes are dropped

    def workspace(self, B: int, T: int):
        key = (B, T)
        if key not in self.ws:
            n = self.lib.mtts_workspace_bytes(self.h, B, T)
            if n == 0:
                raise _lib.MttsError(f"unsupported shape B={B}, T={T}: need 1 <= B <= 2048 and T >= 1")
            while len(self.ws) >= self.MAX_SHAPES:           # bucketed serving sees many (B, T): bound the memory
                old = next(iter(self.ws))
                self._drop_shape(old)
            buf, ptr = _aligned_buffer(n, self.device, 1024)
            self.ws[key] = (buf, ptr, n)
        else:
            self.ws[key] = self.ws.pop(key)                  # most recently used last
        return self.ws[key]

    def _drop_shape(self, key):
        buf, ptr, n = self.ws.pop(key)
        torch.cuda.synchronize(self.device)                   # nothing enqueued may still use it
        _lib.check(self.lib.mtts_release_workspace(self.h, ptr, n))
        for sk in [k for k in self.static if k[:2] == key]:
            del self.static[sk]

    def estimator(self, x, mu, mask, t, spks, out):
        B, _, T = x.shape
        _, ws_ptr, ws_n = self.workspace(B, T)
        _lib.check(self.lib.mtts_estimator_forward(
            self.h, x.data_ptr(), mu.data_ptr(), mask.data_ptr(), t.data_ptr(),
            spks.data_ptr() if spks is not None else None, out.data_ptr(), ws_ptr, ws_n, B, T, self._stream()))

    def solve(self, z, mu, mask, spks, n_timesteps: int, solver: int, use_graph: bool):
        """In-place ODE solve on z.  With use_graph the inputs are staged into per-shape static
        buffers so the captured CUDA graph (keyed on pointers) is reused across calls."""
        B, _, T = z.shape
        _, ws_ptr, ws_n = self.workspace(B, T)
        if not use_graph:
            _lib.check(self.lib.mtts_euler_solve(
                self.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(),
                spks.data_ptr() if spks is not None else None, n_timesteps, solver, ws_ptr, ws_n, B, T, 0,
                self._stream()))
            return z
        key = (B, T, spks is not None)
        st = self.static.get(key)
        if st is None:
            with torch.inference_mode(False):     # plain tensors: reusable inside and outside inference_mode
                st = {"z": torch.empty(z.shape, dtype=z.dtype, device=z.device),
                      "mu": torch.empty(mu.shape, dtype=mu.dtype, device=mu.device),
                      "mask": torch.empty(mask.shape, dtype=mask.dtype, device=mask.device),
                      "spks": None if spks is None else torch.empty(spks.shape, dtype=spks.dtype, device=spks.device)}
            self.static[key] = st
        st["z"].copy_(z); st["mu"].copy_(mu); st["mask"].copy_(mask)
        if spks is not None:
            st["spks"].copy_(spks)
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:                 # the legacy default stream cannot be captured
            if self.side_stream is None:
                self.side_stream = torch.cuda.Stream(self.device)
            self.side_stream.wait_stream(cur)
            run = self.side_stream
        else:
            run = cur
        _lib.check(self.lib.mtts_euler_solve(
            self.h, st["z"].data_ptr(), st["mu"].data_ptr(), st["mask"].data_ptr(),
            st["spks"].data_ptr() if spks is not None else None, n_timesteps, solver, ws_ptr, ws_n, B, T, 1,
            run.cuda_stream))
        if run is not cur:
            cur.wait_stream(run)
        z.copy_(st["z"])
        return z

    def launch_count(self) -> int:
        return self.lib.mtts_last_launch_count(self.h)

    def set_chains(self, n: int):
        """Utterance chains per solve (0 = heuristic; 1 when several solves are kept in flight on several streams)."""
        _lib.check(self.lib.mtts_set_chains(self.h, int(n)))
        for key in list(self.ws):  # the workspace size depends on the chain layout
            self._drop_shape(key)

    def set_lanes(self, n: int):
        """Solves the caller keeps in flight on as many engines / streams: persistent launches take their share of the SMs."""
        _lib.check(self.lib.mtts_set_lanes(self.h, int(n)))


# ----------------------------------------------------------------------------------------------
# Decoder: the U-Net estimator
# ----------------------------------------------------------------------------------------------
class Decoder(nn.Module):
    """1-D U-Net vector-field estimator; constructor/forward as reference model.py:834-1048."""

    MAX_ENGINES = 8
    chains = 0      # utterance chains per solve for engines created from now on (see set_chains)
    lanes = 1       # solves kept in flight at a time on as many streams (see set_lanes)

    def __init__(self, in_channels, out_channels, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                 n_blocks=1, num_mid_blocks=2, num_heads=4, time_emb_dim=None, time_mlp_dim=None, ffn_mult=4,
                 **kwargs):
        super().__init__()
        channels = tuple(channels)
        if len(channels) != 2 or channels[0] != channels[1] or n_blocks != 1:
            raise NotImplementedError("native estimator supports channels=(C, C) with n_blocks=1 "
                                      "(the configuration the reference instantiates, main.py:63-79)")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.channels = channels
        self.num_heads = num_heads
        self.attention_head_dim = attention_head_dim
        self.num_mid_blocks = num_mid_blocks
        self._spec = estimator_param_spec(in_channels, out_channels, channels[0], num_heads, attention_head_dim,
                                          num_mid_blocks)
        shapes = dict(self._spec)
        for key, shape in self._spec:
            node = self
            *path, leaf = key.split(".")
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            node.register_parameter(leaf, nn.Parameter(_init_param(key, shape, shapes), requires_grad=False))
        # one native engine (libmtts handle + packed weights + workspace + CUDA graphs) per (device, CUDA stream):
        # solves issued on different streams are independent "lanes" that overlap on the GPU
        self._engines: Dict[Tuple[torch.device, int], _Engine] = {}

    # -- native plumbing ---------------------------------------------------------------------
    def _cfg(self) -> _lib.MttsConfig:
        return _lib.MttsConfig(self.in_channels, self.out_channels, self.channels[0], self.num_heads,
                               self.attention_head_dim, self.num_mid_blocks)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .half() swap the parameters' storage without touching their version counters
        self._storage_epoch = getattr(self, "_storage_epoch", 0) + 1
        return super()._apply(fn, *args, **kwargs)

    def _weights_version(self):
        """Changes whenever a parameter is written in place (version counters: load_state_dict, copy_, optimiser steps) or
        moved (storage epoch).  ~170 attribute reads per call instead of as many data_ptr() calls."""
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        return (getattr(self, "_storage_epoch", 0),) + tuple(p._version for p in plist)

    def _engine(self, device: torch.device) -> _Engine:
        if device.type != "cuda":
            raise RuntimeError("matcha_tts_b200.Decoder runs on CUDA (sm_100a) only; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= self.MAX_ENGINES:      # bound the handles a stream-hopping caller can create
                self._engines.pop(next(iter(self._engines)))
            eng = _Engine(self._cfg(), device)
            if self.chains:
                eng.set_chains(self.chains)
            if self.lanes > 1:
                eng.set_lanes(self.lanes)
            self._engines[key] = eng
        ver = self._weights_version()
        if eng.packed_version != ver:
            sd = {k: v for k, v in self.state_dict().items()}
            half = self.in_channels // 2
            # host-derived constant, computed exactly like reference model.py:757-758
            step = math.log(10000) / (half - 1)
            sd["@time_freqs"] = torch.exp(torch.arange(half).float() * -step)
            missing = [n for n in eng.names if n not in sd]
            if missing:
                raise KeyError(f"estimator weights missing for the native engine: {missing[:4]}...")
            eng.load_weights(sd)
            eng.packed_version = ver
        return eng

    @staticmethod
    def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None if t is None else t.detach().to(torch.float32).contiguous()

    def _check_inputs(self, x, mask, mu, spks):
        B, F_, T = x.shape
        if F_ != self.out_channels or mu.shape != x.shape:
            raise ValueError(f"x and mu must be (B, {self.out_channels}, T); got {tuple(x.shape)}, {tuple(mu.shape)}")
        if mask.numel() != B * T:
            raise ValueError(f"mask must be (B, 1, T); got {tuple(mask.shape)}")
        n_spk = self.in_channels - 2 * self.out_channels
        if (spks is None) != (n_spk == 0) or (spks is not None and tuple(spks.shape) != (B, n_spk)):
            raise ValueError(f"spks must be {'None' if n_spk == 0 else (B, n_spk)} for in_channels={self.in_channels}")

    # -- reference surface -------------------------------------------------------------------
    def forward(self, x, mask, mu, t, spks=None, cond=None):
        """x, mu: (B, n_feats, T); mask: (B, 1, T); t: (B,) -> (B, n_feats, T)."""
        self._check_inputs(x, mask, mu, spks)
        eng = self._engine(x.device)
        B = x.shape[0]
        t = torch.as_tensor(t, device=x.device, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        x32, mu32, m32, t32, s32 = self._f32c(x), self._f32c(mu), self._f32c(mask), self._f32c(t), self._f32c(spks)
        out = torch.empty_like(x32)
        eng.estimator(x32, mu32, m32, t32, s32, out)
        return out.to(x.dtype)

    def solve(self, z, mu, mask, n_timesteps: int, spks=None, solver: str = "euler", use_graph: bool = True):
        """Fixed-step ODE solve fused on the device (all n_timesteps in one enqueue / CUDA graph)."""
        self._check_inputs(z, mask, mu, spks)
        codes = {"euler": _lib.MTTS_SOLVER_EULER, "midpoint": _lib.MTTS_SOLVER_MIDPOINT}
        if solver not in codes:
            raise NotImplementedError(f"Solver {solver} not implemented")
        eng = self._engine(z.device)
        z32 = z.detach().to(torch.float32).contiguous().clone()
        eng.solve(z32, self._f32c(mu), self._f32c(mask), self._f32c(spks), int(n_timesteps), codes[solver], use_graph)
        return z32.to(z.dtype)

    def set_chains(self, n: int):
        """Utterance chains per solve: 0 = heuristic (best for one solve at a time), 1 = no split (use it when several
        solves are in flight on several CUDA streams: each stream has its own engine and they overlap each other)."""
        self.chains = int(n)
        for eng in self._engines.values():
            eng.set_chains(self.chains)

    def set_lanes(self, n: int):
        """Tell the native engines that the caller keeps `n` solves in flight at a time, each on its own CUDA stream (every
        stream has its own engine): with n > 1 a solve's persistent kernels size their grids for 1/n-th of the SMs (plus a
        quarter), which is what lets n batch-64 solves share the GPU without paying every CTA's fixed time on every SM
        (+15 % at n = 4).  Implies set_chains(1).  Results are bit-identical for every n."""
        self.lanes = max(1, int(n))
        if self.lanes > 1:
            self.set_chains(1)
        for eng in self._engines.values():
            eng.set_lanes(self.lanes)

    def last_launch_count(self, device=None) -> int:
        engs = [e for (d, _), e in self._engines.items() if device is None or d == torch.device(device)]
        return engs[-1].launch_count()


# ----------------------------------------------------------------------------------------------
# flow-matching sampler
# ----------------------------------------------------------------------------------------------
class BASECFM(nn.Module):
    """Fixed-step conditional-flow-matching sampler (reference model.py:1063-1109)."""

    def __init__(self, n_feats, cfm_params, n_spks=1, spk_emb_dim=64):
        super().__init__()
        self.n_feats = n_feats
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        self.solver = cfm_params.get("solver", "euler")
        self.sigma_min = cfm_params.get("sigma_min", 1e-4)
        self.use_cuda_graph = True

    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None):
        z = torch.randn_like(mu) * temperature            # same draw as reference model.py:1085
        return self.solve_from(z, mu, mask, n_timesteps, spks)

    def solve_from(self, z, mu, mask, n_timesteps, spks=None):
        """The loop of model.py:1089-1104 starting from a given z_0 (used for parity tests)."""
        if self.solver not in ("euler", "midpoint"):
            raise NotImplementedError(f"Solver {self.solver} not implemented")
        est = self.estimator
        if not isinstance(est, Decoder):
            raise TypeError("CFM.estimator must be a matcha_tts_b200.Decoder: the sampler runs natively on the GPU")
        return est.solve(z, mu, mask, int(n_timesteps), spks, self.solver, self.use_cuda_graph)


class CFM(BASECFM):
    """reference model.py:1114-1145 (inference surface; `compute_loss` is training-only, not provided)."""

    def __init__(self, n_feats, cfm_params, n_spks=1, spk_emb_dim=64, estimator=None):
        super().__init__(n_feats, cfm_params, n_spks=n_spks, spk_emb_dim=spk_emb_dim)
        if estimator is None:
            raise ValueError("estimator must be provided")
        self.estimator = estimator

    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None):
        return super().forward(mu=mu, mask=mask, n_timesteps=n_timesteps, temperature=temperature, spks=spks,
                               cond=cond)


# ----------------------------------------------------------------------------------------------
# model facade: the boundary the hot path is called through
# ----------------------------------------------------------------------------------------------
class MatchaTTS(nn.Module):
    """The reference's model class (model.py:1172-1300): same constructor, module tree / state-dict keys (`encoder.*`,
    `decoder.estimator.*`, `spk_emb.weight`, `mel_mean`, `mel_std`) and `synthesize` signature.

    `self.encoder` is the native TextEncoder + duration predictor (text_encoder.py; model.py:441-535), `self.decoder` the
    native CFM sampler; the glue between them -- durations -> lengths -> hard alignment -> mu_y -> denormalize -> crop
    (model.py:1272-1300) -- is a handful of torch calls.  `encoder=` substitutes any module with the TextEncoder interface
    `encoder(x, x_lengths, spks) -> (mu, logw, x_mask)`; with `encoder_params` that only carry `n_feats` and no `encoder=`
    the model is decoder-only and `synthesize` raises.
    """

    def __init__(self, n_vocab, n_spks, spk_emb_dim, encoder_params, decoder_params, cfm_params,
                 duration_predictor_params=None, encoder: Optional[nn.Module] = None):
        super().__init__()
        self.n_vocab = n_vocab
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        if n_spks > 1:
            self.spk_emb = nn.Embedding(n_spks, spk_emb_dim)
        self.register_buffer("mel_mean", torch.tensor(0.0))
        self.register_buffer("mel_std", torch.tensor(1.0))
        if encoder is not None:
            self.encoder = encoder
        elif hasattr(encoder_params, "n_channels"):
            from .text_encoder import TextEncoder
            self.encoder = TextEncoder(encoder_type=getattr(encoder_params, "encoder_type", "RoPE Encoder"),
                                       encoder_params=encoder_params, duration_predictor_params=duration_predictor_params,
                                       n_vocab=n_vocab, n_spks=n_spks, spk_emb_dim=spk_emb_dim)
        n_feats = encoder_params.n_feats
        in_ch = 2 * n_feats + (spk_emb_dim if n_spks > 1 else 0)
        est = Decoder(in_channels=in_ch, out_channels=n_feats, channels=decoder_params.channels,
                      dropout=decoder_params.dropout, attention_head_dim=decoder_params.attention_head_dim,
                      n_blocks=decoder_params.n_blocks, num_mid_blocks=decoder_params.num_mid_blocks,
                      num_heads=decoder_params.num_heads, act_fn=getattr(decoder_params, "act_fn", "snakebeta"))
        self.decoder = CFM(n_feats=n_feats, cfm_params=cfm_params, n_spks=n_spks, spk_emb_dim=spk_emb_dim,
                           estimator=est)

    @torch.inference_mode()
    def synthesize(self, x, x_lengths, n_timesteps, temperature=1.0, spks=None, length_scale=1.0):
        if not hasattr(self, "encoder"):
            raise RuntimeError("this MatchaTTS was built decoder-only (encoder_params without n_channels and no encoder=): "
                               "synthesize needs the text encoder, see the class docstring")
        mu, logw, x_mask = self.encoder(x, x_lengths, spks)
        w = torch.exp(logw) * x_mask * length_scale
        w_ceil = torch.ceil(w)
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max_length = int(y_lengths.max())                       # host sync, as model.py:1281
        y_max_length_ = fix_len_compatibility(y_max_length)
        y_mask = sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
        mu_y = expand_by_duration(mu, w_ceil.squeeze(1), x_mask, y_mask)   # == attn^T . mu (model.py:1288), as a gather
        mel = self.decoder(mu_y, y_mask, n_timesteps, temperature, spks, cond=None)
        mel = denormalize(mel, self.mel_mean, self.mel_std)
        return mel[:, :, :y_max_length], y_lengths, attn

    synthesise = synthesize     # British spelling used by BASELINE.json / the upstream package
