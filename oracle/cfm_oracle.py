"""CPU oracle for the Matcha-TTS CFM decoder hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement (plain torch-CPU tensor ops over a flat
state-dict) of the reference's flow-matching decoder.  It is the *checker* for the CUDA
path and the CPU baseline for bench.py; nothing in the product package imports it.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it.

Parity pinning: the reference ships NO golden vectors or tests for this path (SURVEY.md
section 8c), so the oracle is pinned against *outputs of the reference itself*, produced by
importing /root/reference/model.py in the build container (tests/golden/make_golden.py,
vectors committed under tests/golden/).  tests/test_oracle.py re-checks the restatement
against those vectors everywhere, and against the live reference when it is mounted.

Reference lines followed (all in /root/reference/model.py):
  sinusoidal embedding      :753-762        timestep MLP        :828-832
  Block1D                   :773-775        ResnetBlock1D       :785-790
  Attention (mask quirk)    :670-705        BasicTransformerBlock :733-744
  SnakeBeta / FeedForward   :600-609, :641-644
  Downsample/Upsample       :797-798, :812-814
  Decoder.forward wiring    :964-1048       Euler / midpoint loop :1084-1109

`Emu` lets the tests emulate the storage precisions of the CUDA pipeline (fp16 GEMM
operands etc.) on the CPU, to pick formats that stay inside the parity tolerance.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# configuration / deterministic weights
# ----------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DecoderCfg:
    """Hyper-parameters of the estimator (reference main.py:63-79 literals)."""
    in_channels: int = 160          # 2*n_feats (+ spk_emb_dim for multi-speaker)
    out_channels: int = 80
    channels: int = 256             # reference channels=(256, 256): two levels, both 256 wide
    heads: int = 2
    head_dim: int = 64
    n_mid: int = 2
    groups: int = 8

    @property
    def time_dim(self) -> int:
        return self.channels * 4

    @property
    def attn_dim(self) -> int:
        return self.heads * self.head_dim

    @property
    def ff_dim(self) -> int:
        return self.channels * 4


def stage_names(cfg: DecoderCfg):
    """The six resnet+transformer stages in execution order, with their input widths."""
    c = cfg.channels
    return [
        ("down_blocks.0", cfg.in_channels), ("down_blocks.1", c),
        ("mid_blocks.0", c), ("mid_blocks.1", c),
        ("up_blocks.0", 2 * c), ("up_blocks.1", 2 * c),
    ]


def state_dict_spec(cfg: DecoderCfg):
    """(key, shape) list of the estimator state-dict, in a fixed order."""
    c, td, ad, fd = cfg.channels, cfg.time_dim, cfg.attn_dim, cfg.ff_dim
    spec = [
        ("time_mlp.linear_1.weight", (td, cfg.in_channels)), ("time_mlp.linear_1.bias", (td,)),
        ("time_mlp.linear_2.weight", (td, td)), ("time_mlp.linear_2.bias", (td,)),
    ]
    for name, ci in stage_names(cfg):
        r, t = name + ".0", name + ".1.0"
        spec += [
            (r + ".mlp.1.weight", (c, td)), (r + ".mlp.1.bias", (c,)),
            (r + ".block1.block.0.weight", (c, ci, 3)), (r + ".block1.block.0.bias", (c,)),
            (r + ".block1.block.1.weight", (c,)), (r + ".block1.block.1.bias", (c,)),
            (r + ".block2.block.0.weight", (c, c, 3)), (r + ".block2.block.0.bias", (c,)),
            (r + ".block2.block.1.weight", (c,)), (r + ".block2.block.1.bias", (c,)),
            (r + ".res_conv.weight", (c, ci, 1)), (r + ".res_conv.bias", (c,)),
            (t + ".norm1.weight", (c,)), (t + ".norm1.bias", (c,)),
            (t + ".attn1.to_q.weight", (ad, c)), (t + ".attn1.to_k.weight", (ad, c)),
            (t + ".attn1.to_v.weight", (ad, c)),
            (t + ".attn1.to_out.0.weight", (c, ad)), (t + ".attn1.to_out.0.bias", (c,)),
            (t + ".norm3.weight", (c,)), (t + ".norm3.bias", (c,)),
            (t + ".ff.net.0.alpha", (fd,)), (t + ".ff.net.0.beta", (fd,)),
            (t + ".ff.net.0.proj.weight", (fd, c)), (t + ".ff.net.0.proj.bias", (fd,)),
            (t + ".ff.net.2.weight", (c, fd)), (t + ".ff.net.2.bias", (c,)),
        ]
    spec += [
        ("down_blocks.0.2.conv.weight", (c, c, 3)), ("down_blocks.0.2.conv.bias", (c,)),
        ("down_blocks.1.2.weight", (c, c, 3)), ("down_blocks.1.2.bias", (c,)),
        ("up_blocks.0.2.conv.weight", (c, c, 4)), ("up_blocks.0.2.conv.bias", (c,)),   # (in,out,k)
        ("up_blocks.1.2.weight", (c, c, 3)), ("up_blocks.1.2.bias", (c,)),
        ("final_block.block.0.weight", (c, c, 3)), ("final_block.block.0.bias", (c,)),
        ("final_block.block.1.weight", (c,)), ("final_block.block.1.bias", (c,)),
        ("final_proj.weight", (cfg.out_channels, c, 1)), ("final_proj.bias", (cfg.out_channels,)),
    ]
    return spec


def make_state_dict(cfg: DecoderCfg, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic random-init estimator weights, reproducible on any box from the seed.

    Magnitudes follow PyTorch's default init (uniform +-1/sqrt(fan_in)) so activations have
    the same scale as a freshly constructed reference Decoder; norm gains/biases and the
    SnakeBeta log-scales are perturbed away from 1/0 so that every affine term is exercised.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = dict(state_dict_spec(cfg))
    for key, shape in shapes.items():
        leaf = key.rsplit(".", 1)[-1]
        is_norm = (".block.1." in key) or (".norm1." in key) or (".norm3." in key)
        if leaf in ("alpha", "beta"):
            v = 0.2 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            v = 0.1 * torch.randn(shape, generator=g)
        else:
            if leaf == "weight":
                if key == "up_blocks.0.2.conv.weight":        # ConvTranspose1d: (in, out, k)
                    fan_in = shape[1] * shape[2]
                else:
                    fan_in = int(torch.tensor(shape[1:]).prod())
            else:
                wshape = shapes[key[:-4] + "weight"]
                fan_in = int(torch.tensor(wshape[1:]).prod())
                if key == "up_blocks.0.2.conv.bias":
                    fan_in = wshape[1] * wshape[2]
            bound = 1.0 / math.sqrt(fan_in)
            v = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = v.float().contiguous()
    return sd


# ----------------------------------------------------------------------------------------
# precision emulation hooks (identity by default)
# ----------------------------------------------------------------------------------------
@dataclass
class Emu:
    """Where the CUDA pipeline rounds to 16 bit.  All None/False => pure fp32 oracle."""
    operand: Optional[torch.dtype] = None     # GEMM operands: activations AND weights
    conv_out: Optional[torch.dtype] = None    # raw conv output stored before GroupNorm
    attn: Optional[torch.dtype] = None        # q, k, v and the softmax probabilities
    resid: Optional[torch.dtype] = None       # transformer residual stream (resnet out, post-attention)
    trace: Optional[dict] = None              # if a dict: named intermediates are recorded into it
    prefix: str = ""

    def tr(self, name: str, x: Tensor) -> Tensor:
        if self.trace is not None:
            self.trace[self.prefix + name] = x.detach().clone()
        return x

    def op(self, x: Tensor) -> Tensor:
        return x if self.operand is None else x.to(self.operand).float()

    def co(self, x: Tensor) -> Tensor:
        return x if self.conv_out is None else x.to(self.conv_out).float()

    def at(self, x: Tensor) -> Tensor:
        return x if self.attn is None else x.to(self.attn).float()

    def rs(self, x: Tensor) -> Tensor:
        return x if self.resid is None else x.to(self.resid).float()


_NOEMU = Emu()


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def sinusoidal_embedding(t: Tensor, dim: int, scale: float = 1000.0) -> Tensor:
    """model.py:753-762 -- [sin(s*t*w_j), cos(s*t*w_j)], w_j = exp(-j*ln(1e4)/(dim/2-1))."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half).float() * -step)
    arg = scale * t.reshape(-1, 1).float() * freqs.reshape(1, -1)
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def time_embedding(sd, t: Tensor, cfg: DecoderCfg) -> Tensor:
    """model.py:971-972 + :828-832 -- (B,) -> (B, 4*channels); Linear -> SiLU -> Linear."""
    e = sinusoidal_embedding(t, cfg.in_channels)
    h = F.linear(e, sd["time_mlp.linear_1.weight"], sd["time_mlp.linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, sd["time_mlp.linear_2.weight"], sd["time_mlp.linear_2.bias"])


def _conv(x, w, b, emu: Emu, **kw):
    return F.conv1d(emu.op(x), emu.op(w), b, **kw)


def block1d(sd, pfx: str, x: Tensor, m: Tensor, cfg: DecoderCfg, emu: Emu) -> Tensor:
    """model.py:773-775 -- Mish(GroupNorm8(Conv1d k3 p1 (x*m))) * m; GN stats span ALL frames."""
    y = _conv(x * m, sd[pfx + ".block.0.weight"], sd[pfx + ".block.0.bias"], emu, padding=1)
    y = emu.tr("y." + pfx.rsplit(".", 1)[-1], emu.co(y))
    y = F.group_norm(y, cfg.groups, sd[pfx + ".block.1.weight"], sd[pfx + ".block.1.bias"], eps=1e-5)
    return F.mish(y) * m


def resnet_block(sd, pfx: str, x: Tensor, m: Tensor, temb: Tensor, cfg: DecoderCfg, emu: Emu) -> Tensor:
    """model.py:785-790 -- output is NOT masked (padded frames keep res_conv.bias)."""
    h = block1d(sd, pfx + ".block1", x, m, cfg, emu)
    tau = F.linear(F.mish(temb), sd[pfx + ".mlp.1.weight"], sd[pfx + ".mlp.1.bias"])
    h = h + tau.unsqueeze(-1)
    emu.tr("h1", h * m)
    h = block1d(sd, pfx + ".block2", h, m, cfg, emu)
    res = emu.tr("res", _conv(x * m, sd[pfx + ".res_conv.weight"], sd[pfx + ".res_conv.bias"], emu))
    return h + res


def attention(sd, pfx: str, a: Tensor, key_mask: Tensor, cfg: DecoderCfg, emu: Emu) -> Tensor:
    """model.py:670-705.  a: (B,L,C) LayerNorm'ed tokens; key_mask (B,L) float.

    Quirk reproduced on purpose (model.py:697): masked keys are filled with
    -finfo.min == +3.4e38, so a row with >=1 masked key attends uniformly to its MASKED keys.
    """
    B, L, _ = a.shape
    H, D = cfg.heads, cfg.head_dim
    ao = emu.op(a)
    q = F.linear(ao, emu.op(sd[pfx + ".to_q.weight"]))
    k = F.linear(ao, emu.op(sd[pfx + ".to_k.weight"]))
    v = F.linear(ao, emu.op(sd[pfx + ".to_v.weight"]))
    emu.tr("q", q); emu.tr("k", k); emu.tr("v", v)
    q, k, v = (emu.at(z).reshape(B, L, H, D).permute(0, 2, 1, 3) for z in (q, k, v))
    sim = torch.matmul(q, k.transpose(-1, -2)) * (D ** -0.5)
    fill = -torch.finfo(sim.dtype).min
    sim = sim.masked_fill(key_mask.reshape(B, 1, 1, L) == 0, fill)
    p = sim.softmax(dim=-1) if emu.attn is None else _emu_softmax(sim, emu)
    o = emu.tr("o", torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, L, H * D))
    return F.linear(emu.op(o), emu.op(sd[pfx + ".to_out.0.weight"]), sd[pfx + ".to_out.0.bias"])


def _emu_softmax(sim: Tensor, emu: Emu) -> Tensor:
    """Flash-style: un-normalised exp rounded to 16 bit for the PV product, fp32 row sum."""
    mx = sim.amax(dim=-1, keepdim=True)
    e = torch.exp(sim - mx)
    return emu.at(e) / e.sum(dim=-1, keepdim=True)


def snake_ff(sd, pfx: str, c: Tensor, emu: Emu) -> Tensor:
    """model.py:600-609 + :641-644 -- Linear -> u + sin^2(u*e^alpha)/(e^beta+1e-9) -> Linear."""
    u = F.linear(emu.op(c), emu.op(sd[pfx + ".net.0.proj.weight"]), sd[pfx + ".net.0.proj.bias"])
    ea = torch.exp(sd[pfx + ".net.0.alpha"])
    eb = torch.exp(sd[pfx + ".net.0.beta"])
    s = emu.tr("s", u + (1.0 / (eb + 1e-9)) * torch.sin(u * ea) ** 2)
    return F.linear(emu.op(s), emu.op(sd[pfx + ".net.2.weight"]), sd[pfx + ".net.2.bias"])


def transformer_block(sd, pfx: str, x: Tensor, key_mask: Tensor, cfg: DecoderCfg, emu: Emu) -> Tensor:
    """model.py:733-744 -- pre-LN self-attention then pre-LN SnakeBeta FF, both residual."""
    C = cfg.channels
    a = emu.tr("a", F.layer_norm(x, (C,), sd[pfx + ".norm1.weight"], sd[pfx + ".norm1.bias"], eps=1e-5))
    x = emu.tr("xa", emu.rs(attention(sd, pfx + ".attn1", a, key_mask, cfg, emu) + x))
    c = emu.tr("c", F.layer_norm(x, (C,), sd[pfx + ".norm3.weight"], sd[pfx + ".norm3.bias"], eps=1e-5))
    return snake_ff(sd, pfx + ".ff", c, emu) + x


def _stage(sd, name, x, m, temb, cfg, emu):
    emu.prefix = name + ":"
    x = emu.tr("xr", emu.rs(resnet_block(sd, name + ".0", x, m, temb, cfg, emu)))
    x = transformer_block(sd, name + ".1.0", x.transpose(1, 2), m[:, 0, :], cfg, emu)
    x = x.transpose(1, 2)
    emu.tr("out", x * m)          # the CUDA pipeline stores stage outputs already masked
    emu.prefix = ""
    return x


# ----------------------------------------------------------------------------------------
# estimator and solver
# ----------------------------------------------------------------------------------------
def estimator_forward(sd, cfg: DecoderCfg, x: Tensor, mask: Tensor, mu: Tensor, t: Tensor,
                      spks: Optional[Tensor] = None, emu: Emu = _NOEMU) -> Tensor:
    """model.py:964-1048.  x, mu: (B,80,T); mask: (B,1,T) float; t: (B,) -> (B,80,T)."""
    temb = emu.tr("temb", time_embedding(sd, t, cfg))
    x = torch.cat([x, mu], dim=1)
    if spks is not None:
        x = torch.cat([x, spks.unsqueeze(-1).expand(-1, -1, x.shape[-1])], dim=1)
    m0 = mask
    m1 = mask[:, :, ::2]

    x = _stage(sd, "down_blocks.0", x, m0, temb, cfg, emu)
    skip0 = x
    x = _conv(x * m0, sd["down_blocks.0.2.conv.weight"], sd["down_blocks.0.2.conv.bias"], emu,
              stride=2, padding=1)
    emu.tr("xD0", x * m1)
    x = _stage(sd, "down_blocks.1", x, m1, temb, cfg, emu)
    skip1 = x
    x = _conv(x * m1, sd["down_blocks.1.2.weight"], sd["down_blocks.1.2.bias"], emu, padding=1)
    emu.tr("xD1", x * m1)

    for i in range(cfg.n_mid):
        x = _stage(sd, f"mid_blocks.{i}", x, m1, temb, cfg, emu)

    x = _stage(sd, "up_blocks.0", torch.cat([x, skip1], dim=1), m1, temb, cfg, emu)
    x = F.conv_transpose1d(emu.op(x * m1), emu.op(sd["up_blocks.0.2.conv.weight"]),
                           sd["up_blocks.0.2.conv.bias"], stride=2, padding=1)
    if x.shape[-1] != skip0.shape[-1]:                       # odd T: nearest-resize == crop
        x = F.interpolate(x, size=skip0.shape[-1], mode="nearest")
    emu.tr("xU0", x * m0)
    x = _stage(sd, "up_blocks.1", torch.cat([x, skip0], dim=1), m0, temb, cfg, emu)
    x = _conv(x * m0, sd["up_blocks.1.2.weight"], sd["up_blocks.1.2.bias"], emu, padding=1)
    emu.tr("xF", x * m0)

    x = emu.tr("hF", block1d(sd, "final_block", x, m0, cfg, emu))
    out = _conv(x * m0, sd["final_proj.weight"], sd["final_proj.bias"], emu)
    return out * mask


def euler_solve(sd, cfg: DecoderCfg, z0: Tensor, mu: Tensor, mask: Tensor, n_timesteps: int,
                spks: Optional[Tensor] = None, solver: str = "euler", emu: Emu = _NOEMU) -> Tensor:
    """model.py:1084-1109 with the noise z0 injected (the reference draws it at :1085)."""
    z = z0.clone()
    B = z.shape[0]
    dt = 1.0 / n_timesteps
    dt_v = torch.tensor([dt] * B, dtype=z.dtype)
    for i in range(n_timesteps):
        t = torch.tensor([i / n_timesteps] * B, dtype=z.dtype)
        pred = estimator_forward(sd, cfg, z, mask, mu, t, spks, emu)
        if solver == "euler":
            z = z + pred * dt_v.reshape(B, 1, 1)
        elif solver == "midpoint":
            z_mid = z + pred * dt_v.reshape(B, 1, 1) * 0.5
            pred_mid = estimator_forward(sd, cfg, z_mid, mask, mu, t + dt_v * 0.5, spks, emu)
            z = z + pred_mid * dt_v.reshape(B, 1, 1)
        else:
            raise NotImplementedError(f"Solver {solver} not implemented")
    return z


# ----------------------------------------------------------------------------------------
# helpers shared by tests and bench
# ----------------------------------------------------------------------------------------
def sequence_mask(lengths: Tensor, max_length: int) -> Tensor:
    """model.py:42-46."""
    return torch.arange(max_length).unsqueeze(0) < lengths.unsqueeze(1)


def fix_len_compatibility(length: int, num_downsamplings: int = 2) -> int:
    """model.py:49-55 -- round up to a multiple of 2**num_downsamplings."""
    f = 2 ** num_downsamplings
    return int(math.ceil(length / f) * f)


def make_inputs(cfg: DecoderCfg, B: int, T: int, lengths=None, seed: int = 1, temperature: float = 0.667):
    """Seeded synthetic (mu, mask, z0, spks) of the named shape."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(B, cfg.out_channels, T, generator=g)
    z0 = torch.randn(B, cfg.out_channels, T, generator=g) * temperature
    n_spk = cfg.in_channels - 2 * cfg.out_channels
    spks = torch.randn(B, n_spk, generator=g) if n_spk > 0 else None
    if lengths is None:
        lengths = torch.full((B,), T, dtype=torch.long)
    lengths = torch.as_tensor(lengths, dtype=torch.long)
    mask = sequence_mask(lengths, T).unsqueeze(1).float()
    return mu, mask, z0, spks


def parity_errors(a: Tensor, ref: Tensor, mask: Tensor):
    """(max-abs, relative-L2) over valid frames -- the two numbers of the parity bar."""
    m = mask.bool().expand_as(ref)
    d = (a.double() - ref.double())[m]
    r = ref.double()[m]
    return float(d.abs().max()), float(d.norm() / r.norm().clamp_min(1e-30))


TOL_MAX_ABS = 2e-2      # BASELINE.json north_star
TOL_REL_L2 = 1e-3
