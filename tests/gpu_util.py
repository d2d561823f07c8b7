"""Helpers shared by the GPU parity tests (not collected by pytest)."""
import ctypes as C

import torch

from oracle import cfm_oracle as O


def make_decoder(cin=160, seed=0, device="cuda"):
    """Native Decoder loaded with the oracle's seeded weights; returns (decoder, cfg, state_dict)."""
    from matcha_tts_b200 import Decoder
    cfg = O.DecoderCfg(in_channels=cin)
    sd = O.make_state_dict(cfg, seed)
    dec = Decoder(in_channels=cin, out_channels=80, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                  n_blocks=1, num_mid_blocks=2, num_heads=2, act_fn="snakebeta")
    dec.load_state_dict(sd, strict=True)
    return dec.to(device), cfg, sd


def ws_tensor(eng, B, T, name, rows, cols, dtype=torch.float16):
    """View of a named intermediate inside the engine's workspace as a (rows, cols) tensor."""
    buf, ptr, n = eng.workspace(B, T)
    off = eng.lib.mtts_debug_buffer_offset(eng.h, B, T, 0, name.encode())
    assert off >= 0, name
    start = ptr - buf.data_ptr() + off
    esz = torch.empty((), dtype=dtype).element_size()
    return buf[start:start + rows * cols * esz].view(dtype).reshape(rows, cols)


def flat_to_bct(x, B, L, Lp):
    """(B*Lp, C) flat channels-last rows -> (B, C, L), dropping guard rows."""
    C_ = x.shape[1]
    return x.reshape(B, Lp, C_)[:, :L].permute(0, 2, 1).float()


def errs(a, ref):
    d = (a.double().cpu() - ref.double().cpu())
    return float(d.abs().max()), float(d.norm() / ref.double().norm().clamp_min(1e-30))
