"""Golden vectors of the reference TextEncoder + duration predictor, from the LIVE reference.

    python tests/golden/make_text_golden.py        (build container only: imports /root/reference/model.py, read-only)

The oracle's seeded state-dict (oracle.text_encoder_oracle.make_state_dict) is loaded into the reference module with
load_state_dict(strict=True) -- which also pins the key names and shapes -- and the reference's outputs on seeded token
ids are stored in tests/golden/text_golden.npz.  Weights are not stored; a checksum of the state-dict is.
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import model as ref                                     # noqa: E402  (the reference itself)
from oracle import text_encoder_oracle as TO            # noqa: E402

# name, n_spks, B, T_x, lengths, seed
CASES = [("lj_b3", 1, 3, 23, [23, 17, 5], 11), ("lj_b1", 1, 1, 40, [40], 12), ("vctk_b2", 109, 2, 19, [19, 8], 13)]


def checksum(sd):
    return float(sum(float(v.double().abs().sum()) for v in sd.values()))


def inputs(cfg, B, T, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg.n_vocab, (B, T), generator=g)
    spks = torch.randn(B, cfg.spk_emb_dim, generator=g) if cfg.n_spks > 1 else None
    return x, torch.tensor(lengths), spks


def ref_encoder(cfg, sd):
    ep = types.SimpleNamespace(encoder_type="RoPE Encoder", n_feats=cfg.n_feats, n_channels=cfg.n_channels,
                               filter_channels=cfg.filter_channels, n_heads=cfg.n_heads, n_layers=cfg.n_layers,
                               kernel_size=cfg.kernel_size, p_dropout=0.1, prenet=cfg.prenet)
    dp = types.SimpleNamespace(filter_channels_dp=cfg.filter_channels_dp, kernel_size=cfg.kernel_size_dp, p_dropout=0.1)
    enc = ref.TextEncoder("RoPE Encoder", ep, dp, cfg.n_vocab, n_spks=cfg.n_spks, spk_emb_dim=cfg.spk_emb_dim)
    enc.load_state_dict(sd, strict=True)
    return enc.eval()


def main():
    out = {}
    for name, n_spks, B, T, lengths, seed in CASES:
        cfg = TO.TextEncCfg(n_spks=n_spks)
        sd = TO.make_state_dict(cfg, seed=0)
        x, xl, spks = inputs(cfg, B, T, lengths, seed)
        with torch.no_grad():
            mu, logw, mask = ref_encoder(cfg, sd)(x, xl, spks)
            mu_o, logw_o, mask_o = TO.text_encoder_forward(sd, cfg, x, xl, spks)
        for a, b, what in ((mu, mu_o, "mu"), (logw, logw_o, "logw"), (mask, mask_o, "mask")):
            err = float((a - b).abs().max())
            assert err <= 2e-5, (name, what, err)
            print(f"{name}: oracle vs reference {what} max-abs {err:.2e}")
        out[name + ".mu"] = mu.numpy()
        out[name + ".logw"] = logw.numpy()
        out[name + ".mask"] = mask.numpy()
        out[name + ".sd_checksum"] = np.float64(checksum(sd))
    np.savez_compressed(os.path.join(HERE, "text_golden.npz"), **out)
    print("wrote text_golden.npz:", {k: v.shape for k, v in out.items() if v.ndim})


if __name__ == "__main__":
    main()
