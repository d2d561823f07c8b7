"""Generate the committed golden vectors from the LIVE reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/model.py (read-only mount; never copied), instantiates the reference
`Decoder` / `CFM` / sub-blocks, loads the oracle's seeded state-dict into them with
`load_state_dict(strict=True)` and records the reference's outputs on seeded inputs.  The GPU
box has no /root/reference, so these files are what pins the oracle (and through it the CUDA
path) to the reference there.  Every case also asserts oracle == reference to 1e-5 so a
mismatch is caught at generation time.

Weights are NOT stored (11 M parameters): they are regenerated from the seed by
oracle.cfm_oracle.make_state_dict, which is deterministic for a given torch CPU generator.
A checksum of the state-dict is stored so that a drift of the generator would be detected.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import model as ref                                   # noqa: E402  (the reference itself)
from oracle import cfm_oracle as O                    # noqa: E402
from make_golden_cases import CASES, sd_checksum      # noqa: E402


def ref_decoder(cfg, sd):
    dec = ref.Decoder(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                      channels=(cfg.channels, cfg.channels), dropout=0.05,
                      attention_head_dim=cfg.head_dim, n_blocks=1, num_mid_blocks=cfg.n_mid,
                      num_heads=cfg.heads, act_fn="snakebeta")
    dec.load_state_dict(sd, strict=True)
    return dec.eval()


def ref_euler(dec, z0, mu, mask, n, spks, solver):
    """model.py:1086-1104 restated around the reference estimator so z0 can be injected."""
    z = z0.clone()
    B = z.shape[0]
    dt = torch.tensor([1.0 / n] * B, dtype=z.dtype)
    with torch.inference_mode():
        for i in range(n):
            t = torch.tensor([i / n] * B, dtype=z.dtype)
            pred = dec(z, mask, mu, t, spks, None)
            if solver == "euler":
                z = z + pred * dt[:, None, None]
            else:
                z_mid = z + pred * dt[:, None, None] * 0.5
                pred_mid = dec(z_mid, mask, mu, t + dt * 0.5, spks, None)
                z = z + pred_mid * dt[:, None, None]
    return z


def check(name, a, b, tol=2e-5):
    err = float((a - b).abs().max())
    print(f"  {name}: oracle vs reference max-abs {err:.2e}")
    assert err <= tol, (name, err)




def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    out = {}
    for name, cin, B, T, lengths, n, solver, seed in CASES:
        print(name)
        cfg = O.DecoderCfg(in_channels=cin)
        sd = O.make_state_dict(cfg, seed=0)
        dec = ref_decoder(cfg, sd)
        mu, mask, z0, spks = O.make_inputs(cfg, B, T, lengths, seed=seed)
        # single estimator call with per-row distinct t
        t = torch.linspace(0.05, 0.9, B)
        with torch.inference_mode():
            e_ref = dec(z0, mask, mu, t, spks, None)
        e_or = O.estimator_forward(sd, cfg, z0, mask, mu, t, spks)
        check("estimator", e_or, e_ref)
        z_ref = ref_euler(dec, z0, mu, mask, n, spks, solver)
        z_or = O.euler_solve(sd, cfg, z0, mu, mask, n, spks, solver)
        check("solve", z_or, z_ref)
        out[name + ".est"] = e_ref.numpy()
        out[name + ".z"] = z_ref.numpy()
        out[name + ".sdsum"] = np.float64(sd_checksum(sd))

    # --- behavioural pins of the sub-blocks ---------------------
    print("blocks")
    cfg = O.DecoderCfg()
    sd = O.make_state_dict(cfg, seed=0)
    dec = ref_decoder(cfg, sd)
    g = torch.Generator().manual_seed(99)
    B, L = 2, 20
    x = torch.randn(B, 256, L, generator=g)
    m = O.sequence_mask(torch.tensor([20, 13]), L).unsqueeze(1).float()
    temb = torch.randn(B, 1024, generator=g)
    with torch.inference_mode():
        r_ref = dec.mid_blocks[0][0](x, m, temb)
        t_ref = dec.mid_blocks[0][1][0](x.transpose(1, 2), attention_mask=m[:, 0, :], timestep=None)
        d_ref = dec.down_blocks[0][2](x)
        u_ref = dec.up_blocks[0][2](x)
        te_ref = dec.time_mlp(dec.time_embeddings(torch.tensor([0.0, 0.3, 0.9])))
    r_or = O.resnet_block(sd, "mid_blocks.0.0", x, m, temb, cfg, O.Emu())
    t_or = O.transformer_block(sd, "mid_blocks.0.1.0", x.transpose(1, 2), m[:, 0, :], cfg, O.Emu())
    te_or = O.time_embedding(sd, torch.tensor([0.0, 0.3, 0.9]), cfg)
    check("resnet", r_or, r_ref)
    check("transformer", t_or, t_ref)
    check("time_embedding", te_or, te_ref, 1e-4)
    out["blk.resnet"] = r_ref.numpy()
    out["blk.transformer"] = t_ref.numpy()
    out["blk.down"] = d_ref.numpy()
    out["blk.up"] = u_ref.numpy()
    out["blk.temb"] = te_ref.numpy()

    np.savez_compressed(os.path.join(HERE, "cfm_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "cfm_golden.npz"))
    print(f"wrote cfm_golden.npz ({sz/1024:.0f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
