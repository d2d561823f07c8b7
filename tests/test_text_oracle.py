"""Oracle of the reference TextEncoder + duration predictor (model.py:148-535) against the
golden vectors generated from the live reference, plus the properties a native implementation may rely on."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import text_encoder_oracle as TO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "text_golden.npz")
HAVE_REF = os.path.exists("/root/reference/model.py")
# name, n_spks, B, T_x, lengths, seed -- the cases of tests/golden/make_text_golden.py
CASES = [("lj_b3", 1, 3, 23, [23, 17, 5], 11), ("lj_b1", 1, 1, 40, [40], 12), ("vctk_b2", 109, 2, 19, [19, 8], 13)]


def _inputs(cfg, B, T, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg.n_vocab, (B, T), generator=g)
    spks = torch.randn(B, cfg.spk_emb_dim, generator=g) if cfg.n_spks > 1 else None
    return x, torch.tensor(lengths), spks


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_text_oracle_matches_reference_golden(case):
    name, n_spks, B, T, lengths, seed = case
    gold = np.load(GOLD)
    cfg = TO.TextEncCfg(n_spks=n_spks)
    sd = TO.make_state_dict(cfg, seed=0)
    chk = float(sum(float(v.double().abs().sum()) for v in sd.values()))
    assert abs(chk - float(gold[name + ".sd_checksum"])) <= 1e-6 * chk        # same seeded weights as at generation time
    x, xl, spks = _inputs(cfg, B, T, lengths, seed)
    mu, logw, mask = TO.text_encoder_forward(sd, cfg, x, xl, spks)
    assert torch.equal(mask, torch.from_numpy(gold[name + ".mask"]))
    assert float((mu - torch.from_numpy(gold[name + ".mu"])).abs().max()) <= 2e-5
    assert float((logw - torch.from_numpy(gold[name + ".logw"])).abs().max()) <= 2e-5


def test_state_dict_layout():
    cfg = TO.TextEncCfg()
    shapes = TO.param_shapes(cfg)
    assert len(shapes) == 1 + 14 + 6 * 16 + 2 + 10
    assert sum(int(np.prod(s)) for s in shapes.values()) == 18_204_193 - 11_008_848     # MatchaTTS minus the estimator (both measured in SURVEY section 8c)
    assert TO.param_shapes(TO.TextEncCfg(n_spks=109))["encoder.attn_layers.0.conv_q.weight"] == (256, 256, 1)


def test_padding_is_inert_and_outputs_are_masked():
    """Padded tokens contribute nothing (attention fills them with -1e4 -> exp underflows to 0, convs see x*mask): a row
    gives the same mu / logw alone at its own length as inside a longer batch, and both are 0 at padded positions."""
    cfg = TO.TextEncCfg()
    sd = TO.make_state_dict(cfg, seed=0)
    x, xl, _ = _inputs(cfg, 3, 23, [23, 17, 5], 21)
    mu, logw, mask = TO.text_encoder_forward(sd, cfg, x, xl)
    assert float((mu * (1 - mask)).abs().max()) == 0.0 and float((logw * (1 - mask)).abs().max()) == 0.0
    for b, n in enumerate([23, 17, 5]):
        mu1, logw1, _ = TO.text_encoder_forward(sd, cfg, x[b:b + 1, :n], xl[b:b + 1])
        assert float((mu1 - mu[b:b + 1, :, :n]).abs().max()) <= 2e-5
        assert float((logw1 - logw[b:b + 1, :, :n]).abs().max()) <= 2e-5


def test_rope_is_a_rotation():
    """RoPE acts on the first d features only and preserves their norm (model.py:244-289)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 11, 96, generator=g)
    y = TO.rope(x, 48)
    assert torch.equal(y[..., 48:], x[..., 48:])
    assert torch.allclose(y[..., :48].norm(dim=-1), x[..., :48].norm(dim=-1), atol=1e-5)
    assert torch.allclose(y[:, :, 0], x[:, :, 0])                          # position 0: angle 0


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_text_oracle_equals_live_reference():
    import make_text_golden as G
    cfg = TO.TextEncCfg()
    sd = TO.make_state_dict(cfg, seed=5)
    x, xl, _ = _inputs(cfg, 2, 31, [31, 12], 33)
    with torch.no_grad():
        mu_r, logw_r, mask_r = G.ref_encoder(cfg, sd)(x, xl, None)
    mu, logw, mask = TO.text_encoder_forward(sd, cfg, x, xl)
    assert torch.equal(mask, mask_r)
    assert float((mu - mu_r).abs().max()) <= 2e-5 and float((logw - logw_r).abs().max()) <= 2e-5
