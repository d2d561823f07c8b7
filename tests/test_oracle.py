"""The CPU oracle against the committed golden vectors (outputs of the live reference) and the
behavioural pins listed in SURVEY.md section 8(c)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import cfm_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_cases import CASES, sd_checksum  # noqa: E402


@pytest.fixture(scope="module")
def sds():
    return {c: O.make_state_dict(O.DecoderCfg(in_channels=c), seed=0) for c in (160, 224)}


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(case, golden, sds):
    name, cin, B, T, lengths, n, solver, seed = case
    cfg = O.DecoderCfg(in_channels=cin)
    sd = sds[cin]
    # the seeded weights must be the ones the goldens were made with
    assert abs(sd_checksum(sd) - float(golden[name + ".sdsum"])) < 1e-6 * abs(float(golden[name + ".sdsum"])) + 1e-6
    mu, mask, z0, spks = O.make_inputs(cfg, B, T, lengths, seed=seed)
    t = torch.linspace(0.05, 0.9, B)
    est = O.estimator_forward(sd, cfg, z0, mask, mu, t, spks)
    np.testing.assert_allclose(est.numpy(), golden[name + ".est"], rtol=0, atol=2e-5)
    z = O.euler_solve(sd, cfg, z0, mu, mask, n, spks, solver)
    np.testing.assert_allclose(z.numpy(), golden[name + ".z"], rtol=0, atol=2e-5)


def test_block_goldens(golden, sds):
    cfg, sd = O.DecoderCfg(), sds[160]
    g = torch.Generator().manual_seed(99)
    B, L = 2, 20
    x = torch.randn(B, 256, L, generator=g)
    m = O.sequence_mask(torch.tensor([20, 13]), L).unsqueeze(1).float()
    temb = torch.randn(B, 1024, generator=g)
    r = O.resnet_block(sd, "mid_blocks.0.0", x, m, temb, cfg, O.Emu())
    t = O.transformer_block(sd, "mid_blocks.0.1.0", x.transpose(1, 2), m[:, 0, :], cfg, O.Emu())
    te = O.time_embedding(sd, torch.tensor([0.0, 0.3, 0.9]), cfg)
    np.testing.assert_allclose(r.numpy(), golden["blk.resnet"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(t.numpy(), golden["blk.transformer"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(te.numpy(), golden["blk.temb"], atol=1e-4, rtol=0)
    d = torch.nn.functional.conv1d(x, sd["down_blocks.0.2.conv.weight"], sd["down_blocks.0.2.conv.bias"], stride=2, padding=1)
    u = torch.nn.functional.conv_transpose1d(x, sd["up_blocks.0.2.conv.weight"], sd["up_blocks.0.2.conv.bias"], stride=2, padding=1)
    np.testing.assert_allclose(d.numpy(), golden["blk.down"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(u.numpy(), golden["blk.up"], atol=2e-5, rtol=0)


def test_attention_mask_quirk(sds):
    """model.py:697: a row with padding attends uniformly to its PADDED keys; without padding it is
    ordinary softmax attention."""
    cfg, sd = O.DecoderCfg(), sds[160]
    g = torch.Generator().manual_seed(5)
    B, L = 2, 12
    a = torch.randn(B, L, 256, generator=g)
    km = O.sequence_mask(torch.tensor([12, 7]), L).float()
    out = O.attention(sd, "mid_blocks.0.1.0.attn1", a, km, cfg, O.Emu())
    pfx = "mid_blocks.0.1.0.attn1"
    v = a @ sd[pfx + ".to_v.weight"].T
    # padded row: every query gets mean of V over padded keys
    vm = v[1, 7:].mean(0)
    exp1 = vm @ sd[pfx + ".to_out.0.weight"].T + sd[pfx + ".to_out.0.bias"]
    assert torch.allclose(out[1], exp1.expand(L, -1), atol=1e-5)
    # unpadded row: plain attention
    q = (a[0] @ sd[pfx + ".to_q.weight"].T).reshape(L, 2, 64).transpose(0, 1)
    k = (a[0] @ sd[pfx + ".to_k.weight"].T).reshape(L, 2, 64).transpose(0, 1)
    vv = v[0].reshape(L, 2, 64).transpose(0, 1)
    o = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1) @ vv
    exp0 = o.transpose(0, 1).reshape(L, 128) @ sd[pfx + ".to_out.0.weight"].T + sd[pfx + ".to_out.0.bias"]
    assert torch.allclose(out[0], exp0, atol=1e-5)


def test_padded_frames_and_batch_independence(sds):
    cfg, sd = O.DecoderCfg(), sds[160]
    mu, mask, z0, _ = O.make_inputs(cfg, 3, 16, [16, 9, 12], seed=3)
    t = torch.tensor([0.1, 0.5, 0.7])
    out = O.estimator_forward(sd, cfg, z0, mask, mu, t)
    assert float((out * (1 - mask)).abs().max()) == 0.0          # exactly 0 at padded frames
    solo = O.estimator_forward(sd, cfg, z0[1:2], mask[1:2], mu[1:2], t[1:2])
    assert torch.allclose(out[1:2], solo, atol=1e-5)              # rows are independent
    z = O.euler_solve(sd, cfg, z0, mu, mask, 2)
    assert torch.equal(z * (1 - mask), z0 * (1 - mask))           # padded frames keep z0


def test_emulated_fp16_pipeline_within_tolerance(sds):
    """The storage precisions of the CUDA pipeline (fp16 everywhere, fp32 state/statistics) stay
    inside the parity bar; plain bf16 operands do not."""
    cfg, sd = O.DecoderCfg(), sds[160]
    mu, mask, z0, _ = O.make_inputs(cfg, 2, 64, [64, 50], seed=4)
    ref = O.euler_solve(sd, cfg, z0, mu, mask, 10)
    h = torch.float16
    out = O.euler_solve(sd, cfg, z0, mu, mask, 10, emu=O.Emu(operand=h, conv_out=h, attn=h, resid=h))
    ma, rl = O.parity_errors(out, ref, mask)
    assert ma <= O.TOL_MAX_ABS and rl <= O.TOL_REL_L2, (ma, rl)


def test_emulated_bf16_operands_miss_the_bar(sds):
    """Why the GEMM operands are fp16 and not the north star's bf16 (same tcgen05 kind::f16 rate): the same emulation with
    bf16 operands / storage -- fp32 accumulation, statistics and state, exactly the native pipeline's structure -- lands
    ABOVE the 1e-3 relative-L2 bar (SURVEY.md section 7.3 item 1 measured 2.0-3.4e-3), fp16 well below it.  The bf16 figure
    is printed so that it is reported next to the fp16 one."""
    cfg, sd = O.DecoderCfg(), sds[160]
    mu, mask, z0, _ = O.make_inputs(cfg, 2, 64, [64, 50], seed=4)
    ref = O.euler_solve(sd, cfg, z0, mu, mask, 10)
    out = {}
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        z = O.euler_solve(sd, cfg, z0, mu, mask, 10, emu=O.Emu(operand=dt, conv_out=dt, attn=dt, resid=dt))
        out[name] = O.parity_errors(z, ref, mask)
    print(f"emulated 10-step error vs fp32: fp16 max-abs {out['fp16'][0]:.2e} rel-L2 {out['fp16'][1]:.2e}; "
          f"bf16 max-abs {out['bf16'][0]:.2e} rel-L2 {out['bf16'][1]:.2e}")
    assert out["fp16"][1] <= O.TOL_REL_L2 < out["bf16"][1], out
    assert out["bf16"][1] >= 3 * out["fp16"][1]


def test_helpers():
    assert O.fix_len_compatibility(343) == 344 and O.fix_len_compatibility(344) == 344
    m = O.sequence_mask(torch.tensor([2, 0, 3]), 3)
    assert m.tolist() == [[True, True, False], [False, False, False], [True, True, True]]
    e = O.sinusoidal_embedding(torch.tensor([0.0]), 160)
    assert torch.allclose(e[0, :80], torch.zeros(80)) and torch.allclose(e[0, 80:], torch.ones(80))
