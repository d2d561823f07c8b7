"""Parity of the native TextEncoder + duration predictor (mtts_text_* of include/mtts.h, through the C ABI) against the CPU
oracle and the golden vectors of the live reference (reference model.py:441-535), and of
MatchaTTS.synthesize end to end -- tokens in, mel out -- against the same glue run over the two oracles.

Floating point: GEMM operands and stored activations are fp16 (fp32 accumulation, fp32 LayerNorm / softmax), the oracle
fp32.  Bars: mu relative L2 <= 2e-3 and max-abs <= 1e-2 over valid tokens (measured 0.9-1.4e-3 / 2.4-4.6e-3) (the decoder rounds mu to fp16 when it stages its
first operand anyway); logw max-abs <= 1.5e-2 (measured 2.7-6.0e-3); both exactly 0 on padded tokens; x_mask exact."""
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import cfm_oracle as O  # noqa: E402
from oracle import text_encoder_oracle as TO  # noqa: E402

pytestmark = pytest.mark.gpu

MU_REL, MU_ABS, LOGW_ABS = 2e-3, 1e-2, 1.5e-2
GOLD = os.path.join(ROOT, "tests", "golden", "text_golden.npz")
CASES = [("lj_b3", 1, 3, 23, [23, 17, 5], 11), ("lj_b1", 1, 1, 40, [40], 12), ("vctk_b2", 109, 2, 19, [19, 8], 13)]


def _params(n_spks=1):
    enc = types.SimpleNamespace(encoder_type="RoPE Encoder", n_feats=80, n_channels=192, filter_channels=768, n_heads=2, n_layers=6,
                                kernel_size=3, p_dropout=0.1, prenet=True)
    dur = types.SimpleNamespace(filter_channels_dp=256, kernel_size=3, p_dropout=0.1)
    dec = types.SimpleNamespace(channels=(256, 256), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=2, num_heads=2,
                                act_fn="snakebeta")
    return enc, dur, dec


def make_encoder(n_spks=1, seed=0):
    from matcha_tts_b200 import TextEncoder
    cfg = TO.TextEncCfg(n_spks=n_spks)
    sd = TO.make_state_dict(cfg, seed)
    enc_p, dur_p, _ = _params(n_spks)
    enc = TextEncoder("RoPE Encoder", enc_p, dur_p, cfg.n_vocab, n_spks=n_spks, spk_emb_dim=64)
    enc.load_state_dict(sd, strict=True)
    return enc.cuda(), cfg, sd


def _inputs(cfg, B, T, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg.n_vocab, (B, T), generator=g)
    spks = torch.randn(B, cfg.spk_emb_dim, generator=g) if cfg.n_spks > 1 else None
    return x, torch.tensor(lengths), spks


def _check(mu, logw, mask, mu_r, logw_r, mask_r, what=""):
    assert torch.equal(mask.cpu(), mask_r)
    m = mask_r.bool()
    d = (mu.cpu().double() - mu_r.double())[m.expand_as(mu_r)]
    rel = float(d.norm() / mu_r.double()[m.expand_as(mu_r)].norm())
    ma = float(d.abs().max())
    lw = float((logw.cpu() - logw_r)[m].abs().max())
    print(f"{what}: mu rel-L2 {rel:.2e} max-abs {ma:.2e}; logw max-abs {lw:.2e}")
    assert rel <= MU_REL and ma <= MU_ABS and lw <= LOGW_ABS, (what, rel, ma, lw)
    assert float((mu.cpu() * (1 - mask_r)).abs().max()) == 0.0 and float((logw.cpu() * (1 - mask_r)).abs().max()) == 0.0


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_text_encoder_against_reference_golden(case):
    name, n_spks, B, T, lengths, seed = case
    gold = np.load(GOLD)
    enc, cfg, sd = make_encoder(n_spks)
    x, xl, spks = _inputs(cfg, B, T, lengths, seed)
    mu, logw, mask = enc(x.cuda(), xl.cuda(), None if spks is None else spks.cuda())
    _check(mu, logw, mask, torch.from_numpy(gold[name + ".mu"]), torch.from_numpy(gold[name + ".logw"]),
           torch.from_numpy(gold[name + ".mask"]), name)


@pytest.mark.parametrize("n_spks,B,T,lengths,seed", [
    (1, 1, 1, [1], 41), (1, 2, 3, [3, 1], 42), (1, 4, 130, [130, 129, 64, 2], 43), (1, 16, 97, None, 44),
    (109, 5, 77, [77, 50, 33, 9, 1], 45), (1, 2, 300, [300, 171], 46), (1, 64, 120, "ragged", 47),
])
def test_text_encoder_against_oracle(n_spks, B, T, lengths, seed):
    enc, cfg, sd = make_encoder(n_spks)
    if lengths is None:
        lengths = [T] * B
    if lengths == "ragged":
        g = torch.Generator().manual_seed(seed)
        lengths = torch.randint(20, T + 1, (B,), generator=g).tolist()
        lengths[0] = T
    x, xl, spks = _inputs(cfg, B, T, lengths, seed)
    mu_r, logw_r, mask_r = TO.text_encoder_forward(sd, cfg, x, xl, spks)
    mu, logw, mask = enc(x.cuda(), xl.cuda(), None if spks is None else spks.cuda())
    torch.cuda.synchronize()
    _check(mu, logw, mask, mu_r, logw_r, mask_r, f"n_spks={n_spks} B={B} T={T}")
    mu2, logw2, _ = enc(x.cuda(), xl.cuda(), None if spks is None else spks.cuda())
    assert torch.equal(mu, mu2) and torch.equal(logw, logw2)                 # run-to-run bit stability


def test_text_encoder_stage_trace():
    """The encoder's running activation after the prenet and after every layer (workspace buffer X1, read back with a
    growing launch limit) against the oracle's trace: a wrong stage cannot hide behind the output tolerance."""
    enc, cfg, sd = make_encoder(1)
    eng = enc._engine(torch.device("cuda", 0))
    B, T, lengths = 3, 29, [29, 14, 6]
    x, xl, _ = _inputs(cfg, B, T, lengths, 51)
    trace = {}
    TO.text_encoder_forward(sd, cfg, x, xl, None, trace=trace)
    mask = TO.sequence_mask(xl, T).unsqueeze(1).float()
    Lx = T + 2

    def x1():
        buf, ptr, n = eng.workspace(B, T)
        off = eng.lib.mtts_text_debug_buffer_offset(eng.h, B, T, b"X1")
        start = ptr - buf.data_ptr() + off
        flat = buf[start:start + B * Lx * 256 * 2].view(torch.float16).reshape(B, Lx, 256)
        return flat[:, :T, :192].permute(0, 2, 1).float().cpu()

    try:
        steps = [("prenet", 2 + 3 * 2 + 1)] + [(f"layer{i}", 2 + 7 + 7 * (i + 1)) for i in range(cfg.n_layers)]
        for name, limit in steps:
            assert eng.lib.mtts_text_debug_set_launch_limit(eng.h, limit) == 0
            enc(x.cuda(), xl.cuda())
            torch.cuda.synchronize()
            got, want = x1() * mask, trace[name] * mask
            rel = float((got.double() - want.double()).norm() / want.double().norm())
            print(f"{name}: rel-L2 {rel:.2e}")
            assert rel <= 3e-3, (name, rel)
    finally:
        eng.lib.mtts_text_debug_set_launch_limit(eng.h, -1)
    enc(x.cuda(), xl.cuda())
    assert enc.last_launch_count() == 2 + 7 + 7 * cfg.n_layers + 2 + 4


def test_padded_tokens_are_inert():
    enc, cfg, sd = make_encoder(1)
    x, xl, _ = _inputs(cfg, 3, 23, [23, 17, 5], 21)
    mu, logw, _ = enc(x.cuda(), xl.cuda())
    for b, n in enumerate([23, 17, 5]):
        mu1, logw1, _ = enc(x[b:b + 1, :n].cuda(), xl[b:b + 1].cuda())
        assert float((mu1 - mu[b:b + 1, :, :n]).abs().max()) <= 5e-3           # fp16 rows regrouped into other tiles: not bit-equal
        assert float((logw1 - logw[b:b + 1, :, :n]).abs().max()) <= 5e-3


def test_text_encoder_contract_errors():
    from matcha_tts_b200._lib import MttsError
    enc, cfg, _ = make_encoder(1)
    x = torch.zeros(2, 5, dtype=torch.long, device="cuda")
    with pytest.raises(ValueError):
        enc(x, torch.tensor([5], device="cuda"))
    with pytest.raises(ValueError):
        enc(x, torch.tensor([5, 5], device="cuda"), torch.zeros(2, 64, device="cuda"))
    with pytest.raises(RuntimeError):
        enc(x.cpu(), torch.tensor([5, 5]))
    with pytest.raises(MttsError):
        enc(torch.zeros(1, 9000, dtype=torch.long, device="cuda"), torch.tensor([9000], device="cuda"))


def test_synthesize_tokens_to_mel_natively():
    """MatchaTTS.synthesize with NO injected encoder: tokens -> native text encoder -> durations -> alignment -> mu_y ->
    native CFM decoder -> mel (reference model.py:1264-1300), against the same glue over the fp32 oracles.  Durations are
    ceil(exp(logw)): a token whose exp(logw) lies within the encoder's error of an integer may round the other way, so the
    comparison uses the native durations for both sides (and checks that they differ from the oracle's on few tokens)."""
    from matcha_tts_b200 import MatchaTTS
    from matcha_tts_b200.model import expand_by_duration
    enc_p, dur_p, dec_p = _params(1)
    tcfg = TO.TextEncCfg()
    tsd = TO.make_state_dict(tcfg, 0)
    # the seeded weights give exp(logw) ~ 1: shift the duration head so that utterances get a few frames per token
    tsd["proj_w.proj.bias"] = tsd["proj_w.proj.bias"] + 1.2
    dcfg = O.DecoderCfg()
    dsd = O.make_state_dict(dcfg, 0)
    m = MatchaTTS(tcfg.n_vocab, 1, 64, enc_p, dec_p, {"solver": "euler", "sigma_min": 1e-4}, dur_p).cuda()
    full = {"encoder." + k: v for k, v in tsd.items()}
    full.update({"decoder.estimator." + k: v for k, v in dsd.items()})
    full["mel_mean"], full["mel_std"] = torch.tensor(-5.5), torch.tensor(2.1)
    m.load_state_dict(full, strict=True)
    x, xl, _ = _inputs(tcfg, 3, 21, [21, 13, 4], 61)
    torch.manual_seed(5)
    mel, ylen, attn = m.synthesise(x.cuda(), xl.cuda(), n_timesteps=4, temperature=0.667, length_scale=1.0)
    assert mel.shape[:2] == (3, 80) and mel.shape[2] == int(ylen.max()) and attn.shape[:3] == (3, 1, 21)
    # oracle side
    mu_r, logw_r, xm = TO.text_encoder_forward(tsd, tcfg, x, xl)
    w_ceil_r = torch.ceil(torch.exp(logw_r) * xm)
    mu_n, logw_n, _ = m.encoder(x.cuda(), xl.cuda())
    w_ceil_n = torch.ceil(torch.exp(logw_n) * xm.cuda()).cpu()
    flips = int((w_ceil_r != w_ceil_n).sum())
    print(f"duration flips: {flips} of {int(xm.sum())} tokens; y_lengths {ylen.tolist()}")
    assert flips <= max(1, int(0.1 * float(xm.sum())))
    yl = torch.clamp_min(w_ceil_n.sum([1, 2]), 1).long()
    assert torch.equal(yl, ylen.cpu())
    Tm = O.fix_len_compatibility(int(yl.max()))
    y_mask = O.sequence_mask(yl, Tm).unsqueeze(1).float()
    mu_y = expand_by_duration(mu_r, w_ceil_n.squeeze(1), xm, y_mask)
    torch.manual_seed(5)
    Bz, Cz, Tz = mu_y.shape
    z0 = (torch.randn(Bz, Tz, Cz, device="cuda").transpose(1, 2) * 0.667).cpu().contiguous()      # model.py:1085 on a transposed view
    ref = O.euler_solve(dsd, dcfg, z0, mu_y.contiguous(), y_mask, 4) * 2.1 - 5.5
    n = int(yl.max())
    ma, rl = O.parity_errors(mel.cpu(), ref[:, :, :n], y_mask[:, :, :n])
    print(f"synthesize: mel max-abs {ma:.2e} rel-L2 {rl:.2e}")
    assert ma <= 2.1 * 2 * O.TOL_MAX_ABS and rl <= 2 * O.TOL_REL_L2, (ma, rl)      # encoder (3e-3 on mu) + decoder (1e-3) errors add
