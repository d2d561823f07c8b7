"""CPU checks of the vocoder row: the oracle against the golden waveforms of the live reference,
the weight-norm folding, the ConvTranspose-as-three-tap-conv identity the native kernel relies on, and the host-side weight table."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hifigan_oracle as HO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "hifigan_golden.npz")
CASES = [("b2_t24", 2, 24, 21), ("b1_t88", 1, 88, 22), ("b3_t7", 3, 7, 23)]


def mel_input(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return -5.0 + 2.0 * torch.randn(B, 80, T, generator=g)


@pytest.fixture(scope="module")
def sd():
    return HO.make_state_dict(HO.HifiganCfg(), seed=0)


def test_oracle_equals_reference_golden(sd):
    gold = np.load(GOLD)
    cs = float(sum(float(v.double().abs().sum()) for v in sd.values()))
    assert abs(cs - float(gold["sd_checksum"])) < 1e-6 * cs           # same seeded weights as the golden run
    for name, B, T, seed in CASES:
        with torch.no_grad():
            wav = HO.generator_forward(sd, mel_input(B, T, seed))
        ref = torch.from_numpy(gold[name + ".wav"])
        assert wav.shape == ref.shape == (B, 1, 256 * T)
        assert float((wav - ref).abs().max()) <= 2e-5, name


def test_denoiser_oracle_equals_reference_golden(sd):
    gold = np.load(GOLD)
    bias = HO.denoiser_bias_spec(sd)
    assert float((bias - torch.from_numpy(gold["bias_spec"])).abs().max()) <= 1e-4
    for name in ("b2_t24", "b1_t88"):
        wav = torch.from_numpy(gold[name + ".wav"]).squeeze(1)
        for s in (0.0005, 0.05):
            out = HO.denoiser_forward(wav, bias, s)
            ref = torch.from_numpy(gold[f"{name}.denoised.{s}"])
            assert out.shape == ref.shape and float((out - ref).abs().max()) <= 2e-5, (name, s)


def test_weight_norm_folding(sd):
    wn = HO.to_weight_norm(sd)
    assert "conv_pre.weight_g" in wn and "ups.0.weight_v" in wn and "conv_pre.weight" not in wn
    assert wn["ups.0.weight_g"].shape == (512, 1, 1)                   # ConvTranspose1d: the first axis is the input channel
    f = HO.fold_weight_norm(wn)
    assert set(f) == set(sd)
    for k in sd:
        assert float((f[k] - sd[k]).abs().max()) <= 1e-6 * float(sd[k].abs().max()) + 1e-9, k
    from matcha_tts_b200 import hifigan
    f2 = hifigan.fold_weight_norm(wn)                                    # the package's own folding (host side of the boundary)
    for k in sd:
        assert torch.equal(f2[k], f[k]), k


@pytest.mark.parametrize("ci,co,u", [(8, 4, 8), (6, 3, 2)])
def test_conv_transpose_is_a_three_tap_conv(ci, co, u):
    """ConvTranspose1d(k = 2u, stride u, padding u / 2) == 3-tap conv over the input frames with u * Cout outputs: output frame
    q * u + p takes input frame q + s through kernel index j = p + pad - s * u (the packing of mtts_voc.inc)."""
    k, pad = 2 * u, u // 2
    g = torch.Generator().manual_seed(5)
    W, b, x = torch.randn(ci, co, k, generator=g), torch.randn(co, generator=g), torch.randn(2, ci, 13, generator=g)
    ref = F.conv_transpose1d(x, W, b, stride=u, padding=pad)
    Weff = torch.zeros(u * co, ci, 3)
    for p in range(u):
        for s in (-1, 0, 1):
            j = p + pad - s * u
            if 0 <= j < k:
                Weff[p * co:(p + 1) * co, :, s + 1] = W[:, :, j].t()
    y = F.conv1d(x, Weff, b.repeat(u), padding=1)
    y = y.view(2, u, co, 13).permute(0, 2, 3, 1).reshape(2, co, 13 * u)
    assert float((y - ref).abs().max()) <= 1e-5


def test_flops_per_frame():
    f = HO.flops_per_frame()
    assert 6.0e8 < f < 6.3e8          # 613 MFLOP per mel frame: 5.6 x the 10-step ODE solve


def test_generator_module_and_weight_table(libmtts, sd):
    from matcha_tts_b200 import _lib, hifigan
    gen = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
    keys = set(gen.state_dict().keys())
    assert keys == set(HO.to_weight_norm(sd).keys())                     # checkpoint form: main.py:146-147 loads strictly
    gen.load_state_dict(HO.to_weight_norm(sd), strict=True)
    gen.remove_weight_norm()
    assert set(gen.state_dict().keys()) == set(sd.keys())
    for k, v in gen.state_dict().items():
        assert float((v - sd[k]).abs().max()) <= 1e-6 * float(sd[k].abs().max()) + 1e-9, k
    h = C.c_void_p()
    _lib.check(libmtts.mtts_voc_create(C.byref(gen._vcfg), 0, C.byref(h)))
    try:
        n = libmtts.mtts_voc_num_weights(h)
        names = [libmtts.mtts_voc_weight_name(h, i).decode() for i in range(n)]
        shapes = HO.param_shapes()
        assert set(names) == set(shapes) and len(names) == len(shapes) == 156
        for i, nm in enumerate(names):
            assert libmtts.mtts_voc_weight_numel(h, i) == math.prod(shapes[nm]), nm
        assert libmtts.mtts_voc_hop_length(h) == 256
        assert libmtts.mtts_voc_workspace_bytes(h, 64, 344) < 3 << 30
        assert libmtts.mtts_voc_workspace_bytes(h, 0, 344) == 0
    finally:
        libmtts.mtts_voc_destroy(h)
    bad = _lib.MttsVocConfig.from_buffer_copy(gen._vcfg)
    bad.upsample_kernel_sizes[3] = 8
    assert libmtts.mtts_voc_create(C.byref(bad), 0, C.byref(h)) != 0
    assert b"kernel size 2u" in libmtts.mtts_last_error()
