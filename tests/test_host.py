"""Host-side logic and the C-ABI boundary on CPU: the library loads and exports every symbol
include/mtts.h declares, its weight table equals the reference state-dict layout, and the
drop-in modules keep the reference's constructor / error behaviour.  No compute calls here."""
import ctypes as C
import os
import re
import types

import pytest
import torch

from oracle import cfm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists("/root/reference/model.py")


def _ref():
    import sys
    sys.dont_write_bytecode = True
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import model as ref
    return ref


def test_header_symbols_exported(libmtts):
    from matcha_tts_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mtts.h")).read()
    declared = set(re.findall(r"\b(mtts_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no prototypes found in mtts.h"
    for name in declared:
        assert hasattr(libmtts, name), f"libmtts.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes table and mtts.h disagree"
    assert b"sm_100a" in libmtts.mtts_version()


def test_lanes_and_chains_switches_validate_their_argument(libmtts):
    """mtts_set_lanes / mtts_set_chains (include/mtts.h): range-checked, no GPU needed."""
    from matcha_tts_b200 import _lib
    cfg = _lib.MttsConfig(160, 80, 256, 2, 64, 2)
    h = C.c_void_p()
    _lib.check(libmtts.mtts_create(C.byref(cfg), 0, C.byref(h)))
    try:
        for n in (1, 4, 16):
            assert libmtts.mtts_set_lanes(h, n) == 0
        for n in (0, -1, 17):
            assert libmtts.mtts_set_lanes(h, n) < 0
        assert libmtts.mtts_set_lanes(None, 2) < 0
        assert libmtts.mtts_set_chains(h, 9) < 0 and libmtts.mtts_set_chains(h, 1) == 0
        # the SM share of a persistent launch (148 SMs assumed without a device): all SMs for one solve at a time; with n
        # lanes the grid with the least wave-quantisation waste inside [148 / n, 148 * 5/4 / n]
        grid = libmtts.mtts_debug_lane_grid
        _lib.check(libmtts.mtts_set_lanes(h, 1))
        assert [grid(h, t, 1) for t in (0, 3, 148, 173, 692)] == [0, 3, 148, 148, 148] and grid(h, 384, 2) == 296
        _lib.check(libmtts.mtts_set_lanes(h, 4))
        assert grid(h, 173, 1) == 44 and grid(h, 87, 1) == 44          # 4 and 2 full waves (43 CTAs would need 5 and 3)
        assert grid(h, 30, 1) == 30 and grid(h, 37, 1) == 37            # fewer tiles than the share: one CTA each
        assert grid(h, 257, 1) == 43 and grid(h, 129, 1) == 43          # B16 x T2048: 6 and 3 full waves
        assert 74 <= grid(h, 384, 2) <= 93                              # two CTAs per SM: a share of 296
        _lib.check(libmtts.mtts_set_lanes(h, 5))
        assert grid(h, 173, 1) == 35 and 30 <= grid(h, 87, 1) <= 37
        for lanes in (2, 3, 6, 8, 16):
            _lib.check(libmtts.mtts_set_lanes(h, lanes))
            for tiles in (1, 44, 87, 173, 346, 692, 5000):
                g = grid(h, tiles, 1)
                lo, hi = -(-148 // lanes), -(-(148 * 5 // 4) // lanes)
                assert (g == tiles and tiles <= lo) or lo <= g <= hi, (lanes, tiles, g)
        assert grid(None, 4, 1) < 0 and grid(h, 4, 3) < 0
    finally:
        libmtts.mtts_destroy(h)


@pytest.mark.parametrize("cin", [160, 224])
def test_weight_table_matches_state_dict(libmtts, cin):
    from matcha_tts_b200 import _lib
    from matcha_tts_b200.model import estimator_param_spec
    cfg = _lib.MttsConfig(cin, 80, 256, 2, 64, 2)
    h = C.c_void_p()
    _lib.check(libmtts.mtts_create(C.byref(cfg), 0, C.byref(h)))
    n = libmtts.mtts_num_weights(h)
    table = [(libmtts.mtts_weight_name(h, i).decode(), libmtts.mtts_weight_numel(h, i)) for i in range(n)]
    spec = estimator_param_spec(cin, 80, 256, 2, 64, 2)
    assert spec == O.state_dict_spec(O.DecoderCfg(in_channels=cin))      # package mirror == oracle == App. B
    want = [("@time_freqs", cin // 2)] + [(k, int(torch.Size(s).numel())) for k, s in spec]
    assert table == want
    assert libmtts.mtts_weight_arena_bytes(h) > 2 * sum(x[1] for x in want if "block.0.weight" in x[0])
    # workspace sizing: valid and invalid shapes
    assert libmtts.mtts_workspace_bytes(h, 64, 344) > 0
    assert libmtts.mtts_workspace_bytes(h, 1, 2) > 0
    assert libmtts.mtts_workspace_bytes(h, 4, 33) > libmtts.mtts_workspace_bytes(h, 4, 32)   # odd T: one more guard row (reference crop, model.py:1027)
    assert libmtts.mtts_workspace_bytes(h, 4, 0) == 0
    assert libmtts.mtts_workspace_bytes(h, 0, 32) == 0
    assert libmtts.mtts_debug_buffer_offset(h, 2, 32, 0, b"skip0") > 0
    assert libmtts.mtts_debug_buffer_offset(h, 2, 32, 0, b"nope") == -1
    # call-order errors are reported, not ignored
    assert libmtts.mtts_load_weight(h, 0, C.c_void_p(16), cin // 2, None) == -3
    assert b"arena" in libmtts.mtts_last_error()
    libmtts.mtts_destroy(h)


def test_unsupported_config_is_an_error(libmtts):
    from matcha_tts_b200 import _lib
    h = C.c_void_p()
    bad = _lib.MttsConfig(160, 80, 512, 2, 64, 2)
    assert libmtts.mtts_create(C.byref(bad), 0, C.byref(h)) == -1
    assert b"unsupported" in libmtts.mtts_last_error()
    with pytest.raises(_lib.MttsError):
        _lib.check(-1)


def test_decoder_state_dict_layout():
    from matcha_tts_b200 import Decoder
    dec = Decoder(in_channels=160, out_channels=80, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                  n_blocks=1, num_mid_blocks=2, num_heads=2, act_fn="snakebeta")
    sd = dec.state_dict()
    spec = O.state_dict_spec(O.DecoderCfg())
    assert sorted(sd.keys()) == sorted(k for k, _ in spec)
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in spec)
    assert sum(v.numel() for v in sd.values()) == 11_008_848
    dec.load_state_dict(O.make_state_dict(O.DecoderCfg(), 0), strict=True)         # reference-layout weights load
    with pytest.raises(RuntimeError):                                              # no CPU path
        dec(torch.zeros(1, 80, 8), torch.ones(1, 1, 8), torch.zeros(1, 80, 8), torch.zeros(1))


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_decoder_state_dict_equals_reference():
    from matcha_tts_b200 import Decoder
    ref = _ref()
    for cin in (160, 224):
        kw = dict(in_channels=cin, out_channels=80, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                  n_blocks=1, num_mid_blocks=2, num_heads=2, act_fn="snakebeta")
        a, b = Decoder(**kw).state_dict(), ref.Decoder(**kw).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
        Decoder(**kw).load_state_dict(b, strict=True)


def test_cfm_constructor_contract():
    from matcha_tts_b200 import CFM, Decoder
    with pytest.raises(ValueError):
        CFM(80, {"solver": "euler"}, estimator=None)                # model.py:1131-1132
    dec = Decoder(160, 80, num_heads=2)
    cfm = CFM(80, {"solver": "rk4"}, estimator=dec)
    with pytest.raises(NotImplementedError):                         # model.py:1107
        cfm.solve_from(torch.zeros(1, 80, 8), torch.zeros(1, 80, 8), torch.ones(1, 1, 8), 2)
    assert CFM(80, {}, estimator=dec).solver == "euler" and CFM(80, {}, estimator=dec).sigma_min == 1e-4
    with pytest.raises(NotImplementedError):
        Decoder(160, 80, channels=(256, 512))


def test_host_helpers():
    from matcha_tts_b200 import denormalize, fix_len_compatibility, generate_path, sequence_mask
    assert fix_len_compatibility(343) == 344 and fix_len_compatibility(1) == 4
    lens = torch.tensor([3, 1])
    assert sequence_mask(lens, 4).tolist() == [[True, True, True, False], [True, False, False, False]]
    dur = torch.tensor([[2.0, 1.0, 0.0], [1.0, 3.0, 0.0]])
    mask = torch.ones(2, 3, 4)
    p = generate_path(dur, mask)
    assert p[0].tolist() == [[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]]
    assert p[1].tolist() == [[1, 0, 0, 0], [0, 1, 1, 1], [0, 0, 0, 0]]
    x = torch.ones(1, 2, 3)
    assert torch.allclose(denormalize(x, -5.5, 2.0), torch.full((1, 2, 3), -3.5))
    assert torch.allclose(denormalize(x, torch.tensor(1.0), torch.tensor(3.0)), torch.full((1, 2, 3), 4.0))


def test_expand_by_duration_equals_dense_alignment_product():
    """synthesize expands mu by a gather; the reference multiplies by the dense 0/1 path (model.py:1284-1288):
    bit-identical, including zero-duration tokens, masked tokens, padded frames and an all-zero-duration row."""
    from matcha_tts_b200 import generate_path, sequence_mask
    from matcha_tts_b200.model import expand_by_duration, fix_len_compatibility
    g = torch.Generator().manual_seed(3)
    for trial in range(6):
        b, tx, nf = 4, 9, 5
        x_len = torch.tensor([9, 6, 1, 4])
        x_mask = sequence_mask(x_len, tx).unsqueeze(1).float()
        w_ceil = torch.randint(0, 5, (b, 1, tx), generator=g).float() * x_mask
        if trial == 0:
            w_ceil[2] = 0.0                                           # y_length clamps to 1, no token covers frame 0
        mu = torch.randn(b, nf, tx, generator=g)
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        t_y = fix_len_compatibility(int(y_lengths.max()))
        y_mask = sequence_mask(y_lengths, t_y).unsqueeze(1).float()
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1))
        dense = torch.matmul(attn.transpose(1, 2), mu.transpose(1, 2)).transpose(1, 2)
        got = expand_by_duration(mu, w_ceil.squeeze(1), x_mask, y_mask)
        assert torch.equal(got, dense) and got.stride() == dense.stride()    # same memory order: randn_like(mu_y) draws alike


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_host_helpers_equal_reference():
    from matcha_tts_b200 import fix_len_compatibility, generate_path, sequence_mask
    ref = _ref()
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        b, tx = 3, 7
        dur = torch.randint(0, 5, (b, tx), generator=g).float()
        ty = int(dur.sum(1).max()) + 2
        ylen = dur.sum(1).long()
        xm = torch.ones(b, tx)
        ym = ref.sequence_mask(ylen, ty).float()
        am = xm.unsqueeze(-1) * ym.unsqueeze(1)
        assert torch.equal(generate_path(dur, am), ref.generate_path(dur, am))
        assert torch.equal(sequence_mask(ylen, ty), ref.sequence_mask(ylen, ty))
    for n in (1, 4, 5, 343, 344):
        assert fix_len_compatibility(n) == ref.fix_len_compatibility(torch.tensor(n))


def test_matcha_facade_signature():
    import inspect
    from matcha_tts_b200 import MatchaTTS
    sig = inspect.signature(MatchaTTS.synthesize)
    assert list(sig.parameters) == ["self", "x", "x_lengths", "n_timesteps", "temperature", "spks", "length_scale"]
    assert MatchaTTS.synthesise is MatchaTTS.synthesize
    enc = types.SimpleNamespace(n_feats=80)
    dp = types.SimpleNamespace(channels=(256, 256), dropout=0.05, attention_head_dim=64, n_blocks=1,
                               num_mid_blocks=2, num_heads=2, act_fn="snakebeta")
    m = MatchaTTS(178, 1, 64, enc, dp, {"solver": "euler", "sigma_min": 1e-4})
    keys = list(m.state_dict().keys())
    assert keys[0] == "mel_mean" and keys[1] == "mel_std" and keys[2].startswith("decoder.estimator.time_mlp")
    with pytest.raises(RuntimeError):
        m.synthesize(torch.zeros(1, 4, dtype=torch.long), torch.tensor([4]), 2)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU port of the reference path) on a tiny shape: one JSON line with the
    contract's keys, and nothing else on stdout."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "2", "--frames", "32", "--n-timesteps", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "mel-frames/s" and d["higher_is_better"] is True
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "model.py"))    # tools/stage_reference.sh
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["config"]["config1_cpu"]["batch"] == 1 and d["config"]["config1_cpu"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("n_spks", [1, 109])
def test_text_weight_table_matches_state_dict(libmtts, n_spks):
    """The text encoder's weight table (mtts_text_weight_name / _numel) lists the reference TextEncoder's state-dict keys
    (oracle.param_shapes == the live reference's, tests/test_text_oracle.py) in a fixed order, plus the host-derived RoPE
    frequencies; the package's spec mirrors it."""
    from matcha_tts_b200 import _lib
    from matcha_tts_b200.text_encoder import text_encoder_param_spec
    from oracle import text_encoder_oracle as TO
    cfg = TO.TextEncCfg(n_spks=n_spks)
    tc = _lib.MttsTextConfig(cfg.n_vocab, cfg.n_feats, cfg.n_channels, cfg.filter_channels, cfg.n_heads, cfg.n_layers, cfg.kernel_size, 1,
                             cfg.filter_channels_dp, cfg.kernel_size_dp, n_spks, cfg.spk_emb_dim)
    h = C.c_void_p()
    _lib.check(libmtts.mtts_text_create(C.byref(tc), 0, C.byref(h)))
    table = [(libmtts.mtts_text_weight_name(h, i).decode(), libmtts.mtts_text_weight_numel(h, i)) for i in range(libmtts.mtts_text_num_weights(h))]
    shapes = TO.param_shapes(cfg)
    spec = text_encoder_param_spec(cfg.n_vocab, cfg.n_channels, cfg.filter_channels, cfg.n_layers, cfg.kernel_size, True, cfg.n_feats,
                                   cfg.filter_channels_dp, cfg.kernel_size_dp, cfg.width)
    assert dict(spec) == shapes and [k for k, _ in spec] == list(shapes)
    half = (cfg.width // cfg.n_heads) // 4
    want = [(k, int(torch.Size(s).numel())) for k, s in spec]
    want.insert(1, ("@rope_theta", half))
    assert table == want
    assert libmtts.mtts_text_workspace_bytes(h, 64, 120) > 0 and libmtts.mtts_text_workspace_bytes(h, 0, 5) == 0
    assert libmtts.mtts_text_load_weight(h, 0, C.c_void_p(16), cfg.n_vocab * cfg.n_channels, None) == -3      # arena not set
    libmtts.mtts_text_destroy(h)
    bad = _lib.MttsTextConfig(178, 80, 200, 768, 2, 6, 3, 1, 256, 3, 1, 64)                                      # width not a multiple of 64
    h2 = C.c_void_p()
    assert libmtts.mtts_text_create(C.byref(bad), 0, C.byref(h2)) == -1 and b"multiples of 64" in libmtts.mtts_last_error()


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_matcha_state_dict_equals_reference_and_loads_a_lightning_checkpoint():
    """The whole model tree -- native text encoder, estimator, buffers -- has the reference MatchaTTS's keys and shapes, so
    a reference training checkpoint loads with strict=True (reference main.py:94-121)."""
    from matcha_tts_b200 import MatchaTTS, load_lightning_checkpoint
    ref = _ref()
    enc = types.SimpleNamespace(encoder_type="RoPE Encoder", n_feats=80, n_channels=192, filter_channels=768, n_heads=2, n_layers=6,
                                kernel_size=3, p_dropout=0.1, prenet=True)
    dec = types.SimpleNamespace(channels=(256, 256), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=2, num_heads=2,
                                act_fn="snakebeta")
    dur = types.SimpleNamespace(filter_channels_dp=256, kernel_size=3, p_dropout=0.1)
    for n_spks in (1, 109):
        theirs = ref.MatchaTTS(178, n_spks, 64, enc, dec, {"solver": "euler", "sigma_min": 1e-4}, dur)
        mine = MatchaTTS(178, n_spks, 64, enc, dec, {"solver": "euler", "sigma_min": 1e-4}, dur)
        a = {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        assert a == b
        ckpt = {"state_dict": {"model." + k: v for k, v in theirs.state_dict().items()}}
        load_lightning_checkpoint(mine, ckpt)                                  # strict
        for k, v in theirs.state_dict().items():
            assert torch.equal(mine.state_dict()[k], v), k
