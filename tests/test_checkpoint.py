"""Checkpoint / wire formats on CPU: the Lightning-checkpoint loader follows reference
main.py:94-121 (unwrap "state_dict", strip "model.", strict load, eval), the flat weight file round-trips the
estimator in libmtts' table order, and the mel writer stores (n_feats, T) float32."""
import ctypes as C
import os
import types

import numpy as np
import pytest
import torch

from matcha_tts_b200 import Decoder, MatchaTTS, checkpoint as K
from oracle import cfm_oracle as O


def _matcha(n_spks=1):
    enc_p = types.SimpleNamespace(n_feats=80)
    dp = types.SimpleNamespace(channels=(256, 256), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=2,
                               num_heads=2, act_fn="snakebeta")
    return MatchaTTS(178, n_spks, 64, enc_p, dp, {"solver": "euler"}, encoder=torch.nn.Identity())


def test_lightning_checkpoint_is_unwrapped_stripped_and_loaded_strictly(tmp_path):
    src = _matcha()
    sd = O.make_state_dict(O.DecoderCfg(), 3)
    src.decoder.estimator.load_state_dict(sd)
    src.mel_mean.fill_(-5.52)
    src.mel_std.fill_(2.06)
    ckpt = {"epoch": 7, "state_dict": {"model." + k: v.clone() for k, v in src.state_dict().items()}}
    path = os.path.join(tmp_path, "last.ckpt")
    torch.save(ckpt, path)
    for arg in (path, ckpt, ckpt["state_dict"], K.strip_lightning_prefix(ckpt["state_dict"])):
        m = _matcha().train()
        out = K.load_lightning_checkpoint(m, arg)
        assert out is m and not m.training                                   # main.py:120 model.eval()
        assert float(m.mel_mean) == pytest.approx(-5.52) and float(m.mel_std) == pytest.approx(2.06)
        for k, v in sd.items():
            assert torch.equal(m.decoder.estimator.state_dict()[k], v), k
    bad = dict(ckpt["state_dict"])
    bad.pop("model.decoder.estimator.final_proj.bias")
    with pytest.raises(RuntimeError):                                        # strict like main.py:114-118
        K.load_lightning_checkpoint(_matcha(), bad)
    K.load_lightning_checkpoint(_matcha(), bad, strict=False)


def test_weight_file_round_trip_in_table_order(tmp_path, libmtts):
    from matcha_tts_b200 import _lib
    dec = Decoder(in_channels=224, out_channels=80, channels=(256, 256), num_heads=2, num_mid_blocks=2)
    dec.load_state_dict(O.make_state_dict(O.DecoderCfg(in_channels=224), 5))
    path = os.path.join(tmp_path, "estimator.mttsw")
    nbytes = K.save_weight_file(dec, path)
    assert nbytes == os.path.getsize(path)
    table = K.load_weight_file(path)
    cfg = _lib.MttsConfig(224, 80, 256, 2, 64, 2)
    h = C.c_void_p()
    _lib.check(libmtts.mtts_create(C.byref(cfg), 0, C.byref(h)))
    want = [(libmtts.mtts_weight_name(h, i).decode(), libmtts.mtts_weight_numel(h, i)) for i in range(libmtts.mtts_num_weights(h))]
    libmtts.mtts_destroy(h)
    assert [(k, v.numel()) for k, v in table.items()] == want               # exactly what mtts_load_weight expects, in order
    other = Decoder(in_channels=224, out_channels=80, channels=(256, 256), num_heads=2, num_mid_blocks=2)
    K.load_weight_file_into(other, path)
    for k, v in dec.state_dict().items():
        assert torch.equal(other.state_dict()[k], v), k
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        K.load_weight_file(path)


def test_mel_writer(tmp_path):
    mel = torch.randn(1, 80, 37)
    path = os.path.join(tmp_path, "mel.npy")
    K.save_mel_npy(mel, path)
    back = np.load(path)
    assert back.dtype == np.float32 and back.shape == (80, 37) and np.array_equal(back, mel[0].numpy())
    with pytest.raises(ValueError):
        K.save_mel_npy(torch.randn(2, 80, 5), path)
