"""Golden vectors of the vendored HiFi-GAN Generator + Denoiser, from the LIVE reference.

    python tests/golden/make_hifigan_golden.py     (build container only: imports /root/reference/hifigan, read-only)

The oracle's seeded state-dict (oracle.hifigan_oracle.make_state_dict), in its weight-normed form, is loaded into the reference
Generator with load_state_dict(strict=True) -- which pins the checkpoint's key names and shapes -- and the reference's
waveforms on seeded mels, plus the reference Denoiser's output on them, are stored in tests/golden/hifigan_golden.npz.
Weights are not stored; a checksum of the state-dict is.  (hifigan/xutils.py imports matplotlib for a plotting helper the
path never calls; it is not in this image, so an empty stand-in module is registered before the import.)
"""
import contextlib
import io
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

try:
    import matplotlib  # noqa: F401
except ImportError:
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    m.pylab = types.ModuleType("matplotlib.pylab")
    sys.modules["matplotlib"], sys.modules["matplotlib.pylab"] = m, m.pylab

from hifigan.config import v1                           # noqa: E402  (the reference itself)
from hifigan.denoiser import Denoiser                   # noqa: E402
from hifigan.env import AttrDict                        # noqa: E402
from hifigan.models import Generator                    # noqa: E402
from oracle import hifigan_oracle as HO                 # noqa: E402

# name, B, T, seed
CASES = [("b2_t24", 2, 24, 21), ("b1_t88", 1, 88, 22), ("b3_t7", 3, 7, 23)]
STRENGTHS = (0.0005, 0.05)


def checksum(sd):
    return float(sum(float(v.double().abs().sum()) for v in sd.values()))


def mel_input(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return -5.0 + 2.0 * torch.randn(B, 80, T, generator=g)          # log-mel range of the LJSpeech statistics (main.py:74)


def main():
    cfg = HO.HifiganCfg()
    sd = HO.make_state_dict(cfg, seed=0)
    wn = HO.to_weight_norm(sd)
    gen = Generator(AttrDict(v1))
    gen.load_state_dict(wn, strict=True)
    gen.eval()
    folded = HO.fold_weight_norm(wn)
    assert set(folded) == set(sd)
    for k in sd:
        assert float((folded[k] - sd[k]).abs().max()) <= 1e-6 * float(sd[k].abs().max()) + 1e-9, k
    with contextlib.redirect_stdout(io.StringIO()):
        gen.remove_weight_norm()                                     # main.py:149
    for k, v in gen.state_dict().items():
        assert float((v - sd[k]).abs().max()) <= 1e-6, k
    den = Denoiser(gen, mode="zeros")
    out = {"sd_checksum": np.float64(checksum(sd))}
    bias_o = HO.denoiser_bias_spec(sd, cfg)
    err = float((den.bias_spec - bias_o).abs().max())
    assert err <= 1e-5 * float(bias_o.abs().max()), err
    print(f"bias_spec: oracle vs reference max-abs {err:.2e} (max {float(bias_o.abs().max()):.3f})")
    out["bias_spec"] = den.bias_spec.numpy()
    for name, B, T, seed in CASES:
        mel = mel_input(B, T, seed)
        with torch.no_grad():
            wav = gen(mel)
            wav_o = HO.generator_forward(sd, mel, cfg)
        err = float((wav - wav_o).abs().max())
        assert err <= 2e-5, (name, err)
        print(f"{name}: oracle vs reference wav max-abs {err:.2e}  (rms {float(wav.pow(2).mean().sqrt()):.3f}, |max| {float(wav.abs().max()):.3f})")
        out[name + ".wav"] = wav.numpy()
        if T * cfg.hop > 1024:                                        # reflect padding of the centred STFT needs n > n_fft / 2
            for s in STRENGTHS:
                dn = den(wav.squeeze(1), strength=s)
                dn_o = HO.denoiser_forward(wav.squeeze(1), bias_o, s)
                err = float((dn - dn_o).abs().max())
                assert err <= 2e-5, (name, s, err)
                print(f"{name}: oracle vs reference denoised (strength {s}) max-abs {err:.2e}, change {float((dn - wav.squeeze(1)[:, :dn.shape[1]]).abs().max()):.3e}")
                out[f"{name}.denoised.{s}"] = dn.numpy()
    np.savez_compressed(os.path.join(HERE, "hifigan_golden.npz"), **out)
    print("wrote hifigan_golden.npz:", {k: v.shape for k, v in out.items() if v.ndim})


if __name__ == "__main__":
    main()
