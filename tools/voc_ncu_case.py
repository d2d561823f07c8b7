"""One eager forward of the native HiFi-GAN generator, for `ncu` (B200 only): no warm-up, no graph, 79 launches.

    python tools/voc_ncu_case.py [--B 64] [--T 344]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as HO  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--T", type=int, default=344)
a = ap.parse_args()
from matcha_tts_b200 import hifigan  # noqa: E402
gen = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
gen.load_state_dict(HO.to_weight_norm(HO.make_state_dict(HO.HifiganCfg(), 0)), strict=True)
gen = gen.cuda()
gen.use_cuda_graph = False
mel = (-5.0 + 2.0 * torch.randn(a.B, 80, a.T)).cuda()
wav = gen(mel)
torch.cuda.synchronize()
print("ok", tuple(wav.shape), float(wav.abs().max()), gen.last_launch_count())
