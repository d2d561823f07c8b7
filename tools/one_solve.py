"""One eager (non-graph) solve of the bench workload, for `ncu` launch lists / captures:
python tools/one_solve.py [B T n_steps]   (prints nothing but a completion line)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 344, 2)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    eng = dec._engine(dev)
    mu = torch.randn(B, 80, T, device=dev)
    z = torch.randn(B, 80, T, device=dev)
    mask = torch.ones(B, 1, T, device=dev)
    stream = torch.cuda.Stream(dev)
    ws = eng.workspace(B, T)
    with torch.cuda.stream(stream):
        for _ in range(2):
            _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0, ws[1], ws[2],
                                                B, T, 0, stream.cuda_stream))
    torch.cuda.synchronize()
    print(f"one_solve done: B={B} T={T} n={n} launches={eng.launch_count()} finite={bool(torch.isfinite(z).all())}")


if __name__ == "__main__":
    main()
