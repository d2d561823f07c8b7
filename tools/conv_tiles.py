"""Per-tile stamps of ONE conv launch inside the estimator (the launch the limit stops after): when the MMA warp saw the
tile's first operands / issued its last MMA, when the epilogue saw the accumulator / finished -- for the kernels the
unit GEMM of tools/gemm_tiles.py cannot show (EPI_STATS with its GroupNorm partial sums, the dual conv1 + res_conv).
python tools/conv_tiles.py [B T] [limit ...]   limit = launches before the stop: 7 = s0.conv1, 9 = s0.conv2, 14 = s0.levelconv"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 344)
    limits = [int(a) for a in sys.argv[3:]] or [7, 9, 14]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    dec.set_chains(1)
    eng = dec._engine(dev)
    mu = torch.randn(B, 80, T, device=dev)
    z = torch.randn(B, 80, T, device=dev)
    mask = torch.ones(B, 1, T, device=dev)
    t = torch.full((B,), 0.3, device=dev)
    buf = torch.zeros(148, 64, dtype=torch.int64, device=dev)
    for _ in range(2):
        dec(z, mask, mu, t)
    torch.cuda.synchronize()
    for lim in limits:
        buf.zero_()
        _lib.check(eng.lib.mtts_debug_set_tile_timeline(eng.h, buf.data_ptr()))
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, lim))
        dec(z, mask, mu, t)
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, -1))
        _lib.check(eng.lib.mtts_debug_set_tile_timeline(eng.h, None))
        tt = buf.cpu().double()
        used = tt[:, 0] != 0
        tt = tt[used]
        rel = (tt - tt[:, 0:1]) / 1.965e3
        rel[tt == 0] = float("nan")
        print(f"B={B} T={T} launch #{lim}: {int(used.sum())} stamped CTAs; us since the CTA's first operands (mean over CTAs)")
        print(" tile | first-op  last-MMA-issued | acc-seen  epilogue-done | main loop  epilogue")
        for i in range(16):
            m = torch.nanmean(rel[:, 4 * i:4 * i + 4], dim=0)
            if torch.isnan(m[3]):
                break
            print(f" {i:4d} | {m[0]:8.2f} {m[1]:16.2f} | {m[2]:8.2f} {m[3]:14.2f} | {m[1]-m[0]:9.2f} {m[3]-m[2]:9.2f}")


if __name__ == "__main__":
    main()
