"""Per-tile timeline of one persistent GEMM launch at full occupancy (mtts_debug_set_tile_timeline): for CTA 0 and the mean
over CTAs, when the MMA warp saw a tile's first operands / issued its last MMA and when the epilogue saw / finished the
accumulator -- shows whether epilogue(i) overlaps main loop(i+1).   python tools/gemm_tiles.py [rows C N taps]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    rows, Cc, N, taps = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (88576, 256, 256, 3)
    dev = torch.device("cuda", 0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    eng = dec._engine(dev)
    stream = torch.cuda.Stream(dev)
    A = torch.randn(rows, Cc, device=dev).half()
    W = torch.randn(N, taps * Cc, device=dev).half()
    bias = torch.randn(N, device=dev)
    out = torch.empty(rows, N, dtype=torch.float16, device=dev)
    sh = (C.c_int * taps)(*([0] if taps == 1 else [-1, 0, 1]))
    buf = torch.zeros(148, 64, dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        def go():
            eng.lib.mtts_debug_gemm(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), rows, Cc, N, taps, sh,
                                    stream.cuda_stream)
        for _ in range(5):
            go()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_tile_timeline(eng.h, buf.data_ptr()))
        go()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_tile_timeline(eng.h, None))
    t = buf.cpu().double()
    used = t[:, 0] != 0
    t = t[used]
    t0 = t[:, 0:1]
    rel = (t - t0) / 1.9e3      # us at ~1.9 GHz
    rel[t == 0] = float("nan")
    print(f"rows={rows} C={Cc} N={N} taps={taps}: {int(used.sum())} CTAs; us since the CTA's first operands arrived")
    print(" tile | MMA first-op  MMA last-issue | EPI acc-seen  EPI done   (CTA 0)      |  mean over CTAs: first-op  last-issue  acc-seen  done")
    for i in range(16):
        if torch.isnan(rel[0, 4 * i + 3]):
            break
        m = torch.nanmean(rel[:, 4 * i:4 * i + 4], dim=0)
        print(f" {i:4d} | {rel[0, 4*i]:11.2f} {rel[0, 4*i+1]:14.2f} | {rel[0, 4*i+2]:11.2f} {rel[0, 4*i+3]:9.2f}                |"
              f" {m[0]:24.2f} {m[1]:11.2f} {m[2]:9.2f} {m[3]:6.2f}")


if __name__ == "__main__":
    main()
