"""Steady-state cost of one tcgen05 GEMM launch: the same launch repeated back to back (PDL on), CUDA events around the
whole train.  Separates the fixed per-launch latency chain from the K-proportional main loop.
python tools/gemm_repeat.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    eng = dec._engine(dev)
    stream = torch.cuda.Stream(dev)
    reps = 200
    print("rows   C    N   taps | us/launch   GFLOP   TFLOP/s")
    shapes = ((256, 256, 1), (256, 256, 3)) if os.environ.get("MTTS_DBG") else ((256, 256, 1), (256, 256, 3), (512, 256, 3), (256, 128, 1), (256, 384, 1))
    for rows in (128, 11072, 88576):
        for Cc, N, taps in shapes:
            A = torch.randn(rows, Cc, device=dev).half()
            W = torch.randn(N, taps * Cc, device=dev).half()
            bias = torch.randn(N, device=dev)
            out = torch.empty(rows, N, dtype=torch.float16, device=dev)
            sh = (C.c_int * taps)(*([0] if taps == 1 else [-1, 0, 1]))
            with torch.cuda.stream(stream):
                def go():
                    eng.lib.mtts_debug_gemm(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), rows, Cc, N, taps, sh,
                                            stream.cuda_stream)
                for _ in range(20):
                    go()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    go()
                b.record(stream)
                b.synchronize()
            us = a.elapsed_time(b) * 1e3 / reps
            fl = 2.0 * rows * N * taps * Cc
            print(f"{rows:6d} {Cc:4d} {N:4d} {taps:3d}   | {us:8.2f}  {fl / 1e9:7.2f}  {fl / us / 1e6:8.1f}", flush=True)


if __name__ == "__main__":
    main()
