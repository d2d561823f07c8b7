"""Per-launch device times of one native HiFi-GAN generator forward (B200 only).

    python tools/voc_profile.py [--B 64] [--T 344] [--reps 3]

Brackets every launch with a CUDA-event pair (mtts_voc_debug_profile_begin / _end) and prints milliseconds, algorithmic
FLOPs and TFLOP/s per launch plus the per-level sums; then times the whole forward (CUDA graph) with events."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as HO  # noqa: E402


def labels(cfg):
    out = ["pack_mel", "conv_pre"]
    nk, nd = len(cfg.resblock_kernel_sizes), len(cfg.resblock_dilation_sizes[0])
    for i in range(len(cfg.upsample_rates)):
        out.append(f"L{i + 1}.ups")
        for j in range(nk):
            for m in range(nd):
                k, d = cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j][m]
                out += [f"L{i + 1}.rb{j}.c1 k{k} d{d}", f"L{i + 1}.rb{j}.c2 k{k} d1"]
    out.append("conv_post")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--T", type=int, default=344)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from matcha_tts_b200 import hifigan
    cfg = HO.HifiganCfg()
    gen = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
    gen.load_state_dict(HO.to_weight_norm(HO.make_state_dict(cfg, 0)), strict=True)
    gen = gen.cuda()
    mel = (-5.0 + 2.0 * torch.randn(a.B, 80, a.T)).cuda()
    eng = gen._engine(torch.device("cuda", 0))
    lib = eng.lib
    gen.use_cuda_graph = False
    for _ in range(2):
        gen(mel)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    best = None
    for _ in range(a.reps):
        lib.mtts_voc_debug_profile_begin(eng.h, stream)
        gen(mel)
        n = 128
        ms, kind, fl = (C.c_float * n)(), (C.c_int * n)(), (C.c_double * n)()
        cnt = lib.mtts_voc_debug_profile_end(eng.h, n, ms, kind, fl)
        rows = [(ms[i], fl[i]) for i in range(cnt)]
        if best is None or sum(r[0] for r in rows) < sum(r[0] for r in best):
            best = rows
    lab = labels(cfg)
    tot_ms = sum(r[0] for r in best)
    tot_fl = sum(r[1] for r in best)
    lev = {}
    for l, (m, f) in zip(lab, best):
        print(f"{l:24s} {m * 1e3:9.1f} us  {f / 1e9:9.2f} GFLOP  {f / m / 1e9 if m > 0 else 0:8.1f} TFLOP/s")
        key = l.split(".")[0]
        lev[key] = (lev.get(key, (0, 0))[0] + m, lev.get(key, (0, 0))[1] + f)
    for k, (m, f) in lev.items():
        print(f"sum {k:12s} {m:8.3f} ms  {f / 1e12:7.3f} TFLOP  {f / m / 1e9:8.1f} TFLOP/s")
    print(f"launch sum {tot_ms:.3f} ms over {len(best)} launches, {tot_fl / 1e12:.3f} TFLOP -> {tot_fl / tot_ms / 1e9:.1f} TFLOP/s "
          f"(B={a.B} T={a.T}, {a.B * a.T / tot_ms * 1e3 / 1e6:.3f} M mel-frames/s)")
    gen.use_cuda_graph = True
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            gen(mel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gen(mel)
        e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5
    print(f"forward (graph, incl. mel staging + wav clone): {t:.3f} ms -> {a.B * a.T / t * 1e3 / 1e6:.3f} M mel-frames/s, "
          f"{a.B * a.T * 256 / 22050 / (t / 1e3):.0f} x real time at 22.05 kHz")


if __name__ == "__main__":
    main()
