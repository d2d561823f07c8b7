"""Per-launch device times of ONE estimator evaluation inside a solve (CUDA events around every
launch via mtts_debug_profile_*): python tools/profile_solve.py [B T] -> table + gpurun_out/launch_table.txt"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402

# launches of one resnet + transformer stage, in order (switches as in mtts_create)
NAMES = ["conv1"] + ([] if os.environ.get("MTTS_GNA_SPLIT") == "0" else ["gnA"]) + ["conv2"]
NAMES += ["gnB", "qkv"]
NAMES += ["attn", "tail"]


def labels():
    out = ["mask", "x0"]          # per-solve prologue (the time-embedding table is cached per plan)
    per_step = []
    for s in range(6):
        per_step += [f"s{s}.{n}" for n in NAMES]
        if s in (0, 1, 4, 5):
            per_step.append(f"s{s}.levelconv")
    per_step += ["final.conv", "final.gn", "final.proj"]
    return out, per_step


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 344)
    n = 3
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
        _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0, ws[1], ws[2],
                                            B, T, 0, stream.cuda_stream))      # warm-up: builds the cached tables
        for rep in range(2):
            _lib.check(eng.lib.mtts_debug_profile_begin(eng.h, stream.cuda_stream))
            _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0, ws[1], ws[2],
                                                B, T, 0, stream.cuda_stream))
            cap = 4096
            ms, kd, fl = (C.c_float * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
            cnt = eng.lib.mtts_debug_profile_end(eng.h, cap, ms, kd, fl)
    pro, per = labels()
    lines = []
    base = len(pro) + len(per)          # skip prologue and the first step; print the second step
    tot = 0.0
    for i, lab in enumerate(per):
        j = base + i
        tf = fl[j] / (ms[j] * 1e-3) / 1e12 if fl[j] > 0 else 0.0
        lines.append(f"{lab:16s} kind={kd[j]} {ms[j]*1e3:8.1f} us  {fl[j]/1e9:8.2f} GFLOP  {tf:7.1f} TFLOP/s")
        tot += ms[j]
    lines.append(f"step total {tot*1e3:.1f} us over {len(per)} launches (B={B} T={T}); recorded {cnt} launches")
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"launch_table_B{B}_T{T}.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
