"""In-context marginal device time of every launch of ONE estimator evaluation at real clocks:
T(k) = time of mtts_estimator_forward stopped after k launches (mtts_debug_set_launch_limit), averaged
over reps with CUDA events; marginal(k) = T(k) - T(k-1).  Unlike per-launch event pairs this keeps
programmatic dependent launch overlap and adds no per-kernel event overhead.
python tools/marginal_times.py [B T reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from matcha_tts_b200 import Decoder, _lib  # noqa: E402
from profile_solve import labels  # noqa: E402


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 344)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    eng = dec._engine(dev)
    mu = torch.randn(B, 80, T, device=dev)
    x = torch.randn(B, 80, T, device=dev)
    out = torch.empty_like(x)
    mask = torch.ones(B, 1, T, device=dev)
    t = torch.full((B,), 0.3, device=dev)
    stream = torch.cuda.Stream(dev)
    ws = eng.workspace(B, T)
    pro, per = labels()
    n_pro = 6                       # estimator_forward: sinus, lin1, lin2, lin6, mask, x0
    total = n_pro + len(per)

    def run(limit):
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, limit))
        _lib.check(eng.lib.mtts_estimator_forward(eng.h, x.data_ptr(), mu.data_ptr(), mask.data_ptr(), t.data_ptr(), None,
                                                  out.data_ptr(), ws[1], ws[2], B, T, stream.cuda_stream))

    times = []
    with torch.cuda.stream(stream):
        for _ in range(3):
            run(-1)
        torch.cuda.synchronize()
        for k in range(n_pro, total + 1):
            best = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                run(k)
                b.record(stream)
                b.synchronize()
                best.append(a.elapsed_time(b) * 1e3)
            best.sort()
            times.append(sum(best[:max(1, reps // 2)]) / max(1, reps // 2))   # mean of the faster half
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, -1))
    lines = [f"prologue ({n_pro} launches): {times[0]:8.1f} us"]
    by = {}
    for i, lab in enumerate(per):
        d = times[i + 1] - times[i]
        lines.append(f"{lab:16s} {d:8.2f} us   (cumulative {times[i + 1]:9.1f})")
        k = lab.split(".")[1]
        by[k] = by.get(k, 0.0) + d
    lines.append(f"estimator total {times[-1] - times[0]:.1f} us over {len(per)} launches (B={B} T={T}, eager launches, PDL on)")
    lines.append("by layer type: " + ", ".join(f"{k}={v:.1f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"marginal_B{B}_T{T}.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
