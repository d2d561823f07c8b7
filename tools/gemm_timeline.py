"""In-kernel timeline of every GEMM launch of one estimator evaluation (mtts_debug_set_timeline):
per launch, the mean over CTAs of the phase durations (SM cycles -> us at the measured clock) and the
gap between one kernel's last CTA exit and the next GEMM's first dependency-wait release.
python tools/gemm_timeline.py [B T]"""
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
    _, per = labels()
    gemm_labels = [l for l in per if l.split(".")[1] not in ("gnA", "gnB", "attn", "gn", "tail")]
    nl = len(gemm_labels)
    buf = torch.zeros(nl, 148, 16, dtype=torch.int64, device=dev)

    def run():
        _lib.check(eng.lib.mtts_estimator_forward(eng.h, x.data_ptr(), mu.data_ptr(), mask.data_ptr(), t.data_ptr(), None,
                                                  out.data_ptr(), ws[1], ws[2], B, T, stream.cuda_stream))
    with torch.cuda.stream(stream):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_timeline(eng.h, buf.data_ptr(), nl))
        run()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_timeline(eng.h, None, 0))
    tl = buf.cpu()
    lines = ["launch            ctas  setup  depwait  1st-op  mma-issue  acc-ready  epi-done  exit   | span_us  gap_to_next_us   (phase columns: mean us since CTA entry)"]
    prev_exit = None
    for i, lab in enumerate(gemm_labels):
        a = tl[i]
        used = a[:, 0] != 0
        n = int(used.sum())
        if n == 0:
            continue
        a = a[used].double()
        cyc = (a[:, 7] - a[:, 0])
        ns = (a[:, 10] - a[:, 8])
        ghz = float((cyc / ns.clamp_min(1)).median())     # cycles per ns
        def us(col):
            v = (a[:, col] - a[:, 0]) / ghz / 1e3
            return float(v[a[:, col] != 0].mean()) if (a[:, col] != 0).any() else float("nan")
        span = float(a[:, 10].max() - a[:, 8].min()) / 1e3
        first_wait = float(a[:, 9].min())
        gap = "" if prev_exit is None else f"{(first_wait - prev_exit) / 1e3:7.2f} (since prev GEMM's last exit)"
        lines.append(f"{lab:16s} {n:5d} {us(1):6.2f} {us(2):8.2f} {us(3):7.2f} {us(4):10.2f} {us(5):10.2f} {us(6):9.2f} {us(7):6.2f} | {span:7.2f}  {gap}   clk={ghz:.2f}GHz"
                     f"   epi: ld0 {us(11):.2f} chunk0 {us(12):.2f} lastchunk {us(13):.2f} stats {us(14):.2f}")
        prev_exit = float(a[:, 10].max())
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"gemm_timeline_B{B}_T{T}.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
