"""Phase timeline of ff_tail_kernel (first tile of CTA 0 and the mean over CTAs), from clock64 stamps.
python tools/tail_timeline.py [B T stage_limit]  -- runs the estimator up to the tail launch of stage 0 (level T; 6 + 7 launches).
With CTA pairs (the default) the stamps come from the pair leaders, one row per 256-row unit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 344)
    limit = int(sys.argv[3]) if len(sys.argv) > 3 else 6 + 7      # prologue + stage 0 (its tail is the only tail launch)
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
    buf = torch.zeros(148, 128, dtype=torch.int64, device=dev)

    def run():
        _lib.check(eng.lib.mtts_estimator_forward(eng.h, x.data_ptr(), mu.data_ptr(), mask.data_ptr(), t.data_ptr(), None,
                                                  out.data_ptr(), ws[1], ws[2], B, T, stream.cuda_stream))
    with torch.cuda.stream(stream):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_tail_timeline(eng.h, buf.data_ptr()))
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, limit))
        run()
        torch.cuda.synchronize()
        _lib.check(eng.lib.mtts_debug_set_launch_limit(eng.h, -1))
        _lib.check(eng.lib.mtts_debug_set_tail_timeline(eng.h, None))
    tl = buf.cpu().double()
    used = (tl[:, 0] != 0) & (tl[:, 64] != 0)
    n = int(used.sum())
    a = tl[used]
    t0 = a[:, 0:1]                                 # MMA: r_empty passed
    rel = (a - t0) / 1.965e3                       # us at 1.965 GHz
    rel[a == 0] = float("nan")
    m = torch.nanmean(rel, dim=0)
    print(f"{n} CTAs recorded (launch limit {limit}); times in us since the MMA thread entered the tile (mean over CTAs)")
    print(f"MMA : to_out issued {m[1]:.2f}   c_ready seen {m[2]:.2f}")
    print(f"EPI : r_full seen {m[64]:.2f}   c_ready arrive {m[65]:.2f}   r_done seen {m[66]:.2f}   r_empty arrive {m[67]:.2f}")
    print(" j | FF1 start  FF1 issued | FF2 s_ready  FF2 issued | EPI d1_full  EPI s_ready-arrive")
    for j in range(8):
        print(f" {j} | {m[4 + 4 * j]:9.2f} {m[5 + 4 * j]:11.2f} | {m[6 + 4 * j]:11.2f} {m[7 + 4 * j]:11.2f} | {m[64 + 4 + 2 * j]:11.2f} {m[64 + 5 + 2 * j]:18.2f}")


if __name__ == "__main__":
    main()
