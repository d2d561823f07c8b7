"""Phase timeline of gnb_qkv_kernel (worker warp 0 and the MMA warp, first tiles of every CTA) from clock64 stamps.
python tools/gq_timeline.py [B T]  -- runs the estimator up to the first gnb_qkv launch (stage 0, level T)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 344)
    limit = 6 + 4                      # prologue, conv1, gnA, conv2, gnb_qkv
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    dec.set_chains(1)
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
    raw = buf.cpu().double().reshape(148, -1)
    tl = raw[:, :72].reshape(148, 6, 12)
    used = tl[:, 0, 0] != 0
    a = tl[used]
    t0 = a[:, 0:1, 0:1]
    rel = (a - t0) / 1.9e3                          # us at ~1.9 GHz since the CTA's first tile started
    rel[a == 0] = float("nan")
    m = rel.nanmean(0)
    names = ["T start", "batch 0", "batch 1", "batch 2", "batch 3", "slow rows", "a_ready arrive", "-",
             "MMA go", "MMA issued", "E: d_full seen", "E done"]
    print(f"{int(used.sum())} CTAs; us since the CTA's first tile started (mean over CTAs); B={B} T={T}")
    print("tile | " + " | ".join(f"{n:>14s}" for n in names))
    for i in range(6):
        print(f"{i:4d} | " + " | ".join(f"{float(m[i, j]):14.2f}" for j in range(12)))
    pw = raw[used][:, 72:120].reshape(-1, 8, 6)
    relw = (pw - a[:, 0, 0].reshape(-1, 1, 1)) / 1.9e3
    mw = relw.mean(0)
    print("per worker warp (mean over CTAs): T(2) start | T(2) a_ready arrive | E(1) d_full seen | E(1) done")
    for w in range(8):
        print(f"  warp {w + 2} (partition {(w + 2) % 4}): " + "  ".join(f"{float(mw[w, k]):8.2f}" for k in range(4)))


if __name__ == "__main__":
    main()
