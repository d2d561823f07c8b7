"""Solve time (CUDA-graph replay, device-resident inputs) for several chain counts:
python tools/sweep_nsub.py [B T n_steps] -> prints ms per solve for MTTS_NSUB in {1,2,4,8}"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 344, 10)
    subs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    mu = torch.randn(B, 80, T, device=dev)
    z0 = torch.randn(B, 80, T, device=dev)
    mask = torch.ones(B, 1, T, device=dev)
    stream = torch.cuda.Stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for nsub in subs:
        os.environ["MTTS_NSUB"] = str(nsub)
        torch.manual_seed(0)
        dec = Decoder(160, 80, num_heads=2).to(dev)
        eng = dec._engine(dev)
        ws = eng.workspace(B, T)
        z = torch.empty_like(z0)
        with torch.cuda.stream(stream):
            def step():
                z.copy_(z0, non_blocking=True)
                _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0, ws[1], ws[2],
                                                    B, T, 1, stream.cuda_stream))
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); step(); b.record(stream)
                b.synchronize()
                ts.append(a.elapsed_time(b))
        ts.sort()
        out = z.clone()
        if ref is None:
            ref = out
        d = float((out - ref).abs().max())
        print(f"nsub={nsub}: {sum(ts)/len(ts):.3f} ms/solve (min {ts[0]:.3f})  {B*T/(sum(ts)/len(ts))*1e3/1e6:.3f} M frames/s  "
              f"launches={eng.launch_count()}  max|z - z(nsub={subs[0]})|={d:.3e}", flush=True)
        del dec, eng


if __name__ == "__main__":
    main()
