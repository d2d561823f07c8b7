"""Do the kernels of different utterance chains overlap in time?  Captures one graph solve with the
GEMM timeline on and prints, for the first GEMM launches of step 1, [entry, exit] (us, globaltimer)
per chain.  python tools/chain_overlap.py [B T nsub]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsub = int(sys.argv[3]) if len(sys.argv) > 3 else 2
os.environ["MTTS_NSUB"] = str(nsub)
from matcha_tts_b200 import Decoder, _lib  # noqa: E402


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 344)
    n = 2
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dec = Decoder(160, 80, num_heads=2).to(dev)
    eng = dec._engine(dev)
    mu = torch.randn(B, 80, T, device=dev)
    z = torch.randn(B, 80, T, device=dev)
    mask = torch.ones(B, 1, T, device=dev)
    stream = torch.cuda.Stream(dev)
    ws = eng.workspace(B, T)
    per_chain = 30 * n          # GEMM launches per estimator evaluation and chain (conv1, conv2, qkv x6 + 4 level convs + 2 final)... see below
    nl = per_chain * nsub
    buf = torch.zeros(nl, 148, 16, dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        _lib.check(eng.lib.mtts_debug_set_timeline(eng.h, buf.data_ptr(), nl))
        for _ in range(3):
            _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0, ws[1], ws[2],
                                                B, T, 1, stream.cuda_stream))
        torch.cuda.synchronize()
    tl = buf.cpu()
    t0 = None
    for k in range(30, 30 + 30):          # all GEMMs of the second step
        row = []
        for c in range(nsub):
            a = tl[c * per_chain + k]
            used = a[:, 8] != 0
            if not used.any():
                row.append("   (not recorded)   ")
                continue
            ent, ex = int(a[used][:, 8].min()), int(a[used][:, 10].max())
            if t0 is None:
                t0 = ent
            row.append(f"chain{c}: ctas={int(used.sum()):3d} [{(ent - t0) / 1e3:8.2f}, {(ex - t0) / 1e3:8.2f}]")
        print(f"gemm#{k - 30:2d}  " + "   ".join(row))


if __name__ == "__main__":
    main()
