"""The case run under compute-sanitizer (one tool per gpurun call; see profiles/r02_sanitizer_*.txt):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py [B T n]
smoke()'s shape (B=2 x T=64, ragged) and one B x T solve (default 4 x 344, 2 Euler steps), eager launches (no graph:
the sanitizer instruments per launch), for the default kernels and -- MTTS_PAIRS=1 / MTTS_TAIL_PAIRS=1 in the
environment -- the CTA-pair variants.  Exits non-zero on a parity failure."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as U  # noqa: E402
from oracle import cfm_oracle as O  # noqa: E402


def main():
    B, T, n = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4, 344, 2)
    dec, cfg, sd = U.make_decoder(160)
    for (b, t, lengths, steps) in [(2, 64, [64, 41], 2), (B, T, [T] * (B - 1) + [max(1, T - 43)], n)]:
        mu, mask, z0, _ = O.make_inputs(cfg, b, t, lengths, seed=7)
        ref = O.euler_solve(sd, cfg, z0, mu, mask, steps)
        out = dec.solve(z0.cuda(), mu.cuda(), mask.cuda(), steps, None, "euler", use_graph=False)
        torch.cuda.synchronize()
        ma, rl = O.parity_errors(out.cpu(), ref, mask)
        print(f"sanitize_case: B={b} T={t} steps={steps} launches={dec.last_launch_count()} max-abs={ma:.3e} rel-L2={rl:.3e}", flush=True)
        assert ma <= O.TOL_MAX_ABS and rl <= O.TOL_REL_L2, (ma, rl)


if __name__ == "__main__":
    main()
