"""Verbose bring-up diagnostics on a B200: unit GEMMs, then a stage-by-stage comparison of every
intermediate of one estimator call against the oracle trace, then the 10-step solve.
Usage: python tools/gpu_debug.py [B T]"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cfm_oracle as O  # noqa: E402
import gpu_util as U  # noqa: E402


def gemm_case(eng, rows, Cc, N, shifts, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(rows, Cc, generator=g)).half()
    W = (torch.randn(N, len(shifts) * Cc, generator=g) / (len(shifts) * Cc) ** 0.5).half()
    bias = torch.randn(N, generator=g)
    ref = bias[None, :].repeat(rows, 1)
    Af = A.float()
    for i, s in enumerate(shifts):
        sh = torch.zeros_like(Af)
        if s == 0: sh = Af
        elif s > 0: sh[:-s] = Af[s:]
        else: sh[-s:] = Af[:s]
        ref += sh @ W[:, i * Cc:(i + 1) * Cc].float().T
    Ad, Wd, bd = A.cuda(), W.cuda(), bias.cuda()
    out = torch.zeros(rows, N, dtype=torch.float16, device="cuda")
    sh = (C.c_int * len(shifts))(*shifts)
    rc = eng.lib.mtts_debug_gemm(eng.h, Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), out.data_ptr(), rows, Cc, N,
                                 len(shifts), sh, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e = U.errs(out.float(), ref)
    print(f"gemm rows={rows} C={Cc} N={N} shifts={shifts}: rc={rc} max-abs {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    return e


def main():
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 48)
    torch.manual_seed(0)
    dec, cfg, sd = U.make_decoder(160)
    eng = dec._engine(torch.device("cuda", 0))
    print("engine ready; arena bytes", eng.lib.mtts_weight_arena_bytes(eng.h), flush=True)

    for (rows, Cc, N, shifts) in [(128, 64, 256, [0]), (300, 256, 256, [0]), (300, 256, 256, [-1, 0, 1]),
                                  (1000, 128, 128, [0]), (5000, 256, 1024, [0]), (700, 1024, 256, [0])]:
        gemm_case(eng, rows, Cc, N, shifts)

    lengths = [T, max(2, T - 17), max(2, T // 3)][:B] if B <= 3 else None
    mu, mask, z0, spks = O.make_inputs(cfg, B, T, lengths, seed=12)
    t = torch.linspace(0.05, 0.9, B)
    trace = {}
    ref = O.estimator_forward(sd, cfg, z0, mask, mu, t, spks, emu=O.Emu(trace=trace))
    H, LpT, LpH = T // 2, T + 2, T // 2 + 1
    d = lambda x: x.cuda()

    def run(limit):
        eng.lib.mtts_debug_set_launch_limit(eng.h, limit)
        out = dec(d(z0), d(mask), d(mu), d(t), None if spks is None else d(spks))
        torch.cuda.synchronize()
        return out

    def cmp(name, got, want):
        e = U.errs(got, want)
        flag = "" if e[1] < 5e-3 else "   <<<<<<<< MISMATCH"
        print(f"  {name:28s} max-abs {e[0]:.3e} rel {e[1]:.3e}{flag}", flush=True)

    def buf(name, L, Lp, cols):
        return U.flat_to_bct(U.ws_tensor(eng, B, T, name, B * Lp, cols), B, L, Lp)

    # prologue
    run(6)
    x0 = buf("x0", T, LpT, 256)
    xin = torch.cat([z0, mu], 1) * mask
    cmp("x0", x0[:, :160], xin)
    te6 = U.ws_tensor(eng, B, T, "te6", B, 1536, torch.float32)
    for s, (nm, _) in enumerate(O.stage_names(cfg)):
        tau = torch.nn.functional.linear(torch.nn.functional.mish(trace["temb"]), sd[nm + ".0.mlp.1.weight"], sd[nm + ".0.mlp.1.bias"])
        cmp(f"te6[{s}]", te6[:, s * 256:(s + 1) * 256], tau)

    stages = [("down_blocks.0", T, LpT, "skip0"), ("down_blocks.1", H, LpH, "skip1"), ("mid_blocks.0", H, LpH, "xM0"),
              ("mid_blocks.1", H, LpH, "xM1"), ("up_blocks.0", H, LpH, "xU0s"), ("up_blocks.1", T, LpT, "xU1s")]
    lvl_after = {0: ("xD0", H, LpH), 1: ("xD1", H, LpH), 4: ("xU0", T, LpT), 5: ("xF", T, LpT)}
    base = 6
    for si, (nm, L, Lp, outname) in enumerate(stages):
        print(f"stage {si} {nm} (L={L})")
        pf = nm + ":"
        run(base + 1); cmp("y1 (conv1 raw)", buf("y", L, Lp, 256), trace[pf + "y.block1"])
        run(base + 2); cmp("res", buf("res", L, Lp, 256), trace[pf + "res"])
        run(base + 3); cmp("h1", buf("h1", L, Lp, 256), trace[pf + "h1"])
        run(base + 4); cmp("y2 (conv2 raw)", buf("y", L, Lp, 256), trace[pf + "y.block2"])
        run(base + 5); cmp("xr", buf("xr", L, Lp, 256), trace[pf + "xr"]); cmp("a (LN1)", buf("a", L, Lp, 256), trace[pf + "a"].transpose(1, 2))
        run(base + 6)
        cmp("q", buf("q", L, Lp, 128), trace[pf + "q"].transpose(1, 2) * 0.125)
        cmp("k", buf("k", L, Lp, 128), trace[pf + "k"].transpose(1, 2))
        Lpad = (L + 7) // 8 * 8
        vt = U.ws_tensor(eng, B, T, "vt", B * 128, (T + 7) // 8 * 8)  # pitch is LpadT for the T level only
        if L == T:
            cmp("v^T", vt[:, :L].reshape(B, 128, L).float(), trace[pf + "v"].transpose(1, 2))
        else:
            vth = vt.reshape(-1)[:B * 128 * Lpad].reshape(B * 128, Lpad)
            cmp("v^T", vth[:, :L].reshape(B, 128, L).float(), trace[pf + "v"].transpose(1, 2))
        run(base + 7); cmp("o (attention)", buf("o", L, Lp, 128), trace[pf + "o"].transpose(1, 2))
        run(base + 8); cmp("xa", buf("xa", L, Lp, 256), trace[pf + "xa"].transpose(1, 2)); cmp("c (LN3)", buf("a", L, Lp, 256), trace[pf + "c"].transpose(1, 2))
        run(base + 9); cmp("s (snake)", buf("s", L, Lp, 1024), trace[pf + "s"].transpose(1, 2))
        run(base + 10); cmp("stage out", buf(outname, L, Lp, 256), trace[pf + "out"])
        base += 10
        if si in lvl_after:
            nm2, L2, Lp2 = lvl_after[si]
            run(base + 1); cmp("level conv " + nm2, buf(nm2, L2, Lp2, 256), trace[nm2])
            base += 1
    run(base + 2); cmp("hF", buf("h1", T, LpT, 256), trace["hF"])
    out = run(-1)
    cmp("estimator out", out.cpu(), ref)
    print("launches per estimator call:", dec.last_launch_count())

    # full solve
    for use_graph in (False, True):
        zr = O.euler_solve(sd, cfg, z0, mu, mask, 10, spks)
        t0 = time.time()
        zg = dec.solve(d(z0), d(mu), d(mask), 10, None if spks is None else d(spks), "euler", use_graph)
        torch.cuda.synchronize()
        ma, rl = O.parity_errors(zg.cpu(), zr, mask)
        print(f"solve 10 steps graph={use_graph}: max-abs {ma:.3e} rel-L2 {rl:.3e} (bar 2e-2 / 1e-3) wall {time.time()-t0:.3f}s launches {dec.last_launch_count()}", flush=True)


if __name__ == "__main__":
    main()
