"""Launch-by-launch trace of the native HiFi-GAN generator against the fp32 oracle (B200 only).

    python tools/voc_trace.py [--B 2] [--T 24] [--seed 21]

Stops the forward after every launch (mtts_voc_debug_set_launch_limit), reads the buffer that launch wrote out of the
workspace and compares it with the oracle's tensor at the same point of Generator.forward (hifigan/models.py:181-195).
Prints one line per launch; used by tests/test_gpu_hifigan.py::test_stage_trace and for debugging."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as HO  # noqa: E402


def stages(cfg: HO.HifiganCfg):
    """[(launch index (1-based), name, buffer, expected key, activation slope or None)] in launch order."""
    out = [(2, "conv_pre", "a_in", "conv_pre", 0.1)]
    n = 2
    nk, nd = len(cfg.resblock_kernel_sizes), len(cfg.resblock_dilation_sizes[0])
    for i in range(len(cfg.upsample_rates)):
        n += 1
        out.append((n, f"ups.{i} act", "xa", f"ups.{i}", 0.1))
        for j in range(nk):
            for m in range(nd):
                n += 2
                last_pair = m + 1 == nd
                key = f"resblocks.{i * nk + j}.pair{m}"
                if not last_pair:
                    out.append((n, key + " act", "r_act", key, 0.1))
                elif j + 1 < nk:
                    out.append((n, f"xs after resblock {i * nk + j}", "xs", f"xs.{i}.{j}", None))
                else:
                    out.append((n, f"level.{i} act", "a_in", f"level.{i}", 0.01 if i + 1 == len(cfg.upsample_rates) else 0.1))
    return out, n + 1


def run(B=2, T=24, seed=21, verbose=True):
    from matcha_tts_b200 import hifigan
    cfg = HO.HifiganCfg()
    sd = HO.make_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(seed)
    mel = -5.0 + 2.0 * torch.randn(B, 80, T, generator=g)
    trace = {}
    with torch.no_grad():
        wav_ref = HO.generator_forward(sd, mel, cfg, trace)
    nk = len(cfg.resblock_kernel_sizes)
    for i in range(len(cfg.upsample_rates)):       # running sums of the resblock outputs
        acc = None
        for j in range(nk):
            r = trace[f"resblocks.{i * nk + j}.pair{len(cfg.resblock_dilation_sizes[j]) - 1}"]
            acc = r if acc is None else acc + r
            trace[f"xs.{i}.{j}"] = acc
    gen = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
    gen.load_state_dict(HO.to_weight_norm(sd), strict=True)
    gen = gen.cuda()
    gen.use_cuda_graph = False
    eng = gen._engine(torch.device("cuda", torch.cuda.current_device()))
    lib = eng.lib
    st, total = stages(cfg)
    rows = []
    melc = mel.cuda()
    for launch, name, buf, key, slope in st:
        lib.mtts_voc_debug_set_launch_limit(eng.h, launch)
        eng.forward(melc, use_graph=False)
        torch.cuda.synchronize()
        ref = trace[key]
        if slope is not None:
            ref = F.leaky_relu(ref, slope)
        Bc, Cc, Lc = ref.shape
        wbuf, ptr, nbytes = eng.workspace(B, T)
        off = lib.mtts_voc_debug_buffer_offset(eng.h, B, T, buf.encode())
        start = ptr - wbuf.data_ptr() + off
        got = wbuf[start:start + Bc * Lc * Cc * 2].view(torch.float16).reshape(Bc, Lc, Cc).permute(0, 2, 1).float().cpu()
        d = (got.double() - ref.double())
        rows.append((launch, name, float(d.abs().max()), float(d.norm() / ref.double().norm()), float(ref.abs().max())))
        if verbose:
            print(f"launch {launch:3d}  {name:34s} max-abs {rows[-1][2]:.3e}  rel-L2 {rows[-1][3]:.3e}  (|ref| max {rows[-1][4]:.2f})", flush=True)
    lib.mtts_voc_debug_set_launch_limit(eng.h, -1)
    wav = eng.forward(melc, use_graph=False).cpu()
    d = (wav.double() - wav_ref.double())
    rows.append((total, "wav", float(d.abs().max()), float(d.norm() / wav_ref.double().norm()), float(wav_ref.abs().max())))
    if verbose:
        print(f"launch {total:3d}  {'wav':34s} max-abs {rows[-1][2]:.3e}  rel-L2 {rows[-1][3]:.3e}  launches {eng.launch_count()}", flush=True)
    return rows, eng.launch_count()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=24)
    ap.add_argument("--seed", type=int, default=21)
    a = ap.parse_args()
    run(a.B, a.T, a.seed)
