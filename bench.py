#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the CFM decoder Euler solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one full n_timesteps=10 Euler solve of one batch (BASELINE configs[1]: LJSpeech-shape
decoder, B=64 utterances x T_mel=344 frames, random-init weights, synthetic mu / noise).
  value : whole-job mel-frames/s with inputs resident in HBM (CUDA-graph replay of the solve).  The K steps are K
          independent batches; `--in-flight F` (default 5) of them are in flight at a time, each on its own solve
          lane (CUDA stream + native handle told about F through mtts_set_lanes: every persistent launch then takes
          its share of the SMs), the way a serving process overlaps consecutive batches: one solve is a serial chain
          of ~490 latency-bound kernels whose CTAs get ~1 tile each on 148 SMs.  Timed with CUDA
          events around all K steps, max over ranks; the steps rotate over distinct input sets larger than L2.
          config.serial holds the one-solve-at-a-time figure (L2 flushed between steps), config.sustained the same
          loop over 120 steps (under the power cap).
  e2e   : the same metric through the public API CFM.forward(mu, mask, n_timesteps, temperature)
          with mu/mask in pinned HOST memory and the mel read back to the host every step.
  roofline : tcgen05 GEMM kernel (all convs + linears): algorithmic FLOPs / CUDA-event time of its
          launches inside one solve, against the measured sustained bf16 peak.
  cpu_baseline : the reference's own CFM.forward (baseline/_ref/model.py, staged by tools/stage_reference.sh; the oracle
          port if it is not staged) on this box's host cores (rank 0, N=1): a bounded sample of the workload + config 1.
  config5 : BASELINE config 5 (4096 multi-speaker utterances, bucketed, sharded over the ranks, NCCL gather included).
  reference_on_gpu : informative -- the unmodified reference through PyTorch eager on the same GPU (N=1).
`--impl reference` times only the CPU path (bounded sample of the same workload) and prints the
same JSON shape with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
# ... also when a native library printf()s (NCCL's version banner ignores NCCL_DEBUG_FILE): file descriptor 1 is pointed at
# stderr for the whole process and the JSON line goes to a private duplicate of the real stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)

METRIC = "mel-frames/sec for CFM decoder sampling (10 Euler steps)"
UNIT = "mel-frames/s"
GEMM_FLOP_PER_FRAME_STEP = 10_985_472       # SURVEY.md App. C (conv + linear GEMMs, Cin=160)
ATTN_FLOP_PER_FRAME_STEP_PER_T = 1536       # + 1536*T for attention


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1362.7), d.get("hbm_gbs", 6535.4), "measured"
    return 1400.0, 6650.0, "fallback"     # B200_PROFILING.md fallback (sustained ~1.4 PF, 6.65 TB/s)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's OWN implementation when it is staged (baseline/_ref/model.py, copied by
# tools/stage_reference.sh; git-ignored, travels with the gpurun snapshot), else the oracle port
# ------------------------------------------------------------------------------------------------
def load_reference():
    """The reference's model.py (CFM.forward :1136, BASECFM Euler loop :1084-1109, Decoder :834-1048) or None."""
    path = os.path.join(ROOT, "baseline", "_ref", "model.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("matcha_reference_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_hifigan():
    """The reference's vendored HiFi-GAN Generator with config v1 (hifigan/models.py:148, config.py) or None.  xutils.py imports
    matplotlib for a plotting helper the path never calls; when the image lacks it an empty stand-in module is registered."""
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(root, "hifigan", "models.py")):
        return None
    import types
    sys.dont_write_bytecode = True
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        m = types.ModuleType("matplotlib")
        m.use = lambda *a, **k: None
        m.pylab = types.ModuleType("matplotlib.pylab")
        sys.modules["matplotlib"], sys.modules["matplotlib.pylab"] = m, m.pylab
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from hifigan.config import v1
        from hifigan.env import AttrDict
        from hifigan.models import Generator
        import torch
        torch.manual_seed(0)
        gen = Generator(AttrDict(v1)).eval()
    return gen


def reference_cfm(ref, device="cpu"):
    """The reference's CFM around its Decoder with the hyper-parameters of reference main.py:67-75, seed 0, eval()."""
    import torch
    torch.manual_seed(0)
    dec = ref.Decoder(in_channels=160, out_channels=80, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                      n_blocks=1, num_mid_blocks=2, num_heads=2, act_fn="snakebeta")
    return ref.CFM(80, {"solver": "euler", "sigma_min": 1e-4}, estimator=dec).eval().to(device)


def cpu_reference(B, T, n_timesteps, steps, warmup, budget_s):
    """Times the CPU path on a bounded sample (rows of the B x T batch) with every host thread torch will use.
    kind = "reference": reference CFM.forward from baseline/_ref; kind = "port": oracle.euler_solve."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    ref = load_reference()
    if ref is not None:
        cfm = reference_cfm(ref)
        kind = "reference"

        def make(rows):
            g = torch.Generator().manual_seed(1)
            return torch.randn(rows, 80, T, generator=g), torch.ones(rows, 1, T)

        def solve(inp, n):
            cfm(inp[0], inp[1], n, temperature=0.667)
    else:
        from oracle import cfm_oracle as O
        cfg = O.DecoderCfg()
        sd = O.make_state_dict(cfg, 0)
        kind = "port"

        def make(rows):
            mu, mask, z0, _ = O.make_inputs(cfg, rows, T, None, seed=1)
            return mu, mask, z0

        def solve(inp, n):
            with torch.inference_mode():
                O.euler_solve(sd, cfg, inp[2], inp[0], inp[1], n)
    # B=1 x T, 1 warm-up, best of 5
    one = make(1)
    solve(one, 1)                                                     # page in / thread pool start
    solve(one, n_timesteps)
    t1 = []
    for _ in range(5):
        t0 = time.perf_counter()
        solve(one, n_timesteps)
        t1.append(time.perf_counter() - t0)
    config1 = {"batch": 1, "t_mel": T, "seconds_best_of_5": min(t1), "value": T / min(t1), "unit": UNIT}
    # the line's value: a bounded sample of the B x T workload, sized so that (steps + warmup) solves fit the budget
    rows = int(max(1, min(16, B, budget_s / max(1, steps + warmup) / min(t1))))
    inp = make(rows)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        solve(inp, n_timesteps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    what = ("reference CFM.forward (baseline/_ref/model.py, unmodified)" if kind == "reference"
            else "oracle port of the reference PyTorch path (baseline/_ref not staged)")
    return {"value": rows * T / mean, "ms_per_step": mean * 1e3, "cores": cores, "rows": rows, "kind": kind, "config1": config1,
            "sample": f"{what}: {rows} of {B} rows x T={T}, {n_timesteps} Euler steps, fp32 torch-CPU on {cores} threads, "
                      f"{steps} timed solves after {warmup} warm-up (mean)"}


def reference_on_gpu(B, T, n_timesteps, dev, reps=3):
    """Informative (SURVEY.md section 2.1 / BASELINE.md section 2): the unmodified reference through PyTorch eager on this
    B200 -- fp32 with torch's default TF32 convolutions, and autocast(bfloat16) -- "the library kernels to beat"."""
    import torch
    ref = load_reference()
    if ref is None:
        return None
    cfm = reference_cfm(ref, dev)
    g = torch.Generator().manual_seed(1)
    mu = torch.randn(B, 80, T, generator=g).to(dev)
    mask = torch.ones(B, 1, T, device=dev)
    out = {}
    for name, ctx in (("eager_fp32_tf32conv", None), ("autocast_bf16", torch.bfloat16)):
        try:
            def run():
                if ctx is None:
                    return cfm(mu, mask, n_timesteps, temperature=0.667)
                with torch.autocast("cuda", dtype=ctx):
                    return cfm(mu, mask, n_timesteps, temperature=0.667)
            run()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                z = run()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"value": B * T / (ms * 1e-3), "unit": UNIT, "ms_per_solve": ms, "finite": bool(torch.isfinite(z).all())}
        except Exception as exc:      # an eager-path failure must not take the bench line down
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    out["note"] = f"reference CFM.forward, PyTorch eager on the same GPU, batch {B} x T={T}, {n_timesteps} steps, mean of {reps}"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.batch, args.frames, args.n_timesteps, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, extra={"device": "host CPU", "config1_cpu": r["config1"],
                                               "note": "the reference's CPU PyTorch path on this box's host cores, one process "
                                                       "(it does not shard); the reference is pure Python/PyTorch, nothing compiled"}),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, parsed from the committed summary
    of this round's `ncu --set full` capture (profiles/roofline_traffic.json names the capture it came from); None when
    no current capture is committed."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        d = json.load(open(path))
        return {"bytes_per_launch": d["dram_bytes_read"] + d["dram_bytes_write"], "kernel": d["kernel"], "source": d["source"],
                "algorithmic_bytes_per_launch": d.get("algorithmic_bytes")}
    except Exception:
        return None


def workload_config(args, extra=None):
    c = {"workload": "BASELINE configs[1]: Matcha-TTS LJSpeech-shape CFM decoder (11.0 M params, random init), "
                     f"batch {args.batch} x T_mel={args.frames}, n_timesteps={args.n_timesteps}, "
                     + ("ragged lengths U{300..T}" if args.ragged else "all rows full length (mask all ones)"),
         "batch": args.batch, "t_mel": args.frames, "n_timesteps": args.n_timesteps,
         "frames_per_step": args.batch * args.frames, "parallelism": f"dp{args.gpus} (utterance-sharded, no hot-path collective)",
         "l2": "flushed between timed steps (256 MiB write)"}
    if extra:
        c.update(extra)
    return c


def run_synthesize(dev, B, n_timesteps, steps, world):
    """tokens -> mel through the reference's own entry point (model.py:1264-1300) on the native modules: B utterances of 86
    tokens, random-init weights with the duration head biased to ~4 frames per token (a random head gives ~1), so the batch
    decodes ~B x 344 frames like configs[1].  One call at a time (synthesize reads y_lengths.max() on the host, :1281)."""
    import math
    import types
    import torch
    import torch.distributed as dist
    from matcha_tts_b200 import MatchaTTS
    enc_p = types.SimpleNamespace(encoder_type="RoPE Encoder", n_feats=80, n_channels=192, filter_channels=768, n_heads=2, n_layers=6,
                                  kernel_size=3, p_dropout=0.1, prenet=True)
    dec_p = types.SimpleNamespace(channels=(256, 256), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=2, num_heads=2,
                                  act_fn="snakebeta")
    dur_p = types.SimpleNamespace(filter_channels_dp=256, kernel_size=3, p_dropout=0.1)
    torch.manual_seed(0)
    m = MatchaTTS(178, 1, 64, enc_p, dec_p, {"solver": "euler", "sigma_min": 1e-4}, dur_p).to(dev)
    with torch.no_grad():
        m.encoder.proj_w.proj.bias.fill_(math.log(3.5))
    Tx = 86
    g = torch.Generator().manual_seed(7)
    tok_h = torch.randint(0, 178, (B, Tx), generator=g).pin_memory()
    len_h = torch.full((B,), Tx, dtype=torch.long).pin_memory()
    stream = torch.cuda.Stream(dev)
    out_h = None
    frames = 0
    enc_ms = 0.0
    with torch.cuda.stream(stream):
        def once():
            nonlocal out_h
            tok = tok_h.to(dev, non_blocking=True)
            lens = len_h.to(dev, non_blocking=True)
            mel, ylen, _ = m.synthesise(tok, lens, n_timesteps=n_timesteps, temperature=0.667)
            if out_h is None or out_h.shape != mel.shape:
                out_h = torch.empty(mel.shape, dtype=mel.dtype).pin_memory()
            out_h.copy_(mel, non_blocking=True)
            yl = ylen.to("cpu", non_blocking=True)
            stream.synchronize()
            return int(yl.sum())
        for _ in range(3):
            once()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            frames += once()
        dt = time.perf_counter() - t0
        # the text encoder's share: CUDA events around its call alone
        tok, lens = tok_h.to(dev), len_h.to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            m.encoder(tok, lens)
        e1.record(stream)
        e1.synchronize()
        enc_ms = e0.elapsed_time(e1) / 10
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # the whole reference pipeline for ONE sentence ("end-to-end RTF (text -> mel -> HiFi-GAN -> denoiser)",
    # MOS_audiou_generator.ipynb:235-258: mean 0.0173): tokens on the host -> synthesise -> vocoder -> denoiser -> audio on the host
    rtf = None
    if world == 1:
        try:
            from matcha_tts_b200 import hifigan
            from oracle import hifigan_oracle as HO          # seeded vocoder weights only
            voc = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
            voc.load_state_dict(HO.to_weight_norm(HO.make_state_dict(HO.HifiganCfg(), 0)), strict=True)
            voc = voc.to(dev)
            tok1 = torch.randint(0, 178, (1, 120), generator=g).pin_memory()
            len1 = torch.full((1,), 120, dtype=torch.long).pin_memory()
            with torch.cuda.stream(stream):
                den = hifigan.Denoiser(voc, mode="zeros")

                def sentence():
                    mel, ylen, _ = m.synthesise(tok1.to(dev, non_blocking=True), len1.to(dev, non_blocking=True), n_timesteps=n_timesteps,
                                                temperature=0.667)
                    audio = den(voc(mel).clamp(-1, 1).squeeze(1), strength=0.00025)
                    a_h = audio.to("cpu", non_blocking=True)
                    stream.synchronize()
                    return a_h.shape[-1] / 22050.0
                for _ in range(3):
                    secs = sentence()
                t0 = time.perf_counter()
                for _ in range(10):
                    secs = sentence()
                call = (time.perf_counter() - t0) / 10
            rtf = {"value": call / secs, "ms_per_sentence": call * 1e3, "audio_seconds": secs, "x_real_time": secs / call,
                   "what": "one sentence of 120 tokens, host tokens in -> MatchaTTS.synthesise (10 Euler steps) -> hifigan.Generator -> "
                           "Denoiser -> audio on the host, mean of 10 calls; the reference's notebook reports RTF 0.0173 for this chain "
                           "(BASELINE.md section 1, other hardware)"}
            del voc, den
        except Exception as exc:
            rtf = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return {"value": world * frames / dt, "unit": "valid mel-frames/s", "ms_per_call": dt / steps * 1e3, "calls": steps, "batch": B,
            "sentence_rtf": rtf,
            "tokens_per_utterance": Tx, "mel_frames_per_call": frames // steps, "text_encoder_ms": enc_ms,
            "text_encoder_launches": m.encoder.last_launch_count(),
            "h2d_bytes_per_call": tok_h.numel() * 8 + len_h.numel() * 8, "d2h_bytes_per_call": out_h.numel() * 4,
            "api": "MatchaTTS.synthesise(x, x_lengths, n_timesteps, temperature): native text encoder + duration predictor, alignment glue in "
                   "torch (one host read of y_lengths.max()), native CFM decoder; pinned-host tokens in, mels out, one call at a time"}


def run_vocoder(dev, B, T, steps, cpu_budget_s=12.0):
    """The step after the path: hifigan.Generator.forward(mel) -> wav on the native kernels, B x T
    mel frames per call (the batch one solve produces).  `value`: mel resident in HBM, CUDA events on the launching stream;
    `e2e`: pinned-host mel in, pinned-host waveform out inside the timed region; per-launch times for the roofline; the
    reference's own Generator on the host cores (bounded sample) and through PyTorch eager on this GPU next to it."""
    import ctypes as C
    import torch
    from matcha_tts_b200 import hifigan
    from oracle import hifigan_oracle as HO          # weights only (seeded state-dict); the checker, not the thing measured
    cfg = HO.HifiganCfg()
    sd = HO.make_state_dict(cfg, 0)
    gen = hifigan.Generator(hifigan.AttrDict(hifigan.v1))
    gen.load_state_dict(HO.to_weight_norm(sd), strict=True)
    gen = gen.to(dev)
    g = torch.Generator().manual_seed(3)
    NSET = 3
    mels_h = [(-5.0 + 2.0 * torch.randn(B, 80, T, generator=g)).pin_memory() for _ in range(NSET)]
    mels_d = [m.to(dev) for m in mels_h]
    stream = torch.cuda.Stream(dev)
    flop = HO.flops_per_frame(cfg) * B * T
    with torch.cuda.stream(stream):
        for i in range(3):
            wav = gen(mels_d[i % NSET])
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            wav = gen(mels_d[i % NSET])
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        wav_h = torch.empty(wav.shape, dtype=wav.dtype).pin_memory()
        t0 = time.perf_counter()
        for i in range(steps):
            w = gen(mels_h[i % NSET].to(dev, non_blocking=True))
            wav_h.copy_(w, non_blocking=True)
            stream.synchronize()
        e2e_ms = (time.perf_counter() - t0) / steps * 1e3
        finite = bool(torch.isfinite(wav_h).all()) and float(wav_h.abs().max()) <= 1.0
        # per-launch device times (event pairs; graph bypassed meanwhile)
        eng = gen._engine(dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device()))
        lib = eng.lib
        gen.use_cuda_graph = False
        gen(mels_d[0])
        best = None
        for _ in range(2):
            lib.mtts_voc_debug_profile_begin(eng.h, stream.cuda_stream)
            gen(mels_d[0])
            n = 128
            tms, kind, fl = (C.c_float * n)(), (C.c_int * n)(), (C.c_double * n)()
            cnt = lib.mtts_voc_debug_profile_end(eng.h, n, tms, kind, fl)
            rows = [(tms[i], kind[i], fl[i]) for i in range(cnt)]
            if best is None or sum(r[0] for r in rows) < sum(r[0] for r in best):
                best = rows
        gen.use_cuda_graph = True
    gemm_ms = sum(r[0] for r in best if r[1] == 0)
    gemm_fl = sum(r[2] for r in best if r[1] == 0)
    peak, _, peak_src = measured_peaks()
    out = {"value": B * T / (ms * 1e-3), "unit": UNIT, "ms_per_call": ms, "calls": steps, "batch": B, "frames": T,
           "samples_per_call": B * T * cfg.hop, "x_real_time_22050": B * T * cfg.hop / 22050.0 / (ms * 1e-3),
           "model_tflops": flop / (ms * 1e-3) / 1e12, "flop_per_mel_frame": HO.flops_per_frame(cfg), "finite": finite,
           "launches_per_call": len(best),
           "e2e": {"value": B * T / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_call": e2e_ms, "h2d_bytes_per_call": B * 80 * T * 4,
                   "d2h_bytes_per_call": B * T * cfg.hop * 4,
                   "api": "hifigan.Generator.forward(mel): pinned-host mel in, pinned-host waveform out, one call at a time"},
           "roofline": {"bound": "tensor", "achieved": gemm_fl / (gemm_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": gemm_fl / (gemm_ms * 1e-3) / 1e12 / peak, "traffic": None,
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src}): the launches run inside a long step",
                        "what": f"the {sum(1 for r in best if r[1] == 0)} voc_conv_kernel launches of one call, CUDA-event pairs per launch "
                                f"({gemm_ms:.2f} ms of {sum(r[0] for r in best):.2f} ms), dense FLOPs at the real channel widths "
                                "(the zero weights of the ConvTranspose phases and the padded mel channels are not counted)"},
           "api": "matcha_tts_b200.hifigan.Generator(AttrDict(v1)).forward(mel) == reference hifigan/models.py:181-195; fp16 operands "
                  "and activations, fp32 accumulate; 13.9 M parameters, seeded weights"}
    # the reference's own Generator: host cores on a bounded sample, PyTorch eager on this GPU at the full batch
    try:
        ref = load_reference_hifigan()
        if ref is not None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.load_state_dict(HO.to_weight_norm(sd), strict=True)
            with torch.no_grad():
                rows = 1
                x = mels_h[0][:rows].clone()
                t0 = time.perf_counter()
                ref(x)
                one = time.perf_counter() - t0
                reps = max(1, min(5, int(cpu_budget_s / max(one, 1e-3)) - 1))
                t0 = time.perf_counter()
                for _ in range(reps):
                    y_cpu = ref(x)
                cpu_s = (time.perf_counter() - t0) / reps
                out["cpu_baseline"] = {"value": rows * T / cpu_s, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
                                       "sample": f"reference hifigan.models.Generator (staged baseline/_ref/hifigan), {rows} of {B} rows x T={T}, "
                                                 f"fp32 torch-CPU, mean of {reps} after 1 warm-up"}
                d = wav_h[:rows].double() - y_cpu.double()
                out["vs_reference_wav"] = {"max_abs": float(d.abs().max()), "rel_l2": float(d.norm() / y_cpu.double().norm())}
                refg = ref.to(dev)
                res = {}
                for name, ctx in (("eager_fp32_tf32conv", None), ("autocast_bf16", torch.bfloat16)):
                    def run():
                        if ctx is None:
                            return refg(mels_d[0])
                        with torch.autocast("cuda", dtype=ctx):
                            return refg(mels_d[0])
                    run()
                    torch.cuda.synchronize(dev)
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    for _ in range(3):
                        run()
                    a1.record()
                    a1.synchronize()
                    t = a0.elapsed_time(a1) / 3
                    res[name] = {"value": B * T / (t * 1e-3), "unit": UNIT, "ms_per_call": t}
                out["reference_on_gpu"] = res
                del refg
    except Exception as exc:
        out["reference_error"] = f"{type(exc).__name__}: {exc}"[:300]
    del gen, mels_d
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from matcha_tts_b200 import CFM, Decoder, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, n = args.batch, args.frames, args.n_timesteps

    torch.manual_seed(0)
    dec = Decoder(in_channels=160, out_channels=80, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                  n_blocks=1, num_mid_blocks=2, num_heads=2, act_fn="snakebeta").to(dev)
    cfm = CFM(80, {"solver": "euler", "sigma_min": 1e-4}, estimator=dec)
    g = torch.Generator().manual_seed(1 + rank)
    mu_h = torch.randn(B, 80, T, generator=g).pin_memory()
    if args.ragged:
        lengths = torch.randint(300, T + 1, (B,), generator=g)
        lengths[0] = T
    else:
        lengths = torch.full((B,), T)
    mask_h = (torch.arange(T)[None, :] < lengths[:, None]).unsqueeze(1).float().pin_memory()
    z0 = (torch.randn(B, 80, T, generator=g) * 0.667).to(dev)
    mu, mask = mu_h.to(dev), mask_h.to(dev)
    eng = dec._engine(dev)
    stream = torch.cuda.Stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    z = torch.empty_like(z0)
    gathered = torch.empty(world * B, 80, T, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        z.copy_(z0, non_blocking=True)
        _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0,
                                            eng.workspace(B, T)[1], eng.workspace(B, T)[2], B, T, 1, stream.cuda_stream))
        if world > 1:
            dist.all_gather_into_tensor(gathered, z)          # final mel gather (NCCL over NVLink)

    # ---- device-resident, one solve at a time (config.serial) ----
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        barrier()
        launches_per_step = eng.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        clocks = ClockSampler(local)
        clocks.__enter__()                                      # sampled over all timed regions
        time.sleep(0.1)
        barrier()
        for a, b in evs:
            flush.fill_(1)                                      # evict L2 between timed steps
            a.record(stream)
            step_device()
            b.record(stream)
        barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
    serial_ms = sum(ms)
    if world > 1:
        t = torch.tensor([serial_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        serial_ms = float(t.item())
    frames = world * B * T * args.steps
    serial_value = frames / (serial_ms * 1e-3)

    # ---- device-resident throughput: F solves in flight on F lanes, rotating input sets ----
    F = max(1, args.in_flight)
    NSET = 4                                                    # input sets per lane: F*NSET*(mu, z0, z) > L2 (126 MB)
    lanes = [stream] + [torch.cuda.Stream(dev) for _ in range(F - 1)]
    lane_eng, lane_sets, lane_gather = [], [], []
    for li, ls in enumerate(lanes):
        with torch.cuda.stream(ls):
            lane_eng.append(dec._engine(dev))                   # one native engine per stream
        if F > 1:
            lane_eng[-1].set_chains(1)                          # the solves overlap each other: no split inside a solve ...
            lane_eng[-1].set_lanes(F)                           # ... and every persistent launch takes its share of the SMs
        sets = []
        for k in range(NSET):
            gk = torch.Generator().manual_seed(100 + 10 * li + k + 1000 * rank)
            sets.append((torch.randn(B, 80, T, generator=gk).to(dev), (torch.randn(B, 80, T, generator=gk) * 0.667).to(dev),
                         torch.empty(B, 80, T, device=dev)))
        lane_sets.append(sets)
        lane_gather.append(torch.empty(world * B, 80, T, device=dev) if world > 1 else None)
    set_bytes = F * NSET * 3 * B * 80 * T * 4

    def step_lane(i):
        li, k = i % F, (i // F) % NSET
        ls, le = lanes[li], lane_eng[li]
        mu_k, z0_k, z_k = lane_sets[li][k]
        with torch.cuda.stream(ls):
            z_k.copy_(z0_k, non_blocking=True)
            _lib.check(le.lib.mtts_euler_solve(le.h, z_k.data_ptr(), mu_k.data_ptr(), mask.data_ptr(), None, n, 0,
                                               le.workspace(B, T)[1], le.workspace(B, T)[2], B, T, 1, ls.cuda_stream))
            if world > 1:
                dist.all_gather_into_tensor(lane_gather[li], z_k)

    def run_lanes(steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for ls in lanes[1:]:
            ls.wait_event(ev0)
        for i in range(steps):
            step_lane(i)
        for ls in lanes[1:]:
            e = torch.cuda.Event()
            e.record(ls)
            stream.wait_event(e)
        ev1.record(stream)
        return ev0, ev1

    with torch.cuda.stream(stream):
        run_lanes(max(args.warmup, F * NSET))                   # captures the CUDA graph of every (lane, input set)
        launches_per_step = lane_eng[0].launch_count()          # kernels of one solve as the timed region runs it (one chain per solve when F > 1)
        barrier()
        ev0, ev1 = run_lanes(args.steps)
        barrier()
        total_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = frames / (total_ms * 1e-3)

    # ---- the same loop over >= 120 steps: the sustained figure under the power cap (the K-step region above is a burst) ----
    sustained = None
    if args.sustained_steps > 0:
        with torch.cuda.stream(stream):
            barrier()
            s0, s1 = run_lanes(args.sustained_steps)
            barrier()
            sus_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([sus_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sus_ms = float(t.item())
        sustained = {"steps": args.sustained_steps, "value": world * B * T * args.sustained_steps / (sus_ms * 1e-3), "unit": UNIT,
                     "ms_per_step": sus_ms / args.sustained_steps}

    # ---- end to end through the public API: pinned host inputs -> CFM.forward -> host mel ----
    # Every step copies its own inputs host->device and its result device->host inside the timed region.
    # The loop is pipelined the way a serving process is: the copies of step i+1 / i-1 run on copy streams
    # while CFM.forward of step i computes (double-buffered device inputs and pinned host outputs).
    NB = 2 * F                                                  # rotating device-input / pinned-output buffers
    out_h = [torch.empty(B, 80, T).pin_memory() for _ in range(NB)]
    mu_dv = [torch.empty(B, 80, T, device=dev) for _ in range(NB)]
    mask_dv = [torch.empty(B, 1, T, device=dev) for _ in range(NB)]
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run_e2e(steps):
        ev_in = [torch.cuda.Event() for _ in range(steps)]
        ev_out = [torch.cuda.Event() for _ in range(steps)]
        ev_host = [torch.cuda.Event() for _ in range(steps)]      # host result buffer written
        mels = [None] * NB
        for i in range(steps):
            k, ls = i % NB, lanes[i % F]
            if i >= NB:
                ev_host[i - NB].synchronize()                           # buffers k are free again, result i-NB is on the host
            with torch.cuda.stream(h2d):
                mu_dv[k].copy_(mu_h, non_blocking=True)
                mask_dv[k].copy_(mask_h, non_blocking=True)
                ev_in[i].record(h2d)
            with torch.cuda.stream(ls):
                ls.wait_event(ev_in[i])
                mel = cfm(mu_dv[k], mask_dv[k], n, temperature=0.667)    # the public call (reference model.py:1136)
                ev_out[i].record(ls)
            mels[k] = mel                                               # keep alive until copied out
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev_out[i])
                out_h[k].copy_(mel, non_blocking=True)
                ev_host[i].record(d2h)
        for i in range(max(0, steps - NB), steps):
            ev_host[i].synchronize()

    with torch.cuda.stream(stream):
        run_e2e(max(NB, args.warmup))
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    clocks.__exit__(None, None, None)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = frames / e2e_s

    # ---- per-kernel timing of one solve (CUDA events around every launch) for the roofline ----
    roof = None
    kinds = {}
    if F > 1:
        eng.set_lanes(1)                                        # each launch below runs alone on the GPU: full grids
    with torch.cuda.stream(stream):
        z.copy_(z0)
        torch.cuda.synchronize(dev)
        _lib.check(eng.lib.mtts_debug_profile_begin(eng.h, stream.cuda_stream))
        _lib.check(eng.lib.mtts_euler_solve(eng.h, z.data_ptr(), mu.data_ptr(), mask.data_ptr(), None, n, 0,
                                            eng.workspace(B, T)[1], eng.workspace(B, T)[2], B, T, 0, stream.cuda_stream))
        cap = 8192
        msb, kb, fb = (C.c_float * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
        cnt = eng.lib.mtts_debug_profile_end(eng.h, cap, msb, kb, fb)
    if cnt > 0:
        names = {0: "gemm_tc", 1: "attention", 2: "gn_apply", 3: "other"}
        for i in range(min(cnt, cap)):
            k = kinds.setdefault(names[kb[i]], {"launches": 0, "ms": 0.0, "flop": 0.0})
            k["launches"] += 1
            k["ms"] += msb[i]
            k["flop"] += fb[i]
        peak_tf, _, which = measured_peaks()
        gm = kinds["gemm_tc"]
        achieved = gm["flop"] / (gm["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel + ff_tail_kernel (tcgen05 implicit GEMMs: all convs + linears)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": f"{which} sustained bf16 (MEASURED_PEAKS.json); fp16 runs at the same tcgen05 rate",
                "traffic": traffic_from_profiles(),
                "avg_launch_us": gm["ms"] * 1e3 / gm["launches"], "launches_per_solve": gm["launches"],
                "algorithmic_flop_per_solve": gm["flop"],
                "share_of_solve": {k: round(v["ms"] / sum(x["ms"] for x in kinds.values()), 4) for k, v in kinds.items()},
                "ms_per_solve_by_kernel": {k: round(v["ms"], 4) for k, v in kinds.items()}}

    # ---- the same per-launch event timing at the occupancy of the timed region: ONE eager solve over the F*B rows the F in-flight
    #      solves hold together (a single batch-B solve leaves 41 % of the SMs idle at level T/2; the timed region does not) ----
    if roof is not None and F > 1:
        BF = F * B
        muF, zF, maskF = torch.randn(BF, 80, T, device=dev), torch.randn(BF, 80, T, device=dev) * 0.667, mask.repeat(F, 1, 1).contiguous()
        with torch.cuda.stream(stream):
            wsF = eng.workspace(BF, T)
            for rep in range(2):                                # first pass: tables / tensor maps of the new shape
                if rep == 1:
                    torch.cuda.synchronize(dev)
                    _lib.check(eng.lib.mtts_debug_profile_begin(eng.h, stream.cuda_stream))
                _lib.check(eng.lib.mtts_euler_solve(eng.h, zF.data_ptr(), muF.data_ptr(), maskF.data_ptr(), None, n, 0, wsF[1], wsF[2],
                                                    BF, T, 0, stream.cuda_stream))
            cntF = eng.lib.mtts_debug_profile_end(eng.h, cap, msb, kb, fb)
        gms = sum(msb[i] for i in range(min(cntF, cap)) if kb[i] == 0)
        gfl = sum(fb[i] for i in range(min(cntF, cap)) if kb[i] == 0)
        gln = sum(1 for i in range(min(cntF, cap)) if kb[i] == 0)
        allms = sum(msb[i] for i in range(min(cntF, cap)))
        if gms > 0:
            # the headline roofline figure: per-launch event timing at the occupancy of the timed region.  The single batch-B eager solve
            # above (59 % of the SMs busy at level T/2, event pairs between launches) moves to roofline.one_solve.
            one = {k: roof[k] for k in ("achieved", "frac", "avg_launch_us", "launches_per_solve", "algorithmic_flop_per_solve",
                                        "share_of_solve", "ms_per_solve_by_kernel")}
            one["note"] = f"one eager batch-{B} solve, CUDA events around every launch (its {gm['ms']:.2f} ms of GEMM time exceed the timed step: not the timed region's occupancy)"
            for k in ("share_of_solve", "ms_per_solve_by_kernel"):
                roof.pop(k)
            roof.update({"achieved": gfl / (gms * 1e-3) / 1e12, "frac": gfl / (gms * 1e-3) / 1e12 / peak_tf, "avg_launch_us": gms * 1e3 / gln,
                         "launches_per_solve": gln, "algorithmic_flop_per_solve": gfl, "rows": BF * (T + 2),
                         "gemm_share_of_solve": round(gms / allms, 4),
                         "how": f"all tcgen05 GEMM / tail launches of one eager batch-{BF} solve (= the rows of the {F} solves in flight during the "
                                "timed region), CUDA events around every launch on the launching stream",
                         "one_solve": one})
        del muF, zF, maskF

    # ---- the dominant kernel on its own: block2's k3 conv (256 -> 256 channels) over the rows the F in-flight solves hold,
    #      launched back to back (PDL on, operands L2-warm), CUDA events around the train -> against the BURST peak ----
    if roof is not None:
        rows_k, reps_k = F * B * (T + 2), 50
        A_k = torch.randn(rows_k, 256, device=dev).half()
        W_k = (torch.randn(256, 768, device=dev) / 27.7).half()
        b_k = torch.randn(256, device=dev)
        o_k = torch.empty(rows_k, 256, dtype=torch.float16, device=dev)
        sh_k = (C.c_int * 3)(-1, 0, 1)
        with torch.cuda.stream(stream):
            def conv_once():
                _lib.check(eng.lib.mtts_debug_gemm(eng.h, A_k.data_ptr(), W_k.data_ptr(), b_k.data_ptr(), o_k.data_ptr(), rows_k, 256, 256,
                                                   3, sh_k, stream.cuda_stream))
            for _ in range(5):
                conv_once()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps_k):
                conv_once()
            e1.record(stream)
            e1.synchronize()
        us_k = e0.elapsed_time(e1) * 1e3 / reps_k
        tf_k = 2.0 * rows_k * 256 * 768 / us_k / 1e6
        burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1620.4) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1620.4
        roof["kernel_alone"] = {"kernel": "gemm_tc_kernel<256, EPI_PLAIN> as a k3 Conv1d 256->256", "rows": rows_k, "us_per_launch": us_k,
                                "achieved": tf_k, "peak": burst, "unit": "TFLOP/s", "frac": tf_k / burst,
                                "note": f"{reps_k} back-to-back launches over the {rows_k} rows of {F} in-flight solves; burst bf16 peak"}

    # ---- BASELINE config 5 folded into the line: 4096 multi-speaker utterances, bucketed, sharded over the ranks (strong
    #      scaling), the NCCL gather of the mels included, SHA-256 over all mels (equal for every N) ----
    config5 = None
    if not args.no_config5:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config5 as c5
        del lane_sets, lane_gather, mu_dv, mask_dv
        torch.cuda.empty_cache()
        try:
            config5 = c5.run(dev, world, rank, n_utt=args.config5_utts, max_frames=args.config5_frames, lanes=args.config5_lanes)
        except Exception as exc:                                   # never take the headline line down
            config5 = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- the whole reference call: MatchaTTS.synthesise(tokens, lengths, n_timesteps, temperature) -- native text encoder +
    #      duration predictor, alignment glue, native decoder -- tokens in pinned host memory, mels read back to the host ----
    synth = None
    if not args.no_synthesize:
        try:
            synth = run_synthesize(dev, B, n, max(5, args.steps // 2), world)
        except Exception as exc:
            synth = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- the step after the path: the HiFi-GAN generator on the batch one solve produces (rank 0, one GPU) ----
    voc = None
    if not args.no_vocoder and world == 1 and rank == 0:
        try:
            voc = run_vocoder(dev, B, T, max(5, args.steps // 2))
        except Exception as exc:
            voc = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    ref_gpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ref_gpu = reference_on_gpu(B, T, n, dev)

    if rank == 0:
        flop_step = B * T * n * (GEMM_FLOP_PER_FRAME_STEP + ATTN_FLOP_PER_FRAME_STEP_PER_T * T)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(B, T, n, steps=3, warmup=1, budget_s=25.0)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "config1": r["config1"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": workload_config(args, extra={
                "operands": "fp16 tensors, fp32 accumulate (bf16 operands miss the 1e-3 rel-L2 parity bar; same tensor rate)",
                "model_tflops_per_gpu": flop_step / (total_ms / args.steps * 1e-3) / 1e12,
                "in_flight_solves": F, "chains_per_solve": 1 if F > 1 else "heuristic (2)",
                "sm_share": "mtts_set_lanes(%d): persistent launches sized for 148 * 5/4 / %d SMs at most" % (F, F) if F > 1 else None,
                "l2": f"steps rotate over {F * NSET} distinct (mu, z0, z) sets = {set_bytes / 2**20:.0f} MiB > 126 MB L2, and every solve "
                      f"streams its own {eng.workspace(B, T)[2] / 2**20:.0f} MiB workspace; the serial figure flushes L2 (256 MiB write) between steps",
                "sustained": sustained,
                "serial": {"value": serial_value, "ms_per_step": serial_ms / args.steps, "ms_min": min(ms), "ms_max": max(ms),
                           "note": "one solve at a time (latency of a batch-64 solve), L2 flushed between steps"}}),
            "roofline": roof, "cpu_baseline": cpu, "config5": config5, "synthesize": synth, "vocoder": voc, "reference_on_gpu": ref_gpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": mu_h.numel() * 4 + mask_h.numel() * 4,
                    "d2h_bytes_per_step": out_h[0].numel() * 4, "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "CFM.forward(mu, mask, n_timesteps, temperature) per step, pinned-host mu/mask in and mel out per step; "
                           f"{F} steps in flight on {F} streams, copies of neighbouring steps overlap the solves (copy streams)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=344)
    ap.add_argument("--n-timesteps", type=int, default=10)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--in-flight", type=int, default=int(os.environ.get("MTTS_BENCH_INFLIGHT", "5")), help="independent solves (batches) in flight at a time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustained-steps", type=int, default=120, help="extra timed region of this many steps (0 = skip)")
    ap.add_argument("--no-synthesize", action="store_true", help="skip the tokens -> mel leg (MatchaTTS.synthesise)")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the mel -> waveform leg (hifigan.Generator)")
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE config 5 job folded into the line")
    ap.add_argument("--config5-utts", type=int, default=4096)
    ap.add_argument("--config5-frames", type=int, default=64 * 344, help="padded-frame budget per bucket")
    ap.add_argument("--config5-lanes", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
