"""BASELINE config 5: multi-speaker (VCTK-shape, in_channels 224) decoder, 4096 utterances with clipped log-normal
lengths, length-bucketed and sharded over the ranks of one node (one process per GPU, no hot-path collective; ONE NCCL
all-gather of the finished mels at the end).
    python tools/config5.py [n_utt] [max_frames] [lanes]                       (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config5.py [n_utt] ...
Prints one JSON line: valid and padded mel-frames/s of the whole job (bucketing, padding, solves on `lanes` lanes, the
gather) and a SHA-256 over every utterance's mel -- the same for 1 / 2 / 4 / 8 GPUs: bucket
composition is a function of the length list alone and the noise of a bucket is seeded by the bucket.
bench.py folds the same run into its JSON line as `config5`."""
import hashlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(dev, world, rank, n_utt=4096, max_frames=64 * 344, lanes=5, passes=3, with_hash=True):
    from matcha_tts_b200 import CFM, Decoder, batching
    import torch.distributed as dist
    g = torch.Generator().manual_seed(6)
    lengths = torch.exp(torch.randn(n_utt, generator=g) * 0.45 + 5.7).clamp(64, 800).long().tolist()   # median ~300
    torch.manual_seed(0)
    dec = Decoder(in_channels=224, out_channels=80, channels=(256, 256), num_heads=2, num_mid_blocks=2).to(dev)
    dec.set_lanes(lanes)                  # `lanes` buckets in flight: no split inside a solve, every launch takes its share of the SMs
    cfm = CFM(80, {"solver": "euler", "sigma_min": 1e-4}, n_spks=109, spk_emb_dim=64, estimator=dec)
    gd = torch.Generator(device=dev).manual_seed(5)
    mus = [torch.randn(80, n, generator=gd, device=dev) for n in lengths]
    spks = [torch.randn(64, generator=gd, device=dev) for _ in lengths]
    buckets = batching.make_buckets(lengths, max_frames)
    mine = batching.assign_buckets(buckets, world)[rank]

    def solver(mu, mask, s, bk):
        torch.manual_seed(1000 + bk.indices[0])                  # the bucket's noise does not depend on the rank it lands on
        return cfm(mu, mask, 10, temperature=0.667, spks=s)

    def one_pass(mode):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = batching.solve_sharded(mus, solver, spks, max_frames=max_frames, device=dev, gather=mode, lanes=lanes)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)             # slowest rank
        return float(t.item()), out

    # mode "device": every rank ends with every utterance's mel in its HBM (the NCCL gather); "rank0" adds the copy of all
    # mels to rank 0's host memory.  With one rank there is nothing to gather: the mels land in pinned host memory.
    modes = ["device", "rank0"] if world > 1 else [False]
    res = {}
    digest = None
    for mode in modes:
        _, warm = one_pass(mode)                 # warm-up: captures one CUDA graph per (lane, bucket shape)
        if hasattr(warm, "release"):
            warm.release()
        dts = []
        for rep in range(passes):
            dt, out = one_pass(mode)
            dts.append(dt)
            last = rep == passes - 1
            if last and with_hash and rank == 0 and mode in ("rank0", False):
                hsh = hashlib.sha256()
                finite = True
                for i in range(n_utt):
                    v = out[i].contiguous()
                    finite = finite and bool(torch.isfinite(v).all())
                    hsh.update(v.numpy().tobytes())
                digest = hsh.hexdigest()
                res["finite"] = finite
            if hasattr(out, "release"):
                out.release()
            del out
        res[str(mode)] = {"seconds": sorted(dts)[len(dts) // 2], "seconds_all_passes": dts}
    valid = sum(lengths)
    padded = sum(b.padded_frames for b in buckets)
    head = res["device"] if world > 1 else res["False"]
    line = {"config": "BASELINE configs[4]: VCTK-shape decoder (in_channels 224), %d utterances, lengths clipped log-normal [64, 800] "
                      "median %d, bucketed (<= %d padded frames per batch), %d solve lanes per GPU, strong scaling over the ranks"
                      % (n_utt, sorted(lengths)[n_utt // 2], max_frames, lanes),
            "n_gpus": world, "buckets": len(buckets), "buckets_rank0": len(mine), "seconds": head["seconds"],
            "valid_frames_per_s": valid / head["seconds"], "padded_frames_per_s": padded / head["seconds"],
            "padding_overhead": padded / valid - 1.0, "mel_sha256": digest,
            "gather": ("one ncclAllGather (all_gather_into_tensor) of the compacted frames arenas: every rank holds all mels in HBM"
                       if world > 1 else "single rank: mels copied to pinned host memory per bucket"),
            "timing": res}
    if world > 1:
        line["with_rank0_host_copy"] = {"seconds": res["rank0"]["seconds"], "valid_frames_per_s": valid / res["rank0"]["seconds"]}
    del dec, cfm
    return line


def main():
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    max_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64 * 344
    lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    line = run(dev, world, rank, n_utt, max_frames, lanes)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
