"""Host-side batching: length bucketing, rank assignment and the final gather.
The multi-rank path runs here on CPU with the gloo backend (world_size 2) and a stub solver: the
scheduling / gather logic under test is exactly what the GPU ranks execute around the native solver."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matcha_tts_b200 import batching as Bt
from matcha_tts_b200.model import fix_len_compatibility


def _lengths(n, seed, lo=64, hi=800):
    rng = random.Random(seed)
    return [max(lo, min(hi, int(rng.lognormvariate(5.7, 0.45)))) for _ in range(n)]


@pytest.mark.parametrize("n,seed", [(1, 0), (7, 1), (300, 2), (4096, 6)])
def test_buckets_cover_every_utterance_once(n, seed):
    lengths = _lengths(n, seed)
    buckets = Bt.make_buckets(lengths, max_frames=64 * 344)
    seen = sorted(i for b in buckets for i in b.indices)
    assert seen == list(range(n))
    for b in buckets:
        longest = max(lengths[i] for i in b.indices)
        assert b.t_max == fix_len_compatibility(longest) and b.t_max % 4 == 0      # reference rounding (model.py:49-55)
        assert b.padded_frames <= max(64 * 344, b.t_max)
    # sorted by length: padding waste stays small
    if n >= 300:
        valid = sum(lengths)
        padded = sum(b.padded_frames for b in buckets)
        assert padded <= (1.03 if n >= 4096 else 1.2) * valid


def test_buckets_are_deterministic_and_independent_of_world_size():
    lengths = _lengths(500, 3)
    b1 = Bt.make_buckets(lengths)
    b2 = Bt.make_buckets(list(lengths))
    assert b1 == b2
    for world in (1, 2, 4, 8):
        parts = Bt.assign_buckets(b1, world)
        assert sorted(j for p in parts for j in p) == list(range(len(b1)))
        loads = [sum(b1[j].cost for j in p) for p in parts]
        if world > 1 and len(b1) >= 4 * world:
            assert max(loads) <= 1.25 * (sum(loads) / world)                       # LPT keeps ranks balanced


def test_edge_cases():
    assert Bt.make_buckets([]) == []
    assert Bt.make_buckets([5]) == [Bt.Bucket((0,), 8)]
    with pytest.raises(ValueError):
        Bt.make_buckets([10, 0])
    # one utterance longer than the frame budget still gets its own bucket
    bs = Bt.make_buckets([5000, 10, 10], max_frames=1024)
    assert bs[0] == Bt.Bucket((0,), 5000)
    # ties keep input order
    assert Bt.make_buckets([8, 8, 8], max_frames=16)[0].indices == (0, 1)
    mus = [torch.randn(80, 5), torch.randn(80, 8)]
    mu, mask = Bt.pad_batch(mus, Bt.Bucket((1, 0), 8))
    assert mu.shape == (2, 80, 8) and mask[1, 0].tolist() == [1] * 5 + [0] * 3
    assert torch.equal(mu[1, :, :5], mus[0]) and float(mu[1, :, 5:].abs().max()) == 0.0


def _stub_solver(mu, mask, spks, bucket):
    """Depends on the whole batch (like the real decoder): row result = mu * mask + batch mean."""
    out = mu * mask + mu.mean()
    if spks is not None:
        out = out + spks.mean(1)[:, None, None]
    return out


def _worker(rank, world, port, lengths, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)
    mus = [torch.randn(80, n, generator=g) for n in lengths]
    spks = [torch.randn(64, generator=g) for _ in lengths]
    out = Bt.solve_sharded(mus, _stub_solver, spks=spks, max_frames=4096)
    if rank == 0:
        ret.update({i: v for i, v in out.items()})
    # gather="rank0": only rank 0 holds the gathered mels; the lazy dict cuts its views on access
    out0 = Bt.solve_sharded(mus, _stub_solver, spks=spks, max_frames=4096, gather="rank0")
    ret[f"rank0_mode_len_{rank}"] = len(out0)
    if rank == 0:
        ret["rank0_mode_equal"] = all(torch.equal(out0[i], out[i]) for i in range(len(lengths))) and sorted(out0) == sorted(out)
        ret["lazy_items"] = all(v.shape == (80, lengths[i]) for i, v in out0.items()) and len(out0.values()) == len(lengths)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_matches_single_process():
    lengths = _lengths(37, 5, lo=16, hi=300)
    g = torch.Generator().manual_seed(11)
    mus = [torch.randn(80, n, generator=g) for n in lengths]
    spks = [torch.randn(64, generator=g) for _ in lengths]
    single = Bt.solve_sharded(mus, _stub_solver, spks=spks, max_frames=4096)
    assert sorted(single) == list(range(len(lengths)))
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), lengths, ret), nprocs=2, join=True)
        multi = dict(ret)
    assert multi.pop("rank0_mode_len_0") == len(lengths) and multi.pop("rank0_mode_len_1") == 0
    assert multi.pop("rank0_mode_equal") and multi.pop("lazy_items")
    assert sorted(multi) == sorted(single)
    for i in single:
        assert multi[i].shape == (80, lengths[i])
        assert torch.equal(multi[i], single[i])          # sharding must not change any utterance's result


def test_index_plan_reproduces_padding_and_compaction():
    """The CUDA path of solve_sharded pads with one gather and compacts with another (indices from _index_plan):
    on CPU tensors the same indices must reproduce pad_batch exactly and invert it on the valid frames."""
    import numpy as np
    g = torch.Generator().manual_seed(9)
    lengths = [int(x) for x in torch.randint(1, 90, (37,), generator=g)]
    mus = [torch.randn(5, n, generator=g) for n in lengths]
    buckets = Bt.make_buckets(lengths, max_frames=4 * 88, max_batch=6)
    for world in (1, 3):
        for mine in Bt.assign_buckets(buckets, world):
            local_ids = [i for bid in mine for i in buckets[bid].indices]
            if not local_ids:
                continue
            lens = np.asarray([lengths[i] for i in local_ids], dtype=np.int64)
            offs = np.concatenate([[0], np.cumsum(lens)])
            total = int(offs[-1])
            table = torch.cat([mus[i] for i in local_ids] + [torch.zeros(5, 1)], dim=1)
            plan, flat = Bt._index_plan(buckets, mine, lens, offs)
            flat = torch.from_numpy(flat)
            arena = torch.full((total, 5), float("nan"))
            for bk, r0, i0, i1, i2, f0, nv in plan:
                B, T = len(bk.indices), bk.t_max
                idx = flat[i0:i1]
                mu = table.index_select(1, idx).view(5, B, T).permute(1, 0, 2).contiguous()
                mask = (idx != total).float().view(B, 1, T)
                mu_ref, mask_ref = Bt.pad_batch(mus, bk)
                assert torch.equal(mu, mu_ref) and torch.equal(mask, mask_ref)
                assert local_ids[r0:r0 + B] == list(bk.indices)
                comp = mu.permute(0, 2, 1).reshape(B * T, 5).index_select(0, flat[i1:i2])
                assert comp.shape[0] == nv
                arena[f0:f0 + nv] = comp
            for j, i in enumerate(local_ids):
                assert torch.equal(arena[int(offs[j]):int(offs[j + 1])].t(), mus[i])


def test_result_arena_pool_recycles_and_stays_bounded(monkeypatch):
    """MelDict.release() hands the pinned result arena back; the pool reuses the smallest arena that fits, walks views
    back to the whole allocation and never keeps more than _ARENA_POOL_BYTES (pinning stubbed out: no GPU here)."""
    orig = torch.empty

    def unpinned(*a, **k):
        k.pop("pin_memory", None)
        return orig(*a, **k)

    monkeypatch.setattr(torch, "empty", unpinned)
    monkeypatch.setattr(Bt, "_PINNED", {})
    a = Bt._pinned_arena(3_000_000)
    assert a.numel() == 3_000_000 and a._base.numel() == 3 << 20
    d = Bt.MelDict({0: a.view(-1, 80)[:10].t()})
    d.arena = a.view(-1, 80)
    d.release()
    assert len(d) == 0 and d.arena is None and len(Bt._PINNED[("arena", 3 << 20)]) == 1
    d.release()                                                   # idempotent
    b = Bt._pinned_arena(2_900_000)                               # similar size: recycled
    assert b._base.numel() == 3 << 20 and not Bt._PINNED[("arena", 3 << 20)]
    c = Bt._pinned_arena(100)                                     # much smaller: its own arena, not a slice of 12 MB
    assert c._base.numel() == 1 << 20
    Bt._release_arena(b)
    Bt._release_arena(c)
    assert Bt._pinned_arena(500_000)._base.numel() == 1 << 20
    monkeypatch.setattr(Bt, "_ARENA_POOL_BYTES", 13 << 20)
    Bt._release_arena(Bt._pinned_arena(900_000))                  # 12 MB + 4 MB pooled > 13 MB: the oldest goes
    assert sum(k[1] * 4 * len(v) for k, v in Bt._PINNED.items() if k[0] == "arena") <= 13 << 20
