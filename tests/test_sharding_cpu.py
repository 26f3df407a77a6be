"""CPU: the host side of the multi-GPU partitioning, including a world_size-2 gloo run of the
statistics exchange that lets every shard take the same convergence decision."""
import os
import socket

import numpy as np
import pytest

from eaqhm_b200 import engine, sharding


def test_partitions_cover_everything_once():
    for n, w in ((10, 4), (3, 8), (10000, 8), (1, 1)):
        rr = [sharding.partition_round_robin(n, r, w) for r in range(w)]
        assert sorted(i for p in rr for i in p) == list(range(n))
        cb = sharding.contiguous_blocks(n, w)
        assert cb[0][0] == 0 and cb[-1][1] == n and all(a[1] == b[0] for a, b in zip(cb, cb[1:]))
        assert max(b - a for a, b in cb) - min(b - a for a, b in cb) <= 1


def test_shard_plan_alignment_and_cover():
    length, align = 16000 * 7 + 123, 240
    halo = sharding.halo_samples(267, 15, 10, 32, align)
    assert halo % align == 0 and halo >= 480 + 11 * (267 + 480)
    plan = sharding.plan_shards(length, 16000 * 2, halo, align)
    assert plan[0][0] == 0 and plan[-1][1] == length
    for (lo, hi, elo, ehi), nxt in zip(plan, plan[1:] + [None]):
        assert lo % align == 0 and elo % align == 0 and elo <= lo < hi <= ehi <= length
        assert elo == max(0, lo - halo) and ehi == min(length, hi + halo)
        if nxt is not None:
            assert hi == nxt[0]


def test_shard_utterances_share_grid_group_and_stale_f0():
    fs, step, fstep = 16000, 15, 80
    length = fs * 3
    t5 = np.arange(0, length - 1, fstep) / fs
    f0s = np.stack([t5, 150 + 20 * np.sin(2 * np.pi * t5), np.ones_like(t5)], axis=1)
    s = np.random.default_rng(0).standard_normal(length) * 0.1
    voiced5 = np.ones(len(np.arange(1, length, fstep)), bool)
    plan = sharding.plan_shards(length, fs, 2400, 240)
    utts, full = sharding.shard_utterances(s, fs, f0s, voiced5, fstep, 90.0, 62, plan, "sig")
    assert len(utts) == 3 and all(u.group == "sig" and u.f0stale == full.f0stale for u in utts)
    assert all(abs(u.group_std_det - np.std(s)) < 1e-15 for u in utts)
    for u, (lo, hi, elo, ehi) in zip(utts, plan):
        ft = engine.FrameTables(len(u.s), fs, u.f0s, u.voiced5, fstep, 62)
        g = (elo + ft.ti0)                                        # global centre of the excerpt's frames
        own = (g >= lo) & (g < hi)
        gi = g[own] // step
        # inside the owned range the excerpt sees the same frames, window lengths and harmonic counts
        assert (ft.analysed[own] == full.analysed[gi]).all()
        assert (ft.N[own] == full.N[gi]).all() and (ft.K0[own] == full.K0[gi]).all()
        assert np.abs(ft.f0[own] - full.f0[gi]).max() < 1e-10
        assert u.stat_range == (lo - elo, hi - elo)


def test_merge_stats_equals_numpy_std():
    rng = np.random.default_rng(1)
    parts = [rng.standard_normal(n) * sc + off for n, sc, off in ((1000, 1.0, 0.1), (10, 3.0, -2.0), (5000, 0.2, 0.0))]
    rows = [("g", len(p), p.mean(), ((p - p.mean()) ** 2).sum()) for p in parts] + [("other", 3, 1.0, 2.0)]
    acc = engine.merge_stats(rows)
    n, mean, m2 = acc["g"]
    allp = np.concatenate(parts)
    assert n == len(allp) and abs(mean - allp.mean()) < 1e-14 and abs(np.sqrt(m2 / n) - np.std(allp)) < 1e-13
    assert acc["other"] == (3, 1.0, 2.0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(10 + rank)
    data = rng.standard_normal(1000 + 500 * rank) + rank
    rows = [(("g", "long"), float(len(data)), float(data.mean()), float(((data - data.mean()) ** 2).sum()))]
    merged = engine.merge_stats(sharding.make_allgather(dist)(rows))
    first, last = sharding.contiguous_blocks(7, world)[rank]
    q.put((rank, merged[("g", "long")], (first, last)))
    dist.destroy_process_group()


def test_world_size_2_gloo_statistics_exchange():
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d0 = np.random.default_rng(10).standard_normal(1000)
    d1 = np.random.default_rng(11).standard_normal(1500) + 1
    both = np.concatenate([d0, d1])
    for rank, (n, mean, m2), blk in out:
        assert n == len(both) and abs(mean - both.mean()) < 1e-13 and abs(np.sqrt(m2 / n) - np.std(both)) < 1e-13
    assert [b for _, _, b in out] == [(0, 4), (4, 7)]       # every rank: same decision inputs, disjoint shard blocks
