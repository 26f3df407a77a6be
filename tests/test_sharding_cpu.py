"""CPU: the host side of the multi-GPU partitioning, including a world_size-2 gloo run of the
per-adaptation halo exchange of per-frame parameters and of the statistics exchange that lets every rank take the
same convergence decision."""
import os
import socket

import numpy as np
import pytest

from eaqhm_b200 import corpus, engine, longsignal, sharding


def test_partitions_cover_everything_once():
    for n, w in ((10, 4), (3, 8), (10000, 8), (1, 1)):
        rr = [sharding.partition_round_robin(n, r, w) for r in range(w)]
        assert sorted(i for p in rr for i in p) == list(range(n))
        cb = sharding.contiguous_blocks(n, w)
        assert cb[0][0] == 0 and cb[-1][1] == n and all(a[1] == b[0] for a, b in zip(cb, cb[1:]))
        assert max(b - a for a, b in cb) - min(b - a for a, b in cb) <= 1


def test_stratified_chunk_order_is_a_balanced_permutation():
    """Chunk pipeline (EAQHM_PIPELINE=2): the cost-ordered frame list re-dealt into chunks that are stratified samples."""
    from eaqhm_b200.sharding import stratified_chunk_order
    rng = np.random.default_rng(3)
    for n_items, chunk in ((21270, 3545), (1000, 260), (17, 5), (10, 10), (7, 100)):
        perm = stratified_chunk_order(n_items, chunk)
        assert sorted(perm.tolist()) == list(range(n_items))
        cost = np.sort(rng.random(n_items) ** 3)[::-1]            # descending, like the engine's cost order
        c = cost[perm]
        nch = -(-n_items // chunk)
        sums = [c[b * chunk:(b + 1) * chunk].sum() for b in range(nch)]
        if nch >= 2 and n_items % chunk == 0:
            assert max(sums) / min(sums) < 1.05                    # equal work per chunk
            for b in range(nch):                                   # expensive frames first inside every chunk
                assert (np.diff(c[b * chunk:(b + 1) * chunk]) <= 0).all()
        if nch == 1:
            assert (perm == np.arange(n_items)).all()


def test_corpus_batches_respect_the_memory_budget():
    sizes, kmaxes = [48000, 48000, 16000, 96000, 48000, 1000], [60, 97, 40, 97, 50, 97]
    per = lambda n, k: (corpus.TRACK_ARRAYS * k + 4) * n * 8
    budget = per(48000 + 48000 + 16000, 97)
    plan = corpus.plan_batches(sizes, kmaxes, budget)
    assert plan[0][0] == 0 and plan[-1][1] == len(sizes) and all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    for a, b in plan:
        assert b - a == 1 or per(sum(sizes[a:b]), max(kmaxes[a:b])) <= budget
    assert plan[0] == (0, 3)
    assert corpus.plan_batches([10 ** 9], [97], 1) == [(0, 1)]          # an oversized utterance still gets a batch
    assert corpus.plan_batches([], [], 1) == []


def test_range_plan_alignment_and_cover():
    length, align = 16000 * 7 + 123, 240
    halo = longsignal.exchange_halo_samples(267, 15, 32, align)
    # window reach + 32 spline knots, independent of the number of adaptations; the excerpt-edge frames that the
    # reference's gating would drop (analysisWindow*step samples, functions.py:180) stay outside every owned window
    assert halo % align == 0 and halo >= 267 + 32 * 15 and halo - 267 >= 32 * 15
    for world in (1, 2, 3):
        plans = longsignal.plan_ranges(length, 16000 * 2, halo, align, world)
        flat = [p for r in plans for p in r]
        assert len(plans) == world and flat[0][0] == 0 and flat[-1][1] == length
        for (lo, hi, elo, ehi), nxt in zip(flat, flat[1:] + [None]):
            assert lo % align == 0 and elo % 15 == 0 and elo <= lo < hi <= ehi <= length
            assert elo == max(0, lo - halo) and ehi == min(length, hi + halo)
            if nxt is not None:
                assert hi == nxt[0]
    with pytest.raises(ValueError):
        longsignal.plan_ranges(length, 16000 * 4, halo, align, 8)


def test_excerpt_tables_are_slices_of_the_signal_tables():
    fs, step, fstep = 16000, 15, 80
    length = fs * 3 + 77
    t5 = np.arange(0, length - 1, fstep) / fs
    f0s = np.stack([t5, 150 + 20 * np.sin(2 * np.pi * t5), np.ones_like(t5)], axis=1)
    voiced5 = np.ones(len(np.arange(1, length, fstep)), bool)
    voiced5[100:140] = False
    full = engine.FrameTables(length, fs, f0s, voiced5, fstep, 62)
    for lo, hi in ((0, 16000), (14400, 33600), (31200, length)):
        e = full.excerpt(lo, hi)
        g = lo // step
        sl = slice(g, g + e.No_ti)
        assert e.g0 == g and (e.ti0 + lo == full.ti0[sl]).all() and e.ti0[0] == 0
        for name in ("f0", "K0", "N", "analysed", "isSpeech"):
            assert (getattr(e, name) == getattr(full, name)[sl]).all()      # bitwise, not recomputed
        assert e.f0stale == full.f0stale and e.own.sum() == hi - lo - e.ti0[0]
        if hi == length:
            assert g + e.No_ti == full.No_ti
    with pytest.raises(ValueError):
        full.excerpt(7, 16000)


def test_range_limited_tables_equal_the_signal_tables():
    """A rank of a long-signal run builds the per-frame tables of its own frame range only (host.FrameTables
    frame_range): same values as the slice of the whole signal's tables, the stale f0 of quirk Q1 found from the
    signal's end, excerpts cut from them identical."""
    fs, step, fstep = 16000, 15, 80
    length = fs * 4 + 33
    t5 = np.arange(0, length - 1, fstep) / fs
    f0s = np.stack([t5, 140 + 30 * np.sin(2 * np.pi * 0.7 * t5), np.ones_like(t5)], axis=1)
    voiced5 = np.ones(len(np.arange(1, length, fstep)), bool)
    voiced5[300:420] = False
    voiced5[-90:] = False                       # the last analysed frame lies well before the end (Q1 search)
    full = engine.FrameTables(length, fs, f0s, voiced5, fstep, 70)
    for g_lo, g_hi in ((0, 1500), (1200, 3000), (2900, full.No_ti), (-5, 10 ** 9)):
        part = engine.FrameTables(length, fs, f0s, voiced5, fstep, 70, frame_range=(g_lo, g_hi))
        a, b = max(0, g_lo), min(full.No_ti, g_hi)
        assert part.g_base == a and part.No_ti == b - a and part.No_ti_total == full.No_ti
        for name in ("f0", "K0", "N", "analysed", "isSpeech", "ti0", "own"):
            assert (getattr(part, name) == getattr(full, name)[a:b]).all(), name
        assert part.f0stale == full.f0stale and full.f0stale > 0
    part = engine.FrameTables(length, fs, f0s, voiced5, fstep, 70, frame_range=(1200, 3000))
    lo, hi = 1300 * step, 2800 * step
    e1, e2 = full.excerpt(lo, hi), part.excerpt(lo, hi)
    for name in ("f0", "K0", "N", "analysed", "isSpeech", "ti0", "own"):
        assert (getattr(e1, name) == getattr(e2, name)).all()
    assert e1.g0 == e2.g0 and e1.f0stale == e2.f0stale
    with pytest.raises(ValueError):
        part.excerpt(0, 16000)


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
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    merged = engine.merge_stats([r for part in gathered for r in part])
    first, last = sharding.contiguous_blocks(7, world)[rank]
    # neighbour exchange of per-frame parameter rows (longsignal.exchange_halo_rows) on CPU tensors over gloo:
    # every rank owns 50 frames, keeps 8 halo rows per side; row content = global frame index
    import torch
    H, own, K = 8, 50, 5
    f_lo, f_hi = rank * own, (rank + 1) * own
    s_lo, s_hi = max(0, f_lo - H), min(world * own, f_hi + H)
    gidx = torch.arange(s_lo, s_hi, dtype=torch.float64)
    store = [torch.zeros((s_hi - s_lo, K), dtype=torch.float64) for _ in range(3)] + [torch.zeros(s_hi - s_lo, dtype=torch.float64)]
    o0, o1 = f_lo - s_lo, f_hi - s_lo
    for i, t in enumerate(store):
        t[o0:o1] = (gidx[o0:o1] * (i + 1)).reshape((-1,) + (1,) * (t.dim() - 1))
    sent = longsignal.exchange_halo_rows(store, (o0, o1), H, rank, world, dist)
    ok = all(torch.equal(t, (gidx * (i + 1)).reshape((-1,) + (1,) * (t.dim() - 1)).expand_as(t)) for i, t in enumerate(store))
    # fixed-size all-gather of the (n, mean, M2) rows, padded with n = 0 rows (longsignal.allgather_rows)
    rt = torch.zeros((3, 3), dtype=torch.float64)
    rt[0] = torch.tensor(rows[0][1:], dtype=torch.float64)
    allr = longsignal.allgather_rows(rt, world, dist).numpy()
    m2 = engine.merge_stats([("x", r[0], r[1], r[2]) for r in allr if r[0] > 0])["x"]
    ok = ok and len(allr) == 3 * world and np.allclose(m2, merged[("g", "long")], rtol=1e-15, atol=0)
    # the SRER target's standard deviation from per-shard sums (longsignal.target_std): every rank gets the bits a single
    # rank gets from the same shard plan
    sig = np.random.default_rng(99).standard_normal(16000 * 3 + 77) * 0.1 + 0.01
    plan1 = longsignal.plan_ranges(len(sig), 16000, 960, 240, 1)
    planw = longsignal.plan_ranges(len(sig), 16000, 960, 240, world)
    std_w = longsignal.target_std(sig, planw, rank, world, dist, "cpu")
    std_1 = longsignal.target_std(sig, plan1, 0, 1, None, "cpu")
    ok = ok and std_w == std_1 and abs(std_w - np.std(sig)) < 1e-15
    q.put((rank, merged[("g", "long")], (first, last), bool(ok), int(sent)))
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
    for rank, (n, mean, m2), blk, halo_ok, sent in out:
        assert halo_ok and sent == 8 * (3 * 5 + 1) * 8       # one neighbour each: 8 rows of 3*K+1 doubles
        assert n == len(both) and abs(mean - both.mean()) < 1e-13 and abs(np.sqrt(m2 / n) - np.std(both)) < 1e-13
    assert [o[2] for o in out] == [(0, 4), (4, 7)]       # every rank: same decision inputs, disjoint shard blocks
