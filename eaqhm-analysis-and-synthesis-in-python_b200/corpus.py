"""Utterance-sharded corpora (BASELINE configs c3, c5): independent utterances, dealt round-robin
to the ranks, analysed in memory-sized batches that adapt in lockstep on one GPU.  No data-path collective: every
utterance takes its own convergence decision (functions.py:394-396), ranks never wait for each other.

Host preprocessing of an utterance (wav read, SWIPE', voicing flags: everything `load(i)` does) runs in a small
thread pool a few utterances ahead of the GPU, so the device does not idle while the next batch is prepared; the
hand-over of a batch's results (wait for the device -> host copy, per-utterance host arrays) runs on a worker thread
while the next batch is already adapting (two alternating sets of pinned staging buffers).
"""
from concurrent.futures import ThreadPoolExecutor

from .engine import Engine
from .sharding import partition_round_robin

TRACK_ARRAYS = 5          # am, ph, fmc, cos ph, sin ph: the dense per-sample state of a resident batch


def utterance_bytes(u):
    """Device bytes an utterance adds to a resident batch: dense tracks + signal, target, two reconstructions."""
    n = len(u.s)
    return (TRACK_ARRAYS * u.Kmax + 4) * n * 8


def plan_batches(sizes, kmaxes, budget_bytes):
    """Greedy in-order packing: [(first, last+1)] such that the dense tracks of a batch (every utterance padded to
    the batch's largest Kmax) stay within `budget_bytes`; a single oversized utterance forms its own batch."""
    out, first, samples, kmax = [], 0, 0, 0
    for i, (n, k) in enumerate(zip(sizes, kmaxes)):
        new_k = max(kmax, k)
        if i > first and (TRACK_ARRAYS * new_k + 4) * (samples + n) * 8 > budget_bytes:
            out.append((first, i))
            first, samples, new_k = i, 0, k
        samples, kmax = samples + n, new_k
    if len(sizes) > first:
        out.append((first, len(sizes)))
    return out


def analyze_corpus(n_utterances, load, rank=0, world=1, batch_bytes=16 << 30, lookahead=64, workers=4,
                   step=15, maxAdpt=10, pitchPeriods=3, analysisWindow=32, **engine_kw):
    """Generator of (utterance index, engine.Result) for this rank's share of a corpus.

    `load(i) -> engine.Utterance` does the host preprocessing of utterance i (api.make_utterance on a wav, or a
    synthetic generator).  Utterances are analysed in the order of their index within the rank's share."""
    mine = partition_round_robin(n_utterances, rank, world)
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool, ThreadPoolExecutor(max_workers=1) as handover:
        pending = []                                   # [(index, future)] in order
        nxt = 0

        def top_up():
            nonlocal nxt
            while nxt < len(mine) and len(pending) < lookahead:
                pending.append((mine[nxt], pool.submit(load, mine[nxt])))
                nxt += 1
        top_up()
        previous, nbatch = None, 0                     # (batch, future of its results) still being handed over
        while pending:
            batch, kmax = [], 0
            while pending:
                i, fut = pending[0]
                u = fut.result()
                new_k = max(kmax, u.Kmax)
                total = sum(len(b.s) for _, b in batch) + len(u.s)
                # two batches are alive at a time (one adapting, one being handed over): each gets half the budget
                if batch and (TRACK_ARRAYS * new_k + 4) * total * 8 > batch_bytes // 2:
                    break
                batch.append((i, u))
                kmax = new_k
                pending.pop(0)
                top_up()
            eng = Engine([u for _, u in batch], step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods,
                         analysisWindow=analysisWindow, staging_set=nbatch % 2, **engine_kw)
            eng.prepare_readback()
            out = eng.run_device(prefetch_results=True)
            fut = handover.submit(eng.results, *out)
            del eng
            if previous is not None:
                for (i, _), r in zip(previous[0], previous[1].result()):
                    yield i, r
            previous, nbatch = (batch, fut), nbatch + 1
        if previous is not None:
            for (i, _), r in zip(previous[0], previous[1].result()):
                yield i, r
