"""Partitioning of the hot path over GPUs.

* utterance sharding (configs c3, c5): independent units, round-robin over ranks, no data-path collective;
* frame-range sharding of one long signal (config c2): contiguous blocks of shards per rank; the per-adaptation
  neighbour exchange of per-frame parameters and the global convergence decision live in longsignal.py.
"""
SPLINE_HALO_KNOTS = 32      # a cubic-spline knot influences its neighbours like 0.268^distance


def partition_round_robin(n_items, rank, world):
    """Indices of the items rank `rank` of `world` processes (utterance sharding)."""
    return list(range(rank, n_items, world))


def contiguous_blocks(n_items, world):
    """[(first, last+1)] per rank: contiguous blocks of near-equal size (frame-range sharding)."""
    base, rem = divmod(n_items, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        out.append((start, start + cnt))
        start += cnt
    return out


def stratified_chunk_order(n_items, chunk):
    """Positions of a cost-ordered list re-dealt so that consecutive chunks of `chunk` items are stratified samples of
    it: chunk b takes ranks b, b + n, b + 2n, ... (n = number of chunks), in rank order inside the chunk.  Used by the
    chunk pipeline of the analysis stage (EAQHM_PIPELINE=2): equal work per chunk, expensive frames first in each."""
    import numpy as np
    n = -(-n_items // chunk)
    if n < 2:
        return np.arange(n_items)
    perm = np.arange(n * chunk).reshape(chunk, n).T.ravel()
    return perm[perm < n_items]

