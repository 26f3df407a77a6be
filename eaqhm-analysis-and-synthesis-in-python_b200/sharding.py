"""Partitioning of the hot path over GPUs.

* utterance sharding (configs c3, c5): independent units, round-robin over ranks, no data-path collective;
* frame-range sharding of one long signal (config c2): contiguous sample ranges with a halo.  Frames are
  independent inside an adaptation, but adaptation a+1 of a frame reads the tracks of adaptation a over its
  window (+-N_max samples) and the splines couple ~32 knots (0.268^32 influence), so the region a
  shard computes correctly shrinks by N_max + 32*step samples per adaptation.  Round 1 gives every shard a halo
  deep enough for all adaptations and exchanges only the SRER statistics (n, mean, M2 per shard) so that all
  shards take the reference's single global convergence decision (functions.py:388,394-396).
"""
import numpy as np

from .engine import Engine, Result, Utterance
from .host import FrameTables

SPLINE_HALO_KNOTS = 32


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


def halo_samples(N_max, step, maxAdpt, analysisWindow, align):
    """Halo that keeps a shard's owned range exact through maxAdpt+1 adaptations."""
    h = analysisWindow * step + (maxAdpt + 1) * (N_max + SPLINE_HALO_KNOTS * step)
    return int(-(-h // align) * align)


def plan_shards(length, shard_samples, halo, align):
    """[(own_lo, own_hi, ex_lo, ex_hi)]: owned sample range and the excerpt (owned + halo) of every shard.
    Owned ranges and excerpt starts are multiples of `align` = lcm(step, frame_step) so that the frame grid
    (functions.py:148) and the 5 ms f0 grid (functions.py:113,155) of an excerpt coincide with the signal's."""
    shard_samples = max(align, int(shard_samples) // align * align)
    plan, lo = [], 0
    while lo < length:
        hi = min(length, lo + shard_samples)
        if length - hi < shard_samples // 2:      # avoid a tiny tail shard
            hi = length
        plan.append((lo, hi, max(0, lo - halo), min(length, hi + halo)))
        lo = hi
    return plan


def shard_utterances(s, fs, f0s, voiced5, frame_step, f0min, Kmax, plan, group, det_part=None, step=15,
                     pitchPeriods=3, analysisWindow=32):
    """Excerpt Utterances for `plan`, all in one SRER group, with the whole signal's target std and stale f0."""
    s = np.asarray(s, float).reshape(-1)
    det = s if det_part is None else np.asarray(det_part, float).reshape(-1)
    full = FrameTables(len(s), fs, f0s, voiced5, frame_step, Kmax, step, pitchPeriods, analysisWindow)
    std_det = float(np.std(det))
    out = []
    for own_lo, own_hi, ex_lo, ex_hi in plan:
        j0 = ex_lo // frame_step
        out.append(Utterance(s=s[ex_lo:ex_hi], fs=fs, f0s=np.asarray(f0s)[j0:], voiced5=np.asarray(voiced5)[j0:],
                             frame_step=frame_step, f0min=f0min, Kmax=Kmax,
                             det_part=None if det_part is None else det[ex_lo:ex_hi], group=group,
                             stat_range=(own_lo - ex_lo, own_hi - ex_lo), group_std_det=std_det,
                             f0stale=full.f0stale, xoff=ex_lo))
    return out, full


def stitch(results, plan, full, step=15):
    """Owned ranges of the shard results -> one Result for the whole signal."""
    length = full.length
    Kmax = results[0].am.shape[1]
    s_recon = np.zeros(length)
    a0 = np.zeros(full.No_ti)
    am = np.zeros((full.No_ti, Kmax))
    fm = np.zeros((full.No_ti, Kmax))
    ph = np.zeros((full.No_ti, Kmax))
    for r, (own_lo, own_hi, ex_lo, ex_hi) in zip(results, plan):
        s_recon[own_lo:own_hi] = r.s_recon[own_lo - ex_lo:own_hi - ex_lo]
        g_lo, g_hi = -(-own_lo // step), -(-own_hi // step)           # frames with centre in [own_lo, own_hi)
        l_lo = g_lo - ex_lo // step
        n = g_hi - g_lo
        a0[g_lo:g_hi] = r.a0[l_lo:l_lo + n]
        am[g_lo:g_hi] = r.am[l_lo:l_lo + n]
        fm[g_lo:g_hi] = r.fm[l_lo:l_lo + n]
        ph[g_lo:g_hi] = r.ph[l_lo:l_lo + n]
    r0 = results[0]
    return Result(s_recon=s_recon, SRER=list(r0.SRER), ti=full.ti0.copy(), isSpeech=full.isSpeech.copy(),
                  isVoiced=full.analysed.copy(), a0=a0, am=am, fm=fm, ph=ph, n_adapt_best=r0.n_adapt_best,
                  stats=dict(shards=len(plan)))


def make_allgather(dist):
    """stats_allgather callback for Engine over a torch.distributed process group (NCCL or gloo)."""
    def gather(rows):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, rows)
        return [r for part in out for r in part]
    return gather


def analyze_long_signal(s, fs, f0s, voiced5, frame_step, f0min, Kmax, shard_seconds=20.0, det_part=None, step=15,
                        maxAdpt=10, pitchPeriods=3, analysisWindow=32, rank=0, world=1, dist=None, **engine_kw):
    """Frame-range sharded analysis of one long signal.  Every rank analyses a contiguous block of shards
    (all resident, adapting in lockstep); the SRER statistics are all-gathered so that every shard on every
    rank stops at the same adaptation.  Returns (Result for this rank's owned ranges stitched into a
    full-length Result, plan, (first, last+1) shard block of this rank)."""
    align = int(np.lcm(step, frame_step))
    full = FrameTables(len(np.asarray(s).reshape(-1)), fs, f0s, voiced5, frame_step, Kmax, step, pitchPeriods,
                       analysisWindow)
    halo = halo_samples(int(full.N.max()), step, maxAdpt, analysisWindow, align)
    plan = plan_shards(full.length, int(shard_seconds * fs), halo, align)
    utts, full = shard_utterances(s, fs, f0s, voiced5, frame_step, f0min, Kmax, plan, "long", det_part, step,
                                  pitchPeriods, analysisWindow)
    first, last = contiguous_blocks(len(plan), world)[rank]
    mine = utts[first:last]
    gather = make_allgather(dist) if (dist is not None and world > 1) else None
    eng = Engine(mine, step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods, analysisWindow=analysisWindow,
                 stats_allgather=gather, **engine_kw)
    res = eng.run()
    return stitch(res, plan[first:last], full, step), plan, (first, last)
