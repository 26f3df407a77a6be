"""One long signal as contiguous frame ranges (BASELINE config c2): streamed over the shards of a
GPU and split over GPUs with a per-adaptation halo exchange of per-frame parameters.

What persists between adaptations is only the per-frame parameter store (amplitude, phase, frequency of every
harmonic and the DC value at every frame centre: 3*Kmax+1 doubles per frame, frame-major); the dense per-sample
tracks of the reference (functions.py:159-160,169-171; length x Kmax x 3 doubles, 134 GB for the 1-hour signal) are
rebuilt shard by shard from that store whenever a stage needs them.  One adaptation a (functions.py:163-402) is

  pass 1, every shard:  tracks(a-1) <- interpolate(store_cur over the shard's excerpt)          [a > 0]
                        analysis(a) of the frames the shard owns  -> rows of store_new           (functions.py:176-329)
  exchange:             the first / last `halo` owned rows of store_new go to the previous / next rank, which
                        keeps them as its halo rows (torch.distributed send/recv: NCCL over NVLink on GPUs)
  pass 2, every shard:  tracks(a) <- interpolate(store_new over the excerpt)                     (functions.py:337-383)
                        s_hat, residual statistics over the owned samples                        (functions.py:385-388)
                        centre values of adaptation a -> rows of the candidate store             (functions.py:397-402)
  decision:             (n, mean, M2) of all shards of all ranks are all-gathered and merged -> the signal's SRER;
                        candidate becomes best unless SRER[a] <= SRER[a-1], which stops every rank (functions.py:394-396)

An excerpt is the owned range plus `halo` samples on each side: halo >= N_max + 34*step, because an owned frame's
window reaches N_max samples into the neighbour (functions.py:193,196) and the cubic splines couple ~32 knots
(0.268^32 influence).  The halo does not depend on the number of adaptations and no frame is analysed
twice.  With one shard per rank the tracks built in pass 2 are reused by pass 1 of the next adaptation.
"""
import numpy as np
import torch

from . import _lib
from .engine import (BufferPool, Engine, Result, Utterance, _parallel_copies, _parallel_fill, _pinned, _to_host,
                     check_not_pd, merge_stats)
from .host import FrameTables
from .sharding import SPLINE_HALO_KNOTS, contiguous_blocks

NROW = 4      # stores: am, ph, fm [rows][Kmax] + a0 [rows]


def exchange_halo_samples(N_max, step, analysisWindow, align):
    """Halo of an excerpt: window reach + spline coupling, and wide enough that the excerpt-edge frames the local
    gating would drop (functions.py:180) are never inside an owned window."""
    h = N_max + (SPLINE_HALO_KNOTS + 2) * step
    return int(-(-h // align) * align)


def plan_ranges(length, shard_samples, halo, align, world):
    """Per rank: list of (own_lo, own_hi, ex_lo, ex_hi).  Owned ranges tile [0, length), are multiples of `align`
    (a multiple of step, so excerpts start on the frame grid) and every rank gets a contiguous block of them."""
    shard_samples = max(align, int(shard_samples) // align * align)
    own, lo = [], 0
    while lo < length:
        hi = min(length, lo + shard_samples)
        if length - hi < shard_samples // 2:
            hi = length
        own.append((lo, hi))
        lo = hi
    if len(own) < world:
        raise ValueError("fewer shards (%d) than ranks (%d): lower shard_seconds" % (len(own), world))
    out = []
    for first, last in contiguous_blocks(len(own), world):
        out.append([(a, b, max(0, a - halo), min(length, b + halo)) for a, b in own[first:last]])
    return out


def exchange_halo_rows(store, own_rows, halo_rows, rank, world, dist):
    """The per-adaptation neighbour exchange: `store` is a list of tensors whose dim 0 indexes this rank's frames
    [first owned - halo_rows, last owned + halo_rows) (clipped at the signal ends); rows [own_rows[0], own_rows[1])
    are owned.  The first halo_rows owned rows go to rank-1 and land in its right halo; the last ones go to rank+1.
    Works on whatever device the store lives on (NCCL for CUDA tensors, gloo for the CPU test)."""
    if dist is None or world == 1:
        return 0
    o0, o1 = own_rows
    H = halo_rows
    ops, recvs, sent = [], [], 0
    width = sum(int(np.prod(t.shape[1:])) for t in store)

    def pack(r0, r1):
        return torch.cat([t[r0:r1].reshape(r1 - r0, -1) for t in store], dim=1).contiguous()

    def unpack(buf, r0):
        c = 0
        for t in store:
            w = int(np.prod(t.shape[1:]))
            t[r0:r0 + buf.shape[0]] = buf[:, c:c + w].reshape((buf.shape[0],) + tuple(t.shape[1:]))
            c += w
    if rank > 0:
        out = pack(o0, o0 + H)
        inc = torch.empty((H, width), dtype=out.dtype, device=out.device)
        ops += [dist.P2POp(dist.isend, out, rank - 1), dist.P2POp(dist.irecv, inc, rank - 1)]
        recvs.append((inc, o0 - H))
        sent += out.numel() * out.element_size()
    if rank < world - 1:
        out = pack(o1 - H, o1)
        inc = torch.empty((H, width), dtype=out.dtype, device=out.device)
        ops += [dist.P2POp(dist.isend, out, rank + 1), dist.P2POp(dist.irecv, inc, rank + 1)]
        recvs.append((inc, o1))
        sent += out.numel() * out.element_size()
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for inc, r0 in recvs:
        unpack(inc, r0)
    return sent


def allgather_rows(rows, world, dist):
    """All ranks' (n, mean, M2) rows in rank order: one fixed-size all-gather (NCCL on GPUs, gloo in the CPU test)."""
    if dist is None or world == 1:
        return rows
    parts = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(parts, rows.contiguous())
    return torch.cat(parts, dim=0)


def target_std(det, plan_all, rank, world, dist, device):
    """numpy.std of the whole SRER target (functions.py:161) without any rank touching more than its own samples: (n,
    mean, M2) of every shard's owned range, all-gathered once (fixed size, padded with n = 0 rows) and merged in shard
    order with the pairwise-variance formula - the same arithmetic, hence the same bits, for any number of ranks."""
    nmax = max(len(p) for p in plan_all)
    st = np.zeros((nmax, 3))
    for j, (own_lo, own_hi, _, _) in enumerate(plan_all[rank]):
        x = det[own_lo:own_hi]
        mu = x.mean()
        st[j] = (own_hi - own_lo, mu, ((x - mu) ** 2).sum())
    st = allgather_rows(torch.from_numpy(st).to(device), world, dist).cpu().numpy()
    n_all, _, m2_all = merge_stats([("t", r[0], r[1], r[2]) for r in st if r[0] > 0])["t"]
    return float(np.sqrt(m2_all / n_all))


class LongSignalEngine:
    """eaQHM adaptation loop of one long signal on this rank's contiguous frame range (see the module docstring)."""

    def __init__(self, s, fs, f0s, voiced5, frame_step, f0min, Kmax, shard_seconds=20.0, det_part=None, step=15,
                 maxAdpt=10, pitchPeriods=3, analysisWindow=32, rank=0, world=1, dist=None, device="cuda",
                 full_tables=None, **engine_kw):
        if not torch.cuda.is_available():
            raise _lib.EaqhmError("no CUDA device: the eaQHM B200 path has no CPU fallback")
        self.device = torch.device(device)
        self.rank, self.world, self.dist = int(rank), int(world), dist
        self.fs, self.step, self.maxAdpt, self.Kmax = fs, int(step), int(maxAdpt), int(Kmax)
        s = np.asarray(s, float).reshape(-1)
        det = s if det_part is None else np.asarray(det_part, float).reshape(-1)
        length, total = len(s), len(range(1, len(s), self.step))
        # Everything a rank sets up is local to its range: the halo follows from a bound on the half window that every
        # rank derives alike from the f0 grid (the f0 of a frame is a convex combination of two voiced grid values,
        # functions.py:185), the per-frame tables cover the rank's frames only, the target's standard deviation comes
        # from per-shard sums exchanged below.
        f0g = np.asarray(f0s, float)[:, 1]
        m = min(len(f0g), len(voiced5))
        vg = f0g[:m][np.asarray(voiced5, bool)[:m] & np.isfinite(f0g[:m]) & (f0g[:m] > 0)]
        N_bound = int(max(120, np.rint((pitchPeriods / 2) * (fs / vg.min())))) if len(vg) else 120     # functions.py:191
        if full_tables is not None:
            N_bound = int(full_tables.N.max())
        align = int(np.lcm(self.step, int(frame_step)))
        self.halo = exchange_halo_samples(N_bound, self.step, analysisWindow, align)
        self.H = self.halo // self.step
        self.plan_all = plan_ranges(length, int(shard_seconds * fs), self.halo, align, self.world)
        self.plan = self.plan_all[self.rank]
        self.R_lo, self.R_hi = self.plan[0][0], self.plan[-1][1]
        cdiv = lambda a, b: -(-a // b)
        self.F_lo, self.F_hi = cdiv(self.R_lo, self.step), min(total, cdiv(self.R_hi, self.step))
        self.S_lo, self.S_hi = max(0, self.F_lo - self.H), min(total, self.F_hi + self.H)
        self.X_lo, self.X_hi = self.plan[0][2], self.plan[-1][3]          # resident samples: owned range + halo
        # frames of the resident samples (a superset of the store rows: the excerpt tables are sliced from these)
        err = None
        try:
            if self.world > 1 and self.F_hi - self.F_lo < self.H:
                raise ValueError("a rank must own at least the halo (%d frames)" % self.H)
            self.full = full_tables if full_tables is not None else FrameTables(
                length, fs, f0s, voiced5, frame_step, Kmax, step, pitchPeriods, analysisWindow,
                frame_range=(self.X_lo // self.step, cdiv(self.X_hi, self.step)))
            if len(self.full.wrapping_frames()):
                raise ValueError("%d analysed frame(s) have a window that starts before sample 0 (the reference wraps it "
                                 "to the end of the signal): not supported by the frame-range engine; use engine.Engine, or "
                                 "raise analysisWindow so that analysisWindow*step >= the largest half window (%d samples)"
                                 % (len(self.full.wrapping_frames()), int(self.full.N.max())))
        except (ValueError, IndexError) as e:
            err = e
        if self.world > 1:
            # an input error on one rank (window past the end of the signal: the last rank only) must stop all of them
            flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if err is None and int(flag.item()):
                err = ValueError("another rank rejected its range of the signal")
        if err is not None:
            raise err
        full = self.full
        rows = self.S_hi - self.S_lo
        d, K = self.device, self.Kmax
        self.h2d_bytes = 2 * (self.X_hi - self.X_lo) * 8
        self.d2h_bytes = 0
        self.s_dev = torch.from_numpy(np.ascontiguousarray(s[self.X_lo:self.X_hi])).to(d)
        self.det_dev = self.s_dev if det_part is None else torch.from_numpy(
            np.ascontiguousarray(det[self.X_lo:self.X_hi])).to(d)
        self.std_det = target_std(det, self.plan_all, self.rank, self.world, self.dist, d)        # functions.py:161
        f64 = dict(dtype=torch.float64, device=d)

        def store():
            return [torch.zeros((rows, K), **f64) for _ in range(3)] + [torch.zeros(rows, **f64)]
        self.cur, self.new, self.cand, self.best = store(), store(), store(), store()
        self.s_hat = torch.zeros(self.R_hi - self.R_lo, **f64)
        self.s_best = torch.zeros(self.R_hi - self.R_lo, **f64)
        self.pool = BufferPool(d)
        self.shards = []
        for own_lo, own_hi, ex_lo, ex_hi in self.plan:
            ft = full.excerpt(ex_lo, ex_hi)
            u = Utterance(s=self.s_dev[ex_lo - self.X_lo:ex_hi - self.X_lo], fs=fs, f0s=None, voiced5=None,
                          frame_step=frame_step, f0min=f0min, Kmax=K,
                          det_part=None if det_part is None else self.det_dev[ex_lo - self.X_lo:ex_hi - self.X_lo],
                          stat_range=(own_lo - ex_lo, own_hi - ex_lo), group_std_det=self.std_det,
                          f0stale=full.f0stale, xoff=ex_lo, own_range=(own_lo - ex_lo, own_hi - ex_lo))
            e = Engine([u], step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods, analysisWindow=analysisWindow,
                       device=d, frame_tables=[ft], pool=self.pool, **engine_kw)
            e.set_active([True])
            g_lo, g_hi = cdiv(own_lo, self.step), min(full.No_ti_total, cdiv(own_hi, self.step))
            self.shards.append(dict(eng=e, ft=ft, own=(own_lo, own_hi), ex=(ex_lo, ex_hi),
                                    r0=ft.g0 - self.S_lo,                     # store row of the excerpt's first frame
                                    l0=g_lo - ft.g0, l1=g_hi - ft.g0,         # owned frames, excerpt-local
                                    n_stat=float(own_hi - own_lo)))
            self.h2d_bytes += e.h2d_bytes
        if self.world > 1:
            # one exchange of the (still zero) halo rows now: NCCL sets up its peer connections at the first send/recv
            # between two ranks, which otherwise lands inside the first adaptation
            exchange_halo_rows(self.new, (self.F_lo - self.S_lo, self.F_hi - self.S_lo), self.H, self.rank, self.world,
                               self.dist)
        self.tracks_hold = None          # (shard index, adaptation) whose tracks the pool currently holds
        self._copy_stream = self._copy_done = self._copy_prev = self._views = None
        self.exchanged_bytes = 0
        self.frames_owned = int(sum(len(sh["eng"].alists[0]) for sh in self.shards))

    # --------------------------------------------------------------------------------------------------
    def _load_knots(self, sh, store):
        """Excerpt rows of a frame-major store -> the engine's harmonic-major knot buffers [Kmax][G]."""
        e, r0 = sh["eng"], sh["r0"]
        G, K = e.G, self.Kmax
        for i in range(3):
            e.kn[i].view(K, G).copy_(store[i][r0:r0 + G].t())
        e.kn[3].copy_(store[3][r0:r0 + G])

    def _store_owned(self, sh, store):
        """Owned frames of the engine's knot buffers -> rows of a store."""
        e, r0, l0, l1 = sh["eng"], sh["r0"], sh["l0"], sh["l1"]
        G, K = e.G, self.Kmax
        for i in range(3):
            store[i][r0 + l0:r0 + l1].copy_(e.kn[i].view(K, G)[:, l0:l1].t())
        store[3][r0 + l0:r0 + l1].copy_(e.kn[3][l0:l1])

    def _best_am_ptr(self, sh):
        # Q10 (functions.py:210,383,400): the fallback's in-place amplitude lands in the rows of the best store
        return self.best[0].data_ptr() + sh["r0"] * self.Kmax * 8

    def _start_readback(self):
        """Device -> pinned host copy of the owned part of the best store and waveform on the copy stream, ordered
        after everything queued so far; the next adaptation runs meanwhile (it writes the OTHER store)."""
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        o0, o1 = self.F_lo - self.S_lo, self.F_hi - self.S_lo
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            self._views = [_to_host(self.s_best, "ls")] + [_to_host(x[o0:o1], "lk%d" % i) for i, x in enumerate(self.best)]
            self._copy_prev, self._copy_done = self._copy_done, torch.cuda.Event()
            self._copy_done.record(self._copy_stream)

    def run_device(self, on_adaptation=None, fixed_adaptations=None, prefetch_results=False):
        """Adaptation loop; results stay on the device.  Returns (SRER list, best adaptation, frame solves).
        fixed_adaptations (benchmark mode): exactly that many adaptations, no convergence stop (the best store still
        ends at the adaptation the reference would return), so that runs on different ranges do the same work.
        prefetch_results: copy the best store to the host after every improving adaptation, hidden behind the next
        adaptation (run() does; results() then only waits for the last copy)."""
        SRER, best_a, solves = [], 0, 0
        stopped = False
        self._copy_done = self._copy_prev = self._views = None
        ones = np.ones(1, bool)
        nsh = len(self.shards)
        stat_dev = torch.zeros((nsh, 2), dtype=torch.float64, device=self.device)
        # (n, mean, M2) of this rank's shards, padded to the largest shard count of any rank (n = 0 rows are ignored)
        rows_dev = torch.zeros((max(len(p) for p in self.plan_all), 3), dtype=torch.float64, device=self.device)
        rows_dev[:nsh, 0] = torch.tensor([sh["n_stat"] for sh in self.shards], dtype=torch.float64)
        self.pool.get("status", 2, torch.int32).zero_()
        self.pool.get("flops", 3, torch.float64).zero_()
        self.exchanged_bytes = 0
        self.tracks_hold = None
        for a in range(self.maxAdpt + 1 if fixed_adaptations is None else int(fixed_adaptations)):
            # ---- pass 1: analysis of adaptation a on the tracks of adaptation a-1 -----------------------------
            order = list(range(nsh))
            if self.tracks_hold is not None and self.tracks_hold[1] == a - 1:
                j = self.tracks_hold[0]                  # start with the shard whose tracks are still resident
                order = [j] + [i for i in order if i != j]
            for j in order:
                sh = self.shards[j]
                e = sh["eng"]
                if a > 0 and self.tracks_hold != (j, a - 1):
                    self._load_knots(sh, self.cur)
                    e.interpolate()
                e.best_am_ptr = self._best_am_ptr(sh)
                e.analysis(a)
                self.tracks_hold = None
                self._store_owned(sh, self.new)
                solves += e.tab.nA
            # ---- neighbour exchange of the new per-frame parameters --------------------------------------------
            self.exchanged_bytes += exchange_halo_rows(self.new, (self.F_lo - self.S_lo, self.F_hi - self.S_lo),
                                                       self.H, self.rank, self.world, self.dist)
            # ---- pass 2: tracks of adaptation a, resynthesis, statistics, candidate snapshot --------------------
            if self._copy_prev is not None:          # the candidate store was the best store two copies ago
                torch.cuda.current_stream(self.device).wait_event(self._copy_prev)
            for j in range(nsh):
                sh = self.shards[j]
                e = sh["eng"]
                self._load_knots(sh, self.new)
                e.interpolate()
                self.tracks_hold = (j, a)
                stat_dev[j].copy_(e.synthesize())
                e.snapshot(ones)
                own_lo, own_hi = sh["own"]
                ex_lo = sh["ex"][0]
                self.s_hat[own_lo - self.R_lo:own_hi - self.R_lo].copy_(e.s_hat[own_lo - ex_lo:own_hi - ex_lo])
                G, K, r0, l0, l1 = e.G, self.Kmax, sh["r0"], sh["l0"], sh["l1"]
                for i in range(3):
                    self.cand[i][r0 + l0:r0 + l1].copy_(e.best[i].view(G, K)[l0:l1])
                self.cand[3][r0 + l0:r0 + l1].copy_(e.best[3][l0:l1])
            rows_dev[:nsh, 1:].copy_(stat_dev)
            stat = allgather_rows(rows_dev, self.world, self.dist).cpu().numpy()   # the one host sync per adaptation
            self.d2h_bytes += stat.nbytes
            n, _, m2 = merge_stats([("long", r[0], r[1], r[2]) for r in stat if r[0] > 0])["long"]
            with np.errstate(all="ignore"):
                srer = 20 * np.log10(self.std_det / np.sqrt(m2 / n))           # functions.py:388
            SRER.append(srer)
            if on_adaptation is not None:
                on_adaptation(a, SRER)
            if a != 0 and srer <= SRER[a - 1]:                                 # functions.py:394-396
                if fixed_adaptations is None:
                    break
                stopped = True
            self.cur, self.new = self.new, self.cur
            if stopped:
                continue
            self.best, self.cand = self.cand, self.best                       # functions.py:397-402
            self.s_best, self.s_hat = self.s_hat, self.s_best
            best_a = a
            if prefetch_results:
                self._start_readback()
        fl = self.pool.get("flops", 3, torch.float64).cpu().numpy()
        self.useful_flops = dict(gram=float(fl[0]), chol=float(fl[1]), basis=float(fl[2]))
        check_not_pd(int(self.pool.get("status", 2, torch.int32)[0].item()))
        return SRER, best_a, solves

    def results(self, SRER, best_a, solves=0):
        """Owned part of the result as host arrays inside a full-length Result (zeros outside this rank's range)."""
        full, K = self.full, self.Kmax
        status = self.pool.get("status", 2, torch.int32).cpu().numpy()
        # a prefetched copy is final unless the 140 Hz fallback wrote into the best store afterwards (Q10)
        if self._copy_done is None or status[1] > 0:
            self._start_readback()
        self._copy_done.synchronize()
        views, self._copy_done, self._copy_prev, self._views = self._views, None, None, None
        n = self.F_hi - self.F_lo
        if self.world == 1:
            # the rank owns the whole signal: one copy out of the pinned staging buffers
            s_recon, am, ph, fm, a0 = _parallel_copies([views[0].numpy()] + [views[i].numpy().reshape(n, K) for i in (1, 2, 3)]
                                                       + [views[4].numpy()])
        else:
            s_recon = np.zeros(full.length)
            am, fm, ph = (np.zeros((full.No_ti_total, K)) for _ in range(3))
            a0 = np.zeros(full.No_ti_total)
            fr = slice(self.F_lo, self.F_hi)
            _parallel_fill([(s_recon[self.R_lo:self.R_hi], views[0].numpy()), (am[fr], views[1].numpy().reshape(n, K)),
                            (ph[fr], views[2].numpy().reshape(n, K)), (fm[fr], views[3].numpy().reshape(n, K)),
                            (a0[fr], views[4].numpy())])
        self.d2h_bytes += (self.R_hi - self.R_lo) * 8 + n * (3 * K + 1) * 8
        # flags of the owned frames inside full-length arrays (like the parameters: zeros outside this rank's range)
        T, b = full.No_ti_total, full.g_base
        speech, voiced = np.zeros(T, bool), np.zeros(T, bool)
        speech[self.F_lo:self.F_hi] = full.isSpeech[self.F_lo - b:self.F_hi - b]
        voiced[self.F_lo:self.F_hi] = full.analysed[self.F_lo - b:self.F_hi - b]
        return Result(s_recon=s_recon, SRER=list(SRER), ti=self.step * np.arange(T, dtype=np.int64), isSpeech=speech,
                      isVoiced=voiced, a0=a0, am=am, fm=fm, ph=ph, n_adapt_best=int(best_a),
                      fs=int(self.fs), step=self.step,
                      stats=dict(shards=len(self.shards), frames_solved=int(solves), not_pd_frames=int(status[0]),
                                 fallback_flags=int(status[1]), exchanged_bytes=int(self.exchanged_bytes),
                                 owned_samples=(self.R_lo, self.R_hi), owned_frames=(self.F_lo, self.F_hi)))

    def prepare_readback(self):
        """Page-lock the staging buffers of results() ahead of the loop (seconds for an hour of signal)."""
        n = self.F_hi - self.F_lo
        _pinned("ls", self.s_best.numel(), torch.float64)
        for i in range(3):
            _pinned("lk%d" % i, n * self.Kmax, torch.float64)
        _pinned("lk3", n, torch.float64)

    def run(self, on_adaptation=None, fixed_adaptations=None):
        self.prepare_readback()
        return self.results(*self.run_device(on_adaptation, fixed_adaptations, prefetch_results=True))


def analyze_long_signal(s, fs, f0s, voiced5, frame_step, f0min, Kmax, shard_seconds=20.0, det_part=None, step=15,
                        maxAdpt=10, pitchPeriods=3, analysisWindow=32, rank=0, world=1, dist=None, **engine_kw):
    """Frame-range sharded analysis of one long signal.  Returns (Result holding this rank's owned range inside
    full-length arrays, plan of all ranks, this rank's owned sample range)."""
    eng = LongSignalEngine(s, fs, f0s, voiced5, frame_step, f0min, Kmax, shard_seconds, det_part, step, maxAdpt,
                           pitchPeriods, analysisWindow, rank, world, dist, **engine_kw)
    return eng.run(), eng.plan_all, (eng.R_lo, eng.R_hi)
