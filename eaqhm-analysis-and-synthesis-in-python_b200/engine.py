"""Adaptation loop of the eaQHM hot path on one GPU (functions.py:163-402).

Python orchestrates; every numeric stage is a call through the C ABI of
libeaqhm_b200.so (include/eaqhm_b200.h) on device buffers owned by torch tensors.
A batch of utterances (or the frame-range shards of one long signal) is laid out on
one concatenated sample axis; utterances adapt in lockstep and drop out of the
batch when their SRER stops improving (functions.py:394-396).
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .host import FrameTables
from .sharding import stratified_chunk_order


@dataclass
class Utterance:
    """Array-level inputs of the hot path for one signal."""
    s: np.ndarray                   # (length,) float64 in [-1,1)
    fs: int
    f0s: np.ndarray                 # (n5,3) rows [t, f0, strength] on the 5 ms grid (functions.py:113)
    voiced5: np.ndarray             # (n5',) isVoiced per 5 ms frame after the mode handling (functions.py:127-146)
    frame_step: int                 # samples between 5 ms frames (functions.py:642)
    f0min: float                    # lower pitch limit (functions.py:95-109)
    Kmax: int                       # functions.py:117-120
    det_part: np.ndarray = None     # SRER target (functions.py:138,146); default s
    # excerpts of one long signal (longsignal.py) count only their owned samples `stat_range` in the SRER statistics
    # and use the whole signal's target std / stale f0 (Q1)
    stat_range: tuple = None
    group_std_det: float = None
    f0stale: float = None
    xoff: int = 0                   # position of s[0] inside the signal this utterance was cut from
    # streamed long signals (longsignal.py): the utterance is an excerpt (owned range + exchanged halo) that analyses
    # only the frames centred in own_range (local samples); s / det_part may then be device tensors (views of the
    # resident signal), and the per-frame tables come pre-sliced from the whole signal's (Engine(frame_tables=...))
    own_range: tuple = None


@dataclass
class Result:
    s_recon: np.ndarray
    SRER: list
    ti: np.ndarray                  # 0-based centre of every frame
    isSpeech: np.ndarray
    isVoiced: np.ndarray
    a0: np.ndarray                  # (No_ti,)
    am: np.ndarray                  # (No_ti,Kmax)
    fm: np.ndarray
    ph: np.ndarray
    n_adapt_best: int = 0
    stats: dict = field(default_factory=dict)
    fs: int = 0                     # sample rate and hop: what params.resynthesize / save_result need besides the arrays
    step: int = 15


class NotPositiveDefiniteError(np.linalg.LinAlgError):
    """A frame's normal matrix E_w^H E_w was not positive definite.  The reference's explicit inverse
    (functions.py:465,530) raises numpy.linalg.LinAlgError on an exactly singular matrix and returns noise on a
    numerically singular one; the Cholesky solve of this path detects both and refuses to return that noise."""


def check_not_pd(count):
    if count:
        raise NotPositiveDefiniteError(
            "%d frame solve(s) met a normal matrix that is not positive definite (window shorter than the number of "
            "unknowns, e.g. pitchPeriods <= 2, or an all-zero window); the reference's inv() raises LinAlgError or "
            "returns noise for such frames (functions.py:465,530)" % int(count))


_h2d_counter = [0]


def _dev(a, dtype, device):
    a = np.ascontiguousarray(a, dtype=dtype)
    _h2d_counter[0] += a.nbytes
    return torch.from_numpy(a).to(device, non_blocking=False)


def merge_stats(rows):
    """{group: (n, mean, M2)} from rows (group, n, mean, M2): Chan's pairwise combination of the per-shard
    population statistics, so that sqrt(M2/n) of a group equals numpy.std over all its samples (functions.py:388)."""
    acc = {}
    for g, n, mean, m2 in rows:
        if g in acc:
            n0, mean0, m20 = acc[g]
            nt = n0 + n
            dlt = mean - mean0
            acc[g] = (nt, mean0 + dlt * n / nt, m20 + m2 + dlt * dlt * n0 * n / nt)
        else:
            acc[g] = (n, mean, m2)
    return acc


_WORKSPACE = {}


def _workspace(nbytes, device):
    """Grow-only analysis scratch shared by the engines of a process (like a BLAS workspace): stages of different
    engines are stream-ordered on the current stream, and the scratch never carries state between stages."""
    key = str(device)
    t = _WORKSPACE.get(key)
    if t is None or t.numel() < nbytes:
        _WORKSPACE.pop(key, None)
        t = None
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACE[key] = t
    return t


class BufferPool:
    """Grow-only named device buffers shared by the shard engines of one streamed long signal: a shard engine
    owns its (small) tables, while tracks / knots / work areas hold data only during that shard's pass."""

    def __init__(self, device):
        self.device, self.bufs = torch.device(device), {}

    def get(self, name, numel, dtype):
        t = self.bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            self.bufs.pop(name, None)
            t = None
            t = torch.zeros(max(1, numel), dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t[:numel]


_PINNED = {}


_COPY_POOL = None


def _parallel_copies(views):
    """Pageable copies of pinned staging views on a few host threads (numpy releases the GIL for large copies): the
    result arrays own their memory, the staging buffers are reused by the next call."""
    global _COPY_POOL
    if sum(v.nbytes for v in views) < (4 << 20):
        return [v.copy() for v in views]
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=4)
    return list(_COPY_POOL.map(np.copy, views))


def _parallel_fill(pairs):
    """dst[...] = src for (dst view, src view) pairs on the host copy threads (results of a rank inside full-length arrays)."""
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=4)
    list(_COPY_POOL.map(lambda p: np.copyto(p[0], p[1]), pairs))


def _pinned(tag, n, dtype):
    """Cached pinned staging buffer (grow-only, one per tag; page-locking gigabytes takes seconds, so long signals
    call this before their loop starts)."""
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < n or buf.dtype != dtype:
        buf = torch.empty(n, dtype=dtype, pin_memory=True)
        _PINNED[tag] = buf
    return buf


def _to_host(t, tag):
    """Device tensor -> numpy through a cached pinned staging buffer (grow-only, one per tag)."""
    n = t.numel()
    buf = _pinned(tag, n, t.dtype)
    view = buf[:n]
    view.copy_(t.reshape(-1), non_blocking=True)
    return view


class Engine:
    """Holds the device buffers of one batch and runs the adaptation loop."""

    def __init__(self, utterances, step=15, maxAdpt=10, pitchPeriods=3, analysisWindow=32,
                 device="cuda", chunk=None, scratch_budget_bytes=None, frame_tables=None,
                 pool=None, phasor_tracks=True, staging_set=0):
        if not torch.cuda.is_available():
            raise _lib.EaqhmError("no CUDA device: the eaQHM B200 path has no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device(device)
        _h2d_counter[0] = 0
        self.utts = list(utterances)
        U = len(self.utts)
        if U == 0:
            raise ValueError("empty batch")
        fs = self.utts[0].fs
        if any(u.fs != fs for u in self.utts):
            raise ValueError("all utterances of a batch must share the sample rate")
        self.fs, self.step, self.maxAdpt = fs, int(step), int(maxAdpt)
        self.ft = list(frame_tables) if frame_tables is not None else [
            FrameTables(len(u.s), fs, u.f0s, u.voiced5, u.frame_step, u.Kmax, step, pitchPeriods, analysisWindow)
            for u in self.utts]
        self.pool = pool
        self.U = U
        self.Kmax = max(u.Kmax for u in self.utts)
        lens = np.array([f.length for f in self.ft], np.int64)
        cnts = np.array([f.No_ti for f in self.ft], np.int32)
        self.ustart = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
        self.gfirst = np.concatenate(([0], np.cumsum(cnts)[:-1])).astype(np.int32)
        self.T, self.G = int(lens.sum()), int(cnts.sum())
        self.lens, self.cnts = lens, cnts
        d = self.device
        cat = np.concatenate
        self.t_pos = _dev(cat([self.ustart[i] + f.ti0 for i, f in enumerate(self.ft)]), np.int64, d)
        self.t_own = _dev(cat([f.own for f in self.ft]), np.int32, d)
        self.t_utt = _dev(cat([np.full(f.No_ti, i) for i, f in enumerate(self.ft)]), np.int32, d)
        self.t_N = _dev(cat([f.N for f in self.ft]), np.int32, d)
        self.t_K0 = _dev(cat([f.K0 for f in self.ft]), np.int32, d)
        self.t_f0 = _dev(cat([f.f0 for f in self.ft]), np.float64, d)
        self.t_ustart = _dev(self.ustart, np.int64, d)
        self.t_ulen = _dev(lens, np.int64, d)
        self.t_gfirst = _dev(self.gfirst, np.int32, d)
        self.t_gcount = _dev(cnts, np.int32, d)
        self.t_f0stale = _dev([f.f0stale if u.f0stale is None else float(u.f0stale)
                               for u, f in zip(self.utts, self.ft)], np.float64, d)
        sr = [(0, f.length) if u.stat_range is None else tuple(u.stat_range) for u, f in zip(self.utts, self.ft)]
        self.stat_n = np.array([hi - lo for lo, hi in sr], np.float64)
        self.t_xoff = _dev([int(u.xoff) for u in self.utts], np.int64, d)
        self.t_stat_lo = _dev([lo for lo, _ in sr], np.int64, d)
        self.t_stat_hi = _dev([hi for _, hi in sr], np.int64, d)
        self.t_f0min = _dev([float(u.f0min) for u in self.utts], np.float64, d)
        self.t_active = torch.ones(U, dtype=torch.uint8, device=d)
        # frames each utterance analyses: all gated frames (functions.py:180-181), or for an excerpt of a streamed
        # long signal those centred in the range it owns; the 140 Hz fallback flags (Q6) are then also evaluated on
        # the neighbours' frames whose centre an owned window can reach
        def owned(u, f, margin):
            if u.own_range is None:
                return f.analysed
            return f.analysed & (f.ti0 >= u.own_range[0] - margin) & (f.ti0 < u.own_range[1] + margin)
        self.N_max = int(max(1, max(int(f.N.max()) for f in self.ft)))
        sel = [owned(u, f, 0) for u, f in zip(self.utts, self.ft)]
        self.alists = [self.gfirst[i] + np.flatnonzero(m).astype(np.int32) for i, m in enumerate(sel)]
        self.flists = None
        if any(u.own_range is not None for u in self.utts):
            self.flists = [self.gfirst[i] + np.flatnonzero(owned(u, f, self.N_max)).astype(np.int32)
                           for i, (u, f) in enumerate(zip(self.utts, self.ft))]
            self.t_flist = torch.empty(max(1, sum(len(a) for a in self.flists)), dtype=torch.int32, device=d)
        # cost proxy of a frame solve (~ K^3): launches take the expensive frames first (longest-processing-time order)
        self.acost = [(f.K0[m].astype(np.int64) ** 3) for f, m in zip(self.ft, sel)]
        self.sort_frames = os.environ.get("EAQHM_SORT_FRAMES", "1") != "0"
        self.t_alist = torch.empty(max(1, sum(len(a) for a in self.alists)), dtype=torch.int32, device=d)
        if all(torch.is_tensor(u.s) for u in self.utts) and U == 1:
            # excerpt of a signal that is already resident: views, no copies
            u = self.utts[0]
            self.s = u.s
            self.det = u.s if u.det_part is None else u.det_part
            assert self.s.is_cuda and self.s.dtype == torch.float64 and self.s.is_contiguous()
            if u.group_std_det is None:
                raise ValueError("a device-resident excerpt needs the whole signal's target std (group_std_det)")
            self.std_det = np.array([float(u.group_std_det)])
        else:
            self.s = _dev(cat([np.asarray(u.s, float).reshape(-1) for u in self.utts]), np.float64, d)
            dets = [np.asarray(u.s if u.det_part is None else u.det_part, float).reshape(-1) for u in self.utts]
            self.det = _dev(cat(dets), np.float64, d)
            self.std_det = np.array([np.std(x) if u.group_std_det is None else float(u.group_std_det)
                                     for u, x in zip(self.utts, dets)])          # functions.py:161
        self.nnz_max = int(max(1, max(int(f.K0.max()) for f in self.ft)))
        K, T, G = self.Kmax, self.T, self.G

        def buf(name, numel, dtype=torch.float64):
            if self.pool is not None:
                return self.pool.get(name, numel, dtype)
            return torch.zeros(numel, dtype=dtype, device=d)
        # am, ph, fmc [Kmax][T] and, unless switched off (then the analysis always forms the reference's running sum
        # itself: 40 % less track memory, slower basis), the phasor tracks cos ph, sin ph
        self.tr = [buf("tr%d" % i, K * T) for i in range(5 if phasor_tracks else 3)]
        self.kn = [buf("kn%d" % i, K * G) for i in range(3)] + [buf("kn3", G)]               # am, ph, fm, a0
        self.best = [buf("best%d" % i, K * G) for i in range(3)] + [buf("best3", G)]
        self.fm_fin = buf("fm_fin", K * G)
        self.a0mom = buf("a0mom", G)
        self.s_hat = buf("s_hat", T)
        self.s_best = buf("s_best", T)
        self.stat = buf("stat", 2 * U)
        self.improved = buf("improved", U, torch.uint8)
        self.status = buf("status", 2, torch.int32)
        self.flops = buf("flops", 3)
        nA = int(self.t_alist.numel())
        self.pipelined = self.L.eaqhm_analysis_pipeline_sets() >= 2
        if chunk is None:
            if scratch_budget_bytes is None:
                # big launches amortise the tail of the per-frame kernels (profiles/: 40 GB vs 6 GB = -10 % step time, 80 GB
                # - the 21,270 frames of a 20 s c2 range in one chunk - another -1 %); at most half of what is free: the
                # workspace is shared by the engines of a process (one process per GPU)
                ws = _WORKSPACE.get(str(self.device))
                if ws is not None and ws.numel() >= (8 << 30):
                    scratch_budget_bytes = int(ws.numel())      # reuse the process's workspace (no driver query: ~10 ms)
                else:
                    free, _ = torch.cuda.mem_get_info(self.device)
                    # experiment knob (tools only): EAQHM_SCRATCH_GB = budget in GiB, capped at half of what is free
                    env_gb = os.environ.get("EAQHM_SCRATCH_GB")
                    scratch_budget_bytes = int(min(int(float(env_gb) * (1 << 30)) if env_gb else 80 << 30, free // 2))
            per_frame = self.L.eaqhm_analysis_scratch_bytes(0, 1, self.nnz_max, self.N_max)
            chunk = int(max(1, min(nA, scratch_budget_bytes // max(1, per_frame))))
            if self.pipelined:
                # chunk pipeline (Cholesky of one chunk beside the basis + Gram of the next): the first Gram and the last
                # Cholesky overlap nothing, so aim at PIPE_CHUNKS chunks, but keep every launch a few waves of frames
                want = int(os.environ.get("EAQHM_PIPE_CHUNKS", "6"))
                floor = int(os.environ.get("EAQHM_PIPE_MIN_CHUNK", "2400"))
                chunk = int(max(1, min(chunk, max(floor, -(-nA // max(1, want))))))
        self.chunk = int(chunk)
        nb = self.L.eaqhm_analysis_scratch_bytes(G, self.chunk, self.nnz_max, self.N_max)
        self.scratch = _workspace(nb, d)
        self.scratch_bytes = int(nb)
        self.iwork = buf("iwork", self.L.eaqhm_interp_work_bytes(G, K), torch.uint8)
        self.swork = buf("swork", self.L.eaqhm_synth_work_bytes(int(lens.max()), U), torch.uint8)
        self.tab = _lib.Tables()
        t = self.tab
        t.T, t.maxlen, t.G, t.U, t.Kmax, t.step, t.fs = T, int(lens.max()), G, U, K, self.step, float(fs)
        for name in ("pos", "own", "utt", "N", "K0", "f0", "ustart", "ulen", "gfirst", "gcount", "f0stale",
                     "f0min", "active", "stat_lo", "stat_hi", "xoff", "alist"):
            setattr(t, name, getattr(self, "t_" + name).data_ptr())
        t.nA = 0
        t.flist, t.nF = None, 0
        self.best_am_ptr = None         # default: this engine's own snapshot (Q10); a streamed signal points it at its store
        self.c_tr = _lib.Tracks(*([x.data_ptr() for x in self.tr] + [None] * (5 - len(self.tr))))
        self.c_kn = _lib.Knots(*[x.data_ptr() for x in self.kn])
        self.c_best = _lib.Knots(*[x.data_ptr() for x in self.best])
        self.h2d_bytes = _h2d_counter[0]
        self.d2h_bytes = 0
        self.useful_flops = dict(gram=0.0, chol=0.0, basis=0.0)
        # read-back of the best snapshot overlapped with the next adaptation (run(): copy stream + pinned staging)
        self._copy_stream = None
        self._copy_done = None
        self._views = None
        # pinned staging buffers are cached per process and tag; two engines whose read-backs overlap in time (the corpus
        # runner hands the results of one batch to a worker thread while the next batch computes) use different sets
        self._stag = "" if not staging_set else "#%d" % int(staging_set)

    # ------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_active(self, active):
        act = np.asarray(active, bool)
        self.t_active.copy_(torch.from_numpy(act.astype(np.uint8)))
        self.h2d_bytes += self.U
        al = [self.alists[i] for i in range(self.U) if act[i]]
        al = np.concatenate(al) if al else np.zeros(0, np.int32)
        if self.sort_frames and len(al):
            cost = np.concatenate([self.acost[i] for i in range(self.U) if act[i]])
            al = al[np.argsort(-cost, kind="stable")]
            n = -(-len(al) // self.chunk)
            if n >= 2 and self.pipelined:
                # chunk pipeline: every chunk a stratified sample of the cost-ordered list (chunk b takes ranks b, b+n,
                # b+2n, ...: expensive frames first inside a chunk, equal work per chunk), so that the Cholesky of one
                # chunk and the Gram of the next stay matched and the two launches that overlap nothing are 1/n of the work
                al = al[stratified_chunk_order(len(al), self.chunk)]
        if len(al):
            self.t_alist[:len(al)].copy_(torch.from_numpy(al))
            self.h2d_bytes += al.nbytes
        self.tab.nA = int(len(al))
        if self.flists is not None:
            fl = [self.flists[i] for i in range(self.U) if act[i]]
            fl = np.concatenate(fl) if fl else np.zeros(0, np.int32)
            if len(fl):
                self.t_flist[:len(fl)].copy_(torch.from_numpy(fl))
                self.h2d_bytes += fl.nbytes
            self.tab.flist, self.tab.nF = self.t_flist.data_ptr(), int(len(fl))
        return int(len(al))

    def analysis(self, a):
        """Stage A (functions.py:176-329)."""
        for x in self.kn:
            x.zero_()
        with torch.cuda.device(self.device):          # the library launches on the current device
            self._analysis(a)

    def _analysis(self, a):
        _lib.check(self.L.eaqhm_analysis(C.byref(self.tab), self.s.data_ptr(), C.byref(self.c_tr), C.byref(self.c_kn),
                                         self.best_am_ptr or self.best[0].data_ptr(), a, self.nnz_max, self.N_max,
                                         self.chunk,
                                         self.scratch.data_ptr(), self.scratch_bytes, self.status.data_ptr(),
                                         self.flops.data_ptr(), self._stream()))

    def interpolate(self):
        """Stage B (functions.py:337-383)."""
        with torch.cuda.device(self.device):
            self._interpolate()

    def _interpolate(self):
        _lib.check(self.L.eaqhm_interpolate(C.byref(self.tab), C.byref(self.c_kn), C.byref(self.c_tr),
                                            self.fm_fin.data_ptr(), self.a0mom.data_ptr(), self.iwork.data_ptr(),
                                            self.iwork.numel(), self._stream()))

    def synthesize(self):
        """Stage C (functions.py:385-388); returns [mean | M2] of the residual per utterance (device tensor)."""
        with torch.cuda.device(self.device):
            self._synthesize()
        return self.stat

    def _synthesize(self):
        _lib.check(self.L.eaqhm_synthesize(C.byref(self.tab), C.byref(self.c_tr), self.kn[3].data_ptr(),
                                           self.a0mom.data_ptr(), self.det.data_ptr(), self.s_hat.data_ptr(),
                                           self.stat.data_ptr(), self.swork.data_ptr(), self.swork.numel(),
                                           self._stream()))
        return self.stat

    def srer(self, stat):
        """SRER per utterance (functions.py:388) from the residual statistics [mean | M2] of eaqhm_synthesize."""
        with np.errstate(all="ignore"):
            return 20 * np.log10(self.std_det / np.sqrt(stat[self.U:] / self.stat_n))

    def _start_readback(self):
        """Device -> pinned host copy of the current best snapshot on the copy stream, ordered after everything queued on
        the main stream so far; the loop goes on meanwhile.  The next snapshot waits for it (it rewrites the source)."""
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            self._views = [_to_host(self.s_best, "s" + self._stag)] + \
                [_to_host(x, "k%d%s" % (i, self._stag)) for i, x in enumerate(self.best)]
            self._copy_done = torch.cuda.Event()
            self._copy_done.record(self._copy_stream)

    def snapshot(self, improved):
        if self._copy_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._copy_done)
        self.improved.copy_(torch.from_numpy(np.asarray(improved, bool).astype(np.uint8)))
        with torch.cuda.device(self.device):
            self._snapshot()

    def _snapshot(self):
        _lib.check(self.L.eaqhm_snapshot(C.byref(self.tab), self.improved.data_ptr(), C.byref(self.c_tr),
                                         C.byref(self.c_kn), self.fm_fin.data_ptr(), self.a0mom.data_ptr(),
                                         self.s_hat.data_ptr(), self.s_best.data_ptr(), C.byref(self.c_best),
                                         self._stream()))

    # ------------------------------------------------------------------------------
    def run_device(self, on_adaptation=None, fixed_adaptations=None, prefetch_results=False):
        """The adaptation loop (functions.py:163-402) on the resident inputs; results stay on the device.
        Returns (SRER per utterance, index of the best adaptation per utterance, frame solves).
        fixed_adaptations (benchmark mode): exactly that many adaptations for every utterance, no convergence stop;
        the snapshot still ends at the adaptation the reference would return.
        prefetch_results: start the device -> host copy of the snapshot after every improving adaptation, hidden behind
        the next adaptation's kernels (run() does; results() then only waits for the last copy)."""
        U = self.U
        self._copy_done = self._views = None
        active = np.ones(U, bool)
        stopped = np.zeros(U, bool)
        SRER = [[] for _ in range(U)]
        best_a = np.zeros(U, int)
        frames_solved = 0
        self.status.zero_()
        self.flops.zero_()
        for a in range(self.maxAdpt + 1 if fixed_adaptations is None else int(fixed_adaptations)):
            nA = self.set_active(active)
            frames_solved += nA
            self.analysis(a)
            self.interpolate()
            stat = self.synthesize().cpu().numpy()                          # the one host sync per adaptation
            self.d2h_bytes += stat.nbytes
            srer_u = self.srer(stat)
            improved = np.zeros(U, bool)
            with np.errstate(all="ignore"):
                for u in np.flatnonzero(active):
                    srer = srer_u[u]
                    SRER[u].append(srer)
                    if a != 0 and srer <= SRER[u][a - 1]:                   # functions.py:394-396
                        stopped[u] = True
                        active[u] = fixed_adaptations is not None
                    if stopped[u]:
                        pass
                    else:
                        improved[u] = True
                        best_a[u] = a
            self.snapshot(improved)                                         # functions.py:397-402
            if prefetch_results and improved.any():
                self._start_readback()
            self.h2d_bytes += U
            if on_adaptation is not None:
                on_adaptation(a, SRER)
            if not active.any():
                break
        fl = self.flops.cpu().numpy()
        self.useful_flops = dict(gram=float(fl[0]), chol=float(fl[1]), basis=float(fl[2]))
        check_not_pd(int(self.status[0].item()))
        return SRER, best_a, frames_solved

    def prepare_readback(self):
        """Page-lock the staging buffers of results() ahead of the loop."""
        _pinned("s" + self._stag, self.s_best.numel(), torch.float64)
        for i, x in enumerate(self.best):
            _pinned("k%d%s" % (i, self._stag), x.numel(), torch.float64)

    def run(self, on_adaptation=None):
        """Adaptation loop + copy of the results to host arrays.  Returns one Result per utterance."""
        self.prepare_readback()
        return self.results(*self.run_device(on_adaptation, prefetch_results=True))

    def results(self, SRER, best_a, frames_solved=0):
        K, G = self.Kmax, self.G
        status = self.status.cpu().numpy()
        # a prefetched copy is the final snapshot unless the 140 Hz fallback wrote its in-place amplitude into it
        # afterwards (Q10: the write of adaptation a+1 lands in the snapshot of adaptation a)
        if self._copy_done is None or status[1] > 0:
            self._start_readback()
        self._copy_done.synchronize()
        views, self._copy_done, self._views = self._views, None, None
        s_best = views[0].numpy()
        b_am, b_ph, b_fm = [v.numpy().reshape(G, K) for v in views[1:4]]            # frame-major snapshot
        b_a0 = views[4].numpy()
        self.d2h_bytes += s_best.nbytes + 3 * b_am.nbytes + b_a0.nbytes
        out = []
        for i, (u, f) in enumerate(zip(self.utts, self.ft)):
            g0, n = self.gfirst[i], f.No_ti
            sl = slice(g0, g0 + n)
            ku = u.Kmax
            s_u, a0_u, am_u, fm_u, ph_u = _parallel_copies([s_best[self.ustart[i]:self.ustart[i] + f.length], b_a0[sl],
                                                            b_am[sl, :ku], b_fm[sl, :ku], b_ph[sl, :ku]])
            out.append(Result(
                s_recon=s_u, SRER=list(SRER[i]), ti=f.ti0.copy(),
                isSpeech=f.isSpeech.copy(), isVoiced=f.analysed.copy(), a0=a0_u, am=am_u, fm=fm_u, ph=ph_u,
                n_adapt_best=int(best_a[i]), fs=int(self.fs), step=self.step,
                stats=dict(not_pd_frames=int(status[0]), fallback_frames=int(status[1]),
                           frames_solved=int(frames_solved))))
        return out
