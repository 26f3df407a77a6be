/*
 * eaqhm_b200.h — C ABI of the B200-native eaQHM analysis+synthesis hot path.
 *
 * The reference (Antibas/eaQHM-analysis-and-synthesis-in-Python) is pure Python and
 * has no FFI layer; its hot path is the adaptation loop of eaQHMAnalysisAndSynthesis
 * (functions.py:163-402) and the three helpers iqhmLS_complexamps (functions.py:420),
 * eaqhmLS_complexamps (functions.py:472) and phase_integr_interpolation
 * (functions.py:537).  This header declares what a ctypes binding of that path
 * binds: (1) the three helper seams with HOST buffers (unit parity points), and
 * (2) the stage entry points of one adaptation on DEVICE buffers, which the Python
 * orchestrator (eaqhm_b200/engine.py) calls in the order of functions.py:163-402.
 *
 * Conventions: every function returns 0 on success, a negative EAQHM_E_* code on
 * failure (eaqhm_last_error() gives the message).  All floating-point data is
 * IEEE double; complex values are interleaved (re, im).  Device entry points take
 * the CUDA stream as void* (cudaStream_t) and never synchronise.  No torch types.
 *
 * Data layout on the device:
 *   samples   : utterances of a batch (or the frame-range shard of one long signal)
 *               are concatenated on one sample axis of length T;
 *   frames    : global frame index g over the concatenation (G frames); frame g of
 *               utterance u has its centre at sample pos[g] (functions.py:148:
 *               ti = arange(1,length,step), 0-based centre ti-1);
 *   tracks    : dense per-sample tracks am/ph/fmc are harmonic-major [Kmax][T];
 *   knots     : per-frame parameters are harmonic-major [Kmax][G].
 */
#ifndef EAQHM_B200_H
#define EAQHM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAQHM_OK            0
#define EAQHM_E_ARG        -1   /* invalid argument                                   */
#define EAQHM_E_CUDA       -2   /* CUDA runtime error (message in eaqhm_last_error)   */
#define EAQHM_E_NOTPD      -3   /* a frame's normal matrix was not positive definite  */
#define EAQHM_E_NODEVICE   -4   /* no CUDA device: there is no CPU fallback           */

/* Library identification / error text (thread-local). */
int         eaqhm_version(void);
const char* eaqhm_last_error(void);
int         eaqhm_device_count(void);

/* Launch accounting and per-kernel-class device timing (CUDA events on the launching
 * stream).  Counts are always kept; timing only while enabled.  eaqhm_profile_read
 * synchronises the recorded events and returns accumulated milliseconds and launch
 * counts for the eaqhm_profile_classes() kernel classes (names by index). */
int         eaqhm_profile_enable(int on);
int         eaqhm_profile_reset(void);
int         eaqhm_profile_classes(void);
const char* eaqhm_profile_name(int kernel_class);
int         eaqhm_profile_read(double* ms, int64_t* counts, int32_t n);

/* ---------------------------------------------------------------------------
 * Tables describing the frames of a batch (device pointers unless noted).
 * Mirrors the per-frame quantities of functions.py:148-157,176-194.
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t        T;        /* samples on the concatenated axis                               */
    int64_t        maxlen;   /* longest utterance (samples)                                    */
    int32_t        G;        /* frames on the concatenated axis                                */
    int32_t        U;        /* utterances                                                     */
    int32_t        Kmax;     /* harmonic capacity (functions.py:117-120)                       */
    int32_t        step;     /* hop in samples (functions.py:35, default 15)                   */
    double         fs;       /* sample rate (Hz)                                               */
    const int64_t* pos;      /* [G] centre sample of frame g on the concatenated axis          */
    const int32_t* own;      /* [G] samples owned by frame g: [pos, pos+own)                   */
    const int32_t* utt;      /* [G] utterance of frame g                                       */
    const int32_t* N;        /* [G] half window (functions.py:191); 0 = frame not analysed      */
    const int32_t* K0;       /* [G] harmonics at adaptation 0 (functions.py:187)               */
    const double*  f0;       /* [G] f0 at adaptation 0 (functions.py:185)                      */
    const int64_t* ustart;   /* [U] first sample of utterance u                                */
    const int64_t* ulen;     /* [U] samples of utterance u                                     */
    const int32_t* gfirst;   /* [U] first frame of utterance u                                 */
    const int32_t* gcount;   /* [U] frames of utterance u                                      */
    const double*  f0stale;  /* [U] f0 of the last analysed frame (quirk Q1, functions.py:310) */
    const double*  f0min;    /* [U] lower pitch limit (functions.py:95-109,321)                */
    const uint8_t* active;   /* [U] 1 = utterance still adapting (converged ones are skipped)  */
    const int64_t* stat_lo;  /* [U] first local sample counted in the SRER statistics          */
    const int64_t* stat_hi;  /* [U] one past the last one (whole utterance: 0, ulen; a frame-   */
                             /*     range shard counts only the samples it owns, not its halo)  */
    const int64_t* xoff;     /* [U] position of the utterance's first sample inside the signal  */
                             /*     it was cut from (0 for whole signals): frame-range shards   */
                             /*     keep the reference's absolute sample coordinates (Q4 knots  */
                             /*     at samples 0 and step, functions.py:369)                    */
    const int32_t* alist;    /* [nA] analysed frames (global index g), any order               */
    int32_t        nA;
    const int32_t* flist;    /* [nF] frames whose 140 Hz fallback flag is evaluated (Q6); NULL  */
    int32_t        nF;       /*      = alist.  A shard lists its halo frames here as well       */
} eaqhm_tables;

/* Dense per-sample tracks [Kmax][T] (functions.py:159-160,169-171). */
typedef struct {
    double* am;    /* am_recon after interpolation == am_current of the next adaptation */
    double* ph;    /* ph_recon                                                          */
    double* fmc;   /* fm_current (Hz)                                                   */
    double* zc;    /* cos(ph_recon), sin(ph_recon): written by the interpolation stage; */
    double* zs;    /* inside a run the next adaptation's basis exp(j 2 pi cumsum(fm_current)/fs) */
                   /* (functions.py:509-516) is the ratio of two of these phasors.  Both may be  */
                   /* NULL: the analysis then always forms the running sum itself        */
} eaqhm_tracks;

/* Per-frame parameters [Kmax][G] (+ a0 [G]) written by the analysis stage. */
typedef struct {
    double* am;    /* |a_k|         (functions.py:316) 0 = not accepted */
    double* ph;    /* angle(a_k)    (functions.py:317)                  */
    double* fm;    /* frequency Hz  (functions.py:320-324)              */
    double* a0;    /* Re(DC)        (functions.py:303)                  */
} eaqhm_knots;

/* Scratch size for the analysis of a batch with G frames, processed `chunk` frames at
 * a time, with at most nnz_max active harmonics and half window N_max.  The caller
 * allocates (torch / cudaMalloc). */
int64_t eaqhm_analysis_scratch_bytes(int32_t G, int32_t chunk, int32_t nnz_max, int32_t N_max);

/* Copies of the per-chunk scratch inside that size (1 = chunks run one after the other; 2 = while one chunk is
 * factored the next one's basis and Gram are built on a second internal stream: a caller that wants the overlap
 * picks `chunk` so that the analysed frames make several chunks).  Environment EAQHM_PIPELINE overrides. */
int eaqhm_analysis_pipeline_sets(void);

/* ---------------------------------------------------------------------------
 * Stage A — analysis of all frames in tables->alist for adaptation `adapt`
 * (functions.py:176-329): active set + 140 Hz fallback (Q6), track gather, gap
 * fill, time-reversed mirror (Q2), adaptive (or harmonic, adapt==0) basis, weighted
 * complex Gram (weight w^2, Q3), Cholesky solve, frequency mismatch, acceptance,
 * knot write.  `knots` must be zero-filled by the caller beforehand.
 * `best_am` (may be NULL): [G][Kmax] (frame-major) snapshot that receives the in-place fallback
 * amplitude 10e-4 (the am_fin alias of functions.py:210,383,400).
 * `status` (device, int32[2]): [0] += frames whose matrix was not positive definite,
 * [1] += frames that used the fallback.
 * `flop_stats` (device, double[3], may be NULL): += algorithmic flops of the frames
 * analysed (Gram, Cholesky+solves, basis rhs) for the roofline report.
 * ------------------------------------------------------------------------- */
int eaqhm_analysis(const eaqhm_tables* tab, const double* s, const eaqhm_tracks* cur,
                   const eaqhm_knots* knots, double* best_am, int32_t adapt,
                   int32_t nnz_max, int32_t N_max, int32_t chunk,
                   void* scratch, int64_t scratch_bytes, int32_t* status, double* flop_stats,
                   void* stream);

/* ---------------------------------------------------------------------------
 * Stage B — interpolation between frame centres (functions.py:337-383): run
 * detection, linear amplitude, not-a-knot cubic frequency (short-run quirk Q4),
 * phase integration (functions.py:537-575, Q5), new fm_current, DC spline moments.
 * Writes every sample of tracks (zeros outside runs), fm_fin [Kmax][G] (fm_recon at
 * the centres) and a0mom [G] (second derivatives of the DC spline).
 * `work` : device scratch of eaqhm_interp_work_bytes(G, Kmax) bytes.
 * ------------------------------------------------------------------------- */
int64_t eaqhm_interp_work_bytes(int32_t G, int32_t Kmax);
int eaqhm_interpolate(const eaqhm_tables* tab, const eaqhm_knots* knots, const eaqhm_tracks* out,
                      double* fm_fin, double* a0mom, void* work, int64_t work_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Stage C — synthesis and SRER statistics (functions.py:385-388):
 * s_hat = a0 + 2*sum_k am*cos(ph); per utterance, over its statistics range, the mean
 * of e = det - s_hat and M2 = sum (e - mean)^2 (two-pass).  stat_out [2U]: mean in
 * [0,U), M2 in [U,2U).  The caller forms 20*log10(std_det/sqrt(M2/n)), combining the
 * shards of one signal with the pairwise-variance formula.
 * `work`: eaqhm_synth_work_bytes(maxlen, U) bytes.
 * ------------------------------------------------------------------------- */
int64_t eaqhm_synth_work_bytes(int64_t maxlen, int32_t U);
int eaqhm_synthesize(const eaqhm_tables* tab, const eaqhm_tracks* tr, const double* a0k,
                     const double* a0mom, const double* det, double* s_hat, double* stat_out,
                     void* work, int64_t work_bytes, void* stream);

/* Snapshot of the best adaptation (functions.py:397-402) for utterances with
 * improved[u] != 0: s_best <- s_hat, and the centre values of a0/am/fm/ph.  `best` is
 * frame-major, [G][Kmax] (+ a0 [G]): the layout the host reads back per utterance. */
int eaqhm_snapshot(const eaqhm_tables* tab, const uint8_t* improved, const eaqhm_tracks* tr,
                   const eaqhm_knots* knots, const double* fm_fin, const double* a0mom,
                   const double* s_hat, double* s_best, const eaqhm_knots* best, void* stream);

/* ---------------------------------------------------------------------------
 * Helper seams with HOST buffers (unit parity points; each call copies in, runs device
 * kernels on one frame, copies out, synchronises).  The two least-squares seams take
 * explicit column tracks / frequencies like the reference's helpers, so they run the
 * general route (explicit-column basis kernel, complex DMMA Gram k_gram_warp, complex
 * Cholesky k_chol_solve): the SAME Gram and Cholesky kernels as stage A at adaptations
 * >= 1.  Stage A's adaptation 0 uses a cheaper real even/odd formulation of the same
 * normal equations (k_a0_build_real, k_chol_real) and its phase integration is fused into
 * the interpolation kernel (k_interp_hw); those product kernels are covered end to end by
 * the golden-fixture tests (tests/test_gpu_parity.py), the seams by their own.
 * ------------------------------------------------------------------------- */

/* functions.py:420  iqhmLS_complexamps(s, f0range, window, fs) -> amplitudes, slopes.
 * s[L], f0range[M] (Hz, ascending -K..K times f0), window[L]; amp/slope: M complex. */
int eaqhm_iqhmLS_complexamps(const double* s, int32_t L, const double* f0range, int32_t M,
                             const double* window, double fs, double* amp, double* slope);

/* functions.py:472  eaqhmLS_complexamps(s, am, fm, window, fs) -> amplitudes, slopes.
 * am, fm: row-major [L][M] per-sample tracks exactly as the reference passes them. */
int eaqhm_eaqhmLS_complexamps(const double* s, int32_t L, const double* am, const double* fm,
                              int32_t M, const double* window, double fs, double* amp, double* slope);

/* functions.py:537  phase_integr_interpolation(fm_recon, ph_recon, indices) -> phases.
 * fm_recon (rad/sample) and ph_recon are full columns of `length` samples; indices[n]
 * ascending knot samples; out has indices[n-1]-indices[0]+1 samples. */
int eaqhm_phase_integr_interpolation(const double* fm_recon, const double* ph_recon, int64_t length,
                                     const int64_t* indices, int32_t n, double* out);

/* misc.py:167-182 (ellipFilter -> scipy.signal.filtfilt(b, a, x), default odd padding of `padlen` samples and lfilter_zi
 * initial conditions `zi`[ntaps-1]) for `nutt` signals packed in x: signal u is x[start[u] .. start[u]+len[u]), len[u] >
 * padlen.  One thread per signal, scipy's operation order, no FMA: bit-identical output in y (same packing).  work: scratch,
 * signal u uses work[work_start[u] .. +len[u]+2*padlen).  ntaps = len(b) = len(a) = 7 (6th order), a[0] = 1.  DEVICE buffers. */
int eaqhm_filtfilt_batch(const double* x, const int64_t* start, const int64_t* len, int32_t nutt, const double* b,
                         const double* a, const double* zi, int32_t ntaps, int32_t padlen, double* y, double* work,
                         const int64_t* work_start, void* stream);

/* functions.py:616-634 (voicedUnvoicedFrames): energy_db[i] = 20 log10 numpy.std(s[ti + (-mid-1 .. mid-1)]) for the 5 ms
 * frame centres ti = 1 + i*step, window_len = 2*mid + 1 samples (30 ms, odd), likewise for the low-passed signal; NaN
 * for frames whose window leaves the signal.  DEVICE buffers; thresholds and the median filter stay with the caller. */
int eaqhm_vuv_energy(const double* s, const double* s_smooth, int64_t length, int32_t window_len, int32_t step,
                     int64_t nframes, double* energy_db, double* energy_smooth_db, void* stream);

/* ---------------------------------------------------------------------------
 * SWIPE' pitch-strength matrix (host preprocessing of the path: functions.py:111 -> SWIPE.py:14-132), DEVICE buffers.
 * The caller (eaqhm_b200/preprocess.py:swipep_gpu) owns the grids that depend on the sample rate and pitch range only
 * and does the parabolic refinement of the winning candidate (SWIPE.py:106-131) on the host.
 * ------------------------------------------------------------------------- */

/* SWIPE.py:56-72,140-142 for one power-of-two window size ws (<= 8192): frame `seg` covers the zero-padded samples
 * [seg*hop - pad_front, +ws); Hann `window`[ws] with sumw2 = sum(window^2); PSD on the rfft frequencies f[ws/2+1];
 * linear interpolation onto ferb[nerb] (lo[e] = index of the rfft bin at or below ferb[e], <= ws/2 - 1), square root,
 * unit norm per frame.  Ln: [nerb][nseg]. */
int eaqhm_swipe_loudness(const double* x, int64_t n_x, int32_t ws, int32_t hop, int32_t pad_front,
                         const double* window, double fs, double sumw2, const double* f, const int32_t* lo,
                         const double* ferb, int32_t nerb, int64_t nseg, double* Ln, void* stream);

/* SWIPE.py:146: Si[ncand][nseg] = K[ncand][nerb] x Ln[nerb][nseg] (K = candidate kernels, SWIPE.py:166-184). */
int eaqhm_swipe_strength(const double* K, int32_t ncand, int32_t nerb, const double* Ln, int64_t nseg, double* Si,
                         void* stream);

/* SWIPE.py:92-104: S[rows[r]][t] += mu[r] * (linear interpolation of Si[r] at output time t): hi[t] = upper frame index
 * (< 0: t outside the frame times -> NaN), w1[t] = t - tim[hi-1], w2[t] = tim[hi] - tim[hi-1].  S: [npc][nt]. */
int eaqhm_swipe_accumulate(const double* Si, int32_t ncand, int64_t nseg, const int32_t* rows, const double* mu,
                           const int32_t* hi, const double* w1, const double* w2, int64_t nt, double* S, void* stream);

/* SWIPE.py:108: best[t] = numpy.argmax(S[:, t]) (first NaN, else first maximum); y[3][nt] = S[best-1], S[best],
 * S[best+1] at t (NaN beyond the candidate range). */
int eaqhm_swipe_argmax(const double* S, int32_t npc, int64_t nt, int32_t* best, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EAQHM_B200_H */
