// Stage C of one eaQHM adaptation: sinusoidal resynthesis and the SRER statistics
// (reference: functions.py:385-388), plus the best-adaptation snapshot (functions.py:397-402).
#include "common.cuh"

namespace {

constexpr int kSynThreads = 256;

// DC spline value at local sample n of utterance u (functions.py:340, with extrapolation past the last knot)
__device__ __forceinline__ double a0_at(const eaqhm_tables& t, const double* a0k, const double* a0mom, int u, int64_t n) {
    const int g0 = t.gfirst[u], nu = t.gcount[u];
    int i = (int)(n / t.step);
    if (i > nu - 2) i = nu - 2;
    const double h = (double)t.step;
    return spline_eval(a0k[g0 + i], a0k[g0 + i + 1], a0mom[g0 + i], a0mom[g0 + i + 1], h,
                       (double)(n - (int64_t)i * t.step));
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;   // valid in warp 0
}

// s_hat = a0 + 2 * sum_k am*cos(ph); block partial sums of e = det - s_hat
__global__ void __launch_bounds__(kSynThreads) k_synth(eaqhm_tables t, eaqhm_tracks tr, const double* a0k,
                                                       const double* a0mom, const double* det, double* s_hat,
                                                       double* part, int nblk) {
    __shared__ double sh[kSynThreads / 32];
    const int u = blockIdx.y;
    if (!t.active[u]) return;
    const int64_t len = t.ulen[u], us = t.ustart[u];
    double e = 0.0;
    const int64_t n = (int64_t)blockIdx.x * kSynThreads + threadIdx.x;
    if (n < len) {
        double acc = 0.0;
        const int64_t x = us + n;
        // four harmonics per trip: the eight loads are independent, so the HBM latency overlaps (the amplitude is zero
        // outside the runs, where the phase is zero too, so the unconditional cos costs nothing in accuracy)
        int k = 0;
        for (; k + 4 <= t.Kmax; k += 4) {
            double a[4], p[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                a[q] = tr.am[(int64_t)(k + q) * t.T + x];
                p[q] = tr.ph[(int64_t)(k + q) * t.T + x];
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (a[q] != 0.0) acc += a[q] * cos(p[q]);
        }
        for (; k < t.Kmax; k++) {
            const double a = tr.am[(int64_t)k * t.T + x];
            if (a != 0.0) acc += a * cos(tr.ph[(int64_t)k * t.T + x]);
        }
        const double v = a0_at(t, a0k, a0mom, u, n) + 2.0 * acc;
        s_hat[x] = v;
        if (n >= t.stat_lo[u] && n < t.stat_hi[u]) e = det[x] - v;
    }
    const double bs = block_sum(e, sh);
    if (threadIdx.x == 0) part[(int64_t)u * nblk + blockIdx.x] = bs;
}

// mean of e per utterance from the block partials (fixed order -> deterministic)
__global__ void k_reduce_parts(eaqhm_tables t, const double* part, int nblk, double* out, bool divide) {
    __shared__ double sh[kSynThreads / 32];
    const int u = blockIdx.x;
    if (!t.active[u]) return;
    const int nb = (int)((t.ulen[u] + kSynThreads - 1) / kSynThreads);
    double v = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) v += part[(int64_t)u * nblk + b];
    const double r = block_sum(v, sh);
    if (threadIdx.x == 0) out[u] = divide ? r / (double)(t.stat_hi[u] - t.stat_lo[u]) : r;
}

// second pass of the population variance (numpy.std is two-pass, functions.py:388)
__global__ void __launch_bounds__(kSynThreads) k_var_pass(eaqhm_tables t, const double* det, const double* s_hat,
                                                          const double* mean, double* part, int nblk) {
    __shared__ double sh[kSynThreads / 32];
    const int u = blockIdx.y;
    if (!t.active[u]) return;
    const int64_t len = t.ulen[u], us = t.ustart[u];
    const int64_t n = (int64_t)blockIdx.x * kSynThreads + threadIdx.x;
    double q = 0.0;
    if (n < len && n >= t.stat_lo[u] && n < t.stat_hi[u]) {
        const double d = (det[us + n] - s_hat[us + n]) - mean[u];
        q = d * d;
    }
    const double bs = block_sum(q, sh);
    if (threadIdx.x == 0) part[(int64_t)u * nblk + blockIdx.x] = bs;
}

__global__ void k_snapshot_knots(eaqhm_tables t, const uint8_t* improved, eaqhm_tracks tr, const double* fm_fin,
                                 const double* a0k, const double* a0mom, eaqhm_knots best) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (g >= t.G) return;
    const int u = t.utt[g];
    if (!improved[u]) return;
    // the snapshot is frame-major [G][Kmax]: it is what the host reads back and slices per utterance
    const int64_t x = t.pos[g], o = (int64_t)k * t.G + g, ob = (int64_t)g * t.Kmax + k;
    best.am[ob] = tr.am[(int64_t)k * t.T + x];
    best.ph[ob] = tr.ph[(int64_t)k * t.T + x];
    best.fm[ob] = fm_fin[o];
    if (k == 0) best.a0[g] = a0_at(t, a0k, a0mom, u, x - t.ustart[u]);
}

__global__ void k_snapshot_signal(eaqhm_tables t, const uint8_t* improved, const double* s_hat, double* s_best) {
    const int u = blockIdx.y;
    if (!improved[u]) return;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < t.ulen[u]) s_best[t.ustart[u] + n] = s_hat[t.ustart[u] + n];
}

}  // namespace

static int synth_nblk(int64_t maxlen) { return (int)((maxlen + kSynThreads - 1) / kSynThreads); }

extern "C" int64_t eaqhm_synth_work_bytes(int64_t maxlen, int32_t U) {
    if (maxlen <= 0 || U <= 0) return EAQHM_E_ARG;
    // block partials [U][nblk(maxlen)] + mean [U] + spare
    const int64_t nblk = synth_nblk(maxlen);
    return align_up(((int64_t)U * nblk + 2 * (int64_t)U) * 8, 256);
}

extern "C" int eaqhm_synthesize(const eaqhm_tables* tab, const eaqhm_tracks* tr, const double* a0k,
                                const double* a0mom, const double* det, double* s_hat, double* stat_out,
                                void* work, int64_t work_bytes, void* stream) {
    CHECK_ARG(tab && tr && a0k && a0mom && det && s_hat && stat_out && work);
    CHECK_ARG(tab->maxlen > 0 && work_bytes >= eaqhm_synth_work_bytes(tab->maxlen, tab->U));
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = synth_nblk(tab->maxlen);   // blocks past a shorter utterance's end contribute zero
    double* part = (double*)work;
    double* mean = stat_out;
    double* m2 = stat_out + tab->U;
    dim3 grid(nblk, tab->U);
    LAUNCH(KC_SYNTH, st, k_synth<<<grid, kSynThreads, 0, st>>>(*tab, *tr, a0k, a0mom, det, s_hat, part, nblk));
    LAUNCH(KC_REDUCE, st, k_reduce_parts<<<tab->U, kSynThreads, 0, st>>>(*tab, part, nblk, mean, true));
    LAUNCH(KC_VARPASS, st, k_var_pass<<<grid, kSynThreads, 0, st>>>(*tab, det, s_hat, mean, part, nblk));
    LAUNCH(KC_REDUCE, st, k_reduce_parts<<<tab->U, kSynThreads, 0, st>>>(*tab, part, nblk, m2, false));
    return EAQHM_OK;
}

extern "C" int eaqhm_snapshot(const eaqhm_tables* tab, const uint8_t* improved, const eaqhm_tracks* tr,
                              const eaqhm_knots* knots, const double* fm_fin, const double* a0mom,
                              const double* s_hat, double* s_best, const eaqhm_knots* best, void* stream) {
    CHECK_ARG(tab && improved && tr && knots && fm_fin && a0mom && s_hat && s_best && best);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 g1((tab->G + 127) / 128, tab->Kmax);
    LAUNCH(KC_SNAPSHOT, st, k_snapshot_knots<<<g1, 128, 0, st>>>(*tab, improved, *tr, fm_fin, knots->a0, a0mom, *best));
    dim3 g2(synth_nblk(tab->maxlen), tab->U);
    LAUNCH(KC_SNAPSHOT, st, k_snapshot_signal<<<g2, kSynThreads, 0, st>>>(*tab, improved, s_hat, s_best));
    return EAQHM_OK;
}
