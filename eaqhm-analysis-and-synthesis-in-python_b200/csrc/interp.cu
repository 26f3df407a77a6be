// Stage B of one eaQHM adaptation: interpolation of the per-frame parameters between frame
// centres (reference: functions.py:337-383 and phase_integr_interpolation :537-575).
#include <stdlib.h>

#include "common.cuh"

namespace {

// forward-elimination factors of the constant tridiagonal (1,4,1) system of a uniform-knot cubic
// spline: cp[0] = 1/4, cp[i] = 1/(4 - cp[i-1]); converges to 2 - sqrt(3) in < 40 steps.
constexpr int kCpTab = 48;
__constant__ double c_cp[kCpTab];

__device__ __forceinline__ double cp_at(int i) { return c_cp[i < kCpTab ? i : kCpTab - 1]; }

// Second derivatives ("moments") of the not-a-knot cubic spline through uniformly spaced knots (what
// interp1d(kind=3) == make_interp_spline(k=3) builds; functions.py:340,367).  Reduced system: with
// M_0 = 2 M_1 - M_2 the first interior equation gives M_1 = r_1/6 (same at the far end); the remaining
// unknowns obey M_{j-1} + 4 M_j + M_{j+1} = r_j, r_j = 6 (d_j - d_{j-1}) / h.
//
// Parallel form: the influence of a right-hand side decays like (2 - sqrt(3))^distance = 0.268^d, so a
// thread that owns a chunk of kChunk knots solves the system on its chunk extended by kHalo knots on each
// side (with an arbitrary Dirichlet value at an artificial cut); 0.268^30 = 7e-18 is below the rounding of the
// sequential solve.  True run ends inside the window get the exact not-a-knot closure.
constexpr int kChunk = 32;
constexpr int kHalo = 30;
constexpr int kLocalMax = kChunk + 2 * (kHalo + 2) + 4;

template <class Acc>
__device__ void chunk_moments(const double* __restrict__ y, double* __restrict__ m, int g0, int gend, int a, int b,
                              double h, Acc acc) {
    auto rhs = [&](int j) {
        const double d0 = (y[j] - y[j - 1]) / h, d1 = (y[j + 1] - y[j]) / h;
        return 6.0 * (d1 - d0) / h;
    };
    double loc[kLocalMax];
    const int lo_lim = max(g0, a - kHalo - 2), hi_lim = min(gend, b + kHalo + 2);
    int q = a;
    while (q < b) {
        if (!acc(q)) { q++; continue; }
        int ls = q;
        while (ls > lo_lim && acc(ls - 1)) ls--;
        const bool true_left = (ls == g0) || !acc(ls - 1);
        int re = q + 1;
        while (re < hi_lim && acc(re)) re++;
        const bool true_right = (re == gend) || !acc(re);
        if (!(true_left && true_right && re - ls < 4)) {
            const double B0 = true_left ? rhs(ls + 1) / 6.0 : 0.0;
            const double B1 = true_right ? rhs(re - 2) / 6.0 : 0.0;
            const int u0 = true_left ? ls + 2 : ls, u1 = true_right ? re - 3 : re - 1;
            if (true_left) loc[1] = B0;
            if (true_right) loc[re - 2 - ls] = B1;
            double dprev = 0.0;
            for (int j = u0; j <= u1; j++) {
                double r = rhs(j);
                if (j == u0) r -= B0;
                if (j == u1) r -= B1;
                const double d = (j == u0) ? r * 0.25 : (r - dprev) * cp_at(j - u0);
                loc[j - ls] = d;
                dprev = d;
            }
            for (int j = u1 - 1; j >= u0; j--) loc[j - ls] = loc[j - ls] - cp_at(j - u0) * loc[j + 1 - ls];
            if (true_left) loc[0] = 2.0 * loc[1] - loc[2];
            if (true_right) loc[re - 1 - ls] = 2.0 * loc[re - 2 - ls] - loc[re - 3 - ls];
            for (int j = max(ls, a); j < min(re, b); j++) m[j] = loc[j - ls];
        }
        q = re;
    }
}

// one thread per (harmonic k, chunk of frames): moments of the frequency spline of every run of >= 4 knots
__global__ void __launch_bounds__(128) k_run_moments(eaqhm_tables t, eaqhm_knots kn, double* mom) {
    const int k = blockIdx.y;
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) * kChunk;
    if (a >= t.G) return;
    const double* am = kn.am + (int64_t)k * t.G;
    const double* fm = kn.fm + (int64_t)k * t.G;
    double* mo = mom + (int64_t)k * t.G;
    const double h = (double)t.step;
    // a chunk may straddle utterances: handle each utterance's part separately
    int q = a;
    const int bend = min(a + kChunk, t.G);
    while (q < bend) {
        const int u = t.utt[q];
        const int g0 = t.gfirst[u], gend = g0 + t.gcount[u];
        const int b = min(bend, gend);
        if (t.active[u]) chunk_moments(fm, mo, g0, gend, q, b, h, [&](int j) { return am[j] != 0.0; });
        q = b;
    }
}

// DC spline through ALL frames of an utterance (functions.py:340)
__global__ void __launch_bounds__(128) k_a0_moments(eaqhm_tables t, const double* a0k, double* a0mom) {
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) * kChunk;
    if (a >= t.G) return;
    const double h = (double)t.step;
    int q = a;
    const int bend = min(a + kChunk, t.G);
    while (q < bend) {
        const int u = t.utt[q];
        const int g0 = t.gfirst[u], gend = g0 + t.gcount[u];
        const int b = min(bend, gend);
        if (t.active[u]) chunk_moments(a0k, a0mom, g0, gend, q, b, h, [](int) { return true; });
        q = b;
    }
}

// numpy.unwrap applied to one difference (period 2*pi), returned as the corrected difference
__device__ __forceinline__ double unwrap_diff(double dd) {
    double ddmod = fmod(dd + kPi, kTwoPi);
    if (ddmod < 0.0) ddmod += kTwoPi;
    ddmod -= kPi;
    if (ddmod == -kPi && dd > 0.0) ddmod = kPi;
    double corr = ddmod - dd;
    if (fabs(dd) < kPi) corr = 0.0;
    return dd + corr;
}

// cubic through 4 points in Lagrange form (short-run quirk Q4, functions.py:369-371)
__device__ __forceinline__ double lagrange4(const double* X, const double* Y, double x) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double num = 1.0, den = 1.0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j != i) { num *= (x - X[j]); den *= (X[i] - X[j]); }
        acc += Y[i] * (num / den);
    }
    return acc;
}

// one thread per (harmonic k, frame g): owns samples [pos, pos+own) of the three dense tracks
__global__ void __launch_bounds__(128) k_interp(eaqhm_tables t, eaqhm_knots kn, eaqhm_tracks tr, const double* mom,
                                               double* fm_fin) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (g >= t.G) return;
    const int u = t.utt[g];
    if (!t.active[u]) return;
    const int64_t G = t.G, T = t.T;
    const int64_t x0 = t.pos[g];
    const int own = t.own[g];
    const int gl = g - t.gfirst[u], nu = t.gcount[u];
    const double* amk = kn.am + (int64_t)k * G;
    const double* fmk = kn.fm + (int64_t)k * G;
    const double* phk = kn.ph + (int64_t)k * G;
    double* am = tr.am + (int64_t)k * T;
    double* ph = tr.ph + (int64_t)k * T;
    double* fc = tr.fmc + (int64_t)k * T;
    const bool zz = tr.zc != nullptr && tr.zs != nullptr;   // phasor tracks are optional (NULL: analysis takes its slow path)
    double* zc = zz ? tr.zc + (int64_t)k * T : nullptr;     // phasor of the phase track (cos, sin): the next adaptation's
    double* zs = zz ? tr.zs + (int64_t)k * T : nullptr;     // basis inside a run is a ratio of two of these (k_basis)
    auto acc = [&](int d) { const int j = gl + d; return j >= 0 && j < nu && amk[g + d] != 0.0; };
    const double y0 = amk[g];
    const bool a0 = y0 != 0.0, has_prev = a0 && acc(-1), has_next = a0 && acc(1);

    if (!has_next) {
        // not a segment start: zero fill; an isolated accepted frame stays a one-sample impulse
        // (functions.py:350-358 never interpolates it); the last knot of a run is written by its
        // predecessor's thread.
        for (int s = has_prev ? 1 : 0; s < own; s++) {
            am[x0 + s] = 0.0; ph[x0 + s] = 0.0; fc[x0 + s] = 0.0;
            if (zz) { zc[x0 + s] = 0.0; zs[x0 + s] = 0.0; }
        }
        if (a0 && !has_prev) { am[x0] = y0; ph[x0] = phk[g]; }
        if (!has_prev) fm_fin[(int64_t)k * G + g] = a0 ? fmk[g] : 0.0;
        return;
    }

    // segment g -> g+1 of a run
    const int D = t.step;
    const double h = (double)D;
    const double y1 = amk[g + 1];
    const bool next_last = !acc(2);
    const int nback = has_prev ? (acc(-2) ? 2 : 1) : 0;
    const int nfwd = next_last ? 0 : (acc(3) ? 2 : 1);
    const bool shortrun = (nback + 2 + nfwd) < 4;
    double LX[4], LY[4];
    const int64_t xoff = t.xoff[u];                // absolute position of this utterance's first sample
    const double xl0 = (double)(x0 - t.ustart[u] + xoff);
    if (shortrun) {
        // knots: extra ones at ABSOLUTE samples 0, step (functions.py:369) carrying whatever fm_recon holds there
        // (the knot values of the signal's frames 0 and 1: zero, they are never analysed, unless this utterance
        // starts at sample 0 and holds them itself), then the run's own 2 or 3 knots
        const int gs = g - nback;                  // first knot of the run
        const int nrun = nback + 2 + nfwd;
        const int64_t gu = t.gfirst[u];
        int c = 0;
        for (int e = 0; e < 4 - nrun; e++) { LX[c] = (double)(e * D); LY[c] = xoff == 0 ? fmk[gu + e] : 0.0; c++; }
        for (int e = 0; e < nrun; e++) { LX[c] = (double)(t.pos[gs + e] - t.ustart[u] + xoff); LY[c] = fmk[gs + e]; c++; }
    }
    const double f0k = fmk[g], f1k = fmk[g + 1];
    const double m0 = shortrun ? 0.0 : mom[(int64_t)k * G + g], m1 = shortrun ? 0.0 : mom[(int64_t)k * G + g + 1];
    auto fm_at = [&](int s) {
        return shortrun ? lagrange4(LX, LY, xl0 + (double)s) : spline_eval(f0k, f1k, m0, m1, h, (double)s);
    };
    const double wscale = kTwoPi / t.fs;            // functions.py:373: 2*pi/fs*fm_recon
    // pass 1: end value of the running phase sum (functions.py:561-562)
    double csum = 0.0;
    for (int s = 0; s <= D; s++) csum += wscale * fm_at(s);
    const double w0 = wscale * fm_at(0);
    const double ph0 = phk[g], ph1 = phk[g + 1];
    const double shift = ph0 - w0;
    const double pend = csum + shift;
    const double Mr = rint((pend - ph1) / kTwoPi);                              // functions.py:564
    const double er = kPi * ((pend - ph1) - kTwoPi * Mr) / (double)(2 * D);    // functions.py:565
    // pass 2: corrected phase, amplitude, new instantaneous frequency (functions.py:364,566-571,375)
    const double slope = (y1 - y0) / h;
    const double fscale = t.fs / kTwoPi;
    double c2 = 0.0, e2 = 0.0, pprev = 0.0;
    for (int s = 0; s <= D; s++) {
        const double fv = fm_at(s);
        c2 += wscale * fv;
        e2 += sin((kPi * (double)s) / h) * er;
        const double p = (c2 + shift) - e2;
        if (s == 0) {
            if (!has_prev) fc[x0] = fv;                                        // functions.py:375 first sample
            fm_fin[(int64_t)k * G + g] = fv;
        } else if (s < D) {
            fc[x0 + s] = fscale * unwrap_diff(p - pprev);
        }
        if (s < D) {
            am[x0 + s] = (s == 0) ? y0 : slope * (double)s + y0;
            ph[x0 + s] = p;
            if (zz) {
                double sn, cs;
                sincos(p, &sn, &cs);
                zc[x0 + s] = cs; zs[x0 + s] = sn;
            }
            pprev = p;
        } else {
            double pnext;
            if (next_last) {
                pnext = p;
                am[x0 + D] = y1;
                ph[x0 + D] = p;
                if (zz) {
                    double sn, cs;
                    sincos(p, &sn, &cs);
                    zc[x0 + D] = cs; zs[x0 + D] = sn;
                }
                fm_fin[(int64_t)k * G + g + 1] = fv;
            } else {
                // the next segment restarts from the measured phase (functions.py:562,571)
                const double fn0 = shortrun ? fv : spline_eval(f1k, 0.0, m1, 0.0, h, 0.0);
                const double wn0 = wscale * fn0;
                pnext = wn0 + (ph1 - wn0);
            }
            fc[x0 + D] = fscale * unwrap_diff(pnext - pprev);
        }
    }
}

// Coalesced variant: a half-warp (16 lanes) per (harmonic k, frame g), lane = sample inside the frame, so that the
// five tracks are written 128 contiguous bytes at a time (k_interp above writes one sample per lane 15 samples apart).
// The running sums of functions.py:561,569 become half-warp scans (same terms, pairwise instead of sequential order:
// ~1e-16 relative); everything else is the arithmetic of k_interp.
__device__ __forceinline__ double hw_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ double hw_scan(double v, int sl, unsigned mask) {   // inclusive prefix sum over the half-warp
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double t = __shfl_up_sync(mask, v, o, 16);
        if (sl >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(128) k_interp_hw(eaqhm_tables t, eaqhm_knots kn, eaqhm_tracks tr, const double* mom,
                                                  double* fm_fin) {
    const int sl = threadIdx.x & 15;
    const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int k = blockIdx.y;
    if (g >= t.G) return;
    const int u = t.utt[g];
    if (!t.active[u]) return;
    const int64_t G = t.G, T = t.T;
    const int64_t x0 = t.pos[g];
    const int own = t.own[g];
    const int gl = g - t.gfirst[u], nu = t.gcount[u];
    const double* amk = kn.am + (int64_t)k * G;
    const double* fmk = kn.fm + (int64_t)k * G;
    const double* phk = kn.ph + (int64_t)k * G;
    double* am = tr.am + (int64_t)k * T;
    double* ph = tr.ph + (int64_t)k * T;
    double* fc = tr.fmc + (int64_t)k * T;
    const bool zz = tr.zc != nullptr && tr.zs != nullptr;
    double* zc = zz ? tr.zc + (int64_t)k * T : nullptr;
    double* zs = zz ? tr.zs + (int64_t)k * T : nullptr;
    auto acc = [&](int d) { const int j = gl + d; return j >= 0 && j < nu && amk[g + d] != 0.0; };
    const double y0 = amk[g];
    const bool a0 = y0 != 0.0, has_prev = a0 && acc(-1), has_next = a0 && acc(1);

    if (!has_next) {
        for (int s = sl; s < own; s += 16) {
            if (s == 0 && has_prev) continue;              // the last knot of a run was written by its predecessor
            const bool imp = s == 0 && a0;                 // isolated accepted frame: one-sample impulse
            am[x0 + s] = imp ? y0 : 0.0; ph[x0 + s] = imp ? phk[g] : 0.0; fc[x0 + s] = 0.0;
            if (zz) { zc[x0 + s] = 0.0; zs[x0 + s] = 0.0; }
        }
        if (sl == 0 && !has_prev) fm_fin[(int64_t)k * G + g] = a0 ? fmk[g] : 0.0;
        return;
    }

    const int D = t.step;
    const double h = (double)D;
    const double y1 = amk[g + 1];
    const bool next_last = !acc(2);
    const int nback = has_prev ? (acc(-2) ? 2 : 1) : 0;
    const int nfwd = next_last ? 0 : (acc(3) ? 2 : 1);
    const bool shortrun = (nback + 2 + nfwd) < 4;
    double LX[4], LY[4];
    const int64_t xoff = t.xoff[u];
    const double xl0 = (double)(x0 - t.ustart[u] + xoff);
    if (shortrun) {
        const int gs = g - nback;
        const int nrun = nback + 2 + nfwd;
        const int64_t gu = t.gfirst[u];
        int c = 0;
        for (int e = 0; e < 4 - nrun; e++) { LX[c] = (double)(e * D); LY[c] = xoff == 0 ? fmk[gu + e] : 0.0; c++; }
        for (int e = 0; e < nrun; e++) { LX[c] = (double)(t.pos[gs + e] - t.ustart[u] + xoff); LY[c] = fmk[gs + e]; c++; }
    }
    const double f0k = fmk[g], f1k = fmk[g + 1];
    const double m0 = shortrun ? 0.0 : mom[(int64_t)k * G + g], m1 = shortrun ? 0.0 : mom[(int64_t)k * G + g + 1];
    auto fm_at = [&](int s) {
        return shortrun ? lagrange4(LX, LY, xl0 + (double)s) : spline_eval(f0k, f1k, m0, m1, h, (double)s);
    };
    const double wscale = kTwoPi / t.fs, fscale = t.fs / kTwoPi;
    const int nblk = (D + 1 + 15) / 16;                    // samples s = 0..D in blocks of 16 lanes
    // pass 1: end value of the running phase sum (functions.py:561-562)
    double csum = 0.0, fv0 = 0.0;
    for (int b = 0; b < nblk; b++) {
        const int s = b * 16 + sl;
        const double fv = s <= D ? fm_at(s) : 0.0;
        if (b == 0) fv0 = fv;
        csum += s <= D ? wscale * fv : 0.0;
    }
    csum = hw_sum(csum, mask);
    const double w0 = wscale * __shfl_sync(mask, fv0, 0, 16);
    const double ph0 = phk[g], ph1 = phk[g + 1];
    const double shift = ph0 - w0;
    const double pend = csum + shift;
    const double Mr = rint((pend - ph1) / kTwoPi);                              // functions.py:564
    const double er = kPi * ((pend - ph1) - kTwoPi * Mr) / (double)(2 * D);    // functions.py:565
    const double slope = (y1 - y0) / h;
    // pass 2: corrected phase, amplitude, new instantaneous frequency (functions.py:364,566-571,375)
    double c_carry = 0.0, e_carry = 0.0, p_carry = 0.0;
    for (int b = 0; b < nblk; b++) {
        const int s = b * 16 + sl;
        const bool in = s <= D;
        const double fv = (nblk == 1) ? fv0 : (in ? fm_at(s) : 0.0);
        const double c2 = hw_scan(in ? wscale * fv : 0.0, sl, mask) + c_carry;
        const double e2 = hw_scan(in ? sin((kPi * (double)s) / h) * er : 0.0, sl, mask) + e_carry;
        const double p = (c2 + shift) - e2;
        double pprev = __shfl_up_sync(mask, p, 1, 16);
        if (sl == 0) pprev = p_carry;
        c_carry = __shfl_sync(mask, c2, 15, 16);
        e_carry = __shfl_sync(mask, e2, 15, 16);
        p_carry = __shfl_sync(mask, p, 15, 16);
        if (!in) continue;
        if (s == 0) {
            if (!has_prev) fc[x0] = fv;                                        // functions.py:375 first sample
            fm_fin[(int64_t)k * G + g] = fv;
        } else if (s < D) {
            fc[x0 + s] = fscale * unwrap_diff(p - pprev);
        }
        if (s < D) {
            am[x0 + s] = (s == 0) ? y0 : slope * (double)s + y0;
            ph[x0 + s] = p;
            if (zz) { double sn, cs; sincos(p, &sn, &cs); zc[x0 + s] = cs; zs[x0 + s] = sn; }
        } else {
            double pnext;
            if (next_last) {
                pnext = p;
                am[x0 + D] = y1;
                ph[x0 + D] = p;
                if (zz) { double sn, cs; sincos(p, &sn, &cs); zc[x0 + D] = cs; zs[x0 + D] = sn; }
                fm_fin[(int64_t)k * G + g + 1] = fv;
            } else {
                const double fn0 = shortrun ? fv : spline_eval(f1k, 0.0, m1, 0.0, h, 0.0);
                const double wn0 = wscale * fn0;
                pnext = wn0 + (ph1 - wn0);
            }
            fc[x0 + D] = fscale * unwrap_diff(pnext - pprev);
        }
    }
}

// dense-input version of the phase integration for the helper seam (functions.py:537-575)
__global__ void k_phase_integr_dense(const double* fm, const double* phc, const int64_t* idx, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int64_t i0 = idx[i], i1 = idx[i + 1], base = idx[0];
    const int D = (int)(i1 - i0);
    double csum = 0.0;
    for (int s = 0; s <= D; s++) csum += fm[i0 + s];
    const double shift = phc[i0] - fm[i0];
    const double pend = csum + shift;
    const double Mr = rint((pend - phc[i1]) / kTwoPi);
    const double er = kPi * ((pend - phc[i1]) - kTwoPi * Mr) / (double)(2 * D);
    double c2 = 0.0, e2 = 0.0;
    for (int s = 0; s <= D; s++) {
        c2 += fm[i0 + s];
        e2 += sin((kPi * (double)s) / (double)D) * er;
        if (s < D || i == n - 2) out[i0 - base + s] = (c2 + shift) - e2;
    }
}

DeviceOnce g_cp_once;
int ensure_cp_table() {
    const int dev = g_cp_once.claim();          // __constant__ memory is per device
    if (dev == -1) return EAQHM_OK;
    CHECK_ARG(dev >= 0);
    double cp[kCpTab];
    cp[0] = 0.25;
    for (int i = 1; i < kCpTab; i++) cp[i] = 1.0 / (4.0 - cp[i - 1]);
    CUDA_TRY(cudaMemcpyToSymbol(c_cp, cp, sizeof(cp)));
    return EAQHM_OK;
}

}  // namespace

extern "C" int64_t eaqhm_interp_work_bytes(int32_t G, int32_t Kmax) {
    if (G < 0 || Kmax <= 0) return EAQHM_E_ARG;
    return align_up((int64_t)G * Kmax * 8, 256);
}

extern "C" int eaqhm_interpolate(const eaqhm_tables* tab, const eaqhm_knots* knots, const eaqhm_tracks* out,
                                 double* fm_fin, double* a0mom, void* work, int64_t work_bytes, void* stream) {
    CHECK_ARG(tab && knots && out && fm_fin && a0mom && work);
    CHECK_ARG(out->am && out->ph && out->fmc && (out->zc == nullptr) == (out->zs == nullptr));
    CHECK_ARG(work_bytes >= eaqhm_interp_work_bytes(tab->G, tab->Kmax));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_cp_table();
    if (rc != EAQHM_OK) return rc;
    double* mom = (double*)work;
    const int nchunk = (tab->G + kChunk - 1) / kChunk;
    dim3 gm((nchunk + 127) / 128, tab->Kmax);
    LAUNCH(KC_RUN_MOMENTS, st, k_run_moments<<<gm, 128, 0, st>>>(*tab, *knots, mom));
    LAUNCH(KC_A0_MOMENTS, st, k_a0_moments<<<(nchunk + 127) / 128, 128, 0, st>>>(*tab, knots->a0, a0mom));
    // experiment knob (tools only): EAQHM_INTERP_VARIANT=1 selects the thread-per-frame kernel
    static const int variant = getenv("EAQHM_INTERP_VARIANT") ? atoi(getenv("EAQHM_INTERP_VARIANT")) : 0;
    if (variant == 1) {
        dim3 grid((tab->G + 127) / 128, tab->Kmax);
        LAUNCH(KC_INTERP, st, k_interp<<<grid, 128, 0, st>>>(*tab, *knots, *out, mom, fm_fin));
    } else {
        dim3 grid((tab->G + 7) / 8, tab->Kmax);           // 8 half-warps (frames) per CTA
        LAUNCH(KC_INTERP, st, k_interp_hw<<<grid, 128, 0, st>>>(*tab, *knots, *out, mom, fm_fin));
    }
    return EAQHM_OK;
}

extern "C" int eaqhm_phase_integr_interpolation(const double* fm_recon, const double* ph_recon, int64_t length,
                                                const int64_t* indices, int32_t n, double* out) {
    CHECK_ARG(fm_recon && ph_recon && indices && out && n >= 2 && length > 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        eaqhm_set_error("no CUDA device: the eaQHM library has no CPU fallback");
        return EAQHM_E_NODEVICE;
    }
    for (int i = 0; i < n; i++) CHECK_ARG(indices[i] >= 0 && indices[i] < length && (i == 0 || indices[i] > indices[i - 1]));
    const int64_t nout = indices[n - 1] - indices[0] + 1;
    double *d_fm = nullptr, *d_ph = nullptr, *d_out = nullptr;
    int64_t* d_idx = nullptr;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(&d_fm, length * 8)); CUDA_TRY(cudaMalloc(&d_ph, length * 8));
        CUDA_TRY(cudaMalloc(&d_out, nout * 8)); CUDA_TRY(cudaMalloc(&d_idx, (size_t)n * 8));
        CUDA_TRY(cudaMemcpy(d_fm, fm_recon, length * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_ph, ph_recon, length * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_idx, indices, (size_t)n * 8, cudaMemcpyHostToDevice));
        LAUNCH(KC_INTERP, 0, k_phase_integr_dense<<<(n - 1 + 127) / 128, 128>>>(d_fm, d_ph, d_idx, n, d_out));
        CUDA_TRY(cudaMemcpy(out, d_out, nout * 8, cudaMemcpyDeviceToHost));
        return EAQHM_OK;
    };
    const int rc = body();
    cudaFree(d_fm); cudaFree(d_ph); cudaFree(d_out); cudaFree(d_idx);
    return rc;
}
