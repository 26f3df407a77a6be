// Stage A of one eaQHM adaptation: per-frame basis construction, weighted complex Gram,
// Cholesky solve, frequency mismatch and acceptance (reference: functions.py:176-329,
// iqhmLS_complexamps :420-470, eaqhmLS_complexamps :472-535).  sm_100a, FP64.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kBasisThreads = 256;
constexpr int kBasisWarps = kBasisThreads / 32;
constexpr int kNb = 16;   // Cholesky block width
// A pivot below kPivotTol x the matrix's own diagonal entry means the normal matrix is singular to working precision
// (legitimate frames: cond(R) <= ~2e7, pivots >= ~1e-8 of the diagonal); the frame is flagged and the host raises
constexpr double kPivotTol = 1e-13;

struct AnaArgs {
    eaqhm_tables t;
    const double* s;
    const double* am_cur;
    const double* fm_cur;
    const double* zc_cur;           // cos / sin of the phase track (eaqhm_tracks.zc, .zs)
    const double* zs_cur;
    eaqhm_knots k;
    double* best_am;
    int32_t adapt;
    int32_t j0, nf;                 // chunk of the analysed list: [j0, j0+nf)
    int32_t nnz_max, Mxmax, Pmax, Pp, Lp;
    int32_t nwb;                    // warps of the basis kernel that own a column buffer (smem bound for long windows)
    uint8_t* fb;                    // [G] fallback flags of this adaptation
    int32_t* Mcols;                 // [chunk] generic columns of E2 per frame (2*nnz; the DC column is kept apart)
    int32_t* Lrows;                 // [chunk] window samples per frame
    int16_t* act;                   // [chunk][nnz_max] active harmonic indices
    double* X;                      // [chunk][2 planes][Mxmax][Lp]  windowed basis w*E2, real and imaginary planes
    double2* R;                     // [chunk][Pmax+1][Pp]   (row P = conj(rhs))
    double2* x;                     // [chunk][Pmax]
    int32_t* status;
    double* flop_stats;             // [3] += algorithmic flops of the Gram / Cholesky / basis+rhs (may be null)
    uint8_t* slow;                  // [chunk][nnz_max] columns the fast basis kernel left to k_basis<true> (gaps, run ends, fallback)
    int32_t slow_only;              // k_basis<true> builds only the columns flagged in `slow`
    int32_t a0_real;                // adaptation 0 solved as two real systems (k_a0_build_real, k_chol_real)
    int32_t store_X;                // k_basis<false> writes the basis planes (only the dense-Gram path reads them)
};

// Real even / odd systems of adaptation 0 inside a frame's scratch matrix (see k_a0_build_real)
struct A0Real {
    int M, Mp;                      // unknowns per system (2K+1), row pitch in doubles
    double* sys[2];                 // augmented matrices [M+1][Mp] (row M = right-hand side)
    double* sol[2];                 // solutions [M]
};
__device__ __forceinline__ A0Real a0_real_layout(double2* RF, int Mx) {
    A0Real r;
    r.M = Mx + 1;
    r.Mp = (r.M + 15) & ~15;
    double* base = reinterpret_cast<double*>(RF);
    r.sys[0] = base;
    r.sys[1] = base + (size_t)(r.M + 1) * r.Mp;
    r.sol[0] = base + (size_t)2 * (r.M + 1) * r.Mp;
    r.sol[1] = r.sol[0] + r.Mp;
    return r;
}

// ---------------------------------------------------------------------------------------
// fallback flags over ALL analysed frames (functions.py:204-210, quirk Q6): a frame whose
// centre row of fm_current is entirely zero analyses a single 140 Hz / 10e-4 component.
// The in-place amplitude write also lands in the previous adaptation's am_recon, which is
// the am_fin snapshot (functions.py:383,400) -> best_am.
// ---------------------------------------------------------------------------------------
__global__ void k_fallback_flags(AnaArgs A) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int32_t* list = A.t.flist ? A.t.flist : A.t.alist;
    const int nlist = A.t.flist ? A.t.nF : A.t.nA;
    if (w >= nlist) return;
    const int g = list[w];
    const int64_t c = A.t.pos[g];
    bool any = false;
    for (int k = lane; k < A.t.Kmax; k += 32) any |= (A.fm_cur[(int64_t)k * A.t.T + c] != 0.0);
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        A.fb[g] = any ? 0 : 1;
        if (!any) {
            atomicAdd(&A.status[1], 1);
            if (A.best_am) A.best_am[(int64_t)g * A.t.Kmax] = kFallbackAm;   // harmonic 0 of frame g (snapshot is [G][Kmax])
        }
    }
}

// active set of each frame of the chunk (functions.py:202, 246-248); adaptation 0: all K0 harmonics
__global__ void k_active(AnaArgs A) {
    const int jl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (jl >= A.nf) return;
    const int g = A.t.alist[A.j0 + jl];
    int16_t* act = A.act + (int64_t)jl * A.nnz_max;
    int count = 0;
    if (A.adapt == 0) {
        count = A.t.K0[g];
        for (int k = lane; k < count; k += 32) act[k] = (int16_t)k;
    } else if (A.fb[g]) {
        count = 1;
        if (lane == 0) act[0] = 0;
    } else {
        const int64_t c = A.t.pos[g];
        for (int k0 = 0; k0 < A.t.Kmax; k0 += 32) {
            const int k = k0 + lane;
            const bool nzf = k < A.t.Kmax && A.fm_cur[(int64_t)k * A.t.T + c] != 0.0;
            const unsigned b = __ballot_sync(0xffffffffu, nzf);
            if (nzf) {
                const int slot = count + __popc(b & ((1u << lane) - 1u));
                if (slot < A.nnz_max) act[slot] = (int16_t)k;
            }
            count += __popc(b);
        }
        if (count > A.nnz_max) count = A.nnz_max;   // cannot happen: active set only shrinks
    }
    if (lane == 0) {
        A.Mcols[jl] = 2 * count;
        A.Lrows[jl] = 2 * A.t.N[g] + 1;
        if (A.flop_stats) {   // structure-aware minimum work of this frame
            const double M = 2.0 * count + 1.0, L = 2.0 * A.t.N[g] + 1.0, P = 2.0 * M;
            if (A.adapt > 0) atomicAdd(&A.flop_stats[0], 12.0 * L * M * M);   // adaptation 0 uses the Toeplitz tables
            if (A.a0_real) atomicAdd(&A.flop_stats[1], 2.0 * ((1.0 / 3.0) * M * M * M + 2.0 * M * M));   // two real systems
            else atomicAdd(&A.flop_stats[1], (4.0 / 3.0) * P * P * P + 8.0 * P * P);
            atomicAdd(&A.flop_stats[2], 16.0 * L * M);
        }
    }
}

// ---------------------------------------------------------------------------------------
// basis: one CTA per frame, one warp per harmonic column pair (+f, time-reversed -f), plus DC.
// Generic columns of E2 (stored windowed, X = w*E2, as real / imaginary planes [col][n]):
//   [0,nnz) positive harmonics, [nnz,2nnz) their mirrored partners.  The DC column is E2 = 1, i.e.
//   X = w (real); its Gram rows are plain weighted sums of the other columns and are written here.
// Unknown order of the LS system (P = 2*Mx + 2):
//   [0,Mx) amplitudes, [Mx,2Mx) slopes of the generic columns, 2Mx = DC amplitude, 2Mx+1 = DC slope.
// Rows written here into the augmented matrix R~ = [[R, b],[b^H, .]]: 2Mx, 2Mx+1 (DC) and P (conj(b)).
// ---------------------------------------------------------------------------------------
struct ColSums {   // weighted sums of one column X over the window
    double bar, bai, bsr, bsi;          // sum X*ws, sum n*X*ws                 (ws = w*s)     -> rhs
    double d0r, d0i, d1r, d1i, d2r, d2i; // sum X*w, sum n*X*w, sum n^2*X*w                    -> DC rows
    __device__ __forceinline__ void zero() { bar = bai = bsr = bsi = d0r = d0i = d1r = d1i = d2r = d2i = 0.0; }
    __device__ __forceinline__ void add(double xr, double xi, double nc, double wn, double v) {
        bar += xr * v; bai += xi * v;
        const double nxr = nc * xr, nxi = nc * xi;
        bsr += nxr * v; bsi += nxi * v;
        d0r += xr * wn; d0i += xi * wn;
        d1r += nxr * wn; d1i += nxi * wn;
        d2r += nc * nxr * wn; d2i += nc * nxi * wn;
    }
    __device__ __forceinline__ void reduce() {
        bar = warp_sum(bar); bai = warp_sum(bai); bsr = warp_sum(bsr); bsi = warp_sum(bsi);
        d0r = warp_sum(d0r); d0i = warp_sum(d0i); d1r = warp_sum(d1r); d1i = warp_sum(d1i);
        d2r = warp_sum(d2r); d2i = warp_sum(d2i);
    }
    // rows of R~: Rb = row P (conj rhs), Rd0 = row 2Mx, Rd1 = row 2Mx+1 (null without DC)
    __device__ __forceinline__ void store(int col, int Mx, double2* Rb, double2* Rd0, double2* Rd1) const {
        Rb[col] = make_double2(bar, bai);
        Rb[Mx + col] = make_double2(bsr, bsi);
        if (Rd0) {
            Rd0[col] = make_double2(d0r, d0i);       Rd0[Mx + col] = make_double2(d1r, d1i);
            Rd1[col] = make_double2(d1r, d1i);       Rd1[Mx + col] = make_double2(d2r, d2i);
        }
    }
};

template <bool ADAPT>
__global__ void __launch_bounds__(kBasisThreads) k_basis(AnaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jl = blockIdx.x;
    if (ADAPT && A.slow_only) {     // nothing left to this kernel in most frames: leave before touching shared memory
        __shared__ int any_slow;
        if (tid == 0) any_slow = 0;
        __syncthreads();
        const int nzq = A.Mcols[jl] / 2;
        for (int it = tid; it < nzq; it += kBasisThreads) if (A.slow[(size_t)jl * A.nnz_max + it]) any_slow = 1;
        __syncthreads();
        if (!any_slow) return;
    }
    const int g = A.t.alist[A.j0 + jl];
    const int N = A.t.N[g], L = 2 * N + 1;
    const int Mx = A.Mcols[jl], nz = Mx / 2, P = 2 * Mx + 2;
    const int u = A.t.utt[g];
    const int64_t us = A.t.ustart[u], ul = A.t.ulen[u];
    const int64_t cl = A.t.pos[g] - us;
    const double fs = A.t.fs;

    double* w = reinterpret_cast<double*>(smem_raw);
    double* ws = w + A.Lp;
    double* fmw = ws + A.Lp + (size_t)warp * 2 * A.Lp;
    double* amw = fmw + A.Lp;
    int16_t* prv = reinterpret_cast<int16_t*>(ws + A.Lp + (size_t)A.nwb * 2 * A.Lp) + (size_t)warp * 2 * A.Lp;
    int16_t* nxt = prv + A.Lp;

    // analysis window (numpy.blackman / numpy.hamming, functions.py:194,200) and windowed signal
    for (int n = tid; n < L; n += kBasisThreads) {
        const double nn = (double)(1 - L + 2 * n), den = (double)(L - 1);
        const double x1 = (kPi * nn) / den;
        double wn;
        if (ADAPT) wn = 0.54 + 0.46 * cos(x1);
        else wn = 0.42 + 0.5 * cos(x1) + 0.08 * cos(((2.0 * kPi) * nn) / den);
        int64_t il = cl - N + n;
        if (il < 0) il += ul;                 // the reference's negative fancy index wraps
        w[n] = wn;
        ws[n] = wn * A.s[us + il];
    }
    __syncthreads();

    double* Xr = A.X + (size_t)jl * 2 * A.Mxmax * A.Lp;
    double* Xi = Xr + (size_t)A.Mxmax * A.Lp;
    double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    double2* Rb = RF + (size_t)P * A.Pp;
    double2* Rd0 = RF + (size_t)(2 * Mx) * A.Pp;
    double2* Rd1 = RF + (size_t)(2 * Mx + 1) * A.Pp;
    const int16_t* act = A.act + (int64_t)jl * A.nnz_max;
    const double f0 = A.t.f0[g];

    for (int it = warp; it <= nz && warp < A.nwb; it += A.nwb) {
        if (ADAPT && A.slow_only && (it == nz || !A.slow[(size_t)jl * A.nnz_max + it])) continue;   // built by k_basis_fast
        if (it == nz) {   // DC column: E2 = 1 (fm = 0, am = 0 -> (eps+0)/(0+eps) = 1), X = w
            double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
            for (int n = lane; n < L; n += 32) {
                const double wn = w[n], nc = (double)(n - N), v = ws[n];
                s0 += wn * v;        s1 += nc * wn * v;
                q0 += wn * wn;       q1 += nc * wn * wn;      q2 += nc * nc * wn * wn;
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1); q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
            if (lane == 0) {
                Rb[2 * Mx] = make_double2(s0, 0.0);       Rb[2 * Mx + 1] = make_double2(s1, 0.0);
                Rd0[2 * Mx] = make_double2(q0, 0.0);
                Rd1[2 * Mx] = make_double2(q1, 0.0);      Rd1[2 * Mx + 1] = make_double2(q2, 0.0);
            }
            continue;
        }
        const int k = act[it];
        double* __restrict__ pr = Xr + (size_t)it * A.Lp;
        double* __restrict__ pi = Xi + (size_t)it * A.Lp;
        double* __restrict__ nr = Xr + (size_t)(nz + it) * A.Lp;
        double* __restrict__ ni = Xi + (size_t)(nz + it) * A.Lp;
        ColSums cp, cn;
        cp.zero(); cn.zero();

        if (!ADAPT) {
            // functions.py:453-455: t = (n*2*pi*f)/fs with f = (k+1)*f0; the -f column is the conjugate
            const double fk = (double)(k + 1) * f0;
            for (int n = lane; n < L; n += 32) {
                const double nc = (double)(n - N);
                const double t = (((nc * 2.0) * kPi) * fk) / fs;
                double sn, cs;
                sincos(t, &sn, &cs);
                const double wn = w[n], xr = cs * wn, xi = sn * wn, v = ws[n];
                if (A.store_X) { pr[n] = xr; pi[n] = xi; nr[n] = xr; ni[n] = -xi; }
                cp.add(xr, xi, nc, wn, v);
                cn.add(xr, -xi, nc, wn, v);
            }
        } else {
            // 1. gather the window of fm_current / am_current (functions.py:244-245); column 0 sees the
            //    in-place fallback writes of this and earlier frames (Q6, functions.py:209-210)
            const double* __restrict__ fmk = A.fm_cur + (int64_t)k * A.t.T + us;
            const double* __restrict__ amk = A.am_cur + (int64_t)k * A.t.T + us;
            bool clean = A.zc_cur != nullptr && A.zs_cur != nullptr && cl - N >= 0;   // no wrapped indices, no zeros, no fallback substitution
            // four samples per lane and trip: the eight loads are issued before the first use (the kernel is bound by
            // the latency of these gathers: ncu source view, profiles/r01_ncu_summary_final.md)
            for (int n0 = lane; n0 < L; n0 += 128) {
                double fq[4], aq[4];
                int64_t ilq[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int n = n0 + 32 * q;
                    int64_t il = cl - N + (n < L ? n : 0);
                    if (il < 0) il += ul;
                    ilq[q] = il;
                    fq[q] = fmk[il]; aq[q] = amk[il];
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int n = n0 + 32 * q;
                    if (n >= L) break;
                    double f = fq[q], a = aq[q];
                    const int64_t il = ilq[q];
                    if (k == 0 && (il % A.t.step) == 0) {
                        const int64_t gi = A.t.gfirst[u] + il / A.t.step;
                        if (gi <= g && gi < A.t.gfirst[u] + A.t.gcount[u] && A.fb[gi]) { f = kFallbackFm; a = kFallbackAm; clean = false; }
                    }
                    clean &= (f != 0.0);
                    fmw[n] = f; amw[n] = a;
                }
            }
            clean = __all_sync(0xffffffffu, clean);
            __syncwarp();
            if (clean) {
                // Fast path: the window lies inside one run of this harmonic.  There fm_current[m] = fs/2pi *
                // unwrap(ph[m] - ph[m-1]) (functions.py:375), so the running sum telescopes and
                // E1[n] = exp(j 2 pi (S[n] - S[N]) / fs) = z[n] conj(z[N]) with z = exp(j ph_recon), the phasor track the
                // interpolation stage wrote: one complex product instead of the gap scan, the running sum and a
                // sincos per sample (the two forms differ by the rounding of the reference's own running sum, ~1e-14 rad).
                const double* __restrict__ zck = A.zc_cur + (int64_t)k * A.t.T + us + (cl - N);
                const double* __restrict__ zsk = A.zs_cur + (int64_t)k * A.t.T + us + (cl - N);
                double s_lo = 0.0;                           // S[N-1] = sum of fm over the first N samples (for n = L-1 below)
                for (int n = lane; n < N; n += 32) s_lo += fmw[n];
                s_lo = warp_sum(s_lo);
                s_lo = __shfl_sync(0xffffffffu, s_lo, 0);
                const double fmid = fmw[N], aden = amw[N] + kAmEps, ainv = 1.0 / aden;
                const double zcc = zck[N], zsc = zsk[N];
                __syncwarp();
                // amplitude ratio (functions.py:517-518) once per sample, kept in the amplitude buffer for the mirrored
                // column (multiplied by the reciprocal: one rounding away from the reference's division)
                for (int n0 = lane; n0 < L; n0 += 128) {
                    double cq[4], sq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int n = n0 + 32 * q, ns = n < L ? n : 0;
                        cq[q] = zck[ns]; sq[q] = zsk[ns];
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int n = n0 + 32 * q;
                        if (n >= L) break;
                        const double nc = (double)(n - N), wn = w[n], v = ws[n];
                        const double c1 = cq[q], s1 = sq[q];
                        const double cs = c1 * zcc + s1 * zsc, sn = s1 * zcc - c1 * zsc;
                        const double r = (kAmEps + amw[n]) * ainv;
                        const double xr = (r * cs) * wn, xi = (r * sn) * wn;
                        pr[n] = xr; pi[n] = xi;
                        cp.add(xr, xi, nc, wn, v);
                        fmw[n] = cs; ni[n] = sn; amw[n] = r;
                    }
                }
                __syncwarp();
                double rs, rc;
                sincos((kTwoPi * fmid) / fs, &rs, &rc);      // rho = exp(j 2 pi (S[N] - S[N-1]) / fs), S[N] - S[N-1] = fm[N]
                for (int n = lane; n < N; n += 32) {
                    const int n2 = L - 2 - n;
                    const double c1 = fmw[n], s1 = ni[n], c2 = fmw[n2], s2 = ni[n2];
                    {
                        const double er = c2 * rc - s2 * rs, ei = c2 * rs + s2 * rc;
                        const double r = amw[L - 1 - n], wn = w[n];
                        const double xr = (r * er) * wn, xi = (r * ei) * wn;
                        nr[n] = xr; ni[n] = xi;
                        cn.add(xr, xi, (double)(n - N), wn, ws[n]);
                    }
                    {
                        const double er = c1 * rc - s1 * rs, ei = c1 * rs + s1 * rc;
                        const double r = amw[n + 1], wn = w[n2];
                        const double xr = (r * er) * wn, xi = (r * ei) * wn;
                        nr[n2] = xr; ni[n2] = xi;
                        cn.add(xr, xi, (double)(n2 - N), wn, ws[n2]);
                    }
                }
                if (lane == 0) {
                    double sn, cs;
                    sincos((kTwoPi * (0.0 - s_lo)) / fs, &sn, &cs);
                    const double r = amw[0], wn = w[L - 1];
                    const double xr = (r * cs) * wn, xi = (r * sn) * wn;
                    nr[L - 1] = xr; ni[L - 1] = xi;
                    cn.add(xr, xi, (double)N, wn, ws[L - 1]);
                }
                cp.reduce(); cn.reduce();
                if (lane == 0) {
                    cp.store(it, Mx, Rb, Rd0, Rd1);
                    cn.store(nz + it, Mx, Rb, Rd0, Rd1);
                }
                __syncwarp();
                continue;
            }
            // 2. nearest non-zero neighbour on each side (functions.py:253-271)
            int carry = -1;
            for (int base = 0; base < L; base += 32) {
                const int n = base + lane;
                const bool nzf = n < L && fmw[n] != 0.0;
                const unsigned b = __ballot_sync(0xffffffffu, nzf);
                const unsigned m = b & (0xffffffffu >> (31 - lane));
                if (n < L) prv[n] = (int16_t)(m ? base + 31 - __clz(m) : carry);
                if (b) carry = base + 31 - __clz(b);
            }
            carry = -1;
            for (int base = ((L - 1) / 32) * 32; base >= 0; base -= 32) {
                const int n = base + lane;
                const bool nzf = n < L && fmw[n] != 0.0;
                const unsigned b = __ballot_sync(0xffffffffu, nzf);
                const unsigned m = b & (0xffffffffu << lane);
                if (n < L) nxt[n] = (int16_t)(m ? base + __ffs(m) - 1 : carry);
                if (b) carry = base + __ffs(b) - 1;
            }
            __syncwarp();
            // 3. edge hold + linear gap fill through numpy.interp's formula (functions.py:277-278)
            for (int n = lane; n < L; n += 32) {
                if (fmw[n] == 0.0) {
                    const int p = prv[n], q = nxt[n];
                    double f, a;
                    if (p < 0 && q < 0) { f = 0.0; a = amw[n]; }
                    else if (p < 0) { f = fmw[q]; a = amw[q]; }
                    else if (q < 0) { f = fmw[p]; a = amw[p]; }
                    else {
                        const double dx = (double)(q - p), xx = (double)(n - p);
                        f = ((fmw[q] - fmw[p]) / dx) * xx + fmw[p];
                        a = ((amw[q] - amw[p]) / dx) * xx + amw[p];
                    }
                    // neighbours p, q are non-zero samples: never written here; zero samples are read by their own lane only
                    fmw[n] = f; amw[n] = a;
                }
            }
            __syncwarp();
            // 4. running sum of the frequency track (functions.py:509-511; lfilter == cumsum)
            double run = 0.0;
            for (int base = 0; base < L; base += 32) {
                const int n = base + lane;
                double v = n < L ? fmw[n] : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += t;
                }
                v += run;
                if (n < L) fmw[n] = v;
                run = __shfl_sync(0xffffffffu, v, 31);
            }
            __syncwarp();
            // 5. +f column (functions.py:513-519): E1[n] = exp(j 2 pi (S[n] - S[N]) / fs).
            //    Time-reversed -f column (Q2, :284-285): the running sum of -fm[::-1] minus its centre value equals
            //    S[L-2-n] - S[N-1], so E1neg[n] = E1[L-2-n] * rho with rho = exp(j 2 pi (S[N] - S[N-1]) / fs): one rotation
            //    instead of a second sincos per sample.  Pass 1 keeps cos in the (now free) running-sum buffer and parks
            //    sin in the -f output row; pass 2 walks the index pairs (n, L-2-n), which only touch their own two slots.
            const double Smid = fmw[N], Smid1 = fmw[N - 1], aden = amw[N] + kAmEps;
            __syncwarp();
            for (int n = lane; n < L; n += 32) {
                const double nc = (double)(n - N), wn = w[n], v = ws[n];
                double sn, cs;
                sincos((kTwoPi * (fmw[n] - Smid)) / fs, &sn, &cs);
                const double r = (kAmEps + amw[n]) / aden;
                const double xr = (r * cs) * wn, xi = (r * sn) * wn;
                pr[n] = xr; pi[n] = xi;
                cp.add(xr, xi, nc, wn, v);
                fmw[n] = cs; ni[n] = sn;
            }
            __syncwarp();
            double rs, rc;
            sincos((kTwoPi * (Smid - Smid1)) / fs, &rs, &rc);
            for (int n = lane; n < N; n += 32) {
                const int n2 = L - 2 - n;                       // partner index, in [N, 2N-1]
                const double c1 = fmw[n], s1 = ni[n], c2 = fmw[n2], s2 = ni[n2];
                {   // -f column at n: E1[n2] * rho, amplitude from sample L-1-n
                    const double er = c2 * rc - s2 * rs, ei = c2 * rs + s2 * rc;
                    const double r = (kAmEps + amw[L - 1 - n]) / aden, wn = w[n];
                    const double xr = (r * er) * wn, xi = (r * ei) * wn;
                    nr[n] = xr; ni[n] = xi;
                    cn.add(xr, xi, (double)(n - N), wn, ws[n]);
                }
                {   // -f column at n2: E1[n] * rho, amplitude from sample L-1-n2 = n+1
                    const double er = c1 * rc - s1 * rs, ei = c1 * rs + s1 * rc;
                    const double r = (kAmEps + amw[n + 1]) / aden, wn = w[n2];
                    const double xr = (r * er) * wn, xi = (r * ei) * wn;
                    nr[n2] = xr; ni[n2] = xi;
                    cn.add(xr, xi, (double)(n2 - N), wn, ws[n2]);
                }
            }
            if (lane == 0) {   // n = L-1: running sum before the first sample is 0
                double sn, cs;
                sincos((kTwoPi * (0.0 - Smid1)) / fs, &sn, &cs);
                const double r = (kAmEps + amw[0]) / aden, wn = w[L - 1];
                const double xr = (r * cs) * wn, xi = (r * sn) * wn;
                nr[L - 1] = xr; ni[L - 1] = xi;
                cn.add(xr, xi, (double)N, wn, ws[L - 1]);
            }
        }
        cp.reduce(); cn.reduce();
        if (lane == 0) {
            cp.store(it, Mx, Rb, Rd0, Rd1);
            cn.store(nz + it, Mx, Rb, Rd0, Rd1);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// Adaptive basis, fast path only.  Inside one run of a harmonic the basis phasor is a ratio of two phasor-track values
// (see k_basis<true>), so a column needs no gap scan, no running sum and no per-warp column buffers: this kernel builds
// every such column with shared memory for the window alone (four resident CTAs per SM instead of two) and two passes
// over the samples, the second one re-reading the (L1-resident) track values for the mirrored column instead of parking
// them in shared memory.  A column that turns out not to be clean (a zero sample, the 140 Hz fallback, a wrapped
// window) is flagged in A.slow and rebuilt, together with its sums, by k_basis<true> in slow-only mode.  Same
// arithmetic, operation for operation, as the fast path of k_basis<true>.
// ---------------------------------------------------------------------------------------
#ifndef EAQHM_BASIS_FAST_CTAS
#define EAQHM_BASIS_FAST_CTAS 3
#endif
#ifndef EAQHM_BASIS_FAST_UNROLL
#define EAQHM_BASIS_FAST_UNROLL 4
#endif
__global__ void __launch_bounds__(kBasisThreads, EAQHM_BASIS_FAST_CTAS) k_basis_fast(AnaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jl = blockIdx.x;
    const int g = A.t.alist[A.j0 + jl];
    const int N = A.t.N[g], L = 2 * N + 1;
    const int Mx = A.Mcols[jl], nz = Mx / 2, P = 2 * Mx + 2;
    const int u = A.t.utt[g];
    const int64_t us = A.t.ustart[u], ul = A.t.ulen[u];
    const int64_t cl = A.t.pos[g] - us;
    const double fs = A.t.fs;
    double* w = reinterpret_cast<double*>(smem_raw);
    double* ws = w + A.Lp;
    for (int n = tid; n < L; n += kBasisThreads) {
        const double nn = (double)(1 - L + 2 * n), den = (double)(L - 1);
        const double wn = 0.54 + 0.46 * cos((kPi * nn) / den);                 // numpy.hamming (functions.py:200)
        int64_t il = cl - N + n;
        if (il < 0) il += ul;
        w[n] = wn;
        ws[n] = wn * A.s[us + il];
    }
    __syncthreads();
    double* Xr = A.X + (size_t)jl * 2 * A.Mxmax * A.Lp;
    double* Xi = Xr + (size_t)A.Mxmax * A.Lp;
    double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    double2* Rb = RF + (size_t)P * A.Pp;
    double2* Rd0 = RF + (size_t)(2 * Mx) * A.Pp;
    double2* Rd1 = RF + (size_t)(2 * Mx + 1) * A.Pp;
    const int16_t* act = A.act + (int64_t)jl * A.nnz_max;
    uint8_t* slow = A.slow + (size_t)jl * A.nnz_max;
    const bool frame_ok = cl - N >= 0;                                         // no wrapped indices

    for (int it = warp; it <= nz; it += kBasisWarps) {
        if (it == nz) {   // DC column: X = w
            double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
            for (int n = lane; n < L; n += 32) {
                const double wn = w[n], nc = (double)(n - N), v = ws[n];
                s0 += wn * v;        s1 += nc * wn * v;
                q0 += wn * wn;       q1 += nc * wn * wn;      q2 += nc * nc * wn * wn;
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1); q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
            if (lane == 0) {
                Rb[2 * Mx] = make_double2(s0, 0.0);       Rb[2 * Mx + 1] = make_double2(s1, 0.0);
                Rd0[2 * Mx] = make_double2(q0, 0.0);
                Rd1[2 * Mx] = make_double2(q1, 0.0);      Rd1[2 * Mx + 1] = make_double2(q2, 0.0);
            }
            continue;
        }
        const int k = act[it];
        if (!frame_ok) { if (lane == 0) slow[it] = 1; continue; }
        const int64_t off = (int64_t)k * A.t.T + us + (cl - N);
        const double* __restrict__ fmk = A.fm_cur + off;
        const double* __restrict__ amk = A.am_cur + off;
        const double* __restrict__ zck = A.zc_cur + off;
        const double* __restrict__ zsk = A.zs_cur + off;
        double* __restrict__ pr = Xr + (size_t)it * A.Lp;
        double* __restrict__ pi = Xi + (size_t)it * A.Lp;
        double* __restrict__ nr = Xr + (size_t)(nz + it) * A.Lp;
        double* __restrict__ ni = Xi + (size_t)(nz + it) * A.Lp;
        const double fmid = fmk[N], ainv = 1.0 / (amk[N] + kAmEps);
        const double zcc = zck[N], zsc = zsk[N];
        // pass 1: +f column, its sums, the zero test and S[N-1] = sum of fm over the first N samples
        bool clean = true;
        double s_lo = 0.0;
        ColSums cs_;
        cs_.zero();
        constexpr int kU = EAQHM_BASIS_FAST_UNROLL;             // samples per lane in flight
        for (int n0 = lane; n0 < L; n0 += 32 * kU) {
            double fq[kU], aq[kU], cq[kU], sq[kU];
#pragma unroll
            for (int q = 0; q < kU; q++) {
                const int n = n0 + 32 * q, ns = n < L ? n : 0;
                fq[q] = fmk[ns]; aq[q] = amk[ns]; cq[q] = zck[ns]; sq[q] = zsk[ns];
            }
#pragma unroll
            for (int q = 0; q < kU; q++) {
                const int n = n0 + 32 * q;
                if (n >= L) break;
                if (k == 0) {                                                   // in-place fallback of a frame centred here (Q6)
                    const int64_t il = cl - N + n;
                    if ((il % A.t.step) == 0) {
                        const int64_t gi = A.t.gfirst[u] + il / A.t.step;
                        if (gi <= g && gi < A.t.gfirst[u] + A.t.gcount[u] && A.fb[gi]) clean = false;
                    }
                }
                clean &= (fq[q] != 0.0);
                if (n < N) s_lo += fq[q];
                const double nc = (double)(n - N), wn = w[n], v = ws[n];
                const double c1 = cq[q], s1 = sq[q];
                const double cs = c1 * zcc + s1 * zsc, sn = s1 * zcc - c1 * zsc;
                const double r = (kAmEps + aq[q]) * ainv;
                const double xr = (r * cs) * wn, xi = (r * sn) * wn;
                pr[n] = xr; pi[n] = xi;
                cs_.add(xr, xi, nc, wn, v);
            }
        }
        clean = __all_sync(0xffffffffu, clean);
        if (!clean) { if (lane == 0) slow[it] = 1; continue; }
        s_lo = warp_sum(s_lo);
        s_lo = __shfl_sync(0xffffffffu, s_lo, 0);
        cs_.reduce();
        if (lane == 0) cs_.store(it, Mx, Rb, Rd0, Rd1);
        // pass 2: time-reversed -f column (Q2): E1neg[n] = E1[L-2-n] rho, rho = exp(j 2 pi fm[N] / fs), amplitude ratio of
        // sample L-1-n; the index pairs (n, L-2-n) share their four track loads
        cs_.zero();
        double rs, rc;
        sincos((kTwoPi * fmid) / fs, &rs, &rc);
        for (int n = lane; n < N; n += 32) {
            const int n2 = L - 2 - n;
            const double c1 = zck[n], s1 = zsk[n], c2 = zck[n2], s2 = zsk[n2];
            const double cs1 = c1 * zcc + s1 * zsc, sn1 = s1 * zcc - c1 * zsc;   // E1[n]
            const double cs2 = c2 * zcc + s2 * zsc, sn2 = s2 * zcc - c2 * zsc;   // E1[n2]
            {
                const double er = cs2 * rc - sn2 * rs, ei = cs2 * rs + sn2 * rc;
                const double r = (kAmEps + amk[L - 1 - n]) * ainv, wn = w[n];
                const double xr = (r * er) * wn, xi = (r * ei) * wn;
                nr[n] = xr; ni[n] = xi;
                cs_.add(xr, xi, (double)(n - N), wn, ws[n]);
            }
            {
                const double er = cs1 * rc - sn1 * rs, ei = cs1 * rs + sn1 * rc;
                const double r = (kAmEps + amk[n + 1]) * ainv, wn = w[n2];
                const double xr = (r * er) * wn, xi = (r * ei) * wn;
                nr[n2] = xr; ni[n2] = xi;
                cs_.add(xr, xi, (double)(n2 - N), wn, ws[n2]);
            }
        }
        if (lane == 0) {
            double sn, cs;
            sincos((kTwoPi * (0.0 - s_lo)) / fs, &sn, &cs);
            const double r = (kAmEps + amk[0]) * ainv, wn = w[L - 1];
            const double xr = (r * cs) * wn, xi = (r * sn) * wn;
            nr[L - 1] = xr; ni[L - 1] = xi;
            cs_.add(xr, xi, (double)N, wn, ws[L - 1]);
        }
        cs_.reduce();
        if (lane == 0) cs_.store(nz + it, Mx, Rb, Rd0, Rd1);
    }
}

// explicit-column basis for the eaqhmLS_complexamps seam (functions.py:472-519): am, fm [L][M] row-major;
// every column (the DC one included) is a generic column, no separate DC rows (P = 2M)
__global__ void k_basis_explicit(const double* s, const double* am, const double* fm, const double* win,
                                 int L, int M, double fs, double* X, int Lp, double2* Rb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw) + (size_t)(threadIdx.x >> 5) * Lp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int mid = (L - 1) / 2;
    double* Xr = X;
    double* Xi = X + (size_t)M * Lp;
    for (int c = warp; c < M; c += nw) {
        double run = 0.0;
        for (int base = 0; base < L; base += 32) {
            const int n = base + lane;
            double v = n < L ? fm[(size_t)n * M + c] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            v += run;
            if (n < L) S[n] = v;
            run = __shfl_sync(0xffffffffu, v, 31);
        }
        __syncwarp();
        const double Smid = S[mid], aden = am[(size_t)mid * M + c] + kAmEps;
        ColSums cs_;
        cs_.zero();
        for (int n = lane; n < L; n += 32) {
            const double nc = (double)(n - mid), wn = win[n], v = wn * s[n];
            double sn, cs;
            sincos((kTwoPi * (S[n] - Smid)) / fs, &sn, &cs);
            const double r = (kAmEps + am[(size_t)n * M + c]) / aden;
            const double xr = (r * cs) * wn, xi = (r * sn) * wn;
            Xr[(size_t)c * Lp + n] = xr; Xi[(size_t)c * Lp + n] = xi;
            cs_.add(xr, xi, nc, wn, v);
        }
        cs_.reduce();
        if (lane == 0) cs_.store(c, M, Rb, nullptr, nullptr);
        __syncwarp();
    }
}

// explicit-frequency basis for the iqhmLS_complexamps seam (functions.py:420-464)
__global__ void k_basis_harmonic_explicit(const double* s, const double* f0range, const double* win, int L, int M,
                                          double fs, double* X, int Lp, double2* Rb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double midlen = (double)(L - 1) / 2.0;
    double* Xr = X;
    double* Xi = X + (size_t)M * Lp;
    for (int c = warp; c < M; c += nw) {
        const double fk = f0range[c];
        ColSums cs_;
        cs_.zero();
        for (int n = lane; n < L; n += 32) {
            const double nc = (double)n - midlen, wn = win[n], v = wn * s[n];
            double sn, cs;
            sincos((((nc * 2.0) * kPi) * fk) / fs, &sn, &cs);
            const double xr = cs * wn, xi = sn * wn;
            Xr[(size_t)c * Lp + n] = xr; Xi[(size_t)c * Lp + n] = xi;
            cs_.add(xr, xi, nc, wn, v);
        }
        cs_.reduce();
        if (lane == 0) cs_.store(c, M, Rb, nullptr, nullptr);
    }
}

// ---------------------------------------------------------------------------------------
// Gram (functions.py:459,524): R = E_w^H E_w with E_w = [X, n*X] has the block form
// [[A, B],[B^H, C]], A = X^H X, B = X^H (n X), C = X^H (n^2 X), all three Hermitian (the weights
// w^2, n w^2, n^2 w^2 are real), so three Hermitian half-Grams of the Mx generic columns replace one of
// 2Mx columns (12 L Mx^2 instead of 16 L Mx^2 real flops).
//
// FP64 tensor path: mma.sync.m8n8k4.f64 (DMMA; shares the FP64 pipe with DFMA at the same peak but needs
// 8x fewer issue slots and operand reads).  CTA = 4 warps = 32x32 complex outputs x 3 weights; each warp a
// 16x16 tile as 2x2 DMMA sub-tiles; operands staged through shared memory by cp.async, double buffered,
// [col][n] with pitch 36 doubles (conflict-free for the fragment pattern).  Complex product by real parts:
// Gr = Xr^T Xr' + Xi^T Xi', Gi = Xr^T Xi' - Xi^T Xr'; the weights n, n^2 scale the B fragments in registers.
// ---------------------------------------------------------------------------------------
constexpr int kGT = 32;      // CTA tile edge (columns)
constexpr int kGK = 32;      // samples per pipeline stage
constexpr int kGP = 36;      // shared-memory pitch in doubles (== 4 mod 16)
constexpr int kGramSmem = 2 /*stage*/ * 2 /*panel*/ * 2 /*plane*/ * kGT * kGP * 8;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(128, 3) k_gram(SolveArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int jl = blockIdx.y;
    const int Mx = A.Mcols[jl], L = A.Lrows[jl], N = (L - 1) / 2;
    const int tp = blockIdx.x;
    int ti = (int)((sqrt(8.0 * tp + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > tp) ti--;
    while ((ti + 1) * (ti + 2) / 2 <= tp) ti++;
    const int tj = tp - ti * (ti + 1) / 2;
    if (ti * kGT >= Mx) return;
    const bool diag = ti == tj;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 1, wj = warp & 1;
    const int lr = lane >> 2, lc = lane & 3;

    // sm[stage][panel][plane][col][kGP]
    double* sm = reinterpret_cast<double*>(smem_raw);
    auto sm_at = [&](int stage, int panel, int plane) {
        return sm + ((size_t)((stage * 2 + panel) * 2 + plane)) * kGT * kGP;
    };
    const double* Xf = A.X + (size_t)jl * 2 * A.Mxmax * A.Lp;
    const size_t plane_stride = (size_t)A.Mxmax * A.Lp;

    auto load_stage = [&](int chunk, int stage) {
        const int n0 = chunk * kGK;
        // per panel and plane: 32 cols x 32 samples = 512 16-byte pieces; 128 threads x 4
        const int npanel = diag ? 1 : 2;
        for (int pnl = 0; pnl < npanel; pnl++) {
            const int c0 = (pnl == 0 ? ti : tj) * kGT;
#pragma unroll
            for (int plane = 0; plane < 2; plane++) {
                double* dst = sm_at(stage, pnl, plane);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int idx = tid + e * 128;          // 0..511
                    const int col = idx >> 4, pc = idx & 15; // piece pc covers samples n0 + 2pc, +1
                    const int n = n0 + 2 * pc;
                    const int gc = c0 + col;
                    int bytes = 0;
                    if (gc < Mx) bytes = (n + 1 < L) ? 16 : ((n < L) ? 8 : 0);
                    const int gcs = gc < Mx ? gc : 0, ns = n < L ? n : 0;
                    cp_async16(dst + col * kGP + 2 * pc, Xf + plane * plane_stride + (size_t)gcs * A.Lp + ns, bytes);
                }
            }
        }
    };

    // sub-tile activity (warp uniform): rows/cols inside Mx, and on or below the diagonal at 8x8 granularity
    const int i0 = ti * kGT + wi * 16, j0 = tj * kGT + wj * 16;
    bool act_[2][2];
    bool any = false;
#pragma unroll
    for (int si = 0; si < 2; si++)
#pragma unroll
        for (int sj = 0; sj < 2; sj++) {
            act_[si][sj] = (i0 + si * 8 < Mx) && (j0 + sj * 8 < Mx) && (j0 + sj * 8 <= i0 + si * 8);
            any |= act_[si][sj];
        }

    double acc[2][2][3][2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int d = 0; d < 2; d++) acc[a][b][c][d][0] = acc[a][b][c][d][1] = 0.0;

    const int nchunks = (L + kGK - 1) / kGK;
    load_stage(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) {
            load_stage(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (any) {
            const int st = c & 1;
            const double* ar_s = sm_at(st, 0, 0) + (wi * 16 + lr) * kGP + lc;
            const double* ai_s = sm_at(st, 0, 1) + (wi * 16 + lr) * kGP + lc;
            const double* br_s = sm_at(st, diag ? 0 : 1, 0) + (wj * 16 + lr) * kGP + lc;
            const double* bi_s = sm_at(st, diag ? 0 : 1, 1) + (wj * 16 + lr) * kGP + lc;
#pragma unroll
            for (int kk = 0; kk < kGK; kk += 4) {
                const double nc = (double)(c * kGK + kk + lc - N), nc2 = nc * nc;
                double ar[2], ai[2], nai[2], br[2][3], bi[2][3];
#pragma unroll
                for (int si = 0; si < 2; si++) {
                    ar[si] = ar_s[si * 8 * kGP + kk];
                    ai[si] = ai_s[si * 8 * kGP + kk];
                    nai[si] = -ai[si];
                }
#pragma unroll
                for (int sj = 0; sj < 2; sj++) {
                    const double r = br_s[sj * 8 * kGP + kk], i = bi_s[sj * 8 * kGP + kk];
                    br[sj][0] = r; br[sj][1] = r * nc; br[sj][2] = r * nc2;
                    bi[sj][0] = i; bi[sj][1] = i * nc; bi[sj][2] = i * nc2;
                }
#pragma unroll
                for (int si = 0; si < 2; si++)
#pragma unroll
                    for (int sj = 0; sj < 2; sj++)
                        if (act_[si][sj]) {
#pragma unroll
                            for (int wt = 0; wt < 3; wt++) {
                                dmma884(acc[si][sj][wt][0], ar[si], br[sj][wt]);
                                dmma884(acc[si][sj][wt][1], ar[si], bi[sj][wt]);
                                dmma884(acc[si][sj][wt][0], ai[si], bi[sj][wt]);
                                dmma884(acc[si][sj][wt][1], nai[si], br[sj][wt]);
                            }
                        }
            }
        }
        __syncthreads();
    }

    // epilogue: C fragment element (row lr, cols 2*lc, 2*lc+1) of each 8x8 sub-tile
    double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    const int Pp = A.Pp;
#pragma unroll
    for (int si = 0; si < 2; si++)
#pragma unroll
        for (int sj = 0; sj < 2; sj++) {
            if (!act_[si][sj]) continue;
            const int i = i0 + si * 8 + lr;
            const bool offdiag_sub = (j0 + sj * 8) < (i0 + si * 8);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = j0 + sj * 8 + 2 * lc + e;
                if (i >= Mx || j >= Mx) continue;
                const double2 a = make_double2(acc[si][sj][0][0][e], acc[si][sj][0][1][e]);
                const double2 b = make_double2(acc[si][sj][1][0][e], acc[si][sj][1][1][e]);
                const double2 cc = make_double2(acc[si][sj][2][0][e], acc[si][sj][2][1][e]);
                if (j <= i) {
                    RF[(size_t)i * Pp + j] = a;                       // A block
                    RF[(size_t)(Mx + i) * Pp + Mx + j] = cc;          // C block
                }
                RF[(size_t)(Mx + i) * Pp + j] = b;                    // B^H block (= B, Hermitian), full
                if (offdiag_sub) RF[(size_t)(Mx + j) * Pp + i] = make_double2(b.x, -b.y);
            }
        }
}

// ---------------------------------------------------------------------------------------
// Gram, warp-autonomous variant: the same DMMA tile arithmetic as k_gram, but every warp owns a 16x16 output tile
// (x 3 weights) end to end: it streams its two 16-column panels through a private double-buffered cp.async ring and
// never meets a CTA barrier.  Work items (frame, lower-triangular 16x16 tile) are dealt round-robin to all warps of a
// persistent grid, so ragged frame sizes, partially filled edge tiles and the idle upper warp of diagonal CTA tiles
// no longer cost barrier stalls (profiles/r01_ncu_summary.md: barrier was the top stall of k_gram).
// ---------------------------------------------------------------------------------------
#ifndef EAQHM_GRAM_NOBREAK
#define EAQHM_GRAM_NOBREAK 0
#endif
constexpr int kG2K = 16;                 // samples per stage
constexpr int kG2P = 20;                 // shared-memory pitch in doubles (== 4 mod 16: conflict-free fragments)
constexpr int kG2Warps = 4;
constexpr int kG2StageDoubles = 2 /*panel*/ * 2 /*plane*/ * 16 * kG2P;
constexpr int kGram2Smem = kG2Warps * 2 /*stages*/ * kG2StageDoubles * 8;

template <bool k3M>
__global__ void __launch_bounds__(kG2Warps * 32, 2) k_gram_warp(SolveArgs A, int items_per_frame) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * kG2StageDoubles;
    const int nitems = A.nf * items_per_frame;
    const size_t plane_stride = (size_t)A.Mxmax * A.Lp;

    // Work items (frame, lower-triangular tile) come off one global queue in frame order: the (up to 11) tiles that share a
    // 16-column panel of a frame then run at the same time on neighbouring warps and find it in L2, whatever the spread of
    // window lengths and empty tile slots does to the warps' progress (a static round-robin lets them drift apart)
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(A.counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int jl = item / items_per_frame, t = item - jl * items_per_frame;
        int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while (ti * (ti + 1) / 2 > t) ti--;
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        const int tj = t - ti * (ti + 1) / 2;
        const int Mx = A.Mcols[jl];
        const int i0 = ti * 16, j0 = tj * 16;
        if (i0 >= Mx) continue;
        const int L = A.Lrows[jl], N = (L - 1) / 2;
        const bool diag = ti == tj;
        const double* Xf = A.X + (size_t)jl * 2 * A.Mxmax * A.Lp;

        bool act_[2][2];
#pragma unroll
        for (int si = 0; si < 2; si++)
#pragma unroll
            for (int sj = 0; sj < 2; sj++)
                act_[si][sj] = (i0 + si * 8 < Mx) && (j0 + sj * 8 < Mx) && (j0 + sj * 8 <= i0 + si * 8);
        const bool all_act = act_[0][0] && act_[0][1] && act_[1][0] && act_[1][1];

        // plain: acc[..][0] = real part, [1] = imaginary part (4 real products per complex product).
        // 3M:    acc[..][0] = Xr^T Xr', [1] = Xi^T Xi', [2] = (Xr - Xi)^T (Xr' + Xi'): 3 real products;
        //        Re = [0] + [1], Im = [2] - [0] + [1]  (normwise as accurate as the 4-product form)
        constexpr int kAcc = k3M ? 3 : 2;
        double acc[2][2][3][kAcc][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int d = 0; d < kAcc; d++) acc[a][b][c][d][0] = acc[a][b][c][d][1] = 0.0;

        // stage loader: panel 0 = tile rows' columns, panel 1 = tile cols' columns (skipped on the diagonal);
        // per (panel, plane): 16 cols x 16 samples = 128 pieces of 16 B -> 4 per lane
        auto load_stage = [&](int chunk, int slot) {
            const int n0 = chunk * kG2K;
            double* dst0 = ring + (size_t)slot * kG2StageDoubles;
            const int npanel = diag ? 1 : 2;
            for (int pnl = 0; pnl < npanel; pnl++) {
                const int c0 = pnl == 0 ? i0 : j0;
#pragma unroll
                for (int plane = 0; plane < 2; plane++) {
                    double* dst = dst0 + (size_t)(pnl * 2 + plane) * 16 * kG2P;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int idx = lane + e * 32;           // 0..127
                        const int col = idx >> 3, pc = idx & 7;  // piece pc covers samples n0 + 2pc, +1
                        const int n = n0 + 2 * pc, gc = c0 + col;
                        int bytes = 0;
                        if (gc < Mx) bytes = (n + 1 < L) ? 16 : ((n < L) ? 8 : 0);
                        const int gcs = gc < Mx ? gc : 0, ns = n < L ? n : 0;
                        cp_async16(dst + col * kG2P + 2 * pc, Xf + plane * plane_stride + (size_t)gcs * A.Lp + ns, bytes);
                    }
                }
            }
        };

        const int nst = (L + kG2K - 1) / kG2K;
        load_stage(0, 0);
        cp_async_commit();
        for (int c = 0; c < nst; c++) {
            if (c + 1 < nst) {
                load_stage(c + 1, (c + 1) & 1);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const double* st = ring + (size_t)(c & 1) * kG2StageDoubles;
            const double* ar_s = st + (0 * 2 + 0) * 16 * kG2P + lr * kG2P + lc;
            const double* ai_s = st + (0 * 2 + 1) * 16 * kG2P + lr * kG2P + lc;
            const double* br_s = st + ((diag ? 0 : 1) * 2 + 0) * 16 * kG2P + lr * kG2P + lc;
            const double* bi_s = st + ((diag ? 0 : 1) * 2 + 1) * 16 * kG2P + lr * kG2P + lc;
            // one k-step (four samples): fragments from the stage, weights n, n^2 applied to the B side, 9 (3M) or 12 DMMAs
            // per active 8x8 sub-tile.  kAll: all four sub-tiles active - no predicate, no branch
            auto kstep = [&](auto all_c, int kk) {
                constexpr bool kAll = decltype(all_c)::value;
                const double nc = (double)(c * kG2K + kk + lc - N), nc2 = nc * nc;
                double ar[2], ai[2], ax[2], br[2][3], bi[2][3], bx[2][3];
#pragma unroll
                for (int si = 0; si < 2; si++) {
                    ar[si] = ar_s[si * 8 * kG2P + kk];
                    ai[si] = ai_s[si * 8 * kG2P + kk];
                    ax[si] = k3M ? ar[si] - ai[si] : -ai[si];
                }
#pragma unroll
                for (int sj = 0; sj < 2; sj++) {
                    const double r = br_s[sj * 8 * kG2P + kk], i = bi_s[sj * 8 * kG2P + kk];
                    br[sj][0] = r; br[sj][1] = r * nc; br[sj][2] = r * nc2;
                    bi[sj][0] = i; bi[sj][1] = i * nc; bi[sj][2] = i * nc2;
                    if (k3M) { const double t = r + i; bx[sj][0] = t; bx[sj][1] = t * nc; bx[sj][2] = t * nc2; }
                }
#pragma unroll
                for (int si = 0; si < 2; si++)
#pragma unroll
                    for (int sj = 0; sj < 2; sj++)
                        if (kAll || act_[si][sj]) {
#pragma unroll
                            for (int wt = 0; wt < 3; wt++) {
                                if (k3M) {
                                    dmma884(acc[si][sj][wt][0], ar[si], br[sj][wt]);
                                    dmma884(acc[si][sj][wt][1], ai[si], bi[sj][wt]);
                                    dmma884(acc[si][sj][wt][kAcc - 1], ax[si], bx[sj][wt]);
                                } else {
                                    dmma884(acc[si][sj][wt][0], ar[si], br[sj][wt]);
                                    dmma884(acc[si][sj][wt][1], ar[si], bi[sj][wt]);
                                    dmma884(acc[si][sj][wt][0], ai[si], bi[sj][wt]);
                                    dmma884(acc[si][sj][wt][1], ax[si], br[sj][wt]);
                                }
                            }
                        }
            };
#if EAQHM_GRAM_NOBREAK
            // interior tiles (all four sub-tiles active), every stage but the last: four k-steps as one straight-line block,
            // so that a k-step's fragment loads and scalings can be scheduled under the previous one's DMMAs
            if (all_act && c + 1 < nst) {
#pragma unroll
                for (int kk = 0; kk < kG2K; kk += 4) kstep(std::true_type{}, kk);
            } else
#endif
            {
#pragma unroll
                for (int kk = 0; kk < kG2K; kk += 4) {
                    if (c * kG2K + kk >= L) break;             // k-steps past the window end of the last stage
                    kstep(std::false_type{}, kk);
                }
            }
            __syncwarp();
        }

        double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
        const int Pp = A.Pp;
#pragma unroll
        for (int si = 0; si < 2; si++)
#pragma unroll
            for (int sj = 0; sj < 2; sj++) {
                if (!act_[si][sj]) continue;
                const int i = i0 + si * 8 + lr;
                const bool offdiag_sub = (j0 + sj * 8) < (i0 + si * 8);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = j0 + sj * 8 + 2 * lc + e;
                    if (i >= Mx || j >= Mx) continue;
                    auto cplx = [&](int wt) {
                        const double p0 = acc[si][sj][wt][0][e], p1 = acc[si][sj][wt][1][e];
                        return k3M ? make_double2(p0 + p1, (acc[si][sj][wt][kAcc - 1][e] - p0) + p1) : make_double2(p0, p1);
                    };
                    const double2 a = cplx(0), b = cplx(1), cc = cplx(2);
                    if (j <= i) {
                        RF[(size_t)i * Pp + j] = a;
                        RF[(size_t)(Mx + i) * Pp + Mx + j] = cc;
                    }
                    RF[(size_t)(Mx + i) * Pp + j] = b;
                    if (offdiag_sub) RF[(size_t)(Mx + j) * Pp + i] = make_double2(b.x, -b.y);
                }
            }
    }
}

// ---------------------------------------------------------------------------------------
// Adaptation 0: the basis is stationary and harmonic, E2[n,h] = exp(j 2 pi h f0 n / fs) (functions.py:453-454),
// so every Gram entry depends only on the harmonic-number difference d = h_q - h_p:
//   sum_n n^e w_n^2 exp(j d theta n),  theta = 2 pi f0 / fs,  e = 0,1,2   (Toeplitz blocks).
// With the symmetric window the three lag tables are T0(d) = w_0^2 + 2 sum_{n>0} w_n^2 cos(d theta n) (real),
// T1(d) = 2j sum_{n>0} n w_n^2 sin(d theta n) (imaginary, odd), T2(d) = 2 sum_{n>0} n^2 w_n^2 cos(d theta n) (real):
// O(L K) work instead of the O(L K^2) Gram.  One CTA per frame: thread per lag d (rotation recurrence over n), then
// the P x P lower triangle is filled from the tables.  The DC rows and the rhs row come from k_basis<false>.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gram_toeplitz(AnaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int jl = blockIdx.x;
    const int g = A.t.alist[A.j0 + jl];
    const int N = A.t.N[g], L = 2 * N + 1;
    const int Mx = A.Mcols[jl], K = Mx / 2;
    double* w2 = reinterpret_cast<double*>(smem_raw);       // [N+1]  w_n^2 for n = 0..N
    double* T0 = w2 + (A.Lp / 2 + 2);                       // [2K+1]
    double* T1 = T0 + (A.Mxmax + 1);
    double* T2 = T1 + (A.Mxmax + 1);
    for (int n = tid; n <= N; n += 256) {
        const double nn = (double)(2 * n), den = (double)(L - 1);   // numpy.blackman at index N + n
        const double wn = 0.42 + 0.5 * cos((kPi * nn) / den) + 0.08 * cos(((2.0 * kPi) * nn) / den);
        w2[n] = wn * wn;
    }
    __syncthreads();
    const double theta = (kTwoPi * A.t.f0[g]) / A.t.fs;
    for (int d = tid; d <= 2 * K; d += 256) {
        double rs, rc;
        sincos((double)d * theta, &rs, &rc);
        double zr = 1.0, zi = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int n = 1; n <= N; n++) {
            const double nzr = zr * rc - zi * rs;
            zi = zr * rs + zi * rc;
            zr = nzr;
            const double wn2 = w2[n], nw = (double)n * wn2;
            t0 = fma(wn2, zr, t0);
            t1 = fma(nw, zi, t1);
            t2 = fma((double)n * nw, zr, t2);
        }
        T0[d] = w2[0] + 2.0 * t0;
        T1[d] = 2.0 * t1;
        T2[d] = 2.0 * t2;
    }
    __syncthreads();
    double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    const int Pp = A.Pp;
    // harmonic number of generic column c: +1..+K for c < K, -1..-K for c >= K
    for (int e = tid; e < Mx * Mx; e += 256) {
        const int p = e / Mx, q = e - p * Mx;
        const int hp = p < K ? p + 1 : -(p - K + 1), hq = q < K ? q + 1 : -(q - K + 1);
        const int d = hq - hp, ad = d < 0 ? -d : d;
        if (q <= p) {
            RF[(size_t)p * Pp + q] = make_double2(T0[ad], 0.0);
            RF[(size_t)(Mx + p) * Pp + Mx + q] = make_double2(T2[ad], 0.0);
        }
        RF[(size_t)(Mx + p) * Pp + q] = make_double2(0.0, d < 0 ? -T1[ad] : T1[ad]);
    }
}

// ---------------------------------------------------------------------------------------
// Cholesky of the augmented matrix [[R, b],[b^H, .]] (left-looking, block width 16) followed by
// the back substitution L^H x = y.  Replaces inv(R) @ b (functions.py:465,530).
// One CTA of kNW warps per frame, several CTAs per SM.  Rows are owned in blocks of 16 by warp
// (row block % kNW), so a warp only ever rewrites its own rows.  Per block column jb, per owned row block:
//   1. W = R[rows, jb:jb+16] - L[rows, :jb] L[jb:jb+16, :jb]^H on the FP64 tensor path (DMMA m8n8k4), fragments
//      read straight from the factor (16 B per lane), result staged in the warp's shared-memory tile;
//   2. the warp that owns the diagonal block factors it warp-synchronously (no CTA barrier) into D;
//   3. after ONE CTA barrier ("D ready") every warp finishes its staged tiles: W_i <- W_i L_jj^{-H}, one lane per
//      row, and writes the final rows; a second barrier closes the block column.
// Row P of the augmented matrix carries conj(b); after the factorisation it holds y^H with L y = b.
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// Two-level (panel) left-looking Cholesky.  The factorisation advances in column panels of kPanelW = 48 columns.
// For panel J, k_chol_update first applies ALL earlier panels to the panel's tiles,
//     W[i, J] = R[i, J] - L[i, 0:Jb] L[Jb:Jb+48, 0:Jb]^H        (i >= Jb, in place),
// as one dependency-free batched GEMM, then k_chol_solve(panel = J) factors the panel (its own update loop only spans
// the <= 32 columns inside the panel).  80 % of the factorisation's flops move into the update kernel, which is built
// like k_gram_warp: every warp owns one 16-row x 48-column tile end to end, streams its 16 A rows and the panel's 48 B
// rows of the factor through a private double-buffered cp.async ring (8 factor columns per stage) and never meets a
// barrier; 36 DMMAs (3M: three real products per complex product) per eight 16-byte fragment loads, where the
// single-kernel version issued 16 DMMAs per four 16-byte loads straight from L2 / HBM.  Work items (frame, row tile) are
// dealt round-robin to a persistent grid.
// ---------------------------------------------------------------------------------------
constexpr int kPanelW = 48;                           // columns per panel (three 16-wide block columns)
constexpr int kPanelMinP = 512;                       // systems at least this large take the panel factorisation
constexpr int kUK = 8;                                // factor columns per pipeline stage
constexpr int kUPitch = 12;                           // row pitch of a stage in double2: 8 + 4 pad = 192 B == 64 mod 128
constexpr int kUStage = (16 + kPanelW) * kUPitch;     // double2 per stage: 16 A rows, then 48 B rows
constexpr int kUWarps = 4;
constexpr int kUpdSmem = kUWarps * 2 * kUStage * 16;

template <bool k3M>
__global__ void __launch_bounds__(kUWarps * 32, 2) k_chol_update(SolveArgs A, int panel, int items_per_frame) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    double2* ring = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * 2 * kUStage;
    const int nitems = A.nf * items_per_frame;
    const int Jb = panel * kPanelW, rt0 = Jb / kNb;
    const int Pp = A.Pp;

    // Work items (frame, row tile) come off one global queue: a static round-robin over 1,184 warps = 74 x 16 would hand
    // a warp the same row tile of every frame whenever a frame has 16 items (profiles/r02_chol_experiments.md)
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(A.counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int jl = item / items_per_frame, rt = rt0 + (item - jl * items_per_frame);
        const int P = 2 * A.Mcols[jl] + 2 * A.dc;
        const int i0 = rt * kNb;
        if (i0 > P || Jb >= P) continue;              // past the rhs row / frame already factored
        double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * Pp;

        // 8x8 sub-tiles that hold an element on or below the diagonal, inside the matrix
        bool act_[2][3][2];
#pragma unroll
        for (int si = 0; si < 2; si++)
#pragma unroll
            for (int cb = 0; cb < 3; cb++)
#pragma unroll
                for (int sj = 0; sj < 2; sj++) {
                    const int r0 = i0 + si * 8, c0 = Jb + cb * 16 + sj * 8;
                    act_[si][cb][sj] = (r0 <= P) && (c0 < P) && (c0 <= r0 + 7);
                }
        // plain: acc[..][0] = Re, [1] = Im of sum a conj(b) (4 real products).  3M: [0] = sum ar br, [1] = sum ai bi,
        // [2] = sum (ar + ai)(br - bi); Re = [0] + [1], Im = [2] - [0] + [1].
        constexpr int kAcc = k3M ? 3 : 2;
        double acc[2][3][2][kAcc][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int d = 0; d < kAcc; d++) acc[a][b][c][d][0] = acc[a][b][c][d][1] = 0.0;

        // stage loader: 16 A rows (this tile's rows) and 48 B rows (the panel's rows) x 8 factor columns, 16 B pieces;
        // rows past the matrix are clamped to the rhs row (their products are never stored)
        auto load_stage = [&](int chunk, int slot) {
            double2* dst = ring + (size_t)slot * kUStage;
            const int k0 = chunk * kUK;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int idx = lane + e * 32, row = idx >> 3, kc = idx & 7;
                cp_async16(dst + row * kUPitch + kc, RF + (size_t)min(i0 + row, P) * Pp + k0 + kc, 16);
            }
#pragma unroll
            for (int e = 0; e < 12; e++) {
                const int idx = lane + e * 32, row = idx >> 3, kc = idx & 7;
                cp_async16(dst + (16 + row) * kUPitch + kc, RF + (size_t)min(Jb + row, P) * Pp + k0 + kc, 16);
            }
        };

        const int nst = Jb / kUK;
        load_stage(0, 0);
        cp_async_commit();
        for (int c = 0; c < nst; c++) {
            if (c + 1 < nst) {
                load_stage(c + 1, (c + 1) & 1);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const double2* st = ring + (size_t)(c & 1) * kUStage;
            const double2* a_s = st + lr * kUPitch + lc;
            const double2* b_s = st + (16 + lr) * kUPitch + lc;
#pragma unroll
            for (int kk = 0; kk < kUK; kk += 4) {
                double ar[2], ai[2], ax[2];
#pragma unroll
                for (int si = 0; si < 2; si++) {
                    const double2 v = a_s[si * 8 * kUPitch + kk];
                    ar[si] = v.x; ai[si] = v.y;
                    ax[si] = k3M ? v.x + v.y : -v.x;
                }
#pragma unroll
                for (int cb = 0; cb < 3; cb++) {
                    double br[2], bi[2], bx[2];
#pragma unroll
                    for (int sj = 0; sj < 2; sj++) {
                        const double2 v = b_s[(cb * 16 + sj * 8) * kUPitch + kk];
                        br[sj] = v.x; bi[sj] = v.y;
                        bx[sj] = v.x - v.y;
                    }
#pragma unroll
                    for (int si = 0; si < 2; si++)
#pragma unroll
                        for (int sj = 0; sj < 2; sj++)
                            if (act_[si][cb][sj]) {
                                if (k3M) {
                                    dmma884(acc[si][cb][sj][0], ar[si], br[sj]);
                                    dmma884(acc[si][cb][sj][1], ai[si], bi[sj]);
                                    dmma884(acc[si][cb][sj][kAcc - 1], ax[si], bx[sj]);
                                } else {
                                    // a conj(b): Re += ar br + ai bi ; Im += ai br - ar bi
                                    dmma884(acc[si][cb][sj][0], ar[si], br[sj]);
                                    dmma884(acc[si][cb][sj][0], ai[si], bi[sj]);
                                    dmma884(acc[si][cb][sj][1], ai[si], br[sj]);
                                    dmma884(acc[si][cb][sj][1], ax[si], bi[sj]);
                                }
                            }
                }
            }
            __syncwarp();
        }

        // epilogue: W = R - acc in place; C fragment element (row lr, cols 2 lc, 2 lc + 1) of each 8x8 sub-tile
#pragma unroll
        for (int si = 0; si < 2; si++)
#pragma unroll
            for (int cb = 0; cb < 3; cb++)
#pragma unroll
                for (int sj = 0; sj < 2; sj++) {
                    if (!act_[si][cb][sj]) continue;
                    const int i = i0 + si * 8 + lr;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int cidx = Jb + cb * 16 + sj * 8 + 2 * lc + e;
                        if (i > P || cidx >= P || cidx > i) continue;
                        const double p0 = acc[si][cb][sj][0][e], p1 = acc[si][cb][sj][1][e];
                        const double re = k3M ? p0 + p1 : p0;
                        const double im = k3M ? (acc[si][cb][sj][kAcc - 1][e] - p0) + p1 : p1;
                        double2 v = RF[(size_t)i * Pp + cidx];
                        v.x -= re; v.y -= im;
                        RF[(size_t)i * Pp + cidx] = v;
                    }
                }
    }
}

#ifdef EAQHM_CHOL_TIMING
__device__ unsigned long long g_chol_cycles[8];
#define TMARK(slot) do { if (lane == 0) { const long long _t = clock64(); tacc[slot] += _t - tlast; tlast = _t; } } while (0)
#else
#define TMARK(slot) do { } while (0)
#endif

#ifndef EAQHM_CHOL_3M
#define EAQHM_CHOL_3M 0
#endif
constexpr bool kChol3M = EAQHM_CHOL_3M != 0;
constexpr int kOneWarpMaxP = 128; // launches whose largest system is at most this size use one warp per frame (eaqhm_launch_solve)

#ifndef EAQHM_CHOL_ROWSPLIT
#define EAQHM_CHOL_ROWSPLIT 0
#endif
#ifndef EAQHM_CHOL_PREFETCH
#define EAQHM_CHOL_PREFETCH 0
#endif
#ifndef EAQHM_CHOL_THREADS_PER_SM
#define EAQHM_CHOL_THREADS_PER_SM 512
#endif
template <int kNW, bool kQueue = false>
__global__ void __launch_bounds__(kNW * 32, EAQHM_CHOL_THREADS_PER_SM / (kNW * 32)) k_chol_solve(SolveArgs A, int panel) {
    // panel < 0: the whole factorisation and the back substitution in one launch.  panel = J >= 0: only the block columns
    // of column panel J (kPanelW columns), whose tiles k_chol_update has already updated with every earlier panel, so the
    // update loop below starts at the panel's first column; the back substitution runs in the frame's last panel.
    // kQueue (panel < 0 only): persistent grid, CTAs take frames off the queue at A.squeue (the chunk pipeline sizes the
    // grid to the share of an SM it leaves to this kernel; a second launch may join the same queue).
    constexpr int kThreads = kNW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int Pp = A.Pp;
    __shared__ double2 D[kNb][kNb + 1];               // factored diagonal block of the current block column
    __shared__ double2 Tall[kNW][kNb][kNb + 1];       // per-warp staging tile
    double2 (*T)[kNb + 1] = Tall[warp];
    // row groups of the back substitution's dot products: at least four, so that the one-warp and the two-warp instance
    // add the same partial sums in the same order (a frame's result must not depend on which instance its launch got)
    constexpr int kGroups = kThreads / 16 < 4 ? 4 : kThreads / 16;
    __shared__ double2 red[kGroups][17];
    __shared__ double Dorig[kNW][kNb];                // diagonal of R before the update: scale of the pivot test
    __shared__ int bad;
    __shared__ int next_frame;
    extern __shared__ __align__(16) double2 xs[];     // [Pmax] solution
  for (;;) {
    int jl = blockIdx.x;
    if (kQueue) {
        if (tid == 0) next_frame = atomicAdd(A.squeue, 1);
        __syncthreads();
        jl = next_frame;
        if (jl >= A.nf) return;
    }
    const int P = 2 * A.Mcols[jl] + 2 * A.dc;
    double2* RF = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    if (tid == 0) bad = 0;
    __syncthreads();
#ifdef EAQHM_CHOL_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif

    const int nrb = (P + 1 + kNb - 1) / kNb;          // row blocks (the rhs row P is in the last one)
    const int ncb = (P + kNb - 1) / kNb;              // block columns
    const int jbk_lo = panel < 0 ? 0 : panel * (kPanelW / kNb);
    const int jbk_hi = panel < 0 ? ncb : min(ncb, jbk_lo + kPanelW / kNb);
    if (jbk_lo >= ncb) return;                        // this frame's factorisation ended in an earlier panel
    const int kbeg = jbk_lo * kNb;
    for (int jbk = jbk_lo; jbk < jbk_hi; jbk++) {
        const int jb = jbk * kNb, nbc = min(kNb, P - jb);
        int rb = jbk + ((warp - jbk) % kNW + kNW) % kNW;     // first owned row block >= jbk
        bool d_ready = false;
        for (; rb < nrb || !d_ready; rb += kNW) {
            const bool have = rb < nrb;
            const int i0 = rb * kNb;
            if (have) {
                // ---- 1. DMMA update of the 16x16 tile ---------------------------------------------
                // plain: wr, wi accumulate the real and imaginary parts (4 DMMA per complex 8x8x4 product).
                // 3M (kChol3M): wr = -sum ar br, wi = -sum ai bi, wx = -sum (ar + ai)(br - bi) (3 DMMA); afterwards
                // Re = wr + wi, Im = wx - wr + wi.
                double wr[2][2][2], wi[2][2][2], wx[2][2][2];
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
                        wr[a][b][0] = wr[a][b][1] = wi[a][b][0] = wi[a][b][1] = wx[a][b][0] = wx[a][b][1] = 0.0;
                if (jb > kbeg) {
                    const double2* arow[2];
                    const double2* brow[2];
#pragma unroll
                    for (int s2 = 0; s2 < 2; s2++) {
                        arow[s2] = RF + (size_t)min(i0 + s2 * 8 + lr, P) * Pp + lc;
                        brow[s2] = RF + (size_t)min(jb + s2 * 8 + lr, P) * Pp + lc;
                    }
                    // Software pipeline in registers: jb is a multiple of 16, i.e. of four k-steps.  Two operand sets of two
                    // k-steps each; the loads of one set are issued before the 32 DMMAs of the other, so a warp always has
                    // eight 16-byte loads in flight while it computes (the first DMMA after a load group was the top stall).
                    auto kstep = [&](const double2 (&a)[2], const double2 (&b)[2]) {
                        if (kChol3M) {
                            double as[2], bd[2];
#pragma unroll
                            for (int s2 = 0; s2 < 2; s2++) { as[s2] = -(a[s2].x + a[s2].y); bd[s2] = b[s2].x - b[s2].y; }
#pragma unroll
                            for (int si = 0; si < 2; si++)
#pragma unroll
                                for (int sj = 0; sj < 2; sj++) {
                                    dmma884(wr[si][sj], -a[si].x, b[sj].x);
                                    dmma884(wi[si][sj], -a[si].y, b[sj].y);
                                    dmma884(wx[si][sj], as[si], bd[sj]);
                                }
                        } else {
#pragma unroll
                            for (int si = 0; si < 2; si++)
#pragma unroll
                                for (int sj = 0; sj < 2; sj++) {
                                    // W -= a * conj(b):  Wr -= ar br + ai bi ; Wi -= ai br - ar bi
                                    dmma884(wr[si][sj], -a[si].x, b[sj].x);
                                    dmma884(wi[si][sj], -a[si].y, b[sj].x);
                                    dmma884(wr[si][sj], -a[si].y, b[sj].y);
                                    dmma884(wi[si][sj], a[si].x, b[sj].y);
                                }
                        }
                    };
                    double2 pa[2][2], pb[2][2], qa[2][2], qb[2][2];      // [k-step within the set][row half]
                    auto load2 = [&](double2 (&sa)[2][2], double2 (&sb)[2][2], int k0) {
#pragma unroll
                        for (int q = 0; q < 2; q++)
#pragma unroll
                            for (int s2 = 0; s2 < 2; s2++) { sa[q][s2] = arow[s2][k0 + 4 * q]; sb[q][s2] = brow[s2][k0 + 4 * q]; }
                    };
                    load2(pa, pb, kbeg);
                    for (int k0 = kbeg; k0 < jb; k0 += 16) {
#if EAQHM_CHOL_PREFETCH > 0
                        {
                            // L2 prefetch of the A-operand stream kChPf iterations (16 columns each) ahead: 16 rows x 256 B =
                            // 32 lines, one per lane; past the end of this tile's k range it runs into the warp's next tile
                            // of the same block column (same k range, next owned row block)
                            int kp = k0 + EAQHM_CHOL_PREFETCH * 16, ip = i0;
                            if (kp >= jb) { kp = kbeg + (kp - jb); ip = i0 + kNW * kNb; }
                            if (kp < jb && ip <= P) {
                                const double2* pp = RF + (size_t)min(ip + (lane >> 1), P) * Pp + kp + (lane & 1) * 8;
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp));
                            }
                        }
#endif
                        load2(qa, qb, k0 + 8);
                        kstep(pa[0], pb[0]); kstep(pa[1], pb[1]);
                        if (k0 + 16 < jb) load2(pa, pb, k0 + 16);
                        kstep(qa[0], qb[0]); kstep(qa[1], qb[1]);
                    }
                    if (kChol3M) {
#pragma unroll
                        for (int si = 0; si < 2; si++)
#pragma unroll
                            for (int sj = 0; sj < 2; sj++)
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    const double p1 = wr[si][sj][e], p2 = wi[si][sj][e];
                                    wr[si][sj][e] = p1 + p2;
                                    wi[si][sj][e] = (wx[si][sj][e] - p1) + p2;
                                }
                    }
                }
                __syncwarp();
                TMARK(0);   // k loop
#pragma unroll
                for (int si = 0; si < 2; si++)
#pragma unroll
                    for (int sj = 0; sj < 2; sj++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = si * 8 + lr, cidx = sj * 8 + 2 * lc + e, i = i0 + r;
                            double2 v = make_double2(0.0, 0.0);
                            if (i <= P && cidx < nbc && jb + cidx <= i) {
                                v = RF[(size_t)i * Pp + jb + cidx];
                                if (rb == jbk && r == cidx) Dorig[warp][r] = v.x;
                                v.x += wr[si][sj][e]; v.y += wi[si][sj][e];
                            }
                            T[r][cidx] = v;
                        }
                __syncwarp();
                TMARK(1);   // staging
                // ---- 2. diagonal tile: factor it in place (warp-synchronous), publish as D ----------
                if (rb == jbk) {
                    for (int c = 0; c < nbc; c++) {
                        double dd = T[c][c].x;
                        if (!(dd > kPivotTol * Dorig[warp][c])) {   // not (numerically) positive definite: flag the frame (the host raises) and keep
                            if (lane == 0) bad = 1;     // the factor finite so that nothing downstream sees NaN
                            dd = 1.0;
                        }
                        const double dsq = sqrt(dd), inv = 1.0 / dsq;
                        __syncwarp();
                        if (lane == c) T[c][c] = make_double2(dsq, 0.0);
                        else if (lane > c && lane < nbc) { T[lane][c].x *= inv; T[lane][c].y *= inv; }
                        __syncwarp();
                        for (int e = lane; e < kNb * kNb; e += 32) {
                            const int r = e >> 4, c2 = e & 15;
                            if (r > c && r < nbc && c2 > c && c2 <= r) cfms_conj_b(T[r][c2], T[r][c], T[c2][c]);
                        }
                        __syncwarp();
                    }
                    for (int e = lane; e < kNb * kNb; e += 32) {
                        const int r = e >> 4, c2 = e & 15;
                        const double2 v = (r < nbc && c2 <= r) ? T[r][c2] : make_double2(0.0, 0.0);
                        D[r][c2] = v;
                        if (r < nbc && c2 <= r) RF[(size_t)(jb + r) * Pp + jb + c2] = v;
                    }
                }
            }
            TMARK(2);       // factor
            if (!d_ready) {
                __syncthreads();                      // D ready (every warp passes here exactly once per block column)
                d_ready = true;
            }
            TMARK(3);       // wait D
            // ---- 3. rows below the diagonal block: W_i <- W_i L_jj^{-H}, one lane per row ------------
#if EAQHM_CHOL_ROWSPLIT
            // two lanes per row: lane r solves the real parts, lane r + 16 the imaginary parts of row r, exchanging each
            // finished entry by one shuffle; the same fused multiply-adds in the same order as the one-lane form below
            // (bit-identical), on all 32 lanes instead of 16
            if (have) {
                const int row = lane & 15, part = lane >> 4, i = i0 + row;
                const bool valid = i <= P && !(rb == jbk && row < nbc);
                double own[kNb], oth[kNb];
#pragma unroll
                for (int c = 0; c < kNb; c++) {
                    const double2 t = T[row][c];
                    own[c] = part ? t.y : t.x;
                    oth[c] = part ? t.x : t.y;
                }
#pragma unroll
                for (int c = 0; c < kNb; c++) {
                    if (c < nbc) {
                        double v = own[c];
#pragma unroll
                        for (int cc = 0; cc < kNb; cc++)
                            if (cc < c) {
                                const double2 d = D[c][cc];
                                v = fma(-own[cc], d.x, v);
                                v = fma(part ? oth[cc] : -oth[cc], d.y, v);
                            }
                        own[c] = v * (1.0 / D[c][c].x);
                        oth[c] = __shfl_xor_sync(0xffffffffu, own[c], 16);
                    }
                }
                if (valid) {
                    double2* rowp = RF + (size_t)i * Pp + jb;
#pragma unroll
                    for (int c = 0; c < kNb; c++)
                        if (c < nbc && (c >> 3) == part) rowp[c] = part ? make_double2(oth[c], own[c]) : make_double2(own[c], oth[c]);
                }
            }
#else
            if (have && lane < kNb) {
                const int i = i0 + lane;
                // rows of the diagonal tile that lie below the (possibly partial) diagonal block - the rhs row P in
                // the last block column - are ordinary rows
                if (i <= P && !(rb == jbk && lane < nbc)) {
                    double2 W[kNb];
#pragma unroll
                    for (int c = 0; c < kNb; c++) W[c] = T[lane][c];
#pragma unroll
                    for (int c = 0; c < kNb; c++) {
                        if (c < nbc) {
                            double2 v = W[c];
#pragma unroll
                            for (int cc = 0; cc < kNb; cc++) if (cc < c) cfms_conj_b(v, W[cc], D[c][cc]);
                            const double inv = 1.0 / D[c][c].x;
                            W[c] = make_double2(v.x * inv, v.y * inv);
                        }
                    }
                    double2* rowp = RF + (size_t)i * Pp + jb;
#pragma unroll
                    for (int c = 0; c < kNb; c++) if (c < nbc) rowp[c] = W[c];
                }
            }
#endif
            __syncwarp();
            TMARK(4);       // solve rows
        }
        __syncthreads();                              // block column complete: rows final for the next one
        TMARK(5);           // wait end of block column
    }

    if (tid == 0 && bad && A.status) { atomicAdd(&A.status[0], 1); bad = 0; }
    if (jbk_hi < ncb) return;                         // more panels to come
    // back substitution: row P of the factor is y^H, so y_q = conj(R[P][q]); solve L^H x = y.
    // Per block column (last to first): the diagonal block is brought into shared memory with one coalesced load
    // (the 16 dependent steps of the triangular solve then never touch global memory), the contribution of the
    // already solved unknowns is a column-block dot product spread over all threads (independent loads, unrolled).
    for (int b = ncb - 1; b >= 0; b--) {
        const int jb = b * kNb, nbc = min(kNb, P - jb);
        const int c = tid & 15, rr = tid >> 4;
        for (int e = tid; e < kNb * kNb; e += kThreads) {
            const int r = e >> 4, c2 = e & 15;
            D[r][c2] = (r < nbc && c2 <= r) ? RF[(size_t)(jb + r) * Pp + jb + c2] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int grp = rr; grp < kGroups; grp += kThreads / 16) {
            double2 part = make_double2(0.0, 0.0);
            if (c < nbc) {
                constexpr int kStride = kGroups;
                int i = jb + nbc + grp;
                for (; i + 3 * kStride < P; i += 4 * kStride) {
                    const double2 l0 = RF[(size_t)i * Pp + jb + c], l1 = RF[(size_t)(i + kStride) * Pp + jb + c];
                    const double2 l2 = RF[(size_t)(i + 2 * kStride) * Pp + jb + c], l3 = RF[(size_t)(i + 3 * kStride) * Pp + jb + c];
                    cfma_conj_a(part, l0, xs[i]); cfma_conj_a(part, l1, xs[i + kStride]);
                    cfma_conj_a(part, l2, xs[i + 2 * kStride]); cfma_conj_a(part, l3, xs[i + 3 * kStride]);
                }
                for (; i < P; i += kStride) cfma_conj_a(part, RF[(size_t)i * Pp + jb + c], xs[i]);
            }
            red[grp][c] = part;
        }
        const double2 yq = (tid < nbc) ? RF[(size_t)P * Pp + jb + tid] : make_double2(0.0, 0.0);
        __syncthreads();
        if (tid < 32) {
            double2 v = make_double2(0.0, 0.0);
            if (tid < nbc) {
                v = make_double2(yq.x, -yq.y);
                for (int r2 = 0; r2 < kGroups; r2++) { v.x -= red[r2][tid].x; v.y -= red[r2][tid].y; }
            }
            for (int cc = nbc - 1; cc >= 0; cc--) {
                double2 xc = make_double2(0.0, 0.0);
                if (tid == cc) {
                    const double inv = 1.0 / D[cc][cc].x;
                    xc = make_double2(v.x * inv, v.y * inv);
                    xs[jb + cc] = xc;
                }
                xc.x = __shfl_sync(0xffffffffu, xc.x, cc);
                xc.y = __shfl_sync(0xffffffffu, xc.y, cc);
                if (tid < cc) {   // v -= conj(L[jb+cc][jb+tid]) * x_cc
                    const double2 l = D[cc][tid];
                    v.x -= l.x * xc.x + l.y * xc.y;
                    v.y -= l.x * xc.y - l.y * xc.x;
                }
            }
        }
        __syncthreads();
    }
    TMARK(6);               // back substitution
#ifdef EAQHM_CHOL_TIMING
    if (lane == 0) for (int q = 0; q < 8; q++) atomicAdd(&g_chol_cycles[q], (unsigned long long)tacc[q]);
#endif
    double2* xo = A.x + (size_t)jl * A.Pmax;
    for (int p = tid; p < P; p += kThreads) xo[p] = xs[p];
    if (!kQueue) return;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Adaptation 0 in real arithmetic.  The basis of iqhmLS_complexamps (functions.py:453-455) is exp(j h theta n) and
// n exp(j h theta n), h = -K..K, under a symmetric window on the symmetric sample range n = -N..N, and the signal is real.
// The minimiser is conjugate symmetric (a_-h = conj a_h), so the model is
//     sum_k alpha_k cos(k theta n) + gamma_k sin(k theta n) + beta_k n cos(k theta n) + delta_k n sin(k theta n)
// with a_k = (alpha_k - j gamma_k) / 2, b_k = (beta_k - j delta_k) / 2 (k >= 1), a_0 = alpha_0, b_0 = beta_0, and the even
// functions {cos, n sin} are orthogonal to the odd ones {sin, n cos} under the weight w^2: the 2(2K+1) complex normal
// equations split into TWO real symmetric positive definite systems of 2K+1 unknowns whose entries are sums and
// differences of the same three lag tables as the Toeplitz fill:  16 x fewer flops and 4 x fewer bytes than the complex
// Cholesky of the same least-squares problem.
//   even: [alpha_0..alpha_K | delta_1..delta_K]    odd: [gamma_1..gamma_K | beta_0..beta_K]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_a0_build_real(AnaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int jl = blockIdx.x;
    const int g = A.t.alist[A.j0 + jl];
    const int N = A.t.N[g], L = 2 * N + 1;
    const int Mx = A.Mcols[jl], K = Mx / 2;
    double* w2 = reinterpret_cast<double*>(smem_raw);       // [N+1]  w_n^2 for n = 0..N
    double* T0 = w2 + (A.Lp / 2 + 2);                       // [2K+1]  sum w^2 cos(d theta n)
    double* T1 = T0 + (A.Mxmax + 1);                        //         sum n w^2 sin(d theta n)
    double* T2 = T1 + (A.Mxmax + 1);                        //         sum n^2 w^2 cos(d theta n)
    for (int n = tid; n <= N; n += 256) {
        const double nn = (double)(2 * n), den = (double)(L - 1);   // numpy.blackman at index N + n
        const double wn = 0.42 + 0.5 * cos((kPi * nn) / den) + 0.08 * cos(((2.0 * kPi) * nn) / den);
        w2[n] = wn * wn;
    }
    __syncthreads();
    const double theta = (kTwoPi * A.t.f0[g]) / A.t.fs;
    for (int d = tid; d <= 2 * K; d += 256) {
        double rs, rc;
        sincos((double)d * theta, &rs, &rc);
        double zr = 1.0, zi = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int n = 1; n <= N; n++) {
            const double nzr = zr * rc - zi * rs;
            zi = zr * rs + zi * rc;
            zr = nzr;
            const double wn2 = w2[n], nw = (double)n * wn2;
            t0 = fma(wn2, zr, t0);
            t1 = fma(nw, zi, t1);
            t2 = fma((double)n * nw, zr, t2);
        }
        T0[d] = w2[0] + 2.0 * t0;
        T1[d] = 2.0 * t1;
        T2[d] = 2.0 * t2;
    }
    __syncthreads();
    double2* RFc = A.R + (size_t)jl * (A.Pmax + 1) * A.Pp;
    const A0Real ar = a0_real_layout(RFc, Mx);
    const int M = ar.M, Mp = ar.Mp;
    auto S1 = [&](int d) { return d < 0 ? -T1[-d] : T1[d]; };
    for (int e = tid; e < M * M; e += 256) {
        const int p = e / M, q = e - p * M;
        if (q > p) continue;
        {   // even: index <= K is cos(k), k = index; index > K is n sin(k), k = index - K
            const bool pc = p <= K, qc = q <= K;
            const int kp = pc ? p : p - K, kq = qc ? q : q - K;
            const int dm = kp > kq ? kp - kq : kq - kp;
            double v;
            if (pc && qc) v = 0.5 * (T0[dm] + T0[kp + kq]);
            else if (!pc && !qc) v = 0.5 * (T2[dm] - T2[kp + kq]);
            else v = 0.5 * (S1(kp + kq) + S1(kp - kq));            // p = n sin(kp), q = cos(kq)  (q <= p: q is the cosine)
            ar.sys[0][(size_t)p * Mp + q] = v;
        }
        {   // odd: index < K is sin(k), k = index + 1; index >= K is n cos(k), k = index - K
            const bool ps = p < K, qs = q < K;
            const int kp = ps ? p + 1 : p - K, kq = qs ? q + 1 : q - K;
            const int dm = kp > kq ? kp - kq : kq - kp;
            double v;
            if (ps && qs) v = 0.5 * (T0[dm] - T0[kp + kq]);
            else if (!ps && !qs) v = 0.5 * (T2[dm] + T2[kp + kq]);
            else v = 0.5 * (S1(kq + kp) + S1(kq - kp));            // p = n cos(kp), q = sin(kq)
            ar.sys[1][(size_t)p * Mp + q] = v;
        }
    }
    // right-hand sides from the sums k_basis<false> left in row P (and the DC entries) of the complex layout:
    // Rb[c] = (sum w^2 s cos, sum w^2 s sin), Rb[Mx + c] = the same with the extra factor n, column c <-> harmonic c + 1
    const int P = 2 * Mx + 2;
    const double2* Rb = RFc + (size_t)P * A.Pp;
    double* be = ar.sys[0] + (size_t)M * Mp;
    double* bo = ar.sys[1] + (size_t)M * Mp;
    for (int i = tid; i < M; i += 256) {
        be[i] = i == 0 ? Rb[2 * Mx].x : (i <= K ? Rb[i - 1].x : Rb[Mx + i - K - 1].y);
        bo[i] = i < K ? Rb[i].y : (i == K ? Rb[2 * Mx + 1].x : Rb[Mx + i - K - 1].x);
    }
}

// Real twin of k_chol_solve: left-looking blocked Cholesky (block width 16) of one augmented real system
// [[R, b],[b^T, .]] per CTA, DMMA block-column update, warp-synchronous diagonal factor, lane-per-row triangular
// solve, then L^T x = y.  Grid: 2 CTAs per frame (even, odd system).
template <int kNW>
__global__ void __launch_bounds__(kNW * 32, 512 / (kNW * 32)) k_chol_real(AnaArgs A) {
    constexpr int kThreads = kNW * 32;
    const int jl = blockIdx.x >> 1, par = blockIdx.x & 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const A0Real ar = a0_real_layout(A.R + (size_t)jl * (A.Pmax + 1) * A.Pp, A.Mcols[jl]);
    const int P = ar.M, Pp = ar.Mp;
    double* RF = ar.sys[par];
    __shared__ double D[kNb][kNb + 1];
    __shared__ double Tall[kNW][kNb][kNb + 1];
    __shared__ double red[kThreads / 16][17];
    __shared__ double Dorig[kNW][kNb];
    __shared__ int bad;
    extern __shared__ __align__(16) double xsr[];     // [P] solution
    double (*T)[kNb + 1] = Tall[warp];
    if (tid == 0) bad = 0;
    __syncthreads();

    const int nrb = (P + 1 + kNb - 1) / kNb;          // row blocks (the rhs row P is in the last one)
    const int ncb = (P + kNb - 1) / kNb;              // block columns
    for (int jbk = 0; jbk < ncb; jbk++) {
        const int jb = jbk * kNb, nbc = min(kNb, P - jb);
        int rb = jbk + ((warp - jbk) % kNW + kNW) % kNW;
        bool d_ready = false;
        for (; rb < nrb || !d_ready; rb += kNW) {
            const bool have = rb < nrb;
            const int i0 = rb * kNb;
            if (have) {
                double acc[2][2][2];
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
                if (jb > 0) {
                    const double* arow[2];
                    const double* brow[2];
#pragma unroll
                    for (int s2 = 0; s2 < 2; s2++) {
                        arow[s2] = RF + (size_t)min(i0 + s2 * 8 + lr, P) * Pp + lc;
                        brow[s2] = RF + (size_t)min(jb + s2 * 8 + lr, P) * Pp + lc;
                    }
#pragma unroll 4
                    for (int k0 = 0; k0 < jb; k0 += 4) {
                        double a[2], b[2];
#pragma unroll
                        for (int s2 = 0; s2 < 2; s2++) { a[s2] = -arow[s2][k0]; b[s2] = brow[s2][k0]; }
#pragma unroll
                        for (int si = 0; si < 2; si++)
#pragma unroll
                            for (int sj = 0; sj < 2; sj++) dmma884(acc[si][sj], a[si], b[sj]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int si = 0; si < 2; si++)
#pragma unroll
                    for (int sj = 0; sj < 2; sj++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = si * 8 + lr, cidx = sj * 8 + 2 * lc + e, i = i0 + r;
                            double v = 0.0;
                            if (i <= P && cidx < nbc && jb + cidx <= i) {
                                v = RF[(size_t)i * Pp + jb + cidx];
                                if (rb == jbk && r == cidx) Dorig[warp][r] = v;
                                v += acc[si][sj][e];
                            }
                            T[r][cidx] = v;
                        }
                __syncwarp();
                if (rb == jbk) {
                    for (int c = 0; c < nbc; c++) {
                        double dd = T[c][c];
                        if (!(dd > kPivotTol * Dorig[warp][c])) {
                            if (lane == 0) bad = 1;
                            dd = 1.0;
                        }
                        const double dsq = sqrt(dd), inv = 1.0 / dsq;
                        __syncwarp();
                        if (lane == c) T[c][c] = dsq;
                        else if (lane > c && lane < nbc) T[lane][c] *= inv;
                        __syncwarp();
                        for (int e = lane; e < kNb * kNb; e += 32) {
                            const int r = e >> 4, c2 = e & 15;
                            if (r > c && r < nbc && c2 > c && c2 <= r) T[r][c2] = fma(-T[r][c], T[c2][c], T[r][c2]);
                        }
                        __syncwarp();
                    }
                    for (int e = lane; e < kNb * kNb; e += 32) {
                        const int r = e >> 4, c2 = e & 15;
                        const double v = (r < nbc && c2 <= r) ? T[r][c2] : 0.0;
                        D[r][c2] = v;
                        if (r < nbc && c2 <= r) RF[(size_t)(jb + r) * Pp + jb + c2] = v;
                    }
                }
            }
            if (!d_ready) {
                __syncthreads();
                d_ready = true;
            }
            if (have && lane < kNb) {
                const int i = i0 + lane;
                if (i <= P && !(rb == jbk && lane < nbc)) {
                    double W[kNb];
#pragma unroll
                    for (int c = 0; c < kNb; c++) W[c] = T[lane][c];
#pragma unroll
                    for (int c = 0; c < kNb; c++) {
                        if (c < nbc) {
                            double v = W[c];
#pragma unroll
                            for (int cc = 0; cc < kNb; cc++) if (cc < c) v = fma(-W[cc], D[c][cc], v);
                            W[c] = v / D[c][c];
                        }
                    }
                    double* rowp = RF + (size_t)i * Pp + jb;
#pragma unroll
                    for (int c = 0; c < kNb; c++) if (c < nbc) rowp[c] = W[c];
                }
            }
            __syncwarp();
        }
        __syncthreads();
    }

    // back substitution L^T x = y, y = row P of the factor
    for (int b = ncb - 1; b >= 0; b--) {
        const int jb = b * kNb, nbc = min(kNb, P - jb);
        const int c = tid & 15, rr = tid >> 4;
        for (int e = tid; e < kNb * kNb; e += kThreads) {
            const int r = e >> 4, c2 = e & 15;
            D[r][c2] = (r < nbc && c2 <= r) ? RF[(size_t)(jb + r) * Pp + jb + c2] : 0.0;
        }
        double part = 0.0;
        if (c < nbc) {
            constexpr int kStride = kThreads / 16;
            for (int i = jb + nbc + rr; i < P; i += kStride) part = fma(RF[(size_t)i * Pp + jb + c], xsr[i], part);
        }
        red[rr][c] = part;
        const double yq = (tid < nbc) ? RF[(size_t)P * Pp + jb + tid] : 0.0;
        __syncthreads();
        if (tid < 32) {
            double v = 0.0;
            if (tid < nbc) {
                v = yq;
                for (int r2 = 0; r2 < kThreads / 16; r2++) v -= red[r2][tid];
            }
            for (int cc = nbc - 1; cc >= 0; cc--) {
                double xc = 0.0;
                if (tid == cc) {
                    xc = v / D[cc][cc];
                    xsr[jb + cc] = xc;
                }
                xc = __shfl_sync(0xffffffffu, xc, cc);
                if (tid < cc) v = fma(-D[cc][tid], xc, v);
            }
        }
        __syncthreads();
    }
    double* xo = ar.sol[par];
    for (int p = tid; p < P; p += kThreads) xo[p] = xsr[p];
    if (tid == 0 && bad && A.status) atomicAdd(&A.status[0], 1);
}

// ---------------------------------------------------------------------------------------
// frequency mismatch, acceptance test and knot write (functions.py:297-324)
// ---------------------------------------------------------------------------------------
__global__ void k_accept(AnaArgs A) {
    const int jl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (jl >= A.nf) return;
    const int g = A.t.alist[A.j0 + jl];
    const int M = A.Mcols[jl], nz = M / 2;      // M generic columns; solution: amps [0,M), slopes [M,2M), DC at 2M
    const int u = A.t.utt[g];
    const double2* x = A.x + (size_t)jl * A.Pmax;
    const int16_t* act = A.act + (int64_t)jl * A.nnz_max;
    const int64_t G = A.t.G;
    A0Real ar;
    if (A.a0_real) ar = a0_real_layout(A.R + (size_t)jl * (A.Pmax + 1) * A.Pp, M);
    // amplitude of the +f column `it`: complex solution, or (alpha_k - j gamma_k) / 2 of the real even / odd systems
    auto amp_of = [&](int it) {
        return A.a0_real ? make_double2(0.5 * ar.sol[0][it + 1], -0.5 * ar.sol[1][it]) : x[it];
    };
    double mx = 0.0;
    for (int it = lane; it < nz; it += 32) { const double2 a = amp_of(it); mx = fmax(mx, hypot(a.x, a.y)); }
    mx = warp_max(mx);
    const double log_max = 20.0 * log10(mx) - 150.0;                       // functions.py:309
    const double f0u = A.adapt == 0 ? A.t.f0[g] : A.t.f0stale[u];         // stale f0 for a >= 1 (Q1)
    const double h = f0u / (double)(A.adapt + 1);                          // functions.py:310
    const bool correct = f0u > A.t.f0min[u];                               // functions.py:321
    if (lane == 0) A.k.a0[g] = A.a0_real ? ar.sol[0][0] : x[2 * M].x;      // functions.py:303
    const int64_t c = A.t.pos[g];
    for (int it = lane; it < nz; it += 32) {
        const int k = act[it];
        const double2 a = amp_of(it);
        const double mag = hypot(a.x, a.y);
        double df = 0.0;
        if (A.adapt > 0) {
            const double2 b = x[M + it];
            df = (A.t.fs / kTwoPi) * ((a.x * b.y - a.y * b.x) / (mag * mag));   // functions.py:297
        }
        const bool ok = (20.0 * log10(mag) > log_max) && (fabs(df) < h);       // functions.py:315
        if (ok) {
            double f;
            if (A.adapt == 0) f = (double)(k + 1) * f0u;                        // functions.py:320
            else {
                const double fc = A.fb[g] ? kFallbackFm : A.fm_cur[(int64_t)k * A.t.T + c];
                f = correct ? fc + df : fc;                                     // functions.py:322-324
            }
            A.k.am[(int64_t)k * G + g] = mag;
            A.k.ph[(int64_t)k * G + g] = atan2(a.y, a.x);
            A.k.fm[(int64_t)k * G + g] = f;
        }
    }
}

size_t basis_smem_bytes(int Lp, int nwb) {
    return (size_t)Lp * 8 * 2 + (size_t)nwb * 2 * Lp * 8 + (size_t)nwb * 2 * Lp * 2;
}

}  // namespace

// Chunks of frames are pipelined through `nsets` copies of the per-chunk scratch (see eaqhm_analysis): while the Cholesky of
// chunk c runs, the basis and Gram kernels of chunk c+1 fill the other copy.  EAQHM_PIPELINE=1 keeps one copy (serial chunks).
int eaqhm_pipeline_sets() {
    static const int n = [] {
        const char* e = getenv("EAQHM_PIPELINE");
        const int v = e ? atoi(e) : kDefaultSets;
        return v < 1 ? 1 : (v > kMaxSets ? kMaxSets : v);
    }();
    return n;
}

AnaLayout eaqhm_ana_layout(int32_t G, int32_t chunk, int32_t nnz_max, int32_t N_max) {
    AnaLayout l;
    l.nnz_max = nnz_max;
    l.Mxmax = 2 * nnz_max;
    l.Pmax = 2 * l.Mxmax + 2;
    l.Pp = (int32_t)align_up(l.Pmax, kNb);
    l.Lp = (int32_t)align_up(2 * N_max + 1, 4);
    l.nsets = eaqhm_pipeline_sets();
    int64_t o = 0;
    l.off_fb = o;  o = align_up(o + G, 256);
    const int64_t set0 = o;                     // offsets below are those of set 0; set s lies s * set_stride further
    l.off_M = o;   o = align_up(o + (int64_t)chunk * 4 * 2, 256);
    l.off_act = o; o = align_up(o + (int64_t)chunk * nnz_max * 2, 256);
    l.off_slow = o; o = align_up(o + (int64_t)chunk * nnz_max, 256);
    l.off_X = o;   o = align_up(o + (int64_t)chunk * 2 * l.Mxmax * l.Lp * 8, 256);
    l.off_R = o;   o = align_up(o + (int64_t)chunk * (l.Pmax + 1) * l.Pp * 16, 256);
    l.off_x = o;   o = align_up(o + (int64_t)chunk * l.Pmax * 16, 256);
    l.off_cnt = o; o = align_up(o + 256, 256);
    l.set_stride = o - set0;
    l.total = set0 + (int64_t)l.nsets * l.set_stride;
    return l;
}

static int solve_attributes() {
    static DeviceOnce solve_once;               // function attributes are per device
    if (const int dev = solve_once.claim(); dev != -1) {
        CHECK_ARG(dev >= 0);
        CUDA_TRY(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_gram_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGram2Smem));
        CUDA_TRY(cudaFuncSetAttribute(k_gram_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGram2Smem));
        CUDA_TRY(cudaFuncSetAttribute(k_chol_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k_chol_solve<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k_chol_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_chol_update<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdSmem));
    }
    return EAQHM_OK;
}

int eaqhm_launch_gram(const SolveArgs& a, cudaStream_t st, int ctas_per_sm, bool reset_queue) {
    if (int rc = solve_attributes(); rc != EAQHM_OK) return rc;
    // experiment knob (tools only): EAQHM_GRAM_VARIANT=1 selects the CTA-tiled kernel, 3 the warp kernel with four real
    // products per complex product; default: warp kernel with three (3M)
    static int gram_variant = getenv("EAQHM_GRAM_VARIANT") ? atoi(getenv("EAQHM_GRAM_VARIANT")) : 0;
    if (gram_variant == 1) {
        if (!reset_queue) return EAQHM_OK;                      // no queue: the first launch covers every tile
        const int Tt = (a.Mxmax + kGT - 1) / kGT;
        dim3 grid(Tt * (Tt + 1) / 2, a.nf);
        LAUNCH(KC_GRAM, st, k_gram<<<grid, 128, kGramSmem, st>>>(a));
        return EAQHM_OK;
    }
    const int nsm = eaqhm_sm_count();
    const int Tt16 = (a.Mxmax + 15) / 16, ipf = Tt16 * (Tt16 + 1) / 2;
    const long long nitems = (long long)a.nf * ipf;
    CHECK_ARG(nitems < (1LL << 31) - 4096);
    if (reset_queue) CUDA_TRY(cudaMemsetAsync(a.counter, 0, sizeof(int32_t), st));
    int blocks = ctas_per_sm * nsm;                             // persistent grid
    if ((long long)blocks * kG2Warps > nitems) blocks = (int)((nitems + kG2Warps - 1) / kG2Warps);
    if (gram_variant == 3) LAUNCH(KC_GRAM, st, k_gram_warp<false><<<blocks, kG2Warps * 32, kGram2Smem, st>>>(a, ipf));
    else LAUNCH(KC_GRAM, st, k_gram_warp<true><<<blocks, kG2Warps * 32, kGram2Smem, st>>>(a, ipf));
    return EAQHM_OK;
}

int eaqhm_launch_solve(const SolveArgs& a, cudaStream_t st) {
    if (int rc = solve_attributes(); rc != EAQHM_OK) return rc;
    // experiment knob (tools only): EAQHM_CHOL_PAD_KB pads the dynamic shared memory to lower the frames per SM
    static const size_t pad = getenv("EAQHM_CHOL_PAD_KB") ? (size_t)atoi(getenv("EAQHM_CHOL_PAD_KB")) * 1024 : 0;
    const size_t xs_bytes = (size_t)a.Pmax * sizeof(double2) + pad;
    // experiment knob (tools only): EAQHM_CHOL_WARPS in {1,2,4} warps per frame
    static int warps = getenv("EAQHM_CHOL_WARPS") ? atoi(getenv("EAQHM_CHOL_WARPS")) : 2;
    CHECK_ARG(xs_bytes - pad <= 32 * 1024);
    // Two-level (panel) factorisation for large systems: measured cross-over at P ~ 510 (P = 486: single kernel 2.5 %
    // faster; P = 538: panels 6 % faster; P = 586: 20 %; P = 1,186: 43 %), profiles/r02_chol_experiments.md.  Experiment knobs (tools only):
    // EAQHM_CHOL_PANEL=0/1 forces one or the other; EAQHM_CHOL_UPD_3M=0 forms the panel update's complex products from
    // four real products
    static const int panel_env = getenv("EAQHM_CHOL_PANEL") ? atoi(getenv("EAQHM_CHOL_PANEL")) : -1;
    const bool panels = panel_env >= 0 ? panel_env != 0 : a.Pmax >= kPanelMinP;
    static const bool upd3m = !(getenv("EAQHM_CHOL_UPD_3M") && atoi(getenv("EAQHM_CHOL_UPD_3M")) == 0);
    if (panels && warps == 2 && a.Pmax > kPanelW) {
        const int np = (a.Pmax + kPanelW - 1) / kPanelW;
        const int nrbmax = (a.Pmax + 1 + kNb - 1) / kNb;
        const int nsm = eaqhm_sm_count();
        for (int J = 0; J < np; J++) {
            if (J > 0) {
                const int ipf = nrbmax - J * (kPanelW / kNb);
                const long long nitems = (long long)a.nf * ipf;
                CHECK_ARG(nitems < (1LL << 31) - 4096);
                CUDA_TRY(cudaMemsetAsync(a.counter, 0, sizeof(int32_t), st));
                int blocks = 2 * nsm;                                   // persistent: two CTAs per SM
                if ((long long)blocks * kUWarps > nitems) blocks = (int)((nitems + kUWarps - 1) / kUWarps);
                if (upd3m) LAUNCH(KC_CHOL_UPDATE, st, k_chol_update<true><<<blocks, kUWarps * 32, kUpdSmem, st>>>(a, J, ipf));
                else LAUNCH(KC_CHOL_UPDATE, st, k_chol_update<false><<<blocks, kUWarps * 32, kUpdSmem, st>>>(a, J, ipf));
            }
            LAUNCH(KC_CHOL, st, k_chol_solve<2><<<a.nf, 64, xs_bytes, st>>>(a, J));
        }
        return EAQHM_OK;
    }
    // Small systems factor faster with one warp per frame (16 frames per SM, no CTA barrier): measured -21 % at P 62-126
    // (config c5), equal at P 106-194 (SA19), +12 % at P ~ 200-346 (c2, c3) (profiles/r02_chol_experiments.md).  The
    // per-frame P is only known on the device, so the choice is per launch, by the launch's largest possible system; a
    // per-frame split (two launches whose CTAs exit when the frame belongs to the other) was measured and costs the large
    // workloads 1-2 % in empty CTAs.  EAQHM_CHOL_WARPS forces either.
    if (!getenv("EAQHM_CHOL_WARPS") && a.Pmax <= kOneWarpMaxP) {
        LAUNCH(KC_CHOL, st, k_chol_solve<1><<<a.nf, 32, xs_bytes, st>>>(a, -1));
        return EAQHM_OK;
    }
    if (warps == 1) LAUNCH(KC_CHOL, st, k_chol_solve<1><<<a.nf, 32, xs_bytes, st>>>(a, -1));
    else if (warps == 4) LAUNCH(KC_CHOL, st, k_chol_solve<4><<<a.nf, 128, xs_bytes, st>>>(a, -1));
    else if (warps == 8) LAUNCH(KC_CHOL, st, k_chol_solve<8><<<a.nf, 256, xs_bytes, st>>>(a, -1));
    else LAUNCH(KC_CHOL, st, k_chol_solve<2><<<a.nf, 64, xs_bytes, st>>>(a, -1));
    return EAQHM_OK;
}

bool eaqhm_solve_uses_panels(const SolveArgs& a) {
    static const int panel_env = getenv("EAQHM_CHOL_PANEL") ? atoi(getenv("EAQHM_CHOL_PANEL")) : -1;
    static const int warps = getenv("EAQHM_CHOL_WARPS") ? atoi(getenv("EAQHM_CHOL_WARPS")) : 2;
    const bool panels = panel_env >= 0 ? panel_env != 0 : a.Pmax >= kPanelMinP;
    return panels && warps == 2 && a.Pmax > kPanelW;
}

int eaqhm_launch_solve_queue(const SolveArgs& a, cudaStream_t st, int ctas_per_sm, bool reset_queue) {
    if (int rc = solve_attributes(); rc != EAQHM_OK) return rc;
    CHECK_ARG(a.squeue != nullptr && ctas_per_sm >= 1);
    const size_t xs_bytes = (size_t)a.Pmax * sizeof(double2);
    CHECK_ARG(xs_bytes <= 32 * 1024);
    if (reset_queue) CUDA_TRY(cudaMemsetAsync(a.squeue, 0, sizeof(int32_t), st));
    const int blocks = min(a.nf, ctas_per_sm * eaqhm_sm_count());
    LAUNCH(KC_CHOL, st, k_chol_solve<2, true><<<blocks, 64, xs_bytes, st>>>(a, -1));
    return EAQHM_OK;
}

int eaqhm_launch_gram_solve(const SolveArgs& a, cudaStream_t st, bool do_gram) {
    if (do_gram)
        if (int rc = eaqhm_launch_gram(a, st, 2, true); rc != EAQHM_OK) return rc;
    return eaqhm_launch_solve(a, st);
}

// ---- chunk pipeline: two internal streams and their events, per device ------------------------------------------
namespace {
struct Pipe {
    cudaStream_t sg = nullptr, sc = nullptr;    // build (basis, Gram) and solve (Cholesky, acceptance) streams
    cudaEvent_t fork = nullptr, sgdone = nullptr, cqueue = nullptr, built[kMaxSets] = {}, solved[kMaxSets] = {};
};
int pipe_get(Pipe** out) {
    static Pipe pipes[kMaxDevices];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CHECK_ARG(dev >= 0 && dev < kMaxDevices);
    Pipe& P = pipes[dev];
    if (P.sg == nullptr) {
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        // experiment knob (tools only): EAQHM_PIPE_PRIO 0 = equal priorities, 1 = Gram first (default), 2 = Cholesky first
        const int prio = getenv("EAQHM_PIPE_PRIO") ? atoi(getenv("EAQHM_PIPE_PRIO")) : 1;
        CUDA_TRY(cudaStreamCreateWithPriority(&P.sg, cudaStreamNonBlocking, prio == 1 ? greatest : least));
        CUDA_TRY(cudaStreamCreateWithPriority(&P.sc, cudaStreamNonBlocking, prio == 2 ? greatest : least));
        CUDA_TRY(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&P.sgdone, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&P.cqueue, cudaEventDisableTiming));
        for (int i = 0; i < kMaxSets; i++) {
            CUDA_TRY(cudaEventCreateWithFlags(&P.built[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&P.solved[i], cudaEventDisableTiming));
        }
    }
    *out = &P;
    return EAQHM_OK;
}
}  // namespace

extern "C" int64_t eaqhm_analysis_scratch_bytes(int32_t G, int32_t chunk, int32_t nnz_max, int32_t N_max) {
    if (G < 0 || chunk <= 0 || nnz_max <= 0 || N_max <= 0) return EAQHM_E_ARG;
    return eaqhm_ana_layout(G, chunk, nnz_max, N_max).total;
}

extern "C" int eaqhm_analysis_pipeline_sets(void) { return eaqhm_pipeline_sets(); }

extern "C" int eaqhm_analysis(const eaqhm_tables* tab, const double* s, const eaqhm_tracks* cur,
                              const eaqhm_knots* knots, double* best_am, int32_t adapt, int32_t nnz_max,
                              int32_t N_max, int32_t chunk, void* scratch, int64_t scratch_bytes,
                              int32_t* status, double* flop_stats, void* stream) {
    CHECK_ARG(tab && s && cur && knots && scratch && status);
    CHECK_ARG(adapt >= 0 && nnz_max > 0 && N_max > 0 && chunk > 0);
    CHECK_ARG(nnz_max <= 32767 && 2 * N_max + 1 <= 32767);
    const AnaLayout l = eaqhm_ana_layout(tab->G, chunk, nnz_max, N_max);
    CHECK_ARG(scratch_bytes >= l.total);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)scratch;
    AnaArgs A;
    A.t = *tab;
    A.s = s;
    A.am_cur = cur->am;
    A.fm_cur = cur->fmc;
    A.zc_cur = cur->zc;
    A.zs_cur = cur->zs;
    A.k = *knots;
    A.best_am = best_am;
    A.adapt = adapt;
    A.nnz_max = l.nnz_max; A.Mxmax = l.Mxmax; A.Pmax = l.Pmax; A.Pp = l.Pp; A.Lp = l.Lp;
    A.fb = base + l.off_fb;
    A.slow_only = 0;
    A.status = status;
    A.flop_stats = flop_stats;
    if (tab->nA == 0) return EAQHM_OK;
    // per-chunk scratch of set `b`
    auto bind_set = [&](int b) {
        unsigned char* sb = base + (int64_t)b * l.set_stride;
        A.Mcols = (int32_t*)(sb + l.off_M);
        A.Lrows = A.Mcols + chunk;
        A.act = (int16_t*)(sb + l.off_act);
        A.slow = sb + l.off_slow;
        A.X = (double*)(sb + l.off_X);
        A.R = (double2*)(sb + l.off_R);
        A.x = (double2*)(sb + l.off_x);
        return (int32_t*)(sb + l.off_cnt);
    };

    // experiment knob (tools only): EAQHM_BASIS_NWB = warps of a basis CTA that own a column buffer (fewer -> more CTAs per SM)
    static const int nwb_env = getenv("EAQHM_BASIS_NWB") ? atoi(getenv("EAQHM_BASIS_NWB")) : kBasisWarps;
    int nwb = nwb_env >= 1 && nwb_env <= kBasisWarps ? nwb_env : kBasisWarps;
    while (nwb > 1 && basis_smem_bytes(l.Lp, nwb) > 200 * 1024) nwb--;
    A.nwb = nwb;
    const size_t smem = basis_smem_bytes(l.Lp, nwb);
    static DeviceOnce basis_once;               // function attributes are per device
    if (const int dev = basis_once.claim(); dev != -1) {
        CHECK_ARG(dev >= 0);
        CUDA_TRY(cudaFuncSetAttribute(k_basis<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k_basis<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k_basis_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    CHECK_ARG(smem <= 200 * 1024);

    // experiment knobs (tools only): EAQHM_A0_REAL=0 keeps the complex Toeplitz + Cholesky path at adaptation 0
    static const bool use_a0_real = !(getenv("EAQHM_A0_REAL") && atoi(getenv("EAQHM_A0_REAL")) == 0);
    static const bool no_toeplitz = getenv("EAQHM_NO_TOEPLITZ") && atoi(getenv("EAQHM_NO_TOEPLITZ"));
    A.a0_real = (adapt == 0 && use_a0_real && !no_toeplitz) ? 1 : 0;
    A.store_X = (adapt == 0 && !no_toeplitz) ? 0 : 1;       // adaptation 0 builds its normal matrix from lag tables
    CUDA_TRY(cudaMemsetAsync(A.fb, 0, tab->G, st));
    if (adapt > 0) {
        A.j0 = 0; A.nf = tab->nA;
        const int nfl = tab->flist ? tab->nF : tab->nA;
        LAUNCH(KC_FALLBACK, st, k_fallback_flags<<<(nfl * 32 + 255) / 256, 256, 0, st>>>(A));
    }

    // Build half of a chunk (active set, basis planes, rhs / DC rows) on stream `sb`; A is bound to the chunk's set
    auto build_chunk = [&](cudaStream_t sb) -> int {
        LAUNCH(KC_ACTIVE, sb, k_active<<<(A.nf * 32 + 255) / 256, 256, 0, sb>>>(A));
        // adaptive basis: fast-path kernel for the columns inside one run (phasor tracks present), k_basis<true> for the rest
        static const bool two_kernels = !(getenv("EAQHM_BASIS_FAST") && atoi(getenv("EAQHM_BASIS_FAST")) == 0);
        const size_t fsm = (size_t)l.Lp * 16;
        const bool fast = adapt > 0 && two_kernels && A.zc_cur != nullptr && A.zs_cur != nullptr && fsm <= 64 * 1024;
        if (adapt == 0) {
            LAUNCH(KC_BASIS, sb, k_basis<false><<<A.nf, kBasisThreads, smem, sb>>>(A));
        } else if (fast) {
            CUDA_TRY(cudaMemsetAsync(A.slow, 0, (size_t)A.nf * l.nnz_max, sb));
            LAUNCH(KC_BASIS, sb, k_basis_fast<<<A.nf, kBasisThreads, fsm, sb>>>(A));
            A.slow_only = 1;
            LAUNCH(KC_BASIS, sb, k_basis<true><<<A.nf, kBasisThreads, smem, sb>>>(A));
            A.slow_only = 0;
        } else {
            LAUNCH(KC_BASIS, sb, k_basis<true><<<A.nf, kBasisThreads, smem, sb>>>(A));
        }
        return EAQHM_OK;
    };
    auto solve_args = [&](int32_t* counter) {
        SolveArgs S;
        S.X = A.X; S.R = A.R; S.x = A.x; S.Mcols = A.Mcols; S.Lrows = A.Lrows;
        S.nf = A.nf; S.Mxmax = l.Mxmax; S.dc = 1; S.Pmax = l.Pmax; S.Pp = l.Pp; S.Lp = l.Lp; S.status = status;
        S.counter = counter;
        S.squeue = nullptr;
        return S;
    };
    static const bool use_toeplitz = !(getenv("EAQHM_NO_TOEPLITZ") && atoi(getenv("EAQHM_NO_TOEPLITZ")));
    const bool toeplitz = adapt == 0 && use_toeplitz;
    const int nchunks = (tab->nA + chunk - 1) / chunk;

    // (the pipeline needs the queue-driven Gram and the single-kernel Cholesky: not with the panel path or the experiment variants)
    if (adapt > 0 && l.nsets >= 2 && nchunks >= 2 && !eaqhm_solve_uses_panels(solve_args(nullptr)) &&
        !getenv("EAQHM_GRAM_VARIANT") && !getenv("EAQHM_CHOL_WARPS")) {
        // Chunk pipeline on two internal streams.  Phase p runs the Cholesky of chunk p beside the basis + Gram of chunk
        // p+1: the Gram is FP64-pipe bound with few resident warps, the Cholesky load-latency bound with many small CTAs,
        // so each covers the other's stalls.  Both kernels are persistent grids on work queues, sized to share an SM
        // (one Gram CTA = 4 warps at 240 registers + four Cholesky CTAs = 8 warps at 128 registers); when one of the
        // two queues of a phase runs dry, a second launch of the other kernel (queued behind it on the same stream)
        // joins that kernel's queue and takes over the freed half of every SM.
        //   sg: build(p+1) . Gram A(p+1) . Cholesky B(p)            sc: Cholesky A(p) . Gram B(p+1) . accept(p)
        // Same kernels, same per-frame arithmetic: results are bit-identical to the serial order (EAQHM_PIPELINE=1).
        Pipe* P = nullptr;
        if (int rc = pipe_get(&P); rc != EAQHM_OK) return rc;
        static const int g_ctas = getenv("EAQHM_PIPE_GRAM_CTAS") ? atoi(getenv("EAQHM_PIPE_GRAM_CTAS")) : 1;
        static const int c_ctas = getenv("EAQHM_PIPE_CHOL_CTAS") ? atoi(getenv("EAQHM_PIPE_CHOL_CTAS")) : 4;
        static const bool helpers = !(getenv("EAQHM_PIPE_HELPERS") && atoi(getenv("EAQHM_PIPE_HELPERS")) == 0);
        CUDA_TRY(cudaEventRecord(P->fork, st));
        CUDA_TRY(cudaStreamWaitEvent(P->sg, P->fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(P->sc, P->fork, 0));
        AnaArgs Ac[kMaxSets];               // arguments of the chunk that lives in each set
        SolveArgs Sc[kMaxSets];
        auto bind_chunk = [&](int c) {
            const int b = c % l.nsets;
            A.j0 = c * chunk;
            A.nf = min(chunk, tab->nA - A.j0);
            int32_t* counter = bind_set(b);
            Ac[b] = A;
            Sc[b] = solve_args(counter);
            Sc[b].squeue = counter + 16;
            return b;
        };
        // phase p = -1 .. nchunks-1: chunk p is factored (p >= 0) while chunk p+1 is built (p+1 < nchunks)
        for (int p = -1; p < nchunks; p++) {
            const int bs = p >= 0 ? p % l.nsets : -1;                  // set being solved
            const bool build = p + 1 < nchunks;
            int bb = -1;
            if (build) {
                // the set's previous chunk (p+1-nsets) must be solved and accepted before the set is rewritten
                if (p + 1 >= l.nsets) CUDA_TRY(cudaStreamWaitEvent(P->sg, P->solved[(p + 1) % l.nsets], 0));
                bb = bind_chunk(p + 1);
                A = Ac[bb];
                if (int rc = build_chunk(P->sg); rc != EAQHM_OK) return rc;
                CUDA_TRY(cudaMemsetAsync(Sc[bb].counter, 0, sizeof(int32_t), P->sg));
                CUDA_TRY(cudaEventRecord(P->built[bb], P->sg));
            }
            if (p >= 0) {
                // sc: everything of the previous phase on sg (Gram A of chunk p) is done; Gram B of chunk p ran on sc itself
                CUDA_TRY(cudaStreamWaitEvent(P->sc, P->sgdone, 0));
                // ... and the basis kernels of chunk p+1 have had the device to themselves (beside the Cholesky's resident
                // CTAs they crawl: 96 KB of shared memory per CTA; measured 8.9 ms against 2.5 ms per chunk)
                if (build) CUDA_TRY(cudaStreamWaitEvent(P->sc, P->built[bb], 0));
                CUDA_TRY(cudaMemsetAsync(Sc[bs].squeue, 0, sizeof(int32_t), P->sc));
                CUDA_TRY(cudaEventRecord(P->cqueue, P->sc));
                if (int rc = eaqhm_launch_solve_queue(Sc[bs], P->sc, build ? c_ctas : 2 * c_ctas, false); rc != EAQHM_OK) return rc;
            }
            if (build) {
                if (int rc = eaqhm_launch_gram(Sc[bb], P->sg, p >= 0 ? g_ctas : 2 * g_ctas, false); rc != EAQHM_OK) return rc;
                if (p >= 0 && helpers) {
                    if (int rc = eaqhm_launch_gram(Sc[bb], P->sc, g_ctas, false); rc != EAQHM_OK) return rc;
                    CUDA_TRY(cudaStreamWaitEvent(P->sg, P->cqueue, 0));
                    if (int rc = eaqhm_launch_solve_queue(Sc[bs], P->sg, c_ctas, false); rc != EAQHM_OK) return rc;
                }
            }
            CUDA_TRY(cudaEventRecord(P->sgdone, P->sg));
            if (p >= 0) {
                CUDA_TRY(cudaStreamWaitEvent(P->sc, P->sgdone, 0));     // Cholesky B of chunk p (on sg) is done
                A = Ac[bs];
                LAUNCH(KC_ACCEPT, P->sc, k_accept<<<(A.nf * 32 + 255) / 256, 256, 0, P->sc>>>(A));
                CUDA_TRY(cudaEventRecord(P->solved[bs], P->sc));
            }
        }
        // sc is in order and its last phase waited for sg: joining sc joins both
        CUDA_TRY(cudaStreamWaitEvent(st, P->solved[(nchunks - 1) % l.nsets], 0));
        return EAQHM_OK;
    }

    for (int j0 = 0; j0 < tab->nA; j0 += chunk) {
        A.j0 = j0;
        A.nf = min(chunk, tab->nA - j0);
        int32_t* counter = bind_set(0);
        if (int rc = build_chunk(st); rc != EAQHM_OK) return rc;
        const SolveArgs S = solve_args(counter);
        if (A.a0_real) {
            // adaptation 0 as two real systems per frame (see k_a0_build_real)
            const size_t tsm = ((size_t)(l.Lp / 2 + 2) + 3 * (size_t)(l.Mxmax + 1)) * 8;
            CHECK_ARG(tsm <= 48 * 1024);
            LAUNCH(KC_GRAM, st, k_a0_build_real<<<A.nf, 256, tsm, st>>>(A));
            // one warp per real system (2K+1 <= 173 unknowns): 1-3 % faster than two on every configuration
            // (profiles/r02_chol_experiments.md); experiment knob (tools only): EAQHM_CHOL_REAL_WARPS=2
            static const int real_warps = getenv("EAQHM_CHOL_REAL_WARPS") ? atoi(getenv("EAQHM_CHOL_REAL_WARPS")) : 1;
            if (real_warps == 1) LAUNCH(KC_CHOL, st, k_chol_real<1><<<2 * A.nf, 32, (size_t)(l.Mxmax + 2) * 8, st>>>(A));
            else LAUNCH(KC_CHOL, st, k_chol_real<2><<<2 * A.nf, 64, (size_t)(l.Mxmax + 2) * 8, st>>>(A));
            LAUNCH(KC_ACCEPT, st, k_accept<<<(A.nf * 32 + 255) / 256, 256, 0, st>>>(A));
            continue;
        }
        if (toeplitz) {
            const size_t tsm = ((size_t)(l.Lp / 2 + 2) + 3 * (size_t)(l.Mxmax + 1)) * 8;
            CHECK_ARG(tsm <= 48 * 1024);
            LAUNCH(KC_GRAM, st, k_gram_toeplitz<<<A.nf, 256, tsm, st>>>(A));
        }
        int rc = eaqhm_launch_gram_solve(S, st, !toeplitz);
        if (rc != EAQHM_OK) return rc;
        LAUNCH(KC_ACCEPT, st, k_accept<<<(A.nf * 32 + 255) / 256, 256, 0, st>>>(A));
    }
    return EAQHM_OK;
}

// ---- helper seams (host buffers) --------------------------------------------------------
static int seam_solve(const double* s, int L, int M, const double* win, double fs, const double* am,
                      const double* fm, const double* f0range, double* amp, double* slope) {
    CHECK_ARG(s && win && amp && slope && L >= 3 && (L & 1) && M >= 1);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        eaqhm_set_error("no CUDA device: the eaQHM library has no CPU fallback");
        return EAQHM_E_NODEVICE;
    }
    const int P = 2 * M, Pp = (int)align_up(P, kNb), Lp = (int)align_up(L, 4);
    double *d_s = nullptr, *d_win = nullptr, *d_am = nullptr, *d_fm = nullptr, *d_f0 = nullptr;
    double* d_X = nullptr;
    double2 *d_R = nullptr, *d_x = nullptr;
    int32_t* d_meta = nullptr;
    int rc = EAQHM_OK;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(&d_s, L * 8)); CUDA_TRY(cudaMalloc(&d_win, L * 8));
        CUDA_TRY(cudaMemcpy(d_s, s, L * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_win, win, L * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&d_X, (size_t)2 * M * Lp * 8));
        CUDA_TRY(cudaMalloc(&d_R, (size_t)(P + 1) * Pp * 16));
        CUDA_TRY(cudaMemset(d_R, 0, (size_t)(P + 1) * Pp * 16));
        CUDA_TRY(cudaMalloc(&d_x, (size_t)P * 16));
        CUDA_TRY(cudaMalloc(&d_meta, 8 * sizeof(int32_t)));
        const int32_t meta[8] = {M, L, 0, 0, 0, 0, 0, 0};      // [0] columns, [1] samples, [2..3] status, [4] work counter
        CUDA_TRY(cudaMemcpy(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice));
        double2* Rb = d_R + (size_t)P * Pp;
        if (f0range) {
            CUDA_TRY(cudaMalloc(&d_f0, M * 8));
            CUDA_TRY(cudaMemcpy(d_f0, f0range, M * 8, cudaMemcpyHostToDevice));
            LAUNCH(KC_BASIS, 0, k_basis_harmonic_explicit<<<1, 256>>>(d_s, d_f0, d_win, L, M, fs, d_X, Lp, Rb));
        } else {
            CUDA_TRY(cudaMalloc(&d_am, (size_t)L * M * 8)); CUDA_TRY(cudaMalloc(&d_fm, (size_t)L * M * 8));
            CUDA_TRY(cudaMemcpy(d_am, am, (size_t)L * M * 8, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_fm, fm, (size_t)L * M * 8, cudaMemcpyHostToDevice));
            CHECK_ARG((size_t)8 * Lp * 8 <= 48 * 1024);
            LAUNCH(KC_BASIS, 0, k_basis_explicit<<<1, 256, (size_t)8 * Lp * 8>>>(d_s, d_am, d_fm, d_win, L, M, fs, d_X, Lp, Rb));
        }
        SolveArgs S;
        S.X = d_X; S.R = d_R; S.x = d_x; S.Mcols = d_meta; S.Lrows = d_meta + 1;
        S.nf = 1; S.Mxmax = M; S.dc = 0; S.Pmax = P; S.Pp = Pp; S.Lp = Lp; S.status = d_meta + 2;
        S.counter = d_meta + 4;
        int r2 = eaqhm_launch_gram_solve(S, 0, true);
        if (r2 != EAQHM_OK) return r2;
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(amp, d_x, (size_t)M * 16, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(slope, d_x + M, (size_t)M * 16, cudaMemcpyDeviceToHost));
        int32_t meta_out[4];
        CUDA_TRY(cudaMemcpy(meta_out, d_meta, sizeof(meta_out), cudaMemcpyDeviceToHost));
        if (meta_out[2] != 0) {
            eaqhm_set_error("normal matrix is not positive definite");
            return EAQHM_E_NOTPD;
        }
        return EAQHM_OK;
    };
    rc = body();
    cudaFree(d_s); cudaFree(d_win); cudaFree(d_am); cudaFree(d_fm); cudaFree(d_f0);
    cudaFree(d_X); cudaFree(d_R); cudaFree(d_x); cudaFree(d_meta);
    return rc;
}

extern "C" int eaqhm_iqhmLS_complexamps(const double* s, int32_t L, const double* f0range, int32_t M,
                                        const double* window, double fs, double* amp, double* slope) {
    CHECK_ARG(f0range != nullptr);
    return seam_solve(s, L, M, window, fs, nullptr, nullptr, f0range, amp, slope);
}

extern "C" int eaqhm_eaqhmLS_complexamps(const double* s, int32_t L, const double* am, const double* fm,
                                         int32_t M, const double* window, double fs, double* amp, double* slope) {
    CHECK_ARG(am != nullptr && fm != nullptr);
    return seam_solve(s, L, M, window, fs, am, fm, nullptr, amp, slope);
}

#ifdef EAQHM_CHOL_TIMING
extern "C" int eaqhm_debug_chol_cycles(double* out, int reset) {
    unsigned long long h[8];
    if (cudaMemcpyFromSymbol(h, g_chol_cycles, sizeof(h)) != cudaSuccess) return EAQHM_E_CUDA;
    for (int i = 0; i < 8; i++) out[i] = (double)h[i];
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_chol_cycles, z, sizeof(z)); }
    return EAQHM_OK;
}
#endif
