// Host-preprocessing stages of the path on the GPU: the 30 ms frame energies behind the
// speech / voiced flags (functions.py:577-642) and the SWIPE' pitch-strength matrix.
//
// SWIPE' pitch-strength matrix on the GPU (reference: SWIPE.py:14-132, the port of Camacho's estimator the reference
// ships; host-side preprocessing of eaQHMAnalysisAndSynthesis, functions.py:111).  Per power-of-two window size:
//   k_swipe_loudness  one CTA per spectrogram frame: Hann-windowed frame -> radix-2 FFT in shared memory -> one-sided
//                     PSD (SWIPE.py:66, matplotlib specgram: |X|^2, doubled except DC / Nyquist, / fs, / sum w^2) ->
//                     linear interpolation onto the ERB frequency grid, square root (SWIPE.py:71-72) -> unit norm
//                     (SWIPE.py:140-142);
//   k_swipe_strength  strength of every pitch candidate of this window size in every frame: kernel matrix (SWIPE.py:166-184,
//                     built on the host, cached per (fs, pitch range)) times normalised loudness (SWIPE.py:146);
//   k_swipe_accumulate linear interpolation of the strengths from the frame times onto the 1 ms grid (NaN outside,
//                     SWIPE.py:92-100) and weighted accumulation into the candidate x time strength matrix (SWIPE.py:102-104).
// Then k_swipe_argmax picks the strongest candidate of every 1 ms frame with numpy.argmax's NaN rule and hands the three
// strengths around it to the host, which does the parabolic refinement (SWIPE.py:106-131) exactly like the CPU path.
// FP64 throughout; no library FFT.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned brev_bits(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// x: device signal [n_x]; frame seg covers zero-padded samples [seg*hop - pad_front, +ws)
__global__ void __launch_bounds__(512) k_swipe_loudness(const double* __restrict__ x, int64_t n_x, int ws, int log2ws, int hop,
                                                        int pad_front, const double* __restrict__ win, double fs, double sumw2,
                                                        const double* __restrict__ f, const int32_t* __restrict__ lo,
                                                        const double* __restrict__ ferb, int nerb, int64_t nseg,
                                                        double* __restrict__ Ln) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* a = reinterpret_cast<double2*>(smem_raw);                     // [ws] FFT buffer
    double* psd = reinterpret_cast<double*>(a + (ws / 2 + 1));             // [ws/2+1], aliases the unused upper half of a
    double* Lb = reinterpret_cast<double*>(a + ws);                        // [nerb]
    __shared__ double red[16];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t seg = blockIdx.x;
    const int64_t base = seg * hop - pad_front;
    for (int i = tid; i < ws; i += nt) {
        const int64_t g = base + i;
        const double v = (g >= 0 && g < n_x) ? x[g] * win[i] : 0.0;
        a[brev_bits((unsigned)i, log2ws)] = make_double2(v, 0.0);
    }
    __syncthreads();
    for (int s = 1; s <= log2ws; s++) {
        const int half = 1 << (s - 1);
        for (int b = tid; b < ws / 2; b += nt) {
            const int grp = b >> (s - 1), pos = b & (half - 1);
            const int i0 = (grp << s) + pos, i1 = i0 + half;
            double sn, cs;
            sincospi(-(double)pos / (double)half, &sn, &cs);               // exp(-2 pi i pos / (2 half))
            const double2 u = a[i0], v = a[i1];
            const double tr = v.x * cs - v.y * sn, ti = v.x * sn + v.y * cs;
            a[i0] = make_double2(u.x + tr, u.y + ti);
            a[i1] = make_double2(u.x - tr, u.y - ti);
        }
        __syncthreads();
    }
    // one-sided PSD in numpy's operation order (preprocess._psd_spectrogram); psd[] aliases a[k > ws/2]: read first
    for (int k0 = 0; k0 <= ws / 2; k0 += nt) {
        const int k = k0 + tid;
        double p = 0.0;
        if (k <= ws / 2) {
            const double2 v = a[k];
            p = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
            if (k != 0 && k != ws / 2) p *= 2.0;
            p /= fs;
            p /= sumw2;
        }
        __syncthreads();
        if (k <= ws / 2) psd[k] = p;
    }
    __syncthreads();
    double part = 0.0;
    for (int e = tid; e < nerb; e += nt) {
        const int l = lo[e];
        const double slope = (psd[l + 1] - psd[l]) / (f[l + 1] - f[l]);
        const double v = __dadd_rn(__dmul_rn(slope, ferb[e] - f[l]), psd[l]);
        const double L = sqrt(fmax(0.0, v));
        Lb[e] = L;
        part += L * L;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid < 32) {
        double v = tid < (nt >> 5) ? red[tid] : 0.0;
        v = warp_sum(v);
        if (tid == 0) red[0] = v;
    }
    __syncthreads();
    const double nrm = sqrt(red[0]);
    for (int e = tid; e < nerb; e += nt) Ln[(int64_t)e * nseg + seg] = nrm == 0.0 ? 0.0 : Lb[e] / nrm;   // SWIPE.py:141: 0 -> inf
}

// Si[r][seg] = sum_e K[r][e] Ln[e][seg]
__global__ void __launch_bounds__(128) k_swipe_strength(const double* __restrict__ K, int nerb, const double* __restrict__ Ln,
                                                        int64_t nseg, double* __restrict__ Si) {
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (seg >= nseg) return;
    const double* k = K + (int64_t)r * nerb;
    double acc = 0.0;
    for (int e = 0; e < nerb; e++) acc = fma(k[e], Ln[(int64_t)e * nseg + seg], acc);
    Si[(int64_t)r * nseg + seg] = acc;
}

// S[rows[r]][t] += mu[r] * interp(Si[r], t); NaN where t lies outside the frame times of this window size
__global__ void __launch_bounds__(128) k_swipe_accumulate(const double* __restrict__ Si, int64_t nseg, const int32_t* __restrict__ rows,
                                                          const double* __restrict__ mu, const int32_t* __restrict__ hi,
                                                          const double* __restrict__ w1, const double* __restrict__ w2,
                                                          int64_t nt, double* __restrict__ S) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (t >= nt) return;
    const int h = hi[t];                       // < 0: outside
    double v;
    if (h < 0 || nseg <= 1) {
        v = nan("");
    } else {
        const double s0 = Si[(int64_t)r * nseg + h - 1], s1 = Si[(int64_t)r * nseg + h];
        const double sl = (s1 - s0) / w2[t];                     // (Si[hi] - Si[l0]) / (tim[hi] - tim[l0])
        v = __dadd_rn(__dmul_rn(sl, w1[t]), s0);                 // sl * (t - tim[l0]) + Si[l0]
    }
    double* dst = S + (int64_t)rows[r] * nt + t;
    *dst = __dadd_rn(*dst, __dmul_rn(mu[r], v));
}

// numpy.argmax over the candidates of one time frame (the first NaN wins, else the first maximum) and the strengths at
// best-1, best, best+1
__global__ void __launch_bounds__(128) k_swipe_argmax(const double* __restrict__ S, int npc, int64_t nt, int32_t* __restrict__ best,
                                                      double* __restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    int b = 0;
    double m = S[t];
    if (m == m) {
        for (int j = 1; j < npc; j++) {
            const double v = S[(int64_t)j * nt + t];
            if (v != v) { b = j; m = v; break; }
            if (v > m) { m = v; b = j; }
        }
    }
    best[t] = b;
    y[t] = b > 0 ? S[(int64_t)(b - 1) * nt + t] : nan("");
    y[nt + t] = m;
    y[2 * nt + t] = b < npc - 1 ? S[(int64_t)(b + 1) * nt + t] : nan("");
}

// functions.py:618-634: 20 log10 of the population standard deviation (numpy.std) of the 30 ms window around every 5 ms
// frame centre, for the high-passed signal and its low-passed version; one warp per frame, two passes (mean, then
// squared deviations).  Frames whose window leaves the signal keep NaN (the reference leaves their flags False).
__global__ void __launch_bounds__(128) k_vuv_energy(const double* __restrict__ s, const double* __restrict__ sm, int64_t length,
                                                    int wlen, int step, int64_t nframes, double* __restrict__ en,
                                                    double* __restrict__ en_sm) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= nframes) return;
    const int64_t tith = 1 + i * step;                       // functions.py:616: ti = arange(1, length, step)
    const int mid = (wlen - 1) / 2;
    if (!(tith > mid && tith < length - mid)) {
        if (lane == 0) { en[i] = nan(""); en_sm[i] = nan(""); }
        return;
    }
    const int64_t start = tith - mid - 1;                    // s[tith + arange(-mid-1, mid)]
    double a = 0.0, b = 0.0;
    for (int n = lane; n < wlen; n += 32) { a += s[start + n]; b += sm[start + n]; }
    a = warp_sum(a) / (double)wlen;
    b = warp_sum(b) / (double)wlen;
    double va = 0.0, vb = 0.0;
    for (int n = lane; n < wlen; n += 32) {
        const double da = s[start + n] - a, db = sm[start + n] - b;
        va = fma(da, da, va); vb = fma(db, db, vb);
    }
    va = warp_sum(va); vb = warp_sum(vb);
    if (lane == 0) {
        en[i] = 20.0 * log10(sqrt(va / (double)wlen));
        en_sm[i] = 20.0 * log10(sqrt(vb / (double)wlen));
    }
}

// scipy.signal.filtfilt(b, a, x) (default 'pad' method: odd extension by padlen samples, steady-state initial conditions
// lfilter_zi scaled by the first sample, forward pass, backward pass) as the reference uses it for its elliptic filters
// (misc.py:167-182 via functions.py:90-91,579,582), for a BATCH of signals: one thread per signal runs the two recursions
// sequentially in time, in scipy's direct-form-II-transposed operation order and without fused multiply-adds, so the
// output equals scipy's bit for bit.  A corpus filters thousands of utterances per launch; a single long signal is better
// served by the host (the recursion is sequential).
template <int NT>   // taps of b and a (filter order + 1)
__global__ void __launch_bounds__(64) k_filtfilt_batch(const double* __restrict__ x, const int64_t* __restrict__ start,
                                                       const int64_t* __restrict__ len, int nutt, const double* __restrict__ bc,
                                                       const double* __restrict__ ac, const double* __restrict__ zic,
                                                       int padlen, double* __restrict__ y, double* __restrict__ work,
                                                       const int64_t* __restrict__ wstart) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nutt) return;
    double b[NT], a[NT], z[NT - 1];
#pragma unroll
    for (int i = 0; i < NT; i++) { b[i] = bc[i]; a[i] = ac[i]; }
    const double* xs = x + start[u];
    double* ys = y + start[u];
    double* w = work + wstart[u];
    const int64_t L = len[u], E = L + 2 * (int64_t)padlen;
    const double xl = __dmul_rn(2.0, xs[0]), xr = __dmul_rn(2.0, xs[L - 1]);
    // one sample of scipy's _linear_filter (direct form II transposed), its operation order, no contraction
#define EAQHM_FILT_STEP(xn, yn)                                                                            \
    {                                                                                                      \
        yn = __dadd_rn(z[0], __dmul_rn(b[0], xn));                                                         \
        _Pragma("unroll") for (int i = 0; i < NT - 2; i++)                                                 \
            z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(xn, b[i + 1])), __dmul_rn(yn, a[i + 1]));       \
        z[NT - 2] = __dsub_rn(__dmul_rn(xn, b[NT - 1]), __dmul_rn(yn, a[NT - 1]));                         \
    }
    // forward pass over the odd extension (scipy.signal._arraytools.odd_ext) of the signal
    const double x0 = __dsub_rn(xl, xs[padlen]);
#pragma unroll
    for (int i = 0; i < NT - 1; i++) z[i] = __dmul_rn(zic[i], x0);
    for (int64_t i = 0; i < E; i++) {
        double xn, yn;
        if (i < padlen) xn = __dsub_rn(xl, xs[padlen - i]);
        else if (i < padlen + L) xn = xs[i - padlen];
        else xn = __dsub_rn(xr, xs[L - 2 - (i - padlen - L)]);
        EAQHM_FILT_STEP(xn, yn);
        w[i] = yn;
    }
    // backward pass; keep the samples of the signal proper
    const double y0 = w[E - 1];
#pragma unroll
    for (int i = 0; i < NT - 1; i++) z[i] = __dmul_rn(zic[i], y0);
    for (int64_t i = E - 1; i >= padlen; i--) {
        const double xn = w[i];
        double yn;
        EAQHM_FILT_STEP(xn, yn);
        if (i < padlen + L) ys[i - padlen] = yn;
    }
#undef EAQHM_FILT_STEP
}

}  // namespace

extern "C" int eaqhm_vuv_energy(const double* s, const double* s_smooth, int64_t length, int32_t window_len, int32_t step,
                                int64_t nframes, double* energy_db, double* energy_smooth_db, void* stream) {
    CHECK_ARG(s && s_smooth && energy_db && energy_smooth_db && length > 0 && window_len > 0 && (window_len & 1) && step > 0);
    CHECK_ARG(nframes > 0 && nframes * 32 < (1LL << 40));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((nframes * 32 + 127) / 128);
    LAUNCH(KC_SWIPE, st, k_vuv_energy<<<blocks, 128, 0, st>>>(s, s_smooth, length, window_len, step, nframes, energy_db,
                                                              energy_smooth_db));
    return EAQHM_OK;
}

extern "C" int eaqhm_filtfilt_batch(const double* x, const int64_t* start, const int64_t* len, int32_t nutt, const double* b,
                                    const double* a, const double* zi, int32_t ntaps, int32_t padlen, double* y, double* work,
                                    const int64_t* work_start, void* stream) {
    CHECK_ARG(x && start && len && b && a && zi && y && work && work_start && nutt > 0);
    CHECK_ARG(ntaps == 7 && padlen >= 1);        // the reference's filters are all 6th order (misc.py:178)
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH(KC_SWIPE, st, k_filtfilt_batch<7><<<(nutt + 63) / 64, 64, 0, st>>>(x, start, len, nutt, b, a, zi, padlen, y, work,
                                                                            work_start));
    return EAQHM_OK;
}

extern "C" int eaqhm_swipe_loudness(const double* x, int64_t n_x, int32_t ws, int32_t hop, int32_t pad_front,
                                    const double* window, double fs, double sumw2, const double* f, const int32_t* lo,
                                    const double* ferb, int32_t nerb, int64_t nseg, double* Ln, void* stream) {
    CHECK_ARG(x && window && f && lo && ferb && Ln && n_x > 0 && nerb > 0 && nseg > 0 && hop > 0);
    CHECK_ARG(ws >= 8 && ws <= 8192 && (ws & (ws - 1)) == 0);
    int log2ws = 0;
    while ((1 << log2ws) < ws) log2ws++;
    const size_t smem = (size_t)ws * 16 + (size_t)nerb * 8;
    CHECK_ARG(smem <= 200 * 1024 && nseg < (1LL << 31));
    static DeviceOnce once;
    if (const int dev = once.claim(); dev != -1) {
        CHECK_ARG(dev >= 0);
        CUDA_TRY(cudaFuncSetAttribute(k_swipe_loudness, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = ws / 2 >= 512 ? 512 : (ws / 2 < 32 ? 32 : ws / 2);
    LAUNCH(KC_SWIPE, st, k_swipe_loudness<<<(unsigned)nseg, threads, smem, st>>>(x, n_x, ws, log2ws, hop, pad_front, window, fs,
                                                                                  sumw2, f, lo, ferb, nerb, nseg, Ln));
    return EAQHM_OK;
}

extern "C" int eaqhm_swipe_strength(const double* K, int32_t ncand, int32_t nerb, const double* Ln, int64_t nseg, double* Si,
                                    void* stream) {
    CHECK_ARG(K && Ln && Si && ncand > 0 && nerb > 0 && nseg > 0 && ncand <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((nseg + 127) / 128), ncand);
    LAUNCH(KC_SWIPE, st, k_swipe_strength<<<grid, 128, 0, st>>>(K, nerb, Ln, nseg, Si));
    return EAQHM_OK;
}

extern "C" int eaqhm_swipe_accumulate(const double* Si, int32_t ncand, int64_t nseg, const int32_t* rows, const double* mu,
                                      const int32_t* hi, const double* w1, const double* w2, int64_t nt, double* S,
                                      void* stream) {
    CHECK_ARG(Si && rows && mu && hi && w1 && w2 && S && ncand > 0 && ncand <= 65535 && nt > 0 && nseg > 0);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((nt + 127) / 128), ncand);
    LAUNCH(KC_SWIPE, st, k_swipe_accumulate<<<grid, 128, 0, st>>>(Si, nseg, rows, mu, hi, w1, w2, nt, S));
    return EAQHM_OK;
}

extern "C" int eaqhm_swipe_argmax(const double* S, int32_t npc, int64_t nt, int32_t* best, double* y, void* stream) {
    CHECK_ARG(S && best && y && npc > 0 && nt > 0);
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH(KC_SWIPE, st, k_swipe_argmax<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(S, npc, nt, best, y));
    return EAQHM_OK;
}
