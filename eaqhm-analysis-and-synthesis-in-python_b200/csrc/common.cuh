// Shared declarations of the eaQHM B200 library (internal; the public ABI is include/eaqhm_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eaqhm_b200.h"

#define EAQHM_VERSION 100

static constexpr double kPi = 3.141592653589793;
static constexpr double kTwoPi = 2.0 * 3.141592653589793;
static constexpr double kAmEps = 10e-5;       // functions.py:517
static constexpr double kFallbackFm = 140.0;  // functions.py:209
static constexpr double kFallbackAm = 10e-4;  // functions.py:210

void eaqhm_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            eaqhm_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                            cudaGetErrorString(_e));                                     \
            return EAQHM_E_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define CHECK_ARG(cond)                                                                  \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            eaqhm_set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__);  \
            return EAQHM_E_ARG;                                                          \
        }                                                                                \
    } while (0)

// ---- launch accounting / per-kernel-class CUDA-event timing (capi.cu) -----------------------
enum KernelClass {
    KC_FALLBACK = 0, KC_ACTIVE, KC_BASIS, KC_GRAM, KC_CHOL, KC_ACCEPT, KC_RUN_MOMENTS, KC_A0_MOMENTS,
    KC_INTERP, KC_SYNTH, KC_REDUCE, KC_VARPASS, KC_SNAPSHOT, KC_CHOL_UPDATE, KC_SWIPE, KC_COUNT
};
void eaqhm_prof_begin(int kc, cudaStream_t st);
void eaqhm_prof_end(int kc, cudaStream_t st);
// LAUNCH(KC_X, stream, kernel<<<...>>>(...));  counts the launch and, when profiling is on, brackets it with events
#define LAUNCH(kc, st, ...)                                                              \
    do {                                                                                 \
        eaqhm_prof_begin((kc), (st));                                                    \
        __VA_ARGS__;                                                                     \
        eaqhm_prof_end((kc), (st));                                                      \
        CUDA_TRY(cudaGetLastError());                                                    \
    } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- per-device one-time initialisation -------------------------------------------------------
// cudaFuncSetAttribute and __constant__ uploads are per device: every such set-up is keyed by the CURRENT device's
// ordinal (the caller makes the stream's device current: engine.py wraps every call in torch.cuda.device).
constexpr int kMaxDevices = 64;
struct DeviceOnce {
    bool done[kMaxDevices] = {};
    // returns the current device ordinal when its set-up is still to do (and marks it done), -1 otherwise, -2 on error
    int claim() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -2;
        if (done[dev]) return -1;
        done[dev] = true;
        return dev;
    }
};
int eaqhm_sm_count();   // multiprocessors of the current device (cached per device)

// ---- complex helpers on double2 (x = re, y = im) ---------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfma_conj_a(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= a * conj(b)
__device__ __forceinline__ void cfms_conj_b(double2& acc, double2 a, double2 b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.y, b.x, acc.y); acc.y = fma(a.x, b.y, acc.y);
}
// acc += conj(a) * b  (same as cfma_conj_a; named for the back substitution)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- analysis stage internals (analysis.cu) -------------------------------------------
struct AnaLayout {
    int32_t nnz_max, Mxmax, Pmax, Pp, Lp;  // Mxmax = 2*nnz_max generic columns, Pmax = 2*Mxmax+2, Pp = row pitch of R, Lp = pitch of X rows
    int32_t nsets;                         // copies of the per-chunk scratch (chunk pipeline depth)
    int64_t off_fb, off_M, off_act, off_slow, off_X, off_R, off_x, off_cnt, set_stride, total;
};
constexpr int kMaxSets = 3;
constexpr int kDefaultSets = 1;   // chunk pipeline measured and not shipped (profiles/r02_chol_experiments.md): EAQHM_PIPELINE=2 enables it
int eaqhm_pipeline_sets();
AnaLayout eaqhm_ana_layout(int32_t G, int32_t chunk, int32_t nnz_max, int32_t N_max);

// Solves `nf` already-built frames: X [nf][2][Mxmax][Lp] (windowed basis planes), Mcols[nf] generic
// columns, Lrows[nf] window samples; dc = 1 when the DC rows (and the rhs row) were written by the basis
// kernel.  Fills R (P = 2*Mcols + 2*dc), runs the Cholesky solve, leaves x [nf][Pmax].
struct SolveArgs {
    const double* X; double2* R; double2* x; const int32_t* Mcols;
    const int32_t* Lrows;
    int32_t nf, Mxmax, dc, Pmax, Pp, Lp;
    int32_t* status;
    int32_t* counter;      // work-queue head of the Gram / panel update kernel (zeroed by the launcher)
    int32_t* squeue = nullptr;   // frame-queue head of the persistent Cholesky launches of the chunk pipeline (else unused)
};
int eaqhm_launch_gram_solve(const SolveArgs& a, cudaStream_t st, bool do_gram);
// The two halves of the above.  The Gram kernel is a persistent grid that takes (frame, tile) items off the queue at
// a.counter: `ctas_per_sm` sizes the grid, `reset_queue` zeroes the queue head first; a second launch with
// reset_queue = false (any stream, ordered after the reset) joins the same queue.
int eaqhm_launch_gram(const SolveArgs& a, cudaStream_t st, int ctas_per_sm, bool reset_queue);
int eaqhm_launch_solve(const SolveArgs& a, cudaStream_t st);
// Persistent form of the single-kernel Cholesky (systems below the panel threshold): `ctas_per_sm` CTAs per SM take
// frames off the queue at a.squeue; a launch with reset_queue = false joins a queue that is already being served.
int eaqhm_launch_solve_queue(const SolveArgs& a, cudaStream_t st, int ctas_per_sm, bool reset_queue);
bool eaqhm_solve_uses_panels(const SolveArgs& a);

// spline helpers shared by interp.cu and synth.cu
__device__ __forceinline__ double spline_eval(double y0, double y1, double m0, double m1, double h, double t1) {
    // moment form of the cubic piece on [x_j, x_j + h] at offset t1 (functions.py:340,367 restated)
    const double t0 = h - t1;
    return (m0 * t0 * t0 * t0 + m1 * t1 * t1 * t1) / (6.0 * h) + (y0 / h - m0 * h / 6.0) * t0 +
           (y1 / h - m1 * h / 6.0) * t1;
}
