/* mr_oracle.c -- C restatement of Multirate.jl's streaming polyphase FIR loops.
 *
 * TEST INFRASTRUCTURE ONLY.  Not linked into, loaded by or called from the
 * product library (multirate.jl_b200/csrc/libmrb.so).  Used by tests/ as a
 * second checker, and by bench.py's `cpu_baseline` / `--impl reference` legs as
 * the timed CPU baseline ("port": the Julia reference cannot run here -- no
 * julia binary in the image, and the source is Julia-0.3 syntax).
 *
 * Structure is deliberately the reference's: one state machine per channel
 * (the reference has one FIRFilter per Vector), per output a branch on "window
 * straddles history" vs "window inside x" (src/Filters.jl:560-564), one forward
 * dot product over the flipped bank column accumulated in the promoted type
 * (src/support.jl:5-55; `#pragma omp simd reduction` stands in for @simd), scalar
 * phase / index update (src/Filters.jl:567-568, 663-673, 780-792), shiftin! at the
 * end (src/support.jl:61-80).  Channels are split over OpenMP threads.
 *
 * Parity pinning: validated in tests/test_oracle.py against oracle/
 * multirate_oracle.py, which is itself pinned on the reference's known-answer
 * vectors (README.md:58-142 etc.).  Farrow VALUES: parity unpinned upstream.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { K_STANDARD = 0, K_INTERPOLATOR = 1, K_DECIMATOR = 2, K_RATIONAL = 3, K_ARBITRARY = 4, K_FARROW = 5 };
enum { D_F32 = 0, D_F64 = 1, D_C64 = 2, D_C128 = 3 };

typedef struct {
    int kind, th, tx;
    long hLen, L, M, Nphi, T, H, nch;
    int polyorder;
    double rate, delta;
    void *bank;    /* Th[T*Nphi], column phi contiguous (Julia column-major pfb) ; Nphi = 1 for standard/decimator */
    void *dbank;   /* arbitrary: derivative bank */
    double *pnfb;  /* farrow: T x (order+1), lowest order first, values representable in Th */
    /* carried state, 1-based as upstream */
    long phiIdx, inputDeficit, xIdx;
    double acc, alpha; /* arbitrary: phiAccumulator, alpha ; farrow: acc = Float64 phiIdx */
    void *history;     /* Tx[nch*H] */
} mro_filter;

static size_t dsize(int d) { return d == D_F32 ? 4 : d == D_F64 ? 8 : d == D_C64 ? 8 : 16; }

/* src/Filters.jl:284-298 */
static void taps2pfb_d(const double *h, long hLen, long Nphi, long T, double *pfb) {
    long hIdx = 0;
    for (long row = T - 1; row >= 0; --row)
        for (long col = 0; col < Nphi; ++col) {
            pfb[col * T + row] = hIdx < hLen ? h[hIdx] : 0.0;
            ++hIdx;
        }
}

static void *to_th(const double *src, long n, int th) {
    if (th == D_F32) { float *p = malloc(sizeof(float) * (n ? n : 1)); for (long i = 0; i < n; ++i) p[i] = (float)src[i]; return p; }
    double *p = malloc(sizeof(double) * (n ? n : 1)); memcpy(p, src, sizeof(double) * n); return p;
}

/* h: taps already rounded to Th, passed as double.  pnfb: T*(order+1) host-fitted
 * coefficients (farrow only).  kind is chosen by the caller exactly as
 * FIRFilter(h, ratio) does (src/Filters.jl:158-198). */
mro_filter *mro_create(int kind, int th, int tx, const double *h, long hLen, long L, long M, double rate,
                       long Nphi, int polyorder, const double *pnfb, long nch) {
    mro_filter *f = calloc(1, sizeof(*f));
    f->kind = kind; f->th = th; f->tx = tx; f->hLen = hLen; f->L = L; f->M = M; f->nch = nch;
    f->rate = rate; f->polyorder = polyorder;
    if (kind == K_STANDARD || kind == K_DECIMATOR) {
        f->Nphi = 1; f->T = hLen;
        double *fl = malloc(sizeof(double) * hLen);
        for (long i = 0; i < hLen; ++i) fl[i] = h[hLen - 1 - i];   /* flipud :21,:53 */
        f->bank = to_th(fl, hLen, th); free(fl);
    } else {
        f->Nphi = (kind == K_ARBITRARY || kind == K_FARROW) ? Nphi : L;
        f->T = (hLen + f->Nphi - 1) / f->Nphi;
        double *p = malloc(sizeof(double) * f->T * f->Nphi);
        taps2pfb_d(h, hLen, f->Nphi, f->T, p);
        f->bank = to_th(p, f->T * f->Nphi, th);
        if (kind == K_ARBITRARY) {            /* dh = [diff(h); 0] in Th arithmetic, :106 */
            double *dh = malloc(sizeof(double) * hLen);
            for (long i = 0; i + 1 < hLen; ++i) dh[i] = th == D_F32 ? (double)((float)h[i + 1] - (float)h[i]) : h[i + 1] - h[i];
            dh[hLen - 1] = 0.0;
            taps2pfb_d(dh, hLen, f->Nphi, f->T, p);
            f->dbank = to_th(p, f->T * f->Nphi, th); free(dh);
        }
        free(p);
        if (kind == K_FARROW) {
            f->pnfb = malloc(sizeof(double) * f->T * (polyorder + 1));
            memcpy(f->pnfb, pnfb, sizeof(double) * f->T * (polyorder + 1));
        }
    }
    f->H = f->T - 1;                           /* :165,168,171,174,186,195 */
    f->delta = (kind == K_ARBITRARY || kind == K_FARROW) ? (double)f->Nphi / rate : 0.0;
    f->history = calloc((size_t)(f->H * nch + 1), dsize(tx));
    f->phiIdx = 1; f->inputDeficit = 1; f->xIdx = 1; f->acc = 1.0; f->alpha = 0.0;
    return f;
}

void mro_destroy(mro_filter *f) {
    if (!f) return;
    free(f->bank); free(f->dbank); free(f->pnfb); free(f->history); free(f);
}

void mro_reset(mro_filter *f) {
    memset(f->history, 0, (size_t)(f->H * f->nch) * dsize(f->tx));
    f->phiIdx = 1; f->inputDeficit = 1; f->xIdx = 1; f->acc = 1.0; f->alpha = 0.0;
}

void mro_get_state(const mro_filter *f, long *phiIdx, long *deficit, double *acc, double *alpha) {
    *phiIdx = f->phiIdx; *deficit = f->inputDeficit; *acc = f->acc; *alpha = f->alpha;
}

/* exact count + end state are produced by running the loop; this is the reference's
 * own outputlength (an upper bound for arbitrary/farrow), src/Filters.jl:352-385 */
long mro_outputlength(const mro_filter *f, long n_in) {
    switch (f->kind) {
    case K_STANDARD: return n_in;
    case K_INTERPOLATOR: return f->L * n_in;
    case K_DECIMATOR: return (long)ceil((double)(n_in - f->inputDeficit + 1) / (double)f->M);
    case K_RATIONAL: return (long)ceil((double)((n_in - f->inputDeficit + 1) * f->L - f->phiIdx + 1) / (double)f->M);
    default: return (long)ceil((double)(n_in - f->inputDeficit + 1) * f->rate);
    }
}

typedef struct { long phiIdx, inputDeficit, xIdx; double acc, alpha; } kstate;

#define DEFINE_ALL(SFX, TH, TX, TY)                                                                          \
    /* src/support.jl:5-14 / :33-42 : window entirely inside x; a = bank column */                         \
    static inline TY dot_x_##SFX(const TH *a, long aLen, const TX *b, long bLastIdx) {                       \
        const TX *bp = b + (bLastIdx - aLen);                                                                \
        TY acc = (TY)a[0] * (TY)bp[0];                                                                       \
        _Pragma("omp simd reduction(+:acc)")                                                                 \
        for (long i = 1; i < aLen; ++i) acc += (TY)a[i] * (TY)bp[i];                                         \
        return acc;                                                                                          \
    }                                                                                                        \
    /* src/support.jl:16-31 / :44-55 : window straddles history b (len aLen-1) and x = c */                 \
    static inline TY dot_hx_##SFX(const TH *a, long aLen, const TX *b, const TX *c, long cLastIdx) {         \
        TY acc = 0;                                                                                          \
        _Pragma("omp simd reduction(+:acc)")                                                                 \
        for (long i = 0; i < aLen - cLastIdx; ++i) acc += (TY)a[i] * (TY)b[i + cLastIdx - 1];                \
        _Pragma("omp simd reduction(+:acc)")                                                                 \
        for (long i = 0; i < cLastIdx; ++i) acc += (TY)a[aLen - cLastIdx + i] * (TY)c[i];                    \
        return acc;                                                                                          \
    }                                                                                                        \
    static inline TY dot_##SFX(const TH *a, long T, const TX *hist, const TX *x, long idx) {                 \
        return idx < T ? dot_hx_##SFX(a, T, hist, x, idx) : dot_x_##SFX(a, T, x, idx);                       \
    }                                                                                                        \
    /* src/support.jl:61-80 */                                                                               \
    static void shiftin_##SFX(TX *a, long aLen, const TX *b, long bLen) {                                    \
        if (bLen >= aLen) memcpy(a, b + (bLen - aLen), sizeof(TX) * aLen);                                   \
        else { memmove(a, a + bLen, sizeof(TX) * (aLen - bLen)); memcpy(a + (aLen - bLen), b, sizeof(TX) * bLen); } \
    }                                                                                                        \
    /* one channel, one chunk; returns the output count; *s is that channel's private kernel state */       \
    static long chan_##SFX(const mro_filter *f, kstate *s, TX *hist, const TX *x, long xLen, TY *y, TH *cur) { \
        const TH *bank = (const TH *)f->bank; const TH *dbank = (const TH *)f->dbank;                       \
        const long T = f->T, H = f->H; long n = 0;                                                          \
        switch (f->kind) {                                                                                   \
        case K_STANDARD: /* src/Filters.jl:450-473 */                                                        \
            for (long yIdx = 1; yIdx <= xLen; ++yIdx) y[n++] = dot_##SFX(bank, T, hist, x, yIdx);            \
            break;                                                                                           \
        case K_INTERPOLATOR: { /* :489-517 */                                                                \
            long phi = 1, inputIdx = 1; const long outLen = f->L * xLen;                                     \
            for (long yIdx = 1; yIdx <= outLen; ++yIdx) {                                                    \
                y[n++] = dot_##SFX(bank + (phi - 1) * T, T, hist, x, inputIdx);                              \
                if (phi == f->Nphi) { phi = 1; ++inputIdx; } else ++phi;                                     \
            }                                                                                                \
            break; }                                                                                         \
        case K_DECIMATOR: { /* :598-631 (early-out per SURVEY 9.3) */                                        \
            if (xLen < s->inputDeficit) { shiftin_##SFX(hist, H, x, xLen); s->inputDeficit -= xLen; return 0; } \
            long inputIdx = s->inputDeficit;                                                                 \
            while (inputIdx <= xLen) { y[n++] = dot_##SFX(bank, T, hist, x, inputIdx); inputIdx += f->M; }   \
            s->inputDeficit = inputIdx - xLen;                                                               \
            break; }                                                                                         \
        case K_RATIONAL: { /* :536-575 */                                                                    \
            if (xLen < s->inputDeficit) { shiftin_##SFX(hist, H, x, xLen); s->inputDeficit -= xLen; return 0; } \
            const long step = f->M % f->L; long inputIdx = s->inputDeficit;                                  \
            while (inputIdx <= xLen) {                                                                       \
                y[n++] = dot_##SFX(bank + (s->phiIdx - 1) * T, T, hist, x, inputIdx);                        \
                inputIdx += (long)floor((double)(s->phiIdx + f->M - 1) / (double)f->L);   /* :567 */          \
                long nx = s->phiIdx + step; s->phiIdx = nx > f->L ? nx - f->L : nx;      /* :433-439 */      \
            }                                                                                                \
            s->inputDeficit = inputIdx - xLen;                                                               \
            break; }                                                                                         \
        case K_ARBITRARY: { /* :693-742 */                                                                   \
            if (xLen < s->inputDeficit) { shiftin_##SFX(hist, H, x, xLen); s->inputDeficit -= xLen; return 0; } \
            s->xIdx = s->inputDeficit;                                                                       \
            while (s->xIdx <= xLen) {                                                                        \
                TY lo = dot_##SFX(bank + (s->phiIdx - 1) * T, T, hist, x, s->xIdx);                          \
                TY up = dot_##SFX(dbank + (s->phiIdx - 1) * T, T, hist, x, s->xIdx);                         \
                y[n++] = (TY)(lo + up * s->alpha);                                        /* :730 */          \
                s->acc += f->delta;                                                       /* :663-673 */      \
                if (s->acc > (double)f->Nphi) {                                                              \
                    s->xIdx += (long)floor((s->acc - 1.0) / (double)f->Nphi);                                \
                    s->acc = fmod(s->acc - 1.0, (double)f->Nphi) + 1.0;                                      \
                }                                                                                            \
                s->phiIdx = (long)floor(s->acc); s->alpha = s->acc - (double)s->phiIdx;                      \
            }                                                                                                \
            s->inputDeficit = s->xIdx - xLen;                                                                \
            break; }                                                                                         \
        case K_FARROW: { /* :795-836 */                                                                      \
            if (xLen < s->inputDeficit) { shiftin_##SFX(hist, H, x, xLen); s->inputDeficit -= xLen; return 0; } \
            s->xIdx = s->inputDeficit; const int P = f->polyorder;                                           \
            for (long i = 0; i < T; ++i) { /* currentTaps at the carried phase */                            \
                const double *c = f->pnfb + i * (P + 1); double v = c[P];                                    \
                for (int p = P - 1; p >= 0; --p) v = v * s->acc + c[p];                                      \
                cur[i] = (TH)v; }                                                                            \
            while (s->xIdx <= xLen) {                                                                        \
                y[n++] = dot_##SFX(cur, T, hist, x, s->xIdx);                                                \
                s->acc += f->delta;                                                       /* :780-786 */      \
                if (s->acc > (double)f->Nphi) {                                                              \
                    s->xIdx += (long)floor((s->acc - 1.0) / (double)f->Nphi);                                \
                    s->acc = fmod(s->acc - 1.0, (double)f->Nphi) + 1.0;                                      \
                }                                                                                            \
                for (long i = 0; i < T; ++i) {                                            /* :789-791 */      \
                    const double *c = f->pnfb + i * (P + 1); double v = c[P];                                \
                    for (int p = P - 1; p >= 0; --p) v = v * s->acc + c[p];                                  \
                    cur[i] = (TH)v; }                                                                        \
            }                                                                                                \
            s->inputDeficit = s->xIdx - xLen;                                                                \
            break; }                                                                                         \
        }                                                                                                    \
        shiftin_##SFX(hist, H, x, xLen);                                                                     \
        return n;                                                                                            \
    }                                                                                                        \
    static long run_##SFX(mro_filter *f, const void *xv, long ldx, long xLen, void *yv, long ldy, int nthreads) { \
        const TX *x = (const TX *)xv; TY *y = (TY *)yv; long count = 0; kstate fin;                         \
        kstate s0 = { f->phiIdx, f->inputDeficit, f->xIdx, f->acc, f->alpha };                               \
        fin = s0;                                                                                            \
        _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                                   \
        for (long c = 0; c < f->nch; ++c) {                                                                  \
            kstate s = s0; TH *cur = (TH *)malloc(sizeof(TH) * (f->T + 1));                                  \
            long n = chan_##SFX(f, &s, (TX *)f->history + c * f->H, x + c * ldx, xLen, y + c * ldy, cur);    \
            free(cur);                                                                                       \
            if (c == 0) { count = n; fin = s; }                                                              \
        }                                                                                                    \
        f->phiIdx = fin.phiIdx; f->inputDeficit = fin.inputDeficit; f->xIdx = fin.xIdx;                      \
        f->acc = fin.acc; f->alpha = fin.alpha;                                                              \
        return count;                                                                                        \
    }

DEFINE_ALL(f32_f32, float, float, float)
DEFINE_ALL(f32_f64, float, double, double)
DEFINE_ALL(f32_c64, float, float complex, float complex)
DEFINE_ALL(f32_c128, float, double complex, double complex)
DEFINE_ALL(f64_f32, double, float, double)
DEFINE_ALL(f64_f64, double, double, double)
DEFINE_ALL(f64_c64, double, float complex, double complex)
DEFINE_ALL(f64_c128, double, double complex, double complex)

/* x: nch rows of ldx samples (time contiguous); y: nch rows of ldy; y must hold
 * mro_outputlength() samples per row.  Returns the per-channel output count. */
long mro_filt(mro_filter *f, const void *x, long ldx, long n_in, void *y, long ldy, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    switch (f->th * 4 + f->tx) {
    case D_F32 * 4 + D_F32: return run_f32_f32(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F32 * 4 + D_F64: return run_f32_f64(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F32 * 4 + D_C64: return run_f32_c64(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F32 * 4 + D_C128: return run_f32_c128(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F64 * 4 + D_F32: return run_f64_f32(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F64 * 4 + D_F64: return run_f64_f64(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F64 * 4 + D_C64: return run_f64_c64(f, x, ldx, n_in, y, ldy, nthreads);
    case D_F64 * 4 + D_C128: return run_f64_c128(f, x, ldx, n_in, y, ldy, nthreads);
    }
    return -1;
}

int mro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
