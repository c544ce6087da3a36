/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference AKAZE hot path (see akaze_oracle.h for
 * the rules about who may use it, the list of deliberate differences and the pinning status).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is written as fmaf() on purpose and
 * every separately rounded product as a plain '*'; the patterns were read from the SASS of the
 * reference's sm_100a build (nvcc 12.9, default -fmad=true, no fast-math).
 */
#define _GNU_SOURCE
#include "akaze_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_LEVELS 64
#define ORC_NBINS 300           /* akazed.cu:8 */
#define ORC_MAX_DIST 96         /* akazed.cu:11 */

/* reflect-101: akazed.cu:162-170 (borderAdd) and the abs(ix - i) idiom used beside it */
static inline int refl_lo(int i) { return i < 0 ? -i : i; }
static inline int refl_hi(int i, int m) { return i < m ? i : m + m - 2 - i; }
static inline int refl(int i, int m) { return refl_hi(refl_lo(i), m); }

void orc_default_options(orc_options* o)
{
    o->noctaves = 4; o->max_scale = 4; o->per = 0.7f; o->soffset = 1.6f; o->reordering = 1;
    o->derivative_factor = 1.5f; o->dthreshold = 0.001f; o->diffusivity = 1; o->pattern = 10;
    o->kcontrast_override = 0.f; o->threads = 0;
}

/* ---------------------------------------------------------------------------------------------
 * FED time steps — fed.cpp:41-119 (fed_tau_by_process_time -> by_cycle_time -> internal) and
 * the primality helper fed.cpp:122-148.  Mixed float/double arithmetic kept as in the source.
 * ------------------------------------------------------------------------------------------- */
static int orc_is_prime(int v)
{
    if (v <= 1) return 0;
    if (v == 2 || v == 3 || v == 5 || v == 7) return 1;
    if (v % 2 == 0 || v % 3 == 0 || v % 5 == 0 || v % 7 == 0) return 0;
    int upper = (int)sqrt(v + 1.0);
    for (int d = 11; d <= upper; d += 2) if (v % d == 0) return 0;
    return 1;
}

int orc_fed_tau(float T, int M, float tau_max, int reordering, float* tau, int cap)
{
    float t = T / (float)M;
    int n = (int)(ceil(sqrt(3.0 * t / tau_max + 0.25f) - 0.5f - 1.0e-8f) + 0.5f);
    float scale = (float)(3.0 * t / (tau_max * (float)(n * (n + 1))));
    if (n <= 0) return 0;
    if (n > cap) return -n;
    float c = 1.0f / (4.0f * (float)n + 2.0f);
    float d = scale * tau_max / 2.0f;
    float* th = (float*)malloc(sizeof(float) * n);
    for (int k = 0; k < n; ++k) {
        float h = (float)cos(M_PI * (2.0f * (float)k + 1.0f) * c);
        th[k] = d / (h * h);
    }
    if (!reordering) {
        memcpy(tau, th, sizeof(float) * n);
    } else {
        int kappa = n / 2, prime = n + 1;
        while (!orc_is_prime(prime)) prime++;
        for (int k = 0, l = 0; l < n; ++k, ++l) {
            int index;
            while ((index = ((k + 1) * kappa) % prime - 1) >= n) k++;
            tau[l] = th[index];
        }
    }
    free(th);
    return n;
}

/* akazed.cu:2298-2333 (createGaussKernel): host float math, taps 0..radius */
void orc_gauss_taps(float var, int radius, float* k)
{
    float denom = 1.f / (2.f * var);
    float ksum = 0.f;
    for (int i = 0; i <= radius; i++) {
        k[i] = expf(-i * i * denom);
        ksum += (i == 0) ? k[i] : k[i] + k[i];
    }
    ksum = 1 / ksum;
    for (int i = 0; i <= radius; i++) k[i] *= ksum;
}

/* main.cpp:149: cv::Mat::convertTo(CV_32FC1, 1.0/255.0) on CV_8U evaluates float(v)*float(alpha) */
void orc_u8_to_f32(const unsigned char* src, float* dst, int w, int h, int sp, int dp)
{
    const float a = (float)(1.0 / 255.0);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(size_t)y * dp + x] = (float)src[(size_t)y * sp + x] * a;
}

/* one symmetric tap pass: acc = (x-1 + x+1)*k1 ; acc = fma(x0,k0,acc) ; acc = fma(x-i + x+i, ki, acc)
 * — akazed.cu:225-238 / :281-286 as compiled (the k1 product is the separately rounded one) */
#define TAPS(R, K, C, AT)                                             \
    float acc = ((AT(-1)) + (AT(+1))) * (K)[1];                       \
    acc = fmaf((C), (K)[0], acc);                                     \
    for (int i_ = 2; i_ <= (R); i_++) acc = fmaf((AT(-i_)) + (AT(+i_)), (K)[i_], acc);

/* akazed.cu:204-290 gConv2d<R> + akazed.cu:2336-2386 hLowPass (R from ksz) */
void orc_lowpass(const float* src, float* dst, int w, int h, int p, float var, int ksz)
{
    int R = ksz <= 5 ? 2 : ksz <= 7 ? 3 : ksz <= 9 ? 4 : 5;
    float k[8];
    orc_gauss_taps(var, R, k);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float* row = src + (size_t)y * p;
        for (int x = 0; x < w; x++) {
#define AT(o) row[refl(x + (o), w)]
            TAPS(R, k, row[x], AT)
#undef AT
            tmp[(size_t)y * w + x] = acc;
        }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
#define AT(o) tmp[(size_t)refl(y + (o), h) * w + x]
            TAPS(R, k, tmp[(size_t)y * w + x], AT)
#undef AT
            dst[(size_t)y * p + x] = acc;
        }
    }
    free(tmp);
}

/* akazed.cu:449-511 gDownWithSmooth: dst = src(2x,2y); smooth = sigma=1 R=2 blur evaluated on the
 * coarse lattice (taps at source offsets 0,+-2,+-4, reflection in SOURCE coordinates) */
void orc_down_with_smooth(const float* src, float* dst, float* smooth, int sw, int sh, int sp, int dw, int dh, int dp)
{
    float k[3];
    orc_gauss_taps(1.f, 2, k);
    /* row pass is needed at source rows refl(2y + 2j): compute it for every source row once */
    float* tmp = (float*)malloc(sizeof(float) * (size_t)dw * sh);
#pragma omp parallel for schedule(static)
    for (int sy = 0; sy < sh; sy++) {
        const float* row = src + (size_t)sy * sp;
        for (int x = 0; x < dw; x++) {
            int sx = x + x;
#define AT(o) row[refl(sx + 2 * (o), sw)]
            TAPS(2, k, row[sx], AT)
#undef AT
            tmp[(size_t)sy * dw + x] = acc;
        }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; y++) {
        int sy = y + y;
        for (int x = 0; x < dw; x++) {
#define AT(o) tmp[(size_t)refl(sy + 2 * (o), sh) * dw + x]
            TAPS(2, k, tmp[(size_t)sy * dw + x], AT)
#undef AT
            smooth[(size_t)y * dp + x] = acc;
            dst[(size_t)y * dp + x] = src[(size_t)sy * sp + x + x];
        }
    }
    free(tmp);
}

/* 3x3 neighbourhood fetch with reflect-101 at distance s */
#define NB9(src, p, w, h, x, y, s)                                          \
    int x0_ = refl_lo((x) - (s)), x2_ = refl_hi((x) + (s), (w));            \
    int y0_ = refl_lo((y) - (s)), y2_ = refl_hi((y) + (s), (h));            \
    const float* r0_ = (src) + (size_t)y0_ * (p);                           \
    const float* r1_ = (src) + (size_t)(y) * (p);                           \
    const float* r2_ = (src) + (size_t)y2_ * (p);                           \
    float ul = r0_[x0_], uc = r0_[(x)], ur = r0_[x2_];                      \
    float cl = r1_[x0_], cr = r1_[x2_];                                     \
    float ll = r2_[x0_], lc = r2_[(x)], lr = r2_[x2_];

/* akazed.cu:664-665 / :1088-1089 as compiled */
static inline void scharr_dxdy(float ul, float uc, float ur, float cl, float cr, float ll, float lc, float lr, float* dx, float* dy)
{
    *dx = fmaf(cr - cl, 10.f, 3.f * (((ur + lr) - ul) - ll));
    *dy = fmaf(lc - uc, 10.f, 3.f * (((lr + ll) - ul) - ur));
}

/* akazed.cu:644-667 gScharrContrastNaive: sqrt(dx*dx + dy*dy) compiles to sqrt_rn(fma(dx,dx,dy*dy)) */
void orc_scharr_mag(const float* src, float* mag, int w, int h, int p)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            NB9(src, p, w, h, x, y, 1)
            float dx, dy;
            scharr_dxdy(ul, uc, ur, cl, cr, ll, lc, lr, &dx, &dy);
            mag[(size_t)y * p + x] = sqrtf(fmaf(dx, dx, dy * dy));
        }
}

/* akazed.cu:2410-2484 hScharrContrast host logic + :901-938 histogram.  hmax = max(0.03, TRUE max)
 * (B-1), in-image pixels only (B-3).  bin = (int)__fmul_rz(mag, 300/hmax), clamp 299. */
static inline float mul_rz(float a, float b)
{
    double d = (double)a * (double)b;            /* exact: 24x24 bits fit in 53 */
    float f = (float)d;                          /* round to nearest */
    if (fabs((double)f) > fabs(d)) f = nextafterf(f, 0.f);
    return f;
}

float orc_contrast_from_mag(const float* mag, int w, int h, int p, float per, float* hmax_out)
{
    float hmax = 0.03f;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) if (mag[(size_t)y * p + x] > hmax) hmax = mag[(size_t)y * p + x];
    if (hmax_out) *hmax_out = hmax;
    int hist[ORC_NBINS];
    memset(hist, 0, sizeof(hist));
    float hfactor = ORC_NBINS / hmax;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int hi = (int)mul_rz(mag[(size_t)y * p + x], hfactor);
            if (hi >= ORC_NBINS) hi = ORC_NBINS - 1;
            hist[hi]++;
        }
    int thresh = (int)((w * h - hist[0]) * per);
    int cum = 0, k = 1;
    while (k < ORC_NBINS) {
        if (cum >= thresh) break;
        cum += hist[k];
        k++;
    }
    return k / hfactor;
}

/* akazed.cu:1068-1107 gFlowNaive + :2487-2506 hFlow (ikc = 1/(k*k) on the host) */
void orc_flow(const float* src, float* flow, int type, float kcontrast, int w, int h, int p)
{
    float ikc = 1.f / (kcontrast * kcontrast);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            NB9(src, p, w, h, x, y, 1)
            float dx, dy;
            scharr_dxdy(ul, uc, ur, cl, cr, ll, lc, lr, &dx, &dy);
            float d = fmaf(dx, dx, dy * dy) * ikc;
            float g;
            if (type == 0) g = expf(-d);                                   /* PM_G1: __expf on the GPU */
            else if (type == 1) g = 1.f / (1.f + d);                       /* PM_G2: IEEE division     */
            else if (type == 2) g = 1.f - expf(-3.315f / powf(d, 4.f));    /* WEICKERT                 */
            else g = 1.f / sqrtf(1.f + d);                                 /* CHARBONNIER              */
            flow[(size_t)y * p + x] = g;
        }
}

/* akazed.cu:1241-1264 gNldStepNaive as compiled: the LEFT term is the separately rounded product,
 * right/down/up are folded in by FMA in that order; stepfac = 0.5*tau (akazed.cu:2515) */
void orc_nld_step(const float* src, const float* flow, float* dst, float tau, int w, int h, int p)
{
    float stepfac = 0.5f * tau;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        int y0 = refl_lo(y - 1), y2 = refl_hi(y + 1, h);
        for (int x = 0; x < w; x++) {
            int x0 = refl_lo(x - 1), x2 = refl_hi(x + 1, w);
            size_t c = (size_t)y * p + x;
            float L0 = src[c], g0 = flow[c];
            float s = (g0 + flow[(size_t)y * p + x0]) * (src[(size_t)y * p + x0] - L0);
            s = fmaf(g0 + flow[(size_t)y * p + x2], src[(size_t)y * p + x2] - L0, s);
            s = fmaf(g0 + flow[(size_t)y2 * p + x], src[(size_t)y2 * p + x] - L0, s);
            s = fmaf(g0 + flow[(size_t)y0 * p + x], src[(size_t)y0 * p + x] - L0, s);
            dst[c] = fmaf(s, stepfac, L0);
        }
    }
}

/* akazed.cu:1267-1331 gDerivate + gHessianDeterminant, :2531-2560 hHessianDeterminant.
 * NOTE the two kernels contract differently: first derivatives fma(diff,fac2,fac1*sum), second
 * derivatives fma(sum,fac1,fac2*diff). */
void orc_hessian(const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int p)
{
    float wgt = 10.f / 3.f;
    float fac1 = 1.f / (2.f * (wgt + 2.f));
    float fac2 = wgt * fac1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            NB9(smooth, p, w, h, x, y, step)
            (void)uc; (void)lc;
            lx[(size_t)y * p + x] = fmaf(cr - cl, fac2, fac1 * (((ur + lr) - ul) - ll));
            ly[(size_t)y * p + x] = fmaf(lc - uc, fac2, fac1 * (((lr + ll) - ur) - ul));
        }
    float* out = det;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float dxx, dxy, dyy;
            {
                NB9(lx, p, w, h, x, y, step)
                dxx = fmaf(((ur + lr) - ul) - ll, fac1, fac2 * (cr - cl));
                dxy = fmaf(((lr + ll) - ur) - ul, fac1, fac2 * (lc - uc));
            }
            {
                NB9(ly, p, w, h, x, y, step)
                (void)cl; (void)cr;
                dyy = fmaf(((lr + ll) - ur) - ul, fac1, fac2 * (lc - uc));
            }
            out[(size_t)y * p + x] = fmaf(dxx, dyy, -(dxy * dxy));
        }
}

/* ---------------------------------------------------------------------------------------------
 * Pyramid: schedule akaze.cpp:268-363, loop akaze.cpp:300-440
 * ------------------------------------------------------------------------------------------- */
typedef struct orc_level {
    int octave, sub, w, h, pitch, nsteps, sigma_size;
    float esigma, size, border;
    float tau[128];
    float *lt, *det, *lx, *ly;
} orc_level;

struct orc_pyramid {
    orc_options opt;
    int w, h, noct, nlev;
    orc_level lev[ORC_MAX_LEVELS];
    float kcontrast;
    float *smooth, *flow, *tmp;      /* octave-0 sized scratch */
    float *resp; int* layer;         /* full-resolution merge maps */
    int psz;
};

static int align_up(int a, int b) { return (a % b) ? a - a % b + b : a; }

orc_pyramid* orc_pyramid_create(int w, int h, const orc_options* o)
{
    orc_pyramid* P = (orc_pyramid*)calloc(1, sizeof(orc_pyramid));
    P->opt = *o; P->w = w; P->h = h;
    /* akaze.cpp:204-237: octave j dropped when w or h < 80 */
    int ow[16], oh[16];
    ow[0] = w; oh[0] = h; P->noct = 1;
    for (int j = 1; j < o->noctaves; j++) {
        ow[j] = ow[j - 1] >> 1; oh[j] = oh[j - 1] >> 1;
        if (ow[j] < 80 || oh[j] < 80) break;
        P->noct = j + 1;
    }
    int S = o->max_scale;
    float tmax = 0.25f, soffset = o->soffset;
    float last_etime = (float)(0.5 * soffset * soffset);
    float smax = (float)(10.0 * sqrtf(2.0f));                       /* akaze.cpp:279 (M-LDB) */
    int oratio = 1;
    float psz = 10000.f;
    P->nlev = 0;
    for (int i = 0; i < P->noct; i++) {
        for (int j = 0; j < S; j++) {
            orc_level* L = &P->lev[P->nlev++];
            L->octave = i; L->sub = j; L->w = ow[i]; L->h = oh[i]; L->pitch = align_up(ow[i], 32);
            if (i == 0 && j == 0) {
                L->esigma = soffset;
                L->size = soffset * o->derivative_factor;            /* akaze.cpp:341 */
                L->nsteps = 0;
            } else {
                L->esigma = soffset * powf(2, (float)j / S + i);     /* akaze.cpp:357 */
                float cur = 0.5f * L->esigma * L->esigma;
                float ttime = cur - last_etime;
                L->nsteps = orc_fed_tau(ttime, 1, tmax, o->reordering, L->tau, 128);
                L->size = L->esigma * o->derivative_factor / oratio; /* akaze.cpp:361 */
                last_etime = cur;
            }
            L->sigma_size = (int)(L->size + 0.5f);
            L->border = smax * L->sigma_size;
            size_t n = (size_t)L->pitch * L->h;
            L->lt = (float*)calloc(n, 4); L->det = (float*)calloc(n, 4);
            L->lx = (float*)calloc(n, 4); L->ly = (float*)calloc(n, 4);
        }
        float b0 = P->lev[i * S].border * oratio;                    /* akaze.cpp:434 */
        if (b0 < psz) psz = b0;
        oratio *= 2;
    }
    P->psz = (int)psz;
    size_t n0 = (size_t)P->lev[0].pitch * h;
    P->smooth = (float*)calloc(n0, 4); P->flow = (float*)calloc(n0, 4); P->tmp = (float*)calloc(n0, 4);
    P->resp = (float*)calloc((size_t)w * h, 4); P->layer = (int*)calloc((size_t)w * h, 4);
    return P;
}

void orc_pyramid_free(orc_pyramid* P)
{
    if (!P) return;
    for (int i = 0; i < P->nlev; i++) { free(P->lev[i].lt); free(P->lev[i].det); free(P->lev[i].lx); free(P->lev[i].ly); }
    free(P->smooth); free(P->flow); free(P->tmp); free(P->resp); free(P->layer); free(P);
}
int orc_pyramid_levels(const orc_pyramid* P) { return P->nlev; }
void orc_pyramid_level_dims(const orc_pyramid* P, int l, int* w, int* h, int* pitch, int* nsteps, float* size, int* sigma_size)
{
    const orc_level* L = &P->lev[l];
    if (w) *w = L->w; if (h) *h = L->h; if (pitch) *pitch = L->pitch; if (nsteps) *nsteps = L->nsteps;
    if (size) *size = L->size; if (sigma_size) *sigma_size = L->sigma_size;
}
int orc_pyramid_tau(const orc_pyramid* P, int l, float* tau, int cap)
{
    int n = P->lev[l].nsteps;
    for (int i = 0; i < n && i < cap; i++) tau[i] = P->lev[l].tau[i];
    return n;
}
const float* orc_pyramid_plane(const orc_pyramid* P, int l, int which)
{
    const orc_level* L = &P->lev[l];
    return which == 0 ? L->lt : which == 1 ? L->det : which == 2 ? L->lx : L->ly;
}
float orc_pyramid_kcontrast(const orc_pyramid* P) { return P->kcontrast; }

/* akaze.cpp:300-429 */
void orc_pyramid_build(orc_pyramid* P, const float* img, int ipitch)
{
#ifdef _OPENMP
    if (P->opt.threads > 0) omp_set_num_threads(P->opt.threads);
#endif
    int S = P->opt.max_scale;
    float kc = 0.f;
    for (int l = 0; l < P->nlev; l++) {
        orc_level* L = &P->lev[l];
        int w = L->w, h = L->h, p = L->pitch;
        size_t n = (size_t)p * h;
        if (l == 0) {
            /* akaze.cpp:325-346 */
            float var = P->opt.soffset * P->opt.soffset;
            int ksz = (int)(2 * ceilf((P->opt.soffset - 0.8f) / 0.3f) + 3);
            float* in = (float*)malloc(n * 4);
            for (int y = 0; y < h; y++) memcpy(in + (size_t)y * p, img + (size_t)y * ipitch, (size_t)w * 4);
            orc_lowpass(in, P->smooth, w, h, p, 1.f, 5);
            if (P->opt.kcontrast_override > 0.f) kc = P->opt.kcontrast_override;
            else { orc_scharr_mag(P->smooth, P->tmp, w, h, p); kc = orc_contrast_from_mag(P->tmp, w, h, p, P->opt.per, NULL); }
            P->kcontrast = kc;
            orc_lowpass(in, L->lt, w, h, p, var, ksz);
            free(in);
            orc_hessian(L->lt, L->lx, L->ly, L->det, L->sigma_size, w, h, p);
            continue;
        }
        if (L->sub == 0) {
            /* akaze.cpp:371-392: source is sublevel 0 of the previous octave */
            kc *= 0.75f;
            orc_level* Pv = &P->lev[l - S];
            orc_down_with_smooth(Pv->lt, L->lt, P->smooth, Pv->w, Pv->h, Pv->pitch, w, h, p);
            orc_flow(P->smooth, P->flow, P->opt.diffusivity, kc, w, h, p);
            for (int k = 0; k < L->nsteps; k++) {
                orc_nld_step(L->lt, P->flow, P->tmp, L->tau[k], w, h, p);
                memcpy(L->lt, P->tmp, n * 4);
            }
        } else {
            /* akaze.cpp:393-421 */
            orc_level* Pv = &P->lev[l - 1];
            orc_lowpass(Pv->lt, P->smooth, w, h, p, 1.f, 5);
            orc_flow(P->smooth, P->flow, P->opt.diffusivity, kc, w, h, p);
            orc_nld_step(Pv->lt, P->flow, L->lt, L->tau[0], w, h, p);
            for (int k = 1; k < L->nsteps; k++) {
                orc_nld_step(L->lt, P->flow, P->tmp, L->tau[k], w, h, p);
                memcpy(L->lt, P->tmp, n * 4);
            }
        }
        orc_hessian(P->smooth, L->lx, L->ly, L->det, L->sigma_size, w, h, p);   /* akaze.cpp:423 */
    }
}

/* akazed.cu:1334-1393 gCalcExtremaMap (+ hCalcExtremaMap :2563-2587), :1554-1613 gNmsRNaive,
 * :1615-1662 gRefine; deterministic merge (B-2) and raster order (B-5). */
int orc_pyramid_detect(orc_pyramid* P, orc_keypoint* out, int cap)
{
    int W = P->w, H = P->h, S = P->opt.max_scale;
    for (size_t i = 0; i < (size_t)W * H; i++) { P->resp[i] = -1.f; P->layer[i] = -1; }
    for (int l = 0; l < P->nlev; l++) {
        orc_level* L = &P->lev[l];
        int o = L->octave, w = L->w, h = L->h, p = L->pitch;
        int psz = (int)P->lev[o * S].border;                  /* hCalcExtremaMap: (int)params[0] */
        float border = L->border, thr = P->opt.dthreshold;
        int gx = (w - 2 * psz + 15) / 16, gy = (h - 2 * psz + 15) / 16;   /* the launch grid bounds the scan */
        for (int iy = psz; iy < psz + gy * 16; iy++)
            for (int ix = psz; ix < psz + gx * 16; ix++) {
                int left_x = (int)(ix - border + 0.5f) - 1, right_x = (int)(ix + border + 0.5f) + 1;
                int up_y = (int)(iy - border + 0.5f) - 1, down_y = (int)(iy + border + 0.5f) + 1;
                if (left_x < 0 || right_x >= w || up_y < 0 || down_y >= h) continue;
                const float* vp = L->det + (size_t)iy * p + ix;
                float v = *vp;
                if (v > thr && v > vp[-p] && v > vp[p] && v > vp[-1] && v > vp[1] &&
                    v > vp[-p - 1] && v > vp[-p + 1] && v > vp[p - 1] && v > vp[p + 1]) {
                    size_t oi = (size_t)(iy << o) * W + (ix << o);
                    if (P->resp[oi] < v) { P->resp[oi] = v; P->layer[oi] = l; }
                }
            }
    }
    int n = 0, psz = P->psz;
    for (int iy = psz; iy + psz < H; iy++)
        for (int ix = psz; ix + psz < W; ix++) {
            size_t idx = (size_t)iy * W + ix;
            int l = P->layer[idx];
            if (l < 0) continue;
            float fsz = P->lev[l].size, rc = P->resp[idx];
            int isz = (int)(fsz + 0.5f), sq = (int)(fsz * fsz);
            int kill = 0;
            for (int i = -isz; i <= isz && !kill; i++)
                for (int j = -isz; j <= isz; j++) {
                    if (i == 0 && j == 0) continue;
                    if (i * i + j * j >= sq) continue;
                    /* akazed.cu:1578-1581: the `continue` at the centre skips `new_idx++`, so in the centre row every
                     * j > 0 reads the pixel at offset j - 1 (the centre itself for j = 1) under the distance test of j;
                     * the farthest in-radius pixel to the right on the centre row is never examined. */
                    int jj = (i == 0 && j > 0) ? j - 1 : j;
                    size_t ni = (size_t)(iy + i) * W + (ix + jj);
                    if (P->layer[ni] < 0) continue;      /* non-candidates hold a value below any response */
                    float rn = P->resp[ni];
                    if (rn > rc || (rn == rc && i <= 0 && j <= 0)) { kill = 1; break; }
                }
            if (kill) continue;
            if (n < cap) {
                orc_keypoint* k = &out[n];
                memset(k, 0, sizeof(*k));
                k->ix = ix; k->iy = iy; k->x = (float)ix; k->y = (float)iy;
                k->layer = l; k->size = fsz; k->response = rc; k->angle = 0.f;
                /* refine, akazed.cu:1629-1657 */
                orc_level* L = &P->lev[l];
                int o = L->octave, p = L->pitch, x = ix >> o, y = iy >> o;
                const float* d = L->det + (size_t)y * p + x;
                float v2 = d[0] + d[0];
                float dx = 0.5f * (d[1] - d[-1]);
                float dy = 0.5f * (d[p] - d[-p]);
                float dxx = (d[1] + d[-1]) - v2;
                float dyy = (d[p] + d[-p]) - v2;
                float dxy = 0.25f * (((d[p + 1] + d[-p - 1]) - d[-p + 1]) - d[p - 1]);
                float dd = fmaf(dxx, dyy, -(dxy * dxy));
                float idd = dd != 0.f ? 1.f / dd : 0.f;
                float o0 = idd * fmaf(dy, dxy, -(dx * dyy));
                float o1 = idd * fmaf(dx, dxy, -(dy * dxx));
                int weak = o0 < -1.f || o0 > 1.f || o1 < -1.f || o1 > 1.f;
                if (!weak) {
                    float ratio = (float)(1 << o);
                    k->y = ratio * ((float)y + o1);
                    k->x = ratio * ((float)x + o0);
                }
            }
            n++;
        }
    return n < cap ? n : cap;
}

/* akazed.cu:173-185 dFastAtan2 */
static float fast_atan2(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float a = fminf(ax, ay) / fmaxf(ax, ay);
    float s = a * a;
    float r = fmaf(fmaf(fmaf(-0.0464964749f, s, 0.15931422f), s, -0.327622764f), s * a, a);
    r = (ay > ax ? 1.5707963267948966f - r : r);
    r = (x < 0 ? (float)(M_PI - r) : r);
    r = (y < 0 ? -r : r);
    return r;
}

void orc_compare_indices(int* c1, int* c2)
{
    /* akazed.cu:65-159: per grid (2x2 cells 0..3, 3x3 cells 4..12, 4x4 cells 13..28), per channel,
     * all pairs (j, i>j); value index = 3*cell + channel */
    int lo[3] = { 0, 4, 13 }, hi[3] = { 4, 13, 29 }, n = 0;
    for (int g = 0; g < 3; g++)
        for (int ch = 0; ch < 3; ch++)
            for (int j = lo[g]; j < hi[g] - 1; j++)
                for (int i = j + 1; i < hi[g]; i++) { c1[n] = 3 * j + ch; c2[n] = 3 * i + ch; n++; }
}

/* akazed.cu:1665-1736 gCalcOrient (bins summed in ascending thread order, B-4) and
 * akazed.cu:1869-2001 gDescribe2 (64 logical threads, exact accumulation + reduction tree) */
static void orient_one(const orc_pyramid* P, orc_keypoint* k)
{
    int S = P->opt.max_scale; (void)S;
    const orc_level* L = &P->lev[k->layer];
    int o = L->octave, p = L->pitch;
    int step = (int)(k->size + 0.5f);
    int x = (int)(k->x + 0.5f) >> o, y = (int)(k->y + 0.5f) >> o;
    float resx[42], resy[42];
    for (int i = 0; i < 42; i++) resx[i] = resy[i] = 0.f;
    for (int t = 0; t < 208; t++) {
        int i = (t & 15) - 6, j = (t / 16) - 6, r2 = i * i + j * j;
        if (r2 >= 36) continue;
        float gw = expf(-r2 * 0.08f);
        int yy = y + step * j, xx = x + step * i;
        if (yy < 0) yy = 0; if (yy >= L->h) yy = L->h - 1; if (xx < 0) xx = 0; if (xx >= L->w) xx = L->w - 1;
        size_t pos = (size_t)yy * p + xx;
        float dx = gw * L->lx[pos], dy = gw * L->ly[pos];
        float ang = atan2f(dy, dx);
        int a = (int)(ang * (21 / M_PI)) + 21;
        a = a > 41 ? 41 : a; a = a < 0 ? 0 : a;
        resx[a] += dx; resy[a] += dy;
    }
    float maxr = 0.f; int maxk = 0; float bx = 0, by = 0;
    float r8x[42], r8y[42];
    for (int t = 0; t < 42; t++) {
        float sx = resx[t], sy = resy[t];
        for (int kk = t + 1; kk < t + 7; kk++) { sx += resx[kk < 42 ? kk : kk - 42]; sy += resy[kk < 42 ? kk : kk - 42]; }
        r8x[t] = sx; r8y[t] = sy;
    }
    for (int t = 0; t < 42; t++) {
        float r = fmaf(r8x[t], r8x[t], r8y[t] * r8y[t]);
        if (r > maxr) { maxr = r; maxk = t; }
    }
    bx = r8x[maxk]; by = r8y[maxk];
    float ang = fast_atan2(by, bx);
    k->angle = (ang < 0.0f ? (float)(ang + 2.0f * M_PI) : ang);
}

static void describe_one(const orc_pyramid* P, orc_keypoint* k, const float* trig, const int* c1, const int* c2)
{
    const orc_level* L = &P->lev[k->layer];
    int o = L->octave, p = L->pitch;
    int pat = P->opt.pattern;
    int size2 = pat, size3 = (int)ceilf(2.0f * pat / 3.0f), size4 = (int)ceilf(0.5f * pat);
    int win = 3 * size3 > 4 * size4 ? 3 * size3 : 4 * size4;
    float iratio = 1.f / (1 << o);
    int scale = (int)(k->size + 0.5f);
    float xf = k->x * iratio, yf = k->y * iratio;
    float co = trig ? trig[0] : cosf(k->angle), si = trig ? trig[1] : sinf(k->angle);
    static __thread float acc[64][90];
    memset(acc, 0, sizeof(acc));
    for (int t = 0; t < 64; t++)
        for (int i = t; i < win * win; i += 64) {
            int y = i / win, x = i - win * y, m = x > y ? x : y;
            if (m >= win) continue;
            int l = x - size2, kk = y - size2;
            int xp = (int)(fmaf((float)scale, fmaf(co, (float)kk, -(si * (float)l)), xf) + 0.5f);
            int yp = (int)(fmaf((float)scale, fmaf(si, (float)kk, co * (float)l), yf) + 0.5f);
            if (yp < 0) yp = 0; if (yp >= L->h) yp = L->h - 1; if (xp < 0) xp = 0; if (xp >= L->w) xp = L->w - 1;
            size_t pos = (size_t)yp * p + xp;
            float im = L->lt[pos], dx = L->lx[pos], dy = L->ly[pos];
            float rx = fmaf(co, dy, -(si * dx));
            float ry = fmaf(co, dx, si * dy);
            if (m < 2 * size2) {
                int c = ((y < size2 ? 0 : 1) * 2 + (x < size2 ? 0 : 1));
                acc[t][3 * c] += im; acc[t][3 * c + 1] += rx; acc[t][3 * c + 2] += ry;
            }
            if (m < 3 * size3) {
                int x3 = x < size3 ? 0 : (x < 2 * size3 ? 1 : 2), y3 = y < size3 ? 0 : (y < 2 * size3 ? 1 : 2);
                int c = 4 + y3 * 3 + x3;
                acc[t][3 * c] += im; acc[t][3 * c + 1] += rx; acc[t][3 * c + 2] += ry;
            }
            if (m < 4 * size4) {
                int x4 = x < 2 * size4 ? (x < size4 ? 0 : 1) : (x < 3 * size4 ? 2 : 3);
                int y4 = y < 2 * size4 ? (y < size4 ? 0 : 1) : (y < 3 * size4 ? 2 : 3);
                int c = 13 + y4 * 4 + x4;
                acc[t][3 * c] += im; acc[t][3 * c + 1] += rx; acc[t][3 * c + 2] += ry;
            }
        }
    float val[90];
    for (int v = 0; v < 90; v++) {
        float a[32];
        for (int t = 0; t < 32; t++) a[t] = acc[t][v] + acc[t + 32][v];
        for (int d = 1; d < 32; d <<= 1)
            for (int t = 0; t + d < 32; t += 2 * d) a[t] = a[t] + a[t + d];   /* lane-0 cone of the shuffle-down tree */
        val[v] = a[0];
    }
    memset(k->desc, 0, 64);
    for (int b = 0; b < 61; b++) {
        unsigned char r = 0;
        for (int i = 0; i < (b == 60 ? 6 : 8); i++) r |= (unsigned char)((val[c1[8 * b + i]] > val[c2[8 * b + i]] ? 1 : 0) << i);
        k->desc[b] = r;
    }
}

void orc_pyramid_describe(const orc_pyramid* P, orc_keypoint* kps, int n, int with_orientation, const float* trig)
{
    int c1[488], c2[488];
    orc_compare_indices(c1, c2);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; i++) {
        if (with_orientation) orient_one(P, &kps[i]);
        describe_one(P, &kps[i], trig ? trig + 2 * i : NULL, c1, c2);
    }
}

int orc_detect_and_compute(const float* img, int w, int h, int pitch, const orc_options* o,
                           orc_keypoint* out, int cap, int describe)
{
    orc_pyramid* P = orc_pyramid_create(w, h, o);
    orc_pyramid_build(P, img, pitch);
    int n = orc_pyramid_detect(P, out, cap);
    if (describe) orc_pyramid_describe(P, out, n, 1, NULL);
    orc_pyramid_free(P);
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * Matching
 * ------------------------------------------------------------------------------------------- */
static inline int hamming64(const unsigned char* a, const unsigned char* b)
{
    const unsigned long long* x = (const unsigned long long*)a;
    const unsigned long long* y = (const unsigned long long*)b;
    int d = 0;
    for (int i = 0; i < 8; i++) d += __builtin_popcountll(x[i] ^ y[i]);
    return d;
}

/* akazed.cu:2144-2241 gHammingMatch: 16 strided partial minima (first index on ties inside a
 * stride); accept iff the global minimum is strictly unique across the strides and < 96 */
void orc_match_compat(const unsigned char* q, int nq, const unsigned char* t, int nt, int* out2)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        int dmin[16], imin[16];
        for (int c = 0; c < 16; c++) { dmin[c] = 1 << 20; imin[c] = -1; }
        for (int j = 0; j < nt; j++) {
            int d = hamming64(q + (size_t)64 * i, t + (size_t)64 * j), c = j & 15;
            if (d < dmin[c]) { dmin[c] = d; imin[c] = j; }
        }
        int best = 0;
        for (int c = 1; c < 16; c++) if (dmin[c] < dmin[best]) best = c;
        int unique = 1;
        for (int c = 0; c < 16; c++) if (c != best && !(dmin[best] < dmin[c])) unique = 0;
        if (unique && dmin[best] < ORC_MAX_DIST && imin[best] >= 0) { out2[2 * i] = imin[best]; out2[2 * i + 1] = dmin[best]; }
        else { out2[2 * i] = -1; out2[2 * i + 1] = -1; }
    }
}

void orc_match_knn2(const unsigned char* q, int nq, const unsigned char* t, int nt, int* out4)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        int d1 = 1 << 20, i1 = -1, d2 = 1 << 20, i2 = -1;
        for (int j = 0; j < nt; j++) {
            int d = hamming64(q + (size_t)64 * i, t + (size_t)64 * j);
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
            else if (d < d2) { d2 = d; i2 = j; }
        }
        out4[4 * i] = i1; out4[4 * i + 1] = i1 < 0 ? -1 : d1; out4[4 * i + 2] = i2; out4[4 * i + 3] = i2 < 0 ? -1 : d2;
    }
}
