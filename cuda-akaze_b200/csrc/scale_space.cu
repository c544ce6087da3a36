// Stage kernels of the nonlinear scale space: one kernel per reference stage function, batched over
// frames (blockIdx.z) and free of host round trips.  These are the seams the parity tests compare
// against the compiled reference (akazed.h h* functions); the production path in
// scale_space_fused.cu fuses them and is checked bit-for-bit against these.
#include "common.cuh"
#include "kernels.h"

using namespace akz;

namespace {

constexpr int BX = 32, BY = 8;

struct TapsArg { float k[6]; };

__device__ __forceinline__ float load_px(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_px(const unsigned char* p) { return u8_to_unit(__ldg(p)); }

// ---- Gaussian blur, reference gConv2d<R> (akazed.cu:204-290) ---------------------------------------
// tile 32x16 outputs; the row-filtered intermediate (rounded to f32, as the reference's smem copy is)
// is staged in shared memory for rows y0-R .. y0+15+R with reflect-101 applied to the ROW INDEX.
template <int R, typename Tin>
__global__ void __launch_bounds__(256) k_lowpass(const Tin* __restrict__ src, float* __restrict__ dst, int w, int h,
                                                 int sp, long long sstride, int dp, long long dstride, TapsArg t)
{
    constexpr int TH = 16;
    __shared__ float rows[TH + 2 * R][BX];
    const Tin* s = src + (long long)blockIdx.z * sstride;
    float* d = dst + (long long)blockIdx.z * dstride;
    int x = blockIdx.x * BX + threadIdx.x;
    int y0 = blockIdx.y * TH;
    int xc = min(x, w - 1);
    for (int r = threadIdx.y; r < TH + 2 * R; r += BY) {
        int yy = refl(y0 - R + r, h);
        yy = min(max(yy, 0), h - 1);
        const Tin* row = s + (long long)yy * sp;
        float acc = __fmul_rn(__fadd_rn(load_px(row + refl_lo(xc - 1)), load_px(row + refl_hi(xc + 1, w))), t.k[1]);
        acc = __fmaf_rn(load_px(row + xc), t.k[0], acc);
#pragma unroll
        for (int i = 2; i <= R; i++)
            acc = __fmaf_rn(__fadd_rn(load_px(row + refl_lo(xc - i)), load_px(row + refl_hi(xc + i, w))), t.k[i], acc);
        rows[r][threadIdx.x] = acc;
    }
    __syncthreads();
    if (x >= w) return;
    for (int r = threadIdx.y; r < TH; r += BY) {
        int y = y0 + r;
        if (y >= h) break;
        int c = r + R;
        float acc = __fmul_rn(__fadd_rn(rows[c - 1][threadIdx.x], rows[c + 1][threadIdx.x]), t.k[1]);
        acc = __fmaf_rn(rows[c][threadIdx.x], t.k[0], acc);
#pragma unroll
        for (int i = 2; i <= R; i++)
            acc = __fmaf_rn(__fadd_rn(rows[c - i][threadIdx.x], rows[c + i][threadIdx.x]), t.k[i], acc);
        d[(long long)y * dp + x] = acc;
    }
}

// ---- octave transition, reference gDownWithSmooth (akazed.cu:449-511) -------------------------------
__global__ void __launch_bounds__(256) k_down_smooth(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ smooth,
                                                     int sw, int sh, int sp, long long sstride,
                                                     int dw, int dh, int dp, long long dstride, TapsArg t)
{
    constexpr int TH = 16, R = 2;
    __shared__ float rows[TH + 2 * R][BX];
    const float* s = src + (long long)blockIdx.z * sstride;
    float* d = dst + (long long)blockIdx.z * dstride;
    float* sm = smooth + (long long)blockIdx.z * dstride;
    int x = blockIdx.x * BX + threadIdx.x;
    int y0 = blockIdx.y * TH;
    int xc = min(x, dw - 1);
    int sx = xc + xc;
    for (int r = threadIdx.y; r < TH + 2 * R; r += BY) {
        int sy = refl(2 * (y0 - R + r), sh);       // reflection in SOURCE coordinates
        sy = min(max(sy, 0), sh - 1);
        const float* row = s + (long long)sy * sp;
        rows[r][threadIdx.x] = gauss_r2(__ldg(row + refl_lo(sx - 4)), __ldg(row + refl_lo(sx - 2)), __ldg(row + sx),
                                        __ldg(row + refl_hi(sx + 2, sw)), __ldg(row + refl_hi(sx + 4, sw)), t.k[0], t.k[1], t.k[2]);
    }
    __syncthreads();
    if (x >= dw) return;
    for (int r = threadIdx.y; r < TH; r += BY) {
        int y = y0 + r;
        if (y >= dh) break;
        int c = r + R;
        long long o = (long long)y * dp + x;
        d[o] = __ldg(s + (long long)(2 * y) * sp + sx);
        sm[o] = gauss_r2(rows[c - 2][threadIdx.x], rows[c - 1][threadIdx.x], rows[c][threadIdx.x],
                         rows[c + 1][threadIdx.x], rows[c + 2][threadIdx.x], t.k[0], t.k[1], t.k[2]);
    }
}

// 3x3 neighbourhood at distance s with reflect-101
struct Nb9 { float ul, uc, ur, cl, cc, cr, ll, lc, lr; };
__device__ __forceinline__ Nb9 load9(const float* __restrict__ p, int x, int y, int w, int h, int pitch, int s)
{
    int x0 = refl_lo(x - s), x2 = refl_hi(x + s, w);
    int y0 = refl_lo(y - s), y2 = refl_hi(y + s, h);
    const float* r0 = p + (long long)y0 * pitch;
    const float* r1 = p + (long long)y * pitch;
    const float* r2 = p + (long long)y2 * pitch;
    Nb9 n;
    n.ul = __ldg(r0 + x0); n.uc = __ldg(r0 + x); n.ur = __ldg(r0 + x2);
    n.cl = __ldg(r1 + x0); n.cc = __ldg(r1 + x); n.cr = __ldg(r1 + x2);
    n.ll = __ldg(r2 + x0); n.lc = __ldg(r2 + x); n.lr = __ldg(r2 + x2);
    return n;
}

// ---- contrast factor, reference hScharrContrast (akazed.cu:2410-2484) -------------------------------
// pass 1: true maximum of the Scharr magnitude (App. B-1) -> hmax_bits[frame] (seeded with 0.03f)
__global__ void __launch_bounds__(256) k_scharr_max(const float* __restrict__ src, unsigned* __restrict__ hmax_bits,
                                                    int w, int h, int pitch, long long stride)
{
    const float* s = src + (long long)blockIdx.z * stride;
    int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
    float m = 0.f;
    if (x < w && y < h) {
        Nb9 n = load9(s, x, y, w, h, pitch, 1);
        float dx = scharr_dx(n.ul, n.ur, n.cl, n.cr, n.ll, n.lr);
        float dy = scharr_dy(n.ul, n.uc, n.ur, n.ll, n.lc, n.lr);
        m = __fsqrt_rn(grad_sq(dx, dy));
    }
    unsigned b = __float_as_uint(m);                 // m >= 0: the bit patterns order like the floats
    b = __reduce_max_sync(0xffffffffu, b);
    __shared__ unsigned smax[BY];
    if (threadIdx.x == 0) smax[threadIdx.y] = b;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned v = smax[0];
#pragma unroll
        for (int i = 1; i < BY; i++) v = max(v, smax[i]);
        atomicMax(hmax_bits + blockIdx.z, v);
    }
}

// pass 2: 300-bin histogram of mag*300/hmax (truncating multiply), in-image pixels only (App. B-3)
__global__ void __launch_bounds__(256) k_contrast_hist(const float* __restrict__ src, const unsigned* __restrict__ hmax_bits,
                                                       int* __restrict__ hist, int w, int h, int pitch, long long stride)
{
    __shared__ int sh[AKZ_NBINS];
    int tid = threadIdx.y * BX + threadIdx.x;
    for (int i = tid; i < AKZ_NBINS; i += BX * BY) sh[i] = 0;
    __syncthreads();
    const float* s = src + (long long)blockIdx.z * stride;
    float hfactor = __fdiv_rn((float)AKZ_NBINS, __uint_as_float(hmax_bits[blockIdx.z]));
    int x = blockIdx.x * BX + threadIdx.x;
    // each block covers 4 row groups to amortise the shared histogram
    for (int yy = 0; yy < 4; yy++) {
        int y = (blockIdx.y * 4 + yy) * BY + threadIdx.y;
        if (x < w && y < h) {
            Nb9 n = load9(s, x, y, w, h, pitch, 1);
            float dx = scharr_dx(n.ul, n.ur, n.cl, n.cr, n.ll, n.lr);
            float dy = scharr_dy(n.ul, n.uc, n.ur, n.ll, n.lc, n.lr);
            float m = __fsqrt_rn(grad_sq(dx, dy));
            int hi = (int)__fmul_rz(m, hfactor);
            hi = min(hi, AKZ_NBINS - 1);
            atomicAdd(&sh[hi], 1);
        }
    }
    __syncthreads();
    int* g = hist + (long long)blockIdx.z * AKZ_NBINS;
    for (int i = tid; i < AKZ_NBINS; i += BX * BY)
        if (sh[i]) atomicAdd(g + i, sh[i]);
}

// pass 3: the host scan of akazed.cu:2468-2481, one thread per frame
__global__ void k_contrast_scan(const int* __restrict__ hist, const unsigned* __restrict__ hmax_bits, float* __restrict__ kout,
                                float per, int w, int h, int nframes, float override_k)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    if (override_k > 0.f) { kout[f] = override_k; return; }
    const int* hg = hist + (long long)f * AKZ_NBINS;
    float hfactor = __fdiv_rn((float)AKZ_NBINS, __uint_as_float(hmax_bits[f]));
    int thresh = (int)__fmul_rn((float)(w * h - hg[0]), per);
    int cum = 0, k = 1;
    while (k < AKZ_NBINS) {
        if (cum >= thresh) break;
        cum += hg[k];
        k++;
    }
    kout[f] = __fdiv_rn((float)k, hfactor);
}

__global__ void k_contrast_init(unsigned* hmax_bits, int* hist, int nframes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes) hmax_bits[i] = __float_as_uint(0.03f);       // akazed.cu:2413
    if (i < nframes * AKZ_NBINS) hist[i] = 0;
}

// ---- conductance, reference gFlowNaive (akazed.cu:1068-1107) ----------------------------------------
__global__ void __launch_bounds__(256) k_flow(const float* __restrict__ src, float* __restrict__ flow, int type,
                                              const float* __restrict__ kc, float kscale, int nmul,
                                              int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
    if (x >= w || y >= h) return;
    const float* s = src + (long long)blockIdx.z * stride;
    float k = kc[blockIdx.z];
    for (int i = 0; i < nmul; i++) k = __fmul_rn(k, kscale);      // kcontrast *= 0.75f per octave (akaze.cpp:373)
    float ikc = __fdiv_rn(1.f, __fmul_rn(k, k));                  // akazed.cu:2493
    Nb9 n = load9(s, x, y, w, h, pitch, 1);
    float dx = scharr_dx(n.ul, n.ur, n.cl, n.cr, n.ll, n.lr);
    float dy = scharr_dy(n.ul, n.uc, n.ur, n.ll, n.lc, n.lr);
    float d = __fmul_rn(grad_sq(dx, dy), ikc);
    flow[(long long)blockIdx.z * stride + (long long)y * pitch + x] = conductance(type, d);
}

// ---- one explicit step, reference gNldStepNaive (akazed.cu:1241-1264) ---------------------------------
__global__ void __launch_bounds__(256) k_nld_step(const float* __restrict__ src, const float* __restrict__ flow, float* __restrict__ dst,
                                                  float stepfac, int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
    if (x >= w || y >= h) return;
    long long base = (long long)blockIdx.z * stride;
    const float* L = src + base;
    const float* g = flow + base;
    int x0 = refl_lo(x - 1), x2 = refl_hi(x + 1, w), y0 = refl_lo(y - 1), y2 = refl_hi(y + 1, h);
    long long r1 = (long long)y * pitch, r0 = (long long)y0 * pitch, r2 = (long long)y2 * pitch;
    float L0 = __ldg(L + r1 + x), g0 = __ldg(g + r1 + x);
    dst[base + r1 + x] = nld_update(L0, g0, __ldg(L + r1 + x0), __ldg(g + r1 + x0), __ldg(L + r1 + x2), __ldg(g + r1 + x2),
                                    __ldg(L + r2 + x), __ldg(g + r2 + x), __ldg(L + r0 + x), __ldg(g + r0 + x), stepfac);
}

// ---- derivatives, reference gDerivate / gHessianDeterminant (akazed.cu:1267-1331) ----------------------
__global__ void __launch_bounds__(256) k_derivs(const float* __restrict__ src, float* __restrict__ lx, float* __restrict__ ly,
                                                int step, float fac1, float fac2, int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
    if (x >= w || y >= h) return;
    long long base = (long long)blockIdx.z * stride;
    Nb9 n = load9(src + base, x, y, w, h, pitch, step);
    long long o = base + (long long)y * pitch + x;
    lx[o] = deriv1(sum_x(n.ul, n.ur, n.ll, n.lr), __fsub_rn(n.cr, n.cl), fac1, fac2);
    ly[o] = deriv1(sum_y(n.ul, n.ur, n.ll, n.lr), __fsub_rn(n.lc, n.uc), fac1, fac2);
}

__global__ void __launch_bounds__(256) k_hessian_det(const float* __restrict__ lx, const float* __restrict__ ly, float* __restrict__ det,
                                                     int step, float fac1, float fac2, int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
    if (x >= w || y >= h) return;
    long long base = (long long)blockIdx.z * stride;
    Nb9 a = load9(lx + base, x, y, w, h, pitch, step);
    Nb9 b = load9(ly + base, x, y, w, h, pitch, step);
    float dxx = deriv2(sum_x(a.ul, a.ur, a.ll, a.lr), __fsub_rn(a.cr, a.cl), fac1, fac2);
    float dxy = deriv2(sum_y(a.ul, a.ur, a.ll, a.lr), __fsub_rn(a.lc, a.uc), fac1, fac2);
    float dyy = deriv2(sum_y(b.ul, b.ur, b.ll, b.lr), __fsub_rn(b.lc, b.uc), fac1, fac2);
    det[base + (long long)y * pitch + x] = hess_det(dxx, dyy, dxy);
}

TapsArg make_taps(float var, int R)
{
    TapsArg t = {};
    akz_gauss_taps(var, R, t.k);
    return t;
}

inline dim3 grid2d(int w, int h, int th, int n) { return dim3((w + BX - 1) / BX, (h + th - 1) / th, n); }

}  // namespace

// ---- launch wrappers -------------------------------------------------------------------------------
namespace akzk {

int radius_from_ksz(int ksz) { return ksz <= 5 ? 2 : ksz <= 7 ? 3 : ksz <= 9 ? 4 : ksz <= 11 ? 5 : -1; }   // akazed.cu:2347-2381

template <typename Tin>
static int lowpass_t(cudaStream_t st, const Tin* src, float* dst, int w, int h, int sp, long long sstride, int dp, long long dstride,
                     int n, float var, int ksz)
{
    int R = radius_from_ksz(ksz);
    if (R < 0) return akz_set_error(AKZ_E_UNSUPPORTED, "Gaussian kernels larger than 11 are not implemented (akazed.cu:2377)");
    TapsArg t = make_taps(var, R);
    dim3 g = grid2d(w, h, 16, n), b(BX, BY);
    switch (R) {
    case 2: k_lowpass<2, Tin><<<g, b, 0, st>>>(src, dst, w, h, sp, sstride, dp, dstride, t); break;
    case 3: k_lowpass<3, Tin><<<g, b, 0, st>>>(src, dst, w, h, sp, sstride, dp, dstride, t); break;
    case 4: k_lowpass<4, Tin><<<g, b, 0, st>>>(src, dst, w, h, sp, sstride, dp, dstride, t); break;
    default: k_lowpass<5, Tin><<<g, b, 0, st>>>(src, dst, w, h, sp, sstride, dp, dstride, t); break;
    }
    return 1;
}

int lowpass(cudaStream_t st, const float* src, float* dst, int w, int h, int sp, long long sstride, int dp, long long dstride,
            int n, float var, int ksz)
{ return lowpass_t<float>(st, src, dst, w, h, sp, sstride, dp, dstride, n, var, ksz); }

int lowpass_u8(cudaStream_t st, const unsigned char* src, float* dst, int w, int h, int sp, long long sstride, int dp, long long dstride,
               int n, float var, int ksz)
{ return lowpass_t<unsigned char>(st, src, dst, w, h, sp, sstride, dp, dstride, n, var, ksz); }

int down_with_smooth(cudaStream_t st, const float* src, float* dst, float* smooth, int sw, int sh, int sp, long long sstride,
                     int dw, int dh, int dp, long long dstride, int n)
{
    TapsArg t = make_taps(1.f, 2);
    k_down_smooth<<<grid2d(dw, dh, 16, n), dim3(BX, BY), 0, st>>>(src, dst, smooth, sw, sh, sp, sstride, dw, dh, dp, dstride, t);
    return 1;
}

int contrast(cudaStream_t st, const float* src, unsigned* hmax_bits, int* hist, float* kout, float per, float override_k,
             int w, int h, int pitch, long long stride, int n)
{
    int tot = n * AKZ_NBINS;
    k_contrast_init<<<(tot + 255) / 256, 256, 0, st>>>(hmax_bits, hist, n);
    int launches = 2;
    if (!(override_k > 0.f)) {
        k_scharr_max<<<grid2d(w, h, BY, n), dim3(BX, BY), 0, st>>>(src, hmax_bits, w, h, pitch, stride);
        k_contrast_hist<<<grid2d(w, h, BY * 4, n), dim3(BX, BY), 0, st>>>(src, hmax_bits, hist, w, h, pitch, stride);
        launches += 2;
    }
    k_contrast_scan<<<(n + 63) / 64, 64, 0, st>>>(hist, hmax_bits, kout, per, w, h, n, override_k);
    return launches;
}

// only the percentile scan (the histogram and the maximum were produced by base_level2)
int contrast_scan(cudaStream_t st, const unsigned* hmax_bits, const int* hist, float* kout, float per, float override_k, int w, int h, int n)
{
    k_contrast_scan<<<(n + 63) / 64, 64, 0, st>>>(hist, hmax_bits, kout, per, w, h, n, override_k);
    return 1;
}

int flow(cudaStream_t st, const float* src, float* flowp, int type, const float* kc, float kscale, int nmul,
         int w, int h, int pitch, long long stride, int n)
{
    k_flow<<<grid2d(w, h, BY, n), dim3(BX, BY), 0, st>>>(src, flowp, type, kc, kscale, nmul, w, h, pitch, stride);
    return 1;
}

int nld_step(cudaStream_t st, const float* src, const float* flowp, float* dst, float tau, int w, int h, int pitch, long long stride, int n)
{
    float stepfac = 0.5f * tau;                                   // akazed.cu:2515
    k_nld_step<<<grid2d(w, h, BY, n), dim3(BX, BY), 0, st>>>(src, flowp, dst, stepfac, w, h, pitch, stride);
    return 1;
}

void hessian_factors(float* fac1, float* fac2)
{
    float wgt = 10.f / 3.f;                                       // akazed.cu:2537-2539
    *fac1 = 1.f / (2.f * (wgt + 2.f));
    *fac2 = wgt * *fac1;
}

int hessian(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long stride, int n)
{
    float fac1, fac2;
    hessian_factors(&fac1, &fac2);
    dim3 g = grid2d(w, h, BY, n), b(BX, BY);
    k_derivs<<<g, b, 0, st>>>(smooth, lx, ly, step, fac1, fac2, w, h, pitch, stride);
    k_hessian_det<<<g, b, 0, st>>>(lx, ly, det, step, fac1, fac2, w, h, pitch, stride);
    return 2;
}

}  // namespace akzk
