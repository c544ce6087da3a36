// Integer ("fast") scale space: the reference's second pipeline (namespace fastakaze, akazed.cu:2781-4366, driven by
// Akazer::fastDetect akaze.cpp:506-743).  Pixels stay integers in the 0..255 range of the 8-bit input, filter weights are
// 16.16 fixed point ((int)(w * 65536 + 0.5f)) and every product sum is shifted right by 16; the conductance is an integer
// scaled by 65536.  Integer addition is associative, so unlike the float pipeline nothing here depends on the order of
// summation: one straightforward kernel per stage, batched over the frames of a chunk (blockIdx.z), reflect-101 borders
// applied on indices.  The few float expressions (gradient magnitude, conductance) are written in the reference's own
// form so that nvcc contracts them the same way (checked against the compiled reference in tests/test_gpu_fast.py).
#include "common.cuh"
#include "kernels.h"

using namespace akz;

namespace {

constexpr int FBX = 32, FBY = 8;

struct ITaps { int k[6]; };

inline dim3 fgrid(int w, int h, int n) { return dim3((w + FBX - 1) / FBX, (h + FBY - 1) / FBY, n); }

__device__ __forceinline__ int ldpx(const unsigned char* p) { return (int)__ldg(p); }
__device__ __forceinline__ int ldpx(const int* p) { return __ldg(p); }

// row pass of the separable Gaussian (gConv2d<R> / gConv2dR2, akazed.cu:2786-3076): (k0*x0 + sum ki*(x-i + x+i)) >> 16
template <int R, typename Tin>
__global__ void __launch_bounds__(256) k_frow(const Tin* __restrict__ src, int* __restrict__ dst, int w, int h, int sp, long long sstride,
                                              int dp, long long dstride, ITaps t)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    const Tin* row = src + (long long)blockIdx.z * sstride + (long long)y * sp;
    int s = t.k[0] * ldpx(row + x);
#pragma unroll
    for (int i = 1; i <= R; i++) s += t.k[i] * (ldpx(row + refl_lo(x - i)) + ldpx(row + refl_hi(x + i, w)));
    dst[(long long)blockIdx.z * dstride + (long long)y * dp + x] = s >> 16;
}

// column pass over row-filtered data; rows by reflected index (the reference stages reflected rows in shared memory)
template <int R>
__global__ void __launch_bounds__(256) k_fcol(const int* __restrict__ src, int* __restrict__ dst, int w, int h, int p, long long stride, ITaps t)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    const int* s0 = src + (long long)blockIdx.z * stride + x;
    int s = t.k[0] * __ldg(s0 + (long long)y * p);
#pragma unroll
    for (int i = 1; i <= R; i++) s += t.k[i] * (__ldg(s0 + (long long)refl_lo(y - i) * p) + __ldg(s0 + (long long)refl_hi(y + i, h) * p));
    dst[(long long)blockIdx.z * stride + (long long)y * p + x] = s >> 16;
}

// octave transition (gDownWithSmooth akazed.cu:3143-3205): dst = src(2x, 2y); smooth = radius-2 blur on the coarse lattice
// with taps at source offsets 0, +-2, +-4 reflected in SOURCE coordinates
__device__ __forceinline__ int srefl(int c, int m) { c = c < 0 ? -c : c; return c < m ? c : m + m - 2 - c; }
__global__ void __launch_bounds__(256) k_fdown(const int* __restrict__ src, int* __restrict__ dst, int* __restrict__ smooth,
                                               int sw, int sh, int sp, long long sstride, int dw, int dh, int dp, long long dstride, ITaps t)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int* s = src + (long long)blockIdx.z * sstride;
    const int sx = 2 * x, sy = 2 * y;
    const int xs[5] = { srefl(sx - 4, sw), srefl(sx - 2, sw), sx, srefl(sx + 2, sw), srefl(sx + 4, sw) };
    const int ys[5] = { srefl(sy - 4, sh), srefl(sy - 2, sh), sy, srefl(sy + 2, sh), srefl(sy + 4, sh) };
    int rv[5];
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int* row = s + (long long)ys[r] * sp;
        rv[r] = (t.k[0] * __ldg(row + xs[2]) + t.k[1] * (__ldg(row + xs[1]) + __ldg(row + xs[3])) + t.k[2] * (__ldg(row + xs[0]) + __ldg(row + xs[4]))) >> 16;
    }
    long long o = (long long)blockIdx.z * dstride + (long long)y * dp + x;
    dst[o] = __ldg(s + (long long)sy * sp + sx);
    smooth[o] = (t.k[0] * rv[2] + t.k[1] * (rv[1] + rv[3]) + t.k[2] * (rv[0] + rv[4])) >> 16;
}

struct INb { int ul, uc, ur, cl, cr, ll, lc, lr; };
__device__ __forceinline__ INb inb(const int* __restrict__ p, int x, int y, int w, int h, int pitch, int step)
{
    int x0 = refl_lo(x - step), x2 = refl_hi(x + step, w), y0 = refl_lo(y - step), y2 = refl_hi(y + step, h);
    const int* r0 = p + (long long)y0 * pitch;
    const int* r1 = p + (long long)y * pitch;
    const int* r2 = p + (long long)y2 * pitch;
    INb n;
    n.ul = __ldg(r0 + x0); n.uc = __ldg(r0 + x); n.ur = __ldg(r0 + x2);
    n.cl = __ldg(r1 + x0); n.cr = __ldg(r1 + x2);
    n.ll = __ldg(r2 + x0); n.lc = __ldg(r2 + x); n.lr = __ldg(r2 + x2);
    return n;
}

// gradient magnitude (gScharrContrastNaive akazed.cu:3208-3231) and its maximum (true maximum, App. B-1)
__global__ void __launch_bounds__(256) k_fscharr(const int* __restrict__ src, int* __restrict__ mag, int* __restrict__ hmax,
                                                 int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    int m = 0;
    if (x < w && y < h) {
        INb n = inb(src + (long long)blockIdx.z * stride, x, y, w, h, pitch, 1);
        int dx = 10 * (n.cr - n.cl) + 3 * (n.ur + n.lr - n.ul - n.ll);
        int dy = 10 * (n.lc - n.uc) + 3 * (n.ll + n.lr - n.ul - n.ur);
        m = (int)(__fsqrt_rn(dx * dx + dy * dy) + 0.5f);
        mag[(long long)blockIdx.z * stride + (long long)y * pitch + x] = m;
    }
    m = __reduce_max_sync(0xffffffffu, m);
    __shared__ int sm[FBY];
    if (threadIdx.x == 0) sm[threadIdx.y] = m;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int v = sm[0];
#pragma unroll
        for (int i = 1; i < FBY; i++) v = max(v, sm[i]);
        atomicMax(hmax + blockIdx.z, v);
    }
}

// hfactor = (int)(NBINS / (float)hmax * 65536 + 0.5f), host arithmetic of akazed.cu:4137 (no contraction on the host)
__device__ __forceinline__ int fast_hfactor(int hmax) { return (int)__fadd_rn(__fmul_rn(__fdiv_rn((float)AKZ_NBINS, (float)hmax), 65536.f), 0.5f); }

__global__ void __launch_bounds__(256) k_fhist(const int* __restrict__ mag, const int* __restrict__ hmax, int* __restrict__ hist,
                                               int w, int h, int pitch, long long stride)
{
    __shared__ int sh[AKZ_NBINS];
    int tid = threadIdx.y * FBX + threadIdx.x;
    for (int i = tid; i < AKZ_NBINS; i += FBX * FBY) sh[i] = 0;
    __syncthreads();
    const int factor = fast_hfactor(hmax[blockIdx.z]);
    int x = blockIdx.x * FBX + threadIdx.x;
    for (int yy = 0; yy < 4; yy++) {
        int y = (blockIdx.y * 4 + yy) * FBY + threadIdx.y;
        if (x < w && y < h) {
            int hi = (__ldg(mag + (long long)blockIdx.z * stride + (long long)y * pitch + x) * factor) >> 16;      // akazed.cu:3321
            atomicAdd(&sh[min(hi, AKZ_NBINS - 1)], 1);
        }
    }
    __syncthreads();
    int* g = hist + (long long)blockIdx.z * AKZ_NBINS;
    for (int i = tid; i < AKZ_NBINS; i += FBX * FBY)
        if (sh[i]) atomicAdd(g + i, sh[i]);
}

// host scan of akazed.cu:4152-4166 on the device: kcontrast = k * hmax / NBINS (integers)
__global__ void k_fscan(const int* __restrict__ hist, const int* __restrict__ hmax, int* __restrict__ kout, float per, int w, int h, int nframes, int override_k)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    if (override_k > 0) { kout[f] = override_k; return; }
    const int* hg = hist + (long long)f * AKZ_NBINS;
    int thresh = (int)__fmul_rn((float)(w * h - hg[0]), per);
    int cum = 0, k = 1;
    while (k < AKZ_NBINS) {
        if (cum >= thresh) break;
        cum += hg[k];
        k++;
    }
    kout[f] = k * hmax[f] / AKZ_NBINS;
}

__global__ void k_finit(int* hmax, int* hist, int nframes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes) hmax[i] = 1;                                  // akazed.cu:4099
    if (i < nframes * AKZ_NBINS) hist[i] = 0;
}

// conductance (gFlowNaive akazed.cu:3406-3446); kc: per-frame integer contrast factor, scaled per octave on the device as
// akaze.cpp:649 does on the host: k = (int)(k * 0.75f + 0.5f)
__global__ void __launch_bounds__(256) k_fflow(const int* __restrict__ src, int* __restrict__ flow, int type, const int* __restrict__ kc, int nmul,
                                               int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    int k = kc[blockIdx.z];
    for (int i = 0; i < nmul; i++) k = (int)__fadd_rn(__fmul_rn((float)k, 0.75f), 0.5f);
    const float ikc = __fdiv_rn(1.f, (float)(k * k));              // akazed.cu:4218 (host)
    INb n = inb(src + (long long)blockIdx.z * stride, x, y, w, h, pitch, 1);
    int dx = 10 * (n.cr - n.cl) + 3 * (n.ur + n.lr - n.ul - n.ll);
    int dy = 10 * (n.lc - n.uc) + 3 * (n.ll + n.lr - n.ul - n.ur);
    float dif2 = (dx * dx + dy * dy) * ikc;
    int g;
    if (type == 0) g = (int)(__expf(-dif2) * 65536 + 0.5f);
    else if (type == 1) g = (int)(1.f / (1.f + dif2) * 65536 + 0.5f);
    else if (type == 2) g = (int)((1.f - __expf(-3.315f / __powf(dif2, 4))) * 65536 + 0.5f);
    else g = (int)(1.f / __fsqrt_rn(1.f + dif2) * 65536 + 0.5f);
    flow[(long long)blockIdx.z * stride + (long long)y * pitch + x] = g;
}

// explicit diffusion step (gNldStepNaive akazed.cu:3449-3473)
__global__ void __launch_bounds__(256) k_fnld(const int* __restrict__ src, const int* __restrict__ flow, int* __restrict__ dst, int stepfac,
                                              int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    const int* L = src + (long long)blockIdx.z * stride;
    const int* G = flow + (long long)blockIdx.z * stride;
    int x0 = refl_lo(x - 1), x2 = refl_hi(x + 1, w), y0 = refl_lo(y - 1), y2 = refl_hi(y + 1, h);
    long long c = (long long)y * pitch + x;
    int L0 = __ldg(L + c), g0 = __ldg(G + c);
    long long r = (long long)y * pitch + x2, l = (long long)y * pitch + x0, d = (long long)y2 * pitch + x, u = (long long)y0 * pitch + x;
    int step = ((g0 + __ldg(G + r)) * (__ldg(L + r) - L0) + (g0 + __ldg(G + l)) * (__ldg(L + l) - L0) +
                (g0 + __ldg(G + d)) * (__ldg(L + d) - L0) + (g0 + __ldg(G + u)) * (__ldg(L + u) - L0)) >> 16;
    dst[(long long)blockIdx.z * stride + c] = ((stepfac * step) >> 16) + L0;
}

// first derivatives (gDerivate akazed.cu:3339-3368) and determinant (gHessianDeterminant :3371-3403)
__global__ void __launch_bounds__(256) k_fderiv(const int* __restrict__ src, int* __restrict__ lx, int* __restrict__ ly, int step, int fac1, int fac2,
                                                int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    INb n = inb(src + (long long)blockIdx.z * stride, x, y, w, h, pitch, step);
    long long o = (long long)blockIdx.z * stride + (long long)y * pitch + x;
    lx[o] = (fac1 * (n.ur + n.lr - n.ul - n.ll) + fac2 * (n.cr - n.cl)) >> 16;
    ly[o] = (fac1 * (n.lr + n.ll - n.ur - n.ul) + fac2 * (n.lc - n.uc)) >> 16;
}

__global__ void __launch_bounds__(256) k_fhess(const int* __restrict__ lx, const int* __restrict__ ly, int* __restrict__ det, int step, int fac1, int fac2,
                                               int w, int h, int pitch, long long stride)
{
    int x = blockIdx.x * FBX + threadIdx.x, y = blockIdx.y * FBY + threadIdx.y;
    if (x >= w || y >= h) return;
    INb a = inb(lx + (long long)blockIdx.z * stride, x, y, w, h, pitch, step);
    INb b = inb(ly + (long long)blockIdx.z * stride, x, y, w, h, pitch, step);
    int dxx = (fac1 * (a.ur + a.lr - a.ul - a.ll) + fac2 * (a.cr - a.cl)) >> 16;
    int dxy = (fac1 * (a.lr + a.ll - a.ur - a.ul) + fac2 * (a.lc - a.uc)) >> 16;
    int dyy = (fac1 * (b.lr + b.ll - b.ur - b.ul) + fac2 * (b.lc - b.uc)) >> 16;
    det[(long long)blockIdx.z * stride + (long long)y * pitch + x] = dxx * dyy - dxy * dxy;
}

ITaps itaps(float var, int R)
{
    float k[6];
    akz_gauss_taps(var, R, k);
    ITaps t = {};
    for (int i = 0; i <= R; i++) t.k[i] = (int)(k[i] * 65536 + 0.5f);       // akazed.cu:3896 (kernel * 65536 is exact)
    return t;
}

}  // namespace

namespace akzk {

// separable blur u8 -> int or int -> int; tmp is a scratch plane batch
int fast_lowpass(cudaStream_t st, const void* src, int src_u8, int* dst, int* tmp, int w, int h, int sp, long long sstride,
                 int dp, long long dstride, int n, float var, int ksz)
{
    int R = radius_from_ksz(ksz);
    if (R < 0) return akz_set_error(AKZ_E_UNSUPPORTED, "Gaussian kernels larger than 11 are not implemented (akazed.cu:4007)");
    ITaps t = itaps(var, R);
    dim3 g = fgrid(w, h, n), b(FBX, FBY);
#define AKZ_FROW(RR)                                                                                                        \
    if (src_u8) k_frow<RR, unsigned char><<<g, b, 0, st>>>((const unsigned char*)src, tmp, w, h, sp, sstride, dp, dstride, t); \
    else k_frow<RR, int><<<g, b, 0, st>>>((const int*)src, tmp, w, h, sp, sstride, dp, dstride, t);                          \
    k_fcol<RR><<<g, b, 0, st>>>(tmp, dst, w, h, dp, dstride, t);
    switch (R) {
    case 2: AKZ_FROW(2) break;
    case 3: AKZ_FROW(3) break;
    case 4: AKZ_FROW(4) break;
    default: AKZ_FROW(5) break;
    }
#undef AKZ_FROW
    return 2;
}

int fast_down(cudaStream_t st, const int* src, int* dst, int* smooth, int sw, int sh, int sp, long long sstride,
              int dw, int dh, int dp, long long dstride, int n)
{
    k_fdown<<<fgrid(dw, dh, n), dim3(FBX, FBY), 0, st>>>(src, dst, smooth, sw, sh, sp, sstride, dw, dh, dp, dstride, itaps(1.f, 2));
    return 1;
}

// contrast factor of the integer pipeline: magnitude plane `mag` (scratch), per-frame k -> kout (device ints)
int fast_contrast(cudaStream_t st, const int* src, int* mag, int* hmax, int* hist, int* kout, float per, int override_k,
                  int w, int h, int pitch, long long stride, int n)
{
    int tot = n * AKZ_NBINS, launches = 2;
    k_finit<<<(tot + 255) / 256, 256, 0, st>>>(hmax, hist, n);
    if (override_k <= 0) {
        k_fscharr<<<fgrid(w, h, n), dim3(FBX, FBY), 0, st>>>(src, mag, hmax, w, h, pitch, stride);
        k_fhist<<<dim3((w + FBX - 1) / FBX, (h + 4 * FBY - 1) / (4 * FBY), n), dim3(FBX, FBY), 0, st>>>(mag, hmax, hist, w, h, pitch, stride);
        launches += 2;
    }
    k_fscan<<<(n + 63) / 64, 64, 0, st>>>(hist, hmax, kout, per, w, h, n, override_k);
    return launches;
}

int fast_contrast_init(cudaStream_t st, int* hmax, int* hist, int n)
{
    int tot = n * AKZ_NBINS;
    k_finit<<<(tot + 255) / 256, 256, 0, st>>>(hmax, hist, n);
    return 1;
}

int fast_contrast_tail(cudaStream_t st, const int* mag, const int* hmax, int* hist, int* kout, float per, int override_k,
                       int w, int h, int pitch, long long stride, int n)
{
    int launches = 1;
    if (override_k <= 0) {
        k_fhist<<<dim3((w + FBX - 1) / FBX, (h + 4 * FBY - 1) / (4 * FBY), n), dim3(FBX, FBY), 0, st>>>(mag, hmax, hist, w, h, pitch, stride);
        launches++;
    }
    k_fscan<<<(n + 63) / 64, 64, 0, st>>>(hist, hmax, kout, per, w, h, n, override_k);
    return launches;
}

int fast_flow(cudaStream_t st, const int* src, int* flow, int type, const int* kc, int nmul, int w, int h, int pitch, long long stride, int n)
{
    k_fflow<<<fgrid(w, h, n), dim3(FBX, FBY), 0, st>>>(src, flow, type, kc, nmul, w, h, pitch, stride);
    return 1;
}

int fast_nld_step(cudaStream_t st, const int* src, const int* flow, int* dst, float tau, int w, int h, int pitch, long long stride, int n)
{
    int stepfac = (int)(0.5f * tau * 65536 + 0.5f);               // akazed.cu:4240 (0.5f*tau*65536 is exact up to the first product)
    k_fnld<<<fgrid(w, h, n), dim3(FBX, FBY), 0, st>>>(src, flow, dst, stepfac, w, h, pitch, stride);
    return 1;
}

int fast_hessian(cudaStream_t st, const int* smooth, int* lx, int* ly, int* det, int step, int w, int h, int pitch, long long stride, int n)
{
    float f1, f2;
    hessian_factors(&f1, &f2);
    int fac1 = (int)(f1 * 65536 + 0.5f), fac2 = (int)(f2 * 65536 + 0.5f);       // akazed.cu:4184-4185
    dim3 g = fgrid(w, h, n), b(FBX, FBY);
    k_fderiv<<<g, b, 0, st>>>(smooth, lx, ly, step, fac1, fac2, w, h, pitch, stride);
    k_fhess<<<g, b, 0, st>>>(lx, ly, det, step, fac1, fac2, w, h, pitch, stride);
    return 2;
}

}  // namespace akzk
