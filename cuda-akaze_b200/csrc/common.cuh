// Shared device helpers: reflect-101 indexing and the per-pixel arithmetic of every stage written
// with explicit rounding intrinsics.  The operation order of each expression reproduces what the
// reference's sm_100a build executes (read from its SASS, see DESIGN.md "pinned arithmetic"), so
// planes are bit-identical no matter how the kernels around these expressions are tiled or fused.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>

#define AKZ_NBINS 300          // akazed.cu:8
#define AKZ_MAX_DIST 96        // akazed.cu:11
#define AKZ_MAX_LEVELS 40
#define AKZ_MAX_STEPS 128

namespace akz {

// reflect-101, reference: abs(ix - i) and borderAdd (akazed.cu:162-170)
__device__ __forceinline__ int refl_lo(int i) { return i < 0 ? -i : i; }
__device__ __forceinline__ int refl_hi(int i, int m) { return i < m ? i : m + m - 2 - i; }
__device__ __forceinline__ int refl(int i, int m) { return refl_hi(refl_lo(i), m); }

// u8 -> [0,1] float exactly as cv::Mat::convertTo(CV_32F, 1/255.) evaluates it (main.cpp:149)
__device__ __forceinline__ float u8_to_unit(unsigned char v) { return __fmul_rn((float)v, (float)(1.0 / 255.0)); }

// ---- Gaussian taps (akazed.cu:225-238, :281-286) -----------------------------------------------
// acc = (x-1 + x+1)*k1 ; acc = fma(x0,k0,acc) ; acc = fma(x-i + x+i, ki, acc), i = 2..R
template <int R>
struct Taps { float k[R + 1]; };

__device__ __forceinline__ float gauss_r2(float m2, float m1, float c, float p1, float p2, float k0, float k1, float k2)
{
    float acc = __fmul_rn(__fadd_rn(m1, p1), k1);
    acc = __fmaf_rn(c, k0, acc);
    return __fmaf_rn(__fadd_rn(m2, p2), k2, acc);
}

// ---- Scharr (akazed.cu:664-665, :1088-1089) ------------------------------------------------------
__device__ __forceinline__ float scharr_dx(float ul, float ur, float cl, float cr, float ll, float lr)
{
    float s = __fsub_rn(__fsub_rn(__fadd_rn(ur, lr), ul), ll);
    return __fmaf_rn(__fsub_rn(cr, cl), 10.f, __fmul_rn(3.f, s));
}
__device__ __forceinline__ float scharr_dy(float ul, float uc, float ur, float ll, float lc, float lr)
{
    float s = __fsub_rn(__fsub_rn(__fadd_rn(lr, ll), ul), ur);
    return __fmaf_rn(__fsub_rn(lc, uc), 10.f, __fmul_rn(3.f, s));
}
// dx*dx + dy*dy as compiled: the dy product is rounded, the dx product is fused
__device__ __forceinline__ float grad_sq(float dx, float dy) { return __fmaf_rn(dx, dx, __fmul_rn(dy, dy)); }

// conductance (akazed.cu:1090-1106); d = ikc * |grad|^2
__device__ __forceinline__ float conductance(int type, float d)
{
    if (type == 1) return __fdiv_rn(1.f, __fadd_rn(1.f, d));                      // PM_G2
    if (type == 0) return __expf(-d);                                             // PM_G1
    if (type == 2) return 1.f - __expf(-3.315f / __powf(d, 4));                   // WEICKERT
    return __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(1.f, d)));                         // CHARBONNIER
}

// ---- explicit diffusion step (akazed.cu:1257-1262): left product rounded, then right, down, up fused
__device__ __forceinline__ float nld_update(float L0, float g0, float LL, float gL, float LR, float gR,
                                            float LD, float gD, float LU, float gU, float stepfac)
{
    float s = __fmul_rn(__fadd_rn(g0, gL), __fsub_rn(LL, L0));
    s = __fmaf_rn(__fadd_rn(g0, gR), __fsub_rn(LR, L0), s);
    s = __fmaf_rn(__fadd_rn(g0, gD), __fsub_rn(LD, L0), s);
    s = __fmaf_rn(__fadd_rn(g0, gU), __fsub_rn(LU, L0), s);
    return __fmaf_rn(s, stepfac, L0);
}

// ---- derivative filters (akazed.cu:1294-1295 and :1326-1330 contract differently) -----------------
__device__ __forceinline__ float sum_x(float ul, float ur, float ll, float lr) { return __fsub_rn(__fsub_rn(__fadd_rn(ur, lr), ul), ll); }
__device__ __forceinline__ float sum_y(float ul, float ur, float ll, float lr) { return __fsub_rn(__fsub_rn(__fadd_rn(lr, ll), ur), ul); }
__device__ __forceinline__ float deriv1(float sum, float diff, float fac1, float fac2) { return __fmaf_rn(diff, fac2, __fmul_rn(fac1, sum)); }
__device__ __forceinline__ float deriv2(float sum, float diff, float fac1, float fac2) { return __fmaf_rn(sum, fac1, __fmul_rn(fac2, diff)); }
__device__ __forceinline__ float hess_det(float dxx, float dyy, float dxy) { return __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy)); }

// ordered key for the cross-level arg-max merge: larger response wins, ties go to the lower layer
__device__ __forceinline__ unsigned long long merge_key(float resp, int layer)
{
    return ((unsigned long long)__float_as_uint(resp) << 32) | (unsigned)(0xFFFF - layer);
}
__device__ __forceinline__ float key_resp(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_layer(unsigned long long k) { return 0xFFFF - (int)(k & 0xFFFFu); }

}  // namespace akz

// ---- host side shared structs --------------------------------------------------------------------
struct AkzLevel {
    int octave, sub, w, h, pitch, nsteps, sigma_size, tau_off;
    float esigma, size, border;
    long long plane;              // pitch*h elements
    float *lt, *det, *lx, *ly;    // chunk bases: frame f at base + f*plane
};

// what keypoint kernels need to find a plane: lives in constant memory of the translation unit that uses it
struct AkzLevelDev {
    const float *lt, *det, *lx, *ly;
    long long plane;
    int w, h, pitch, octave;
    float size;
    int pad;
};

#define AKZ_CUDA_TRY(expr)                                                            \
    do {                                                                              \
        cudaError_t e_ = (expr);                                                      \
        if (e_ != cudaSuccess) return akz_set_cuda_error(e_, #expr, __FILE__, __LINE__); \
    } while (0)

// cudaFuncSetAttribute and __constant__ uploads are per device.  `if (akz_once_guard g{flag}) { ... }` runs its body exactly
// once per device, also when several host threads create contexts at the same time: the first caller holds the flag's mutex
// while the body runs and publishes the device bit afterwards; later callers see the bit (acquire) and skip.
struct akz_once_t {
    std::atomic<unsigned long long> done{0};
    std::mutex m;
};
struct akz_once_guard {
    akz_once_t& f;
    unsigned long long bit;
    bool first;
    explicit akz_once_guard(akz_once_t& fl) : f(fl), bit(0), first(false)
    {
        int d = 0;
        cudaGetDevice(&d);
        bit = 1ull << (d & 63);
        if (f.done.load(std::memory_order_acquire) & bit) return;
        f.m.lock();
        if (f.done.load(std::memory_order_relaxed) & bit) { f.m.unlock(); return; }
        first = true;
    }
    ~akz_once_guard()
    {
        if (first) { f.done.fetch_or(bit, std::memory_order_release); f.m.unlock(); }
    }
    explicit operator bool() const { return first; }
    akz_once_guard(const akz_once_guard&) = delete;
    akz_once_guard& operator=(const akz_once_guard&) = delete;
};

// restores the caller's current device when an entry point returns (the library switches to the context's device)
struct akz_device_guard {
    int prev;
    akz_device_guard() : prev(-1) { cudaGetDevice(&prev); }
    ~akz_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};

int akz_set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
int akz_set_error(int code, const char* fmt, ...);
