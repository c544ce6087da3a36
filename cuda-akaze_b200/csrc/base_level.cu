// k_base2<Tin>: the base level (0,0) in one pass over the input frame.
//
// One read of a 64x64 input tile (+4 halo, float [0,1] or raw u8) produces
//   * Lt(0,0)   = Gaussian blur with sigma0^2 = soffset^2, radius 4        (akaze.cpp:331, gConv2d<4> akazed.cu:204)
//   * the Scharr gradient magnitude of the sigma = 1 blur (radius 2), written to a scratch plane, and its
//     per-frame maximum (atomicMax on the float bits)                       (akaze.cpp:329-330, akazed.cu:644-667, :2410-2449)
// k_hist2 then bins the magnitude plane with the now known maximum          (akazed.cu:901-938, in-image pixels only)
// and k_contrast_scan (scale_space.cu) turns the histogram into k on the device.
// Replaces 2 x k_lowpass + k_scharr_max + k_contrast_hist (24 B/px of traffic, two recomputations of the gradient)
// by 16 B/px.  Both blurs are symmetric operators, so evaluating them on the reflect-101 extended tile gives the
// value at the mirrored pixel bit for bit; no border fix-up is needed here.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>

using namespace akz;

namespace {

constexpr int B2_T = 64, B2_O = 8, B2_SP = B2_T + 2 * B2_O;      // tile, frame offset (cols), pitch 80
constexpr int B2_R = B2_T + 8;                                   // 72 rows: Y0-4 .. Y0+67
constexpr int B2_NT = 512;

struct Base2Args {
    const void* img;
    float* lt;            // Lt(0,0)
    float* mag;           // gradient magnitude scratch plane (may be null: skip the contrast part)
    unsigned* hmax_bits;
    long long istride, plane;
    int w, h, ipitch, pitch, vec_ok;
    float a0, a1, a2, a3, a4;      // sigma0 taps (radius 4)
    float k0, k1, k2;              // sigma = 1 taps (radius 2)
    int ia[5], ik[3];              // INT: the same taps in 16.16 fixed point (akazed.cu:3896)
};

__device__ __forceinline__ float4 lds4(const float* p)
{
    float4 v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// radius-4 Gaussian in the reference's operation order (akazed.cu:225-238): (x-1 + x+1)*k1, fma(x0,k0), fma(x-i + x+i, ki) i = 2..4
__device__ __forceinline__ float gauss_r4(float m4, float m3, float m2, float m1, float c, float p1, float p2, float p3, float p4,
                                          float k0, float k1, float k2, float k3, float k4)
{
    float acc = __fmul_rn(__fadd_rn(m1, p1), k1);
    acc = __fmaf_rn(c, k0, acc);
    acc = __fmaf_rn(__fadd_rn(m2, p2), k2, acc);
    acc = __fmaf_rn(__fadd_rn(m3, p3), k3, acc);
    return __fmaf_rn(__fadd_rn(m4, p4), k4, acc);
}

__device__ __forceinline__ float to_unit(float v) { return v; }
__device__ __forceinline__ float to_unit(unsigned char v) { return u8_to_unit(v); }

// INT = true: the integer pipeline's base level (fastakaze, akaze.cpp:593-617): pixels stay 0..255 integers whose bit patterns
// travel through the float tiles; a blur pass is (k0 x0 + sum ki (x-i + x+i)) >> 16 (gConv2d akazed.cu:2786-3076), the gradient
// magnitude (int)(sqrt(dx^2 + dy^2) + 0.5f) (gScharrContrastNaive akazed.cu:3208-3231).  Integer sums are associative.
__device__ __forceinline__ int b2i(float v) { return __float_as_int(v); }
template <bool INT, typename T> __device__ __forceinline__ float b2_px(T v) { return INT ? __int_as_float((int)v) : to_unit(v); }
template <bool INT>
__device__ __forceinline__ float b2_gauss2(float m2, float m1, float c, float p1, float p2, const Base2Args& a)
{
    if (!INT) return gauss_r2(m2, m1, c, p1, p2, a.k0, a.k1, a.k2);
    return __int_as_float((a.ik[0] * b2i(c) + a.ik[1] * (b2i(m1) + b2i(p1)) + a.ik[2] * (b2i(m2) + b2i(p2))) >> 16);
}
template <bool INT>
__device__ __forceinline__ float b2_gauss4(float m4, float m3, float m2, float m1, float c, float p1, float p2, float p3, float p4, const Base2Args& a)
{
    if (!INT) return gauss_r4(m4, m3, m2, m1, c, p1, p2, p3, p4, a.a0, a.a1, a.a2, a.a3, a.a4);
    return __int_as_float((a.ia[0] * b2i(c) + a.ia[1] * (b2i(m1) + b2i(p1)) + a.ia[2] * (b2i(m2) + b2i(p2)) + a.ia[3] * (b2i(m3) + b2i(p3)) +
                           a.ia[4] * (b2i(m4) + b2i(p4))) >> 16);
}
template <bool INT>
__device__ __forceinline__ float b2_mag(float ul, float uc, float ur, float cl, float cr, float ll, float lc, float lr)
{
    if (!INT) {
        float dx = scharr_dx(ul, ur, cl, cr, ll, lr);
        float dy = scharr_dy(ul, uc, ur, ll, lc, lr);
        return __fsqrt_rn(grad_sq(dx, dy));
    }
    const int dx = 10 * (b2i(cr) - b2i(cl)) + 3 * (b2i(ur) + b2i(lr) - b2i(ul) - b2i(ll));
    const int dy = 10 * (b2i(lc) - b2i(uc)) + 3 * (b2i(ll) + b2i(lr) - b2i(ul) - b2i(ur));
    return __int_as_float((int)(__fsqrt_rn(dx * dx + dy * dy) + 0.5f));
}

template <typename Tin, bool INT>
__global__ void __launch_bounds__(B2_NT, 2) k_base2(const __grid_constant__ Base2Args a)
{
    constexpr int SP = B2_SP;
    extern __shared__ __align__(16) float sm[];
    float* In = sm;                         // [72][80]  input tile; later the sigma = 1 blur S1 (rows 3..68)
    float* R0 = In + B2_R * SP;             // sigma0 row pass, cols [8, 72)
    float* R1 = R0 + B2_R * SP;             // sigma1 row pass, cols [4, 76)
    const int tid = threadIdx.x, frame = blockIdx.z;
    const int X0 = blockIdx.x * B2_T, Y0 = blockIdx.y * B2_T;
    const int w = a.w, h = a.h;
    const bool interior = X0 - B2_O >= 0 && X0 + B2_T + B2_O <= w && Y0 - 4 >= 0 && Y0 + B2_T + 4 <= h;
    const bool fast = interior && a.vec_ok;
    const Tin* __restrict__ src = (const Tin*)a.img + (long long)frame * a.istride;

    // ---- 1. input tile ------------------------------------------------------------------------------------
    if (fast && sizeof(Tin) == 4) {
        for (int i = tid; i < B2_R * (SP / 4); i += B2_NT) {
            int r = i / (SP / 4), g = i - r * (SP / 4);
            cp_async16(In + r * SP + 4 * g, (const float*)src + (long long)(Y0 - 4 + r) * a.ipitch + (X0 - B2_O + 4 * g));
        }
        cp_async_wait_all();
    } else if (fast) {
        for (int i = tid; i < B2_R * (SP / 8); i += B2_NT) {               // 8 bytes = 8 pixels per item
            int r = i / (SP / 8), g = i - r * (SP / 8);
            uint2 v = __ldg(reinterpret_cast<const uint2*>((const unsigned char*)src + (long long)(Y0 - 4 + r) * a.ipitch + (X0 - B2_O + 8 * g)));
            float* d = In + r * SP + 8 * g;
            sts4(d, b2_px<INT>((unsigned char)(v.x)), b2_px<INT>((unsigned char)(v.x >> 8)), b2_px<INT>((unsigned char)(v.x >> 16)), b2_px<INT>((unsigned char)(v.x >> 24)));
            sts4(d + 4, b2_px<INT>((unsigned char)(v.y)), b2_px<INT>((unsigned char)(v.y >> 8)), b2_px<INT>((unsigned char)(v.y >> 16)), b2_px<INT>((unsigned char)(v.y >> 24)));
        }
    } else {
        for (int i = tid; i < B2_R * SP; i += B2_NT) {
            int r = i / SP, c = i - r * SP;
            int sy = min(max(refl(Y0 - 4 + r, h), 0), h - 1), sx = min(max(refl(X0 - B2_O + c, w), 0), w - 1);
            In[i] = b2_px<INT>(__ldg(src + (long long)sy * a.ipitch + sx));
        }
    }
    __syncthreads();

    // ---- 2. both row passes from the same three float4 loads: cols [4, 76), 18 items per row (24 lanes) --------
    for (int i = tid; i < B2_R * 24; i += B2_NT) {
        int r = i / 24, g = i - r * 24;
        if (g >= 18) continue;
        const float* p = In + r * SP + 4 + 4 * g;
        float4 v0 = lds4(p - 4), v1 = lds4(p), v2 = lds4(p + 4);
        const float e[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
        float o1[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o1[j] = b2_gauss2<INT>(e[2 + j], e[3 + j], e[4 + j], e[5 + j], e[6 + j], a);
        sts4(R1 + r * SP + 4 + 4 * g, o1[0], o1[1], o1[2], o1[3]);
        if (g >= 1 && g <= 16) {
            float o0[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                o0[j] = b2_gauss4<INT>(e[j], e[1 + j], e[2 + j], e[3 + j], e[4 + j], e[5 + j], e[6 + j], e[7 + j], e[8 + j], a);
            sts4(R0 + r * SP + 4 + 4 * g, o0[0], o0[1], o0[2], o0[3]);
        }
    }
    __syncthreads();

    // ---- 3a. sigma0 column pass -> Lt(0,0), two rows per item ---------------------------------------------------
    {
        float* ltg = a.lt + (long long)frame * a.plane;
        for (int i = tid; i < (B2_T / 2) * 16; i += B2_NT) {
            int rb = i >> 4, g = i & 15;
            int r = 4 + 2 * rb;                                           // tile row of the first output
            const float* p = R0 + (r - 4) * SP + B2_O + 4 * g;
            float4 q[10];
#pragma unroll
            for (int k = 0; k < 10; k++) q[k] = lds4(p + k * SP);
            float o[2][4];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                o[t][0] = b2_gauss4<INT>(q[t].x, q[t + 1].x, q[t + 2].x, q[t + 3].x, q[t + 4].x, q[t + 5].x, q[t + 6].x, q[t + 7].x, q[t + 8].x, a);
                o[t][1] = b2_gauss4<INT>(q[t].y, q[t + 1].y, q[t + 2].y, q[t + 3].y, q[t + 4].y, q[t + 5].y, q[t + 6].y, q[t + 7].y, q[t + 8].y, a);
                o[t][2] = b2_gauss4<INT>(q[t].z, q[t + 1].z, q[t + 2].z, q[t + 3].z, q[t + 4].z, q[t + 5].z, q[t + 6].z, q[t + 7].z, q[t + 8].z, a);
                o[t][3] = b2_gauss4<INT>(q[t].w, q[t + 1].w, q[t + 2].w, q[t + 3].w, q[t + 4].w, q[t + 5].w, q[t + 6].w, q[t + 7].w, q[t + 8].w, a);
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                int y = Y0 + 2 * rb + t, x = X0 + 4 * g;
                float* d = ltg + (long long)y * a.pitch + x;
                if (fast) *reinterpret_cast<float4*>(d) = make_float4(o[t][0], o[t][1], o[t][2], o[t][3]);
                else if (y < h) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (x + j < w) d[j] = o[t][j];
                }
            }
        }
    }
    if (a.mag == nullptr) return;

    // ---- 3b. sigma = 1 column pass -> S1 (over In): tile rows 3..68 (Y0-1 .. Y0+64), cols [4, 76) ----------------
    for (int i = tid; i < 33 * 24; i += B2_NT) {
        int rb = i / 24, g = i - rb * 24;
        if (g >= 18) continue;
        int r = 3 + 2 * rb;
        const float* p = R1 + (r - 2) * SP + 4 + 4 * g;
        float4 b0 = lds4(p), b1 = lds4(p + SP), b2 = lds4(p + 2 * SP), b3 = lds4(p + 3 * SP), b4 = lds4(p + 4 * SP), b5 = lds4(p + 5 * SP);
        float* q = In + r * SP + 4 + 4 * g;
        sts4(q, b2_gauss2<INT>(b0.x, b1.x, b2.x, b3.x, b4.x, a), b2_gauss2<INT>(b0.y, b1.y, b2.y, b3.y, b4.y, a),
             b2_gauss2<INT>(b0.z, b1.z, b2.z, b3.z, b4.z, a), b2_gauss2<INT>(b0.w, b1.w, b2.w, b3.w, b4.w, a));
        sts4(q + SP, b2_gauss2<INT>(b1.x, b2.x, b3.x, b4.x, b5.x, a), b2_gauss2<INT>(b1.y, b2.y, b3.y, b4.y, b5.y, a),
             b2_gauss2<INT>(b1.z, b2.z, b3.z, b4.z, b5.z, a), b2_gauss2<INT>(b1.w, b2.w, b3.w, b4.w, b5.w, a));
    }
    __syncthreads();

    // ---- 4. Scharr magnitude of S1 -> scratch plane, maximum -> hmax_bits[frame] -----------------------------------
    {
        float* mg = a.mag + (long long)frame * a.plane;
        unsigned best = 0u;
        for (int i = tid; i < B2_T * 16; i += B2_NT) {
            int r = i >> 4, g = i & 15;
            const float* p = In + (r + 4) * SP + B2_O + 4 * g;
            float u[6], c[6], l[6];
#define AKZ_ROW6(arr, q)                                                                         \
            {                                                                                        \
                float4 v = lds4(q);                                                                  \
                float lft = __shfl_up_sync(0xffffffffu, v.w, 1), rgt = __shfl_down_sync(0xffffffffu, v.x, 1); \
                if (g == 0) lft = (q)[-1];                                                           \
                if (g == 15) rgt = (q)[4];                                                           \
                arr[0] = lft; arr[1] = v.x; arr[2] = v.y; arr[3] = v.z; arr[4] = v.w; arr[5] = rgt; \
            }
            AKZ_ROW6(u, p - SP)
            AKZ_ROW6(c, p)
            AKZ_ROW6(l, p + SP)
#undef AKZ_ROW6
            float o[4];
            int y = Y0 + r, x = X0 + 4 * g;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                o[j] = b2_mag<INT>(u[j], u[j + 1], u[j + 2], c[j], c[j + 2], l[j], l[j + 1], l[j + 2]);
                if (y < h && x + j < w) best = max(best, __float_as_uint(o[j]));      // o >= 0: bit patterns order like the floats
            }
            float* d = mg + (long long)y * a.pitch + x;
            if (fast) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            else if (y < h) {
#pragma unroll
                for (int j = 0; j < 4; j++) if (x + j < w) d[j] = o[j];
            }
        }
        best = __reduce_max_sync(0xffffffffu, best);
        __shared__ unsigned smax[B2_NT / 32];
        if ((tid & 31) == 0) smax[tid >> 5] = best;
        __syncthreads();
        if (tid < 32) {
            unsigned v = tid < B2_NT / 32 ? smax[tid] : 0u;
            v = __reduce_max_sync(0xffffffffu, v);
            if (tid == 0) atomicMax(a.hmax_bits + frame, v);
        }
    }
}

// =====================================================================================================================
// k_base4<Tin>: the same base level as a streaming warp kernel (structure of k_blur4 / k_fed4: a warp owns a strip of 128
// columns, 4 per lane, one halo lane at either end for the radius-4 row pass, and marches down a band of rows; rows arrive
// through a cp.async landing ring; vertical windows live in registers).  At row time t: both row passes of input row t from
// the lane's four pixels and the two neighbour lanes' (8 shuffles), the sigma0 column pass over nine row-filtered rows ->
// Lt row t-4, the sigma = 1 column pass over five -> blurred row t-2, and the Scharr magnitude of row t-3 from three
// blurred rows (+ running maximum, one atomicMax per warp).  Both blurs are symmetric, so the mirrored rows (reflected row
// index) and the mirrored neighbours of the border lanes give the reference's reflect-101 values bit for bit, also for the
// Scharr taps on the blurred plane.  The row loop is unrolled by 12 (ring of the nine-row window).
// =====================================================================================================================
constexpr int B4_WARPS = 4, B4_COLS = 120, B4_RING = 6;

struct Base4Args {
    Base2Args b;
    int nstrips, nbands, band_h, nunits;
};

template <typename Tin> struct B4in { static constexpr int LANEB = sizeof(Tin) == 4 ? 16 : 4; static constexpr int SLOTB = 32 * LANEB; };

struct Base4Regs {
    float R0[12][4];        // sigma0 row pass, rows t-8 .. t (ring by row time mod 12)
    float R1[6][4];         // sigma = 1 row pass, rows t-4 .. t (mod 6)
    float S1[3][6];         // blurred rows t-4, t-3, t-2 with their left / right neighbour (production time mod 3)
};

struct Base4Lane {
    const unsigned char* psrc;      // frame base + clamped column of this lane, in bytes
    long long obase;
    unsigned ring;
    int y0, y1, t0;
    bool bl, br, store;
};

template <typename Tin>
__device__ __forceinline__ void b4_request(const Base2Args& a, const Base4Lane& ln, int row_time, int slot)
{
    const int r = min(max(refl(row_time, a.h), 0), a.h - 1);
    const unsigned char* g = ln.psrc + (long long)r * a.ipitch * (int)sizeof(Tin);
    const unsigned d = ln.ring + slot * B4in<Tin>::SLOTB;
    if (sizeof(Tin) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(g) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <typename Tin, bool INT, bool MAG, int PH>
__device__ __forceinline__ void b4_row(Base4Regs& R, const Base2Args& a, const Base4Lane& ln, int t, unsigned& best)
{
    constexpr unsigned FULL = 0xffffffffu;
    b4_request<Tin>(a, ln, t + B4_RING - 1, (PH + B4_RING - 1) % B4_RING);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(B4_RING - 1) : "memory");
    float v[4];
    {
        const unsigned sa = ln.ring + (PH % B4_RING) * B4in<Tin>::SLOTB;
        if (sizeof(Tin) == 4) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(sa) : "memory");
        } else {
            unsigned w4;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w4) : "r"(sa) : "memory");
            v[0] = b2_px<INT>((unsigned char)w4); v[1] = b2_px<INT>((unsigned char)(w4 >> 8));
            v[2] = b2_px<INT>((unsigned char)(w4 >> 16)); v[3] = b2_px<INT>((unsigned char)(w4 >> 24));
        }
    }
    // ---- both row passes
    {
        float e[12];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            e[j] = __shfl_up_sync(FULL, v[j], 1);
            e[4 + j] = v[j];
            e[8 + j] = __shfl_down_sync(FULL, v[j], 1);
        }
        if (ln.bl) { const float r0 = e[8]; e[0] = r0; e[1] = v[3]; e[2] = v[2]; e[3] = v[1]; }           // x = -4 .. -1 -> 4, 3, 2, 1
        if (ln.br) { const float l3 = e[3]; e[8] = v[2]; e[9] = v[1]; e[10] = v[0]; e[11] = l3; }         // x = w .. w+3 -> w-2 .. w-5
        float* r0 = R.R0[PH % 12];
#pragma unroll
        for (int j = 0; j < 4; j++) r0[j] = b2_gauss4<INT>(e[j], e[1 + j], e[2 + j], e[3 + j], e[4 + j], e[5 + j], e[6 + j], e[7 + j], e[8 + j], a);
        if (MAG) {
            float* r1 = R.R1[PH % 6];
#pragma unroll
            for (int j = 0; j < 4; j++) r1[j] = b2_gauss2<INT>(e[2 + j], e[3 + j], e[4 + j], e[5 + j], e[6 + j], a);
        }
    }
    // ---- sigma0 column pass -> Lt row t-4
    {
        const int rl = t - 4;
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            o[c] = b2_gauss4<INT>(R.R0[(PH + 4) % 12][c], R.R0[(PH + 5) % 12][c], R.R0[(PH + 6) % 12][c], R.R0[(PH + 7) % 12][c], R.R0[(PH + 8) % 12][c],
                                  R.R0[(PH + 9) % 12][c], R.R0[(PH + 10) % 12][c], R.R0[(PH + 11) % 12][c], R.R0[PH % 12][c], a);
        if (ln.store && rl >= ln.y0 && rl < ln.y1)
            *reinterpret_cast<float4*>(a.lt + ln.obase + (long long)rl * a.pitch) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (MAG) {
        // ---- sigma = 1 column pass -> blurred row t-2 (kept with its left / right neighbours)
        float* s = R.S1[PH % 3];
#pragma unroll
        for (int c = 0; c < 4; c++)
            s[1 + c] = b2_gauss2<INT>(R.R1[(PH + 2) % 6][c], R.R1[(PH + 3) % 6][c], R.R1[(PH + 4) % 6][c], R.R1[(PH + 5) % 6][c], R.R1[PH % 6][c], a);
        s[0] = __shfl_up_sync(FULL, s[4], 1);
        s[5] = __shfl_down_sync(FULL, s[1], 1);
        if (ln.bl) s[0] = s[2];
        if (ln.br) s[5] = s[3];
        // ---- Scharr magnitude of row t-3
        const int rm = t - 3;
        if (ln.store && rm >= ln.y0 && rm < ln.y1) {
            const float* up = R.S1[(PH + 1) % 3];
            const float* ce = R.S1[(PH + 2) % 3];
            const float* dn = R.S1[PH % 3];
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                o[c] = b2_mag<INT>(up[c], up[c + 1], up[c + 2], ce[c], ce[c + 2], dn[c], dn[c + 1], dn[c + 2]);
                best = max(best, __float_as_uint(o[c]));
            }
            *reinterpret_cast<float4*>(a.mag + ln.obase + (long long)rm * a.pitch) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <typename Tin, bool INT, bool MAG>
__global__ void __launch_bounds__(32 * B4_WARPS, 4) k_base4(const __grid_constant__ Base4Args aa)
{
    const Base2Args& a = aa.b;
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * B4_WARPS + (threadIdx.x >> 5);
    if (unit >= aa.nunits) return;
    const int per = aa.nstrips * aa.nbands;
    const int frame = unit / per, rem = unit - frame * per;
    const int band = rem / aa.nstrips, strip = rem - band * aa.nstrips;
    Base4Lane ln;
    const int gx0 = strip * B4_COLS - 4 + 4 * lane;
    const int gxl = min(max(gx0, 0), min(a.pitch, a.ipitch) - 4);
    ln.psrc = (const unsigned char*)a.img + ((long long)frame * a.istride + gxl) * (int)sizeof(Tin);
    ln.obase = (long long)frame * a.plane + gx0;
    ln.y0 = band * aa.band_h; ln.y1 = min(a.h, ln.y0 + aa.band_h); ln.t0 = ln.y0 - 4;
    ln.bl = gx0 == 0;
    ln.br = gx0 + 3 == a.w - 1;
    ln.store = lane >= 1 && lane <= 30 && gx0 >= 0 && gx0 < a.w;
    __shared__ __align__(16) unsigned char ring_mem[B4_WARPS * B4_RING * B4in<Tin>::SLOTB];
    ln.ring = (unsigned)__cvta_generic_to_shared(ring_mem + (threadIdx.x >> 5) * (B4_RING * B4in<Tin>::SLOTB) + lane * B4in<Tin>::LANEB);
    Base4Regs R;
#pragma unroll
    for (int i = 0; i < 12; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) { R.R0[i][c] = 0.f; R.R1[i % 6][c] = 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < 6; c++) R.S1[i][c] = 0.f;
#pragma unroll
    for (int k = 0; k < B4_RING - 1; k++) b4_request<Tin>(a, ln, ln.t0 + k, k);
    unsigned best = 0u;
    // Lt row y0 needs input rows y0-4 .. y0+4; Lt row y1-1 is complete at row time y1+3
    const int T = (ln.y1 - ln.y0) + 8;
    for (int it = 0; it < T; it += 12) {
        b4_row<Tin, INT, MAG, 0>(R, a, ln, ln.t0 + it, best);
        b4_row<Tin, INT, MAG, 1>(R, a, ln, ln.t0 + it + 1, best);
        b4_row<Tin, INT, MAG, 2>(R, a, ln, ln.t0 + it + 2, best);
        b4_row<Tin, INT, MAG, 3>(R, a, ln, ln.t0 + it + 3, best);
        b4_row<Tin, INT, MAG, 4>(R, a, ln, ln.t0 + it + 4, best);
        b4_row<Tin, INT, MAG, 5>(R, a, ln, ln.t0 + it + 5, best);
        b4_row<Tin, INT, MAG, 6>(R, a, ln, ln.t0 + it + 6, best);
        b4_row<Tin, INT, MAG, 7>(R, a, ln, ln.t0 + it + 7, best);
        b4_row<Tin, INT, MAG, 8>(R, a, ln, ln.t0 + it + 8, best);
        b4_row<Tin, INT, MAG, 9>(R, a, ln, ln.t0 + it + 9, best);
        b4_row<Tin, INT, MAG, 10>(R, a, ln, ln.t0 + it + 10, best);
        b4_row<Tin, INT, MAG, 11>(R, a, ln, ln.t0 + it + 11, best);
    }
    if (MAG) {
        best = __reduce_max_sync(0xffffffffu, best);
        if (lane == 0) atomicMax(a.hmax_bits + frame, best);
    }
}

// 300-bin histogram of mag * 300 / hmax (truncating multiply, clamp to 299) over the in-image pixels (akazed.cu:901-938,
// App. B-3); one block bins 8 rows; per-warp privatised shared histograms keep the atomics apart.
__global__ void __launch_bounds__(256) k_hist2(const float* __restrict__ mag, const unsigned* __restrict__ hmax_bits, int* __restrict__ hist,
                                               int w, int h, int pitch, long long plane, int vec_ok)
{
    __shared__ int sh[8][AKZ_NBINS + 4];
    const int tid = threadIdx.x, wid = tid >> 5, frame = blockIdx.y;
    for (int i = tid; i < 8 * (AKZ_NBINS + 4); i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const float hfactor = __fdiv_rn((float)AKZ_NBINS, __uint_as_float(hmax_bits[frame]));
    const float* m = mag + (long long)frame * plane;
    const int y0 = blockIdx.x * 8;
    if (vec_ok && (w & 3) == 0) {
        const int w4 = w >> 2;
        for (int i = tid; i < 8 * w4; i += 256) {
            int r = i / w4, g = i - r * w4, y = y0 + r;
            if (y >= h) break;
            float4 v = __ldg(reinterpret_cast<const float4*>(m + (long long)y * pitch + 4 * g));
            atomicAdd(&sh[wid][min((int)__fmul_rz(v.x, hfactor), AKZ_NBINS - 1)], 1);
            atomicAdd(&sh[wid][min((int)__fmul_rz(v.y, hfactor), AKZ_NBINS - 1)], 1);
            atomicAdd(&sh[wid][min((int)__fmul_rz(v.z, hfactor), AKZ_NBINS - 1)], 1);
            atomicAdd(&sh[wid][min((int)__fmul_rz(v.w, hfactor), AKZ_NBINS - 1)], 1);
        }
    } else {
        for (int i = tid; i < 8 * w; i += 256) {
            int r = i / w, x = i - r * w, y = y0 + r;
            if (y >= h) break;
            atomicAdd(&sh[wid][min((int)__fmul_rz(__ldg(m + (long long)y * pitch + x), hfactor), AKZ_NBINS - 1)], 1);
        }
    }
    __syncthreads();
    int* g = hist + (long long)frame * AKZ_NBINS;
    for (int i = tid; i < AKZ_NBINS; i += 256) {
        int s = sh[0][i] + sh[1][i] + sh[2][i] + sh[3][i] + sh[4][i] + sh[5][i] + sh[6][i] + sh[7][i];
        if (s) atomicAdd(g + i, s);
    }
}

__global__ void k_contrast_init2(unsigned* hmax_bits, int* hist, int nframes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes) hmax_bits[i] = __float_as_uint(0.03f);       // akazed.cu:2413
    if (i < nframes * AKZ_NBINS) hist[i] = 0;
}

constexpr int B2_SMEM = 3 * B2_R * B2_SP * (int)sizeof(float);

}  // namespace

namespace akzk {

// Fused base level: Lt(0,0) and (unless mag == nullptr) the gradient-magnitude plane + per-frame maximum + histogram.
// Covers sigma0 kernels of radius 4 (ksz0 = 9, the reference default soffset = 1.6); returns 0 when not applicable.
int base_level2(cudaStream_t st, const void* img, int dtype, int w, int h, int ipitch, long long istride,
                float* lt, int pitch, long long plane, float* mag, unsigned* hmax_bits, int* hist, float var0, int ksz0, int n, int int_planes)
{
    if (radius_from_ksz(ksz0) != 4 || w < 16 || h < 16) return 0;
    if (int_planes && dtype != AKZ_U8) return 0;
    static akz_once_t attr;
    if (akz_once_guard once{attr}) {
        cudaFuncSetAttribute(k_base2<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
        cudaFuncSetAttribute(k_base2<unsigned char, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
        cudaFuncSetAttribute(k_base2<unsigned char, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    }
    Base2Args a = {};
    a.img = img; a.lt = lt; a.mag = mag; a.hmax_bits = hmax_bits; a.istride = istride; a.plane = plane;
    a.w = w; a.h = h; a.ipitch = ipitch; a.pitch = pitch;
    float t0[5], t1[3];
    akz_gauss_taps(var0, 4, t0);
    akz_gauss_taps(1.f, 2, t1);
    a.a0 = t0[0]; a.a1 = t0[1]; a.a2 = t0[2]; a.a3 = t0[3]; a.a4 = t0[4];
    a.k0 = t1[0]; a.k1 = t1[1]; a.k2 = t1[2];
    for (int i = 0; i < 5; i++) a.ia[i] = (int)(t0[i] * 65536 + 0.5f);           // akazed.cu:3896 (kernel * 65536 is exact)
    for (int i = 0; i < 3; i++) a.ik[i] = (int)(t1[i] * 65536 + 0.5f);
    const size_t esz = dtype == AKZ_U8 ? 1 : 4;
    a.vec_ok = (pitch % 4 == 0) && (plane % 4 == 0) && ((uintptr_t)lt % 16 == 0) && (mag == nullptr || (uintptr_t)mag % 16 == 0) &&
               ((uintptr_t)img % 16 == 0) && (((size_t)ipitch * esz) % 16 == 0) && (((size_t)istride * esz) % 16 == 0);
    int launches = 0;
    dim3 g((w + B2_T - 1) / B2_T, (h + B2_T - 1) / B2_T, n);
    // streaming warp kernel when the rows are 16-byte aligned and there are enough (strip, band, frame) units; else the tile kernel
    Base4Args s4 = {};
    s4.b = a;
    s4.nstrips = (w + B4_COLS - 1) / B4_COLS;
    {
        const int nb = std::max(1, (h + 50) / 100);                       // bands of ~100 rows: 8 rows of warm-up each, row loop unrolled by 12
        s4.band_h = (h + nb - 1) / nb;
        s4.nbands = (h + s4.band_h - 1) / s4.band_h;
    }
    const long long units = (long long)n * s4.nstrips * s4.nbands;
    static const int min_units = [] { const char* e = getenv("AKZ_BASE_MIN_UNITS"); return e ? atoi(e) : 1024; }();
    const bool stream = a.vec_ok && (w % 4) == 0 && w >= 32 && h >= 16 && units >= min_units && units < (1ll << 30);
    s4.nunits = (int)units;
    const int g4 = (int)((units + B4_WARPS - 1) / B4_WARPS);
    if (int_planes) {
        // integer pipeline: the caller zeroes the maximum / histogram before and runs the integer histogram + scan after
        if (stream) { if (mag) k_base4<unsigned char, true, true><<<g4, 32 * B4_WARPS, 0, st>>>(s4); else k_base4<unsigned char, true, false><<<g4, 32 * B4_WARPS, 0, st>>>(s4); }
        else k_base2<unsigned char, true><<<g, B2_NT, B2_SMEM, st>>>(a);
        return 1;
    }
    if (mag) { int tot = n * AKZ_NBINS; k_contrast_init2<<<(tot + 255) / 256, 256, 0, st>>>(hmax_bits, hist, n); launches++; }
    if (stream) {
        if (dtype == AKZ_U8) { if (mag) k_base4<unsigned char, false, true><<<g4, 32 * B4_WARPS, 0, st>>>(s4); else k_base4<unsigned char, false, false><<<g4, 32 * B4_WARPS, 0, st>>>(s4); }
        else { if (mag) k_base4<float, false, true><<<g4, 32 * B4_WARPS, 0, st>>>(s4); else k_base4<float, false, false><<<g4, 32 * B4_WARPS, 0, st>>>(s4); }
    } else if (dtype == AKZ_U8) k_base2<unsigned char, false><<<g, B2_NT, B2_SMEM, st>>>(a);
    else k_base2<float, false><<<g, B2_NT, B2_SMEM, st>>>(a);
    launches++;
    if (mag) {
        k_hist2<<<dim3((h + 7) / 8, n), 256, 0, st>>>(mag, hmax_bits, hist, w, h, pitch, plane, a.vec_ok);
        launches++;
    }
    return launches;
}

}  // namespace akzk
