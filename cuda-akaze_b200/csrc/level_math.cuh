// Per-pixel arithmetic of the level kernels (k_prep2 / k_prep3 in level_prep.cu, k_deriv4 in deriv_stream.cu), float and
// integer pipeline.  Every expression is the pinned sequence of common.cuh (float) or the reference's integer form.
#pragma once
#include "common.cuh"

namespace akz {

struct LevelMathArgs {
    float fac1, fac2, k0, k1, k2;
    int ik0, ik1, ik2, ifac1, ifac2;     // INT: 16.16 fixed-point taps and derivative factors (akazed.cu:3896, :4184)
};

// ---- arithmetic of the two pipelines ---------------------------------------------------------------------------------------
// INT = false: the float pipeline, pinned operation order of common.cuh.  INT = true: the integer pipeline (namespace fastakaze,
// akazed.cu:2781-4366): planes are int32 whose bit patterns travel through the same float registers / shared-memory tiles;
// every product sum is shifted right by 16, integer addition is associative, so only the truncation points matter.
__device__ __forceinline__ int fi(float v) { return __float_as_int(v); }
__device__ __forceinline__ float fb(int v) { return __int_as_float(v); }

template <bool INT>
__device__ __forceinline__ float p2_gauss(float m2, float m1, float c, float p1, float p2, const LevelMathArgs& a)
{
    if (!INT) return gauss_r2(m2, m1, c, p1, p2, a.k0, a.k1, a.k2);
    return fb((a.ik0 * fi(c) + a.ik1 * (fi(m1) + fi(p1)) + a.ik2 * (fi(m2) + fi(p2))) >> 16);        // akazed.cu:2786-3076
}
// conductance of one pixel from its 3 x 3 neighbourhood of the smoothed level; ikc = 1 / k^2
template <bool INT>
__device__ __forceinline__ float p2_flow(float ul, float uc, float ur, float cl, float cr, float ll, float lc, float lr, int type, float ikc)
{
    if (!INT) {
        float dx = scharr_dx(ul, ur, cl, cr, ll, lr);
        float dy = scharr_dy(ul, uc, ur, ll, lc, lr);
        return conductance(type, __fmul_rn(grad_sq(dx, dy), ikc));
    }
    // gFlowNaive akazed.cu:3406-3446, written in the reference's form (same contraction by nvcc as k_fflow)
    const int dx = 10 * (fi(cr) - fi(cl)) + 3 * (fi(ur) + fi(lr) - fi(ul) - fi(ll));
    const int dy = 10 * (fi(lc) - fi(uc)) + 3 * (fi(ll) + fi(lr) - fi(ul) - fi(ur));
    const float dif2 = (dx * dx + dy * dy) * ikc;
    int g;
    if (type == 0) g = (int)(__expf(-dif2) * 65536 + 0.5f);
    else if (type == 1) g = (int)(1.f / (1.f + dif2) * 65536 + 0.5f);
    else if (type == 2) g = (int)((1.f - __expf(-3.315f / __powf(dif2, 4))) * 65536 + 0.5f);
    else g = (int)(1.f / __fsqrt_rn(1.f + dif2) * 65536 + 0.5f);
    return fb(g);
}
// first derivatives (x: sum_x / cr - cl, y: sum_y / lc - uc)
template <bool INT>
__device__ __forceinline__ void p2_deriv1(float ul, float uc, float ur, float cl, float cr, float ll, float lc, float lr, const LevelMathArgs& a, float& vx, float& vy)
{
    if (!INT) {
        vx = deriv1(sum_x(ul, ur, ll, lr), __fsub_rn(cr, cl), a.fac1, a.fac2);
        vy = deriv1(sum_y(ul, ur, ll, lr), __fsub_rn(lc, uc), a.fac1, a.fac2);
    } else {                                                                                         // gDerivate akazed.cu:3339-3368
        vx = fb((a.ifac1 * (fi(ur) + fi(lr) - fi(ul) - fi(ll)) + a.ifac2 * (fi(cr) - fi(cl))) >> 16);
        vy = fb((a.ifac1 * (fi(lr) + fi(ll) - fi(ur) - fi(ul)) + a.ifac2 * (fi(lc) - fi(uc))) >> 16);
    }
}
// determinant of the Hessian from the neighbourhoods of Lx (xu*, xc*, xl*) and Ly (yu*, yl*)
template <bool INT>
__device__ __forceinline__ float p2_det(float xul, float xuc, float xur, float xcl, float xcr, float xll, float xlc, float xlr,
                                        float yul, float yuc, float yur, float yll, float ylc, float ylr, const LevelMathArgs& a)
{
    if (!INT) {
        float dxx = deriv2(sum_x(xul, xur, xll, xlr), __fsub_rn(xcr, xcl), a.fac1, a.fac2);
        float dxy = deriv2(sum_y(xul, xur, xll, xlr), __fsub_rn(xlc, xuc), a.fac1, a.fac2);
        float dyy = deriv2(sum_y(yul, yur, yll, ylr), __fsub_rn(ylc, yuc), a.fac1, a.fac2);
        return hess_det(dxx, dyy, dxy);
    }
    // gHessianDeterminant akazed.cu:3371-3403
    const int dxx = (a.ifac1 * (fi(xur) + fi(xlr) - fi(xul) - fi(xll)) + a.ifac2 * (fi(xcr) - fi(xcl))) >> 16;
    const int dxy = (a.ifac1 * (fi(xlr) + fi(xll) - fi(xur) - fi(xul)) + a.ifac2 * (fi(xlc) - fi(xuc))) >> 16;
    const int dyy = (a.ifac1 * (fi(ylr) + fi(yll) - fi(yur) - fi(yul)) + a.ifac2 * (fi(ylc) - fi(yuc))) >> 16;
    return fb(dxx * dyy - dxy * dxy);
}

}  // namespace akz
