// TEST INFRASTRUCTURE.  C window onto the PRODUCT's drop-in C++ surface (include/akaze.h, akazed.h, fed.h as
// implemented by cuda-akaze_b200/csrc/akaze_compat.cu), so that the GPU tests and bench.py can drive
// akaze::Akazer / initAkazeData / cuMatch and the h* stage functions exactly as the reference's main.cpp:192-209,
// :300-317 does -- the twin of oracle/ref_shim.cu, which opens the same window onto the compiled reference.
// Built by tests/cpp/Makefile against lib/libakaze_b200.so; nothing here computes.
#include "akaze.h"
#include "akazed.h"
#include "fed.h"
#include <chrono>
#include <cstring>
#include <vector>

#define DROPIN_API extern "C" __attribute__((visibility("default")))

DROPIN_API int dropin_sizeof_point() { return (int)sizeof(akaze::AkazePoint); }

// ---- AkazeData through initAkazeData / freeAkazeData (akaze.h:12-13; main.cpp:190-191, :231-232) ----------------
DROPIN_API void* dropin_data_create(int max_pts, int host, int dev)
{
    akaze::AkazeData* d = new akaze::AkazeData;
    akaze::initAkazeData(*d, max_pts, host != 0, dev != 0);
    return d;
}
DROPIN_API void dropin_data_free(void* p)
{
    akaze::AkazeData* d = (akaze::AkazeData*)p;
    akaze::freeAkazeData(*d);
    delete d;
}
DROPIN_API int   dropin_data_num(void* p) { return ((akaze::AkazeData*)p)->num_pts; }
DROPIN_API int   dropin_data_max(void* p) { return ((akaze::AkazeData*)p)->max_pts; }
DROPIN_API void* dropin_data_host(void* p) { return ((akaze::AkazeData*)p)->h_data; }
DROPIN_API void* dropin_data_dev(void* p) { return ((akaze::AkazeData*)p)->d_data; }
DROPIN_API void  dropin_data_set_num(void* p, int n) { ((akaze::AkazeData*)p)->num_pts = n; }

// ---- Akazer (akaze.h:19-30) --------------------------------------------------------------------------------------
DROPIN_API void* dropin_akazer_create(int w, int h, int p, int noct, int S, float per, float kc, float soffset,
                                      int reordering, float dfac, float dthr, int diffusivity, int pattern)
{
    akaze::Akazer* a = new akaze::Akazer;
    a->init(make_int3(w, h, p), noct, S, per, kc, soffset, reordering != 0, dfac, dthr, diffusivity, pattern);
    return a;
}
DROPIN_API void dropin_akazer_destroy(void* a) { delete (akaze::Akazer*)a; }

DROPIN_API int dropin_akazer_detectAndCompute(void* a, float* d_img, void* data, int w, int h, int p, int desc)
{
    akaze::AkazeData* d = (akaze::AkazeData*)data;
    ((akaze::Akazer*)a)->detectAndCompute(d_img, *d, make_int3(w, h, p), desc != 0);
    return d->num_pts;
}
DROPIN_API int dropin_akazer_fastDetectAndCompute(void* a, unsigned char* d_img, void* data, int w, int h, int p, int desc)
{
    akaze::AkazeData* d = (akaze::AkazeData*)data;
    ((akaze::Akazer*)a)->fastDetectAndCompute(d_img, *d, make_int3(w, h, p), desc != 0);
    return d->num_pts;
}
// the reference's timed loop (main.cpp:199-205): iters calls, host clock around them (each call is synchronous)
DROPIN_API float dropin_akazer_time(void* a, void* d_img, void* data, int w, int h, int p, int desc, int fast, int iters)
{
    akaze::AkazeData* d = (akaze::AkazeData*)data;
    akaze::Akazer* az = (akaze::Akazer*)a;
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) {
        if (fast) az->fastDetectAndCompute((unsigned char*)d_img, *d, make_int3(w, h, p), desc != 0);
        else az->detectAndCompute((float*)d_img, *d, make_int3(w, h, p), desc != 0);
    }
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    return (float)(std::chrono::duration<double, std::milli>(t1 - t0).count() / (iters > 0 ? iters : 1));
}

DROPIN_API void dropin_cuMatch(void* data1, void* data2)
{
    akaze::cuMatch(*(akaze::AkazeData*)data1, *(akaze::AkazeData*)data2);
}
DROPIN_API void dropin_hMatch(void* data1, void* data2)
{
    akaze::hMatch(*(akaze::AkazeData*)data1, *(akaze::AkazeData*)data2);
}

// ---- fed.h ----------------------------------------------------------------------------------------------------------
DROPIN_API int dropin_fed_tau(float T, int M, float tau_max, int reordering, float* out, int cap)
{
    std::vector<float> tau;
    int n = fed_tau_by_process_time(T, M, tau_max, reordering != 0, tau);
    for (int i = 0; i < n && i < cap; i++) out[i] = tau[i];
    return n;
}
DROPIN_API int dropin_fed_tau_internal(int n, float scale, float tau_max, int reordering, float* out, int cap)
{
    std::vector<float> tau;
    int m = fed_tau_internal(n, scale, tau_max, reordering != 0, tau);
    for (int i = 0; i < m && i < cap; i++) out[i] = tau[i];
    return m;
}

// ---- akazed.h stage functions, float (akazed.h:32-77) --------------------------------------------------------------
DROPIN_API void dropin_hLowPass(float* src, float* dst, int w, int h, int p, float var, int ksz) { akaze::hLowPass(src, dst, w, h, p, var, ksz); }
DROPIN_API void dropin_hDownWithSmooth(float* src, float* dst, float* smooth, int sw, int sh, int sp, int dw, int dh, int dp)
{ akaze::hDownWithSmooth(src, dst, smooth, make_int3(sw, sh, sp), make_int3(dw, dh, dp)); }
DROPIN_API float dropin_hScharrContrast(float* src, float* grad, float per, int w, int h, int p)
{ float k = 0.03f; akaze::hScharrContrast(src, grad, k, per, w, h, p); return k; }
DROPIN_API void dropin_hFlow(float* src, float* flow, int type, float k, int w, int h, int p)
{ akaze::hFlow(src, flow, (akaze::DiffusivityType)type, k, w, h, p); }
DROPIN_API void dropin_hNldStep(float* img, float* flow, float* dst, float tau, int w, int h, int p) { akaze::hNldStep(img, flow, dst, tau, w, h, p); }
DROPIN_API void dropin_hHessianDeterminant(float* src, float* dx, float* dy, int step, int w, int h, int p)
{ akaze::hHessianDeterminant(src, dx, dy, step, w, h, p); }

// ---- akazed.h stage functions, integer (akazed.h:88-110) ------------------------------------------------------------
DROPIN_API void dropin_fast_hConv2dR2_u8(unsigned char* src, int* dst, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, w, h, p, var); }
DROPIN_API void dropin_fast_hConv2dR2_i(int* src, int* dst, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, w, h, p, var); }
DROPIN_API void dropin_fast_hConv2dR2_u8_t(unsigned char* src, int* dst, int* tmp, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, tmp, w, h, p, var); }
DROPIN_API void dropin_fast_hConv2dR2_i_t(int* src, int* dst, int* tmp, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, tmp, w, h, p, var); }
DROPIN_API void dropin_fast_hLowPass(unsigned char* src, int* dst, int w, int h, int p, float var, int ksz) { fastakaze::hLowPass(src, dst, w, h, p, var, ksz); }
DROPIN_API void dropin_fast_hLowPass_t(unsigned char* src, int* dst, int* tmp, int w, int h, int p, float var, int ksz) { fastakaze::hLowPass(src, dst, tmp, w, h, p, var, ksz); }
DROPIN_API void dropin_fast_hDownWithSmooth(int* src, int* dst, int* smooth, int sw, int sh, int sp, int dw, int dh, int dp)
{ fastakaze::hDownWithSmooth(src, dst, smooth, make_int3(sw, sh, sp), make_int3(dw, dh, dp)); }
DROPIN_API int dropin_fast_hScharrContrast(int* src, int* grad, float per, int w, int h, int p)
{ int k = 1; fastakaze::hScharrContrast(src, grad, k, per, w, h, p); return k; }
DROPIN_API void dropin_fast_hFlow(int* src, int* flow, int type, int k, int w, int h, int p)
{ fastakaze::hFlow(src, flow, (akaze::DiffusivityType)type, k, w, h, p); }
DROPIN_API void dropin_fast_hNldStep(int* img, int* flow, int* dst, float tau, int w, int h, int p) { fastakaze::hNldStep(img, flow, dst, tau, w, h, p); }
DROPIN_API void dropin_fast_hHessianDeterminant(int* src, int* dx, int* dy, int step, int w, int h, int p)
{ fastakaze::hHessianDeterminant(src, dx, dy, step, w, h, p); }
