// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// C-ABI window onto the UNMODIFIED reference (Accustomer/CUDA-AKAZE) so that the parity
// tests can drive it through ctypes on the GPU box.  This file is ours; it is compiled
// TOGETHER with the reference sources where they lie under /root/reference (see
// oracle/Makefile) into oracle/_ref/libref_akaze.so.  No reference source is copied.
//
// Two build-time tricks, both on the compiler command line only:
//   * this TU sees Akazer's private section (ORACLE_OPEN_AKAZER) so it can call
//     Akazer::allocMemory / Akazer::detect (akaze.cpp:204, :240) with a buffer we own and
//     read the pyramid back before detectAndCompute's trailing cudaMemset (akaze.cpp:144)
//     would wipe it;
//   * akaze.cpp alone is compiled with -DhScharrContrast=hScharrContrastHook, so the
//     pipeline's call at akaze.cpp:330 lands in the hook below, which forwards to the real
//     akaze::hScharrContrast (akazed.cu:2410) and can record or override the contrast
//     factor.  The reference's max-reduction is racy; recording k is
//     what makes stage-exact comparisons of a whole run possible.
#define private public
#include "akaze.h"
#undef private
#include "akazed.h"
#include "fed.h"
#include <vector>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
int   g_k_mode = 0;        // 0 = pass through and record, 1 = override with g_k_inject
float g_k_inject = 0.f;
float g_k_seen = 0.f;
}

namespace akaze {
void hScharrContrastHook(float* src, float* grad, float& kcontrast, float per, int width, int height, int pitch)
{
    hScharrContrast(src, grad, kcontrast, per, width, height, pitch);
    g_k_seen = kcontrast;
    if (g_k_mode == 1) kcontrast = g_k_inject;
}
}

namespace {
int g_ik_mode = 0, g_ik_inject = 0, g_ik_seen = 0;
}
namespace fastakaze {
// the -D rename also hits the integer overload's declaration in akazed.h: same record / override hook for the
// integer pipeline's contrast factor (akaze.cpp:599)
void hScharrContrastHook(int* src, int* grad, int& kcontrast, float per, int width, int height, int pitch)
{
    hScharrContrast(src, grad, kcontrast, per, width, height, pitch);
    g_ik_seen = kcontrast;
    if (g_ik_mode == 1) kcontrast = g_ik_inject;
}
}
REF_API void ref_fast_set_kcontrast_mode(int mode, int k) { g_ik_mode = mode; g_ik_inject = k; }
REF_API int ref_fast_last_kcontrast() { return g_ik_seen; }

REF_API void ref_set_kcontrast_mode(int mode, float k) { g_k_mode = mode; g_k_inject = k; }
REF_API float ref_last_kcontrast() { return g_k_seen; }

REF_API int ref_sizeof_point() { return (int)sizeof(akaze::AkazePoint); }

REF_API int ref_fed_tau(float T, int M, float tau_max, int reordering, float* out, int cap)
{
    std::vector<float> tau;
    int n = fed_tau_by_process_time(T, M, tau_max, reordering != 0, tau);
    for (int i = 0; i < n && i < cap; i++) out[i] = tau[i];
    return n;
}

REF_API int ref_fed_tau_internal(int n, float scale, float tau_max, int reordering, float* out, int cap)
{
    std::vector<float> tau;
    int m = fed_tau_internal(n, scale, tau_max, reordering != 0, tau);
    for (int i = 0; i < m && i < cap; i++) out[i] = tau[i];
    return m;
}

// ---- stage seams (akazed.h) -------------------------------------------------------------
REF_API void ref_setCompareIndices() { setCompareIndices(); }
REF_API void ref_setMaxNumPoints(int n) { setMaxNumPoints(n); }
REF_API void ref_setOparam(const int* p, int n) { setOparam(p, n); }
REF_API void ref_resetPointCounter()
{
    void* a; getPointCounter(&a); cudaMemset(a, 0, sizeof(unsigned int));
}
REF_API unsigned int ref_readPointCounter()
{
    void* a; getPointCounter(&a); unsigned int v = 0;
    cudaMemcpy(&v, a, sizeof(v), cudaMemcpyDeviceToHost); return v;
}
REF_API void ref_hLowPass(float* src, float* dst, int w, int h, int p, float var, int ksz)
{ akaze::hLowPass(src, dst, w, h, p, var, ksz); }
REF_API void ref_hDownWithSmooth(float* src, float* dst, float* smooth, int sw, int sh, int sp, int dw, int dh, int dp)
{ akaze::hDownWithSmooth(src, dst, smooth, make_int3(sw, sh, sp), make_int3(dw, dh, dp)); }
REF_API float ref_hScharrContrast(float* src, float* grad, float per, int w, int h, int p)
{ float k = 0.03f; akaze::hScharrContrast(src, grad, k, per, w, h, p); return k; }
REF_API void ref_hFlow(float* src, float* flow, int type, float k, int w, int h, int p)
{ akaze::hFlow(src, flow, (akaze::DiffusivityType)type, k, w, h, p); }
REF_API void ref_hNldStep(float* img, float* flow, float* dst, float tau, int w, int h, int p)
{ akaze::hNldStep(img, flow, dst, tau, w, h, p); }
REF_API void ref_hHessianDeterminant(float* src, float* dx, float* dy, int step, int w, int h, int p)
{ akaze::hHessianDeterminant(src, dx, dy, step, w, h, p); }
REF_API void ref_hCalcExtremaMap(float* dets, float* resp, float* size, int* layer, float* params,
                                 int octave, int max_scale, float thr, int w, int h, int p, int op)
{ akaze::hCalcExtremaMap(dets, resp, size, layer, params, octave, max_scale, thr, w, h, p, op); }
REF_API void ref_hNmsR(void* pts, float* resp, float* size, int* layer, int psz, int neigh, int w, int h, int p)
{ akaze::hNmsR((akaze::AkazePoint*)pts, resp, size, layer, psz, neigh, w, h, p); }

static akaze::AkazeData mkdata(void* d_pts, int n, int cap)
{ akaze::AkazeData d; d.num_pts = n; d.max_pts = cap; d.h_data = NULL; d.d_data = (akaze::AkazePoint*)d_pts; return d; }

REF_API void ref_hRefine(void* d_pts, int n, int cap, float* tmem, int noct, int S)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) akaze::hRefine(d, tmem, noct, S); }
REF_API void ref_hCalcOrient(void* d_pts, int n, int cap, float* tmem, int noct, int S)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) akaze::hCalcOrient(d, tmem, noct, S); }
REF_API void ref_hDescribe(void* d_pts, int n, int cap, float* tmem, int noct, int S, int pat)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) akaze::hDescribe(d, tmem, noct, S, pat); }
REF_API void ref_hMatch(void* d_q, int nq, void* d_t, int nt)
{
    akaze::AkazeData a = mkdata(d_q, nq, nq), b = mkdata(d_t, nt, nt);
    if (nq > 0) akaze::hMatch(a, b);
}

// ---- integer pipeline stage seams (akazed.h:88-124) ------------------------------------------
REF_API void ref_fast_hConv2dR2_u8(unsigned char* src, int* dst, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, w, h, p, var); }
REF_API void ref_fast_hConv2dR2_i(int* src, int* dst, int w, int h, int p, float var) { fastakaze::hConv2dR2(src, dst, w, h, p, var); }
REF_API void ref_fast_hLowPass(unsigned char* src, int* dst, int w, int h, int p, float var, int ksz) { fastakaze::hLowPass(src, dst, w, h, p, var, ksz); cudaDeviceSynchronize(); }
REF_API void ref_fast_hDownWithSmooth(int* src, int* dst, int* smooth, int sw, int sh, int sp, int dw, int dh, int dp)
{ fastakaze::hDownWithSmooth(src, dst, smooth, make_int3(sw, sh, sp), make_int3(dw, dh, dp)); }
REF_API int ref_fast_hScharrContrast(int* src, int* grad, float per, int w, int h, int p)
{ int k = 1; fastakaze::hScharrContrast(src, grad, k, per, w, h, p); return k; }
REF_API void ref_fast_hFlow(int* src, int* flow, int type, int k, int w, int h, int p)
{ fastakaze::hFlow(src, flow, (akaze::DiffusivityType)type, k, w, h, p); }
REF_API void ref_fast_hNldStep(int* img, int* flow, int* dst, float tau, int w, int h, int p) { fastakaze::hNldStep(img, flow, dst, tau, w, h, p); }
REF_API void ref_fast_hHessianDeterminant(int* src, int* dx, int* dy, int step, int w, int h, int p)
{ fastakaze::hHessianDeterminant(src, dx, dy, step, w, h, p); }
REF_API void ref_fast_hCalcExtremaMap(int* dets, int* resp, float* size, int* layer, float* params,
                                      int octave, int max_scale, int thr, int w, int h, int p, int op)
{ fastakaze::hCalcExtremaMap(dets, resp, size, layer, params, octave, max_scale, thr, w, h, p, op); }
REF_API void ref_fast_hNmsR(void* pts, int* resp, float* size, int* layer, int psz, int neigh, int w, int h, int p)
{ fastakaze::hNmsR((akaze::AkazePoint*)pts, resp, size, layer, psz, neigh, w, h, p); }

// ---- whole pipeline (akaze.h) -------------------------------------------------------------
struct RefAkazer {
    akaze::Akazer az;
    int oparams[5 * 8 + 1];
    int noct_used;
};

REF_API void* ref_akazer_create(int w, int h, int p, int noct, int S, float per, float kc, float soffset,
                                int reordering, float dfac, float dthr, int diffusivity, int pattern)
{
    RefAkazer* r = new RefAkazer;
    memset(r->oparams, 0, sizeof(r->oparams));
    r->az.init(make_int3(w, h, p), noct, S, per, kc, soffset, reordering != 0, dfac, dthr, diffusivity, pattern);
    r->noct_used = noct;
    return r;
}
REF_API void ref_akazer_destroy(void* hnd) { delete (RefAkazer*)hnd; }

// The reference's public entry point, as main.cpp:199-205 drives it.
REF_API int ref_akazer_detectAndCompute(void* hnd, float* d_img, int w, int h, int p, int desc,
                                        void* d_pts, void* h_pts, int cap)
{
    RefAkazer* r = (RefAkazer*)hnd;
    akaze::AkazeData d; d.num_pts = 0; d.max_pts = cap;
    d.d_data = (akaze::AkazePoint*)d_pts; d.h_data = (akaze::AkazePoint*)h_pts;
    r->az.detectAndCompute(d_img, d, make_int3(w, h, p), desc != 0);
    return d.num_pts;
}

// Same pipeline, but the pyramid lives in a buffer the caller can read afterwards.
// Returns num_pts; *tmem_out receives a cudaMalloc'ed pointer (free with ref_cuda_free),
// oparams_out receives [osizes(n) | offsets(n+1) | owhps(n*3)] as akaze.cpp:104-107 lays them out.
REF_API int ref_akazer_detect_keep(void* hnd, float* d_img, int w, int h, int p, int desc,
                                   void* d_pts, int cap, float** tmem_out, int* oparams_out, int* noct_out)
{
    RefAkazer* r = (RefAkazer*)hnd;
    akaze::Akazer& az = r->az;
    int3 whp0 = make_int3(w, h, p);
    // learn how many octaves survive BEFORE laying out the scratch ints: akaze.cpp:215-219
    // shrinks noctaves inside allocMemory, after :104-107 already sliced the array (App. B-12)
    {
        int wq = w, hq = h, n = 1;
        for (int j = 1; j < az.noctaves; j++) { wq >>= 1; hq >>= 1; if (wq < 80 || hq < 80) break; n++; }
        az.noctaves = n;
    }
    int n = az.noctaves;
    int* osizes = r->oparams;
    int* offsets = osizes + n;
    int3* owhps = (int3*)(offsets + n + 1);
    float* tmem = NULL;
    az.allocMemory((void**)&tmem, whp0, owhps, osizes, offsets, false);
    akaze::AkazeData d = mkdata(d_pts, 0, cap);
    az.detect(d, tmem, d_img, owhps, osizes, offsets);
    if (desc && d.num_pts > 0) {
        akaze::hCalcOrient(d, tmem, az.noctaves, az.max_scale);
        akaze::hDescribe(d, tmem, az.noctaves, az.max_scale, az.descriptor_pattern_size);
    }
    *tmem_out = tmem;
    memcpy(oparams_out, r->oparams, sizeof(int) * (5 * n + 1));
    *noct_out = n;
    return d.num_pts;
}
// Integer pipeline (akaze.cpp:153-201, :506-743) with the pyramid kept, as ref_akazer_detect_keep does for the float one.
REF_API int ref_akazer_fast_detect_keep(void* hnd, unsigned char* d_img, int w, int h, int p, int desc,
                                        void* d_pts, int cap, void** tmem_out, int* oparams_out, int* noct_out)
{
    RefAkazer* r = (RefAkazer*)hnd;
    akaze::Akazer& az = r->az;
    int3 whp0 = make_int3(w, h, p);
    {
        int wq = w, hq = h, n = 1;
        for (int j = 1; j < az.noctaves; j++) { wq >>= 1; hq >>= 1; if (wq < 80 || hq < 80) break; n++; }
        az.noctaves = n;
    }
    int n = az.noctaves;
    int* osizes = r->oparams;
    int* offsets = osizes + n;
    int3* owhps = (int3*)(offsets + n + 1);
    void* tmem = NULL;
    az.allocMemory(&tmem, whp0, owhps, osizes, offsets, false);
    akaze::AkazeData d = mkdata(d_pts, 0, cap);
    az.fastDetect(d, tmem, d_img, owhps, osizes, offsets);
    if (desc && d.num_pts > 0) {
        fastakaze::hCalcOrient(d, tmem, az.noctaves, az.max_scale);
        fastakaze::hDescribe(d, tmem, az.noctaves, az.max_scale, az.descriptor_pattern_size);
    }
    *tmem_out = tmem;
    memcpy(oparams_out, r->oparams, sizeof(int) * (5 * n + 1));
    *noct_out = n;
    return d.num_pts;
}
REF_API int ref_akazer_fastDetectAndCompute(void* hnd, unsigned char* d_img, int w, int h, int p, int desc, void* d_pts, void* h_pts, int cap)
{
    RefAkazer* r = (RefAkazer*)hnd;
    akaze::AkazeData d; d.num_pts = 0; d.max_pts = cap;
    d.d_data = (akaze::AkazePoint*)d_pts; d.h_data = (akaze::AkazePoint*)h_pts;
    r->az.fastDetectAndCompute(d_img, d, make_int3(w, h, p), desc != 0);
    return d.num_pts;
}
REF_API void ref_fast_hRefine(void* d_pts, int n, int cap, void* tmem, int noct, int S)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) fastakaze::hRefine(d, tmem, noct, S); }
REF_API void ref_fast_hCalcOrient(void* d_pts, int n, int cap, void* tmem, int noct, int S)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) fastakaze::hCalcOrient(d, tmem, noct, S); }
REF_API void ref_fast_hDescribe(void* d_pts, int n, int cap, void* tmem, int noct, int S, int pat)
{ akaze::AkazeData d = mkdata(d_pts, n, cap); if (n > 0) fastakaze::hDescribe(d, tmem, noct, S, pat); }
REF_API void ref_cuda_free(void* p) { cudaFree(p); }

// gHammingMatch reads 3 bytes of shared memory it never wrote: whatever earlier kernels left there is
// added to every distance of a query.  This helper (ours, test infrastructure) fills the shared memory of every SM with
// zeros so that a following hMatch / cuMatch is deterministic and comparable bit for bit.
__global__ void k_scrub_shared(int words)
{
    extern __shared__ int scrub[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) scrub[i] = 0;
    __syncthreads();
    if (scrub[(threadIdx.x * 97) % words] != 0) printf("scrub\n");        // keeps the stores alive
}
REF_API int ref_scrub_shared_memory()
{
    int dev = 0, nsm = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cudaFuncSetAttribute(k_scrub_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm) != cudaSuccess) return -1;
    k_scrub_shared<<<nsm * 8, 256, maxsm>>>(maxsm / 4);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

REF_API void ref_cuMatch(void* d_q, void* h_q, int nq, void* d_t, int nt)
{
    akaze::AkazeData a = mkdata(d_q, nq, nq), b = mkdata(d_t, nt, nt);
    a.h_data = (akaze::AkazePoint*)h_q;
    if (nq > 0) akaze::cuMatch(a, b);
}
