// Drop-in header: the stage-level host functions of the reference (akazed.h:7-124).  The float
// plane stages are served by the per-stage B200 kernels (scale_space.cu); they are synchronous like
// the reference's.  Stages that operate on the reference's dense maps / AoS pyramid layout
// (hCalcExtremaMap, hNms, hNmsR, hRefine, hCalcOrient, hDescribe) and the integer "fastakaze" stages
// are declared for source compatibility; the B200 pipeline does not route through them — use
// akaze::Akazer or the C ABI (akaze_b200.h).  Calling one of those reports the fact and exits, the
// reference's own error convention.
#pragma once
#include "akaze_structures.h"
#include "cuda_utils.h"

// the product library is built with -fvisibility=hidden: the drop-in surface is exported explicitly
#pragma GCC visibility push(default)

void setMaxNumPoints(const int num);
void getPointCounter(void** addr);
void getMaxContrastAddr(void** addr);
void setHistogram(const int* h_hist);
void setExtremaParam(const float* param, const int n);
void setOparam(const int* oparams, const int n);
void setCompareIndices();

namespace akaze
{
    void setLowPassKernel(const float* kernel, const int ksz);

    void hConv2d(float* src, float* dst, int width, int height, int pitch);
    void hSepConv2d(float* src, float* dst, int width, int height, int pitch);
    void hLowPass(float* src, float* dst, int width, int height, int pitch, float var, int ksz);
    void hDownWithSmooth(float* src, float* dst, float* smooth, int3 swhp, int3 dwhp);
    void hScharrContrast(float* src, float* grad, float& kcontrast, float per, int width, int height, int pitch);
    void hFlow(float* src, float* flow, DiffusivityType type, float kcontrast, int width, int height, int pitch);
    void hNldStep(float* img, float* flow, float* temp, float step_size, int width, int height, int pitch);
    void hHessianDeterminant(float* src, float* dx, float* dy, int step, int width, int height, int pitch);
    void hCalcExtremaMap(float* dets, float* response_map, float* size_map, int* layer_map, float* params,
        int octave, int max_scale, float threshold, int width, int height, int pitch, int opitch);
    void hNms(AkazePoint* points, float* response_map, float* size_map, int* layer_map, int psz, int width, int height, int pitch);
    void hNmsR(AkazePoint* points, float* response_map, float* size_map, int* layer_map, int psz, int neigh, int width, int height, int pitch);
    void hRefine(AkazeData& result, float* tmem, int noctaves, int max_scale);
    void hCalcOrient(AkazeData& result, float* tmem, int noctaves, int max_scale);
    void hDescribe(AkazeData& result, float* tmem, int noctaves, int max_scale, int patsize);
    void hMatch(AkazeData& result1, AkazeData& result2);
}

namespace fastakaze
{
    void hConv2dR2(unsigned char* src, int* dst, int width, int height, int pitch, float var);
    void hConv2dR2(int* src, int* dst, int width, int height, int pitch, float var);
    void hConv2dR2(unsigned char* src, int* dst, int* temp, int width, int height, int pitch, float var);
    void hConv2dR2(int* src, int* dst, int* temp, int width, int height, int pitch, float var);
    void hLowPass(unsigned char* src, int* dst, int width, int height, int pitch, float var, int ksz);
    void hLowPass(unsigned char* src, int* dst, int* temp, int width, int height, int pitch, float var, int ksz);
    void hDownWithSmooth(int* src, int* dst, int* smooth, int3 swhp, int3 dwhp);
    void hScharrContrast(int* src, int* grad, int& kcontrast, float per, int width, int height, int pitch);
    void hHessianDeterminant(int* src, int* dx, int* dy, int step, int width, int height, int pitch);
    void hFlow(int* src, int* flow, akaze::DiffusivityType type, int kcontrast, int width, int height, int pitch);
    void hNldStep(int* img, int* flow, int* temp, float step_size, int width, int height, int pitch);
    void hCalcExtremaMap(int* dets, int* response_map, float* size_map, int* layer_map, float* params,
        int octave, int max_scale, int threshold, int width, int height, int pitch, int opitch);
    void hNmsR(akaze::AkazePoint* points, int* response_map, float* size_map, int* layer_map, int psz, int neigh, int width, int height, int pitch);
    void hRefine(akaze::AkazeData& result, void* tmem, int noctaves, int max_scale);
    void hCalcOrient(akaze::AkazeData& result, void* tmem, int noctaves, int max_scale);
    void hDescribe(akaze::AkazeData& result, void* tmem, int noctaves, int max_scale, int patsize);
}

#pragma GCC visibility pop
