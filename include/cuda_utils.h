// Drop-in header: the small CUDA utility surface main.cpp uses from the reference's cuda_utils.h
// (CHECK / CheckMsg :10-37, initDevice :41-67, cpuTimer :71-77, GpuTimer :81-108, warp shuffles
// :111-142, iAlignUp / iDivUp / iExp2UpP :160-182).  Same names, arguments and error behaviour
// (message on stderr, then exit(-1)).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <cuda_runtime.h>

#define H_PI 1.5707963267948966f

#define CHECK(err) __check(err, __FILE__, __LINE__)
#define CheckMsg(msg) __checkMsg(msg, __FILE__, __LINE__)

inline void __check(cudaError err, const char* file, const int line)
{
    if (err == cudaSuccess) return;
    fprintf(stderr, "CHECK() Runtime API error in file <%s>, line %i : %s.\n", file, line, cudaGetErrorString(err));
    exit(-1);
}

inline void __checkMsg(const char* msg, const char* file, const int line)
{
    const cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) return;
    fprintf(stderr, "CheckMsg() CUDA error: %s in file <%s>, line %i : %s.\n", msg, file, line, cudaGetErrorString(err));
    exit(-1);
}

// Select (and report) the CUDA device; the ordinal is clamped into the valid range.
inline bool initDevice(int dev)
{
    int count = 0;
    CHECK(cudaGetDeviceCount(&count));
    if (count == 0) { fprintf(stderr, "CUDA error: no devices supporting CUDA.\n"); return false; }
    dev = std::max(0, std::min(dev, count - 1));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 1) { fprintf(stderr, "error: device does not support CUDA.\n"); return false; }
    CHECK(cudaSetDevice(dev));
    int drv = 0, rt = 0;
    CHECK(cudaDriverGetVersion(&drv));
    CHECK(cudaRuntimeGetVersion(&rt));
    fprintf(stderr, "Using Device %d: %s, CUDA Driver Version: %d.%d, Runtime Version: %d.%d\n", dev, prop.name,
            drv / 1000, drv % 1000, rt / 1000, rt % 1000);
    return true;
}

// wall clock in microseconds
inline long long cpuTimer()
{
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// milliseconds since construction, measured with CUDA events on `stream`
class GpuTimer
{
public:
    GpuTimer(cudaStream_t stream_ = 0) : stream(stream_)
    {
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
        cudaEventRecord(start, stream);
    }
    ~GpuTimer() { cudaEventDestroy(start); cudaEventDestroy(stop); }
    float read()
    {
        cudaEventRecord(stop, stream);
        cudaEventSynchronize(stop);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, start, stop);
        return ms;
    }
private:
    cudaEvent_t start, stop;
    cudaStream_t stream;
};

#ifdef __CUDACC__
template <class T> __device__ __inline__ T shiftDown(T var, unsigned int delta, int width = 32) { return __shfl_down_sync(0xffffffff, var, delta, width); }
template <class T> __device__ __inline__ T shiftUp(T var, unsigned int delta, int width = 32) { return __shfl_up_sync(0xffffffff, var, delta, width); }
template <class T> __device__ __inline__ T shuffle(T var, unsigned int lane, int width = 32) { return __shfl_sync(0xffffffff, var, lane, width); }
inline __device__ int __uchar2int(unsigned char data) { return ((data << 23) >> 23); }
inline __device__ int __char2int(signed char data) { return ((data << 24) >> 24); }
#endif

inline int iAlignUp(const int a, const int b) { return (a % b != 0) ? (a - a % b + b) : a; }
inline int iDivUp(int a, int b) { return (a % b != 0) ? (a / b + 1) : (a / b); }
inline int iExp2UpP(const int a)
{
    int p = 0;
    for (int v = 1; v < a; v <<= 1) p++;
    return p;
}
