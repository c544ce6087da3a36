// Dominant orientation and the 486-bit M-LDB descriptor, plus the AoS bridge for the akaze.h shim.
// Reference: gCalcOrient (akazed.cu:1665-1736), dFastAtan2 (:173-185), gDescribe2 (:1869-2001),
// comparison tables (:65-159), AkazePoint layout (akaze_structures.h:19-39).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>
#include <math_constants.h>

using namespace akz;

namespace {

// ---- orientation sample table: the 109 (i,j) of the 13x16 thread grid with i*i+j*j < 36, in thread
// order, and their weights exp(-r2*0.08f) evaluated ON THE DEVICE with the same expf the reference uses
__device__ float g_orient_w[36];            // weight by r2
__device__ float g_orient_wf[36];           // integer pipeline: __expf(-r2 * 0.08f) (akazed.cu:3681)
__device__ signed char g_orient_ij[128][2]; // (i, j) by sample rank; 109 valid
__device__ int g_orient_n;

__global__ void k_orient_table()
{
    if (threadIdx.x < 36) {
        int r2 = threadIdx.x;
        g_orient_w[r2] = exp(-r2 * 0.08f);                       // akazed.cu:1697
        g_orient_wf[r2] = __expf(-r2 * 0.08f);
    }
    if (threadIdx.x == 0) {
        int n = 0;
        for (int t = 0; t < 208; t++) {
            int i = (t & 15) - 6, j = (t / 16) - 6;                // akazed.cu:1692-1693
            if (i * i + j * j < 36) { g_orient_ij[n][0] = (signed char)i; g_orient_ij[n][1] = (signed char)j; n++; }
        }
        g_orient_n = n;
    }
}

__device__ __forceinline__ float fast_atan2(float y, float x)     // akazed.cu:173-185
{
    const float absx = fabsf(x), absy = fabsf(y);
    const float a = __fdiv_rn(fminf(absx, absy), fmaxf(absx, absy));
    const float s = a * a;
    float r = __fmaf_rn(__fmaf_rn(__fmaf_rn(-0.0464964749f, s, 0.15931422f), s, -0.327622764f), s * a, a);
    r = (absy > absx ? 1.5707963267948966f - r : r);
    r = (x < 0 ? (float)(3.14159265358979323846 - r) : r);
    r = (y < 0 ? -r : r);
    return r;
}

// locate keypoint g of the chunk: frame by linear search over the prefix (n <= a few dozen)
__device__ __forceinline__ int find_frame(const int* __restrict__ prefix, int n, int g)
{
    int f = 0;
    while (f + 1 < n && g >= prefix[f + 1]) f++;
    return f;
}

constexpr int ORI_WARPS = 4;
#define AKZ_MAX_FRAMES_SEARCH 256

// one warp per keypoint; bins are summed in ascending sample order => deterministic (App. B-4)
// FAST = true: the integer pipeline's gCalcOrient (akazed.cu:3649-3720): int Lx/Ly planes, __expf weights, and the
// polynomial dFastAtan2 for the bin as well as for the final angle
template <bool FAST>
__global__ void __launch_bounds__(ORI_WARPS * 32) k_orient(const __grid_constant__ AkzLevelTable tab, const int* __restrict__ prefix, int nframes,
                                                          akz_keypoint* __restrict__ kpts, int max_pts)
{
    __shared__ float4 s_samp[ORI_WARPS][128];
    __shared__ float s_res[ORI_WARPS][2][64];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int total = prefix[nframes];
    for (int g = blockIdx.x * ORI_WARPS + wid; g < total; g += gridDim.x * ORI_WARPS) {
        int frame = find_frame(prefix, nframes, g);
        akz_keypoint* kp = kpts + (long long)frame * max_pts + (g - prefix[frame]);
        const AkzLevelDev& L = tab.lv[kp->layer];
        int o = L.octave, p = L.pitch;
        const float* lx = L.lx + (long long)frame * L.plane;
        const float* ly = L.ly + (long long)frame * L.plane;
        int step = (int)__fadd_rn(kp->size, 0.5f);
        int x = (int)__fadd_rn(kp->x, 0.5f) >> o;
        int y = (int)__fadd_rn(kp->y, 0.5f) >> o;
        for (int s = lane; s < 128; s += 32) {
            float4 v = make_float4(0.f, 0.f, __int_as_float(-1), 0.f);
            if (s < 109) {
                int i = g_orient_ij[s][0], j = g_orient_ij[s][1];
                float gw = FAST ? g_orient_wf[i * i + j * j] : g_orient_w[i * i + j * j];
                int yy = min(max(y + step * j, 0), L.h - 1), xx = min(max(x + step * i, 0), L.w - 1);
                long long pos = (long long)yy * p + xx;
                float dx, dy, ang;
                if (FAST) {
                    dx = gw * __ldg(reinterpret_cast<const int*>(lx) + pos);
                    dy = gw * __ldg(reinterpret_cast<const int*>(ly) + pos);
                    ang = fast_atan2(dy, dx);
                } else {
                    dx = gw * __ldg(lx + pos);
                    dy = gw * __ldg(ly + pos);
                    ang = atan2f(dy, dx);
                }
                int a = max(min((int)(ang * (21 / 3.14159265358979323846)) + 21, 41), 0);   // akazed.cu:1702
                v = make_float4(dx, dy, __int_as_float(a), 0.f);
            }
            s_samp[wid][s] = v;
        }
        __syncwarp();
        // lane b owns bins b and b+32
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
        for (int s = 0; s < 109; s++) {
            float4 v = s_samp[wid][s];
            int a = __float_as_int(v.z);
            if (a == lane) { ax = __fadd_rn(ax, v.x); ay = __fadd_rn(ay, v.y); }
            if (a == lane + 32) { bx = __fadd_rn(bx, v.x); by = __fadd_rn(by, v.y); }
        }
        s_res[wid][0][lane] = ax; s_res[wid][1][lane] = ay;
        s_res[wid][0][lane + 32] = bx; s_res[wid][1][lane + 32] = by;
        __syncwarp();
        // sliding window of 7 bins (akazed.cu:1708-1718), two windows per lane (k = lane, lane+32 < 42)
        float best = -1.f, wx0 = 0.f, wy0 = 0.f;
        int bestk = 0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            int k = lane + 32 * half;
            if (k < 42) {
                float sx = s_res[wid][0][k], sy = s_res[wid][1][k];
                for (int q = k + 1; q < k + 7; q++) {
                    int qq = q < 42 ? q : q - 42;
                    sx = __fadd_rn(sx, s_res[wid][0][qq]);
                    sy = __fadd_rn(sy, s_res[wid][1][qq]);
                }
                float r = __fmaf_rn(sx, sx, __fmul_rn(sy, sy));
                if (r > best) { best = r; bestk = k; wx0 = sx; wy0 = sy; }
            }
        }
        // arg-max with "first strictly greater wins" (akazed.cu:1723-1733): larger r, then smaller k
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, d);
            int ok = __shfl_xor_sync(0xffffffffu, bestk, d);
            float ox = __shfl_xor_sync(0xffffffffu, wx0, d);
            float oy = __shfl_xor_sync(0xffffffffu, wy0, d);
            if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; wx0 = ox; wy0 = oy; }
        }
        if (lane == 0) {
            // maxr starts at 0 and the test is '>': when every window is zero the reference keeps k = 0
            if (!(best > 0.f)) {
                float sx = s_res[wid][0][0], sy = s_res[wid][1][0];
                for (int q = 1; q < 7; q++) { sx = __fadd_rn(sx, s_res[wid][0][q]); sy = __fadd_rn(sy, s_res[wid][1][q]); }
                wx0 = sx; wy0 = sy;
            }
            float ang = fast_atan2(wy0, wx0);
            kp->angle = (ang < 0.0f ? (float)(ang + 2.0f * 3.14159265358979323846) : ang);
        }
        __syncwarp();
    }
}

// ---- M-LDB -------------------------------------------------------------------------------------------
__constant__ short c_cmp[2][488];

// ---- M-LDB, pattern size fixed at compile time (PAT = 10 is the reference default, akaze.h:54) ---------------
// Same numbers as k_describe, produced with a third of the instructions (ncu r01b: 2160 thread-instructions per
// thread per keypoint, 18.7 % issue utilisation, stalls on the global gathers and on shared memory):
//   * all 7 x 3 gathers of a thread are issued before the first accumulation (memory-level parallelism 21);
//   * window size and cell geometry are immediates (no runtime division);
//   * the reduction is transposed: thread v owns output value v and evaluates the reference's tree itself,
//     a_t = acc_t + acc_{t+32}, then the shuffle-down pairing ((a0+a1)+(a2+a3))+... serially from shared memory
//     (row stride 65: conflict-free both for the per-thread accumulation columns and for the transposed reads).
template <bool INT> __device__ __forceinline__ float dsum(float a, float b)
{
    return INT ? __int_as_float(__float_as_int(a) + __float_as_int(b)) : __fadd_rn(a, b);
}
template <bool INT> __device__ __forceinline__ bool dgreater(float a, float b)
{
    return INT ? __float_as_int(a) > __float_as_int(b) : a > b;
}

// INT = true: the integer pipeline's M-LDB (gDescribe2 akazed.cu:3723-3855): int planes (bit patterns in the float registers and
// accumulators), sample positions and rotated derivatives in the reference's own float expressions truncated to int, int
// cell sums (associative: any reduction order), int comparisons.
template <int PAT, bool INT>
__global__ void __launch_bounds__(64) k_describe_t(const __grid_constant__ AkzLevelTable tab, const int* __restrict__ prefix, int nframes,
                                                   const akz_keypoint* __restrict__ kpts, unsigned char* __restrict__ desc, int max_pts)
{
    constexpr int S2 = PAT, S3 = (2 * PAT + 2) / 3, S4 = (PAT + 1) / 2;      // akazed.cu:2681-2683 (ceil)
    constexpr int WIN = (3 * S3 > 4 * S4) ? 3 * S3 : 4 * S4;
    constexpr int NS = WIN * WIN, NK = (NS + 63) / 64, RS = 65;
    __shared__ __align__(16) float acc[87 * RS + 4];
    __shared__ float val[96];
    const int tix = threadIdx.x;
    // frame of a keypoint: binary search in a shared-memory copy of the prefix (the linear search through global memory
    // was 14 % of the stall samples, ncu r01f; one block per frame instead leaves the busiest frame as a long tail)
    __shared__ int s_prefix[AKZ_MAX_FRAMES_SEARCH + 1];
    for (int i = tix; i <= nframes && i <= AKZ_MAX_FRAMES_SEARCH; i += 64) s_prefix[i] = prefix[i];
    __syncthreads();
    const int total = s_prefix[min(nframes, AKZ_MAX_FRAMES_SEARCH)];
    // this thread's 8 comparison pairs, fetched once (per-lane different constant addresses serialise in the constant cache)
    unsigned cmp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int b = min(tix, 60) * 8 + i;
        cmp[i] = (unsigned)c_cmp[0][b] | ((unsigned)c_cmp[1][b] << 16);
    }
    // 1. gather of keypoint g into registers (all 7 x 3 loads issued back to back)
    struct Kp { float co, si; int frame, local; };
    auto gather = [&](int g, Kp& K, float (&im)[NK], float (&dx)[NK], float (&dy)[NK]) {
        int lo = 0, hi = nframes - 1;                      // largest f with s_prefix[f] <= g
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_prefix[mid] <= g) lo = mid; else hi = mid - 1; }
        K.frame = lo; K.local = g - s_prefix[lo];
        const akz_keypoint* kp = kpts + (long long)K.frame * max_pts + K.local;
        const AkzLevelDev& L = tab.lv[kp->layer];
        const int o = L.octave, p = L.pitch;
        const float iratio = 1.f / (1 << o);
        const float fscale = (float)(int)__fadd_rn(kp->size, 0.5f);
        const int iscale = (int)(kp->size + 0.5f);
        const float xf = INT ? kp->x * iratio : __fmul_rn(kp->x, iratio), yf = INT ? kp->y * iratio : __fmul_rn(kp->y, iratio);
        const float ang = kp->angle;
        const float co = __cosf(ang), si = __sinf(ang);
        K.co = co; K.si = si;
        const float* imd = L.lt + (long long)K.frame * L.plane;
        const float* dxd = L.lx + (long long)K.frame * L.plane;
        const float* dyd = L.ly + (long long)K.frame * L.plane;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            int i = tix + 64 * k;
            im[k] = 0.f; dx[k] = 0.f; dy[k] = 0.f;
            if (i < NS) {
                int y = i / WIN, x = i - WIN * y;
                float l = (float)(x - S2), kk = (float)(y - S2);
                int xp, yp;
                if (!INT) {
                    xp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(co, kk, -__fmul_rn(si, l)), xf), 0.5f);
                    yp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(si, kk, __fmul_rn(co, l)), yf), 0.5f);
                } else {                                                   // the reference's expressions, as k_describe_int writes them
                    const int li = x - S2, ki = y - S2;
                    xp = (int)(xf + iscale * (ki * co - li * si) + 0.5f);
                    yp = (int)(yf + iscale * (ki * si + li * co) + 0.5f);
                }
                xp = min(max(xp, 0), L.w - 1); yp = min(max(yp, 0), L.h - 1);       // no-op for in-range patches (border test)
                long long pos = (long long)yp * p + xp;
                im[k] = __ldg(imd + pos); dx[k] = __ldg(dxd + pos); dy[k] = __ldg(dyd + pos);
            }
        }
    };
    // Software pipeline: the gathers of the block's NEXT keypoint are issued before the current one is accumulated, so their
    // DRAM latency overlaps the shared-memory work (ncu r01f: 14 % issue utilisation, 8 warps per issue on the long scoreboard
    // with ten 64-thread blocks per SM and a strictly serial gather -> accumulate -> reduce loop per block).
    Kp cur, nxt;
    float im[NK], dx[NK], dy[NK], nim[NK], ndx[NK], ndy[NK];
    if ((int)blockIdx.x < total) gather(blockIdx.x, cur, im, dx, dy);
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        const bool more = g + (int)gridDim.x < total;
        if (more) gather(g + gridDim.x, nxt, nim, ndx, ndy);
        const float co = cur.co, si = cur.si;
        const int frame = cur.frame, local = cur.local;
        // 2. clear the accumulators: the block zeroes the whole array with 128-bit stores (22 per thread instead of 87 scalar)
        for (int i = tix; i < (87 * RS + 3) / 4; i += 64) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        // 3. accumulate in sample order (each thread touches only its own column: no synchronisation needed)
#pragma unroll
        for (int k = 0; k < NK; k++) {
            int i = tix + 64 * k;
            if (i < NS) {
                int y = i / WIN, x = i - WIN * y, m = max(x, y);
                float rx, ry;
                if (!INT) {
                    rx = __fmaf_rn(co, dy[k], -__fmul_rn(si, dx[k]));
                    ry = __fmaf_rn(co, dx[k], __fmul_rn(si, dy[k]));
                } else {
                    const int dxi = __float_as_int(dx[k]), dyi = __float_as_int(dy[k]);
                    const int rxi = -dxi * si + dyi * co, ryi = dxi * co + dyi * si;          // float expressions truncated to int
                    rx = __int_as_float(rxi); ry = __int_as_float(ryi);
                }
                // The cells of the three grids live in disjoint rows of acc, but the compiler cannot know that and would
                // serialise the three read-modify-write groups of a sample (load after the previous group's store: ncu r01j,
                // 6.3 warps per issue on the short scoreboard).  All nine loads first, then the adds, then the stores.
                const bool g2 = m < 2 * S2, g3 = m < 3 * S3, g4 = m < 4 * S4;
                const int x3 = (x < S3 ? 0 : (x < 2 * S3 ? 1 : 2)), y3 = (y < S3 ? 0 : (y < 2 * S3 ? 1 : 2));
                const int x4 = (x < 2 * S4 ? (x < S4 ? 0 : 1) : (x < 3 * S4 ? 2 : 3));
                const int y4 = (y < 2 * S4 ? (y < S4 ? 0 : 1) : (y < 3 * S4 ? 2 : 3));
                float* a2 = acc + (3 * ((y < S2 ? 0 : 1) * 2 + (x < S2 ? 0 : 1))) * RS + tix;
                float* a3 = acc + (3 * (4 + y3 * 3 + x3)) * RS + tix;
                float* a4 = acc + (3 * (13 + y4 * 4 + x4)) * RS + tix;
                float v2[3] = { 0.f, 0.f, 0.f }, v3[3] = { 0.f, 0.f, 0.f }, v4[3] = { 0.f, 0.f, 0.f };
                if (g2) { v2[0] = a2[0]; v2[1] = a2[RS]; v2[2] = a2[2 * RS]; }
                if (g3) { v3[0] = a3[0]; v3[1] = a3[RS]; v3[2] = a3[2 * RS]; }
                if (g4) { v4[0] = a4[0]; v4[1] = a4[RS]; v4[2] = a4[2 * RS]; }
                if (g2) { a2[0] = dsum<INT>(v2[0], im[k]); a2[RS] = dsum<INT>(v2[1], rx); a2[2 * RS] = dsum<INT>(v2[2], ry); }
                if (g3) { a3[0] = dsum<INT>(v3[0], im[k]); a3[RS] = dsum<INT>(v3[1], rx); a3[2 * RS] = dsum<INT>(v3[2], ry); }
                if (g4) { a4[0] = dsum<INT>(v4[0], im[k]); a4[RS] = dsum<INT>(v4[1], rx); a4[2 * RS] = dsum<INT>(v4[2], ry); }
            }
        }
        __syncthreads();
        // 4. transposed reduction: the reference's tree for value v, evaluated by one thread
        for (int v = tix; v < 87; v += 64) {
            const float* a = acc + v * RS;
            float r[32];
#pragma unroll
            for (int t = 0; t < 32; t++) r[t] = dsum<INT>(a[t], a[t + 32]);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
#pragma unroll
                for (int t = 0; t + d < 32; t += 2 * d) r[t] = dsum<INT>(r[t], r[t + d]);
            val[v] = r[0];
        }
        __syncthreads();
        unsigned char* out = desc + ((long long)frame * max_pts + local) * 64;
        unsigned rbits = 0;
        if (tix < 61) {
            const int nb = (tix == 60 ? 6 : 8);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < nb) rbits |= (dgreater<INT>(val[cmp[i] & 0xFFFFu], val[cmp[i] >> 16]) ? 1u : 0u) << i;
        }
        out[tix] = (unsigned char)rbits;                  // bytes 61..63 are written as zero
        __syncthreads();
        if (more) {
            cur = nxt;
#pragma unroll
            for (int k = 0; k < NK; k++) { im[k] = nim[k]; dx[k] = ndx[k]; dy[k] = ndy[k]; }
        }
    }
}

// ---- M-LDB, sample-major: k_describe_s ----------------------------------------------------------------------------------------
// The reference accumulates the (2 P + 1)^2 samples of a keypoint with 64 threads -- thread t takes samples t, t + 64, ... in
// order and adds each to the cell of the 2x2, 3x3 and 4x4 grid it falls into -- and reduces every cell value over the 64
// threads with a fixed tree: a_t = acc_t + acc_(t+32), then pairs, fours, ... (gDescribe2 akazed.cu:1869-2001).  The float
// sums pin that order.  WHICH samples of which thread fall into which cell is static (sample i sits at column i mod W, row
// i / W of the unrotated grid), so the whole reduction is a fixed expression over the sample values.  k_describe_s evaluates
// it directly: 128 threads gather the samples into a sample-major shared array (5.3 KB per channel set at P = 10, not the
// 87 x 64 accumulator matrix: 22.6 KB, read-modify-write per sample, zero fill, 64 reads per output value), then four threads
// per cell walk a table of 16-bit masks -- bit m of mask[cell][lane] = sample lane + 64 m belongs to the cell -- rebuild the per-
// lane sums in sample order, add the two halves, run the tree over their eight lanes in registers and join by shuffle.
// ~1240 sample reads per channel instead of ~8200 shared-memory operations per keypoint; the same bits.
// Any pattern size with (2 P + 1)^2 <= 1024 samples (P <= 14); the tables are built on the host for the context's pattern.
constexpr int DS_NT = 128;
constexpr int DS_CELLS = 29;                                 // 4 + 9 + 16
constexpr int DS_MAXNS = 1024;                               // mask bits: sample = lane + 64 m, m < 16

struct DescTables {
    unsigned short mask[DS_CELLS][4][16];                    // [cell][quarter q][slot k]: lane = 8 q + (k >> 1) + 32 (k & 1)
};

struct DescArgs {
    const int* prefix;
    const akz_keypoint* kpts;
    unsigned char* desc;
    const DescTables* tables;
    int nframes, max_pts;
    int s2, s3, s4, win, ns;
};

// sample i lives at ds_val[i + (i >> 5)]: the reduction reads samples 64 m apart (lane + 64 m), which would all sit in one bank
__device__ __forceinline__ int dsi(int i) { return i + (i >> 5); }

template <bool INT, int NK>
__global__ void __launch_bounds__(DS_NT, (NK <= 4 ? 8 : 4)) k_describe_s(const __grid_constant__ AkzLevelTable tab, const __grid_constant__ DescArgs a)
{
    extern __shared__ __align__(16) float ds_val[];          // [3][nsp]: im, rx, ry of every sample
    __shared__ __align__(16) DescTables s_tab;
    __shared__ float s_cell[96];
    __shared__ int s_prefix[AKZ_MAX_FRAMES_SEARCH + 1];
    const int tid = threadIdx.x;
    const int nsp = (a.ns + (a.ns >> 5) + 4) & ~3;
    for (int i = tid; i <= a.nframes && i <= AKZ_MAX_FRAMES_SEARCH; i += DS_NT) s_prefix[i] = a.prefix[i];
    for (int i = tid; i < (int)(sizeof(DescTables) / 4); i += DS_NT) reinterpret_cast<unsigned*>(&s_tab)[i] = reinterpret_cast<const unsigned*>(a.tables)[i];
    __syncthreads();
    const int total = s_prefix[min(a.nframes, AKZ_MAX_FRAMES_SEARCH)];
    // this thread's 8 comparison pairs, fetched once (per-lane different constant addresses serialise in the constant cache)
    unsigned cmp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int b = min(tid, 60) * 8 + i;
        cmp[i] = (unsigned)c_cmp[0][b] | ((unsigned)c_cmp[1][b] << 16);
    }
    const int S2 = a.s2, WIN = a.win, NS = a.ns;
    struct Kp { float co, si; int frame, local; };
    // gather of keypoint g into registers: all NK x 3 loads are issued back to back
    auto gather = [&](int g, Kp& K, float (&im)[NK], float (&dx)[NK], float (&dy)[NK]) {
        int lo = 0, hi = a.nframes - 1;                    // largest f with s_prefix[f] <= g
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_prefix[mid] <= g) lo = mid; else hi = mid - 1; }
        K.frame = lo; K.local = g - s_prefix[lo];
        const akz_keypoint* kp = a.kpts + (long long)K.frame * a.max_pts + K.local;
        const AkzLevelDev& L = tab.lv[kp->layer];
        const int o = L.octave, p = L.pitch;
        const float iratio = 1.f / (1 << o);
        const float fscale = (float)(int)__fadd_rn(kp->size, 0.5f);
        const int iscale = (int)(kp->size + 0.5f);
        const float xf = INT ? kp->x * iratio : __fmul_rn(kp->x, iratio), yf = INT ? kp->y * iratio : __fmul_rn(kp->y, iratio);
        const float ang = kp->angle;
        const float co = __cosf(ang), si = __sinf(ang);
        K.co = co; K.si = si;
        const float* imd = L.lt + (long long)K.frame * L.plane;
        const float* dxd = L.lx + (long long)K.frame * L.plane;
        const float* dyd = L.ly + (long long)K.frame * L.plane;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            int i = tid + DS_NT * k;
            im[k] = 0.f; dx[k] = 0.f; dy[k] = 0.f;
            if (i < NS) {
                int y = i / WIN, x = i - WIN * y;
                int xp, yp;
                if (!INT) {
                    float l = (float)(x - S2), kk = (float)(y - S2);
                    xp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(co, kk, -__fmul_rn(si, l)), xf), 0.5f);
                    yp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(si, kk, __fmul_rn(co, l)), yf), 0.5f);
                } else {                                   // integer pipeline: the reference's own float expressions truncated to int
                    const int li = x - S2, ki = y - S2;
                    xp = (int)(xf + iscale * (ki * co - li * si) + 0.5f);
                    yp = (int)(yf + iscale * (ki * si + li * co) + 0.5f);
                }
                xp = min(max(xp, 0), L.w - 1); yp = min(max(yp, 0), L.h - 1);       // no-op for in-range patches (border test)
                long long pos = (long long)yp * p + xp;
                im[k] = __ldg(imd + pos); dx[k] = __ldg(dxd + pos); dy[k] = __ldg(dyd + pos);
            }
        }
    };
    // the gathers of the block's NEXT keypoint are in flight while the current one is reduced
    Kp cur, nxt;
    float im[NK], dx[NK], dy[NK], nim[NK], ndx[NK], ndy[NK];
    if ((int)blockIdx.x < total) gather(blockIdx.x, cur, im, dx, dy);
    // role of this thread in the reduction: quarter q of cell
    const int cell = min(tid >> 2, DS_CELLS - 1), q = tid & 3;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        const bool more = g + (int)gridDim.x < total;
        if (more) gather(g + gridDim.x, nxt, nim, ndx, ndy);
        const float co = cur.co, si = cur.si;
        // 1. rotated derivatives, sample-major into shared memory
#pragma unroll
        for (int k = 0; k < NK; k++) {
            int i = tid + DS_NT * k;
            if (i < NS) {
                float rx, ry;
                if (!INT) {
                    rx = __fmaf_rn(co, dy[k], -__fmul_rn(si, dx[k]));
                    ry = __fmaf_rn(co, dx[k], __fmul_rn(si, dy[k]));
                } else {
                    const int dxi = __float_as_int(dx[k]), dyi = __float_as_int(dy[k]);
                    const int rxi = -dxi * si + dyi * co, ryi = dxi * co + dyi * si;          // float expressions truncated to int
                    rx = __int_as_float(rxi); ry = __int_as_float(ryi);
                }
                const int si_ = dsi(i);
                ds_val[si_] = im[k]; ds_val[nsp + si_] = rx; ds_val[2 * nsp + si_] = ry;
            }
        }
        __syncthreads();
        // 2. the reference's reduction for this thread's eight lane pairs of its cell
        {
            const uint4 w0 = *reinterpret_cast<const uint4*>(&s_tab.mask[cell][q][0]);
            const uint4 w1 = *reinterpret_cast<const uint4*>(&s_tab.mask[cell][q][8]);
            const unsigned mw[8] = { w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w };       // two 16-bit masks per word: lane tt (low), lane tt + 32 (high)
            float r[8][3];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int tt = 8 * q + j;
                float plo[3] = { 0.f, 0.f, 0.f }, phi[3] = { 0.f, 0.f, 0.f };
                unsigned mk = mw[j] & 0xFFFFu;
                while (mk) {                                              // samples tt + 64 m of lane tt, in order
                    const int m = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int id = dsi(tt + 64 * m);
                    plo[0] = dsum<INT>(plo[0], ds_val[id]); plo[1] = dsum<INT>(plo[1], ds_val[nsp + id]); plo[2] = dsum<INT>(plo[2], ds_val[2 * nsp + id]);
                }
                mk = mw[j] >> 16;
                while (mk) {                                              // lane tt + 32
                    const int m = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int id = dsi(tt + 32 + 64 * m);
                    phi[0] = dsum<INT>(phi[0], ds_val[id]); phi[1] = dsum<INT>(phi[1], ds_val[nsp + id]); phi[2] = dsum<INT>(phi[2], ds_val[2 * nsp + id]);
                }
#pragma unroll
                for (int c = 0; c < 3; c++) r[j][c] = dsum<INT>(plo[c], phi[c]);              // a_t = acc_t + acc_(t+32)
            }
#pragma unroll
            for (int d = 1; d < 8; d <<= 1)
#pragma unroll
                for (int j = 0; j + d < 8; j += 2 * d)
#pragma unroll
                    for (int c = 0; c < 3; c++) r[j][c] = dsum<INT>(r[j][c], r[j + d][c]);
            // lanes 8 q .. 8 q + 7 are done; the quarters of a cell sit in four consecutive threads: steps d = 8 and d = 16 of the tree
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float v = r[0][c];
                v = dsum<INT>(v, __shfl_down_sync(0xffffffffu, v, 1));        // q0 + q1, q2 + q3
                v = dsum<INT>(v, __shfl_down_sync(0xffffffffu, v, 2));        // (q0 + q1) + (q2 + q3)
                if (q == 0 && tid < 4 * DS_CELLS) s_cell[3 * cell + c] = v;
            }
        }
        __syncthreads();
        // 3. the 486 comparisons
        if (tid < 64) {
            unsigned char* out = a.desc + ((long long)cur.frame * a.max_pts + cur.local) * 64;
            unsigned rbits = 0;
            if (tid < 61) {
                const int nb = (tid == 60 ? 6 : 8);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (i < nb) rbits |= (dgreater<INT>(s_cell[cmp[i] & 0xFFFFu], s_cell[cmp[i] >> 16]) ? 1u : 0u) << i;
            }
            out[tid] = (unsigned char)rbits;                  // bytes 61..63 are written as zero
        }
        if (more) {
            cur = nxt;
#pragma unroll
            for (int k = 0; k < NK; k++) { im[k] = nim[k]; dx[k] = ndx[k]; dy[k] = ndy[k]; }
        }
    }
}

// masks of the static reduction for a pattern (host)
static void build_desc_tables(int s2, int s3, int s4, int win, DescTables& T)
{
    memset(&T, 0, sizeof(T));
    const int ns = win * win;
    for (int i = 0; i < ns; i++) {
        const int y = i / win, x = i - win * y, m = x > y ? x : y;
        const int lane = i & 63, mm = i >> 6;
        const int qd = (lane & 31) >> 3, k = 2 * (lane & 7) + (lane >> 5);
        auto set = [&](int cell) { T.mask[cell][qd][k] |= (unsigned short)(1u << mm); };
        if (m < 2 * s2) set((y < s2 ? 0 : 1) * 2 + (x < s2 ? 0 : 1));
        if (m < 3 * s3) set(4 + (y < s3 ? 0 : (y < 2 * s3 ? 1 : 2)) * 3 + (x < s3 ? 0 : (x < 2 * s3 ? 1 : 2)));
        if (m < 4 * s4) set(13 + (y < 2 * s4 ? (y < s4 ? 0 : 1) : (y < 3 * s4 ? 2 : 3)) * 4 + (x < 2 * s4 ? (x < s4 ? 0 : 1) : (x < 3 * s4 ? 2 : 3)));
    }
}

// ---- AoS bridge: reference AkazePoint (104 B) ------------------------------------------------------------
struct RefPoint {
    float x, y; int octave; float response, size, angle;
    unsigned char features[61];
    int match, distance; float match_x, match_y;
};
static_assert(sizeof(RefPoint) == 104, "AkazePoint layout (SURVEY App. C)");

__global__ void k_pack(const int* __restrict__ count, const akz_keypoint* __restrict__ kpts, const unsigned char* __restrict__ desc,
                       RefPoint* __restrict__ pts, int max_pts, int with_desc)
{
    int n = min(*count, max_pts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        akz_keypoint k = kpts[i];
        RefPoint* p = pts + i;
        p->x = k.x; p->y = k.y; p->octave = k.layer; p->size = k.size; p->angle = k.angle;
        if (with_desc) {
            // 61 bytes = 15 words + 1 byte; features[] starts at offset 24 of a 104-byte record: word stores are aligned.
            // Bytes 85..87 (padding before `match`) are left alone.
            const uint4* d4 = reinterpret_cast<const uint4*>(desc + (long long)i * 64);
            const uint4 a = __ldg(d4), b = __ldg(d4 + 1), c = __ldg(d4 + 2), d = __ldg(d4 + 3);
            unsigned* f = reinterpret_cast<unsigned*>(p->features);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
            f[8] = c.x; f[9] = c.y; f[10] = c.z; f[11] = c.w; f[12] = d.x; f[13] = d.y; f[14] = d.z;
            p->features[60] = (unsigned char)(d.w & 0xFFu);
        }
    }
}

__global__ void k_unpack(const RefPoint* __restrict__ pts, int n, unsigned char* __restrict__ desc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int b = 0; b < 61; b++) desc[(long long)i * 64 + b] = pts[i].features[b];
    desc[(long long)i * 64 + 61] = 0; desc[(long long)i * 64 + 62] = 0; desc[(long long)i * 64 + 63] = 0;
}

__global__ void k_scatter(const akz_match_t* __restrict__ m, int nq, RefPoint* __restrict__ pq, const RefPoint* __restrict__ pt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    akz_match_t r = m[i];
    if (r.idx1 >= 0) { pq[i].match = r.idx1; pq[i].distance = r.dist1; pq[i].match_x = pt[r.idx1].x; pq[i].match_y = pt[r.idx1].y; }
    else { pq[i].match = -1; pq[i].distance = -1; pq[i].match_x = -1.f; pq[i].match_y = -1.f; }     // akazed.cu:2231-2237
}

akz_once_t g_cmp_uploaded;                // per device: constant memory is per device

}  // namespace

namespace akzk {

int orient_table_init(cudaStream_t st)
{
    k_orient_table<<<1, 64, 0, st>>>();
    if (akz_once_guard once{g_cmp_uploaded}) {
        int c1[488], c2[488];
        short h[2][488];
        akz_compare_indices(c1, c2);
        for (int i = 0; i < 488; i++) { h[0][i] = (short)(i < 486 ? c1[i] : 0); h[1][i] = (short)(i < 486 ? c2[i] : 0); }
        cudaMemcpyToSymbolAsync(c_cmp, h, sizeof(h), 0, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
    }
    return 1;
}

int orient(cudaStream_t st, const AkzLevelTable& tab, const int* counts, const int* prefix, akz_keypoint* kpts, int max_pts, int n, int fast)
{
    (void)counts;
    if (fast) k_orient<true><<<148 * 8, ORI_WARPS * 32, 0, st>>>(tab, prefix, n, kpts, max_pts);
    else k_orient<false><<<148 * 8, ORI_WARPS * 32, 0, st>>>(tab, prefix, n, kpts, max_pts);
    return 1;
}

// per-device cache of the reduction tables of a pattern size (tiny, built on first use; never freed)
static const DescTables* desc_tables_for(int pattern, int s2, int s3, int s4, int win)
{
    static std::mutex mu;
    static std::vector<std::pair<long long, const DescTables*>> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const long long key = (long long)dev * 1000 + pattern;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache) if (e.first == key) return e.second;
    DescTables h;
    build_desc_tables(s2, s3, s4, win, h);
    DescTables* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(DescTables)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, &h, sizeof(DescTables), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    cache.push_back({ key, d });
    return d;
}

// warm the table cache (akz_create): no allocation or blocking copy later, e.g. inside a stream capture
int describe_prepare(int pattern)
{
    const int s2 = pattern, s3 = (int)ceilf(2.0f * pattern / 3.0f), s4 = (int)ceilf(0.5f * pattern);
    const int win = std::max(3 * s3, 4 * s4);
    if (win * win > DS_MAXNS) return akz_set_error(AKZ_E_UNSUPPORTED, "descriptor_pattern_size %d: (%d x %d samples) exceeds the %d the M-LDB kernel supports", pattern, win, win, DS_MAXNS);
    return desc_tables_for(pattern, s2, s3, s4, win) ? 0 : akz_set_error(AKZ_E_NOMEM, "descriptor tables");
}

int describe(cudaStream_t st, const AkzLevelTable& tab, const int* counts, const int* prefix, const akz_keypoint* kpts,
             unsigned char* desc, int max_pts, int n, int pattern, int fast)
{
    (void)counts;
    const int s2 = pattern;                                       // akazed.cu:2681-2683
    const int s3 = (int)ceilf(2.0f * pattern / 3.0f);
    const int s4 = (int)ceilf(0.5f * pattern);
    const int win = std::max(3 * s3, 4 * s4), ns = win * win;
    if (ns > DS_MAXNS) return akz_set_error(AKZ_E_UNSUPPORTED, "descriptor_pattern_size %d is not supported (at most %d samples)", pattern, DS_MAXNS);
    if (n > AKZ_MAX_FRAMES_SEARCH) return akz_set_error(AKZ_E_UNSUPPORTED, "more than %d frames per chunk", AKZ_MAX_FRAMES_SEARCH);
    DescArgs a;
    a.prefix = prefix; a.kpts = kpts; a.desc = desc; a.nframes = n; a.max_pts = max_pts;
    a.s2 = s2; a.s3 = s3; a.s4 = s4; a.win = win; a.ns = ns;
    a.tables = desc_tables_for(pattern, s2, s3, s4, win);
    if (!a.tables) return akz_set_error(AKZ_E_NOMEM, "descriptor tables");
    const size_t smem = 3 * (size_t)((ns + (ns >> 5) + 4) & ~3) * sizeof(float);
    const int grid = 148 * 8;
    if (ns <= 4 * DS_NT) {
        if (fast) k_describe_s<true, 4><<<grid, DS_NT, smem, st>>>(tab, a);
        else k_describe_s<false, 4><<<grid, DS_NT, smem, st>>>(tab, a);
    } else {
        if (fast) k_describe_s<true, 8><<<grid, DS_NT, smem, st>>>(tab, a);
        else k_describe_s<false, 8><<<grid, DS_NT, smem, st>>>(tab, a);
    }
    return 1;
}

int pack_points(cudaStream_t st, const int* count, const akz_keypoint* kpts, const unsigned char* desc, void* points, int max_pts, int with_desc)
{
    k_pack<<<64, 256, 0, st>>>(count, kpts, desc, (RefPoint*)points, max_pts, with_desc);
    return 1;
}

int unpack_desc(cudaStream_t st, const void* points, int n, unsigned char* desc)
{
    if (n <= 0) return 0;
    k_unpack<<<(n + 255) / 256, 256, 0, st>>>((const RefPoint*)points, n, desc);
    return 1;
}

int scatter_matches(cudaStream_t st, const akz_match_t* m, int nq, void* pq, const void* pt)
{
    if (nq <= 0) return 0;
    k_scatter<<<(nq + 255) / 256, 256, 0, st>>>(m, nq, (RefPoint*)pq, (const RefPoint*)pt);
    return 1;
}

}  // namespace akzk
