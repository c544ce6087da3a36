// Dominant orientation and the 486-bit M-LDB descriptor, plus the AoS bridge for the akaze.h shim.
// Reference: gCalcOrient (akazed.cu:1665-1736), dFastAtan2 (:173-185), gDescribe2 (:1869-2001),
// comparison tables (:65-159), AkazePoint layout (akaze_structures.h:19-39).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
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

// Gathers of the keypoint stages.  ncu (profiles/r02n_*): k_orient and both M-LDB kernels sit at the same 1.1 TB/s of DRAM reads
// whatever their instruction count: scattered 32-byte sector fetches bound them, not bytes.  The patches of neighbouring
// keypoints cover every sector of a row segment sooner or later, so a miss asks L2 to fetch the whole 256-byte neighbourhood.
__device__ __forceinline__ float ldg_wide(const float* p)
{
    float v;
    asm("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// locate keypoint g of the chunk: frame by linear search over the prefix (n <= a few dozen)
__device__ __forceinline__ int find_frame(const int* __restrict__ prefix, int n, int g)
{
    int f = 0;
    while (f + 1 < n && g >= prefix[f + 1]) f++;
    return f;
}

// ---- processing order of the keypoint stages ------------------------------------------------------------------------------
// Keypoints are stored in raster order (App. B-5); consecutive ones sit on different levels, i.e. in different planes.  Gathering
// in that order thrashes: scripts/probes/gather_probe.cu -- the M-LDB access pattern alone, 19 k keypoints per frame -- takes
// 3.1 ns per keypoint when they share one level, 7-9 ns when twelve levels alternate.  order[frame][j] lists the frame's
// keypoints grouped by layer (counting sort, one block per frame; within a layer raster order is kept from one group of 1024 to
// the next, which is all locality needs).  Orientation and descriptors walk the frame through it; results and their positions
// in the output do not depend on it.
__global__ void __launch_bounds__(1024) k_layer_order(const int* __restrict__ counts, const akz_keypoint* __restrict__ kpts, int* __restrict__ order, int max_pts)
{
    __shared__ int s_cnt[AKZ_MAX_LEVELS], s_cur[AKZ_MAX_LEVELS];
    const int frame = blockIdx.x, n = min(counts[frame], max_pts);
    const akz_keypoint* kp = kpts + (long long)frame * max_pts;
    int* ord = order + (long long)frame * max_pts;
    if (threadIdx.x < AKZ_MAX_LEVELS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) atomicAdd(&s_cnt[kp[i].layer], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int l = 0; l < AKZ_MAX_LEVELS; l++) { s_cur[l] = sum; sum += s_cnt[l]; }
    }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        if (i < n) ord[atomicAdd(&s_cur[kp[i].layer], 1)] = i;
        __syncthreads();
    }
}

constexpr int ORI_WARPS = 4;
#define AKZ_MAX_FRAMES_SEARCH 256

// one warp per keypoint; bins are summed in ascending sample order => deterministic (App. B-4)
// FAST = true: the integer pipeline's gCalcOrient (akazed.cu:3649-3720): int Lx/Ly planes, __expf weights, and the
// polynomial dFastAtan2 for the bin as well as for the final angle
// Bin sums.  Lane l holds samples l, l + 32, l + 64, l + 96 (109 valid).  Round by round the lanes whose samples fall into the
// same bin find each other (__match_any_sync) and every one of them adds the group's values in lane order on top of the bin's
// running sum; the lowest lane writes it back.  Rounds ascend and lanes ascend within a round: exactly the ascending-sample
// order (the first version had every lane scan all 109 samples for its two bins: 900 of its 1570 warp instructions per
// keypoint, ncu r02n), same bits.
template <bool FAST>
__global__ void __launch_bounds__(ORI_WARPS * 32) k_orient(const __grid_constant__ AkzLevelTable tab, const int* __restrict__ prefix, int nframes,
                                                          akz_keypoint* __restrict__ kpts, int max_pts, const int* __restrict__ order)
{
    __shared__ float s_res[ORI_WARPS][2][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int total = prefix[nframes];
    // this lane's four samples: offsets and weight (constant for the kernel)
    int si[4], sj[4];
    float sw[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int sidx = lane + 32 * r;
        si[r] = 0; sj[r] = 0; sw[r] = 0.f;
        if (sidx < 109) {
            si[r] = g_orient_ij[sidx][0]; sj[r] = g_orient_ij[sidx][1];
            sw[r] = FAST ? g_orient_wf[si[r] * si[r] + sj[r] * sj[r]] : g_orient_w[si[r] * si[r] + sj[r] * sj[r]];
        }
    }
    for (int g = blockIdx.x * ORI_WARPS + wid; g < total; g += gridDim.x * ORI_WARPS) {
        int frame = find_frame(prefix, nframes, g);
        int local = g - prefix[frame];
        if (order) local = order[(long long)frame * max_pts + local];
        akz_keypoint* kp = kpts + (long long)frame * max_pts + local;
        const AkzLevelDev& L = tab.lv[kp->layer];
        int o = L.octave, p = L.pitch;
        const float* lx = L.lx + (long long)frame * L.plane;
        const float* ly = L.ly + (long long)frame * L.plane;
        int step = (int)__fadd_rn(kp->size, 0.5f);
        int x = (int)__fadd_rn(kp->x, 0.5f) >> o;
        int y = (int)__fadd_rn(kp->y, 0.5f) >> o;
        float gx[4], gy[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {                       // all eight gathers in flight
            int yy = min(max(y + step * sj[r], 0), L.h - 1), xx = min(max(x + step * si[r], 0), L.w - 1);
            long long pos = (long long)yy * p + xx;
            gx[r] = ldg_wide(lx + pos); gy[r] = ldg_wide(ly + pos);
        }
        s_res[wid][0][lane] = 0.f; s_res[wid][0][lane + 32] = 0.f; s_res[wid][1][lane] = 0.f; s_res[wid][1][lane + 32] = 0.f;
        float dx[4], dy[4];
        int bin[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float ang;
            if (FAST) {
                dx[r] = sw[r] * __float_as_int(gx[r]);
                dy[r] = sw[r] * __float_as_int(gy[r]);
                ang = fast_atan2(dy[r], dx[r]);
            } else {
                dx[r] = sw[r] * gx[r];
                dy[r] = sw[r] * gy[r];
                ang = atan2f(dy[r], dx[r]);
            }
            bin[r] = max(min((int)(ang * (21 / 3.14159265358979323846)) + 21, 41), 0);   // akazed.cu:1702
            if (lane + 32 * r >= 109) bin[r] = 64 + lane;                                // no sample: a group of its own, never stored
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const bool valid = bin[r] < 64;
            const unsigned peers = __match_any_sync(0xffffffffu, bin[r]);
            const int maxc = __reduce_max_sync(0xffffffffu, valid ? __popc(peers) : 0);
            float ax = 0.f, ay = 0.f;
            if (valid) { ax = s_res[wid][0][bin[r]]; ay = s_res[wid][1][bin[r]]; }
            unsigned rem = peers;
            for (int k = 0; k < maxc; k++) {
                const int src = rem ? __ffs(rem) - 1 : lane;
                const float vx = __shfl_sync(0xffffffffu, dx[r], src), vy = __shfl_sync(0xffffffffu, dy[r], src);
                if (rem) { ax = __fadd_rn(ax, vx); ay = __fadd_rn(ay, vy); }
                rem &= rem - 1;
            }
            __syncwarp();                                    // every lane has read the bin's running sum
            if (valid && lane == __ffs(peers) - 1) { s_res[wid][0][bin[r]] = ax; s_res[wid][1][bin[r]] = ay; }
            __syncwarp();
        }
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
                float rr = __fmaf_rn(sx, sx, __fmul_rn(sy, sy));
                if (rr > best) { best = rr; bestk = k; wx0 = sx; wy0 = sy; }
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

// float / integer pipeline arithmetic of the cell sums: INT = true is the integer pipeline's M-LDB (gDescribe2 akazed.cu:3723-3855):
// int planes (bit patterns travel in float registers), sample positions and rotated derivatives in the reference's own float
// expressions truncated to int, int cell sums (associative), int comparisons.
template <bool INT> __device__ __forceinline__ float dsum(float a, float b)
{
    return INT ? __int_as_float(__float_as_int(a) + __float_as_int(b)) : __fadd_rn(a, b);
}
template <bool INT> __device__ __forceinline__ bool dgreater(float a, float b)
{
    return INT ? __float_as_int(a) > __float_as_int(b) : a > b;
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
    const int* order;                                        // processing order within a frame (k_layer_order), or null
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
        if (a.order) K.local = a.order[(long long)K.frame * a.max_pts + K.local];
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
                im[k] = ldg_wide(imd + pos); dx[k] = ldg_wide(dxd + pos); dy[k] = ldg_wide(dyd + pos);
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

constexpr int DW_RS = 64;                                    // floats per row of the (cell value x reference thread) matrix: bank = thread mod 32
constexpr int DW_ROWS = 90;                                  // 87 values + 3 trash rows for samples outside a grid

// ---- M-LDB, block per keypoint with register run sums: k_describe_b (pattern size 10) ---------------------------------------
// ncu of k_describe_s on the keypoint-heavy input (profiles/r02n_k_describe_s_raw.txt): 6950 warp instructions per keypoint at
// 61 % issue utilisation -- its table-driven reduction walks masks with divergent loops.  Here everything that depends on a
// sample's place in the 21 x 21 grid is a per-thread constant computed once per kernel:
//   gather:     as k_describe_s -- 128 threads gather the 441 samples (the NEXT keypoint's gathers are in flight while the
//               current one is reduced) and store them sample-major, rotated, in shared memory;
//   accumulate: reference thread t's samples (t, t + 64, ...) visit every cell of a grid in ONE contiguous run (x advances by 1
//               and y by 3 per sample, with at most one wrap), so the per-(cell, thread) sums are running sums in registers,
//               stored once when the run ends into an 87 x 64 matrix in shared memory (row = 3 cell + channel, column =
//               reference thread: conflict-free; samples outside a grid go to three trash rows).  Threads 0..63 take grids
//               2 x 2 and 3 x 3, threads 64..127 grid 4 x 4.  Starting every run from zero is exact (0 + s = s);
//   reduce:     thread v < 87 reads row v (16 LDS.128, float4 index XOR v so that the eight threads of a quarter warp hit
//               different banks), writes zeros back (the matrix is clean for the next keypoint) and runs the reference's tree
//               in registers: a_t = acc_t + acc_(t+32), pairs, fours, ...  The tree is a butterfly, invariant under an XOR
//               permutation of its leaves, and fadd commutes: same bits;
//   compare:    61 threads form one descriptor byte each.
// What is left is the gather itself (scripts/probes/gather_probe.cu: the access pattern alone costs 3-4 ns per keypoint; with the
// samples forced onto one pixel the kernel runs four times faster): misses fetch 256-byte neighbourhoods (ldg_wide) and the
// frame is walked level by level (k_layer_order).  Five CTAs per SM: six need register spills that cost more than they hide.
constexpr int DB_NSP = 448;                                  // floats per channel of the sample array
constexpr int DB_CTAS = 5;                                   // CTAs per SM (102 registers: no spills; 6 spills the row reduce)
constexpr size_t DB_SMEM = (size_t)(3 * DB_NSP + DW_ROWS * DW_RS + 96 + AKZ_MAX_FRAMES_SEARCH + 4) * 4;

template <bool INT>
__global__ void __launch_bounds__(DS_NT, DB_CTAS) k_describe_b(const __grid_constant__ AkzLevelTable tab, const __grid_constant__ DescArgs a)
{
    constexpr int P = 10, S3 = 7, S4 = 5, WIN = 21, NS = WIN * WIN, NK = 4;
    extern __shared__ __align__(16) float db_smem[];
    float* ds_val = db_smem;                                 // [3][DB_NSP]: im, rx, ry of every sample
    float* s_acc = ds_val + 3 * DB_NSP;                      // [90][64]
    float* s_cell = s_acc + DW_ROWS * DW_RS;                 // [96]
    int* s_prefix = reinterpret_cast<int*>(s_cell + 96);
    const int tid = threadIdx.x;
    for (int i = tid; i < DW_ROWS * DW_RS / 4; i += DS_NT) reinterpret_cast<float4*>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i <= a.nframes && i <= AKZ_MAX_FRAMES_SEARCH; i += DS_NT) s_prefix[i] = a.prefix[i];
    __syncthreads();
    const int total = s_prefix[min(a.nframes, AKZ_MAX_FRAMES_SEARCH)];
    unsigned cmp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int b = min(tid, 60) * 8 + i;
        cmp[i] = (unsigned)c_cmp[0][b] | ((unsigned)c_cmp[1][b] << 16);
    }
    // reference thread t = tid & 63; threads 0..63 own grids 2 x 2 and 3 x 3 (offA, offB), threads 64..127 grid 4 x 4 (offA)
    const int col = tid & 63, hi = tid >> 6;
    unsigned offs[7];
    unsigned lastA = 0, lastB = 0;
#pragma unroll
    for (int m = 0; m < 7; m++) {
        const int i = col + 64 * m;
        const int y = i / WIN, x = i - WIN * y, mx = max(x, y);
        const bool in = i < NS;
        const int r2 = (in && mx < 2 * P) ? 3 * ((y < P ? 0 : 1) * 2 + (x < P ? 0 : 1)) : 87;
        const int r3 = in ? 3 * (4 + (y < S3 ? 0 : (y < 2 * S3 ? 1 : 2)) * 3 + (x < S3 ? 0 : (x < 2 * S3 ? 1 : 2))) : 87;
        const int r4 = (in && mx < 4 * S4) ? 3 * (13 + (y < 2 * S4 ? (y < S4 ? 0 : 1) : (y < 3 * S4 ? 2 : 3)) * 4 + (x < 2 * S4 ? (x < S4 ? 0 : 1) : (x < 3 * S4 ? 2 : 3))) : 87;
        offs[m] = hi ? (unsigned)(r4 * DW_RS + col) : ((unsigned)(r2 * DW_RS + col) | ((unsigned)(r3 * DW_RS + col) << 16));
    }
#pragma unroll
    for (int m = 0; m < 7; m++) {
        const int mn = m == 6 ? m : m + 1;
        if (m == 6 || (offs[mn] & 0xFFFFu) != (offs[m] & 0xFFFFu)) lastA |= 1u << m;
        if (m == 6 || (offs[mn] >> 16) != (offs[m] >> 16)) lastB |= 1u << m;
    }
    struct Kp { float co, si; int frame, local; };
    auto gather = [&](int g, Kp& K, float (&im)[NK], float (&dx)[NK], float (&dy)[NK]) {
        int lo = 0, hh = a.nframes - 1;                    // largest f with s_prefix[f] <= g
        while (lo < hh) { int mid = (lo + hh + 1) >> 1; if (s_prefix[mid] <= g) lo = mid; else hh = mid - 1; }
        K.frame = lo; K.local = g - s_prefix[lo];
        if (a.order) K.local = a.order[(long long)K.frame * a.max_pts + K.local];
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
                    float l = (float)(x - P), kk = (float)(y - P);
                    xp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(co, kk, -__fmul_rn(si, l)), xf), 0.5f);
                    yp = (int)__fadd_rn(__fmaf_rn(fscale, __fmaf_rn(si, kk, __fmul_rn(co, l)), yf), 0.5f);
                } else {                                   // integer pipeline: the reference's own float expressions truncated to int
                    const int li = x - P, ki = y - P;
                    xp = (int)(xf + iscale * (ki * co - li * si) + 0.5f);
                    yp = (int)(yf + iscale * (ki * si + li * co) + 0.5f);
                }
                xp = min(max(xp, 0), L.w - 1); yp = min(max(yp, 0), L.h - 1);       // no-op for in-range patches (border test)
                long long pos = (long long)yp * p + xp;
                im[k] = ldg_wide(imd + pos); dx[k] = ldg_wide(dxd + pos); dy[k] = ldg_wide(dyd + pos);
            }
        }
    };
    Kp cur, nxt;
    float im[NK], dx[NK], dy[NK], nim[NK], ndx[NK], ndy[NK];
    if ((int)blockIdx.x < total) gather(blockIdx.x, cur, im, dx, dy);
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        const bool more = g + (int)gridDim.x < total;
        if (more) gather(g + gridDim.x, nxt, nim, ndx, ndy);
        const float co = cur.co, si = cur.si;
        // 1. rotated derivatives, sample-major into shared memory (samples 441..447 of a channel are never read as members of a cell)
#pragma unroll
        for (int k = 0; k < NK; k++) {
            int i = tid + DS_NT * k;
            if (i < DB_NSP) {
                float rx, ry;
                if (!INT) {
                    rx = __fmaf_rn(co, dy[k], -__fmul_rn(si, dx[k]));
                    ry = __fmaf_rn(co, dx[k], __fmul_rn(si, dy[k]));
                } else {
                    const int dxi = __float_as_int(dx[k]), dyi = __float_as_int(dy[k]);
                    const int rxi = -dxi * si + dyi * co, ryi = dxi * co + dyi * si;          // float expressions truncated to int
                    rx = __int_as_float(rxi); ry = __int_as_float(ryi);
                }
                ds_val[i] = im[k]; ds_val[DB_NSP + i] = rx; ds_val[2 * DB_NSP + i] = ry;
            }
        }
        __syncthreads();
        // 2. running sums of reference thread col over its seven samples, in order; a sum is stored when its run ends
        {
            float aA[3] = { 0.f, 0.f, 0.f }, aB[3] = { 0.f, 0.f, 0.f };
            bool eA = false, eB = false;
#pragma unroll
            for (int m = 0; m < 7; m++) {
                const int i = col + 64 * m;
                const float s3[3] = { ds_val[i], ds_val[DB_NSP + i], ds_val[2 * DB_NSP + i] };
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    aA[c] = dsum<INT>(eA ? 0.f : aA[c], s3[c]);
                    aB[c] = dsum<INT>(eB ? 0.f : aB[c], s3[c]);
                }
                eA = (lastA >> m) & 1u; eB = (lastB >> m) & 1u;
                if (eA) { float* q = s_acc + (offs[m] & 0xFFFFu); q[0] = aA[0]; q[DW_RS] = aA[1]; q[2 * DW_RS] = aA[2]; }
                if (eB && !hi) { float* q = s_acc + (offs[m] >> 16); q[0] = aB[0]; q[DW_RS] = aB[1]; q[2 * DW_RS] = aB[2]; }
            }
        }
        __syncthreads();
        // 3. the reference's reduction tree for value tid, in registers; the row is handed back zeroed
        if (tid < 87) {
            float4* row = reinterpret_cast<float4*>(s_acc + tid * DW_RS);
            const int sw = tid & 7;
            float r[32];
#pragma unroll
            for (int k = 0; k < 8; k++) {                                    // a_t = acc_t + acc_(t+32): float4 k and k + 8 of the row
                const float4 q = row[k ^ sw], u = row[(k ^ sw) + 8];
                r[4 * k] = dsum<INT>(q.x, u.x); r[4 * k + 1] = dsum<INT>(q.y, u.y); r[4 * k + 2] = dsum<INT>(q.z, u.z); r[4 * k + 3] = dsum<INT>(q.w, u.w);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) row[k ^ sw] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
#pragma unroll
                for (int t = 0; t + d < 32; t += 2 * d) r[t] = dsum<INT>(r[t], r[t + d]);
            s_cell[tid] = r[0];
        }
        __syncthreads();
        // 4. the 486 comparisons
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

akz_once_t g_dw_attr;                     // per device: the opt-in to more than 48 KB of dynamic shared memory

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
std::atomic<int> g_describe_generic{getenv("AKZ_DESCRIBE_S") != nullptr ? 1 : 0};      // akz_set_describe_kernel: the generic kernel also for pattern 10

}  // namespace

namespace akzk {

void set_describe_kernel(int which) { g_describe_generic.store(which ? 1 : 0, std::memory_order_relaxed); }

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

int layer_order(cudaStream_t st, const int* counts, const akz_keypoint* kpts, int* order, int max_pts, int n)
{
    k_layer_order<<<n, 1024, 0, st>>>(counts, kpts, order, max_pts);
    return 1;
}

int orient(cudaStream_t st, const AkzLevelTable& tab, const int* counts, const int* prefix, akz_keypoint* kpts, int max_pts, int n, int fast, const int* order)
{
    (void)counts;
    static const int ctas = getenv("AKZ_ORI_CTAS") ? atoi(getenv("AKZ_ORI_CTAS")) : 8;      // tuning knob
    if (fast) k_orient<true><<<148 * ctas, ORI_WARPS * 32, 0, st>>>(tab, prefix, n, kpts, max_pts, order);
    else k_orient<false><<<148 * ctas, ORI_WARPS * 32, 0, st>>>(tab, prefix, n, kpts, max_pts, order);
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
             unsigned char* desc, int max_pts, int n, int pattern, int fast, const int* order)
{
    (void)counts;
    const int s2 = pattern;                                       // akazed.cu:2681-2683
    const int s3 = (int)ceilf(2.0f * pattern / 3.0f);
    const int s4 = (int)ceilf(0.5f * pattern);
    const int win = std::max(3 * s3, 4 * s4), ns = win * win;
    if (ns > DS_MAXNS) return akz_set_error(AKZ_E_UNSUPPORTED, "descriptor_pattern_size %d is not supported (at most %d samples)", pattern, DS_MAXNS);
    if (n > AKZ_MAX_FRAMES_SEARCH) return akz_set_error(AKZ_E_UNSUPPORTED, "more than %d frames per chunk", AKZ_MAX_FRAMES_SEARCH);
    DescArgs a;
    a.prefix = prefix; a.kpts = kpts; a.desc = desc; a.nframes = n; a.max_pts = max_pts; a.order = order;
    a.s2 = s2; a.s3 = s3; a.s4 = s4; a.win = win; a.ns = ns;
    a.tables = desc_tables_for(pattern, s2, s3, s4, win);
    if (!a.tables) return akz_set_error(AKZ_E_NOMEM, "descriptor tables");
    if (pattern == 10 && !g_describe_generic.load(std::memory_order_relaxed)) {                          // the reference default: k_describe_b
        if (akz_once_guard once{g_dw_attr}) {
            cudaFuncSetAttribute(k_describe_b<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM);
            cudaFuncSetAttribute(k_describe_b<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM);
        }
        static const int ctas = getenv("AKZ_DB_CTAS") ? atoi(getenv("AKZ_DB_CTAS")) : DB_CTAS;      // tuning knob
        if (fast) k_describe_b<true><<<148 * ctas, DS_NT, DB_SMEM, st>>>(tab, a);
        else k_describe_b<false><<<148 * ctas, DS_NT, DB_SMEM, st>>>(tab, a);
        return 1;
    }
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
