// k_prep2<S, MODE>: the per-level "everything but the diffusion" kernel, second generation.
//
// Same contract as k_level_prep (scale_space_fused.cu): one read of the predecessor tile produces the
// conductance plane g, the first derivatives Lx, Ly and the Hessian determinant of a level; the sigma=1
// blur, its row-filtered intermediate and the derivative tiles live in shared memory only.
// Subsumes gConv2d<2> | gDownWithSmooth (akazed.cu:204, :449), gFlowNaive (:1068), gDerivate (:1267) and
// gHessianDeterminant (:1299).
//
// What changed against the first generation (ncu, profiles/r01_ncu_summary.md: 533 thread-instructions per
// pixel, ALU pipe 54 % busy with index arithmetic, FMA pipe 30 %):
//   * the derivative step S and the mode are template parameters and all tiles share ONE shared-memory
//     coordinate frame (column = gx - X0 + 12, row = gy - Y0 + 2S + 2, pitch 88), so every neighbour offset
//     is an immediate and every phase is a flat loop over "4 consecutive pixels" items: LDS.128 / STS.128 /
//     STG.128 only, no per-element index arithmetic;
//   * reflect-101 handling left the hot loops: interior tiles (82 % at 1080p) stream their input with
//     cp.async 16-byte copies; border tiles load with reflected indices and, after the blur and after the
//     first-derivative phase, overwrite the out-of-image cells of the tile with the value of their mirror
//     cell (a copy, so operators are never evaluated on a mirrored extension: Lx/Ly are antisymmetric and
//     the octave transition reflects in SOURCE coordinates);
//   * 64x48 output tile, 384 threads, 3 CTAs per SM (first version: 64x64, 512 threads, 2 CTAs: the six barrier-separated
//     phases overlap better across three CTAs);
//   * the column pass works on four-row items and the conductance phase on row pairs (fewer shared-memory loads per output:
//     the kernel sits on the LSU wavefront pipe, 85 % busy);
//   * INT = true instantiates the same kernel for the integer pipeline (int32 planes as bit patterns, 16.16 fixed point).
// The arithmetic is the pinned sequence of common.cuh: results are bit-identical to the staged kernels.
#include "common.cuh"
#include "kernels.h"
#include "level_math.cuh"

using namespace akz;

namespace {

constexpr int P2_W = 64, P2_H = 48;          // output tile
constexpr int P2_OX = 12;                    // shared-memory column of output column 0 (multiple of 4, >= 2*4 + 2)
constexpr int P2_SP = P2_W + 2 * P2_OX;      // 88 floats per shared-memory row
constexpr int P2_NT = 384;
enum { PM_BASE = 0, PM_BLUR = 1, PM_DOWN = 2 };

struct Prep2Args {
    const float* src;                    // predecessor Lt (same resolution) or source octave (PM_DOWN)
    float* ltdst;                        // PM_DOWN: subsampled Lt
    float *flow, *lx, *ly, *det;         // flow may be null
    const float* kc;                     // per-frame contrast factor
    long long splane, plane;
    LevelMathArgs m;                     // taps and derivative factors, float and 16.16
    float kscale;
    int nmul, type, vec_ok;
    int sw, sh, sp;                      // source dims (== w, h, pitch unless PM_DOWN)
    int w, h, pitch;
};

// 128-bit shared-memory load that the compiler may not narrow: when only some components are used (S = 2, 3) nvcc
// splits a float4 load into LDS.32 / LDS.64 pieces, which at a 16-byte lane stride are 4-way / 2-way bank conflicts
// (ncu r01b source page: 36 % of the shared-memory wavefronts of this kernel were conflict replays).
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
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// source-coordinate reflect of a coarse index (gDownWithSmooth reflects 2x+-2, 2x+-4 in the SOURCE image)
__device__ __forceinline__ int coarse_src2(int q, int sdim)
{
    int c = 2 * q;
    if (c < 0) c = -c;
    if (c >= sdim) c = sdim + sdim - 2 - c;
    return min(max(c, 0), sdim - 1);
}

// out-of-image cells of a tile := value of their reflect-101 mirror cell (border tiles only).
// Only the ghost band is visited: the rows above the image / below it over the full width, then the columns left /
// right of it over the remaining rows (visiting every cell of the region was 9 % of the kernel's instructions, ncu r01f).
template <int OY, int C0, int C1>
__device__ __forceinline__ void ghost_fix(float* T, int r0, int r1, int X0, int Y0, int w, int h, int tid)
{
    constexpr int nc = C1 - C0;
    const int nr = r1 - r0;
    const int gy0 = Y0 - OY + r0, gx0 = X0 - P2_OX + C0;           // image coordinates of the region's first cell
    const int top = min(max(-gy0, 0), nr), bot = min(max(h - gy0, top), nr);      // in-image rows of the region: [top, bot)
    const int lft = min(max(-gx0, 0), nc), rgt = min(max(w - gx0, lft), nc);      // in-image columns:           [lft, rgt)
    auto fix = [&](int rr, int cc) {
        int r = r0 + rr, c = C0 + cc;
        int mr = min(max(refl(gy0 + rr, h) - Y0 + OY, r0), r1 - 1);
        int mc = min(max(refl(gx0 + cc, w) - X0 + P2_OX, C0), C1 - 1);
        T[r * P2_SP + c] = T[mr * P2_SP + mc];
    };
    const int nghost_rows = top + (nr - bot);
    for (int i = tid; i < nghost_rows * nc; i += P2_NT) {
        int rr = i / nc, cc = i - rr * nc;
        if (rr >= top) rr += bot - top;
        fix(rr, cc);
    }
    const int nghost_cols = lft + (nc - rgt);
    if (nghost_cols > 0) {
        for (int i = tid; i < (bot - top) * nghost_cols; i += P2_NT) {
            int rr = i / nghost_cols, cc = i - rr * nghost_cols;
            if (cc >= lft) cc += rgt - lft;
            fix(top + rr, cc);
        }
    }
}

template <int S, int MODE, bool INT>
__global__ void __launch_bounds__(P2_NT, 3) k_prep2(const __grid_constant__ Prep2Args a)
{
    constexpr int OY = 2 * S + 2;                     // shared-memory row of output row 0
    constexpr int AR = P2_H + 2 * OY;                 // rows of the input tile
    constexpr int SP = P2_SP;
    extern __shared__ __align__(16) float sm[];
    float* A = sm;                                    // input tile            rows [0, AR)
    float* Bf = A + AR * SP;                          // row-filtered          rows [0, AR)
    float* Sm = Bf + AR * SP;                         // sigma = 1 blur        rows [2, AR-2)
    float* LX = A;                                    // first derivatives     rows [OY-S, OY+64+S)   (alias A / Bf)
    float* LY = Bf;

    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int X0 = blockIdx.x * P2_W, Y0 = blockIdx.y * P2_H;
    const int w = a.w, h = a.h;
    const bool interior = X0 - P2_OX >= 0 && X0 + P2_W + P2_OX <= w && Y0 - OY >= 0 && Y0 + P2_H + OY <= h;
    const bool fast = interior && a.vec_ok;
    const float* __restrict__ src = a.src + (long long)frame * a.splane;

    // ---- 1. input tile ------------------------------------------------------------------------------
    {
        float* dstp = (MODE == PM_BASE) ? Sm : A;
        if (MODE != PM_DOWN && fast) {
            for (int i = tid; i < AR * (SP / 4); i += P2_NT) {
                int r = i / (SP / 4), g = i - r * (SP / 4);
                cp_async16(dstp + r * SP + 4 * g, src + (long long)(Y0 - OY + r) * a.sp + (X0 - P2_OX + 4 * g));
            }
            cp_async_wait_all();
        } else if (MODE == PM_DOWN) {
            for (int i = tid; i < AR * SP; i += P2_NT) {
                int r = i / SP, c = i - r * SP;
                int sy = coarse_src2(Y0 - OY + r, a.sh), sx = coarse_src2(X0 - P2_OX + c, a.sw);
                dstp[i] = __ldg(src + (long long)sy * a.sp + sx);
            }
        } else {
            // border tile: rows by reflected index; a group of 4 columns that lies inside the image is one float4 load
            for (int i = tid; i < AR * (SP / 4); i += P2_NT) {
                int r = i / (SP / 4), g = i - r * (SP / 4);
                int sy = min(max(refl(Y0 - OY + r, h), 0), h - 1), gx = X0 - P2_OX + 4 * g;
                const float* row = src + (long long)sy * a.sp;
                float* d = dstp + r * SP + 4 * g;
                if (a.vec_ok && gx >= 0 && gx + 3 < w) *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(row + gx));
                else {
#pragma unroll
                    for (int j = 0; j < 4; j++) d[j] = __ldg(row + min(max(refl(gx + j, w), 0), w - 1));
                }
            }
        }
    }
    __syncthreads();

    if (MODE == PM_DOWN) {
        // subsampled plane: dst(x, y) = src(2x, 2y)   (akazed.cu:505)
        float* ltd = a.ltdst + (long long)frame * a.plane;
        for (int i = tid; i < P2_H * P2_W; i += P2_NT) {
            int r = i >> 6, c = i & 63;
            int y = Y0 + r, x = X0 + c;
            if (y < h && x < w) ltd[(long long)y * a.pitch + x] = A[(r + OY) * SP + P2_OX + c];
        }
    }

    if (MODE != PM_BASE) {
        // ---- 2. row pass: Bf[r][4..84) from A[r][2..86) -----------------------------------------------
        // (24 lanes per row, 20 active: a quarter-warp never straddles two rows, which would be a 2-way bank conflict)
        for (int i = tid; i < AR * 24; i += P2_NT) {
            int r = i / 24, g = i - r * 24;
            if (g >= 20) continue;
            const float* p = A + r * SP + 4 + 4 * g;
            float4 v0 = lds4(p - 4), v1 = lds4(p), v2 = lds4(p + 4);
            sts4(Bf + r * SP + 4 + 4 * g,
                 p2_gauss<INT>(v0.z, v0.w, v1.x, v1.y, v1.z, a.m), p2_gauss<INT>(v0.w, v1.x, v1.y, v1.z, v1.w, a.m),
                 p2_gauss<INT>(v1.x, v1.y, v1.z, v1.w, v2.x, a.m), p2_gauss<INT>(v1.y, v1.z, v1.w, v2.x, v2.y, a.m));
        }
        __syncthreads();
        // ---- 3. column pass: Sm rows [2, AR-2), four rows per item (8 loads + 4 stores per 4 rows; two-row items: 2 x (6 + 2)) ----
        static_assert((AR - 4) % 4 == 0, "column pass works in groups of four rows");
        for (int i = tid; i < ((AR - 4) / 4) * 24; i += P2_NT) {
            int rb = i / 24, g = i - rb * 24;
            if (g >= 20) continue;
            int r = 2 + 4 * rb;
            const float* p = Bf + (r - 2) * SP + 4 + 4 * g;
            float4 b[8];
#pragma unroll
            for (int k = 0; k < 8; k++) b[k] = lds4(p + k * SP);
            float* q = Sm + r * SP + 4 + 4 * g;
#pragma unroll
            for (int t = 0; t < 4; t++)
                sts4(q + t * SP, p2_gauss<INT>(b[t].x, b[t + 1].x, b[t + 2].x, b[t + 3].x, b[t + 4].x, a.m), p2_gauss<INT>(b[t].y, b[t + 1].y, b[t + 2].y, b[t + 3].y, b[t + 4].y, a.m),
                     p2_gauss<INT>(b[t].z, b[t + 1].z, b[t + 2].z, b[t + 3].z, b[t + 4].z, a.m), p2_gauss<INT>(b[t].w, b[t + 1].w, b[t + 2].w, b[t + 3].w, b[t + 4].w, a.m));
        }
        __syncthreads();
    }
    if (!interior) {
        ghost_fix<OY, 4, 84>(Sm, 2, AR - 2, X0, Y0, w, h, tid);
        __syncthreads();
    }

    // ---- 4. conductance of the output pixels ----------------------------------------------------------
    if (a.flow) {
        float ikc;
        if (!INT) {
            float k = a.kc[frame];
            for (int i = 0; i < a.nmul; i++) k = __fmul_rn(k, a.kscale);
            ikc = __fdiv_rn(1.f, __fmul_rn(k, k));
        } else {
            // integer contrast factor, scaled per octave as akaze.cpp:649 does on the host: k = (int)(k * 0.75f + 0.5f)
            int k = reinterpret_cast<const int*>(a.kc)[frame];
            for (int i = 0; i < a.nmul; i++) k = (int)__fadd_rn(__fmul_rn((float)k, 0.75f), 0.5f);
            ikc = __fdiv_rn(1.f, (float)(k * k));                       // akazed.cu:4218 (host)
        }
        float* fl = a.flow + (long long)frame * a.plane;
        // two output rows per item: four row loads for two rows instead of six
        static_assert(P2_H % 2 == 0, "conductance phase works on row pairs");
        for (int i = tid; i < (P2_H / 2) * 16; i += P2_NT) {
            int r = 2 * (i >> 4), g = i & 15;
            const float* p = Sm + (r + OY) * SP + P2_OX + 4 * g;
            // the two values beside the float4 come from the neighbouring lanes (same row pair: 16 items per row, all 32
            // lanes active); only the row ends read shared memory (a 4-float-strided scalar LDS is a 4-way bank conflict)
            float rw[4][6];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* q = p + (k - 1) * SP;
                float4 v = lds4(q);
                float lft = __shfl_up_sync(0xffffffffu, v.w, 1), rgt = __shfl_down_sync(0xffffffffu, v.x, 1);
                if (g == 0) lft = q[-1];
                if (g == 15) rgt = q[4];
                rw[k][0] = lft; rw[k][1] = v.x; rw[k][2] = v.y; rw[k][3] = v.z; rw[k][4] = v.w; rw[k][5] = rgt;
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    o[j] = p2_flow<INT>(rw[t][j], rw[t][j + 1], rw[t][j + 2], rw[t + 1][j], rw[t + 1][j + 2], rw[t + 2][j], rw[t + 2][j + 1], rw[t + 2][j + 2], a.type, ikc);
                int y = Y0 + r + t, x = X0 + 4 * g;
                float* d = fl + (long long)y * a.pitch + x;
                if (fast) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
                else if (y < h) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (x + j < w) d[j] = o[j];
                }
            }
        }
    }

    // ---- 5. first derivatives on the tile extended by S: rows [OY-S, OY+64+S), columns [8, 80) ----------
    {
        float* lxg = a.lx + (long long)frame * a.plane;
        float* lyg = a.ly + (long long)frame * a.plane;
        for (int i = tid; i < (P2_H + 2 * S) * 24; i += P2_NT) {
            int r = i / 24, g = i - r * 24;
            if (g >= 18) continue;
            int sr = OY - S + r, c4 = 8 + 4 * g;
            const float* p = Sm + sr * SP + c4;
            float u[12], c[12], l[12];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                float4 vu = lds4(p - S * SP - 4 + 4 * q), vl = lds4(p + S * SP - 4 + 4 * q);
                u[4 * q] = vu.x; u[4 * q + 1] = vu.y; u[4 * q + 2] = vu.z; u[4 * q + 3] = vu.w;
                if (S < 4 || q != 1) {                       // the centre row is read at m - S and m + S only
                    float4 vc = lds4(p - 4 + 4 * q);
                    c[4 * q] = vc.x; c[4 * q + 1] = vc.y; c[4 * q + 2] = vc.z; c[4 * q + 3] = vc.w;
                }
                l[4 * q] = vl.x; l[4 * q + 1] = vl.y; l[4 * q + 2] = vl.z; l[4 * q + 3] = vl.w;
            }
            float vx[4], vy[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int m = 4 + j;
                float ul = u[m - S], uc = u[m], ur = u[m + S], cl = c[m - S], cr = c[m + S], ll = l[m - S], lc = l[m], lr = l[m + S];
                p2_deriv1<INT>(ul, uc, ur, cl, cr, ll, lc, lr, a.m, vx[j], vy[j]);
            }
            sts4(LX + sr * SP + c4, vx[0], vx[1], vx[2], vx[3]);
            sts4(LY + sr * SP + c4, vy[0], vy[1], vy[2], vy[3]);
            if (r >= S && r < S + P2_H && g >= 1 && g <= 16) {
                int y = Y0 + r - S, x = X0 + 4 * (g - 1);
                long long o = (long long)y * a.pitch + x;
                if (fast) {
                    *reinterpret_cast<float4*>(lxg + o) = make_float4(vx[0], vx[1], vx[2], vx[3]);
                    *reinterpret_cast<float4*>(lyg + o) = make_float4(vy[0], vy[1], vy[2], vy[3]);
                } else if (y < h) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (x + j < w) { lxg[o + j] = vx[j]; lyg[o + j] = vy[j]; }
                }
            }
        }
    }
    __syncthreads();
    if (!interior) {
        ghost_fix<OY, 8, 80>(LX, OY - S, OY + P2_H + S, X0, Y0, w, h, tid);
        ghost_fix<OY, 8, 80>(LY, OY - S, OY + P2_H + S, X0, Y0, w, h, tid);
        __syncthreads();
    }

    // ---- 6. second derivatives and determinant ----------------------------------------------------------
    {
        float* dg = a.det + (long long)frame * a.plane;
        for (int i = tid; i < P2_H * 16; i += P2_NT) {
            int r = i >> 4, g = i & 15;
            const float* px = LX + (r + OY) * SP + P2_OX + 4 * g;
            const float* py = LY + (r + OY) * SP + P2_OX + 4 * g;
            float xu[12], xc[12], xl[12], yu[12], yl[12];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                float4 t;
                t = lds4(px - S * SP - 4 + 4 * q); xu[4 * q] = t.x; xu[4 * q + 1] = t.y; xu[4 * q + 2] = t.z; xu[4 * q + 3] = t.w;
                t = lds4(px + S * SP - 4 + 4 * q); xl[4 * q] = t.x; xl[4 * q + 1] = t.y; xl[4 * q + 2] = t.z; xl[4 * q + 3] = t.w;
                t = lds4(py - S * SP - 4 + 4 * q); yu[4 * q] = t.x; yu[4 * q + 1] = t.y; yu[4 * q + 2] = t.z; yu[4 * q + 3] = t.w;
                t = lds4(py + S * SP - 4 + 4 * q); yl[4 * q] = t.x; yl[4 * q + 1] = t.y; yl[4 * q + 2] = t.z; yl[4 * q + 3] = t.w;
                if (S < 4 || q != 1) { t = lds4(px - 4 + 4 * q); xc[4 * q] = t.x; xc[4 * q + 1] = t.y; xc[4 * q + 2] = t.z; xc[4 * q + 3] = t.w; }
            }
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int m = 4 + j;
                o[j] = p2_det<INT>(xu[m - S], xu[m], xu[m + S], xc[m - S], xc[m + S], xl[m - S], xl[m], xl[m + S],
                                   yu[m - S], yu[m], yu[m + S], yl[m - S], yl[m], yl[m + S], a.m);
            }
            int y = Y0 + r, x = X0 + 4 * g;
            float* d = dg + (long long)y * a.pitch + x;
            if (fast) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            else if (y < h) {
#pragma unroll
                for (int j = 0; j < 4; j++) if (x + j < w) d[j] = o[j];
            }
        }
    }
}

// =====================================================================================================================
// k_prep3<MODE>: the blur half of a level for the split pipeline -- sigma = 1 blur (or octave transition) of the predecessor
// tile, the conductance plane g, and the blurred plane itself written to global memory for the streaming derivative kernel
// (deriv_stream.cu).  Same phases 1-4, arithmetic and border treatment as k_prep2, on a frame that only carries the halo
// the blur (2) and the Scharr conductance (1) need: 64 x 48 output tile, rows [-4, 52), columns [-8, 72), 80 floats per
// shared-memory row; the blurred tile overwrites the input tile.  Subsumes gConv2d<2> | gDownWithSmooth (akazed.cu:204, :449)
// and gFlowNaive (:1068).
// =====================================================================================================================
constexpr int P3_OX = 8, P3_OY = 4;
constexpr int P3_SP = P2_W + 2 * P3_OX;       // 80
constexpr int P3_AR = P2_H + 2 * P3_OY;       // 56
constexpr int P3_C0 = 4, P3_C1 = 76;          // columns of the frame that are computed (18 groups of 4)
constexpr int P3_NG = (P3_C1 - P3_C0) / 4;

// out-of-image cells of the blurred tile := value of their reflect-101 mirror cell (border tiles only; see ghost_fix)
__device__ __forceinline__ void ghost_fix3(float* T, int r0, int r1, int X0, int Y0, int w, int h, int tid)
{
    constexpr int GC0 = P3_OX - 1, GC1 = P3_OX + P2_W + 1;          // the conductance stencil reads one column beyond the tile
    constexpr int nc = GC1 - GC0;
    const int nr = r1 - r0;
    const int gy0 = Y0 - P3_OY + r0, gx0 = X0 - P3_OX + GC0;
    const int top = min(max(-gy0, 0), nr), bot = min(max(h - gy0, top), nr);
    const int lft = min(max(-gx0, 0), nc), rgt = min(max(w - gx0, lft), nc);
    auto fix = [&](int rr, int cc) {
        int r = r0 + rr, c = GC0 + cc;
        int mr = min(max(refl(gy0 + rr, h) - Y0 + P3_OY, r0), r1 - 1);
        int mc = min(max(refl(gx0 + cc, w) - X0 + P3_OX, GC0), GC1 - 1);
        T[r * P3_SP + c] = T[mr * P3_SP + mc];
    };
    const int nghost_rows = top + (nr - bot);
    for (int i = tid; i < nghost_rows * nc; i += P2_NT) {
        int rr = i / nc, cc = i - rr * nc;
        if (rr >= top) rr += bot - top;
        fix(rr, cc);
    }
    const int nghost_cols = lft + (nc - rgt);
    if (nghost_cols > 0) {
        for (int i = tid; i < (bot - top) * nghost_cols; i += P2_NT) {
            int rr = i / nghost_cols, cc = i - rr * nghost_cols;
            if (cc >= lft) cc += rgt - lft;
            fix(top + rr, cc);
        }
    }
}

struct Prep3Args {
    const float* src;                    // predecessor Lt (same resolution) or source octave (PM_DOWN)
    float* ltdst;                        // PM_DOWN: subsampled Lt
    float *flow, *smooth;                // conductance and blurred plane of the level
    const float* kc;                     // per-frame contrast factor
    long long splane, plane;
    LevelMathArgs m;
    float kscale;
    int nmul, type, vec_ok;
    int sw, sh, sp;                      // source dims (== w, h, pitch unless PM_DOWN)
    int w, h, pitch;
};

template <int MODE, bool INT>
__global__ void __launch_bounds__(P2_NT, 3) k_prep3(const __grid_constant__ Prep3Args a)
{
    constexpr int OY = P3_OY, AR = P3_AR, SP = P3_SP, OX = P3_OX;
    extern __shared__ __align__(16) float sm[];
    float* A = sm;                                    // input tile, later the blurred tile
    float* Bf = A + AR * SP;                          // row-filtered
    float* Sm = A;

    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int X0 = blockIdx.x * P2_W, Y0 = blockIdx.y * P2_H;
    const int w = a.w, h = a.h;
    const bool interior = X0 - OX >= 0 && X0 + P2_W + OX <= w && Y0 - OY >= 0 && Y0 + P2_H + OY <= h;
    const bool fast = interior && a.vec_ok;
    const float* __restrict__ src = a.src + (long long)frame * a.splane;

    // ---- 1. input tile: rows [0, AR), columns [C0, C1) ----------------------------------------------
    if (MODE != PM_DOWN && fast) {
        for (int i = tid; i < AR * P3_NG; i += P2_NT) {
            int r = i / P3_NG, g = i - r * P3_NG;
            cp_async16(A + r * SP + P3_C0 + 4 * g, src + (long long)(Y0 - OY + r) * a.sp + (X0 - OX + P3_C0 + 4 * g));
        }
        cp_async_wait_all();
    } else if (MODE == PM_DOWN) {
        for (int i = tid; i < AR * (P3_C1 - P3_C0); i += P2_NT) {
            int r = i / (P3_C1 - P3_C0), c = P3_C0 + i - r * (P3_C1 - P3_C0);
            int sy = coarse_src2(Y0 - OY + r, a.sh), sx = coarse_src2(X0 - OX + c, a.sw);
            A[r * SP + c] = __ldg(src + (long long)sy * a.sp + sx);
        }
    } else {
        for (int i = tid; i < AR * P3_NG; i += P2_NT) {
            int r = i / P3_NG, g = i - r * P3_NG;
            int sy = min(max(refl(Y0 - OY + r, h), 0), h - 1), gx = X0 - OX + P3_C0 + 4 * g;
            const float* row = src + (long long)sy * a.sp;
            float* d = A + r * SP + P3_C0 + 4 * g;
            if (a.vec_ok && gx >= 0 && gx + 3 < w) *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(row + gx));
            else {
#pragma unroll
                for (int j = 0; j < 4; j++) d[j] = __ldg(row + min(max(refl(gx + j, w), 0), w - 1));
            }
        }
    }
    __syncthreads();

    if (MODE == PM_DOWN) {
        // subsampled plane: dst(x, y) = src(2x, 2y)   (akazed.cu:505)
        float* ltd = a.ltdst + (long long)frame * a.plane;
        for (int i = tid; i < P2_H * P2_W; i += P2_NT) {
            int r = i >> 6, c = i & 63;
            int y = Y0 + r, x = X0 + c;
            if (y < h && x < w) ltd[(long long)y * a.pitch + x] = A[(r + OY) * SP + OX + c];
        }
    }

    // ---- 2. row pass: Bf[r][8, 72) from A[r][4, 76) (16 groups; the two outer groups only feed it) ----------------
    for (int i = tid; i < AR * 16; i += P2_NT) {
        int r = i >> 4, g = i & 15;
        const float* p = A + r * SP + OX + 4 * g;
        float4 v0 = lds4(p - 4), v1 = lds4(p), v2 = lds4(p + 4);
        sts4(Bf + r * SP + OX + 4 * g,
             p2_gauss<INT>(v0.z, v0.w, v1.x, v1.y, v1.z, a.m), p2_gauss<INT>(v0.w, v1.x, v1.y, v1.z, v1.w, a.m),
             p2_gauss<INT>(v1.x, v1.y, v1.z, v1.w, v2.x, a.m), p2_gauss<INT>(v1.y, v1.z, v1.w, v2.x, v2.y, a.m));
    }
    // the blur is also needed one column beyond the tile on either side (conductance stencil): columns 7 and 72
    for (int i = tid; i < AR * 2; i += P2_NT) {
        int r = i >> 1, c = (i & 1) ? OX + P2_W : OX - 1;
        const float* p = A + r * SP + c;
        Bf[r * SP + c] = p2_gauss<INT>(p[-2], p[-1], p[0], p[1], p[2], a.m);
    }
    __syncthreads();
    // ---- 3. column pass: Sm rows [2, AR-2) (over the input tile), four rows per item; 16 groups + the two edge columns ----
    static_assert((AR - 4) % 4 == 0, "column pass works in groups of four rows");
    for (int i = tid; i < ((AR - 4) / 4) * 16; i += P2_NT) {
        int rb = i >> 4, g = i & 15;
        int r = 2 + 4 * rb;
        const float* p = Bf + (r - 2) * SP + OX + 4 * g;
        float4 b[8];
#pragma unroll
        for (int k = 0; k < 8; k++) b[k] = lds4(p + k * SP);
        float* q = Sm + r * SP + OX + 4 * g;
#pragma unroll
        for (int t = 0; t < 4; t++)
            sts4(q + t * SP, p2_gauss<INT>(b[t].x, b[t + 1].x, b[t + 2].x, b[t + 3].x, b[t + 4].x, a.m), p2_gauss<INT>(b[t].y, b[t + 1].y, b[t + 2].y, b[t + 3].y, b[t + 4].y, a.m),
                 p2_gauss<INT>(b[t].z, b[t + 1].z, b[t + 2].z, b[t + 3].z, b[t + 4].z, a.m), p2_gauss<INT>(b[t].w, b[t + 1].w, b[t + 2].w, b[t + 3].w, b[t + 4].w, a.m));
    }
    for (int i = tid; i < (AR - 4) * 2; i += P2_NT) {
        int r = 2 + (i >> 1), c = (i & 1) ? OX + P2_W : OX - 1;
        const float* p = Bf + r * SP + c;
        Sm[r * SP + c] = p2_gauss<INT>(p[-2 * SP], p[-SP], p[0], p[SP], p[2 * SP], a.m);
    }
    __syncthreads();
    if (!interior) {
        ghost_fix3(Sm, OY - 1, OY + P2_H + 1, X0, Y0, w, h, tid);
        __syncthreads();
    }

    // ---- 4. blurred plane to global memory + conductance of the output pixels -------------------------------------
    float ikc;
    if (!INT) {
        float k = a.kc[frame];
        for (int i = 0; i < a.nmul; i++) k = __fmul_rn(k, a.kscale);
        ikc = __fdiv_rn(1.f, __fmul_rn(k, k));
    } else {
        int k = reinterpret_cast<const int*>(a.kc)[frame];
        for (int i = 0; i < a.nmul; i++) k = (int)__fadd_rn(__fmul_rn((float)k, 0.75f), 0.5f);       // akaze.cpp:649
        ikc = __fdiv_rn(1.f, (float)(k * k));                                                        // akazed.cu:4218 (host)
    }
    float* fl = a.flow + (long long)frame * a.plane;
    float* smg = a.smooth + (long long)frame * a.plane;
    static_assert(P2_H % 2 == 0, "conductance phase works on row pairs");
    for (int i = tid; i < (P2_H / 2) * 16; i += P2_NT) {
        int r = 2 * (i >> 4), g = i & 15;
        const float* p = Sm + (r + OY) * SP + OX + 4 * g;
        float rw[4][6];
        float4 ctr[2];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float* q = p + (k - 1) * SP;
            float4 v = lds4(q);
            float lft = __shfl_up_sync(0xffffffffu, v.w, 1), rgt = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (g == 0) lft = q[-1];
            if (g == 15) rgt = q[4];
            rw[k][0] = lft; rw[k][1] = v.x; rw[k][2] = v.y; rw[k][3] = v.z; rw[k][4] = v.w; rw[k][5] = rgt;
            if (k == 1) ctr[0] = v;
            if (k == 2) ctr[1] = v;
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                o[j] = p2_flow<INT>(rw[t][j], rw[t][j + 1], rw[t][j + 2], rw[t + 1][j], rw[t + 1][j + 2], rw[t + 2][j], rw[t + 2][j + 1], rw[t + 2][j + 2], a.type, ikc);
            int y = Y0 + r + t, x = X0 + 4 * g;
            long long off = (long long)y * a.pitch + x;
            if (fast) {
                *reinterpret_cast<float4*>(fl + off) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(smg + off) = ctr[t];
            } else if (y < h) {
                const float cv[4] = { ctr[t].x, ctr[t].y, ctr[t].z, ctr[t].w };
#pragma unroll
                for (int j = 0; j < 4; j++) if (x + j < w) { fl[off + j] = o[j]; smg[off + j] = cv[j]; }
            }
        }
    }
}

constexpr int prep3_smem() { return 2 * P3_AR * P3_SP * (int)sizeof(float); }

template <int MODE, bool INT>
void prep3_launch(cudaStream_t st, const Prep3Args& a, int n)
{
    dim3 g((a.w + P2_W - 1) / P2_W, (a.h + P2_H - 1) / P2_H, n);
    k_prep3<MODE, INT><<<g, P2_NT, prep3_smem(), st>>>(a);
}

template <int S>
constexpr int prep2_smem() { return 3 * (P2_H + 2 * (2 * S + 2)) * P2_SP * (int)sizeof(float); }

template <int S, int MODE, bool INT>
void prep2_launch(cudaStream_t st, const Prep2Args& a, int n)
{
    static akz_once_t attr;
    if (akz_once_guard once{attr}) cudaFuncSetAttribute(k_prep2<S, MODE, INT>, cudaFuncAttributeMaxDynamicSharedMemorySize, prep2_smem<S>());
    dim3 g((a.w + P2_W - 1) / P2_W, (a.h + P2_H - 1) / P2_H, n);
    k_prep2<S, MODE, INT><<<g, P2_NT, prep2_smem<S>(), st>>>(a);
}

template <int MODE, bool INT>
bool prep2_dispatch(cudaStream_t st, const Prep2Args& a, int step, int n)
{
    switch (step) {
    case 2: prep2_launch<2, MODE, INT>(st, a, n); return true;
    case 3: prep2_launch<3, MODE, INT>(st, a, n); return true;
    case 4: prep2_launch<4, MODE, INT>(st, a, n); return true;
    default: return false;
    }
}

}  // namespace

namespace akzk {

// mode: 0 = base level (smooth := src, no blur), 1 = same-resolution blur, 2 = octave transition.
// Returns 1 when launched, 0 when this (step, size) combination is not covered (caller falls back to k_level_prep).
int level_prep2(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                float* ltdst, float* flowp, float* lx, float* ly, float* det, int type,
                const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n, int int_planes)
{
    if (step < 2 || step > 4 || w < 24 || h < 24) return 0;       // reflections must stay single (halo <= 10 + 2)
    Prep2Args a = {};
    a.src = src; a.ltdst = ltdst; a.flow = flowp; a.lx = lx; a.ly = ly; a.det = det; a.kc = kc;
    a.splane = splane; a.plane = plane; a.kscale = kscale; a.nmul = nmul; a.type = type;
    a.sw = sw; a.sh = sh; a.sp = sp; a.w = w; a.h = h; a.pitch = pitch;
    hessian_factors(&a.m.fac1, &a.m.fac2);
    float k[3];
    akz_gauss_taps(1.f, 2, k);
    a.m.k0 = k[0]; a.m.k1 = k[1]; a.m.k2 = k[2];
    auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p % 16) == 0; };
    a.vec_ok = (pitch % 4 == 0) && (plane % 4 == 0) && (sp % 4 == 0) && (splane % 4 == 0) &&
               al16(src) && al16(flowp) && al16(lx) && al16(ly) && al16(det);
    a.m.ik0 = (int)(k[0] * 65536 + 0.5f); a.m.ik1 = (int)(k[1] * 65536 + 0.5f); a.m.ik2 = (int)(k[2] * 65536 + 0.5f);      // akazed.cu:3896
    a.m.ifac1 = (int)(a.m.fac1 * 65536 + 0.5f); a.m.ifac2 = (int)(a.m.fac2 * 65536 + 0.5f);                                  // akazed.cu:4184-4185
    bool ok;
    if (int_planes)
        ok = mode == 0 ? prep2_dispatch<PM_BASE, true>(st, a, step, n) : mode == 1 ? prep2_dispatch<PM_BLUR, true>(st, a, step, n)
                                                                                 : prep2_dispatch<PM_DOWN, true>(st, a, step, n);
    else
        ok = mode == 0 ? prep2_dispatch<PM_BASE, false>(st, a, step, n) : mode == 1 ? prep2_dispatch<PM_BLUR, false>(st, a, step, n)
                                                                                  : prep2_dispatch<PM_DOWN, false>(st, a, step, n);
    return ok ? 1 : 0;
}

// Blur half of a level (k_prep3): mode 1 = same-resolution blur, 2 = octave transition.  Writes the conductance plane and
// the blurred plane (for deriv_stream).  Returns 1 when launched, 0 when not covered.
int level_blur_flow(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                    float* ltdst, float* flowp, float* smooth, int type, const float* kc, float kscale, int nmul,
                    int w, int h, int pitch, long long plane, int n, int int_planes)
{
    if ((mode != 1 && mode != 2) || w < 24 || h < 24 || !flowp || !smooth) return 0;
    Prep3Args a = {};
    a.src = src; a.ltdst = ltdst; a.flow = flowp; a.smooth = smooth; a.kc = kc;
    a.splane = splane; a.plane = plane; a.kscale = kscale; a.nmul = nmul; a.type = type;
    a.sw = sw; a.sh = sh; a.sp = sp; a.w = w; a.h = h; a.pitch = pitch;
    float k[3];
    akz_gauss_taps(1.f, 2, k);
    a.m.k0 = k[0]; a.m.k1 = k[1]; a.m.k2 = k[2];
    a.m.ik0 = (int)(k[0] * 65536 + 0.5f); a.m.ik1 = (int)(k[1] * 65536 + 0.5f); a.m.ik2 = (int)(k[2] * 65536 + 0.5f);      // akazed.cu:3896
    auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p % 16) == 0; };
    a.vec_ok = (pitch % 4 == 0) && (plane % 4 == 0) && (sp % 4 == 0) && (splane % 4 == 0) && al16(src) && al16(flowp) && al16(smooth);
    if (int_planes) { if (mode == 1) prep3_launch<PM_BLUR, true>(st, a, n); else prep3_launch<PM_DOWN, true>(st, a, n); }
    else { if (mode == 1) prep3_launch<PM_BLUR, false>(st, a, n); else prep3_launch<PM_DOWN, false>(st, a, n); }
    return 1;
}

}  // namespace akzk
