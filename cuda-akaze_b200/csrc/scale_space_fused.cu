// Production kernels of the nonlinear scale space.
//
//   k_level_prep<DOWN>  one pass per level: reads a tile (+halo) of the predecessor plane once and
//                       produces everything of the level that does not depend on the diffusion:
//                       the sigma=1 blur (shared memory only), the conductance plane g, the first
//                       derivatives Lx, Ly and the Hessian determinant.  DOWN=1 additionally performs
//                       the octave transition (point subsample + coarse-lattice blur).
//                       Subsumes gConv2d<2> | gDownWithSmooth, gFlowNaive, gDerivate,
//                       gHessianDeterminant (akazed.cu:204, :449, :1068, :1267, :1299).
//   k_fed               all (up to 8) explicit diffusion steps of a FED cycle in one launch: the tile
//                       and its halo live in shared memory, the valid region shrinks by one ring per
//                       step (temporal blocking).  Subsumes n x gNldStepNaive + n D2D copies
//                       (akazed.cu:1241, akaze.cpp:383-391, :412-420).
//
// Bit-exactness: every value is produced by the same rounded operations as in scale_space.cu; only
// the data movement differs.  Reflect-101 is applied on INDICES (never by evaluating an operator on a
// mirrored extension) wherever an operator is not symmetric: Lx/Ly are antisymmetric, the diffusion
// update sums left/right/down/up in a fixed order, and the octave transition reflects in source
// coordinates.
#include "common.cuh"
#include "kernels.h"
#include <cstring>

using namespace akz;

namespace {

// =====================================================================================================
// level prep
// =====================================================================================================
constexpr int PT_W = 64, PT_H = 32;     // output tile
constexpr int PB_X = 32, PB_Y = 8;      // threads

struct PrepArgs {
    const float* src;                    // predecessor Lt (same resolution) or source octave (DOWN)
    float* ltdst;                        // DOWN: subsampled Lt
    float *flow, *lx, *ly, *det;         // flow may be null (base level)
    const float* kc;                     // per-frame contrast factor
    long long splane, plane;
    float kscale, fac1, fac2, k0, k1, k2;
    int nmul, type, step, blur;
    int sw, sh, sp;                      // source dims (== w,h,pitch unless DOWN)
    int w, h, pitch;
};

// source-coordinate reflect of a coarse index (gDownWithSmooth reflects 2x+-2, 2x+-4 in the SOURCE image)
__device__ __forceinline__ int coarse_src(int q, int sdim)
{
    int c = 2 * q;
    if (c < 0) c = -c;
    if (c >= sdim) c = sdim + sdim - 2 - c;
    return c;
}

template <bool DOWN>
__global__ void __launch_bounds__(PB_X * PB_Y) k_level_prep(const __grid_constant__ PrepArgs a)
{
    extern __shared__ float sm[];
    const int s = a.step;
    const int hb = a.blur ? 2 : 0;                    // blur halo
    const int AW = PT_W + 4 * s + 2 * hb, AH = PT_H + 4 * s + 2 * hb;   // input tile
    const int SW = PT_W + 4 * s, SH = PT_H + 4 * s;                     // smooth tile
    const int DW = PT_W + 2 * s, DH = PT_H + 2 * s;                     // Lx/Ly tiles
    float* A = sm;                                    // [AH][AW]
    float* B = A + AW * AH;                           // [AH][SW]  row-filtered
    float* S = B + SW * AH;                           // [SH][SW]
    float* LX = sm;                                   // aliases A/B once S is complete
    float* LY = LX + DW * DH;

    const int frame = blockIdx.z;
    const int X0 = blockIdx.x * PT_W, Y0 = blockIdx.y * PT_H;
    const int w = a.w, h = a.h;
    const float* src = a.src + (long long)frame * a.splane;
    const int tx = threadIdx.x, ty = threadIdx.y;

    // ---- 1. input tile, extended by index reflection -------------------------------------------------
    {
        const int ax0 = X0 - 2 * s - hb, ay0 = Y0 - 2 * s - hb;
        float* dstp = a.blur ? A : S;
        for (int r = ty; r < AH; r += PB_Y) {
            int gy = ay0 + r, sy;
            if (DOWN) sy = min(max(coarse_src(gy, a.sh), 0), a.sh - 1);
            else sy = min(max(refl(gy, h), 0), h - 1);
            const float* row = src + (long long)sy * a.sp;
            for (int c = tx; c < AW; c += PB_X) {
                int gx = ax0 + c, sx;
                if (DOWN) sx = min(max(coarse_src(gx, a.sw), 0), a.sw - 1);
                else sx = min(max(refl(gx, w), 0), w - 1);
                dstp[r * AW + c] = __ldg(row + sx);
            }
        }
    }
    __syncthreads();

    if (DOWN) {
        // subsampled plane: dst(x,y) = src(2x,2y)   (akazed.cu:505)
        float* ltd = a.ltdst + (long long)frame * a.plane;
        for (int r = ty; r < PT_H; r += PB_Y) {
            int y = Y0 + r;
            if (y >= h) break;
            for (int c = tx; c < PT_W; c += PB_X) {
                int x = X0 + c;
                if (x < w) ltd[(long long)y * a.pitch + x] = A[(r + 2 * s + hb) * AW + (c + 2 * s + hb)];
            }
        }
    }

    if (a.blur) {
        // ---- 2. row pass, evaluated at the REFLECTED column for cells outside the image -----------------
        const int sx0 = X0 - 2 * s;
        for (int r = ty; r < AH; r += PB_Y) {
            const float* Ar = A + r * AW;
            for (int c = tx; c < SW; c += PB_X) {
                int q = min(max(refl(sx0 + c, w), 0), w - 1) - (X0 - 2 * s - hb);      // column in A
                q = min(max(q, 2), AW - 3);
                B[r * SW + c] = gauss_r2(Ar[q - 2], Ar[q - 1], Ar[q], Ar[q + 1], Ar[q + 2], a.k0, a.k1, a.k2);
            }
        }
        __syncthreads();
        // ---- 3. column pass, evaluated at the REFLECTED row ---------------------------------------------
        const int sy0 = Y0 - 2 * s;
        for (int r = ty; r < SH; r += PB_Y) {
            int q = min(max(refl(sy0 + r, h), 0), h - 1) - (Y0 - 2 * s - hb);          // row in B
            q = min(max(q, 2), AH - 3);
            for (int c = tx; c < SW; c += PB_X)
                S[r * SW + c] = gauss_r2(B[(q - 2) * SW + c], B[(q - 1) * SW + c], B[q * SW + c], B[(q + 1) * SW + c], B[(q + 2) * SW + c],
                                         a.k0, a.k1, a.k2);
        }
        __syncthreads();
    }

    // ---- 4. conductance for the output pixels ----------------------------------------------------------
    if (a.flow) {
        float k = a.kc[frame];
        for (int i = 0; i < a.nmul; i++) k = __fmul_rn(k, a.kscale);
        const float ikc = __fdiv_rn(1.f, __fmul_rn(k, k));
        float* fl = a.flow + (long long)frame * a.plane;
        for (int r = ty; r < PT_H; r += PB_Y) {
            int y = Y0 + r;
            if (y >= h) break;
            const float* S1 = S + (r + 2 * s) * SW + 2 * s;
            for (int c = tx; c < PT_W; c += PB_X) {
                int x = X0 + c;
                if (x >= w) continue;
                const float* p = S1 + c;
                float ul = p[-SW - 1], uc = p[-SW], ur = p[-SW + 1], cl = p[-1], cr = p[1], ll = p[SW - 1], lc = p[SW], lr = p[SW + 1];
                float dx = scharr_dx(ul, ur, cl, cr, ll, lr);
                float dy = scharr_dy(ul, uc, ur, ll, lc, lr);
                fl[(long long)y * a.pitch + x] = conductance(a.type, __fmul_rn(grad_sq(dx, dy), ikc));
            }
        }
    }

    // ---- 5. first derivatives on the tile extended by s, computed AT the reflected position -------------
    // (LX/LY alias A/B: make sure every thread is done reading B)
    __syncthreads();
    {
        const int dx0 = X0 - s, dy0 = Y0 - s;
        float* lxg = a.lx + (long long)frame * a.plane;
        float* lyg = a.ly + (long long)frame * a.plane;
        for (int r = ty; r < DH; r += PB_Y) {
            int gy = dy0 + r;
            int qy = min(max(refl(gy, h), 0), h - 1) - (Y0 - 2 * s);                  // row in S
            qy = min(max(qy, s), SH - 1 - s);
            for (int c = tx; c < DW; c += PB_X) {
                int gx = dx0 + c;
                int qx = min(max(refl(gx, w), 0), w - 1) - (X0 - 2 * s);              // column in S
                qx = min(max(qx, s), SW - 1 - s);
                const float* p = S + qy * SW + qx;
                float ul = p[-s * SW - s], uc = p[-s * SW], ur = p[-s * SW + s], cl = p[-s], cr = p[s];
                float ll = p[s * SW - s], lc = p[s * SW], lr = p[s * SW + s];
                float vx = deriv1(sum_x(ul, ur, ll, lr), __fsub_rn(cr, cl), a.fac1, a.fac2);
                float vy = deriv1(sum_y(ul, ur, ll, lr), __fsub_rn(lc, uc), a.fac1, a.fac2);
                // LX/LY overlap A/B but not S, and S is read-only from here on
                LX[r * DW + c] = vx;
                LY[r * DW + c] = vy;
                if (gx >= X0 && gx < X0 + PT_W && gx < w && gy >= Y0 && gy < Y0 + PT_H && gy < h) {
                    lxg[(long long)gy * a.pitch + gx] = vx;
                    lyg[(long long)gy * a.pitch + gx] = vy;
                }
            }
        }
    }
    __syncthreads();

    // ---- 6. second derivatives and determinant ---------------------------------------------------------
    {
        float* dg = a.det + (long long)frame * a.plane;
        for (int r = ty; r < PT_H; r += PB_Y) {
            int y = Y0 + r;
            if (y >= h) break;
            for (int c = tx; c < PT_W; c += PB_X) {
                int x = X0 + c;
                if (x >= w) continue;
                const float* px = LX + (r + s) * DW + (c + s);
                const float* py = LY + (r + s) * DW + (c + s);
                float xul = px[-s * DW - s], xuc = px[-s * DW], xur = px[-s * DW + s], xcl = px[-s], xcr = px[s];
                float xll = px[s * DW - s], xlc = px[s * DW], xlr = px[s * DW + s];
                float yul = py[-s * DW - s], yuc = py[-s * DW], yur = py[-s * DW + s];
                float yll = py[s * DW - s], ylc = py[s * DW], ylr = py[s * DW + s];
                float dxx = deriv2(sum_x(xul, xur, xll, xlr), __fsub_rn(xcr, xcl), a.fac1, a.fac2);
                float dxy = deriv2(sum_y(xul, xur, xll, xlr), __fsub_rn(xlc, xuc), a.fac1, a.fac2);
                float dyy = deriv2(sum_y(yul, yur, yll, ylr), __fsub_rn(ylc, yuc), a.fac1, a.fac2);
                dg[(long long)y * a.pitch + x] = hess_det(dxx, dyy, dxy);
            }
        }
    }
}

size_t prep_smem(int s, int blur)
{
    int hb = blur ? 2 : 0;
    int AW = PT_W + 4 * s + 2 * hb, AH = PT_H + 4 * s + 2 * hb, SW = PT_W + 4 * s, SH = PT_H + 4 * s, DW = PT_W + 2 * s, DH = PT_H + 2 * s;
    size_t ab = (size_t)AW * AH + (size_t)SW * AH;
    size_t d2 = 2 * (size_t)DW * DH;
    size_t front = ab > d2 ? ab : d2;
    return (front + (size_t)SW * SH) * sizeof(float);
}

// =====================================================================================================
// FED cycle, temporally blocked
// =====================================================================================================
constexpr int FE_W = 128, FE_H = 64;       // shared-memory tile (outputs + halo)
constexpr int FE_TY = 4, FE_RUN = FE_H / FE_TY;
constexpr int FE_MAXK = 8;

struct FedArgs {
    const float *src, *flow;
    float* dst;
    long long plane;
    int w, h, pitch, n;
    float stepfac[FE_MAXK];
};

__global__ void __launch_bounds__(FE_W * FE_TY) k_fed(const __grid_constant__ FedArgs a)
{
    extern __shared__ float sm[];
    float* La = sm;
    float* Lb = sm + FE_W * FE_H;
    float* G = sm + 2 * FE_W * FE_H;
    const int n = a.n, w = a.w, h = a.h;
    const int TWo = FE_W - 2 * n, THo = FE_H - 2 * n;
    const int X0 = blockIdx.x * TWo, Y0 = blockIdx.y * THo;
    const int GX0 = X0 - n, GY0 = Y0 - n;
    const long long base = (long long)blockIdx.z * a.plane;
    const float* src = a.src + base;
    const float* flw = a.flow + base;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = GX0 + tx;
    const bool xin = gx >= 0 && gx < w;

    // load: column tx, all rows (coalesced along x)
    for (int r = ty; r < FE_H; r += FE_TY) {
        int gy = GY0 + r;
        float lv = 0.f, gv = 0.f;
        if (xin && gy >= 0 && gy < h) {
            long long o = (long long)gy * a.pitch + gx;
            lv = __ldg(src + o);
            gv = __ldg(flw + o);
        }
        La[r * FE_W + tx] = lv;
        G[r * FE_W + tx] = gv;
    }
    __syncthreads();

    // reflect-101 neighbour offsets of this column (akazed.cu:1251-1254)
    const int offl = (gx == 0) ? 1 : -1;
    const int offr = (gx == w - 1) ? -1 : 1;
    const int r0 = ty * FE_RUN;

    float* Lin = La;
    float* Lout = Lb;
    for (int st = 1; st <= n; st++) {
        // cells still exact after `st` steps: one ring lost per step on every side that is not an image border
        const int xlo = (GX0 <= 0) ? -GX0 : st, xhi = (GX0 + FE_W >= w) ? w - GX0 : FE_W - st;
        const int ylo = (GY0 <= 0) ? -GY0 : st, yhi = (GY0 + FE_H >= h) ? h - GY0 : FE_H - st;
        const float sf = a.stepfac[st - 1];
        if (tx >= xlo && tx < xhi) {
            int ra = max(r0, ylo), rb = min(r0 + FE_RUN, yhi);
            for (int r = ra; r < rb; r++) {
                int gy = GY0 + r;
                int up = (gy == 0) ? 1 : -1, dn = (gy == h - 1) ? -1 : 1;
                const float* Lc = Lin + r * FE_W + tx;
                const float* Gc = G + r * FE_W + tx;
                float L0 = Lc[0], g0 = Gc[0];
                Lout[r * FE_W + tx] = nld_update(L0, g0, Lc[offl], Gc[offl], Lc[offr], Gc[offr],
                                                 Lc[dn * FE_W], Gc[dn * FE_W], Lc[up * FE_W], Gc[up * FE_W], sf);
            }
        }
        __syncthreads();
        float* t = Lin; Lin = Lout; Lout = t;
    }

    // store the interior
    float* dst = a.dst + base;
    if (gx >= X0 && gx < X0 + TWo && gx < w) {
        for (int r = ty; r < FE_H; r += FE_TY) {
            int gy = GY0 + r;
            if (gy >= Y0 && gy < Y0 + THo && gy < h) dst[(long long)gy * a.pitch + gx] = Lin[r * FE_W + tx];
        }
    }
}


// -----------------------------------------------------------------------------------------------------
// k_fed2: register-blocked variant.  A 64x64 tile per CTA (512 threads), every thread owns a 4 (x) by
// 2 (y) block of pixels for the whole launch: its Lt values AND the four conductance pair sums
// g0+gL, g0+gR, g0+gD, g0+gU of every pixel stay in registers (g is frozen during a cycle, so the sums
// are computed once; fadd is commutative, so neighbouring pixels share them bit-exactly: 22 sums per
// 8 pixels).  Per step a thread exchanges only the rim of its block through shared memory (two row
// float4s, two column float2s) and evaluates 9 FP instructions per pixel (4 sub, 1 mul, 4 fma) instead
// of 13 FP + 10 shared-memory loads.  The valid region shrinks by one ring per step as in k_fed; cells
// that have lost validity keep being computed (their garbage never reaches a valid cell).
// Tile origin is aligned (x to 4, y to 2) so the image's left/top border always falls on a block edge.
// -----------------------------------------------------------------------------------------------------
constexpr int F2_T = 64;                     // tile edge
constexpr int F2_BX = 16, F2_BY = 32;        // threads: 16 x 32, block 4 x 2 pixels
constexpr int F2_CP = 66;                    // column-array pitch (floats): odd number of 8-byte units
constexpr int F2_BUF = F2_T * F2_T + 2 * F2_BX * F2_CP;      // floats per exchange buffer
constexpr int F2_SMEM = 2 * F2_BUF * (int)sizeof(float);

struct Fed2Geom { int nhx, nhy, two, tho; };
__host__ __device__ inline Fed2Geom fed2_geom(int n, int rb = 2)
{
    Fed2Geom g;                                  // rb = rows of a thread's block: the tile origin is aligned to it
    g.nhx = (n + 3) & ~3; g.nhy = (n + rb - 1) & ~(rb - 1);
    g.two = (F2_T - g.nhx - n) & ~3; g.tho = (F2_T - g.nhy - n) & ~(rb - 1);
    return g;
}

__device__ __forceinline__ float nld_update_s(float L0, float sL, float LL, float sR, float LR, float sD, float LD, float sU, float LU, float sf)
{
    float s = __fmul_rn(sL, __fsub_rn(LL, L0));
    s = __fmaf_rn(sR, __fsub_rn(LR, L0), s);
    s = __fmaf_rn(sD, __fsub_rn(LD, L0), s);
    s = __fmaf_rn(sU, __fsub_rn(LU, L0), s);
    return __fmaf_rn(s, sf, L0);
}

template <bool BORDER>
__device__ __forceinline__ void fed2_body(const FedArgs& a, float* sm, int vec_ok)
{
    const int n = a.n, w = a.w, h = a.h;
    const Fed2Geom ge = fed2_geom(n);
    const int X0 = blockIdx.x * ge.two, Y0 = blockIdx.y * ge.tho;
    const int GX0 = X0 - ge.nhx, GY0 = Y0 - ge.nhy;
    const long long base = (long long)blockIdx.z * a.plane;
    const float* __restrict__ src = a.src + base;
    const float* __restrict__ flw = a.flow + base;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * F2_BX + tx;
    const int bx = 4 * tx, by = 2 * ty;                 // block origin in the tile
    const int gx0 = GX0 + bx, gy0 = GY0 + by;

    float* T0 = sm;                                     // buffer 0: tile [64][64], CL [16][66], CR [16][66]
    float* T1 = sm + F2_BUF;                            // buffer 1 (its tile doubles as the G staging area)
    float* Gs = T1;

    // ---- conductance tile -> shared (reflected indices: pair sums at the image border come out right)
    float L[2][4];
    if (!BORDER && vec_ok) {
        for (int i = tid; i < F2_T * (F2_T / 4); i += F2_BX * F2_BY) {
            int r = i >> 4, c4 = (i & 15) * 4;
            *(float4*)(Gs + r * F2_T + c4) = __ldg((const float4*)(flw + (long long)(GY0 + r) * a.pitch + GX0 + c4));
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            float4 v = __ldg((const float4*)(src + (long long)(gy0 + r) * a.pitch + gx0));
            L[r][0] = v.x; L[r][1] = v.y; L[r][2] = v.z; L[r][3] = v.w;
        }
    } else {
        for (int i = tid; i < F2_T * F2_T; i += F2_BX * F2_BY) {
            int r = i >> 6, c = i & 63;
            int sy = min(max(refl(GY0 + r, h), 0), h - 1), sx = min(max(refl(GX0 + c, w), 0), w - 1);
            Gs[r * F2_T + c] = __ldg(flw + (long long)sy * a.pitch + sx);
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            int sy = min(max(refl(gy0 + r, h), 0), h - 1);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int sx = min(max(refl(gx0 + c, w), 0), w - 1);
                L[r][c] = __ldg(src + (long long)sy * a.pitch + sx);
            }
        }
    }
    __syncthreads();

    // ---- pair sums: sh[r][j] = g(r, j-1) + g(r, j), j = 0..4 ; sv[k][c] = g(k-1, c) + g(k, c), k = 0..2
    float sh[2][5], sv[3][4];
    {
        float g[4][6];                                   // rows by-1..by+2, cols bx-1..bx+4 (clamped at the tile rim)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int rr = min(max(by - 1 + k, 0), F2_T - 1);
            const float* gr = Gs + rr * F2_T;
            float4 v = *(const float4*)(gr + bx);
            g[k][0] = gr[max(bx - 1, 0)]; g[k][1] = v.x; g[k][2] = v.y; g[k][3] = v.z; g[k][4] = v.w; g[k][5] = gr[min(bx + 4, F2_T - 1)];
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int j = 0; j < 5; j++) sh[r][j] = __fadd_rn(g[r + 1][j + 1], g[r + 1][j]);
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int c = 0; c < 4; c++) sv[k][c] = __fadd_rn(g[k + 1][c + 1], g[k][c + 1]);
    }

    // border bookkeeping (BORDER tiles only)
    bool bl = false, bt = false;
    int ir = -1, jb = -1;
    if (BORDER) {
        bl = (gx0 == 0); bt = (gy0 == 0);
        ir = (w - 1) - gx0; if (ir < 0 || ir > 3) ir = -1;
        jb = (h - 1) - gy0; if (jb < 0 || jb > 1) jb = -1;
    }

    const int txl = max(tx - 1, 0), txr = min(tx + 1, F2_BX - 1);
    const int ru = max(by - 1, 0), rd = min(by + 2, F2_T - 1);
    const int o_row0 = by * F2_T + bx, o_up = ru * F2_T + bx, o_dn = rd * F2_T + bx;
    const int o_cl = F2_T * F2_T + tx * F2_CP + by, o_cr = o_cl + F2_BX * F2_CP;
    const int o_lf = F2_T * F2_T + F2_BX * F2_CP + txl * F2_CP + by;      // CR of the left neighbour
    const int o_rt = F2_T * F2_T + txr * F2_CP + by;                      // CL of the right neighbour

    // publish the rim of the block in buffer 0
    *(float4*)(T0 + o_row0) = make_float4(L[0][0], L[0][1], L[0][2], L[0][3]);
    *(float4*)(T0 + o_row0 + F2_T) = make_float4(L[1][0], L[1][1], L[1][2], L[1][3]);
    *(float2*)(T0 + o_cl) = make_float2(L[0][0], L[1][0]);
    *(float2*)(T0 + o_cr) = make_float2(L[0][3], L[1][3]);
    __syncthreads();

    float* cur = T0;
    float* nxt = T1;
    for (int st = 0; st < n; st++) {
        const float sf = a.stepfac[st];
        const float4 up4 = *(const float4*)(cur + o_up);
        const float4 dn4 = *(const float4*)(cur + o_dn);
        const float2 lf2 = *(const float2*)(cur + o_lf);
        const float2 rt2 = *(const float2*)(cur + o_rt);
        const float up[4] = { up4.x, up4.y, up4.z, up4.w }, dn[4] = { dn4.x, dn4.y, dn4.z, dn4.w };
        const float lf[2] = { lf2.x, lf2.y }, rt[2] = { rt2.x, rt2.y };
        float N[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float LL = (c == 0) ? lf[r] : L[r][c - 1];
                float LR = (c == 3) ? rt[r] : L[r][c + 1];
                float LU = (r == 0) ? up[c] : L[0][c];
                float LD = (r == 1) ? dn[c] : L[1][c];
                if (BORDER) {
                    if (c == 0 && bl) LL = LR;                 // x = 0: left neighbour is x = 1
                    if (c == ir) LR = LL;                      // x = w-1: right neighbour is x = w-2
                    if (r == 0 && bt) LU = LD;                 // y = 0
                    if (r == jb) LD = LU;                      // y = h-1
                }
                N[r][c] = nld_update_s(L[r][c], sh[r][c], LL, sh[r][c + 1], LR, sv[r + 1][c], LD, sv[r][c], LU, sf);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) L[r][c] = N[r][c];
        if (st + 1 < n) {
            *(float4*)(nxt + o_row0) = make_float4(L[0][0], L[0][1], L[0][2], L[0][3]);
            *(float4*)(nxt + o_row0 + F2_T) = make_float4(L[1][0], L[1][1], L[1][2], L[1][3]);
            *(float2*)(nxt + o_cl) = make_float2(L[0][0], L[1][0]);
            *(float2*)(nxt + o_cr) = make_float2(L[0][3], L[1][3]);
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
    }

    // ---- store the output region of the tile
    if (gx0 >= X0 && gx0 < X0 + ge.two && gx0 < w) {
        float* dst = a.dst + base;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            int gy = gy0 + r;
            if (gy >= Y0 && gy < Y0 + ge.tho && gy < h) {
                float* d = dst + (long long)gy * a.pitch + gx0;
                if (vec_ok && gx0 + 3 < w) *(float4*)d = make_float4(L[r][0], L[r][1], L[r][2], L[r][3]);
                else {
#pragma unroll
                    for (int c = 0; c < 4; c++) if (gx0 + c < w) d[c] = L[r][c];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(F2_BX * F2_BY, 2) k_fed2(const __grid_constant__ FedArgs a, int vec_ok)
{
    extern __shared__ __align__(16) float sm[];
    const Fed2Geom ge = fed2_geom(a.n);
    const int GX0 = blockIdx.x * ge.two - ge.nhx, GY0 = blockIdx.y * ge.tho - ge.nhy;
    const bool border = GX0 <= 0 || GY0 <= 0 || GX0 + F2_T >= a.w || GY0 + F2_T >= a.h;
    if (border) fed2_body<true>(a, sm, vec_ok);
    else fed2_body<false>(a, sm, vec_ok);
}

// -----------------------------------------------------------------------------------------------------
// k_fed3: k_fed2 made persistent with a software prefetch.  ncu on k_fed2 (profiles/r01a_k_fed2_raw.txt):
// issue slots 51 % busy, 5.1 warps per issue waiting on the global loads of the tile prologue -- with two
// register-limited CTAs per SM nothing overlaps a tile's load latency.  Here each CTA walks a strided list
// of tiles; while it runs the steps of tile t, cp.async copies the Lt and g tiles of tile t+1 into a
// shared-memory staging area (interior tiles; border tiles fill the stage with reflected scalar loads).
// Arithmetic and tiling are those of k_fed2.
// -----------------------------------------------------------------------------------------------------
constexpr int F3_STAGE = 2 * F2_T * F2_T;                               // floats: Lt tile, g tile
constexpr int F3_SMEM = (2 * F2_BUF + F3_STAGE) * (int)sizeof(float);

struct Fed3Args {
    FedArgs f;
    int gx, gy, ntiles, vec_ok;
    int nhx, nhy, two, tho;                       // fed2_geom(n), computed on the host
    unsigned long long inv_per, inv_gx;           // floor(2^40 / d) + 1: t / d == (t * inv) >> 40 for t < 2^24, d < 2^16
};

__device__ __forceinline__ int f3_div(int t, unsigned long long inv) { return (int)(((unsigned long long)(unsigned)t * inv) >> 40); }

__device__ __forceinline__ void f3_cp_async16(float* smem_dst, const float* gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

struct F3Tile { int X0, Y0, GX0, GY0; long long base; bool border; };

// (ncu source view of the first version: the two integer divisions of this decode were 14 % of all executed
// instructions of the kernel, hence the multiplicative inverses)
__device__ __forceinline__ F3Tile f3_decode(const Fed3Args& a, int t)
{
    F3Tile T;
    int per = a.gx * a.gy;
    int frame = f3_div(t, a.inv_per), rem = t - frame * per;
    int by = f3_div(rem, a.inv_gx), bx = rem - by * a.gx;
    T.X0 = bx * a.two; T.Y0 = by * a.tho;
    T.GX0 = T.X0 - a.nhx; T.GY0 = T.Y0 - a.nhy;
    T.base = (long long)frame * a.f.plane;
    T.border = T.GX0 <= 0 || T.GY0 <= 0 || T.GX0 + F2_T >= a.f.w || T.GY0 + F2_T >= a.f.h || !a.vec_ok;
    return T;
}

// fill the staging area with the Lt and g tiles of T (asynchronously for interior tiles)
template <int NT>
__device__ __forceinline__ void f3_stage(const Fed3Args& a, const F3Tile& T, float* St, int tid)
{
    const float* __restrict__ src = a.f.src + T.base;
    const float* __restrict__ flw = a.f.flow + T.base;
    if (!T.border) {
        for (int i = tid; i < F2_T * (F2_T / 4); i += NT) {
            int r = i >> 4, c4 = (i & 15) * 4;
            long long o = (long long)(T.GY0 + r) * a.f.pitch + T.GX0 + c4;
            f3_cp_async16(St + r * F2_T + c4, src + o);
            f3_cp_async16(St + F2_T * F2_T + r * F2_T + c4, flw + o);
        }
    } else {
        // border tile: rows by reflected index; a group of 4 columns inside the image is one float4 load per plane
        const int w = a.f.w, h = a.f.h;
        for (int i = tid; i < F2_T * (F2_T / 4); i += NT) {
            int r = i >> 4, c4 = (i & 15) * 4;
            int sy = min(max(refl(T.GY0 + r, h), 0), h - 1), gx = T.GX0 + c4;
            long long ro = (long long)sy * a.f.pitch;
            float* dl = St + r * F2_T + c4;
            float* dg = dl + F2_T * F2_T;
            if (a.vec_ok && gx >= 0 && gx + 3 < w) {
                *(float4*)dl = __ldg((const float4*)(src + ro + gx));
                *(float4*)dg = __ldg((const float4*)(flw + ro + gx));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int sx = min(max(refl(gx + j, w), 0), w - 1);
                    dl[j] = __ldg(src + ro + sx);
                    dg[j] = __ldg(flw + ro + sx);
                }
            }
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

// INT = true: the integer pipeline's step (gNldStepNaive akazed.cu:3449-3473) on int32 planes; the values travel through
// the same float registers / shared-memory buffers as bit patterns (only moves, loads and stores touch them).
template <bool INT>
__device__ __forceinline__ float f3_sum(float a, float b)
{
    if (INT) return __int_as_float(__float_as_int(a) + __float_as_int(b));
    return __fadd_rn(a, b);
}
template <bool INT>
__device__ __forceinline__ float f3_upd(float L0, float sL, float LL, float sR, float LR, float sD, float LD, float sU, float LU, float sf)
{
    if (!INT) return nld_update_s(L0, sL, LL, sR, LR, sD, LD, sU, LU, sf);
    const int l0 = __float_as_int(L0);
    const int step = (__float_as_int(sR) * (__float_as_int(LR) - l0) + __float_as_int(sL) * (__float_as_int(LL) - l0) +
                      __float_as_int(sD) * (__float_as_int(LD) - l0) + __float_as_int(sU) * (__float_as_int(LU) - l0)) >> 16;
    return __int_as_float(((__float_as_int(sf) * step) >> 16) + l0);
}

// RB = rows of the block a thread owns (4 x RB pixels).  RB = 4 halves the number of threads per tile: the per-thread fixed
// work of a tile (decode, addresses, border flags, rim exchange: two thirds of the executed instructions at 3-4 steps, ncu
// r01i per-line table) is paid once per 16 pixels instead of once per 8.
template <bool INT, int RB>
__global__ void __launch_bounds__(F2_BX * (F2_T / RB), 2) k_fed3(const __grid_constant__ Fed3Args a)
{
    constexpr int NT = F2_BX * (F2_T / RB);
    extern __shared__ __align__(16) float sm[];
    float* T0 = sm;
    float* T1 = sm + F2_BUF;
    float* St = sm + 2 * F2_BUF;
    float* Gs = St + F2_T * F2_T;
    const int n = a.f.n, w = a.f.w, h = a.f.h;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * F2_BX + tx;
    const int bx = 4 * tx, by = RB * ty;

    const int txl = max(tx - 1, 0), txr = min(tx + 1, F2_BX - 1);
    const int ru = max(by - 1, 0), rd = min(by + RB, F2_T - 1);
    const int o_row0 = by * F2_T + bx, o_up = ru * F2_T + bx, o_dn = rd * F2_T + bx;
    const int o_cl = F2_T * F2_T + tx * F2_CP + by, o_cr = o_cl + F2_BX * F2_CP;
    const int o_lf = F2_T * F2_T + F2_BX * F2_CP + txl * F2_CP + by;      // CR of the left neighbour
    const int o_rt = F2_T * F2_T + txr * F2_CP + by;                      // CL of the right neighbour

    // rim of the block: first and last row as float4, first and last column as RB / 2 float2
    auto publish = [&](float* buf, const float (&L)[RB][4]) {
        *(float4*)(buf + o_row0) = make_float4(L[0][0], L[0][1], L[0][2], L[0][3]);
        *(float4*)(buf + o_row0 + (RB - 1) * F2_T) = make_float4(L[RB - 1][0], L[RB - 1][1], L[RB - 1][2], L[RB - 1][3]);
#pragma unroll
        for (int r = 0; r < RB; r += 2) {
            *(float2*)(buf + o_cl + r) = make_float2(L[r][0], L[r + 1][0]);
            *(float2*)(buf + o_cr + r) = make_float2(L[r][3], L[r + 1][3]);
        }
    };

    int t = blockIdx.x;
    if (t < a.ntiles) { F3Tile Tn = f3_decode(a, t); f3_stage<NT>(a, Tn, St, tid); }
    for (; t < a.ntiles; t += gridDim.x) {
        const F3Tile T = f3_decode(a, t);
        const int gx0 = T.GX0 + bx, gy0 = T.GY0 + by;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                                                   // stage(t) complete and visible

        float L[RB][4];
#pragma unroll
        for (int r = 0; r < RB; r++) {
            float4 v = *(const float4*)(St + (by + r) * F2_T + bx);
            L[r][0] = v.x; L[r][1] = v.y; L[r][2] = v.z; L[r][3] = v.w;
        }
        float sh[RB][5], sv[RB + 1][4];
        {
            float g[RB + 2][6];
#pragma unroll
            for (int k = 0; k < RB + 2; k++) {
                int rr = min(max(by - 1 + k, 0), F2_T - 1);
                const float* gr = Gs + rr * F2_T;
                float4 v = *(const float4*)(gr + bx);
                g[k][0] = gr[max(bx - 1, 0)]; g[k][1] = v.x; g[k][2] = v.y; g[k][3] = v.z; g[k][4] = v.w; g[k][5] = gr[min(bx + 4, F2_T - 1)];
            }
#pragma unroll
            for (int r = 0; r < RB; r++)
#pragma unroll
                for (int j = 0; j < 5; j++) sh[r][j] = f3_sum<INT>(g[r + 1][j + 1], g[r + 1][j]);
#pragma unroll
            for (int k = 0; k < RB + 1; k++)
#pragma unroll
                for (int c = 0; c < 4; c++) sv[k][c] = f3_sum<INT>(g[k + 1][c + 1], g[k][c + 1]);
        }
        // publish the rim of the block in buffer 0 (its last readers finished before the barrier above)
        publish(T0, L);
        __syncthreads();                                                   // stage fully consumed, rim visible
        if (t + (int)gridDim.x < a.ntiles) { F3Tile Tn = f3_decode(a, t + gridDim.x); f3_stage<NT>(a, Tn, St, tid); }

        // image-border bookkeeping (only tiles that touch the border pay for it)
        const bool border = T.GX0 <= 0 || T.GY0 <= 0 || T.GX0 + F2_T >= w || T.GY0 + F2_T >= h;
        bool bl = false, bt = false;
        int ir = -1, jb = -1;
        if (border) {
            bl = (gx0 == 0); bt = (gy0 == 0);
            ir = (w - 1) - gx0; if (ir < 0 || ir > 3) ir = -1;
            jb = (h - 1) - gy0; if (jb < 0 || jb > RB - 1) jb = -1;
        }

        float* cur = T0;
        float* nxt = T1;
        for (int st = 0; st < n; st++) {
            const float sf = a.f.stepfac[st];
            const float4 up4 = *(const float4*)(cur + o_up);
            const float4 dn4 = *(const float4*)(cur + o_dn);
            float lf[RB], rt[RB];
#pragma unroll
            for (int r = 0; r < RB; r += 2) {
                const float2 l2 = *(const float2*)(cur + o_lf + r), r2 = *(const float2*)(cur + o_rt + r);
                lf[r] = l2.x; lf[r + 1] = l2.y; rt[r] = r2.x; rt[r + 1] = r2.y;
            }
            const float up[4] = { up4.x, up4.y, up4.z, up4.w }, dn[4] = { dn4.x, dn4.y, dn4.z, dn4.w };
            float N[RB][4];
            if (!border) {
#pragma unroll
                for (int r = 0; r < RB; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float LL = (c == 0) ? lf[r] : L[r][c - 1];
                        float LR = (c == 3) ? rt[r] : L[r][c + 1];
                        float LU = (r == 0) ? up[c] : L[r - 1 < 0 ? 0 : r - 1][c];
                        float LD = (r == RB - 1) ? dn[c] : L[r + 1 > RB - 1 ? RB - 1 : r + 1][c];
                        N[r][c] = f3_upd<INT>(L[r][c], sh[r][c], LL, sh[r][c + 1], LR, sv[r + 1][c], LD, sv[r][c], LU, sf);
                    }
            } else {
#pragma unroll
                for (int r = 0; r < RB; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float LL = (c == 0) ? lf[r] : L[r][c - 1];
                        float LR = (c == 3) ? rt[r] : L[r][c + 1];
                        float LU = (r == 0) ? up[c] : L[r - 1 < 0 ? 0 : r - 1][c];
                        float LD = (r == RB - 1) ? dn[c] : L[r + 1 > RB - 1 ? RB - 1 : r + 1][c];
                        if (c == 0 && bl) LL = LR;                 // x = 0: left neighbour is x = 1
                        if (c == ir) LR = LL;                      // x = w-1: right neighbour is x = w-2
                        if (r == 0 && bt) LU = LD;                 // y = 0
                        if (r == jb) LD = LU;                      // y = h-1
                        N[r][c] = f3_upd<INT>(L[r][c], sh[r][c], LL, sh[r][c + 1], LR, sv[r + 1][c], LD, sv[r][c], LU, sf);
                    }
            }
#pragma unroll
            for (int r = 0; r < RB; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) L[r][c] = N[r][c];
            if (st + 1 < n) {
                publish(nxt, L);
                __syncthreads();
                float* tt = cur; cur = nxt; nxt = tt;
            }
        }

        // store the output region of the tile
        if (gx0 >= T.X0 && gx0 < T.X0 + a.two && gx0 < w) {
            float* dst = a.f.dst + T.base;
#pragma unroll
            for (int r = 0; r < RB; r++) {
                int gy = gy0 + r;
                if (gy >= T.Y0 && gy < T.Y0 + a.tho && gy < h) {
                    float* d = dst + (long long)gy * a.f.pitch + gx0;
                    if (a.vec_ok && gx0 + 3 < w) *(float4*)d = make_float4(L[r][0], L[r][1], L[r][2], L[r][3]);
                    else {
#pragma unroll
                        for (int c = 0; c < 4; c++) if (gx0 + c < w) d[c] = L[r][c];
                    }
                }
            }
        }
    }
}

akz_once_t g_attr_done;
int g_fed_rb = 4;                            // rows per thread block of k_fed3 (AKZ_FED_RB=2 selects the 4 x 2 variant)

}  // namespace

namespace akzk {

static void set_attrs()
{
    akz_once_guard once{g_attr_done};
    if (!once) return;
    if (const char* e = getenv("AKZ_FED_RB")) g_fed_rb = atoi(e) == 2 ? 2 : 4;
    cudaFuncSetAttribute(k_fed, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * FE_W * FE_H * (int)sizeof(float));
    cudaFuncSetAttribute(k_fed2, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM);
    cudaFuncSetAttribute(k_fed3<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_level_prep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_level_prep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

static int prep_common(cudaStream_t st, PrepArgs& a, bool down, int n)
{
    set_attrs();
    hessian_factors(&a.fac1, &a.fac2);
    float k[3];
    akz_gauss_taps(1.f, 2, k);
    a.k0 = k[0]; a.k1 = k[1]; a.k2 = k[2];
    size_t smem = prep_smem(a.step, a.blur);
    if (smem > 160 * 1024) return akz_set_error(AKZ_E_UNSUPPORTED, "derivative step %d too large for the fused level kernel", a.step);
    dim3 g((a.w + PT_W - 1) / PT_W, (a.h + PT_H - 1) / PT_H, n), b(PB_X, PB_Y);
    if (down) k_level_prep<true><<<g, b, smem, st>>>(a);
    else k_level_prep<false><<<g, b, smem, st>>>(a);
    return 1;
}

int level_prep(cudaStream_t st, const float* ltprev, float* flowp, float* lx, float* ly, float* det, int blur, int type,
               const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n)
{
    PrepArgs a = {};
    a.src = ltprev; a.ltdst = nullptr; a.flow = flowp; a.lx = lx; a.ly = ly; a.det = det; a.kc = kc;
    a.splane = plane; a.plane = plane; a.kscale = kscale; a.nmul = nmul; a.type = type; a.step = step; a.blur = blur;
    a.sw = w; a.sh = h; a.sp = pitch; a.w = w; a.h = h; a.pitch = pitch;
    return prep_common(st, a, false, n);
}

int level_prep_down(cudaStream_t st, const float* ltsrc, int sw, int sh, int sp, long long splane,
                    float* ltdst, float* flowp, float* lx, float* ly, float* det, int type,
                    const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n)
{
    PrepArgs a = {};
    a.src = ltsrc; a.ltdst = ltdst; a.flow = flowp; a.lx = lx; a.ly = ly; a.det = det; a.kc = kc;
    a.splane = splane; a.plane = plane; a.kscale = kscale; a.nmul = nmul; a.type = type; a.step = step; a.blur = 1;
    a.sw = sw; a.sh = sh; a.sp = sp; a.w = w; a.h = h; a.pitch = pitch;
    return prep_common(st, a, true, n);
}

// dst receives the result of n steps applied to src; tmp is a scratch plane batch (never aliases src/dst).
// fused == 0: n single-step launches; fused == 1: ceil(n/8) launches of the persistent, prefetching k_fed3;
// fused == 2: the same split with the shared-memory-resident k_fed, fused == 3: with k_fed2 (kept as cross-checks).
int fed_cycle(cudaStream_t st, const float* src, const float* flowp, float* dst, float* tmp, const float* tau, int nsteps,
              int w, int h, int pitch, long long plane, int n, int fused, int int_planes)
{
    if (nsteps <= 0) return 0;
    set_attrs();
    int launches = 0;
    if (!fused) {
        const float* cur = src;
        for (int k = 0; k < nsteps; k++) {
            float* out = ((nsteps - 1 - k) % 2 == 0) ? dst : tmp;
            nld_step(st, cur, flowp, out, tau[k], w, h, pitch, plane, n);
            cur = out;
            launches++;
        }
        return launches;
    }
    int m = (nsteps + FE_MAXK - 1) / FE_MAXK;
    int done = 0;
    const float* cur = src;
    for (int i = 0; i < m; i++) {
        int cnt = (nsteps - done + (m - i) - 1) / (m - i);          // balanced split
        FedArgs a = {};
        a.src = cur; a.flow = flowp; a.dst = ((m - 1 - i) % 2 == 0) ? dst : tmp;
        a.plane = plane; a.w = w; a.h = h; a.pitch = pitch; a.n = cnt;
        for (int k = 0; k < cnt; k++) {
            if (int_planes) { int sfi = (int)(0.5f * tau[done + k] * 65536 + 0.5f); memcpy(&a.stepfac[k], &sfi, 4); }      // akazed.cu:4240
            else a.stepfac[k] = 0.5f * tau[done + k];                                                                  // akazed.cu:2515
        }
        if (fused == 2) {
            int TWo = FE_W - 2 * cnt, THo = FE_H - 2 * cnt;
            dim3 g((w + TWo - 1) / TWo, (h + THo - 1) / THo, n), b(FE_W, FE_TY);
            k_fed<<<g, b, 3 * FE_W * FE_H * sizeof(float), st>>>(a);
        } else {
            const int rb = (fused == 3 || g_fed_rb == 2) ? 2 : 4;
            Fed2Geom ge = fed2_geom(cnt, rb);
            dim3 g((w + ge.two - 1) / ge.two, (h + ge.tho - 1) / ge.tho, n), b(F2_BX, F2_T / rb);
            int vec_ok = (pitch % 4 == 0) && (plane % 4 == 0) && (((uintptr_t)a.src | (uintptr_t)flowp | (uintptr_t)a.dst) % 16 == 0);
            if (fused == 3) k_fed2<<<g, b, F2_SMEM, st>>>(a, vec_ok);
            else {
                Fed3Args a3;
                a3.f = a; a3.gx = g.x; a3.gy = g.y; a3.ntiles = (int)(g.x * g.y * g.z); a3.vec_ok = vec_ok;
                a3.nhx = ge.nhx; a3.nhy = ge.nhy; a3.two = ge.two; a3.tho = ge.tho;
                a3.inv_per = (1ull << 40) / (unsigned long long)(g.x * g.y) + 1; a3.inv_gx = (1ull << 40) / (unsigned long long)g.x + 1;
                if (a3.ntiles >= (1 << 24) || g.x * g.y >= (1u << 16)) return akz_set_error(AKZ_E_UNSUPPORTED, "FED tile count out of range");
                int nb = a3.ntiles < 2 * 148 ? a3.ntiles : 2 * 148;
                if (rb == 4) {
                    if (int_planes) k_fed3<true, 4><<<nb, b, F3_SMEM, st>>>(a3);
                    else k_fed3<false, 4><<<nb, b, F3_SMEM, st>>>(a3);
                } else {
                    if (int_planes) k_fed3<true, 2><<<nb, b, F3_SMEM, st>>>(a3);
                    else k_fed3<false, 2><<<nb, b, F3_SMEM, st>>>(a3);
                }
            }
        }
        cur = a.dst;
        done += cnt;
        launches++;
    }
    return launches;
}

}  // namespace akzk
