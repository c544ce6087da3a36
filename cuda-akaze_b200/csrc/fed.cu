// Explicit-diffusion (FED) cycles of the nonlinear scale space: n steps of gNldStepNaive (akazed.cu:1241-1264; integer twin
// :3449-3473) with the conductance frozen, plus the n device-to-device copies of akaze.cpp:386,415, in as few launches as
// possible.  Two kernels, same pinned arithmetic (results are bit-identical to n single-step launches):
//
//   k_fed4<N>  PRODUCTION.  Streaming, register-resident, no shared memory, no block barrier.  A WARP owns a strip of 128
//              columns (4 per lane; one halo lane at either end, 120 columns of output) and marches down a band of rows.  At
//              row time t it loads row t of Lt and g (two 16-byte loads per lane, prefetched), forms the conductance pair sums
//              of that row once (g is frozen during a cycle and fadd is commutative, so neighbouring pixels share the sums
//              bit-exactly) and advances a wavefront of the N steps: step s produces its row t - s from the three rows
//              t-s-1 .. t-s+1 of step s - 1 held in rotating register windows, left / right neighbours come from the adjacent
//              lanes by shuffle, and row t - N of the last step is stored.  Per step-pixel that is the 9 FP instructions the
//              pinned arithmetic needs plus ~2 of data movement, against 28 in the tile kernel (ncu r01l: two thirds of its
//              instructions were per-tile fixed work and rim exchange at the 3-4 steps of octave 0); recomputation is 8 of 128
//              columns and 2 N rows per band instead of a ring of n cells around every 64 x 64 tile.
//   k_fed3     tile kernel with temporal blocking in shared memory (round 1).  Kept for planes the streaming kernel does
//              not take: rows that are not 16-byte aligned (stage seams called with arbitrary pitches).
#include "common.cuh"
#include "kernels.h"
#include <cstring>

using namespace akz;

namespace {

constexpr int FE_MAXK = 8;                   // steps per launch of the tile kernel

struct FedArgs {
    const float *src, *flow;
    float* dst;
    long long plane;
    int w, h, pitch, n;
    float stepfac[FE_MAXK];
};

// ---- tile geometry of k_fed3: a 64x64 tile per CTA, a thread owns a 4 (x) by RB (y) block of pixels for the whole launch ----
constexpr int F2_T = 64;                     // tile edge
constexpr int F2_BX = 16;                    // threads in x: blocks of 4 pixels
constexpr int F2_CP = 66;                    // column-array pitch (floats): odd number of 8-byte units
constexpr int F2_BUF = F2_T * F2_T + 2 * F2_BX * F2_CP;      // floats per exchange buffer

struct Fed2Geom { int nhx, nhy, two, tho; };
inline Fed2Geom fed2_geom(int n, int rb)
{
    Fed2Geom g;                                  // rb = rows of a thread's block: the tile origin is aligned to it
    g.nhx = (n + 3) & ~3; g.nhy = (n + rb - 1) & ~(rb - 1);
    g.two = (F2_T - g.nhx - n) & ~3; g.tho = (F2_T - g.nhy - n) & ~(rb - 1);
    return g;
}

// one explicit step of a pixel from the pair sums s* = g0 + g* and the four neighbours (akazed.cu:1257-1262: left product
// rounded, then right, down, up fused)
__device__ __forceinline__ float nld_update_s(float L0, float sL, float LL, float sR, float LR, float sD, float LD, float sU, float LU, float sf)
{
    float s = __fmul_rn(sL, __fsub_rn(LL, L0));
    s = __fmaf_rn(sR, __fsub_rn(LR, L0), s);
    s = __fmaf_rn(sD, __fsub_rn(LD, L0), s);
    s = __fmaf_rn(sU, __fsub_rn(LU, L0), s);
    return __fmaf_rn(s, sf, L0);
}

// =====================================================================================================
// k_fed3: persistent tile kernel.  Each CTA walks a strided list of tiles; while it runs the steps of tile t, cp.async
// copies the Lt and g tiles of tile t+1 into a shared-memory staging area.  Up to 8 steps per launch; the valid region
// shrinks by one ring per step.
// =====================================================================================================
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

// =====================================================================================================
// k_fed4: streaming warp kernel (see the header comment)
// =====================================================================================================
constexpr int F4_COLS = 120;        // output columns of a strip: lanes 1..30 x 4 (lanes 0 and 31 carry the halo)
constexpr int F4_WARPS = 4;         // warps per CTA; every warp works on its own (strip, band, frame) unit
constexpr int F4_MAXN = 4;          // steps per launch: the halo of 4 columns is used up by 4 steps
constexpr int F4_RING = 6;          // rows of Lt / g per warp in the shared-memory landing ring: row t + 5 is requested while row t is used
constexpr int F4_ROWB = 2 * 32 * 16; // bytes of one ring slot: 32 lanes x 16 bytes of Lt, then 32 x 16 bytes of g (conflict-free LDS.128)

struct Fed4Args {
    const float *src, *flow;
    float* dst;
    long long plane;
    int w, h, pitch;
    int nstrips, nbands, band_h, nunits;
    float stepfac[F4_MAXN];
};

// register state of a warp lane: everything is indexed with compile-time constants (the row loop is unrolled by 6)
template <int N>
struct Fed4Regs {
    float L0[3][4];                 // input rows t-2, t-1, t (ring by row time mod 3)
    float G[2][4];                  // conductance rows t-1, t (ring by row time mod 2)
    float Lr[N][3][4];              // Lr[s-1] = rows of step s = 1 .. N-1, ring by production time mod 3 (the last entry is unused)
    float SH[6][5], SV[6][4];       // pair sums of row t: SH[j] = g(x+j) + g(x+j-1), SV[c] = g(y) + g(y-1); ring by row time mod 6
    float NL[N], NR[N];             // left / right neighbour (adjacent lanes) of the newest row of step s = 0 .. N-1: shuffled when the
                                    // row is produced, used one row time later when that row is the centre row of step s + 1
};

struct Fed4Lane {
    const float *psrc, *pflow;      // frame base + clamped column of this lane
    float* pdst;                    // frame base + column of this lane
    unsigned ring;                  // shared-memory address of this lane's 16 bytes of Lt in slot 0 of its warp's landing ring (g: + 512)
    // TMA variant (A/B, AKZ_FED_TMA=1): the row segment of the whole warp is one bulk copy per plane issued by lane 0
    const float *wsrc, *wflow;      // frame base + first in-row column of the warp's segment
    unsigned wring, mbar;           // shared-memory address of the segment's first byte in slot 0; of the warp's six mbarriers
    unsigned wbytes;                // bytes of the segment that lie inside the row
    unsigned par;                   // parity of the current pass over the ring
    int y0, y1, t0;                 // band: output rows [y0, y1), first row time
    int ir;                         // GEN: index of the image's last column inside this lane's four (-1: none)
    bool bl, br, store;             // lane holds x = 0 / x = w-1 (as its last column) / lane writes output
};

// Rows travel global -> shared memory with cp.async (LDGSTS): 16 bytes of Lt and of g per lane and row, five rows ahead of the
// row being consumed.  A lane only ever reads back the bytes it copied itself, so cp.async.wait_group is all the
// synchronisation there is (no barrier, no cross-lane visibility).  Five rows x 1 KB x 16 warps = 80 KB in flight per SM:
// with loads issued from registers two rows ahead (first version) the kernel sat at 2.7 TB/s, bound by the memory latency.
template <bool TMA>
__device__ __forceinline__ void f4_request(const Fed4Args& a, const Fed4Lane& ln, int row_time, int slot)
{
    const int row = min(max(row_time, 0), a.h - 1);
    const long long o = (long long)row * a.pitch;
    if (!TMA) {
        const unsigned d = ln.ring + slot * F4_ROWB;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(ln.psrc + o) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + 512u), "l"(ln.pflow + o) : "memory");
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    } else {
        // one elected lane: expect 2 x wbytes on the slot's mbarrier, then one bulk copy (TMA, SASS UBLKCP) per plane
        if ((threadIdx.x & 31) == 0) {
            const unsigned mb = ln.mbar + slot * 8, d = ln.wring + slot * F4_ROWB;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(2u * ln.wbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(d), "l"(ln.wsrc + o), "r"(ln.wbytes), "r"(mb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(d + 512u), "l"(ln.wflow + o), "r"(ln.wbytes), "r"(mb) : "memory");
        }
    }
}

template <bool TMA>
__device__ __forceinline__ void f4_wait(const Fed4Lane& ln, int slot)
{
    if (!TMA) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(F4_RING - 1) : "memory");
    } else {
        const unsigned mb = ln.mbar + slot * 8;
        asm volatile("{\n.reg .pred p;\nAKZ_W_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra AKZ_W_%=;\n}\n" ::"r"(mb), "r"(ln.par) : "memory");
    }
}

// one row time: request row t + 5, take row t out of the landing ring, form its pair sums, then step s = 1..N advances to its
// row t - s.
// PH = (row time - first row time) mod 6.  SLOW = this row time touches the first / last image row or lies in the warm-up /
// padding of the band: border substitutions are applied and the store is range-checked.  Rows outside the image or the band
// are computed like any other (their values are never consumed by a valid row).
template <int N, bool INT, bool GEN, bool TMA, int PH, bool SLOW>
__device__ __forceinline__ void f4_row(Fed4Regs<N>& R, const Fed4Args& a, const Fed4Lane& ln, int t)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int h = a.h;
    // ---- landing ring
    {
        if (TMA) __syncwarp();                 // every lane has read the slot that is requested again (row t - 1)
        f4_request<TMA>(a, ln, t + F4_RING - 1, (PH + F4_RING - 1) % F4_RING);
        f4_wait<TMA>(ln, PH);
        float4 v, g;
        const unsigned sa = ln.ring + PH * F4_ROWB;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "r"(sa + 512u) : "memory");
        R.L0[PH % 3][0] = v.x; R.L0[PH % 3][1] = v.y; R.L0[PH % 3][2] = v.z; R.L0[PH % 3][3] = v.w;
        R.G[PH % 2][0] = g.x; R.G[PH % 2][1] = g.y; R.G[PH % 2][2] = g.z; R.G[PH % 2][3] = g.w;
    }
    // neighbours of the centre rows, exchanged one row time ago (the shuffle latency is off the critical path)
    float nl[N], nr[N];
#pragma unroll
    for (int s = 0; s < N; s++) { nl[s] = R.NL[s]; nr[s] = R.NR[s]; }
    R.NL[0] = __shfl_up_sync(FULL, R.L0[PH % 3][3], 1);
    R.NR[0] = __shfl_down_sync(FULL, R.L0[PH % 3][0], 1);
    // ---- pair sums of row t
    {
        constexpr int c0 = PH % 6, gc = PH % 2, gp = (PH + 1) % 2;
        const float gl = __shfl_up_sync(FULL, R.G[gc][3], 1), gr = __shfl_down_sync(FULL, R.G[gc][0], 1);
        R.SH[c0][0] = f3_sum<INT>(R.G[gc][0], gl);
#pragma unroll
        for (int j = 1; j < 4; j++) R.SH[c0][j] = f3_sum<INT>(R.G[gc][j], R.G[gc][j - 1]);
        R.SH[c0][4] = f3_sum<INT>(gr, R.G[gc][3]);
#pragma unroll
        for (int c = 0; c < 4; c++) R.SV[c0][c] = f3_sum<INT>(R.G[gc][c], R.G[gp][c]);
        // image border in x: the missing neighbour is the mirror pixel, so its pair sum is the opposite one
        if (ln.bl) R.SH[c0][0] = R.SH[c0][1];
        if (!GEN) { if (ln.br) R.SH[c0][4] = R.SH[c0][3]; }
        else {
#pragma unroll
            for (int c = 0; c < 4; c++) if (c == ln.ir) R.SH[c0][c + 1] = R.SH[c0][c];
        }
    }
    // ---- the wavefront of steps
#pragma unroll
    for (int s = 1; s <= N; s++) {
        const int r = t - s;
        const float* up = (s == 1) ? R.L0[(PH + 1) % 3] : R.Lr[s >= 2 ? s - 2 : 0][(PH + 1) % 3];
        const float* ce = (s == 1) ? R.L0[(PH + 2) % 3] : R.Lr[s >= 2 ? s - 2 : 0][(PH + 2) % 3];
        const float* dn = (s == 1) ? R.L0[PH % 3] : R.Lr[s >= 2 ? s - 2 : 0][PH % 3];
        const float* sh = R.SH[(PH + 6 - s) % 6];
        const float* su = R.SV[(PH + 6 - s) % 6];              // between rows r-1 and r
        const float* sd = R.SV[(PH + 7 - s) % 6];              // between rows r and r+1
        float ll = nl[s - 1], rr = nr[s - 1];
        if (ln.bl) ll = ce[1];                                 // x = 0: the left neighbour is x = 1
        if (!GEN) { if (ln.br) rr = ce[2]; }                   // x = w-1: the right neighbour is x = w-2
        const float sf = a.stepfac[s - 1];
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float LL = (c == 0) ? ll : ce[c > 0 ? c - 1 : 0];
            float LR = (c == 3) ? rr : ce[c < 3 ? c + 1 : 3];
            float LU = up[c], LD = dn[c], sU = su[c], sD = sd[c];
            if (GEN) { if (c == ln.ir) LR = LL; }
            if (SLOW) {
                if (r == 0) { LU = LD; sU = sD; }              // y = 0
                if (r == h - 1) { LD = LU; sD = sU; }          // y = h-1
            }
            o[c] = f3_upd<INT>(ce[c], sh[c], LL, sh[c + 1], LR, sD, LD, sU, LU, sf);
        }
        if (s < N) {
#pragma unroll
            for (int c = 0; c < 4; c++) R.Lr[s - 1][PH % 3][c] = o[c];
            R.NL[s < N ? s : 0] = __shfl_up_sync(FULL, o[3], 1);
            R.NR[s < N ? s : 0] = __shfl_down_sync(FULL, o[0], 1);
        } else if (ln.store && (!SLOW || (r >= ln.y0 && r < ln.y1))) {
            float* d = ln.pdst + (long long)r * a.pitch;
            if (!GEN) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int c = 0; c < 4; c++) if (ln.ir < 0 || c <= ln.ir) d[c] = o[c];
            }
        }
    }
}

template <int N, bool INT, bool GEN, bool TMA, int PH>
__device__ __forceinline__ void f4_phase(Fed4Regs<N>& R, const Fed4Args& a, const Fed4Lane& ln, int it, int T)
{
    // every row time writes the same ring slots on both paths (no early outs: a skipped write would keep the old value of
    // the slot alive around the loop and cost its register for the whole iteration)
    const int i = it + PH, t = ln.t0 + i;
    if (i <= 2 * N || t >= a.h || i >= T) f4_row<N, INT, GEN, TMA, PH, true>(R, a, ln, t);
    else f4_row<N, INT, GEN, TMA, PH, false>(R, a, ln, t);
}

template <int N, bool INT, bool GEN, bool TMA>
__global__ void __launch_bounds__(32 * F4_WARPS, (N <= 3 ? 4 : 3)) k_fed4(const __grid_constant__ Fed4Args a)
{
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * F4_WARPS + (threadIdx.x >> 5);
    if (unit >= a.nunits) return;                       // whole warps leave: no block barrier is used
    // strips fastest: the warps of a CTA work on neighbouring strips of one band (their halo columns overlap in L2)
    const int per = a.nstrips * a.nbands;
    const int frame = unit / per, rem = unit - frame * per;
    const int band = rem / a.nstrips, strip = rem - band * a.nstrips;
    Fed4Lane ln;
    const int gx0 = strip * F4_COLS - 4 + 4 * lane;
    const int gxl = min(max(gx0, 0), a.pitch - 4);
    const long long base = (long long)frame * a.plane;
    ln.psrc = a.src + base + gxl; ln.pflow = a.flow + base + gxl; ln.pdst = a.dst + base + gx0;
    ln.y0 = band * a.band_h; ln.y1 = min(a.h, ln.y0 + a.band_h); ln.t0 = ln.y0 - N;
    ln.bl = gx0 == 0;
    ln.br = gx0 + 3 == a.w - 1;
    ln.ir = (a.w - 1) - gx0; if (ln.ir < 0 || ln.ir > 3) ln.ir = -1;
    ln.store = lane >= 1 && lane <= 30 && gx0 >= 0 && gx0 < a.w;
    if (!GEN) ln.ir = -1;

    __shared__ __align__(16) unsigned char ring_mem[F4_WARPS * F4_RING * F4_ROWB];
    ln.ring = (unsigned)__cvta_generic_to_shared(ring_mem + (threadIdx.x >> 5) * (F4_RING * F4_ROWB) + lane * 16);
    ln.wsrc = ln.wflow = nullptr; ln.wring = ln.mbar = ln.wbytes = ln.par = 0;
    if (TMA) {
        __shared__ __align__(8) unsigned long long mbars[F4_WARPS * F4_RING];
        const int gs = strip * F4_COLS - 4;                                // first column of the warp's 128-column segment
        const int c0 = max(gs, 0), c1 = min(gs + 128, a.pitch);            // part of it inside the row (multiples of 4 columns)
        ln.wsrc = a.src + base + c0; ln.wflow = a.flow + base + c0;
        ln.wbytes = (unsigned)(c1 - c0) * 4u;
        ln.wring = (unsigned)__cvta_generic_to_shared(ring_mem + (threadIdx.x >> 5) * (F4_RING * F4_ROWB)) + (unsigned)(c0 - gs) * 4u;
        ln.mbar = (unsigned)__cvta_generic_to_shared(mbars + (threadIdx.x >> 5) * F4_RING);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < F4_RING; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ln.mbar + k * 8) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
    }

    Fed4Regs<N> R;
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int c = 0; c < 4; c++) R.SV[i][c] = 0.f;
#pragma unroll
        for (int c = 0; c < 5; c++) R.SH[i][c] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) { R.L0[i][c] = 0.f; R.G[i & 1][c] = 0.f; }
#pragma unroll
    for (int s = 0; s < N; s++)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) R.Lr[s][i][c] = 0.f;
#pragma unroll
    for (int s = 0; s < N; s++) { R.NL[s] = 0.f; R.NR[s] = 0.f; }
    // rows t0 .. t0 + 4 are requested up front, one commit group per row
#pragma unroll
    for (int k = 0; k < F4_RING - 1; k++) f4_request<TMA>(a, ln, ln.t0 + k, k);
    const int T = (ln.y1 - ln.y0) + 2 * N;
    for (int it = 0; it < T; it += 6) {
        f4_phase<N, INT, GEN, TMA, 0>(R, a, ln, it, T);
        f4_phase<N, INT, GEN, TMA, 1>(R, a, ln, it, T);
        f4_phase<N, INT, GEN, TMA, 2>(R, a, ln, it, T);
        f4_phase<N, INT, GEN, TMA, 3>(R, a, ln, it, T);
        f4_phase<N, INT, GEN, TMA, 4>(R, a, ln, it, T);
        f4_phase<N, INT, GEN, TMA, 5>(R, a, ln, it, T);
        ln.par ^= 1u;
    }
}

akz_once_t g_attr_done;
int g_fed_rb = 4;                            // rows per thread block of k_fed3 (AKZ_FED_RB=2 selects the 4 x 2 variant)
int g_fed_band = 0;                          // rows per band of k_fed4 (AKZ_FED_BAND; 0 = chosen per launch)
int g_fed_stream = 1;                        // AKZ_FED_STREAM=0 forces the tile kernel (A/B measurements)
int g_fed_tma = 0;                           // AKZ_FED_TMA=1: landing ring filled by cp.async.bulk (TMA) + mbarrier instead of cp.async
int g_fed_maxn = 4;                          // steps per launch of k_fed4 (AKZ_FED_MAXN, 2..4; tuning knob)
int g_fed_min_units = 1024;                  // fewer (strip, band, frame) units than this: the tile kernel (AKZ_FED_MIN_UNITS)

}  // namespace

namespace akzk {

static void set_attrs()
{
    akz_once_guard once{g_attr_done};
    if (!once) return;
    if (const char* e = getenv("AKZ_FED_RB")) g_fed_rb = atoi(e) == 2 ? 2 : 4;
    if (const char* e = getenv("AKZ_FED_BAND")) g_fed_band = std::max(8, atoi(e));
    if (const char* e = getenv("AKZ_FED_MIN_UNITS")) g_fed_min_units = atoi(e);
    if (const char* e = getenv("AKZ_FED_MAXN")) g_fed_maxn = std::min(F4_MAXN, std::max(2, atoi(e)));
    if (const char* e = getenv("AKZ_FED_TMA")) g_fed_tma = atoi(e);
    if (const char* e = getenv("AKZ_FED_STREAM")) g_fed_stream = atoi(e);
    cudaFuncSetAttribute(k_fed3<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    cudaFuncSetAttribute(k_fed3<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
}

template <int N>
static void fed4_launch(cudaStream_t st, const Fed4Args& a, bool int_planes, bool gen)
{
    const int grid = (a.nunits + F4_WARPS - 1) / F4_WARPS;
    if (g_fed_tma && !int_planes && !gen) {                 // measured alternative: rows by TMA bulk copies + mbarriers
        k_fed4<N, false, false, true><<<grid, 32 * F4_WARPS, 0, st>>>(a);
        return;
    }
    if (int_planes) {
        if (gen) k_fed4<N, true, true, false><<<grid, 32 * F4_WARPS, 0, st>>>(a);
        else k_fed4<N, true, false, false><<<grid, 32 * F4_WARPS, 0, st>>>(a);
    } else {
        if (gen) k_fed4<N, false, true, false><<<grid, 32 * F4_WARPS, 0, st>>>(a);
        else k_fed4<N, false, false, false><<<grid, 32 * F4_WARPS, 0, st>>>(a);
    }
}

// dst receives the result of n steps applied to src; tmp is a scratch plane batch (never aliases src/dst).
// fused == 0: n single-step launches.  fused != 0: ceil(n / 4) launches of the streaming kernel k_fed4 when the rows are
// 16-byte aligned, else ceil(n / 8) launches of the tile kernel k_fed3; long cycles are split evenly and ping-pong through tmp.
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
    auto stepfac = [&](int k) {
        float f;
        if (int_planes) { int sfi = (int)(0.5f * tau[k] * 65536 + 0.5f); memcpy(&f, &sfi, 4); }      // akazed.cu:4240
        else f = 0.5f * tau[k];                                                                    // akazed.cu:2515
        return f;
    };
    const bool vec_ok = (pitch % 4 == 0) && (plane % 4 == 0) && (((uintptr_t)src | (uintptr_t)flowp | (uintptr_t)dst | (uintptr_t)tmp) % 16 == 0);
    // Streaming kernel: units of (strip of 120 columns) x (band of rows) x frame, one per warp.  The band height trades the
    // 2 N warm-up rows of every band against the number of units.  Measured on a B200 (scripts/fed_probe.py, 32 frames): bands
    // of ~96 rows are best as soon as there are ~1200 units (1920x1080 and 960x540: less than one round of resident warps is
    // fine, a warp's own row pipeline is what has to stay full); 480x270 wants bands of 32 rows (1152 units); 240x135 (320
    // units) is faster on the tile kernel, as are rows that are not 16-byte aligned.
    const int nstrips = (w + F4_COLS - 1) / F4_COLS;
    // band layout of a launch of N steps: (waves of CTAs) x (row times of a unit), bands of 32 rows or more: 1920x1080 x 32 frames,
    // N = 3 -> 9 bands of 120 rows = 1152 CTAs = two waves (0.152 ms per cycle; 11 bands of 99 rows = 2.4 waves: 0.160 ms); 480x270
    // -> bands of 30-32 rows.  Per LAUNCH, not per cycle: a 6-step cycle is two launches of 3 steps (four CTAs per SM, six warm-up
    // rows), not of 4 (960x540: 0.098 -> 0.089 ms)
    auto layout = [&](int N, int& band_h, int& nbands) {
        band_h = g_fed_band;
        if (band_h <= 0) {
            static const int nsm = [] { int d = 0, v = 148; cudaGetDevice(&d); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148; return v; }();
            const long long slots = (long long)nsm * (N <= 3 ? 4 : 3);
            long long best = -1;
            band_h = h;
            for (int nb = 1; nb <= std::max(1, h / 32); nb++) {
                const int bh = (h + nb - 1) / nb;
                const int nbd = (h + bh - 1) / bh;
                const long long ctas = ((long long)n * nstrips * nbd + F4_WARPS - 1) / F4_WARPS;
                const long long cost = ((ctas + slots - 1) / slots) * (bh + 2 * N);
                if (best < 0 || cost < best) { best = cost; band_h = bh; }
            }
            nbands = (h + band_h - 1) / band_h;
        } else {
            nbands = std::max(1, (h + band_h / 2) / band_h);
            band_h = (h + nbands - 1) / nbands;
            nbands = (h + band_h - 1) / band_h;
        }
    };
    int band_h, nbands;
    layout(std::min(nsteps, g_fed_maxn), band_h, nbands);
    const long long units = (long long)n * nstrips * nbands;
    const bool stream = vec_ok && g_fed_stream && w >= 8 && h >= 8 && units >= g_fed_min_units && units < (1ll << 30);
    const int maxk = stream ? g_fed_maxn : FE_MAXK;
    const int m = (nsteps + maxk - 1) / maxk;
    int done = 0;
    const float* cur = src;
    for (int i = 0; i < m; i++) {
        const int cnt = (nsteps - done + (m - i) - 1) / (m - i);          // balanced split
        float* out = ((m - 1 - i) % 2 == 0) ? dst : tmp;
        if (stream) {
            Fed4Args a4 = {};
            a4.src = cur; a4.flow = flowp; a4.dst = out; a4.plane = plane; a4.w = w; a4.h = h; a4.pitch = pitch;
            int bh, nb;
            layout(cnt, bh, nb);
            a4.nstrips = nstrips; a4.nbands = nb; a4.band_h = bh; a4.nunits = n * nstrips * nb;
            for (int k = 0; k < cnt; k++) a4.stepfac[k] = stepfac(done + k);
            const bool gen = (w % 4) != 0;
            switch (cnt) {
            case 1: fed4_launch<1>(st, a4, int_planes != 0, gen); break;
            case 2: fed4_launch<2>(st, a4, int_planes != 0, gen); break;
            case 3: fed4_launch<3>(st, a4, int_planes != 0, gen); break;
            default: fed4_launch<4>(st, a4, int_planes != 0, gen); break;
            }
        } else {
            FedArgs a = {};
            a.src = cur; a.flow = flowp; a.dst = out;
            a.plane = plane; a.w = w; a.h = h; a.pitch = pitch; a.n = cnt;
            for (int k = 0; k < cnt; k++) a.stepfac[k] = stepfac(done + k);
            const int rb = g_fed_rb == 2 ? 2 : 4;
            Fed2Geom ge = fed2_geom(cnt, rb);
            dim3 g((w + ge.two - 1) / ge.two, (h + ge.tho - 1) / ge.tho, n), b(F2_BX, F2_T / rb);
            Fed3Args a3;
            a3.f = a; a3.gx = g.x; a3.gy = g.y; a3.ntiles = (int)(g.x * g.y * g.z); a3.vec_ok = vec_ok ? 1 : 0;
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
        cur = out;
        done += cnt;
        launches++;
    }
    return launches;
}

}  // namespace akzk
