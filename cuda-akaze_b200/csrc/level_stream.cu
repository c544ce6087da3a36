// k_level4<S>: a whole level's "everything but the diffusion" in ONE streaming kernel -- sigma = 1 blur of the predecessor level,
// conductance, Lx, Ly and the Hessian determinant -- without the blurred plane's round trip through HBM that the pair
// k_blur4 (blur_stream.cu) + k_deriv4 (deriv_stream.cu) pays: 20 bytes per pixel instead of 28 (see STATUS below).  Subsumes gConv2d<2>
// (akazed.cu:204), gFlowNaive (:1068), gDerivate (:1267) and gHessianDeterminant (:1299) and their integer twins.
//
// The two halves want the rows in different orders: the 5-tap blur and the 3 x 3 conductance talk to NEIGHBOURING rows, the
// derivative stencils (tap distance S = sigma_size) only to rows of the same residue mod S.  So a CTA of S warps owns a strip of
// 128 columns (4 per lane) of a band of rows, warp k the rows of residue k, and the rows cross warps through shared memory:
//   iteration i, warp k, rows in blocks of S:  h = base + i S + k
//     1. input row h (cp.async landing ring, five iterations ahead) -> horizontal filter (neighbours by shuffle) -> ring HF
//     2. vertical filter of row a = h - 2S from the five HF rows a-2 .. a+2 (other warps', written before the last barrier)
//        -> blurred row a -> ring BL, and into this warp's register pipeline with its +-S column halo
//     3. Lx, Ly of row a - S and det of row a - 2S from the register pipeline, exactly as k_deriv4 (halos by shuffle)
//     4. conductance of row a - 2S: centre row from the register pipeline, rows a - 2S -+ 1 from ring BL
//     5. one block barrier
// Both rings hold six blocks of rows: a slot is rewritten three iterations after its last reader.  Every stage runs in every
// iteration, warm-up included: guarded by (warp-uniform) branches the stages no longer overlap and the kernel is 25 % slower.
//
// STATUS: opt-in (AKZ_LEVEL4=1; 2 = at any size, for the parity test).  Bit-exact in both pipelines, 20.4 bytes per pixel of HBM
// traffic instead of 27.2 -- and 5 % slower than the pair on a B200 (profiles/r02z_level4_ab.txt): the level is co-limited by
// issue slots, the fused form needs as many instructions as the pair plus wider halos (128 loaded columns per 112 stored on both
// halves), ~0.9 shared-memory wavefronts per pixel, and runs 12-15 warps per SM.  Kept as the measured answer to "why two kernels".
//
// Borders as in the two kernels it replaces: the blur runs on the mirrored input (bit-identical to reflect-101, the filter is
// symmetric), the conductance substitutes the mirror row / column of the BLURRED plane, and the derivative pixels whose taps
// leave the image are recomputed by the border-ring kernels of deriv_stream.cu -- for them the blurred plane IS written, but only
// within 8 pixels of the image border.
#include "common.cuh"
#include "kernels.h"
#include "level_math.cuh"
#include <algorithm>
#include <cstdlib>

using namespace akz;

namespace {

constexpr int L4_LAND = 6;                  // landing ring slots per warp
constexpr int L4_BLK = 6;                   // depth of the two row rings, in blocks of S rows
constexpr int L4_HFB = 512;                 // bytes of a row-filtered row (16 per lane)
constexpr int L4_BLB = 512 + 32;            // bytes of a blurred row: 16 bytes of padding at either end (halo reads of lanes 0 / 31)
constexpr int L4_LAG = 7;                   // iterations from a row's input to its determinant / conductance, plus the blur's warm-up

struct Level4Args {
    const float* src;
    float *flow, *smooth, *lx, *ly, *det;
    unsigned char* hot;                     // optional, see k_deriv4
    const float* kc;
    long long plane;
    LevelMathArgs m;
    float kscale, thr;
    int ithr, nmul, type;
    int w, h, pitch;
    int nstrips, nbands, band_h, nunits;
};

template <int S> struct L4 {
    static constexpr int HL = S <= 3 ? 2 : 3;            // halo lanes at either end: 2 + 2 S columns
    static constexpr int OC = (32 - 2 * HL) * 4;         // output columns of a strip
    static constexpr int RW = 4 + 2 * S;                 // a register row: own 4 values with S neighbours on either side
    static constexpr int ROWS = L4_BLK * S;              // rows per ring
    static constexpr int SMEM = S * L4_LAND * 512 + ROWS * (L4_HFB + L4_BLB);
    static constexpr int CTAS = S == 2 ? 8 : (S == 3 ? 5 : 3);
};

template <int S>
struct Level4Regs {
    float Sm[3][L4<S>::RW];                 // blurred rows a - 2S, a - S, a               (ring by iteration mod 3)
    float Lx[3][L4<S>::RW], Ly[3][L4<S>::RW];   // derivative rows a - 3S, a - 2S, a - S   (ring by iteration mod 3)
};

struct Level4Lane {
    const float* psrc;                      // frame base + clamped column of this lane
    long long obase, hbase;                 // frame base + column (outputs) / group (hot plane) of this lane
    unsigned land, hf, bl;                  // shared-memory addresses of this lane's 16 bytes: landing slot 0 of its warp, row 0 of ring HF, of ring BL
    int y0, y1, yb, k;                      // band rows [y0, y1), row of iteration 0 of warp 0, warp (= row residue)
    float ikc;
    bool lb, rb, store, edge;               // lane holds column 0 / w - 1; writes outputs; lies within 8 columns of the left / right border
};

__device__ __forceinline__ float l4_lds1(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float2 l4_lds2(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 l4_lds4(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void l4_sts4(unsigned a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ void l4_request(const Level4Args& a, const Level4Lane& ln, int row, int slot)
{
    const int r = min(max(refl(row, a.h), 0), a.h - 1);
    const unsigned d = ln.land + slot * 512;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(ln.psrc + (long long)r * a.pitch) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int S>
__device__ __forceinline__ int l4_wrap(int q) { return q < 0 ? q + L4<S>::ROWS : (q >= L4<S>::ROWS ? q - L4<S>::ROWS : q); }

// one iteration; PH = iteration mod 6
template <int S, bool INT, int TYPE, int PH>
__device__ __forceinline__ void l4_row(Level4Regs<S>& R, const Level4Args& a, const Level4Lane& ln, int i)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RW = L4<S>::RW;
    const int type = TYPE >= 0 ? TYPE : a.type;
    const int hrow = ln.yb + i * S + ln.k;
    // ---- 1. input row hrow -> horizontal filter -> ring HF
    l4_request(a, ln, hrow + (L4_LAND - 1) * S, (PH + L4_LAND - 1) % L4_LAND);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(L4_LAND - 1) : "memory");
    {
        const float4 p = l4_lds4(ln.land + PH * 512);
        float l2 = __shfl_up_sync(FULL, p.z, 1), l1 = __shfl_up_sync(FULL, p.w, 1);
        float r1 = __shfl_down_sync(FULL, p.x, 1), r2 = __shfl_down_sync(FULL, p.y, 1);
        if (ln.lb) { l1 = p.y; l2 = p.z; }                                 // x = -1, -2 -> 1, 2
        if (ln.rb) { r1 = p.z; r2 = p.y; }                                 // w, w + 1 -> w - 2, w - 3
        const float b0 = p2_gauss<INT>(l2, l1, p.x, p.y, p.z, a.m);
        const float b1 = p2_gauss<INT>(l1, p.x, p.y, p.z, p.w, a.m);
        const float b2 = p2_gauss<INT>(p.x, p.y, p.z, p.w, r1, a.m);
        const float b3 = p2_gauss<INT>(p.y, p.z, p.w, r1, r2, a.m);
        l4_sts4(ln.hf + (PH * S + ln.k) * L4_HFB, b0, b1, b2, b3);
    }
    // ---- 2. vertical filter -> blurred row arow (ring BL; register pipeline with its column halo)
    const int arow = hrow - 2 * S;
    const int qa = ((PH + 4) % 6) * S + ln.k;
    {
        const float4 m2 = l4_lds4(ln.hf + l4_wrap<S>(qa - 2) * L4_HFB);
        const float4 m1 = l4_lds4(ln.hf + l4_wrap<S>(qa - 1) * L4_HFB);
        const float4 c0 = l4_lds4(ln.hf + qa * L4_HFB);
        const float4 p1 = l4_lds4(ln.hf + l4_wrap<S>(qa + 1) * L4_HFB);
        const float4 p2 = l4_lds4(ln.hf + l4_wrap<S>(qa + 2) * L4_HFB);
        float* sm = R.Sm[PH % 3];
        sm[S] = p2_gauss<INT>(m2.x, m1.x, c0.x, p1.x, p2.x, a.m);
        sm[S + 1] = p2_gauss<INT>(m2.y, m1.y, c0.y, p1.y, p2.y, a.m);
        sm[S + 2] = p2_gauss<INT>(m2.z, m1.z, c0.z, p1.z, p2.z, a.m);
        sm[S + 3] = p2_gauss<INT>(m2.w, m1.w, c0.w, p1.w, p2.w, a.m);
        const unsigned sa = ln.bl + qa * L4_BLB;
        l4_sts4(sa, sm[S], sm[S + 1], sm[S + 2], sm[S + 3]);
        // the border-ring kernels read the blurred plane within 2 S <= 8 pixels of the image border
        if (ln.store && arow >= ln.y0 && arow < ln.y1 && (ln.edge || arow < 8 || arow >= a.h - 8))
            *reinterpret_cast<float4*>(a.smooth + ln.obase + (long long)arow * a.pitch) = make_float4(sm[S], sm[S + 1], sm[S + 2], sm[S + 3]);
        if (S == 4) {
            __syncwarp();
            const float4 l = l4_lds4(sa - 16), r = l4_lds4(sa + 16);
            sm[0] = l.x; sm[1] = l.y; sm[2] = l.z; sm[3] = l.w;
            sm[8 % RW] = r.x; sm[9 % RW] = r.y; sm[10 % RW] = r.z; sm[11 % RW] = r.w;
        } else {
            // a 32- or 64-bit shared load at a 16-byte lane stride costs four wavefronts, a shuffle one
#pragma unroll
            for (int j = 0; j < S; j++) {
                const float lo = __shfl_up_sync(FULL, sm[S + 4 - S + j], 1);      // columns x - S + j of the left neighbour lane
                const float hi = __shfl_down_sync(FULL, sm[S + j], 1);            // columns x + 4 + j of the right neighbour lane
                sm[j] = lo; sm[(S + 4 + j) % RW] = hi;
            }
        }
    }
    // ---- 3. first derivatives of row arow - S, determinant of row arow - 2S (k_deriv4's row step)
    const int r1 = arow - S;
    {
        const float* up = R.Sm[(PH + 1) % 3];
        const float* ce = R.Sm[(PH + 2) % 3];
        const float* dn = R.Sm[PH % 3];
        float vx[4], vy[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int m = c + S;
            p2_deriv1<INT>(up[m - S], up[m], up[m + S], ce[m - S], ce[m + S], dn[m - S], dn[m], dn[m + S], a.m, vx[c], vy[c]);
        }
        if (ln.store && r1 >= ln.y0 && r1 < ln.y1) {
            const long long o = ln.obase + (long long)r1 * a.pitch;
            *reinterpret_cast<float4*>(a.lx + o) = make_float4(vx[0], vx[1], vx[2], vx[3]);
            *reinterpret_cast<float4*>(a.ly + o) = make_float4(vy[0], vy[1], vy[2], vy[3]);
        }
        float* lx = R.Lx[PH % 3];
        float* ly = R.Ly[PH % 3];
#pragma unroll
        for (int c = 0; c < 4; c++) { lx[S + c] = vx[c]; ly[S + c] = vy[c]; }
#pragma unroll
        for (int j = 0; j < S; j++) {
            lx[j] = __shfl_up_sync(FULL, vx[4 - S + j], 1);
            ly[j] = __shfl_up_sync(FULL, vy[4 - S + j], 1);
            lx[(S + 4 + j) % RW] = __shfl_down_sync(FULL, vx[j], 1);
            ly[(S + 4 + j) % RW] = __shfl_down_sync(FULL, vy[j], 1);
        }
    }
    // (every stage runs in every iteration, also in the warm-up of a band: straight-line code lets the stages overlap; with
    // branches around them the kernel was 25 % slower)
    const int r2 = arow - 2 * S;
    const bool out2 = ln.store && r2 >= ln.y0 && r2 < ln.y1;
    {
        const float* xu = R.Lx[(PH + 1) % 3];
        const float* xc = R.Lx[(PH + 2) % 3];
        const float* xl = R.Lx[PH % 3];
        const float* yu = R.Ly[(PH + 1) % 3];
        const float* yl = R.Ly[PH % 3];
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int m = c + S;
            o[c] = p2_det<INT>(xu[m - S], xu[m], xu[m + S], xc[m - S], xc[m + S], xl[m - S], xl[m], xl[m + S],
                               yu[m - S], yu[m], yu[m + S], yl[m - S], yl[m], yl[m + S], a.m);
        }
        if (out2) {
            *reinterpret_cast<float4*>(a.det + ln.obase + (long long)r2 * a.pitch) = make_float4(o[0], o[1], o[2], o[3]);
            if (a.hot) {
                const bool h = INT ? (fi(o[0]) > a.ithr || fi(o[1]) > a.ithr || fi(o[2]) > a.ithr || fi(o[3]) > a.ithr)
                                   : (o[0] > a.thr || o[1] > a.thr || o[2] > a.thr || o[3] > a.thr);
                a.hot[ln.hbase + (long long)r2 * (a.pitch >> 2)] = h ? 1 : 0;
            }
        }
    }
    // ---- 4. conductance of row r2: centre row from the register pipeline, rows r2 -+ 1 from ring BL (their left / right
    // neighbours by shuffle: a 32-bit shared load at a 16-byte lane stride costs four wavefronts)
    {
        const int qf = ((PH + 2) % 6) * S + ln.k;
        const float4 u4 = l4_lds4(ln.bl + l4_wrap<S>(qf - 1) * L4_BLB), d4 = l4_lds4(ln.bl + l4_wrap<S>(qf + 1) * L4_BLB);
        float u[6] = { __shfl_up_sync(FULL, u4.w, 1), u4.x, u4.y, u4.z, u4.w, __shfl_down_sync(FULL, u4.x, 1) };
        float d[6] = { __shfl_up_sync(FULL, d4.w, 1), d4.x, d4.y, d4.z, d4.w, __shfl_down_sync(FULL, d4.x, 1) };
        if (out2) {
            const float* ce = R.Sm[(PH + 1) % 3];
            float cl = ce[S - 1], cr = ce[S + 4];
            if (ln.lb) { u[0] = u[2]; d[0] = d[2]; cl = ce[S + 1]; }     // blurred(-1) := blurred(1)   (index reflection, akazed.cu:1076-1083)
            if (ln.rb) { u[5] = u[3]; d[5] = d[3]; cr = ce[S + 2]; }     // blurred(w)  := blurred(w - 2)
            if (r2 == 0) {                                                 // row -1 := row 1
#pragma unroll
                for (int c = 0; c < 6; c++) u[c] = d[c];
            }
            if (r2 == a.h - 1) {                                           // row h := row h - 2
#pragma unroll
                for (int c = 0; c < 6; c++) d[c] = u[c];
            }
            float f[4];
            f[0] = p2_flow<INT>(u[0], u[1], u[2], cl, ce[S + 1], d[0], d[1], d[2], type, ln.ikc);
            f[1] = p2_flow<INT>(u[1], u[2], u[3], ce[S], ce[S + 2], d[1], d[2], d[3], type, ln.ikc);
            f[2] = p2_flow<INT>(u[2], u[3], u[4], ce[S + 1], ce[S + 3], d[2], d[3], d[4], type, ln.ikc);
            f[3] = p2_flow<INT>(u[3], u[4], u[5], ce[S + 2], cr, d[3], d[4], d[5], type, ln.ikc);
            *reinterpret_cast<float4*>(a.flow + ln.obase + (long long)r2 * a.pitch) = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
    // ---- 5. rows written in this iteration become visible to the other warps; slots read in it may be rewritten
    __syncthreads();
}

template <int S, bool INT, int TYPE, int CTAS>
__global__ void __launch_bounds__(32 * S, CTAS) k_level4(const __grid_constant__ Level4Args a)
{
    const int lane = threadIdx.x & 31;
    // strips fastest, then bands, then frames
    int rem = blockIdx.x;
    const int strip = rem % a.nstrips; rem /= a.nstrips;
    const int band = rem % a.nbands;
    const int frame = rem / a.nbands;
    Level4Lane ln;
    ln.k = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);          // provably warp-uniform: the branches on row numbers stay uniform branches
    const int gx0 = strip * L4<S>::OC - 4 * L4<S>::HL + 4 * lane;
    const int gxl = min(max(gx0, 0), a.pitch - 4);
    const long long base = (long long)frame * a.plane;
    ln.psrc = a.src + base + gxl;
    ln.obase = base + gx0;
    ln.hbase = (long long)frame * (a.pitch >> 2) * a.h + (gx0 >> 2);
    ln.y0 = band * a.band_h; ln.y1 = min(a.h, ln.y0 + a.band_h);
    ln.yb = ln.y0 - 3 * S;
    ln.lb = gx0 == 0;
    ln.rb = gx0 + 3 == a.w - 1;
    ln.store = lane >= L4<S>::HL && lane < 32 - L4<S>::HL && gx0 >= 0 && gx0 < a.w;
    ln.edge = gx0 < 8 || gx0 + 4 > a.w - 8;
    if (!INT) {
        float k = a.kc[frame];
        for (int i = 0; i < a.nmul; i++) k = __fmul_rn(k, a.kscale);
        ln.ikc = __fdiv_rn(1.f, __fmul_rn(k, k));
    } else {
        int k = reinterpret_cast<const int*>(a.kc)[frame];
        for (int i = 0; i < a.nmul; i++) k = (int)__fadd_rn(__fmul_rn((float)k, 0.75f), 0.5f);       // akaze.cpp:649
        ln.ikc = __fdiv_rn(1.f, (float)(k * k));                                                     // akazed.cu:4218 (host)
    }
    __shared__ __align__(16) unsigned char smem[L4<S>::SMEM];
    ln.land = (unsigned)__cvta_generic_to_shared(smem + ln.k * (L4_LAND * 512) + lane * 16);
    ln.hf = (unsigned)__cvta_generic_to_shared(smem + S * L4_LAND * 512 + lane * 16);
    ln.bl = (unsigned)__cvta_generic_to_shared(smem + S * L4_LAND * 512 + L4<S>::ROWS * L4_HFB + 16 + lane * 16);

    Level4Regs<S> R;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < L4<S>::RW; c++) { R.Sm[i][c] = 0.f; R.Lx[i][c] = 0.f; R.Ly[i][c] = 0.f; }
#pragma unroll
    for (int q = 0; q < L4_LAND - 1; q++) l4_request(a, ln, ln.yb + q * S + ln.k, q);
    // the determinant / conductance of the band's first row of a residue leave in iteration L4_LAG
    const int T = (ln.y1 - ln.y0 + S - 1) / S + L4_LAG;
    for (int i = 0; i < T; i += 6) {                                      // T is the same for every warp of the CTA (block barrier inside)
        l4_row<S, INT, TYPE, 0>(R, a, ln, i);
        if (i + 1 >= T) break;
        l4_row<S, INT, TYPE, 1>(R, a, ln, i + 1);
        if (i + 2 >= T) break;
        l4_row<S, INT, TYPE, 2>(R, a, ln, i + 2);
        if (i + 3 >= T) break;
        l4_row<S, INT, TYPE, 3>(R, a, ln, i + 3);
        if (i + 4 >= T) break;
        l4_row<S, INT, TYPE, 4>(R, a, ln, i + 4);
        if (i + 5 >= T) break;
        l4_row<S, INT, TYPE, 5>(R, a, ln, i + 5);
    }
}

template <int S, bool INT>
int level4_launch(cudaStream_t st, Level4Args& a, int n, int force)
{
    a.nstrips = (a.w + L4<S>::OC - 1) / L4<S>::OC;
    // bands: multiples of 12 rows (every band starts on residue 0 of every S), L4_LAG iterations of warm-up each; their number by
    // the model of k_deriv4: (waves of CTAs the GPU needs) x (iterations of a CTA)
    static const int nsm = [] { int d = 0, v = 148; cudaGetDevice(&d); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148; return v; }();
    const long long per_band = (long long)n * a.nstrips;
    // S = 3: five CTAs per SM mean 128 registers and a few spilled values, four mean 162 registers and were 9 % faster (AKZ_L4_CTAS=5 selects five)
    static const int ctas3 = [] { const char* e = getenv("AKZ_L4_CTAS"); return e && atoi(e) == 5 ? 5 : 4; }();
    const int ctas = S == 3 ? ctas3 : L4<S>::CTAS;
    const long long slots = (long long)nsm * ctas;
    const int nb_lo = std::max(1, (a.h + 287) / 288), nb_hi = std::max(nb_lo, (a.h + 47) / 48);
    long long best_cost = -1;
    int best_bh = a.h;
    for (int nb = nb_lo; nb <= nb_hi; nb++) {
        const int bh = std::max(48, ((a.h + nb - 1) / nb + 11) / 12 * 12);
        const int nbands = (a.h + bh - 1) / bh;
        const long long ctas = per_band * nbands;
        const long long cost = ((ctas + slots - 1) / slots) * (bh / S + L4_LAG);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bh = bh; }
    }
    a.band_h = best_bh;
    a.nbands = (a.h + a.band_h - 1) / a.band_h;
    const long long units = per_band * a.nbands;
    if (units >= (1ll << 30)) return 0;
    // few CTAs (single frames, small levels): the serial march of a CTA over its band is the critical path; the pair of kernels
    // (or the tile kernels behind them) is faster there
    if (!force && units < slots) return 0;
    a.nunits = (int)units;
    if (S == 3 && ctas == 4) {
        if (a.type == 1) k_level4<S, INT, 1, (S == 3 ? 4 : L4<S>::CTAS)><<<a.nunits, 32 * S, 0, st>>>(a);
        else k_level4<S, INT, -1, (S == 3 ? 4 : L4<S>::CTAS)><<<a.nunits, 32 * S, 0, st>>>(a);
    } else {
        if (a.type == 1) k_level4<S, INT, 1, L4<S>::CTAS><<<a.nunits, 32 * S, 0, st>>>(a);
        else k_level4<S, INT, -1, L4<S>::CTAS><<<a.nunits, 32 * S, 0, st>>>(a);
    }
    return 1;
}

}  // namespace

namespace akzk {

// One level from its predecessor (same resolution): conductance, Lx, Ly, det (+ the hot plane), border rings included.
// `smooth` receives the blurred plane only near the image border (what the ring kernels read).  Returns the number of launches,
// 0 when the case is not covered (the caller then runs blur_stream + deriv_stream).  force: skip the size heuristics (tests).
int level_stream(cudaStream_t st, const float* src, float* flowp, float* smooth, float* lx, float* ly, float* det, int type,
                 const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n, int int_planes,
                 cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join, unsigned char* hot, float thr, int ithr, int force)
{
    if (step < 2 || step > 4 || w < 64 || h < 48 || (w % 4) != 0 || (pitch % 4) != 0 || (plane % 4) != 0) return 0;
    if (!src || !flowp || !smooth || !lx || !ly || !det) return 0;
    if ((((uintptr_t)src | (uintptr_t)flowp | (uintptr_t)smooth | (uintptr_t)lx | (uintptr_t)ly | (uintptr_t)det) % 16) != 0) return 0;
    if (src == smooth || src == flowp || src == lx || src == ly || src == det) return 0;
    Level4Args a = {};
    a.src = src; a.flow = flowp; a.smooth = smooth; a.lx = lx; a.ly = ly; a.det = det; a.hot = hot; a.kc = kc;
    a.plane = plane; a.kscale = kscale; a.thr = thr; a.ithr = ithr; a.nmul = nmul; a.type = type;
    a.w = w; a.h = h; a.pitch = pitch;
    float k[3];
    akz_gauss_taps(1.f, 2, k);
    a.m.k0 = k[0]; a.m.k1 = k[1]; a.m.k2 = k[2];
    a.m.ik0 = (int)(k[0] * 65536 + 0.5f); a.m.ik1 = (int)(k[1] * 65536 + 0.5f); a.m.ik2 = (int)(k[2] * 65536 + 0.5f);      // akazed.cu:3896
    hessian_factors(&a.m.fac1, &a.m.fac2);
    a.m.ifac1 = (int)(a.m.fac1 * 65536 + 0.5f); a.m.ifac2 = (int)(a.m.fac2 * 65536 + 0.5f);                              // akazed.cu:4184-4185
    int r;
    if (int_planes) r = step == 2 ? level4_launch<2, true>(st, a, n, force) : step == 3 ? level4_launch<3, true>(st, a, n, force) : level4_launch<4, true>(st, a, n, force);
    else r = step == 2 ? level4_launch<2, false>(st, a, n, force) : step == 3 ? level4_launch<3, false>(st, a, n, force) : level4_launch<4, false>(st, a, n, force);
    if (r <= 0) return r;
    return r + deriv_rings(st, smooth, lx, ly, det, step, w, h, pitch, plane, n, int_planes, ring_st, ev_fork, ev_join);
}

}  // namespace akzk
