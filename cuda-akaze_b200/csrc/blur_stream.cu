// k_blur4<MODE>: the blur half of a level as a streaming warp kernel -- sigma = 1 Gaussian blur of the predecessor level (or
// the octave transition: point subsample + blur on the coarse lattice), the conductance plane g, and the blurred plane itself
// for the streaming derivative kernel (deriv_stream.cu).  Subsumes gConv2d<2> | gDownWithSmooth (akazed.cu:204, :449; integer
// twins) and gFlowNaive (:1068).  Same structure as k_fed4 (fed.cu): a WARP owns a strip of 128 columns (4 per lane, one halo
// lane at either end: the blur reaches 2 columns, the Scharr conductance 1) and marches down a band of rows.  At row time t it
// takes row t of the input (cp.async landing ring, five rows ahead), filters it horizontally with the neighbours' values by
// shuffle, filters vertically over the five row-filtered rows held in registers (blurred row t-2, stored), and evaluates the
// conductance of row t-3 from the three blurred rows in registers (their left / right neighbours were exchanged when each row
// was produced).  No shared-memory tile, no block barrier: ~36 FP instructions of pinned arithmetic per pixel plus ~9 of data
// movement, against a tile kernel (k_prep3) that ran at 3.0 TB/s of algorithmic bytes, bound by issue slots and shared memory.
//
// Borders.  The blur is symmetric and fadd commutes, so evaluating it on the mirrored input IS the reflect-101 (BLUR) /
// source-coordinate reflection (DOWN, akazed.cu:474-494) of the reference, bit for bit: rows arrive by reflected index, the
// lanes holding column 0 / w-1 take their missing neighbours from their own registers.  The conductance reflects indices of the
// BLURRED plane: its missing neighbour column / row is the mirror one (substituted at use).
#include "common.cuh"
#include "kernels.h"
#include "level_math.cuh"
#include <algorithm>
#include <cstdlib>

using namespace akz;

namespace {

constexpr int B4_WARPS = 4;
constexpr int B4_COLS = 120;                // output columns of a strip: lanes 1..30
constexpr int B4_RING = 6;                  // landing ring slots per warp
enum { BM_BLUR = 1, BM_DOWN = 2 };

struct Blur4Args {
    const float* src;                       // predecessor Lt (BLUR) or the finer octave's level 0 (DOWN: source is 2 w x 2 h)
    float* ltdst;                           // DOWN: subsampled Lt
    float *flow, *smooth;
    const float* kc;
    long long splane, plane;
    LevelMathArgs m;
    float kscale;
    int nmul, type;
    int sw, sh, sp;
    int w, h, pitch;
    int nstrips, nbands, band_h, nunits;
};

template <int MODE> struct B4 {
    static constexpr int LANEB = MODE == BM_DOWN ? 32 : 16;          // bytes of input per lane and row
    static constexpr int SLOTB = 32 * LANEB;
};

struct Blur4Regs {
    float Bf[6][4];                         // row-filtered rows t-4 .. t     (ring by row time mod 6)
    float Sm[3][6];                         // blurred rows t-4, t-3, t-2 with their left / right neighbour (ring by production time mod 3)
};

struct Blur4Lane {
    const float* psrc;                      // frame base + clamped (source) column of this lane
    long long obase;                        // frame base + column of this lane (outputs)
    unsigned ring;
    int y0, y1, t0;
    float ikc;
    bool bl, br, store;
};

template <int MODE>
__device__ __forceinline__ void b4_request(const Blur4Args& a, const Blur4Lane& ln, int row_time, int slot)
{
    int r;
    if (MODE == BM_DOWN) {                                  // source row of coarse row t, reflected in source coordinates
        int c = 2 * row_time;
        if (c < 0) c = -c;
        if (c >= a.sh) c = a.sh + a.sh - 2 - c;
        r = min(max(c, 0), a.sh - 1);
    } else {
        r = min(max(refl(row_time, a.h), 0), a.h - 1);
    }
    const float* g = ln.psrc + (long long)r * a.sp;
    const unsigned d = ln.ring + slot * B4<MODE>::SLOTB;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g) : "memory");
    if (MODE == BM_DOWN) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + 16u), "l"(g + 4) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int MODE, bool INT, int TYPE, int PH>
__device__ __forceinline__ void b4_row(Blur4Regs& R, const Blur4Args& a, const Blur4Lane& ln, int t)
{
    const int type = TYPE >= 0 ? TYPE : a.type;       // TYPE >= 0: diffusivity fixed at compile time (one conductance formula in the loop)
    constexpr unsigned FULL = 0xffffffffu;
    // ---- input row t
    b4_request<MODE>(a, ln, t + B4_RING - 1, (PH + B4_RING - 1) % B4_RING);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(B4_RING - 1) : "memory");
    float v[4];
    {
        const unsigned sa = ln.ring + PH * B4<MODE>::SLOTB;
        float4 p, q;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w) : "r"(sa) : "memory");
        if (MODE == BM_DOWN) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(sa + 16u) : "memory");
            v[0] = p.x; v[1] = p.z; v[2] = q.x; v[3] = q.z;              // dst(x, y) = src(2x, 2y)   (akazed.cu:505)
        } else {
            v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
        }
    }
    if (MODE == BM_DOWN && ln.store && t >= ln.y0 && t < ln.y1)
        *reinterpret_cast<float4*>(a.ltdst + ln.obase + (long long)t * a.pitch) = make_float4(v[0], v[1], v[2], v[3]);
    // ---- horizontal pass -> Bf[t]
    {
        float l2 = __shfl_up_sync(FULL, v[2], 1), l1 = __shfl_up_sync(FULL, v[3], 1);
        float r1 = __shfl_down_sync(FULL, v[0], 1), r2 = __shfl_down_sync(FULL, v[1], 1);
        if (ln.bl) { l1 = v[1]; l2 = v[2]; }                               // x = -1, -2 -> 1, 2
        if (ln.br) {
            if (MODE == BM_DOWN) { r1 = v[3]; r2 = v[2]; }                 // coarse w, w+1 -> source 2w-2, 2w-4 = coarse w-1, w-2
            else { r1 = v[2]; r2 = v[1]; }                                 // w, w+1 -> w-2, w-3
        }
        float* b = R.Bf[PH % 6];
        b[0] = p2_gauss<INT>(l2, l1, v[0], v[1], v[2], a.m);
        b[1] = p2_gauss<INT>(l1, v[0], v[1], v[2], v[3], a.m);
        b[2] = p2_gauss<INT>(v[0], v[1], v[2], v[3], r1, a.m);
        b[3] = p2_gauss<INT>(v[1], v[2], v[3], r1, r2, a.m);
    }
    // ---- vertical pass -> blurred row t-2 (stored, and kept with its left / right neighbours)
    const int rs = t - 2;
    {
        const float* m2 = R.Bf[(PH + 2) % 6];
        const float* m1 = R.Bf[(PH + 3) % 6];
        const float* c0 = R.Bf[(PH + 4) % 6];
        const float* p1 = R.Bf[(PH + 5) % 6];
        const float* p2 = R.Bf[PH % 6];
        float* s = R.Sm[PH % 3];
#pragma unroll
        for (int c = 0; c < 4; c++) s[1 + c] = p2_gauss<INT>(m2[c], m1[c], c0[c], p1[c], p2[c], a.m);
        if (ln.store && rs >= ln.y0 && rs < ln.y1)
            *reinterpret_cast<float4*>(a.smooth + ln.obase + (long long)rs * a.pitch) = make_float4(s[1], s[2], s[3], s[4]);
        s[0] = __shfl_up_sync(FULL, s[4], 1);
        s[5] = __shfl_down_sync(FULL, s[1], 1);
        if (ln.bl) s[0] = s[2];                                            // blurred(-1) := blurred(1)   (index reflection, akazed.cu:1076-1083)
        if (ln.br) s[5] = s[3];                                            // blurred(w)  := blurred(w-2)
    }
    // ---- conductance of row t-3 from the blurred rows t-4, t-3, t-2
    const int rg = t - 3;
    if (ln.store && rg >= ln.y0 && rg < ln.y1) {
        const float* up = R.Sm[(PH + 1) % 3];
        const float* ce = R.Sm[(PH + 2) % 3];
        const float* dn = R.Sm[PH % 3];
        float o[4];
        if (rg == 0 || rg == a.h - 1) {                                    // warp-uniform: first / last image row
            float u[6], d[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                u[c] = rg == 0 ? dn[c] : up[c];                            // row -1 := row 1
                d[c] = rg == a.h - 1 ? up[c] : dn[c];                      // row h  := row h-2
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                o[c] = p2_flow<INT>(u[c], u[c + 1], u[c + 2], ce[c], ce[c + 2], d[c], d[c + 1], d[c + 2], type, ln.ikc);
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                o[c] = p2_flow<INT>(up[c], up[c + 1], up[c + 2], ce[c], ce[c + 2], dn[c], dn[c + 1], dn[c + 2], type, ln.ikc);
        }
        *reinterpret_cast<float4*>(a.flow + ln.obase + (long long)rg * a.pitch) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

template <int MODE, bool INT, int TYPE>
__global__ void __launch_bounds__(32 * B4_WARPS, 4) k_blur4(const __grid_constant__ Blur4Args a)
{
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * B4_WARPS + (threadIdx.x >> 5);
    if (unit >= a.nunits) return;
    const int per = a.nstrips * a.nbands;
    const int frame = unit / per, rem = unit - frame * per;
    const int band = rem / a.nstrips, strip = rem - band * a.nstrips;
    Blur4Lane ln;
    const int gx0 = strip * B4_COLS - 4 + 4 * lane;
    // clamped column of the loads (halo lanes beyond the image read valid memory; their values are never used)
    const int gxl = min(max(gx0, 0), (MODE == BM_DOWN ? min(a.pitch, a.sp / 2) : a.pitch) - 4);
    ln.psrc = a.src + (long long)frame * a.splane + (MODE == BM_DOWN ? 2 * gxl : gxl);
    ln.obase = (long long)frame * a.plane + gx0;
    ln.y0 = band * a.band_h; ln.y1 = min(a.h, ln.y0 + a.band_h); ln.t0 = ln.y0 - 3;
    ln.bl = gx0 == 0;
    ln.br = gx0 + 3 == a.w - 1;
    ln.store = lane >= 1 && lane <= 30 && gx0 >= 0 && gx0 < a.w;
    if (!INT) {
        float k = a.kc[frame];
        for (int i = 0; i < a.nmul; i++) k = __fmul_rn(k, a.kscale);
        ln.ikc = __fdiv_rn(1.f, __fmul_rn(k, k));
    } else {
        int k = reinterpret_cast<const int*>(a.kc)[frame];
        for (int i = 0; i < a.nmul; i++) k = (int)__fadd_rn(__fmul_rn((float)k, 0.75f), 0.5f);       // akaze.cpp:649
        ln.ikc = __fdiv_rn(1.f, (float)(k * k));                                                     // akazed.cu:4218 (host)
    }
    __shared__ __align__(16) unsigned char ring_mem[B4_WARPS * B4_RING * B4<MODE>::SLOTB];
    ln.ring = (unsigned)__cvta_generic_to_shared(ring_mem + (threadIdx.x >> 5) * (B4_RING * B4<MODE>::SLOTB) + lane * B4<MODE>::LANEB);

    Blur4Regs R;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) R.Bf[i][c] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < 6; c++) R.Sm[i][c] = 0.f;
#pragma unroll
    for (int k = 0; k < B4_RING - 1; k++) b4_request<MODE>(a, ln, ln.t0 + k, k);
    // blurred row y0-1 needs input rows from y0-3; the conductance of row y1-1 is complete at row time y1+2
    const int T = (ln.y1 - ln.y0) + 6;
    for (int it = 0; it < T; it += 6) {
        b4_row<MODE, INT, TYPE, 0>(R, a, ln, ln.t0 + it);
        b4_row<MODE, INT, TYPE, 1>(R, a, ln, ln.t0 + it + 1);
        b4_row<MODE, INT, TYPE, 2>(R, a, ln, ln.t0 + it + 2);
        b4_row<MODE, INT, TYPE, 3>(R, a, ln, ln.t0 + it + 3);
        b4_row<MODE, INT, TYPE, 4>(R, a, ln, ln.t0 + it + 4);
        b4_row<MODE, INT, TYPE, 5>(R, a, ln, ln.t0 + it + 5);
    }
}

}  // namespace

namespace akzk {

// Streaming blur half of a level: mode 1 = same-resolution blur, 2 = octave transition (source exactly twice the size).
// Returns 1 when launched, 0 when the case is not covered (the tile kernel k_prep3 then takes the level).
int blur_stream(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                float* ltdst, float* flowp, float* smooth, int type, const float* kc, float kscale, int nmul,
                int w, int h, int pitch, long long plane, int n, int int_planes)
{
    if ((mode != 1 && mode != 2) || w < 32 || h < 16 || !flowp || !smooth || (w % 4) != 0) return 0;
    if ((pitch % 4) != 0 || (plane % 4) != 0 || (sp % 4) != 0 || (splane % 4) != 0) return 0;
    if ((((uintptr_t)src | (uintptr_t)flowp | (uintptr_t)smooth | (uintptr_t)ltdst) % 16) != 0) return 0;
    if (mode == 2 && (sw != 2 * w || sh != 2 * h || !ltdst || (sp % 8) != 0)) return 0;
    if (mode == 1 && (sw != w || sh != h)) return 0;
    if (src == smooth || src == flowp) return 0;
    Blur4Args a = {};
    a.src = src; a.ltdst = ltdst; a.flow = flowp; a.smooth = smooth; a.kc = kc;
    a.splane = splane; a.plane = plane; a.kscale = kscale; a.nmul = nmul; a.type = type;
    a.sw = sw; a.sh = sh; a.sp = sp; a.w = w; a.h = h; a.pitch = pitch;
    float k[3];
    akz_gauss_taps(1.f, 2, k);
    a.m.k0 = k[0]; a.m.k1 = k[1]; a.m.k2 = k[2];
    a.m.ik0 = (int)(k[0] * 65536 + 0.5f); a.m.ik1 = (int)(k[1] * 65536 + 0.5f); a.m.ik2 = (int)(k[2] * 65536 + 0.5f);      // akazed.cu:3896
    a.nstrips = (w + B4_COLS - 1) / B4_COLS;
    // Bands cost six rows of warm-up each.  Their number by the model of k_deriv4 / k_fed4: (waves of CTAs the GPU needs) x (row
    // times of a unit); AKZ_BLUR_BAND=<rows> overrides (tuning knob)
    {
        static const int nsm = [] { int d = 0, v = 148; cudaGetDevice(&d); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148; return v; }();
        static const int forced = [] { const char* e = getenv("AKZ_BLUR_BAND"); return e ? atoi(e) : 0; }();
        const long long per_band = (long long)n * a.nstrips, slots = (long long)nsm * 4;       // resident CTAs (launch bounds of k_blur4)
        const int nb_lo = std::max(1, (h + 359) / 360), nb_hi = std::max(nb_lo, (h + 31) / 32);
        long long best_cost = -1;
        int best_bh = h;
        for (int nb = nb_lo; nb <= nb_hi; nb++) {
            const int bh = std::max(32, (h + nb - 1) / nb);
            const int nbands = (h + bh - 1) / bh;
            const long long ctas = (per_band * nbands + B4_WARPS - 1) / B4_WARPS;
            const long long cost = ((ctas + slots - 1) / slots) * (bh + 6);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bh = bh; }
        }
        if (forced >= 16) best_bh = std::min(h, forced);
        a.band_h = best_bh; a.nbands = (h + best_bh - 1) / best_bh;
    }
    const long long units = (long long)n * a.nstrips * a.nbands;
    if (units >= (1ll << 30)) return 0;
    a.nunits = (int)units;
    // few units (small batches, small levels): a warp's serial march over its band is the critical path; the tile kernel is faster
    static const int min_units = [] { const char* e = getenv("AKZ_BLUR_MIN_UNITS"); return e ? atoi(e) : 1024; }();
    if (a.nunits < min_units) return 0;
    const int grid = (a.nunits + B4_WARPS - 1) / B4_WARPS;
    // the reference default PM_G2 (akaze.h:53) has the diffusivity as a compile-time constant; the others share a generic instance
#define AKZ_B4(M, I) do { if (type == 1) k_blur4<M, I, 1><<<grid, 32 * B4_WARPS, 0, st>>>(a); else k_blur4<M, I, -1><<<grid, 32 * B4_WARPS, 0, st>>>(a); } while (0)
    if (int_planes) { if (mode == 1) AKZ_B4(BM_BLUR, true); else AKZ_B4(BM_DOWN, true); }
    else { if (mode == 1) AKZ_B4(BM_BLUR, false); else AKZ_B4(BM_DOWN, false); }
#undef AKZ_B4
    return 1;
}

}  // namespace akzk
