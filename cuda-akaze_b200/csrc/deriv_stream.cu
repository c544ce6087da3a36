// k_deriv4<S>: first derivatives Lx, Ly and the Hessian determinant of a level from its smoothed plane, as a streaming warp
// kernel.  Subsumes gDerivate (akazed.cu:1267-1296) and gHessianDeterminant (:1299-1331; integer twins :3339-3403).
//
// Both stencils are 3 x 3 with tap distance S (the level's sigma_size, 2..4): Lx, Ly(x, y) read the smoothed plane at
// (x +- S, y +- S), det(x, y) reads Lx, Ly at (x +- S, y +- S).  Rows therefore only ever talk to rows of the same residue
// mod S: the rows of one residue class form an ordinary stride-1 pipeline.  A WARP owns a strip of 128 columns (4 per lane)
// and the rows of ONE residue class of a band: at row time j it takes row R_j = base + j S of the smoothed plane (16 bytes per
// lane through a cp.async landing ring in shared memory, four rows ahead), reads it back with its +-S column halo (the
// neighbouring lanes' bytes of the same ring slot), forms Lx, Ly of row R_(j-1) from the three rows held in registers, exchanges
// their +-S halo with the adjacent lanes by shuffle, and forms det of row R_(j-2).  No block barrier, no shared-memory tile:
// per output pixel 30 FP instructions of pinned arithmetic plus ~6 of data movement (the tile kernel k_prep2 spent 24 LDS.128
// per four pixels on these two stencils and sat on the shared-memory pipe, ncu r01l: LSU wavefronts 85 % busy).
//
// Borders.  The reference reflects INDICES (reflect-101) when a tap leaves the image, and neither operator commutes with
// mirroring (Lx / Ly are antisymmetric, the row sums are subtracted in a fixed order), so the streaming kernel does not
// try: pixels within S of the image border (Lx, Ly) and within 2 S (det) are recomputed afterwards from the planes in global
// memory by two small kernels with reflected indices (k_deriv_ring, k_det_ring: ~2 % of the pixels).
#include "common.cuh"
#include "kernels.h"
#include "level_math.cuh"
#include <algorithm>

using namespace akz;

namespace {

constexpr int D4_WARPS = 4;                 // warps per CTA, each on its own (strip, band, residue, frame) unit
constexpr int D4_RING = 6;                  // landing ring slots per warp: row j + 4 is requested while row j is used
constexpr int D4_SLOTB = 512 + 32;          // bytes per slot: 16 bytes of padding, 32 lanes x 16 bytes, 16 bytes of padding

struct Deriv4Args {
    const float* sm;
    float *lx, *ly, *det;
    unsigned char* hot;                     // optional: one byte per four pixels, "some determinant of the group passes the detector
                                            // threshold" -- k_extrema reads it instead of the determinant plane
    float thr; int ithr;
    long long plane;
    int w, h, pitch;
    int nstrips, nbands, band_h, nunits;
    LevelMathArgs m;
};

template <int S> struct D4 {
    static constexpr int HL = (2 * S + 3) / 4;          // halo lanes at either end of a strip (2 S columns)
    static constexpr int OC = (32 - 2 * HL) * 4;        // output columns of a strip
    static constexpr int RW = 4 + 2 * S;                // a register row: own 4 values with S neighbours on either side
};

template <int S>
struct Deriv4Regs {
    float Sm[3][D4<S>::RW];                 // smoothed rows R_(j-2), R_(j-1), R_j            (ring by row time mod 3)
    float Lx[3][D4<S>::RW], Ly[3][D4<S>::RW];   // derivative rows R_(j-3), R_(j-2), R_(j-1)  (ring by production time mod 3)
};

struct Deriv4Lane {
    const float* psm;                       // frame base + clamped column of this lane
    long long obase;                        // frame base + column of this lane (outputs)
    long long hbase;                        // frame base + group of this lane in the `hot` plane
    unsigned ring;                          // shared-memory address of this lane's 16 bytes in slot 0 of its warp's ring
    int y0, y1, r0;                         // band rows [y0, y1), row of time 0
    bool store;                             // lane writes output columns
};

__device__ __forceinline__ void d4_request(const Deriv4Args& a, const Deriv4Lane& ln, int row, int slot)
{
    const int r = min(max(row, 0), a.h - 1);
    const unsigned d = ln.ring + slot * D4_SLOTB;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(ln.psm + (long long)r * a.pitch) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__device__ __forceinline__ float d4_lds1(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float2 d4_lds2(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 d4_lds4(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

// columns x - S .. x + 3 + S of a ring slot (x = this lane's first column) with the widest aligned loads
template <int S>
__device__ __forceinline__ void d4_read_row(unsigned sa, float (&r)[D4<S>::RW])
{
    if (S == 4) {
        const float4 a = d4_lds4(sa - 16), b = d4_lds4(sa), c = d4_lds4(sa + 16);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
        r[8 % D4<S>::RW] = c.x; r[9 % D4<S>::RW] = c.y; r[10 % D4<S>::RW] = c.z; r[11 % D4<S>::RW] = c.w;
    } else if (S == 3) {
        const float a = d4_lds1(sa - 12);
        const float2 b = d4_lds2(sa - 8);
        const float4 c = d4_lds4(sa);
        const float2 d = d4_lds2(sa + 16);
        const float e = d4_lds1(sa + 24);
        r[0] = a; r[1] = b.x; r[2] = b.y; r[3] = c.x; r[4] = c.y; r[5] = c.z; r[6] = c.w; r[7] = d.x; r[8 % D4<S>::RW] = d.y; r[9 % D4<S>::RW] = e;
    } else {
        const float2 a = d4_lds2(sa - 8);
        const float4 b = d4_lds4(sa);
        const float2 c = d4_lds2(sa + 16);
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = b.z; r[5] = b.w; r[6] = c.x; r[7] = c.y;
    }
}

// one row time of a residue class; PH = row time mod 6
template <int S, bool INT, int PH>
__device__ __forceinline__ void d4_row(Deriv4Regs<S>& R, const Deriv4Args& a, const Deriv4Lane& ln, int j)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RW = D4<S>::RW;
    // ---- landing ring: request row j + 4 (its slot held row j - 2, read two row times ago by every lane that has passed the
    // warp barrier of row time j - 1), wait for this lane's copy of row j, make all lanes' copies visible
    d4_request(a, ln, ln.r0 + (j + D4_RING - 2) * S, (PH + D4_RING - 2) % D4_RING);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(D4_RING - 2) : "memory");
    __syncwarp();
    d4_read_row<S>(ln.ring + PH * D4_SLOTB, R.Sm[PH % 3]);
    // ---- first derivatives of row R_(j-1)
    const int r1 = ln.r0 + (j - 1) * S;
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
        for (int i = 0; i < S; i++) {
            lx[i] = __shfl_up_sync(FULL, vx[4 - S + i], 1);             // columns x - S + i of the left neighbour lane
            ly[i] = __shfl_up_sync(FULL, vy[4 - S + i], 1);
            lx[(S + 4 + i) % RW] = __shfl_down_sync(FULL, vx[i], 1);   // columns x + 4 + i of the right neighbour lane
            ly[(S + 4 + i) % RW] = __shfl_down_sync(FULL, vy[i], 1);
        }
    }
    // ---- determinant of row R_(j-2)
    const int r2 = ln.r0 + (j - 2) * S;
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
        if (ln.store && r2 >= ln.y0 && r2 < ln.y1) {
            *reinterpret_cast<float4*>(a.det + ln.obase + (long long)r2 * a.pitch) = make_float4(o[0], o[1], o[2], o[3]);
            if (a.hot) {
                const bool h = INT ? (fi(o[0]) > a.ithr || fi(o[1]) > a.ithr || fi(o[2]) > a.ithr || fi(o[3]) > a.ithr)
                                   : (o[0] > a.thr || o[1] > a.thr || o[2] > a.thr || o[3] > a.thr);
                a.hot[ln.hbase + (long long)r2 * (a.pitch >> 2)] = h ? 1 : 0;
            }
        }
    }
}

template <int S, bool INT>
__global__ void __launch_bounds__(32 * D4_WARPS, (S <= 3 ? 4 : 3)) k_deriv4(const __grid_constant__ Deriv4Args a)
{
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * D4_WARPS + (threadIdx.x >> 5);
    if (unit >= a.nunits) return;                       // whole warps leave: no block barrier is used
    // strips fastest, then the residue classes of a band (their rows interleave in memory), then bands, then frames
    int rem = unit;
    const int strip = rem % a.nstrips; rem /= a.nstrips;
    const int rho = rem % S; rem /= S;
    const int band = rem % a.nbands;
    const int frame = rem / a.nbands;
    Deriv4Lane ln;
    const int gx0 = strip * D4<S>::OC - 4 * D4<S>::HL + 4 * lane;
    const int gxl = min(max(gx0, 0), a.pitch - 4);
    const long long base = (long long)frame * a.plane;
    ln.psm = a.sm + base + gxl;
    ln.obase = base + gx0;
    ln.hbase = (long long)frame * (a.pitch >> 2) * a.h + (gx0 >> 2);
    ln.y0 = band * a.band_h; ln.y1 = min(a.h, ln.y0 + a.band_h);
    ln.r0 = ln.y0 + rho - 2 * S;
    ln.store = lane >= D4<S>::HL && lane < 32 - D4<S>::HL && gx0 >= 0 && gx0 < a.w;
    __shared__ __align__(16) unsigned char ring_mem[D4_WARPS * D4_RING * D4_SLOTB];
    ln.ring = (unsigned)__cvta_generic_to_shared(ring_mem + (threadIdx.x >> 5) * (D4_RING * D4_SLOTB) + 16 + lane * 16);

    Deriv4Regs<S> R;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < D4<S>::RW; c++) { R.Sm[i][c] = 0.f; R.Lx[i][c] = 0.f; R.Ly[i][c] = 0.f; }
    // rows of the band in this residue class, plus two row times before (derivative rows the first determinant needs) and two after
    // (the loop ends when the determinant row of row time j, r0 + (j - 2) S, has left the band: written with the values that are live
    // anyway -- a trip count kept in a register was spilled by ptxas at 128 registers and reloaded every six rows: 17 % of the
    // kernel's stall samples sat on that local load, ncu r02z)
#pragma unroll
    for (int k = 0; k < D4_RING - 2; k++) d4_request(a, ln, ln.r0 + k * S, k);
    for (int j = 0; ln.r0 + (j - 2) * S < ln.y1; j += 6) {
        d4_row<S, INT, 0>(R, a, ln, j);
        d4_row<S, INT, 1>(R, a, ln, j + 1);
        d4_row<S, INT, 2>(R, a, ln, j + 2);
        d4_row<S, INT, 3>(R, a, ln, j + 3);
        d4_row<S, INT, 4>(R, a, ln, j + 4);
        d4_row<S, INT, 5>(R, a, ln, j + 5);
    }
}

// ---- border ring: the pixels whose taps leave the image, recomputed with reflected indices -----------------------------------
// pixel i of the ring of width B of a w x h image: rows [0, B) and [h-B, h) over the full width, then columns [0, B) and
// [w-B, w) over the rows between
__device__ __forceinline__ bool ring_pixel(int i, int B, int w, int h, int& x, int& y)
{
    const int bt = min(B, (h + 1) / 2), bb = min(B, h - bt);          // top / bottom rows (they never overlap)
    if (i < bt * w) { y = i / w; x = i - y * w; return true; }
    i -= bt * w;
    if (i < bb * w) { y = i / w; x = i - y * w; y += h - bb; return true; }
    i -= bb * w;
    const int mid = h - bt - bb, bl = min(B, (w + 1) / 2), br = min(B, w - bl);
    if (mid <= 0) return false;
    if (i < mid * bl) { y = i / bl; x = i - y * bl; y += bt; return true; }
    i -= mid * bl;
    if (i < mid * br) { y = i / br; x = i - y * br; y += bt; x += w - br; return true; }
    return false;
}
__host__ __device__ inline int ring_count(int B, int w, int h)
{
    const int bt = B < (h + 1) / 2 ? B : (h + 1) / 2, bb = B < h - bt ? B : h - bt;
    const int mid = h - bt - bb, bl = B < (w + 1) / 2 ? B : (w + 1) / 2, br = B < w - bl ? B : w - bl;
    return (bt + bb) * w + (mid > 0 ? mid * (bl + br) : 0);
}

struct Nb8 { float ul, uc, ur, cl, cr, ll, lc, lr; };
__device__ __forceinline__ Nb8 ring_load(const float* __restrict__ p, int x, int y, int w, int h, int pitch, int S)
{
    const int x0 = refl(x - S, w), x2 = refl(x + S, w), y0 = refl(y - S, h), y2 = refl(y + S, h);
    const float* r0 = p + (long long)y0 * pitch;
    const float* r1 = p + (long long)y * pitch;
    const float* r2 = p + (long long)y2 * pitch;
    Nb8 n;
    n.ul = r0[x0]; n.uc = r0[x]; n.ur = r0[x2]; n.cl = r1[x0]; n.cr = r1[x2]; n.ll = r2[x0]; n.lc = r2[x]; n.lr = r2[x2];
    return n;
}

template <bool INT>
__global__ void __launch_bounds__(256) k_deriv_ring(const float* __restrict__ sm, float* __restrict__ lx, float* __restrict__ ly, int S, LevelMathArgs m,
                                                    int w, int h, int pitch, long long plane, int count)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    int x, y;
    if (i >= count || !ring_pixel(i, S, w, h, x, y)) return;
    const long long base = (long long)blockIdx.y * plane;
    const Nb8 n = ring_load(sm + base, x, y, w, h, pitch, S);
    float vx, vy;
    p2_deriv1<INT>(n.ul, n.uc, n.ur, n.cl, n.cr, n.ll, n.lc, n.lr, m, vx, vy);
    lx[base + (long long)y * pitch + x] = vx;
    ly[base + (long long)y * pitch + x] = vy;
}

template <bool INT>
__global__ void __launch_bounds__(256) k_det_ring(const float* __restrict__ lx, const float* __restrict__ ly, float* __restrict__ det, int S, LevelMathArgs m,
                                                  int w, int h, int pitch, long long plane, int count)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    int x, y;
    if (i >= count || !ring_pixel(i, 2 * S, w, h, x, y)) return;
    const long long base = (long long)blockIdx.y * plane;
    const Nb8 a = ring_load(lx + base, x, y, w, h, pitch, S);
    const Nb8 b = ring_load(ly + base, x, y, w, h, pitch, S);
    det[base + (long long)y * pitch + x] = p2_det<INT>(a.ul, a.uc, a.ur, a.cl, a.cr, a.ll, a.lc, a.lr, b.ul, b.uc, b.ur, b.ll, b.lc, b.lr, m);
}

// The two border-ring kernels are a few microseconds of work that depend on the main kernel: with a side stream they are forked
// off (event after the main kernel) and run under whatever the main stream does next -- the diffusion cycle of the level, which
// touches none of these planes.  The caller joins ev_join before it overwrites the blurred plane or reads Lx / Ly / det.
template <bool INT>
void rings_launch(cudaStream_t st, const float* sm, float* lx, float* ly, float* det, int S, const LevelMathArgs& m, int w, int h, int pitch, long long plane,
                  int n, cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    const int c1 = ring_count(S, w, h), c2 = ring_count(2 * S, w, h);
    cudaStream_t rs = st;
    if (ring_st && ev_fork && ev_join) {
        cudaEventRecord(ev_fork, st);
        cudaStreamWaitEvent(ring_st, ev_fork, 0);
        rs = ring_st;
    }
    k_deriv_ring<INT><<<dim3((c1 + 255) / 256, n), 256, 0, rs>>>(sm, lx, ly, S, m, w, h, pitch, plane, c1);
    k_det_ring<INT><<<dim3((c2 + 255) / 256, n), 256, 0, rs>>>(lx, ly, det, S, m, w, h, pitch, plane, c2);
    if (rs != st) cudaEventRecord(ev_join, rs);
}

template <int S, bool INT>
void deriv4_launch(cudaStream_t st, Deriv4Args& a, int n, cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    a.nstrips = (a.w + D4<S>::OC - 1) / D4<S>::OC;
    // Bands are multiples of 12 rows (every band starts on residue 0 of every S) and cost two row times of warm-up before and
    // after per residue class.  Their number is chosen by a small model: (waves of CTAs the GPU needs) x (row times of a unit).
    // Many frames: ~270-row bands (octave 0 of 32 frames: 4 bands, 3 waves).  Few frames or small levels: shorter bands until the
    // units fill one wave -- but not 1.1 waves (960 x 540, S = 3, 32 frames: 3 bands = 648 CTAs for 592 slots took 58 us, 2 bands
    // 48 us; one 1080p frame with 270-row bands: 120-240 units, 22-30 us per launch on the critical path of the graph).
    static const int nsm = [] { int d = 0, v = 148; cudaGetDevice(&d); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148; return v; }();
    const long long per_band = (long long)n * S * a.nstrips;
    const long long slots = (long long)nsm * (S <= 3 ? 4 : 3);          // resident CTAs (launch bounds of k_deriv4)
    const int nb_lo = std::max(1, (a.h + 269) / 270), nb_hi = std::max(nb_lo, (a.h + 23) / 24);
    long long best_cost = -1;
    int best_bh = a.h;
    for (int nb = nb_lo; nb <= nb_hi; nb++) {
        const int bh = std::max(24, ((a.h + nb - 1) / nb + 11) / 12 * 12);
        const int nbands = (a.h + bh - 1) / bh;
        const long long ctas = (per_band * nbands + D4_WARPS - 1) / D4_WARPS;
        const long long cost = ((ctas + slots - 1) / slots) * (bh / S + 4);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bh = bh; }
    }
    a.band_h = best_bh;
    a.nbands = (a.h + a.band_h - 1) / a.band_h;
    a.nunits = n * a.nbands * S * a.nstrips;
    k_deriv4<S, INT><<<(a.nunits + D4_WARPS - 1) / D4_WARPS, 32 * D4_WARPS, 0, st>>>(a);
    rings_launch<INT>(st, a.sm, a.lx, a.ly, a.det, S, a.m, a.w, a.h, a.pitch, a.plane, n, ring_st, ev_fork, ev_join);
}

}  // namespace

namespace akzk {

// Lx, Ly, det of a level from its smoothed plane.  ring_st / ev_fork / ev_join (optional): side stream for the border-ring
// kernels, see deriv4_launch.  Returns the number of launches (3), or 0 when the case is not covered
// (derivative step outside 2..4, rows not 16-byte aligned, width not a multiple of 4, tiny planes): the caller then uses the
// tile kernel or the per-stage kernels.
int deriv_stream(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long plane,
                 int n, int int_planes, cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join, unsigned char* hot, float thr, int ithr)
{
    if (step < 2 || step > 4 || w < 32 || h < 32 || (w % 4) != 0 || (pitch % 4) != 0 || (plane % 4) != 0) return 0;
    if ((((uintptr_t)smooth | (uintptr_t)lx | (uintptr_t)ly | (uintptr_t)det) % 16) != 0) return 0;
    if (smooth == det || smooth == lx || smooth == ly) return 0;      // not an in-place kernel
    Deriv4Args a = {};
    a.sm = smooth; a.lx = lx; a.ly = ly; a.det = det; a.plane = plane; a.w = w; a.h = h; a.pitch = pitch;
    a.hot = hot; a.thr = thr; a.ithr = ithr;
    hessian_factors(&a.m.fac1, &a.m.fac2);
    a.m.ifac1 = (int)(a.m.fac1 * 65536 + 0.5f); a.m.ifac2 = (int)(a.m.fac2 * 65536 + 0.5f);          // akazed.cu:4184-4185
    if (int_planes) {
        if (step == 2) deriv4_launch<2, true>(st, a, n, ring_st, ev_fork, ev_join); else if (step == 3) deriv4_launch<3, true>(st, a, n, ring_st, ev_fork, ev_join); else deriv4_launch<4, true>(st, a, n, ring_st, ev_fork, ev_join);
    } else {
        if (step == 2) deriv4_launch<2, false>(st, a, n, ring_st, ev_fork, ev_join); else if (step == 3) deriv4_launch<3, false>(st, a, n, ring_st, ev_fork, ev_join); else deriv4_launch<4, false>(st, a, n, ring_st, ev_fork, ev_join);
    }
    return 3;
}

// the border-ring kernels alone (k_level4 in level_stream.cu computes the interior itself); returns the number of launches
int deriv_rings(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long plane,
                int n, int int_planes, cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    LevelMathArgs m = {};
    hessian_factors(&m.fac1, &m.fac2);
    m.ifac1 = (int)(m.fac1 * 65536 + 0.5f); m.ifac2 = (int)(m.fac2 * 65536 + 0.5f);
    if (int_planes) rings_launch<true>(st, smooth, lx, ly, det, step, m, w, h, pitch, plane, n, ring_st, ev_fork, ev_join);
    else rings_launch<false>(st, smooth, lx, ly, det, step, m, w, h, pitch, plane, n, ring_st, ev_fork, ev_join);
    return 2;
}

}  // namespace akzk
