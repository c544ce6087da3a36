// Detector: per-level 3x3 extrema -> deterministic cross-level arg-max merge on a full-resolution
// key map -> radius NMS -> sub-pixel refinement -> raster-ordered compaction.
// Reference: gCalcExtremaMap (akazed.cu:1334-1393), gNmsRNaive (:1554-1613), gRefine (:1615-1662).
// Differences from the reference, all deliberate:
//   - the three float/int maps are one 64-bit key map merged with atomicMax: arg-max over levels,
//     ties to the lowest layer, no tearing;
//   - survivors are compacted in raster order (row counts -> scan -> emit), not atomicInc order;
//   - the keypoint count never leaves the device;
//   - the key map is sparse (a few thousand candidates per 2 M pixels): k_extrema also sets the candidate's bit in an occupancy
//     bitmap (one word per 32 pixels); the NMS pass reads the bitmap and touches the 64-bit map only around candidates, and
//     k_clear_map zeroes exactly the entries that were written.  Round 1 cleared the whole map (531 MB per 32 frames) and
//     read all of it again (k_nms_mark: 2.5 ms per 256 frames, 1.7 TB/s).
#include "common.cuh"
#include "kernels.h"

using namespace akz;

namespace {

// grid: (ceil((w - x_base)/128), ceil((h-2psz)/8), nframes*nsub); 4 consecutive pixels per thread.
// Almost every pixel fails the threshold, so a thread first looks at one float4 of the centre row and only
// candidates pay for the two neighbour rows (ncu r01b: the one-pixel-per-thread version was latency bound,
// 9.4 warps stalled on the long scoreboard per issue).
// T = float: the float pipeline; T = int: the integer "fast" pipeline (gCalcExtremaMap akazed.cu:3476, threshold 65)
template <typename T> struct ExtVec;
template <> struct ExtVec<float> { typedef float4 type; };
template <> struct ExtVec<int> { typedef int4 type; };
__device__ __forceinline__ unsigned long long ext_key(float v, int layer) { return merge_key(v, layer); }
__device__ __forceinline__ unsigned long long ext_key(int v, int layer) { return ((unsigned long long)(unsigned)v << 32) | (unsigned)(0xFFFF - layer); }

__device__ __forceinline__ float ext_thr(const AkzExtremaLevel& L, float) { return L.threshold; }
__device__ __forceinline__ int ext_thr(const AkzExtremaLevel& L, int) { return L.ithreshold; }

// the 3 x 3 test of one group of four pixels (x0 .. x0 + 3 of row iy) and the merge of its maxima into the key map
template <typename T>
__device__ __forceinline__ void ext_eval_group(const AkzExtremaArgs& a, const AkzExtremaLevel& L, int frame, int iy, int x0,
                                               unsigned long long* __restrict__ map, int mpitch, long long mplane,
                                               unsigned* __restrict__ occ, int mwords, int H)
{
    typedef typename ExtVec<T>::type V4;
    const T* row = reinterpret_cast<const T*>(L.det) + (long long)frame * L.plane + (long long)iy * a.pitch + x0;
    const V4 c4 = __ldg(reinterpret_cast<const V4*>(row));
    const T thr = ext_thr(L, T());
    if (!(c4.x > thr || c4.y > thr || c4.z > thr || c4.w > thr)) return;
    const V4 u4 = __ldg(reinterpret_cast<const V4*>(row - a.pitch));
    const V4 d4 = __ldg(reinterpret_cast<const V4*>(row + a.pitch));
    const T u[6] = { __ldg(row - a.pitch - 1), u4.x, u4.y, u4.z, u4.w, __ldg(row - a.pitch + 4) };
    const T c[6] = { __ldg(row - 1), c4.x, c4.y, c4.z, c4.w, __ldg(row + 4) };
    const T d[6] = { __ldg(row + a.pitch - 1), d4.x, d4.y, d4.z, d4.w, __ldg(row + a.pitch + 4) };
    const float border = L.border;
    // akazed.cu:1346-1353 (float arithmetic, truncating casts)
    int up_y = (int)(__fadd_rn(__fsub_rn((float)iy, border), 0.5f)) - 1;
    int down_y = (int)(__fadd_rn(__fadd_rn((float)iy, border), 0.5f)) + 1;
    if (up_y < 0 || down_y >= a.h) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int ix = x0 + j;
        T v = c[j + 1];
        if (ix < a.psz || ix >= a.w - 1 || !(v > thr)) continue;
        int left_x = (int)(__fadd_rn(__fsub_rn((float)ix, border), 0.5f)) - 1;
        int right_x = (int)(__fadd_rn(__fadd_rn((float)ix, border), 0.5f)) + 1;
        if (left_x < 0 || right_x >= a.w) continue;
        if (v > u[j + 1] && v > d[j + 1] && v > c[j] && v > c[j + 2] && v > u[j] && v > u[j + 2] && v > d[j] && v > d[j + 2]) {
            const int Y = iy << a.octave, X = ix << a.octave;
            long long oi = (long long)frame * mplane + (long long)Y * mpitch + X;
            atomicMax(map + oi, ext_key(v, L.layer));
            atomicOr(occ + ((long long)frame * H + Y) * mwords + (X >> 5), 1u << (X & 31));
        }
    }
}

// grid: (ceil((w - x_base)/128), ceil((h-2psz)/8), nframes*nsub); 4 consecutive pixels per thread, one row
template <typename T>
__global__ void __launch_bounds__(256) k_extrema(const __grid_constant__ AkzExtremaArgs a, unsigned long long* __restrict__ map, int mpitch, long long mplane,
                                                 unsigned* __restrict__ occ, int mwords, int H)
{
    int frame = blockIdx.z / a.nsub, sub = blockIdx.z - frame * a.nsub;
    const AkzExtremaLevel& L = a.lv[sub];
    const int x0 = (a.psz & ~3) + (blockIdx.x * 32 + threadIdx.x) * 4;
    const int iy = blockIdx.y * 8 + threadIdx.y + a.psz;
    if (x0 >= a.w - 1 || iy >= a.h - 1) return;
    ext_eval_group<T>(a, L, frame, iy, x0, map, mpitch, mplane, occ, mwords, H);
}

// k_extrema_h: the levels whose derivative kernel left a `hot` plane (one byte per four pixels: "a determinant of the group
// passes the threshold").  A thread reads eight of those bytes (32 pixels) with one load and goes on only for the groups that
// are set -- under 1 % on real frames.  k_extrema is one thread, one 128-bit load and one exit per four pixels: 63 M threads
// per octave-0 level of 32 frames, every SM turning over 8-warp blocks that live for one memory round trip: 293 us whatever
// the load brings (with the byte instead of the four determinants: the same 293 us).  This form launches an eighth of the
// threads and moves a sixteenth of the bytes.
template <typename T>
__global__ void __launch_bounds__(256) k_extrema_h(const __grid_constant__ AkzExtremaArgs a, unsigned long long* __restrict__ map, int mpitch, long long mplane,
                                                   unsigned* __restrict__ occ, int mwords, int H)
{
    int frame = blockIdx.z / a.nsub, sub = blockIdx.z - frame * a.nsub;
    const AkzExtremaLevel& L = a.lv[sub];
    const int lane = threadIdx.x;
    const int wx0 = blockIdx.x * 1024;                                  // a warp covers 1024 pixels of one row
    const int x0 = wx0 + lane * 32;                                     // 8 groups of 4 pixels; rows are padded to 32 pixels
    const int iy = blockIdx.y * 8 + threadIdx.y + a.psz;
    if (iy >= a.h - 1) return;                                          // whole warps leave
    uint2 hb = make_uint2(0u, 0u);
    if (x0 < a.w - 1) hb = __ldg(reinterpret_cast<const uint2*>(L.hot + ((long long)frame * a.h + iy) * (a.pitch >> 2) + (x0 >> 2)));
    if (!__any_sync(0xffffffffu, (hb.x | hb.y) != 0u)) return;
    // the warp's set groups are dealt out to its lanes (one 3 x 3 test per lane and round) instead of every lane walking its own
    // eight: a warp holds ~5 of them, spread over as many lanes and groups.  (Four rows per warp -- four loads in flight, 32
    // masks to deal out -- was 2.4 times slower: the dealing is executed by nearly every warp.)
    unsigned m[8];
    int total = 0;
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const unsigned byte = ((g < 4 ? hb.x : hb.y) >> (8 * (g & 3))) & 0xFFu;
        m[g] = __ballot_sync(0xffffffffu, byte != 0u && x0 + 4 * g < a.w - 1);
        total += __popc(m[g]);
    }
    for (int k = lane; k < total; k += 32) {
        // the k-th set group of the warp: which mask, then which bit of it (one __fns: it is a software loop)
        int off = k, gi = 0;
        unsigned msel = m[0];
#pragma unroll
        for (int g = 0; g < 7; g++) {
            const int cnt = __popc(m[g]);
            if (gi == g && off >= cnt) { off -= cnt; gi = g + 1; msel = m[g + 1]; }
        }
        const int xg = wx0 + 32 * (int)__fns(msel, 0u, off + 1) + 4 * gi;
        ext_eval_group<T>(a, L, frame, iy, xg, map, mpitch, mplane, occ, mwords, H);
    }
}

// one thread per word of the occupancy bitmap (32 pixels of a row): radius NMS decision for its candidates -> survivor mask.
// (no block barrier and no counter: the row counts are popcounts of the masks, taken by k_row_scan)
__global__ void __launch_bounds__(256) k_nms_mark(const unsigned long long* __restrict__ map, int mpitch, long long mplane,
                                                  int W, int H, int psz, const __grid_constant__ AkzLevelTable tab,
                                                  const unsigned* __restrict__ occ, unsigned* __restrict__ rowmask, int mwords)
{
    const int frame = blockIdx.y;
    const int wi = blockIdx.x * 256 + threadIdx.x;               // word of rows [psz, H - psz)
    if (wi >= (H - 2 * psz) * mwords) return;
    const int iy = psz + wi / mwords, word = wi - (iy - psz) * mwords;
    const long long wpos = ((long long)frame * H + iy) * mwords + word;
    unsigned cand = __ldg(occ + wpos), keepmask = 0;
    const unsigned long long* m = map + (long long)frame * mplane;
    const int xend = W - psz;                                     // ix + psz < W
    while (cand) {
        const int bit = __ffs(cand) - 1;
        cand &= cand - 1;
        const int ix = word * 32 + bit;
        if (ix < psz || ix >= xend) continue;
        const unsigned long long key = m[(long long)iy * mpitch + ix];
        if (key == 0ull) continue;
        const unsigned rc = (unsigned)(key >> 32);                // positive float bits and positive ints order like unsigned
        const float fsz = tab.lv[key_layer(key)].size;
        const int isz = (int)__fadd_rn(fsz, 0.5f);
        const int sq = (int)__fmul_rn(fsz, fsz);
        bool keep = true;
        if (isz <= 4) {
            // Neighbours exist only where the occupancy bitmap has a bit: the window's rows of the bitmap first (at most 9 x 2
            // words, all in flight together), then the keys of the occupied cells only (a candidate has ~0.5 per window row).
            // The first version read the 64-bit map cell by cell with an early exit: up to 81 dependent round trips per candidate
            // (58 us for one 1080p frame) and 8 bytes per cell.
            const int xlo = ix - isz, w0 = xlo >> 5, sh = xlo & 31, span = 2 * isz + 1;
            const bool two = sh + span > 32 && w0 + 1 < mwords;
            unsigned field[9];
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int i = t - 4;
                field[t] = 0u;
                if (i >= -isz && i <= isz) {
                    const unsigned* orow = occ + ((long long)frame * H + iy + i) * mwords + w0;
                    const unsigned lo = __ldg(orow), hi = two ? __ldg(orow + 1) : 0u;
                    field[t] = __funnelshift_r(lo, hi, sh) & ((1u << span) - 1u);        // bit b: column ix - isz + b
                }
            }
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int i = t - 4;
                if (!keep || field[t] == 0u) continue;
                const unsigned long long* row = m + (long long)(iy + i) * mpitch + ix;
                for (int j = -isz; j <= isz; j++) {
                    if ((i == 0 && j == 0) || i * i + j * j >= sq) continue;
                    // akazed.cu:1578-1581: the reference's `continue` at the centre skips its `new_idx++`, so on
                    // the centre row every j > 0 examines the pixel at offset j-1 (the centre itself for j = 1)
                    // under the distance test of j.  Reproduced: the keypoint SET must equal the reference's.
                    const int e = (i == 0 && j > 0) ? j - 1 : j;
                    if (!((field[t] >> (e + isz)) & 1u)) continue;
                    const unsigned long long kn = row[e];
                    if (kn == 0ull) continue;
                    const unsigned rn = (unsigned)(kn >> 32);
                    if (rn > rc || (rn == rc && i <= 0 && j <= 0)) { keep = false; break; }
                }
            }
        } else {
            for (int i = -isz; i <= isz && keep; i++) {
                const unsigned long long* row = m + (long long)(iy + i) * mpitch + ix;
                for (int j = -isz; j <= isz; j++) {
                    if ((i == 0 && j == 0) || i * i + j * j >= sq) continue;
                    const unsigned long long kn = row[(i == 0 && j > 0) ? j - 1 : j];
                    if (kn == 0ull) continue;
                    const unsigned rn = (unsigned)(kn >> 32);
                    if (rn > rc || (rn == rc && i <= 0 && j <= 0)) { keep = false; break; }
                }
            }
        }
        if (keep) keepmask |= 1u << bit;
    }
    rowmask[wpos] = keepmask;
}

// zero the map entries that k_extrema wrote and the bitmap itself: the map is clean for the next chunk without a memset
__global__ void __launch_bounds__(256) k_clear_map(unsigned long long* __restrict__ map, int mpitch, long long mplane, int H,
                                                   unsigned* __restrict__ occ, int mwords)
{
    const int frame = blockIdx.y;
    const int wi = blockIdx.x * 256 + threadIdx.x;
    if (wi >= H * mwords) return;
    const int iy = wi / mwords, word = wi - iy * mwords;
    unsigned* o = occ + (long long)frame * H * mwords + wi;
    unsigned cand = *o;
    if (!cand) return;
    *o = 0u;
    unsigned long long* row = map + (long long)frame * mplane + (long long)iy * mpitch + word * 32;
    while (cand) {
        const int bit = __ffs(cand) - 1;
        cand &= cand - 1;
        row[bit] = 0ull;
    }
}

// one block per frame: exclusive scan of the row counts (in place) -> per-frame total
__global__ void __launch_bounds__(1024) k_row_scan(int* __restrict__ rowcount, const unsigned* __restrict__ rowmask, int mwords, int H, int psz,
                                                   int* __restrict__ counts, int* __restrict__ totals, int max_pts)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    int frame = blockIdx.x;
    int* rc = rowcount + (long long)frame * H;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lo = psz, hi = H - psz;
    for (int base = lo; base < hi; base += 1024) {
        int y = base + threadIdx.x;
        int v = 0;
        if (y < hi) {                                     // survivors of row y = set bits of its mask words
            const unsigned* mw = rowmask + ((long long)frame * H + y) * mwords;
            if ((mwords & 3) == 0) {                      // rows of 16-byte multiples: 128-bit loads, four in flight
                const uint4* m4 = reinterpret_cast<const uint4*>(mw);
#pragma unroll 4
                for (int k = 0; k < mwords / 4; k++) { const uint4 q = __ldg(m4 + k); v += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w); }
            } else {
                for (int k = 0; k < mwords; k++) v += __popc(__ldg(mw + k));
            }
        }
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int ws = warp_sums[threadIdx.x], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, d);
                if (threadIdx.x >= d) wi += t;
            }
            warp_sums[threadIdx.x] = wi - ws;            // exclusive
        }
        __syncthreads();
        int excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if (y < hi) rc[y] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[frame] = carry;                            // survivors before clamping (reference: the raw counter)
        counts[frame] = min(carry, max_pts);              // akaze.cpp:451
    }
}

// exclusive prefix of the clamped counts over the frames of the chunk: prefix[0..n], prefix[n] = total
__global__ void k_frame_prefix(const int* __restrict__ counts, int* __restrict__ prefix, int n)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < n; i++) { prefix[i] = s; s += counts[i]; }
        prefix[n] = s;
    }
}

// one block (128 threads) per (row, frame): rank the survivors of the row and write refined keypoints
template <bool INT>
__global__ void __launch_bounds__(128) k_emit_refine(const unsigned long long* __restrict__ map, int mpitch, long long mplane,
                                                     int H, int psz, const __grid_constant__ AkzLevelTable tab, const unsigned* __restrict__ rowmask, int mwords,
                                                     const int* __restrict__ rowoff, akz_keypoint* __restrict__ kpts, int max_pts)
{
    int iy = blockIdx.x + psz, frame = blockIdx.y;
    __shared__ int s_pref[129];
    unsigned word = 0;
    if ((int)threadIdx.x < mwords) word = rowmask[((long long)frame * H + iy) * mwords + threadIdx.x];
    // mwords <= 128 (W <= 4096)
    int c = __popc(word);
    s_pref[threadIdx.x + 1] = c;
    if (threadIdx.x == 0) s_pref[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 1; i <= 128; i++) s_pref[i] += s_pref[i - 1];
    __syncthreads();
    if (!word) return;
    int rank = rowoff[(long long)frame * H + iy] + s_pref[threadIdx.x];
    const unsigned long long* m = map + (long long)frame * mplane + (long long)iy * mpitch;
    akz_keypoint* out = kpts + (long long)frame * max_pts;
    while (word) {
        int bit = __ffs(word) - 1;
        word &= word - 1;
        if (rank >= max_pts) break;
        int ix = threadIdx.x * 32 + bit;
        unsigned long long key = m[ix];
        int layer = key_layer(key);
        const AkzLevelDev& L = tab.lv[layer];
        int o = L.octave, p = L.pitch;
        int x = ix >> o, y = iy >> o;
        akz_keypoint k;
        k.ix = ix; k.iy = iy; k.layer = layer; k.size = L.size; k.angle = 0.f;
        if (!INT) {
            const float* d = L.det + (long long)frame * L.plane + (long long)y * p + x;
            // akazed.cu:1636-1657, operation order as compiled
            float d0 = __ldg(d), dl = __ldg(d - 1), dr = __ldg(d + 1), du = __ldg(d - p), dd_ = __ldg(d + p);
            float v2 = __fadd_rn(d0, d0);
            float gx = __fmul_rn(0.5f, __fsub_rn(dr, dl));
            float gy = __fmul_rn(0.5f, __fsub_rn(dd_, du));
            float dxx = __fsub_rn(__fadd_rn(dr, dl), v2);
            float dyy = __fsub_rn(__fadd_rn(dd_, du), v2);
            float dxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(__ldg(d + p + 1), __ldg(d - p - 1)), __ldg(d - p + 1)), __ldg(d + p - 1)));
            float det = __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy));
            float idd = det != 0.f ? __fdiv_rn(1.f, det) : 0.f;
            float o0 = __fmul_rn(idd, __fmaf_rn(gy, dxy, -__fmul_rn(gx, dyy)));
            float o1 = __fmul_rn(idd, __fmaf_rn(gx, dxy, -__fmul_rn(gy, dxx)));
            bool weak = o0 < -1.f || o0 > 1.f || o1 < -1.f || o1 > 1.f;
            k.response = key_resp(key);
            if (weak) { k.x = (float)ix; k.y = (float)iy; }
            else {
                float ratio = (float)(1 << o);
                k.x = __fmul_rn(ratio, __fadd_rn((float)x, o0));
                k.y = __fmul_rn(ratio, __fadd_rn((float)y, o1));
            }
        } else {
            // integer pipeline, gRefine akazed.cu:3600-3646: integer differences (arithmetic shifts), float quotient
            const int* d = reinterpret_cast<const int*>(L.det) + (long long)frame * L.plane + (long long)y * p + x;
            int d0 = __ldg(d), dl = __ldg(d - 1), dr = __ldg(d + 1), du = __ldg(d - p), dd_ = __ldg(d + p);
            int v2 = d0 + d0;
            int gx = (dr - dl) >> 1;
            int gy = (dd_ - du) >> 1;
            int dxx = dr + dl - v2;
            int dyy = dd_ + du - v2;
            int dxy = (__ldg(d + p + 1) + __ldg(d - p - 1) - __ldg(d - p + 1) - __ldg(d + p - 1)) >> 2;
            int det = dxx * dyy - dxy * dxy;
            float idd = det != 0 ? (1.f / det) : 0.f;
            float o0 = idd * (dxy * gy - dyy * gx);
            float o1 = idd * (dxy * gx - dxx * gy);
            bool weak = o0 < -1.f || o0 > 1.f || o1 < -1.f || o1 > 1.f;
            k.response = (float)(int)(key >> 32);
            if (weak) { k.x = (float)ix; k.y = (float)iy; }
            else {
                int ratio = 1 << o;
                k.y = ratio * (y + o1);
                k.x = ratio * (x + o0);
            }
        }
        out[rank] = k;
        rank++;
    }
}

}  // namespace

namespace akzk {

int extrema(cudaStream_t st, const AkzExtremaArgs& a, unsigned long long* map, int mpitch, long long mplane, unsigned* occ, int mwords, int H, int n)
{
    int ew = a.w - 2 * a.psz, eh = a.h - 2 * a.psz;
    if (ew <= 0 || eh <= 0) return 0;
    int xb = a.psz & ~3;
    bool all_hot = (a.pitch % 32) == 0;
    for (int j = 0; j < a.nsub; j++) all_hot = all_hot && a.lv[j].hot != nullptr;
    if (all_hot) {
        dim3 gh((a.w + 1023) / 1024, (eh + 7) / 8, n * a.nsub);
        if (a.int_planes) k_extrema_h<int><<<gh, dim3(32, 8), 0, st>>>(a, map, mpitch, mplane, occ, mwords, H);
        else k_extrema_h<float><<<gh, dim3(32, 8), 0, st>>>(a, map, mpitch, mplane, occ, mwords, H);
        return 1;
    }
    dim3 g((a.w - xb + 127) / 128, (eh + 7) / 8, n * a.nsub);
    if (a.int_planes) k_extrema<int><<<g, dim3(32, 8), 0, st>>>(a, map, mpitch, mplane, occ, mwords, H);
    else k_extrema<float><<<g, dim3(32, 8), 0, st>>>(a, map, mpitch, mplane, occ, mwords, H);
    return 1;
}

int frame_prefix(cudaStream_t st, const int* counts, int* prefix, int n)
{
    k_frame_prefix<<<1, 32, 0, st>>>(counts, prefix, n);
    return 1;
}

int nms_emit(cudaStream_t st, unsigned long long* map, int mpitch, long long mplane, int W, int H, int psz,
             const AkzLevelTable& tab, unsigned* occ, unsigned* rowmask, int* rowcount, int* counts, int* prefix,
             akz_keypoint* kpts, int max_pts, int n, int int_planes)
{
    int mwords = (W + 31) / 32;
    if (mwords > 128) return akz_set_error(AKZ_E_UNSUPPORTED, "frame width above 4096 is not supported by the compaction kernel");
    int rows = H - 2 * psz;
    int launches = 0;
    if (rows > 0) {
        k_nms_mark<<<dim3((rows * mwords + 255) / 256, n), 256, 0, st>>>(map, mpitch, mplane, W, H, psz, tab, occ, rowmask, mwords);
        launches++;
    }
    k_row_scan<<<n, 1024, 0, st>>>(rowcount, rowmask, mwords, H, rows > 0 ? psz : H, counts, prefix + n + 1, max_pts);
    k_frame_prefix<<<1, 32, 0, st>>>(counts, prefix, n);
    launches += 2;
    if (rows > 0) {
        if (int_planes) k_emit_refine<true><<<dim3(rows, n), 128, 0, st>>>(map, mpitch, mplane, H, psz, tab, rowmask, mwords, rowcount, kpts, max_pts);
        else k_emit_refine<false><<<dim3(rows, n), 128, 0, st>>>(map, mpitch, mplane, H, psz, tab, rowmask, mwords, rowcount, kpts, max_pts);
        launches++;
    }
    return launches;
}

// the map entries written by k_extrema and the bitmap back to zero (after the survivors' keys were read by k_emit_refine)
int clear_map(cudaStream_t st, unsigned long long* map, int mpitch, long long mplane, int W, int H, unsigned* occ, int n)
{
    const int mwords = (W + 31) / 32;
    k_clear_map<<<dim3((H * mwords + 255) / 256, n), 256, 0, st>>>(map, mpitch, mplane, H, occ, mwords);
    return 1;
}

}  // namespace akzk
