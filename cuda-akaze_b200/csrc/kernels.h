// Internal launch-wrapper declarations (host side).  Every wrapper enqueues on the given stream,
// never synchronises, and returns the number of kernels it launched (or a negative AKZ_E_* code).
#pragma once
#include <cuda_runtime.h>
#include "../../include/akaze_b200.h"
#include "common.cuh"

struct AkzLevelTable {
    AkzLevelDev lv[AKZ_MAX_LEVELS];
    int nlevels, max_scale;
};

struct AkzExtremaLevel {
    const float* det;
    const unsigned char* hot;           // optional (k_deriv4): one byte per four pixels, set when a determinant of the group passes the threshold
    long long plane;
    float border, threshold;
    int layer, ithreshold;              // ithreshold: integer pipeline (akaze.cpp:560: 65)
};
struct AkzExtremaArgs {
    AkzExtremaLevel lv[8];
    int nsub, w, h, pitch, octave, psz, int_planes;
};

namespace akzk {

int radius_from_ksz(int ksz);
void hessian_factors(float* fac1, float* fac2);

// ---- scale_space.cu: one kernel per reference stage --------------------------------------------
int lowpass(cudaStream_t st, const float* src, float* dst, int w, int h, int sp, long long sstride, int dp, long long dstride,
            int n, float var, int ksz);
int lowpass_u8(cudaStream_t st, const unsigned char* src, float* dst, int w, int h, int sp, long long sstride, int dp, long long dstride,
               int n, float var, int ksz);
int down_with_smooth(cudaStream_t st, const float* src, float* dst, float* smooth, int sw, int sh, int sp, long long sstride,
                     int dw, int dh, int dp, long long dstride, int n);
int contrast(cudaStream_t st, const float* src, unsigned* hmax_bits, int* hist, float* kout, float per, float override_k,
             int w, int h, int pitch, long long stride, int n);
int flow(cudaStream_t st, const float* src, float* flowp, int type, const float* kc, float kscale, int nmul,
         int w, int h, int pitch, long long stride, int n);
int nld_step(cudaStream_t st, const float* src, const float* flowp, float* dst, float tau, int w, int h, int pitch, long long stride, int n);
int hessian(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long stride, int n);

// ---- fused production kernels ---------------------------------------------------------------------
// base_level.cu: Lt(0,0), gradient-magnitude plane, its maximum and histogram in one pass over the input (+ one over the
// magnitude plane).  Returns the number of launches, 0 when the sigma0 radius is not 4 (caller uses the staged kernels).
int base_level2(cudaStream_t st, const void* img, int dtype, int w, int h, int ipitch, long long istride,
                float* lt, int pitch, long long plane, float* mag, unsigned* hmax_bits, int* hist, float var0, int ksz0, int n, int int_planes = 0);
int contrast_scan(cudaStream_t st, const unsigned* hmax_bits, const int* hist, float* kout, float per, float override_k, int w, int h, int n);
// level_prep.cu: second-generation level kernel (templated on the derivative step, 64x64 tiles, vector shared-memory
// traffic).  mode 0 = base level (no blur), 1 = blur, 2 = octave transition.  Returns 1 if launched, 0 if the
// (step, size) combination is not covered and the caller must run the per-stage kernels.
int level_prep2(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                float* ltdst, float* flowp, float* lx, float* ly, float* det, int type,
                const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n, int int_planes = 0);
// split level pipeline: k_prep3 (level_prep.cu: blur or octave transition + conductance, blurred plane to global memory) and
// k_deriv4 + border-ring kernels (deriv_stream.cu: Lx, Ly, det from the blurred plane).  Both return 0 when the case is not covered.
int level_blur_flow(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                    float* ltdst, float* flowp, float* smooth, int type, const float* kc, float kscale, int nmul,
                    int w, int h, int pitch, long long plane, int n, int int_planes = 0);
// blur_stream.cu: the same half as a streaming warp kernel (k_blur4); 0 when not covered
int blur_stream(cudaStream_t st, int mode, const float* src, int sw, int sh, int sp, long long splane,
                float* ltdst, float* flowp, float* smooth, int type, const float* kc, float kscale, int nmul,
                int w, int h, int pitch, long long plane, int n, int int_planes = 0);
int deriv_stream(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long plane,
                 int n, int int_planes = 0, cudaStream_t ring_st = nullptr, cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr,
                 unsigned char* hot = nullptr, float thr = 0.f, int ithr = 0);     // hot: one byte per 4 pixels, "a determinant of the group > threshold"
int deriv_rings(cudaStream_t st, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long plane,
                int n, int int_planes, cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join);
// level_stream.cu: both halves of a same-resolution level in one streaming kernel (k_level4) + the border rings; the blurred plane
// is written only near the image border.  0 when not covered (few CTAs unless force, derivative step outside 2..4, alignment)
int level_stream(cudaStream_t st, const float* src, float* flowp, float* smooth, float* lx, float* ly, float* det, int type,
                 const float* kc, float kscale, int nmul, int step, int w, int h, int pitch, long long plane, int n, int int_planes,
                 cudaStream_t ring_st, cudaEvent_t ev_fork, cudaEvent_t ev_join, unsigned char* hot, float thr, int ithr, int force);
// fed.cu: all n FED steps of a level (frozen conductance) in ceil(n / 4) launches of the streaming warp kernel (k_fed4), or of the
// tile kernel (k_fed3) when the rows are not 16-byte aligned
int fed_cycle(cudaStream_t st, const float* src, const float* flowp, float* dst, float* tmp, const float* tau, int nsteps,
              int w, int h, int pitch, long long plane, int n, int fused, int int_planes = 0);

// ---- fast_pipeline.cu: integer ("fast") scale-space stages, reference namespace fastakaze (akazed.cu:2781-4366) ----------
int fast_lowpass(cudaStream_t st, const void* src, int src_u8, int* dst, int* tmp, int w, int h, int sp, long long sstride,
                 int dp, long long dstride, int n, float var, int ksz);
int fast_down(cudaStream_t st, const int* src, int* dst, int* smooth, int sw, int sh, int sp, long long sstride,
              int dw, int dh, int dp, long long dstride, int n);
// fast_contrast split for the fused integer base level: zero the maximum / histogram; histogram of a magnitude plane + scan
int fast_contrast_init(cudaStream_t st, int* hmax, int* hist, int n);
int fast_contrast_tail(cudaStream_t st, const int* mag, const int* hmax, int* hist, int* kout, float per, int override_k,
                       int w, int h, int pitch, long long stride, int n);
int fast_contrast(cudaStream_t st, const int* src, int* mag, int* hmax, int* hist, int* kout, float per, int override_k,
                  int w, int h, int pitch, long long stride, int n);
int fast_flow(cudaStream_t st, const int* src, int* flow, int type, const int* kc, int nmul, int w, int h, int pitch, long long stride, int n);
int fast_nld_step(cudaStream_t st, const int* src, const int* flow, int* dst, float tau, int w, int h, int pitch, long long stride, int n);
int fast_hessian(cudaStream_t st, const int* smooth, int* lx, int* ly, int* det, int step, int w, int h, int pitch, long long stride, int n);

// ---- detect.cu ---------------------------------------------------------------------------------------
int extrema(cudaStream_t st, const AkzExtremaArgs& a, unsigned long long* map, int mpitch, long long mplane, unsigned* occ, int mwords, int H, int n);
int nms_emit(cudaStream_t st, unsigned long long* map, int mpitch, long long mplane, int W, int H, int psz,
             const AkzLevelTable& tab, unsigned* occ, unsigned* rowmask, int* rowcount, int* counts, int* prefix,
             akz_keypoint* kpts, int max_pts, int n, int int_planes = 0);

int clear_map(cudaStream_t st, unsigned long long* map, int mpitch, long long mplane, int W, int H, unsigned* occ, int n);
int frame_prefix(cudaStream_t st, const int* counts, int* prefix, int n);

// ---- describe.cu -------------------------------------------------------------------------------------
int orient_table_init(cudaStream_t st);
int layer_order(cudaStream_t st, const int* counts, const akz_keypoint* kpts, int* order, int max_pts, int n);     // order[frame][j]: keypoints grouped by layer
int orient(cudaStream_t st, const AkzLevelTable& tab, const int* counts, const int* prefix, akz_keypoint* kpts, int max_pts, int n, int fast = 0, const int* order = nullptr);
int describe(cudaStream_t st, const AkzLevelTable& tab, const int* counts, const int* prefix, const akz_keypoint* kpts,
             unsigned char* desc, int max_pts, int n, int pattern, int fast = 0, const int* order = nullptr);
void set_describe_kernel(int which);    // 1 = the generic kernel for every pattern size
int describe_prepare(int pattern);      // builds the pattern's reduction tables on the current device (call outside stream capture)
int pack_points(cudaStream_t st, const int* count, const akz_keypoint* kpts, const unsigned char* desc, void* points, int max_pts, int with_desc);
int unpack_desc(cudaStream_t st, const void* points, int n, unsigned char* desc);
int scatter_matches(cudaStream_t st, const akz_match_t* m, int nq, void* pq, const void* pt);

// ---- match.cu ----------------------------------------------------------------------------------------
int match_partial(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode,
                  int nsplit, akz_match_t* parts, int use_mma = 0);
int match_pairs(cudaStream_t st, const unsigned char* desc, const int* counts, int nframes, int max_pts, int mode, int nsplit,
                akz_match_t* parts, akz_match_t* out);
// match_tc5.cu: tcgen05 / tensor-memory matcher (same partial-result contract)
int match_tc5_plan(int nq, int nt, int* nsplit, int* tiles, int* per_cta, int* grid);     // returns the number of partial results per query
int match_partial_tc5(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode, akz_match_t* parts,
                      int filter_mode = -1);     // -1 = chunk filter by range length (default), 0 / 1 = forced off / on (tests)
int match_merge(cudaStream_t st, const akz_match_t* parts, int nparts, int nq, int mode, int finalize, akz_match_t* out);

}  // namespace akzk
