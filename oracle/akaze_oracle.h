/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference AKAZE hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  It is the checker, never the product: nothing under cuda-akaze_b200/ links,
 * imports or calls it.
 *
 * Every function restates one reference function (Accustomer/CUDA-AKAZE, file:line given at each
 * definition in akaze_oracle.c) in plain C with the floating-point contraction pattern of the
 * reference's sm_100a build written out explicitly (fmaf / separately rounded products), so that
 * the scale space, detector response, extrema, NMS and refinement are bit-comparable with the
 * GPU.  Deliberate differences (each is a reference bug with undefined behaviour):
 *   B-1  contrast maximum = the true maximum (the reference's reduction is racy)
 *   B-2  extrema merge    = arg-max, ties to the lowest layer (the reference races)
 *   B-3  histogram        = in-image pixels only
 *   B-4  orientation      = histogram bins summed in ascending sample order
 *   B-5  keypoint order   = raster order of the full-resolution integer position
 *   B-6  Hamming distance = 486 bits, padding bytes zero
 *   B-7  blur halos       = correct reflect-101 everywhere
 * expf/atan2f/cosf/sinf come from libm here and from CUDA on the GPU (MUFU for __cosf/__sinf);
 * the orientation and descriptor stages are therefore pinned against the compiled reference on
 * the GPU box (oracle/_ref), and against this file only with the trig values passed in.
 *
 * Pinning status: scale-space / detector functions are pinned bit-for-bit against oracle/_ref on a
 * B200 by tests/test_gpu_parity.py (the reference ships no golden vectors of its own).
 */
#ifndef AKAZE_ORACLE_H
#define AKAZE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_options {
    int   noctaves;            /* akaze.h:36  default 4   */
    int   max_scale;           /* akaze.h:38  default 4   */
    float per;                 /* akaze.h:40  default 0.7 */
    float soffset;             /* akaze.h:44  default 1.6 */
    int   reordering;          /* akaze.h:46  default 1   */
    float derivative_factor;   /* akaze.h:48  default 1.5 */
    float dthreshold;          /* akaze.h:50  default 1e-3*/
    int   diffusivity;         /* akaze.h:52  default 1 (PM_G2) */
    int   pattern;             /* akaze.h:54  default 10  */
    float kcontrast_override;  /* >0: use instead of the percentile estimate (parity hook)    */
    int   threads;             /* OpenMP threads, <=0: all                                    */
} orc_options;

typedef struct orc_keypoint {
    float x, y;                /* full-resolution, refined                                    */
    float response;            /* det(H) at the integer position (the reference never writes  */
                               /* AkazePoint::response; this is the response_map value)       */
    float size;                /* octave-relative derivative scale (akaze.cpp:341,361)        */
    float angle;
    int   layer;               /* octave*max_scale + sublevel (akazed.cu:1608)                */
    int   ix, iy;              /* full-resolution integer position before refinement          */
    unsigned char desc[64];    /* 61 bytes M-LDB + 3 zero bytes                               */
} orc_keypoint;

typedef struct orc_pyramid orc_pyramid;

void  orc_default_options(orc_options* o);

/* fed.cpp:41-119 */
int   orc_fed_tau(float T, int M, float tau_max, int reordering, float* tau, int cap);
/* akazed.cu:2298-2333 */
void  orc_gauss_taps(float var, int radius, float* taps);
/* main.cpp:149 (cv::Mat::convertTo CV_8U -> CV_32F, alpha = 1/255) */
void  orc_u8_to_f32(const unsigned char* src, float* dst, int w, int h, int sp, int dp);

/* stage functions: planes are row-major float, pitch in elements */
void  orc_lowpass(const float* src, float* dst, int w, int h, int p, float var, int ksz);
void  orc_down_with_smooth(const float* src, float* dst, float* smooth, int sw, int sh, int sp, int dw, int dh, int dp);
void  orc_scharr_mag(const float* src, float* mag, int w, int h, int p);
float orc_contrast_from_mag(const float* mag, int w, int h, int p, float per, float* hmax_out);
void  orc_flow(const float* src, float* flow, int type, float kcontrast, int w, int h, int p);
void  orc_nld_step(const float* src, const float* flow, float* dst, float tau, int w, int h, int p);
void  orc_hessian(const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int p);

/* whole pipeline */
orc_pyramid* orc_pyramid_create(int w, int h, const orc_options* o);
void  orc_pyramid_free(orc_pyramid* P);
int   orc_pyramid_levels(const orc_pyramid* P);
void  orc_pyramid_level_dims(const orc_pyramid* P, int level, int* w, int* h, int* pitch, int* nsteps, float* size, int* sigma_size);
int   orc_pyramid_tau(const orc_pyramid* P, int level, float* tau, int cap);
/* which: 0 Lt, 1 det, 2 Lx, 3 Ly (the reference's plane groups, akaze.cpp:315-320) */
const float* orc_pyramid_plane(const orc_pyramid* P, int level, int which);
float orc_pyramid_kcontrast(const orc_pyramid* P);
void  orc_pyramid_build(orc_pyramid* P, const float* img, int pitch);
int   orc_pyramid_detect(orc_pyramid* P, orc_keypoint* out, int cap);
/* trig: optional [n][2] array of (cos, sin) to use instead of libm (GPU MUFU values) */
void  orc_pyramid_describe(const orc_pyramid* P, orc_keypoint* kps, int n, int with_orientation, const float* trig);

int   orc_detect_and_compute(const float* img, int w, int h, int pitch, const orc_options* o,
                             orc_keypoint* out, int cap, int describe);

/* matching over [n][64]-byte descriptors (bytes 61..63 zero).
 * compat (akazed.cu:2144-2241): out[q] = {match or -1, distance or -1}
 * knn2   (akazed.cu:2028-2122 intent, OpenCV knnMatch ties): out[q] = {i1,d1,i2,d2} */
void  orc_match_compat(const unsigned char* q, int nq, const unsigned char* t, int nt, int* out2);
void  orc_match_knn2(const unsigned char* q, int nq, const unsigned char* t, int nt, int* out4);

/* akazed.cu:65-159 */
void  orc_compare_indices(int* idx1, int* idx2);

#ifdef __cplusplus
}
#endif
#endif
