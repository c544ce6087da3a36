/*
 * akaze_b200.h — C ABI of the B200-native AKAZE hot path (detect / describe / match).
 *
 * This is the drop-in boundary.  The C++ surface of the reference (akaze.h / akazed.h:
 * akaze::Akazer, initAkazeData, freeAkazeData, cuMatch and the h* stage functions) is re-created in
 * include/akaze.h ... on top of these entry points; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every "d_" pointer is CUDA device memory on the context's
 *     device, every "h_" pointer is host memory (pinned gives the best transfer rate);
 *   - images are row-major, `pitch` is in ELEMENTS (reference: whp.z, main.cpp:174-188), frame f of a
 *     batch starts at base + f*frame_stride elements;
 *   - every call returns AKZ_OK (0) or a negative AKZ_E_* code; akz_last_error() has the text.
 *     (The reference prints and exit(-1)s instead, cuda_utils.h:18-37; the C++ shim restores that.)
 *   - calls are asynchronous on the context's stream unless they return data to the host;
 *     akz_sync() waits.  There is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with AKZ_E_CUDA.
 */
#ifndef AKAZE_B200_H
#define AKAZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKZ_API __attribute__((visibility("default")))

enum {
    AKZ_OK = 0,
    AKZ_E_INVALID = -1,     /* bad argument */
    AKZ_E_CUDA = -2,        /* CUDA runtime error (text in akz_last_error) */
    AKZ_E_NOMEM = -3,
    AKZ_E_UNSUPPORTED = -4
};

enum { AKZ_F32 = 0, AKZ_U8 = 1 };                /* input pixel type: [0,1] float (main.cpp:149) or raw u8 */
enum { AKZ_MATCH_COMPAT = 0, AKZ_MATCH_KNN2 = 1, AKZ_MATCH_UNIQUE2 = 2 };
enum { AKZ_PLANE_LT = 0, AKZ_PLANE_DET = 1, AKZ_PLANE_LX = 2, AKZ_PLANE_LY = 3 };   /* akaze.cpp:315-320 */

/* The reference has no options struct: these are the 11 arguments of Akazer::init (akaze.h:25-26)
 * and the member defaults (akaze.h:34-54), plus the capacities a batched context needs. */
typedef struct akz_options {
    int   width, height;            /* frame size the context is sized for; 0 x 0 = matcher-only     */
                                    /* context (no pyramid buffers)                                  */
    int   noctaves;                 /* 4   */
    int   max_scale;                /* 4 sublevels per octave, <= 5 (MAX_SCALE, akazed.cu:9)          */
    float per;                      /* 0.7 percentile of the contrast histogram                       */
    float kcontrast;                /* 0.03 — kept for signature parity; recomputed per frame         */
    float soffset;                  /* 1.6  */
    int   reordering;               /* 1    */
    float derivative_factor;        /* 1.5  */
    float dthreshold;               /* 0.001, must be >= 0                                            */
    int   diffusivity;              /* 1 = PM_G2 (akaze_structures.h:51-57)                           */
    int   descriptor_pattern_size;  /* 10   */
    int   max_pts;                  /* per-frame keypoint capacity (main.cpp:157: 10000)              */
    int   max_batch;                /* frames processed together (chunk size)                         */
    int   device;                   /* CUDA device ordinal, -1 = current                              */
    float kcontrast_override;       /* > 0: skip the percentile estimate and use this k (parity hook, */
                                    /*      SURVEY App. B-1)                                          */
    int   fused;                    /* 1 = fused production kernels, 0 = one kernel per reference     */
                                    /*     stage (same results; used as a cross-check)                */
    int   fast_kcontrast_override;  /* > 0: integer pipeline only, use this integer contrast factor   */
    int   lanes;                    /* 1 (default) or 2: with 2, batches larger than max_batch keep two */
                                    /*     chunks in flight on two streams with their own pyramids    */
} akz_options;

/* One detected keypoint (32 bytes, device or host). */
typedef struct akz_keypoint {
    float x, y;          /* refined full-resolution position (akazed.cu:1655-1656)                    */
    float response;      /* det(H) at the integer position (response_map entry)                       */
    float size;          /* octave-relative derivative scale, as the reference stores it              */
    float angle;         /* [0, 2*pi)                                                                  */
    int   layer;         /* octave*max_scale + sublevel  (reference: AkazePoint::octave)              */
    int   ix, iy;        /* full-resolution integer position before refinement                        */
} akz_keypoint;

/* Result of one matcher query (16 bytes).
 * COMPAT: idx1 = match or -1, dist1 = distance or -1 (akazed.cu:2222-2237); idx2 = class mask, dist2 = 0.
 * KNN2  : best and second best (distance, index), lowest index first on ties; -1 when absent.
 * UNIQUE2: as KNN2, but idx1/dist1 = -1 unless dist1 < dist2 and dist1 < 96 (the reference's gMatch rule, akazed.cu:2103). */
typedef struct akz_match_t {
    int idx1, dist1, idx2, dist2;
} akz_match_t;

typedef struct akz_ctx akz_ctx;

/* ---- library ------------------------------------------------------------------------------ */
AKZ_API int         akz_version(void);
AKZ_API const char* akz_last_error(void);
AKZ_API void        akz_default_options(akz_options* o);

/* FED time steps, host only — replaces fed_tau_by_process_time (fed.cpp:41). Returns n (or -n if
 * cap is too small). */
AKZ_API int akz_fed_tau(float T, int M, float tau_max, int reordering, float* tau, int cap);
/* the time steps of one cycle for a given step count and scale — replaces fed_tau_internal (fed.cpp:64) */
AKZ_API int akz_fed_tau_internal(int n, float scale, float tau_max, int reordering, float* tau, int cap);
/* Normalised Gaussian taps k[0..radius] — replaces createGaussKernel (akazed.cu:2298). */
AKZ_API void akz_gauss_taps(float var, int radius, float* taps);
/* M-LDB comparison table (486 pairs) — replaces setCompareIndices (akazed.cu:65). */
AKZ_API void akz_compare_indices(int* idx1, int* idx2);

/* ---- context ------------------------------------------------------------------------------ */
AKZ_API int  akz_create(const akz_options* o, akz_ctx** out);
AKZ_API void akz_destroy(akz_ctx* c);
AKZ_API int  akz_sync(akz_ctx* c);
AKZ_API void* akz_stream(akz_ctx* c);                 /* cudaStream_t */
/* schedule introspection (replaces the scalars computed in akaze.cpp:268-363) */
AKZ_API int  akz_num_levels(const akz_ctx* c);
AKZ_API int  akz_level_info(const akz_ctx* c, int level, int* w, int* h, int* pitch, int* nsteps,
                            float* size, int* sigma_size, float* tau, int tau_cap);
/* device pointer of a plane of frame `frame` of the LAST processed chunk */
AKZ_API const float* akz_level_plane(const akz_ctx* c, int level, int which, int frame);
AKZ_API int  akz_launch_count(const akz_ctx* c);      /* kernels launched since creation */

/* ---- device timing per kernel class (bench.py roofline): while enabled, every kernel group of the pipeline is
 * bracketed by CUDA events on the context's stream; akz_profile_read synchronises, returns the accumulated
 * milliseconds and launch counts per class (AKZ_K_*) and resets them. */
enum { AKZ_K_BASE = 0, AKZ_K_BLUR, AKZ_K_CONTRAST, AKZ_K_PREP, AKZ_K_HESSIAN, AKZ_K_FLOW, AKZ_K_FED, AKZ_K_EXTREMA,
       AKZ_K_NMS, AKZ_K_ORIENT, AKZ_K_DESCRIBE, AKZ_K_MATCH, AKZ_K_MISC, AKZ_NUM_KCLASS };
AKZ_API int         akz_profile_enable(akz_ctx* c, int on);
AKZ_API int         akz_profile_read(akz_ctx* c, int ncls, double* ms, long long* launches);
/* the times of the last akz_profile_read by octave: ms[cls * noct + octave] (scale-space classes; the others report octave 0) */
AKZ_API int         akz_profile_octaves(akz_ctx* c, int ncls, int noct, double* ms);
AKZ_API const char* akz_profile_class_name(int cls);

/* ---- the hot path: replaces Akazer::detectAndCompute (akaze.cpp:101-150) over a batch ------- */
/* d_images: nframes frames of `dtype` on the device.  Results stay on the device:
 *   d_counts[nframes]                      number of keypoints per frame (clamped to max_pts)
 *   d_kpts  [nframes][max_pts]             akz_keypoint, raster order of (iy, ix)
 *   d_desc  [nframes][max_pts][64]         M-LDB, 61 bytes + 3 zero bytes (may be NULL if !describe)
 * nframes may exceed max_batch; frames are processed in chunks. */
AKZ_API int akz_detect_and_compute(akz_ctx* c, const void* d_images, int dtype, int nframes,
                                   int width, int height, int pitch, long long frame_stride,
                                   int describe, int* d_counts, akz_keypoint* d_kpts, uint8_t* d_desc);

/* Same call with HOST buffers: copies the frames in, runs the path, copies counts / keypoints /
 * descriptors out, synchronises.  h_kpts / h_desc are [nframes][max_pts] strided like the device
 * version; only the first counts[f] entries of each frame are valid. */
AKZ_API int akz_detect_and_compute_host(akz_ctx* c, const void* h_images, int dtype, int nframes,
                                        int width, int height, int pitch, long long frame_stride,
                                        int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc);

/* ---- the integer ("fast") pipeline: replaces Akazer::fastDetectAndCompute (akaze.cpp:153-201, :506-743) --------------
 * Raw 8-bit frames in, the reference's 16.16 fixed-point arithmetic (namespace fastakaze, akazed.cu:2781-4366): integer
 * planes (readable through akz_level_plane as int32), integer contrast factor, determinant threshold 65.  Results have
 * the layout of akz_detect_and_compute; `response` holds the integer determinant converted to float. */
AKZ_API int akz_fast_detect_and_compute(akz_ctx* c, const uint8_t* d_images, int nframes, int width, int height, int pitch,
                                        long long frame_stride, int describe, int* d_counts, akz_keypoint* d_kpts, uint8_t* d_desc);
/* host buffers in / out, pipelined like akz_detect_and_compute_host */
AKZ_API int akz_fast_detect_and_compute_host(akz_ctx* c, const uint8_t* h_images, int nframes, int width, int height, int pitch,
                                             long long frame_stride, int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc);
AKZ_API int akz_fast_build_scale_space(akz_ctx* c, const uint8_t* d_images, int nframes, int width, int height, int pitch, long long frame_stride);
AKZ_API int akz_fast_get_kcontrast(akz_ctx* c, int* h_k, int nframes);
/* stage seams of the integer pipeline (akazed.h:88-110), batched like the float ones; tmp: scratch plane batch */
AKZ_API int akz_fast_lowpass(akz_ctx* c, const void* src, int src_is_u8, int* dst, int* tmp, int w, int h, int pitch, long long stride,
                             int nframes, float var, int ksz);                              /* hConv2dR2 / hLowPass */
AKZ_API int akz_fast_down_with_smooth(akz_ctx* c, const int* src, int* dst, int* smooth, int sw, int sh, int sp, long long sstride,
                                      int dw, int dh, int dp, long long dstride, int nframes);
AKZ_API int akz_fast_scharr_contrast(akz_ctx* c, const int* src, int* mag, int* d_k, float per, int w, int h, int pitch, long long stride, int nframes);
AKZ_API int akz_fast_flow(akz_ctx* c, const int* src, int* flow, int type, const int* d_k, int w, int h, int pitch, long long stride, int nframes);
AKZ_API int akz_fast_nld_step(akz_ctx* c, const int* src, const int* flow, int* dst, float tau, int w, int h, int pitch, long long stride, int nframes);
AKZ_API int akz_fast_hessian(akz_ctx* c, const int* smooth, int* lx, int* ly, int* det, int step, int w, int h, int pitch, long long stride, int nframes);

/* Build the scale space only (planes readable through akz_level_plane); nframes <= max_batch. */
AKZ_API int akz_build_scale_space(akz_ctx* c, const void* d_images, int dtype, int nframes,
                                  int width, int height, int pitch, long long frame_stride);
/* contrast factors of the last chunk (device -> host, synchronises) */
AKZ_API int akz_get_kcontrast(akz_ctx* c, float* h_k, int nframes);

/* ---- stage seams: one entry per reference stage function (akazed.h:32-77), batched ----------- */
/* Every plane argument is a batch: frame f at base + f*stride elements. stream = context stream. */
AKZ_API int akz_lowpass(akz_ctx* c, const float* src, float* dst, int w, int h, int pitch, long long stride, int nframes,
                        float var, int ksz);                                       /* hLowPass :2336 */
AKZ_API int akz_down_with_smooth(akz_ctx* c, const float* src, float* dst, float* smooth,
                                 int sw, int sh, int sp, long long sstride,
                                 int dw, int dh, int dp, long long dstride, int nframes);   /* hDownWithSmooth :2389 */
/* writes one contrast factor per frame to d_k; grad may be NULL (the gradient plane is not kept) */
AKZ_API int akz_scharr_contrast(akz_ctx* c, const float* src, float* d_k, float per,
                                int w, int h, int pitch, long long stride, int nframes);   /* hScharrContrast :2410 */
/* d_k: per-frame contrast factor on the device; kscale multiplies it (0.75^octave, akaze.cpp:373) */
AKZ_API int akz_flow(akz_ctx* c, const float* src, float* flow, int type, const float* d_k, float kscale,
                     int w, int h, int pitch, long long stride, int nframes);               /* hFlow :2487 */
AKZ_API int akz_nld_step(akz_ctx* c, const float* src, const float* flow, float* dst, float tau,
                         int w, int h, int pitch, long long stride, int nframes);           /* hNldStep :2509 */
/* n explicit steps with frozen conductance: dst = FED cycle applied to src (src != dst) */
AKZ_API int akz_fed_cycle(akz_ctx* c, const float* src, const float* flow, float* dst, float* tmp,
                          const float* tau, int n, int w, int h, int pitch, long long stride, int nframes);
AKZ_API int akz_hessian(akz_ctx* c, const float* smooth, float* lx, float* ly, float* det, int step,
                        int w, int h, int pitch, long long stride, int nframes);            /* hHessianDeterminant :2531 */

/* keypoint stages on the planes of the LAST processed chunk (akz_build_scale_space / akz_detect_and_compute):
 * d_counts[nframes], d_kpts[nframes][max_pts] as produced by akz_detect_and_compute (or supplied by a test) */
AKZ_API int akz_orient(akz_ctx* c, const int* d_counts, akz_keypoint* d_kpts, int nframes);                 /* hCalcOrient :2655 */
AKZ_API int akz_describe(akz_ctx* c, const int* d_counts, const akz_keypoint* d_kpts, uint8_t* d_desc, int nframes); /* hDescribe :2675 */
/* extrema + NMS + refinement on the planes of the last processed chunk                     hCalcExtremaMap/hNmsR/hRefine */
AKZ_API int akz_detect_keypoints(akz_ctx* c, int nframes, int* d_counts, akz_keypoint* d_kpts);

/* ---- matcher: replaces akaze::cuMatch / hMatch (akaze.cpp:55, akazed.cu:2758) -------------------- */
/* d_q [nq][64], d_t [nt][64] descriptors; train indices reported as t_index_base + local index
 * (for sharded train sets).  d_out[nq].  finalize != 0 applies the acceptance rule (single shard);
 * finalize == 0 leaves the associative partial form for akz_match_merge after a gather. */
AKZ_API int akz_match(akz_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt,
                      int t_index_base, int mode, int finalize, akz_match_t* d_out);
/* merge `nparts` partial results (each [nq], e.g. gathered from shards) into d_out[nq];
 * finalize != 0 applies the acceptance rule of the mode (COMPAT: unique class and < 96) */
AKZ_API int akz_match_merge(akz_ctx* c, const akz_match_t* d_parts, int nparts, int nq, int mode,
                            int finalize, akz_match_t* d_out);
/* Consecutive-frame matching of a batch (BASELINE configs[4], the 4K stream): the descriptors of frame f (queries) against
 * those of frame f - 1 (train) for f = 1 .. nframes-1, in one batched launch; the keypoint counts are read on the device, so the
 * call can follow akz_detect_and_compute on the same stream without a host round trip.  d_desc [nframes][max_pts][64] and
 * d_counts [nframes] as akz_detect_and_compute wrote them (max_pts = the context's); d_out [nframes][max_pts], row 0 unused. */
AKZ_API int akz_match_pairs(akz_ctx* c, const uint8_t* d_desc, const int* d_counts, int nframes, int mode, akz_match_t* d_out);

/* ---- train-sharded matching across GPUs (one process per GPU; SURVEY 8e, BASELINE configs[3]) -------------------------------
 * Every rank holds all nq queries and a contiguous range of the train set starting at global index t_index_base.
 * akz_match_sharded = akz_match(finalize = 0) on the local range -> ONE ncclAllGather of nq x 16 bytes per rank on the context's
 * stream (NCCL over NVLink / NVSwitch) -> akz_match_merge(finalize = 1): every rank ends with the same final result in d_out.
 * Nothing synchronises with the host; the only exchange is the gather of the per-shard candidates.
 * The communicator belongs to the context.  akz_comm_unique_id fills a 128-byte id on one rank (ncclGetUniqueId); the caller
 * hands it to the other ranks by its own means (MPI, torch.distributed, a file ...) and every rank calls akz_comm_init.
 * akz_comm_attach uses a communicator the caller already owns (an ncclComm_t of the NCCL library loaded in the process).
 * NCCL is loaded at run time (libnccl.so.2, the copy already in the process if there is one): the library has no link-time
 * dependency on it and everything else works without it. */
#define AKZ_COMM_ID_BYTES 128
AKZ_API int akz_comm_unique_id(void* id128);
AKZ_API int akz_comm_init(akz_ctx* c, int nranks, int rank, const void* id128);
AKZ_API int akz_comm_attach(akz_ctx* c, void* nccl_comm, int nranks, int rank);
AKZ_API int akz_comm_destroy(akz_ctx* c);
AKZ_API int akz_match_sharded(akz_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_t_local, int nt_local,
                              int t_index_base, int mode, akz_match_t* d_out);

/* kernel selection of akz_match: 0 = by problem size (default), 1 = LOP3/POPC kernel, 2 = mma.sync (IMMA) kernel,
 * 3 = tcgen05 kernel (tensor-memory accumulators).
 * Both produce identical results; the switch exists for tests and measurements. */
AKZ_API void akz_set_match_kernel(int which);
/* kernel selection of the M-LDB stage: 0 = by pattern size (default: k_describe_b for the reference's pattern 10, the generic
 * k_describe_s otherwise), 1 = always the generic kernel.  Identical results; for tests and measurements. */
AKZ_API void akz_set_describe_kernel(int which);
/* Host-side planning, exposed for tests (no device work).
 * akz_plan_chunks: the chunk plan of a batch of nframes frames (max_batch per chunk, ramp_up = the host pipeline's short first
 *   chunks); writes starts / sizes (up to cap entries) and returns the number of chunks.
 * akz_plan_match: work decomposition of the tcgen05 matcher for nq x nt descriptors: out = { partial results per query,
 *   items per query block, tiles of 128 train descriptors per item, tiles per CTA, CTAs }; returns 0 or an error code. */
AKZ_API int akz_plan_chunks(int nframes, int max_batch, int ramp_up, int* starts, int* sizes, int cap);
AKZ_API int akz_plan_match(int nq, int nt, int* out5);
AKZ_API int akz_match_host(akz_ctx* c, const uint8_t* h_q, int nq, const uint8_t* h_t, int nt,
                           int mode, akz_match_t* h_out);

/* ---- host format in OpenCV conventions (beside cv::AKAZE; reference main.cpp:373-388 uses cv::KeyPoint / cv::DMatch) -- */
/* out: n rows of 7 floats = pt.x, pt.y, size (full-resolution pixels: size * 2^octave), angle (degrees), response,
 * octave, class_id (sublevel).  Host arrays. */
AKZ_API int akz_keypoints_to_opencv(const akz_keypoint* h_kpts, int n, int max_scale, float* out);
/* out: rows (queryIdx, trainIdx, distance) of the accepted matches; returns their number.  Host arrays. */
AKZ_API int akz_matches_to_opencv(const akz_match_t* h_m, int nq, int* out);

/* ---- AoS bridge for the akaze.h shim ------------------------------------------------------------ */
/* writes x,y,octave,size,angle,features into reference-layout AkazePoint records (104 B, App. D) */
AKZ_API int akz_pack_points(akz_ctx* c, const int* d_count, const akz_keypoint* d_kpts, const uint8_t* d_desc,
                            void* d_points, int max_pts, int with_desc);
/* gathers descriptors out of AkazePoint records into [n][64] */
AKZ_API int akz_unpack_desc(akz_ctx* c, const void* d_points, int n, uint8_t* d_desc);
/* scatters match results back into AkazePoint::match/distance/match_x/match_y (akazed.cu:2222-2237) */
AKZ_API int akz_scatter_matches(akz_ctx* c, const akz_match_t* d_m, int nq, void* d_points_q, const void* d_points_t);

#ifdef __cplusplus
}
#endif
#endif /* AKAZE_B200_H */
