// C ABI implementation: context, scale schedule, the batched detect+describe pipeline and the
// matcher front end.  Replaces Akazer::{init,allocMemory,detect,detectAndCompute} (akaze.cpp:80-503)
// and akaze::cuMatch (akaze.cpp:55-64); the per-frame host round trips of the reference (contrast
// maximum, histogram, keypoint counter) are gone: every data-dependent scalar stays on the device.
#include "common.cuh"
#include "kernels.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <dlfcn.h>

struct akz_nccl_id { char internal[128]; };      // layout of ncclUniqueId (nccl.h): passed by value to ncclCommInitRank

// ---- errors ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int akz_set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int akz_set_cuda_error(cudaError_t e, const char* what, const char* file, int line)
{
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return AKZ_E_CUDA;
}

extern "C" {

int akz_version(void) { return 100; }
const char* akz_last_error(void) { return g_err; }

void akz_default_options(akz_options* o)
{
    memset(o, 0, sizeof(*o));
    o->width = 0; o->height = 0;
    o->noctaves = 4; o->max_scale = 4; o->per = 0.7f; o->kcontrast = 0.03f; o->soffset = 1.6f; o->reordering = 1;
    o->derivative_factor = 1.5f; o->dthreshold = 0.001f; o->diffusivity = 1; o->descriptor_pattern_size = 10;
    o->max_pts = 10000; o->max_batch = 8; o->device = -1; o->kcontrast_override = 0.f; o->fused = 1; o->fast_kcontrast_override = 0; o->lanes = 1;
}

// ---- host math ----------------------------------------------------------------------------------------
static bool is_prime_int(int v)
{
    if (v <= 1) return false;
    if (v == 2 || v == 3 || v == 5 || v == 7) return true;
    if (!(v % 2) || !(v % 3) || !(v % 5) || !(v % 7)) return false;
    int upper = (int)sqrt(v + 1.0);
    for (int d = 11; d <= upper; d += 2)
        if (v % d == 0) return false;
    return true;
}

// FED cycle time steps (Grewenig/Weickert/Bruhn): the smallest n with a stable cycle of stopping time
// T/M, tau_k = d / cos^2(pi (2k+1)/(4n+2)), optionally permuted by the kappa-cycle (kappa = n/2) modulo
// the next prime >= n+1.  The float/double mix mirrors fed.cpp:48-76 so the values are bit-identical.
int akz_fed_tau(float T, int M, float tau_max, int reordering, float* tau, int cap)
{
    const float t = T / (float)M;
    const int n = (int)(ceil(sqrt(3.0 * t / tau_max + 0.25f) - 0.5f - 1.0e-8f) + 0.5f);
    if (n <= 0) return 0;
    const float scale = (float)(3.0 * t / (tau_max * (float)(n * (n + 1))));
    return akz_fed_tau_internal(n, scale, tau_max, reordering, tau, cap);
}

// the n time steps of one cycle for a given (n, scale): replaces fed_tau_internal (fed.cpp:64-119)
int akz_fed_tau_internal(int n, float scale, float tau_max, int reordering, float* tau, int cap)
{
    if (n <= 0) return 0;
    if (n > cap) return -n;
    const float c = 1.0f / (4.0f * (float)n + 2.0f);
    const float d = scale * tau_max / 2.0f;
    std::vector<float> plain(n);
    for (int k = 0; k < n; ++k) {
        const float hc = (float)cos(M_PI * (2.0f * (float)k + 1.0f) * c);
        plain[k] = d / (hc * hc);
    }
    if (!reordering) {
        std::copy(plain.begin(), plain.end(), tau);
        return n;
    }
    const int kappa = n / 2;
    int prime = n + 1;
    while (!is_prime_int(prime)) prime++;
    for (int k = 0, l = 0; l < n; ++k, ++l) {
        int index;
        while ((index = ((k + 1) * kappa) % prime - 1) >= n) k++;
        tau[l] = plain[index];
    }
    return n;
}

void akz_gauss_taps(float var, int radius, float* k)
{
    const float denom = 1.f / (2.f * var);
    float ksum = 0.f;
    for (int i = 0; i <= radius; i++) {
        k[i] = expf(-i * i * denom);
        ksum += (i == 0) ? k[i] : k[i] + k[i];
    }
    ksum = 1 / ksum;
    for (int i = 0; i <= radius; i++) k[i] *= ksum;
}

void akz_compare_indices(int* c1, int* c2)
{
    // three grids (2x2: cells 0..3, 3x3: 4..12, 4x4: 13..28) x three channels, all ordered pairs (j<i);
    // a value lives at 3*cell + channel
    static const int lo[3] = { 0, 4, 13 }, hi[3] = { 4, 13, 29 };
    int n = 0;
    for (int g = 0; g < 3; g++)
        for (int ch = 0; ch < 3; ch++)
            for (int j = lo[g]; j + 1 < hi[g]; j++)
                for (int i = j + 1; i < hi[g]; i++) { c1[n] = 3 * j + ch; c2[n] = 3 * i + ch; n++; }
}

}  // extern "C"

// ---- context -----------------------------------------------------------------------------------------
constexpr int AKZ_NSET = 4;                 // staging buffers / result sets of the host pipeline

struct akz_ctx {
    akz_options opt;
    int device;
    cudaStream_t stream;
    int nlev, noct, psz;
    AkzLevel lev[AKZ_MAX_LEVELS];
    float tau[AKZ_MAX_LEVELS * AKZ_MAX_STEPS];
    int mpitch;
    long long mplane;
    // device memory
    std::vector<void*> allocs;
    float *smooth, *flow, *tmpA, *tmpB;            // scratch planes of the octave whose launches are being issued (= sc[k])
    // Small-batch contexts (max_batch <= 4, e.g. the one behind akaze::Akazer): every octave has its own scratch planes and
    // stream, so the octave chains -- octave o + 1 only needs level (o, 0) -- run side by side, and a whole chunk is captured
    // once per argument set as a CUDA graph and replayed (a single frame is launch bound: ~90 launches of a few microseconds)
    struct Scratch { float *smooth, *flow, *tmpA, *tmpB; } sc[8];
    cudaStream_t cur;                               // stream the pipeline launches go to (c->stream, or an octave stream)
    cudaStream_t ostream[8];
    cudaEvent_t ev_lvl0[8], ev_oct[8];
    // side streams of the border-ring kernels (deriv_stream.cu), one per octave stream; rk = the one in use, ring_pending[k]:
    // ev_ring_join[k] has to be waited for before the blurred plane is overwritten / the level's planes are read
    cudaStream_t ring_stream[8];
    cudaEvent_t ev_ring_fork[8], ev_ring_join[8];
    bool ring_pending[8];
    int rk;
    bool opar, opar_saved;
    struct GraphEntry { unsigned long long key[10]; int seen; cudaGraphExec_t exec; int launches; };
    std::vector<GraphEntry> graphs;
    bool graph_ok;
    unsigned long long* map;
    unsigned char* hot[AKZ_MAX_LEVELS];   // per level: one byte per four pixels written by k_deriv4 ("a determinant of the group > threshold")
    bool hot_valid[AKZ_MAX_LEVELS];       // ... of the chunk in flight (set by the level's derivative kernel, used by detect_chunk right after)
    unsigned *rowmask, *occ;      // survivor masks; occupancy bitmap of the key map (one bit per pixel with a candidate)
    int *rowcount, *prefix, *order, *hist, *counts_own;      // counts_own / kpts_own / desc_own: AKZ_NSET result sets (host API pipeline)
    unsigned* hmax;
    float* kc;
    akz_keypoint* kpts_own;
    unsigned char* desc_own;
    void* img_stage[AKZ_NSET];
    size_t img_stage_bytes;
    cudaStream_t h2d_stream, d2h_stream;
    cudaEvent_t ev_h2d[AKZ_NSET], ev_comp[AKZ_NSET], ev_cnt[AKZ_NSET], ev_d2h[AKZ_NSET];
    int* h_cnt_pinned;                               // [AKZ_NSET][max_batch] pinned landing zone of the per-chunk counts
    // second lane (akz_options::lanes == 2): a child context with its own pyramid, scratch planes and stream; chunks
    // alternate between the lanes so that two chunks are in flight (kernels bound by different pipes overlap)
    akz_ctx* lane1;
    cudaEvent_t ev_fork, ev_join;
    akz_match_t* match_parts;
    size_t match_parts_n;
    void* match_stage; size_t match_stage_bytes;
    // train-sharded matching: communicator (NCCL, loaded at run time) and the gather buffers
    void* comm; int comm_ranks, comm_rank; bool comm_owned;
    akz_match_t* shard_buf; size_t shard_buf_n;          // [1 + nranks][nq]: own partial result, then the gathered ones
    int last_frames;
    long long launches;
    AkzLevelTable tab;
    // per-kernel-class device timing (akz_profile_*): event pairs around every wrapper call while enabled
    bool prof_on;
    struct ProfPair { cudaEvent_t a, b; int cls, launches, oct; };
    std::vector<ProfPair> prof_pairs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[AKZ_NUM_KCLASS];
    long long prof_launches[AKZ_NUM_KCLASS];
    double prof_oct_ms[AKZ_NUM_KCLASS][8], prof_oct_last[AKZ_NUM_KCLASS][8];      // the same times by octave (keypoint stages: octave 0)
    int cur_oct = 0;                               // octave whose launches are being issued
};

static cudaEvent_t prof_event(akz_ctx* c)
{
    cudaEvent_t e;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}

template <typename T>
static int dalloc(akz_ctx* c, T** p, size_t n)
{
    void* q = nullptr;
    AKZ_CUDA_TRY(cudaMalloc(&q, n * sizeof(T)));
    c->allocs.push_back(q);
    *p = (T*)q;
    return AKZ_OK;
}

static int align_up(int a, int b) { return (a % b) ? a - a % b + b : a; }

// scale schedule: akaze.cpp:204-237 (octave sizes) and :268-363 (sigma, evolution times, FED steps,
// derivative scales, borders)
static int build_schedule(akz_ctx* c)
{
    const akz_options& o = c->opt;
    int ow[16], oh[16];
    ow[0] = o.width; oh[0] = o.height; c->noct = 1;
    for (int j = 1; j < o.noctaves; j++) {
        ow[j] = ow[j - 1] >> 1; oh[j] = oh[j - 1] >> 1;
        if (ow[j] < 80 || oh[j] < 80) break;
        c->noct = j + 1;
    }
    const int S = o.max_scale;
    if (c->noct * S > AKZ_MAX_LEVELS) return akz_set_error(AKZ_E_INVALID, "too many levels");
    float last_etime = (float)(0.5 * o.soffset * o.soffset);
    const float smax = (float)(10.0 * sqrtf(2.0f));
    float psz = 10000.f;
    int oratio = 1, toff = 0;
    c->nlev = 0;
    for (int i = 0; i < c->noct; i++) {
        for (int j = 0; j < S; j++) {
            AkzLevel& L = c->lev[c->nlev++];
            memset(&L, 0, sizeof(L));
            L.octave = i; L.sub = j; L.w = ow[i]; L.h = oh[i]; L.pitch = align_up(ow[i], 32);
            L.plane = (long long)L.pitch * L.h;
            L.tau_off = toff;
            if (i == 0 && j == 0) {
                L.esigma = o.soffset;
                L.size = o.soffset * o.derivative_factor;
                L.nsteps = 0;
            } else {
                L.esigma = o.soffset * powf(2, (float)j / S + i);
                const float cur = 0.5f * L.esigma * L.esigma;
                const float ttime = cur - last_etime;
                int n = akz_fed_tau(ttime, 1, 0.25f, o.reordering, c->tau + toff, AKZ_MAX_STEPS);
                if (n < 0) return akz_set_error(AKZ_E_UNSUPPORTED, "FED cycle of %d steps exceeds %d", -n, AKZ_MAX_STEPS);
                L.nsteps = n; toff += n;
                L.size = L.esigma * o.derivative_factor / oratio;
                last_etime = cur;
            }
            L.sigma_size = (int)(L.size + 0.5f);
            L.border = smax * L.sigma_size;
        }
        psz = std::min(psz, c->lev[i * S].border * oratio);
        oratio *= 2;
    }
    c->psz = (int)psz;
    return AKZ_OK;
}

extern "C" {

int akz_create(const akz_options* o, akz_ctx** out)
{
    if (!o || !out) return akz_set_error(AKZ_E_INVALID, "null argument");
    const bool matcher_only = (o->width == 0 && o->height == 0);
    if (!matcher_only && (o->width < 16 || o->height < 16 || o->width > 4096)) return akz_set_error(AKZ_E_INVALID, "frame size %dx%d unsupported", o->width, o->height);
    if (o->max_scale < 1 || o->max_scale > 5 || o->noctaves < 1 || o->noctaves > 8) return akz_set_error(AKZ_E_INVALID, "octaves/sublevels out of range");
    if (o->dthreshold < 0.f) return akz_set_error(AKZ_E_INVALID, "dthreshold must be >= 0");
    if (o->max_pts < 1 || o->max_batch < 1) return akz_set_error(AKZ_E_INVALID, "max_pts / max_batch must be positive");
    if (!matcher_only && o->height > 4096) return akz_set_error(AKZ_E_INVALID, "frame size %dx%d unsupported", o->width, o->height);
    if (o->lanes < 0 || o->lanes > 2) return akz_set_error(AKZ_E_INVALID, "lanes must be 1 or 2");
    if (o->diffusivity < 0 || o->diffusivity > 3) return akz_set_error(AKZ_E_INVALID, "diffusivity must be 0..3 (akaze_structures.h:51-57)");
    if (o->descriptor_pattern_size < 1 || o->descriptor_pattern_size > 64) return akz_set_error(AKZ_E_INVALID, "descriptor_pattern_size out of range");
    if (!(o->per > 0.f && o->per <= 1.f) || !(o->soffset > 0.f) || !(o->derivative_factor > 0.f)) return akz_set_error(AKZ_E_INVALID, "per / soffset / derivative_factor out of range");
    int ndev = 0;
    AKZ_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return akz_set_error(AKZ_E_CUDA, "no CUDA device: this library has no CPU fallback");
    akz_ctx* c = new akz_ctx();
    c->opt = *o;
    c->launches = 0; c->last_frames = 0; c->prof_on = false;
    memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_launches, 0, sizeof(c->prof_launches));
    memset(c->prof_oct_ms, 0, sizeof(c->prof_oct_ms)); memset(c->prof_oct_last, 0, sizeof(c->prof_oct_last));
    for (int i = 0; i < AKZ_NSET; i++) c->img_stage[i] = nullptr;
    c->lane1 = nullptr; c->ev_fork = c->ev_join = nullptr;
    c->img_stage_bytes = 0; c->match_parts = nullptr; c->match_parts_n = 0;
    c->h2d_stream = c->d2h_stream = nullptr; c->h_cnt_pinned = nullptr;
    for (int i = 0; i < AKZ_NSET; i++) { c->ev_h2d[i] = c->ev_comp[i] = c->ev_cnt[i] = c->ev_d2h[i] = nullptr; }
    c->match_stage = nullptr; c->match_stage_bytes = 0;
    c->cur = nullptr; c->opar = false; c->graph_ok = false;
    for (int i = 0; i < 8; i++) { c->ostream[i] = nullptr; c->ev_lvl0[i] = c->ev_oct[i] = nullptr; c->sc[i] = { nullptr, nullptr, nullptr, nullptr }; }
    for (int i = 0; i < 8; i++) { c->ring_stream[i] = nullptr; c->ev_ring_fork[i] = c->ev_ring_join[i] = nullptr; c->ring_pending[i] = false; }
    c->rk = 0;
    c->comm = nullptr; c->comm_ranks = 0; c->comm_rank = 0; c->comm_owned = false; c->shard_buf = nullptr; c->shard_buf_n = 0;
    int rc = AKZ_OK;
    do {
        if (o->device >= 0) { if (cudaSetDevice(o->device) != cudaSuccess) { rc = akz_set_error(AKZ_E_CUDA, "cudaSetDevice(%d) failed", o->device); break; } }
        cudaGetDevice(&c->device);
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault) != cudaSuccess) { rc = akz_set_error(AKZ_E_CUDA, "stream creation failed"); break; }
        c->cur = c->stream;
        const int B = o->max_batch;
        // small per-frame scalars: needed by the stage seams even without a pyramid
        if ((rc = dalloc(c, &c->prefix, (size_t)2 * B + 2)) != AKZ_OK) break;
        if (!matcher_only && (rc = dalloc(c, &c->order, (size_t)B * o->max_pts)) != AKZ_OK) break;      // processing order of the keypoint stages
        if ((rc = dalloc(c, &c->hist, (size_t)AKZ_NBINS * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->hmax, (size_t)B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->kc, (size_t)B)) != AKZ_OK) break;
        if (matcher_only) { c->nlev = 0; c->noct = 0; break; }
        if ((rc = build_schedule(c)) != AKZ_OK) break;
        for (int l = 0; l < c->nlev && rc == AKZ_OK; l++) {
            AkzLevel& L = c->lev[l];
            size_t n = (size_t)L.plane * B;
            if ((rc = dalloc(c, &L.lt, n)) != AKZ_OK) break;
            if ((rc = dalloc(c, &L.det, n)) != AKZ_OK) break;
            if ((rc = dalloc(c, &L.lx, n)) != AKZ_OK) break;
            if ((rc = dalloc(c, &L.ly, n)) != AKZ_OK) break;
            if (o->fused == 1 && (L.pitch % 4) == 0 && (rc = dalloc(c, &c->hot[l], (size_t)(L.pitch / 4) * L.h * B)) != AKZ_OK) break;
        }
        if (rc != AKZ_OK) break;
        size_t n0 = (size_t)c->lev[0].plane * B;
        if ((rc = dalloc(c, &c->smooth, n0)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->flow, n0)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->tmpA, n0)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->tmpB, n0)) != AKZ_OK) break;
        {
            static const bool ring_side = [] { const char* e = getenv("AKZ_RING_SIDE"); return !e || atoi(e) != 0; }();
            const int nring = ring_side ? c->noct : 0;
            for (int k = 0; k < nring && k < 8; k++) {
                if (cudaStreamCreateWithFlags(&c->ring_stream[k], cudaStreamNonBlocking) != cudaSuccess) { rc = akz_set_error(AKZ_E_CUDA, "stream creation failed"); break; }
                cudaEventCreateWithFlags(&c->ev_ring_fork[k], cudaEventDisableTiming);
                cudaEventCreateWithFlags(&c->ev_ring_join[k], cudaEventDisableTiming);
            }
            if (rc != AKZ_OK) break;
        }
        c->sc[0] = { c->smooth, c->flow, c->tmpA, c->tmpB };
        for (int k = 1; k < 8; k++) c->sc[k] = c->sc[0];
        {
            static const bool small_on = [] { const char* e = getenv("AKZ_SMALL_BATCH"); return !e || atoi(e) != 0; }();
            // octave chains on their own streams for every batch size (AKZ_OPAR_ALL=0: only for batches of up to 4 frames): with
            // chunks of 32 frames the resident rate is unchanged (two lanes already fill the machine) but the host pipeline, whose
            // first chunks are short and whose last chunk runs alone, gains 1.6 % (float frames) / 3.7 % (u8 frames)
            static const bool opar_all = [] { const char* e = getenv("AKZ_OPAR_ALL"); return !e || atoi(e) != 0; }();
            if (small_on && (B <= 4 || opar_all) && o->fused == 1 && c->noct > 1) {
                for (int k = 1; k < c->noct && rc == AKZ_OK; k++) {
                    const size_t nk = (size_t)c->lev[k * o->max_scale].plane * B;
                    if ((rc = dalloc(c, &c->sc[k].smooth, nk)) != AKZ_OK) break;
                    if ((rc = dalloc(c, &c->sc[k].flow, nk)) != AKZ_OK) break;
                    if ((rc = dalloc(c, &c->sc[k].tmpA, nk)) != AKZ_OK) break;
                    if ((rc = dalloc(c, &c->sc[k].tmpB, nk)) != AKZ_OK) break;
                }
                if (rc != AKZ_OK) break;
                c->ostream[0] = c->stream;
                for (int k = 0; k < c->noct; k++) {
                    if (k > 0 && cudaStreamCreateWithFlags(&c->ostream[k], cudaStreamNonBlocking) != cudaSuccess) { rc = akz_set_error(AKZ_E_CUDA, "stream creation failed"); break; }
                    cudaEventCreateWithFlags(&c->ev_lvl0[k], cudaEventDisableTiming);
                    cudaEventCreateWithFlags(&c->ev_oct[k], cudaEventDisableTiming);
                }
                if (rc != AKZ_OK) break;
                c->opar = true;
                c->graph_ok = B <= 4;
            }
        }
        c->mpitch = c->lev[0].pitch;
        c->mplane = (long long)c->mpitch * o->height;
        if ((rc = dalloc(c, &c->map, (size_t)c->mplane * B)) != AKZ_OK) break;
        int mwords = (o->width + 31) / 32;
        if ((rc = dalloc(c, &c->rowmask, (size_t)mwords * o->height * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->occ, (size_t)mwords * o->height * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->rowcount, (size_t)o->height * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->counts_own, (size_t)AKZ_NSET * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->kpts_own, (size_t)AKZ_NSET * o->max_pts * B)) != AKZ_OK) break;
        if ((rc = dalloc(c, &c->desc_own, (size_t)AKZ_NSET * o->max_pts * B * 64)) != AKZ_OK) break;
        if (cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = akz_set_error(AKZ_E_CUDA, "stream creation failed"); break; }
        for (int i = 0; i < AKZ_NSET; i++) {
            cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_cnt[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming);
        }
        if (cudaMallocHost((void**)&c->h_cnt_pinned, sizeof(int) * AKZ_NSET * B) != cudaSuccess) { rc = akz_set_error(AKZ_E_NOMEM, "pinned allocation failed"); break; }
        cudaMemsetAsync(c->rowcount, 0, sizeof(int) * (size_t)o->height * B, c->stream);
        cudaMemsetAsync(c->rowmask, 0, sizeof(unsigned) * (size_t)mwords * o->height * B, c->stream);
        // the key map and its occupancy bitmap start clean and every chunk leaves them clean (k_clear_map)
        cudaMemsetAsync(c->occ, 0, sizeof(unsigned) * (size_t)mwords * o->height * B, c->stream);
        cudaMemsetAsync(c->map, 0, sizeof(unsigned long long) * (size_t)c->mplane * B, c->stream);
        // level table for the keypoint kernels
        memset(&c->tab, 0, sizeof(c->tab));
        c->tab.nlevels = c->nlev; c->tab.max_scale = o->max_scale;
        for (int l = 0; l < c->nlev; l++) {
            const AkzLevel& L = c->lev[l];
            AkzLevelDev& D = c->tab.lv[l];
            D.lt = L.lt; D.det = L.det; D.lx = L.lx; D.ly = L.ly; D.plane = L.plane;
            D.w = L.w; D.h = L.h; D.pitch = L.pitch; D.octave = L.octave; D.size = L.size;
        }
        akzk::orient_table_init(c->stream);
        if ((rc = akzk::describe_prepare(o->descriptor_pattern_size)) != AKZ_OK) break;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = akz_set_cuda_error(cudaGetLastError(), "context init", __FILE__, __LINE__); break; }
    } while (0);
    if (rc == AKZ_OK && !matcher_only && o->lanes >= 2) {
        akz_options o1 = *o;
        o1.lanes = 1;
        rc = akz_create(&o1, &c->lane1);
        if (rc == AKZ_OK) {
            cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
        }
    }
    if (rc != AKZ_OK) { akz_destroy(c); return rc; }
    *out = c;
    return AKZ_OK;
}

void akz_destroy(akz_ctx* c)
{
    if (!c) return;
    if (c->lane1) akz_destroy(c->lane1);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (int k = 0; k < 8; k++) {
        if (c->ring_stream[k]) { cudaStreamSynchronize(c->ring_stream[k]); cudaStreamDestroy(c->ring_stream[k]); }
        if (c->ev_ring_fork[k]) cudaEventDestroy(c->ev_ring_fork[k]);
        if (c->ev_ring_join[k]) cudaEventDestroy(c->ev_ring_join[k]);
    }
    for (int k = 0; k < 8; k++) {
        if (k > 0 && c->ostream[k]) { cudaStreamSynchronize(c->ostream[k]); cudaStreamDestroy(c->ostream[k]); }
        if (c->ev_lvl0[k]) cudaEventDestroy(c->ev_lvl0[k]);
        if (c->ev_oct[k]) cudaEventDestroy(c->ev_oct[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (void* p : c->allocs) cudaFree(p);
    for (int i = 0; i < AKZ_NSET; i++) {
        if (c->img_stage[i]) cudaFree(c->img_stage[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_cnt[i]) cudaEventDestroy(c->ev_cnt[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    if (c->h_cnt_pinned) cudaFreeHost(c->h_cnt_pinned);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    akz_comm_destroy(c);
    if (c->shard_buf) cudaFree(c->shard_buf);
    if (c->match_parts) cudaFree(c->match_parts);
    if (c->match_stage) cudaFree(c->match_stage);
    for (auto& pp : c->prof_pairs) { cudaEventDestroy(pp.a); cudaEventDestroy(pp.b); }
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int akz_sync(akz_ctx* c)
{
    if (!c) return akz_set_error(AKZ_E_INVALID, "null context");
    if (c->lane1) AKZ_CUDA_TRY(cudaStreamSynchronize(c->lane1->stream));
    AKZ_CUDA_TRY(cudaStreamSynchronize(c->stream));
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}
void* akz_stream(akz_ctx* c) { return (void*)c->stream; }
int akz_num_levels(const akz_ctx* c) { return c->nlev; }
int akz_launch_count(const akz_ctx* c) { return (int)(c->launches + (c->lane1 ? c->lane1->launches : 0)); }

int akz_level_info(const akz_ctx* c, int l, int* w, int* h, int* pitch, int* nsteps, float* size, int* sigma_size, float* tau, int cap)
{
    if (l < 0 || l >= c->nlev) return akz_set_error(AKZ_E_INVALID, "level out of range");
    const AkzLevel& L = c->lev[l];
    if (w) *w = L.w; if (h) *h = L.h; if (pitch) *pitch = L.pitch; if (nsteps) *nsteps = L.nsteps;
    if (size) *size = L.size; if (sigma_size) *sigma_size = L.sigma_size;
    if (tau) for (int i = 0; i < L.nsteps && i < cap; i++) tau[i] = c->tau[L.tau_off + i];
    return AKZ_OK;
}

const float* akz_level_plane(const akz_ctx* c, int l, int which, int frame)
{
    if (l < 0 || l >= c->nlev || frame < 0 || frame >= c->opt.max_batch) return nullptr;
    const AkzLevel& L = c->lev[l];
    const float* b = which == AKZ_PLANE_LT ? L.lt : which == AKZ_PLANE_DET ? L.det : which == AKZ_PLANE_LX ? L.lx : L.ly;
    return b + (long long)frame * L.plane;
}

}  // extern "C"

// ---- pipeline ------------------------------------------------------------------------------------------
#define LAUNCHED(cls_, expr) do {                                                              \
        cudaEvent_t ea_ = nullptr;                                                             \
        if (c->prof_on) { ea_ = prof_event(c); cudaEventRecord(ea_, c->cur); }              \
        int r_ = (expr);                                                                       \
        if (r_ < 0) return r_;                                                                 \
        c->launches += r_;                                                                     \
        if (ea_) { cudaEvent_t eb_ = prof_event(c); cudaEventRecord(eb_, c->cur);           \
                   c->prof_pairs.push_back({ ea_, eb_, (cls_), r_, c->cur_oct }); }                        \
    } while (0)

static int check_frame_args(akz_ctx* c, const void* img, int dtype, int nframes, int w, int h, int pitch, long long stride)
{
    if (!c || !img) return akz_set_error(AKZ_E_INVALID, "null argument");
    if (dtype != AKZ_F32 && dtype != AKZ_U8) return akz_set_error(AKZ_E_INVALID, "dtype must be AKZ_F32 or AKZ_U8");
    if (c->nlev == 0) return akz_set_error(AKZ_E_INVALID, "matcher-only context");
    if (w != c->opt.width || h != c->opt.height) return akz_set_error(AKZ_E_INVALID, "frame size %dx%d differs from the context's %dx%d", w, h, c->opt.width, c->opt.height);
    if (pitch < w || nframes < 0) return akz_set_error(AKZ_E_INVALID, "bad pitch or frame count");
    if (nframes > 1 && stride < (long long)pitch * h) return akz_set_error(AKZ_E_INVALID, "frame stride %lld smaller than pitch * height = %lld", stride, (long long)pitch * h);
    return AKZ_OK;
}

static bool prep_split_enabled()
{
    static const bool on = [] { const char* e = getenv("AKZ_PREP_SPLIT"); return !e || atoi(e) != 0; }();      // A/B knob, read once
    return on;
}

// Split level pipeline (k_prep3 + k_deriv4).  Returns the number of launches, 0 when the level is not covered.
static int prep_level_split(akz_ctx* c, int mode, const float* src, int sw, int sh, int sp, long long splane, float* ltdst, float* flowp,
                            float* lx, float* ly, float* det, int nmul, int step, int w, int h, int pitch, long long plane, int nf, int int_planes)
{
    // the context's own level?  then its `hot` plane is written along with the determinant (and is stale until it has been)
    int li = -1;
    for (int l = 0; l < c->nlev; l++) if (c->lev[l].det == det && c->lev[l].w == w && c->lev[l].h == h && c->lev[l].pitch == pitch) li = l;
    if (li >= 0) c->hot_valid[li] = false;
    if (!prep_split_enabled() || step < 2 || step > 4 || w < 32 || h < 32 || (w % 4) != 0) return 0;
    cudaStream_t st = c->cur;
    const int rk = c->rk;
    cudaStream_t rst = c->prof_on ? nullptr : c->ring_stream[rk];         // per-class timing keeps the ring kernels on the main stream
    // the ring kernels of the previous level (side stream) read the blurred plane this level is about to overwrite
    if (c->ring_pending[rk]) { AKZ_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_ring_join[rk], 0)); c->ring_pending[rk] = false; }
    unsigned char* hot = li >= 0 ? c->hot[li] : nullptr;
    const float thr = c->opt.dthreshold;
    const int ithr = 65;                                                  // akaze.cpp:560
    if (mode == 0) {
        int r0 = akzk::deriv_stream(st, src, lx, ly, det, step, w, h, pitch, plane, nf, int_planes, rst, c->ev_ring_fork[rk], c->ev_ring_join[rk], hot, thr, ithr);
        if (r0 > 0 && rst) c->ring_pending[rk] = true;
        if (r0 > 0 && hot) c->hot_valid[li] = true;
        return r0;
    }
    if (!c->smooth || !flowp) return 0;
    // both halves must take the level: probe the derivative half's conditions first (it has the stricter ones)
    if ((pitch % 4) != 0 || (plane % 4) != 0 || (((uintptr_t)c->smooth | (uintptr_t)lx | (uintptr_t)ly | (uintptr_t)det) % 16) != 0) return 0;
    // same-resolution levels: both halves in one streaming kernel (no round trip of the blurred plane) when there are enough CTAs
    static const int level4 = [] { const char* e = getenv("AKZ_LEVEL4"); return e ? atoi(e) : 0; }();      // A/B knob: 0 off (default until it beats the pair), 1 on, 2 forced (tests)
    if (mode == 1 && level4 && src != c->smooth) {
        const int r4 = akzk::level_stream(st, src, flowp, c->smooth, lx, ly, det, c->opt.diffusivity, c->kc, 0.75f, nmul, step, w, h, pitch, plane, nf, int_planes,
                                          rst, c->ev_ring_fork[rk], c->ev_ring_join[rk], hot, thr, ithr, level4 == 2);
        if (r4 < 0) return r4;
        if (r4 > 0) {
            if (rst) c->ring_pending[rk] = true;
            if (hot) c->hot_valid[li] = true;
            return r4;
        }
    }
    static const bool blur_stream_on = [] { const char* e = getenv("AKZ_BLUR_STREAM"); return !e || atoi(e) != 0; }();      // A/B knob
    int r1 = blur_stream_on ? akzk::blur_stream(st, mode, src, sw, sh, sp, splane, ltdst, flowp, c->smooth, c->opt.diffusivity, c->kc, 0.75f, nmul,
                                                w, h, pitch, plane, nf, int_planes) : 0;
    if (r1 == 0)
        r1 = akzk::level_blur_flow(st, mode, src, sw, sh, sp, splane, ltdst, flowp, c->smooth, c->opt.diffusivity, c->kc, 0.75f, nmul,
                                   w, h, pitch, plane, nf, int_planes);
    if (r1 <= 0) return r1;
    int r2 = akzk::deriv_stream(st, c->smooth, lx, ly, det, step, w, h, pitch, plane, nf, int_planes, rst, c->ev_ring_fork[rk], c->ev_ring_join[rk], hot, thr, ithr);
    if (r2 < 0) return r2;
    if (r2 == 0) return akz_set_error(AKZ_E_UNSUPPORTED, "split level pipeline: derivative half refused a level the blur half took");
    if (rst) c->ring_pending[rk] = true;
    if (hot) c->hot_valid[li] = true;
    return r1 + r2;
}

// One level's "everything but the diffusion": the fused kernel k_prep2 when it covers the case (fused == 1, derivative step 2..4,
// level at least 24 x 24); otherwise one kernel per reference stage (blur or octave transition, conductance, derivatives).
// mode: 0 = base level (smooth := src), 1 = same-resolution blur, 2 = octave transition.
static int prep_level(akz_ctx* c, int mode, const float* src, int sw, int sh, int sp, long long splane, float* ltdst, float* flowp,
                      float* lx, float* ly, float* det, int nmul, int step, int w, int h, int pitch, long long plane, int nf)
{
    const akz_options& o = c->opt;
    cudaStream_t st = c->cur;
    if (o.fused == 1) {
        // split pipeline: tile kernel for the blur / octave transition + conductance (blurred plane to c->smooth), streaming warp
        // kernel for the derivatives and the determinant; the single tile kernel k_prep2 takes what they do not cover
        int r = prep_level_split(c, mode, src, sw, sh, sp, splane, ltdst, flowp, lx, ly, det, nmul, step, w, h, pitch, plane, nf, 0);
        if (r != 0) return r;
        r = akzk::level_prep2(st, mode, src, sw, sh, sp, splane, ltdst, flowp, lx, ly, det, o.diffusivity, c->kc, 0.75f, nmul,
                              step, w, h, pitch, plane, nf);
        if (r != 0) return r;
    }
    int launches = 0, r = 0;
    const float* sm = src;
    if (mode != 0) {
        if (!c->smooth) return akz_set_error(AKZ_E_INVALID, "this context has no scratch planes for the staged level path");
        if (mode == 2) r = akzk::down_with_smooth(st, src, ltdst, c->smooth, sw, sh, sp, splane, w, h, pitch, plane, nf);
        else r = akzk::lowpass(st, src, c->smooth, w, h, pitch, plane, pitch, plane, nf, 1.f, 5);
        if (r < 0) return r;
        launches += r;
        sm = c->smooth;
    }
    if (flowp) {
        if ((r = akzk::flow(st, sm, flowp, o.diffusivity, c->kc, 0.75f, nmul, w, h, pitch, plane, nf)) < 0) return r;
        launches += r;
    }
    if ((r = akzk::hessian(st, sm, lx, ly, det, step, w, h, pitch, plane, nf)) < 0) return r;
    return launches + r;
}

// Octave switch of the pipeline: the launches that follow go to the octave's stream and scratch planes.  Without per-octave
// resources (large-batch contexts) every octave uses the context's stream and the one scratch set.
static int enter_octave(akz_ctx* c, int oc)
{
    const int k = c->opar ? oc : 0;
    c->cur_oct = oc & 7;
    c->smooth = c->sc[k].smooth; c->flow = c->sc[k].flow; c->tmpA = c->sc[k].tmpA; c->tmpB = c->sc[k].tmpB;
    c->cur = c->opar ? c->ostream[oc] : c->stream;
    c->rk = c->opar ? oc : 0;
    // octave oc starts from level (oc - 1, 0) (akaze.cpp:371-392) and the contrast factor: both are ready at that octave's event
    if (c->opar && oc > 0) AKZ_CUDA_TRY(cudaStreamWaitEvent(c->cur, c->ev_lvl0[oc - 1], 0));
    return AKZ_OK;
}

static int scale_space_levels(akz_ctx* c, const void* img, int dtype, int nf, int ipitch, long long istride)
{
    const akz_options& o = c->opt;
    const int S = o.max_scale, fused = o.fused;
    AkzLevel& L0 = c->lev[0];
    const int w0 = L0.w, h0 = L0.h, p0 = L0.pitch;
    int rc = enter_octave(c, 0);
    if (rc != AKZ_OK) return rc;
    cudaStream_t st = c->cur;
    // level (0,0): akaze.cpp:325-346
    const float var0 = o.soffset * o.soffset;
    const int ksz0 = (int)(2 * ceilf((o.soffset - 0.8f) / 0.3f) + 3);
    bool base_done = false;
    if (fused == 1) {
        // one pass over the input: Lt(0,0) + gradient magnitude plane (c->smooth) + maximum + histogram
        int r = 0;
        LAUNCHED(AKZ_K_BASE, (r = akzk::base_level2(st, img, dtype, w0, h0, ipitch, istride, L0.lt, p0, L0.plane,
                                                    o.kcontrast_override > 0.f ? nullptr : c->smooth, c->hmax, c->hist, var0, ksz0, nf)));
        if (r > 0) {
            base_done = true;
            LAUNCHED(AKZ_K_CONTRAST, akzk::contrast_scan(st, c->hmax, c->hist, c->kc, o.per, o.kcontrast_override, w0, h0, nf));
        }
    }
    if (!base_done) {
        if (dtype == AKZ_U8) {
            if (!(o.kcontrast_override > 0.f))
                LAUNCHED(AKZ_K_BASE, akzk::lowpass_u8(st, (const unsigned char*)img, c->smooth, w0, h0, ipitch, istride, p0, L0.plane, nf, 1.f, 5));
            LAUNCHED(AKZ_K_BASE, akzk::lowpass_u8(st, (const unsigned char*)img, L0.lt, w0, h0, ipitch, istride, p0, L0.plane, nf, var0, ksz0));
        } else {
            if (!(o.kcontrast_override > 0.f))
                LAUNCHED(AKZ_K_BASE, akzk::lowpass(st, (const float*)img, c->smooth, w0, h0, ipitch, istride, p0, L0.plane, nf, 1.f, 5));
            LAUNCHED(AKZ_K_BASE, akzk::lowpass(st, (const float*)img, L0.lt, w0, h0, ipitch, istride, p0, L0.plane, nf, var0, ksz0));
        }
        LAUNCHED(AKZ_K_CONTRAST, akzk::contrast(st, c->smooth, c->hmax, c->hist, c->kc, o.per, o.kcontrast_override, w0, h0, p0, L0.plane, nf));
    }
    if (c->opar) AKZ_CUDA_TRY(cudaEventRecord(c->ev_lvl0[0], st));            // Lt(0,0) and the contrast factors are ready
    if (fused) LAUNCHED(AKZ_K_PREP, prep_level(c, 0, L0.lt, w0, h0, p0, L0.plane, nullptr, nullptr, L0.lx, L0.ly, L0.det, 0, L0.sigma_size, w0, h0, p0, L0.plane, nf));
    else LAUNCHED(AKZ_K_HESSIAN, akzk::hessian(st, L0.lt, L0.lx, L0.ly, L0.det, L0.sigma_size, w0, h0, p0, L0.plane, nf));

    for (int l = 1; l < c->nlev; l++) {
        AkzLevel& L = c->lev[l];
        const float* tau = c->tau + L.tau_off;
        const int w = L.w, h = L.h, p = L.pitch;
        if (L.sub == 0) {
            // new octave (akaze.cpp:371-392): source is sublevel 0 of the previous octave; kcontrast *= 0.75
            if ((rc = enter_octave(c, L.octave)) != AKZ_OK) return rc;
            st = c->cur;
            AkzLevel& P = c->lev[l - S];
            if (fused) {
                LAUNCHED(AKZ_K_PREP, prep_level(c, 2, P.lt, P.w, P.h, P.pitch, P.plane, c->tmpB, c->flow, L.lx, L.ly, L.det, L.octave,
                                                L.sigma_size, w, h, p, L.plane, nf));
            } else {
                LAUNCHED(AKZ_K_BLUR, akzk::down_with_smooth(st, P.lt, c->tmpB, c->smooth, P.w, P.h, P.pitch, P.plane, w, h, p, L.plane, nf));
                LAUNCHED(AKZ_K_FLOW, akzk::flow(st, c->smooth, c->flow, o.diffusivity, c->kc, 0.75f, L.octave, w, h, p, L.plane, nf));
            }
            LAUNCHED(AKZ_K_FED, akzk::fed_cycle(st, c->tmpB, c->flow, L.lt, c->tmpA, tau, L.nsteps, w, h, p, L.plane, nf, fused));
            if (c->opar) AKZ_CUDA_TRY(cudaEventRecord(c->ev_lvl0[L.octave], st));
        } else {
            // next sublevel (akaze.cpp:393-421)
            AkzLevel& P = c->lev[l - 1];
            if (fused) {
                LAUNCHED(AKZ_K_PREP, prep_level(c, 1, P.lt, w, h, p, L.plane, nullptr, c->flow, L.lx, L.ly, L.det, L.octave, L.sigma_size, w, h, p, L.plane, nf));
            } else {
                LAUNCHED(AKZ_K_BLUR, akzk::lowpass(st, P.lt, c->smooth, w, h, p, L.plane, p, L.plane, nf, 1.f, 5));
                LAUNCHED(AKZ_K_FLOW, akzk::flow(st, c->smooth, c->flow, o.diffusivity, c->kc, 0.75f, L.octave, w, h, p, L.plane, nf));
            }
            LAUNCHED(AKZ_K_FED, akzk::fed_cycle(st, P.lt, c->flow, L.lt, c->tmpA, tau, L.nsteps, w, h, p, L.plane, nf, fused));
        }
        if (!fused) LAUNCHED(AKZ_K_HESSIAN, akzk::hessian(st, c->smooth, L.lx, L.ly, L.det, L.sigma_size, w, h, p, L.plane, nf));   // akaze.cpp:423
    }
    c->last_frames = nf;
    return AKZ_OK;
}

// the side streams of the border-ring kernels join the stream of their octave (also after an error: nothing stays forked)
static int join_rings(akz_ctx* c)
{
    int rj = AKZ_OK;
    for (int k = 0; k < 8; k++)
        if (c->ring_pending[k]) {
            if (cudaStreamWaitEvent(c->opar ? c->ostream[k] : c->stream, c->ev_ring_join[k], 0) != cudaSuccess) rj = akz_set_error(AKZ_E_CUDA, "joining a ring stream failed");
            c->ring_pending[k] = false;
        }
    return rj;
}

static int scale_space_chunk(akz_ctx* c, const void* img, int dtype, int nf, int ipitch, long long istride)
{
    const int rc = scale_space_levels(c, img, dtype, nf, ipitch, istride);
    // back to the context's stream and scratch set; the octave streams join it (also after an error: nothing stays forked)
    int rj = join_rings(c);
    if (c->opar)
        for (int k = 1; k < c->noct; k++) {
            if (cudaEventRecord(c->ev_oct[k], c->ostream[k]) != cudaSuccess || cudaStreamWaitEvent(c->stream, c->ev_oct[k], 0) != cudaSuccess)
                rj = akz_set_error(AKZ_E_CUDA, "joining the octave streams failed");
        }
    c->cur = c->stream; c->rk = 0; c->cur_oct = 0;
    c->smooth = c->sc[0].smooth; c->flow = c->sc[0].flow; c->tmpA = c->sc[0].tmpA; c->tmpB = c->sc[0].tmpB;
    return rc != AKZ_OK ? rc : rj;
}

// integer pipeline, Akazer::fastDetect (akaze.cpp:506-743): the plane buffers of the context are reused as int32 planes
static int fast_scale_space_levels(akz_ctx* c, const unsigned char* img, int nf, int ipitch, long long istride)
{
    cudaStream_t st = c->stream;
    const akz_options& o = c->opt;
    const int S = o.max_scale;
    AkzLevel& L0 = c->lev[0];
    const int w0 = L0.w, h0 = L0.h, p0 = L0.pitch;
    int* smooth = (int*)c->smooth; int* flow = (int*)c->flow; int* tA = (int*)c->tmpA; int* tB = (int*)c->tmpB;
    int* ikc = (int*)c->kc; int* ihmax = (int*)c->hmax;
    const float var0 = o.soffset * o.soffset;
    const int ksz0 = (int)(2 * ceilf((o.soffset - 0.8f) / 0.3f) + 3);
    // level (0,0): akaze.cpp:593-617
    bool base_done = false;
    if (o.fused == 1) {
        // one pass over the input (the float pipeline's k_base2 instantiated for the integer arithmetic): Lt(0,0), the gradient
        // magnitude plane of the sigma = 1 blur and its maximum; then the integer histogram and the percentile scan
        int r = 0;
        LAUNCHED(AKZ_K_CONTRAST, akzk::fast_contrast_init(st, ihmax, c->hist, nf));
        LAUNCHED(AKZ_K_BASE, (r = akzk::base_level2(st, img, AKZ_U8, w0, h0, ipitch, istride, L0.lt, p0, L0.plane,
                                                    o.fast_kcontrast_override > 0 ? nullptr : (float*)tB, (unsigned*)ihmax, c->hist, var0, ksz0, nf, 1)));
        if (r > 0) {
            base_done = true;
            LAUNCHED(AKZ_K_CONTRAST, akzk::fast_contrast_tail(st, tB, ihmax, c->hist, ikc, o.per, o.fast_kcontrast_override, w0, h0, p0, L0.plane, nf));
        }
    }
    if (!base_done) {
        LAUNCHED(AKZ_K_BASE, akzk::fast_lowpass(st, img, 1, smooth, tA, w0, h0, ipitch, istride, p0, L0.plane, nf, 1.f, 5));
        LAUNCHED(AKZ_K_CONTRAST, akzk::fast_contrast(st, smooth, tB, ihmax, c->hist, ikc, o.per, o.fast_kcontrast_override, w0, h0, p0, L0.plane, nf));
        LAUNCHED(AKZ_K_BASE, akzk::fast_lowpass(st, img, 1, (int*)L0.lt, tA, w0, h0, ipitch, istride, p0, L0.plane, nf, var0, ksz0));
    }
    // fused == 1: blur / octave transition + conductance + derivatives + determinant of a level in ONE kernel, the float
    // pipeline's k_prep2 instantiated for the integer arithmetic (level_prep.cu); 0 = one kernel per reference stage
    auto iprep = [&](int mode, const int* src, int sw, int sh, int sp, long long splane, int* ltdst, int* flowp, AkzLevel& L, int nmul) -> int {
        if (o.fused != 1) return 0;
        int rs = prep_level_split(c, mode, (const float*)src, sw, sh, sp, splane, (float*)ltdst, (float*)flowp, L.lx, L.ly, L.det, nmul,
                                  L.sigma_size, L.w, L.h, L.pitch, L.plane, nf, 1);
        if (rs != 0) return rs;
        return akzk::level_prep2(st, mode, (const float*)src, sw, sh, sp, splane, (float*)ltdst, (float*)flowp, L.lx, L.ly, L.det, o.diffusivity,
                                 c->kc, 0.75f, nmul, L.sigma_size, L.w, L.h, L.pitch, L.plane, nf, 1);
    };
    {
        int r = 0;
        LAUNCHED(AKZ_K_PREP, (r = iprep(0, (const int*)L0.lt, w0, h0, p0, L0.plane, nullptr, nullptr, L0, 0)));
        if (r == 0) LAUNCHED(AKZ_K_HESSIAN, akzk::fast_hessian(st, (const int*)L0.lt, (int*)L0.lx, (int*)L0.ly, (int*)L0.det, L0.sigma_size, w0, h0, p0, L0.plane, nf));
    }
    for (int l = 1; l < c->nlev; l++) {
        AkzLevel& L = c->lev[l];
        const float* tau = c->tau + L.tau_off;
        const int w = L.w, h = L.h, p = L.pitch;
        const int* cur;
        int r = 0;
        if (L.sub == 0) {
            AkzLevel& P = c->lev[l - S];                                   // sublevel 0 of the previous octave (akaze.cpp:650)
            LAUNCHED(AKZ_K_PREP, (r = iprep(2, (const int*)P.lt, P.w, P.h, P.pitch, P.plane, tB, flow, L, L.octave)));
            if (r == 0) LAUNCHED(AKZ_K_BLUR, akzk::fast_down(st, (const int*)P.lt, tB, smooth, P.w, P.h, P.pitch, P.plane, w, h, p, L.plane, nf));
            cur = tB;
        } else {
            AkzLevel& P = c->lev[l - 1];
            LAUNCHED(AKZ_K_PREP, (r = iprep(1, (const int*)P.lt, w, h, p, L.plane, nullptr, flow, L, L.octave)));
            if (r == 0) LAUNCHED(AKZ_K_BLUR, akzk::fast_lowpass(st, P.lt, 0, smooth, tA, w, h, p, L.plane, p, L.plane, nf, 1.f, 5));
            cur = (const int*)P.lt;
        }
        if (r == 0) LAUNCHED(AKZ_K_FLOW, akzk::fast_flow(st, smooth, flow, o.diffusivity, ikc, L.octave, w, h, p, L.plane, nf));
        if (o.fused == 1) {
            // the temporally blocked, register-resident FED kernel of the float pipeline, instantiated for int32 planes
            LAUNCHED(AKZ_K_FED, akzk::fed_cycle(st, (const float*)cur, (const float*)flow, L.lt, (float*)tA, tau, L.nsteps, w, h, p, L.plane, nf, 1, 1));
        } else {
            for (int k = 0; k < L.nsteps; k++) {
                int* out = ((L.nsteps - 1 - k) % 2 == 0) ? (int*)L.lt : tA;
                LAUNCHED(AKZ_K_FED, akzk::fast_nld_step(st, cur, flow, out, tau[k], w, h, p, L.plane, nf));
                cur = out;
            }
        }
        if (r == 0) LAUNCHED(AKZ_K_HESSIAN, akzk::fast_hessian(st, smooth, (int*)L.lx, (int*)L.ly, (int*)L.det, L.sigma_size, w, h, p, L.plane, nf));
    }
    c->last_frames = nf;
    return AKZ_OK;
}

static int fast_scale_space_chunk(akz_ctx* c, const unsigned char* img, int nf, int ipitch, long long istride)
{
    c->rk = 0;                                   // the integer pipeline runs its octaves on the context's stream
    const int rc = fast_scale_space_levels(c, img, nf, ipitch, istride);
    const int rj = join_rings(c);                // ring index 0 belongs to the context's stream in either mode
    return rc != AKZ_OK ? rc : rj;
}

// fresh: the scale space of these frames was built by the same entry point just before (the `hot` planes describe these determinants)
static int detect_chunk(akz_ctx* c, int nf, int describe, int* d_counts, akz_keypoint* d_kpts, unsigned char* d_desc, int fast = 0, bool fresh = true)
{
    cudaStream_t st = c->stream;
    const akz_options& o = c->opt;
    const int S = o.max_scale;
    const int mwords = (o.width + 31) / 32;
    // Small-batch contexts (octave streams): the four extrema kernels merge into the same map with atomicMax and are independent,
    // so octaves 1.. run beside octave 0 on their octave streams; k_clear_map runs on a side stream under the keypoint stages.
    const bool par = c->opar && c->noct > 1;
    if (par) AKZ_CUDA_TRY(cudaEventRecord(c->ev_lvl0[0], st));
    for (int oc = 0; oc < c->noct; oc++) {
        AkzExtremaArgs a;
        memset(&a, 0, sizeof(a));
        const AkzLevel& F = c->lev[oc * S];
        a.nsub = S; a.w = F.w; a.h = F.h; a.pitch = F.pitch; a.octave = oc; a.psz = (int)F.border;     // akazed.cu:2571
        a.int_planes = fast;
        for (int j = 0; j < S; j++) {
            const AkzLevel& L = c->lev[oc * S + j];
            a.lv[j].det = L.det; a.lv[j].plane = L.plane; a.lv[j].border = L.border; a.lv[j].threshold = o.dthreshold;
            a.lv[j].layer = oc * S + j; a.lv[j].ithreshold = 65;                      // akaze.cpp:560
            a.lv[j].hot = (fresh && c->hot_valid[oc * S + j]) ? c->hot[oc * S + j] : nullptr;
        }
        cudaStream_t so = (par && oc > 0) ? c->ostream[oc] : st;
        if (so != st) AKZ_CUDA_TRY(cudaStreamWaitEvent(so, c->ev_lvl0[0], 0));
        c->cur = so;
        LAUNCHED(AKZ_K_EXTREMA, akzk::extrema(so, a, c->map, c->mpitch, c->mplane, c->occ, mwords, o.height, nf));
        c->cur = st;
        if (so != st) { AKZ_CUDA_TRY(cudaEventRecord(c->ev_oct[oc], so)); AKZ_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_oct[oc], 0)); }
    }
    LAUNCHED(AKZ_K_NMS, akzk::nms_emit(st, c->map, c->mpitch, c->mplane, o.width, o.height, c->psz, c->tab, c->occ, c->rowmask, c->rowcount,
                            d_counts, c->prefix, d_kpts, o.max_pts, nf, fast));
    cudaStream_t sc = (par && describe) ? c->ostream[1] : st;
    if (sc != st) { AKZ_CUDA_TRY(cudaEventRecord(c->ev_lvl0[0], st)); AKZ_CUDA_TRY(cudaStreamWaitEvent(sc, c->ev_lvl0[0], 0)); }
    c->cur = sc;
    LAUNCHED(AKZ_K_NMS, akzk::clear_map(sc, c->map, c->mpitch, c->mplane, o.width, o.height, c->occ, nf));
    c->cur = st;
    if (sc != st) AKZ_CUDA_TRY(cudaEventRecord(c->ev_oct[1], sc));
    if (describe) {
        LAUNCHED(AKZ_K_ORIENT, akzk::layer_order(st, d_counts, d_kpts, c->order, o.max_pts, nf));
        LAUNCHED(AKZ_K_ORIENT, akzk::orient(st, c->tab, d_counts, c->prefix, d_kpts, o.max_pts, nf, fast, c->order));
        LAUNCHED(AKZ_K_DESCRIBE, akzk::describe(st, c->tab, d_counts, c->prefix, d_kpts, d_desc, o.max_pts, nf, o.descriptor_pattern_size, fast, c->order));
    }
    if (sc != st) AKZ_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_oct[1], 0));
    return AKZ_OK;
}

// Replay of a whole chunk as a CUDA graph (small-batch contexts).  The first call with an argument set runs eagerly (lazy
// per-device initialisation stays outside any capture) and remembers the set; the second one captures the same launches --
// kernels, memsets and the fork / join of the octave streams -- instantiates the graph and launches it; later ones launch it.
template <class F>
static int run_graphed(akz_ctx* c, const unsigned long long (&key)[10], F&& body)
{
    if (!c->graph_ok || c->prof_on) return body();
    for (auto& g : c->graphs) {
        if (memcmp(g.key, key, sizeof(key)) != 0) continue;
        if (!g.exec) {
            cudaGraph_t graph = nullptr;
            const long long l0 = c->launches;
            if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_ok = false; return body(); }
            const int rc = body();
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            const int nl = (int)(c->launches - l0);
            c->launches = l0;
            if (rc != AKZ_OK || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                c->graph_ok = false;                       // this context keeps working without graphs
                return rc != AKZ_OK ? rc : body();
            }
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess || !exec) { cudaGetLastError(); c->graph_ok = false; return body(); }
            g.exec = exec; g.launches = nl;
        }
        AKZ_CUDA_TRY(cudaGraphLaunch(g.exec, c->stream));
        c->launches += g.launches;
        return AKZ_OK;
    }
    if (c->graphs.size() >= 8) {
        if (c->graphs.front().exec) { cudaStreamSynchronize(c->stream); cudaGraphExecDestroy(c->graphs.front().exec); }
        c->graphs.erase(c->graphs.begin());
    }
    akz_ctx::GraphEntry g = {};
    memcpy(g.key, key, sizeof(key));
    g.seen = 1; g.exec = nullptr; g.launches = 0;
    c->graphs.push_back(g);
    return body();
}

extern "C" {

int akz_build_scale_space(akz_ctx* c, const void* d_images, int dtype, int nframes, int w, int h, int pitch, long long stride)
{
    int rc = check_frame_args(c, d_images, dtype, nframes, w, h, pitch, stride);
    if (rc != AKZ_OK) return rc;
    if (nframes > c->opt.max_batch) return akz_set_error(AKZ_E_INVALID, "nframes exceeds max_batch");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    rc = scale_space_chunk(c, d_images, dtype, nframes, pitch, stride);
    if (rc != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}

int akz_get_kcontrast(akz_ctx* c, float* h_k, int nframes)
{
    AKZ_CUDA_TRY(cudaMemcpyAsync(h_k, c->kc, sizeof(float) * nframes, cudaMemcpyDeviceToHost, c->stream));
    return akz_sync(c);
}

// Chunk plan of a batch: full chunks of B frames; small chunks (B/8, B/4, B/2) at the start when the first upload has nothing
// to hide behind (host pipeline).  A mirrored ramp-down (so that the last chunk, which runs alone on its lane, is short) is
// implemented but off: measured at 256 frames it costs more in small-chunk inefficiency than the shorter tail returns
// (round 1: 4705 -> 4636 images/s resident, 4308 -> 4322 host to host; round 2, copy-bound float frames: 5754 -> 5683,
// scripts/probes/e2e_rampdown.sh).
static void chunk_plan(int nframes, int B, bool ramp_up, bool ramp_down, std::vector<int>& start, std::vector<int>& size)
{
    start.clear(); size.clear();
    int f = 0;
    auto push = [&](int n) { if (n > 0) { start.push_back(f); size.push_back(n); f += n; } };
    if (ramp_up)
        for (int sz = std::max(1, B / 8); sz < B && f + sz <= nframes; sz *= 2) push(sz);
    int tail[3] = { B / 2, B / 4, B / 8 }, T = 0;
    if (ramp_down) T = tail[0] + tail[1] + tail[2];
    while (nframes - f > T + B) push(B);
    int rest = nframes - f;
    if (!ramp_down || T == 0) { while (rest > 0) { int n = std::min(rest, B); push(n); rest -= n; } return; }
    if (rest > T) { push(rest - T); rest = T; }
    for (int k = 0; k < 3 && rest > 0; k++) { int n = std::min(rest, std::max(1, tail[k])); push(n); rest -= n; }
    while (rest > 0) { int n = std::min(rest, B); push(n); rest -= n; }
}

int akz_detect_and_compute(akz_ctx* c, const void* d_images, int dtype, int nframes, int w, int h, int pitch, long long stride,
                           int describe, int* d_counts, akz_keypoint* d_kpts, uint8_t* d_desc)
{
    int rc = check_frame_args(c, d_images, dtype, nframes, w, h, pitch, stride);
    if (rc != AKZ_OK) return rc;
    if (!d_counts || !d_kpts || (describe && !d_desc)) return akz_set_error(AKZ_E_INVALID, "null result buffer");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    const int B = c->opt.max_batch;
    const size_t esz = dtype == AKZ_U8 ? 1 : 4;
    if (nframes >= 1 && nframes <= B && c->graph_ok) {
        const unsigned long long key[10] = { (unsigned long long)(uintptr_t)d_images, (unsigned long long)dtype, (unsigned long long)nframes,
                                             (unsigned long long)pitch, (unsigned long long)stride, (unsigned long long)describe,
                                             (unsigned long long)(uintptr_t)d_counts, (unsigned long long)(uintptr_t)d_kpts,
                                             (unsigned long long)(uintptr_t)d_desc, 0ull };
        return run_graphed(c, key, [&]() -> int {
            int r = scale_space_chunk(c, d_images, dtype, nframes, pitch, stride);
            if (r != AKZ_OK) return r;
            return detect_chunk(c, nframes, describe, d_counts, d_kpts, d_desc);
        });
    }
    // two lanes: chunks alternate between the parent and the child context, each on its own stream
    const bool two = c->lane1 != nullptr && nframes > B;
    if (two) {
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
        AKZ_CUDA_TRY(cudaStreamWaitEvent(c->lane1->stream, c->ev_fork, 0));
    }
    std::vector<int> cstart, csize;
    chunk_plan(nframes, B, false, false, cstart, csize);
    for (int k = 0; k < (int)cstart.size(); k++) {
        const int f0 = cstart[k], nf = csize[k];
        akz_ctx* L = (two && (k & 1)) ? c->lane1 : c;
        const char* img = (const char*)d_images + (size_t)f0 * stride * esz;
        if ((rc = scale_space_chunk(L, img, dtype, nf, pitch, stride)) != AKZ_OK) break;
        if ((rc = detect_chunk(L, nf, describe, d_counts + f0, d_kpts + (size_t)f0 * c->opt.max_pts,
                               d_desc ? d_desc + (size_t)f0 * c->opt.max_pts * 64 : nullptr)) != AKZ_OK) break;
    }
    if (two) {                                   // also on an error: the second lane is always joined
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_join, c->lane1->stream));
        AKZ_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    if (rc != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}

int akz_fast_build_scale_space(akz_ctx* c, const uint8_t* d_images, int nframes, int w, int h, int pitch, long long stride)
{
    int rc = check_frame_args(c, d_images, AKZ_U8, nframes, w, h, pitch, stride);
    if (rc != AKZ_OK) return rc;
    if (nframes > c->opt.max_batch) return akz_set_error(AKZ_E_INVALID, "nframes exceeds max_batch");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    if ((rc = fast_scale_space_chunk(c, d_images, nframes, pitch, stride)) != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}

int akz_fast_get_kcontrast(akz_ctx* c, int* h_k, int nframes)
{
    AKZ_CUDA_TRY(cudaMemcpyAsync(h_k, c->kc, sizeof(int) * nframes, cudaMemcpyDeviceToHost, c->stream));
    return akz_sync(c);
}

int akz_fast_detect_and_compute(akz_ctx* c, const uint8_t* d_images, int nframes, int w, int h, int pitch, long long stride,
                                int describe, int* d_counts, akz_keypoint* d_kpts, uint8_t* d_desc)
{
    int rc = check_frame_args(c, d_images, AKZ_U8, nframes, w, h, pitch, stride);
    if (rc != AKZ_OK) return rc;
    if (!d_counts || !d_kpts || (describe && !d_desc)) return akz_set_error(AKZ_E_INVALID, "null result buffer");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    const int B = c->opt.max_batch;
    if (nframes >= 1 && nframes <= B && c->graph_ok) {
        const unsigned long long key[10] = { (unsigned long long)(uintptr_t)d_images, (unsigned long long)AKZ_U8, (unsigned long long)nframes,
                                             (unsigned long long)pitch, (unsigned long long)stride, (unsigned long long)describe,
                                             (unsigned long long)(uintptr_t)d_counts, (unsigned long long)(uintptr_t)d_kpts,
                                             (unsigned long long)(uintptr_t)d_desc, 1ull };
        return run_graphed(c, key, [&]() -> int {
            int r = fast_scale_space_chunk(c, d_images, nframes, pitch, stride);
            if (r != AKZ_OK) return r;
            return detect_chunk(c, nframes, describe, d_counts, d_kpts, d_desc, 1);
        });
    }
    const bool two = c->lane1 != nullptr && nframes > B;
    if (two) {
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
        AKZ_CUDA_TRY(cudaStreamWaitEvent(c->lane1->stream, c->ev_fork, 0));
    }
    std::vector<int> cstart, csize;
    chunk_plan(nframes, B, false, false, cstart, csize);
    for (int k = 0; k < (int)cstart.size(); k++) {
        const int f0 = cstart[k], nf = csize[k];
        akz_ctx* L = (two && (k & 1)) ? c->lane1 : c;
        if ((rc = fast_scale_space_chunk(L, d_images + (size_t)f0 * stride, nf, pitch, stride)) != AKZ_OK) break;
        if ((rc = detect_chunk(L, nf, describe, d_counts + f0, d_kpts + (size_t)f0 * c->opt.max_pts,
                               d_desc ? d_desc + (size_t)f0 * c->opt.max_pts * 64 : nullptr, 1)) != AKZ_OK) break;
    }
    if (two) {                                   // also on an error: the second lane is always joined
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_join, c->lane1->stream));
        AKZ_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    if (rc != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}

static int detect_and_compute_host_impl(akz_ctx* c, const void* h_images, int dtype, int nframes, int w, int h, int pitch, long long stride,
                                        int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc, int fast);

// On an error in the middle of the pipeline nothing may stay in flight: copies into the caller's buffers and both lanes are
// drained before the error is returned (the message of the first failure is kept).
static int detect_and_compute_host(akz_ctx* c, const void* h_images, int dtype, int nframes, int w, int h, int pitch, long long stride,
                                   int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc, int fast)
{
    const int rc = detect_and_compute_host_impl(c, h_images, dtype, nframes, w, h, pitch, stride, describe, h_counts, h_kpts, h_desc, fast);
    if (rc != AKZ_OK && c && c->nlev > 0) {
        if (c->h2d_stream) cudaStreamSynchronize(c->h2d_stream);
        cudaStreamSynchronize(c->stream);
        if (c->lane1) cudaStreamSynchronize(c->lane1->stream);
        if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
        cudaGetLastError();
    }
    return rc;
}

static int detect_and_compute_host_impl(akz_ctx* c, const void* h_images, int dtype, int nframes, int w, int h, int pitch, long long stride,
                                        int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc, int fast)
{
    int rc = check_frame_args(c, h_images, dtype, nframes, w, h, pitch, stride);
    if (rc != AKZ_OK) return rc;
    if (!h_counts || !h_kpts || (describe && !h_desc)) return akz_set_error(AKZ_E_INVALID, "null result buffer");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    const int B = c->opt.max_batch, MP = c->opt.max_pts;
    const size_t esz = dtype == AKZ_U8 ? 1 : 4;
    const size_t need = (size_t)B * stride * esz;
    if (c->img_stage_bytes < need) {
        for (int i = 0; i < AKZ_NSET; i++) {
            if (c->img_stage[i]) cudaFree(c->img_stage[i]);
            c->img_stage[i] = nullptr;
        }
        c->img_stage_bytes = 0;
        for (int i = 0; i < AKZ_NSET; i++) AKZ_CUDA_TRY(cudaMalloc(&c->img_stage[i], need));
        c->img_stage_bytes = need;
    }
    // Pipeline over chunks of max_batch frames with AKZ_NSET staging buffers / result sets (chunk i uses set i % AKZ_NSET):
    //   h2d stream   : frames of chunk i+LAG -> staging                  (overlaps the kernels of the chunks before it)
    //   lane streams : scale space + detector + descriptors of chunk i -> result set; with two lanes (akz_options::lanes)
    //                  consecutive chunks run on different streams with their own pyramids, so two are in flight
    //   d2h stream   : counts of chunk i, then (once the host knows them) one strided copy of keypoints and one of
    //                  descriptors, width = the largest count of the chunk
    // The host drains chunk i-LAG after enqueueing chunk i, so LAG chunks are always queued ahead of the GPU.
    const int NL = c->lane1 ? 2 : 1, LAG = NL;
    // the upload of the first chunk is not hidden behind anything (a full first chunk of 32 float frames is 265 MB = 5 ms
    // exposed): see chunk_plan
    std::vector<int> cstart, csize;
    static const bool ramp_down = getenv("AKZ_RAMP_DOWN") != nullptr;      // tuning knob (scripts/probes/e2e_rampdown.sh): off, see chunk_plan
    // Float frames are bound by the host link (55.6 GB/s): what counts there is how soon the first kernel starts and how
    // little is left after the last byte, so their chunks stay at 32 frames however large max_batch is (256 frames, max_batch 64:
    // 5430 -> 5930 images/s).  u8 frames are bound by the kernels and take the full chunk.  AKZ_HOST_CHUNK_F32 overrides.
    static const int f32_cap = [] { const char* e = getenv("AKZ_HOST_CHUNK_F32"); const int v = e ? atoi(e) : 32; return v > 0 ? v : 32; }();
    const int Bh = (dtype == AKZ_F32 && !fast) ? std::min(B, f32_cap) : B;
    chunk_plan(nframes, Bh, true, ramp_down, cstart, csize);
    const int nchunks = (int)cstart.size();
    auto chunk_frames = [&](int i) { return csize[i]; };
    auto issue_h2d = [&](int i) -> int {
        const int s = i % AKZ_NSET;
        // staging[s] was last read by the kernels of chunk i - AKZ_NSET
        if (i >= AKZ_NSET) AKZ_CUDA_TRY(cudaStreamWaitEvent(c->h2d_stream, c->ev_comp[s], 0));
        const char* src = (const char*)h_images + (size_t)cstart[i] * stride * esz;
        AKZ_CUDA_TRY(cudaMemcpyAsync(c->img_stage[s], src, (size_t)chunk_frames(i) * stride * esz, cudaMemcpyHostToDevice, c->h2d_stream));
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_h2d[s], c->h2d_stream));
        return AKZ_OK;
    };
    auto drain = [&](int i) -> int {          // host side of the result download of chunk i
        const int s = i % AKZ_NSET, nf = chunk_frames(i), f0 = cstart[i];
        AKZ_CUDA_TRY(cudaEventSynchronize(c->ev_cnt[s]));
        int maxc = 0;
        for (int f = 0; f < nf; f++) { h_counts[f0 + f] = c->h_cnt_pinned[s * B + f]; maxc = std::max(maxc, h_counts[f0 + f]); }
        if (maxc > 0) {
            AKZ_CUDA_TRY(cudaMemcpy2DAsync(h_kpts + (size_t)f0 * MP, sizeof(akz_keypoint) * (size_t)MP, c->kpts_own + (size_t)s * B * MP,
                                           sizeof(akz_keypoint) * (size_t)MP, sizeof(akz_keypoint) * (size_t)maxc, nf, cudaMemcpyDeviceToHost, c->d2h_stream));
            if (describe)
                AKZ_CUDA_TRY(cudaMemcpy2DAsync(h_desc + (size_t)f0 * MP * 64, (size_t)MP * 64, c->desc_own + (size_t)s * B * MP * 64, (size_t)MP * 64,
                                               (size_t)maxc * 64, nf, cudaMemcpyDeviceToHost, c->d2h_stream));
        }
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_d2h[s], c->d2h_stream));
        return AKZ_OK;
    };
    for (int i = 0; i < std::min(LAG, nchunks); i++)
        if ((rc = issue_h2d(i)) != AKZ_OK) return rc;
    for (int i = 0; i < nchunks; i++) {
        const int s = i % AKZ_NSET, nf = chunk_frames(i);
        akz_ctx* L = (NL == 2 && (i & 1)) ? c->lane1 : c;
        cudaStream_t st = L->stream;
        AKZ_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_h2d[s], 0));
        if (i >= AKZ_NSET) AKZ_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_d2h[s], 0));   // result set s is free again
        if (fast) rc = fast_scale_space_chunk(L, (const unsigned char*)c->img_stage[s], nf, pitch, stride);
        else rc = scale_space_chunk(L, c->img_stage[s], dtype, nf, pitch, stride);
        if (rc != AKZ_OK) return rc;
        if ((rc = detect_chunk(L, nf, describe, c->counts_own + (size_t)s * B, c->kpts_own + (size_t)s * B * MP,
                               c->desc_own + (size_t)s * B * MP * 64, fast)) != AKZ_OK) return rc;
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_comp[s], st));
        if (i + LAG < nchunks && (rc = issue_h2d(i + LAG)) != AKZ_OK) return rc;
        // results of an earlier chunk first (the GPU is busy with later ones), so that they do not queue behind the wait below
        if (i >= LAG && (rc = drain(i - LAG)) != AKZ_OK) return rc;
        AKZ_CUDA_TRY(cudaStreamWaitEvent(c->d2h_stream, c->ev_comp[s], 0));
        AKZ_CUDA_TRY(cudaMemcpyAsync(c->h_cnt_pinned + s * B, c->counts_own + (size_t)s * B, sizeof(int) * nf, cudaMemcpyDeviceToHost, c->d2h_stream));
        AKZ_CUDA_TRY(cudaEventRecord(c->ev_cnt[s], c->d2h_stream));
    }
    for (int i = std::max(0, nchunks - LAG); i < nchunks; i++)
        if ((rc = drain(i)) != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaStreamSynchronize(c->d2h_stream));
    AKZ_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->lane1) AKZ_CUDA_TRY(cudaStreamSynchronize(c->lane1->stream));
    AKZ_CUDA_TRY(cudaGetLastError());
    return AKZ_OK;
}

int akz_detect_and_compute_host(akz_ctx* c, const void* h_images, int dtype, int nframes, int w, int h, int pitch, long long stride,
                                int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc)
{
    return detect_and_compute_host(c, h_images, dtype, nframes, w, h, pitch, stride, describe, h_counts, h_kpts, h_desc, 0);
}

int akz_fast_detect_and_compute_host(akz_ctx* c, const uint8_t* h_images, int nframes, int w, int h, int pitch, long long stride,
                                     int describe, int* h_counts, akz_keypoint* h_kpts, uint8_t* h_desc)
{
    return detect_and_compute_host(c, h_images, AKZ_U8, nframes, w, h, pitch, stride, describe, h_counts, h_kpts, h_desc, 1);
}

// ---- stage seams ------------------------------------------------------------------------------------------
#define STAGE_PROLOGUE()                                                    \
    if (!c) return akz_set_error(AKZ_E_INVALID, "null context");            \
    akz_device_guard dev_guard_;      /* the caller's current device is restored on return */ \
    AKZ_CUDA_TRY(cudaSetDevice(c->device))
#define STAGE_EPILOGUE() do { AKZ_CUDA_TRY(cudaGetLastError()); return AKZ_OK; } while (0)

int akz_lowpass(akz_ctx* c, const float* src, float* dst, int w, int h, int pitch, long long stride, int n, float var, int ksz)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_BLUR, akzk::lowpass(c->stream, src, dst, w, h, pitch, stride, pitch, stride, n, var, ksz));
    STAGE_EPILOGUE();
}

int akz_down_with_smooth(akz_ctx* c, const float* src, float* dst, float* smooth, int sw, int sh, int sp, long long sstride,
                         int dw, int dh, int dp, long long dstride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_BLUR, akzk::down_with_smooth(c->stream, src, dst, smooth, sw, sh, sp, sstride, dw, dh, dp, dstride, n));
    STAGE_EPILOGUE();
}

int akz_scharr_contrast(akz_ctx* c, const float* src, float* d_k, float per, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    if (n > c->opt.max_batch) return akz_set_error(AKZ_E_INVALID, "nframes exceeds max_batch");
    LAUNCHED(AKZ_K_CONTRAST, akzk::contrast(c->stream, src, c->hmax, c->hist, d_k, per, 0.f, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}

int akz_flow(akz_ctx* c, const float* src, float* flow, int type, const float* d_k, float kscale, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_FLOW, akzk::flow(c->stream, src, flow, type, d_k, kscale, 1, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}

int akz_nld_step(akz_ctx* c, const float* src, const float* flow, float* dst, float tau, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_FED, akzk::nld_step(c->stream, src, flow, dst, tau, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}

int akz_fed_cycle(akz_ctx* c, const float* src, const float* flow, float* dst, float* tmp, const float* tau, int nsteps,
                  int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    if (src == dst || src == tmp || dst == tmp) return akz_set_error(AKZ_E_INVALID, "src, dst and tmp must be distinct");
    LAUNCHED(AKZ_K_FED, akzk::fed_cycle(c->stream, src, flow, dst, tmp, tau, nsteps, w, h, pitch, stride, n, c->opt.fused));
    STAGE_EPILOGUE();
}

int akz_hessian(akz_ctx* c, const float* smooth, float* lx, float* ly, float* det, int step, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    if (c->opt.fused && (det == smooth || lx == smooth || ly == smooth))
        return akz_set_error(AKZ_E_INVALID, "the fused level kernel cannot work in place: outputs must not alias the input (use fused = 0)");
    if (c->opt.fused) LAUNCHED(AKZ_K_PREP, prep_level(c, 0, smooth, w, h, pitch, stride, nullptr, nullptr, lx, ly, det, 0, step, w, h, pitch, stride, n));
    else LAUNCHED(AKZ_K_HESSIAN, akzk::hessian(c->stream, smooth, lx, ly, det, step, w, h, pitch, stride, n));
    {
        const int rj = join_rings(c);
        if (rj != AKZ_OK) return rj;
    }
    STAGE_EPILOGUE();
}

// ---- integer pipeline stage seams ---------------------------------------------------------------------------------------
int akz_fast_lowpass(akz_ctx* c, const void* src, int src_is_u8, int* dst, int* tmp, int w, int h, int pitch, long long stride, int n, float var, int ksz)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_BLUR, akzk::fast_lowpass(c->stream, src, src_is_u8, dst, tmp, w, h, pitch, stride, pitch, stride, n, var, ksz));
    STAGE_EPILOGUE();
}
int akz_fast_down_with_smooth(akz_ctx* c, const int* src, int* dst, int* smooth, int sw, int sh, int sp, long long sstride,
                              int dw, int dh, int dp, long long dstride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_BLUR, akzk::fast_down(c->stream, src, dst, smooth, sw, sh, sp, sstride, dw, dh, dp, dstride, n));
    STAGE_EPILOGUE();
}
int akz_fast_scharr_contrast(akz_ctx* c, const int* src, int* mag, int* d_k, float per, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    if (n > c->opt.max_batch) return akz_set_error(AKZ_E_INVALID, "nframes exceeds max_batch");
    LAUNCHED(AKZ_K_CONTRAST, akzk::fast_contrast(c->stream, src, mag, (int*)c->hmax, c->hist, d_k, per, 0, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}
int akz_fast_flow(akz_ctx* c, const int* src, int* flow, int type, const int* d_k, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_FLOW, akzk::fast_flow(c->stream, src, flow, type, d_k, 0, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}
int akz_fast_nld_step(akz_ctx* c, const int* src, const int* flow, int* dst, float tau, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_FED, akzk::fast_nld_step(c->stream, src, flow, dst, tau, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}
int akz_fast_hessian(akz_ctx* c, const int* smooth, int* lx, int* ly, int* det, int step, int w, int h, int pitch, long long stride, int n)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_HESSIAN, akzk::fast_hessian(c->stream, smooth, lx, ly, det, step, w, h, pitch, stride, n));
    STAGE_EPILOGUE();
}

int akz_orient(akz_ctx* c, const int* d_counts, akz_keypoint* d_kpts, int n)
{
    STAGE_PROLOGUE();
    if (n < 1 || n > c->opt.max_batch || c->nlev == 0) return akz_set_error(AKZ_E_INVALID, "bad frame count");
    LAUNCHED(AKZ_K_MISC, akzk::frame_prefix(c->stream, d_counts, c->prefix, n));
    LAUNCHED(AKZ_K_ORIENT, akzk::orient(c->stream, c->tab, d_counts, c->prefix, d_kpts, c->opt.max_pts, n));
    STAGE_EPILOGUE();
}

int akz_describe(akz_ctx* c, const int* d_counts, const akz_keypoint* d_kpts, uint8_t* d_desc, int n)
{
    STAGE_PROLOGUE();
    if (n < 1 || n > c->opt.max_batch || c->nlev == 0) return akz_set_error(AKZ_E_INVALID, "bad frame count");
    LAUNCHED(AKZ_K_MISC, akzk::frame_prefix(c->stream, d_counts, c->prefix, n));
    LAUNCHED(AKZ_K_DESCRIBE, akzk::describe(c->stream, c->tab, d_counts, c->prefix, d_kpts, d_desc, c->opt.max_pts, n, c->opt.descriptor_pattern_size));
    STAGE_EPILOGUE();
}

int akz_detect_keypoints(akz_ctx* c, int n, int* d_counts, akz_keypoint* d_kpts)
{
    STAGE_PROLOGUE();
    if (n < 1 || n > c->opt.max_batch || c->nlev == 0) return akz_set_error(AKZ_E_INVALID, "bad frame count");
    int rc = detect_chunk(c, n, 0, d_counts, d_kpts, nullptr, 0, false);      // the planes may have been written by the caller
    if (rc != AKZ_OK) return rc;
    STAGE_EPILOGUE();
}

// ---- matcher ----------------------------------------------------------------------------------------------
static std::atomic<int> g_match_kernel{0};   // 0 = by problem size, 1 = POPC/LOP3 kernel, 2 = mma.sync kernel, 3 = tcgen05 kernel
void akz_set_match_kernel(int which) { g_match_kernel.store(which, std::memory_order_relaxed); }
void akz_set_describe_kernel(int which) { akzk::set_describe_kernel(which); }

int akz_match(akz_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int t_index_base, int mode, int finalize, akz_match_t* d_out)
{
    STAGE_PROLOGUE();
    if (mode != AKZ_MATCH_COMPAT && mode != AKZ_MATCH_KNN2 && mode != AKZ_MATCH_UNIQUE2) return akz_set_error(AKZ_E_INVALID, "bad matcher mode");
    if (nq < 0 || nt < 0 || !d_out) return akz_set_error(AKZ_E_INVALID, "bad matcher arguments");
    if (nq == 0) return AKZ_OK;
    // Kernel choice (akz_set_match_kernel overrides): small problems -> LOP3/POPC kernel (2000 x 2300: 29 us against 55 us for
    // the tensor-core kernels); large ones -> tcgen05 kernel (match_tc5.cu); 2 selects the legacy mma.sync kernel.
    const bool large = (long long)nq * nt >= (1ll << 24) && nq >= 1024;
    const int forced = g_match_kernel.load(std::memory_order_relaxed);
    int kern = forced == 0 ? (large ? 3 : 1) : forced;
    const int tc5_filter = kern == 4 ? 1 : kern == 5 ? 0 : -1;               // 4 / 5: tcgen05 kernel with the chunk filter forced on / off (tests)
    if (kern > 3) kern = 3;
    const int use_mma = kern == 2;
    const int qtile = 256;                           // queries per block of every kernel
    int qblocks = (nq + qtile - 1) / qtile;
    int nsplit = std::max(1, std::min((8 * 148 + qblocks - 1) / qblocks, (nt + 127) / 128));
    if (kern >= 2) {
        // one block per SM: pick the split whose block count fills whole waves of 148 best, with at least ~4 train tiles
        // per block so the query-tile expansion is amortised
        const int max_split = std::max(1, std::min((nt + 127) / 128 / 4, 64));
        double best = -1.0;
        for (int sp = 1; sp <= max_split; sp++) {
            long long blocks = (long long)qblocks * sp;
            double eff = (double)blocks / (double)(((blocks + 147) / 148) * 148);
            if (blocks < 148) eff *= 0.5;
            if (eff > best + 1e-9) { best = eff; nsplit = sp; }
        }
    }
    if (kern == 3) {
        // k_match_tc5 balances the tiles over one CTA per SM itself; it reports how many partial results per query it writes
        int sp, T, S, grid;
        nsplit = akzk::match_tc5_plan(nq, nt, &sp, &T, &S, &grid);
        if (nsplit < 0) return nsplit;
    }
    size_t need = (size_t)nsplit * nq;
    if (c->match_parts_n < need) {
        if (c->match_parts) { cudaStreamSynchronize(c->stream); cudaFree(c->match_parts); }
        c->match_parts = nullptr; c->match_parts_n = 0;
        AKZ_CUDA_TRY(cudaMalloc((void**)&c->match_parts, need * sizeof(akz_match_t)));
        c->match_parts_n = need;
    }
    if (kern == 3) { LAUNCHED(AKZ_K_MATCH, akzk::match_partial_tc5(c->stream, d_q, nq, d_t, nt, t_index_base, mode, c->match_parts, tc5_filter)); }
    else { LAUNCHED(AKZ_K_MATCH, akzk::match_partial(c->stream, d_q, nq, d_t, nt, t_index_base, mode, nsplit, c->match_parts, use_mma)); }
    LAUNCHED(AKZ_K_MATCH, akzk::match_merge(c->stream, c->match_parts, nsplit, nq, mode, finalize, d_out));
    STAGE_EPILOGUE();
}

int akz_match_pairs(akz_ctx* c, const uint8_t* d_desc, const int* d_counts, int nframes, int mode, akz_match_t* d_out)
{
    STAGE_PROLOGUE();
    if (mode != AKZ_MATCH_COMPAT && mode != AKZ_MATCH_KNN2 && mode != AKZ_MATCH_UNIQUE2) return akz_set_error(AKZ_E_INVALID, "bad matcher mode");
    if (!d_desc || !d_counts || !d_out || nframes < 0) return akz_set_error(AKZ_E_INVALID, "bad matcher arguments");
    if (nframes < 2) return AKZ_OK;
    const int npairs = nframes - 1, mp = c->opt.max_pts;
    const int nsplit = npairs >= 16 ? 2 : npairs >= 4 ? 4 : 8;        // enough blocks for the 148 SMs when there are few pairs
    const size_t need = (size_t)npairs * nsplit * mp;
    if (c->match_parts_n < need) {
        if (c->match_parts) { cudaStreamSynchronize(c->stream); cudaFree(c->match_parts); }
        c->match_parts = nullptr; c->match_parts_n = 0;
        AKZ_CUDA_TRY(cudaMalloc((void**)&c->match_parts, need * sizeof(akz_match_t)));
        c->match_parts_n = need;
    }
    LAUNCHED(AKZ_K_MATCH, akzk::match_pairs(c->stream, d_desc, d_counts, nframes, mp, mode, nsplit, c->match_parts, d_out));
    STAGE_EPILOGUE();
}

// ---- train-sharded matching: NCCL through dlopen ----------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, akz_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};
std::mutex g_nccl_mutex;
NcclApi g_nccl;

// dlopen by soname: glibc returns the copy already mapped into the process (e.g. the one a framework brought) before it
// searches the library path, so a communicator handed to akz_comm_attach and the calls below go to the same NCCL
int nccl_api(NcclApi** out)
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!g_nccl.lib) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return akz_set_error(AKZ_E_UNSUPPORTED, "NCCL is not available: %s", dlerror());
        NcclApi a;
        a.lib = h;
        a.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(void**, int, akz_nccl_id, int))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
        a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        a.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
            return akz_set_error(AKZ_E_UNSUPPORTED, "libnccl.so.2 lacks an expected entry point");
        g_nccl = a;
    }
    *out = &g_nccl;
    return AKZ_OK;
}
int nccl_fail(NcclApi* n, int code, const char* what) { return akz_set_error(AKZ_E_CUDA, "NCCL error %d (%s) in %s", code, n->GetErrorString(code), what); }
}  // namespace

int akz_comm_unique_id(void* id128)
{
    if (!id128) return akz_set_error(AKZ_E_INVALID, "null argument");
    NcclApi* n;
    int rc = nccl_api(&n);
    if (rc != AKZ_OK) return rc;
    static_assert(sizeof(akz_nccl_id) == AKZ_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    int e = n->GetUniqueId(id128);
    return e == 0 ? AKZ_OK : nccl_fail(n, e, "ncclGetUniqueId");
}

int akz_comm_init(akz_ctx* c, int nranks, int rank, const void* id128)
{
    STAGE_PROLOGUE();
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return akz_set_error(AKZ_E_INVALID, "bad communicator arguments");
    NcclApi* n;
    int rc = nccl_api(&n);
    if (rc != AKZ_OK) return rc;
    akz_comm_destroy(c);
    akz_nccl_id id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    int e = n->CommInitRank(&comm, nranks, id, rank);
    if (e != 0) return nccl_fail(n, e, "ncclCommInitRank");
    c->comm = comm; c->comm_ranks = nranks; c->comm_rank = rank; c->comm_owned = true;
    return AKZ_OK;
}

int akz_comm_attach(akz_ctx* c, void* nccl_comm, int nranks, int rank)
{
    if (!c || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return akz_set_error(AKZ_E_INVALID, "bad communicator arguments");
    NcclApi* n;
    int rc = nccl_api(&n);
    if (rc != AKZ_OK) return rc;
    akz_comm_destroy(c);
    c->comm = nccl_comm; c->comm_ranks = nranks; c->comm_rank = rank; c->comm_owned = false;
    return AKZ_OK;
}

int akz_comm_destroy(akz_ctx* c)
{
    if (!c) return akz_set_error(AKZ_E_INVALID, "null context");
    if (c->comm && c->comm_owned && g_nccl.CommDestroy) {
        if (c->stream) cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->comm);
    }
    c->cur = nullptr; c->opar = false; c->graph_ok = false;
    for (int i = 0; i < 8; i++) { c->ostream[i] = nullptr; c->ev_lvl0[i] = c->ev_oct[i] = nullptr; c->sc[i] = { nullptr, nullptr, nullptr, nullptr }; }
    for (int i = 0; i < 8; i++) { c->ring_stream[i] = nullptr; c->ev_ring_fork[i] = c->ev_ring_join[i] = nullptr; c->ring_pending[i] = false; }
    c->rk = 0;
    c->comm = nullptr; c->comm_ranks = 0; c->comm_rank = 0; c->comm_owned = false;
    return AKZ_OK;
}

int akz_match_sharded(akz_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_t_local, int nt_local, int t_index_base, int mode, akz_match_t* d_out)
{
    STAGE_PROLOGUE();
    if (!c->comm) return akz_set_error(AKZ_E_INVALID, "no communicator: call akz_comm_init or akz_comm_attach first");
    if (nq < 0 || nt_local < 0 || !d_out) return akz_set_error(AKZ_E_INVALID, "bad matcher arguments");
    if (nq == 0) return AKZ_OK;
    NcclApi* n;
    int rc = nccl_api(&n);
    if (rc != AKZ_OK) return rc;
    const size_t need = (size_t)(1 + c->comm_ranks) * nq;
    if (c->shard_buf_n < need) {
        if (c->shard_buf) { cudaStreamSynchronize(c->stream); cudaFree(c->shard_buf); }
        c->shard_buf = nullptr; c->shard_buf_n = 0;
        AKZ_CUDA_TRY(cudaMalloc((void**)&c->shard_buf, need * sizeof(akz_match_t)));
        c->shard_buf_n = need;
    }
    akz_match_t* mine = c->shard_buf;
    akz_match_t* all = c->shard_buf + nq;
    // local range -> associative partial form (KNN2: two best (distance, index); COMPAT: best + the mask of index classes)
    if ((rc = akz_match(c, d_q, nq, d_t_local, nt_local, t_index_base, mode, 0, mine)) != AKZ_OK) return rc;
    // the one exchange of the path: nq x 16 bytes per rank, stream-ordered
    int e = n->AllGather(mine, all, (size_t)nq * sizeof(akz_match_t), 0 /* ncclInt8 / ncclChar */, c->comm, c->stream);
    if (e != 0) return nccl_fail(n, e, "ncclAllGather");
    LAUNCHED(AKZ_K_MATCH, akzk::match_merge(c->stream, all, c->comm_ranks, nq, mode, 1, d_out));
    STAGE_EPILOGUE();
}

int akz_plan_chunks(int nframes, int max_batch, int ramp_up, int* starts, int* sizes, int cap)
{
    if (nframes < 0 || max_batch < 1 || !starts || !sizes) return akz_set_error(AKZ_E_INVALID, "bad arguments");
    std::vector<int> st, sz;
    chunk_plan(nframes, max_batch, ramp_up != 0, false, st, sz);
    for (size_t i = 0; i < st.size() && (int)i < cap; i++) { starts[i] = st[i]; sizes[i] = sz[i]; }
    return (int)st.size();
}

int akz_plan_match(int nq, int nt, int* out5)
{
    if (nq < 1 || nt < 0 || !out5) return akz_set_error(AKZ_E_INVALID, "bad arguments");
    int nsplit, T, S, grid;
    const int nparts = akzk::match_tc5_plan(nq, nt, &nsplit, &T, &S, &grid);
    if (nparts < 0) return nparts;
    out5[0] = nparts; out5[1] = nsplit; out5[2] = T; out5[3] = S; out5[4] = grid;
    return AKZ_OK;
}

int akz_match_merge(akz_ctx* c, const akz_match_t* d_parts, int nparts, int nq, int mode, int finalize, akz_match_t* d_out)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_MATCH, akzk::match_merge(c->stream, d_parts, nparts, nq, mode, finalize, d_out));
    STAGE_EPILOGUE();
}

int akz_match_host(akz_ctx* c, const uint8_t* h_q, int nq, const uint8_t* h_t, int nt, int mode, akz_match_t* h_out)
{
    STAGE_PROLOGUE();
    if (nq <= 0) return AKZ_OK;
    size_t need = (size_t)(nq + nt) * 64 + (size_t)nq * sizeof(akz_match_t);
    if (c->match_stage_bytes < need) {
        if (c->match_stage) { cudaStreamSynchronize(c->stream); cudaFree(c->match_stage); }
        c->match_stage = nullptr; c->match_stage_bytes = 0;
        AKZ_CUDA_TRY(cudaMalloc(&c->match_stage, need));
        c->match_stage_bytes = need;
    }
    uint8_t* dq = (uint8_t*)c->match_stage;
    uint8_t* dt = dq + (size_t)nq * 64;
    akz_match_t* dm = (akz_match_t*)(dt + (size_t)nt * 64);
    AKZ_CUDA_TRY(cudaMemcpyAsync(dq, h_q, (size_t)nq * 64, cudaMemcpyHostToDevice, c->stream));
    if (nt > 0) AKZ_CUDA_TRY(cudaMemcpyAsync(dt, h_t, (size_t)nt * 64, cudaMemcpyHostToDevice, c->stream));
    int rc = akz_match(c, dq, nq, dt, nt, 0, mode, 1, dm);
    if (rc != AKZ_OK) return rc;
    AKZ_CUDA_TRY(cudaMemcpyAsync(h_out, dm, (size_t)nq * sizeof(akz_match_t), cudaMemcpyDeviceToHost, c->stream));
    return akz_sync(c);
}

// ---- per-kernel-class device timing ------------------------------------------------------------------------------
static const char* const k_class_names[AKZ_NUM_KCLASS] = { "base", "blur", "contrast", "prep", "hessian", "flow", "fed", "extrema", "nms",
                                                           "orient", "describe", "match", "misc" };
const char* akz_profile_class_name(int cls) { return (cls >= 0 && cls < AKZ_NUM_KCLASS) ? k_class_names[cls] : ""; }

int akz_profile_enable(akz_ctx* c, int on)
{
    if (!c) return akz_set_error(AKZ_E_INVALID, "null context");
    // A kernel's time is its own only when nothing runs beside it: while profiling, contexts of more than 4 frames put the octave
    // chains back on the context's stream (small-batch contexts keep theirs: their per-class times are read as a share).
    for (akz_ctx* x : { c, c->lane1 }) {
        if (!x) continue;
        if (on && !x->prof_on && x->opt.max_batch > 4) { x->opar_saved = x->opar; x->opar = false; }
        if (!on && x->prof_on && x->opt.max_batch > 4) x->opar = x->opar_saved;
        x->prof_on = on != 0;
    }
    return AKZ_OK;
}

int akz_profile_read(akz_ctx* c, int ncls, double* ms, long long* launches)
{
    if (!c || !ms || !launches) return akz_set_error(AKZ_E_INVALID, "null argument");
    akz_device_guard dev_guard_;
    AKZ_CUDA_TRY(cudaSetDevice(c->device));
    AKZ_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (auto& pp : c->prof_pairs) {
        float t = 0.f;
        AKZ_CUDA_TRY(cudaEventElapsedTime(&t, pp.a, pp.b));
        c->prof_ms[pp.cls] += t; c->prof_launches[pp.cls] += pp.launches; c->prof_oct_ms[pp.cls][pp.oct & 7] += t;
        c->prof_pool.push_back(pp.a); c->prof_pool.push_back(pp.b);
    }
    c->prof_pairs.clear();
    for (int i = 0; i < ncls && i < AKZ_NUM_KCLASS; i++) { ms[i] = c->prof_ms[i]; launches[i] = c->prof_launches[i]; }
    memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_launches, 0, sizeof(c->prof_launches));
    memcpy(c->prof_oct_last, c->prof_oct_ms, sizeof(c->prof_oct_ms)); memset(c->prof_oct_ms, 0, sizeof(c->prof_oct_ms));
    if (c->lane1) {
        // with two lanes the per-class times are summed over both streams (they overlap in wall-clock time)
        double ms1[AKZ_NUM_KCLASS]; long long l1[AKZ_NUM_KCLASS];
        int rc = akz_profile_read(c->lane1, AKZ_NUM_KCLASS, ms1, l1);
        if (rc != AKZ_OK) return rc;
        for (int i = 0; i < ncls && i < AKZ_NUM_KCLASS; i++) { ms[i] += ms1[i]; launches[i] += l1[i]; }
        for (int i = 0; i < AKZ_NUM_KCLASS; i++)
            for (int k = 0; k < 8; k++) c->prof_oct_last[i][k] += c->lane1->prof_oct_last[i][k];
    }
    return AKZ_OK;
}

// the times of the last akz_profile_read split by octave: ms[cls * noct + octave] (scale-space classes; everything else is octave 0)
int akz_profile_octaves(akz_ctx* c, int ncls, int noct, double* ms)
{
    if (!c || !ms || ncls < 0 || noct < 0) return akz_set_error(AKZ_E_INVALID, "bad argument");
    for (int i = 0; i < ncls; i++)
        for (int k = 0; k < noct; k++) ms[i * noct + k] = (i < AKZ_NUM_KCLASS && k < 8) ? c->prof_oct_last[i][k] : 0.0;
    return AKZ_OK;
}

// ---- host format: OpenCV conventions --------------------------------------------------------------
// cv::KeyPoint fields for keypoints already on the host: pt = (x, y), size = derivative scale in FULL-resolution pixels
// (size * 2^octave), angle in degrees, response, octave, class_id = sublevel.  out: n rows of 7 floats.
int akz_keypoints_to_opencv(const akz_keypoint* h_kpts, int n, int max_scale, float* out)
{
    if (!h_kpts || !out || n < 0 || max_scale < 1) return akz_set_error(AKZ_E_INVALID, "bad argument");
    for (int i = 0; i < n; i++) {
        const akz_keypoint& k = h_kpts[i];
        int oct = k.layer / max_scale, sub = k.layer - oct * max_scale;
        float* o = out + 7 * (size_t)i;
        o[0] = k.x; o[1] = k.y; o[2] = k.size * (float)(1 << oct); o[3] = k.angle * (float)(180.0 / M_PI);
        o[4] = k.response; o[5] = (float)oct; o[6] = (float)sub;
    }
    return AKZ_OK;
}
// cv::DMatch fields for accepted matches: out rows (queryIdx, trainIdx, distance); returns the number of rows written
int akz_matches_to_opencv(const akz_match_t* h_m, int nq, int* out)
{
    if (!h_m || !out || nq < 0) return akz_set_error(AKZ_E_INVALID, "bad argument");
    int n = 0;
    for (int i = 0; i < nq; i++)
        if (h_m[i].idx1 >= 0) { out[3 * n] = i; out[3 * n + 1] = h_m[i].idx1; out[3 * n + 2] = h_m[i].dist1; n++; }
    return n;
}

// ---- AoS bridge -----------------------------------------------------------------------------------------------
int akz_pack_points(akz_ctx* c, const int* d_count, const akz_keypoint* d_kpts, const uint8_t* d_desc, void* d_points, int max_pts, int with_desc)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_MISC, akzk::pack_points(c->stream, d_count, d_kpts, d_desc, d_points, max_pts, with_desc));
    STAGE_EPILOGUE();
}
int akz_unpack_desc(akz_ctx* c, const void* d_points, int n, uint8_t* d_desc)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_MISC, akzk::unpack_desc(c->stream, d_points, n, d_desc));
    STAGE_EPILOGUE();
}
int akz_scatter_matches(akz_ctx* c, const akz_match_t* d_m, int nq, void* d_points_q, const void* d_points_t)
{
    STAGE_PROLOGUE();
    LAUNCHED(AKZ_K_MISC, akzk::scatter_matches(c->stream, d_m, nq, d_points_q, d_points_t));
    STAGE_EPILOGUE();
}

}  // extern "C"
