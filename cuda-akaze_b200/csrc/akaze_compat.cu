// C++ drop-in surface of the reference (akaze.h / akazed.h / fed.h) implemented over the C ABI.
// Nothing here computes: it translates the reference's calling convention (synchronous calls, AoS
// AkazePoint records, print-and-exit error handling) into akz_* calls.
#include "../../include/akaze.h"
#include "../../include/akazed.h"
#include "../../include/fed.h"
#include "../../include/akaze_b200.h"
#include <cstring>
#include <cmath>
#include <mutex>

namespace {

[[noreturn]] void die(const char* where)
{
    fprintf(stderr, "akaze_b200: %s failed: %s\n", where, akz_last_error());
    exit(-1);                                   // the reference's convention (cuda_utils.h:18-37)
}
#define AKZ_DO(call) do { if ((call) != AKZ_OK) die(#call); } while (0)

[[noreturn]] void not_routed(const char* name)
{
    fprintf(stderr, "akaze_b200: %s is declared for source compatibility only; the B200 pipeline keeps keypoints "
                    "in SoA buffers and never materialises the reference's dense maps / pyramid layout. "
                    "Use akaze::Akazer or the C ABI (akaze_b200.h).\n", name);
    exit(-1);
}

// One lazily created context PER DEVICE for the matcher and the synchronous stage functions (unfused kernels:
// hHessianDeterminant writes the determinant over its input, which needs separate passes).  The reference runs on
// whatever device is current (cuda_utils.h initDevice), so the context is looked up by the current device; creation
// is serialised, and callers that share a device share the context the way the reference's callers share the default stream.
constexpr int kMaxDevices = 64;
std::mutex g_stage_mutex;
akz_ctx* g_stage_ctx[kMaxDevices] = {};

int current_device()
{
    int d = 0;
    CHECK(cudaGetDevice(&d));
    return d & (kMaxDevices - 1);
}

akz_ctx* stage_ctx()
{
    const int d = current_device();
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (!g_stage_ctx[d]) {
        akz_options o;
        akz_default_options(&o);
        o.width = 0; o.height = 0; o.fused = 0; o.max_batch = 1; o.device = d;
        AKZ_DO(akz_create(&o, &g_stage_ctx[d]));
    }
    return g_stage_ctx[d];
}

// scratch device buffers of the shims, one set per device (grown on demand, never shrunk)
struct Scratch {
    void* p[kMaxDevices] = {}; size_t n[kMaxDevices] = {};
    void* get(size_t bytes)
    {
        const int d = current_device();
        std::lock_guard<std::mutex> lock(g_stage_mutex);
        if (n[d] < bytes) { if (p[d]) cudaFree(p[d]); CHECK(cudaMalloc(&p[d], bytes)); n[d] = bytes; }
        return p[d];
    }
};
Scratch g_match_scratch, g_k_scratch, g_tmp_scratch;

void match_points(akaze::AkazeData& r1, akaze::AkazeData& r2)
{
    akz_ctx* c = stage_ctx();
    const int n1 = r1.num_pts, n2 = r2.num_pts;
    if (n1 <= 0) return;                         // the reference would launch an empty grid and exit (App. B-9)
    size_t bytes = (size_t)(n1 + n2) * 64 + (size_t)n1 * sizeof(akz_match_t);
    uint8_t* dq = (uint8_t*)g_match_scratch.get(bytes);
    uint8_t* dt = dq + (size_t)n1 * 64;
    akz_match_t* dm = (akz_match_t*)(dt + (size_t)n2 * 64);
    AKZ_DO(akz_unpack_desc(c, r1.d_data, n1, dq));
    AKZ_DO(akz_unpack_desc(c, r2.d_data, n2, dt));
    AKZ_DO(akz_match(c, dq, n1, dt, n2, 0, AKZ_MATCH_COMPAT, 1, dm));
    AKZ_DO(akz_scatter_matches(c, dm, n1, r1.d_data, r2.d_data));
    AKZ_DO(akz_sync(c));
}

}  // namespace

// ---- fed.h -------------------------------------------------------------------------------------------
int fed_tau_by_process_time(const float T, const int M, const float tau_max, const bool reordering, std::vector<float>& tau)
{
    float buf[4096];
    int n = akz_fed_tau(T, M, tau_max, reordering ? 1 : 0, buf, 4096);
    if (n <= 0) { tau.clear(); return 0; }
    tau.assign(buf, buf + n);
    return n;
}
int fed_tau_by_cycle_time(const float t, const float tau_max, const bool reordering, std::vector<float>& tau)
{
    return fed_tau_by_process_time(t, 1, tau_max, reordering, tau);
}
int fed_tau_internal(const int n, const float scale, const float tau_max, const bool reordering, std::vector<float>& tau)
{
    if (n <= 0) return 0;
    tau.assign((size_t)n, 0.f);
    const int m = akz_fed_tau_internal(n, scale, tau_max, reordering ? 1 : 0, tau.data(), n);
    if (m <= 0) { tau.clear(); return 0; }
    return m;
}
bool fed_is_prime_internal(const int number)
{
    if (number <= 1) return false;
    for (int d = 2; (long long)d * d <= number; d++) if (number % d == 0) return false;
    return true;
}

// ---- akazed.h: global setters (the B200 pipeline keeps this state per context; these keep callers linking)
namespace { unsigned int* g_counter = nullptr; unsigned int* g_maxc = nullptr; }
void setMaxNumPoints(const int) {}
void getPointCounter(void** addr)
{
    if (!g_counter) { CHECK(cudaMalloc((void**)&g_counter, sizeof(unsigned int))); CHECK(cudaMemset(g_counter, 0, sizeof(unsigned int))); }
    *addr = g_counter;
}
void getMaxContrastAddr(void** addr)
{
    if (!g_maxc) { CHECK(cudaMalloc((void**)&g_maxc, sizeof(unsigned int))); CHECK(cudaMemset(g_maxc, 0, sizeof(unsigned int))); }
    *addr = g_maxc;
}
void setHistogram(const int*) {}
void setExtremaParam(const float*, const int) {}
void setOparam(const int*, const int) {}
void setCompareIndices() {}

namespace akaze
{
    // ---- akaze.h -------------------------------------------------------------------------------------
    void initAkazeData(AkazeData& data, const int max_pts, const bool host, const bool dev)
    {
        data.num_pts = 0;
        data.max_pts = max_pts;
        const size_t bytes = sizeof(AkazePoint) * (size_t)max_pts;
        data.h_data = host ? (AkazePoint*)malloc(bytes) : NULL;
        data.d_data = NULL;
        if (dev) {
            CHECK(cudaMalloc((void**)&data.d_data, bytes));
            CHECK(cudaMemset(data.d_data, 0, bytes));       // the reference leaves the padding bytes undefined (App. B-6)
        }
    }

    void freeAkazeData(AkazeData& data)
    {
        if (data.d_data != NULL) CHECK(cudaFree(data.d_data));
        if (data.h_data != NULL) free(data.h_data);
        data.d_data = NULL; data.h_data = NULL;
        data.num_pts = 0; data.max_pts = 0;
    }

    void cuMatch(AkazeData& result1, AkazeData& result2)
    {
        match_points(result1, result2);
        if (result1.h_data && result1.num_pts > 0) {
            // match, distance, match_x, match_y: 16 bytes per record (akaze.cpp:58-63)
            CHECK(cudaMemcpy2D(&result1.h_data[0].match, sizeof(AkazePoint), &result1.d_data[0].match, sizeof(AkazePoint),
                               4 * sizeof(float), result1.num_pts, cudaMemcpyDeviceToHost));
        }
    }

    struct Akazer::State {
        akz_options opt;
        akz_ctx* ctx = nullptr;
        int* d_count = nullptr;
        akz_keypoint* d_kpts = nullptr;
        uint8_t* d_desc = nullptr;
        int* h_count = nullptr;                 // pinned landing zone of the keypoint count
        unsigned char* h_stage = nullptr;       // pinned landing zone of the AkazePoint records (cap x 104 bytes)
        int cap = 0;
        int last_n = 0;                         // keypoints of the previous frame (size of the speculative record copy)
        void release()
        {
            if (ctx) akz_destroy(ctx);
            if (d_count) cudaFree(d_count);
            if (d_kpts) cudaFree(d_kpts);
            if (d_desc) cudaFree(d_desc);
            if (h_count) cudaFreeHost(h_count);
            if (h_stage) cudaFreeHost(h_stage);
            ctx = nullptr; d_count = nullptr; d_kpts = nullptr; d_desc = nullptr; h_count = nullptr; h_stage = nullptr; cap = 0;
        }
        void ensure(int w, int h, int max_pts)
        {
            if (ctx && opt.width == w && opt.height == h && cap == max_pts) return;
            release();
            opt.width = w; opt.height = h; opt.max_pts = max_pts; opt.max_batch = 1;
            AKZ_DO(akz_create(&opt, &ctx));
            CHECK(cudaMalloc((void**)&d_count, sizeof(int)));
            CHECK(cudaMalloc((void**)&d_kpts, sizeof(akz_keypoint) * (size_t)max_pts));
            CHECK(cudaMalloc((void**)&d_desc, (size_t)64 * max_pts));
            CHECK(cudaMallocHost((void**)&h_count, sizeof(int)));
            CHECK(cudaMallocHost((void**)&h_stage, sizeof(AkazePoint) * (size_t)max_pts));
            cap = max_pts;
        }
        // One frame through the batched pipeline (a context of batch 1: octave chains on their own streams, the whole chunk
        // replayed as a CUDA graph from the second call with the same buffers on), then the reference's result contract
        // (akaze.cpp:128-139): num_pts, AkazePoint records in result.d_data, and the first 24 (+61) bytes of every record in
        // result.h_data.  The host copy is one contiguous transfer of the records into pinned memory and a strided copy on the
        // host (the reference's cudaMemcpy2D of 85-byte rows is a row-by-row transfer).
        void run(const void* image, int dtype, AkazeData& result, int3 whp0, bool desc, bool fast = false)
        {
            ensure(whp0.x, whp0.y, result.max_pts);
            cudaStream_t st = (cudaStream_t)akz_stream(ctx);
            if (fast)
                AKZ_DO(akz_fast_detect_and_compute(ctx, (const uint8_t*)image, 1, whp0.x, whp0.y, whp0.z, (long long)whp0.y * whp0.z,
                                                   desc ? 1 : 0, d_count, d_kpts, d_desc));
            else
                AKZ_DO(akz_detect_and_compute(ctx, image, dtype, 1, whp0.x, whp0.y, whp0.z, (long long)whp0.y * whp0.z,
                                              desc ? 1 : 0, d_count, d_kpts, d_desc));
            AKZ_DO(akz_pack_points(ctx, d_count, d_kpts, d_desc, result.d_data, result.max_pts, desc ? 1 : 0));
            CHECK(cudaMemcpyAsync(h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            // The records follow the count in the same stream: as many as the previous frame had (+ 25 %), so that a stream of
            // similar frames needs one synchronisation per frame, not two; the rest, if any, in a second copy.
            const bool want = result.h_data != NULL;
            const size_t guess = want ? std::min((size_t)result.max_pts, (size_t)last_n + (size_t)last_n / 4 + 256) : 0;
            if (guess) CHECK(cudaMemcpyAsync(h_stage, result.d_data, guess * sizeof(AkazePoint), cudaMemcpyDeviceToHost, st));
            AKZ_DO(akz_sync(ctx));
            result.num_pts = *h_count;
            last_n = result.num_pts;
            if (want && result.num_pts > 0) {
                const size_t n = (size_t)result.num_pts;
                const size_t ncopy = (desc ? FLEN * sizeof(unsigned char) : 0) + 6 * sizeof(float);      // akaze.cpp:134-139
                if (n > guess) {
                    CHECK(cudaMemcpyAsync(h_stage + guess * sizeof(AkazePoint), result.d_data + guess, (n - guess) * sizeof(AkazePoint), cudaMemcpyDeviceToHost, st));
                    AKZ_DO(akz_sync(ctx));
                }
                unsigned char* dst = reinterpret_cast<unsigned char*>(result.h_data);
                for (size_t i = 0; i < n; i++) memcpy(dst + i * sizeof(AkazePoint), h_stage + i * sizeof(AkazePoint), ncopy);
            }
        }
    };

    Akazer::Akazer() : state(new State)
    {
        akz_default_options(&state->opt);
    }

    Akazer::~Akazer()
    {
        state->release();
        delete state;
    }

    void Akazer::init(int3 whp0, int _noctaves, int _max_scale, float _per, float _kcontrast, float _soffset, bool _reordering,
        float _derivative_factor, float _dthreshold, int _diffusivity, int _descriptor_pattern_size)
    {
        akz_options& o = state->opt;
        o.width = whp0.x; o.height = whp0.y;
        o.noctaves = _noctaves; o.max_scale = _max_scale; o.per = _per; o.kcontrast = _kcontrast; o.soffset = _soffset;
        o.reordering = _reordering ? 1 : 0; o.derivative_factor = _derivative_factor; o.dthreshold = _dthreshold;
        o.diffusivity = _diffusivity; o.descriptor_pattern_size = _descriptor_pattern_size;
        if (state->ctx) state->release();            // options changed: the context is rebuilt on the next call
    }

    void Akazer::detectAndCompute(float* image, AkazeData& result, int3 whp0, const bool desc)
    {
        state->run(image, AKZ_F32, result, whp0, desc);
    }

    // Integer entry point of the reference (akaze.cpp:153-201): the 16.16 fixed-point pipeline (fast_pipeline.cu)
    void Akazer::fastDetectAndCompute(unsigned char* image, AkazeData& result, int3 whp0, const bool desc)
    {
        state->run(image, AKZ_U8, result, whp0, desc, true);
    }

    // ---- akazed.h: float stage functions, synchronous like the reference's wrappers --------------------------
    void setLowPassKernel(const float*, const int) {}      // taps are kernel arguments in this build

    void hLowPass(float* src, float* dst, int width, int height, int pitch, float var, int ksz)
    {
        akz_ctx* c = stage_ctx();
        if (ksz > 11) { std::cerr << "Kernels larger than 11 not implemented" << std::endl; return; }   // akazed.cu:2377-2380
        AKZ_DO(akz_lowpass(c, src, dst, width, height, pitch, (long long)pitch * height, 1, var, ksz));
        AKZ_DO(akz_sync(c));
    }

    void hDownWithSmooth(float* src, float* dst, float* smooth, int3 swhp, int3 dwhp)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_down_with_smooth(c, src, dst, smooth, swhp.x, swhp.y, swhp.z, (long long)swhp.y * swhp.z,
                                    dwhp.x, dwhp.y, dwhp.z, (long long)dwhp.y * dwhp.z, 1));
        AKZ_DO(akz_sync(c));
    }

    // grad is accepted for signature compatibility; the gradient plane is recomputed on the fly and not stored
    void hScharrContrast(float* src, float* /*grad*/, float& kcontrast, float per, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        float* dk = (float*)g_k_scratch.get(sizeof(float));
        AKZ_DO(akz_scharr_contrast(c, src, dk, per, width, height, pitch, (long long)pitch * height, 1));
        CHECK(cudaMemcpyAsync(&kcontrast, dk, sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)akz_stream(c)));
        AKZ_DO(akz_sync(c));
    }

    void hFlow(float* src, float* flow, DiffusivityType type, float kcontrast, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        float* dk = (float*)g_k_scratch.get(sizeof(float));
        CHECK(cudaMemcpyAsync(dk, &kcontrast, sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)akz_stream(c)));
        AKZ_DO(akz_flow(c, src, flow, (int)type, dk, 1.0f, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    void hNldStep(float* img, float* flow, float* temp, float step_size, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_nld_step(c, img, flow, temp, step_size, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    // like the reference, the determinant overwrites src (akazed.cu:2550)
    void hHessianDeterminant(float* src, float* dx, float* dy, int step, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_hessian(c, src, dx, dy, src, step, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    void hMatch(AkazeData& result1, AkazeData& result2) { match_points(result1, result2); }

    void hConv2d(float*, float*, int, int, int) { not_routed("akaze::hConv2d"); }
    void hSepConv2d(float*, float*, int, int, int) { not_routed("akaze::hSepConv2d"); }
    void hCalcExtremaMap(float*, float*, float*, int*, float*, int, int, float, int, int, int, int) { not_routed("akaze::hCalcExtremaMap"); }
    void hNms(AkazePoint*, float*, float*, int*, int, int, int, int) { not_routed("akaze::hNms"); }
    void hNmsR(AkazePoint*, float*, float*, int*, int, int, int, int, int) { not_routed("akaze::hNmsR"); }
    void hRefine(AkazeData&, float*, int, int) { not_routed("akaze::hRefine"); }
    void hCalcOrient(AkazeData&, float*, int, int) { not_routed("akaze::hCalcOrient"); }
    void hDescribe(AkazeData&, float*, int, int, int) { not_routed("akaze::hDescribe"); }
}

namespace fastakaze
{
    // Integer stage functions (akazed.h:88-110), synchronous like the reference's wrappers.  The overloads without a
    // `temp` argument run the same separable kernels on a scratch plane of the shim (the results are identical: integer sums).
    static void blur(const void* src, int src_u8, int* dst, int* temp, int width, int height, int pitch, float var, int ksz)
    {
        akz_ctx* c = stage_ctx();
        if (ksz > 11) { std::cerr << "Kernels larger than 11 not implemented" << std::endl; return; }   // akazed.cu:4007
        if (!temp) temp = (int*)g_tmp_scratch.get(sizeof(int) * (size_t)pitch * height);
        AKZ_DO(akz_fast_lowpass(c, src, src_u8, dst, temp, width, height, pitch, (long long)pitch * height, 1, var, ksz));
        AKZ_DO(akz_sync(c));
    }
    void hConv2dR2(unsigned char* src, int* dst, int width, int height, int pitch, float var) { blur(src, 1, dst, nullptr, width, height, pitch, var, 5); }
    void hConv2dR2(int* src, int* dst, int width, int height, int pitch, float var) { blur(src, 0, dst, nullptr, width, height, pitch, var, 5); }
    void hConv2dR2(unsigned char* src, int* dst, int* temp, int width, int height, int pitch, float var) { blur(src, 1, dst, temp, width, height, pitch, var, 5); }
    void hConv2dR2(int* src, int* dst, int* temp, int width, int height, int pitch, float var) { blur(src, 0, dst, temp, width, height, pitch, var, 5); }
    void hLowPass(unsigned char* src, int* dst, int width, int height, int pitch, float var, int ksz) { blur(src, 1, dst, nullptr, width, height, pitch, var, ksz); }
    void hLowPass(unsigned char* src, int* dst, int* temp, int width, int height, int pitch, float var, int ksz) { blur(src, 1, dst, temp, width, height, pitch, var, ksz); }

    void hDownWithSmooth(int* src, int* dst, int* smooth, int3 swhp, int3 dwhp)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_fast_down_with_smooth(c, src, dst, smooth, swhp.x, swhp.y, swhp.z, (long long)swhp.y * swhp.z,
                                         dwhp.x, dwhp.y, dwhp.z, (long long)dwhp.y * dwhp.z, 1));
        AKZ_DO(akz_sync(c));
    }

    // grad receives the gradient-magnitude plane, as in the reference (akazed.cu:4098)
    void hScharrContrast(int* src, int* grad, int& kcontrast, float per, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        int* dk = (int*)g_k_scratch.get(sizeof(int));
        AKZ_DO(akz_fast_scharr_contrast(c, src, grad, dk, per, width, height, pitch, (long long)pitch * height, 1));
        CHECK(cudaMemcpyAsync(&kcontrast, dk, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)akz_stream(c)));
        AKZ_DO(akz_sync(c));
    }

    // like the reference, the determinant overwrites src (akazed.cu:4175-4206)
    void hHessianDeterminant(int* src, int* dx, int* dy, int step, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_fast_hessian(c, src, dx, dy, src, step, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    void hFlow(int* src, int* flow, akaze::DiffusivityType type, int kcontrast, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        int* dk = (int*)g_k_scratch.get(sizeof(int));
        CHECK(cudaMemcpyAsync(dk, &kcontrast, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)akz_stream(c)));
        AKZ_DO(akz_fast_flow(c, src, flow, (int)type, dk, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    void hNldStep(int* img, int* flow, int* temp, float step_size, int width, int height, int pitch)
    {
        akz_ctx* c = stage_ctx();
        AKZ_DO(akz_fast_nld_step(c, img, flow, temp, step_size, width, height, pitch, (long long)pitch * height, 1));
        AKZ_DO(akz_sync(c));
    }

    void hCalcExtremaMap(int*, int*, float*, int*, float*, int, int, int, int, int, int, int) { not_routed("fastakaze::hCalcExtremaMap"); }
    void hNmsR(akaze::AkazePoint*, int*, float*, int*, int, int, int, int, int) { not_routed("fastakaze::hNmsR"); }
    void hRefine(akaze::AkazeData&, void*, int, int) { not_routed("fastakaze::hRefine"); }
    void hCalcOrient(akaze::AkazeData&, void*, int, int) { not_routed("fastakaze::hCalcOrient"); }
    void hDescribe(akaze::AkazeData&, void*, int, int, int) { not_routed("fastakaze::hDescribe"); }
}
