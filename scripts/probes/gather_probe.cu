// Probe: cost of the M-LDB gather pattern (441 samples on a rotated lattice around raster-ordered keypoints) under different
// plane layouts.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Kp { float x, y, co, si, scale; int frame; int level; int pad; };
constexpr int W = 1920, H = 1080, PITCH = 1920;
__constant__ int NF;

// MODE 0: three planes, three LDG.32.  1: Lt plane + interleaved (Lx, Ly) float2.  2: float4 records.  3: three planes through tex2D
template <int MODE, int SLOTS>
__global__ void __launch_bounds__(64) k_gather(const Kp* __restrict__ kp, int n, const float* __restrict__ p0, const float* __restrict__ p1,
                                              const float* __restrict__ p2, const cudaTextureObject_t* __restrict__ tex, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int g = w; g < n; g += nw) {
        const Kp k = kp[g];
        const long long fo = ((long long)k.level * NF + k.frame) * PITCH * H;
        float v[SLOTS][3];
#pragma unroll
        for (int j = 0; j < SLOTS; j++) {
            const int i = lane + 32 * j;
            const int y = i / 21, x = i - 21 * y;
            const float l = (float)(x - 10), kk = (float)(y - 10);
            int xp = (int)(k.scale * (k.co * kk - k.si * l) + k.x + 0.5f);
            int yp = (int)(k.scale * (k.si * kk + k.co * l) + k.y + 0.5f);
            xp = min(max(xp, 0), W - 1); yp = min(max(yp, 0), H - 1);
            const long long pos = fo + (long long)yp * PITCH + xp;
            if (MODE == 0) { v[j][0] = __ldg(p0 + pos); v[j][1] = __ldg(p1 + pos); v[j][2] = __ldg(p2 + pos); }
            else if (MODE == 1) { v[j][0] = __ldg(p0 + pos); const float2 q = __ldg(reinterpret_cast<const float2*>(p1) + pos); v[j][1] = q.x; v[j][2] = q.y; }
            else if (MODE == 2) { const float4 q = __ldg(reinterpret_cast<const float4*>(p0) + pos); v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; }
            else {
                const float fx = xp + 0.5f, fy = yp + 0.5f + (float)(k.frame * H);   // (texture variant ignores levels)
                v[j][0] = tex2D<float>(tex[0], fx, fy); v[j][1] = tex2D<float>(tex[1], fx, fy); v[j][2] = tex2D<float>(tex[2], fx, fy);
            }
        }
#pragma unroll
        for (int j = 0; j < SLOTS; j++) acc += v[j][0] + v[j][1] * k.co + v[j][2] * k.si;
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char** argv)
{
    const int nf = argc > 1 ? atoi(argv[1]) : 16, per = argc > 2 ? atoi(argv[2]) : 19000;
    const int n = nf * per;
    const int nlev = argc > 3 ? atoi(argv[3]) : 1, sortlev = argc > 4 ? atoi(argv[4]) : 0;
    CK(cudaMemcpyToSymbol(NF, &nf, sizeof(int)));
    std::vector<Kp> h(n);
    srand(1);
    for (int f = 0; f < nf; f++) {
        std::vector<std::pair<int, int>> pts(per);
        for (auto& p : pts) { p.second = 60 + rand() % (W - 120); p.first = 60 + rand() % (H - 120); }
        std::sort(pts.begin(), pts.end());
        std::vector<int> lev(per); for (auto& l : lev) l = rand() % nlev;
        if (sortlev) { std::vector<int> idx(per); for (int i = 0; i < per; i++) idx[i] = i; std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return lev[a] < lev[b]; }); std::vector<std::pair<int,int>> p2(per); std::vector<int> l2(per); for (int i = 0; i < per; i++) { p2[i] = pts[idx[i]]; l2[i] = lev[idx[i]]; } pts = p2; lev = l2; }
        for (int i = 0; i < per; i++) {
            Kp& k = h[(size_t)f * per + i];
            const float a = 6.2831853f * (rand() / (float)RAND_MAX);
            k.x = pts[i].second; k.y = pts[i].first; k.co = cosf(a); k.si = sinf(a); k.scale = 2 + rand() % 3; k.frame = f; k.level = lev[i];
        }
    }
    Kp* d; CK(cudaMalloc(&d, n * sizeof(Kp))); CK(cudaMemcpy(d, h.data(), n * sizeof(Kp), cudaMemcpyHostToDevice));
    const size_t plane = (size_t)PITCH * H * nf * nlev;
    float *p0, *p1, *p2, *out;
    CK(cudaMalloc(&p0, plane * (nlev > 1 ? 4 : 16))); CK(cudaMalloc(&p1, plane * (nlev > 1 ? 8 : 8))); CK(cudaMalloc(&p2, plane * 4)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(p0, 0, plane * (nlev > 1 ? 4 : 16))); CK(cudaMemset(p1, 0, plane * 8)); CK(cudaMemset(p2, 0, plane * 4));
    cudaTextureObject_t ht[3], *dt;
    float* tp[3] = { p0, p1, p2 };
    for (int i = 0; i < 3; i++) {
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = tp[i]; rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        rd.res.pitch2D.width = W; rd.res.pitch2D.height = (size_t)H * nf; rd.res.pitch2D.pitchInBytes = PITCH * 4;
        cudaTextureDesc td = {}; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        CK(cudaCreateTextureObject(&ht[i], &rd, &td, nullptr));
    }
    CK(cudaMalloc(&dt, sizeof(ht))); CK(cudaMemcpy(dt, ht, sizeof(ht), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float* flush; CK(cudaMalloc(&flush, 512 << 20));
    auto run = [&](const char* name, auto launch) {
        float best = 1e9;
        for (int it = 0; it < 4; it++) {
            cudaMemsetAsync(flush, it, 512 << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("%-44s %8.3f ms  %6.2f ns/kp\n", name, best, best * 1e6 / n);
    };
    for (int bps = 4; bps <= 16; bps *= 2) {
        const int grid = 148 * bps;
        printf("-- %d CTAs of 64 threads per SM\n", bps);
        run("3 planes, 14 slots in flight", [&] { k_gather<0, 14><<<grid, 64>>>(d, n, p0, p1, p2, dt, out); });
        run("Lt + float2(Lx,Ly), 14 slots", [&] { k_gather<1, 14><<<grid, 64>>>(d, n, p0, p1, p2, dt, out); });
        if (nlev == 1) run("float4 records, 14 slots", [&] { k_gather<2, 14><<<grid, 64>>>(d, n, p0, p1, p2, dt, out); });
        if (nlev == 1) run("3 planes through tex2D, 14 slots", [&] { k_gather<3, 14><<<grid, 64>>>(d, n, p0, p1, p2, dt, out); });
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
