// Compile-only check (tests/test_cpu.py) and, when built with nvcc and linked against
// libakaze_b200.so, a runnable demo: it drives the drop-in surface the way the reference's
// main.cpp:128-233 does, minus OpenCV (images come from a binary PGM or a synthetic pattern).
#include "akaze.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

static bool read_pgm(const char* path, std::vector<unsigned char>& px, int& w, int& h)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int maxv = 0;
    if (fscanf(f, "P5 %d %d %d", &w, &h, &maxv) != 3) { fclose(f); return false; }
    fgetc(f);
    px.resize((size_t)w * h);
    bool ok = fread(px.data(), 1, px.size(), f) == px.size();
    fclose(f);
    return ok;
}

static void synthetic(std::vector<unsigned char>& px, int w, int h, int shift)
{
    px.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int xs = x + shift;
            int v = ((xs / 37 + y / 29) & 1) * 120 + ((xs * 7 + y * 13) % 61) + (((xs / 11) ^ (y / 17)) & 3) * 15;
            px[(size_t)y * w + x] = (unsigned char)(v > 255 ? 255 : v);
        }
}

int main(int argc, char** argv)
{
    int devNum = argc > 1 ? std::atoi(argv[1]) : 0;
    std::vector<unsigned char> l8, r8;
    int w = 640, h = 480, w2 = 640, h2 = 480;
    if (!(argc > 3 && read_pgm(argv[2], l8, w, h) && read_pgm(argv[3], r8, w2, h2))) {
        synthetic(l8, w, h, 0);
        synthetic(r8, w2, h2, 9);
    }
    std::vector<float> limg(l8.size()), rimg(r8.size());
    for (size_t i = 0; i < l8.size(); i++) limg[i] = (float)l8[i] * (float)(1.0 / 255.0);
    for (size_t i = 0; i < r8.size(); i++) rimg[i] = (float)r8[i] * (float)(1.0 / 255.0);

    int max_npts = 10000, noctaves = 4, max_scale = 4;
    float per = 0.7f, kcontrast = 0.03f, soffset = 1.6f, derivative_factor = 1.5f, dthreshold = 0.001f;
    bool reordering = true;
    int diffusivity = 1, descriptor_pattern_size = 10;

    if (!initDevice(devNum)) return 1;
    GpuTimer timer(0);
    int3 whp1, whp2;
    whp1.x = w; whp1.y = h; whp1.z = iAlignUp(whp1.x, 128);
    whp2.x = w2; whp2.y = h2; whp2.z = iAlignUp(whp2.x, 128);
    float *img1 = NULL, *img2 = NULL;
    CHECK(cudaMalloc((void**)&img1, sizeof(float) * whp1.y * whp1.z));
    CHECK(cudaMalloc((void**)&img2, sizeof(float) * whp2.y * whp2.z));
    CHECK(cudaMemcpy2D(img1, sizeof(float) * whp1.z, limg.data(), sizeof(float) * whp1.x, sizeof(float) * whp1.x, whp1.y, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy2D(img2, sizeof(float) * whp2.z, rimg.data(), sizeof(float) * whp2.x, sizeof(float) * whp2.x, whp2.y, cudaMemcpyHostToDevice));

    akaze::AkazeData data1, data2;
    akaze::initAkazeData(data1, max_npts, true, true);
    akaze::initAkazeData(data2, max_npts, true, true);

    std::unique_ptr<akaze::Akazer> detector(new akaze::Akazer);
    detector->init(whp1, noctaves, max_scale, per, kcontrast, soffset, reordering, derivative_factor, dthreshold, diffusivity, descriptor_pattern_size);

    int nrepeats = 10;
    float t1 = timer.read();
    for (int i = 0; i < nrepeats; i++) {
        detector->detectAndCompute(img1, data1, whp1, true);
        detector->detectAndCompute(img2, data2, whp2, true);
    }
    float t2 = timer.read();
    akaze::cuMatch(data1, data2);
    float t3 = timer.read();

    int matched = 0;
    for (int i = 0; i < data1.num_pts; i++) {
        const akaze::AkazePoint& p = data1.h_data[i];
        if (p.match >= 0) { matched++; (void)p.x; (void)p.y; (void)p.size; }
    }
    printf("features1 %d features2 %d matched %d detect_ms_per_pair %.3f match_ms %.3f\n",
           data1.num_pts, data2.num_pts, matched, (t2 - t1) / nrepeats, t3 - t2);

    // the integer entry point must exist and run (main.cpp:311-312)
    std::vector<unsigned char> padded((size_t)whp1.y * whp1.z, 0);
    for (int y = 0; y < h; y++) memcpy(&padded[(size_t)y * whp1.z], &l8[(size_t)y * w], w);
    unsigned char* img8 = NULL;
    CHECK(cudaMalloc((void**)&img8, padded.size()));
    CHECK(cudaMemcpy(img8, padded.data(), padded.size(), cudaMemcpyHostToDevice));
    detector->fastDetectAndCompute(img8, data2, whp1, true);
    printf("fast_features %d\n", data2.num_pts);

    akaze::freeAkazeData(data1);
    akaze::freeAkazeData(data2);
    CHECK(cudaFree(img1));
    CHECK(cudaFree(img2));
    CHECK(cudaFree(img8));
    return 0;
}
