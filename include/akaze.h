// Drop-in header: the public C++ entry points of the reference (akaze.h:10-30), implemented on top of
// the C ABI in akaze_b200.h by cuda-akaze_b200/csrc/akaze_compat.cu.  main.cpp of the reference
// compiles against this header unchanged.
#pragma once
#include "akaze_structures.h"
#include "cuda_utils.h"

// the product library is built with -fvisibility=hidden: the drop-in surface is exported explicitly
#pragma GCC visibility push(default)

namespace akaze
{
    // Allocate / release the host and device arrays of an AkazeData (reference akaze.cpp:26-52).
    void initAkazeData(AkazeData& data, const int max_pts, const bool host, const bool dev);
    void freeAkazeData(AkazeData& data);

    // Reference-compatible matching (akaze.cpp:55-64): writes match / distance / match_x / match_y of
    // every point of result1 (device, and host when h_data is set).
    void cuMatch(AkazeData& result1, AkazeData& result2);

    class Akazer
    {
    public:
        Akazer();
        ~Akazer();

        // Same eleven options as the reference (akaze.h:25-26); whp0 = {width, height, pitch in elements}.
        void init(int3 whp0, int _noctaves, int _max_scale, float _per, float _kcontrast, float _soffset, bool _reordering,
            float _derivative_factor, float _dthreshold, int _diffusivity, int _descriptor_pattern_size);

        // image: device pointer, row pitch whp0.z elements; float in [0,1] or raw 8-bit grey.
        void detectAndCompute(float* image, AkazeData& result, int3 whp0, const bool desc = true);
        void fastDetectAndCompute(unsigned char* image, AkazeData& result, int3 whp0, const bool desc = true);

    private:
        struct State;
        State* state;
        Akazer(const Akazer&);
        Akazer& operator=(const Akazer&);
    };
}

#pragma GCC visibility pop
