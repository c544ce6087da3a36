// Drop-in header: result records of the AKAZE path, layout-compatible with the reference
// (Accustomer/CUDA-AKAZE akaze_structures.h:15-59).  The B200 pipeline itself works on SoA buffers
// (akaze_b200.h: akz_keypoint + [n][64] descriptors); these AoS records exist only at this boundary.
#pragma once

namespace akaze
{

// Only the M-LDB descriptor (type 5 of the reference's list: 0 SURF_UPRIGHT, 1 SURF, 2 MSURF_UPRIGHT,
// 3 MSURF, 4 MLDB_UPRIGHT, 5 MLDB) is implemented by the reference and by this build.
#define FEATURE_TYPE 5
#define FLEN 61                       // 486 bits

    // 104 bytes; offsets x0 y4 octave8 response12 size16 angle20 features24 match88 distance92 match_x96 match_y100
    struct AkazePoint
    {
        float x, y;                   // full-resolution position
        int   octave;                 // layer index: octave * sublevels + sublevel
        float response;               // never written (as in the reference)
        float size;                   // octave-relative derivative scale
        float angle;                  // radians, [0, 2pi)
        unsigned char features[FLEN];
        int   match;                  // index into the train set, or -1
        int   distance;               // Hamming distance, or -1
        float match_x, match_y;       // position of the matched train point, or -1
    };
    static_assert(sizeof(AkazePoint) == 104, "AkazePoint must keep the reference layout");

    struct AkazeData
    {
        int num_pts;                  // valid points
        int max_pts;                  // capacity of h_data / d_data
        AkazePoint* h_data;           // host copy (may be NULL)
        AkazePoint* d_data;           // device copy
    };

    enum DiffusivityType { PM_G1 = 0, PM_G2 = 1, WEICKERT = 2, CHARBONNIER = 3 };

}
