// Brute-force Hamming matcher over [n][64]-byte M-LDB descriptors (486 valid bits, zero padding).
// Reference: gHammingMatch (akazed.cu:2144-2241, the live 1-NN with the 16-stride uniqueness gate
// and the <96 gate) and gMatch (akazed.cu:2028-2122, the top-2 variant).
//
// Layout: each thread keeps ONE query descriptor in 16 registers; the block stages tiles of train
// descriptors in shared memory and every thread walks the tile with broadcast LDS.128 reads, so the
// inner loop is 16 LOP3 + 16 POPC + adds per pair and touches no global memory.  The train range is
// split across blockIdx.y so the grid covers all 148 SMs even for a few thousand queries; partial
// results are merged by k_match_merge with an associative rule, which is also what the train-sharded
// multi-GPU path applies to the per-shard results after the NCCL gather.
//
// Partial-result encoding (akz_match_t):
//   KNN2   (idx1,dist1) best, (idx2,dist2) second best; lexicographic (distance, index) order
//   COMPAT dist1 = minimum distance, idx1 = lowest index attaining it, idx2 = bit mask of the
//          (index mod 16) classes attaining it.  The reference accepts iff exactly one class attains
//          the minimum and it is < 96; that test is applied once, after the last merge.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int QPB = 128;        // queries per block (one per thread)
constexpr int TILE = 128;       // train descriptors per shared-memory tile (8 KB)

struct Best { int d1, i1, d2, i2; };

template <int MODE>
__device__ __forceinline__ void consider(Best& b, int d, int j)
{
    if (MODE == AKZ_MATCH_KNN2) {
        if (d < b.d1) { b.d2 = b.d1; b.i2 = b.i1; b.d1 = d; b.i1 = j; }
        else if (d < b.d2) { b.d2 = d; b.i2 = j; }
    } else {
        if (d < b.d1) { b.d1 = d; b.i1 = j; b.i2 = 1 << (j & 15); }
        else if (d == b.d1) b.i2 |= 1 << (j & 15);
    }
}

template <int MODE>
__global__ void __launch_bounds__(QPB) k_match(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int tbase,
                                               int per_split, akz_match_t* __restrict__ parts)
{
    __shared__ uint4 tile[TILE * 4];
    int qi = blockIdx.x * QPB + threadIdx.x;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    if (qi < nq) { a0 = __ldg(q + 4 * qi); a1 = __ldg(q + 4 * qi + 1); a2 = __ldg(q + 4 * qi + 2); a3 = __ldg(q + 4 * qi + 3); }
    int t0 = blockIdx.y * per_split, t1 = min(nt, t0 + per_split);
    Best b;
    b.d1 = 1 << 20; b.i1 = -1; b.d2 = (MODE == AKZ_MATCH_KNN2) ? (1 << 20) : 0; b.i2 = (MODE == AKZ_MATCH_KNN2) ? -1 : 0;
    for (int base = t0; base < t1; base += TILE) {
        int cnt = min(TILE, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 4; i += QPB) tile[i] = __ldg(t + 4 * (long long)base + i);
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < cnt; j++) {
            uint4 b0 = tile[4 * j], b1 = tile[4 * j + 1], b2 = tile[4 * j + 2], b3 = tile[4 * j + 3];
            int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w)
                  + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w)
                  + __popc(a2.x ^ b2.x) + __popc(a2.y ^ b2.y) + __popc(a2.z ^ b2.z) + __popc(a2.w ^ b2.w)
                  + __popc(a3.x ^ b3.x) + __popc(a3.y ^ b3.y) + __popc(a3.z ^ b3.z) + __popc(a3.w ^ b3.w);
            consider<MODE>(b, d, tbase + base + j);
        }
    }
    if (qi < nq) {
        akz_match_t m;
        m.idx1 = b.i1; m.dist1 = b.i1 < 0 ? -1 : b.d1; m.idx2 = b.i2;
        m.dist2 = (MODE == AKZ_MATCH_KNN2) ? (b.i2 < 0 ? -1 : b.d2) : 0;
        parts[(long long)blockIdx.y * nq + qi] = m;
    }
}

__device__ __forceinline__ bool lex_less(int d, int i, int d2, int i2) { return d < d2 || (d == d2 && i < i2); }

__global__ void k_match_merge(const akz_match_t* __restrict__ parts, int nparts, int nq, int mode, int finalize, akz_match_t* __restrict__ out)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    akz_match_t r;
    r.idx1 = -1; r.dist1 = -1; r.idx2 = (mode == AKZ_MATCH_KNN2) ? -1 : 0; r.dist2 = (mode == AKZ_MATCH_KNN2) ? -1 : 0;
    for (int p = 0; p < nparts; p++) {
        akz_match_t m = parts[(long long)p * nq + qi];
        if (m.idx1 < 0) continue;
        if (mode == AKZ_MATCH_KNN2) {
            // merge two sorted pairs, lowest (distance, index) first
            int cd[4] = { r.dist1, r.dist2, m.dist1, m.dist2 };
            int ci[4] = { r.idx1, r.idx2, m.idx1, m.idx2 };
            int bd1 = 1 << 20, bi1 = -1, bd2 = 1 << 20, bi2 = -1;
            for (int k = 0; k < 4; k++) {
                if (ci[k] < 0) continue;
                if (bi1 < 0 || lex_less(cd[k], ci[k], bd1, bi1)) { bd2 = bd1; bi2 = bi1; bd1 = cd[k]; bi1 = ci[k]; }
                else if (bi2 < 0 || lex_less(cd[k], ci[k], bd2, bi2)) { bd2 = cd[k]; bi2 = ci[k]; }
            }
            r.idx1 = bi1; r.dist1 = bi1 < 0 ? -1 : bd1; r.idx2 = bi2; r.dist2 = bi2 < 0 ? -1 : bd2;
        } else {
            if (r.idx1 < 0 || m.dist1 < r.dist1) r = m;
            else if (m.dist1 == r.dist1) { r.idx1 = min(r.idx1, m.idx1); r.idx2 |= m.idx2; }
        }
    }
    if (finalize && mode == AKZ_MATCH_COMPAT) {
        // akazed.cu:2222: the minimum must be strictly unique across the 16 strides and below MAX_DIST
        bool ok = r.idx1 >= 0 && __popc((unsigned)r.idx2) == 1 && r.dist1 < AKZ_MAX_DIST;
        if (!ok) { r.idx1 = -1; r.dist1 = -1; }
    }
    out[qi] = r;
}

}  // namespace

namespace akzk {

int match_partial(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode,
                  int nsplit, akz_match_t* parts)
{
    if (nq <= 0) return 0;
    int per = (nt + nsplit - 1) / nsplit;
    per = ((per + TILE - 1) / TILE) * TILE;
    if (per <= 0) per = TILE;
    dim3 g((nq + QPB - 1) / QPB, nsplit);
    if (mode == AKZ_MATCH_KNN2)
        k_match<AKZ_MATCH_KNN2><<<g, QPB, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    else
        k_match<AKZ_MATCH_COMPAT><<<g, QPB, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    return 1;
}

int match_merge(cudaStream_t st, const akz_match_t* parts, int nparts, int nq, int mode, int finalize, akz_match_t* out)
{
    if (nq <= 0) return 0;
    k_match_merge<<<(nq + 127) / 128, 128, 0, st>>>(parts, nparts, nq, mode, finalize, out);
    return 1;
}

}  // namespace akzk
