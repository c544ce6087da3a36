// Brute-force Hamming matcher over [n][64]-byte M-LDB descriptors (486 valid bits, zero padding).
// Reference: gHammingMatch (akazed.cu:2144-2241, the live 1-NN with the 16-stride uniqueness gate
// and the <96 gate) and gMatch (akazed.cu:2028-2122, the top-2 variant).
//
// Layout: each thread keeps TWO query descriptors in 32 registers; the block stages tiles of train
// descriptors in shared memory and every thread walks the tile with broadcast LDS.128 reads (one read serves
// both queries), so the inner loop touches no global memory.  The 512-bit distance is a carry-save
// (Harley-Seal) tree: 16 XOR + 30 LOP3 + 5 POPC per pair instead of 16 XOR + 16 POPC + 15 IADD.  The train
// range is split across blockIdx.y so the grid covers all 148 SMs even for a few thousand queries; partial
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

constexpr int QPT = 2;                   // queries per thread
constexpr int NTH = 128;                 // threads per block
constexpr int QPB = QPT * NTH;           // queries per block
constexpr int TILE = 128;                // train descriptors per shared-memory tile (8 KB)

// 3-input logic, written as PTX so that ptxas keeps the carry-save structure (left to itself the compiler re-associates
// the tree into ~95 LOP3 per pair instead of 46: ncu r01e, ALU pipe 89.5 % busy)
__device__ __forceinline__ unsigned xor3(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#define AKZ_CSA(h, l, a, b, c) { unsigned a_ = (a), b_ = (b), c_ = (c); (h) = maj3(a_, b_, c_); (l) = xor3(a_, b_, c_); }
#define AKZ_HA(h, l, a, b)     { unsigned a_ = (a), b_ = (b); (h) = a_ & b_; (l) = a_ ^ b_; }

// Hamming distance of two 512-bit strings: one carry-save level, then POPC.
//   plain      : 16 XOR + 16 POPC + adds   POPC issues at 4 lanes/clk/SMSP: 128 SMSP cycles per warp-pair, POPC bound
//   full tree  : 16 XOR + 30 LOP3 + 5 POPC  ALU pipe (16 lanes/clk/SMSP) bound: ~116 cycles (ncu r01g: ALU 84.5 % busy)
//   this form  : 16 XOR + 10 LOP3 + 11 POPC the two pipes are balanced: 2 x 36 ALU instructions vs 8 x 11 POPC cycles
// five 3:2 compressors turn 15 words into 5 sums (weight 1) and 5 carries (weight 2); word 15 is counted directly.
__device__ __forceinline__ int hamming512(const unsigned (&q)[16], const uint4& t0, const uint4& t1, const uint4& t2, const uint4& t3)
{
    const unsigned x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const unsigned x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    const unsigned x8 = q[8] ^ t2.x, x9 = q[9] ^ t2.y, x10 = q[10] ^ t2.z, x11 = q[11] ^ t2.w;
    const unsigned x12 = q[12] ^ t3.x, x13 = q[13] ^ t3.y, x14 = q[14] ^ t3.z, x15 = q[15] ^ t3.w;
    unsigned s0, s1, s2, s3, s4, c0, c1, c2, c3, c4;
    AKZ_CSA(c0, s0, x0, x1, x2)
    AKZ_CSA(c1, s1, x3, x4, x5)
    AKZ_CSA(c2, s2, x6, x7, x8)
    AKZ_CSA(c3, s3, x9, x10, x11)
    AKZ_CSA(c4, s4, x12, x13, x14)
    const int ones = __popc(s0) + __popc(s1) + __popc(s2) + __popc(s3) + __popc(s4) + __popc(x15);
    const int twos = __popc(c0) + __popc(c1) + __popc(c2) + __popc(c3) + __popc(c4);
    return ones + 2 * twos;
}

// Running best of a query.  A candidate is ordered by key = distance << 22 | (index relative to the block's train
// range): unsigned order of the key IS the lexicographic (distance, index) order the matcher contract asks for.
//   KNN2  : two smallest keys (min / max network, 3 instructions per pair)
//   COMPAT: smallest key + the 16-bit mask of (index mod 16) classes that attain the smallest DISTANCE
struct Best { unsigned k1, k2; };
constexpr unsigned KEY_NONE = 0xFFFFFFFFu;
constexpr int KEY_IDX_BITS = 22;

template <int MODE>
__device__ __forceinline__ void consider(Best& b, int d, int jrel, unsigned classbit)
{
    const unsigned key = ((unsigned)d << KEY_IDX_BITS) | (unsigned)jrel;
    if (MODE == AKZ_MATCH_KNN2) {
        const unsigned hi = max(b.k1, key);
        b.k1 = min(b.k1, key);
        b.k2 = min(b.k2, hi);
    } else {
        const unsigned dcur = b.k1 >> KEY_IDX_BITS;                    // KEY_NONE >> 22 = 1023 > any distance
        b.k2 = (unsigned)d < dcur ? classbit : ((unsigned)d == dcur ? (b.k2 | classbit) : b.k2);
        b.k1 = min(b.k1, key);
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTH) k_match(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int tbase,
                                               int per_split, akz_match_t* __restrict__ parts)
{
    __shared__ uint4 tile[TILE * 4];
    unsigned qa[QPT][16];
    int qi[QPT];
    Best b[QPT];
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        qi[k] = blockIdx.x * QPB + k * NTH + threadIdx.x;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 w = qi[k] < nq ? __ldg(q + 4 * (long long)qi[k] + v) : make_uint4(0, 0, 0, 0);
            qa[k][4 * v] = w.x; qa[k][4 * v + 1] = w.y; qa[k][4 * v + 2] = w.z; qa[k][4 * v + 3] = w.w;
        }
        b[k].k1 = KEY_NONE; b[k].k2 = (MODE == AKZ_MATCH_KNN2) ? KEY_NONE : 0u;
    }
    const int t0 = blockIdx.y * per_split, t1 = min(nt, t0 + per_split);
    for (int base = t0; base < t1; base += TILE) {
        int cnt = min(TILE, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 4; i += NTH) tile[i] = __ldg(t + 4 * (long long)base + i);
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < cnt; j++) {
            const uint4 b0 = tile[4 * j], b1 = tile[4 * j + 1], b2 = tile[4 * j + 2], b3 = tile[4 * j + 3];
            const int jrel = base - t0 + j;
            const unsigned classbit = 1u << ((tbase + base + j) & 15);
#pragma unroll
            for (int k = 0; k < QPT; k++) consider<MODE>(b[k], hamming512(qa[k], b0, b1, b2, b3), jrel, classbit);
        }
    }
    const unsigned imask = (1u << KEY_IDX_BITS) - 1;
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        if (qi[k] < nq) {
            akz_match_t m;
            const bool has1 = b[k].k1 != KEY_NONE;
            m.idx1 = has1 ? tbase + t0 + (int)(b[k].k1 & imask) : -1;
            m.dist1 = has1 ? (int)(b[k].k1 >> KEY_IDX_BITS) : -1;
            if (MODE == AKZ_MATCH_KNN2) {
                const bool has2 = b[k].k2 != KEY_NONE;
                m.idx2 = has2 ? tbase + t0 + (int)(b[k].k2 & imask) : -1;
                m.dist2 = has2 ? (int)(b[k].k2 >> KEY_IDX_BITS) : -1;
            } else {
                m.idx2 = (int)b[k].k2; m.dist2 = 0;
            }
            parts[(long long)blockIdx.y * nq + qi[k]] = m;
        }
    }
}

__device__ __forceinline__ bool lex_less(int d, int i, int d2, int i2) { return d < d2 || (d == d2 && i < i2); }

__global__ void k_match_merge(const akz_match_t* __restrict__ parts, int nparts, int nq, int mode, int finalize, akz_match_t* __restrict__ out)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    akz_match_t r;
    const bool top2 = mode != AKZ_MATCH_COMPAT;                // KNN2 and UNIQUE2 share the partial form
    r.idx1 = -1; r.dist1 = -1; r.idx2 = top2 ? -1 : 0; r.dist2 = top2 ? -1 : 0;
    for (int p = 0; p < nparts; p++) {
        akz_match_t m = parts[(long long)p * nq + qi];
        if (m.idx1 < 0) continue;
        if (top2) {
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
    if (finalize && mode == AKZ_MATCH_UNIQUE2) {
        // gMatch (akazed.cu:2103): the best must be strictly better than the second best and below MAX_DIST
        bool ok = r.idx1 >= 0 && r.dist1 < AKZ_MAX_DIST && (r.idx2 < 0 || r.dist1 < r.dist2);
        if (!ok) { r.idx1 = -1; r.dist1 = -1; }
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
    if (per >= (1 << KEY_IDX_BITS)) return akz_set_error(AKZ_E_UNSUPPORTED, "train range per block exceeds 2^22 descriptors: shard the train set");
    dim3 g((nq + QPB - 1) / QPB, nsplit);
    if (mode != AKZ_MATCH_COMPAT)
        k_match<AKZ_MATCH_KNN2><<<g, NTH, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    else
        k_match<AKZ_MATCH_COMPAT><<<g, NTH, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    return 1;
}

int match_merge(cudaStream_t st, const akz_match_t* parts, int nparts, int nq, int mode, int finalize, akz_match_t* out)
{
    if (nq <= 0) return 0;
    k_match_merge<<<(nq + 127) / 128, 128, 0, st>>>(parts, nparts, nq, mode, finalize, out);
    return 1;
}

}  // namespace akzk
