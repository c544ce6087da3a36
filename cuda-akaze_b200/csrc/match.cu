// Brute-force Hamming matcher over [n][64]-byte M-LDB descriptors (486 valid bits, zero padding).
// Reference: gHammingMatch (akazed.cu:2144-2241, the live 1-NN with the 16-stride uniqueness gate
// and the <96 gate) and gMatch (akazed.cu:2028-2122, the top-2 variant).
//
// Two kernels with identical (integer) results; akz_match picks by problem size:
//   k_match      LOP3/POPC: each thread keeps TWO query descriptors in 32 registers; the block stages tiles of train
//                descriptors in shared memory and every thread walks the tile with broadcast LDS.128 reads.  The 512-bit
//                distance is one carry-save level + 11 POPC, balanced between the ALU pipe and the POPC unit.
//   k_match_mma  tensor cores: popc(q ^ t) = popc(q) + popc(t) - 2 <q, t> as an int8 GEMM (mma.sync m16n8k32, IMMA) on
//                descriptors expanded to one byte per bit in shared memory.  10k x 10k: 0.203 ms against 0.318 ms.
// The train range is split across blockIdx.y so the grid fills whole waves of the 148 SMs; partial results are merged
// by k_match_merge with an associative rule, which is also what the train-sharded multi-GPU path applies to the
// per-shard results after the NCCL gather.
//
// Partial-result encoding (akz_match_t):
//   KNN2   (idx1,dist1) best, (idx2,dist2) second best; lexicographic (distance, index) order
//   COMPAT dist1 = minimum distance, idx1 = lowest index attaining it, idx2 = bit mask of the
//          (index mod 16) classes attaining it.  The reference accepts iff exactly one class attains
//          the minimum and it is < 96; that test is applied once, after the last merge.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

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

// body of the LOP3/POPC kernel: query block bx against the train range [by * per_split, (by + 1) * per_split) of t;
// the partial results of that range go to parts[qi]
template <int MODE>
__device__ __forceinline__ void match_block(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int tbase,
                                            int per_split, akz_match_t* __restrict__ parts, int bx, int by, uint4* tile)
{
    unsigned qa[QPT][16];
    int qi[QPT];
    Best b[QPT];
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        qi[k] = bx * QPB + k * NTH + threadIdx.x;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 w = qi[k] < nq ? __ldg(q + 4 * (long long)qi[k] + v) : make_uint4(0, 0, 0, 0);
            if (v == 3) w.w &= 0x3Fu;                      // bits 486..511 are padding: every matcher kernel ignores them
            qa[k][4 * v] = w.x; qa[k][4 * v + 1] = w.y; qa[k][4 * v + 2] = w.z; qa[k][4 * v + 3] = w.w;
        }
        b[k].k1 = KEY_NONE; b[k].k2 = (MODE == AKZ_MATCH_KNN2) ? KEY_NONE : 0u;
    }
    const int t0 = by * per_split, t1 = min(nt, t0 + per_split);
    for (int base = t0; base < t1; base += TILE) {
        int cnt = min(TILE, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 4; i += NTH) {
            uint4 w = __ldg(t + 4 * (long long)base + i);
            if ((i & 3) == 3) w.w &= 0x3Fu;                // padding bits (as k_match_tc5)
            tile[i] = w;
        }
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
            parts[qi[k]] = m;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTH) k_match(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int tbase,
                                               int per_split, akz_match_t* __restrict__ parts)
{
    __shared__ uint4 tile[TILE * 4];
    match_block<MODE>(q, nq, t, nt, tbase, per_split, parts + (long long)blockIdx.y * nq, blockIdx.x, blockIdx.y, tile);
}

// Batched matching of consecutive frames of a stream (BASELINE configs[4]): pair p = blockIdx.z matches the descriptors of
// frame p + 1 (queries) against those of frame p (train).  The counts are read on the device: no host round trip between the
// detector and the matcher.  desc: [nframes][max_pts][64]; parts: [npairs][nsplit][max_pts].
template <int MODE>
__global__ void __launch_bounds__(NTH) k_match_pairs(const uint4* __restrict__ desc, const int* __restrict__ counts, int max_pts, int nsplit,
                                                     akz_match_t* __restrict__ parts)
{
    __shared__ uint4 tile[TILE * 4];
    const int p = blockIdx.z;
    const int nq = min(counts[p + 1], max_pts), nt = min(counts[p], max_pts);
    if ((int)blockIdx.x * QPB >= nq) return;                        // whole blocks leave before any barrier
    int per = (nt + nsplit - 1) / nsplit;
    per = max(((per + TILE - 1) / TILE) * TILE, TILE);
    match_block<MODE>(desc + (long long)(p + 1) * max_pts * 4, nq, desc + (long long)p * max_pts * 4, nt, 0, per,
                      parts + ((long long)p * nsplit + blockIdx.y) * max_pts, blockIdx.x, blockIdx.y, tile);
}

// =====================================================================================================================
// Tensor-core variant for large problems.  popc(q ^ t) = popc(q) + popc(t) - 2 * <q, t> with the descriptors read as
// 512-long 0/1 vectors, so the pairwise part is an integer GEMM: mma.sync.m16n8k32 (u8 x u8 -> s32, SASS IMMA.16832).
// The 64-byte descriptors are expanded to one byte per bit in shared memory on the fly (nibble * 0x00204081 & 0x01010101
// spreads 4 bits over 4 bytes); expanding in global memory instead would multiply the L2/HBM traffic by 8.
//   block = 16 warps, 256 queries x a range of train descriptors in tiles of 128
//   warp  = 16 queries: its A fragments (16 x 512 bytes) stay in 64 registers for the whole kernel
//   per tile: 16 column blocks of 8 train descriptors x 16 k-steps -> 256 IMMA per warp for 2048 pairs (0.125 per pair),
//             against ~36 ALU + 11 POPC instructions per pair in k_match
// The epilogue keeps the same keyed top-2 / class-mask state as k_match (two query rows per thread), merged across the
// four threads that share a row with shuffles at the end.  Results are bit-identical to k_match (integers).
// =====================================================================================================================
constexpr int MQ = 256, MT = 128, MROW = 528, MNT = 2 * MQ;    // 16 warps x 16 query rows                    // row pitch in bytes: 512 + 16 (conflict-free fragment loads)
constexpr int MM_SMEM = (MQ + MT) * MROW + (MQ + MT) * 4;

__device__ __forceinline__ void mma_u8(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NR packed descriptors = 16 NR words, NR / 32 per thread (512 threads): item i = tid + 512 * k -> descriptor i >> 4, word i & 15
template <int NR>
__device__ __forceinline__ void mm_load(const unsigned* __restrict__ src, int cnt, unsigned (&w)[NR / 32], int tid)
{
#pragma unroll
    for (int k = 0; k < NR / 32; k++) {
        int i = tid + MNT * k, d = i >> 4;
        w[k] = d < cnt ? __ldg(src + (long long)d * 16 + (i & 15)) : 0u;
        if ((i & 15) == 15) w[k] &= 0x3Fu;                 // padding bits 486..511 are ignored (as k_match_tc5)
    }
}
// expand to one byte per bit: rows[d][32 * word + bit], popcounts -> pc[d]
template <int NR>
__device__ __forceinline__ void mm_store(const unsigned (&wv)[NR / 32], unsigned char* rows, int* pc, int tid)
{
#pragma unroll
    for (int k = 0; k < NR / 32; k++) {
        int i = tid + MNT * k, d = i >> 4, wi = i & 15;
        unsigned w = wv[k];
        uint4 lo, hi;
        lo.x = ((w & 0xFu) * 0x00204081u) & 0x01010101u;         lo.y = (((w >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
        lo.z = (((w >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;  lo.w = (((w >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
        hi.x = (((w >> 16) & 0xFu) * 0x00204081u) & 0x01010101u; hi.y = (((w >> 20) & 0xFu) * 0x00204081u) & 0x01010101u;
        hi.z = (((w >> 24) & 0xFu) * 0x00204081u) & 0x01010101u; hi.w = ((w >> 28) * 0x00204081u) & 0x01010101u;
        uint4* dst = reinterpret_cast<uint4*>(rows + d * MROW + wi * 32);
        dst[0] = lo; dst[1] = hi;
        int p = __popc(w);                                        // 16 consecutive lanes hold one descriptor
        p += __shfl_xor_sync(0xffffffffu, p, 1); p += __shfl_xor_sync(0xffffffffu, p, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 4); p += __shfl_xor_sync(0xffffffffu, p, 8);
        if (wi == 0) pc[d] = p;
    }
}

template <int MODE>
__global__ void __launch_bounds__(MNT, 1) k_match_mma(const unsigned* __restrict__ q, int nq, const unsigned* __restrict__ t, int nt, int tbase,
                                                      int per_split, akz_match_t* __restrict__ parts)
{
    extern __shared__ __align__(16) unsigned char msm[];
    unsigned char* Qs = msm;
    unsigned char* Ts = msm + MQ * MROW;
    int* pqs = reinterpret_cast<int*>(msm + (MQ + MT) * MROW);
    int* pts = pqs + MQ;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int q0 = blockIdx.x * MQ;
    {
        unsigned wq[MQ / 32];
        mm_load<MQ>(q + (long long)q0 * 16, min(MQ, nq - q0), wq, tid);
        mm_store<MQ>(wq, Qs, pqs, tid);
    }
    unsigned wv[MT / 32];
    __syncthreads();
    unsigned a[16][4];
    {
        const unsigned char* r0 = Qs + (wid * 16 + gid) * MROW + tig * 4;
        const unsigned char* r1 = r0 + 8 * MROW;
#pragma unroll
        for (int ks = 0; ks < 16; ks++) {
            a[ks][0] = *reinterpret_cast<const unsigned*>(r0 + ks * 32);
            a[ks][1] = *reinterpret_cast<const unsigned*>(r1 + ks * 32);
            a[ks][2] = *reinterpret_cast<const unsigned*>(r0 + ks * 32 + 16);
            a[ks][3] = *reinterpret_cast<const unsigned*>(r1 + ks * 32 + 16);
        }
    }
    const int pq0 = pqs[wid * 16 + gid], pq1 = pqs[wid * 16 + gid + 8];
    Best b0, b1;
    b0.k1 = b1.k1 = KEY_NONE; b0.k2 = b1.k2 = (MODE == AKZ_MATCH_KNN2) ? KEY_NONE : 0u;
    const int t0 = blockIdx.y * per_split, t1 = min(nt, t0 + per_split);
    // Two tile buffers: the query tile's bytes are dead once the A fragments sit in registers, so its region becomes the
    // second buffer.  A warp that has finished the MMAs of tile i expands tile i+1 into the other buffer straight away
    // (its words were prefetched during the MMAs), so expansion overlaps the other warps' tensor work: one barrier per tile.
    __syncthreads();                                              // every warp holds its A fragments and row popcounts
    if (t0 < t1) {
        mm_load<MT>(t + (long long)t0 * 16, min(MT, t1 - t0), wv, tid);
        mm_store<MT>(wv, Ts, pts, tid);
        if (t0 + MT < t1) mm_load<MT>(t + (long long)(t0 + MT) * 16, min(MT, t1 - t0 - MT), wv, tid);
    }
    __syncthreads();
    int p = 0;
    for (int base = t0; base < t1; base += MT, p ^= 1) {
        const int cnt = min(MT, t1 - base);
        const unsigned char* Tcur = p ? Qs : Ts;
        const int* ptc = p ? pqs : pts;
        // four column blocks (4 x 8 train descriptors) at a time: four independent accumulator chains per warp
        for (int nb = 0; nb < MT / 8; nb += 4) {
            if (nb * 8 >= cnt) break;
            int c[4][4];
#pragma unroll
            for (int u = 0; u < 4; u++) { c[u][0] = 0; c[u][1] = 0; c[u][2] = 0; c[u][3] = 0; }
            const unsigned char* br = Tcur + (nb * 8 + gid) * MROW + tig * 4;
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    mma_u8(c[u], a[ks], *reinterpret_cast<const unsigned*>(br + u * 8 * MROW + ks * 32),
                           *reinterpret_cast<const unsigned*>(br + u * 8 * MROW + ks * 32 + 16));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = (nb + u) * 8 + tig * 2;             // columns j, j+1 of the tile
                const int pt0 = ptc[j], pt1 = ptc[j + 1];
                const int jrel = base - t0 + j;
                const unsigned cb0 = 1u << ((tbase + base + j) & 15), cb1 = 1u << ((tbase + base + j + 1) & 15);
                if (j < cnt) {
                    consider<MODE>(b0, pq0 + pt0 - 2 * c[u][0], jrel, cb0);
                    consider<MODE>(b1, pq1 + pt0 - 2 * c[u][2], jrel, cb0);
                }
                if (j + 1 < cnt) {
                    consider<MODE>(b0, pq0 + pt1 - 2 * c[u][1], jrel + 1, cb1);
                    consider<MODE>(b1, pq1 + pt1 - 2 * c[u][3], jrel + 1, cb1);
                }
            }
        }
        if (base + MT < t1) {
            mm_store<MT>(wv, p ? Ts : Qs, p ? pts : pqs, tid);
            if (base + 2 * MT < t1) mm_load<MT>(t + (long long)(base + 2 * MT) * 16, min(MT, t1 - base - 2 * MT), wv, tid);
        }
        __syncthreads();
    }
    // merge the four threads that share a query row
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            Best& b = r ? b1 : b0;
            unsigned o1 = __shfl_xor_sync(0xffffffffu, b.k1, d), o2 = __shfl_xor_sync(0xffffffffu, b.k2, d);
            if (MODE == AKZ_MATCH_KNN2) {
                unsigned hi = max(b.k1, o1);
                b.k1 = min(b.k1, o1);
                b.k2 = min(min(b.k2, o2), hi);
            } else {
                unsigned da = b.k1 >> KEY_IDX_BITS, db = o1 >> KEY_IDX_BITS;
                b.k2 = da < db ? b.k2 : (da == db ? (b.k2 | o2) : o2);
                b.k1 = min(b.k1, o1);
            }
        }
    }
    if (tig == 0) {
        const unsigned imask = (1u << KEY_IDX_BITS) - 1;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const Best& b = r ? b1 : b0;
            int qi = q0 + wid * 16 + gid + 8 * r;
            if (qi < nq) {
                akz_match_t m;
                const bool has1 = b.k1 != KEY_NONE;
                m.idx1 = has1 ? tbase + t0 + (int)(b.k1 & imask) : -1;
                m.dist1 = has1 ? (int)(b.k1 >> KEY_IDX_BITS) : -1;
                if (MODE == AKZ_MATCH_KNN2) {
                    const bool has2 = b.k2 != KEY_NONE;
                    m.idx2 = has2 ? tbase + t0 + (int)(b.k2 & imask) : -1;
                    m.dist2 = has2 ? (int)(b.k2 >> KEY_IDX_BITS) : -1;
                } else {
                    m.idx2 = (int)b.k2; m.dist2 = 0;
                }
                parts[(long long)blockIdx.y * nq + qi] = m;
            }
        }
    }
}

__device__ __forceinline__ bool lex_less(int d, int i, int d2, int i2) { return d < d2 || (d == d2 && i < i2); }

__global__ void k_match_merge(const akz_match_t* __restrict__ parts, int nparts, int nq, int mode, int finalize, akz_match_t* __restrict__ out)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    // launched with programmatic stream serialisation (match_merge below): the blocks may become resident while the kernel that
    // writes `parts` is still running; its results are visible after this wait (a no-op under a plain launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (qi >= nq) return;
    akz_match_t r;
    const bool top2 = mode != AKZ_MATCH_COMPAT;                // KNN2 and UNIQUE2 share the partial form
    r.idx1 = -1; r.dist1 = -1; r.idx2 = top2 ? -1 : 0; r.dist2 = top2 ? -1 : 0;
    // the partial results of a query are 16-byte records nq apart: fetch them eight at a time (independent 128-bit loads in
    // flight) before the serial merge -- one load per iteration left the kernel latency bound (9.7 us for 10 x 10k records)
    const int4* p4 = reinterpret_cast<const int4*>(parts);
    for (int p0 = 0; p0 < nparts; p0 += 8) {
        int4 buf[8];
#pragma unroll
        for (int k = 0; k < 8; k++) buf[k] = p0 + k < nparts ? __ldg(p4 + (long long)(p0 + k) * nq + qi) : make_int4(-1, -1, -1, -1);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            akz_match_t m;
            m.idx1 = buf[k].x; m.dist1 = buf[k].y; m.idx2 = buf[k].z; m.dist2 = buf[k].w;
            if (m.idx1 < 0) continue;
            if (top2) {
                // merge two sorted pairs, lowest (distance, index) first
                int cd[4] = { r.dist1, r.dist2, m.dist1, m.dist2 };
                int ci[4] = { r.idx1, r.idx2, m.idx1, m.idx2 };
                int bd1 = 1 << 20, bi1 = -1, bd2 = 1 << 20, bi2 = -1;
                for (int c = 0; c < 4; c++) {
                    if (ci[c] < 0) continue;
                    if (bi1 < 0 || lex_less(cd[c], ci[c], bd1, bi1)) { bd2 = bd1; bi2 = bi1; bd1 = cd[c]; bi1 = ci[c]; }
                    else if (bi2 < 0 || lex_less(cd[c], ci[c], bd2, bi2)) { bd2 = cd[c]; bi2 = ci[c]; }
                }
                r.idx1 = bi1; r.dist1 = bi1 < 0 ? -1 : bd1; r.idx2 = bi2; r.dist2 = bi2 < 0 ? -1 : bd2;
            } else {
                if (r.idx1 < 0 || m.dist1 < r.dist1) r = m;
                else if (m.dist1 == r.dist1) { r.idx1 = min(r.idx1, m.idx1); r.idx2 |= m.idx2; }
            }
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

// merge of k_match_pairs: out[(p + 1) * max_pts + qi] for the queries of frame p + 1 (acceptance rule applied)
__global__ void k_match_merge_pairs(const akz_match_t* __restrict__ parts, const int* __restrict__ counts, int max_pts, int nsplit, int mode,
                                    akz_match_t* __restrict__ out)
{
    const int p = blockIdx.y;
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = min(counts[p + 1], max_pts);
    if (qi >= nq) return;
    const akz_match_t* pp = parts + (long long)p * nsplit * max_pts;
    akz_match_t r;
    const bool top2 = mode != AKZ_MATCH_COMPAT;
    r.idx1 = -1; r.dist1 = -1; r.idx2 = top2 ? -1 : 0; r.dist2 = top2 ? -1 : 0;
    for (int k = 0; k < nsplit; k++) {
        const int4 b = __ldg(reinterpret_cast<const int4*>(pp + (long long)k * max_pts + qi));
        akz_match_t m;
        m.idx1 = b.x; m.dist1 = b.y; m.idx2 = b.z; m.dist2 = b.w;
        if (m.idx1 < 0) continue;
        if (top2) {
            int cd[4] = { r.dist1, r.dist2, m.dist1, m.dist2 };
            int ci[4] = { r.idx1, r.idx2, m.idx1, m.idx2 };
            int bd1 = 1 << 20, bi1 = -1, bd2 = 1 << 20, bi2 = -1;
            for (int c = 0; c < 4; c++) {
                if (ci[c] < 0) continue;
                if (bi1 < 0 || lex_less(cd[c], ci[c], bd1, bi1)) { bd2 = bd1; bi2 = bi1; bd1 = cd[c]; bi1 = ci[c]; }
                else if (bi2 < 0 || lex_less(cd[c], ci[c], bd2, bi2)) { bd2 = cd[c]; bi2 = ci[c]; }
            }
            r.idx1 = bi1; r.dist1 = bi1 < 0 ? -1 : bd1; r.idx2 = bi2; r.dist2 = bi2 < 0 ? -1 : bd2;
        } else {
            if (r.idx1 < 0 || m.dist1 < r.dist1) r = m;
            else if (m.dist1 == r.dist1) { r.idx1 = min(r.idx1, m.idx1); r.idx2 |= m.idx2; }
        }
    }
    if (mode == AKZ_MATCH_UNIQUE2) {
        bool ok = r.idx1 >= 0 && r.dist1 < AKZ_MAX_DIST && (r.idx2 < 0 || r.dist1 < r.dist2);
        if (!ok) { r.idx1 = -1; r.dist1 = -1; }
    }
    if (mode == AKZ_MATCH_COMPAT) {
        bool ok = r.idx1 >= 0 && __popc((unsigned)r.idx2) == 1 && r.dist1 < AKZ_MAX_DIST;
        if (!ok) { r.idx1 = -1; r.dist1 = -1; }
    }
    out[(long long)(p + 1) * max_pts + qi] = r;
}

}  // namespace

namespace akzk {

// consecutive-frame matching of a batch (see k_match_pairs); parts must hold npairs * nsplit * max_pts records
int match_pairs(cudaStream_t st, const unsigned char* desc, const int* counts, int nframes, int max_pts, int mode, int nsplit,
                akz_match_t* parts, akz_match_t* out)
{
    const int npairs = nframes - 1;
    if (npairs <= 0) return 0;
    if (max_pts >= (1 << KEY_IDX_BITS)) return akz_set_error(AKZ_E_UNSUPPORTED, "max_pts too large for the batched matcher");
    dim3 g((max_pts + QPB - 1) / QPB, nsplit, npairs);
    if (mode != AKZ_MATCH_COMPAT) k_match_pairs<AKZ_MATCH_KNN2><<<g, NTH, 0, st>>>((const uint4*)desc, counts, max_pts, nsplit, parts);
    else k_match_pairs<AKZ_MATCH_COMPAT><<<g, NTH, 0, st>>>((const uint4*)desc, counts, max_pts, nsplit, parts);
    k_match_merge_pairs<<<dim3((max_pts + 127) / 128, npairs), 128, 0, st>>>(parts, counts, max_pts, nsplit, mode, out);
    return 2;
}

// use_mma != 0: tensor-core kernel (large problems); the split is in units of 128 train descriptors either way
int match_partial(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode,
                  int nsplit, akz_match_t* parts, int use_mma)
{
    if (nq <= 0) return 0;
    int per = (nt + nsplit - 1) / nsplit;
    per = ((per + TILE - 1) / TILE) * TILE;
    if (per <= 0) per = TILE;
    if (per >= (1 << KEY_IDX_BITS)) return akz_set_error(AKZ_E_UNSUPPORTED, "train range per block exceeds 2^22 descriptors: shard the train set");
    if (use_mma) {
        static akz_once_t attr;
        if (akz_once_guard once{attr}) {
            cudaFuncSetAttribute(k_match_mma<AKZ_MATCH_KNN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
            cudaFuncSetAttribute(k_match_mma<AKZ_MATCH_COMPAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
        }
        dim3 g((nq + MQ - 1) / MQ, nsplit);
        if (mode != AKZ_MATCH_COMPAT)
            k_match_mma<AKZ_MATCH_KNN2><<<g, MNT, MM_SMEM, st>>>((const unsigned*)q, nq, (const unsigned*)t, nt, tbase, per, parts);
        else
            k_match_mma<AKZ_MATCH_COMPAT><<<g, MNT, MM_SMEM, st>>>((const unsigned*)q, nq, (const unsigned*)t, nt, tbase, per, parts);
        return 1;
    }
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
    // programmatic dependent launch: k_match_tc5 releases its dependents when it starts (griddepcontrol.launch_dependents), so the
    // merge blocks are resident and waiting when the last tile is written -- the launch gap of a 3 us kernel behind a 38 us one
    static const bool pdl = [] { const char* e = getenv("AKZ_MATCH_PDL"); return !e || atoi(e) != 0; }();      // A/B knob
    if (!pdl) {
        k_match_merge<<<(nq + 127) / 128, 128, 0, st>>>(parts, nparts, nq, mode, finalize, out);
        return 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((nq + 127) / 128); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_match_merge, parts, nparts, nq, mode, finalize, out) != cudaSuccess) {
        cudaGetLastError();
        k_match_merge<<<(nq + 127) / 128, 128, 0, st>>>(parts, nparts, nq, mode, finalize, out);
    }
    return 1;
}

}  // namespace akzk
