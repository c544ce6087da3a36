// k_match_tc5: brute-force Hamming matcher on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// tensor memory).  Same contract and same integer results as k_match / k_match_mma (match.cu); replaces gHammingMatch
// (akazed.cu:2144-2241) and gMatch (akazed.cu:2028-2122) for large problems.
//
//   popc(q ^ t) = popc(q) + popc(t) - 2 <q, t>      with the descriptors read as 512-long 0/1 vectors
//
// so the pairwise part is a u8 x u8 -> s32 GEMM with K = 512.  One CTA per SM owns 256 queries (two operand-A tiles of
// MMA M = 128) and walks a range of train descriptors in tiles of 128 (MMA N); a train tile is 2 x 16 tcgen05.mma of
// 128 x 128 x 32 into two accumulators.
//
// Warp roles (13 warps, no block-wide barrier inside the main loop; everything is handed over through mbarriers):
//   warps 9-12  expanders: read packed descriptors (64 B) and write them one byte per bit into shared memory in the
//               K-major SWIZZLE_128B operand layout the tensor core reads (expanding in global memory would multiply
//               the L2/HBM traffic by 8).  The unit is a K-block: 128 rows x 128 bytes = bits [128 kb, 128 kb + 128) of
//               every descriptor of the tile, 16 KB; four K-block slots form a ring (one tile of look-ahead).  A slot
//               feeds both query tiles, which halves the expansion work and the shared-memory stores per MMA.
//   warp 8      one thread issues the MMAs: per K-block 2 x 4 tcgen05.mma (K = 32 bytes each), tcgen05.commit releases
//               the slot; after the fourth K-block a second commit publishes the accumulator pair.
//   warps 0-7   epilogue: thread = query row; warp w reads TMEM lanes 32 (w % 4) ... of accumulator w / 4.  tcgen05.ld
//               brings 32 accumulator columns at a time (the next load is in flight while a chunk is processed); the
//               running top-2 (or minimum + class mask) is kept on keys
//                     key = (popc(t) + 512 - 2 dot) << 20 | (train index relative to the CTA's range)
//               = ptk[column] - (dot << 21): one integer instruction per pair, then a 3-instruction min/max network on
//               two independent chains (even / odd columns).  popc(q) is constant along a row and is added once at the
//               end.  Two accumulator-pair buffers (2 x 256 TMEM columns) decouple the epilogue from the MMAs.
// Any fixed permutation of the 512 bits gives the same dot product as long as both operands use it, so the expansion uses
// the cheapest one: output word = (input word >> j) & 0x01010101.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int T5_M = 128, T5_N = 128;
constexpr int T5_QT = 2;                         // operand-A tiles (128 queries each) per CTA
constexpr int T5_Q = T5_QT * T5_M;               // queries per CTA
constexpr int T5_SLOTS = 4;                      // ring of K-block slots
constexpr int T5_NACC = 2;                       // accumulator-pair buffers in tensor memory
constexpr int T5_KB = 128 * 128;                 // bytes of one K-block slot
constexpr int T5_NEPI = T5_Q, T5_NEXP = 128;
constexpr int T5_NT = T5_NEPI + 32 + T5_NEXP;    // 416 threads
constexpr int T5_MMA_WARP = T5_NEPI / 32;
static_assert(T5_SLOTS == 4, "the expander and issuer loops use slot = K-block");
constexpr int T5_IDX_BITS = 20;
constexpr unsigned T5_NONE = 0xFFFFFFFFu;
constexpr int T5_DOFF = 512;                     // keeps popc(t) - 2 dot non-negative

constexpr int T5_OFF_A = 0;
constexpr int T5_OFF_RING = T5_QT * 4 * T5_KB;
constexpr int T5_OFF_PTK = T5_OFF_RING + T5_SLOTS * T5_KB;
constexpr int T5_OFF_BAR = T5_OFF_PTK + T5_NACC * T5_N * 4;
constexpr int T5_NBAR = 2 * T5_SLOTS + 2 * T5_NACC;
constexpr int T5_OFF_TMEM = T5_OFF_BAR + T5_NBAR * 8;
constexpr int T5_SMEM = T5_OFF_TMEM + 16 + 1024;             // + slack for the 1024-byte alignment of the operand tiles

// instruction descriptor (kind::i8): D = s32, A = B = u8, both K-major, N = 128, M = 128
constexpr unsigned T5_IDESC = (2u << 4) | ((unsigned)(T5_N >> 3) << 17) | ((unsigned)(T5_M >> 4) << 24);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory operand descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (SBO), LBO = 1 (unused), version 1
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr)
{
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(T5_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 128 bits of descriptor `row` -> 128 bytes (0 / 1) of row `row` of a K-block, SWIZZLE_128B: the 16-byte chunk j of a row sits
// at chunk position j ^ (row & 7).  Chunk j = { (x.x >> j) & M, (x.y >> j) & M, (x.z >> j) & M, (x.w >> j) & M }.
__device__ __forceinline__ void expand_store(unsigned char* kblock, int row, uint4 x)
{
    unsigned char* rp = kblock + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint4 o;
        o.x = (x.x >> j) & 0x01010101u; o.y = (x.y >> j) & 0x01010101u;
        o.z = (x.z >> j) & 0x01010101u; o.w = (x.w >> j) & 0x01010101u;
        *reinterpret_cast<uint4*>(rp + ((j ^ sw) << 4)) = o;
    }
}

struct Best5 { unsigned k1, k2; };

template <int MODE>
__device__ __forceinline__ void consider5(Best5& b, unsigned key, unsigned classbit)
{
    if (MODE == AKZ_MATCH_KNN2) {
        const unsigned hi = max(b.k1, key);
        b.k1 = min(b.k1, key);
        b.k2 = min(b.k2, hi);
    } else {
        const unsigned d = key >> T5_IDX_BITS, dcur = b.k1 >> T5_IDX_BITS;
        b.k2 = d < dcur ? classbit : (d == dcur ? (b.k2 | classbit) : b.k2);
        b.k1 = min(b.k1, key);
    }
}

// merge of two top-2 states (KNN2) or two (minimum, class mask) states (COMPAT)
template <int MODE>
__device__ __forceinline__ void merge5(Best5& a, const Best5& o)
{
    if (MODE == AKZ_MATCH_KNN2) {
        const unsigned hi = max(a.k1, o.k1);
        a.k1 = min(a.k1, o.k1);
        a.k2 = min(min(a.k2, o.k2), hi);
    } else {
        const unsigned da = a.k1 >> T5_IDX_BITS, db = o.k1 >> T5_IDX_BITS;
        a.k2 = da < db ? a.k2 : (da == db ? (a.k2 | o.k2) : o.k2);
        a.k1 = min(a.k1, o.k1);
    }
}

template <int MODE>
__global__ void __launch_bounds__(T5_NT, 1) k_match_tc5(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int tbase,
                                                        int per_split, akz_match_t* __restrict__ parts)
{
    extern __shared__ unsigned char t5raw[];
    unsigned char* sm = t5raw + ((1024u - (smem_u32(t5raw) & 1023u)) & 1023u);
    unsigned char* As = sm + T5_OFF_A;
    unsigned char* Ring = sm + T5_OFF_RING;
    unsigned* ptk = reinterpret_cast<unsigned*>(sm + T5_OFF_PTK);
    const unsigned bar0 = smem_u32(sm + T5_OFF_BAR);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(sm + T5_OFF_TMEM);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (T5_SLOTS + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * T5_SLOTS + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * T5_SLOTS + T5_NACC + b); };

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * T5_Q;
    const int t0 = blockIdx.y * per_split, t1 = min(nt, t0 + per_split);
    const int ntiles = t1 > t0 ? (t1 - t0 + T5_N - 1) / T5_N : 0;

    // ---- prologue: barriers, tensor memory, the query tiles as operand A ---------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < T5_SLOTS; s++) { mbar_init(bar_full(s), T5_NEXP); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < T5_NACC; b++) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), T5_NEPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == T5_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 4 * T5_Q; i += T5_NT) {
        const int row = i & (T5_Q - 1), kb = i / T5_Q;                 // query row of the CTA, K-block
        const uint4 x = q0 + row < nq ? __ldg(q + 4 * (long long)(q0 + row) + kb) : make_uint4(0, 0, 0, 0);
        expand_store(As + ((row >> 7) * 4 + kb) * T5_KB, row & 127, x);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    if (wid < T5_MMA_WARP) {
        // ================================ epilogue: thread = query row ==========================================
        const int row = tid;                                         // A tile row >> 7, TMEM lane row & 127
        int pq = 0;
        if (q0 + row < nq) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 w = __ldg(q + 4 * (long long)(q0 + row) + v);
                pq += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            }
        }
        Best5 be, bo;                                                // even / odd columns: two independent chains
        be.k1 = bo.k1 = T5_NONE; be.k2 = bo.k2 = (MODE == AKZ_MATCH_KNN2) ? T5_NONE : 0u;
        const unsigned cb0 = (unsigned)tbase & 15u;                  // t0 and the tile width are multiples of 16
        for (int j = 0; j < ntiles; j++) {
            const int b = j % T5_NACC;
            // key base of the tile's columns (the first 128 epilogue threads serve one column each)
            if (tid < T5_N) {
                const int col = t0 + j * T5_N + tid;
                unsigned key = T5_NONE;
                if (col < t1) {
                    int pt = 0;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const uint4 w = __ldg(t + 4 * (long long)col + v);
                        pt += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
                    }
                    key = ((unsigned)(pt + T5_DOFF) << T5_IDX_BITS) | (unsigned)(j * T5_N + tid);
                }
                ptk[b * T5_N + tid] = key;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(T5_NEPI) : "memory");
            mbar_wait(bar_tfull(b), (unsigned)(j / T5_NACC) & 1u);
            tc_fence_after();
            const unsigned taddr = tmem + ((unsigned)((wid & 3) * 32) << 16) + (unsigned)(b * T5_QT * T5_N + (wid >> 2) * T5_N);
            const uint4* pk4 = reinterpret_cast<const uint4*>(ptk + b * T5_N);
            unsigned va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < T5_N / 32; c++) {
                unsigned (&v)[32] = (c & 1) ? vb : va;
                unsigned (&vn)[32] = (c & 1) ? va : vb;
                if (c + 1 < T5_N / 32) tmem_ld32(taddr + (c + 1) * 32, vn);        // in flight while this chunk is processed
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint4 k4 = pk4[c * 8 + g];
                    const unsigned kk[4] = { k4.x, k4.y, k4.z, k4.w };
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int i = 4 * g + e;
                        const unsigned key = kk[e] - (v[i] << (T5_IDX_BITS + 1));
                        consider5<MODE>((e & 1) ? bo : be, key, 1u << ((cb0 + i) & 15u));
                    }
                }
                if (c + 1 < T5_N / 32) tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(bar_tempty(b));
        }
        merge5<MODE>(be, bo);
        if (q0 + row < nq) {
            const unsigned imask = (1u << T5_IDX_BITS) - 1;
            akz_match_t m;
            const bool has1 = be.k1 != T5_NONE;
            m.idx1 = has1 ? tbase + t0 + (int)(be.k1 & imask) : -1;
            m.dist1 = has1 ? (int)(be.k1 >> T5_IDX_BITS) - T5_DOFF + pq : -1;
            if (MODE == AKZ_MATCH_KNN2) {
                const bool has2 = be.k2 != T5_NONE;
                m.idx2 = has2 ? tbase + t0 + (int)(be.k2 & imask) : -1;
                m.dist2 = has2 ? (int)(be.k2 >> T5_IDX_BITS) - T5_DOFF + pq : -1;
            } else {
                m.idx2 = has1 ? (int)be.k2 : 0; m.dist2 = 0;
            }
            parts[(long long)blockIdx.y * nq + q0 + row] = m;
        }
    } else if (wid == T5_MMA_WARP) {
        // ================================ MMA issue: one thread ==================================================
        if (lane == 0) {
            const unsigned long long adesc0 = umma_desc(smem_u32(As));
            const unsigned long long bdesc0 = umma_desc(smem_u32(Ring));
            for (int j = 0; j < ntiles; j++) {
                const int b = j % T5_NACC;
                mbar_wait(bar_tempty(b), ((unsigned)(j / T5_NACC) & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < 4; kb++) {
                    const int s = kb;
                    mbar_wait(bar_full(s), (unsigned)j & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < T5_QT; h++) {
                        const unsigned d_tmem = tmem + (unsigned)(b * T5_QT * T5_N + h * T5_N);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma_i8(d_tmem, adesc0 + (unsigned long long)((h * 4 + kb) * (T5_KB >> 4) + 2 * k),
                                    bdesc0 + (unsigned long long)(s * (T5_KB >> 4) + 2 * k), (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(bar_empty(s));
                }
                umma_commit(bar_tfull(b));
            }
        }
        __syncwarp();
    } else {
        // ================================ expanders: thread = row of the train tile, all four K-blocks ===============
        const int row = tid - (T5_NEPI + 32);
        uint4 x[4], nx[4];
#pragma unroll
        for (int kb = 0; kb < 4; kb++) x[kb] = (ntiles > 0 && t0 + row < t1) ? __ldg(t + 4 * (long long)(t0 + row) + kb) : make_uint4(0, 0, 0, 0);
        for (int j = 0; j < ntiles; j++) {
            const int nrow = t0 + (j + 1) * T5_N + row;
#pragma unroll
            for (int kb = 0; kb < 4; kb++) nx[kb] = (j + 1 < ntiles && nrow < t1) ? __ldg(t + 4 * (long long)nrow + kb) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                mbar_wait(bar_empty(kb), ((unsigned)j & 1u) ^ 1u);                  // T5_SLOTS == 4: slot = K-block
                expand_store(Ring + kb * T5_KB, row, x[kb]);
                fence_async_smem();
                mbar_arrive(bar_full(kb));
            }
#pragma unroll
            for (int kb = 0; kb < 4; kb++) x[kb] = nx[kb];
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (wid == T5_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace

namespace akzk {

// Tensor-memory matcher; `per` (train descriptors per blockIdx.y) is a multiple of 128 chosen by the caller.
int match_partial_tc5(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode,
                      int nsplit, akz_match_t* parts)
{
    if (nq <= 0) return 0;
    int per = (nt + nsplit - 1) / nsplit;
    per = ((per + T5_N - 1) / T5_N) * T5_N;
    if (per <= 0) per = T5_N;
    if (per >= (1 << T5_IDX_BITS)) return akz_set_error(AKZ_E_UNSUPPORTED, "train range per block exceeds 2^20 descriptors: raise the split or shard the train set");
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_KNN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_COMPAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
        attr = true;
    }
    dim3 g((nq + T5_Q - 1) / T5_Q, nsplit);
    if (mode != AKZ_MATCH_COMPAT)
        k_match_tc5<AKZ_MATCH_KNN2><<<g, T5_NT, T5_SMEM, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    else
        k_match_tc5<AKZ_MATCH_COMPAT><<<g, T5_NT, T5_SMEM, st>>>((const uint4*)q, nq, (const uint4*)t, nt, tbase, per, parts);
    return 1;
}

}  // namespace akzk
