// k_match_tc5: brute-force Hamming matcher on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// tensor memory).  Same contract and same integer results as k_match / k_match_mma (match.cu); replaces gHammingMatch
// (akazed.cu:2144-2241) and gMatch (akazed.cu:2028-2122) for large problems.
//
//   popc(q ^ t) = popc(q) + popc(t) - 2 <q, t>      with the descriptors read as 512-long 0/1 vectors
//
// so the pairwise part is a u8 x u8 -> s32 GEMM with K = 512.  The grid is one persistent CTA per SM; a CTA works on one item
// at a time -- 256 queries (two operand-A tiles of MMA M = 128) against a range of train descriptors in tiles of 128 (MMA N);
// a train tile is 2 x 16 tcgen05.mma of 128 x 128 x 32 into two accumulators (see "Work decomposition" at the kernel).
//
// Warp roles (13 warps, no block-wide barrier inside the main loop; everything is handed over through mbarriers):
//   warps 9-12  expanders: read packed descriptors (64 B, two tiles of look-ahead in registers) and write them one byte per
//               bit into shared memory in the K-major SWIZZLE_128B operand layout the tensor core reads (expanding in global
//               memory would multiply the L2/HBM traffic by 8).  The unit is a K-block: 128 rows x 128 bytes = bits
//               [128 kb, 128 kb + 128) of every descriptor of the tile, 16 KB; four K-block slots form a ring.  A slot feeds
//               both query tiles, which halves the expansion work and the shared-memory stores per MMA.
//   warp 8      issues the MMAs: per K-block 2 x 4 tcgen05.mma (K = 32 bytes each), tcgen05.commit releases the slot; after
//               the fourth K-block a second commit publishes the accumulator pair.  The whole warp runs the loop and
//               elect.sync picks the issuing lane inside the asm statement (see umma_i8).
//   warps 0-7   epilogue: thread = query row; warp w reads TMEM lanes 32 (w % 4) ... of accumulator w / 4.  A packed
//               tcgen05.ld brings 64 accumulator columns (two per register) with the next load in flight.  The padding
//               positions of the GEMM carry popc(t) and the column index, so an accumulator IS the sortable part of a key
//               (see "keys"): the top-2 of a chunk is taken on the packed 16-bit values and only the winners are unpacked.
//               Two accumulator-pair buffers (2 x 256 TMEM columns) decouple the epilogue from the MMAs.
// Any fixed permutation of the 512 bits gives the same dot product as long as both operands use it, and the operand bytes
// need not be 0 / 1 as long as every product of two set bits is the same constant: see expand_store.
// For long train ranges (FILTER) the epilogue first takes the maximum of a 64-column chunk with a packed VIMNMX3 tree and runs
// the exact top-2 only when that maximum can change the state (a new top-2 entry is rare once a few thousand candidates have
// been seen).  The reference-compatible mode needs one update per chunk in any case.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

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
constexpr int T5_RANGE_UNIT = 1024;              // the train range of a CTA is a multiple of this (16 half tiles: see the class bits)
constexpr int T5_MAX_RANGE = 8192 * 64;          // 13-bit half-tile ordinal

constexpr int T5_OFF_A = 0;
constexpr int T5_OFF_RING = T5_QT * 4 * T5_KB;
constexpr int T5_OFF_BAR = T5_OFF_RING + T5_SLOTS * T5_KB;
constexpr int T5_NBAR = 2 * T5_SLOTS + 2 * T5_NACC + 1;
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
// spinning wait (test_wait never suspends the thread): for the MMA issuer and the expanders, whose wake-up latency is on the
// critical path of the slot ring (a slot is refilled while the three other K-blocks of the tile are multiplied)
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_spin(unsigned bar, unsigned parity)
{
    if (mbar_test(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_test(bar, parity)) {
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
// Issued by a whole, converged warp: one elected lane executes the instruction.  With warp-uniform control flow and operands
// the compiler keeps descriptors and addresses in uniform registers; issued from inside `if (lane == 0)` every MMA cost four
// R2UR moves and a waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY), about as long as the 64 clocks the MMA itself takes.
__device__ __forceinline__ void umma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(T5_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
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
// the same load with two adjacent columns packed into one register (their low 16 bits): the accumulators are at most
// 64 * 998 + 63 < 2^16, and the TMEM -> register traffic shares the SM's 128 B/clk memory datapath with the operand reads of the
// tensor core and the expanders' stores (ncu r01j: the three add up to the kernel's duration), so half the bytes is time won
__device__ __forceinline__ void tmem_ld32_pack16(unsigned taddr, unsigned (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 128 bits of descriptor `row` -> 128 bytes of row `row` of a K-block, SWIZZLE_128B: the 16-byte chunk j of a row sits at chunk
// position j ^ (row & 7).  Chunk j carries bit (8 b + j) of each of the four words in byte b.  The operands are weighted so
// that the train side needs no shift: a train byte is 0 or 2^j (a plain AND with a mask), a query byte is 0 or 2^(7-j), and a
// product is 0 or 128: the accumulator receives 128 <q, t>.
// `extra` (last K-block only) is OR-ed into the bytes of the four padding bits 486..489, see the key layout below.
template <bool TRAIN>
__device__ __forceinline__ void expand_store(unsigned char* kblock, int row, uint4 x, bool last, unsigned e6, unsigned e7, unsigned e0, unsigned e1)
{
    unsigned char* rp = kblock + row * 128;
    const int sw = row & 7;
    if (last) x.w &= 0x3Fu;                                       // bits 486..511 are padding (zero by contract): ignored
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint4 o;
        if (TRAIN) {
            const unsigned m = 0x01010101u << j;
            o.x = x.x & m; o.y = x.y & m; o.z = x.z & m; o.w = x.w & m;
        } else {
            const unsigned m = 0x01010101u << (7 - j);            // bit j of a byte moves to bit 7 - j of the same byte
            if (7 - 2 * j >= 0) { o.x = (x.x << (7 - 2 * j)) & m; o.y = (x.y << (7 - 2 * j)) & m; o.z = (x.z << (7 - 2 * j)) & m; o.w = (x.w << (7 - 2 * j)) & m; }
            else { o.x = (x.x >> (2 * j - 7)) & m; o.y = (x.y >> (2 * j - 7)) & m; o.z = (x.z >> (2 * j - 7)) & m; o.w = (x.w >> (2 * j - 7)) & m; }
        }
        if (last) {
            if (j == 6) o.w |= e6;
            if (j == 7) o.w |= e7;
            if (j == 0) o.w |= e0 << 8;
            if (j == 1) o.w |= e1 << 8;
        }
        *reinterpret_cast<uint4*>(rp + ((j ^ sw) << 4)) = o;
    }
}

// ---- keys ----------------------------------------------------------------------------------------------------------------
// The four padding positions carry, on the query side, the constants 64, 64, 64, 1 and, on the train side, three bytes that sum
// to 512 - popc(t) and the byte 63 - c (c = column within its half tile of 64), so the tensor core itself delivers
//        acc = 128 dot + 64 (512 - popc(t)) + (63 - c) = 64 E + (63 - c),     E = 512 - (popc(t) - 2 dot)  in [26, 1024]
// and the epilogue needs ONE integer multiply-add per pair and no shared-memory traffic:
//        key = acc << 13 | (8191 - ordinal of the half tile)            bits [19,30) E, [13,19) 63 - c, [0,13) 8191 - ordinal
// The train index inside the CTA's range is  rel = c * H + ordinal  (H = number of half tiles of the range): the column is the
// MOST significant part, so the unsigned order of the keys is exactly (smaller distance, then smaller train index) and the
// running top-2 is a max network.  A row of zeros (beyond the range) gives key < 8192 = "none".
// Hamming distance = popc(q) + 512 - E.
constexpr unsigned T5_EMASK = (1u << 19) - 1;

struct Best5 { unsigned k1, k2; };

template <int MODE>
__device__ __forceinline__ void consider5(Best5& b, unsigned key, unsigned classbit)
{
    if (MODE == AKZ_MATCH_KNN2) {
        const unsigned lo = min(b.k1, key);
        b.k1 = max(b.k1, key);
        b.k2 = max(b.k2, lo);
    } else {
        // E(key) > E(k1)  <=>  key > (k1 | EMASK);   E(key) >= E(k1)  <=>  key >= (k1 & ~EMASK)
        const bool gt = key > (b.k1 | T5_EMASK), ge = key >= (b.k1 & ~T5_EMASK);
        b.k2 = gt ? 0u : b.k2;
        if (ge) b.k2 |= classbit;
        b.k1 = max(b.k1, key);
    }
}
// two candidates at once (KNN2): five instructions with the three-input maximum
__device__ __forceinline__ void consider5_pair(Best5& b, unsigned ka, unsigned kb)
{
    const unsigned lo = min(ka, kb), hi = max(ka, kb);
    const unsigned t = min(b.k1, hi);
    b.k1 = max(b.k1, hi);
    b.k2 = __vimax3_u32(b.k2, t, lo);
}

// merge of two top-2 states (KNN2) or two (best, class mask) states (COMPAT)
template <int MODE>
__device__ __forceinline__ void merge5(Best5& a, const Best5& o)
{
    if (MODE == AKZ_MATCH_KNN2) {
        const unsigned lo = min(a.k1, o.k1);
        a.k1 = max(a.k1, o.k1);
        a.k2 = max(max(a.k2, o.k2), lo);
    } else {
        const unsigned ea = a.k1 >> 19, eb = o.k1 >> 19;
        a.k2 = ea > eb ? a.k2 : (ea == eb ? (a.k2 | o.k2) : o.k2);
        a.k1 = max(a.k1, o.k1);
    }
}

__device__ __forceinline__ int popc128(const uint4& w) { return __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w); }

// Work decomposition.  An ITEM is (query block of 256, split of the train set): T tiles of 128 train descriptors whose index
// inside the item's range is c * H + ordinal (H = 2 T half tiles, see above).  The tiles of all items form one linear space
// (item-major); the grid is one CTA per SM and CTA g walks the tiles [g S, (g + 1) S): every SM gets the same number of tiles
// whatever the problem size, and a CTA pays its prologue once.  A CTA that crosses into the next item writes the partial
// result of the item it leaves (slot = its rank among the CTAs that cover that item), re-expands operand A and goes on; the
// expanders and the MMA issuer run ahead across the boundary.  k_match_merge folds the slots.
struct T5Args {
    const uint4* q; const uint4* t; akz_match_t* parts;
    int nq, nt, tbase;
    int T;                  // tiles per item (multiple of 8: the 64 columns of a half tile share their index class)
    int nsplit;             // items per query block
    int S;                  // tiles per CTA
    int total;              // tiles in the linear space
    int maxslots;           // partial results per item
};

template <int MODE, bool FILTER>
__global__ void __launch_bounds__(T5_NT, 1) k_match_tc5(const __grid_constant__ T5Args a)
{
    // k_match_merge is launched behind this kernel as a programmatic dependent: let its blocks become resident now (they block in
    // griddepcontrol.wait until this grid has finished and its writes are visible)
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ unsigned char t5raw[];
    unsigned char* sm = t5raw + ((1024u - (smem_u32(t5raw) & 1023u)) & 1023u);
    unsigned char* As = sm + T5_OFF_A;
    unsigned char* Ring = sm + T5_OFF_RING;
    const unsigned bar0 = smem_u32(sm + T5_OFF_BAR);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(sm + T5_OFF_TMEM);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (T5_SLOTS + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * T5_SLOTS + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * T5_SLOTS + T5_NACC + b); };
    const unsigned bar_afull = bar0 + 8u * (2 * T5_SLOTS + 2 * T5_NACC);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const uint4* __restrict__ q = a.q;
    const uint4* __restrict__ t = a.t;
    const int nq = a.nq, nt = a.nt, T = a.T, H = 2 * a.T;
    const int lin0 = blockIdx.x * a.S, lin1 = min(a.total, lin0 + a.S);

    // ---- prologue: barriers and tensor memory; the roles start right after one block barrier ---------------------------
    if (tid == 0) {
        for (int s = 0; s < T5_SLOTS; s++) { mbar_init(bar_full(s), T5_NEXP); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < T5_NACC; b++) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), T5_NEPI); }
        mbar_init(bar_afull, T5_NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == T5_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    if (wid < T5_MMA_WARP) {
        // ================================ epilogue: thread = query row ==========================================
        const int row = tid;                                         // A tile row >> 7, TMEM lane row & 127
        const unsigned taddr0 = tmem + ((unsigned)((wid & 3) * 32) << 16) + (unsigned)((wid >> 2) * T5_N);
        int gt = 0;                                                  // tiles this CTA has consumed (barrier phases)
        for (int lin = lin0; lin < lin1;) {
            const int item = lin / T, j0 = lin - item * T, j1 = min(T, j0 + (lin1 - lin));
            const int q0 = (item / a.nsplit) * T5_Q, t0 = (item % a.nsplit) * T * T5_N;
            // operand A of this item: every epilogue thread expands its own query.  The MMAs of the previous item are complete
            // (this thread has consumed the accumulator of its last tile), the expanders are already at work on the train tiles.
            int pq = 0;
            {
                uint4 x[4];
#pragma unroll
                for (int kb = 0; kb < 4; kb++) x[kb] = q0 + row < nq ? __ldg(q + 4 * (long long)(q0 + row) + kb) : make_uint4(0, 0, 0, 0);
                x[3].w &= 0x3Fu;
                pq = popc128(x[0]) + popc128(x[1]) + popc128(x[2]) + popc128(x[3]);
#pragma unroll
                for (int kb = 0; kb < 4; kb++) expand_store<false>(As + ((row >> 7) * 4 + kb) * T5_KB, row & 127, x[kb], kb == 3, 64u, 64u, 64u, 1u);
                fence_async_smem();
                mbar_arrive(bar_afull);
            }
            Best5 be, bo;                                            // two independent chains
            be.k1 = bo.k1 = 0u; be.k2 = bo.k2 = 0u;
            for (int j = j0; j < j1; j++, gt++) {
                const int b = gt % T5_NACC;
                mbar_wait(bar_tfull(b), (unsigned)(gt / T5_NACC) & 1u);
                tc_fence_after();
                const unsigned taddr = taddr0 + (unsigned)(b * T5_QT * T5_N);
                unsigned va[32], vb[32];
                tmem_ld32_pack16(taddr, va);                          // 64 columns = one half tile, two per register
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    unsigned (&v)[32] = c ? vb : va;
                    if (c == 0) tmem_ld32_pack16(taddr + 64, vb);      // in flight while the first half is processed
                    const int ordinal = 2 * j + c;
                    const unsigned ordinv = 8191u - (unsigned)ordinal;
                    // packed maximum of the 64 accumulators (even columns in the low halves, odd ones in the high halves)
                    unsigned m = 0;
                    if (FILTER || MODE != AKZ_MATCH_KNN2) {
                        m = __vimax3_u16x2(v[0], v[1], v[2]);
#pragma unroll
                        for (int i = 3; i + 1 < 32; i += 2) m = __vimax3_u16x2(m, v[i], v[i + 1]);
                        m = __vmaxu2(m, v[31]);
                        m = max(m & 0xFFFFu, m >> 16);
                    }
                    if (MODE != AKZ_MATCH_KNN2) {
                        // reference-compatible state = (best key, classes that attain the best distance).  H is a multiple of 16,
                        // so the 64 columns of a half tile share their index class: only the best key of the chunk can change
                        // the state, and one update per chunk is exact.
                        consider5<MODE>(be, m * 8192u + ordinv, 1u << ((unsigned)(a.tbase + ordinal) & 15u));
                    } else {
                        // a key is acc << 13 | ordinv, so (max << 13) | 8191 bounds the keys of the chunk
                        const bool update = !FILTER || ((m << 13) | 8191u) > max(be.k2, bo.k2);
                        if (update) {
                            // top-2 of the chunk on packed 16-bit pairs (even columns in the low halves, odd ones in the high
                            // halves; the accumulators of a chunk are distinct, they carry their column): 16 sorted pairs, then a
                            // merge tree of three packed instructions per node -- 1.2 ALU instructions per column instead of 4.5
                            // for unpacking every accumulator into a key.  Only the four winners become keys.
                            // (a per-register test against the current second best was tried first: 32 divergent branches cost
                            // more than the updates they skip.)
                            unsigned hi[16], lo[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) { hi[i] = __vmaxu2(v[2 * i], v[2 * i + 1]); lo[i] = __vminu2(v[2 * i], v[2 * i + 1]); }
#pragma unroll
                            for (int n = 16; n > 1; n >>= 1) {
#pragma unroll
                                for (int i = 0; i < n / 2; i++) {
                                    const unsigned a1 = hi[2 * i], a2 = lo[2 * i], b1 = hi[2 * i + 1], b2 = lo[2 * i + 1];
                                    hi[i] = __vmaxu2(a1, b1);
                                    lo[i] = __vimax3_u16x2(__vminu2(a1, b1), a2, b2);
                                }
                            }
                            consider5_pair(be, (hi[0] & 0xFFFFu) * 8192u + ordinv, (lo[0] & 0xFFFFu) * 8192u + ordinv);
                            consider5_pair(bo, (hi[0] >> 16) * 8192u + ordinv, (lo[0] >> 16) * 8192u + ordinv);
                        }
                    }
                    if (c == 0) tmem_ld_wait();
                }
                tc_fence_before();
                mbar_arrive(bar_tempty(b));
            }
            merge5<MODE>(be, bo);
            if (q0 + row < nq) {
                akz_match_t m;
                auto rel = [&](unsigned k) { return (int)(63u - ((k >> 13) & 63u)) * H + (int)(8191u - (k & 8191u)); };
                auto dist = [&](unsigned k) { return pq + 512 - (int)(k >> 19); };
                const bool has1 = be.k1 >= 8192u;
                m.idx1 = has1 ? a.tbase + t0 + rel(be.k1) : -1;
                m.dist1 = has1 ? dist(be.k1) : -1;
                if (MODE == AKZ_MATCH_KNN2) {
                    const bool has2 = be.k2 >= 8192u;
                    m.idx2 = has2 ? a.tbase + t0 + rel(be.k2) : -1;
                    m.dist2 = has2 ? dist(be.k2) : -1;
                } else {
                    m.idx2 = has1 ? (int)be.k2 : 0; m.dist2 = 0;
                }
                // slot = rank of this CTA among the CTAs that cover the item; the first of them also voids the unused slots
                const int gfirst = (item * T) / a.S, glast = ((item + 1) * T - 1) / a.S;
                akz_match_t* dst = a.parts + (long long)(item % a.nsplit) * a.maxslots * nq + q0 + row;
                dst[(long long)((int)blockIdx.x - gfirst) * nq] = m;
                if ((int)blockIdx.x == gfirst) {
                    akz_match_t none; none.idx1 = -1; none.dist1 = -1; none.idx2 = (MODE == AKZ_MATCH_KNN2) ? -1 : 0; none.dist2 = (MODE == AKZ_MATCH_KNN2) ? -1 : 0;
                    for (int sl = glast - gfirst + 1; sl < a.maxslots; sl++) dst[(long long)sl * nq] = none;
                }
            }
            lin += j1 - j0;
        }
    } else if (wid == T5_MMA_WARP) {
        // ================================ MMA issue: one warp, one elected lane per instruction =====================
        {
            const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const unsigned long long adesc0 = umma_desc(smem_u32(As));
            const unsigned long long bdesc0 = umma_desc(smem_u32(Ring));
            int gt = 0, seg = 0, cur_item = -1;
            int item = lin0 / T, j = lin0 - item * T;
            for (int lin = lin0; lin < lin1; lin++, gt++) {
                if (item != cur_item) {                                   // operand A of the new item
                    mbar_spin(bar_afull, (unsigned)seg & 1u);
                    cur_item = item; seg++;
                }
                const int b = gt % T5_NACC;
                mbar_spin(bar_tempty(b), ((unsigned)(gt / T5_NACC) & 1u) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 4; kb++) {
                    mbar_spin(bar_full(kb), (unsigned)gt & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < T5_QT; h++) {
                        const unsigned d_tmem = tmem_u + (unsigned)(b * T5_QT * T5_N + h * T5_N);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma_i8(d_tmem, adesc0 + (unsigned long long)((h * 4 + kb) * (T5_KB >> 4) + 2 * k),
                                    bdesc0 + (unsigned long long)(kb * (T5_KB >> 4) + 2 * k), (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(bar_empty(kb));
                }
                umma_commit(bar_tfull(b));
                if (++j == T) { j = 0; item++; }
            }
        }
        __syncwarp();
    } else {
        // ================================ expanders: thread = row of the train tile, all four K-blocks ===============
        // The serial instruction chain of an expander thread per tile (wait, expand, store, fence, arrive: four times) paces the
        // whole kernel (measured: a division per tile in this loop cost 13 %; moving the popcount behind its own loads 20 %), so
        // the loop is kept minimal: the next tile's row is prefetched, its position advances incrementally.
        const int row = tid - (T5_NEPI + 32);
        const int c64 = row & 63, half = row >> 6;
        const unsigned rowoff = (unsigned)c64 * (unsigned)H + (unsigned)half;      // index of the row inside the item's range, tile 0
        int nitem = lin0 / T, nj = lin0 - nitem * T;                               // position of the NEXT row to load
        long long ibase = (long long)(nitem % a.nsplit) * T * T5_N;
        auto load_next = [&](uint4 (&x)[4], bool live) {
            const long long g = ibase + rowoff + 2u * (unsigned)nj;               // the column is the major part of the index
            const bool ok = live && g < nt;
#pragma unroll
            for (int kb = 0; kb < 4; kb++) x[kb] = ok ? __ldg(t + 4 * g + kb) : make_uint4(0, 0, 0, 0);
            if (++nj == T) { nj = 0; nitem++; ibase = (long long)(nitem % a.nsplit) * T * T5_N; }
            return ok;
        };
        // two tiles of look-ahead: the rows of a tile lie H descriptors apart (up to 500 KB), one tile time does not cover
        // the latency of such a scattered read
        uint4 x[4], nx[4], nnx[4];
        bool ok = load_next(x, lin0 < lin1);
        bool nok = load_next(nx, lin0 + 1 < lin1);
        int gt = 0;
        for (int lin = lin0; lin < lin1; lin++, gt++) {
            const bool nnok = load_next(nnx, lin + 2 < lin1);
            // padding bytes of the row: three bytes summing to 512 - popc(t), and 63 - column
            x[3].w &= 0x3Fu;
            const int rest = 512 - (popc128(x[0]) + popc128(x[1]) + popc128(x[2]) + popc128(x[3]));
            const unsigned e6 = ok ? (unsigned)min(rest, 255) : 0u;
            const unsigned e7 = ok ? (unsigned)min(rest - (int)e6, 255) : 0u;
            const unsigned e0 = ok ? (unsigned)(rest - (int)e6 - (int)e7) : 0u;
            const unsigned e1 = ok ? (unsigned)(63 - c64) : 0u;
            const unsigned par = ((unsigned)gt & 1u) ^ 1u;
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                mbar_spin(bar_empty(kb), par);                                      // T5_SLOTS == 4: slot = K-block
                expand_store<true>(Ring + kb * T5_KB, row, x[kb], kb == 3, e6, e7, e0, e1);
                fence_async_smem();
                mbar_arrive(bar_full(kb));
            }
#pragma unroll
            for (int kb = 0; kb < 4; kb++) { x[kb] = nx[kb]; nx[kb] = nnx[kb]; }
            ok = nok; nok = nnok;
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

// tuning knob (environment AKZ_TC5_MAX_TILES), read once
static int tc5_max_tiles()
{
    static const int v = [] { const char* e = getenv("AKZ_TC5_MAX_TILES"); return e ? std::max(8, atoi(e) / 8 * 8) : 4096; }();
    return v;
}
// SM count of the current device (cached per device ordinal; a process may drive several GPUs)
static int tc5_sm_count()
{
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

// Tensor-memory matcher.  Writes plan->nparts partial results per query to `parts` (see the work decomposition above);
// match_tc5_plan gives the number of slots so that the caller can size the buffer first.
int match_tc5_plan(int nq, int nt, int* nsplit_out, int* tiles_out, int* per_cta_out, int* grid_out)
{
    const int nsm = tc5_sm_count();
    const int g_tc5_max_tiles = tc5_max_tiles();
    const int nqb = (nq + T5_Q - 1) / T5_Q;
    // items of at most g_tc5_max_tiles tiles: the rows of a tile are H = 2 T descriptors apart, and long ranges spread the 128 rows
    // of every tile over the whole train set
    const long long max_range = std::min<long long>(T5_MAX_RANGE, (long long)g_tc5_max_tiles * T5_N);
    int nsplit = (int)(((long long)nt + max_range - 1) / max_range);
    if (nsplit < 1) nsplit = 1;
    int T = (int)((((long long)nt + nsplit - 1) / nsplit + T5_RANGE_UNIT - 1) / T5_RANGE_UNIT) * (T5_RANGE_UNIT / T5_N);
    if (T < T5_RANGE_UNIT / T5_N) T = T5_RANGE_UNIT / T5_N;
    const long long total = (long long)nqb * nsplit * T;
    if (total >= (1ll << 30)) return akz_set_error(AKZ_E_UNSUPPORTED, "matching problem too large for one launch: shard the train set");
    const int grid = (int)std::min<long long>(nsm, total);
    const int S = (int)((total + grid - 1) / grid);
    const int maxslots = (T + S - 1) / S + 1;
    *nsplit_out = nsplit; *tiles_out = T; *per_cta_out = S; *grid_out = (int)((total + S - 1) / S);
    return nsplit * maxslots;
}

int match_partial_tc5(cudaStream_t st, const unsigned char* q, int nq, const unsigned char* t, int nt, int tbase, int mode, akz_match_t* parts, int filter_mode)
{
    if (nq <= 0) return 0;
    int nsplit, T, S, grid;
    const int nparts = match_tc5_plan(nq, nt, &nsplit, &T, &S, &grid);
    if (nparts < 0) return nparts;
    static akz_once_t attr;
    if (akz_once_guard once{attr}) {
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_KNN2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_COMPAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_KNN2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
        cudaFuncSetAttribute(k_match_tc5<AKZ_MATCH_COMPAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
    }
    T5Args a;
    a.q = (const uint4*)q; a.t = (const uint4*)t; a.parts = parts; a.nq = nq; a.nt = nt; a.tbase = tbase;
    a.T = T; a.nsplit = nsplit; a.S = S; a.total = ((nq + T5_Q - 1) / T5_Q) * nsplit * T; a.maxslots = nparts / nsplit;
    // a chunk rarely holds a new top-2 entry after ~2000 candidates; the reference-compatible state changes only on a new best
    // distance or a tie with it, which is rarer still: its filter pays off at every range length
    const bool filter = filter_mode < 0 ? (T * T5_N >= 2048 || mode == AKZ_MATCH_COMPAT) : filter_mode != 0;
    if (mode != AKZ_MATCH_COMPAT) {
        if (filter) k_match_tc5<AKZ_MATCH_KNN2, true><<<grid, T5_NT, T5_SMEM, st>>>(a);
        else k_match_tc5<AKZ_MATCH_KNN2, false><<<grid, T5_NT, T5_SMEM, st>>>(a);
    } else {
        if (filter) k_match_tc5<AKZ_MATCH_COMPAT, true><<<grid, T5_NT, T5_SMEM, st>>>(a);
        else k_match_tc5<AKZ_MATCH_COMPAT, false><<<grid, T5_NT, T5_SMEM, st>>>(a);
    }
    return 1;
}

}  // namespace akzk
