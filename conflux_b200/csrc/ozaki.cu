// conflux_b200/csrc/ozaki.cu -- the trailing update  C -= L * U  on the 5th-generation tensor cores (tcgen05, int8).
//
// Replaces cblas_dgemm at /root/reference/src/conflux/lu/conflux_opt.hpp:1628-1632 on the path selected by
// CFLX_GEMM=ozaki (gemm.cu's DMMA kernel stays the anchor and serves the TRSM sweeps).  tcgen05.mma has no f64 kind, so
// FP64 accuracy is recovered with error-free slicing (Ozaki scheme):
//   * every row m of L and every column n of U gets ONE power-of-two scale 2^ea[m] / 2^eb[n] (its largest magnitude);
//   * the scaled entries are cut into S = 8 signed digits (first 6 bits, then 7 bits each, round-to-nearest so every
//     digit fits [-64, 64]): x = 2^(e-6) * sum_s d_s 2^(-7s), exact to 55 bits relative to the row/column maximum;
//   * every digit-plane product  sum_k da_s[m][k] * db_t[k][n]  is an EXACT int8 x int8 -> int32 GEMM
//     (tcgen05.mma.kind::i8, accumulators in tensor memory); the planes with s + t = g share one accumulator (their
//     weight 2^(-7g) is the same; at most 8 products of <= 2^12 * K, K <= 512: < 2^25), only g <= 7 is kept (36 MMAs);
//   * the epilogue recombines  sum_g 2^(-7g) acc_g  in FP64, applies 2^(ea+eb-12) and subtracts from C in place.
// The neglected planes (g >= 8) are below 2^-56 of |row max| * |column max| per term: the same order as the rounding of
// a native FP64 dot product of that length (tools/ozaki_study.py: residual and pivots unchanged at 8 slices).
//
// Kernel structure (one CTA per SM, 640 threads, all 512 TMEM columns):
//   warp 0     TMA producer: digit planes are K-major int8 ([plane][row][K], 128-byte swizzled boxes of 128 k) fetched
//              with cp.async.bulk.tensor.3d through two tensor maps; per 128-k chunk the 8 B planes stay resident
//              (double-buffered set, ONE barrier pair per set) while the 8 A planes stream through a ring of three
//              two-plane slots: every plane chunk is loaded exactly once per tile;
//   warp 1     MMA issuer: one thread, UMMA 128 x (64..256) x 32 (up to four B planes per instruction), smem descriptors (SWIZZLE_128B, K-major), tcgen05.commit
//              onto the mbarriers that free operand slots and publish finished accumulators;
//   warp 2     TMEM allocation / release;
//   warps 4-19 epilogue: tcgen05.ld 16x256b (accumulator-fragment layout), software-pipelined int32 -> FP64 recombination in
//              registers; the negated update is staged in shared memory and ADDED to C by the L2 through bulk reduce
//              operations (cp.reduce.async.bulk .add.f64, SASS UBLKRED): the SM never reads C.
// The 8 accumulators are the pipeline between MMA and epilogue: the epilogue drains group g while the MMAs of the later
// groups still run, and the C update of a tile overlaps the MMAs of the next one.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace cflx {

namespace {
constexpr int OZ_S = 8;           // digit planes per operand
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_KC = 128;  // CTA tile, k-chunk (bytes = int8 elements)
constexpr int OZ_A_SLOTS = 2;      // each slot holds TWO consecutive A planes (one full/empty barrier pair per two rows)
constexpr int OZ_A_BYTES = OZ_BM * OZ_KC, OZ_B_BYTES = OZ_BN * OZ_KC;
constexpr int OZ_A_SLOT_BYTES = 2 * OZ_A_BYTES;
constexpr int OZ_EPI_WARPS = 16;  // 4 per TMEM lane quadrant, 16 columns each
constexpr int OZ_THREADS = 128 + 32 * OZ_EPI_WARPS;
constexpr size_t OZ_SMEM_OPERANDS = (size_t)2 * OZ_S * OZ_B_BYTES + (size_t)OZ_A_SLOTS * OZ_A_SLOT_BYTES;
constexpr int OZ_STAGE_DOUBLES = 16 * OZ_BN;   // per TMEM lane quadrant: 16 rows x 64 columns of (-update), source of the bulk reduce-add
constexpr size_t OZ_SMEM = 1024 /*alignment slack*/ + OZ_SMEM_OPERANDS + (size_t)4 * OZ_STAGE_DOUBLES * sizeof(double) + 512 /*barriers*/;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D(tmem) (+)= A(smem desc) * B(smem desc), int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors given as (low word, shared high word); ACC = true: D += A*B unconditionally
template <bool ACC>
__device__ __forceinline__ void umma_i8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    if constexpr (ACC) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.eq.u32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// 16 TMEM lanes x 16 columns in the accumulator-fragment layout (shape 16x256b, two 8-column repeats): thread T receives
//   r[4*c8 + 0..1] = lane T/4,     columns 8*c8 + 2*(T%4) + {0, 1}
//   r[4*c8 + 2..3] = lane T/4 + 8, same columns
// i.e. four lanes cover one 64-byte row segment of 8 int32 -- the mapping a coalesced C access wants.
__device__ __forceinline__ void tmem_ld_frag16(uint32_t taddr, int (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16]) {  // lane = thread, 16 consecutive columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor of a K-major, 128-byte-swizzled operand tile (rows of 128 bytes, 8-row groups 1024 B
// apart): start address >> 4 | SBO = 1024 >> 4 at bit 32 | version 1 at bit 46 | SWIZZLE_128B (2) at bit 61
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::i8: D = S32 (2 at bit 4), A = B = signed int8 (1 at bits 7 / 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t oz_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
}

// C[...] += smem[...] (FP64, round-to-nearest) done by the L2 through the TMA engine (SASS UBLKRED.ADD.F64): no C load, no
// load latency on the SM; bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_reduce_add_f64(double* gdst, const double* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void l2_prefetch_128(const void* p) {  // 16-byte aligned, 128 bytes
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], 128;" ::"l"(p) : "memory");
}
__host__ __device__ constexpr double pow2c(int e) {  // compile-time 2^e, e <= 0
    double x = 1.0;
    for (int i = 0; i < -e; ++i) x *= 0.5;
    return x;
}
__device__ __forceinline__ double pow2i(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }  // |e| < 1022

struct OzArgs {
    int M, N, K;           // C is M x N, K = contraction length (multiple of 128)
    int a_row0;            // first row of the A planes that belongs to row 0 of this C window
    int b_row0;            // first row of the B planes that belongs to column 0 of this C window
    double* C;             // in place: C -= A^T-planes * B-planes
    int64_t ldc;
    const int* ea;         // [M]   row exponents of A
    const int* eb;         // [...] column exponents of B, indexed like the B planes (b_row0 + n)
    int tiles_m, tiles_n;
    long long* dbg;        // optional [16] cycle counters of CTA 0: see cflx_dbg_ozaki_cycles
};

template <bool DBG>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, OzArgs g) {
    extern __shared__ unsigned char oz_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)oz_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = base;                                        // [2][8][64 x 128 B]
    unsigned char* sA = base + (size_t)2 * OZ_S * OZ_B_BYTES;         // [3][128 x 128 B]
    double* stage = reinterpret_cast<double*>(base + OZ_SMEM_OPERANDS);   // [4 lane quadrants][16 rows][64 columns]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * OZ_STAGE_DOUBLES);
    uint64_t* fullB = bars;                 // [2]  one per B set (8 planes land on it)
    uint64_t* emptyB = bars + 2;            // [2]
    uint64_t* fullA = bars + 4;             // [2..3]  one per A slot (2 planes)
    uint64_t* emptyA = bars + 7;            // [2..3]
    uint64_t* tfull = bars + 10;            // [8]  accumulator (group) g is complete
    uint64_t* tdone = bars + 18;            // [1]  the epilogue has drained all 8 accumulators of the tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = g.tiles_m * g.tiles_n;
    const int KC = g.K / OZ_KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&fullB[i], 1);
            mbar_init(&emptyB[i], 1);
        }
        for (int i = 0; i < OZ_A_SLOTS; ++i) {
            mbar_init(&fullA[i], 1);
            mbar_init(&emptyA[i], 1);
        }
        for (int i = 0; i < OZ_S; ++i) mbar_init(&tfull[i], 1);
        mbar_init(tdone, OZ_EPI_WARPS);  // one arrival per epilogue warp
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    long long tm[4] = {0, 0, 0, 0};
    long long tc0 = 0;
#define OZ_T0() if constexpr (DBG) tc0 = clock64();
#define OZ_T1(i) if constexpr (DBG) tm[i] += clock64() - tc0;
    const long long t_begin = DBG ? clock64() : 0;
    if (warp == 0) {
        // ================================================================== TMA producer
        uint32_t q = 0, acnt = 0;  // k-chunks and A slots (plane pairs) loaded so far by this CTA
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / g.tiles_n) * OZ_BM, n0 = (tile % g.tiles_n) * OZ_BN;
            for (int kc = 0; kc < KC; ++kc, ++q) {
                const int set = q & 1;
                const uint32_t useB = q >> 1;
                OZ_T0() mbar_wait(&emptyB[set], (useB & 1) ^ 1); OZ_T1(0)
                if (lane == 0) {
                    mbar_arrive_expect_tx(&fullB[set], OZ_S * OZ_B_BYTES);
                    for (int t = 0; t < OZ_S; ++t)
                        tma_load_3d(sB + (size_t)(set * 8 + t) * OZ_B_BYTES, &mapB, kc * OZ_KC, g.b_row0 + n0, t, &fullB[set]);
                }
                for (int sp = 0; sp < OZ_S / 2; ++sp, ++acnt) {
                    const int slot = acnt % OZ_A_SLOTS;
                    const uint32_t useA = acnt / OZ_A_SLOTS;
                    OZ_T0() mbar_wait(&emptyA[slot], (useA & 1) ^ 1); OZ_T1(1)
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&fullA[slot], OZ_A_SLOT_BYTES);
                        tma_load_3d(sA + (size_t)slot * OZ_A_SLOT_BYTES, &mapA, kc * OZ_KC, g.a_row0 + m0, 2 * sp, &fullA[slot]);
                        tma_load_3d(sA + (size_t)slot * OZ_A_SLOT_BYTES + OZ_A_BYTES, &mapA, kc * OZ_KC, g.a_row0 + m0, 2 * sp + 1, &fullA[slot]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        // One CTA per SM owns all 512 TMEM columns, so the allocation starts at column 0 and every accumulator address is
        // a compile-time constant; the plane loop is fully unrolled so that each UMMA is issued from immediates + two
        // descriptor low words (the issue path of this single thread is what the tensor pipe waits for).
        if (tmem_base != 0) __trap();
        const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
        constexpr uint32_t DESC_HI = (uint32_t)((((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61)) >> 32);
        uint32_t q = 0, acnt = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            for (int kc = 0; kc < KC; ++kc, ++q) {
                const int set = q & 1;
                const uint32_t useB = q >> 1;
                const uint32_t b_lo = ((sB0 + (uint32_t)set * 8u * OZ_B_BYTES) & 0x3FFFF) >> 4;
                OZ_T0() mbar_wait(&fullB[set], useB & 1); OZ_T1(1)
                OZ_T0() if (kc == 0) mbar_wait(tdone, (it & 1) ^ 1); OZ_T1(2)   // the epilogue drained the previous tile
#pragma unroll
                for (int sp = 0; sp < OZ_S / 2; ++sp, ++acnt) {
                    const int slot = acnt % OZ_A_SLOTS;
                    OZ_T0() mbar_wait(&fullA[slot], (acnt / OZ_A_SLOTS) & 1); OZ_T1(0)
                    tc_fence_after();
                    if (lane == 0) {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int s = 2 * sp + half;
                            const uint32_t a_lo = ((sA0 + (uint32_t)slot * OZ_A_SLOT_BYTES + (uint32_t)half * OZ_A_BYTES) & 0x3FFFF) >> 4;
                            // The B planes t, t+1, ... of a set are contiguous 64-row tiles and their accumulators (groups
                            // s+t, s+t+1, ...) are contiguous 64-column blocks of tensor memory: up to four planes go through
                            // ONE UMMA of N = 64 * planes, so A_s is read from shared memory once per four products.
#pragma unroll
                            for (int t = 0; t < OZ_S - s; t += 4) {
                                const int np = (OZ_S - s - t) < 4 ? (OZ_S - s - t) : 4;
                                const uint32_t bt_lo = b_lo + (uint32_t)t * (OZ_B_BYTES >> 4);
                                const uint32_t d = (uint32_t)(s + t) * OZ_BN;
#pragma unroll
                                for (int k = 0; k < OZ_KC / 32; ++k) {  // +32 bytes along K inside the swizzle atom: +2 in the descriptor
                                    if (s == 0 && k == 0) umma_i8_lo<false>(d, a_lo, bt_lo, DESC_HI, oz_idesc(np * OZ_BN), kc > 0 ? 1u : 0u);
                                    else umma_i8_lo<true>(d, a_lo + 2 * k, bt_lo + 2 * k, DESC_HI, oz_idesc(np * OZ_BN), 1u);
                                }
                            }
                            if (kc == KC - 1) tc_commit(&tfull[s]);       // groups <= s are complete
                        }
                        tc_commit(&emptyA[slot]);                         // planes 2sp, 2sp+1 of A are consumed
                        if (sp == OZ_S / 2 - 1) tc_commit(&emptyB[set]);  // the whole B set of this k chunk is consumed
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ================================================================== epilogue (16 warps)
        // warp -> TMEM lane quadrant lq (32 rows) x 16-column quarter cq.  The accumulators are read in the fragment layout
        // (tmem_ld_frag16): thread T holds rows {T/4, T/4+8, T/4+16, T/4+24} x column pairs {2(T%4), 8+2(T%4)}, so four
        // lanes cover a 64-byte row segment and the C read-modify-write needs no shared-memory transpose.
        const int ew = warp - 4;
        const int lq = warp & 3;
        const int cq = ew >> 2;
        const int cp = lane & 3, r8 = lane >> 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int m0 = (tile / g.tiles_n) * OZ_BM, n0 = (tile % g.tiles_n) * OZ_BN;
            {   // pull my row's 128-byte segment of the C tile into L2 now; it is needed after the 8 accumulators
                const int prow = m0 + lq * 32 + lane, pcol = n0 + cq * 16;
                if (prow < g.M && pcol + 16 <= g.N) l2_prefetch_128(g.C + (int64_t)prow * g.ldc + pcol);
            }
            const int rowbase = m0 + lq * 32 + r8;             // my rows: rowbase + 8 * j, j = 0..3
            const int colbase = n0 + cq * 16 + 2 * cp;          // my column pairs: colbase + 8 * c8
            // exponents of my 4 rows and 4 columns (kept as ints: half the registers; the powers of two are rebuilt at use)
            int erow[4], ecol[2][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) erow[j] = (rowbase + 8 * j < g.M) ? g.ea[g.a_row0 + rowbase + 8 * j] - 12 : 0;
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
                const int col = colbase + 8 * c8;
                ecol[c8][0] = col < g.N ? g.eb[g.b_row0 + col] : 0;
                ecol[c8][1] = col < g.N ? g.eb[g.b_row0 + col + 1] : 0;
            }
            // sum[h][i]: lane half h (rows 16h..16h+15 of the quadrant), i = 4*c8 + 2*rsel + x as delivered by tmem_ld_frag16
            double sum[2][8];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 8; ++i) sum[h][i] = 0.0;
            // 16 half-steps (group g, lane half h), two 8-register buffers: the TMEM load of the NEXT half-step is issued
            // right after the current one has landed and BEFORE its FP64 work, so load latency and recombination overlap.
            int acc[2][8];
            const uint32_t ta0 = ((uint32_t)(lq * 32) << 16) + (uint32_t)(cq * 16);
            OZ_T0() mbar_wait(&tfull[0], it & 1); OZ_T1(0)
            tc_fence_after();
            tmem_ld_frag16(ta0, acc[0]);
#pragma unroll
            for (int step = 0; step < 2 * OZ_S; ++step) {
                const int grp = step >> 1, h = step & 1;
                OZ_T0()
                tmem_ld_wait();                                  // acc[step & 1] has landed
                bool issued = true;
                if (step + 1 < 2 * OZ_S) {
                    const uint32_t tn = ta0 + (uint32_t)(((step + 1) >> 1) * OZ_BN) + ((uint32_t)(((step + 1) & 1) * 16) << 16);
                    if (h == 0) {
                        tmem_ld_frag16(tn, acc[(step + 1) & 1]);  // other lane half of the same group
                    } else if (mbar_test(&tfull[grp + 1], it & 1)) {
                        tc_fence_after();
                        tmem_ld_frag16(tn, acc[(step + 1) & 1]);  // next group is already complete
                    } else {
                        issued = false;
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tdone);            // all 8 accumulators are in registers
                }
                const double w = pow2c(-7 * grp);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // exact int32 -> double without the conversion pipe: 2^52 + 2^31 + x, then subtract the bias
                    const double x = __hiloint2double(0x43300000, acc[step & 1][i] ^ 0x80000000) - 4503601774854144.0;
                    sum[h][i] = fma(x, w, sum[h][i]);
                }
                OZ_T1(1)
                if (!issued) {
                    OZ_T0() mbar_wait(&tfull[grp + 1], it & 1); OZ_T1(0)
                    tc_fence_after();
                    tmem_ld_frag16(ta0 + (uint32_t)((grp + 1) * OZ_BN), acc[(step + 1) & 1]);
                }
            }
            OZ_T0()
            // ---- C -= sum * 2^(ea - 12) * 2^eb: the NEGATED update is staged in shared memory (16 rows x 16 columns per lane
            // half) and added to C by the L2 through a bulk reduce (one 128-byte row segment per lane): the SM never loads C
            // (the rows were pulled into L2 by the prefetch at the top of the tile).
            // The four warps of a lane quadrant share one 16-row x 64-column staging tile, so that ONE bulk reduce moves a whole
            // 512-byte row segment (the TMA engine's cost is per operation: 128 operations per tile instead of 512).
            double* stg = stage + (size_t)lq * OZ_STAGE_DOUBLES;
            const int segcols = min(OZ_BN, g.N - n0);            // valid columns of the tile (even)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                bulk_wait_read();                                 // my previous reduces have read their staging rows
                named_bar_sync(1 + lq, 128);                      // ... and so have those of the other three warps
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel) {
                    const double sr = -pow2i(erow[2 * h + rsel]);
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        double2 o;
                        o.x = (sum[h][4 * c8 + 2 * rsel] * sr) * pow2i(ecol[c8][0]);
                        o.y = (sum[h][4 * c8 + 2 * rsel + 1] * sr) * pow2i(ecol[c8][1]);
                        *reinterpret_cast<double2*>(stg + (8 * rsel + r8) * OZ_BN + cq * 16 + 8 * c8 + 2 * cp) = o;
                    }
                }
                fence_proxy_async();
                named_bar_sync(1 + lq, 128);                      // the whole 16 x 64 tile is staged
                if (lane < 4) {                                   // warp cq issues rows 4cq .. 4cq+3 of this half
                    const int lr = 4 * cq + lane;
                    const int r = m0 + lq * 32 + 16 * h + lr;
                    if (r < g.M && segcols > 0)
                        bulk_reduce_add_f64(g.C + (int64_t)r * g.ldc + n0, stg + lr * OZ_BN, (uint32_t)segcols * 8u);
                    bulk_commit();
                }
            }
            OZ_T1(2)
        }
    }
    if (DBG && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
        const int o = warp == 0 ? 0 : (warp == 1 ? 4 : 8);
        for (int i = 0; i < 3; ++i) g.dbg[o + i] = tm[i];
        g.dbg[o + 3] = clock64() - t_begin;
    }
#undef OZ_T0
#undef OZ_T1
    if (warp >= 4) bulk_wait_all();   // every reduce-add of this thread has been performed
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------- raw UMMA rate probe
// Back-to-back tcgen05.mma on one resident shared-memory tile per SM (no TMA, no epilogue): the tensor pipe's own rate
// for kind::i8 (and kind::f16 for calibration against the bf16 peak in MEASURED_PEAKS.json).
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int n, int use_f16, int iters) {
    extern __shared__ unsigned char pk_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)pk_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (threadIdx.x == 0) {
        const uint64_t adesc = smem_desc_sw128(smem_u32(base)), bdesc = smem_desc_sw128(smem_u32(base + 16384));
        // f16 kind: D = F32 (1 at bit 4), A = B = BF16 (1 at bits 7 / 10); i8 kind as in the GEMM
        const uint32_t idesc = use_f16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24)) : oz_idesc(n);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (use_f16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (uint32_t)((it & 1) * 256)),
                                 "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u)
                                 : "memory");
                else
                    umma_i8(tb + (uint32_t)((it & 1) * 256), adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// ---------------------------------------------------------------------------------------------- digit planes
// src[k][o] (K x ld doubles, o = row of L / column of U) -> planes[s][o0 + o][k] int8 (K contiguous), exps[o0 + o].
// One CTA per 32 outer indices: pass 1 = largest magnitude per outer index, pass 2 = digits, transposed through smem.
constexpr int OZ_SPLIT_OUT = 32;
__global__ void __launch_bounds__(256) ozaki_split_kernel(const double* __restrict__ src, int64_t ld, int n_outer, int K,
                                                          int8_t* __restrict__ planes, int64_t plane_stride, int o0,
                                                          int* __restrict__ exps) {
    constexpr int PITCH = OZ_KC + 4;  // bytes per smem row: 33 words, conflict-free byte scatter
    __shared__ __align__(16) unsigned char tile[OZ_S][OZ_SPLIT_OUT][PITCH];
    __shared__ double red[8][OZ_SPLIT_OUT];
    __shared__ int s_exp[OZ_SPLIT_OUT];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int o = blockIdx.x * OZ_SPLIT_OUT + tx;
    const bool valid = o < n_outer;
    double mx = 0.0;
    if (valid)
        for (int k = ty; k < K; k += 8) mx = fmax(mx, fabs(src[(int64_t)k * ld + o]));
    red[ty][tx] = mx;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmax(mx, red[i][tx]);
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1): |x| * 2^-e < 1 for the whole row
        s_exp[tx] = e;
        if (valid) exps[o0 + o] = e;
    }
    __syncthreads();
    const double inv = pow2i(-s_exp[tx]);
    for (int k0 = 0; k0 < K; k0 += OZ_KC) {
        for (int kk = ty; kk < OZ_KC; kk += 8) {
            double r = valid ? src[(int64_t)(k0 + kk) * ld + o] * inv : 0.0;
            r *= 64.0;  // first digit: 6 bits
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) {
                const double d = rint(r);
                tile[s][tx][kk] = (unsigned char)(signed char)(int)d;
                r = (r - d) * 128.0;
            }
        }
        __syncthreads();
        // write-out: plane s, outer row, 128 contiguous bytes = 8 x 16 B
        for (int e = threadIdx.x; e < OZ_S * OZ_SPLIT_OUT * 8; e += 256) {
            const int s = e / (OZ_SPLIT_OUT * 8), rowi = (e / 8) % OZ_SPLIT_OUT, c16 = e % 8;
            const int oo = blockIdx.x * OZ_SPLIT_OUT + rowi;
            if (oo < n_outer) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&tile[s][rowi][c16 * 16]);
                uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(planes + (int64_t)s * plane_stride + (int64_t)(o0 + oo) * K + k0 + c16 * 16) = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p) fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}
// planes: [OZ_S][cap_rows][K] int8; box = 128 bytes of k x box_rows rows of one plane, 128-byte swizzle
int make_plane_map(CUtensorMap* map, const int8_t* planes, int K, int cap_rows, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled is not available from the driver");
        return CFLX_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)cap_rows, (cuuint64_t)OZ_S};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)K * (cuuint64_t)cap_rows};
    cuuint32_t box[3] = {(cuuint32_t)OZ_KC, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d) for K=%d rows=%d", (int)r, K, cap_rows);
        return CFLX_ERR_CUDA;
    }
    return CFLX_OK;
}
}  // namespace

struct OzakiWorkspace::Maps {
    CUtensorMap a, b;
};

int ozaki_workspace_create(OzakiWorkspace* ws, int max_rows, int max_cols, int K) {
    *ws = OzakiWorkspace{};
    if (K <= 0 || K % OZ_KC != 0 || K > 512) {
        set_last_error("ozaki: contraction length %d unsupported (multiple of 128, <= 512)", K);
        return CFLX_ERR_UNSUPPORTED;
    }
    ws->K = K;
    ws->cap_a = (int)round_up(max_rows, OZ_BM);
    ws->cap_b = (int)round_up(max_cols, OZ_BN);
    CFLX_CUDA(cudaMalloc((void**)&ws->planesA, (size_t)OZ_S * ws->cap_a * K));
    CFLX_CUDA(cudaMalloc((void**)&ws->planesB, (size_t)OZ_S * ws->cap_b * K));
    CFLX_CUDA(cudaMemset(ws->planesA, 0, (size_t)OZ_S * ws->cap_a * K));
    CFLX_CUDA(cudaMemset(ws->planesB, 0, (size_t)OZ_S * ws->cap_b * K));
    CFLX_CUDA(cudaMalloc((void**)&ws->ea, sizeof(int) * ws->cap_a));
    CFLX_CUDA(cudaMalloc((void**)&ws->eb, sizeof(int) * ws->cap_b));
    CFLX_CUDA(cudaMemset(ws->ea, 0, sizeof(int) * ws->cap_a));
    CFLX_CUDA(cudaMemset(ws->eb, 0, sizeof(int) * ws->cap_b));
    if (getenv("CFLX_OZAKI_DBG")) {
        CFLX_CUDA(cudaMalloc((void**)&ws->dbg, sizeof(long long) * 16));
        CFLX_CUDA(cudaMemset(ws->dbg, 0, sizeof(long long) * 16));
    }
    ws->maps = new OzakiWorkspace::Maps;
    CFLX_TRY(make_plane_map(&ws->maps->a, ws->planesA, K, ws->cap_a, OZ_BM));
    CFLX_TRY(make_plane_map(&ws->maps->b, ws->planesB, K, ws->cap_b, OZ_BN));
    static PerDeviceMax cfg;
    if (cfg.raise(OZ_SMEM)) {
        CFLX_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
        CFLX_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
    }
    int dev = 0;
    CFLX_CUDA(cudaGetDevice(&dev));
    CFLX_CUDA(cudaDeviceGetAttribute(&ws->sms, cudaDevAttrMultiProcessorCount, dev));
    return CFLX_OK;
}
void ozaki_workspace_destroy(OzakiWorkspace* ws) {
    cudaFree(ws->planesA);
    cudaFree(ws->planesB);
    cudaFree(ws->ea);
    cudaFree(ws->eb);
    cudaFree(ws->dbg);
    delete ws->maps;
    *ws = OzakiWorkspace{};
}

// Tera-MACs/s of back-to-back UMMA 128 x n x (32 bytes of K) instructions, one CTA per SM: use_f16 = 0 -> kind::i8
// (32 int8 per instruction), 1 -> kind::f16 on bf16 (16 elements per instruction).
int umma_peak_probe(int n, int use_f16, double* tmacs_out) {
    if (n < 16 || n > 256 || n % 16) return CFLX_ERR_ARG;
    int dev = 0, sms = 0;
    CFLX_CUDA(cudaGetDevice(&dev));
    CFLX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = 1024 + 16384 + 32768;
    CFLX_CUDA(cudaFuncSetAttribute(umma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CFLX_CUDA(cudaEventCreate(&e0));
    CFLX_CUDA(cudaEventCreate(&e1));
    const int iters = 20000;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CFLX_CUDA(cudaEventRecord(e0));
        umma_peak_kernel<<<sms, 128, smem>>>(n, use_f16, iters);
        CFLX_CUDA(cudaEventRecord(e1));
        CFLX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CFLX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double kper = use_f16 ? 16.0 : 32.0;
        const double macs = (double)sms * iters * 4 * 128.0 * n * kper;
        best = std::max(best, macs / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CFLX_CUDA(cudaGetLastError());
    *tmacs_out = best;
    return CFLX_OK;
}

// digit planes of rows [0, n) of L^T (LT[k][row], ld) -> A planes
int ozaki_split_a(OzakiWorkspace* ws, const double* LT, int64_t ld, int n, cudaStream_t s) {
    if (n <= 0) return CFLX_OK;
    if (n > ws->cap_a) return CFLX_ERR_ARG;
    ozaki_split_kernel<<<(n + OZ_SPLIT_OUT - 1) / OZ_SPLIT_OUT, 256, 0, s>>>(LT, ld, n, ws->K, ws->planesA, (int64_t)ws->cap_a * ws->K, 0, ws->ea);
    CFLX_CUDA(cudaGetLastError());
    return CFLX_OK;
}
// digit planes of columns [col0, col0 + n) of U (U[k][col], ld) -> B planes, rows col0..
int ozaki_split_b(OzakiWorkspace* ws, const double* U, int64_t ld, int col0, int n, cudaStream_t s) {
    if (n <= 0) return CFLX_OK;
    if (col0 < 0 || col0 + n > ws->cap_b) return CFLX_ERR_ARG;
    ozaki_split_kernel<<<(n + OZ_SPLIT_OUT - 1) / OZ_SPLIT_OUT, 256, 0, s>>>(U + col0, ld, n, ws->K, ws->planesB, (int64_t)ws->cap_b * ws->K, col0, ws->eb);
    CFLX_CUDA(cudaGetLastError());
    return CFLX_OK;
}
// C[0..M) x [0..N) -= (rows row0..row0+M of the A planes) * (rows col0..col0+N of the B planes)
int launch_ozaki_gemm(OzakiWorkspace* ws, int M, int N, int row0, int col0, double* C, int64_t ldc, int max_ctas, cudaStream_t s) {
    if (M <= 0 || N <= 0) return CFLX_OK;
    if ((N & 1) || (ldc & 1) || col0 < 0 || row0 < 0 || row0 + M > ws->cap_a || col0 + N > ws->cap_b) {
        set_last_error("ozaki_gemm: unsupported window M=%d N=%d row0=%d col0=%d ldc=%lld", M, N, row0, col0, (long long)ldc);
        return CFLX_ERR_UNSUPPORTED;
    }
    OzArgs g{};
    g.M = M; g.N = N; g.K = ws->K;
    g.a_row0 = row0;
    g.b_row0 = col0;
    g.C = C; g.ldc = ldc;
    g.ea = ws->ea; g.eb = ws->eb;
    g.tiles_m = (M + OZ_BM - 1) / OZ_BM;
    g.tiles_n = (N + OZ_BN - 1) / OZ_BN;
    g.dbg = ws->dbg;
    int grid = g.tiles_m * g.tiles_n;
    int cap = ws->sms;
    if (max_ctas > 0 && max_ctas < cap) cap = max_ctas;
    if (grid > cap) grid = cap;
    if (ws->dbg) ozaki_gemm_kernel<true><<<grid, OZ_THREADS, OZ_SMEM, s>>>(ws->maps->a, ws->maps->b, g);
    else ozaki_gemm_kernel<false><<<grid, OZ_THREADS, OZ_SMEM, s>>>(ws->maps->a, ws->maps->b, g);
    CFLX_CUDA(cudaGetLastError());
    return CFLX_OK;
}

}  // namespace cflx
