// K4-TC: the dense transforms on the 5th-generation tensor cores (tcgen05), float32-accurate.
//
//   C[M,N] (+)= sum_s A_s[M,K_s] * W_s[N,K_s]^T (+ bias)        A_s, W_s row-major, K contiguous
//
// The parity bound of the hot path is rel 1e-5 in float32; a single TF32 product is 1e-3.  Each
// operand is therefore split in shared memory into hi = x with the low 13 mantissa bits cleared
// (exactly a TF32 value) and lo = x - hi (exact in float32), and three tcgen05.mma kind::tf32
// products  hi*hi + lo*hi + hi*lo  accumulate in TMEM in float32 ("3xTF32"): error ~2^-21.
// The kernel stays HBM-bound (AI of [N,128]x[128,128] is 32 flop/B, a third of the TF32 tensor
// peak still exceeds what HBM can feed).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) of A / W k-blocks [128 x 32 f32]
//   warp 1      MMA issuer: one elected lane issues the tcgen05.mma's, commits to mbarriers
//   warps 2-5   splitters: hi/lo split of the landed tiles (generic proxy -> fence.proxy.async)
//   warps 6-9   epilogue: tcgen05.ld of the accumulator (double-buffered in TMEM), bias /
//               accumulate, 128-bit global stores
// Pipelines: smem ring (TMA -> split -> MMA, kStages deep), TMEM ring (MMA -> epilogue, 2 deep).
#include <cuda.h>

#include "agx_common.cuh"

namespace agx {

constexpr int kTcBM = 128;                 // rows per tile
constexpr int kTcBK = 32;                  // float32 per k-block = one 128 B swizzle span
constexpr int kTcStages = 3;
constexpr int kTcMaxSegs = 8;
constexpr int kTcMaxBN = 128;               // output columns per tile (UMMA N), TMEM: 2 x 128 columns
constexpr int kTcThreads = 320;            // 10 warps
constexpr int kTileBytes = kTcBM * kTcBK * 4;   // 16 KB

struct TcSeg {
    CUtensorMap map_a;                     // [M, K] f32, box [32, 128], swizzle 128B
    CUtensorMap map_w;                     // [N, K] f32, box [32, BN]
    int32_t k_blocks;
    int32_t pad_[15];
};

struct TcProb {
    CUtensorMap map_c;                     // [M, N] f32, box [32, 128], swizzle 128B (TMA store)
    const float* bias;
    int32_t seg_begin, n_seg;
    int32_t M, N, BN;                      // BN = N rounded up to 16 (UMMA N), <= kTcMaxBN
    int32_t accumulate;
    int32_t tile_begin;                    // first global tile index of this problem
    int32_t total_kb;                      // k-blocks per tile (all segments)
    int32_t pad_[6];
};

constexpr int kTcMaxProbs = 24;
constexpr int kTcMaxSegsTotal = 40;

struct TcParams {
    TcSeg seg[kTcMaxSegsTotal];
    TcProb prob[kTcMaxProbs];
    int32_t n_prob;
    int32_t total_tiles;
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x,
                                            int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Round-to-nearest onto the TF32 grid (10 explicit mantissa bits), low 13 bits cleared so that the
// tensor core's own treatment of those bits cannot matter.  x = hi + lo exactly with
// |lo| <= 2^-12 |x|; rounding lo as well leaves 2^-24 |x|: float32-level products from 3 MMAs.
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// K-major operand tile, 128 B rows, 128B swizzle, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address      [0,14)
    d |= (uint64_t)1 << 16;                           // LBO (unused)       [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO = 1024 B       [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version [46,48)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B       [61,64)
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format  : F32
    d |= 2u << 7;                    // A format  : TF32
    d |= 2u << 10;                   // B format  : TF32
    d |= (uint32_t)(N >> 3) << 17;   // N
    d |= (uint32_t)(M >> 4) << 24;   // M
    return d;                        // A, B K-major, no negate, dense
}

struct __align__(1024) TcSmem {
    float a_hi[kTcStages][kTcBM * kTcBK];
    float a_lo[kTcStages][kTcBM * kTcBK];
    float w_hi[kTcStages][kTcMaxBN * kTcBK];
    float w_lo[kTcStages][kTcMaxBN * kTcBK];
    float stage[2][kTcBM * 32];             // epilogue staging: 128 rows x 32 columns, 128B swizzle
    uint64_t full[kTcStages];               // TMA landed
    uint64_t split[kTcStages];              // hi/lo written
    uint64_t empty[kTcStages];              // MMAs of the stage retired
    uint64_t acc_full[2];                   // accumulator complete
    uint64_t acc_empty[2];                  // accumulator drained by the epilogue
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_tc(const __grid_constant__ TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.split[s], 128);
            mbar_init(&S.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&S.acc_full[a], 1);
            mbar_init(&S.acc_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 2 accumulators x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
            smem_u32(&S.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    // every role walks the same static tile sequence: tile -> (problem, row tile)
#define AGX_TC_FOR_TILES(...)                                                           \
    for (int gt = blockIdx.x, pi = 0; gt < P.total_tiles; gt += gridDim.x) {            \
        while (pi + 1 < P.n_prob && gt >= P.prob[pi + 1].tile_begin) ++pi;              \
        const TcProb& Q = P.prob[pi];                                                   \
        const int tile = gt - Q.tile_begin;                                             \
        __VA_ARGS__                                                                     \
    }

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            AGX_TC_FOR_TILES({
                const uint32_t w_tile_bytes = (uint32_t)Q.BN * kTcBK * 4;
                for (int s = Q.seg_begin; s < Q.seg_begin + Q.n_seg; ++s) {
                    for (int kb = 0; kb < P.seg[s].k_blocks; ++kb, ++it) {
                        const int st = it % kTcStages;
                        const uint32_t ph = (it / kTcStages) & 1;
                        mbar_wait(&S.empty[st], ph ^ 1);
                        mbar_expect_tx(&S.full[st], kTileBytes + w_tile_bytes);
                        tma_load_2d(&P.seg[s].map_a, &S.full[st], S.a_hi[st], kb * kTcBK,
                                    tile * kTcBM);
                        tma_load_2d(&P.seg[s].map_w, &S.full[st], S.w_hi[st], kb * kTcBK, 0);
                    }
                }
            })
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t it = 0, tl = 0;
        AGX_TC_FOR_TILES({
            (void)tile;
            const uint32_t idesc = make_idesc(kTcBM, Q.BN);
            const int total_kb = Q.total_kb;
            const int acc = tl & 1;
            const uint32_t acc_ph = (tl >> 1) & 1;
            mbar_wait(&S.acc_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem + (uint32_t)acc * kTcMaxBN;
            for (int kbt = 0; kbt < total_kb; ++kbt, ++it) {
                const int st = it % kTcStages;
                const uint32_t ph = (it / kTcStages) & 1;
                mbar_wait(&S.split[st], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(S.a_hi[st]), a_lo = smem_u32(S.a_lo[st]);
                    const uint32_t w_hi = smem_u32(S.w_hi[st]), w_lo = smem_u32(S.w_lo[st]);
#pragma unroll
                    for (int k = 0; k < kTcBK / 8; ++k) {       // 8 tf32 = 32 B per MMA
                        const uint32_t off = k * 32;
                        const uint32_t first = (kbt == 0 && k == 0) ? 0u : 1u;
                        tc_mma_tf32(d_tmem, make_desc(a_lo + off), make_desc(w_hi + off), idesc, first);
                        tc_mma_tf32(d_tmem, make_desc(a_hi + off), make_desc(w_lo + off), idesc, 1u);
                        tc_mma_tf32(d_tmem, make_desc(a_hi + off), make_desc(w_hi + off), idesc, 1u);
                    }
                    tc_commit(&S.empty[st]);                     // frees the smem stage
                    if (kbt == total_kb - 1) tc_commit(&S.acc_full[acc]);
                }
                __syncwarp();
            }
            ++tl;
        })
    } else if (warp < 6) {
        // ================= splitters (128 threads) =================
        const int t = threadIdx.x - 64;
        uint32_t it = 0;
        AGX_TC_FOR_TILES({
            (void)tile;
            const int BN = Q.BN;
            for (int kbt = 0; kbt < Q.total_kb; ++kbt, ++it) {
                const int st = it % kTcStages;
                const uint32_t ph = (it / kTcStages) & 1;
                mbar_wait(&S.full[st], ph);
                // elementwise, layout-agnostic: the swizzled byte order is kept as it landed
                float4* ah = reinterpret_cast<float4*>(S.a_hi[st]);
                float4* al = reinterpret_cast<float4*>(S.a_lo[st]);
#pragma unroll 4
                for (int i = t; i < kTcBM * kTcBK / 4; i += 128) {
                    const float4 x = ah[i];
                    float4 h, l;
                    h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
                    h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
                    h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
                    h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
                    ah[i] = h;
                    al[i] = l;
                }
                float4* wh = reinterpret_cast<float4*>(S.w_hi[st]);
                float4* wl = reinterpret_cast<float4*>(S.w_lo[st]);
                for (int i = t; i < BN * kTcBK / 4; i += 128) {
                    const float4 x = wh[i];
                    float4 h, l;
                    h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
                    h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
                    h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
                    h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
                    wh[i] = h;
                    wl[i] = l;
                }
                fence_proxy_async();             // generic-proxy writes -> visible to tcgen05.mma
                mbar_arrive(&S.split[st]);
            }
        })
    } else {
        // ================= epilogue (warps 6..9; TMEM lane quarter = warp % 4) =================
        // accumulator row r of the tile lives in TMEM lane r: thread (q, lane) owns row 32q+lane.
        // Each 32-column slab goes registers -> swizzled smem -> one TMA store (or TMA reduce-add
        // when accumulating): full 128 B lines leave the SM instead of 32 scattered 16 B pieces.
        const int q = warp & 3;
        const int r_in_tile = q * 32 + lane;
        const bool store_thread = (warp == 6 && lane == 0);
        uint32_t tl = 0, chunk_ctr = 0;
        AGX_TC_FOR_TILES({
            const int BN = Q.BN;
            const int acc = tl & 1;
            const uint32_t acc_ph = (tl >> 1) & 1;
            mbar_wait(&S.acc_full[acc], acc_ph);
            tc_fence_after();
            for (int c0 = 0; c0 < BN; c0 += 32, ++chunk_ctr) {
                float* stg = S.stage[chunk_ctr & 1];
                // the bulk store that last read this staging buffer must have drained it
                if (store_thread) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                uint32_t v[32];
                const uint32_t taddr = tmem + (uint32_t)acc * kTcMaxBN + (uint32_t)c0 + ((uint32_t)(q * 32) << 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
                    "[%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                      "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                      "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                      "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                      "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 32 >= BN) {             // last slab read: hand the accumulator back
                    tc_fence_before();
                    mbar_arrive(&S.acc_empty[acc]);
                }
                // row r_in_tile, 16-byte chunk j -> physical chunk j ^ (row % 8)  (128B swizzle)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    if (Q.bias) {
                        const int c = c0 + 4 * j;
                        o.x += c + 0 < Q.N ? Q.bias[c + 0] : 0.f;
                        o.y += c + 1 < Q.N ? Q.bias[c + 1] : 0.f;
                        o.z += c + 2 < Q.N ? Q.bias[c + 2] : 0.f;
                        o.w += c + 3 < Q.N ? Q.bias[c + 3] : 0.f;
                    }
                    *reinterpret_cast<float4*>(stg + r_in_tile * 32 + ((j ^ (r_in_tile & 7)) << 2)) = o;
                }
                fence_proxy_async();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (store_thread) {
                    if (Q.accumulate)
                        asm volatile(
                            "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group "
                            "[%0, {%2, %3}], [%1];" ::"l"(&Q.map_c),
                            "r"(smem_u32(stg)), "r"(c0), "r"(tile * kTcBM)
                            : "memory");
                    else
                        asm volatile(
                            "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group "
                            "[%0, {%2, %3}], [%1];" ::"l"(&Q.map_c),
                            "r"(smem_u32(stg)), "r"(c0), "r"(tile * kTcBM)
                            : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            ++tl;
        })
        if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
    }
}

// =================================================================================================
// Long reductions on the tensor cores: C[M,N] = A^T B with A [K, M], B [K, N] row-major and K the
// (very long) node dimension -- the weight gradients dW = dOut^T X over 10^5 rows.
//
//   * operands are MN-major for the MMA (the contiguous dimension is M resp. N, not K): TMA boxes
//     of [32 k-rows x 32 floats] land in the 128B-swizzle-with-32B-atoms layout (the only one the
//     tensor core takes for MN-major TF32), MN blocks 4 KB apart (LBO), 4-row k groups 512 B apart
//     (SBO); the instruction descriptor carries a_major = b_major = MN;
//   * split-K over CTAs: each CTA reduces kLkBlocksPerCta k-blocks and writes one [M, N] partial
//     (combined in split order by gemm_splitk_reduce: deterministic);
//   * the tensor core's float32 accumulator truncates (error ~1.2e-8 per accumulated K element),
//     so an accumulator is only trusted for kLkFlush k-blocks (256 rows): the epilogue warps drain
//     it (double-buffered in TMEM) into round-to-nearest float32 running sums held in registers.
// =================================================================================================
constexpr int kLkStages = 3;
constexpr int kLkFlush = 8;                 // k-blocks (of 32 rows) per TMEM accumulator
constexpr int kLkBlocksPerCta = 24;         // 768 rows of K per CTA
constexpr int kLkMaxProbs = 8;
constexpr int kLkBlockFloats = 32 * 32;     // one TMA box: 32 k-rows x 32 floats = 4 KB

struct LkProb {
    CUtensorMap map_a;                      // [K, M] f32, box [32, 32], swizzle 128B
    CUtensorMap map_b;                      // [K, N] f32, box [32, 32]
    float* partial;                         // [splits][M][N]
    int32_t M, N, BN;
    int32_t total_kb, item_begin, n_items, per;     // per = k-blocks per CTA
    int32_t pad_[7];
};

struct LkParams {
    LkProb prob[kLkMaxProbs];
    int32_t n_prob;
    int32_t total_items;
};

struct __align__(1024) LkSmem {
    float a_hi[kLkStages][4 * kLkBlockFloats];
    float a_lo[kLkStages][4 * kLkBlockFloats];
    float b_hi[kLkStages][4 * kLkBlockFloats];
    float b_lo[kLkStages][4 * kLkBlockFloats];
    uint64_t full[kLkStages];
    uint64_t split[kLkStages];
    uint64_t empty[kLkStages];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

// MN-major TF32 operand: the only shared-memory layout the tensor core accepts is the 128B swizzle
// with 32 B atomicity (SWIZZLE_128B_BASE32B; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- atoms of
// 4 k-rows x 128 B.  LBO = 4096 B between 32-float MN blocks (one TMA box each), SBO = 512 B
// between 4-row k groups.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                           // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_tc_longk(const __grid_constant__ LkParams P) {
    extern __shared__ uint8_t smem_raw[];
    LkSmem& S = *reinterpret_cast<LkSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int pi = 0;
    while (pi + 1 < P.n_prob && (int)blockIdx.x >= P.prob[pi + 1].item_begin) ++pi;
    const LkProb& Q = P.prob[pi];
    const int sp = blockIdx.x - Q.item_begin;
    const int kb0 = sp * Q.per;
    const int nkb = min(Q.total_kb, kb0 + Q.per) - kb0;      // >= 1 by construction
    const int nA = (Q.M + 31) / 32, nB = (Q.N + 31) / 32;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kLkStages; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.split[s], 128);
            mbar_init(&S.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&S.acc_full[a], 1);
            mbar_init(&S.acc_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
            smem_u32(&S.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 2 && warp < 6 && (nA < 4 || nB < 4)) {
        // MN blocks beyond M / N are never loaded: give the MMA zeros there, not stale shared memory
        const int t = threadIdx.x - 64;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int st = 0; st < kLkStages; ++st) {
            for (int i = nA * 256 + t; i < 4 * 256; i += 128) {
                reinterpret_cast<float4*>(S.a_hi[st])[i] = z;
                reinterpret_cast<float4*>(S.a_lo[st])[i] = z;
            }
            for (int i = nB * 256 + t; i < 4 * 256; i += 128) {
                reinterpret_cast<float4*>(S.b_hi[st])[i] = z;
                reinterpret_cast<float4*>(S.b_lo[st])[i] = z;
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int st = it % kLkStages;
                const uint32_t ph = (it / kLkStages) & 1;
                mbar_wait(&S.empty[st], ph ^ 1);
                mbar_expect_tx(&S.full[st], (uint32_t)(nA + nB) * kLkBlockFloats * 4);
                const int krow = (kb0 + it) * 32;
                for (int b = 0; b < nA; ++b)
                    tma_load_2d(&Q.map_a, &S.full[st], S.a_hi[st] + b * kLkBlockFloats, b * 32, krow);
                for (int b = 0; b < nB; ++b)
                    tma_load_2d(&Q.map_b, &S.full[st], S.b_hi[st] + b * kLkBlockFloats, b * 32, krow);
            }
        }
    } else if (warp == 1) {
        // M = 128 rows of D, N = BN columns; A and B MN-major (bits 15, 16)
        const uint32_t idesc = make_idesc(kTcBM, Q.BN) | (1u << 15) | (1u << 16);
        int it = 0;
        for (int g = 0; it < nkb; ++g) {
            const int acc = g & 1;
            const uint32_t acc_ph = (g >> 1) & 1;
            mbar_wait(&S.acc_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem + (uint32_t)acc * kTcMaxBN;
            const int gend = min(nkb, it + kLkFlush);
            for (int j = 0; it < gend; ++it, ++j) {
                const int st = it % kLkStages;
                const uint32_t ph = (it / kLkStages) & 1;
                mbar_wait(&S.split[st], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(S.a_hi[st]), a_lo = smem_u32(S.a_lo[st]);
                    const uint32_t b_hi = smem_u32(S.b_hi[st]), b_lo = smem_u32(S.b_lo[st]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {               // 8 k-rows (one 1 KB atom) per MMA
                        const uint32_t off = k * 1024;
                        const uint32_t first = (j == 0 && k == 0) ? 0u : 1u;
                        tc_mma_tf32(d_tmem, make_desc_mn(a_lo + off), make_desc_mn(b_hi + off), idesc, first);
                        tc_mma_tf32(d_tmem, make_desc_mn(a_hi + off), make_desc_mn(b_lo + off), idesc, 1u);
                        tc_mma_tf32(d_tmem, make_desc_mn(a_hi + off), make_desc_mn(b_hi + off), idesc, 1u);
                    }
                    tc_commit(&S.empty[st]);
                    if (it == gend - 1) tc_commit(&S.acc_full[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        const int t = threadIdx.x - 64;
        for (int it = 0; it < nkb; ++it) {
            const int st = it % kLkStages;
            const uint32_t ph = (it / kLkStages) & 1;
            mbar_wait(&S.full[st], ph);
            float4* ah = reinterpret_cast<float4*>(S.a_hi[st]);
            float4* al = reinterpret_cast<float4*>(S.a_lo[st]);
#pragma unroll 4
            for (int i = t; i < nA * 256; i += 128) {
                const float4 x = ah[i];
                float4 h, l;
                h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
                h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
                h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
                h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
                ah[i] = h;
                al[i] = l;
            }
            float4* bh = reinterpret_cast<float4*>(S.b_hi[st]);
            float4* bl = reinterpret_cast<float4*>(S.b_lo[st]);
#pragma unroll 4
            for (int i = t; i < nB * 256; i += 128) {
                const float4 x = bh[i];
                float4 h, l;
                h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
                h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
                h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
                h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
                bh[i] = h;
                bl[i] = l;
            }
            fence_proxy_async();
            mbar_arrive(&S.split[st]);
        }
    } else {
        // epilogue: thread (q, lane) owns accumulator row 32q + lane; running sums in registers
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int BN = Q.BN;
        float run[kTcMaxBN];
#pragma unroll
        for (int i = 0; i < kTcMaxBN; ++i) run[i] = 0.f;
        const int groups = (nkb + kLkFlush - 1) / kLkFlush;
        for (int g = 0; g < groups; ++g) {
            const int acc = g & 1;
            const uint32_t acc_ph = (g >> 1) & 1;
            mbar_wait(&S.acc_full[acc], acc_ph);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kTcMaxBN; c0 += 32) {
                if (c0 < BN) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem + (uint32_t)acc * kTcMaxBN + (uint32_t)c0 + ((uint32_t)(q * 32) << 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
                        "[%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                          "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                          "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) run[c0 + j] += __uint_as_float(v[j]);
                }
            }
            tc_fence_before();
            mbar_arrive(&S.acc_empty[acc]);
        }
        if (row < Q.M) {
            float* out = Q.partial + ((size_t)sp * Q.M + row) * Q.N;
#pragma unroll
            for (int c = 0; c < kTcMaxBN; c += 4) {
                if (c + 3 < Q.N && (Q.N & 3) == 0) {
                    *reinterpret_cast<float4*>(out + c) = make_float4(run[c], run[c + 1], run[c + 2], run[c + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (c + i < Q.N) out[c + i] = run[c + i];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
    }
}

// ---- host ----------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows, int box_cols = kTcBK,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return AGX_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                  (long long)rows, (long long)cols, (long long)ld);
        return AGX_ERR_CUDA;
    }
    return AGX_OK;
}

static bool tc_seg_ok(const agx_gemm_seg_t& S) {
    // A[M,K] K-contiguous, B[k][n] = W[n][k] K-contiguous, 16-byte aligned rows, no masks
    return S.a_cs == 1 && S.b_rs == 1 && !S.A_mask && !S.B_mask && S.K >= 8 && (S.a_rs % 4) == 0 &&
           (S.b_cs % 4) == 0 && (reinterpret_cast<uintptr_t>(S.A) % 16) == 0 &&
           (reinterpret_cast<uintptr_t>(S.B) % 16) == 0;
}

bool gemm_tc_eligible(const agx_gemm_problem_t& Q, const agx_gemm_seg_t* segs) {
    if (Q.split_k > 1 || Q.row_scale || Q.skip_flag || Q.seg_count < 1 || Q.seg_count > kTcMaxSegs)
        return false;
    // Measured on B200 (profiles/probes/tc_err_probe.py): the tensor core's float32 accumulator truncates,
    // the error grows ~1.2e-8 per accumulated K element (1.5e-6 at K=128, 1.2e-5 at K=1024).
    // Only reductions up to 256 stay an order of magnitude inside the 1e-5 parity bound; longer
    // ones (the heads' K = 896 / 2176) and the tiny node types (whose 18-row BatchNorm amplifies
    // input error ~100x) take the exact FFMA kernel.
    if (Q.N < 8 || Q.N > kTcMaxBN || Q.M < 512) return false;
    int64_t ktot = 0;
    for (int s = 0; s < Q.seg_count; ++s) ktot += segs[Q.seg_begin + s].K;
    if (ktot > 256) return false;
    if ((reinterpret_cast<uintptr_t>(Q.C) % 16) != 0 || (Q.ldc % 4) != 0) return false;
    for (int s = 0; s < Q.seg_count; ++s)
        if (!tc_seg_ok(segs[Q.seg_begin + s])) return false;
    return true;
}

// All eligible problems of one agx_gemm_grouped call go into ONE persistent launch.
int gemm_tc_launch(const agx_gemm_problem_t* probs, const int* idx, int cnt,
                   const agx_gemm_seg_t* segs, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(TcSmem) + 1024;
    if (!attr_set) {
        AGX_CUDA(cudaFuncSetAttribute(gemm_tf32x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    int done = 0;
    while (done < cnt) {
        TcParams P;
        P.n_prob = 0;
        P.total_tiles = 0;
        int nseg = 0;
        while (done < cnt && P.n_prob < kTcMaxProbs &&
               nseg + probs[idx[done]].seg_count <= kTcMaxSegsTotal) {
            const agx_gemm_problem_t& Q = probs[idx[done]];
            TcProb& T = P.prob[P.n_prob];
            T.bias = Q.bias;
            T.seg_begin = nseg;
            T.n_seg = Q.seg_count;
            T.M = Q.M;
            T.N = Q.N;
            T.BN = (Q.N + 15) / 16 * 16;
            T.accumulate = Q.accumulate;
            T.tile_begin = P.total_tiles;
            T.total_kb = 0;
            int rc = make_map(&T.map_c, Q.C, Q.M, Q.N, Q.ldc, kTcBM);
            if (rc) return rc;
            for (int s = 0; s < Q.seg_count; ++s) {
                const agx_gemm_seg_t& G = segs[Q.seg_begin + s];
                TcSeg& Sg = P.seg[nseg++];
                rc = make_map(&Sg.map_a, G.A, Q.M, G.K, G.a_rs, kTcBM);
                if (rc) return rc;
                rc = make_map(&Sg.map_w, G.B, Q.N, G.K, G.b_cs, T.BN);
                if (rc) return rc;
                Sg.k_blocks = (G.K + kTcBK - 1) / kTcBK;
                T.total_kb += Sg.k_blocks;
            }
            P.total_tiles += (Q.M + kTcBM - 1) / kTcBM;
            ++P.n_prob;
            ++done;
        }
        if (P.n_prob == 0) {
            set_error("gemm_tc_launch: a problem does not fit the segment table");
            return AGX_ERR_INVALID;
        }
        const int grid = P.total_tiles < kNumSMs ? P.total_tiles : kNumSMs;
        gemm_tf32x3_tc<<<grid, kTcThreads, smem, st>>>(P);
        AGX_LAUNCH_CHECK("gemm_tf32x3_tc");
    }
    return AGX_OK;
}

// ---- long-K (weight gradient) problems --------------------------------------------------------
bool gemm_tc_longk_eligible(const agx_gemm_problem_t& Q, const agx_gemm_seg_t* segs) {
    if (Q.split_k <= 1 || Q.seg_count != 1 || Q.skip_flag || !Q.partial) return false;
    if (Q.M < 1 || Q.M > kTcBM || Q.N < 8 || Q.N > kTcMaxBN) return false;
    const agx_gemm_seg_t& S = segs[Q.seg_begin];
    // A^T given as opA[M,K] with element (m,k) at A[m + k*a_cs]; B [K,N] row-major
    if (S.K < 2048 || S.a_rs != 1 || S.b_cs != 1 || S.A_mask || S.B_mask) return false;
    if ((S.a_cs % 4) != 0 || (S.b_rs % 4) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(S.A) % 16) != 0 || (reinterpret_cast<uintptr_t>(S.B) % 16) != 0)
        return false;
    // the caller sized the partial buffer for split_k slabs: enough for the splits used here?
    // (the launch gives a CTA at least total_kb / 140 + 1 k-blocks -- one wave -- so a problem
    // never takes more than ~140 slabs however long K is: 1.9 M rows on the 16x graph)
    const int total_kb = (S.K + 31) / 32;
    const int per = total_kb / 140 + 1 > kLkBlocksPerCta ? total_kb / 140 + 1 : kLkBlocksPerCta;
    const int items = (total_kb + per - 1) / per;
    return items <= Q.split_k && (reinterpret_cast<uintptr_t>(Q.partial) % 16) == 0;
}

// used_split[i] receives the number of partial slabs written for problem idx[i]
int gemm_tc_longk_launch(const agx_gemm_problem_t* probs, const int* idx, int cnt,
                         const agx_gemm_seg_t* segs, int* used_split, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(LkSmem) + 1024;
    if (!attr_set) {
        AGX_CUDA(cudaFuncSetAttribute(gemm_tf32x3_tc_longk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    // one wave: aim at <= ~144 CTAs over all problems of the launch (at least 24 k-blocks each)
    int64_t all_kb = 0;
    for (int i = 0; i < cnt; ++i) all_kb += (segs[probs[idx[i]].seg_begin].K + 31) / 32;
    const int kb_per_cta = (int)(all_kb / 140 + 1) > kLkBlocksPerCta ? (int)(all_kb / 140 + 1)
                                                                     : kLkBlocksPerCta;
    int done = 0;
    while (done < cnt) {
        LkParams P;
        P.n_prob = 0;
        P.total_items = 0;
        while (done < cnt && P.n_prob < kLkMaxProbs) {
            const agx_gemm_problem_t& Q = probs[idx[done]];
            const agx_gemm_seg_t& G = segs[Q.seg_begin];
            LkProb& T = P.prob[P.n_prob];
            T.partial = Q.partial;
            T.M = Q.M;
            T.N = Q.N;
            T.BN = (Q.N + 15) / 16 * 16;
            T.total_kb = (G.K + 31) / 32;
            T.n_items = (T.total_kb + kb_per_cta - 1) / kb_per_cta;
            // every CTA must own at least one k-block with the ceil split used by the kernel
            const int per = (T.total_kb + T.n_items - 1) / T.n_items;
            T.n_items = (T.total_kb + per - 1) / per;
            T.per = per;
            T.item_begin = P.total_items;
            int rc = make_map(&T.map_a, G.A, G.K, Q.M, G.a_cs, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
            if (rc) return rc;
            rc = make_map(&T.map_b, G.B, G.K, Q.N, G.b_rs, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
            if (rc) return rc;
            used_split[done] = T.n_items;
            P.total_items += T.n_items;
            ++P.n_prob;
            ++done;
        }
        gemm_tf32x3_tc_longk<<<P.total_items, kTcThreads, smem, st>>>(P);
        AGX_LAUNCH_CHECK("gemm_tf32x3_tc_longk");
    }
    return AGX_OK;
}

}  // namespace agx
