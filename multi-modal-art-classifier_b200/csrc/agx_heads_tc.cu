// K5/K6 on the tensor cores: one fused kernel per mini-batch for the fusion / projector heads
// (agx_head_step, include/agx.h).
//
//   reference, per batch (src/models/models_kg.py:237-243, src/train_new_multimodal_multitask.py:76-83,
//   under fp16 autocast):   comb = cat(feat, emb) ; out = Linear(Dropout(comb)) ; CE ; backward
//
// Here, per 128-row tile and head, ONE CTA (1 CTA per SM, persistent over its tiles):
//
//   pass 1  for every 64-column block of the virtual concatenation [feat | emb]:
//             TMA (cp.async.bulk.tensor) stages the float32 block [128 x 64] in shared memory,
//             8 converter warps apply the dropout mask (Philox4x32-10 from (key, step, row, column)
//             generated here, or an explicit mask), round to bf16 and write the K-major 128B-swizzled
//             operand tile + the matching [64 x 64] bf16 weight block; one elected thread issues
//             4 tcgen05.mma.kind::f16 (M=128 rows, N=64 classes, K=16) into the logits accumulator
//             in TMEM.
//   epilogue  4 warps read the logits with tcgen05.ld (thread = row): + bias, softmax cross entropy
//             (class weights, global normaliser) or SmoothL1, write the logits, and leave the logit
//             gradient as a bf16 [128 rows x 64 classes] tile in shared memory.
//   pass 2  the same input blocks again (L2 hits), two per step: the bf16 tile [row][column] that was
//             the K-major A operand of pass 1 IS the MN-major A operand of
//                 d_weight^T[128 columns, 64 classes] += x^T[128 columns, 128 rows] . dlogits[128 rows, 64]
//             (no transposition in shared memory; the logit-gradient tile is the MN-major B operand);
//             up to 7 such accumulators live in TMEM (64 + 7 * 64 = 512 columns) across all tiles of
//             the CTA and are drained once at the end into this CTA's partial.
//   agx_head_reduce then adds the partials of the CTAs in fixed order (deterministic).
//
// Heads wider than 64 outputs are passed as several heads over column slices (the projector's 128
// outputs = 2 heads); reductions longer than 7 * 128 columns (ResNet features, K = 2176) are split
// over CTAs that repeat the forward pass for their column range.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "agx_common.cuh"

namespace agx {

constexpr int kHdRows = 128;               // rows per tile = UMMA M of the forward
constexpr int kHdBK = 64;                  // columns per block (64 bf16 = one 128 B swizzle span)
constexpr int kHdCP = 64;                  // classes, padded = UMMA N
constexpr int kHdXkStages = 4;             // bf16 operand ring (pass 2 consumes aligned PAIRS)
constexpr int kHdXfStages = 3;             // float32 TMA staging ring
constexpr int kHdMaxChunks = 7;            // 128-column d_weight accumulators per CTA
constexpr int kHdConvThreads = 256;
constexpr int kHdThreads = 32 * 14;        // warp 0 MMA, warp 1 TMA, warps 2-9 convert, 10-13 epilogue
constexpr int kHdXfBytes = kHdRows * kHdBK * 4;       // 32 KB
constexpr int kHdXkBytes = kHdRows * kHdBK * 2;       // 16 KB
constexpr int kHdWkBytes = kHdCP * kHdBK * 2;         //  8 KB
constexpr int kHdBitmapBytes = 2 * kHdMaxChunks * kHdConvThreads * 4;   // 14 KB

struct HdHeadDev {
    CUtensorMap map[2];                    // [B, width] float32, box [64 columns, 128 rows], no swizzle
    const float* mask;
    int64_t ld_mask;
    const __nv_bfloat16* wbf;              // [64][K] bf16, rows >= C are zero
    const float* bias;
    const int64_t* labels;
    const float* class_w;
    const float* norm;                     // device scalar: sum of class_w[labels] over the GLOBAL batch
    const float* target;
    int64_t ld_target;
    float* logits;
    int64_t ld_logits;
    float* dw_part;                        // [G][C][K]
    float* db_part;                        // [G][64]
    float* loss_part;                      // [G]
    float coef, inv_count;
    int32_t width0, C, loss, pad_;
};

struct HdParams {
    HdHeadDev h[AGX_MAX_HEADS];
    const int64_t* seed_state;
    int32_t n_heads, B, K, n_tiles, G, n_split, chunks_per_split, use_philox;
    uint32_t thresh;
    float keep_scale;
};

struct HdSmem {
    float xf[kHdXfStages][kHdRows * kHdBK];
    __nv_bfloat16 xk[kHdXkStages][kHdRows * kHdBK];
    __nv_bfloat16 wk[kHdXkStages][kHdCP * kHdBK];
    __nv_bfloat16 dl[kHdRows * kHdCP];
    uint32_t bitmap[2 * kHdMaxChunks][kHdConvThreads];
    uint64_t xfull[kHdXfStages];           // TMA landed (tx bytes)
    uint64_t xfree[kHdXfStages];           // 256 converters have read the float32 stage
    uint64_t conv[kHdXkStages];            // 256 converters have written the bf16 stage
    uint64_t empty[kHdXkStages];           // MMAs that read the bf16 stage have retired
    uint64_t logits_full;                  // logits accumulator complete
    uint64_t dl_ready;                     // 128 epilogue threads: logits drained, dlogits tile written
    uint64_t dw_full;                      // all d_weight accumulators complete
    float s_db[4][kHdCP];
    float s_loss[4];
    uint32_t tmem_base;
};

// ---- PTX helpers (same conventions as agx_gemm_tc.cu) ------------------------------------------
__device__ __forceinline__ uint32_t hd_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hd_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hd_smem(bar)), "r"(count));
}
__device__ __forceinline__ void hd_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hd_smem(bar)) : "memory");
}
__device__ __forceinline__ void hd_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hd_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hd_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(hd_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void hd_tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(hd_smem(dst)), "l"(map), "r"(hd_smem(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void hd_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void hd_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hd_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hd_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hd_smem(bar)) : "memory");
}
__device__ __forceinline__ void hd_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// shared-memory matrix descriptor, 128B swizzle, 8-row groups 1024 B apart (SBO); LBO = distance of
// the next 64-element atom along M/N (MN-major operands only)
__device__ __forceinline__ uint64_t hd_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t hd_idesc(int M, int N, bool mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D: F32
    d |= 1u << 7;                    // A: BF16
    d |= 1u << 10;                   // B: BF16
    if (mn_major) d |= (1u << 15) | (1u << 16);
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void hd_tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void hd_philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ uint32_t hd_pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kHdThreads, 1)
head_fused_bf16(const __grid_constant__ HdParams P) {
    extern __shared__ uint8_t hd_raw[];
    HdSmem& S = *reinterpret_cast<HdSmem*>((reinterpret_cast<uintptr_t>(hd_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // CTA -> (head, column split, slot): the CTA owns tiles slot, slot + G, ...
    const int unit = (int)blockIdx.x / P.G, slot = (int)blockIdx.x % P.G;
    const int hi = unit / P.n_split, split = unit % P.n_split;
    const HdHeadDev& H = P.h[hi];
    const int K = P.K;
    const int nb1 = K / kHdBK;                                   // blocks of pass 1
    const int c0 = split * P.chunks_per_split;                   // this CTA's d_weight chunks
    const int c1 = min(K / 128, c0 + P.chunks_per_split);
    const int nch = c1 - c0;
    const int my_tiles = slot < P.n_tiles ? (P.n_tiles - slot + P.G - 1) / P.G : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kHdXfStages; ++s) {
            hd_mbar_init(&S.xfull[s], 1);
            hd_mbar_init(&S.xfree[s], kHdConvThreads);
        }
        for (int s = 0; s < kHdXkStages; ++s) {
            hd_mbar_init(&S.conv[s], kHdConvThreads);
            hd_mbar_init(&S.empty[s], 1);
        }
        hd_mbar_init(&S.logits_full, 1);
        hd_mbar_init(&S.dl_ready, 128);
        hd_mbar_init(&S.dw_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(hd_smem(&S.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    hd_tc_before();
    __syncthreads();
    hd_tc_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 1) {
        // ================= TMA producer: float32 blocks, pass 1 then pass 2 of every tile ==========
        if (lane == 0 && my_tiles > 0) {
            uint32_t fi = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int row0 = (slot + t * P.G) * kHdRows;
                const int total = nb1 + 2 * nch;
                for (int i = 0; i < total; ++i, ++fi) {
                    const int kb = i < nb1 ? i : 2 * c0 + (i - nb1);
                    const int col = kb * kHdBK;
                    const int fs = fi % kHdXfStages;
                    hd_mbar_wait(&S.xfree[fs], ((fi / kHdXfStages) & 1) ^ 1);
                    hd_mbar_expect_tx(&S.xfull[fs], kHdXfBytes);
                    if (col < H.width0)
                        hd_tma_load_2d(&H.map[0], &S.xfull[fs], S.xf[fs], col, row0);
                    else
                        hd_tma_load_2d(&H.map[1], &S.xfull[fs], S.xf[fs], col - H.width0, row0);
                }
            }
        }
    } else if (warp == 0) {
        // ================= MMA issuer ===============================================================
        if (my_tiles > 0) {
            const uint32_t idesc_f = hd_idesc(kHdRows, kHdCP, false);
            const uint32_t idesc_b = hd_idesc(128, kHdCP, true);
            uint32_t it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                // pass 1: logits[128, 64] = x[128, K] w[64, K]^T.  The previous tile's epilogue has
                // drained the accumulator (dl_ready was awaited before that tile's pass 2)
                for (int kb = 0; kb < nb1; ++kb, ++it) {
                    const int st = it % kHdXkStages;
                    hd_mbar_wait(&S.conv[st], (it / kHdXkStages) & 1);
                    hd_tc_after();
                    if (lane == 0) {
                        const uint32_t a = hd_smem(S.xk[st]), b = hd_smem(S.wk[st]);
#pragma unroll
                        for (int k = 0; k < kHdBK / 16; ++k)
                            hd_mma_bf16(tmem, hd_desc(a + k * 32, 16), hd_desc(b + k * 32, 16), idesc_f,
                                        (kb == 0 && k == 0) ? 0u : 1u);
                        hd_commit(&S.empty[st]);
                        if (kb == nb1 - 1) hd_commit(&S.logits_full);
                    }
                    __syncwarp();
                }
                // pass 2: d_weight^T chunk [128 columns, 64] += x^T . dlogits over the tile's 128 rows
                hd_mbar_wait(&S.dl_ready, t & 1);
                hd_tc_after();
                for (int j = 0; j < nch; ++j, it += 2) {
                    const int st = it % kHdXkStages;             // even: (st, st + 1) is a pair
                    hd_mbar_wait(&S.conv[st], (it / kHdXkStages) & 1);
                    hd_mbar_wait(&S.conv[st + 1], (it / kHdXkStages) & 1);
                    hd_tc_after();
                    if (lane == 0) {
                        const uint32_t a = hd_smem(S.xk[st]), b = hd_smem(S.dl);
                        const uint32_t d = tmem + kHdCP + (uint32_t)j * kHdCP;
#pragma unroll
                        for (int kk = 0; kk < kHdRows / 16; ++kk)       // 16 rows of K per MMA
                            hd_mma_bf16(d, hd_desc(a + kk * 2048, kHdXkBytes), hd_desc(b + kk * 2048, 16),
                                        idesc_b, (t == 0 && kk == 0) ? 0u : 1u);
                        hd_commit(&S.empty[st]);
                        hd_commit(&S.empty[st + 1]);
                        if (t == my_tiles - 1 && j == nch - 1) hd_commit(&S.dw_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 10) {
        // ================= converters (256 threads) =================================================
        // unit u = tid + 256 i (i < 8) of a block: row u / 16, float4 column u % 16
        const int tid = threadIdx.x - 64;
        const int c4 = tid & 15, rbase = tid >> 4;
        uint32_t key0 = 0, key1 = 0, step = 0;
        if (P.use_philox) {
            const uint64_t key = (uint64_t)P.seed_state[0];
            key0 = (uint32_t)key;
            key1 = (uint32_t)(key >> 32);
            step = (uint32_t)P.seed_state[1];
        }
        uint32_t fi = 0, it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int row0 = (slot + t * P.G) * kHdRows;
            const int total = nb1 + 2 * nch;
            for (int i = 0; i < total; ++i, ++fi, ++it) {
                const bool pass1 = i < nb1;
                const int kb = pass1 ? i : 2 * c0 + (i - nb1);
                const int fs = fi % kHdXfStages, st = it % kHdXkStages;
                // weight block of pass 1: straight from the (L2-resident) bf16 copy
                uint4 wv[2];
                if (pass1) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = tid + q * kHdConvThreads;          // 16 B chunk of the [64 x 64] block
                        wv[q] = __ldg(reinterpret_cast<const uint4*>(H.wbf + (size_t)(ch >> 3) * K + kb * kHdBK) + (ch & 7));
                    }
                }
                hd_mbar_wait(&S.xfull[fs], (fi / kHdXfStages) & 1);
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    v[q] = *reinterpret_cast<const float4*>(&S.xf[fs][(rbase + 16 * q) * kHdBK + c4 * 4]);
                // dropout
                const bool in_range = kb >= 2 * c0 && kb < 2 * c1;        // block is revisited by pass 2
                if (H.mask) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = row0 + rbase + 16 * q;
                        if (r < P.B) {
                            const float4 m = __ldg(reinterpret_cast<const float4*>(H.mask + (size_t)r * H.ld_mask + kb * kHdBK) + c4);
                            v[q].x *= m.x; v[q].y *= m.y; v[q].z *= m.z; v[q].w *= m.w;
                        }
                    }
                } else if (P.use_philox) {
                    uint32_t bits = 0;
                    if (pass1) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            uint32_t c[4] = {(uint32_t)(row0 + rbase + 16 * q), (uint32_t)(kb * 16 + c4), (uint32_t)hi, step};
                            hd_philox(c, key0, key1);
#pragma unroll
                            for (int e = 0; e < 4; ++e) bits |= (c[e] >= P.thresh ? 1u : 0u) << (4 * q + e);
                        }
                        if (in_range) S.bitmap[kb - 2 * c0][tid] = bits;
                    } else {
                        bits = S.bitmap[kb - 2 * c0][tid];
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[q].x = (bits >> (4 * q + 0)) & 1u ? v[q].x * P.keep_scale : 0.f;
                        v[q].y = (bits >> (4 * q + 1)) & 1u ? v[q].y * P.keep_scale : 0.f;
                        v[q].z = (bits >> (4 * q + 2)) & 1u ? v[q].z * P.keep_scale : 0.f;
                        v[q].w = (bits >> (4 * q + 3)) & 1u ? v[q].w * P.keep_scale : 0.f;
                    }
                }
                uint2 o[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    o[q] = make_uint2(hd_pack2(v[q].x, v[q].y), hd_pack2(v[q].z, v[q].w));
                    // (volatile asm statements keep their order: the packed values -- hence the
                    // loads they depend on -- are complete before the arrive below is issued)
                    asm volatile("" ::"r"(o[q].x), "r"(o[q].y));
                }
                // The float32 stage is released only AFTER its values have been consumed (packed):
                // mbarrier arrives are not ordered behind shared-memory loads still in flight -- with
                // the load/store unit busy (the epilogue's logit stores) the arrive overtook the loads
                // and the next TMA block landed in the stage before some rows had been read
                // (measured: rows of one converter warp held the block loaded three steps later).
                hd_mbar_arrive(&S.xfree[fs]);
                hd_mbar_wait(&S.empty[st], ((it / kHdXkStages) & 1) ^ 1);
                uint8_t* xk = reinterpret_cast<uint8_t*>(S.xk[st]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = rbase + 16 * q;
                    *reinterpret_cast<uint2*>(xk + r * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + ((c4 & 1) << 3)) = o[q];
                }
                if (pass1) {
                    uint8_t* wk = reinterpret_cast<uint8_t*>(S.wk[st]);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = tid + q * kHdConvThreads;
                        const int n = ch >> 3, c8 = ch & 7;
                        *reinterpret_cast<uint4*>(wk + n * 128 + ((c8 ^ (n & 7)) << 4)) = wv[q];
                    }
                }
                hd_fence_async();
                hd_mbar_arrive(&S.conv[st]);
            }
        }
    } else {
        // ================= epilogue (warps 10-13; TMEM lane quarter = warp % 4) =======================
        const int q = warp & 3;
        const int row_l = q * 32 + lane;
        const int C = H.C;
        float loss_acc = 0.f;
        float db_acc[2] = {0.f, 0.f};
        const bool lead = split == 0;                            // writes logits / loss / d_bias
        for (int t = 0; t < my_tiles; ++t) {
            const int row = (slot + t * P.G) * kHdRows + row_l;
            const bool valid = row < P.B;
            hd_mbar_wait(&S.logits_full, t & 1);
            hd_tc_after();
            float g[64];
            {
                uint32_t u[64];
                hd_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16), u);
                hd_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32, u + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 64; ++c) g[c] = __uint_as_float(u[c]);
            }
            if (H.bias) {
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < C) g[c] += __ldg(H.bias + c);
            }
            if (valid && lead && H.logits) {
                float* o = H.logits + (size_t)row * H.ld_logits;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < C) o[c] = g[c];
            }
            if (H.loss == AGX_HEAD_CE) {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < C) m = fmaxf(m, g[c]);
                int y = 0;
                float w = 0.f;
                if (valid) {
                    y = (int)__ldg(H.labels + row);
                    w = H.class_w ? __ldg(H.class_w + y) : 1.0f;
                }
                float s = 0.f, ly = 0.f;
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    if (c == y) ly = g[c] - m;
                    g[c] = c < C ? __expf(g[c] - m) : 0.f;
                    s += g[c];
                }
                const float inv_s = 1.0f / s;
                const float scale = valid ? H.coef * w / __ldg(H.norm) : 0.f;
#pragma unroll
                for (int c = 0; c < 64; ++c) g[c] = scale * (g[c] * inv_s - (c == y ? 1.0f : 0.f));
                if (valid) loss_acc += scale * (__logf(s) - ly);         // -scale * log softmax[y]
            } else {
                const float* tg = H.target + (size_t)(valid ? row : 0) * H.ld_target;
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    float d = 0.f;
                    if (valid && c < C) d = g[c] - __ldg(tg + c);
                    const float ad = fabsf(d);
                    loss_acc += H.inv_count * (ad < 1.0f ? 0.5f * d * d : ad - 0.5f);
                    g[c] = H.inv_count * (ad < 1.0f ? d : (d > 0.f ? 1.0f : -1.0f));
                }
            }
            // d_bias: column sums over the tile's rows (fixed butterfly order)
            if (lead && H.db_part) {
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const float sum = warp_sum(g[c]);
                    if (lane == (c & 31)) db_acc[c >> 5] += sum;
                }
            }
            // logit gradient tile, bf16 [row][class], 128 B rows, 128B swizzle (MN-major B operand)
            uint8_t* dl = reinterpret_cast<uint8_t*>(S.dl);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 o = make_uint4(hd_pack2(g[8 * j], g[8 * j + 1]), hd_pack2(g[8 * j + 2], g[8 * j + 3]),
                                           hd_pack2(g[8 * j + 4], g[8 * j + 5]), hd_pack2(g[8 * j + 6], g[8 * j + 7]));
                *reinterpret_cast<uint4*>(dl + row_l * 128 + ((j ^ (row_l & 7)) << 4)) = o;
            }
            hd_fence_async();
            hd_tc_before();
            hd_mbar_arrive(&S.dl_ready);
        }
        if (my_tiles > 0) {
            // ---- loss / d_bias partials of this CTA ------------------------------------------------
            if (lead) {
                const float ls = warp_sum(loss_acc);
                if (lane == 0) S.s_loss[q] = ls;
                S.s_db[q][lane] = db_acc[0];
                S.s_db[q][32 + lane] = db_acc[1];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (lead && q == 0) {
                if (lane == 0) H.loss_part[slot] = (S.s_loss[0] + S.s_loss[1]) + (S.s_loss[2] + S.s_loss[3]);
                if (H.db_part) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int c = lane + 32 * k;
                        H.db_part[(size_t)slot * kHdCP + c] = (S.s_db[0][c] + S.s_db[1][c]) + (S.s_db[2][c] + S.s_db[3][c]);
                    }
                }
            }
            // ---- d_weight accumulators -> this CTA's partial [C][K] -------------------------------------
            hd_mbar_wait(&S.dw_full, 0);
            hd_tc_after();
            for (int j = 0; j < nch; ++j) {
                uint32_t u[64];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + kHdCP + (uint32_t)j * kHdCP;
                hd_tmem_ld32(ta, u);
                hd_tmem_ld32(ta + 32, u + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int col = (c0 + j) * 128 + row_l;          // accumulator lane = input column
                float* o = H.dw_part + (size_t)slot * C * K + col;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < C) o[(size_t)c * K] = __uint_as_float(u[c]);
            }
        }
    }

    hd_tc_before();
    __syncthreads();
    if (warp == 0) {
        hd_tc_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- phase 1: bf16 weights + label normaliser ---------------------------------------------------
struct HdPrep {
    const float* weight[AGX_MAX_HEADS];
    int64_t ldw[AGX_MAX_HEADS];
    __nv_bfloat16* wbf[AGX_MAX_HEADS];
    const int64_t* labels[AGX_MAX_HEADS];
    const float* class_w[AGX_MAX_HEADS];
    float* norm[AGX_MAX_HEADS];                    // NULL: not a CE head
    int32_t C[AGX_MAX_HEADS];
    int32_t n_heads, B, K, wblocks;
};

__global__ void __launch_bounds__(1024) head_prepare(const __grid_constant__ HdPrep P) {
    const int hi = (int)blockIdx.x / (P.wblocks + 1), b = (int)blockIdx.x % (P.wblocks + 1);
    if (b < P.wblocks) {
        // [64, K] bf16, rows >= C zero
        const int64_t n4 = (int64_t)kHdCP * P.K / 4;
        for (int64_t i = (int64_t)b * 1024 + threadIdx.x; i < n4; i += (int64_t)P.wblocks * 1024) {
            const int64_t e = i * 4;
            const int r = (int)(e / P.K), c = (int)(e % P.K);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < P.C[hi]) v = __ldg(reinterpret_cast<const float4*>(P.weight[hi] + (size_t)r * P.ldw[hi] + c));
            *reinterpret_cast<uint2*>(P.wbf[hi] + e) = make_uint2(hd_pack2(v.x, v.y), hd_pack2(v.z, v.w));
        }
        return;
    }
    if (!P.norm[hi]) return;
    __shared__ float red[1024];
    float s = 0.f;
    for (int r = threadIdx.x; r < P.B; r += 1024)
        s += P.class_w[hi] ? __ldg(P.class_w[hi] + __ldg(P.labels[hi] + r)) : 1.0f;
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *P.norm[hi] = red[0];
}

// ---- partials of the CTAs -> gradients (fixed order) -------------------------------------------------
struct HdReduce {
    const float* dw_part[AGX_MAX_HEADS];
    const float* db_part[AGX_MAX_HEADS];
    const float* loss_part[AGX_MAX_HEADS];
    float* d_weight[AGX_MAX_HEADS];
    int64_t ld_dw[AGX_MAX_HEADS];
    float* d_bias[AGX_MAX_HEADS];
    int32_t C[AGX_MAX_HEADS];
    float* loss;
    int32_t n_heads, K, G, accumulate, blocks_per_head;
};

__global__ void __launch_bounds__(256) head_reduce(const __grid_constant__ HdReduce P) {
    const int hi = (int)blockIdx.x / P.blocks_per_head, b = (int)blockIdx.x % P.blocks_per_head;
    const int C = P.C[hi], K = P.K;
    const int64_t n = (int64_t)C * K;
    // one element per thread; the G partials are summed in CTA order (deterministic), eight loads
    // in flight (a serial chain of G L2 round trips took 25-35 us at G = 32)
    for (int64_t e = (int64_t)b * 256 + threadIdx.x; e < n; e += (int64_t)P.blocks_per_head * 256) {
        const float* src = P.dw_part[hi] + e;
        float s = 0.f;
        int g = 0;
        for (; g + 8 <= P.G; g += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (size_t)(g + u) * n);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; g < P.G; ++g) s += __ldcg(src + (size_t)g * n);
        float* o = P.d_weight[hi] + (size_t)(e / K) * P.ld_dw[hi] + (e % K);
        *o = P.accumulate ? *o + s : s;
    }
    if (b == 0 && P.d_bias[hi] && (int)threadIdx.x < C) {
        float s = 0.f;
        for (int g = 0; g < P.G; ++g) s += __ldcg(P.db_part[hi] + (size_t)g * kHdCP + threadIdx.x);
        float* o = P.d_bias[hi] + threadIdx.x;
        *o = P.accumulate ? *o + s : s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // loss: sum over heads and CTAs, lane-strided then a fixed butterfly
        float s = 0.f;
        for (int h = 0; h < P.n_heads; ++h)
            for (int g = threadIdx.x; g < P.G; g += 32) s += __ldcg(P.loss_part[h] + g);
        s = warp_sum(s);
        if (threadIdx.x == 0) *P.loss = P.accumulate ? *P.loss + s : s;
    }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_hdEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

static PFN_hdEncodeTiled hd_get_encode() {
    static PFN_hdEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_hdEncodeTiled)p;
    }
    return fn;
}

static int hd_make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    PFN_hdEncodeTiled enc = hd_get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return AGX_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kHdBK, (cuuint32_t)kHdRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("agx_head_step: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                  (long long)rows, (long long)cols, (long long)ld);
        return AGX_ERR_CUDA;
    }
    return AGX_OK;
}

struct HdPlan {
    int K, n_tiles, n_split, chunks_per_split, G;
    size_t off_wbf[AGX_MAX_HEADS], off_dw[AGX_MAX_HEADS], off_db[AGX_MAX_HEADS], off_loss[AGX_MAX_HEADS];
    size_t bytes;
};

static int hd_plan(const agx_head_t* h, int n_heads, int32_t B, HdPlan* pl) {
    AGX_CHECK_ARG(h && n_heads >= 1 && n_heads <= AGX_MAX_HEADS, "agx_head_step: n_heads=%d out of [1,%d]",
                  n_heads, AGX_MAX_HEADS);
    AGX_CHECK_ARG(B >= 1, "agx_head_step: B=%d", B);
    const int K = h[0].width[0] + h[0].width[1];
    for (int i = 0; i < n_heads; ++i) {
        AGX_CHECK_ARG(h[i].width[0] + h[i].width[1] == K, "agx_head_step: head %d: all heads must share K", i);
        AGX_CHECK_ARG(h[i].width[0] > 0 && h[i].width[0] % kHdBK == 0 && h[i].width[1] >= 0 && K % 128 == 0,
                      "agx_head_step: head %d: width[0]=%d must be a multiple of 64 and K=%d of 128", i,
                      h[i].width[0], K);
        AGX_CHECK_ARG(h[i].C >= 1 && h[i].C <= kHdCP, "agx_head_step: head %d: C=%d out of [1,64]", i, h[i].C);
        AGX_CHECK_ARG(h[i].part[0] && (h[i].width[1] == 0 || h[i].part[1]) && h[i].weight && h[i].d_weight,
                      "agx_head_step: head %d: null pointer", i);
        AGX_CHECK_ARG(h[i].loss == AGX_HEAD_CE ? h[i].labels != nullptr : h[i].target != nullptr,
                      "agx_head_step: head %d: labels / target missing", i);
        AGX_CHECK_ARG((h[i].ldw % 4) == 0 && (reinterpret_cast<uintptr_t>(h[i].weight) % 16) == 0,
                      "agx_head_step: head %d: weight rows must be 16-byte aligned", i);
    }
    pl->K = K;
    pl->n_tiles = (int)ceil_div(B, kHdRows);
    const int chunks = K / 128;
    pl->n_split = (int)ceil_div(chunks, kHdMaxChunks);
    pl->chunks_per_split = (int)ceil_div(chunks, pl->n_split);
    const int units = n_heads * pl->n_split;
    int G = kNumSMs / units;
    if (G < 1) G = 1;
    if (G > pl->n_tiles) G = pl->n_tiles;
    static const char* env_g = getenv("AGX_HEAD_G");          // debugging: CTAs per (head, split)
    if (env_g && atoi(env_g) >= 1 && atoi(env_g) < G) G = atoi(env_g);
    pl->G = G;
    size_t off = 0;
    for (int i = 0; i < n_heads; ++i) {
        pl->off_wbf[i] = off;
        off = align_up(off + (size_t)kHdCP * K * 2, 256);
        pl->off_dw[i] = off;
        off = align_up(off + (size_t)G * h[i].C * K * 4, 256);
        pl->off_db[i] = off;
        off = align_up(off + (size_t)G * kHdCP * 4, 256);
        pl->off_loss[i] = off;
        off = align_up(off + (size_t)G * 4, 256);
    }
    pl->bytes = off;
    return AGX_OK;
}

}  // namespace agx

using namespace agx;

extern "C" size_t agx_head_step_workspace_bytes(const agx_head_t* h_heads, int n_heads, int32_t B) {
    HdPlan pl;
    if (hd_plan(h_heads, n_heads, B, &pl) != AGX_OK) return 0;
    return pl.bytes;
}

extern "C" int agx_head_step_prepare(const agx_head_t* h, int n_heads, int32_t B, float* norm,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    HdPlan pl;
    int rc = hd_plan(h, n_heads, B, &pl);
    if (rc) return rc;
    AGX_CHECK_ARG(workspace && workspace_bytes >= pl.bytes && (reinterpret_cast<uintptr_t>(workspace) % 256) == 0,
                  "agx_head_step_prepare: workspace %zu < required %zu (or not 256-byte aligned)", workspace_bytes,
                  pl.bytes);
    HdPrep P;
    P.n_heads = n_heads;
    P.B = B;
    P.K = pl.K;
    P.wblocks = 4;
    for (int i = 0; i < n_heads; ++i) {
        P.weight[i] = h[i].weight;
        P.ldw[i] = h[i].ldw;
        P.wbf[i] = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(workspace) + pl.off_wbf[i]);
        P.labels[i] = h[i].labels;
        P.class_w[i] = h[i].class_w;
        P.norm[i] = (h[i].loss == AGX_HEAD_CE && norm) ? norm + i : nullptr;
        P.C[i] = h[i].C;
    }
    head_prepare<<<n_heads * (P.wblocks + 1), 1024, 0, (cudaStream_t)stream>>>(P);
    AGX_LAUNCH_CHECK("head_prepare");
    return AGX_OK;
}

extern "C" int agx_head_step(const agx_head_t* h, int n_heads, int32_t B, float p_drop,
                             const int64_t* seed_state, const float* norm, float* loss, int accumulate,
                             void* workspace, size_t workspace_bytes, void* stream) {
    HdPlan pl;
    int rc = hd_plan(h, n_heads, B, &pl);
    if (rc) return rc;
    AGX_CHECK_ARG(workspace && workspace_bytes >= pl.bytes && (reinterpret_cast<uintptr_t>(workspace) % 256) == 0,
                  "agx_head_step: workspace %zu < required %zu (or not 256-byte aligned)", workspace_bytes, pl.bytes);
    AGX_CHECK_ARG(loss, "agx_head_step: loss must not be null");
    AGX_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "agx_head_step: p_drop=%f", (double)p_drop);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    HdParams P;
    memset(&P, 0, sizeof(P));
    P.n_heads = n_heads;
    P.B = B;
    P.K = pl.K;
    P.n_tiles = pl.n_tiles;
    P.G = pl.G;
    P.n_split = pl.n_split;
    P.chunks_per_split = pl.chunks_per_split;
    P.seed_state = seed_state;
    P.use_philox = (p_drop > 0.f && seed_state) ? 1 : 0;
    P.thresh = (uint32_t)fmin(4294967295.0, (double)p_drop * 4294967296.0);
    P.keep_scale = 1.0f / (1.0f - p_drop);
    for (int i = 0; i < n_heads; ++i) {
        HdHeadDev& D = P.h[i];
        AGX_CHECK_ARG(h[i].loss != AGX_HEAD_CE || norm, "agx_head_step: norm must not be null for CE heads");
        for (int k = 0; k < 2; ++k) {
            if (h[i].width[k] == 0) continue;
            AGX_CHECK_ARG((h[i].ld[k] % 4) == 0 && (reinterpret_cast<uintptr_t>(h[i].part[k]) % 16) == 0,
                          "agx_head_step: head %d part %d: rows must be 16-byte aligned", i, k);
            rc = hd_make_map(&D.map[k], h[i].part[k], B, h[i].width[k], h[i].ld[k]);
            if (rc) return rc;
        }
        AGX_CHECK_ARG(!h[i].mask || ((h[i].ld_mask % 4) == 0 && (reinterpret_cast<uintptr_t>(h[i].mask) % 16) == 0),
                      "agx_head_step: head %d: mask rows must be 16-byte aligned", i);
        D.mask = h[i].mask;
        D.ld_mask = h[i].ld_mask;
        D.wbf = reinterpret_cast<const __nv_bfloat16*>(ws + pl.off_wbf[i]);
        D.bias = h[i].bias;
        D.labels = h[i].labels;
        D.class_w = h[i].class_w;
        D.norm = norm ? norm + i : nullptr;
        D.target = h[i].target;
        D.ld_target = h[i].ld_target;
        D.logits = h[i].logits;
        D.ld_logits = h[i].ld_logits;
        D.dw_part = reinterpret_cast<float*>(ws + pl.off_dw[i]);
        D.db_part = h[i].d_bias ? reinterpret_cast<float*>(ws + pl.off_db[i]) : nullptr;
        D.loss_part = reinterpret_cast<float*>(ws + pl.off_loss[i]);
        D.coef = h[i].coef;
        D.inv_count = h[i].inv_count;
        D.width0 = h[i].width[0];
        D.C = h[i].C;
        D.loss = h[i].loss;
    }
    static bool attr_set = false;
    const size_t smem = sizeof(HdSmem) + 1024;
    if (!attr_set) {
        AGX_CUDA(cudaFuncSetAttribute(head_fused_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    head_fused_bf16<<<n_heads * pl.n_split * pl.G, kHdThreads, smem, st>>>(P);
    AGX_LAUNCH_CHECK("head_fused_bf16");

    HdReduce R;
    memset(&R, 0, sizeof(R));
    R.n_heads = n_heads;
    R.K = pl.K;
    R.G = pl.G;
    R.accumulate = accumulate ? 1 : 0;
    R.loss = loss;
    R.blocks_per_head = (int)ceil_div((int64_t)kHdCP * pl.K, 256);
    for (int i = 0; i < n_heads; ++i) {
        R.dw_part[i] = P.h[i].dw_part;
        R.db_part[i] = P.h[i].db_part;
        R.loss_part[i] = P.h[i].loss_part;
        R.d_weight[i] = h[i].d_weight;
        R.ld_dw[i] = h[i].ld_dw;
        R.d_bias[i] = h[i].d_bias;
        R.C[i] = h[i].C;
    }
    head_reduce<<<n_heads * R.blocks_per_head, 256, 0, st>>>(R);
    AGX_LAUNCH_CHECK("head_reduce");
    return AGX_OK;
}
