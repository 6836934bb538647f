// K2 / K3: neighbour aggregation over CSR (forward) and CSC (transpose, backward).
//
//   agg_rows    one (sub-)warp per destination row; the row's neighbour lists of up to 8
//               relations are reduced and summed in one pass, so the per-destination-type
//               torch.add chain of to_hetero never materialises per-relation outputs.
//               128-bit feature loads (lane l owns columns 4l..4l+3), neighbour ids read 4-8
//               ahead so several 512 B row gathers are in flight per warp.
//   agg_chunks  edge-balanced variant for relations with few, very long rows (artwork -> style /
//               genre / tag ...): one warp per 128 consecutive CSR edges, partial row sums
//               ("fragments") combined by agg_chunks_fixup in fixed chunk order.
//
// Both accumulate in float32, in CSR (= edge list) order inside a row: no float atomics, results
// are run-to-run reproducible.  HBM-bound: algorithmic traffic per edge = F*sizeof(T) + 4 B.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "agx_common.cuh"

namespace agx {

constexpr int kAggThreads = 256;
constexpr int kAggWarps = kAggThreads / 32;

struct RowGroups {
    agx_row_group_t g[AGX_MAX_GROUPS];
    int32_t slot_start[AGX_MAX_GROUPS + 1];
    int32_t n;
    int32_t F;
};

struct ChunkSegs {
    agx_chunk_seg_t s[AGX_MAX_CHUNK_SEGS];
    int32_t chunk_start[AGX_MAX_CHUNK_SEGS + 1];
    int32_t n;
    int32_t F;
};

template <typename T, int VEC>
struct Vec {
    float v[VEC];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_vec(const T* p);

template <>
__device__ __forceinline__ Vec<float, 4> load_vec<float, 4>(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec<float, 4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
template <>
__device__ __forceinline__ Vec<float, 1> load_vec<float, 1>(const float* p) {
    Vec<float, 1> r;
    r.v[0] = __ldg(p);
    return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16, 8> load_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    Vec<__nv_bfloat16, 8> r;
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16, 1> load_vec<__nv_bfloat16, 1>(const __nv_bfloat16* p) {
    Vec<__nv_bfloat16, 1> r;
    r.v[0] = __bfloat162float(*p);
    return r;
}

// Gather loads that must stay batched: ptxas interleaves ordinary loads with their consumers
// (load, use, load, use: two rows in flight per warp); volatile asm keeps the U loads of a batch
// back to back so that all of them are in flight before the first use.
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> gather_vec(const T* p);

template <>
__device__ __forceinline__ Vec<float, 4> gather_vec<float, 4>(const float* p) {
    Vec<float, 4> r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ Vec<float, 1> gather_vec<float, 1>(const float* p) {
    Vec<float, 1> r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16, 8> gather_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p) {
    uint32_t w[4];
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    Vec<__nv_bfloat16, 8> r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16, 1> gather_vec<__nv_bfloat16, 1>(const __nv_bfloat16* p) {
    uint16_t h;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(h) : "l"(p));
    Vec<__nv_bfloat16, 1> r;
    r.v[0] = __uint_as_float((uint32_t)h << 16);
    return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const Vec<T, VEC>& a);

template <>
__device__ __forceinline__ void store_vec<float, 4>(float* p, const Vec<float, 4>& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
template <>
__device__ __forceinline__ void store_vec<float, 1>(float* p, const Vec<float, 1>& a) {
    *p = a.v[0];
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 8>(__nv_bfloat16* p,
                                                            const Vec<__nv_bfloat16, 8>& a) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 1>(__nv_bfloat16* p,
                                                            const Vec<__nv_bfloat16, 1>& a) {
    *p = __float2bfloat16(a.v[0]);
}

template <typename T>
__device__ __forceinline__ void store_elem(T* p, float v);
template <>
__device__ __forceinline__ void store_elem<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16(v);
}
// four consecutive output elements (16 B float / 8 B bfloat16, aligned)
template <typename T>
__device__ __forceinline__ void store_row4(T* p, const Vec<T, 4>& a);
template <>
__device__ __forceinline__ void store_row4<float>(float* p, const Vec<float, 4>& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
template <>
__device__ __forceinline__ void store_row4<__nv_bfloat16>(__nv_bfloat16* p,
                                                          const Vec<__nv_bfloat16, 4>& a) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.v[0], a.v[1]);
    const __nv_bfloat162 h1 = __floats2bfloat162_rn(a.v[2], a.v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0),
                                             *reinterpret_cast<const uint32_t*>(&h1));
}

// multiplier of the row gathered for CSR slot e from neighbour c: 1/nbr_scale[c] (transpose of
// scatter-mean) and / or the per-edge weight (GATConv attention).  Exactly 1.0f when neither is set.
__device__ __forceinline__ float edge_scale(const agx_rel_t& R, int e, int c) {
    float s = R.nbr_scale ? 1.0f / __ldg(R.nbr_scale + c) : 1.0f;
    if (R.edge_w) s *= __ldg(R.edge_w + (R.edge_w_idx ? __ldg(R.edge_w_idx + e) : e));
    return s;
}

// ------------------------------------------------------------------------------------------------
// agg_rows
// ------------------------------------------------------------------------------------------------
// Direct (un-staged) evaluation of one output row by LPR lanes: the slow path for rows whose
// neighbour list does not fit the shared-memory staging of agg_rows.
template <typename T, int VEC, int LPR>
__device__ __noinline__ void agg_row_direct(const agx_row_group_t& G, int row, int F, int l) {
    constexpr int U = 4;                           // neighbour rows in flight
    for (int c0 = l * VEC; c0 < F; c0 += LPR * VEC) {
        Vec<T, VEC> acc;
        acc.zero();
        for (int r = 0; r < G.n_rel; ++r) {
            const agx_rel_t& R = G.rel[r];
            const T* __restrict__ x = reinterpret_cast<const T*>(R.x) + c0;
            Vec<T, VEC> racc;
            racc.zero();
            const int end_r = __ldg(R.rowptr + row + 1);
            int e = __ldg(R.rowptr + row);
            for (; e + U <= end_r; e += U) {
                int c[U];
                Vec<T, VEC> v[U];
                float s[U];
#pragma unroll
                for (int u = 0; u < U; ++u) c[u] = __ldg(R.col + e + u);
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = load_vec<T, VEC>(x + (int64_t)c[u] * R.ldx);
                if (R.nbr_scale || R.edge_w) {
#pragma unroll
                    for (int u = 0; u < U; ++u) s[u] = edge_scale(R, e + u, c[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) racc.v[i] += v[u].v[i] * s[u];
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) racc.v[i] += v[u].v[i];
                }
            }
            for (; e < end_r; ++e) {
                const int c = __ldg(R.col + e);
                const Vec<T, VEC> v = load_vec<T, VEC>(x + (int64_t)c * R.ldx);
                const float s = edge_scale(R, e, c);
#pragma unroll
                for (int i = 0; i < VEC; ++i) racc.v[i] += v.v[i] * s;
            }
            if (R.row_cnt) {
                const float d = __ldg(R.row_cnt + row);
#pragma unroll
                for (int i = 0; i < VEC; ++i) racc.v[i] = racc.v[i] / d;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc.v[i] += racc.v[i];
        }
        T* o = reinterpret_cast<T*>(G.out) + (int64_t)row * G.ldo + c0;
        if (G.accumulate) {
            const Vec<T, VEC> old = load_vec<T, VEC>(o);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc.v[i] += old.v[i];
        }
        if (G.bias) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc.v[i] += __ldg(G.bias + c0 + i);
        }
        store_vec<T, VEC>(o, acc);
    }
}

// agg_rows: a warp owns kRowsPerWarp (16) consecutive output rows.
//   index phase   lane i walks the row extents and neighbour ids of ITS row for every relation of
//                 the group (coalesced rowptr reads across the warp, all relations issued before
//                 the first use) and stages up to kRowCap (id, relation, scale) triples in shared
//                 memory: two dependent memory round trips per 32 rows instead of three per row
//                 and relation -- short rows are latency-, not bandwidth-bound.
//   gather phase  the rows are processed one per sub-warp of LPR lanes: ALL feature rows of an
//                 output row are requested back to back (<= 16 x 512 B in flight per warp), then
//                 summed in edge order per relation, divided by the relation's count (mean) and
//                 added in relation order.
constexpr int kRowCap = 16;
constexpr int kRowsWarps = 4;
constexpr int kRowsThreads = kRowsWarps * 32;
// output rows per warp.  16, not 32: the big groups are ~1.25 "waves" of resident warps, and the
// work left after the first wave runs on a quarter-full machine; half-size units halve that tail.
constexpr int kRowsPerWarp = 16;

// MINB = resident CTAs per SM the register allocation is bounded for (launch bounds): 5 leaves the
// compiler ~100 registers (20 warps per SM), 8 caps it at 64 (32 warps per SM).
template <typename T, int VEC, int LPR, int MINB>
__global__ void __launch_bounds__(kRowsThreads, MINB)
agg_rows(const __grid_constant__ RowGroups P) {
    constexpr int RPW = 32 / LPR;                  // rows gathered concurrently by one warp
    // per staged edge: address of the source row, multiplier, and "closes its relation" marker
    // (0 = more edges of the relation follow; d >= 1 = last edge, divide the relation sum by d)
    __shared__ const T* s_ptr[kRowsWarps][kRowsPerWarp][kRowCap];
    __shared__ float s_scl[kRowsWarps][kRowsPerWarp][kRowCap];
    __shared__ float s_div[kRowsWarps][kRowsPerWarp][kRowCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    const int F = P.F;
    const int64_t slot = ((int64_t)blockIdx.x * kRowsWarps + w) * kRowsPerWarp + lane;
    const bool valid = lane < kRowsPerWarp && slot < P.slot_start[P.n];
    int gi = 0, row = 0, tot = 0;
    bool overflow = false;
    if (valid) {
        while (slot >= P.slot_start[gi + 1]) ++gi;
        row = (int)(slot - P.slot_start[gi]);
        const agx_row_group_t& G = P.g[gi];
        int beg[AGX_MAX_REL_PER_GROUP], end[AGX_MAX_REL_PER_GROUP];
#pragma unroll
        for (int r = 0; r < AGX_MAX_REL_PER_GROUP; ++r) {
            beg[r] = end[r] = 0;
            if (r < G.n_rel) {
                beg[r] = __ldg(G.rel[r].rowptr + row);
                end[r] = __ldg(G.rel[r].rowptr + row + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < AGX_MAX_REL_PER_GROUP; ++r) {
            if (r < G.n_rel && beg[r] < end[r]) {
                const agx_rel_t& R = G.rel[r];
                const float d = R.row_cnt ? __ldg(R.row_cnt + row) : 1.0f;
                const T* xb = reinterpret_cast<const T*>(R.x);
                for (int e = beg[r]; e < end[r]; ++e) {
                    if (tot < kRowCap) {
                        const int c = __ldg(R.col + e);
                        s_ptr[w][lane][tot] = xb + (int64_t)c * R.ldx;
                        s_scl[w][lane][tot] = edge_scale(R, e, c);
                        s_div[w][lane][tot] = (e + 1 == end[r]) ? d : 0.f;
                        ++tot;
                    } else {
                        overflow = true;
                    }
                }
            }
        }
    }
    __syncwarp();

    for (int j0 = 0; j0 < kRowsPerWarp; j0 += RPW) {
        const int j = j0 + sub;
        const bool vj = __shfl_sync(0xffffffffu, valid ? 1 : 0, j) != 0;
        const bool oj = __shfl_sync(0xffffffffu, overflow ? 1 : 0, j) != 0;
        const int gj = __shfl_sync(0xffffffffu, gi, j);
        const int rj = __shfl_sync(0xffffffffu, row, j);
        const int nj = __shfl_sync(0xffffffffu, tot, j);
        if (!vj) continue;
        const agx_row_group_t& G = P.g[gj];
        if (oj) {
            agg_row_direct<T, VEC, LPR>(G, rj, F, l);
            continue;
        }
        for (int c0 = l * VEC; c0 < F; c0 += LPR * VEC) {
            Vec<T, VEC> acc, racc;
            acc.zero();
            racc.zero();
            for (int base = 0; base < nj; base += 8) {          // <= 2 rounds (kRowCap = 16)
                Vec<T, VEC> v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (base + u < nj) v[u] = load_vec<T, VEC>(s_ptr[w][j][base + u] + c0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (base + u < nj) {
                        const float s = s_scl[w][j][base + u];      // 1.0 when unscaled: exact
                        const float d = s_div[w][j][base + u];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) racc.v[i] = fmaf(v[u].v[i], s, racc.v[i]);
                        if (d != 0.f) {                             // last edge of its relation
                            if (d > 1.f) {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) racc.v[i] = racc.v[i] / d;
                            }
#pragma unroll
                            for (int i = 0; i < VEC; ++i) {
                                acc.v[i] += racc.v[i];
                                racc.v[i] = 0.f;
                            }
                        }
                    }
                }
            }
            T* o = reinterpret_cast<T*>(G.out) + (int64_t)rj * G.ldo + c0;
            if (G.accumulate) {
                const Vec<T, VEC> old = load_vec<T, VEC>(o);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc.v[i] += old.v[i];
            }
            if (G.bias) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc.v[i] += __ldg(G.bias + c0 + i);
            }
            store_vec<T, VEC>(o, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// agg_chunks: edge-balanced, ONE launch.  A warp owns 128 consecutive CSR edges of one relation.
//
//   * the chunk's neighbour ids (and 1/nbr_scale) are staged once in shared memory, the first row
//     is found by a warp-wide 32-ary search, row extents are read 32 rows at a time;
//   * LPR = F/VEC lanes cover one feature row; the 32/LPR sub-warps take alternating edges of a row
//     segment (F=32: four 128 B gathers per load instruction instead of one) and their partial
//     sums are added in sub-warp order -- for LPR = 32 (F >= 128) that is plain edge order;
//   * a row that lies inside the chunk is written directly (mean: divided by its count); a row
//     that crosses chunk boundaries leaves per-chunk fragments (lead[chunk] = the part of a row
//     that started in an earlier chunk, trail[chunk] = the start of a row that continues), and
//     the LAST warp to deliver a fragment of that row (one counter per row, self-resetting) adds
//     them up in chunk order: trail[c0] + lead[c0+1] + ... -- a fixed order whoever does it, so
//     results are reproducible; no float atomics, no second kernel;
//   * rows without edges are zero-filled by the kernel too (no memset), dealt out evenly over the
//     relation's warps.
// ------------------------------------------------------------------------------------------------
constexpr int kGatherDepth = 4;                    // feature-row loads in flight per lane

// fragment loads: from L2 (other SMs wrote them), volatile so that a batch stays in flight together
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_frag(const float* p) {
    Vec<T, VEC> r;
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(r.v[i]), "=f"(r.v[i + 1]), "=f"(r.v[i + 2]), "=f"(r.v[i + 3])
                         : "l"(p + i));
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r.v[i]) : "l"(p + i));
    }
    return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void store_frag(float* p, const Vec<T, VEC>& a) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
            *reinterpret_cast<float4*>(p + i) = make_float4(a.v[i], a.v[i + 1], a.v[i + 2], a.v[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = a.v[i];
    }
}

// rows [r0, r1) of the segment's output have no edges: write zeros (whole warp, VEC-wide stores)
template <typename T, int VEC>
__device__ __forceinline__ void zero_rows(const agx_chunk_seg_t& S, int F, int r0, int r1, int lane) {
    Vec<T, VEC> z;
    z.zero();
    for (int r = r0; r < r1; ++r)
        for (int c0 = lane * VEC; c0 < F; c0 += 32 * VEC)
            store_vec<T, VEC>(reinterpret_cast<T*>(S.out) + (int64_t)r * S.ldo + c0, z);
}

// ---- TMA row ring (agg_chunks<.., TMA = true>) ------------------------------------------------
// Feature rows are not gathered through registers but by the TMA engine: per warp a ring of
// kRingGroups x kRingRows row slots in shared memory, one cp.async.bulk (global -> shared, one whole
// feature row, <= 512 B) per neighbour, completion counted in bytes on one mbarrier per group.  Up
// to kRingGroups * kRingRows rows (12 KB at F = 128) are in flight per warp without holding a
// single register, and the summation reads 16 B per lane from shared memory in edge order.
constexpr int kRingRows = 8;
constexpr int kRingGroups = 3;
constexpr int kRingMaxRowBytes = 512;

__device__ __forceinline__ uint32_t agg_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void agg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(agg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void agg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(agg_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void agg_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(agg_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void agg_bulk_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            agg_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(agg_smem_u32(bar))
        : "memory");
}

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_smem_vec(const uint8_t* p) {
    Vec<T, VEC> r;
    if constexpr (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t wd[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r.v[2 * i] = __uint_as_float(wd[i] << 16);
            r.v[2 * i + 1] = __uint_as_float(wd[i] & 0xffff0000u);
        }
    }
    return r;
}

constexpr int kCtaEdges = kAggWarps * AGX_CHUNK_EDGES;      // 1024 edges per CTA (smallest unit)
constexpr int kMaxChunkF = 256;                              // columns per launch (host loops)
constexpr int kChunkCtasPerSm = 4;                           // residency of the register path

// CE = edges per warp (128, 192 or 256): the host picks the unit that fills whole waves of
// resident CTAs (876 CTAs of 1024 edges on 592 slots ran as 1.48 waves: mean SM-active time 45 us
// of a 77 us launch; 584 CTAs of 1536 edges are one wave).
template <typename T, int VEC, int LPR, bool TMA, int CE>
__global__ void __launch_bounds__(kAggThreads, TMA ? 2 : kChunkCtasPerSm)
agg_chunks(const __grid_constant__ ChunkSegs P) {
    constexpr int SUB = 32 / LPR;
    constexpr int U = kGatherDepth;
    constexpr int CTA_EDGES = kAggWarps * CE;
    static_assert(CE % 32 == 0 && (!TMA || CE == AGX_CHUNK_EDGES), "chunk size");
    __shared__ int s_col[TMA ? kAggWarps : 1][CE];
    __shared__ float s_scl[TMA ? kAggWarps : 1][CE];
    __shared__ uint64_t s_bar[kAggWarps][kRingGroups];
    // row pieces a warp could not finish alone: [warp][0] = head (its first row began in an earlier
    // warp's edges), [warp][1] = tail (its last row goes on); stitched by warp 0 after the barrier
    __shared__ __align__(16) float s_part[kAggWarps][2][kMaxChunkF];
    __shared__ int s_head_row[kAggWarps], s_head_open[kAggWarps], s_tail_row[kAggWarps];
    __shared__ int s_row_beg[kAggWarps][2], s_row_end[kAggWarps][2];
    extern __shared__ __align__(128) uint8_t s_ring[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    // CTA -> (relation, 1024-edge block); warp w owns edges [128 w, 128 w + 128) of the block
    int si = 0;
    while ((int)blockIdx.x >= P.chunk_start[si + 1]) ++si;
    const agx_chunk_seg_t& S = P.s[si];
    const agx_rel_t& R = S.rel;
    const int F = P.F;
    const int cta = (int)blockIdx.x - P.chunk_start[si];          // block index inside the relation
    const int nctas = P.chunk_start[si + 1] - P.chunk_start[si];
    const int start = cta * CTA_EDGES + w * CE;
    const int end = min(S.n_edges, start + CE);
    const int n_rows = S.n_rows;
    float* lead = S.frag;                                         // [nctas][F]
    float* trail = S.frag + (size_t)nctas * F;
    const bool has_work = start < S.n_edges;
    if (lane == 0) {
        s_head_row[w] = -1;
        s_head_open[w] = 0;
        s_tail_row[w] = -1;
    }

    // ---- rows without edges: zeros.  The relation's rows are dealt out evenly over ALL its warps
    // (not to the chunk that holds the preceding edge: a source table that carries other ranks'
    // boundary rows has runs of tens of thousands of empty rows, which one warp would clear alone)
    {
        const int64_t gw = (int64_t)cta * kAggWarps + w, GW = (int64_t)nctas * kAggWarps;
        const int z0 = (int)((int64_t)n_rows * gw / GW), z1 = (int)((int64_t)n_rows * (gw + 1) / GW);
        for (int r = z0; r < z1; r += 32) {
            const int rr = r + lane;
            bool empty = false;
            if (rr < z1) empty = __ldg(R.rowptr + rr) == __ldg(R.rowptr + rr + 1);
            unsigned m = __ballot_sync(0xffffffffu, empty);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                zero_rows<T, VEC>(S, F, r + b, r + b + 1, lane);
            }
        }
    }

    if (has_work) {
    // ---- stage neighbour ids / scales ----------------------------------------------------------
    // lane i of register k holds edge k*32 + i of the chunk (register path: read by shuffles);
    // the TMA path issues its copies from the shared-memory copy
    int cr[CE / 32];
    float sr[CE / 32];
#pragma unroll
    for (int k = 0; k < CE / 32; ++k) {
        const int i = k * 32 + lane;
        cr[k] = start + i < end ? __ldg(R.col + start + i) : 0;
    }
#pragma unroll
    for (int k = 0; k < CE / 32; ++k) {
        const int i = k * 32 + lane;
        sr[k] = start + i < end ? edge_scale(R, start + i, cr[k]) : 1.0f;
        if constexpr (TMA) {
            s_col[w][i] = cr[k];
            s_scl[w][i] = sr[k];
        }
    }
    // ---- row of the first edge: largest r with rowptr[r] <= start (32-ary search) ---------------
    int lo = 0, hi = n_rows;                       // rowptr[lo] <= start < rowptr[hi]
    while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int p = lo + (lane + 1) * step;
        const int v = p < hi ? __ldg(R.rowptr + p) : INT_MAX;
        const int cnt = __popc(__ballot_sync(0xffffffffu, v <= start));
        lo += cnt * step;
        hi = min(hi, lo + step);
    }
    int row = lo;
    __syncwarp();
    // ---- TMA ring: start the first kRingGroups groups of row copies -------------------------------
    const int n_chunk = end - start;
    const int ngroups = (n_chunk + kRingRows - 1) / kRingRows;
    const uint32_t RB = (uint32_t)F * sizeof(T);               // bytes per feature row
    uint8_t* ring = s_ring + (size_t)w * kRingGroups * kRingRows * RB;
    int waited = 0;                                              // groups whose rows have landed
    auto issue_group = [&](int g) {
        const int r0 = g * kRingRows;
        const int cnt = min(kRingRows, n_chunk - r0);
        const int slot = g % kRingGroups;
        if (lane == 0) agg_mbar_expect_tx(&s_bar[w][slot], (uint32_t)cnt * RB);
        __syncwarp();
        if (lane < cnt)
            agg_bulk_row(ring + (size_t)(slot * kRingRows + lane) * RB,
                         reinterpret_cast<const uint8_t*>(R.x) +
                             (int64_t)s_col[TMA ? w : 0][r0 + lane] * R.ldx * (int64_t)sizeof(T),
                         RB, &s_bar[w][slot]);
    };
    if constexpr (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < kRingGroups; ++g) agg_mbar_init(&s_bar[w][g], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        for (int g = 0; g < kRingGroups && g < ngroups; ++g) issue_group(g);
    }

    int wrow = row;                                // row-extent window: lane i holds rowptr[wrow+1+i]
    int rp = wrow + 1 + lane <= n_rows ? __ldg(R.rowptr + wrow + 1 + lane) : INT_MAX;
    int rbeg = __ldg(R.rowptr + row);
    int e = start;
    while (true) {
        if (row - wrow >= 32) {
            wrow = row;
            rp = wrow + 1 + lane <= n_rows ? __ldg(R.rowptr + wrow + 1 + lane) : INT_MAX;
        }
        const int rend = __shfl_sync(0xffffffffu, rp, row - wrow);
        const int e1 = min(rend, end);
        const bool starts_here = rbeg >= start, ends_here = rend <= end;
        const int i0 = e - start, i1 = e1 - start;
        // ---- the row segment [e, e1) (F <= kMaxChunkF: at most LPR*VEC*? column blocks) ---------
        for (int cb = 0; cb < F; cb += LPR * VEC) {
            const int c0 = cb + l * VEC;
            const bool active = c0 < F;
            // lanes past the last column read column block 0 (in bounds) and drop the result
            const T* x = reinterpret_cast<const T*>(R.x) + (active ? c0 : 0);
            Vec<T, VEC> acc;
            acc.zero();
            if constexpr (TMA) {
                // rows arrive in the ring in edge order; one (segment x group) interval at a time
                int i = i0;
                while (i < i1) {
                    const int g = i / kRingRows;
                    const int b = min(i1, (g + 1) * kRingRows);
                    if (waited <= g) {
                        agg_mbar_wait(&s_bar[w][g % kRingGroups], (uint32_t)(g / kRingGroups) & 1u);
                        waited = g + 1;
                    }
                    const uint8_t* rows = ring + (size_t)(g % kRingGroups) * kRingRows * RB +
                                          (size_t)l * VEC * sizeof(T);
                    Vec<T, VEC> v[kRingRows / SUB];
#pragma unroll
                    for (int u = 0; u < kRingRows / SUB; ++u) {
                        const int e_ = i + u * SUB + sub;
                        if (e_ < b && active)
                            v[u] = load_smem_vec<T, VEC>(rows + (size_t)(e_ - g * kRingRows) * RB);
                    }
#pragma unroll
                    for (int u = 0; u < kRingRows / SUB; ++u) {
                        const int e_ = i + u * SUB + sub;
                        if (e_ < b && active) {
                            const float sc = s_scl[TMA ? w : 0][e_];
#pragma unroll
                            for (int k = 0; k < VEC; ++k) acc.v[k] += v[u].v[k] * sc;
                        }
                    }
                    i = b;
                    if (b == (g + 1) * kRingRows) {              // group drained: refill its slot
                        __syncwarp();
                        if (g + kRingGroups < ngroups) issue_group(g + kRingGroups);
                    }
                }
            } else {
                // batches of SUB*U edges aligned inside the chunk (never straddling a 32-edge
                // register); slots outside [i0, i1) are loaded too (their ids are valid rows) but
                // not added, so the U gathers of a batch are unconditional and stay in flight together
                constexpr int BATCH = SUB * U;
                for (int base = (i0 / BATCH) * BATCH; base < i1; base += BATCH) {
                    const int kreg = base >> 5;                  // warp-uniform
                    int ck = cr[0];
                    float sk = sr[0];
#pragma unroll
                    for (int q = 1; q < CE / 32; ++q)
                        if (kreg == q) {
                            ck = cr[q];
                            sk = sr[q];
                        }
                    const int j0 = (base & 31) + sub;
                    int ci[U];
                    float sc[U];
                    bool ok[U];
                    Vec<T, VEC> v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int idx = base + u * SUB + sub;
                        ci[u] = __shfl_sync(0xffffffffu, ck, j0 + u * SUB);
                        sc[u] = __shfl_sync(0xffffffffu, sk, j0 + u * SUB);
                        ok[u] = idx >= i0 && idx < i1;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = gather_vec<T, VEC>(x + (int64_t)ci[u] * R.ldx);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (ok[u]) {
#pragma unroll
                            for (int k = 0; k < VEC; ++k) acc.v[k] += v[u].v[k] * sc[u];
                        }
                }
            }
            if (SUB > 1) {                         // partial sums of the sub-warps, in sub-warp order
                Vec<T, VEC> tot;
#pragma unroll
                for (int k = 0; k < VEC; ++k) tot.v[k] = __shfl_sync(0xffffffffu, acc.v[k], l);
#pragma unroll
                for (int sw = 1; sw < SUB; ++sw)
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        tot.v[k] += __shfl_sync(0xffffffffu, acc.v[k], sw * LPR + l);
                acc = tot;
            }
            if (active && sub == 0) {
                if (starts_here && ends_here) {
                    if (R.row_cnt) {
                        const float d = __ldg(R.row_cnt + row);
#pragma unroll
                        for (int k = 0; k < VEC; ++k) acc.v[k] = acc.v[k] / d;
                    }
                    store_vec<T, VEC>(reinterpret_cast<T*>(S.out) + (int64_t)row * S.ldo + c0, acc);
                } else {
                    store_frag<T, VEC>(&s_part[w][starts_here ? 1 : 0][c0], acc);
                }
            }
        }
        if (!(starts_here && ends_here) && lane == 0) {          // hand the piece to the stitcher
            if (starts_here) {
                s_tail_row[w] = row;
                s_row_beg[w][1] = rbeg;
                s_row_end[w][1] = rend;
            } else {
                s_head_row[w] = row;
                s_head_open[w] = ends_here ? 0 : 1;
                s_row_beg[w][0] = rbeg;
                s_row_end[w][0] = rend;
            }
        }
        e = e1;
        if (rend > end) break;                     // the row continues in the next chunk
        // ---- next row with edges (the empty rows in between were cleared above) ------------------
        int nr = row + 1;
        while (nr < n_rows) {
            if (nr - wrow >= 32) {
                wrow = nr;
                rp = wrow + 1 + lane <= n_rows ? __ldg(R.rowptr + wrow + 1 + lane) : INT_MAX;
            }
            const unsigned m = __ballot_sync(0xffffffffu, rp > rend) & (0xffffffffu << (nr - wrow));
            const int stop = m ? wrow + (__ffs(m) - 1) : wrow + 32;
            nr = min(stop, n_rows);
            if (m) break;
        }
        if (e >= end || nr >= n_rows) break;
        row = nr;
        rbeg = rend;
    }
    }   // has_work

    // ---- stitch the pieces of rows that cross warps (warp 0, warps in edge order) ---------------
    __syncthreads();
    if (w != 0) return;
    Vec<float, 4> carry[kMaxChunkF / 128];          // lane owns columns 4*lane + 128*j
    int carry_row = -1, carry_beg = 0, carry_end = 0;
    bool carry_outside = false;
    // finished row piece: a whole row -> out; the start / continuation of a row that crosses CTAs
    // -> global fragment + arrival counter, the last CTA to arrive adds the fragments in CTA order
    auto emit = [&](bool complete, bool is_lead) {
        const float d = R.row_cnt ? __ldg(R.row_cnt + carry_row) : 1.0f;
        if (complete) {
#pragma unroll
            for (int j = 0; j < kMaxChunkF / 128; ++j) {
                const int c0 = j * 128 + lane * 4;
                if (c0 < F) {
                    Vec<T, 4> o;
#pragma unroll
                    for (int k = 0; k < 4; ++k) o.v[k] = R.row_cnt ? carry[j].v[k] / d : carry[j].v[k];
                    if (c0 + 3 < F && VEC >= 4) {
                        store_row4<T>(reinterpret_cast<T*>(S.out) + (int64_t)carry_row * S.ldo + c0, o);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (c0 + k < F)
                                store_elem<T>(reinterpret_cast<T*>(S.out) + (int64_t)carry_row * S.ldo + c0 + k, o.v[k]);
                    }
                }
            }
            return;
        }
        float* f = (is_lead ? lead : trail) + (size_t)cta * F;
#pragma unroll
        for (int j = 0; j < kMaxChunkF / 128; ++j) {
            const int c0 = j * 128 + lane * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c0 + k < F) f[c0 + k] = carry[j].v[k];
        }
        const int c_first = carry_beg / CTA_EDGES, c_last = (carry_end - 1) / CTA_EDGES;
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(S.counters + c_first, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != c_last - c_first) return;
        __threadfence();
        constexpr int UC = U;
        const bool v4ok = (F & 3) == 0;                          // fragment rows 16 B aligned
        for (int c0 = lane * 4; c0 < F; c0 += 128) {
            const int nc = min(4, F - c0);
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < nc; ++k) tot[k] = __ldcg(trail + (size_t)c_first * F + c0 + k);
            for (int c = c_first + 1; c <= c_last; c += UC) {
                float v[UC][4];
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const float* q = lead + (size_t)min(c + u, c_last) * F + c0;
                    if (v4ok) {
                        asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]) : "l"(q));
                    } else {
                        for (int k = 0; k < 4; ++k) v[u][k] = k < nc ? __ldcg(q + k) : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < UC; ++u)
                    if (c + u <= c_last) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) tot[k] += v[u][k];
                    }
            }
            for (int k = 0; k < nc; ++k)
                store_elem<T>(reinterpret_cast<T*>(S.out) + (int64_t)carry_row * S.ldo + c0 + k,
                              R.row_cnt ? tot[k] / d : tot[k]);
        }
        if (lane == 0) S.counters[c_first] = 0;                  // ready for the next launch
    };
    auto load_part = [&](int ww, int which, bool add) {
#pragma unroll
        for (int j = 0; j < kMaxChunkF / 128; ++j) {
            const int c0 = j * 128 + lane * 4;
            if (c0 < F) {
                const float4 t = *reinterpret_cast<const float4*>(&s_part[ww][which][c0]);
                if (add) {
                    carry[j].v[0] += t.x; carry[j].v[1] += t.y; carry[j].v[2] += t.z; carry[j].v[3] += t.w;
                } else {
                    carry[j].v[0] = t.x; carry[j].v[1] = t.y; carry[j].v[2] = t.z; carry[j].v[3] = t.w;
                }
            }
        }
    };
    for (int ww = 0; ww < kAggWarps; ++ww) {
        const int hr = s_head_row[ww], tr = s_tail_row[ww];
        if (hr >= 0) {
            if (carry_row < 0) {                    // the row began before this CTA's edges
                load_part(ww, 0, false);
                carry_row = hr;
                carry_beg = s_row_beg[ww][0];
                carry_end = s_row_end[ww][0];
                carry_outside = true;
            } else {
                load_part(ww, 0, true);
            }
            if (!s_head_open[ww]) {                 // the row ends inside warp ww's edges
                emit(!carry_outside, true);
                carry_row = -1;
            }
        }
        if (tr >= 0) {
            load_part(ww, 1, false);
            carry_row = tr;
            carry_beg = s_row_beg[ww][1];
            carry_end = s_row_end[ww][1];
            carry_outside = false;
        }
    }
    if (carry_row >= 0) emit(false, carry_outside);   // goes on in the next CTA
}

static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T, int VEC>
static int launch_rows(const RowGroups& P, int lpr, int64_t slots, cudaStream_t st) {
    static const bool occ8 = getenv("AGX_ROWS_OCC8") != nullptr;       // A/B switch (round 2)
#define AGX_ROWS_CASE(L)                                                                  \
    case L: {                                                                             \
        const int64_t per_block = (int64_t)kRowsWarps * kRowsPerWarp;                      \
        const unsigned grid = (unsigned)ceil_div(slots, per_block);                        \
        if (occ8) agg_rows<T, VEC, L, 8><<<grid, kRowsThreads, 0, st>>>(P);                \
        else agg_rows<T, VEC, L, 5><<<grid, kRowsThreads, 0, st>>>(P);                     \
        break;                                                                            \
    }
    switch (lpr) {
        AGX_ROWS_CASE(32)
        AGX_ROWS_CASE(16)
        AGX_ROWS_CASE(8)
        AGX_ROWS_CASE(4)
        default:
            set_error("agx_aggregate_rows: bad lanes-per-row %d", lpr);
            return AGX_ERR_INVALID;
    }
#undef AGX_ROWS_CASE
    AGX_LAUNCH_CHECK("agg_rows");
    return AGX_OK;
}

}  // namespace agx

using namespace agx;

static int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

extern "C" int agx_aggregate_rows(const agx_row_group_t* h_groups, int n_groups, int F, int dtype,
                                  void* stream) {
    AGX_CHECK_ARG(h_groups && n_groups >= 1 && n_groups <= AGX_MAX_GROUPS,
                  "agx_aggregate_rows: n_groups=%d out of [1,%d]", n_groups, AGX_MAX_GROUPS);
    AGX_CHECK_ARG(F >= 1, "agx_aggregate_rows: F=%d", F);
    AGX_CHECK_ARG(dtype == AGX_F32 || dtype == AGX_BF16, "agx_aggregate_rows: bad dtype %d", dtype);
    const size_t esz = dtype == AGX_F32 ? 4 : 2;
    const int wide = dtype == AGX_F32 ? 4 : 8;
    RowGroups P;
    P.n = n_groups;
    P.F = F;
    P.slot_start[0] = 0;
    bool vec_ok = (F % wide) == 0;
    for (int g = 0; g < n_groups; ++g) {
        const agx_row_group_t& G = h_groups[g];
        AGX_CHECK_ARG(G.n_rows >= 0 && G.n_rel >= 0 && G.n_rel <= AGX_MAX_REL_PER_GROUP,
                      "agx_aggregate_rows: group %d: n_rows=%d n_rel=%d", g, G.n_rows, G.n_rel);
        AGX_CHECK_ARG(G.n_rows == 0 || G.out, "agx_aggregate_rows: group %d: null out", g);
        vec_ok = vec_ok && aligned_to(G.out, 16) && (G.ldo * esz) % 16 == 0;
        for (int r = 0; r < G.n_rel; ++r) {
            const agx_rel_t& R = G.rel[r];
            AGX_CHECK_ARG(R.rowptr && (R.col || G.n_rows == 0) && R.x,
                          "agx_aggregate_rows: group %d rel %d: null pointer", g, r);
            vec_ok = vec_ok && aligned_to(R.x, 16) && (R.ldx * esz) % 16 == 0;
        }
        P.g[g] = G;
        P.slot_start[g + 1] = P.slot_start[g] + G.n_rows;
    }
    const int64_t slots = P.slot_start[n_groups];
    if (slots == 0) return AGX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AGX_F32) {
        if (vec_ok) return launch_rows<float, 4>(P, max(4, min(32, pow2_ceil(F / 4))), slots, st);
        return launch_rows<float, 1>(P, max(4, min(32, pow2_ceil(F))), slots, st);
    }
    if (vec_ok)
        return launch_rows<__nv_bfloat16, 8>(P, max(4, min(32, pow2_ceil(F / 8))), slots, st);
    return launch_rows<__nv_bfloat16, 1>(P, max(4, min(32, pow2_ceil(F))), slots, st);
}

constexpr int kMinCtaEdges = kAggWarps * 32;                 // smallest unit a launch may pick

extern "C" size_t agx_chunk_frag_floats(int64_t n_edges, int F) {
    const int fb = F < kMaxChunkF ? F : kMaxChunkF;             // one column block at a time
    return (size_t)2 * (size_t)ceil_div(n_edges > 0 ? n_edges : 1, kMinCtaEdges) * (size_t)fb;
}

extern "C" size_t agx_chunk_counters(int64_t n_edges) {
    return (size_t)ceil_div(n_edges > 0 ? n_edges : 1, kMinCtaEdges);
}

template <typename T, int VEC, int LPR, int CE>
static int launch_chunks_ce(const ChunkSegs& P, unsigned grid, bool tma, cudaStream_t st) {
    if constexpr (CE == AGX_CHUNK_EDGES) {
        if (tma) {
            const size_t smem = (size_t)kAggWarps * kRingGroups * kRingRows * (size_t)P.F * sizeof(T);
            static bool attr_set = false;
            if (!attr_set) {
                AGX_CUDA(cudaFuncSetAttribute(agg_chunks<T, VEC, LPR, true, CE>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kAggWarps * kRingGroups * kRingRows * kRingMaxRowBytes));
                attr_set = true;
            }
            agg_chunks<T, VEC, LPR, true, CE><<<grid, kAggThreads, smem, st>>>(P);
            AGX_LAUNCH_CHECK("agg_chunks");
            return AGX_OK;
        }
    }
    agg_chunks<T, VEC, LPR, false, CE><<<grid, kAggThreads, 0, st>>>(P);
    AGX_LAUNCH_CHECK("agg_chunks");
    return AGX_OK;
}

template <typename T, int VEC, int LPR>
static int launch_chunks_lpr(const ChunkSegs& P, unsigned grid, bool tma, int ce, cudaStream_t st) {
    switch (ce) {
        case 256: return launch_chunks_ce<T, VEC, LPR, 256>(P, grid, false, st);
        case 192: return launch_chunks_ce<T, VEC, LPR, 192>(P, grid, false, st);
        case 64: return launch_chunks_ce<T, VEC, LPR, 64>(P, grid, false, st);
        case 32: return launch_chunks_ce<T, VEC, LPR, 32>(P, grid, false, st);
        default: return launch_chunks_ce<T, VEC, LPR, AGX_CHUNK_EDGES>(P, grid, tma, st);
    }
}

template <typename T, int VEC>
static int launch_chunks(const ChunkSegs& P, int lpr, unsigned grid, bool tma, int ce,
                         cudaStream_t st) {
    switch (lpr) {
        case 32: return launch_chunks_lpr<T, VEC, 32>(P, grid, tma, ce, st);
        case 16: return launch_chunks_lpr<T, VEC, 16>(P, grid, tma, ce, st);
        default: return launch_chunks_lpr<T, VEC, 8>(P, grid, tma, ce, st);
    }
}

// Edges per warp for this launch.  Measured on B200 (round 2, full graph, 7 launches per step):
//     edges per warp    64      128      192      256
//     us per step       436     553      649      783
// Filling exactly one wave of resident CTAs with bigger units (the round-1 hypothesis) is SLOWER,
// smaller units are faster: a warp walks its edges in serial batches of kGatherDepth gathers (one
// L2 / DRAM round trip per batch), so a CTA lasts as long as its unit and the launch ends with the
// slowest CTA, while the gathers themselves are far from the L2 ceiling at any residency
// (profiles/probes/gather_probe2.cu: 12.5 TB/s with 2 CTAs per SM).  The same holds on launches
// of many waves (16x replicated graph, 14 M edges per launch: 5.5 ms per step at 64, 6.2 at 128,
// 6.5 at 256).  Default 64; AGX_CHUNK_CE=32|64|128|192|256 overrides.
static int pick_chunk_edges(const agx_chunk_seg_t* segs, int n_segs, bool tma) {
    static const char* env = getenv("AGX_CHUNK_CE");
    (void)segs;
    (void)n_segs;
    if (tma) return AGX_CHUNK_EDGES;
    if (env) {
        const int v = atoi(env);
        if (v == 32 || v == 64 || v == 128 || v == 192 || v == 256) return v;
    }
    return 64;
}

extern "C" int agx_aggregate_chunks(const agx_chunk_seg_t* h_segs, int n_segs, int F, int dtype,
                                    void* stream) {
    AGX_CHECK_ARG(h_segs && n_segs >= 1 && n_segs <= AGX_MAX_CHUNK_SEGS,
                  "agx_aggregate_chunks: n_segs=%d out of [1,%d]", n_segs, AGX_MAX_CHUNK_SEGS);
    AGX_CHECK_ARG(F >= 1, "agx_aggregate_chunks: F=%d", F);
    AGX_CHECK_ARG(dtype == AGX_F32 || dtype == AGX_BF16, "agx_aggregate_chunks: bad dtype %d",
                  dtype);
    const size_t esz = dtype == AGX_F32 ? 4 : 2;
    const int wide = dtype == AGX_F32 ? 4 : 8;
    cudaStream_t st = (cudaStream_t)stream;
    for (int s = 0; s < n_segs; ++s) {
        const agx_chunk_seg_t& S = h_segs[s];
        AGX_CHECK_ARG(S.n_rows >= 0 && S.n_edges >= 0, "agx_aggregate_chunks: seg %d sizes", s);
        AGX_CHECK_ARG(S.rel.rowptr && S.rel.x && (S.out || S.n_rows == 0),
                      "agx_aggregate_chunks: seg %d: null pointer", s);
        AGX_CHECK_ARG(S.n_edges == 0 || (S.rel.col && S.frag && S.counters),
                      "agx_aggregate_chunks: seg %d: null col/frag/counters", s);
        // a relation without edges is visited by no chunk: clear its output here
        if (S.n_edges == 0 && S.n_rows > 0)
            AGX_CUDA(cudaMemset2DAsync(S.out, (size_t)S.ldo * esz, 0, (size_t)F * esz,
                                       (size_t)S.n_rows, st));
    }
    // Measured on B200 (profiles/probes/gather_probe.cu, 895k random 512 B rows of an L2-resident 60 MB
    // table): register gathers 15.7 TB/s, TMA row ring 11.7-13.7 TB/s, and the ring costs ~3x the
    // instructions per row (one UBLKCP per row is issued lane by lane): the ring is kept as an
    // option (AGX_TMA_GATHER=1), the default is the register path.
    static const bool want_tma = getenv("AGX_TMA_GATHER") != nullptr;
    // the kernel covers up to kMaxChunkF columns (the row pieces are stitched in shared memory):
    // wider features take one launch per column block
    for (int cb = 0; cb < F; cb += kMaxChunkF) {
        const int Fb = F - cb < kMaxChunkF ? F - cb : kMaxChunkF;
        ChunkSegs P;
        P.n = n_segs;
        P.F = Fb;
        P.chunk_start[0] = 0;
        bool vec_ok = (Fb % wide) == 0;
        for (int s = 0; s < n_segs; ++s) {
            agx_chunk_seg_t S = h_segs[s];
            S.rel.x = reinterpret_cast<const char*>(S.rel.x) + (size_t)cb * esz;
            S.out = reinterpret_cast<char*>(S.out) + (size_t)cb * esz;
            vec_ok = vec_ok && aligned_to(S.out, 16) && (S.ldo * esz) % 16 == 0 &&
                     aligned_to(S.rel.x, 16) && (S.rel.ldx * esz) % 16 == 0 &&
                     (S.n_edges == 0 || aligned_to(S.frag, 16));
            P.s[s] = S;
        }
        const bool tma = vec_ok && want_tma && (size_t)Fb * esz <= (size_t)kRingMaxRowBytes;
        const int ce = pick_chunk_edges(h_segs, n_segs, tma);     // edges per warp of this launch
        for (int s = 0; s < n_segs; ++s)
            P.chunk_start[s + 1] =
                P.chunk_start[s] + (int32_t)ceil_div(P.s[s].n_edges, (int64_t)kAggWarps * ce);
        const unsigned grid = (unsigned)P.chunk_start[n_segs];
        if (grid == 0) return AGX_OK;
        int rc;
        if (dtype == AGX_F32) {
            rc = vec_ok ? launch_chunks<float, 4>(P, max(8, min(32, pow2_ceil(Fb / 4))), grid, tma, ce, st)
                        : launch_chunks<float, 1>(P, max(8, min(32, pow2_ceil(Fb))), grid, false, ce, st);
        } else {
            rc = vec_ok ? launch_chunks<__nv_bfloat16, 8>(P, max(8, min(32, pow2_ceil(Fb / 8))), grid,
                                                          tma, ce, st)
                        : launch_chunks<__nv_bfloat16, 1>(P, max(8, min(32, pow2_ceil(Fb))), grid,
                                                          false, ce, st);
        }
        if (rc) return rc;
    }
    return AGX_OK;
}
