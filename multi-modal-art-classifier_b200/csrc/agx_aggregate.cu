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
                if (R.nbr_scale) {
#pragma unroll
                    for (int u = 0; u < U; ++u) s[u] = 1.0f / __ldg(R.nbr_scale + c[u]);
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
                const float s = R.nbr_scale ? 1.0f / __ldg(R.nbr_scale + c) : 1.0f;
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
        store_vec<T, VEC>(o, acc);
    }
}

// agg_rows: a warp owns 32 consecutive output rows.
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

template <typename T, int VEC, int LPR>
__global__ void __launch_bounds__(kRowsThreads)
agg_rows(const __grid_constant__ RowGroups P) {
    constexpr int RPW = 32 / LPR;                  // rows gathered concurrently by one warp
    // per staged edge: address of the source row, multiplier, and "closes its relation" marker
    // (0 = more edges of the relation follow; d >= 1 = last edge, divide the relation sum by d)
    __shared__ const T* s_ptr[kRowsWarps][32][kRowCap];
    __shared__ float s_scl[kRowsWarps][32][kRowCap];
    __shared__ float s_div[kRowsWarps][32][kRowCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    const int F = P.F;
    const int64_t slot = ((int64_t)blockIdx.x * kRowsWarps + w) * 32 + lane;
    const bool valid = slot < P.slot_start[P.n];
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
                        s_scl[w][lane][tot] = R.nbr_scale ? 1.0f / __ldg(R.nbr_scale + c) : 1.0f;
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

    for (int j0 = 0; j0 < 32; j0 += RPW) {
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
            store_vec<T, VEC>(o, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// agg_chunks: warp per 128 CSR edges.  frag layout per segment: lead[chunks][F], trail[chunks][F]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_i32(const int32_t* a, int n, int key) {
    int lo = 0, hi = n;            // first index with a[idx] > key
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <typename T, int VEC>
__device__ __forceinline__ void accumulate_edges(Vec<T, VEC>& acc, const agx_rel_t& R, const T* x,
                                                 int e0, int e1, int lane, bool active) {
    // sequential (edge order) sum of rows col[e0..e1); neighbour ids fetched 32 at a time
    for (int e = e0; e < e1; e += 32) {
        const int n = min(32, e1 - e);
        int c = lane < n ? __ldg(R.col + e + lane) : 0;
        float s = 1.0f;
        if (R.nbr_scale && lane < n) s = 1.0f / __ldg(R.nbr_scale + c);
        for (int j0 = 0; j0 < n; j0 += 8) {
            Vec<T, VEC> v[8];
            float sj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int cj = __shfl_sync(0xffffffffu, c, (j0 + u) & 31);
                sj[u] = __shfl_sync(0xffffffffu, s, (j0 + u) & 31);
                if (j0 + u < n && active) v[u] = load_vec<T, VEC>(x + (int64_t)cj * R.ldx);
                else v[u].zero();
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j0 + u < n) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc.v[i] += v[u].v[i] * sj[u];
                }
        }
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kAggThreads)
agg_chunks(const __grid_constant__ ChunkSegs P) {
    const int lane = threadIdx.x & 31;
    const int64_t gchunk = (int64_t)blockIdx.x * kAggWarps + (threadIdx.x >> 5);
    if (gchunk >= P.chunk_start[P.n]) return;
    int si = 0;
    while (gchunk >= P.chunk_start[si + 1]) ++si;
    const agx_chunk_seg_t& S = P.s[si];
    const agx_rel_t& R = S.rel;
    const int F = P.F;
    const int chunk = (int)(gchunk - P.chunk_start[si]);
    const int nchunks = P.chunk_start[si + 1] - P.chunk_start[si];
    const int start = chunk * AGX_CHUNK_EDGES;
    const int end = min(S.n_edges, start + AGX_CHUNK_EDGES);
    float* lead = S.frag;
    float* trail = S.frag + (size_t)nchunks * F;

    for (int c0 = lane * VEC; c0 < ((F + 32 * VEC - 1) / (32 * VEC)) * (32 * VEC); c0 += 32 * VEC) {
        const bool active = c0 < F;
        const T* x = reinterpret_cast<const T*>(R.x) + c0;
        int row = upper_bound_i32(R.rowptr, S.n_rows + 1, start) - 1;
        int e = start;
        while (e < end) {
            const int rbeg = __ldg(R.rowptr + row), rend = __ldg(R.rowptr + row + 1);
            const int e1 = min(rend, end);
            Vec<T, VEC> acc;
            acc.zero();
            accumulate_edges<T, VEC>(acc, R, x, e, e1, lane, active);
            const bool starts_here = rbeg >= start, ends_here = rend <= end;
            if (active) {
                if (starts_here && ends_here) {
                    if (R.row_cnt) {
                        const float d = __ldg(R.row_cnt + row);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc.v[i] = acc.v[i] / d;
                    }
                    store_vec<T, VEC>(reinterpret_cast<T*>(S.out) + (int64_t)row * S.ldo + c0, acc);
                } else {
                    float* f = (starts_here ? trail : lead) + (size_t)chunk * F + c0;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) f[i] = acc.v[i];
                }
            }
            e = e1;
            if (e < end) {
                ++row;
                while (__ldg(R.rowptr + row + 1) <= e) ++row;     // skip empty rows
            }
        }
    }
}

// One warp per chunk that owns a spanning row (the row starts in the chunk and runs past its end):
// total = trail[c] + lead[c+1] + ... + lead[c_last], in chunk order.
template <typename T, int VEC>
__global__ void __launch_bounds__(kAggThreads)
agg_chunks_fixup(const __grid_constant__ ChunkSegs P) {
    const int lane = threadIdx.x & 31;
    // one CTA per chunk; the (rare) owning CTAs spread the row's fragments over their 8 warps:
    // warp w adds lead[chunk+1+w], lead[chunk+1+w+8], ... and warp 0 combines trail + the 8 warp
    // sums in warp order -- a fixed order, so the result is reproducible.
    __shared__ float wsum[kAggWarps][128 + 4];
    const int w = threadIdx.x >> 5;
    const int64_t gchunk = blockIdx.x;
    if (gchunk >= P.chunk_start[P.n]) return;
    int si = 0;
    while (gchunk >= P.chunk_start[si + 1]) ++si;
    const agx_chunk_seg_t& S = P.s[si];
    const agx_rel_t& R = S.rel;
    const int F = P.F;
    const int chunk = (int)(gchunk - P.chunk_start[si]);
    const int nchunks = P.chunk_start[si + 1] - P.chunk_start[si];
    const int start = chunk * AGX_CHUNK_EDGES;
    const int end = min(S.n_edges, start + AGX_CHUNK_EDGES);
    if (end >= S.n_edges) return;                       // last chunk cannot own a spanning row
    const int row = upper_bound_i32(R.rowptr, S.n_rows + 1, end - 1) - 1;   // row of the last edge
    const int rbeg = __ldg(R.rowptr + row), rend = __ldg(R.rowptr + row + 1);
    if (!(rend > end && rbeg >= start)) return;          // uniform over the CTA
    const int c_last = (rend - 1) / AGX_CHUNK_EDGES;
    const float* lead = S.frag;
    const float* trail = S.frag + (size_t)nchunks * F;
    for (int cb = 0; cb < F; cb += 128) {                // 128 columns per round (smem staging)
        const int cw = min(128, F - cb);
        for (int c0 = lane * VEC; c0 < cw; c0 += 32 * VEC) {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
            int c = chunk + 1 + w;
            for (; c + 3 * kAggWarps <= c_last; c += 4 * kAggWarps) {
                float t[4][VEC];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int i = 0; i < VEC; ++i)
                        t[u][i] = lead[(size_t)(c + u * kAggWarps) * F + cb + c0 + i];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[i] += t[u][i];
            }
            for (; c <= c_last; c += kAggWarps)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += lead[(size_t)c * F + cb + c0 + i];
#pragma unroll
            for (int i = 0; i < VEC; ++i) wsum[w][c0 + i] = acc[i];
        }
        __syncthreads();
        if (w == 0) {
            for (int c0 = lane * VEC; c0 < cw; c0 += 32 * VEC) {
                Vec<T, VEC> o;
                const float d = R.row_cnt ? __ldg(R.row_cnt + row) : 1.0f;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float a = trail[(size_t)chunk * F + cb + c0 + i];
#pragma unroll
                    for (int k = 0; k < kAggWarps; ++k) a += wsum[k][c0 + i];
                    o.v[i] = R.row_cnt ? a / d : a;
                }
                store_vec<T, VEC>(reinterpret_cast<T*>(S.out) + (int64_t)row * S.ldo + cb + c0, o);
            }
        }
        __syncthreads();
    }
}

static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T, int VEC>
static int launch_rows(const RowGroups& P, int lpr, int64_t slots, cudaStream_t st) {
#define AGX_ROWS_CASE(L)                                                                  \
    case L: {                                                                             \
        const int64_t per_block = (int64_t)kRowsWarps * 32;   /* 32 row slots per warp */   \
        agg_rows<T, VEC, L><<<(unsigned)ceil_div(slots, per_block), kRowsThreads, 0, st>>>(P); \
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

extern "C" size_t agx_chunk_frag_floats(int64_t n_edges, int F) {
    return (size_t)2 * (size_t)ceil_div(n_edges > 0 ? n_edges : 1, AGX_CHUNK_EDGES) * (size_t)F;
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
    ChunkSegs P;
    P.n = n_segs;
    P.F = F;
    P.chunk_start[0] = 0;
    bool vec_ok = (F % wide) == 0;
    for (int s = 0; s < n_segs; ++s) {
        const agx_chunk_seg_t& S = h_segs[s];
        AGX_CHECK_ARG(S.n_rows >= 0 && S.n_edges >= 0, "agx_aggregate_chunks: seg %d sizes", s);
        AGX_CHECK_ARG(S.rel.rowptr && S.rel.x && (S.out || S.n_rows == 0),
                      "agx_aggregate_chunks: seg %d: null pointer", s);
        AGX_CHECK_ARG(S.n_edges == 0 || (S.rel.col && S.frag),
                      "agx_aggregate_chunks: seg %d: null col/frag", s);
        vec_ok = vec_ok && aligned_to(S.out, 16) && (S.ldo * esz) % 16 == 0 &&
                 aligned_to(S.rel.x, 16) && (S.rel.ldx * esz) % 16 == 0;
        P.s[s] = S;
        P.chunk_start[s + 1] = P.chunk_start[s] + (int32_t)ceil_div(S.n_edges, AGX_CHUNK_EDGES);
        // rows without edges are never visited by a chunk: clear the output first
        if (S.n_rows > 0)
            AGX_CUDA(cudaMemset2DAsync(S.out, (size_t)S.ldo * esz, 0, (size_t)F * esz,
                                       (size_t)S.n_rows, st));
    }
    const int64_t chunks = P.chunk_start[n_segs];
    if (chunks == 0) return AGX_OK;
    const unsigned grid = (unsigned)ceil_div(chunks, kAggWarps);
    if (dtype == AGX_F32) {
        if (vec_ok) {
            agg_chunks<float, 4><<<grid, kAggThreads, 0, st>>>(P);
            AGX_LAUNCH_CHECK("agg_chunks");
            agg_chunks_fixup<float, 4><<<(unsigned)chunks, kAggThreads, 0, st>>>(P);
        } else {
            agg_chunks<float, 1><<<grid, kAggThreads, 0, st>>>(P);
            AGX_LAUNCH_CHECK("agg_chunks");
            agg_chunks_fixup<float, 1><<<(unsigned)chunks, kAggThreads, 0, st>>>(P);
        }
    } else {
        if (vec_ok) {
            agg_chunks<__nv_bfloat16, 8><<<grid, kAggThreads, 0, st>>>(P);
            AGX_LAUNCH_CHECK("agg_chunks");
            agg_chunks_fixup<__nv_bfloat16, 8><<<(unsigned)chunks, kAggThreads, 0, st>>>(P);
        } else {
            agg_chunks<__nv_bfloat16, 1><<<grid, kAggThreads, 0, st>>>(P);
            AGX_LAUNCH_CHECK("agg_chunks");
            agg_chunks_fixup<__nv_bfloat16, 1><<<(unsigned)chunks, kAggThreads, 0, st>>>(P);
        }
    }
    AGX_LAUNCH_CHECK("agg_chunks_fixup");
    return AGX_OK;
}
