// GATConv attention (the reference script's default --operator,
// /root/reference/src/train_gnn_embeddings.py:15,99; PyG 2.0.2 GATConv, one head):
//
//   e_ij = leaky_relu(a_l[j] + a_r[i]),  alpha_ij = exp(e_ij - max_i) / (sum_i + 1e-16),
//   out_i = sum_j alpha_ij x_l[j] (+ bias)
//
// over the CSR (by destination) of the relation's edge list with self loops added.  ArtGraph's
// relations into style / genre / tag / media have a few rows of 10^3..10^4 edges, so nothing here
// lets one warp walk a row of feature vectors:
//
//   scalar passes (this file, 4-12 B per edge)
//     gat_edge_softmax      alpha per CSR slot; 8 lanes per row up to 8 edges, a warp per row up to
//                           AGX_GAT_LONG_ROW, every longer row by a 1024-thread CTA of its own
//                           (three strided passes over the row, values parked in alpha)
//     gat_edge_softmax_bwd  de_ij = alpha_ij (dalpha_ij - sum_j alpha_ij dalpha_ij) leaky'(.),
//                           da_r[i] = sum_j de_ij, same row mapping
//     sddmm                 dalpha_ij = <dout[i], x_l[j]>, a warp per 32 CSR slots (edge-balanced)
//   wide passes (agx_aggregate.cu, 4 F B per edge)
//     out  = weighted neighbour sum over the CSR   (agx_rel_t.edge_w = alpha)
//     dx_l = weighted neighbour sum over the CSC   (edge_w = alpha through edge_w_idx)
//     da_l = sum over the CSC of de (F = 1)
//
// All reductions run in a fixed order (lane-strided partial sums, shuffle tree, warp order): results
// are reproducible run to run; float32 throughout.
#include "agx_common.cuh"

namespace agx {

constexpr int kGatThreads = 1024;                  // softmax kernels: a hub row gets all of them
constexpr int kGatWarps = kGatThreads / 32;
constexpr int kSddmmThreads = 256;
constexpr int kSddmmWarps = kSddmmThreads / 32;

struct GatRels {
    agx_gat_rel_t r[AGX_MAX_GAT_RELS];
    int32_t hub_start[AGX_MAX_GAT_RELS + 1];       // first one CTA per long row, per relation ...
    int32_t blk_start[AGX_MAX_GAT_RELS + 1];       // ... then CTA -> relation (128 rows per CTA)
    int32_t n;
    float slope;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : slope * v; }

// reduction over the NT threads that share a row: NT = 32 one warp, NT = kGatThreads the CTA
// (warp results combined in warp order by every thread; s_red is reused, hence the leading barrier)
template <int NT, bool MAX>
__device__ __forceinline__ float group_reduce(float v, float* s_red) {
    v = MAX ? warp_max(v) : warp_sum(v);
    if constexpr (NT > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        v = s_red[0];
#pragma unroll
        for (int w = 1; w < kGatWarps; ++w) v = MAX ? fmaxf(v, s_red[w]) : v + s_red[w];
    }
    return v;
}

// softmax of one destination row by NT threads (t = this thread's index among them)
template <int NT>
__device__ __forceinline__ void softmax_row(const agx_gat_rel_t& R, int row, int beg, int end, int t,
                                            float slope, float* s_red) {
    const float ar = __ldg(R.a_r + row);
    float m = -INFINITY;
#pragma unroll 4
    for (int e = beg + t; e < end; e += NT) {
        const float l = leaky(__ldg(R.a_l + __ldg(R.col + e)) + ar, slope);
        R.alpha[e] = l;                            // parked: the same thread reads it back
        m = fmaxf(m, l);
    }
    m = group_reduce<NT, true>(m, s_red);
    float s = 0.f;
#pragma unroll 4
    for (int e = beg + t; e < end; e += NT) {
        const float p = expf(R.alpha[e] - m);
        R.alpha[e] = p;
        s += p;
    }
    s = group_reduce<NT, false>(s, s_red);
    const float inv = 1.0f / (s + 1e-16f);
#pragma unroll 4
    for (int e = beg + t; e < end; e += NT) R.alpha[e] *= inv;
}

template <int NT>
__device__ __forceinline__ void softmax_bwd_row(const agx_gat_rel_t& R, int row, int beg, int end,
                                                int t, float slope, float* s_red) {
    const float ar = __ldg(R.a_r + row);
    float s = 0.f;
#pragma unroll 4
    for (int e = beg + t; e < end; e += NT) s = fmaf(R.alpha[e], __ldg(R.dalpha + e), s);
    s = group_reduce<NT, false>(s, s_red);
    float d = 0.f;
#pragma unroll 4
    for (int e = beg + t; e < end; e += NT) {
        const float raw = __ldg(R.a_l + __ldg(R.col + e)) + ar;
        const float de = R.alpha[e] * (__ldg(R.dalpha + e) - s) * (raw > 0.f ? 1.0f : slope);
        R.de[e] = de;
        d += de;
    }
    d = group_reduce<NT, false>(d, s_red);
    if (t == 0) R.da_r[row] = d;
}

// A row of at most NL (8 or 32) edges by NL lanes, one edge per lane, everything in registers:
// three dependent memory round trips (row extent, source id, logit) instead of the nine of the
// strided three-pass walk -- rows of 1..8 edges are >95% of ArtGraph's destination rows and their
// cost is latency, not bandwidth.  Every lane of the warp calls this (shuffles); `row_ok` = the
// lane's row exists and has <= NL edges, `sl` = lane index inside its group of NL.
template <int NL, bool BWD>
__device__ __forceinline__ void softmax_small(const agx_gat_rel_t& R, int row, int beg, int deg,
                                              bool row_ok, int sl, float slope) {
    const bool act = row_ok && sl < deg;
    const int e = beg + sl;
    const float ar = row_ok ? __ldg(R.a_r + row) : 0.f;
    const int c = act ? __ldg(R.col + e) : 0;
    const float raw = act ? __ldg(R.a_l + c) + ar : 0.f;
    if constexpr (!BWD) {
        const float l = act ? leaky(raw, slope) : -INFINITY;
        float m = l;
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float p = act ? expf(l - m) : 0.f;
        float s = p;
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (act) R.alpha[e] = p * (1.0f / (s + 1e-16f));
    } else {
        const float al = act ? R.alpha[e] : 0.f;
        const float dal = act ? __ldg(R.dalpha + e) : 0.f;
        float s = al * dal;
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float de = al * (dal - s) * (raw > 0.f ? 1.0f : slope);      // 0 for idle lanes
        if (act) R.de[e] = de;
        float d = de;
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (row_ok && sl == 0) R.da_r[row] = d;
    }
}

constexpr int kGatGroup = 8;                              // lanes per short row
constexpr int kGatRowsPerWarp = 32 / kGatGroup;           // 4

// The CTAs behind the hub CTAs own 128 consecutive destination rows of one relation each, a warp 4:
//   deg <= 8             8 lanes per row, four rows of the warp at once (registers)
//   8 < deg <= 32        the whole warp, one edge per lane (registers)
//   32 < deg <= LONG     the whole warp, strided three-pass walk
//   deg > LONG           skipped: row long_rows[b] belongs to hub CTA hub_start[rel] + b (all 1024
//                        threads on one row; style / genre / tag hubs of 10^3..10^4 edges)
template <bool BWD>
__global__ void __launch_bounds__(kGatThreads) gat_edge_softmax(const __grid_constant__ GatRels P) {
    __shared__ float s_red[kGatWarps];
    // hub CTAs are the FIRST blocks of the grid: the longest rows start at once and finish behind
    // the short rows instead of after them (they were the tail: 56 us of SM-active time in an 87 us
    // launch)
    if ((int)blockIdx.x < P.hub_start[P.n]) {                    // hub CTA (CTA-uniform branch)
        int ri = 0;
        while ((int)blockIdx.x >= P.hub_start[ri + 1]) ++ri;
        const agx_gat_rel_t& R = P.r[ri];
        const int lrow = __ldg(R.long_rows + ((int)blockIdx.x - P.hub_start[ri]));
        const int lbeg = __ldg(R.rowptr + lrow), lend = __ldg(R.rowptr + lrow + 1);
        if constexpr (BWD)
            softmax_bwd_row<kGatThreads>(R, lrow, lbeg, lend, threadIdx.x, P.slope, s_red);
        else
            softmax_row<kGatThreads>(R, lrow, lbeg, lend, threadIdx.x, P.slope, s_red);
        return;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane / kGatGroup, gl = lane % kGatGroup;
    int ri = 0;
    while ((int)blockIdx.x >= P.blk_start[ri + 1]) ++ri;
    const agx_gat_rel_t& R = P.r[ri];
    const int row0 = (((int)blockIdx.x - P.blk_start[ri]) * kGatWarps + w) * kGatRowsPerWarp;
    const int row = row0 + g;
    int beg = 0, deg = 0;
    if (row < R.n_rows) {
        beg = __ldg(R.rowptr + row);
        deg = __ldg(R.rowptr + row + 1) - beg;
    }
    softmax_small<kGatGroup, BWD>(R, row, beg, deg, row < R.n_rows && deg <= kGatGroup, gl, P.slope);
#pragma unroll
    for (int gg = 0; gg < kGatRowsPerWarp; ++gg) {       // warp-uniform: values of group gg's row
        const int rdeg = __shfl_sync(0xffffffffu, deg, gg * kGatGroup);
        const int rbeg = __shfl_sync(0xffffffffu, beg, gg * kGatGroup);
        const int rrow = row0 + gg;
        // (without a long_rows list the warp walks a hub row itself: slow, still correct)
        if (rdeg <= kGatGroup || (rdeg > AGX_GAT_LONG_ROW && R.n_long > 0)) continue;
        if (rdeg <= 32) {
            softmax_small<32, BWD>(R, rrow, rbeg, rdeg, true, lane, P.slope);
        } else if constexpr (BWD) {
            softmax_bwd_row<32>(R, rrow, rbeg, rbeg + rdeg, lane, P.slope, s_red);
        } else {
            softmax_row<32>(R, rrow, rbeg, rbeg + rdeg, lane, P.slope, s_red);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sddmm: out[e] = <a[row[e]], b[col[e]]>.  A warp owns 32 consecutive slots; LPR lanes cover one
// pair of feature rows (VEC columns per lane and step), the 32 / LPR sub-warps take different
// slots, kSddmmDepth slots per sub-warp are loaded before the first is reduced.
// ------------------------------------------------------------------------------------------------
constexpr int kSddmmDepth = 4;

struct SddmmSegs {
    agx_sddmm_seg_t s[AGX_MAX_SDDMM_SEGS];
    int32_t blk_start[AGX_MAX_SDDMM_SEGS + 1];     // CTA -> segment (256 slots per CTA)
    int32_t n;
    int32_t F;
};

template <int VEC, int LPR>
__global__ void __launch_bounds__(kSddmmThreads) sddmm(const __grid_constant__ SddmmSegs P) {
    constexpr int SUB = 32 / LPR;
    constexpr int U = kSddmmDepth;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    int si = 0;
    while ((int)blockIdx.x >= P.blk_start[si + 1]) ++si;
    const agx_sddmm_seg_t& S = P.s[si];
    const int e0 = (((int)blockIdx.x - P.blk_start[si]) * kSddmmWarps + w) * 32;
    const int n = min(32, S.n_edges - e0);
    if (n <= 0) return;
    const int F = P.F;
    const int my_r = lane < n ? __ldg(S.row + e0 + lane) : 0;
    const int my_c = lane < n ? __ldg(S.col + e0 + lane) : 0;
    for (int j0 = 0; j0 < n; j0 += SUB * U) {
        float part[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(j0 + u * SUB + sub, n - 1);          // clamped: valid rows, not stored
            const int rj = __shfl_sync(0xffffffffu, my_r, j);
            const int cj = __shfl_sync(0xffffffffu, my_c, j);
            const float* pa = S.a + (int64_t)rj * S.lda;
            const float* pb = S.b + (int64_t)cj * S.ldb;
            float acc = 0.f;
            for (int c0 = l * VEC; c0 < F; c0 += LPR * VEC) {
                if constexpr (VEC == 4) {
                    const float4 x = ldg_f4(pa + c0), y = ldg_f4(pb + c0);
                    acc = fmaf(x.x, y.x, acc);
                    acc = fmaf(x.y, y.y, acc);
                    acc = fmaf(x.z, y.z, acc);
                    acc = fmaf(x.w, y.w, acc);
                } else {
                    acc = fmaf(__ldg(pa + c0), __ldg(pb + c0), acc);
                }
            }
            part[u] = acc;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v = part[u];
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int j = j0 + u * SUB + sub;
            if (l == 0 && j < n) S.out[e0 + j] = v;
        }
    }
}

static bool gat_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool BWD>
static int launch_edge_softmax(const agx_gat_rel_t* h_rels, int n_rels, float slope, void* stream,
                               const char* what) {
    AGX_CHECK_ARG(h_rels && n_rels >= 1 && n_rels <= AGX_MAX_GAT_RELS, "%s: n_rels=%d out of [1,%d]",
                  what, n_rels, AGX_MAX_GAT_RELS);
    GatRels P;
    P.n = n_rels;
    P.slope = slope;
    P.hub_start[0] = 0;
    for (int i = 0; i < n_rels; ++i) P.hub_start[i + 1] = P.hub_start[i] + h_rels[i].n_long;
    P.blk_start[0] = P.hub_start[n_rels];
    for (int i = 0; i < n_rels; ++i) {
        const agx_gat_rel_t& R = h_rels[i];
        AGX_CHECK_ARG(R.n_rows >= 0 && R.n_long >= 0 && R.n_long <= R.n_rows,
                      "%s: relation %d: n_rows=%d n_long=%d", what, i, R.n_rows, R.n_long);
        AGX_CHECK_ARG(R.n_rows == 0 || (R.rowptr && R.col && R.a_l && R.a_r && R.alpha),
                      "%s: relation %d: null pointer", what, i);
        AGX_CHECK_ARG(R.n_long == 0 || R.long_rows, "%s: relation %d: null long_rows", what, i);
        AGX_CHECK_ARG(!BWD || R.n_rows == 0 || (R.dalpha && R.de && R.da_r),
                      "%s: relation %d: null backward pointer", what, i);
        P.r[i] = R;
        P.blk_start[i + 1] =
            P.blk_start[i] + (int32_t)ceil_div(R.n_rows, kGatWarps * kGatRowsPerWarp);
    }
    if (P.blk_start[n_rels] == 0) return AGX_OK;
    gat_edge_softmax<BWD><<<(unsigned)P.blk_start[n_rels], kGatThreads, 0, (cudaStream_t)stream>>>(P);
    AGX_LAUNCH_CHECK(what);
    return AGX_OK;
}

}  // namespace agx

using namespace agx;

extern "C" int agx_gat_edge_softmax(const agx_gat_rel_t* h_rels, int n_rels, float slope,
                                    void* stream) {
    return launch_edge_softmax<false>(h_rels, n_rels, slope, stream, "agx_gat_edge_softmax");
}

extern "C" int agx_gat_edge_softmax_bwd(const agx_gat_rel_t* h_rels, int n_rels, float slope,
                                        void* stream) {
    return launch_edge_softmax<true>(h_rels, n_rels, slope, stream, "agx_gat_edge_softmax_bwd");
}

extern "C" int agx_sddmm(const agx_sddmm_seg_t* h_segs, int n_segs, int32_t F, void* stream) {
    AGX_CHECK_ARG(h_segs && n_segs >= 1 && n_segs <= AGX_MAX_SDDMM_SEGS,
                  "agx_sddmm: n_segs=%d out of [1,%d]", n_segs, AGX_MAX_SDDMM_SEGS);
    AGX_CHECK_ARG(F >= 1, "agx_sddmm: F=%d", F);
    SddmmSegs P;
    P.n = n_segs;
    P.F = F;
    P.blk_start[0] = 0;
    bool vec_ok = (F & 3) == 0;
    for (int i = 0; i < n_segs; ++i) {
        const agx_sddmm_seg_t& S = h_segs[i];
        AGX_CHECK_ARG(S.n_edges >= 0, "agx_sddmm: segment %d: n_edges=%d", i, S.n_edges);
        AGX_CHECK_ARG(S.n_edges == 0 || (S.row && S.col && S.a && S.b && S.out),
                      "agx_sddmm: segment %d: null pointer", i);
        vec_ok = vec_ok && gat_aligned16(S.a) && gat_aligned16(S.b) && (S.lda & 3) == 0 &&
                 (S.ldb & 3) == 0;
        P.s[i] = S;
        P.blk_start[i + 1] = P.blk_start[i] + (int32_t)ceil_div(S.n_edges, kSddmmThreads);
    }
    const unsigned grid = (unsigned)P.blk_start[n_segs];
    if (grid == 0) return AGX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!vec_ok)
        sddmm<1, 32><<<grid, kSddmmThreads, 0, st>>>(P);
    else if (F > 64)
        sddmm<4, 32><<<grid, kSddmmThreads, 0, st>>>(P);
    else if (F > 32)
        sddmm<4, 16><<<grid, kSddmmThreads, 0, st>>>(P);
    else
        sddmm<4, 8><<<grid, kSddmmThreads, 0, st>>>(P);
    AGX_LAUNCH_CHECK("sddmm");
    return AGX_OK;
}
