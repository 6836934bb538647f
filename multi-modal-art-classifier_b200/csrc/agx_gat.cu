// GATConv attention aggregation (the reference script's default --operator,
// /root/reference/src/train_gnn_embeddings.py:15,99; PyG 2.0.2 GATConv, one head):
//
//   e_ij = leaky_relu(a_l[j] + a_r[i]),  alpha_ij = exp(e_ij - max_i) / (sum_i + 1e-16),
//   out_i = sum_j alpha_ij x_l[j] (+ bias)
//
// over the CSR (by destination) of the relation's edge list with self loops added.  One warp per
// row, neighbours in edge order, float32, no atomics: reproducible.  The backward pass is the
// softmax / leaky-relu chain per destination row plus the transpose (CSC, by source) of the
// weighted sum; per-edge quantities are exchanged between the two in ORIGINAL edge order
// (alpha_e[eid], de_e[eid]) so no inverse permutation is needed.
#include "agx_common.cuh"

namespace agx {

constexpr int kGatMaxF = 256;                 // feature width handled in registers (8 per lane)
constexpr int kGatCols = kGatMaxF / 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : slope * v; }

__global__ void __launch_bounds__(256)
gat_fwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
        const int32_t* __restrict__ eid, const float* __restrict__ a_l,
        const float* __restrict__ a_r, const float* __restrict__ x_l, int64_t ldx, int F,
        float slope, const float* __restrict__ bias, float* __restrict__ out, int64_t ldo,
        float* __restrict__ alpha_e, int n_rows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float ar = __ldg(a_r + row);
    float m = -INFINITY;
    for (int e = beg + lane; e < end; e += 32)
        m = fmaxf(m, leaky(__ldg(a_l + __ldg(col + e)) + ar, slope));
    m = warp_max(m);
    float s = 0.f;
    for (int e = beg + lane; e < end; e += 32)
        s += expf(leaky(__ldg(a_l + __ldg(col + e)) + ar, slope) - m);
    s = warp_sum(s);
    const float inv = 1.0f / (s + 1e-16f);
    float acc[kGatCols];
#pragma unroll
    for (int k = 0; k < kGatCols; ++k) acc[k] = 0.f;
    for (int e0 = beg; e0 < end; e0 += 32) {
        const int n = min(32, end - e0);
        int c = 0;
        float al = 0.f;
        if (lane < n) {
            c = __ldg(col + e0 + lane);
            al = expf(leaky(__ldg(a_l + c) + ar, slope) - m) * inv;
            alpha_e[__ldg(eid + e0 + lane)] = al;
        }
        for (int j = 0; j < n; ++j) {
            const int cj = __shfl_sync(0xffffffffu, c, j);
            const float aj = __shfl_sync(0xffffffffu, al, j);
            const float* xr = x_l + (int64_t)cj * ldx;
#pragma unroll
            for (int k = 0; k < kGatCols; ++k) {
                const int cc = lane + 32 * k;
                if (cc < F) acc[k] = fmaf(aj, __ldg(xr + cc), acc[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kGatCols; ++k) {
        const int cc = lane + 32 * k;
        if (cc < F) out[(int64_t)row * ldo + cc] = acc[k] + (bias ? __ldg(bias + cc) : 0.f);
    }
}

// per destination row: d alpha_ij = dout_i . x_l[j];  s_i = sum_j alpha_ij d alpha_ij;
// de_ij = alpha_ij (d alpha_ij - s_i) * leaky'(a_l[j] + a_r[i]);  da_r[i] = sum_j de_ij
__global__ void __launch_bounds__(256)
gat_bwd_dst(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
            const int32_t* __restrict__ eid, const float* __restrict__ a_l,
            const float* __restrict__ a_r, const float* __restrict__ x_l, int64_t ldx, int F,
            float slope, const float* __restrict__ dout, int64_t ldd,
            const float* __restrict__ alpha_e, float* __restrict__ de_e, float* __restrict__ da_r,
            int n_rows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float ar = __ldg(a_r + row);
    float g[kGatCols];
#pragma unroll
    for (int k = 0; k < kGatCols; ++k) {
        const int cc = lane + 32 * k;
        g[k] = cc < F ? __ldg(dout + (int64_t)row * ldd + cc) : 0.f;
    }
    float s = 0.f;                                   // same value in every lane
    for (int e = beg; e < end; ++e) {
        const int c = __ldg(col + e);
        const float* xr = x_l + (int64_t)c * ldx;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < kGatCols; ++k) {
            const int cc = lane + 32 * k;
            if (cc < F) d = fmaf(g[k], __ldg(xr + cc), d);
        }
        d = warp_sum(d);
        const int id = __ldg(eid + e);
        s = fmaf(alpha_e[id], d, s);
        if (lane == 0) de_e[id] = d;                 // d alpha for now
    }
    __syncwarp();
    float t = 0.f;
    for (int e = beg + lane; e < end; e += 32) {
        const int id = __ldg(eid + e);
        const float raw = __ldg(a_l + __ldg(col + e)) + ar;
        const float de = alpha_e[id] * (de_e[id] - s) * (raw > 0.f ? 1.0f : slope);
        de_e[id] = de;
        t += de;
    }
    t = warp_sum(t);
    if (lane == 0) da_r[row] = t;
}

// per source row (CSC): dx_l[j] = sum_i alpha_ij dout_i ;  da_l[j] = sum_i de_ij
__global__ void __launch_bounds__(256)
gat_bwd_src(const int32_t* __restrict__ cscptr, const int32_t* __restrict__ dstid,
            const int32_t* __restrict__ eid, const float* __restrict__ alpha_e,
            const float* __restrict__ de_e, const float* __restrict__ dout, int64_t ldd, int F,
            float* __restrict__ dx_l, int64_t ldx, float* __restrict__ da_l, int n_src) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_src) return;
    const int beg = __ldg(cscptr + row), end = __ldg(cscptr + row + 1);
    float acc[kGatCols];
#pragma unroll
    for (int k = 0; k < kGatCols; ++k) acc[k] = 0.f;
    float sde = 0.f;
    for (int q = beg; q < end; ++q) {
        const int i = __ldg(dstid + q), id = __ldg(eid + q);
        const float a = alpha_e[id];
        sde += de_e[id];
        const float* gr = dout + (int64_t)i * ldd;
#pragma unroll
        for (int k = 0; k < kGatCols; ++k) {
            const int cc = lane + 32 * k;
            if (cc < F) acc[k] = fmaf(a, __ldg(gr + cc), acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kGatCols; ++k) {
        const int cc = lane + 32 * k;
        if (cc < F) dx_l[(int64_t)row * ldx + cc] = acc[k];
    }
    if (lane == 0) da_l[row] = sde;
}

}  // namespace agx

using namespace agx;

extern "C" int agx_gat_forward(const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                               const float* a_l, const float* a_r, const float* x_l, int64_t ldx,
                               int32_t F, float slope, const float* bias, float* out, int64_t ldo,
                               float* alpha_e, int32_t n_rows, void* stream) {
    AGX_CHECK_ARG(F >= 1 && F <= kGatMaxF, "agx_gat_forward: F=%d out of [1,%d]", F, kGatMaxF);
    AGX_CHECK_ARG(n_rows >= 0, "agx_gat_forward: n_rows=%d", n_rows);
    if (n_rows == 0) return AGX_OK;
    AGX_CHECK_ARG(rowptr && a_l && a_r && x_l && out, "agx_gat_forward: null pointer");
    gat_fwd<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(
        rowptr, col, eid, a_l, a_r, x_l, ldx, F, slope, bias, out, ldo, alpha_e, n_rows);
    AGX_LAUNCH_CHECK("gat_fwd");
    return AGX_OK;
}

extern "C" int agx_gat_backward_dst(const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                                    const float* a_l, const float* a_r, const float* x_l,
                                    int64_t ldx, int32_t F, float slope, const float* dout,
                                    int64_t ldd, const float* alpha_e, float* de_e, float* da_r,
                                    int32_t n_rows, void* stream) {
    AGX_CHECK_ARG(F >= 1 && F <= kGatMaxF, "agx_gat_backward_dst: F=%d out of [1,%d]", F, kGatMaxF);
    if (n_rows <= 0) return AGX_OK;
    AGX_CHECK_ARG(rowptr && a_l && a_r && x_l && dout && da_r, "agx_gat_backward_dst: null pointer");
    gat_bwd_dst<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(
        rowptr, col, eid, a_l, a_r, x_l, ldx, F, slope, dout, ldd, alpha_e, de_e, da_r, n_rows);
    AGX_LAUNCH_CHECK("gat_bwd_dst");
    return AGX_OK;
}

extern "C" int agx_gat_backward_src(const int32_t* cscptr, const int32_t* dstid, const int32_t* eid,
                                    const float* alpha_e, const float* de_e, const float* dout,
                                    int64_t ldd, int32_t F, float* dx_l, int64_t ldx, float* da_l,
                                    int32_t n_src, void* stream) {
    AGX_CHECK_ARG(F >= 1 && F <= kGatMaxF, "agx_gat_backward_src: F=%d out of [1,%d]", F, kGatMaxF);
    if (n_src <= 0) return AGX_OK;
    AGX_CHECK_ARG(cscptr && dout && dx_l && da_l, "agx_gat_backward_src: null pointer");
    gat_bwd_src<<<(unsigned)ceil_div(n_src, 8), 256, 0, (cudaStream_t)stream>>>(
        cscptr, dstid, eid, alpha_e, de_e, dout, ldd, F, dx_l, ldx, da_l, n_src);
    AGX_LAUNCH_CHECK("gat_bwd_src");
    return AGX_OK;
}
