// Batched small kernels of the hetero-GNN training step and the heads: per-node-type BatchNorm1d
// (training statistics in two passes with float64 combination, so they match the reference's
// double-accumulating CPU kernel), ReLU + dropout mask, log_softmax + nll, Adam, SmoothL1, and
// utility copies.  All reductions are two-stage (per-CTA partials -> one combining CTA) in a fixed
// order: no float atomics, results are reproducible.
#include "agx_common.cuh"

namespace agx {

// ------------------------------------------------------------------------------------------------
// sum of up to 8 arrays
// ------------------------------------------------------------------------------------------------
struct SumParams {
    agx_sum_desc_t d[AGX_MAX_TENSORS];
    int64_t start[AGX_MAX_TENSORS + 1];
    int32_t n;
};

__global__ void __launch_bounds__(256) sum_arrays(const __grid_constant__ SumParams P) {
    const int64_t total = P.start[P.n];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int di = 0;
        while (i >= P.start[di + 1]) ++di;
        const agx_sum_desc_t& D = P.d[di];
        const int64_t e = i - P.start[di];
        float v = D.in[0][e];
        for (int k = 1; k < D.n_in; ++k) v += D.in[k][e];
        if (D.bias) v += __ldg(D.bias + e % D.bias_F);
        D.out[e] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d over [n_rows, F] per descriptor.  Slabs of kBnSlab rows per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kBnSlab = 128;
constexpr int kBnThreads = 256;

struct BnParams {
    agx_bn_desc_t d[AGX_MAX_GROUPS];
    int32_t slab_start[AGX_MAX_GROUPS + 1];
    int32_t n;
    int32_t F;
    double* ws;           // [total_slabs][2F] float64 partials: sum x, sum x^2
    int32_t training;
    float momentum, eps;
    // multi-GPU (SyncBN): mode 1 = only reduce the slab partials into sums[n][2F]; mode 2 = finalize
    // from sums (already all-reduced across ranks) with the global row counts; mode 0 = fused
    int32_t mode;
    double* sums;
    const double* counts;
};

// One pass over x: per slab the float64 column sums of x and of x*x (a float32 square is exact in
// float64, so var = E[x^2] - mean^2 loses nothing that matters: 53-bit sums of <= 2^24-bit terms).
// part[slab][0..F) = sum x, part[slab][F..2F) = sum x^2.
__global__ void __launch_bounds__(kBnThreads) bn_stats(const __grid_constant__ BnParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.slab_start[di + 1]) ++di;
    const agx_bn_desc_t& D = P.d[di];
    const int slab = blockIdx.x - P.slab_start[di];
    const int r0 = slab * kBnSlab, r1 = min(D.n_rows, r0 + kBnSlab);
    const int F = P.F;
    double* part = P.ws + (size_t)blockIdx.x * 2 * F;
    if ((F & 3) == 0 && F <= 4 * kBnThreads) {
        // 128-bit path: thread owns 4 consecutive columns, row groups stride over t / (F/4)
        __shared__ double red1[4][kBnThreads], red2[4][kBnThreads];
        const int cols4 = F >> 2;
        const int groups = kBnThreads / cols4;
        const int c = (threadIdx.x % cols4) * 4, g = threadIdx.x / cols4;
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        if (g < groups) {
            int r = r0 + g;
            for (; r + 3 * groups < r1; r += 4 * groups) {          // four rows in flight
                float4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    x[u] = *reinterpret_cast<const float4*>(D.x + (int64_t)(r + u * groups) * F + c);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double v0 = x[u].x, v1 = x[u].y, v2 = x[u].z, v3 = x[u].w;
                    s1[0] += v0; s2[0] += v0 * v0;
                    s1[1] += v1; s2[1] += v1 * v1;
                    s1[2] += v2; s2[2] += v2 * v2;
                    s1[3] += v3; s2[3] += v3 * v3;
                }
            }
            for (; r < r1; r += groups) {
                const float4 x = *reinterpret_cast<const float4*>(D.x + (int64_t)r * F + c);
                const double v0 = x.x, v1 = x.y, v2 = x.z, v3 = x.w;
                s1[0] += v0; s2[0] += v0 * v0;
                s1[1] += v1; s2[1] += v1 * v1;
                s1[2] += v2; s2[2] += v2 * v2;
                s1[3] += v3; s2[3] += v3 * v3;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            red1[i][threadIdx.x] = s1[i];
            red2[i][threadIdx.x] = s2[i];
        }
        __syncthreads();
        if (threadIdx.x < cols4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t1 = 0.0, t2 = 0.0;
                for (int gg = 0; gg < groups; ++gg) {
                    t1 += red1[i][gg * cols4 + threadIdx.x];
                    t2 += red2[i][gg * cols4 + threadIdx.x];
                }
                part[4 * threadIdx.x + i] = t1;
                part[F + 4 * threadIdx.x + i] = t2;
            }
        }
        return;
    }
    __shared__ double red[2][kBnThreads];
    for (int c0 = 0; c0 < F; c0 += kBnThreads) {
        const int cols = min(kBnThreads, F - c0);
        const int groups = kBnThreads / cols > 0 ? kBnThreads / cols : 1;   // row groups
        const int c = threadIdx.x % cols, g = threadIdx.x / cols;
        double s1 = 0.0, s2 = 0.0;
        if (g < groups) {
            for (int r = r0 + g; r < r1; r += groups) {
                const double v = (double)D.x[(int64_t)r * F + c0 + c];
                s1 += v;
                s2 += v * v;
            }
        }
        red[0][threadIdx.x] = s1;
        red[1][threadIdx.x] = s2;
        __syncthreads();
        if (threadIdx.x < cols) {
            double t1 = 0.0, t2 = 0.0;
            for (int gg = 0; gg < groups; ++gg) {
                t1 += red[0][gg * cols + threadIdx.x];
                t2 += red[1][gg * cols + threadIdx.x];
            }
            part[c0 + threadIdx.x] = t1;
            part[F + c0 + threadIdx.x] = t2;
        }
        __syncthreads();
    }
}

// Ordered float64 total of per-slab partials for 32 consecutive columns: 1024 threads =
// 32 columns x 32 slab groups; group g adds slabs g, g+32, ... and the 32 group sums are then added
// in group order by the column's first thread.  part[s * stride + col].  Returns the total in the
// threads with threadIdx.x < 32 (column = col0 + threadIdx.x).
__device__ __forceinline__ double slab_total_32x32(const double* __restrict__ part, int slabs,
                                                   size_t stride, int col0, int ncols,
                                                   double (*sm)[33]) {
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    double a = 0.0;
    if (col0 + c < ncols) {
        // four independent partial sums keep four loads in flight (fixed order: still reproducible)
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int s = g;
        for (; s + 96 < slabs; s += 128) {
            a0 += part[(size_t)s * stride + col0 + c];
            a1 += part[(size_t)(s + 32) * stride + col0 + c];
            a2 += part[(size_t)(s + 64) * stride + col0 + c];
            a3 += part[(size_t)(s + 96) * stride + col0 + c];
        }
        for (; s < slabs; s += 32) a0 += part[(size_t)s * stride + col0 + c];
        a = (a0 + a1) + (a2 + a3);
    }
    sm[g][c] = a;
    __syncthreads();
    double t = 0.0;
    if (g == 0)
        for (int k = 0; k < 32; ++k) t += sm[k][c];
    __syncthreads();
    return t;
}

// The same for TWO column blocks `off` apart in every slab (sum x and sum x^2; sum dy and sum dy*xhat)
// in ONE pass: both chains of loads are in flight together, one barrier pair instead of two.  Same
// summation order per total as slab_total_32x32.
__device__ __forceinline__ void slab_total2_32x32(const double* __restrict__ part, int slabs,
                                                  size_t stride, size_t off, int col0, int ncols,
                                                  double (*sm)[33], double (*sm2)[33], double& t1,
                                                  double& t2) {
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    double a = 0.0, b = 0.0;
    if (col0 + c < ncols) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int s = g;
        for (; s + 96 < slabs; s += 128) {
            const double* q = part + (size_t)s * stride + col0 + c;
            a0 += q[0];
            a1 += q[32 * stride];
            a2 += q[64 * stride];
            a3 += q[96 * stride];
            b0 += q[off];
            b1 += q[32 * stride + off];
            b2 += q[64 * stride + off];
            b3 += q[96 * stride + off];
        }
        for (; s < slabs; s += 32) {
            a0 += part[(size_t)s * stride + col0 + c];
            b0 += part[(size_t)s * stride + off + col0 + c];
        }
        a = (a0 + a1) + (a2 + a3);
        b = (b0 + b1) + (b2 + b3);
    }
    sm[g][c] = a;
    sm2[g][c] = b;
    __syncthreads();
    t1 = t2 = 0.0;
    if (g == 0)
        for (int k = 0; k < 32; ++k) {
            t1 += sm[k][c];
            t2 += sm2[k][c];
        }
    __syncthreads();
}

// grid (descriptor, 32-column block); float64 combination of slab partials.
// P.mode: 0 = finalize from the local partials; 1 = only write the local column sums to
// sums[n][2F] (multi-GPU: all-reduced by the caller); 2 = finalize from sums with the global counts
__global__ void __launch_bounds__(1024) bn_finalize(const __grid_constant__ BnParams P) {
    __shared__ double sm[32][33], sm2[32][33];
    const agx_bn_desc_t& D = P.d[blockIdx.x];
    const int s0 = P.slab_start[blockIdx.x], s1 = P.slab_start[blockIdx.x + 1];
    const int F = P.F;
    const int col0 = blockIdx.y * 32;
    const int c = col0 + (int)threadIdx.x;
    if (!P.training) {
        if (threadIdx.x < 32 && c < F) {
            D.save_mean[c] = D.running_mean[c];
            D.save_invstd[c] = (float)(1.0 / sqrt((double)D.running_var[c] + (double)P.eps));
        }
        return;
    }
    double a1 = 0.0, a2 = 0.0;
    if (P.mode == 2) {
        if (threadIdx.x < 32 && c < F) {
            a1 = P.sums[(size_t)blockIdx.x * 2 * F + c];
            a2 = P.sums[(size_t)blockIdx.x * 2 * F + F + c];
        }
    } else {
        const double* part = P.ws + (size_t)s0 * 2 * F;
        slab_total2_32x32(part, s1 - s0, (size_t)2 * F, (size_t)F, col0, F, sm, sm2, a1, a2);
    }
    if (P.mode == 1) {
        if (threadIdx.x < 32 && c < F) {
            P.sums[(size_t)blockIdx.x * 2 * F + c] = a1;
            P.sums[(size_t)blockIdx.x * 2 * F + F + c] = a2;
        }
        return;
    }
    if (threadIdx.x < 32 && c < F) {
        const double n = P.counts ? P.counts[blockIdx.x] : (double)D.n_rows;
        const double mean = a1 / n;
        double var = a2 / n - mean * mean;               // biased (normalisation)
        if (var < 0.0) var = 0.0;
        D.save_mean[c] = (float)mean;
        D.save_invstd[c] = (float)(1.0 / sqrt(var + (double)P.eps));
        if (D.running_mean) {
            const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
            D.running_mean[c] = (float)((1.0 - P.momentum) * D.running_mean[c] + P.momentum * mean);
            D.running_var[c] = (float)((1.0 - P.momentum) * D.running_var[c] +
                                       P.momentum * unbiased);
        }
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
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

// keep / drop decision of agx_dropout_mask for the four elements of Philox block `ctr`
__device__ __forceinline__ void drop_keep4(uint64_t key, uint64_t ctr, float p, float keep, float k[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float u = (float)(c[i] >> 8) * (1.0f / 16777216.0f);   // [0, 1)
        k[i] = u >= p ? keep : 0.f;
    }
}

__global__ void __launch_bounds__(256) bn_apply(const __grid_constant__ BnParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.slab_start[di + 1]) ++di;
    const agx_bn_desc_t& D = P.d[di];
    const int slab = blockIdx.x - P.slab_start[di];
    const int F = P.F;
    const int64_t e0 = (int64_t)slab * kBnSlab * F;
    const int64_t e1 = min((int64_t)D.n_rows * F, e0 + (int64_t)kBnSlab * F);
    const bool philox = D.y_act && !D.dmask && D.drop_seed && D.drop_p > 0.f;
    const uint64_t key = philox ? D.drop_seed[0] : 0, base = philox ? D.drop_seed[1] : 0;
    const float keep = 1.0f / (1.0f - D.drop_p);
    if ((F & 3) == 0) {      // 128-bit path (rows are 16-byte aligned: torch allocations, F % 4 == 0)
        for (int64_t e = e0 + 4 * (int64_t)threadIdx.x; e < e1; e += 4 * (int64_t)blockDim.x) {
            const int c = (int)(e % F);
            const float4 x = *reinterpret_cast<const float4*>(D.x + e);
            const float4 m = *reinterpret_cast<const float4*>(D.save_mean + c);
            const float4 is = *reinterpret_cast<const float4*>(D.save_invstd + c);
            const float4 wv = *reinterpret_cast<const float4*>(D.weight + c);
            const float4 bv = *reinterpret_cast<const float4*>(D.bias + c);
            float4 y;
            y.x = (x.x - m.x) * is.x * wv.x + bv.x;
            y.y = (x.y - m.y) * is.y * wv.y + bv.y;
            y.z = (x.z - m.z) * is.z * wv.z + bv.z;
            y.w = (x.w - m.w) * is.w * wv.w + bv.w;
            if (D.y) *reinterpret_cast<float4*>(D.y + e) = y;
            if (D.y_act) {
                float4 a = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f),
                                       fmaxf(y.w, 0.f));
                if (D.dmask) {
                    const float4 k = *reinterpret_cast<const float4*>(D.dmask + e);
                    a.x *= k.x; a.y *= k.y; a.z *= k.z; a.w *= k.w;
                } else if (philox) {
                    float k[4];
                    drop_keep4(key, base + (uint64_t)((D.drop_offset + e) >> 2), D.drop_p, keep, k);
                    a.x *= k[0]; a.y *= k[1]; a.z *= k[2]; a.w *= k[3];
                }
                *reinterpret_cast<float4*>(D.y_act + e) = a;
            }
        }
        return;
    }
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int c = (int)(e % F);
        const float xh = (D.x[e] - D.save_mean[c]) * D.save_invstd[c];
        const float y = xh * D.weight[c] + D.bias[c];
        if (D.y) D.y[e] = y;
        if (D.y_act) {
            float a = y > 0.f ? y : 0.f;
            if (D.dmask) {
                a *= D.dmask[e];
            } else if (philox) {
                float k[4];
                const int64_t ge = D.drop_offset + e;
                drop_keep4(key, base + (uint64_t)(ge >> 2), D.drop_p, keep, k);
                a *= k[ge & 3];
            }
            D.y_act[e] = a;
        }
    }
}

struct BnBwdParams {
    agx_bn_bwd_desc_t d[AGX_MAX_GROUPS];
    int32_t slab_start[AGX_MAX_GROUPS + 1];
    int32_t n;
    int32_t F;
    double* ws;           // [total_slabs][2F] float64 partials: sum dy, sum dy*xhat
    double* totals;       // [n][2F] column totals (all-reduced across ranks for SyncBN)
    const double* counts; // nullable: global rows per descriptor
    int32_t training;
};

__device__ __forceinline__ float bn_dy_total(const agx_bn_bwd_desc_t& D, int64_t e) {
    float g = D.dy ? D.dy[e] : 0.f;
    if (D.dy_act) {
        float a = D.dy_act[e];
        if (D.dmask) a *= D.dmask[e];
        else if (D.act_scale != 0.f) a *= D.act_scale;
        if (!(D.y[e] > 0.f)) a = 0.f;
        g += a;
    }
    return g;
}

__device__ __forceinline__ float4 bn_dy_total4(const agx_bn_bwd_desc_t& D, int64_t e) {
    float4 g = D.dy ? *reinterpret_cast<const float4*>(D.dy + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (D.dy_act) {
        float4 a = *reinterpret_cast<const float4*>(D.dy_act + e);
        if (D.dmask) {
            const float4 k = *reinterpret_cast<const float4*>(D.dmask + e);
            a.x *= k.x; a.y *= k.y; a.z *= k.z; a.w *= k.w;
        } else {
            const float sc = D.act_scale != 0.f ? D.act_scale : 1.0f;
            a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
        }
        const float4 y = *reinterpret_cast<const float4*>(D.y + e);
        g.x += y.x > 0.f ? a.x : 0.f;
        g.y += y.y > 0.f ? a.y : 0.f;
        g.z += y.z > 0.f ? a.z : 0.f;
        g.w += y.w > 0.f ? a.w : 0.f;
    }
    return g;
}

__global__ void __launch_bounds__(kBnThreads, 4) bn_bwd_reduce(const __grid_constant__ BnBwdParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.slab_start[di + 1]) ++di;
    const agx_bn_bwd_desc_t& D = P.d[di];
    const int slab = blockIdx.x - P.slab_start[di];
    const int r0 = slab * kBnSlab, r1 = min(D.n_rows, r0 + kBnSlab);
    const int F = P.F;
    double* part = P.ws + (size_t)blockIdx.x * 2 * F;
    __shared__ double red0[kBnThreads], red1[kBnThreads];
    if ((F & 3) == 0 && F <= 4 * kBnThreads) {
        __shared__ double r40[4][kBnThreads], r41[4][kBnThreads];
        const int cols4 = F >> 2;
        const int groups = kBnThreads / cols4;
        const int c = (threadIdx.x % cols4) * 4, g = threadIdx.x / cols4;
        double s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0};
        if (g < groups) {
            const float4 m = *reinterpret_cast<const float4*>(D.save_mean + c);
            const float4 is = *reinterpret_cast<const float4*>(D.save_invstd + c);
            int r = r0 + g;
            for (; r + 3 * groups < r1; r += 4 * groups) {       // four rows in flight, same order
                float4 gy[4], x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t e = (int64_t)(r + u * groups) * F + c;
                    gy[u] = bn_dy_total4(D, e);
                    x[u] = *reinterpret_cast<const float4*>(D.x + e);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    s0[0] += (double)gy[u].x; s1[0] += (double)gy[u].x * (double)((x[u].x - m.x) * is.x);
                    s0[1] += (double)gy[u].y; s1[1] += (double)gy[u].y * (double)((x[u].y - m.y) * is.y);
                    s0[2] += (double)gy[u].z; s1[2] += (double)gy[u].z * (double)((x[u].z - m.z) * is.z);
                    s0[3] += (double)gy[u].w; s1[3] += (double)gy[u].w * (double)((x[u].w - m.w) * is.w);
                }
            }
            for (; r < r1; r += groups) {
                const int64_t e = (int64_t)r * F + c;
                const float4 gy = bn_dy_total4(D, e);
                const float4 x = *reinterpret_cast<const float4*>(D.x + e);
                s0[0] += (double)gy.x; s1[0] += (double)gy.x * (double)((x.x - m.x) * is.x);
                s0[1] += (double)gy.y; s1[1] += (double)gy.y * (double)((x.y - m.y) * is.y);
                s0[2] += (double)gy.z; s1[2] += (double)gy.z * (double)((x.z - m.z) * is.z);
                s0[3] += (double)gy.w; s1[3] += (double)gy.w * (double)((x.w - m.w) * is.w);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r40[i][threadIdx.x] = s0[i];
            r41[i][threadIdx.x] = s1[i];
        }
        __syncthreads();
        if (threadIdx.x < cols4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t0 = 0.0, t1 = 0.0;
                for (int gg = 0; gg < groups; ++gg) {
                    t0 += r40[i][gg * cols4 + threadIdx.x];
                    t1 += r41[i][gg * cols4 + threadIdx.x];
                }
                part[4 * threadIdx.x + i] = t0;
                part[F + 4 * threadIdx.x + i] = t1;
            }
        }
        return;
    }
    for (int c0 = 0; c0 < F; c0 += kBnThreads) {
        const int cols = min(kBnThreads, F - c0);
        const int groups = kBnThreads / cols > 0 ? kBnThreads / cols : 1;
        const int c = threadIdx.x % cols, g = threadIdx.x / cols;
        double s0 = 0.0, s1 = 0.0;
        if (g < groups) {
            const float m = D.save_mean[c0 + c], is = D.save_invstd[c0 + c];
            for (int r = r0 + g; r < r1; r += groups) {
                const int64_t e = (int64_t)r * F + c0 + c;
                const float gy = bn_dy_total(D, e);
                s0 += (double)gy;
                s1 += (double)gy * (double)((D.x[e] - m) * is);
            }
        }
        red0[threadIdx.x] = s0;
        red1[threadIdx.x] = s1;
        __syncthreads();
        if (threadIdx.x < cols) {
            double t0 = 0.0, t1 = 0.0;
            for (int gg = 0; gg < groups; ++gg) {
                t0 += red0[gg * cols + threadIdx.x];
                t1 += red1[gg * cols + threadIdx.x];
            }
            part[c0 + threadIdx.x] = t0;
            part[F + c0 + threadIdx.x] = t1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) bn_bwd_finalize(const __grid_constant__ BnBwdParams P) {
    __shared__ double sm[32][33], sm2[32][33];
    const agx_bn_bwd_desc_t& D = P.d[blockIdx.x];
    const int s0 = P.slab_start[blockIdx.x], s1 = P.slab_start[blockIdx.x + 1];
    const int F = P.F;
    const int col0 = blockIdx.y * 32;
    const double* part = P.ws + (size_t)s0 * 2 * F;
    double a0, a1;
    slab_total2_32x32(part, s1 - s0, (size_t)2 * F, (size_t)F, col0, F, sm, sm2, a0, a1);
    const int c = col0 + (int)threadIdx.x;
    if (threadIdx.x < 32 && c < F) {
        P.totals[(size_t)blockIdx.x * 2 * F + c] = a0;
        P.totals[(size_t)blockIdx.x * 2 * F + F + c] = a1;
        if (D.dbias) D.dbias[c] += (float)a0;
        if (D.dweight) D.dweight[c] += (float)a1;
    }
}

constexpr int kBnApplyMaxF = 512;      // per-column coefficients staged in shared memory up to this F

__global__ void __launch_bounds__(256) bn_bwd_apply(const __grid_constant__ BnBwdParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.slab_start[di + 1]) ++di;
    const agx_bn_bwd_desc_t& D = P.d[di];
    if (!D.dx) return;
    const int slab = blockIdx.x - P.slab_start[di];
    const int F = P.F;
    const double* tot = P.totals + (size_t)di * 2 * F;
    const float inv_n = (float)(1.0 / (P.counts ? P.counts[di] : (double)D.n_rows));
    const int64_t e0 = (int64_t)slab * kBnSlab * F;
    const int64_t e1 = min((int64_t)D.n_rows * F, e0 + (int64_t)kBnSlab * F);
    if ((F & 3) == 0 && F <= kBnApplyMaxF) {
        // dx = a*gy - b - xh*c  with per-column a = w*invstd, b = a*sum(dy)/n, c = a*sum(dy*xh)/n
        // (same float operations per element as the scalar path below, in the same order)
        __shared__ __align__(16) float s_is[kBnApplyMaxF], s_m[kBnApplyMaxF], s_w[kBnApplyMaxF],
            s_t0[kBnApplyMaxF], s_t1[kBnApplyMaxF];
        for (int c = threadIdx.x; c < F; c += blockDim.x) {
            s_is[c] = D.save_invstd[c];
            s_m[c] = D.save_mean[c];
            s_w[c] = D.weight[c];
            s_t0[c] = P.training ? (float)tot[c] * inv_n : 0.f;
            s_t1[c] = P.training ? (float)tot[F + c] * inv_n : 0.f;
        }
        __syncthreads();
        for (int64_t e = e0 + 4 * (int64_t)threadIdx.x; e < e1; e += 4 * (int64_t)blockDim.x) {
            const int c = (int)(e % F);
            const float4 gy = bn_dy_total4(D, e);
            const float4 is = *reinterpret_cast<const float4*>(s_is + c);
            const float4 w = *reinterpret_cast<const float4*>(s_w + c);
            float4 dx;
            if (P.training) {
                const float4 x = *reinterpret_cast<const float4*>(D.x + e);
                const float4 m = *reinterpret_cast<const float4*>(s_m + c);
                const float4 t0 = *reinterpret_cast<const float4*>(s_t0 + c);
                const float4 t1 = *reinterpret_cast<const float4*>(s_t1 + c);
                dx.x = w.x * is.x * (gy.x - t0.x - (x.x - m.x) * is.x * t1.x);
                dx.y = w.y * is.y * (gy.y - t0.y - (x.y - m.y) * is.y * t1.y);
                dx.z = w.z * is.z * (gy.z - t0.z - (x.z - m.z) * is.z * t1.z);
                dx.w = w.w * is.w * (gy.w - t0.w - (x.w - m.w) * is.w * t1.w);
            } else {
                dx.x = w.x * is.x * gy.x;
                dx.y = w.y * is.y * gy.y;
                dx.z = w.z * is.z * gy.z;
                dx.w = w.w * is.w * gy.w;
            }
            *reinterpret_cast<float4*>(D.dx + e) = dx;
        }
        return;
    }
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int c = (int)(e % F);
        const float gy = bn_dy_total(D, e);
        const float is = D.save_invstd[c];
        float dx;
        if (P.training) {
            const float xh = (D.x[e] - D.save_mean[c]) * is;
            dx = D.weight[c] * is * (gy - (float)tot[c] * inv_n - xh * (float)tot[F + c] * inv_n);
        } else {
            dx = D.weight[c] * is * gy;
        }
        D.dx[e] = dx;
    }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): partial per 256-row slab, then ordered combine
// ------------------------------------------------------------------------------------------------
struct ColsumParams {
    agx_colsum_desc_t d[AGX_MAX_TENSORS];
    int32_t slab_start[AGX_MAX_TENSORS + 1];
    int32_t part_start[AGX_MAX_TENSORS + 1];     // float offset of the descriptor's partials
    int32_t n;
    double* ws;
};

constexpr int kColsumSlab = 128;

__global__ void __launch_bounds__(256) colsum_partial(const __grid_constant__ ColsumParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.slab_start[di + 1]) ++di;
    const agx_colsum_desc_t& D = P.d[di];
    const int slab = blockIdx.x - P.slab_start[di];
    const int r0 = slab * kColsumSlab, r1 = min(D.n_rows, r0 + kColsumSlab);
    const int F = D.F;
    double* part = P.ws + P.part_start[di] + (size_t)slab * F;
    __shared__ double red[256];
    if ((F & 3) == 0 && F <= 1024 && (D.ldx & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(D.x) & 15) == 0) {
        // 128-bit path: thread owns 4 consecutive columns, four rows in flight
        __shared__ double red4[4][256];
        const int cols4 = F >> 2;
        const int groups = 256 / cols4;
        const int c = (threadIdx.x % cols4) * 4, g = threadIdx.x / cols4;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        if (g < groups) {
            int r = r0 + g;
            for (; r + 3 * groups < r1; r += 4 * groups) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = *reinterpret_cast<const float4*>(D.x + (int64_t)(r + u * groups) * D.ldx + c);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    s[0] += (double)v[u].x; s[1] += (double)v[u].y;
                    s[2] += (double)v[u].z; s[3] += (double)v[u].w;
                }
            }
            for (; r < r1; r += groups) {
                const float4 v = *reinterpret_cast<const float4*>(D.x + (int64_t)r * D.ldx + c);
                s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) red4[i][threadIdx.x] = s[i];
        __syncthreads();
        if (threadIdx.x < cols4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t = 0.0;
                for (int gg = 0; gg < groups; ++gg) t += red4[i][gg * cols4 + threadIdx.x];
                part[4 * threadIdx.x + i] = t;
            }
        }
        return;
    }
    for (int c0 = 0; c0 < F; c0 += 256) {
        const int cols = min(256, F - c0);
        const int groups = 256 / cols > 0 ? 256 / cols : 1;
        const int c = threadIdx.x % cols, g = threadIdx.x / cols;
        double s = 0.0;
        if (g < groups)
            for (int r = r0 + g; r < r1; r += groups) s += (double)D.x[(int64_t)r * D.ldx + c0 + c];
        red[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x < cols) {
            double t = 0.0;
            for (int gg = 0; gg < groups; ++gg) t += red[gg * cols + threadIdx.x];
            part[c0 + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) colsum_final(const __grid_constant__ ColsumParams P) {
    __shared__ double sm[32][33];
    const agx_colsum_desc_t& D = P.d[blockIdx.x];
    const int slabs = P.slab_start[blockIdx.x + 1] - P.slab_start[blockIdx.x];
    const double* part = P.ws + P.part_start[blockIdx.x];
    for (int col0 = blockIdx.y * 32; col0 < D.F; col0 += gridDim.y * 32) {
        const double a = slab_total_32x32(part, slabs, D.F, col0, D.F, sm);
        const int c = col0 + (int)threadIdx.x;
        if (threadIdx.x < 32 && c < D.F) D.out[c] = D.accumulate ? D.out[c] + (float)a : (float)a;
    }
}

// ------------------------------------------------------------------------------------------------
// log_softmax + (weighted) nll ; warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
log_softmax_nll(const float* __restrict__ logits, int64_t ld, int n_rows, int C,
                const int64_t* __restrict__ labels, const float* __restrict__ class_w,
                float* __restrict__ logp, int64_t ldp, float* __restrict__ row_ws) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float* x = logits + (int64_t)row * ld;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    if (logp)
        for (int c = lane; c < C; c += 32) logp[(int64_t)row * ldp + c] = x[c] - lse;
    if (labels && lane == 0) {
        const int64_t y = labels[row];
        const bool ok = y >= 0 && y < C;                 // ignore_index (-100) / invalid -> weight 0
        const float w = ok ? (class_w ? class_w[y] : 1.0f) : 0.f;
        row_ws[2 * (int64_t)row] = ok ? -(x[y] - lse) * w : 0.f;
        row_ws[2 * (int64_t)row + 1] = w;
    }
}

// C <= 32, C % 4 == 0 (the 32 style classes of the GNN's output layer): 8 lanes x float4 per row, four
// rows per warp -- a quarter of the warp instructions per row of the warp-per-row kernel above
__global__ void __launch_bounds__(256)
log_softmax_nll_v4(const float* __restrict__ logits, int64_t ld, int n_rows, int C,
                   const int64_t* __restrict__ labels, const float* __restrict__ class_w,
                   float* __restrict__ logp, int64_t ldp, float* __restrict__ row_ws) {
    const int lane = threadIdx.x & 31, l = lane & 7;
    const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const bool rv = row < n_rows;
    const bool has = rv && l * 4 < C;
    const float* x = logits + (int64_t)(rv ? row : 0) * ld;
    float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (has) v = *reinterpret_cast<const float4*>(x + l * 4);
    float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = has ? (expf(v.x - mx) + expf(v.y - mx)) + (expf(v.z - mx) + expf(v.w - mx)) : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = mx + logf(s);
    if (logp && has)
        *reinterpret_cast<float4*>(logp + (int64_t)row * ldp + l * 4) =
            make_float4(v.x - lse, v.y - lse, v.z - lse, v.w - lse);
    if (labels && rv && l == 0) {
        const int64_t y = labels[row];
        const bool ok = y >= 0 && y < C;
        const float w = ok ? (class_w ? class_w[y] : 1.0f) : 0.f;
        row_ws[2 * (int64_t)row] = ok ? -(x[y] - lse) * w : 0.f;
        row_ws[2 * (int64_t)row + 1] = w;
    }
}

__global__ void __launch_bounds__(256)
log_softmax_nll_bwd_v4(const float* __restrict__ logp, int64_t ldp, int n_rows, int C,
                       const int64_t* __restrict__ labels, const float* __restrict__ class_w,
                       const float* __restrict__ loss_sum, const float* __restrict__ gscale,
                       float coef, const float* __restrict__ dlogp, int64_t lddp,
                       float* __restrict__ dlogits, int64_t ld) {
    const int lane = threadIdx.x & 31, l = lane & 7;
    const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const bool rv = row < n_rows;
    const bool has = rv && l * 4 < C;
    float g = 0.f;
    int y = -1;
    if (labels && rv) {
        const int64_t yy = labels[row];
        const bool ok = yy >= 0 && yy < C;
        const float w = ok ? (class_w ? class_w[yy] : 1.0f) : 0.f;
        g = w * coef * (gscale ? gscale[0] : 1.0f) / loss_sum[1];
        y = ok ? (int)yy : -1;
    }
    float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
    float sd = 0.f;
    if (dlogp) {
        if (has) dl = *reinterpret_cast<const float4*>(dlogp + (int64_t)row * lddp + l * 4);
        sd = (dl.x + dl.y) + (dl.z + dl.w);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    if (!has) return;
    const float4 lp = *reinterpret_cast<const float4*>(logp + (int64_t)row * ldp + l * 4);
    const float pv[4] = {expf(lp.x), expf(lp.y), expf(lp.z), expf(lp.w)};
    const float dv[4] = {dl.x, dl.y, dl.z, dl.w};
    float o4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float d = g * (pv[i] - (l * 4 + i == y ? 1.0f : 0.f));
        if (dlogp) d += dv[i] - pv[i] * sd;
        o4[i] = d;
    }
    *reinterpret_cast<float4*>(dlogits + (int64_t)row * ld + l * 4) =
        make_float4(o4[0], o4[1], o4[2], o4[3]);
}

static bool lsm_v4_ok(int C, int64_t a, int64_t b, int64_t c, const void* p0, const void* p1,
                      const void* p2) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
    return C <= 32 && (C & 3) == 0 && (a & 3) == 0 && (b & 3) == 0 && (c & 3) == 0 && al(p0) &&
           al(p1) && al(p2);
}

// ordered float64 reduction of interleaved pairs: out[0] = sum in[2i], out[1] = sum in[2i+1].
// One thread-block CLUSTER of 8 CTAs (8 SMs pull the row workspace instead of one): CTA r reduces
// the r-th eighth of the pairs (thread t: elements t, t + 1024, ... of its range, four loads in
// flight; then a shared-memory tree), CTA 0 adds the eight CTA totals in rank order out of their
// shared memory (DSMEM).  The order is fixed by (n, 8, 1024) alone: reproducible.
constexpr int kPairCtas = 8;
__global__ void __cluster_dims__(kPairCtas, 1, 1) __launch_bounds__(1024)
reduce_pairs(const float* __restrict__ in, int64_t n, float* __restrict__ out, int accumulate) {
    __shared__ double s0[1024], s1[1024];
    __shared__ double tot[2];
    const unsigned rank = blockIdx.x;                      // grid = one cluster
    const int64_t per = (n + kPairCtas - 1) / kPairCtas;
    const int64_t lo = per * rank, hi = lo + per < n ? lo + per : n;
    const float2* in2 = reinterpret_cast<const float2*>(in);
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    int64_t i = lo + threadIdx.x;
    for (; i + 3 * 1024 < hi; i += 4 * 1024) {
        const float2 v0 = in2[i], v1 = in2[i + 1024], v2 = in2[i + 2048], v3 = in2[i + 3072];
        a0 += (double)v0.x; a1 += (double)v0.y;
        b0 += (double)v1.x; b1 += (double)v1.y;
        c0 += (double)v2.x; c1 += (double)v2.y;
        d0 += (double)v3.x; d1 += (double)v3.y;
    }
    for (; i < hi; i += 1024) {
        const float2 v = in2[i];
        a0 += (double)v.x;
        a1 += (double)v.y;
    }
    s0[threadIdx.x] = (a0 + b0) + (c0 + d0);
    s1[threadIdx.x] = (a1 + b1) + (c1 + d1);
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] += s0[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tot[0] = s0[0];
        tot[1] = s1[0];
    }
    // cluster barrier (release / acquire): every CTA's total is visible through DSMEM
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (rank == 0 && threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(tot);
        for (unsigned r = 0; r < kPairCtas; ++r) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            double v0, v1;
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v0) : "r"(remote));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v1) : "r"(remote + 8));
            t0 += v0;
            t1 += v1;
        }
        out[0] = (accumulate ? out[0] : 0.f) + (float)t0;
        out[1] = (accumulate ? out[1] : 0.f) + (float)t1;
    }
    // nobody leaves while CTA 0 may still read its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__global__ void __launch_bounds__(256)
log_softmax_nll_bwd(const float* __restrict__ logp, int64_t ldp, int n_rows, int C,
                    const int64_t* __restrict__ labels, const float* __restrict__ class_w,
                    const float* __restrict__ loss_sum, const float* __restrict__ gscale,
                    float coef, const float* __restrict__ dlogp, int64_t lddp,
                    float* __restrict__ dlogits, int64_t ld) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float* lp = logp + (int64_t)row * ldp;
    float g = 0.f;
    int64_t y = -1;
    if (labels) {
        y = labels[row];
        const bool ok = y >= 0 && y < C;
        const float w = ok ? (class_w ? class_w[y] : 1.0f) : 0.f;
        g = w * coef * (gscale ? gscale[0] : 1.0f) / loss_sum[1];
        if (!ok) y = -1;
    }
    // optional upstream gradient on log-probabilities: d/dx = dlogp - softmax * sum(dlogp)
    float sd = 0.f;
    if (dlogp) {
        for (int c = lane; c < C; c += 32) sd += dlogp[(int64_t)row * lddp + c];
        sd = warp_sum(sd);
    }
    for (int c = lane; c < C; c += 32) {
        const float p = expf(lp[c]);
        float d = g * (p - (c == y ? 1.0f : 0.f));
        if (dlogp) d += dlogp[(int64_t)row * lddp + c] - p * sd;
        dlogits[(int64_t)row * ld + c] = d;
    }
}

// nll on given log-probabilities (F.nll_loss): thread per row
__global__ void __launch_bounds__(256)
nll_forward(const float* __restrict__ logp, int64_t ld, int n_rows, int C,
            const int64_t* __restrict__ labels, const float* __restrict__ class_w,
            float* __restrict__ row_ws) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        const int64_t y = labels[row];
        const bool ok = y >= 0 && y < C;
        const float w = ok ? (class_w ? class_w[y] : 1.0f) : 0.f;
        row_ws[2 * (int64_t)row] = ok ? -logp[(int64_t)row * ld + y] * w : 0.f;
        row_ws[2 * (int64_t)row + 1] = w;
    }
}

__global__ void __launch_bounds__(256)
nll_backward(int n_rows, int C, const int64_t* __restrict__ labels, const float* __restrict__ class_w,
             const float* __restrict__ loss_sum, const float* __restrict__ gscale, float coef,
             float* __restrict__ dlogp, int64_t ld) {
    const int64_t total = (int64_t)n_rows * C;
    const float gs = coef * (gscale ? gscale[0] : 1.0f) / loss_sum[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / C), c = (int)(i % C);
        const int64_t y = labels[row];
        float v = 0.f;
        if (y == c) v = -gs * (class_w ? class_w[y] : 1.0f);
        dlogp[(int64_t)row * ld + c] = v;
    }
}

__global__ void finish_loss(const float* __restrict__ loss_sum, float coef, float* __restrict__ loss,
                            int accumulate) {
    const float v = coef * loss_sum[0] / loss_sum[1];
    loss[0] = accumulate ? loss[0] + v : v;
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam single-tensor arithmetic, amsgrad=False, maximize=False)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
          float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
          const int32_t* __restrict__ step) {
    const int t = *step;
    const double bc1 = 1.0 - pow((double)b1, (double)t);
    const double bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    auto upd = [&](float& pi, float gi, float& mi_, float& vi_) {
        float grad = gi;
        if (wd != 0.f) grad += wd * pi;
        const float mi = mi_ + (grad - mi_) * (1.0f - b1);             // lerp_
        const float vi = vi_ * b2 + (1.0f - b2) * grad * grad;         // mul_ + addcmul_
        mi_ = mi;
        vi_ = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi - step_size * (mi / denom);
    };
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
          reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
        // 128-bit accesses, several elements per thread (the two pow() above are per thread)
        const int64_t n4 = n >> 2;
        for (int64_t q = tid; q < n4; q += nthr) {
            float4 P4 = reinterpret_cast<float4*>(p)[q];
            const float4 G4 = reinterpret_cast<const float4*>(g)[q];
            float4 M4 = reinterpret_cast<float4*>(m)[q];
            float4 V4 = reinterpret_cast<float4*>(v)[q];
            upd(P4.x, G4.x, M4.x, V4.x);
            upd(P4.y, G4.y, M4.y, V4.y);
            upd(P4.z, G4.z, M4.z, V4.z);
            upd(P4.w, G4.w, M4.w, V4.w);
            reinterpret_cast<float4*>(m)[q] = M4;
            reinterpret_cast<float4*>(v)[q] = V4;
            reinterpret_cast<float4*>(p)[q] = P4;
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthr) upd(p[i], g[i], m[i], v[i]);
}

// ------------------------------------------------------------------------------------------------
// dropout mask: Philox4x32-10, 4 uniforms per counter
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
dropout_mask(float* __restrict__ mask, int64_t n, float p, const uint64_t* __restrict__ seed) {
    const uint64_t key = seed[0], base = seed[1];
    const float keep = 1.0f / (1.0f - p);
    const int64_t n4 = (n + 3) / 4;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4;
         q += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t ctr = base + (uint64_t)q;
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t e = q * 4 + i;
            if (e < n) {
                const float u = (float)(c[i] >> 8) * (1.0f / 16777216.0f);   // [0, 1)
                mask[e] = u >= p ? keep : 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SmoothL1 (beta = 1), mean
// ------------------------------------------------------------------------------------------------
// squared = 0: SmoothL1 (Huber, beta 1); squared = 1: MSE
__global__ void __launch_bounds__(256)
smooth_l1_partial(const float* __restrict__ out, const float* __restrict__ target, int64_t n,
                  float* __restrict__ dout, float* __restrict__ part, int squared) {
    __shared__ float red[256];
    float s = 0.f;
    const float inv_n = 1.0f / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float d = out[i] - target[i];
        const float a = fabsf(d);
        if (squared) {
            s += d * d;
            if (dout) dout[i] = 2.0f * d * inv_n;
        } else {
            s += a < 1.0f ? 0.5f * d * d : a - 0.5f;
            if (dout) dout[i] = (a < 1.0f ? d : (d > 0.f ? 1.0f : -1.0f)) * inv_n;
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256)
smooth_l1_final(const float* __restrict__ part, int nparts, int64_t n, float* __restrict__ loss) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += (double)part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(red[0] / (double)n);
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_f32(float* p, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void __launch_bounds__(256)
scale_mask(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        y[i] = x[i] * mask[i];
}

__global__ void __launch_bounds__(256)
gather_rows(const float* __restrict__ table, int64_t ld, const int64_t* __restrict__ idx, int64_t n,
            int F, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
        const float* s = table + idx[r] * ld;
        for (int c = lane; c < F; c += 32) out[r * ldo + c] = s[c];
    }
}

__global__ void __launch_bounds__(256)
pack_rows(const float* __restrict__ x, int64_t ld, const int32_t* __restrict__ idx, int n, int F,
          float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
        const float* s = x + (int64_t)idx[r] * ld;
        for (int c = lane; c < F; c += 32) out[r * F + c] = s[c];
    }
}

__global__ void __launch_bounds__(256)
unpack_rows_add(float* __restrict__ x, int64_t ld, const int32_t* __restrict__ idx, int n, int F,
                const float* __restrict__ in) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
        float* d = x + (int64_t)idx[r] * ld;
        for (int c = lane; c < F; c += 32) d[c] += in[r * F + c];
    }
}

__global__ void __launch_bounds__(256)
is_identity(const float* __restrict__ x, int64_t ld, int n, int32_t* __restrict__ not_identity) {
    const int64_t total = (int64_t)n * n;
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n), c = (int)(i % n);
        const float v = x[(int64_t)r * ld + c];
        if (v != (r == c ? 1.0f : 0.f)) bad = true;
    }
    if (bad) *not_identity = 1;
}

__global__ void __launch_bounds__(256)
flag_invert(const int32_t* __restrict__ not_identity, int32_t* __restrict__ flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *flag = *not_identity ? 0 : 1;
}

__global__ void __launch_bounds__(256)
transpose(const float* __restrict__ in, int64_t ld_in, int rows, int cols, float* __restrict__ out,
          int64_t ld_out, const int32_t* __restrict__ only_if_flag) {
    if (only_if_flag && *only_if_flag == 0) return;
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = by + j, c = bx + tx;
        tile[j][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = bx + j, r = by + tx;          // out[c][r]
        if (c < cols && r < rows) out[(int64_t)c * ld_out + r] = tile[tx][j];
    }
}

static inline int grid_for(int64_t n, int per_block) {
    const int64_t b = ceil_div(n, per_block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace agx

using namespace agx;

extern "C" int agx_sum_arrays(const agx_sum_desc_t* h_descs, int n, void* stream) {
    AGX_CHECK_ARG(h_descs && n >= 1 && n <= AGX_MAX_TENSORS, "agx_sum_arrays: n=%d", n);
    SumParams P;
    P.n = n;
    P.start[0] = 0;
    for (int i = 0; i < n; ++i) {
        const agx_sum_desc_t& D = h_descs[i];
        AGX_CHECK_ARG(D.n_in >= 1 && D.n_in <= 8 && D.numel >= 0 && (D.numel == 0 || D.out),
                      "agx_sum_arrays: desc %d invalid", i);
        for (int k = 0; k < D.n_in; ++k)
            AGX_CHECK_ARG(D.numel == 0 || D.in[k], "agx_sum_arrays: desc %d input %d null", i, k);
        AGX_CHECK_ARG(!D.bias || D.bias_F >= 1, "agx_sum_arrays: desc %d: bias_F=%lld", i,
                      (long long)D.bias_F);
        P.d[i] = D;
        P.start[i + 1] = P.start[i] + D.numel;
    }
    if (P.start[n] == 0) return AGX_OK;
    sum_arrays<<<grid_for(P.start[n], 256), 256, 0, (cudaStream_t)stream>>>(P);
    AGX_LAUNCH_CHECK("sum_arrays");
    return AGX_OK;
}

extern "C" size_t agx_bn_workspace_floats(int64_t total_rows, int n_descs, int F) {
    const int64_t slabs = ceil_div(total_rows, kBnSlab) + n_descs;
    return 2 * (size_t)(slabs * 2 * F) + 2 * (size_t)n_descs * 2 * F;   // float64 partials + totals
}

// phases (bit mask): 1 = first-moment pass, 2 = second-moment pass, 4 = normalise.  With `sums`
// (multi-GPU SyncBN) each statistics phase ends by writing the per-type float64 column sums to
// sums[n][F]; the caller all-reduces them and the next phase finalises from them with the global
// row `counts`.  Without `sums` the phases run back to back and finalise from local partials.
static int bn_forward_impl(const agx_bn_desc_t* h_descs, int n, int F, int training, float momentum,
                           float eps, float* workspace, size_t workspace_floats, int phases,
                           double* sums, const double* counts, cudaStream_t st) {
    AGX_CHECK_ARG(h_descs && n >= 1 && n <= AGX_MAX_GROUPS, "agx_bn_forward: n=%d", n);
    AGX_CHECK_ARG(F >= 1, "agx_bn_forward: F=%d", F);
    BnParams P;
    P.n = n;
    P.F = F;
    P.ws = reinterpret_cast<double*>(workspace);
    P.mode = 0;
    P.sums = sums;
    P.counts = counts;
    P.training = training;
    P.momentum = momentum;
    P.eps = eps;
    P.slab_start[0] = 0;
    int64_t rows = 0;
    for (int i = 0; i < n; ++i) {
        const agx_bn_desc_t& D = h_descs[i];
        AGX_CHECK_ARG(D.n_rows >= 0 && D.x && (D.y || D.y_act) && D.weight && D.bias &&
                          D.save_mean && D.save_invstd,
                      "agx_bn_forward: desc %d has null pointers", i);
        AGX_CHECK_ARG(!(D.drop_seed && !D.dmask && D.drop_p > 0.f) ||
                          (D.drop_p < 1.f && (D.drop_offset & 3) == 0 && D.drop_offset >= 0),
                      "agx_bn_forward: desc %d: drop_p in [0,1), drop_offset a multiple of 4", i);
        AGX_CHECK_ARG(training || (D.running_mean && D.running_var),
                      "agx_bn_forward: desc %d: eval mode needs running stats", i);
        AGX_CHECK_ARG(!training || D.n_rows > 1 || counts,
                      "agx_bn_forward: desc %d: Expected more than 1 value per channel when "
                      "training, got input size [%d, %d]", i, D.n_rows, F);
        P.d[i] = D;
        P.slab_start[i + 1] = P.slab_start[i] + (int32_t)ceil_div(D.n_rows, kBnSlab);
        rows += D.n_rows;
    }
    if (training && workspace_floats < agx_bn_workspace_floats(rows, n, F)) {
        set_error("agx_bn_forward: workspace too small");
        return AGX_ERR_WORKSPACE;
    }
    const int slabs = P.slab_start[n];
    const dim3 fgrid(n, (F + 31) / 32);
    if (!training) {
        if (phases & 4) {
            bn_finalize<<<fgrid, 1024, 0, st>>>(P);
            AGX_LAUNCH_CHECK("bn_finalize");
        }
    } else {
        if (phases & 1) {
            if (slabs > 0) {
                bn_stats<<<slabs, kBnThreads, 0, st>>>(P);
                AGX_LAUNCH_CHECK("bn_stats");
            }
            if (sums) {                      // hand the local sums to the caller's all-reduce
                P.mode = 1;
                bn_finalize<<<fgrid, 1024, 0, st>>>(P);
                AGX_LAUNCH_CHECK("bn_finalize");
            }
        }
        if (phases & 4) {
            P.mode = sums ? 2 : 0;
            bn_finalize<<<fgrid, 1024, 0, st>>>(P);
            AGX_LAUNCH_CHECK("bn_finalize");
        }
    }
    if ((phases & 4) && slabs > 0) {
        bn_apply<<<slabs, 256, 0, st>>>(P);
        AGX_LAUNCH_CHECK("bn_apply");
    }
    return AGX_OK;
}

extern "C" int agx_bn_forward(const agx_bn_desc_t* h_descs, int n, int F, int training,
                              float momentum, float eps, float* workspace, size_t workspace_floats,
                              void* stream) {
    return bn_forward_impl(h_descs, n, F, training, momentum, eps, workspace, workspace_floats, 7,
                           nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int agx_bn_forward_phase(const agx_bn_desc_t* h_descs, int n, int F, int training,
                                    float momentum, float eps, float* workspace,
                                    size_t workspace_floats, int phases, double* sums,
                                    const double* counts, void* stream) {
    AGX_CHECK_ARG(sums && counts, "agx_bn_forward_phase: sums/counts must not be null");
    return bn_forward_impl(h_descs, n, F, training, momentum, eps, workspace, workspace_floats,
                           phases, sums, counts, (cudaStream_t)stream);
}

// phases: 1 = column totals (sum dy, sum dy*xhat) into `totals` + local dweight/dbias, 2 = dx
static int bn_backward_impl(const agx_bn_bwd_desc_t* h_descs, int n, int F, int training,
                            float* workspace, size_t workspace_floats, int phases, double* totals,
                            const double* counts, cudaStream_t st) {
    AGX_CHECK_ARG(h_descs && n >= 1 && n <= AGX_MAX_GROUPS, "agx_bn_backward: n=%d", n);
    BnBwdParams P;
    P.n = n;
    P.F = F;
    P.training = training;
    P.slab_start[0] = 0;
    int64_t rows = 0;
    for (int i = 0; i < n; ++i) {
        const agx_bn_bwd_desc_t& D = h_descs[i];
        AGX_CHECK_ARG(D.n_rows >= 0 && D.x && D.weight && D.save_mean && D.save_invstd,
                      "agx_bn_backward: desc %d has null pointers", i);
        AGX_CHECK_ARG(!D.dy_act || D.y, "agx_bn_backward: desc %d: dy_act needs y", i);
        P.d[i] = D;
        P.slab_start[i + 1] = P.slab_start[i] + (int32_t)ceil_div(D.n_rows, kBnSlab);
        rows += D.n_rows;
    }
    if (workspace_floats < agx_bn_workspace_floats(rows, n, F)) {
        set_error("agx_bn_backward: workspace too small");
        return AGX_ERR_WORKSPACE;
    }
    const int slabs = P.slab_start[n];
    P.ws = reinterpret_cast<double*>(workspace);
    P.totals = totals ? totals : reinterpret_cast<double*>(workspace) + (size_t)slabs * 2 * F;
    P.counts = counts;
    if (phases & 1) {
        if (slabs > 0) {
            bn_bwd_reduce<<<slabs, kBnThreads, 0, st>>>(P);
            AGX_LAUNCH_CHECK("bn_bwd_reduce");
        }
        bn_bwd_finalize<<<dim3(n, (F + 31) / 32), 1024, 0, st>>>(P);
        AGX_LAUNCH_CHECK("bn_bwd_finalize");
    }
    if ((phases & 2) && slabs > 0) {
        bn_bwd_apply<<<slabs, 256, 0, st>>>(P);
        AGX_LAUNCH_CHECK("bn_bwd_apply");
    }
    return AGX_OK;
}

extern "C" int agx_bn_backward(const agx_bn_bwd_desc_t* h_descs, int n, int F, int training,
                               float* workspace, size_t workspace_floats, void* stream) {
    return bn_backward_impl(h_descs, n, F, training, workspace, workspace_floats, 3, nullptr, nullptr,
                            (cudaStream_t)stream);
}

extern "C" int agx_bn_backward_phase(const agx_bn_bwd_desc_t* h_descs, int n, int F, int training,
                                     float* workspace, size_t workspace_floats, int phases,
                                     double* totals, const double* counts, void* stream) {
    AGX_CHECK_ARG(totals && counts, "agx_bn_backward_phase: totals/counts must not be null");
    return bn_backward_impl(h_descs, n, F, training, workspace, workspace_floats, phases, totals,
                            counts, (cudaStream_t)stream);
}

extern "C" size_t agx_colsum_workspace_floats(int64_t total_rows, int n_descs, int max_F) {
    return 2 * (size_t)(ceil_div(total_rows, kColsumSlab) + n_descs) * (size_t)max_F;   // float64
}

extern "C" int agx_colsum(const agx_colsum_desc_t* h_descs, int n, float* workspace,
                          size_t workspace_floats, void* stream) {
    AGX_CHECK_ARG(h_descs && n >= 1 && n <= AGX_MAX_TENSORS, "agx_colsum: n=%d", n);
    ColsumParams P;
    P.n = n;
    P.ws = reinterpret_cast<double*>(workspace);
    P.slab_start[0] = 0;
    P.part_start[0] = 0;
    for (int i = 0; i < n; ++i) {
        AGX_CHECK_ARG(h_descs[i].out && (h_descs[i].x || h_descs[i].n_rows == 0) && h_descs[i].F >= 1,
                      "agx_colsum: desc %d invalid", i);
        P.d[i] = h_descs[i];
        const int slabs = (int)ceil_div(h_descs[i].n_rows, kColsumSlab);
        P.slab_start[i + 1] = P.slab_start[i] + slabs;
        P.part_start[i + 1] = P.part_start[i] + slabs * h_descs[i].F;
    }
    if (2 * (size_t)P.part_start[n] > workspace_floats || (P.part_start[n] > 0 && !workspace)) {
        set_error("agx_colsum: workspace too small (%d floats needed)", P.part_start[n]);
        return AGX_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (P.slab_start[n] > 0) {
        colsum_partial<<<P.slab_start[n], 256, 0, st>>>(P);
        AGX_LAUNCH_CHECK("colsum_partial");
    }
    colsum_final<<<dim3(n, 4), 1024, 0, st>>>(P);
    AGX_LAUNCH_CHECK("colsum_final");
    return AGX_OK;
}

extern "C" int agx_log_softmax_nll(const float* logits, int64_t ld, int32_t n_rows, int32_t C,
                                   const int64_t* labels, const float* class_w, float* logp,
                                   int64_t ldp, float* loss_sum, float* row_ws, void* stream) {
    AGX_CHECK_ARG(logits && n_rows >= 0 && C >= 1, "agx_log_softmax_nll: bad arguments");
    AGX_CHECK_ARG(!labels || (loss_sum && row_ws),
                  "agx_log_softmax_nll: labels need loss_sum and row_ws");
    AGX_CHECK_ARG((reinterpret_cast<uintptr_t>(row_ws) % 8) == 0,
                  "agx_log_softmax_nll: row_ws must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows > 0) {
        if (lsm_v4_ok(C, ld, logp ? ldp : 0, 0, logits, logp, nullptr))
            log_softmax_nll_v4<<<(unsigned)ceil_div(n_rows, 32), 256, 0, st>>>(
                logits, ld, n_rows, C, labels, class_w, logp, ldp, row_ws);
        else
            log_softmax_nll<<<(unsigned)ceil_div(n_rows, 8), 256, 0, st>>>(
                logits, ld, n_rows, C, labels, class_w, logp, ldp, row_ws);
        AGX_LAUNCH_CHECK("log_softmax_nll");
    }
    if (labels) {
        reduce_pairs<<<kPairCtas, 1024, 0, st>>>(row_ws, n_rows, loss_sum, 0);
        AGX_LAUNCH_CHECK("reduce_pairs");
    }
    return AGX_OK;
}

extern "C" int agx_nll_forward(const float* logp, int64_t ld, int32_t n_rows, int32_t C,
                               const int64_t* labels, const float* class_w, float* loss_sum,
                               float* row_ws, void* stream) {
    AGX_CHECK_ARG(logp && labels && loss_sum && row_ws && n_rows >= 0 && C >= 1 &&
                      (reinterpret_cast<uintptr_t>(row_ws) % 8) == 0,
                  "agx_nll_forward: bad arguments (row_ws: 8-byte aligned)");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows > 0) {
        nll_forward<<<grid_for(n_rows, 256), 256, 0, st>>>(logp, ld, n_rows, C, labels, class_w,
                                                           row_ws);
        AGX_LAUNCH_CHECK("nll_forward");
    }
    reduce_pairs<<<kPairCtas, 1024, 0, st>>>(row_ws, n_rows, loss_sum, 0);
    AGX_LAUNCH_CHECK("reduce_pairs");
    return AGX_OK;
}

extern "C" int agx_nll_backward(int32_t n_rows, int32_t C, const int64_t* labels,
                                const float* class_w, const float* loss_sum, const float* gscale,
                                float coef, float* dlogp, int64_t ld, void* stream) {
    AGX_CHECK_ARG(labels && loss_sum && dlogp && n_rows >= 0 && C >= 1,
                  "agx_nll_backward: bad arguments");
    if (n_rows == 0) return AGX_OK;
    nll_backward<<<grid_for((int64_t)n_rows * C, 256), 256, 0, (cudaStream_t)stream>>>(
        n_rows, C, labels, class_w, loss_sum, gscale, coef, dlogp, ld);
    AGX_LAUNCH_CHECK("nll_backward");
    return AGX_OK;
}

extern "C" int agx_loss_finish(const float* loss_sum, float coef, float* loss, int accumulate,
                               void* stream) {
    AGX_CHECK_ARG(loss_sum && loss, "agx_loss_finish: null pointer");
    finish_loss<<<1, 1, 0, (cudaStream_t)stream>>>(loss_sum, coef, loss, accumulate);
    AGX_LAUNCH_CHECK("finish_loss");
    return AGX_OK;
}

extern "C" int agx_log_softmax_nll_bwd(const float* logp, int64_t ldp, int32_t n_rows, int32_t C,
                                       const int64_t* labels, const float* class_w,
                                       const float* loss_sum, const float* gscale, float coef,
                                       const float* dlogp, int64_t lddp, float* dlogits, int64_t ld,
                                       void* stream) {
    AGX_CHECK_ARG(logp && dlogits && n_rows >= 0 && C >= 1, "agx_log_softmax_nll_bwd: bad arguments");
    AGX_CHECK_ARG(!labels || loss_sum, "agx_log_softmax_nll_bwd: labels need loss_sum");
    if (n_rows == 0) return AGX_OK;
    if (lsm_v4_ok(C, ldp, ld, dlogp ? lddp : 0, logp, dlogits, dlogp))
        log_softmax_nll_bwd_v4<<<(unsigned)ceil_div(n_rows, 32), 256, 0, (cudaStream_t)stream>>>(
            logp, ldp, n_rows, C, labels, class_w, loss_sum, gscale, coef, dlogp, lddp, dlogits, ld);
    else
        log_softmax_nll_bwd<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(
            logp, ldp, n_rows, C, labels, class_w, loss_sum, gscale, coef, dlogp, lddp, dlogits, ld);
    AGX_LAUNCH_CHECK("log_softmax_nll_bwd");
    return AGX_OK;
}

extern "C" int agx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t numel, float lr, float beta1, float beta2, float eps,
                             float weight_decay, const int32_t* step, void* stream) {
    AGX_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && numel >= 0,
                  "agx_adam_step: null pointer");
    if (numel == 0) return AGX_OK;
    int grid = grid_for((numel + 3) / 4, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    adam_step<<<grid, 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, step);
    AGX_LAUNCH_CHECK("adam_step");
    return AGX_OK;
}

extern "C" int agx_dropout_mask(float* mask, int64_t numel, float p, const uint64_t* seed,
                                void* stream) {
    AGX_CHECK_ARG((mask && seed) || numel == 0, "agx_dropout_mask: null pointer");
    AGX_CHECK_ARG(p >= 0.f && p < 1.f, "agx_dropout_mask: p=%f out of [0,1)", (double)p);
    if (numel <= 0) return AGX_OK;
    dropout_mask<<<grid_for((numel + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(mask, numel, p,
                                                                                  seed);
    AGX_LAUNCH_CHECK("dropout_mask");
    return AGX_OK;
}

extern "C" size_t agx_smooth_l1_workspace_floats(void) { return 1024; }

extern "C" int agx_smooth_l1(const float* out, const float* target, int64_t numel, float* loss,
                             float* dout, float* workspace, void* stream) {
    AGX_CHECK_ARG(out && target && loss && workspace && numel > 0, "agx_smooth_l1: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(numel, 256);
    if (grid > 1024) grid = 1024;
    smooth_l1_partial<<<grid, 256, 0, st>>>(out, target, numel, dout, workspace, 0);
    AGX_LAUNCH_CHECK("smooth_l1_partial");
    smooth_l1_final<<<1, 256, 0, st>>>(workspace, grid, numel, loss);
    AGX_LAUNCH_CHECK("smooth_l1_final");
    return AGX_OK;
}

extern "C" int agx_mse(const float* out, const float* target, int64_t numel, float* loss,
                       float* dout, float* workspace, void* stream) {
    AGX_CHECK_ARG(out && target && loss && workspace && numel > 0, "agx_mse: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(numel, 256);
    if (grid > 1024) grid = 1024;
    smooth_l1_partial<<<grid, 256, 0, st>>>(out, target, numel, dout, workspace, 1);
    AGX_LAUNCH_CHECK("mse_partial");
    smooth_l1_final<<<1, 256, 0, st>>>(workspace, grid, numel, loss);
    AGX_LAUNCH_CHECK("mse_final");
    return AGX_OK;
}

__global__ void __launch_bounds__(256)
tanh_fwd(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        y[i] = tanhf(x[i]);
}

__global__ void __launch_bounds__(256)
tanh_bwd(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * (1.0f - y[i] * y[i]);
}

extern "C" int agx_tanh(const float* x, float* y, int64_t numel, void* stream) {
    AGX_CHECK_ARG((x && y) || numel == 0, "agx_tanh: null pointer");
    if (numel <= 0) return AGX_OK;
    tanh_fwd<<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(x, y, numel);
    AGX_LAUNCH_CHECK("tanh_fwd");
    return AGX_OK;
}

extern "C" int agx_tanh_bwd(const float* y, const float* dy, float* dx, int64_t numel, void* stream) {
    AGX_CHECK_ARG((y && dy && dx) || numel == 0, "agx_tanh_bwd: null pointer");
    if (numel <= 0) return AGX_OK;
    tanh_bwd<<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, numel);
    AGX_LAUNCH_CHECK("tanh_bwd");
    return AGX_OK;
}

// torch.optim.SGD(momentum, dampening 0, no nesterov): buf = momentum*buf + g ; p -= lr*buf
// (buf starts at zero, so the first step uses buf = g exactly like torch's lazy initialisation)
__global__ void __launch_bounds__(256)
sgd_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n,
         float lr, float momentum, float weight_decay) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);
        const float b = momentum * buf[i] + gi;
        buf[i] = b;
        p[i] -= lr * b;
    }
}

extern "C" int agx_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t numel,
                            float lr, float momentum, float weight_decay, void* stream) {
    AGX_CHECK_ARG((param && grad && momentum_buf) || numel == 0, "agx_sgd_step: null pointer");
    if (numel <= 0) return AGX_OK;
    sgd_step<<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, numel,
                                                                    lr, momentum, weight_decay);
    AGX_LAUNCH_CHECK("sgd_step");
    return AGX_OK;
}

extern "C" int agx_fill_f32(float* p, int64_t numel, float v, void* stream) {
    AGX_CHECK_ARG(p || numel == 0, "agx_fill_f32: null pointer");
    if (numel <= 0) return AGX_OK;
    fill_f32<<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(p, numel, v);
    AGX_LAUNCH_CHECK("fill_f32");
    return AGX_OK;
}

extern "C" int agx_scale_mask(const float* x, const float* mask, float* y, int64_t numel,
                              void* stream) {
    AGX_CHECK_ARG((x && mask && y) || numel == 0, "agx_scale_mask: null pointer");
    if (numel <= 0) return AGX_OK;
    scale_mask<<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, y, numel);
    AGX_LAUNCH_CHECK("scale_mask");
    return AGX_OK;
}

extern "C" int agx_gather_rows(const float* table, int64_t ld, const int64_t* idx, int64_t n,
                               int32_t F, float* out, int64_t ldo, void* stream) {
    AGX_CHECK_ARG((table && idx && out) || n == 0, "agx_gather_rows: null pointer");
    if (n <= 0) return AGX_OK;
    gather_rows<<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>(table, ld, idx, n, F, out, ldo);
    AGX_LAUNCH_CHECK("gather_rows");
    return AGX_OK;
}

extern "C" int agx_pack_rows(const float* x, int64_t ld, const int32_t* idx, int32_t n, int32_t F,
                             float* out, void* stream) {
    AGX_CHECK_ARG((x && idx && out) || n == 0, "agx_pack_rows: null pointer");
    if (n <= 0) return AGX_OK;
    pack_rows<<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, idx, n, F, out);
    AGX_LAUNCH_CHECK("pack_rows");
    return AGX_OK;
}

extern "C" int agx_unpack_rows_add(float* x, int64_t ld, const int32_t* idx, int32_t n, int32_t F,
                                   const float* in, void* stream) {
    AGX_CHECK_ARG((x && idx && in) || n == 0, "agx_unpack_rows_add: null pointer");
    if (n <= 0) return AGX_OK;
    unpack_rows_add<<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, idx, n, F, in);
    AGX_LAUNCH_CHECK("unpack_rows_add");
    return AGX_OK;
}

extern "C" int agx_is_identity(const float* x, int64_t ld, int32_t n, int32_t* flag,
                               int32_t* scratch, void* stream) {
    AGX_CHECK_ARG(x && flag && scratch && n >= 1, "agx_is_identity: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    AGX_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int32_t), st));
    is_identity<<<grid_for((int64_t)n * n, 256), 256, 0, st>>>(x, ld, n, scratch);
    AGX_LAUNCH_CHECK("is_identity");
    flag_invert<<<1, 32, 0, st>>>(scratch, flag);
    AGX_LAUNCH_CHECK("flag_invert");
    return AGX_OK;
}

struct TransposeParams {
    agx_transpose_desc_t d[AGX_MAX_TENSORS];
    int32_t tile_start[AGX_MAX_TENSORS + 1];
    int32_t n;
};

// every matrix of the batch in one launch: CTA -> (descriptor, 32x32 tile)
__global__ void __launch_bounds__(256) transpose_batched(const __grid_constant__ TransposeParams P) {
    int di = 0;
    while ((int)blockIdx.x >= P.tile_start[di + 1]) ++di;
    const agx_transpose_desc_t& D = P.d[di];
    const int t = blockIdx.x - P.tile_start[di];
    const int tiles_x = (D.cols + 31) / 32;
    const int bx = (t % tiles_x) * 32, by = (t / tiles_x) * 32;
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = by + j, c = bx + tx;
        tile[j][tx] = (r < D.rows && c < D.cols) ? D.in[(int64_t)r * D.ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = bx + j, r = by + tx;          // out[c][r]
        if (c < D.cols && r < D.rows) {
            float* o = D.out + (int64_t)c * D.ld_out + r;
            *o = D.accumulate ? *o + tile[tx][j] : tile[tx][j];
        }
    }
}

extern "C" int agx_transpose_batched(const agx_transpose_desc_t* h_descs, int n, void* stream) {
    AGX_CHECK_ARG(h_descs && n >= 1 && n <= AGX_MAX_TENSORS, "agx_transpose_batched: n=%d", n);
    TransposeParams P;
    P.n = n;
    P.tile_start[0] = 0;
    for (int i = 0; i < n; ++i) {
        const agx_transpose_desc_t& D = h_descs[i];
        AGX_CHECK_ARG(D.rows >= 0 && D.cols >= 0 && ((D.in && D.out) || D.rows == 0 || D.cols == 0),
                      "agx_transpose_batched: desc %d: bad arguments", i);
        P.d[i] = D;
        P.tile_start[i + 1] = P.tile_start[i] +
                              (int32_t)(ceil_div(D.rows, 32) * ceil_div(D.cols, 32));
    }
    if (P.tile_start[n] == 0) return AGX_OK;
    transpose_batched<<<P.tile_start[n], 256, 0, (cudaStream_t)stream>>>(P);
    AGX_LAUNCH_CHECK("transpose_batched");
    return AGX_OK;
}

extern "C" int agx_transpose(const float* in, int64_t ld_in, int32_t rows, int32_t cols, float* out,
                             int64_t ld_out, const int32_t* only_if_flag, void* stream) {
    AGX_CHECK_ARG(in && out && rows >= 0 && cols >= 0, "agx_transpose: bad arguments");
    if (rows == 0 || cols == 0) return AGX_OK;
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    transpose<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld_in, rows, cols, out, ld_out,
                                                      only_if_flag);
    AGX_LAUNCH_CHECK("transpose");
    return AGX_OK;
}
