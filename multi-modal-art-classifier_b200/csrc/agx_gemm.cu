// K4: grouped, multi-segment float32 GEMM on the FP32 pipes (FFMA), float32 accumulation.
//
//   C_p[M,N] (+)= sum_s opA_s[M,K_s] * opB_s[K_s,N] (+ bias) (/ row_scale)      for every problem p
//
// One launch covers every dense transform of a hetero layer: the per-relation lin_l products, the
// root (lin_r) product and the cross-relation sum are segments of ONE output tile, so the
// per-relation [N_dst, out] results of the reference are never written to HBM.  Operands are
// addressed by (row stride, column stride), which expresses X*W^T (forward), dY*W (input
// gradient) and dY^T*X (weight gradient, split over the long row dimension into slabs that are
// combined in slab order -- deterministic, no float atomics).
//
// This is the exact-fp32 path required by the 1e-5 parity bound (TF32 tensor-core products are
// 1e-3).  128x128x16 (or 128x32x16) CTA tiles, 8x8 (4x4... ) register micro-tiles, 128-bit shared
// memory reads, register-staged global prefetch of the next k-tile.
#include <stdlib.h>

#include "agx_common.cuh"

namespace agx {

// agx_gemm_tc.cu: tcgen05 path for X W^T problems with many rows
bool gemm_tc_eligible(const agx_gemm_problem_t& Q, const agx_gemm_seg_t* segs);
int gemm_tc_launch(const agx_gemm_problem_t* probs, const int* idx, int cnt,
                   const agx_gemm_seg_t* segs, cudaStream_t st);
// agx_gemm_tc.cu: tcgen05 path for A^T B with a very long K (weight gradients), split over CTAs
bool gemm_tc_longk_eligible(const agx_gemm_problem_t& Q, const agx_gemm_seg_t* segs);
int gemm_tc_longk_launch(const agx_gemm_problem_t* probs, const int* idx, int cnt,
                         const agx_gemm_seg_t* segs, int* used_split, cudaStream_t st);

constexpr int kGemmThreads = 256;
constexpr int BK = 16;

struct GemmParams {
    agx_gemm_problem_t p[AGX_MAX_GEMM_PROBLEMS];
    int32_t tile_start[AGX_MAX_GEMM_PROBLEMS + 1];
    agx_gemm_seg_t s[AGX_MAX_GEMM_SEGS];
    int32_t n;
};

// Load a [ROWS x BK] operand tile into smem laid out sm[k][row] (row stride LDS).
// element (row, k) = g[row * rs + k * cs]; rows >= rows_valid and k >= k_valid are zero-filled.
template <int ROWS, int LDS>
struct TileLoader {
    static constexpr int kVecs = ROWS * BK / 4;                               // float4 per tile
    static constexpr int kIters = (kVecs + kGemmThreads - 1) / kGemmThreads;  // float4 per thread
    float reg[kIters * 4];

    static __device__ __forceinline__ int mode(const float* g, const float* mk, int64_t rs, int64_t cs) {
        const bool al = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mk)) & 15) == 0;
        if (cs == 1 && (rs & 3) == 0 && al) return 0;     // k contiguous, float4 along k
        if (rs == 1 && (cs & 3) == 0 && al) return 1;     // row contiguous, float4 along rows
        return 2;                                         // generic strides, scalar
    }

    // mk (nullable) is an elementwise multiplier with the same addressing as g
    __device__ __forceinline__ void fetch(const float* __restrict__ g, const float* __restrict__ mk,
                                          int64_t rs, int64_t cs, int rows_valid, int k_valid, int tid) {
        const int md = mode(g, mk, rs, cs);
        const int64_t mo = mk ? (mk - g) : 0;
        if (md == 0) {
#pragma unroll
            for (int j = 0; j < kIters; ++j) {
                const int i = tid + j * kGemmThreads;
                const int row = i >> 2, k4 = (i & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < kVecs && row < rows_valid) {
                    const float* p = g + (int64_t)row * rs + k4;
                    if (k4 + 3 < k_valid) {
                        v = __ldg(reinterpret_cast<const float4*>(p));
                        if (mk) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(p + mo));
                            v.x *= q.x; v.y *= q.y; v.z *= q.z; v.w *= q.w;
                        }
                    } else {
                        if (k4 + 0 < k_valid) v.x = __ldg(p + 0) * (mk ? __ldg(p + mo + 0) : 1.f);
                        if (k4 + 1 < k_valid) v.y = __ldg(p + 1) * (mk ? __ldg(p + mo + 1) : 1.f);
                        if (k4 + 2 < k_valid) v.z = __ldg(p + 2) * (mk ? __ldg(p + mo + 2) : 1.f);
                    }
                }
                reg[4 * j + 0] = v.x; reg[4 * j + 1] = v.y; reg[4 * j + 2] = v.z; reg[4 * j + 3] = v.w;
            }
        } else if (md == 1) {
#pragma unroll
            for (int j = 0; j < kIters; ++j) {
                const int i = tid + j * kGemmThreads;
                const int k = i / (ROWS / 4), r4 = (i % (ROWS / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < kVecs && k < k_valid) {
                    const float* p = g + (int64_t)k * cs + r4;
                    if (r4 + 3 < rows_valid) {
                        v = __ldg(reinterpret_cast<const float4*>(p));
                        if (mk) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(p + mo));
                            v.x *= q.x; v.y *= q.y; v.z *= q.z; v.w *= q.w;
                        }
                    } else {
                        if (r4 + 0 < rows_valid) v.x = __ldg(p + 0) * (mk ? __ldg(p + mo + 0) : 1.f);
                        if (r4 + 1 < rows_valid) v.y = __ldg(p + 1) * (mk ? __ldg(p + mo + 1) : 1.f);
                        if (r4 + 2 < rows_valid) v.z = __ldg(p + 2) * (mk ? __ldg(p + mo + 2) : 1.f);
                    }
                }
                reg[4 * j + 0] = v.x; reg[4 * j + 1] = v.y; reg[4 * j + 2] = v.z; reg[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kIters * 4; ++j) {
                const int i = tid + j * kGemmThreads;
                int row, k;
                if (cs == 1) { row = i / BK; k = i % BK; } else { k = i / ROWS; row = i % ROWS; }
                const int64_t off = (int64_t)row * rs + (int64_t)k * cs;
                reg[j] = (i < ROWS * BK && row < rows_valid && k < k_valid)
                             ? __ldg(g + off) * (mk ? __ldg(g + mo + off) : 1.f) : 0.f;
            }
        }
    }

    __device__ __forceinline__ void commit(float* sm, int64_t rs, int64_t cs, const float* g,
                                           const float* mk, int tid) {
        const int md = mode(g, mk, rs, cs);
        if (md == 0) {
#pragma unroll
            for (int j = 0; j < kIters; ++j) {
                const int i = tid + j * kGemmThreads;
                const int row = i >> 2, k4 = (i & 3) * 4;
                if (i < kVecs) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) sm[(k4 + q) * LDS + row] = reg[4 * j + q];
                }
            }
        } else if (md == 1) {
#pragma unroll
            for (int j = 0; j < kIters; ++j) {
                const int i = tid + j * kGemmThreads;
                const int k = i / (ROWS / 4), r4 = (i % (ROWS / 4)) * 4;
                if (i < kVecs)
                    *reinterpret_cast<float4*>(&sm[k * LDS + r4]) =
                        make_float4(reg[4 * j], reg[4 * j + 1], reg[4 * j + 2], reg[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kIters * 4; ++j) {
                const int i = tid + j * kGemmThreads;
                int row, k;
                if (cs == 1) { row = i / BK; k = i % BK; } else { k = i / ROWS; row = i % ROWS; }
                if (i < ROWS * BK) sm[k * LDS + row] = reg[j];
            }
        }
    }
};

// BM x BN CTA tile; each thread owns TM x TN outputs split in two halves per dimension
// (rows ty*TM/2 + {0..TM/2-1} and BM/2 + ty*TM/2 + ..) so that float4 smem reads are conflict-free.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_f32(const __grid_constant__ GemmParams P) {
    static_assert((BM / TM) * (BN / TN) == kGemmThreads, "thread tiling");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[BK * LDA];
    __shared__ __align__(16) float Bs[BK * LDB];

    int pi = 0;
    while ((int)blockIdx.x >= P.tile_start[pi + 1]) ++pi;
    const agx_gemm_problem_t& Q = P.p[pi];
    if (Q.skip_flag && *Q.skip_flag != 0) return;
    const int tiles_n = (Q.N + BN - 1) / BN;
    const int tiles_m = (Q.M + BM - 1) / BM;
    int t = blockIdx.x - P.tile_start[pi];
    const int slab = t / (tiles_m * tiles_n);
    t -= slab * tiles_m * tiles_n;
    const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int mv = min(BM, Q.M - m0), nv = min(BN, Q.N - n0);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    TileLoader<BM, LDA> la;
    TileLoader<BN, LDB> lb;

    for (int si = 0; si < Q.seg_count; ++si) {
        const agx_gemm_seg_t& S = P.s[Q.seg_begin + si];
        int k_begin = 0, k_end = S.K;
        if (Q.split_k > 1) {
            const int per = (((S.K + Q.split_k - 1) / Q.split_k) + BK - 1) / BK * BK;
            k_begin = min(S.K, slab * per);
            k_end = min(S.K, k_begin + per);
        }
        const float* Ag = S.A + (int64_t)m0 * S.a_rs;
        const float* Bg = S.B + (int64_t)n0 * S.b_cs;
        const float* Am = S.A_mask ? S.A_mask + (int64_t)m0 * S.a_rs : nullptr;
        const float* Bm = S.B_mask ? S.B_mask + (int64_t)n0 * S.b_cs : nullptr;
#define AGX_AOFF(k) ((int64_t)(k) * S.a_cs)
#define AGX_BOFF(k) ((int64_t)(k) * S.b_rs)
        if (k_begin < k_end) {
            la.fetch(Ag + AGX_AOFF(k_begin), Am ? Am + AGX_AOFF(k_begin) : nullptr, S.a_rs, S.a_cs, mv,
                     k_end - k_begin, tid);
            lb.fetch(Bg + AGX_BOFF(k_begin), Bm ? Bm + AGX_BOFF(k_begin) : nullptr, S.b_cs, S.b_rs, nv,
                     k_end - k_begin, tid);
        }
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
            __syncthreads();
            la.commit(As, S.a_rs, S.a_cs, Ag + AGX_AOFF(k0), Am ? Am + AGX_AOFF(k0) : nullptr, tid);
            lb.commit(Bs, S.b_cs, S.b_rs, Bg + AGX_BOFF(k0), Bm ? Bm + AGX_BOFF(k0) : nullptr, tid);
            __syncthreads();
            if (k0 + BK < k_end) {
                la.fetch(Ag + AGX_AOFF(k0 + BK), Am ? Am + AGX_AOFF(k0 + BK) : nullptr, S.a_rs, S.a_cs,
                         mv, k_end - k0 - BK, tid);
                lb.fetch(Bg + AGX_BOFF(k0 + BK), Bm ? Bm + AGX_BOFF(k0 + BK) : nullptr, S.b_cs, S.b_rs,
                         nv, k_end - k0 - BK, tid);
            }
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if constexpr (TM / 2 == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(&As[k * LDA + h * (BM / 2) + ty * 4]);
                        a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < TM / 2; ++i)
                            a[h * (TM / 2) + i] = As[k * LDA + h * (BM / 2) + ty * (TM / 2) + i];
                    }
                    if constexpr (TN / 2 == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(&Bs[k * LDB + h * (BN / 2) + tx * 4]);
                        b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < TN / 2; ++j)
                            b[h * (TN / 2) + j] = Bs[k * LDB + h * (BN / 2) + tx * (TN / 2) + j];
                    }
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }

    // epilogue
    const bool to_partial = Q.split_k > 1;
    float* Cbase = to_partial ? Q.partial + (size_t)slab * Q.M * Q.N : Q.C;
    const int64_t ldc = to_partial ? Q.N : Q.ldc;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / (TM / 2)) * (BM / 2) + ty * (TM / 2) + (i % (TM / 2));
        if (m >= Q.M) continue;
        const float rsd = (!to_partial && Q.row_scale) ? Q.row_scale[m] : 1.0f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j / (TN / 2)) * (BN / 2) + tx * (TN / 2) + (j % (TN / 2));
            if (n >= Q.N) continue;
            float v = acc[i][j];
            if (!to_partial) {
                if (Q.row_scale) v = v / rsd;
                if (Q.bias) v += Q.bias[n];
                if (Q.accumulate) v += Cbase[(int64_t)m * ldc + n];
            }
            Cbase[(int64_t)m * ldc + n] = v;
        }
    }
}

struct ReduceParams {
    agx_gemm_problem_t p[AGX_MAX_GEMM_PROBLEMS];
    int64_t elem_start[AGX_MAX_GEMM_PROBLEMS + 1];
    int32_t n;
};

// 64 consecutive outputs per CTA; thread (g, j) adds slabs g, g+4, ... of output j, the four
// group sums are then added in group order: a fixed order, independent of the grid.
__global__ void __launch_bounds__(256)
gemm_splitk_reduce(const __grid_constant__ ReduceParams P) {
    __shared__ float part[4][64];
    const int64_t total = P.elem_start[P.n];
    const int j = threadIdx.x & 63, g = threadIdx.x >> 6;
    for (int64_t base = (int64_t)blockIdx.x * 64; base < total; base += (int64_t)gridDim.x * 64) {
        const int64_t i = base + j;
        int pi = 0;
        float v = 0.f;
        int64_t e = 0;
        const bool live = i < total;
        if (live) {
            while (i >= P.elem_start[pi + 1]) ++pi;
            const agx_gemm_problem_t& Q = P.p[pi];
            e = i - P.elem_start[pi];
            const size_t mn = (size_t)Q.M * Q.N;
            int s = g;
            for (; s + 12 < Q.split_k; s += 16) {
                const float a = Q.partial[(size_t)s * mn + e], b = Q.partial[(size_t)(s + 4) * mn + e];
                const float c = Q.partial[(size_t)(s + 8) * mn + e], d = Q.partial[(size_t)(s + 12) * mn + e];
                v += a; v += b; v += c; v += d;
            }
            for (; s < Q.split_k; s += 4) v += Q.partial[(size_t)s * mn + e];
        }
        part[g][j] = v;
        __syncthreads();
        if (g == 0 && live) {
            const agx_gemm_problem_t& Q = P.p[pi];
            if (!(Q.skip_flag && *Q.skip_flag != 0)) {
                float t = ((part[0][j] + part[1][j]) + part[2][j]) + part[3][j];
                const int m = (int)(e / Q.N), n = (int)(e % Q.N);
                if (Q.row_scale) t = t / Q.row_scale[m];
                if (Q.bias) t += Q.bias[n];
                float* c = Q.C + (int64_t)m * Q.ldc + n;
                if (Q.accumulate) t += *c;
                *c = t;
            }
        }
        __syncthreads();
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_class(const agx_gemm_problem_t* probs, const int* idx, int cnt,
                        const agx_gemm_seg_t* segs, int n_segs, cudaStream_t st) {
    if (cnt == 0) return AGX_OK;
    GemmParams P;
    P.n = cnt;
    for (int s = 0; s < n_segs; ++s) P.s[s] = segs[s];
    P.tile_start[0] = 0;
    for (int i = 0; i < cnt; ++i) {
        const agx_gemm_problem_t& Q = probs[idx[i]];
        P.p[i] = Q;
        const int64_t tiles = ceil_div(Q.M, BM) * ceil_div(Q.N, BN) * (Q.split_k > 1 ? Q.split_k : 1);
        P.tile_start[i + 1] = P.tile_start[i] + (int32_t)tiles;
    }
    if (P.tile_start[cnt] == 0) return AGX_OK;
    gemm_f32<BM, BN, TM, TN><<<P.tile_start[cnt], kGemmThreads, 0, st>>>(P);
    AGX_LAUNCH_CHECK("gemm_f32");
    return AGX_OK;
}

}  // namespace agx

using namespace agx;

extern "C" int agx_gemm_grouped(const agx_gemm_problem_t* h_problems, int n_problems,
                                const agx_gemm_seg_t* h_segs, int n_segs, void* stream) {
    AGX_CHECK_ARG(h_problems && n_problems >= 1 && n_problems <= AGX_MAX_GEMM_PROBLEMS,
                  "agx_gemm_grouped: n_problems=%d out of [1,%d]", n_problems,
                  AGX_MAX_GEMM_PROBLEMS);
    AGX_CHECK_ARG(h_segs && n_segs >= 1 && n_segs <= AGX_MAX_GEMM_SEGS,
                  "agx_gemm_grouped: n_segs=%d out of [1,%d]", n_segs, AGX_MAX_GEMM_SEGS);
    cudaStream_t st = (cudaStream_t)stream;
    int wide[AGX_MAX_GEMM_PROBLEMS], narrow[AGX_MAX_GEMM_PROBLEMS], nw = 0, nn = 0;
    int shortm[AGX_MAX_GEMM_PROBLEMS], ns = 0;
    int64_t short_tiles = 0;
    int tc[AGX_MAX_GEMM_PROBLEMS], ntc = 0;
    int lk[AGX_MAX_GEMM_PROBLEMS], nlk = 0, lk_used[AGX_MAX_GEMM_PROBLEMS];
    static const bool use_tc = getenv("AGX_DISABLE_TC") == nullptr;   // A/B switch for tests
    bool any_split = false;
    for (int i = 0; i < n_problems; ++i) {
        const agx_gemm_problem_t& Q = h_problems[i];
        AGX_CHECK_ARG(Q.M >= 0 && Q.N >= 0 && Q.seg_count >= 0 && Q.seg_begin >= 0 &&
                          Q.seg_begin + Q.seg_count <= n_segs,
                      "agx_gemm_grouped: problem %d: bad sizes / segment range", i);
        AGX_CHECK_ARG((Q.M == 0 || Q.N == 0) || Q.C, "agx_gemm_grouped: problem %d: null C", i);
        if (Q.split_k > 1) {
            AGX_CHECK_ARG(Q.seg_count == 1 && Q.partial,
                          "agx_gemm_grouped: problem %d: split_k needs one segment and a partial "
                          "buffer", i);
            any_split = true;
        }
        for (int s = 0; s < Q.seg_count; ++s) {
            const agx_gemm_seg_t& S = h_segs[Q.seg_begin + s];
            AGX_CHECK_ARG(S.K >= 0 && (S.K == 0 || (S.A && S.B)),
                          "agx_gemm_grouped: problem %d segment %d: null operand", i, s);
        }
        if (Q.M == 0 || Q.N == 0) continue;
        if (use_tc && gemm_tc_eligible(Q, h_segs)) {      // tall X W^T products: tcgen05 3xTF32
            tc[ntc++] = i;
            continue;
        }
        if (use_tc && gemm_tc_longk_eligible(Q, h_segs)) { // dOut^T X over 10^5 rows: tcgen05, split-K
            lk[nlk++] = i;
            continue;
        }
        // few-row problems (the small node types, their weight gradients) would sit on one or two
        // SMs with 128-row tiles: 32-row tiles spread them over 4x as many CTAs.  When even that
        // leaves the launch under two CTAs per SM (the full graph: 6..68 CTAs of 32 x 128 per
        // launch at ~25 us, bound by the instruction issue of one CTA per SM) the launch takes
        // 32 x 32 tiles -- a quarter of the FMAs per k-step on four times the CTAs, same
        // summation order.  Launches that already fill the machine (split-K weight gradients of
        // the 16x graph) keep 32 x 128: narrower tiles re-read the long operand four times
        // (measured there: 17.5 ms of GEMM per step with 32 x 32 everywhere, 10.4 with this rule).
        if (Q.N <= 48) narrow[nn++] = i;
        else if (Q.M <= 256) {
            shortm[ns++] = i;
            short_tiles += ceil_div(Q.M, 32) * ceil_div(Q.N, 128) * (Q.split_k > 1 ? Q.split_k : 1);
        } else wide[nw++] = i;
    }
    // The problems of a call are independent (distinct outputs).  The few-row and narrow classes run
    // on a handful of CTAs each, bound by the serial k-loop of one CTA (~20-30 us per launch on the
    // full graph, nine such launches per step): they go to a side stream NEXT TO the tensor-core
    // launches instead of behind them; the split-K reduction joins both.
    SideStream fk;
    const bool big = ntc > 0 || nlk > 0 || nw > 0, small = ns > 0 || nn > 0;
    int rc = fk.fork(st, 0);
    if (rc) return rc;
    if (!(big && small)) {               // nothing to overlap: stay on the caller's stream
        rc = fk.join();
        if (rc) return rc;
        fk.side = st;
    }
    if (short_tiles < 2 * kNumSMs)
        rc = launch_class<32, 32, 2, 2>(h_problems, shortm, ns, h_segs, n_segs, fk.side);
    else
        rc = launch_class<32, 128, 2, 8>(h_problems, shortm, ns, h_segs, n_segs, fk.side);
    if (rc) return rc;
    rc = launch_class<128, 32, 4, 4>(h_problems, narrow, nn, h_segs, n_segs, fk.side);
    if (rc) return rc;
    if (ntc > 0) {
        const int rc_tc = gemm_tc_launch(h_problems, tc, ntc, h_segs, st);
        if (rc_tc) return rc_tc;
    }
    if (nlk > 0) {
        const int rc_lk = gemm_tc_longk_launch(h_problems, lk, nlk, h_segs, lk_used, st);
        if (rc_lk) return rc_lk;
    }
    rc = launch_class<128, 128, 8, 8>(h_problems, wide, nw, h_segs, n_segs, st);
    if (rc) return rc;
    rc = fk.join();
    if (rc) return rc;
    if (any_split) {
        ReduceParams R;
        R.n = 0;
        R.elem_start[0] = 0;
        for (int i = 0; i < n_problems; ++i) {
            const agx_gemm_problem_t& Q = h_problems[i];
            if (Q.split_k > 1 && Q.M > 0 && Q.N > 0) {
                R.p[R.n] = Q;
                for (int k = 0; k < nlk; ++k)               // tensor-core path wrote fewer slabs
                    if (lk[k] == i) R.p[R.n].split_k = lk_used[k];
                R.elem_start[R.n + 1] = R.elem_start[R.n] + (int64_t)Q.M * Q.N;
                ++R.n;
            }
        }
        if (R.n > 0) {
            const int64_t total = R.elem_start[R.n];
            const int grid = (int)(ceil_div(total, 64) < 148 * 8 ? ceil_div(total, 64) : 148 * 8);
            gemm_splitk_reduce<<<grid, 256, 0, st>>>(R);
            AGX_LAUNCH_CHECK("gemm_splitk_reduce");
        }
    }
    return AGX_OK;
}
