// Layer-level entry points of the C ABI (include/agx.h): the graph plan handle and one SAGEConv /
// GraphConv relation forward + backward, composed from the K1-K4 kernels of this library on the
// host side in C++ -- what a C / C++ caller uses instead of the Python planning layer
// (mmac_b200/functional.py).  The Python operators route their standalone (single-relation) calls
// through these entry points as well (mmac_b200/nn.py: MessagePassing.forward).
//
//   reference: conv((x_src, x_dst), edge_index) at /root/reference/src/models/models_graph.py:30,38
//              and its autograd backward (src/train_gnn_embeddings.py:47).
#include <stdlib.h>
#include <string.h>

#include <new>

#include "agx_common.cuh"

namespace agx {

constexpr double kLongRowAvgDegree = 12.0;     // same policy as mmac_b200/ops.py: CSR.long_rows
constexpr int kLongRowMaxDegree = 256;

__global__ void __launch_bounds__(256)
max_degree_kernel(const int32_t* __restrict__ rowptr, int64_t n_slots, const int32_t* __restrict__ rel_of_slot_start,
                  int n_rels, int32_t* __restrict__ out) {
    // rowptr of relation r occupies slots [start[r], start[r+1]) = n_rows_r + 1 entries
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i + 1 < n_slots;
         i += (int64_t)gridDim.x * blockDim.x) {
        int r = 0;
        while (r + 1 < n_rels && i >= rel_of_slot_start[r + 1]) ++r;
        if (i + 1 >= rel_of_slot_start[r + 1]) continue;           // last entry of relation r
        atomicMax(out + r, rowptr[i + 1] - rowptr[i]);
    }
}

}  // namespace agx

using namespace agx;

struct agx_graph_plan {
    int n_rels;
    void* buf;                  // one device allocation
    agx_plan_rel_t rel[AGX_MAX_CSR_RELS / 2];
};

extern "C" int agx_graph_plan_create(const agx_edge_list_t* h_rels, int n_rels, void* stream,
                                     agx_graph_plan_t** plan_out) {
    AGX_CHECK_ARG(h_rels && plan_out && n_rels >= 1 && 2 * n_rels <= AGX_MAX_CSR_RELS,
                  "agx_graph_plan_create: n_rels=%d out of [1,%d]", n_rels, AGX_MAX_CSR_RELS / 2);
    cudaStream_t st = (cudaStream_t)stream;
    // lists 0..R-1: CSR (key = destination), R..2R-1: CSC (key = source)
    agx_edge_list_t lists[AGX_MAX_CSR_RELS];
    int64_t E = 0, NR = 0;
    for (int r = 0; r < n_rels; ++r) {
        lists[r] = h_rels[r];
        lists[n_rels + r].keys = h_rels[r].vals;
        lists[n_rels + r].vals = h_rels[r].keys;
        lists[n_rels + r].n_edges = h_rels[r].n_edges;
        lists[n_rels + r].n_rows = h_rels[r].n_cols;
        lists[n_rels + r].n_cols = h_rels[r].n_rows;
        E += 2 * h_rels[r].n_edges;
        NR += h_rels[r].n_rows + h_rels[r].n_cols;
    }
    const int L2 = 2 * n_rels;
    const size_t ws_bytes = agx_csr_workspace_bytes(E, NR);
    const size_t b_rowptr = align_up((size_t)(NR + L2) * 4, 256);
    const size_t b_col = align_up((size_t)(E > 0 ? E : 1) * 4, 256);
    const size_t b_cnt = align_up((size_t)(NR > 0 ? NR : 1) * 4, 256);
    const size_t b_misc = 1024;
    const size_t total = b_rowptr + 2 * b_col + b_cnt + b_misc;
    char* base = nullptr;
    AGX_CUDA(cudaMalloc((void**)&base, total));
    void* ws = nullptr;
    cudaError_t e = cudaMalloc(&ws, ws_bytes);
    if (e != cudaSuccess) {
        cudaFree(base);
        return cuda_fail(e, "cudaMalloc(sort workspace)");
    }
    int32_t* rowptr = reinterpret_cast<int32_t*>(base);
    int32_t* col = reinterpret_cast<int32_t*>(base + b_rowptr);
    int32_t* eid = reinterpret_cast<int32_t*>(base + b_rowptr + b_col);
    float* cnt = reinterpret_cast<float*>(base + b_rowptr + 2 * b_col);
    int32_t* misc = reinterpret_cast<int32_t*>(base + b_rowptr + 2 * b_col + b_cnt);   // err, maxdeg, starts
    cudaMemsetAsync(misc, 0, b_misc, st);
    int rc = agx_csr_build(lists, L2, rowptr, col, eid, cnt, misc, ws, ws_bytes, st);
    if (rc == AGX_OK) {
        int32_t h_start[AGX_MAX_CSR_RELS + 1];
        h_start[0] = 0;
        for (int r = 0; r < L2; ++r) h_start[r + 1] = h_start[r] + (int32_t)lists[r].n_rows + 1;
        int32_t* d_start = misc + 64;
        int32_t* d_max = misc + 8;
        cudaMemcpyAsync(d_start, h_start, sizeof(int32_t) * (L2 + 1), cudaMemcpyHostToDevice, st);
        max_degree_kernel<<<148 * 4, 256, 0, st>>>(rowptr, NR + L2, d_start, L2, d_max);
        count_launch();
        int32_t h_misc[8 + AGX_MAX_CSR_RELS];
        e = cudaMemcpyAsync(h_misc, misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "agx_graph_plan_create");
        else if (h_misc[0] != 0) {
            set_error("agx_graph_plan_create: edge list contains node ids outside [0, n)");
            rc = AGX_ERR_INVALID;
        } else {
            agx_graph_plan* P = new (std::nothrow) agx_graph_plan;
            if (!P) {
                set_error("agx_graph_plan_create: out of host memory");
                rc = AGX_ERR_INVALID;
            } else {
                P->n_rels = n_rels;
                P->buf = base;
                int64_t e0 = 0, r0 = 0;
                int64_t eoff[AGX_MAX_CSR_RELS], roff[AGX_MAX_CSR_RELS];
                for (int r = 0; r < L2; ++r) {
                    eoff[r] = e0;
                    roff[r] = r0;
                    e0 += lists[r].n_edges;
                    r0 += lists[r].n_rows;
                }
                for (int r = 0; r < n_rels; ++r) {
                    agx_plan_rel_t& R = P->rel[r];
                    const int t = n_rels + r;
                    R.rowptr = rowptr + roff[r] + r;
                    R.col = col + eoff[r];
                    R.cnt = cnt + roff[r];
                    R.t_rowptr = rowptr + roff[t] + t;
                    R.t_col = col + eoff[t];
                    R.n_src = (int32_t)h_rels[r].n_cols;
                    R.n_dst = (int32_t)h_rels[r].n_rows;
                    R.n_edges = (int32_t)h_rels[r].n_edges;
                    const double avg = R.n_dst > 0 ? (double)R.n_edges / R.n_dst : 0.0;
                    const double tavg = R.n_src > 0 ? (double)R.n_edges / R.n_src : 0.0;
                    R.long_rows = (avg > kLongRowAvgDegree || h_misc[8 + r] > kLongRowMaxDegree) ? 1 : 0;
                    R.t_long_rows = (tavg > kLongRowAvgDegree || h_misc[8 + t] > kLongRowMaxDegree) ? 1 : 0;
                    R.pad_ = 0;
                }
                *plan_out = P;
            }
        }
    }
    cudaFree(ws);
    if (rc != AGX_OK) cudaFree(base);
    return rc;
}

extern "C" int agx_graph_plan_relation(const agx_graph_plan_t* plan, int r, agx_plan_rel_t* out) {
    AGX_CHECK_ARG(plan && out && r >= 0 && r < plan->n_rels, "agx_graph_plan_relation: bad plan / index %d", r);
    *out = plan->rel[r];
    return AGX_OK;
}

extern "C" int agx_graph_plan_destroy(agx_graph_plan_t* plan) {
    if (!plan) return AGX_OK;
    cudaFree(plan->buf);
    delete plan;
    return AGX_OK;
}

// ---- workspace layout of a layer call ----------------------------------------------------------------
namespace {

struct LayerWs {
    size_t frag, counters, colsum, part_l, part_r, dg, tmp, total;
    int split_dst;
};

int split_k_for(int64_t k_rows) {
    int64_t s = (k_rows + 383) / 384;
    if (s < 1) s = 1;
    if (s > 512) s = 512;
    return (int)s;
}

LayerWs layer_ws(const agx_sage_layer_t& L) {
    LayerWs W;
    const agx_plan_rel_t& R = L.rel;
    size_t off = 0;
    W.counters = off;       // int32 arrival counters (zeroed before every chunk launch)
    off = align_up(off + agx_chunk_counters(R.n_edges) * 4, 256);
    W.frag = off;
    off = align_up(off + agx_chunk_frag_floats(R.n_edges, L.f_src) * 4, 256);
    W.colsum = off;
    off = align_up(off + agx_colsum_workspace_floats(R.n_dst, 1, L.out_channels) * 4, 256);
    W.split_dst = split_k_for(R.n_dst);
    W.part_l = off;
    off = align_up(off + (size_t)W.split_dst * L.out_channels * L.f_src * 4, 256);
    W.part_r = off;
    off = align_up(off + (size_t)W.split_dst * L.out_channels * (L.w_r ? L.f_dst : 0) * 4, 256);
    W.dg = off;
    off = align_up(off + (size_t)R.n_dst * L.f_src * 4, 256);
    W.tmp = off;
    off = align_up(off + (size_t)R.n_src * L.f_src * 4, 256);
    W.total = off;
    return W;
}

int check_layer(const agx_sage_layer_t* L, const char* who) {
    AGX_CHECK_ARG(L, "%s: null layer", who);
    const agx_plan_rel_t& R = L->rel;
    AGX_CHECK_ARG(R.rowptr && R.t_rowptr && (R.n_edges == 0 || (R.col && R.t_col)) && R.n_src >= 0 &&
                      R.n_dst >= 0 && R.n_edges >= 0,
                  "%s: incomplete relation plan", who);
    AGX_CHECK_ARG(!L->mean || R.cnt, "%s: scatter-mean needs rel.cnt", who);
    AGX_CHECK_ARG(L->f_src >= 1 && L->out_channels >= 1 && L->x_src && L->w_l, "%s: x_src / w_l / sizes", who);
    AGX_CHECK_ARG((L->w_r == nullptr) == (L->x_dst == nullptr) || L->w_r == nullptr,
                  "%s: w_r needs x_dst", who);
    AGX_CHECK_ARG(!L->w_r || L->f_dst >= 1, "%s: f_dst", who);
    return AGX_OK;
}

// out[rows of the CSR] = (scaled) neighbour sums of x, picking the row-parallel or the edge-balanced kernel
int aggregate(const int32_t* rowptr, const int32_t* col, int n_rows, int n_edges, bool long_rows,
              const float* x, int64_t ldx, int F, const float* row_cnt, const float* nbr_scale, float* out,
              int accumulate, char* ws, const LayerWs& W, cudaStream_t st) {
    agx_rel_t rel;
    memset(&rel, 0, sizeof(rel));
    rel.rowptr = rowptr;
    rel.col = col;
    rel.x = x;
    rel.ldx = ldx;
    rel.row_cnt = row_cnt;
    rel.nbr_scale = nbr_scale;
    if (long_rows && !accumulate) {
        agx_chunk_seg_t S;
        memset(&S, 0, sizeof(S));
        S.rel = rel;
        S.out = out;
        S.ldo = F;
        S.n_rows = n_rows;
        S.n_edges = n_edges;
        S.frag = reinterpret_cast<float*>(ws + W.frag);
        S.counters = reinterpret_cast<int32_t*>(ws + W.counters);
        AGX_CUDA(cudaMemsetAsync(S.counters, 0, agx_chunk_counters(n_edges) * 4, st));
        return agx_aggregate_chunks(&S, 1, F, AGX_F32, st);
    }
    agx_row_group_t G;
    memset(&G, 0, sizeof(G));
    G.out = out;
    G.ldo = F;
    G.n_rows = n_rows;
    G.n_rel = 1;
    G.accumulate = accumulate;
    G.rel[0] = rel;
    return agx_aggregate_rows(&G, 1, F, AGX_F32, st);
}

}  // namespace

extern "C" size_t agx_sage_layer_workspace_bytes(const agx_sage_layer_t* layer) {
    if (check_layer(layer, "agx_sage_layer_workspace_bytes") != AGX_OK) return 0;
    return layer_ws(*layer).total;
}

extern "C" int agx_sage_layer_fwd(const agx_sage_layer_t* L, float* out, int64_t ldo, int accumulate,
                                  float* agg, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_layer(L, "agx_sage_layer_fwd");
    if (rc) return rc;
    const LayerWs W = layer_ws(*L);
    AGX_CHECK_ARG(out && agg && workspace && workspace_bytes >= W.total,
                  "agx_sage_layer_fwd: null buffer or workspace %zu < %zu", workspace_bytes, W.total);
    cudaStream_t st = (cudaStream_t)stream;
    const agx_plan_rel_t& R = L->rel;
    char* ws = static_cast<char*>(workspace);
    if (R.n_dst == 0) return AGX_OK;
    // 1. agg = mean / sum of the neighbours' rows (deterministic edge-order accumulation)
    rc = aggregate(R.rowptr, R.col, R.n_dst, R.n_edges, R.long_rows != 0, L->x_src, L->ld_src, L->f_src,
                   L->mean ? R.cnt : nullptr, nullptr, agg, 0, ws, W, st);
    if (rc) return rc;
    // 2. out (+)= agg w_l^T + x_dst w_r^T + b_l  (one two-segment GEMM)
    agx_gemm_seg_t seg[2];
    memset(seg, 0, sizeof(seg));
    seg[0].A = agg; seg[0].a_rs = L->f_src; seg[0].a_cs = 1;
    seg[0].B = L->w_l; seg[0].b_rs = 1; seg[0].b_cs = L->f_src;
    seg[0].K = L->f_src;
    int nseg = 1;
    if (L->w_r) {
        seg[1].A = L->x_dst; seg[1].a_rs = L->ld_dst; seg[1].a_cs = 1;
        seg[1].B = L->w_r; seg[1].b_rs = 1; seg[1].b_cs = L->f_dst;
        seg[1].K = L->f_dst;
        nseg = 2;
    }
    agx_gemm_problem_t Q;
    memset(&Q, 0, sizeof(Q));
    Q.C = out; Q.ldc = ldo; Q.bias = L->b_l;
    Q.M = R.n_dst; Q.N = L->out_channels;
    Q.accumulate = accumulate ? 1 : 0;
    Q.seg_begin = 0; Q.seg_count = nseg; Q.split_k = 1;
    return agx_gemm_grouped(&Q, 1, seg, nseg, st);
}

extern "C" int agx_sage_layer_bwd(const agx_sage_layer_t* L, const float* agg, const float* dout,
                                  int64_t lddo, float* d_w_l, float* d_b_l, float* d_w_r, float* d_x_src,
                                  int64_t ld_dxs, float* d_x_dst, int64_t ld_dxd, int accumulate_dx,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_layer(L, "agx_sage_layer_bwd");
    if (rc) return rc;
    const LayerWs W = layer_ws(*L);
    AGX_CHECK_ARG(agg && dout && d_w_l && workspace && workspace_bytes >= W.total,
                  "agx_sage_layer_bwd: null buffer or workspace %zu < %zu", workspace_bytes, W.total);
    AGX_CHECK_ARG(!d_w_r || L->w_r, "agx_sage_layer_bwd: d_w_r without w_r");
    AGX_CHECK_ARG(!d_x_dst || L->w_r, "agx_sage_layer_bwd: d_x_dst without w_r");
    AGX_CHECK_ARG(!d_x_src || ld_dxs == L->f_src, "agx_sage_layer_bwd: d_x_src must be dense [n_src, f_src]");
    cudaStream_t st = (cudaStream_t)stream;
    const agx_plan_rel_t& R = L->rel;
    char* ws = static_cast<char*>(workspace);
    const int O = L->out_channels;
    if (R.n_dst == 0) {
        AGX_CUDA(cudaMemsetAsync(d_w_l, 0, (size_t)O * L->f_src * 4, st));
        if (d_b_l) AGX_CUDA(cudaMemsetAsync(d_b_l, 0, (size_t)O * 4, st));
        if (d_w_r) AGX_CUDA(cudaMemsetAsync(d_w_r, 0, (size_t)O * L->f_dst * 4, st));
        return AGX_OK;
    }
    // 1. bias gradient: column sums of dout
    if (d_b_l) {
        agx_colsum_desc_t D;
        memset(&D, 0, sizeof(D));
        D.x = dout; D.ldx = lddo; D.out = d_b_l; D.n_rows = R.n_dst; D.F = O;
        rc = agx_colsum(&D, 1, reinterpret_cast<float*>(ws + W.colsum),
                        agx_colsum_workspace_floats(R.n_dst, 1, O), st);
        if (rc) return rc;
    }
    // 2. weight gradients (split over the node dimension, combined in slab order) and
    //    dG = dout w_l (the gradient of the aggregated neighbourhood)
    agx_gemm_seg_t seg[4];
    agx_gemm_problem_t Q[4];
    memset(seg, 0, sizeof(seg));
    memset(Q, 0, sizeof(Q));
    int np = 0;
    auto add = [&](float* C, int64_t ldc, int M, int N, const float* A, int64_t a_rs, int64_t a_cs,
                   const float* B, int64_t b_rs, int64_t b_cs, int K, int split, float* partial, int acc) {
        seg[np].A = A; seg[np].a_rs = a_rs; seg[np].a_cs = a_cs;
        seg[np].B = B; seg[np].b_rs = b_rs; seg[np].b_cs = b_cs; seg[np].K = K;
        Q[np].C = C; Q[np].ldc = ldc; Q[np].M = M; Q[np].N = N; Q[np].accumulate = acc;
        Q[np].seg_begin = np; Q[np].seg_count = 1; Q[np].split_k = split; Q[np].partial = split > 1 ? partial : nullptr;
        ++np;
    };
    add(d_w_l, L->f_src, O, L->f_src, dout, 1, lddo, agg, L->f_src, 1, R.n_dst, W.split_dst,
        reinterpret_cast<float*>(ws + W.part_l), 0);
    if (d_w_r)
        add(d_w_r, L->f_dst, O, L->f_dst, dout, 1, lddo, L->x_dst, L->ld_dst, 1, R.n_dst, W.split_dst,
            reinterpret_cast<float*>(ws + W.part_r), 0);
    float* dG = reinterpret_cast<float*>(ws + W.dg);
    if (d_x_src) add(dG, L->f_src, R.n_dst, L->f_src, dout, lddo, 1, L->w_l, L->f_src, 1, O, 1, nullptr, 0);
    if (d_x_dst)
        add(d_x_dst, ld_dxd, R.n_dst, L->f_dst, dout, lddo, 1, L->w_r, L->f_dst, 1, O, 1, nullptr,
            accumulate_dx ? 1 : 0);
    rc = agx_gemm_grouped(Q, np, seg, np, st);
    if (rc) return rc;
    // 3. d_x_src (+)= A^T dG: the transpose aggregation over the CSC, mean: neighbour i scaled by 1 / cnt[i]
    if (d_x_src && R.n_src > 0) {
        const bool lng = R.t_long_rows != 0;
        if (lng && accumulate_dx) {
            float* tmp = reinterpret_cast<float*>(ws + W.tmp);
            rc = aggregate(R.t_rowptr, R.t_col, R.n_src, R.n_edges, true, dG, L->f_src, L->f_src, nullptr,
                           L->mean ? R.cnt : nullptr, tmp, 0, ws, W, st);
            if (rc) return rc;
            agx_sum_desc_t S;
            memset(&S, 0, sizeof(S));
            S.out = d_x_src; S.in[0] = d_x_src; S.in[1] = tmp; S.n_in = 2;
            S.numel = (int64_t)R.n_src * L->f_src;
            rc = agx_sum_arrays(&S, 1, st);
        } else {
            rc = aggregate(R.t_rowptr, R.t_col, R.n_src, R.n_edges, lng, dG, L->f_src, L->f_src, nullptr,
                           L->mean ? R.cnt : nullptr, d_x_src, accumulate_dx ? 1 : 0, ws, W, st);
        }
        if (rc) return rc;
    }
    return AGX_OK;
}
