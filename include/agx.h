/*
 * agx.h -- C ABI of libagx.so: the B200 (sm_100a) kernels of the ArtGraph hetero-GNN + fusion-head
 * hot path.  No torch / C++ types cross this boundary: plain pointers, sizes and POD descriptor
 * structs.  Pointers are DEVICE pointers unless the parameter name starts with `h_`.
 *
 * Conventions
 *   - every function returns 0 on success, a negative agx_status otherwise; agx_last_error()
 *     returns a thread-local message.  No C++ exception leaves the library.
 *   - every function only ENQUEUES work on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises, allocates or frees.  Workspaces are caller-owned (size query functions).
 *     Descriptor tables (h_*) are read on the host during the call and are passed to the kernels
 *     by value, so a call can be captured into a CUDA graph.
 *   - the library is re-entrant and keeps no state besides the thread-local error string.
 *   - features are row-major float32 (`AGX_F32`) or bfloat16 (`AGX_BF16`, aggregation only);
 *     accumulation is always float32.  Indices are int32 inside the library; the reference's
 *     int64 `edge_index` tensors are consumed by agx_csr_build only.
 *
 * The reference has no native interface of its own (it is pure Python on PyG / torch); each entry
 * point cites the reference call site (path relative to the reference repo root) whose arithmetic
 * it replaces.  INTEGRATION.md shows the ctypes binding.
 */
#ifndef AGX_H_
#define AGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_VERSION 100          /* 0.1.0 */

typedef enum {
    AGX_OK = 0,
    AGX_ERR_INVALID = -1,        /* bad argument (message says which) */
    AGX_ERR_CUDA = -2,           /* a CUDA runtime call or launch failed */
    AGX_ERR_WORKSPACE = -3,      /* caller workspace too small */
    AGX_ERR_UNSUPPORTED = -4
} agx_status;

enum { AGX_F32 = 0, AGX_BF16 = 1, AGX_F64 = 2 /* agx_peer_allreduce only */ };
enum { AGX_SUM = 0, AGX_MEAN = 1 };

int agx_version(void);
const char* agx_last_error(void);
/* number of kernels this library has launched (or captured into a CUDA graph) in this process */
uint64_t agx_launch_count(void);
/* fills a comma separated list of the __global__ kernels compiled into the library */
int agx_kernel_inventory(char* h_buf, size_t h_buf_bytes);

/* ------------------------------------------------------------------------------------------
 * K1  CSR / CSC construction (stable counting sort of the edge lists)
 *   replaces: the implicit neighbour ordering of PyG `MessagePassing.propagate` on a dense
 *   `edge_index` (gather + torch_scatter.scatter in edge order), called from
 *   src/models/models_graph.py:30,38; edge_index layout of src/data/artgraph.py:106-112.
 * ------------------------------------------------------------------------------------------ */
#define AGX_MAX_CSR_RELS 40

typedef struct {
    const int64_t* keys;      /* [n_edges] row key per edge (dst for CSR, src for CSC)          */
    const int64_t* vals;      /* [n_edges] column value per edge (src for CSR, dst for CSC)     */
    int64_t n_edges;
    int64_t n_rows;           /* keys must lie in [0, n_rows)                                    */
    int64_t n_cols;           /* vals must lie in [0, n_cols)                                    */
} agx_edge_list_t;

/* Workspace bytes for agx_csr_build over relations with the given totals. */
size_t agx_csr_workspace_bytes(int64_t total_edges, int64_t total_rows);

/* Sorts all `n_rels` edge lists in ONE stable LSD radix sort over the composite key
 * (relation, key).  Outputs, relation r at offsets rows_before(r)+r / edges_before(r):
 *   rowptr [sum(n_rows_r + 1)] int32   local edge positions, rowptr_r[n_rows_r] == n_edges_r
 *   col    [sum(n_edges_r)]    int32   vals in sorted order
 *   eid    [sum(n_edges_r)]    int32   original edge id (the permutation); nullable
 *   cnt    [sum(n_rows_r)]     float   max(degree, 1) (the clamp of scatter-mean); nullable
 *   err    [1]                 int32   set to 1 if any key/val is out of range (never cleared)
 * Bit-exact with a stable sort by key: within a row, neighbours keep edge-list order. */
int agx_csr_build(const agx_edge_list_t* h_rels, int n_rels, int32_t* rowptr, int32_t* col,
                  int32_t* eid, float* cnt, int32_t* err, void* workspace, size_t workspace_bytes,
                  void* stream);

/* a-2: `T.ToUndirected()` for a non-bipartite store (src/train_gnn_embeddings.py:117-120):
 * coalesce(cat([row,col]), cat([col,row])): unique pairs ascending in row*n+col.
 * out_row/out_col [2*n_edges] int64; out_count[1] int64 = number of unique pairs. */
size_t agx_coalesce_workspace_bytes(int64_t n_edges);
int agx_coalesce_undirected(const int64_t* row, const int64_t* col, int64_t n_edges, int64_t n_nodes,
                            int64_t* out_row, int64_t* out_col, int64_t* out_count,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2/K3  neighbour aggregation (gather + segmented reduce), forward and transpose
 *   replaces: `x_src.index_select(0, ei[0])` + `torch_scatter.scatter(.., reduce)` inside
 *   SAGEConv/GraphConv.propagate (call sites src/models/models_graph.py:30,38) and, run over the
 *   CSC with neighbour scales, their autograd transposes (`index_add_`, gather of grad_out).
 * ------------------------------------------------------------------------------------------ */
#define AGX_MAX_REL_PER_GROUP 8
#define AGX_MAX_GROUPS 24
#define AGX_MAX_CHUNK_SEGS 24
#define AGX_CHUNK_EDGES 128

typedef struct {
    const int32_t* rowptr;    /* [n_rows+1]                                                      */
    const int32_t* col;       /* [n_edges]                                                       */
    const void* x;            /* source rows [*, F], leading dimension ldx (elements)            */
    int64_t ldx;
    const float* row_cnt;     /* nullable: divide the row sum by row_cnt[row]   (scatter-mean)   */
    const float* nbr_scale;   /* nullable: multiply neighbour j's row by 1/nbr_scale[j]
                                 (transpose of scatter-mean)                                      */
    const float* edge_w;      /* nullable: multiply the row gathered for slot e (position in `col`)
                                 by edge_w[e] (GATConv attention coefficients)                    */
    const int32_t* edge_w_idx;/* nullable: the weight of slot e is edge_w[edge_w_idx[e]] (the CSC
                                 reads the coefficients stored in CSR order)                      */
} agx_rel_t;

/* One output row = sum over up to 8 relations of that relation's (scaled) neighbour sum:
 * the per-destination-type `torch.add` chain of to_hetero (a-3) fused into the gather. */
typedef struct {
    void* out;                /* [n_rows, F], leading dimension ldo                              */
    int64_t ldo;
    int32_t n_rows;
    int32_t n_rel;
    int32_t accumulate;       /* 1: out += result                                                 */
    int32_t relu_dmask;       /* reserved                                                         */
    agx_rel_t rel[AGX_MAX_REL_PER_GROUP];
    const float* bias;        /* nullable: [F] float32 added to every output row                  */
} agx_row_group_t;

/* warp (or sub-warp) per destination row; deterministic edge-order accumulation */
int agx_aggregate_rows(const agx_row_group_t* h_groups, int n_groups, int F, int dtype,
                       void* stream);

typedef struct {
    agx_rel_t rel;
    void* out;                /* [n_rows, F]; rows of degree 0 are written as 0                  */
    int64_t ldo;
    int32_t n_rows;
    int32_t n_edges;
    float* frag;              /* [agx_chunk_frag_floats(n_edges, F)] scratch                     */
    int32_t* counters;        /* [agx_chunk_counters(n_edges)] arrival counters: ALL ZERO on entry,
                                 all zero again when the kernel has finished                      */
} agx_chunk_seg_t;

size_t agx_chunk_frag_floats(int64_t n_edges, int F);
size_t agx_chunk_counters(int64_t n_edges);
/* edge-balanced variant for relations with long rows (few destination rows, many edges): one warp
 * per 128 consecutive CSR edges; rows that cross chunks are completed, in fixed chunk order, by the
 * last warp to deliver a fragment (one launch, no float atomics, reproducible) */
int agx_aggregate_chunks(const agx_chunk_seg_t* h_segs, int n_segs, int F, int dtype,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  dense transforms (lin_l / lin_r / head Linear), float32 with float32 accumulation
 *   replaces: F.linear inside PyG Linear (SAGEConv.lin_l/lin_r, GraphConv.lin_rel/lin_root) and
 *   nn.Linear of src/models/models_kg.py:148-150,174-180,225-231,254,272, plus their autograd
 *   backward GEMMs.
 *   C[M,N] (+)= sum_s opA_s[M,K_s] * opB_s[K_s,N] (+ bias[N]) ; element (i,k) of opA_s is
 *   A[i*a_rs + k*a_cs], element (k,j) of opB_s is B[k*b_rs + j*b_cs].
 * ------------------------------------------------------------------------------------------ */
#define AGX_MAX_GEMM_PROBLEMS 24
#define AGX_MAX_GEMM_SEGS 64

typedef struct {
    const float* A; int64_t a_rs, a_cs;
    const float* B; int64_t b_rs, b_cs;
    const float* A_mask;      /* nullable: opA is multiplied elementwise by A_mask (same strides);
                                 the Dropout in front of the head Linear, never materialised     */
    const float* B_mask;      /* nullable: same for opB                                          */
    int32_t K;
    int32_t pad_;
} agx_gemm_seg_t;

typedef struct {
    float* C; int64_t ldc;
    const float* bias;        /* nullable, [N]                                                    */
    const float* row_scale;   /* nullable, [M]: C row i is divided by row_scale[i] in the epilogue */
    int32_t M, N;
    int32_t accumulate;       /* 1: C += result                                                   */
    int32_t seg_begin, seg_count;  /* into the segment table                                      */
    int32_t split_k;          /* >1: reduction of the single segment is split into this many
                                 slabs through `partial`, combined in slab order               */
    float* partial;           /* [split_k, M, N] when split_k > 1                                 */
    const int32_t* skip_flag; /* nullable: if *skip_flag != 0 the problem is skipped on device    */
} agx_gemm_problem_t;

int agx_gemm_grouped(const agx_gemm_problem_t* h_problems, int n_problems,
                     const agx_gemm_seg_t* h_segs, int n_segs, void* stream);

/* ------------------------------------------------------------------------------------------
 * small batched elementwise / reduction kernels of the GNN step
 * ------------------------------------------------------------------------------------------ */
#define AGX_MAX_TENSORS 48

/* out = sum of up to 8 equally-sized float arrays (sum of lin_r weights / lin_l biases that share
 * a destination type; exact restatement of adding the per-relation outputs, a-3) */
typedef struct {
    float* out; const float* in[8]; int32_t n_in; int64_t numel;
    const float* bias;        /* nullable: + bias[i % bias_F] (row-broadcast bias of [rows, bias_F]) */
    int64_t bias_F;
} agx_sum_desc_t;
int agx_sum_arrays(const agx_sum_desc_t* h_descs, int n, void* stream);

/* a-7  BatchNorm1d per node type (src/models/models_graph.py:19,32-33), batched over types, with
 * the following activation + dropout (:34-37) fused into the normalising pass.
 * Training statistics: two passes (mean, then centred second moment), per-slab float32 partials
 * combined in float64 (the reference's CPU kernel accumulates in double). */
typedef struct {
    const float* x; float* y;            /* [n_rows, F] contiguous; y nullable if y_act is given   */
    float* y_act;                         /* nullable: relu(y) * dropout mask                     */
    const float* dmask;                   /* nullable: [n_rows, F] multiplicative dropout mask    */
    const float* weight; const float* bias;   /* [F] affine                                       */
    float* running_mean; float* running_var;  /* [F]; updated in training mode, read in eval mode */
    float* save_mean; float* save_invstd;     /* [F] saved for backward                           */
    int32_t n_rows;
    /* dropout WITHOUT a materialised mask (dmask == NULL, drop_seed != NULL, drop_p > 0): element e
     * is kept iff agx_dropout_mask(mask, numel, drop_p, drop_seed) would have kept element
     * drop_offset + e of its buffer -- the same Philox4x32-10 stream, generated in the normalising
     * pass (kept values are scaled by 1 / (1 - drop_p)) */
    float drop_p;
    const uint64_t* drop_seed;            /* device: {key, counter}                               */
    int64_t drop_offset;                  /* multiple of 4                                        */
} agx_bn_desc_t;
size_t agx_bn_workspace_floats(int64_t total_rows, int n_descs, int F);
int agx_bn_forward(const agx_bn_desc_t* h_descs, int n, int F, int training, float momentum,
                   float eps, float* workspace, size_t workspace_floats, void* stream);

typedef struct {
    const float* x;                       /* BN input saved from forward                          */
    const float* y;                       /* BN output OR y_act (only "> 0" is read: the relu
                                           * gate; dropped elements get no gradient either way);
                                           * required if dy_act                                    */
    const float* dy;                      /* grad wrt y (nullable -> 0)                           */
    const float* dy_act;                  /* grad wrt y_act (nullable -> 0)                       */
    const float* dmask;                   /* nullable                                             */
    const float* weight;
    const float* save_mean; const float* save_invstd;
    float* dx;                            /* nullable                                             */
    float* dweight; float* dbias;         /* nullable; accumulated (+=)                           */
    int32_t n_rows;
    float act_scale;                      /* dmask == NULL: dy_act is multiplied by act_scale
                                           * (dropout without a mask tensor: y = y_act, scale
                                           * 1 / (1 - p) in float32); 0 = no scaling               */
} agx_bn_bwd_desc_t;
int agx_bn_backward(const agx_bn_bwd_desc_t* h_descs, int n, int F, int training, float* workspace,
                    size_t workspace_floats, void* stream);

/* Multi-GPU BatchNorm (the node type's rows are spread over the ranks; statistics over ALL rows,
 * as on one GPU).  phases is a bit mask; between phases the caller all-reduces the float64 buffer:
 *   forward : 1 -> sums[n][2F] = local column sums of x and of x*x            (all-reduce sums)
 *             4 -> mean / invstd / running stats from sums and counts; normalise (+ReLU, dropout)
 *   backward: 1 -> totals[n][2F] = local (sum dy, sum dy*xhat); dweight/dbias += local sums
 *                                                                             (all-reduce totals)
 *             2 -> dx with the global totals / counts
 * counts[n] = global row count per descriptor (float64, device). */
int agx_bn_forward_phase(const agx_bn_desc_t* h_descs, int n, int F, int training, float momentum,
                         float eps, float* workspace, size_t workspace_floats, int phases,
                         double* sums, const double* counts, void* stream);
int agx_bn_backward_phase(const agx_bn_bwd_desc_t* h_descs, int n, int F, int training,
                          float* workspace, size_t workspace_floats, int phases, double* totals,
                          const double* counts, void* stream);

/* column sums: out[F] (+)= sum_rows x[rows, F] (bias gradients) */
typedef struct { const float* x; int64_t ldx; float* out; int32_t n_rows; int32_t F; int32_t accumulate; int32_t pad_; } agx_colsum_desc_t;
size_t agx_colsum_workspace_floats(int64_t total_rows, int n_descs, int max_F);
int agx_colsum(const agx_colsum_desc_t* h_descs, int n, float* workspace, size_t workspace_floats,
               void* stream);

/* a-8  log_softmax(dim=1) + nll_loss (src/models/models_graph.py:39,
 * src/train_gnn_embeddings.py:29-30) and, with class weights, CrossEntropyLoss of the heads
 * (src/train_new_multimodal_multitask.py:48-55,79-81).  logp (nullable) receives the
 * log-probabilities.  With labels: row_ws[2*n_rows] scratch, loss_sum[2] = (sum w_y*nll, sum w_y)
 * reduced in row order in float64.  Labels outside [0,C) (ignore_index) get weight 0. */
int agx_log_softmax_nll(const float* logits, int64_t ld, int32_t n_rows, int32_t C,
                        const int64_t* labels, const float* class_w, float* logp, int64_t ldp,
                        float* loss_sum, float* row_ws, void* stream);
/* F.nll_loss on given log-probabilities (src/train_gnn_embeddings.py:29-30, the reference's call
 * shape): loss_sum[2] as above; backward writes the full dlogp [n_rows, C]
 * (-coef*gscale*w_y/loss_sum[1] at the label, 0 elsewhere). */
int agx_nll_forward(const float* logp, int64_t ld, int32_t n_rows, int32_t C, const int64_t* labels,
                    const float* class_w, float* loss_sum, float* row_ws, void* stream);
int agx_nll_backward(int32_t n_rows, int32_t C, const int64_t* labels, const float* class_w,
                     const float* loss_sum, const float* gscale, float coef, float* dlogp,
                     int64_t ld, void* stream);
/* loss[0] (+)= coef * loss_sum[0] / loss_sum[1]   (mean reduction) */
int agx_loss_finish(const float* loss_sum, float coef, float* loss, int accumulate, void* stream);
/* dlogits = coef * gscale[0] * w_y * (softmax - onehot) / loss_sum[1]
 *           (+ dlogp - softmax * rowsum(dlogp) when an upstream gradient on logp is given) */
int agx_log_softmax_nll_bwd(const float* logp, int64_t ldp, int32_t n_rows, int32_t C,
                            const int64_t* labels, const float* class_w, const float* loss_sum,
                            const float* gscale, float coef, const float* dlogp, int64_t lddp,
                            float* dlogits, int64_t ld, void* stream);

/* a-9  torch.optim.Adam step over a flat parameter arena (src/train_gnn_embeddings.py:144,
 * src/train_projector.py:34): exp_avg.lerp_, exp_avg_sq.mul_().addcmul_(), bias corrections from
 * the device step counter (so a captured CUDA graph advances). */
int agx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  int64_t numel, float lr, float beta1, float beta2, float eps, float weight_decay,
                  const int32_t* step /* device scalar, 1-based */, void* stream);

/* F.dropout / nn.Dropout mask (src/models/models_graph.py:37, src/models/models_kg.py:148,175):
 * mask[i] = (u_i >= p) / (1 - p), u_i from Philox4x32-10 keyed by seed[0], counter seed[1] + i/4.
 * seed is a DEVICE pointer to two uint64 so that graph replays draw fresh masks. */
int agx_dropout_mask(float* mask, int64_t numel, float p, const uint64_t* seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5/K6  fusion / projector heads
 *   cat -> Dropout -> Linear of src/models/models_kg.py:237-243 (and :158-162, :208-215) is a
 *   two-segment agx_gemm_grouped problem (feat and emb against column slices of W, A_mask = the
 *   dropout mask): the concatenation is never materialised.  The losses:
 *   CrossEntropyLoss = agx_log_softmax_nll above; SmoothL1Loss (src/train_projector.py:33,52) here.
 * ------------------------------------------------------------------------------------------ */
/* SmoothL1 (beta = 1, mean): loss[0] = mean(huber(out - target)); dout = clamp(diff,-1,1)/numel */
size_t agx_smooth_l1_workspace_floats(void);
int agx_smooth_l1(const float* out, const float* target, int64_t numel, float* loss,
                  float* dout /*nullable*/, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * GATConv attention aggregation
 *   replaces: GATConv.propagate / message / torch_geometric.utils.softmax (PyG 2.0.2, one head),
 *   selected by the reference's default --operator (src/train_gnn_embeddings.py:15,99).
 *   CSR / CSC of the relation's edge list WITH the self loops GATConv adds (agx_csr_build);
 *   a_l [n_src], a_r [n_dst] attention logits, x_l [n_src, F] transformed sources.
 *   The layer is split into per-edge SCALAR passes (softmax and its backward, below) and WIDE
 *   passes that run on the edge-balanced aggregation kernels with per-edge weights
 *   (agx_rel_t.edge_w): out = sum_j alpha_ij x_l[j] over the CSR, dx_l = sum_i alpha_ij dout[i]
 *   and da_l[j] = sum_i de_ij (F = 1) over the CSC, d alpha by agx_sddmm.  No row of any
 *   degree is ever walked by a single warp.
 * ------------------------------------------------------------------------------------------ */
#define AGX_MAX_GAT_RELS 24
#define AGX_GAT_LONG_ROW 1024    /* rows with more edges are reduced by a whole CTA, not a warp */

/* One relation of an attention layer.  All per-edge arrays are in CSR order (slot e of `col`). */
typedef struct {
    const int32_t* rowptr;    /* [n_rows+1] CSR by destination                                    */
    const int32_t* col;       /* [n_edges] source ids                                             */
    const float* a_l;         /* [n_src]  <x_l[j], att_l>                                         */
    const float* a_r;         /* [n_rows] <x_r[i], att_r>                                         */
    float* alpha;             /* [n_edges] softmax: written;  backward: read                      */
    const float* dalpha;      /* backward: [n_edges] d loss / d alpha (agx_sddmm of dout and x_l) */
    float* de;                /* backward: [n_edges] d loss / d (a_l[j] + a_r[i])                 */
    float* da_r;              /* backward: [n_rows]  sum_j de_ij                                  */
    const int32_t* long_rows; /* [n_long] the rows with more than AGX_GAT_LONG_ROW edges, found once
                                 per graph by the caller: each gets a CTA of its own                */
    int32_t n_rows;
    int32_t n_long;
} agx_gat_rel_t;

/* alpha_ij = softmax_i(leaky_relu(a_l[j] + a_r[i])) per destination row, for up to 24 relations in
 * ONE launch.  Eight lanes per row of up to 8 edges, a warp per row up to AGX_GAT_LONG_ROW edges;
 * every longer row (artwork -> style / genre / tag hubs, listed in `long_rows`) is reduced by a
 * 1024-thread CTA of its own, so hub rows neither queue behind one another nor behind short rows.
 * Fixed reduction order: reproducible. */
int agx_gat_edge_softmax(const agx_gat_rel_t* h_rels, int n_rels, float slope, void* stream);
/* de_ij = alpha_ij (dalpha_ij - sum_j alpha_ij dalpha_ij) leaky_relu'(a_l[j] + a_r[i]);
 * da_r[i] = sum_j de_ij */
int agx_gat_edge_softmax_bwd(const agx_gat_rel_t* h_rels, int n_rels, float slope, void* stream);

/* Sampled dense-dense product: out[e] = <a[row[e]], b[col[e]]> over F columns, edge-balanced (a warp
 * per 32 slots) -- d loss / d alpha_ij = <dout[i], x_l[j]> of the attention layer. */
#define AGX_MAX_SDDMM_SEGS 24
typedef struct {
    const int32_t* row;       /* [n_edges] row of `a` per slot (the destination of CSR slot e)    */
    const int32_t* col;       /* [n_edges] row of `b` per slot                                    */
    const float* a; int64_t lda;
    const float* b; int64_t ldb;
    float* out;               /* [n_edges]                                                        */
    int32_t n_edges;
    int32_t pad_;
} agx_sddmm_seg_t;
int agx_sddmm(const agx_sddmm_seg_t* h_segs, int n_segs, int32_t F, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer level: one SAGEConv / GraphConv relation, forward and backward, and the graph plan it
 * runs on -- the calls a C / C++ host makes instead of the Python planning layer.
 *   replaces: `conv((x_src, x_dst), edge_index)` at src/models/models_graph.py:30,38 (forward) and
 *   its autograd backward (src/train_gnn_embeddings.py:47):
 *       SAGEConv   out = lin_l(mean_{j in N(i)} x_src[j]) + lin_r(x_dst[i])        (PyG 2.0.2)
 *       GraphConv  out = lin_rel(sum_{j in N(i)} x_src[j]) + lin_root(x_dst[i])
 *   `accumulate` adds into `out`: the relation sum of to_hetero(aggr='sum') (models_graph.py:45).
 *   Ownership: agx_graph_plan_create allocates the CSR / CSC device buffers of all relations (the
 *   only allocations libagx.so ever owns) and agx_graph_plan_destroy frees them; everything else
 *   (features, weights, outputs, workspaces) is the caller's.  Calls enqueue on `stream`.
 * ------------------------------------------------------------------------------------------ */
typedef struct agx_graph_plan agx_graph_plan_t;      /* opaque */

typedef struct {
    const int32_t* rowptr;    /* CSR by destination: [n_dst + 1]                                  */
    const int32_t* col;       /* [n_edges] source ids, edge-list order inside a row                */
    const float* cnt;         /* [n_dst] max(in-degree, 1)                                        */
    const int32_t* t_rowptr;  /* CSC by source (the transpose): [n_src + 1]                        */
    const int32_t* t_col;     /* [n_edges] destination ids                                        */
    int32_t n_src, n_dst, n_edges;
    int32_t long_rows;        /* CSR has few / very long rows: edge-balanced kernel               */
    int32_t t_long_rows;      /* same for the CSC                                                 */
    int32_t pad_;
} agx_plan_rel_t;

/* rels[r]: keys = destination id, vals = source id of every edge (int64, device), n_rows = n_dst,
 * n_cols = n_src.  Sorts all relations (one batched K1 sort each way), reads the largest degrees
 * back once (synchronises `stream`). */
int agx_graph_plan_create(const agx_edge_list_t* h_rels, int n_rels, void* stream,
                          agx_graph_plan_t** plan);
int agx_graph_plan_relation(const agx_graph_plan_t* plan, int r, agx_plan_rel_t* out);
int agx_graph_plan_destroy(agx_graph_plan_t* plan);

typedef struct {
    agx_plan_rel_t rel;
    int32_t mean;             /* 1: SAGEConv (scatter-mean); 0: GraphConv (scatter-add)           */
    int32_t f_src, f_dst, out_channels;
    const float* x_src;       /* [n_src, f_src] float32 */
    int64_t ld_src;
    const float* x_dst;       /* [n_dst, f_dst]; may be NULL together with w_r (no root term)      */
    int64_t ld_dst;
    const float* w_l;         /* [out, f_src]  lin_l.weight / lin_rel.weight                       */
    const float* b_l;         /* [out] or NULL                                                    */
    const float* w_r;         /* [out, f_dst]  lin_r.weight / lin_root.weight, or NULL             */
} agx_sage_layer_t;

size_t agx_sage_layer_workspace_bytes(const agx_sage_layer_t* layer);
/* out[n_dst, out] (+)= aggr(x_src) w_l^T + b_l + x_dst w_r^T ; agg [n_dst, f_src] receives the
 * aggregated neighbourhood (needed by the backward pass) */
int agx_sage_layer_fwd(const agx_sage_layer_t* layer, float* out, int64_t ldo, int accumulate,
                       float* agg, void* workspace, size_t workspace_bytes, void* stream);
/* d_w_l [out, f_src], d_b_l [out], d_w_r [out, f_dst] are written; d_x_src [n_src, f_src] /
 * d_x_dst [n_dst, f_dst] (each nullable) are written, or added to when accumulate_dx != 0 */
int agx_sage_layer_bwd(const agx_sage_layer_t* layer, const float* agg, const float* dout,
                       int64_t lddo, float* d_w_l, float* d_b_l, float* d_w_r, float* d_x_src,
                       int64_t ld_dxs, float* d_x_dst, int64_t ld_dxd, int accumulate_dx,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5/K6 fused head step on the tensor cores (tcgen05, bf16 operands, float32 accumulation)
 *   replaces, per mini-batch, the whole of
 *     comb = cat(feat, emb); out = Linear(Dropout(comb)); loss = coef * CE(out, y; w); backward
 *       src/models/models_kg.py:237-243 (multitask), :158-162 / :208-215 (single task)
 *       src/train_new_multimodal_multitask.py:76-83 (run there under fp16 autocast)
 *     out = encoder(feat); loss = SmoothL1(out, emb); backward
 *       src/models/models_kg.py:261,278; src/train_projector.py:49-54 (fp16 autocast)
 *   One "head" = one Linear with <= 64 output columns over the virtual concatenation of up to two
 *   input parts; wider outputs (the projector's 128) are given as several heads over column slices.
 *   Per 128-row tile ONE CTA: parts staged by TMA, dropout applied while converting to bf16 (Philox
 *   from (seed, step) in-kernel, or an explicit mask), logits on tcgen05.mma into TMEM, softmax-CE /
 *   SmoothL1 + the logit gradient in the TMEM epilogue, d_weight^T accumulated in TMEM by a second
 *   tcgen05.mma pass over the same (L2-hot) tile; partials of the CTAs are combined in fixed order.
 *   Requirements: width[0] % 64 == 0, (width[0] + width[1]) % 128 == 0, C <= 64, rows 16-byte
 *   aligned.  Input gradients are NOT produced (precomputed backbone features, BASELINE config 4).
 * ------------------------------------------------------------------------------------------ */
#define AGX_HEAD_CE 0
#define AGX_HEAD_SMOOTH_L1 1
#define AGX_MAX_HEADS 4
typedef struct agx_head_t {
    const float* part[2];    /* [B, width[i]] float32 row-major; part[1] NULL when width[1] == 0 */
    int64_t ld[2];
    int32_t width[2];
    int32_t C;               /* output columns, 1..64 */
    int32_t loss;            /* AGX_HEAD_CE | AGX_HEAD_SMOOTH_L1 */
    const float* weight;     /* [C, K] float32 row-major, K = width[0] + width[1] */
    int64_t ldw;
    const float* bias;       /* [C] or NULL */
    const float* mask;       /* optional explicit multiplicative dropout mask [B, K]; overrides p_drop */
    int64_t ld_mask;
    const int64_t* labels;   /* CE: [B] */
    const float* class_w;    /* CE: [C] or NULL */
    float coef;              /* CE: weight of this head's loss term (0.5 in the multitask script) */
    float inv_count;         /* SMOOTH_L1: 1 / (global rows * total output columns) */
    const float* target;     /* SMOOTH_L1: [B, C] (column slice of the embedding) */
    int64_t ld_target;
    float* logits;           /* optional output [B, C] float32 */
    int64_t ld_logits;
    float* d_weight;         /* [C, K] gradient */
    int64_t ld_dw;
    float* d_bias;           /* [C] or NULL */
} agx_head_t;

/* bytes of caller-owned workspace for agx_head_step_prepare / agx_head_step (same n_heads, B) */
size_t agx_head_step_workspace_bytes(const agx_head_t* h_heads, int n_heads, int32_t B);
/* phase 1: bf16 copies of the weights into the workspace and norm[h] = sum_rows class_w[labels]
 * (= B without class weights; untouched for SMOOTH_L1 heads).  The caller may all-reduce norm
 * (batch-sharded heads) before phase 2. */
int agx_head_step_prepare(const agx_head_t* h_heads, int n_heads, int32_t B, float* norm,
                          void* workspace, size_t workspace_bytes, void* stream);
/* phase 2: forward + loss + backward.  loss[0] (+)= sum_h coef_h * sum_rows w nll / norm[h]
 * (resp. inv_count * sum huber); d_weight / d_bias are written, or added to when accumulate != 0.
 * seed_state: device int64 [2] (Philox key, step counter), NULL or p_drop == 0: no dropout. */
int agx_head_step(const agx_head_t* h_heads, int n_heads, int32_t B, float p_drop,
                  const int64_t* seed_state, const float* norm, float* loss, int accumulate,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * ContextNet / Castellano encoder heads
 *   replaces: nn.Tanh of the encoder (src/models/models_kg.py:80-85,120-125), MSELoss and
 *   SGD(momentum=0.9) of src/train_baseline_context.py:47-54.
 * ------------------------------------------------------------------------------------------ */
/* MSE (mean): loss[0] = mean((out - target)^2); dout = 2*(out - target)/numel; workspace as
 * agx_smooth_l1 */
int agx_mse(const float* out, const float* target, int64_t numel, float* loss,
            float* dout /*nullable*/, float* workspace, void* stream);
int agx_tanh(const float* x, float* y, int64_t numel, void* stream);
/* dx = dy * (1 - y^2) */
int agx_tanh_bwd(const float* y, const float* dy, float* dx, int64_t numel, void* stream);
/* torch.optim.SGD with momentum (dampening 0, no nesterov): buf = momentum*buf + g; p -= lr*buf */
int agx_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t numel, float lr,
                 float momentum, float weight_decay, void* stream);

/* misc elementwise */
int agx_fill_f32(float* p, int64_t numel, float v, void* stream);
int agx_scale_mask(const float* x, const float* mask, float* y, int64_t numel, void* stream);
/* rows gather: out[i,:] = table[idx[i],:]  (a-11 head input selection, src/data/data_kg.py:169-178)*/
int agx_gather_rows(const float* table, int64_t ld, const int64_t* idx, int64_t n, int32_t F,
                    float* out, int64_t ldo, void* stream);
/* identity detection for one-hot inputs (src/data/artgraph.py:93-95): flag[0] = 1 iff x == eye(n);
 * scratch[1] int32 */
int agx_is_identity(const float* x, int64_t ld, int32_t n, int32_t* flag, int32_t* scratch,
                    void* stream);
/* out[j, i] = in[i, j] ; in [rows, cols]; skipped on device when only_if_flag && *only_if_flag==0 */
int agx_transpose(const float* in, int64_t ld_in, int32_t rows, int32_t cols, float* out,
                  int64_t ld_out, const int32_t* only_if_flag, void* stream);
/* up to AGX_MAX_TENSORS transposes in one launch (the W^T copies of a hetero layer) */
typedef struct {
    const float* in; int64_t ld_in; float* out; int64_t ld_out; int32_t rows; int32_t cols;
    int32_t accumulate;       /* out += in^T (a weight gradient added into its optimizer buffer) */
    int32_t pad_;
} agx_transpose_desc_t;
int agx_transpose_batched(const agx_transpose_desc_t* h_descs, int n, void* stream);

/* halo exchange support (config 5): pack rows listed in idx into a contiguous send buffer, and
 * scatter-add received gradient rows back (deterministic: idx is unique per call) */
int agx_pack_rows(const float* x, int64_t ld, const int32_t* idx, int32_t n, int32_t F, float* out,
                  void* stream);
int agx_unpack_rows_add(float* x, int64_t ld, const int32_t* idx, int32_t n, int32_t F,
                        const float* in, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small all-reduce (sum) over NVLink / NVSwitch peer memory: one kernel, no NCCL call, for the
 * latency-bound reductions on the critical path of a multi-GPU step (BatchNorm statistics of the
 * rows of all ranks, the loss, the heads' label normaliser and gradient arena).  Push protocol:
 * every rank writes its elements, each 32-bit word tagged with the call's epoch, into its receive
 * slot on every peer and then collects the same elements of all ranks from its own buffer.
 * d_peer_bufs: DEVICE array of `world` base pointers to the ranks' symmetric buffers (each
 * agx_peer_allreduce_buffer_bytes(max payload, world) bytes, zero-initialised, mapped by the
 * caller -- torch.distributed._symmetric_memory in the Python binding); d_epoch: device int64[4],
 * zero-initialised, private to this rank.  Every rank must make the same sequence of calls.  Ranks
 * are added in rank order: bit-identical results on all ranks.  CUDA-graph capturable (the epoch
 * is advanced on the device).
 * ------------------------------------------------------------------------------------------ */
size_t agx_peer_allreduce_buffer_bytes(size_t max_payload_bytes, int world);
int agx_peer_allreduce(void* const* d_peer_bufs, int rank, int world, const void* in, void* out,
                       int64_t numel, int dtype /* AGX_F32 | AGX_F64 */, int64_t* d_epoch,
                       size_t max_payload_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGX_H_ */
