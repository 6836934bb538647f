/* A C host program on the layer-level C ABI of libagx.so (include/agx.h): builds a graph plan,
 * runs one SAGEConv relation forward + backward on the GPU and checks every result against loops
 * written here from the operator's definition (PyG 2.0.2 SAGEConv: out = lin_l(mean_j x_j) +
 * lin_r(x_i); call site /root/reference/src/models/models_graph.py:30).  No Python, no torch.
 *   gcc -std=c99 test_sage_layer.c -I../../include -I/usr/local/cuda/include -L<pkg> -lagx \
 *       -L/usr/local/cuda/lib64 -lcudart -lm
 * Exit code 0 and "OK" on success. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "agx.h"

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            return 2;                                                                    \
        }                                                                                \
    } while (0)
#define AGX(call)                                                                  \
    do {                                                                           \
        int rc_ = (call);                                                          \
        if (rc_ != 0) {                                                            \
            fprintf(stderr, "agx error %d (%s) at line %d\n", rc_, agx_last_error(), __LINE__); \
            return 3;                                                              \
        }                                                                          \
    } while (0)

static uint64_t rng_state = 88172645463325252ull;
static double urand(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

static double max_rel(const float* a, const double* b, size_t n) {
    double err = 0.0, scale = 1e-30;
    for (size_t i = 0; i < n; ++i) {
        const double d = fabs((double)a[i] - b[i]);
        if (d > err) err = d;
        if (fabs(b[i]) > scale) scale = fabs(b[i]);
    }
    return err / scale;
}

static void* to_device(const void* h, size_t bytes) {
    void* d = NULL;
    if (cudaMalloc(&d, bytes ? bytes : 4) != cudaSuccess) return NULL;
    if (bytes) cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    return d;
}

static int run_case(int mean, int n_src, int n_dst, int n_edges, int hub_edges, int FS, int FD, int O) {
    /* edges: the first hub_edges point at destination 3 (a long row: edge-balanced kernel), the
     * rest are uniform; destinations >= n_dst - 5 stay isolated (mean of nothing = 0) */
    int64_t* src = malloc(sizeof(int64_t) * n_edges);
    int64_t* dst = malloc(sizeof(int64_t) * n_edges);
    for (int e = 0; e < n_edges; ++e) {
        src[e] = (int64_t)(urand() * n_src);
        dst[e] = e < hub_edges ? 3 : (int64_t)(urand() * (n_dst - 5));
    }
    float* xs = malloc(sizeof(float) * n_src * FS);
    float* xd = malloc(sizeof(float) * n_dst * FD);
    float* wl = malloc(sizeof(float) * O * FS);
    float* bl = malloc(sizeof(float) * O);
    float* wr = malloc(sizeof(float) * O * FD);
    float* go = malloc(sizeof(float) * n_dst * O);
    for (int i = 0; i < n_src * FS; ++i) xs[i] = (float)(urand() * 2 - 1);
    for (int i = 0; i < n_dst * FD; ++i) xd[i] = (float)(urand() * 2 - 1);
    for (int i = 0; i < O * FS; ++i) wl[i] = (float)((urand() * 2 - 1) / sqrt(FS));
    for (int i = 0; i < O; ++i) bl[i] = (float)(urand() * 0.2 - 0.1);
    for (int i = 0; i < O * FD; ++i) wr[i] = (float)((urand() * 2 - 1) / sqrt(FD));
    for (int i = 0; i < n_dst * O; ++i) go[i] = (float)(urand() * 2 - 1);

    /* ---- reference (double) ---- */
    double* agg = calloc((size_t)n_dst * FS, sizeof(double));
    double* cnt = calloc(n_dst, sizeof(double));
    for (int e = 0; e < n_edges; ++e) {
        cnt[dst[e]] += 1.0;
        for (int k = 0; k < FS; ++k) agg[dst[e] * FS + k] += xs[src[e] * FS + k];
    }
    for (int i = 0; i < n_dst; ++i) {
        if (cnt[i] < 1.0) cnt[i] = 1.0;
        if (mean)
            for (int k = 0; k < FS; ++k) agg[i * FS + k] /= cnt[i];
    }
    double* out = calloc((size_t)n_dst * O, sizeof(double));
    for (int i = 0; i < n_dst; ++i)
        for (int o = 0; o < O; ++o) {
            double s = bl[o];
            for (int k = 0; k < FS; ++k) s += agg[i * FS + k] * wl[o * FS + k];
            for (int k = 0; k < FD; ++k) s += (double)xd[i * FD + k] * wr[o * FD + k];
            out[i * O + o] = s;
        }
    double* dwl = calloc((size_t)O * FS, sizeof(double));
    double* dbl = calloc(O, sizeof(double));
    double* dwr = calloc((size_t)O * FD, sizeof(double));
    double* dagg = calloc((size_t)n_dst * FS, sizeof(double));
    double* dxs = calloc((size_t)n_src * FS, sizeof(double));
    double* dxd = calloc((size_t)n_dst * FD, sizeof(double));
    for (int i = 0; i < n_dst; ++i)
        for (int o = 0; o < O; ++o) {
            const double g = go[i * O + o];
            dbl[o] += g;
            for (int k = 0; k < FS; ++k) {
                dwl[o * FS + k] += g * agg[i * FS + k];
                dagg[i * FS + k] += g * wl[o * FS + k];
            }
            for (int k = 0; k < FD; ++k) {
                dwr[o * FD + k] += g * xd[i * FD + k];
                dxd[i * FD + k] += g * wr[o * FD + k];
            }
        }
    for (int e = 0; e < n_edges; ++e)
        for (int k = 0; k < FS; ++k)
            dxs[src[e] * FS + k] += dagg[dst[e] * FS + k] / (mean ? cnt[dst[e]] : 1.0);

    /* ---- libagx.so ---- */
    int64_t* d_src = to_device(src, sizeof(int64_t) * n_edges);
    int64_t* d_dst = to_device(dst, sizeof(int64_t) * n_edges);
    agx_edge_list_t el;
    el.keys = d_dst;
    el.vals = d_src;
    el.n_edges = n_edges;
    el.n_rows = n_dst;
    el.n_cols = n_src;
    agx_graph_plan_t* plan = NULL;
    AGX(agx_graph_plan_create(&el, 1, NULL, &plan));
    agx_sage_layer_t L;
    memset(&L, 0, sizeof(L));
    AGX(agx_graph_plan_relation(plan, 0, &L.rel));
    if (hub_edges > 256 && !L.rel.long_rows) {
        fprintf(stderr, "plan did not flag the hub row\n");
        return 4;
    }
    L.mean = mean;
    L.f_src = FS;
    L.f_dst = FD;
    L.out_channels = O;
    L.x_src = to_device(xs, sizeof(float) * n_src * FS);
    L.ld_src = FS;
    L.x_dst = to_device(xd, sizeof(float) * n_dst * FD);
    L.ld_dst = FD;
    L.w_l = to_device(wl, sizeof(float) * O * FS);
    L.b_l = to_device(bl, sizeof(float) * O);
    L.w_r = to_device(wr, sizeof(float) * O * FD);
    const size_t ws_bytes = agx_sage_layer_workspace_bytes(&L);
    void* ws = NULL;
    CK(cudaMalloc(&ws, ws_bytes ? ws_bytes : 4));
    float *d_out, *d_agg, *d_go, *d_dwl, *d_dbl, *d_dwr, *d_dxs, *d_dxd;
    CK(cudaMalloc((void**)&d_out, sizeof(float) * n_dst * O));
    CK(cudaMalloc((void**)&d_agg, sizeof(float) * n_dst * FS));
    d_go = to_device(go, sizeof(float) * n_dst * O);
    CK(cudaMalloc((void**)&d_dwl, sizeof(float) * O * FS));
    CK(cudaMalloc((void**)&d_dbl, sizeof(float) * O));
    CK(cudaMalloc((void**)&d_dwr, sizeof(float) * O * FD));
    CK(cudaMalloc((void**)&d_dxs, sizeof(float) * n_src * FS));
    CK(cudaMalloc((void**)&d_dxd, sizeof(float) * n_dst * FD));
    AGX(agx_sage_layer_fwd(&L, d_out, O, 0, d_agg, ws, ws_bytes, NULL));
    AGX(agx_sage_layer_bwd(&L, d_agg, d_go, O, d_dwl, d_dbl, d_dwr, d_dxs, FS, d_dxd, FD, 0, ws, ws_bytes, NULL));
    /* accumulate-into-dst (the relation sum of to_hetero): a second forward doubles the output */
    AGX(agx_sage_layer_fwd(&L, d_out, O, 1, d_agg, ws, ws_bytes, NULL));
    CK(cudaDeviceSynchronize());

    float* h = malloc(sizeof(float) * ((size_t)n_dst * O + (size_t)n_src * FS + (size_t)n_dst * FD + (size_t)O * (FS + FD + 1)));
    int bad = 0;
#define CHECK(dev, ref, n, scale, what)                                                       \
    do {                                                                                      \
        CK(cudaMemcpy(h, dev, sizeof(float) * (n), cudaMemcpyDeviceToHost));                  \
        for (size_t q_ = 0; q_ < (size_t)(n); ++q_) h[q_] = (float)(h[q_] / (scale));         \
        const double e_ = max_rel(h, ref, n);                                                 \
        printf("  %-10s rel err %.2e\n", what, e_);                                           \
        if (!(e_ <= 1e-5)) bad = 1;                                                           \
    } while (0)
    printf("%s  n_src=%d n_dst=%d E=%d hub=%d  F=%d/%d -> %d  long_rows=%d/%d\n", mean ? "SAGEConv (mean)" : "GraphConv (add)",
           n_src, n_dst, n_edges, hub_edges, FS, FD, O, L.rel.long_rows, L.rel.t_long_rows);
    CHECK(d_out, out, (size_t)n_dst * O, 2.0, "out");
    CHECK(d_dwl, dwl, (size_t)O * FS, 1.0, "d_w_l");
    CHECK(d_dbl, dbl, (size_t)O, 1.0, "d_b_l");
    CHECK(d_dwr, dwr, (size_t)O * FD, 1.0, "d_w_r");
    CHECK(d_dxs, dxs, (size_t)n_src * FS, 1.0, "d_x_src");
    CHECK(d_dxd, dxd, (size_t)n_dst * FD, 1.0, "d_x_dst");
    AGX(agx_graph_plan_destroy(plan));
    return bad;
}

int main(void) {
    if (agx_version() != 100) {
        fprintf(stderr, "unexpected agx_version %d\n", agx_version());
        return 1;
    }
    int rc = run_case(1, 300, 200, 3000, 400, 48, 20, 64);     /* hub row: edge-balanced kernel */
    if (rc) return rc;
    rc = run_case(0, 5000, 900, 6000, 0, 128, 128, 32);        /* short rows: row-parallel kernel */
    if (rc) return rc;
    rc = run_case(1, 40, 3000, 9000, 0, 16, 128, 128);         /* few sources feeding many rows */
    if (rc) return rc;
    printf("OK\n");
    return 0;
}
