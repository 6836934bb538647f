// K1: CSR / CSC construction for all relations of a heterograph in one stable LSD radix sort,
// plus the coalesce step of ToUndirected for non-bipartite stores.
//
// The sort key is the composite (relation, row) = row_base[relation] + row, the payload the
// global edge id.  A stable sort by that key keeps, inside every row, the order of the edge
// list -- the order in which the reference's CPU scatter_add_ accumulates, so
// downstream sums are reproducible and indices are bit-exact with torch.sort(stable=True).
//
// HBM-bound integer work: each radix pass reads 8 B and writes 8 B per edge; tiles of 2048
// edges per CTA, 128 B coalesced loads per warp step, warp-private digit counters in shared
// memory (no global atomics, no float atomics anywhere).
#include "agx_common.cuh"

namespace agx {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortSteps = 8;                                  // 32-lane steps per warp
constexpr int kSortTile = kSortThreads * kSortSteps;           // 2048 keys per CTA

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct CsrRels {
    const int64_t* keys[AGX_MAX_CSR_RELS];
    const int64_t* vals[AGX_MAX_CSR_RELS];
    int64_t ebase[AGX_MAX_CSR_RELS + 1];    // edges before relation r
    int64_t rbase[AGX_MAX_CSR_RELS + 1];    // rows before relation r
    int64_t ncols[AGX_MAX_CSR_RELS];
    int n;
};

__device__ __forceinline__ int find_rel(const int64_t* base, int n, int64_t g) {
    // last r with base[r] <= g  (relations with zero edges are skipped naturally)
    int lo = 0, hi = n;        // invariant: base[lo] <= g < base[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (base[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
csr_make_keys(const __grid_constant__ CsrRels R, uint32_t* __restrict__ key, uint32_t* __restrict__ val,
              int32_t* __restrict__ err) {
    const int64_t total = R.ebase[R.n];
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int r = find_rel(R.ebase, R.n, g);
        const int64_t le = g - R.ebase[r];
        int64_t k = R.keys[r][le];
        const int64_t v = R.vals[r][le];
        const int64_t nrows = R.rbase[r + 1] - R.rbase[r];
        if (k < 0 || k >= nrows || v < 0 || v >= R.ncols[r]) {
            *err = 1;
            k = 0;
        }
        key[g] = (uint32_t)(R.rbase[r] + k);
        val[g] = (uint32_t)g;
    }
}

// ---- per-tile digit histogram: hist[digit * nb + tile] --------------------------------------
__global__ void __launch_bounds__(kSortThreads)
radix_hist(const uint32_t* __restrict__ key, int64_t n, int shift, uint32_t* __restrict__ hist, int nb) {
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t t0 = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const int64_t i = t0 + s * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(key[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// ---- stable scatter: rank = (elements with the same digit earlier in the tile) ---------------
__global__ void __launch_bounds__(kSortThreads)
radix_scatter(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
              uint32_t* __restrict__ kout, uint32_t* __restrict__ vout,
              const uint32_t* __restrict__ offs, int64_t n, int shift, int nb) {
    __shared__ uint32_t cnt[kSortWarps][kRadix];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
    const uint32_t gbase = offs[(size_t)threadIdx.x * nb + blockIdx.x];
    __syncthreads();

    const int64_t w0 = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (32 * kSortSteps);
    uint32_t k[kSortSteps], v[kSortSteps], rank[kSortSteps];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const int64_t i = w0 + s * 32 + lane;
        const bool valid = i < n;
        k[s] = valid ? kin[i] : 0u;
        v[s] = valid ? vin[i] : 0u;
        const uint32_t d = (k[s] >> shift) & (kRadix - 1);
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
        const uint32_t r = __popc(m & lt);
        rank[s] = valid ? cnt[w][d] + r : 0u;
        __syncwarp();
        if (valid && r == 0) cnt[w][d] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x, seeded with the tile's global offset
        uint32_t run = gbase;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
            const uint32_t t = cnt[ww][threadIdx.x];
            cnt[ww][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const int64_t i = w0 + s * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[s] >> shift) & (kRadix - 1);
            const uint32_t pos = cnt[w][d] + rank[s];
            kout[pos] = k[s];
            vout[pos] = v[s];
        }
    }
}

// ---- exclusive scan (3 phases) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t s = lane < nw ? wsum[lane] : 0u;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        wsum[lane] = si - s;                       // exclusive warp offsets
        if (lane == 31 && total) *total = si;
    }
    __syncthreads();
    const uint32_t r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += in[base + i];
    __shared__ uint32_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_spine(uint32_t* __restrict__ sums, int n) {
    __shared__ uint32_t carry, tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < n; b += 1024) {
        const int i = b + threadIdx.x;
        const uint32_t v = i < n ? sums[i] : 0u;
        const uint32_t e = block_exclusive_scan(v, &tot);
        if (i < n) sums[i] = e + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? data[base + i] : 0u;
        s += v[i];
    }
    uint32_t run = block_exclusive_scan(s, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

static int exclusive_scan_u32(uint32_t* data, int64_t n, uint32_t* sums, cudaStream_t st) {
    const int nblk = (int)ceil_div(n, kScanTile);
    scan_reduce<<<nblk, kScanThreads, 0, st>>>(data, n, sums);
    AGX_LAUNCH_CHECK("scan_reduce");
    scan_spine<<<1, 1024, 0, st>>>(sums, nblk);
    AGX_LAUNCH_CHECK("scan_spine");
    scan_apply<<<nblk, kScanThreads, 0, st>>>(data, n, sums);
    AGX_LAUNCH_CHECK("scan_apply");
    return AGX_OK;
}

static size_t sort_scratch_bytes(int64_t n) {
    const int64_t nb = ceil_div(n, kSortTile);
    const int64_t nh = nb * kRadix;
    return align_up((size_t)nh * 4, 256) + align_up((size_t)ceil_div(nh, kScanTile) * 4 + 4, 256);
}

// Stable LSD radix sort of (key, val) pairs on the low `bits` bits.  Ping-pongs between (k0,v0)
// and (k1,v1); returns in *result_in_0 which pair holds the result.
static int radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, int64_t n,
                            int bits, void* scratch, cudaStream_t st, bool* result_in_0) {
    *result_in_0 = true;
    if (n == 0) return AGX_OK;
    const int nb = (int)ceil_div(n, kSortTile);
    uint32_t* hist = (uint32_t*)scratch;
    uint32_t* sums = (uint32_t*)((char*)scratch + align_up((size_t)nb * kRadix * 4, 256));
    uint32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
    for (int shift = 0; shift < bits; shift += kRadixBits) {
        radix_hist<<<nb, kSortThreads, 0, st>>>(ki, n, shift, hist, nb);
        AGX_LAUNCH_CHECK("radix_hist");
        int rc = exclusive_scan_u32(hist, (int64_t)nb * kRadix, sums, st);
        if (rc) return rc;
        radix_scatter<<<nb, kSortThreads, 0, st>>>(ki, vi, ko, vo, hist, n, shift, nb);
        AGX_LAUNCH_CHECK("radix_scatter");
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
        *result_in_0 = !*result_in_0;
    }
    return AGX_OK;
}

static int bits_for(int64_t n_values) {   // bits needed to represent values in [0, n_values)
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n_values) ++b;
    return b;
}

// ---- finalize: col / eid from the sorted payload, rowptr / cnt by binary search ----------------
__global__ void __launch_bounds__(256)
csr_finalize(const __grid_constant__ CsrRels R, const uint32_t* __restrict__ vsorted,
             int32_t* __restrict__ col, int32_t* __restrict__ eid) {
    const int64_t total = R.ebase[R.n];
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = vsorted[p];
        const int r = find_rel(R.ebase, R.n, g);
        const int64_t le = g - R.ebase[r];
        int64_t v = R.vals[r][le];
        if (v < 0 || v >= R.ncols[r]) v = 0;
        col[p] = (int32_t)v;
        if (eid) eid[p] = (int32_t)le;
    }
}

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t* a, int64_t n, uint32_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// rowptr / cnt from the sorted keys (key = rows before the relation + row id), three small passes
// instead of two binary searches over all edges per row (60 us on the full graph):
//   A  thread per sorted position p: where the key changes, the rows in (key[p-1], key[p]] start at
//      p -- the first kRowptrRun of them are written here (a run longer than that is a block of rows
//      without edges, left at -1);
//   B  thread per rowptr slot: the closing slot of every relation; slots still at -1 do one search;
//   C  cnt[row] = max(rowptr[row + 1] - rowptr[row], 1)   (the divisor of scatter-mean).
constexpr int kRowptrRun = 8;

__global__ void __launch_bounds__(256)
csr_rowptr_edges(const __grid_constant__ CsrRels R, const uint32_t* __restrict__ ksorted,
                 int32_t* __restrict__ rowptr) {
    const int64_t total_rows = R.rbase[R.n];
    const int64_t E = R.ebase[R.n];
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= E;
         p += (int64_t)gridDim.x * blockDim.x) {
        int64_t cur = p < E ? (int64_t)ksorted[p] : total_rows - 1;
        if (cur > total_rows - 1) cur = total_rows - 1;          // (ids flagged in err: stay in range)
        const int64_t lo = p > 0 ? (int64_t)ksorted[p - 1] + 1 : 0;
        if (lo > cur) continue;                                  // same key as the edge before
        const int64_t hi = cur < lo + kRowptrRun - 1 ? cur : lo + kRowptrRun - 1;
        int r = find_rel(R.rbase, R.n, lo);
        for (int64_t g = lo; g <= hi; ++g) {
            while (g >= R.rbase[r + 1]) ++r;
            rowptr[g + r] = (int32_t)(p - R.ebase[r]);
        }
    }
}

__global__ void __launch_bounds__(256)
csr_rowptr_fill(const __grid_constant__ CsrRels R, const uint32_t* __restrict__ ksorted,
                int32_t* __restrict__ rowptr) {
    const int64_t total = R.rbase[R.n] + R.n;         // sum(n_rows_r + 1)
    const int64_t total_edges = R.ebase[R.n];
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        // relation of slot q: last r with rbase[r] + r <= q
        int lo = 0, hi = R.n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (R.rbase[mid] + mid <= q) lo = mid; else hi = mid;
        }
        const int r = lo;
        const int64_t i = q - (R.rbase[r] + r);
        const int64_t nrows = R.rbase[r + 1] - R.rbase[r];
        if (i >= nrows) {
            rowptr[q] = (int32_t)(R.ebase[r + 1] - R.ebase[r]);
        } else if (rowptr[q] < 0) {
            const int64_t pos = lower_bound_u32(ksorted, total_edges, (uint32_t)(R.rbase[r] + i));
            rowptr[q] = (int32_t)(pos - R.ebase[r]);
        }
    }
}

__global__ void __launch_bounds__(256)
csr_row_counts(const __grid_constant__ CsrRels R, const int32_t* __restrict__ rowptr,
               float* __restrict__ cnt) {
    const int64_t total_rows = R.rbase[R.n];
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total_rows;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int r = find_rel(R.rbase, R.n, g);
        const int32_t d = rowptr[g + r + 1] - rowptr[g + r];
        cnt[g] = (float)(d < 1 ? 1 : d);
    }
}

// ---- coalesce (ToUndirected on a non-bipartite store) -----------------------------------------
__global__ void __launch_bounds__(256)
coalesce_prepare(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t e,
                 uint32_t* __restrict__ key, uint32_t* __restrict__ val, int by_row,
                 const uint32_t* __restrict__ perm) {
    // symmetrised list: element g < e is (row[g], col[g]); element g >= e is (col[g-e], row[g-e])
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < 2 * e;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = perm ? perm[p] : p;
        const int64_t a = g < e ? row[g] : col[g - e];
        const int64_t b = g < e ? col[g] : row[g - e];
        key[p] = (uint32_t)(by_row ? a : b);
        val[p] = (uint32_t)g;
    }
}

__global__ void __launch_bounds__(256)
coalesce_flag(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t e,
              const uint32_t* __restrict__ perm, uint32_t* __restrict__ flag) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < 2 * e;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = perm[p];
        const int64_t a = g < e ? row[g] : col[g - e];
        const int64_t b = g < e ? col[g] : row[g - e];
        uint32_t f = 1;
        if (p > 0) {
            const int64_t h = perm[p - 1];
            const int64_t a0 = h < e ? row[h] : col[h - e];
            const int64_t b0 = h < e ? col[h] : row[h - e];
            f = (a0 != a || b0 != b) ? 1u : 0u;
        }
        flag[p] = f;
    }
}

__global__ void __launch_bounds__(256)
coalesce_compact(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t e,
                 const uint32_t* __restrict__ perm, const uint32_t* __restrict__ pos,
                 int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                 int64_t* __restrict__ out_count) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < 2 * e;
         p += (int64_t)gridDim.x * blockDim.x) {
        const bool last = p == 2 * e - 1;
        const bool keep = last ? true : pos[p + 1] != pos[p];   // flag[p] == 1 (exclusive scan)
        const int64_t g = perm[p];
        const int64_t a = g < e ? row[g] : col[g - e];
        const int64_t b = g < e ? col[g] : row[g - e];
        bool is_first = keep;
        if (last) {
            // flag of the last element: differs from its predecessor?
            if (p > 0) {
                const int64_t h = perm[p - 1];
                const int64_t a0 = h < e ? row[h] : col[h - e];
                const int64_t b0 = h < e ? col[h] : row[h - e];
                is_first = (a0 != a || b0 != b);
            }
            *out_count = (int64_t)pos[p] + (is_first ? 1 : 0);
        }
        if (is_first) {
            out_row[pos[p]] = a;
            out_col[pos[p]] = b;
        }
    }
}

}  // namespace agx

using namespace agx;

extern "C" size_t agx_csr_workspace_bytes(int64_t total_edges, int64_t total_rows) {
    (void)total_rows;
    const size_t e = align_up((size_t)(total_edges > 0 ? total_edges : 1) * 4, 256);
    return 4 * e + sort_scratch_bytes(total_edges > 0 ? total_edges : 1);
}

extern "C" int agx_csr_build(const agx_edge_list_t* h_rels, int n_rels, int32_t* rowptr,
                             int32_t* col, int32_t* eid, float* cnt, int32_t* err,
                             void* workspace, size_t workspace_bytes, void* stream) {
    AGX_CHECK_ARG(h_rels && n_rels >= 1 && n_rels <= AGX_MAX_CSR_RELS,
                  "agx_csr_build: n_rels=%d out of [1,%d]", n_rels, AGX_MAX_CSR_RELS);
    AGX_CHECK_ARG(rowptr && err, "agx_csr_build: rowptr/err must not be null");
    cudaStream_t st = (cudaStream_t)stream;
    CsrRels R;
    R.n = n_rels;
    R.ebase[0] = 0;
    R.rbase[0] = 0;
    for (int r = 0; r < n_rels; ++r) {
        const agx_edge_list_t& L = h_rels[r];
        AGX_CHECK_ARG(L.n_edges >= 0 && L.n_rows >= 0 && L.n_cols >= 0,
                      "agx_csr_build: relation %d has negative sizes", r);
        AGX_CHECK_ARG(L.n_edges == 0 || (L.keys && L.vals),
                      "agx_csr_build: relation %d has null edge arrays", r);
        R.keys[r] = L.keys;
        R.vals[r] = L.vals;
        R.ncols[r] = L.n_cols;
        R.ebase[r + 1] = R.ebase[r] + L.n_edges;
        R.rbase[r + 1] = R.rbase[r] + L.n_rows;
    }
    const int64_t E = R.ebase[n_rels], NR = R.rbase[n_rels];
    AGX_CHECK_ARG(E < ((int64_t)1 << 31) && NR < ((int64_t)1 << 31),
                  "agx_csr_build: more than 2^31 edges or rows");
    AGX_CHECK_ARG(E == 0 || col, "agx_csr_build: col must not be null");
    if (workspace_bytes < agx_csr_workspace_bytes(E, NR)) {
        set_error("agx_csr_build: workspace %zu < required %zu", workspace_bytes,
                  agx_csr_workspace_bytes(E, NR));
        return AGX_ERR_WORKSPACE;
    }
    const size_t eb = align_up((size_t)(E > 0 ? E : 1) * 4, 256);
    uint32_t* k0 = (uint32_t*)workspace;
    uint32_t* v0 = (uint32_t*)((char*)workspace + eb);
    uint32_t* k1 = (uint32_t*)((char*)workspace + 2 * eb);
    uint32_t* v1 = (uint32_t*)((char*)workspace + 3 * eb);
    void* scratch = (char*)workspace + 4 * eb;

    const uint32_t* ks = k0;
    const uint32_t* vs = v0;
    if (E > 0) {
        const int grid = (int)(ceil_div(E, 256) < 148 * 16 ? ceil_div(E, 256) : 148 * 16);
        csr_make_keys<<<grid, 256, 0, st>>>(R, k0, v0, err);
        AGX_LAUNCH_CHECK("csr_make_keys");
        bool in0 = true;
        int rc = radix_sort_pairs(k0, v0, k1, v1, E, bits_for(NR), scratch, st, &in0);
        if (rc) return rc;
        ks = in0 ? k0 : k1;
        vs = in0 ? v0 : v1;
        csr_finalize<<<grid, 256, 0, st>>>(R, vs, col, eid);
        AGX_LAUNCH_CHECK("csr_finalize");
    }
    {
        const int64_t slots = NR + n_rels;
        const int grid = (int)(ceil_div(slots, 256) < 148 * 16 ? ceil_div(slots, 256) : 148 * 16);
        AGX_CUDA(cudaMemsetAsync(rowptr, 0xFF, (size_t)slots * sizeof(int32_t), st));
        const int ge = (int)(ceil_div(E + 1, 256) < 148 * 16 ? ceil_div(E + 1, 256) : 148 * 16);
        if (NR > 0) {
            csr_rowptr_edges<<<ge, 256, 0, st>>>(R, ks, rowptr);
            AGX_LAUNCH_CHECK("csr_rowptr_edges");
        }
        csr_rowptr_fill<<<grid, 256, 0, st>>>(R, ks, rowptr);
        AGX_LAUNCH_CHECK("csr_rowptr_fill");
        if (cnt && NR > 0) {
            csr_row_counts<<<grid, 256, 0, st>>>(R, rowptr, cnt);
            AGX_LAUNCH_CHECK("csr_row_counts");
        }
    }
    return AGX_OK;
}

extern "C" size_t agx_coalesce_workspace_bytes(int64_t n_edges) {
    const int64_t n = 2 * (n_edges > 0 ? n_edges : 1);
    return 4 * align_up((size_t)n * 4, 256) + sort_scratch_bytes(n) + 256;
}

extern "C" int agx_coalesce_undirected(const int64_t* row, const int64_t* col, int64_t n_edges,
                                       int64_t n_nodes, int64_t* out_row, int64_t* out_col,
                                       int64_t* out_count, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    AGX_CHECK_ARG(n_edges >= 0 && n_nodes >= 0 && n_nodes < ((int64_t)1 << 32),
                  "agx_coalesce_undirected: bad sizes");
    AGX_CHECK_ARG(out_count, "agx_coalesce_undirected: out_count is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_edges == 0) {
        AGX_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st));
        return AGX_OK;
    }
    AGX_CHECK_ARG(row && col && out_row && out_col, "agx_coalesce_undirected: null arrays");
    AGX_CHECK_ARG(2 * n_edges < ((int64_t)1 << 31), "agx_coalesce_undirected: too many edges");
    if (workspace_bytes < agx_coalesce_workspace_bytes(n_edges)) {
        set_error("agx_coalesce_undirected: workspace too small");
        return AGX_ERR_WORKSPACE;
    }
    const int64_t n = 2 * n_edges;
    const size_t eb = align_up((size_t)n * 4, 256);
    uint32_t* k0 = (uint32_t*)workspace;
    uint32_t* v0 = (uint32_t*)((char*)workspace + eb);
    uint32_t* k1 = (uint32_t*)((char*)workspace + 2 * eb);
    uint32_t* v1 = (uint32_t*)((char*)workspace + 3 * eb);
    void* scratch = (char*)workspace + 4 * eb;
    const int grid = (int)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
    const int bits = bits_for(n_nodes);
    bool in0;
    // LSD over the pair (row, col): stable sort by col first, then by row
    coalesce_prepare<<<grid, 256, 0, st>>>(row, col, n_edges, k0, v0, 0, nullptr);
    AGX_LAUNCH_CHECK("coalesce_prepare");
    int rc = radix_sort_pairs(k0, v0, k1, v1, n, bits, scratch, st, &in0);
    if (rc) return rc;
    uint32_t* perm1 = in0 ? v0 : v1;
    uint32_t* ka = in0 ? k1 : k0;       // free pair
    uint32_t* va = in0 ? v1 : v0;
    uint32_t* kb = in0 ? k0 : k1;       // holds perm1 in its val half: keep perm1 intact while preparing
    coalesce_prepare<<<grid, 256, 0, st>>>(row, col, n_edges, ka, va, 1, perm1);
    AGX_LAUNCH_CHECK("coalesce_prepare");
    rc = radix_sort_pairs(ka, va, kb, perm1, n, bits, scratch, st, &in0);
    if (rc) return rc;
    uint32_t* perm = in0 ? va : perm1;
    uint32_t* flag = in0 ? kb : ka;     // reuse a key buffer not holding the result
    coalesce_flag<<<grid, 256, 0, st>>>(row, col, n_edges, perm, flag);
    AGX_LAUNCH_CHECK("coalesce_flag");
    rc = exclusive_scan_u32(flag, n, (uint32_t*)scratch, st);   // hist region: >= n/2048 words
    if (rc) return rc;
    coalesce_compact<<<grid, 256, 0, st>>>(row, col, n_edges, perm, flag, out_row, out_col,
                                           out_count);
    AGX_LAUNCH_CHECK("coalesce_compact");
    return AGX_OK;
}
