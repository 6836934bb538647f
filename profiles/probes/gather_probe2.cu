// Micro-benchmark (round 2): how many bytes must be in flight per SM for random 512 B / 128 B row
// gathers out of an L2-resident table to reach the L2 ceiling, and whether an LDGSTS (cp.async)
// ring in shared memory gets there at the occupancy the aggregation kernels run at.
//   ldg<U>      register gathers, U rows in flight per warp, occupancy limited by dynamic smem
//   cpasync<NG> per-warp ring of NG groups x 4 rows in shared memory (cp.async.cg 16 B per lane),
//               wait_group pipelining: NG*4 rows in flight per warp without registers
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe2 gather_probe2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(256) gather_ldg(const float4* __restrict__ x, const int* __restrict__ col, int n_per_warp, float4* out) {
    extern __shared__ uint8_t dummy[];
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int* c = col + w * n_per_warp;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < n_per_warp; i += 32) {
        const int mine = c[i + lane];
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int cj = __shfl_sync(0xffffffffu, mine, j0 + u);
                const float4* p = x + (int64_t)cj * 32 + lane;
                asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    out[w * 32 + lane] = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ring: [warp][NG][G rows][512 B]
template <int NG, int G>
__global__ void __launch_bounds__(256) gather_cpasync(const float4* __restrict__ x, const int* __restrict__ col, int n_per_warp, float4* out) {
    extern __shared__ __align__(128) uint8_t ring_all[];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * 8 + wq;
    const int* c = col + w * n_per_warp;
    uint8_t* ring = ring_all + (size_t)wq * NG * G * 512;
    const int ngroups = n_per_warp / G;          // n_per_warp = 128
    int cr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cr[k] = c[k * 32 + lane];
    auto issue = [&](int g) {
        const int slot = g % NG;
#pragma unroll
        for (int r = 0; r < G; ++r) {
            const int e = g * G + r;
            const int reg = e >> 5;
            int ck = cr[0];
            if (reg == 1) ck = cr[1];
            if (reg == 2) ck = cr[2];
            if (reg == 3) ck = cr[3];
            const int cj = __shfl_sync(0xffffffffu, ck, e & 31);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(ring + (slot * G + r) * 512 + lane * 16)), "l"(x + (int64_t)cj * 32 + lane) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int g = 0; g < NG - 1; ++g) issue(g);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int g = 0; g < ngroups; ++g) {
        if (g + NG - 1 < ngroups) issue(g + NG - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(NG - 1) : "memory");
        const int slot = g % NG;
#pragma unroll
        for (int r = 0; r < G; ++r) { const float4 v = *reinterpret_cast<const float4*>(ring + (slot * G + r) * 512 + lane * 16); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    out[w * 32 + lane] = acc;
}

template <typename F> float timeit(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

template <int U> int run_ldg(const float4* x, const int* col, float4* out, int warps, double bytes) {
    CK(cudaFuncSetAttribute(gather_ldg<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    for (int ctas : {8, 4, 3, 2}) {       // resident CTAs per SM forced through dynamic smem
        const int smem = ctas == 8 ? 0 : (ctas == 4 ? 54 : ctas == 3 ? 72 : 110) * 1024;
        float ms = timeit([&] { gather_ldg<U><<<warps / 8, 256, smem>>>(x, col, 128, out); });
        printf("  LDG U=%-2d  %d CTAs/SM (%3d KB in flight/SM) %7.1f us  %6.2f TB/s\n", U, ctas, ctas * 8 * U / 2, ms * 1e3, bytes / ms / 1e9);
    }
    return 0;
}
template <int NG, int G> int run_cp(const float4* x, const int* col, float4* out, int warps, double bytes) {
    const int smem = 8 * NG * G * 512;
    CK(cudaFuncSetAttribute(gather_cpasync<NG, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gather_cpasync<NG, G>, 256, smem);
    float ms = timeit([&] { gather_cpasync<NG, G><<<warps / 8, 256, smem>>>(x, col, 128, out); });
    printf("  cp.async NG=%d G=%d (%2d KB/CTA, %d CTAs/SM, %3d KB in flight/SM) %7.1f us  %6.2f TB/s\n", NG, G, smem / 1024, nb, nb * 8 * (NG - 1) * G / 2, ms * 1e3, bytes / ms / 1e9);
    return 0;
}

int main() {
    const int n_per_warp = 128; const int warps = 6992; const int64_t E = (int64_t)warps * n_per_warp;
    for (int64_t rows : {116475LL, 1164750LL}) {
        float4 *x, *out; int* col;
        CK(cudaMalloc(&x, rows * 512)); CK(cudaMalloc(&out, (size_t)warps * 512)); CK(cudaMalloc(&col, E * 4));
        CK(cudaMemset(x, 0, rows * 512));
        std::vector<int> h(E); std::mt19937 rng(1); for (auto& v : h) v = rng() % rows;
        CK(cudaMemcpy(col, h.data(), E * 4, cudaMemcpyHostToDevice));
        const double bytes = (double)E * 516;
        printf("table %.0f MB, %lld gathered rows (%.0f MB)\n", rows * 512 / 1e6, (long long)E, bytes / 1e6);
        run_ldg<4>(x, col, out, warps, bytes);
        run_ldg<8>(x, col, out, warps, bytes);
        run_ldg<16>(x, col, out, warps, bytes);
        run_cp<2, 4>(x, col, out, warps, bytes);
        run_cp<3, 4>(x, col, out, warps, bytes);
        run_cp<4, 4>(x, col, out, warps, bytes);
        run_cp<5, 4>(x, col, out, warps, bytes);
        run_cp<3, 8>(x, col, out, warps, bytes);
        run_cp<4, 8>(x, col, out, warps, bytes);
        cudaFree(x); cudaFree(out); cudaFree(col);
    }
    return 0;
}
