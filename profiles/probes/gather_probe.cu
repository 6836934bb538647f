// Micro-benchmark: ceiling of random 512 B row gathers on B200 (register LDG vs TMA bulk ring).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(256) gather_ldg(const float4* __restrict__ x, const int* __restrict__ col, int n_per_warp, float4* out) {
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
template <int NG>
__global__ void __launch_bounds__(256) gather_tma(const float4* __restrict__ x, const int* __restrict__ col, int n_per_warp, float4* out) {
    extern __shared__ __align__(128) uint8_t ring_all[];
    __shared__ uint64_t bar[8][NG];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * 8 + wq;
    const int* c = col + w * n_per_warp;
    uint8_t* ring = ring_all + (size_t)wq * NG * 8 * 512;
    if (lane == 0) { for (int g = 0; g < NG; ++g) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[wq][g]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int ngroups = n_per_warp / 8;
    auto issue = [&](int g) {
        const int slot = g % NG;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[wq][slot])), "r"(8 * 512) : "memory");
        __syncwarp();
        if (lane < 8) {
            const int cj = c[g * 8 + lane];
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (slot * 8 + lane) * 512)), "l"(x + (int64_t)cj * 32), "r"(512), "r"(s32(&bar[wq][slot])) : "memory");
        }
    };
    for (int g = 0; g < NG && g < ngroups; ++g) issue(g);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int g = 0; g < ngroups; ++g) {
        const int slot = g % NG; const uint32_t par = (g / NG) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&bar[wq][slot])), "r"(par) : "memory");
#pragma unroll
        for (int r = 0; r < 8; ++r) { const float4 v = *reinterpret_cast<const float4*>(ring + (slot * 8 + r) * 512 + lane * 16); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        __syncwarp();
        if (g + NG < ngroups) issue(g + NG);
    }
    out[w * 32 + lane] = acc;
}

__global__ void stream_copy(const float4* __restrict__ a, float4* __restrict__ b, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F> float timeit(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
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
        float ms;
        ms = timeit([&] { gather_ldg<4><<<warps / 8, 256>>>(x, col, n_per_warp, out); });  printf("  LDG  U=4   %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        ms = timeit([&] { gather_ldg<8><<<warps / 8, 256>>>(x, col, n_per_warp, out); });  printf("  LDG  U=8   %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        ms = timeit([&] { gather_ldg<16><<<warps / 8, 256>>>(x, col, n_per_warp, out); }); printf("  LDG  U=16  %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        ms = timeit([&] { gather_ldg<32><<<warps / 8, 256>>>(x, col, n_per_warp, out); }); printf("  LDG  U=32  %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        cudaFuncSetAttribute(gather_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8 * 512);
        cudaFuncSetAttribute(gather_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 8 * 512);
        cudaFuncSetAttribute(gather_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 8 * 512);
        ms = timeit([&] { gather_tma<2><<<warps / 8, 256, 8 * 2 * 8 * 512>>>(x, col, n_per_warp, out); }); printf("  TMA  NG=2  %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        ms = timeit([&] { gather_tma<3><<<warps / 8, 256, 8 * 3 * 8 * 512>>>(x, col, n_per_warp, out); }); printf("  TMA  NG=3  %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        ms = timeit([&] { gather_tma<4><<<warps / 8, 256, 8 * 4 * 8 * 512>>>(x, col, n_per_warp, out); }); printf("  TMA  NG=4  %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
        float4* y; CK(cudaMalloc(&y, E * 512 / 2));
        float4* z; CK(cudaMalloc(&z, E * 512 / 2));
        ms = timeit([&] { stream_copy<<<148 * 8, 256>>>(y, z, E * 512 / 2 / 16); }); printf("  stream copy of the same bytes (r+w) %7.1f us  %6.2f TB/s\n", ms * 1e3, (double)E * 512 / ms / 1e9);
        cudaFree(x); cudaFree(out); cudaFree(col); cudaFree(y); cudaFree(z);
    }
    return 0;
}
