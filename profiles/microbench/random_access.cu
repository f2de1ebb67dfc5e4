// Microbenchmark behind DESIGN.md's memory-access choices: how fast can a B200 do random
// 1-byte gathers and random atomics over tables much larger than L2, per load flavour and
// L2 fetch granularity, and how fast is the same gather when the table fits in L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_access random_access.cu && ./random_access
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

template <int MODE>
__device__ __forceinline__ uint32_t load_byte(const uint8_t *p) {
    uint32_t v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) v = __ldcg(p);
    else if (MODE == 2) v = __ldcs(p);
    else if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p)); }
    else if (MODE == 4) { asm volatile("ld.global.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p)); }
    else { v = *(volatile const uint8_t *)p; }
    return v;
}

template <int MODE, int ILP>
__global__ void gather(const uint8_t *tab, uint64_t mask, uint64_t n_per_thread, uint32_t *out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += ILP) {
        uint32_t v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) v[j] = load_byte<MODE>(tab + (mix(tid * n_per_thread + i + j) & mask));
#pragma unroll
        for (int j = 0; j < ILP; j++) acc += v[j];
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

// two-level: summary (L2 resident) says "maybe"; only then touch the big table
template <int ILP>
__global__ void gather2(const uint32_t *summary, const uint8_t *tab, uint64_t mask, uint64_t n_per_thread, uint32_t *out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += ILP) {
        uint64_t idx[ILP]; uint32_t s[ILP], v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) { idx[j] = mix(tid * n_per_thread + i + j) & mask; s[j] = __ldg(summary + (idx[j] >> 7)); }
#pragma unroll
        for (int j = 0; j < ILP; j++) { v[j] = 0; if ((s[j] >> ((idx[j] >> 2) & 31)) & 1) v[j] = __ldg(tab + idx[j]); }
#pragma unroll
        for (int j = 0; j < ILP; j++) acc += v[j];
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <int MODE>
__global__ void atomics(uint32_t *tab, uint64_t mask_words, uint64_t n_per_thread) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i = 0; i < n_per_thread; i += 8) {
        uint32_t old[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t h = mix(tid * n_per_thread + i + j);
            uint32_t *p = tab + (h & mask_words);
            uint32_t one = 1u << (8 * ((h >> 40) & 3));
            if (MODE == 0) old[j] = atomicCAS(p, 0u, one);
            else if (MODE == 1) old[j] = atomicAdd(p, one);
            else { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(one) : "memory"); old[j] = 0; }
        }
        uint32_t a = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) a += old[j];
        if (a == 0xdeadbeef) tab[0] = a;
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    const uint64_t GiB = 1ULL << 30;
    uint8_t *big; uint32_t *out; uint32_t *summary;
    cudaMalloc(&big, 8 * GiB); cudaMalloc(&out, 64); cudaMalloc(&summary, 32 << 20);
    cudaMemset(big, 0, 8 * GiB); cudaMemset(summary, 0, 32 << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 8, threads = 256; const uint64_t T = (uint64_t)blocks * threads;
    const uint64_t per = 512; const double N = (double)T * per;
    size_t lim = 0;
    for (int gran : {0, 32, 64, 128}) {
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
        printf("== cudaLimitMaxL2FetchGranularity = %zu\n", lim);
#define RUN(NAME, KERNEL, ...) do { KERNEL<<<blocks, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); \
        cudaEventRecord(a); KERNEL<<<blocks, threads>>>(__VA_ARGS__); cudaEventRecord(b); cudaEventSynchronize(b); \
        printf("  %-44s %8.3f ms  %7.1f G/s  (%s)\n", NAME, time_ms(a, b), N / time_ms(a, b) / 1e6, cudaGetErrorString(cudaGetLastError())); } while (0)
        RUN("gather 1GiB ldg(nc)        ILP8", (gather<0, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB ldcg           ILP8", (gather<1, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB ldcs           ILP8", (gather<2, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB nc.L1::no_alloc ILP8", (gather<3, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB L1::no_alloc   ILP8", (gather<4, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB volatile       ILP8", (gather<5, 8>), big, GiB - 1, per, out);
        RUN("gather 1GiB ldg(nc)        ILP1", (gather<0, 1>), big, GiB - 1, per, out);
        RUN("gather 1GiB ldg(nc)        ILP16", (gather<0, 16>), big, GiB - 1, per, out);
        RUN("gather 8GiB ldg(nc)        ILP8", (gather<0, 8>), big, 8 * GiB - 1, per, out);
    }
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    printf("== L2-resident tables (granularity 32)\n");
    RUN("gather 32MiB ldg           ILP8", (gather<0, 8>), big, (32ULL << 20) - 1, per, out);
    RUN("gather 32MiB ldcg          ILP8", (gather<1, 8>), big, (32ULL << 20) - 1, per, out);
    RUN("gather 64MiB ldg           ILP8", (gather<0, 8>), big, (64ULL << 20) - 1, per, out);
    RUN("gather 16MiB ldg           ILP8", (gather<0, 8>), big, (16ULL << 20) - 1, per, out);
    RUN("gather 4MiB  ldg           ILP8", (gather<0, 8>), big, (4ULL << 20) - 1, per, out);
    RUN("two-level: 32MiB summary all-zero + 1GiB   ", (gather2<8>), summary, big, GiB - 1, per, out);
    cudaMemset(summary, 0x11, 32 << 20); // 25 % of the summary bits set
    RUN("two-level: 25% pass to 1GiB table          ", (gather2<8>), summary, big, GiB - 1, per, out);
    printf("== random atomics on u32 words, 8 in flight per thread\n");
    const uint64_t sizes[4] = {8 * GiB, GiB, 64ULL << 20, 16ULL << 20};
    for (uint64_t sz : sizes) {
        char nm[64];
        cudaMemset(big, 0, sz);
        snprintf(nm, 64, "CAS(0->1) over %5llu MiB", (unsigned long long)(sz >> 20)); RUN(nm, (atomics<0>), (uint32_t *)big, sz / 4 - 1, per);
        snprintf(nm, 64, "atomicAdd over %5llu MiB", (unsigned long long)(sz >> 20)); RUN(nm, (atomics<1>), (uint32_t *)big, sz / 4 - 1, per);
        snprintf(nm, 64, "red.add   over %5llu MiB", (unsigned long long)(sz >> 20)); RUN(nm, (atomics<2>), (uint32_t *)big, sz / 4 - 1, per);
    }
    return 0;
}
