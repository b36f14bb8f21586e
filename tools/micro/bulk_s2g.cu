// Microbenchmark: rate of small cp.async.bulk shared->global copies (UBLKCP) per SM, vs. plain vector stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int SZ, int MODE>
__global__ void __launch_bounds__(512) k(uint8_t *dst, uint64_t dst_bytes, int rounds, int active_lanes)
{
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < 65536 / 4; i += 512) ((uint32_t *)sm)[i] = i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint64_t x = (blockIdx.x * 512ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < rounds; ++r) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const uint64_t off = ((x >> 20) % (dst_bytes / 128)) * 128;       // random 128-B aligned destination
        const uint32_t so = (threadIdx.x * 128u) & 65535u;
        if (MODE == 0) {
            if (lane < active_lanes) bulk_s2g(dst + off, sm + so, SZ);
            bulk_commit();
            bulk_wait_read();
        } else {
            // SZ/16 lanes... here: each lane stores its own SZ bytes with 16-B vector stores
            if (lane < active_lanes) {
#pragma unroll
                for (int q = 0; q < SZ / 16; ++q) *(uint4 *)(dst + off + 16 * q) = *(const uint4 *)(sm + so + 16 * q);
            }
        }
    }
}

template <int SZ, int MODE> void run(uint8_t *d, uint64_t bytes, int rounds, int lanes, const char *name)
{
    cudaFuncSetAttribute(k<SZ, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<SZ, MODE><<<296, 512, 65536>>>(d, bytes, 4, lanes);
    cudaEventRecord(e0);
    k<SZ, MODE><<<296, 512, 65536>>>(d, bytes, rounds, lanes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double copies = 296.0 * 16 * lanes * rounds;
    printf("%-10s SZ=%3d lanes=%2d: %8.3f ms  %8.2f G copies/s  %7.1f GB/s  per-SM %.3f copies/clk (1.965GHz)  %s\n", name, SZ, lanes, ms,
           copies / ms / 1e6, copies * SZ / ms / 1e6, copies / ms / 1e6 / 148 / 1.965, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
    uint8_t *d;
    const uint64_t bytes = 6ull << 30;
    cudaMalloc(&d, bytes);
    for (int lanes : {32, 8, 1}) {
        run<16, 0>(d, bytes, 2000, lanes, "bulk");
        run<32, 0>(d, bytes, 2000, lanes, "bulk");
        run<64, 0>(d, bytes, 2000, lanes, "bulk");
        run<128, 0>(d, bytes, 2000, lanes, "bulk");
    }
    run<16, 1>(d, bytes, 2000, 32, "st.v4");
    run<32, 1>(d, bytes, 2000, 32, "st.v4");
    run<64, 1>(d, bytes, 2000, 32, "st.v4");
    return 0;
}
