// Microbenchmark 2: completion latency of cp.async.bulk shared -> global stores to COLD rows (each store targets a row
// never touched before, like a snapshot ring that keeps every frame), the cost of wait_group when nothing is pending,
// and whether two back-to-back stores to the SAME row land in issue order.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void issue(void* dst, const void* src, int bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__global__ void probe(unsigned char* dst, size_t stride, int bytes, int gap, long long* out, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* src = sm + (size_t)warp * 2048;
    for (int i = lane; i < 512; i += 32) reinterpret_cast<int*>(src)[i] = i;
    __syncwarp();
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    long long t_read = 0, t_full = 0, t_idle = 0;
    for (int it = 0; it < iters; it++) {
        unsigned char* row = dst + (w * iters + it) * stride;
        __syncwarp();
        long long t0 = clock64();
        if (lane == 0) { issue(row, src, bytes); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
        __syncwarp();
        long long t1 = clock64();
        for (int k = 0; k < gap; k++) __nanosleep(20);   // unrelated work between the issue and the drain
        long long t1b = clock64();
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        long long t2 = clock64();
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        long long t3 = clock64();
        t_read += t1 - t0; t_full += t2 - t1b; t_idle += t3 - t2;
    }
    if (lane == 0 && warp == 0 && blockIdx.x == 0) { out[0] = t_read / iters; out[1] = t_full / iters; out[2] = t_idle / iters; }
}
// order test: store A (all words = 2*it+1) then store B (all words = 2*it+2) to the same row, no wait in between
__global__ void order_probe(int* dst, int words, int iters, int* bad) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* a = reinterpret_cast<int*>(sm + (size_t)warp * 4096);
    int* b = a + 512;
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    int* row = dst + w * 512;
    for (int it = 0; it < iters; it++) {
        for (int i = lane; i < words; i += 32) { a[i] = 2 * it + 1; b[i] = 2 * it + 2; }
        __syncwarp();
        if (lane == 0) { issue(row, a, words * 4); issue(row, b, words * 4); asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
        __syncwarp();
        int v = 0;
        for (int i = lane; i < words; i += 32) v |= __ldcg(row + i) != 2 * it + 2;
        if (__any_sync(0xffffffffu, v) && lane == 0) atomicAdd(bad, 1);
        __syncwarp();
    }
}
int main() {
    const size_t stride = 896; const int iters = 400, warps = 8;
    unsigned char* dst; long long* out; long long h[3]; int* bad; int hb = 0;
    cudaMalloc(&dst, (size_t)148 * warps * iters * stride + 4096);
    cudaMalloc(&out, 24); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    for (int gap : {0, 8, 32, 128}) {
        probe<<<148, warps * 32, warps * 2048>>>(dst, stride, 896, gap, out, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
        printf("cold rows, gap %3d sleeps: source-free %lld cyc, drain after gap %lld cyc, idle wait_group %lld cyc (%s)\n", gap, h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    }
    order_probe<<<148, warps * 32, warps * 4096>>>(reinterpret_cast<int*>(dst), 224, 2000, bad);
    cudaDeviceSynchronize();
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("same-row back-to-back stores: %d of %d rows ended with the FIRST store's data (%s)\n", hb, 148 * warps * 2000, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
