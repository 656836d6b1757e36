// bench_barrier.cu -- latency of the synchronisation primitives the merge kernel can choose from,
// measured on the device it runs on: the hand-written grid barrier of eliminate.cu (one atomic,
// one acquire poll), cooperative_groups grid.sync(), a thread-block-cluster barrier, __syncthreads
// and a dependent chain of L2 / DRAM gathers (the other half of a merge pass).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/bench_barrier tools/bench_barrier.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// (a) the barrier of eliminate.cu
__global__ void k_hand(unsigned *bar, int iters, long long *cyc)
{
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned target = (unsigned)(it + 1) * gridDim.x;
            __threadfence();
            atomicAdd(bar, 1u);
            while (ld_acquire(bar) < target) { }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (b) release-add + relaxed poll + one acquire fence
__global__ void k_hand2(unsigned *bar, int iters, long long *cyc)
{
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned target = (unsigned)(it + 1) * gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
            while (ld_relaxed(bar) < target) { }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (c) cooperative groups
__global__ void k_cg(int iters, long long *cyc)
{
    cg::grid_group g = cg::this_grid();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (d) cluster barrier (one cluster)
__global__ void k_cluster(int iters, long long *cyc)
{
    cg::cluster_group c = cg::this_cluster();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) c.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (d2) cluster barrier with a gpu-scope fence before it (what global-memory hand-over needs)
__global__ void k_cluster_fence(int iters, long long *cyc)
{
    cg::cluster_group c = cg::this_cluster();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) { __threadfence(); c.sync(); }
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (e) __syncthreads
__global__ void k_block(int iters, long long *cyc)
{
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

// (f) dependent gather chain through a random permutation (one thread)
__global__ void k_chase(const unsigned *perm, int steps, long long *cyc, unsigned *sink)
{
    unsigned p = 0;
    const long long t0 = clock64();
    for (int i = 0; i < steps; i++) p = __ldcg(perm + p);
    *cyc = clock64() - t0;
    *sink = p;
}

// (g) the same chain, 32 lanes of every warp of a block each following their own chain
__global__ void k_chase_wide(const unsigned *perm, unsigned n, int steps, long long *cyc, unsigned *sink)
{
    unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) * 977u % n;
    const long long t0 = clock64();
    for (int i = 0; i < steps; i++) p = __ldcg(perm + p);
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
    if (p == 0xffffffffu) *sink = p;
}

int main()
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const double mhz = prop.clockRate / 1e3;
    printf("%s, %d SMs, %.0f MHz\n", prop.name, prop.multiProcessorCount, mhz);
    unsigned *bar;
    long long *cyc, h;
    CK(cudaMalloc(&bar, 256));
    CK(cudaMalloc(&cyc, 8));
    const int iters = 2000;
    const int nSM = prop.multiProcessorCount;

    for (int threads : {128, 512}) {
        for (int blocks : {nSM / 4, nSM / 2, nSM, 2 * nSM}) {
            void *args[] = {&bar, (void *)&iters, &cyc};
            int perSM = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_hand, threads, 0));
            if (blocks > perSM * nSM) continue;
            CK(cudaMemset(bar, 0, 256));
            CK(cudaLaunchCooperativeKernel((void *)k_hand, dim3(blocks), dim3(threads), args, 0, 0));
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("hand  barrier  %4d blocks x %3d: %7.2f us\n", blocks, threads, h / mhz / iters);
            CK(cudaMemset(bar, 0, 256));
            CK(cudaLaunchCooperativeKernel((void *)k_hand2, dim3(blocks), dim3(threads), args, 0, 0));
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("hand2 barrier  %4d blocks x %3d: %7.2f us\n", blocks, threads, h / mhz / iters);
            void *args2[] = {(void *)&iters, &cyc};
            CK(cudaLaunchCooperativeKernel((void *)k_cg, dim3(blocks), dim3(threads), args2, 0, 0));
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("cg grid.sync   %4d blocks x %3d: %7.2f us\n", blocks, threads, h / mhz / iters);
        }
    }
    for (int csize : {1, 2, 4, 8, 16}) {
        for (int threads : {256, 1024}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize);
            cfg.blockDim = dim3(threads);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (csize > 8) {
                CK(cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                CK(cudaFuncSetAttribute(k_cluster_fence, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            }
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, iters, cyc);
            if (e != cudaSuccess) { printf("cluster %d x %d: %s\n", csize, threads, cudaGetErrorString(e)); cudaGetLastError(); continue; }
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("cluster.sync   %4d CTAs  x %4d: %7.3f us", csize, threads, h / mhz / iters);
            CK(cudaLaunchKernelEx(&cfg, k_cluster_fence, iters, cyc));
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("   with __threadfence: %7.3f us\n", h / mhz / iters);
        }
    }
    k_block<<<1, 1024>>>(iters, cyc);
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("__syncthreads  1024 threads: %7.3f us\n", h / mhz / iters);

    // pointer chase: one cycle through n entries (Sattolo), n*4 bytes
    for (size_t mb : {8, 64, 512, 2048}) {
        const unsigned n = (unsigned)(mb << 20) / 4;
        unsigned *hp = (unsigned *)malloc((size_t)n * 4);
        for (unsigned i = 0; i < n; i++) hp[i] = i;
        unsigned long long s = 88172645463325252ull;
        for (unsigned i = n - 1; i > 0; i--) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            unsigned j = (unsigned)(s % i);
            unsigned t = hp[i]; hp[i] = hp[j]; hp[j] = t;
        }
        unsigned *dp, *sink;
        CK(cudaMalloc(&dp, (size_t)n * 4));
        CK(cudaMalloc(&sink, 4));
        CK(cudaMemcpy(dp, hp, (size_t)n * 4, cudaMemcpyHostToDevice));
        const int steps = 4000;
        k_chase<<<1, 1>>>(dp, steps, cyc, sink);     // warm (partially)
        k_chase<<<1, 1>>>(dp, steps, cyc, sink);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("dependent gather, %4zu MB table, 1 thread : %7.3f us per step\n", mb, h / mhz / steps);
        k_chase_wide<<<nSM, 512>>>(dp, n, 200, cyc, sink);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("dependent gather, %4zu MB table, %d x 512 : %7.3f us per step\n", mb, nSM, h / mhz / 200);
        cudaFree(dp); cudaFree(sink); free(hp);
    }
    return 0;
}
