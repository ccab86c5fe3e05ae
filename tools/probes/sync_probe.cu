// Micro-probe: cost of a grid-wide barrier (cooperative groups) and of a cluster barrier on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sync_probe tools/probes/sync_probe.cu && /tmp/sync_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void grid_loop(int iters, unsigned long long *sink)
{
    cg::grid_group g = cg::this_grid();
    unsigned long long acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        g.sync();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

__global__ void cluster_loop(int iters, unsigned long long *sink)
{
    cg::cluster_group c = cg::this_cluster();
    unsigned long long acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        c.sync();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

int main()
{
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2000;
    for (int blocks : {37, 74, 148, 296, 592}) {
        for (int threads : {128, 256}) {
            int it = iters;
            void *args[] = {&it, &sink};
            cudaLaunchCooperativeKernel((void *)grid_loop, dim3(blocks), dim3(threads), args, 0, 0);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel((void *)grid_loop, dim3(blocks), dim3(threads), args, 0, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("grid.sync  blocks=%4d threads=%3d : %.3f us per barrier (%s)\n", blocks, threads, ms * 1e3 / iters, cudaGetErrorString(e));
        }
    }
    for (int csize : {2, 4, 8, 16}) {
        for (int threads : {256, 1024}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize);
            cfg.blockDim = dim3(threads);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (csize > 8) cudaFuncSetAttribute(cluster_loop, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            int it = iters;
            cudaLaunchKernelEx(&cfg, cluster_loop, it, sink);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_loop, it, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("cluster.sync size=%2d threads=%4d : %.3f us per barrier (%s)\n", csize, threads, ms * 1e3 / iters, cudaGetErrorString(e));
        }
    }
    return 0;
}
