// How many thread-block clusters of each size can be co-resident on this GPU (GPC packing)?
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k512(int *p) { if (p) p[0] = 1; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs %d smem/SM %zu smem/block optin %zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin);
    cudaFuncSetAttribute(k512, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int threads : {512, 1024}) for (size_t smem : {(size_t)40 * 1024, (size_t)100 * 1024, (size_t)200 * 1024}) {
        cudaFuncSetAttribute(k512, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int cs : {1, 2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k512, &cfg);
            printf("threads %4d smem %3zu KB cluster %2d: max active clusters %d (CTAs %d) %s\n", threads, smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
