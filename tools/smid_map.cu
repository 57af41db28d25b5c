// prints blockIdx -> smid for a 148-CTA, 1-CTA-per-SM launch (plain and as clusters of 2)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned *out) {
    extern __shared__ char big[];
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) out[blockIdx.x] = smid;
    big[threadIdx.x] = 1;
    __nanosleep(2000000);
}
int main() {
    unsigned *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<148, 128, 200 * 1024>>>(d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("plain:");
    for (int i = 0; i < 148; ++i) printf(" %u", h[i]);
    printf("\n");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("cluster2:");
    for (int i = 0; i < 148; ++i) printf(" %u", h[i]);
    printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
