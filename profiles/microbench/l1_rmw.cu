// Microbenchmark: latency of a dependent read-modify-write chain on 8-byte rows, one thread per block,
// rows in (a) shared memory, (b) global memory behind L1.  Answers: do global stores keep the L1 line valid?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_rmw l1_rmw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_smem(int iters, int rows, unsigned long long* out) {
    extern __shared__ uint2 srows[];
    for (int i = 0; i < rows; ++i) srows[i] = make_uint2(i * 7 + 1, i);
    unsigned idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint2 r = srows[idx];
        r.x = r.x * 1664525u + 1013904223u;
        srows[idx] = r;
        idx = (r.x >> 8) % rows;
    }
    long long t1 = clock64();
    out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = idx;
}

__global__ void k_gmem(int iters, int rows, uint2* g, unsigned long long* out) {
    uint2* grow = g + (size_t)blockIdx.x * rows;
    for (int i = 0; i < rows; ++i) grow[i] = make_uint2(i * 7 + 1, i);
    __threadfence_block();
    unsigned idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint2 r = grow[idx];
        r.x = r.x * 1664525u + 1013904223u;
        grow[idx] = r;
        idx = (r.x >> 8) % rows;
    }
    long long t1 = clock64();
    out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = idx;
}

int main() {
    unsigned long long* d_out; uint2* d_g;
    const int blocks = 148 * 7;
    cudaMalloc(&d_out, blocks * 16); cudaMalloc(&d_g, (size_t)blocks * 8192 * 8);
    unsigned long long h[2];
    const int iters = 200000;
    for (int rows : {1, 64, 512, 2048, 7926}) {
        if (rows * 8 <= 64 * 1024) {
            cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            k_smem<<<1, 1, rows * 8>>>(iters, rows, d_out);
            cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
            printf("smem  rows=%5d  1 block : %.1f cycles/iter\n", rows, (double)h[0] / iters);
        }
        for (int nb : {1, 148, 148 * 7}) {
            k_gmem<<<nb, 1>>>(iters, rows, d_g, d_out);
            cudaError_t e = cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
            printf("gmem  rows=%5d  %4d blocks: %.1f cycles/iter %s\n", rows, nb, (double)h[0] / iters, e ? cudaGetErrorString(e) : "");
        }
    }
    return 0;
}
