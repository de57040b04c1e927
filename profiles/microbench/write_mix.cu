// write_mix.cu -- what HBM gives a streaming kernel with the front end's traffic mix on B200: 1 byte read and 4 bytes
// written per sample (K1), against a plain copy (the MEASURED_PEAKS.json figure: equal reads and writes), a pure write
// and a pure read.  Grid-stride, 128-bit accesses, buffers far larger than the 126 MB L2.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_mix write_mix.cu && ./write_mix
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mix(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_in) {   // 16 B in -> 64 B out
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i];
        out[4 * i + 0] = make_uint4(v.x, v.x + 1, v.x + 2, v.x + 3);
        out[4 * i + 1] = make_uint4(v.y, v.y + 1, v.y + 2, v.y + 3);
        out[4 * i + 2] = make_uint4(v.z, v.z + 1, v.z + 2, v.z + 3);
        out[4 * i + 3] = make_uint4(v.w, v.w + 1, v.w + 2, v.w + 3);
    }
}
// the same traffic, but every store instruction of a warp writes 512 contiguous bytes (whole sectors): word j of the
// warp's 128 output words comes from lane j / 4
__global__ void k_mix_coalesced(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_in) {
    const int lane = threadIdx.x & 31;
    for (size_t i0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) & ~(size_t)31; i0 < n_in; i0 += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i0 + lane];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = k * 32 + lane, src = j >> 2, c = j & 3;
            const uint32_t a = __shfl_sync(0xFFFFFFFFu, v.x, src), b = __shfl_sync(0xFFFFFFFFu, v.y, src);
            const uint32_t cc = __shfl_sync(0xFFFFFFFFu, v.z, src), d = __shfl_sync(0xFFFFFFFFu, v.w, src);
            const uint32_t e = c == 0 ? a : c == 1 ? b : c == 2 ? cc : d;
            out[4 * i0 + j] = make_uint4(e, e + 1, e + 2, e + 3);
        }
    }
}
__global__ void k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void k_write(uint4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_uint4((uint32_t)i, 1, 2, 3);
}
__global__ void k_read(const uint4* __restrict__ in, uint32_t* sink, size_t n) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <class F>
static float time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t samples = 1024ull * 1024 * 1024 * 3;        // configs[3]: 3.22 G samples
    const size_t n_in = samples / 16;                        // uint4 of pixels
    uint4 *in, *out; uint32_t* sink;
    if (cudaMalloc(&in, samples) != cudaSuccess || cudaMalloc(&out, 4 * samples) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 4);
    cudaMemset(in, 1, samples); cudaMemset(out, 0, 4 * samples);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int per_sm : {4, 8, 16}) {
        const int grid = sms * per_sm, block = 256;
        const float mix = time_ms([&] { k_mix<<<grid, block>>>(in, out, n_in); }, 5);
        const float mixc = time_ms([&] { k_mix_coalesced<<<grid, block>>>(in, out, n_in); }, 5);
        printf("grid %d x %d: mix 1R:4W with coalesced stores %.3f ms = %.0f GB/s\n", grid, block, mixc, 5.0 * samples / mixc / 1e6);
        const float cpy = time_ms([&] { k_copy<<<grid, block>>>(out, out + 2 * n_in, 2 * n_in); }, 5);   // 6.4 GB -> 6.4 GB
        const float wr = time_ms([&] { k_write<<<grid, block>>>(out, 4 * n_in); }, 5);
        const float rd = time_ms([&] { k_read<<<grid, block>>>(out, sink, 4 * n_in); }, 5);
        printf("grid %d x %d: mix 1R:4W %.3f ms = %.0f GB/s | copy %.3f ms = %.0f GB/s | write %.3f ms = %.0f GB/s | read %.3f ms = %.0f GB/s\n",
               grid, block, mix, 5.0 * samples / mix / 1e6, cpy, 2.0 * 2 * n_in * 16 / cpy / 1e6, wr, 4.0 * samples / wr / 1e6,
               rd, 4.0 * samples / rd / 1e6);
    }
    printf("cudaMemsetAsync 12.9 GB: ");
    const float ms = time_ms([&] { cudaMemsetAsync(out, 0, 4 * samples); }, 3);
    printf("%.3f ms = %.0f GB/s\n", ms, 4.0 * samples / ms / 1e6);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
