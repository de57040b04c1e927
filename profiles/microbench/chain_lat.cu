// chain_lat.cu -- what one decision of the encoder's range recurrence costs a lone warp on B200, by formulation and by
// how its operands arrive.  One warp per CTA, one CTA per SM; cycles = clock64 around N dependent decisions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chain_lat chain_lat.cu && ./chain_lat
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kN = 1 << 16;          // decisions per measurement
constexpr int kRing = 256;

__device__ __forceinline__ uint32_t step_int(uint32_t y, const uint4& op) {      // coder.cu chain_step, integer form
    const uint32_t nz = y >> 24;
    const uint32_t a = (y >> 8) - 0xFF00u;
    const uint32_t mf = nz * op.y + op.x;
    return a * mf + op.z;
}
__device__ __forceinline__ uint32_t step_fp(uint32_t y, const uint4& op) {       // fp32 form (LLC_CHAIN_V == 2)
    const float x = __uint_as_float(y);
    const float zm = __fadd_rz(x, 2147483648.f);
    const float nz = __saturatef(x - 65535.f);
    const float z = zm - 2147483648.f;
    const float mf = fmaf(nz, __uint_as_float(op.y), __uint_as_float(op.x));
    return __float_as_uint(fmaf(z, mf, __uint_as_float(op.z)));
}

// mode 0: operands from shared memory, one LDS.128 per decision, x stored every 4 (the coder's loop)
// mode 1: the same without the stores     mode 2: operands in registers (the bare recurrence)
// mode 3: x of a group of four stored one decision late     mode 4: four STS.32 instead of one STS.128
// mode 5: x >> 8 (all the byte side needs) packed two per word, one STS.64 per four decisions, stored late
// mode 6: as 3 but two decisions late
template <int FP, int MODE>
__global__ void __launch_bounds__(32) k_chain(uint32_t* out, long long* cycles, uint32_t seed) {
    __shared__ uint4 ring[kRing + 8];
    __shared__ uint4 xs[kRing / 4];
    for (int i = threadIdx.x; i < kRing + 8; i += 32) {
        const uint32_t m = 7 + ((seed + 37u * i) % 241u), bit = (seed >> (i & 15)) & 1u;
        if (FP) ring[i] = make_uint4(__float_as_uint((float)m), __float_as_uint((float)m * -0.99609375f),
                                     __float_as_uint(bit ? 0.f : 255.f), 0u);
        else ring[i] = make_uint4(m << 8, m * 0xFFFFFF01u, (bit ? 0u : 255u) + 0xFF0000u, 0u);
    }
    __syncthreads();
    uint32_t y = FP ? __float_as_uint(16711680.f) : (0xFF00u << 8) + 0xFF0000u;
    const long long t0 = clock64();
    if (MODE == 2) {
        const uint4 a = ring[threadIdx.x & 7], b = ring[8 + (threadIdx.x & 7)];
#pragma unroll 1
        for (int i = 0; i < kN / 8; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                y = FP ? step_fp(y, a) : step_int(y, a);
                y = FP ? step_fp(y, b) : step_int(y, b);
            }
        }
    } else {
#pragma unroll 1
        for (int blk = 0; blk < kN / kRing; ++blk) {
            const uint4* rp = ring;
            uint4* xo = xs;
            uint4 a0 = rp[0], a1 = rp[1], a2 = rp[2], a3 = rp[3];
            uint4 w = make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int v = 0; v < kRing / 8; ++v) {
                const uint4 b0 = rp[4], b1 = rp[5], b2 = rp[6], b3 = rp[7];
                uint4 x;
                x.x = y = FP ? step_fp(y, a0) : step_int(y, a0);
                if (MODE == 3 && v) xo[-1] = w;
                if (MODE == 5 && v) reinterpret_cast<uint2*>(xo)[-1] = make_uint2(__byte_perm(w.x >> 8, w.y >> 8, 0x5410), __byte_perm(w.z >> 8, w.w >> 8, 0x5410));
                x.y = y = FP ? step_fp(y, a1) : step_int(y, a1);
                if (MODE == 6 && v) xo[-1] = w;
                x.z = y = FP ? step_fp(y, a2) : step_int(y, a2);
                x.w = y = FP ? step_fp(y, a3) : step_int(y, a3);
                if (MODE == 0) xo[0] = x;
                if (MODE == 4) { volatile uint32_t* p = reinterpret_cast<volatile uint32_t*>(xo); p[0] = x.x; p[1] = x.y; p[2] = x.z; p[3] = x.w; }
                rp += 8;
                a0 = rp[0]; a1 = rp[1]; a2 = rp[2]; a3 = rp[3];
                w.x = y = FP ? step_fp(y, b0) : step_int(y, b0);
                if (MODE == 3 || MODE == 6 && false) xo[0] = x;
                if (MODE == 5) reinterpret_cast<uint2*>(xo)[0] = make_uint2(__byte_perm(x.x >> 8, x.y >> 8, 0x5410), __byte_perm(x.z >> 8, x.w >> 8, 0x5410));
                w.y = y = FP ? step_fp(y, b1) : step_int(y, b1);
                if (MODE == 6) xo[0] = x;
                w.z = y = FP ? step_fp(y, b2) : step_int(y, b2);
                w.w = y = FP ? step_fp(y, b3) : step_int(y, b3);
                if (MODE == 0) xo[1] = w;
                if (MODE == 4) { volatile uint32_t* p = reinterpret_cast<volatile uint32_t*>(xo + 1); p[0] = w.x; p[1] = w.y; p[2] = w.z; p[3] = w.w; }
                xo += 2;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; out[blockIdx.x] = y; }
}

// latency of one dependent instruction kind: 64 of them back to back
template <int OP>
__global__ void __launch_bounds__(32) k_op(uint32_t* out, long long* cycles, uint32_t a, uint32_t b) {
    uint32_t y = a + threadIdx.x;
    float f = (float)y;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == 0) y = y * b + a;                                  // IMAD
            if (OP == 1) y = (y >> 3) ^ 0x5u;                            // SHF + LOP3 (two ALU ops)
            if (OP == 2) y = __umulhi(y, b) + a;                         // IMAD.HI
            if (OP == 3) f = fmaf(f, 1.0001f, 0.5f);                     // FFMA
            if (OP == 4) f = __fadd_rz(f, 3.0f);                         // FADD.RZ
            if (OP == 5) f = __saturatef(f - 0.25f) + 0.0f;              // FADD.SAT (+ FADD)
            if (OP == 6) y = __byte_perm(y, b, 0x3120 + (k & 1));        // PRMT
            if (OP == 7) y = min(y ^ b, 0x7FFFFFFFu);                    // LOP3 + VIMNMX
            if (OP == 8) { y = (y >> 8) * b + a; }                       // SHF -> IMAD (cross pipe)
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; out[blockIdx.x] = y + __float_as_uint(f); }
}

int main() {
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 4096); cudaMalloc(&d_cyc, 4096 * 8);
    long long h[8];
    auto report = [&](const char* name, double per) {
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-58s %7.2f cycles\n", name, (double)h[0] / per);
    };
#define CHAIN(FP, MODE, name) k_chain<FP, MODE><<<4, 32>>>(d_out, d_cyc, 12345u); cudaDeviceSynchronize(); \
    k_chain<FP, MODE><<<4, 32>>>(d_out, d_cyc, 12345u); cudaDeviceSynchronize(); report(name, kN)
    CHAIN(0, 2, "integer recurrence, operands in registers / decision");
    CHAIN(0, 1, "integer recurrence, LDS.128 operand per decision");
    CHAIN(0, 0, "integer recurrence, LDS.128 operand + x stored (coder loop)");
    CHAIN(0, 3, "integer, LDS.128 + x stored one decision late");
    CHAIN(0, 6, "integer, LDS.128 + x stored two decisions late");
    CHAIN(0, 4, "integer, LDS.128 + x stored as four STS.32");
    CHAIN(0, 5, "integer, LDS.128 + (x >> 8) packed, STS.64 per four, late");
    CHAIN(1, 2, "fp32 recurrence, operands in registers / decision");
    CHAIN(1, 1, "fp32 recurrence, LDS.128 operand per decision");
    CHAIN(1, 0, "fp32 recurrence, LDS.128 operand + x stored");
#define OPL(OP, name, n) k_op<OP><<<1, 32>>>(d_out, d_cyc, 3u, 5u); cudaDeviceSynchronize(); report(name, 1024.0 * 16 * n)
    OPL(0, "IMAD dependent", 1);
    OPL(1, "SHF + LOP3 dependent (per op)", 2);
    OPL(2, "IMAD.HI + IADD dependent (per pair)", 1);
    OPL(3, "FFMA dependent", 1);
    OPL(4, "FADD.RZ dependent", 1);
    OPL(5, "FADD.SAT + FADD dependent (per pair)", 1);
    OPL(6, "PRMT dependent", 1);
    OPL(7, "LOP3 + VIMNMX dependent (per pair)", 1);
    OPL(8, "SHF -> IMAD dependent (per pair)", 1);
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
