// frontend.cu -- K1: pixels -> one (hash, diff) record per sample.
//
// Re-creates the per-sample stage of llcomp::compressImage (/root/reference/llcomp.hpp:390-436):
// reversible colour transform (:396-414), neighbour fetch with border substitution (:417-422),
// quantised 5-neighbour context hash (:424-429, quantisers :297-341), median predictor (:430, :343),
// residual and sign fold (:431-436).  Every neighbour is an ORIGINAL (post-transform) value, so all
// samples are independent and the stage is a pure streaming kernel: 1 byte read, one 4-byte record
// written per sample.
#include "common.cuh"
#include "kernels.cuh"

namespace llc {

// Closed forms of quant11_table / quant5_table (llcomp.hpp:297-341).  Saturating at |x| >= 35 (resp. 4)
// makes the clamp to [-128,127] of the reference redundant.
__device__ __forceinline__ int quant11(int x) {
    const int a = abs(x);
    const int q = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35);
    return x < 0 ? -q : q;
}
__device__ __forceinline__ int quant5(int x) {
    const int a = abs(x);
    const int q = (a >= 1) + (a >= 4);
    return x < 0 ? -q : q;
}
__device__ __forceinline__ int median3(int a, int b, int c) {
    return max(min(a, b), min(max(a, b), c));
}

// Value of plane i of the pixel at p (llcomp.hpp:396-414): planes 0..2 of a >=3-channel image are
// (R-G, G + trunc((B-G + R-G)/4), B-G); everything else is the raw byte.
template <int CT>
__device__ __forceinline__ int plane_value(const uint8_t* __restrict__ p, int C, int i) {
    if ((CT >= 3 || (CT == 0 && C >= 3)) && i < 3) {
        const int g = p[1];
        const int r = (int)p[0] - g, b = (int)p[2] - g;
        if (i == 0) return r;
        if (i == 2) return b;
        return g + (b + r) / 4;   // C++ division: truncates toward zero (:402)
    }
    return p[i];
}

// v1 mapping: one thread per pixel, neighbours re-read through L1/L2.
template <int CT>
__global__ void __launch_bounds__(256) k_frontend_simple(const uint8_t* __restrict__ pixels, Geom g,
                                                         uint32_t* __restrict__ sym,
                                                         unsigned long long* __restrict__ slice_bins) {
    const int C = CT ? CT : g.C;
    const int xi = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int img = blockIdx.z;
    const bool valid = xi < g.W;
    const int x = valid ? xi : g.W - 1;          // out-of-range lanes shadow the last pixel and store nothing

    const int tx = x / g.tw, ty = y / g.th;
    const int x0 = tx * g.tw, y0 = ty * g.th;
    const int w = x - x0, h = y - y0;
    const int sw = min(g.tw, g.W - x0), sh = min(g.th, g.H - y0);

    const size_t pitch = (size_t)g.W * C;
    const uint8_t* row0 = pixels + (size_t)img * g.H * pitch + (size_t)y * pitch;
    const uint8_t* p = row0 + (size_t)x * C;
    uint32_t* out = sym + (size_t)img * g.image_samples() + ((size_t)y0 * g.W + (size_t)x0 * sh) * C +
                    ((size_t)h * sw + w) * C;
    unsigned bins = 0;

#pragma unroll
    for (int i = 0; i < (CT ? CT : C); ++i) {
        const int cur = plane_value<CT>(p, C, i);
        // Border substitution chain of llcomp.hpp:417-422.
        const int l = w > 0 ? plane_value<CT>(p - C, C, i) : (h > 0 ? plane_value<CT>(p - pitch, C, i) : 128);
        const int t = h > 0 ? plane_value<CT>(p - pitch, C, i) : l;
        const int L = w > 1 ? plane_value<CT>(p - 2 * C, C, i) : l;
        const int tl = (h > 0 && w > 0) ? plane_value<CT>(p - pitch - C, C, i) : t;
        const int tr = (h > 0 && w < sw - 1) ? plane_value<CT>(p - pitch + C, C, i) : t;
        const int T = h > 1 ? plane_value<CT>(p - 2 * pitch, C, i) : t;

        int hash = quant11(l - tl) + 11 * quant11(tl - t) + 121 * quant11(t - tr) + 605 * quant5(L - l) +
                   3025 * quant5(T - t);                       // :424-429 (aliased multipliers are normative)
        int diff = cur - median3(l, l + t - tl, t);            // :430-431
        if (hash < 0) { hash = -hash; diff = -diff; }          // :433-436
        if (valid) out[i] = pack_symbol(hash, diff);
        // decisions putSymbol will emit for this residual (llcomp.hpp:183-204): 1, or 2*ilog2|d|+3
        bins += diff ? 2u * (31 - __clz(abs(diff))) + 3u : 1u;
    }

    if (slice_bins) {
        // one counter per slice; a warp that lies inside one tile adds once
        const unsigned slice = (unsigned)img * g.slices_per_image() + ty * g.tiles_x + tx;
        const unsigned mine = valid ? bins : 0u;
        const unsigned first = __shfl_sync(0xFFFFFFFFu, slice, 0);
        if (__all_sync(0xFFFFFFFFu, slice == first)) {
            const unsigned sum = __reduce_add_sync(0xFFFFFFFFu, mine);
            if ((threadIdx.x & 31) == 0 && sum) atomicAdd(slice_bins + slice, (unsigned long long)sum);
        } else if (mine) {
            atomicAdd(slice_bins + slice, (unsigned long long)mine);
        }
    }
}

cudaError_t launch_frontend(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, unsigned long long* d_slice_bins,
                            cudaStream_t st) {
    dim3 block(256);
    dim3 grid((g.W + 255) / 256, g.H, g.n_images);
    if (g.H > 65535 || g.n_images > 65535) return cudaErrorInvalidValue;
    switch (g.C) {
        case 1: k_frontend_simple<1><<<grid, block, 0, st>>>(d_pixels, g, d_sym, d_slice_bins); break;
        case 2: k_frontend_simple<2><<<grid, block, 0, st>>>(d_pixels, g, d_sym, d_slice_bins); break;
        case 3: k_frontend_simple<3><<<grid, block, 0, st>>>(d_pixels, g, d_sym, d_slice_bins); break;
        case 4: k_frontend_simple<4><<<grid, block, 0, st>>>(d_pixels, g, d_sym, d_slice_bins); break;
        default: k_frontend_simple<0><<<grid, block, 0, st>>>(d_pixels, g, d_sym, d_slice_bins); break;
    }
    return cudaGetLastError();
}

}  // namespace llc
