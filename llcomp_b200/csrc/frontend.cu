// frontend.cu -- K1: pixels -> one (hash, diff) record per sample.
//
// Re-creates the per-sample stage of llcomp::compressImage (/root/reference/llcomp.hpp:390-436):
// reversible colour transform (:396-414), neighbour fetch with border substitution (:417-422),
// quantised 5-neighbour context hash (:424-429, quantisers :297-341), median predictor (:430, :343),
// residual and sign fold (:431-436).  Every neighbour is an ORIGINAL (post-transform) value, so all
// samples are independent and the stage is a pure streaming kernel: 1 byte read, one 4-byte record
// written per sample.
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------
// Tiled mapping for 3- and 4-channel images: a CTA stages the post-transform planes of an image region
// (kTileH x kTileW pixels plus a 2-row / 2+1-column halo) in shared memory ONCE, as one 16-byte pixel
// (4 x int32, so that a neighbour is one LDS.128 and needs no unpacking), then every thread codes pixels of
// that region from shared memory.  Interior pixels of a slice take a branch-free path with table-driven
// quantisers (premultiplied by their hash weights); pixels on a slice border take the substitution chain of
// llcomp.hpp:417-422.  No integer division in the loops: the region meets at most one slice boundary per axis
// when tiles are at least as large as the region (smaller tiles fall back to the simple kernel).
// ---------------------------------------------------------------------------------------------------
constexpr int kTileW = 64, kTileH = 16;
constexpr int kHaloL = 2, kHaloR = 1, kHaloT = 2;
constexpr int kSmemW = kTileW + kHaloL + kHaloR;           // 67 pixels per staged row
constexpr int kSmemH = kTileH + kHaloT;                    // 18 rows

// Quantisers as byte tables over the whole range a difference of two staged plane values can take (planes lie in
// [-255, 382] after the colour transform, so differences lie in [-637, 637]): no clamp, the bias sits in the
// load's immediate offset, and the weight a term carries in the hash is applied by the multiply-add that sums it.
constexpr int kQBias = 640;
struct QuantLuts {
    int8_t q11[2 * kQBias];   // q11(x), x + kQBias
    int8_t q5[2 * kQBias];    // q5(x)
};
constexpr QuantLuts make_quant_luts() {
    QuantLuts t{};
    for (int i = 0; i < 2 * kQBias; ++i) {
        const int x = i - kQBias, a = x < 0 ? -x : x;
        const int m11 = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35), m5 = (a >= 1) + (a >= 4);
        t.q11[i] = (int8_t)(x < 0 ? -m11 : m11);
        t.q5[i] = (int8_t)(x < 0 ? -m5 : m5);
    }
    return t;
}
__constant__ QuantLuts c_quant_luts = make_quant_luts();
static_assert(sizeof(QuantLuts) % 16 == 0 && sizeof(QuantLuts) / 16 <= 256, "copied by one uint4 per thread");

template <int CT, bool kCount>
__global__ void __launch_bounds__(256, 5) k_frontend_tiled(const uint8_t* __restrict__ pixels, Geom g,
                                                        uint32_t* __restrict__ sym,
                                                        unsigned long long* __restrict__ slice_bins) {
    static_assert(CT == 3 || CT == 4, "staged pixels hold 3 or 4 planes");
    __shared__ int4 tile[kSmemH][kSmemW];
    __shared__ __align__(16) QuantLuts lut;
    __shared__ unsigned int cta_bins;                      // decisions of the region's first slice (<= 1024*4*19)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.z;
    const int rx0 = blockIdx.x * kTileW, ry0 = blockIdx.y * kTileH;     // region origin in the image
    const size_t pitch = (size_t)g.W * CT;
    const uint8_t* base = pixels + (size_t)img * g.H * pitch;

    if (kCount && tid == 0) cta_bins = 0;
    if (tid < (int)(sizeof(QuantLuts) / 16))
        reinterpret_cast<uint4*>(&lut)[tid] = reinterpret_cast<const uint4*>(&c_quant_luts)[tid];
    // stage the planes (llcomp.hpp:396-409): warp w takes rows w, w+8, w+16; coordinates clamped to the image
    // (clamped copies are never used as neighbours of a valid in-slice pixel)
    // (all loads of a thread are issued before the first use: 3 rows x 3 columns, fully unrolled)
    {
        // 32-bit offsets inside the image (an image is at most 16384 x 16384 x 4 bytes): one add per staged pixel
        uint32_t raw[3][3][4] = {};
        uint32_t rowoff[3], coloff[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            rowoff[a] = (uint32_t)min(max(ry0 + warp + 8 * a - kHaloT, 0), g.H - 1) * (uint32_t)pitch;
            coloff[a] = (uint32_t)min(max(rx0 + lane + 32 * a - kHaloL, 0), g.W - 1) * CT;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int sy = warp + 8 * a;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int sx = lane + 32 * b;
                const uint8_t* p = base + (rowoff[a] + coloff[b]);
                if (sy < kSmemH && sx < kSmemW) {
#pragma unroll
                    for (int c = 0; c < CT; ++c) raw[a][b][c] = p[c];
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int sy = warp + 8 * a;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int sx = lane + 32 * b;
                if (sy < kSmemH && sx < kSmemW) {
                    const int gg = (int)raw[a][b][1];
                    const int r = (int)raw[a][b][0] - gg, bb = (int)raw[a][b][2] - gg;
                    tile[sy][sx] = make_int4(r, gg + (bb + r) / 4, bb, CT == 4 ? (int)raw[a][b][3] : 0);
                }
            }
        }
    }
    __syncthreads();

    // slice geometry of this thread's column and of the (at most two) tile rows the region touches
    const int lx = tid & (kTileW - 1);                                  // column inside the region
    const int x = rx0 + lx;
    const int tx = min(x, g.W - 1) / g.tw;
    const int x0 = tx * g.tw;
    const int sw = min(g.tw, g.W - x0);
    const int w = x - x0;
    const int ty_a = ry0 / g.th;                                        // tile row of the region's first row
    const int yb = (ty_a + 1) * g.th;                                   // first image row of the next tile row
    const unsigned primary = (unsigned)img * g.slices_per_image() + ty_a * g.tiles_x + rx0 / g.tw;
    uint32_t* const img_out = sym + (size_t)img * g.image_samples();
    unsigned acc = 0;
#pragma unroll 1
    for (int ly = tid / kTileW; ly < kTileH; ly += 256 / kTileW) {
        const int y = ry0 + ly;
        if (x >= g.W || y >= g.H) continue;
        const int ty = ty_a + (y >= yb);
        const int y0 = y >= yb ? yb : ty_a * g.th;
        const int sh = min(g.th, g.H - y0);
        const int h = y - y0;
        // record index inside the image in 32 bits (an image has at most 2^30 samples)
        uint32_t* out = img_out + (((uint32_t)y0 * (uint32_t)g.W + (uint32_t)x0 * (uint32_t)sh) +
                                   ((uint32_t)h * (uint32_t)sw + (uint32_t)w)) * (uint32_t)CT;
        const int sy = ly + kHaloT, sx = lx + kHaloL;
        const int4 pc = tile[sy][sx];
        int4 pl, pL, ptl, pt, ptr, pT;
        if (w >= 2 && w < sw - 1 && h >= 2) {                           // interior of the slice
            pl = tile[sy][sx - 1]; pL = tile[sy][sx - 2];
            ptl = tile[sy - 1][sx - 1]; pt = tile[sy - 1][sx]; ptr = tile[sy - 1][sx + 1];
            pT = tile[sy - 2][sx];
        } else {                                                        // llcomp.hpp:417-422, all planes at once
            const int4 c128 = make_int4(128, 128, 128, 128);
            pl = w > 0 ? tile[sy][sx - 1] : (h > 0 ? tile[sy - 1][sx] : c128);
            pt = h > 0 ? tile[sy - 1][sx] : pl;
            pL = w > 1 ? tile[sy][sx - 2] : pl;
            ptl = (h > 0 && w > 0) ? tile[sy - 1][sx - 1] : pt;
            ptr = (h > 0 && w < sw - 1) ? tile[sy - 1][sx + 1] : pt;
            pT = h > 1 ? tile[sy - 2][sx] : pt;
        }
        const int cur[4] = {pc.x, pc.y, pc.z, pc.w};
        const int l[4] = {pl.x, pl.y, pl.z, pl.w}, L[4] = {pL.x, pL.y, pL.z, pL.w};
        const int tl[4] = {ptl.x, ptl.y, ptl.z, ptl.w}, t[4] = {pt.x, pt.y, pt.z, pt.w};
        const int tr[4] = {ptr.x, ptr.y, ptr.z, ptr.w}, T[4] = {pT.x, pT.y, pT.z, pT.w};
        unsigned bins = 0;
#pragma unroll
        for (int i = 0; i < CT; ++i) {
            const int8_t* q11 = lut.q11 + kQBias;
            const int8_t* q5 = lut.q5 + kQBias;
            int hash = q11[l[i] - tl[i]] + 11 * q11[tl[i] - t[i]] + 121 * q11[t[i] - tr[i]] +
                       605 * q5[L[i] - l[i]] + 3025 * q5[T[i] - t[i]];                  // :424-429
            int diff = cur[i] - median3(l[i], l[i] + t[i] - tl[i], t[i]);               // :430-431
            if (hash < 0) { hash = -hash; diff = -diff; }                               // :433-436
            out[i] = pack_symbol(hash, diff);
            if (kCount) bins += diff ? 2u * (31 - __clz(abs(diff))) + 3u : 1u;
        }
        if (kCount) {
            const unsigned slice = (unsigned)img * g.slices_per_image() + ty * g.tiles_x + tx;
            if (slice == primary) acc += bins;
            else atomicAdd(slice_bins + slice, (unsigned long long)bins);
        }
    }
    if (kCount) {
        const unsigned sum = __reduce_add_sync(0xFFFFFFFFu, acc);
        if (lane == 0 && sum) atomicAdd(&cta_bins, sum);
        __syncthreads();
        if (tid == 0 && cta_bins) atomicAdd(slice_bins + primary, (unsigned long long)cta_bins);
    }
}

cudaError_t launch_frontend(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, unsigned long long* d_slice_bins,
                            cudaStream_t st) {
    if (g.H > 65535 * kTileH || g.n_images > 65535) return cudaErrorInvalidValue;
    const Switches& sw = switches();
    const bool no_rows = sw.frontend_tiled || sw.frontend_simple;
    if (!d_slice_bins && !no_rows && frontend_rows_applicable(d_pixels, g)) return launch_frontend_rows(d_pixels, g, d_sym, st);
    // the tiled kernel indexes records inside an image with 32 bits
    if (g.image_samples() < (1ull << 32) && (g.C == 3 || g.C == 4) && g.th >= kTileH && !sw.frontend_simple) {
        dim3 grid((g.W + kTileW - 1) / kTileW, (g.H + kTileH - 1) / kTileH, g.n_images);
        if (g.C == 3 && d_slice_bins) k_frontend_tiled<3, true><<<grid, 256, 0, st>>>(d_pixels, g, d_sym, d_slice_bins);
        else if (g.C == 3) k_frontend_tiled<3, false><<<grid, 256, 0, st>>>(d_pixels, g, d_sym, nullptr);
        else if (d_slice_bins) k_frontend_tiled<4, true><<<grid, 256, 0, st>>>(d_pixels, g, d_sym, d_slice_bins);
        else k_frontend_tiled<4, false><<<grid, 256, 0, st>>>(d_pixels, g, d_sym, nullptr);
        return cudaGetLastError();
    }
    dim3 block(256);
    dim3 grid((g.W + 255) / 256, g.H, g.n_images);
    if (g.H > 65535) return cudaErrorInvalidValue;
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
