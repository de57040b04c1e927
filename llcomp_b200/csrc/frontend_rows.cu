// frontend_rows.cu -- K1, streaming form for 3- and 4-channel images: pixels -> one (hash, diff) record per sample.
//
// Same arithmetic as frontend.cu (/root/reference/llcomp.hpp:390-436: colour transform :396-414, neighbours and
// border rules :417-422, context hash :424-429, median predictor :430, residual and sign fold :431-436); what
// changes is how the bytes move:
//   * a CTA takes a region of kFrThreads*PX pixels x kFrRows rows.  Its raw rows (+2 rows above, +16 bytes either
//     side) are brought HBM -> shared memory by the TMA unit, one cp.async.bulk per row, completion on an mbarrier
//     (UBLKCP in SASS); no thread touches a pixel in global memory;
//   * a thread owns PX consecutive pixels and walks DOWN the rows: the planes of row y-1 and the row y-2 term of
//     the hash stay in registers, so a row costs each thread a few 32-bit shared loads of raw bytes, one colour
//     transform per pixel it sees, and per sample five byte look-ups in the two quantiser tables;
//   * the PX*C records of a thread-row are PX*C*4 contiguous bytes, written with 128-bit stores.
// Rows h < 2 of a slice take the substitution chain of llcomp.hpp:417-422 (a warp-uniform branch); the first two and
// the last pixel of a slice row are patched with selects.
// Needs W*C % 16 == 0, tile_w % 4 == 0 and a 16-byte aligned pixel pointer; everything else uses frontend.cu.
#include <cstdint>

#include "common.cuh"
#include "kernels.cuh"

namespace llc {

namespace {

constexpr int kFrThreads = 128;
constexpr int kFrRows = 32;
constexpr int kQB = 640;                                    // plane differences lie in [-637, 637]

struct QuantBytes {
    int8_t q11[2 * kQB];
    int8_t q5[2 * kQB];
};
constexpr QuantBytes make_quant_bytes() {
    QuantBytes t{};
    for (int i = 0; i < 2 * kQB; ++i) {
        const int x = i - kQB, a = x < 0 ? -x : x;
        const int m11 = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35), m5 = (a >= 1) + (a >= 4);
        t.q11[i] = (int8_t)(x < 0 ? -m11 : m11);              // quant11_table, llcomp.hpp:316-333 (closed form)
        t.q5[i] = (int8_t)(x < 0 ? -m5 : m5);                 // quant5_table,  llcomp.hpp:297-314
    }
    return t;
}
__constant__ QuantBytes c_quant_bytes = make_quant_bytes();

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread for a bounded time by itself; a copy that never completes (it cannot: the launcher
    // checks sizes and alignment) traps instead of hanging the device
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}
// HBM -> shared memory through the TMA unit (1-D bulk copy); src, dst and bytes are multiples of 16.
__device__ __forceinline__ void tma_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// One sample from its seven plane values (all neighbours already substituted): llcomp.hpp:424-436.
// d2 is taken as t - tl, so that the predictor's l + t - tl is l + d2; q11 is odd, hence the -11.
__device__ __forceinline__ uint32_t code_sample(int cur, int l, int L, int tl, int t, int tr, int n5,
                                                const int8_t* __restrict__ q11, const int8_t* __restrict__ q5) {
    const int d1 = l - tl, d2 = t - tl, d3 = t - tr, d4 = L - l;
    int hash = (q11[d3] * 11 - q11[d2]) * 11 + q11[d1] + 605 * q5[d4] + n5;    // :424-429 (n5 = 3025 q5(T - t))
    const int hi = max(l, t), lo = min(l, t);
    const int pred = max(min(l + d2, hi), lo);                                    // median(l, l+t-tl, t), :430
    int diff = cur - pred;                                                        // :431
    const int s = hash >> 31;                                                     // :433-436
    hash = abs(hash);
    diff = (diff ^ s) - s;
    return ((uint32_t)hash << 11) | ((uint32_t)diff & 0x7FFu);
}

template <int CT>
struct Px {
    int v[CT];
};

// Planes of the pixel whose first byte is local byte k of the word array (llcomp.hpp:396-409).
template <int CT, int NW>
__device__ __forceinline__ Px<CT> planes_at(const uint32_t (&w)[NW], int k) {
    auto byte = [&](int i) -> int { return (int)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu); };
    Px<CT> p;
    const int g = byte(k + 1);
    const int r = byte(k) - g, b = byte(k + 2) - g;
    p.v[0] = r;
    const int s2 = b + r;                                       // (b + r) / 4, truncating toward zero (:402)
    p.v[1] = g + ((s2 + ((s2 >> 31) & 3)) >> 2);
    p.v[2] = b;
    if (CT == 4) p.v[3] = byte(k + 3);
    return p;
}

template <int CT, int PX>
struct FrShape {
    static constexpr int kRegionW = kFrThreads * PX;           // pixels of a region row
    static constexpr int kRowBytes = kRegionW * CT + 32;       // + one 16-byte chunk either side
    static constexpr int kRawBytes = (kFrRows + 2) * kRowBytes;
    static constexpr int kSmem = kRawBytes + (int)sizeof(QuantBytes) + 16;
    // bytes a thread reads per row: pixels x-2 .. x+PX, starting at the word that holds byte 16 - 2 CT + PX CT tid
    static constexpr int kShift = (16 - 2 * CT) & 3;
    static constexpr int kWords = (kShift + (PX + 3) * CT + 3) / 4;
    static_assert((PX * CT) % 4 == 0, "a thread's pixel group starts on a word");
};

template <int CT, int PX>
__global__ void __launch_bounds__(kFrThreads, 4) k_frontend_rows(const uint8_t* __restrict__ pixels, Geom g,
                                                                 uint32_t* __restrict__ sym) {
    using S = FrShape<CT, PX>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* raw = smem;
    QuantBytes* lut = reinterpret_cast<QuantBytes*>(smem + S::kRawBytes);
    const uint32_t bar = smem_u32(smem + S::kRawBytes + sizeof(QuantBytes));

    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int rx0 = blockIdx.x * S::kRegionW, ry0 = blockIdx.y * kFrRows;
    const size_t pitch = (size_t)g.W * CT;
    const uint8_t* base = pixels + (size_t)img * g.H * pitch;

    // ---- rows ry0-2 .. ry0+kFrRows-1 of the region -> shared memory, by the TMA unit
    const bool pre = rx0 > 0, post = rx0 + S::kRegionW < g.W;
    const uint32_t row_bytes = (uint32_t)(min(S::kRegionW, g.W - rx0) * CT) + (pre ? 16u : 0u) + (post ? 16u : 0u);
    const int r_first = max(0, 2 - ry0);                                   // rows above the image do not exist
    const int r_end = min(kFrRows + 2, g.H - ry0 + 2);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, row_bytes * (uint32_t)(r_end - r_first));
    }
    for (int i = tid; i < (int)(sizeof(QuantBytes) / 16); i += kFrThreads)
        reinterpret_cast<uint4*>(lut)[i] = reinterpret_cast<const uint4*>(&c_quant_bytes)[i];
    __syncthreads();
    if (tid >= r_first && tid < r_end) {
        const uint8_t* src = base + (size_t)(ry0 - 2 + tid) * pitch + (size_t)rx0 * CT - (pre ? 16 : 0);
        tma_row(smem_u32(raw + tid * S::kRowBytes + (pre ? 0 : 16)), src, row_bytes, bar);
    }

    // ---- this thread's column group
    const int x = rx0 + tid * PX;
    const bool active = x < g.W;
    const int tx = min(x, g.W - 1) / g.tw;
    const int x0 = tx * g.tw;
    const int sw = min(g.tw, g.W - x0);
    const int w = x - x0;
    const bool wl = w == 0, wr = w + PX == sw;                             // group at the slice's left / right edge
    int ty = ry0 / g.th;
    int y0 = ty * g.th;
    int sh = min(g.th, g.H - y0);
    int h = ry0 - y0;
    uint32_t* const img_out = sym + (size_t)img * g.image_samples();
    // records of slice row h start at ((y0 W + x0 sh) + h sw + w) C   (32 bits: an image has < 2^32 samples)
    uint32_t out_idx = (((uint32_t)y0 * (uint32_t)g.W + (uint32_t)x0 * (uint32_t)sh) + ((uint32_t)h * (uint32_t)sw + (uint32_t)w)) * CT;

    const int8_t* q11 = lut->q11 + kQB;
    const int8_t* q5 = lut->q5 + kQB;
    const uint32_t* my_row = reinterpret_cast<const uint32_t*>(raw + ((16 - 2 * CT) & ~3) + PX * CT * tid);

    mbar_wait(bar, 0);

    // planes of row y-1 for pixels x-1 .. x+PX (index j+1), and the row y-2 term of the hash for pixels x .. x+PX-1
    Px<CT> top[PX + 2];
    int n5[PX][CT];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j)
#pragma unroll
        for (int c = 0; c < CT; ++c) top[j].v[c] = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int c = 0; c < CT; ++c) n5[j][c] = 0;

#pragma unroll 1
    for (int r = 0; r < kFrRows + 2; ++r) {
        uint32_t wv[S::kWords];
        const uint32_t* rp = my_row + r * (S::kRowBytes / 4);
#pragma unroll
        for (int k = 0; k < S::kWords; ++k) wv[k] = rp[k];
        Px<CT> cur[PX + 3];                                                // pixels x-2 .. x+PX (index j+2)
#pragma unroll
        for (int j = 0; j < PX + 3; ++j) cur[j] = planes_at<CT>(wv, S::kShift + j * CT);

        const int y = ry0 + r - 2;
        if (r >= 2 && y < g.H && active) {
            uint32_t rec[PX * CT];
            if (h >= 2) {
#pragma unroll
                for (int j = 0; j < PX; ++j)
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        // the slice's first two and last columns: llcomp.hpp:417-422 with h > 1
                        int l = cur[j + 1].v[c], L = cur[j].v[c];
                        const int t = top[j + 1].v[c];
                        int tl = top[j].v[c], tr = top[j + 2].v[c];
                        if (j == 0) { l = wl ? t : l; tl = wl ? t : tl; }
                        if (j <= 1) L = wl ? l : L;
                        if (j == PX - 1) tr = wr ? t : tr;
                        rec[j * CT + c] = code_sample(cur[j + 2].v[c], l, L, tl, t, tr, n5[j][c], q11, q5);
                    }
            } else {
#pragma unroll
                for (int j = 0; j < PX; ++j)
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const int wj = w + j;                               // llcomp.hpp:417-422
                        const int top_c = top[j + 1].v[c];
                        const int l = wj > 0 ? cur[j + 1].v[c] : (h > 0 ? top_c : 128);
                        const int t = h > 0 ? top_c : l;
                        const int L = wj > 1 ? cur[j].v[c] : l;
                        const int tl = (h > 0 && wj > 0) ? top[j].v[c] : t;
                        const int tr = (h > 0 && wj < sw - 1) ? top[j + 2].v[c] : t;
                        // T = t on the first two rows of a slice: q5(0) = 0
                        rec[j * CT + c] = code_sample(cur[j + 2].v[c], l, L, tl, t, tr, 0, q11, q5);
                    }
            }
            uint4* out = reinterpret_cast<uint4*>(img_out + out_idx);
#pragma unroll
            for (int k = 0; k < PX * CT / 4; ++k)
                out[k] = make_uint4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
            // next row of the slice, or the first row of the tile row below
            ++h;
            out_idx += (uint32_t)sw * CT;
            if (h == sh) {
                y0 += g.th;
                sh = min(g.th, g.H - y0);
                h = 0;
                out_idx = ((uint32_t)y0 * (uint32_t)g.W + (uint32_t)x0 * (uint32_t)sh + (uint32_t)w) * CT;
            }
        }
        // this row becomes row y-1; its difference to the old row y-1 is the next row's T - t
#pragma unroll
        for (int j = 0; j < PX; ++j)
#pragma unroll
            for (int c = 0; c < CT; ++c) n5[j][c] = 3025 * q5[top[j + 1].v[c] - cur[j + 2].v[c]];
#pragma unroll
        for (int j = 0; j < PX + 2; ++j) top[j] = cur[j + 1];
    }
}

}  // namespace

// true when the streaming kernel can take this geometry (the caller falls back to frontend.cu otherwise)
bool frontend_rows_applicable(const uint8_t* d_pixels, const Geom& g) {
    if (g.C != 3 && g.C != 4) return false;
    if (((size_t)g.W * g.C) % 16 != 0 || g.tw % 4 != 0 || g.W % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(d_pixels) & 15) != 0) return false;
    if (g.image_samples() >= (1ull << 32)) return false;                   // 32-bit record index inside an image
    if ((g.H + kFrRows - 1) / kFrRows > 65535 || g.n_images > 65535) return false;
    return true;
}

cudaError_t configure_frontend_rows() {
    cudaError_t e = cudaFuncSetAttribute(k_frontend_rows<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FrShape<3, 4>::kSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_frontend_rows<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrShape<4, 4>::kSmem);
    return e;
}

cudaError_t launch_frontend_rows(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, cudaStream_t st) {
    if (g.C == 3) {
        using S = FrShape<3, 4>;
        dim3 grid((g.W + S::kRegionW - 1) / S::kRegionW, (g.H + kFrRows - 1) / kFrRows, g.n_images);
        k_frontend_rows<3, 4><<<grid, kFrThreads, S::kSmem, st>>>(d_pixels, g, d_sym);
    } else {
        using S = FrShape<4, 4>;
        dim3 grid((g.W + S::kRegionW - 1) / S::kRegionW, (g.H + kFrRows - 1) / kFrRows, g.n_images);
        k_frontend_rows<4, 4><<<grid, kFrThreads, S::kSmem, st>>>(d_pixels, g, d_sym);
    }
    return cudaGetLastError();
}

}  // namespace llc
