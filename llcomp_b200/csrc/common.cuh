// common.cuh -- constants, model tables and slice geometry shared by every kernel.
//
// Format constants follow /root/reference/llcomp.hpp:17-32; the adaptive bit model follows
// llcomp.hpp:250-294 (nextStateMps / nextStateLps / stateProbability).  The tables are generated
// from their pair structure and checked entry-by-entry against the oracle in tests/test_abi.py (test_model_tables_equal_oracle)
// (through llcomp_b200_debug_table).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace llc {

constexpr int kSubstates = 8;                      // llcomp.hpp:25
constexpr int kContexts = 7926;                    // reachable |hash| values: 0..7925
constexpr int kStateBytes = kContexts * kSubstates;  // 63,408 B of adaptive state per slice
constexpr int kELim = 4, kRLim = 6, kSignCtx = 7;  // llcomp.hpp:22-24

// Record written by the front end, one per sample.
__host__ __device__ __forceinline__ uint32_t pack_symbol(int hash, int diff) {
    return ((uint32_t)hash << 11) | ((uint32_t)diff & 0x7FFu);
}

// ---- model tables -------------------------------------------------------------------------
// One u32 per state: P(bit=1)*256 | next-if-MPS << 8 | next-if-LPS << 16.
struct ModelTables {
    uint32_t entry[128];
};

constexpr uint8_t kEvenProb[64] = {
    123, 117, 111, 106, 101, 96, 91, 87, 83, 79, 75, 72, 68, 66, 63, 60, 57, 54, 52, 49, 48, 45,
    43,  41,  40,  38,  36,  35, 33, 32, 30, 30, 28, 27, 26, 25, 24, 23, 22, 21, 21, 20, 19, 18,
    18,  17,  17,  16,  16,  15, 15, 14, 14, 13, 13, 13, 12, 12, 12, 11, 11, 11, 11, 7};
constexpr uint8_t kLpsPair[64] = {0,  0,  1,  2,  2,  4,  4,  5,  6,  7,  8,  9,  9,  11, 11, 12,
                                  13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
                                  24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33,
                                  33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 38, 38, 38, 38, 39};

constexpr ModelTables make_tables() {
    ModelTables t{};
    for (int s = 0; s < 128; ++s) {
        const uint32_t p = (s & 1) ? 254u - kEvenProb[s >> 1] : kEvenProb[s >> 1];
        const uint32_t mps = s < 126 ? s + 2 : s;
        const uint32_t lps = s < 2 ? (s ^ 1) : 2u * kLpsPair[s >> 1] + (s & 1);
        t.entry[s] = p | (mps << 8) | (lps << 16);
    }
    return t;
}

// ---- geometry -----------------------------------------------------------------------------
struct Geom {
    int W, H, C;         // image
    int tw, th;          // nominal tile
    int tiles_x, tiles_y;
    int n_images;
    __host__ __device__ uint32_t slices_per_image() const { return (uint32_t)tiles_x * tiles_y; }
    __host__ __device__ uint64_t n_slices() const { return (uint64_t)slices_per_image() * n_images; }
    __host__ __device__ uint64_t image_samples() const { return (uint64_t)W * H * C; }
    __host__ __device__ uint64_t n_samples() const { return image_samples() * n_images; }
};

struct Slice {
    int img, x0, y0, w, h;   // tile rectangle inside image `img`
    uint64_t sym_off;        // first record of the slice in the slice-major symbol array
    uint64_t n;              // samples in the slice
};

__host__ __device__ __forceinline__ Slice slice_of(const Geom& g, uint64_t s) {
    Slice r;
    const uint32_t spi = g.slices_per_image();
    r.img = (int)(s / spi);
    const uint32_t k = (uint32_t)(s % spi);
    const int ty = k / g.tiles_x, tx = k % g.tiles_x;
    r.x0 = tx * g.tw;
    r.y0 = ty * g.th;
    r.w = min(g.tw, g.W - r.x0);
    r.h = min(g.th, g.H - r.y0);
    r.sym_off = (uint64_t)r.img * g.image_samples() +
                ((uint64_t)r.y0 * g.W + (uint64_t)r.x0 * r.h) * g.C;
    r.n = (uint64_t)r.w * r.h * g.C;
    return r;
}

// Scratch given to slice s by the coder: room for 2x the raw size plus slack (uniform noise codes to
// ~1.25x raw); a slice that still outgrows it raises LLCOMP_ERR_OVERFLOW instead of the reference's
// heap overflow (llcomp.hpp:362).
constexpr uint64_t kScratchSlack = 384;   // >= one refill block of the range pass (256 decisions) + finish()
__host__ __device__ __forceinline__ uint64_t scratch_off(const Slice& sl, uint64_t s) {
    return 2 * sl.sym_off + kScratchSlack * s;
}
__host__ __device__ __forceinline__ uint64_t scratch_cap(const Slice& sl) {
    return 2 * sl.n + kScratchSlack;
}

// Bin queue between the model pass and the range pass: slack entries after every slice (alignment, tail vector).
constexpr int kQueuePad = 64;

// Device-side status word: first error wins.
enum : int { kDevOk = 0, kDevOverflow = 5, kDevBadExponent = 2 };

}  // namespace llc
