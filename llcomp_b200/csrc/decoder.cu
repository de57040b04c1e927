// decoder.cu -- K5: one adaptive range decoder per slice.
//
// Re-creates llcomp::decompressImage after its header parse (/root/reference/llcomp.hpp:475-545):
// RangeDecoder (:91-127) with zero fill past the end of the slice (:475-479), getSymbol (:219-247),
// the adaptive bit model (:283-293), neighbour fetch / hash / predictor identical to the encoder
// (:494-509), sign unfold (:511-515, :526-528), reconstruction into the 3-row int16 ring (:483, :529)
// and the inverse colour transform with clamp (:532-543).  channels < 3 skips the transform, mirroring
// the encoder (:410-414) -- the reference decoder reads out of bounds there (SURVEY.md defect D2).
//
// In decode the left neighbour is the sample just reconstructed, so context, bin decode and state
// update form one serial chain per slice: one CTA (one warp) per slice, lane 0 runs the chain with the
// 63,408-byte state and (when they fit) the three rows in shared memory.
#include "common.cuh"
#include "kernels.cuh"

namespace llc {

__constant__ ModelTables c_tables_dec = make_tables();

constexpr int kDecBaseSmem = kStateBytes + 128 * 4;
constexpr int kDecMaxLineSmem = 96 * 1024;                   // rows beyond this live in global scratch

__device__ __forceinline__ int dq11(int x) {
    const int a = abs(x);
    const int q = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35);
    return x < 0 ? -q : q;
}
__device__ __forceinline__ int dq5(int x) {
    const int a = abs(x);
    const int q = (a >= 1) + (a >= 4);
    return x < 0 ? -q : q;
}

struct RangeDec {
    uint32_t low, range;
    const uint8_t* p;
    uint32_t pos, len;
    __device__ __forceinline__ uint32_t next_byte() {        // llcomp.hpp:475-479
        const uint32_t b = pos < len ? (uint32_t)__ldg(p + pos) : 0u;
        ++pos;
        return b;
    }
    __device__ __forceinline__ void init(const uint8_t* src, uint32_t n) {   // llcomp.hpp:93-96
        p = src; pos = 0; len = n; range = 0xFF00u;
        low = next_byte() << 8;
        low |= next_byte();
    }
    __device__ __forceinline__ uint32_t get(uint32_t prob) { // llcomp.hpp:106-121
        const uint32_t r1 = (range * prob) >> 8;
        range -= r1;
        uint32_t bit = 0;
        if (low >= range) { low -= range; range = r1; bit = 1; }
        if (range < 0x100u) { range <<= 8; low = (low << 8) + next_byte(); }   // :98-104
        return bit;
    }
};

template <bool kSmemLines>
__global__ void __launch_bounds__(32) k_slice_decoder(const uint8_t* __restrict__ payload,
                                                      const uint64_t* __restrict__ offsets, Geom g,
                                                      uint8_t* __restrict__ pixels,
                                                      int16_t* __restrict__ line_scratch,
                                                      int* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* state = smem;
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + kStateBytes);

    const int lane = threadIdx.x;
    const uint64_t s = blockIdx.x;
    const Slice sl = slice_of(g, s);
    const int C = g.C;
    const int stride = sl.w * C;
    // every slice gets room for a nominal-width tile
    const size_t line_room = (size_t)3 * min(g.tw, g.W) * C;
    int16_t* lines = kSmemLines ? reinterpret_cast<int16_t*>(smem + kDecBaseSmem)
                                : line_scratch + s * line_room;

    for (int i = lane; i < kStateBytes / 16; i += 32) reinterpret_cast<uint4*>(state)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 128; i += 32) tab[i] = c_tables_dec.entry[i];
    __syncwarp();
    if (lane != 0) return;

    RangeDec dec;
    dec.init(payload + offsets[s], (uint32_t)(offsets[s + 1] - offsets[s]));

    const size_t pitch = (size_t)g.W * C;
    uint8_t* out0 = pixels + (size_t)sl.img * g.H * pitch + (size_t)sl.y0 * pitch + (size_t)sl.x0 * C;

    for (int h = 0; h < sl.h; ++h) {
        int16_t* r0 = lines + (size_t)(h % 3) * stride;                     // llcomp.hpp:487-489
        const int16_t* r1 = lines + (size_t)((h + 2) % 3) * stride;
        const int16_t* r2 = lines + (size_t)((h + 1) % 3) * stride;
        uint8_t* dst = out0 + (size_t)h * pitch;
        for (int w = 0; w < sl.w; ++w) {
            const int x = w * C;
            for (int i = 0; i < C; ++i) {
                const int l = w > 0 ? r0[x - C + i] : (h > 0 ? r1[x + i] : 128);       // :494-499
                const int t = h > 0 ? r1[x + i] : l;
                const int L = w > 1 ? r0[x - 2 * C + i] : l;
                const int tl = (h > 0 && w > 0) ? r1[x - C + i] : t;
                const int tr = (h > 0 && w < sl.w - 1) ? r1[x + C + i] : t;
                const int T = h > 1 ? r2[x + i] : t;
                int hash = dq11(l - tl) + 11 * dq11(tl - t) + 121 * dq11(t - tr) + 605 * dq5(L - l) +
                           3025 * dq5(T - t);                                            // :501-507
                const int predict = max(min(l, l + t - tl), min(max(l, l + t - tl), t));  // median, :509
                const bool neg = hash < 0;                                               // :511-515
                if (neg) hash = -hash;

                uint64_t* rowp = reinterpret_cast<uint64_t*>(state + hash * kSubstates);
                uint64_t row = *rowp;
                auto bin = [&](int ctx) -> uint32_t {                                     // :517-523
                    const int sh = ctx * 8;
                    const uint32_t st = (uint32_t)(row >> sh) & 0xFFu;
                    const uint32_t e = tab[st];
                    const uint32_t bit = dec.get(e & 0xFFu);
                    const uint32_t ns = (bit == (st & 1u)) ? (e >> 8) & 0xFFu : (e >> 16) & 0xFFu;
                    row ^= (uint64_t)(st ^ ns) << sh;
                    return bit;
                };

                int diff = 0;
                if (!bin(0)) {                                                           // :225
                    int e = 0, ctx = 1;
                    while (bin(min(ctx++, kELim))) {                                      // :230-235
                        if (++e > 31) {
                            atomicCAS(status, kDevOk, kDevBadExponent);
                            return;
                        }
                    }
                    uint32_t value = 1;
                    ctx = kELim + 1;
                    for (int k = e - 1; k >= 0; --k) value += value + bin(min(ctx++, kRLim));   // :237-240
                    diff = bin(kSignCtx) ? -(int)value : (int)value;                      // :242-245
                }
                *rowp = row;
                if (neg) diff = -diff;                                                    // :526-528
                r0[x + i] = (int16_t)(predict + diff);                                    // :529
            }
            if (C >= 3) {                                                                 // :532-543
                int r = r0[x], gg = r0[x + 1], b = r0[x + 2];
                gg -= (r + b) / 4;
                r += gg;
                b += gg;
                dst[x + 0] = (uint8_t)max(0, min(255, r));
                dst[x + 1] = (uint8_t)max(0, min(255, gg));
                dst[x + 2] = (uint8_t)max(0, min(255, b));
                for (int i = 3; i < C; ++i) dst[x + i] = (uint8_t)r0[x + i];
            } else {
                for (int i = 0; i < C; ++i) dst[x + i] = (uint8_t)r0[x + i];
            }
        }
    }
}

static bool lines_fit_smem(const Geom& g) {
    return (uint64_t)3 * min(g.tw, g.W) * g.C * 2 <= (uint64_t)kDecMaxLineSmem;
}

uint64_t decoder_line_scratch_bytes(const Geom& g) {
    if (lines_fit_smem(g)) return 0;
    return g.n_slices() * 3ull * min(g.tw, g.W) * g.C * 2ull;
}

cudaError_t configure_slice_decoder() {
    cudaError_t e = cudaFuncSetAttribute(k_slice_decoder<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kDecBaseSmem + kDecMaxLineSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_slice_decoder<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecBaseSmem);
}

cudaError_t launch_slice_decoder(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                 uint8_t* d_pixels, int16_t* d_line_scratch, int* d_status, cudaStream_t st) {
    const uint64_t ns = g.n_slices();
    if (ns == 0 || ns > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    if (lines_fit_smem(g)) {
        const int line_bytes = (3 * min(g.tw, g.W) * g.C * 2 + 15) & ~15;
        k_slice_decoder<true><<<(unsigned)ns, 32, kDecBaseSmem + line_bytes, st>>>(d_payload, d_offsets, g, d_pixels,
                                                                                  nullptr, d_status);
    } else {
        k_slice_decoder<false><<<(unsigned)ns, 32, kDecBaseSmem, st>>>(d_payload, d_offsets, g, d_pixels,
                                                                       d_line_scratch, d_status);
    }
    return cudaGetLastError();
}

}  // namespace llc
