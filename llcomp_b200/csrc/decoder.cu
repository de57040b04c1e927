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
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------
// Fast decoder for 1..4 channels.  Still one serial chain per slice (lane 0), but everything that does not
// depend on the sample being decoded is taken out of it and given to the whole warp:
//   * before a row is decoded, the lanes compute for every sample of the row the part of the context hash that
//     only involves the rows above, 11*q11(tl-t) + 121*q11(t-tr) + 3025*q5(T-t) with the border rules of
//     llcomp.hpp:495-499, and store it IN PLACE of the row y-2 value it consumed.  That buffer then receives
//     the reconstructed row y sample by sample, so two row buffers replace the reference's three;
//   * the payload is staged into a shared-memory ring by all lanes (coalesced), the chain reads bytes from it;
//   * after a row is decoded, the lanes do the inverse colour transform and the pixel stores.
// The chain keeps, per plane, l / L / tl in registers (sliding window), fetches the table entries of the
// sub-states a residual will need together, and updates the 8 sub-states of the context row in registers.
// ---------------------------------------------------------------------------------------------------
constexpr int kRing = 512;                      // bytes of payload staged ahead of the chain
constexpr int kChunkSamples = 36;               // samples decoded between two refills (<= 12.4 B each, worst case)
constexpr int kFastBase = kStateBytes + 128 * 4 + kRing;   // + 2 rows: 76,720 B for a 1024-wide RGB tile -> 3 per SM

struct Q11Lut { int8_t v[256]; };
constexpr Q11Lut make_q11lut() {
    Q11Lut t{};
    for (int i = 0; i < 256; ++i) {
        const int x = i - 128, a = x < 0 ? -x : x;
        const int q = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35);
        t.v[i] = (int8_t)(x < 0 ? -q : q);
    }
    return t;
}
__constant__ Q11Lut c_q11lut = make_q11lut();

// kGlobalState: the slice's state rows live in global memory (pre-zeroed by the host) and are reached through
// L1 instead of shared memory.  Measured on B200 (profiles/microbench/l1_rmw.cu): a dependent 8-byte
// read-modify-write chain costs 116 cycles through L1 against 97 in shared memory, and stores keep the L1 line
// valid, as long as the hot rows of the SM's slices fit L1.  Without the 63 KB of shared memory per slice, 7+
// slices fit an SM instead of 3, so a 1024-slice batch decodes in one wave.
template <int CT, bool kGlobalState>
__global__ void __launch_bounds__(32) k_slice_decoder_fast(const uint8_t* __restrict__ payload,
                                                           const uint64_t* __restrict__ offsets, Geom g,
                                                           uint8_t* __restrict__ pixels, int* __restrict__ status,
                                                           uint2* __restrict__ gstate) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kStateSmem = kGlobalState ? 0 : kStateBytes;
    uint2* state = kGlobalState ? gstate + (size_t)blockIdx.x * kContexts : reinterpret_cast<uint2*>(smem);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + kStateSmem);    // P | next_if_0 << 8 | next_if_1 << 16
    const int8_t* q11lut = c_q11lut.v;
    uint8_t* ring = smem + kStateSmem + 512;

    const int lane = threadIdx.x;
    const uint64_t s = blockIdx.x;
    const Slice sl = slice_of(g, s);
    const int stride = sl.w * CT;
    int16_t* bufA = reinterpret_cast<int16_t*>(smem + kStateSmem + 512 + kRing);   // row y-1
    int16_t* bufB = bufA + ((min(g.tw, g.W) * CT + 7) & ~7);            // row y-2 -> hash part -> row y

    if (!kGlobalState)
        for (int i = lane; i < kStateBytes / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 128; i += 32) {
        const uint32_t e = c_tables_dec.entry[i], p = e & 0xFFu, nm = (e >> 8) & 0xFFu, nl = (e >> 16) & 0xFFu;
        const uint32_t mps = i & 1u;                                     // llcomp.hpp:285, :290-292
        tab[i] = p | ((mps == 0 ? nm : nl) << 8) | ((mps == 1 ? nm : nl) << 16);
    }

    const uint8_t* src = payload + offsets[s];
    const uint32_t len = (uint32_t)(offsets[s + 1] - offsets[s]);
    uint32_t filled = 0;                                                 // bytes [0, filled) are (or were) in the ring
    uint32_t pos = 0;                                                    // bytes consumed by the chain (warp-uniform copy)
    auto refill = [&]() {                                                // zero fill past the end, llcomp.hpp:475-479
        const uint32_t want = pos + kRing;
        for (uint32_t k = filled + lane; k < want; k += 32) ring[k & (kRing - 1)] = k < len ? __ldg(src + k) : 0;
        filled = want;
        __syncwarp();
    };
    refill();

    // decoder registers (meaningful in lane 0)
    uint32_t low = 0, range = 0xFF00u;                                   // llcomp.hpp:93-96
    low = ((uint32_t)ring[0] << 8) | ring[1];
    pos = 2;                                                             // ring[pos] is the next unread byte ...
    uint32_t nextb = ring[2];                                            // ... and lane 0 keeps it in a register
    bool bad = false;
    // the row every plane wrote back last (store-to-load forwarding for the rows requested ahead)
    int wr_hash[CT];
    uint2 wr_row[CT];
#pragma unroll
    for (int i = 0; i < CT; ++i) { wr_hash[i] = -1; wr_row[i] = make_uint2(0, 0); }

    const size_t pitch = (size_t)g.W * CT;
    uint8_t* out0 = pixels + (size_t)sl.img * g.H * pitch + (size_t)sl.y0 * pitch + (size_t)sl.x0 * CT;

    for (int h = 0; h < sl.h; ++h) {
        // ---- all lanes: hash part from the rows above, in place of row y-2 (llcomp.hpp:495-507)
        for (int j = lane; j < stride; j += 32) {
            int pre = 0;
            if (h > 0) {
                const int w = j / CT;
                const int t = bufA[j];
                const int tl = w > 0 ? bufA[j - CT] : t;
                const int tr = w < sl.w - 1 ? bufA[j + CT] : t;
                const int T = h > 1 ? bufB[j] : t;
                pre = 11 * dq11(tl - t) + 121 * dq11(t - tr) + 3025 * dq5(T - t);
            }
            bufB[j] = (int16_t)pre;
        }
        __syncwarp();

        // ---- the chain, in chunks so that the ring can be topped up by the whole warp.
        // Software pipeline over pixels: as soon as plane i of pixel w is reconstructed, everything plane i of
        // pixel w+1 needs that does not depend on the bitstream is prepared -- neighbours, context hash
        // (llcomp.hpp:494-507), median prediction (:509) -- and its state row is requested.  The row then has the
        // CT-1 samples in between to arrive (it comes from L1 or L2 when the rows live in global memory); if one
        // of those samples updates the same row, the fresh copy is taken from registers instead (wr_*).
        // (Measured and dropped: requesting the table entries of all 8 sub-states ahead as well, and mask-based
        // selections instead of compare + select in a decision.  One thread issued in order pays ~4.3 cycles per
        // instruction whatever it is, and both variants add instructions: 1722 -> 1950 ms on configs[3].)
        int l[CT], L[CT], tl[CT];
#pragma unroll
        for (int i = 0; i < CT; ++i) { l[i] = 128; L[i] = 128; tl[i] = 0; }
        int nt[CT] = {}, nhash[CT] = {}, npred[CT] = {};
        uint2 nrow[CT] = {};
        auto prepare = [&](int i, int w) {                   // lane 0: plane i of pixel w, from l / L / tl as they stand
            const int j = w * CT + i;
            // neighbours: first row -> t = tl = l; first column -> l = L = tl = t
            const int t = h > 0 ? (int)bufA[j] : l[i];
            if (w == 0 && h > 0) { l[i] = t; L[i] = t; tl[i] = t; }
            const int tli = h > 0 ? tl[i] : l[i];
            nhash[i] = (int)bufB[j] + q11lut[max(-128, min(127, l[i] - tli)) + 128] + 605 * dq5(L[i] - l[i]);   // :501-507
            const int lt = l[i] + t - tli;
            npred[i] = max(min(l[i], lt), min(max(l[i], lt), t));                            // median, :509
            nt[i] = t;
            nrow[i] = state[abs(nhash[i])];
        };
        if (lane == 0 && !bad) {
#pragma unroll
            for (int i = 0; i < CT; ++i) prepare(i, 0);
        }
        const int px_per_chunk = kChunkSamples / CT;
        for (int w0 = 0; w0 < sl.w; w0 += px_per_chunk) {
            refill();
            if (lane == 0 && !bad) {
                nextb = ring[pos & (kRing - 1)];                         // the ring may have been topped up
                const int w1 = min(sl.w, w0 + px_per_chunk);
                for (int w = w0; w < w1; ++w) {
                    const int j = w * CT;
#pragma unroll
                    for (int i = 0; i < CT; ++i) {
                        const int t = nt[i], predict = npred[i];
                        const bool neg = nhash[i] < 0;                                      // :511-515
                        const int hash = abs(nhash[i]);
                        uint2 row = nrow[i];
#pragma unroll
                        for (int k = CT - 1; k >= 1; --k) {                                 // updated since requested?
                            const int p = (i + CT - k) % CT;                                // k samples ago; newest last
                            if (hash == wr_hash[p]) row = wr_row[p];
                        }

                        // Table entries of the sub-states a residual can touch once, fetched together so that
                        // their shared-memory latency overlaps (ctx 4 and 6 repeat: fetched when reached).
                        const uint32_t e0 = tab[row.x & 0xFFu], e1 = tab[(row.x >> 8) & 0xFFu];
                        const uint32_t e2 = tab[(row.x >> 16) & 0xFFu], e3 = tab[row.x >> 24];
                        const uint32_t e5 = tab[(row.y >> 8) & 0xFFu], e7 = tab[row.y >> 24];

                        // one decision with table entry e of sub-state byte kB of `half` (llcomp.hpp:106-121, :517-523)
                        auto bin = [&](uint32_t e, uint32_t& half, int kB) -> uint32_t {
                            const uint32_t r1 = (range * (e & 0xFFu)) >> 8;
                            const uint32_t r0v = range - r1;
                            const uint32_t bit = low >= r0v ? 1u : 0u;
                            low = min(low, low - r0v);                   // low - r0v wraps above low when low < r0v
                            range = bit ? r1 : r0v;
                            if (range < 0x100u) {                                            // :98-104
                                range <<= 8;
                                low = (low << 8) + nextb;
                                ++pos;
                                nextb = ring[pos & (kRing - 1)];                             // needed at the NEXT refill
                                asm volatile("" ::: "memory");       // keep this a branch: one decision in ten refills
                            }
                            const uint32_t ns = __byte_perm(e, 0, 0x4441 + bit);
                            half = __byte_perm(half, ns, kB == 0 ? 0x3214 : kB == 1 ? 0x3240 : kB == 2 ? 0x3410 : 0x4210);
                            return bit;
                        };

                        int diff = 0;
                        if (!bin(e0, row.x, 0)) {                                            // :225
                            int e = 0;                                                       // :227-235, ctx min(1+k,4)
                            if (bin(e1, row.x, 1)) {
                                e = 1;
                                if (bin(e2, row.x, 2)) {
                                    e = 2;
                                    if (bin(e3, row.x, 3)) {
                                        e = 3;
                                        while (bin(tab[row.y & 0xFFu], row.y, 0)) {
                                            if (++e > 31) { bad = true; break; }
                                        }
                                    }
                                }
                            }
                            if (bad) break;
                            uint32_t value = 1;                                              // :237-240
                            if (e >= 1) value += value + bin(e5, row.y, 1);
                            for (int k = e - 2; k >= 0; --k) value += value + bin(tab[(row.y >> 16) & 0xFFu], row.y, 2);
                            diff = bin(e7, row.y, 3) ? -(int)value : (int)value;             // :242-245
                        }
                        state[hash] = row;
                        wr_hash[i] = hash; wr_row[i] = row;
                        const int cur = (int16_t)(predict + (neg ? -diff : diff));           // :526-529
                        bufB[j + i] = (int16_t)cur;
                        L[i] = w == 0 ? cur : l[i];                                          // w == 1: L = l (:496)
                        l[i] = cur;
                        tl[i] = t;
                        if (w + 1 < sl.w) prepare(i, w + 1);
                    }
                    if (bad) break;
                }
            }
            bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
            if (bad) {
                if (lane == 0) atomicCAS(status, kDevOk, kDevBadExponent);
                return;
            }
        }
        __syncwarp();

        // ---- all lanes: inverse colour transform, clamp, store (llcomp.hpp:532-543)
        uint8_t* dst = out0 + (size_t)h * pitch;
        for (int w = lane; w < sl.w; w += 32) {
            const int16_t* p = bufB + w * CT;
            if (CT >= 3) {
                int r = p[0], gg = p[1], b = p[2];
                gg -= (r + b) / 4;
                r += gg;
                b += gg;
                dst[w * CT + 0] = (uint8_t)max(0, min(255, r));
                dst[w * CT + 1] = (uint8_t)max(0, min(255, gg));
                dst[w * CT + 2] = (uint8_t)max(0, min(255, b));
                if (CT == 4) dst[w * CT + 3] = (uint8_t)p[3];
            } else {
#pragma unroll
                for (int i = 0; i < CT; ++i) dst[w * CT + i] = (uint8_t)p[i];
            }
        }
        __syncwarp();
        int16_t* tmp = bufA; bufA = bufB; bufB = tmp;                     // row y becomes row y-1, row y-1 becomes y-2
    }
}

static int fast_line_bytes(const Geom& g) { return 2 * (((min(g.tw, g.W) * g.C + 7) & ~7) * 2); }
// shared memory of one slice with the state rows in it: the chain (default) or the round-1 fast kernel
static int fast_smem_with_state(const Geom& g) {
    return switches().decoder_v1 ? kFastBase + fast_line_bytes(g) : chain_decoder_smem_bytes(g, false);
}
constexpr int kChainMaxSmem = 226 * 1024;                   // a CTA can have 227 KB on sm_100, static shared memory included
static bool fast_decoder_fits(const Geom& g) {
    if (g.C < 1 || g.C > 4) return false;
    if (switches().decoder_v1) return fast_smem_with_state(g) <= 200 * 1024;
    return chain_decoder_smem_bytes(g, true) <= kChainMaxSmem;          // the rows can always go behind L1
}
// State in shared memory while every slice of the call finds a shared-memory slot at once (3 per SM for
// <= 1024-wide RGB tiles), else state in global memory behind L1 (7+ slices per SM, one wave).
static bool decoder_wants_global_state(const Geom& g, bool shared_launch) {
    const int per_sm = (228 * 1024) / (fast_smem_with_state(g) + 1024);
    if (!fast_decoder_fits(g)) return false;
    if (!switches().decoder_v1 && fast_smem_with_state(g) > kChainMaxSmem) return true;   // wide tiles: no room for the rows
    if (switches().decoder_smem_state) return false;
    // a launch that runs beside other launches of the same call (pipelined host-buffer decode) must not take
    // shared-memory slots away from them
    return shared_launch || g.n_slices() > (uint64_t)per_sm * 148;
}
uint64_t decoder_global_state_bytes(const Geom& g, bool shared_launch) {
    return decoder_wants_global_state(g, shared_launch) ? g.n_slices() * (uint64_t)kStateBytes : 0;
}

static bool lines_fit_smem(const Geom& g) {
    return (uint64_t)3 * min(g.tw, g.W) * g.C * 2 <= (uint64_t)kDecMaxLineSmem;
}

uint64_t decoder_line_scratch_bytes(const Geom& g) {
    if (fast_decoder_fits(g) || lines_fit_smem(g)) return 0;
    return g.n_slices() * 3ull * min(g.tw, g.W) * g.C * 2ull;
}

cudaError_t configure_slice_decoder() {
    cudaError_t e = cudaFuncSetAttribute(k_slice_decoder<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kDecBaseSmem + kDecMaxLineSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_slice_decoder<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecBaseSmem);
    if (e != cudaSuccess) return e;
    const int fast_max = 200 * 1024;
#define LLC_SET(CT, G) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_slice_decoder_fast<CT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_max)
    LLC_SET(1, false); LLC_SET(2, false); LLC_SET(3, false); LLC_SET(4, false);
    LLC_SET(1, true); LLC_SET(2, true); LLC_SET(3, true); LLC_SET(4, true);
#undef LLC_SET
    return e;
}

cudaError_t launch_slice_decoder(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                 uint8_t* d_pixels, int16_t* d_line_scratch, uint8_t* d_gstate, int* d_status,
                                 cudaStream_t st, bool shared_launch) {
    const uint64_t ns = g.n_slices();
    if (ns == 0 || ns > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    if (fast_decoder_fits(g) && !switches().decoder_simple && !switches().decoder_v1) {
        uint8_t* gs = nullptr;
        if (decoder_wants_global_state(g, shared_launch)) {
            gs = d_gstate;
            const cudaError_t e = cudaMemsetAsync(d_gstate, 0, ns * (uint64_t)kStateBytes, st);   // all states start at 0
            if (e != cudaSuccess) return e;
        }
        return launch_slice_decoder_chain(d_payload, d_offsets, g, d_pixels, gs, d_status, st, shared_launch);
    }
    if (fast_decoder_fits(g) && !switches().decoder_simple) {
        const unsigned n = (unsigned)ns;
        if (decoder_wants_global_state(g, shared_launch)) {
            const int smem = kFastBase - kStateBytes + fast_line_bytes(g);
            uint2* gs = reinterpret_cast<uint2*>(d_gstate);
            cudaError_t e = cudaMemsetAsync(d_gstate, 0, ns * (uint64_t)kStateBytes, st);   // all states start at 0
            if (e != cudaSuccess) return e;
            switch (g.C) {
                case 1: k_slice_decoder_fast<1, true><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, gs); break;
                case 2: k_slice_decoder_fast<2, true><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, gs); break;
                case 3: k_slice_decoder_fast<3, true><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, gs); break;
                default: k_slice_decoder_fast<4, true><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, gs); break;
            }
            return cudaGetLastError();
        }
        const int smem = kFastBase + fast_line_bytes(g);
        switch (g.C) {
            case 1: k_slice_decoder_fast<1, false><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, nullptr); break;
            case 2: k_slice_decoder_fast<2, false><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, nullptr); break;
            case 3: k_slice_decoder_fast<3, false><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, nullptr); break;
            default: k_slice_decoder_fast<4, false><<<n, 32, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, nullptr); break;
        }
        return cudaGetLastError();
    }
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
