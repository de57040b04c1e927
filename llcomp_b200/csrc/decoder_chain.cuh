// decoder_chain.cuh -- the serial chain of one slice of the decoder (K5, second form), written so that the same source
// compiles for the device (k_slice_decoder_chain, decoder_chain.cu) and for the host (tests/host/chain_host.cpp, g++): the logic
// is checked against the oracle on the CPU box, the GPU runs confirm it and time it.
//
// Re-creates llcomp::decompressImage after its header parse (/root/reference/llcomp.hpp:475-545): RangeDecoder
// (:91-127), getSymbol (:219-247), the adaptive bit model (:283-293), neighbours / hash / predictor (:494-509), sign
// unfold (:511-515, :526-528), reconstruction (:529), inverse colour transform (:532-543).
//
// One thread decodes a slice; what it executes per binary decision and per sample is what bounds the decoder, so the
// chain is built around its dependency graph rather than around the reference's loop:
//   * range/low.  With Q = 256 - P(1), t = range Q + 255:  r0 = range - (range P >> 8) = t >> 8 (llcomp.hpp:107-108),
//     and with the code value carried as cS = 256 low + 255 the decision `low >= r0` is `cS >= t`: one multiply-add
//     feeds the compare, the shift is off the path.
//   * zero flag and exponent decisions steer control flow anyway (getSymbol's tree), so each is a real branch with the
//     update of range / low / sub-state written out in both arms (constant byte-permute selectors, no select
//     instructions); mantissa and sign decisions are branch-free.
//   * the table entries {Q, next states} of the sub-states a residual can touch are requested at the start of the
//     sample (volatile loads: they stay where they are written), never inside the decision chain.
//   * the context of the plane's next sample is prepared in two parts placed where their inputs have had time to arrive:
//     part 1 right after the sample is reconstructed (differences, the two quantiser look-ups, median), part 2 (hash,
//     request of the state row) after the next sample's table requests are on their way; the row then has CT-1 samples
//     to arrive and is patched from registers if one of them rewrote it.
//   * renormalisation (one decision in ten) is a branch; the next payload byte is already in a register.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LLC_HD __device__ __forceinline__          // under nvcc the chain is device code only
#else
#define LLC_HD inline
#endif

namespace llc {
namespace dchain {

constexpr int kQC = 64;                          // quantiser tables cover [-64, 63]; q11 saturates at 35, q5 at 4
constexpr int kRingBytes = 1024;                 // payload bytes staged ahead of the chain
constexpr int kMaxBytesPerSample = 66;           // a decision consumes at most one byte; <= 2*31+3 decisions + slack

struct alignas(8) Ent { uint32_t q, nx; };       // q = 256 - P(bit = 1); nx = next state if 0 | next state if 1 << 8
struct alignas(8) Row { uint32_t x, y; };        // the 8 sub-states of one context, one byte each

// Layout of the chain's shared memory: 32-bit addresses (device: shared-window addresses; host: offsets into an arena).
// The tables are shared by the warps of a CTA (one slice per warp); every warp has its own payload ring and row buffers.
// The entry table and the rings start on multiples of 1024, so that `index | table` is the address of an element (one
// three-input logic instruction instead of a mask and an add).  Seven slices of 1024 RGB pixels take 94 KB: an SM that
// decodes them as one CTA keeps 128 KB of L1 for their state rows (as seven CTAs with their own tables: 96 KB; the
// chain's speed follows the rows' L1 hit rate).
struct Layout {
    uint32_t ent, q11, q5, ring, bufA, bufB, end;            // ring, bufA, bufB: this warp's; end: of the whole CTA
};
constexpr uint32_t kLayoutAlign = 1024;
#if defined(__CUDACC__)
__host__
#endif
LLC_HD Layout make_layout(int row_elems, uint32_t base, int warp = 0, int n_warps = 1) {   // row_elems = tile width * channels
    Layout L;
    L.ent = (base + kLayoutAlign - 1) & ~(kLayoutAlign - 1);
    L.q11 = L.ent + 128 * 8;                     // 128-byte tables on multiples of 128
    L.q5 = L.q11 + 2 * kQC;
    const uint32_t rings = L.ent + 2 * kLayoutAlign;
    L.ring = rings + (uint32_t)warp * kRingBytes;
    const uint32_t row_bytes = (uint32_t)((row_elems + 4 + 7) & ~7) * 2;   // + one pixel of padding (read ahead)
    const uint32_t rows = rings + (uint32_t)n_warps * kRingBytes;
    L.bufA = rows + (uint32_t)warp * 2 * row_bytes;
    L.bufB = L.bufA + row_bytes;
    L.end = rows + (uint32_t)n_warps * 2 * row_bytes;
    return L;
}
// bytes to reserve when the base address is only known to be 16-byte aligned
#if defined(__CUDACC__)
__host__
#endif
LLC_HD uint32_t layout_bytes(int row_elems, int n_warps = 1) { return make_layout(row_elems, 0, 0, n_warps).end + kLayoutAlign; }

// ---- memory access -------------------------------------------------------------------------------------------
// Device: volatile PTX on 32-bit shared-window addresses.  Volatile keeps a load where it is written (a table request at
// the start of a sample must not sink into the decision that consumes it) and no 64-bit generic address arithmetic is
// carried through the chain.  Host: the same addresses index an arena.
#if defined(__CUDACC__)
struct Smem {
    __device__ __forceinline__ uint32_t u8(uint32_t a) const {
        uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
    }
    __device__ __forceinline__ int s8(uint32_t a) const {
        int v; asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
    }
    __device__ __forceinline__ int s16(uint32_t a) const {
        int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
    }
    __device__ __forceinline__ void st16(uint32_t a, int v) const {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
    }
    __device__ __forceinline__ void st8(uint32_t a, uint32_t v) const {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
    }
    __device__ __forceinline__ Ent ent(uint32_t a) const {
        Ent e; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.q), "=r"(e.nx) : "r"(a) : "memory"); return e;
    }
    __device__ __forceinline__ void st_ent(uint32_t a, Ent e) const {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(e.q), "r"(e.nx) : "memory");
    }
};
// state rows: global memory behind L1 (kGlobal; plain cached accesses) or shared memory
template <bool kGlobal>
struct StateMem {
    uint64_t g;                                  // global address of the slice's rows
    uint32_t s;                                  // shared-window address of the slice's rows
    __device__ __forceinline__ Row load(uint32_t ctx) const {
        Row r;
        if (kGlobal) asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(g + 8ull * ctx) : "memory");
        else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(s + 8u * ctx) : "memory");
        return r;
    }
    __device__ __forceinline__ void store(uint32_t ctx, Row r) const {
        if (kGlobal) asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(g + 8ull * ctx), "r"(r.x), "r"(r.y) : "memory");
        else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s + 8u * ctx), "r"(r.x), "r"(r.y) : "memory");
    }
};
__device__ __forceinline__ uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
struct Smem {
    uint8_t* base;
    uint32_t u8(uint32_t a) const { return base[a]; }
    int s8(uint32_t a) const { return (int8_t)base[a]; }
    int s16(uint32_t a) const { return *reinterpret_cast<const int16_t*>(base + a); }
    void st16(uint32_t a, int v) const { *reinterpret_cast<int16_t*>(base + a) = (int16_t)v; }
    void st8(uint32_t a, uint32_t v) const { base[a] = (uint8_t)v; }
    Ent ent(uint32_t a) const { return *reinterpret_cast<const Ent*>(base + a); }
    void st_ent(uint32_t a, Ent e) const { *reinterpret_cast<Ent*>(base + a) = e; }
};
template <bool kGlobal>
struct StateMem {
    Row* rows;
    Row load(uint32_t ctx) const { return rows[ctx]; }
    void store(uint32_t ctx, Row r) const { rows[ctx] = r; }
};
inline uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) {          // __byte_perm, default mode
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int k = 0; k < 4; ++k) d |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return d;
}
#endif

#if defined(__CUDACC__)
#define LLC_UNLIKELY(x) __builtin_expect(!!(x), 0)
// range * q + 255, pinned where it is written (the compiler would otherwise sink it below the renormalisation test)
__device__ __forceinline__ uint32_t mad255(uint32_t r, uint32_t q) {
    uint32_t t; asm volatile("mad.lo.u32 %0, %1, %2, 255;" : "=r"(t) : "r"(r), "r"(q)); return t;
}
#else
#define LLC_UNLIKELY(x) (x)
inline uint32_t mad255(uint32_t r, uint32_t q) { return r * q + 255u; }
#endif

// Payload bytes and pixels pass once: streaming accesses (evict-first), so that they do not push the slices' state rows
// -- 65 MB for 1024 slices, re-read all the time -- out of L2.
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t load_stream_u8(const uint8_t* p) { return __ldcs(p); }
__device__ __forceinline__ void store_stream_u8(uint8_t* p, uint32_t v) { __stcs(p, (uint8_t)v); }
#else
inline uint32_t load_stream_u8(const uint8_t* p) { return *p; }
inline void store_stream_u8(uint8_t* p, uint32_t v) { *p = (uint8_t)v; }
#endif

LLC_HD int iabs(int v) { return v < 0 ? -v : v; }
LLC_HD int imin(int a, int b) { return a < b ? a : b; }
LLC_HD int imax(int a, int b) { return a > b ? a : b; }
LLC_HD int q11_of(int x) {                                              // quant11_table, llcomp.hpp:297-320
    const int a = iabs(x);
    const int q = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35);
    return x < 0 ? -q : q;
}
LLC_HD int q5_of(int x) {                                               // quant5_table, llcomp.hpp:322-333
    const int a = iabs(x);
    const int q = (a >= 1) + (a >= 4);
    return x < 0 ? -q : q;
}

// ---- phases of a slice that every lane takes part in (lane, nl = number of lanes; host: 0, 1) ----------------------
// entry[s] of the model tables: P | next-if-MPS << 8 | next-if-LPS << 16 (common.cuh); MPS of state s is s & 1.
LLC_HD void fill_tables(const Smem& m, const Layout& L, const uint32_t* entry, int lane, int nl) {
    for (int i = lane; i < 128; i += nl) {
        const uint32_t e = entry[i], p = e & 0xFFu, nm = (e >> 8) & 0xFFu, nlps = (e >> 16) & 0xFFu;
        const uint32_t mps = (uint32_t)i & 1u;                           // llcomp.hpp:285, :290-292
        Ent t;
        t.q = 256u - p;
        t.nx = (mps == 0 ? nm : nlps) | ((mps == 1 ? nm : nlps) << 8);
        m.st_ent(L.ent + 8u * i, t);
    }
    for (int i = lane; i < 2 * kQC; i += nl) {                           // element d of a table sits at d & 127
        const int d = i < kQC ? i : i - 2 * kQC;
        m.st8(L.q11 + i, (uint32_t)q11_of(d) & 0xFFu);
        m.st8(L.q5 + i, (uint32_t)q5_of(d) & 0xFFu);
    }
}

// Payload bytes [filled, pos + kRingBytes) -> ring, zero fill past the end of the slice (llcomp.hpp:475-479).
LLC_HD uint32_t ring_refill(const Smem& m, const Layout& L, const uint8_t* src, uint32_t len, uint32_t filled,
                            uint32_t pos, int lane, int nl) {
    const uint32_t want = pos + kRingBytes;
    for (uint32_t k = filled + lane; k < want; k += nl) m.st8(L.ring | (k & (kRingBytes - 1)), k < len ? load_stream_u8(src + k) : 0u);
    return want;
}

// Before row h is decoded: for every sample the part of the hash that only involves the rows above,
// 11 q11(tl - t) + 121 q11(t - tr) + 3025 q5(T - t) with the border rules of llcomp.hpp:495-499, IN PLACE of the row
// h-2 value it consumed (bufA = row h-1, bufB = row h-2 -> hash part -> row h).  The pixel after the row reads as zero.
template <int CT>
LLC_HD void row_prehash(const Smem& m, uint32_t bufA, uint32_t bufB, int w, int h, int lane, int nl) {
    const int stride = w * CT;
    for (int j = lane; j < stride + CT; j += nl) {
        int pre = 0;
        if (h > 0 && j < stride) {
            const int x = j / CT;
            const int t = m.s16(bufA + 2 * j);
            const int tl = x > 0 ? m.s16(bufA + 2 * (j - CT)) : t;
            const int tr = x < w - 1 ? m.s16(bufA + 2 * (j + CT)) : t;
            const int T = h > 1 ? m.s16(bufB + 2 * j) : t;
            pre = 11 * q11_of(tl - t) + 121 * q11_of(t - tr) + 3025 * q5_of(T - t);
        }
        m.st16(bufB + 2 * j, pre);
    }
}

// After row h is decoded: inverse colour transform, clamp, store (llcomp.hpp:532-543).
template <int CT>
LLC_HD void row_output(const Smem& m, uint32_t bufB, int w, uint8_t* dst, int lane, int nl) {
    for (int x = lane; x < w; x += nl) {
        const uint32_t a = bufB + 2 * (x * CT);
        if (CT >= 3) {
            int r = m.s16(a), g = m.s16(a + 2), b = m.s16(a + 4);
            g -= (r + b) / 4;
            r += g;
            b += g;
            store_stream_u8(dst + x * CT + 0, (uint32_t)imax(0, imin(255, r)));
            store_stream_u8(dst + x * CT + 1, (uint32_t)imax(0, imin(255, g)));
            store_stream_u8(dst + x * CT + 2, (uint32_t)imax(0, imin(255, b)));
            if (CT == 4) store_stream_u8(dst + x * CT + 3, (uint32_t)m.s16(a + 6));
        } else {
            for (int i = 0; i < CT; ++i) store_stream_u8(dst + x * CT + i, (uint32_t)m.s16(a + 2 * i));
        }
    }
}

// ---- the chain ----------------------------------------------------------------------------------------------------
// Renormalisation out of line (kV & 1): written in line, the assembler predicates its six instructions into every
// decision although one decision in ten executes them.
struct Renormed { uint32_t R, cS, pos, nextb; };
#if defined(__CUDACC__)
__device__ __noinline__
#else
inline
#endif
Renormed renorm_out(uint32_t R, uint32_t cS, uint32_t pos, uint32_t nextb, uint32_t ring) {
    Smem m{};
    Renormed r;
    r.R = R << 8;
    r.cS = ((cS - 255u + nextb) << 8) | 255u;
    r.pos = pos + 1;
    r.nextb = m.u8(ring | (r.pos & (kRingBytes - 1)));
    return r;
}

// kV: variant bits for measurements.  1: renormalisation out of line; 2: part 2 of the context left to the assembler's
// own placement; 4: products taken ahead of the renormalisation test (residual_spec).
template <int CT, bool kGlobal, int kV = 0>
struct Chain {
    Smem m;
    Layout L;
    StateMem<kGlobal> st;
    uint32_t bufA, bufB;             // row h-1; row h-2 -> hash part -> row h
    // range decoder (llcomp.hpp:91-127): R = range, cS = 256 low + 255; ring[pos] is the next unread byte = nextb
    uint32_t R, cS, pos, nextb;
    bool bad;
    // per plane: the sample the plane decodes next
    int l[CT], t[CT];                // left neighbour; top neighbour (becomes the next sample's tl)
    int pred[CT], pre[CT];           // median prediction; hash part from the rows above
    int q11v[CT], q5v[CT];           // q11(l - tl), q5(L - l), requested in part 1
    int nhash[CT], nah[CT];          // signed context hash and its magnitude (part 2)
    Row nrow[CT];                    // its state row as requested in part 2
    int whash[CT];                   // |hash| and row the plane wrote back last (forwarding)
    Row wrow[CT];
    int tn[CT], pn[CT];              // top neighbour and hash part of the sample after that, requested ahead

    LLC_HD void init(const Smem& mem, const Layout& lay, const StateMem<kGlobal>& state) {
        m = mem; L = lay; st = state;
        bufA = L.bufA; bufB = L.bufB;
        R = 0xFF00u;                                                    // llcomp.hpp:93-96
        const uint32_t low = (m.u8(L.ring + 0) << 8) | m.u8(L.ring + 1);
        cS = (low << 8) | 255u;
        pos = 2;
        nextb = m.u8(L.ring + 2);
        bad = false;
        for (int i = 0; i < CT; ++i) { whash[i] = -1; wrow[i].x = 0; wrow[i].y = 0; }
    }

    LLC_HD void renorm() {                                              // llcomp.hpp:98-104
#if defined(__CUDACC__)
        if (kV & 1) {
            const Renormed r = renorm_out(R, cS, pos, nextb, L.ring);
            R = r.R; cS = r.cS; pos = r.pos; nextb = r.nextb;
            return;
        }
#endif
        R <<= 8;
        cS = ((cS - 255u + nextb) << 8) | 255u;
        ++pos;
        nextb = m.u8(L.ring | (pos & (kRingBytes - 1)));
    }

    // part 1 of the context of plane i's next sample: differences, quantiser look-ups, median.  l[i], t[i] (the top
    // neighbour of the sample just decoded, i.e. the new top-left) and `left2` (the new left-left) are in registers.
    template <bool kFirstRow>
    LLC_HD void part1(int i, int cur, int left2, int t_new, int pre_new) {
        const int tl = kFirstRow ? cur : t[i];                          // first row: t = tl = l (llcomp.hpp:495-499)
        const int tt = kFirstRow ? cur : t_new;
        const int d1 = imax(-kQC, imin(kQC - 1, cur - tl));
        const int d4 = imax(-kQC, imin(kQC - 1, left2 - cur));
        q11v[i] = m.s8(L.q11 | ((uint32_t)d1 & (2 * kQC - 1)));      // tables in wrapped order: index d & 127
        q5v[i] = m.s8(L.q5 | ((uint32_t)d4 & (2 * kQC - 1)));
        const int lt = cur + tt - tl;
        pred[i] = imax(imin(cur, lt), imin(imax(cur, lt), tt));         // median, llcomp.hpp:509
        l[i] = cur;
        t[i] = tt;
        pre[i] = pre_new;
    }
    // part 2: hash (llcomp.hpp:501-507) and the request of its state row
    // `after`: a value that is zero but that the assembler cannot know to be (a table entry's q >> 9): it ties the
    // position of part 2 to the arrival of that entry, i.e. behind the table requests of the sample in hand.  Left to
    // itself the assembler hoists the row request to where part 1 ends and the thread then waits there for the two
    // quantiser look-ups it has just issued.
    LLC_HD void part2(int i, uint32_t after = 0) {
        const int hsh = pre[i] + q11v[i] + 605 * q5v[i] + (int)after;
        nhash[i] = hsh;
        nah[i] = iabs(hsh);
        nrow[i] = st.load((uint32_t)nah[i]);
    }

    // Start of row h: every plane's first sample.  First column: l = L = tl = t (h > 0) or 128 (h == 0).
    template <bool kFirstRow>
    LLC_HD void row_begin() {
        for (int i = 0; i < CT; ++i) {
            const int t0 = kFirstRow ? 128 : m.s16(bufA + 2 * i);
            const int p0 = m.s16(bufB + 2 * i);
            // l = L = tl = t0: d1 = d4 = 0, median = t0
            q11v[i] = 0; q5v[i] = 0;
            pred[i] = t0; l[i] = t0; t[i] = t0; pre[i] = p0;
        }
        for (int i = 0; i + 1 < CT; ++i) part2(i);
    }

    // One decision with a branch-free update (mantissa, sign): returns the bit, writes the next state into byte kB of w.
    template <int kB>
    LLC_HD uint32_t bin_flat(const Ent e, uint32_t& w) {
        const uint32_t tq = R * e.q + 255u;
        const uint32_t r0 = tq >> 8, t0 = tq & ~0xFFu;
        const bool bit = cS >= tq;                                      // low >= range - r1, llcomp.hpp:110
        R = bit ? R - r0 : r0;
        const uint32_t cd = cS - t0;                                    // wraps above cS exactly when the bit is 0
        cS = cd < cS ? cd : cS;                                         // (cS >= tq <=> cS >= t0: the low byte of cS is 0xFF)
        constexpr uint32_t s0 = kB == 0 ? 0x3214u : kB == 1 ? 0x3240u : kB == 2 ? 0x3410u : 0x4210u;
        constexpr uint32_t s1 = kB == 0 ? 0x3215u : kB == 1 ? 0x3250u : kB == 2 ? 0x3510u : 0x5210u;
        w = bperm(w, e.nx, bit ? s1 : s0);
        if (R < 0x100u) renorm();
        return bit ? 1u : 0u;
    }

// A decision that steers control flow: both arms written out.  ARM1 / ARM0 are statements.
#define LLC_DCHAIN_BIN(E, W, S0, S1, ARM1, ARM0)                   \
    {                                                              \
        const uint32_t tq_ = R * (E).q + 255u;                     \
        const uint32_t r0_ = tq_ >> 8;                             \
        if (cS >= tq_) {                                           \
            cS -= r0_ << 8;                                        \
            R -= r0_;                                              \
            W = bperm(W, (E).nx, S1);                              \
            if (R < 0x100u) renorm();                              \
            ARM1                                                   \
        } else {                                                   \
            R = r0_;                                               \
            W = bperm(W, (E).nx, S0);                              \
            if (R < 0x100u) renorm();                              \
            ARM0                                                   \
        }                                                          \
    }

    LLC_HD Ent ent_of(uint32_t word, int k) const {                     // entry of the sub-state in byte k of word
        const uint32_t off = k == 0 ? (word << 3) & 0x7F8u : (word >> (8 * k - 3)) & 0x7F8u;
        return m.ent(L.ent | off);
    }

    // magnitude bits below the leading one, then the sign (llcomp.hpp:237-245): E = exponent (number of mantissa bits)
    template <int E>
    LLC_HD int mantissa_sign(Row& row, const Ent e5, Ent e6, const Ent e7) {
        uint32_t value = 1;
        if (E >= 1) value = 2u + bin_flat<1>(e5, row.y);
        if (E >= 2) value += value + bin_flat<2>(e6, row.y);
        if (E >= 3) {
            e6 = ent_of(row.y, 2);
            value += value + bin_flat<2>(e6, row.y);
        }
        const uint32_t sgn = bin_flat<3>(e7, row.y);
        return sgn ? -(int)value : (int)value;
    }

    // The residual of one sample (getSymbol, llcomp.hpp:219-247) with the sub-states of `row`.
    LLC_HD int residual(Row& row) {
        const Ent e0 = ent_of(row.x, 0), e1 = ent_of(row.x, 1), e2 = ent_of(row.x, 2), e3 = ent_of(row.x, 3);
        const Ent e5 = ent_of(row.y, 1), e6 = ent_of(row.y, 2), e7 = ent_of(row.y, 3);
        return residual_with(row, e0, e1, e2, e3, e5, e6, e7);
    }
    LLC_HD int residual_with(Row& row, const Ent e0, const Ent e1, const Ent e2, const Ent e3, const Ent e5, const Ent e6,
                             const Ent e7) {
        int diff = 0;
        LLC_DCHAIN_BIN(e0, row.x, 0x3214u, 0x3215u, { diff = 0; }, {
            LLC_DCHAIN_BIN(e1, row.x, 0x3240u, 0x3250u, {
                LLC_DCHAIN_BIN(e2, row.x, 0x3410u, 0x3510u, {
                    LLC_DCHAIN_BIN(e3, row.x, 0x4210u, 0x5210u, {
                        // exponent >= 3: context 4 repeats (llcomp.hpp:230-235), then the general mantissa loop
                        int e = 3;
                        for (;;) {
                            const Ent e4 = ent_of(row.y, 0);
                            if (!bin_flat<0>(e4, row.y)) break;
                            if (++e > 31) { bad = true; pos = 0x80000000u; break; }   // ends the pixel loop at its next check
                        }
                        uint32_t value = 2u + bin_flat<1>(e5, row.y);
                        for (int k = e - 2; k >= 0; --k) {
                            const Ent e6d = ent_of(row.y, 2);
                            value += value + bin_flat<2>(e6d, row.y);
                        }
                        const uint32_t sgn = bin_flat<3>(e7, row.y);
                        diff = sgn ? -(int)value : (int)value;
                    }, { diff = mantissa_sign<2>(row, e5, e6, e7); })
                }, { diff = mantissa_sign<1>(row, e5, e6, e7); })
            }, { diff = mantissa_sign<0>(row, e5, e6, e7); })
        })
        return diff;
    }

    // ---- the decisions, second form (kV & 4): the product of the decision that FOLLOWS is taken as soon as the range is
    // known, before it is known whether the range has to be renormalised first.  One decision in ten renormalises: a
    // cold loop does it and takes the product again.  Written in line and predicated by the assembler, the test, the
    // shift and the product sit between every two decisions with their full fixed latencies, taken or not.
#define LLC_PREP(TQ, Q)                                            \
    uint32_t TQ = mad255(R, (Q));                                  \
    while (LLC_UNLIKELY(R < 0x100u)) { renorm(); TQ = mad255(R, (Q)); }
#define LLC_SETTLE() while (LLC_UNLIKELY(R < 0x100u)) renorm();
#define LLC_SBIN(TQ, E, W, S0, S1, ARM1, ARM0)                     \
    {                                                              \
        const uint32_t r0_ = (TQ) >> 8;                            \
        if (cS >= (TQ)) {                                          \
            cS -= (TQ) & ~0xFFu;                                   \
            R -= r0_;                                              \
            W = bperm(W, (E).nx, S1);                              \
            ARM1                                                   \
        } else {                                                   \
            R = r0_;                                               \
            W = bperm(W, (E).nx, S0);                              \
            ARM0                                                   \
        }                                                          \
    }
    // branch-free decision whose product tq is already taken; leaves the range settled
    template <int kB>
    LLC_HD uint32_t flat_s(uint32_t tq, const Ent e, uint32_t& w) {
        const uint32_t r0 = tq >> 8, t0 = tq & ~0xFFu;
        const bool bit = cS >= tq;
        R = bit ? R - r0 : r0;
        cS = bit ? cS - t0 : cS;
        constexpr uint32_t s0 = kB == 0 ? 0x3214u : kB == 1 ? 0x3240u : kB == 2 ? 0x3410u : 0x4210u;
        constexpr uint32_t s1 = kB == 0 ? 0x3215u : kB == 1 ? 0x3250u : kB == 2 ? 0x3510u : 0x5210u;
        w = bperm(w, e.nx, bit ? s1 : s0);
        return bit ? 1u : 0u;
    }
    // mantissa and sign for an exponent E <= 2 reached with the range not yet settled
    template <int E>
    LLC_HD int mantissa_sign_s(Row& row, const Ent e5, const Ent e6, const Ent e7) {
        uint32_t value = 1;
        if (E >= 1) {
            LLC_PREP(t5, e5.q)
            value = 2u + flat_s<1>(t5, e5, row.y);
        }
        if (E >= 2) {
            LLC_PREP(t6, e6.q)
            value += value + flat_s<2>(t6, e6, row.y);
        }
        LLC_PREP(t7, e7.q)
        const uint32_t sgn = flat_s<3>(t7, e7, row.y);
        LLC_SETTLE()
        return sgn ? -(int)value : (int)value;
    }
    LLC_HD int residual_spec(Row& row, const Ent e0, const Ent e1, const Ent e2, const Ent e3, const Ent e5, const Ent e6,
                             const Ent e7) {
        int diff = 0;
        const uint32_t t0q = mad255(R, e0.q);                            // the range is settled between samples
        LLC_SBIN(t0q, e0, row.x, 0x3214u, 0x3215u, { LLC_SETTLE() diff = 0; }, {
            LLC_PREP(t1q, e1.q)
            LLC_SBIN(t1q, e1, row.x, 0x3240u, 0x3250u, {
                LLC_PREP(t2q, e2.q)
                LLC_SBIN(t2q, e2, row.x, 0x3410u, 0x3510u, {
                    LLC_PREP(t3q, e3.q)
                    LLC_SBIN(t3q, e3, row.x, 0x4210u, 0x5210u, {
                        // exponent >= 3: context 4 repeats (llcomp.hpp:230-235), then the general mantissa loop
                        LLC_SETTLE()
                        int e = 3;
                        for (;;) {
                            const Ent e4 = ent_of(row.y, 0);
                            if (!bin_flat<0>(e4, row.y)) break;
                            if (++e > 31) { bad = true; pos = 0x80000000u; break; }   // ends the pixel loop at its next check
                        }
                        uint32_t value = 2u + bin_flat<1>(e5, row.y);
                        for (int k = e - 2; k >= 0; --k) {
                            const Ent e6d = ent_of(row.y, 2);
                            value += value + bin_flat<2>(e6d, row.y);
                        }
                        const uint32_t sgn = bin_flat<3>(e7, row.y);
                        diff = sgn ? -(int)value : (int)value;
                    }, { diff = mantissa_sign_s<2>(row, e5, e6, e7); })
                }, { diff = mantissa_sign_s<1>(row, e5, e6, e7); })
            }, { diff = mantissa_sign_s<0>(row, e5, e6, e7); })
        })
        return diff;
    }
#undef LLC_PREP
#undef LLC_SETTLE
#undef LLC_SBIN

    // Pixels [w0, w_end) of the row in hand; stops early (returns the pixel reached) when the payload staged in the
    // ring could run out (pos > pos_limit).  A bad stream (exponent > 31) sets `bad` and a huge `pos`: the loop then
    // stops at the next pixel (what it decodes until then is discarded by the caller), so the sample loop does not
    // carry a test of the flag.
    template <bool kFirstRow>
    LLC_HD int run(int w0, int w_end, uint32_t pos_limit) {
        int w = w0;
        for (; w < w_end; ++w) {
            if (pos > pos_limit) break;
            const uint32_t aA = bufA + 2u * (uint32_t)(w * CT), aB = bufB + 2u * (uint32_t)(w * CT);   // pixel w in the row buffers
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int i = 0; i < CT; ++i) {
                if (CT == 1) part2(0);
                // ---- [A] the row as requested, or as rewritten since; table entries; operands of the sample after
                const int hsh = nhash[i];
                const int ah = nah[i];
                Row& row = wrow[i];                                      // the plane's "last written" row is the working copy
                row = nrow[i];
                for (int k = CT - 1; k >= 1; --k) {                      // k samples ago; newest last
                    const int p = (i + CT - k) % CT;
                    if (ah == whash[p]) row = wrow[p];
                }
                whash[i] = ah;                                           // (only the other planes compare against it)
                const Ent e0 = ent_of(row.x, 0), e1 = ent_of(row.x, 1), e2 = ent_of(row.x, 2), e3 = ent_of(row.x, 3);
                const Ent e5 = ent_of(row.y, 1), e6 = ent_of(row.y, 2), e7 = ent_of(row.y, 3);
                const int t_next = kFirstRow ? 0 : m.s16(aA + 2u * (CT + i));
                const int p_next = m.s16(aB + 2u * (CT + i));
                // ---- [B] part 2 of the plane before this one (its quantiser look-ups have arrived)
                if (CT > 1) part2((i + CT - 1) % CT, (kV & 2) ? 0u : e0.q >> 9);
                // ---- [C] the decisions
                const int pr = pred[i], left = l[i];
                int diff = (kV & 4) ? residual_spec(row, e0, e1, e2, e3, e5, e6, e7)
                                    : residual_with(row, e0, e1, e2, e3, e5, e6, e7);
                // ---- [D] write back, reconstruct, part 1 of the plane's next sample
                st.store((uint32_t)ah, row);
                const int cur = (int)(int16_t)(pr + (hsh < 0 ? -diff : diff));   // llcomp.hpp:526-529
                m.st16(aB + 2u * i, cur);
                part1<kFirstRow>(i, cur, w == 0 ? cur : left, t_next, p_next);   // w == 0: the next sample's L = l (:496)
            }
        }
        return w;
    }
#undef LLC_DCHAIN_BIN
};

// A whole slice with `nl` lanes (host: one).  Returns false on a bad stream (exponent > 31, llcomp.hpp:232).
// dst = first pixel of the slice, pitch in bytes.  sync() separates the phases (device: __syncwarp).
template <int CT, bool kGlobal, int kV = 0, class Sync>
LLC_HD bool decode_slice_rows(const Smem& m, const Layout& L, const StateMem<kGlobal>& state, const uint32_t* entry,
                              const uint8_t* src, uint32_t len, int w, int h, uint8_t* dst, size_t pitch, int lane, int nl,
                              Sync sync, bool tables_ready = false) {
    if (!tables_ready) fill_tables(m, L, entry, lane, nl);     // (a CTA of several warps fills them once, before)
    uint32_t filled = ring_refill(m, L, src, len, 0, 0, lane, nl);
    sync();
    Chain<CT, kGlobal, kV> c;
    c.init(m, L, state);
    bool ok = true;
    for (int y = 0; y < h; ++y) {
        row_prehash<CT>(m, c.bufA, c.bufB, w, y, lane, nl);
        sync();
        if (lane == 0) {
            if (y == 0) c.template row_begin<true>(); else c.template row_begin<false>();
        }
        int x = 0;
        while (x < w) {
            filled = ring_refill(m, L, src, len, filled, c.pos, lane, nl);
            sync();
            if (lane == 0) {
                const uint32_t limit = filled - CT * kMaxBytesPerSample - 1;   // ring[pos] itself stays staged
                x = y == 0 ? c.template run<true>(x, w, limit) : c.template run<false>(x, w, limit);
            }
#if defined(__CUDACC__)
            x = __shfl_sync(0xFFFFFFFFu, x, 0);
            c.pos = __shfl_sync(0xFFFFFFFFu, c.pos, 0);
            ok = !__shfl_sync(0xFFFFFFFFu, (int)c.bad, 0);
#else
            ok = !c.bad;
#endif
            if (!ok) return false;
        }
        sync();
        row_output<CT>(m, c.bufB, w, dst + (size_t)y * pitch, lane, nl);
        sync();
        const uint32_t tmp = c.bufA; c.bufA = c.bufB; c.bufB = tmp;     // row y becomes row y-1, row y-1 becomes y-2
    }
    return ok;
}

template <int CT, bool kGlobal, class Sync>
LLC_HD bool decode_slice(const Smem& m, const StateMem<kGlobal>& state, const uint32_t* entry, const uint8_t* src,
                         uint32_t len, int w, int h, uint8_t* dst, size_t pitch, int lane, int nl, Sync sync) {
    return decode_slice_rows<CT, kGlobal, 0>(m, make_layout(w * CT, 0), state, entry, src, len, w, h, dst, pitch, lane, nl, sync);
}

}  // namespace dchain
}  // namespace llc
