// sample.cuh -- the per-sample stage of llcomp::compressImage as device functions shared by the front-end kernels and by
// the fused coder (which computes its records from the pixels itself): colour transform (/root/reference/llcomp.hpp:396-414),
// neighbours with the border rules (:417-422), context hash (:424-429, quantisers :297-341), median predictor (:430),
// residual and sign fold (:431-436).
#pragma once
#include "common.cuh"

namespace llc {

constexpr int kQB = 640;                                    // plane differences lie in [-637, 637]

// quant11_table / quant5_table (llcomp.hpp:297-333) as closed forms, tabulated over the whole range a difference of
// two plane values can take: no clamp, the bias sits in the load's immediate offset.
struct QuantBytes {
    int8_t q11[2 * kQB];
    int8_t q5[2 * kQB];
};
constexpr QuantBytes make_quant_bytes() {
    QuantBytes t{};
    for (int i = 0; i < 2 * kQB; ++i) {
        const int x = i - kQB, a = x < 0 ? -x : x;
        const int m11 = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35), m5 = (a >= 1) + (a >= 4);
        t.q11[i] = (int8_t)(x < 0 ? -m11 : m11);
        t.q5[i] = (int8_t)(x < 0 ? -m5 : m5);
    }
    return t;
}
static_assert(sizeof(QuantBytes) % 16 == 0, "copied with 16-byte words");

// One sample from its plane values (all neighbours already substituted): llcomp.hpp:424-436.  n5 = 3025 q5(T - t).
// d2 is taken as t - tl, so that the predictor's l + t - tl is l + d2; q11 is odd, hence the -11.
__device__ __forceinline__ uint32_t code_sample(int cur, int l, int L, int tl, int t, int tr, int n5,
                                                const int8_t* __restrict__ q11, const int8_t* __restrict__ q5) {
    const int d1 = l - tl, d2 = t - tl, d3 = t - tr, d4 = L - l;
    int hash = (q11[d3] * 11 - q11[d2]) * 11 + q11[d1] + 605 * q5[d4] + n5;    // :424-429
    const int hi = max(l, t), lo = min(l, t);
    const int pred = max(min(l + d2, hi), lo);                                    // median(l, l+t-tl, t), :430
    int diff = cur - pred;                                                        // :431
    const int s = hash >> 31;                                                     // :433-436
    hash = abs(hash);
    diff = (diff ^ s) - s;
    return ((uint32_t)hash << 11) | ((uint32_t)diff & 0x7FFu);
}

// The same record from a signed hash and the unfolded residual, arranged for the pipes: with sigma = +-1 the sign of
// the hash, w = (hash 2048 + diff) sigma = |hash| 2048 + folded diff, and the record is w with its hash field put right
// when the folded residual is negative (bit 10 of w set): w + 2 (w & 0x400).  Three multiply-adds and three logic /
// shift operations instead of one and five: the front end is bound by the ALU pipe.
__device__ __forceinline__ uint32_t record_of(int hash, int diff) {
    const int sigma = (hash >> 31) | 1;
    const int w = (hash * 2048 + diff) * sigma;
    return (uint32_t)(w + 2 * (w & 0x400));
}

template <int CT>
struct Px {
    int v[CT];
};

// Planes of the pixel at p (llcomp.hpp:396-409): (R-G, G + trunc((B-G + R-G)/4), B-G, extra...) for three or more
// channels, the raw bytes otherwise.
template <int CT>
__device__ __forceinline__ Px<CT> planes_of(const uint8_t* __restrict__ p) {
    Px<CT> o;
    if (CT >= 3) {
        const int g = p[1];
        const int r = (int)p[0] - g, b = (int)p[2] - g;
        const int s2 = b + r;                                  // (b + r) / 4, truncating toward zero (:402)
        o.v[0] = r;
        o.v[1] = g + ((s2 + ((s2 >> 31) & 3)) >> 2);
        o.v[2] = b;
#pragma unroll
        for (int c = 3; c < CT; ++c) o.v[c] = p[c];
    } else {
#pragma unroll
        for (int c = 0; c < CT; ++c) o.v[c] = p[c];
    }
    return o;
}

// The CT records of the pixel at (w, h) of a slice sw wide whose first byte is p (row pitch in bytes): any position,
// neighbours that do not exist are neither read nor used (llcomp.hpp:417-422).
template <int CT>
__device__ __forceinline__ void records_of_pixel(const uint8_t* __restrict__ p, size_t pitch, int w, int h, int sw,
                                                 const int8_t* __restrict__ q11, const int8_t* __restrict__ q5,
                                                 uint32_t (&rec)[CT]) {
    const Px<CT> cur = planes_of<CT>(p);
    if (w > 1 && h > 1 && w < sw - 1) {                      // inside the slice: every neighbour exists
        const Px<CT> pl = planes_of<CT>(p - CT), pL = planes_of<CT>(p - 2 * CT);
        const Px<CT> pt = planes_of<CT>(p - pitch), ptl = planes_of<CT>(p - pitch - CT), ptr = planes_of<CT>(p - pitch + CT);
        const Px<CT> pT = planes_of<CT>(p - 2 * pitch);
#pragma unroll
        for (int c = 0; c < CT; ++c)
            rec[c] = code_sample(cur.v[c], pl.v[c], pL.v[c], ptl.v[c], pt.v[c], ptr.v[c], 3025 * q5[pT.v[c] - pt.v[c]], q11, q5);
        return;
    }
    const bool hl = w > 0, hL = w > 1, ht = h > 0, hT = h > 1, hr = w < sw - 1;
    const Px<CT> pl = planes_of<CT>(hl ? p - CT : p), pL = planes_of<CT>(hL ? p - 2 * CT : p);
    const uint8_t* up = ht ? p - pitch : p;
    const Px<CT> pt = planes_of<CT>(up), ptl = planes_of<CT>(ht && hl ? up - CT : up), ptr = planes_of<CT>(ht && hr ? up + CT : up);
    const Px<CT> pT = planes_of<CT>(hT ? p - 2 * pitch : p);
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int l = hl ? pl.v[c] : (ht ? pt.v[c] : 128);
        const int t = ht ? pt.v[c] : l;
        const int L = hL ? pL.v[c] : l;
        const int tl = (ht && hl) ? ptl.v[c] : t;
        const int tr = (ht && hr) ? ptr.v[c] : t;
        const int T = hT ? pT.v[c] : t;
        rec[c] = code_sample(cur.v[c], l, L, tl, t, tr, 3025 * q5[T - t], q11, q5);
    }
}

}  // namespace llc
