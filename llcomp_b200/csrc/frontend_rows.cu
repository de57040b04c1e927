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
#include "sample.cuh"

namespace llc {

namespace {

constexpr int kFrThreads = 128;
#ifndef LLC_FR_PX3
#define LLC_FR_PX3 8
#define LLC_FR_ROWS3 16
#define LLC_FR_MINB3 3
#endif
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

template <int CT, int PX, int ROWS>
struct FrShape {
    static constexpr int kRegionW = kFrThreads * PX;           // pixels of a region row
    static constexpr int kRowBytes = kRegionW * CT + 32;       // + one 16-byte chunk either side
    static constexpr int kSmem = (ROWS + 2) * kRowBytes;       // dynamic: the raw rows (tables and barrier are static)
    // bytes a thread reads per row: pixels x-2 .. x+PX, starting at the word that holds byte 16 - 2 CT + PX CT tid
    static constexpr int kShift = (16 - 2 * CT) & 3;
    static constexpr int kWords = (kShift + (PX + 3) * CT + 3) / 4;
    static_assert((PX * CT) % 4 == 0, "a thread's pixel group starts on a word");
    static_assert(ROWS % 2 == 0, "the row loop is unrolled by two");
};

// kV, variants kept for cross-checks and measurements (0 = default).
// bit 0 = the warp lays its row of records out in shared memory and the TMA unit stores it (UBLKCP.G.S), instead of
// every thread storing its own 96 contiguous bytes with 128-bit stores.  A streaming kernel with this 1-read-4-write mix
// and per-thread-contiguous stores tops out at 4.1 TB/s against 6.0 TB/s with whole-sector stores
// (profiles/microbench/write_mix.cu), which is where this kernel sits -- but it sits there for another reason: with
// TMA stores it takes 4.32 ms against 4.20, and with the staged row read back by the lanes for coalesced 128-bit stores
// 4.63: its issue slots and its ALU pipe are 77 % busy with 12 warps per SM, it is bound by instructions per sample.
// bit 2 = the round-2a sample code (five look-ups per sample, code_sample)
// instead of the shared look-ups below.
// (Measured and dropped: raw bytes fetched one by one instead of words + extraction, 4.49 against 4.17 ms; the colour
// transform's division through a byte table, 4.24 ms.)
template <int CT, int PX, int ROWS, int MINB, int kV>
__global__ void __launch_bounds__(kFrThreads, MINB) k_frontend_rows(const uint8_t* __restrict__ pixels, Geom g,
                                                                    uint32_t* __restrict__ sym) {
    using S = FrShape<CT, PX, ROWS>;
    extern __shared__ __align__(128) uint8_t raw[];
    // static, so that a look-up is one LDS whose table address is an immediate: the difference that indexes it then
    // comes straight from the multiply-add pipe instead of a three-input add on the (busier) ALU pipe
    __shared__ __align__(16) QuantBytes lut;
    constexpr int NV = PX * CT / 4;                                        // 16-byte words of records per thread and row
    __shared__ __align__(128) uint4 stage[(kV & 1) ? kFrThreads * NV : 1];   // per warp: its row of records as it lies in HBM
    __shared__ __align__(8) unsigned long long bar_word;
    const uint32_t bar = smem_u32(&bar_word);

    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int rx0 = blockIdx.x * S::kRegionW, ry0 = blockIdx.y * ROWS;
    const size_t pitch = (size_t)g.W * CT;
    const uint8_t* base = pixels + (size_t)img * g.H * pitch;

    // ---- rows ry0-2 .. ry0+ROWS-1 of the region -> shared memory, by the TMA unit
    const bool pre = rx0 > 0, post = rx0 + S::kRegionW < g.W;
    const uint32_t row_bytes = (uint32_t)(min(S::kRegionW, g.W - rx0) * CT) + (pre ? 16u : 0u) + (post ? 16u : 0u);
    const int r_first = max(0, 2 - ry0);                                   // rows above the image do not exist
    const int r_end = min(ROWS + 2, g.H - ry0 + 2);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, row_bytes * (uint32_t)(r_end - r_first));
    }
    for (int i = tid; i < (int)(sizeof(QuantBytes) / 16); i += kFrThreads)
        reinterpret_cast<uint4*>(&lut)[i] = reinterpret_cast<const uint4*>(&c_quant_bytes)[i];
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

    const int8_t* q11 = lut.q11 + kQB;
    const int8_t* q5 = lut.q5 + kQB;
    const uint32_t* my_row = reinterpret_cast<const uint32_t*>(raw + ((16 - 2 * CT) & ~3) + PX * CT * tid);

    mbar_wait(bar, 0);

    // Two plane arrays that swap roles every row (the loop is unrolled by two): pixels x-2 .. x+PX of the row in hand
    // (index j+2) and of the row above; n5 carries the row y-2 term of the hash for pixels x .. x+PX-1.
    Px<CT> pa[PX + 3], pb[PX + 3];
    int n5[PX][CT];
#pragma unroll
    for (int j = 0; j < PX + 3; ++j)
#pragma unroll
        for (int c = 0; c < CT; ++c) pb[j].v[c] = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int c = 0; c < CT; ++c) n5[j][c] = 0;

    auto do_row = [&](int r, Px<CT> (&cur)[PX + 3], const Px<CT> (&top)[PX + 3]) {
        uint32_t wv[S::kWords];
        const uint32_t* rp = my_row + r * (S::kRowBytes / 4);
#pragma unroll
        for (int k = 0; k < S::kWords; ++k) wv[k] = rp[k];
#pragma unroll
        for (int j = 0; j < PX + 3; ++j) cur[j] = planes_at<CT>(wv, S::kShift + j * CT);

        const int y = ry0 + r - 2;
        const int h_row = h;                                                // slice row of this image row (h moves on below)
        if (r >= 2 && y < g.H) {                                            // warp-uniform
            uint32_t rec[PX * CT];
            if (!active) {
#pragma unroll
                for (int k = 0; k < PX * CT; ++k) rec[k] = 0;
            } else if (h >= 2 && !(kV & 4)) {
                // Interior rows, shared look-ups.  With a3[j] = q11(t - tr) of sample j: q11(t - tl) of sample j is
                // -a3[j-1] (the same two values of the row above, swapped); with v[j] = q11(top - cur) at column j:
                // q11(l - tl) of sample j is -v[j-1], and 3025 q5 of the same difference is the next row's T - t term.
                // Four look-ups and four subtractions per sample instead of five and five.
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    int a3_prev = wl ? 0 : (int)q11[top[1].v[c] - top[2].v[c]];          // sample -1's t - tr; left edge: tl = t
                    int v_prev = wl ? 0 : (int)q11[top[1].v[c] - cur[1].v[c]];          // column -1;        left edge: l = tl = t
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        const int t = top[j + 2].v[c], x = cur[j + 2].v[c];
                        int l = cur[j + 1].v[c], L = cur[j].v[c], tl = top[j + 1].v[c];
                        if (j == 0) { l = wl ? t : l; tl = wl ? t : tl; }
                        if (j <= 1) L = wl ? l : L;
                        const int a3 = (j == PX - 1 && wr) ? 0 : (int)q11[t - top[j + 3].v[c]];   // right edge: tr = t
                        const int vd = t - x;
                        const int v = q11[vd];
                        const int hash = (a3 * 11 + a3_prev) * 11 - v_prev + 605 * q5[L - l] + n5[j][c];   // :424-429
                        n5[j][c] = 3025 * q5[vd];                                                          // next row's T - t
                        a3_prev = a3;
                        v_prev = v;
                        const int hi = max(l, t), lo = min(l, t);
                        const int pred = max(min(l + t - tl, hi), lo);                                    // median, :430
                        rec[j * CT + c] = record_of(hash, x - pred);                                       // :431-436
                    }
                }
            } else if (h >= 2) {
#pragma unroll
                for (int j = 0; j < PX; ++j)
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        // the slice's first two and last columns: llcomp.hpp:417-422 with h > 1
                        int l = cur[j + 1].v[c], L = cur[j].v[c];
                        const int t = top[j + 2].v[c];
                        int tl = top[j + 1].v[c], tr = top[j + 3].v[c];
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
                        const int top_c = top[j + 2].v[c];
                        const int l = wj > 0 ? cur[j + 1].v[c] : (h > 0 ? top_c : 128);
                        const int t = h > 0 ? top_c : l;
                        const int L = wj > 1 ? cur[j].v[c] : l;
                        const int tl = (h > 0 && wj > 0) ? top[j + 1].v[c] : t;
                        const int tr = (h > 0 && wj < sw - 1) ? top[j + 3].v[c] : t;
                        // T = t on the first two rows of a slice: q5(0) = 0
                        rec[j * CT + c] = code_sample(cur[j + 2].v[c], l, L, tl, t, tr, 0, q11, q5);
                    }
            }
            if (!(kV & 1)) {
                if (active) {
                    uint4* out = reinterpret_cast<uint4*>(img_out + out_idx);
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        out[k] = make_uint4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
                }
            } else {
                // (variant) The warp lays its row of records out in shared memory as it lies in HBM and the TMA unit
                // stores it: one bulk copy per run of lanes inside one slice (the whole warp, 3 KB, when the slice is at
                // least 256 pixels wide), issued by the run's first lane.  Warp-level synchronisation only; the previous
                // row's copy has long read its source when the wait is reached.
                const int lane = tid & 31;
                uint4* const st = stage + (tid >> 5) * 32 * NV + lane * NV;
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int k = 0; k < NV; ++k) st[k] = make_uint4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (active && (lane == 0 || wl)) {
                    const int run = min(32 - lane, (sw - w) / PX);            // lanes of this slice from here on
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(img_out + out_idx),
                                 "r"(smem_u32(st)), "r"((uint32_t)run * NV * 16u) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (active) {
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
        }
        // the difference of this row to the one above is the next row's T - t (the shared-look-up form has set it already)
        if (!(r >= 2 && y < g.H && active && h_row >= 2 && !(kV & 4))) {
#pragma unroll
            for (int j = 0; j < PX; ++j)
#pragma unroll
                for (int c = 0; c < CT; ++c) n5[j][c] = 3025 * q5[top[j + 2].v[c] - cur[j + 2].v[c]];
        }
    };
#pragma unroll 1
    for (int r = 0; r < ROWS + 2; r += 2) {
        do_row(r, pa, pb);
        do_row(r + 1, pb, pa);
    }
    if (kV & 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the CTA's shared memory outlives its bulk stores
}

// PX pixels per thread and ROWS rows per region: 3 channels -> 8 x 16 (a thread's row is 24 samples, 6 x 128-bit
// stores; the region row is 1024 pixels), 4 channels -> 4 x 32.
template <int CT> struct FrPick;
template <> struct FrPick<3> { static constexpr int PX = LLC_FR_PX3, ROWS = LLC_FR_ROWS3, MINB = LLC_FR_MINB3; };
template <> struct FrPick<4> { static constexpr int PX = 4, ROWS = 32, MINB = 3; };

}  // namespace

// true when the streaming kernel can take this geometry (the caller falls back to frontend.cu otherwise)
bool frontend_rows_applicable(const uint8_t* d_pixels, const Geom& g) {
    if (g.C != 3 && g.C != 4) return false;
    const int px = g.C == 3 ? FrPick<3>::PX : FrPick<4>::PX, rows = g.C == 3 ? FrPick<3>::ROWS : FrPick<4>::ROWS;
    if (((size_t)g.W * g.C) % 16 != 0 || g.tw % px != 0 || g.W % px != 0) return false;
    if ((reinterpret_cast<uintptr_t>(d_pixels) & 15) != 0) return false;
    if (g.image_samples() >= (1ull << 32)) return false;                   // 32-bit record index inside an image
    if ((g.H + rows - 1) / rows > 65535 || g.n_images > 65535) return false;
    return true;
}

template <int CT, int kV = 0>
static cudaError_t launch_rows(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, cudaStream_t st) {
    using P = FrPick<CT>;
    using S = FrShape<CT, P::PX, P::ROWS>;
    const cudaError_t configured = ensure_dynamic_smem<k_frontend_rows<CT, P::PX, P::ROWS, P::MINB, kV>>(S::kSmem);
    if (configured != cudaSuccess) return configured;
    dim3 grid((g.W + S::kRegionW - 1) / S::kRegionW, (g.H + P::ROWS - 1) / P::ROWS, g.n_images);
    k_frontend_rows<CT, P::PX, P::ROWS, P::MINB, kV><<<grid, kFrThreads, S::kSmem, st>>>(d_pixels, g, d_sym);
    return cudaGetLastError();
}

cudaError_t configure_frontend_rows() { return cudaSuccess; }   // the kernels configure themselves at first launch

cudaError_t launch_frontend_rows(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, cudaStream_t st) {
    if (g.C == 3) {
        switch (switches().frontend_variant) {                 // LLCOMP_FRONTEND_VARIANT: measurement variants (kV)
            case 1: return launch_rows<3, 1>(d_pixels, g, d_sym, st);
            case 4: return launch_rows<3, 4>(d_pixels, g, d_sym, st);
            case 5: return launch_rows<3, 5>(d_pixels, g, d_sym, st);
            default: return launch_rows<3, 0>(d_pixels, g, d_sym, st);
        }
    }
    switch (switches().frontend_variant) {
        case 1: return launch_rows<4, 1>(d_pixels, g, d_sym, st);
        case 4: return launch_rows<4, 4>(d_pixels, g, d_sym, st);
        default: return launch_rows<4, 0>(d_pixels, g, d_sym, st);
    }
}

}  // namespace llc
