// coder.cu -- K2: one adaptive range coder per slice, context state resident in shared memory.
//
// Re-creates the sequential back half of llcomp::compressImage (/root/reference/llcomp.hpp:439-449):
// binarisation of the residual (putSymbol, :166-206), the 128-state adaptive bit model
// (cabac::State, :283-293, tables :252-281) indexed hash*8+ctx (:440-441), and RangeEncoder
// (:33-89) including its carry propagation (outstanding_byte / outstanding_count) and finish().
//
// One CTA (one warp) per slice.  Lane 0 runs the serial chain; the whole warp clears the state,
// streams the slice's records HBM -> shared memory a chunk ahead of the coder, and owns nothing else.
// The 63,408 reachable state bytes stay in shared memory for the whole slice, so three slices are
// resident per SM.
#include "common.cuh"
#include "kernels.cuh"

namespace llc {

__constant__ ModelTables c_tables = make_tables();

constexpr int kChunk = 256;                                  // records staged per step
constexpr int kCoderSmem = kStateBytes + 128 * 4 + 2 * kChunk * 4;

struct RangeEnc {
    uint32_t low, range;
    int held;           // outstanding_byte (llcomp.hpp:85), -1 until the first byte is latched
    uint32_t pending;   // outstanding_count (llcomp.hpp:84)
    uint8_t* out;
    uint32_t pos, cap;

    __device__ __forceinline__ void emit(uint32_t b) {
        if (pos < cap) out[pos] = (uint8_t)b;
        ++pos;                                               // keeps counting so overflow is detectable
    }
    // One pass of the loop body of renorm_encoder (llcomp.hpp:39-57).
    __device__ __forceinline__ void shift_out() {
        if (held < 0) {
            held = (int)(low >> 8);
        } else if (low <= 0xFF00u) {
            emit((uint32_t)held);
            for (; pending; --pending) emit(0xFFu);
            held = (int)(low >> 8);
        } else if (low >= 0x10000u) {
            emit((uint32_t)held + 1u);
            for (; pending; --pending) emit(0x00u);
            held = (int)((low >> 8) & 0xFFu);
        } else {
            ++pending;
        }
        low = (low & 0xFFu) << 8;
        range <<= 8;
    }
    // llcomp.hpp:60-73.  range >= 1 after the update, so one renormalisation step always suffices.
    __device__ __forceinline__ void put(uint32_t bit, uint32_t prob) {
        const uint32_t r1 = (range * prob) >> 8;
        if (bit) { low += range - r1; range = r1; } else { range -= r1; }
        if (range < 0x100u) shift_out();
    }
    __device__ __forceinline__ void finish() {               // llcomp.hpp:75-81
        range = 0xFFu; low += 0xFFu; shift_out();
        range = 0xFFu; shift_out();
    }
};

__global__ void __launch_bounds__(32) k_slice_coder(const uint32_t* __restrict__ sym, Geom g,
                                                    uint8_t* __restrict__ scratch,
                                                    uint32_t* __restrict__ slice_bytes,
                                                    int* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* state = smem;
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + kStateBytes);
    uint32_t* ring = tab + 128;

    const int lane = threadIdx.x;
    const uint64_t s = blockIdx.x;
    const Slice sl = slice_of(g, s);
    const uint32_t* in = sym + sl.sym_off;
    const uint64_t n = sl.n;

    for (int i = lane; i < kStateBytes / 16; i += 32) reinterpret_cast<uint4*>(state)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 128; i += 32) tab[i] = c_tables.entry[i];

    RangeEnc enc;
    enc.low = 0; enc.range = 0xFF00u; enc.held = -1; enc.pending = 0;     // llcomp.hpp:35
    enc.out = scratch + scratch_off(sl, s);
    enc.pos = 0; enc.cap = (uint32_t)min(scratch_cap(sl), (uint64_t)0xFFFFFFFFu);

    // stage chunk 0
    constexpr int kPerLane = kChunk / 32;
    uint32_t pre[kPerLane];
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) {
        const uint64_t k = (uint64_t)j * 32 + lane;
        ring[j * 32 + lane] = k < n ? in[k] : 0u;
    }
    __syncwarp();

    const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
    for (uint64_t ch = 0; ch < n_chunks; ++ch) {
        const uint32_t* cur = ring + (ch & 1) * kChunk;
        // issue the loads of the next chunk before the serial section so they fly underneath it
        const uint64_t nb = (ch + 1) * kChunk;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            const uint64_t k = nb + (uint64_t)j * 32 + lane;
            pre[j] = k < n ? in[k] : 0u;
        }
        if (lane == 0) {
            const int m = (int)min((uint64_t)kChunk, n - ch * kChunk);
            for (int q = 0; q < m; ++q) {
                const uint32_t rec = cur[q];
                const uint32_t hash = rec >> 11;
                const int d = ((int)(rec << 21)) >> 21;                   // sign-extend the 11-bit residual
                uint64_t* rowp = reinterpret_cast<uint64_t*>(state + hash * kSubstates);
                uint64_t row = *rowp;                                     // the 8 sub-states of this context

                auto code = [&](int ctx, uint32_t bit) {
                    const int sh = ctx * 8;
                    const uint32_t st = (uint32_t)(row >> sh) & 0xFFu;
                    const uint32_t e = tab[st];
                    enc.put(bit, e & 0xFFu);                              // llcomp.hpp:442
                    const uint32_t ns = (bit == (st & 1u)) ? (e >> 8) & 0xFFu : (e >> 16) & 0xFFu;   // :290-292
                    row ^= (uint64_t)(st ^ ns) << sh;
                };

                if (d == 0) {
                    code(0, 1u);                                          // llcomp.hpp:204
                } else {
                    const uint32_t a = (uint32_t)abs(d);
                    const int e = 31 - __clz(a);                          // :148
                    code(0, 0u);                                          // :187
                    for (int k = 0; k < e; ++k) code(min(1 + k, kELim), 1u);   // :190-192
                    code(min(1 + e, kELim), 0u);                          // :193
                    for (int k = e - 1, c = kELim + 1; k >= 0; --k, ++c)  // :195-198
                        code(min(c, kRLim), (a >> k) & 1u);
                    code(kSignCtx, d < 0 ? 1u : 0u);                      // :200-202
                }
                *rowp = row;
            }
        }
        __syncwarp();
        uint32_t* nxt = ring + ((ch + 1) & 1) * kChunk;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) nxt[j * 32 + lane] = pre[j];
        __syncwarp();
    }

    if (lane == 0) {
        enc.finish();                                                     // llcomp.hpp:449
        slice_bytes[s] = enc.pos;
        if (enc.pos > enc.cap) atomicCAS(status, kDevOk, kDevOverflow);
    }
}

cudaError_t configure_slice_coder() {
    return cudaFuncSetAttribute(k_slice_coder, cudaFuncAttributeMaxDynamicSharedMemorySize, kCoderSmem);
}

cudaError_t launch_slice_coder(const uint32_t* d_sym, const Geom& g, uint8_t* d_scratch, uint32_t* d_slice_bytes,
                               int* d_status, cudaStream_t st) {
    const uint64_t ns = g.n_slices();
    if (ns == 0 || ns > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    k_slice_coder<<<(unsigned)ns, 32, kCoderSmem, st>>>(d_sym, g, d_scratch, d_slice_bytes, d_status);
    return cudaGetLastError();
}

}  // namespace llc
