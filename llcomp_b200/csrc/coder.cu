// coder.cu -- K2: the adaptive binary range coder of every slice, split where the algorithm splits.
//
// Re-creates the sequential back half of llcomp::compressImage (/root/reference/llcomp.hpp:439-449):
// binarisation of the residual (putSymbol, :166-206), the 128-state adaptive bit model
// (cabac::State, :283-293, tables :252-281) indexed hash*8+ctx (:440-441), and RangeEncoder (:33-89)
// with its carry propagation (outstanding_byte / outstanding_count) and finish().
//
// The reference interleaves two recurrences per binary decision; they are independent of each other:
//   (a) the probability state of context (hash, ctx) depends only on the earlier bits of THAT context;
//   (b) low/range of the coder depend on the (bit, probability) sequence, not on the states.
// K2a  k_model_pass   one warp per slice, the slice's 63,408 state bytes in shared memory.  32 samples per
//                     step, one per lane; lanes whose samples share a context are chained in sample order
//                     (match.any groups, the group leader walks its members with the 8 sub-states in
//                     registers).  Emits one 16-bit entry per decision into the bin queue in HBM.
// K2b  k_range_pass   the irreducible serial chain: x = range*M + A, renormalise, carry/byte emission.
//                     No shared-memory state, so every slice of the batch is resident at once; S slices
//                     share a warp (32/S lanes each) when there are more slices than schedulers.
#include "common.cuh"
#include "kernels.cuh"

namespace llc {

__constant__ ModelTables c_tables = make_tables();

constexpr unsigned kFull = 0xFFFFFFFFu;

// Bin-queue entry (u16): bits 0..8 = M, bit 15 = "bit is 0".  With P = P(bit=1)*256 of the context:
//   bit 1: M = P,       A = 0     range' = (range*M + A) >> 8 = range*P >> 8              (llcomp.hpp:62,69)
//   bit 0: M = 256 - P, A = 255   range' = range - (range*P >> 8) = ceil(range*(256-P)/256) (llcomp.hpp:66)
// and low += range - range' exactly when bit is 1 (llcomp.hpp:68).  M = 256, A = 0 is a no-op (padding).
constexpr uint32_t kNoopEntry = 0x0100u;
// The front end counts the decisions of every slice exactly, so the host lays the queue out without slack
// beyond kQueuePad no-op entries per slice (whole-block reads of the range pass) and 16-byte alignment.


// ---------------------------------------------------------------------------------------------------
// K2a
// ---------------------------------------------------------------------------------------------------
constexpr int kModelSmem = kStateBytes + 256 * 4;

// One decision of sub-state `ctx` (compile-time byte of the row half `half`): look the entry up, store it,
// advance the state (llcomp.hpp:440-443).  tab2[s*2+bit] = entry | next_state << 16.
template <int kByte>
__device__ __forceinline__ void model_step(bool active, uint32_t& half, uint32_t bit, const uint32_t* tab2,
                                           uint16_t* q) {
    if (active) {
        const uint32_t s = (half >> (8 * kByte)) & 0xFFu;
        const uint32_t w = tab2[s * 2 + bit];
        *q = (uint16_t)w;
        half = __byte_perm(half, w, kByte == 0 ? 0x3216 : kByte == 1 ? 0x3260 : kByte == 2 ? 0x3610 : 0x6210);
    }
}

__global__ void __launch_bounds__(32) k_model_pass(const uint32_t* __restrict__ sym, Geom g, uint64_t s0,
                                                   uint16_t* __restrict__ queue,
                                                   const uint64_t* __restrict__ q_off) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint2* state = reinterpret_cast<uint2*>(smem);                         // one 8-byte row per context
    uint32_t* tab2 = reinterpret_cast<uint32_t*>(smem + kStateBytes);

    const int lane = threadIdx.x;
    const uint64_t s = s0 + blockIdx.x;
    const Slice sl = slice_of(g, s);
    const uint32_t* in = sym + sl.sym_off;
    const uint64_t n = sl.n;
    uint16_t* q = queue + q_off[s];

    for (int i = lane; i < kStateBytes / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 256; i += 32) {
        const uint32_t st = i >> 1, b = i & 1, e = c_tables.entry[st];
        const uint32_t p = e & 0xFFu;
        const uint32_t ns = (b == (st & 1u)) ? (e >> 8) & 0xFFu : (e >> 16) & 0xFFu;   // llcomp.hpp:290-292
        tab2[i] = (b ? p : (256u - p) | 0x8000u) | (ns << 16);
    }
    __syncwarp();

    uint64_t qpos = 0;                                                    // entries written so far (warp-uniform)
    uint32_t rec_next = lane < n ? in[lane] : 0u;
    for (uint64_t base = 0; base < n; base += 32) {
        const uint32_t rec = rec_next;
        const bool valid = base + lane < n;
        {
            const uint64_t k = base + 32 + lane;
            rec_next = k < n ? in[k] : 0u;                                // next step's record flies under this step
        }
        const uint32_t hash = rec >> 11;
        const int d = ((int)(rec << 21)) >> 21;
        const uint32_t a = (uint32_t)abs(d);
        const int e = a ? 31 - __clz(a) : 0;
        const uint32_t nb = valid ? (a ? 2u * e + 3u : 1u) : 0u;

        uint32_t off = nb;                                                // exclusive scan of the bin counts
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, off, dlt);
            if (lane >= dlt) off += y;
        }
        const uint32_t total = __shfl_sync(kFull, off, 31);
        off -= nb;

        // lanes with the same context form a chain; its first lane carries the row through the members
        const uint32_t key = valid ? hash : (0x10000u | lane);
        uint32_t members = __match_any_sync(kFull, key);
        const bool leader = valid && (__ffs(members) - 1 == lane);
        const int rounds = __reduce_max_sync(kFull, valid ? __popc(members) : 0);
        uint2 row = make_uint2(0, 0);
        if (leader) row = state[hash];

        for (int r = 0; r < rounds; ++r) {
            const bool has = leader && members != 0;
            const int src = has ? __ffs(members) - 1 : lane;
            members &= members - 1;
            const int md = __shfl_sync(kFull, d, src);
            const uint32_t mo = __shfl_sync(kFull, off, src);
            const uint32_t ma = (uint32_t)abs(md);
            const int me = ma ? 31 - __clz(ma) : 0;
            uint16_t* qs = q + qpos + mo;

            model_step<0>(has, row.x, ma == 0, tab2, qs);                                 // :187 / :204
            if (__any_sync(kFull, has && ma)) {
                model_step<1>(has && ma, row.x, me >= 1, tab2, qs + 1);                   // :190-193, ctx min(1+k,4)
                model_step<2>(has && me >= 1, row.x, me >= 2, tab2, qs + 2);
                model_step<3>(has && me >= 2, row.x, me >= 3, tab2, qs + 3);
                for (int j = 0; __any_sync(kFull, has && me >= 3 && j <= me - 3); ++j)
                    model_step<0>(has && me >= 3 && j <= me - 3, row.y, j < me - 3, tab2, qs + 4 + j);   // ctx 4
                model_step<1>(has && me >= 1, row.y, (ma >> max(me - 1, 0)) & 1u, tab2, qs + me + 2);         // ctx 5, :195-198
                for (int j = 0; __any_sync(kFull, has && me >= 2 && j <= me - 2); ++j)
                    model_step<2>(has && me >= 2 && j <= me - 2, row.y, (ma >> max(me - 2 - j, 0)) & 1u, tab2,
                                  qs + me + 3 + j);                                                       // ctx 6
                model_step<3>(has && ma, row.y, md < 0, tab2, qs + 2 * me + 2);                         // ctx 7, :200-202
            }
        }
        if (leader) state[hash] = row;
        __syncwarp();
        qpos += total;
    }

    // pad with no-ops so the range pass can read whole blocks
    for (int i = lane; i < kQueuePad; i += 32) q[qpos + i] = (uint16_t)kNoopEntry;
}

// ---------------------------------------------------------------------------------------------------
// K2b
// ---------------------------------------------------------------------------------------------------
struct RangeEnc {
    uint32_t low, range;
    int held;           // outstanding_byte (llcomp.hpp:85), -1 until the first byte is latched
    uint32_t pending;   // outstanding_count (llcomp.hpp:84)
    uint8_t* out;
    uint32_t pos, cap;
    bool owner;         // one lane of the slice's lane group writes; all of them count

    __device__ __forceinline__ void emit(uint32_t b) {
        if (owner && pos < cap) out[pos] = (uint8_t)b;
        ++pos;                                               // keeps counting so overflow is detectable
    }
    // Byte/carry half of one pass of renorm_encoder's loop body (llcomp.hpp:40-55).
    __device__ __forceinline__ void shift_low() {
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
    }
    // One decision, llcomp.hpp:60-73 in the (M, A) form described above.
    __device__ __forceinline__ void put(uint32_t entry) {
        const uint32_t m = entry & 0x1FFu;
        const uint32_t a = (entry & 0x8000u) ? 255u : 0u;
        const uint32_t x = range * m + a;
        const uint32_t r = x >> 8;
        if (a == 0) low += range - r;
        if (x < 0x10000u) {                                  // range' < 0x100: renormalise once (range' >= 1)
            shift_low();
            range = x & 0xFFFFFF00u;                         // == r << 8
        } else {
            range = r;
        }
    }
    __device__ __forceinline__ void finish() {               // llcomp.hpp:75-81
        low += 0xFFu; shift_low();                           // range = 0xFF both times: always renormalises
        shift_low();
    }
};

template <int S>
__global__ void __launch_bounds__(32) k_range_pass(const uint16_t* __restrict__ queue,
                                                   const uint64_t* __restrict__ q_off,
                                                   const unsigned long long* __restrict__ n_bins, Geom g, uint64_t s0,
                                                   uint32_t n_launch, uint8_t* __restrict__ scratch,
                                                   uint32_t* __restrict__ slice_bytes, int* __restrict__ status) {
    constexpr int L = 32 / S;                                // lanes per slice
    constexpr int kBlock = L * 8;                            // entries per refill of one slice
    __shared__ uint4 stage[2][32];

    const int lane = threadIdx.x, grp = lane / L, sub = lane % L;
    const uint32_t k = blockIdx.x * S + grp;                 // slice index within this launch
    const bool live = k < n_launch;
    const uint64_t s = s0 + (live ? k : 0);
    const Slice sl = slice_of(g, s);
    const uint4* src = reinterpret_cast<const uint4*>(queue + q_off[s]) + sub;
    const uint64_t nb = live ? n_bins[s] : 0;
    const uint32_t my_blocks = (uint32_t)((nb + kBlock - 1) / kBlock);
    const uint32_t n_blocks = __reduce_max_sync(kFull, my_blocks);
    const uint4 noop = make_uint4(kNoopEntry * 0x10001u, kNoopEntry * 0x10001u, kNoopEntry * 0x10001u,
                                  kNoopEntry * 0x10001u);

    RangeEnc enc;
    enc.low = 0; enc.range = 0xFF00u; enc.held = -1; enc.pending = 0;     // llcomp.hpp:35
    enc.out = scratch + scratch_off(sl, s);
    enc.pos = 0; enc.cap = (uint32_t)min(scratch_cap(sl), (uint64_t)0xFFFFFFFFu);
    enc.owner = live && sub == 0;

    uint4 r0 = 0 < my_blocks ? src[0] : noop;
    uint4 r1 = 1 < my_blocks ? src[L] : noop;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        stage[b & 1][lane] = r0;
        __syncwarp();
        r0 = r1;
        r1 = b + 2 < my_blocks ? src[(size_t)(b + 2) * L] : noop;         // two refills ahead of the chain
#pragma unroll 1
        for (int j = 0; j < L; ++j) {
            const uint4 w = stage[b & 1][grp * L + j];
            enc.put(w.x & 0xFFFFu); enc.put(w.x >> 16);
            enc.put(w.y & 0xFFFFu); enc.put(w.y >> 16);
            enc.put(w.z & 0xFFFFu); enc.put(w.z >> 16);
            enc.put(w.w & 0xFFFFu); enc.put(w.w >> 16);
        }
    }
    enc.finish();                                                         // llcomp.hpp:449
    if (enc.owner) {
        slice_bytes[s] = enc.pos;
        if (enc.pos > enc.cap) atomicCAS(status, kDevOk, kDevOverflow);
    }
}

// ---------------------------------------------------------------------------------------------------
cudaError_t configure_slice_coder() {
    return cudaFuncSetAttribute(k_model_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, kModelSmem);
}

cudaError_t launch_model_pass(const uint32_t* d_sym, const Geom& g, uint64_t s0, uint64_t count, uint16_t* d_queue,
                              const uint64_t* d_qoff, cudaStream_t st) {
    if (count == 0 || count > 0x3FFFFFFFull) return cudaErrorInvalidValue;
    k_model_pass<<<(unsigned)count, 32, kModelSmem, st>>>(d_sym, g, s0, d_queue, d_qoff);
    return cudaGetLastError();
}

cudaError_t launch_range_pass(const uint16_t* d_queue, const uint64_t* d_qoff, const unsigned long long* d_nbins,
                              const Geom& g, uint64_t s0, uint64_t count, uint8_t* d_scratch, uint32_t* d_slice_bytes,
                              int* d_status, cudaStream_t st) {
    if (count == 0 || count > 0x3FFFFFFFull) return cudaErrorInvalidValue;
    // one slice per warp until the warps outnumber ~2 per scheduler (4 x 148 of them), then share warps
    const unsigned n = (unsigned)count;
    if (n <= 1536) k_range_pass<1><<<n, 32, 0, st>>>(d_queue, d_qoff, d_nbins, g, s0, n, d_scratch, d_slice_bytes, d_status);
    else if (n <= 3072) k_range_pass<2><<<(n + 1) / 2, 32, 0, st>>>(d_queue, d_qoff, d_nbins, g, s0, n, d_scratch, d_slice_bytes, d_status);
    else k_range_pass<4><<<(n + 3) / 4, 32, 0, st>>>(d_queue, d_qoff, d_nbins, g, s0, n, d_scratch, d_slice_bytes, d_status);
    return cudaGetLastError();
}

}  // namespace llc
