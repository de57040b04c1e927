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
// and (b) itself splits into the range recurrence (serial, 4 instructions per decision) and the low/carry/byte
// side, which is a sum of per-decision increments between renormalisations and runs lane-parallel.
//
//   k_slice_coder_fused   default.  Warp roles: model (a), range chain, bytes; the warps hand 256-decision blocks
//                         to each other through shared memory; one chain warp serves up to four slices.
//   k_model_pass + k_range_pass_ws   the same work as two kernels with a 2-byte-per-decision queue in HBM
//                         between them (LLCOMP_CODER_SPLIT=1; kept as a cross-check of the fused kernel).
// The slice's 63,408 bytes of state rows live in shared memory while every slice of the launch finds a slot,
// else in global memory behind L1.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "sample.cuh"

namespace llc {

__constant__ ModelTables c_tables = make_tables();
__constant__ QuantBytes c_quant_bytes_coder = make_quant_bytes();

constexpr unsigned kFull = 0xFFFFFFFFu;

// Shared memory through 32-bit shared-window addresses: a generic pointer that is bumped or indexed in a loop makes
// the compiler carry 64-bit address arithmetic through the warp-serial loops of the fused coder.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// Bin-queue entry (u16): bits 0..7 = M, bit 15 = "the decision is a 0".  With P = P(bit=1)*256 of the context:
//   bit 1: M = P,       A = 0     range' = (range*M + A) >> 8 = range*P >> 8                (llcomp.hpp:62,69)
//   bit 0: M = 256 - P, A = 255   range' = range - (range*P >> 8) = ceil(range*(256-P)/256)   (llcomp.hpp:66)
// and low += range - range' exactly when the decision is a 1 (llcomp.hpp:68).  P is in [7,247], so M fits a byte.
// In the split form the front end counts the decisions of every slice exactly, so the host lays the queue out
// without slack beyond kQueuePad entries per slice (16-byte alignment, reads of the last partial vector).

// ---------------------------------------------------------------------------------------------------
// Model pass
// ---------------------------------------------------------------------------------------------------
constexpr int kRowBytesSmem = kStateBytes;                   // 7926 rows x 8 sub-states, a multiple of 16
constexpr int kModelSmem = kRowBytesSmem + 256 * 4;          // + tab2

// Where the model pass puts its 16-bit entries: the bin queue in HBM (split kernels) or a shared-memory FIFO
// (fused kernel).  `pos` counts from the first decision of the current 32-sample step.
// A sink turns the position of a sample's first decision into an address once (at) and then takes the sample's
// entries at small offsets from it (put): the offsets are compile-time constants or running sums, so an entry costs
// one store and no address arithmetic.
struct QueueSink {
    uint16_t* q;
    typedef uint16_t* Addr;
    __device__ __forceinline__ Addr at(uint32_t pos) const { return q + pos; }
    __device__ __forceinline__ void put(Addr a, int k, uint32_t w) const { a[k] = (uint16_t)w; }
};

// One step of the model pass: 32 consecutive samples, one per lane (rec = packed record of this lane's sample),
// in two halves so that a caller can run the first half of step i+1 (warp votes and the ~330-cycle match, none of
// which touch the state rows) under the second half of step i.
// llcomp.hpp:166-206 (binarisation), :440-443 (state).
struct StepPlan {
    uint32_t off;        // position of the lane's first decision within the step
    uint32_t total;      // decisions of the step
    uint32_t members;    // lanes with this lane's context (lanes without a sample: themselves)
    uint32_t zero_mask;  // lanes whose residual is zero
};

__device__ __forceinline__ int residual_of(uint32_t rec) { return ((int)(rec << 21)) >> 21; }

__device__ __forceinline__ StepPlan model_plan(uint32_t rec, bool valid, int lane) {
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t a = (uint32_t)abs(residual_of(rec));
    const int e = a ? 31 - __clz(a) : -1;                                 // -1 marks a zero residual
    const uint32_t nb = valid ? 2u * e + 3u : 0u;                         // 1 decision for zero, else 2e+3
    StepPlan p;
    // exclusive scan of the decision counts (<= 19, five bit planes, no dependent shuffles)
    p.off = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const uint32_t m = __ballot_sync(kFull, (nb >> b) & 1u);
        p.off += __popc(m & lt_mask) << b;
    }
    p.total = __shfl_sync(kFull, p.off + nb, 31);
    // lanes with the same context form a chain; its first lane carries the row through the members
    p.members = __match_any_sync(kFull, valid ? rec >> 11 : (0x10000u | lane));
    p.zero_mask = __ballot_sync(kFull, valid && a == 0);
    return p;
}

// The row of the lane's context, fetched by the first lane of every chain (zero elsewhere).
__device__ __forceinline__ uint2 model_row(uint32_t rec, bool valid, const StepPlan& plan, const uint2* state, int lane) {
    const bool leader = valid && __ffs(plan.members) - 1 == lane;
    uint2 row = make_uint2(0, 0);
    if (leader) row = state[rec >> 11];
    return row;
}

template <class Sink>
__device__ __forceinline__ uint32_t model_apply(uint32_t rec, bool valid, const StepPlan& plan, uint2 row, uint2* state,
                                                const uint32_t* tab2, int lane, Sink sink) {
    const uint32_t hash = rec >> 11;
    const int d = residual_of(rec);
    const uint32_t off = plan.off;
    uint32_t members = plan.members;
    const bool mine = valid;
    const int head = __ffs(members) - 1;
    const bool leader = mine && head == lane;
    // Smooth content: every member of the chain has a zero residual (one decision, "is zero" = 1, sub-state 0) and
    // that sub-state sits in the saturated state 127 (MPS 1, next-if-MPS 127, llcomp.hpp:258).  Then every member
    // gets the same entry and the row does not change: no need to walk the chain member by member.
    const bool flat = leader && (members & ~plan.zero_mask) == 0 && (row.x & 0xFFu) == 127u;
    if (__any_sync(kFull, flat)) {
        const bool flat_chain = __shfl_sync(kFull, flat, head);          // every lane takes part in the shuffle
        if (mine && flat_chain) sink.put(sink.at(off), 0, tab2[127 * 2 + 1]);
        if (flat) members = 0;
    }
    const int rounds = __reduce_max_sync(kFull, leader ? __popc(members) : 0);
    uint32_t s0b = row.x & 0xFFu, s1b = (row.x >> 8) & 0xFFu, s2b = (row.x >> 16) & 0xFFu, s3b = row.x >> 24;
    uint32_t s4b = row.y & 0xFFu, s5b = (row.y >> 8) & 0xFFu, s6b = (row.y >> 16) & 0xFFu, s7b = row.y >> 24;

    for (int r = 0; r < rounds; ++r) {
        const bool has = leader && members != 0;
        const int src = has ? __ffs(members) - 1 : lane;
        members &= members - 1;
        const int md = __shfl_sync(kFull, d, src);
        const uint32_t mo = __shfl_sync(kFull, off, src);
        const uint32_t ma = (uint32_t)abs(md);
        const int me = ma ? 31 - __clz(ma) : -1;
        const int maxe = __reduce_max_sync(kFull, has ? me : -1);
        // entries of the member: zero flag at +0, exponent at +1 .. +e+1, mantissa at +e+2 .. +2e+1, sign at +2e+2
        const typename Sink::Addr a0 = sink.at(mo);
        const typename Sink::Addr ae = a0 + max(me, 0);

        // the sub-states are independent of one another: issue every look-up of the straight part first
        const uint32_t w0 = tab2[s0b * 2 + (me < 0)];                                   // ctx 0, :187 / :204
        const uint32_t w1 = tab2[s1b * 2 + (me >= 1)];                                  // ctx 1..3, :190-193
        const uint32_t w2 = tab2[s2b * 2 + (me >= 2)];
        const uint32_t w3 = tab2[s3b * 2 + (me >= 3)];
        const uint32_t w5 = tab2[s5b * 2 + ((ma >> max(me - 1, 0)) & 1u)];             // ctx 5, first mantissa bit
        const uint32_t w7 = tab2[s7b * 2 + (md < 0)];                                   // ctx 7, sign, :200-202
        if (has) { sink.put(a0, 0, w0); s0b = w0 >> 16; }
        if (maxe >= 0) {
            if (has && me >= 0) {
                sink.put(a0, 1, w1); s1b = w1 >> 16;
                sink.put(ae + max(me, 0), 2, w7); s7b = w7 >> 16;
            }
            if (has && me >= 1) {
                sink.put(a0, 2, w2); s2b = w2 >> 16;
                sink.put(ae, 2, w5); s5b = w5 >> 16;
            }
            if (has && me >= 2) { sink.put(a0, 3, w3); s3b = w3 >> 16; }
            typename Sink::Addr a4 = a0;
            for (int j = 0; j <= maxe - 3; ++j) {                                       // ctx 4: positions 4..e+1
                const uint32_t w4 = tab2[s4b * 2 + (j < me - 3)];
                if (has && j <= me - 3) { sink.put(a4, 4, w4); s4b = w4 >> 16; }
                a4 += 1;
            }
            typename Sink::Addr a6 = ae;
            for (int j = 0; j <= maxe - 2; ++j) {                                       // ctx 6: mantissa bits e-2..0
                const uint32_t w6 = tab2[s6b * 2 + ((ma >> max(me - 2 - j, 0)) & 1u)];
                if (has && j <= me - 2) { sink.put(a6, 3, w6); s6b = w6 >> 16; }
                a6 += 1;
            }
        }
    }
    if (leader)
        state[hash] = make_uint2(s0b | (s1b << 8) | (s2b << 16) | (s3b << 24),
                                     s4b | (s5b << 8) | (s6b << 16) | (s7b << 24));
    __syncwarp();
    return plan.total;
}

template <class Sink>
__device__ __forceinline__ uint32_t model_chunk(uint32_t rec, bool valid, uint2* state,
                                                const uint32_t* tab2, int lane, Sink sink) {
    const StepPlan plan = model_plan(rec, valid, lane);
    return model_apply(rec, valid, plan, model_row(rec, valid, plan, state, lane), state, tab2, lane, sink);
}

// Table of the model pass: [state*2 + bit] = queue entry | next_state << 16 (llcomp.hpp:252-281, :290-292).
__device__ __forceinline__ void fill_tab2(uint32_t* tab2, int lane) {
    for (int i = lane; i < 256; i += 32) {
        const uint32_t st = i >> 1, b = i & 1, e = c_tables.entry[st];
        const uint32_t p = e & 0xFFu;
        const uint32_t ns = (b == (st & 1u)) ? (e >> 8) & 0xFFu : (e >> 16) & 0xFFu;
        tab2[i] = (b ? p : (256u - p) | 0x8000u) | (ns << 16);
    }
}

// kGlobalState: the rows live in global memory, pre-zeroed by the host, and are reached through L1
// (116 vs 97 cycles per dependent read-modify-write, profiles/microbench/l1_rmw.cu); without 63 KB of shared
// memory per slice every slice of a 1024-image batch is resident at once.
template <bool kGlobalState>
__global__ void __launch_bounds__(32) k_model_pass(const uint32_t* __restrict__ sym, Geom g, uint64_t s0,
                                                   uint16_t* __restrict__ queue,
                                                   const uint64_t* __restrict__ q_off,
                                                   uint2* __restrict__ gstate) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kRowBytes = kGlobalState ? 0 : kRowBytesSmem;
    uint2* state = kGlobalState ? gstate + (size_t)blockIdx.x * kContexts
                                : reinterpret_cast<uint2*>(smem);         // one 8-byte row per context
    uint32_t* tab2 = reinterpret_cast<uint32_t*>(smem + kRowBytes);       // [state*2 + bit] = entry | next << 16

    const int lane = threadIdx.x;
    const uint64_t s = s0 + blockIdx.x;
    const Slice sl = slice_of(g, s);
    const uint32_t* in = sym + sl.sym_off;
    const uint64_t n = sl.n;
    uint16_t* q = queue + q_off[s];

    for (int i = lane; i < kRowBytes / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fill_tab2(tab2, lane);
    __syncwarp();

    uint32_t rec_next = lane < n ? in[lane] : 0u;
    for (uint64_t base = 0; base < n; base += 32) {
        const uint32_t rec = rec_next;
        const bool valid = base + lane < n;
        {
            const uint64_t k = base + 32 + lane;
            rec_next = k < n ? in[k] : 0u;                                // next step's record flies under this step
        }
        q += model_chunk(rec, valid, state, tab2, lane, QueueSink{q});
    }
}

// ---------------------------------------------------------------------------------------------------
// K2b
// ---------------------------------------------------------------------------------------------------
struct ByteTail {           // byte/carry side of the encoder
    uint32_t low;
    uint32_t hp;            // bits 0..7 outstanding_byte (llcomp.hpp:85), bit 8 = nothing latched yet,
                            // bits 9.. outstanding_count (llcomp.hpp:84); "plain" state <=> hp < 0x100
    uint8_t* outp;          // where the next byte goes (all lanes of a slice store the same byte there)
};
constexpr uint32_t kHpEmpty = 0x100u;

// Everything of renorm_encoder's loop body (llcomp.hpp:40-54) that the inlined fast path does not cover:
// the first latch, deferred 0xFF bytes and their flush.  Rare, so kept out of line to keep the chain's code small.
__device__ __noinline__ ByteTail renorm_slow(ByteTail t) {
    uint32_t held = t.hp & 0xFFu, pending = t.hp >> 9;
    if (t.hp & kHpEmpty) {
        held = t.low >> 8;                                   // first byte: just latch (low < 0x10000 here)
    } else if (t.low <= 0xFF00u) {
        *t.outp++ = (uint8_t)held;
        for (; pending; --pending) *t.outp++ = 0xFFu;
        held = t.low >> 8;
    } else if (t.low >= 0x10000u) {
        *t.outp++ = (uint8_t)(held + 1);
        for (; pending; --pending) *t.outp++ = 0x00u;
        held = (t.low >> 8) & 0xFFu;
    } else {
        ++pending;
    }
    t.hp = held | (pending << 9);
    return t;
}

// prmt with the selector's "replicate the sign of the selected byte" bit (bit 3 of a nibble), which
// __byte_perm masks off: turns the flag bit of an entry into A = 0x00 / 0xFF in one instruction.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t prmt2(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// One pass of renorm_encoder's loop body (llcomp.hpp:40-55) on the byte side.
__device__ __forceinline__ void shift_low(ByteTail& t) {
    // plain byte or carry with nothing deferred: emit held (+1 on carry), latch the next byte
    if (t.hp < kHpEmpty && (t.low - 0xFF01u) >= 0xFFu) {
        *t.outp++ = (uint8_t)(t.hp + (t.low >> 16));
        t.hp = (t.low >> 8) & 0xFFu;
    } else {
        t = renorm_slow(t);
    }
    t.low = (t.low & 0xFFu) << 8;
}

constexpr int kBlk = 256;                                    // decisions per block of the split range pass

// Byte side of one block of `cnt` decisions (helper warp, lane-parallel over 32 decisions at a time): x values
// from the chain, A operands (0 <=> the decision is a 1) from the operand ring.  Exact re-statement of the
// low/carry half of RangeEncoder::put + renorm_encoder (llcomp.hpp:38-73).
__device__ __forceinline__ void byte_side_block(ByteTail& t, bool& overflow, uint32_t& x_carry, const uint32_t* xr,
                                                const uint2* inr, uint32_t cnt, int lane, uint8_t* out0,
                                                uint8_t* out_end) {
    for (uint32_t base = 0; base < cnt; base += 32) {
        const bool live = base + lane < cnt;
        const uint32_t x = live ? xr[base + lane] : 0x01000000u;      // inert: no renorm, delta 0
        const uint32_t a = live ? inr[base + lane].y : 255u;
        uint32_t xp = __shfl_up_sync(kFull, x, 1);
        if (lane == 0) xp = x_carry;
        x_carry = __shfl_sync(kFull, x, 31);
        const bool any_dead = __any_sync(kFull, !live);
        if (any_dead) {                          // keep the carry at the last live decision
            const uint32_t last = cnt - base - 1;
            x_carry = __shfl_sync(kFull, x, last);
        }
        const uint32_t r_before = xp < 0x10000u ? (xp & 0xFFFFFF00u) : (xp >> 8);
        const uint32_t delta = (live && a == 0) ? r_before - (x >> 8) : 0u;
        const uint32_t F = __ballot_sync(kFull, live && x < 0x10000u);   // renormalising decisions
        uint32_t pre = delta;                    // inclusive prefix sum of the low increments
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, pre, d);
            if (lane >= d) pre += y;
        }
        const uint32_t total = __shfl_sync(kFull, pre, 31);
        if (F == 0) { t.low += total; continue; }
        // segment sums: segment of lane i starts after the last renormalisation below i
        const uint32_t below = F & ((1u << lane) - 1u);
        const int p = below ? 31 - __clz(below) : -1;           // previous renormalising lane
        const uint32_t pre_p = __shfl_sync(kFull, pre, p < 0 ? 0 : p);
        const uint32_t seg = pre - (p < 0 ? 0u : pre_p);         // S_j when lane renormalises
        const int first = __ffs(F) - 1;
        // low at this lane's renormalisation (valid on lanes of F)
        const uint32_t seg_p = __shfl_sync(kFull, seg, p < 0 ? 0 : p);
        const uint32_t low_first_part = t.low;                   // low entering the sub-block
        uint32_t low_ev;
        if (p < 0) low_ev = low_first_part + seg;
        else low_ev = (((seg_p + (p == first ? low_first_part : 0u)) & 0xFFu) << 8) + seg;
        // feed the renormalisations through the byte/carry machine (llcomp.hpp:40-55)
        if (t.outp + (t.hp >> 9) + 32 + 8 > out_end) { overflow = true; t.outp = out0; t.hp &= 0x1FFu; }
        const bool is_ev = (F >> lane) & 1u;
        const int last_lane = 31 - __clz(F);
        // Plain regime: a byte is latched, nothing is deferred, and no renormalisation of this sub-block
        // defers one (low in 0xFF01..0xFFFF).  Then renormalisation j emits the byte latched by j-1 plus
        // its own carry, independently of all the others: every lane writes its own byte.
        const bool defers = is_ev && (low_ev - 0xFF01u) < 0xFFu;
        if (t.hp < kHpEmpty && !__any_sync(kFull, defers)) {
            const uint32_t low_p = __shfl_sync(kFull, low_ev, p < 0 ? 0 : p);
            const uint32_t held = p < 0 ? t.hp : (low_p >> 8) & 0xFFu;
            if (is_ev) t.outp[__popc(below)] = (uint8_t)(held + (low_ev >> 16));
            const uint32_t low_last = __shfl_sync(kFull, low_ev, last_lane);
            t.outp += __popc(F);
            t.hp = (low_last >> 8) & 0xFFu;
            t.low = (low_last & 0xFFu) << 8;
        } else {
            uint32_t todo = F;
            do {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                t.low = __shfl_sync(kFull, low_ev, i);
                shift_low(t);
            } while (todo);
        }
        // decisions after the last renormalisation of the sub-block
        const uint32_t pre_last = __shfl_sync(kFull, pre, last_lane);
        t.low += total - pre_last;
    }
}

// ---------------------------------------------------------------------------------------------------
// K2b, warp-specialised form (one CTA of two warps per slice).
//   chain warp   only the range recurrence x = range*M + A; range' = x < 0x10000 ? x & ~0xFF : x >> 8, four
//                instructions per decision, operands pre-expanded to (M, A) pairs in shared memory, x written
//                back to shared memory.  This is the irreducible serial dependency of the encoder.
//   helper warp  everything else, lane-parallel over 32 decisions at a time: expands the next block of queue
//                entries; from the x values of the previous block derives range-before/after, the low
//                increments (segmented sums between renormalisations), low at every renormalisation
//                (low_j = ((S_{j-1} & 0xFF) << 8) + S_j, because low mod 256 only depends on the last segment),
//                and feeds those through the byte/carry machine of renorm_encoder (llcomp.hpp:38-58).
// One __syncthreads per 256-decision block; both rings are double buffered.
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void pair_sync() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__global__ void __launch_bounds__(128) k_range_pass_ws(const uint16_t* __restrict__ queue,
                                                      const uint64_t* __restrict__ q_off,
                                                      const unsigned long long* __restrict__ n_bins, Geom g, uint64_t s0,
                                                      uint8_t* __restrict__ scratch, uint32_t* __restrict__ slice_bytes,
                                                      int* __restrict__ status) {
    __shared__ __align__(16) uint2 in_ring[2][kBlk + 4];     // (M, A) per decision (+ slack for the look-ahead load)
    __shared__ __align__(16) uint32_t x_ring[2][kBlk];       // x per decision

    // Four warp slots per CTA, two used: which two rotates with the CTA index so that the chain warps of the
    // CTAs sharing an SM spread over its four schedulers (warp slot w of a CTA issues on scheduler w).
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const int chain_slot = blockIdx.x & 3, helper_slot = (blockIdx.x + 2) & 3;
    if (wslot != chain_slot && wslot != helper_slot) return;
    const bool is_chain = wslot == chain_slot;
    const uint64_t s = s0 + blockIdx.x;
    const Slice sl = slice_of(g, s);
    const uint16_t* q = queue + q_off[s];
    const uint64_t nb = n_bins[s];
    const uint32_t n_blk = (uint32_t)((nb + kBlk - 1) / kBlk);

    // helper state
    uint8_t* const out0 = scratch + scratch_off(sl, s);
    uint8_t* const out_end = out0 + scratch_cap(sl);
    ByteTail t;
    t.low = 0; t.hp = kHpEmpty; t.outp = out0;               // llcomp.hpp:35
    bool overflow = false;
    uint32_t x_carry = 0xFF00u << 8;                         // pseudo-x whose successor range is the initial 0xFF00
    // chain state
    uint32_t range = 0xFF00u;

    auto load_entries = [&](uint32_t b) -> uint4 {          // lane's 8 entries of block b (0 beyond the end)
        const uint64_t first = (uint64_t)b * kBlk + lane * 8;
        if (first + 8 <= nb) return reinterpret_cast<const uint4*>(q)[first / 8];
        uint32_t w[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; ++k)
            if (first + k < nb) w[k >> 1] |= (uint32_t)q[first + k] << (16 * (k & 1));
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    auto expand = [&](uint4 e, int buf) {                    // 8 entries -> 8 (M, A) pairs
        const uint32_t w[4] = {e.x, e.y, e.z, e.w};
        uint4* dst = reinterpret_cast<uint4*>(&in_ring[buf][lane * 8]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            dst[k] = make_uint4(w[k] & 0xFFu, prmt(w[k], 0x4449), prmt(w[k], 0x4442), prmt(w[k], 0x444B));
    };

    uint4 e_next = make_uint4(0, 0, 0, 0);
    if (!is_chain) {
        if (n_blk > 0) expand(load_entries(0), 0);
        if (n_blk > 1) e_next = load_entries(1);
    }
    pair_sync();

    for (uint32_t b = 0; b <= n_blk; ++b) {                  // iteration b: chain on block b, helper on block b-1
        if (is_chain) {
            if (b < n_blk) {
                const uint32_t cnt = (uint32_t)min((uint64_t)kBlk, nb - (uint64_t)b * kBlk);
                const uint4* in = reinterpret_cast<const uint4*>(in_ring[b & 1]);
                uint4* xo = reinterpret_cast<uint4*>(x_ring[b & 1]);
                const uint32_t n4 = (cnt + 3) / 4;           // garbage beyond cnt is computed and ignored
                // operands of the next 4 decisions are fetched while the current 4 run (the ring has slack)
                uint4 p0 = in[0], p1 = in[1];
#pragma unroll 4
                for (uint32_t v = n4; v > 0; --v) {
                    in += 2;
                    const uint4 q0 = in[0], q1 = in[1];
                    uint4 xs;
                    xs.x = range * p0.x + p0.y; range = xs.x < 0x10000u ? (xs.x & 0xFFFFFF00u) : (xs.x >> 8);
                    xs.y = range * p0.z + p0.w; range = xs.y < 0x10000u ? (xs.y & 0xFFFFFF00u) : (xs.y >> 8);
                    xs.z = range * p1.x + p1.y; range = xs.z < 0x10000u ? (xs.z & 0xFFFFFF00u) : (xs.z >> 8);
                    xs.w = range * p1.z + p1.w; range = xs.w < 0x10000u ? (xs.w & 0xFFFFFF00u) : (xs.w >> 8);
                    *xo++ = xs;
                    p0 = q0; p1 = q1;
                }
            }
        } else {
            if (b > 0) {                                     // byte side of block b-1
                const uint32_t pb = b - 1;
                const uint32_t cnt = (uint32_t)min((uint64_t)kBlk, nb - (uint64_t)pb * kBlk);
                byte_side_block(t, overflow, x_carry, x_ring[pb & 1], in_ring[pb & 1], cnt, lane, out0, out_end);
            }
            if (b + 1 < n_blk) {                             // operands of block b+1, entries of block b+2
                expand(e_next, (b + 1) & 1);
                if (b + 2 < n_blk) e_next = load_entries(b + 2);
            }
        }
        pair_sync();
    }

    if (!is_chain) {
        if (t.outp + (t.hp >> 9) + 8 > out_end) { overflow = true; t.outp = out0; t.hp &= 0x1FFu; }
        // finish(), llcomp.hpp:75-81: range = 0xFF both times, so each renorm_encoder call shifts exactly once
        t.low += 0xFFu;
        shift_low(t);
        shift_low(t);
        if (lane == 0) {
            slice_bytes[s] = overflow ? 0xFFFFFFFFu : (uint32_t)(t.outp - out0);
            if (overflow) atomicCAS(status, kDevOk, kDevOverflow);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2 fused: records -> slice payload in ONE kernel, no bin queue in HBM.  NS slices per CTA, 1 + 2 NS warps:
//   model warp   (one per slice) the model pass above, its entries go to a shared-memory FIFO instead of HBM; it
//                runs to the end of the slice at its own pace, as far ahead as the FIFO has room;
//   chain warp   (one per CTA) the range recurrence of all NS slices in lock step, 32/NS lanes each;
//   helper warp  (one per slice) expands FIFO entries to operands for the chain and runs the byte side.
// Chain and helpers meet at one barrier per 256-decision block: in iteration b the helper turns block b-1 into
// bytes and expands block b+1 (waiting for the model warp if it has to), the chain runs block b.  With the state rows behind L1 (kGlobalState) a slice needs
// ~14 KB of shared memory, so every slice of a 1024-image batch is resident at once; with the rows in shared
// memory (NS = 1, few slices) the CTA takes 79 KB.
//
// Who is the bottleneck was measured with per-role cycle counters: the serial recurrence is,
// at 13 cycles of dependent latency per decision.  Everything here is arranged to keep that warp lean:
// operands arrive pre-multiplied, one 16-byte shared load per decision, and no predicates on the dependent path.
// ---------------------------------------------------------------------------------------------------
constexpr int kBlkF = 256;                                   // decisions per block of the fused coder
constexpr int kPerLane = kBlkF / 32;                         // helper: consecutive decisions per lane
#ifndef LLC_SPIN_NS
#define LLC_SPIN_NS 1024
#endif
#ifndef LLC_AHEAD
#define LLC_AHEAD 3
#endif
constexpr int kAhead = LLC_AHEAD;                            // blocks the model warp runs ahead of the chain
constexpr int kFifoF = kAhead <= 3 ? 2048 : 4096;            // FIFO entries: kAhead blocks + one step (<= 608) fit
static_assert(kAhead * kBlkF + 32 * 19 + kBlkF <= kFifoF, "the model warp must not overrun the block being expanded");
// The FIFO is a ring of kFifoF entries followed by kFifoSpill more: the entries of one sample are stored at consecutive
// addresses from the (wrapped) position of its first decision, so the one sample per lap that straddles the end of the
// ring spills into the tail, and its lane copies the spilled entries to the start of the ring afterwards.
constexpr int kFifoSpill = 24;                               // >= 18 (a sample has at most 19 entries), 16-byte multiple
struct FifoAddr {                                            // shared-window byte address of a 16-bit entry
    uint32_t a;
    __device__ __forceinline__ FifoAddr operator+(int k) const { return FifoAddr{a + 2u * (uint32_t)k}; }
    __device__ __forceinline__ FifoAddr& operator+=(int k) { a += 2u * (uint32_t)k; return *this; }
};
struct FifoSink {
    uint32_t fifo_s;     // shared-window address of the FIFO
    uint32_t base;
    typedef FifoAddr Addr;
    __device__ __forceinline__ Addr at(uint32_t pos) const { return FifoAddr{fifo_s + (((base + pos) & (kFifoF - 1)) << 1)}; }
    __device__ __forceinline__ void put(Addr p, int k, uint32_t w) const { sts16(p.a + 2u * (uint32_t)k, w); }
};
// after a step: the lane whose sample crossed the end of the ring brings its spilled entries home
__device__ __forceinline__ void fifo_unspill(uint32_t fifo_s, uint32_t base, uint32_t off, uint32_t nb, bool valid) {
    const uint32_t start = (base + off) & (kFifoF - 1);
    const bool spilled = valid && start + nb > (uint32_t)kFifoF;
    if (__any_sync(kFull, spilled)) {
        if (spilled) {
            const uint32_t n = start + nb - kFifoF;
            for (uint32_t k = 0; k < n; ++k) {
                uint16_t v;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(fifo_s + 2u * (kFifoF + k)) : "memory");
                sts16(fifo_s + 2u * k, v);
            }
        }
        __syncwarp();
    }
}

// One decision of the range recurrence.  With x = range * M + A (24 bits),
//   range' = x < 0x10000 ? x & ~0xFF : x >> 8   (RangeEncoder::put + the renormalisation shift, llcomp.hpp:57-73)
// is (x >> 8) * f with f = 256 or 1, so x' = (x >> 8) * (M * f) + A.  The chain carries y = x + kChainBias, whose
// bit 24 is [x >= 0x10000]: the selection of f is then arithmetic, three dependent operations per decision
// (shift, multiply-add, multiply-add).  A compare feeding a select costs ~13 cycles of predicate latency, and
// this recurrence is the critical path of the whole coder.  Operand: (256 M, -255 M, A + kChainBias).
constexpr uint32_t kChainBias = 0xFF0000u;
#ifndef LLC_CHAIN_V
#define LLC_CHAIN_V 0
#endif
#if LLC_CHAIN_V == 2
// The same recurrence in fp32 (every value is an integer below 2^24, so every operation is exact), all on the FMA
// pipe: 4 + 4 + 4 cycles instead of 5 + 4 + 5 for shift -> multiply-add -> multiply-add, because a result that crosses
// between the two integer pipes costs one more cycle.  With z = x with its low byte cleared,
//   range' M' = z M' (renormalisation: range' = z)   or   (z / 256) M' (range' = x >> 8),
// so the shift moves into the multiplier's exponent: Mf = M' / 256 when x >= 0x10000 (nz = 1), else M'.
//   z  = (x + 2^31 rounded toward zero) - 2^31      ulp(2^31) = 256: the rounding clears the low byte
//   nz = saturate(x - 65535)                         0 or 1 for an integer x
// Operand: (M', M'/256 - M', A) as floats; x is carried unbiased, as a float.
__device__ __forceinline__ uint32_t chain_step(uint32_t y, const uint4& op, uint32_t) {
    const float x = __uint_as_float(y);
    const float zm = __fadd_rz(x, 2147483648.f);
    const float nz = __saturatef(x - 65535.f);
    const float z = zm - 2147483648.f;
    const float mf = fmaf(nz, __uint_as_float(op.y), __uint_as_float(op.x));
    return __float_as_uint(fmaf(z, mf, __uint_as_float(op.z)));
}
#else
__constant__ uint32_t c_two8 = 256u;                         // a multiplier ptxas cannot fold into a shift
__device__ __forceinline__ uint32_t chain_step(uint32_t y, const uint4& op, uint32_t two8) {
#if LLC_CHAIN_V == 1
    // (measured and dropped: bit 24 through the multiplier, IMAD.HI, so that the recurrence stays on one pipe: 12%
    // slower, the instruction's latency is not the 4 cycles ptxas schedules for)
    uint32_t nz;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(nz) : "r"(y), "r"(two8));
#else
    (void)two8;
    const uint32_t nz = y >> 24;                             // 1: no renormalisation
#endif
    const uint32_t a = (y >> 8) - (kChainBias >> 8);         // x >> 8
    const uint32_t mf = nz * op.y + op.x;                    // M f
    return a * mf + op.z;
}
#endif

// Operand ring of one block: decision e sits in slot (e % 8) * 32 + e / 8, so that the helper lane that expands
// decisions 8 l .. 8 l + 7 writes its j-th operand next to its neighbours' j-th operands (no bank conflicts);
// the chain reads one slot per decision either way.
constexpr int kRingSlots = kBlkF + 8;                        // + look-ahead of the chain's operand prefetch
__device__ __forceinline__ int ring_slot(int e) { return (e & 7) * 32 + (e >> 3); }

// Byte side of one block of the fused coder, lane-serial: lane l owns decisions [8 l, 8 l + 8) of the block,
// so the low increments and the renormalisation flags are found without any cross-lane traffic; only the
// renormalisations themselves (about one per ten decisions) are then compacted and handled one per lane.
//   xr      biased x values of the block from the chain; the lane's own words are reused as its event staging
//   evl     scratch for the compacted event list (kBlkF + 3 words; aliases the consumed operand ring)
//   nodelta bit 15 - i/2 (i even) / 31 - i/2 (i odd) set <=> decision 8 l + i codes a 0 (no low increment)
// Same arithmetic as byte_side_block: exact re-statement of the low/carry half of RangeEncoder::put +
// renorm_encoder (llcomp.hpp:38-73).  With E_j = low entering the block + all increments up to renormalisation j
// (E_-1 = E_-2 = E_-3 = 0), low at renormalisation j is (((E_j-1 - E_j-2) & 0xFF) << 8) + E_j - E_j-1: low mod 256
// only depends on the last segment.
__device__ __forceinline__ void byte_side_lanes(ByteTail& t, bool& overflow, uint32_t& x_carry, uint32_t* xr,
                                                uint32_t* evl, uint32_t nodelta, uint32_t cnt, int lane,
                                                uint8_t* out0, uint8_t* out_end) {
    uint32_t x[kPerLane];
    {
        const uint4* xv = reinterpret_cast<const uint4*>(xr + kPerLane * lane);
#pragma unroll
        for (int k = 0; k < kPerLane / 4; ++k) {
            const uint4 v = xv[k];
#if LLC_CHAIN_V == 2
            x[4 * k] = __float2uint_rz(__uint_as_float(v.x)); x[4 * k + 1] = __float2uint_rz(__uint_as_float(v.y));
            x[4 * k + 2] = __float2uint_rz(__uint_as_float(v.z)); x[4 * k + 3] = __float2uint_rz(__uint_as_float(v.w));
#else
            x[4 * k] = v.x - kChainBias; x[4 * k + 1] = v.y - kChainBias;      // the chain stores x + kChainBias
            x[4 * k + 2] = v.z - kChainBias; x[4 * k + 3] = v.w - kChainBias;
#endif
        }
    }
    if (cnt < (uint32_t)kBlkF) {                              // last block of a slice: the rest is inert
        const int mine = (int)cnt - kPerLane * lane;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i)
            if (i >= mine) { x[i] = 0x01000000u; nodelta |= 1u << ((i & 1) ? 31 - i / 2 : 15 - i / 2); }
    }
    uint32_t xp = __shfl_up_sync(kFull, x[kPerLane - 1], 1);
    if (lane == 0) xp = x_carry;
    x_carry = __shfl_sync(kFull, x[kPerLane - 1], 31);
    uint32_t r = xp < 0x10000u ? (xp & 0xFFFFFF00u) : (xp >> 8);   // range before the lane's first decision
    uint32_t sum = 0;
    const uint32_t stage_s = smem_addr(xr) + 4u * kPerLane * lane;
    uint32_t sp = stage_s;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const uint32_t sh = x[i] >> 8;
        if (!(nodelta & (1u << ((i & 1) ? 31 - i / 2 : 15 - i / 2)))) sum += r - sh;
        const bool ev = x[i] < 0x10000u;
        r = ev ? (x[i] & 0xFFFFFF00u) : sh;
        if (ev) { sts32(sp, sum); sp += 4; }
    }
    const uint32_t k_mine = (sp - stage_s) >> 2;
    uint32_t inc_s = sum, inc_k = k_mine;                    // inclusive scans over lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ys = __shfl_up_sync(kFull, inc_s, d), yk = __shfl_up_sync(kFull, inc_k, d);
        if (lane >= d) { inc_s += ys; inc_k += yk; }
    }
    const uint32_t e_tot = t.low + __shfl_sync(kFull, inc_s, 31);  // low after the block if nothing renormalised
    const uint32_t n_ev = __shfl_sync(kFull, inc_k, 31);
    if (n_ev == 0) { t.low = e_tot; return; }
    {
        const uint32_t off = t.low + inc_s - sum;
        const uint32_t dst_s = smem_addr(evl) + 4u * (3 + (inc_k - k_mine));
        for (uint32_t k = 0; k < k_mine; ++k) sts32(dst_s + 4 * k, lds32(stage_s + 4 * k) + off);
        if (lane < 3) evl[lane] = 0;
    }
    __syncwarp();
    uint32_t low_last = 0, e_last = 0;
    for (uint32_t j0 = 0; j0 < n_ev; j0 += 32) {
        const uint32_t nv = min(32u, n_ev - j0);
        const bool valid = (uint32_t)lane < nv;
        const uint32_t* e = evl + j0 + (valid ? lane : 0);
        const uint32_t e0 = e[3], e1 = e[2], e2 = e[1], e3 = e[0];
        const uint32_t low_j = (((e1 - e2) & 0xFFu) << 8) + (e0 - e1);
        const uint32_t low_p = (((e2 - e3) & 0xFFu) << 8) + (e1 - e2);
        if (t.outp + (t.hp >> 9) + 32 + 8 > out_end) { overflow = true; t.outp = out0; t.hp &= 0x1FFu; }
        const bool defers = valid && (low_j - 0xFF01u) < 0xFFu;
        if (t.hp < kHpEmpty && !__any_sync(kFull, defers)) {
            // plain regime (see byte_side_block): every renormalisation emits the byte latched by its
            // predecessor plus its own carry
            const uint32_t held = lane == 0 ? t.hp : (low_p >> 8) & 0xFFu;
            if (valid) t.outp[lane] = (uint8_t)(held + (low_j >> 16));
            t.outp += nv;
            const uint32_t ll = __shfl_sync(kFull, low_j, nv - 1);
            t.hp = (ll >> 8) & 0xFFu;
        } else {
            for (uint32_t i = 0; i < nv; ++i) {
                t.low = __shfl_sync(kFull, low_j, i);
                shift_low(t);
            }
        }
        low_last = __shfl_sync(kFull, low_j, nv - 1);
        e_last = __shfl_sync(kFull, e0, nv - 1);
    }
    t.low = ((low_last & 0xFFu) << 8) + (e_tot - e_last);     // decisions after the last renormalisation
    __syncwarp();
}

// Which warp of the CTA takes which role.  The recurrence warp is the critical path and its speed depends on what
// else issues from its SM sub-partition (measured: 353M cycles alone, 455M next to a second chain, 540M next to
// two), so roles are dealt by the sub-partition each warp actually sits on (%warpid & 3): the k-th CTA to arrive
// on an SM puts its chain on sub-partition k & 3, its model warps (the other heavy role) on the following ones,
// and the light helper warps wherever is left.
__device__ uint32_t g_sm_arrivals[1024];                      // CTAs started per SM, never reset: only k & 3 matters

template <int NS>
__device__ __forceinline__ int assign_role(int wslot, int lane) {
    constexpr int kWarps = 1 + 2 * NS;
    __shared__ uint32_t s_part[kWarps];
    __shared__ uint32_t s_ord;
    uint32_t smid, wid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (lane == 0) s_part[wslot] = wid & 3u;
    if (threadIdx.x == 0) s_ord = atomicAdd(&g_sm_arrivals[smid & 1023u], 1u);
    __syncthreads();
    const uint32_t c = s_ord & 3u;
    uint32_t free_mask = (1u << kWarps) - 1u;
    auto take = [&](uint32_t part) -> int {                  // a free warp on that sub-partition, else any free warp
        int pick = __ffs(free_mask) - 1;
#pragma unroll
        for (int w = kWarps - 1; w >= 0; --w)
            if (((free_mask >> w) & 1u) && s_part[w] == (part & 3u)) pick = w;
        free_mask &= ~(1u << pick);
        return pick;
    };
    int role = 0;
    if (take(c) == wslot) role = 0;
    constexpr uint32_t kModelAt[4] = {NS == 1 ? 2u : 1u, 2u, 3u, 2u};
#pragma unroll
    for (int k = 0; k < NS; ++k)
        if (take(c + kModelAt[k & 3]) == wslot) role = 1 + k;
#pragma unroll
    for (int k = 0; k < NS; ++k)
        if (take(c + 3u + k) == wslot) role = 1 + NS + k;
    return role;
}

// One CTA per SM (NS >= 3: the whole launch is one wave of <= 148 CTAs).  The warp scheduler of a sub-partition
// prefers the eligible warp with the highest warp id, so the recurrence warp -- the critical path -- is made the
// highest-numbered warp of the least populated sub-partition and shares it only with (light) helper warps; the
// model warps are spread evenly over the other three.
// kAlone: the CTA is launched with 4 ceil(2 NS / 3) warps, so that the model and helper warps fit the other three
// sub-partitions; the recurrence warp then has a sub-partition to itself (its spare warps leave at once: role -1).
// Roles: 0 recurrence, 1..NS model, NS+1..2NS helper.
__host__ __device__ constexpr int fused_warps(int ns, int mode) { return mode == 2 ? 4 * ((2 * ns + 2) / 3) : 1 + 2 * ns; }
template <int NS, bool kAlone>
__device__ __forceinline__ int assign_role_solo(int wslot, int lane) {
    constexpr int kWarps = fused_warps(NS, kAlone ? 2 : 1);
    constexpr int kRoles = 2 * NS;
    __shared__ uint32_t s_wid[kWarps];
    uint32_t wid;
    asm("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (lane == 0) s_wid[wslot] = wid;
    __syncthreads();
    int cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t p = s_wid[w] & 3u;
#pragma unroll
        for (int k = 0; k < 4; ++k) cnt[k] += (p == (uint32_t)k);
    }
    int c = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cnt[k] > 0 && (c < 0 || cnt[k] <= cnt[c])) c = k;
    int chain = -1;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if ((s_wid[w] & 3u) == (uint32_t)c && (chain < 0 || s_wid[w] > s_wid[chain])) chain = w;
    if (wslot == chain) return 0;
    const uint32_t mine = s_wid[wslot];
    const bool on_c = (mine & 3u) == (uint32_t)c;
    if (kAlone && on_c) return -1;
    // order of the warps off the chain's sub-partition: by rank inside their sub-partition, then by sub-partition
    auto key_of = [&](int w) -> uint32_t {
        uint32_t rank = 0;
#pragma unroll
        for (int v = 0; v < kWarps; ++v) rank += ((s_wid[v] & 3u) == (s_wid[w] & 3u) && s_wid[v] < s_wid[w]);
        return rank * 4u + (s_wid[w] & 3u);
    };
    int off_c = 0, before = 0;
    const uint32_t my_key = key_of(wslot);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        if (w == chain) continue;
        const bool w_on_c = (s_wid[w] & 3u) == (uint32_t)c;
        off_c += !w_on_c;
        if (w_on_c == on_c && key_of(w) < my_key) ++before;
    }
    // models first (off the recurrence's sub-partition), then the helpers
    const int idx = on_c ? off_c + before : before;
    return idx < kRoles ? 1 + idx : -1;
}

constexpr int kFifoBytes = (kFifoF + kFifoSpill) * 2;
constexpr int kFusedPerSlice = kFifoBytes + 2 * kRingSlots * 16 + 2 * kBlkF * 4 + 64;   // + control words
// kPixels (the CTA computes its records from the pixels, no record array in HBM): a ring of four chunks of 32 pixels x
// <= 4 planes of records per slice, and the quantiser tables once per CTA
constexpr int kRecChunk = 32 * 4;
constexpr int kRecRing = 4;
constexpr int kRecBytes = kRecRing * kRecChunk * 4;
constexpr int fused_smem_bytes(int ns, bool global_state, bool pixels) {
    return 1024 + ns * kFusedPerSlice + (global_state ? 0 : kRowBytesSmem) + (pixels ? ns * kRecBytes + (int)sizeof(QuantBytes) : 0);
}

// kWide: the launch is one wave of one CTA per SM, so the kernel may take the registers it wants (65 instead of 64 and no
// spill: 176.7 -> 174.7 ms on configs[3]); otherwise two CTAs must fit an SM (with more CTAs than SMs the second resident
// CTA is worth more: 4096 slices code at 22.1 GB/s against 16.0), which caps the registers at 64.
template <int NS, bool kGlobalState, int kSolo, bool kPixels, bool kWide = false>
__global__ void __launch_bounds__(32 * fused_warps(NS, kSolo), kWide ? 1 : 2) k_slice_coder_fused(const uint32_t* __restrict__ sym,
                                                                       const uint8_t* __restrict__ pixels, Geom g,
                                                                       uint8_t* __restrict__ scratch,
                                                                       uint32_t* __restrict__ slice_bytes,
                                                                       int* __restrict__ status,
                                                                       uint2* __restrict__ gstate, uint32_t n_slices) {
    static_assert(kGlobalState || NS == 1, "the state rows of one slice fill the shared memory of a CTA");
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int L = NS == 1 ? 32 : NS == 2 ? 16 : NS <= 4 ? 8 : 4;   // chain lanes per slice (NS L <= 32)
    static_assert(NS * L <= 32 && NS <= 8, "one chain warp serves all slices of the CTA");
    uint32_t* tab2 = reinterpret_cast<uint32_t*>(smem);
    auto slice_smem = [&](int q) { return smem + 1024 + q * kFusedPerSlice; };
    auto fifo_of = [&](int q) { return reinterpret_cast<uint16_t*>(slice_smem(q)); };
    auto ring_of = [&](int q, int buf) { return reinterpret_cast<uint4*>(slice_smem(q) + kFifoBytes) + buf * kRingSlots; };
    auto x_of = [&](int q, int buf) {
        return reinterpret_cast<uint32_t*>(slice_smem(q) + kFifoBytes + 2 * kRingSlots * 16) + buf * kBlkF;
    };
    // Flow control between the roles of a slice: shared-memory words with one writer each.
    //   ctl[0]          model -> helper         decisions produced so far (mod 2^32)
    //   ctl[2]          model -> helper         1 once the model warp has reached the end of the slice
    //   ctl[1]          helper -> model         FIFO positions below this one are expanded and may be overwritten
    //   ctl[4 + (b&3)]  helper -> chain/helper  decisions of block b of the slice (<= 0: none)
    // The model warp runs to the end of its slice at its own pace, as far ahead as the FIFO has room; only the chain
    // warp and the helpers meet at a barrier per block.  (With one CTA barrier for all roles every iteration took as
    // long as its slowest warp -- a model warp that met an L2 miss -- although the FIFO still held blocks of slack:
    // per-role counters showed every role under 75 % busy at seven slices per CTA.)
    auto ctl_of = [&](int q) {
        return reinterpret_cast<volatile uint32_t*>(slice_smem(q) + kFifoBytes + 2 * kRingSlots * 16 + 2 * kBlkF * 4);
    };
    auto role_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * (1 + NS)) : "memory"); };

    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    // 0 chain, 1..NS model, NS+1..2NS helper
    const int role = kSolo ? assign_role_solo<NS, kSolo == 2>(wslot, lane) : assign_role<NS>(wslot, lane);
    // slice (within the CTA) this warp / lane group serves; spare chain lanes shadow the last slice
    const int q = role == 0 ? min(lane / L, NS - 1) : (role - 1) % NS;
    const uint32_t sidx = blockIdx.x * NS + q;
    const bool live = sidx < n_slices;                        // surplus slices of the last CTA: nothing to do
    const uint64_t s = live ? sidx : n_slices - 1;
    const Slice sl = slice_of(g, s);
    const uint32_t* in = sym + sl.sym_off;
    const uint64_t n = live ? sl.n : 0;
    uint2* state = kGlobalState ? gstate + (size_t)s * kContexts
                                : reinterpret_cast<uint2*>(smem + 1024 + NS * kFusedPerSlice);
    uint8_t* const px_smem = smem + 1024 + NS * kFusedPerSlice + (kGlobalState ? 0 : kRowBytesSmem);
    QuantBytes* const quant = reinterpret_cast<QuantBytes*>(px_smem + NS * kRecBytes);
    uint16_t* fifo = fifo_of(q);
    volatile uint32_t* ctl = ctl_of(q);

    if (role == 0) fill_tab2(tab2, lane);
    if (kPixels)
        for (int i = threadIdx.x; i < (int)(sizeof(QuantBytes) / 16); i += blockDim.x)
            reinterpret_cast<uint4*>(quant)[i] = reinterpret_cast<const uint4*>(&c_quant_bytes_coder)[i];
    if (role > NS && lane < 16) ctl[lane] = 0;
    const bool spare = role < 0;
    if (!kGlobalState)
        for (int i = threadIdx.x; i < kRowBytesSmem / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(state)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (spare) return;

    const bool is_model = role >= 1 && role <= NS, is_helper = role > NS;
    // kPixels: the front end runs inside the CTA, in the model warp.  Records are made 32 pixels at a time (lane = pixel,
    // all its planes, any position of the slice: records_of_pixel) into a ring of chunks in shared memory, from which
    // the steps read them two steps ahead, as they read the record array otherwise.  (Measured on configs[3], seven
    // slices per CTA, against 176 + 4 ms with K1 and the record array: 193 ms this way; 210 ms with the helper warp
    // making the records before its barrier -- its global loads then delay the recurrence warp; 270 ms with one
    // producer warp for the seven slices of the CTA -- it cannot keep up.)
    const int C = g.C;                                        // 3 or 4 with kPixels (the launcher checks)
    const size_t pitch = (size_t)g.W * C;
    const uint8_t* const sl_px = kPixels ? pixels + ((size_t)sl.img * g.H + sl.y0) * pitch + (size_t)sl.x0 * C : nullptr;
    const uint32_t n_px = live ? (uint32_t)sl.w * (uint32_t)sl.h : 0u;
    uint32_t* const recbuf = reinterpret_cast<uint32_t*>(px_smem + q * kRecBytes);
    auto produce = [&](uint32_t chunk) {
        const int8_t* const q11 = quant->q11 + kQB;
        const int8_t* const q5 = quant->q5 + kQB;
        const uint32_t pi = chunk * 32u + lane;
        if (pi < n_px) {
            const uint32_t y = pi / (uint32_t)sl.w, x = pi - y * (uint32_t)sl.w;
            const uint8_t* p = sl_px + y * pitch + (size_t)x * C;
            uint32_t* dst = recbuf + (chunk & (kRecRing - 1)) * kRecChunk + lane * C;
            if (C == 3) {
                uint32_t r[3];
                records_of_pixel<3>(p, pitch, (int)x, (int)y, sl.w, q11, q5, r);
                dst[0] = r[0]; dst[1] = r[1]; dst[2] = r[2];
            } else {
                uint32_t r[4];
                records_of_pixel<4>(p, pitch, (int)x, (int)y, sl.w, q11, q5, r);
                *reinterpret_cast<uint4*>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
            }
            if (x + 32u < (uint32_t)sl.w) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32 * C));
        }
        __syncwarp();
    };

    // ---- model warp: software pipeline over 32-sample steps.  Records two steps ahead; the plan (votes, match)
    // and an L1 prefetch of the state rows one step ahead: a working set of 8 slices x ~4000 live rows does not
    // fit L1, and a row fetched from L2 in the middle of a step costs ~700 cycles of a warp with nothing else to do.
    if (is_model) {
#ifdef LLC_ROLE_TIMING
        long long t_spin = 0, t_all0 = clock64();
#endif
        uint64_t base = 0;
        uint32_t produced = 0, cons_seen = 0;                 // positions mod 2^32: the roles are never a ring apart
        uint32_t rec_seen = 0;                                // kPixels: chunks made so far
        // record of this lane's sample in step k (32 samples per step; a step never straddles two chunks)
        auto step_record = [&](uint64_t k) -> uint32_t {
            const uint64_t first = k * 32u;
            if (!kPixels) return first + lane < n ? __ldcs(in + first + lane) : 0u;   // read once: streaming, the state rows keep L2
            if (first >= n) return 0u;
            const uint32_t chunk = (uint32_t)(k / (uint32_t)C), sub = (uint32_t)(k - (uint64_t)chunk * C);
            while (rec_seen <= chunk) produce(rec_seen++);    // (a chunk two back is in registers by now: ring of four)
            return first + lane < n ? recbuf[(chunk & (kRecRing - 1)) * kRecChunk + sub * 32u + lane] : 0u;
        };
        uint32_t rec_cur = step_record(0);
        uint32_t rec_next = step_record(1);
        StepPlan plan_cur = model_plan(rec_cur, lane < n, lane);
        while (base < n) {
            if (produced - cons_seen + 32 * 19 > (uint32_t)kFifoF) {       // no room for a whole step: wait for the helper
#ifdef LLC_ROLE_TIMING
                const long long t_in = clock64();
#endif
                // (room comes a block at a time, every ~5000 cycles; polling faster only takes issue slots from the
                // recurrence warp: with 256 ns sleeps the seven model warps of a CTA spent more instructions polling
                // than the whole CTA spent working)
                while (produced - (cons_seen = ctl[1]) + 32 * 19 > (uint32_t)kFifoF) __nanosleep(LLC_SPIN_NS);
#ifdef LLC_ROLE_TIMING
                t_spin += clock64() - t_in;
#endif
            }
            const bool valid = base + lane < n, valid_next = base + 32 + lane < n;
            const uint32_t rec_after = step_record(base / 32 + 2);
            // this step's rows first: their latency (L1, often L2) runs under the votes and the match of the next step
            const uint2 row = model_row(rec_cur, valid, plan_cur, state, lane);
            if (kGlobalState && valid_next) asm volatile("prefetch.global.L1 [%0];" ::"l"(state + (rec_next >> 11)));
            const StepPlan plan_next = model_plan(rec_next, valid_next, lane);
            model_apply(rec_cur, valid, plan_cur, row, state, tab2, lane, FifoSink{smem_addr(fifo), produced});
            {
                const uint32_t a_cur = (uint32_t)abs(residual_of(rec_cur));
                fifo_unspill(smem_addr(fifo), produced, plan_cur.off, valid ? (a_cur ? 65u - 2u * __clz(a_cur) : 1u) : 0u, valid);
            }
            produced += plan_cur.total;
            rec_cur = rec_next; rec_next = rec_after; plan_cur = plan_next;
            base += 32;
            __syncwarp();                                     // every lane's entries before the count
            if (lane == 0) { __threadfence_block(); ctl[0] = produced; }
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); ctl[2] = 1u; }
#ifdef LLC_ROLE_TIMING
        if (lane == 0) {
            uint32_t smid, wid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            asm("mov.u32 %0, %%warpid;" : "=r"(wid));
            const long long tot = clock64() - t_all0;
            printf("cta %d sm %u role %d work %lld total %lld wid %u\n", blockIdx.x, smid, role, tot - t_spin, tot, wid);
        }
#endif
        return;                                               // the other roles meet at a named barrier that does not count this warp
    }
    // ---- helper warp
    uint8_t* const out0 = scratch + scratch_off(sl, s);
    uint8_t* const out_end = out0 + scratch_cap(sl);
    ByteTail t;
    t.low = 0; t.hp = kHpEmpty; t.outp = out0;               // llcomp.hpp:35
    bool overflow = false;
    uint32_t x_carry = 0xFF00u << 8;                         // pseudo-x whose successor range is the initial 0xFF00
    uint32_t nd_prev = 0, nd_cur = 0;                        // no-delta masks of blocks b-1 and b
    // FIFO entries of block blk -> chain operands (256 M, -255 M, A + bias); returns the lane's no-delta mask
    auto expand = [&](uint32_t blk) -> uint32_t {
        uint4* ring = ring_of(q, blk & 1) + lane;            // the lane's j-th decision goes to slot j * 32 + lane
        const uint4 e = *reinterpret_cast<const uint4*>(fifo + ((blk * kBlkF) & (kFifoF - 1)) + lane * kPerLane);
        const uint32_t w[4] = {e.x, e.y, e.z, e.w};
        uint32_t nd = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#if LLC_CHAIN_V == 2
            // entry -> (M, M/256 - M, A) as floats; 0x4B000000 | m is the float 2^23 + m
            const float m_lo = __uint_as_float(0x4B000000u | (w[k] & 0xFFu)) - 8388608.f;
            const float m_hi = __uint_as_float(0x4B000000u | ((w[k] >> 16) & 0xFFu)) - 8388608.f;
            const float a_lo = (w[k] & 0x8000u) ? 255.f : 0.f, a_hi = (w[k] & 0x80000000u) ? 255.f : 0.f;
            ring[(2 * k) * 32] = make_uint4(__float_as_uint(m_lo), __float_as_uint(m_lo * -0.99609375f), __float_as_uint(a_lo), 0u);
            ring[(2 * k + 1) * 32] = make_uint4(__float_as_uint(m_hi), __float_as_uint(m_hi * -0.99609375f), __float_as_uint(a_hi), 0u);
#else
            // A = 0xFF where the entry's flag (bit 15) is set: sign-replicating byte select, bias from the 2nd source
            const uint32_t m_lo = w[k] & 0xFFu, m_hi = prmt(w[k], 0x4442);
            ring[(2 * k) * 32] = make_uint4(m_lo << 8, m_lo * 0xFFFFFF01u, prmt2(w[k], 0x0000FF00u, 0x4549), 0u);
            ring[(2 * k + 1) * 32] = make_uint4(m_hi << 8, m_hi * 0xFFFFFF01u, prmt2(w[k], 0x0000FF00u, 0x454B), 0u);
#endif
            nd |= (w[k] & 0x80008000u) >> k;
        }
        return nd;
    };
    // ---- chain warp (per lane group)
#if LLC_CHAIN_V == 2
    uint32_t yc = __float_as_uint((float)(0xFF00u << 8));    // pseudo-x whose successor range is 0xFF00
    const uint32_t two8 = 0;
#else
    uint32_t yc = (0xFF00u << 8) + kChainBias;               // biased pseudo-x whose successor range is 0xFF00
    const uint32_t two8 = c_two8;
#endif

    // decisions of block blk once the model warp has produced all of it (256), or what is left of the slice (<= 0: none)
    auto wait_block = [&](uint32_t blk) -> int {
        const uint32_t first = blk * (uint32_t)kBlkF;
        for (;;) {
            const uint32_t done = ctl[2];
            const int ahead = (int)(ctl[0] - first);          // read after the flag: the last count precedes it
            if (ahead >= kBlkF) return kBlkF;
            if (done) return ahead;
            __nanosleep(128);
        }
    };
    if (is_helper) {
        const int c0 = wait_block(0);
        __threadfence_block();
        nd_cur = c0 > 0 ? expand(0) : 0u;
        __syncwarp();
        if (lane == 0) { ctl[1] = (uint32_t)kBlkF; ctl[4] = (uint32_t)c0; ctl[7] = 0u; }
    }
    role_sync();

#ifdef LLC_ROLE_TIMING
    long long t_work = 0, t_all0 = clock64();
#endif
    for (uint32_t b = 0;; ++b) {
#ifdef LLC_ROLE_TIMING
        const long long t_in = clock64();
#endif
        // decisions of blocks b and b-1, longest slice of the CTA; identical in every warp
        int max_cur = -1, max_prev = -1;
#pragma unroll
        for (int qq = 0; qq < NS; ++qq) {
            max_cur = max(max_cur, (int)ctl_of(qq)[4 + (b & 3)]);
            max_prev = max(max_prev, (int)ctl_of(qq)[4 + ((b + 3) & 3)]);
        }
        if (max_cur <= 0 && (b == 0 || max_prev <= 0)) break;            // nothing in block b nor in block b-1

        if (role == 0) {
            if (max_cur > 0) {
                // eight decisions per trip; slots e%8*32 + e/8.  Operands of the next four are in flight while four
                // run.  Shorter slices of the CTA (and the tail of a last block) compute garbage nobody reads.
                const uint4* rp = ring_of(q, b & 1);
                uint4* xo = reinterpret_cast<uint4*>(x_of(q, b & 1));
                const int trips = (max_cur + 7) / 8;
                uint4 a0 = rp[0], a1 = rp[32], a2 = rp[64], a3 = rp[96];
                // x of a group of four is stored one decision late: a store issued right behind the multiply-add that
                // produces its last word waits for it, and the in-order warp with it (profiles/microbench/chain_lat.cu:
                // 16.8 -> 15.8 cycles per decision; 14.9 without any store)
                uint4 xs, xw = make_uint4(0, 0, 0, 0);
                for (int v = 0; v < trips; ++v) {
                    const uint4 b0 = rp[128], b1 = rp[160], b2 = rp[192], b3 = rp[224];
                    xs.x = yc = chain_step(yc, a0, two8);
                    if (v) xo[-1] = xw;
                    xs.y = yc = chain_step(yc, a1, two8);
                    xs.z = yc = chain_step(yc, a2, two8);
                    xs.w = yc = chain_step(yc, a3, two8);
                    ++rp;
                    a0 = rp[0]; a1 = rp[32]; a2 = rp[64]; a3 = rp[96];
                    xw.x = yc = chain_step(yc, b0, two8);
                    xo[0] = xs;
                    xw.y = yc = chain_step(yc, b1, two8);
                    xw.z = yc = chain_step(yc, b2, two8);
                    xw.w = yc = chain_step(yc, b3, two8);
                    xo += 2;
                }
                xo[-1] = xw;
            }
        } else {
            const int cnt_prev = b > 0 ? (int)ctl[4 + ((b + 3) & 3)] : 0;
            if (cnt_prev > 0)
                byte_side_lanes(t, overflow, x_carry, x_of(q, (b - 1) & 1),
                                reinterpret_cast<uint32_t*>(ring_of(q, (b - 1) & 1)), nd_prev, (uint32_t)cnt_prev, lane,
                                out0, out_end);
            nd_prev = nd_cur;
            const int cnt_next = wait_block(b + 1);
            __threadfence_block();
            nd_cur = cnt_next > 0 ? expand(b + 1) : 0u;
            __syncwarp();
            if (lane == 0) { ctl[1] = (b + 2) * (uint32_t)kBlkF; ctl[4 + ((b + 1) & 3)] = (uint32_t)cnt_next; }
        }
#ifdef LLC_ROLE_TIMING
        t_work += clock64() - t_in;
#endif
        role_sync();
    }
#ifdef LLC_ROLE_TIMING                                       // per-role busy cycles
    if (lane == 0) {
        uint32_t smid, wid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        asm("mov.u32 %0, %%warpid;" : "=r"(wid));
        printf("cta %d sm %u role %d work %lld total %lld wid %u\n", blockIdx.x, smid, role, t_work, clock64() - t_all0, wid);
    }
#endif

    if (is_helper && live) {
        if (t.outp + (t.hp >> 9) + 8 > out_end) { overflow = true; t.outp = out0; t.hp &= 0x1FFu; }
        // finish(), llcomp.hpp:75-81: range = 0xFF both times, so each renorm_encoder call shifts exactly once
        t.low += 0xFFu;
        shift_low(t);
        shift_low(t);
        if (lane == 0) {
            slice_bytes[s] = overflow ? 0xFFFFFFFFu : (uint32_t)(t.outp - out0);
            if (overflow) atomicCAS(status, kDevOk, kDevOverflow);
        }
    }
}

static int sm_count() {                                      // of the current device (the devices of a box are alike)
    static const int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1)
            v = 148;
        return v;
    }();
    return n;
}

// The fused CTA with the state rows in shared memory takes 79 KB.  While every slice can have an SM of its own the rows
// stay there; beyond that they go behind L1 and one CTA per SM serves several slices (two such CTAs on an SM measured
// 162 ms per 1024^2 RGB slice against 150 ms for one CTA with two slices and the rows behind L1).
uint64_t fused_global_state_bytes(uint64_t n_slices) {
    if (switches().model_smem_state) return 0;
    // (a forced slices-per-CTA count implies the rows behind L1: the shared-memory form is one slice per CTA)
    return (n_slices > (uint64_t)sm_count() || switches().fused_ns) ? n_slices * (uint64_t)kStateBytes : 0;
}

// (Measured and dropped: asking for a smaller shared-memory carve-out so that the rows behind L1 get more of it.
// The default carve-out leaves them a 35% L1 hit rate, but the model warp hides that behind its look-ahead and the
// kernel time did not move, while co-resident launches of a pipelined batch got slower.)
template <int NS, bool kGlobalState, int kSolo, bool kPixels>
static cudaError_t launch_fused_from(const uint32_t* d_sym, const uint8_t* d_pixels, const Geom& g, uint8_t* d_scratch,
                                     uint32_t* d_slice_bytes, int* d_status, uint2* gs, unsigned n, cudaStream_t st) {
    constexpr int kSmem = fused_smem_bytes(NS, kGlobalState, kPixels);
    const cudaError_t configured = ensure_dynamic_smem<k_slice_coder_fused<NS, kGlobalState, kSolo, kPixels>>(kSmem);
    if (configured != cudaSuccess) return configured;
    // Shared memory / L1 split: a kernel that has opted in to large dynamic shared memory gets the largest carve-out by
    // default (233 KB, ~20 KB of L1); one CTA per SM needs kSmem, the rest is better spent on the state rows' L1.
    // (Only while the launch is one wave of one CTA per SM: with more CTAs than SMs the largest carve-out lets two of
    // them share an SM, which is worth more -- 4096 slices of 256^2: 22.1 GB/s against 15.9.)
    const bool one_wave = kSolo && kGlobalState && (n + NS - 1) / NS <= (unsigned)sm_count();
    const dim3 grid((n + NS - 1) / NS), block(32 * fused_warps(NS, kSolo));
    if (one_wave) {
        constexpr auto kernel = k_slice_coder_fused<NS, kGlobalState, kSolo, kPixels, (kSolo && kGlobalState)>;
        const cudaError_t wide_configured = ensure_dynamic_smem<kernel>(kSmem);
        if (wide_configured != cudaSuccess) return wide_configured;
        const int pct = switches().coder_max_carveout ? (int)cudaSharedmemCarveoutMaxShared
                                                      : std::min(100, ((kSmem + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024));
        (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        (void)cudaGetLastError();                            // a hint: never fails a launch
        kernel<<<grid, block, kSmem, st>>>(d_sym, d_pixels, g, d_scratch, d_slice_bytes, d_status, gs, n);
        return cudaGetLastError();
    }
    k_slice_coder_fused<NS, kGlobalState, kSolo, kPixels><<<grid, block, kSmem, st>>>(
        d_sym, d_pixels, g, d_scratch, d_slice_bytes, d_status, gs, n);
    return cudaGetLastError();
}

// The default forms exist twice: reading K1's record array (d_pixels == nullptr) or the pixels themselves.
struct FusedArgs {
    const uint32_t* d_sym; const uint8_t* d_pixels; const Geom& g; uint8_t* d_scratch; uint32_t* d_slice_bytes;
    int* d_status; uint2* gs; unsigned n; cudaStream_t st;
};
template <int NS, bool kGlobalState, int kSolo>
static cudaError_t launch_fused(const FusedArgs& a) {
    if (a.d_pixels)
        return launch_fused_from<NS, kGlobalState, kSolo, true>(nullptr, a.d_pixels, a.g, a.d_scratch, a.d_slice_bytes, a.d_status, a.gs, a.n, a.st);
    return launch_fused_from<NS, kGlobalState, kSolo, false>(a.d_sym, nullptr, a.g, a.d_scratch, a.d_slice_bytes, a.d_status, a.gs, a.n, a.st);
}
template <int NS, bool kGlobalState, int kSolo>
static cudaError_t launch_fused_records(const FusedArgs& a) {     // round-1 arrangements: record array only
    if (a.d_pixels) return cudaErrorInvalidValue;
    return launch_fused_from<NS, kGlobalState, kSolo, false>(a.d_sym, nullptr, a.g, a.d_scratch, a.d_slice_bytes, a.d_status, a.gs, a.n, a.st);
}


// true when the fused coder can compute its records from the pixels of this geometry itself (else: K1 + record array)
bool fused_coder_can_take_pixels(const Geom& g) {
    return (g.C == 3 || g.C == 4) && (switches().fused_ns == 0 || switches().fused_ns >= 10);
}
// Records from the pixels cost time (measured on 1024^2 RGB slices, K1 + record array against pixels: 138.4 / 140.3 ms
// at one slice per SM, 142.5 / 155.0 at two, 156.5 / 187.0 at four, 181.0 / 199.4 at seven), because the model warp that
// makes them is the second-busiest role of the CTA; they save the 4-byte-per-sample record array and its 8 bytes per
// sample of HBM traffic.  So the default takes K1's records while the array fits and the pixels when it does not.
bool fused_coder_takes_pixels(const Geom& g, bool record_array_fits) {
    if (!fused_coder_can_take_pixels(g) || switches().coder_records) return false;
    return switches().coder_pixels || !record_array_fits;
}

cudaError_t launch_slice_coder_fused(const uint32_t* d_sym, const uint8_t* d_pixels, const Geom& g, uint8_t* d_scratch,
                                     uint32_t* d_slice_bytes, int* d_status, uint8_t* d_gstate, cudaStream_t st,
                                     uint64_t n_concurrent) {
    const uint64_t ns = g.n_slices();
    if (ns == 0 || ns > 0x0FFFFFFFull || (d_sym == nullptr) == (d_pixels == nullptr)) return cudaErrorInvalidValue;
    if (d_pixels && !fused_coder_can_take_pixels(g)) return cudaErrorInvalidValue;
    const unsigned n = (unsigned)ns;
    if (!d_gstate) return launch_fused<1, false, 2>(FusedArgs{d_sym, d_pixels, g, d_scratch, d_slice_bytes, d_status, nullptr, n, st});
    // the caller decides where the rows live (fused_global_state_bytes); all states start at 0
    cudaError_t e = cudaMemsetAsync(d_gstate, 0, ns * (uint64_t)kStateBytes, st);
    if (e != cudaSuccess) return e;
    // One CTA per SM, as few slices per CTA as one wave allows (1024 slices on 148 SMs: 7); beyond 7 per SM the
    // launch takes several waves.  Up to five slices per CTA the recurrence warp gets a sub-partition of its own
    // (20+: measured 154 against 160 ms at four per SM); with six or seven the model warps would then be too crowded on
    // the other three (no gain at seven), so it shares its sub-partition with two helper warps (10+).
    // LLCOMP_FUSED_NS=1|2|4 selects the round-1 arrangement (several CTAs per SM, roles dealt by arrival order on the
    // SM) that the default is tested against; 11..17 and 22..27 force a solo form.
    const uint64_t n_resident = std::max<uint64_t>(n, n_concurrent);
    const int per_sm = (int)std::min<uint64_t>(7, std::max<uint64_t>(2, (n_resident + sm_count() - 1) / sm_count()));
    int per_cta = (per_sm <= 5 ? 20 : 10) + per_sm;
    if (switches().fused_ns) per_cta = switches().fused_ns;
    uint2* gs = reinterpret_cast<uint2*>(d_gstate);
    const FusedArgs a{d_sym, d_pixels, g, d_scratch, d_slice_bytes, d_status, gs, n, st};
    switch (per_cta) {
        case 1: return launch_fused_records<1, true, 0>(a);
        case 2: return launch_fused_records<2, true, 0>(a);
        case 4: return launch_fused_records<4, true, 0>(a);
        case 11: return launch_fused<1, true, 1>(a);
        case 12: return launch_fused<2, true, 1>(a);
        case 13: return launch_fused<3, true, 1>(a);
        case 14: return launch_fused<4, true, 1>(a);
        case 15: return launch_fused<5, true, 1>(a);
        case 16: return launch_fused<6, true, 1>(a);
        case 17: return launch_fused<7, true, 1>(a);
        case 22: return launch_fused<2, true, 2>(a);
        case 23: return launch_fused<3, true, 2>(a);
        case 24: return launch_fused<4, true, 2>(a);
        case 25: return launch_fused<5, true, 2>(a);
        case 27: return launch_fused<7, true, 2>(a);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------------
cudaError_t configure_slice_coder() {
    cudaError_t e = cudaFuncSetAttribute(k_model_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kModelSmem);
    return e;                                                // the fused kernels configure themselves at first launch
}

uint64_t model_global_state_bytes(uint64_t count) {
    // state in shared memory while every slice of the launch finds a slot (3 per SM), else behind L1
    return (count > 3 * 148 && !switches().model_smem_state) ? count * (uint64_t)kStateBytes : 0;
}

cudaError_t launch_model_pass(const uint32_t* d_sym, const Geom& g, uint64_t s0, uint64_t count, uint16_t* d_queue,
                              const uint64_t* d_qoff, uint8_t* d_gstate, cudaStream_t st) {
    if (count == 0 || count > 0x0FFFFFFFull) return cudaErrorInvalidValue;
    const unsigned n = (unsigned)count;
    if (model_global_state_bytes(count)) {
        cudaError_t e = cudaMemsetAsync(d_gstate, 0, count * (uint64_t)kStateBytes, st);   // all states start at 0
        if (e != cudaSuccess) return e;
        k_model_pass<true><<<n, 32, 256 * 4, st>>>(d_sym, g, s0, d_queue, d_qoff, reinterpret_cast<uint2*>(d_gstate));
        return cudaGetLastError();
    }
    k_model_pass<false><<<n, 32, kModelSmem, st>>>(d_sym, g, s0, d_queue, d_qoff, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_range_pass(const uint16_t* d_queue, const uint64_t* d_qoff, const unsigned long long* d_nbins,
                              const Geom& g, uint64_t s0, uint64_t count, uint8_t* d_scratch, uint32_t* d_slice_bytes,
                              int* d_status, cudaStream_t st) {
    if (count == 0 || count > 0x3FFFFFFFull) return cudaErrorInvalidValue;
    // One slice per warp while every warp can have a scheduler of its own (4 x 148 of them); beyond that slices
    // share warps: the chain is latency-bound, so a second slice in the same warp is nearly free.
    const unsigned n = (unsigned)count;
    k_range_pass_ws<<<n, 128, 0, st>>>(d_queue, d_qoff, d_nbins, g, s0, d_scratch, d_slice_bytes, d_status);
    return cudaGetLastError();
}

}  // namespace llc
