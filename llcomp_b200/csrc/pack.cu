// pack.cu -- K3: exclusive scan of slice byte counts; K4: compaction of the per-slice scratch
// payloads into one contiguous bitstream.  New in this build (the reference codes one image into one
// growing std::vector, llcomp.hpp:362-372); required because slices are coded concurrently.
#include "common.cuh"
#include "kernels.cuh"

namespace llc {

// ---- K3 ------------------------------------------------------------------------------------
// Single CTA, chunked: slice counts are small (<= a few thousand per image, 1024*S per batch).
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads) k_scan_sizes(const uint32_t* __restrict__ sizes, uint64_t n,
                                                             uint64_t* __restrict__ offsets, uint64_t capacity,
                                                             int* __restrict__ status) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += kScanThreads) {
        const uint64_t i = base + tid;
        const uint64_t v = i < n ? sizes[i] : 0;
        uint64_t x = v;                                      // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += y;
            }
            warp_sum[lane] = w;                              // inclusive over warps
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t excl = carry + (wid ? warp_sum[wid - 1] : 0) + x - v;
        if (i < n) offsets[i] = excl;
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        offsets[n] = carry_s;
        if (carry_s > capacity) atomicCAS(status, kDevOk, kDevOverflow);
    }
}

cudaError_t launch_scan(const uint32_t* d_slice_bytes, uint64_t n_slices, uint64_t* d_offsets, uint64_t capacity,
                        int* d_status, cudaStream_t st) {
    k_scan_sizes<<<1, kScanThreads, 0, st>>>(d_slice_bytes, n_slices, d_offsets, capacity, d_status);
    return cudaGetLastError();
}

// ---- K4 ------------------------------------------------------------------------------------
// grid = (n_slices, parts): CTA (s, p) copies a strided set of 4 KB pieces of slice s.  The destination is
// written with 16-byte stores on 16-byte boundaries; the (arbitrarily aligned) source is read as aligned
// 32-bit words and realigned with funnel shifts.
constexpr int kCompactThreads = 256;
constexpr uint32_t kPiece = kCompactThreads * 16;

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}

__global__ void __launch_bounds__(kCompactThreads) k_compact(const uint8_t* __restrict__ scratch, Geom g,
                                                             const uint64_t* __restrict__ offsets,
                                                             uint8_t* __restrict__ payload, uint64_t capacity) {
    const uint64_t s = blockIdx.x;
    const Slice sl = slice_of(g, s);
    const uint64_t o0 = offsets[s], o1 = offsets[s + 1];
    if (o1 > capacity) return;                               // overflow already flagged by the scan
    const uint64_t len = o1 - o0;
    if (len > scratch_cap(sl)) return;                       // slice overflowed its scratch (flagged by the coder)
    const uint8_t* src = scratch + scratch_off(sl, s);
    uint8_t* dst = payload + o0;

    // bytes up to the first 16-byte boundary of dst, then whole 16-byte units, then the tail
    const uint64_t head = min(len, (uint64_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    const uint64_t body = (len - head) / 16;
    if (blockIdx.y == 0) {
        if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
        const uint64_t tail0 = head + body * 16;
        if (tail0 + threadIdx.x < len) dst[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
    }
    const uint8_t* sb = src + head;
    uint4* db = reinterpret_cast<uint4*>(dst + head);
    for (uint64_t u = (uint64_t)blockIdx.y * kCompactThreads + threadIdx.x; u < body;
         u += (uint64_t)gridDim.y * kCompactThreads) {
        const uint8_t* q = sb + u * 16;
        uint4 v;
        v.x = load_u32_unaligned(q);
        v.y = load_u32_unaligned(q + 4);
        v.z = load_u32_unaligned(q + 8);
        v.w = load_u32_unaligned(q + 12);
        db[u] = v;
    }
}

cudaError_t launch_compact(const uint8_t* d_scratch, const Geom& g, const uint64_t* d_offsets, uint8_t* d_payload,
                           uint64_t capacity, cudaStream_t st) {
    const uint64_t ns = g.n_slices();
    if (ns > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    // enough CTAs per slice to cover ~2x raw in 4 KB pieces, bounded so the grid stays near 8 CTAs/SM
    const uint64_t max_bytes = 2ull * (uint64_t)min(g.tw, g.W) * min(g.th, g.H) * g.C + kScratchSlack;
    uint64_t parts = (max_bytes + kPiece - 1) / kPiece;
    const uint64_t want = (148ull * 8 + ns - 1) / ns;
    if (parts > want) parts = want;
    if (parts < 1) parts = 1;
    if (parts > 65535) parts = 65535;
    dim3 grid((unsigned)ns, (unsigned)parts);
    k_compact<<<grid, kCompactThreads, 0, st>>>(d_scratch, g, d_offsets, d_payload, capacity);
    return cudaGetLastError();
}

}  // namespace llc
