// decoder_chain.cu -- K5, default form for 1..4 channels: one warp per slice, lane 0 runs the chain of
// decoder_chain.cuh (see there for the design), the whole warp stages the payload, prepares the row-above part of the
// context hashes before a row and does the inverse colour transform and the pixel stores after it.
// /root/reference/llcomp.hpp:475-545.
#include "common.cuh"
#include "decoder_chain.cuh"
#include "kernels.cuh"

namespace llc {

__constant__ ModelTables c_tables_chain = make_tables();

// kGlobalState: the slices' state rows live in global memory (pre-zeroed by the host) behind L1 and a CTA decodes one
// slice per warp (up to seven: blockDim.x / 32), the warps sharing the tables; else one slice per CTA with its rows in
// shared memory in front of the chain's buffers.  After the tables are filled the warps never meet again.
constexpr int kChainMaxWarps = 7;
template <int CT, bool kGlobalState, int kV>
__global__ void __launch_bounds__(32 * kChainMaxWarps, 1) k_slice_decoder_chain(const uint8_t* __restrict__ payload,
                                                                             const uint64_t* __restrict__ offsets, Geom g,
                                                                             uint8_t* __restrict__ pixels,
                                                                             int* __restrict__ status,
                                                                             uint2* __restrict__ gstate, uint32_t n_slices) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    // The window address of the shared memory goes through a volatile shared-memory word: a value the compiler has
    // loaded cannot be rematerialised (it otherwise re-reads SR_CgaCtaId and rebuilds the base inside the sample loop,
    // in front of every table request).
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = (uint32_t)__cvta_generic_to_shared(smem);
    __syncthreads();
    const uint32_t base = *reinterpret_cast<volatile uint32_t*>(&s_base);
    const int row_elems = min(g.tw, g.W) * CT;                            // one layout for every slice of the launch
    const dchain::Layout L = dchain::make_layout(row_elems, base + (kGlobalState ? 0u : (uint32_t)kStateBytes), warp, n_warps);
    dchain::Smem m;
    dchain::fill_tables(m, L, c_tables_chain.entry, threadIdx.x, blockDim.x);
    if (!kGlobalState) {
        uint4* rows = reinterpret_cast<uint4*>(smem);                     // rows first (16-byte aligned), then the chain's buffers
        for (int i = threadIdx.x; i < kStateBytes / 16; i += blockDim.x) rows[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint64_t s = (uint64_t)blockIdx.x * n_warps + warp;
    if (s >= n_slices) return;
    const Slice sl = slice_of(g, s);
    dchain::StateMem<kGlobalState> st;
    st.g = reinterpret_cast<uint64_t>(gstate + (size_t)s * kContexts);
    st.s = base;
    const size_t pitch = (size_t)g.W * CT;
    uint8_t* dst = pixels + (size_t)sl.img * g.H * pitch + (size_t)sl.y0 * pitch + (size_t)sl.x0 * CT;
    const uint8_t* src = payload + offsets[s];
    const uint32_t len = (uint32_t)(offsets[s + 1] - offsets[s]);
    const bool ok = dchain::decode_slice_rows<CT, kGlobalState, kV>(m, L, st, c_tables_chain.entry, src, len, sl.w, sl.h, dst,
                                                                     pitch, lane, 32, [] { __syncwarp(); }, true);
    if (!ok && lane == 0) atomicCAS(status, kDevOk, kDevBadExponent);
}

// slices per CTA: as few as one wave of one CTA per SM allows, at most seven; one when the rows live in shared memory
// or when the warps' buffers would not fit together
static int chain_warps_per_cta(const Geom& g, bool global_state, bool shared_launch) {
    if (!global_state) return 1;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    // (a launch that runs beside another one of the same call -- pipelined host-buffer decode, two groups -- leaves it half of the SMs)
    const uint64_t resident = g.n_slices() * (shared_launch ? 2 : 1);
    int w = (int)std::min<uint64_t>(kChainMaxWarps, std::max<uint64_t>(1, (resident + sms - 1) / sms));
    while (w > 1 && dchain::layout_bytes(std::min(g.tw, g.W) * g.C, w) > 226u * 1024u) --w;
    return w;
}

int chain_decoder_smem_bytes(const Geom& g, bool global_state) {
    return (int)dchain::layout_bytes(std::min(g.tw, g.W) * g.C, 1) + (global_state ? 0 : kStateBytes);   // one slice per CTA: the least a launch needs
}

template <int CT, bool kGlobalState, int kV = 0>
static cudaError_t launch_chain(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g, uint8_t* d_pixels,
                                int* d_status, uint2* gs, unsigned n, cudaStream_t st, bool shared_launch = false) {
    const int warps = chain_warps_per_cta(g, kGlobalState, shared_launch);
    const int smem = (int)dchain::layout_bytes(std::min(g.tw, g.W) * g.C, warps) + (kGlobalState ? 0 : kStateBytes);
    const cudaError_t configured = ensure_dynamic_smem<k_slice_decoder_chain<CT, kGlobalState, kV>>(226 * 1024);
    if (configured != cudaSuccess) return configured;
    // Shared memory / L1 split.  Left to itself the driver gives a kernel that has opted in to large dynamic shared memory
    // the largest carve-out (233 KB: ncu launch__shared_mem_config_size), i.e. ~20 KB of L1 -- and the chain's speed
    // follows the L1 hit rate of its state rows.  Ask for what the resident CTAs need and no more.
    const unsigned ctas = (n + warps - 1) / warps;
    if (!switches().decoder_max_carveout) {
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        const int by_smem = std::max(1, (228 * 1024) / (smem + 1024 + 16));
        const int per_sm = std::min<int>(by_smem, std::max<int>(1, (int)((ctas * (shared_launch ? 2u : 1u) + sms - 1) / sms)));
        const int pct = std::min(100, (per_sm * (smem + 1024 + 16) * 100 + 228 * 1024 - 1) / (228 * 1024));
        (void)cudaFuncSetAttribute(k_slice_decoder_chain<CT, kGlobalState, kV>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        (void)cudaGetLastError();                            // a hint: never fails a launch
    } else {
        (void)cudaFuncSetAttribute(k_slice_decoder_chain<CT, kGlobalState, kV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        (void)cudaGetLastError();
    }
    k_slice_decoder_chain<CT, kGlobalState, kV><<<ctas, 32 * warps, smem, st>>>(d_payload, d_offsets, g, d_pixels, d_status, gs, n);
    return cudaGetLastError();
}

cudaError_t launch_slice_decoder_chain(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                       uint8_t* d_pixels, uint8_t* d_gstate, int* d_status, cudaStream_t st,
                                       bool shared_launch) {
    const unsigned n = (unsigned)g.n_slices();
    uint2* gs = reinterpret_cast<uint2*>(d_gstate);
    // measurement variants (LLCOMP_DECODER_VARIANT=1, 2, 4, three channels, rows behind L1): see Chain's kV
    if (switches().decoder_variant && g.C == 3 && d_gstate) {
        switch (switches().decoder_variant) {
            case 1: return launch_chain<3, true, 1>(d_payload, d_offsets, g, d_pixels, d_status, gs, n, st);
            case 2: return launch_chain<3, true, 2>(d_payload, d_offsets, g, d_pixels, d_status, gs, n, st);
            case 4: return launch_chain<3, true, 4>(d_payload, d_offsets, g, d_pixels, d_status, gs, n, st);
        }
    }
#define LLC_CASE(CT)                                                                                                    \
    case CT:                                                                                                            \
        return d_gstate ? launch_chain<CT, true>(d_payload, d_offsets, g, d_pixels, d_status, gs, n, st, shared_launch) \
                        : launch_chain<CT, false>(d_payload, d_offsets, g, d_pixels, d_status, nullptr, n, st);
    switch (g.C) {
        LLC_CASE(1) LLC_CASE(2) LLC_CASE(3) LLC_CASE(4)
    }
#undef LLC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace llc
