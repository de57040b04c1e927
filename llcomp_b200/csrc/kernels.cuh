// kernels.cuh -- host-callable launchers of the device stages (internal to the library).
#pragma once
#include "common.cuh"

#include <atomic>

namespace llc {

// Opt a kernel into more than 48 KB of dynamic shared memory, once per device (the attribute is per device, and the
// multi-device entry points launch the same kernel on several of them from several host threads).
template <auto kernel>                                       // the kernel itself, not its type: kernels of one signature must not share the flag
inline cudaError_t ensure_dynamic_smem(int bytes) {
    static std::atomic<unsigned long long> done{0};         // one static per kernel: bit d = device d
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// Test switches (DESIGN.md "Test switches"): each selects the plain variant of one stage so that tests can check the
// default kernels against it; none selects a CPU path.  Read from the environment (LLCOMP_*) when a context is
// created and again by llcomp_b200_reload_switches(); the launchers never call getenv themselves.
struct Switches {
    bool frontend_simple = false;     // LLCOMP_FRONTEND_SIMPLE   one thread per pixel
    bool frontend_tiled = false;      // LLCOMP_FRONTEND_TILED    round-1 tiled front end instead of the streaming one
    bool decoder_simple = false;      // LLCOMP_DECODER_SIMPLE    plain chain, three row buffers
    bool decoder_v1 = false;          // LLCOMP_DECODER_V1        round-1 fast decoder (k_slice_decoder_fast) instead of the chain of decoder_chain.cuh
    bool coder_split = false;         // LLCOMP_CODER_SPLIT       model pass -> HBM queue -> range pass
    bool decoder_smem_state = false;  // LLCOMP_DECODER_SMEM_STATE
    bool model_smem_state = false;    // LLCOMP_MODEL_SMEM_STATE
    bool coder_records = false;       // LLCOMP_CODER_RECORDS     fused coder always reads K1's record array
    bool coder_pixels = false;        // LLCOMP_CODER_PIXELS      fused coder always computes its records from the pixels
    bool decoder_max_carveout = false; // LLCOMP_DECODER_MAX_CARVEOUT  chain decoder with the largest shared-memory carve-out (least L1), for measurements
    bool coder_max_carveout = false;  // LLCOMP_CODER_MAX_CARVEOUT  fused coder with the largest shared-memory carve-out (least L1), for measurements
    int frontend_variant = 0;         // LLCOMP_FRONTEND_VARIANT  measurement variants of the streaming front end (frontend_rows.cu)
    int decoder_variant = 0;          // LLCOMP_DECODER_VARIANT   measurement variants of the chain decoder (decoder_chain.cu)
    int fused_ns = 0;                 // LLCOMP_FUSED_NS          slices per coder CTA; 0 = automatic
    int groups = 0;                   // LLCOMP_GROUPS            image groups of the pipelined host-buffer calls; 0 = automatic
};
const Switches& switches();
void reload_switches();

// K1  pixels -> records, plus the exact number of binary decisions of every slice (frontend.cu).
//     d_slice_bins (n_slices counters) must be zero on entry; nullptr skips the counting.
cudaError_t launch_frontend(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym,
                            unsigned long long* d_slice_bins, cudaStream_t st);

// K1, streaming form (frontend_rows.cu): TMA row loads, rows carried in registers, 128-bit stores.
bool frontend_rows_applicable(const uint8_t* d_pixels, const Geom& g);
cudaError_t launch_frontend_rows(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, cudaStream_t st);
cudaError_t configure_frontend_rows();

// K2a records -> bin queue (model pass, state in shared memory); K2b bin queue -> per-slice scratch payloads
//     and byte counts (range pass).  Both work on slices [s0, s0+count); d_qoff[s] = first queue entry of slice s.
cudaError_t launch_model_pass(const uint32_t* d_sym, const Geom& g, uint64_t s0, uint64_t count, uint16_t* d_queue,
                              const uint64_t* d_qoff, uint8_t* d_gstate, cudaStream_t st);
// bytes of global state the model pass wants for a launch of `count` slices (0: state stays in shared memory)
uint64_t model_global_state_bytes(uint64_t count);
cudaError_t launch_range_pass(const uint16_t* d_queue, const uint64_t* d_qoff, const unsigned long long* d_nbins,
                              const Geom& g, uint64_t s0, uint64_t count, uint8_t* d_scratch, uint32_t* d_slice_bytes,
                              int* d_status, cudaStream_t st);
// K2 fused: records -> per-slice scratch payloads in one kernel (model / chain / helper warp roles), no bin
//     queue in HBM.  d_gstate != nullptr: the state rows live there (n_slices x 63,408 B, zeroed by the launch);
//     nullptr: in shared memory.  fused_global_state_bytes() says which one a stand-alone launch should use.
//     Exactly one of d_sym / d_pixels: with d_pixels (3- and 4-channel images, fused_coder_takes_pixels) the model
//     warps compute the records from the pixels themselves and no record array exists.
//     n_concurrent: slices of all fused-coder launches that run at the same time as this one (pipelined host-buffer
//     encode), this launch included; 0 = this launch has the device to itself.  It sizes the CTAs (slices per CTA).
cudaError_t launch_slice_coder_fused(const uint32_t* d_sym, const uint8_t* d_pixels, const Geom& g, uint8_t* d_scratch,
                                     uint32_t* d_slice_bytes, int* d_status, uint8_t* d_gstate, cudaStream_t st,
                                     uint64_t n_concurrent = 0);
// can: the geometry allows it (3 or 4 channels, a solo arrangement).  takes: ... and it is what a call should do --
// forced by a switch, or because the record array (4 bytes per sample) would not fit beside everything else.
bool fused_coder_can_take_pixels(const Geom& g);
bool fused_coder_takes_pixels(const Geom& g, bool record_array_fits = true);
uint64_t fused_global_state_bytes(uint64_t n_slices);
cudaError_t configure_slice_coder();

// K3  exclusive scan of slice byte counts; K4 compaction into one contiguous payload (pack.cu)
cudaError_t launch_scan(const uint32_t* d_slice_bytes, uint64_t n_slices, uint64_t* d_offsets,
                        uint64_t capacity, int* d_status, cudaStream_t st);
cudaError_t launch_compact(const uint8_t* d_scratch, const Geom& g, const uint64_t* d_offsets,
                           uint8_t* d_payload, uint64_t capacity, cudaStream_t st);

// K5  payloads -> pixels (decoder.cu)
//     shared_launch: the launch runs beside other decoder launches of the same call (pipelined host-buffer decode);
//     its state rows then always live in d_gstate.
cudaError_t launch_slice_decoder(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                 uint8_t* d_pixels, int16_t* d_line_scratch, uint8_t* d_gstate, int* d_status,
                                 cudaStream_t st, bool shared_launch = false);
// bytes of global state the decoder wants for this geometry (0 when the state stays in shared memory)
uint64_t decoder_global_state_bytes(const Geom& g, bool shared_launch = false);
cudaError_t configure_slice_decoder();
// bytes of global line scratch the decoder needs for this geometry (0 when the rows fit in shared memory)
uint64_t decoder_line_scratch_bytes(const Geom& g);
// K5, default form for 1..4 channels (decoder_chain.cu): d_gstate != nullptr -> state rows there (zeroed by the caller)
cudaError_t launch_slice_decoder_chain(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                       uint8_t* d_pixels, uint8_t* d_gstate, int* d_status, cudaStream_t st,
                                       bool shared_launch = false);
int chain_decoder_smem_bytes(const Geom& g, bool global_state);

}  // namespace llc
