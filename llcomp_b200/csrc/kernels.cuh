// kernels.cuh -- host-callable launchers of the device stages (internal to the library).
#pragma once
#include "common.cuh"

namespace llc {

// K1  pixels -> records (frontend.cu)
cudaError_t launch_frontend(const uint8_t* d_pixels, const Geom& g, uint32_t* d_sym, cudaStream_t st);

// K2  records -> per-slice scratch payloads + byte counts (coder.cu)
cudaError_t launch_slice_coder(const uint32_t* d_sym, const Geom& g, uint8_t* d_scratch,
                               uint32_t* d_slice_bytes, int* d_status, cudaStream_t st);
cudaError_t configure_slice_coder();

// K3  exclusive scan of slice byte counts; K4 compaction into one contiguous payload (pack.cu)
cudaError_t launch_scan(const uint32_t* d_slice_bytes, uint64_t n_slices, uint64_t* d_offsets,
                        uint64_t capacity, int* d_status, cudaStream_t st);
cudaError_t launch_compact(const uint8_t* d_scratch, const Geom& g, const uint64_t* d_offsets,
                           uint8_t* d_payload, uint64_t capacity, cudaStream_t st);

// K5  payloads -> pixels (decoder.cu)
cudaError_t launch_slice_decoder(const uint8_t* d_payload, const uint64_t* d_offsets, const Geom& g,
                                 uint8_t* d_pixels, int16_t* d_line_scratch, int* d_status,
                                 cudaStream_t st);
cudaError_t configure_slice_decoder();
// bytes of global line scratch the decoder needs for this geometry (0 when the rows fit in shared memory)
uint64_t decoder_line_scratch_bytes(const Geom& g);

}  // namespace llc
