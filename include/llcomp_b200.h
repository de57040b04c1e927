/*
 * llcomp_b200.h -- C ABI of the B200-native llcomp encode/decode hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types,
 * no exceptions.  It is what a binding for the reference's two entry points
 *
 *     std::vector<uint8_t> llcomp::compressImage(rgb, width, height, channels)   llcomp.hpp:358
 *     llcomp::RawImage     llcomp::decompressImage(data)                          llcomp.hpp:461
 *
 * (called from llcompc.cpp:33 and llcompd.cpp:26) links against.  The C++ header
 * llcomp_b200/host/llcomp.hpp re-creates those two signatures on top of this
 * ABI; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Bitstreams
 *   grid 1x1 and W,H <= 65535:  the reference's revision-2 stream, byte-identical
 *                               79 | C | W u16le | H u16le | payload       (llcomp.hpp:375-378)
 *   otherwise (sliced):         B2 | 01 | C | 00 | W u32 | H u32 | tile_w u32 | tile_h u32 |
 *                               n_slices u32 | n_slices x u32 payload bytes | payloads
 *                               where payload k == reference compressImage(tile k)[6:],
 *                               tiles in row-major order (new container; the reference has none).
 *
 * Every function needs the CUDA device; there is no CPU fallback.
 */
#ifndef LLCOMP_B200_H
#define LLCOMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLCOMP_B200_ABI_VERSION 1

#define LLCOMP_MAGIC_REV2 0x79   /* llcomp.hpp:19-20: 0x77 + revision */
#define LLCOMP_MAGIC_SLICED 0xB2

typedef enum {
    LLCOMP_OK = 0,
    LLCOMP_ERR_BAD_MAGIC = 1,     /* host wrapper throws std::runtime_error("Invalid magic number"), llcomp.hpp:466 */
    LLCOMP_ERR_BAD_EXPONENT = 2,  /* host wrapper throws std::runtime_error("Invalid exponent"),     llcomp.hpp:233 */
    LLCOMP_ERR_BAD_ARG = 3,
    LLCOMP_ERR_NOMEM = 4,
    LLCOMP_ERR_OVERFLOW = 5,      /* a slice payload outgrew its scratch (the reference overflows its heap here: llcomp.hpp:362) */
    LLCOMP_ERR_CUDA = 6,
    LLCOMP_ERR_TRUNCATED = 7      /* sliced container shorter than its own header/table */
} llcomp_status;

/* One batch of equally sized images cut into a regular tile grid.
 * tile_w/tile_h == 0 means "whole image" (one slice per image).  Edge tiles are smaller. */
typedef struct {
    int32_t width, height, channels;
    int32_t tile_w, tile_h;
    int32_t n_images;
} llcomp_geometry;

typedef struct llcomp_ctx llcomp_ctx;   /* per-device context: scratch buffers, tables, last error.
                                          * Calls on one context are serialised by a lock inside it (the reference's
                                          * functions are re-entrant, llcomp.hpp:358/:461; this keeps a shared context
                                          * safe).  The device-resident calls are asynchronous: two of them in flight
                                          * on different streams need two contexts. */
typedef struct llcomp_multi llcomp_multi; /* several devices of one box, one context and one host thread per device */

int llcomp_b200_abi_version(void);
const char *llcomp_b200_status_string(int status);
const char *llcomp_b200_last_error(const llcomp_ctx *ctx);   /* CUDA error text of the last LLCOMP_ERR_CUDA */

int llcomp_b200_ctx_create(int device, llcomp_ctx **ctx);
void llcomp_b200_ctx_destroy(llcomp_ctx *ctx);

/* ---- geometry helpers (host only, no device work) ------------------------ */
uint64_t llcomp_b200_slice_count(const llcomp_geometry *g);
uint64_t llcomp_b200_sample_count(const llcomp_geometry *g);      /* W*H*C*n_images */
uint64_t llcomp_b200_payload_capacity(const llcomp_geometry *g);  /* bytes d_payload must hold for encode_device */
uint64_t llcomp_b200_stream_bound(const llcomp_geometry *g);      /* payload capacity + all headers */

/* ---- host-buffer entry points: replace llcomp.hpp:358 / :461 ------------- */
/* *stream is malloc'd by the library; release with llcomp_b200_free. */
int llcomp_b200_encode(llcomp_ctx *ctx, const uint8_t *pixels, int width, int height, int channels,
                       int tile_w, int tile_h, uint8_t **stream, size_t *stream_len);
int llcomp_b200_decode(llcomp_ctx *ctx, const uint8_t *stream, size_t stream_len,
                       uint8_t **pixels, int *width, int *height, int *channels);
/* Header only (llcomp.hpp:463-470); tile_w/tile_h = image size for a reference stream. */
int llcomp_b200_peek(const uint8_t *stream, size_t stream_len, int *width, int *height, int *channels,
                     int *tile_w, int *tile_h);

/* Batch of n_images equally sized images, pixels contiguous image after image.  Writes one complete
 * stream per image, back to back, into the caller's buffer `out` (out_cap bytes; llcomp_b200_stream_bound
 * is always enough; a pinned buffer avoids a staging copy).  offsets has n_images+1 entries. */
int llcomp_b200_encode_batch(llcomp_ctx *ctx, const uint8_t *pixels, const llcomp_geometry *g,
                             uint8_t *out, uint64_t out_cap, uint64_t *offsets);
/* Inverse; every stream must describe the same geometry; pixels_out is caller-allocated. */
int llcomp_b200_decode_batch(llcomp_ctx *ctx, const uint8_t *streams, const uint64_t *offsets, int n_images,
                             uint8_t *pixels_out, uint64_t pixels_cap, llcomp_geometry *g_out);
void llcomp_b200_free(void *p);
/* Page-locked host memory for the buffers of the batch calls (pixels in, streams out and back): copies from and to it
 * run asynchronously at full PCIe speed and overlap with the coding of other image groups; with pageable memory the
 * driver stages every copy.  NULL when it cannot be had (the caller falls back to ordinary memory). */
void *llcomp_b200_host_alloc(size_t bytes);
void llcomp_b200_host_free(void *p);

/* ---- several GPUs of one box ---------------- */
/* Slices share nothing, so the work is dealt to the devices in contiguous blocks with no exchange between them:
 * a batch by images, a single image by bands of tile rows.  Output bytes and offsets are exactly those of the
 * single-device calls above (one stream per image; a banded image is one container with the slice table of the
 * whole image).  devices[] may name a device more than once. */
int llcomp_b200_multi_create(const int *devices, int n_devices, llcomp_multi **multi);
void llcomp_b200_multi_destroy(llcomp_multi *multi);
int llcomp_b200_multi_device_count(const llcomp_multi *multi);
llcomp_ctx *llcomp_b200_multi_ctx(llcomp_multi *multi, int k);   /* context of the k-th device (last_error, launch counts) */
int llcomp_b200_multi_encode_batch(llcomp_multi *multi, const uint8_t *pixels, const llcomp_geometry *g,
                                   uint8_t *out, uint64_t out_cap, uint64_t *offsets);
int llcomp_b200_multi_decode_batch(llcomp_multi *multi, const uint8_t *streams, const uint64_t *offsets, int n_images,
                                   uint8_t *pixels_out, uint64_t pixels_cap, llcomp_geometry *g_out);

/* ---- device-resident entry points (inputs and outputs stay in HBM) ------- */
/* Asynchronous on `cuda_stream` (a cudaStream_t; NULL = default stream).  Errors raised on the
 * device (overflow, invalid exponent) are collected by llcomp_b200_finish.
 *
 * encode: d_pixels [n_images][H][W][C] u8  ->  d_payload: slice payloads back to back,
 *         d_offsets[n_slices+1] u64 exclusive scan of payload bytes (slice k = [off[k], off[k+1])). */
int llcomp_b200_encode_device(llcomp_ctx *ctx, const uint8_t *d_pixels, const llcomp_geometry *g,
                              uint8_t *d_payload, uint64_t payload_capacity, uint64_t *d_offsets,
                              void *cuda_stream);
int llcomp_b200_decode_device(llcomp_ctx *ctx, const uint8_t *d_payload, const uint64_t *d_offsets,
                              const llcomp_geometry *g, uint8_t *d_pixels, void *cuda_stream);
/* Synchronises `cuda_stream` and returns the first device-side error since the last finish. */
int llcomp_b200_finish(llcomp_ctx *ctx, void *cuda_stream);

/* ---- single stages, for parity tests and profiling ----------------------- */
/* Front end alone (llcomp.hpp:390-436): one u32 record (hash<<11 | diff&0x7FF) per sample,
 * slice-major, raster, channel-interleaved.  d_symbols holds sample_count records. */
int llcomp_b200_frontend_device(llcomp_ctx *ctx, const uint8_t *d_pixels, const llcomp_geometry *g,
                                uint32_t *d_symbols, void *cuda_stream);
/* Number of kernels launched through this context so far (bench.py reports it as gpu_launches). */
uint64_t llcomp_b200_launch_count(const llcomp_ctx *ctx);
/* Device time of the kernels of the LAST encode_device / decode_device call, in milliseconds, measured
 * with CUDA events on the caller's stream when profiling was switched on; names in llcomp_b200_stage_name. */
#define LLCOMP_B200_N_STAGES 6
void llcomp_b200_set_profiling(llcomp_ctx *ctx, int on);
int llcomp_b200_stage_times(llcomp_ctx *ctx, float *ms_out /* LLCOMP_B200_N_STAGES */);
const char *llcomp_b200_stage_name(int stage);
/* Binary decisions coded by the last encode call on this context (bins/s reporting). */
uint64_t llcomp_b200_last_bin_count(const llcomp_ctx *ctx);
/* Only for the two-kernel coder kept behind LLCOMP_CODER_SPLIT=1 (the default fused coder has no queue):
 * bytes of HBM its bin queue may take; default 40 % of the device.  Tests shrink it to force launch groups. */
void llcomp_b200_set_queue_budget(llcomp_ctx *ctx, uint64_t bytes);
/* The front end's records (4 bytes per sample) go through an array in HBM while it fits `bytes` (default: a third of the
 * device) and the free memory; beyond that the coder computes them from the pixels inside its CTAs (3- and 4-channel
 * images; a few percent slower, no array).  0 = always from the pixels.  The second call says what the last encode did. */
void llcomp_b200_set_record_budget(llcomp_ctx *ctx, uint64_t bytes);
int llcomp_b200_last_encode_from_pixels(const llcomp_ctx *ctx);
/* Test switches (LLCOMP_FRONTEND_SIMPLE, LLCOMP_FRONTEND_TILED, LLCOMP_DECODER_SIMPLE, LLCOMP_CODER_SPLIT,
 * LLCOMP_DECODER_SMEM_STATE, LLCOMP_MODEL_SMEM_STATE, LLCOMP_FUSED_NS) select the plain GPU variant of a stage.  They are
 * read from the environment when a context is created; call this after changing them in a live process. */
void llcomp_b200_reload_switches(void);
/* Model table entry of state s: P(bit=1)*256 | next_if_mps<<8 | next_if_lps<<16 (llcomp.hpp:252-281). */
uint32_t llcomp_b200_debug_table(int s);

#ifdef __cplusplus
}
#endif
#endif /* LLCOMP_B200_H */
