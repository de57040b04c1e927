/*
 * llcomp_oracle.h -- CPU restatement of the llcomp (revision 2) codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (llcomp_b200/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker or as the CPU baseline being reported.
 *
 * Parity pinned: every function here is differentially tested against the
 * unmodified reference header (oracle/_ref/libllcomp_ref.so, built from
 * /root/reference/llcomp.hpp by oracle/Makefile) and against the known-answer
 * vectors in tests/golden/ (see tests/test_oracle.py).
 *
 * Differences from the reference, both only where the reference is undefined
 * (defects D1/D2):
 *   D1  output buffer grows instead of being fixed at the raw size
 *       (llcomp.hpp:362 overflows when the stream is longer than the image);
 *   D2  the decoder skips the inverse colour transform when channels < 3,
 *       mirroring the encoder (llcomp.hpp:410-414 vs :532-540).
 */
#ifndef LLCOMP_ORACLE_H
#define LLCOMP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    LLO_OK = 0,
    LLO_BAD_MAGIC = 1,        /* llcomp.hpp:465-467 "Invalid magic number" */
    LLO_BAD_EXPONENT = 2,     /* llcomp.hpp:232-234 "Invalid exponent"     */
    LLO_NOMEM = 3,
    LLO_BAD_ARG = 4
};

/* One packed front-end record per sample: (hash << 11) | (diff & 0x7FF),
 * hash in [0,7925] after the sign fold, diff in [-510,510] (llcomp.hpp:424-436). */
#define LLO_SYM_PACK(hash, diff) (((uint32_t)(hash) << 11) | ((uint32_t)(diff) & 0x7FFu))

/* Table accessors so tests can compare with the reference's arrays. */
int llo_quant11(int x);                 /* llcomp.hpp:335-337 */
int llo_quant5(int x);                  /* llcomp.hpp:339-341 */
int llo_median(int a, int b, int c);    /* llcomp.hpp:343-356 */
int llo_next_state_mps(int s);          /* llcomp.hpp:252-259 */
int llo_next_state_lps(int s);          /* llcomp.hpp:261-268 */
int llo_state_probability(int s);       /* llcomp.hpp:270-281 */

/* Front end only (llcomp.hpp:390-436): pixels of one tile -> one record per
 * sample in raster, channel-interleaved order.  `pitch` is the byte distance
 * between rows of the enclosing image (== w*c for a whole image). */
int llo_frontend_tile(const uint8_t *px, size_t pitch, int w, int h, int c,
                      uint32_t *sym_out);

/* Bins of one residual (llcomp.hpp:166-206): writes up to 19 (ctx<<1|bit)
 * bytes, returns the count. */
int llo_binarize(int diff, uint8_t *ctxbit_out);

/* Headerless payload of one tile == compressImage(tile)[6:] (llcomp.hpp:380-450).
 * *out is malloc'd; caller frees with llo_free. */
int llo_encode_tile(const uint8_t *px, size_t pitch, int w, int h, int c,
                    uint8_t **out, size_t *out_len);

/* Range-code a pre-computed record array (what the GPU coder kernel does). */
int llo_encode_symbols(const uint32_t *sym, size_t n, uint8_t **out, size_t *out_len);

/* Inverse of llo_encode_tile (llcomp.hpp:475-545 with fix D2).  Bytes past
 * `len` read as zero (llcomp.hpp:476-477). */
int llo_decode_tile(const uint8_t *payload, size_t len, int w, int h, int c,
                    uint8_t *px_out, size_t pitch);

/* Whole-image stream, reference format: 79 C Wlo Whi Hlo Hhi payload
 * (llcomp.hpp:375-378, :463-470). */
int llo_compress(const uint8_t *px, int w, int h, int c, uint8_t **out, size_t *out_len);
int llo_decompress(const uint8_t *stream, size_t len, uint8_t **px_out,
                   int *w, int *h, int *c);

/* Counts binary decisions of a tile (for bins/s reporting). */
uint64_t llo_count_bins(const uint8_t *px, size_t pitch, int w, int h, int c);

void llo_free(void *p);

/* FNV-1a 64 over a byte range (SURVEY.md appendix B hashes). */
uint64_t llo_fnv1a64(const uint8_t *p, size_t n);

/* Synthetic generator G(W,H,C,n,seed) of SURVEY.md appendix C, with its own
 * MT19937 so that it matches std::mt19937.  noise < 0 selects the
 * high-entropy variant (pixel = rng() & 0xFF). */
void llo_generate(uint8_t *px, int w, int h, int c, int noise, uint32_t seed);

/* Multi-threaded batch helpers for the CPU baseline (each thread codes whole
 * independent images; the reference has no internal threading). Returns total
 * stream bytes, or 0 on error. */
uint64_t llo_compress_batch_mt(const uint8_t *px, int n_images, int w, int h, int c,
                               int n_threads);

/* Decode n_images streams (stream k = streams[offsets[k]..offsets[k+1])); returns samples decoded. */
uint64_t llo_decompress_batch_mt(const uint8_t *streams, const uint64_t *offsets, int n_images,
                                 int n_threads);

#ifdef __cplusplus
}
#endif
#endif
