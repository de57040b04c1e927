// api.cu -- the extern "C" layer of include/llcomp_b200.h: per-device context, stream containers,
// host-buffer and device-resident entry points.  Replaces the call sites llcompc.cpp:33
// (compressImage) and llcompd.cpp:26 (decompressImage) of the reference; header layout follows
// llcomp.hpp:375-378 / :463-470.
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "../../include/llcomp_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace llc;

constexpr int llcomp_ctx_groups = 4;   // streams a host-buffer call may pipeline its image groups over

namespace llc {
static Switches g_switches;
const Switches& switches() { return g_switches; }
void reload_switches() {
    auto on = [](const char* name) { const char* v = getenv(name); return v && *v && std::strcmp(v, "0") != 0; };
    Switches s;
    s.frontend_simple = on("LLCOMP_FRONTEND_SIMPLE");
    s.frontend_tiled = on("LLCOMP_FRONTEND_TILED");
    s.decoder_simple = on("LLCOMP_DECODER_SIMPLE");
    s.decoder_v1 = on("LLCOMP_DECODER_V1");
    s.coder_max_carveout = on("LLCOMP_CODER_MAX_CARVEOUT");
    s.decoder_max_carveout = on("LLCOMP_DECODER_MAX_CARVEOUT");
    s.coder_split = on("LLCOMP_CODER_SPLIT");
    s.decoder_smem_state = on("LLCOMP_DECODER_SMEM_STATE");
    s.model_smem_state = on("LLCOMP_MODEL_SMEM_STATE");
    s.coder_records = on("LLCOMP_CODER_RECORDS");
    s.coder_pixels = on("LLCOMP_CODER_PIXELS");
    if (const char* v = getenv("LLCOMP_FUSED_NS")) s.fused_ns = atoi(v);
    if (const char* v = getenv("LLCOMP_DECODER_VARIANT")) s.decoder_variant = atoi(v);
    if (const char* v = getenv("LLCOMP_FRONTEND_VARIANT")) s.frontend_variant = atoi(v);
    if (const char* v = getenv("LLCOMP_GROUPS")) s.groups = std::max(0, std::min(atoi(v), (int)llcomp_ctx_groups));
    g_switches = s;
}
}  // namespace llc

namespace {

enum Stage { kStFrontend = 0, kStModel, kStRange, kStScan, kStCompact, kStDecoder };
const char* const kStageNames[LLCOMP_B200_N_STAGES] = {"frontend", "model_pass", "slice_coder", "scan", "compact",
                                                       "slice_decoder"};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct llcomp_ctx {
    std::recursive_mutex mu;                // one call at a time per context (the work buffers below are shared)
    int device = 0;
    cudaStream_t stream = nullptr;          // used by the host-buffer entry points
    static constexpr int kGroups = llcomp_ctx_groups;   // host-buffer calls: image groups pipelined over this many streams
    cudaStream_t group_stream[kGroups] = {};
    DevBuf<uint32_t> sym;                   // K1 -> K2a records
    DevBuf<unsigned long long> slice_bins;  // K1: exact number of binary decisions per slice
    DevBuf<uint64_t> qoff;                  // first bin-queue entry of every slice (within its launch group)
    DevBuf<uint16_t> queue;                 // K2a -> K2b bin queue
    PinnedBuf<unsigned long long> h_bins;
    PinnedBuf<uint64_t> h_qoff;
    bool from_pixels = false;               // the current fused encode computes its records from the pixels (no record array)
    uint64_t last_bins = 0;                 // decisions coded by the last encode call
    uint64_t queue_budget = 0;              // bytes the bin queue may take; slices are processed in groups that fit
    uint64_t record_budget = 0;             // bytes K1's record array may take; beyond, the coder works from the pixels
    DevBuf<uint8_t> scratch;                // K2b per-slice payloads before compaction
    DevBuf<uint32_t> slice_bytes;
    DevBuf<int16_t> lines;                  // K5 row scratch when a tile row does not fit in smem
    DevBuf<uint8_t> gstate;                 // per-slice state rows when they live in global memory (behind L1)
    DevBuf<uint8_t> pixels, payload;        // staging of the host-buffer entry points
    DevBuf<uint64_t> offsets;
    int* d_status = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    bool profiling = false;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;                // one per kernel launch of the last device call (profiling only)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

namespace {

int fail_cuda(llcomp_ctx* c, cudaError_t e, const char* where) {
    if (c) c->last_error = std::string(where) + ": " + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? LLCOMP_ERR_NOMEM : LLCOMP_ERR_CUDA;
}
#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);    \
    } while (0)

bool make_geom(const llcomp_geometry* in, Geom* g) {
    if (!in) return false;
    if (in->width < 1 || in->height < 1 || in->channels < 1 || in->channels > 255 || in->n_images < 1) return false;
    if (in->tile_w < 0 || in->tile_h < 0) return false;
    g->W = in->width; g->H = in->height; g->C = in->channels; g->n_images = in->n_images;
    g->tw = (in->tile_w == 0 || in->tile_w > in->width) ? in->width : in->tile_w;
    g->th = (in->tile_h == 0 || in->tile_h > in->height) ? in->height : in->tile_h;
    g->tiles_x = (g->W + g->tw - 1) / g->tw;
    g->tiles_y = (g->H + g->th - 1) / g->th;
    if ((uint64_t)g->tw * g->th * g->C >= (1ull << 30)) return false;   // per-slice byte counts are u32
    if ((uint64_t)g->tiles_x * (uint64_t)g->tiles_y >= (1ull << 31)) return false;   // slices_per_image() is u32
    if (g->n_slices() >= (1ull << 31)) return false;
    return true;
}

uint64_t payload_capacity(const Geom& g) { return 2 * g.n_samples() + kScratchSlack * g.n_slices(); }

void put_u32(uint8_t* p, uint32_t v) { p[0] = v & 0xFF; p[1] = (v >> 8) & 0xFF; p[2] = (v >> 16) & 0xFF; p[3] = v >> 24; }
uint32_t get_u32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

// A single-slice image that fits u16 dimensions is written in the reference's own format.
bool is_reference_layout(const Geom& g) { return g.slices_per_image() == 1 && g.W <= 0xFFFF && g.H <= 0xFFFF; }
size_t header_bytes(const Geom& g) { return is_reference_layout(g) ? 6 : 24 + 4 * (size_t)g.slices_per_image(); }

void write_header(const Geom& g, const uint64_t* off /* offsets of this image's slices, spi+1 */, uint8_t* h) {
    if (is_reference_layout(g)) {                        // llcomp.hpp:375-378
        h[0] = LLCOMP_MAGIC_REV2; h[1] = (uint8_t)g.C;
        h[2] = g.W & 0xFF; h[3] = (g.W >> 8) & 0xFF;
        h[4] = g.H & 0xFF; h[5] = (g.H >> 8) & 0xFF;
        return;
    }
    const uint32_t spi = g.slices_per_image();
    h[0] = LLCOMP_MAGIC_SLICED; h[1] = 1; h[2] = (uint8_t)g.C; h[3] = 0;
    put_u32(h + 4, g.W); put_u32(h + 8, g.H); put_u32(h + 12, g.tw); put_u32(h + 16, g.th); put_u32(h + 20, spi);
    for (uint32_t k = 0; k < spi; ++k) put_u32(h + 24 + 4 * k, (uint32_t)(off[k + 1] - off[k]));
}

// Parses either header.  lens receives the per-slice payload byte counts (reference stream: one entry).
int parse_header(const uint8_t* s, size_t n, llcomp_geometry* g, size_t* hdr, std::vector<uint32_t>* lens) {
    if (!s || n < 1) return LLCOMP_ERR_BAD_ARG;
    if (s[0] == LLCOMP_MAGIC_REV2) {                     // llcomp.hpp:463-470
        if (n < 6) return LLCOMP_ERR_TRUNCATED;
        g->channels = s[1];
        g->width = s[2] | (s[3] << 8);
        g->height = s[4] | (s[5] << 8);
        g->tile_w = g->width; g->tile_h = g->height; g->n_images = 1;
        *hdr = 6;
        if (lens) lens->assign(1, (uint32_t)std::min<size_t>(n - 6, 0xFFFFFFFFu));
        return LLCOMP_OK;
    }
    if (s[0] != LLCOMP_MAGIC_SLICED) return LLCOMP_ERR_BAD_MAGIC;
    if (n < 24) return LLCOMP_ERR_TRUNCATED;
    if (s[1] != 1) return LLCOMP_ERR_BAD_MAGIC;
    g->channels = s[2];
    g->width = (int32_t)get_u32(s + 4); g->height = (int32_t)get_u32(s + 8);
    g->tile_w = (int32_t)get_u32(s + 12); g->tile_h = (int32_t)get_u32(s + 16);
    g->n_images = 1;
    const uint32_t spi = get_u32(s + 20);
    Geom gg;
    if (g->width < 1 || g->height < 1 || g->tile_w < 1 || g->tile_h < 1 || !make_geom(g, &gg)) return LLCOMP_ERR_BAD_ARG;
    if (gg.slices_per_image() != spi || gg.tw != g->tile_w || gg.th != g->tile_h) return LLCOMP_ERR_BAD_ARG;
    if (n < 24 + 4 * (size_t)spi) return LLCOMP_ERR_TRUNCATED;
    *hdr = 24 + 4 * (size_t)spi;
    if (lens) {
        lens->resize(spi);
        for (uint32_t k = 0; k < spi; ++k) (*lens)[k] = get_u32(s + 24 + 4 * k);
    }
    return LLCOMP_OK;
}

cudaEvent_t take_event(llcomp_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}
void begin_call(llcomp_ctx* ctx) { ctx->spans.clear(); ctx->ev_used = 0; }
// Brackets one kernel launch with events on the launching stream (profiling only) and counts it.
struct StageScope {
    llcomp_ctx* ctx; cudaStream_t st; cudaEvent_t b = nullptr;
    StageScope(llcomp_ctx* c, cudaStream_t s, int stage) : ctx(c), st(s) {
        ctx->launches++;
        if (!ctx->profiling) return;
        cudaEvent_t a = take_event(ctx);
        b = take_event(ctx);
        cudaEventRecord(a, st);
        ctx->spans.push_back({stage, a, b});
    }
    ~StageScope() { if (b) cudaEventRecord(b, st); }
};

// The record array K1 writes for the coder (4 bytes per sample), unless the fused coder is to compute its records from
// the pixels: forced by a switch, or because the array does not fit (it is then not worth a third of the device either).
cudaError_t reserve_records(llcomp_ctx* ctx, const Geom& g) {
    ctx->from_pixels = false;
    const size_t need = g.n_samples();
    if (switches().coder_split) return ctx->sym.reserve(need);
    if (fused_coder_takes_pixels(g, true)) { ctx->from_pixels = true; return cudaSuccess; }   // forced by a switch
    const bool may_fall_back = fused_coder_takes_pixels(g, false);
    if (may_fall_back && need * 4 > ctx->record_budget) { ctx->sym.release(); ctx->from_pixels = true; return cudaSuccess; }
    if (need <= ctx->sym.cap) return cudaSuccess;
    if (may_fall_back) {
        size_t free_b = 0, total_b = 0;
        const bool roomy = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need * 4 + (1ull << 30) <= free_b + ctx->sym.cap * 4;
        if (!roomy) { ctx->sym.release(); ctx->from_pixels = true; return cudaSuccess; }
    }
    const cudaError_t e = ctx->sym.reserve(need);
    if (e == cudaErrorMemoryAllocation && may_fall_back) {
        (void)cudaGetLastError();
        ctx->from_pixels = true;
        return cudaSuccess;
    }
    return e;
}

// Front end + fused coder + scan + compaction of the images [first_image, first_image + g.n_images) of a batch
// whose workspace (records, scratch, byte counts, state rows) was reserved for the whole batch: every buffer is
// linear in the image index, so a group of images simply works on its own stretch of each.
int encode_fused_on(llcomp_ctx* ctx, const uint8_t* d_pixels, const Geom& g, uint64_t first_image, bool global_state,
                    uint8_t* d_payload, uint64_t capacity, uint64_t* d_offsets, cudaStream_t st, uint64_t n_concurrent = 0) {
    const uint64_t spi = g.slices_per_image(), first_slice = first_image * spi;
    const bool from_pixels = ctx->from_pixels;               // no record array then (the caller has not reserved one)
    uint32_t* sym = from_pixels ? nullptr : ctx->sym.p + first_image * g.image_samples();
    uint8_t* scratch = ctx->scratch.p + first_image * (2 * g.image_samples() + kScratchSlack * spi);
    uint32_t* slice_bytes = ctx->slice_bytes.p + first_slice;
    uint8_t* gstate = global_state ? ctx->gstate.p + first_slice * (uint64_t)kStateBytes : nullptr;
    if (!from_pixels) {
        StageScope sc(ctx, st, kStFrontend);
        CK(launch_frontend(d_pixels, g, sym, nullptr, st));
    }
    {
        StageScope sc(ctx, st, kStRange);
        CK(launch_slice_coder_fused(sym, from_pixels ? d_pixels : nullptr, g, scratch, slice_bytes, ctx->d_status, gstate, st,
                                    n_concurrent));
    }
    {
        StageScope sc(ctx, st, kStScan);
        CK(launch_scan(slice_bytes, g.n_slices(), d_offsets, capacity, ctx->d_status, st));
    }
    {
        StageScope sc(ctx, st, kStCompact);
        CK(launch_compact(scratch, g, d_offsets, d_payload, capacity, st));
    }
    return LLCOMP_OK;
}


// ---- host-buffer decode ----------------------------------------------------------------------------
// Headers of a batch of streams, parsed on the host: the slice table of the whole batch (offsets inside one contiguous
// device payload) and, per image, which host bytes exist (a truncated stream reads as zero beyond its end,
// llcomp.hpp:476-477).
struct ParsedBatch {
    llcomp_geometry gi{};
    std::vector<uint64_t> off;          // n_slices + 1
    struct Piece { uint64_t src, dst, have, want; };
    std::vector<Piece> pieces;          // one per image
};

int parse_batch(const uint8_t* streams, const uint64_t* offsets, int n_images, ParsedBatch& pb) {
    std::vector<uint32_t> lens;
    uint64_t total = 0;
    pb.off.assign(1, 0);
    pb.pieces.clear();
    for (int k = 0; k < n_images; ++k) {
        const uint8_t* s = streams + offsets[k];
        const size_t n = (size_t)(offsets[k + 1] - offsets[k]);
        llcomp_geometry gk; size_t hdr;
        const int rc = parse_header(s, n, &gk, &hdr, &lens);
        if (rc) return rc;
        if (k == 0) pb.gi = gk;
        else if (gk.width != pb.gi.width || gk.height != pb.gi.height || gk.channels != pb.gi.channels ||
                 gk.tile_w != pb.gi.tile_w || gk.tile_h != pb.gi.tile_h) return LLCOMP_ERR_BAD_ARG;
        uint64_t want = 0;
        for (uint32_t L : lens) { want += L; pb.off.push_back(total + want); }
        pb.pieces.push_back({offsets[k] + hdr, total, std::min<uint64_t>(want, n - hdr), want});
        total += want;
    }
    pb.gi.n_images = n_images;
    return LLCOMP_OK;
}

// Decodes images [first, first + count) of a parsed batch into pixels_out (their pixels, back to back).  Image groups
// are pipelined over a few streams: payload upload, slice decoder and pixel download of different groups overlap.
int decode_parsed(llcomp_ctx* ctx, const uint8_t* streams, const ParsedBatch& pb, const Geom& g_all, int first, int count,
                  uint8_t* pixels_out) {
    CK(cudaSetDevice(ctx->device));
    Geom g = g_all;
    g.n_images = count;
    const uint64_t spi = g.slices_per_image(), ns = g.n_slices(), img_bytes = g.image_samples();
    const uint64_t s0 = (uint64_t)first * spi;
    const uint64_t pay0 = pb.off[s0], pay_bytes = pb.off[s0 + ns] - pay0;
    // (two groups: the copies are a twentieth of the decode time, and four launches side by side decode slower than two)
    const int want_groups = switches().groups ? switches().groups : 2;
    const int n_groups = count >= 2 * want_groups ? want_groups : 1;
    const bool shared = n_groups > 1;
    CK(ctx->payload.reserve(pay_bytes + 16));
    CK(ctx->offsets.reserve(ns + n_groups));
    CK(ctx->pixels.reserve(g.n_samples()));
    const uint64_t lb = decoder_line_scratch_bytes(g);
    if (lb) CK(ctx->lines.reserve(lb / 2));
    const uint64_t gb = decoder_global_state_bytes(g, shared);
    if (gb) CK(ctx->gstate.reserve(gb));
    // offsets relative to this call's payload buffer, in pinned memory (uploads from pageable memory would serialise)
    CK(ctx->h_qoff.reserve(ns + n_groups));
    begin_call(ctx);
    int img = 0;
    for (int k = 0; k < n_groups; ++k) {
        const int cnt = count / n_groups + (k < count % n_groups ? 1 : 0);
        cudaStream_t st = n_groups == 1 ? ctx->stream : ctx->group_stream[k];
        const uint64_t sl0 = (uint64_t)img * spi, nsl = (uint64_t)cnt * spi;
        uint64_t* h_off = ctx->h_qoff.p + sl0 + k;
        for (uint64_t i = 0; i <= nsl; ++i) h_off[i] = pb.off[s0 + sl0 + i] - pay0;
        uint64_t* d_off = ctx->offsets.p + sl0 + k;
        CK(cudaMemcpyAsync(d_off, h_off, (nsl + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        for (int i = 0; i < cnt; ++i) {
            const ParsedBatch::Piece& pc = pb.pieces[first + img + i];
            uint8_t* dst = ctx->payload.p + (pc.dst - pay0);
            if (pc.have) CK(cudaMemcpyAsync(dst, streams + pc.src, pc.have, cudaMemcpyHostToDevice, st));
            if (pc.have < pc.want) CK(cudaMemsetAsync(dst + pc.have, 0, pc.want - pc.have, st));
        }
        Geom gg = g;
        gg.n_images = cnt;
        uint8_t* d_px = ctx->pixels.p + (uint64_t)img * img_bytes;
        int16_t* lines = lb ? ctx->lines.p + sl0 * (3ull * std::min(g.tw, g.W) * g.C) : nullptr;
        uint8_t* gstate = gb ? ctx->gstate.p + sl0 * (uint64_t)kStateBytes : nullptr;
        {
            StageScope sc(ctx, st, kStDecoder);
            CK(launch_slice_decoder(ctx->payload.p, d_off, gg, d_px, lines, gstate, ctx->d_status, st, shared));
        }
        CK(cudaMemcpyAsync(pixels_out + (uint64_t)img * img_bytes, d_px, (uint64_t)cnt * img_bytes, cudaMemcpyDeviceToHost, st));
        img += cnt;
    }
    cudaError_t first_err = cudaSuccess;
    for (int k = 0; k < n_groups; ++k) {
        const cudaError_t e = cudaStreamSynchronize(n_groups == 1 ? ctx->stream : ctx->group_stream[k]);
        if (e != cudaSuccess && first_err == cudaSuccess) first_err = e;
    }
    if (first_err != cudaSuccess) return fail_cuda(ctx, first_err, "decode_parsed: stream synchronize");
    int dev = 0;
    CK(cudaMemcpy(&dev, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (dev != 0) { CK(cudaMemset(ctx->d_status, 0, sizeof(int))); return dev; }
    return LLCOMP_OK;
}

}  // namespace

extern "C" {

int llcomp_b200_abi_version(void) { return LLCOMP_B200_ABI_VERSION; }
void llcomp_b200_reload_switches(void) { reload_switches(); }

const char* llcomp_b200_status_string(int s) {
    switch (s) {
        case LLCOMP_OK: return "ok";
        case LLCOMP_ERR_BAD_MAGIC: return "Invalid magic number";     // llcomp.hpp:466
        case LLCOMP_ERR_BAD_EXPONENT: return "Invalid exponent";      // llcomp.hpp:233
        case LLCOMP_ERR_BAD_ARG: return "bad argument";
        case LLCOMP_ERR_NOMEM: return "out of memory";
        case LLCOMP_ERR_OVERFLOW: return "slice payload exceeds its buffer";
        case LLCOMP_ERR_CUDA: return "CUDA error";
        case LLCOMP_ERR_TRUNCATED: return "truncated container header";
        default: return "unknown status";
    }
}

const char* llcomp_b200_last_error(const llcomp_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
const char* llcomp_b200_stage_name(int i) { return (i >= 0 && i < LLCOMP_B200_N_STAGES) ? kStageNames[i] : ""; }

int llcomp_b200_ctx_create(int device, llcomp_ctx** out) {
    if (!out) return LLCOMP_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        (void)cudaGetLastError();
        return LLCOMP_ERR_CUDA;                                      // no device: there is no CPU fallback
    }
    llcomp_ctx* ctx = new (std::nothrow) llcomp_ctx;
    if (!ctx) return LLCOMP_ERR_NOMEM;
    reload_switches();
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int k = 0; k < llcomp_ctx::kGroups && e == cudaSuccess; ++k)
        e = cudaStreamCreateWithFlags(&ctx->group_stream[k], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_status), sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(int));
    if (e == cudaSuccess) {
        size_t free_b = 0, total_b = 0;
        e = cudaMemGetInfo(&free_b, &total_b);
        ctx->queue_budget = (uint64_t)total_b * 2 / 5;       // 40 % of HBM (72 GB on B200)
        ctx->record_budget = (uint64_t)total_b / 3;          // a third of HBM (60 GB: 15 G samples per call)
    }
    if (e == cudaSuccess) e = configure_frontend_rows();
    if (e == cudaSuccess) e = configure_slice_coder();
    if (e == cudaSuccess) e = configure_slice_decoder();
    if (e != cudaSuccess) {
        fprintf(stderr, "llcomp_b200: context creation failed: %s\n", cudaGetErrorString(e));
        llcomp_b200_ctx_destroy(ctx);
        return LLCOMP_ERR_CUDA;
    }
    *out = ctx;
    return LLCOMP_OK;
}

void llcomp_b200_ctx_destroy(llcomp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->sym.release(); ctx->scratch.release(); ctx->slice_bytes.release(); ctx->lines.release();
    ctx->pixels.release(); ctx->payload.release(); ctx->offsets.release();
    ctx->gstate.release();
    ctx->slice_bins.release(); ctx->qoff.release(); ctx->queue.release(); ctx->h_bins.release(); ctx->h_qoff.release();
    if (ctx->d_status) cudaFree(ctx->d_status);
    for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    for (auto& gs : ctx->group_stream) if (gs) cudaStreamDestroy(gs);
    delete ctx;
}

uint64_t llcomp_b200_slice_count(const llcomp_geometry* g) { Geom gg; return make_geom(g, &gg) ? gg.n_slices() : 0; }
uint64_t llcomp_b200_sample_count(const llcomp_geometry* g) { Geom gg; return make_geom(g, &gg) ? gg.n_samples() : 0; }
uint64_t llcomp_b200_payload_capacity(const llcomp_geometry* g) { Geom gg; return make_geom(g, &gg) ? payload_capacity(gg) : 0; }
uint64_t llcomp_b200_stream_bound(const llcomp_geometry* g) {
    Geom gg;
    if (!make_geom(g, &gg)) return 0;
    return payload_capacity(gg) + (uint64_t)header_bytes(gg) * gg.n_images;
}
uint64_t llcomp_b200_launch_count(const llcomp_ctx* ctx) { return ctx ? ctx->launches : 0; }
void llcomp_b200_set_profiling(llcomp_ctx* ctx, int on) { if (ctx) ctx->profiling = on != 0; }

int llcomp_b200_stage_times(llcomp_ctx* ctx, float* ms) {
    if (!ctx || !ms) return LLCOMP_ERR_BAD_ARG;
    for (int i = 0; i < LLCOMP_B200_N_STAGES; ++i) ms[i] = 0.f;
    for (const auto& sp : ctx->spans) {
        float t = 0.f;
        CK(cudaEventSynchronize(sp.b));
        CK(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.stage] += t;
    }
    return LLCOMP_OK;
}

uint64_t llcomp_b200_last_bin_count(const llcomp_ctx* ctx) { return ctx ? ctx->last_bins : 0; }
void llcomp_b200_set_queue_budget(llcomp_ctx* ctx, uint64_t bytes) { if (ctx && bytes) ctx->queue_budget = bytes; }
void llcomp_b200_set_record_budget(llcomp_ctx* ctx, uint64_t bytes) { if (ctx) ctx->record_budget = bytes; }
int llcomp_b200_last_encode_from_pixels(const llcomp_ctx* ctx) { return ctx && ctx->from_pixels; }

// ---- device-resident path --------------------------------------------------------------------
int llcomp_b200_frontend_device(llcomp_ctx* ctx, const uint8_t* d_pixels, const llcomp_geometry* gi, uint32_t* d_sym,
                                void* cuda_stream) {
    Geom g;
    if (!ctx || !d_pixels || !d_sym || !make_geom(gi, &g)) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    begin_call(ctx);
    {
        StageScope sc(ctx, st, kStFrontend);
        CK(launch_frontend(d_pixels, g, d_sym, nullptr, st));
    }
    return LLCOMP_OK;
}

int llcomp_b200_encode_device(llcomp_ctx* ctx, const uint8_t* d_pixels, const llcomp_geometry* gi, uint8_t* d_payload,
                              uint64_t capacity, uint64_t* d_offsets, void* cuda_stream) {
    Geom g;
    if (!ctx || !d_pixels || !d_payload || !d_offsets || !make_geom(gi, &g)) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint64_t ns = g.n_slices();
    CK(reserve_records(ctx, g));
    CK(ctx->scratch.reserve(payload_capacity(g) + 64));
    CK(ctx->slice_bytes.reserve(ns));
    CK(ctx->slice_bins.reserve(ns));
    CK(ctx->qoff.reserve(ns));
    CK(ctx->h_bins.reserve(ns));
    CK(ctx->h_qoff.reserve(ns));
    begin_call(ctx);

    if (!switches().coder_split) {
        // Default: front end, then ONE fused coder kernel (model + range chain + bytes per slice); fully asynchronous.
        const uint64_t gsb = fused_global_state_bytes(ns);
        if (gsb) CK(ctx->gstate.reserve(gsb));
        const int rc = encode_fused_on(ctx, d_pixels, g, 0, gsb != 0, d_payload, capacity, d_offsets, st);
        ctx->last_bins = 0;
        return rc;
    } else {
    // K1: records + exact decision count of every slice
        CK(cudaMemsetAsync(ctx->slice_bins.p, 0, ns * sizeof(unsigned long long), st));
        {
            StageScope sc(ctx, st, kStFrontend);
            CK(launch_frontend(d_pixels, g, ctx->sym.p, ctx->slice_bins.p, st));
        }
        // The one host round trip of the encoder: the counts size the bin queue and split the slices into
        // launch groups that fit the queue budget (one group unless the batch is huge or very noisy).
        CK(cudaMemcpyAsync(ctx->h_bins.p, ctx->slice_bins.p, ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint64_t budget_entries = ctx->queue_budget / 2;
        struct Group { uint64_t s0, count, entries; };
        std::vector<Group> groups;
        uint64_t run = 0, start = 0;
        ctx->last_bins = 0;
        for (uint64_t k = 0; k < ns; ++k) {
            ctx->last_bins += ctx->h_bins.p[k];
            const uint64_t need = (ctx->h_bins.p[k] + kQueuePad + 7) & ~7ull;
            if (need > budget_entries) { ctx->last_error = "bin queue budget too small for one slice"; return LLCOMP_ERR_NOMEM; }
            if (run + need > budget_entries) { groups.push_back({start, k - start, run}); start = k; run = 0; }
            ctx->h_qoff.p[k] = run;
            run += need;
        }
        groups.push_back({start, ns - start, run});
        uint64_t max_entries = 0;
        for (const Group& gr : groups) max_entries = std::max(max_entries, gr.entries);
        CK(ctx->queue.reserve(max_entries + 64));
        CK(cudaMemcpyAsync(ctx->qoff.p, ctx->h_qoff.p, ns * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        uint64_t gstate_bytes = 0;
        for (const Group& gr : groups) gstate_bytes = std::max(gstate_bytes, model_global_state_bytes(gr.count));
        if (gstate_bytes) CK(ctx->gstate.reserve(gstate_bytes));

        for (const Group& gr : groups) {
            {
                StageScope sc(ctx, st, kStModel);
                CK(launch_model_pass(ctx->sym.p, g, gr.s0, gr.count, ctx->queue.p, ctx->qoff.p, ctx->gstate.p, st));
            }
            {
                StageScope sc(ctx, st, kStRange);
                CK(launch_range_pass(ctx->queue.p, ctx->qoff.p, ctx->slice_bins.p, g, gr.s0, gr.count, ctx->scratch.p,
                                     ctx->slice_bytes.p, ctx->d_status, st));
            }
        }
    }
    {
        StageScope sc(ctx, st, kStScan);
        CK(launch_scan(ctx->slice_bytes.p, ns, d_offsets, capacity, ctx->d_status, st));
    }
    {
        StageScope sc(ctx, st, kStCompact);
        CK(launch_compact(ctx->scratch.p, g, d_offsets, d_payload, capacity, st));
    }
    return LLCOMP_OK;
}

int llcomp_b200_decode_device(llcomp_ctx* ctx, const uint8_t* d_payload, const uint64_t* d_offsets,
                              const llcomp_geometry* gi, uint8_t* d_pixels, void* cuda_stream) {
    Geom g;
    if (!ctx || !d_payload || !d_offsets || !d_pixels || !make_geom(gi, &g)) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint64_t lb = decoder_line_scratch_bytes(g);
    if (lb) CK(ctx->lines.reserve(lb / 2));
    const uint64_t gb = decoder_global_state_bytes(g);
    if (gb) CK(ctx->gstate.reserve(gb));
    begin_call(ctx);
    {
        StageScope sc(ctx, st, kStDecoder);
        CK(launch_slice_decoder(d_payload, d_offsets, g, d_pixels, ctx->lines.p, ctx->gstate.p, ctx->d_status, st));
    }
    return LLCOMP_OK;
}

int llcomp_b200_finish(llcomp_ctx* ctx, void* cuda_stream) {
    if (!ctx) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    int st = 0;
    CK(cudaMemcpy(&st, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st != 0) CK(cudaMemset(ctx->d_status, 0, sizeof(int)));
    return st;
}

// ---- host-buffer path ------------------------------------------------------------------------
int llcomp_b200_encode_batch(llcomp_ctx* ctx, const uint8_t* pixels, const llcomp_geometry* gi, uint8_t* out,
                             uint64_t out_cap, uint64_t* offsets) {
    Geom g;
    if (!ctx || !pixels || !out || !offsets || !make_geom(gi, &g)) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const uint64_t ns = g.n_slices(), cap = payload_capacity(g);
    const uint32_t spi = g.slices_per_image();
    const size_t hb = header_bytes(g);
    CK(ctx->pixels.reserve(g.n_samples()));
    CK(ctx->payload.reserve(cap));

    // Image groups pipelined over a few streams: the upload of group k+1 and the download of group k-1 run under
    // the coding of group k, and -- the coder being latency-bound per slice -- the groups' kernels overlap each
    // other on the GPU.  The state rows then have to live behind L1 (a shared-memory slot per slice would
    // serialise the groups).  The split coder (LLCOMP_CODER_SPLIT) and small batches take the single-call path.
    // (the groups' coder launches run side by side: each is told how many slices are resident in all, so that together
    // they fill the device once instead of each sizing its CTAs as if it were alone)
    const int want_groups = switches().groups ? switches().groups : llcomp_ctx::kGroups;
    const int n_groups = (switches().coder_split || g.n_images < 2 * want_groups) ? 1 : want_groups;
    if (n_groups == 1) {
        CK(ctx->offsets.reserve(ns + 1));
        cudaStream_t st = ctx->stream;
        CK(cudaMemcpyAsync(ctx->pixels.p, pixels, g.n_samples(), cudaMemcpyHostToDevice, st));
        int rc = llcomp_b200_encode_device(ctx, ctx->pixels.p, gi, ctx->payload.p, cap, ctx->offsets.p, st);
        if (rc) return rc;
        std::vector<uint64_t> off(ns + 1);
        CK(cudaMemcpyAsync(off.data(), ctx->offsets.p, (ns + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        rc = llcomp_b200_finish(ctx, st);
        if (rc) return rc;
        uint64_t pos = 0;
        for (int k = 0; k < g.n_images; ++k) {
            const uint64_t p0 = off[(uint64_t)k * spi], p1 = off[(uint64_t)(k + 1) * spi];
            offsets[k] = pos;
            if (pos + hb + (p1 - p0) > out_cap) return LLCOMP_ERR_OVERFLOW;
            write_header(g, &off[(uint64_t)k * spi], out + pos);
            CK(cudaMemcpyAsync(out + pos + hb, ctx->payload.p + p0, p1 - p0, cudaMemcpyDeviceToHost, st));
            pos += hb + (p1 - p0);
        }
        offsets[g.n_images] = pos;
        CK(cudaStreamSynchronize(st));
        return LLCOMP_OK;
    }

    CK(reserve_records(ctx, g));
    CK(ctx->scratch.reserve(cap + 64));
    CK(ctx->slice_bytes.reserve(ns));
    CK(ctx->gstate.reserve(ns * (uint64_t)kStateBytes));
    CK(ctx->offsets.reserve(ns + n_groups));
    begin_call(ctx);
    struct Part { int first, count; uint64_t slice0; };
    std::vector<Part> parts;
    for (int k = 0, first = 0; k < n_groups; ++k) {
        const int count = g.n_images / n_groups + (k < g.n_images % n_groups ? 1 : 0);
        parts.push_back({first, count, (uint64_t)first * spi});
        first += count;
    }
    const uint64_t img_bytes = g.image_samples(), img_cap = 2 * img_bytes + kScratchSlack * spi;
    // offsets land in pinned memory: a copy to pageable memory would block the host until the group is coded
    CK(ctx->h_qoff.reserve(ns + n_groups));
    std::vector<uint64_t*> off(n_groups);
    for (int k = 0; k < n_groups; ++k) {
        const Part& pt = parts[k];
        cudaStream_t st = ctx->group_stream[k];
        Geom gg = g;
        gg.n_images = pt.count;
        uint8_t* d_px = ctx->pixels.p + (uint64_t)pt.first * img_bytes;
        uint64_t* d_off = ctx->offsets.p + pt.slice0 + k;
        CK(cudaMemcpyAsync(d_px, pixels + (uint64_t)pt.first * img_bytes, (uint64_t)pt.count * img_bytes,
                           cudaMemcpyHostToDevice, st));
        const int rc = encode_fused_on(ctx, d_px, gg, pt.first, true, ctx->payload.p + (uint64_t)pt.first * img_cap,
                                       (uint64_t)pt.count * img_cap, d_off, st, ns);
        if (rc) return rc;
        off[k] = ctx->h_qoff.p + pt.slice0 + k;
        CK(cudaMemcpyAsync(off[k], d_off, ((uint64_t)pt.count * spi + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    uint64_t pos = 0;
    int status = LLCOMP_OK;
    for (int k = 0; k < n_groups; ++k) {
        const Part& pt = parts[k];
        cudaStream_t st = ctx->group_stream[k];
        CK(cudaStreamSynchronize(st));                       // this group's offsets are on the host now
        const uint8_t* d_pay = ctx->payload.p + (uint64_t)pt.first * img_cap;
        for (int i = 0; i < pt.count && status == LLCOMP_OK; ++i) {
            const uint64_t* o = &off[k][(uint64_t)i * spi];
            const uint64_t p0 = o[0], p1 = o[spi];
            offsets[pt.first + i] = pos;
            if (p1 > (uint64_t)pt.count * img_cap || pos + hb + (p1 - p0) > out_cap) { status = LLCOMP_ERR_OVERFLOW; break; }
            write_header(g, o, out + pos);
            CK(cudaMemcpyAsync(out + pos + hb, d_pay + p0, p1 - p0, cudaMemcpyDeviceToHost, st));
            pos += hb + (p1 - p0);
        }
    }
    offsets[g.n_images] = pos;
    for (int k = 0; k < n_groups; ++k) CK(cudaStreamSynchronize(ctx->group_stream[k]));
    int dev = 0;
    CK(cudaMemcpy(&dev, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (dev != 0) { CK(cudaMemset(ctx->d_status, 0, sizeof(int))); return dev; }
    return status;
}

int llcomp_b200_encode(llcomp_ctx* ctx, const uint8_t* pixels, int width, int height, int channels, int tile_w,
                       int tile_h, uint8_t** stream, size_t* stream_len) {
    if (!stream || !stream_len) return LLCOMP_ERR_BAD_ARG;
    *stream = nullptr; *stream_len = 0;
    llcomp_geometry gi = {width, height, channels, tile_w, tile_h, 1};
    const uint64_t bound = llcomp_b200_stream_bound(&gi);
    if (!bound) return LLCOMP_ERR_BAD_ARG;
    uint8_t* buf = static_cast<uint8_t*>(malloc(bound));
    if (!buf) return LLCOMP_ERR_NOMEM;
    uint64_t off[2];
    const int rc = llcomp_b200_encode_batch(ctx, pixels, &gi, buf, bound, off);
    if (rc) { free(buf); return rc; }
    uint8_t* fit = static_cast<uint8_t*>(realloc(buf, off[1] ? off[1] : 1));
    *stream = fit ? fit : buf;
    *stream_len = off[1];
    return LLCOMP_OK;
}

int llcomp_b200_peek(const uint8_t* stream, size_t n, int* w, int* h, int* c, int* tw, int* th) {
    llcomp_geometry g; size_t hdr;
    const int rc = parse_header(stream, n, &g, &hdr, nullptr);
    if (rc) return rc;
    if (w) *w = g.width; if (h) *h = g.height; if (c) *c = g.channels;
    if (tw) *tw = g.tile_w; if (th) *th = g.tile_h;
    return LLCOMP_OK;
}

int llcomp_b200_decode_batch(llcomp_ctx* ctx, const uint8_t* streams, const uint64_t* offsets, int n_images,
                             uint8_t* pixels_out, uint64_t pixels_cap, llcomp_geometry* g_out) {
    if (!ctx || !streams || !offsets || n_images < 1 || !pixels_out) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ParsedBatch pb;
    int rc = parse_batch(streams, offsets, n_images, pb);
    if (rc) return rc;
    if (g_out) *g_out = pb.gi;
    if (pb.gi.width == 0 || pb.gi.height == 0 || pb.gi.channels == 0) return LLCOMP_OK;   // empty image: nothing to decode
    Geom g;
    if (!make_geom(&pb.gi, &g)) return LLCOMP_ERR_BAD_ARG;
    if (g.n_samples() > pixels_cap) return LLCOMP_ERR_OVERFLOW;
    return decode_parsed(ctx, streams, pb, g, 0, g.n_images, pixels_out);
}

int llcomp_b200_decode(llcomp_ctx* ctx, const uint8_t* stream, size_t n, uint8_t** pixels, int* w, int* h, int* c) {
    if (!pixels || !w || !h || !c) return LLCOMP_ERR_BAD_ARG;
    *pixels = nullptr;
    llcomp_geometry g; size_t hdr;
    int rc = parse_header(stream, n, &g, &hdr, nullptr);
    if (rc) return rc;
    const uint64_t bytes = (uint64_t)g.width * g.height * g.channels;
    uint8_t* buf = static_cast<uint8_t*>(malloc(bytes ? bytes : 1));
    if (!buf) return LLCOMP_ERR_NOMEM;
    const uint64_t off[2] = {0, n};
    rc = llcomp_b200_decode_batch(ctx, stream, off, 1, buf, bytes, &g);
    if (rc) { free(buf); return rc; }
    *pixels = buf; *w = g.width; *h = g.height; *c = g.channels;
    return LLCOMP_OK;
}

void llcomp_b200_free(void* p) { free(p); }

void* llcomp_b200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void llcomp_b200_host_free(void* p) {
    if (p) (void)cudaFreeHost(p);
}

// Model table entry s: P | next_mps<<8 | next_lps<<16 (tests compare it with the oracle's tables).
uint32_t llcomp_b200_debug_table(int s) {
    static const ModelTables t = make_tables();
    return (s >= 0 && s < 128) ? t.entry[s] : 0;
}

}  // extern "C"

// ---- several devices of one box ---------------------------------------------------------------------
// Slices are independent (own payload, own state rows, own line buffers), so a batch or a large image is dealt to
// the devices in contiguous blocks with no data-path exchange between them: one host thread and one
// context per device; the only thing that crosses devices is the slice byte counts, on the host, to place every
// device's payload in the one output.  The bytes are the single-device call's bytes.
struct llcomp_multi {
    std::vector<llcomp_ctx*> ctx;
    std::mutex mu;                          // one call at a time
};

namespace {

struct ShardPlan {
    llcomp_ctx* ctx = nullptr;
    llcomp_geometry gi{};
    const uint8_t* px = nullptr;
    int first_image = 0;
    const uint64_t* off = nullptr;          // phase 1 result: slice offsets (pinned, owned by ctx)
    int rc = LLCOMP_OK;
    struct Copy { uint64_t p0, n; uint8_t* dst; };
    std::vector<Copy> copies;               // phase 2: payload ranges -> their place in the caller's buffer
};

// Phase 1 on one device: pixels up, encode, slice offsets back.  The payload stays in ctx->payload.
int shard_encode(ShardPlan& sh) {
    llcomp_ctx* ctx = sh.ctx;
    Geom g;
    if (!make_geom(&sh.gi, &g)) return LLCOMP_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t ns = g.n_slices(), cap = payload_capacity(g);
    CK(ctx->pixels.reserve(g.n_samples()));
    CK(ctx->payload.reserve(cap));
    CK(ctx->offsets.reserve(ns + 1));
    CK(ctx->h_qoff.reserve(ns + 1));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->pixels.p, sh.px, g.n_samples(), cudaMemcpyHostToDevice, st));
    int rc = llcomp_b200_encode_device(ctx, ctx->pixels.p, &sh.gi, ctx->payload.p, cap, ctx->offsets.p, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_qoff.p, ctx->offsets.p, (ns + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    rc = llcomp_b200_finish(ctx, st);
    if (rc) return rc;
    sh.off = ctx->h_qoff.p;
    return LLCOMP_OK;
}

int shard_fetch(ShardPlan& sh) {
    llcomp_ctx* ctx = sh.ctx;
    CK(cudaSetDevice(ctx->device));
    for (const ShardPlan::Copy& c : sh.copies)
        if (c.n) CK(cudaMemcpyAsync(c.dst, ctx->payload.p + c.p0, c.n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LLCOMP_OK;
}

// Contiguous block [lo, hi) of part k out of n parts; blocks differ by at most one item.
void block_of(int items, int k, int n, int* lo, int* hi) {
    const int base = items / n, extra = items % n;
    *lo = k * base + std::min(k, extra);
    *hi = *lo + base + (k < extra ? 1 : 0);
}

// Runs fn(k) for k in [0, n) on n host threads, twice, with `between` on the calling thread in between (all first
// halves have finished when it runs; it returns false to skip the second halves).  last(k) ends every thread.
template <class F1, class Mid, class F2, class F3>
void two_phase(int n, F1 first, Mid between, F2 second, F3 last) {
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0, go = 0;                // go: 0 wait, 1 run the second half, 2 skip it
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k)
        th.emplace_back([&, k] {
            first(k);
            std::unique_lock<std::mutex> lk(m);
            if (++arrived == n) cv.notify_all();
            cv.wait(lk, [&] { return go != 0; });
            const bool run = go == 1;
            lk.unlock();
            if (run) second(k);
            last(k);                                         // on the worker thread either way
        });
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return arrived == n; });
    }
    const bool ok = between();
    {
        std::lock_guard<std::mutex> lk(m);
        go = ok ? 1 : 2;
    }
    cv.notify_all();
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int llcomp_b200_multi_create(const int* devices, int n_devices, llcomp_multi** out) {
    if (!out) return LLCOMP_ERR_BAD_ARG;
    *out = nullptr;
    if (!devices || n_devices < 1) return LLCOMP_ERR_BAD_ARG;
    llcomp_multi* m = new (std::nothrow) llcomp_multi;
    if (!m) return LLCOMP_ERR_NOMEM;
    for (int k = 0; k < n_devices; ++k) {
        llcomp_ctx* c = nullptr;
        const int rc = llcomp_b200_ctx_create(devices[k], &c);   // the same device may appear more than once
        if (rc) { llcomp_b200_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
    }
    *out = m;
    return LLCOMP_OK;
}

void llcomp_b200_multi_destroy(llcomp_multi* m) {
    if (!m) return;
    for (llcomp_ctx* c : m->ctx) llcomp_b200_ctx_destroy(c);
    delete m;
}

int llcomp_b200_multi_device_count(const llcomp_multi* m) { return m ? (int)m->ctx.size() : 0; }
llcomp_ctx* llcomp_b200_multi_ctx(llcomp_multi* m, int k) { return (m && k >= 0 && k < (int)m->ctx.size()) ? m->ctx[k] : nullptr; }

int llcomp_b200_multi_encode_batch(llcomp_multi* m, const uint8_t* pixels, const llcomp_geometry* gi, uint8_t* out,
                                   uint64_t out_cap, uint64_t* offsets) {
    Geom g;
    if (!m || !pixels || !out || !offsets || !make_geom(gi, &g)) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> call_lock(m->mu);
    const int G = (int)m->ctx.size();
    const uint32_t spi = g.slices_per_image();
    const size_t hb = header_bytes(g);
    const uint64_t img_bytes = g.image_samples();
    // a batch goes image-wise; one image goes by bands of tile rows (each band is an image of its own to its device:
    // tiles do not look outside themselves)
    const bool bands = g.n_images == 1 && g.tiles_y >= 2 && G >= 2;
    const int n_sh = bands ? std::min(G, g.tiles_y) : std::min(G, g.n_images);
    std::vector<ShardPlan> sh(n_sh);
    for (int k = 0; k < n_sh; ++k) {
        int lo, hi;
        block_of(bands ? g.tiles_y : g.n_images, k, n_sh, &lo, &hi);
        sh[k].ctx = m->ctx[k];
        sh[k].gi = *gi;
        sh[k].gi.tile_w = g.tw;
        sh[k].gi.tile_h = g.th;
        if (bands) {
            const int y0 = lo * g.th, y1 = std::min(g.H, hi * g.th);
            sh[k].gi.height = y1 - y0;
            sh[k].px = pixels + (uint64_t)y0 * g.W * g.C;
        } else {
            sh[k].gi.n_images = hi - lo;
            sh[k].first_image = lo;
            sh[k].px = pixels + (uint64_t)lo * img_bytes;
        }
    }
    int status = LLCOMP_OK;
    two_phase(
        n_sh,
        [&](int k) {
            sh[k].ctx->mu.lock();                            // held across both halves: the payload stays in the context
            sh[k].rc = shard_encode(sh[k]);
        },
        [&]() -> bool {
            for (int k = 0; k < n_sh; ++k)
                if (sh[k].rc) { status = sh[k].rc; return false; }
            uint64_t pos = 0;
            if (bands) {
                std::vector<uint64_t> all(1, 0);             // slice offsets of the whole image
                for (int k = 0; k < n_sh; ++k) {
                    Geom gk;
                    make_geom(&sh[k].gi, &gk);
                    const uint64_t nk = gk.n_slices();
                    for (uint64_t i = 1; i <= nk; ++i) all.push_back(all.back() + (sh[k].off[i] - sh[k].off[i - 1]));
                }
                if (all.size() != (size_t)spi + 1) { status = LLCOMP_ERR_BAD_ARG; return false; }
                if (hb + all.back() > out_cap) { status = LLCOMP_ERR_OVERFLOW; return false; }
                write_header(g, all.data(), out);
                pos = hb;
                for (int k = 0; k < n_sh; ++k) {
                    Geom gk;
                    make_geom(&sh[k].gi, &gk);
                    const uint64_t nb = sh[k].off[gk.n_slices()];
                    sh[k].copies.push_back({0, nb, out + pos});
                    pos += nb;
                }
                offsets[0] = 0;
                offsets[1] = pos;
                return true;
            }
            for (int k = 0; k < n_sh; ++k)
                for (int i = 0; i < sh[k].gi.n_images; ++i) {
                    const uint64_t* o = sh[k].off + (uint64_t)i * spi;
                    const uint64_t p0 = o[0], p1 = o[spi];
                    offsets[sh[k].first_image + i] = pos;
                    if (pos + hb + (p1 - p0) > out_cap) { status = LLCOMP_ERR_OVERFLOW; return false; }
                    write_header(g, o, out + pos);
                    sh[k].copies.push_back({p0, p1 - p0, out + pos + hb});
                    pos += hb + (p1 - p0);
                }
            offsets[g.n_images] = pos;
            return true;
        },
        [&](int k) { sh[k].rc = shard_fetch(sh[k]); },
        [&](int k) { sh[k].ctx->mu.unlock(); });
    for (int k = 0; k < n_sh; ++k)
        if (status == LLCOMP_OK && sh[k].rc) status = sh[k].rc;
    return status;
}

int llcomp_b200_multi_decode_batch(llcomp_multi* m, const uint8_t* streams, const uint64_t* offsets, int n_images,
                                   uint8_t* pixels_out, uint64_t pixels_cap, llcomp_geometry* g_out) {
    if (!m || !streams || !offsets || n_images < 1 || !pixels_out) return LLCOMP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> call_lock(m->mu);
    ParsedBatch pb;
    int rc = parse_batch(streams, offsets, n_images, pb);
    if (rc) return rc;
    if (g_out) *g_out = pb.gi;
    if (pb.gi.width == 0 || pb.gi.height == 0 || pb.gi.channels == 0) return LLCOMP_OK;
    Geom g;
    if (!make_geom(&pb.gi, &g)) return LLCOMP_ERR_BAD_ARG;
    if (g.n_samples() > pixels_cap) return LLCOMP_ERR_OVERFLOW;
    const int G = (int)m->ctx.size();
    const bool bands = n_images == 1 && g.tiles_y >= 2 && G >= 2;
    const int n_sh = bands ? std::min(G, g.tiles_y) : std::min(G, n_images);
    std::vector<int> rcs(n_sh, LLCOMP_OK);
    std::vector<std::thread> th;
    for (int k = 0; k < n_sh; ++k)
        th.emplace_back([&, k] {
            int lo, hi;
            block_of(bands ? g.tiles_y : n_images, k, n_sh, &lo, &hi);
            llcomp_ctx* ctx = m->ctx[k];
            std::lock_guard<std::recursive_mutex> lock(ctx->mu);
            if (!bands) {
                rcs[k] = decode_parsed(ctx, streams, pb, g, lo, hi - lo, pixels_out + (uint64_t)lo * g.image_samples());
                return;
            }
            // the band as a one-image batch of its own: tile rows [lo, hi) of the image
            const int y0 = lo * g.th, y1 = std::min(g.H, hi * g.th);
            const uint64_t sl0 = (uint64_t)lo * g.tiles_x, sl1 = (uint64_t)hi * g.tiles_x;
            ParsedBatch band;
            band.gi = pb.gi;
            band.gi.height = y1 - y0;
            band.gi.n_images = 1;
            const uint64_t b0 = pb.off[sl0], b1 = pb.off[sl1];
            for (uint64_t s = sl0; s <= sl1; ++s) band.off.push_back(pb.off[s] - b0);
            const ParsedBatch::Piece& whole = pb.pieces[0];
            const uint64_t have = whole.have > b0 ? std::min(whole.have, b1) - b0 : 0;
            band.pieces.push_back({whole.src + b0, 0, have, b1 - b0});
            Geom gb;
            if (!make_geom(&band.gi, &gb)) { rcs[k] = LLCOMP_ERR_BAD_ARG; return; }
            rcs[k] = decode_parsed(ctx, streams, band, gb, 0, 1, pixels_out + (uint64_t)y0 * g.W * g.C);
        });
    for (auto& t : th) t.join();
    for (int k = 0; k < n_sh; ++k)
        if (rcs[k]) return rcs[k];
    return LLCOMP_OK;
}

}  // extern "C"
