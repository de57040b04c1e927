// llcomp.hpp -- host-side mirror of the reference interface on top of the C ABI.
//
// Same namespace, names, signatures, ownership and exceptions as /root/reference/llcomp.hpp:
//   std::vector<uint8_t> llcomp::compressImage(const std::vector<uint8_t>& rgb, int width, int height, int channels)  (:358)
//   llcomp::RawImage     llcomp::decompressImage(const std::vector<uint8_t>& data)                                   (:461)
//   struct RawImage { pixels, width, height, channels }  in this member order (:454-459; llcompd.cpp:26 binds it)
//   constants ext / revision / magic_revision (:18-20)
// so llcompc.cpp:33 and llcompd.cpp:26 compile against it unchanged.  All arithmetic happens in
// libllcomp_b200.so (CUDA, sm_100a); nothing here computes a single sample and there is no CPU fallback.
//
// New (the reference has no slicing / batching / devices): Options{tile_w, tile_h, device, devices} overloads and
// compressBatch / decompressBatch.  Options::devices with two or more entries spreads a batch (image-wise) or one
// tiled image (by bands of tile rows) over several GPUs of the box; the bytes do not depend on it.
#pragma once
#include <cstdint>
#include <memory>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/llcomp_b200.h"

namespace llcomp {

constexpr inline auto ext = ".llcomp";                       // llcomp.hpp:18
constexpr inline uint8_t revision = 2;                       // llcomp.hpp:19
constexpr inline uint8_t magic_revision = 0x77 + revision;   // llcomp.hpp:20

struct RawImage {                                            // llcomp.hpp:454-459
    std::vector<uint8_t> pixels;
    uint16_t width;
    uint16_t height;
    uint8_t channels;
};

struct Options {
    int tile_w = 0, tile_h = 0;   // 0 = one slice per image: byte-identical to the reference stream
    int device = 0;
    std::vector<int> devices;     // >= 2 entries: shard over these GPUs (llcomp_b200_multi_*); else `device` alone
};

namespace detail {
struct CtxDeleter { void operator()(llcomp_ctx* c) const { llcomp_b200_ctx_destroy(c); } };

inline llcomp_ctx* context(int device) {
    static std::mutex mu;
    static std::vector<std::unique_ptr<llcomp_ctx, CtxDeleter>> ctxs;
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0) throw std::invalid_argument("llcomp: negative device index");
    if ((size_t)device >= ctxs.size()) ctxs.resize(device + 1);
    if (!ctxs[device]) {
        llcomp_ctx* c = nullptr;
        if (llcomp_b200_ctx_create(device, &c) != LLCOMP_OK)
            throw std::runtime_error("llcomp: no usable CUDA device (this build has no CPU path)");
        ctxs[device].reset(c);
    }
    return ctxs[device].get();
}

struct MultiDeleter { void operator()(llcomp_multi* m) const { llcomp_b200_multi_destroy(m); } };

inline llcomp_multi* multi_context(const std::vector<int>& devices) {
    static std::mutex mu;
    static std::vector<std::pair<std::vector<int>, std::unique_ptr<llcomp_multi, MultiDeleter>>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache)
        if (e.first == devices) return e.second.get();
    llcomp_multi* m = nullptr;
    if (llcomp_b200_multi_create(devices.data(), (int)devices.size(), &m) != LLCOMP_OK)
        throw std::runtime_error("llcomp: no usable CUDA device (this build has no CPU path)");
    cache.emplace_back(devices, std::unique_ptr<llcomp_multi, MultiDeleter>(m));
    return m;
}

// Staging memory of the batch calls: page-locked when it can be had (llcomp_b200_host_alloc), so that the library's
// copies run asynchronously and overlap with the coding of other image groups; ordinary memory otherwise.
struct HostBuffer {
    uint8_t* p = nullptr;
    size_t n = 0;
    bool pinned = false;
    explicit HostBuffer(size_t bytes) : n(bytes) {
        if (bytes == 0) return;
        p = static_cast<uint8_t*>(llcomp_b200_host_alloc(bytes));
        pinned = p != nullptr;
        if (!p) p = static_cast<uint8_t*>(std::malloc(bytes));
        if (!p) throw std::bad_alloc();
    }
    ~HostBuffer() {
        if (pinned) llcomp_b200_host_free(p);
        else std::free(p);
    }
    HostBuffer(const HostBuffer&) = delete;
    HostBuffer& operator=(const HostBuffer&) = delete;
    uint8_t* data() { return p; }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

// RawImage carries 16-bit dimensions (llcomp.hpp:454-459); the reference truncates larger ones silently
// (SURVEY.md defect D3), this refuses them.
inline void check_u16(int w, int h) {
    if (w > 0xFFFF || h > 0xFFFF)
        throw std::length_error("llcomp: image dimensions above 65535 do not fit llcomp::RawImage; use the C ABI");
}

// The two reference exceptions keep their exact text (llcomp.hpp:233, :466) so callers that print
// e.what() (llcompd.cpp:33) behave the same.
[[noreturn]] inline void raise(llcomp_ctx* c, int status) {
    std::string msg = llcomp_b200_status_string(status);
    if (status == LLCOMP_ERR_CUDA) msg += std::string(": ") + llcomp_b200_last_error(c);
    throw std::runtime_error(msg);
}
}  // namespace detail

inline std::vector<uint8_t> compressImage(const std::vector<uint8_t>& rgb, int width, int height, int channels,
                                          const Options& opt) {
    if (rgb.size() != (size_t)width * height * channels)     // assert at llcomp.hpp:361
        throw std::invalid_argument("llcomp: rgb.size() != width*height*channels");
    if (opt.devices.size() >= 2) {                           // bands of tile rows over several GPUs
        llcomp_multi* m = detail::multi_context(opt.devices);
        llcomp_geometry g{width, height, channels, opt.tile_w, opt.tile_h, 1};
        std::vector<uint8_t> out(llcomp_b200_stream_bound(&g));
        uint64_t off[2] = {0, 0};
        const int rc = llcomp_b200_multi_encode_batch(m, rgb.data(), &g, out.data(), out.size(), off);
        if (rc != LLCOMP_OK) detail::raise(llcomp_b200_multi_ctx(m, 0), rc);
        out.resize(off[1]);
        return out;
    }
    llcomp_ctx* c = detail::context(opt.device);
    uint8_t* s = nullptr;
    size_t n = 0;
    const int rc = llcomp_b200_encode(c, rgb.data(), width, height, channels, opt.tile_w, opt.tile_h, &s, &n);
    if (rc != LLCOMP_OK) detail::raise(c, rc);
    std::vector<uint8_t> out(s, s + n);
    llcomp_b200_free(s);
    return out;
}

inline std::vector<uint8_t> compressImage(const std::vector<uint8_t>& rgb, int width, int height, int channels) {
    return compressImage(rgb, width, height, channels, Options{});
}

inline RawImage decompressImage(const std::vector<uint8_t>& data, const Options& opt) {
    llcomp_ctx* c = detail::context(opt.devices.size() >= 2 ? opt.devices[0] : opt.device);
    {
        int w = 0, h = 0, ch = 0, tw = 0, th = 0;
        const int rc = llcomp_b200_peek(data.data(), data.size(), &w, &h, &ch, &tw, &th);
        if (rc != LLCOMP_OK) detail::raise(c, rc);
        detail::check_u16(w, h);
        if (opt.devices.size() >= 2) {
            llcomp_multi* m = detail::multi_context(opt.devices);
            RawImage img{std::vector<uint8_t>((size_t)w * h * ch), (uint16_t)w, (uint16_t)h, (uint8_t)ch};
            const uint64_t off[2] = {0, data.size()};
            llcomp_geometry g{};
            const int rc2 = llcomp_b200_multi_decode_batch(m, data.data(), off, 1, img.pixels.data(), img.pixels.size(), &g);
            if (rc2 != LLCOMP_OK) detail::raise(llcomp_b200_multi_ctx(m, 0), rc2);
            return img;
        }
    }
    uint8_t* px = nullptr;
    int w = 0, h = 0, ch = 0;
    const int rc = llcomp_b200_decode(c, data.data(), data.size(), &px, &w, &h, &ch);
    if (rc != LLCOMP_OK) detail::raise(c, rc);
    RawImage img{std::vector<uint8_t>(px, px + (size_t)w * h * ch), (uint16_t)w, (uint16_t)h, (uint8_t)ch};
    llcomp_b200_free(px);
    return img;
}

inline RawImage decompressImage(const std::vector<uint8_t>& data) { return decompressImage(data, Options{}); }

// n equally sized images, pixels back to back -> one complete stream per image.  `pixels` may be any host memory; the
// tools stage theirs in a detail::HostBuffer (page-locked).  The streams come back through a page-locked buffer.
inline std::vector<std::vector<uint8_t>> compressBatch(const uint8_t* pixels, size_t n_bytes, int n_images, int width,
                                                       int height, int channels, const Options& opt = Options{}) {
    llcomp_geometry g{width, height, channels, opt.tile_w, opt.tile_h, n_images};
    if (n_bytes != llcomp_b200_sample_count(&g)) throw std::invalid_argument("llcomp: batch size mismatch");
    detail::HostBuffer buf(llcomp_b200_stream_bound(&g));
    std::vector<uint64_t> off(n_images + 1);
    if (opt.devices.size() >= 2) {
        llcomp_multi* m = detail::multi_context(opt.devices);
        const int rc = llcomp_b200_multi_encode_batch(m, pixels, &g, buf.data(), buf.size(), off.data());
        if (rc != LLCOMP_OK) detail::raise(llcomp_b200_multi_ctx(m, 0), rc);
    } else {
        llcomp_ctx* c = detail::context(opt.device);
        const int rc = llcomp_b200_encode_batch(c, pixels, &g, buf.data(), buf.size(), off.data());
        if (rc != LLCOMP_OK) detail::raise(c, rc);
    }
    std::vector<std::vector<uint8_t>> out(n_images);
    for (int k = 0; k < n_images; ++k) out[k].assign(buf.data() + off[k], buf.data() + off[k + 1]);
    return out;
}
inline std::vector<std::vector<uint8_t>> compressBatch(const std::vector<uint8_t>& pixels, int n_images, int width,
                                                       int height, int channels, const Options& opt = Options{}) {
    return compressBatch(pixels.data(), pixels.size(), n_images, width, height, channels, opt);
}

// Inverse: n complete streams of the same geometry (same header, same tile grid) -> n images in one call.
inline std::vector<RawImage> decompressBatch(const std::vector<std::vector<uint8_t>>& streams,
                                             const Options& opt = Options{}) {
    std::vector<RawImage> out;
    if (streams.empty()) return out;
    int w = 0, h = 0, ch = 0, tw = 0, th = 0;
    llcomp_ctx* c = detail::context(opt.devices.size() >= 2 ? opt.devices[0] : opt.device);
    int rc = llcomp_b200_peek(streams[0].data(), streams[0].size(), &w, &h, &ch, &tw, &th);
    if (rc != LLCOMP_OK) detail::raise(c, rc);
    detail::check_u16(w, h);
    std::vector<uint64_t> off(streams.size() + 1, 0);
    for (size_t k = 0; k < streams.size(); ++k) off[k + 1] = off[k] + streams[k].size();
    detail::HostBuffer cat(off.back());                      // page-locked staging: streams in, pixels out
    for (size_t k = 0; k < streams.size(); ++k) std::memcpy(cat.data() + off[k], streams[k].data(), streams[k].size());
    const size_t per_image = (size_t)w * h * ch;
    detail::HostBuffer px(per_image * streams.size());
    llcomp_geometry g{};
    if (opt.devices.size() >= 2)
        rc = llcomp_b200_multi_decode_batch(detail::multi_context(opt.devices), cat.data(), off.data(), (int)streams.size(),
                                            px.data(), px.size(), &g);
    else
        rc = llcomp_b200_decode_batch(c, cat.data(), off.data(), (int)streams.size(), px.data(), px.size(), &g);
    if (rc != LLCOMP_OK) detail::raise(c, rc);
    out.reserve(streams.size());
    for (size_t k = 0; k < streams.size(); ++k)
        out.push_back(RawImage{std::vector<uint8_t>(px.data() + k * per_image, px.data() + (k + 1) * per_image),
                               (uint16_t)w, (uint16_t)h, (uint8_t)ch});
    return out;
}

}  // namespace llcomp
