// ref_shim.cpp -- C-callable doorway to the UNMODIFIED reference header.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into
// oracle/_ref/libllcomp_ref.so with `-I/root/reference` so that
// `#include "llcomp.hpp"` resolves to the reference's own file where it lies;
// no reference source is copied into this repository.  Used to (1) pin the C
// restatement in llcomp_oracle.c and (2) serve as the "reference" CPU baseline
// in bench.py (flags match BASELINE.md: g++ -O2 -DNDEBUG -std=c++17).
//
// The reference is undefined for streams longer than the raw image
// (llcomp.hpp:362, defect D1) and for decoding channels < 3 (:532-540, D2);
// callers must not route such inputs here (oracle/__init__.py guards this).
#include "llcomp.hpp"

#include <atomic>
#include <cstring>
#include <thread>

extern "C" {

int ref_magic(void) { return llcomp::magic_revision; }
int ref_states_nb(void) { return (int)llcomp::getStatesNb(); }
int ref_quant11(int x) { return llcomp::quant11(x); }
int ref_quant5(int x) { return llcomp::quant5(x); }
int ref_median(int a, int b, int c) { return llcomp::median(a, b, c); }
int ref_next_state_mps(int s) { return llcomp::cabac::nextStateMps[s]; }
int ref_next_state_lps(int s) { return llcomp::cabac::nextStateLps[s]; }
int ref_state_probability(int s) { return llcomp::cabac::stateProbability[s]; }

// llcomp::compressImage (llcomp.hpp:358).  Returns the stream length, or 0 if
// `cap` is too small.
size_t ref_compress(const uint8_t* px, int w, int h, int c, uint8_t* out, size_t cap) {
    std::vector<uint8_t> in(px, px + (size_t)w * h * c);
    std::vector<uint8_t> s = llcomp::compressImage(in, w, h, c);
    if (s.size() > cap) return 0;
    std::memcpy(out, s.data(), s.size());
    return s.size();
}

// llcomp::decompressImage (llcomp.hpp:461).  0 ok, 1/2 = the two exceptions.
int ref_decompress(const uint8_t* stream, size_t len, uint8_t* px_out, size_t cap,
                   int* w, int* h, int* c) {
    try {
        std::vector<uint8_t> in(stream, stream + len);
        llcomp::RawImage img = llcomp::decompressImage(in);
        *w = img.width; *h = img.height; *c = img.channels;
        if (img.pixels.size() > cap) return 3;
        std::memcpy(px_out, img.pixels.data(), img.pixels.size());
        return 0;
    } catch (const std::runtime_error& e) {
        return std::strcmp(e.what(), "Invalid magic number") == 0 ? 1 : 2;
    }
}

// Binarisation alone (llcomp.hpp:166): (ctx<<1|bit) per bin, returns count.
int ref_binarize(int diff, uint8_t* out) {
    int n = 0;
    llcomp::binarization::putSymbol<true, llcomp::param_e_lim, llcomp::param_r_lim,
                                    llcomp::param_s_bit>(diff, [&](int ctx, bool bit) {
        out[n++] = (uint8_t)((ctx << 1) | (bit ? 1 : 0));
    });
    return n;
}

// n_threads workers each code whole independent images (the reference has no
// internal threading).  Returns total stream bytes.
uint64_t ref_compress_batch_mt(const uint8_t* px, int n_images, int w, int h, int c,
                               int n_threads) {
    const size_t img = (size_t)w * h * c;
    std::atomic<int> next{0};
    std::atomic<uint64_t> total{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); ++t)
        pool.emplace_back([&] {
            for (int k; (k = next.fetch_add(1)) < n_images;) {
                std::vector<uint8_t> in(px + img * k, px + img * (k + 1));
                total += llcomp::compressImage(in, w, h, c).size();
            }
        });
    for (auto& th : pool) th.join();
    return total.load();
}

// Decode n_images copies of one stream on n_threads workers; returns pixels decoded.
uint64_t ref_decompress_batch_mt(const uint8_t* streams, const uint64_t* offsets, int n_images,
                                 int n_threads) {
    std::atomic<int> next{0};
    std::atomic<uint64_t> total{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); ++t)
        pool.emplace_back([&] {
            for (int k; (k = next.fetch_add(1)) < n_images;) {
                std::vector<uint8_t> in(streams + offsets[k], streams + offsets[k + 1]);
                total += llcomp::decompressImage(in).pixels.size();
            }
        });
    for (auto& th : pool) th.join();
    return total.load();
}

}  // extern "C"
