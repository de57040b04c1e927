// llcompc -- encoder tool; same contract as /root/reference/llcompc.cpp: `llcompc <image>` writes
// `<image>.llcomp`, exit 0 on success, 1 on a load/open failure.  Image loading is PNM/PAM instead of
// stb_image (not vendored by the reference); the codec call is the reference's own line (llcompc.cpp:33).
// Extra, optional: --tile WxH (sliced container), --device N, --gpus N (shard a batch image-wise, or one tiled image by
// bands of tile rows, over GPUs 0..N-1; the bytes do not change), and more than one image: consecutive files of the
// same size are coded in ONE batch call (one slice per image is one serial chain on the GPU, so a batch is
// where the throughput comes from); every `<image>.llcomp` is the same bytes as a single-file run writes.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "llcomp.hpp"
#include "pnm.hpp"

static bool write_stream(const std::string& image_path, const std::vector<uint8_t>& compressed) {
    const std::string outputFile = image_path + llcomp::ext;
    std::ofstream outFile(outputFile, std::ios::binary);
    if (!outFile) {
        std::cerr << "Error opening output file: " << outputFile << std::endl;
        return false;
    }
    outFile.write(reinterpret_cast<const char*>(compressed.data()), (std::streamsize)compressed.size());
    return true;
}

int main(int argc, char** argv) {
    llcomp::Options opt;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--tile") && i + 1 < argc) {
            if (std::sscanf(argv[++i], "%dx%d", &opt.tile_w, &opt.tile_h) != 2) { std::cerr << "--tile wants WxH\n"; return 1; }
        } else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) {
            opt.device = std::atoi(argv[++i]);
        } else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) {
            const int n = std::atoi(argv[++i]);
            if (n < 1) { std::cerr << "--gpus wants a positive count\n"; return 1; }
            opt.devices.clear();
            for (int d = 0; d < n && n > 1; ++d) opt.devices.push_back(d);
        } else if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) {        // explicit list, e.g. 0,0 or 2,3
            opt.devices.clear();
            for (const char* p = argv[++i]; *p;) {
                opt.devices.push_back(std::atoi(p));
                while (*p && *p != ',') ++p;
                if (*p == ',') ++p;
            }
        } else {
            files.push_back(argv[i]);
        }
    }
    if (files.empty()) {
        std::cerr << "Usage: " << argv[0] << " <image_path> [more images ...] [--tile WxH] [--device N] [--gpus N]" << std::endl;
        return 1;
    }
    std::vector<pnm::Image> imgs(files.size());
    for (size_t k = 0; k < files.size(); ++k) {
        std::string err;
        if (!pnm::read(files[k].c_str(), imgs[k], err)) {
            std::cerr << "Error loading image: " << err << std::endl;
            return 1;
        }
    }
    try {
        for (size_t k = 0; k < files.size();) {
            size_t e = k + 1;                                // run of equally sized images
            while (e < files.size() && imgs[e].width == imgs[k].width && imgs[e].height == imgs[k].height &&
                   imgs[e].channels == imgs[k].channels)
                ++e;
            if (e - k == 1) {
                if (!write_stream(files[k], llcomp::compressImage(imgs[k].pixels, imgs[k].width, imgs[k].height,
                                                                  imgs[k].channels, opt)))
                    return 1;
            } else {
                const size_t per_image = imgs[k].pixels.size();
                llcomp::detail::HostBuffer all(per_image * (e - k));     // page-locked: the upload overlaps with the coding
                for (size_t i = k; i < e; ++i) {
                    std::memcpy(all.data() + (i - k) * per_image, imgs[i].pixels.data(), per_image);
                    std::vector<uint8_t>().swap(imgs[i].pixels);          // the staged copy is the only one kept
                }
                const auto streams = llcomp::compressBatch(all.data(), all.size(), (int)(e - k), imgs[k].width,
                                                           imgs[k].height, imgs[k].channels, opt);
                for (size_t i = k; i < e; ++i)
                    if (!write_stream(files[i], streams[i - k])) return 1;
            }
            k = e;
        }
    } catch (const std::exception& e) {
        std::cerr << "Error compressing image: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
