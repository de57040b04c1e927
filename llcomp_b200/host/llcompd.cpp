// llcompd -- decoder tool; same contract as /root/reference/llcompd.cpp: `llcompd <file.llcomp>` writes the
// decoded image next to it, exit 1 on open/decompress errors (message "Error decompressing image: <what>",
// llcompd.cpp:33), 2 on unknown exceptions.  Output is PNM/PAM (`<file>.ppm|.pgm|.pam`) instead of PNG:
// stb_image_write is not vendored by the reference.  Extra: more than one file; consecutive files with the same
// header (same size and tile grid) are decoded in one batch call; --device N / --gpus N as in llcompc.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "llcomp.hpp"
#include "pnm.hpp"

static void write_image(const std::string& stream_path, const llcomp::RawImage& img) {
    const std::string outputFile = stream_path + pnm::extension_for(img.channels);
    if (!pnm::write(outputFile, img.pixels.data(), img.width, img.height, img.channels))
        std::cerr << "Error writing output file: " << outputFile << std::endl;
}

int main(int argc, char** argv) {
    llcomp::Options opt;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--device") && i + 1 < argc) {
            opt.device = std::atoi(argv[++i]);
        } else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) {
            const int n = std::atoi(argv[++i]);
            opt.devices.clear();
            for (int d = 0; d < n && n > 1; ++d) opt.devices.push_back(d);
        } else if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) {
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
        std::cerr << "Usage: " << argv[0] << " <image_path> [more files ...] [--device N] [--gpus N]" << std::endl;
        return 1;
    }
    std::vector<std::vector<uint8_t>> streams(files.size());
    for (size_t k = 0; k < files.size(); ++k) {
        std::ifstream inFile(files[k], std::ios::binary);
        if (!inFile) {
            std::cerr << "Error opening input file: " << files[k] << std::endl;
            return 1;
        }
        streams[k].assign(std::istreambuf_iterator<char>(inFile), std::istreambuf_iterator<char>());
    }
    try {
        if (files.size() == 1) {
            auto [pixels, width, height, channels] = llcomp::decompressImage(streams[0], opt);   // llcompd.cpp:26
            write_image(files[0], llcomp::RawImage{std::move(pixels), width, height, channels});
            return 0;
        }
        // the bytes before the payloads (6 for a reference stream, 24 + 4 n for the sliced container minus the
        // per-slice lengths) say whether two streams share a geometry; comparing the fixed part is enough
        auto same_geometry = [](const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
            if (a.empty() || b.empty() || a[0] != b[0]) return false;
            const size_t fixed = a[0] == llcomp::magic_revision ? 6 : 24;
            return a.size() >= fixed && b.size() >= fixed && std::equal(a.begin(), a.begin() + fixed, b.begin());
        };
        for (size_t k = 0; k < files.size();) {
            size_t e = k + 1;
            while (e < files.size() && same_geometry(streams[k], streams[e])) ++e;
            if (e - k == 1) {
                write_image(files[k], llcomp::decompressImage(streams[k], opt));
            } else {
                const std::vector<std::vector<uint8_t>> group(streams.begin() + k, streams.begin() + e);
                const auto imgs = llcomp::decompressBatch(group, opt);
                for (size_t i = k; i < e; ++i) write_image(files[i], imgs[i - k]);
            }
            k = e;
        }
    } catch (const std::exception& e) {
        std::cerr << "Error decompressing image: " << e.what() << std::endl;
        return 1;
    } catch (...) {
        std::cerr << "Unknown error occurred" << std::endl;
        return 2;
    }
    return 0;
}
