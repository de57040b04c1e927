// llcompd -- decoder tool; same contract as /root/reference/llcompd.cpp: `llcompd <file.llcomp>` writes the
// decoded image next to it, exit 1 on open/decompress errors (message "Error decompressing image: <what>",
// llcompd.cpp:33), 2 on unknown exceptions.  Output is PNM/PAM (`<file>.ppm|.pgm|.pam`) instead of PNG:
// stb_image_write is not vendored by the reference.  Extra: more than one file; consecutive files with the same
// header (same size and tile grid) are decoded in one batch call.
#include <algorithm>
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
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <image_path> [more files ...]" << std::endl;
        return 1;
    }
    std::vector<std::string> files(argv + 1, argv + argc);
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
            auto [pixels, width, height, channels] = llcomp::decompressImage(streams[0]);   // llcompd.cpp:26
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
                write_image(files[k], llcomp::decompressImage(streams[k]));
            } else {
                const std::vector<std::vector<uint8_t>> group(streams.begin() + k, streams.begin() + e);
                const auto imgs = llcomp::decompressBatch(group);
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
