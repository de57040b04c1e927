// llcompc -- encoder tool; same contract as /root/reference/llcompc.cpp: `llcompc <image>` writes
// `<image>.llcomp`, exit 0 on success, 1 on a load/open failure.  Image loading is PNM/PAM instead of
// stb_image (not vendored by the reference); the codec call is the reference's own line (llcompc.cpp:33).
// Extra, optional: --tile WxH (sliced container), --device N.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "llcomp.hpp"
#include "pnm.hpp"

int main(int argc, char** argv) {
    llcomp::Options opt;
    const char* filename = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--tile") && i + 1 < argc) {
            if (std::sscanf(argv[++i], "%dx%d", &opt.tile_w, &opt.tile_h) != 2) { std::cerr << "--tile wants WxH\n"; return 1; }
        } else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) {
            opt.device = std::atoi(argv[++i]);
        } else if (!filename) {
            filename = argv[i];
        }
    }
    if (!filename) {
        std::cerr << "Usage: " << argv[0] << " <image_path> [--tile WxH] [--device N]" << std::endl;
        return 1;
    }
    pnm::Image img;
    std::string err;
    if (!pnm::read(filename, img, err)) {
        std::cerr << "Error loading image: " << err << std::endl;
        return 1;
    }
    try {
        std::vector<uint8_t> compressed = llcomp::compressImage(img.pixels, img.width, img.height, img.channels, opt);
        std::string outputFile = std::string(filename) + llcomp::ext;
        std::ofstream outFile(outputFile, std::ios::binary);
        if (!outFile) {
            std::cerr << "Error opening output file: " << outputFile << std::endl;
            return 1;
        }
        outFile.write(reinterpret_cast<const char*>(compressed.data()), (std::streamsize)compressed.size());
    } catch (const std::exception& e) {
        std::cerr << "Error compressing image: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
