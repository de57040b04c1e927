// llcompd -- decoder tool; same contract as /root/reference/llcompd.cpp: `llcompd <file.llcomp>` writes the
// decoded image next to it, exit 1 on open/decompress errors (message "Error decompressing image: <what>",
// llcompd.cpp:33), 2 on unknown exceptions.  Output is PNM/PAM (`<file>.ppm|.pgm|.pam`) instead of PNG:
// stb_image_write is not vendored by the reference.
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "llcomp.hpp"
#include "pnm.hpp"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <image_path>" << std::endl;
        return 1;
    }
    const char* filename = argv[1];
    std::ifstream inFile(filename, std::ios::binary);
    if (!inFile) {
        std::cerr << "Error opening input file: " << filename << std::endl;
        return 1;
    }
    std::vector<uint8_t> compressed((std::istreambuf_iterator<char>(inFile)), std::istreambuf_iterator<char>());
    inFile.close();
    try {
        auto [pixels, width, height, channels] = llcomp::decompressImage(compressed);   // llcompd.cpp:26
        std::string outputFile = std::string(filename) + pnm::extension_for(channels);
        if (!pnm::write(outputFile, pixels.data(), width, height, channels)) {
            std::cerr << "Error writing output file: " << outputFile << std::endl;
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
