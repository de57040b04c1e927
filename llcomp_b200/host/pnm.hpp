// pnm.hpp -- minimal binary PNM/PAM reader and writer for the llcompc / llcompd tools.
// Stands in for stb_image / stb_image_write (llcompc.cpp:25, llcompd.cpp:29), which the reference does not
// vendor.  P5 (gray), P6 (RGB) and P7 (PAM, any depth) with maxval 255 only: the codec is 8-bit.
#pragma once
#include <cctype>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pnm {

struct Image {
    std::vector<uint8_t> pixels;
    int width = 0, height = 0, channels = 0;
};

inline bool next_token(std::istream& in, std::string& tok) {
    tok.clear();
    int ch;
    while ((ch = in.get()) != EOF) {
        if (ch == '#') { while ((ch = in.get()) != EOF && ch != '\n') {} continue; }
        if (!std::isspace(ch)) { tok.push_back((char)ch); break; }
    }
    while ((ch = in.peek()) != EOF && !std::isspace(ch)) tok.push_back((char)in.get());
    return !tok.empty();
}

inline bool read(const std::string& path, Image& img, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open file"; return false; }
    std::string magic, tok;
    if (!next_token(in, magic)) { err = "empty file"; return false; }
    int maxval = 0;
    if (magic == "P5" || magic == "P6") {
        img.channels = magic == "P5" ? 1 : 3;
        if (!next_token(in, tok)) { err = "bad header"; return false; }
        img.width = std::atoi(tok.c_str());
        if (!next_token(in, tok)) { err = "bad header"; return false; }
        img.height = std::atoi(tok.c_str());
        if (!next_token(in, tok)) { err = "bad header"; return false; }
        maxval = std::atoi(tok.c_str());
        in.get();   // the single whitespace byte after maxval
    } else if (magic == "P7") {
        std::string line;
        std::getline(in, line);
        while (std::getline(in, line)) {
            if (line.rfind("ENDHDR", 0) == 0) break;
            std::istringstream ls(line);
            std::string key;
            ls >> key;
            if (key == "WIDTH") ls >> img.width;
            else if (key == "HEIGHT") ls >> img.height;
            else if (key == "DEPTH") ls >> img.channels;
            else if (key == "MAXVAL") ls >> maxval;
        }
    } else {
        err = "not a binary PNM/PAM file (P5, P6 or P7)";
        return false;
    }
    if (img.width <= 0 || img.height <= 0 || img.channels <= 0 || maxval != 255) {
        err = "unsupported image (need 8-bit samples, positive size)";
        return false;
    }
    img.pixels.resize((size_t)img.width * img.height * img.channels);
    in.read(reinterpret_cast<char*>(img.pixels.data()), (std::streamsize)img.pixels.size());
    if ((size_t)in.gcount() != img.pixels.size()) { err = "truncated pixel data"; return false; }
    return true;
}

inline std::string extension_for(int channels) { return channels == 1 ? ".pgm" : channels == 3 ? ".ppm" : ".pam"; }

inline bool write(const std::string& path, const uint8_t* px, int width, int height, int channels) {
    std::ofstream out(path, std::ios::binary);
    if (!out) return false;
    if (channels == 1 || channels == 3) {
        out << (channels == 1 ? "P5" : "P6") << "\n" << width << " " << height << "\n255\n";
    } else {
        const char* tt = channels == 2 ? "GRAYSCALE_ALPHA" : channels == 4 ? "RGB_ALPHA" : "MULTICHANNEL";
        out << "P7\nWIDTH " << width << "\nHEIGHT " << height << "\nDEPTH " << channels
            << "\nMAXVAL 255\nTUPLTYPE " << tt << "\nENDHDR\n";
    }
    out.write(reinterpret_cast<const char*>(px), (std::streamsize)((size_t)width * height * channels));
    return (bool)out;
}

}  // namespace pnm
