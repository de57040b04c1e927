// Test stub standing in for stb_image_write.h (included, not vendored, by the reference's llcompd.cpp).  Only
// stbi_write_png as llcompd.cpp:29 calls it; it writes a binary PNM/PAM under the name it is given.
#pragma once
#include <cstdio>

int stbi_write_png(char const* filename, int w, int h, int comp, const void* data, int stride_in_bytes);

#ifdef STB_IMAGE_WRITE_IMPLEMENTATION
int stbi_write_png(char const* filename, int w, int h, int comp, const void* data, int stride_in_bytes) {
    std::FILE* f = std::fopen(filename, "wb");
    if (!f) return 0;
    if (comp == 1) std::fprintf(f, "P5\n%d %d\n255\n", w, h);
    else if (comp == 3) std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    else std::fprintf(f, "P7\nWIDTH %d\nHEIGHT %d\nDEPTH %d\nMAXVAL 255\nTUPLTYPE %s\nENDHDR\n", w, h, comp,
                      comp == 4 ? "RGB_ALPHA" : "GRAYSCALE_ALPHA");
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (int y = 0; y < h; ++y) std::fwrite(p + (size_t)y * stride_in_bytes, 1, (size_t)w * comp, f);
    std::fclose(f);
    return 1;
}
#endif
