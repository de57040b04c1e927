// Test stub standing in for stb_image.h, which the reference's llcompc.cpp includes but does not vendor
// (CMakeLists.txt:9 finds it through vcpkg).  Only what llcompc.cpp:25-31 calls, reading binary PNM (P5/P6) instead
// of PNG/JPEG: enough to compile, link and run the UNMODIFIED tool against llcomp_b200/host/llcomp.hpp.
#pragma once
#include <cstdio>
#include <cstdlib>

typedef unsigned char stbi_uc;
stbi_uc* stbi_load(char const* filename, int* x, int* y, int* channels_in_file, int desired_channels);
const char* stbi_failure_reason(void);
void stbi_image_free(void* retval_from_stbi_load);

#ifdef STB_IMAGE_IMPLEMENTATION
static const char* stub_stbi_reason = "";
const char* stbi_failure_reason(void) { return stub_stbi_reason; }
void stbi_image_free(void* p) { std::free(p); }
stbi_uc* stbi_load(char const* filename, int* x, int* y, int* comp, int) {
    std::FILE* f = std::fopen(filename, "rb");
    if (!f) { stub_stbi_reason = "can't fopen"; return nullptr; }
    char magic[3] = {0, 0, 0};
    int maxv = 0;
    if (std::fscanf(f, "%2s %d %d %d", magic, x, y, &maxv) != 4 || magic[0] != 'P' || (magic[1] != '5' && magic[1] != '6') ||
        maxv != 255 || *x < 1 || *y < 1) {
        std::fclose(f);
        stub_stbi_reason = "not a binary PGM/PPM";
        return nullptr;
    }
    std::fgetc(f);                                            // the single whitespace after maxval
    *comp = magic[1] == '6' ? 3 : 1;
    const size_t n = (size_t)*x * *y * *comp;
    stbi_uc* px = static_cast<stbi_uc*>(std::malloc(n));
    if (!px || std::fread(px, 1, n, f) != n) {
        std::free(px);
        std::fclose(f);
        stub_stbi_reason = "short file";
        return nullptr;
    }
    std::fclose(f);
    return px;
}
#endif
