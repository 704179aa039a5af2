// png_io.hpp -- PNG <-> float buffers with the reference's exact semantics (includes/utils.hpp:32-150):
// libpng's simplified API forced to 8-bit RGBA on read (so 16-bit velocity fields get libpng's own
// gamma handling), float = byte/255.0, and byte = (png_byte)(x*255.0) (truncation) on
// write.  libpng has no development headers in this image, so the five entry points are resolved
// with dlopen() and the png_image struct is declared by hand (libpng 1.6 ABI, x86-64).
#pragma once

#include <dlfcn.h>
#include <glob.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace pngio {

struct png_image {          // libpng 1.6 png.h: struct png_image
    void *opaque;
    uint32_t version, width, height, format, flags, colormap_entries, warning_or_error;
    char message[64];
};
constexpr uint32_t kVersion = 1;      // PNG_IMAGE_VERSION
constexpr uint32_t kFormatRGBA = 0x03;  // PNG_FORMAT_RGBA = COLOR | ALPHA, 8 bits per channel

struct Api {
    int (*begin_read_from_file)(png_image *, const char *) = nullptr;
    int (*finish_read)(png_image *, const void *, void *, int32_t, void *) = nullptr;
    int (*write_to_file)(png_image *, const char *, int, const void *, int32_t, const void *) = nullptr;
    void (*image_free)(png_image *) = nullptr;
    int (*sig_cmp)(const unsigned char *, size_t, size_t) = nullptr;
    bool ok() const { return begin_read_from_file && finish_read && write_to_file && image_free && sig_cmp; }
};

inline const Api &api()
{
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    std::vector<std::string> cands;
    if (const char *e = getenv("PFS_LIBPNG")) cands.push_back(e);
    cands.push_back("libpng16.so.16");
    cands.push_back("libpng16.so");
#ifdef PFS_LIBPNG_HINT
    // a glob pattern found by the build (host/Makefile looks for the libpng bundled with Pillow when the system has none)
    {
        glob_t g;
        if (glob(PFS_LIBPNG_HINT, 0, nullptr, &g) == 0)
            for (size_t i = 0; i < g.gl_pathc; i++) cands.push_back(g.gl_pathv[i]);
        globfree(&g);
    }
#endif
    for (const auto &c : cands) {
        void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) continue;
        a.begin_read_from_file = (decltype(a.begin_read_from_file))dlsym(h, "png_image_begin_read_from_file");
        a.finish_read = (decltype(a.finish_read))dlsym(h, "png_image_finish_read");
        a.write_to_file = (decltype(a.write_to_file))dlsym(h, "png_image_write_to_file");
        a.image_free = (decltype(a.image_free))dlsym(h, "png_image_free");
        a.sig_cmp = (decltype(a.sig_cmp))dlsym(h, "png_sig_cmp");
        if (a.ok()) break;
        a = Api();
    }
    return a;
}

// utils.hpp:32-108.  `alloc` provides the float buffer (pinned host memory in the CUDA driver).
inline int read_png_to_array(png_image *image, const char *fn, float **x, void *(*alloc)(size_t))
{
    const Api &p = api();
    if (!p.ok()) {
        std::fprintf(stderr, "libpng16 not found (set PFS_LIBPNG)\n");
        return 1;
    }
    FILE *fp = std::fopen(fn, "rb");
    if (!fp) return 1;
    unsigned char sig[8] = {0};
    size_t got = std::fread(sig, 1, 8, fp);
    std::fclose(fp);
    if (got != 8 || p.sig_cmp(sig, 0, 8) != 0) return 1;    // png_check_sig(sig, 8)
    std::memset(image, 0, sizeof(*image));
    image->version = kVersion;
    if (!p.begin_read_from_file(image, fn)) return 1;
    image->format = kFormatRGBA;
    const size_t n = (size_t)image->height * image->width * 4;  // PNG_IMAGE_SIZE for RGBA8, default stride
    std::vector<unsigned char> buffer(n);
    if (!p.finish_read(image, nullptr, buffer.data(), 0, nullptr)) return 1;
    *x = (float *)alloc(sizeof(float) * n);
    if (!*x) return 1;
    for (size_t i = 0; i < n; i++) (*x)[i] = (float)buffer[i] / 255.0;   // utils.hpp:83 (double divide)
    return 0;
}

// utils.hpp:120-150
inline int write_png_from_array(png_image *image, const char *fn, const float *x)
{
    const Api &p = api();
    const size_t n = (size_t)image->height * image->width * 4;
    std::vector<unsigned char> buffer(n);
    for (size_t i = 0; i < n; i++) buffer[i] = (unsigned char)(x[i] * 255.0);   // truncation, utils.hpp:130
    return p.write_to_file(image, fn, 0, buffer.data(), 0, nullptr) ? 0 : 1;
}

// Same file as write_png_from_array would produce, from bytes already converted on the device.
inline int write_png_from_bytes(png_image *image, const char *fn, const unsigned char *rgba)
{
    return api().write_to_file(image, fn, 0, rgba, 0, nullptr) ? 0 : 1;
}

}  // namespace pngio
