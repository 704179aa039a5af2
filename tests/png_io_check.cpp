// png_io_check.cpp -- drives probabilistic_fluid_simulation_b200/host/png_io.hpp (the C++ driver's PNG boundary,
// reference includes/utils.hpp:32-150) from the CPU test suite.  No CUDA: the float buffer comes from malloc.
//   png_io_check read  <in.png>  <out.bin>   -> int32 width, int32 height, then width*height*4 float32
//   png_io_check write <in.bin>  <out.png>   <- the same layout
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../probabilistic_fluid_simulation_b200/host/png_io.hpp"

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    pngio::png_image img;
    if (!std::strcmp(argv[1], "read")) {
        float *x = nullptr;
        if (pngio::read_png_to_array(&img, argv[2], &x, std::malloc) != 0) return 1;
        FILE *f = std::fopen(argv[3], "wb");
        if (!f) return 3;
        int32_t wh[2] = {(int32_t)img.width, (int32_t)img.height};
        std::fwrite(wh, sizeof(int32_t), 2, f);
        std::fwrite(x, sizeof(float), (size_t)img.width * img.height * 4, f);
        std::fclose(f);
        std::free(x);
        return 0;
    }
    if (!std::strcmp(argv[1], "write")) {
        FILE *f = std::fopen(argv[2], "rb");
        if (!f) return 3;
        int32_t wh[2];
        if (std::fread(wh, sizeof(int32_t), 2, f) != 2) return 3;
        std::vector<float> x((size_t)wh[0] * wh[1] * 4);
        if (std::fread(x.data(), sizeof(float), x.size(), f) != x.size()) return 3;
        std::fclose(f);
        std::memset(&img, 0, sizeof(img));
        img.version = pngio::kVersion;
        img.width = (uint32_t)wh[0];
        img.height = (uint32_t)wh[1];
        img.format = pngio::kFormatRGBA;
        return pngio::write_png_from_array(&img, argv[3], x.data());
    }
    return 2;
}
