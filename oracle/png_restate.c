/* png_restate.c -- TEST INFRASTRUCTURE (parity checker for the INPUT boundary of the hot path).
 *
 * The reference reads every input through libpng's simplified API forced to 8-bit RGBA
 * (/root/reference/includes/utils.hpp:49-66: png_image_begin_read_from_file, format = PNG_FORMAT_RGBA,
 * png_image_finish_read) and turns bytes into floats with byte/255.0 (utils.hpp:82-84).  libpng is a system
 * dependency of the reference (CMakeLists.txt:15 find_package(PNG); vcpkg.json pins it only on Windows), not
 * vendored, and no reference test pins what it does to the 16-bit velocity fields the repository ships.  This file
 * restates that conversion from libpng's published algorithm (libpng 1.6 pngrtran.c: png_build_gamma_table,
 * png_build_16to8_table, png_gamma_16bit_correct, png_do_gamma, png_do_scale_16_to_8; png.c: png_reciprocal,
 * png_reciprocal2, png_gamma_significant, floating-point arithmetic build) for the pixel formats that occur:
 * non-interlaced, bit depth 8 or 16, colour types 0 (grey), 2 (RGB), 4 (grey+alpha), 6 (RGBA), optional gAMA / sRGB
 * chunk.  It is pinned against the real libpng 1.6.56 of this image on every input the reference ships and on
 * synthetic files covering all 65536 sample values (tests/test_png_restatement.py).
 *
 * What libpng does on that path, in the order it does it:
 *   1. reconstruct the filtered scanlines (PNG specification, filter types 0-4);
 *   2. 16-bit files are LINEAR light unless a gAMA/sRGB chunk says otherwise (file gamma 1.0), 8-bit files are
 *      sRGB (file gamma 0.45455); the output is sRGB = screen gamma 2.2;
 *   3. if the correction reciprocal2(file, screen) differs from 1 by more than 5 %, colour samples go through a
 *      table: 8-bit files a 256-entry one; 16-bit files the "16 to 8" table, which looks at the top 11 bits of the
 *      sample only (PNG_MAX_GAMMA_8 = 11) and stores the 8-bit result replicated to 16 bits;
 *   4. every 16-bit sample (alpha included, which is never gamma-corrected) is scaled to 8 bits by
 *      hi + (((lo - hi + 128) * 65535) >> 24)  ==  round(v * 255 / 65535);
 *   5. grey becomes R = G = B, a missing alpha channel becomes 255.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PNG_FP_1 100000
#define PNG_GAMMA_sRGB 220000          /* screen gamma libpng assumes for sRGB output */
#define PNG_GAMMA_sRGB_INVERSE 45455   /* file gamma libpng assumes for 8-bit input without gAMA */
#define PNG_GAMMA_LINEAR PNG_FP_1      /* file gamma libpng assumes for 16-bit input without gAMA */
#define PNG_MAX_GAMMA_8 11

/* png.c: png_reciprocal / png_reciprocal2 (floating-point build): floor(x + .5) of the quotient */
static int32_t png_reciprocal(int32_t a)
{
    double r = floor(1E10 / a + .5);
    return (r <= 2147483647. && r >= -2147483648.) ? (int32_t)r : 0;
}
static int32_t png_reciprocal2(int32_t a, int32_t b)
{
    double r = 1E15 / a;
    r /= b;
    r = floor(r + .5);
    return (r <= 2147483647. && r >= -2147483648.) ? (int32_t)r : 0;
}
/* png.c: png_gamma_significant -- PNG_GAMMA_THRESHOLD_FIXED = 5000 */
static int png_gamma_significant(int32_t g) { return g < PNG_FP_1 - 5000 || g > PNG_FP_1 + 5000; }

/* png.c: png_gamma_8bit_correct / png_gamma_16bit_correct (PNG_FLOATING_ARITHMETIC_SUPPORTED) */
static unsigned gamma_8bit_correct(unsigned value, int32_t gamma_val)
{
    if (value > 0 && value < 255) return (unsigned)floor(255 * pow((int)value / 255., gamma_val * .00001) + .5);
    return value;
}
static unsigned gamma_16bit_correct(unsigned value, int32_t gamma_val)
{
    if (value > 0 && value < 65535) return (unsigned)floor(65535. * pow((int32_t)value / 65535., gamma_val * .00001) + .5);
    return value;
}

/* pngrtran.c: png_build_16to8_table with shift = 16 - PNG_MAX_GAMMA_8 = 5, flattened: the sample's top 11 bits
 * (hi << 3 | lo >> 5) index `table` directly (libpng stores entry `last` at [last & 7][last >> 3] and looks up
 * [lo >> 5][hi]).  gamma_val is the INVERSE correction: for output value i the table finds the largest input that
 * still maps to it.  Exported so that the test can compare the table itself with libpng's output. */
void png_restate_16to8_table(uint16_t table[1 << PNG_MAX_GAMMA_8], int32_t gamma_val)
{
    const unsigned shift = 16 - PNG_MAX_GAMMA_8;
    const unsigned num = 1U << (8U - shift), max = (1U << (16U - shift)) - 1U;
    unsigned last = 0;
    for (unsigned i = 0; i < 255; ++i) {
        const unsigned out = i * 257U;                                  /* 16-bit value of 8-bit i */
        unsigned bound = gamma_16bit_correct(out + 128U, gamma_val);    /* input that maps to i + .5 */
        bound = (bound * max + 32768U) / 65535U + 1U;                   /* ... in table entries */
        while (last < bound) table[last++] = (uint16_t)out;
    }
    while (last < (num << 8)) table[last++] = 65535U;
}

/* pngrtran.c: png_do_scale_16_to_8 */
static uint8_t scale_16_to_8(unsigned v)
{
    int32_t tmp = (int32_t)(v >> 8);
    tmp += (((int32_t)(v & 0xff) - tmp + 128) * 65535) >> 24;          /* arithmetic shift, as compiled by every libpng */
    return (uint8_t)tmp;
}

/* PNG specification 9.2-9.4: reconstruct h scanlines of (1 filter byte + rowbytes) in place.  0 = ok. */
int png_restate_unfilter(uint8_t *raw, int h, size_t rowbytes, int bpp)
{
    const size_t stride = rowbytes + 1;
    uint8_t *zero = (uint8_t *)calloc(rowbytes ? rowbytes : 1, 1);
    if (!zero) return 2;
    const uint8_t *prev = zero;
    for (int j = 0; j < h; j++) {
        uint8_t *row = raw + (size_t)j * stride + 1;
        switch (row[-1]) {
        case 0: break;
        case 1:
            for (size_t i = (size_t)bpp; i < rowbytes; i++) row[i] = (uint8_t)(row[i] + row[i - bpp]);
            break;
        case 2:
            for (size_t i = 0; i < rowbytes; i++) row[i] = (uint8_t)(row[i] + prev[i]);
            break;
        case 3:
            for (size_t i = 0; i < rowbytes; i++) {
                const unsigned a = i >= (size_t)bpp ? row[i - bpp] : 0;
                row[i] = (uint8_t)(row[i] + ((a + prev[i]) >> 1));
            }
            break;
        case 4:
            for (size_t i = 0; i < rowbytes; i++) {
                const int a = i >= (size_t)bpp ? row[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                const int pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                row[i] = (uint8_t)(row[i] + pred);
            }
            break;
        default: free(zero); return 1;
        }
        prev = row;
    }
    free(zero);
    return 0;
}

/* Reconstructed scanlines -> 8-bit RGBA exactly as png_image_finish_read(format = PNG_FORMAT_RGBA) delivers them.
 * file_gamma: value of the gAMA chunk (x 100000), 45455 for an sRGB chunk, 0 if the file has neither.
 * Returns 0, or 1 for a format outside the restated set. */
int png_restate_to_rgba8(const uint8_t *rows, int w, int h, int bit_depth, int color_type, int32_t file_gamma, uint8_t *out)
{
    if ((bit_depth != 8 && bit_depth != 16) || (color_type != 0 && color_type != 2 && color_type != 4 && color_type != 6))
        return 1;
    const int channels = (color_type == 0) ? 1 : (color_type == 2) ? 3 : (color_type == 4) ? 2 : 4;
    const int colors = (color_type & 2) ? 3 : 1, has_alpha = (color_type & 4) != 0;
    const int bps = bit_depth / 8;
    const size_t stride = (size_t)w * channels * bps + 1;
    /* pngread.c png_image_read_direct: default input gamma by bit depth, output sRGB */
    if (file_gamma <= 0) file_gamma = (bit_depth == 16) ? PNG_GAMMA_LINEAR : PNG_GAMMA_sRGB_INVERSE;
    const int32_t correction = png_reciprocal2(file_gamma, PNG_GAMMA_sRGB);
    const int do_gamma = png_gamma_significant(correction);

    uint8_t t8[256];
    uint16_t *t16 = NULL;
    if (do_gamma && bit_depth == 8) {
        for (unsigned i = 0; i < 256; i++) t8[i] = (uint8_t)gamma_8bit_correct(i, correction);      /* png_build_8bit_table */
    } else if (do_gamma) {
        t16 = (uint16_t *)malloc(sizeof(uint16_t) << PNG_MAX_GAMMA_8);
        if (!t16) return 2;
        /* The inverse correction the table builder receives: reciprocal of the (already rounded) correction, for
         * linear 16-bit input reciprocal(45455) = 219998.  Pinned numerically: the libpng 1.6.53, 1.6.55 and 1.6.56
         * builds in this image all decode the 65536 sample values to the table built from 219998 (219999 gives the
         * same table; 220000 = file gamma x screen gamma would differ for the 64 samples 6560..6591, 14272..14303). */
        png_restate_16to8_table(t16, png_reciprocal(correction));
    }
    for (int j = 0; j < h; j++) {
        const uint8_t *p = rows + (size_t)j * stride + 1;
        uint8_t *o = out + (size_t)j * w * 4;
        for (int i = 0; i < w; i++, o += 4) {
            uint8_t c[4] = {0, 0, 0, 255};
            for (int k = 0; k < channels; k++, p += bps) {
                const int is_alpha = has_alpha && k == channels - 1;
                uint8_t v8;
                if (bit_depth == 8) {
                    v8 = (do_gamma && !is_alpha) ? t8[p[0]] : p[0];
                } else {
                    unsigned v = ((unsigned)p[0] << 8) | p[1];
                    if (do_gamma && !is_alpha) v = t16[v >> (16 - PNG_MAX_GAMMA_8)];                     /* png_do_gamma */
                    v8 = scale_16_to_8(v);
                }
                if (is_alpha) c[3] = v8;
                else if (colors == 1) c[0] = c[1] = c[2] = v8;                                           /* png_do_gray_to_rgb */
                else c[k] = v8;
            }
            memcpy(o, c, 4);
        }
    }
    free(t16);
    return 0;
}
