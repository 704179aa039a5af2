/* png.h stand-in -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
 *
 * This image ships libpng 1.6 shared objects (inside the pillow wheel) but no development headers, so the
 * reference's includes/utils.hpp cannot be compiled as it is.  This header declares exactly the part of
 * libpng's "simplified API" that utils.hpp:32-150 uses, with the libpng 1.6 ABI on x86-64, so that the
 * UNMODIFIED reference driver (src/main.cpp) can be built where it lies under /root/reference and linked
 * against the real library.  Nothing in the product includes this file (the product's driver resolves
 * libpng with dlopen, host/png_io.hpp). */
#ifndef PFS_COMPAT_PNG_H
#define PFS_COMPAT_PNG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char png_byte;
typedef png_byte *png_bytep;
typedef const png_byte *png_const_bytep;
typedef uint32_t png_uint_32;
typedef int32_t png_int_32;
typedef struct png_control *png_controlp;

typedef struct {
    png_controlp opaque;
    png_uint_32 version;
    png_uint_32 width;
    png_uint_32 height;
    png_uint_32 format;
    png_uint_32 flags;
    png_uint_32 colormap_entries;
    png_uint_32 warning_or_error;
    char message[64];
} png_image, *png_imagep;

#define PNG_IMAGE_VERSION 1

#define PNG_FORMAT_FLAG_ALPHA 0x01U
#define PNG_FORMAT_FLAG_COLOR 0x02U
#define PNG_FORMAT_FLAG_LINEAR 0x04U
#define PNG_FORMAT_FLAG_COLORMAP 0x08U
#define PNG_FORMAT_RGBA (PNG_FORMAT_FLAG_COLOR | PNG_FORMAT_FLAG_ALPHA)

/* bytes of a decoded image with the default row stride */
#define PNG_IMAGE_SAMPLE_CHANNELS(fmt) (((fmt) & (PNG_FORMAT_FLAG_COLOR | PNG_FORMAT_FLAG_ALPHA)) + 1)
#define PNG_IMAGE_SAMPLE_COMPONENT_SIZE(fmt) ((((fmt) & PNG_FORMAT_FLAG_LINEAR) >> 2) + 1)
#define PNG_IMAGE_PIXEL_CHANNELS(fmt) (((fmt) & PNG_FORMAT_FLAG_COLORMAP) ? 1 : PNG_IMAGE_SAMPLE_CHANNELS(fmt))
#define PNG_IMAGE_PIXEL_COMPONENT_SIZE(fmt) (((fmt) & PNG_FORMAT_FLAG_COLORMAP) ? 1 : PNG_IMAGE_SAMPLE_COMPONENT_SIZE(fmt))
#define PNG_IMAGE_ROW_STRIDE(image) (PNG_IMAGE_PIXEL_CHANNELS((image).format) * (image).width)
#define PNG_IMAGE_BUFFER_SIZE(image, row_stride) \
    (PNG_IMAGE_PIXEL_COMPONENT_SIZE((image).format) * (image).height * (row_stride))
#define PNG_IMAGE_SIZE(image) PNG_IMAGE_BUFFER_SIZE(image, PNG_IMAGE_ROW_STRIDE(image))

int png_sig_cmp(png_const_bytep sig, size_t start, size_t num_to_check);
#define png_check_sig(sig, n) (!png_sig_cmp((sig), 0, (n)))

int png_image_begin_read_from_file(png_imagep image, const char *file_name);
int png_image_finish_read(png_imagep image, const void *background, void *buffer, png_int_32 row_stride,
                          void *colormap);
int png_image_write_to_file(png_imagep image, const char *file, int convert_to_8bit, const void *buffer,
                            png_int_32 row_stride, const void *colormap);
void png_image_free(png_imagep image);

#ifdef __cplusplus
}
#endif
#endif /* PFS_COMPAT_PNG_H */
