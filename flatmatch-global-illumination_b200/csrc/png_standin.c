/* zlib-based stand-in for the reference's two libpng entry points (png_helper.h:13-15,
 * implemented there by png_helper.c:111-161 and :255-287 on top of libpng).
 *
 * The hot path does not touch PNG files; this exists so that the reference's untouched main.c /
 * image.c / rectangle.c can be linked into a complete `globalIllumination` binary on machines
 * without libpng headers (this image), which is what the "example.png bake wall time" half of the
 * BASELINE metric needs: layout PNG in, tiles/tile_N.png out.
 *
 * Same contract as the reference functions: 8-bit RGB (colour type 2) or RGBA (6), non-interlaced;
 * read returns one malloc'ed, row-contiguous buffer (png_helper.c:61-66); an unreadable file
 * prints and exit(0)s (png_helper.c:118-123).  Decoded pixels are identical to libpng's; encoded
 * files differ from libpng's byte for byte (filter / deflate choices) but decode to the same pixels.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static const int PNG_STANDIN_RGB = 2, PNG_STANDIN_RGBA = 6;

static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void put32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

void read_png_file(const char *file_name, int *width, int *height, int *color_type, uint8_t **pixel_buffer)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    FILE *fp = fopen(file_name, "rb");
    if (!fp) {
        printf("File '%s' could not be opened, exiting ...\n", file_name);
        exit(0);
    }
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint8_t *file = (uint8_t *)malloc(size > 0 ? (size_t)size : 1);
    if (size < 8 || fread(file, 1, (size_t)size, fp) != (size_t)size) {
        printf("Error reading file header, exiting ...\n");
        fclose(fp); free(file);
        return;
    }
    fclose(fp);
    if (memcmp(file, sig, 8) != 0) { free(file); return; }

    uint32_t w = 0, h = 0;
    int ctype = -1, channels = 0;
    uint8_t *idat = (uint8_t *)malloc((size_t)size);
    size_t idat_len = 0;
    for (long pos = 8; pos + 12 <= size;) {
        uint32_t len = be32(file + pos);
        const uint8_t *type = file + pos + 4, *data = file + pos + 8;
        if (pos + 12 + (long)len > size) break;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data); h = be32(data + 4);
            ctype = data[9];
            if (data[8] != 8 || (ctype != PNG_STANDIN_RGB && ctype != PNG_STANDIN_RGBA) || data[12] != 0) {
                printf("[Err] '%s': only 8-bit non-interlaced RGB/RGBA PNG files are supported, exiting ...\n", file_name);
                exit(0);
            }
            channels = ctype == PNG_STANDIN_RGB ? 3 : 4;
        } else if (!memcmp(type, "IDAT", 4)) {
            memcpy(idat + idat_len, data, len);
            idat_len += len;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (long)len;
    }
    if (!channels || !w || !h) { free(file); free(idat); return; }

    const size_t stride = (size_t)w * channels;
    uLongf raw_len = (uLongf)((stride + 1) * h);
    uint8_t *raw = (uint8_t *)malloc(raw_len);
    if (uncompress(raw, &raw_len, idat, (uLong)idat_len) != Z_OK || raw_len != (stride + 1) * h) {
        printf("[Err] '%s': corrupt image data, exiting ...\n", file_name);
        exit(0);
    }
    uint8_t *out = (uint8_t *)malloc(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *src = raw + (stride + 1) * y;
        uint8_t *dst = out + stride * y;
        const uint8_t *up = y ? dst - stride : NULL;
        const int filter = src[0];
        src++;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)channels ? dst[x - channels] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)channels) ? up[x - channels] : 0;
            int v = src[x];
            switch (filter) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
            }
            dst[x] = (uint8_t)v;
        }
    }
    free(raw); free(idat); free(file);
    *width = (int)w; *height = (int)h; *color_type = ctype; *pixel_buffer = out;
}

static void write_chunk(FILE *fp, const char *type, const uint8_t *data, uint32_t len)
{
    uint8_t hdr[8], crcb[4];
    put32(hdr, len);
    memcpy(hdr + 4, type, 4);
    uint32_t crc = (uint32_t)crc32(0L, hdr + 4, 4);
    if (len) crc = (uint32_t)crc32(crc, data, len);
    put32(crcb, crc);
    fwrite(hdr, 1, 8, fp);
    if (len) fwrite(data, 1, len, fp);
    fwrite(crcb, 1, 4, fp);
}

void write_png_file(const char *file_name, int width, int height, int color_type, uint8_t *pixel_buffer)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (width <= 0 || height <= 0 || (color_type != PNG_STANDIN_RGB && color_type != PNG_STANDIN_RGBA)) return;
    FILE *fp = fopen(file_name, "wb");
    if (!fp) return;                                   /* png_helper.c:261-262 */
    const int channels = color_type == PNG_STANDIN_RGB ? 3 : 4;
    const size_t stride = (size_t)width * channels;
    uint8_t *raw = (uint8_t *)malloc((stride + 1) * (size_t)height);
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;                     /* filter type 0 (none) */
        memcpy(raw + (stride + 1) * y + 1, pixel_buffer + stride * y, stride);
    }
    uLongf zlen = compressBound((uLong)((stride + 1) * height));
    uint8_t *z = (uint8_t *)malloc(zlen);
    if (compress2(z, &zlen, raw, (uLong)((stride + 1) * height), Z_BEST_SPEED) != Z_OK) {
        free(raw); free(z); fclose(fp);
        return;
    }
    uint8_t ihdr[13];
    put32(ihdr, (uint32_t)width); put32(ihdr + 4, (uint32_t)height);
    ihdr[8] = 8; ihdr[9] = (uint8_t)color_type; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    fwrite(sig, 1, 8, fp);
    write_chunk(fp, "IHDR", ihdr, 13);
    write_chunk(fp, "IDAT", z, (uint32_t)zlen);
    write_chunk(fp, "IEND", NULL, 0);
    fclose(fp);
    free(raw); free(z);
}
